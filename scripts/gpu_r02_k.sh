#!/bin/bash
# Round-2 closing visit (one GPU) on the final kernels (lockstep + 64 KB stages): full parity suite, smoke, bench lines
# of every workload (verify / cpu legs / clustered leg), reference arm, drop-in latency, launch lists, and the ncu
# --set full capture of the batch-2048 scan on the bench's own routed federation.  Lands in gpurun_out/r02k_*.
mkdir -p gpurun_out
bash scripts/gpu_tests_split.sh
timeout 300 python -m pytest tests/test_gpu_glue.py tests/test_gpu_multi.py -q -p no:cacheprovider > gpurun_out/glue.log 2>&1; echo "glue+multi exit=$? $(tail -1 gpurun_out/glue.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? $(tail -1 gpurun_out/smoke.log)"
timeout 900 python bench.py > gpurun_out/r02k_bench_federation_1gpu.json 2> gpurun_out/r02k_bench_federation_1gpu.err; echo "bench exit=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02k_bench_reference_arm.json 2> gpurun_out/r02k_bench_reference_arm.err; echo "reference arm exit=$?"
for wl in textbooks medqa_ts; do
  timeout 600 python bench.py --workload $wl --batch 64 --steps 200 --warmup 20 > gpurun_out/r02k_bench_$wl.json 2> gpurun_out/r02k_bench_$wl.err; echo "$wl exit=$?"
done
python - <<'PY'
import json
for f in ('federation_1gpu','textbooks','medqa_ts'):
    try:
        d=json.load(open(f'gpurun_out/r02k_bench_{f}.json')); r=d['roofline']
        print(f, {k:d.get(k) for k in ('value','ms_per_step','whole_step_frac','exact_rescans')}, 'verify', d['verify']['mismatches'], 'e2e', round(d['e2e']['value'],1), round(d['e2e']['p50_ms'],4), 'roofline', round(r['frac'],3), d['clocks']['sm_mhz'])
        print('   ', [(s['batch'], round(s['p50_ms'],4), round(s['frac'],3), round(s['whole_step_frac'],3), s['exact_rescans'], s['verify']['mismatches']) for s in d.get('sweep',[])])
        c=d.get('clustered'); cb=d.get('cpu_baseline')
        if c: print('    clustered', round(c['qps'],1), c['exact_rescans_per_step'], c['verify']['mismatches'])
        if cb: print('    cpu', round(cb['value'],2), round(cb['as_deployed']['value'],3), cb['cores'])
    except Exception as e: print(f,'parse failed',e)
try: print(open('gpurun_out/r02k_bench_reference_arm.json').read()[:300])
except Exception as e: print(e)
PY
timeout 300 python scripts/gpu_dropin_latency.py > gpurun_out/r02k_dropin_latency.jsonl 2> gpurun_out/r02k_dropin.err; echo "dropin exit=$?"; cut -c1-230 gpurun_out/r02k_dropin_latency.jsonl
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-clustered-leg --verify-queries 0 --sweep=8,64"
timeout 200 $CMD > gpurun_out/r02k_plain.json 2> gpurun_out/r02k_plain.err &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:rr_[^s]' -c 400 --csv --log-file gpurun_out/r02k_launches.csv $CMD > gpurun_out/r02k_ncu_l.out 2>&1
echo "launch list exit=$?"
CMD2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-clustered-leg --verify-queries 0 --sweep="
timeout 200 $CMD2 > gpurun_out/r02k_plain2.json 2> gpurun_out/r02k_plain2.err &&
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:rr_gemm_scan_kernel<\(int\)2, \(bool\)0' -s 7 -c 1 -f -o gpurun_out/r02k_gemm_routed $CMD2 > gpurun_out/r02k_ncu2.out 2>&1; echo "gemm routed capture exit=$?"
ncu -i gpurun_out/r02k_gemm_routed.ncu-rep --page raw --csv > gpurun_out/r02k_gemm_routed_raw.csv 2>/dev/null
python scripts/ncu_keys.py gpurun_out/r02k_gemm_routed.ncu-rep 2>/dev/null | grep -E "Kernel Name|dram__bytes|time_duration|tensor_cycles|hit_rate"
