#!/bin/bash
# Round-2 visit H (one GPU): the lockstep of the units sharing a corpus slice (rr_scan_gemm.cu).
# Full parity suite on the new build, then A/B of the lockstep window on the bench's own workloads, then the
# ncu --set full capture of the routed-federation launch whose DRAM traffic was 2.63x the algorithmic bytes.
mkdir -p gpurun_out
bash scripts/gpu_tests_split.sh
timeout 300 python -m pytest tests/test_gpu_glue.py tests/test_gpu_multi.py -q -p no:cacheprovider > gpurun_out/glue.log 2>&1; echo "glue+multi exit=$? $(tail -1 gpurun_out/glue.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? $(tail -1 gpurun_out/smoke.log)"
AB="--no-cpu-baseline --no-clustered-leg --steps 20 --warmup 5 --verify-queries 2"
for w in 0 4 1; do
  RR_LOCKSTEP=$w timeout 300 python bench.py --sweep= $AB > gpurun_out/r02h_fed_w$w.json 2> gpurun_out/r02h_fed_w$w.err; echo "federation window $w exit=$?"
done
for w in 0 4; do
  RR_LOCKSTEP=$w timeout 300 python bench.py --workload pubmed --sweep=1024,4096 $AB > gpurun_out/r02h_pubmed_w$w.json 2> gpurun_out/r02h_pubmed_w$w.err; echo "pubmed window $w exit=$?"
done
python - <<'PY'
import json
for f in ('fed_w0','fed_w4','fed_w1','pubmed_w0','pubmed_w4'):
    try:
        d=json.load(open(f'gpurun_out/r02h_{f}.json')); r=d['roofline']
        print(f, round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(r['frac'],4), 'whole', round(d['whole_step_frac'],4), 'verify', d['verify']['mismatches'], 'rescans', d['exact_rescans'], 'clk', d['clocks']['sm_mhz'],
              [(s['batch'], round(s['p50_ms'],3), round(s['frac'],3), s['verify']['mismatches']) for s in d.get('sweep',[])])
    except Exception as e: print(f,'parse failed',e)
PY
# the default bench line of the new build (everything on)
timeout 900 python bench.py > gpurun_out/r02h_bench_federation_1gpu.json 2> gpurun_out/r02h_bench_federation_1gpu.err; echo "bench exit=$?"
CMD2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-clustered-leg --verify-queries 0 --sweep="
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:rr_gemm_scan_kernel<\(int\)2, \(bool\)0' -s 7 -c 1 -f -o gpurun_out/r02h_gemm_routed $CMD2 > gpurun_out/r02h_ncu2.out 2>&1; echo "gemm routed capture exit=$?"
ncu -i gpurun_out/r02h_gemm_routed.ncu-rep --page raw --csv > gpurun_out/r02h_gemm_routed_raw.csv 2>/dev/null
python scripts/ncu_keys.py gpurun_out/r02h_gemm_routed.ncu-rep 2>/dev/null | head -40
