#!/bin/bash
# Round-2 visit F (one GPU): validation of the adaptive rescoring set + sort-based merge, small-corpus benches,
# default bench, and the ncu evidence on the bench's OWN routed configuration (launch list + --set full captures).
mkdir -p gpurun_out
bash scripts/gpu_tests_split.sh
timeout 300 python -m pytest tests/test_gpu_glue.py -q -p no:cacheprovider > gpurun_out/glue.log 2>&1; echo "glue exit=$? $(tail -1 gpurun_out/glue.log)"
for wl in textbooks medqa_ts; do
  timeout 600 python bench.py --workload $wl --batch 64 --steps 200 --warmup 20 > gpurun_out/r02f_bench_$wl.json 2> gpurun_out/r02f_bench_$wl.err; echo "$wl exit=$?"
done
python - <<'PY'
import json
for wl in ('textbooks','medqa_ts'):
    try:
        d=json.load(open(f'gpurun_out/r02f_bench_{wl}.json'))
        print(wl, 'e2e p50', round(d['e2e']['p50_ms'],4), [(s['batch'], round(s['p50_ms'],4), round(s['whole_step_frac'],3), s['exact_rescans'], s['verify']['mismatches']) for s in d['sweep']], 'cpu', round(d['cpu_baseline']['value'],1), round(d['cpu_baseline']['as_deployed']['value'],2))
    except Exception as e: print(wl,'parse failed',e)
PY
timeout 300 python scripts/gpu_dropin_latency.py > gpurun_out/r02f_dropin_latency.jsonl 2> gpurun_out/r02f_dropin.err; echo "dropin exit=$?"; cut -c1-230 gpurun_out/r02f_dropin_latency.jsonl
timeout 900 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench exit=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','whole_step_frac','exact_rescans')}, d['verify']['mismatches'], d['roofline']['frac']); print([(s['batch'], round(s['p50_ms'],3), round(s['frac'],3), s['exact_rescans'], s['verify']['mismatches']) for s in d['sweep']]); c=d.get('clustered'); print(c['qps'], c['exact_rescans_per_step'], c['verify']['mismatches'], c.get('batch64',{}).get('qps')); print(d['cpu_baseline']['value'], d['cpu_baseline']['as_deployed']['value'])
PY
# ---- ncu on the bench's own configuration (routed federation)
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-clustered-leg --verify-queries 0 --sweep=8,64"
$CMD > gpurun_out/r02f_plain.json 2> gpurun_out/r02f_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rr_ -c 700 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02f_ncu_l.out 2>&1
echo "launch list exit=$?"
CMD2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-clustered-leg --verify-queries 0 --sweep="
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:rr_gemm_scan_kernel<\(int\)2, \(bool\)0' -s 6 -c 1 -f -o gpurun_out/r02_gemm_routed $CMD2 > gpurun_out/r02f_ncu2.out 2>&1; echo "gemm routed capture exit=$?"
CMD3="python bench.py --batch 64 --steps 2 --warmup 3 --no-cpu-baseline --no-clustered-leg --verify-queries 0 --sweep="
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:rr_gemm_scan_kernel<\(int\)1, \(bool\)0' -s 6 -c 1 -f -o gpurun_out/r02_gemm64_routed $CMD3 > gpurun_out/r02f_ncu3.out 2>&1; echo "gemm64 routed capture exit=$?"
CMD4="python bench.py --batch 8 --steps 2 --warmup 3 --no-cpu-baseline --no-clustered-leg --verify-queries 0 --sweep="
ncu --set full --clock-control none --import-source on -k regex:rr_tc_scan -s 6 -c 1 -f -o gpurun_out/r02_scan_routed $CMD4 > gpurun_out/r02f_ncu4.out 2>&1; echo "scan routed capture exit=$?"
CMD5="python bench.py --workload textbooks --batch 1 --steps 3 --warmup 3 --no-cpu-baseline --no-clustered-leg --verify-queries 0 --sweep=64"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rr_ -c 300 --csv --log-file gpurun_out/r02_launches_small.csv $CMD5 > gpurun_out/r02f_ncu_l5.out 2>&1; echo "small launch list exit=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:rr_finalize_kernel<\(bool\)1' -s 2 -c 1 -f -o gpurun_out/r02_finalize_b1 $CMD5 > gpurun_out/r02f_ncu6.out 2>&1; echo "finalize capture exit=$?"
ls -la gpurun_out/*.ncu-rep | tail -6
