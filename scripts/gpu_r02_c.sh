#!/bin/bash
# Round-2 visit C: PDL everywhere + fused begin/prep: parity suite, small-corpus benches, timeline trace of the gemm scan.
mkdir -p gpurun_out
bash scripts/gpu_tests_split.sh
RR_TRACE=1 timeout 300 python scripts/gpu_trace.py 125000 64 > gpurun_out/r02c_trace_64.txt 2>&1; echo "trace exit=$?"; head -70 gpurun_out/r02c_trace_64.txt
for wl in textbooks medqa_ts; do
  timeout 600 python bench.py --workload $wl --batch 64 --steps 200 --warmup 20 --no-cpu-baseline --no-clustered-leg > gpurun_out/r02c_bench_$wl.json 2> gpurun_out/r02c_bench_$wl.err; echo "$wl exit=$?"
  RR_NO_PDL=1 timeout 600 python bench.py --workload $wl --batch 64 --steps 200 --warmup 20 --no-cpu-baseline --no-clustered-leg > gpurun_out/r02c_bench_${wl}_nopdl.json 2> gpurun_out/r02c_bench_${wl}_nopdl.err; echo "$wl nopdl exit=$?"
done
python - <<'PY'
import json
for wl in ('textbooks','textbooks_nopdl','medqa_ts','medqa_ts_nopdl'):
    try:
        d=json.load(open(f'gpurun_out/r02c_bench_{wl}.json'))
        print(wl, 'e2e p50', d['e2e']['p50_ms'], [(s['batch'], round(s['p50_ms'],4), s['verify']['mismatches']) for s in d['sweep']])
    except Exception as e: print(wl,'parse failed',e)
PY
timeout 300 python scripts/gpu_dropin_latency.py > gpurun_out/r02c_dropin_latency.jsonl 2> gpurun_out/r02c_dropin.err; echo "dropin exit=$?"; cat gpurun_out/r02c_dropin_latency.jsonl
timeout 900 python bench.py --no-cpu-baseline --no-clustered-leg --steps 10 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench exit=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','whole_step_frac','verify')}); print([(s['batch'], round(s['p50_ms'],3), round(s['frac'],3), s['verify']['mismatches']) for s in d['sweep']])
PY
