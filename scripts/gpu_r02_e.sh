#!/bin/bash
# Round-2 visit E: validation of the small-corpus work (tests, small benches, A/B of the two-accumulator layout), default bench.
mkdir -p gpurun_out
bash scripts/gpu_tests_split.sh
for wl in textbooks medqa_ts; do
  timeout 600 python bench.py --workload $wl --batch 64 --steps 200 --warmup 20 --no-cpu-baseline --no-clustered-leg > gpurun_out/r02e_bench_$wl.json 2> gpurun_out/r02e_bench_$wl.err; echo "$wl exit=$?"
  timeout 600 python bench.py --workload $wl --batch 64 --steps 200 --warmup 20 --no-cpu-baseline --no-clustered-leg --index-param gemm_tmem_all=0 > gpurun_out/r02e_bench_${wl}_2acc.json 2> gpurun_out/r02e_bench_${wl}_2acc.err; echo "$wl 2acc exit=$?"
done
python - <<'PY'
import json
for wl in ('textbooks','textbooks_2acc','medqa_ts','medqa_ts_2acc'):
    try:
        d=json.load(open(f'gpurun_out/r02e_bench_{wl}.json'))
        print(wl, 'e2e p50', round(d['e2e']['p50_ms'],4), [(s['batch'], round(s['p50_ms'],4), round(s['whole_step_frac'],3), s['exact_rescans'], s['verify']['mismatches']) for s in d['sweep']])
    except Exception as e: print(wl,'parse failed',e)
PY
timeout 300 python scripts/gpu_dropin_latency.py > gpurun_out/r02e_dropin_latency.jsonl 2> gpurun_out/r02e_dropin.err; echo "dropin exit=$?"; cut -c1-230 gpurun_out/r02e_dropin_latency.jsonl
CMD1="python bench.py --workload textbooks --batch 1 --steps 4 --warmup 3 --no-cpu-baseline --no-clustered-leg --verify-queries 0 --sweep=64"
$CMD1 > gpurun_out/r02e_plain1.json 2> gpurun_out/r02e_plain1.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rr_ -c 400 --csv --log-file gpurun_out/r02e_launches_small.csv $CMD1 > gpurun_out/r02e_ncu_l1.out 2>&1
echo "launch list exit=$?"
timeout 900 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench exit=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02e_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','whole_step_frac','exact_rescans')}, d['verify']['mismatches'], d['roofline']['frac']); print([(s['batch'], round(s['p50_ms'],3), round(s['frac'],3), s['exact_rescans'], s['verify']['mismatches']) for s in d['sweep']]); c=d.get('clustered'); print(c['qps'], c['exact_rescans_per_step'], c['verify']['mismatches'], c.get('batch64',{}).get('qps'))
PY
timeout 900 python bench.py --workload feb4rag16 --no-cpu-baseline > gpurun_out/r02e_bench_feb4rag16.json 2> gpurun_out/r02e_bench_feb4rag16.err; echo "feb exit=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02e_bench_feb4rag16.json'))
print({k:d.get(k) for k in ('value','ms_per_step','whole_step_frac','exact_rescans')}, d['verify']['mismatches'], d['e2e']['value']); print([(s['batch'], round(s['p50_ms'],3), round(s['frac'],3), s['exact_rescans'], s['verify']['mismatches']) for s in d['sweep']])
PY
