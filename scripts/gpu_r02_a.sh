#!/bin/bash
# Round-2 visit A (one GPU): parity suite, smoke, default bench (verify / clustered / cpu legs), small corpora, configs[2] and [4].
mkdir -p gpurun_out
bash scripts/gpu_tests_split.sh
timeout 600 python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider > gpurun_out/multi.log 2>&1; echo "multi exit=$? $(tail -1 gpurun_out/multi.log)"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? $(tail -1 gpurun_out/smoke.log)"
t0=$(date +%s)
timeout 1200 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench exit=$? in $(( $(date +%s) - t0 ))s"; tail -5 gpurun_out/r02a_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02a_bench.json'))
    print({k:d.get(k) for k in ('value','ms_per_step','p50_ms','p99_ms','gpu_launches','exact_rescans','whole_step_frac')}); print(d['e2e']); print(d['roofline']); print('verify',d['verify']); print(d['clocks']); print(d['cpu_baseline']); print('clustered', d.get('clustered'))
    for s in d.get('sweep',[]): print(s)
except Exception as e: print('bench parse failed', e)
PY
for wl in textbooks medqa_ts; do
  timeout 600 python bench.py --workload $wl --batch 64 --steps 200 --warmup 20 > gpurun_out/r02a_bench_$wl.json 2> gpurun_out/r02a_bench_$wl.err; echo "$wl exit=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02a_bench_$wl.json'))
    print({k:d.get(k) for k in ('value','ms_per_step','p50_ms','p99_ms','whole_step_frac','verify')}); print(d['e2e']); print(d['roofline'])
    for s in d.get('sweep',[]): print(s)
    print('clustered', d.get('clustered'))
except Exception as e: print('parse failed', e)
PY
done
timeout 900 python bench.py --workload feb4rag16 > gpurun_out/r02a_bench_feb4rag16.json 2> gpurun_out/r02a_bench_feb4rag16.err; echo "feb4rag16 exit=$?"; tail -3 gpurun_out/r02a_bench_feb4rag16.err
timeout 900 python bench.py --workload pubmed --no-cpu-baseline > gpurun_out/r02a_bench_pubmed.json 2> gpurun_out/r02a_bench_pubmed.err; echo "pubmed exit=$?"; tail -3 gpurun_out/r02a_bench_pubmed.err
python - <<'PY'
import json
for wl in ('feb4rag16','pubmed'):
    try:
        d=json.load(open(f'gpurun_out/r02a_bench_{wl}.json'))
        print(wl, {k:d.get(k) for k in ('value','ms_per_step','p50_ms','p99_ms','whole_step_frac','verify','exact_rescans')}); print(d['e2e']); print(d['roofline'])
        for s in d.get('sweep',[]): print(s)
    except Exception as e: print(wl, 'parse failed', e)
PY
timeout 300 python scripts/gpu_dropin_latency.py > gpurun_out/r02a_dropin_latency.jsonl 2> gpurun_out/r02a_dropin.err; echo "dropin exit=$?"; cat gpurun_out/r02a_dropin_latency.jsonl
