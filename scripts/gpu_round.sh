#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (default = full federation), pubmed-slice sweep.
mkdir -p gpurun_out
bash scripts/gpu_tests_split.sh
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? $(tail -1 gpurun_out/smoke.log)"
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','p50_ms','p99_ms','gpu_launches','exact_rescans')}); print(d['e2e']); print(d['roofline']); print(d['clocks']); print(d['cpu_baseline'])
for s in d.get('sweep',[]): print(s)
PY
timeout 600 python bench.py --workload pubmed --rows-scale 0.25 --steps 10 --warmup 3 --no-cpu-baseline --routing all --sweep 64,128,256,512,1024,2048 > gpurun_out/gemm_bench.json 2> gpurun_out/gemm_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/gemm_bench.json'))
for s in d.get('sweep',[]): print(s)
PY
