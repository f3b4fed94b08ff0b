#!/bin/bash
# Parity tests, then the 1/8-scale federation (the per-GPU shard of an 8-GPU run) at small batches, where
# the fixed cost per step matters most.
mkdir -p gpurun_out
bash scripts/gpu_tests_split.sh
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? $(tail -1 gpurun_out/smoke.log)"
timeout 600 python bench.py --rows-scale 0.125 --batch 2048 --steps 20 --warmup 5 --no-cpu-baseline --sweep 1,8,32,64,256,1024 > gpurun_out/fused_small.json 2> gpurun_out/fused_small.err; echo "bench exit=$?"; tail -3 gpurun_out/fused_small.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/fused_small.json'))
print({k:d[k] for k in ('value','ms_per_step','p50_ms','p99_ms','gpu_launches','exact_rescans')}); print(d['e2e']); print(d['roofline'])
for s in d.get('sweep',[]): print(s)
PY
