#!/bin/bash
# Parity tests, then A/B of the <=128-query layout of the queries-in-TMEM scan on the 1/8-scale and full federation.
mkdir -p gpurun_out
bash scripts/gpu_tests_split.sh
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? $(tail -1 gpurun_out/smoke.log)"
for P in 1 0; do
for SC in 0.125 1.0; do
timeout 600 python bench.py --rows-scale $SC --batch 64 --steps 20 --warmup 5 --no-cpu-baseline --sweep 1,8,64,128 --index-param gemm_tmem_all=$P > gpurun_out/ab_${P}_${SC}.json 2> gpurun_out/ab_${P}_${SC}.err; echo "bench exit=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/ab_${P}_${SC}.json'))
print("tmem_all=$P scale=$SC", {k:round(d[k],4) if isinstance(d[k],float) else d[k] for k in ('value','ms_per_step','gpu_launches','exact_rescans')}, 'e2e', round(d['e2e']['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), 'share', round(d['roofline']['kernel_share_of_step'],4))
for s in d.get('sweep',[]): print('   ', s['batch'], round(s['qps'],1), round(s['p50_ms'],4), round(s['p99_ms'],4), round(s['frac'],4))
PY
done; done
