#!/bin/bash
# small-batch regime: swap-AB kernel (path 1) vs queries-in-TMEM kernel (path 2 / 3), full and 1/8-scale federation
mkdir -p gpurun_out
for SCALE in 1.0 0.125; do
for P in 1 2 3; do
python bench.py --rows-scale $SCALE --batch 64 --steps 10 --warmup 3 --no-cpu-baseline --sweep 1,8,32,64,128 --scan-path $P > gpurun_out/pathcmp_${SCALE}_$P.json 2> gpurun_out/pathcmp_${SCALE}_$P.err
python - <<PY
import json
d=json.load(open('gpurun_out/pathcmp_${SCALE}_$P.json'))
print("scale $SCALE path $P:", " | ".join("B=%d %.0f q/s p50 %.2f ms (%s %.2f)" % (s['batch'], s['qps'], s['p50_ms'], s['bound'], s['frac'] or 0) for s in d['sweep']))
PY
done; done
