#!/bin/bash
# swap-AB scan (path 1) vs queries-in-TMEM scan (path 2) at small batches, 1/8-scale and full federation
mkdir -p gpurun_out
for P in 1 2; do
for SC in 0.125 1.0; do
python bench.py --rows-scale $SC --batch 32 --steps 20 --warmup 5 --no-cpu-baseline --sweep 1,4,8,16,32,48 --scan-path $P > gpurun_out/abs.json 2> gpurun_out/abs.err
python - <<PY
import json
d=json.load(open('gpurun_out/abs.json'))
print("path $P scale=$SC:", " | ".join("B=%d %.0f q/s %.3f ms" % (s['batch'], s['qps'], s['p50_ms']) for s in d['sweep']))
PY
done; done
