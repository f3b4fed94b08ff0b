#!/bin/bash
mkdir -p gpurun_out
for CFG in "--rows-scale 1.0" "--rows-scale 0.125" "--workload pubmed --rows-scale 0.25 --routing all"; do
for P in 2 3; do
python bench.py $CFG --batch 256 --steps 10 --warmup 3 --no-cpu-baseline --sweep 192,256,384,512,768,1024 --scan-path $P > gpurun_out/pc2.json 2> gpurun_out/pc2.err
python - <<PY
import json
d=json.load(open('gpurun_out/pc2.json'))
print("$CFG path $P:", " | ".join("B=%d %.0f q/s %.2f ms" % (s['batch'], s['qps'], s['p50_ms']) for s in d['sweep']))
PY
done; done
