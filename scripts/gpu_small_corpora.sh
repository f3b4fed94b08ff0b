#!/bin/bash
# configs[0] / configs[1] of BASELINE.json: the small MedRAG sources alone (latency-bound, not roofline cases)
mkdir -p gpurun_out
for W in textbooks medqa_ts; do
python bench.py --workload $W --batch 64 --steps 50 --warmup 10 --no-cpu-baseline --sweep 1,8,64,256,2048 > gpurun_out/w_$W.json 2> gpurun_out/w_$W.err
python - <<PY
import json
d=json.load(open('gpurun_out/w_$W.json'))
print("$W", d['config']['rows'], "e2e B=64 %.0f q/s %.3f ms |" % (d['e2e']['value'], d['e2e']['p50_ms']), " | ".join("B=%d %.0f q/s p50 %.3f p99 %.3f ms" % (s['batch'], s['qps'], s['p50_ms'], s['p99_ms']) for s in d['sweep']))
PY
done
