#!/bin/bash
# Parity tests + mid-batch sweep (full and 1/8-scale federation, PubMed slice all-routed)
mkdir -p gpurun_out
bash scripts/gpu_tests_split.sh
for CFG in "--rows-scale 1.0" "--rows-scale 0.125" "--workload pubmed --rows-scale 0.25 --routing all"; do
python bench.py $CFG --batch 2048 --steps 10 --warmup 3 --no-cpu-baseline --sweep 64,128,192,256,384,512,1024 > gpurun_out/mid.json 2> gpurun_out/mid.err
python - <<PY
import json
d=json.load(open('gpurun_out/mid.json'))
print("$CFG: B=2048 %.0f q/s %.2f ms frac %.3f |" % (d['value'], d['ms_per_step'], d['roofline']['frac']), " | ".join("B=%d %.0f q/s %.2f ms" % (s['batch'], s['qps'], s['p50_ms']) for s in d['sweep']), "rescans", d['exact_rescans'])
PY
done
