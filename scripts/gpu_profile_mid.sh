#!/bin/bash
# ncu --set full of the queries-in-TMEM scan in its two non-headline regimes, PubMed-shaped slice (5.975 M x 768):
#  a) one CTA pair per corpus slice (<= 256 routed queries, 74 slices): tensor and HBM both near saturation
#  b) one CTA per slice at 64 queries (HBM-bound)
TAG=${1:-r01m}
mkdir -p gpurun_out
BASE="python bench.py --workload pubmed --rows-scale 0.25 --routing all --steps 3 --warmup 3 --no-cpu-baseline --sweep="
for CFG in "256 3 pair256" "512 3 pair512" "64 0 one64" "128 0 one128"; do
set -- $CFG
$BASE --batch $1 --scan-path $2 > gpurun_out/${TAG}_$3.json 2> gpurun_out/${TAG}_$3.err
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_$3.json'))
print("$3", round(d['ms_per_step'],4), d['roofline']['bound'], round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), 'tflops', d['roofline'].get('tflops'), 'hbm', d['roofline'].get('hbm_gbs'))
PY
done
ncu --set full --clock-control none --import-source on -k regex:rr_gemm_scan -s 3 -c 1 -f -o gpurun_out/${TAG}_pair256 $BASE --batch 256 --scan-path 3 > gpurun_out/${TAG}_ncu_a.out 2>&1; echo "capture a exit=$?"
ncu --set full --clock-control none --import-source on -k regex:rr_gemm_scan -s 3 -c 1 -f -o gpurun_out/${TAG}_one64 $BASE --batch 64 > gpurun_out/${TAG}_ncu_b.out 2>&1; echo "capture b exit=$?"
