#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --rows-scale 0.125 --batch 64 --steps 10 --warmup 3 --no-cpu-baseline --sweep="
$CMD > gpurun_out/small64_plain.json 2> gpurun_out/small64_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rr_ -c 400 --csv \
    --log-file gpurun_out/small64_launches.csv $CMD > gpurun_out/small64_ncu.out 2>&1
echo "exit=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/small64_plain.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['roofline']['avg_launch_ms'])
PY
