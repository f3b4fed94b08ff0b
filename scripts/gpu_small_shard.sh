#!/bin/bash
# proxy for one rank of the 8-GPU run: the federation at 1/8 scale on one GPU (no NCCL)
mkdir -p gpurun_out
CMD="python bench.py --rows-scale 0.125 --steps 10 --warmup 3 --no-cpu-baseline --sweep 64"
$CMD > gpurun_out/small_plain.json 2> gpurun_out/small_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rr_ -c 500 --csv \
    --log-file gpurun_out/small_launches.csv $CMD > gpurun_out/small_ncu.out 2>&1
echo "exit=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/small_plain.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['roofline']['avg_launch_ms'])
for s in d['sweep']: print(s)
PY
