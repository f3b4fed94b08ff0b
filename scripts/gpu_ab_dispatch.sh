#!/bin/bash
mkdir -p gpurun_out
for P in model nomodel; do
if [ $P = nomodel ]; then export RR_NO_DISPATCH_MODEL=1; else unset RR_NO_DISPATCH_MODEL; fi
for SC in 0.125 1.0; do
python bench.py --rows-scale $SC --batch 64 --steps 20 --warmup 5 --no-cpu-baseline --sweep 64,96,128 > gpurun_out/abd.json 2> gpurun_out/abd.err
python - <<PY
import json
d=json.load(open('gpurun_out/abd.json'))
print("$P scale=$SC: B=64 %.0f q/s %.3f ms frac %.3f share %.3f |" % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_share_of_step']), " | ".join("B=%d %.0f q/s %.3f ms" % (s['batch'], s['qps'], s['p50_ms']) for s in d['sweep']))
PY
done
CMD="python bench.py --rows-scale 0.125 --batch 64 --steps 3 --warmup 3 --no-cpu-baseline --sweep="
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rr_ -c 200 --csv --log-file gpurun_out/abd_${P}_launches.csv $CMD > gpurun_out/abd_ncu.out 2>&1
done
