#!/bin/bash
# Round-2 visit G (one GPU): A/B of the large-batch step on ONE box: default vs RR_NO_PDL=1; pubmed batch 2048.
mkdir -p gpurun_out
ARGS="--steps 12 --warmup 4 --no-cpu-baseline --no-clustered-leg --verify-queries 4 --sweep="
for tag in a_default b_nopdl c_default d_nopdl; do
  if [[ $tag == *nopdl ]]; then export RR_NO_PDL=1; else unset RR_NO_PDL; fi
  timeout 300 python bench.py $ARGS > gpurun_out/r02g_$tag.json 2> gpurun_out/r02g_$tag.err; echo "$tag exit=$?"
done
unset RR_NO_PDL
timeout 300 python bench.py --workload pubmed $ARGS > gpurun_out/r02g_pubmed.json 2> gpurun_out/r02g_pubmed.err; echo "pubmed exit=$?"
RR_NO_PDL=1 timeout 300 python bench.py --workload pubmed $ARGS > gpurun_out/r02g_pubmed_nopdl.json 2> gpurun_out/r02g_pubmed_nopdl.err; echo "pubmed nopdl exit=$?"
python - <<'PY'
import json
for t in ('a_default','b_nopdl','c_default','d_nopdl','pubmed','pubmed_nopdl'):
    try:
        d=json.load(open(f'gpurun_out/r02g_{t}.json')); r=d['roofline']
        print(t, round(d['value'],1), round(d['ms_per_step'],2), 'scan ms', round(r['avg_launch_ms'],3), 'frac', round(r['frac'],3), d['clocks']['sm_mhz'], d['verify']['mismatches'])
    except Exception as e: print(t,'failed',e)
PY
