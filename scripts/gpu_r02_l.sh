#!/bin/bash
# Round-2 visit L (one GPU): from how many ROUTED queries should a batch above 128 take the CTA-pair variant?
mkdir -p gpurun_out
AB="--no-cpu-baseline --no-clustered-leg --steps 60 --warmup 5 --verify-queries 2 --sweep="
ab() { local tag=$1; shift
  env "$@" timeout 300 python bench.py $AB $BARGS > gpurun_out/r02l_$tag.json 2> gpurun_out/r02l_$tag.err; echo "$tag exit=$?"; }
BARGS="--batch 256"
for m in 129 97 65 33; do ab fed256_p$m RR_PAIR_MIN=$m; done
BARGS="--batch 160"
for m in 129 65 33; do ab fed160_p$m RR_PAIR_MIN=$m; done
python - <<'PY'
import json
for f in ('fed256_p129','fed256_p97','fed256_p65','fed256_p33','fed160_p129','fed160_p65','fed160_p33'):
    try:
        d=json.load(open(f'gpurun_out/r02l_{f}.json')); r=d['roofline']
        print(f, round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'p50', round(d['p50_ms'],3), r['bound'], 'frac', round(r['frac'],4), 'whole', round(d['whole_step_frac'],4), 'verify', d['verify']['mismatches'], 'rescans', d['exact_rescans'], 'clk', d['clocks']['sm_mhz'])
    except Exception as e: print(f,'parse failed',e)
PY
