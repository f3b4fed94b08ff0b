#!/bin/bash
# Round-2 visit J (one GPU): 64 KB stages also for single-tile pair launches (129 ... 256 routed queries, the ridge)?
mkdir -p gpurun_out
run() { local name=$1; shift
  timeout 900 python -m pytest "$@" -q --timeout 600 -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "$name exit=$? $(tail -1 gpurun_out/$name.log)"; }
export RR_STAGE_KC_MINQ=129
run j_minq_gemm tests/test_gpu_index.py -k "gemm or fuzz or batches or auto"
unset RR_STAGE_KC_MINQ
AB="--no-cpu-baseline --no-clustered-leg --steps 60 --warmup 5 --verify-queries 2 --sweep="
ab() { local tag=$1; shift
  env "$@" timeout 300 python bench.py $AB $BARGS > gpurun_out/r02j_$tag.json 2> gpurun_out/r02j_$tag.err; echo "$tag exit=$?"; }
BARGS="--batch 256"
ab fed256_base RR_X=0
ab fed256_minq RR_STAGE_KC_MINQ=129
BARGS="--workload pubmed --batch 256"
ab pub256_base RR_X=0
ab pub256_minq RR_STAGE_KC_MINQ=129
BARGS="--workload pubmed --batch 200"
ab pub200_base RR_X=0
ab pub200_minq RR_STAGE_KC_MINQ=129
python - <<'PY'
import json
for f in ('fed256_base','fed256_minq','pub256_base','pub256_minq','pub200_base','pub200_minq'):
    try:
        d=json.load(open(f'gpurun_out/r02j_{f}.json')); r=d['roofline']
        print(f, round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'p50', round(d['p50_ms'],3), r['bound'], 'frac', round(r['frac'],4), 'whole', round(d['whole_step_frac'],4), 'verify', d['verify']['mismatches'], 'rescans', d['exact_rescans'], 'clk', d['clocks']['sm_mhz'])
    except Exception as e: print(f,'parse failed',e)
PY
