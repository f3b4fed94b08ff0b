#!/bin/bash
# Round-2 visit I (one GPU): 64 KB TMA stages for CTA pairs (gemm_stage_kc = 4) and the lockstep window.
# New gemm tests on the default build, then the WHOLE parity suite with RR_STAGE_KC=4 forced (what a flipped default
# would run), then A/B on the bench's own workloads.
mkdir -p gpurun_out
run() { # name, pytest args...
  local name=$1; shift
  timeout 900 python -m pytest "$@" -q --timeout 600 -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "$name exit=$? $(tail -1 gpurun_out/$name.log)"
}
run i_gemm_default tests/test_gpu_index.py -k "gemm"
export RR_STAGE_KC=4
run i_skc4_index tests/test_gpu_index.py
run i_skc4_federation tests/test_gpu_federation.py tests/test_gpu_glue.py
unset RR_STAGE_KC
AB="--no-cpu-baseline --no-clustered-leg --steps 20 --warmup 5 --verify-queries 2"
ab() { # tag, env..., -- bench args
  local tag=$1; shift
  env "$@" timeout 300 python bench.py $AB $BARGS > gpurun_out/r02i_$tag.json 2> gpurun_out/r02i_$tag.err; echo "$tag exit=$?"
}
BARGS="--sweep="
ab fed_k2_w4 RR_STAGE_KC=2 RR_LOCKSTEP=4
ab fed_k4_w4 RR_STAGE_KC=4 RR_LOCKSTEP=4
ab fed_k4_w8 RR_STAGE_KC=4 RR_LOCKSTEP=8
ab fed_k4_w2 RR_STAGE_KC=4 RR_LOCKSTEP=2
ab fed_k2_w8 RR_STAGE_KC=2 RR_LOCKSTEP=8
BARGS="--workload pubmed --sweep=512,1024,4096"
ab pubmed_k2 RR_STAGE_KC=2
ab pubmed_k4 RR_STAGE_KC=4
python - <<'PY'
import json
for f in ('fed_k2_w4','fed_k4_w4','fed_k4_w8','fed_k4_w2','fed_k2_w8','pubmed_k2','pubmed_k4'):
    try:
        d=json.load(open(f'gpurun_out/r02i_{f}.json')); r=d['roofline']
        print(f, round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(r['frac'],4), 'whole', round(d['whole_step_frac'],4), 'verify', d['verify']['mismatches'], 'rescans', d['exact_rescans'], 'clk', d['clocks']['sm_mhz'],
              [(s['batch'], round(s['p50_ms'],3), round(s['frac'],3), s['verify']['mismatches']) for s in d.get('sweep',[])])
    except Exception as e: print(f,'parse failed',e)
PY
