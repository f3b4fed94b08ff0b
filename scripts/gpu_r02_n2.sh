#!/bin/bash
# short 2-rank sanity of the bench (every rank must reach the final barrier): federation incl. clustered leg, feb4rag16
mkdir -p gpurun_out
N=2
run() { name=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N "$@" > gpurun_out/r02_bench_${name}_n$N.json 2> gpurun_out/r02_bench_${name}_n$N.err; echo "$name exit=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench_${name}_n$N.json'))
    print('$name', {k:d.get(k) for k in ('n_gpus','value','ms_per_step','whole_step_frac','exact_rescans')}, 'verify', d['verify']['mismatches'], d['verify']['queries'], d['verify']['rows_checked'], 'e2e', round(d['e2e']['value'],1), 'roofline', round(d['roofline']['frac'],3), round(d['roofline']['kernel_share_of_step'],3))
    print([(s['batch'], round(s['p50_ms'],3), round(s['frac'],3), round(s['whole_step_frac'],3), s['verify']['mismatches']) for s in d.get('sweep',[])])
    c=d.get('clustered')
    if c: print('clustered', round(c['qps'],1), c['exact_rescans_per_step'], c['verify']['mismatches'])
except Exception as e: print('$name parse failed', e)
PY
}
run federation --steps 20 --warmup 5
