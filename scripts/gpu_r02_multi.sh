#!/bin/bash
# Round-2 multi-GPU visit (N = number of visible GPUs): real-NCCL parity (pytest -m gpu torchrun test), then the
# bench lines of configs[3] (federation), configs[2] (pubmed) and configs[4] (feb4rag16) at N ranks, each with verify.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r02_multi_gpus_n$N.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_multi.py -q -s -p no:cacheprovider > gpurun_out/r02_multi_gpu_check_n$N.log 2>&1; echo "multi test exit=$? $(tail -1 gpurun_out/r02_multi_gpu_check_n$N.log)"; grep "world=" gpurun_out/r02_multi_gpu_check_n$N.log | tail -12
run() { # name, args...
  local name=$1; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/r02_bench_${name}_n$N.json 2> gpurun_out/r02_bench_${name}_n$N.err; echo "$name exit=$?"
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
run pubmed --workload pubmed --steps 10 --warmup 5 --no-clustered-leg
run feb4rag16 --workload feb4rag16 --steps 20 --warmup 5
