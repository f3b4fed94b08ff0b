#!/bin/bash
# Multi-GPU check: N ranks over NCCL, federation row-sharded (strong scaling)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_$N.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29510 \
   scripts/multi_gpu_check.py > gpurun_out/check_n$N.log 2>&1
echo "parity N=$N exit=$?"; grep -v "^\[W\|NCCL" gpurun_out/check_n$N.log | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 20 --warmup 5 ${BENCH_ARGS} > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N exit=$?"; tail -5 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
