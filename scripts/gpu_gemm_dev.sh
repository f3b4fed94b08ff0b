#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_index.py -k "gemm" -x -q --timeout 300 -p no:cacheprovider > gpurun_out/gemm_tests.log 2>&1
echo "gemm tests exit=$? $(tail -1 gpurun_out/gemm_tests.log)"
tail -30 gpurun_out/gemm_tests.log
timeout 600 python bench.py --workload pubmed --rows-scale 0.25 --steps 10 --warmup 3 --no-cpu-baseline --routing all --sweep 64,128,256,512,1024,2048 > gpurun_out/gemm_bench.json 2> gpurun_out/gemm_bench.err
echo "bench exit=$?"; tail -3 gpurun_out/gemm_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/gemm_bench.json'))
for s in d.get('sweep',[]): print(s)
PY
