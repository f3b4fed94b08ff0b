#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_index.py tests/test_gpu_federation.py -k "gemm or federation or routing or smoke" -x -q --timeout 300 -p no:cacheprovider > gpurun_out/gemm_tests.log 2>&1
echo "gemm tests exit=$? $(tail -1 gpurun_out/gemm_tests.log)"
tail -30 gpurun_out/gemm_tests.log | grep -v "^$" | tail -15
timeout 600 python bench.py --workload pubmed --rows-scale 0.25 --steps 10 --warmup 3 --no-cpu-baseline --routing ${ROUTING:-all} --sweep ${SWEEP:-64,128,256,512,1024,2048} > gpurun_out/gemm_bench.json 2> gpurun_out/gemm_bench.err
echo "bench exit=$?"; tail -3 gpurun_out/gemm_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/gemm_bench.json'))
for s in d.get('sweep',[]): print(s)
PY
if [ -n "$FED" ]; then
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sweep 64,256,1024 > gpurun_out/fed_bench.json 2> gpurun_out/fed_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/fed_bench.json'))
for s in d.get('sweep',[]): print(s)
PY
fi
