#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (default = full federation), reference arm.
mkdir -p gpurun_out
bash scripts/gpu_tests_split.sh
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? $(tail -1 gpurun_out/smoke.log)"
timeout 900 python bench.py --sweep 1,8,64,256,1024 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
nproc; free -g | head -2
