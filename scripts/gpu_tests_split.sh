#!/bin/bash
# Run the GPU suite as separate processes so one CUDA fault does not hide the rest.
# Usage (on the GPU box, from the repo root): bash scripts/gpu_tests_split.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, pytest args...
  local name=$1; shift
  timeout 900 python -m pytest "$@" -q --timeout 600 -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "$name exit=$? $(tail -1 gpurun_out/$name.log)"
}
run router_rerank tests/test_gpu_router_rerank.py
run index_exact tests/test_gpu_index.py -k "(exact or reconstruct or normalize or errors or offset) and not fuzz"
run index_tcscan tests/test_gpu_index.py -k "tc_scan_candidates"
run index_auto tests/test_gpu_index.py -k "not exact and not tc_scan_candidates and not reconstruct and not normalize and not errors and not offset and not gemm and not config0 and not multi_million and not near_duplicate and not fuzz"
run index_gemm tests/test_gpu_index.py -k "gemm"
run index_sizes tests/test_gpu_index.py -k "config0 or multi_million or near_duplicate"
run index_fuzz tests/test_gpu_index.py -k "fuzz"
run federation tests/test_gpu_federation.py
