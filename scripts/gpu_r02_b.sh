#!/bin/bash
# Round-2 visit B (one GPU): new parity tests; where the time of a small-corpus step goes (BASELINE configs[0]):
# launch list + ncu --set full of the swap-AB scan (batch 1), the queries-in-TMEM scan (batch 64), finalize, router.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_glue.py tests/test_gpu_index.py -k "glue or read_index or collector or roundtrip" -q -p no:cacheprovider > gpurun_out/r02b_tests.log 2>&1; echo "tests exit=$? $(tail -1 gpurun_out/r02b_tests.log)"
CMD1="python bench.py --workload textbooks --batch 1 --steps 4 --warmup 3 --no-cpu-baseline --no-clustered-leg --verify-queries 0 --sweep="
CMD64="python bench.py --workload textbooks --batch 64 --steps 4 --warmup 3 --no-cpu-baseline --no-clustered-leg --verify-queries 0 --sweep="
$CMD1 > gpurun_out/r02b_plain1.json 2> gpurun_out/r02b_plain1.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rr_ -c 300 --csv --log-file gpurun_out/r02b_launches_b1.csv $CMD1 > gpurun_out/r02b_ncu_l1.out 2>&1
echo "launch list b1 exit=$?"
$CMD64 > gpurun_out/r02b_plain64.json 2> gpurun_out/r02b_plain64.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rr_ -c 300 --csv --log-file gpurun_out/r02b_launches_b64.csv $CMD64 > gpurun_out/r02b_ncu_l64.out 2>&1
echo "launch list b64 exit=$?"
ncu --set full --clock-control none --import-source on -k regex:rr_tc_scan -s 8 -c 1 -f -o gpurun_out/r02b_tc_b1 $CMD1 > gpurun_out/r02b_ncu_tc.out 2>&1; echo "tc capture exit=$?"
ncu --set full --clock-control none --import-source on -k regex:rr_finalize -s 16 -c 1 -f -o gpurun_out/r02b_fin_b1 $CMD1 > gpurun_out/r02b_ncu_fin.out 2>&1; echo "fin capture exit=$?"
ncu --set full --clock-control none --import-source on -k regex:rr_router -s 8 -c 1 -f -o gpurun_out/r02b_router_b1 $CMD1 > gpurun_out/r02b_ncu_router.out 2>&1; echo "router capture exit=$?"
ncu --set full --clock-control none --import-source on -k regex:rr_gemm_scan -s 8 -c 1 -f -o gpurun_out/r02b_gemm_b64 $CMD64 > gpurun_out/r02b_ncu_gemm.out 2>&1; echo "gemm capture exit=$?"
ncu --set full --clock-control none --import-source on -k regex:rr_merge -s 8 -c 1 -f -o gpurun_out/r02b_merge_b64 $CMD64 > gpurun_out/r02b_ncu_merge.out 2>&1; echo "merge capture exit=$?"
ls -la gpurun_out/*.ncu-rep
