#!/bin/bash
# ncu evidence for the dominant kernels (one GPU).  Usage: bash scripts/gpu_profile.sh <tag>
# 1) launch list (gpu__time_duration per launch of every rr_* kernel) of the default bench command, short
# 2) launch list of the 1/8-scale federation (the per-GPU shard of an 8-GPU run) at batch 64 and 1
# 3) --set full capture of the large-batch scan kernel (CTA pairs) on a PubMed-shaped slice (6M x 768, >> L2), batch 2048
# 4) --set full capture of the small-batch (swap-AB) scan kernel on the same slice, batch 16
# 5) --set full capture of the queries-in-TMEM scan in its HBM-bound regime (one CTA per slice), batch 64
# Every ncu command runs only after the same command exited 0 without ncu.
TAG=${1:-r01}
mkdir -p gpurun_out
CMD1="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --sweep 8,64"
$CMD1 > gpurun_out/${TAG}_plain1.json 2> gpurun_out/${TAG}_plain1.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rr_ -c 600 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD1 > gpurun_out/${TAG}_ncu1.out 2>&1
echo "launch list exit=$?"
CMD1B="python bench.py --rows-scale 0.125 --batch 64 --steps 4 --warmup 3 --no-cpu-baseline --sweep 1"
$CMD1B > gpurun_out/${TAG}_plain1b.json 2> gpurun_out/${TAG}_plain1b.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rr_ -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches_small.csv $CMD1B > gpurun_out/${TAG}_ncu1b.out 2>&1
echo "small launch list exit=$?"
CMD2="python bench.py --workload pubmed --rows-scale 0.25 --routing all --batch 2048 --steps 3 --warmup 3 --no-cpu-baseline --sweep="
$CMD2 > gpurun_out/${TAG}_plain2.json 2> gpurun_out/${TAG}_plain2.err &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:rr_gemm_scan_kernel<\(int\)2, \(bool\)0>' -s 3 -c 1 \
    -f -o gpurun_out/${TAG}_gemm $CMD2 > gpurun_out/${TAG}_ncu2.out 2>&1
# (fallback: per step the matching launches are <1,0> no-op, <2,0>, <1,1> no-op, <2,1> no-op)
[ -f gpurun_out/${TAG}_gemm.ncu-rep ] || ncu --set full --clock-control none --import-source on -k regex:rr_gemm_scan -s 13 -c 1 \
    -f -o gpurun_out/${TAG}_gemm $CMD2 > gpurun_out/${TAG}_ncu2.out 2>&1
echo "gemm capture exit=$?"
CMD3="python bench.py --workload pubmed --rows-scale 0.25 --routing all --batch 16 --steps 3 --warmup 3 --no-cpu-baseline --sweep="
$CMD3 > gpurun_out/${TAG}_plain3.json 2> gpurun_out/${TAG}_plain3.err &&
ncu --set full --clock-control none --import-source on -k regex:rr_tc_scan -s 3 -c 1 \
    -f -o gpurun_out/${TAG}_scan $CMD3 > gpurun_out/${TAG}_ncu3.out 2>&1
echo "scan capture exit=$?"
CMD4="python bench.py --workload pubmed --rows-scale 0.25 --routing all --batch 64 --steps 3 --warmup 3 --no-cpu-baseline --sweep="
$CMD4 > gpurun_out/${TAG}_plain4.json 2> gpurun_out/${TAG}_plain4.err &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:rr_gemm_scan_kernel<\(int\)1, \(bool\)0>' -s 3 -c 1 \
    -f -o gpurun_out/${TAG}_gemm64 $CMD4 > gpurun_out/${TAG}_ncu4.out 2>&1
[ -f gpurun_out/${TAG}_gemm64.ncu-rep ] || ncu --set full --clock-control none --import-source on -k regex:rr_gemm_scan -s 6 -c 1 \
    -f -o gpurun_out/${TAG}_gemm64 $CMD4 > gpurun_out/${TAG}_ncu4.out 2>&1
echo "gemm64 capture exit=$?"
