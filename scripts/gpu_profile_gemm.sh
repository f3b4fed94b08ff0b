#!/bin/bash
# ncu --set full of the large-batch scan kernel at two batch sizes (pubmed slice 6M x 768)
TAG=${1:-r01}
mkdir -p gpurun_out
for B in ${BATCHES:-1024 128}; do
CMD="python bench.py --workload pubmed --rows-scale 0.25 --routing all --batch $B --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_gemm_plain_$B.json 2> gpurun_out/${TAG}_gemm_plain_$B.err &&
ncu --set full --clock-control none --import-source on -k regex:rr_gemm_scan -s 3 -c 1 \
    -f -o gpurun_out/${TAG}_gemm_$B $CMD > gpurun_out/${TAG}_gemm_ncu_$B.out 2>&1
echo "B=$B exit=$?"
done
