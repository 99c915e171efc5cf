#!/bin/bash
# DRAM traffic of one GEMM shape under different rasterisation group sizes (ncu, two metrics only)
M=$1; N=$2; K=$3
for mode in 0 1; do
  for gm in 2 4 8 16 32 64; do
    MM_GEMM_2CTA=$mode MM_GEMM_GM=$gm ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tcgen05 -s 2 -c 1 --csv python scripts/gpu_gemm_one.py $M $N $K 2>/dev/null | grep -E "dram__bytes_read|gpu__time" | awk -F'","' -v m=$mode -v g=$gm '{printf "2cta=%s gm=%s %s %s %s\n", m, g, $(NF-2), $(NF-1), $NF}'
  done
done
