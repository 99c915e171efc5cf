#!/bin/bash
# TMA-store epilogue on/off, same box, alternating (MM_GEMM_TMA_STORE is read once per process).
mkdir -p gpurun_out
for r in 1 2; do
  for s in 1 0; do
    echo "== MM_GEMM_TMA_STORE=$s"
    MM_GEMM_TMA_STORE=$s timeout 120 python scripts/gpu_gemm_ab.py 2>&1 | grep TFLOP
  done
done | tee gpurun_out/r02_gemm_store_ab.txt
