#!/bin/bash
# Whole train step under MM_GEMM_TMA_STORE = 1 (all plain bf16 results), 0 (none), 2 (all but the wgrad layout): one box, alternating.
mkdir -p gpurun_out
one() {
  MM_GEMM_TMA_STORE=$1 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-decode 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('store=$1', round(d['ms_per_step'],1), 'ms  gemm', round(d['roofline']['achieved']), 'TF/s  share', round(d['roofline']['gemm_share_of_step'],3), d['clocks']['sm_mhz'])"
}
for r in 1 2; do for s in 1 0 2; do one $s; done; done 2>&1 | tee gpurun_out/r02_step_store_ab.txt
