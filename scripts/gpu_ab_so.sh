#!/bin/bash
# A/B of two builds of the C-ABI library on ONE box (box-to-box variation is a few per cent): ab/_C_prev.so vs the tree's.
mkdir -p gpurun_out
cp metamorph_b200/_C.so /tmp/_C_new.so
one() {
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-decode 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],1), 'ms  gemm', round(d['roofline']['achieved']), 'TF/s  share', round(d['roofline']['gemm_share_of_step'],3), d['clocks']['sm_mhz'])"
}
for r in 1 2; do
  cp ab/_C_prev.so metamorph_b200/_C.so; one prev
  cp /tmp/_C_new.so metamorph_b200/_C.so; one new
done 2>&1 | tee gpurun_out/r02_ab_so.txt
