#!/bin/bash
# A-panel L2 budget of the 2-CTA GEMM rasterisation (MM_GEMM_PANEL_MB): DRAM bytes + time of the gate/up launch, then the step.
mkdir -p gpurun_out
{
for mb in 32 40 48 64; do
  echo "== MM_GEMM_PANEL_MB=$mb"
  MM_GEMM_PANEL_MB=$mb timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
     -k regex:gemm_tcgen05_2cta -s 2 -c 1 python scripts/gpu_gemm_one.py 16384 28672 4096 2>&1 | grep -E "dram__bytes|gpu__time"
  MM_GEMM_PANEL_MB=$mb timeout 100 python scripts/gpu_gemm_ab.py 2>&1 | grep -E "fwd gate|fwd qkv"
done
one() {
  MM_GEMM_PANEL_MB=$1 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-decode 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('panel_mb=$1', round(d['ms_per_step'],1), 'ms  gemm', round(d['roofline']['achieved']), 'TF/s', d['clocks']['sm_mhz'])"
}
one 32; one 48; one 32; one 48
} 2>&1 | tee gpurun_out/r02_panel_sweep.txt
