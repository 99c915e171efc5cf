#!/bin/bash
# Round-2 evidence run (one B200): kernel timings without a profiler, A/B of the GEMM schedule switches, full ncu sets of
# the two dominant kernels, the launch list of one train step, compute-sanitizer over the small cases.
mkdir -p gpurun_out
o=gpurun_out
timeout 120 python scripts/gpu_attn_bench.py > $o/r02_attn_bench.txt 2>&1
for dyn in 1 0; do for hint in 1 0; do
  MM_GEMM_DYNAMIC=$dyn MM_GEMM_L2HINT=$hint timeout 120 python scripts/gpu_gemm_ab.py >> $o/r02_gemm_ab.txt 2>&1
done; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_2cta -s 2 -c 1 -f -o $o/r02_gemm2cta \
  python scripts/gpu_gemm_one.py 16384 28672 4096 > $o/r02_ncu_gemm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flash_bwd -s 2 -c 2 -f -o $o/r02_attn_bwd \
  python scripts/gpu_attn_one.py > $o/r02_ncu_attn.log 2>&1
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $o/r02_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-decode > $o/r02_launches_bench.log 2>&1
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python scripts/gpu_sanitizer_cases.py > $o/r02_memcheck.txt 2>&1
timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python scripts/gpu_sanitizer_cases.py gemm attn > $o/r02_racecheck.txt 2>&1
tail -3 $o/r02_attn_bench.txt $o/r02_gemm_ab.txt $o/r02_memcheck.txt $o/r02_racecheck.txt
ls -la $o | grep r02
