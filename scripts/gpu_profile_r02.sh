#!/bin/bash
# Round-2 evidence run (one B200): kernel timings without a profiler, full ncu sets of the dominant kernels, the launch
# list of one train step and of one decode step, compute-sanitizer over the small cases.
mkdir -p gpurun_out
o=gpurun_out
timeout 120 python scripts/gpu_attn_bench.py > $o/r02_attn_bench.txt 2>&1
timeout 120 python scripts/gpu_gemm_ab.py > $o/r02_gemm_final.txt 2>&1
timeout 120 python scripts/gpu_skinny_bench.py > $o/r02_skinny_bench.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_2cta -s 2 -c 1 -f -o $o/r02_gemm2cta \
  python scripts/gpu_gemm_one.py 16384 28672 4096 > $o/r02_ncu_gemm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flash_ -s 3 -c 3 -f -o $o/r02_attn \
  python scripts/gpu_attn_one.py > $o/r02_ncu_attn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:skinny_gemm -c 8 -f -o $o/r02_skinny \
  python scripts/gpu_skinny_one.py > $o/r02_ncu_skinny.log 2>&1
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $o/r02_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-decode > $o/r02_launches_bench.log 2>&1
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python scripts/gpu_sanitizer_cases.py > $o/r02_memcheck.txt 2>&1
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python scripts/gpu_sanitizer_cases.py gemm attn decode > $o/r02_racecheck.txt 2>&1
tail -n 4 $o/r02_memcheck.txt $o/r02_racecheck.txt
ls -la $o | grep r02_ | tail -20
