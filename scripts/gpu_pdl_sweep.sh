for m in 0 1 3 9 11; do
  MM_PDL_MODE=$m timeout 200 python scripts/gpu_decode_bench.py > gpurun_out/decode_mode_$m.json 2> gpurun_out/decode_mode_$m.err
done
MM_PDL_MODE=1 MM_SKINNY_GU_V1=1 timeout 200 python scripts/gpu_decode_bench.py > gpurun_out/decode_mode_1guv1.json 2> gpurun_out/decode_mode_1guv1.err
MM_PDL_MODE=9 MM_SKINNY_GU_V1=1 timeout 200 python scripts/gpu_decode_bench.py > gpurun_out/decode_mode_9guv1.json 2> gpurun_out/decode_mode_9guv1.err
