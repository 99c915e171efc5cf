# A/B runs of the decode step on one box
timeout 300 python -m pytest tests/test_decode_gpu.py -q -m gpu -x 2>&1 | tail -4 > gpurun_out/ds_tests.txt
MM_DECODE_STACK=0 timeout 200 python scripts/gpu_decode_bench.py > gpurun_out/ds_perop.json 2> gpurun_out/ds_perop.err
timeout 200 python scripts/gpu_decode_bench.py > gpurun_out/ds_stack.json 2> gpurun_out/ds_stack.err
timeout 200 python scripts/gpu_ds_trace.py 400 > gpurun_out/ds_trace.txt 2>&1
