timeout 200 python -m pytest -q -m gpu tests/test_kernels_gpu.py -x -k "attention" 2>&1 | tail -2
timeout 120 python scripts/gpu_attn_bench.py 2>&1 | grep bwd
timeout 120 python scripts/gpu_attn_bench.py 2>&1 | grep "bwd tc=True"
