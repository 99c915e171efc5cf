mkdir -p gpurun_out
timeout 120 python scripts/gpu_norm_bench.py > gpurun_out/r02_norm_bench2.txt 2>&1; cat gpurun_out/r02_norm_bench2.txt
timeout 120 python scripts/gpu_gemm_ab.py > gpurun_out/r02_gemm_box2.txt 2>&1; cat gpurun_out/r02_gemm_box2.txt
timeout 200 python -m pytest -q -m gpu tests/test_kernels_gpu.py -k "not gemm and not attention" 2>&1 | tail -2
timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-decode > gpurun_out/r02_bench_box2.json 2> gpurun_out/r02_bench_box2.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_box2.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['achieved'], d['clocks'])"
nvidia-smi --query-gpu=name,power.limit,power.max_limit,temperature.gpu,clocks.sm --format=csv
