mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decode_gpu.py -x -q -m gpu > gpurun_out/r02_dec_tests.txt 2>&1; tail -5 gpurun_out/r02_dec_tests.txt
timeout 200 python scripts/gpu_skinny_bench.py --batches > gpurun_out/r02_skinny_batches.txt 2>&1; cat gpurun_out/r02_skinny_batches.txt
timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python scripts/gpu_sanitizer_cases.py attn > gpurun_out/r02_racecheck_attn.txt 2>&1; tail -3 gpurun_out/r02_racecheck_attn.txt
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_dec.json 2> gpurun_out/r02_bench_dec.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_dec.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], json.dumps(d['decode'])[:1500])"
