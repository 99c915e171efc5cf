#!/bin/bash
# 2 GPUs, final tree: the 2-rank == 1-rank equality tests and one bench line through the fused NVSwitch path.
mkdir -p gpurun_out
o=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 python -m pytest tests/test_dp_nccl_gpu.py -m gpu -q -rs > $o/r02_nccl_test_final.log 2>&1
tail -4 $o/r02_nccl_test_final.log | cut -c1-300
timeout 300 $TR --master-port 29521 bench.py --gpus 2 --steps 4 --warmup 3 > $o/r02_bench_n2_final.json 2> $o/r02_bench_n2_final.err
python - "$o/r02_bench_n2_final.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("N=2 ms/step", round(d["ms_per_step"], 1), "tok/s", round(d["value"]), "gemm TF/s", round(d["roofline"]["achieved"]), d["config"]["optimizer"][:160])
except Exception as e:
    print("unreadable:", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
