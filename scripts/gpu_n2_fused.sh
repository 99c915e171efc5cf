#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 500 python -m pytest tests/test_dp_nccl_gpu.py -m gpu -q -rs -s > $o/r02_nccl_test2.log 2>&1
tail -12 $o/r02_nccl_test2.log | cut -c1-300
timeout 400 $TR --master-port 29521 bench.py --gpus 2 --steps 6 --warmup 3 > $o/r02_bench_n2_fused.json 2> $o/r02_bench_n2_fused.err
MM_FUSED_REDUCE=0 timeout 400 $TR --master-port 29522 bench.py --gpus 2 --steps 6 --warmup 3 > $o/r02_bench_n2_fused_ag_only.json 2> $o/r02_bench_n2_fused_ag_only.err
for f in r02_bench_n2_fused r02_bench_n2_fused_ag_only; do python - "$o/$f.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 1), "tok/s", round(d["value"]), "gemm TF/s", round(d["roofline"]["achieved"]), d["config"]["optimizer"][:200])
except Exception as e:
    print(sys.argv[1], "unreadable:", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
done
