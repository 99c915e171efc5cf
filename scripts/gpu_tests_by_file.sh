#!/bin/bash
# Runs every GPU test file in its own process (a CUDA fault in one file must not take the others down) and collects the
# tails in gpurun_out/<tag>_tests.log.   usage: scripts/gpu_tests_by_file.sh <tag> [pytest args]
tag=${1:-run}; shift
mkdir -p gpurun_out
log=gpurun_out/${tag}_tests.log
: > "$log"
for f in tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_realwidth_gpu.py tests/test_trainer_gpu.py \
         tests/test_decode_gpu.py tests/test_preprocess_gpu.py tests/test_dp_nccl_gpu.py; do
  echo "===== $f" >> "$log"
  timeout ${MM_TEST_TIMEOUT:-420} python -m pytest "$f" -m gpu -q -rs -s "$@" 2>&1 | grep -v "^$" | tail -40 >> "$log"
done
grep -E "passed|failed|error" "$log" | tail -12
