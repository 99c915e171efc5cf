"""Standalone run of bench.preprocess_bench (SigLIP image pre-processing, host uint8 -> device fp32) on one B200."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402


def main():
    peaks = {}
    p = os.path.join(os.path.dirname(os.path.abspath(bench.__file__)), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            peaks = json.load(f)
    print(json.dumps(bench.preprocess_bench(torch.device("cuda", 0), peaks)))


if __name__ == "__main__":
    main()
