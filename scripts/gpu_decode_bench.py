"""Standalone decode bench (BASELINE config 4) on one B200: same routine bench.py reports under "decode"."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from metamorph_b200 import synthetic  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    model = synthetic.build_model(synthetic.make_config(), device=dev)
    peaks = {}
    p = os.path.join(os.path.dirname(os.path.abspath(bench.__file__)), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            peaks = json.load(f)
    hb = peaks.get("hbm_gbs") or peaks.get("hbm_gbs_sustained")
    out = bench.decode_bench(model, dev, {"hbm_gbs": hb} if hb else {})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
