"""Generates tests/golden/reference_state_keys.json: names, shapes and dtypes of the REFERENCE model's own
`state_dict()` (metamorph.model.MetaMorphLlamaForCausalLM imported from /root/reference, tiny LLaMA dims, real-width
2-layer SigLIP) — the checkpoint contract tests/test_checkpoint.py holds the product's saved files to.
The tower is built with the survey's shim (SiglipVisionModel wrapper), so its keys carry an extra `vision_model.` level
that the real reference (`model.vision_model`, siglip_encoder.py:122) does not have; it is stripped here.

    PYTHONDONTWRITEBYTECODE=1 WANDB_MODE=disabled python -m oracle.make_golden_checkpoint
"""
import json
import os

from oracle.ref_model import activate, build_reference
from oracle.weights import TINY, make_weights


def main():
    activate("/root/reference")
    model = build_reference(TINY, make_weights(TINY))
    tp = "model.vision_tower.vision_tower."
    out = {}
    for k, v in model.state_dict().items():
        if k.startswith(tp + "vision_model."):
            k = tp + k[len(tp + "vision_model."):]
        out[k] = {"shape": list(v.shape), "dtype": str(v.dtype).replace("torch.", "")}
    path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "reference_state_keys.json")
    with open(path, "w") as fh:
        json.dump({"config": "oracle.weights.TINY", "state": out}, fh, indent=1, sort_keys=True)
        fh.write("\n")
    print("wrote", os.path.abspath(path), len(out), "tensors")


if __name__ == "__main__":
    main()
