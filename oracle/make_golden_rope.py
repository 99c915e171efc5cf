"""Generates tests/golden/rope_tables.pt from HF's own LlamaRotaryEmbedding (the arithmetic the reference gets through
`LlamaModel`, metamorph_llama.py:349 -> HF modeling_llama.py rotary classes; transformers 5.5 here, 4.45 pinned — the
RoPE definitions are unchanged between them): cos/sin at selected positions for the plain LLaMA-3 setting
(theta 5e5) and for LLaMA-3.1's `llama3` frequency scaling.

    python -m oracle.make_golden_rope
"""
import os

import torch
from transformers import LlamaConfig
from transformers.models.llama.modeling_llama import LlamaRotaryEmbedding

SCALING = {"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
           "original_max_position_embeddings": 8192}
POSITIONS = [0, 1, 2, 3, 7, 63, 64, 127, 511, 1000, 4095, 4096, 8191, 8192, 12345, 20000, 65535, 131071]


def tables(scaling):
    kw = dict(hidden_size=4096, num_attention_heads=32, num_key_value_heads=8, head_dim=128, rope_theta=500000.0,
              max_position_embeddings=131072)
    try:
        cfg = LlamaConfig(**kw, rope_scaling=dict(scaling) if scaling else None)
    except Exception:
        cfg = LlamaConfig(**kw)
    if scaling and getattr(cfg, "rope_parameters", None) is not None:
        cfg.rope_parameters = {**dict(scaling), "rope_theta": 500000.0}
    rot = LlamaRotaryEmbedding(cfg)
    pos = torch.tensor(POSITIONS)[None]
    cos, sin = rot(torch.zeros(1, len(POSITIONS), 128, dtype=torch.float32), pos)
    return cos[0, :, :64].contiguous(), sin[0, :, :64].contiguous(), rot.inv_freq.clone()


def main():
    out = {"positions": torch.tensor(POSITIONS), "scaling": SCALING}
    for name, sc in (("plain", None), ("llama3", SCALING)):
        cos, sin, inv = tables(sc)
        out[name] = {"cos": cos, "sin": sin, "inv_freq": inv}
    assert not torch.equal(out["plain"]["inv_freq"], out["llama3"]["inv_freq"]), "scaling was not applied"
    path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "rope_tables.pt")
    torch.save(out, path)
    print("wrote", os.path.abspath(path), os.path.getsize(path))


if __name__ == "__main__":
    main()
