"""CPU: RoPE tables — the oracle restatement (oracle/restatement.py::rope_cos_sin) and the product's host table
builder (engine/llama.py::rope_tables) against HF's own LlamaRotaryEmbedding (tests/golden/rope_tables.pt, made by
oracle/make_golden_rope.py), for plain LLaMA-3 (theta 5e5) and LLaMA-3.1's `llama3` frequency scaling."""
import os

import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "rope_tables.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


@pytest.mark.parametrize("name", ["plain", "llama3"])
def test_oracle_rope_matches_hf(gold, name):
    from oracle.restatement import rope_cos_sin
    pos = gold["positions"]
    cos, sin = rope_cos_sin(128, 500000.0, pos[None], gold["scaling"] if name == "llama3" else None)
    cos, sin = cos.reshape(len(pos), -1)[:, :64], sin.reshape(len(pos), -1)[:, :64]
    torch.testing.assert_close(cos, gold[name]["cos"], rtol=0, atol=2e-6)
    torch.testing.assert_close(sin, gold[name]["sin"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("name", ["plain", "llama3"])
def test_product_rope_tables_match_hf(gold, name):
    from metamorph_b200.engine.llama import LlamaDims, rope_tables
    pos = gold["positions"]
    dims = LlamaDims(hidden=4096, n_layers=1, n_heads=32, n_kv_heads=8, head_dim=128, intermediate=14336, vocab=8,
                     rope_theta=500000.0, rope_scaling=gold["scaling"] if name == "llama3" else None)
    n = int(pos.max()) + 1
    cos, sin = rope_tables(dims, n, "cpu", round_bf16=False)
    # large angles lose absolute precision in fp32 (pos * inv_freq ~ 1e5): HF computes the same product in fp32
    torch.testing.assert_close(cos[pos], gold[name]["cos"], rtol=0, atol=5e-3)
    torch.testing.assert_close(sin[pos], gold[name]["sin"], rtol=0, atol=5e-3)
    small = pos < 8192
    torch.testing.assert_close(cos[pos][small], gold[name]["cos"][small], rtol=0, atol=2e-4)
    torch.testing.assert_close(sin[pos][small], gold[name]["sin"][small], rtol=0, atol=2e-4)
    # what the kernels consume: HF casts cos/sin to the activation dtype (bf16)
    cos16, _ = rope_tables(dims, n, "cpu", round_bf16=True)
    exact = (cos16[pos][small] == gold[name]["cos"][small].bfloat16().float()).float().mean()
    assert exact > 0.99
