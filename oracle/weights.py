"""ORACLE (test infrastructure): deterministic synthetic weights in the reference's HF parameter
naming, generated from numpy's version-stable PCG64 stream so fixtures do not have to store them."""
from __future__ import annotations

import numpy as np
import torch

TINY = dict(hidden=256, layers=2, heads=2, kv_heads=1, head_dim=128, inter=512, vocab=128258,
            rms_eps=1e-5, rope_theta=500000.0, siglip_width=1152, siglip_inter=4304, siglip_layers=2,
            siglip_heads=16, image_size=384, image_tokens=64, max_len=4096, vision_coef=1.0)


def make_weights(cfg=TINY, seed: int = 0, bf16_round: bool = True):
    rng = np.random.default_rng(seed)

    def t(shape, std):
        x = torch.from_numpy((rng.standard_normal(shape, dtype=np.float32) * std))
        return x.bfloat16().float() if bf16_round else x

    H, I, V = cfg["hidden"], cfg["inter"], cfg["vocab"]
    nq, nkv, dh = cfg["heads"], cfg["kv_heads"], cfg["head_dim"]
    p = {}
    p["model.embed_tokens.weight"] = t((V, H), 0.02)
    p["lm_head.weight"] = t((V, H), 0.02)
    p["model.norm.weight"] = 1 + t((H,), 0.05)
    for i in range(cfg["layers"]):
        q = f"model.layers.{i}."
        p[q + "input_layernorm.weight"] = 1 + t((H,), 0.05)
        p[q + "post_attention_layernorm.weight"] = 1 + t((H,), 0.05)
        p[q + "self_attn.q_proj.weight"] = t((nq * dh, H), 0.05)
        p[q + "self_attn.k_proj.weight"] = t((nkv * dh, H), 0.05)
        p[q + "self_attn.v_proj.weight"] = t((nkv * dh, H), 0.05)
        p[q + "self_attn.o_proj.weight"] = t((H, nq * dh), 0.05)
        p[q + "mlp.gate_proj.weight"] = t((I, H), 0.05)
        p[q + "mlp.up_proj.weight"] = t((I, H), 0.05)
        p[q + "mlp.down_proj.weight"] = t((H, I), 0.05)
    C, CI = cfg["siglip_width"], cfg["siglip_inter"]
    for name, din, dout in (("model.mm_projector.0", C, H), ("model.mm_projector.2", H, H),
                            ("vision_head.0", H, H), ("vision_head.2", H, C)):
        p[name + ".weight"] = t((dout, din), 1.0 / np.sqrt(din))
        p[name + ".bias"] = t((dout,), 0.02)
    p["model.vision_proj.weight"] = t((H, 4096), 0.01)
    p["model.vision_proj.bias"] = t((H,), 0.01)
    tp = "model.vision_tower.vision_tower."
    P = (cfg["image_size"] // 14) ** 2
    p[tp + "embeddings.patch_embedding.weight"] = t((C, 3, 14, 14), 1.0 / np.sqrt(588))
    p[tp + "embeddings.patch_embedding.bias"] = t((C,), 0.02)
    p[tp + "embeddings.position_embedding.weight"] = t((P, C), 0.02)
    for i in range(cfg["siglip_layers"]):
        q = f"{tp}encoder.layers.{i}."
        for ln in ("layer_norm1", "layer_norm2"):
            p[q + ln + ".weight"] = 1 + t((C,), 0.05)
            p[q + ln + ".bias"] = t((C,), 0.02)
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            p[q + f"self_attn.{nm}.weight"] = t((C, C), 1.0 / np.sqrt(C))
            p[q + f"self_attn.{nm}.bias"] = t((C,), 0.02)
        p[q + "mlp.fc1.weight"] = t((CI, C), 1.0 / np.sqrt(C))
        p[q + "mlp.fc1.bias"] = t((CI,), 0.02)
        p[q + "mlp.fc2.weight"] = t((C, CI), 1.0 / np.sqrt(CI))
        p[q + "mlp.fc2.bias"] = t((C,), 0.02)
    p[tp + "post_layernorm.weight"] = torch.ones(C)
    p[tp + "post_layernorm.bias"] = torch.zeros(C)
    return p


def make_batch(cfg=TINY, seed: int = 1):
    """3 samples (SURVEY.md §8c cross-check shape): prompt-image + answer-image; text-only (dummy
    image); shorter answer-image sample (right padded). Returns input_ids, attention_mask, labels, images."""
    rng = np.random.default_rng(seed)
    IMG, START, END = -200, 128256, 128257

    def txt(n):
        return rng.integers(0, 128000, size=n).tolist()

    s0 = [128000] + txt(9) + [START, IMG, END] + txt(7)
    l0 = [-100] * len(s0)
    a0 = txt(5) + [START, IMG, END] + txt(4)
    s0, l0 = s0 + a0, l0 + a0
    s1 = [128000] + txt(30)
    l1 = [-100] * 12 + s1[12:]
    s2 = [128000] + txt(6)
    l2 = [-100] * len(s2)
    a2 = txt(3) + [START, IMG, END] + txt(2)
    s2, l2 = s2 + a2, l2 + a2
    L = max(len(s0), len(s1), len(s2))
    ids = torch.full((3, L), 128001, dtype=torch.long)
    labs = torch.full((3, L), -100, dtype=torch.long)
    mask = torch.zeros((3, L), dtype=torch.bool)
    for b, (s, l) in enumerate(((s0, l0), (s1, l1), (s2, l2))):
        ids[b, :len(s)] = torch.tensor(s)
        labs[b, :len(l)] = torch.tensor(l)
        mask[b, :len(s)] = True
    n_img = 4  # s0: 2 images, s1: dummy, s2: 1
    images = torch.from_numpy(rng.standard_normal((n_img, 3, cfg["image_size"], cfg["image_size"]),
                                                  dtype=np.float32)).bfloat16().float()
    images[2] = 0  # dummy image of the text-only sample (train.py:1239-1242)
    return ids, mask, labs, images
