"""ORACLE (test infrastructure): deterministic synthetic weights in the reference's HF parameter
naming, generated from numpy's version-stable PCG64 stream so fixtures do not have to store them."""
from __future__ import annotations

import numpy as np
import torch

TINY = dict(hidden=256, layers=2, heads=2, kv_heads=1, head_dim=128, inter=512, vocab=128258,
            rms_eps=1e-5, rope_theta=500000.0, siglip_width=1152, siglip_inter=4304, siglip_layers=2,
            siglip_heads=16, image_size=384, image_tokens=64, max_len=4096, vision_coef=1.0)


# Real-width cases (LLaMA-3-8B layer dims, real vocab, real-width SigLIP at depth 2): the kernel combination the
# benchmark runs (32/8 heads with GQA group 4, I=14336 SwiGLU interleave, 2-CTA GEMM path, V=128258 compaction,
# 32 key tiles) at a depth the reference finishes on CPU in about a minute (oracle/make_golden_realwidth.py).
REAL_A = dict(TINY, hidden=4096, layers=1, heads=32, kv_heads=8, inter=14336, w_std=1.0 / 64, w_std_down=0.00835)
REAL_B = dict(REAL_A, layers=2)


def make_weights(cfg=TINY, seed: int = 0, bf16_round: bool = True):
    rng = np.random.default_rng(seed)
    ws, wsd = cfg.get("w_std", 0.05), cfg.get("w_std_down", cfg.get("w_std", 0.05))

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
        p[q + "self_attn.q_proj.weight"] = t((nq * dh, H), ws)
        p[q + "self_attn.k_proj.weight"] = t((nkv * dh, H), ws)
        p[q + "self_attn.v_proj.weight"] = t((nkv * dh, H), ws)
        p[q + "self_attn.o_proj.weight"] = t((H, nq * dh), ws)
        p[q + "mlp.gate_proj.weight"] = t((I, H), ws)
        p[q + "mlp.up_proj.weight"] = t((I, H), ws)
        p[q + "mlp.down_proj.weight"] = t((H, I), wsd)
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


def interleaved_sample(rng, total, n_prompt_img, n_answer_img, prompt_frac=0.55):
    """One synthetic sample of exactly `total` interleaved positions (each image = <image_start>, 64 visual tokens,
    <image_end> = 66 positions, 3 input ids): BOS, prompt text with the prompt-side images (labels -100), then the answer
    with the answer-side images (labels = ids). Returns (input_ids, labels) as lists."""
    IMG, START, END = -200, 128256, 128257

    def txt(n):
        return rng.integers(0, 128000, size=n).tolist()

    n_img = n_prompt_img + n_answer_img
    n_text = total - 66 * n_img - 1
    n_prompt_text = int(n_text * prompt_frac)
    segs_p = np.diff(np.linspace(0, n_prompt_text, n_prompt_img + 2).astype(int)).tolist()
    segs_a = np.diff(np.linspace(0, n_text - n_prompt_text, n_answer_img + 2).astype(int)).tolist()
    s = [128000] + txt(segs_p[0])
    for i in range(n_prompt_img):
        s += [START, IMG, END] + txt(segs_p[i + 1])
    l = [-100] * len(s)
    a = txt(segs_a[0])
    for i in range(n_answer_img):
        a += [START, IMG, END] + txt(segs_a[i + 1])
    return s + a, l + a


def make_batch_real(case: str, cfg=REAL_A, seed: int = 11):
    """Real-width batches (B=2). case "A": sample 0 is EXACTLY 4096 positions after interleaving (2 prompt + 2 answer
    images), sample 1 is ragged (2501 positions, 1 + 1 images) and right-padded. case "B": T = 1501 (T % 4 != 0, the
    shape that used to route the attention backward to a fallback kernel): sample 0 has 1 + 1 images, sample 1 is text
    only (dummy all-zero image, train.py:1239-1242)."""
    rng = np.random.default_rng(seed + (0 if case == "A" else 1))

    def sample(total, n_p, n_a):
        return interleaved_sample(rng, total, n_p, n_a)

    if case == "A":
        samples, n_img = [sample(4096, 2, 2), sample(2501, 1, 1)], 6
    else:
        s1 = [128000] + rng.integers(0, 128000, size=906).tolist()
        samples, n_img = [sample(1501, 1, 1), (s1, [-100] * 300 + s1[300:])], 3
    L = max(len(s) for s, _ in samples)
    ids = torch.full((2, L), 128001, dtype=torch.long)
    labs = torch.full((2, L), -100, dtype=torch.long)
    mask = torch.zeros((2, L), dtype=torch.bool)
    for b, (s, l) in enumerate(samples):
        ids[b, :len(s)] = torch.tensor(s)
        labs[b, :len(l)] = torch.tensor(l)
        mask[b, :len(s)] = True
    images = torch.from_numpy(rng.standard_normal((n_img, 3, cfg["image_size"], cfg["image_size"]),
                                                  dtype=np.float32)).bfloat16().float()
    if case == "B":
        images[2] = 0
    return ids, mask, labs, images


def with_sparse_lm_head(W, k: int = 16, seed: int = 5):
    """Copy of a weight dict whose lm_head keeps only `k` (seeded) vocabulary rows, all others zero. With 128 k random
    rows the top-1/top-2 logit gap of a random model is ~0.2 % — below bf16 noise, so free-running decodes of a bf16
    implementation cannot be held to the fp32 reference token by token. With k live rows (logit exactly 0 everywhere
    else) the gap is two orders of magnitude larger while the decode path (lm_head GEMM over the full vocabulary,
    argmax, embedding feedback, image mode) is exercised unchanged. Returns (weights, live token ids)."""
    rng = np.random.default_rng(seed)
    ids = np.sort(rng.choice(128000, size=k, replace=False))
    out = dict(W)
    lm = torch.zeros_like(W["lm_head.weight"])
    lm[ids] = W["lm_head.weight"][ids]
    out["lm_head.weight"] = lm
    return out, ids.tolist()
