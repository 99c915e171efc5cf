"""ORACLE — test infrastructure, NOT product code.

CPU restatement (plain torch fp32 / Python loops) of the reference algorithm for the MetaMorph hot
path. Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this package; nothing under `metamorph_b200/` does.

Parity pinning: the reference ships no tests or golden vectors, so this
restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the build container by
`oracle/make_golden.py` (imports /root/reference read-only) and committed under `tests/golden/`;
`tests/test_oracle_golden.py` checks every function here against those fixtures.

Each function cites the reference lines it follows:
  interleave_reference        metamorph/model/metamorph_arch.py:245-425
  siglip_tower_forward        metamorph/model/multimodal_encoder/siglip_encoder.py:138-213
                              + HF modeling_siglip.py (SiglipVisionEmbeddings:116, SiglipAttention:252,
                                SiglipMLP:315, SiglipEncoderLayer:330)  [transformers==4.45.0 pinned,
                                un-vendored third-party: algorithm restated from its published source]
  mlp_gelu                    metamorph/model/multimodal_projector/builder.py:52-59,
                              language_model/metamorph_llama.py:252-256
  llama_forward               HF modeling_llama.py (LlamaRMSNorm:53, rotary:73-168, LlamaMLP:171,
                              LlamaAttention:225, LlamaDecoderLayer:292) called at metamorph_llama.py:349
  losses                      metamorph_llama.py:398-474
  greedy_decode_nocache       metamorph_llama.py:502-597 (+ decoding branch :363-377)
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200


# ------------------------------------------------------------------------------------------------
# index logic — token-by-token walk (deliberately structured differently from the product code)
# ------------------------------------------------------------------------------------------------
def interleave_reference(input_ids: List[List[int]], attention_mask: List[List[bool]],
                         labels: List[List[int]], n_images: int, image_len: int, max_len: int,
                         padding_side: str = "right", start_id: int = 128256):
    """Returns dict(rows=[[('t', id) | ('i', img, j) | ('p',)]], labels, image_positions, mask,
    position_ids, placeholder) with the semantics of metamorph_arch.py:259-423."""
    img_cursor = 0
    placeholder: List[int] = []
    seqs = []
    for ids_row, m_row, l_row in zip(input_ids, attention_mask, labels):
        ids = [t for t, m in zip(ids_row, m_row) if m]
        labs = [t for t, m in zip(l_row, m_row) if m]
        rows, out_l, out_p = [], [], []
        if IMAGE_TOKEN_INDEX not in ids:                                   # :275-284
            placeholder.append(img_cursor)
            img_cursor += 1
            seqs.append(([("t", t) for t in ids], list(labs), [0] * len(ids)))
            continue
        stopped = False
        prev_label_in_chunk: Optional[int] = None
        for t, lab in zip(ids, labs):
            if t != IMAGE_TOKEN_INDEX:
                if not stopped:
                    rows.append(("t", t)); out_l.append(lab); out_p.append(0)
                prev_label_in_chunk = lab
                continue
            if prev_label_in_chunk is None:                                # cur_labels_noim[i][-1] on empty chunk
                raise IndexError("empty text chunk before <image>")
            answer = prev_label_in_chunk == start_id                       # :317
            if len(rows) + image_len > max_len:                            # :324-326
                stopped = True
                placeholder.append(img_cursor)
            else:
                for j in range(image_len):
                    rows.append(("i", img_cursor, j)); out_l.append(IGNORE_INDEX); out_p.append(1 if answer else 0)
                if not answer:
                    placeholder.append(img_cursor)                         # :335
            img_cursor += 1
            prev_label_in_chunk = None
        seqs.append((rows, out_l, out_p))
    seqs = [(r[:max_len], l[:max_len], p[:max_len]) for r, l, p in seqs]    # :355-358
    T = max(len(r) for r, _, _ in seqs)
    out = dict(rows=[], labels=[], image_positions=[], mask=[], position_ids=[], placeholder=placeholder)
    for r, l, p in seqs:
        n = len(r)
        pad = T - n
        if padding_side == "left":                                         # :375-386
            out["rows"].append([("p",)] * pad + r)
            out["labels"].append([IGNORE_INDEX] * pad + l)
            out["image_positions"].append([0] * pad + p)
            out["mask"].append([False] * pad + [True] * n)
            out["position_ids"].append([0] * pad + list(range(n)))
        else:                                                              # :388-397
            out["rows"].append(r + [("p",)] * pad)
            out["labels"].append(l + [IGNORE_INDEX] * pad)
            out["image_positions"].append(p + [0] * pad)
            out["mask"].append([True] * n + [False] * pad)
            out["position_ids"].append(list(range(n)) + [0] * pad)
    out["targets"] = [i for i in range(n_images) if i not in set(placeholder)]  # :415-423
    return out


def materialize_embeds(rows, embed_w: torch.Tensor, img_feats: torch.Tensor) -> torch.Tensor:
    """rows from interleave_reference -> [B, T, H] (padding rows are zeros, :376-391)."""
    B, T, H = len(rows), len(rows[0]), embed_w.shape[1]
    out = torch.zeros(B, T, H, dtype=embed_w.dtype)
    for b in range(B):
        for t, r in enumerate(rows[b]):
            if r[0] == "t":
                out[b, t] = embed_w[r[1]]
            elif r[0] == "i":
                out[b, t] = img_feats[r[1], r[2]]
    return out


# ------------------------------------------------------------------------------------------------
# SigLIP tower (fp32)
# ------------------------------------------------------------------------------------------------
def siglip_tower_forward(p: Dict[str, torch.Tensor], images: torch.Tensor, n_layers: int, n_heads: int,
                         out_tokens: int, normalize: bool, eps: float = 1e-6, prefix: str = "") -> torch.Tensor:
    """images [N,3,S,S] fp32 -> [N, out_tokens, C]. hidden_states[-1] (no post_layernorm), bilinear
    27x27 -> sqrt(out_tokens)^2 in fp32, optional L2 normalise (siglip_encoder.py:151-163, 206-208)."""
    g = lambda k: p[prefix + k].float()
    x = F.conv2d(images.float(), g("embeddings.patch_embedding.weight"), g("embeddings.patch_embedding.bias"), stride=14)
    x = x.flatten(2).transpose(1, 2) + g("embeddings.position_embedding.weight")[None]
    N, P, C = x.shape
    dh = C // n_heads
    for i in range(n_layers):
        q = f"encoder.layers.{i}."
        h = F.layer_norm(x, (C,), g(q + "layer_norm1.weight"), g(q + "layer_norm1.bias"), eps)
        qq = F.linear(h, g(q + "self_attn.q_proj.weight"), g(q + "self_attn.q_proj.bias")).view(N, P, n_heads, dh).transpose(1, 2)
        kk = F.linear(h, g(q + "self_attn.k_proj.weight"), g(q + "self_attn.k_proj.bias")).view(N, P, n_heads, dh).transpose(1, 2)
        vv = F.linear(h, g(q + "self_attn.v_proj.weight"), g(q + "self_attn.v_proj.bias")).view(N, P, n_heads, dh).transpose(1, 2)
        a = torch.softmax(qq @ kk.transpose(-1, -2) * dh ** -0.5, dim=-1) @ vv
        a = a.transpose(1, 2).reshape(N, P, C)
        x = x + F.linear(a, g(q + "self_attn.out_proj.weight"), g(q + "self_attn.out_proj.bias"))
        h = F.layer_norm(x, (C,), g(q + "layer_norm2.weight"), g(q + "layer_norm2.bias"), eps)
        h = F.gelu(F.linear(h, g(q + "mlp.fc1.weight"), g(q + "mlp.fc1.bias")), approximate="tanh")
        x = x + F.linear(h, g(q + "mlp.fc2.weight"), g(q + "mlp.fc2.bias"))
    if P != out_tokens:
        s, t = int(math.sqrt(P)), int(math.sqrt(out_tokens))
        x = x.view(N, s, s, C).permute(0, 3, 1, 2)
        x = F.interpolate(x, size=(t, t), mode="bilinear", align_corners=False)
        x = x.permute(0, 2, 3, 1).flatten(1, 2)
    if normalize:
        x = F.normalize(x, p=2, dim=-1)
    return x


def mlp_gelu(p: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor) -> torch.Tensor:
    h = F.gelu(F.linear(x, p[prefix + "0.weight"].float(), p[prefix + "0.bias"].float()))
    return F.linear(h, p[prefix + "2.weight"].float(), p[prefix + "2.bias"].float())


# ------------------------------------------------------------------------------------------------
# LLaMA (fp32)
# ------------------------------------------------------------------------------------------------
def rope_cos_sin(head_dim: int, theta: float, positions: torch.Tensor, scaling: Optional[dict] = None):
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
    if scaling and scaling.get("rope_type", scaling.get("type")) == "llama3":
        f, lo, hi, old = (scaling["factor"], scaling["low_freq_factor"], scaling["high_freq_factor"],
                          scaling["original_max_position_embeddings"])
        wl = 2 * math.pi / inv
        inv_l = torch.where(wl > old / lo, inv / f, inv)
        sm = (old / wl - lo) / (hi - lo)
        mid = (1 - sm) * inv_l / f + sm * inv_l
        inv = torch.where(~(wl < old / hi) & ~(wl > old / lo), mid, inv_l)
    ang = positions.float()[..., None] * inv
    ang = torch.cat([ang, ang], dim=-1)
    return ang.cos(), ang.sin()


def llama_forward(p: Dict[str, torch.Tensor], x: torch.Tensor, position_ids: torch.Tensor,
                  key_mask: torch.Tensor, n_layers: int, n_heads: int, n_kv: int, eps: float, theta: float,
                  scaling: Optional[dict] = None, prefix: str = "model.") -> torch.Tensor:
    """x [B,T,H] fp32, key_mask [B,T] bool (True = attend). Returns the final-norm output [B,T,H]."""
    B, T, H = x.shape
    dh = p[prefix + "layers.0.self_attn.q_proj.weight"].shape[0] // n_heads
    cos, sin = rope_cos_sin(dh, theta, position_ids, scaling)        # [B,T,dh]
    cos, sin = cos[:, None], sin[:, None]

    def rms(v, w):
        return w.float() * (v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + eps))

    def rot(v):
        return torch.cat([-v[..., dh // 2:], v[..., :dh // 2]], dim=-1)

    causal = torch.tril(torch.ones(T, T, dtype=torch.bool))
    allow = causal[None, None] & key_mask[:, None, None, :]
    for i in range(n_layers):
        q_ = f"{prefix}layers.{i}."
        h = rms(x, p[q_ + "input_layernorm.weight"])
        q = F.linear(h, p[q_ + "self_attn.q_proj.weight"].float()).view(B, T, n_heads, dh).transpose(1, 2)
        k = F.linear(h, p[q_ + "self_attn.k_proj.weight"].float()).view(B, T, n_kv, dh).transpose(1, 2)
        v = F.linear(h, p[q_ + "self_attn.v_proj.weight"].float()).view(B, T, n_kv, dh).transpose(1, 2)
        q, k = q * cos + rot(q) * sin, k * cos + rot(k) * sin
        k = k.repeat_interleave(n_heads // n_kv, dim=1)
        v = v.repeat_interleave(n_heads // n_kv, dim=1)
        s = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
        s = s.masked_fill(~allow, float("-inf"))
        a = torch.nan_to_num(torch.softmax(s, dim=-1)) @ v
        a = a.transpose(1, 2).reshape(B, T, n_heads * dh)
        x = x + F.linear(a, p[q_ + "self_attn.o_proj.weight"].float())
        h = rms(x, p[q_ + "post_attention_layernorm.weight"])
        g = F.linear(h, p[q_ + "mlp.gate_proj.weight"].float())
        u = F.linear(h, p[q_ + "mlp.up_proj.weight"].float())
        x = x + F.linear(F.silu(g) * u, p[q_ + "mlp.down_proj.weight"].float())
    return rms(x, p[prefix + "norm.weight"])


def losses(p: Dict[str, torch.Tensor], hidden: torch.Tensor, labels: torch.Tensor,
           image_positions: torch.Tensor, targets: Optional[torch.Tensor], vision_coef: float,
           use_vision_ar: bool = True):
    """metamorph_llama.py:398-474 -> (loss, loss_language, loss_image_ar, logits)."""
    logits = F.linear(hidden, p["lm_head.weight"].float())
    V = logits.shape[-1]
    loss_lang = F.cross_entropy(logits[:, :-1].reshape(-1, V), labels[:, 1:].reshape(-1), ignore_index=IGNORE_INDEX)
    sel = image_positions[:, 1:] != 0
    pred = hidden[:, :-1][sel]
    pred = F.normalize(mlp_gelu(p, "vision_head.", pred), p=2, dim=-1)
    if targets is None:
        loss_img = loss_lang
    else:
        tgt = targets.reshape(-1, targets.shape[-1]).float()
        try:
            loss_img = -F.cosine_similarity(tgt, pred, dim=-1).mean()
        except RuntimeError:
            loss_img = loss_lang
    loss = loss_lang + vision_coef * loss_img if use_vision_ar else loss_lang
    return loss, loss_lang, loss_img, logits


def full_forward(p, cfg, input_ids, attention_mask, labels, images):
    """End-to-end fp32 restatement of MetaMorphLlamaForCausalLM.forward for the scripts' config."""
    tp = "model.vision_tower.vision_tower."
    feats = siglip_tower_forward(p, images, cfg["siglip_layers"], cfg["siglip_heads"], cfg["image_tokens"],
                                 True, prefix=tp)
    ar = mlp_gelu(p, "model.mm_projector.", feats)
    plan = interleave_reference(input_ids.tolist(), attention_mask.bool().tolist(), labels.tolist(),
                                images.shape[0], cfg["image_tokens"], cfg["max_len"])
    x = materialize_embeds(plan["rows"], p["model.embed_tokens.weight"].float(), ar)
    pos = torch.tensor(plan["position_ids"])
    mask = torch.tensor(plan["mask"])
    hidden = llama_forward(p, x, pos, mask, cfg["layers"], cfg["heads"], cfg["kv_heads"], cfg["rms_eps"],
                           cfg["rope_theta"])
    tl = torch.tensor(plan["labels"])
    ip = torch.tensor(plan["image_positions"])
    tgt = feats[plan["targets"]] if plan["targets"] else feats[:0]
    loss, ll, li, logits = losses(p, hidden, tl, ip, tgt, cfg.get("vision_coef", 1.0))
    return dict(loss=loss, loss_language=ll, loss_image_ar=li, logits=logits, hidden=hidden, plan=plan,
                inputs_embeds=x, feats=feats, ar_feats=ar)


@torch.no_grad()
def greedy_decode_nocache(p, cfg, inputs_embeds: torch.Tensor, max_new_tokens: int, start_id=128256,
                          end_id=128257, eos=(128001, 128009), trace=None):
    """metamorph_llama.py:502-597: re-runs the whole growing prefix every step (batch 1).
    `trace` (a list) receives one dict per step: argmax token, top-1 minus top-2 logit, the largest |logit| (the
    scale bf16 noise is relative to: image-mode steps see the small projector output), and the mode flags BEFORE the
    step's state update (used by oracle/make_golden_decode_quirks.py to pick cases with a safe logit margin)."""
    x = inputs_embeds.float()
    ids, imgs = [], []
    in_image, n_img_tok, n_out = False, 0, 0
    ntok = cfg["image_tokens"]
    while True:
        T = x.shape[1]
        hidden = llama_forward(p, x, torch.arange(T)[None], torch.ones(1, T, dtype=torch.bool), cfg["layers"],
                               cfg["heads"], cfg["kv_heads"], cfg["rms_eps"], cfg["rope_theta"])
        pred_z = None
        if in_image:                                                          # :363-377
            pred_z = F.normalize(mlp_gelu(p, "vision_head.", hidden[:, -1]), p=2, dim=-1)
            hidden = hidden.clone()
            hidden[:, -1] = mlp_gelu(p, "model.mm_projector.", pred_z)
        logits = F.linear(hidden[:, -1], p["lm_head.weight"].float())
        tok = int(logits.argmax(-1))
        if trace is not None:
            top2 = logits[0].topk(2).values
            trace.append(dict(tok=tok, margin=float(top2[0] - top2[1]), scale=float(logits[0].abs().max()),
                              in_image=in_image, n_img_tok=n_img_tok))
        emb = p["model.embed_tokens.weight"].float()[tok][None, None]
        if not in_image and tok == start_id:
            in_image = True; ids.append(tok); x = torch.cat([x, emb], 1)
        elif in_image and n_img_tok < ntok:
            n_img_tok += 1; imgs.append(pred_z); x = torch.cat([x, hidden[:, -1:, :]], 1)
            if n_img_tok == ntok:
                in_image = False
        elif tok == end_id:
            in_image = False; n_img_tok = 0; ids.append(tok); x = torch.cat([x, emb], 1)
        else:
            x = torch.cat([x, emb], 1); ids.append(tok)
        n_out += 1
        if tok in eos or n_out > max_new_tokens:
            break
    return ids, (torch.cat(imgs, 0) if imgs else torch.zeros(0, 1152))
