"""The kernels of one KV-cached decode step, shared by DecodeEngine (one batch to completion) and
ContinuousBatcher (a stream of requests): decoder stack on the weight-streaming GEMMs, then the heads
(final norm, vision head -> projector feedback branch, lm_head, argmax). Reference arithmetic:
metamorph_llama.py:363-377, 482-490 (decoding branch of llm_forward) and 526-582 (greedy_decode loop body)."""
from __future__ import annotations

import torch

from .. import ops


def decoder_stack_step(layers, x, kc, vc, cur_pos, stack):
    """x [B, H] -> [B, H] through all layers; K/V of the fed position are appended to kc/vc [L, B, Hkv, Tmax, dh]."""
    d = stack.dims
    Hq, Hkv, dh = d.n_heads, d.n_kv_heads, d.head_dim
    for i, w in enumerate(layers):
        n1 = ops.rmsnorm(x, w.ln1, d.rms_eps)
        qkv = ops.skinny_gemm(n1, w.wqkv)
        attn = ops.decode_attn(qkv, kc[i], vc[i], cur_pos, stack.cos, stack.sin, Hq, Hkv, dh, stack.scale)
        hmid = ops.skinny_gemm(attn, w.wo, resid=x, epilogue=ops.SK_RESID)
        n2 = ops.rmsnorm(hmid, w.ln2, d.rms_eps)
        act = ops.skinny_gemm(n2, w.wgu, epilogue=ops.SK_SWIGLU)
        x = ops.skinny_gemm(act, w.wd, resid=hmid, epilogue=ops.SK_RESID)
    return x


def decode_heads(m, h_pre_norm, in_image_mode, logits, V):
    """-> (argmax token [B] int32, pred_z [B, C] (normalised visual embedding), prediction [B, H] (its projection)).
    The image-mode branch is computed for every sequence and selected per sequence (graph friendly)."""
    inner = m.get_model()
    d = m.stack.dims
    hidden = ops.rmsnorm(h_pre_norm, inner.norm.weight.data, d.rms_eps)
    vh, pj = m.vision_head, inner.mm_projector
    z = ops.skinny_gemm(hidden, vh.fc1.weight.data, bias=vh.fc1.bias.data, epilogue=ops.SK_BIAS_GELU)
    z = ops.skinny_gemm(z, vh.fc2.weight.data, bias=vh.fc2.bias.data, epilogue=ops.SK_BIAS)
    pred_z = ops.l2norm_rows(z) if m.normalize_vision else z
    p1 = ops.skinny_gemm(pred_z, pj.fc1.weight.data, bias=pj.fc1.bias.data, epilogue=ops.SK_BIAS_GELU)
    prediction = ops.skinny_gemm(p1, pj.fc2.weight.data, bias=pj.fc2.bias.data, epilogue=ops.SK_BIAS)
    h_eff = torch.empty_like(hidden)
    ops.decode_select_hidden(in_image_mode, hidden, prediction, h_eff)
    ops.skinny_gemm(h_eff, m.lm_head.weight.data, out=logits[:, :V])
    tok = ops.argmax_rows(logits, V)
    return tok, pred_z, prediction
