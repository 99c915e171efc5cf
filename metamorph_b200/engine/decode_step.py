"""The kernels of one KV-cached decode step, shared by DecodeEngine (one batch to completion) and
ContinuousBatcher (a stream of requests): decoder stack on the weight-streaming GEMMs, then the heads
(final norm, vision head -> projector feedback branch, lm_head, argmax). Reference arithmetic:
metamorph_llama.py:363-377, 482-490 (decoding branch of llm_forward) and 526-582 (greedy_decode loop body)."""
from __future__ import annotations

import os

import torch

from .. import ops


# MM_DECODE_PREFETCH=0 switches the L2 hints of the decode chain off (A/B measurements)
PREFETCH = os.environ.get("MM_DECODE_PREFETCH", "1") != "0"


def decoder_stack_step(layers, x, kc, vc, cur_pos, stack, after=None):
    """x [B, H] -> [B, H] through all layers; K/V of the fed position are appended to kc/vc [L, B, Hkv, Tmax, dh].
    Every kernel of the chain names the weight matrix its successor will stream (`prefetch=`): the head of that stream
    is pulled into L2 across the kernel boundary (csrc/decode.cu l2_prefetch_share). `after`: the first weight matrix
    the caller streams after the stack (the heads)."""
    d = stack.dims
    Hq, Hkv, dh = d.n_heads, d.n_kv_heads, d.head_dim
    cap = ops.PREFETCH_CAP
    for i, w in enumerate(layers):
        nxt = layers[i + 1].wqkv if i + 1 < len(layers) else after
        pf = (lambda t, off=0, n=cap: (t, off, n) if (PREFETCH and t is not None) else None)
        n1 = ops.rmsnorm(x, w.ln1, d.rms_eps)
        qkv = ops.skinny_gemm(n1, w.wqkv, prefetch=pf(w.wo))
        attn = ops.decode_attn(qkv, kc[i], vc[i], cur_pos, stack.cos, stack.sin, Hq, Hkv, dh, stack.scale,
                               prefetch=pf(w.wgu))                                   # HBM is idle during attention
        hmid = ops.skinny_gemm(attn, w.wo, resid=x, epilogue=ops.SK_RESID, prefetch=pf(w.wgu, cap, cap // 2))
        n2 = ops.rmsnorm(hmid, w.ln2, d.rms_eps)
        act = ops.skinny_gemm(n2, w.wgu, epilogue=ops.SK_SWIGLU, prefetch=pf(w.wd))
        x = ops.skinny_gemm(act, w.wd, resid=hmid, epilogue=ops.SK_RESID, prefetch=pf(nxt))
    return x


def heads_first_weight(m):
    """The first weight matrix `decode_heads` streams (what the last decoder layer prefetches)."""
    return m.vision_head.fc1.weight.data


def decode_heads(m, h_pre_norm, in_image_mode, logits, V, next_step_first=None):
    """-> (argmax token [B] int32, pred_z [B, C] (normalised visual embedding), prediction [B, H] (its projection)).
    The image-mode branch is computed for every sequence and selected per sequence (graph friendly).
    `next_step_first`: the first weight matrix of the NEXT decode step (layer 0's qkv projection), prefetched into L2
    while lm_head's stream drains."""
    inner = m.get_model()
    d = m.stack.dims
    pf = (lambda t: t if PREFETCH else None)
    hidden = ops.rmsnorm(h_pre_norm, inner.norm.weight.data, d.rms_eps)
    vh, pj = m.vision_head, inner.mm_projector
    z = ops.skinny_gemm(hidden, vh.fc1.weight.data, bias=vh.fc1.bias.data, epilogue=ops.SK_BIAS_GELU,
                        prefetch=pf(vh.fc2.weight.data))
    z = ops.skinny_gemm(z, vh.fc2.weight.data, bias=vh.fc2.bias.data, epilogue=ops.SK_BIAS, prefetch=pf(pj.fc1.weight.data))
    pred_z = ops.l2norm_rows(z) if m.normalize_vision else z
    p1 = ops.skinny_gemm(pred_z, pj.fc1.weight.data, bias=pj.fc1.bias.data, epilogue=ops.SK_BIAS_GELU,
                         prefetch=pf(pj.fc2.weight.data))
    prediction = ops.skinny_gemm(p1, pj.fc2.weight.data, bias=pj.fc2.bias.data, epilogue=ops.SK_BIAS,
                                 prefetch=pf(m.lm_head.weight.data))
    h_eff = torch.empty_like(hidden)
    ops.decode_select_hidden(in_image_mode, hidden, prediction, h_eff)
    ops.skinny_gemm(h_eff, m.lm_head.weight.data, out=logits[:, :V], prefetch=pf(next_step_first))
    tok = ops.argmax_rows(logits, V)
    return tok, pred_z, prediction
