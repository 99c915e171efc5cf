"""Weight packing between the reference (HF) parameter layout and the fused device layout.

The SwiGLU GEMM epilogue (csrc/gemm_tcgen05.cu, EPI_SWIGLU) needs gate and up columns of the same
intermediate channel inside one 32-column accumulator chunk, so W_gate/W_up [I, H] are stored as one
[2I, H] matrix whose rows alternate in blocks of 16: rows [32b, 32b+16) = gate[16b:16b+16],
rows [32b+16, 32b+32) = up[16b:16b+16].  (HF names: model.layers.{i}.mlp.{gate,up}_proj.weight.)
"""
from __future__ import annotations

import torch

BLK = 16


def interleave_gate_up(gate: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    I, H = gate.shape
    assert up.shape == (I, H) and I % BLK == 0
    return torch.stack([gate.view(I // BLK, BLK, H), up.view(I // BLK, BLK, H)], dim=1).reshape(2 * I, H).contiguous()


def deinterleave_gate_up(gu: torch.Tensor):
    I2, H = gu.shape
    I = I2 // 2
    t = gu.view(I // BLK, 2, BLK, H)
    return t[:, 0].reshape(I, H).contiguous(), t[:, 1].reshape(I, H).contiguous()
