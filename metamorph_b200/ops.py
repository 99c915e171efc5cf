"""Python wrappers over the C-ABI kernels. Each wrapper validates shapes, allocates outputs with
torch (caller-owned memory) and enqueues the kernel on torch's current stream."""
from __future__ import annotations

from typing import Optional

import torch

from ._lib import call, c_float, c_int, ll, ptr, require_cuda, stream_ptr

EPI_STORE, EPI_BIAS, EPI_BIAS_GELU_ERF, EPI_BIAS_GELU_TANH, EPI_RESID, EPI_BIAS_RESID, EPI_SWIGLU = range(7)


def _ld(t: torch.Tensor) -> int:
    assert t.dim() == 2 and t.stride(1) == 1, "expected a row-major 2-D view"
    return t.stride(0)


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False,
         out: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
         resid: Optional[torch.Tensor] = None, aux: Optional[torch.Tensor] = None,
         epilogue: int = EPI_STORE, out_dtype: torch.dtype = torch.bfloat16,
         accumulate: bool = False, alpha: float = 1.0, force_bn: int = 0) -> torch.Tensor:
    """bf16 tensor-core GEMM (tcgen05). Operand storage:
         a_mn=False: a is [M, K] row-major;  a_mn=True: a is stored [K, M] row-major (A = a^T)
         b_mn=False: b is [N, K] row-major (C = A b^T, nn.Linear);  b_mn=True: b is [K, N] (C = A b)
    """
    require_cuda(a, b, out, bias, resid, aux)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    if a_mn:
        K, M = a.shape
    else:
        M, K = a.shape
    if b_mn:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    assert K == Kb, f"contraction mismatch {K} vs {Kb}"
    n_out = N // 2 if epilogue == EPI_SWIGLU else N
    if out is None:
        out = torch.empty((M, n_out), dtype=out_dtype, device=a.device)
    assert out.shape == (M, n_out) and out.dtype == out_dtype
    call("mm_gemm_bf16", ptr(a), ptr(b), ptr(out), ptr(bias), ptr(resid), ptr(aux),
         ll(M), ll(N), ll(K), ll(_ld(a)), ll(_ld(b)), ll(_ld(out)),
         ll(_ld(resid) if resid is not None else 0), ll(_ld(aux) if aux is not None else 0),
         c_int(int(a_mn)), c_int(int(b_mn)), c_int(epilogue),
         c_int(1 if out_dtype == torch.float32 else 0), c_int(int(accumulate)), c_float(alpha),
         c_int(force_bn), stream_ptr())
    return out
