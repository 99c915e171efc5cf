"""Weight-streaming GEMM (decode) microbench: cycles through distinct weight buffers (> L2) like the real step."""
import sys

import torch

sys.path.insert(0, ".")
from metamorph_b200 import ops  # noqa: E402


def bench(N, K, epi=ops.SK_STORE, copies=12, iters=5, m=8):
    ws = [torch.randn(N, K, device="cuda").bfloat16() for _ in range(copies)]
    x = torch.randn(m, K, device="cuda").bfloat16()
    res = torch.randn(m, N, device="cuda").bfloat16()
    for w in ws:
        ops.skinny_gemm(x, w, resid=res if epi == ops.SK_RESID else None, epilogue=epi)
    torch.cuda.synchronize()
    # replay through a CUDA graph (as the decode engine does) so host-side launch cost is not measured
    out = torch.empty((m, N // 2 if epi == ops.SK_SWIGLU else N), dtype=torch.bfloat16, device="cuda")
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for w in ws:
            ops.skinny_gemm(x, w, resid=res if epi == ops.SK_RESID else None, epilogue=epi, out=out)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (iters * copies)
    print(f"skinny m={m:2d} N={N:6d} K={K:6d} epi={epi}: {us:7.1f} us  {N*K*2/us/1e3:7.1f} GB/s", flush=True)


for m in ([8, 16, 32] if "--batches" in sys.argv else [8]):
    bench(6144, 4096, m=m)
    bench(4096, 4096, ops.SK_RESID, m=m)
    bench(28672, 4096, ops.SK_SWIGLU, m=m)
    bench(4096, 14336, ops.SK_RESID, m=m)
    bench(128258, 4096, copies=3, m=m)

import os
