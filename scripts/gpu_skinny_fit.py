"""time = a + bytes / BW fit of the weight-streaming GEMM kernels (graph replay over distinct buffers > L2)."""
import sys

import torch

sys.path.insert(0, ".")
from metamorph_b200 import ops  # noqa: E402


def bench(N, K, epi, copies, iters=5, m=8):
    ws = [torch.randn(N, K, device="cuda").bfloat16() for _ in range(copies)]
    x = torch.randn(m, K, device="cuda").bfloat16()
    res = torch.randn(m, N, device="cuda").bfloat16()
    out = torch.empty((m, N // 2 if epi == ops.SK_SWIGLU else N), dtype=torch.bfloat16, device="cuda")
    for w in ws:
        ops.skinny_gemm(x, w, resid=res if epi == ops.SK_RESID else None, epilogue=epi, out=out)
    torch.cuda.synchronize()
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
    return e0.elapsed_time(e1) * 1e3 / (iters * copies)


for name, K, epi, Ns in (("swiglu (32-row TMA kernel)", 4096, ops.SK_SWIGLU, (3584, 7168, 14336, 28672, 57344)),
                         ("store K=4096 (16-row TMA kernel)", 4096, ops.SK_STORE, (1024, 2048, 4096, 8192, 16384)),
                         ("resid K=14336 (16-row TMA kernel)", 14336, ops.SK_RESID, (1024, 2048, 4096, 8192))):
    pts = []
    for N in Ns:
        mb = N * K * 2 / 1e6
        copies = max(3, min(16, int(1500 / mb)))
        us = bench(N, K, epi, copies)
        pts.append((mb, us))
        print(f"{name}: N={N:6d}  {mb:7.1f} MB  {us:7.2f} us  {mb / us * 1e3 / 1e3:6.2f} TB/s", flush=True)
    (x0, y0), (x1, y1) = pts[-2], pts[-1]
    slope = (y1 - y0) / (x1 - x0)
    print(f"   -> last two points: {1 / slope / 1e3 * 1e3:.2f} TB/s marginal, intercept {y1 - slope * x1:.1f} us", flush=True)
