"""Does the weight row stride limit the weight-streaming GEMM? N = 4096 rows (256 CTAs of 16 rows), K and the row
pitch varied; CUDA-graph replay over distinct buffers larger than L2."""
import sys

import torch

sys.path.insert(0, ".")
from metamorph_b200 import ops  # noqa: E402


def bench(N, K, pad=0, copies=10, iters=5, m=8):
    bufs = [torch.randn(N, K + pad, device="cuda").bfloat16() for _ in range(copies)]
    ws = [b[:, :K] for b in bufs]
    x = torch.randn(m, K, device="cuda").bfloat16()
    out = torch.empty((m, N), dtype=torch.bfloat16, device="cuda")
    for w in ws:
        ops.skinny_gemm(x, w, out=out)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for w in ws:
            ops.skinny_gemm(x, w, out=out)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (iters * copies)
    print(f"N={N:6d} K={K:6d} pitch={2*(K+pad):6d} B: {us:7.1f} us  {N*K*2/us/1e3:7.1f} GB/s", flush=True)


for K, pad in [(4096, 0), (8192, 0), (12288, 0), (14336, 0), (16384, 0), (14336, 64), (14336, 512), (14336, 2048),
               (8192, 64), (4096, 64)]:
    bench(4096, K, pad)
bench(2048, 14336)
bench(8192, 14336, copies=6)
