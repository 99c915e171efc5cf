"""The four weight-streaming GEMM shapes of one decoder layer, each launched on two distinct weight buffers
(for `ncu --set full -k regex:skinny_gemm`: the second launch of a shape is the cold-cache steady-state one)."""
import sys

import torch

sys.path.insert(0, ".")
from metamorph_b200 import ops  # noqa: E402

m = 8
for N, K, epi in ((6144, 4096, ops.SK_STORE), (4096, 4096, ops.SK_RESID), (28672, 4096, ops.SK_SWIGLU), (4096, 14336, ops.SK_RESID)):
    x = torch.randn(m, K, device="cuda").bfloat16()
    res = torch.randn(m, N, device="cuda").bfloat16()
    for _ in range(2):
        w = torch.randn(N, K, device="cuda").bfloat16()
        ops.skinny_gemm(x, w, resid=res if epi == ops.SK_RESID else None, epilogue=epi)
torch.cuda.synchronize()
print("done")
