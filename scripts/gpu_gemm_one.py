"""Launch one tcgen05 GEMM shape a few times (for `ncu --set full -s 2 -c 1`)."""
import sys

import torch

sys.path.insert(0, ".")
from metamorph_b200 import ops  # noqa: E402

M, N, K = (int(x) for x in sys.argv[1:4])
a = torch.randn(M, K, device="cuda").bfloat16()
b = torch.randn(N, K, device="cuda").bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    ops.gemm(a, b, out=out)
torch.cuda.synchronize()
print("done")
