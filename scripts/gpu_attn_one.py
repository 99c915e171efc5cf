"""One forward + backward of the tcgen05 attention at the train-step shape (for `ncu -k regex:flash_`)."""
import math
import sys

import torch

sys.path.insert(0, ".")
from metamorph_b200 import ops  # noqa: E402

B, T, Hq, Hkv, d = 4, 4096, 32, 8, 128
qkv = (torch.randn(B * T, (Hq + 2 * Hkv) * d, device="cuda") * 0.5).bfloat16()
q, k, v = qkv[:, :Hq * d], qkv[:, Hq * d:(Hq + Hkv) * d], qkv[:, (Hq + Hkv) * d:]
scale = 1 / math.sqrt(d)
dqkv = torch.empty_like(qkv)
for _ in range(2):
    o, lse = ops.attn_fwd(q, k, v, B, T, Hq, Hkv, d, True, scale)
    ops.attn_bwd(q, k, v, o, torch.randn_like(o), lse, dqkv[:, :Hq * d], dqkv[:, Hq * d:(Hq + Hkv) * d],
                 dqkv[:, (Hq + Hkv) * d:], B, T, Hq, Hkv, d, scale)
torch.cuda.synchronize()
print("done")
