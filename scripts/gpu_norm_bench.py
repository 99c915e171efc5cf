"""HBM-bound elementwise kernels of the train step at its shapes (M = 16384 rows, H = 4096): time per launch and the
fraction of the measured HBM peak, cycling through buffers larger than L2 like the real step."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from metamorph_b200 import ops  # noqa: E402

M, H = 16384, 4096
peak = 6579.6
try:
    with open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) as f:
        peak = float(json.load(f).get("hbm_gbs", peak))
except Exception:  # noqa: BLE001
    pass


def timed(fn, n_bufs, iters=10):
    for i in range(n_bufs):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(iters):
        for i in range(n_bufs):
            fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (iters * n_bufs)


nb = 4
xs = [torch.randn(M, H, device="cuda").bfloat16() for _ in range(nb)]
dys = [torch.randn(M, H, device="cuda").bfloat16() for _ in range(nb)]
res = [torch.randn(M, H, device="cuda").bfloat16() for _ in range(nb)]
outs = [torch.empty(M, H, device="cuda", dtype=torch.bfloat16) for _ in range(nb)]
w = torch.randn(H, device="cuda").bfloat16()
dw = torch.zeros(H, device="cuda", dtype=torch.float32)
us = timed(lambda i: ops.rmsnorm(xs[i], w, 1e-5, out=outs[i]), nb)
print(f"rmsnorm_fwd  {us:7.1f} us  {2 * M * H * 2 / us / 1e3:7.1f} GB/s  {2 * M * H * 2 / us / 1e3 / peak:.2f} of HBM peak")
us = timed(lambda i: ops.rmsnorm_bwd(dys[i], xs[i], w, 1e-5, dres_in=res[i], dw_accum=dw, out=outs[i]), nb)
print(f"rmsnorm_bwd  {us:7.1f} us  {4 * M * H * 2 / us / 1e3:7.1f} GB/s  {4 * M * H * 2 / us / 1e3 / peak:.2f} of HBM peak")
W = 6144
qkv = [torch.randn(M, W, device="cuda").bfloat16() for _ in range(nb)]
pos = (torch.arange(M, device="cuda", dtype=torch.int32) % 4096).contiguous()
inv = 1.0 / (500000.0 ** (torch.arange(0, 128, 2).float() / 128))
ang = torch.arange(4096).float()[:, None] * inv[None]
cos, sin = ang.cos().cuda().contiguous(), ang.sin().cuda().contiguous()
us = timed(lambda i: ops.rope_(qkv[i], pos, cos, sin, 40, 128), nb)
print(f"rope         {us:7.1f} us  {2 * M * 40 * 128 * 2 / us / 1e3:7.1f} GB/s  {2 * M * 40 * 128 * 2 / us / 1e3 / peak:.2f} of HBM peak")
