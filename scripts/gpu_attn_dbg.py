import math, os, sys, torch
sys.path.insert(0, ".")
from metamorph_b200 import ops
from scripts.gpu_attn_bench import timeit
B, T, Hq, Hkv, d = 4, 4096, 32, 8, 128
torch.manual_seed(0)
qkv = (torch.randn(B * T, (Hq + 2 * Hkv) * d, device="cuda") * 0.5).bfloat16()
q, k, v = qkv[:, :Hq * d], qkv[:, Hq * d:(Hq + Hkv) * d], qkv[:, (Hq + Hkv) * d:]
scale = 1 / math.sqrt(d)
o, lse = ops.attn_fwd(q, k, v, B, T, Hq, Hkv, d, True, scale)
dout = torch.randn_like(o)
dqkv = torch.zeros_like(qkv)
fl = 2.5 * 4 * B * Hq * T * T * d / 2
for dbg in (0, 1, 2, 4, 8, 3, 7, 15):
    os.environ["MM_ATTN_DBG"] = str(dbg)
    fn = lambda: ops.attn_bwd(q, k, v, o, dout, lse, dqkv[:, :Hq * d], dqkv[:, Hq * d:(Hq + Hkv) * d], dqkv[:, (Hq + Hkv) * d:], B, T, Hq, Hkv, d, scale, tc=True)
    ms = timeit(fn, iters=3)
    print(f"bwd_tc dbg={dbg:2d}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
