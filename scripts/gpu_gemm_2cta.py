"""Bring-up + A/B of the 2-CTA (cta_group::2) GEMM variant vs the 1-CTA kernel on B200."""
import sys
import time

import torch

sys.path.insert(0, ".")
from metamorph_b200 import ops  # noqa: E402


def check(M, N, K, a_mn, b_mn, epi=0):
    torch.manual_seed(1)
    a = torch.randn((K, M) if a_mn else (M, K), device="cuda").bfloat16()
    b = torch.randn((K, N) if b_mn else (N, K), device="cuda").bfloat16()
    ref = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, force_bn=256)
    out = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, force_bn=512)
    torch.cuda.synchronize()
    err = (out.float() - ref.float()).abs().max().item()
    print(f"2cta M={M} N={N} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)} max|2cta-1cta|={err:.5f} "
          f"{'OK' if err <= 1e-2 * ref.float().abs().max().item() else 'FAIL'}", flush=True)
    return err


def perf(M, N, K, bn, iters):
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(a, b, out=out, force_bn=bn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.gemm(a, b, out=out, force_bn=bn)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return 2 * M * N * K / ms / 1e9


def main():
    for (a_mn, b_mn) in [(False, False), (False, True), (True, True)]:
        for (M, N, K) in [(256, 256, 64), (512, 512, 256), (1024, 1152, 4096), (776, 520, 200)]:
            try:
                check(M, N, K, a_mn, b_mn)
            except Exception as e:  # noqa: BLE001
                print("EXC", M, N, K, a_mn, b_mn, str(e)[:200], flush=True)
                return 1
    for (M, N, K) in [(16384, 4096, 4096), (16384, 28672, 4096), (16384, 4096, 14336)]:
        for bn, name in ((256, "1cta"), (512, "2cta")):
            burst = perf(M, N, K, bn, 10)
            sustained = perf(M, N, K, bn, 300)
            print(f"PERF {name} M={M} N={N} K={K}: burst {burst:.0f} TF/s, sustained(300 it) {sustained:.0f} TF/s", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
