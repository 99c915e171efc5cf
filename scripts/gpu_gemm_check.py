"""Bring-up check for the tcgen05 GEMM on a real B200 (run under gpurun). Not a pytest: prints a
table so a failing operand-layout variant can be told apart from a broken pipeline."""
import sys
import time

import torch

sys.path.insert(0, ".")
from metamorph_b200 import ops  # noqa: E402


def ref(a, b, a_mn, b_mn):
    A = a.float().t() if a_mn else a.float()
    B = b.float() if b_mn else b.float().t()
    return A @ B


def run_case(M, N, K, a_mn, b_mn, bn, epi=0):
    torch.manual_seed(M + N + K)
    dev = "cuda"
    a = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16()
    b = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
    out = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, force_bn=bn)
    torch.cuda.synchronize()
    r = ref(a, b, a_mn, b_mn)
    err = (out.float() - r).abs().max().item()
    scale = r.abs().max().item()
    ok = err <= 2e-2 * scale
    print(f"M={M:6d} N={N:6d} K={K:6d} a_mn={int(a_mn)} b_mn={int(b_mn)} bn={bn:3d} "
          f"max_err={err:.4f} ref_max={scale:.2f} {'OK' if ok else 'FAIL'}", flush=True)
    return ok


def main():
    print(torch.cuda.get_device_name(0), flush=True)
    allok = True
    for (a_mn, b_mn) in [(False, False), (False, True), (True, True)]:
        for bn in (128, 256):
            for (M, N, K) in [(128, 256, 64), (128, 256, 256), (256, 512, 1024),
                              (300, 520, 200) if not (a_mn or b_mn) else (304, 520, 200),
                              (1024, 1152, 4304 if not (a_mn or b_mn) else 4096)]:
                try:
                    allok &= run_case(M, N, K, a_mn, b_mn, bn)
                except Exception as e:  # noqa: BLE001
                    allok = False
                    print(f"EXC M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} bn={bn}: {e}", flush=True)
                    return 1
    # throughput
    for (M, N, K) in [(8192, 8192, 8192), (16384, 4096, 4096), (16384, 28672, 4096), (16384, 4096, 14336)]:
        a = torch.randn(M, K, device="cuda").bfloat16()
        b = torch.randn(N, K, device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        for _ in range(3):
            ops.gemm(a, b, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        iters = 10
        for _ in range(iters):
            ops.gemm(a, b, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"PERF M={M} N={N} K={K}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
        t0 = time.time()
        ref_out = a @ b.t()
        torch.cuda.synchronize()
        err = (out.float() - ref_out.float()).abs().max().item()
        print(f"     vs cuBLAS max_err={err:.4f}", flush=True)
    print("ALL OK" if allok else "SOME FAILED", flush=True)
    return 0 if allok else 1


if __name__ == "__main__":
    sys.exit(main())
