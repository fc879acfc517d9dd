#!/usr/bin/env python
"""GEMM A/B: the library's tcgen05 GEMM against torch.matmul (cuBLAS) on the DiT shapes, each run back to back for
about a second so both sit in the same power-capped regime (diagnostics, not a bench value)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ltx2_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, seconds=1.0):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    while True:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
        if time.time() - t0 > seconds:
            break
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for M, N, K in [(3456, 12288, 4096), (3456, 4096, 4096), (3456, 16384, 4096), (3456, 4096, 16384), (8192, 8192, 8192),
                (1728, 16384, 4096), (432, 16384, 4096)]:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = torch.randn(N, K, device=dev).to(torch.bfloat16) * K ** -0.5
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    os.environ.pop("LTX2_GEMM_2CTA", None)
    ms_ours = timed(lambda: ops.gemm(a, w, None, out=out))
    os.environ["LTX2_GEMM_2CTA"] = "2"
    ms_pair = timed(lambda: ops.gemm(a, w, None, out=out))
    os.environ.pop("LTX2_GEMM_2CTA", None)
    ms_cublas = timed(lambda: torch.matmul(a, w.t(), out=out))
    print(f"M={M:5d} N={N:5d} K={K:5d}  1-CTA {fl / ms_ours / 1e9:7.1f} TF/s ({ms_ours * 1e3:7.1f} us)   "
          f"2-CTA {fl / ms_pair / 1e9:7.1f} TF/s ({ms_pair * 1e3:7.1f} us)   "
          f"cuBLAS {fl / ms_cublas / 1e9:7.1f} TF/s ({ms_cublas * 1e3:7.1f} us)", flush=True)
