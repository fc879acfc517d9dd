#!/usr/bin/env python
"""GEMM timing on the context-parallel shard shapes (432 / 864 / 1728 token rows): standard 128-row tiles vs the
transposed tiles the host picks (LTX2_GEMM_T=0 disables them) vs cuBLAS.  Diagnostics, not a bench value."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ltx2_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, iters=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for M in (432, 864, 1728, 3456):
    for N, K, resid in [(12288, 4096, False), (4096, 4096, True), (4096, 4096, False), (16384, 4096, False), (4096, 16384, True)]:
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = torch.randn(N, K, device=dev).to(torch.bfloat16) * K ** -0.5
        fl = 2.0 * M * N * K
        if resid:
            y = torch.zeros(M, N, device=dev)
            fn = lambda: ops.gemm(a, w, None, mode=ops.EPI_F32_RESIDUAL, out=y, max_splits=8)  # noqa: E731
        else:
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            fn = lambda: ops.gemm(a, w, None, out=out)  # noqa: E731
        os.environ["LTX2_GEMM_T"] = "0"
        os.environ["LTX2_GEMM_2CTA"] = "0"
        ms_std = timed(fn)
        os.environ.pop("LTX2_GEMM_T")
        os.environ.pop("LTX2_GEMM_2CTA")
        os.environ["LTX2_GEMM_WIDE"] = "0"
        ms_nowide = timed(fn)
        os.environ["LTX2_GEMM_WIDE"] = "2"
        ms_wide = timed(fn)
        os.environ.pop("LTX2_GEMM_WIDE")
        ms_auto = timed(fn)
        o2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        ms_cublas = timed(lambda: torch.matmul(a, w.t(), out=o2))
        print(f"M={M:5d} N={N:5d} K={K:5d} {'residual' if resid else 'bf16    '}  standard {ms_std * 1e3:7.1f} us   T/pair {ms_nowide * 1e3:7.1f} us   wide {ms_wide * 1e3:7.1f} us   "
              f"auto {ms_auto * 1e3:7.1f} us ({fl / ms_auto / 1e9:6.0f} TF/s)   cuBLAS {ms_cublas * 1e3:7.1f} us", flush=True)
