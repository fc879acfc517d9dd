#!/usr/bin/env python
"""GEMM timing with COLD weights: every launch uses a different weight matrix out of a pool larger than the L2, the
way a denoising step streams 25.8 GB of weights through the 126 MB L2.  Context-parallel shard shapes.  Diagnostics."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ltx2_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timed(fns, rounds=3):
    for f in fns:
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rounds):
        for f in fns:
            f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (rounds * len(fns)) * 1e3


for M in (432, 864, 3456):
    for N, K, resid in [(12288, 4096, False), (4096, 4096, True), (4096, 4096, False), (16384, 4096, False), (4096, 16384, True)]:
        pool = max(2, int(1.5e9 // (N * K * 2)))          # >= 1.5 GB of distinct weights
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        ws = [torch.randn(N, K, device=dev).to(torch.bfloat16) * K ** -0.5 for _ in range(pool)]
        fl = 2.0 * M * N * K
        if resid:
            y = torch.zeros(M, N, device=dev)
            mk = lambda w: (lambda: ops.gemm(a, w, None, mode=ops.EPI_F32_RESIDUAL, out=y, max_splits=8))  # noqa: E731
        else:
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            mk = lambda w: (lambda: ops.gemm(a, w, None, out=out))  # noqa: E731
        cold = timed([mk(w) for w in ws])
        warm = timed([mk(ws[0])] * pool)
        o2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        cub_cold = timed([(lambda w=w: torch.matmul(a, w.t(), out=o2)) for w in ws])
        cub_warm = timed([(lambda: torch.matmul(a, ws[0].t(), out=o2))] * pool)
        print(f"M={M:5d} N={N:5d} K={K:5d} {'residual' if resid else 'bf16    '}  ours cold {cold:7.1f} us warm {warm:7.1f} us "
              f"({fl / cold / 1e6:6.0f} / {fl / warm / 1e6:6.0f} TF/s)   cuBLAS cold {cub_cold:7.1f} warm {cub_warm:7.1f} us   "
              f"weights {N * K * 2 / 1e6:5.0f} MB = {N * K * 2 / cold / 1e6:5.2f} TB/s cold", flush=True)
        del ws
        torch.cuda.empty_cache()
