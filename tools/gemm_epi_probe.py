#!/usr/bin/env python
"""Fixed per-item cost of the C^T (transposed) epilogue: the wide / pair / transposed GEMM kernels at K = 64 (one K block)
against K = 4096, every epilogue mode.  Diagnostics."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ltx2_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, iters=100):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for M, N in ((432, 4096), (432, 16384)):
    for K in (64, 1024, 4096):
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = torch.randn(N, K, device=dev).to(torch.bfloat16) * K ** -0.5
        row = []
        for mode, name in ((ops.EPI_BF16, "bf16"), (ops.EPI_F32, "f32"), (ops.EPI_F32_RESIDUAL, "resid")):
            if mode == ops.EPI_BF16:
                out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            else:
                out = torch.zeros(M, N, device=dev)
            fn = lambda: ops.gemm(a, w, None, mode=mode, out=out)  # noqa: E731
            for env, tag in (({"LTX2_GEMM_WIDE": "2"}, "wide"), ({"LTX2_GEMM_2CTA": "2"}, "pair"), ({"LTX2_GEMM_T": "2"}, "T"),
                             ({"LTX2_GEMM_T": "0", "LTX2_GEMM_2CTA": "0"}, "std")):
                for k in ("LTX2_GEMM_WIDE", "LTX2_GEMM_2CTA", "LTX2_GEMM_T"):
                    os.environ.pop(k, None)
                os.environ.update(env)
                row.append(f"{name}/{tag} {timed(fn):6.1f}")
        print(f"M={M} N={N} K={K}: " + "  ".join(row), flush=True)
