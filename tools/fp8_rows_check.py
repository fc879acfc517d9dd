#!/usr/bin/env python
"""FP8 GEMM (ltx2_gemm_e4m3): the first M rows of a 3456-row launch against a launch on those M rows alone, for the
context-parallel shard sizes (different tile widths are chosen) -- must be bit-identical.  Diagnostics."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ltx2_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
for K in (4096, 16384):
    for N in (4096, 12288, 16384):
        a = torch.randn(3456, K, device=dev)
        w = torch.randn(N, K, device=dev) * K ** -0.5
        a8, as_ = ops.quantize_rows_e4m3(a)
        w8, ws = ops.quantize_rows_e4m3(w)
        for mode in (ops.EPI_BF16, ops.EPI_F32):
            full = ops.gemm_e4m3(a8, as_, w8, ws, None, mode=mode)
            for M in (432, 864, 1728):
                part = ops.gemm_e4m3(a8[:M].contiguous(), as_[:M].contiguous(), w8, ws, None, mode=mode)
                same = torch.equal(full[:M], part)
                d = float((full[:M].float() - part.float()).abs().max())
                print(f"K={K} N={N} mode={mode} M={M}: bit-equal to rows of the M=3456 launch: {same}  max diff {d:.3e}", flush=True)
