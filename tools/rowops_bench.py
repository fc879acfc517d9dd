#!/usr/bin/env python
"""Timing of the HBM-bound row kernels of the DiT at the bench shape (3456 x 4096): achieved GB/s against the
algorithmic bytes of DESIGN.md section 4 (diagnostics)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ltx2_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
M, D, H, Dh = 3456, 4096, 32, 128


def timed(fn, iters=50):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()                                       # evict the 126 MB L2 between iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


x = torch.randn(M, D, device=dev)
mod = torch.randn(1, 9, D, device=dev) * 0.1
cls = torch.zeros(M, dtype=torch.int32, device=dev)
ms = timed(lambda: ops.norm_modulate(x, kind=ops.NORM_RMS, mod=mod, shift_row=0, scale_row=1, row_cls=cls))
print(f"norm_modulate      {ms * 1e3:6.1f} us  {M * D * 6 / ms / 1e6:7.0f} GB/s (fp32 in + bf16 out = {M * D * 6 / 1e6:.0f} MB)")
qkv = torch.randn(1, M, 3 * D, device=dev).to(torch.bfloat16)
w = torch.ones(D, device=dev)
pos = torch.rand(1, 3, M, 2, device=dev)
cos = torch.randn(1, M, D // 2, device=dev)
sin = torch.randn(1, M, D // 2, device=dev)
if hasattr(ops, "qkv_head_scatter"):
    ms = timed(lambda: ops.qkv_head_scatter(qkv, w, w, cos, sin, 1, M, H, Dh))
    by = M * D * 2 * 2 + M * D * 4 + M * D * 2 * 2
    print(f"qkv_head_scatter   {ms * 1e3:6.1f} us  {by / ms / 1e6:7.0f} GB/s (q,k in + cos,sin + q,k out = {by / 1e6:.0f} MB)")
q = torch.randn(M, D, device=dev).to(torch.bfloat16)
ms = timed(lambda: ops.headnorm_rope(q, w, 1, M, H, Dh, cos=cos, sin=sin))
by = M * D * 4 + M * D * 4
print(f"headnorm + rope    {ms * 1e3:6.1f} us  {by / ms / 1e6:7.0f} GB/s (bf16 in + out + cos,sin = {by / 1e6:.0f} MB)")
ms = timed(lambda: ops.headnorm_rope(q, w, 1, M, H, Dh))
print(f"headnorm (no rope) {ms * 1e3:6.1f} us  {M * D * 4 / ms / 1e6:7.0f} GB/s (bf16 in + out = {M * D * 4 / 1e6:.0f} MB)")
