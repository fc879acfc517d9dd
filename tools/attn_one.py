#!/usr/bin/env python
"""Launch the production self-attention shape a few times (target for `ncu -k regex:attention`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ltx2_b200 import ops  # noqa: E402

H, Dh = 32, 128
Tq = int(sys.argv[1]) if len(sys.argv) > 1 else 3456
Tk = int(sys.argv[2]) if len(sys.argv) > 2 else Tq
dev = torch.device("cuda:0")
q = torch.randn(1, H, Tq, Dh, device=dev).to(torch.bfloat16)
k = torch.randn(1, H, Tk, Dh, device=dev).to(torch.bfloat16)
qkv = torch.randn(1, Tk, 3 * H * Dh, device=dev).to(torch.bfloat16)
for _ in range(3):
    out = ops.attention_vrows(q, k, qkv[:, :, 2 * H * Dh:], H, Dh)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
