#!/usr/bin/env python
"""Print the pipeline timeline of one attention CTA (diagnostics for kernel tuning)."""
import ctypes as C
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

os.environ["LTX2_ATTN_KERNEL"] = "single"   # this tool reads the single-tile kernel's trace layout; see attn_bench.py

from ltx2_b200._lib import check, lib, ptr, stream_ptr  # noqa: E402

B, H, T, Dh = 1, 32, 3456, 128
dev = torch.device("cuda:0")
q = torch.randn(B, H, T, Dh, device=dev).to(torch.bfloat16)
k = torch.randn(B, H, T, Dh, device=dev).to(torch.bfloat16)
vt = torch.randn(B, H, Dh, T, device=dev).to(torch.bfloat16)
out = torch.empty(B, T, H * Dh, device=dev, dtype=torch.bfloat16)
nkv = (T + 127) // 128
trace = torch.zeros(nkv * 8, device=dev, dtype=torch.int64)
for _ in range(2):
    check(lib().ltx2_attention_trace(ptr(q), ptr(k), ptr(vt), ptr(out), B, H, T, T, T, Dh, C.c_float(1 / math.sqrt(Dh)),
                                     ptr(trace), stream_ptr()))
torch.cuda.synchronize()
t = trace.cpu().reshape(nkv, 8)
t0 = int(t[0][t[0] > 0].min())
names = ["QK_j issued", "P_j seen(MMA)", "PV_j issued", "S_j seen", "S_j in regs", "exps done", "PV_j-1 retired", "P_j published"]
print("block " + " ".join(f"{n:>15s}" for n in names))
for j in range(min(nkv, 12)):
    print(f"{j:5d} " + " ".join(f"{int(x) - t0:15d}" for x in t[j]))
d = (t[1:, 7] - t[:-1, 7]).float()
print("period (P published -> next):", d[2:].mean().item(), "cycles")
print("softmax: wait S", (t[3:, 3] - t[2:-1, 7]).float().mean().item(), " ld", (t[3:, 4] - t[3:, 3]).float().mean().item(),
      " max+exp", (t[3:, 5] - t[3:, 4]).float().mean().item(), " wait PV", (t[3:, 6] - t[3:, 5]).float().mean().item(),
      " st P", (t[3:, 7] - t[3:, 6]).float().mean().item())
print("MMA: P seen -> PV issued", (t[3:, 2] - t[3:, 1]).float().mean().item(), " PV_j issued -> QK_{j+2} issued",
      (t[5:, 0] - t[3:-2, 2]).float().mean().item())
