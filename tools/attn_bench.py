#!/usr/bin/env python
"""Attention kernel A/B timing + pipeline timeline (diagnostics for kernel tuning, not a bench value).

Times the production shapes (self-attention 3456x3456 and text cross-attention 3456x1024, 32 heads x 128) with CUDA
events for the single-tile kernel and the two-stream kernel under its tuning knobs, checks each against a torch fp32
reference on a slice, and prints the per-block timeline of CTA 0 of the two-stream kernel.
"""
import ctypes as C
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ltx2_b200 import ops  # noqa: E402
from ltx2_b200._lib import check, lib, ptr, stream_ptr  # noqa: E402

dev = torch.device("cuda:0")
H, Dh = 32, 128


def setenv(**kw):
    for k in ("LTX2_ATTN_KERNEL", "LTX2_ATTN_POLY", "LTX2_ATTN_PAIRS"):
        os.environ.pop(k, None)
    for k, v in kw.items():
        os.environ[k] = str(v)


def run(Tq, Tk, heads=H, iters=20):
    g = torch.Generator(device=dev).manual_seed(1)
    q = torch.randn(1, heads, Tq, Dh, device=dev, generator=g).to(torch.bfloat16)
    k = torch.randn(1, heads, Tk, Dh, device=dev, generator=g).to(torch.bfloat16)
    qkv = torch.randn(1, Tk, 3 * heads * Dh, device=dev, generator=g).to(torch.bfloat16)
    v_rows = qkv[:, :, 2 * heads * Dh:]
    v = v_rows.reshape(1, Tk, heads, Dh).permute(0, 2, 1, 3)
    hs = slice(0, 2)
    s = (q[:, hs].float() @ k[:, hs].float().transpose(-1, -2)) / math.sqrt(Dh)
    ref = (torch.softmax(s, -1) @ v[:, hs].float()).permute(0, 2, 1, 3).reshape(1, Tq, 2 * Dh)
    flops = 4.0 * heads * Tq * Tk * Dh
    cfgs = [("single-tile", dict(LTX2_ATTN_KERNEL="single")), ("pair poly3 (default)", {}),
            ("pair poly0", dict(LTX2_ATTN_POLY=0)), ("pair poly2", dict(LTX2_ATTN_POLY=2)),
            ("pair poly4", dict(LTX2_ATTN_POLY=4)), ("pair NO softmax math", dict(LTX2_ATTN_POLY=8)), ("all split-KV", dict(LTX2_ATTN_PAIRS=0)),
            ("max pairs", dict(LTX2_ATTN_PAIRS=-1))]
    for name, env in cfgs:
        setenv(**env)
        out = ops.attention_vrows(q, k, v_rows, heads, Dh)
        torch.cuda.synchronize()
        err = float((out[:, :, :2 * Dh].float() - ref).norm() / ref.norm())
        for _ in range(3):
            ops.attention_vrows(q, k, v_rows, heads, Dh)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ops.attention_vrows(q, k, v_rows, heads, Dh)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"  Tq={Tq} Tk={Tk} H={heads} {name:22s} {ms * 1e3:8.1f} us  {flops / ms / 1e9:7.1f} TF/s  rel.err {err:.2e}",
              flush=True)
    setenv()


def timeline(T=3456):
    setenv(LTX2_ATTN_PAIRS=-1)
    q = torch.randn(1, H, T, Dh, device=dev).to(torch.bfloat16)
    k = torch.randn(1, H, T, Dh, device=dev).to(torch.bfloat16)
    vt = torch.randn(1, H, Dh, T, device=dev).to(torch.bfloat16)
    out = torch.empty(1, T, H * Dh, device=dev, dtype=torch.bfloat16)
    nkv = (T + 127) // 128
    trace = torch.zeros(nkv * 16 + 16, device=dev, dtype=torch.int64)
    for _ in range(2):
        check(lib().ltx2_attention_trace(ptr(q), ptr(k), ptr(vt), ptr(out), 1, H, T, T, T, Dh,
                                         C.c_float(1 / math.sqrt(Dh)), ptr(trace), stream_ptr()))
    torch.cuda.synchronize()
    ph = trace.cpu()[nkv * 16:nkv * 16 + 4]
    t = trace.cpu()[:nkv * 16].reshape(nkv, 16)
    print('CTA 0 phases (cycles): set-up', int(ph[1] - ph[0]), ' key loop', int(ph[2] - ph[1]), ' epilogue', int(ph[3] - ph[2]))
    t0 = int(t[0][t[0] > 0].min())
    names = ["P0a seen", "PV0a+S iss", "P1a seen", "PV1a+S iss", "S0a seen", "S0a regs", "exp0a done", "P0a pub",
             "S0b seen", "S0b regs", "exp0b done", "P0b pub", "S1a seen", "S1a regs", "exp1a done", "P1a pub"]
    print("tile " + " ".join(f"{n:>10s}" for n in names))
    for j in range(min(nkv, 10)):
        print(f"{j:4d} " + " ".join(f"{int(x) - t0:10d}" for x in t[j][:16]))
    d = (t[3:-1, 7] - t[2:-2, 7]).float()
    print("period of stream 0 per 128-key tile (P0a published -> next):", d.mean().item(), "cycles  (tensor floor 2048)")
    print("softmax 0, half a: S seen->regs", (t[2:-1, 5] - t[2:-1, 4]).float().mean().item(), " regs->exp done",
          (t[2:-1, 6] - t[2:-1, 5]).float().mean().item(), " exp done->published",
          (t[2:-1, 7] - t[2:-1, 6]).float().mean().item(), " published->S0b seen",
          (t[2:-1, 8] - t[2:-1, 7]).float().mean().item(), " P0b published->next S0a seen",
          (t[3:-1, 4] - t[2:-2, 11]).float().mean().item())
    print("MMA: P0a seen->issued", (t[2:-1, 1] - t[2:-1, 0]).float().mean().item(), " issued->P1a seen",
          (t[2:-1, 2] - t[2:-1, 1]).float().mean().item(), " P1a seen->issued", (t[2:-1, 3] - t[2:-1, 2]).float().mean().item())
    setenv()


if __name__ == "__main__":
    print("== self-attention", flush=True)
    run(3456, 3456)
    print("== text cross-attention", flush=True)
    run(3456, 1024)
    print("== context-parallel head shards (8 ranks: 4 heads, 2 ranks: 16 heads)", flush=True)
    run(3456, 3456, heads=4)
    run(3456, 3456, heads=16)
    print("== 1024x768x121 (N = 12288), 8 heads", flush=True)
    run(12288, 12288, heads=8, iters=5)
    timeline()
