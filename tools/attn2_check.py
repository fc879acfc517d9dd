#!/usr/bin/env python
"""SM-pair attention kernel (attention_2cta_sm100.cu) against the one-SM two-stream kernel: timing with CUDA events and a
check against a torch fp32 reference on two heads, for the 19B shapes, the context-parallel head shards and the
N = 12288 configuration; then the pipeline timeline of cluster 0 (ltx2_attention_vrows_trace).  Diagnostics for kernel
tuning, not a bench value.  Rows: 2cta = the dispatcher's schedule, -ns = whole items per SM pair, -fs = stream-K forced,
-p2 / -p4 = polynomial share of the exponentials, pair = the one-SM kernel."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ltx2_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
Dh = 128


def case(Tq, Tk, heads, B=1, iters=20, check_heads=2, gate=False):
    g = torch.Generator(device=dev).manual_seed(1)
    q = torch.randn(B, heads, Tq, Dh, device=dev, generator=g).to(torch.bfloat16)
    k = torch.randn(B, heads, Tk, Dh, device=dev, generator=g).to(torch.bfloat16)
    qkv = torch.randn(B, Tk, 3 * heads * Dh, device=dev, generator=g).to(torch.bfloat16)
    v_rows = qkv[:, :, 2 * heads * Dh:]
    v = v_rows.reshape(B, Tk, heads, Dh).permute(0, 2, 1, 3)
    check_heads = min(check_heads, heads)
    hs = slice(0, check_heads)
    s = (q[:, hs].float() @ k[:, hs].float().transpose(-1, -2)) / math.sqrt(Dh)
    ref = (torch.softmax(s, -1) @ v[:, hs].float()).permute(0, 2, 1, 3).reshape(B, Tq, check_heads * Dh)
    flops = 4.0 * B * heads * Tq * Tk * Dh
    res = {}
    for name, env in (("2cta", {"LTX2_ATTN_2CTA": "1"}), ("2cta-ns", {"LTX2_ATTN_2CTA": "1", "LTX2_ATTN_SPLIT": "0"}), ("2cta-fs", {"LTX2_ATTN_2CTA": "1", "LTX2_ATTN_SPLIT": "1"}), ("2cta-p2", {"LTX2_ATTN_2CTA": "1", "LTX2_ATTN_POLY": "2"}), ("2cta-p4", {"LTX2_ATTN_2CTA": "1", "LTX2_ATTN_POLY": "4"}), ("pair", {"LTX2_ATTN_2CTA": "0"})):
        os.environ.pop("LTX2_ATTN_2CTA", None)
        os.environ.pop("LTX2_ATTN_POLY", None)
        os.environ.pop("LTX2_ATTN_SPLIT", None)
        os.environ.pop("LTX2_ATTN_DBG", None)
        os.environ.update(env)
        out = ops.attention_vrows(q, k, v_rows, heads, Dh)
        torch.cuda.synchronize()
        err = float((out[:, :, :check_heads * Dh].float() - ref).norm() / ref.norm())
        res[name] = out
        for _ in range(3):
            ops.attention_vrows(q, k, v_rows, heads, Dh)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ops.attention_vrows(q, k, v_rows, heads, Dh)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"Tq={Tq} Tk={Tk} H={heads} B={B} {name:8s} {ms*1e3:8.1f} us {flops/ms/1e9:7.1f} TF/s rel.err {err:.2e} finite {bool(torch.isfinite(out).all())}", flush=True)
    d = float((res["2cta"].float() - res["pair"].float()).abs().max())
    print(f"   max |2cta - pair| = {d:.3e}", flush=True)


for shape in [(384, 1000, 2), (3456, 3456, 4), (3456, 3456, 16), (3456, 3456, 32), (3456, 1024, 32), (12288, 12288, 8), (6144, 6144, 8)]:
    case(*shape)
case(640, 333, 4, B=2)


def timeline(Tq, Tk, heads):
    import ctypes as C
    from ltx2_b200._lib import check, lib, ptr, stream_ptr
    os.environ["LTX2_ATTN_2CTA"] = "1"
    g = torch.Generator(device=dev).manual_seed(1)
    q = torch.randn(1, heads, Tq, Dh, device=dev, generator=g).to(torch.bfloat16)
    k = torch.randn(1, heads, Tk, Dh, device=dev, generator=g).to(torch.bfloat16)
    qkv = torch.randn(1, Tk, 3 * heads * Dh, device=dev, generator=g).to(torch.bfloat16)
    v_rows = qkv[:, :, 2 * heads * Dh:]
    nblk = (Tk + 127) // 128
    tr = torch.zeros(16 * nblk, device=dev, dtype=torch.int64)
    out = torch.empty(1, Tq, heads * Dh, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        check(lib().ltx2_attention_vrows_trace(ptr(q), ptr(k), ptr(v_rows), C.c_int64(v_rows.stride(1)), C.c_int64(Dh),
                                               C.c_int64(v_rows.stride(0)), ptr(out), 1, heads, Tq, Tk, Dh,
                                               C.c_float(1.0 / math.sqrt(Dh)), ptr(tr), stream_ptr()), "trace")
    torch.cuda.synchronize()
    t = tr.cpu().view(nblk, 16)
    t0 = int(t[0, 2])
    names = ["P seen(MMA)", "PV+S issued", "A: S seen", "A: exp done", "A: published", "B: S seen", "B: exp done", "B: published"]
    print("block " + " ".join(f"{n:>13s}" for n in names))
    for kb in range(min(nblk, 6)):
        print(f"{kb:5d} " + " ".join(f"{int(t[kb, e]) - t0:13d}" for e in range(8)))
    print("MMA warp, relative to 'P seen': entry acquired, PV issued, S issued, empty committed, s_full committed")
    for kb in range(2, min(nblk, 6)):
        print(f"{kb:5d} " + " ".join(f"{int(t[kb, e]) - int(t[kb, 0]):8d}" for e in (8, 9, 10, 11, 1)))
    if nblk > 6:
        per = (int(t[nblk - 2, 4]) - int(t[2, 4])) / (nblk - 4)
        print(f"period per 128-key block: {per:.0f} clk (tensor floor 1024)")


if os.environ.get("ATTN2_TIMELINE", "1") != "0":
    for dbg in ("0",):
        os.environ["LTX2_ATTN_DBG"] = dbg
        print("== LTX2_ATTN_DBG =", dbg, "(1: no S MMAs in the loop, 2: no P*V MMAs)")
        timeline(3456, 3456, 32)
