"""GPU parity of each sm_100a kernel, called through the C ABI, against a torch fp32 reference of the
same op on the same (bf16-rounded) inputs and against the oracle's functions."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def dev():
    return torch.device("cuda:0")


def rnd(*shape, seed=0, std=1.0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * std).to(dtype).to(dev())


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


# bf16 output rounding is 2^-9 relative per element; fp32-accumulated GEMMs differ only by summation order.
TOL_BF16_OUT = 6e-3
TOL_F32_OUT = 2e-5


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 256), (200, 384, 320), (1000, 128, 128),
                                   (96, 32, 4096), (130, 64, 192), (3456, 4096, 4096), (864, 12288, 512)])
def test_gemm_bf16_store(M, N, K):
    from ltx2_b200 import ops
    a, w = rnd(M, K, seed=1, dtype=torch.bfloat16), rnd(N, K, seed=2, std=K ** -0.5, dtype=torch.bfloat16)
    bias = rnd(N, seed=3)
    ref = a.float() @ w.float().T + bias
    out = ops.gemm(a, w, bias, mode=ops.EPI_BF16)
    assert rel_err(out.float(), ref) < TOL_BF16_OUT
    out32 = ops.gemm(a, w, bias, mode=ops.EPI_F32)
    assert rel_err(out32, ref) < TOL_F32_OUT
    assert float((out32 - ref).abs().max()) < 2e-4 * max(1.0, float(ref.abs().max()))


def test_gemm_epilogues_gelu_and_gated_residual():
    from ltx2_b200 import ops
    M, N, K = 520, 512, 256
    a, w = rnd(M, K, seed=4, dtype=torch.bfloat16), rnd(N, K, seed=5, std=K ** -0.5, dtype=torch.bfloat16)
    bias = rnd(N, seed=6)
    lin = a.float() @ w.float().T + bias
    out = ops.gemm(a, w, bias, mode=ops.EPI_BF16_GELU)
    assert rel_err(out.float(), torch.nn.functional.gelu(lin, approximate="tanh")) < TOL_BF16_OUT
    x = rnd(M, N, seed=7)
    gate = rnd(3, 2 * N, seed=8)                       # 3 classes, row pitch 2N, gate lives in the 2nd half
    cls = (torch.arange(M, device=dev()) % 3).to(torch.int32)
    ref = x + 0.5 * gate[:, N:][cls.long()] * lin
    y = x.clone()
    ops.gemm(a, w, bias, mode=ops.EPI_F32_RESIDUAL, out=y, gate=gate[:, N:], row_cls=cls, alpha=0.5)
    assert rel_err(y, ref) < TOL_F32_OUT
    y2 = x.clone()
    ops.gemm(a, w, None, mode=ops.EPI_F32_RESIDUAL, out=y2)      # no bias, no gate
    assert rel_err(y2, x + a.float() @ w.float().T) < TOL_F32_OUT


def test_gemm_strided_operands_fused_qkv_layout():
    from ltx2_b200 import ops
    M, D = 300, 256
    a = rnd(M, D, seed=9, dtype=torch.bfloat16)
    wqkv = rnd(3 * D, D, seed=10, std=D ** -0.5, dtype=torch.bfloat16)
    out = torch.zeros(M, 3 * D, device=dev(), dtype=torch.bfloat16)
    ops.gemm(a, wqkv, None, mode=ops.EPI_BF16, out=out)
    assert rel_err(out.float(), a.float() @ wqkv.float().T) < TOL_BF16_OUT
    # second GEMM reads a column slice (row pitch 3D) as its A operand
    wo = rnd(D, D, seed=11, std=D ** -0.5, dtype=torch.bfloat16)
    k_slice = out[:, D:2 * D]
    assert not k_slice.is_contiguous()
    y = torch.empty(M, D, device=dev(), dtype=torch.bfloat16)
    from ltx2_b200._lib import lib, ptr, stream_ptr, check
    import ctypes as C
    check(lib().ltx2_gemm_bf16(ptr(k_slice), C.c_int64(3 * D), ptr(wo), C.c_int64(D), M, D, D, 0, None, ptr(y),
                               C.c_int64(D), None, C.c_int64(0), None, C.c_float(1.0), stream_ptr()))
    assert rel_err(y.float(), k_slice.float() @ wo.float().T) < TOL_BF16_OUT


def _attn_ref(q, k, v, gate=None):
    B, H, Tq, Dh = q.shape
    s = (q.float() @ k.float().transpose(-1, -2)) / math.sqrt(Dh)
    o = torch.softmax(s, -1) @ v.float()                      # (B,H,Tq,Dh)
    lse = torch.logsumexp(s, -1)
    if gate is not None:
        o = o * (2 * torch.sigmoid(gate)).reshape(B, Tq, H).permute(0, 2, 1)[..., None]
    return o.permute(0, 2, 1, 3).reshape(B, Tq, H * Dh), lse


@pytest.mark.parametrize("B,H,Tq,Tk,Dh", [(1, 2, 128, 128, 128), (1, 2, 256, 384, 128), (2, 3, 200, 333, 128),
                                          (1, 4, 130, 65, 64), (1, 2, 65, 520, 64), (1, 2, 864, 864, 128),
                                          (1, 32, 3456, 3456, 128), (1, 8, 384, 1024, 128)])
def test_attention(B, H, Tq, Tk, Dh):
    from ltx2_b200 import ops
    q = rnd(B, H, Tq, Dh, seed=20, dtype=torch.bfloat16)
    k = rnd(B, H, Tk, Dh, seed=21, dtype=torch.bfloat16)
    v = rnd(B, H, Tk, Dh, seed=22, dtype=torch.bfloat16)
    Tp = (Tk + 63) // 64 * 64
    vt = torch.zeros(B, H, Dh, Tp, device=dev(), dtype=torch.bfloat16)
    vt[..., :Tk] = v.transpose(-1, -2)
    out, lse = ops.attention(q, k, vt.contiguous(), Tk, want_lse=True)
    ref, ref_lse = _attn_ref(q, k, v)
    # P is rounded to bf16 before P*V and the output is bf16: ~2^-8 relative
    assert rel_err(out.float(), ref) < 1.2e-2
    assert float((lse - ref_lse).abs().max()) < 2e-3


@pytest.mark.parametrize("pairs", ["-1", "0", "1"])
@pytest.mark.parametrize("B,H,Tq,Tk", [(1, 2, 256, 384), (2, 3, 200, 333), (1, 2, 640, 128), (1, 3, 384, 1000),
                                       (1, 2, 130, 65)])
def test_attention_two_stream_modes(monkeypatch, pairs, B, H, Tq, Tk):
    """head_dim 128: force the pair (two query tiles share K/V) and split-KV (one tile, merged halves) work items on
    shapes small enough that the scheduler would not pick them: ragged tails, odd tile counts, a single key block."""
    from ltx2_b200 import ops
    monkeypatch.setenv("LTX2_ATTN_PAIRS", pairs)
    Dh = 128
    q = rnd(B, H, Tq, Dh, seed=27, std=2.0, dtype=torch.bfloat16)
    k = rnd(B, H, Tk, Dh, seed=28, std=2.0, dtype=torch.bfloat16)
    qkv = rnd(B, Tk, 3 * H * Dh, seed=29, dtype=torch.bfloat16)
    v_rows = qkv[:, :, 2 * H * Dh:]
    gate = rnd(B * Tq, H, seed=26)
    out = ops.attention_vrows(q, k, v_rows, H, Dh, gate_logits=gate)
    v = v_rows.reshape(B, Tk, H, Dh).permute(0, 2, 1, 3)
    ref, ref_lse = _attn_ref(q, k, v, gate)
    assert rel_err(out.float(), ref) < 1.5e-2
    Tp = (Tk + 63) // 64 * 64
    vt = torch.zeros(B, H, Dh, Tp, device=dev(), dtype=torch.bfloat16)
    vt[..., :Tk] = v.transpose(-1, -2)
    out2, lse = ops.attention(q, k, vt.contiguous(), Tk, want_lse=True)
    ref2, _ = _attn_ref(q, k, v)
    assert rel_err(out2.float(), ref2) < 1.5e-2
    assert float((lse - ref_lse).abs().max()) < 2e-3


@pytest.mark.parametrize("split", ["0", "1"])
@pytest.mark.parametrize("B,H,Tq,Tk,std", [(1, 2, 256, 128, 1.0), (1, 1, 256, 40, 1.0), (2, 3, 300, 333, 2.0),
                                           (1, 2, 384, 1000, 4.0), (1, 3, 640, 1024, 1.0), (1, 2, 130, 2000, 2.0),
                                           (2, 4, 1000, 700, 1.0)])
def test_attention_sm_pair_kernel(monkeypatch, split, B, H, Tq, Tk, std):
    """head_dim 128 SM-pair kernel (cta_group::2, attention_2cta_sm100.cu), forced on shapes the dispatcher would give to
    the one-SM kernel: odd query-tile counts (the second CTA of the last pair runs on an out-of-range tile), ragged key
    tails inside the first / second 64 keys of a block, a single key block, batch > 1, sharp logits (running maximum moves,
    lazy rescale), with whole items per cluster (split 0) and with (item, key block) ranges cut across clusters and
    merged by the last finisher (split 1)."""
    from ltx2_b200 import ops
    monkeypatch.setenv("LTX2_ATTN_2CTA", "1")
    monkeypatch.setenv("LTX2_ATTN_SPLIT", split)
    Dh = 128
    q = rnd(B, H, Tq, Dh, seed=31, std=std, dtype=torch.bfloat16)
    k = rnd(B, H, Tk, Dh, seed=32, std=2.0, dtype=torch.bfloat16)
    qkv = rnd(B, Tk, 3 * H * Dh, seed=33, dtype=torch.bfloat16)
    v_rows = qkv[:, :, 2 * H * Dh:]
    gate = rnd(B * Tq, H, seed=34)
    v = v_rows.reshape(B, Tk, H, Dh).permute(0, 2, 1, 3)
    out = ops.attention_vrows(q, k, v_rows, H, Dh, gate_logits=gate)
    ref, ref_lse = _attn_ref(q, k, v, gate)
    assert rel_err(out.float(), ref) < 1.5e-2
    out2, lse = ops.attention_vrows(q, k, v_rows, H, Dh, want_lse=True)
    ref2, _ = _attn_ref(q, k, v)
    assert rel_err(out2.float(), ref2) < 1.5e-2
    assert float((lse - ref_lse).abs().max()) < 2e-3
    # the same launch again: the merge counters of the split schedule were reset by the mergers
    out3 = ops.attention_vrows(q, k, v_rows, H, Dh)
    assert torch.equal(out2, out3)


@pytest.mark.parametrize("H,Tq,Tk", [(4, 3456, 3456), (32, 3456, 1024)])
def test_attention_sm_pair_kernel_production_shapes(monkeypatch, H, Tq, Tk):
    """The SM-pair kernel at the 19B shapes (4 heads = one rank of 8 under context parallelism: 56 items over 74 SM
    pairs, every item cut in two or three parts), against torch on two heads and bit-for-bit between its two schedules'
    own repeat runs (the merge order is fixed, so a schedule is deterministic)."""
    from ltx2_b200 import ops
    monkeypatch.setenv("LTX2_ATTN_2CTA", "1")
    Dh = 128
    q = rnd(1, H, Tq, Dh, seed=35, dtype=torch.bfloat16)
    k = rnd(1, H, Tk, Dh, seed=36, dtype=torch.bfloat16)
    qkv = rnd(1, Tk, 3 * H * Dh, seed=37, dtype=torch.bfloat16)
    v_rows = qkv[:, :, 2 * H * Dh:]
    v = v_rows.reshape(1, Tk, H, Dh).permute(0, 2, 1, 3)
    ref, _ = _attn_ref(q[:, :2], k[:, :2], v[:, :2])
    outs = {}
    for split in ("0", "1"):
        monkeypatch.setenv("LTX2_ATTN_SPLIT", split)
        out = ops.attention_vrows(q, k, v_rows, H, Dh)
        assert rel_err(out[:, :, :2 * Dh].float(), ref) < 1.2e-2
        assert torch.equal(out, ops.attention_vrows(q, k, v_rows, H, Dh))
        outs[split] = out
    assert rel_err(outs["0"].float(), outs["1"].float()) < 5e-3


def test_attention_long_sequences_dispatch_to_the_sm_pair_kernel(monkeypatch):
    """From 8192 queries and keys on, head_dim-128 attention takes the SM-pair kernel by itself (the N = 12288
    configuration of BASELINE.json); checked against torch on one head and against the one-SM kernel on all."""
    from ltx2_b200 import ops
    monkeypatch.delenv("LTX2_ATTN_2CTA", raising=False)
    monkeypatch.delenv("LTX2_ATTN_SPLIT", raising=False)
    B, H, T, Dh = 1, 3, 8192 + 128, 128
    q = rnd(B, H, T, Dh, seed=41, dtype=torch.bfloat16)
    k = rnd(B, H, T, Dh, seed=42, dtype=torch.bfloat16)
    qkv = rnd(B, T, 3 * H * Dh, seed=43, dtype=torch.bfloat16)
    v_rows = qkv[:, :, 2 * H * Dh:]
    v = v_rows.reshape(B, T, H, Dh).permute(0, 2, 1, 3)
    out = ops.attention_vrows(q, k, v_rows, H, Dh)
    ref, _ = _attn_ref(q[:, :1], k[:, :1], v[:, :1])
    assert rel_err(out[:, :, :Dh].float(), ref) < 1.2e-2
    monkeypatch.setenv("LTX2_ATTN_2CTA", "0")
    out1 = ops.attention_vrows(q, k, v_rows, H, Dh)
    assert not torch.equal(out, out1)                     # a different kernel did run
    assert rel_err(out.float(), out1.float()) < 5e-3


def test_attention_single_tile_kernel_still_matches(monkeypatch):
    """LTX2_ATTN_KERNEL=single keeps the one-tile kernel (the head_dim 64 path) reachable at head_dim 128 for A/B runs."""
    from ltx2_b200 import ops
    monkeypatch.setenv("LTX2_ATTN_KERNEL", "single")
    B, H, Tq, Tk, Dh = 1, 2, 300, 500, 128
    q, k = rnd(B, H, Tq, Dh, seed=20, dtype=torch.bfloat16), rnd(B, H, Tk, Dh, seed=21, dtype=torch.bfloat16)
    qkv = rnd(B, Tk, 3 * H * Dh, seed=22, dtype=torch.bfloat16)
    v_rows = qkv[:, :, 2 * H * Dh:]
    out = ops.attention_vrows(q, k, v_rows, H, Dh)
    ref, _ = _attn_ref(q, k, v_rows.reshape(B, Tk, H, Dh).permute(0, 2, 1, 3))
    assert rel_err(out.float(), ref) < 1.2e-2


def test_attention_sharp_softmax_and_gate():
    from ltx2_b200 import ops
    B, H, Tq, Tk, Dh = 1, 2, 256, 300, 128
    q = rnd(B, H, Tq, Dh, seed=23, std=4.0, dtype=torch.bfloat16)     # large logits -> running max moves a lot
    k = rnd(B, H, Tk, Dh, seed=24, std=2.0, dtype=torch.bfloat16)
    v = rnd(B, H, Tk, Dh, seed=25, dtype=torch.bfloat16)
    gate = rnd(B * Tq, H, seed=26)
    Tp = (Tk + 63) // 64 * 64
    vt = torch.zeros(B, H, Dh, Tp, device=dev(), dtype=torch.bfloat16)
    vt[..., :Tk] = v.transpose(-1, -2)
    out = ops.attention(q, k, vt, Tk, gate_logits=gate)
    ref, _ = _attn_ref(q, k, v, gate)
    assert rel_err(out.float(), ref) < 1.5e-2


@pytest.mark.parametrize("D", [128, 512, 4096])
def test_norm_modulate(D):
    from ltx2_b200 import ops
    from oracle import dit_oracle as O
    M = 77
    x = rnd(M, D, seed=30, std=3.0)
    mod = rnd(4, 9, D, seed=31, std=0.3)
    cls = (torch.arange(M, device=dev()) % 4).to(torch.int32)
    shift, scale = mod[cls.long(), 3], mod[cls.long(), 4]
    ref = O.rms_norm(x) * (1 + scale) + shift
    out = ops.norm_modulate(x, kind=ops.NORM_RMS, mod=mod, shift_row=3, scale_row=4, row_cls=cls)
    assert rel_err(out.float(), ref) < 4e-3
    assert rel_err(ops.norm_modulate(x, kind=ops.NORM_RMS).float(), O.rms_norm(x)) < 4e-3
    ln = torch.nn.functional.layer_norm(x, (D,), eps=1e-6) * (1 + mod[cls.long(), 1]) + mod[cls.long(), 0]
    out = ops.norm_modulate(x, kind=ops.NORM_LAYER, mod=mod, shift_row=0, scale_row=1, row_cls=cls)
    assert rel_err(out.float(), ln) < 4e-3
    xb = x.to(torch.bfloat16)
    out = ops.norm_modulate(xb, kind=ops.NORM_NONE, mod=mod, shift_row=0, scale_row=1, row_cls=cls)
    assert rel_err(out.float(), xb.float() * (1 + mod[cls.long(), 1]) + mod[cls.long(), 0]) < 4e-3


def test_rope_tables_match_reference_golden():
    from ltx2_b200 import ops
    g = np.load(os.path.join(GOLDEN, "rope.npz"))
    pos = torch.from_numpy(g["positions"]).to(dev())
    cos, sin = ops.rope_tables(pos, 4096, (20, 2048, 2048))
    ref_c = torch.from_numpy(g["cos"]).permute(0, 2, 1, 3).reshape(cos.shape).to(dev())   # (B,H,T,64)->(B,T,2048)
    ref_s = torch.from_numpy(g["sin"]).permute(0, 2, 1, 3).reshape(sin.shape).to(dev())
    # fp32 arguments up to ~1.5e4 rad: an ulp of the frequency grid moves cos/sin by ~1e-3
    assert float((cos - ref_c).abs().max()) < 1.5e-2 and float((sin - ref_s).abs().max()) < 1.5e-2
    assert float((cos - ref_c).abs().mean()) < 2e-4
    c1, s1 = ops.rope_tables(pos[:, 0:1].contiguous(), 2048, (20,))
    r1 = torch.from_numpy(g["cos_1d"]).permute(0, 2, 1, 3).reshape(c1.shape).to(dev())
    assert float((c1 - r1).abs().max()) < 1.5e-2


@pytest.mark.parametrize("H,Dh", [(4, 128), (32, 128), (4, 64)])
def test_headnorm_rope_and_v_transpose(H, Dh):
    from ltx2_b200 import ops
    from oracle import dit_oracle as O
    B, T = 2, 70
    inner = H * Dh
    qkv = rnd(B * T, 3 * inner, seed=40, dtype=torch.bfloat16)
    w = 1 + 0.1 * rnd(inner, seed=41)
    cos = torch.cos(rnd(B, T, inner // 2, seed=42, std=3.0))
    sin = torch.sin(rnd(B, T, inner // 2, seed=42, std=3.0))
    q = qkv[:, :inner]
    out = ops.headnorm_rope(q, w, B, T, H, Dh, cos, sin)
    cos_h = cos.reshape(B, T, H, Dh // 2).permute(0, 2, 1, 3).cpu()
    sin_h = sin.reshape(B, T, H, Dh // 2).permute(0, 2, 1, 3).cpu()
    qn = O.rms_norm(q.float().cpu().reshape(B, T, inner), w.cpu())
    ref = O.apply_split_rope(qn, cos_h, sin_h).reshape(B, T, H, Dh).permute(0, 2, 1, 3)
    assert rel_err(out.float().cpu(), ref) < 4e-3
    out = ops.headnorm_rope(q, w, B, T, H, Dh)                      # no RoPE (cross-attention)
    assert rel_err(out.float().cpu(), qn.reshape(B, T, H, Dh).permute(0, 2, 1, 3)) < 4e-3
    v = qkv[:, 2 * inner:]
    vt = ops.v_transpose(v, B, T, H, Dh)
    ref_v = v.reshape(B, T, H, Dh).permute(0, 2, 3, 1)
    assert torch.equal(vt[..., :T], ref_v)


def test_timestep_embedding_pieces():
    from ltx2_b200 import ops
    from oracle import dit_oracle as O
    t = torch.tensor([1.0, 0.725, 0.05], device=dev())
    s = ops.timestep_sinusoid(t, 1000.0)
    ref = O.sinusoid_dit(t.cpu() * 1000.0)
    assert float((s.cpu() - ref).abs().max()) < 2e-4      # fp32 sin/cos of arguments up to 1000 rad
    x = rnd(3, 512, seed=50)
    w = rnd(1536, 512, seed=51, std=512 ** -0.5, dtype=torch.bfloat16)
    b = rnd(1536, seed=52)
    y = ops.small_linear(x, w, b, act_in=1)
    ref = torch.nn.functional.silu(x) @ w.float().T + b
    assert rel_err(y, ref) < 1e-5


def test_x0_and_reference_metal_kernels():
    from ltx2_b200 import ops
    from oracle import dit_oracle as O
    lat, vel = rnd(50, 128, seed=60), rnd(50, 128, seed=61)
    t = torch.rand(50, device=dev())
    assert torch.allclose(ops.x0_from_velocity(lat, vel, t), lat - t[:, None] * vel, atol=1e-6)
    for dt, tol in ((torch.float32, 1e-6), (torch.bfloat16, 1e-2), (torch.float16, 2e-3)):
        a, b = rnd(3, 7, 64, seed=62, dtype=dt), rnd(3, 7, 64, seed=63, dtype=dt)
        assert rel_err(ops.silu_mul(a, b).float(), O.silu_mul(a.float(), b.float())) < tol
        assert rel_err(ops.gelu_mul(a, b).float(), O.gelu_mul(a.float(), b.float())) < tol
        th = torch.rand(3, 7, 32, device=dev()) * 6.28
        cos, sin = torch.cos(th).repeat_interleave(2, -1).to(dt), torch.sin(th).repeat_interleave(2, -1).to(dt)
        assert rel_err(ops.interleaved_rope(a, cos, sin).float(),
                       O.interleaved_rope(a.float(), cos.float(), sin.float())) < tol
    # empty input (the reference asserts equal shapes and launches a zero-size grid)
    e = torch.empty(0, 8, device=dev())
    assert ops.silu_mul(e, e).shape == (0, 8)
    with pytest.raises(AssertionError):
        ops.silu_mul(rnd(2, 4), rnd(2, 5))


@pytest.mark.parametrize("B,H,Tq,Tk,Dh", [(1, 2, 256, 384, 128), (2, 3, 200, 333, 128), (1, 4, 130, 65, 64),
                                          (1, 32, 3456, 3456, 128)])
def test_attention_v_in_row_form_from_fused_qkv(B, H, Tq, Tk, Dh):
    """V consumed directly from a token-major [B, T, 3*inner] buffer (MN-major tcgen05 operand), no transpose pass."""
    from ltx2_b200 import ops
    inner = H * Dh
    q = rnd(B, H, Tq, Dh, seed=70, dtype=torch.bfloat16)
    k = rnd(B, H, Tk, Dh, seed=71, dtype=torch.bfloat16)
    qkv = rnd(B, Tk, 3 * inner, seed=72, dtype=torch.bfloat16)
    v_rows = qkv[:, :, 2 * inner:]
    out = ops.attention_vrows(q, k, v_rows, H, Dh)
    v = v_rows.reshape(B, Tk, H, Dh).permute(0, 2, 1, 3)
    ref, _ = _attn_ref(q, k, v)
    assert rel_err(out.float(), ref) < 1.2e-2


@pytest.mark.parametrize("M", [3456, 3500, 3328, 2049, 432, 864, 65, 1000])
@pytest.mark.parametrize("mode", ["bf16", "gelu", "f32", "residual"])
def test_gemm_two_cta_transposed_tiles(monkeypatch, M, mode):
    """Pair kernel (tcgen05 cta_group::2): two SMs compute a C^T tile of 256 weight rows x tw tokens, tw a multiple of
    32 picked by the host (432 -> 3 x 160 ...), a narrow last tile scheduled as a half item (3456 = 13.5 x 256), rows
    past M zero-filled (3500); every epilogue mode.  Forced on here; the standard kernel must agree to rounding."""
    from ltx2_b200 import ops
    monkeypatch.setenv("LTX2_GEMM_2CTA", "2")          # take the pair kernel as soon as it fills the machine
    N, K = 2560, 320
    a, w = rnd(M, K, seed=85, dtype=torch.bfloat16), rnd(N, K, seed=86, std=K ** -0.5, dtype=torch.bfloat16)
    bias, x = rnd(N, seed=87), rnd(M, N, seed=88)
    gate = rnd(2, N, seed=89)
    cls = (torch.arange(M, device=dev()) % 2).to(torch.int32)
    acc = a.float() @ w.float().T + bias

    def run():
        if mode == "bf16":
            return ops.gemm(a, w, bias).float(), acc, TOL_BF16_OUT
        if mode == "gelu":
            return (ops.gemm(a, w, bias, mode=ops.EPI_BF16_GELU).float(),
                    torch.nn.functional.gelu(acc, approximate="tanh"), TOL_BF16_OUT)
        if mode == "f32":
            return ops.gemm(a, w, bias, mode=ops.EPI_F32), acc, TOL_F32_OUT
        y = x.clone()
        ops.gemm(a, w, bias, mode=ops.EPI_F32_RESIDUAL, out=y, gate=gate, row_cls=cls, alpha=0.5)
        return y, x + 0.5 * gate[cls.long()] * acc, TOL_F32_OUT

    out2, ref, tol = run()
    assert rel_err(out2, ref) < tol
    monkeypatch.setenv("LTX2_GEMM_2CTA", "0")
    monkeypatch.setenv("LTX2_GEMM_T", "0")
    out1, _, _ = run()
    assert rel_err(out2, out1) < 1e-5 if mode in ("f32", "residual") else rel_err(out2, out1) < 2e-3


@pytest.mark.parametrize("M,N,K", [(432, 1536, 512), (432, 4096, 2048), (65, 2048, 256), (864, 12288, 256), (200, 384, 640),
                                   (1000, 1024, 4096)])
@pytest.mark.parametrize("mode", ["bf16", "gelu", "f32", "residual", "residual_splitk"])
def test_gemm_transposed_tiles_for_ragged_token_counts(monkeypatch, M, N, K, mode):
    """Token counts that do not fill 128-row tiles (context-parallel shards, audio tokens) run as C^T tiles: 128 weight
    rows x nt tokens, nt a multiple of 16 chosen by the host (432 -> 3 x 144).  Forced on here for every shape and
    epilogue mode; the standard kernel on the same inputs must agree."""
    from ltx2_b200 import ops
    monkeypatch.setenv("LTX2_GEMM_T", "2")
    a, w = rnd(M, K, seed=90, dtype=torch.bfloat16), rnd(N, K, seed=91, std=K ** -0.5, dtype=torch.bfloat16)
    bias, x = rnd(N, seed=92), rnd(M, N, seed=93)
    gate = rnd(3, N, seed=94)
    cls = (torch.arange(M, device=dev()) % 3).to(torch.int32)
    acc = a.float() @ w.float().T + bias

    def run():
        if mode == "bf16":
            return ops.gemm(a, w, bias).float(), acc, TOL_BF16_OUT
        if mode == "gelu":
            return (ops.gemm(a, w, bias, mode=ops.EPI_BF16_GELU).float(),
                    torch.nn.functional.gelu(acc, approximate="tanh"), TOL_BF16_OUT)
        if mode == "f32":
            return ops.gemm(a, w, bias, mode=ops.EPI_F32), acc, TOL_F32_OUT
        y = x.clone()
        ops.gemm(a, w, bias, mode=ops.EPI_F32_RESIDUAL, out=y, gate=gate, row_cls=cls, alpha=0.5,
                 max_splits=8 if mode == "residual_splitk" else 1)
        return y, x + 0.5 * gate[cls.long()] * acc, TOL_F32_OUT

    out_t, ref, tol = run()
    assert rel_err(out_t, ref) < tol
    monkeypatch.setenv("LTX2_GEMM_T", "0")
    out_s, _, _ = run()
    assert rel_err(out_t, out_s) < (2e-3 if mode in ("bf16", "gelu") else 2e-5)


@pytest.mark.parametrize("M,N,K", [(432, 1536, 512), (432, 4096, 2048), (65, 2048, 256), (864, 2560, 320), (200, 512, 640),
                                   (1000, 1024, 4096), (512, 768, 1024), (513, 768, 1024), (1500, 512, 576)])
@pytest.mark.parametrize("mode", ["bf16", "gelu", "f32", "residual", "residual_splitk"])
def test_gemm_wide_pair_tiles_for_shards(monkeypatch, M, N, K, mode):
    """Wide pair kernel (cta_group::2, 256 weight rows x up to 512 tokens in two accumulators, K split for the residual
    epilogue): the context-parallel shard shapes, token counts just above / below one tile (512, 513), a ragged K tail
    (576 = 9 x 64, 320), every epilogue mode.  Forced on; the standard kernel on the same inputs must agree."""
    from ltx2_b200 import ops
    monkeypatch.setenv("LTX2_GEMM_WIDE", "2")
    a, w = rnd(M, K, seed=95, dtype=torch.bfloat16), rnd(N, K, seed=96, std=K ** -0.5, dtype=torch.bfloat16)
    bias, x = rnd(N, seed=97), rnd(M, N, seed=98)
    gate = rnd(3, N, seed=99)
    cls = (torch.arange(M, device=dev()) % 3).to(torch.int32)
    acc = a.float() @ w.float().T + bias

    def run():
        if mode == "bf16":
            return ops.gemm(a, w, bias).float(), acc, TOL_BF16_OUT
        if mode == "gelu":
            return (ops.gemm(a, w, bias, mode=ops.EPI_BF16_GELU).float(),
                    torch.nn.functional.gelu(acc, approximate="tanh"), TOL_BF16_OUT)
        if mode == "f32":
            return ops.gemm(a, w, bias, mode=ops.EPI_F32), acc, TOL_F32_OUT
        y = x.clone()
        ops.gemm(a, w, bias, mode=ops.EPI_F32_RESIDUAL, out=y, gate=gate, row_cls=cls, alpha=0.5,
                 max_splits=8 if mode == "residual_splitk" else 1)
        return y, x + 0.5 * gate[cls.long()] * acc, TOL_F32_OUT

    out_w, ref, tol = run()
    assert rel_err(out_w, ref) < tol
    monkeypatch.setenv("LTX2_GEMM_WIDE", "0")
    monkeypatch.setenv("LTX2_GEMM_2CTA", "0")
    monkeypatch.setenv("LTX2_GEMM_T", "0")
    out_s, _, _ = run()
    assert rel_err(out_w, out_s) < (2e-3 if mode in ("bf16", "gelu") else 2e-5)


@pytest.mark.parametrize("M,N,K", [(256, 1024, 4096), (432, 4096, 4096), (432, 4096, 16384), (100, 512, 640)])
def test_gemm_residual_split_k_small_m(M, N, K):
    """Small-M residual GEMMs (context-parallel ranks) split K across CTAs and accumulate with vector reductions."""
    from ltx2_b200 import ops
    a, w = rnd(M, K, seed=80, dtype=torch.bfloat16), rnd(N, K, seed=81, std=K ** -0.5, dtype=torch.bfloat16)
    bias, x = rnd(N, seed=82), rnd(M, N, seed=83)
    gate = rnd(2, N, seed=84)
    cls = (torch.arange(M, device=dev()) % 2).to(torch.int32)
    ref = x + gate[cls.long()] * (a.float() @ w.float().T + bias)
    y = x.clone()
    ops.gemm(a, w, bias, mode=ops.EPI_F32_RESIDUAL, out=y, gate=gate, row_cls=cls, max_splits=8)
    assert rel_err(y, ref) < TOL_F32_OUT
    assert float((y - ref).abs().max()) < 3e-4 * max(1.0, float(ref.abs().max()))
