"""Thin torch-tensor wrappers over the per-op C-ABI entry points (unit parity, glue).

Every function launches on torch's current CUDA stream and raises Ltx2Error on failure.
Tensors must be CUDA and contiguous in the layout documented in include/ltx2_b200.h.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr, dtype_code

EPI_BF16, EPI_BF16_GELU, EPI_F32, EPI_F32_RESIDUAL = 0, 1, 2, 3
NORM_NONE, NORM_RMS, NORM_LAYER = 0, 1, 2


def _cuda(*ts):
    for t in ts:
        if t is not None:
            ok = t.is_contiguous() or (t.ndim == 2 and t.stride(1) == 1)     # row-pitched 2-D views are fine
            assert t.is_cuda and ok, "ltx2_b200 ops need CUDA tensors with a contiguous innermost dimension"


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, mode: int = EPI_BF16,
         out: Optional[torch.Tensor] = None, gate: Optional[torch.Tensor] = None,
         row_cls: Optional[torch.Tensor] = None, alpha: float = 1.0, max_splits: int = 1) -> torch.Tensor:
    """C = A W^T (+ epilogue).  a [M,K] bf16, w [N,K] bf16, bias [N] fp32."""
    _cuda(a, w, bias, out, gate, row_cls)
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        assert mode != EPI_F32_RESIDUAL
        out = torch.empty(M, N, device=a.device, dtype=torch.float32 if mode == EPI_F32 else torch.bfloat16)
    if max_splits > 1:
        assert mode == EPI_F32_RESIDUAL
        check(lib().ltx2_gemm_bf16_splitk(ptr(a), C.c_int64(a.stride(0)), ptr(w), C.c_int64(w.stride(0)), M, N, K,
                                          ptr(bias), ptr(out), C.c_int64(out.stride(0)), ptr(gate),
                                          C.c_int64(gate.stride(0) if gate is not None else 0), ptr(row_cls),
                                          C.c_float(alpha), max_splits, stream_ptr()), "ltx2_gemm_bf16_splitk")
        return out
    check(lib().ltx2_gemm_bf16(ptr(a), C.c_int64(a.stride(0)), ptr(w), C.c_int64(w.stride(0)), M, N, K, mode,
                               ptr(bias), ptr(out), C.c_int64(out.stride(0)), ptr(gate),
                               C.c_int64(gate.stride(0) if gate is not None else 0), ptr(row_cls),
                               C.c_float(alpha), stream_ptr()), "ltx2_gemm_bf16")
    return out


def attention(q: torch.Tensor, k: torch.Tensor, vt: torch.Tensor, tk: int, *, gate_logits=None, want_lse=False):
    """q [B,H,Tq,Dh], k [B,H,Tk,Dh], vt [B,H,Dh,Tkp] bf16 -> out [B,Tq,H*Dh] bf16 (and lse [B,H,Tq])."""
    _cuda(q, k, vt, gate_logits)
    B, H, Tq, Dh = q.shape
    Tkp = vt.shape[-1]
    out = torch.empty(B, Tq, H * Dh, device=q.device, dtype=torch.bfloat16)
    lse = torch.empty(B, H, Tq, device=q.device, dtype=torch.float32) if want_lse else None
    check(lib().ltx2_attention(ptr(q), ptr(k), ptr(vt), ptr(out), B, H, Tq, tk, Tkp, Dh,
                               C.c_float(1.0 / math.sqrt(Dh)), ptr(gate_logits), ptr(lse), stream_ptr()),
          "ltx2_attention")
    return (out, lse) if want_lse else out


def attention_vrows(q: torch.Tensor, k: torch.Tensor, v_rows: torch.Tensor, H: int, Dh: int, *, gate_logits=None,
                    want_lse: bool = False):
    """V straight from a token-major buffer: v_rows [B, Tk, >= H*Dh] view (last dim contiguous), e.g. qkv[..., 2*inner:]."""
    _cuda(q, k, gate_logits)
    B, _, Tq, _ = q.shape
    Tk = k.shape[2]
    assert v_rows.is_cuda and v_rows.stride(2) == 1
    out = torch.empty(B, Tq, H * Dh, device=q.device, dtype=torch.bfloat16)
    lse = torch.empty(B, H, Tq, device=q.device, dtype=torch.float32) if want_lse else None
    check(lib().ltx2_attention_vrows(ptr(q), ptr(k), ptr(v_rows), C.c_int64(v_rows.stride(1)), C.c_int64(Dh),
                                     C.c_int64(v_rows.stride(0)), ptr(out), B, H, Tq, Tk, Dh,
                                     C.c_float(1.0 / math.sqrt(Dh)), ptr(gate_logits), ptr(lse), stream_ptr()),
          "ltx2_attention_vrows")
    return (out, lse) if want_lse else out


def norm_modulate(x: torch.Tensor, *, kind: int, eps: float = 1e-6, mod: Optional[torch.Tensor] = None,
                  shift_row: int = 0, scale_row: int = 1, row_cls: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [M,D] fp32|bf16; mod [n_cls, rows, D] fp32 -> bf16 [M,D]."""
    _cuda(x, mod, row_cls)
    M, D = x.shape
    out = torch.empty(M, D, device=x.device, dtype=torch.bfloat16)
    ms = mod.stride(0) if mod is not None else 0
    check(lib().ltx2_norm_modulate(ptr(x), dtype_code(x), C.c_int64(x.stride(0)), ptr(out), C.c_int64(D), M, D, kind,
                                   C.c_float(eps), ptr(mod), C.c_int64(ms), C.c_int64(shift_row * D),
                                   C.c_int64(scale_row * D), ptr(row_cls), stream_ptr()), "ltx2_norm_modulate")
    return out


def headnorm_rope(x: torch.Tensor, weight: torch.Tensor, B: int, T: int, H: int, Dh: int,
                  cos: Optional[torch.Tensor] = None, sin: Optional[torch.Tensor] = None, eps: float = 1e-6):
    """x [B*T, >=H*Dh] bf16 (row pitch = x.stride(0)) -> [B,H,T,Dh] bf16."""
    _cuda(weight, cos, sin)
    assert x.is_cuda and x.stride(1) == 1
    out = torch.empty(B, H, T, Dh, device=x.device, dtype=torch.bfloat16)
    check(lib().ltx2_headnorm_rope(ptr(x), C.c_int64(x.stride(0)), ptr(weight), ptr(cos), ptr(sin), ptr(out), B, T, H,
                                   Dh, C.c_float(eps), stream_ptr()), "ltx2_headnorm_rope")
    return out


def v_transpose(v: torch.Tensor, B: int, T: int, H: int, Dh: int) -> torch.Tensor:
    assert v.is_cuda and v.stride(1) == 1
    Tp = (T + 63) // 64 * 64
    out = torch.empty(B, H, Dh, Tp, device=v.device, dtype=torch.bfloat16)
    check(lib().ltx2_v_transpose(ptr(v), C.c_int64(v.stride(0)), ptr(out), B, T, Tp, H, Dh, stream_ptr()),
          "ltx2_v_transpose")
    return out


def rope_tables(positions: torch.Tensor, dim: int, max_pos, theta: float = 10000.0):
    """positions [B,n_dims,T,2] fp32 -> cos, sin [B,T,dim/2] fp32 (token-major)."""
    _cuda(positions)
    B, n_dims, T, _ = positions.shape
    cos = torch.empty(B, T, dim // 2, device=positions.device, dtype=torch.float32)
    sin = torch.empty_like(cos)
    mp = (C.c_float * 3)(*([float(m) for m in max_pos] + [1.0] * (3 - len(max_pos))))
    check(lib().ltx2_rope_tables(ptr(positions), B, n_dims, T, dim, mp, C.c_float(theta), ptr(cos), ptr(sin),
                                 stream_ptr()), "ltx2_rope_tables")
    return cos, sin


def timestep_sinusoid(t: torch.Tensor, multiplier: float = 1000.0) -> torch.Tensor:
    _cuda(t)
    out = torch.empty(t.numel(), 256, device=t.device, dtype=torch.float32)
    check(lib().ltx2_timestep_sinusoid(ptr(t), t.numel(), C.c_float(multiplier), ptr(out), stream_ptr()),
          "ltx2_timestep_sinusoid")
    return out


def small_linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], act_in: int = 0) -> torch.Tensor:
    _cuda(x, w, bias)
    R, K = x.shape
    N = w.shape[0]
    y = torch.empty(R, N, device=x.device, dtype=torch.float32)
    check(lib().ltx2_small_linear(ptr(x), R, K, ptr(w), ptr(bias), ptr(y), N, act_in, stream_ptr()),
          "ltx2_small_linear")
    return y


def x0_from_velocity(latent: torch.Tensor, velocity: torch.Tensor, t_row: torch.Tensor) -> torch.Tensor:
    _cuda(latent, velocity, t_row)
    M, Cc = latent.shape
    out = torch.empty_like(latent)
    check(lib().ltx2_x0_from_velocity(ptr(latent), ptr(velocity), ptr(t_row), ptr(out), M, Cc, stream_ptr()),
          "ltx2_x0_from_velocity")
    return out


def _binary(fn_name, a, b):
    _cuda(a, b)
    assert a.shape == b.shape and a.dtype == b.dtype, f"Shape mismatch: {tuple(a.shape)} vs {tuple(b.shape)}"
    out = torch.empty_like(a)
    check(getattr(lib(), fn_name)(ptr(a), ptr(b), ptr(out), C.c_int64(a.numel()), dtype_code(a), stream_ptr()), fn_name)
    return out


def silu_mul(a, b):
    return _binary("ltx2_silu_mul", a, b)


def gelu_mul(a, b):
    return _binary("ltx2_gelu_mul", a, b)


def interleaved_rope(x, cos, sin):
    _cuda(x, cos, sin)
    out = torch.empty_like(x)
    check(lib().ltx2_interleaved_rope(ptr(x), ptr(cos), ptr(sin), ptr(out), C.c_int64(x.numel()), dtype_code(x),
                                      stream_ptr()), "ltx2_interleaved_rope")
    return out


def conv3d(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, causal: bool = False) -> torch.Tensor:
    """Conv3dSimple (simple_decoder.py:90-180) as one op: x [B,T,H,W,Cin] bf16 channels-last, weight [Cout,Cin,3,3,3]
    (PyTorch layout, any float dtype), bias [Cout] -> [B,T,H,W,Cout] bf16."""
    _cuda(x, weight, bias)
    assert x.dtype == torch.bfloat16 and x.ndim == 5 and weight.ndim == 5
    B, T, H, W, Cin = x.shape
    Cout = weight.shape[0]
    out = torch.empty(B, T, H, W, Cout, device=x.device, dtype=torch.bfloat16)
    nbytes = int(lib().ltx2_conv3d_workspace_bytes(B, T, H, W, Cin, Cout))
    ws = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
    check(lib().ltx2_conv3d(ptr(x), ptr(weight), dtype_code(weight), ptr(bias), dtype_code(bias), ptr(out), B, T, H, W,
                            Cin, Cout, int(causal), ptr(ws), stream_ptr()), "ltx2_conv3d")
    return out


# ---- FP8 (E4M3) path --------------------------------------------------------------------------------------------
def quantize_rows_e4m3(w: torch.Tensor):
    """[rows, K] float tensor -> (E4M3 bytes as torch.float8_e4m3fn [rows, K], per-row scales fp32 [rows])."""
    _cuda(w)
    rows, K = w.shape
    out = torch.empty(rows, K, device=w.device, dtype=torch.uint8)
    scale = torch.empty(rows, device=w.device, dtype=torch.float32)
    check(lib().ltx2_quantize_rows_e4m3(ptr(w), dtype_code(w), rows, K, ptr(out), ptr(scale), stream_ptr()),
          "ltx2_quantize_rows_e4m3")
    return out.view(torch.float8_e4m3fn), scale


def norm_modulate_q8(x: torch.Tensor, *, kind: int, eps: float = 1e-6, mod: Optional[torch.Tensor] = None,
                     shift_row: int = 0, scale_row: int = 1, row_cls: Optional[torch.Tensor] = None,
                     want_bf16: bool = False):
    """norm_modulate whose output row is quantised to E4M3 with a per-row scale: (q [M,D] float8_e4m3fn, scale [M]
    [, bf16 copy])."""
    _cuda(x, mod, row_cls)
    M, D = x.shape
    q = torch.empty(M, D, device=x.device, dtype=torch.uint8)
    sc = torch.empty(M, device=x.device, dtype=torch.float32)
    o16 = torch.empty(M, D, device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    ms = mod.stride(0) if mod is not None else 0
    check(lib().ltx2_norm_modulate_q8(ptr(x), dtype_code(x), x.stride(0), ptr(q), D, ptr(sc), ptr(o16), D, M, D, kind,
                                      eps, ptr(mod), ms, shift_row * D, scale_row * D, ptr(row_cls), stream_ptr()),
          "ltx2_norm_modulate_q8")
    q = q.view(torch.float8_e4m3fn)
    return (q, sc, o16) if want_bf16 else (q, sc)


def gemm_e4m3(a8: torch.Tensor, a_scale: torch.Tensor, w8: torch.Tensor, w_scale: torch.Tensor,
              bias: Optional[torch.Tensor] = None, *, mode: int = EPI_BF16) -> torch.Tensor:
    """C = (A8 W8^T) * a_scale[:, None] * w_scale[None, :] (+ bias, epilogue modes 0..2) on the FP8 tensor pipe."""
    _cuda(a8, w8, a_scale, w_scale, bias)
    M, K = a8.shape
    N = w8.shape[0]
    out = torch.empty(M, N, device=a8.device, dtype=torch.float32 if mode == EPI_F32 else torch.bfloat16)
    check(lib().ltx2_gemm_e4m3(ptr(a8), a8.stride(0), ptr(w8), w8.stride(0), M, N, K, mode, ptr(bias), ptr(a_scale),
                               ptr(w_scale), ptr(out), out.stride(0), stream_ptr()), "ltx2_gemm_e4m3")
    return out
