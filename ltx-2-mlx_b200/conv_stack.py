"""Shared host plumbing of the VAE encoder and the latent upscaler: packed conv weights and thin wrappers over the C-ABI
building blocks of include/ltx2_b200.h (section "ENCODER and SPATIAL UPSCALER building blocks").  Nothing here
computes; every function launches kernels of libltx2_b200.so on torch's current stream."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from ._lib import check, dtype_code, lib, ptr, stream_ptr

HW_REFLECT, HW_ZERO = 0, 1
T_REPLICATE, T_CAUSAL, T_ZERO = 0, 1, 2
ACT_NONE, ACT_PIXELNORM_SILU, ACT_GROUPNORM, ACT_GROUPNORM_SILU = 0, 1, 2, 3


class PackedConv:
    """One 3x3x3 conv in the tcgen05 kernel's weight layout (bf16 [Cout_pad, 27*Cin], fp32 bias [Cout_pad])."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor, device):
        weight, bias = weight.to(device), bias.to(device)
        if weight.ndim == 4:                                   # per-frame conv2d (spatial.py:294-323): only the centre
            w5 = torch.zeros(weight.shape[0], weight.shape[1], 3, 3, 3, device=device, dtype=weight.dtype)   # time tap
            w5[:, :, 1] = weight
            weight = w5
        cout, cin = weight.shape[:2]
        if cin % 64:                                           # the conv's K granule is 64 channels: zero-pad C_in
            pad = 64 - cin % 64
            weight = torch.cat([weight, torch.zeros(cout, pad, 3, 3, 3, device=device, dtype=weight.dtype)], dim=1)
        if weight.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            weight = weight.float()
        if bias.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            bias = bias.float()
        weight, bias = weight.contiguous(), bias.contiguous()
        self.cin, self.cout = weight.shape[1], cout
        self.cout_pad = (cout + 31) // 32 * 32
        self.w = torch.empty(self.cout_pad, 27 * self.cin, device=device, dtype=torch.bfloat16)
        self.b = torch.empty(self.cout_pad, device=device, dtype=torch.float32)
        check(lib().ltx2_conv3d_pack(ptr(weight), dtype_code(weight), ptr(bias), dtype_code(bias), cout, self.cout_pad,
                                     self.cin, ptr(self.w), ptr(self.b), stream_ptr()), "ltx2_conv3d_pack")

    def __call__(self, x_padded: torch.Tensor, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        B, Tp, Hp, Wp, C = x_padded.shape
        assert C == self.cin, (C, self.cin)
        T, H, W = Tp - 2, Hp - 2, Wp - 2
        out = torch.empty(B, T, H, W, self.cout, device=x_padded.device, dtype=torch.bfloat16)
        check(lib().ltx2_conv3d_packed(ptr(x_padded), ptr(self.w), ptr(self.b), ptr(out), ptr(residual), B, T, H, W,
                                       self.cin, self.cout, self.cout_pad, stream_ptr()), "ltx2_conv3d_packed")
        return out


def pad_act(x: torch.Tensor, *, hw_mode: int, t_mode: int, act: int = ACT_NONE, dup_first: bool = False,
            gn: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor, int]] = None, eps: float = 1e-6,
            residual: Optional[torch.Tensor] = None, want_plain: bool = False):
    """x [B,T,H,W,C] bf16 -> padded [B,T(+1)+2,H+2,W+2,C] (and the un-padded activated tensor when want_plain)."""
    B, T, H, W, C = x.shape
    TL = T + (1 if dup_first else 0)
    out = torch.empty(B, TL + 2, H + 2, W + 2, C, device=x.device, dtype=torch.bfloat16)
    plain = torch.empty_like(x) if want_plain else None
    stats, gw, gb, groups = gn if gn is not None else (None, None, None, 0)
    check(lib().ltx2_pad_act(ptr(x), ptr(out), ptr(plain), B, T, H, W, C, hw_mode, t_mode, int(dup_first), act, ptr(stats),
                             ptr(gw), ptr(gb), groups, eps, ptr(residual), stream_ptr()), "ltx2_pad_act")
    return (out, plain) if want_plain else out


def group_stats(x: torch.Tensor, groups: int, eps: float = 1e-5) -> torch.Tensor:
    B, T, H, W, C = x.shape
    stats = torch.empty(B, groups, 2, device=x.device, dtype=torch.float32)
    check(lib().ltx2_group_stats(ptr(x), B, T * H * W, C, groups, eps, ptr(stats), stream_ptr()), "ltx2_group_stats")
    return stats


def to_ncdhw_f32(x: torch.Tensor, channels: int, mean: Optional[torch.Tensor] = None,
                 std: Optional[torch.Tensor] = None) -> torch.Tensor:
    B, T, H, W, Cs = x.shape
    out = torch.empty(B, channels, T, H, W, device=x.device, dtype=torch.float32)
    check(lib().ltx2_ndhwc_to_ncdhw(ptr(x), ptr(out), B, channels, Cs, T * H * W, ptr(mean), ptr(std), stream_ptr()),
          "ltx2_ndhwc_to_ncdhw")
    return out


def to_ndhwc_bf16(x: torch.Tensor) -> torch.Tensor:
    B, C, T, H, W = x.shape
    out = torch.empty(B, T, H, W, C, device=x.device, dtype=torch.bfloat16)
    check(lib().ltx2_ncdhw_to_ndhwc(ptr(x), dtype_code(x), ptr(out), B, C, T * H * W, stream_ptr()), "ltx2_ncdhw_to_ndhwc")
    return out


class ConvCollector:
    """Pairs `<prefix>.weight` / `<prefix>.bias` tensors as they stream in and packs each conv once both are there."""

    def __init__(self, device):
        self.device = device
        self.pending: Dict[str, Dict[str, torch.Tensor]] = {}
        self.convs: Dict[str, PackedConv] = {}

    def add(self, prefix: str, kind: str, tensor: torch.Tensor, out_rows: Optional[int] = None) -> None:
        slot = self.pending.setdefault(prefix, {})
        slot[kind] = tensor if out_rows is None else tensor[:out_rows]
        if "weight" in slot and "bias" in slot:
            self.convs[prefix] = PackedConv(slot["weight"], slot["bias"], self.device)
            del self.pending[prefix]
