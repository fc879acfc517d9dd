"""Host-side mirror of the reference's 2x latent SPATIAL UPSCALER, backed by the sm_100a kernels.

Same names and call surface as ``LTX_2_MLX/model/upscaler/spatial.py``: ``SpatialUpscaler`` (:326-412,
``upscaler(latent (B,128,F,H,W)) -> (B,128,F,2H,2W)``) and ``load_spatial_upscaler_weights(upscaler, path)`` (:414-538) --
stage 2 of the distilled / two-stage pipelines (pipelines/distilled.py:394-405).  The 3x3x3 convs (and the per-frame 3x3
conv of the resampler, as a 3x3x3 conv whose outer time taps are zero) run on the tcgen05 implicit-GEMM kernel;
GroupNorm statistics, the fused GroupNorm-affine + residual + SiLU + zero padding pass and the pixel shuffle are
csrc/vae_aux.cu.
"""
from __future__ import annotations

from typing import Any, Dict, Iterable, List, Tuple

import torch

from . import conv_stack as cs
from ._lib import check, lib, ptr, stream_ptr
from .transformer import to_device


class SpatialUpscaler:
    def __init__(self, in_channels: int = 128, mid_channels: int = 1024, num_blocks_per_stage: int = 4,
                 num_groups: int = 32, device="cuda"):
        self.in_channels, self.mid_channels = in_channels, mid_channels
        self.num_blocks_per_stage, self.num_groups = num_blocks_per_stage, num_groups
        self.device = torch.device(device)
        self._convs = cs.ConvCollector(self.device)
        self._norms: Dict[str, Dict[str, torch.Tensor]] = {}

    # ---- weights (checkpoint keys of spatial.py:414-538) --------------------------------------------------------
    def _conv_names(self) -> List[str]:
        names = ["initial_conv", "final_conv", "upsampler.conv"]
        for stage in ("res_blocks", "post_upsample_res_blocks"):
            names += [f"{stage}.{i}.conv{k}" for i in range(self.num_blocks_per_stage) for k in (1, 2)]
        return names

    def _norm_names(self) -> List[str]:
        names = ["initial_norm"]
        for stage in ("res_blocks", "post_upsample_res_blocks"):
            names += [f"{stage}.{i}.norm{k}" for i in range(self.num_blocks_per_stage) for k in (1, 2)]
        return names

    def load_weights(self, weights: Iterable[Tuple[str, Any]]) -> int:
        n = 0
        convs, norms = set(self._conv_names()), set(self._norm_names())
        with torch.cuda.device(self.device):
            for key, value in (weights.items() if isinstance(weights, dict) else weights):
                if "." not in key:
                    continue
                prefix, kind = key.rsplit(".", 1)
                if prefix in convs and kind in ("weight", "bias"):
                    self._convs.add(prefix, kind, to_device(value, self.device))
                elif prefix in norms and kind in ("weight", "bias"):
                    self._norms.setdefault(prefix, {})[kind] = to_device(value, self.device, torch.float32)
                else:
                    continue
                n += 1
            torch.cuda.current_stream().synchronize()
        return n

    def missing_weights(self) -> List[str]:
        miss = [p for p in self._conv_names() if p not in self._convs.convs]
        miss += [p for p in self._norm_names() if set(self._norms.get(p, {})) != {"weight", "bias"}]
        return miss

    # ---- forward ------------------------------------------------------------------------------------------------
    def _gn(self, h: torch.Tensor, name: str):
        nw = self._norms[name]
        return cs.group_stats(h, self.num_groups, 1e-5), nw["weight"], nw["bias"], self.num_groups

    def _res_block(self, prefix: str, xp: torch.Tensor, cur: torch.Tensor):
        """ResBlock3d (spatial.py:158-181): conv1 -> norm1 -> SiLU -> conv2 -> norm2 -> SiLU(. + x).
        xp = zero-padded `cur`; returns the padded and plain block output."""
        conv, pad = self._convs.convs, dict(hw_mode=cs.HW_ZERO, t_mode=cs.T_ZERO)
        h = conv[f"{prefix}.conv1"](xp)
        hp = cs.pad_act(h, act=cs.ACT_GROUPNORM_SILU, gn=self._gn(h, f"{prefix}.norm1"), eps=1e-5, **pad)
        h2 = conv[f"{prefix}.conv2"](hp)
        return cs.pad_act(h2, act=cs.ACT_GROUPNORM_SILU, gn=self._gn(h2, f"{prefix}.norm2"), eps=1e-5, residual=cur,
                          want_plain=True, **pad)

    def __call__(self, x) -> torch.Tensor:
        lat = to_device(x, self.device)
        if lat.ndim != 5 or lat.shape[1] != self.in_channels:
            raise ValueError(f"latent must be (B, {self.in_channels}, F, H, W); got {tuple(lat.shape)}")
        miss = self.missing_weights()
        if miss:
            raise RuntimeError(f"upscaler weights missing: {miss[:4]}")
        conv, pad = self._convs.convs, dict(hw_mode=cs.HW_ZERO, t_mode=cs.T_ZERO)
        with torch.cuda.device(self.device):
            cin = conv["initial_conv"].cin
            if lat.shape[1] < cin:            # the conv's K granule is 64 channels: zero channels match zero weight columns
                lat = torch.cat([lat, torch.zeros(lat.shape[0], cin - lat.shape[1], *lat.shape[2:], device=self.device,
                                                  dtype=lat.dtype)], dim=1).contiguous()
            h = conv["initial_conv"](cs.pad_act(cs.to_ndhwc_bf16(lat), **pad))
            xp, cur = cs.pad_act(h, act=cs.ACT_GROUPNORM_SILU, gn=self._gn(h, "initial_norm"), eps=1e-5, want_plain=True,
                                 **pad)
            for i in range(self.num_blocks_per_stage):
                xp, cur = self._res_block(f"res_blocks.{i}", xp, cur)
            # SpatialRationalResampler (scale 2): per-frame conv C -> 4C, pixel shuffle; the stride-1 blur is an identity
            y = conv["upsampler.conv"](xp)
            B, F, H, W, C4 = y.shape
            cur = torch.empty(B, F, 2 * H, 2 * W, C4 // 4, device=self.device, dtype=torch.bfloat16)
            check(lib().ltx2_pixel_shuffle2(ptr(y), ptr(cur), B * F, H, W, C4 // 4, stream_ptr()), "ltx2_pixel_shuffle2")
            xp = cs.pad_act(cur, **pad)
            for i in range(self.num_blocks_per_stage):
                xp, cur = self._res_block(f"post_upsample_res_blocks.{i}", xp, cur)
            return cs.to_ncdhw_f32(conv["final_conv"](xp), self.in_channels)


def load_spatial_upscaler_weights(upscaler: SpatialUpscaler, weights_path: str) -> None:
    """Drop-in for spatial.load_spatial_upscaler_weights: safetensors file -> engine."""
    from safetensors import safe_open

    print(f"Loading Spatial Upscaler weights from {weights_path}...")
    with safe_open(weights_path, framework="pt") as f:
        n = upscaler.load_weights((k, f.get_tensor(k)) for k in f.keys())
    print(f"  Loaded {n} weight tensors")
