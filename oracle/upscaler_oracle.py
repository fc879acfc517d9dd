"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch fp32) of the reference's 2x latent SPATIAL UPSCALER.

SURVEY.md 8(f) rank 3 (stage 2 of the distilled / two-stage pipelines): the checker of the CUDA path
ltx-2-mlx_b200/upscaler.py (tests/test_encoder_upscaler_gpu.py).  Imported only by tests/.  Pinned by tests/golden/upscaler.npz, produced by the reference's own `SpatialUpscaler` (weights
through its own `load_spatial_upscaler_weights`) over the restated mlx primitives of oracle/_mlx_shim.

Reference map (/root/reference/LTX_2_MLX/model/upscaler/spatial.py):
  conv3d (zero padding in T, H and W) ................. :21-88
  group_norm_5d (statistics over C/groups, T, H, W) ... :91-128
  ResBlock3d: conv-norm-silu-conv-norm, silu(x + res) .. :131-181
  SpatialRationalResampler: per-frame 3x3 conv2d -> pixel shuffle (c, r_h, r_w); the stride-1 blur is an identity
                                                        :184-323
  SpatialUpscaler forward ............................. :326-412;  checkpoint keys :414-538
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F


def conv3d(x, weight, bias):
    return F.conv3d(x, weight.to(x.dtype), bias.to(x.dtype), padding=1)


def group_norm_5d(x, groups: int, weight, bias, eps: float = 1e-5):
    b, c, t, h, w = x.shape
    g = x.reshape(b, groups, c // groups, t, h, w)
    mean = g.mean(dim=(2, 3, 4, 5), keepdim=True)
    var = g.var(dim=(2, 3, 4, 5), keepdim=True, unbiased=False)
    g = (g - mean) / torch.sqrt(var + eps)
    return g.reshape(b, c, t, h, w) * weight.reshape(1, -1, 1, 1, 1) + bias.reshape(1, -1, 1, 1, 1)


def res_block(w: Dict[str, torch.Tensor], p: str, x, groups: int):
    h = conv3d(x, w[p + ".conv1.weight"], w[p + ".conv1.bias"])
    h = F.silu(group_norm_5d(h, groups, w[p + ".norm1.weight"], w[p + ".norm1.bias"]))
    h = conv3d(h, w[p + ".conv2.weight"], w[p + ".conv2.bias"])
    h = group_norm_5d(h, groups, w[p + ".norm2.weight"], w[p + ".norm2.bias"])
    return F.silu(h + x)


def resample2x(w, x):
    """Per-frame conv2d (C -> 4C) and pixel shuffle with channel packing (c, r_h, r_w)."""
    b, c, f, h, ww = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, ww)
    y = F.conv2d(y, w["upsampler.conv.weight"].to(y.dtype), w["upsampler.conv.bias"].to(y.dtype), padding=1)
    y = F.pixel_shuffle(y, 2)
    return y.reshape(b, f, c, 2 * h, 2 * ww).permute(0, 2, 1, 3, 4)


def upscale(w: Dict[str, torch.Tensor], latent: torch.Tensor, *, groups: int = 32, blocks: int = 4) -> torch.Tensor:
    """(B,128,F,H,W) -> (B,128,F,2H,2W)"""
    x = conv3d(latent.float(), w["initial_conv.weight"], w["initial_conv.bias"])
    x = F.silu(group_norm_5d(x, groups, w["initial_norm.weight"], w["initial_norm.bias"]))
    for i in range(blocks):
        x = res_block(w, f"res_blocks.{i}", x, groups)
    x = resample2x(w, x)
    for i in range(blocks):
        x = res_block(w, f"post_upsample_res_blocks.{i}", x, groups)
    return conv3d(x, w["final_conv.weight"], w["final_conv.bias"])
