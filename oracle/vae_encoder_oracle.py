"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch fp32) of the reference video-VAE ENCODER.

SURVEY.md 8(f) rank 3 (image conditioning / stage 2 need the encoder): the checker of the CUDA path
ltx-2-mlx_b200/video_vae_encoder.py (tests/test_encoder_upscaler_gpu.py).  Imported only by tests/.  Pinned by
tests/golden/vae_encoder.npz, produced by the reference's own `SimpleVideoEncoder` (weights through its own
`load_vae_encoder_weights`) over the restated mlx primitives of oracle/_mlx_shim (tests/golden/make_golden.py).

Reference map (under /root/reference/LTX_2_MLX/model/video_vae/):
  patchify (c, p_t, r_w, r_h packing) ........ ops.py:9-68
  Conv3dSimple (ZERO pad H/W, causal = first frame twice) . simple_encoder.py:18-117
  EncoderResBlock3d / pixel norm ............. simple_encoder.py:12-15, 120-154
  SpaceToDepthDownsample3d (+ group-mean residual) ... simple_encoder.py:172-257
  encoder forward / latent normalisation ..... simple_encoder.py:260-405, ops.py:173-186
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

# (kind, channels in, channels out or block count, stride)   simple_encoder.py:296-305
ENCODER_BLOCKS = [("res", 128, 4, None), ("down", 128, 256, (1, 2, 2)), ("res", 256, 6, None),
                  ("down", 256, 512, (2, 1, 1)), ("res", 512, 6, None), ("down", 512, 1024, (2, 2, 2)),
                  ("res", 1024, 2, None), ("down", 1024, 1024, (2, 2, 2)), ("res", 1024, 2, None)]


def patchify(x: torch.Tensor, q: int = 4) -> torch.Tensor:
    """ops.py:44-68 -- (B,C,F,H,W) -> (B, C*q*q, F, H/q, W/q), channel order (c, r_w, r_h)."""
    b, c, f, h, w = x.shape
    x = x.reshape(b, c, f, 1, h // q, q, w // q, q)
    x = x.permute(0, 1, 3, 7, 5, 2, 4, 6)                 # (B, C, p, r_w, r_h, F, H/q, W/q)
    return x.reshape(b, c * q * q, f, h // q, w // q)


def conv3d(x, weight, bias, causal: bool = True):
    """simple_encoder.py:46-117 -- 3x3x3, zero padding in H/W, first frame repeated twice in front when causal."""
    x = F.pad(x, (1, 1, 1, 1))
    if causal:
        x = torch.cat([x[:, :, :1], x[:, :, :1], x], dim=2)
    else:
        x = F.pad(x, (0, 0, 0, 0, 1, 1))
    return F.conv3d(x, weight.to(x.dtype), bias.to(x.dtype))


def pixel_norm(x, eps: float = 1e-6):
    """simple_encoder.py:12-15"""
    return x * torch.rsqrt(torch.mean(x * x, dim=1, keepdim=True) + eps)


def res_block(w: Dict[str, torch.Tensor], prefix: str, x):
    """simple_encoder.py:132-154"""
    h = conv3d(F.silu(pixel_norm(x)), w[prefix + ".conv1.conv.weight"], w[prefix + ".conv1.conv.bias"])
    h = conv3d(F.silu(pixel_norm(h)), w[prefix + ".conv2.conv.weight"], w[prefix + ".conv2.conv.bias"])
    return h + x


def space_to_depth(x, stride):
    """simple_encoder.py:207-224 -- channel order (c, st, sh, sw)."""
    b, c, t, h, w = x.shape
    st, sh, sw = stride
    x = x.reshape(b, c, t // st, st, h // sh, sh, w // sw, sw)
    x = x.permute(0, 1, 3, 5, 7, 2, 4, 6)
    return x.reshape(b, c * st * sh * sw, t // st, h // sh, w // sw)


def downsample(w, prefix: str, x, c_out: int, stride):
    """simple_encoder.py:226-257 -- conv, space-to-depth, plus the space-to-depth of the input averaged over
    channel groups; a temporal stride of 2 first duplicates the first frame."""
    if stride[0] == 2:
        x = torch.cat([x[:, :, :1], x], dim=2)
    r = space_to_depth(x, stride)
    b, cr, t, h, ww = r.shape
    r = r.reshape(b, c_out, cr // c_out, t, h, ww).mean(dim=2)
    y = conv3d(x, w[prefix + ".conv.conv.weight"], w[prefix + ".conv.conv.bias"])
    return space_to_depth(y, stride) + r


def vae_encode(w: Dict[str, torch.Tensor], video: torch.Tensor) -> torch.Tensor:
    """simple_encoder.py:308-405 -- video (B,3,F,H,W) in [-1,1], F = 1 + 8k -> normalised latent (B,128,1+k,H/32,W/32)."""
    if (video.shape[2] - 1) % 8 != 0:
        raise ValueError(f"Invalid number of frames: {video.shape[2]}. "
                         "Encoder input must have 1 + 8*k frames (e.g., 1, 9, 17, 25, 33...).")
    x = patchify(video.float(), 4)
    x = conv3d(x, w["vae.encoder.conv_in.conv.weight"], w["vae.encoder.conv_in.conv.bias"])
    for idx, (kind, c_in, n_or_cout, stride) in enumerate(ENCODER_BLOCKS):
        P = f"vae.encoder.down_blocks.{idx}"
        if kind == "res":
            for j in range(n_or_cout):
                x = res_block(w, f"{P}.res_blocks.{j}", x)
        else:
            x = downsample(w, P, x, n_or_cout, stride)
    x = F.silu(pixel_norm(x))
    x = conv3d(x, w["vae.encoder.conv_out.conv.weight"], w["vae.encoder.conv_out.conv.bias"])
    means = x[:, :128]
    std = w["vae.per_channel_statistics.std-of-means"].float().reshape(1, -1, 1, 1, 1)
    mean = w["vae.per_channel_statistics.mean-of-means"].float().reshape(1, -1, 1, 1, 1)
    return (means - mean) / std                                    # PerChannelStatistics.normalize, ops.py:173-186
