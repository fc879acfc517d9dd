"""Host-side mirror of the reference's video-VAE ENCODER, backed by the sm_100a kernels.

Same names, call surface and error behaviour as ``LTX_2_MLX/model/video_vae/simple_encoder.py``: ``SimpleVideoEncoder``
(:258-405, ``encoder(video) -> normalised latent``), ``load_vae_encoder_weights(encoder, path)`` (:407-532) and
``encode_video`` -- what image conditioning and the second stage of the two-stage pipelines call
(pipelines/distilled.py:394-405).  Every conv runs on the tcgen05 implicit-GEMM kernel (csrc/conv3d_sm100.cu); padding,
pixel-norm + SiLU, space-to-depth with the group-mean residual and the latent normalisation are csrc/vae_aux.cu.
Activations are channels-last bf16 between convs (fp32 accumulation), like the decoder engine.
"""
from __future__ import annotations

from typing import Any, Iterable, List, Tuple

import torch

from . import conv_stack as cs
from ._lib import check, lib, ptr, stream_ptr
from .transformer import to_device

# simple_encoder.py:296-305: (kind, channels in, block count | channels out, stride)
ENCODER_BLOCKS = [("res", 128, 4, None), ("down", 128, 256, (1, 2, 2)), ("res", 256, 6, None),
                  ("down", 256, 512, (2, 1, 1)), ("res", 512, 6, None), ("down", 512, 1024, (2, 2, 2)),
                  ("res", 1024, 2, None), ("down", 1024, 1024, (2, 2, 2)), ("res", 1024, 2, None)]


class SimpleVideoEncoder:
    """video (B, 3, F, H, W) in [-1, 1], F = 1 + 8k  ->  normalised latent (B, 128, 1 + k, H/32, W/32) fp32 on the GPU."""

    def __init__(self, compute_dtype: Any = None, device="cuda"):
        self.compute_dtype = compute_dtype
        self.patch_size = 4
        self.device = torch.device(device)
        self._convs = cs.ConvCollector(self.device)
        self.mean_of_means = None
        self.std_of_means = None

    # ---- weights ----------------------------------------------------------------------------------------------
    def expected_convs(self) -> List[str]:
        names = ["vae.encoder.conv_in.conv", "vae.encoder.conv_out.conv"]
        for idx, (kind, _, n, _) in enumerate(ENCODER_BLOCKS):
            P = f"vae.encoder.down_blocks.{idx}"
            if kind == "res":
                names += [f"{P}.res_blocks.{j}.conv{k}.conv" for j in range(n) for k in (1, 2)]
            else:
                names.append(f"{P}.conv.conv")
        return names

    def load_weights(self, weights: Iterable[Tuple[str, Any]]) -> int:
        """(checkpoint_key, tensor) pairs under the names simple_encoder.py:407-532 reads; others are skipped."""
        n = 0
        with torch.cuda.device(self.device):
            for key, value in (weights.items() if isinstance(weights, dict) else weights):
                if key == "vae.per_channel_statistics.mean-of-means":
                    self.mean_of_means = to_device(value, self.device, torch.float32)
                elif key == "vae.per_channel_statistics.std-of-means":
                    self.std_of_means = to_device(value, self.device, torch.float32)
                elif key.startswith("vae.encoder.") and (key.endswith(".weight") or key.endswith(".bias")):
                    prefix, kind = key.rsplit(".", 1)
                    # conv_out has 129 rows (128 means + 1 shared log-variance); only the means are used (:396-398)
                    rows = 128 if prefix == "vae.encoder.conv_out.conv" else None
                    self._convs.add(prefix, kind, to_device(value, self.device), rows)
                else:
                    continue
                n += 1
            torch.cuda.current_stream().synchronize()
        return n

    def missing_weights(self) -> List[str]:
        miss = [p for p in self.expected_convs() if p not in self._convs.convs]
        if self.mean_of_means is None or self.std_of_means is None:
            miss.append("vae.per_channel_statistics")
        return miss

    # ---- forward ----------------------------------------------------------------------------------------------
    def __call__(self, video, show_progress: bool = True) -> torch.Tensor:
        v = to_device(video, self.device, torch.float32)
        if v.ndim != 5 or v.shape[1] != 3:
            raise ValueError(f"video must be (B, 3, F, H, W); got {tuple(v.shape)}")
        B, _, F, H, W = v.shape
        if (F - 1) % 8 != 0:
            raise ValueError(f"Invalid number of frames: {F}. "
                             f"Encoder input must have 1 + 8*k frames (e.g., 1, 9, 17, 25, 33...).")
        if H % 32 or W % 32:
            raise ValueError(f"video height/width must be multiples of 32; got {H}x{W}")
        miss = self.missing_weights()
        if miss:
            raise RuntimeError(f"encoder weights missing: {miss[:4]}")
        conv = self._convs.convs
        pad = dict(hw_mode=cs.HW_ZERO, t_mode=cs.T_CAUSAL)
        with torch.cuda.device(self.device):
            x = torch.empty(B, F, H // 4, W // 4, 64, device=self.device, dtype=torch.bfloat16)
            check(lib().ltx2_patchify_video(ptr(v), ptr(x), B, F, H, W, 64, stream_ptr()), "ltx2_patchify_video")
            cur = conv["vae.encoder.conv_in.conv"](cs.pad_act(x, **pad))
            for idx, (kind, c_in, n_or_cout, stride) in enumerate(ENCODER_BLOCKS):
                P = f"vae.encoder.down_blocks.{idx}"
                if kind == "res":
                    for j in range(n_or_cout):        # EncoderResBlock3d, simple_encoder.py:132-154
                        h = conv[f"{P}.res_blocks.{j}.conv1.conv"](cs.pad_act(cur, act=cs.ACT_PIXELNORM_SILU, **pad))
                        cur = conv[f"{P}.res_blocks.{j}.conv2.conv"](cs.pad_act(h, act=cs.ACT_PIXELNORM_SILU, **pad),
                                                                     residual=cur)
                else:                                  # SpaceToDepthDownsample3d, simple_encoder.py:226-257
                    st, sh, sw = stride
                    dup = st == 2
                    y = conv[f"{P}.conv.conv"](cs.pad_act(cur, dup_first=dup, **pad))
                    Bc, T, Hc, Wc, Cx = cur.shape
                    TL = T + (1 if dup else 0)
                    out = torch.empty(Bc, TL // st, Hc // sh, Wc // sw, n_or_cout, device=self.device,
                                      dtype=torch.bfloat16)
                    check(lib().ltx2_space_to_depth_residual(ptr(y), ptr(cur), ptr(out), Bc, T, Hc, Wc, Cx, n_or_cout, st,
                                                             sh, sw, int(dup), stream_ptr()),
                          "ltx2_space_to_depth_residual")
                    cur = out
            y = conv["vae.encoder.conv_out.conv"](cs.pad_act(cur, act=cs.ACT_PIXELNORM_SILU, **pad))
            return cs.to_ncdhw_f32(y, 128, self.mean_of_means, self.std_of_means)


def load_vae_encoder_weights(encoder: SimpleVideoEncoder, weights_path: str) -> None:
    """Drop-in for simple_encoder.load_vae_encoder_weights: safetensors file -> engine."""
    from safetensors import safe_open

    print(f"Loading VAE encoder weights from {weights_path}...")
    with safe_open(weights_path, framework="pt") as f:
        def gen():
            for k in f.keys():
                if k.startswith("vae.encoder.") or k.startswith("vae.per_channel_statistics."):
                    yield k, f.get_tensor(k)
        n = encoder.load_weights(gen())
    print(f"  Loaded {n} weight tensors")


def encode_video(video, encoder: SimpleVideoEncoder) -> torch.Tensor:
    """simple_encoder.py:535-560: frames (F, H, W, 3) uint8 or float in [0, 1|255] -> latent (1, 128, F', H', W')."""
    v = to_device(video, encoder.device, torch.float32)
    if v.ndim == 4:
        if float(v.max()) > 1.0:
            v = v / 255.0
        v = (v * 2.0 - 1.0).permute(3, 0, 1, 2)[None]
    return encoder(v.contiguous())
