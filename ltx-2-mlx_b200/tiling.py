"""Tiled VAE decoding -- drop-in for ``LTX_2_MLX/model/video_vae/tiling.py``.

Same names and semantics: ``SpatialTilingConfig`` / ``TemporalTilingConfig`` / ``TilingConfig`` (:55-122, same
validation errors), ``TileSpec`` / ``generate_tile_specs`` (:125-249), ``compute_trapezoidal_mask_1d`` (:9-52) and
``decode_tiled(latent, decoder_fn, tiling_config, timestep)`` (:252-412), which yields the blended video once.
Differences that do not change results: every tile is decoded ONCE (the reference decodes each tile twice and discards
the first pass, :299-347), and the weighted accumulation / normalisation run as two CUDA kernels on one preallocated
buffer (``ltx2_tile_accumulate`` / ``ltx2_tile_normalize``) instead of rebuilding the output by concatenation per tile.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterator, List, Optional, Tuple

import torch

from ._lib import check, lib, ptr, stream_ptr
from .transformer import to_device


def compute_trapezoidal_mask_1d(length: int, ramp_left: int, ramp_right: int, left_starts_from_0: bool = False) -> torch.Tensor:
    if length <= 0:
        raise ValueError("Mask length must be positive.")
    ramp_left = max(0, min(ramp_left, length))
    ramp_right = max(0, min(ramp_right, length))
    mask = torch.ones(length, dtype=torch.float32)
    if ramp_left > 0:
        n = ramp_left + 1 if left_starts_from_0 else ramp_left + 2
        fade_in = torch.linspace(0.0, 1.0, n)[:-1]
        if not left_starts_from_0:
            fade_in = fade_in[1:]
        mask = torch.cat([fade_in, mask[ramp_left:]])
    if ramp_right > 0:
        mask = torch.cat([mask[:-ramp_right], torch.linspace(1.0, 0.0, ramp_right + 2)[1:-1]])
    return mask.clamp(0, 1)


@dataclass(frozen=True)
class SpatialTilingConfig:
    tile_size_in_pixels: int
    tile_overlap_in_pixels: int = 0

    def __post_init__(self) -> None:
        if self.tile_size_in_pixels < 64:
            raise ValueError(f"tile_size_in_pixels must be at least 64, got {self.tile_size_in_pixels}")
        if self.tile_size_in_pixels % 32 != 0:
            raise ValueError(f"tile_size_in_pixels must be divisible by 32, got {self.tile_size_in_pixels}")
        if self.tile_overlap_in_pixels % 32 != 0:
            raise ValueError(f"tile_overlap_in_pixels must be divisible by 32, got {self.tile_overlap_in_pixels}")
        if self.tile_overlap_in_pixels >= self.tile_size_in_pixels:
            raise ValueError(f"Overlap must be less than tile size, got {self.tile_overlap_in_pixels} and "
                             f"{self.tile_size_in_pixels}")


@dataclass(frozen=True)
class TemporalTilingConfig:
    tile_size_in_frames: int
    tile_overlap_in_frames: int = 0

    def __post_init__(self) -> None:
        if self.tile_size_in_frames < 16:
            raise ValueError(f"tile_size_in_frames must be at least 16, got {self.tile_size_in_frames}")
        if self.tile_size_in_frames % 8 != 0:
            raise ValueError(f"tile_size_in_frames must be divisible by 8, got {self.tile_size_in_frames}")
        if self.tile_overlap_in_frames % 8 != 0:
            raise ValueError(f"tile_overlap_in_frames must be divisible by 8, got {self.tile_overlap_in_frames}")
        if self.tile_overlap_in_frames >= self.tile_size_in_frames:
            raise ValueError(f"Overlap must be less than tile size, got {self.tile_overlap_in_frames} and "
                             f"{self.tile_size_in_frames}")


@dataclass(frozen=True)
class TilingConfig:
    spatial_config: Optional[SpatialTilingConfig] = None
    temporal_config: Optional[TemporalTilingConfig] = None

    @classmethod
    def default(cls) -> "TilingConfig":
        return cls(spatial_config=SpatialTilingConfig(tile_size_in_pixels=512, tile_overlap_in_pixels=64),
                   temporal_config=TemporalTilingConfig(tile_size_in_frames=64, tile_overlap_in_frames=24))


@dataclass
class TileSpec:
    in_t_start: int
    in_t_end: int
    in_h_start: int
    in_h_end: int
    in_w_start: int
    in_w_end: int
    out_t_start: int
    out_t_end: int
    out_h_start: int
    out_h_end: int
    out_w_start: int
    out_w_end: int
    ramp_t_left: int
    ramp_t_right: int
    ramp_h_left: int
    ramp_h_right: int
    ramp_w_left: int
    ramp_w_right: int


def _tiles_1d(length: int, tile: int, overlap: int) -> List[Tuple[int, int, int, int]]:
    """(start, end, ramp_left, ramp_right) per tile; the last tile is shifted back to keep its full size."""
    if length <= tile:
        return [(0, length, 0, 0)]
    out, pos, stride = [], 0, tile - overlap
    while pos < length:
        end = min(pos + tile, length)
        start = max(0, end - tile)
        out.append((start, end, overlap if start > 0 else 0, overlap if end < length else 0))
        if end >= length:
            break
        pos += stride
    return out


def generate_tile_specs(latent_shape, tiling_config: TilingConfig, scale_factors=(8, 32, 32)) -> List[TileSpec]:
    _, _, t, h, w = latent_shape
    st, sh, sw = scale_factors
    sc, tc = tiling_config.spatial_config, tiling_config.temporal_config
    th, tw = (sc.tile_size_in_pixels // sh, sc.tile_size_in_pixels // sw) if sc else (h, w)
    oh, ow = (sc.tile_overlap_in_pixels // sh, sc.tile_overlap_in_pixels // sw) if sc else (0, 0)
    tt, ot = (tc.tile_size_in_frames // st, tc.tile_overlap_in_frames // st) if tc else (t, 0)
    specs = []
    for (t0, t1, rtl, rtr) in _tiles_1d(t, tt, ot):
        for (h0, h1, rhl, rhr) in _tiles_1d(h, th, oh):
            for (w0, w1, rwl, rwr) in _tiles_1d(w, tw, ow):
                specs.append(TileSpec(
                    in_t_start=t0, in_t_end=t1, in_h_start=h0, in_h_end=h1, in_w_start=w0, in_w_end=w1,
                    out_t_start=t0 * st if t0 > 0 else 0, out_t_end=(t1 - 1) * st + 1 if t1 > 1 else 1,
                    out_h_start=h0 * sh, out_h_end=h1 * sh, out_w_start=w0 * sw, out_w_end=w1 * sw,
                    ramp_t_left=rtl * st, ramp_t_right=rtr * st, ramp_h_left=rhl * sh, ramp_h_right=rhr * sh,
                    ramp_w_left=rwl * sw, ramp_w_right=rwr * sw))
    return specs


def decode_tiled(latent, decoder_fn, tiling_config: TilingConfig, timestep: Optional[float] = 0.05,
                 show_progress: bool = True, key=None, group=None) -> Iterator[torch.Tensor]:
    """Yields the blended video (B, 3, 8(T-1)+1, 32H, 32W) fp32 once, like the reference generator.

    With a torch.distributed `group` the tiles (independent units) are decoded round-robin across the ranks; each
    rank accumulates its own tiles and the weighted sums are all-reduced before the normalisation, so every rank
    yields the full video."""
    world, rank = 1, 0
    if group is not None:
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = getattr(decoder_fn, "device", None) or torch.device("cuda", torch.cuda.current_device())
    x = to_device(latent, dev)
    b, _, t, h, w = x.shape
    To, Ho, Wo = (t - 1) * 8 + 1, h * 32, w * 32
    out = torch.zeros(b, 3, To, Ho, Wo, device=dev, dtype=torch.float32)
    wsum = torch.zeros(To, Ho, Wo, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        for i, s in enumerate(generate_tile_specs(x.shape, tiling_config)):
            if i % world != rank:
                continue
            tile = decoder_fn(x[:, :, s.in_t_start:s.in_t_end, s.in_h_start:s.in_h_end,
                                s.in_w_start:s.in_w_end].contiguous(), timestep=timestep)
            tile = to_device(tile, dev, torch.float32)
            _, _, dt, dh, dw = tile.shape
            tt = min(dt, s.out_t_end - s.out_t_start)
            th = min(dh, s.out_h_end - s.out_h_start)
            tw = min(dw, s.out_w_end - s.out_w_start)
            mt = compute_trapezoidal_mask_1d(tt, min(s.ramp_t_left, tt), min(s.ramp_t_right, tt),
                                             left_starts_from_0=(s.out_t_start == 0)).to(dev)
            mh = compute_trapezoidal_mask_1d(th, min(s.ramp_h_left, th), min(s.ramp_h_right, th)).to(dev)
            mw = compute_trapezoidal_mask_1d(tw, min(s.ramp_w_left, tw), min(s.ramp_w_right, tw)).to(dev)
            check(lib().ltx2_tile_accumulate(ptr(out), ptr(wsum), ptr(tile), b * 3, To, Ho, Wo, dt, dh, dw,
                                             s.out_t_start, s.out_h_start, s.out_w_start, tt, th, tw, ptr(mt), ptr(mh),
                                             ptr(mw), stream_ptr()), "ltx2_tile_accumulate")
        if world > 1:
            dist.all_reduce(out, group=group)
            dist.all_reduce(wsum, group=group)
        check(lib().ltx2_tile_normalize(ptr(out), ptr(wsum), b * 3, To * Ho * Wo, stream_ptr()), "ltx2_tile_normalize")
    yield out
