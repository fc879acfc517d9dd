"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference video-VAE decode.

Oracle for the second half of the hot path.  Plain
torch-CPU, imported only by tests/, smoke() and bench.py's CPU legs.  Pinned the
same way as oracle/dit_oracle.py (reference code over restated mlx primitives;
see that module's header and tests/golden/make_golden.py).

Reference map (under /root/reference/LTX_2_MLX/model/video_vae/):
  sinusoid [cos,sin] / MLP ......... simple_decoder.py:12-59
  Conv3dSimple (reflect HW, replicate T) ... simple_decoder.py:90-180
  ResBlock3d / pixel norm .......... simple_decoder.py:194-240, 339-342
  DepthToSpaceUpsample3d ........... simple_decoder.py:274-313
  decoder forward .................. simple_decoder.py:446-563
  decode_latent (chunk + cross-fade) simple_decoder.py:676-800
  unpatchify ....................... ops.py:70-131
  decode_tiled / trapezoid masks ... tiling.py:9-52, 154-249, 252-412
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

STRIDES = {"compress_all": (2, 2, 2), "compress_time": (2, 1, 1), "compress_space": (1, 2, 2)}


def sinusoid_vae(t: torch.Tensor, dim: int = 256) -> torch.Tensor:
    """simple_decoder.py:12-39 -- [cos, sin], exponent i/half."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    a = t.reshape(-1, 1).to(torch.float32) * freqs[None]
    return torch.cat([torch.cos(a), torch.sin(a)], dim=-1)


def _mlp(w, prefix, x):
    h = x @ w[prefix + ".linear_1.weight"].to(x.dtype).T + w[prefix + ".linear_1.bias"].to(x.dtype)
    h = F.silu(h)
    return h @ w[prefix + ".linear_2.weight"].to(x.dtype).T + w[prefix + ".linear_2.bias"].to(x.dtype)


def conv3d(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, causal: bool = False) -> torch.Tensor:
    """Conv3dSimple: reflect-pad H,W by 1; pad T by first/last-frame replication
    (causal: 2 before; non-causal: 1 before + 1 after); 3x3x3 cross-correlation; bias."""
    x = F.pad(x, (1, 1, 1, 1, 0, 0), mode="reflect")
    if causal:
        x = torch.cat([x[:, :, :1].expand(-1, -1, 2, -1, -1), x], dim=2)
    else:
        x = torch.cat([x[:, :, :1], x, x[:, :, -1:]], dim=2)
    return F.conv3d(x, weight.to(x.dtype), bias.to(x.dtype))


def pixel_norm(x, eps=1e-6):
    return x * torch.rsqrt(torch.mean(x * x, dim=1, keepdim=True) + eps)


def depth_to_space(x, c_out, stride):
    b, c, t, h, w = x.shape
    ft, fh, fw = stride
    x = x.reshape(b, c_out, ft, fh, fw, t, h, w).permute(0, 1, 5, 2, 6, 3, 7, 4)
    return x.reshape(b, c_out, t * ft, h * fh, w * fw)


def unpatchify(x, r=4):
    """ops.py:108-125 -- channel packing (c, p_t=1, r_w, r_h): width factor BEFORE height."""
    b, cp, f, h, w = x.shape
    c = cp // (r * r)
    x = x.reshape(b, c, 1, r, r, f, h, w).permute(0, 1, 5, 2, 6, 4, 7, 3)
    return x.reshape(b, c, f, h * r, w * r)


def res_block(w, prefix, x, time_emb, causal):
    c = x.shape[1]
    tab = w[prefix + ".scale_shift_table"].to(x.dtype)
    if time_emb is not None:
        ss = tab[None] + time_emb.reshape(-1, 4, c)
    else:
        ss = tab[None]
    sh1, sc1, sh2, sc2 = (ss[:, j, :, None, None, None] for j in range(4))
    h = F.silu(pixel_norm(x) * (1 + sc1) + sh1)
    h = conv3d(h, w[prefix + ".conv1.conv.weight"], w[prefix + ".conv1.conv.bias"], causal)
    h = F.silu(pixel_norm(h) * (1 + sc2) + sh2)
    h = conv3d(h, w[prefix + ".conv2.conv.weight"], w[prefix + ".conv2.conv.bias"], causal)
    return h + x


def upsample(w, prefix, x, stride, multiplier, residual, causal):
    ft, fh, fw = stride
    sp = ft * fh * fw
    c_in = x.shape[1]
    res = None
    if residual:
        res = depth_to_space(x, c_in // sp, stride)
        if ft > 1:
            res = res[:, :, 1:]
        res = res.repeat(1, sp // multiplier, 1, 1, 1)
    y = conv3d(x, w[prefix + ".conv.conv.weight"], w[prefix + ".conv.conv.bias"], causal)
    y = depth_to_space(y, c_in // multiplier, stride)
    if ft > 1:
        y = y[:, :, 1:]
    return y if res is None else y + res


def stages(decoder_blocks, base_channels) -> List[Tuple[str, dict, int]]:
    c = base_channels * 8
    out = []
    for name, params in reversed(decoder_blocks):
        p = {"num_layers": params} if isinstance(params, int) else dict(params)
        if name == "res_x":
            out.append(("res", p, c))
        else:
            p["stride"] = STRIDES[name]
            p.setdefault("multiplier", 1)
            p.setdefault("residual", False)
            out.append(("up", p, c))
            c //= p["multiplier"]
    return out


def vae_decode(w: Dict[str, torch.Tensor], latent: torch.Tensor, *, decoder_blocks, base_channels: int = 128,
               timestep: Optional[float] = 0.05, timestep_conditioning: bool = True,
               decode_noise_scale: float = 0.0, noise: Optional[torch.Tensor] = None,
               causal: bool = False, dtype=torch.float32) -> torch.Tensor:
    """SimpleVideoDecoder.__call__ (simple_decoder.py:446-563). latent (B,128,T,H,W) ->
    (B,3,8(T-1)+1,32H,32W) fp32.  `noise` (same shape as latent, N(0,1)) is consumed only when
    decode_noise_scale != 0 (the reference draws it from mx.random, :497)."""
    x = latent.to(dtype)
    B = x.shape[0]
    st = None
    if timestep_conditioning and timestep is not None:
        mult = float(w.get("vae.decoder.timestep_scale_multiplier", torch.tensor(1000.0)))
        st = torch.full((B,), float(timestep), dtype=torch.float32) * mult
    x = x * w["vae.per_channel_statistics.std-of-means"].to(dtype)[None, :, None, None, None]
    x = x + w["vae.per_channel_statistics.mean-of-means"].to(dtype)[None, :, None, None, None]
    if timestep_conditioning and timestep is not None and decode_noise_scale != 0.0:
        assert noise is not None
        x = noise.to(dtype) * decode_noise_scale + (1.0 - decode_noise_scale) * x
    x = conv3d(x, w["vae.decoder.conv_in.conv.weight"], w["vae.decoder.conv_in.conv.bias"], causal)
    for idx, (kind, p, c) in enumerate(stages(decoder_blocks, base_channels)):
        U = f"vae.decoder.up_blocks.{idx}"
        if kind == "res":
            te = None
            tk = U + ".time_embedder.timestep_embedder"
            if st is not None and (tk + ".linear_1.weight") in w:
                te = _mlp(w, tk, sinusoid_vae(st).to(dtype))
            for j in range(p["num_layers"]):
                x = res_block(w, f"{U}.res_blocks.{j}", x, te, causal)
        else:
            x = upsample(w, U, x, p["stride"], p["multiplier"], p["residual"], causal)
    x = pixel_norm(x)
    cf = x.shape[1]
    tab = w["vae.decoder.last_scale_shift_table"].to(dtype)
    lk = "vae.decoder.last_time_embedder.timestep_embedder"
    if st is not None and (lk + ".linear_1.weight") in w:
        ss = tab[None] + _mlp(w, lk, sinusoid_vae(st).to(dtype)).reshape(B, 2, cf)
    else:
        ss = tab[None]
    shift, scale = ss[:, 0, :, None, None, None], 1 + ss[:, 1, :, None, None, None]
    x = F.silu(x * scale + shift)
    x = conv3d(x, w["vae.decoder.conv_out.conv.weight"], w["vae.decoder.conv_out.conv.bias"], causal)
    return unpatchify(x, 4).to(torch.float32)


def chunk_plan(T: int, chunk: int = 7, overlap: int = 2) -> List[Tuple[int, int]]:
    """Temporal chunk schedule of decode_latent (simple_decoder.py:728-747)."""
    out = []
    stride = chunk - overlap
    t = 0
    while t < T:
        end = min(t + chunk, T)
        if end - t < overlap + 1 and t > 0:
            t = max(0, end - chunk)
            end = min(t + chunk, T)
        out.append((t, end))
        if end >= T:
            break
        t += stride
    return out


def _pix_t(lt: int) -> int:
    for _ in range(3):
        lt = lt * 2 - 1
    return lt


def blend_chunks(chunks: List[torch.Tensor], T_latent: int, overlap: int = 2) -> torch.Tensor:
    """Linear cross-fade stitching of decoded chunks (simple_decoder.py:749-790)."""
    total = _pix_t(T_latent)
    if len(chunks) == 1:
        return chunks[0][:, :, :total]
    ref = _pix_t(overlap)
    video = chunks[0]
    for cur in chunks[1:]:
        ov = min(ref, cur.shape[2], video.shape[2])
        if ov <= 1:
            video = torch.cat([video, cur], dim=2)
            continue
        ramp = torch.linspace(0.0, 1.0, ov).reshape(1, 1, ov, 1, 1)
        blended = video[:, :, -ov:] * (1.0 - ramp) + cur[:, :, :ov] * ramp
        video = torch.cat([video[:, :, :-ov], blended, cur[:, :, ov:]], dim=2)
    return video[:, :, :total]


def to_uint8_frames(video: torch.Tensor) -> torch.Tensor:
    """simple_decoder.py:793-798: clip((v+1)/2,0,1)*255 -> uint8 (truncation), (T,H,W,3)."""
    v = (torch.clamp((video + 1) / 2, 0, 1) * 255).to(torch.uint8)
    return v[0].permute(1, 2, 3, 0).contiguous()


def decode_latent(w, latent: torch.Tensor, *, decoder_blocks, base_channels=128, timestep=0.05,
                  timestep_conditioning=True, decode_noise_scale=0.0, chunk=7, overlap=2,
                  dtype=torch.float32) -> torch.Tensor:
    """decode_latent (simple_decoder.py:676-800) with decode_noise_scale fixed (0 for parity)."""
    if latent.ndim == 4:
        latent = latent[None]
    T = latent.shape[2]
    kw = dict(decoder_blocks=decoder_blocks, base_channels=base_channels, timestep=timestep,
              timestep_conditioning=timestep_conditioning, decode_noise_scale=decode_noise_scale, dtype=dtype)
    if T <= chunk:
        video = vae_decode(w, latent, **kw)
    else:
        parts = [vae_decode(w, latent[:, :, a:b], **kw) for a, b in chunk_plan(T, chunk, overlap)]
        video = blend_chunks(parts, T, overlap)
    return to_uint8_frames(video)


def trapezoid_mask_1d(length: int, ramp_left: int, ramp_right: int, left_starts_from_0: bool = False) -> torch.Tensor:
    """compute_trapezoidal_mask_1d (tiling.py:9-52): linear ramps that exclude the 0 and 1
    end points (fade-in keeps the 0 only when left_starts_from_0)."""
    ramp_left = max(0, min(ramp_left, length))
    ramp_right = max(0, min(ramp_right, length))
    m = torch.ones(length, dtype=torch.float32)
    if ramp_left > 0:
        n = ramp_left + 1 if left_starts_from_0 else ramp_left + 2
        fade = torch.linspace(0.0, 1.0, n)[:-1]
        if not left_starts_from_0:
            fade = fade[1:]
        m = torch.cat([fade, m[ramp_left:]])
    if ramp_right > 0:
        m = torch.cat([m[:-ramp_right], torch.linspace(1.0, 0.0, ramp_right + 2)[1:-1]])
    return m.clamp(0, 1)


def tiles_1d(length: int, tile: int, overlap: int):
    """gen_tiles_1d of generate_tile_specs (tiling.py:203-226)."""
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


def decode_tiled(decode_fn, latent: torch.Tensor, *, tile_px: Optional[int], overlap_px: int = 0,
                 tile_frames: Optional[int] = None, overlap_frames: int = 0) -> torch.Tensor:
    """decode_tiled (tiling.py:252-412): trapezoid-weighted blend of independently decoded tiles.
    decode_fn(latent_tile) -> (B,3,T',H',W')."""
    b, _, t, h, w = latent.shape
    th, oh = (tile_px // 32, overlap_px // 32) if tile_px else (max(h, w), 0)
    tt, ot = (tile_frames // 8, overlap_frames // 8) if tile_frames else (t, 0)
    To, Ho, Wo = (t - 1) * 8 + 1, h * 32, w * 32
    out = torch.zeros(b, 3, To, Ho, Wo)
    ws = torch.zeros(1, 1, To, Ho, Wo)
    h_tiles = tiles_1d(h, th, oh) if tile_px else [(0, h, 0, 0)]
    w_tiles = tiles_1d(w, th, oh) if tile_px else [(0, w, 0, 0)]
    for (t0, t1, rtl, rtr) in tiles_1d(t, tt, ot):
        for (h0, h1, rhl, rhr) in h_tiles:
            for (w0, w1, rwl, rwr) in w_tiles:
                tile = decode_fn(latent[:, :, t0:t1, h0:h1, w0:w1])
                ot0 = t0 * 8 if t0 > 0 else 0
                ot1 = (t1 - 1) * 8 + 1 if t1 > 1 else 1
                nt = min(tile.shape[2], ot1 - ot0)
                nh = min(tile.shape[3], (h1 - h0) * 32)
                nw = min(tile.shape[4], (w1 - w0) * 32)
                mt = trapezoid_mask_1d(nt, min(rtl * 8, nt), min(rtr * 8, nt), left_starts_from_0=(ot0 == 0))
                mh = trapezoid_mask_1d(nh, min(rhl * 32, nh), min(rhr * 32, nh))
                mw = trapezoid_mask_1d(nw, min(rwl * 32, nw), min(rwr * 32, nw))
                m = mt[None, None, :, None, None] * mh[None, None, None, :, None] * mw[None, None, None, None, :]
                out[:, :, ot0:ot0 + nt, h0 * 32:h0 * 32 + nh, w0 * 32:w0 * 32 + nw] += tile[:, :, :nt, :nh, :nw] * m
                ws[:, :, ot0:ot0 + nt, h0 * 32:h0 * 32 + nh, w0 * 32:w0 * 32 + nw] += m
    return out / torch.clamp(ws, min=1e-8)
