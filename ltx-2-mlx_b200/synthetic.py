"""Seeded synthetic checkpoints in the reference's *checkpoint* key names.

There is no network, so neither the 43 GB LTX-2 checkpoint nor the reference's
PyTorch parity fixtures exist here.  Every parity test and
bench run therefore uses weights drawn by this module, written under the same
names the reference loaders consume:

  * DiT: ``model.diffusion_model.<pytorch name>`` -- the names that
    ``LTX_2_MLX/loader/weight_converter.py:318-446`` reads and renames
    (``to_out.0 -> to_out``, ``ff.net.0.proj -> ff.project_in.proj``,
    ``ff.net.2 -> ff.project_out``; weight_converter.py:300-313).
  * VAE: ``vae.decoder.*`` / ``vae.per_channel_statistics.*`` -- the names that
    ``LTX_2_MLX/model/video_vae/simple_decoder.py:566-673`` reads.

Distributions follow SURVEY.md 8(d): Linear weight/bias ~ U(-1/sqrt(in), 1/sqrt(in))
(MLX ``nn.Linear`` default), q/k-norm weight = 1 + 0.1 N(0,1), every
scale_shift_table ~ 0.1 N(0,1) (the reference's zero init would hide modulation
bugs), VAE conv weight ~ N(0, 1/(27 C_in)).  Each tensor has its own generator
seeded from crc32(key), so generation order does not matter.
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Optional, Tuple

import torch

DIT_PREFIX = "model.diffusion_model."


@dataclass
class DitConfig:
    """Architecture of the DiT (defaults = LTX-2 19B, model.py:436-461)."""

    num_attention_heads: int = 32
    attention_head_dim: int = 128
    in_channels: int = 128
    out_channels: int = 128
    num_layers: int = 48
    cross_attention_dim: int = 4096
    caption_channels: Optional[int] = 3840
    cross_attention_adaln: bool = False      # V2 / LTX-2.3
    apply_gated_attention: bool = False      # V2 / LTX-2.3
    audio: bool = False                      # LTXModelType.AudioVideo
    audio_heads: int = 32                    # model.py:428
    audio_head_dim: int = 64                 # model.py:429
    audio_in_channels: int = 128
    audio_out_channels: int = 128

    @property
    def dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim

    @property
    def audio_dim(self) -> int:
        return self.audio_heads * self.audio_head_dim

    @property
    def adaln_params(self) -> int:
        return 9 if self.cross_attention_adaln else 6


# default V2.0 decoder stack, simple_decoder.py:353-361
DEFAULT_DECODER_BLOCKS = [
    ["res_x", {"num_layers": 5}],
    ["compress_all", {"multiplier": 2, "residual": True}],
    ["res_x", {"num_layers": 5}],
    ["compress_all", {"multiplier": 2, "residual": True}],
    ["res_x", {"num_layers": 5}],
    ["compress_all", {"multiplier": 2, "residual": True}],
    ["res_x", {"num_layers": 5}],
]

STRIDES = {"compress_all": (2, 2, 2), "compress_time": (2, 1, 1), "compress_space": (1, 2, 2)}


@dataclass
class VaeConfig:
    decoder_blocks: List = field(default_factory=lambda: [list(b) for b in DEFAULT_DECODER_BLOCKS])
    base_channels: int = 128
    timestep_conditioning: bool = True
    latent_channels: int = 128

    def stages(self) -> List[Tuple[str, dict, int]]:
        """(kind, params, in_channels) in execution order (simple_decoder.py:403-427)."""
        c = self.base_channels * 8
        out = []
        for name, params in reversed(self.decoder_blocks):
            p = {"num_layers": params} if isinstance(params, int) else dict(params)
            if name == "res_x":
                out.append(("res", p, c))
            elif name in STRIDES:
                p = dict(p)
                p["stride"] = STRIDES[name]
                p.setdefault("multiplier", 1)
                p.setdefault("residual", False)
                out.append(("up", p, c))
                c = c // p["multiplier"]
            else:
                raise ValueError(f"Unknown decoder block: {name}")
        return out

    @property
    def final_channels(self) -> int:
        c = self.base_channels * 8
        for name, params in reversed(self.decoder_blocks):
            if name in STRIDES:
                p = {"multiplier": 1} if isinstance(params, int) else params
                c = c // p.get("multiplier", 1)
        return c


def _gen(key: str, seed: int, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _uniform(key, shape, bound, seed, device, dtype):
    g = _gen(key, seed, device)
    t = torch.empty(shape, device=device, dtype=torch.float32)
    t.uniform_(-bound, bound, generator=g)
    return t.to(dtype)


def _normal(key, shape, std, mean, seed, device, dtype):
    g = _gen(key, seed, device)
    t = torch.empty(shape, device=device, dtype=torch.float32)
    t.normal_(mean, std, generator=g)
    return t.to(dtype)


def _linear(prefix, out_f, in_f, seed, device, dtype, bias=True):
    b = 1.0 / math.sqrt(in_f)
    yield prefix + ".weight", _uniform(prefix + ".weight", (out_f, in_f), b, seed, device, dtype)
    if bias:
        yield prefix + ".bias", _uniform(prefix + ".bias", (out_f,), b, seed, device, dtype)


def _adaln_single(prefix, dim, n_emb, seed, device, dtype):
    yield from _linear(prefix + ".emb.timestep_embedder.linear_1", dim, 256, seed, device, dtype)
    yield from _linear(prefix + ".emb.timestep_embedder.linear_2", dim, dim, seed, device, dtype)
    yield from _linear(prefix + ".linear", n_emb * dim, dim, seed, device, dtype)


def _attention(prefix, query_dim, context_dim, inner, heads, gated, seed, device, dtype):
    yield from _linear(prefix + ".to_q", inner, query_dim, seed, device, dtype)
    yield from _linear(prefix + ".to_k", inner, context_dim, seed, device, dtype)
    yield from _linear(prefix + ".to_v", inner, context_dim, seed, device, dtype)
    yield from _linear(prefix + ".to_out.0", query_dim, inner, seed, device, dtype)
    yield prefix + ".q_norm.weight", _normal(prefix + ".q_norm.weight", (inner,), 0.1, 1.0, seed, device, dtype)
    yield prefix + ".k_norm.weight", _normal(prefix + ".k_norm.weight", (inner,), 0.1, 1.0, seed, device, dtype)
    if gated:
        yield from _linear(prefix + ".to_gate_logits", heads, query_dim, seed, device, dtype)


def _table(key, rows, dim, seed, device):
    return key, _normal(key, (rows, dim), 0.1, 0.0, seed, device, torch.float32)


def iter_dit_weights(cfg: DitConfig, seed: int = 0, device="cpu", dtype=torch.float32
                     ) -> Iterator[Tuple[str, torch.Tensor]]:
    """Yield (checkpoint_key, tensor) for the DiT, one tensor at a time (streamable)."""
    P = DIT_PREFIX
    D, Da = cfg.dim, cfg.audio_dim
    n = cfg.adaln_params
    v2, gated = cfg.cross_attention_adaln, cfg.apply_gated_attention

    yield from _linear(P + "patchify_proj", D, cfg.in_channels, seed, device, dtype)
    yield from _adaln_single(P + "adaln_single", D, n, seed, device, dtype)
    if v2:
        yield from _adaln_single(P + "prompt_adaln_single", D, 2, seed, device, dtype)
    if cfg.caption_channels is not None:
        yield from _linear(P + "caption_projection.linear_1", D, cfg.caption_channels, seed, device, dtype)
        yield from _linear(P + "caption_projection.linear_2", D, D, seed, device, dtype)
    yield _table(P + "scale_shift_table", 2, D, seed, device)
    yield from _linear(P + "proj_out", cfg.out_channels, D, seed, device, dtype)

    if cfg.audio:
        yield from _linear(P + "audio_patchify_proj", Da, cfg.audio_in_channels, seed, device, dtype)
        yield from _adaln_single(P + "audio_adaln_single", Da, n, seed, device, dtype)
        if v2:
            yield from _adaln_single(P + "audio_prompt_adaln_single", Da, 2, seed, device, dtype)
        if cfg.caption_channels is not None:
            yield from _linear(P + "audio_caption_projection.linear_1", Da, cfg.caption_channels, seed, device, dtype)
            yield from _linear(P + "audio_caption_projection.linear_2", Da, Da, seed, device, dtype)
        yield _table(P + "audio_scale_shift_table", 2, Da, seed, device)
        yield from _linear(P + "audio_proj_out", cfg.audio_out_channels, Da, seed, device, dtype)
        yield from _adaln_single(P + "av_ca_video_scale_shift_adaln_single", D, 4, seed, device, dtype)
        yield from _adaln_single(P + "av_ca_a2v_gate_adaln_single", D, 1, seed, device, dtype)
        yield from _adaln_single(P + "av_ca_audio_scale_shift_adaln_single", Da, 4, seed, device, dtype)
        yield from _adaln_single(P + "av_ca_v2a_gate_adaln_single", Da, 1, seed, device, dtype)

    for i in range(cfg.num_layers):
        B = f"{P}transformer_blocks.{i}"
        H = cfg.num_attention_heads
        yield from _attention(B + ".attn1", D, D, D, H, gated, seed, device, dtype)
        yield from _attention(B + ".attn2", D, cfg.cross_attention_dim, D, H, gated, seed, device, dtype)
        yield from _linear(B + ".ff.net.0.proj", 4 * D, D, seed, device, dtype)
        yield from _linear(B + ".ff.net.2", D, 4 * D, seed, device, dtype)
        yield _table(B + ".scale_shift_table", n, D, seed, device)
        if v2:
            yield _table(B + ".prompt_scale_shift_table", 2, D, seed, device)
        if cfg.audio:
            Ha = cfg.audio_heads
            yield from _attention(B + ".audio_attn1", Da, Da, Da, Ha, gated, seed, device, dtype)
            yield from _attention(B + ".audio_attn2", Da, Da, Da, Ha, gated, seed, device, dtype)
            yield from _linear(B + ".audio_ff.net.0.proj", 4 * Da, Da, seed, device, dtype)
            yield from _linear(B + ".audio_ff.net.2", Da, 4 * Da, seed, device, dtype)
            yield _table(B + ".audio_scale_shift_table", n, Da, seed, device)
            if v2:
                yield _table(B + ".audio_prompt_scale_shift_table", 2, Da, seed, device)
            # a2v: Q video, K/V audio; v2a: Q audio, K/V video (transformer.py:339-361)
            yield from _attention(B + ".audio_to_video_attn", D, Da, Da, Ha, gated, seed, device, dtype)
            yield from _attention(B + ".video_to_audio_attn", Da, D, Da, Ha, gated, seed, device, dtype)
            yield _table(B + ".scale_shift_table_a2v_ca_audio", 5, Da, seed, device)
            yield _table(B + ".scale_shift_table_a2v_ca_video", 5, D, seed, device)


def _conv3d(prefix, c_out, c_in, seed, device, dtype):
    std = 1.0 / math.sqrt(27 * c_in)
    yield prefix + ".weight", _normal(prefix + ".weight", (c_out, c_in, 3, 3, 3), std, 0.0, seed, device, dtype)
    yield prefix + ".bias", _normal(prefix + ".bias", (c_out,), 0.02, 0.0, seed, device, dtype)


def iter_vae_weights(cfg: VaeConfig, seed: int = 0, device="cpu", dtype=torch.float32
                     ) -> Iterator[Tuple[str, torch.Tensor]]:
    """Yield (checkpoint_key, tensor) for the video-VAE decoder."""
    L = cfg.latent_channels
    yield "vae.per_channel_statistics.mean-of-means", _normal("vae.mom", (L,), 0.1, 0.0, seed, device, torch.float32)
    yield "vae.per_channel_statistics.std-of-means", _normal("vae.som", (L,), 0.05, 1.0, seed, device, torch.float32)
    c0 = cfg.base_channels * 8
    yield from _conv3d("vae.decoder.conv_in.conv", c0, L, seed, device, dtype)
    for idx, (kind, p, c) in enumerate(cfg.stages()):
        U = f"vae.decoder.up_blocks.{idx}"
        if kind == "res":
            for j in range(p["num_layers"]):
                yield from _conv3d(f"{U}.res_blocks.{j}.conv1.conv", c, c, seed, device, dtype)
                yield from _conv3d(f"{U}.res_blocks.{j}.conv2.conv", c, c, seed, device, dtype)
                yield _table(f"{U}.res_blocks.{j}.scale_shift_table", 4, c, seed, device)
            if cfg.timestep_conditioning:
                T = f"{U}.time_embedder.timestep_embedder"
                yield from _linear(T + ".linear_1", 4 * c, 256, seed, device, dtype)
                yield from _linear(T + ".linear_2", 4 * c, 4 * c, seed, device, dtype)
        else:
            s = p["stride"]
            c_conv = s[0] * s[1] * s[2] * c // p["multiplier"]
            yield from _conv3d(f"{U}.conv.conv", c_conv, c, seed, device, dtype)
    cf = cfg.final_channels
    yield from _conv3d("vae.decoder.conv_out.conv", 48, cf, seed, device, dtype)
    yield _table("vae.decoder.last_scale_shift_table", 2, cf, seed, device)
    if cfg.timestep_conditioning:
        yield "vae.decoder.timestep_scale_multiplier", torch.tensor(1000.0, device=device)
        T = "vae.decoder.last_time_embedder.timestep_embedder"
        yield from _linear(T + ".linear_1", 256, 256, seed, device, dtype)
        yield from _linear(T + ".linear_2", 2 * cf, 256, seed, device, dtype)


def iter_vae_encoder_weights(seed: int = 0, device="cpu", dtype=torch.float32, channel_div: int = 1
                             ) -> Iterator[Tuple[str, torch.Tensor]]:
    """Yield (checkpoint_key, tensor) for the video-VAE ENCODER under the key names `load_vae_encoder_weights` reads
    (simple_encoder.py:408-528).  The reference encoder's widths are fixed (128 .. 1024); `channel_div` > 1 only serves
    block-level tests."""
    L = 128
    yield "vae.per_channel_statistics.mean-of-means", _normal("vae.mom", (L,), 0.1, 0.0, seed, device, torch.float32)
    yield "vae.per_channel_statistics.std-of-means", _normal("vae.som", (L,), 0.05, 1.0, seed, device, torch.float32)
    d = channel_div
    yield from _conv3d("vae.encoder.conv_in.conv", 128 // d, 48, seed, device, dtype)
    blocks = [("res", 128, 4, None), ("down", 128, 256, (1, 2, 2)), ("res", 256, 6, None), ("down", 256, 512, (2, 1, 1)),
              ("res", 512, 6, None), ("down", 512, 1024, (2, 2, 2)), ("res", 1024, 2, None),
              ("down", 1024, 1024, (2, 2, 2)), ("res", 1024, 2, None)]
    for idx, (kind, c_in, n_or_cout, stride) in enumerate(blocks):
        P = f"vae.encoder.down_blocks.{idx}"
        if kind == "res":
            for j in range(n_or_cout):
                yield from _conv3d(f"{P}.res_blocks.{j}.conv1.conv", c_in // d, c_in // d, seed, device, dtype)
                yield from _conv3d(f"{P}.res_blocks.{j}.conv2.conv", c_in // d, c_in // d, seed, device, dtype)
        else:
            sp = stride[0] * stride[1] * stride[2]
            yield from _conv3d(f"{P}.conv.conv", n_or_cout // d // sp, c_in // d, seed, device, dtype)
    yield from _conv3d("vae.encoder.conv_out.conv", 129, 1024 // d, seed, device, dtype)


def iter_upscaler_weights(seed: int = 0, device="cpu", dtype=torch.float32, in_channels: int = 128,
                          mid_channels: int = 1024, blocks: int = 4) -> Iterator[Tuple[str, torch.Tensor]]:
    """Yield (checkpoint_key, tensor) for the 2x latent spatial upscaler under the key names
    `load_spatial_upscaler_weights` reads (model/upscaler/spatial.py:414-538)."""
    def conv(prefix, c_out, c_in):
        w = _normal(prefix + ".weight", (c_out, c_in, 3, 3, 3), 1.0 / math.sqrt(27 * c_in), 0.0, seed, device, dtype)
        return [(prefix + ".weight", w), (prefix + ".bias", _normal(prefix + ".bias", (c_out,), 0.02, 0.0, seed, device, dtype))]

    def norm(prefix, c):
        return [(prefix + ".weight", _normal(prefix + ".weight", (c,), 0.1, 1.0, seed, device, torch.float32)),
                (prefix + ".bias", _normal(prefix + ".bias", (c,), 0.05, 0.0, seed, device, torch.float32))]

    yield from conv("initial_conv", mid_channels, in_channels)
    yield from norm("initial_norm", mid_channels)
    for stage in ("res_blocks", "post_upsample_res_blocks"):
        for i in range(blocks):
            yield from conv(f"{stage}.{i}.conv1", mid_channels, mid_channels)
            yield from norm(f"{stage}.{i}.norm1", mid_channels)
            yield from conv(f"{stage}.{i}.conv2", mid_channels, mid_channels)
            yield from norm(f"{stage}.{i}.norm2", mid_channels)
    yield "upsampler.conv.weight", _normal("upsampler.conv.weight", (4 * mid_channels, mid_channels, 3, 3),
                                           1.0 / math.sqrt(9 * mid_channels), 0.0, seed, device, dtype)
    yield "upsampler.conv.bias", _normal("upsampler.conv.bias", (4 * mid_channels,), 0.02, 0.0, seed, device, dtype)
    yield from conv("final_conv", in_channels, mid_channels)


def iter_connector_weights(heads: int = 30, head_dim: int = 128, layers: int = 2, registers: int = 128, gated: bool = True,
                           seed: int = 0, device="cpu", dtype=torch.float32) -> Iterator[Tuple[str, torch.Tensor]]:
    """Yield (module_key, tensor) for the text-embeddings connector (model/text_encoder/connector.py:104-176) under the
    reference's module attribute names, i.e. the checkpoint keys `model.diffusion_model.video_embeddings_connector.*`
    after loader/weight_converter.py:146-162,300-313."""
    D = heads * head_dim
    for i in range(layers):
        P = f"transformer_1d_blocks.{i}"
        for k, t in _attention(P + ".attn1", D, D, D, heads, gated, seed, device, dtype):
            yield k.replace(".to_out.0.", ".to_out."), t
        yield from _linear(P + ".ff.project_in.proj", 4 * D, D, seed, device, dtype)
        yield from _linear(P + ".ff.project_out", D, 4 * D, seed, device, dtype)
    if registers:
        yield "learnable_registers", _uniform("learnable_registers", (registers, D), 1.0, seed, device, torch.float32)


def dit_weights(cfg: DitConfig, seed: int = 0, device="cpu", dtype=torch.float32) -> Dict[str, torch.Tensor]:
    return dict(iter_dit_weights(cfg, seed, device, dtype))


def vae_weights(cfg: VaeConfig, seed: int = 0, device="cpu", dtype=torch.float32) -> Dict[str, torch.Tensor]:
    return dict(iter_vae_weights(cfg, seed, device, dtype))


# ---------------------------------------------------------------------------------
# synthetic activations
# ---------------------------------------------------------------------------------

def video_positions(batch: int, frames: int, height: int, width: int, fps: Optional[float] = 24.0,
                    device="cpu") -> torch.Tensor:
    """[start, end) pixel/second bounds per token, shape (B, 3, N, 2) fp32.

    Restates VideoLatentPatchifier.get_patch_grid_bounds (patchifiers.py:147-199) +
    get_pixel_coords(causal_fix=True) (patchifiers.py:202-240) with scale factors
    (8, 32, 32) and the temporal axis divided by fps (conditioning/tools.py:69-78).
    Pass fps=None for the unscaled pixel coordinates test_parity.py:280-284 feeds.
    """
    f = torch.arange(frames, dtype=torch.float32)
    h = torch.arange(height, dtype=torch.float32)
    w = torch.arange(width, dtype=torch.float32)
    gf, gh, gw = torch.meshgrid(f, h, w, indexing="ij")
    starts = torch.stack([gf, gh, gw], 0).reshape(3, -1)
    ends = starts + 1.0
    coords = torch.stack([starts, ends], -1)                       # (3, N, 2)
    scale = torch.tensor([8.0, 32.0, 32.0]).reshape(3, 1, 1)
    px = coords * scale
    px[0] = torch.clamp(px[0] + 1.0 - 8.0, min=0.0)
    if fps is not None:
        px[0] = px[0] / float(fps)
    return px[None].expand(batch, 3, px.shape[1], 2).contiguous().to(device)


def audio_positions(batch: int, n_tokens: int, tokens_per_second: float = 25.0, device="cpu") -> torch.Tensor:
    """1-D temporal [start, end) bounds in seconds for audio tokens, shape (B, 1, N_a, 2)."""
    t = torch.arange(n_tokens, dtype=torch.float32) / tokens_per_second
    pos = torch.stack([t, t + 1.0 / tokens_per_second], -1)[None, None]
    return pos.expand(batch, 1, n_tokens, 2).contiguous().to(device)


def latents(shape, seed: int = 42, device="cpu", std: float = 1.0) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return (torch.randn(shape, generator=g, dtype=torch.float32) * std).to(device)
