"""Host-side mirror of the reference's video-VAE decode interface, backed by the sm_100a engine.

Same names, argument meaning and defaults as ``LTX_2_MLX/model/video_vae/simple_decoder.py``:
``SimpleVideoDecoder`` (:364-563; ``__call__(latent, timestep=0.05, show_progress=True, causal=False)``),
``load_vae_decoder_weights(decoder, path)`` (:566-673) and
``decode_latent(latent, decoder, timestep=0.05, key=None, temporal_chunk_size=7, temporal_overlap=2)`` (:676-800).
Attributes pipelines touch from outside are kept: ``decode_noise_scale`` (tests/test_parity.py:359),
``std_of_means`` / ``mean_of_means`` (pipelines/one_stage.py:980-981).

All arithmetic (convs, norms, depth-to-space, unpatchify, chunk cross-fade, uint8 conversion) runs in the C-ABI
library; this module only schedules chunks and moves buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch

from ._lib import LtxVaeConfig, LtxVaeStage, check, dtype_code, lib, ptr, stream_ptr
from .synthetic import DEFAULT_DECODER_BLOCKS, STRIDES
from .transformer import to_device

_STRIDE_MAP = STRIDES


class SimpleVideoDecoder:
    """Config-driven VAE decoder (V2.0 and V2.3 stacks) -- drop-in for simple_decoder.py:364-563."""

    def __init__(self, decoder_blocks: Optional[List] = None, base_channels: int = 128,
                 timestep_conditioning: bool = True, compute_dtype: Any = None, device="cuda"):
        self.device = torch.device(device)
        self.compute_dtype = compute_dtype
        self.timestep_conditioning = timestep_conditioning
        self.decode_noise_scale = 0.025                       # simple_decoder.py:391
        self.decoder_blocks = [list(b) for b in (decoder_blocks or DEFAULT_DECODER_BLOCKS)]
        self.base_channels = base_channels
        self.mean_of_means = torch.zeros(128)
        self.std_of_means = torch.zeros(128)
        self._noise_generator: Optional[torch.Generator] = None
        self._shards = None            # (rank, world, group) once enable_temporal_shards() ran

        cfg = LtxVaeConfig()
        cfg.base_channels, cfg.latent_channels = base_channels, 128
        cfg.timestep_conditioning = int(timestep_conditioning)
        self.block_types: List[str] = []
        feature = base_channels * 8
        n = 0
        for name, params in reversed(self.decoder_blocks):
            p = {"num_layers": params} if isinstance(params, int) else dict(params)
            st = LtxVaeStage()
            if name == "res_x":
                st.kind, st.num_layers = 0, int(p["num_layers"])
                st.stride_t = st.stride_h = st.stride_w = st.multiplier = 1
                self.block_types.append("res")
            elif name in _STRIDE_MAP:
                st.kind = 1
                st.stride_t, st.stride_h, st.stride_w = _STRIDE_MAP[name]
                st.multiplier = int(p.get("multiplier", 1))
                st.residual = int(bool(p.get("residual", False)))
                feature //= st.multiplier
                self.block_types.append("upsample")
            else:
                raise ValueError(f"Unknown decoder block: {name}")           # simple_decoder.py:427
            cfg.stages[n] = st
            n += 1
        cfg.num_stages = n
        self.final_channels = feature
        self._cfg = cfg
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().ltx2_vae_create(C.byref(cfg), C.byref(self._h)), "ltx2_vae_create")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                lib().ltx2_vae_destroy(h)
            except Exception:
                pass
            self._h = C.c_void_p()

    # ---- weights ------------------------------------------------------------------------------
    def load_weights(self, weights: Iterable[Tuple[str, Any]]) -> int:
        """(checkpoint_key, tensor) pairs with the names simple_decoder.py:592-671 reads; others are skipped."""
        n = 0
        with torch.cuda.device(self.device):
            for key, value in (weights.items() if isinstance(weights, dict) else weights):
                if not key.startswith("vae."):
                    continue
                t = to_device(value, self.device)
                shape = (C.c_int64 * max(t.ndim, 1))(*(t.shape if t.ndim else (1,)))
                st = lib().ltx2_vae_set_weight(self._h, key.encode(), ptr(t), dtype_code(t), shape, max(t.ndim, 1),
                                               stream_ptr())
                if st == -2:       # LTX2_ERR_NOKEY: encoder / unused tensors
                    continue
                check(st, f"vae set_weight({key})")
                if key.endswith("mean-of-means"):
                    self.mean_of_means = t.float().cpu()
                elif key.endswith("std-of-means"):
                    self.std_of_means = t.float().cpu()
                n += 1
            torch.cuda.current_stream().synchronize()
        return n

    def missing_weights(self) -> List[str]:
        buf = C.create_string_buffer(8192)
        n = lib().ltx2_vae_missing_weights(self._h, buf, C.c_int64(len(buf)))
        return [s for s in buf.value.decode().split("\n") if s] if n else []

    def output_shape(self, latent_shape) -> Tuple[int, ...]:
        i = (C.c_int64 * 5)(*latent_shape)
        o = (C.c_int64 * 5)()
        check(lib().ltx2_vae_output_shape(self._h, i, o))
        return tuple(o)

    # ---- forward ------------------------------------------------------------------------------
    def __call__(self, latent, timestep: Optional[float] = 0.05, show_progress: bool = True, causal: bool = False,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """latent (B,128,T,H,W) -> video (B,3,8(T-1)+1,32H,32W) fp32 on the GPU (simple_decoder.py:446-563)."""
        with torch.cuda.device(self.device):
            x = to_device(latent, self.device)
            if x.ndim != 5:
                raise ValueError(f"latent must be (B, C, T, H, W); got {tuple(x.shape)}")
            oshape = self.output_shape(x.shape)
            if out is None:
                out = torch.empty(oshape, device=self.device, dtype=torch.float32)
            noise = None
            s = float(self.decode_noise_scale)
            if self.timestep_conditioning and timestep is not None and s != 0.0:
                # the reference draws mx.random.normal here (:497); the draw is host plumbing, the blend is in-kernel
                noise = torch.randn(x.shape, device=self.device, dtype=torch.float32, generator=self._noise_generator)
            shape = (C.c_int64 * 5)(*x.shape)
            check(lib().ltx2_vae_decode(self._h, ptr(x), dtype_code(x), shape,
                                        C.c_float(-1.0 if timestep is None else float(timestep)), C.c_float(s),
                                        ptr(noise), int(bool(causal)), ptr(out), stream_ptr()), "ltx2_vae_decode")
        return out


def shard_frames(decoder: SimpleVideoDecoder, latent_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """(first output frame, output frame count) of `rank` when a `latent_frames`-frame latent is decoded in temporal
    shards over `world` ranks (host logic of the C ABI, no GPU work)."""
    t0, tn = C.c_int64(), C.c_int64()
    check(lib().ltx2_vae_shard_frames(decoder._h, latent_frames, rank, world, C.byref(t0), C.byref(tn)),
          "ltx2_vae_shard_frames")
    return int(t0.value), int(tn.value)


def enable_temporal_shards(decoder: SimpleVideoDecoder, max_latent_shape, group=None) -> None:
    """Collective: set up the exchange region for decoding ONE clip over the ranks of `group` in temporal shards
    (include/ltx2_b200.h, ltx2_vae_cp_*).  `max_latent_shape` = the largest (B, 128, T, H, W) latent (or chunk) that will
    be decoded; every rank must pass the same latents and hold the same weights afterwards."""
    import torch.distributed as dist
    from .context_parallel import exchange_handles
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return
    buf = C.create_string_buffer(64)
    shp = (C.c_int64 * 5)(*[int(v) for v in max_latent_shape])
    with torch.cuda.device(decoder.device):
        check(lib().ltx2_vae_cp_init(decoder._h, rank, world, shp, buf), "ltx2_vae_cp_init")
        handles = exchange_handles(buf.raw, group)
        check(lib().ltx2_vae_cp_connect(decoder._h, handles), "ltx2_vae_cp_connect")
        torch.cuda.synchronize()
    dist.barrier(group=group)
    decoder._shards = (rank, world, group)
    # the decode-noise draws (simple_decoder.py:496-498) must be identical on every rank: same seed, same call sequence
    decoder._shard_noise = torch.Generator(device=decoder.device)
    decoder._shard_noise.manual_seed(0x5EED)


def disable_temporal_shards(decoder: SimpleVideoDecoder) -> None:
    """Collective teardown: close the imported mappings, barrier, free the own region."""
    import torch.distributed as dist
    if decoder._shards is None:
        return
    group = decoder._shards[2]
    with torch.cuda.device(decoder.device):
        check(lib().ltx2_vae_cp_shutdown(decoder._h, 0), "ltx2_vae_cp_shutdown")
        dist.barrier(group=group)
        check(lib().ltx2_vae_cp_shutdown(decoder._h, 1), "ltx2_vae_cp_shutdown")
    dist.barrier(group=group)
    decoder._shards = None


def decode_sharded(decoder: SimpleVideoDecoder, latent, timestep: Optional[float] = 0.05, dst: Optional[int] = None,
                   group_ranks: Optional[Tuple[int, int]] = None, slot: int = 0,
                   collect: bool = True) -> Optional[torch.Tensor]:
    """SimpleVideoDecoder.__call__ over the ranks enabled by enable_temporal_shards(): every rank of the rank group
    `group_ranks` = (first, size) (default: all ranks) passes the same latent, computes its frame range and stores it into
    clip slot `slot` on rank `dst` (or on every rank, dst=None) through peer memory -- no torch.distributed call on the
    data path.  collect=True finishes with collect_clip() (a barrier over ALL ranks: every rank must get there); with
    collect=False the caller does that itself, e.g. after two groups have decoded two chunks side by side.
    Bit-identical to the single-GPU decode (noise injection off, or the same noise on every rank)."""
    rank, world, group = decoder._shards
    first, size = group_ranks if group_ranks is not None else (0, world)
    with torch.cuda.device(decoder.device):
        x = to_device(latent, decoder.device)
        if x.ndim != 5:
            raise ValueError(f"latent must be (B, C, T, H, W); got {tuple(x.shape)}")
        noise = None
        s = float(decoder.decode_noise_scale)
        if decoder.timestep_conditioning and timestep is not None and s != 0.0:
            # every rank must blend the SAME noise into the latent (each also builds its neighbours' boundary frames of
            # the first conv input): enable_temporal_shards() seeded one generator identically on all ranks
            noise = torch.randn(x.shape, device=decoder.device, dtype=torch.float32, generator=decoder._shard_noise)
        if first <= rank < first + size:
            shape = (C.c_int64 * 5)(*x.shape)
            check(lib().ltx2_vae_decode_sharded(decoder._h, ptr(x), dtype_code(x), shape,
                                                -1.0 if timestep is None else float(timestep), s, ptr(noise), first, size,
                                                slot, -1 if dst is None else int(dst), stream_ptr()),
                  "ltx2_vae_decode_sharded")
        if not collect:
            return None
        return collect_clip(decoder, x.shape, dst, slot)


def collect_clip(decoder: SimpleVideoDecoder, latent_shape, dst: Optional[int] = None, slot: int = 0):
    """Collective over ALL ranks: wait until every rank group has stored its frames of clip slot `slot`, then return the
    clip (B, 3, T', 32H, 32W) on rank `dst` (None elsewhere) or on every rank (dst=None)."""
    rank, world, group = decoder._shards
    with torch.cuda.device(decoder.device):
        oshape = decoder.output_shape(tuple(latent_shape))
        receive = dst is None or rank == dst
        video = torch.empty(oshape, device=decoder.device, dtype=torch.float32) if receive else None
        check(lib().ltx2_vae_cp_collect(decoder._h, slot, (C.c_int64 * 5)(*oshape), -1 if dst is None else int(dst),
                                        ptr(video), stream_ptr()), "ltx2_vae_cp_collect")
        return video


def load_vae_decoder_weights(decoder: SimpleVideoDecoder, weights_path: str) -> None:
    """Drop-in for simple_decoder.load_vae_decoder_weights: safetensors file -> engine."""
    from safetensors import safe_open

    print(f"Loading VAE decoder weights from {weights_path}...")
    with safe_open(weights_path, framework="pt") as f:
        def gen():
            for k in f.keys():
                if k.startswith("vae.decoder.") or k.startswith("vae.per_channel_statistics."):
                    yield k, f.get_tensor(k)
        n = decoder.load_weights(gen())
    print(f"  Loaded {n} weight tensors")


def chunk_plan(T: int, chunk: int = 7, overlap: int = 2) -> List[Tuple[int, int]]:
    """Temporal chunk schedule of decode_latent (simple_decoder.py:728-747)."""
    out, stride, t = [], chunk - overlap, 0
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


def unit_owner(index: int, world: int) -> int:
    """Round-robin owner of the index-th independent decode unit (temporal chunk or tile), SURVEY.md 8(e)."""
    return index % world


def _exchange_unit(t: torch.Tensor, owner: int, group, dst: Optional[int]) -> None:
    """Move one decoded unit from its owner to `dst` (point to point) or to every rank (broadcast)."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    g = lambda r: dist.get_global_rank(group, r)  # noqa: E731
    if dst is None:
        dist.broadcast(t, src=g(owner), group=group)
    elif owner != dst:
        if rank == owner:
            dist.send(t, dst=g(dst), group=group)
        elif rank == dst:
            dist.recv(t, src=g(owner), group=group)


def decode_latent_video(latent, decoder: SimpleVideoDecoder, timestep: Optional[float] = 0.05,
                        temporal_chunk_size: int = 7, temporal_overlap: int = 2, group=None,
                        dst: Optional[int] = None) -> Optional[torch.Tensor]:
    """The float video (B,3,T,H,W) decode_latent builds before its uint8 conversion (:704-790).

    With a torch.distributed `group` (one rank per GPU, every rank holding the same latent and decoder weights) the
    reference's temporal chunks are decoded round-robin across the ranks and exchanged before the cross-fade: to all
    ranks (dst=None, every rank returns the video) or only to rank `dst` (the others return None)."""
    x = to_device(latent, decoder.device)
    if x.ndim == 4:
        x = x[None]
    T = x.shape[2]
    world, rank = 1, 0
    if group is not None:              # e.g. torch.distributed.group.WORLD
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    sharded = decoder._shards is not None and world > 1
    if T <= temporal_chunk_size:
        if sharded:
            return decode_sharded(decoder, x, timestep, dst)
        if world > 1 and dst is not None and rank != dst:
            return None
        return decoder(x, timestep=timestep, show_progress=False)
    total = _pix_t(T)
    plan = chunk_plan(T, temporal_chunk_size, temporal_overlap)
    ref_overlap = _pix_t(temporal_overlap)
    B, _, _, H, W = decoder.output_shape(x.shape)
    # chunks land at frame (len_so_far - overlap); the reference concatenates, we blend into one buffer
    pieces = []
    length = 0
    # temporal shards with >= 4 ranks and >= 2 chunks: the two halves of the ranks decode two chunks side by side (a chunk
    # of 7 latent frames cannot use more than 7 ranks, and the small early stages do not speed up below one frame per
    # rank, so two 4-rank groups finish two chunks sooner than 8 ranks finish them one after the other)
    halves = sharded and world >= 4 and world % 2 == 0 and len(plan) >= 2
    shard_out = {}
    if halves:
        hw = world // 2
        for i0 in range(0, len(plan), 2):
            pair = plan[i0:i0 + 2]
            s0 = 2 * ((i0 // 2) % 2)          # clip slots alternate between rounds (see kVaeClipSlots, vae_engine.cu)
            for j, (a, b) in enumerate(pair):
                # (every rank calls this for every chunk, so the shared noise generator advances identically everywhere;
                #  only the ranks of the chunk's group launch kernels)
                decode_sharded(decoder, x[:, :, a:b].contiguous(), timestep, dst, group_ranks=(j * hw, hw), slot=s0 + j,
                               collect=False)
            for j, (a, b) in enumerate(pair):
                shard_out[i0 + j] = collect_clip(decoder, (x.shape[0], x.shape[1], b - a, x.shape[3], x.shape[4]), dst,
                                                 s0 + j)
    for i, (a, b) in enumerate(plan):
        if halves:
            v = shard_out[i]
        elif sharded:
            # temporal shards: ALL ranks work on every chunk (its frames are split over them); collected on dst / all
            v = decode_sharded(decoder, x[:, :, a:b].contiguous(), timestep, dst)
        elif unit_owner(i, world) == rank:
            v = decoder(x[:, :, a:b].contiguous(), timestep=timestep, show_progress=False)
        elif dst is None or rank == dst:
            v = torch.empty(decoder.output_shape((x.shape[0], x.shape[1], b - a, x.shape[3], x.shape[4])),
                            device=decoder.device, dtype=torch.float32)
        else:
            v = None
        n_t = _pix_t(b - a)
        ov = 0 if not pieces else min(ref_overlap, n_t, length)
        if ov <= 1:
            ov = 0
        pieces.append((v, length - ov, ov, n_t))
        length = length - ov + n_t
    if world > 1:
        if not sharded:
            for i, (v, _, _, _) in enumerate(pieces):
                if v is not None:
                    _exchange_unit(v, unit_owner(i, world), group, dst)
        if dst is not None and rank != dst:
            return None
    video = torch.empty(B, 3, length, H, W, device=decoder.device, dtype=torch.float32)
    with torch.cuda.device(decoder.device):
        for v, t0, ov, n_t in pieces:
            check(lib().ltx2_blend_chunk(ptr(video), ptr(v), B * 3, length, n_t, H * W, t0, ov, stream_ptr()),
                  "ltx2_blend_chunk")
    return video[:, :, :total]


def decode_latent(latent, decoder: SimpleVideoDecoder, timestep: Optional[float] = 0.05, key=None,
                  temporal_chunk_size: int = 7, temporal_overlap: int = 2, group=None,
                  dst: Optional[int] = None) -> Optional[torch.Tensor]:
    """Decode latent to uint8 frames (T,H,W,3) -- drop-in for simple_decoder.decode_latent (:676-800).
    `group` / `dst`: multi-GPU chunk sharding, see decode_latent_video."""
    video = decode_latent_video(latent, decoder, timestep, temporal_chunk_size, temporal_overlap, group, dst)
    if video is None:
        return None
    video = video.contiguous()
    B, _, T, H, W = video.shape
    out = torch.empty(T, H, W, 3, device=decoder.device, dtype=torch.uint8)
    with torch.cuda.device(decoder.device):
        check(lib().ltx2_video_to_uint8(ptr(video[0].contiguous()), ptr(out), T, H, W, stream_ptr()),
              "ltx2_video_to_uint8")
    return out
