"""Host-side mirror of the reference's DiT interface, backed by the sm_100a engine.

Same names, argument meaning and error behaviour as
``LTX_2_MLX/model/transformer/model.py``: ``LTXModelType`` (:18-29), ``Modality`` (:59-69),
``LTXModel`` (:413-881, ``__call__`` :776-881) and ``X0Model`` (:884-936), plus the STG
perturbation carriers of ``LTX_2_MLX/components/perturbations.py``.  Pipelines call these
objects exactly as they call the reference's (``transformer(video_modality[, audio_modality],
perturbations=...)``, pipelines/distilled.py:229-239); arrays may be torch tensors (any device)
or anything ``numpy.asarray`` accepts (the reference converts mx->numpy the same way,
tests/test_parity.py:322).  Outputs are fp32 CUDA torch tensors, stream-ordered: the caller
synchronises by reading them (``.cpu()``), the analogue of ``mx.eval``.

All arithmetic happens in the C-ABI engine (include/ltx2_b200.h); nothing here computes.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass
from enum import Enum
from typing import Any, Dict, Iterable, List, Optional, Tuple, Union

import numpy as np
import torch

from . import _lib
from ._lib import LtxDitConfig, LtxDitSkip, LtxModalityView, check, dtype_code, lib, ptr, stream_ptr


class LTXModelType(Enum):
    AudioVideo = "ltx av model"
    VideoOnly = "ltx video only model"
    AudioOnly = "ltx audio only model"

    def is_video_enabled(self) -> bool:
        return self in (LTXModelType.AudioVideo, LTXModelType.VideoOnly)

    def is_audio_enabled(self) -> bool:
        return self in (LTXModelType.AudioVideo, LTXModelType.AudioOnly)


class LTXRopeType(Enum):
    INTERLEAVED = "interleaved"
    SPLIT = "split"


@dataclass
class Modality:
    """Input modality data (video or audio) -- model.py:59-69."""

    latent: Any            # (B, T, C) patchified latents
    context: Any           # (B, S, C_ctx) text context
    context_mask: Any      # None in practice (pipelines/common.py:223-231)
    timesteps: Any         # (B,), (B, T) or (B, T, 1)
    positions: Any         # (B, n_dims, T, 2) [start, end) bounds
    enabled: bool = True
    sigma: Any = None      # (B,) scalar noise level (V2 prompt adaLN)
    # Engine extension (not in the reference): (class_values (n_cls,) fp32, row_cls (B, T) int32) with
    # timesteps[b, t] == class_values[row_cls[b, t]].  A sampling loop with a fixed denoise mask builds row_cls once
    # and rescales class_values per step, so the forward needs no host-side de-duplication of (B, T) timesteps
    # (sampling.euler_denoising_loop).  When given, `timesteps` is only used for its shape checks.
    timestep_classes: Any = None


# ---- STG perturbations (components/perturbations.py) ------------------------------------
class PerturbationType(Enum):
    SKIP_A2V_CROSS_ATTN = "skip_a2v_cross_attn"
    SKIP_V2A_CROSS_ATTN = "skip_v2a_cross_attn"
    SKIP_VIDEO_SELF_ATTN = "skip_video_self_attn"
    SKIP_AUDIO_SELF_ATTN = "skip_audio_self_attn"


@dataclass(frozen=True)
class Perturbation:
    type: PerturbationType
    blocks: Optional[List[int]] = None      # None = every block

    def is_perturbed(self, perturbation_type, block: int) -> bool:
        if self.type != perturbation_type:
            return False
        return self.blocks is None or block in self.blocks


@dataclass(frozen=True)
class PerturbationConfig:
    perturbations: Optional[List[Perturbation]] = None

    def is_perturbed(self, perturbation_type, block: int) -> bool:
        return bool(self.perturbations) and any(p.is_perturbed(perturbation_type, block) for p in self.perturbations)

    @staticmethod
    def empty() -> "PerturbationConfig":
        return PerturbationConfig(perturbations=[])


@dataclass(frozen=True)
class BatchedPerturbationConfig:
    perturbations: List[PerturbationConfig]

    def any_in_batch(self, perturbation_type, block: int) -> bool:
        return any(p.is_perturbed(perturbation_type, block) for p in self.perturbations)

    def all_in_batch(self, perturbation_type, block: int) -> bool:
        return all(p.is_perturbed(perturbation_type, block) for p in self.perturbations)

    @staticmethod
    def empty(batch_size: int) -> "BatchedPerturbationConfig":
        return BatchedPerturbationConfig(perturbations=[PerturbationConfig.empty() for _ in range(batch_size)])


def _type_member(perturbations, name: str):
    """The enum member called `name` in whatever enum class the caller's config uses (ours or the reference's)."""
    for cfg in getattr(perturbations, "perturbations", []) or []:
        for p in getattr(cfg, "perturbations", None) or []:
            return getattr(type(p.type), name)
    return getattr(PerturbationType, name)


def _skip_masks(perturbations, num_layers: int) -> Optional[LtxDitSkip]:
    if perturbations is None:
        return None
    sk = LtxDitSkip(0, 0, 0, 0)
    for field, name in (("video_self_attn", "SKIP_VIDEO_SELF_ATTN"), ("audio_self_attn", "SKIP_AUDIO_SELF_ATTN"),
                        ("a2v_cross_attn", "SKIP_A2V_CROSS_ATTN"), ("v2a_cross_attn", "SKIP_V2A_CROSS_ATTN")):
        member = _type_member(perturbations, name)
        bits = 0
        for i in range(num_layers):
            if perturbations.all_in_batch(member, i):      # transformer.py:486-501
                bits |= 1 << i
        setattr(sk, field, bits)
    return sk


# ---- array plumbing --------------------------------------------------------------------
_FLOATS = (torch.float32, torch.bfloat16, torch.float16)


def to_device(a, device, dtype=None) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(a)))
    if dtype is not None:
        t = t.to(dtype)
    elif t.dtype not in _FLOATS:
        t = t.to(torch.float32)
    return t.to(device, non_blocking=True).contiguous()


class _BlockHandle:
    """Stand-in for `model.transformer_blocks[i]`: OneStagePipeline sets `_cross_attn_scale`
    on it (pipelines/one_stage.py:207-222; transformer.py:526-528)."""

    def __init__(self, model: "LTXModel", idx: int):
        object.__setattr__(self, "_model", model)
        object.__setattr__(self, "idx", idx)
        object.__setattr__(self, "_scale", None)

    def __setattr__(self, name, value):
        if name == "_cross_attn_scale":
            object.__setattr__(self, "_scale", value)
            self._model._set_cross_attn_scale(self.idx, value)
        else:
            object.__setattr__(self, name, value)

    def __getattr__(self, name):
        if name == "_cross_attn_scale":
            scale = object.__getattribute__(self, "_scale")
            if scale is None:
                raise AttributeError(name)
            return scale
        raise AttributeError(name)

    def __delattr__(self, name):
        if name == "_cross_attn_scale":
            object.__setattr__(self, "_scale", None)
            self._model._set_cross_attn_scale(self.idx, None)
        else:
            object.__delattr__(self, name)


class LTXModel:
    """LTX-2 transformer (velocity model) -- drop-in for model.py:413-881."""

    AUDIO_ATTENTION_HEADS = 32
    AUDIO_HEAD_DIM = 64
    AUDIO_IN_CHANNELS = 128
    AUDIO_OUT_CHANNELS = 128
    AUDIO_CROSS_PE_MAX_POS = 20

    def __init__(
        self,
        model_type: LTXModelType = LTXModelType.VideoOnly,
        num_attention_heads: int = 32,
        attention_head_dim: int = 128,
        in_channels: int = 128,
        out_channels: int = 128,
        num_layers: int = 48,
        cross_attention_dim: int = 4096,
        norm_eps: float = 1e-6,
        caption_channels: Optional[int] = 3840,
        positional_embedding_theta: float = 10000.0,
        positional_embedding_max_pos: Optional[List[int]] = None,
        timestep_scale_multiplier: int = 1000,
        av_ca_timestep_scale_multiplier: int = 1,
        use_middle_indices_grid: bool = True,
        rope_type: LTXRopeType = LTXRopeType.SPLIT,
        compute_dtype: Any = None,
        low_memory: bool = False,
        fast_mode: bool = False,
        cross_attention_adaln: bool = False,
        apply_gated_attention: bool = False,
        device: Union[str, torch.device] = "cuda",
        fp8_linear: Optional[bool] = None,
    ):
        """`fp8_linear` (engine option, default: environment LTX2_FP8=1): keep the norm-fed linears (self-attention QKV,
        text-attention Q, FFN up) as E4M3 with their weight_scale and run them on the FP8 tensor pipe with per-token
        dynamic activation scales; FP8 checkpoint tensors of those layers are then ingested byte-for-byte."""
        if getattr(model_type, "name", None) in LTXModelType.__members__ and not isinstance(model_type, LTXModelType):
            model_type = LTXModelType[model_type.name]          # accept the reference's own enum
        if model_type == LTXModelType.AudioOnly:
            raise NotImplementedError("AudioOnly is not on the sampling hot path (SURVEY.md section 8)")
        if getattr(rope_type, "name", "SPLIT") != "SPLIT":
            raise NotImplementedError("the DiT uses SPLIT RoPE (model.py:455); INTERLEAVED is exposed as "
                                      "ltx2_b200.kernels.interleaved_rope only")
        if not use_middle_indices_grid:
            raise NotImplementedError("use_middle_indices_grid=False is never used by the reference pipelines")
        self.model_type = model_type
        self.rope_type = rope_type
        self.num_attention_heads = num_attention_heads
        self.num_layers = num_layers
        self.inner_dim = self.video_inner_dim = num_attention_heads * attention_head_dim
        self.audio_inner_dim = self.AUDIO_ATTENTION_HEADS * self.AUDIO_HEAD_DIM
        self.in_channels, self.out_channels = in_channels, out_channels
        self.caption_channels = caption_channels
        self.cross_attention_adaln = cross_attention_adaln
        self.apply_gated_attention = apply_gated_attention
        self.timestep_scale_multiplier = timestep_scale_multiplier
        self.positional_embedding_theta = positional_embedding_theta
        self.positional_embedding_max_pos = positional_embedding_max_pos or [20, 2048, 2048]
        self.norm_eps = norm_eps
        self.compute_dtype = compute_dtype
        self.device = torch.device(device)

        cfg = LtxDitConfig()
        cfg.num_attention_heads, cfg.attention_head_dim = num_attention_heads, attention_head_dim
        cfg.in_channels, cfg.out_channels, cfg.num_layers = in_channels, out_channels, num_layers
        cfg.cross_attention_dim = cross_attention_dim
        cfg.caption_channels = caption_channels or 0
        cfg.cross_attention_adaln = int(cross_attention_adaln)
        cfg.apply_gated_attention = int(apply_gated_attention)
        cfg.audio_enabled = int(model_type.is_audio_enabled())
        cfg.audio_heads, cfg.audio_head_dim = self.AUDIO_ATTENTION_HEADS, self.AUDIO_HEAD_DIM
        cfg.audio_in_channels, cfg.audio_out_channels = self.AUDIO_IN_CHANNELS, self.AUDIO_OUT_CHANNELS
        cfg.norm_eps = norm_eps
        cfg.positional_embedding_theta = positional_embedding_theta
        cfg.max_pos = (C.c_float * 3)(*[float(v) for v in self.positional_embedding_max_pos])
        cfg.audio_max_pos = float(self.AUDIO_CROSS_PE_MAX_POS)
        cfg.timestep_scale_multiplier = float(timestep_scale_multiplier)
        cfg.av_ca_timestep_scale_multiplier = float(av_ca_timestep_scale_multiplier)
        if fp8_linear is None:
            fp8_linear = os.environ.get("LTX2_FP8", "0") == "1"
        self.fp8_linear = bool(fp8_linear)
        cfg.fp8_linear = int(self.fp8_linear)
        self._cfg = cfg
        # text-context reuse across steps (V1 models; ltx2_dit_set_context_tag): on unless LTX2_CTX_CACHE=0
        self.reuse_context = os.environ.get("LTX2_CTX_CACHE", "1") != "0"
        self._ctx_ref = None                 # keeps the tagged context tensor alive so its id() stays unique
        self._cp = None                      # (rank, world, group, batch, n_total) once context_parallel.enable() ran
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().ltx2_dit_create(C.byref(cfg), C.byref(self._h)), "ltx2_dit_create")
        self.transformer_blocks = [_BlockHandle(self, i) for i in range(num_layers)]

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                lib().ltx2_dit_destroy(h)
            except Exception:
                pass
            self._h = C.c_void_p()

    # ---- weights ------------------------------------------------------------------------
    def load_weights(self, weights: Iterable[Tuple[str, Any]], strict: bool = False) -> int:
        """Mirror of mlx ``Module.load_weights(list_of_(key, array))`` used by the LoRA fuse/restore
        paths (scripts/generate.py:1198-1200, pipelines/two_stage.py:180-186).  Keys are the
        reference's MLX-side names.  Returns the number of tensors consumed."""
        n = 0
        with torch.cuda.device(self.device):
            for item in (weights.items() if isinstance(weights, dict) else weights):
                # (key, tensor) or (key, fp8_tensor, weight_scale): an FP8 checkpoint tensor with its per-tensor scale
                key, value = item[0], item[1]
                scale = float(item[2]) if len(item) > 2 else 1.0
                if isinstance(value, torch.Tensor) and value.dtype == torch.float8_e4m3fn:
                    t = value.to(self.device, non_blocking=True).contiguous()
                else:
                    t = to_device(value, self.device)
                shape = (C.c_int64 * max(t.ndim, 1))(*(t.shape if t.ndim else (1,)))
                st = lib().ltx2_dit_set_weight_scaled(self._h, key.encode(), ptr(t), dtype_code(t), shape,
                                                      max(t.ndim, 1), scale, stream_ptr())
                if st == -2 and not strict:       # LTX2_ERR_NOKEY: not a tensor of this model
                    continue
                check(st, f"set_weight({key})")
                n += 1
            torch.cuda.current_stream().synchronize()
        return n

    def update(self, nested: Dict[str, Any]) -> "LTXModel":
        """Mirror of mlx ``Module.update(nested_dict)`` (weight_converter.py:438-443)."""
        flat: List[Tuple[str, Any]] = []

        def walk(prefix, node):
            if isinstance(node, dict):
                for k, v in node.items():
                    walk(f"{prefix}.{k}" if prefix else str(k), v)
            elif isinstance(node, (list, tuple)):
                for i, v in enumerate(node):
                    walk(f"{prefix}.{i}", v)
            else:
                flat.append((prefix, node))

        walk("", nested)
        self.load_weights(flat)
        return self

    def weight_keys(self) -> List[str]:
        n = lib().ltx2_dit_weight_keys(self._h, None, C.c_int64(0))
        buf = C.create_string_buffer(int(n) + 16)
        lib().ltx2_dit_weight_keys(self._h, buf, C.c_int64(len(buf)))
        return sorted(k for k in buf.value.decode().split("\n") if k)

    def get_weight(self, key: str, dtype: torch.dtype = torch.float32) -> torch.Tensor:
        """One tensor of the flat {reference_key: array} view (engine storage converted to `dtype`)."""
        shape = (C.c_int64 * 2)()
        nd = lib().ltx2_dit_weight_shape(self._h, key.encode(), shape)
        if nd < 0:
            raise KeyError(key)
        dims = tuple(shape[:nd])
        out = torch.empty(dims, device=self.device, dtype=dtype)
        with torch.cuda.device(self.device):
            check(lib().ltx2_dit_get_weight(self._h, key.encode(), ptr(out), dtype_code(out), C.c_int64(out.numel()),
                                            stream_ptr()), f"get_weight({key})")
        return out

    def parameters(self) -> Dict[str, torch.Tensor]:
        """Flat {reference_key: tensor} view, materialised lazily per key (the reference returns a nested dict that
        callers immediately flatten with mlx.utils.tree_flatten)."""
        model = self

        class _Lazy(dict):
            def __missing__(self, key):
                return model.get_weight(key)

            def keys(self):
                return model.weight_keys()

            def __iter__(self):
                return iter(model.weight_keys())

            def __len__(self):
                return len(model.weight_keys())

            def items(self):
                return ((k, model.get_weight(k)) for k in model.weight_keys())

        return _Lazy()

    def missing_weights(self) -> List[str]:
        buf = C.create_string_buffer(8192)
        n = lib().ltx2_dit_missing_weights(self._h, buf, C.c_int64(len(buf)))
        return [s for s in buf.value.decode().split("\n") if s] if n else []

    def _context_tag(self, context) -> int:
        """Non-zero tag when `context` is provably the tensor (same object, same torch version counter) whose projected
        K/V the engine may hold from the previous call; 0 = recompute.  Only torch tensors carry a version counter, so
        numpy / mx arrays are never reused.  (Writes that bypass torch, e.g. through a numpy view of the same memory,
        are invisible to the counter: call reset_context_cache() after such a write.)"""
        if not self.reuse_context or not isinstance(context, torch.Tensor):
            self._ctx_ref = None
            return 0
        self._ctx_ref = context
        key = (id(context), context._version, context.data_ptr(), tuple(context.shape), str(context.dtype),
               getattr(self, "_ctx_salt", 0))
        return (hash(key) & 0x7FFFFFFFFFFFFFFF) | 1

    def reset_context_cache(self) -> None:
        self._ctx_ref = None
        check(lib().ltx2_dit_set_context_tag(self._h, 0), "ltx2_dit_set_context_tag")
        # a forward with tag 0 drops the cached K/V; until then make sure no stale tag can match
        self._ctx_salt = getattr(self, "_ctx_salt", 0) + 1

    def set_layer_limit(self, n: Optional[int]) -> None:
        """Diagnostics (bench.py parity leg): run only the first n blocks, then the head; None/0 = all."""
        check(lib().ltx2_dit_set_layer_limit(self._h, int(n or 0)), "ltx2_dit_set_layer_limit")

    def _set_cross_attn_scale(self, idx: int, value) -> None:
        check(lib().ltx2_dit_set_cross_attn_scale(self._h, idx, C.c_float(float("nan") if value is None else float(value))),
              "set_cross_attn_scale")

    # ---- forward ------------------------------------------------------------------------
    def _view(self, m: Modality, n_dims: int, keep: list) -> LtxModalityView:
        dev = self.device
        if m.context_mask is not None:
            mask = to_device(m.context_mask, dev, torch.float32)
            is_bool = isinstance(m.context_mask, torch.Tensor) and m.context_mask.dtype == torch.bool
            trivially_open = bool((mask != 0).all()) if is_bool or mask.max() <= 1 else bool((mask == 0).all())
            if not trivially_open:
                raise NotImplementedError("context_mask that hides tokens: the reference pipelines always pass None "
                                          "(pipelines/common.py:223-231)")
        latent = to_device(m.latent, dev)
        context = to_device(m.context, dev)
        if n_dims == 3:
            check(lib().ltx2_dit_set_context_tag(self._h, self._context_tag(m.context)), "ltx2_dit_set_context_tag")
        if latent.ndim != 3 or context.ndim != 3:
            raise ValueError(f"latent/context must be (B, T, C); got {tuple(latent.shape)} / {tuple(context.shape)}")
        B, N, _ = latent.shape
        ts = to_device(m.timesteps, dev, torch.float32)
        if ts.ndim == 0:
            ts = ts.reshape(1).expand(B)
        if ts.ndim == 3:
            ts = ts[:, :, 0]
        ts = ts.reshape(B, -1).contiguous()
        if ts.shape[1] not in (1, N):
            raise ValueError(f"timesteps must be (B,), (B, T) or (B, T, 1); got {tuple(ts.shape)} for T={N}")
        cls_vals = row_cls = None
        if getattr(m, "timestep_classes", None) is not None:
            cls_vals, row_cls = m.timestep_classes
            cls_vals = to_device(cls_vals, dev, torch.float32).reshape(-1)
            row_cls = to_device(row_cls, dev, torch.int32).reshape(B, N)
            if not 1 <= cls_vals.numel() <= 64:
                raise ValueError("timestep_classes: 1..64 classes")
            if m.sigma is None:
                raise ValueError("timestep_classes needs Modality.sigma")
        pos = to_device(m.positions, dev, torch.float32)
        if pos.ndim != 4 or pos.shape[1] != n_dims or pos.shape[-1] != 2:
            # rope.py:229,262-263 assert the same
            raise AssertionError(f"positions must be (B, {n_dims}, T, 2); got {tuple(pos.shape)}")
        sigma = None
        if m.sigma is not None:
            sigma = to_device(m.sigma, dev, torch.float32).reshape(-1)
            if sigma.numel() == 1 and B > 1:
                sigma = sigma.expand(B).contiguous()
        if self._cp is not None and n_dims == 3:
            from .context_parallel import slice_tokens
            rank, world, _, cp_b, cp_n = self._cp
            if (B, N) != (cp_b, cp_n):
                raise ValueError(f"context parallel was enabled for batch {cp_b} x {cp_n} tokens, got {B} x {N}")
            if sigma is None and ts.shape[1] != 1:
                # the scalar sigma of a modality defaults to its FIRST token's timestep (model.py:250-260,394-399);
                # take it before slicing so every rank uses token 0 of the whole sequence
                sigma = ts[:, 0].contiguous()
            latent, ts, pos = slice_tokens(latent, ts, pos, rank, world)
            if row_cls is not None:
                a = rank * latent.shape[1]
                row_cls = row_cls[:, a:a + latent.shape[1]].contiguous()
            N = latent.shape[1]
        keep.extend([latent, context, ts, pos, sigma, cls_vals, row_cls])
        v = LtxModalityView()
        v.latent, v.latent_dtype = latent.data_ptr(), dtype_code(latent)
        v.context, v.context_dtype = context.data_ptr(), dtype_code(context)
        v.timesteps = cls_vals.data_ptr() if cls_vals is not None else ts.data_ptr()
        v.row_cls = row_cls.data_ptr() if row_cls is not None else None
        v.n_cls = int(cls_vals.numel()) if cls_vals is not None else 0
        v.sigma = sigma.data_ptr() if sigma is not None else None
        v.positions = pos.data_ptr()
        v.batch, v.tokens, v.context_tokens = B, N, context.shape[1]
        v.n_t, v.n_dims = ts.shape[1], n_dims
        return v

    def _forward(self, video, audio, perturbations, x0: bool):
        if self.model_type.is_video_enabled() and video is None:
            raise ValueError("Video modality required for video-enabled model")        # model.py:824
        keep: list = []
        with torch.cuda.device(self.device):
            vv = self._view(video, 3, keep)
            av = None
            if audio is not None and self.model_type.is_audio_enabled() and getattr(audio, "enabled", True):
                a_lat = np.asarray(audio.latent.shape if hasattr(audio.latent, "shape") else np.shape(audio.latent))
                if int(np.prod(a_lat)) > 0:
                    av = self._view(audio, 1, keep)
            out_v = torch.empty(vv.batch, vv.tokens, self.out_channels, device=self.device, dtype=torch.float32)
            out_a = None
            if av is not None:
                out_a = torch.empty(av.batch, av.tokens, self.AUDIO_OUT_CHANNELS, device=self.device,
                                    dtype=torch.float32)
            sk = _skip_masks(perturbations, self.num_layers)
            check(lib().ltx2_dit_forward(self._h, C.byref(vv), C.byref(av) if av is not None else None,
                                         C.byref(sk) if sk is not None else None, int(x0), ptr(out_v), ptr(out_a),
                                         stream_ptr()), "ltx2_dit_forward")
        if self._cp is not None:
            from .context_parallel import gather_tokens
            out_v = gather_tokens(out_v, self._cp[2])
        if self.model_type == LTXModelType.VideoOnly:
            return out_v
        if out_a is None:
            out_a = torch.zeros(vv.batch, 0, self.AUDIO_OUT_CHANNELS, device=self.device)
        return out_v, out_a

    def __call__(self, video: Optional[Modality] = None, audio: Optional[Modality] = None, perturbations=None):
        """VideoOnly: video velocity (B,N,C) fp32.  AudioVideo: (video_velocity, audio_velocity)."""
        return self._forward(video, audio, perturbations, x0=False)


class X0Model:
    """Wrapper that returns denoised outputs instead of velocities -- model.py:884-936."""

    def __init__(self, velocity_model: LTXModel):
        self.velocity_model = velocity_model

    def __call__(self, video: Optional[Modality] = None, audio: Optional[Modality] = None, perturbations=None):
        out = self.velocity_model._forward(video, audio, perturbations, x0=True)
        if isinstance(out, tuple):
            if audio is None:
                return out[0]            # video-only inference on the AV model returns video only (:926-928)
            return out
        return out


LTXAVModel = LTXModel
X0AVModel = X0Model
