"""ctypes binding of the C ABI in include/ltx2_b200.h.

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a) as
``ltx-2-mlx_b200/libltx2_b200.so``.  There is no fallback: if the library is missing or a
call fails, an exception is raised -- the product path never routes through the oracle or
a CPU/eager implementation.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libltx2_b200.so")

F32, BF16, F16 = 0, 1, 2


class Ltx2Error(RuntimeError):
    pass


class LtxDitConfig(C.Structure):
    _fields_ = [
        ("num_attention_heads", C.c_int32), ("attention_head_dim", C.c_int32), ("in_channels", C.c_int32),
        ("out_channels", C.c_int32), ("num_layers", C.c_int32), ("cross_attention_dim", C.c_int32),
        ("caption_channels", C.c_int32), ("cross_attention_adaln", C.c_int32), ("apply_gated_attention", C.c_int32),
        ("audio_enabled", C.c_int32), ("audio_heads", C.c_int32), ("audio_head_dim", C.c_int32),
        ("audio_in_channels", C.c_int32), ("audio_out_channels", C.c_int32), ("norm_eps", C.c_float),
        ("positional_embedding_theta", C.c_float), ("max_pos", C.c_float * 3), ("audio_max_pos", C.c_float),
        ("timestep_scale_multiplier", C.c_float), ("av_ca_timestep_scale_multiplier", C.c_float),
    ]


class LtxModalityView(C.Structure):
    _fields_ = [
        ("latent", C.c_void_p), ("latent_dtype", C.c_int32), ("context", C.c_void_p), ("context_dtype", C.c_int32),
        ("timesteps", C.c_void_p), ("sigma", C.c_void_p), ("positions", C.c_void_p), ("batch", C.c_int32),
        ("tokens", C.c_int32), ("context_tokens", C.c_int32), ("n_t", C.c_int32), ("n_dims", C.c_int32),
    ]


class LtxDitSkip(C.Structure):
    _fields_ = [("video_self_attn", C.c_uint64), ("audio_self_attn", C.c_uint64), ("a2v_cross_attn", C.c_uint64),
                ("v2a_cross_attn", C.c_uint64)]


class LtxVaeStage(C.Structure):
    _fields_ = [("kind", C.c_int32), ("num_layers", C.c_int32), ("stride_t", C.c_int32), ("stride_h", C.c_int32),
                ("stride_w", C.c_int32), ("multiplier", C.c_int32), ("residual", C.c_int32)]


class LtxVaeConfig(C.Structure):
    _fields_ = [("num_stages", C.c_int32), ("stages", LtxVaeStage * 16), ("base_channels", C.c_int32),
                ("latent_channels", C.c_int32), ("timestep_conditioning", C.c_int32)]


# every symbol include/ltx2_b200.h declares; tests check the library exports all of them
EXPORTS = [
    "ltx2_version", "ltx2_last_error",
    "ltx2_dit_create", "ltx2_dit_destroy", "ltx2_dit_set_weight", "ltx2_dit_missing_weights", "ltx2_dit_weight_keys", "ltx2_dit_weight_shape", "ltx2_dit_get_weight", "ltx2_dit_forward",
    "ltx2_dit_set_cross_attn_scale", "ltx2_dit_cp_init", "ltx2_dit_cp_connect", "ltx2_dit_set_profile", "ltx2_dit_profile_read", "ltx2_launch_count",
    "ltx2_vae_create", "ltx2_vae_destroy", "ltx2_vae_set_weight", "ltx2_vae_missing_weights", "ltx2_vae_output_shape",
    "ltx2_vae_decode", "ltx2_vae_set_profile", "ltx2_vae_profile_read", "ltx2_blend_chunk", "ltx2_video_to_uint8", "ltx2_tile_accumulate", "ltx2_tile_normalize",
    "ltx2_gemm_bf16", "ltx2_gemm_bf16_splitk", "ltx2_attention", "ltx2_attention_trace", "ltx2_attention_vrows", "ltx2_norm_modulate", "ltx2_headnorm_rope", "ltx2_v_transpose",
    "ltx2_rope_tables", "ltx2_timestep_sinusoid", "ltx2_small_linear", "ltx2_x0_from_velocity", "ltx2_denoise_update", "ltx2_gemm_plan", "ltx2_attention_plan", "ltx2_silu_mul",
    "ltx2_gelu_mul", "ltx2_interleaved_rope", "ltx2_cast",
]

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load libltx2_b200.so (once).  Raises Ltx2Error if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Ltx2Error(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(nvcc, sm_100a). There is no CPU fallback.")
        _lib = C.CDLL(LIB_PATH)
        _lib.ltx2_last_error.restype = C.c_char_p
        _lib.ltx2_dit_destroy.restype = None
        _lib.ltx2_launch_count.restype = C.c_int64
        _lib.ltx2_dit_weight_keys.restype = C.c_int64
        if hasattr(_lib, "ltx2_vae_destroy"):
            _lib.ltx2_vae_destroy.restype = None
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().ltx2_last_error()
        raise Ltx2Error(f"{what or 'ltx2 call'} failed ({status}): {msg.decode() if msg else ''}")


def ptr(t) -> C.c_void_p:
    """Device (or host) pointer of a torch tensor, None -> NULL."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def dtype_code(t) -> int:
    import torch
    return {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}[t.dtype]


def stream_ptr() -> C.c_void_p:
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
