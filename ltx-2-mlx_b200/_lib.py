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

F32, BF16, F16, F8E4M3 = 0, 1, 2, 3


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
        ("fp8_linear", C.c_int32),
    ]


class LtxModalityView(C.Structure):
    _fields_ = [
        ("latent", C.c_void_p), ("latent_dtype", C.c_int32), ("context", C.c_void_p), ("context_dtype", C.c_int32),
        ("timesteps", C.c_void_p), ("sigma", C.c_void_p), ("positions", C.c_void_p), ("batch", C.c_int32),
        ("tokens", C.c_int32), ("context_tokens", C.c_int32), ("n_t", C.c_int32), ("n_dims", C.c_int32),
        ("row_cls", C.c_void_p), ("n_cls", C.c_int32),
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


# Every symbol include/ltx2_b200.h declares, with its C signature (restype, argtypes).  Declaring them once here means
# ctypes converts Python ints/floats to the right width at every call site (an undeclared `int64_t` argument would be
# marshalled as a 32-bit int and silently truncated); tests check the library exports all of them.
_P, _I32, _I64, _U64, _F, _D, _S = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_double, C.c_char_p
SIGNATURES = {
    "ltx2_version": (_I32, []),
    "ltx2_last_error": (_S, []),
    "ltx2_dit_create": (_I32, [_P, _P]),
    "ltx2_dit_destroy": (None, [_P]),
    "ltx2_dit_set_weight": (_I32, [_P, _S, _P, _I32, _P, _I32, _P]),
    "ltx2_dit_set_weight_scaled": (_I32, [_P, _S, _P, _I32, _P, _I32, _F, _P]),
    "ltx2_dit_weight_keys": (_I64, [_P, _P, _I64]),
    "ltx2_dit_weight_shape": (_I32, [_P, _S, _P]),
    "ltx2_dit_get_weight": (_I32, [_P, _S, _P, _I32, _I64, _P]),
    "ltx2_dit_missing_weights": (_I32, [_P, _P, _I64]),
    "ltx2_dit_forward": (_I32, [_P, _P, _P, _P, _I32, _P, _P, _P]),
    "ltx2_dit_set_context_tag": (_I32, [_P, _U64]),
    "ltx2_dit_set_layer_limit": (_I32, [_P, _I32]),
    "ltx2_dit_set_cross_attn_scale": (_I32, [_P, _I32, _F]),
    "ltx2_dit_set_profile": (_I32, [_P, _I32]),
    "ltx2_dit_profile_read": (_I32, [_P, _P, _P, _P, _I32]),
    "ltx2_dit_profile_launch": (_I32, [_P, _I32, _P, _P, _P]),
    "ltx2_launch_count": (_I64, []),
    "ltx2_dit_cp_init": (_I32, [_P, _I32, _I32, _I32, _I32, _I32, _P]),
    "ltx2_dit_cp_connect": (_I32, [_P, _P]),
    "ltx2_dit_cp_set_split_k": (_I32, [_P, _I32]),
    "ltx2_dit_cp_shutdown": (_I32, [_P, _I32]),
    "ltx2_vae_create": (_I32, [_P, _P]),
    "ltx2_vae_destroy": (None, [_P]),
    "ltx2_vae_set_weight": (_I32, [_P, _S, _P, _I32, _P, _I32, _P]),
    "ltx2_vae_missing_weights": (_I32, [_P, _P, _I64]),
    "ltx2_vae_output_shape": (_I32, [_P, _P, _P]),
    "ltx2_vae_decode": (_I32, [_P, _P, _I32, _P, _F, _F, _P, _I32, _P, _P]),
    "ltx2_vae_cp_init": (_I32, [_P, _I32, _I32, _P, _P]),
    "ltx2_vae_cp_connect": (_I32, [_P, _P]),
    "ltx2_vae_cp_shutdown": (_I32, [_P, _I32]),
    "ltx2_vae_shard_frames": (_I32, [_P, _I64, _I32, _I32, _P, _P]),
    "ltx2_vae_decode_sharded": (_I32, [_P, _P, _I32, _P, _F, _F, _P, _I32, _I32, _I32, _I32, _P]),
    "ltx2_vae_cp_collect": (_I32, [_P, _I32, _P, _I32, _P, _P]),
    "ltx2_conv3d_workspace_bytes": (_I64, [_I32] * 6),
    "ltx2_conv3d": (_I32, [_P, _P, _I32, _P, _I32, _P] + [_I32] * 7 + [_P, _P]),
    "ltx2_pad_act": (_I32, [_P, _P, _P] + [_I32] * 9 + [_P, _P, _P, _I32, _F, _P, _P]),
    "ltx2_group_stats": (_I32, [_P, _I32, _I64, _I32, _I32, _F, _P, _P]),
    "ltx2_conv3d_pack": (_I32, [_P, _I32, _P, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "ltx2_conv3d_packed": (_I32, [_P, _P, _P, _P, _P] + [_I32] * 7 + [_P]),
    "ltx2_patchify_video": (_I32, [_P, _P] + [_I32] * 5 + [_P]),
    "ltx2_space_to_depth_residual": (_I32, [_P, _P, _P] + [_I32] * 10 + [_P]),
    "ltx2_pixel_shuffle2": (_I32, [_P, _P, _I64, _I32, _I32, _I32, _P]),
    "ltx2_ndhwc_to_ncdhw": (_I32, [_P, _P, _I32, _I32, _I32, _I64, _P, _P, _P]),
    "ltx2_ncdhw_to_ndhwc": (_I32, [_P, _I32, _P, _I32, _I32, _I64, _P]),
    "ltx2_vae_set_profile": (_I32, [_P, _I32]),
    "ltx2_vae_profile_read": (_I32, [_P, _P, _P, _P]),
    "ltx2_vae_profile_launch": (_I32, [_P, _I32, _P, _P]),
    "ltx2_blend_chunk": (_I32, [_P, _P] + [_I32] * 6 + [_P]),
    "ltx2_video_to_uint8": (_I32, [_P, _P, _I32, _I32, _I32, _P]),
    "ltx2_tile_accumulate": (_I32, [_P, _P, _P] + [_I32] * 13 + [_P, _P, _P, _P]),
    "ltx2_tile_normalize": (_I32, [_P, _P, _I32, _I64, _P]),
    "ltx2_gemm_bf16": (_I32, [_P, _I64, _P, _I64, _I32, _I32, _I32, _I32, _P, _P, _I64, _P, _I64, _P, _F, _P]),
    "ltx2_gemm_e4m3": (_I32, [_P, _I64, _P, _I64, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _I64, _P]),
    "ltx2_norm_modulate_q8": (_I32, [_P, _I32, _I64, _P, _I64, _P, _P, _I64, _I32, _I32, _I32, _F, _P, _I64, _I64, _I64,
                                     _P, _P]),
    "ltx2_quantize_rows_e4m3": (_I32, [_P, _I32, _I64, _I64, _P, _P, _P]),
    "ltx2_gemm_bf16_splitk": (_I32, [_P, _I64, _P, _I64, _I32, _I32, _I32, _P, _P, _I64, _P, _I64, _P, _F, _I32, _P]),
    "ltx2_attention": (_I32, [_P, _P, _P, _P] + [_I32] * 6 + [_F, _P, _P, _P]),
    "ltx2_attention_vrows": (_I32, [_P, _P, _P, _I64, _I64, _I64, _P] + [_I32] * 5 + [_F, _P, _P, _P]),
    "ltx2_attention_vrows_trace": (_I32, [_P, _P, _P, _I64, _I64, _I64, _P] + [_I32] * 5 + [_F, _P, _P]),
    "ltx2_attention_trace": (_I32, [_P, _P, _P, _P] + [_I32] * 6 + [_F, _P, _P]),
    "ltx2_norm_modulate": (_I32, [_P, _I32, _I64, _P, _I64, _I32, _I32, _I32, _F, _P, _I64, _I64, _I64, _P, _P]),
    "ltx2_headnorm_rope": (_I32, [_P, _I64, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _F, _P]),
    "ltx2_v_transpose": (_I32, [_P, _I64, _P] + [_I32] * 5 + [_P]),
    "ltx2_rope_tables": (_I32, [_P, _I32, _I32, _I32, _I32, _P, _F, _P, _P, _P]),
    "ltx2_timestep_sinusoid": (_I32, [_P, _I32, _F, _P, _P]),
    "ltx2_small_linear": (_I32, [_P, _I32, _I32, _P, _P, _P, _I32, _I32, _P]),
    "ltx2_x0_from_velocity": (_I32, [_P, _P, _P, _P, _I32, _I32, _P]),
    "ltx2_gemm_plan": (_I32, [_I32] * 5 + [_P]),
    "ltx2_attention_plan": (_I32, [_I32, _I32, _P, _P]),
    "ltx2_attention_sm_pair_plan": (_I32, [_I32, _I32, _I32, _P, _P]),
    "ltx2_attention_sm_pair_segments": (_I32, [_I32, _I32, _I32, _I32, _P, _I32]),
    "ltx2_denoise_update": (_I32, [_P, _P, _P, _F, _P, _P, _F, _F, _P, _P, _I32, _I32, _P]),
    "ltx2_silu_mul": (_I32, [_P, _P, _P, _I64, _I32, _P]),
    "ltx2_gelu_mul": (_I32, [_P, _P, _P, _I64, _I32, _P]),
    "ltx2_interleaved_rope": (_I32, [_P, _P, _P, _P, _I64, _I32, _P]),
    "ltx2_cast": (_I32, [_P, _I32, _P, _I32, _I64, _P]),
}
EXPORTS = list(SIGNATURES)

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load libltx2_b200.so (once) and declare every export's signature.  Raises Ltx2Error if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Ltx2Error(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name, None)
            if fn is None:
                raise Ltx2Error(f"{LIB_PATH} does not export {name}: rebuild it (__graft_entry__.build())")
            fn.restype = res
            fn.argtypes = args
        _lib = L
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
    return {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16, torch.float8_e4m3fn: F8E4M3}[t.dtype]


def stream_ptr() -> C.c_void_p:
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
