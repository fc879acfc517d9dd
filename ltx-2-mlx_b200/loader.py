"""Checkpoint ingest: reference safetensors names -> engine weight slots.

Mirrors the call surface of ``load_transformer_weights(model, path, ...)``
(LTX_2_MLX/loader/weight_converter.py:318-326) and ``load_vae_decoder_weights(decoder, path)``
(LTX_2_MLX/model/video_vae/simple_decoder.py:566).  Tensors stream one at a time from the
file to the device; the engine converts to its storage type (bf16 matrices, fp32
biases/norm weights/adaLN tables) on the GPU, so there is no torch->numpy->fp32 hop.
FP8 (E4M3) checkpoint tensors are handed over as bytes + weight_scale (``loader/fp8_loader.py:14-130``
semantics): an engine built with ``fp8_linear`` keeps them quantised for its FP8 GEMMs.
"""
from __future__ import annotations

import re
from typing import Dict, Iterable, Iterator, Optional, Tuple

import torch

DIT_PREFIX = "model.diffusion_model."

# weight_converter.py:300-313 -- PyTorch Sequential names -> the reference's module attribute names
_RENAMES = (
    (re.compile(r"\.to_out\.0\."), ".to_out."),
    (re.compile(r"\.(audio_)?ff\.net\.0\.proj\."), r".\1ff.project_in.proj."),
    (re.compile(r"\.(audio_)?ff\.net\.2\."), r".\1ff.project_out."),
)


def convert_pytorch_key_to_mlx(pytorch_key: str, include_audio: bool = False) -> Optional[str]:
    """Same contract as weight_converter.py:277-315: key without the ``model.diffusion_model.`` prefix in,
    engine/MLX-side key out; None for tensors the DiT does not own."""
    key = pytorch_key
    if not include_audio and ("av_ca" in key or "a2v" in key or "audio" in key.lower()):
        return None
    if "video_embeddings_connector" in key or "audio_embeddings_connector" in key:
        return None
    for rx, rep in _RENAMES:
        key = rx.sub(rep, key)
    return key


def iter_engine_weights(tensors: Iterable[Tuple[str, torch.Tensor]], include_audio: bool,
                        fp8_scales: Optional[Dict[str, float]] = None) -> Iterator[tuple]:
    for ck, t in tensors:
        if not ck.startswith(DIT_PREFIX) or ck.endswith(".weight_scale"):
            continue
        key = convert_pytorch_key_to_mlx(ck[len(DIT_PREFIX):], include_audio=include_audio)
        if key is None:
            continue
        if t.dtype == torch.float8_e4m3fn:
            # fp8_loader.py:14-32: value = weight_fp8 * weight_scale.  The bytes and the scale go to the engine as they
            # are: FP8-computed linears keep them (cfg.fp8_linear), all other layers widen to bf16 on the device.
            yield key, t, float(fp8_scales.get(ck, 1.0)) if fp8_scales else 1.0
            continue
        if fp8_scales and ck in fp8_scales:
            t = t.to(torch.float32) * fp8_scales[ck]
        elif t.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            t = t.to(torch.float32)
        yield key, t


def load_transformer_state_dict(model, weights: Dict[str, torch.Tensor], include_audio: Optional[bool] = None,
                                strict: bool = True) -> int:
    """Load an in-memory checkpoint (reference checkpoint key names) into an ltx2_b200 LTXModel."""
    if include_audio is None:
        include_audio = model.model_type.is_audio_enabled()
    n = model.load_weights(iter_engine_weights(weights.items(), include_audio), strict=False)
    missing = model.missing_weights()
    if strict and missing:
        raise KeyError(f"{len(missing)} DiT tensors missing from the checkpoint, e.g. {missing[:4]}")
    return n


def load_transformer_weights(model, weights_path: str, strict: bool = False, use_fp8: bool = False,
                             include_audio: bool = False, streaming: bool = True, target_dtype: str = "float16") -> None:
    """Drop-in for weight_converter.load_transformer_weights.  `target_dtype` is accepted for signature
    compatibility; the engine always stores matrices as bf16 and small tensors as fp32."""
    from safetensors import safe_open

    fp8_scales: Dict[str, float] = {}
    with safe_open(weights_path, framework="pt") as f:
        keys = list(f.keys())
        if use_fp8:
            for k in keys:
                if k.endswith(".weight_scale"):
                    fp8_scales[k.replace(".weight_scale", ".weight")] = float(f.get_tensor(k).item())

        def gen():
            for k in keys:
                if k.startswith(DIT_PREFIX):
                    yield k, f.get_tensor(k)

        n = model.load_weights(iter_engine_weights(gen(), include_audio, fp8_scales), strict=False)
    missing = model.missing_weights()
    print(f"  Converted {n} weight tensors ({len(missing)} engine tensors still unset)")
    if strict and missing:
        raise KeyError(f"missing DiT tensors: {missing[:8]}")
