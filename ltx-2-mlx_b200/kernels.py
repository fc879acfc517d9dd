"""Drop-in for ``LTX_2_MLX/kernels/fused_ops.py``: ``silu_mul`` (:50), ``gelu_mul`` (:95), ``interleaved_rope`` (:183).

Same names, shapes and assertions; the Metal shader strings are replaced by CUDA kernels behind the C ABI
(``ltx2_silu_mul`` / ``ltx2_gelu_mul`` / ``ltx2_interleaved_rope``).  Inputs may be torch tensors (any device) or
array-likes; outputs are CUDA torch tensors of the input dtype (float32, bfloat16 or float16)."""
from __future__ import annotations

import torch

from . import ops
from .transformer import to_device


def _prep(*arrays):
    dev = torch.device("cuda", torch.cuda.current_device())
    ts = [to_device(a, dev) for a in arrays]
    return ts


def silu_mul(a, b) -> torch.Tensor:
    """silu(a) * b"""
    a, b = _prep(a, b)
    assert a.shape == b.shape, f"Shape mismatch: {tuple(a.shape)} vs {tuple(b.shape)}"
    return ops.silu_mul(a, b.to(a.dtype))


def gelu_mul(a, b) -> torch.Tensor:
    """gelu_approx(a) * b (tanh form)"""
    a, b = _prep(a, b)
    assert a.shape == b.shape, f"Shape mismatch: {tuple(a.shape)} vs {tuple(b.shape)}"
    return ops.gelu_mul(a, b.to(a.dtype))


def interleaved_rope(x, cos_freqs, sin_freqs) -> torch.Tensor:
    """Rotate pairs (x[2i], x[2i+1]); cos/sin broadcastable to x's shape (fused_ops.py:183-242)."""
    x, c, s = _prep(x, cos_freqs, sin_freqs)
    c = torch.broadcast_to(c.to(x.dtype), x.shape).contiguous()
    s = torch.broadcast_to(s.to(x.dtype), x.shape).contiguous()
    return ops.interleaved_rope(x, c, s)


__all__ = ["silu_mul", "gelu_mul", "interleaved_rope"]
