"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch fp32) of the reference's text-embeddings connector
(SURVEY.md 8(f) rank 4).  Imported only by tests/.  Pinned by tests/golden/connector.npz, produced by the reference's own
`Embeddings1DConnector` over the restated mlx primitives of oracle/_mlx_shim (tests/golden/make_golden.py).

Reference map (/root/reference/LTX_2_MLX/model/text_encoder/connector.py):
  BasicTransformerBlock1D: rms_norm -> Attention(self, RoPE) -> +x; rms_norm -> FeedForward -> +x ..... :13-101
  _append_learnable_registers: tile the registers to max(1024, T) rounded up, append rows [T:] ........ :175-228
  __call__: registers -> 1-D RoPE over arange(T) (max_pos [1]) -> blocks -> rms_norm .................. :230-283
RoPE tables: model/transformer/rope.py:365-418 with indices_grid (B,1,T), use_middle_indices_grid=False; INTERLEAVED
(:331-362, applied on (B,T,H*Dh) before the head split, :51-89) or SPLIT (:292-328).
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch

from . import dit_oracle as O


def append_registers(x: torch.Tensor, registers: torch.Tensor) -> torch.Tensor:
    B, T, D = x.shape
    n = registers.shape[0]
    dup = math.ceil(max(1024, T) / n)
    tiled = registers.repeat(dup, 1)
    extra = tiled[T:]
    if extra.shape[0] > 0:
        x = torch.cat([x, extra[None].expand(B, -1, -1).to(x.dtype)], dim=1)
    return x


def rope_1d(T: int, dim: int, heads: int, theta: float, max_pos: float, rope_type: str,
            double_precision: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """precompute_freqs_cis for positions arange(T): INTERLEAVED -> cos/sin (1,T,dim); SPLIT -> (1,H,T,dim/(2H)).

    With max_pos = [1] the angles reach theta*pi/2 * (2T-3) ~ 3e7 rad, where ONE float32 ulp of the frequency grid is a
    phase error of about 2 rad: the tables are only reproducible if the grid is computed by the very same float32
    operations.  numpy float32 is used here because that is what pins the golden vectors (the shim evaluates
    `theta ** mx.linspace(...)` with numpy); `double_precision` is the reference's generate_freq_grid_np (rope.py:147-178)."""
    n = dim // 2
    if double_precision:
        idx = (np.power(theta, np.linspace(0.0, 1.0, n, dtype=np.float64)) * math.pi / 2).astype(np.float32)
    else:
        lin = np.linspace(0.0, 1.0, n).astype(np.float32)
        idx = ((np.float32(theta) ** lin) * np.float32(math.pi / 2)).astype(np.float32)
    scaled = (np.arange(T, dtype=np.float32) / np.float32(max_pos)) * np.float32(2) - np.float32(1)
    freqs = (idx[None, :] * scaled[:, None]).astype(np.float32)[None]   # (1,T,dim/2)
    cos, sin = torch.from_numpy(np.cos(freqs)), torch.from_numpy(np.sin(freqs))
    if rope_type == "split":
        cos = cos.reshape(1, T, heads, -1).permute(0, 2, 1, 3)
        sin = sin.reshape(1, T, heads, -1).permute(0, 2, 1, 3)
        return cos.contiguous(), sin.contiguous()
    return cos.repeat_interleave(2, dim=-1), sin.repeat_interleave(2, dim=-1)     # dim % 2 == 0: no identity padding


def attention(w, prefix: str, x: torch.Tensor, heads: int, pe, rope_type: str) -> torch.Tensor:
    q = O.linear(w, prefix + ".to_q", x)
    k = O.linear(w, prefix + ".to_k", x)
    v = O.linear(w, prefix + ".to_v", x)
    q = O.rms_norm(q, w[prefix + ".q_norm.weight"].to(x.dtype))
    k = O.rms_norm(k, w[prefix + ".k_norm.weight"].to(x.dtype))
    if rope_type == "split":
        q, k = O.apply_split_rope(q, *pe), O.apply_split_rope(k, *pe)
    else:
        q, k = O.interleaved_rope(q, *pe), O.interleaved_rope(k, *pe)
    out = O.sdpa(q, k, v, heads)
    if prefix + ".to_gate_logits.weight" in w:
        g = 2.0 * torch.sigmoid(O.linear(w, prefix + ".to_gate_logits", x))
        B, T, inner = out.shape
        out = (out.reshape(B, T, heads, inner // heads) * g[..., None]).reshape(B, T, inner)
    return O.linear(w, prefix + ".to_out", out)


def connector(w: Dict[str, torch.Tensor], x: torch.Tensor, *, heads: int, layers: int, theta: float = 10000.0,
              max_pos: float = 1.0, rope_type: str = "interleaved", eps: float = 1e-6) -> torch.Tensor:
    """Embeddings1DConnector.__call__: (B,T,D) -> (B,max(1024,T) rounded up to the register count,D)."""
    x = x.float()
    if "learnable_registers" in w:
        x = append_registers(x, w["learnable_registers"].float())
    T, D = x.shape[1], x.shape[2]
    pe = rope_1d(T, D, heads, theta, max_pos, rope_type)
    for i in range(layers):
        P = f"transformer_1d_blocks.{i}"
        x = x + attention(w, P + ".attn1", O.rms_norm(x, None, eps), heads, pe, rope_type)
        x = x + O.feed_forward(w, P + ".ff", O.rms_norm(x, None, eps))
    return O.rms_norm(x, None, eps)
