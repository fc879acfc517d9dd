"""Host-side mirror of the reference's text-embeddings connector, backed by the sm_100a kernels.

Same names and call surface as ``LTX_2_MLX/model/text_encoder/connector.py``: ``Embeddings1DConnector`` (:104-283,
``connector(hidden_states (B,T,D), attention_mask) -> (hidden_states (B,T',D), mask)``) -- the 2-block 1-D transformer
that refines the Gemma features into the DiT's text context once per prompt; the only consumer of the reference's
``interleaved_rope`` Metal kernel (rope.py:51-89).  The Gemma encoder itself stays with the caller (out of scope:
SURVEY.md section 2).

Every linear runs on the tcgen05 GEMM with its fused epilogue (bias / GELU-tanh / fp32 residual accumulate), attention on
the tcgen05 flash-attention kernel (head_dim 64 or 128, per-head 2*sigmoid gate), q/k RMSNorm + head split on
``ltx2_headnorm_rope`` and the interleaved RoPE on ``ltx2_interleaved_rope``.  The cos/sin tables are built on the host
with the reference's own float32 operation sequence (they are chaotic in the frequency grid's last bit: angles reach
3e7 rad with max_pos = [1]; see oracle/connector_oracle.py) and uploaded once per sequence length.
"""
from __future__ import annotations

import math
from typing import Any, Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch

from . import ops
from .transformer import LTXRopeType, to_device


class Embeddings1DConnector:
    def __init__(self, attention_head_dim: int = 128, num_attention_heads: int = 30, num_layers: int = 2,
                 positional_embedding_theta: float = 10000.0, positional_embedding_max_pos: Optional[List[int]] = None,
                 num_learnable_registers: Optional[int] = 128, rope_type: Any = LTXRopeType.INTERLEAVED,
                 norm_eps: float = 1e-6, apply_gated_attention: bool = False, double_precision_rope: bool = False,
                 device="cuda"):
        if attention_head_dim not in (64, 128):
            raise ValueError("attention_head_dim must be 64 or 128")
        self.num_attention_heads, self.head_dim = num_attention_heads, attention_head_dim
        self.inner_dim = num_attention_heads * attention_head_dim
        self.num_layers = num_layers
        self.positional_embedding_theta = positional_embedding_theta
        self.positional_embedding_max_pos = positional_embedding_max_pos or [1]
        self.rope_type = LTXRopeType[getattr(rope_type, "name", str(rope_type)).upper()]
        self.norm_eps = norm_eps
        self.apply_gated_attention = apply_gated_attention
        self.double_precision_rope = double_precision_rope
        self.num_learnable_registers = num_learnable_registers
        self.device = torch.device(device)
        self._w: Dict[str, torch.Tensor] = {}
        self._tables: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}

    # ---- weights: the reference's module attribute names ("transformer_1d_blocks.0.attn1.to_q.weight", ...) ----------
    def expected_keys(self) -> List[str]:
        keys = []
        for i in range(self.num_layers):
            P = f"transformer_1d_blocks.{i}"
            for lin in ("attn1.to_q", "attn1.to_k", "attn1.to_v", "attn1.to_out", "ff.project_in.proj", "ff.project_out"):
                keys += [f"{P}.{lin}.weight", f"{P}.{lin}.bias"]
            keys += [f"{P}.attn1.q_norm.weight", f"{P}.attn1.k_norm.weight"]
            if self.apply_gated_attention:
                keys += [f"{P}.attn1.to_gate_logits.weight", f"{P}.attn1.to_gate_logits.bias"]
        if self.num_learnable_registers:
            keys.append("learnable_registers")
        return keys

    def load_weights(self, weights: Iterable[Tuple[str, Any]]) -> int:
        want, n = set(self.expected_keys()), 0
        for key, value in (weights.items() if isinstance(weights, dict) else weights):
            for prefix in ("model.diffusion_model.video_embeddings_connector.", "model.diffusion_model.embeddings_connector.",
                           "embeddings_connector."):
                if key.startswith(prefix):
                    key = key[len(prefix):]
            key = key.replace(".to_out.0.", ".to_out.").replace(".ff.net.0.proj.", ".ff.project_in.proj.") \
                     .replace(".ff.net.2.", ".ff.project_out.")
            if key not in want:
                continue
            t = to_device(value, self.device)
            is_matrix = t.ndim == 2 and key != "learnable_registers"
            self._w[key] = t.to(torch.bfloat16).contiguous() if is_matrix else t.float().contiguous()
            n += 1
        if all(k in self._w for k in self.expected_keys()):
            self._pack()
        return n

    def missing_weights(self) -> List[str]:
        return [k for k in self.expected_keys() if k not in self._w]

    def _pack(self) -> None:
        """q/k/v adjacent (one GEMM, N = 3*inner); gate projection padded to the GEMM's 32-column granule."""
        for i in range(self.num_layers):
            P = f"transformer_1d_blocks.{i}.attn1"
            self._w[P + ".qkv.weight"] = torch.cat([self._w[f"{P}.to_{n}.weight"] for n in "qkv"], dim=0).contiguous()
            self._w[P + ".qkv.bias"] = torch.cat([self._w[f"{P}.to_{n}.bias"] for n in "qkv"], dim=0).contiguous()
            if self.apply_gated_attention:
                H, D = self.num_attention_heads, self.inner_dim
                Hp = (H + 31) // 32 * 32
                gw = torch.zeros(Hp, D, device=self.device, dtype=torch.bfloat16)
                gb = torch.zeros(Hp, device=self.device, dtype=torch.float32)
                gw[:H], gb[:H] = self._w[P + ".to_gate_logits.weight"], self._w[P + ".to_gate_logits.bias"]
                self._w[P + ".gate.weight"], self._w[P + ".gate.bias"] = gw, gb

    # ---- RoPE tables (rope.py:365-418 with indices_grid = arange(T)[None, None, :]) ------------------------------------
    def _rope(self, T: int):
        if T not in self._tables:
            D, H, Dh = self.inner_dim, self.num_attention_heads, self.head_dim
            n, theta = D // 2, self.positional_embedding_theta
            if self.double_precision_rope:                    # generate_freq_grid_np, rope.py:147-178
                idx = (np.power(theta, np.linspace(0.0, 1.0, n, dtype=np.float64)) * math.pi / 2).astype(np.float32)
            else:                                             # generate_freq_grid, rope.py:181-211 (float32 throughout)
                idx = ((np.float32(theta) ** np.linspace(0.0, 1.0, n).astype(np.float32)) *
                       np.float32(math.pi / 2)).astype(np.float32)
            scaled = (np.arange(T, dtype=np.float32) / np.float32(self.positional_embedding_max_pos[0])) * np.float32(2) \
                - np.float32(1)
            freqs = (idx[None, :] * scaled[:, None]).astype(np.float32)                   # (T, D/2)
            cos, sin = torch.from_numpy(np.cos(freqs)), torch.from_numpy(np.sin(freqs))
            if self.rope_type == LTXRopeType.SPLIT:           # token-major (T, D/2) fp32, as ltx2_headnorm_rope reads it
                tab = (cos.to(self.device).contiguous(), sin.to(self.device).contiguous())
            else:                                             # each value twice (rope.py:331-362), head-split (H, T, Dh)
                def lay(t):
                    t = t.repeat_interleave(2, dim=-1).reshape(T, H, Dh).permute(1, 0, 2)
                    return t.to(self.device, torch.bfloat16).contiguous()
                tab = (lay(cos), lay(sin))
            self._tables[T] = tab
        return self._tables[T]

    def _append_learnable_registers(self, x: torch.Tensor) -> torch.Tensor:
        """connector.py:175-228: registers tiled to max(1024, T) (rounded up), rows [T:] appended; the mask is cleared."""
        B, T, D = x.shape
        n = self.num_learnable_registers
        dup = math.ceil(max(1024, T) / n)
        extra = self._w["learnable_registers"].repeat(dup, 1)[T:]
        if extra.shape[0] > 0:
            x = torch.cat([x, extra[None].expand(B, -1, -1).to(x.dtype)], dim=1)
        return x

    # ---- forward --------------------------------------------------------------------------------------------------------
    def __call__(self, hidden_states, attention_mask=None):
        miss = self.missing_weights()
        if miss:
            raise RuntimeError(f"connector weights missing: {miss[:4]}")
        x = to_device(hidden_states, self.device, torch.float32)
        if x.ndim == 4:
            x = x.squeeze(1)
        if x.ndim != 3 or x.shape[2] != self.inner_dim:
            raise ValueError(f"hidden_states must be (B, T, {self.inner_dim}); got {tuple(x.shape)}")
        if self.num_learnable_registers:
            x = self._append_learnable_registers(x)
        elif attention_mask is not None:
            raise NotImplementedError("a connector without learnable registers keeps the padding mask; the reference "
                                      "pipelines always use registers (connector.py:117)")
        B, T, D = x.shape
        H, Dh, M = self.num_attention_heads, self.head_dim, B * T
        cos, sin = self._rope(T)
        split = self.rope_type == LTXRopeType.SPLIT
        with torch.cuda.device(self.device):
            x = x.reshape(M, D).contiguous()                                  # fp32 residual stream
            w = self._w
            for i in range(self.num_layers):
                P = f"transformer_1d_blocks.{i}"
                xn = ops.norm_modulate(x, kind=ops.NORM_RMS, eps=self.norm_eps)
                qkv = ops.gemm(xn, w[P + ".attn1.qkv.weight"], w[P + ".attn1.qkv.bias"])          # [M, 3D] bf16
                if split:
                    c = cos[None].expand(B, -1, -1).contiguous()
                    s = sin[None].expand(B, -1, -1).contiguous()
                    qh = ops.headnorm_rope(qkv[:, :D], w[P + ".attn1.q_norm.weight"], B, T, H, Dh, c, s, self.norm_eps)
                    kh = ops.headnorm_rope(qkv[:, D:2 * D], w[P + ".attn1.k_norm.weight"], B, T, H, Dh, c, s, self.norm_eps)
                else:
                    qh = ops.headnorm_rope(qkv[:, :D], w[P + ".attn1.q_norm.weight"], B, T, H, Dh, eps=self.norm_eps)
                    kh = ops.headnorm_rope(qkv[:, D:2 * D], w[P + ".attn1.k_norm.weight"], B, T, H, Dh, eps=self.norm_eps)
                    c = cos[None].expand(B, -1, -1, -1).contiguous()
                    s = sin[None].expand(B, -1, -1, -1).contiguous()
                    qh, kh = ops.interleaved_rope(qh, c, s), ops.interleaved_rope(kh, c, s)
                gate = None
                if self.apply_gated_attention:
                    gl = ops.gemm(xn, w[P + ".attn1.gate.weight"], w[P + ".attn1.gate.bias"], mode=ops.EPI_F32)
                    gate = gl[:, :H].contiguous()
                attn = ops.attention_vrows(qh, kh, qkv.view(B, T, 3 * D)[:, :, 2 * D:], H, Dh, gate_logits=gate)
                ops.gemm(attn.view(M, D), w[P + ".attn1.to_out.weight"], w[P + ".attn1.to_out.bias"],
                         mode=ops.EPI_F32_RESIDUAL, out=x)
                xn = ops.norm_modulate(x, kind=ops.NORM_RMS, eps=self.norm_eps)
                h = ops.gemm(xn, w[P + ".ff.project_in.proj.weight"], w[P + ".ff.project_in.proj.bias"],
                             mode=ops.EPI_BF16_GELU)
                ops.gemm(h, w[P + ".ff.project_out.weight"], w[P + ".ff.project_out.bias"], mode=ops.EPI_F32_RESIDUAL, out=x)
            out = ops.norm_modulate(x, kind=ops.NORM_RMS, eps=self.norm_eps).float().view(B, T, D)
        mask = torch.zeros(B, 1, 1, T, device=self.device)
        return out, mask
