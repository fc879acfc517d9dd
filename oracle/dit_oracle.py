"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference DiT forward.

This is the *oracle* for the hot path named by BASELINE.json:north_star: a plain
torch-CPU (fp32 or fp64) restatement of what the reference computes, written from
the cited files.  It is imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product path never calls it.

Pinning: the reference's arithmetic lives in the third-party `mlx==0.30.1`
(uv.lock:537-538), which is absent here, and the reference ships no golden vectors
for this path (tests/test_parity.py needs a 43 GB checkpoint + un-vendored PyTorch
fixtures).  The oracle is pinned against golden vectors produced by running the
reference's OWN Python modules (LTX_2_MLX/model/transformer/*.py, imported from
/root/reference) over a numpy restatement of the mlx primitives
(oracle/_mlx_shim, generator tests/golden/make_golden.py).  That pins op order,
layouts, table-row order and constants to the reference's code; the primitive
arithmetic itself (matmul/softmax/rms_norm) is "published semantics", not mlx's
binary.  DESIGN.md states this as "pinned to reference code over restated mlx
primitives".

Reference map (all under /root/reference/LTX_2_MLX/model/transformer/):
  timestep embedding ........ timestep_embedding.py:10-60, 166-202
  RoPE tables (SPLIT) ....... rope.py:181-211, 214-289, 292-328, 365-418
  RoPE apply (SPLIT) ........ rope.py:92-144
  attention ................. attention.py:203-253
  FFN ....................... feed_forward.py:18-54
  block ..................... transformer.py:369-455, 457-648
  preprocess / head / X0 .... model.py:113-161, 203-281, 320-410, 744-758, 776-881, 895-936
"""
from __future__ import annotations

import math
import re
from typing import Dict, Optional, Sequence, Tuple

import torch

EPS = 1e-6            # model.py:445, attention.py:160
THETA = 10000.0       # model.py:447
MAX_POS = (20, 2048, 2048)  # model.py:507-508
AUDIO_MAX_POS = 20    # model.py:434
TS_MULT = 1000.0      # model.py:449


# ---------------------------------------------------------------------------------
# key handling (weight_converter.py:277-313, re-expressed)
# ---------------------------------------------------------------------------------
_RENAMES = (
    (re.compile(r"\.to_out\.0\."), ".to_out."),
    (re.compile(r"\.(audio_)?ff\.net\.0\.proj\."), r".\1ff.project_in.proj."),
    (re.compile(r"\.(audio_)?ff\.net\.2\."), r".\1ff.project_out."),
)


def engine_key(checkpoint_key: str) -> Optional[str]:
    """'model.diffusion_model.X' -> the reference's MLX attribute path, None if not a DiT key."""
    p = "model.diffusion_model."
    if not checkpoint_key.startswith(p):
        return None
    k = checkpoint_key[len(p):]
    for rx, rep in _RENAMES:
        k = rx.sub(rep, k)
    return k


def to_engine_keys(weights: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in weights.items():
        ek = engine_key(k)
        if ek is not None:
            out[ek] = v
    return out


# ---------------------------------------------------------------------------------
# primitives
# ---------------------------------------------------------------------------------

def rms_norm(x, weight=None, eps=EPS):
    y = x * torch.rsqrt(torch.mean(x * x, dim=-1, keepdim=True) + eps)
    return y if weight is None else y * weight


def gelu_tanh(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x * x * x)))


def linear(w, prefix, x):
    y = x @ w[prefix + ".weight"].to(x.dtype).T
    b = w.get(prefix + ".bias")
    return y if b is None else y + b.to(x.dtype)


def sinusoid_dit(t: torch.Tensor, dim: int = 256) -> torch.Tensor:
    """get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0): [cos, sin]."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half
    arg = t.reshape(-1, 1).to(torch.float32) * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)


def adaln_single(w, prefix, t_flat, dtype):
    """AdaLayerNormSingle: (Linear(SiLU(e)), e) with e = MLP(sinusoid(t))."""
    s = sinusoid_dit(t_flat).to(dtype)
    e = linear(w, prefix + ".emb.timestep_embedder.linear_1", s)
    e = torch.nn.functional.silu(e)
    e = linear(w, prefix + ".emb.timestep_embedder.linear_2", e)
    return linear(w, prefix + ".linear", torch.nn.functional.silu(e)), e


def freq_grid(theta: float, n_pos_dims: int, dim: int) -> torch.Tensor:
    """generate_freq_grid (rope.py:181-211): theta**linspace(0,1,dim//(2*n)) * pi/2, fp32."""
    n = dim // (2 * n_pos_dims)
    lin = torch.linspace(0.0, 1.0, n, dtype=torch.float32)
    return (torch.pow(torch.tensor(theta, dtype=torch.float32), lin) * (math.pi / 2)).to(torch.float32)


def rope_tables(positions: torch.Tensor, dim: int, heads: int, max_pos: Sequence[int],
                theta: float = THETA) -> Tuple[torch.Tensor, torch.Tensor]:
    """precompute_freqs_cis(rope_type=SPLIT, use_middle_indices_grid=True).

    positions (B, n_dims, T, 2) -> cos, sin each (B, H, T, dim/(2H)) fp32.
    """
    pos = positions.to(torch.float32)
    B, n_dims, T, _ = pos.shape
    assert n_dims == len(max_pos)
    mid = (pos[..., 0] + pos[..., 1]) / 2.0                        # (B, n_dims, T)
    frac = torch.stack([mid[:, i, :] / max_pos[i] for i in range(n_dims)], dim=-1)  # (B,T,n_dims)
    scaled = frac * 2 - 1
    idx = freq_grid(theta, n_dims, dim)                            # (n_freq,)
    freqs = idx[None, None, None, :] * scaled[..., None]           # (B,T,n_dims,n_freq)
    freqs = freqs.permute(0, 1, 3, 2).reshape(B, T, -1)            # freq-major, axis-minor
    cos, sin = torch.cos(freqs), torch.sin(freqs)
    pad = dim // 2 - freqs.shape[-1]
    if pad:
        cos = torch.cat([torch.ones(B, T, pad), cos], dim=-1)      # identity entries in FRONT
        sin = torch.cat([torch.zeros(B, T, pad), sin], dim=-1)
    cos = cos.reshape(B, T, heads, -1).permute(0, 2, 1, 3)
    sin = sin.reshape(B, T, heads, -1).permute(0, 2, 1, 3)
    return cos.contiguous(), sin.contiguous()


def apply_split_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """apply_split_rotary_emb: x (B,T,H*Dh), cos/sin (B,H,T,Dh/2)."""
    B, H, T, half = cos.shape
    xh = x.reshape(B, T, H, 2, half).permute(0, 2, 1, 3, 4)        # (B,H,T,2,half)
    x1, x2 = xh[..., 0, :], xh[..., 1, :]
    c, s = cos.to(x.dtype), sin.to(x.dtype)
    o = torch.stack([x1 * c - x2 * s, x2 * c + x1 * s], dim=-2)    # (B,H,T,2,half)
    return o.permute(0, 2, 1, 3, 4).reshape(B, T, H * 2 * half)


def sdpa(q, k, v, heads):
    B, Tq, inner = q.shape
    Tk = k.shape[1]
    d = inner // heads
    qh = q.reshape(B, Tq, heads, d).permute(0, 2, 1, 3)
    kh = k.reshape(B, Tk, heads, d).permute(0, 2, 1, 3)
    vh = v.reshape(B, Tk, heads, d).permute(0, 2, 1, 3)
    scale = 1.0 / math.sqrt(d)
    # same arithmetic per query row; query blocks only bound the (B,H,Tq,Tk) score matrix (19 GB at N = 12288)
    step = Tq if B * heads * Tq * Tk <= (1 << 29) else max(1, (1 << 29) // (B * heads * Tk))
    outs = []
    for q0 in range(0, Tq, step):
        s = (qh[:, :, q0:q0 + step] @ kh.transpose(-1, -2)) * scale
        outs.append(torch.softmax(s, dim=-1) @ vh)
    o = outs[0] if len(outs) == 1 else torch.cat(outs, dim=2)
    return o.permute(0, 2, 1, 3).reshape(B, Tq, inner)


def attention(w, prefix, x, heads, context=None, pe=None, k_pe=None):
    """Attention.__call__ (attention.py:203-253): proj -> full-width RMSNorm(weight) -> RoPE
    -> SDPA -> optional 2*sigmoid per-head gate -> out proj."""
    ctx = x if context is None else context
    q = linear(w, prefix + ".to_q", x)
    k = linear(w, prefix + ".to_k", ctx)
    v = linear(w, prefix + ".to_v", ctx)
    q = rms_norm(q, w[prefix + ".q_norm.weight"].to(x.dtype))
    k = rms_norm(k, w[prefix + ".k_norm.weight"].to(x.dtype))
    if pe is not None:
        q = apply_split_rope(q, *pe)
        k = apply_split_rope(k, *(pe if k_pe is None else k_pe))
    out = sdpa(q, k, v, heads)
    if prefix + ".to_gate_logits.weight" in w:
        g = 2.0 * torch.sigmoid(linear(w, prefix + ".to_gate_logits", x))    # (B,T,H)
        B, T, inner = out.shape
        out = (out.reshape(B, T, heads, inner // heads) * g[..., None]).reshape(B, T, inner)
    return linear(w, prefix + ".to_out", out)


def feed_forward(w, prefix, x):
    return linear(w, prefix + ".project_out", gelu_tanh(linear(w, prefix + ".project_in.proj", x)))


# ---------------------------------------------------------------------------------
# preprocessing (model.py:231-281, 368-410)
# ---------------------------------------------------------------------------------

def _prepare_timestep(w, prefix, t, batch, dim, dtype):
    emb, e = adaln_single(w, prefix, (t.to(torch.float32) * TS_MULT).flatten(), dtype)
    n = emb.shape[-1] // dim
    return emb.reshape(batch, -1, n, dim), e.reshape(batch, -1, dim)


def _scalar_sigma(mod):
    s = mod.get("sigma")
    if s is None:
        s = mod["timesteps"]
    if s.ndim > 1:
        s = s[:, 0]
    return s


def prepare(w, mod: dict, *, audio_side: bool, dim: int, heads: int, max_pos, v2: bool, dtype,
            cross: Optional[dict] = None, cross_dim: Optional[int] = None,
            av_ca_timestep_scale_multiplier: float = 1.0):
    pfx = "audio_" if audio_side else ""
    x = linear(w, pfx + "patchify_proj", mod["latent"].to(dtype))
    B = x.shape[0]
    ts, emb_t = _prepare_timestep(w, pfx + "adaln_single", mod["timesteps"], B, dim, dtype)
    prompt_ts = None
    if v2:
        prompt_ts, _ = _prepare_timestep(w, pfx + "prompt_adaln_single", _scalar_sigma(mod), B, dim, dtype)
    ctx = mod["context"].to(dtype)
    if (pfx + "caption_projection.linear_1.weight") in w:
        ctx = linear(w, pfx + "caption_projection.linear_2",
                     gelu_tanh(linear(w, pfx + "caption_projection.linear_1", ctx)))
    ctx = ctx.reshape(B, -1, dim)
    pe = rope_tables(mod["positions"], dim, heads, max_pos)
    args = dict(x=x, context=ctx, timesteps=ts, embedded_timestep=emb_t, pe=pe, prompt_timestep=prompt_ts,
                cross_pe=None, cross_ss=None, cross_gate=None)
    if cross is not None:
        # cross-modal RoPE uses THIS modality's temporal axis only; timestep uses the OTHER's sigma
        args["cross_pe"] = rope_tables(mod["positions"][:, 0:1], cross_dim, heads, (AUDIO_MAX_POS,))
        st = _scalar_sigma(cross).to(torch.float32) * TS_MULT
        side = "audio" if audio_side else "video"
        gate = "v2a" if audio_side else "a2v"
        ss, _ = adaln_single(w, f"av_ca_{side}_scale_shift_adaln_single", st.flatten(), dtype)
        args["cross_ss"] = ss.reshape(B, -1, 4, dim)
        factor = av_ca_timestep_scale_multiplier / TS_MULT
        g, _ = adaln_single(w, f"av_ca_{gate}_gate_adaln_single", (st * factor).flatten(), dtype)
        args["cross_gate"] = g.reshape(B, -1, 1, dim)
    return args


# ---------------------------------------------------------------------------------
# block (transformer.py:457-648)
# ---------------------------------------------------------------------------------

def _ada(table, ts, start, end):
    v = table[start:end].to(ts.dtype)[None, None] + ts[:, :, start:end, :]
    return tuple(v[:, :, i, :] for i in range(end - start))


def _text_cross_attention(w, B, attn_prefix, x, args, table, prompt_table, heads, v2):
    if v2:
        shift_q, scale_q, gate = _ada(table, args["timesteps"], 6, 9)
        kv = prompt_table.to(x.dtype)[None, None] + args["prompt_timestep"]
        shift_kv, scale_kv = kv[:, :, 0, :], kv[:, :, 1, :]
        a_in = rms_norm(x) * (1 + scale_q) + shift_q
        ctx = args["context"] * (1 + scale_kv) + shift_kv
        return attention(w, attn_prefix, a_in, heads, context=ctx) * gate
    return attention(w, attn_prefix, rms_norm(x), heads, context=args["context"])


def block(w, i: int, video: Optional[dict], audio: Optional[dict], *, heads: int, audio_heads: int,
          v2: bool, skip: Optional[dict] = None, cross_attn_scale: Optional[float] = None):
    B_ = f"transformer_blocks.{i}"
    skip = skip or {}
    vx = video["x"] if video is not None else None
    ax = audio["x"] if audio is not None else None
    run_vx = vx is not None and vx.numel() > 0
    run_ax = ax is not None and ax.numel() > 0
    run_a2v = run_vx and run_ax
    run_v2a = run_ax and run_vx

    if run_vx:
        tab = w[B_ + ".scale_shift_table"]
        shift, scale, gate = _ada(tab, video["timesteps"], 0, 3)
        if not skip.get("video_self"):
            n = rms_norm(vx) * (1 + scale) + shift
            vx = vx + attention(w, B_ + ".attn1", n, heads, pe=video["pe"]) * gate
        c = _text_cross_attention(w, B_, B_ + ".attn2", vx, video, tab,
                                  w.get(B_ + ".prompt_scale_shift_table"), heads, v2)
        if cross_attn_scale is not None:
            c = c * cross_attn_scale
        vx = vx + c

    if run_ax:
        atab = w[B_ + ".audio_scale_shift_table"]
        shift, scale, gate = _ada(atab, audio["timesteps"], 0, 3)
        if not skip.get("audio_self"):
            n = rms_norm(ax) * (1 + scale) + shift
            ax = ax + attention(w, B_ + ".audio_attn1", n, audio_heads, pe=audio["pe"]) * gate
        ax = ax + _text_cross_attention(w, B_, B_ + ".audio_attn2", ax, audio, atab,
                                        w.get(B_ + ".audio_prompt_scale_shift_table"), audio_heads, v2)

    if run_a2v or run_v2a:
        vn, an = rms_norm(vx), rms_norm(ax)

        def av(table, ss, g):
            t = table.to(ss.dtype)
            s4 = t[:4][None, None] + ss
            gg = t[4:][None, None] + g
            return tuple(s4[:, :, j, :] for j in range(4)) + (gg[:, :, 0, :],)

        # rows are (scale_a2v, shift_a2v, scale_v2a, shift_v2a, gate)  -- scale FIRST
        sa_a2v, ha_a2v, sa_v2a, ha_v2a, gate_v2a = av(w[B_ + ".scale_shift_table_a2v_ca_audio"],
                                                       audio["cross_ss"], audio["cross_gate"])
        sv_a2v, hv_a2v, sv_v2a, hv_v2a, gate_a2v = av(w[B_ + ".scale_shift_table_a2v_ca_video"],
                                                       video["cross_ss"], video["cross_gate"])
        if run_a2v and not skip.get("a2v"):
            vq = vn * (1 + sv_a2v) + hv_a2v
            ak = an * (1 + sa_a2v) + ha_a2v
            vx = vx + attention(w, B_ + ".audio_to_video_attn", vq, audio_heads, context=ak,
                                pe=video["cross_pe"], k_pe=audio["cross_pe"]) * gate_a2v
        if run_v2a and not skip.get("v2a"):
            aq = an * (1 + sa_v2a) + ha_v2a
            vk = vn * (1 + sv_v2a) + hv_v2a
            ax = ax + attention(w, B_ + ".video_to_audio_attn", aq, audio_heads, context=vk,
                                pe=audio["cross_pe"], k_pe=video["cross_pe"]) * gate_v2a

    if run_vx:
        shift, scale, gate = _ada(w[B_ + ".scale_shift_table"], video["timesteps"], 3, 6)
        n = rms_norm(vx) * (1 + scale) + shift
        vx = vx + feed_forward(w, B_ + ".ff", n) * gate
    if run_ax:
        shift, scale, gate = _ada(w[B_ + ".audio_scale_shift_table"], audio["timesteps"], 3, 6)
        n = rms_norm(ax) * (1 + scale) + shift
        ax = ax + feed_forward(w, B_ + ".audio_ff", n) * gate

    if video is not None:
        video = dict(video, x=vx)
    if audio is not None:
        audio = dict(audio, x=ax)
    return video, audio


def _output(w, pfx, x, emb_t):
    ss = w[pfx + "scale_shift_table"].to(x.dtype)[None, None] + emb_t[:, :, None, :]
    shift, scale = ss[:, :, 0, :], ss[:, :, 1, :]
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    xn = (x - mu) * torch.rsqrt(var + EPS)
    return linear(w, pfx + "proj_out", xn * (1 + scale) + shift)


# ---------------------------------------------------------------------------------
# model
# ---------------------------------------------------------------------------------

def dit_forward(w: Dict[str, torch.Tensor], video: dict, audio: Optional[dict] = None, *,
                num_layers: int, heads: int = 32, audio_heads: int = 32,
                v2: bool = False, skip_blocks: Optional[Dict[str, Sequence[int]]] = None,
                av_ca_timestep_scale_multiplier: float = 1.0,
                dtype=torch.float32, return_hidden: bool = False):
    """LTXModel.__call__ (model.py:776-881). `w` uses engine keys (see to_engine_keys).

    video/audio: dict(latent (B,N,C), context (B,S,Cc), timesteps (B,)|(B,N)|(B,N,1),
    positions (B,n_dims,N,2), sigma optional (B,)).  Returns velocity (B,N,C_out) fp32
    [and audio velocity when audio is given].
    skip_blocks: {"video_self"|"audio_self"|"a2v"|"v2a": [block indices]} (STG perturbations).
    """
    dim = w["patchify_proj.weight"].shape[0]
    has_audio = audio is not None and audio["latent"].numel() > 0
    v_args = prepare(w, video, audio_side=False, dim=dim, heads=heads, max_pos=MAX_POS, v2=v2, dtype=dtype,
                     cross=audio if has_audio else None,
                     cross_dim=w["audio_patchify_proj.weight"].shape[0] if has_audio else None,
                     av_ca_timestep_scale_multiplier=av_ca_timestep_scale_multiplier)
    a_args = None
    if has_audio:
        adim = w["audio_patchify_proj.weight"].shape[0]
        a_args = prepare(w, audio, audio_side=True, dim=adim, heads=audio_heads, max_pos=(AUDIO_MAX_POS,),
                         v2=v2, dtype=dtype, cross=video, cross_dim=adim,
                         av_ca_timestep_scale_multiplier=av_ca_timestep_scale_multiplier)
    skip_blocks = skip_blocks or {}
    for i in range(num_layers):
        skip = {k: (i in v) for k, v in skip_blocks.items()}
        v_args, a_args = block(w, i, v_args, a_args, heads=heads, audio_heads=audio_heads, v2=v2, skip=skip)
    if return_hidden:
        return v_args["x"], (a_args["x"] if a_args is not None else None)
    v_out = _output(w, "", v_args["x"], v_args["embedded_timestep"]).to(torch.float32)
    if a_args is None:
        return v_out
    a_out = _output(w, "audio_", a_args["x"], a_args["embedded_timestep"]).to(torch.float32)
    return v_out, a_out


def to_x0(latent: torch.Tensor, timesteps: torch.Tensor, velocity: torch.Tensor) -> torch.Tensor:
    """X0Model.denoise (model.py:912-918): latent - t * v with t broadcast per token."""
    t = timesteps.to(torch.float32)
    if t.ndim == 1:
        t = t[:, None, None]
    elif t.ndim == 2:
        t = t[:, :, None]
    return latent.to(torch.float32) - t * velocity


def x0_forward(w, video: dict, audio: Optional[dict] = None, **kw):
    out = dit_forward(w, video, audio, **kw)
    if isinstance(out, tuple):
        return to_x0(video["latent"], video["timesteps"], out[0]), to_x0(audio["latent"], audio["timesteps"], out[1])
    return to_x0(video["latent"], video["timesteps"], out)


# ---------------------------------------------------------------------------------
# the three Metal kernels (kernels/fused_ops.py:12-26, 30-47, 136-180)
# ---------------------------------------------------------------------------------

def silu_mul(a, b):
    return a * torch.sigmoid(a) * b


def gelu_mul(a, b):
    return gelu_tanh(a) * b


def interleaved_rope(x, cos, sin):
    """pairs (x[2i], x[2i+1]); cos/sin already broadcast to x's shape."""
    xe, xo = x[..., 0::2], x[..., 1::2]
    ce, se = cos[..., 0::2], sin[..., 0::2]
    co, so = cos[..., 1::2], sin[..., 1::2]
    out = torch.empty_like(x)
    out[..., 0::2] = xe * ce - xo * se
    out[..., 1::2] = xo * co + xe * so
    return out
