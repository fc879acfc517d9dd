"""Pin the oracle (oracle/*.py) against golden vectors produced by the reference's own
code (tests/golden/make_golden.py).  CPU only."""
import ast
import os

import numpy as np
import pytest
import torch

from ltx2_b200 import synthetic
from oracle import dit_oracle as D
from oracle import vae_oracle as V

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def T(a):
    return torch.from_numpy(np.asarray(a))


def checksum(w):
    return float(sum(float(v.double().abs().sum()) for v in w.values()))


def close(a, b, rtol, atol):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b).max()
    assert np.allclose(a, b, rtol=rtol, atol=atol), f"max abs err {err:.3e}, ref max {np.abs(b).max():.3e}"


V1 = synthetic.DitConfig(num_attention_heads=4, attention_head_dim=32, in_channels=16, out_channels=16,
                         num_layers=2, cross_attention_dim=128, caption_channels=48)
V2 = synthetic.DitConfig(num_attention_heads=4, attention_head_dim=32, in_channels=16, out_channels=16,
                         num_layers=2, cross_attention_dim=128, caption_channels=None,
                         cross_attention_adaln=True, apply_gated_attention=True, audio=True,
                         audio_heads=4, audio_head_dim=16)


def test_ops_conv3d_unpatchify_d2s_masks():
    g = load("ops.npz")
    y = V.conv3d(T(g["conv_x"]), T(g["conv_w"]), T(g["conv_b"]), causal=False)
    close(y, g["conv_y"], 1e-4, 1e-5)
    y = V.conv3d(T(g["conv_x"]), T(g["conv_w"]), T(g["conv_b"]), causal=True)
    close(y, g["conv_y_causal"], 1e-4, 1e-5)
    assert np.array_equal(V.unpatchify(T(g["unpatchify_x"]), 4).numpy(), g["unpatchify_y"])
    assert np.array_equal(V.depth_to_space(T(g["d2s_x"]), 8, (2, 2, 2)).numpy(), g["d2s_y"])
    for k, v in g.items():
        if k.startswith("mask_"):
            _, l, a, b, z = k.split("_")
            close(V.trapezoid_mask_1d(int(l), int(a), int(b), bool(int(z))), v, 0, 1e-6)


def test_rope_tables_and_apply():
    g = load("rope.npz")
    cos, sin = D.rope_tables(T(g["positions"]), 4096, 32, D.MAX_POS)
    # fp32 arguments reach ~1.5e4 rad, where one ulp of the argument is ~1e-3 in cos/sin; numpy's
    # and torch's float32 pow() differ by an ulp or two in the frequency grid -> few e-3 abs.
    close(cos, g["cos"], 0, 1.5e-2)
    close(sin, g["sin"], 0, 1.5e-2)
    # the 2 identity entries sit at the FRONT of head 0 (rope.py:311-326)
    assert torch.all(cos[0, 0, :, :2] == 1) and torch.all(sin[0, 0, :, :2] == 0)
    y = D.apply_split_rope(T(g["x"]), T(g["cos"]), T(g["sin"]))
    close(y, g["y"], 1e-5, 1e-5)
    c1, s1 = D.rope_tables(T(g["positions"])[:, 0:1], 2048, 32, (D.AUDIO_MAX_POS,))
    close(c1, g["cos_1d"], 0, 1.5e-2)
    close(s1, g["sin_1d"], 0, 1.5e-2)


@pytest.mark.parametrize("mode", ["scalar", "pertoken"])
def test_dit_v1(mode):
    g = load("dit_v1.npz")
    w = synthetic.dit_weights(V1, seed=1)
    assert abs(checksum(w) - float(g["weight_checksum"])) < 1e-6 * float(g["weight_checksum"])
    w = D.to_engine_keys(w)
    video = dict(latent=T(g["latent"]), context=T(g["context"]), timesteps=T(g[f"timesteps_{mode}"]),
                 positions=T(g["positions"]))
    vel = D.dit_forward(w, video, num_layers=2, heads=4)
    close(vel, g[f"velocity_{mode}"], 2e-3, 2e-4)
    x0 = D.x0_forward(w, video, num_layers=2, heads=4)
    close(x0, g[f"x0_{mode}"], 2e-3, 2e-4)


def test_dit_v2_audio_video_and_stg():
    g = load("dit_v2_av.npz")
    w = synthetic.dit_weights(V2, seed=2)
    assert abs(checksum(w) - float(g["weight_checksum"])) < 1e-6 * float(g["weight_checksum"])
    w = D.to_engine_keys(w)
    video = dict(latent=T(g["latent"]), context=T(g["context"]), timesteps=T(g["sigma_video"]),
                 positions=T(g["positions"]), sigma=T(g["sigma_video"]))
    audio = dict(latent=T(g["audio_latent"]), context=T(g["audio_context"]), timesteps=T(g["sigma_audio"]),
                 positions=T(g["audio_positions"]), sigma=T(g["sigma_audio"]))
    kw = dict(num_layers=2, heads=4, audio_heads=4, v2=True, av_ca_timestep_scale_multiplier=1000)
    vv, av = D.dit_forward(w, video, audio, **kw)
    close(vv, g["velocity_video"], 2e-3, 2e-4)
    close(av, g["velocity_audio"], 2e-3, 2e-4)
    x0v, x0a = D.x0_forward(w, video, audio, **kw)
    close(x0v, g["x0_video"], 2e-3, 2e-4)
    close(x0a, g["x0_audio"], 2e-3, 2e-4)
    close(D.dit_forward(w, video, None, **kw), g["velocity_video_only"], 2e-3, 2e-4)
    pv, pa = D.dit_forward(w, video, audio, skip_blocks={"video_self": [1], "a2v": [0]}, **kw)
    close(pv, g["velocity_video_stg"], 2e-3, 2e-4)
    close(pa, g["velocity_audio_stg"], 2e-3, 2e-4)
    assert not np.allclose(g["velocity_video_stg"], g["velocity_video"], atol=1e-3)


def test_vae_v20_decode_and_chunked_frames():
    g = load("vae_v20.npz")
    blocks = ast.literal_eval(str(g["decoder_blocks"]))
    cfg = synthetic.VaeConfig(decoder_blocks=blocks, base_channels=8)
    w = synthetic.vae_weights(cfg, seed=3)
    assert abs(checksum(w) - float(g["weight_checksum"])) < 1e-6 * float(g["weight_checksum"])
    kw = dict(decoder_blocks=blocks, base_channels=8)
    close(V.vae_decode(w, T(g["latent"]), timestep=0.05, **kw), g["video"], 2e-3, 2e-4)
    close(V.vae_decode(w, T(g["latent"]), timestep=0.05, causal=True, **kw), g["video_causal"], 2e-3, 2e-4)
    close(V.vae_decode(w, T(g["latent"]), timestep=None, **kw), g["video_no_timestep"], 2e-3, 2e-4)
    f9 = V.decode_latent(w, T(g["latent9"]), **kw).numpy().astype(np.int32)
    d = np.abs(f9 - g["frames9"].astype(np.int32))
    assert f9.shape == g["frames9"].shape == (65, 64, 64, 3)
    assert d.max() <= 1 and (d > 0).mean() < 0.01          # uint8 truncation of ~1e-5 float noise
    f3 = V.decode_latent(w, T(g["latent"])[0], **kw).numpy().astype(np.int32)
    assert np.abs(f3 - g["frames3"].astype(np.int32)).max() <= 1
    assert V.chunk_plan(9) == [(0, 7), (5, 9)] and V.chunk_plan(16) == [(0, 7), (5, 12), (10, 16)]
    assert V.chunk_plan(7) == [(0, 7)]


def test_vae_v23_style_stack():
    g = load("vae_v23.npz")
    blocks = ast.literal_eval(str(g["decoder_blocks"]))
    cfg = synthetic.VaeConfig(decoder_blocks=blocks, base_channels=16, timestep_conditioning=False)
    w = synthetic.vae_weights(cfg, seed=4)
    assert abs(checksum(w) - float(g["weight_checksum"])) < 1e-6 * float(g["weight_checksum"])
    out = V.vae_decode(w, T(g["latent"]), decoder_blocks=blocks, base_channels=16, timestep=None,
                       timestep_conditioning=False)
    assert out.shape == g["video"].shape
    close(out, g["video"], 2e-3, 2e-4)


def test_metal_kernel_formulas():
    # kernels/fused_ops.py:12-47,136-180 cannot execute off-Apple; the oracle restates the shader math.
    a, b = torch.randn(3, 5, 8), torch.randn(3, 5, 8)
    close(D.silu_mul(a, b), (a / (1 + torch.exp(-a))) * b, 1e-6, 1e-6)
    close(D.gelu_mul(a, b), torch.nn.functional.gelu(a, approximate="tanh") * b, 1e-5, 1e-6)
    th = torch.rand(3, 5, 4) * 6.28
    cos, sin = torch.cos(th).repeat_interleave(2, -1), torch.sin(th).repeat_interleave(2, -1)
    y = D.interleaved_rope(a, cos, sin)
    z = torch.view_as_real(torch.view_as_complex(a.reshape(3, 5, 4, 2).contiguous()) * torch.polar(torch.ones_like(th), th))
    close(y, z.reshape(3, 5, 8), 1e-5, 1e-6)


def test_vae_decode_tiled_spatial_and_temporal():
    g = load("vae_tiled.npz")
    blocks = ast.literal_eval(str(g["decoder_blocks"]))
    cfg = synthetic.VaeConfig(decoder_blocks=blocks, base_channels=8)
    w = synthetic.vae_weights(cfg, seed=3)
    assert abs(checksum(w) - float(g["weight_checksum"])) < 1e-6 * float(g["weight_checksum"])
    dec = lambda x: V.vae_decode(w, x, decoder_blocks=blocks, base_channels=8, timestep=0.05)  # noqa: E731
    out = V.decode_tiled(dec, T(g["latent_spatial"]), tile_px=64, overlap_px=32)
    close(out, g["video_spatial"], 2e-3, 2e-4)
    out_t = V.decode_tiled(dec, T(g["latent_temporal"]), tile_px=None, tile_frames=16, overlap_frames=8)
    close(out_t, g["video_temporal"], 2e-3, 2e-4)
