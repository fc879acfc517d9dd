#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by executing the REFERENCE's own code.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_golden.py

What runs: the reference's modules are imported from /root/reference unchanged
(LTX_2_MLX/model/transformer/*.py, model/video_vae/{simple_decoder,ops}.py,
components/{patchifiers,perturbations}.py, types.py).  Their only dependency that is
missing here, the third-party `mlx` package, is provided by the numpy restatement in
oracle/_mlx_shim (see its docstring).  Weights are the seeded synthetic checkpoint of
ltx-2-mlx_b200/synthetic.py; the VAE weights go through the reference's own
`load_vae_decoder_weights` via a temporary safetensors file.

Two known defects of the reference at HEAD are worked around, not fixed:
  * LTXModel.__call__ calls prepare(video, audio) on a preprocessor whose signature is
    prepare(modality) for VideoOnly (model.py:825 vs :231) -> wrapped to ignore `audio`.
  * nothing else.

Each .npz stores inputs, outputs and a weight checksum; tests regenerate the weights
from the seed and verify the checksum before comparing.
"""
from __future__ import annotations

import importlib
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

sys.path.insert(0, os.path.join(ROOT, "oracle", "_mlx_shim"))
sys.path.insert(0, ROOT)


def _stub_package(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m


# skip package __init__ files that pull in unrelated subsystems (encoders, guiders, ...)
sys.path.insert(0, REF)
_stub_package("LTX_2_MLX.model.video_vae", f"{REF}/LTX_2_MLX/model/video_vae")
_stub_package("LTX_2_MLX.components", f"{REF}/LTX_2_MLX/components")
import LTX_2_MLX  # noqa: E402  (root __init__: types + core_utils only)
sys.modules["LTX_2_MLX.model.video_vae"].__package__ = "LTX_2_MLX.model.video_vae"

import mlx.core as mx  # noqa: E402  (the shim)
import torch  # noqa: E402

from LTX_2_MLX.model.transformer import model as ref_model  # noqa: E402
from LTX_2_MLX.model.transformer import rope as ref_rope  # noqa: E402
from LTX_2_MLX.components import patchifiers as ref_patch  # noqa: E402
from LTX_2_MLX.components import perturbations as ref_pert  # noqa: E402
from LTX_2_MLX.types import VideoLatentShape, SpatioTemporalScaleFactors  # noqa: E402
ref_vae = importlib.import_module("LTX_2_MLX.model.video_vae.simple_decoder")
ref_ops = importlib.import_module("LTX_2_MLX.model.video_vae.ops")
ref_tiling = importlib.import_module("LTX_2_MLX.model.video_vae.tiling")

synthetic = importlib.import_module("ltx2_b200.synthetic")
from oracle.dit_oracle import engine_key  # noqa: E402  (key renames only)


def A(t):
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().numpy()
    return np.asarray(t)


def set_by_key(obj, key, value):
    parts = key.split(".")
    for p in parts[:-1]:
        obj = obj[int(p)] if p.isdigit() else getattr(obj, p)
    assert hasattr(obj, parts[-1]) or True
    setattr(obj, parts[-1], mx.array(A(value).astype(np.float32)))


def checksum(weights) -> float:
    return float(sum(float(v.double().abs().sum()) for v in weights.values()))


def load_dit(model, cfg, seed):
    w = synthetic.dit_weights(cfg, seed=seed)
    for ck, v in w.items():
        set_by_key(model, engine_key(ck), v)
    return checksum(w)


def ref_positions(batch, f, h, w, fps):
    patchifier = ref_patch.VideoLatentPatchifier(patch_size=1)
    shape = VideoLatentShape(batch=batch, channels=128, frames=f, height=h, width=w)
    coords = patchifier.get_patch_grid_bounds(output_shape=shape)
    px = ref_patch.get_pixel_coords(coords, SpatioTemporalScaleFactors.default(), causal_fix=True)
    px = px.astype(mx.float32)
    if fps is not None:
        # conditioning/tools.py:69-78: temporal axis in seconds
        px = mx.concatenate([px[:, 0:1] / fps, px[:, 1:]], axis=1)
    return px


def rnd(shape, seed, std=1.0):
    return (np.random.default_rng(seed).standard_normal(shape) * std).astype(np.float32)


# ---------------------------------------------------------------------------------
def golden_dit_v1():
    cfg = synthetic.DitConfig(num_attention_heads=4, attention_head_dim=32, in_channels=16, out_channels=16,
                              num_layers=2, cross_attention_dim=128, caption_channels=48)
    m = ref_model.LTXModel(model_type=ref_model.LTXModelType.VideoOnly, num_attention_heads=4,
                           attention_head_dim=32, in_channels=16, out_channels=16, num_layers=2,
                           cross_attention_dim=128, caption_channels=48)
    orig = m._video_args_preprocessor.prepare
    m._video_args_preprocessor.prepare = lambda v, a=None: orig(v)      # model.py:825 defect
    cs = load_dit(m, cfg, seed=1)
    B, F_, H, W, S = 2, 3, 4, 5, 24
    N = F_ * H * W
    lat, ctx = rnd((B, N, 16), 10), rnd((B, S, 48), 11, 0.5)
    pos = ref_positions(B, F_, H, W, 24.0)
    out = {}
    for name, ts in (("scalar", np.array([0.9, 0.4], np.float32)),
                     ("pertoken", np.where(np.arange(N)[None, :, None] < H * W, 0.0,
                                           np.array([0.725, 0.25], np.float32)[:, None, None]).astype(np.float32))):
        mod = ref_model.Modality(latent=mx.array(lat), context=mx.array(ctx), context_mask=None,
                                 timesteps=mx.array(ts), positions=pos)
        vel = m(mod)
        x0 = ref_model.X0Model(m)(mod)
        out[f"timesteps_{name}"] = ts
        out[f"velocity_{name}"] = A(vel)
        out[f"x0_{name}"] = A(x0)
    np.savez_compressed(os.path.join(HERE, "dit_v1.npz"), latent=lat, context=ctx, positions=A(pos),
                        weight_checksum=cs, **out)
    print("dit_v1", {k: v.shape for k, v in out.items()})


def golden_dit_v2_av():
    class SmallAV(ref_model.LTXModel):
        AUDIO_ATTENTION_HEADS = 4
        AUDIO_HEAD_DIM = 16

    cfg = synthetic.DitConfig(num_attention_heads=4, attention_head_dim=32, in_channels=16, out_channels=16,
                              num_layers=2, cross_attention_dim=128, caption_channels=None,
                              cross_attention_adaln=True, apply_gated_attention=True, audio=True,
                              audio_heads=4, audio_head_dim=16, audio_in_channels=128, audio_out_channels=128)
    m = SmallAV(model_type=ref_model.LTXModelType.AudioVideo, num_attention_heads=4, attention_head_dim=32,
                in_channels=16, out_channels=16, num_layers=2, cross_attention_dim=128, caption_channels=None,
                cross_attention_adaln=True, apply_gated_attention=True, av_ca_timestep_scale_multiplier=1000)
    cs = load_dit(m, cfg, seed=2)
    B, F_, H, W, S, Na = 2, 3, 4, 5, 24, 7
    N = F_ * H * W
    lat, ctx = rnd((B, N, 16), 20), rnd((B, S, 128), 21, 0.5)
    alat, actx = rnd((B, Na, 128), 22), rnd((B, S, 64), 23, 0.5)
    pos = ref_positions(B, F_, H, W, 25.0)
    apos = A(synthetic.audio_positions(B, Na, 25.0))
    sig_v, sig_a = np.array([0.9, 0.4], np.float32), np.array([0.8, 0.3], np.float32)
    vmod = ref_model.Modality(latent=mx.array(lat), context=mx.array(ctx), context_mask=None,
                              timesteps=mx.array(sig_v), positions=pos, sigma=mx.array(sig_v))
    amod = ref_model.Modality(latent=mx.array(alat), context=mx.array(actx), context_mask=None,
                              timesteps=mx.array(sig_a), positions=mx.array(apos), sigma=mx.array(sig_a))
    vv, av = m(vmod, amod)
    x0v, x0a = ref_model.X0Model(m)(vmod, amod)
    # video-only inference on the AV model (audio=None path, model.py:829-851)
    v_only = m(vmod, None)[0]
    # STG perturbation: skip video self-attention in block 1 and a2v in block 0 for the whole batch
    pc = ref_pert.PerturbationConfig(perturbations=[
        ref_pert.Perturbation(type=ref_pert.PerturbationType.SKIP_VIDEO_SELF_ATTN, blocks=[1]),
        ref_pert.Perturbation(type=ref_pert.PerturbationType.SKIP_A2V_CROSS_ATTN, blocks=[0]),
    ])
    bp = ref_pert.BatchedPerturbationConfig(perturbations=[pc, pc])
    pv, pa = m(vmod, amod, perturbations=bp)
    np.savez_compressed(os.path.join(HERE, "dit_v2_av.npz"), latent=lat, context=ctx, positions=A(pos),
                        audio_latent=alat, audio_context=actx, audio_positions=apos, sigma_video=sig_v,
                        sigma_audio=sig_a, velocity_video=A(vv), velocity_audio=A(av), x0_video=A(x0v),
                        x0_audio=A(x0a), velocity_video_only=A(v_only), velocity_video_stg=A(pv),
                        velocity_audio_stg=A(pa), weight_checksum=cs)
    print("dit_v2_av", A(vv).shape, A(av).shape)


def golden_rope():
    # production-size table (D=4096, 32 heads, 768x512x65 grid), sampled tokens
    pos = A(ref_positions(1, 9, 16, 24, 24.0))
    sel = np.arange(0, pos.shape[2], 173)
    p = pos[:, :, sel, :]
    cos, sin = ref_rope.precompute_freqs_cis(mx.array(p), dim=4096, out_dtype=mx.float32, theta=10000.0,
                                             max_pos=[20, 2048, 2048], use_middle_indices_grid=True,
                                             num_attention_heads=32, rope_type=ref_rope.LTXRopeType.SPLIT)
    x = rnd((1, len(sel), 4096), 30)
    y = ref_rope.apply_split_rotary_emb(mx.array(x), cos, sin)
    # 1-D temporal cross-modal table (model.py:320-343)
    c1, s1 = ref_rope.precompute_freqs_cis(mx.array(p[:, 0:1]), dim=2048, out_dtype=mx.float32, theta=10000.0,
                                           max_pos=[20], use_middle_indices_grid=True, num_attention_heads=32,
                                           rope_type=ref_rope.LTXRopeType.SPLIT)
    np.savez_compressed(os.path.join(HERE, "rope.npz"), positions=p, cos=A(cos), sin=A(sin), x=x, y=A(y),
                        cos_1d=A(c1), sin_1d=A(s1), token_index=sel)
    print("rope", A(cos).shape, A(c1).shape)


def _ref_decoder(cfg, seed):
    w = synthetic.vae_weights(cfg, seed=seed)
    dec = ref_vae.SimpleVideoDecoder(decoder_blocks=cfg.decoder_blocks, base_channels=cfg.base_channels,
                                     timestep_conditioning=cfg.timestep_conditioning)
    from safetensors.torch import save_file
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "vae.safetensors")
        save_file({k: v.contiguous() for k, v in w.items()}, path)
        ref_vae.load_vae_decoder_weights(dec, path)        # the reference's own loader
    dec.decode_noise_scale = 0.0
    return dec, checksum(w)


def golden_vae():
    blocks = [["res_x", {"num_layers": 2}], ["compress_all", {"multiplier": 2, "residual": True}],
              ["res_x", {"num_layers": 2}], ["compress_all", {"multiplier": 2, "residual": True}],
              ["res_x", {"num_layers": 1}], ["compress_all", {"multiplier": 2, "residual": True}],
              ["res_x", {"num_layers": 1}]]
    cfg = synthetic.VaeConfig(decoder_blocks=blocks, base_channels=8, timestep_conditioning=True)
    dec, cs = _ref_decoder(cfg, seed=3)
    lat = rnd((1, 128, 3, 2, 3), 40)
    out = A(dec(mx.array(lat), timestep=0.05, show_progress=False))
    out_c = A(dec(mx.array(lat), timestep=0.05, show_progress=False, causal=True))
    out_nt = A(dec(mx.array(lat), timestep=None, show_progress=False))
    # chunked decode_latent: 9 latent frames -> chunks (0,7),(5,9), cross-fade, uint8
    lat9 = rnd((1, 128, 9, 2, 2), 41)
    frames = A(ref_vae.decode_latent(mx.array(lat9), dec, timestep=0.05))
    frames3 = A(ref_vae.decode_latent(mx.array(lat[0]), dec, timestep=0.05))
    np.savez_compressed(os.path.join(HERE, "vae_v20.npz"), latent=lat, video=out, video_causal=out_c,
                        video_no_timestep=out_nt, latent9=lat9, frames9=frames, frames3=frames3,
                        weight_checksum=cs, decoder_blocks=repr(blocks))
    print("vae_v20", out.shape, frames.shape, frames3.shape)

    # decode_tiled (tiling.py:252-412) over the small V2.0 decoder above: 3x3 spatial tiles of 64 px / 32 px overlap
    lat_t = rnd((1, 128, 2, 4, 4), 44)
    cfg_t = ref_tiling.TilingConfig(spatial_config=ref_tiling.SpatialTilingConfig(64, 32), temporal_config=None)
    tiled = A(next(ref_tiling.decode_tiled(mx.array(lat_t), lambda x, timestep=None: dec(x, timestep=timestep,
                                           show_progress=False), cfg_t, timestep=0.05, show_progress=False)))
    # temporal tiling needs >= 16-frame tiles: 5 latent frames, tiles of 2 latent frames (16 px frames), overlap 1
    lat_tt = rnd((1, 128, 5, 2, 2), 45)
    cfg_tt = ref_tiling.TilingConfig(spatial_config=None,
                                     temporal_config=ref_tiling.TemporalTilingConfig(16, 8))
    tiled_t = A(next(ref_tiling.decode_tiled(mx.array(lat_tt), lambda x, timestep=None: dec(x, timestep=timestep,
                                             show_progress=False), cfg_tt, timestep=0.05, show_progress=False)))
    np.savez_compressed(os.path.join(HERE, "vae_tiled.npz"), latent_spatial=lat_t, video_spatial=tiled,
                        latent_temporal=lat_tt, video_temporal=tiled_t, weight_checksum=cs,
                        decoder_blocks=repr(blocks))
    print("vae_tiled", tiled.shape, tiled_t.shape)
    # a V2.3-style stack: separate temporal / spatial upsamplers, no timestep conditioning
    blocks23 = [["res_x", {"num_layers": 1}], ["compress_space", {"multiplier": 2, "residual": True}],
                ["res_x", {"num_layers": 1}], ["compress_time", {"multiplier": 2, "residual": False}],
                ["res_x", {"num_layers": 1}], ["compress_all", {"multiplier": 1, "residual": True}],
                ["res_x", {"num_layers": 1}]]
    cfg23 = synthetic.VaeConfig(decoder_blocks=blocks23, base_channels=16, timestep_conditioning=False)
    dec23, cs23 = _ref_decoder(cfg23, seed=4)
    lat23 = rnd((1, 128, 2, 2, 2), 42)
    out23 = A(dec23(mx.array(lat23), timestep=None, show_progress=False))
    np.savez_compressed(os.path.join(HERE, "vae_v23.npz"), latent=lat23, video=out23, weight_checksum=cs23,
                        decoder_blocks=repr(blocks23))
    print("vae_v23", out23.shape)


def golden_ops():
    conv = ref_vae.Conv3dSimple(6, 10)
    wt, b = rnd((10, 6, 3, 3, 3), 50, 0.1), rnd((10,), 51, 0.1)
    conv.weight, conv.bias = mx.array(wt), mx.array(b)
    x = rnd((2, 6, 4, 5, 7), 52)
    y_nc = A(conv(mx.array(x), causal=False))
    y_c = A(conv(mx.array(x), causal=True))
    u = rnd((1, 48, 3, 4, 5), 53)
    un = A(ref_ops.unpatchify(mx.array(u), patch_size_hw=4, patch_size_t=1))
    up = ref_vae.DepthToSpaceUpsample3d(16, stride=(2, 2, 2), residual=True, out_channels_reduction_factor=2)
    d = rnd((1, 64, 2, 3, 4), 54)
    d2s = A(up._depth_to_space(mx.array(d), 8))
    masks = {f"mask_{l}_{a}_{b_}_{int(z)}": A(ref_tiling.compute_trapezoidal_mask_1d(l, a, b_, z))
             for (l, a, b_, z) in [(16, 4, 4, False), (16, 4, 0, True), (8, 0, 3, False), (5, 8, 8, False)]}
    np.savez_compressed(os.path.join(HERE, "ops.npz"), conv_w=wt, conv_b=b, conv_x=x, conv_y=y_nc, conv_y_causal=y_c,
                        unpatchify_x=u, unpatchify_y=un, d2s_x=d, d2s_y=d2s, **masks)
    print("ops", y_nc.shape, un.shape, d2s.shape)


def golden_vae_encoder():
    """The reference's SimpleVideoEncoder at its fixed widths (128..1024 channels), weights through its own loader,
    on a 9-frame 32x32 clip and a single 64x32 image: (1,3,9,32,32) -> (1,128,2,1,1), (1,3,1,64,32) -> (1,128,1,2,1)."""
    ref_enc = importlib.import_module("LTX_2_MLX.model.video_vae.simple_encoder")
    w = dict(synthetic.iter_vae_encoder_weights(seed=11))
    enc = ref_enc.SimpleVideoEncoder()
    from safetensors.torch import save_file
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "enc.safetensors")
        save_file({k: v.contiguous() for k, v in w.items()}, path)
        ref_enc.load_vae_encoder_weights(enc, path)       # the reference's own loader
    clip = np.clip(rnd((1, 3, 9, 32, 32), 70, 0.5), -1, 1).astype(np.float32)
    image = np.clip(rnd((1, 3, 1, 64, 32), 71, 0.5), -1, 1).astype(np.float32)
    lat_clip = A(enc(mx.array(clip), show_progress=False))
    lat_image = A(enc(mx.array(image), show_progress=False))
    # building blocks on small tensors (channel order of patchify / space-to-depth, group-mean residual)
    x = rnd((1, 3, 2, 8, 12), 72)
    pat = A(ref_ops.patchify(mx.array(x), patch_size_hw=4, patch_size_t=1))
    down = ref_enc.SpaceToDepthDownsample3d(8, 16, stride=(2, 2, 2))
    dw, db = rnd((2, 8, 3, 3, 3), 73, 0.1), rnd((2,), 74, 0.1)
    down.conv.weight, down.conv.bias = mx.array(dw), mx.array(db)
    dx = rnd((1, 8, 3, 4, 6), 75)
    dy = A(down(mx.array(dx), causal=True))
    np.savez_compressed(os.path.join(HERE, "vae_encoder.npz"), clip=clip, image=image, latent_clip=lat_clip,
                        latent_image=lat_image, weight_checksum=checksum(w), patchify_x=x, patchify_y=pat,
                        down_w=dw, down_b=db, down_x=dx, down_y=dy)
    print("vae_encoder", lat_clip.shape, lat_image.shape, pat.shape, dy.shape)


def golden_upscaler():
    """The reference's SpatialUpscaler (64 mid channels, 8 groups, 2 blocks per stage to keep the fixture small; the
    class takes these as constructor arguments), weights through its own loader: (1,16,3,4,5) -> (1,16,3,8,10)."""
    ref_up = importlib.import_module("LTX_2_MLX.model.upscaler.spatial")
    cin, mid, groups, blocks = 16, 64, 8, 2
    w = dict(synthetic.iter_upscaler_weights(seed=13, in_channels=cin, mid_channels=mid, blocks=blocks))
    up = ref_up.SpatialUpscaler(in_channels=cin, mid_channels=mid, num_blocks_per_stage=blocks, num_groups=groups)
    from safetensors.torch import save_file
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "up.safetensors")
        save_file({k: v.contiguous() for k, v in w.items()}, path)
        ref_up.load_spatial_upscaler_weights(up, path)
    lat = rnd((1, cin, 3, 4, 5), 80)
    out = A(up(mx.array(lat)))
    np.savez_compressed(os.path.join(HERE, "upscaler.npz"), latent=lat, upscaled=out, weight_checksum=checksum(w),
                        cfg=np.asarray([cin, mid, groups, blocks]))
    print("upscaler", out.shape)


def golden_connector():
    """Embeddings1DConnector (model/text_encoder/connector.py) -- the reference's own class, both RoPE layouts, gated
    attention on, learnable registers (sequence extended to 1024).  The Metal interleaved-RoPE kernel cannot run here, so
    the reference's own naive fallback (rope.py:76-89) is selected."""
    _stub_package("LTX_2_MLX.model.text_encoder", f"{REF}/LTX_2_MLX/model/text_encoder")
    ref_conn = importlib.import_module("LTX_2_MLX.model.text_encoder.connector")
    ref_rope._HAS_FUSED_ROPE = False
    heads, hd, layers, regs = 4, 64, 2, 128
    w = dict(synthetic.iter_connector_weights(heads=heads, head_dim=hd, layers=layers, registers=regs, gated=True, seed=17))
    x = rnd((2, 20, heads * hd), 700)
    out = {}
    for name, rt in (("interleaved", ref_rope.LTXRopeType.INTERLEAVED), ("split", ref_rope.LTXRopeType.SPLIT)):
        conn = ref_conn.Embeddings1DConnector(attention_head_dim=hd, num_attention_heads=heads, num_layers=layers,
                                              num_learnable_registers=regs, rope_type=rt, apply_gated_attention=True)
        for k, v in w.items():
            set_by_key(conn, k, v)
        # SPLIT with B > 1 breaks inside the reference (apply_split_rotary_emb reshapes with the table's batch of 1)
        y, mask = conn(mx.array(x if name == "interleaved" else x[:1]), None)
        out[name] = A(y).astype(np.float32)
        assert float(np.abs(A(mask)).max()) == 0.0
    np.savez_compressed(os.path.join(HERE, "connector.npz"), x=x, weight_checksum=checksum(w),
                        cfg=np.array([heads, hd, layers, regs]), **{"y_" + k: v for k, v in out.items()})
    print("connector", {k: v.shape for k, v in out.items()})


def _ref_function(path, name):
    """Compile ONE function of a reference file that cannot be imported as a module here (pipelines/common.py pulls
    in PIL, the encoder, ...): its source text is taken from the reference file at run time and executed as is."""
    import ast
    src = open(path).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"mx": mx}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def golden_sampling():
    """The elementwise tail of the reference's denoise loops (pipelines/distilled.py:243-251, one_stage.py:284-320):
    CFGGuider.guide -> post_process_latent -> EulerDiffusionStep.step, run by the reference's own code."""
    from LTX_2_MLX.components import diffusion_steps as ref_steps
    from LTX_2_MLX.components import guiders as ref_guiders
    from LTX_2_MLX.components import schedulers as ref_sched
    post = _ref_function(f"{REF}/LTX_2_MLX/pipelines/common.py", "post_process_latent")
    B, T, C = 2, 24, 16
    sample, cond, uncond, clean = rnd((B, T, C), 60), rnd((B, T, C), 61), rnd((B, T, C), 62), rnd((B, T, C), 63)
    mask = (np.random.default_rng(64).random((B, T)) > 0.3).astype(np.float32)
    mask[0, :3] = 0.25                                    # fractional masks occur with strength < 1 conditioning
    sigmas = np.asarray(ref_sched.DISTILLED_SIGMA_VALUES, np.float32)
    out = {}
    stepper = ref_steps.EulerDiffusionStep()
    for idx in (0, 5, 7):
        plain = stepper.step(mx.array(sample), mx.array(cond), mx.array(sigmas), idx)
        guided = ref_guiders.CFGGuider(3.0).guide(mx.array(cond), mx.array(uncond))
        blended = post(guided, mx.array(mask), mx.array(clean))
        full = stepper.step(mx.array(sample), blended, mx.array(sigmas), idx)
        out[f"plain_{idx}"], out[f"full_{idx}"] = A(plain), A(full)
    out["guided"], out["blended"] = A(guided), A(blended)
    np.savez_compressed(os.path.join(HERE, "sampling.npz"), sample=sample, cond=cond, uncond=uncond, clean=clean,
                        mask=mask, sigmas=sigmas, cfg_scale=np.float32(3.0), **out)
    print("sampling", out["full_0"].shape)


if __name__ == "__main__":
    if "--connector-only" in sys.argv:
        golden_connector()
        sys.exit(0)
    if "--upscaler-only" in sys.argv:
        golden_upscaler()
        sys.exit(0)
    if "--encoder-only" in sys.argv:
        golden_vae_encoder()
        sys.exit(0)
    golden_sampling()
    if "--sampling-only" in sys.argv:
        sys.exit(0)
    golden_vae_encoder()
    golden_upscaler()
    golden_ops()
    golden_rope()
    golden_dit_v1()
    golden_dit_v2_av()
    golden_vae()
    golden_connector()
