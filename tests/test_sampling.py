"""Fused denoise-step update (ltx2_denoise_update) and its Python mirrors vs the reference's own code (golden vectors
from tests/golden/make_golden.py: EulerDiffusionStep.step, CFGGuider.guide, post_process_latent)."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _g():
    g = np.load(os.path.join(GOLDEN, "sampling.npz"))
    return {k: torch.from_numpy(g[k]) if g[k].ndim else float(g[k]) for k in g.files}


def test_oracle_matches_reference_golden():
    from oracle import sampling_oracle as O
    g = _g()
    guided = O.cfg_guide(g["cond"], g["uncond"], g["cfg_scale"])
    assert torch.allclose(guided, g["guided"], atol=1e-6)
    blended = O.post_process_latent(guided, g["mask"], g["clean"])
    assert torch.allclose(blended, g["blended"], atol=1e-6)
    for idx in (0, 5, 7):
        assert torch.allclose(O.euler_step(g["sample"], g["cond"], g["sigmas"], idx), g[f"plain_{idx}"], atol=2e-5)
        assert torch.allclose(O.euler_step(g["sample"], blended, g["sigmas"], idx), g[f"full_{idx}"], atol=2e-5)
    with pytest.raises(ValueError):
        O.to_velocity(g["sample"], 0.0, g["cond"])


@pytest.mark.gpu
def test_denoise_update_matches_reference_golden():
    from ltx2_b200 import sampling
    g = _g()
    s = g["sigmas"]
    for idx in (0, 5, 7):
        out = sampling.EulerDiffusionStep().step(g["sample"], g["cond"], s, idx)
        assert torch.allclose(out.cpu(), g[f"plain_{idx}"], atol=2e-5)
        out, den = sampling.denoise_update(g["sample"], g["cond"], float(s[idx]), float(s[idx + 1]), uncond_x0=g["uncond"],
                                           cfg_scale=g["cfg_scale"], denoise_mask=g["mask"], clean_latent=g["clean"],
                                           return_denoised=True)
        assert torch.allclose(den.cpu(), g["blended"], atol=1e-5)
        assert torch.allclose(out.cpu(), g[f"full_{idx}"], atol=3e-5)
    with pytest.raises(ValueError, match="Sigma can't be 0.0"):
        sampling.denoise_update(g["sample"], g["cond"], 0.0, 0.0)
    cuda = g["cond"].cuda()
    assert torch.allclose(sampling.CFGGuider(3.0).guide(cuda, g["uncond"].cuda()).cpu(), g["guided"], atol=1e-6)
    assert torch.allclose(sampling.post_process_latent(g["guided"], g["mask"], g["clean"]), g["blended"], atol=1e-6)


@pytest.mark.gpu
def test_euler_denoising_loop_matches_oracle_loop():
    """Eight distilled steps on a small DiT: the device-resident loop (one X0Model call + one fused update per step)
    against the same loop written with the oracle's step functions around the same engine calls."""
    from ltx2_b200 import sampling, synthetic
    from ltx2_b200.transformer import LTXModel, LTXModelType, Modality, X0Model
    from oracle import sampling_oracle as O
    dev = torch.device("cuda:0")
    heads, hd = 4, 64
    cfg = synthetic.DitConfig(num_attention_heads=heads, attention_head_dim=hd, in_channels=32, out_channels=32,
                              num_layers=2, cross_attention_dim=heads * hd, caption_channels=96)
    model = LTXModel(model_type=LTXModelType.VideoOnly, num_attention_heads=heads, attention_head_dim=hd, in_channels=32,
                     out_channels=32, num_layers=2, cross_attention_dim=heads * hd, caption_channels=96, device=dev)
    from ltx2_b200.loader import iter_engine_weights
    model.load_weights(iter_engine_weights(synthetic.iter_dit_weights(cfg, seed=5, device=dev, dtype=torch.bfloat16), False))
    x0m = X0Model(model)
    B, F, H, W = 1, 2, 4, 6
    N = F * H * W
    lat = synthetic.latents((B, N, 32), seed=7).to(dev)
    ctx = (0.1 * synthetic.latents((B, 16, 96), seed=8)).to(dev)
    nctx = (0.1 * synthetic.latents((B, 16, 96), seed=9)).to(dev)
    pos = synthetic.video_positions(B, F, H, W, fps=24.0).to(dev)
    mask = torch.ones(B, N, device=dev)
    mask[:, :H * W] = 0.0                                         # first latent frame is conditioning
    clean = synthetic.latents((B, N, 32), seed=10).to(dev)
    sigmas = [1.0, 0.99375, 0.9875, 0.98125, 0.975, 0.909375, 0.725, 0.421875, 0.0]
    out = sampling.euler_denoising_loop(x0m, lat, ctx, pos, sigmas, denoise_mask=mask, clean_latent=clean,
                                        negative_context=nctx, cfg_scale=3.0)
    x = lat.clone()
    for i in range(len(sigmas) - 1):
        sig = torch.full((B,), sigmas[i], device=dev)
        md = dict(context_mask=None, timesteps=mask * sigmas[i], positions=pos, sigma=sig)
        cond = x0m(Modality(latent=x, context=ctx, **md))
        unc = x0m(Modality(latent=x, context=nctx, **md))
        d = O.post_process_latent(O.cfg_guide(cond, unc, 3.0), mask, clean)
        x = O.euler_step(x, d, sigmas, i)
    assert torch.isfinite(out).all()
    assert float((out - x).abs().max()) < 1e-4 * max(1.0, float(x.abs().max()))


def test_timestep_classes_from_mask_cpu():
    """Host logic of the sync-free loop: the (batch, mask value) classes reproduce timesteps_from_mask
    (pipelines/common.py:193-203: mask * sigma) exactly, per batch element."""
    from ltx2_b200 import sampling
    mask = torch.ones(3, 40)
    mask[:, :8] = 0.0
    mask[1, 8:16] = 0.25
    mask[2] = 1.0
    vals, rows = sampling.timestep_classes_from_mask(mask)
    assert rows.dtype == torch.int32 and rows.shape == (3, 40)
    assert torch.equal(vals[rows.long()], mask)
    # classes never mix batch elements (sigma differs per batch element in general)
    owners = [set(rows[b].tolist()) for b in range(3)]
    assert not (owners[0] & owners[1]) and not (owners[1] & owners[2]) and len(owners[2]) == 1
    for sigma in (1.0, 0.421875):
        assert torch.equal((vals * sigma)[rows.long()], mask * sigma)
    many = torch.rand(1, 100)
    assert sampling.timestep_classes_from_mask(many) is None          # > 64 classes: caller falls back to per-token


@pytest.mark.gpu
def test_graph_captured_loop_replays_bit_identically():
    """the 8-step loop (8 forwards + 8 fused updates) captured as one CUDA graph equals the eager
    loop bit for bit, also for a second sample pushed through the same graph and with a denoise mask + CFG."""
    from ltx2_b200 import sampling, synthetic
    from ltx2_b200.loader import iter_engine_weights
    from ltx2_b200.transformer import LTXModel, LTXModelType, X0Model
    dev = torch.device("cuda:0")
    heads, hd = 4, 64
    cfg = synthetic.DitConfig(num_attention_heads=heads, attention_head_dim=hd, in_channels=32, out_channels=32,
                              num_layers=2, cross_attention_dim=heads * hd, caption_channels=96)
    model = LTXModel(model_type=LTXModelType.VideoOnly, num_attention_heads=heads, attention_head_dim=hd, in_channels=32,
                     out_channels=32, num_layers=2, cross_attention_dim=heads * hd, caption_channels=96, device=dev)
    model.load_weights(iter_engine_weights(synthetic.iter_dit_weights(cfg, seed=5, device=dev, dtype=torch.bfloat16), False))
    x0m = X0Model(model)
    B, F, H, W = 1, 2, 4, 6
    N = F * H * W
    pos = synthetic.video_positions(B, F, H, W, fps=24.0).to(dev)
    sigmas = [1.0, 0.99375, 0.9875, 0.98125, 0.975, 0.909375, 0.725, 0.421875, 0.0]
    mask = torch.ones(B, N, device=dev)
    mask[:, :H * W] = 0.0
    for use_mask in (False, True):
        g = sampling.GraphedDenoiser(x0m, sigmas, cfg_scale=3.0 if use_mask else 1.0)
        for seed in (7, 17, 27):
            lat = synthetic.latents((B, N, 32), seed=seed).to(dev)
            ctx = (0.1 * synthetic.latents((B, 16, 96), seed=seed + 1)).to(dev)
            kw = {}
            if use_mask:
                kw = dict(denoise_mask=mask, clean_latent=synthetic.latents((B, N, 32), seed=seed + 2).to(dev),
                          negative_context=(0.1 * synthetic.latents((B, 16, 96), seed=seed + 3)).to(dev))
            out = g(lat, ctx, pos, **kw)
            model.reset_context_cache()
            ref = sampling.euler_denoising_loop(x0m, lat, ctx, pos, sigmas, cfg_scale=g.cfg_scale, **kw)
            assert torch.isfinite(out).all() and torch.equal(out, ref), (use_mask, seed)
