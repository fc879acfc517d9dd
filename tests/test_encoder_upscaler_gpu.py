"""CUDA paths of the video-VAE encoder and the 2x latent spatial upscaler against
  (1) the golden vectors produced by the reference's OWN SimpleVideoEncoder / SpatialUpscaler (tests/golden/make_golden.py)
  (2) the oracles on the same bf16-rounded weights at larger shapes, and the building-block ops against torch.
Tolerances as for the decoder (tests/test_vae_gpu.py): bf16 activations between convs, fp32 accumulation -> relative
L2 <= 3e-2 and Pearson r >= 0.999 on full networks; data-movement ops exact."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm())


def pearson(a, b):
    return float(np.corrcoef(a.double().flatten().cpu().numpy(), b.double().flatten().cpu().numpy())[0, 1])


def bf16_round(w):
    return {k: (v.to(torch.bfloat16).float() if v.ndim >= 4 else v) for k, v in w.items()}


def test_encoder_matches_reference_golden_and_oracle():
    from ltx2_b200 import synthetic
    from ltx2_b200.video_vae_encoder import SimpleVideoEncoder
    from oracle import vae_encoder_oracle as E
    g = np.load(os.path.join(GOLDEN, "vae_encoder.npz"))
    w = dict(synthetic.iter_vae_encoder_weights(seed=11))
    enc = SimpleVideoEncoder()
    assert len(enc.missing_weights()) > 0
    enc.load_weights(w)
    assert enc.missing_weights() == []
    for x, y in (("clip", "latent_clip"), ("image", "latent_image")):
        out = enc(torch.from_numpy(g[x]))
        ref = torch.from_numpy(g[y])                    # the reference's own encoder (fp32 weights)
        assert out.shape == ref.shape and out.dtype == torch.float32 and out.is_cuda
        assert rel(out, ref) < 3e-2, (x, rel(out, ref))
        assert pearson(out, ref) > 0.999
    # a larger clip against the oracle on the bf16-rounded weights: 17 frames @ 128x192 -> latent (1,128,3,4,6)
    video = synthetic.latents((1, 3, 17, 128, 192), seed=600).clamp(-1, 1)
    with torch.no_grad():
        ref = E.vae_encode(bf16_round(w), video)
    out = enc(video)
    assert out.shape == ref.shape == (1, 128, 3, 4, 6)
    assert rel(out, ref) < 3e-2, rel(out, ref)
    assert pearson(out, ref) > 0.999
    with pytest.raises(ValueError, match=r"1 \+ 8\*k frames"):
        enc(torch.zeros(1, 3, 4, 32, 32))


def test_encoder_loader_from_safetensors(tmp_path):
    from safetensors.torch import save_file
    from ltx2_b200 import synthetic
    from ltx2_b200.video_vae_encoder import SimpleVideoEncoder, encode_video, load_vae_encoder_weights
    w = {k: v.contiguous() for k, v in synthetic.iter_vae_encoder_weights(seed=11)}
    path = str(tmp_path / "vae.safetensors")
    save_file(w, path)
    a, b = SimpleVideoEncoder(), SimpleVideoEncoder()
    a.load_weights(w)
    load_vae_encoder_weights(b, path)
    assert b.missing_weights() == []
    frames = (synthetic.latents((9, 64, 64, 3), seed=601).clamp(-1, 1) * 0.5 + 0.5)
    la, lb = encode_video(frames, a), encode_video(frames, b)
    assert la.shape == (1, 128, 2, 2, 2) and torch.equal(la, lb)


def test_upscaler_matches_reference_golden_and_oracle():
    from ltx2_b200 import synthetic
    from ltx2_b200.upscaler import SpatialUpscaler
    from oracle import upscaler_oracle as U
    g = np.load(os.path.join(GOLDEN, "upscaler.npz"))
    cin, mid, groups, blocks = (int(v) for v in g["cfg"])
    w = dict(synthetic.iter_upscaler_weights(seed=13, in_channels=cin, mid_channels=mid, blocks=blocks))
    up = SpatialUpscaler(in_channels=cin, mid_channels=mid, num_blocks_per_stage=blocks, num_groups=groups)
    up.load_weights(w)
    assert up.missing_weights() == []
    out = up(torch.from_numpy(g["latent"]))
    ref = torch.from_numpy(g["upscaled"])               # the reference's own SpatialUpscaler (fp32 weights)
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert rel(out, ref) < 3e-2, rel(out, ref)
    assert pearson(out, ref) > 0.999
    # production widths (128 -> 1024, 32 groups) on a small latent, against the oracle on bf16-rounded weights
    w = dict(synthetic.iter_upscaler_weights(seed=14, blocks=2))
    up = SpatialUpscaler(num_blocks_per_stage=2)
    up.load_weights(w)
    lat = synthetic.latents((1, 128, 2, 6, 9), seed=610)
    with torch.no_grad():
        ref = U.upscale(bf16_round(w), lat, groups=32, blocks=2)
    out = up(lat)
    assert out.shape == ref.shape == (1, 128, 2, 12, 18)
    assert rel(out, ref) < 3e-2, rel(out, ref)
    assert pearson(out, ref) > 0.999


def test_building_blocks_against_torch():
    from ltx2_b200 import conv_stack as cs
    from ltx2_b200._lib import check, lib, ptr, stream_ptr
    from oracle import vae_encoder_oracle as E
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    # patchify: exact data movement
    v = torch.randn(2, 3, 3, 8, 12, device=dev)
    out = torch.empty(2, 3, 2, 3, 64, device=dev, dtype=torch.bfloat16)
    check(lib().ltx2_patchify_video(ptr(v), ptr(out), 2, 3, 8, 12, 64, stream_ptr()))
    ref = E.patchify(v.cpu(), 4).permute(0, 2, 3, 4, 1)
    assert torch.equal(out[..., :48].float().cpu(), ref.to(torch.bfloat16).float()) and float(out[..., 48:].abs().max()) == 0
    # zero / causal padding + pixel-norm + SiLU
    x = torch.randn(1, 3, 4, 5, 64, device=dev).to(torch.bfloat16)
    p = cs.pad_act(x, hw_mode=cs.HW_ZERO, t_mode=cs.T_CAUSAL, act=cs.ACT_PIXELNORM_SILU)
    xf = x.float()
    y = torch.nn.functional.silu(xf * torch.rsqrt((xf * xf).mean(-1, keepdim=True) + 1e-6))
    assert p.shape == (1, 5, 6, 7, 64)
    assert rel(p[:, 2:, 1:-1, 1:-1].float(), y) < 5e-3 and rel(p[:, 0, 1:-1, 1:-1].float(), y[:, 0]) < 5e-3
    assert float(p[:, :, 0].abs().max()) == 0 and float(p[:, :, :, -1].abs().max()) == 0
    # GroupNorm statistics and the fused affine + residual + SiLU pass
    h = (torch.randn(2, 2, 3, 4, 64, device=dev) * 2 + 0.5).to(torch.bfloat16)
    res = torch.randn(2, 2, 3, 4, 64, device=dev).to(torch.bfloat16)
    gw, gb = torch.randn(64, device=dev) * 0.1 + 1, torch.randn(64, device=dev) * 0.1
    st = cs.group_stats(h, 8, 1e-5)
    hg = h.float().reshape(2, 24, 8, 8)                           # [B, THW, groups, C/groups]
    mean = hg.mean(dim=(1, 3))
    var = hg.var(dim=(1, 3), unbiased=False)
    assert torch.allclose(st[..., 0], mean, atol=1e-4) and torch.allclose(st[..., 1], torch.rsqrt(var + 1e-5), rtol=1e-4)
    pp, plain = cs.pad_act(h, hw_mode=cs.HW_ZERO, t_mode=cs.T_ZERO, act=cs.ACT_GROUPNORM_SILU, gn=(st, gw, gb, 8),
                           eps=1e-5, residual=res, want_plain=True)
    yn = ((hg - mean[:, None, :, None]) * torch.rsqrt(var + 1e-5)[:, None, :, None]).reshape(2, 2, 3, 4, 64) * gw + gb
    yr = torch.nn.functional.silu(yn + res.float())
    assert rel(plain.float(), yr) < 5e-3 and torch.equal(pp[:, 1:-1, 1:-1, 1:-1], plain)
    assert float(pp[:, 0].abs().max()) == 0 and float(pp[:, -1].abs().max()) == 0
    # space-to-depth + group-mean residual with the duplicated first frame, against the oracle's downsample tail
    xx = torch.randn(1, 3, 4, 6, 64, device=dev).to(torch.bfloat16)           # [B,T,H,W,Cx]
    yy = torch.randn(1, 4, 4, 6, 16, device=dev).to(torch.bfloat16)           # conv output on T+1 frames, Cout/sp = 16
    o = torch.empty(1, 2, 2, 3, 128, device=dev, dtype=torch.bfloat16)
    check(lib().ltx2_space_to_depth_residual(ptr(yy), ptr(xx), ptr(o), 1, 3, 4, 6, 64, 128, 2, 2, 2, 1, stream_ptr()))
    xn = xx.float().permute(0, 4, 1, 2, 3).cpu()
    xn = torch.cat([xn[:, :, :1], xn], dim=2)
    r = E.space_to_depth(xn, (2, 2, 2))
    r = r.reshape(1, 128, r.shape[1] // 128, 2, 2, 3).mean(dim=2)
    ref = E.space_to_depth(yy.float().permute(0, 4, 1, 2, 3).cpu(), (2, 2, 2)) + r
    assert rel(o.float().permute(0, 4, 1, 2, 3), ref) < 5e-3
    # pixel shuffle: exact
    ys = torch.randn(3, 2, 3, 32, device=dev).to(torch.bfloat16)             # [BF,H,W,4C], C = 8
    os_ = torch.empty(3, 4, 6, 8, device=dev, dtype=torch.bfloat16)
    check(lib().ltx2_pixel_shuffle2(ptr(ys), ptr(os_), 3, 2, 3, 8, stream_ptr()))
    ref = torch.nn.functional.pixel_shuffle(ys.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    assert torch.equal(os_.float(), ref)
