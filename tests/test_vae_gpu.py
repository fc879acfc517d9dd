"""GPU parity of the video-VAE decode path (SimpleVideoDecoder / decode_latent mirrors -> C ABI) against the
oracle on the same seeded synthetic checkpoint.  The oracle runs fp32 on bf16-rounded conv weights; the engine keeps
activations in bf16 between convs (fp32 accumulation), so the tolerance is a relative L2 error of 3e-2 plus the
reference's own metric, Pearson r (tests/test_parity.py:53-59; its gate is r >= 0.95), >= 0.999."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BLOCKS_V20 = [["res_x", {"num_layers": 2}], ["compress_all", {"multiplier": 2, "residual": True}],
              ["res_x", {"num_layers": 1}], ["compress_all", {"multiplier": 2, "residual": True}],
              ["res_x", {"num_layers": 1}], ["compress_all", {"multiplier": 2, "residual": True}],
              ["res_x", {"num_layers": 1}]]
BLOCKS_V23 = [["res_x", {"num_layers": 1}], ["compress_space", {"multiplier": 2, "residual": True}],
              ["res_x", {"num_layers": 1}], ["compress_time", {"multiplier": 2, "residual": False}],
              ["res_x", {"num_layers": 1}], ["compress_all", {"multiplier": 1, "residual": True}],
              ["res_x", {"num_layers": 1}]]


def pearson(a, b):
    return float(np.corrcoef(a.double().flatten().cpu().numpy(), b.double().flatten().cpu().numpy())[0, 1])


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm())


def bf16_round(w):
    return {k: (v.to(torch.bfloat16).float() if v.ndim >= 2 and "scale_shift_table" not in k else v) for k, v in w.items()}


def build(blocks, base, tc, seed):
    from ltx2_b200 import synthetic
    from ltx2_b200.video_vae import SimpleVideoDecoder
    cfg = synthetic.VaeConfig(decoder_blocks=blocks, base_channels=base, timestep_conditioning=tc)
    w = synthetic.vae_weights(cfg, seed=seed)
    dec = SimpleVideoDecoder(decoder_blocks=blocks, base_channels=base, timestep_conditioning=tc)
    dec.load_weights(w)
    assert dec.missing_weights() == []
    dec.decode_noise_scale = 0.0          # deterministic, as tests/test_parity.py:359
    return dec, bf16_round(w)


@pytest.mark.parametrize("causal", [False, True])
def test_vae_v20_decode_matches_oracle(causal):
    from ltx2_b200 import synthetic
    from oracle import vae_oracle as V
    dec, w = build(BLOCKS_V20, 64, True, seed=21)
    lat = synthetic.latents((1, 128, 3, 2, 3), seed=200)
    ref = V.vae_decode(w, lat, decoder_blocks=BLOCKS_V20, base_channels=64, timestep=0.05, causal=causal)
    out = dec(lat, timestep=0.05, causal=causal)
    assert out.shape == ref.shape == (1, 3, 17, 64, 96) and out.dtype == torch.float32
    assert rel(out, ref) < 3e-2, rel(out, ref)
    assert pearson(out, ref) > 0.999


def test_vae_no_timestep_and_batch2():
    from ltx2_b200 import synthetic
    from oracle import vae_oracle as V
    dec, w = build(BLOCKS_V20, 64, True, seed=22)
    lat = synthetic.latents((2, 128, 2, 3, 2), seed=201)
    ref = V.vae_decode(w, lat, decoder_blocks=BLOCKS_V20, base_channels=64, timestep=None)
    out = dec(lat, timestep=None)
    assert rel(out, ref) < 3e-2 and pearson(out, ref) > 0.999
    # bf16 / numpy latents are accepted
    out2 = dec(lat.numpy(), timestep=None)
    assert torch.equal(out2, out)
    # noise injection changes the result and is reproducible with a seeded generator
    dec.decode_noise_scale = 0.025
    dec._noise_generator = torch.Generator(device="cuda").manual_seed(5)
    n1 = dec(lat, timestep=0.05)
    dec._noise_generator = torch.Generator(device="cuda").manual_seed(5)
    n2 = dec(lat, timestep=0.05)
    assert torch.equal(n1, n2) and not torch.allclose(n1, dec(lat, timestep=None))
    g = torch.Generator(device="cuda").manual_seed(5)
    noise = torch.randn(lat.shape, device="cuda", generator=g)
    ref_n = V.vae_decode(w, lat, decoder_blocks=BLOCKS_V20, base_channels=64, timestep=0.05, decode_noise_scale=0.025,
                         noise=noise.cpu())
    assert rel(n1, ref_n) < 3e-2


def test_vae_v23_style_stack_matches_oracle():
    from ltx2_b200 import synthetic
    from oracle import vae_oracle as V
    dec, w = build(BLOCKS_V23, 64, False, seed=23)
    lat = synthetic.latents((1, 128, 2, 2, 2), seed=202)
    ref = V.vae_decode(w, lat, decoder_blocks=BLOCKS_V23, base_channels=64, timestep=None, timestep_conditioning=False)
    out = dec(lat, timestep=None)
    assert out.shape == ref.shape
    assert rel(out, ref) < 3e-2 and pearson(out, ref) > 0.999


def test_decode_latent_chunked_blend_and_uint8():
    from ltx2_b200 import synthetic
    from ltx2_b200.video_vae import chunk_plan, decode_latent, decode_latent_video
    from oracle import vae_oracle as V
    dec, w = build(BLOCKS_V20, 64, True, seed=24)
    assert chunk_plan(9) == V.chunk_plan(9) == [(0, 7), (5, 9)]
    assert chunk_plan(16) == V.chunk_plan(16) and chunk_plan(31) == V.chunk_plan(31)
    lat = synthetic.latents((1, 128, 9, 2, 2), seed=203)
    kw = dict(decoder_blocks=BLOCKS_V20, base_channels=64)
    parts = [V.vae_decode(w, lat[:, :, a:b], timestep=0.05, **kw) for a, b in V.chunk_plan(9)]
    ref_video = V.blend_chunks(parts, 9)
    video = decode_latent_video(lat, dec)
    assert video.shape == ref_video.shape == (1, 3, 65, 64, 64)
    assert rel(video, ref_video) < 3e-2
    frames = decode_latent(lat[0], dec)                 # 4-D latent, like generate.py:2085
    assert frames.shape == (65, 64, 64, 3) and frames.dtype == torch.uint8
    ref_frames = V.to_uint8_frames(ref_video)
    d = (frames.cpu().int() - ref_frames.int()).abs()
    assert d.float().mean() < 2.0 and d.max() <= 24      # uint8 of bf16-path pixels: a few grey levels
    # uint8 conversion itself is exact on identical input (truncation, not rounding)
    assert torch.equal(V.to_uint8_frames(video.cpu()), frames.cpu())
    # single-pass path (T <= 7)
    short = decode_latent(lat[:, :, :3], dec)
    assert short.shape == (17, 64, 64, 3)


def test_vae_errors():
    from ltx2_b200._lib import Ltx2Error
    from ltx2_b200.video_vae import SimpleVideoDecoder
    with pytest.raises(ValueError, match="Unknown decoder block"):
        SimpleVideoDecoder(decoder_blocks=[["bogus", 1]], base_channels=64)
    dec = SimpleVideoDecoder(decoder_blocks=BLOCKS_V20, base_channels=64)
    with pytest.raises(Ltx2Error, match="has not been set"):
        dec(torch.zeros(1, 128, 2, 2, 2))


def test_decode_tiled_matches_oracle():
    from ltx2_b200 import synthetic
    from ltx2_b200.tiling import SpatialTilingConfig, TemporalTilingConfig, TilingConfig, decode_tiled
    from oracle import vae_oracle as V
    dec, w = build(BLOCKS_V20, 64, True, seed=25)
    ref_dec = lambda x: V.vae_decode(w, x, decoder_blocks=BLOCKS_V20, base_channels=64, timestep=0.05)  # noqa: E731
    lat = synthetic.latents((1, 128, 2, 4, 4), seed=204)
    out = next(decode_tiled(lat, dec, TilingConfig(SpatialTilingConfig(64, 32), None), timestep=0.05))
    ref = V.decode_tiled(ref_dec, lat, tile_px=64, overlap_px=32)
    assert out.shape == ref.shape == (1, 3, 9, 128, 128)
    assert rel(out, ref) < 3e-2 and pearson(out, ref) > 0.999
    lat_t = synthetic.latents((1, 128, 5, 2, 2), seed=205)
    out_t = next(decode_tiled(lat_t, dec, TilingConfig(None, TemporalTilingConfig(16, 8)), timestep=0.05))
    ref_t = V.decode_tiled(ref_dec, lat_t, tile_px=None, tile_frames=16, overlap_frames=8)
    assert rel(out_t, ref_t) < 3e-2
    # no tiling needed -> identical to a plain decode
    one = next(decode_tiled(lat, dec, TilingConfig(SpatialTilingConfig(512, 64), None), timestep=0.05))
    assert torch.allclose(one, dec(lat, timestep=0.05), atol=1e-6)


def test_load_vae_decoder_weights_from_safetensors_file(tmp_path):
    """load_vae_decoder_weights(decoder, path) (simple_decoder.py:566 call surface) on a temporary safetensors file."""
    from safetensors.torch import save_file
    from ltx2_b200 import synthetic
    from ltx2_b200.video_vae import SimpleVideoDecoder, load_vae_decoder_weights
    dec, _ = build(BLOCKS_V20, 64, True, seed=27)
    cfg = synthetic.VaeConfig(decoder_blocks=BLOCKS_V20, base_channels=64, timestep_conditioning=True)
    sd = {k: v.contiguous() for k, v in synthetic.vae_weights(cfg, seed=27).items()}
    sd["model.diffusion_model.proj_out.bias"] = torch.zeros(4)            # foreign tensors are ignored
    path = str(tmp_path / "vae.safetensors")
    save_file(sd, path)
    dec2 = SimpleVideoDecoder(decoder_blocks=BLOCKS_V20, base_channels=64, timestep_conditioning=True)
    load_vae_decoder_weights(dec2, path)
    assert dec2.missing_weights() == []
    dec2.decode_noise_scale = 0.0
    lat = synthetic.latents((1, 128, 2, 2, 3), seed=270)
    assert torch.equal(dec2(lat, timestep=0.05), dec(lat, timestep=0.05))


def test_fused_conv_producer_matches_unfused_sequence(monkeypatch):
    """The conv epilogue that writes the next conv's normalised / activated / padded input (128- and 256-channel
    stages) against the separate norm_act_pad pass it replaces (LTX2_VAE_FUSE=0): same stored bf16 activations go into
    the norm, so the two agree to bf16 rounding of the normalised values; both match the oracle."""
    from ltx2_b200 import synthetic
    from oracle import vae_oracle as V
    dec, w = build(BLOCKS_V20, 64, True, seed=28)
    for shape, causal in (((1, 128, 3, 2, 3), False), ((2, 128, 1, 3, 2), True), ((1, 128, 2, 5, 4), False)):
        lat = synthetic.latents(shape, seed=280)
        monkeypatch.setenv("LTX2_VAE_FUSE", "1")
        fused = dec(lat, timestep=0.05, causal=causal)
        monkeypatch.setenv("LTX2_VAE_FUSE", "0")
        plain = dec(lat, timestep=0.05, causal=causal)
        monkeypatch.delenv("LTX2_VAE_FUSE")
        ref = V.vae_decode(w, lat, decoder_blocks=BLOCKS_V20, base_channels=64, timestep=0.05, causal=causal)
        assert rel(fused, plain) < 1e-2, rel(fused, plain)
        assert rel(fused, ref) < 3e-2 and rel(plain, ref) < 3e-2, (rel(fused, ref), rel(plain, ref))
        assert pearson(fused, ref) > 0.999
