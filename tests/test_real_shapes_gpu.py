"""GPU parity at the shapes BASELINE.json is benchmarked on: the 19B-width DiT
(D = 4096 = 32 heads x 128, S = 1024 text tokens, caption 3840) and the base-128 V2.0 VAE stack, through the C ABI,
against the oracle on the same seeded weights.

Weights are drawn on the GPU in bf16 (seconds instead of minutes for 0.5-2 G parameters) and the oracle receives the
same bf16 values widened to fp32, so what is compared is bf16 activations at the GEMM/attention/conv inputs with fp32
accumulation (engine) against fp32 everywhere (oracle).  Tolerances as in test_dit_gpu.py / test_vae_gpu.py: relative
L2 <= 2e-2 (DiT) / 3e-2 (VAE) and the reference's own metric, Pearson r (tests/test_parity.py:53-59; its gate is
0.95), >= 0.999."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def pearson(a, b):
    return float(np.corrcoef(a.double().flatten().cpu().numpy(), b.double().flatten().cpu().numpy())[0, 1])


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm())


def build_19b_width(layers, *, av=False, seed=5):
    """Engine model + oracle weight dict (fp32 CPU copies of the same bf16 values) at the production widths."""
    from ltx2_b200 import synthetic
    from ltx2_b200.loader import iter_engine_weights
    from ltx2_b200.transformer import LTXModel, LTXModelType
    from oracle import dit_oracle as O
    dev = torch.device("cuda:0")
    cfg = synthetic.DitConfig(num_layers=layers, caption_channels=None if av else 3840, cross_attention_adaln=av,
                              apply_gated_attention=av, audio=av)
    m = LTXModel(model_type=LTXModelType.AudioVideo if av else LTXModelType.VideoOnly, num_layers=layers,
                 caption_channels=None if av else 3840, cross_attention_adaln=av, apply_gated_attention=av,
                 av_ca_timestep_scale_multiplier=1000, device=dev)
    w_cpu = {}

    def tee():
        for k, t in synthetic.iter_dit_weights(cfg, seed=seed, device=dev, dtype=torch.bfloat16):
            w_cpu[k] = t.float().cpu()
            yield k, t

    m.load_weights(iter_engine_weights(tee(), include_audio=av))
    assert m.missing_weights() == []
    return m, O.to_engine_keys(w_cpu)


def video_inputs(B, F, H, W, S, ctx_dim, seed):
    from ltx2_b200 import synthetic
    lat = synthetic.latents((B, F * H * W, 128), seed=seed)
    ctx = synthetic.latents((B, S, ctx_dim), seed=seed + 1, std=0.1).to(torch.bfloat16).float()
    pos = synthetic.video_positions(B, F, H, W, fps=24.0)
    return lat, ctx, pos


def test_dit_19b_width_single_block_configs0():
    """BASELINE.json configs[0]: one DiT block at 256x384 px x 9 latent frames (N = 864), D = 4096, S = 1024."""
    from ltx2_b200.transformer import Modality, X0Model
    from oracle import dit_oracle as O
    torch.set_num_threads(max(1, torch.get_num_threads()))
    m, w = build_19b_width(1)
    lat, ctx, pos = video_inputs(1, 9, 8, 12, 1024, 3840, 400)
    ts = torch.tensor([0.725])
    ref = O.dit_forward(w, dict(latent=lat, context=ctx, timesteps=ts, positions=pos), num_layers=1, heads=32)
    mod = Modality(latent=lat, context=ctx, context_mask=None, timesteps=ts, positions=pos)
    out = m(mod)
    assert out.shape == ref.shape == (1, 864, 128)
    assert rel(out, ref) < 2e-2, rel(out, ref)
    assert pearson(out, ref) > 0.999
    x0 = X0Model(m)(mod)
    assert rel(x0, O.to_x0(lat, ts, ref)) < 2e-2


def test_dit_19b_width_four_blocks_full_token_count():
    """Four blocks of the benchmarked configuration (configs[1] shape: N = 3456, S = 1024, D = 4096, B = 1): covers
    qkv_head_scatter at 32 heads, the K = 16384 residual GEMM, the 3456 x 3456 x 32-head attention and the text
    cross-attention at production size inside a real forward."""
    from ltx2_b200.transformer import Modality
    from oracle import dit_oracle as O
    m, w = build_19b_width(4)
    lat, ctx, pos = video_inputs(1, 9, 16, 24, 1024, 3840, 410)
    ts = torch.tensor([0.909375])
    ref = O.dit_forward(w, dict(latent=lat, context=ctx, timesteps=ts, positions=pos), num_layers=4, heads=32)
    mod = Modality(latent=lat, context=ctx, context_mask=None, timesteps=ts, positions=pos)
    out = m(mod)
    assert rel(out, ref) < 2e-2, rel(out, ref)
    assert pearson(out, ref) > 0.999
    # second call reuses the cached text K/V of this context tensor (V1): bit-identical
    assert torch.equal(m(mod), out)
    # per-token timesteps (image conditioning: first latent frame clean) at the same size
    ts_tok = torch.full((1, 3456), 0.909375)
    ts_tok[:, :16 * 24] = 0.0
    ref_tok = O.dit_forward(w, dict(latent=lat, context=ctx, timesteps=ts_tok, positions=pos), num_layers=4, heads=32)
    out_tok = m(Modality(latent=lat, context=ctx, context_mask=None, timesteps=ts_tok, positions=pos))
    assert rel(out_tok, ref_tok) < 2e-2, rel(out_tok, ref_tok)


def test_dit_v23_av_production_widths():
    """LTX-2.3-style audio+video model at D = 4096 / D_a = 2048 (configs[2] widths): cross_attention_adaln, gated
    attention, a2v / v2a cross-modal attention; 2 blocks, N = 864 video tokens, 65 audio tokens, S = 1024."""
    from ltx2_b200 import synthetic
    from ltx2_b200.transformer import Modality
    from oracle import dit_oracle as O
    m, w = build_19b_width(2, av=True, seed=6)
    lat, ctx, pos = video_inputs(1, 9, 8, 12, 1024, 4096, 420)
    Na = 65
    alat = synthetic.latents((1, Na, 128), seed=430)
    actx = synthetic.latents((1, 1024, 2048), seed=431, std=0.1).to(torch.bfloat16).float()
    apos = synthetic.audio_positions(1, Na)
    sv, sa = torch.tensor([0.725]), torch.tensor([0.6])
    kw = dict(num_layers=2, heads=32, audio_heads=32, v2=True, av_ca_timestep_scale_multiplier=1000)
    rv, ra = O.dit_forward(w, dict(latent=lat, context=ctx, timesteps=sv, positions=pos, sigma=sv),
                           dict(latent=alat, context=actx, timesteps=sa, positions=apos, sigma=sa), **kw)
    ov, oa = m(Modality(latent=lat, context=ctx, context_mask=None, timesteps=sv, positions=pos, sigma=sv),
               Modality(latent=alat, context=actx, context_mask=None, timesteps=sa, positions=apos, sigma=sa))
    assert rel(ov, rv) < 2e-2 and rel(oa, ra) < 2e-2, (rel(ov, rv), rel(oa, ra))
    assert pearson(ov, rv) > 0.999 and pearson(oa, ra) > 0.999


# ---- VAE ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(1024, 1024, 3, 16, 24), (128, 128, 9, 128, 192), (128, 48, 9, 128, 192),
                                   (512, 2048, 2, 32, 48)])
@pytest.mark.parametrize("causal", [False, True])
def test_conv3d_op_production_shapes(shape, causal):
    """ltx2_conv3d (the tcgen05 implicit-GEMM kernel alone) against the oracle's Conv3dSimple at the decoder's real
    channel counts / grids: K = 27*1024 = 27648, W = 192 rows, the 128 -> 48 conv_out width."""
    if causal and shape[0] != 128:
        pytest.skip("causal variant checked on the last-stage shapes only")
    from ltx2_b200 import ops, synthetic
    from oracle import vae_oracle as V
    cin, cout, T, H, W = shape
    x = synthetic.latents((1, cin, T, H, W), seed=500).to(torch.bfloat16)
    wt = (synthetic.latents((cout, cin, 3, 3, 3), seed=501) / (27 * cin) ** 0.5).to(torch.bfloat16)
    b = synthetic.latents((cout,), seed=502) * 0.1
    ref = V.conv3d(x.float(), wt.float(), b, causal=causal)                          # NCDHW fp32
    out = ops.conv3d(x.permute(0, 2, 3, 4, 1).contiguous().cuda(), wt.cuda(), b.cuda(), causal=causal)
    out = out.float().permute(0, 4, 1, 2, 3).cpu()
    assert out.shape == ref.shape
    # bf16 output rounding (2^-9 relative) on O(1) values + fp32 accumulation order
    assert rel(out, ref) < 4e-3, rel(out, ref)
    assert float((out - ref).abs().max()) < 4e-2 * max(1.0, float(ref.abs().max()))


def test_vae_base128_v20_stack_matches_oracle():
    """The benchmarked decoder (base 128: 1024/512/256/128 channels, 5 res blocks per group, timestep conditioning)
    on a 1x128x2x16x24 latent -> 9 frames @ 512x768: the conv kernel at C_in = 1024 inside the real stack,
    norm_act_pad at the 128-channel 9x128x192 activation, the depth-to-space and unpatchify epilogues at full width."""
    from ltx2_b200 import synthetic
    from ltx2_b200.video_vae import SimpleVideoDecoder
    from oracle import vae_oracle as V
    dev = torch.device("cuda:0")
    cfg = synthetic.VaeConfig()
    dec = SimpleVideoDecoder(device=dev)
    w_cpu = {}

    def tee():
        for k, t in synthetic.iter_vae_weights(cfg, seed=0, device=dev, dtype=torch.bfloat16):
            w_cpu[k] = t.float().cpu()
            yield k, t

    dec.load_weights(tee())
    assert dec.missing_weights() == []
    dec.decode_noise_scale = 0.0
    lat = synthetic.latents((1, 128, 2, 16, 24), seed=43)
    ref = V.vae_decode(w_cpu, lat, decoder_blocks=synthetic.DEFAULT_DECODER_BLOCKS, base_channels=128, timestep=0.05)
    out = dec(lat, timestep=0.05)
    assert out.shape == ref.shape == (1, 3, 9, 512, 768)
    assert rel(out, ref) < 3e-2, rel(out, ref)
    assert pearson(out, ref) > 0.999


def test_fp8_linears_19b_width_four_blocks():
    """The FP8 linear path (E4M3 weights kept quantised, per-token E4M3 activations, tcgen05.mma kind::f8f6f4) at the
    benchmarked width: 4 blocks, N = 3456, against the oracle on the DEQUANTISED weights read back from the engine.
    Tolerance: rel L2 <= 6e-2, Pearson >= 0.995 (tests/test_fp8_gpu.py explains the bound)."""
    from ltx2_b200 import synthetic
    from ltx2_b200.loader import iter_engine_weights
    from ltx2_b200.transformer import LTXModel, LTXModelType, Modality
    from oracle import dit_oracle as O
    dev = torch.device("cuda:0")
    cfg = synthetic.DitConfig(num_layers=4)
    m = LTXModel(model_type=LTXModelType.VideoOnly, num_layers=4, device=dev, fp8_linear=True)
    m.load_weights(iter_engine_weights(synthetic.iter_dit_weights(cfg, seed=5, device=dev, dtype=torch.bfloat16), False))
    assert m.missing_weights() == []
    w = {k: m.get_weight(k).cpu() for k in m.weight_keys()}
    lat, ctx, pos = video_inputs(1, 9, 16, 24, 1024, 3840, 410)
    ts = torch.tensor([0.909375])
    ref = O.dit_forward(w, dict(latent=lat, context=ctx, timesteps=ts, positions=pos), num_layers=4, heads=32)
    out = m(Modality(latent=lat, context=ctx, context_mask=None, timesteps=ts, positions=pos))
    r, p = rel(out, ref), pearson(out, ref)
    print(f"fp8 4 blocks @ D=4096, N=3456: rel L2 {r:.3e}, pearson {p:.5f}")
    assert r < 6e-2 and p > 0.995, (r, p)
