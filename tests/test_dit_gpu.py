"""GPU parity of the full DiT forward (through LTXModel -> C ABI) against the oracle on the same seeded
synthetic checkpoint and inputs.  The oracle runs in fp32 on the bf16-rounded weights, so the remaining
difference is bf16 activations at the GEMM/attention inputs (fp32 accumulation, fp32 residual stream).

Tolerances: relative L2 error <= 2e-2 and the reference's own metric, Pearson r (tests/test_parity.py:53-59),
>= 0.999 (the reference's gate is r >= 0.95, test_parity.py:38)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def pearson(a, b):
    a, b = a.double().flatten().cpu().numpy(), b.double().flatten().cpu().numpy()
    return float(np.corrcoef(a, b)[0, 1])


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm())


def bf16_round(w):
    return {k: (v.to(torch.bfloat16).float() if v.ndim == 2 and "scale_shift_table" not in k else v) for k, v in w.items()}


def build(cfg, seed):
    from ltx2_b200 import synthetic
    from ltx2_b200.loader import load_transformer_state_dict
    from ltx2_b200.transformer import LTXModel, LTXModelType
    from oracle import dit_oracle as O

    class Small(LTXModel):
        AUDIO_ATTENTION_HEADS = cfg.audio_heads
        AUDIO_HEAD_DIM = cfg.audio_head_dim

    w = synthetic.dit_weights(cfg, seed=seed)
    m = Small(model_type=LTXModelType.AudioVideo if cfg.audio else LTXModelType.VideoOnly,
              num_attention_heads=cfg.num_attention_heads, attention_head_dim=cfg.attention_head_dim,
              in_channels=cfg.in_channels, out_channels=cfg.out_channels, num_layers=cfg.num_layers,
              cross_attention_dim=cfg.cross_attention_dim, caption_channels=cfg.caption_channels,
              cross_attention_adaln=cfg.cross_attention_adaln, apply_gated_attention=cfg.apply_gated_attention,
              av_ca_timestep_scale_multiplier=1000)
    load_transformer_state_dict(m, w)
    assert m.missing_weights() == []
    return m, O.to_engine_keys(bf16_round(w))


def video_inputs(cfg, B, F, H, W, S, seed, ctx_dim):
    from ltx2_b200 import synthetic
    N = F * H * W
    lat = synthetic.latents((B, N, cfg.in_channels), seed=seed)
    ctx = synthetic.latents((B, S, ctx_dim), seed=seed + 1, std=0.5)
    pos = synthetic.video_positions(B, F, H, W, fps=24.0)
    return lat, ctx, pos


@pytest.mark.parametrize("mode", ["scalar", "pertoken"])
def test_dit_v1_forward_matches_oracle(mode):
    from ltx2_b200 import synthetic
    from ltx2_b200.transformer import Modality, X0Model
    from oracle import dit_oracle as O
    cfg = synthetic.DitConfig(num_attention_heads=4, attention_head_dim=128, in_channels=32, out_channels=32,
                              num_layers=2, cross_attention_dim=512, caption_channels=64)
    m, w = build(cfg, seed=11)
    B, F, H, W, S = 2, 3, 4, 6, 40
    lat, ctx, pos = video_inputs(cfg, B, F, H, W, S, 100, 64)
    N = F * H * W
    if mode == "scalar":
        ts = torch.tensor([0.9, 0.4])
    else:
        ts = torch.where(torch.arange(N)[None, :, None] < H * W, torch.zeros(1), torch.tensor([0.725, 0.25])[:, None, None])
    ref = O.dit_forward(w, dict(latent=lat, context=ctx, timesteps=ts, positions=pos), num_layers=2, heads=4)
    mod = Modality(latent=lat, context=ctx, context_mask=None, timesteps=ts, positions=pos)
    out = m(mod)
    assert out.shape == ref.shape and out.dtype == torch.float32 and out.is_cuda
    assert rel(out, ref) < 2e-2, rel(out, ref)
    assert pearson(out, ref) > 0.999
    x0 = X0Model(m)(mod)
    assert rel(x0, O.to_x0(lat, ts, ref)) < 2e-2
    # numpy inputs and bf16 inputs are accepted like the reference's mx.array
    out_np = m(Modality(latent=lat.numpy(), context=ctx.numpy(), context_mask=None, timesteps=ts.numpy(),
                        positions=pos.numpy()))
    assert torch.equal(out_np, out)


def test_dit_v1_cross_attn_scale_and_errors():
    from ltx2_b200 import synthetic
    from ltx2_b200._lib import Ltx2Error
    from ltx2_b200.transformer import LTXModel, Modality
    from oracle import dit_oracle as O
    cfg = synthetic.DitConfig(num_attention_heads=2, attention_head_dim=64, in_channels=32, out_channels=32,
                              num_layers=1, cross_attention_dim=128, caption_channels=64)
    m, w = build(cfg, seed=12)
    lat, ctx, pos = video_inputs(cfg, 1, 2, 3, 4, 24, 110, 64)
    ts = torch.tensor([0.6])
    mod = Modality(latent=lat, context=ctx, context_mask=None, timesteps=ts, positions=pos)
    base = m(mod)
    m.transformer_blocks[0]._cross_attn_scale = 0.0
    off = m(mod)
    del m.transformer_blocks[0]._cross_attn_scale
    assert torch.equal(m(mod), base) and not torch.allclose(off, base)
    with pytest.raises(ValueError, match="Video modality required"):
        m(None)
    fresh = LTXModel(num_attention_heads=2, attention_head_dim=64, in_channels=32, out_channels=32, num_layers=1,
                     cross_attention_dim=128, caption_channels=64)
    with pytest.raises(Ltx2Error, match="have not been set"):
        fresh(mod)


def test_dit_v2_audio_video_and_stg_match_oracle():
    from ltx2_b200 import synthetic
    from ltx2_b200.transformer import (BatchedPerturbationConfig, Modality, Perturbation, PerturbationConfig,
                                       PerturbationType, X0Model)
    from oracle import dit_oracle as O
    cfg = synthetic.DitConfig(num_attention_heads=4, attention_head_dim=128, in_channels=32, out_channels=32,
                              num_layers=2, cross_attention_dim=512, caption_channels=None,
                              cross_attention_adaln=True, apply_gated_attention=True, audio=True,
                              audio_heads=4, audio_head_dim=64)
    m, w = build(cfg, seed=13)
    B, F, H, W, S, Na = 2, 3, 4, 6, 40, 9
    lat, ctx, pos = video_inputs(cfg, B, F, H, W, S, 120, 512)
    alat = synthetic.latents((B, Na, 128), seed=130)
    actx = synthetic.latents((B, S, 256), seed=131, std=0.5)
    apos = synthetic.audio_positions(B, Na)
    sv, sa = torch.tensor([0.9, 0.4]), torch.tensor([0.8, 0.3])
    kw = dict(num_layers=2, heads=4, audio_heads=4, v2=True, av_ca_timestep_scale_multiplier=1000)
    vd = dict(latent=lat, context=ctx, timesteps=sv, positions=pos, sigma=sv)
    ad = dict(latent=alat, context=actx, timesteps=sa, positions=apos, sigma=sa)
    rv, ra = O.dit_forward(w, vd, ad, **kw)
    vm = Modality(latent=lat, context=ctx, context_mask=None, timesteps=sv, positions=pos, sigma=sv)
    am = Modality(latent=alat, context=actx, context_mask=None, timesteps=sa, positions=apos, sigma=sa)
    ov, oa = m(vm, am)
    assert rel(ov, rv) < 2e-2 and rel(oa, ra) < 2e-2, (rel(ov, rv), rel(oa, ra))
    assert pearson(ov, rv) > 0.999 and pearson(oa, ra) > 0.999
    # video-only inference on the AV model
    v_only, a_empty = m(vm, None)
    assert a_empty.shape == (B, 0, 128)
    assert rel(v_only, O.dit_forward(w, vd, None, **kw)) < 2e-2
    assert X0Model(m)(vm, None).shape == lat.shape
    # STG: skip video self-attention in block 1 and a2v in block 0 for the whole batch
    pc = PerturbationConfig([Perturbation(PerturbationType.SKIP_VIDEO_SELF_ATTN, [1]),
                             Perturbation(PerturbationType.SKIP_A2V_CROSS_ATTN, [0])])
    pv, pa = m(vm, am, perturbations=BatchedPerturbationConfig([pc, pc]))
    sv_, sa_ = O.dit_forward(w, vd, ad, skip_blocks={"video_self": [1], "a2v": [0]}, **kw)
    assert rel(pv, sv_) < 2e-2 and rel(pa, sa_) < 2e-2
    # a perturbation that only applies to half the batch is NOT applied (all_in_batch, transformer.py:486-501)
    half = BatchedPerturbationConfig([pc, PerturbationConfig.empty()])
    hv, _ = m(vm, am, perturbations=half)
    assert torch.equal(hv, ov)


def test_weight_readback_and_lora_style_update():
    """parameters()/load_weights round trip used by the LoRA fuse/restore flows (generate.py:1198-1200)."""
    from ltx2_b200 import synthetic
    from ltx2_b200.transformer import Modality
    cfg = synthetic.DitConfig(num_attention_heads=2, attention_head_dim=64, in_channels=32, out_channels=32,
                              num_layers=1, cross_attention_dim=128, caption_channels=64)
    m, w = build(cfg, seed=14)
    params = m.parameters()
    keys = list(params.keys())
    assert "transformer_blocks.0.attn1.to_out.weight" in keys and len(keys) == len(w)
    k = "transformer_blocks.0.ff.project_in.proj.weight"
    base = params[k]
    assert base.shape == (512, 128) and torch.equal(base.cpu(), w[k].to(torch.bfloat16).float())
    assert torch.equal(params["transformer_blocks.0.scale_shift_table"].cpu(), w["transformer_blocks.0.scale_shift_table"])
    lat, ctx, pos = video_inputs(cfg, 1, 2, 3, 4, 24, 140, 64)
    mod = Modality(latent=lat, context=ctx, context_mask=None, timesteps=torch.tensor([0.5]), positions=pos)
    before = m(mod)
    delta = torch.randn_like(base) * 0.05
    m.load_weights([(k, base + delta)])                      # fuse
    assert not torch.allclose(m(mod), before)
    m.load_weights([(k, base)])                              # restore
    assert torch.equal(m(mod), before)
    with pytest.raises(KeyError):
        m.get_weight("no.such.key")


def test_timestep_classes_equal_per_token_form_and_context_cache_is_safe():
    """Engine extensions behind the reference call surface: (1) Modality.timestep_classes (pre-computed (batch, sigma)
    classes, no host round trip) is bit-identical to the per-token (B, T) timesteps it stands for; (2) the V1 text
    K/V reuse keyed on the context tensor's identity + version: same object again -> same bits; in-place change, new
    tensor, or a weight update -> recomputed."""
    from ltx2_b200 import sampling, synthetic
    from ltx2_b200.transformer import Modality
    cfg = synthetic.DitConfig(num_attention_heads=4, attention_head_dim=128, in_channels=32, out_channels=32,
                              num_layers=2, cross_attention_dim=512, caption_channels=64)
    m, w = build(cfg, seed=15)
    B, F, H, W, S = 2, 3, 4, 6, 40
    N = F * H * W
    lat, ctx, pos = video_inputs(cfg, B, F, H, W, S, 150, 64)
    dev = torch.device("cuda:0")
    ctx_d = ctx.to(dev)
    mask = torch.ones(B, N, device=dev)
    mask[:, :H * W] = 0.0
    mask[1, H * W:2 * H * W] = 0.5
    sig = torch.tensor([0.725, 0.725], device=dev)
    vals, rows = sampling.timestep_classes_from_mask(mask)
    assert vals.numel() == 5 and rows.shape == (B, N)
    assert torch.equal(vals[rows.long()], mask)
    per_tok = m(Modality(latent=lat, context=ctx_d, context_mask=None, timesteps=mask * 0.725, positions=pos, sigma=sig))
    by_cls = m(Modality(latent=lat, context=ctx_d, context_mask=None, timesteps=sig, positions=pos, sigma=sig,
                        timestep_classes=(vals * 0.725, rows)))
    assert torch.equal(per_tok, by_cls)

    # ---- context K/V reuse ----
    mod = Modality(latent=lat, context=ctx_d, context_mask=None, timesteps=torch.tensor([0.9, 0.4]), positions=pos)
    first = m(mod)
    assert torch.equal(m(mod), first)                       # hit
    m.reuse_context = False
    assert torch.equal(m(mod), first)                       # recomputed: same bits as the cached path
    m.reuse_context = True
    m(mod)
    ctx_d.mul_(0.5)                                         # in-place change bumps the version -> miss
    changed = m(mod)
    assert not torch.equal(changed, first)
    m.reuse_context = False
    assert torch.equal(m(mod), changed)
    m.reuse_context = True
    ctx2 = (ctx_d * 2.0).contiguous()                       # the original values in a new tensor
    again = m(Modality(latent=lat, context=ctx2, context_mask=None, timesteps=torch.tensor([0.9, 0.4]), positions=pos))
    assert torch.equal(again, first)
    # a weight update of the text K projection invalidates the cached K/V
    k = "transformer_blocks.0.attn2.to_k.weight"
    base = m.get_weight(k)
    mod2 = Modality(latent=lat, context=ctx2, context_mask=None, timesteps=torch.tensor([0.9, 0.4]), positions=pos)
    m.load_weights([(k, base * 1.5)])
    assert not torch.equal(m(mod2), first)
    m.load_weights([(k, base)])
    assert torch.equal(m(mod2), first)


def test_load_transformer_weights_from_safetensors_file(tmp_path):
    """load_transformer_weights(model, path) (loader/weight_converter.py:318-326 call surface) on a temporary safetensors
    file with the reference's checkpoint key names, including an FP8-style tensor with its weight_scale
    (loader/fp8_loader.py:14-32: weight * weight_scale)."""
    from safetensors.torch import save_file
    from ltx2_b200 import synthetic
    from ltx2_b200.loader import load_transformer_weights
    from ltx2_b200.transformer import LTXModel, Modality
    cfg = synthetic.DitConfig(num_attention_heads=2, attention_head_dim=64, in_channels=32, out_channels=32,
                              num_layers=1, cross_attention_dim=128, caption_channels=64)
    w = synthetic.dit_weights(cfg, seed=16)
    sd = {k: v.to(torch.bfloat16).contiguous() for k, v in w.items()}
    fp8_key = "model.diffusion_model.transformer_blocks.0.ff.net.0.proj.weight"
    scale = 0.03125
    sd[fp8_key] = (w[fp8_key] / scale).to(torch.float8_e4m3fn)
    sd[fp8_key.replace(".weight", ".weight_scale")] = torch.tensor(scale)
    sd["vae.decoder.conv_in.conv.bias"] = torch.zeros(4)                 # foreign tensors are ignored
    path = str(tmp_path / "ckpt.safetensors")
    save_file(sd, path)
    kw = dict(num_attention_heads=2, attention_head_dim=64, in_channels=32, out_channels=32, num_layers=1,
              cross_attention_dim=128, caption_channels=64)
    m2 = LTXModel(**kw)
    load_transformer_weights(m2, path, strict=True, use_fp8=True)
    assert m2.missing_weights() == []
    deq = sd[fp8_key].float() * scale
    got = m2.get_weight("transformer_blocks.0.ff.project_in.proj.weight").cpu()
    assert torch.equal(got, deq.to(torch.bfloat16).float())
    # same forward as a model loaded from the same (bf16 / dequantised FP8) values held in memory
    from ltx2_b200.loader import load_transformer_state_dict
    m3 = LTXModel(**kw)
    mem = {k: v for k, v in sd.items() if not k.endswith(".weight_scale")}
    mem[fp8_key] = deq
    load_transformer_state_dict(m3, mem)
    lat, ctx, pos = video_inputs(cfg, 1, 2, 3, 4, 24, 160, 64)
    mod = Modality(latent=lat, context=ctx, context_mask=None, timesteps=torch.tensor([0.5]), positions=pos)
    assert torch.equal(m2(mod), m3(mod))


def test_profile_records_per_launch_and_per_class():
    """ltx2_dit_set_profile brackets every GEMM / attention launch with events; ltx2_dit_profile_launch returns them one
    by one (launch order) and their sums are what ltx2_dit_profile_read reports per class."""
    import ctypes as C
    from ltx2_b200 import _lib, synthetic
    from ltx2_b200.transformer import Modality

    cfg = synthetic.DitConfig(num_attention_heads=4, attention_head_dim=128, in_channels=32, out_channels=32, num_layers=2,
                              cross_attention_dim=512, caption_channels=64)
    m, _ = build(cfg, seed=5)
    lat, ctx, pos = video_inputs(cfg, 1, 2, 4, 6, 24, seed=50, ctx_dim=64)
    mod = Modality(latent=lat, context=ctx, context_mask=None, timesteps=torch.tensor([0.7]), positions=pos)
    L = _lib.lib()
    _lib.check(L.ltx2_dit_set_profile(m._h, 1))
    m(mod)
    ms, fl, ln = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_int64 * 3)()
    _lib.check(L.ltx2_dit_profile_read(m._h, ms, fl, ln, 3))
    one_ms, one_fl, one_cls = C.c_double(), C.c_double(), C.c_int32()
    sums, counts, i = [0.0, 0.0, 0.0], [0, 0, 0], 0
    while L.ltx2_dit_profile_launch(m._h, i, C.byref(one_ms), C.byref(one_fl), C.byref(one_cls)) == 0:
        assert one_ms.value > 0 and one_fl.value > 0 and 0 <= one_cls.value < 3
        sums[one_cls.value] += one_ms.value
        counts[one_cls.value] += 1
        i += 1
    _lib.check(L.ltx2_dit_set_profile(m._h, 0))
    assert i == sum(ln) and counts == list(ln)
    assert counts[1] == 2 * cfg.num_layers                 # self + text attention per block
    for c in range(3):
        assert abs(sums[c] - ms[c]) <= 1e-6 * max(1.0, ms[c])
