"""__graft_entry__.smoke(): one small invocation of each half of the hot path on cuda:0, checked against the
oracle (the oracle is the checker here, never the thing shipped)."""
from __future__ import annotations

import numpy as np
import torch


def _rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm())


def run() -> None:
    from oracle import dit_oracle as O
    from oracle import vae_oracle as V

    from . import synthetic
    from .loader import load_transformer_state_dict
    from .transformer import LTXModel, LTXModelType, Modality, X0Model
    from .video_vae import SimpleVideoDecoder, decode_latent

    torch.cuda.set_device(0)
    # ---- DiT: 2 blocks, D = 512, N = 72 tokens ----
    cfg = synthetic.DitConfig(num_attention_heads=4, attention_head_dim=128, in_channels=32, out_channels=32,
                              num_layers=2, cross_attention_dim=512, caption_channels=64)
    w = synthetic.dit_weights(cfg, seed=3)
    model = LTXModel(model_type=LTXModelType.VideoOnly, num_attention_heads=4, attention_head_dim=128, in_channels=32,
                     out_channels=32, num_layers=2, cross_attention_dim=512, caption_channels=64)
    load_transformer_state_dict(model, w)
    B, F, H, W, S = 1, 3, 4, 6, 40
    lat = synthetic.latents((B, F * H * W, 32), seed=1)
    ctx = synthetic.latents((B, S, 64), seed=2, std=0.5)
    pos = synthetic.video_positions(B, F, H, W)
    ts = torch.tensor([0.725])
    x0 = X0Model(model)(Modality(latent=lat, context=ctx, context_mask=None, timesteps=ts, positions=pos))
    wr = O.to_engine_keys({k: (v.to(torch.bfloat16).float() if v.ndim == 2 and "scale_shift_table" not in k else v)
                           for k, v in w.items()})
    ref = O.x0_forward(wr, dict(latent=lat, context=ctx, timesteps=ts, positions=pos), num_layers=2, heads=4)
    e = _rel(x0, ref)
    assert e < 2e-2, f"DiT smoke mismatch: rel err {e}"

    # ---- VAE: 64-channel base, 1x128x2x2x2 latent -> 9 frames of 64x64 ----
    blocks = [["res_x", {"num_layers": 1}], ["compress_all", {"multiplier": 2, "residual": True}],
              ["res_x", {"num_layers": 1}], ["compress_all", {"multiplier": 2, "residual": True}],
              ["res_x", {"num_layers": 1}], ["compress_all", {"multiplier": 2, "residual": True}],
              ["res_x", {"num_layers": 1}]]
    vcfg = synthetic.VaeConfig(decoder_blocks=blocks, base_channels=64)
    vw = synthetic.vae_weights(vcfg, seed=4)
    dec = SimpleVideoDecoder(decoder_blocks=blocks, base_channels=64)
    dec.load_weights(vw)
    dec.decode_noise_scale = 0.0
    vlat = synthetic.latents((1, 128, 2, 2, 2), seed=5)
    video = dec(vlat, timestep=0.05)
    vwr = {k: (v.to(torch.bfloat16).float() if v.ndim >= 2 and "scale_shift_table" not in k else v) for k, v in vw.items()}
    vref = V.vae_decode(vwr, vlat, decoder_blocks=blocks, base_channels=64, timestep=0.05)
    e2 = _rel(video, vref)
    assert e2 < 3e-2, f"VAE smoke mismatch: rel err {e2}"
    frames = decode_latent(vlat, dec)
    assert frames.shape == (9, 64, 64, 3) and frames.dtype == torch.uint8
    torch.cuda.synchronize()
    print(f"smoke ok: DiT x0 rel err {e:.2e}, VAE rel err {e2:.2e}")
