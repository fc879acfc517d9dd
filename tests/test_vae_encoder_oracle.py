"""Video-VAE encoder oracle vs golden vectors from the reference's own SimpleVideoEncoder (SURVEY.md 8(f) rank 3;
pins the oracle the CUDA encoder path is tested against in tests/test_encoder_upscaler_gpu.py)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    return float((a - b).norm() / b.norm())


def test_encoder_blocks_match_reference_golden():
    from oracle import vae_encoder_oracle as E
    g = np.load(os.path.join(GOLDEN, "vae_encoder.npz"))
    pat = E.patchify(torch.from_numpy(g["patchify_x"]), 4)
    assert torch.equal(pat, torch.from_numpy(g["patchify_y"]))                     # pure data movement: exact
    w = {"d.conv.conv.weight": torch.from_numpy(g["down_w"]), "d.conv.conv.bias": torch.from_numpy(g["down_b"])}
    dy = E.downsample(w, "d", torch.from_numpy(g["down_x"]), 16, (2, 2, 2))
    assert dy.shape == g["down_y"].shape and _rel(dy, torch.from_numpy(g["down_y"])) < 1e-5


def test_encoder_forward_matches_reference_golden():
    from ltx2_b200 import synthetic
    from oracle import vae_encoder_oracle as E
    g = np.load(os.path.join(GOLDEN, "vae_encoder.npz"))
    w = dict(synthetic.iter_vae_encoder_weights(seed=11))
    cs = float(sum(float(v.double().abs().sum()) for v in w.values()))
    assert abs(cs - float(g["weight_checksum"])) <= 1e-6 * cs, "synthetic encoder weights changed; regenerate the golden"
    with torch.no_grad():
        for x, y in (("clip", "latent_clip"), ("image", "latent_image")):
            out = E.vae_encode(w, torch.from_numpy(g[x]))
            ref = torch.from_numpy(g[y])
            assert out.shape == ref.shape
            assert _rel(out, ref) < 2e-3, (x, _rel(out, ref))
    try:
        E.vae_encode(w, torch.zeros(1, 3, 4, 32, 32))
        assert False, "4 frames must be rejected"
    except ValueError as e:
        assert "1 + 8*k frames" in str(e)
