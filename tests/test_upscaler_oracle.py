"""Latent spatial upscaler oracle vs a golden vector from the reference's own SpatialUpscaler (SURVEY.md 8(f) rank 3;
pins the oracle the CUDA upscaler path is tested against in tests/test_encoder_upscaler_gpu.py)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_upscaler_matches_reference_golden():
    from ltx2_b200 import synthetic
    from oracle import upscaler_oracle as U
    g = np.load(os.path.join(GOLDEN, "upscaler.npz"))
    cin, mid, groups, blocks = (int(v) for v in g["cfg"])
    w = dict(synthetic.iter_upscaler_weights(seed=13, in_channels=cin, mid_channels=mid, blocks=blocks))
    cs = float(sum(float(v.double().abs().sum()) for v in w.values()))
    assert abs(cs - float(g["weight_checksum"])) <= 1e-6 * cs, "synthetic upscaler weights changed; regenerate the golden"
    with torch.no_grad():
        out = U.upscale(w, torch.from_numpy(g["latent"]), groups=groups, blocks=blocks)
    ref = torch.from_numpy(g["upscaled"])
    assert out.shape == ref.shape == (1, cin, 3, 8, 10)
    assert float((out - ref).norm() / ref.norm()) < 1e-4
