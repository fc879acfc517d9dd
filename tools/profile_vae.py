#!/usr/bin/env python
"""One 7-latent-frame VAE chunk decode (49 frames @ 512x768) between cudaProfilerStart/Stop, for ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ltx2_b200 import synthetic  # noqa: E402
from ltx2_b200.video_vae import SimpleVideoDecoder  # noqa: E402

dev = torch.device("cuda:0")
dec = SimpleVideoDecoder(device=dev)
dec.load_weights(synthetic.iter_vae_weights(synthetic.VaeConfig(), seed=0, device=dev, dtype=torch.bfloat16))
lat = synthetic.latents((1, 128, 7, 16, 24), seed=43).to(dev)
dec(lat, timestep=0.05)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = dec(lat, timestep=0.05)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("finite", bool(torch.isfinite(out).all()), tuple(out.shape))
