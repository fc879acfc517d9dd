#!/usr/bin/env python
"""Does any kernel read memory it (or the engine) never wrote?  Build the model in a fresh process, run; free it; fill
most of the device memory with NaN bit patterns, free that; build the same model again (its cudaMalloc'd buffers now
hold garbage), run: the outputs must be bit-identical.  Diagnostics."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from ltx2_b200 import synthetic  # noqa: E402
from ltx2_b200.loader import iter_engine_weights  # noqa: E402
from ltx2_b200.transformer import LTXModel, LTXModelType, Modality, X0Model  # noqa: E402

dev = torch.device("cuda:0")
c = dict(bench.CONFIGS["19b"])
L = int(os.environ.get("PROBE_LAYERS", "6"))
D = c["heads"] * c["head_dim"]
cfg = synthetic.DitConfig(num_attention_heads=c["heads"], attention_head_dim=c["head_dim"], num_layers=L,
                          cross_attention_dim=D, caption_channels=c["caption"])


def run(fp8, N_grid):
    F, H, W = N_grid
    m = LTXModel(model_type=LTXModelType.VideoOnly, num_attention_heads=c["heads"], attention_head_dim=c["head_dim"],
                 num_layers=L, cross_attention_dim=D, caption_channels=c["caption"], device=dev, fp8_linear=fp8)
    m.load_weights(iter_engine_weights(synthetic.iter_dit_weights(cfg, seed=0, device=dev, dtype=torch.bfloat16), False))
    N, S = F * H * W, c["S"]
    lat = synthetic.latents((1, N, 128), seed=42).to(dev)
    ctx = synthetic.latents((1, S, c["caption"]), seed=7, std=0.1).to(torch.bfloat16).to(dev)
    pos = synthetic.video_positions(1, F, H, W).to(dev)
    mod = Modality(latent=lat, context=ctx, context_mask=None, timesteps=torch.tensor([0.9], device=dev), positions=pos)
    out = X0Model(m)(mod).clone()
    torch.cuda.synchronize()
    del m
    torch.cuda.empty_cache()
    return out


def dirty():
    free, _ = torch.cuda.mem_get_info()
    n = int(free * 0.9) // 4
    junk = torch.full((n,), float("nan"), device=dev)
    torch.cuda.synchronize()
    del junk
    torch.cuda.empty_cache()


for fp8 in (False, True):
    for grid in ((9, 16, 24), (1, 18, 24)):
        a = run(fp8, grid)
        dirty()
        b = run(fp8, grid)
        print(f"fp8={fp8} tokens {grid[0] * grid[1] * grid[2]}: fresh vs dirty device memory: "
              f"{'bit-identical' if torch.equal(a, b) else f'DIFFERENT max {float((a - b).abs().max()):.3e} finite {bool(torch.isfinite(b).all())}'}",
              flush=True)
