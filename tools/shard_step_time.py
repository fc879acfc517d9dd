#!/usr/bin/env python
"""Step time of the 19B DiT on ONE GPU at a context-parallel rank's token count (432 / 864 / 1728 tokens, all 32 heads,
LTX2_SHARD_SPLIT_K=8): a rank's GEMM and row kernels without the peer traffic, barriers and the 3456-key attention --
the part of an N-GPU step that is local work.  Diagnostics."""
import os
import sys

os.environ.setdefault("LTX2_SHARD_SPLIT_K", "8")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from ltx2_b200 import synthetic  # noqa: E402
from ltx2_b200.loader import iter_engine_weights  # noqa: E402
from ltx2_b200.transformer import LTXModel, LTXModelType, Modality, X0Model  # noqa: E402

c = dict(bench.CONFIGS["19b"])
dev = torch.device("cuda:0")
D = c["heads"] * c["head_dim"]
cfg = synthetic.DitConfig(num_attention_heads=c["heads"], attention_head_dim=c["head_dim"], num_layers=c["layers"],
                          cross_attention_dim=D, caption_channels=c["caption"])
model = LTXModel(model_type=LTXModelType.VideoOnly, num_attention_heads=c["heads"], attention_head_dim=c["head_dim"],
                 num_layers=c["layers"], cross_attention_dim=D, caption_channels=c["caption"], device=dev)
model.load_weights(iter_engine_weights(synthetic.iter_dit_weights(cfg, seed=0, device=dev, dtype=torch.bfloat16), False))
x0 = X0Model(model)
for grid in ((1, 18, 24), (2, 18, 24), (4, 18, 24), (9, 16, 24)):
    F, H, W = grid
    N, S = F * H * W, c["S"]
    lat = synthetic.latents((1, N, 128), seed=42).to(dev)
    ctx = synthetic.latents((1, S, c["caption"]), seed=7, std=0.1).to(torch.bfloat16).to(dev)
    pos = synthetic.video_positions(1, F, H, W).to(dev)
    sig = torch.tensor([1.0], device=dev)
    mod = Modality(latent=lat, context=ctx, context_mask=None, timesteps=sig, positions=pos)
    for _ in range(3):
        x0(mod)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 20
    e0.record()
    for _ in range(steps):
        x0(mod)
    e1.record()
    torch.cuda.synchronize()
    print(f"tokens {N:5d}: {e0.elapsed_time(e1) / steps:7.2f} ms per step", flush=True)
