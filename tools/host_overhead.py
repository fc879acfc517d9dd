#!/usr/bin/env python
"""Host enqueue time vs GPU time of one forward at a context-parallel-shard-sized problem (432 tokens, 19B model) on one
GPU: tells whether small problems are bound by the CPU launch path (diagnostics)."""
import os
import sys
import time

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
for F, H, W in ((1, 18, 24), (9, 16, 24)):
    N, S = F * H * W, c["S"]
    lat = synthetic.latents((1, N, 128), seed=42).to(dev)
    ctx = (0.1 * synthetic.latents((1, S, c["caption"]), seed=43)).to(dev).to(torch.bfloat16)
    pos = synthetic.video_positions(1, F, H, W, fps=24.0).to(dev)
    sig = torch.full((1,), 0.9, device=dev)
    mod = Modality(latent=lat, context=ctx, context_mask=None, timesteps=sig.reshape(1, 1), positions=pos, sigma=sig)
    for _ in range(3):
        x0(mod)
    torch.cuda.synchronize()
    host, gpu = [], []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        x0(mod)
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        host.append((t1 - t0) * 1e3)
        gpu.append(e0.elapsed_time(e1))
    print(f"N={N}: host enqueue {min(host):.2f} ms, GPU {min(gpu):.2f} ms per forward", flush=True)
