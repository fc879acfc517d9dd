#!/usr/bin/env python
"""Run-to-run determinism of the 48-block FP8 forward on one GPU (DET_TOKENS=3456 | 432, DET_RUNS=n), with the default
and with the pinned kernel choice: every run must be bit-identical to the first.  Diagnostics."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from ltx2_b200 import synthetic  # noqa: E402
from ltx2_b200.loader import iter_engine_weights  # noqa: E402
from ltx2_b200.transformer import LTXModel, LTXModelType, Modality, X0Model  # noqa: E402

dev = torch.device("cuda:0")
c = dict(bench.CONFIGS["19b"]); L = 48
D = c["heads"] * c["head_dim"]
cfg = synthetic.DitConfig(num_attention_heads=c["heads"], attention_head_dim=c["head_dim"], num_layers=L, cross_attention_dim=D, caption_channels=c["caption"])
m = LTXModel(model_type=LTXModelType.VideoOnly, num_attention_heads=c["heads"], attention_head_dim=c["head_dim"], num_layers=L, cross_attention_dim=D, caption_channels=c["caption"], device=dev, fp8_linear=True)
m.load_weights(iter_engine_weights(synthetic.iter_dit_weights(cfg, seed=0, device=dev, dtype=torch.bfloat16), False))
F, H, W = (9, 16, 24) if os.environ.get("DET_TOKENS", "3456") == "3456" else (1, 18, 24)
N, S = F*H*W, c["S"]
lat = synthetic.latents((1, N, 128), seed=42).to(dev)
ctx = synthetic.latents((1, S, c["caption"]), seed=7, std=0.1).to(torch.bfloat16).to(dev)
pos = synthetic.video_positions(1, F, H, W).to(dev)
x0 = X0Model(m)
for env in ({}, {"LTX2_GEMM_T": "0", "LTX2_GEMM_2CTA": "0", "LTX2_ATTN_PAIRS": "0"}):
    os.environ.update(env)
    outs = []
    for i in range(int(os.environ.get("DET_RUNS", "6"))):
        m.reset_context_cache()
        mod = Modality(latent=lat, context=ctx, context_mask=None, timesteps=torch.tensor([0.9], device=dev), positions=pos)
        outs.append(x0(mod).clone())
    torch.cuda.synchronize()
    print("env", env, "runs identical to run 0:", [bool(torch.equal(outs[0], o)) for o in outs[1:]], "max diff", max(float((outs[0]-o).abs().max()) for o in outs[1:]), flush=True)
