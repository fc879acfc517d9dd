#!/usr/bin/env python
"""One denoising step of the bench workload between cudaProfilerStart/Stop, for ncu:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py
    ncu --profile-from-start off --set full --clock-control none --import-source on \
        -k regex:gemm_bf16_kernel -s 8 -c 3 -o gpurun_out/prof_gemm python tools/profile_step.py
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from ltx2_b200 import synthetic  # noqa: E402
from ltx2_b200.loader import iter_engine_weights  # noqa: E402
from ltx2_b200.transformer import LTXModel, LTXModelType, Modality, X0Model  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="19b")
    ap.add_argument("--layers", type=int, default=None, help="override block count (ncu --set full replays are slow)")
    ap.add_argument("--grid", type=int, nargs=3, default=None, metavar=("F", "H", "W"),
                    help="override the latent grid, e.g. 1 18 24 = the 432 tokens of an 8-way context-parallel shard")
    ap.add_argument("--fp8", action="store_true", help="the FP8 linear path (LTXModel(fp8_linear=True))")
    args = ap.parse_args()
    c = dict(bench.CONFIGS[args.config])
    if args.layers:
        c["layers"] = args.layers
    if args.grid:
        c["F"], c["H"], c["W"] = args.grid
    dev = torch.device("cuda:0")
    D = c["heads"] * c["head_dim"]
    cfg = synthetic.DitConfig(num_attention_heads=c["heads"], attention_head_dim=c["head_dim"], num_layers=c["layers"],
                              cross_attention_dim=D, caption_channels=c["caption"])
    model = LTXModel(model_type=LTXModelType.VideoOnly, num_attention_heads=c["heads"], attention_head_dim=c["head_dim"],
                     num_layers=c["layers"], cross_attention_dim=D, caption_channels=c["caption"], device=dev,
                     fp8_linear=args.fp8)
    model.load_weights(iter_engine_weights(synthetic.iter_dit_weights(cfg, seed=0, device=dev, dtype=torch.bfloat16), False))
    N, S = c["F"] * c["H"] * c["W"], c["S"]
    lat = synthetic.latents((1, N, 128), seed=42).to(dev)
    ctx = synthetic.latents((1, S, c["caption"]), seed=7, std=0.1).to(torch.bfloat16).to(dev)
    pos = synthetic.video_positions(1, c["F"], c["H"], c["W"]).to(dev)
    sig = torch.tensor([1.0], device=dev)
    mod = Modality(latent=lat, context=ctx, context_mask=None, timesteps=sig, positions=pos)
    x0 = X0Model(model)
    x0(mod)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    out = x0(mod)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("finite", bool(torch.isfinite(out).all()))


if __name__ == "__main__":
    main()
