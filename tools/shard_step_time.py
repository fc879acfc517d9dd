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
    if N == 432 or N == 3456:
        # per-launch kernel times of one block (events around every launch: ltx2_dit_set_profile)
        import ctypes as C
        from ltx2_b200 import _lib
        L = _lib.lib()
        _lib.check(L.ltx2_dit_set_profile(model._h, 1))
        x0(mod)
        ms, fl, cl = C.c_double(), C.c_double(), C.c_int32()
        recs = []
        i = 0
        while L.ltx2_dit_profile_launch(model._h, i, C.byref(ms), C.byref(fl), C.byref(cl)) == 0:
            recs.append((ms.value * 1e3, fl.value, cl.value))
            i += 1
        _lib.check(L.ltx2_dit_set_profile(model._h, 0))
        per = (len(recs) - 4) // c["layers"] if len(recs) > 100 else len(recs)
        blk = recs[2 + 10 * per: 2 + 11 * per]
        print(f"   {len(recs)} profiled launches; block 10 ({per} launches): " +
              "  ".join(f"{'A' if k == 1 else 'G'} {t:5.1f}us/{f / t / 1e6 if t > 0 else 0:4.0f}TF" for t, f, k in blk), flush=True)
        print(f"   sums: GEMM {sum(t for t, f, k in recs if k != 1) / 1e3:6.2f} ms, attention {sum(t for t, f, k in recs if k == 1) / 1e3:6.2f} ms", flush=True)

# the same step replayed from a CUDA graph: what the launch path costs at this kernel size
for grid in ((1, 18, 24),):
    F, H, W = grid
    N, S = F * H * W, c["S"]
    lat = synthetic.latents((1, N, 128), seed=42).to(dev)
    ctx = synthetic.latents((1, S, c["caption"]), seed=7, std=0.1).to(torch.bfloat16).to(dev)
    pos = synthetic.video_positions(1, F, H, W).to(dev)
    sig = torch.tensor([1.0], device=dev)
    mod = Modality(latent=lat, context=ctx, context_mask=None, timesteps=sig, positions=pos)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            x0(mod)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            out = x0(mod)
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"tokens {N:5d}: {e0.elapsed_time(e1) / 20:7.2f} ms per step replayed from a CUDA graph", flush=True)
