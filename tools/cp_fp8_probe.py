#!/usr/bin/env python
"""Context parallelism with the FP8 linears at the 19B width: sharded vs un-sharded forward with the kernel choice pinned,
block by block (set_layer_limit), for token grids that give a rank 432 / 1728 rows.  torchrun, >= 2 GPUs.  Diagnostics."""
import os
import sys

for k, v in {"LTX2_GEMM_T": "0", "LTX2_GEMM_2CTA": "0", "LTX2_ATTN_PAIRS": "0", "LTX2_ATTN_SPLIT": "0"}.items():
    os.environ[k] = v
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from ltx2_b200 import context_parallel, synthetic  # noqa: E402
from ltx2_b200.loader import iter_engine_weights  # noqa: E402
from ltx2_b200.transformer import LTXModel, LTXModelType, Modality, X0Model  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
c = dict(bench.CONFIGS["19b"])
L = int(os.environ.get("PROBE_LAYERS", "4"))
D = c["heads"] * c["head_dim"]
cfg = synthetic.DitConfig(num_attention_heads=c["heads"], attention_head_dim=c["head_dim"], num_layers=L,
                          cross_attention_dim=D, caption_channels=c["caption"])
for fp8 in ((True,) if os.environ.get("PROBE_FP8_ONLY") else (True, False)):
    per_rank = int(os.environ.get("PROBE_ROWS_PER_RANK", "432"))
    for grid in (((per_rank * world // (18 * 24), 18, 24),) if os.environ.get("PROBE_FP8_ONLY") else ((432 * world // (18 * 24), 18, 24), (1728 * world // (18 * 24), 18, 24))):
        F, H, W = grid
        N, S = F * H * W, c["S"]
        models = []
        for _ in range(2):
            m = LTXModel(model_type=LTXModelType.VideoOnly, num_attention_heads=c["heads"], attention_head_dim=c["head_dim"],
                         num_layers=L, cross_attention_dim=D, caption_channels=c["caption"], device=dev, fp8_linear=fp8)
            m.load_weights(iter_engine_weights(synthetic.iter_dit_weights(cfg, seed=0, device=dev, dtype=torch.bfloat16), False))
            models.append(m)
        single, sharded = models
        context_parallel.enable(sharded, batch=1, n_total=N, context_tokens=S)
        context_parallel.set_split_k(sharded, 1)
        lat = synthetic.latents((1, N, 128), seed=42).to(dev)
        ctx = synthetic.latents((1, S, c["caption"]), seed=7, std=0.1).to(torch.bfloat16).to(dev)
        pos = synthetic.video_positions(1, F, H, W).to(dev)
        sig = torch.tensor([0.9], device=dev)
        res = []
        for lim in range(1, L + 1, int(os.environ.get("PROBE_LAYER_STEP", "1"))):
            outs = []
            for m in (single, sharded):
                m.set_layer_limit(lim)
                m.reset_context_cache()
                mod = Modality(latent=lat, context=ctx, context_mask=None, timesteps=sig, positions=pos)
                outs.append(X0Model(m)(mod).clone())
            torch.cuda.synchronize()
            res.append((lim, bool(torch.equal(outs[0], outs[1])), float((outs[0] - outs[1]).abs().max())))
        t = torch.tensor([[0.0 if e else 1.0, d] for _, e, d in res], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"fp8={fp8} tokens {N} ({N // world} per rank), blocks 1..{L}: " +
                  "  ".join(f"{lim}:{'exact' if t[i, 0] == 0 else f'{float(t[i, 1]):.1e}'}" for i, (lim, _, _) in enumerate(res)),
                  flush=True)
        context_parallel.disable(sharded)
        del single, sharded, models
        torch.cuda.empty_cache()
dist.destroy_process_group()
