import os, sys
sys.path.insert(0, "/root/repo")
import torch
import bench
from ltx2_b200 import synthetic
from ltx2_b200.loader import iter_engine_weights
from ltx2_b200.transformer import LTXModel, LTXModelType, Modality, X0Model
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
