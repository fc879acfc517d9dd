import sys, torch
sys.path.insert(0, "/root/repo")
from ltx2_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
for K in (4096, 16384):
    for N in (4096, 12288, 16384):
        a = torch.randn(3456, K, device=dev)
        w = torch.randn(N, K, device=dev) * K ** -0.5
        a8, as_ = ops.quantize_rows_e4m3(a)
        w8, ws = ops.quantize_rows_e4m3(w)
        for mode in (ops.EPI_BF16, ops.EPI_F32):
            full = ops.gemm_e4m3(a8, as_, w8, ws, None, mode=mode)
            for M in (432, 864, 1728):
                part = ops.gemm_e4m3(a8[:M].contiguous(), as_[:M].contiguous(), w8, ws, None, mode=mode)
                same = torch.equal(full[:M], part)
                d = float((full[:M].float() - part.float()).abs().max())
                print(f"K={K} N={N} mode={mode} M={M}: bit-equal to rows of the M=3456 launch: {same}  max diff {d:.3e}", flush=True)
