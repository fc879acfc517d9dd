"""Text-embeddings connector: oracle vs the golden vectors of the reference's own
Embeddings1DConnector (CPU), and the CUDA-backed mirror vs both (GPU).  Tolerance of the CUDA path: bf16 GEMM / attention
operands and bf16 RoPE tables (the reference casts its tables to the activation dtype too, connector.py:262) -> relative
L2 <= 2e-2, Pearson >= 0.999."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm())


def _setup():
    from ltx2_b200 import synthetic
    g = np.load(os.path.join(GOLDEN, "connector.npz"))
    heads, hd, layers, regs = (int(v) for v in g["cfg"])
    w = dict(synthetic.iter_connector_weights(heads=heads, head_dim=hd, layers=layers, registers=regs, gated=True, seed=17))
    cs = float(sum(float(v.double().abs().sum()) for v in w.values()))
    assert abs(cs - float(g["weight_checksum"])) <= 1e-6 * cs, "synthetic connector weights changed; regenerate the golden"
    return g, w, heads, hd, layers, regs


def test_connector_oracle_matches_reference_golden():
    from oracle import connector_oracle as C
    g, w, heads, hd, layers, regs = _setup()
    x = torch.from_numpy(g["x"])
    for name in ("interleaved", "split"):
        xin = x if name == "interleaved" else x[:1]
        y = C.connector(w, xin, heads=heads, layers=layers, rope_type=name)
        ref = torch.from_numpy(g["y_" + name])
        assert y.shape == ref.shape == (xin.shape[0], 1024, heads * hd)
        assert rel(y, ref) < 1e-5, (name, rel(y, ref))


@pytest.mark.gpu
@pytest.mark.parametrize("rope", ["interleaved", "split"])
def test_connector_cuda_matches_reference_golden_and_oracle(rope):
    from ltx2_b200 import synthetic
    from ltx2_b200.text_connector import Embeddings1DConnector
    from ltx2_b200.transformer import LTXRopeType
    from oracle import connector_oracle as C
    g, w, heads, hd, layers, regs = _setup()
    conn = Embeddings1DConnector(attention_head_dim=hd, num_attention_heads=heads, num_layers=layers,
                                 num_learnable_registers=regs, rope_type=LTXRopeType[rope.upper()],
                                 apply_gated_attention=True)
    assert conn.missing_weights()
    # the checkpoint spelling of the same keys is accepted too (loader/weight_converter.py:146-162, 300-313)
    conn.load_weights(("model.diffusion_model.video_embeddings_connector." + k.replace(".to_out.", ".to_out.0."), v)
                      for k, v in w.items())
    assert conn.missing_weights() == []
    x = torch.from_numpy(g["x"])
    xin = x if rope == "interleaved" else x[:1]
    y, mask = conn(xin, torch.zeros(xin.shape[0], 1, 1, xin.shape[1]))
    ref = torch.from_numpy(g["y_" + rope])                     # the reference's own connector, fp32
    assert y.shape == ref.shape and mask.shape == (xin.shape[0], 1, 1, 1024) and float(mask.abs().max()) == 0
    assert rel(y, ref) < 2e-2, rel(y, ref)
    assert float(np.corrcoef(y.cpu().flatten().numpy(), ref.flatten().numpy())[0, 1]) > 0.999
    # production width (30 heads x 128 = 3840, 1024 tokens, batch 2 for the interleaved layout) against the oracle
    wp = dict(synthetic.iter_connector_weights(seed=18, layers=1))
    wr = {k: (v.to(torch.bfloat16).float() if v.ndim == 2 and k != "learnable_registers" else v) for k, v in wp.items()}
    big = Embeddings1DConnector(num_layers=1, rope_type=LTXRopeType[rope.upper()], apply_gated_attention=True)
    big.load_weights(wp)
    xb = synthetic.latents((2 if rope == "interleaved" else 1, 300, 3840), seed=701)
    with torch.no_grad():
        refb = C.connector(wr, xb, heads=30, layers=1, rope_type=rope)
    yb, _ = big(xb)
    assert yb.shape == refb.shape == (xb.shape[0], 1024, 3840)
    assert rel(yb, refb) < 2e-2, rel(yb, refb)
