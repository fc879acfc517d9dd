"""CPU-only checks: the C-ABI library builds/loads and exports every symbol include/ltx2_b200.h declares; host-side
logic (key renames, STG bit masks, chunk schedule) matches the reference's semantics.  No compute calls."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ltx2_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ltx2_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from ltx2_b200 import _lib
    L = _lib.lib()
    declared = header_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert sorted(_lib.EXPORTS) == declared, set(_lib.EXPORTS) ^ set(declared)
    assert L.ltx2_version() == 100


def test_product_fails_loudly_without_gpu_or_library():
    from ltx2_b200 import _lib
    from ltx2_b200._lib import Ltx2Error
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ltx2_b200.transformer import LTXModel
    with pytest.raises((Ltx2Error, RuntimeError, AssertionError)):
        LTXModel(num_layers=1, num_attention_heads=2, attention_head_dim=64, cross_attention_dim=128)
    # a missing shared library is an error, never a fallback
    saved, _lib._lib = _lib._lib, None
    path, _lib.LIB_PATH = _lib.LIB_PATH, _lib.LIB_PATH + ".absent"
    try:
        with pytest.raises(Ltx2Error, match="no CPU fallback"):
            _lib.lib()
    finally:
        _lib.LIB_PATH, _lib._lib = path, saved


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ltx-2-mlx_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) and f != "smoke.py":
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
    assert "oracle" not in open(os.path.join(ROOT, "ltx2_b200", "__init__.py")).read()


def test_checkpoint_key_renames_match_reference_rules():
    # weight_converter.py:277-315 and tests/test_loaders.py:33-165 (to_out.0, ff.net.0.proj, ff.net.2, audio filter)
    from ltx2_b200.loader import convert_pytorch_key_to_mlx as conv
    assert conv("transformer_blocks.0.attn1.to_out.0.weight") == "transformer_blocks.0.attn1.to_out.weight"
    assert conv("transformer_blocks.5.ff.net.0.proj.bias") == "transformer_blocks.5.ff.project_in.proj.bias"
    assert conv("transformer_blocks.5.ff.net.2.weight") == "transformer_blocks.5.ff.project_out.weight"
    assert conv("transformer_blocks.1.attn2.to_q.weight") == "transformer_blocks.1.attn2.to_q.weight"
    assert conv("transformer_blocks.1.audio_attn1.to_q.weight") is None
    assert conv("transformer_blocks.1.audio_ff.net.0.proj.weight", include_audio=True) == \
        "transformer_blocks.1.audio_ff.project_in.proj.weight"
    assert conv("transformer_blocks.1.audio_ff.net.2.bias", include_audio=True) == \
        "transformer_blocks.1.audio_ff.project_out.bias"
    assert conv("av_ca_video_scale_shift_adaln_single.linear.weight") is None
    assert conv("video_embeddings_connector.foo", include_audio=True) is None
    assert conv("adaln_single.emb.timestep_embedder.linear_1.weight") == "adaln_single.emb.timestep_embedder.linear_1.weight"
    # the oracle's own rename (independent restatement) agrees
    from oracle.dit_oracle import engine_key
    for k in ("transformer_blocks.0.attn1.to_out.0.weight", "transformer_blocks.5.audio_ff.net.2.weight"):
        assert engine_key("model.diffusion_model." + k) == conv(k, include_audio=True)


def test_every_synthetic_key_maps_to_one_engine_slot():
    from ltx2_b200 import synthetic
    from ltx2_b200.loader import iter_engine_weights
    cfg = synthetic.DitConfig(num_attention_heads=2, attention_head_dim=64, in_channels=32, out_channels=32,
                              num_layers=2, cross_attention_dim=128, caption_channels=None, cross_attention_adaln=True,
                              apply_gated_attention=True, audio=True, audio_heads=2, audio_head_dim=64)
    w = synthetic.dit_weights(cfg)
    keys = [k for k, _ in iter_engine_weights(w.items(), include_audio=True)]
    assert len(keys) == len(set(keys)) == len(w)
    video_only = [k for k, _ in iter_engine_weights(w.items(), include_audio=False)]
    assert all("audio" not in k and "a2v" not in k and "av_ca" not in k for k in video_only)


def test_stg_masks_follow_all_in_batch():
    from ltx2_b200.transformer import (BatchedPerturbationConfig, Perturbation, PerturbationConfig, PerturbationType,
                                       _skip_masks)
    pc = PerturbationConfig([Perturbation(PerturbationType.SKIP_VIDEO_SELF_ATTN, [1, 3]),
                             Perturbation(PerturbationType.SKIP_V2A_CROSS_ATTN, None)])
    sk = _skip_masks(BatchedPerturbationConfig([pc, pc]), 4)
    assert sk.video_self_attn == 0b1010 and sk.v2a_cross_attn == 0b1111 and sk.a2v_cross_attn == 0 and sk.audio_self_attn == 0
    half = _skip_masks(BatchedPerturbationConfig([pc, PerturbationConfig.empty()]), 4)
    assert half.video_self_attn == 0 and half.v2a_cross_attn == 0          # transformer.py:486-501: ALL samples
    assert _skip_masks(None, 4) is None
    assert BatchedPerturbationConfig.empty(2).all_in_batch(PerturbationType.SKIP_A2V_CROSS_ATTN, 0) is False


def test_chunk_schedule_and_positions_match_oracle():
    from ltx2_b200.video_vae import chunk_plan
    from oracle import vae_oracle as V
    for T in range(1, 40):
        assert chunk_plan(T) == V.chunk_plan(T)
        plan = chunk_plan(T)
        assert plan[0][0] == 0 and plan[-1][1] == T and all(b - a <= 7 for a, b in plan)
    from ltx2_b200 import synthetic
    pos = synthetic.video_positions(1, 3, 2, 2, fps=None)
    # causal_fix: first frame spans [0,1), later frames [8f-7, 8f+1) (patchifiers.py:228-238)
    assert pos[0, 0, 0].tolist() == [0.0, 1.0] and pos[0, 0, 4].tolist() == [1.0, 9.0] and pos[0, 0, 8].tolist() == [9.0, 17.0]
    assert pos[0, 1, 2].tolist() == [32.0, 64.0] and pos[0, 2, 1].tolist() == [32.0, 64.0]


def test_tile_specs_and_masks_match_reference_semantics():
    # tiling.py:154-249: default config on a 16x24x31 latent (768x512x241) and the error messages of :55-103
    from ltx2_b200.tiling import (SpatialTilingConfig, TemporalTilingConfig, TilingConfig, compute_trapezoidal_mask_1d,
                                  generate_tile_specs)
    from oracle import vae_oracle as V
    specs = generate_tile_specs((1, 128, 31, 16, 24), TilingConfig.default())
    t_tiles = sorted({(s.in_t_start, s.in_t_end) for s in specs})
    w_tiles = sorted({(s.in_w_start, s.in_w_end) for s in specs})
    assert t_tiles == [(a, b) for a, b, _, _ in V.tiles_1d(31, 8, 3)] and t_tiles[-1] == (23, 31)
    assert w_tiles == [(a, b) for a, b, _, _ in V.tiles_1d(24, 16, 2)] and sorted({(s.in_h_start, s.in_h_end) for s in specs}) == [(0, 16)]
    first, last = specs[0], specs[-1]
    assert (first.out_t_start, first.out_t_end) == (0, 57) and first.ramp_t_left == 0 and first.ramp_t_right == 24
    assert last.out_t_end == 241 and last.ramp_w_left == 64 and last.ramp_w_right == 0
    for (l, a, b, z) in [(16, 4, 4, False), (16, 4, 0, True), (8, 0, 3, False), (5, 8, 8, False)]:
        assert torch.allclose(compute_trapezoidal_mask_1d(l, a, b, z), V.trapezoid_mask_1d(l, a, b, z))
    with pytest.raises(ValueError, match="at least 64"):
        SpatialTilingConfig(32)
    with pytest.raises(ValueError, match="divisible by 8"):
        TemporalTilingConfig(20)
    with pytest.raises(ValueError, match="Overlap must be less"):
        SpatialTilingConfig(64, 64)
    with pytest.raises(ValueError, match="positive"):
        compute_trapezoidal_mask_1d(0, 0, 0)
