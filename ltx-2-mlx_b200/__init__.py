"""ltx2_b200 -- Blackwell-native drop-in for the LTX-2-MLX sampling hot path.

Only what the hot path needs lives here:
  csrc/          hand-written sm_100a CUDA kernels + the C-ABI (include/ltx2_b200.h)
  _lib.py        ctypes binding of the C-ABI shared library (fails loudly if it is missing)
  transformer.py LTXModel / X0Model / Modality mirror (reference: model/transformer/model.py)
  video_vae.py   SimpleVideoDecoder / decode_latent mirror (reference: model/video_vae/simple_decoder.py)
  kernels.py     silu_mul / gelu_mul / interleaved_rope (reference: kernels/fused_ops.py)
  loader.py      safetensors key mapping (reference: loader/weight_converter.py, simple_decoder.py:566)
  synthetic.py   seeded synthetic checkpoints for tests and the bench
"""
__version__ = "0.1.0"
