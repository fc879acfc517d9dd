"""ltx2_b200 -- Blackwell-native drop-in for the LTX-2-MLX sampling hot path.

Only what the hot path needs lives here:
  csrc/          hand-written sm_100a CUDA kernels + the C-ABI (include/ltx2_b200.h)
  _lib.py        ctypes binding of the C-ABI shared library (fails loudly if it is missing)
  transformer.py LTXModel / X0Model / Modality mirror (reference: model/transformer/model.py)
  video_vae.py   SimpleVideoDecoder / decode_latent mirror (reference: model/video_vae/simple_decoder.py)
  kernels.py     silu_mul / gelu_mul / interleaved_rope (reference: kernels/fused_ops.py)
  sampling.py    fused denoise-step update: CFG guide + masked blend + Euler step (reference: components/diffusion_steps.py,
                 components/guiders.py, pipelines/common.py) and the device-resident distilled loop
  tiling.py      decode_tiled mirror with native accumulate / normalise (reference: model/video_vae/tiling.py)
  context_parallel.py  token-sharded DiT over the GPUs of one box
  loader.py      safetensors key mapping (reference: loader/weight_converter.py, simple_decoder.py:566)
  synthetic.py   seeded synthetic checkpoints for tests and the bench
"""
__version__ = "0.1.0"
