/*
 * ltx2_b200 -- C ABI of the Blackwell-native LTX-2 denoising hot path.
 *
 * The reference (Acelogic/LTX-2-MLX) has no FFI: its boundary is Python duck typing
 * (SURVEY.md section 8(b)).  This header is what a binding for that boundary links against:
 * every entry point names the reference interface it replaces (paths relative to the
 * reference repository root).  INTEGRATION.md shows the ctypes stub the reference would add.
 *
 * Conventions
 *   - all data pointers are DEVICE pointers unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); every call is
 *     stream-ordered and returns without synchronising;
 *   - return value: 0 on success, negative errno-style code otherwise (LTX2_ERR_* below);
 *     ltx2_last_error() returns a thread-local message for the last failure;
 *   - no exceptions, no torch types, no hidden global state besides opaque handles;
 *   - handles are not thread-safe (one stream, one caller), like the reference's modules.
 *   - dtype codes: 0 = float32, 1 = bfloat16, 2 = float16.
 */
#ifndef LTX2_B200_H_
#define LTX2_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LTX2_OK 0
#define LTX2_ERR_INVALID (-22)
#define LTX2_ERR_NOMEM (-12)
#define LTX2_ERR_CUDA (-5)
#define LTX2_ERR_NOKEY (-2)
#define LTX2_ERR_STATE (-1)

#define LTX2_DTYPE_F32 0
#define LTX2_DTYPE_BF16 1
#define LTX2_DTYPE_F16 2
#define LTX2_DTYPE_F8E4M3 3 /* only as the source dtype of ltx2_dit_set_weight_scaled / the operands of ltx2_gemm_e4m3 */

int ltx2_version(void);
const char* ltx2_last_error(void);

/* =====================================================================================
 * DiT engine  --  replaces LTX_2_MLX/model/transformer/model.py: LTXModel (:413-881) and
 * X0Model (:884-936), i.e. the object every pipeline calls as `transformer(video_modality
 * [, audio_modality], perturbations=...)` (pipelines/distilled.py:229-239, one_stage.py:274).
 * ===================================================================================== */

typedef struct LtxDit LtxDit;

/* LTXModel.__init__ arguments (model.py:436-461) that shape the network. */
typedef struct LtxDitConfig {
  int32_t num_attention_heads;      /* 32 */
  int32_t attention_head_dim;       /* 128 (64 or 128 supported) */
  int32_t in_channels;              /* 128 */
  int32_t out_channels;             /* 128 */
  int32_t num_layers;               /* 48 */
  int32_t cross_attention_dim;      /* 4096 */
  int32_t caption_channels;         /* 3840; 0 = no caption projection (V2) */
  int32_t cross_attention_adaln;    /* V2: 9-row adaLN + prompt KV modulation */
  int32_t apply_gated_attention;    /* V2: per-head 2*sigmoid gate */
  int32_t audio_enabled;            /* LTXModelType.AudioVideo */
  int32_t audio_heads;              /* 32 (model.py:428) */
  int32_t audio_head_dim;           /* 64 (model.py:429) */
  int32_t audio_in_channels;        /* 128 */
  int32_t audio_out_channels;       /* 128 */
  float norm_eps;                   /* 1e-6 */
  float positional_embedding_theta; /* 10000 */
  float max_pos[3];                 /* {20, 2048, 2048} */
  float audio_max_pos;              /* 20 (model.py:434) */
  float timestep_scale_multiplier;  /* 1000 */
  float av_ca_timestep_scale_multiplier; /* 1 (model.py:452); the CLI passes 1000 */
  /* Engine option (no reference counterpart; the reference widens FP8 checkpoints to bf16/fp16 at load,
   * loader/fp8_loader.py:14-130): keep the linears that are fed by a norm kernel -- self-attention QKV, text-attention Q,
   * FFN up-projection -- as E4M3 bytes with their weight_scale, quantise their input rows to E4M3 with a per-token
   * dynamic scale inside the adaLN/RMSNorm kernel, and run them on tcgen05.mma kind::f8f6f4 (twice the bf16 tensor
   * rate); both scales are applied to the fp32 accumulator in the epilogue.  All other linears stay bf16. */
  int32_t fp8_linear;
} LtxDitConfig;

/* One `Modality` (model.py:59-69), flattened.  latent/context dtype per *_dtype. */
typedef struct LtxModalityView {
  const void* latent;        /* [B, N, C_in] */
  int32_t latent_dtype;
  const void* context;       /* [B, S, C_ctx] */
  int32_t context_dtype;
  const float* timesteps;    /* [B * n_t] fp32: n_t == 1 (Modality.timesteps (B,)) or n_t == N ((B,N) / (B,N,1)) */
  const float* sigma;        /* [B] fp32 or NULL (Modality.sigma) */
  const float* positions;    /* [B, n_dims, N, 2] fp32 [start,end) bounds */
  int32_t batch;
  int32_t tokens;            /* N */
  int32_t context_tokens;    /* S */
  int32_t n_t;               /* 1 or N */
  int32_t n_dims;            /* 3 video, 1 audio */
  /* Optional pre-computed timestep classes (n_cls > 0): `timesteps` then holds the n_cls DISTINCT (batch, sigma) values
   * and row_cls[B*N] maps every token row to its class.  The per-token form (n_t == N) makes the engine read the
   * timesteps back and de-duplicate them on the host every call; a sampling loop whose denoise mask is fixed builds
   * the map once and only rescales the class values per step (ltx-2-mlx_b200/sampling.py), so the forward never
   * synchronises with the host.  Same result as timesteps[b,t] = class_value[row_cls[b,t]]. */
  const int32_t* row_cls;    /* [B * N] or NULL */
  int32_t n_cls;             /* 0 = not given; <= 64 */
} LtxModalityView;

int ltx2_dit_create(const LtxDitConfig* cfg, LtxDit** out);
void ltx2_dit_destroy(LtxDit* dit);

/* Replaces load_transformer_weights -> model.update (loader/weight_converter.py:318-446).
 * `key` is the reference's MLX-side name, i.e. the checkpoint key minus "model.diffusion_model."
 * after the renames of weight_converter.py:300-313 (e.g. "transformer_blocks.0.attn1.to_out.weight").
 * The tensor is copied (and converted to the engine's storage type) before the call returns
 * control of `data` to the caller on `stream`. */
int ltx2_dit_set_weight(LtxDit* dit, const char* key, const void* data, int32_t dtype, const int64_t* shape,
                        int32_t ndim, void* stream);
/* An FP8 checkpoint tensor (dtype LTX2_DTYPE_F8E4M3, value = e4m3 * weight_scale; loader/fp8_loader.py:14-32).  With
 * cfg.fp8_linear the FP8-computed linears keep these bytes as they are; every other Linear weight is widened to bf16 on
 * the device.  Float dtypes are accepted with weight_scale == 1 (same as ltx2_dit_set_weight). */
int ltx2_dit_set_weight_scaled(LtxDit* dit, const char* key, const void* data, int32_t dtype, const int64_t* shape,
                               int32_t ndim, float weight_scale, void* stream);
/* Read-back of the flat {reference_key: tensor} view the LoRA fuse/restore code needs (`velocity_model.parameters()`,
 * scripts/generate.py:1198-1200; pipelines/two_stage.py:180-186): list of keys, shape of one key (returns ndim), and
 * a copy of its values converted to dst_dtype (matrices are stored as bf16, so that is their precision). */
int64_t ltx2_dit_weight_keys(LtxDit* dit, char* names_out, int64_t names_cap);
int ltx2_dit_weight_shape(LtxDit* dit, const char* key, int64_t* shape_out /* [2] */);
int ltx2_dit_get_weight(LtxDit* dit, const char* key, void* dst, int32_t dst_dtype, int64_t n, void* stream);
/* number of weight tensors still missing; forward refuses to run until it is 0 */
int ltx2_dit_missing_weights(LtxDit* dit, char* names_out, int64_t names_cap);

/* LTXModel.__call__ (model.py:776-881): velocity [B,N,C_out] fp32 (and audio velocity).
 * skip_* : per-block bit masks of the STG perturbations that apply to the WHOLE batch
 * (BatchedPerturbationConfig.all_in_batch, transformer.py:486-501); bit i = block i; NULL = none.
 * When x0 != 0 the outputs are denoised samples latent - t*velocity (X0Model, model.py:895-936). */
typedef struct LtxDitSkip {
  uint64_t video_self_attn;
  uint64_t audio_self_attn;
  uint64_t a2v_cross_attn;
  uint64_t v2a_cross_attn;
} LtxDitSkip;

int ltx2_dit_forward(LtxDit* dit, const LtxModalityView* video, const LtxModalityView* audio /* nullable */,
                     const LtxDitSkip* skip /* nullable */, int32_t x0, float* out_video, float* out_audio /* nullable */,
                     void* stream);

/* Text-context reuse across denoising steps.  In the V1 model the text context is sigma-independent
 * (transformer.py:427-455: K/V come from the caption-projected context without modulation), so the caption projection
 * and the 48 text K/V projections + k-norm are identical for every step of a sample.  A caller that passes the SAME
 * context contents again may say so with a non-zero tag: when the tag, batch and S match the previous forward, the
 * engine reuses the cached projected K/V (per block: K [B,H,S,Dh] after k-norm, V rows) instead of recomputing them.
 * tag 0 (default) = never reuse.  Ignored for cross_attention_adaln models (their K/V depend on sigma).  The host layer
 * derives the tag from the identity and version counter of the caller's context tensor. */
int ltx2_dit_set_context_tag(LtxDit* dit, uint64_t tag);

/* Diagnostics (bench.py parity leg): run only the first `n` transformer blocks (then the output head); n <= 0 or
 * n >= num_layers restores the full model. */
int ltx2_dit_set_layer_limit(LtxDit* dit, int32_t n);

/* OneStagePipeline pokes block._cross_attn_scale (one_stage.py:207-222; transformer.py:526-528). */
int ltx2_dit_set_cross_attn_scale(LtxDit* dit, int32_t block, float scale /* NaN = unset */);

/* Measurement hooks (bench.py): per-class CUDA-event timing of one forward (class 0 = bf16 GEMM launches,
 * 1 = attention launches, 2 = FP8 GEMM launches -- counted in class 0 when n_classes == 2) and the number of kernels
 * this library has launched since load. */
int ltx2_dit_set_profile(LtxDit* dit, int32_t on);
int ltx2_dit_profile_read(LtxDit* dit, double* ms_out, double* flops_out, int64_t* launches_out, int32_t n_classes);
/* Profiled launch i (launch order) of the last forward: kernel time, algorithmic FLOPs, class (0 GEMM, 1 attention,
 * 2 FP8 GEMM); LTX2_ERR_INVALID past the last record. */
int ltx2_dit_profile_launch(LtxDit* dit, int32_t i, double* ms_out, double* flops_out, int32_t* class_out);
int64_t ltx2_launch_count(void);

/* Context parallelism over the token axis (no reference counterpart -- the reference is
 * single-device).  Rank r owns tokens [r*N/P, (r+1)*N/P) and, inside self-attention, heads [r*H/P, (r+1)*H/P).
 * The two re-shards per block are fused into the producing kernels as stores to peer memory over NVLink (q/k-norm+
 * RoPE kernel: token->head; attention epilogue: head->token) and ordered by flag barriers in peer memory.
 *   1. every rank: ltx2_dit_cp_init(...) allocates its exchange region and returns a 64-byte CUDA IPC handle;
 *   2. the host layer all-gathers the handles (any transport) and calls ltx2_dit_cp_connect on every rank,
 *      followed by a host barrier;
 *   3. ltx2_dit_forward is then called with the LOCAL token slice (tokens = N/P) on every rank, collectively.
 * When context_tokens (S) is given and S % P == 0, the text-context K/V projection of every block is also sharded:
 * each rank projects S/P context rows and its GEMM epilogue stores them into all ranks' buffers.
 * The audio+video (V2.3) model is supported as well: the audio stream is replicated, the v2a keys/values are projected
 * per token slice and broadcast to all ranks.  Limits: N % P == 0, H % P == 0, P <= 8. */
int ltx2_dit_cp_init(LtxDit* dit, int32_t rank, int32_t world, int32_t batch, int32_t n_total, int32_t context_tokens,
                     char* handle_out);
int ltx2_dit_cp_connect(LtxDit* dit, const char* handles);
/* Orderly teardown: phase 0 on every rank (close the imported peer mappings), a host barrier, then phase 1 (free the own
 * exchange region; the engine is single-GPU again).  An exported allocation must not be freed while a peer maps it. */
int ltx2_dit_cp_shutdown(LtxDit* dit, int32_t phase);
/* Split-K cap of the residual GEMMs on sharded ranks (their M = N/P rows do not fill the SMs otherwise): 1 = off -- the
 * sharded forward is then BIT-IDENTICAL to the single-GPU forward (bench.py cp_parity, tests/test_cp_gpu.py) --,
 * 0 = default (8, or the LTX2_CP_SPLIT_K environment variable). */
int ltx2_dit_cp_set_split_k(LtxDit* dit, int32_t max_splits);

/* =====================================================================================
 * Video-VAE decoder engine -- replaces LTX_2_MLX/model/video_vae/simple_decoder.py:
 * SimpleVideoDecoder (:364-563), load_vae_decoder_weights (:566-673) and the device work of
 * decode_latent (:676-800).  Called by pipelines as `decode_latent(latent, decoder[, timestep])`
 * (pipelines/distilled.py:496, scripts/generate.py:2085) and `decoder_fn(tile, timestep=...)`
 * (model/video_vae/tiling.py:363).
 * ===================================================================================== */

typedef struct LtxVae LtxVae;

/* one entry of the reversed decoder_blocks list (simple_decoder.py:403-427) */
typedef struct LtxVaeStage {
  int32_t kind;        /* 0 = "res_x" group, 1 = "compress_*" depth-to-space upsample */
  int32_t num_layers;  /* res_x: number of ResBlock3d */
  int32_t stride_t, stride_h, stride_w; /* upsample: (2,2,2) all, (2,1,1) time, (1,2,2) space */
  int32_t multiplier;  /* upsample: out_channels_reduction_factor */
  int32_t residual;    /* upsample: add depth-to-space residual */
} LtxVaeStage;

typedef struct LtxVaeConfig {
  int32_t num_stages;
  LtxVaeStage stages[16];          /* execution order (latent -> pixels) */
  int32_t base_channels;           /* 128; feature channels start at 8x (multiple of 64 required) */
  int32_t latent_channels;         /* 128 */
  int32_t timestep_conditioning;   /* V2.0-style decoders */
} LtxVaeConfig;

int ltx2_vae_create(const LtxVaeConfig* cfg, LtxVae** out);
void ltx2_vae_destroy(LtxVae* vae);
/* `key` is the CHECKPOINT key the reference loader reads, e.g. "vae.decoder.up_blocks.0.res_blocks.1.conv1.conv.weight"
 * (simple_decoder.py:592-671).  Conv weights arrive in PyTorch layout [C_out, C_in, 3, 3, 3]. */
int ltx2_vae_set_weight(LtxVae* vae, const char* key, const void* data, int32_t dtype, const int64_t* shape,
                        int32_t ndim, void* stream);
int ltx2_vae_missing_weights(LtxVae* vae, char* names_out, int64_t names_cap);
int ltx2_vae_output_shape(LtxVae* vae, const int64_t in_shape[5], int64_t out_shape[5]);

/* SimpleVideoDecoder.__call__ (:446-563).  latent [B,128,T,H,W] (NCDHW, dtype code) -> out [B,3,T',32H,32W] fp32.
 * timestep < 0 means None.  noise (fp32, latent-shaped, N(0,1)) is blended as noise*s + (1-s)*x when
 * timestep conditioning is on (:496-498); pass NULL or s = 0 for the deterministic decode the parity test uses. */
int ltx2_vae_decode(LtxVae* vae, const void* latent, int32_t dtype, const int64_t shape[5], float timestep,
                    float noise_scale, const float* noise, int32_t causal, float* out, void* stream);

/* Temporal shards of ONE decode over the GPUs of an NVLink box (no reference counterpart -- the
 * reference is single-device).  Every rank holds the whole latent and the decoder weights and computes a contiguous
 * range of frames of every activation: latent frames are split evenly over the first min(world, T) ranks, and frame t
 * of a stage becomes frames 2t-1, 2t of the next, so ownership follows the depth-to-space upsampling without any
 * re-distribution.  A 3x3x3 conv needs one frame from each neighbour: the padded conv inputs live in a CUDA-IPC exchange
 * region; before every conv a rank stores its first / last frame into its neighbours' pad slots (peer-memory stores over
 * NVLink) and all ranks pass a flag barrier.  No collective library on the data path; results are bit-identical to
 * ltx2_vae_decode.  Non-causal decode only.
 *   1. every rank: ltx2_vae_cp_init(rank, world, largest latent shape) -> 64-byte CUDA IPC handle;
 *   2. the host layer all-gathers the handles, calls ltx2_vae_cp_connect on every rank, then a host barrier;
 *   3. ltx2_vae_decode_sharded, collectively on the ranks of a GROUP [group_first, group_first + group_size) (the whole
 *      world, or an aligned equal-size part of it: e.g. the two halves of 8 ranks decode two temporal chunks of
 *      decode_latent at the same time), with the SAME latent on every rank of the group: conv_out's epilogue writes each
 *      rank's frames at their place in the clip and the spans are stored into clip slot `slot` of rank `dst` (or of every
 *      rank, dst = -1) over NVLink; ltx2_vae_shard_frames tells which output frames a rank computes;
 *   4. ltx2_vae_cp_collect, on ALL ranks: a barrier over the whole world, then the receiving ranks copy the assembled
 *      clip [B,3,T',32H,32W] fp32 of that slot into `out`;
 *   5. teardown: ltx2_vae_cp_shutdown(0) on every rank, host barrier, ltx2_vae_cp_shutdown(1). */
int ltx2_vae_cp_init(LtxVae* vae, int32_t rank, int32_t world, const int64_t max_latent_shape[5], char* handle_out);
int ltx2_vae_cp_connect(LtxVae* vae, const char* handles);
int ltx2_vae_cp_shutdown(LtxVae* vae, int32_t phase);
int ltx2_vae_shard_frames(LtxVae* vae, int64_t latent_frames, int32_t rank, int32_t world, int64_t* t0, int64_t* tn);
int ltx2_vae_decode_sharded(LtxVae* vae, const void* latent, int32_t dtype, const int64_t shape[5], float timestep,
                            float noise_scale, const float* noise, int32_t group_first, int32_t group_size, int32_t slot,
                            int32_t dst, void* stream);
int ltx2_vae_cp_collect(LtxVae* vae, int32_t slot, const int64_t clip_shape[5], int32_t dst, float* out, void* stream);

/* Conv3dSimple.__call__ (simple_decoder.py:90-180) as one op, for unit parity of the implicit-GEMM conv kernel at
 * production shapes: x [B,T,H,W,Cin] bf16 channels-last, weight in PyTorch layout [Cout,Cin,3,3,3] and bias [Cout]
 * (dtype codes), out [B,T,H,W,Cout] bf16.  Padding as in the decoder: reflect H/W, replicate T (causal: frame 0 twice
 * in front, nothing behind).  workspace: >= ltx2_conv3d_workspace_bytes(...) bytes of device memory (padded input,
 * packed weight, packed bias).  C_in % 64 == 0, C_out % 8 == 0. */
int64_t ltx2_conv3d_workspace_bytes(int32_t B, int32_t T, int32_t H, int32_t W, int32_t Cin, int32_t Cout);
int ltx2_conv3d(const void* x, const void* weight, int32_t w_dtype, const void* bias, int32_t b_dtype, void* out,
                int32_t B, int32_t T, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t causal, void* workspace,
                void* stream);

/* Measurement hooks (bench.py): CUDA-event timing of the conv launches of one decode. */
int ltx2_vae_set_profile(LtxVae* vae, int32_t on);
int ltx2_vae_profile_read(LtxVae* vae, double* ms_out, double* flops_out, int64_t* launches_out);
int ltx2_vae_profile_launch(LtxVae* vae, int32_t i, double* ms_out, double* flops_out);   /* conv launch i of that decode */

/* decode_latent's chunk stitching (:749-790): dst [BC,T_dst,HW] <- cross-fade of src [BC,T_src,HW] placed at frame t0,
 * linear ramp over the first `overlap` frames, plain copy after; and the uint8 conversion (:793-798):
 * video [1,3,T,H,W] fp32 in [-1,1] -> uint8 [T,H,W,3]. */
int ltx2_blend_chunk(float* dst, const float* src, int32_t BC, int32_t T_dst, int32_t T_src, int32_t HW, int32_t t0,
                     int32_t overlap, void* stream);
int ltx2_video_to_uint8(const float* video, uint8_t* out, int32_t T, int32_t H, int32_t W, void* stream);

/* decode_tiled's blend (model/video_vae/tiling.py:354-412): out [BC,To,Ho,Wo] += tile[:, :tt, :th, :tw] * mask_t x mask_h x
 * mask_w placed at (t0,h0,w0) (tile has pitch dt,dh,dw), wsum [To,Ho,Wo] += mask; then out /= max(wsum, 1e-8). */
int ltx2_tile_accumulate(float* out, float* wsum, const float* tile, int32_t BC, int32_t To, int32_t Ho, int32_t Wo,
                         int32_t dt, int32_t dh, int32_t dw, int32_t t0, int32_t h0, int32_t w0, int32_t tt, int32_t th,
                         int32_t tw, const float* mask_t, const float* mask_h, const float* mask_w, void* stream);
int ltx2_tile_normalize(float* out, const float* wsum, int32_t BC, int64_t plane, void* stream);

/* =====================================================================================
 * Video-VAE ENCODER and 2x latent SPATIAL UPSCALER building blocks.  Both are 3x3x3 conv stacks:
 * the convs run on the same tcgen05 implicit-GEMM kernel as the decoder (ltx2_conv3d_packed); the ops below are what
 * differs.  The host mirrors (ltx-2-mlx_b200/video_vae_encoder.py: SimpleVideoEncoder, upscaler.py: SpatialUpscaler)
 * sequence them like model/video_vae/simple_encoder.py:306-405 and model/upscaler/spatial.py:376-412.
 * Activations: channels-last bf16 [B,T,H,W,C].
 * ===================================================================================== */

/* x -> padded [B, TL+2, H+2, W+2, C] (TL = T, or T+1 with dup_first: the first frame duplicated in front,
 * simple_encoder.py:237-240), with an optional activation applied on the way:
 *   hw_mode 0 reflect (decoder) / 1 zero (encoder :62-74, upscaler :57-66);
 *   t_mode  0 replicate 1+1 / 1 causal: first frame twice in front (simple_encoder.py:76-84) / 2 zero 1+1 (upscaler);
 *   act     0 none / 1 pixel-norm + SiLU (simple_encoder.py:12-15,145-151) / 2 GroupNorm affine / 3 GroupNorm + SiLU
 *           (spatial.py:160-181: `residual` is added after the norm, before the SiLU);
 *   gn_stats [B,groups,2] = (mean, rstd) from ltx2_group_stats.  out_plain (optional) receives the un-padded result. */
int ltx2_pad_act(const void* x, void* out_padded, void* out_plain, int32_t B, int32_t T, int32_t H, int32_t W, int32_t C,
                 int32_t hw_mode, int32_t t_mode, int32_t dup_first, int32_t act, const float* gn_stats,
                 const float* gn_weight, const float* gn_bias, int32_t groups, float eps, const void* residual,
                 void* stream);
/* group_norm_5d statistics (spatial.py:91-128): over (C/groups, T, H, W) per (batch, group) */
int ltx2_group_stats(const void* x, int32_t B, int64_t thw, int32_t C, int32_t groups, float eps, float* stats,
                     void* stream);
/* conv weight [Cout,Cin,3,3,3] + bias -> the kernel's packed form (bf16 [Cout_pad, 27*Cin], fp32 [Cout_pad]) */
int ltx2_conv3d_pack(const void* weight, int32_t w_dtype, const void* bias, int32_t b_dtype, int32_t Cout,
                     int32_t Cout_pad, int32_t Cin, void* w_packed, float* b_packed, void* stream);
/* 3x3x3 conv on a padded input [B,T+2,H+2,W+2,Cin] -> out [B,T,H,W,Cout] (+ residual [B,T,H,W,Cout] if given) */
int ltx2_conv3d_packed(const void* x_padded, const void* w_packed, const float* b_packed, void* out, const void* residual,
                       int32_t B, int32_t T, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t Cout_pad,
                       void* stream);
/* patchify (ops.py:44-68): video fp32 [B,3,F,H,W] -> bf16 [B,F,H/4,W/4,Cp], channel (c*4 + r_w)*4 + r_h, zeros >= 48 */
int ltx2_patchify_video(const float* video, void* out, int32_t B, int32_t F, int32_t H, int32_t W, int32_t Cp,
                        void* stream);
/* SpaceToDepthDownsample3d tail (simple_encoder.py:207-257): space_to_depth(conv output y) + group-mean of
 * space_to_depth(x); y [B,TL,H,W,Cout/sp], x [B,T,H,W,Cx], out [B,TL/st,H/sh,W/sw,Cout] */
int ltx2_space_to_depth_residual(const void* y, const void* x, void* out, int32_t B, int32_t T, int32_t H, int32_t W,
                                 int32_t Cx, int32_t Cout, int32_t st, int32_t sh, int32_t sw, int32_t dup_first,
                                 void* stream);
/* PixelShuffle2d (spatial.py:184-218): y [BF,H,W,4C] -> out [BF,2H,2W,C], source channel c*4 + r_h*2 + r_w */
int ltx2_pixel_shuffle2(const void* y, void* out, int64_t BF, int32_t H, int32_t W, int32_t C, void* stream);
/* layout changes at the network boundaries: channels-last bf16 [B,thw,Cs] -> fp32 NCDHW (first C channels, optional
 * (x - mean) / std: PerChannelStatistics.normalize, ops.py:173-186) and NCDHW (dtype code) -> channels-last bf16 */
int ltx2_ndhwc_to_ncdhw(const void* x, float* out, int32_t B, int32_t C, int32_t Cs, int64_t thw, const float* mean,
                        const float* stdv, void* stream);
int ltx2_ncdhw_to_ndhwc(const void* x, int32_t dtype, void* out, int32_t B, int32_t C, int64_t thw, void* stream);

/* =====================================================================================
 * Per-op entry points (unit parity against the oracle)
 * ===================================================================================== */

/* nn.Linear + fused epilogue (attention.py:190-201, feed_forward.py:23-49).
 * C[M,N] = A[M,K] W[N,K]^T; mode 0: bf16 out = acc+bias; 1: bf16 gelu_tanh(acc+bias);
 * 2: f32 out = acc+bias; 3: f32 out += alpha * gate[row_cls[row], col] * (acc+bias). */
int ltx2_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int32_t M, int32_t N, int32_t K,
                   int32_t mode, const float* bias, void* out, int64_t ldo, const float* gate, int64_t gate_stride,
                   const int32_t* row_cls, float alpha, void* stream);

/* The same GEMM with E4M3 operands (A8 [M,K], W8 [N,K] bytes) on tcgen05.mma kind::f8f6f4:
 * C = (A8 W8^T) * row_scale[m] * col_scale[n] (+ bias, epilogue modes 0..2 as above).  K % 16 == 0. */
int ltx2_gemm_e4m3(const void* A8, int64_t lda, const void* W8, int64_t ldw, int32_t M, int32_t N, int32_t K,
                   int32_t mode, const float* bias, const float* row_scale, const float* col_scale, void* out,
                   int64_t ldo, void* stream);
/* FP8 operand preparation: norm_modulate with per-row E4M3 quantisation (out8 [M,D] bytes, row_scale [M] = absmax/448;
 * out_bf16 optional), and a weight matrix [rows,K] (dtype code) -> E4M3 with one scale per row. */
int ltx2_norm_modulate_q8(const void* x, int32_t x_dtype, int64_t ldx, void* out8, int64_t ldo8, float* row_scale,
                          void* out_bf16, int64_t ldo16, int32_t M, int32_t D, int32_t norm_kind, float eps,
                          const float* mod, int64_t mod_stride, int64_t shift_off, int64_t scale_off,
                          const int32_t* row_cls, void* stream);
int ltx2_quantize_rows_e4m3(const void* w, int32_t dtype, int64_t rows, int64_t K, void* out8, float* row_scale,
                            void* stream);

/* Mode 3 with split-K allowed (up to max_splits K slices accumulate into `out` with vector reductions).  Used by the
 * context-parallel ranks, whose M = N/P rows do not fill the SMs otherwise; accumulation order is not deterministic. */
int ltx2_gemm_bf16_splitk(const void* A, int64_t lda, const void* W, int64_t ldw, int32_t M, int32_t N, int32_t K,
                          const float* bias, float* out, int64_t ldo, const float* gate, int64_t gate_stride,
                          const int32_t* row_cls, float alpha, int32_t max_splits, void* stream);

/* mx.fast.scaled_dot_product_attention + head merge + V2 gate (attention.py:12-34, 243-250).
 * q [B,H,Tq,Dh], k [B,H,Tk,Dh], vt [B,H,Dh,Tkp] (V transposed), out [B,Tq,H*Dh]; all bf16. */
int ltx2_attention(const void* q, const void* k, const void* vt, void* out, int32_t B, int32_t H, int32_t Tq,
                   int32_t Tk, int32_t Tkp, int32_t Dh, float scale, const float* gate_logits, float* lse_out,
                   void* stream);

/* Same attention with V in row form (no transpose pass): element (b,h,t,d) at v + b*stride_b + h*stride_h +
 * t*stride_t + d, e.g. the V third of a fused QKV projection output (stride_t = 3*inner, stride_h = Dh). */
int ltx2_attention_vrows(const void* q, const void* k, const void* v, int64_t v_stride_t, int64_t v_stride_h,
                         int64_t v_stride_b, void* out, int32_t B, int32_t H, int32_t Tq, int32_t Tk, int32_t Dh,
                         float scale, const float* gate_logits, float* lse_out, void* stream);

/* Diagnostics: ltx2_attention plus a clock64 timeline of CTA 0.  head_dim 128 (two-stream kernel): trace must hold
 * 16*(key tiles + 1) words; trace[t*16 + e] are the events of key tile t (tools/attn_bench.py names them) and the last
 * 16 words hold the CTA phases (entry, set-up done, key loop done, epilogue done).  With LTX2_ATTN_KERNEL=single or
 * head_dim 64: trace[j*8 + e], e = 0 QK_j issued, 1 P_j seen by the MMA thread, 2 PV_j issued, 3 S_j seen by softmax,
 * 4 S_j in registers, 5 exps done, 6 PV_{j-1} retired, 7 P_j published (tools/attn_trace.py). */
int ltx2_attention_trace(const void* q, const void* k, const void* vt, void* out, int32_t B, int32_t H, int32_t Tq,
                         int32_t Tk, int32_t Tkp, int32_t Dh, float scale, long long* trace, void* stream);

/* _compiled_adaln_forward / rms_norm / LayerNorm+modulate (transformer.py:16-31, model.py:744-758).
 * norm_kind 0 none, 1 RMS, 2 LayerNorm(no affine). */
int ltx2_norm_modulate(const void* x, int32_t x_dtype, int64_t ldx, void* out_bf16, int64_t ldo, int32_t M, int32_t D,
                       int32_t norm_kind, float eps, const float* mod, int64_t mod_stride, int64_t shift_off,
                       int64_t scale_off, const int32_t* row_cls, void* stream);

/* q_norm/k_norm + apply_split_rotary_emb + head split (attention.py:231-237, rope.py:92-144). */
int ltx2_headnorm_rope(const void* in_bf16, int64_t ld, const float* weight, const float* cos, const float* sin,
                       void* out_bf16, int32_t B, int32_t T, int32_t H, int32_t Dh, float eps, void* stream);
int ltx2_v_transpose(const void* v_bf16, int64_t ld, void* vt_bf16, int32_t B, int32_t T, int32_t Tp, int32_t H,
                     int32_t Dh, void* stream);

/* precompute_freqs_cis(rope_type=SPLIT, use_middle_indices_grid=True) (rope.py:365-418).
 * cos/sin out: [B, T, dim/2] fp32 token-major (== reference (B,H,T,dim/2/H) transposed back). */
int ltx2_rope_tables(const float* positions, int32_t B, int32_t n_dims, int32_t T, int32_t dim,
                     const float* max_pos_host, float theta, float* cos, float* sin, void* stream);

/* AdaLayerNormSingle pieces (timestep_embedding.py:10-60, 166-202). */
int ltx2_timestep_sinusoid(const float* t, int32_t R, float multiplier, float* out256, void* stream);
int ltx2_small_linear(const float* x, int32_t R, int32_t K, const void* W_bf16, const float* bias, float* y, int32_t N,
                      int32_t act_in, void* stream);

/* X0Model.denoise (model.py:912-918) */
int ltx2_x0_from_velocity(const float* latent, const float* velocity, const float* t_row, float* x0, int32_t M,
                          int32_t C, void* stream);

/* Diagnostics: ltx2_attention_vrows plus a clock64 timeline of CTA 0.  SM-pair kernel (head_dim 128, Tq > 128): trace
 * must hold 16 * ceil(Tk / 128) words; per key block k, trace[16k + e]: e = 0 P(k) seen by the MMA thread, 1 P(k)V and
 * S(k+2) issued, 2 / 3 / 4 S(k) seen / exponentials done / P(k) published by the softmax warp of keys [0,64), 5 / 6 / 7
 * the same for keys [64,128) (tools/attn2_check.py prints it). */
int ltx2_attention_vrows_trace(const void* q, const void* k, const void* v, int64_t v_stride_t, int64_t v_stride_h,
                               int64_t v_stride_b, void* out, int32_t B, int32_t H, int32_t Tq, int32_t Tk, int32_t Dh,
                               float scale, long long* trace, void* stream);

/* Host-side planners (no GPU needed; 148 SMs are assumed when no device is visible).  They expose which kernel and
 * tiling a launch will take, so the scheduling logic is testable on a CPU-only machine.
 *   ltx2_gemm_plan: out6 = {kernel (0 standard 128-token-row tiles, 1 transposed 128-weight-row tiles, 2 SM-pair tiles),
 *                           bn, K splits, token tile width, last token tile width, token tiles}
 *   ltx2_attention_plan: pair work items per (batch, head) slice of the two-stream attention kernel (the remaining
 *                           128-query tiles run as split-KV items) and the resulting CTA count. */
int ltx2_gemm_plan(int32_t M, int32_t N, int32_t K, int32_t mode, int32_t max_splits, int32_t* out6);
int ltx2_attention_plan(int32_t Tq, int32_t BH, int32_t* pairs_per_slice, int32_t* n_ctas);
/*   ltx2_attention_sm_pair_plan: the SM-pair attention kernel's persistent grid -- clusters (SM pairs) launched and
 *                           whether the (work item, key block) space is cut into equal ranges per cluster whose partial
 *                           results are merged (split = 1, stream-K) or whole items go round-robin (split = 0). */
int ltx2_attention_sm_pair_plan(int32_t Tq, int32_t Tk, int32_t BH, int32_t* n_clusters, int32_t* split);
/*   ltx2_attention_sm_pair_segments: the segments one cluster of that grid walks, 7 ints each {item, first key block,
 *                           end key block, part, parts, scratch slot, scratch slot of part 0 (-1: whole item)};
 *                           returns the segment count (>= 0) or a negative status. */
int ltx2_attention_sm_pair_segments(int32_t Tq, int32_t Tk, int32_t BH, int32_t cluster, int32_t* out7,
                                    int32_t max_segments);

/* The elementwise tail of one denoising step of the reference's host loops (pipelines/distilled.py:243-251,
 * pipelines/one_stage.py:284-320), fused into one pass over fp32 [M, C] tensors:
 *   CFGGuider.guide (components/guiders.py:40-44)            d = cond + (cfg_scale - 1)(cond - uncond)   [uncond_x0 != NULL]
 *   post_process_latent (pipelines/common.py:169-190)        d = d * mask[row] + clean * (1 - mask[row]) [mask, clean != NULL]
 *   EulerDiffusionStep.step (components/diffusion_steps.py:55-67, to_velocity core_utils.py:34-62)
 *                                                            out = sample + (sample - d) / sigma * (sigma_next - sigma)
 * denoised_out (optional) receives d.  sigma == 0 is LTX2_ERR_INVALID ("Sigma can't be 0.0", core_utils.py:54-55). */
int ltx2_denoise_update(const float* sample, const float* cond_x0, const float* uncond_x0, float cfg_scale,
                        const float* denoise_mask, const float* clean_latent, float sigma, float sigma_next, float* out,
                        float* denoised_out, int32_t M, int32_t C, void* stream);

/* The reference's own three kernels (kernels/fused_ops.py:50, 95, 183). */
int ltx2_silu_mul(const void* a, const void* b, void* out, int64_t n, int32_t dtype, void* stream);
int ltx2_gelu_mul(const void* a, const void* b, void* out, int64_t n, int32_t dtype, void* stream);
int ltx2_interleaved_rope(const void* x, const void* cos, const void* sin, void* out, int64_t n, int32_t dtype,
                          void* stream);

int ltx2_cast(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LTX2_B200_H_ */
