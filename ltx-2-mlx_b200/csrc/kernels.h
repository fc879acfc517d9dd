// Internal (C++) launch interface of the sm_100a kernels.  The public C ABI in
// include/ltx2_b200.h is a thin shim over these.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ltx2 {

// ---------------------------------------------------------------------------------
// GEMM  C[M,N] = A[M,K] W[N,K]^T with fused epilogue (gemm_sm100.cu)
// ---------------------------------------------------------------------------------
enum GemmEpilogueMode {
  GEMM_EPI_BF16 = 0,          // out_bf16 = acc + bias
  GEMM_EPI_BF16_GELU = 1,     // out_bf16 = gelu_tanh(acc + bias)
  GEMM_EPI_F32 = 2,           // out_f32  = acc + bias
  GEMM_EPI_F32_RESIDUAL = 3,  // out_f32 += alpha * gate[row_cls[row], col] * (acc + bias)
  GEMM_EPI_E4M3_GELU = 4,     // out_e4m3 = gelu_tanh(acc + bias) / (out_l2[row] * out_coef[0] + out_coef[1])
};

struct GemmEpilogue {
  int mode = GEMM_EPI_BF16;
  const float* bias = nullptr;     // [N] fp32 or null
  void* out = nullptr;             // bf16 or fp32, row pitch ldo elements
  int64_t ldo = 0;
  const float* gate = nullptr;     // [n_cls, gate_stride] fp32 or null (= 1)
  int64_t gate_stride = 0;
  const int* row_cls = nullptr;    // [M] modulation class of each row, or null (= 0)
  float alpha = 1.0f;
  void* out_peers[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int n_out_peers = 0;             // bf16 modes: > 0 replicates the store into every out_peers[i] (peer-memory broadcast)
  int max_splits = 1;              // > 1 allows split-K for the residual mode (accumulation order then varies run to run)
  // FP8 path (gemm_e4m3): acc * row_scale[row] * col_scale[col] before the bias -- the dynamic per-token scale of the
  // E4M3 activation row and the (per-tensor or per-output-channel) scale of the E4M3 weight
  const float* row_scale = nullptr;
  const float* col_scale = nullptr;
  // row_coef != null: the row factor is row_scale[row] * row_coef[0] + row_coef[1] (two DEVICE floats) -- the scale of
  // an E4M3 activation that was written by a GEMM epilogue (mode 4) against a bound instead of a measured absmax:
  // |gelu(a.w + b)| <= |a|_2 max_j|w_j|_2 + max|b|  (Cauchy-Schwarz), with |a|_2 per row from the norm kernel
  const float* row_coef = nullptr;
  const float* out_l2 = nullptr;     // mode 4: [M] L2 norm of the GEMM's own input rows
  const float* out_coef = nullptr;   // mode 4: two device floats
};

int gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, const GemmEpilogue& ep,
              cudaStream_t stream);

// Same contract with E4M3 operands (A [M,K], W [N,K] bytes; tcgen05.mma kind::f8f6f4, fp32 accumulate, K = 32 per
// instruction: twice the bf16 tensor rate); ep.row_scale / ep.col_scale are required.  K % 16 == 0.
int gemm_e4m3(const void* A8, int64_t lda, const void* W8, int64_t ldw, int M, int N, int K, const GemmEpilogue& ep,
              cudaStream_t stream);

// Which kernel and tiling gemm_bf16 takes for a problem (pure host logic: needs no GPU, uses 148 SMs when no device
// is visible).  standard: 128 token rows x bn columns, K split `splits` ways;  transposed: 128 weight rows x tile_w
// tokens (last tile last_w), num_t token tiles, K split `splits` ways;  pair: SM pairs, 256 weight rows x tile_w tokens;
// wide: SM pairs, 256 weight rows x tile_w <= 512 tokens in two accumulators, K split `splits` ways (shard shapes).
enum GemmKernel { GEMM_KERNEL_STANDARD = 0, GEMM_KERNEL_TRANSPOSED = 1, GEMM_KERNEL_PAIR = 2, GEMM_KERNEL_WIDE = 3 };
struct GemmPlan {
  int kernel = GEMM_KERNEL_STANDARD;
  int bn = 256, splits = 1;
  int tile_w = 0, last_w = 0, num_t = 0;
};
GemmPlan plan_gemm(int M, int N, int K, int mode, int max_splits, int n_out_peers);

// ---------------------------------------------------------------------------------
// attention (attention_sm100.cu):  O = softmax(Q K^T * scale) V, no mask
//   q  [B,H,Tq,Dh] bf16, k [B,H,Tk,Dh] bf16, vt [B,H,Dh,Tkp] bf16 (V transposed, pitch Tkp >= Tk, Tkp % 8 == 0)
//   out [B,Tq,H*Dh] bf16 (token-major).  gate_logits [B*Tq, H] fp32 or null: out *= 2*sigmoid(logit)
//   partial (optional, for ring attention): when lse_out != null the kernel also writes the
//   log-sum-exp (natural log, scaled scores) per row to lse_out [B,H,Tq] fp32.
// ---------------------------------------------------------------------------------
// V operand of the attention: either transposed [B,H,Dh,Tkp] (rows = 0) or in row form, element (b,h,t,d) at
// ptr + b*stride_b + h*stride_h + t*stride_t + d (rows = 1) -- e.g. straight out of the fused QKV GEMM output
// (stride_t = 3*inner, stride_h = Dh, stride_b = T*3*inner), consumed as an MN-major tcgen05 B operand.
struct AttnV {
  const void* ptr = nullptr;
  int rows = 0;
  int64_t Tkp = 0;
  int64_t stride_t = 0, stride_h = 0, stride_b = 0;
};
// Context-parallel output re-shard fused into the attention epilogue: query row t of (local) head h is stored to
// peer[t / rows_per_rank] at [(b*rows_per_rank + t % rows_per_rank) * pitch + (head0 + h)*Dh] (rows_per_rank = 0: off).
constexpr int kMaxCpRanks = 8;
struct AttnOutScatter {
  __nv_bfloat16* peer[kMaxCpRanks] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int rows_per_rank = 0;
  int pitch = 0;
  int head0 = 0;
};
int attention_bf16_v(const void* q, const void* k, const AttnV& v, void* out, int B, int H, int Tq, int Tk, int Dh,
                     float scale, const float* gate_logits, float* lse_out, cudaStream_t stream,
                     long long* trace = nullptr, const AttnOutScatter* scatter = nullptr);

// head_dim 128 only: two softmax streams per CTA (attention_pair_sm100.cu); attention_bf16_v dispatches to it.
// trace (diagnostics): CTA 0 writes clock64 stamps to trace[16 * key_blocks].
// Work items of the two-stream attention kernel for Tq queries and BH (batch x head) slices: pair items per slice
// (two 128-query tiles sharing K/V) -- the remaining tiles run as split-KV items -- chosen for the smallest makespan
// over the SMs (pure host logic).
int attention_pair_items(int Tq, int BH, int* n_ctas);

int attention_pair_bf16(const void* q, const void* k, const AttnV& v, void* out, int B, int H, int Tq, int Tk,
                        float scale, const float* gate_logits, float* lse_out, cudaStream_t stream, long long* trace,
                        const AttnOutScatter& sc);

// head_dim 128, V in row form, more than one query tile: SM-pair kernel (attention_2cta_sm100.cu, cta_group::2; two
// adjacent query tiles of a head per cluster; persistent, stream-K over the key blocks).  attention_bf16_v dispatches
// to it for very long sequences (>= 8192 queries and keys); LTX2_ATTN_2CTA=1 forces it, =0 removes it.
// trace (diagnostics): cluster 0 / CTA 0 writes clock64 stamps to trace[16 * key_blocks].
bool attention_2cta_applies(const AttnV& v, int Tq, int Tk, int Dh);
// the pair kernel's work decomposition (pure host logic): clusters in the grid and whether (item, key block) ranges are
// split across clusters (stream-K) or whole items go round-robin
void attention_2cta_plan(int Tq, int Tk, int BH, int* n_clusters, int* split);
// the segments cluster `cluster` walks, 7 ints each: item, first key block, end key block, part, parts, scratch slot of
// the segment, scratch slot of part 0 (-1 for whole items); returns the segment count, -1 for a bad cluster index
int attention_2cta_segments(int Tq, int Tk, int BH, int cluster, int* out7, int max_segments);
int attention_2cta_bf16(const void* q, const void* k, const AttnV& v, void* out, int B, int H, int Tq, int Tk,
                        float scale, const float* gate_logits, float* lse_out, cudaStream_t stream, long long* trace,
                        const AttnOutScatter& sc);

// Self-attention head preparation in ONE pass over a token row of the fused QKV projection (pitch ld):
//   q,k: RMSNorm over the full row (learned weight) + split RoPE;  v: copy (skipped when dst.v[0] == null);
// each head h is written to dst.{q,k,v}[h / heads_per_rank] at [(b*heads_per_rank + h % heads_per_rank) * n_total +
// t_offset + t] * Dh.  With one rank this is the plain (B,T,H*D)->(B,H,T,D) split; with context parallelism the
// destinations are peer pointers, i.e. the token->head re-shard (all-to-all) is fused into this kernel's stores.
struct HeadScatter {
  __nv_bfloat16* q[kMaxCpRanks];
  __nv_bfloat16* k[kMaxCpRanks];
  __nv_bfloat16* v[kMaxCpRanks];
  int heads_per_rank;
  int n_total;
  int t_offset;
};
int qkv_head_scatter(const void* qkv, int64_t ld, const float* wq, const float* wk, const float* cos, const float* sin,
                     const HeadScatter& dst, int B, int T, int H, int Dh, float eps, cudaStream_t stream);

// copy `bytes` from src to dst_peers[0..n_peers) (peer pointers), coalesced
int gate_scatter(const float* logits, float* const* dst_peers, int B, int n_local, int H, int heads_per_rank,
                 int n_total, int t_offset, cudaStream_t stream);
int peer_broadcast(const void* src, void* const* dst_peers, int n_peers, int64_t bytes, cudaStream_t stream);

// cross-GPU barrier on flag words in peer memory: signal epoch to every rank's flags[rank], wait for all of mine
int cp_barrier(uint32_t* const* peer_flags_dev, uint32_t* my_flags, int rank, int world, uint32_t epoch,
               cudaStream_t stream);
// barrier among the ranks [first, first + count) only; slot0 = first flag word of the barrier domain (8 words each)
int cp_barrier_group(uint32_t* const* peer_flags_dev, uint32_t* my_flags, int rank, int first, int count, int slot0,
                     uint32_t epoch, cudaStream_t stream);
int attention_bf16(const void* q, const void* k, const void* vt, void* out, int B, int H, int Tq, int Tk, int Tkp,
                   int Dh, float scale, const float* gate_logits, float* lse_out, cudaStream_t stream,
                   long long* trace = nullptr);

// ---------------------------------------------------------------------------------
// row kernels (rowops.cu)
// ---------------------------------------------------------------------------------
enum NormKind { NORM_NONE = 0, NORM_RMS = 1, NORM_LAYER = 2 };

// out_bf16[row,:] = norm(x[row,:]) * (1 + scale) + shift, with
//   shift = mod[cls*mod_stride + shift_off + col], scale = mod[cls*mod_stride + scale_off + col], cls = row_cls[row]
//   (mod == null -> plain norm).  x is fp32 (x_is_bf16 = 0) or bf16.
int norm_modulate(const void* x, int x_is_bf16, int64_t ldx, void* out_bf16, int64_t ldo, int M, int D, int norm_kind,
                  float eps, const float* mod, int64_t mod_stride, int64_t shift_off, int64_t scale_off,
                  const int* row_cls, cudaStream_t stream);

// RMSNorm over the full row with a learned weight, optional split-RoPE, then scatter to heads:
//   in [B*T, inner] bf16 (pitch ld) -> out [B,H,T,Dh] bf16.   cos/sin [B,T,inner/2] fp32 or null.
int headnorm_rope(const void* in, int64_t ld, const float* weight, const float* cos, const float* sin, void* out,
                  int B, int T, int H, int Dh, float eps, cudaStream_t stream);

// V [B*T, inner] bf16 (pitch ld) -> Vt [B,H,Dh,Tp] bf16 (zero-padded to pitch Tp)
int v_transpose(const void* v, int64_t ld, void* vt, int B, int T, int Tp, int H, int Dh, cudaStream_t stream);

// y[r, n] = act_out( sum_k act_in(x[r,k]) * W[n,k] + bias[n] ), x fp32 [R<=8 rows, K], W bf16 [N,K], y fp32
//   act_in: 0 none, 1 SiLU.  Used for the timestep-embedding MLPs (timestep_embedding.py:166-202).
int small_linear(const float* x, int R, int K, const void* W, const float* bias, float* y, int N, int act_in,
                 cudaStream_t stream);

// out[m,h] = x[m,:].W[h,:] + b[h], x/W bf16, out fp32 (small-H fallback for to_gate_logits, attention.py:244)
int rowdot_bf16(const void* x, int64_t ldx, const void* W, const float* bias, float* out, int M, int H, int K,
                cudaStream_t stream);

// sinusoidal timestep features [cos | sin] of 1000*sigma, 256 wide (timestep_embedding.py:10-60 with
// flip_sin_to_cos=True, shift 0;  simple_decoder.py:12-39 is the same formula)
int timestep_sinusoid(const float* t, int R, float multiplier, float* out256, cudaStream_t stream);

// out[l, c, k, :] = tables[l*table_layer_stride + k*D + :] + emb[c*emb_cls_stride + k*emb_row_stride + :]
// written at out + l*out_layer_stride + c*out_cls_stride + k*D   (adaLN table row + timestep-embedding row;
// get_ada_values transformer.py:369-392, output head model.py:750-754)
int build_modulation_ex(const float* tables, int64_t table_layer_stride, const float* emb, int64_t emb_cls_stride,
                        int64_t emb_row_stride, float* out, int64_t out_layer_stride, int64_t out_cls_stride, int L,
                        int C, int R, int D, cudaStream_t stream);

// split-RoPE tables from [start,end) position bounds (rope.py:365-418): positions [B,n_dims,T,2] fp32 ->
// cos,sin [B,T,dim/2] fp32 in token-major order (head h owns columns [h*dim/2/H, (h+1)*dim/2/H) ).
// freq_grid_dev: device array of n_freq = dim/(2*n_dims) floats, theta^linspace(0,1,n_freq) * pi/2.
// positions carry pos_dims axes per batch element; only the first n_dims are used (the cross-modal table
// uses the temporal axis of the 3-axis video positions, model.py:330-331).
int rope_tables_dev(const float* positions, int B, int pos_dims, int n_dims, int T, int dim, const float* max_pos_host,
                    const float* freq_grid_dev, int n_freq, float* cos, float* sin, cudaStream_t stream);

// x0 = latent - t[row] * velocity  (model.py:912-918), fp32
int x0_from_velocity(const float* latent, const float* velocity, const float* t_row, float* x0, int M, int C,
                     cudaStream_t stream);

// One fused pass for the elementwise tail of a denoising step (the reference's host loop, pipelines/distilled.py:243-251,
// one_stage.py:284-320): CFG guide (uncond != null) -> masked blend with the clean latent (mask [M] per row, clean
// [M,C]; both null = off) -> Euler step.  All fp32 [M,C]; denoised_out (optional) receives the guided+blended x0.
int denoise_update(const float* sample, const float* cond, const float* uncond, float cfg_scale, const float* mask,
                   const float* clean, float sigma, float sigma_next, float* out, float* denoised_out, int M, int C,
                   cudaStream_t stream);

// ---- FP8 operand preparation (fp8ops.cu) ----
// norm_modulate + per-row E4M3 quantisation: out8 [M,D] bytes, row_scale [M] = absmax/448; out_bf16 (optional) also
// receives the unquantised bf16 row (consumers that stay bf16, e.g. the gate-logit projection)
int norm_modulate_q8(const void* x, int x_is_bf16, int64_t ldx, void* out8, int64_t ldo8, float* row_scale,
                     void* out_bf16, int64_t ldo16, int M, int D, int norm_kind, float eps, const float* mod,
                     int64_t mod_stride, int64_t shift_off, int64_t scale_off, const int* row_cls, cudaStream_t stream,
                     float* row_l2 = nullptr /* optional [M]: L2 norm of the output row before quantisation */);
// coef[0] = 1.07 * max_row |w_row|_2 / 448, coef[1] = max|bias| / 448 for an E4M3 weight [rows,K] with row scales:
// the per-row scale factors of the E4M3 output of the GEMM that uses this weight (GEMM_EPI_E4M3_GELU)
int e4m3_bound_coef(const void* w8, const float* row_scale, const float* bias, int64_t rows, int64_t K, float* coef,
                    cudaStream_t stream);
// weight [rows,K] (dtype code) -> E4M3 bytes with one scale per row
int quantize_rows_e4m3(const void* w, int dtype, int64_t rows, int64_t K, void* out8, float* row_scale,
                       cudaStream_t stream);
// dst = e4m3(src) * scale[row * scale_stride]  (scale_stride 0 = per-tensor scale); dst fp32 or bf16
int dequant_e4m3(const void* src8, const float* scale, int scale_stride, int64_t rows, int64_t K, void* dst,
                 int dst_dtype, cudaStream_t stream);
int fill_f32(float* dst, float v, int64_t n, cudaStream_t stream);

int cast_to_bf16(const void* src, int src_dtype, void* dst, int64_t n, cudaStream_t stream);
int cast_to_f32(const void* src, int src_dtype, float* dst, int64_t n, cudaStream_t stream);

// The reference's three Metal kernels (kernels/fused_ops.py), fp32/bf16/fp16 by dtype code
int silu_mul(const void* a, const void* b, void* out, int64_t n, int dtype, cudaStream_t stream);
int gelu_mul(const void* a, const void* b, void* out, int64_t n, int dtype, cudaStream_t stream);
int interleaved_rope(const void* x, const void* cos, const void* sin, void* out, int64_t n, int dtype,
                     cudaStream_t stream);

// ---------------------------------------------------------------------------------
// video-VAE decoder kernels (conv3d_sm100.cu, vae_rowops.cu)
// ---------------------------------------------------------------------------------
enum ConvEpilogueMode {
  CONV_EPI_PLAIN = 0,       // out[B,T,H,W,Cout] bf16 = acc + bias
  CONV_EPI_RESIDUAL = 1,    // ... + residual[B,T,H,W,Cout] (may alias out)
  CONV_EPI_D2S = 2,         // depth-to-space scatter (+ tiled d2s residual gathered from the conv input)
  CONV_EPI_UNPATCHIFY = 3,  // conv_out: fp32 [B, Cout/16, T, 4H, 4W]
};

struct ConvParams {
  int B = 0, T = 0, H = 0, W = 0;      // output (= unpadded input) grid
  int Cin = 0, Cout = 0, Cout_pad = 0; // Cout_pad: rows of the packed weight / bias (multiple of 32)
  int mode = CONV_EPI_PLAIN;
  const float* bias = nullptr;          // [Cout_pad]
  __nv_bfloat16* out = nullptr;
  float* out_f32 = nullptr;
  const __nv_bfloat16* residual = nullptr;   // RESIDUAL: [.,Cout];  D2S: the UNPADDED conv input [.,Cin]
  int ft = 1, fh = 1, fw = 1;           // D2S strides
  int c_d2s = 0;                        // D2S residual: Cin / (ft*fh*fw); 0 = no residual
  // Fused producer (PLAIN / RESIDUAL, C_out == 128 or 256 so that one accumulator row holds every channel of its
  // position): the epilogue ALSO writes the padded input of the next conv, [B,T+2,H+2,W+2,Cout] bf16 with reflect H/W
  // and replicate T borders, optionally through pixel-norm * (1+scale) + shift and SiLU -- what norm_act_pad would do
  // in a separate pass over HBM (simple_decoder.py:105-134, 229-231, 339-342).  `out` may then be null (the raw
  // activation of a ResBlock's first conv has no other reader).
  __nv_bfloat16* pad_out = nullptr;
  int pad_act = 0;                      // 1: silu(pixel_norm(v) * (1 + scale[b]) + shift[b]); 0: plain copy
  const float* pad_mod = nullptr;       // fp32 [B, pad_mod_stride]
  int64_t pad_mod_stride = 0, pad_shift_off = 0, pad_scale_off = 0;
  float pad_eps = 1e-6f;
  int pad_causal = 0;
  int pad_skip_front = 0, pad_skip_back = 0;   // temporal shards: that T pad slot belongs to the neighbour rank
  int out_t_total = 0, out_t0 = 0;      // UNPATCHIFY into frames [out_t0, out_t0 + T) of a clip of out_t_total frames
  int d2s_keep_first = 0;               // temporal shards: this rank's first input frame is not the clip's frame 0, so
                                        // the depth-to-space keeps both of its output frames (no first-frame drop)
};

// x_padded: bf16 [B, T+2, H+2, W+2, Cin]; w_packed: bf16 [Cout_pad, 27*Cin] (tap-major, channel-minor)
int conv3d_bf16(const void* x_padded, const void* w_packed, const ConvParams& p, cudaStream_t stream);

// PyTorch conv weight [Cout, Cin, 3,3,3] (any float dtype) -> packed bf16 [Cout_pad, 27*Cin]; row permutation for
// depth-to-space convs: packed row (sub*Cf + c) <- source row (c*sp + sub); rows >= Cout are zero.  Same for bias.
int pack_conv_weight(const void* w, int dtype, void* packed, int Cout, int Cout_pad, int Cin, int sp,
                     cudaStream_t stream);
int pack_conv_bias(const void* b, int dtype, float* packed, int Cout, int Cout_pad, int sp, cudaStream_t stream);

// latent NCDHW (fp32/bf16/fp16) -> padded channels-last bf16 [B,T+2,H+2,W+2,C] with de-normalisation
// (x*std+mean, simple_decoder.py:492-493) and optional noise blend (:496-498): noise*s + (1-s)*x
int latent_to_padded(const void* latent, int dtype, const float* std_, const float* mean, const float* noise,
                     float noise_scale, void* out, int B, int C, int T, int H, int W, int causal,
                     cudaStream_t stream, int t0 = 0, int Tn = -1 /* window [t0, t0+Tn) of the T frames; -1 = all */);

// x [B,T,H,W,C] bf16 -> padded [B,T+2,H+2,W+2,C] bf16 applying (when act != 0)
//   silu(pixel_norm(x) * (1 + scale[b]) + shift[b])   (simple_decoder.py:229-231, 339-342, 528-542)
// mod: fp32 [B, mod_stride] with rows shift at shift_off and scale at scale_off; reflect pad H/W, replicate pad T.
int norm_act_pad(const void* x, void* out, int B, int T, int H, int W, int C, int act, const float* mod,
                 int64_t mod_stride, int64_t shift_off, int64_t scale_off, float eps, int causal, cudaStream_t stream,
                 int skip_front = 0, int skip_back = 0 /* temporal shards: leave that pad slot to the neighbour rank */);
// temporal shards: my first / last real frame -> the previous rank's back pad slot / the next rank's front pad slot
int halo_push(const void* mine, void* prev, void* next, int B, int n, int n_prev, int n_next, int64_t frame_bytes,
              cudaStream_t stream);

// decode_latent post-processing (simple_decoder.py:749-798)
//   blend: dst[:, :, t0+i] = dst*(1-r_i) + src[:, :, i]*r_i for i < overlap (r = linspace(0,1,overlap)), then copy tail
int blend_chunk(float* dst, const float* src, int BC, int T_dst, int T_src, int HW, int t0, int overlap,
                cudaStream_t stream);
// decode_tiled (tiling.py:354-412): weighted accumulation of one decoded tile and the final normalisation
int tile_accumulate(float* out, float* wsum, const float* tile, int BC, int To, int Ho, int Wo, int dt, int dh, int dw,
                    int t0, int h0, int w0, int tt, int th, int tw, const float* mt, const float* mh, const float* mw,
                    cudaStream_t stream);
int tile_normalize(float* out, const float* wsum, int BC, int64_t plane, cudaStream_t stream);
//   video [1,3,T,H,W] fp32 in [-1,1] -> uint8 [T,H,W,3]
int video_to_uint8(const float* video, uint8_t* out, int T, int H, int W, cudaStream_t stream);

// dtype codes shared with the C ABI
enum { LTX2_F32 = 0, LTX2_BF16 = 1, LTX2_F16 = 2, LTX2_F8E4M3 = 3 };

}  // namespace ltx2
