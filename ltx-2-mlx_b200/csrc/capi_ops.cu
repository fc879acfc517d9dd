// extern "C" shims of the per-op entry points declared in include/ltx2_b200.h.
#include "common.cuh"
#include "kernels.h"
#include "../../include/ltx2_b200.h"

#include <math.h>

#include <map>
#include <mutex>
#include <tuple>
#include <vector>

using namespace ltx2;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

int ltx2_version(void) { return 100; }
const char* ltx2_last_error(void) { return last_error(); }

int ltx2_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int32_t M, int32_t N, int32_t K,
                   int32_t mode, const float* bias, void* out, int64_t ldo, const float* gate, int64_t gate_stride,
                   const int32_t* row_cls, float alpha, void* stream) {
  LTX2_REQUIRE(mode >= 0 && mode <= 3, "gemm: bad epilogue mode %d", mode);
  GemmEpilogue ep;
  ep.mode = mode;
  ep.bias = bias;
  ep.out = out;
  ep.ldo = ldo;
  ep.gate = gate;
  ep.gate_stride = gate_stride;
  ep.row_cls = row_cls;
  ep.alpha = alpha;
  return gemm_bf16(A, lda, W, ldw, M, N, K, ep, S(stream));
}

int ltx2_gemm_bf16_splitk(const void* A, int64_t lda, const void* W, int64_t ldw, int32_t M, int32_t N, int32_t K,
                          const float* bias, float* out, int64_t ldo, const float* gate, int64_t gate_stride,
                          const int32_t* row_cls, float alpha, int32_t max_splits, void* stream) {
  GemmEpilogue ep;
  ep.mode = GEMM_EPI_F32_RESIDUAL;
  ep.bias = bias;
  ep.out = out;
  ep.ldo = ldo;
  ep.gate = gate;
  ep.gate_stride = gate_stride;
  ep.row_cls = row_cls;
  ep.alpha = alpha;
  ep.max_splits = max_splits;
  return gemm_bf16(A, lda, W, ldw, M, N, K, ep, S(stream));
}

int ltx2_gemm_e4m3(const void* A8, int64_t lda, const void* W8, int64_t ldw, int32_t M, int32_t N, int32_t K,
                   int32_t mode, const float* bias, const float* row_scale, const float* col_scale, void* out,
                   int64_t ldo, void* stream) {
  LTX2_REQUIRE(mode >= 0 && mode <= 2, "gemm_e4m3: epilogue mode %d unsupported", mode);
  GemmEpilogue ep;
  ep.mode = mode;
  ep.bias = bias;
  ep.out = out;
  ep.ldo = ldo;
  ep.row_scale = row_scale;
  ep.col_scale = col_scale;
  return gemm_e4m3(A8, lda, W8, ldw, M, N, K, ep, S(stream));
}

int ltx2_norm_modulate_q8(const void* x, int32_t x_dtype, int64_t ldx, void* out8, int64_t ldo8, float* row_scale,
                          void* out_bf16, int64_t ldo16, int32_t M, int32_t D, int32_t norm_kind, float eps,
                          const float* mod, int64_t mod_stride, int64_t shift_off, int64_t scale_off,
                          const int32_t* row_cls, void* stream) {
  LTX2_REQUIRE(x_dtype == LTX2_F32 || x_dtype == LTX2_BF16, "norm_modulate_q8: x must be f32 or bf16");
  return norm_modulate_q8(x, x_dtype == LTX2_BF16, ldx, out8, ldo8, row_scale, out_bf16, ldo16, M, D, norm_kind, eps,
                          mod, mod_stride, shift_off, scale_off, row_cls, S(stream));
}

int ltx2_quantize_rows_e4m3(const void* w, int32_t dtype, int64_t rows, int64_t K, void* out8, float* row_scale,
                            void* stream) {
  return quantize_rows_e4m3(w, dtype, rows, K, out8, row_scale, S(stream));
}

int ltx2_attention(const void* q, const void* k, const void* vt, void* out, int32_t B, int32_t H, int32_t Tq,
                   int32_t Tk, int32_t Tkp, int32_t Dh, float scale, const float* gate_logits, float* lse_out,
                   void* stream) {
  return attention_bf16(q, k, vt, out, B, H, Tq, Tk, Tkp, Dh, scale, gate_logits, lse_out, S(stream));
}

int ltx2_attention_vrows(const void* q, const void* k, const void* v, int64_t v_stride_t, int64_t v_stride_h,
                         int64_t v_stride_b, void* out, int32_t B, int32_t H, int32_t Tq, int32_t Tk, int32_t Dh,
                         float scale, const float* gate_logits, float* lse_out, void* stream) {
  AttnV av;
  av.ptr = v;
  av.rows = 1;
  av.stride_t = v_stride_t;
  av.stride_h = v_stride_h;
  av.stride_b = v_stride_b;
  return attention_bf16_v(q, k, av, out, B, H, Tq, Tk, Dh, scale, gate_logits, lse_out, S(stream));
}

// diagnostics: ltx2_attention_vrows plus the clock64 timeline of CTA 0 (SM-pair kernel: trace[16 * key_blocks])
int ltx2_attention_vrows_trace(const void* q, const void* k, const void* v, int64_t v_stride_t, int64_t v_stride_h,
                               int64_t v_stride_b, void* out, int32_t B, int32_t H, int32_t Tq, int32_t Tk, int32_t Dh,
                               float scale, long long* trace, void* stream) {
  AttnV av;
  av.ptr = v;
  av.rows = 1;
  av.stride_t = v_stride_t;
  av.stride_h = v_stride_h;
  av.stride_b = v_stride_b;
  return attention_bf16_v(q, k, av, out, B, H, Tq, Tk, Dh, scale, nullptr, nullptr, S(stream), trace);
}

// diagnostics: same as ltx2_attention, and CTA (0,0) writes clock64 stamps of its pipeline events to trace[nkv*8]
int ltx2_attention_trace(const void* q, const void* k, const void* vt, void* out, int32_t B, int32_t H, int32_t Tq,
                         int32_t Tk, int32_t Tkp, int32_t Dh, float scale, long long* trace, void* stream) {
  return attention_bf16(q, k, vt, out, B, H, Tq, Tk, Tkp, Dh, scale, nullptr, nullptr, S(stream), trace);
}

int ltx2_norm_modulate(const void* x, int32_t x_dtype, int64_t ldx, void* out, int64_t ldo, int32_t M, int32_t D,
                       int32_t norm_kind, float eps, const float* mod, int64_t mod_stride, int64_t shift_off,
                       int64_t scale_off, const int32_t* row_cls, void* stream) {
  LTX2_REQUIRE(x_dtype == LTX2_F32 || x_dtype == LTX2_BF16, "norm_modulate: x must be f32 or bf16");
  return norm_modulate(x, x_dtype == LTX2_BF16, ldx, out, ldo, M, D, norm_kind, eps, mod, mod_stride, shift_off,
                       scale_off, row_cls, S(stream));
}

int ltx2_headnorm_rope(const void* in, int64_t ld, const float* weight, const float* cos, const float* sin, void* out,
                       int32_t B, int32_t T, int32_t H, int32_t Dh, float eps, void* stream) {
  return headnorm_rope(in, ld, weight, cos, sin, out, B, T, H, Dh, eps, S(stream));
}

int ltx2_v_transpose(const void* v, int64_t ld, void* vt, int32_t B, int32_t T, int32_t Tp, int32_t H, int32_t Dh,
                     void* stream) {
  return v_transpose(v, ld, vt, B, T, Tp, H, Dh, S(stream));
}

int ltx2_rope_tables(const float* positions, int32_t B, int32_t n_dims, int32_t T, int32_t dim,
                     const float* max_pos_host, float theta, float* cos, float* sin, void* stream) {
  LTX2_REQUIRE(n_dims >= 1 && n_dims <= 3 && dim % (2 * n_dims) >= 0, "rope_tables: bad n_dims");
  typedef std::tuple<int, float, int, int> Key;      // the cached grid is a DEVICE pointer: key it by device
  static std::map<Key, float*> cache;
  static std::mutex mu;
  const int n = dim / (2 * n_dims);
  float* dev = nullptr;
  {
    std::lock_guard<std::mutex> lock(mu);
    Key key(current_device(), theta, n_dims, dim);
    auto it = cache.find(key);
    if (it == cache.end()) {
      std::vector<float> g(n);
      for (int i = 0; i < n; ++i) {
        const float lin = n > 1 ? static_cast<float>(static_cast<double>(i) / (n - 1)) : 0.f;
        g[i] = static_cast<float>(pow(static_cast<double>(theta), static_cast<double>(lin)) * (M_PI / 2.0));
      }
      LTX2_CUDA_CHECK(cudaMalloc(&dev, n * 4 + 16));
      LTX2_CUDA_CHECK(cudaMemcpy(dev, g.data(), n * 4, cudaMemcpyHostToDevice));
      cache[key] = dev;
    } else {
      dev = it->second;
    }
  }
  return rope_tables_dev(positions, B, n_dims, n_dims, T, dim, max_pos_host, dev, n, cos, sin, S(stream));
}

int ltx2_timestep_sinusoid(const float* t, int32_t R, float multiplier, float* out256, void* stream) {
  return timestep_sinusoid(t, R, multiplier, out256, S(stream));
}

int ltx2_small_linear(const float* x, int32_t R, int32_t K, const void* W, const float* bias, float* y, int32_t N,
                      int32_t act_in, void* stream) {
  return small_linear(x, R, K, W, bias, y, N, act_in, S(stream));
}

int ltx2_x0_from_velocity(const float* latent, const float* velocity, const float* t_row, float* x0, int32_t M,
                          int32_t C, void* stream) {
  return x0_from_velocity(latent, velocity, t_row, x0, M, C, S(stream));
}

int ltx2_gemm_plan(int32_t M, int32_t N, int32_t K, int32_t mode, int32_t max_splits, int32_t* out6) {
  if (out6 == nullptr || M <= 0 || N <= 0 || K <= 0) {
    set_error("gemm_plan: bad argument");
    return LTX2_ERR_INVALID;
  }
  const GemmPlan p = plan_gemm(M, N, K, mode, max_splits, 0);
  out6[0] = p.kernel; out6[1] = p.bn; out6[2] = p.splits; out6[3] = p.tile_w; out6[4] = p.last_w; out6[5] = p.num_t;
  return LTX2_OK;
}

int ltx2_attention_plan(int32_t Tq, int32_t BH, int32_t* pairs_per_slice, int32_t* n_ctas) {
  if (Tq <= 0 || BH <= 0 || pairs_per_slice == nullptr) {
    set_error("attention_plan: bad argument");
    return LTX2_ERR_INVALID;
  }
  int n = 0;
  *pairs_per_slice = attention_pair_items(Tq, BH, &n);
  if (n_ctas != nullptr) *n_ctas = n;
  return LTX2_OK;
}

int ltx2_attention_sm_pair_plan(int32_t Tq, int32_t Tk, int32_t BH, int32_t* n_clusters, int32_t* split) {
  if (Tq <= 0 || Tk <= 0 || BH <= 0 || n_clusters == nullptr || split == nullptr) {
    set_error("attention_sm_pair_plan: bad argument");
    return LTX2_ERR_INVALID;
  }
  int c = 0, s = 0;
  attention_2cta_plan(Tq, Tk, BH, &c, &s);
  *n_clusters = c;
  *split = s;
  return LTX2_OK;
}

int ltx2_attention_sm_pair_segments(int32_t Tq, int32_t Tk, int32_t BH, int32_t cluster, int32_t* out7,
                                    int32_t max_segments) {
  if (Tq <= 0 || Tk <= 0 || BH <= 0 || out7 == nullptr || max_segments < 0) {
    set_error("attention_sm_pair_segments: bad argument");
    return LTX2_ERR_INVALID;
  }
  const int n = attention_2cta_segments(Tq, Tk, BH, cluster, out7, max_segments);
  if (n < 0) {
    set_error("attention_sm_pair_segments: cluster %d out of range", cluster);
    return LTX2_ERR_INVALID;
  }
  return n;
}

int ltx2_denoise_update(const float* sample, const float* cond_x0, const float* uncond_x0, float cfg_scale,
                        const float* denoise_mask, const float* clean_latent, float sigma, float sigma_next, float* out,
                        float* denoised_out, int32_t M, int32_t C, void* stream) {
  return denoise_update(sample, cond_x0, uncond_x0, cfg_scale, denoise_mask, clean_latent, sigma, sigma_next, out,
                        denoised_out, M, C, S(stream));
}

int ltx2_silu_mul(const void* a, const void* b, void* out, int64_t n, int32_t dtype, void* stream) {
  return silu_mul(a, b, out, n, dtype, S(stream));
}
int ltx2_gelu_mul(const void* a, const void* b, void* out, int64_t n, int32_t dtype, void* stream) {
  return gelu_mul(a, b, out, n, dtype, S(stream));
}
int ltx2_interleaved_rope(const void* x, const void* cos, const void* sin, void* out, int64_t n, int32_t dtype,
                          void* stream) {
  return interleaved_rope(x, cos, sin, out, n, dtype, S(stream));
}

int ltx2_cast(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype, int64_t n, void* stream) {
  if (dst_dtype == LTX2_BF16) return cast_to_bf16(src, src_dtype, dst, n, S(stream));
  if (dst_dtype == LTX2_F32) return cast_to_f32(src, src_dtype, reinterpret_cast<float*>(dst), n, S(stream));
  set_error("cast: destination dtype %d unsupported", dst_dtype);
  return LTX2_ERR_INVALID;
}

}  // extern "C"
