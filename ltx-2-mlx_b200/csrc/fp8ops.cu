// FP8 (E4M3) operand preparation for the tcgen05 kind::f8f6f4 GEMM path.
//
// Reference semantics being kept: loader/fp8_loader.py:14-32 defines an FP8 checkpoint tensor as
// weight_fp8 * weight_scale (per-tensor scale).  The reference widens to bf16/fp16 at load; this engine keeps the E4M3
// bytes of the norm-fed linears (self-attention QKV, text-attention Q, FFN up) and multiplies by the scale in the GEMM
// epilogue instead.  The activation operand of those GEMMs is produced by the adaLN/RMSNorm kernel, which sees a whole
// token row: it quantises the row to E4M3 with a per-ROW dynamic scale (absmax / 448), also applied in the epilogue:
//     out[m, n] = (sum_k a8[m,k] w8[n,k]) * a_scale[m] * w_scale[n] + bias[n]
//
//   norm_modulate_q8     _compiled_adaln_forward / rms_norm (transformer.py:16-31) + row quantisation
//   quantize_rows_e4m3   weight [N,K] (fp32/bf16/fp16) -> E4M3 with one scale per output row (bf16 checkpoints)
//   dequant_e4m3         E4M3 * scale -> fp32 / bf16 (weight read-back, FP8 tensors of layers that stay bf16)
#include "common.cuh"
#include "kernels.h"

#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include <algorithm>

namespace ltx2 {
namespace {

constexpr int kQThreads = 256;
constexpr int kQMaxUnits = 4;     // D <= 256 * 8 * 4 = 8192
constexpr float kE4M3Max = 448.0f;

__device__ __forceinline__ float warp_sum_q(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_q(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int THREADS>
__device__ __forceinline__ float2 block_sum2_q(float a, float b) {
  __shared__ float sa[THREADS / 32], sb[THREADS / 32];
  a = warp_sum_q(a);
  b = warp_sum_q(b);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sa[w] = a; sb[w] = b; }
  __syncthreads();
  float ra = 0.f, rb = 0.f;
#pragma unroll
  for (int i = 0; i < THREADS / 32; ++i) { ra += sa[i]; rb += sb[i]; }
  __syncthreads();
  return make_float2(ra, rb);
}
template <int THREADS>
__device__ __forceinline__ float block_max_q(float a) {
  __shared__ float sm[THREADS / 32];
  a = warp_max_q(a);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sm[w] = a;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < THREADS / 32; ++i) r = fmaxf(r, sm[i]);
  __syncthreads();
  return r;
}

__device__ __forceinline__ void ld8_bf16(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void ld8_f32(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
// 8 floats (already divided by the scale) -> 8 E4M3 bytes, round-to-nearest-even, saturating
__device__ __forceinline__ uint2 pack8_e4m3(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    w[i] = __nv_cvt_float2_to_fp8x2(make_float2(f[2 * i], f[2 * i + 1]), __NV_SATFINITE, __NV_E4M3);   // .x in the low byte
  return make_uint2(w[0] | (w[1] << 16), w[2] | (w[3] << 16));
}
__device__ __forceinline__ float e4m3_to_float(uint8_t b) {
  const __half_raw h = __nv_cvt_fp8_to_halfraw(b, __NV_E4M3);
  return __half2float(*reinterpret_cast<const __half*>(&h));
}

// one CTA per row; the row stays in registers between the statistics, the modulation and the quantisation
template <bool X_BF16>
__global__ void __launch_bounds__(kQThreads)
norm_modulate_q8_kernel(const void* __restrict__ x_, int64_t ldx, uint8_t* __restrict__ out8, int64_t ldo8,
                        float* __restrict__ row_scale, __nv_bfloat16* __restrict__ out16, int64_t ldo16, int D,
                        int norm_kind, float eps, const float* __restrict__ mod, int64_t mod_stride, int64_t shift_off,
                        int64_t scale_off, const int* __restrict__ row_cls, float* __restrict__ row_l2) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x;
  const int units = D / 8;
  float v[kQMaxUnits][8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int u = 0; u < kQMaxUnits; ++u) {
    const int idx = threadIdx.x + u * kQThreads;
    if (idx < units) {
      if (X_BF16) ld8_bf16(reinterpret_cast<const __nv_bfloat16*>(x_) + row * ldx + idx * 8, v[u]);
      else ld8_f32(reinterpret_cast<const float*>(x_) + row * ldx + idx * 8, v[u]);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s1 += v[u][i]; s2 += v[u][i] * v[u][i]; }
    }
  }
  float mean = 0.f, rstd = 1.f;
  if (norm_kind != NORM_NONE) {
    const float2 s = block_sum2_q<kQThreads>(s1, s2);
    if (norm_kind == NORM_LAYER) {
      mean = s.x / D;
      float d2 = 0.f;
#pragma unroll
      for (int u = 0; u < kQMaxUnits; ++u) {
        const int idx = threadIdx.x + u * kQThreads;
        if (idx < units) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { const float d = v[u][i] - mean; d2 += d * d; }
        }
      }
      const float2 t = block_sum2_q<kQThreads>(d2, 0.f);
      rstd = rsqrtf(t.x / D + eps);
    } else {
      rstd = rsqrtf(s.y / D + eps);
    }
  }
  const float* mrow = nullptr;
  if (mod != nullptr) mrow = mod + static_cast<int64_t>(row_cls ? row_cls[row] : 0) * mod_stride;
  float amax = 0.f, l2 = 0.f;
#pragma unroll
  for (int u = 0; u < kQMaxUnits; ++u) {
    const int idx = threadIdx.x + u * kQThreads;
    if (idx < units) {
      if (mrow != nullptr) {
        float sh[8], sc[8];
        ld8_f32(mrow + shift_off + idx * 8, sh);
        ld8_f32(mrow + scale_off + idx * 8, sc);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[u][i] = (v[u][i] - mean) * rstd * (1.f + sc[i]) + sh[i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[u][i] = (v[u][i] - mean) * rstd;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) { amax = fmaxf(amax, fabsf(v[u][i])); l2 = fmaf(v[u][i], v[u][i], l2); }
      if (out16 != nullptr) {
        uint4 q;
        q.x = pack_bf16x2(v[u][0], v[u][1]);
        q.y = pack_bf16x2(v[u][2], v[u][3]);
        q.z = pack_bf16x2(v[u][4], v[u][5]);
        q.w = pack_bf16x2(v[u][6], v[u][7]);
        *reinterpret_cast<uint4*>(out16 + row * ldo16 + idx * 8) = q;
      }
    }
  }
  amax = block_max_q<kQThreads>(amax);
  if (row_l2 != nullptr) {                      // warp-uniform
    const float2 t = block_sum2_q<kQThreads>(l2, 0.f);
    if (threadIdx.x == 0) row_l2[row] = sqrtf(t.x);
  }
  const float scale = amax > 0.f ? amax * (1.0f / kE4M3Max) : 1.0f;
  const float inv = 1.0f / scale;
  if (threadIdx.x == 0) row_scale[row] = scale;
#pragma unroll
  for (int u = 0; u < kQMaxUnits; ++u) {
    const int idx = threadIdx.x + u * kQThreads;
    if (idx < units) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = v[u][i] * inv;
      *reinterpret_cast<uint2*>(out8 + row * ldo8 + idx * 8) = pack8_e4m3(o);
    }
  }
}

// weight rows: one CTA per output row, two passes over the row (absmax, then convert)
template <typename Tin>
__global__ void __launch_bounds__(kQThreads)
quantize_rows_kernel(const Tin* __restrict__ w, int64_t K, uint8_t* __restrict__ out8, float* __restrict__ row_scale) {
  const int64_t row = blockIdx.x;
  const Tin* src = w + row * K;
  float amax = 0.f;
  for (int64_t k = threadIdx.x; k < K; k += kQThreads) amax = fmaxf(amax, fabsf(static_cast<float>(src[k])));
  amax = block_max_q<kQThreads>(amax);
  const float scale = amax > 0.f ? amax * (1.0f / kE4M3Max) : 1.0f;
  const float inv = 1.0f / scale;
  if (threadIdx.x == 0) row_scale[row] = scale;
  for (int64_t k = threadIdx.x; k < K; k += kQThreads)
    out8[row * K + k] = static_cast<uint8_t>(__nv_cvt_float_to_fp8(static_cast<float>(src[k]) * inv, __NV_SATFINITE, __NV_E4M3));
}

// dst[r, k] = e4m3(src[r, k]) * scale[r * scale_stride]   (scale_stride 0: one scale for the tensor)
template <typename Tout>
__global__ void dequant_e4m3_kernel(const uint8_t* __restrict__ src, const float* __restrict__ scale, int scale_stride,
                                    int64_t rows, int64_t K, Tout* __restrict__ dst) {
  const int64_t n = rows * K;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float s = scale[(i / K) * scale_stride];
    dst[i] = static_cast<Tout>(e4m3_to_float(src[i]) * s);
  }
}

// bound coefficients of an E4M3 weight: max over rows of |w_row|_2 (dequantised) and max |bias|
__global__ void __launch_bounds__(kQThreads)
e4m3_row_norm_max_kernel(const uint8_t* __restrict__ w8, const float* __restrict__ row_scale,
                         const float* __restrict__ bias, int64_t K, unsigned int* __restrict__ acc /* [2] float bits */) {
  const int64_t row = blockIdx.x;
  float ss = 0.f;
  for (int64_t k = threadIdx.x; k < K; k += kQThreads) {
    const float f = e4m3_to_float(w8[row * K + k]);
    ss = fmaf(f, f, ss);
  }
  const float2 t = block_sum2_q<kQThreads>(ss, 0.f);
  if (threadIdx.x == 0) {
    atomicMax(acc, __float_as_uint(sqrtf(t.x) * row_scale[row]));          // non-negative floats order like their bits
    atomicMax(acc + 1, __float_as_uint(bias != nullptr ? fabsf(bias[row]) : 0.f));
  }
}
__global__ void e4m3_bound_finish_kernel(const unsigned int* acc, float* coef) {
  coef[0] = 1.07f * __uint_as_float(acc[0]) * (1.0f / kE4M3Max);
  coef[1] = fmaxf(__uint_as_float(acc[1]) * (1.0f / kE4M3Max), 1e-30f);
}

__global__ void fill_f32_kernel(float* dst, float v, int64_t n) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) dst[i] = v;
}

}  // namespace

int norm_modulate_q8(const void* x, int x_is_bf16, int64_t ldx, void* out8, int64_t ldo8, float* row_scale,
                     void* out_bf16, int64_t ldo16, int M, int D, int norm_kind, float eps, const float* mod,
                     int64_t mod_stride, int64_t shift_off, int64_t scale_off, const int* row_cls, cudaStream_t stream,
                     float* row_l2) {
  if (M == 0) return LTX2_OK;
  LTX2_REQUIRE(D % 8 == 0 && D <= kQThreads * 8 * kQMaxUnits, "norm_modulate_q8: D=%d unsupported", D);
  LTX2_REQUIRE(ldx % 8 == 0 && ldo8 % 16 == 0 && ldo16 % 8 == 0 && shift_off % 4 == 0 && scale_off % 4 == 0 &&
                   mod_stride % 4 == 0,
               "norm_modulate_q8: pitches/offsets must keep 16-byte alignment");
  LTX2_REQUIRE(out8 != nullptr && row_scale != nullptr, "norm_modulate_q8: null output");
  if (x_is_bf16)
    LTX2_CUDA_CHECK(launch_pdl(norm_modulate_q8_kernel<true>, dim3(M), dim3(kQThreads), 0, stream, x, ldx,
                               reinterpret_cast<uint8_t*>(out8), ldo8, row_scale,
                               reinterpret_cast<__nv_bfloat16*>(out_bf16), ldo16, D, norm_kind, eps, mod, mod_stride,
                               shift_off, scale_off, row_cls, row_l2));
  else
    LTX2_CUDA_CHECK(launch_pdl(norm_modulate_q8_kernel<false>, dim3(M), dim3(kQThreads), 0, stream, x, ldx,
                               reinterpret_cast<uint8_t*>(out8), ldo8, row_scale,
                               reinterpret_cast<__nv_bfloat16*>(out_bf16), ldo16, D, norm_kind, eps, mod, mod_stride,
                               shift_off, scale_off, row_cls, row_l2));
  count_launch();
  return LTX2_OK;
}

int quantize_rows_e4m3(const void* w, int dtype, int64_t rows, int64_t K, void* out8, float* row_scale,
                       cudaStream_t stream) {
  if (rows == 0) return LTX2_OK;
  uint8_t* o = reinterpret_cast<uint8_t*>(out8);
  switch (dtype) {
    case LTX2_F32:
      quantize_rows_kernel<float><<<static_cast<unsigned>(rows), kQThreads, 0, stream>>>(
          reinterpret_cast<const float*>(w), K, o, row_scale);
      break;
    case LTX2_BF16:
      quantize_rows_kernel<__nv_bfloat16><<<static_cast<unsigned>(rows), kQThreads, 0, stream>>>(
          reinterpret_cast<const __nv_bfloat16*>(w), K, o, row_scale);
      break;
    case LTX2_F16:
      quantize_rows_kernel<__half><<<static_cast<unsigned>(rows), kQThreads, 0, stream>>>(
          reinterpret_cast<const __half*>(w), K, o, row_scale);
      break;
    default: set_error("quantize_rows_e4m3: bad dtype %d", dtype); return LTX2_ERR_INVALID;
  }
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int dequant_e4m3(const void* src8, const float* scale, int scale_stride, int64_t rows, int64_t K, void* dst,
                 int dst_dtype, cudaStream_t stream) {
  const int64_t n = rows * K;
  if (n == 0) return LTX2_OK;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>((n + 255) / 256, 148 * 16));
  const uint8_t* s = reinterpret_cast<const uint8_t*>(src8);
  if (dst_dtype == LTX2_F32)
    dequant_e4m3_kernel<float><<<grid, 256, 0, stream>>>(s, scale, scale_stride, rows, K, reinterpret_cast<float*>(dst));
  else if (dst_dtype == LTX2_BF16)
    dequant_e4m3_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(s, scale, scale_stride, rows, K,
                                                                reinterpret_cast<__nv_bfloat16*>(dst));
  else {
    set_error("dequant_e4m3: destination dtype %d unsupported", dst_dtype);
    return LTX2_ERR_INVALID;
  }
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int e4m3_bound_coef(const void* w8, const float* row_scale, const float* bias, int64_t rows, int64_t K, float* coef,
                    cudaStream_t stream) {
  // coef doubles as the two-word accumulator of the reduction (bit patterns of non-negative floats)
  LTX2_CUDA_CHECK(cudaMemsetAsync(coef, 0, 8, stream));
  e4m3_row_norm_max_kernel<<<static_cast<unsigned>(rows), kQThreads, 0, stream>>>(
      reinterpret_cast<const uint8_t*>(w8), row_scale, bias, K, reinterpret_cast<unsigned int*>(coef));
  e4m3_bound_finish_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<const unsigned int*>(coef), coef);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch(2);
  return LTX2_OK;
}

int fill_f32(float* dst, float v, int64_t n, cudaStream_t stream) {
  if (n == 0) return LTX2_OK;
  fill_f32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(dst, v, n);
  LTX2_CUDA_CHECK(cudaGetLastError());
  return LTX2_OK;
}

}  // namespace ltx2
