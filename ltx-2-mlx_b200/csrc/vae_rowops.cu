// HBM-bound kernels of the video-VAE decode path (everything that is not the conv MMA).
//
// Reference ops replaced (model/video_vae/simple_decoder.py):
//   latent_to_padded   de-normalise (:492-493) + noise blend (:496-498) + Conv3dSimple padding (:105-134)
//   norm_act_pad       _pixel_norm (:339-342) + scale/shift (:218-231, :535-541) + SiLU + the padding of the
//                      following conv (:105-134): reflect H/W, first/last-frame replication in T
//   pack_conv_weight   layout change of PyTorch [Cout,Cin,3,3,3] weights (the reference re-slices per call, :146-150)
//   blend_chunk / video_to_uint8   decode_latent's cross-fade and uint8 conversion (:749-798)
#include "common.cuh"
#include "kernels.h"

#include <cuda_fp16.h>

namespace ltx2 {
namespace {

__device__ __forceinline__ int reflect1(int i, int n) {      // index into [0,n) for a pad-1 reflect
  if (i < 0) return -i;                                       // -1 -> 1
  if (i >= n) return 2 * n - 2 - i;                           // n -> n-2
  return i;
}
__device__ __forceinline__ int src_t(int tp, int T, int causal) {
  int t = causal ? tp - 2 : tp - 1;
  return t < 0 ? 0 : (t >= T ? T - 1 : t);
}

template <typename Tin> __device__ __forceinline__ float ldf(const Tin* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <> __device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(*p); }

// one thread per output element; channels innermost so writes coalesce (the latent itself is tiny)
template <typename Tin>
__global__ void latent_to_padded_kernel(const Tin* __restrict__ lat, const float* __restrict__ std_,
                                        const float* __restrict__ mean, const float* __restrict__ noise, float ns,
                                        __nv_bfloat16* __restrict__ out, int B, int C, int T, int H, int W, int causal) {
  const int64_t total = static_cast<int64_t>(B) * (T + 2) * (H + 2) * (W + 2) * C;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = i % C;
    int64_t r = i / C;
    const int wp = r % (W + 2); r /= (W + 2);
    const int hp = r % (H + 2); r /= (H + 2);
    const int tp = r % (T + 2);
    const int b = r / (T + 2);
    const int t = src_t(tp, T, causal), h = reflect1(hp - 1, H), w = reflect1(wp - 1, W);
    const int64_t src = (((static_cast<int64_t>(b) * C + c) * T + t) * H + h) * W + w;
    float v = ldf<Tin>(lat + src) * std_[c] + mean[c];
    if (noise != nullptr) v = noise[src] * ns + (1.0f - ns) * v;
    out[i] = __float2bfloat16(v);
  }
}

// one warp per padded output position; the source row (C <= 1024 channels) lives in registers
constexpr int kMaxCPerLane = 32;   // C <= 1024

__global__ void __launch_bounds__(256)
norm_act_pad_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int T, int H, int W,
                    int C, int act, const float* __restrict__ mod, int64_t mod_stride, int64_t shift_off,
                    int64_t scale_off, float eps, int causal) {
  const int64_t npos = static_cast<int64_t>(B) * (T + 2) * (H + 2) * (W + 2);
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int units = C / 8;                     // 16-byte units per row
  for (int64_t pi = warp0; pi < npos; pi += nwarps) {
    int64_t r = pi;
    const int wp = r % (W + 2); r /= (W + 2);
    const int hp = r % (H + 2); r /= (H + 2);
    const int tp = r % (T + 2);
    const int b = r / (T + 2);
    const int t = src_t(tp, T, causal), h = reflect1(hp - 1, H), w = reflect1(wp - 1, W);
    const __nv_bfloat16* src = x + (((static_cast<int64_t>(b) * T + t) * H + h) * W + w) * C;
    __nv_bfloat16* dst = out + pi * C;
    if (!act) {
      for (int u = lane; u < units; u += 32)
        *reinterpret_cast<uint4*>(dst + u * 8) = *reinterpret_cast<const uint4*>(src + u * 8);
      continue;
    }
    float v[kMaxCPerLane];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxCPerLane / 8; ++k) {
      const int u = lane + k * 32;
      if (u < units) {
        const uint4 q = *reinterpret_cast<const uint4*>(src + u * 8);
        const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(hh[j]);
          v[k * 8 + 2 * j] = f.x;
          v[k * 8 + 2 * j + 1] = f.y;
          ss += f.x * f.x + f.y * f.y;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss / C + eps);
    const float* mrow = mod + static_cast<int64_t>(b) * mod_stride;
#pragma unroll
    for (int k = 0; k < kMaxCPerLane / 8; ++k) {
      const int u = lane + k * 32;
      if (u < units) {
        float o8[8];
#pragma unroll
        for (int j = 0; j < 8; j += 4) {
          const float4 sh = *reinterpret_cast<const float4*>(mrow + shift_off + u * 8 + j);
          const float4 sc = *reinterpret_cast<const float4*>(mrow + scale_off + u * 8 + j);
          const float shv[4] = {sh.x, sh.y, sh.z, sh.w}, scv[4] = {sc.x, sc.y, sc.z, sc.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float y = v[k * 8 + j + q] * rstd * (1.f + scv[q]) + shv[q];
            o8[j + q] = y / (1.f + __expf(-y));
          }
        }
        uint4 q;
        q.x = pack_bf16x2(o8[0], o8[1]);
        q.y = pack_bf16x2(o8[2], o8[3]);
        q.z = pack_bf16x2(o8[4], o8[5]);
        q.w = pack_bf16x2(o8[6], o8[7]);
        *reinterpret_cast<uint4*>(dst + u * 8) = q;
      }
    }
  }
}

template <typename Tin>
__global__ void pack_conv_weight_kernel(const Tin* __restrict__ w, __nv_bfloat16* __restrict__ packed, int Cout,
                                        int Cout_pad, int Cin, int sp) {
  const int64_t total = static_cast<int64_t>(Cout_pad) * 27 * Cin;
  const int Cf = Cout / sp;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ci = i % Cin;
    const int tap = (i / Cin) % 27;
    const int row = i / (static_cast<int64_t>(Cin) * 27);
    float v = 0.f;
    if (row < Cout) {
      const int src_row = sp > 1 ? (row % Cf) * sp + row / Cf : row;
      v = ldf<Tin>(w + (static_cast<int64_t>(src_row) * Cin + ci) * 27 + tap);
    }
    packed[i] = __float2bfloat16(v);
  }
}

template <typename Tin>
__global__ void pack_conv_bias_kernel(const Tin* __restrict__ b, float* __restrict__ packed, int Cout, int Cout_pad,
                                      int sp) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= Cout_pad) return;
  const int Cf = Cout / sp;
  float v = 0.f;
  if (row < Cout) v = ldf<Tin>(b + (sp > 1 ? (row % Cf) * sp + row / Cf : row));
  packed[row] = v;
}

__global__ void blend_chunk_kernel(float* __restrict__ dst, const float* __restrict__ src, int T_dst, int T_src,
                                   int HW, int t0, int overlap, int64_t total) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int hw = i % HW;
    const int ts = (i / HW) % T_src;
    const int bc = i / (static_cast<int64_t>(HW) * T_src);
    const int td = t0 + ts;
    if (td >= T_dst) continue;
    float* d = dst + (static_cast<int64_t>(bc) * T_dst + td) * HW + hw;
    const float s = src[i];
    if (ts < overlap) {
      const float ramp = overlap > 1 ? static_cast<float>(ts) / static_cast<float>(overlap - 1) : 0.f;
      *d = *d * (1.0f - ramp) + s * ramp;
    } else {
      *d = s;
    }
  }
}

__global__ void video_to_uint8_kernel(const float* __restrict__ v, uint8_t* __restrict__ out, int T, int H, int W) {
  const int64_t total = static_cast<int64_t>(T) * H * W;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float x = (v[c * total + i] + 1.0f) / 2.0f;
      x = fminf(fmaxf(x, 0.f), 1.f) * 255.0f;
      out[i * 3 + c] = static_cast<uint8_t>(x);          // truncation, like astype(uint8)
    }
  }
}

inline int grid_for(int64_t n, int threads) {
  int64_t g = (n + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 32;
  return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

int latent_to_padded(const void* latent, int dtype, const float* std_, const float* mean, const float* noise,
                     float noise_scale, void* out, int B, int C, int T, int H, int W, int causal,
                     cudaStream_t stream) {
  LTX2_REQUIRE(H >= 2 && W >= 2, "reflect padding needs H, W >= 2 (got %d x %d)", H, W);
  const int64_t total = static_cast<int64_t>(B) * (T + 2) * (H + 2) * (W + 2) * C;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  const int g = grid_for(total, 256);
  switch (dtype) {
    case LTX2_F32:
      latent_to_padded_kernel<<<g, 256, 0, stream>>>(reinterpret_cast<const float*>(latent), std_, mean, noise,
                                                     noise_scale, o, B, C, T, H, W, causal);
      break;
    case LTX2_BF16:
      latent_to_padded_kernel<<<g, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(latent), std_, mean, noise,
                                                     noise_scale, o, B, C, T, H, W, causal);
      break;
    case LTX2_F16:
      latent_to_padded_kernel<<<g, 256, 0, stream>>>(reinterpret_cast<const __half*>(latent), std_, mean, noise,
                                                     noise_scale, o, B, C, T, H, W, causal);
      break;
    default: set_error("latent_to_padded: bad dtype %d", dtype); return LTX2_ERR_INVALID;
  }
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int norm_act_pad(const void* x, void* out, int B, int T, int H, int W, int C, int act, const float* mod,
                 int64_t mod_stride, int64_t shift_off, int64_t scale_off, float eps, int causal, cudaStream_t stream) {
  LTX2_REQUIRE(C % 8 == 0 && C <= 32 * kMaxCPerLane, "norm_act_pad: C=%d unsupported", C);
  LTX2_REQUIRE(H >= 2 && W >= 2, "reflect padding needs H, W >= 2 (got %d x %d)", H, W);
  LTX2_REQUIRE(!act || mod != nullptr, "norm_act_pad: modulation table required");
  const int64_t npos = static_cast<int64_t>(B) * (T + 2) * (H + 2) * (W + 2);
  const int g = grid_for(npos * 32, 256);
  norm_act_pad_kernel<<<g, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                             reinterpret_cast<__nv_bfloat16*>(out), B, T, H, W, C, act, mod, mod_stride,
                                             shift_off, scale_off, eps, causal);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int pack_conv_weight(const void* w, int dtype, void* packed, int Cout, int Cout_pad, int Cin, int sp,
                     cudaStream_t stream) {
  const int64_t total = static_cast<int64_t>(Cout_pad) * 27 * Cin;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(packed);
  const int g = grid_for(total, 256);
  switch (dtype) {
    case LTX2_F32: pack_conv_weight_kernel<<<g, 256, 0, stream>>>(reinterpret_cast<const float*>(w), o, Cout, Cout_pad, Cin, sp); break;
    case LTX2_BF16: pack_conv_weight_kernel<<<g, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(w), o, Cout, Cout_pad, Cin, sp); break;
    case LTX2_F16: pack_conv_weight_kernel<<<g, 256, 0, stream>>>(reinterpret_cast<const __half*>(w), o, Cout, Cout_pad, Cin, sp); break;
    default: set_error("pack_conv_weight: bad dtype %d", dtype); return LTX2_ERR_INVALID;
  }
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int pack_conv_bias(const void* b, int dtype, float* packed, int Cout, int Cout_pad, int sp, cudaStream_t stream) {
  const int g = (Cout_pad + 127) / 128;
  switch (dtype) {
    case LTX2_F32: pack_conv_bias_kernel<<<g, 128, 0, stream>>>(reinterpret_cast<const float*>(b), packed, Cout, Cout_pad, sp); break;
    case LTX2_BF16: pack_conv_bias_kernel<<<g, 128, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(b), packed, Cout, Cout_pad, sp); break;
    case LTX2_F16: pack_conv_bias_kernel<<<g, 128, 0, stream>>>(reinterpret_cast<const __half*>(b), packed, Cout, Cout_pad, sp); break;
    default: set_error("pack_conv_bias: bad dtype %d", dtype); return LTX2_ERR_INVALID;
  }
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int blend_chunk(float* dst, const float* src, int BC, int T_dst, int T_src, int HW, int t0, int overlap,
                cudaStream_t stream) {
  const int64_t total = static_cast<int64_t>(BC) * T_src * HW;
  if (total == 0) return LTX2_OK;
  blend_chunk_kernel<<<grid_for(total, 256), 256, 0, stream>>>(dst, src, T_dst, T_src, HW, t0, overlap, total);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int video_to_uint8(const float* video, uint8_t* out, int T, int H, int W, cudaStream_t stream) {
  const int64_t total = static_cast<int64_t>(T) * H * W;
  if (total == 0) return LTX2_OK;
  video_to_uint8_kernel<<<grid_for(total, 256), 256, 0, stream>>>(video, out, T, H, W);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

}  // namespace ltx2
