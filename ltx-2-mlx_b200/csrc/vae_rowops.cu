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
                                        __nv_bfloat16* __restrict__ out, int B, int C, int T, int H, int W, int causal,
                                        int t0, int Tn) {
  // T = frames of the latent; the output holds the window [t0, t0 + Tn) plus one pad slot on each side, which is the
  // true neighbour frame when the window is a shard of a longer clip (replication only at the clip's own ends)
  const int64_t total = static_cast<int64_t>(B) * (Tn + 2) * (H + 2) * (W + 2) * C;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = i % C;
    int64_t r = i / C;
    const int wp = r % (W + 2); r /= (W + 2);
    const int hp = r % (H + 2); r /= (H + 2);
    const int tp = r % (Tn + 2);
    const int b = r / (Tn + 2);
    const int t = src_t(tp + t0, T, causal), h = reflect1(hp - 1, H), w = reflect1(wp - 1, W);
    const int64_t src = (((static_cast<int64_t>(b) * C + c) * T + t) * H + h) * W + w;
    float v = ldf<Tin>(lat + src) * std_[c] + mean[c];
    if (noise != nullptr) v = noise[src] * ns + (1.0f - ns) * v;
    out[i] = __float2bfloat16(v);
  }
}

// One CTA per padded output row (b, tp, hp); P = min(C/8, 32) lanes cooperate on one position (16 B per lane and
// unit), so a warp handles 32/P positions per step.  A lane's channels are the same for every position of the row:
// its shift/scale values are loaded once.  All index math is 32-bit and done once per row.
constexpr int kPadMaxU = 4;        // units of 8 channels per lane: C <= 32 * 8 * 4 = 1024

template <int U>
__global__ void __launch_bounds__(256)
norm_act_pad_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int T, int H, int W, int C,
                    int P, int act, const float* __restrict__ mod, int mod_stride, int shift_off, int scale_off,
                    float eps, int causal, int skip_front, int skip_back) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x;                       // (b*(T+2) + tp)*(H+2) + hp
  const int hp = row % (H + 2);
  const int tp = (row / (H + 2)) % (T + 2);
  // temporal shards: a pad slot that holds a NEIGHBOUR rank's frame is filled by that rank (halo_push), not replicated
  if ((skip_front && tp == 0) || (skip_back && tp == T + 1)) return;
  const int b = row / ((H + 2) * (T + 2));
  const int t = src_t(tp, T, causal), h = reflect1(hp - 1, H);
  const __nv_bfloat16* src_row = x + (static_cast<int64_t>(b) * T + t) * H * static_cast<int64_t>(W) * C +
                                 static_cast<int64_t>(h) * W * C;
  __nv_bfloat16* dst_row = out + static_cast<int64_t>(row) * (W + 2) * C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / P, pl = lane % P;          // position slot inside the warp, lane inside the position group
  const int per_warp = 32 / P;
  float sh[U][8], sc[U][8];
  if (act) {
    const float* mrow = mod + static_cast<int64_t>(b) * mod_stride;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int c0 = (pl + u * P) * 8;
#pragma unroll
      for (int j = 0; j < 8; j += 4) {
        const float4 a = *reinterpret_cast<const float4*>(mrow + shift_off + c0 + j);
        const float4 d = *reinterpret_cast<const float4*>(mrow + scale_off + c0 + j);
        sh[u][j] = a.x; sh[u][j + 1] = a.y; sh[u][j + 2] = a.z; sh[u][j + 3] = a.w;
        sc[u][j] = 1.f + d.x; sc[u][j + 1] = 1.f + d.y; sc[u][j + 2] = 1.f + d.z; sc[u][j + 3] = 1.f + d.w;
      }
    }
  }
  const float inv_c = 1.0f / C;
  for (int wp0 = warp * per_warp; wp0 < W + 2; wp0 += 8 * per_warp) {
    const int wp = wp0 + sub;
    const bool ok = wp < W + 2;
    const int w = reflect1((ok ? wp : 0) - 1, W);
    const __nv_bfloat16* src = src_row + static_cast<int64_t>(w) * C;
    uint4 q[U];
#pragma unroll
    for (int u = 0; u < U; ++u) q[u] = ok ? *reinterpret_cast<const uint4*>(src + (pl + u * P) * 8) : make_uint4(0, 0, 0, 0);
    if (!act) {
      if (ok) {
#pragma unroll
        for (int u = 0; u < U; ++u) *reinterpret_cast<uint4*>(dst_row + static_cast<int64_t>(wp) * C + (pl + u * P) * 8) = q[u];
      }
      continue;
    }
    float v[U][8];
    float ss = 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&q[u]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(hh[j]);
        v[u][2 * j] = f.x;
        v[u][2 * j + 1] = f.y;
        ss = fmaf(f.x, f.x, fmaf(f.y, f.y, ss));
      }
    }
    for (int o = P >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss * inv_c + eps);
    if (ok) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float o8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float y = fmaf(v[u][j] * rstd, sc[u][j], sh[u][j]);
          o8[j] = __fdividef(y, 1.f + __expf(-y));
        }
        uint4 w4;
        w4.x = pack_bf16x2(o8[0], o8[1]);
        w4.y = pack_bf16x2(o8[2], o8[3]);
        w4.z = pack_bf16x2(o8[4], o8[5]);
        w4.w = pack_bf16x2(o8[6], o8[7]);
        *reinterpret_cast<uint4*>(dst_row + static_cast<int64_t>(wp) * C + (pl + u * P) * 8) = w4;
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

// decode_tiled accumulation (tiling.py:354-407): out[b,c,t0+t,h0+h,w0+w] += tile[b,c,t,h,w] * mt[t]*mh[h]*mw[w];
// weights[t0+t,h0+h,w0+w] += mt[t]*mh[h]*mw[w]   (weights has no batch/channel axis, tiling.py:357)
__global__ void tile_accumulate_kernel(float* __restrict__ out, float* __restrict__ wsum, const float* __restrict__ tile,
                                       int BC, int To, int Ho, int Wo, int dt, int dh, int dw, int t0, int h0, int w0,
                                       int tt, int th, int tw, const float* __restrict__ mt, const float* __restrict__ mh,
                                       const float* __restrict__ mw) {
  const int64_t n = static_cast<int64_t>(tt) * th * tw;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int w = i % tw;
    const int h = (i / tw) % th;
    const int t = i / (static_cast<int64_t>(tw) * th);
    const float m = mt[t] * mh[h] * mw[w];
    const int64_t o = (static_cast<int64_t>(t0 + t) * Ho + (h0 + h)) * Wo + (w0 + w);
    wsum[o] += m;
    for (int bc = 0; bc < BC; ++bc)
      out[static_cast<int64_t>(bc) * To * Ho * Wo + o] += tile[((static_cast<int64_t>(bc) * dt + t) * dh + h) * dw + w] * m;
  }
}

__global__ void tile_normalize_kernel(float* __restrict__ out, const float* __restrict__ wsum, int BC, int64_t plane) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < plane;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float inv = 1.0f / fmaxf(wsum[i], 1e-8f);
    for (int bc = 0; bc < BC; ++bc) out[static_cast<int64_t>(bc) * plane + i] *= inv;
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
                     cudaStream_t stream, int t0, int Tn) {
  LTX2_REQUIRE(H >= 2 && W >= 2, "reflect padding needs H, W >= 2 (got %d x %d)", H, W);
  if (Tn < 0) { t0 = 0; Tn = T; }
  LTX2_REQUIRE(t0 >= 0 && Tn >= 1 && t0 + Tn <= T && (!causal || (t0 == 0 && Tn == T)), "latent_to_padded: bad window");
  const int64_t total = static_cast<int64_t>(B) * (Tn + 2) * (H + 2) * (W + 2) * C;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  const int g = grid_for(total, 256);
  switch (dtype) {
    case LTX2_F32:
      latent_to_padded_kernel<<<g, 256, 0, stream>>>(reinterpret_cast<const float*>(latent), std_, mean, noise,
                                                     noise_scale, o, B, C, T, H, W, causal, t0, Tn);
      break;
    case LTX2_BF16:
      latent_to_padded_kernel<<<g, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(latent), std_, mean, noise,
                                                     noise_scale, o, B, C, T, H, W, causal, t0, Tn);
      break;
    case LTX2_F16:
      latent_to_padded_kernel<<<g, 256, 0, stream>>>(reinterpret_cast<const __half*>(latent), std_, mean, noise,
                                                     noise_scale, o, B, C, T, H, W, causal, t0, Tn);
      break;
    default: set_error("latent_to_padded: bad dtype %d", dtype); return LTX2_ERR_INVALID;
  }
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int norm_act_pad(const void* x, void* out, int B, int T, int H, int W, int C, int act, const float* mod,
                 int64_t mod_stride, int64_t shift_off, int64_t scale_off, float eps, int causal, cudaStream_t stream,
                 int skip_front, int skip_back) {
  LTX2_REQUIRE(C % 64 == 0 && C <= 32 * 8 * kPadMaxU && (C / 8 >= 32 ? (C / 8) % 32 == 0 : (32 % (C / 8)) == 0),
               "norm_act_pad: C=%d unsupported (64, 128, 256, 512, 768 or 1024)", C);
  LTX2_REQUIRE(H >= 2 && W >= 2, "reflect padding needs H, W >= 2 (got %d x %d)", H, W);
  LTX2_REQUIRE(!act || mod != nullptr, "norm_act_pad: modulation table required");
  const int units = C / 8;
  const int P = units >= 32 ? 32 : units;
  const int U = units / P;
  const unsigned rows = static_cast<unsigned>(B) * (T + 2) * (H + 2);
  const __nv_bfloat16* xi = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* xo = reinterpret_cast<__nv_bfloat16*>(out);
  const int ms = static_cast<int>(mod_stride), so = static_cast<int>(shift_off), sc = static_cast<int>(scale_off);
  switch (U) {
    case 1: LTX2_CUDA_CHECK(launch_pdl(norm_act_pad_kernel<1>, dim3(rows), dim3(256), 0, stream, xi, xo, T, H, W, C, P, act, mod, ms, so, sc, eps, causal, skip_front, skip_back)); break;
    case 2: LTX2_CUDA_CHECK(launch_pdl(norm_act_pad_kernel<2>, dim3(rows), dim3(256), 0, stream, xi, xo, T, H, W, C, P, act, mod, ms, so, sc, eps, causal, skip_front, skip_back)); break;
    case 3: LTX2_CUDA_CHECK(launch_pdl(norm_act_pad_kernel<3>, dim3(rows), dim3(256), 0, stream, xi, xo, T, H, W, C, P, act, mod, ms, so, sc, eps, causal, skip_front, skip_back)); break;
    case 4: LTX2_CUDA_CHECK(launch_pdl(norm_act_pad_kernel<4>, dim3(rows), dim3(256), 0, stream, xi, xo, T, H, W, C, P, act, mod, ms, so, sc, eps, causal, skip_front, skip_back)); break;
    default: set_error("norm_act_pad: C=%d unsupported", C); return LTX2_ERR_INVALID;
  }
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

// Temporal shards of the decoder (vae_engine.cu): after a rank has produced its padded conv input [B, n+2, Hp*Wp*C], its
// first real frame (slot 1) goes into the PREVIOUS rank's back pad slot (n_prev + 1) and its last real frame (slot n)
// into the NEXT rank's front pad slot (0) -- stores to peer memory over NVLink, 16 bytes per thread.
__global__ void halo_push_kernel(const uint4* __restrict__ mine, uint4* __restrict__ prev, uint4* __restrict__ next, int B,
                                 int n, int n_prev, int n_next, int64_t frame16) {
  const int64_t per = static_cast<int64_t>(B) * frame16;
  const int64_t total = per * ((prev != nullptr ? 1 : 0) + (next != nullptr ? 1 : 0));
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const bool to_prev = prev != nullptr && i < per;
    const int64_t j = to_prev ? i : i - (prev != nullptr ? per : 0);
    const int b = j / frame16;
    const int64_t e = j % frame16;
    if (to_prev)
      prev[(static_cast<int64_t>(b) * (n_prev + 2) + n_prev + 1) * frame16 + e] =
          mine[(static_cast<int64_t>(b) * (n + 2) + 1) * frame16 + e];
    else
      next[(static_cast<int64_t>(b) * (n_next + 2)) * frame16 + e] = mine[(static_cast<int64_t>(b) * (n + 2) + n) * frame16 + e];
  }
}

int halo_push(const void* mine, void* prev, void* next, int B, int n, int n_prev, int n_next, int64_t frame_bytes,
              cudaStream_t stream) {
  if (prev == nullptr && next == nullptr) return LTX2_OK;
  LTX2_REQUIRE(frame_bytes % 16 == 0 && n >= 1, "halo_push: bad frame size");
  const int64_t total = static_cast<int64_t>(B) * (frame_bytes / 16) * 2;
  halo_push_kernel<<<grid_for(total, 256), 256, 0, stream>>>(reinterpret_cast<const uint4*>(mine),
                                                            reinterpret_cast<uint4*>(prev), reinterpret_cast<uint4*>(next),
                                                            B, n, n_prev, n_next, frame_bytes / 16);
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

int tile_accumulate(float* out, float* wsum, const float* tile, int BC, int To, int Ho, int Wo, int dt, int dh, int dw,
                    int t0, int h0, int w0, int tt, int th, int tw, const float* mt, const float* mh, const float* mw,
                    cudaStream_t stream) {
  LTX2_REQUIRE(t0 >= 0 && h0 >= 0 && w0 >= 0 && t0 + tt <= To && h0 + th <= Ho && w0 + tw <= Wo && tt <= dt &&
                   th <= dh && tw <= dw,
               "tile_accumulate: tile [%d,%d,%d]+[%d,%d,%d] does not fit output [%d,%d,%d]", t0, h0, w0, tt, th, tw, To,
               Ho, Wo);
  const int64_t n = static_cast<int64_t>(tt) * th * tw;
  if (n == 0) return LTX2_OK;
  tile_accumulate_kernel<<<grid_for(n, 256), 256, 0, stream>>>(out, wsum, tile, BC, To, Ho, Wo, dt, dh, dw, t0, h0, w0,
                                                                tt, th, tw, mt, mh, mw);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int tile_normalize(float* out, const float* wsum, int BC, int64_t plane, cudaStream_t stream) {
  if (plane == 0) return LTX2_OK;
  tile_normalize_kernel<<<grid_for(plane, 256), 256, 0, stream>>>(out, wsum, BC, plane);
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
