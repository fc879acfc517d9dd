// Kernels around the tcgen05 conv for the video-VAE ENCODER and the 2x latent SPATIAL UPSCALER.
// Both networks are 3x3x3 conv stacks like the decoder, so their FLOPs run on conv3d_sm100.cu; this file holds what
// differs from the decoder: zero / causal padding, pixel-norm + SiLU without modulation, GroupNorm, space-to-depth with
// the group-mean residual, pixel shuffle, patchify, latent normalisation.  Activations are channels-last bf16.
//
// Reference ops replaced:
//   encoder   model/video_vae/simple_encoder.py: Conv3dSimple padding (:46-117), EncoderResBlock3d (:120-154),
//             SpaceToDepthDownsample3d (:172-257), forward + latent normalisation (:306-405); ops.py:9-68 patchify
//   upscaler  model/upscaler/spatial.py: conv3d zero padding (:21-88), group_norm_5d (:91-128), ResBlock3d (:131-181),
//             PixelShuffle2d / SpatialRationalResampler (:184-323), SpatialUpscaler forward (:376-412)
#include "common.cuh"
#include "kernels.h"
#include "../../include/ltx2_b200.h"

#include <cuda_fp16.h>

#include <algorithm>

using namespace ltx2;
typedef __nv_bfloat16 bf16;

namespace {

constexpr int kAuxThreads = 256;

__device__ __forceinline__ float silu_f(float y) { return __fdividef(y, 1.f + __expf(-y)); }

// One warp per OUTPUT position of the padded tensor [B, Tp, H+2, W+2, C]; lanes stride over 8-channel units.
//   hw_mode 0 reflect / 1 zero;  t_mode 0 replicate (1 front, 1 back) / 1 causal (first frame twice in front, nothing
//   behind) / 2 zero (1 front, 1 back);  dup_first: the logical input is [x[0], x[0], x[1], ...] (T + 1 frames).
//   act 0 none, 1 pixel-norm + SiLU, 2 GroupNorm affine, 3 GroupNorm affine + SiLU; `residual` is added after the norm,
//   before the SiLU.  `plain` (optional) receives the un-padded result.
__global__ void __launch_bounds__(kAuxThreads)
pad_act_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, bf16* __restrict__ plain, int B, int T, int H, int W,
               int C, int hw_mode, int t_mode, int dup_first, int act, const float* __restrict__ gn_stats,
               const float* __restrict__ gn_w, const float* __restrict__ gn_b, int groups, float eps,
               const bf16* __restrict__ residual) {
  const int TL = T + (dup_first ? 1 : 0);                   // logical frames
  const int Tp = TL + 2, Hp = H + 2, Wp = W + 2;
  const int64_t n_pos = static_cast<int64_t>(B) * Tp * Hp * Wp;
  const int lane = threadIdx.x & 31;
  const int units = C / 8;
  for (int64_t pos = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; pos < n_pos;
       pos += (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5) {
    int64_t r = pos;
    const int wp = r % Wp; r /= Wp;
    const int hp = r % Hp; r /= Hp;
    const int tp = r % Tp;
    const int b = r / Tp;
    int tl = t_mode == 1 ? tp - 2 : tp - 1;                  // logical source frame
    int h = hp - 1, w = wp - 1;
    bool zero = false;
    if (t_mode == 2) { if (tl < 0 || tl >= TL) zero = true; }
    else tl = tl < 0 ? 0 : (tl >= TL ? TL - 1 : tl);
    if (hw_mode == 1) { if (h < 0 || h >= H || w < 0 || w >= W) zero = true; }
    else {
      h = h < 0 ? -h : (h >= H ? 2 * H - 2 - h : h);
      w = w < 0 ? -w : (w >= W ? 2 * W - 2 - w : w);
    }
    bf16* dst = out + pos * C;
    if (zero) {
      for (int u = lane; u < units; u += 32) *reinterpret_cast<uint4*>(dst + u * 8) = make_uint4(0, 0, 0, 0);
      continue;
    }
    const int t = dup_first ? (tl > 0 ? tl - 1 : 0) : tl;    // physical frame
    const int64_t spos = ((static_cast<int64_t>(b) * T + t) * H + h) * W + w;
    const bf16* src = x + spos * C;
    float rstd = 1.f;
    if (act == 1) {
      float ss = 0.f;
      for (int u = lane; u < units; u += 32) {
        const uint4 q = *reinterpret_cast<const uint4*>(src + u * 8);
        const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(hh[j]);
          ss = fmaf(f.x, f.x, fmaf(f.y, f.y, ss));
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      rstd = rsqrtf(ss / C + eps);
    }
    // is this padded position the interior copy of the source position? (then it also feeds `plain`)
    const bool interior = plain != nullptr && hp >= 1 && hp <= H && wp >= 1 && wp <= W &&
                          (t_mode == 1 ? tp >= 2 : (tp >= 1 && tp <= TL)) && !dup_first;
    const int cpg = C / (groups > 0 ? groups : 1);
    for (int u = lane; u < units; u += 32) {
      const uint4 q = *reinterpret_cast<const uint4*>(src + u * 8);
      const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&q);
      float v[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(hh[j]);
        v[2 * j] = f.x;
        v[2 * j + 1] = f.y;
      }
      if (act == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = silu_f(v[j] * rstd);
      } else if (act >= 2) {
        const int c0 = u * 8;
        const float* st = gn_stats + (static_cast<int64_t>(b) * groups + c0 / cpg) * 2;   // 8 | cpg: one group per unit
        const float mean = st[0], rs = st[1];
        float rv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (residual != nullptr) {
          const uint4 rq = *reinterpret_cast<const uint4*>(residual + spos * C + c0);
          const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(&rq);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(rh[j]);
            rv[2 * j] = f.x;
            rv[2 * j + 1] = f.y;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float y = (v[j] - mean) * rs * gn_w[c0 + j] + gn_b[c0 + j] + rv[j];
          v[j] = act == 3 ? silu_f(y) : y;
        }
      }
      uint4 o;
      o.x = pack_bf16x2(v[0], v[1]);
      o.y = pack_bf16x2(v[2], v[3]);
      o.z = pack_bf16x2(v[4], v[5]);
      o.w = pack_bf16x2(v[6], v[7]);
      *reinterpret_cast<uint4*>(dst + u * 8) = o;
      if (interior) *reinterpret_cast<uint4*>(plain + spos * C + u * 8) = o;
    }
  }
}

// GroupNorm statistics over (C/groups, T, H, W) per (batch, group): stats[b, g] = (mean, rstd)   (spatial.py:91-128)
__global__ void __launch_bounds__(kAuxThreads)
group_stats_kernel(const bf16* __restrict__ x, int64_t thw, int C, int groups, float eps, float* __restrict__ stats) {
  const int b = blockIdx.x / groups, g = blockIdx.x % groups;
  const int cpg = C / groups;                               // multiple of 8
  const int upg = cpg / 8;
  const bf16* base = x + static_cast<int64_t>(b) * thw * C + g * cpg;
  double s1 = 0.0, s2 = 0.0;
  const int64_t n_units = thw * upg;
  for (int64_t i = threadIdx.x; i < n_units; i += kAuxThreads) {
    const int64_t p = i / upg;
    const int u = i % upg;
    const uint4 q = *reinterpret_cast<const uint4*>(base + p * C + u * 8);
    const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&q);
    float a = 0.f, c = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(hh[j]);
      a += f.x + f.y;
      c = fmaf(f.x, f.x, fmaf(f.y, f.y, c));
    }
    s1 += a;
    s2 += c;
  }
  __shared__ double sh1[kAuxThreads / 32], sh2[kAuxThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) { sh1[threadIdx.x >> 5] = s1; sh2[threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int i = 0; i < kAuxThreads / 32; ++i) { a += sh1[i]; c += sh2[i]; }
    const double n = static_cast<double>(thw) * cpg;
    const double mean = a / n;
    const double var = fmax(c / n - mean * mean, 0.0);
    stats[blockIdx.x * 2] = static_cast<float>(mean);
    stats[blockIdx.x * 2 + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
}

// video fp32 [B,3,F,H,W] -> channels-last bf16 [B,F,H/4,W/4,Cp] with channel (c*4 + r_w)*4 + r_h (ops.py:44-68);
// channels >= 48 (padding to the conv's 64-channel K granule) are zero
__global__ void patchify_video_kernel(const float* __restrict__ v, bf16* __restrict__ out, int B, int F, int H, int W,
                                      int Cp) {
  const int Hq = H / 4, Wq = W / 4;
  const int64_t n = static_cast<int64_t>(B) * F * Hq * Wq * Cp;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = i % Cp;
    int64_t r = i / Cp;
    const int wq = r % Wq; r /= Wq;
    const int hq = r % Hq; r /= Hq;
    const int f = r % F;
    const int b = r / F;
    float val = 0.f;
    if (ch < 48) {
      const int c = ch / 16, rw = (ch / 4) % 4, rh = ch % 4;
      val = v[(((static_cast<int64_t>(b) * 3 + c) * F + f) * H + hq * 4 + rh) * W + wq * 4 + rw];
    }
    out[i] = __float2bfloat16(val);
  }
}

// SpaceToDepthDownsample3d tail (simple_encoder.py:207-257): out[b,t',h',w',co] = y[b, t'st+a, h'sh+bh, w'sw+bw, co/sp]
//   + mean over the G = Cx*sp/Cout consecutive space-to-depth channels of x that fold onto co, with (a,bh,bw) = co % sp
// y [B,TL,H,W,Cy] (Cy = Cout/sp), x [B,T,H,W,Cx]; dup_first: logical frame l of x is physical max(l-1, 0), TL = T + 1
__global__ void s2d_residual_kernel(const bf16* __restrict__ y, const bf16* __restrict__ x, bf16* __restrict__ out, int B,
                                    int T, int TL, int H, int W, int Cx, int Cout, int st, int sh, int sw, int dup_first) {
  const int sp = st * sh * sw;
  const int Cy = Cout / sp;
  const int To = TL / st, Ho = H / sh, Wo = W / sw;
  const int G = Cx * sp / Cout;
  const int64_t n = static_cast<int64_t>(B) * To * Ho * Wo * Cout;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int co = i % Cout;
    int64_t r = i / Cout;
    const int wo = r % Wo; r /= Wo;
    const int ho = r % Ho; r /= Ho;
    const int to = r % To;
    const int b = r / To;
    const int sub = co % sp;
    const int a = sub / (sh * sw), bh = (sub / sw) % sh, bw = sub % sw;
    const int64_t ypos = ((static_cast<int64_t>(b) * TL + to * st + a) * H + ho * sh + bh) * W + wo * sw + bw;
    float acc = __bfloat162float(y[ypos * Cy + co / sp]);
    float m = 0.f;
    for (int g = 0; g < G; ++g) {
      const int j = co * G + g;                              // channel of space_to_depth(x): ci * sp + sub2
      const int ci = j / sp, s2 = j % sp;
      const int a2 = s2 / (sh * sw), bh2 = (s2 / sw) % sh, bw2 = s2 % sw;
      int tl = to * st + a2;
      if (dup_first) tl = tl > 0 ? tl - 1 : 0;
      const int64_t xpos = ((static_cast<int64_t>(b) * T + tl) * H + ho * sh + bh2) * W + wo * sw + bw2;
      m += __bfloat162float(x[xpos * Cx + ci]);
    }
    out[i] = __float2bfloat16(acc + m / G);
  }
}

// PixelShuffle2d (spatial.py:184-218): y [B,F,H,W,4C] -> out [B,F,2H,2W,C], source channel c*4 + r_h*2 + r_w
__global__ void pixel_shuffle2_kernel(const bf16* __restrict__ y, bf16* __restrict__ out, int64_t BF, int H, int W, int C) {
  const int64_t n = BF * 2 * H * 2 * W * C;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = i % C;
    int64_t r = i / C;
    const int w2 = r % (2 * W); r /= 2 * W;
    const int h2 = r % (2 * H);
    const int64_t bf = r / (2 * H);
    out[i] = y[((bf * H + h2 / 2) * W + w2 / 2) * (4 * static_cast<int64_t>(C)) + c * 4 + (h2 & 1) * 2 + (w2 & 1)];
  }
}

// channels-last bf16 [B,T,H,W,Cs] -> fp32 NCDHW [B,C,T,H,W] (first C channels), optional (x - mean[c]) / std[c]
__global__ void ndhwc_to_ncdhw_kernel(const bf16* __restrict__ x, float* __restrict__ out, int B, int C, int Cs, int64_t thw,
                                      const float* __restrict__ mean, const float* __restrict__ stdv) {
  const int64_t n = static_cast<int64_t>(B) * C * thw;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t p = i % thw;
    const int c = (i / thw) % C;
    const int b = i / (thw * C);
    float v = __bfloat162float(x[(static_cast<int64_t>(b) * thw + p) * Cs + c]);
    if (mean != nullptr) v = (v - mean[c]) / stdv[c];
    out[i] = v;
  }
}

template <typename Tin>
__global__ void ncdhw_to_ndhwc_kernel(const Tin* __restrict__ x, bf16* __restrict__ out, int B, int C, int64_t thw) {
  const int64_t n = static_cast<int64_t>(B) * C * thw;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = i % C;
    const int64_t p = (i / C) % thw;
    const int b = i / (thw * C);
    out[i] = __float2bfloat16(static_cast<float>(x[(static_cast<int64_t>(b) * C + c) * thw + p]));
  }
}

inline unsigned grid_for(int64_t n, int per_block = kAuxThreads) {
  return static_cast<unsigned>(std::min<int64_t>((n + per_block - 1) / per_block, static_cast<int64_t>(num_sms()) * 32));
}
inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace

extern "C" {

int ltx2_pad_act(const void* x, void* out_padded, void* out_plain, int32_t B, int32_t T, int32_t H, int32_t W, int32_t C,
                 int32_t hw_mode, int32_t t_mode, int32_t dup_first, int32_t act, const float* gn_stats,
                 const float* gn_weight, const float* gn_bias, int32_t groups, float eps, const void* residual,
                 void* stream) {
  LTX2_REQUIRE(x && out_padded && B >= 1 && T >= 1 && H >= 1 && W >= 1 && C % 8 == 0, "pad_act: bad argument");
  LTX2_REQUIRE(hw_mode >= 0 && hw_mode <= 1 && t_mode >= 0 && t_mode <= 2 && act >= 0 && act <= 3, "pad_act: bad mode");
  LTX2_REQUIRE(hw_mode == 1 || (H >= 2 && W >= 2), "pad_act: reflect padding needs H, W >= 2");
  if (act >= 2)
    LTX2_REQUIRE(gn_stats && gn_weight && gn_bias && groups >= 1 && C % groups == 0 && (C / groups) % 8 == 0,
                 "pad_act: GroupNorm needs stats/weight/bias and 8 | C/groups");
  LTX2_REQUIRE(out_plain == nullptr || !dup_first, "pad_act: the un-padded copy is not available with dup_first");
  const int TL = T + (dup_first ? 1 : 0);
  const int64_t n_pos = static_cast<int64_t>(B) * (TL + 2) * (H + 2) * (W + 2);
  pad_act_kernel<<<grid_for(n_pos, kAuxThreads / 32), kAuxThreads, 0, S(stream)>>>(
      reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(out_padded), reinterpret_cast<bf16*>(out_plain), B, T, H,
      W, C, hw_mode, t_mode, dup_first, act, gn_stats, gn_weight, gn_bias, groups, eps,
      reinterpret_cast<const bf16*>(residual));
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int ltx2_group_stats(const void* x, int32_t B, int64_t thw, int32_t C, int32_t groups, float eps, float* stats,
                     void* stream) {
  LTX2_REQUIRE(x && stats && B >= 1 && thw >= 1 && groups >= 1 && C % groups == 0 && (C / groups) % 8 == 0,
               "group_stats: bad argument (8 | C/groups)");
  group_stats_kernel<<<B * groups, kAuxThreads, 0, S(stream)>>>(reinterpret_cast<const bf16*>(x), thw, C, groups, eps,
                                                              stats);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

// packed weight/bias for ltx2_conv3d_packed: weight [Cout,Cin,3,3,3] (dtype code) -> bf16 [Cout_pad, 27*Cin]
int ltx2_conv3d_pack(const void* weight, int32_t w_dtype, const void* bias, int32_t b_dtype, int32_t Cout,
                     int32_t Cout_pad, int32_t Cin, void* w_packed, float* b_packed, void* stream) {
  LTX2_REQUIRE(weight && bias && w_packed && b_packed && Cout_pad % 32 == 0 && Cout_pad >= Cout && Cin % 64 == 0,
               "conv3d_pack: bad argument (C_in %% 64, C_out_pad %% 32)");
  LTX2_PROPAGATE(pack_conv_weight(weight, w_dtype, w_packed, Cout, Cout_pad, Cin, 1, S(stream)));
  return pack_conv_bias(bias, b_dtype, b_packed, Cout, Cout_pad, 1, S(stream));
}

// the tcgen05 implicit-GEMM conv on an already padded input [B,T+2,H+2,W+2,Cin]: out [B,T,H,W,Cout] bf16
// (+ residual [B,T,H,W,Cout] when given)
int ltx2_conv3d_packed(const void* x_padded, const void* w_packed, const float* b_packed, void* out, const void* residual,
                       int32_t B, int32_t T, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t Cout_pad,
                       void* stream) {
  LTX2_REQUIRE(x_padded && w_packed && b_packed && out && Cout % 8 == 0, "conv3d_packed: bad argument (C_out %% 8)");
  ConvParams p;
  p.B = B; p.T = T; p.H = H; p.W = W;
  p.Cin = Cin; p.Cout = Cout; p.Cout_pad = Cout_pad;
  p.mode = residual ? CONV_EPI_RESIDUAL : CONV_EPI_PLAIN;
  p.bias = b_packed;
  p.out = reinterpret_cast<bf16*>(out);
  p.residual = reinterpret_cast<const bf16*>(residual);
  return conv3d_bf16(x_padded, w_packed, p, S(stream));
}

int ltx2_patchify_video(const float* video, void* out, int32_t B, int32_t F, int32_t H, int32_t W, int32_t Cp,
                        void* stream) {
  LTX2_REQUIRE(video && out && H % 4 == 0 && W % 4 == 0 && Cp >= 48, "patchify_video: bad argument");
  const int64_t n = static_cast<int64_t>(B) * F * (H / 4) * (W / 4) * Cp;
  patchify_video_kernel<<<grid_for(n), kAuxThreads, 0, S(stream)>>>(video, reinterpret_cast<bf16*>(out), B, F, H, W, Cp);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int ltx2_space_to_depth_residual(const void* y, const void* x, void* out, int32_t B, int32_t T, int32_t H, int32_t W,
                                 int32_t Cx, int32_t Cout, int32_t st, int32_t sh, int32_t sw, int32_t dup_first,
                                 void* stream) {
  const int sp = st * sh * sw;
  const int TL = T + (dup_first ? 1 : 0);
  LTX2_REQUIRE(y && x && out && sp >= 1 && Cout % sp == 0 && (Cx * sp) % Cout == 0 && TL % st == 0 && H % sh == 0 &&
                   W % sw == 0,
               "space_to_depth_residual: shapes do not divide");
  const int64_t n = static_cast<int64_t>(B) * (TL / st) * (H / sh) * (W / sw) * Cout;
  s2d_residual_kernel<<<grid_for(n), kAuxThreads, 0, S(stream)>>>(reinterpret_cast<const bf16*>(y),
                                                                 reinterpret_cast<const bf16*>(x),
                                                                 reinterpret_cast<bf16*>(out), B, T, TL, H, W, Cx, Cout, st,
                                                                 sh, sw, dup_first);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int ltx2_pixel_shuffle2(const void* y, void* out, int64_t BF, int32_t H, int32_t W, int32_t C, void* stream) {
  LTX2_REQUIRE(y && out && BF >= 1 && C >= 1, "pixel_shuffle2: bad argument");
  const int64_t n = BF * 4 * H * W * C;
  pixel_shuffle2_kernel<<<grid_for(n), kAuxThreads, 0, S(stream)>>>(reinterpret_cast<const bf16*>(y),
                                                                   reinterpret_cast<bf16*>(out), BF, H, W, C);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int ltx2_ndhwc_to_ncdhw(const void* x, float* out, int32_t B, int32_t C, int32_t Cs, int64_t thw, const float* mean,
                        const float* stdv, void* stream) {
  LTX2_REQUIRE(x && out && C <= Cs && (mean == nullptr) == (stdv == nullptr), "ndhwc_to_ncdhw: bad argument");
  const int64_t n = static_cast<int64_t>(B) * C * thw;
  ndhwc_to_ncdhw_kernel<<<grid_for(n), kAuxThreads, 0, S(stream)>>>(reinterpret_cast<const bf16*>(x), out, B, C, Cs, thw,
                                                                   mean, stdv);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int ltx2_ncdhw_to_ndhwc(const void* x, int32_t dtype, void* out, int32_t B, int32_t C, int64_t thw, void* stream) {
  LTX2_REQUIRE(x && out, "ncdhw_to_ndhwc: null argument");
  const int64_t n = static_cast<int64_t>(B) * C * thw;
  bf16* o = reinterpret_cast<bf16*>(out);
  switch (dtype) {
    case LTX2_F32: ncdhw_to_ndhwc_kernel<float><<<grid_for(n), kAuxThreads, 0, S(stream)>>>(reinterpret_cast<const float*>(x), o, B, C, thw); break;
    case LTX2_BF16: ncdhw_to_ndhwc_kernel<bf16><<<grid_for(n), kAuxThreads, 0, S(stream)>>>(reinterpret_cast<const bf16*>(x), o, B, C, thw); break;
    case LTX2_F16: ncdhw_to_ndhwc_kernel<__half><<<grid_for(n), kAuxThreads, 0, S(stream)>>>(reinterpret_cast<const __half*>(x), o, B, C, thw); break;
    default: set_error("ncdhw_to_ndhwc: bad dtype %d", dtype); return LTX2_ERR_INVALID;
  }
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

}  // extern "C"
