// 3x3x3 convolution of the video-VAE decoder as an implicit GEMM on tcgen05 tensor cores.
//
// Reference op replaced: Conv3dSimple.__call__ (model/video_vae/simple_decoder.py:90-180), i.e. 3 x mx.conv2d
// per conv over a reflect(H,W)/replicate(T)-padded input, plus what follows it in the decoder:
//   bias (:178), ResBlock residual add (:240), DepthToSpaceUpsample3d rearrange + first-frame drop + tiled
//   residual (:274-313), conv_out + unpatchify (:545-552, ops.py:70-131).
//
// Data layout: activations are channels-last bf16 [B, T, H, W, C].  The producer of a conv input writes it
// already PADDED ([B, T+2, H+2, W+2, C], vae_rowops.cu), so every filter tap is a plain shifted TMA box:
//   A tile (128 output positions = 16 rows x 8 columns of one frame) for tap (kt,kh,kw), channels [c0,c0+64):
//     4-D box {64, 8, 16, 1} at {c0, w0+kw, h0+kh, b*(T+2)+t+kt}  ->  128 x 128 B rows, 128B swizzle (K-major)
//   B tile: weights re-laid at load time as [C_out, 27*C_in] (tap-major, channel-minor) -> 2-D box {64, BN}
// so the main loop is the GEMM's (gemm_sm100.cu): TMA producer warp, single-thread tcgen05.mma issuer with fp32
// accumulators in TMEM (double-buffered), 4 epilogue warps.  K = 27*C_in.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace ltx2 {

namespace {

constexpr int CBM = 128, CBK = 64, CTW = 8, CTH = 16;
constexpr int kConvThreads = 192;

template <int BN>
struct ConvCfg {
  static constexpr int kStages = (BN >= 256) ? 4 : 6;
  static constexpr int kABytes = CBM * CBK * 2;
  static constexpr int kBBytes = BN * CBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + 3 * BN * 4;
};

// Rows the fused epilogue needs for every position of batch element b, staged once in shared memory by the 128
// epilogue threads (r = 0..127): bias, shift and 1 + scale.  Named barrier 1 orders it against the readers.
template <int BN>
__device__ __forceinline__ void conv_stage_mod(const ConvParams& p, int b, float* s_mod, int r) {
  asm volatile("bar.sync 1, 128;" ::: "memory");
  for (int c = r; c < BN; c += 128) {
    s_mod[c] = p.bias[c];
    if (p.pad_act) {
      const float* m = p.pad_mod + static_cast<int64_t>(b) * p.pad_mod_stride;
      s_mod[BN + c] = m[p.pad_shift_off + c];
      s_mod[2 * BN + c] = 1.f + m[p.pad_scale_off + c];
    }
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");
}

// Epilogue of one accumulator tile for the calling thread's row (= one output position): TMEM -> bias / residual /
// depth-to-space / unpatchify (/ fused padded producer) -> global.  Shared by the 1-CTA and the SM-pair kernels.
template <int BN>
__device__ __forceinline__ void conv_epilogue_row(const ConvParams& p, uint32_t t_row, int mt, int nt, int r, bool tile_ok,
                                                  const float* s_mod, uint64_t* acc_full_bar, uint32_t acc_phase) {
  const int tiles_w = (p.W + CTW - 1) / CTW;
  const int tiles_h = (p.H + CTH - 1) / CTH;
  const int sp = p.ft * p.fh * p.fw;
  const int wx = mt % tiles_w;
  const int hy = (mt / tiles_w) % tiles_h;
  const int bt = mt / (tiles_w * tiles_h);
  const int b = bt / p.T, t = bt % p.T;
  const int h = hy * CTH + r / CTW, w = wx * CTW + r % CTW;
  const bool pos_ok = tile_ok && h < p.H && w < p.W;
  const int64_t pos = ((static_cast<int64_t>(bt) * p.H + h) * p.W + w);     // unpadded NDHWC position index
  if (p.pad_out != nullptr) {
    // ---- fused producer: this thread's accumulator row holds ALL channels of its position (BN == C_out) ----
    // s_mod (shared memory, staged by conv_stage_mod): [0,BN) bias, [BN,2BN) shift, [2BN,3BN) 1 + scale
    // The residual row comes from DRAM (written two convs ago): its first chunk is requested BEFORE the wait for the
    // accumulator, the rest of the row is pulled into L2 meanwhile, and chunk c+1 is loaded while chunk c is processed.
    const bool has_res = p.mode == CONV_EPI_RESIDUAL && pos_ok;
    const __nv_bfloat16* rs = p.residual + pos * p.Cout;
    uint4 rnext[4];
    if (has_res) {
#pragma unroll
      for (int j = 0; j < 4; ++j) rnext[j] = *reinterpret_cast<const uint4*>(rs + 8 * j);
#pragma unroll
      for (int l = 128; l < BN * 2; l += 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(rs) + l));
    }
    mbar_wait(acc_full_bar, acc_phase);
    tc_fence_after();
    // pass 1: v = bf16(acc + bias [+ residual]) -> raw output (optional), sum of squares, v back into TMEM
    float ss = 0.f;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t rr[32];
      tmem_ld_32x32(t_row + c, rr);
      uint4 rcur[4];
      if (has_res) {
#pragma unroll
        for (int j = 0; j < 4; ++j) rcur[j] = rnext[j];
        if (c + 32 < BN) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rnext[j] = *reinterpret_cast<const uint4*>(rs + c + 32 + 8 * j);
        }
      }
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 bb = *reinterpret_cast<const float4*>(s_mod + c + j);
        v[j] = __uint_as_float(rr[j]) + bb.x;
        v[j + 1] = __uint_as_float(rr[j + 1]) + bb.y;
        v[j + 2] = __uint_as_float(rr[j + 2]) + bb.z;
        v[j + 3] = __uint_as_float(rr[j + 3]) + bb.w;
      }
      if (has_res) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&rcur[j]);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = __bfloat1622float2(hh[q]);
            v[8 * j + 2 * q] += f.x;
            v[8 * j + 2 * q + 1] += f.y;
          }
        }
      }
      // the normalisation sees exactly the stored (bf16) activation, like the separate norm_act_pad pass
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk[j]));
        rr[2 * j] = __float_as_uint(f.x);
        rr[2 * j + 1] = __float_as_uint(f.y);
        ss = fmaf(f.x, f.x, fmaf(f.y, f.y, ss));
      }
      if (p.out != nullptr && pos_ok) {
        __nv_bfloat16* o = p.out + pos * p.Cout + c;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(o + 8 * j) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      }
      tmem_st_32x32(t_row + c, rr);
    }
    tmem_st_wait();
    const float rstd = rsqrtf(ss * (1.0f / BN) + p.pad_eps);
    // destinations in the padded tensor: the interior copy plus the border copies this position feeds
    // (reflect: row 1 -> pad row 0, row H-2 -> pad row H+1; T: first / last frame replicated)
    const int Hp = p.H + 2, Wp = p.W + 2;
    int tps[3], nt_ = 0, hps[3], nh_ = 0, wps[3], nw_ = 0;
    if (p.pad_causal) {
      tps[nt_++] = t + 2;
      if (t == 0) { tps[nt_++] = 0; tps[nt_++] = 1; }
    } else {
      tps[nt_++] = t + 1;
      if (t == 0 && !p.pad_skip_front) tps[nt_++] = 0;
      if (t == p.T - 1 && !p.pad_skip_back) tps[nt_++] = p.T + 1;
    }
    hps[nh_++] = h + 1;
    if (h == 1) hps[nh_++] = 0;
    if (h == p.H - 2) hps[nh_++] = p.H + 1;
    wps[nw_++] = w + 1;
    if (w == 1) wps[nw_++] = 0;
    if (w == p.W - 2) wps[nw_++] = p.W + 1;
    // (with H == 3 row 1 is both "1" and "H-2": three copies along that axis)
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t rr[32];
      tmem_ld_32x32(t_row + c, rr);
      tmem_ld_wait();
      if (!pos_ok) continue;
      uint32_t pk[16];
      if (p.pad_act) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 sh = *reinterpret_cast<const float4*>(s_mod + BN + c + j);
          const float4 sc = *reinterpret_cast<const float4*>(s_mod + 2 * BN + c + j);       // 1 + scale
          float y0 = fmaf(__uint_as_float(rr[j]) * rstd, sc.x, sh.x);
          float y1 = fmaf(__uint_as_float(rr[j + 1]) * rstd, sc.y, sh.y);
          float y2 = fmaf(__uint_as_float(rr[j + 2]) * rstd, sc.z, sh.z);
          float y3 = fmaf(__uint_as_float(rr[j + 3]) * rstd, sc.w, sh.w);
          y0 = __fdividef(y0, 1.f + __expf(-y0));
          y1 = __fdividef(y1, 1.f + __expf(-y1));
          y2 = __fdividef(y2, 1.f + __expf(-y2));
          y3 = __fdividef(y3, 1.f + __expf(-y3));
          pk[j / 2] = pack_bf16x2(y0, y1);
          pk[j / 2 + 1] = pack_bf16x2(y2, y3);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(__uint_as_float(rr[2 * j]), __uint_as_float(rr[2 * j + 1]));
      }
      for (int a = 0; a < nt_; ++a)
        for (int bh = 0; bh < nh_; ++bh)
          for (int bw = 0; bw < nw_; ++bw) {
            __nv_bfloat16* o = p.pad_out +
                (((static_cast<int64_t>(b) * (p.T + 2) + tps[a]) * Hp + hps[bh]) * Wp + wps[bw]) * p.Cout + c;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(o + 8 * j) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          }
    }
    return;
  }
  mbar_wait(acc_full_bar, acc_phase);
  tc_fence_after();
#pragma unroll 1
  for (int c = 0; c < BN; c += 32) {
    uint32_t rr[32];
    tmem_ld_32x32(t_row + c, rr);
    tmem_ld_wait();
    const int col0 = nt * BN + c;
    if (!pos_ok || col0 >= p.Cout_pad) continue;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
      v[j] = __uint_as_float(rr[j]) + bb.x;
      v[j + 1] = __uint_as_float(rr[j + 1]) + bb.y;
      v[j + 2] = __uint_as_float(rr[j + 2]) + bb.z;
      v[j + 3] = __uint_as_float(rr[j + 3]) + bb.w;
    }
    if (p.mode == CONV_EPI_PLAIN || p.mode == CONV_EPI_RESIDUAL) {
      const int nv = p.Cout - col0;                  // valid columns of this group (C_out % 8 == 0; rows >= C_out are padding)
      if (nv <= 0) continue;
      if (p.mode == CONV_EPI_RESIDUAL) {
        const __nv_bfloat16* rs = p.residual + pos * p.Cout + col0;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (j >= nv) break;
          const uint4 u = *reinterpret_cast<const uint4*>(rs + j);
          const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = __bfloat1622float2(hh[q]);
            v[j + 2 * q] += f.x;
            v[j + 2 * q + 1] += f.y;
          }
        }
      }
      __nv_bfloat16* o = p.out + pos * p.Cout + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        if (j >= nv) break;
        uint4 q;
        q.x = pack_bf16x2(v[j], v[j + 1]);
        q.y = pack_bf16x2(v[j + 2], v[j + 3]);
        q.z = pack_bf16x2(v[j + 4], v[j + 5]);
        q.w = pack_bf16x2(v[j + 6], v[j + 7]);
        *reinterpret_cast<uint4*>(o + j) = q;
      }
    } else if (p.mode == CONV_EPI_D2S) {
      // weight rows were permuted at load time to (sub-position major, output channel minor):
      // col = sub * Cf + c, sub = (a*fh + bb)*fw + d  ->  32 consecutive columns share one output pixel.
      const int Cf = p.Cout / sp;
      const int sub = col0 / Cf, c0 = col0 % Cf;
      const int a = sub / (p.fh * p.fw), bb2 = (sub / p.fw) % p.fh, d = sub % p.fw;
      const int drop = (p.ft > 1 && !p.d2s_keep_first) ? 1 : 0;
      const int to = t * p.ft + a - drop;
      if (to < 0) continue;
      const int To = p.T * p.ft - drop, Ho = p.H * p.fh, Wo = p.W * p.fw;
      if (p.c_d2s > 0) {
        // residual = depth_to_space(x) tiled along channels: source channel (c % c_d2s)*sp + sub of THIS position
        const __nv_bfloat16* rs = p.residual + pos * p.Cin;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += __bfloat162float(rs[((c0 + j) % p.c_d2s) * sp + sub]);
      }
      __nv_bfloat16* o = p.out + (((static_cast<int64_t>(b) * To + to) * Ho + (h * p.fh + bb2)) * Wo + (w * p.fw + d)) * Cf + c0;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 q;
        q.x = pack_bf16x2(v[j], v[j + 1]);
        q.y = pack_bf16x2(v[j + 2], v[j + 3]);
        q.z = pack_bf16x2(v[j + 4], v[j + 5]);
        q.w = pack_bf16x2(v[j + 6], v[j + 7]);
        *reinterpret_cast<uint4*>(o + j) = q;
      }
    } else {  // CONV_EPI_UNPATCHIFY: col = (ch*4 + rw)*4 + rh -> out_f32[b, ch, t, h*4+rh, w*4+rw]
      const int Hp = p.H * 4, Wp = p.W * 4;
#pragma unroll
      for (int j = 0; j < 32; j += 16) {
        const int ch = (col0 + j) / 16;
        if (ch >= p.Cout / 16) break;
        // out_t_total > 0: this launch produces frames [out_t0, out_t0 + T) of a longer clip (temporal shards)
        const int Tall = p.out_t_total > 0 ? p.out_t_total : p.T;
        float* o = p.out_f32 +
                   ((static_cast<int64_t>(b) * (p.Cout / 16) + ch) * Tall + p.out_t0 + t) * Hp * static_cast<int64_t>(Wp);
#pragma unroll
        for (int rh = 0; rh < 4; ++rh)
          *reinterpret_cast<float4*>(o + static_cast<int64_t>(h * 4 + rh) * Wp + w * 4) =
              make_float4(v[j + rh], v[j + 4 + rh], v[j + 8 + rh], v[j + 12 + rh]);
      }
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3d_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, ConvParams p) {
  using Cfg = ConvCfg<BN>;
  pdl_trigger();                                     // programmatic dependent launch, see common.cuh
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::kStages * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* acc_full = bars + 2 * Cfg::kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_mod = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes + 256);   // [3][BN] (fused epilogue)

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform role (see gemm_sm100.cu)
  const int lane = threadIdx.x & 31;
  const int tiles_w = (p.W + CTW - 1) / CTW;
  const int tiles_h = (p.H + CTH - 1) / CTH;
  const int num_m = p.B * p.T * tiles_h * tiles_w;
  const int num_n = (p.Cout_pad + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int cchunks = p.Cin / CBK;
  const int num_kb = 27 * cchunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                        // the set-up above overlapped the previous kernel's tail

  if (warp == 0) {
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mt = tile % num_m, nt = tile / num_m;
      const int wx = mt % tiles_w;
      const int hy = (mt / tiles_w) % tiles_h;
      const int bt = mt / (tiles_w * tiles_h);          // b*T + t
      const int b = bt / p.T, t = bt % p.T;
      const int plane0 = b * (p.T + 2) + t;
      int tap = 0, cc = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int kt = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) {
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_4d(smem_a + stage * Cfg::kABytes, &tmap_x, &full_bar[stage], cc * CBK, wx * CTW + kw,
                      hy * CTH + kh, plane0 + kt);
          tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmap_w, &full_bar[stage], kb * CBK, nt * BN);
        }
        if (++cc == cchunks) { cc = 0; ++tap; }
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    constexpr uint32_t idesc = umma_idesc_bf16(CBM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t adesc = umma_desc_k_sw128(smem_u32(smem_a + stage * Cfg::kABytes));
        const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smem_b + stage * Cfg::kBBytes));
        if (leader) {
#pragma unroll
          for (int k = 0; k < CBK / 16; ++k) umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      if (leader) umma_commit(&acc_full[acc]);
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    int staged_b = -1;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mt = tile % num_m, nt = tile / num_m;
      if (p.pad_out != nullptr) {
        const int b = mt / (tiles_w * tiles_h * p.T);
        if (b != staged_b) {
          conv_stage_mod<BN>(p, b, s_mod, quarter * 32 + lane);
          staged_b = b;
        }
      }
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
      conv_epilogue_row<BN>(p, t_row, mt, nt, quarter * 32 + lane, true, s_mod, &acc_full[acc], acc_phase);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// SM-pair variant (tcgen05 cta_group::2) for convs whose C_out fits one weight tile (BN == C_out_pad: the 128- and
// 256-channel stages).  With 128 x 128 tiles a single CTA reads 4 KB of A and 4 KB of B from shared memory per 64-clock
// MMA -- 128 B/clk, the shared-memory limit (tensor pipe 67 % active on the last stage, profiles/r1c_kernels.json).  A
// cluster of two CTAs issues ONE MMA with M = 256: CTA r supplies ITS 128 output positions (A) and HALF of the weight
// rows (B: BN/2 rows), so each SM reads 4 + 2 KB per MMA and receives 24 instead of 32 KB per K block from TMA / L2.
// Roles as in gemm2_bf16_kernel (gemm_sm100.cu):
//   warp 0 (both CTAs)  TMA producer of the CTA's A tile and weight half -> own `full` barriers
//   warp 1, CTA 1       relay: own `full` complete -> remote arrive on CTA 0's `peer_full`
//   warp 1, CTA 0       MMA issuer: waits full + peer_full, commits multicast to `empty` / `acc_full` of both CTAs
//   warps 2-5           epilogue of the CTA's own 128 positions; `acc_empty` lives in CTA 0, counts all 8 epilogue warps
// ---------------------------------------------------------------------------------------------------------------
// KH3: one pipeline stage carries the THREE kernel rows kh = 0, 1, 2 of a (kt, kw, channel chunk): the activation box is
// 18 rows high instead of 16 (18 KB instead of 3 x 16 KB), and tap kh reads it at row offset kh -- 8 positions x 128 B =
// 1024 B, exactly one swizzle atom, so the shifted operand needs nothing but a start address.  The implicit GEMM
// otherwise re-reads every activation 27 times from L2 (9.9 TB/s on the 128-channel stage: the measured bound behind
// the 66 % tensor-pipe activity, profiles/r2_kernels.json); this cuts the L2 -> SM activation traffic by 2.67x.
template <int BN, bool KH3 = true>
struct ConvPairCfg {
  static constexpr int kTaps = KH3 ? 3 : 1;
  static constexpr int kABytes = KH3 ? (CTH + 2) * CTW * CBK * 2 : CBM * CBK * 2;     // 18 KB / 16 KB
  static constexpr int kBTap = (BN / 2) * CBK * 2;                                      // one tap's weight half
  static constexpr int kBBytes = kTaps * kBTap;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = KH3 ? (BN >= 256 ? 3 : 5) : 6;
  static constexpr int kTmemCols = 2 * BN;                  // 256 or 512
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 512 + 3 * BN * 4;
};

template <int BN, bool KH3>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
conv3d_pair_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, ConvParams p) {
  using Cfg = ConvPairCfg<BN, KH3>;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::kStages * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                          // [stages]  own TMA -> MMA (CTA 0) / relay (CTA 1)
  uint64_t* peer_full = full_bar + Cfg::kStages;      // [stages]  CTA 1 relay -> CTA 0 MMA (used in CTA 0 only)
  uint64_t* empty_bar = peer_full + Cfg::kStages;     // [stages]  MMA (multicast) -> own TMA
  uint64_t* acc_full = empty_bar + Cfg::kStages;      // [2]       MMA (multicast) -> own epilogue
  uint64_t* acc_empty = acc_full + 2;                 // [2]       all 8 epilogue warps -> MMA (used in CTA 0 only)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_mod = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes + 512);   // [3][BN] (fused epilogue)

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int tiles_w = (p.W + CTW - 1) / CTW;
  const int tiles_h = (p.H + CTH - 1) / CTH;
  const int num_m = p.B * p.T * tiles_h * tiles_w;
  const int num_pairs = (num_m + 1) / 2;              // pair pt = position tiles 2 pt (CTA 0) and 2 pt + 1 (CTA 1)
  const int cchunks = p.Cin / CBK;
  const int num_kb = (KH3 ? 9 : 27) * cchunks;        // KH3: a K block = (kt, kw, channel chunk) with all three kh

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&peer_full[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 8);
    }
    fence_barrier_init();
  }
  __syncthreads();
  cluster_sync_all();                                  // both CTAs are running and their barriers exist
  if (warp == 1) tmem_alloc2(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ============ TMA producer (both CTAs): own position tile, own half of the weight rows ============
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int pt = cluster_id; pt < num_pairs; pt += num_clusters) {
      const int mt = min(2 * pt + static_cast<int>(rank), num_m - 1);     // odd tile count: CTA 1 repeats the last tile
      const int wx = mt % tiles_w;
      const int hy = (mt / tiles_w) % tiles_h;
      const int bt = mt / (tiles_w * tiles_h);
      const int b = bt / p.T, t = bt % p.T;
      const int plane0 = b * (p.T + 2) + t;
      int tap = 0, cc = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) {
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          if (KH3) {
            // tap = kt*3 + kw; the 18-row box covers kh = 0..2; the weights of the three taps (kt, kh, kw) sit 3*C_in
            // apart in the packed [C_out, 27*C_in] matrix
            const int kt = tap / 3, kw = tap % 3;
            tma_load_4d(smem_a + stage * Cfg::kABytes, &tmap_x, &full_bar[stage], cc * CBK, wx * CTW + kw, hy * CTH,
                        plane0 + kt);
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
              tma_load_2d(smem_b + stage * Cfg::kBBytes + kh * Cfg::kBTap, &tmap_w, &full_bar[stage],
                          ((kt * 3 + kh) * 3 + kw) * p.Cin + cc * CBK, static_cast<int>(rank) * (BN / 2));
          } else {
            const int kt = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
            tma_load_4d(smem_a + stage * Cfg::kABytes, &tmap_x, &full_bar[stage], cc * CBK, wx * CTW + kw, hy * CTH + kh,
                        plane0 + kt);
            tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmap_w, &full_bar[stage], kb * CBK,
                        static_cast<int>(rank) * (BN / 2));
          }
        }
        if (++cc == cchunks) { cc = 0; ++tap; }
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && rank == 1) {
    // ===================== relay (CTA 1): my tiles have landed -> tell the MMA issuer in CTA 0 =====================
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int pt = cluster_id; pt < num_pairs; pt += num_clusters) {
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        if (leader) mbar_arrive_cluster_relaxed(mapa_u32(&peer_full[stage], 0));
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (CTA 0): M = 256 positions of the pair, N = BN channels =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc = umma_idesc_bf16(256, BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int pt = cluster_id; pt < num_pairs; pt += num_clusters) {
      mbar_wait_cluster(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        mbar_wait_cluster(&peer_full[stage], phase);
        tc_fence_after();
        const uint64_t adesc = umma_desc_k_sw128(smem_u32(smem_a + stage * Cfg::kABytes));
        const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smem_b + stage * Cfg::kBBytes));
        if (leader) {
#pragma unroll
          for (int kh = 0; kh < Cfg::kTaps; ++kh)
#pragma unroll
            for (int k = 0; k < CBK / 16; ++k)
              // tap kh: activation rows shifted by kh * 8 positions (one 1024-byte swizzle atom), its own weight tile
              umma2_bf16_ss(d_tmem, adesc + kh * (1024 >> 4) + 2 * k, bdesc + kh * (Cfg::kBTap >> 4) + 2 * k, idesc,
                            (kb | kh | k) != 0);
          umma2_commit_both(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      if (leader) umma2_commit_both(&acc_full[acc]);
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..5, both CTAs): the CTA's own 128 positions =====================
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    int staged_b = -1;
    for (int pt = cluster_id; pt < num_pairs; pt += num_clusters) {
      const int mt = 2 * pt + static_cast<int>(rank);
      const int mtc = min(mt, num_m - 1);
      if (p.pad_out != nullptr) {
        const int b = mtc / (tiles_w * tiles_h * p.T);
        if (b != staged_b) {
          conv_stage_mod<BN>(p, b, s_mod, quarter * 32 + lane);
          staged_b = b;
        }
      }
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
      conv_epilogue_row<BN>(p, t_row, mtc, 0, quarter * 32 + lane, mt < num_m, s_mod, &acc_full[acc], acc_phase);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(&acc_empty[acc], 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                  // nobody leaves while the peer may still touch my barriers / smem
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN, bool KH3>
int launch_conv_pair(const CUtensorMap& tx, const CUtensorMap& tw, const ConvParams& p, cudaStream_t stream) {
  using Cfg = ConvPairCfg<BN, KH3>;
  static PerDeviceOnce configured;
  if (configured.first()) {
    LTX2_CUDA_CHECK(cudaFuncSetAttribute(conv3d_pair_kernel<BN, KH3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes));
  }
  const int num_m = p.B * p.T * ((p.H + CTH - 1) / CTH) * ((p.W + CTW - 1) / CTW);
  const int pairs = (num_m + 1) / 2, clusters = num_sms() / 2;
  const int grid = 2 * (pairs < clusters ? pairs : clusters);
  LTX2_CUDA_CHECK(launch_pdl(conv3d_pair_kernel<BN, KH3>, dim3(grid), dim3(kConvThreads), Cfg::kSmemBytes, stream, tx, tw,
                             p));
  count_launch();
  return LTX2_OK;
}

template <int BN>
int launch_conv(const CUtensorMap& tx, const CUtensorMap* tw, const ConvParams& p, cudaStream_t stream) {
  using Cfg = ConvCfg<BN>;
  static PerDeviceOnce configured;
  if (configured.first()) {
    LTX2_CUDA_CHECK(cudaFuncSetAttribute(conv3d_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes));
  }
  const int tiles = p.B * p.T * ((p.H + CTH - 1) / CTH) * ((p.W + CTW - 1) / CTW) * ((p.Cout_pad + BN - 1) / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  LTX2_CUDA_CHECK(launch_pdl(conv3d_kernel<BN>, dim3(grid), dim3(kConvThreads), Cfg::kSmemBytes, stream, tx, *tw, p));
  count_launch();
  return LTX2_OK;
}

}  // namespace

int conv3d_bf16(const void* x_padded, const void* w_packed, const ConvParams& p, cudaStream_t stream) {
  LTX2_REQUIRE(p.Cin % 64 == 0, "conv3d: C_in=%d must be a multiple of 64", p.Cin);
  LTX2_REQUIRE(p.Cout_pad % 32 == 0 && p.Cout_pad >= p.Cout, "conv3d: padded C_out=%d invalid", p.Cout_pad);
  LTX2_REQUIRE(p.B > 0 && p.T > 0 && p.H > 0 && p.W > 0, "conv3d: empty input");
  CUtensorMap tx;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(p.Cin), static_cast<uint64_t>(p.W + 2), static_cast<uint64_t>(p.H + 2),
                        static_cast<uint64_t>(p.B) * (p.T + 2)};
    uint64_t str[3] = {static_cast<uint64_t>(p.Cin) * 2, static_cast<uint64_t>(p.W + 2) * p.Cin * 2,
                       static_cast<uint64_t>(p.H + 2) * (p.W + 2) * p.Cin * 2};
    uint32_t box[4] = {CBK, CTW, CTH, 1};
    LTX2_PROPAGATE(make_tensor_map_bf16(&tx, x_padded, 4, dims, str, box));
  }
  if (p.pad_out != nullptr) {
    LTX2_REQUIRE((p.mode == CONV_EPI_PLAIN || p.mode == CONV_EPI_RESIDUAL) && p.Cout == p.Cout_pad &&
                     (p.Cout == 128 || p.Cout == 256) && p.H >= 2 && p.W >= 2,
                 "conv3d: the fused padded output needs C_out = 128 or 256 (got %d), H, W >= 2 and a plain/residual epilogue",
                 p.Cout);
    LTX2_REQUIRE(!p.pad_act || p.pad_mod != nullptr, "conv3d: fused activation without modulation rows");
  } else {
    LTX2_REQUIRE(p.out != nullptr || p.out_f32 != nullptr, "conv3d: null output");
  }
  int bn = 256;
  if (p.Cout_pad % 256 != 0) bn = 128;
  if (p.Cout_pad % 128 != 0) bn = 64;
  if (p.Cout_pad % 64 != 0) bn = 32;
  if (p.pad_out == nullptr && bn > 64) {
    // Few position tiles (the 1024- / 512-channel stages of a temporal shard: 3-24 tiles): a tile's K loop is serial, so
    // narrower weight tiles put more SMs on the conv.  Cost model in clocks per K step of the 1-CTA kernel as MEASURED
    // on whole convs (N 256 -> 128; N 128 -> 86 and N 64 -> 56: shared-memory / L2 bound, profiles/r2_vae_ab_kh3.txt),
    // times the number of waves -- a conv with many tiles keeps the 256-wide tile.
    const long m_tiles = static_cast<long>(p.B) * p.T * ((p.H + CTH - 1) / CTH) * ((p.W + CTW - 1) / CTW);
    const int nsm = num_sms();
    long best = -1;
    int best_bn = bn;
    for (int cand = bn; cand >= 64; cand >>= 1) {
      if (p.Cout_pad % cand != 0) continue;
      const long tiles = m_tiles * (p.Cout_pad / cand);
      const long cost = ((tiles + nsm - 1) / nsm) * (cand == 256 ? 128 : cand == 128 ? 86 : 56);
      if (best < 0 || cost < best) { best = cost; best_bn = cand; }
    }
    bn = best_bn;
  }
  if (p.mode == CONV_EPI_D2S) {
    const int Cf = p.Cout / (p.ft * p.fh * p.fw);
    LTX2_REQUIRE(Cf % 32 == 0, "conv3d: depth-to-space needs C_out/stride_product %% 32 == 0 (got %d)", Cf);
  }
  // SM-pair kernel when one weight tile covers C_out (LTX2_CONV_PAIR: 0 = never, 1 = 128-channel convs only,
  // 2 = also 256-channel convs (default)).  Measured on B200 (tools/vae_ab.py, profiles/r2_vae_ab.txt): with the fused
  // epilogue 1603 -> 1714 (128) -> 1746 frames/s (128 + 256) for the 65-frame decode.
  {
    const char* env = getenv("LTX2_CONV_PAIR");
    const int lvl = env ? atoi(env) : 2;
    const bool pair = p.Cout_pad == bn && ((bn == 128 && lvl >= 1) || (bn == 256 && lvl >= 2));
    if (pair) {
      CUtensorMap twp;
      LTX2_PROPAGATE(get_tensor_map_2d(&twp, w_packed, p.Cout_pad, static_cast<uint64_t>(27) * p.Cin,
                                       static_cast<uint64_t>(27) * p.Cin, bn / 2));
      // LTX2_CONV_KH3=0: one tap per pipeline stage (the round-2a kernel) for A/B runs
      const char* e3 = getenv("LTX2_CONV_KH3");
      if (e3 && e3[0] == '0')
        return bn == 128 ? launch_conv_pair<128, false>(tx, twp, p, stream) : launch_conv_pair<256, false>(tx, twp, p, stream);
      CUtensorMap tx18;     // the same padded input with an 18-row box: kh = 0..2 of a tile in one load
      {
        uint64_t dims[4] = {static_cast<uint64_t>(p.Cin), static_cast<uint64_t>(p.W + 2), static_cast<uint64_t>(p.H + 2),
                            static_cast<uint64_t>(p.B) * (p.T + 2)};
        uint64_t str[3] = {static_cast<uint64_t>(p.Cin) * 2, static_cast<uint64_t>(p.W + 2) * p.Cin * 2,
                           static_cast<uint64_t>(p.H + 2) * (p.W + 2) * p.Cin * 2};
        uint32_t box[4] = {CBK, CTW, CTH + 2, 1};
        LTX2_PROPAGATE(make_tensor_map_bf16(&tx18, x_padded, 4, dims, str, box));
      }
      return bn == 128 ? launch_conv_pair<128, true>(tx18, twp, p, stream) : launch_conv_pair<256, true>(tx18, twp, p, stream);
    }
  }
  CUtensorMap tw_map;
  const CUtensorMap* tw = &tw_map;
  LTX2_PROPAGATE(get_tensor_map_2d(&tw_map, w_packed, p.Cout_pad, static_cast<uint64_t>(27) * p.Cin,
                                   static_cast<uint64_t>(27) * p.Cin, bn));
  switch (bn) {
    case 256: return launch_conv<256>(tx, tw, p, stream);
    case 128: return launch_conv<128>(tx, tw, p, stream);
    case 64: return launch_conv<64>(tx, tw, p, stream);
    default: return launch_conv<32>(tx, tw, p, stream);
  }
}

}  // namespace ltx2
