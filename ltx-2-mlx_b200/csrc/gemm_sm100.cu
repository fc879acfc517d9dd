// tcgen05 GEMM for the DiT linears:  C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)
//
// Reference ops replaced: every nn.Linear on the hot path (attention.py:190-201,
// feed_forward.py:23,49, model.py:49-50,529,554) together with the elementwise op that
// follows it (bias, GELU-tanh feed_forward.py:26, gated residual transformer.py:35-46).
//
// Design (sm_100a): persistent grid of one CTA per SM, 6 warps:
//   warp 0  TMA producer   A/W tiles (128 x 64 and BN x 64 bf16, 128B swizzle) -> 4-stage smem ring
//   warp 1  MMA issuer     tcgen05.mma cta_group::1 kind::f16, M=128, N=BN, fp32 accumulators in TMEM,
//                          two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 2-5 epilogue     tcgen05.ld (32 lanes x 32 columns per warp) -> bias/GELU/gate/residual -> global
// Both operands are K-major (row-major activations, nn.Linear [out,in] weights), so no transposes exist anywhere.
#include <stdlib.h>

#include <algorithm>

#include <cuda_fp8.h>

#include "common.cuh"
#include "kernels.h"

namespace ltx2 {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kGemmThreads = 192;

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN >= 256) ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// Epilogue of one accumulator tile for the calling thread's row: TMEM (lane = row, BN fp32 columns at t_row) ->
// bias / GELU / gate / residual -> global.  Shared by the 1-CTA and the 2-CTA kernels.
template <int BN>
__device__ __forceinline__ void epilogue_row(const GemmEpilogue& ep, uint32_t t_row, int row, bool row_ok, int n0, int N,
                                             bool first_slice) {
  int cls = 0;
  if (ep.row_cls != nullptr && row_ok) cls = ep.row_cls[row];
#pragma unroll 1
  for (int c = 0; c < BN; c += 32) {
    uint32_t r[32];
    tmem_ld_32x32(t_row + c, r);
    tmem_ld_wait();
    const int col0 = n0 + c;
    if (row_ok && col0 < N) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (ep.row_scale != nullptr) {            // FP8 operands: de-scale the accumulator (fp8ops.cu)
        float rsc = __ldg(ep.row_scale + row);
        if (ep.row_coef != nullptr) rsc = fmaf(rsc, __ldg(ep.row_coef), __ldg(ep.row_coef + 1));
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 cs = __ldg(reinterpret_cast<const float4*>(ep.col_scale + col0 + j));
          v[j] *= rsc * cs.x; v[j + 1] *= rsc * cs.y; v[j + 2] *= rsc * cs.z; v[j + 3] *= rsc * cs.w;
        }
      }
      if (ep.bias != nullptr && first_slice) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + j));
          v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
        }
      }
      if (ep.mode == GEMM_EPI_E4M3_GELU) {
        // FFN up-projection whose consumer is an FP8 GEMM: the activation is written as E4M3 against the row's bound
        const float inv = 1.0f / fmaf(__ldg(ep.out_l2 + row), __ldg(ep.out_coef), __ldg(ep.out_coef + 1));
        uint32_t q[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t lo = __nv_cvt_float2_to_fp8x2(
              make_float2(gelu_tanh_f(v[4 * j]) * inv, gelu_tanh_f(v[4 * j + 1]) * inv), __NV_SATFINITE, __NV_E4M3);
          const uint32_t hi = __nv_cvt_float2_to_fp8x2(
              make_float2(gelu_tanh_f(v[4 * j + 2]) * inv, gelu_tanh_f(v[4 * j + 3]) * inv), __NV_SATFINITE, __NV_E4M3);
          q[j] = lo | (hi << 16);
        }
        uint8_t* o = reinterpret_cast<uint8_t*>(ep.out) + static_cast<int64_t>(row) * ep.ldo + col0;
        *reinterpret_cast<uint4*>(o) = make_uint4(q[0], q[1], q[2], q[3]);
        *reinterpret_cast<uint4*>(o + 16) = make_uint4(q[4], q[5], q[6], q[7]);
      } else if (ep.mode == GEMM_EPI_BF16 || ep.mode == GEMM_EPI_BF16_GELU) {
        if (ep.mode == GEMM_EPI_BF16_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_f(v[j]);
        }
        uint4 q[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          q[j].x = pack_bf16x2(v[8 * j], v[8 * j + 1]);
          q[j].y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
          q[j].z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
          q[j].w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
        }
        const int64_t off = static_cast<int64_t>(row) * ep.ldo + col0;
        if (ep.n_out_peers == 0) {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ep.out) + off;
#pragma unroll
          for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(o + 8 * j) = q[j];
        } else {
          // context parallel: the same tile is stored into every rank's buffer (stores to peer memory over NVLink)
          for (int pr = 0; pr < ep.n_out_peers; ++pr) {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ep.out_peers[pr]) + off;
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(o + 8 * j) = q[j];
          }
        }
      } else if (ep.mode == GEMM_EPI_F32) {
        float* o = reinterpret_cast<float*>(ep.out) + static_cast<int64_t>(row) * ep.ldo + col0;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {  // GEMM_EPI_F32_RESIDUAL: out += alpha * gate[cls, col] * (acc + bias)
        // the add is performed by the L2 (red.global.add.v4.f32): no read of the residual on the SM, and K slices
        // of a split-K launch can accumulate into the same rows
        float* o = reinterpret_cast<float*>(ep.out) + static_cast<int64_t>(row) * ep.ldo + col0;
        const float* g = ep.gate ? ep.gate + static_cast<int64_t>(cls) * ep.gate_stride + col0 : nullptr;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 gg = g ? __ldg(reinterpret_cast<const float4*>(g + j)) : make_float4(1.f, 1.f, 1.f, 1.f);
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + j), "f"(ep.alpha * gg.x * v[j]),
                       "f"(ep.alpha * gg.y * v[j + 1]), "f"(ep.alpha * gg.z * v[j + 2]),
                       "f"(ep.alpha * gg.w * v[j + 3])
                       : "memory");
        }
      }
    }
  }
}

// FP8 = true: the same pipeline with E4M3 operands -- a 128-byte smem row holds 128 K elements instead of 64, the MMA is
// kind::f8f6f4 with K = 32 per instruction, so a K block feeds twice the FLOPs for the same bytes and tensor clocks.
template <int BN, bool FP8 = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 int M, int N, int K, int splits, GemmEpilogue ep) {
  using Cfg = GemmCfg<BN>;
  constexpr int BKE = FP8 ? 2 * BK : BK;              // K elements per 128-byte row
  pdl_trigger();                                     // the next kernel may become resident under my tail
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::kStages * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                        // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + Cfg::kStages;        // [kStages]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * Cfg::kStages;     // [2]        MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;               // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  // warp-uniform role index (the shuffle makes the uniformity visible to the compiler, so the producer / MMA loops
  // below run on the uniform datapath and tcgen05.mma is issued without a per-lane election loop)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int num_m = (M + BM - 1) / BM;
  const int num_n = (N + BN - 1) / BN;
  // split-K (residual-accumulate epilogue only): work item = (m tile, n tile, K slice); slices add into the fp32
  // output with vector reductions, so small-M GEMMs (context-parallel ranks) still fill the SMs
  const int num_mn = num_m * num_n;
  const int num_tiles = num_mn * splits;
  const int num_kb = (K + BKE - 1) / BKE;
  const int kb_per = (num_kb + splits - 1) / splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
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
  pdl_wait();                                        // everything above overlapped the previous kernel's tail

  if (warp == 0) {
    // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile % num_m) * BM;
      const int n0 = ((tile / num_m) % num_n) * BN;
      const int kb0 = (tile / num_mn) * kb_per;
      const int kb1 = min(num_kb, kb0 + kb_per);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) {
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(smem_a + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], kb * BKE, m0);
          tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], kb * BKE, n0);
        }
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp loops, one elected lane issues) =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc = FP8 ? umma_idesc_e4m3(BM, BN) : umma_idesc_bf16(BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      const int kb0 = (tile / num_mn) * kb_per;
      const int kb1 = min(num_kb, kb0 + kb_per);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t adesc = umma_desc_k_sw128(smem_u32(smem_a + stage * Cfg::kABytes));
        const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smem_b + stage * Cfg::kBBytes));
        if (leader) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 B per MMA K step (16 bf16 / 32 E4M3 elements) inside the 128B swizzle atom (address in 16 B units)
            if (FP8) umma_f8_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, ((kb - kb0) | k) != 0);
            else umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, ((kb - kb0) | k) != 0);
          }
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
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;               // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile % num_m) * BM;
      const int n0 = ((tile / num_m) % num_n) * BN;
      const bool first_slice = tile < num_mn;          // the bias is added by K slice 0 only
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < M;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
      epilogue_row<BN>(ep, t_row, row, row_ok, n0, N, first_slice);
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
// 2-CTA variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 tile of C^T.  CTA r loads weight
// rows [128 r, 128 r + 128) of the tile and half of the tile's tokens into ITS shared memory; one thread of CTA 0
// issues tcgen05.mma.cta_group::2 (M 256 = weight rows, N 256 = tokens), which reads the "A" half and half of "B" from
// each CTA and accumulates output columns [128 r, ...) of the tile into CTA r's tensor memory.  Per MMA an SM now reads 8 KB of operands instead of 12 KB and
// receives 32 KB instead of 48 KB per K block from TMA -- shared-memory bandwidth is what held the 1-CTA kernel at
// ~82 % tensor-pipe activity (profiles/r1c_kernels.json).
//   warp 0 (both CTAs)  TMA producer for the CTA's halves -> own `full` barriers
//   warp 1, CTA 1       relay: own `full` complete -> remote arrive on CTA 0's `peer_full`
//   warp 1, CTA 0       MMA issuer: waits full + peer_full, commits multicast to `empty` / `acc_full` of both CTAs
//   warps 2-5           epilogue of the CTA's 128 rows; `acc_empty` lives in CTA 0 and counts all 8 epilogue warps
// ---------------------------------------------------------------------------------------------------------------
constexpr int kStages2 = 6;
constexpr int kHalfBytes = 128 * BK * 2;                       // 16 KB: 128 rows x 64 k
constexpr int kStageBytes2 = 2 * kHalfBytes;                   // A half + W half
constexpr int kSmemBytes2 = kStages2 * kStageBytes2 + 1024 + 256;

// Transposed epilogue of the 2-CTA kernel: the accumulator tile is C^T, i.e. TMEM lane = output column (weight row),
// TMEM column = token.  Thread `lane` of a warp owns one output column; the 32 lanes of a warp store 32 consecutive
// output columns of one token per instruction (64 B bf16 / 128 B fp32 runs).
__device__ __forceinline__ int lane_id() { return static_cast<int>(threadIdx.x & 31); }

__device__ __forceinline__ void epilogue_col_t(const GemmEpilogue& ep, uint32_t t_row, int col, int tok0, int ntok, int M,
                                               bool first_slice = true) {
  const float b = (ep.bias != nullptr && first_slice) ? __ldg(ep.bias + col) : 0.f;
#pragma unroll 1
  for (int c = 0; c < ntok; c += 32) {
    uint32_t r[32];
    tmem_ld_32x32(t_row + c, r);
    tmem_ld_wait();
    const int t0 = tok0 + c;
    const int tend = min(M, tok0 + ntok);                 // tokens of THIS tile only (ntok need not be a multiple of 32)
    if (t0 >= tend) break;
    const int nv = min(32, tend - t0);                    // warp-uniform
    if (ep.mode == GEMM_EPI_BF16 || ep.mode == GEMM_EPI_BF16_GELU) {
      // two tokens per store instruction: lanes 2i / 2i+1 swap one value, so an even lane holds columns (c, c+1) of
      // token j and its odd neighbour columns (c-1, c) of token j+1 -- 4-byte stores, half as many instructions
      const bool odd = lane_id() & 1;
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ep.out) + static_cast<int64_t>(t0) * ep.ldo + (col & ~1);
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        float v0 = __uint_as_float(r[j]) + b, v1 = __uint_as_float(r[j + 1]) + b;
        if (ep.mode == GEMM_EPI_BF16_GELU) {
          v0 = gelu_tanh_f(v0);
          v1 = gelu_tanh_f(v1);
        }
        const float got = __shfl_xor_sync(0xffffffffu, odd ? v0 : v1, 1);
        const uint32_t pk = odd ? pack_bf16x2(got, v1) : pack_bf16x2(v0, got);
        const int jj = j + (odd ? 1 : 0);
        if (jj < nv) *reinterpret_cast<uint32_t*>(o + static_cast<int64_t>(jj) * ep.ldo) = pk;
      }
    } else if (ep.mode == GEMM_EPI_F32) {
      float* o = reinterpret_cast<float*>(ep.out) + static_cast<int64_t>(t0) * ep.ldo + col;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nv) o[static_cast<int64_t>(j) * ep.ldo] = __uint_as_float(r[j]) + b;
    } else {  // GEMM_EPI_F32_RESIDUAL: out += alpha * gate[cls(token), col] * (acc + bias), added by the L2
      float* o = reinterpret_cast<float*>(ep.out) + static_cast<int64_t>(t0) * ep.ldo + col;
      // the modulation class belongs to the TOKEN (uniform over the lanes) and the gate to (class, this lane's column):
      // lane j fetches the class of token j once per 32 tokens; a chunk of one class (the common case: classes are
      // long runs of tokens) needs one gate load, otherwise one per token -- two dependent global loads per token in
      // the inner loop made this epilogue cost ~100 clocks per token
      const int my_cls = (ep.row_cls != nullptr && lane_id() < nv) ? __ldg(ep.row_cls + t0 + lane_id()) : 0;
      const int cls0 = __shfl_sync(0xffffffffu, my_cls, 0);
      const bool uniform = __all_sync(0xffffffffu, lane_id() >= nv || my_cls == cls0);
      const float g0 = ep.gate != nullptr ? __ldg(ep.gate + static_cast<int64_t>(cls0) * ep.gate_stride + col) : 1.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float g = g0;
        if (!uniform) {
          const int cls = __shfl_sync(0xffffffffu, my_cls, j);
          if (ep.gate != nullptr) g = __ldg(ep.gate + static_cast<int64_t>(cls) * ep.gate_stride + col);
        }
        if (j < nv) {
          asm volatile("red.global.add.f32 [%0], %1;" ::"l"(o + static_cast<int64_t>(j) * ep.ldo),
                       "f"(ep.alpha * g * (__uint_as_float(r[j]) + b))
                       : "memory");
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Transposed 1-CTA variant for token counts that do not fill 128-row tiles (context-parallel shards: 432 rows per rank
// waste 16 % of every 128-row tile, and 4 row tiles x N/256 columns quantise badly over 148 SMs).  The CTA computes a
// C^T tile: MMA M = 128 WEIGHT rows, MMA N = nt tokens, nt any multiple of 16 up to 256 chosen by the host so that
// (a) no token column is wasted (432 = 3 x 144) and (b) the tile count fills whole waves.  Same warp roles and pipeline
// as gemm_bf16_kernel; the accumulator is read with the transposed epilogue (lane = output column).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kStagesT = 4;
constexpr int kStageBytesT = BM * BK * 2 + 256 * BK * 2;       // 16 KB weights + up to 32 KB tokens
constexpr int kSmemBytesT = kStagesT * kStageBytesT + 1024 + 256;

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_t_bf16_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, int M, int N,
                   int K, int nt, int num_t, int splits, GemmEpilogue ep) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                       // [stages][128 weight rows x 64]
  uint8_t* smem_b = smem + kStagesT * (BM * BK * 2);            // [stages][<= 256 tokens x 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStagesT * kStageBytesT);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStagesT;
  uint64_t* acc_full = bars + 2 * kStagesT;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int num_w = N / BM;
  const int num_wt = num_w * num_t;
  const int num_tiles = num_wt * splits;             // work item = (token tile [fastest], weight tile, K slice)
  const int num_kb = (K + BK - 1) / BK;
  const int kb_per = (num_kb + splits - 1) / splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < kStagesT; ++s) {
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
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t stage_tx = BM * BK * 2 + nt * BK * 2;          // the token box is nt rows (out-of-range rows count too)
  pdl_wait();

  if (warp == 0) {
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int t0 = (tile % num_t) * nt;
      const int w0 = ((tile / num_t) % num_w) * BM;
      const int kb0 = (tile / num_wt) * kb_per;
      const int kb1 = min(num_kb, kb0 + kb_per);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) {
          mbar_expect_tx(&full_bar[stage], stage_tx);
          tma_load_2d(smem_a + stage * (BM * BK * 2), &tmap_w, &full_bar[stage], kb * BK, w0);
          tma_load_2d(smem_b + stage * (256 * BK * 2), &tmap_x, &full_bar[stage], kb * BK, t0);
        }
        if (++stage == kStagesT) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int t0 = (tile % num_t) * nt;
      const int n_mma = min(nt, (M - t0 + 15) & ~15);           // tokens of this tile, rounded up to the MMA N step
      const uint32_t idesc = umma_idesc_bf16(BM, static_cast<uint32_t>(n_mma));
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
      const int kb0 = (tile / num_wt) * kb_per;
      const int kb1 = min(num_kb, kb0 + kb_per);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t adesc = umma_desc_k_sw128(smem_u32(smem_a + stage * (BM * BK * 2)));
        const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smem_b + stage * (256 * BK * 2)));
        if (leader) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, ((kb - kb0) | k) != 0);
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == kStagesT) { stage = 0; phase ^= 1; }
      }
      if (leader) umma_commit(&acc_full[acc]);
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int t0 = (tile % num_t) * nt;
      const int col = ((tile / num_t) % num_w) * BM + quarter * 32 + lane;
      const bool first_slice = tile < num_wt;                   // the bias is added by K slice 0 only
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * 256;
      epilogue_col_t(ep, t_row, col, t0, min(nt, M - t0), M, first_slice);
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
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_gemm_t(const CUtensorMap* tw, const CUtensorMap* tx, int M, int N, int K, int nt, int num_t, int splits,
                  const GemmEpilogue& ep, cudaStream_t stream) {
  static PerDeviceOnce configured;
  if (configured.first()) {
    LTX2_CUDA_CHECK(cudaFuncSetAttribute(gemm_t_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesT));
  }
  const int tiles = (N / BM) * num_t * splits;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  LTX2_CUDA_CHECK(launch_pdl(gemm_t_bf16_kernel, dim3(grid), dim3(kGemmThreads), kSmemBytesT, stream, *tw, *tx, M, N, K, nt,
                             num_t, splits, ep));
  count_launch();
  return LTX2_OK;
}

// Static tile schedule of the 2-CTA kernel.  Token tiles are tw wide (a multiple of 32, <= 256); the last one is lw wide
// (the remaining tokens rounded up to 32).  When lw <= tw / 2 the last tile of every weight slab is a "half" item
// (about half the time): full tiles go round-robin over the clusters (token tile fastest, so a weight slab is reused
// from L2) and the half tiles then go two at a time to the clusters that got one full tile fewer, so that e.g.
// 16 x 13.5 tiles finish in 3 tile times on 74 SM pairs instead of 4.
struct PairSched {
  int nc, c;          // clusters, this cluster
  int tw, lw;         // tile widths
  int n_full_t;       // token tiles scheduled as full items
  int num_t;          // token tiles incl. a half one
  int nfull, nhalf;   // work items
  int my_full;        // full items of this cluster
  int lh0;            // half items this cluster takes as a "light" cluster
  int li;             // its index among the light clusters
  int n_light_halves;
};
__device__ __forceinline__ PairSched make_sched(int M, int N, int tw, int nc, int c) {
  PairSched s;
  s.nc = nc;
  s.c = c;
  s.tw = tw;
  s.num_t = (M + tw - 1) / tw;
  s.lw = min(tw, (M - (s.num_t - 1) * tw + 31) & ~31);
  const bool half = 2 * s.lw <= tw;
  s.n_full_t = half ? s.num_t - 1 : s.num_t;
  const int num_w = N / 256;
  s.nfull = s.n_full_t * num_w;
  s.nhalf = half ? num_w : 0;
  s.my_full = c < s.nfull ? (s.nfull - c + nc - 1) / nc : 0;
  const int rem = s.nfull % nc;
  const int light = rem == 0 ? 0 : nc - rem;
  s.n_light_halves = min(s.nhalf, 2 * light);
  s.li = c - rem;
  s.lh0 = (light > 0 && s.li >= 0) ? max(0, min(2, s.n_light_halves - 2 * s.li)) : 0;
  return s;
}
// i-th work item of this cluster: weight tile wi, token tile ti, nt tokens wide; false when the cluster is done
__device__ __forceinline__ bool sched_tile(const PairSched& s, int i, int& wi, int& ti, int& nt) {
  if (i < s.my_full) {
    const int f = s.c + i * s.nc;
    wi = f / s.n_full_t;
    ti = f % s.n_full_t;
    nt = ti == s.num_t - 1 ? s.lw : s.tw;
    return true;
  }
  const int j = i - s.my_full;
  const int h = j < s.lh0 ? 2 * s.li + j : s.n_light_halves + s.c + (j - s.lh0) * s.nc;
  if (h >= s.nhalf) return false;
  wi = h;
  ti = s.num_t - 1;
  nt = s.lw;
  return true;
}

// The pair computes C^T tiles: MMA M = 256 WEIGHT rows (every DiT width is a multiple of 256), MMA N = 256 tokens, or
// 128 for the last token tile when at most 128 tokens remain -- so a token count that is an odd multiple of 128
// (3456 = 27 x 128) costs nothing, where 256-token-row tiles would waste half a tile per weight slab.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x128,
                  const __grid_constant__ CUtensorMap tmap_x64, int M, int N, int K, int tw, GemmEpilogue ep) {
  // tmap_x128 / tmap_x64: the token operand with a box of tw/2 rows (full tiles) and lw/2 rows (last tile)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                       // [stages][128 weight rows x 64]
  uint8_t* smem_b = smem + kStages2 * kHalfBytes;               // [stages][128 (or 64) tokens x 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages2 * kStageBytes2);
  uint64_t* full_bar = bars;                         // [stages]  own TMA -> MMA (CTA 0) / relay (CTA 1)
  uint64_t* peer_full = full_bar + kStages2;         // [stages]  CTA 1 relay -> CTA 0 MMA (used in CTA 0 only)
  uint64_t* empty_bar = peer_full + kStages2;        // [stages]  MMA (multicast) -> own TMA
  uint64_t* acc_full = empty_bar + kStages2;         // [2]       MMA (multicast) -> own epilogue
  uint64_t* acc_empty = acc_full + 2;                // [2]       all 8 epilogue warps -> MMA (used in CTA 0 only)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const PairSched sched = make_sched(M, N, tw, num_clusters, cluster_id);
  const int num_kb = (K + BK - 1) / BK;
  int wi, ti, nt;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x128);
    tma_prefetch_desc(&tmap_x64);
    for (int s = 0; s < kStages2; ++s) {
      mbar_init(&full_bar[s], 1);
      // CTA 0: own TMA bytes (one expect_tx arrival) AND the relay of CTA 1 -- one barrier, one wait per K block
      mbar_init(&peer_full[s], 2);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 8);
    }
    fence_barrier_init();
  }
  __syncthreads();
  cluster_sync_all();                                 // both CTAs are running and their barriers exist
  if (warp == 1) tmem_alloc2(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============ TMA producer (both CTAs: this CTA's 128 weight rows and its half of the token tile) ============
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; sched_tile(sched, it, wi, ti, nt); ++it) {
      const int w0 = wi * 256 + static_cast<int>(rank) * 128;
      const bool last = ti == sched.num_t - 1;
      const int t0 = ti * tw + static_cast<int>(rank) * (nt / 2);
      const uint32_t bytes = kHalfBytes + (nt / 2) * BK * 2;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) {
          // CTA 0's bytes complete on the barrier the MMA issuer waits on (it also counts CTA 1's relay)
          uint64_t* fb = rank == 0 ? &peer_full[stage] : &full_bar[stage];
          mbar_expect_tx(fb, bytes);
          tma_load_2d(smem_a + stage * kHalfBytes, &tmap_w, fb, kb * BK, w0);
          tma_load_2d(smem_b + stage * kHalfBytes, last ? &tmap_x64 : &tmap_x128, fb, kb * BK, t0);
        }
        if (++stage == kStages2) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && rank == 1) {
    // ===================== relay (CTA 1): my halves have landed -> tell the MMA issuer in CTA 0 =====================
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; sched_tile(sched, it, wi, ti, nt); ++it) {
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        if (leader) mbar_arrive_cluster_relaxed(mapa_u32(&peer_full[stage], 0));
        __syncwarp();
        if (++stage == kStages2) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (CTA 0) =====================
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    bool ready = false;                               // the stage about to be consumed was complete when probed
    for (int it = 0; sched_tile(sched, it, wi, ti, nt); ++it) {
      const uint32_t idesc = umma_idesc_bf16(256, static_cast<uint32_t>(nt));
      mbar_wait_cluster(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
      for (int kb = 0; kb < num_kb; ++kb) {
        // ONE plain (CTA-scope) wait per K block: `peer_full` completes when this CTA's TMA bytes have landed and CTA 1's
        // relay has arrived.  (The operands are read by the tensor core through the async proxy, never by this thread; an
        // acquire at cluster scope compiles to an L1 invalidation per wait.)  The barrier round trip costs ~100 clocks
        // even when the phase is complete, so the NEXT stage is probed before this stage's instructions are issued.
        if (!ready) mbar_wait(&peer_full[stage], phase);
        tc_fence_after();
        const int nstage = stage + 1 == kStages2 ? 0 : stage + 1;
        const uint32_t nphase = stage + 1 == kStages2 ? phase ^ 1 : phase;
        ready = mbar_test_wait(&peer_full[nstage], nphase);
        const uint64_t adesc = umma_desc_k_sw128(smem_u32(smem_a + stage * kHalfBytes));
        const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smem_b + stage * kHalfBytes));
        if (leader) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma2_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma2_commit_both(&empty_bar[stage]);
        }
        __syncwarp();
        stage = nstage;
        phase = nphase;
      }
      if (leader) umma2_commit_both(&acc_full[acc]);
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ============ epilogue (warps 2..5, both CTAs: this CTA's 128 output columns x all tokens of the tile) ============
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int it = 0; sched_tile(sched, it, wi, ti, nt); ++it) {
      const int col = wi * 256 + static_cast<int>(rank) * 128 + quarter * 32 + lane;
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * 256;
      epilogue_col_t(ep, t_row, col, ti * tw, nt, M);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(&acc_empty[acc], 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                 // nobody leaves while the peer may still touch my barriers / smem
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// "Wide" 2-CTA variant for the context-parallel shard shapes (432 / 864 token rows per rank).  There the kernels above
// are bound by the L2 -> SM bandwidth, not by the tensor pipe: with 3 token tiles per weight slab every weight byte
// leaves the L2 three times and every token byte once per weight tile (QKV at 432 rows: 470 MB per launch = 44 us at
// the ~10 TB/s the L2 delivers, against 22 us of MMA time).  Here a cluster computes 256 weight rows x UP TO 512
// TOKENS: two accumulators of tw/2 tokens each fill the 512 tensor-memory columns, so a 432-row shard is ONE token tile,
// weights leave the L2 once, and per K block an SM receives 16 KB of weights + tw/2 token rows (44 KB for tw = 448)
// for 2 tw tensor clocks (49 B/clk).  There is no second accumulator stage (the epilogue of an item is not overlapped
// with the next item's MMAs -- a cluster has one or two items), and the K loop can be split (fp32 residual epilogue:
// the partial sums are added by the L2), which is what fills the machine when N / 256 is small (N = 4096: 16 slabs).
//   warp 0 (both CTAs)  TMA producer: this CTA's 128 weight rows, and its quarter of the tokens for either accumulator
//   warp 1, CTA 1       relay: own `full` complete -> remote arrive on CTA 0's `peer_full`
//   warp 1, CTA 0       MMA issuer: 2 x 4 tcgen05.mma.cta_group::2 per K block (M 256, N tw/2)
//   warps 2-5           epilogue of the CTA's 128 output columns x all tokens of the tile
// ---------------------------------------------------------------------------------------------------------------
constexpr int kStagesW = 4;
constexpr int kStageBytesW = kHalfBytes + 256 * BK * 2;         // 16 KB weights + up to 256 token rows (32 KB)
constexpr int kSmemBytesW = kStagesW * kStageBytesW + 1024 + 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm2w_bf16_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, int M, int N,
                   int K, int tw, int num_t, int splits, GemmEpilogue ep) {
  // tmap_x: the token operand with a box of tw/4 rows (one CTA's share of one accumulator)
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                       // [stages][128 weight rows x 64]
  uint8_t* smem_b = smem + kStagesW * kHalfBytes;               // [stages][2 x tw/4 token rows x 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStagesW * kStageBytesW);
  uint64_t* full_bar = bars;                         // [stages]  own TMA -> MMA (CTA 0) / relay (CTA 1)
  uint64_t* peer_full = full_bar + kStagesW;         // [stages]  CTA 1 relay -> CTA 0 MMA (used in CTA 0 only)
  uint64_t* empty_bar = peer_full + kStagesW;        // [stages]  MMA (multicast) -> own TMA
  uint64_t* acc_full = empty_bar + kStagesW;         // MMA (multicast) -> own epilogue
  uint64_t* acc_empty = acc_full + 1;                // all 8 epilogue warps -> MMA (used in CTA 0 only)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_kb = (K + BK - 1) / BK;
  const int kb_per = (num_kb + splits - 1) / splits;
  const int num_items = (N / 256) * splits * num_t;  // item = (token tile [fastest], K slice, weight slab)
  const int qrows = tw / 4;                          // token rows per CTA and accumulator
  const uint32_t b_acc_bytes = static_cast<uint32_t>(qrows) * BK * 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < kStagesW; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&peer_full[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 8);
    fence_barrier_init();
  }
  __syncthreads();
  cluster_sync_all();                                 // both CTAs are running and their barriers exist
  if (warp == 1) tmem_alloc2(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ============ TMA producer (both CTAs) ============
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t bytes = kHalfBytes + 2 * b_acc_bytes;
    for (int it = cluster_id; it < num_items; it += num_clusters) {
      const int ti = it % num_t, si = (it / num_t) % splits, wi = it / (num_t * splits);
      const int w0 = wi * 256 + static_cast<int>(rank) * 128;
      const int t0 = ti * tw + static_cast<int>(rank) * qrows;   // accumulator j: + j * tw / 2
      const int kb0 = si * kb_per, kb1 = min(num_kb, kb0 + kb_per);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) {
          mbar_expect_tx(&full_bar[stage], bytes);
          tma_load_2d(smem_a + stage * kHalfBytes, &tmap_w, &full_bar[stage], kb * BK, w0);
          uint8_t* sb = smem_b + stage * (256 * BK * 2);
          tma_load_2d(sb, &tmap_x, &full_bar[stage], kb * BK, t0);
          tma_load_2d(sb + b_acc_bytes, &tmap_x, &full_bar[stage], kb * BK, t0 + tw / 2);
        }
        if (++stage == kStagesW) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && rank == 1) {
    // ===================== relay (CTA 1) =====================
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int it = cluster_id; it < num_items; it += num_clusters) {
      const int si = (it / num_t) % splits;
      const int kb0 = si * kb_per, kb1 = min(num_kb, kb0 + kb_per);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        if (leader) mbar_arrive_cluster_relaxed(mapa_u32(&peer_full[stage], 0));
        __syncwarp();
        if (++stage == kStagesW) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (CTA 0) =====================
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0, acc_phase = 0;
    const uint32_t idesc = umma_idesc_bf16(256, static_cast<uint32_t>(tw / 2));
    for (int it = cluster_id; it < num_items; it += num_clusters) {
      const int si = (it / num_t) % splits;
      const int kb0 = si * kb_per, kb1 = min(num_kb, kb0 + kb_per);
      mbar_wait(acc_empty, acc_phase ^ 1);            // the previous item's epilogue has read the accumulators
      tc_fence_after();
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        mbar_wait(&peer_full[stage], phase);
        tc_fence_after();
        const uint64_t adesc = umma_desc_k_sw128(smem_u32(smem_a + stage * kHalfBytes));
        const uint64_t bdesc0 = umma_desc_k_sw128(smem_u32(smem_b + stage * (256 * BK * 2)));
        const uint64_t bdesc1 = umma_desc_k_sw128(smem_u32(smem_b + stage * (256 * BK * 2) + b_acc_bytes));
        if (leader) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma2_bf16_ss(tmem_base, adesc + 2 * k, bdesc0 + 2 * k, idesc, (kb > kb0) || k != 0);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma2_bf16_ss(tmem_base + 256, adesc + 2 * k, bdesc1 + 2 * k, idesc, (kb > kb0) || k != 0);
          umma2_commit_both(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == kStagesW) { stage = 0; phase ^= 1; }
      }
      if (leader) umma2_commit_both(acc_full);
      __syncwarp();
      acc_phase ^= 1;
    }
  } else {
    // ============ epilogue (warps 2..5, both CTAs: this CTA's 128 output columns x all tokens of the tile) ============
    const int quarter = warp & 3;
    uint32_t acc_phase = 0;
    for (int it = cluster_id; it < num_items; it += num_clusters) {
      const int ti = it % num_t, si = (it / num_t) % splits, wi = it / (num_t * splits);
      const int col = wi * 256 + static_cast<int>(rank) * 128 + quarter * 32 + lane;
      mbar_wait(acc_full, acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
      const int tok0 = ti * tw;
#pragma unroll 1
      for (int j = 0; j < 2; ++j) {
        const int tj = tok0 + j * (tw / 2);
        if (tj < M) epilogue_col_t(ep, t_row + j * 256, col, tj, min(tw / 2, M - tj), M, si == 0);
      }
      tc_fence_before();
      __syncwarp();
      // the accumulator reads are complete (wait::ld inside the epilogue) and fenced: no memory ordering is needed on
      // the arrival itself (a release at cluster scope costs ~1000 clocks)
      if (lane == 0) {
        if (rank == 0) mbar_arrive(acc_empty);
        else mbar_arrive_cluster_relaxed(mapa_u32(acc_empty, 0));
      }
      acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                 // nobody leaves while the peer may still touch my barriers / smem
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

int launch_gemm2w(const CUtensorMap* tw_map, const CUtensorMap* tx, int M, int N, int K, int tile_w, int num_t, int splits,
                  const GemmEpilogue& ep, cudaStream_t stream) {
  static PerDeviceOnce configured;
  if (configured.first()) {
    LTX2_CUDA_CHECK(cudaFuncSetAttribute(gemm2w_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesW));
  }
  const int items = (N / 256) * num_t * splits;
  const int pairs = num_sms() / 2;
  const int grid = 2 * (items < pairs ? items : pairs);
  LTX2_CUDA_CHECK(launch_pdl(gemm2w_bf16_kernel, dim3(grid), dim3(kGemmThreads), kSmemBytesW, stream, *tw_map, *tx, M, N, K,
                             tile_w, num_t, splits, ep));
  count_launch();
  return LTX2_OK;
}

int gemm2_tiles(int M, int N, int tile_w = 256) { return ((M + tile_w - 1) / tile_w) * (N / 256); }

int launch_gemm2(const CUtensorMap* tw, const CUtensorMap* tx128, const CUtensorMap* tx64, int M, int N, int K,
                 int tile_w, const GemmEpilogue& ep, cudaStream_t stream) {
  static PerDeviceOnce configured;
  if (configured.first()) {
    LTX2_CUDA_CHECK(cudaFuncSetAttribute(gemm2_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes2));
  }
  const int tiles = gemm2_tiles(M, N, tile_w);
  const int pairs = num_sms() / 2;
  const int grid = 2 * (tiles < pairs ? tiles : pairs);
  gemm2_bf16_kernel<<<grid, kGemmThreads, kSmemBytes2, stream>>>(*tw, *tx128, *tx64, M, N, K, tile_w, ep);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

template <int BN, bool FP8 = false>
int launch_gemm(const CUtensorMap* ta, const CUtensorMap* tb, int M, int N, int K, int splits, const GemmEpilogue& ep,
                cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static PerDeviceOnce configured;
  if (configured.first()) {
    LTX2_CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_kernel<BN, FP8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes));
  }
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN) * splits;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  LTX2_CUDA_CHECK(launch_pdl(gemm_bf16_kernel<BN, FP8>, dim3(grid), dim3(kGemmThreads), Cfg::kSmemBytes, stream, *ta, *tb, M,
                             N, K, splits, ep));
  count_launch();
  return LTX2_OK;
}

}  // namespace

GemmPlan plan_gemm(int M, int N, int K, int mode, int max_splits, int n_out_peers) {
  GemmPlan p;
  // BN=256 fills the tensor pipe best; fall back to narrower tiles when N is small or when 128x256 tiles would
  // leave most SMs idle.  The residual epilogue accumulates with reductions, so there the K loop is split instead.
  int bn = 256, splits = 1;
  const int num_kb = (K + BK - 1) / BK;
  const int tiles256 = ((M + BM - 1) / BM) * ((N + 255) / 256);
  if (N % 256 == 0 && mode == GEMM_EPI_F32_RESIDUAL && max_splits > 1 && tiles256 < num_sms()) {
    splits = num_sms() / tiles256;
    if (splits > max_splits) splits = max_splits;
    if (splits > 8) splits = 8;
    while (splits > 1 && num_kb / splits < 8) --splits;     // keep >= 8 K blocks (512 of K) per slice
    if (splits < 1) splits = 1;
  } else {
    if (N % 256 != 0 || tiles256 < num_sms() / 2) bn = 128;
    if (N % 128 != 0 || (bn == 128 && ((M + BM - 1) / BM) * ((N + 127) / 128) < num_sms() / 2)) bn = 64;
    if (N % 64 != 0) bn = 32;
  }
  // Kernel choice by a small cycle model.  One work item costs max(MMA cycles, operand bytes / ~77 B per clock -- what
  // the L2 feeds one SM; the standard 128 x 256 tile is bound by it: 48 KB against 512 MMA clocks per K block, the 82 %
  // tensor-pipe activity of profiles/r1c_kernels.json) plus a fixed cost, and a launch costs that times the number
  // of waves.  Candidates:
  //   standard     128 token rows x bn columns (above)
  //   transposed   128 weight rows x nt tokens, nt a multiple of 16 (no wasted token rows, flexible tile count)
  //   pair         SM pairs (cta_group::2): 256 weight rows x tw tokens, tw a multiple of 32; each SM loads only half
  //                of the token operand, so the small tiles a 432-row shard needs stay fed
  // Measured (tools/gemm_cp_shapes.py, tools/gemm_vs_cublas.py): the alternatives win on the wide bf16-output
  // projections (QKV N = 12288, FFN up N = 16384: 4-22 % at 432 / 864 / 1728 / 3456 token rows) and lose on the
  // N = 4096 launches (few weight slabs) and on the residual epilogue (scalar reductions, no split-K in the pair
  // kernel), so unless forced they are only considered for bf16 outputs with N >= 8192, and the standard kernel keeps
  // a 10 % bonus.  At the full 3456 rows they gain 3-4 % stand-alone but nothing inside the power-capped step (11.20 vs
  // 11.23 steps/s, same box), so up to 2048 rows only: the single-GPU path stays on the standard kernel.
  //   LTX2_GEMM_T=0 / LTX2_GEMM_2CTA=0 remove a candidate, =2 force it when it applies (tests)
  {
    const char* envt = getenv("LTX2_GEMM_T");
    const char* env2 = getenv("LTX2_GEMM_2CTA");
    const int nsm = num_sms();
    auto item_cost = [](long kb, long mma_per_kb, long bytes_per_kb, long fixed) {
      const long feed = bytes_per_kb / 77;
      return kb * (mma_per_kb > feed ? mma_per_kb : feed) + fixed;
    };
    const long kb_std = (num_kb + splits - 1) / splits;
    const long items_std = static_cast<long>((M + BM - 1) / BM) * ((N + bn - 1) / bn) * splits;
    const long cost_std = ((items_std + nsm - 1) / nsm) * item_cost(kb_std, 4 * (bn / 2), (BM + bn) * BK * 2, 1500);
    const bool may_split = mode == GEMM_EPI_F32_RESIDUAL && max_splits > 1;
    // transposed 1-CTA
    long best_t = -1;
    int t_nt = 0, t_num = 0, t_splits = 1;
    if (N % BM == 0 && n_out_peers == 0 && !(envt && envt[0] == '0')) {
      for (int num_t = (M + 255) / 256; num_t <= (M + 63) / 64 && num_t <= 64; ++num_t) {
        const int nt = (((M + num_t - 1) / num_t) + 15) & ~15;
        if (nt > 256 || static_cast<long>(nt) * (num_t - 1) >= M) continue;
        const long base = static_cast<long>(N / BM) * num_t;
        for (int sp = 1; sp <= (may_split ? 8 : 1); ++sp) {
          if (sp > 1 && (sp > max_splits || num_kb / sp < 8)) break;
          const long cost = ((base * sp + nsm - 1) / nsm) *
                            item_cost((num_kb + sp - 1) / sp, 4 * (nt / 2), (BM + nt) * BK * 2, 2500);
          if (best_t < 0 || cost < best_t) {
            best_t = cost;
            t_nt = nt;
            t_num = num_t;
            t_splits = sp;
          }
        }
      }
    }
    // pair (2-CTA) kernel: no split-K
    long best_p = -1;
    int p_tw = 0;
    if (N % 256 == 0 && n_out_peers == 0 && splits == 1 && M >= 64 && !(env2 && env2[0] == '0')) {
      const int pairs = nsm / 2;
      for (int num_t = (M + 255) / 256; num_t <= (M + 63) / 64 && num_t <= 64; ++num_t) {
        const int tw = (((M + num_t - 1) / num_t) + 31) & ~31;
        if (tw > 256 || static_cast<long>(tw) * (num_t - 1) >= M) continue;
        const int lw = std::min(tw, (M - (num_t - 1) * tw + 31) & ~31);
        const long num_w = N / 256;
        // items in units of full tiles (a last tile of at most half width counts as half)
        const long units2 = 2 * num_w * (num_t - 1) + (2 * lw <= tw ? num_w : 2 * num_w);
        const long waves2 = (units2 + 2 * pairs - 1) / (2 * pairs);   // in half-tile steps
        const long cost = waves2 * item_cost(num_kb, 4 * (tw / 2), (BM + tw / 2) * BK * 2, 3000) / 2 +
                          item_cost(0, 0, 0, 1500);
        if (best_p < 0 || cost < best_p) {
          best_p = cost;
          p_tw = tw;
        }
      }
    }
    // wide pair kernel (256 weight rows x up to 512 tokens, K split for the residual epilogue).  Its K loop runs at the
    // tensor rate (930 clocks per K block of 2 x 4 N = 224 instructions, tools/gemm_epi_probe.py), but a 432-row shard
    // gives it only N / 256 items (16-64 for 74 SM pairs) and its epilogue is not overlapped, so measured
    // (tools/gemm_cp_shapes.py, profiles/r2_gemm_cp_shapes.txt) it loses to the kernels above on every DiT shape
    // (QKV at 432 rows 55.6 vs 43.6 us, FFN down 68.5 vs 62.9).  It is therefore only taken when forced:
    // LTX2_GEMM_WIDE=2 (tests, A/B runs).
    {
      const char* envw = getenv("LTX2_GEMM_WIDE");
      const bool off = (envw && envw[0] == '0') || (env2 && (env2[0] == '0' || env2[0] == '2')) || (envt && envt[0] == '2');
      const bool force_w = envw && envw[0] == '2';
      const bool modes_ok = mode == GEMM_EPI_BF16 || mode == GEMM_EPI_BF16_GELU || mode == GEMM_EPI_F32 ||
                            mode == GEMM_EPI_F32_RESIDUAL;
      if (!off && force_w && modes_ok && N % 256 == 0 && n_out_peers == 0 && M >= 64) {
        const int pairs = nsm / 2;
        const int num_t = (M + 511) / 512;
        const int tw = (((M + num_t - 1) / num_t) + 31) & ~31;
        const int slabs = N / 256;
        int best_s = 1;
        long best_c = -1;
        for (int sp = 1; sp <= (may_split ? 8 : 1); ++sp) {
          if (sp > 1 && (sp > max_splits || num_kb / sp < 8)) break;
          const long items = static_cast<long>(slabs) * num_t * sp;
          const long waves = (items + pairs - 1) / pairs;
          // per K block 2 tw tensor clocks; per item the epilogue (not overlapped) and the pipeline fill
          const long c = waves * (static_cast<long>((num_kb + sp - 1) / sp) * 2 * tw + 6 * tw + 2000);
          if (best_c < 0 || c < best_c) {
            best_c = c;
            best_s = sp;
          }
        }
        // fewer than half of the SM pairs busy: the standard kernel's smaller tiles spread better
        const long items = static_cast<long>(slabs) * num_t * best_s;
        if (force_w || items * 2 >= pairs) {
          p.kernel = GEMM_KERNEL_WIDE;
          p.tile_w = tw;
          p.last_w = tw;
          p.num_t = num_t;
          p.splits = best_s;
          return p;
        }
      }
    }
    const bool force_t = envt && envt[0] == '2', force_p = env2 && env2[0] == '2';
    const bool wide_bf16 = (mode == GEMM_EPI_BF16 || mode == GEMM_EPI_BF16_GELU) && N >= 8192 && M <= 2048;
    const bool take_p = best_p > 0 && (force_p || (!force_t && wide_bf16 && best_p * 10 < cost_std * 9 &&
                                                   (best_t < 0 || best_p <= best_t)));
    const bool take_t = !take_p && best_t > 0 && (force_t || (wide_bf16 && best_t * 10 < cost_std * 9));
    if (take_p) {
      const int num_t = (M + p_tw - 1) / p_tw;
      p.kernel = GEMM_KERNEL_PAIR;
      p.tile_w = p_tw;
      p.last_w = std::min(p_tw, (M - (num_t - 1) * p_tw + 31) & ~31);
      p.num_t = num_t;
      p.splits = 1;
      return p;
    }
    if (take_t) {
      p.kernel = GEMM_KERNEL_TRANSPOSED;
      p.tile_w = t_nt;
      p.last_w = std::min(t_nt, M - (t_num - 1) * t_nt);
      p.num_t = t_num;
      p.splits = t_splits;
      return p;
    }
  }
  p.kernel = GEMM_KERNEL_STANDARD;
  p.bn = bn;
  p.splits = splits;
  return p;
}

int gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, const GemmEpilogue& ep,
              cudaStream_t stream) {
  LTX2_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  LTX2_REQUIRE(N % 32 == 0, "gemm: N=%d must be a multiple of 32", N);
  LTX2_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm: K/lda/ldw must be multiples of 8 (16 B rows)");
  LTX2_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
               "gemm: operands must be 16-byte aligned");
  LTX2_REQUIRE(ep.out != nullptr || ep.n_out_peers > 0, "gemm: null output");
  // the C^T kernels store bf16 outputs as column pairs: 4-byte aligned rows (n_out_peers = 1 leaves the standard kernel)
  const bool pair_stores_ok = (ep.ldo % 2 == 0) && (reinterpret_cast<uintptr_t>(ep.out) % 4 == 0);
  const GemmPlan plan = plan_gemm(M, N, K, ep.mode, ep.max_splits, pair_stores_ok ? ep.n_out_peers : 1);
  if (plan.kernel == GEMM_KERNEL_WIDE) {
    CUtensorMap tw, tx;
    LTX2_PROPAGATE(get_tensor_map_2d(&tw, W, N, K, ldw, 128));
    LTX2_PROPAGATE(get_tensor_map_2d(&tx, A, M, K, lda, plan.tile_w / 4));
    return launch_gemm2w(&tw, &tx, M, N, K, plan.tile_w, plan.num_t, plan.splits, ep, stream);
  }
  if (plan.kernel == GEMM_KERNEL_PAIR) {
    CUtensorMap tw, txf, txl;
    LTX2_PROPAGATE(get_tensor_map_2d(&tw, W, N, K, ldw, 128));
    LTX2_PROPAGATE(get_tensor_map_2d(&txf, A, M, K, lda, plan.tile_w / 2));
    LTX2_PROPAGATE(get_tensor_map_2d(&txl, A, M, K, lda, plan.last_w / 2));
    return launch_gemm2(&tw, &txf, &txl, M, N, K, plan.tile_w, ep, stream);
  }
  if (plan.kernel == GEMM_KERNEL_TRANSPOSED) {
    CUtensorMap tw, tx;
    LTX2_PROPAGATE(get_tensor_map_2d(&tw, W, N, K, ldw, BM));
    LTX2_PROPAGATE(get_tensor_map_2d(&tx, A, M, K, lda, plan.tile_w));
    return launch_gemm_t(&tw, &tx, M, N, K, plan.tile_w, plan.num_t, plan.splits, ep, stream);
  }
  const int bn = plan.bn, splits = plan.splits;
  CUtensorMap ta, tb;
  LTX2_PROPAGATE(get_tensor_map_2d(&ta, A, M, K, lda, BM));
  LTX2_PROPAGATE(get_tensor_map_2d(&tb, W, N, K, ldw, bn));
  switch (bn) {
    case 256: return launch_gemm<256>(&ta, &tb, M, N, K, splits, ep, stream);
    case 128: return launch_gemm<128>(&ta, &tb, M, N, K, splits, ep, stream);
    case 64: return launch_gemm<64>(&ta, &tb, M, N, K, splits, ep, stream);
    default: return launch_gemm<32>(&ta, &tb, M, N, K, splits, ep, stream);
  }
}

int gemm_e4m3(const void* A8, int64_t lda, const void* W8, int64_t ldw, int M, int N, int K, const GemmEpilogue& ep,
              cudaStream_t stream) {
  LTX2_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_e4m3: empty problem M=%d N=%d K=%d", M, N, K);
  LTX2_REQUIRE(N % 32 == 0, "gemm_e4m3: N=%d must be a multiple of 32", N);
  LTX2_REQUIRE(K % 16 == 0 && lda % 16 == 0 && ldw % 16 == 0, "gemm_e4m3: K/lda/ldw must be multiples of 16 (16 B rows)");
  LTX2_REQUIRE((reinterpret_cast<uintptr_t>(A8) & 15) == 0 && (reinterpret_cast<uintptr_t>(W8) & 15) == 0,
               "gemm_e4m3: operands must be 16-byte aligned");
  LTX2_REQUIRE(ep.out != nullptr && ep.row_scale != nullptr && ep.col_scale != nullptr,
               "gemm_e4m3: output, row_scale and col_scale are required");
  // the standard 128-token-row kernel only (n_out_peers = 1 removes the shard-shaped bf16 candidates from the plan)
  const GemmPlan plan = plan_gemm(M, N, (K + 1) / 2, ep.mode, ep.max_splits, 1);
  CUtensorMap ta, tb;
  LTX2_PROPAGATE(get_tensor_map_2d(&ta, A8, M, K, lda, BM, 1));
  LTX2_PROPAGATE(get_tensor_map_2d(&tb, W8, N, K, ldw, plan.bn, 1));
  switch (plan.bn) {
    case 256: return launch_gemm<256, true>(&ta, &tb, M, N, K, plan.splits, ep, stream);
    case 128: return launch_gemm<128, true>(&ta, &tb, M, N, K, plan.splits, ep, stream);
    case 64: return launch_gemm<64, true>(&ta, &tb, M, N, K, plan.splits, ep, stream);
    default: return launch_gemm<32, true>(&ta, &tb, M, N, K, plan.splits, ep, stream);
  }
}

}  // namespace ltx2
