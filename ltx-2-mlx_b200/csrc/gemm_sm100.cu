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

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 int M, int N, int K, int splits, GemmEpilogue ep) {
  using Cfg = GemmCfg<BN>;
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
  const int num_kb = (K + BK - 1) / BK;
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
          tma_load_2d(smem_a + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], kb * BK, m0);
          tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], kb * BK, n0);
        }
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp loops, one elected lane issues) =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
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
            // +32 B per K=16 step inside the 128B swizzle atom (descriptor address is in 16 B units)
            umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, ((kb - kb0) | k) != 0);
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
          if (ep.bias != nullptr && first_slice) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + j));
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          }
          if (ep.mode == GEMM_EPI_BF16 || ep.mode == GEMM_EPI_BF16_GELU) {
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

template <int BN>
int launch_gemm(const CUtensorMap* ta, const CUtensorMap* tb, int M, int N, int K, int splits, const GemmEpilogue& ep,
                cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;
  if (!configured) {
    LTX2_CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes));
    configured = true;
  }
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN) * splits;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_bf16_kernel<BN><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(*ta, *tb, M, N, K, splits, ep);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

}  // namespace

int gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, const GemmEpilogue& ep,
              cudaStream_t stream) {
  LTX2_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  LTX2_REQUIRE(N % 32 == 0, "gemm: N=%d must be a multiple of 32", N);
  LTX2_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm: K/lda/ldw must be multiples of 8 (16 B rows)");
  LTX2_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
               "gemm: operands must be 16-byte aligned");
  LTX2_REQUIRE(ep.out != nullptr || ep.n_out_peers > 0, "gemm: null output");
  // BN=256 fills the tensor pipe best; fall back to narrower tiles when N is small or when 128x256 tiles would
  // leave most SMs idle.  The residual epilogue accumulates with reductions, so there the K loop is split instead.
  int bn = 256, splits = 1;
  const int num_kb = (K + BK - 1) / BK;
  const int tiles256 = ((M + BM - 1) / BM) * ((N + 255) / 256);
  if (N % 256 == 0 && ep.mode == GEMM_EPI_F32_RESIDUAL && ep.max_splits > 1 && tiles256 < num_sms()) {
    splits = num_sms() / tiles256;
    if (splits > ep.max_splits) splits = ep.max_splits;
    if (splits > 8) splits = 8;
    while (splits > 1 && num_kb / splits < 8) --splits;     // keep >= 8 K blocks (512 of K) per slice
    if (splits < 1) splits = 1;
  } else {
    if (N % 256 != 0 || tiles256 < num_sms() / 2) bn = 128;
    if (N % 128 != 0 || (bn == 128 && ((M + BM - 1) / BM) * ((N + 127) / 128) < num_sms() / 2)) bn = 64;
    if (N % 64 != 0) bn = 32;
  }
  const CUtensorMap *ta, *tb;
  LTX2_PROPAGATE(get_tensor_map_2d(&ta, A, M, K, lda, BM));
  LTX2_PROPAGATE(get_tensor_map_2d(&tb, W, N, K, ldw, bn));
  switch (bn) {
    case 256: return launch_gemm<256>(ta, tb, M, N, K, splits, ep, stream);
    case 128: return launch_gemm<128>(ta, tb, M, N, K, splits, ep, stream);
    case 64: return launch_gemm<64>(ta, tb, M, N, K, splits, ep, stream);
    default: return launch_gemm<32>(ta, tb, M, N, K, splits, ep, stream);
  }
}

}  // namespace ltx2
