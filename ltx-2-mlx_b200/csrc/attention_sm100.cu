// Flash-style attention on tcgen05 tensor cores:  O = softmax(Q K^T / sqrt(d)) V, non-causal, no mask.
//
// Reference op replaced: mx.fast.scaled_dot_product_attention as called from
// _compiled_attention_core_no_mask (attention.py:12-34), plus the head merge
// (B,H,T,D)->(B,T,H*D) (:34) and the V2 per-head gate 2*sigmoid(logits) (:243-250).
//
// One CTA per (128-query tile, batch*head).  6 warps:
//   warp 0   TMA producer: K_j / V^T_j tiles (128 keys) through 3-stage smem rings
//   warp 1   MMA issuer:   S_j = Q K_j^T  (fp32 in TMEM, double-buffered) and O += P_j V_j (one TMEM accumulator).
//                          BOTH A operands (Q and P) are read from tensor memory (tcgen05.mma with a TMEM A operand),
//                          so shared memory only carries K and V: an SS-mode M=128,N=128 MMA would need the full
//                          128 B/clk of shared-memory bandwidth and starve next to the TMA writes.
//   warps 2-9 softmax:     TWO threads per query row (warps w and w+4 share a TMEM lane quarter and split the 128
//                          key columns), so every SM sub-partition has two softmax warps to hide latencies; the
//                          pair exchanges its block maximum through shared memory (named barrier of 64 threads).
//                          Q's row is loaded from global and parked in TMEM once.
//                          S_j is read from TMEM ONCE into registers; p = exp2(s*c - m*c) is written back to TMEM
//                          as packed bf16 (64 columns).
//                          The running output stays in TMEM; it is rescaled only when the row maximum grew by more
//                          than 2^8 since the scale in use (lazy rescale: the stale maximum only changes the common
//                          factor of P and l, which cancels in O / l).
// K is consumed [keys, d] (K-major for S), V is consumed TRANSPOSED [d, keys] (K-major for P V), so
// both MMAs use the same K-major/128B-swizzle descriptor form as the GEMM.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace ltx2 {

namespace {

constexpr int kAttnThreads = 320;   // TMA warp, MMA warp, 8 softmax warps
constexpr int BQ = 128;    // queries per CTA
constexpr int BKV = 128;   // keys per block

constexpr int kKVStages = 3;

template <int DH>
struct AttnCfg {
  static constexpr int kKBytes = BKV * DH * 2;
  static constexpr int kVBytes = DH * BKV * 2;
  static constexpr int kSmemBytes = kKVStages * (kKBytes + kVBytes) + 1024 + 256 + 3072;
  static constexpr int kTmemCols = 512;
  static constexpr int kOCol = 256;            // O accumulator (DH fp32 columns) after the two S buffers
  static constexpr int kPCol = 384;            // P: 128 x 128 bf16 = 64 columns
  static constexpr int kQCol = 448;            // Q: 128 x DH bf16 = DH/2 columns
};

template <int DH, bool VROWS>
__global__ void __launch_bounds__(kAttnThreads, 1)
attention_kernel(const __nv_bfloat16* __restrict__ q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v, __nv_bfloat16* __restrict__ out, int H, int Tq, int Tk,
                 float scale_log2, float scale, const float* __restrict__ gate_logits, float* __restrict__ lse_out,
                 long long* __restrict__ trace, AttnOutScatter sc) {
  using Cfg = AttnCfg<DH>;
  // optional timeline trace of CTA (0,0): trace[j*8 + k] = clock64 at event k of block j (diagnostics only)
  const bool tr = trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + kKVStages * Cfg::kKBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kKVStages * Cfg::kVBytes);
  uint64_t* q_ready = bars;                        // 1   softmax -> MMA: Q parked in TMEM
  uint64_t* k_full = bars + 1;                     // kKVStages
  uint64_t* k_empty = k_full + kKVStages;
  uint64_t* v_full = k_empty + kKVStages;
  uint64_t* v_empty = v_full + kKVStages;
  uint64_t* s_full = v_empty + kKVStages;          // 2   MMA -> softmax: S_j in TMEM
  uint64_t* s_empty = s_full + 2;                  // 2   softmax -> MMA: S buffer read into registers
  uint64_t* p_full = s_empty + 2;                  // 1   softmax -> MMA: P_j in TMEM (and O rescaled if needed)
  uint64_t* pv_done = p_full + 1;                  // 1   MMA -> softmax: P_j V_j retired (P reusable, O up to date)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);
  float* xmax = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [2 parity][2 half][128 rows]
  float* xsum = xmax + 512;                                                          // [2 half][128 rows]

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform role (see gemm_sm100.cu)
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ;
  const int bh = blockIdx.y;
  const int nkv = (Tk + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_ready, 256);
    for (int i = 0; i < kKVStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 256);
    }
    mbar_init(p_full, 256);
    mbar_init(pv_done, 1);
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
    int st = 0;
    uint32_t ph = 0;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&k_empty[st], ph ^ 1);
      if (leader) {
        mbar_expect_tx(&k_full[st], Cfg::kKBytes);
#pragma unroll
        for (int i = 0; i < DH / 64; ++i)
          tma_load_3d(sK + st * Cfg::kKBytes + i * (BKV * 128), &tmap_k, &k_full[st], i * 64, j * BKV, bh);
      }
      mbar_wait(&v_empty[st], ph ^ 1);
      if (leader) {
        mbar_expect_tx(&v_full[st], Cfg::kVBytes);
        if (VROWS) {   // V rows [keys, d]: one 128-key x 64-channel box per 64 channels (MN-major B operand)
#pragma unroll
          for (int i = 0; i < DH / 64; ++i)
            tma_load_4d(sV + st * Cfg::kVBytes + i * (BKV * 128), &tmap_v, &v_full[st], i * 64, j * BKV, bh % H, bh / H);
        } else {       // V^T [d, keys]: one DH x 64-key box per 64 keys (K-major B operand)
#pragma unroll
          for (int i = 0; i < BKV / 64; ++i)
            tma_load_3d(sV + st * Cfg::kVBytes + i * (DH * 128), &tmap_v, &v_full[st], j * BKV + i * 64, 0, bh);
        }
      }
      if (++st == kKVStages) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp loops, one elected lane issues) =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV);
    constexpr uint32_t idesc_o = umma_idesc_bf16(BQ, DH, VROWS);
    int ks_st = 0, vs_st = 0;                 // K / V ring positions
    uint32_t ks_ph = 0, vs_ph = 0;
    auto issue_s = [&](int j) {
      const int b = j & 1;
      mbar_wait(&k_full[ks_st], ks_ph);
      mbar_wait(&s_empty[b], ((j >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint64_t kd = umma_desc_k_sw128(smem_u32(sK + ks_st * Cfg::kKBytes));
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks)
          umma_bf16_ts(tmem_base + b * BKV, tmem_base + Cfg::kQCol + ks * 8,
                       kd + ((ks / 4) * (BKV * 128) >> 4) + 2 * (ks % 4), idesc_s, ks != 0);
        umma_commit(&s_full[b]);
        umma_commit(&k_empty[ks_st]);
        if (tr) trace[j * 8 + 0] = clock64();
      }
      __syncwarp();
      if (++ks_st == kKVStages) { ks_st = 0; ks_ph ^= 1; }
    };
    mbar_wait(q_ready, 0);
    tc_fence_after();
    // Q*K^T runs TWO blocks ahead of the softmax: S_{j+2} is issued as soon as the softmax has pulled S_j out of the
    // shared TMEM buffer (early in its iteration), so it executes during softmax_j instead of queueing behind P_j*V_j
    issue_s(0);
    if (nkv > 1) issue_s(1);
    for (int j = 0; j < nkv; ++j) {
      if (j + 2 < nkv) issue_s(j + 2);
      mbar_wait(p_full, j & 1);
      if (tr && leader) trace[j * 8 + 1] = clock64();
      mbar_wait(&v_full[vs_st], vs_ph);
      tc_fence_after();
      const uint32_t v_addr = smem_u32(sV + vs_st * Cfg::kVBytes);
      // K-major V^T: 64-key boxes of DH rows; MN-major V: 16 keys = 2048 B per K step, 64-channel boxes 16 KB apart
      const uint64_t vd = VROWS ? umma_desc_mn_sw128(v_addr, BKV * 128, 1024) : umma_desc_k_sw128(v_addr);
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < BKV / 16; ++ks)
          umma_bf16_ts(tmem_base + Cfg::kOCol, tmem_base + Cfg::kPCol + ks * 8,
                       vd + (VROWS ? ks * (2048 >> 4) : ((ks / 4) * (DH * 128) >> 4) + 2 * (ks % 4)), idesc_o,
                       (j | ks) != 0);
        umma_commit(&v_empty[vs_st]);
        umma_commit(pv_done);
        if (tr) trace[j * 8 + 2] = clock64();
      }
      __syncwarp();
      if (++vs_st == kKVStages) { vs_st = 0; vs_ph ^= 1; }
    }
  } else {
    // ===================== softmax + output (warps 2..9) =====================
    constexpr int HC = BKV / 2;                         // key columns per thread
    constexpr int OC = DH / 2;                          // output columns per thread
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;                  // query row inside the tile
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_o = t_lane + Cfg::kOCol + half * OC;
    const uint32_t bar_id = 1 + quarter;                // named barrier shared by the two warps of a row group
    float m_run = -INFINITY;                            // true running row maximum (raw scores)
    float m_used = 0.f;                                 // maximum the current scale of P, l and O refers to
    float l = 0.f;                                      // this thread's half of the row sum
    if (half == 0) {
      // park this row of Q in tensor memory: column c of the Q region holds elements (2c, 2c+1)
      const int row = q0 + r;
      const uint4* qrow = reinterpret_cast<const uint4*>(q + (static_cast<int64_t>(bh) * Tq + (row < Tq ? row : 0)) * DH);
#pragma unroll
      for (int c = 0; c < DH / 2; c += 32) {
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint4 u = make_uint4(0u, 0u, 0u, 0u);
          if (row < Tq) u = __ldg(qrow + c / 4 + i);
          v[4 * i] = u.x; v[4 * i + 1] = u.y; v[4 * i + 2] = u.z; v[4 * i + 3] = u.w;
        }
        tmem_st_32x32(t_lane + Cfg::kQCol + c, v);
      }
      tmem_st_wait();
      tc_fence_before();
    }
    mbar_arrive(q_ready);

    for (int j = 0; j < nkv; ++j) {
      const int b = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const int kv_valid = Tk - j * BKV - half * HC;    // valid columns of this thread's half (may be <= 0)
      mbar_wait(&s_full[b], ph);
      const bool trs = tr && warp == 2 && lane == 0;
      if (trs) trace[j * 8 + 3] = clock64();
      tc_fence_after();
      uint32_t s[HC];
      {
        uint32_t (&s0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[0]);
        uint32_t (&s1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[32]);
        tmem_ld_32x32(t_lane + b * BKV + half * HC + 0, s0);
        tmem_ld_32x32(t_lane + b * BKV + half * HC + 32, s1);
        tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(&s_empty[b]);                         // S_j now lives in registers
      if (trs) trace[j * 8 + 4] = clock64();
      if (kv_valid < HC) {
#pragma unroll
        for (int i = 0; i < HC; ++i)
          if (i >= kv_valid) s[i] = 0xff800000u;        // -inf
      }
      float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[1]), mx2 = __uint_as_float(s[2]),
            mx3 = __uint_as_float(s[3]);
#pragma unroll
      for (int i = 4; i < HC; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(s[i]));
        mx1 = fmaxf(mx1, __uint_as_float(s[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(s[i + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(s[i + 3]));
      }
      float bm = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      // exchange the block maximum with the thread that owns the other half of this row
      float* xm = xmax + (j & 1) * 256;
      xm[half * 128 + r] = bm;
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      bm = fmaxf(bm, xm[(half ^ 1) * 128 + r]);
      m_run = fmaxf(m_run, bm);
      float alpha = 1.f;
      bool need = false;
      if (j == 0) {
        m_used = m_run;
      } else if ((m_run - m_used) * scale_log2 > 8.0f) {
        alpha = ex2_approx((m_used - m_run) * scale_log2);
        m_used = m_run;
        need = true;
      }
      const float mb = m_used * scale_log2;
      // p = exp2(s*c - m*c) on pairs with packed fp32 FMA/ADD (FFMA2/FADD2); 3 of every 8 pairs take the polynomial
      // exp2 on the FMA pipe, the other 5 the MUFU unit, so neither pipe alone bounds the loop
      const float2 sl2 = make_float2(scale_log2, scale_log2), nmb = make_float2(-mb, -mb);
      float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
      uint32_t pk[HC / 2];
#pragma unroll
      for (int i = 0; i < HC / 2; i += 2) {
        float2 a = __ffma2_rn(make_float2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), sl2, nmb);
        float2 b2 = __ffma2_rn(make_float2(__uint_as_float(s[2 * i + 2]), __uint_as_float(s[2 * i + 3])), sl2, nmb);
        if ((i & 7) == 2) {
          a = ex2_poly2(a);
        } else {
          a.x = ex2_approx(a.x);
          a.y = ex2_approx(a.y);
        }
        if (((i + 1) & 7) == 5 || ((i + 1) & 7) == 7) {
          b2 = ex2_poly2(b2);
        } else {
          b2.x = ex2_approx(b2.x);
          b2.y = ex2_approx(b2.y);
        }
        acc0 = __fadd2_rn(acc0, a);
        acc1 = __fadd2_rn(acc1, b2);
        pk[i] = pack_bf16x2(a.x, a.y);
        pk[i + 1] = pack_bf16x2(b2.x, b2.y);
      }
      l = l * alpha + ((acc0.x + acc0.y) + (acc1.x + acc1.y));
      if (trs) trace[j * 8 + 5] = clock64();
      if (j > 0) mbar_wait(pv_done, (j - 1) & 1);       // P V_{j-1} retired: P is free, O is complete
      if (trs) trace[j * 8 + 6] = clock64();
      if (__any_sync(0xffffffffu, need)) {
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < OC; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(t_o + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st_32x32(t_o + c, v);
        }
      }
      tmem_st_32x32(t_lane + Cfg::kPCol + half * (HC / 2), pk);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
      if (trs) trace[j * 8 + 7] = clock64();
    }
    // ---- normalise, gate, store ----
    {
      xsum[half * 128 + r] = l;
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      l += xsum[(half ^ 1) * 128 + r];
      mbar_wait(pv_done, (nkv - 1) & 1);
      tc_fence_after();
      const float inv_l = 1.0f / l;
      const int row = q0 + r;
      const int b_idx = bh / H, h_idx = bh % H;
      float g = 1.f;
      if (gate_logits != nullptr && row < Tq) {
        const float z = gate_logits[(static_cast<int64_t>(b_idx) * Tq + row) * H + h_idx];
        g = 2.0f / (1.0f + __expf(-z));
      }
      const float f = inv_l * g;
      // context parallel: row `row` of head h belongs to the rank that owns that token; the store goes straight into
      // that rank's buffer over NVLink (peer pointer), fusing the head->token re-shard into this epilogue
      __nv_bfloat16* o;
      if (sc.rows_per_rank > 0) {
        const int dest = row / sc.rows_per_rank, row_l = row % sc.rows_per_rank;
        o = sc.peer[row < Tq ? dest : 0] + (static_cast<int64_t>(b_idx) * sc.rows_per_rank + row_l) * sc.pitch +
            (sc.head0 + h_idx) * DH + half * OC;
      } else {
        o = out + (static_cast<int64_t>(b_idx) * Tq + row) * (static_cast<int64_t>(H) * DH) + h_idx * DH + half * OC;
      }
#pragma unroll
      for (int c = 0; c < OC; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_o + c, v);
        tmem_ld_wait();
        if (row < Tq) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(v[i + 0]) * f, __uint_as_float(v[i + 1]) * f);
            w.y = pack_bf16x2(__uint_as_float(v[i + 2]) * f, __uint_as_float(v[i + 3]) * f);
            w.z = pack_bf16x2(__uint_as_float(v[i + 4]) * f, __uint_as_float(v[i + 5]) * f);
            w.w = pack_bf16x2(__uint_as_float(v[i + 6]) * f, __uint_as_float(v[i + 7]) * f);
            *reinterpret_cast<uint4*>(o + c + i) = w;
          }
        }
      }
      if (half == 0 && lse_out != nullptr && row < Tq)
        lse_out[static_cast<int64_t>(bh) * Tq + row] = m_used * scale + logf(l);
      tc_fence_before();
    }
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int DH, bool VROWS>
int launch_attention(const void* q, const void* k, const AttnV& v, void* out, int B, int H, int Tq, int Tk,
                     float scale, const float* gate_logits, float* lse_out, long long* trace, const AttnOutScatter& sc,
                     cudaStream_t stream) {
  using Cfg = AttnCfg<DH>;
  static PerDeviceOnce configured;
  if (configured.first()) {
    LTX2_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel<DH, VROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes));
  }
  const uint64_t BH = static_cast<uint64_t>(B) * H;
  CUtensorMap mk, mv;
  {
    uint64_t dims[3] = {static_cast<uint64_t>(DH), static_cast<uint64_t>(Tk), BH};
    uint64_t str[2] = {static_cast<uint64_t>(DH) * 2, static_cast<uint64_t>(Tk) * DH * 2};
    uint32_t box[3] = {64, BKV, 1};
    LTX2_PROPAGATE(make_tensor_map_bf16(&mk, k, 3, dims, str, box));
  }
  if (VROWS) {
    uint64_t dims[4] = {static_cast<uint64_t>(DH), static_cast<uint64_t>(Tk), static_cast<uint64_t>(H),
                        static_cast<uint64_t>(B)};
    uint64_t str[3] = {static_cast<uint64_t>(v.stride_t) * 2, static_cast<uint64_t>(v.stride_h) * 2,
                       static_cast<uint64_t>(v.stride_b) * 2};
    uint32_t box[4] = {64, BKV, 1, 1};
    LTX2_PROPAGATE(make_tensor_map_bf16(&mv, v.ptr, 4, dims, str, box));
  } else {
    uint64_t dims[3] = {static_cast<uint64_t>(Tk), static_cast<uint64_t>(DH), BH};
    uint64_t str[2] = {static_cast<uint64_t>(v.Tkp) * 2, static_cast<uint64_t>(v.Tkp) * DH * 2};
    uint32_t box[3] = {64, static_cast<uint32_t>(DH), 1};
    LTX2_PROPAGATE(make_tensor_map_bf16(&mv, v.ptr, 3, dims, str, box));
  }
  dim3 grid((Tq + BQ - 1) / BQ, static_cast<unsigned>(BH));
  const float kLog2e = 1.4426950408889634f;
  attention_kernel<DH, VROWS><<<grid, kAttnThreads, Cfg::kSmemBytes, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(q), mk, mv, reinterpret_cast<__nv_bfloat16*>(out), H, Tq, Tk,
      scale * kLog2e, scale, gate_logits, lse_out, trace, sc);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

}  // namespace

int attention_bf16_v(const void* q, const void* k, const AttnV& v, void* out, int B, int H, int Tq, int Tk, int Dh,
                     float scale, const float* gate_logits, float* lse_out, cudaStream_t stream, long long* trace,
                     const AttnOutScatter* scatter) {
  AttnOutScatter sc;
  if (scatter) sc = *scatter;
  LTX2_REQUIRE(B > 0 && H > 0 && Tq > 0 && Tk > 0, "attention: empty problem");
  LTX2_REQUIRE(static_cast<int64_t>(B) * H <= 65535, "attention: B*H too large for grid.y");
  LTX2_REQUIRE(Dh == 64 || Dh == 128, "attention: head_dim %d unsupported (64 or 128)", Dh);
  if (v.rows)
    LTX2_REQUIRE(v.stride_t % 8 == 0 && v.stride_h % 8 == 0 && v.stride_b % 8 == 0,
                 "attention: V strides must be multiples of 8 elements (16 B)");
  else
    LTX2_REQUIRE(v.Tkp >= Tk && v.Tkp % 8 == 0, "attention: V^T pitch %lld must be >= Tk=%d and a multiple of 8",
                 (long long)v.Tkp, Tk);
  // head_dim 128 (every video-stream attention): two-stream ping-pong kernel (attention_pair_sm100.cu).
  // LTX2_ATTN_KERNEL=single keeps the one-tile kernel below for A/B measurements.
  const char* env_kernel = getenv("LTX2_ATTN_KERNEL");
  const bool use_pair = !(env_kernel && strcmp(env_kernel, "single") == 0);
  if (Dh == 128 && use_pair && attention_2cta_applies(v, Tq, Tk, Dh))
    return attention_2cta_bf16(q, k, v, out, B, H, Tq, Tk, scale, gate_logits, lse_out, stream, trace, sc);
  if (Dh == 128 && use_pair)
    return attention_pair_bf16(q, k, v, out, B, H, Tq, Tk, scale, gate_logits, lse_out, stream, trace, sc);
  if (v.rows) {
    return Dh == 128 ? launch_attention<128, true>(q, k, v, out, B, H, Tq, Tk, scale, gate_logits, lse_out, trace, sc, stream)
                     : launch_attention<64, true>(q, k, v, out, B, H, Tq, Tk, scale, gate_logits, lse_out, trace, sc, stream);
  }
  return Dh == 128 ? launch_attention<128, false>(q, k, v, out, B, H, Tq, Tk, scale, gate_logits, lse_out, trace, sc, stream)
                   : launch_attention<64, false>(q, k, v, out, B, H, Tq, Tk, scale, gate_logits, lse_out, trace, sc, stream);
}

int attention_bf16(const void* q, const void* k, const void* vt, void* out, int B, int H, int Tq, int Tk, int Tkp,
                   int Dh, float scale, const float* gate_logits, float* lse_out, cudaStream_t stream, long long* trace) {
  AttnV v;
  v.ptr = vt;
  v.rows = 0;
  v.Tkp = Tkp;
  return attention_bf16_v(q, k, v, out, B, H, Tq, Tk, Dh, scale, gate_logits, lse_out, stream, trace, nullptr);
}

}  // namespace ltx2
