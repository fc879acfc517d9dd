// Flash-style attention on tcgen05 tensor cores, head_dim 128, TWO softmax streams per CTA (ping-pong):
//   O = softmax(Q K^T / sqrt(d)) V, non-causal, no mask.
//
// Reference op replaced: mx.fast.scaled_dot_product_attention as called from
// _compiled_attention_core_no_mask (attention.py:12-34), plus the head merge
// (B,H,T,D)->(B,T,H*D) (:34) and the V2 per-head gate 2*sigmoid(logits) (:243-250).
//
// Why two streams: with one 128-query tile per CTA every 128x128 key block needs its own 64 KB of K/V from L2
// (64 B/clk/SM at tensor-pipe speed -- beyond what the L2 delivers to 148 SMs) and the softmax of block j sits on
// the critical path between S_j and P_j*V_j.  Here a CTA owns two streams that share the tensor pipe:
//   pair mode   streams = two adjacent 128-query tiles of one head over ALL keys; every K/V tile fetched from L2 is
//               used by both (half the L2->SM traffic per FLOP);
//   split mode  streams = ONE query tile over the first / second half of the keys, merged in the CTA at the end
//               (used for the odd tile of a head and for small grids such as the context-parallel head shards).
// While the softmax warpgroup of stream 0 turns S0_j into P0_j, the tensor pipe runs P1_{j-1}*V and S1_j for
// stream 1, and vice versa, so neither side waits for the other in steady state.
//
// Warps (384 threads = 3 warpgroups): 0-3 = softmax of stream 0, 4-7 = softmax of stream 1 (one thread per query
// row: no cross-thread reductions), 8 = TMA producer, 9 = MMA issuer, 10-11 idle.  The third warpgroup hands its
// registers to the softmax warpgroups (setmaxnreg 72 / 216) so a whole 128-column score row stays in registers.  Tensor memory (512 columns):
//   [0,128) S0 / P0   [128,256) S1 / P1   [256,384) O0   [384,512) O1
// P_j (bf16, 64 columns) overwrites the first half of S_j once the row has been pulled into registers; S_{j+1} is
// issued behind P_j*V_j on the in-order tensor pipe, so the overwrite is safe.  Q lives in shared memory (SS-mode
// MMA for S), P in tensor memory (TS-mode MMA for P*V).  K/V tiles (32 KB each) flow through one 5-stage ring.
// The running output is rescaled lazily (only when the row maximum grew by more than 2^8), by the softmax thread
// itself: S_j complete implies P_{j-1}*V retired (same commit group), so O is quiescent during softmax j.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace ltx2 {

namespace {

constexpr int kPairThreads = 384;
constexpr int kPairStages = 5;
constexpr int DHP = 128;                    // head dim
constexpr int BQP = 128;                    // queries per stream
constexpr int BKP = 128;                    // keys per block
constexpr int kTileBytes = 128 * 128 * 2;   // one Q / K / V tile
constexpr int kPairSmem = 1024 + 2 * kTileBytes + kPairStages * kTileBytes + 256;
constexpr int kTraceStride = 16;

// registers -> TMEM: this warp's 32 lanes x 16 consecutive columns
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

struct PairGrid {
  int n_q;        // 128-query tiles per (batch, head)
  int n_pairs;    // tiles [0, 2*n_pairs) of every head run as pairs, the rest as split-KV singles
  int BH;
};

template <bool VROWS, int POLY>
__global__ void __launch_bounds__(kPairThreads, 1)
attention_pair_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                      const __grid_constant__ CUtensorMap tmap_v, __nv_bfloat16* __restrict__ out, int H, int Tq, int Tk,
                      float scale_log2, float scale, const float* __restrict__ gate_logits,
                      float* __restrict__ lse_out, long long* __restrict__ trace, AttnOutScatter sc, PairGrid pg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // [2 tiles][2 x (128 rows x 128 B)]
  uint8_t* sRing = sQ + 2 * kTileBytes;                 // [stages][32 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRing + kPairStages * kTileBytes);
  uint64_t* q_full = bars;                              // 1
  uint64_t* full = bars + 1;                            // stages   TMA -> MMA
  uint64_t* empty = full + kPairStages;                 // stages   MMA -> TMA
  uint64_t* s_full = empty + kPairStages;               // 2        MMA -> softmax i: S_j in TMEM
  uint64_t* p_full = s_full + 2;                        // 2        softmax i -> MMA: P_j in TMEM (O rescaled if needed)
  uint64_t* o_done = p_full + 2;                        // 2        MMA -> softmax i: last P*V retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform role
  const int lane = threadIdx.x & 31;

  // ---- work item: pairs of all heads first (longest first), then the split-KV singles ----
  const int total_pairs = pg.BH * pg.n_pairs;
  bool pair;
  int bh, qa;
  {
    const int idx = blockIdx.x;
    if (idx < total_pairs) {
      pair = true;
      bh = idx / pg.n_pairs;
      qa = 2 * (idx % pg.n_pairs);
    } else {
      const int n_split = pg.n_q - 2 * pg.n_pairs;
      const int r = idx - total_pairs;
      pair = false;
      bh = r / n_split;
      qa = 2 * pg.n_pairs + r % n_split;
    }
  }
  const int nkv = (Tk + BKP - 1) / BKP;
  const int n0 = pair ? nkv : (nkv + 1) / 2;            // key blocks of stream 0
  const int n1 = pair ? nkv : nkv - n0;                 // key blocks of stream 1 (may be 0 in split mode)
  const int kv1 = pair ? 0 : n0;                        // first key block of stream 1
  const bool tr = trace != nullptr && blockIdx.x == 0;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < kPairStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8) {
  // third warpgroup: one setmaxnreg for all four warps, then the roles
  asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
  if (warp == 8) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(q_full, pair ? 2 * kTileBytes : kTileBytes);
      for (int t = 0; t < (pair ? 2 : 1); ++t)
#pragma unroll
        for (int i = 0; i < DHP / 64; ++i)
          tma_load_3d(sQ + t * kTileBytes + i * (BQP * 128), &tmap_q, q_full, i * 64, (qa + t) * BQP, bh);
    }
    int slot = 0;
    uint32_t ph = 0;
    auto load_k = [&](int blk) {
      mbar_wait(&empty[slot], ph ^ 1);
      if (leader) {
        uint8_t* dst = sRing + slot * kTileBytes;
        mbar_expect_tx(&full[slot], kTileBytes);
#pragma unroll
        for (int i = 0; i < DHP / 64; ++i) tma_load_3d(dst + i * (BKP * 128), &tmap_k, &full[slot], i * 64, blk * BKP, bh);
      }
      if (++slot == kPairStages) { slot = 0; ph ^= 1; }
    };
    auto load_v = [&](int blk) {
      mbar_wait(&empty[slot], ph ^ 1);
      if (leader) {
        uint8_t* dst = sRing + slot * kTileBytes;
        mbar_expect_tx(&full[slot], kTileBytes);
        if (VROWS) {   // V rows [keys, d]: one 128-key x 64-channel box per 64 channels (MN-major B operand)
#pragma unroll
          for (int i = 0; i < DHP / 64; ++i)
            tma_load_4d(dst + i * (BKP * 128), &tmap_v, &full[slot], i * 64, blk * BKP, bh % H, bh / H);
        } else {       // V^T [d, keys]: one DH x 64-key box per 64 keys (K-major B operand)
#pragma unroll
          for (int i = 0; i < BKP / 64; ++i)
            tma_load_3d(dst + i * (DHP * 128), &tmap_v, &full[slot], blk * BKP + i * 64, 0, bh);
        }
      }
      if (++slot == kPairStages) { slot = 0; ph ^= 1; }
    };
    // tile order = the order the MMA warp consumes them (see below)
    load_k(0);
    if (!pair && n1 > 0) load_k(kv1);
    for (int j = 0; j < n0; ++j) {
      load_v(j);
      if (j + 1 < n0) load_k(j + 1);
      if (!pair) {
        if (j < n1) load_v(kv1 + j);
        if (j + 1 < n1) load_k(kv1 + j + 1);
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = umma_idesc_bf16(BQP, BKP);
    constexpr uint32_t idesc_o = umma_idesc_bf16(BQP, DHP, VROWS);
    int slot = 0;
    uint32_t ph = 0;
    int cur_slot = 0;
    uint32_t cur_addr = 0;
    auto acquire = [&]() {
      mbar_wait(&full[slot], ph);
      tc_fence_after();
      cur_slot = slot;
      cur_addr = smem_u32(sRing + slot * kTileBytes);
      if (++slot == kPairStages) { slot = 0; ph ^= 1; }
    };
    auto release = [&](int s) {
      if (leader) umma_commit(&empty[s]);
    };
    auto issue_s = [&](int i, uint32_t k_addr) {
      const uint64_t qd = umma_desc_k_sw128(smem_u32(sQ + ((pair && i) ? kTileBytes : 0)));
      const uint64_t kd = umma_desc_k_sw128(k_addr);
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < DHP / 16; ++ks) {
          const uint64_t off = ((ks / 4) * (128 * 128) >> 4) + 2 * (ks % 4);
          umma_bf16_ss(tmem_base + i * BKP, qd + off, kd + off, idesc_s, ks != 0);
        }
        umma_commit(&s_full[i]);
      }
    };
    auto issue_pv = [&](int i, uint32_t v_addr, bool acc) {
      const uint64_t vd = VROWS ? umma_desc_mn_sw128(v_addr, BKP * 128, 1024) : umma_desc_k_sw128(v_addr);
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < BKP / 16; ++ks)
          umma_bf16_ts(tmem_base + 256 + i * DHP, tmem_base + i * BKP + ks * 8,
                       vd + (VROWS ? ks * (2048 >> 4) : ((ks / 4) * (DHP * 128) >> 4) + 2 * (ks % 4)), idesc_o,
                       acc || ks != 0);
      }
    };
    mbar_wait(q_full, 0);
    tc_fence_after();
    acquire();
    issue_s(0, cur_addr);
    if (pair) {
      issue_s(1, cur_addr);
      release(cur_slot);
    } else {
      release(cur_slot);
      if (n1 > 0) {
        acquire();
        issue_s(1, cur_addr);
        release(cur_slot);
      }
    }
    __syncwarp();
    for (int j = 0; j < n0; ++j) {
      int v_slot, k_slot = 0;
      uint32_t v_addr, k_addr = 0;
      // ---- stream 0: O0 += P0_j V_j, then S0_{j+1} ----
      mbar_wait(&p_full[0], j & 1);
      tc_fence_after();
      if (tr && leader) trace[j * kTraceStride + 0] = clock64();
      acquire();
      v_slot = cur_slot;
      v_addr = cur_addr;
      issue_pv(0, v_addr, j > 0);
      if (!pair) release(v_slot);
      if (j + 1 < n0) {
        acquire();
        k_slot = cur_slot;
        k_addr = cur_addr;
        issue_s(0, k_addr);
        if (!pair) release(k_slot);
      } else if (leader) {
        umma_commit(&o_done[0]);
      }
      if (tr && leader) trace[j * kTraceStride + 1] = clock64();
      __syncwarp();
      // ---- stream 1 ----
      if (j < n1) {
        mbar_wait(&p_full[1], j & 1);
        tc_fence_after();
        if (tr && leader) trace[j * kTraceStride + 2] = clock64();
        if (!pair) {
          acquire();
          v_slot = cur_slot;
          v_addr = cur_addr;
        }
        issue_pv(1, v_addr, j > 0);
        release(v_slot);
        if (j + 1 < n1) {
          if (!pair) {
            acquire();
            k_slot = cur_slot;
            k_addr = cur_addr;
          }
          issue_s(1, k_addr);
          release(k_slot);
        } else if (leader) {
          umma_commit(&o_done[1]);
        }
        if (tr && leader) trace[j * kTraceStride + 3] = clock64();
        __syncwarp();
      }
    }
  }
  } else {
    // ===================== softmax + output (warps 0..7) =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int i = warp >> 2;                            // stream
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;                  // query row inside the tile
    const int qtile = pair ? qa + i : qa;
    const int row = qtile * BQP + r;
    const int nblk = i == 0 ? n0 : n1;
    const int kvb = i == 0 ? 0 : kv1;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_s = t_lane + i * BKP;
    const uint32_t t_o = t_lane + 256 + i * DHP;
    const bool trs = tr && lane == 0 && quarter == 0;   // warps 0 and 4
    const int tro = 4 + 4 * i;
    float m_run = -INFINITY;                            // true running row maximum (raw scores)
    float m_used = -INFINITY;                           // maximum the current scale of P, l and O refers to
    float l = 0.f;

    for (int j = 0; j < nblk; ++j) {
      const int kv_valid = Tk - (kvb + j) * BKP;
      mbar_wait(&s_full[i], j & 1);
      if (trs) trace[j * kTraceStride + tro + 0] = clock64();
      tc_fence_after();
      uint32_t s[BKP];
      {
        uint32_t (&s0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[0]);
        uint32_t (&s1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[32]);
        uint32_t (&s2)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[64]);
        uint32_t (&s3)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[96]);
        tmem_ld_32x32(t_s + 0, s0);
        tmem_ld_32x32(t_s + 32, s1);
        tmem_ld_32x32(t_s + 64, s2);
        tmem_ld_32x32(t_s + 96, s3);
        tmem_ld_wait();
      }
      if (trs) trace[j * kTraceStride + tro + 1] = clock64();
      if (POLY == 8) {   // diagnostics: no softmax arithmetic -> the tensor-pipe + hand-off floor of this structure
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) pk[e] = s[e] & 0x3f803f80u;
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_st_32x16(t_s + c * 16, pk);
        l = 1.f;
        m_used = 0.f;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[i]);
        continue;
      }
      if (kv_valid < BKP) {
#pragma unroll
        for (int e = 0; e < BKP; ++e)
          if (e >= kv_valid) s[e] = 0xff800000u;        // -inf
      }
      float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[1]), mx2 = __uint_as_float(s[2]),
            mx3 = __uint_as_float(s[3]);
#pragma unroll
      for (int e = 4; e < BKP; e += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(s[e]));
        mx1 = fmaxf(mx1, __uint_as_float(s[e + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(s[e + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(s[e + 3]));
      }
      m_run = fmaxf(m_run, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)));
      float alpha = 1.f;
      bool need = false;
      if (j == 0) {
        m_used = m_run;
      } else if ((m_run - m_used) * scale_log2 > 8.0f) {
        alpha = ex2_approx((m_used - m_run) * scale_log2);
        m_used = m_run;
        need = true;
      }
      if (__any_sync(0xffffffffu, need)) {
        // O_i is quiescent here (P_{j-1}*V retired before S_j was signalled)
#pragma unroll
        for (int c = 0; c < DHP; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(t_o + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
          tmem_st_32x32(t_o + c, v);
        }
      }
      const float mb = m_used * scale_log2;
      // p = exp2(s*c - m*c) on pairs with packed fp32 FMA/ADD; POLY of every 8 pairs take the polynomial exp2 on the
      // FMA pipe, the others the MUFU unit, so neither pipe alone bounds the loop
      const float2 sl2 = make_float2(scale_log2, scale_log2), nmb = make_float2(-mb, -mb);
      float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int ii = 0; ii < 16; ii += 2) {
          const int e = c * 32 + 2 * ii;
          float2 a = __ffma2_rn(make_float2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), sl2, nmb);
          float2 b2 = __ffma2_rn(make_float2(__uint_as_float(s[e + 2]), __uint_as_float(s[e + 3])), sl2, nmb);
          const bool pa = (POLY >= 1 && (ii & 7) == 2) || (POLY >= 4 && (ii & 7) == 6);
          const bool pb = (POLY >= 2 && ((ii + 1) & 7) == 5) || (POLY >= 3 && ((ii + 1) & 7) == 7);
          if (pa) {
            a = ex2_poly2(a);
          } else {
            a.x = ex2_approx(a.x);
            a.y = ex2_approx(a.y);
          }
          if (pb) {
            b2 = ex2_poly2(b2);
          } else {
            b2.x = ex2_approx(b2.x);
            b2.y = ex2_approx(b2.y);
          }
          acc0 = __fadd2_rn(acc0, a);
          acc1 = __fadd2_rn(acc1, b2);
          pk[ii] = pack_bf16x2(a.x, a.y);
          pk[ii + 1] = pack_bf16x2(b2.x, b2.y);
        }
        tmem_st_32x16(t_s + c * 16, pk);
      }
      l = l * alpha + ((acc0.x + acc0.y) + (acc1.x + acc1.y));
      if (trs) trace[j * kTraceStride + tro + 2] = clock64();
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[i]);
      if (trs) trace[j * kTraceStride + tro + 3] = clock64();
    }

    // ---- normalise, gate, store ----
    const int b_idx = bh / H, h_idx = bh % H;
    float g = 1.f;
    if (gate_logits != nullptr && row < Tq) {
      const float z = gate_logits[(static_cast<int64_t>(b_idx) * Tq + row) * H + h_idx];
      g = 2.0f / (1.0f + __expf(-z));
    }
    // context parallel: row `row` of head h belongs to the rank that owns that token; the store goes straight into
    // that rank's buffer over NVLink (peer pointer), fusing the head->token re-shard into this epilogue
    __nv_bfloat16* o;
    if (sc.rows_per_rank > 0) {
      const int dest = row / sc.rows_per_rank, row_l = row % sc.rows_per_rank;
      o = sc.peer[row < Tq ? dest : 0] + (static_cast<int64_t>(b_idx) * sc.rows_per_rank + row_l) * sc.pitch +
          (sc.head0 + h_idx) * DHP;
    } else {
      o = out + (static_cast<int64_t>(b_idx) * Tq + row) * (static_cast<int64_t>(H) * DHP) + h_idx * DHP;
    }
    if (pair) {
      mbar_wait(&o_done[i], 0);
      tc_fence_after();
      const float f = g / l;
#pragma unroll
      for (int c = 0; c < DHP; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_o + c, v);
        tmem_ld_wait();
        if (row < Tq) {
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(v[e + 0]) * f, __uint_as_float(v[e + 1]) * f);
            w.y = pack_bf16x2(__uint_as_float(v[e + 2]) * f, __uint_as_float(v[e + 3]) * f);
            w.z = pack_bf16x2(__uint_as_float(v[e + 4]) * f, __uint_as_float(v[e + 5]) * f);
            w.w = pack_bf16x2(__uint_as_float(v[e + 6]) * f, __uint_as_float(v[e + 7]) * f);
            *reinterpret_cast<uint4*>(o + c + e) = w;
          }
        }
      }
      if (lse_out != nullptr && row < Tq) lse_out[static_cast<int64_t>(bh) * Tq + row] = m_used * scale + logf(l);
    } else {
      // merge the two key halves of this query tile: both accumulators are visible to either warpgroup (same TMEM
      // lanes), so stream i's threads finish output columns [64 i, 64 i + 64) of the merged row
      mbar_wait(&o_done[0], 0);
      if (n1 > 0) mbar_wait(&o_done[1], 0);
      tc_fence_after();
      float2* xs = reinterpret_cast<float2*>(sQ);       // Q is dead: every MMA has retired
      xs[i * 128 + r] = make_float2(m_used, l);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float2 a = xs[r], b = xs[128 + r];
      float fa = 1.f, fb = 0.f, m = a.x;
      if (n1 > 0) {
        m = fmaxf(a.x, b.x);
        fa = ex2_approx((a.x - m) * scale_log2);
        fb = ex2_approx((b.x - m) * scale_log2);
      }
      const float lt = a.y * fa + (n1 > 0 ? b.y * fb : 0.f);
      const float f = g / lt;
      fa *= f;
      fb *= f;
      const uint32_t t_oa = t_lane + 256 + i * 64, t_ob = t_lane + 384 + i * 64;
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t va[32], vb[32];
        tmem_ld_32x32(t_oa + c, va);
        if (n1 > 0) tmem_ld_32x32(t_ob + c, vb);
        tmem_ld_wait();
        if (n1 > 0) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            va[e] = __float_as_uint(__uint_as_float(va[e]) * fa + __uint_as_float(vb[e]) * fb);
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) va[e] = __float_as_uint(__uint_as_float(va[e]) * fa);
        }
        if (row < Tq) {
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(va[e + 0]), __uint_as_float(va[e + 1]));
            w.y = pack_bf16x2(__uint_as_float(va[e + 2]), __uint_as_float(va[e + 3]));
            w.z = pack_bf16x2(__uint_as_float(va[e + 4]), __uint_as_float(va[e + 5]));
            w.w = pack_bf16x2(__uint_as_float(va[e + 6]), __uint_as_float(va[e + 7]));
            *reinterpret_cast<uint4*>(o + i * 64 + c + e) = w;
          }
        }
      }
      if (i == 0 && lse_out != nullptr && row < Tq)
        lse_out[static_cast<int64_t>(bh) * Tq + row] = m * scale + logf(lt);
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// list-scheduling makespan (in half-tile units) of P pair items (cost 2) followed by S split items (cost 1) on n SMs
int makespan(int P, int S, int n) {
  const int r = P / n, rem = P % n;
  // rem SMs carry 2(r+1), n-rem carry 2r; splits fill the lightest first
  int lo = n - rem;                // SMs at load 2r
  int span = rem > 0 ? 2 * (r + 1) : 2 * r;
  if (S <= 0) return span;
  if (rem == 0) return 2 * r + (S + n - 1) / n;
  // fill the lo SMs up to 2r+2 (two splits each), then everyone evenly
  const int cap = 2 * lo;
  if (S <= cap) return span > 2 * r + (S + lo - 1) / lo ? span : 2 * r + (S + lo - 1) / lo;
  return 2 * (r + 1) + (S - cap + n - 1) / n;
}

template <bool VROWS>
int launch_pair(const void* q, const void* k, const AttnV& v, void* out, int B, int H, int Tq, int Tk, float scale,
                const float* gate_logits, float* lse_out, long long* trace, const AttnOutScatter& sc,
                cudaStream_t stream) {
  // diagnostics / tuning knobs, read per launch (a getenv is noise next to a kernel launch):
  //   LTX2_ATTN_POLY   pairs of every 8 that take the polynomial exp2 (0, 2, 3 = default, 4)
  //   LTX2_ATTN_PAIRS  -1 = as many pair items as possible, n >= 0 = exactly n pair items per head (0 = all split-KV)
  const char* env_poly = getenv("LTX2_ATTN_POLY");
  const int variant = env_poly ? atoi(env_poly) : 3;
  const char* env_pairs = getenv("LTX2_ATTN_PAIRS");
  const int force_pairs = env_pairs ? atoi(env_pairs) : -2;
  static bool configured = false;
  if (!configured) {
    LTX2_CUDA_CHECK(cudaFuncSetAttribute(attention_pair_kernel<VROWS, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem));
    LTX2_CUDA_CHECK(cudaFuncSetAttribute(attention_pair_kernel<VROWS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem));
    LTX2_CUDA_CHECK(cudaFuncSetAttribute(attention_pair_kernel<VROWS, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem));
    LTX2_CUDA_CHECK(cudaFuncSetAttribute(attention_pair_kernel<VROWS, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem));
    LTX2_CUDA_CHECK(cudaFuncSetAttribute(attention_pair_kernel<VROWS, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem));
    configured = true;
  }
  const uint64_t BH = static_cast<uint64_t>(B) * H;
  CUtensorMap mq, mk, mv;
  {
    uint64_t dims[3] = {static_cast<uint64_t>(DHP), static_cast<uint64_t>(Tq), BH};
    uint64_t str[2] = {static_cast<uint64_t>(DHP) * 2, static_cast<uint64_t>(Tq) * DHP * 2};
    uint32_t box[3] = {64, BQP, 1};
    LTX2_PROPAGATE(make_tensor_map_bf16(&mq, q, 3, dims, str, box));
  }
  {
    uint64_t dims[3] = {static_cast<uint64_t>(DHP), static_cast<uint64_t>(Tk), BH};
    uint64_t str[2] = {static_cast<uint64_t>(DHP) * 2, static_cast<uint64_t>(Tk) * DHP * 2};
    uint32_t box[3] = {64, BKP, 1};
    LTX2_PROPAGATE(make_tensor_map_bf16(&mk, k, 3, dims, str, box));
  }
  if (VROWS) {
    uint64_t dims[4] = {static_cast<uint64_t>(DHP), static_cast<uint64_t>(Tk), static_cast<uint64_t>(H),
                        static_cast<uint64_t>(B)};
    uint64_t str[3] = {static_cast<uint64_t>(v.stride_t) * 2, static_cast<uint64_t>(v.stride_h) * 2,
                       static_cast<uint64_t>(v.stride_b) * 2};
    uint32_t box[4] = {64, BKP, 1, 1};
    LTX2_PROPAGATE(make_tensor_map_bf16(&mv, v.ptr, 4, dims, str, box));
  } else {
    uint64_t dims[3] = {static_cast<uint64_t>(Tk), static_cast<uint64_t>(DHP), BH};
    uint64_t str[2] = {static_cast<uint64_t>(v.Tkp) * 2, static_cast<uint64_t>(v.Tkp) * DHP * 2};
    uint32_t box[3] = {64, static_cast<uint32_t>(DHP), 1};
    LTX2_PROPAGATE(make_tensor_map_bf16(&mv, v.ptr, 3, dims, str, box));
  }
  // choose how many tiles of each head run as pairs: minimal makespan over the SMs, ties -> more pairs (less L2 traffic)
  PairGrid pg;
  pg.n_q = (Tq + BQP - 1) / BQP;
  pg.BH = static_cast<int>(BH);
  const int nsm = num_sms();
  int best_np = 0, best_span = 1 << 30;
  for (int np = pg.n_q / 2; np >= 0; --np) {
    const int span = makespan(pg.BH * np, pg.BH * (pg.n_q - 2 * np), nsm);
    if (span < best_span) {
      best_span = span;
      best_np = np;
    }
  }
  if (force_pairs == -1) best_np = pg.n_q / 2;
  if (force_pairs >= 0 && force_pairs <= pg.n_q / 2) best_np = force_pairs;
  pg.n_pairs = best_np;
  const unsigned grid = static_cast<unsigned>(pg.BH * (pg.n_q - pg.n_pairs));
  const float kLog2e = 1.4426950408889634f;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
#define LTX2_LAUNCH_PAIR(P)                                                                                        \
  attention_pair_kernel<VROWS, P><<<grid, kPairThreads, kPairSmem, stream>>>(mq, mk, mv, o, H, Tq, Tk, scale * kLog2e, \
                                                                             scale, gate_logits, lse_out, trace, sc, pg)
  switch (variant) {
    case 0: LTX2_LAUNCH_PAIR(0); break;
    case 2: LTX2_LAUNCH_PAIR(2); break;
    case 4: LTX2_LAUNCH_PAIR(4); break;
    case 8: LTX2_LAUNCH_PAIR(8); break;
    default: LTX2_LAUNCH_PAIR(3); break;
  }
#undef LTX2_LAUNCH_PAIR
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

}  // namespace

int attention_pair_bf16(const void* q, const void* k, const AttnV& v, void* out, int B, int H, int Tq, int Tk,
                        float scale, const float* gate_logits, float* lse_out, cudaStream_t stream, long long* trace,
                        const AttnOutScatter& sc) {
  return v.rows ? launch_pair<true>(q, k, v, out, B, H, Tq, Tk, scale, gate_logits, lse_out, trace, sc, stream)
                : launch_pair<false>(q, k, v, out, B, H, Tq, Tk, scale, gate_logits, lse_out, trace, sc, stream);
}

}  // namespace ltx2
