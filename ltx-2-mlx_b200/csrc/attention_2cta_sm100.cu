// Flash-style attention on an SM PAIR (tcgen05 cta_group::2), head_dim 128, V in row form:
//   O = softmax(Q K^T / sqrt(d)) V, non-causal, no mask.
//
// Reference op replaced: mx.fast.scaled_dot_product_attention as called from
// _compiled_attention_core_no_mask (attention.py:12-34), plus the head merge
// (B,H,T,D)->(B,T,H*D) (:34) and the V2 per-head gate 2*sigmoid(logits) (:243-250).
//
// Why an SM pair.  The one-SM kernel (attention_pair_sm100.cu) is bound by its tensor pipe even with the softmax
// arithmetic removed (profiles/r1c_attn_bench.txt, "NO softmax math": 0.70 of the burst peak): its S = Q K^T
// instructions are M 128 x N 64 with Q as a shared-memory operand -- 6 KB of operand reads per 32 tensor clocks, which
// the 128 B/clk shared memory serves in 48.  Here a cluster of two CTAs owns two adjacent 128-query tiles of one head
// and every tcgen05.mma is issued for the pair (M 256): each SM still reads its own 4 KB of Q per instruction but only
// HALF of the key operand (64 of 128 keys), so S runs at N = 128 with 6 KB per 64 clocks, and every K/V byte that
// leaves the L2 feeds 256 queries.  One thread of CTA 0 issues for both SMs.
//
// Per CTA: ONE query tile, keys in blocks of 128, S double-buffered in tensor memory (S(k+1) is computed while the
// softmax works on S(k)).  The 8 softmax warps split a block by COLUMNS: warps 0-3 (group A) own keys [0,64) of every
// block, warps 4-7 (group B) keys [64,128) -- one thread per (row, group), no cross-thread reduction.  The two groups
// keep independent running maxima / sums and accumulate into separate outputs O_A, O_B (like split-KV halves) that are
// merged once at the end, so two softmax warps per scheduler stay busy on a single query tile.
// Tensor memory (512 columns):  [0,128) S buffer 0   [128,256) S buffer 1   [256,384) O_A   [384,512) O_B
// P_g(k) (bf16, 32 columns) overwrites the first half of the group's own 64 S columns.
// Warps (384 threads per CTA): 0-7 softmax, 8 TMA producer (each CTA loads ITS query tile, ITS 64 keys of every K block
// and ITS 64 channels of every V block), 9 = MMA issuer in CTA 0 / relay ("my bytes have landed") in CTA 1, 10-11 idle.
// Barriers: K/V ring full (own TMA) + peer_full (relay -> CTA 0), empty / s_full / o_done (tcgen05.commit multicast to
// both CTAs), p_full in CTA 0 (one arrival per softmax warp of either CTA).
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace ltx2 {

namespace {

constexpr int kThreads = 384;
constexpr int kRing = 10;                   // 16 KB half tiles (K: 64 keys x 128 ch, V: 128 keys x 64 ch)
constexpr int DH = 128;
constexpr int BQ = 128;                     // queries per CTA
constexpr int BK2 = 128;                    // keys per block
constexpr int GRP = 64;                     // keys per softmax group and block
constexpr int kHalf = 64 * 128 * 2;         // 16 KB
constexpr int kQBytes = BQ * DH * 2;        // 32 KB
constexpr int kSmem = 1024 + kQBytes + kRing * kHalf + 512;
constexpr int kTrace = 16;                  // trace words per key block

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

__device__ __forceinline__ void wait_lean(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity))
    if (++spins > (1u << 22)) __trap();
}

template <int POLY, bool TRACE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
attention_2cta_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                      const __grid_constant__ CUtensorMap tmap_v, __nv_bfloat16* __restrict__ out, int H, int Tq, int Tk,
                      float scale_log2, float scale, const float* __restrict__ gate_logits,
                      float* __restrict__ lse_out, long long* __restrict__ trace,
                      const __grid_constant__ AttnOutScatter sc, int n_cl, int dbg) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // [2 x (128 rows x 128 B)]
  uint8_t* sRing = sQ + kQBytes;                        // [kRing][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRing + kRing * kHalf);
  uint64_t* q_full = bars;                              // own TMA: query tile
  uint64_t* peer_q = bars + 1;                          // CTA 0 only: CTA 1's query tile has landed
  uint64_t* full = bars + 2;                            // [kRing] own TMA
  uint64_t* peer_full = full + kRing;                   // [kRing] CTA 0 only: relay of CTA 1
  uint64_t* empty = peer_full + kRing;                  // [kRing] MMA (multicast) -> own TMA
  uint64_t* s_full = empty + kRing;                     // [2]     MMA (multicast) -> softmax: S block in buffer b
  uint64_t* p_full = s_full + 2;                        // [2]     CTA 0 only: 16 softmax warps -> MMA: P in buffer b
  uint64_t* o_done = p_full + 2;                        // MMA (multicast): the last P*V has retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x >> 1;
  const int bh = cid / n_cl;
  const int qtile = 2 * (cid % n_cl) + static_cast<int>(rank);
  const int nblk = (Tk + BK2 - 1) / BK2;
  const bool tr = TRACE && trace != nullptr && blockIdx.x == 0;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    mbar_init(peer_q, 1);
    for (int i = 0; i < kRing; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&peer_full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 16);
    }
    mbar_init(o_done, 1);
    fence_barrier_init();
  }
  __syncthreads();
  cluster_sync_all();                                   // both CTAs run and their barriers exist
  if (warp == 9) tmem_alloc2(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                           // set-up above overlapped the previous kernel's tail

  if (warp == 8) {
    // ===================== TMA producer (both CTAs) =====================
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(q_full, kQBytes);
#pragma unroll
      for (int cc = 0; cc < DH / 64; ++cc)
        tma_load_3d(sQ + cc * (BQ * 128), &tmap_q, q_full, cc * 64, qtile * BQ, bh);
    }
    int slot = 0;
    uint32_t ph = 0;
    auto load = [&](bool is_v, int blk) {
      wait_lean(&empty[slot], ph ^ 1);
      if (leader) {
        uint8_t* dst = sRing + slot * kHalf;
        mbar_expect_tx(&full[slot], kHalf);
        if (!is_v) {                                    // my 64 keys of the block: two 64-channel chunks of 8 KB
#pragma unroll
          for (int cc = 0; cc < DH / 64; ++cc)
            tma_load_3d(dst + cc * (GRP * 128), &tmap_k, &full[slot], cc * 64, blk * BK2 + static_cast<int>(rank) * GRP, bh);
        } else {                                        // my 64 channels of all 128 keys (MN-major B operand)
          tma_load_4d(dst, &tmap_v, &full[slot], static_cast<int>(rank) * 64, blk * BK2, bh % H, bh / H);
        }
      }
      __syncwarp();
      if (++slot == kRing) { slot = 0; ph ^= 1; }
    };
    load(false, 0);
    if (nblk > 1) load(false, 1);
    for (int k = 0; k < nblk; ++k) {
      load(true, k);
      if (k + 2 < nblk) load(false, k + 2);
    }
  } else if (warp == 9 && rank == 1) {
    // ===================== relay (CTA 1): my bytes have landed -> CTA 0's MMA issuer =====================
    const bool leader = elect_one();
    wait_lean(q_full, 0);
    if (leader) mbar_arrive_cluster_relaxed(mapa_u32(peer_q, 0));
    __syncwarp();
    int slot = 0;
    uint32_t ph = 0;
    for (int i = 0; i < 2 * nblk; ++i) {
      wait_lean(&full[slot], ph);
      if (leader) mbar_arrive_cluster_relaxed(mapa_u32(&peer_full[slot], 0));
      __syncwarp();
      if (++slot == kRing) { slot = 0; ph ^= 1; }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer (CTA 0, for the pair) =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = umma_idesc_bf16(2 * BQ, BK2);
    constexpr uint32_t idesc_o = umma_idesc_bf16(2 * BQ, DH, true);
    int slot = 0;
    uint32_t ph = 0;
    auto acquire = [&]() -> int {
      const int s = slot;
      // plain (CTA-scope) waits: the operands are read by the tensor core through the async proxy, never by this
      // thread; an acquire at cluster scope compiles to an L1 invalidation (CCTL.IVALL) per wait, which turned the
      // softmax warps' few local-memory reloads into L2 round trips (~300 clocks per key block, measured)
      wait_lean(&full[s], ph);
      wait_lean(&peer_full[s], ph);
      tc_fence_after();
      if (++slot == kRing) { slot = 0; ph ^= 1; }
      return s;
    };
    const uint64_t qd = umma_desc_k_sw128(smem_u32(sQ));
    // S(buffer b) = Q K^T: M 256 (two query tiles), N 128 (64 keys from each CTA), K 128 in 8 steps
    auto issue_s = [&](int b, int ks_slot) {
      const uint64_t kd = umma_desc_k_sw128(smem_u32(sRing + ks_slot * kHalf));
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks) {
          const uint64_t offq = ((ks / 4) * (BQ * 128) >> 4) + 2 * (ks % 4);
          const uint64_t offk = ((ks / 4) * (GRP * 128) >> 4) + 2 * (ks % 4);
          umma2_bf16_ss(tmem_base + b * BK2, qd + offq, kd + offk, idesc_s, ks != 0);
        }
      }
    };
    // O_g += P_g V[keys 64 g .. 64 g + 63]: M 256, N 128 (64 channels from each CTA), K 64 in 4 steps, g = A, B
    auto issue_pv = [&](int b, int v_slot, bool acc) {
      const uint64_t vd = umma_desc_mn_sw128(smem_u32(sRing + v_slot * kHalf), BK2 * 128, 1024);
      if (leader) {
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
          for (int ks = 0; ks < GRP / 16; ++ks) {
            const int kk = g * (GRP / 16) + ks;
            umma2_bf16_ts(tmem_base + 256 + g * DH, tmem_base + b * BK2 + g * GRP + ks * 8, vd + kk * (2048 >> 4),
                          idesc_o, acc || ks != 0);
          }
      }
    };
    wait_lean(q_full, 0);
    wait_lean(peer_q, 0);
    tc_fence_after();
    for (int b = 0; b < 2 && b < nblk; ++b) {
      const int s = acquire();
      issue_s(b, s);
      if (leader) {
        umma2_commit_both(&empty[s]);
        umma2_commit_both(&s_full[b]);
      }
      __syncwarp();
    }
    for (int k = 0; k < nblk; ++k) {
      const int b = k & 1;
      wait_lean(&p_full[b], (k >> 1) & 1);
      tc_fence_after();
      if (tr && leader) trace[k * kTrace + 0] = clock64();
      {
        const int s = acquire();
        if (tr && leader) trace[k * kTrace + 8] = clock64();
        if (!(TRACE && (dbg & 2))) issue_pv(b, s, k > 0);
        if (tr && leader) trace[k * kTrace + 9] = clock64();
        if (leader) umma2_commit_both(&empty[s]);
        if (tr && leader) trace[k * kTrace + 10] = clock64();
      }
      if (k + 2 < nblk) {
        const int s = acquire();
        if (tr && leader) trace[k * kTrace + 11] = clock64();
        if (!(TRACE && (dbg & 1))) issue_s(b, s);
        if (tr && leader) trace[k * kTrace + 12] = clock64();
        if (leader) umma2_commit_both(&empty[s]);
        if (tr && leader) trace[k * kTrace + 13] = clock64();
      }
      if (leader) {
        // also signalled when no further S goes into this buffer: "P(k)*V retired" is what a rescale waits for
        umma2_commit_both(&s_full[b]);
        if (k == nblk - 1) umma2_commit_both(o_done);
        if (tr) trace[k * kTrace + 1] = clock64();
      }
      __syncwarp();
    }
  } else if (warp < 8) {
    // ===================== softmax + output (warps 0..7) =====================
    const int g = warp >> 2;                            // column group: keys [64 g, 64 g + 64) of every block
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;                  // query row inside the tile
    const int row = qtile * BQ + r;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_o = t_lane + 256 + g * DH;
    const uint32_t p_remote = mapa_u32(p_full, 0);
    const bool trs = tr && lane == 0 && quarter == 0;   // warps 0 and 4 of cluster 0, CTA 0
    float m_run = -INFINITY;                            // true running row maximum (raw scores)
    float m_used = -INFINITY;                           // maximum the current scale of P, l and O refers to
    float l = 0.f;

    for (int k = 0; k < nblk; ++k) {
      const int b = k & 1;
      const int kv_valid = Tk - k * BK2 - g * GRP;      // valid keys of this group's 64 columns (may be <= 0)
      const uint32_t t_s = t_lane + b * BK2 + g * GRP;
      wait_lean(&s_full[b], (k >> 1) & 1);
      if (trs) trace[k * kTrace + 2 + 3 * g] = clock64();
      tc_fence_after();
      uint32_t s[GRP];
      {
        uint32_t (&s0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[0]);
        uint32_t (&s1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[32]);
        tmem_ld_32x32(t_s + 0, s0);
        tmem_ld_32x32(t_s + 32, s1);
        tmem_ld_wait();
      }
      if (kv_valid <= 0) {
        // nothing of this block belongs to the group: P = 0 (the MMA still runs for the other group / the other CTA)
        uint32_t z[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) z[e] = 0u;
        tmem_st_32x16(t_s, z);
        tmem_st_32x16(t_s + 16, z);
      } else {
        if (kv_valid < GRP) {
#pragma unroll
          for (int e = 0; e < GRP; ++e)
            if (e >= kv_valid) s[e] = 0xff800000u;      // -inf
        }
        float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[1]), mx2 = __uint_as_float(s[2]),
              mx3 = __uint_as_float(s[3]);
#pragma unroll
        for (int e = 4; e < GRP; e += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(s[e]));
          mx1 = fmaxf(mx1, __uint_as_float(s[e + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(s[e + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(s[e + 3]));
        }
        m_run = fmaxf(m_run, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)));
        float alpha = 1.f;
        bool need = false;
        if (k == 0) {
          m_used = m_run;
        } else if ((m_run - m_used) * scale_log2 > 8.0f) {
          alpha = ex2_approx((m_used - m_run) * scale_log2);
          m_used = m_run;
          need = true;
        }
        if (__any_sync(0xffffffffu, need)) {
          // O_g must be quiescent: P(k-1)*V is retired once the OTHER S buffer's next completion (S(k+1), or the bare
          // commit when no S(k+1) exists) is signalled; P(k)*V cannot start before this thread publishes P(k)
          wait_lean(&s_full[b ^ 1], ((k + 1) >> 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < DH; c += 32) {
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
        for (int c = 0; c < 2; ++c) {
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
      }
      if (trs) trace[k * kTrace + 3 + 3 * g] = clock64();
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      // the tensor-memory stores are complete (wait::st) and fenced; the arrival itself needs no memory ordering -- a
      // release at cluster scope here costs ~1000 clocks per block (measured)
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&p_full[b]);
        else mbar_arrive_cluster_relaxed(p_remote + 8 * b);
      }
      if (trs) trace[k * kTrace + 4 + 3 * g] = clock64();
    }

    // ---- merge the two column groups, normalise, gate, store ----
    const int b_idx = bh / H, h_idx = bh % H;
    float gt = 1.f;
    if (gate_logits != nullptr && row < Tq) {
      const float z = gate_logits[(static_cast<int64_t>(b_idx) * Tq + row) * H + h_idx];
      gt = 2.0f / (1.0f + __expf(-z));
    }
    // context parallel: row `row` of head h belongs to the rank that owns that token; the store goes straight into
    // that rank's buffer over NVLink (peer pointer), fusing the head->token re-shard into this epilogue
    __nv_bfloat16* o;
    if (sc.rows_per_rank > 0) {
      const int dest = row / sc.rows_per_rank, row_l = row % sc.rows_per_rank;
      o = sc.peer[row < Tq ? dest : 0] + (static_cast<int64_t>(b_idx) * sc.rows_per_rank + row_l) * sc.pitch +
          (sc.head0 + h_idx) * DH;
    } else {
      o = out + (static_cast<int64_t>(b_idx) * Tq + row) * (static_cast<int64_t>(H) * DH) + h_idx * DH;
    }
    wait_lean(o_done, 0);
    tc_fence_after();
    // both accumulators are visible to either group (same TMEM lanes): group g finishes output columns
    // [64 g, 64 g + 64) of the merged row
    float2* xs = reinterpret_cast<float2*>(sQ);         // Q is dead: every MMA has retired
    xs[g * 128 + r] = make_float2(m_used, l);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float2 a = xs[r], bb = xs[128 + r];
    // group B has seen no key at all when Tk <= 64: its maximum is -inf and its sum 0
    const float m = fmaxf(a.x, bb.x);
    const float fa0 = ex2_approx((a.x - m) * scale_log2);
    const float fb0 = bb.y > 0.f ? ex2_approx((bb.x - m) * scale_log2) : 0.f;
    const float lt = a.y * fa0 + bb.y * fb0;
    const float f = gt / lt;
    const float fa = fa0 * f, fb = fb0 * f;
    const uint32_t t_oa = t_lane + 256 + g * 64, t_ob = t_lane + 384 + g * 64;
#pragma unroll
    for (int c = 0; c < 64; c += 32) {
      uint32_t va[32], vb[32];
      tmem_ld_32x32(t_oa + c, va);
      tmem_ld_32x32(t_ob + c, vb);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e)
        va[e] = __float_as_uint(__uint_as_float(va[e]) * fa + __uint_as_float(vb[e]) * fb);
      if (row < Tq) {
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(va[e + 0]), __uint_as_float(va[e + 1]));
          w.y = pack_bf16x2(__uint_as_float(va[e + 2]), __uint_as_float(va[e + 3]));
          w.z = pack_bf16x2(__uint_as_float(va[e + 4]), __uint_as_float(va[e + 5]));
          w.w = pack_bf16x2(__uint_as_float(va[e + 6]), __uint_as_float(va[e + 7]));
          *reinterpret_cast<uint4*>(o + g * 64 + c + e) = w;
        }
      }
    }
    if (g == 0 && lse_out != nullptr && row < Tq) lse_out[static_cast<int64_t>(bh) * Tq + row] = m * scale + logf(lt);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // nobody leaves while the peer may still touch my barriers / smem
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

}  // namespace

bool attention_2cta_applies(const AttnV& v, int Tq, int Dh) {
  const char* env = getenv("LTX2_ATTN_2CTA");
  if (!(env && env[0] == '1')) return false;           // opt-in while the kernel is being tuned
  return Dh == 128 && v.rows != 0 && Tq > BQ;
}

int attention_2cta_bf16(const void* q, const void* k, const AttnV& v, void* out, int B, int H, int Tq, int Tk,
                        float scale, const float* gate_logits, float* lse_out, cudaStream_t stream, long long* trace,
                        const AttnOutScatter& sc) {
  const char* env_poly = getenv("LTX2_ATTN_POLY");
  const int variant = env_poly ? atoi(env_poly) : 3;
  const char* env_dbg = getenv("LTX2_ATTN_DBG");   // traced kernel only: 1 = no S MMAs in the loop, 2 = no P*V MMAs
  const int dbg = env_dbg ? atoi(env_dbg) : 0;
  static PerDeviceOnce configured;
  if (configured.first()) {
    LTX2_CUDA_CHECK(cudaFuncSetAttribute(attention_2cta_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    LTX2_CUDA_CHECK(cudaFuncSetAttribute(attention_2cta_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    LTX2_CUDA_CHECK(cudaFuncSetAttribute(attention_2cta_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    LTX2_CUDA_CHECK(cudaFuncSetAttribute(attention_2cta_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    LTX2_CUDA_CHECK(cudaFuncSetAttribute(attention_2cta_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    LTX2_CUDA_CHECK(cudaFuncSetAttribute(attention_2cta_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  }
  const uint64_t BH = static_cast<uint64_t>(B) * H;
  CUtensorMap mq, mk, mv;
  {
    uint64_t dims[3] = {static_cast<uint64_t>(DH), static_cast<uint64_t>(Tq), BH};
    uint64_t str[2] = {static_cast<uint64_t>(DH) * 2, static_cast<uint64_t>(Tq) * DH * 2};
    uint32_t box[3] = {64, BQ, 1};
    LTX2_PROPAGATE(make_tensor_map_bf16(&mq, q, 3, dims, str, box));
  }
  {
    uint64_t dims[3] = {static_cast<uint64_t>(DH), static_cast<uint64_t>(Tk), BH};
    uint64_t str[2] = {static_cast<uint64_t>(DH) * 2, static_cast<uint64_t>(Tk) * DH * 2};
    uint32_t box[3] = {64, GRP, 1};
    LTX2_PROPAGATE(make_tensor_map_bf16(&mk, k, 3, dims, str, box));
  }
  {
    uint64_t dims[4] = {static_cast<uint64_t>(DH), static_cast<uint64_t>(Tk), static_cast<uint64_t>(H),
                        static_cast<uint64_t>(B)};
    uint64_t str[3] = {static_cast<uint64_t>(v.stride_t) * 2, static_cast<uint64_t>(v.stride_h) * 2,
                       static_cast<uint64_t>(v.stride_b) * 2};
    uint32_t box[4] = {64, BK2, 1, 1};
    LTX2_PROPAGATE(make_tensor_map_bf16(&mv, v.ptr, 4, dims, str, box));
  }
  const int n_q = (Tq + BQ - 1) / BQ;
  const int n_cl = (n_q + 1) / 2;
  const unsigned grid = 2u * static_cast<unsigned>(BH) * static_cast<unsigned>(n_cl);
  const float kLog2e = 1.4426950408889634f;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
#define LTX2_LAUNCH_2CTA(P, T)                                                                                   \
  LTX2_CUDA_CHECK(launch_pdl(attention_2cta_kernel<P, T>, dim3(grid), dim3(kThreads), kSmem, stream, mq, mk, mv, o, H, \
                             Tq, Tk, scale * kLog2e, scale, gate_logits, lse_out, trace, sc, n_cl, dbg))
  if (trace != nullptr) {
    LTX2_LAUNCH_2CTA(3, true);
  } else {
    switch (variant) {
      case 0: LTX2_LAUNCH_2CTA(0, false); break;
      case 1: LTX2_LAUNCH_2CTA(1, false); break;
      case 2: LTX2_LAUNCH_2CTA(2, false); break;
      case 4: LTX2_LAUNCH_2CTA(4, false); break;
      default: LTX2_LAUNCH_2CTA(3, false); break;
    }
  }
#undef LTX2_LAUNCH_2CTA
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

}  // namespace ltx2
