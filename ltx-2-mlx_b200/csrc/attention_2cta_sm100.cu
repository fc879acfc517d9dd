// Flash-style attention on an SM PAIR (tcgen05 cta_group::2), head_dim 128, V in row form; persistent, stream-K:
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
//
// Persistent + stream-K.  A work item is (batch*head, pair of query tiles) x all key blocks.  The grid is one cluster
// per SM pair; the (item, key block) space is cut into equal contiguous ranges, one per cluster ("split" schedule), or
// whole items go round-robin over the clusters when that is as good (host decision).  A range therefore starts and/or
// ends inside an item: such SEGMENTS write their unnormalised output and (max, sum) per row to a scratch slot, and the
// last segment of an item to finish (a counter per query tile) merges the parts in index order -- deterministic --
// and stores the result.  32 heads x 14 pairs = 448 items no longer cost 7 waves on 74 SM pairs but 448 / 74 = 6.05.
// Barriers are initialised, tensor memory allocated and the pipeline filled ONCE per CTA: the next segment's Q / K / V
// loads and its first two S blocks run under the previous segment's last softmax blocks and epilogue.
//
// Warps (384 threads per CTA): 0-7 softmax + epilogue, 8 TMA producer (each CTA loads ITS query tile, ITS 64 keys of
// every K block and ITS 64 channels of every V block), 9 = MMA issuer in CTA 0 / relay ("my bytes have landed") in
// CTA 1, 10-11 idle.
// Shared memory: Q tile 32 KB, ring of 5 entries x (V half 16 KB + K half 16 KB): entry of step j = V(j) and K(j+2).
// Barriers: ring full (own TMA) + peer_full (relay -> CTA 0); empty / q_empty (tcgen05.commit to CTA 0, whose
// producer forwards them to CTA 1 -- a multicast commit costs the issuing thread ~350 clocks, measured); s_full /
// o_done (commit multicast to both CTAs); p_full in CTA 0 (one arrival per softmax warp of either CTA).
#include <stdlib.h>

#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"
#include "kernels.h"

namespace ltx2 {

namespace {

constexpr int kThreads = 384;
constexpr int kRing = 5;                    // entries of 32 KB: [V half: 128 keys x 64 ch][K half: 64 keys x 128 ch]
constexpr int DH = 128;
constexpr int BQ = 128;                     // queries per CTA
constexpr int BK2 = 128;                    // keys per block
constexpr int GRP = 64;                     // keys per softmax group and block
constexpr int kHalf = 64 * 128 * 2;         // 16 KB
constexpr int kEntry = 2 * kHalf;
constexpr int kQBytes = BQ * DH * 2;        // 32 KB
constexpr int kXsBytes = 2 * BQ * 8 + 64;   // (max, sum) exchange between the column groups + the merge flag
constexpr int kSmem = 1024 + kQBytes + kRing * kEntry + kXsBytes + 512;
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

// kernel parameter: the work decomposition (pure integers, identical in every CTA)
struct Sched2 {
  int n_items;   // (batch * head) x query-tile pairs
  int n_cl;      // query-tile pairs per (batch, head)
  int nblk;      // key blocks per item
  int C;         // clusters in the grid
  int split;     // 1: equal contiguous (item, key block) ranges per cluster, 0: whole items round-robin
};

struct Seg {
  int item, kb0, kb1;       // key blocks [kb0, kb1) of item
  int part, parts;          // this segment is part `part` of `parts` of its item (parts == 1: direct epilogue)
  int slot, slot0, c_first; // scratch slot of this segment / of part 0 (parts p >= 1 use slot 2 * (c_first + p))
};

struct Walker {
  long long G, g, g0, g1;
  int C, c, nblk, n_items, it;
  bool split;
  __host__ __device__ Walker(const Sched2& s, int cluster)
      : G(static_cast<long long>(s.n_items) * s.nblk), C(s.C), c(cluster), nblk(s.nblk), n_items(s.n_items),
        it(cluster), split(s.split != 0) {
    g0 = G * c / C;
    g1 = G * (c + 1) / C;
    g = g0;
  }
  __host__ __device__ bool next(Seg& sg) {
    if (!split) {
      if (it >= n_items) return false;
      sg.item = it;
      sg.kb0 = 0;
      sg.kb1 = nblk;
      sg.part = 0;
      sg.parts = 1;
      sg.slot = sg.slot0 = sg.c_first = 0;
      it += C;
      return true;
    }
    if (g >= g1) return false;
    const int item = static_cast<int>(g / nblk);
    const long long x0 = static_cast<long long>(item) * nblk;
    const int kb0 = static_cast<int>(g - x0);
    const long long rem = g1 - g;
    const int len = rem < nblk - kb0 ? static_cast<int>(rem) : nblk - kb0;
    // the cluster that owns position x is floor(((x + 1) C - 1) / G)
    const int c_first = static_cast<int>(((x0 + 1) * C - 1) / G);
    const int c_last = static_cast<int>(((x0 + nblk) * C - 1) / G);
    sg.item = item;
    sg.kb0 = kb0;
    sg.kb1 = kb0 + len;
    sg.parts = c_last - c_first + 1;
    sg.part = c - c_first;
    sg.c_first = c_first;
    sg.slot = 2 * c + (g == g0 ? 0 : 1);
    sg.slot0 = 2 * c_first + (x0 == G * c_first / C ? 0 : 1);
    g += len;
    return true;
  }
};

struct Scratch2 {
  float* part_o;     // [slots * 2][128 rows][128] unnormalised partial outputs
  float2* part_ml;   // [slots * 2][128 rows] (max, sum) the partial output refers to
  int* sem;          // [slots * 2] finished parts per query tile (indexed by the slot of part 0); reset by the merger
};

template <int POLY, bool TRACE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
attention_2cta_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                      const __grid_constant__ CUtensorMap tmap_v, __nv_bfloat16* __restrict__ out, int H, int Tq, int Tk,
                      float scale_log2, float scale, const float* __restrict__ gate_logits,
                      float* __restrict__ lse_out, long long* __restrict__ trace,
                      const __grid_constant__ AttnOutScatter sc, const __grid_constant__ Sched2 sched, Scratch2 scr,
                      int dbg) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // [2 x (128 rows x 128 B)]
  uint8_t* sRing = sQ + kQBytes;                        // [kRing][V half | K half]
  float2* xs = reinterpret_cast<float2*>(sRing + kRing * kEntry);   // [2 groups][128 rows]
  int* xflag = reinterpret_cast<int*>(xs + 2 * BQ);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xs) + kXsBytes);
  uint64_t* q_full = bars;                              // own TMA: query tile
  uint64_t* peer_q = bars + 1;                          // CTA 0 only: CTA 1's query tile has landed
  uint64_t* q_empty = bars + 2;                         // the segment's last S has been issued and retired
  uint64_t* full = bars + 3;                            // [kRing] own TMA
  uint64_t* peer_full = full + kRing;                   // [kRing] CTA 0 only: relay of CTA 1
  uint64_t* empty = peer_full + kRing;                  // [kRing] MMA commit (CTA 0) / CTA 0's producer (CTA 1)
  uint64_t* s_full = empty + kRing;                     // [2]     MMA (multicast) -> softmax: S block in buffer b
  uint64_t* p_full = s_full + 2;                        // [2]     CTA 0 only: 16 softmax warps -> MMA: P in buffer b
  uint64_t* o_done = p_full + 2;                        // MMA (multicast): the segment's last P*V has retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x >> 1;
  const int nblk = sched.nblk;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    mbar_init(peer_q, 1);
    mbar_init(q_empty, 1);
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

  Walker walk(sched, cid);
  Seg sg;

  if (warp == 8) {
    // ===================== TMA producer (both CTAs) =====================
    const bool leader = elect_one();
    int slot = 0;
    uint32_t ph = 0, qph = 0;
    bool wrapped = false, first_seg = true;
    // ring entry: V half of key block `vb` (none: -1) and K half of key block `kb` (none: -1)
    auto load = [&](int bh, int vb, int kb) {
      wait_lean(&empty[slot], ph ^ 1);
      if (leader) {
        // the MMA warp's commit reaches CTA 0 only; forward it (no memory ordering needed: the slot is rewritten by
        // TMA, after the tensor core has finished reading it).  Not on the first pass over the ring: those waits
        // return on the fresh barrier, there was no commit to forward.
        if (rank == 0 && wrapped) mbar_arrive_cluster_relaxed(mapa_u32(&empty[slot], 1));
        uint8_t* dst = sRing + slot * kEntry;
        mbar_expect_tx(&full[slot], (vb >= 0 ? kHalf : 0) + (kb >= 0 ? kHalf : 0));
        if (vb >= 0)                                    // my 64 channels of all 128 keys (MN-major B operand)
          tma_load_4d(dst, &tmap_v, &full[slot], static_cast<int>(rank) * 64, vb * BK2, bh % H, bh / H);
        if (kb >= 0) {                                  // my 64 keys of the block: two 64-channel chunks of 8 KB
#pragma unroll
          for (int cc = 0; cc < DH / 64; ++cc)
            tma_load_3d(dst + kHalf + cc * (GRP * 128), &tmap_k, &full[slot], cc * 64,
                        kb * BK2 + static_cast<int>(rank) * GRP, bh);
        }
      }
      __syncwarp();
      if (++slot == kRing) { slot = 0; ph ^= 1; wrapped = true; }
    };
    while (walk.next(sg)) {
      const int bh = sg.item / sched.n_cl;
      const int qtile = 2 * (sg.item % sched.n_cl) + static_cast<int>(rank);
      const int n = sg.kb1 - sg.kb0;
      wait_lean(q_empty, qph ^ 1);
      if (leader) {
        if (rank == 0 && !first_seg) mbar_arrive_cluster_relaxed(mapa_u32(q_empty, 1));
        mbar_expect_tx(q_full, kQBytes);
#pragma unroll
        for (int cc = 0; cc < DH / 64; ++cc)
          tma_load_3d(sQ + cc * (BQ * 128), &tmap_q, q_full, cc * 64, qtile * BQ, bh);
      }
      __syncwarp();
      qph ^= 1;
      first_seg = false;
      load(bh, -1, sg.kb0);
      if (n > 1) load(bh, -1, sg.kb0 + 1);
      for (int j = 0; j < n; ++j) load(bh, sg.kb0 + j, j + 2 < n ? sg.kb0 + j + 2 : -1);
    }
  } else if (warp == 9 && rank == 1) {
    // ===================== relay (CTA 1): my bytes have landed -> CTA 0's MMA issuer =====================
    const bool leader = elect_one();
    int slot = 0;
    uint32_t ph = 0, qph = 0;
    while (walk.next(sg)) {
      const int n = sg.kb1 - sg.kb0;
      wait_lean(q_full, qph);
      if (leader) mbar_arrive_cluster_relaxed(mapa_u32(peer_q, 0));
      __syncwarp();
      qph ^= 1;
      const int entries = n + (n > 1 ? 2 : 1);
      for (int i = 0; i < entries; ++i) {
        wait_lean(&full[slot], ph);
        if (leader) mbar_arrive_cluster_relaxed(mapa_u32(&peer_full[slot], 0));
        __syncwarp();
        if (++slot == kRing) { slot = 0; ph ^= 1; }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer (CTA 0, for the pair) =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = umma_idesc_bf16(2 * BQ, BK2);
    constexpr uint32_t idesc_o = umma_idesc_bf16(2 * BQ, DH, true);
    int slot = 0;
    uint32_t ph = 0, qph = 0;
    uint32_t np0 = 0, np1 = 0;                          // p_full phases consumed per S buffer
    // plain (CTA-scope) waits: the operands are read by the tensor core through the async proxy, never by this
    // thread; an acquire at cluster scope compiles to an L1 invalidation (CCTL.IVALL) per wait
    auto acquire = [&]() -> int {
      const int s = slot;
      wait_lean(&full[s], ph);
      wait_lean(&peer_full[s], ph);
      tc_fence_after();
      if (++slot == kRing) { slot = 0; ph ^= 1; }
      return s;
    };
    const uint64_t qd = umma_desc_k_sw128(smem_u32(sQ));
    // S(buffer b) = Q K^T: M 256 (two query tiles), N 128 (64 keys from each CTA), K 128 in 8 steps
    auto issue_s = [&](int b, int e) {
      const uint64_t kd = umma_desc_k_sw128(smem_u32(sRing + e * kEntry + kHalf));
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
    auto issue_pv = [&](int b, int e, bool acc) {
      const uint64_t vd = umma_desc_mn_sw128(smem_u32(sRing + e * kEntry), BK2 * 128, 1024);
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
    bool first = true;
    while (walk.next(sg)) {
      const int n = sg.kb1 - sg.kb0;
      const bool tr = TRACE && trace != nullptr && blockIdx.x == 0 && first;
      first = false;
      wait_lean(q_full, qph);
      wait_lean(peer_q, qph);
      qph ^= 1;
      tc_fence_after();
      for (int b = 0; b < 2 && b < n; ++b) {
        const int e = acquire();
        issue_s(b, e);
        if (leader) {
          umma2_commit_local(&empty[e]);
          if (b + 1 == n || b == 1) {
            if (n <= 2) umma2_commit_local(q_empty);    // no further S in this segment: the Q tile may be replaced
          }
          umma2_commit_both(&s_full[b]);
        }
        __syncwarp();
      }
      for (int j = 0; j < n; ++j) {
        const int b = j & 1;
        wait_lean(&p_full[b], (b ? np1 : np0) & 1);
        if (b) ++np1; else ++np0;
        tc_fence_after();
        if (tr && leader) trace[j * kTrace + 0] = clock64();
        const int e = acquire();
        if (tr && leader) trace[j * kTrace + 8] = clock64();
        if (!(TRACE && (dbg & 2))) issue_pv(b, e, j > 0);
        if (tr && leader) trace[j * kTrace + 9] = clock64();
        if (j + 2 < n && !(TRACE && (dbg & 1))) issue_s(b, e);
        if (tr && leader) trace[j * kTrace + 10] = clock64();
        if (leader) {
          umma2_commit_local(&empty[e]);
          if (j + 2 < n && j + 3 >= n) umma2_commit_local(q_empty);   // that was the segment's last S
          if (tr) trace[j * kTrace + 11] = clock64();
          // also signalled when no further S goes into this buffer: "P(j)*V retired" is what a rescale waits for
          umma2_commit_both(&s_full[b]);
          if (j == n - 1) umma2_commit_both(o_done);
          if (tr) trace[j * kTrace + 1] = clock64();
        }
        __syncwarp();
      }
    }
  } else if (warp < 8) {
    // ===================== softmax + output (warps 0..7) =====================
    const int g = warp >> 2;                            // column group: keys [64 g, 64 g + 64) of every block
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;                  // query row inside the tile
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_o = t_lane + 256 + g * DH;
    const uint32_t t_oa = t_lane + 256 + g * 64, t_ob = t_lane + 384 + g * 64;
    const uint32_t p_remote = mapa_u32(p_full, 0);
    uint32_t ns0 = 0, ns1 = 0;                          // s_full completions accounted for per S buffer
    uint32_t oph = 0;
    bool first = true;

    while (walk.next(sg)) {
      const int bh = sg.item / sched.n_cl;
      const int qtile = 2 * (sg.item % sched.n_cl) + static_cast<int>(rank);
      const int row = qtile * BQ + r;
      const int n = sg.kb1 - sg.kb0;
      const bool trs = TRACE && trace != nullptr && blockIdx.x == 0 && first && lane == 0 && quarter == 0;
      first = false;
      float m_run = -INFINITY;                          // true running row maximum (raw scores)
      float m_used = -INFINITY;                         // maximum the current scale of P, l and O refers to
      float l = 0.f;

      bool s_ready = false;                             // the next S block was already complete when probed
      for (int j = 0; j < n; ++j) {
        const int b = j & 1;
        const int kv_valid = Tk - (sg.kb0 + j) * BK2 - g * GRP;   // valid keys of this group's 64 columns (may be <= 0)
        const uint32_t t_s = t_lane + b * BK2 + g * GRP;
        if (!s_ready) wait_lean(&s_full[b], (b ? ns1 : ns0) & 1);
        if (b) ++ns1; else ++ns0;
        if (trs) trace[j * kTrace + 2 + 3 * g] = clock64();
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
              if (e >= kv_valid) s[e] = 0xff800000u;    // -inf
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
          if (m_used == -INFINITY) {                    // the group's first block with a key
            m_used = m_run;
          } else if ((m_run - m_used) * scale_log2 > 8.0f) {
            alpha = ex2_approx((m_used - m_run) * scale_log2);
            m_used = m_run;
            need = true;
          }
          if (__any_sync(0xffffffffu, need)) {
            // O_g must be quiescent: P(j-1)*V is retired once the OTHER S buffer's next completion (S(j+1), or the
            // bare commit when no S(j+1) exists) is signalled; P(j)*V cannot start before this thread publishes P(j)
            wait_lean(&s_full[b ^ 1], (b ? ns0 : ns1) & 1);
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
          // p = exp2(s*c - m*c) on pairs with packed fp32 FMA/ADD; POLY of every 8 pairs take the polynomial exp2 on
          // the FMA pipe, the others the MUFU unit, so neither pipe alone bounds the loop
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
        if (trs) trace[j * kTrace + 3 + 3 * g] = clock64();
        // probe the next block's barrier now: the round trip (~100 clocks) overlaps the publication below
        s_ready = j + 1 < n && mbar_test_wait(&s_full[b ^ 1], (b ? ns0 : ns1) & 1);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        // the tensor-memory stores are complete (wait::st) and fenced; the arrival itself needs no memory ordering --
        // a release at cluster scope here costs ~1000 clocks per block (measured)
        if (lane == 0) {
          if (rank == 0) mbar_arrive(&p_full[b]);
          else mbar_arrive_cluster_relaxed(p_remote + 8 * b);
        }
        if (trs) trace[j * kTrace + 4 + 3 * g] = clock64();
      }
      // the bare completions of this segment (one per S buffer it used) are not waited for as S blocks; o_done below is
      // committed after them, so they have happened before the next segment's first wait
      ++ns0;
      if (n > 1) ++ns1;

      // ---- merge the two column groups; then either normalise + gate + store, or hand the part to the merger ----
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
      o += g * 64;
      wait_lean(o_done, oph);
      oph ^= 1;
      tc_fence_after();
      // both accumulators are visible to either group (same TMEM lanes): group g finishes output columns
      // [64 g, 64 g + 64) of the merged row
      xs[g * BQ + r] = make_float2(m_used, l);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float2 a = xs[r], bb = xs[BQ + r];
      // group B has seen no key at all when the segment ends inside the first 64 keys of the last block
      const float m = fmaxf(a.x, bb.x);
      const float fa0 = a.y > 0.f ? ex2_approx((a.x - m) * scale_log2) : 0.f;
      const float fb0 = bb.y > 0.f ? ex2_approx((bb.x - m) * scale_log2) : 0.f;
      const float lt = a.y * fa0 + bb.y * fb0;
      const bool direct = sg.parts == 1;
      const float f = direct ? gt / lt : 1.f;
      const float fa = fa0 * f, fb = fb0 * f;
      // scratch tile layout [32 column quads][128 rows] of float4: the lanes of a warp (consecutive rows) write and
      // read consecutive 16 B
      uint4* po = nullptr;
      if (!direct)
        po = reinterpret_cast<uint4*>(scr.part_o) + ((static_cast<int64_t>(sg.slot) * 2 + rank) * 32 + g * 16) * BQ + r;
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t va[32], vb[32];
        tmem_ld_32x32(t_oa + c, va);
        tmem_ld_32x32(t_ob + c, vb);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e)
          va[e] = __float_as_uint(__uint_as_float(va[e]) * fa + __uint_as_float(vb[e]) * fb);
        if (direct) {
          if (row < Tq) {
#pragma unroll
            for (int e = 0; e < 32; e += 8) {
              uint4 w;
              w.x = pack_bf16x2(__uint_as_float(va[e + 0]), __uint_as_float(va[e + 1]));
              w.y = pack_bf16x2(__uint_as_float(va[e + 2]), __uint_as_float(va[e + 3]));
              w.z = pack_bf16x2(__uint_as_float(va[e + 4]), __uint_as_float(va[e + 5]));
              w.w = pack_bf16x2(__uint_as_float(va[e + 6]), __uint_as_float(va[e + 7]));
              *reinterpret_cast<uint4*>(o + c + e) = w;
            }
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            __stcg(po + ((c + e) >> 2) * BQ, make_uint4(va[e], va[e + 1], va[e + 2], va[e + 3]));
        }
      }
      // the accumulators are read: the next segment's first P*V (behind this thread's next P publication) may overwrite
      tc_fence_before();
      if (direct) {
        if (g == 0 && lse_out != nullptr && row < Tq)
          lse_out[static_cast<int64_t>(bh) * Tq + row] = m * scale + logf(lt);
      } else {
        if (g == 0) __stcg(scr.part_ml + (static_cast<int64_t>(sg.slot) * 2 + rank) * BQ + r, make_float2(m, lt));
        __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        int* sem = scr.sem + sg.slot0 * 2 + rank;
        if (threadIdx.x == 0) *xflag = atomicAdd(sem, 1);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const bool last = *xflag == sg.parts - 1;
        asm volatile("bar.sync 1, 256;" ::: "memory");   // xflag may be rewritten by the next segment
        if (last) {
          // every part of this query tile is in the scratch: merge them in part order (deterministic)
          __threadfence();
          constexpr int kMaxParts = 8;                  // the host keeps a range at least nblk / 6 blocks long
          float mp[kMaxParts], wp[kMaxParts];
          float M = -INFINITY;
          const int parts = sg.parts < kMaxParts ? sg.parts : kMaxParts;
#pragma unroll
          for (int p = 0; p < kMaxParts; ++p) {
            if (p < parts) {
              const int sl = p == 0 ? sg.slot0 : 2 * (sg.c_first + p);
              const float2 ml = __ldcg(scr.part_ml + (static_cast<int64_t>(sl) * 2 + rank) * BQ + r);
              mp[p] = ml.x;
              wp[p] = ml.y;
              M = fmaxf(M, ml.x);
            }
          }
          float L = 0.f;
#pragma unroll
          for (int p = 0; p < kMaxParts; ++p) {
            if (p < parts) {
              const float w = ex2_approx((mp[p] - M) * scale_log2);
              L += wp[p] * w;
              wp[p] = w;
            }
          }
          const float fn = gt / L;
          // one part at a time: its 16 quads are loaded back to back (one L2 round trip), then accumulated
          float acc[64];
#pragma unroll
          for (int e = 0; e < 64; ++e) acc[e] = 0.f;
          for (int p = 0; p < parts; ++p) {
            const int sl = p == 0 ? sg.slot0 : 2 * (sg.c_first + p);
            const float4* src =
                reinterpret_cast<const float4*>(scr.part_o) + ((static_cast<int64_t>(sl) * 2 + rank) * 32 + g * 16) * BQ + r;
            float w = wp[0];
#pragma unroll
            for (int pp = 1; pp < kMaxParts; ++pp) w = p == pp ? wp[pp] : w;
            float4 x[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) x[e] = __ldcg(src + e * BQ);
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              acc[4 * e] = fmaf(x[e].x, w, acc[4 * e]);
              acc[4 * e + 1] = fmaf(x[e].y, w, acc[4 * e + 1]);
              acc[4 * e + 2] = fmaf(x[e].z, w, acc[4 * e + 2]);
              acc[4 * e + 3] = fmaf(x[e].w, w, acc[4 * e + 3]);
            }
          }
          if (row < Tq) {
#pragma unroll
            for (int e = 0; e < 64; e += 8) {
              uint4 w;
              w.x = pack_bf16x2(acc[e + 0] * fn, acc[e + 1] * fn);
              w.y = pack_bf16x2(acc[e + 2] * fn, acc[e + 3] * fn);
              w.z = pack_bf16x2(acc[e + 4] * fn, acc[e + 5] * fn);
              w.w = pack_bf16x2(acc[e + 6] * fn, acc[e + 7] * fn);
              *reinterpret_cast<uint4*>(o + e) = w;
            }
          }
          if (g == 0 && lse_out != nullptr && row < Tq)
            lse_out[static_cast<int64_t>(bh) * Tq + row] = M * scale + logf(L);
          if (threadIdx.x == 0) *sem = 0;               // ready for the next launch
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // nobody leaves while the peer may still touch my barriers / smem
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// scratch of the split schedule, one per (device, stream): two launches on the same stream are ordered, launches on
// different streams must not share it
struct ScratchEntry {
  Scratch2 s;
  int slots = 0;
};
std::mutex g_scratch_mu;
std::map<std::pair<int, cudaStream_t>, ScratchEntry> g_scratch;

// nullptr members when the scratch does not exist yet and cannot be created now (stream capture)
int get_scratch(cudaStream_t stream, int slots, Scratch2* out) {
  std::lock_guard<std::mutex> lock(g_scratch_mu);
  ScratchEntry& e = g_scratch[std::make_pair(current_device(), stream)];
  if (e.slots < slots) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    LTX2_CUDA_CHECK(cudaStreamIsCapturing(stream, &cap));
    if (cap != cudaStreamCaptureStatusNone) {
      *out = Scratch2{nullptr, nullptr, nullptr};
      return LTX2_OK;
    }
    if (e.slots > 0) {
      LTX2_CUDA_CHECK(cudaStreamSynchronize(stream));
      cudaFree(e.s.part_o);
      cudaFree(e.s.part_ml);
      cudaFree(e.s.sem);
      e.slots = 0;
    }
    const size_t tiles = static_cast<size_t>(slots) * 2;
    LTX2_CUDA_CHECK(cudaMalloc(&e.s.part_o, tiles * BQ * DH * sizeof(float)));
    LTX2_CUDA_CHECK(cudaMalloc(&e.s.part_ml, tiles * BQ * sizeof(float2)));
    LTX2_CUDA_CHECK(cudaMalloc(&e.s.sem, tiles * sizeof(int)));
    LTX2_CUDA_CHECK(cudaMemset(e.s.sem, 0, tiles * sizeof(int)));
    e.slots = slots;
  }
  *out = e.s;
  return LTX2_OK;
}

}  // namespace

// Where the SM-pair kernel is used.  Measured on B200 against the one-SM two-stream kernel (tools/attn2_check.py,
// profiles/r2_attn2_check.txt): both are bound by the softmax instruction stream (~10-11 scores per clock and SM,
// against the 16 the tensor pipe could take at head_dim 128), so the pair's cheaper S instructions only pay where the
// one-SM kernel has its own problem -- very long key loops (12288 x 12288: 531 vs 582 us); at 3456 keys the one-SM
// kernel keeps two DE-synchronised softmax warps per scheduler and wins (175 vs 200 us).  LTX2_ATTN_2CTA=1 forces the
// pair kernel wherever it applies (tests), =0 removes it.
bool attention_2cta_applies(const AttnV& v, int Tq, int Tk, int Dh) {
  if (!(Dh == 128 && v.rows != 0 && Tq > BQ)) return false;
  const char* env = getenv("LTX2_ATTN_2CTA");
  if (env && env[0] == '0') return false;
  if (env && env[0] == '1') return true;
  return Tq >= 8192 && Tk >= 8192;
}

void attention_2cta_plan(int Tq, int Tk, int BH, int* n_clusters, int* split) {
  const int n_q = (Tq + BQ - 1) / BQ;
  const int n_cl = (n_q + 1) / 2;
  const int nblk = (Tk + BK2 - 1) / BK2;
  const long long n_items = static_cast<long long>(BH) * n_cl;
  const long long G = n_items * nblk;
  const int pairs = num_sms() / 2;
  int C = n_items < pairs ? static_cast<int>(n_items) : pairs;
  // whole items round-robin: ceil(n_items / C) items of nblk blocks; equal ranges over all SM pairs: ceil(G / pairs)
  // blocks plus about two blocks' worth of partial epilogue + merge.  Split when that is at least 5 % shorter.
  const char* env = getenv("LTX2_ATTN_SPLIT");
  const long long whole = (n_items + C - 1) / C * nblk;
  // a range is at least 4 key blocks and a third of an item long: an item then has at most 4 parts (the merge
  // handles 8) and the per-segment costs (pipeline fill, partial epilogue, merge) stay small against its key loop
  const long long l_min = nblk / 3 > 4 ? (nblk + 2) / 3 : 4;
  long long cs = G / l_min < pairs ? G / l_min : pairs;
  if (cs < 1) cs = 1;
  const int Cs = static_cast<int>(cs);
  const long long ranged = (G + Cs - 1) / Cs + 2;
  bool sp = ranged * 105 < whole * 100;
  if (env) sp = env[0] != '0' && Cs > 1;
  if (sp) C = Cs;
  *n_clusters = C;
  *split = sp ? 1 : 0;
}

int attention_2cta_segments(int Tq, int Tk, int BH, int cluster, int* out7, int max_segments) {
  // the walk of one cluster through the work decomposition, exactly as the kernel does it (pure integers)
  Sched2 sd;
  const int n_q = (Tq + BQ - 1) / BQ;
  sd.n_cl = (n_q + 1) / 2;
  sd.n_items = BH * sd.n_cl;
  sd.nblk = (Tk + BK2 - 1) / BK2;
  attention_2cta_plan(Tq, Tk, BH, &sd.C, &sd.split);
  if (cluster < 0 || cluster >= sd.C) return -1;
  Walker w(sd, cluster);
  Seg sg;
  int n = 0;
  while (w.next(sg)) {
    if (n < max_segments) {
      int* o = out7 + 7 * n;
      o[0] = sg.item; o[1] = sg.kb0; o[2] = sg.kb1; o[3] = sg.part; o[4] = sg.parts;
      o[5] = sg.parts > 1 ? sg.slot : -1;
      o[6] = sg.parts > 1 ? sg.slot0 : -1;
    }
    ++n;
  }
  return n;
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
  Sched2 sd;
  const int n_q = (Tq + BQ - 1) / BQ;
  sd.n_cl = (n_q + 1) / 2;
  sd.n_items = static_cast<int>(BH) * sd.n_cl;
  sd.nblk = (Tk + BK2 - 1) / BK2;
  attention_2cta_plan(Tq, Tk, static_cast<int>(BH), &sd.C, &sd.split);
  Scratch2 scr{nullptr, nullptr, nullptr};
  if (sd.split) {
    LTX2_PROPAGATE(get_scratch(stream, 2 * (num_sms() / 2), &scr));
    if (scr.part_o == nullptr) {                        // capturing before the first eager call: whole items
      sd.split = 0;
      const int pairs = num_sms() / 2;
      sd.C = sd.n_items < pairs ? sd.n_items : pairs;
    }
  }
  const unsigned grid = 2u * static_cast<unsigned>(sd.C);
  const float kLog2e = 1.4426950408889634f;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
#define LTX2_LAUNCH_2CTA(P, T)                                                                                   \
  LTX2_CUDA_CHECK(launch_pdl(attention_2cta_kernel<P, T>, dim3(grid), dim3(kThreads), kSmem, stream, mq, mk, mv, o, H, \
                             Tq, Tk, scale * kLog2e, scale, gate_logits, lse_out, trace, sc, sd, scr, dbg))
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
