// Flash-style attention on tcgen05 tensor cores, head_dim 128, TWO softmax streams per CTA:
//   O = softmax(Q K^T / sqrt(d)) V, non-causal, no mask.
//
// Reference op replaced: mx.fast.scaled_dot_product_attention as called from
// _compiled_attention_core_no_mask (attention.py:12-34), plus the head merge
// (B,H,T,D)->(B,T,H*D) (:34) and the V2 per-head gate 2*sigmoid(logits) (:243-250).
//
// Why two streams: with one 128-query tile per CTA every 128x128 key block needs its own 64 KB of K/V from L2 and
// the softmax of block j sits alone on its SM sub-partition (one warp per scheduler: measured IPC 0.4).  Here a CTA
// owns two streams that share the tensor pipe and put two independent softmax warps on every scheduler:
//   pair mode   streams = two adjacent 128-query tiles of one head over ALL keys; every K/V tile fetched from L2 is
//               used by both (half the L2->SM traffic per FLOP);
//   split mode  streams = ONE query tile over the first / second half of the keys, merged in the CTA at the end
//               (used for the odd tile of a head and for small grids such as the context-parallel head shards).
// Keys are processed in SUB-BLOCKS of 64 (half a K/V tile) and each stream's S region is two 64-column buffers, so
// S(k+2) is computed while the softmax still works on S(k+1): in steady state a softmax warp never waits for the
// tensor pipe and the tensor pipe never waits for one particular softmax (tools/attn_bench.py prints the timeline).
//
// Warps (384 threads): 0-3 = softmax of stream 0, 4-7 = softmax of stream 1 (one thread per query row: no
// cross-thread reductions), 8 = TMA producer, 9 / 10 = MMA issuers of stream 0 / 1 (the scalar work around a group
// of tcgen05.mma costs about as much as the MMAs, so one issuing warp for both streams was the bottleneck), 11 idle.
// Tensor memory (512 columns):
//   [0,64) [64,128) S buffers a, b of stream 0   [128,192) [192,256) the same for stream 1   [256,384) O0   [384,512) O1
// P(k) (bf16, 32 columns) overwrites the first half of its S buffer once the row has been pulled into registers;
// S(k+2) is issued behind P(k)*V on the in-order tensor pipe, so the overwrite is safe.  Q lives in shared memory
// (SS-mode MMA for S: M 128, N 64 -- shared-memory bound at 48 instead of 32 clk per instruction, tools/probe), P in
// tensor memory (TS-mode MMA for P*V).  K/V tiles (32 KB each) flow through one 5-stage ring.
// The running output is rescaled lazily (only when the row maximum grew by more than 2^8), by the softmax thread
// itself, after waiting for the completion that proves P(k-1)*V has retired.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "kernels.h"

namespace ltx2 {

namespace {

constexpr int kPairThreads = 384;
constexpr int kPairStages = 5;
constexpr int DHP = 128;                    // head dim
constexpr int BQP = 128;                    // queries per stream
constexpr int BKP = 128;                    // keys per K/V tile
constexpr int SUB = 64;                     // keys per softmax sub-block (half a tile)
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

// Everything the TMA producer and the MMA issuer need to walk the tile schedule.
struct PipeCtx {
  const CUtensorMap* tmap_q;
  const CUtensorMap* tmap_k;
  const CUtensorMap* tmap_v;
  uint8_t* sQ;
  uint8_t* sRing;
  uint64_t* q_full;
  uint64_t* full;
  uint64_t* empty;
  uint64_t* s_full;
  uint64_t* p_full;
  uint64_t* o_done;
  uint32_t tmem_base;
  int nsub0, nsub1, kv1, qa, bh, H;
  long long* trace;   // null unless this CTA is traced
};

// mbarrier wait without the printf watchdog of common.cuh (keeps the issuing warps' loops lean); still traps on a hang
__device__ __forceinline__ void mbar_wait_lean(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity))
    if (++spins > (1u << 22)) __trap();
}

// The tile schedule, walked with identical control flow by the TMA producer (ROLE 0) and by the two MMA issuers
// (ROLE 1: stream 0, ROLE 2: stream 1 -- one issuing warp per stream, because the scalar work around a group of
// tcgen05.mma (barrier waits, descriptors, commits) costs about as many cycles as the MMAs themselves).
// Keys are processed in sub-blocks of 64 (half a K/V tile).  Per tile step t, in the fixed order
// (stream 0, half a), (1, a), (0, b), (1, b):   O_i += P_i(2t+h) * V_t[half h];   S_i(2t+2+h) = Q_i K_{t+1}[half h]^T
// into the S buffer the softmax just drained.  A tile is fetched at its first use and released after its last.
// PAIR: both streams use the same K/V tiles (each MMA warp waits for a tile at its own first use and commits to the
// tile's `empty` barrier, which then counts two arrivals); otherwise every stream has its own tiles and an MMA warp
// only steps over the ring slots of the other stream.  Everything is compile-time per (stream, half) so that the
// tcgen05.mma operands stay in uniform registers.
template <bool VROWS, bool PAIR, int ROLE>
__device__ __forceinline__ void walk_schedule(const PipeCtx& c) {
  constexpr bool PRODUCER = ROLE == 0;
  constexpr int ME = ROLE - 1;
  const bool leader = elect_one();
  if (PRODUCER && leader) {
    mbar_expect_tx(c.q_full, PAIR ? 2 * kTileBytes : kTileBytes);
#pragma unroll
    for (int t = 0; t < (PAIR ? 2 : 1); ++t)
#pragma unroll
      for (int cc = 0; cc < DHP / 64; ++cc)
        tma_load_3d(c.sQ + t * kTileBytes + cc * (BQP * 128), c.tmap_q, c.q_full, cc * 64, (c.qa + t) * BQP, c.bh);
  }
  constexpr uint32_t idesc_s = umma_idesc_bf16(BQP, SUB);
  constexpr uint32_t idesc_o = umma_idesc_bf16(BQP, DHP, VROWS);
  const int nsub0 = c.nsub0, nsub1 = c.nsub1;
  int slot = 0;
  uint32_t ph = 0;
  // tile bookkeeping per owner (owner 0 = stream 0, or the shared tiles in PAIR mode; owner 1 = stream 1)
  int k_t0 = -1, k_t1 = -1, v_t0 = -1, v_t1 = -1;
  int k_slot0 = 0, k_slot1 = 0, v_slot0 = 0, v_slot1 = 0;

  // next ring slot: the producer fills it with K (is_v = false) or V tile `tile`; an MMA warp waits for it
  auto acquire = [&](bool is_v, int tile) -> int {
    const int s = slot;
    if (PRODUCER) {
      mbar_wait_lean(&c.empty[s], ph ^ 1);
      if (leader) {
        uint8_t* dst = c.sRing + s * kTileBytes;
        mbar_expect_tx(&c.full[s], kTileBytes);
        if (!is_v) {
#pragma unroll
          for (int cc = 0; cc < DHP / 64; ++cc)
            tma_load_3d(dst + cc * (BKP * 128), c.tmap_k, &c.full[s], cc * 64, tile * BKP, c.bh);
        } else if (VROWS) {   // V rows [keys, d]: one 128-key x 64-channel box per 64 channels (MN-major B operand)
#pragma unroll
          for (int cc = 0; cc < DHP / 64; ++cc)
            tma_load_4d(dst + cc * (BKP * 128), c.tmap_v, &c.full[s], cc * 64, tile * BKP, c.bh % c.H, c.bh / c.H);
        } else {              // V^T [d, keys]: one DH x 64-key box per 64 keys (K-major B operand)
#pragma unroll
          for (int cc = 0; cc < BKP / 64; ++cc)
            tma_load_3d(dst + cc * (DHP * 128), c.tmap_v, &c.full[s], tile * BKP + cc * 64, 0, c.bh);
        }
      }
    } else {
      mbar_wait_lean(&c.full[s], ph);
      tc_fence_after();
    }
    if (++slot == kPairStages) { slot = 0; ph ^= 1; }
    return s;
  };
  auto skip = [&]() -> int {   // a ring slot that belongs to the other stream's MMA warp
    const int s = slot;
    if (++slot == kPairStages) { slot = 0; ph ^= 1; }
    return s;
  };
  auto release = [&](int s) {
    if (!PRODUCER && leader) umma_commit(&c.empty[s]);
  };
  // S_i[half h] = Q_i K[64h .. 64h+63]^T  (M 128, N 64, K 128) into S buffer h of stream i
  auto issue_s = [&](auto I_, auto H_, int k_slot) {
    constexpr int i = decltype(I_)::value, h = decltype(H_)::value;
    const uint64_t qd = umma_desc_k_sw128(smem_u32(c.sQ + ((PAIR && i) ? kTileBytes : 0)));
    const uint64_t kd = umma_desc_k_sw128(smem_u32(c.sRing + k_slot * kTileBytes) + h * (SUB * 128));
    if (leader) {
#pragma unroll
      for (int ks = 0; ks < DHP / 16; ++ks) {
        const uint64_t off = ((ks / 4) * (128 * 128) >> 4) + 2 * (ks % 4);
        umma_bf16_ss(c.tmem_base + i * BKP + h * SUB, qd + off, kd + off, idesc_s, ks != 0);
      }
    }
  };
  // O_i += P_i[half h] V[64h .. 64h+63]  (M 128, N 128, K 64); P sits in the first 32 columns of S buffer h
  auto issue_pv = [&](auto I_, auto H_, int v_slot, bool acc) {
    constexpr int i = decltype(I_)::value, h = decltype(H_)::value;
    const uint32_t v_addr = smem_u32(c.sRing + v_slot * kTileBytes);
    const uint64_t vd = VROWS ? umma_desc_mn_sw128(v_addr, BKP * 128, 1024) : umma_desc_k_sw128(v_addr);
    if (leader) {
#pragma unroll
      for (int ks = 0; ks < SUB / 16; ++ks) {
        constexpr int kk0 = h * (SUB / 16);
        const int kk = kk0 + ks;
        umma_bf16_ts(c.tmem_base + 256 + i * DHP, c.tmem_base + i * BKP + h * SUB + ks * 8,
                     vd + (VROWS ? kk * (2048 >> 4) : ((kk / 4) * (DHP * 128) >> 4) + 2 * (kk % 4)), idesc_o,
                     acc || ks != 0);
      }
    }
  };
  auto has = [&](int i, int k) { return k < (i == 0 ? nsub0 : nsub1); };
  using I0 = std::integral_constant<int, 0>;
  using I1 = std::integral_constant<int, 1>;

  if (!PRODUCER) {
    mbar_wait_lean(c.q_full, 0);
    tc_fence_after();
  }
  // ---- prologue: the S of tile 0 ----
  auto pro = [&](auto I_, auto H_) {
    constexpr int i = decltype(I_)::value, h = decltype(H_)::value;
    constexpr bool mine = PRODUCER || i == ME;
    if (PAIR && !mine) return;                             // shared tiles: I fetch them at my own first use
    constexpr bool own1 = !PAIR && i == 1;                 // which owner's tile this stream reads
    if (!has(i, h)) return;
    int& k_t = own1 ? k_t1 : k_t0;
    int& k_slot = own1 ? k_slot1 : k_slot0;
    if (k_t != 0) {
      k_slot = mine ? acquire(false, own1 ? c.kv1 : 0) : skip();
      k_t = 0;
    }
    if (PRODUCER || !mine) return;
    issue_s(I_, H_, k_slot);
    if (leader) umma_commit(&c.s_full[i * 2 + h]);
    if (h == 1 || !has(i, 1)) release(k_slot);
  };
  pro(I0{}, I0{});
  pro(I1{}, I0{});
  pro(I0{}, I1{});
  pro(I1{}, I1{});
  __syncwarp();
  // ---- main loop over tile steps ----
  const int nt = (nsub0 + 1) >> 1;                         // stream 0 never has fewer sub-blocks than stream 1
  for (int t = 0; t < nt; ++t) {
    auto step = [&](auto I_, auto H_) {
      constexpr int i = decltype(I_)::value, h = decltype(H_)::value;
      constexpr bool mine = PRODUCER || i == ME;
      if (PAIR && !mine) return;
      constexpr bool own1 = !PAIR && i == 1;
      const int k = 2 * t + h;
      if (!has(i, k)) return;
      int& k_t = own1 ? k_t1 : k_t0;
      int& k_slot = own1 ? k_slot1 : k_slot0;
      int& v_t = own1 ? v_t1 : v_t0;
      int& v_slot = own1 ? v_slot1 : v_slot0;
      const int tile0 = own1 ? c.kv1 : 0;
      if (!PRODUCER && mine) {
        mbar_wait_lean(&c.p_full[i * 2 + h], t & 1);
        tc_fence_after();
        if (c.trace && leader && h == 0) c.trace[t * kTraceStride + 2 * i] = clock64();
      }
      if (v_t != t) {
        v_slot = mine ? acquire(true, tile0 + t) : skip();
        v_t = t;
      }
      if (!PRODUCER && mine) {
        issue_pv(I_, H_, v_slot, k > 0);
        if (h == 1 || !has(i, 2 * t + 1)) release(v_slot);
      }
      if (has(i, k + 2)) {
        if (k_t != t + 1) {
          k_slot = mine ? acquire(false, tile0 + t + 1) : skip();
          k_t = t + 1;
        }
        if (!PRODUCER && mine) {
          issue_s(I_, H_, k_slot);
          if (h == 1 || !has(i, 2 * t + 3)) release(k_slot);
        }
      }
      if (!PRODUCER && mine) {
        if (leader) {
          // also signalled when no further S goes into this buffer: "P_i(k)*V retired" is what a rescale waits for
          umma_commit(&c.s_full[i * 2 + h]);
          if (k == (i == 0 ? nsub0 : nsub1) - 1) umma_commit(&c.o_done[i]);
          if (c.trace && h == 0) c.trace[t * kTraceStride + 2 * i + 1] = clock64();
        }
        __syncwarp();
      }
    };
    step(I0{}, I0{});
    step(I1{}, I0{});
    step(I0{}, I1{});
    step(I1{}, I1{});
  }
}

template <bool VROWS, int POLY>
__global__ void __launch_bounds__(kPairThreads, 1)
attention_pair_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                      const __grid_constant__ CUtensorMap tmap_v, __nv_bfloat16* __restrict__ out, int H, int Tq, int Tk,
                      float scale_log2, float scale, const float* __restrict__ gate_logits,
                      float* __restrict__ lse_out, long long* __restrict__ trace,
                      const __grid_constant__ AttnOutScatter sc, PairGrid pg) {
  pdl_trigger();                                        // the next kernel may become resident under my tail
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // [2 tiles][2 x (128 rows x 128 B)]
  uint8_t* sRing = sQ + 2 * kTileBytes;                 // [stages][32 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRing + kPairStages * kTileBytes);
  uint64_t* q_full = bars;                              // 1
  uint64_t* full = bars + 1;                            // stages   TMA -> MMA
  uint64_t* empty = full + kPairStages;                 // stages   MMA -> TMA
  uint64_t* s_full = empty + kPairStages;               // [stream][half]  MMA -> softmax: S sub-block in TMEM buffer `half`
  uint64_t* p_full = s_full + 4;                        // [stream][half]  softmax -> MMA: P sub-block in TMEM (O rescaled)
  uint64_t* o_done = p_full + 4;                        // [stream]        MMA -> softmax: last P*V retired
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
  const int keys0 = pair ? Tk : min(n0 * BKP, Tk);      // keys of stream 0 / 1, processed in sub-blocks of 64
  const int keys1 = pair ? Tk : Tk - n0 * BKP;
  const int nsub0 = (keys0 + SUB - 1) / SUB;
  const int nsub1 = keys1 > 0 ? (keys1 + SUB - 1) / SUB : 0;
  const bool tr = trace != nullptr && blockIdx.x == 0;
  if (tr && threadIdx.x == 0) trace[nkv * kTraceStride + 0] = clock64();           // kernel entry

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < kPairStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], pair ? 2 : 1);               // pair items: both MMA warps release a shared tile
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
    }
    mbar_init(&o_done[0], 1);
    mbar_init(&o_done[1], 1);
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
  pdl_wait();                                           // set-up above overlapped the previous kernel's tail

  if (warp >= 8) {
    if (warp < 11) {
      // ============ TMA producer (warp 8), MMA issuers (warp 9: stream 0, warp 10: stream 1) ============
      PipeCtx c;
      c.tmap_q = &tmap_q; c.tmap_k = &tmap_k; c.tmap_v = &tmap_v;
      c.sQ = sQ; c.sRing = sRing;
      c.q_full = q_full; c.full = full; c.empty = empty; c.s_full = s_full; c.p_full = p_full; c.o_done = o_done;
      c.tmem_base = tmem_base;
      c.nsub0 = nsub0; c.nsub1 = nsub1; c.kv1 = kv1; c.qa = qa; c.bh = bh; c.H = H;
      c.trace = tr ? trace : nullptr;
      if (warp == 8) {
        if (pair) walk_schedule<VROWS, true, 0>(c);
        else walk_schedule<VROWS, false, 0>(c);
      } else if (warp == 9) {
        if (pair) walk_schedule<VROWS, true, 1>(c);
        else walk_schedule<VROWS, false, 1>(c);
      } else {
        if (pair) walk_schedule<VROWS, true, 2>(c);
        else walk_schedule<VROWS, false, 2>(c);
      }
    }
  } else {
    // ===================== softmax + output (warps 0..7) =====================
    const int i = warp >> 2;                            // stream
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;                  // query row inside the tile
    const int qtile = pair ? qa + i : qa;
    const int row = qtile * BQP + r;
    const int nsub = i == 0 ? nsub0 : nsub1;
    const int keys = i == 0 ? keys0 : keys1;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_s = t_lane + i * BKP;
    const uint32_t t_o = t_lane + 256 + i * DHP;
    const bool trs = tr && lane == 0 && quarter == 0;   // warps 0 and 4
    float m_run = -INFINITY;                            // true running row maximum (raw scores)
    float m_used = -INFINITY;                           // maximum the current scale of P, l and O refers to
    float l = 0.f;

    if (trs && i == 0) trace[nkv * kTraceStride + 1] = clock64();   // set-up done (TMEM, barriers)
    bool s_ready = false;                               // the next S sub-block was already complete when probed
    for (int k = 0; k < nsub; ++k) {
      const int h = k & 1, t = k >> 1;
      const int kv_valid = keys - k * SUB;
      const int tro = i == 0 ? 4 + 4 * h : 12;
      const bool trk = trs && (i == 0 || h == 0);
      if (!s_ready) mbar_wait_lean(&s_full[i * 2 + h], t & 1);
      if (trk) trace[t * kTraceStride + tro + 0] = clock64();
      tc_fence_after();
      uint32_t s[SUB];
      {
        uint32_t (&s0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[0]);
        uint32_t (&s1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[32]);
        tmem_ld_32x32(t_s + h * SUB + 0, s0);
        tmem_ld_32x32(t_s + h * SUB + 32, s1);
        tmem_ld_wait();
      }
      if (trk) trace[t * kTraceStride + tro + 1] = clock64();
      if (POLY == 8) {   // diagnostics: no softmax arithmetic -> the tensor-pipe + hand-off floor of this structure
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) pk[e] = s[e] & 0x3f803f80u;
#pragma unroll
        for (int c = 0; c < 2; ++c) tmem_st_32x16(t_s + h * SUB + c * 16, pk);
        l = 1.f;
        m_used = 0.f;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[i * 2 + h]);
        s_ready = false;
        continue;
      }
      if (kv_valid < SUB) {
#pragma unroll
        for (int e = 0; e < SUB; ++e)
          if (e >= kv_valid) s[e] = 0xff800000u;        // -inf
      }
      float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[1]), mx2 = __uint_as_float(s[2]),
            mx3 = __uint_as_float(s[3]);
#pragma unroll
      for (int e = 4; e < SUB; e += 4) {
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
        // O_i must be quiescent: P(k-1)*V is retired once the OTHER S buffer's next completion (S(k+1), or the
        // bare commit when no S(k+1) exists) is signalled; P(k)*V cannot start before this thread publishes P(k)
        mbar_wait_lean(&s_full[i * 2 + (h ^ 1)], ((k + 1) >> 1) & 1);
        tc_fence_after();
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
        tmem_st_32x16(t_s + h * SUB + c * 16, pk);
      }
      l = l * alpha + ((acc0.x + acc0.y) + (acc1.x + acc1.y));
      if (trk) trace[t * kTraceStride + tro + 2] = clock64();
      // probe the next sub-block's barrier now: the round trip (~100 clocks) overlaps the publication below
      s_ready = k + 1 < nsub && mbar_test_wait(&s_full[i * 2 + (h ^ 1)], ((k + 1) >> 1) & 1);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[i * 2 + h]);
      if (trk) trace[t * kTraceStride + tro + 3] = clock64();
    }

    if (trs && i == 0) trace[nkv * kTraceStride + 2] = clock64();   // key loop done
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
      mbar_wait_lean(&o_done[i], 0);
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
      mbar_wait_lean(&o_done[0], 0);
      if (n1 > 0) mbar_wait_lean(&o_done[1], 0);
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
    if (trs && i == 0) trace[nkv * kTraceStride + 3] = clock64();   // epilogue done
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
  static PerDeviceOnce configured;
  if (configured.first()) {
    LTX2_CUDA_CHECK(cudaFuncSetAttribute(attention_pair_kernel<VROWS, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem));
    LTX2_CUDA_CHECK(cudaFuncSetAttribute(attention_pair_kernel<VROWS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem));
    LTX2_CUDA_CHECK(cudaFuncSetAttribute(attention_pair_kernel<VROWS, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem));
    LTX2_CUDA_CHECK(cudaFuncSetAttribute(attention_pair_kernel<VROWS, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem));
    LTX2_CUDA_CHECK(cudaFuncSetAttribute(attention_pair_kernel<VROWS, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem));
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
  PairGrid pg;
  pg.n_q = (Tq + BQP - 1) / BQP;
  pg.BH = static_cast<int>(BH);
  int best_np = attention_pair_items(Tq, pg.BH, nullptr);
  if (force_pairs == -1) best_np = pg.n_q / 2;
  if (force_pairs >= 0 && force_pairs <= pg.n_q / 2) best_np = force_pairs;
  pg.n_pairs = best_np;
  const unsigned grid = static_cast<unsigned>(pg.BH * (pg.n_q - pg.n_pairs));
  const float kLog2e = 1.4426950408889634f;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
#define LTX2_LAUNCH_PAIR(P)                                                                                        \
  launch_pdl(attention_pair_kernel<VROWS, P>, dim3(grid), dim3(kPairThreads), kPairSmem, stream, mq, mk, mv, o, H, Tq, Tk, \
             scale * kLog2e, scale, gate_logits, lse_out, trace, sc, pg)
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

int attention_pair_items(int Tq, int BH, int* n_ctas) {
  // how many tiles of each (batch, head) slice run as pairs: minimal makespan over the SMs, ties -> more pairs (less
  // L2 traffic)
  const int n_q = (Tq + BQP - 1) / BQP;
  const int nsm = num_sms();
  int best_np = 0, best_span = 1 << 30;
  for (int np = n_q / 2; np >= 0; --np) {
    const int span = makespan(BH * np, BH * (n_q - 2 * np), nsm);
    if (span < best_span) {
      best_span = span;
      best_np = np;
    }
  }
  if (n_ctas != nullptr) *n_ctas = BH * (n_q - best_np);
  return best_np;
}

int attention_pair_bf16(const void* q, const void* k, const AttnV& v, void* out, int B, int H, int Tq, int Tk,
                        float scale, const float* gate_logits, float* lse_out, cudaStream_t stream, long long* trace,
                        const AttnOutScatter& sc) {
  return v.rows ? launch_pair<true>(q, k, v, out, B, H, Tq, Tk, scale, gate_logits, lse_out, trace, sc, stream)
                : launch_pair<false>(q, k, v, out, B, H, Tq, Tk, scale, gate_logits, lse_out, trace, sc, stream);
}

}  // namespace ltx2
