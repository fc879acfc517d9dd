// HBM-bound row / elementwise kernels of the DiT path.  Each is one pass over its operands with
// 16-byte vector accesses; the roofline for all of them is HBM bandwidth.
//
// Reference ops replaced:
//   norm_modulate    _compiled_adaln_forward (transformer.py:16-31), rms_norm (attention.py:88-100),
//                    LayerNorm + modulate of the output head (model.py:744-758), V2 KV modulation (transformer.py:452)
//   headnorm_rope    q_norm/k_norm (attention.py:186-187,231-232) + apply_split_rotary_emb (rope.py:92-144)
//                    + the (B,T,H*D)->(B,H,T,D) transpose (attention.py:25-27)
//   v_transpose      the V half of that transpose, emitted K-major for the P*V MMA
//   small_linear / timestep_sinusoid / build_modulation
//                    AdaLayerNormSingle (timestep_embedding.py:166-202) + get_ada_values (transformer.py:369-392)
//   rope_tables      precompute_freqs_cis, SPLIT (rope.py:365-418)
//   x0_from_velocity X0Model.denoise (model.py:912-918)
//   silu_mul / gelu_mul / interleaved_rope   the three Metal kernels (kernels/fused_ops.py)
#include "common.cuh"
#include "kernels.h"

#include <cuda_fp16.h>

namespace ltx2 {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of up to two values; every thread gets the result
template <int THREADS>
__device__ __forceinline__ float2 block_sum2(float a, float b) {
  __shared__ float sa[THREADS / 32], sb[THREADS / 32];
  a = warp_sum(a);
  b = warp_sum(b);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sa[w] = a; sb[w] = b; }
  __syncthreads();
  float ra = 0.f, rb = 0.f;
#pragma unroll
  for (int i = 0; i < THREADS / 32; ++i) { ra += sa[i]; rb += sb[i]; }
  __syncthreads();
  return make_float2(ra, rb);
}

__device__ __forceinline__ void load8_bf16(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void store8_bf16(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void load8_f32(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// ---------------------------------------------------------------------------------
// norm_modulate: one CTA per row, 8 elements per thread per step, row kept in registers
// ---------------------------------------------------------------------------------
constexpr int kRowThreads = 256;
constexpr int kMaxUnits = 4;   // D <= 256 * 8 * 4 = 8192

template <bool X_BF16>
__global__ void __launch_bounds__(kRowThreads)
norm_modulate_kernel(const void* __restrict__ x_, int64_t ldx, __nv_bfloat16* __restrict__ out, int64_t ldo, int D,
                     int norm_kind, float eps, const float* __restrict__ mod, int64_t mod_stride, int64_t shift_off,
                     int64_t scale_off, const int* __restrict__ row_cls) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x;
  const int units = D / 8;
  float v[kMaxUnits][8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int u = 0; u < kMaxUnits; ++u) {
    const int idx = threadIdx.x + u * kRowThreads;
    if (idx < units) {
      if (X_BF16) load8_bf16(reinterpret_cast<const __nv_bfloat16*>(x_) + row * ldx + idx * 8, v[u]);
      else load8_f32(reinterpret_cast<const float*>(x_) + row * ldx + idx * 8, v[u]);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s1 += v[u][i]; s2 += v[u][i] * v[u][i]; }
    }
  }
  float mean = 0.f, rstd = 1.f;
  if (norm_kind != NORM_NONE) {
    const float2 s = block_sum2<kRowThreads>(s1, s2);
    if (norm_kind == NORM_LAYER) {
      mean = s.x / D;
      // two-pass variance from registers (matches nn.LayerNorm numerics better than E[x^2]-E[x]^2)
      float d2 = 0.f;
#pragma unroll
      for (int u = 0; u < kMaxUnits; ++u) {
        const int idx = threadIdx.x + u * kRowThreads;
        if (idx < units) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { const float d = v[u][i] - mean; d2 += d * d; }
        }
      }
      const float2 t = block_sum2<kRowThreads>(d2, 0.f);
      rstd = rsqrtf(t.x / D + eps);
    } else {
      rstd = rsqrtf(s.y / D + eps);
    }
  }
  const float* mrow = nullptr;
  if (mod != nullptr) mrow = mod + static_cast<int64_t>(row_cls ? row_cls[row] : 0) * mod_stride;
#pragma unroll
  for (int u = 0; u < kMaxUnits; ++u) {
    const int idx = threadIdx.x + u * kRowThreads;
    if (idx < units) {
      float o[8];
      if (mrow != nullptr) {
        float sh[8], sc[8];
        load8_f32(mrow + shift_off + idx * 8, sh);
        load8_f32(mrow + scale_off + idx * 8, sc);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (v[u][i] - mean) * rstd * (1.f + sc[i]) + sh[i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (v[u][i] - mean) * rstd;
      }
      store8_bf16(out + row * ldo + idx * 8, o);
    }
  }
}

// ---------------------------------------------------------------------------------
// headnorm_rope: one CTA per token row
// ---------------------------------------------------------------------------------
constexpr int kHeadUnits = 2;  // inner <= 256 * 16 * 2 = 8192

__global__ void __launch_bounds__(kRowThreads)
headnorm_rope_kernel(const __nv_bfloat16* __restrict__ in, int64_t ld, const float* __restrict__ weight,
                     const float* __restrict__ cosb, const float* __restrict__ sinb, __nv_bfloat16* __restrict__ out,
                     int T, int H, int Dh, float eps) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x;             // b*T + t
  const int b = row / T, t = row % T;
  const int inner = H * Dh, half = Dh / 2;
  const int units = inner / 16;           // a unit = 8 first-half + the 8 matching second-half elements
  const int upr = half / 8;               // units per head
  float x1[kHeadUnits][8], x2[kHeadUnits][8];
  float ss = 0.f;
#pragma unroll
  for (int u = 0; u < kHeadUnits; ++u) {
    const int idx = threadIdx.x + u * kRowThreads;
    if (idx < units) {
      const int h = idx / upr, j0 = (idx % upr) * 8;
      const __nv_bfloat16* p = in + row * ld + h * Dh + j0;
      load8_bf16(p, x1[u]);
      load8_bf16(p + half, x2[u]);
#pragma unroll
      for (int i = 0; i < 8; ++i) ss += x1[u][i] * x1[u][i] + x2[u][i] * x2[u][i];
    }
  }
  const float2 s = block_sum2<kRowThreads>(ss, 0.f);
  const float rstd = rsqrtf(s.x / inner + eps);
#pragma unroll
  for (int u = 0; u < kHeadUnits; ++u) {
    const int idx = threadIdx.x + u * kRowThreads;
    if (idx < units) {
      const int h = idx / upr, j0 = (idx % upr) * 8;
      float w1[8], w2[8], o1[8], o2[8];
      load8_f32(weight + h * Dh + j0, w1);
      load8_f32(weight + h * Dh + half + j0, w2);
#pragma unroll
      for (int i = 0; i < 8; ++i) { x1[u][i] *= rstd * w1[i]; x2[u][i] *= rstd * w2[i]; }
      if (cosb != nullptr) {
        float c[8], sn[8];
        const int64_t off = static_cast<int64_t>(row) * (inner / 2) + h * half + j0;
        load8_f32(cosb + off, c);
        load8_f32(sinb + off, sn);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          o1[i] = x1[u][i] * c[i] - x2[u][i] * sn[i];
          o2[i] = x2[u][i] * c[i] + x1[u][i] * sn[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) { o1[i] = x1[u][i]; o2[i] = x2[u][i]; }
      }
      __nv_bfloat16* q = out + ((static_cast<int64_t>(b) * H + h) * T + t) * Dh + j0;
      store8_bf16(q, o1);
      store8_bf16(q + half, o2);
    }
  }
}

// ---------------------------------------------------------------------------------
// qkv_head_scatter: one CTA per token row of the fused QKV output; q/k norm + RoPE, v copy, per-head destinations
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRowThreads)
qkv_head_scatter_kernel(const __nv_bfloat16* __restrict__ qkv, int64_t ld, const float* __restrict__ wq,
                        const float* __restrict__ wk, const float* __restrict__ cosb, const float* __restrict__ sinb,
                        HeadScatter dst, int T, int H, int Dh, float eps) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x;             // b*T + t (local tokens)
  const int b = row / T, t = row % T;
  const int inner = H * Dh, half = Dh / 2;
  const int units = inner / 16;
  const int upr = half / 8;
  const __nv_bfloat16* base = qkv + row * ld;
  float q1[kHeadUnits][8], q2[kHeadUnits][8], k1[kHeadUnits][8], k2[kHeadUnits][8];
  float sq = 0.f, sk = 0.f;
#pragma unroll
  for (int u = 0; u < kHeadUnits; ++u) {
    const int idx = threadIdx.x + u * kRowThreads;
    if (idx < units) {
      const int h = idx / upr, j0 = (idx % upr) * 8;
      const __nv_bfloat16* p = base + h * Dh + j0;
      load8_bf16(p, q1[u]);
      load8_bf16(p + half, q2[u]);
      load8_bf16(p + inner, k1[u]);
      load8_bf16(p + inner + half, k2[u]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sq += q1[u][i] * q1[u][i] + q2[u][i] * q2[u][i];
        sk += k1[u][i] * k1[u][i] + k2[u][i] * k2[u][i];
      }
    }
  }
  // the RoPE rows (HBM) are fetched BEFORE the block reduction so their latency overlaps it (ncu: 42.7 -> 38.5 us;
  // the same change in norm_modulate / headnorm_rope cost occupancy and was slower, so it lives only here)
  float cs[kHeadUnits][8], sns[kHeadUnits][8];
#pragma unroll
  for (int u = 0; u < kHeadUnits; ++u) {
    const int idx = threadIdx.x + u * kRowThreads;
    if (idx < units) {
      const int64_t off = static_cast<int64_t>(row) * (inner / 2) + (idx / upr) * half + (idx % upr) * 8;
      load8_f32(cosb + off, cs[u]);
      load8_f32(sinb + off, sns[u]);
    }
  }
  const float2 s = block_sum2<kRowThreads>(sq, sk);
  const float rq = rsqrtf(s.x / inner + eps), rk = rsqrtf(s.y / inner + eps);
#pragma unroll
  for (int u = 0; u < kHeadUnits; ++u) {
    const int idx = threadIdx.x + u * kRowThreads;
    if (idx < units) {
      const int h = idx / upr, j0 = (idx % upr) * 8;
      float a1[8], a2[8], o1[8], o2[8];
      const float (&c)[8] = cs[u];
      const float (&sn)[8] = sns[u];
      const int dr = h / dst.heads_per_rank, hl = h % dst.heads_per_rank;
      const int64_t o = ((static_cast<int64_t>(b) * dst.heads_per_rank + hl) * dst.n_total + dst.t_offset + t) * Dh + j0;
      load8_f32(wq + h * Dh + j0, a1);
      load8_f32(wq + h * Dh + half + j0, a2);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x1 = q1[u][i] * rq * a1[i], x2 = q2[u][i] * rq * a2[i];
        o1[i] = x1 * c[i] - x2 * sn[i];
        o2[i] = x2 * c[i] + x1 * sn[i];
      }
      store8_bf16(dst.q[dr] + o, o1);
      store8_bf16(dst.q[dr] + o + half, o2);
      load8_f32(wk + h * Dh + j0, a1);
      load8_f32(wk + h * Dh + half + j0, a2);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x1 = k1[u][i] * rk * a1[i], x2 = k2[u][i] * rk * a2[i];
        o1[i] = x1 * c[i] - x2 * sn[i];
        o2[i] = x2 * c[i] + x1 * sn[i];
      }
      store8_bf16(dst.k[dr] + o, o1);
      store8_bf16(dst.k[dr] + o + half, o2);
      if (dst.v[0] != nullptr) {
        const __nv_bfloat16* pv = base + 2 * inner + h * Dh + j0;
        *reinterpret_cast<uint4*>(dst.v[dr] + o) = *reinterpret_cast<const uint4*>(pv);
        *reinterpret_cast<uint4*>(dst.v[dr] + o + half) = *reinterpret_cast<const uint4*>(pv + half);
      }
    }
  }
}

// coalesced copy of a contiguous block to the same offset of every peer buffer (consecutive threads write
// consecutive 16 B, so NVLink sees full-size write packets -- row-strided 16 B stores from a GEMM epilogue do not)
struct PeerList { uint4* p[kMaxCpRanks]; };
__global__ void peer_broadcast_kernel(const uint4* __restrict__ src, PeerList dst, int n_peers, int64_t n16) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n16;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint4 v = src[i];
    for (int r = 0; r < n_peers; ++r) dst.p[r][i] = v;
  }
}

// per-head gate logits [B*n_local, H] (fp32) -> the rank owning each head: dst[r][(b*Nt + t0 + t)*Hl + h%Hl]
struct GatePeers { float* p[kMaxCpRanks]; };
__global__ void gate_scatter_kernel(const float* __restrict__ src, GatePeers dst, int n_local, int H, int Hl, int Nt,
                                    int t0, int64_t total) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int h = i % H;
  const int64_t row = i / H;
  const int t = row % n_local, b = row / n_local;
  dst.p[h / Hl][(static_cast<int64_t>(b) * Nt + t0 + t) * Hl + h % Hl] = src[i];
}

// signal `epoch` into every rank's flag array (slot = my rank), then wait until all my slots reached it
__global__ void cp_barrier_kernel(uint32_t* const* __restrict__ peer_flags, uint32_t* my_flags, int rank, int world,
                                  uint32_t epoch) {
  const int i = threadIdx.x;
  if (i < world) {
    __threadfence_system();
    volatile uint32_t* f = peer_flags[i] + rank;
    *f = epoch;
    volatile uint32_t* m = my_flags + i;
    const long long t0 = clock64();
    while (static_cast<int32_t>(*m - epoch) < 0) {
      if (clock64() - t0 > 40000000000LL) {      // ~20 s: a lost peer becomes an error, not a hung GPU
        printf("ltx2: context-parallel barrier timeout (rank %d waiting for rank %d, epoch %u, have %u)\n", rank, i,
               epoch, *m);
        __trap();
      }
    }
    __threadfence_system();
  }
}

// the same among the ranks [first, first + count) of a larger world; `slot0` = first flag word of this barrier domain
// (every group size has its own words and epoch, so groups of different sizes never see each other's counts)
__global__ void cp_barrier_group_kernel(uint32_t* const* __restrict__ peer_flags, uint32_t* my_flags, int rank, int first,
                                        int count, int slot0, uint32_t epoch) {
  const int i = threadIdx.x;
  if (i < count) {
    const int peer = first + i;
    __threadfence_system();
    volatile uint32_t* f = peer_flags[peer] + slot0 + rank;
    *f = epoch;
    volatile uint32_t* m = my_flags + slot0 + peer;
    const long long t0 = clock64();
    while (static_cast<int32_t>(*m - epoch) < 0) {
      if (clock64() - t0 > 40000000000LL) {
        printf("ltx2: group barrier timeout (rank %d waiting for rank %d, epoch %u, have %u)\n", rank, peer, epoch, *m);
        __trap();
      }
    }
    __threadfence_system();
  }
}

// ---------------------------------------------------------------------------------
// v_transpose: [B*T, inner] -> [B,H,Dh,Tp], 64x64 tiles through shared memory
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
v_transpose_kernel(const __nv_bfloat16* __restrict__ v, int64_t ld, __nv_bfloat16* __restrict__ vt, int T, int Tp,
                   int H, int Dh) {
  __shared__ __nv_bfloat16 tile[64][64 + 2];
  const int t0 = blockIdx.x * 64, c0 = blockIdx.y * 64, b = blockIdx.z;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;    // 64 x 4
  for (int i = ty; i < 64; i += 4) {
    const int t = t0 + i;
    tile[i][tx] = (t < T) ? v[(static_cast<int64_t>(b) * T + t) * ld + c0 + tx] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int i = ty; i < 64; i += 4) {
    const int c = c0 + i;                                     // channel = h*Dh + d
    const int t = t0 + tx;
    if (t < Tp) vt[(static_cast<int64_t>(b) * H * Dh + c) * Tp + t] = tile[tx][i];
  }
}

// ---------------------------------------------------------------------------------
// small_linear: warp per output feature, R <= 8 rows share each weight read
// ---------------------------------------------------------------------------------
constexpr int kMaxSmallRows = 8;

__global__ void __launch_bounds__(256)
small_linear_kernel(const float* __restrict__ x, int R, int K, const __nv_bfloat16* __restrict__ W,
                    const float* __restrict__ bias, float* __restrict__ y, int N, int act_in) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  float acc[kMaxSmallRows];
#pragma unroll
  for (int r = 0; r < kMaxSmallRows; ++r) acc[r] = 0.f;
  const __nv_bfloat16* w = W + static_cast<int64_t>(n) * K;
  for (int k = lane * 8; k < K; k += 256) {
    float wf[8];
    load8_bf16(w + k, wf);
#pragma unroll
    for (int r = 0; r < kMaxSmallRows; ++r) {
      if (r < R) {
        float xf[8];
        load8_f32(x + static_cast<int64_t>(r) * K + k, xf);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float xv = xf[i];
          if (act_in == 1) xv = xv / (1.f + __expf(-xv));
          acc[r] = fmaf(xv, wf[i], acc[r]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kMaxSmallRows; ++r) {
    if (r < R) {
      const float s = warp_sum(acc[r]);
      if (lane == 0) y[static_cast<int64_t>(r) * N + n] = s + (bias ? bias[n] : 0.f);
    }
  }
}

// out[m, h] = x[m,:] . W[h,:] + b[h]   (to_gate_logits for head counts the tensor-core GEMM cannot tile)
__global__ void __launch_bounds__(256)
rowdot_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ W,
              const float* __restrict__ bias, float* __restrict__ out, int M, int H, int K) {
  const int64_t w = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= static_cast<int64_t>(M) * H) return;
  const int m = w / H, h = w % H;
  float acc = 0.f;
  for (int k = lane * 8; k < K; k += 256) {
    float a[8], b[8];
    load8_bf16(x + m * ldx + k, a);
    load8_bf16(W + static_cast<int64_t>(h) * K + k, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc = fmaf(a[i], b[i], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[w] = acc + (bias ? bias[h] : 0.f);
}

__global__ void timestep_sinusoid_kernel(const float* __restrict__ t, int R, float mult, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * 128) return;
  const int r = i / 128, f = i % 128;
  const float freq = expf(-9.210340371976184f * static_cast<float>(f) / 128.0f);   // ln(10000)
  const float a = t[r] * mult * freq;
  out[r * 256 + f] = cosf(a);
  out[r * 256 + 128 + f] = sinf(a);
}

__global__ void build_modulation_kernel(const float* __restrict__ tables, int64_t table_layer_stride,
                                        const float* __restrict__ emb, int64_t emb_cls_stride, int64_t emb_row_stride,
                                        float* __restrict__ out, int64_t out_layer_stride, int64_t out_cls_stride,
                                        int L, int C, int R, int D) {
  const int64_t n4 = static_cast<int64_t>(L) * C * R * D / 4;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t e = i * 4;
    const int d = e % D;
    const int k = (e / D) % R;
    const int c = (e / (static_cast<int64_t>(D) * R)) % C;
    const int l = e / (static_cast<int64_t>(D) * R * C);
    const float4 a = *reinterpret_cast<const float4*>(tables + l * table_layer_stride + static_cast<int64_t>(k) * D + d);
    const float4 b = *reinterpret_cast<const float4*>(emb + c * emb_cls_stride + k * emb_row_stride + d);
    *reinterpret_cast<float4*>(out + l * out_layer_stride + c * out_cls_stride + static_cast<int64_t>(k) * D + d) =
        make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}

// positions [B,n_dims,T,2] -> cos/sin [B,T,dim/2]; freq_grid [n_freq] = theta^linspace(0,1,n_freq) * pi/2
__global__ void rope_tables_kernel(const float* __restrict__ pos, int pos_dims, int n_dims, int T, int half, int n_freq,
                                   const float* __restrict__ freq_grid, float3 max_pos, float* __restrict__ cosb,
                                   float* __restrict__ sinb, int64_t total) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int col = i % half;
  const int64_t bt = i / half;
  const int t = bt % T, b = bt / T;
  const int pad = half - n_freq * n_dims;
  float c = 1.f, s = 0.f;
  if (col >= pad) {
    const int f = col - pad;
    const int fi = f / n_dims, ax = f % n_dims;       // frequency-major, axis-minor (rope.py:285-287)
    const float* p = pos + ((static_cast<int64_t>(b) * pos_dims + ax) * T + t) * 2;
    const float mid = (p[0] + p[1]) / 2.0f;
    const float mp = ax == 0 ? max_pos.x : (ax == 1 ? max_pos.y : max_pos.z);
    const float frac = mid / mp;
    const float a = freq_grid[fi] * (frac * 2.0f - 1.0f);
    c = cosf(a);
    s = sinf(a);
  }
  cosb[i] = c;
  sinb[i] = s;
}

__global__ void x0_kernel(const float* __restrict__ latent, const float* __restrict__ vel,
                          const float* __restrict__ t_row, float* __restrict__ x0, int64_t n, int C) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  x0[i] = latent[i] - t_row[i / C] * vel[i];
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half(v); }

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ s, D* __restrict__ d, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    d[i] = from_f<D>(to_f<S>(s[i]));
}

// op: 0 silu_mul, 1 gelu_mul
template <typename T, int OP>
__global__ void act_mul_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ o, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float x = to_f<T>(a[i]), y = to_f<T>(b[i]);
    const float act = OP == 0 ? x / (1.f + expf(-x)) : gelu_tanh_f(x);
    o[i] = from_f<T>(act * y);
  }
}

// pairs (x[2i], x[2i+1]); cos/sin pre-broadcast to x's shape (fused_ops.py:136-180)
template <typename T>
__global__ void interleaved_rope_kernel(const T* __restrict__ x, const T* __restrict__ c, const T* __restrict__ s,
                                        T* __restrict__ o, int64_t npairs) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < npairs;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float xe = to_f<T>(x[2 * i]), xo = to_f<T>(x[2 * i + 1]);
    o[2 * i] = from_f<T>(xe * to_f<T>(c[2 * i]) - xo * to_f<T>(s[2 * i]));
    o[2 * i + 1] = from_f<T>(xo * to_f<T>(c[2 * i + 1]) + xe * to_f<T>(s[2 * i + 1]));
  }
}

inline int ew_grid(int64_t n, int threads) {
  int64_t g = (n + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

int norm_modulate(const void* x, int x_is_bf16, int64_t ldx, void* out, int64_t ldo, int M, int D, int norm_kind,
                  float eps, const float* mod, int64_t mod_stride, int64_t shift_off, int64_t scale_off,
                  const int* row_cls, cudaStream_t stream) {
  if (M == 0) return LTX2_OK;
  LTX2_REQUIRE(D % 8 == 0 && D <= kRowThreads * 8 * kMaxUnits, "norm_modulate: D=%d unsupported", D);
  LTX2_REQUIRE(ldx % 8 == 0 && ldo % 8 == 0 && shift_off % 4 == 0 && scale_off % 4 == 0 && mod_stride % 4 == 0,
               "norm_modulate: pitches/offsets must keep 16-byte alignment");
  if (x_is_bf16)
    LTX2_CUDA_CHECK(launch_pdl(norm_modulate_kernel<true>, dim3(M), dim3(kRowThreads), 0, stream, x, ldx,
                               reinterpret_cast<__nv_bfloat16*>(out), ldo, D, norm_kind, eps, mod, mod_stride, shift_off,
                               scale_off, row_cls));
  else
    LTX2_CUDA_CHECK(launch_pdl(norm_modulate_kernel<false>, dim3(M), dim3(kRowThreads), 0, stream, x, ldx,
                               reinterpret_cast<__nv_bfloat16*>(out), ldo, D, norm_kind, eps, mod, mod_stride, shift_off,
                               scale_off, row_cls));
  count_launch();
  return LTX2_OK;
}

int headnorm_rope(const void* in, int64_t ld, const float* weight, const float* cos, const float* sin, void* out,
                  int B, int T, int H, int Dh, float eps, cudaStream_t stream) {
  if (B * T == 0) return LTX2_OK;
  const int inner = H * Dh;
  LTX2_REQUIRE(Dh % 16 == 0 && inner <= kRowThreads * 16 * kHeadUnits, "headnorm_rope: H=%d Dh=%d unsupported", H, Dh);
  LTX2_REQUIRE(ld % 8 == 0, "headnorm_rope: pitch must be a multiple of 8");
  LTX2_CUDA_CHECK(launch_pdl(headnorm_rope_kernel, dim3(B * T), dim3(kRowThreads), 0, stream,
                             reinterpret_cast<const __nv_bfloat16*>(in), ld, weight, cos, sin,
                             reinterpret_cast<__nv_bfloat16*>(out), T, H, Dh, eps));
  count_launch();
  return LTX2_OK;
}

int qkv_head_scatter(const void* qkv, int64_t ld, const float* wq, const float* wk, const float* cos, const float* sin,
                     const HeadScatter& dst, int B, int T, int H, int Dh, float eps, cudaStream_t stream) {
  if (B * T == 0) return LTX2_OK;
  const int inner = H * Dh;
  LTX2_REQUIRE(Dh % 16 == 0 && inner <= kRowThreads * 16 * kHeadUnits, "qkv_head_scatter: H=%d Dh=%d unsupported", H, Dh);
  LTX2_REQUIRE(ld % 8 == 0 && cos != nullptr && sin != nullptr, "qkv_head_scatter: bad pitch or missing RoPE tables");
  LTX2_REQUIRE(dst.heads_per_rank > 0 && H % dst.heads_per_rank == 0 && H / dst.heads_per_rank <= kMaxCpRanks,
               "qkv_head_scatter: %d heads cannot be split %d per rank", H, dst.heads_per_rank);
  LTX2_CUDA_CHECK(launch_pdl(qkv_head_scatter_kernel, dim3(B * T), dim3(kRowThreads), 0, stream,
                             reinterpret_cast<const __nv_bfloat16*>(qkv), ld, wq, wk, cos, sin, dst, T, H, Dh, eps));
  count_launch();
  return LTX2_OK;
}

int peer_broadcast(const void* src, void* const* dst_peers, int n_peers, int64_t bytes, cudaStream_t stream) {
  LTX2_REQUIRE(bytes % 16 == 0 && n_peers <= kMaxCpRanks, "peer_broadcast: size must be a multiple of 16 bytes");
  if (bytes == 0 || n_peers == 0) return LTX2_OK;
  PeerList pl;
  for (int r = 0; r < n_peers; ++r) pl.p[r] = reinterpret_cast<uint4*>(dst_peers[r]);
  const int64_t n16 = bytes / 16;
  int grid = static_cast<int>((n16 + 255) / 256);
  if (grid > num_sms() * 2) grid = num_sms() * 2;
  peer_broadcast_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(src), pl, n_peers, n16);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int gate_scatter(const float* logits, float* const* dst_peers, int B, int n_local, int H, int heads_per_rank,
                 int n_total, int t_offset, cudaStream_t stream) {
  LTX2_REQUIRE(H % heads_per_rank == 0 && H / heads_per_rank <= kMaxCpRanks, "gate_scatter: bad head split");
  const int64_t total = static_cast<int64_t>(B) * n_local * H;
  if (total == 0) return LTX2_OK;
  GatePeers gp;
  for (int r = 0; r < H / heads_per_rank; ++r) gp.p[r] = dst_peers[r];
  gate_scatter_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(logits, gp, n_local, H, heads_per_rank,
                                                                                    n_total, t_offset, total);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int cp_barrier(uint32_t* const* peer_flags_dev, uint32_t* my_flags, int rank, int world, uint32_t epoch,
               cudaStream_t stream) {
  cp_barrier_kernel<<<1, 32, 0, stream>>>(peer_flags_dev, my_flags, rank, world, epoch);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int cp_barrier_group(uint32_t* const* peer_flags_dev, uint32_t* my_flags, int rank, int first, int count, int slot0,
                     uint32_t epoch, cudaStream_t stream) {
  cp_barrier_group_kernel<<<1, 32, 0, stream>>>(peer_flags_dev, my_flags, rank, first, count, slot0, epoch);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int v_transpose(const void* v, int64_t ld, void* vt, int B, int T, int Tp, int H, int Dh, cudaStream_t stream) {
  if (B * T == 0) return LTX2_OK;
  LTX2_REQUIRE((H * Dh) % 64 == 0, "v_transpose: inner dim must be a multiple of 64");
  dim3 grid((Tp + 63) / 64, (H * Dh) / 64, B);
  v_transpose_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(v), ld,
                                               reinterpret_cast<__nv_bfloat16*>(vt), T, Tp, H, Dh);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int small_linear(const float* x, int R, int K, const void* W, const float* bias, float* y, int N, int act_in,
                 cudaStream_t stream) {
  LTX2_REQUIRE(R >= 1 && R <= kMaxSmallRows, "small_linear: R=%d rows unsupported (1..8)", R);
  LTX2_REQUIRE(K % 8 == 0, "small_linear: K must be a multiple of 8");
  small_linear_kernel<<<(N + 7) / 8, 256, 0, stream>>>(x, R, K, reinterpret_cast<const __nv_bfloat16*>(W), bias, y, N,
                                                       act_in);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int rowdot_bf16(const void* x, int64_t ldx, const void* W, const float* bias, float* out, int M, int H, int K,
                cudaStream_t stream) {
  LTX2_REQUIRE(K % 8 == 0 && ldx % 8 == 0, "rowdot: K and pitch must be multiples of 8");
  const int64_t warps = static_cast<int64_t>(M) * H;
  if (warps == 0) return LTX2_OK;
  rowdot_kernel<<<static_cast<unsigned>((warps + 7) / 8), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, reinterpret_cast<const __nv_bfloat16*>(W), bias, out, M, H, K);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int timestep_sinusoid(const float* t, int R, float multiplier, float* out256, cudaStream_t stream) {
  timestep_sinusoid_kernel<<<(R * 128 + 127) / 128, 128, 0, stream>>>(t, R, multiplier, out256);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int build_modulation_ex(const float* tables, int64_t table_layer_stride, const float* emb, int64_t emb_cls_stride,
                        int64_t emb_row_stride, float* out, int64_t out_layer_stride, int64_t out_cls_stride, int L,
                        int C, int R, int D, cudaStream_t stream) {
  LTX2_REQUIRE(D % 4 == 0 && table_layer_stride % 4 == 0 && emb_cls_stride % 4 == 0 && emb_row_stride % 4 == 0 &&
                   out_layer_stride % 4 == 0 && out_cls_stride % 4 == 0,
               "build_modulation: D and strides must be multiples of 4");
  const int64_t n4 = static_cast<int64_t>(L) * C * R * D / 4;
  if (n4 == 0) return LTX2_OK;
  build_modulation_kernel<<<ew_grid(n4, 256), 256, 0, stream>>>(tables, table_layer_stride, emb, emb_cls_stride,
                                                               emb_row_stride, out, out_layer_stride, out_cls_stride,
                                                               L, C, R, D);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

// freq_grid is a device array of n_freq floats (see make_freq_grid in dit_engine.cu)
int rope_tables_dev(const float* positions, int B, int pos_dims, int n_dims, int T, int dim, const float* max_pos_host,
                    const float* freq_grid_dev, int n_freq, float* cos, float* sin, cudaStream_t stream) {
  LTX2_REQUIRE(pos_dims >= n_dims, "rope_tables: positions carry %d axes, %d requested", pos_dims, n_dims);
  LTX2_REQUIRE(n_dims >= 1 && n_dims <= 3, "rope_tables: n_dims=%d unsupported", n_dims);
  const int half = dim / 2;
  const int64_t total = static_cast<int64_t>(B) * T * half;
  if (total == 0) return LTX2_OK;
  float3 mp = make_float3(max_pos_host[0], n_dims > 1 ? max_pos_host[1] : 1.f, n_dims > 2 ? max_pos_host[2] : 1.f);
  rope_tables_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(positions, pos_dims, n_dims, T, half, n_freq,
                                                                                   freq_grid_dev, mp, cos, sin, total);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

namespace {

// The elementwise tail of a denoising step in one pass (all fp32):
//   d = cond + (cfg_scale - 1) * (cond - uncond)           CFGGuider.guide        (components/guiders.py:40-44)
//   d = d * mask[row] + clean * (1 - mask[row])             post_process_latent    (pipelines/common.py:169-190)
//   out = sample + (sample - d) / sigma * (sigma_next - sigma)   EulerDiffusionStep.step (diffusion_steps.py:55-67,
//                                                                 to_velocity core_utils.py:34-62)
// uncond / mask / clean may be null (that stage is skipped); denoised_out (optional) receives d.
__global__ void denoise_update_kernel(const float* __restrict__ sample, const float* __restrict__ cond,
                                      const float* __restrict__ uncond, float cfg_m1, const float* __restrict__ mask,
                                      const float* __restrict__ clean, float inv_sigma, float dt,
                                      float* __restrict__ out, float* __restrict__ denoised_out, int64_t n4, int C4) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 x = reinterpret_cast<const float4*>(sample)[i];
  float4 d = reinterpret_cast<const float4*>(cond)[i];
  if (uncond != nullptr) {
    const float4 u = reinterpret_cast<const float4*>(uncond)[i];
    d.x = d.x + cfg_m1 * (d.x - u.x);
    d.y = d.y + cfg_m1 * (d.y - u.y);
    d.z = d.z + cfg_m1 * (d.z - u.z);
    d.w = d.w + cfg_m1 * (d.w - u.w);
  }
  if (mask != nullptr) {
    const float m = mask[i / C4];
    const float4 c = reinterpret_cast<const float4*>(clean)[i];
    d.x = d.x * m + c.x * (1.f - m);
    d.y = d.y * m + c.y * (1.f - m);
    d.z = d.z * m + c.z * (1.f - m);
    d.w = d.w * m + c.w * (1.f - m);
  }
  if (denoised_out != nullptr) reinterpret_cast<float4*>(denoised_out)[i] = d;
  float4 o;
  o.x = x.x + ((x.x - d.x) * inv_sigma) * dt;
  o.y = x.y + ((x.y - d.y) * inv_sigma) * dt;
  o.z = x.z + ((x.z - d.z) * inv_sigma) * dt;
  o.w = x.w + ((x.w - d.w) * inv_sigma) * dt;
  reinterpret_cast<float4*>(out)[i] = o;
}

}  // namespace

int denoise_update(const float* sample, const float* cond, const float* uncond, float cfg_scale, const float* mask,
                   const float* clean, float sigma, float sigma_next, float* out, float* denoised_out, int M, int C,
                   cudaStream_t stream) {
  LTX2_REQUIRE(sigma != 0.f, "denoise_update: sigma can't be 0.0");          // core_utils.py:54-55
  LTX2_REQUIRE(C % 4 == 0, "denoise_update: channel count %d must be a multiple of 4", C);
  LTX2_REQUIRE((mask == nullptr) == (clean == nullptr), "denoise_update: mask and clean latent go together");
  const int64_t n4 = static_cast<int64_t>(M) * C / 4;
  if (n4 == 0) return LTX2_OK;
  denoise_update_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, stream>>>(
      sample, cond, uncond, cfg_scale - 1.f, mask, clean, 1.f / sigma, sigma_next - sigma, out, denoised_out, n4, C / 4);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int x0_from_velocity(const float* latent, const float* velocity, const float* t_row, float* x0, int M, int C,
                     cudaStream_t stream) {
  const int64_t n = static_cast<int64_t>(M) * C;
  if (n == 0) return LTX2_OK;
  x0_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(latent, velocity, t_row, x0, n, C);
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int cast_to_bf16(const void* src, int src_dtype, void* dst, int64_t n, cudaStream_t stream) {
  if (n == 0) return LTX2_OK;
  __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst);
  const int g = ew_grid(n, 256);
  switch (src_dtype) {
    case LTX2_F32: cast_kernel<<<g, 256, 0, stream>>>(reinterpret_cast<const float*>(src), d, n); break;
    case LTX2_BF16: LTX2_CUDA_CHECK(cudaMemcpyAsync(dst, src, n * 2, cudaMemcpyDeviceToDevice, stream)); break;
    case LTX2_F16: cast_kernel<<<g, 256, 0, stream>>>(reinterpret_cast<const __half*>(src), d, n); break;
    default: set_error("cast_to_bf16: bad dtype %d", src_dtype); return LTX2_ERR_INVALID;
  }
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int cast_to_f32(const void* src, int src_dtype, float* dst, int64_t n, cudaStream_t stream) {
  if (n == 0) return LTX2_OK;
  const int g = ew_grid(n, 256);
  switch (src_dtype) {
    case LTX2_F32: LTX2_CUDA_CHECK(cudaMemcpyAsync(dst, src, n * 4, cudaMemcpyDeviceToDevice, stream)); break;
    case LTX2_BF16: cast_kernel<<<g, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(src), dst, n); break;
    case LTX2_F16: cast_kernel<<<g, 256, 0, stream>>>(reinterpret_cast<const __half*>(src), dst, n); break;
    default: set_error("cast_to_f32: bad dtype %d", src_dtype); return LTX2_ERR_INVALID;
  }
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

template <int OP>
static int act_mul(const void* a, const void* b, void* out, int64_t n, int dtype, cudaStream_t stream) {
  if (n == 0) return LTX2_OK;
  const int g = ew_grid(n, 256);
  switch (dtype) {
    case LTX2_F32:
      act_mul_kernel<float, OP><<<g, 256, 0, stream>>>(reinterpret_cast<const float*>(a),
                                                       reinterpret_cast<const float*>(b), reinterpret_cast<float*>(out), n);
      break;
    case LTX2_BF16:
      act_mul_kernel<__nv_bfloat16, OP><<<g, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(a),
                                                               reinterpret_cast<const __nv_bfloat16*>(b),
                                                               reinterpret_cast<__nv_bfloat16*>(out), n);
      break;
    case LTX2_F16:
      act_mul_kernel<__half, OP><<<g, 256, 0, stream>>>(reinterpret_cast<const __half*>(a),
                                                        reinterpret_cast<const __half*>(b),
                                                        reinterpret_cast<__half*>(out), n);
      break;
    default: set_error("act_mul: bad dtype %d", dtype); return LTX2_ERR_INVALID;
  }
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

int silu_mul(const void* a, const void* b, void* out, int64_t n, int dtype, cudaStream_t stream) {
  return act_mul<0>(a, b, out, n, dtype, stream);
}
int gelu_mul(const void* a, const void* b, void* out, int64_t n, int dtype, cudaStream_t stream) {
  return act_mul<1>(a, b, out, n, dtype, stream);
}

int interleaved_rope(const void* x, const void* c, const void* s, void* out, int64_t n, int dtype,
                     cudaStream_t stream) {
  if (n == 0) return LTX2_OK;
  LTX2_REQUIRE(n % 2 == 0, "interleaved_rope: element count must be even");
  const int64_t np = n / 2;
  const int g = ew_grid(np, 256);
  switch (dtype) {
    case LTX2_F32:
      interleaved_rope_kernel<<<g, 256, 0, stream>>>(reinterpret_cast<const float*>(x), reinterpret_cast<const float*>(c),
                                                     reinterpret_cast<const float*>(s), reinterpret_cast<float*>(out), np);
      break;
    case LTX2_BF16:
      interleaved_rope_kernel<<<g, 256, 0, stream>>>(
          reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(c),
          reinterpret_cast<const __nv_bfloat16*>(s), reinterpret_cast<__nv_bfloat16*>(out), np);
      break;
    case LTX2_F16:
      interleaved_rope_kernel<<<g, 256, 0, stream>>>(reinterpret_cast<const __half*>(x), reinterpret_cast<const __half*>(c),
                                                     reinterpret_cast<const __half*>(s), reinterpret_cast<__half*>(out), np);
      break;
    default: set_error("interleaved_rope: bad dtype %d", dtype); return LTX2_ERR_INVALID;
  }
  LTX2_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return LTX2_OK;
}

}  // namespace ltx2
