// DiT engine: owns the packed weights and the activation workspace, and issues the kernel
// sequence of one LTXModel forward on a CUDA stream.
//
// Reference path replaced (LTX_2_MLX/model/transformer/):
//   model.py:231-281,368-410   TransformerArgsPreprocessor / MultiModal...prepare
//   transformer.py:457-648     BasicAVTransformerBlock.__call__ (the 48-iteration hot loop, model.py:720-728)
//   model.py:744-774           output heads,   model.py:895-936  X0Model
//
// HBM layout
//   weights   one arena; per Linear a bf16 [out,in] matrix (nn.Linear layout == K-major B operand) and an
//             fp32 bias; q/k/v of a self-attention are adjacent so one GEMM (N = 3*inner) produces all three;
//             k/v of a cross-attention are adjacent (N = 2*inner).  Norm weights and adaLN tables are fp32.
//   residual  x: fp32 [B*N, D] (updated in place by the GEMM epilogues that end each sub-layer)
//   GEMM inputs  bf16 row-major (written by the norm kernels / previous epilogues)
//   attention operands  Q,K: bf16 [B,H,T,Dh];  V: bf16 [B,H,Dh,T] (transposed so P*V is K-major)
//   modulation  fp32 [layer, class, row, D]: adaLN table row + timestep-embedding row, one "class" per
//             distinct (batch, sigma); rows map to classes through an int32 array (per-token timesteps
//             of image conditioning become 2 classes instead of a (B,N,6,D) tensor).
#include "common.cuh"
#include "kernels.h"
#include "../../include/ltx2_b200.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>

#include <map>
#include <string>
#include <unordered_map>
#include <vector>

namespace ltx2 {

typedef __nv_bfloat16 bf16;

namespace {

struct Slot {
  void* dst = nullptr;
  float* cscale = nullptr;        // storage LTX2_F8E4M3: one scale per output row
  int storage = LTX2_F32;
  int64_t rows = 0, cols = 0;     // cols == 0 -> 1-D of `rows`
  bool loaded = false;
};

// Bump allocator.  A dry pass (base == 0) only measures; addresses are formed with integer arithmetic so the
// dry and the real pass run exactly the same code.
struct Arena {
  uintptr_t base = 0;
  size_t off = 0;
  void* take(size_t bytes) {
    size_t a = (off + 255) & ~size_t(255);
    off = a + bytes;
    return reinterpret_cast<void*>(base + a);
  }
};
template <typename T>
inline T* offset_ptr(T* p, size_t elems) {
  return reinterpret_cast<T*>(reinterpret_cast<uintptr_t>(p) + elems * sizeof(T));
}

struct LinearW {
  bf16* w = nullptr;
  uint8_t* w8 = nullptr;          // FP8 mode (cfg.fp8_linear) for the norm-fed linears: E4M3 [out,in] instead of `w`
  float* cscale = nullptr;        //   and one scale per output row (checkpoint weight_scale, or absmax/448 per row)
  float* b = nullptr;
  int out = 0, in = 0;
};

struct AttnW {
  LinearW q, kv, o, gate;      // kv: [2*inner, ctx_dim]; for self-attention q.w is followed by kv.w (fused QKV)
  float* qnorm = nullptr;
  float* knorm = nullptr;
  bool fused_qkv = false;
  int heads = 0, dh = 0, inner = 0;
};

struct AdaLNW {                // AdaLayerNormSingle
  LinearW l1, l2, lin;
  int n_emb = 0;
};

struct StreamW {               // per-modality weights outside the blocks
  LinearW patchify, cap1, cap2, proj_out;
  AdaLNW adaln, prompt_adaln;
  float* head_table = nullptr;   // [2, dim]
  bool has_caption = false;
};

struct BlockStreamW {          // per-block, per-modality
  AttnW attn1, attn2;
  LinearW ff1, ff2;
  float* ff_coef = nullptr;      // FP8 mode: two device floats, scale bound of the E4M3 FFN hidden (e4m3_bound_coef)
  float* table = nullptr;        // [n_ada, dim]
  float* prompt_table = nullptr; // [2, dim]
};

struct BlockW {
  BlockStreamW v, a;
  AttnW a2v, v2a;
  float* table_ca_audio = nullptr;  // [5, Da]
  float* table_ca_video = nullptr;  // [5, D]
};

// activation workspace of one modality
struct StreamBuf {
  int B = 0, N = 0, S = 0, dim = 0, heads = 0, dh = 0, n_cls = 0;
  float* x = nullptr;           // [M, dim] fp32 residual
  bf16* xn = nullptr;           // [M, dim]
  uint8_t* xq = nullptr;        // [M, dim] E4M3 row-quantised GEMM input (FP8 mode)
  float* xs = nullptr;          // [M] its per-row scales
  float* xl2 = nullptr;         // [M] L2 norm of those rows (bound of the E4M3 FFN hidden)
  uint8_t* hq = nullptr;        // [M, 4*dim] E4M3 FFN hidden (written by the up-projection's epilogue)
  bf16* qkv = nullptr;          // [M, 3*dim]
  bf16* attn = nullptr;         // [M, dim]
  bf16* hidden = nullptr;       // [M, 4*dim]
  bf16* qh = nullptr;           // [B,H,N,Dh]
  bf16* kh = nullptr;           // [B,H,max(N,S,Nother),Dh]
  bf16* vt = nullptr;           // [B,H,Dh,Tp]
  bf16* lat = nullptr;          // [M, C_in] bf16
  bf16* ctx_in = nullptr;       // [B*S, C_ctx]
  bf16* ctx_mid = nullptr;      // [B*S, dim]
  bf16* ctx = nullptr;          // [B*S, dim]
  bf16* ctx_mod = nullptr;      // [B*S, dim] (V2)
  bf16* kv = nullptr;           // [max(B*S, M_other), 2*dim_kv]
  float* gate_logits = nullptr; // [M, H]
  float* cos = nullptr;         // [B,N,dim/2]
  float* sin = nullptr;
  float* ccos = nullptr;        // cross-modal 1-D rope [B,N,Da/2]
  float* csin = nullptr;
  float* mod = nullptr;         // [L, n_cls, n_ada, dim]
  float* prompt_mod = nullptr;  // [L, B, 2, dim]
  float* ca_mod = nullptr;      // [L, B, 5, dim]
  float* head_mod = nullptr;    // [n_cls, 2, dim]
  float* emb_t = nullptr;       // [n_cls, dim] embedded timestep
  int* row_cls = nullptr;       // [M]
  int* row_batch = nullptr;     // [max(M, B*S)] = row / tokens
  int* ctx_batch = nullptr;     // [B*S]
  float* t_cls = nullptr;       // [n_cls] sigma per class
  float* t_row = nullptr;       // [M] sigma per row
  float* vel = nullptr;         // [M, C_out]
};

}  // namespace

}  // namespace ltx2

using namespace ltx2;

// Optional per-kernel-class timing (bench.py's roofline leg): CUDA events around every GEMM and attention
// launch of one forward, on the launching stream.  Off by default -- the timed bench steps run without it.
struct DitProfiler {
  bool on = false;
  std::vector<cudaEvent_t> events;
  size_t used = 0;
  struct Rec { int cat; double work; size_t e0; };
  std::vector<Rec> recs;
  size_t begin(int cat, double work, cudaStream_t st) {
    if (used + 2 > events.size()) {
      const size_t old = events.size();
      events.resize(old + 512);
      for (size_t i = old; i < events.size(); ++i) cudaEventCreate(&events[i]);
    }
    recs.push_back({cat, work, used});
    cudaEventRecord(events[used], st);
    used += 2;
    return used - 2;
  }
  void end(size_t e0, cudaStream_t st) { cudaEventRecord(events[e0 + 1], st); }
};
static thread_local DitProfiler* g_prof = nullptr;
static thread_local int g_split_k = 1;
struct ProfScope {
  size_t e0 = 0;
  cudaStream_t st;
  bool on;
  ProfScope(int cat, double work, cudaStream_t s) : st(s), on(g_prof && g_prof->on) {
    if (on) e0 = g_prof->begin(cat, work, st);
  }
  ~ProfScope() {
    if (on) g_prof->end(e0, st);
  }
};
enum { PROF_GEMM = 0, PROF_ATTN = 1, PROF_GEMM8 = 2 };   // class 2: FP8 GEMM launches (folded into 0 by 2-class readers)

// Context-parallel state: a cudaMalloc'ed exchange region with the same layout on every rank, mapped into each
// process through CUDA IPC, so kernels can store into a peer's buffers over NVLink.
struct CpState {
  int rank = 0, world = 1;
  int B = 0, n_total = 0, n_local = 0, heads_local = 0;
  char* region = nullptr;
  size_t region_bytes = 0;
  char* peer_base[kMaxCpRanks] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool opened[kMaxCpRanks] = {false, false, false, false, false, false, false, false};
  size_t off_q = 0, off_k = 0, off_v = 0, off_o = 0, off_flags = 0;
  size_t off_ckv[2] = {0, 0};      // text-context K/V of the current / next block, [B*S, 2*D] bf16 each
  size_t off_gate = 0;              // gated self-attention: per-head gate logits of all tokens for MY heads, fp32
  size_t off_akv = 0;               // AV: video-side K|V projection of the v2a attention for ALL tokens [B*Nt, 2*Da]
  size_t off_crope[2][2] = {{0, 0}, {0, 0}};   // AV: cross-modal RoPE cos/sin of all video tokens, double-buffered per forward
  int fwd_parity = 0;
  int split_k = 8;                  // split-K cap of the residual GEMMs on sharded ranks (LTX2_CP_SPLIT_K, 1 = off)
  int ctx_tokens = 0;               // > 0: every rank projects S/P context rows and broadcasts them to all peers
  uint32_t** peer_flags_dev = nullptr;
  uint32_t epoch = 0;
  bool connected = false;
};

struct LtxDit {
  CpState cp;
  DitProfiler prof;
  LtxDitConfig cfg;
  int D = 0, Da = 0, n_ada = 6;
  std::unordered_map<std::string, Slot> slots;
  char* arena = nullptr;
  size_t arena_bytes = 0;
  StreamW vw, aw;
  AdaLNW av_v_ss, av_v_gate, av_a_ss, av_a_gate;
  std::vector<BlockW> blocks;
  std::vector<float> cross_attn_scale;
  float* table_arena_v = nullptr;   // [L, n_ada, D]   (contiguous per-layer tables -> one build_modulation)
  float* table_arena_a = nullptr;
  float* ptable_arena_v = nullptr;  // [L, 2, D]
  float* ptable_arena_a = nullptr;
  float* catable_arena_v = nullptr; // [L, 5, D]
  float* catable_arena_a = nullptr; // [L, 5, Da]
  // frequency grids (device)
  float* fg_video = nullptr; int nf_video = 0;       // dim D, 3 axes
  float* fg_audio = nullptr; int nf_audio = 0;       // dim Da, 1 axis
  // workspace
  char* ws = nullptr;
  size_t ws_bytes = 0;
  StreamBuf vb, ab;
  float* scratch = nullptr;         // small fp32 scratch for timestep MLPs
  float* scale1 = nullptr;          // one device float: per-tensor scale of an FP8 tensor being widened at ingest
  std::vector<float> h_ts;          // host staging for per-token timestep dedupe
  // text-context reuse across steps (ltx2_dit_set_context_tag): per block K|V projection of the context
  // [B*S, 2*inner] and the k-normed, head-split K [B,H,S,Dh]; valid for (cached_tag, cached_B, cached_S)
  uint64_t ctx_tag = 0, cached_tag = 0;
  int cached_B = 0, cached_S = 0;
  char* kvc = nullptr;
  size_t kvc_bytes = 0;
  bool ctx_cache_on = false, ctx_cache_hit = false;   // state of the forward in flight
  int layer_limit = 0;              // diagnostics: run only the first n blocks (0 = all)
  bool coef_dirty = true;           // FP8 mode: a weight changed -> recompute the FFN bound coefficients
  // audio+video models: the audio stream's self/text attention and FFN (65 tokens: ~30 latency-bound launches per
  // block) run on a side stream next to the video stream's kernels and join it around the cross-modal attention
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bf16* kvc_kv(int layer, int B, int S) const {
    return reinterpret_cast<bf16*>(kvc) + size_t(layer) * 3 * size_t(B) * S * D;
  }
  bf16* kvc_kh(int layer, int B, int S) const { return kvc_kv(layer, B, S) + size_t(2) * size_t(B) * S * D; }
};

namespace ltx2 {
namespace {

// ------------------------------------------------------------------------------
// weight layout
// ------------------------------------------------------------------------------
struct Layout {
  LtxDit* e;
  Arena* ar;
  bool dry;

  void slot(const std::string& key, void* dst, int storage, int64_t rows, int64_t cols) {
    if (dry) return;
    Slot s;
    s.dst = dst; s.storage = storage; s.rows = rows; s.cols = cols;
    e->slots[key] = s;
  }
  float* f32(const std::string& key, int64_t rows, int64_t cols) {
    float* p = reinterpret_cast<float*>(ar->take(sizeof(float) * rows * (cols ? cols : 1)));
    slot(key, p, LTX2_F32, rows, cols);
    return p;
  }
  // registers an fp32 tensor inside a pre-taken region
  float* f32_at(const std::string& key, float* p, int64_t rows, int64_t cols) {
    slot(key, p, LTX2_F32, rows, cols);
    return p;
  }
  LinearW linear_at(const std::string& prefix, int out, int in, bf16* w_at, float* b_at) {
    LinearW L;
    L.out = out; L.in = in;
    L.w = w_at;
    L.b = b_at;
    slot(prefix + ".weight", L.w, LTX2_BF16, out, in);
    slot(prefix + ".bias", L.b, LTX2_F32, out, 0);
    return L;
  }
  LinearW linear(const std::string& prefix, int out, int in, bool q8 = false) {
    if (q8) {
      uint8_t* w8 = reinterpret_cast<uint8_t*>(ar->take(size_t(out) * in));
      float* cs = reinterpret_cast<float*>(ar->take(sizeof(float) * out));
      float* b = reinterpret_cast<float*>(ar->take(sizeof(float) * out));
      return linear_q8_at(prefix, out, in, w8, cs, b);
    }
    bf16* w = reinterpret_cast<bf16*>(ar->take(sizeof(bf16) * size_t(out) * in));
    float* b = reinterpret_cast<float*>(ar->take(sizeof(float) * out));
    return linear_at(prefix, out, in, w, b);
  }
  LinearW linear_q8_at(const std::string& prefix, int out, int in, uint8_t* w8_at, float* cs_at, float* b_at) {
    LinearW L;
    L.out = out; L.in = in;
    L.w8 = w8_at; L.cscale = cs_at; L.b = b_at;
    if (!dry) {
      Slot s;
      s.dst = w8_at; s.cscale = cs_at; s.storage = LTX2_F8E4M3; s.rows = out; s.cols = in;
      e->slots[prefix + ".weight"] = s;
    }
    slot(prefix + ".bias", L.b, LTX2_F32, out, 0);
    return L;
  }
  AdaLNW adaln(const std::string& prefix, int dim, int n_emb) {
    AdaLNW a;
    a.n_emb = n_emb;
    a.l1 = linear(prefix + ".emb.timestep_embedder.linear_1", dim, 256);
    a.l2 = linear(prefix + ".emb.timestep_embedder.linear_2", dim, dim);
    a.lin = linear(prefix + ".linear", n_emb * dim, dim);
    return a;
  }
  AttnW attention(const std::string& prefix, int query_dim, int ctx_dim, int heads, int dh, bool self_attn,
                  bool gated, bool q8 = false) {
    AttnW a;
    a.heads = heads; a.dh = dh; a.inner = heads * dh;
    const int inner = a.inner;
    a.fused_qkv = self_attn;
    if (self_attn && q8) {
      // fused QKV in E4M3: one byte matrix [3*inner, query_dim], scales and biases [3*inner]
      uint8_t* w8 = reinterpret_cast<uint8_t*>(ar->take(size_t(3) * inner * query_dim));
      float* cs = reinterpret_cast<float*>(ar->take(sizeof(float) * 3 * inner));
      float* b = reinterpret_cast<float*>(ar->take(sizeof(float) * 3 * inner));
      a.q = linear_q8_at(prefix + ".to_q", inner, query_dim, w8, cs, b);
      a.kv = linear_q8_at(prefix + ".to_k", inner, ctx_dim, w8 + size_t(inner) * query_dim, cs + inner, b + inner);
      linear_q8_at(prefix + ".to_v", inner, ctx_dim, w8 + size_t(2) * inner * query_dim, cs + 2 * inner, b + 2 * inner);
    } else if (self_attn) {
      bf16* w = reinterpret_cast<bf16*>(ar->take(sizeof(bf16) * size_t(3) * inner * query_dim));
      float* b = reinterpret_cast<float*>(ar->take(sizeof(float) * 3 * inner));
      a.q = linear_at(prefix + ".to_q", inner, query_dim, w, b);
      a.kv = linear_at(prefix + ".to_k", inner, ctx_dim, offset_ptr(w, size_t(inner) * query_dim), offset_ptr(b, inner));
      linear_at(prefix + ".to_v", inner, ctx_dim, offset_ptr(w, size_t(2) * inner * query_dim), offset_ptr(b, 2 * inner));
    } else {
      a.q = linear(prefix + ".to_q", inner, query_dim, q8);
      bf16* w = reinterpret_cast<bf16*>(ar->take(sizeof(bf16) * size_t(2) * inner * ctx_dim));
      float* b = reinterpret_cast<float*>(ar->take(sizeof(float) * 2 * inner));
      a.kv = linear_at(prefix + ".to_k", inner, ctx_dim, w, b);
      linear_at(prefix + ".to_v", inner, ctx_dim, offset_ptr(w, size_t(inner) * ctx_dim), offset_ptr(b, inner));
    }
    a.kv.out = 2 * inner;
    a.o = linear(prefix + ".to_out", query_dim, inner);
    a.qnorm = f32(prefix + ".q_norm.weight", inner, 0);
    a.knorm = f32(prefix + ".k_norm.weight", inner, 0);
    if (gated) a.gate = linear(prefix + ".to_gate_logits", heads, query_dim);
    return a;
  }
};

void build_layout(LtxDit* e, Arena* ar, bool dry) {
  Layout L{e, ar, dry};
  const LtxDitConfig& c = e->cfg;
  const int D = e->D, Da = e->Da, n = e->n_ada, nl = c.num_layers;
  const bool v2 = c.cross_attention_adaln != 0, gated = c.apply_gated_attention != 0, audio = c.audio_enabled != 0;
  const bool q8 = c.fp8_linear != 0;     // E4M3 storage + FP8 MMA for the linears fed by a norm kernel

  auto stream_w = [&](StreamW& s, const std::string& p, int dim, int in_ch, int out_ch) {
    s.patchify = L.linear(p + "patchify_proj", dim, in_ch);
    s.adaln = L.adaln(p + "adaln_single", dim, n);
    if (v2) s.prompt_adaln = L.adaln(p + "prompt_adaln_single", dim, 2);
    s.has_caption = c.caption_channels > 0;
    if (s.has_caption) {
      s.cap1 = L.linear(p + "caption_projection.linear_1", dim, c.caption_channels);
      s.cap2 = L.linear(p + "caption_projection.linear_2", dim, dim);
    }
    s.head_table = L.f32(p + "scale_shift_table", 2, dim);
    s.proj_out = L.linear(p + "proj_out", out_ch, dim);
  };
  stream_w(e->vw, "", D, c.in_channels, c.out_channels);
  if (audio) {
    stream_w(e->aw, "audio_", Da, c.audio_in_channels, c.audio_out_channels);
    e->av_v_ss = L.adaln("av_ca_video_scale_shift_adaln_single", D, 4);
    e->av_v_gate = L.adaln("av_ca_a2v_gate_adaln_single", D, 1);
    e->av_a_ss = L.adaln("av_ca_audio_scale_shift_adaln_single", Da, 4);
    e->av_a_gate = L.adaln("av_ca_v2a_gate_adaln_single", Da, 1);
  }
  // contiguous per-layer table arenas
  e->table_arena_v = reinterpret_cast<float*>(ar->take(sizeof(float) * size_t(nl) * n * D));
  if (v2) e->ptable_arena_v = reinterpret_cast<float*>(ar->take(sizeof(float) * size_t(nl) * 2 * D));
  if (audio) {
    e->table_arena_a = reinterpret_cast<float*>(ar->take(sizeof(float) * size_t(nl) * n * Da));
    if (v2) e->ptable_arena_a = reinterpret_cast<float*>(ar->take(sizeof(float) * size_t(nl) * 2 * Da));
    e->catable_arena_v = reinterpret_cast<float*>(ar->take(sizeof(float) * size_t(nl) * 5 * D));
    e->catable_arena_a = reinterpret_cast<float*>(ar->take(sizeof(float) * size_t(nl) * 5 * Da));
  }
  if (!dry) e->blocks.resize(nl);
  for (int i = 0; i < nl; ++i) {
    BlockW tmp;
    BlockW& b = dry ? tmp : e->blocks[i];
    const std::string P = "transformer_blocks." + std::to_string(i) + ".";
    b.v.attn1 = L.attention(P + "attn1", D, D, c.num_attention_heads, c.attention_head_dim, true, gated, q8);
    b.v.attn2 = L.attention(P + "attn2", D, c.cross_attention_dim, c.num_attention_heads, c.attention_head_dim,
                            false, gated, q8);
    b.v.ff1 = L.linear(P + "ff.project_in.proj", 4 * D, D, q8);
    b.v.ff2 = L.linear(P + "ff.project_out", D, 4 * D, q8);
    if (q8) b.v.ff_coef = reinterpret_cast<float*>(ar->take(16));
    b.v.table = L.f32_at(P + "scale_shift_table", offset_ptr(e->table_arena_v, size_t(i) * n * D), n, D);
    if (v2)
      b.v.prompt_table = L.f32_at(P + "prompt_scale_shift_table", offset_ptr(e->ptable_arena_v, size_t(i) * 2 * D), 2, D);
    if (audio) {
      b.a.attn1 = L.attention(P + "audio_attn1", Da, Da, c.audio_heads, c.audio_head_dim, true, gated, q8);
      b.a.attn2 = L.attention(P + "audio_attn2", Da, Da, c.audio_heads, c.audio_head_dim, false, gated, q8);
      b.a.ff1 = L.linear(P + "audio_ff.project_in.proj", 4 * Da, Da, q8);
      b.a.ff2 = L.linear(P + "audio_ff.project_out", Da, 4 * Da, q8);
      if (q8) b.a.ff_coef = reinterpret_cast<float*>(ar->take(16));
      b.a.table = L.f32_at(P + "audio_scale_shift_table", offset_ptr(e->table_arena_a, size_t(i) * n * Da), n, Da);
      if (v2)
        b.a.prompt_table = L.f32_at(P + "audio_prompt_scale_shift_table",
                                    offset_ptr(e->ptable_arena_a, size_t(i) * 2 * Da), 2, Da);
      b.a2v = L.attention(P + "audio_to_video_attn", D, Da, c.audio_heads, c.audio_head_dim, false, gated);
      b.v2a = L.attention(P + "video_to_audio_attn", Da, D, c.audio_heads, c.audio_head_dim, false, gated);
      b.table_ca_audio = L.f32_at(P + "scale_shift_table_a2v_ca_audio",
                                  offset_ptr(e->catable_arena_a, size_t(i) * 5 * Da), 5, Da);
      b.table_ca_video = L.f32_at(P + "scale_shift_table_a2v_ca_video",
                                  offset_ptr(e->catable_arena_v, size_t(i) * 5 * D), 5, D);
    }
  }
}

std::vector<float> make_freq_grid(float theta, int n_dims, int dim) {
  // generate_freq_grid (rope.py:181-211): theta ** linspace(0, 1, dim // (2*n_dims)) * pi/2, stored as fp32
  const int n = dim / (2 * n_dims);
  std::vector<float> g(n);
  for (int i = 0; i < n; ++i) {
    const float lin = n > 1 ? static_cast<float>(static_cast<double>(i) / (n - 1)) : 0.f;
    g[i] = static_cast<float>(pow(static_cast<double>(theta), static_cast<double>(lin)) * (M_PI / 2.0));
  }
  return g;
}

__global__ void fill_row_index_kernel(int* row_batch, int M, int tokens) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M) row_batch[i] = i / tokens;
}
__global__ void gather_f32_kernel(const float* src, const int* idx, float* dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

// ------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------
struct Shapes {
  int B, N, S, n_cls;        // video
  int Na, Sa, n_cls_a;       // audio (0 when absent)
};

size_t layout_stream_buf(StreamBuf& sb, Arena& ar, const LtxDitConfig& c, int dim, int heads, int dh, int in_ch,
                         int out_ch, int ctx_ch, int B, int N, int S, int n_cls, int n_ada, int other_tokens,
                         int other_dim, int cross_dim, bool v2, bool gated, bool av) {
  sb.B = B; sb.N = N; sb.S = S; sb.dim = dim; sb.heads = heads; sb.dh = dh; sb.n_cls = n_cls;
  const size_t M = size_t(B) * N;
  const int L = c.num_layers;
  const size_t maxT = std::max(std::max(N, S), other_tokens);
  const size_t Tp = (maxT + 63) / 64 * 64;
  const int kv_dim = std::max(dim, cross_dim);   // inner dim of the widest K/V this stream's buffers must hold
  auto T = [&](size_t bytes) { return ar.take(bytes); };
  sb.x = (float*)T(M * dim * 4);
  sb.xn = (bf16*)T(M * dim * 2);
  sb.xq = c.fp8_linear ? (uint8_t*)T(M * dim) : nullptr;
  sb.xs = c.fp8_linear ? (float*)T(M * 4) : nullptr;
  sb.xl2 = c.fp8_linear ? (float*)T(M * 4) : nullptr;
  sb.hq = c.fp8_linear ? (uint8_t*)T(M * 4 * dim) : nullptr;
  sb.qkv = (bf16*)T(M * 3 * dim * 2);
  sb.attn = (bf16*)T(M * dim * 2);
  sb.hidden = (bf16*)T(M * 4 * dim * 2);
  sb.qh = (bf16*)T(M * dim * 2);
  sb.kh = (bf16*)T(size_t(B) * maxT * kv_dim * 2);
  sb.vt = (bf16*)T(size_t(B) * Tp * kv_dim * 2);
  sb.lat = (bf16*)T(M * in_ch * 2);
  sb.ctx_in = (bf16*)T(size_t(B) * S * ctx_ch * 2);
  sb.ctx_mid = (bf16*)T(size_t(B) * S * dim * 2);
  sb.ctx = (bf16*)T(size_t(B) * S * dim * 2);
  sb.ctx_mod = v2 ? (bf16*)T(size_t(B) * S * dim * 2) : nullptr;
  sb.kv = (bf16*)T(size_t(B) * std::max<size_t>(S, other_tokens) * 2 * kv_dim * 2);
  sb.gate_logits = gated ? (float*)T(M * std::max(heads, 32) * 4) : nullptr;
  sb.cos = (float*)T(M * (dim / 2) * 4);
  sb.sin = (float*)T(M * (dim / 2) * 4);
  sb.ccos = av ? (float*)T(M * (cross_dim / 2) * 4) : nullptr;
  sb.csin = av ? (float*)T(M * (cross_dim / 2) * 4) : nullptr;
  sb.mod = (float*)T(size_t(L) * n_cls * n_ada * dim * 4);
  sb.prompt_mod = v2 ? (float*)T(size_t(L) * B * 2 * dim * 4) : nullptr;
  sb.ca_mod = av ? (float*)T(size_t(L) * B * 5 * dim * 4) : nullptr;
  sb.head_mod = (float*)T(size_t(n_cls) * 2 * dim * 4);
  sb.emb_t = (float*)T(size_t(n_cls) * dim * 4);
  sb.row_cls = (int*)T(M * 4);
  sb.row_batch = (int*)T(std::max(M, size_t(B) * S) * 4);
  sb.ctx_batch = (int*)T(size_t(B) * S * 4);
  sb.t_cls = (float*)T(size_t(n_cls) * 4);
  sb.t_row = (float*)T(M * 4);
  sb.vel = (float*)T(M * out_ch * 4);
  return ar.off;
}

int ensure_workspace(LtxDit* e, const Shapes& s) {
  const LtxDitConfig& c = e->cfg;
  const bool v2 = c.cross_attention_adaln != 0, gated = c.apply_gated_attention != 0;
  const bool av = s.Na > 0;
  const int ctx_ch_v = c.caption_channels > 0 ? c.caption_channels : e->D;
  const int ctx_ch_a = c.caption_channels > 0 ? c.caption_channels : e->Da;
  for (int pass = 0; pass < 2; ++pass) {
    Arena ar;
    ar.base = pass == 0 ? 0 : reinterpret_cast<uintptr_t>(e->ws);
    StreamBuf vb, ab;
    layout_stream_buf(vb, ar, c, e->D, c.num_attention_heads, c.attention_head_dim, c.in_channels, c.out_channels,
                      ctx_ch_v, s.B, s.N, s.S, s.n_cls, e->n_ada, s.Na, e->Da, e->Da, v2, gated, av);
    if (av)
      layout_stream_buf(ab, ar, c, e->Da, c.audio_heads, c.audio_head_dim, c.audio_in_channels, c.audio_out_channels,
                        ctx_ch_a, s.B, s.Na, s.Sa, s.n_cls_a, e->n_ada, e->cp.world > 1 ? e->cp.n_total : s.N, e->D, e->Da,
                        v2, gated, av);
    float* scratch = (float*)ar.take(size_t(64) * (256 + 12 * std::max(e->D, 1)) * 4);
    if (pass == 0) {
      const size_t need = ar.off + 1024;
      if (need > e->ws_bytes) {
        if (e->ws) cudaFree(e->ws);
        e->ws = nullptr;
        e->ws_bytes = 0;
        if (cudaMalloc(&e->ws, need) != cudaSuccess) {
          set_error("workspace allocation of %zu bytes failed", need);
          return LTX2_ERR_NOMEM;
        }
        e->ws_bytes = need;
      }
    } else {
      e->vb = vb;
      e->ab = ab;
      e->scratch = scratch;
    }
  }
  return LTX2_OK;
}

// ------------------------------------------------------------------------------
// forward helpers
// ------------------------------------------------------------------------------
inline int linear_bf16(const bf16* A, int64_t lda, const LinearW& L, int M, bf16* out, int64_t ldo, bool gelu,
                       cudaStream_t st, int n_rows = -1, int row_off = 0) {
  GemmEpilogue ep;
  ep.mode = gelu ? GEMM_EPI_BF16_GELU : GEMM_EPI_BF16;
  ep.bias = L.b + row_off;
  ep.out = out;
  ep.ldo = ldo;
  const int N = n_rows < 0 ? L.out : n_rows;
  ProfScope ps(PROF_GEMM, 2.0 * M * double(N) * L.in, st);
  return gemm_bf16(A, lda, L.w + size_t(row_off) * L.in, L.in, M, N, L.in, ep, st);
}

// the same linear with E4M3 operands: A8 [M, in] row-quantised by norm_modulate_q8 (scales a_scale[M]), weight L.w8
inline int linear_q8(const uint8_t* A8, const float* a_scale, int64_t lda, const LinearW& L, int M, bf16* out, int64_t ldo,
                     bool gelu, cudaStream_t st, int n_rows = -1, int row_off = 0) {
  GemmEpilogue ep;
  ep.mode = gelu ? GEMM_EPI_BF16_GELU : GEMM_EPI_BF16;
  ep.bias = L.b + row_off;
  ep.out = out;
  ep.ldo = ldo;
  ep.row_scale = a_scale;
  ep.col_scale = L.cscale + row_off;
  const int N = n_rows < 0 ? L.out : n_rows;
  ProfScope ps(PROF_GEMM8, 2.0 * M * double(N) * L.in, st);
  return gemm_e4m3(A8, lda, L.w8 + size_t(row_off) * L.in, L.in, M, N, L.in, ep, st);
}

inline int linear_residual(const bf16* A, int64_t lda, const LinearW& L, int M, float* x, int64_t ldx,
                           const float* gate, int64_t gate_stride, const int* row_cls, float alpha, cudaStream_t st) {
  GemmEpilogue ep;
  ep.mode = GEMM_EPI_F32_RESIDUAL;
  ep.bias = L.b;
  ep.out = x;
  ep.ldo = ldx;
  ep.gate = gate;
  ep.gate_stride = gate_stride;
  ep.row_cls = row_cls;
  ep.alpha = alpha;
  ep.max_splits = g_split_k;        // > 1 only for context-parallel ranks (small M): see gemm_sm100.cu
  ProfScope ps(PROF_GEMM, 2.0 * M * double(L.out) * L.in, st);
  return gemm_bf16(A, lda, L.w, L.in, M, L.out, L.in, ep, st);
}

inline int linear_f32(const bf16* A, int64_t lda, const LinearW& L, int M, float* out, int64_t ldo, cudaStream_t st) {
  GemmEpilogue ep;
  ep.mode = GEMM_EPI_F32;
  ep.bias = L.b;
  ep.out = out;
  ep.ldo = ldo;
  ProfScope ps(PROF_GEMM, 2.0 * M * double(L.out) * L.in, st);
  return gemm_bf16(A, lda, L.w, L.in, M, L.out, L.in, ep, st);
}

// AdaLayerNormSingle on R (<= 64) scalar timesteps already on the device: emb [R, n_emb*dim], e [R, dim]
int run_adaln(LtxDit* e, const AdaLNW& a, const float* t_dev, int R, float mult, int dim, float* emb_out,
              float* e_out, cudaStream_t st) {
  float* sinus = e->scratch;                 // [R,256]
  float* h1 = sinus + size_t(64) * 256;      // [R,dim]
  float* e_tmp = h1 + size_t(64) * dim;      // [R,dim]
  float* eo = e_out ? e_out : e_tmp;
  for (int r0 = 0; r0 < R; r0 += 8) {
    const int r = std::min(8, R - r0);
    LTX2_PROPAGATE(timestep_sinusoid(t_dev + r0, r, mult, sinus, st));
    LTX2_PROPAGATE(small_linear(sinus, r, 256, a.l1.w, a.l1.b, h1, dim, 0, st));
    LTX2_PROPAGATE(small_linear(h1, r, dim, a.l2.w, a.l2.b, eo + size_t(r0) * dim, dim, 1, st));
    LTX2_PROPAGATE(small_linear(eo + size_t(r0) * dim, r, dim, a.lin.w, a.lin.b,
                                emb_out + size_t(r0) * a.n_emb * dim, a.n_emb * dim, 1, st));
  }
  return LTX2_OK;
}

struct AttnCall {
  const AttnW* w;
  const bf16* xq; int64_t ldq; int Mq; int Tq;          // query-side input [B*Tq, query_dim]
  const bf16* xkv; int64_t ldkv; int Tk;               // key/value-side input [B*Tk, ctx_dim] (== xq for self)
  const float *qcos, *qsin, *kcos, *ksin;              // rope tables or null
  const bf16* kv_pre = nullptr;                        // K|V projection already available ([B*Tk, 2*inner]): skip that GEMM
  const bf16* kh_pre = nullptr;                        // K already normalised and head-split ([B,H,Tk,Dh]): skip the k-norm
  const uint8_t* xq8 = nullptr;                        // FP8 mode: the query-side input row-quantised to E4M3 ...
  const float* xq8_scale = nullptr;                    // ... and its per-row scales (norm_modulate_q8)
};

// Attention.__call__ up to (not including) to_out: writes sb.attn [B*Tq, inner]
int run_attention_core(LtxDit* e, StreamBuf& sb, const AttnCall& c, int B, cudaStream_t st) {
  const AttnW& w = *c.w;
  const int inner = w.inner, H = w.heads, Dh = w.dh;
  const float eps = e->cfg.norm_eps;
  const bf16 *qp, *kp, *vp;
  int64_t ldq, ldk;
  LTX2_REQUIRE(w.q.w8 == nullptr || c.xq8 != nullptr, "attention: FP8 projection without a quantised input");
  if (w.fused_qkv) {
    if (w.q.w8 != nullptr)
      LTX2_PROPAGATE(linear_q8(c.xq8, c.xq8_scale, c.ldq, w.q, c.Mq, sb.qkv, 3 * inner, false, st, 3 * inner, 0));
    else
      LTX2_PROPAGATE(linear_bf16(c.xq, c.ldq, w.q, c.Mq, sb.qkv, 3 * inner, false, st, 3 * inner, 0));
    qp = sb.qkv; kp = sb.qkv + inner; vp = sb.qkv + 2 * inner;
    ldq = ldk = 3 * inner;
  } else {
    if (w.q.w8 != nullptr)
      LTX2_PROPAGATE(linear_q8(c.xq8, c.xq8_scale, c.ldq, w.q, c.Mq, sb.qkv, inner, false, st));
    else
      LTX2_PROPAGATE(linear_bf16(c.xq, c.ldq, w.q, c.Mq, sb.qkv, inner, false, st));
    const bf16* kvp = c.kv_pre;
    if (kvp == nullptr) {
      LTX2_PROPAGATE(linear_bf16(c.xkv, c.ldkv, w.kv, B * c.Tk, sb.kv, 2 * inner, false, st));
      kvp = sb.kv;
    }
    qp = sb.qkv; kp = kvp; vp = kvp + inner;
    ldq = inner; ldk = 2 * inner;
  }
  if (w.fused_qkv && c.qcos != nullptr && c.qcos == c.kcos) {
    // self-attention: q and k share the token row and the RoPE table -> one pass over the QKV row
    HeadScatter hs = {};
    hs.q[0] = sb.qh; hs.k[0] = sb.kh; hs.v[0] = nullptr;
    hs.heads_per_rank = H; hs.n_total = c.Tq; hs.t_offset = 0;
    LTX2_PROPAGATE(qkv_head_scatter(sb.qkv, 3 * inner, w.qnorm, w.knorm, c.qcos, c.qsin, hs, B, c.Tq, H, Dh, eps, st));
  } else {
    LTX2_PROPAGATE(headnorm_rope(qp, ldq, w.qnorm, c.qcos, c.qsin, sb.qh, B, c.Tq, H, Dh, eps, st));
    if (c.kh_pre == nullptr)
      LTX2_PROPAGATE(headnorm_rope(kp, ldk, w.knorm, c.kcos, c.ksin, sb.kh, B, c.Tk, H, Dh, eps, st));
  }
  const bf16* kh = c.kh_pre != nullptr ? c.kh_pre : sb.kh;
  // V is consumed in place from the projection output (row form, MN-major MMA operand): no transpose pass
  AttnV av;
  av.ptr = vp;
  av.rows = 1;
  av.stride_t = ldk;
  av.stride_h = Dh;
  av.stride_b = int64_t(c.Tk) * ldk;
  const float* gl = nullptr;
  if (w.gate.w != nullptr) {
    // to_gate_logits(x): [Mq, H]; H < 32 is padded by the GEMM's N%32 rule, so heads must be a multiple of 32
    // for the tensor-core path; small head counts (tests) use the row-wise linear instead.
    if (H % 32 == 0) {
      LTX2_PROPAGATE(linear_f32(c.xq, c.ldq, w.gate, c.Mq, sb.gate_logits, H, st));
    } else {
      LTX2_PROPAGATE(rowdot_bf16(c.xq, c.ldq, w.gate.w, w.gate.b, sb.gate_logits, c.Mq, H, w.gate.in, st));
    }
    gl = sb.gate_logits;
  }
  ProfScope ps(PROF_ATTN, 4.0 * B * H * double(c.Tq) * c.Tk * Dh, st);
  return attention_bf16_v(sb.qh, kh, av, sb.attn, B, H, c.Tq, c.Tk, Dh, 1.0f / sqrtf((float)Dh), gl, nullptr, st);
}

}  // namespace
}  // namespace ltx2

// =====================================================================================
// C ABI
// =====================================================================================
extern "C" {

int ltx2_dit_create(const LtxDitConfig* cfg, LtxDit** out) {
  LTX2_REQUIRE(cfg != nullptr && out != nullptr, "dit_create: null argument");
  LTX2_REQUIRE(cfg->attention_head_dim == 64 || cfg->attention_head_dim == 128,
               "dit_create: attention_head_dim must be 64 or 128 (got %d)", cfg->attention_head_dim);
  LTX2_REQUIRE(!cfg->audio_enabled || cfg->audio_head_dim == 64 || cfg->audio_head_dim == 128,
               "dit_create: audio_head_dim must be 64 or 128");
  LTX2_REQUIRE(cfg->num_layers >= 1 && cfg->num_layers <= 64, "dit_create: num_layers must be in 1..64");
  LTX2_REQUIRE(cfg->in_channels % 8 == 0 && cfg->out_channels % 32 == 0, "dit_create: in_channels %% 8, out_channels %% 32");
  LTX2_REQUIRE(!cfg->fp8_linear || ((cfg->num_attention_heads * cfg->attention_head_dim) % 16 == 0 &&
                                    (!cfg->audio_enabled || (cfg->audio_heads * cfg->audio_head_dim) % 16 == 0)),
               "dit_create: fp8_linear needs model widths that are multiples of 16");
  LtxDit* e = new LtxDit();
  e->cfg = *cfg;
  e->D = cfg->num_attention_heads * cfg->attention_head_dim;
  e->Da = cfg->audio_enabled ? cfg->audio_heads * cfg->audio_head_dim : 0;
  e->n_ada = cfg->cross_attention_adaln ? 9 : 6;
  e->cross_attn_scale.assign(cfg->num_layers, NAN);
  Arena dry;
  build_layout(e, &dry, true);
  e->arena_bytes = dry.off + 1024;
  if (cudaMalloc(&e->arena, e->arena_bytes) != cudaSuccess) {
    set_error("weight arena allocation of %zu bytes failed", e->arena_bytes);
    delete e;
    return LTX2_ERR_NOMEM;
  }
  Arena real;
  real.base = reinterpret_cast<uintptr_t>(e->arena);
  build_layout(e, &real, false);
  auto upload = [&](const std::vector<float>& v, float** dst) -> int {
    LTX2_CUDA_CHECK(cudaMalloc(dst, v.size() * 4 + 16));
    LTX2_CUDA_CHECK(cudaMemcpy(*dst, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
    return LTX2_OK;
  };
  std::vector<float> gv = make_freq_grid(cfg->positional_embedding_theta, 3, e->D);
  e->nf_video = (int)gv.size();
  int s = upload(gv, &e->fg_video);
  if (s == LTX2_OK && cudaMalloc(&e->scale1, 16) != cudaSuccess) s = LTX2_ERR_NOMEM;
  if (s == LTX2_OK && cfg->audio_enabled) {
    std::vector<float> ga = make_freq_grid(cfg->positional_embedding_theta, 1, e->Da);
    e->nf_audio = (int)ga.size();
    s = upload(ga, &e->fg_audio);
  }
  if (s != LTX2_OK) {
    ltx2_dit_destroy(e);
    return s;
  }
  *out = e;
  return LTX2_OK;
}

void ltx2_dit_destroy(LtxDit* e) {
  if (!e) return;
  for (auto ev : e->prof.events) cudaEventDestroy(ev);
  if (e->cp.region) {
    for (int r = 0; r < e->cp.world; ++r)
      if (e->cp.opened[r]) cudaIpcCloseMemHandle(e->cp.peer_base[r]);
    cudaFree(e->cp.region);
    if (e->cp.peer_flags_dev) cudaFree(e->cp.peer_flags_dev);
  }
  if (e->arena) cudaFree(e->arena);
  if (e->ws) cudaFree(e->ws);
  if (e->kvc) cudaFree(e->kvc);
  if (e->side) cudaStreamDestroy(e->side);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  if (e->scale1) cudaFree(e->scale1);
  if (e->fg_video) cudaFree(e->fg_video);
  if (e->fg_audio) cudaFree(e->fg_audio);
  delete e;
}

static int set_weight_impl(LtxDit* e, const char* key, const void* data, int32_t dtype, const int64_t* shape,
                           int32_t ndim, float scale, void* stream) {
  LTX2_REQUIRE(e && key && data, "dit_set_weight: null argument");
  auto it = e->slots.find(key);
  if (it == e->slots.end()) {
    set_error("dit_set_weight: unknown key '%s'", key);
    return LTX2_ERR_NOKEY;
  }
  Slot& s = it->second;
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= shape[i];
  const int64_t expect = s.rows * (s.cols ? s.cols : 1);
  const bool shape_ok = (s.cols == 0) ? (n == s.rows)
                                      : (ndim == 2 && shape[0] == s.rows && shape[1] == s.cols);
  if (!shape_ok) {
    set_error("dit_set_weight: '%s' expects [%lld,%lld], got %lld elements (ndim %d)", key, (long long)s.rows,
              (long long)s.cols, (long long)n, ndim);
    return LTX2_ERR_INVALID;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int r;
  if (dtype == LTX2_F8E4M3) {
    // an FP8 checkpoint tensor: value = e4m3 * weight_scale (loader/fp8_loader.py:14-32)
    LTX2_REQUIRE(s.cols != 0, "dit_set_weight: '%s' is not a matrix; FP8 data is only accepted for Linear weights", key);
    if (s.storage == LTX2_F8E4M3) {
      // kept quantised: the checkpoint's own bytes feed the FP8 MMA, the scale goes to the epilogue
      LTX2_CUDA_CHECK(cudaMemcpyAsync(s.dst, data, size_t(expect), cudaMemcpyDeviceToDevice, st));
      r = fill_f32(s.cscale, scale, s.rows, st);
    } else if (s.storage == LTX2_BF16) {
      LTX2_PROPAGATE(fill_f32(e->scale1, scale, 1, st));
      r = dequant_e4m3(data, e->scale1, 0, s.rows, s.cols, s.dst, LTX2_BF16, st);
    } else {
      set_error("dit_set_weight: '%s' is stored as fp32; FP8 data not accepted", key);
      return LTX2_ERR_INVALID;
    }
  } else {
    LTX2_REQUIRE(scale == 1.0f, "dit_set_weight: a weight_scale is only meaningful for FP8 data");
    if (s.storage == LTX2_F8E4M3) r = quantize_rows_e4m3(data, dtype, s.rows, s.cols, s.dst, s.cscale, st);
    else if (s.storage == LTX2_BF16) r = cast_to_bf16(data, dtype, s.dst, expect, st);
    else r = cast_to_f32(data, dtype, reinterpret_cast<float*>(s.dst), expect, st);
  }
  if (r == LTX2_OK) s.loaded = true;
  e->cached_tag = 0;          // cached context K/V were projected with the old weights (LoRA fuse / restore)
  e->coef_dirty = true;
  return r;
}

int ltx2_dit_set_weight(LtxDit* e, const char* key, const void* data, int32_t dtype, const int64_t* shape,
                        int32_t ndim, void* stream) {
  return set_weight_impl(e, key, data, dtype, shape, ndim, 1.0f, stream);
}

int ltx2_dit_set_weight_scaled(LtxDit* e, const char* key, const void* data, int32_t dtype, const int64_t* shape,
                               int32_t ndim, float weight_scale, void* stream) {
  return set_weight_impl(e, key, data, dtype, shape, ndim, weight_scale, stream);
}

// read a tensor back (engine storage -> dst_dtype); used by the LoRA fuse/restore flows that read
// velocity_model.parameters() (scripts/generate.py:1198-1200, pipelines/two_stage.py:180-186)
int ltx2_dit_get_weight(LtxDit* e, const char* key, void* dst, int32_t dst_dtype, int64_t n, void* stream) {
  LTX2_REQUIRE(e && key && dst, "dit_get_weight: null argument");
  auto it = e->slots.find(key);
  if (it == e->slots.end()) {
    set_error("dit_get_weight: unknown key '%s'", key);
    return LTX2_ERR_NOKEY;
  }
  const Slot& s = it->second;
  const int64_t expect = s.rows * (s.cols ? s.cols : 1);
  LTX2_REQUIRE(n == expect, "dit_get_weight: '%s' has %lld elements, buffer holds %lld", key, (long long)expect, (long long)n);
  LTX2_REQUIRE(s.loaded, "dit_get_weight: '%s' has not been set", key);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (s.storage == LTX2_F8E4M3) return dequant_e4m3(s.dst, s.cscale, 1, s.rows, s.cols, dst, dst_dtype, st);
  if (dst_dtype == LTX2_F32) return cast_to_f32(s.dst, s.storage, reinterpret_cast<float*>(dst), n, st);
  if (dst_dtype == LTX2_BF16) return cast_to_bf16(s.dst, s.storage, dst, n, st);
  set_error("dit_get_weight: destination dtype %d unsupported", dst_dtype);
  return LTX2_ERR_INVALID;
}

// shape of a weight slot: returns ndim (1 or 2) and fills shape_out[2]; negative status on unknown key
int ltx2_dit_weight_shape(LtxDit* e, const char* key, int64_t* shape_out) {
  LTX2_REQUIRE(e && key && shape_out, "dit_weight_shape: null argument");
  auto it = e->slots.find(key);
  if (it == e->slots.end()) return LTX2_ERR_NOKEY;
  shape_out[0] = it->second.rows;
  shape_out[1] = it->second.cols;
  return it->second.cols ? 2 : 1;
}

// newline-separated list of every weight key of this configuration
int64_t ltx2_dit_weight_keys(LtxDit* e, char* names_out, int64_t names_cap) {
  std::string acc;
  for (auto& kv : e->slots) acc += kv.first + "\n";
  if (names_out && names_cap > 0) {
    strncpy(names_out, acc.c_str(), names_cap - 1);
    names_out[names_cap - 1] = 0;
  }
  return static_cast<int64_t>(acc.size()) + 1;
}

int ltx2_dit_missing_weights(LtxDit* e, char* names_out, int64_t names_cap) {
  int missing = 0;
  std::string acc;
  for (auto& kv : e->slots)
    if (!kv.second.loaded) {
      ++missing;
      if (acc.size() < 4000) acc += kv.first + "\n";
    }
  if (names_out && names_cap > 0) {
    strncpy(names_out, acc.c_str(), names_cap - 1);
    names_out[names_cap - 1] = 0;
  }
  return missing;
}

int ltx2_dit_set_cross_attn_scale(LtxDit* e, int32_t block, float scale) {
  LTX2_REQUIRE(e && block >= 0 && block < e->cfg.num_layers, "set_cross_attn_scale: bad block %d", block);
  e->cross_attn_scale[block] = scale;
  return LTX2_OK;
}

}  // extern "C"

namespace ltx2 {
namespace {

// Modulation classes = distinct (batch, sigma) pairs.  Scalar timesteps (n_t == 1): one class per batch element,
// no host round trip.  Per-token timesteps (image conditioning, common.py:193-203): the B*N values are read back
// once and de-duplicated on the host (typically 2 classes per batch element).
int prepare_classes(LtxDit* e, const LtxModalityView& m, int* n_cls, std::vector<float>& cls_vals,
                    std::vector<int>& row_cls_host, cudaStream_t st) {
  const int B = m.batch, N = m.tokens;
  cls_vals.clear();
  row_cls_host.clear();
  if (m.n_cls > 0) {         // the caller supplies the classes: nothing to read back
    *n_cls = m.n_cls;
    return LTX2_OK;
  }
  if (m.n_t == 1) {
    *n_cls = B;
    return LTX2_OK;
  }
  e->h_ts.resize(size_t(B) * N);
  LTX2_CUDA_CHECK(cudaMemcpyAsync(e->h_ts.data(), m.timesteps, e->h_ts.size() * 4, cudaMemcpyDeviceToHost, st));
  LTX2_CUDA_CHECK(cudaStreamSynchronize(st));
  row_cls_host.resize(size_t(B) * N);
  for (int b = 0; b < B; ++b) {
    std::map<uint32_t, int> seen;
    for (int t = 0; t < N; ++t) {
      const float v = e->h_ts[size_t(b) * N + t];
      uint32_t bits;
      memcpy(&bits, &v, 4);
      auto it = seen.find(bits);
      if (it == seen.end()) {
        it = seen.emplace(bits, (int)cls_vals.size()).first;
        cls_vals.push_back(v);
      }
      row_cls_host[size_t(b) * N + t] = it->second;
    }
  }
  *n_cls = (int)cls_vals.size();
  return LTX2_OK;
}

struct Prepared {
  int n_cls;
};

int upload_classes(StreamBuf& sb, const LtxModalityView& m, const std::vector<float>& cls_vals,
                   const std::vector<int>& row_cls_host, cudaStream_t st) {
  const int M = m.batch * m.tokens;
  if (m.n_cls > 0) {
    LTX2_CUDA_CHECK(cudaMemcpyAsync(sb.t_cls, m.timesteps, size_t(m.n_cls) * 4, cudaMemcpyDeviceToDevice, st));
  } else if (m.n_t == 1) {
    LTX2_CUDA_CHECK(cudaMemcpyAsync(sb.t_cls, m.timesteps, size_t(m.batch) * 4, cudaMemcpyDeviceToDevice, st));
  } else {
    LTX2_CUDA_CHECK(cudaMemcpyAsync(sb.t_cls, cls_vals.data(), cls_vals.size() * 4, cudaMemcpyHostToDevice, st));
  }
  fill_row_index_kernel<<<(std::max(M, m.batch * m.context_tokens) + 255) / 256, 256, 0, st>>>(
      sb.row_batch, std::max(M, m.batch * m.context_tokens), m.tokens);
  fill_row_index_kernel<<<(m.batch * m.context_tokens + 255) / 256, 256, 0, st>>>(sb.ctx_batch,
                                                                                 m.batch * m.context_tokens,
                                                                                 m.context_tokens);
  if (m.n_cls > 0) {
    LTX2_CUDA_CHECK(cudaMemcpyAsync(sb.row_cls, m.row_cls, size_t(M) * 4, cudaMemcpyDeviceToDevice, st));
  } else if (m.n_t == 1) {
    LTX2_CUDA_CHECK(cudaMemcpyAsync(sb.row_cls, sb.row_batch, size_t(M) * 4, cudaMemcpyDeviceToDevice, st));
  } else {
    LTX2_CUDA_CHECK(cudaMemcpyAsync(sb.row_cls, row_cls_host.data(), size_t(M) * 4, cudaMemcpyHostToDevice, st));
  }
  gather_f32_kernel<<<(M + 255) / 256, 256, 0, st>>>(sb.t_cls, sb.row_cls, sb.t_row, M);
  LTX2_CUDA_CHECK(cudaGetLastError());
  return LTX2_OK;
}

// TransformerArgsPreprocessor.prepare for one modality
int prepare_stream(LtxDit* e, StreamBuf& sb, const StreamW& w, const LtxModalityView& m, bool audio_side,
                   float* table_arena, float* ptable_arena, cudaStream_t st) {
  const LtxDitConfig& c = e->cfg;
  const int dim = sb.dim, B = m.batch, N = m.tokens, S = m.context_tokens, M = B * N, L = c.num_layers;
  const int n = e->n_ada;
  // patchify projection -> fp32 residual stream
  LTX2_PROPAGATE(cast_to_bf16(m.latent, m.latent_dtype, sb.lat, int64_t(M) * w.patchify.in, st));
  LTX2_PROPAGATE(linear_f32(sb.lat, w.patchify.in, w.patchify, M, sb.x, dim, st));
  // timestep adaLN -> per-layer modulation tables
  float* emb = e->scratch + size_t(64) * (256 + 2 * dim);      // [n_cls, n*dim] (after run_adaln's own scratch)
  LTX2_PROPAGATE(run_adaln(e, w.adaln, sb.t_cls, sb.n_cls, c.timestep_scale_multiplier, dim, emb, sb.emb_t, st));
  LTX2_PROPAGATE(build_modulation_ex(table_arena, int64_t(n) * dim, emb, int64_t(n) * dim, dim, sb.mod,
                                     int64_t(sb.n_cls) * n * dim, int64_t(n) * dim, L, sb.n_cls, n, dim, st));
  LTX2_PROPAGATE(build_modulation_ex(w.head_table, 0, sb.emb_t, dim, 0, sb.head_mod, 0, int64_t(2) * dim, 1,
                                     sb.n_cls, 2, dim, st));
  if (c.cross_attention_adaln) {
    // prompt adaLN uses the scalar sigma per batch element (model.py:250-260)
    float* sig = e->scratch + size_t(64) * (256 + 11 * dim);
    if (m.sigma != nullptr) {
      LTX2_CUDA_CHECK(cudaMemcpyAsync(sig, m.sigma, size_t(B) * 4, cudaMemcpyDeviceToDevice, st));
    } else {
      // first token's timestep of every batch element
      LTX2_REQUIRE(m.n_cls == 0, "timestep classes need Modality.sigma for the prompt adaLN");
      LTX2_CUDA_CHECK(cudaMemcpy2DAsync(sig, 4, m.timesteps, size_t(m.n_t) * 4, 4, B, cudaMemcpyDeviceToDevice, st));
    }
    float* pemb = emb;   // reuse: [B, 2*dim]
    LTX2_PROPAGATE(run_adaln(e, w.prompt_adaln, sig, B, c.timestep_scale_multiplier, dim, pemb, nullptr, st));
    LTX2_PROPAGATE(build_modulation_ex(ptable_arena, int64_t(2) * dim, pemb, int64_t(2) * dim, dim, sb.prompt_mod,
                                       int64_t(B) * 2 * dim, int64_t(2) * dim, L, B, 2, dim, st));
  }
  // context (caption projection for V1); nothing to do when the projected K/V of this context are cached
  const int ctx_ch = w.has_caption ? w.cap1.in : dim;
  const bool ctx_cached = !audio_side && e->ctx_cache_on && e->ctx_cache_hit;
  if (!ctx_cached)
    LTX2_PROPAGATE(cast_to_bf16(m.context, m.context_dtype, w.has_caption ? sb.ctx_in : sb.ctx,
                                int64_t(B) * S * ctx_ch, st));
  if (w.has_caption && !ctx_cached) {
    LTX2_PROPAGATE(linear_bf16(sb.ctx_in, ctx_ch, w.cap1, B * S, sb.ctx_mid, dim, true, st));
    LTX2_PROPAGATE(linear_bf16(sb.ctx_mid, dim, w.cap2, B * S, sb.ctx, dim, false, st));
  }
  // RoPE tables
  if (!audio_side) {
    LTX2_PROPAGATE(rope_tables_dev(m.positions, B, 3, 3, N, dim, c.max_pos, e->fg_video, e->nf_video, sb.cos, sb.sin, st));
  } else {
    const float mp[1] = {c.audio_max_pos};
    LTX2_PROPAGATE(rope_tables_dev(m.positions, B, 1, 1, N, dim, mp, e->fg_audio, e->nf_audio, sb.cos, sb.sin, st));
  }
  return LTX2_OK;
}

}  // namespace
}  // namespace ltx2

namespace ltx2 {
namespace {

// scalar sigma per batch element of a modality (Modality.sigma, else timesteps[:,0]) -> dst[B] (device)
int scalar_sigma(const LtxModalityView& m, float* dst, cudaStream_t st) {
  if (m.sigma != nullptr) {
    LTX2_CUDA_CHECK(cudaMemcpyAsync(dst, m.sigma, size_t(m.batch) * 4, cudaMemcpyDeviceToDevice, st));
  } else {
    LTX2_REQUIRE(m.n_cls == 0, "timestep classes need Modality.sigma for the cross-modal adaLN");
    LTX2_CUDA_CHECK(cudaMemcpy2DAsync(dst, 4, m.timesteps, size_t(m.n_t) * 4, 4, m.batch, cudaMemcpyDeviceToDevice, st));
  }
  return LTX2_OK;
}

// cross-modal adaLN (model.py:345-366): ca_mod[l, b, 0..3] = table[l][0..3] + ss_emb[b], [4] = table[l][4] + gate_emb[b]
int prepare_cross_mod(LtxDit* e, StreamBuf& sb, const AdaLNW& ss, const AdaLNW& gate, const float* table_arena,
                      const LtxModalityView& other, cudaStream_t st) {
  const LtxDitConfig& c = e->cfg;
  const int dim = sb.dim, B = sb.B, L = c.num_layers;
  float* sig = e->scratch + size_t(64) * (256 + 11 * std::max(e->D, 1));
  float* emb = e->scratch + size_t(64) * (256 + 2 * std::max(e->D, 1));
  LTX2_PROPAGATE(scalar_sigma(other, sig, st));
  LTX2_PROPAGATE(run_adaln(e, ss, sig, B, c.timestep_scale_multiplier, dim, emb, nullptr, st));
  LTX2_PROPAGATE(build_modulation_ex(table_arena, int64_t(5) * dim, emb, int64_t(4) * dim, dim, sb.ca_mod,
                                     int64_t(B) * 5 * dim, int64_t(5) * dim, L, B, 4, dim, st));
  // gate timestep: sigma * ts_mult * (av_ca_mult / ts_mult)
  LTX2_PROPAGATE(run_adaln(e, gate, sig, B, c.av_ca_timestep_scale_multiplier, dim, emb, nullptr, st));
  LTX2_PROPAGATE(build_modulation_ex(table_arena + 4 * dim, int64_t(5) * dim, emb, dim, dim, sb.ca_mod + 4 * dim,
                                     int64_t(B) * 5 * dim, int64_t(5) * dim, L, B, 1, dim, st));
  return LTX2_OK;
}

// RMSNorm + modulation of the residual stream as the E4M3 input of an FP8 linear: sb.xq / sb.xs (and the bf16 copy in
// `out16` when a bf16 consumer of the same rows exists, e.g. the gate-logit projection)
int rms_mod_q8(const LtxDit* e, const StreamBuf& sb, bf16* out16, const float* mod, int64_t mod_stride, int shift_row,
               int scale_row, const int* cls, cudaStream_t st, float* row_l2 = nullptr) {
  const int M = sb.B * sb.N;
  return norm_modulate_q8(sb.x, 0, sb.dim, sb.xq, sb.dim, sb.xs, out16, sb.dim, M, sb.dim, NORM_RMS, e->cfg.norm_eps, mod,
                          mod_stride, int64_t(shift_row) * sb.dim, int64_t(scale_row) * sb.dim, cls, st, row_l2);
}

int rms_mod(const LtxDit* e, const StreamBuf& sb, bf16* out, const float* mod, int64_t mod_stride, int shift_row,
            int scale_row, const int* cls, cudaStream_t st) {
  const int M = sb.B * sb.N;
  return norm_modulate(sb.x, 0, sb.dim, out, sb.dim, M, sb.dim, NORM_RMS, e->cfg.norm_eps, mod, mod_stride,
                       int64_t(shift_row) * sb.dim, int64_t(scale_row) * sb.dim, cls, st);
}

// self-attention + text cross-attention of one modality (transformer.py:503-553)
int run_self_and_text(LtxDit* e, StreamBuf& sb, const BlockStreamW& w, int layer, bool skip_self, float ca_scale,
                      cudaStream_t st) {
  const LtxDitConfig& c = e->cfg;
  const int dim = sb.dim, B = sb.B, N = sb.N, S = sb.S, M = B * N, n = e->n_ada;
  const bool v2 = c.cross_attention_adaln != 0;
  const float* mod = sb.mod + size_t(layer) * sb.n_cls * n * dim;
  const int64_t ms = int64_t(n) * dim;
  if (!skip_self && e->cp.world > 1 && &sb == &e->vb) {
    // ---- context-parallel self-attention: tokens are sharded, heads are re-sharded for the attention ----
    CpState& cp = e->cp;
    const AttnW& aw = w.attn1;
    const int H = aw.heads, Dh = aw.dh, inner = aw.inner, Hl = cp.heads_local, Nt = cp.n_total;
    if (aw.q.w8 != nullptr) {
      LTX2_PROPAGATE(rms_mod_q8(e, sb, aw.gate.w != nullptr ? sb.xn : nullptr, mod, ms, 0, 1, sb.row_cls, st));
      LTX2_PROPAGATE(linear_q8(sb.xq, sb.xs, dim, aw.q, M, sb.qkv, 3 * inner, false, st, 3 * inner, 0));
    } else {
      LTX2_PROPAGATE(rms_mod(e, sb, sb.xn, mod, ms, 0, 1, sb.row_cls, st));
      LTX2_PROPAGATE(linear_bf16(sb.xn, dim, aw.q, M, sb.qkv, 3 * inner, false, st, 3 * inner, 0));
    }
    HeadScatter hs = {};
    for (int r = 0; r < cp.world; ++r) {
      hs.q[r] = reinterpret_cast<bf16*>(cp.peer_base[r] + cp.off_q);
      hs.k[r] = reinterpret_cast<bf16*>(cp.peer_base[r] + cp.off_k);
      hs.v[r] = reinterpret_cast<bf16*>(cp.peer_base[r] + cp.off_v);
    }
    hs.heads_per_rank = Hl; hs.n_total = Nt; hs.t_offset = cp.rank * cp.n_local;
    // q/k-norm + RoPE + the token->head all-to-all in one kernel: every head is stored into its owner's buffers
    LTX2_PROPAGATE(qkv_head_scatter(sb.qkv, 3 * inner, aw.qnorm, aw.knorm, sb.cos, sb.sin, hs, B, N, H, Dh,
                                    e->cfg.norm_eps, st));
    const float* gl = nullptr;
    if (aw.gate.w != nullptr) {
      // per-head gate logits of my token rows (attention.py:243-250), sent to the ranks that own each head
      if (H % 32 == 0) {
        LTX2_PROPAGATE(linear_f32(sb.xn, dim, aw.gate, M, sb.gate_logits, H, st));
      } else {
        LTX2_PROPAGATE(rowdot_bf16(sb.xn, dim, aw.gate.w, aw.gate.b, sb.gate_logits, M, H, aw.gate.in, st));
      }
      float* peers[kMaxCpRanks];
      for (int r = 0; r < cp.world; ++r) peers[r] = reinterpret_cast<float*>(cp.peer_base[r] + cp.off_gate);
      LTX2_PROPAGATE(gate_scatter(sb.gate_logits, peers, B, N, H, Hl, Nt, cp.rank * cp.n_local, st));
      gl = reinterpret_cast<const float*>(cp.region + cp.off_gate);
    }
    if (cp.ctx_tokens == S && !(e->ctx_cache_on && e->ctx_cache_hit)) {
      // text-context K/V for this block: each rank projects S/P context rows and stores the result into EVERY rank's
      // buffer (peer-memory broadcast) instead of all ranks repeating the full projection.  V2 models project the
      // sigma-modulated context (transformer.py:449-452), which is formed here, before the projection.
      const int Sl = S / cp.world;
      const AttnW& cw = w.attn2;
      const bf16* ctx_src = sb.ctx;
      if (v2) {
        const float* pm = sb.prompt_mod + size_t(layer) * B * 2 * dim;
        LTX2_PROPAGATE(norm_modulate(sb.ctx, 1, dim, sb.ctx_mod, dim, B * S, dim, NORM_NONE, c.norm_eps, pm,
                                     int64_t(2) * dim, 0, dim, sb.ctx_batch, st));
        ctx_src = sb.ctx_mod;
      }
      for (int b = 0; b < B; ++b) {
        const size_t row0 = size_t(b) * S + size_t(cp.rank) * Sl;
        const size_t byte0 = cp.off_ckv[layer & 1] + row0 * 2 * cw.inner * sizeof(bf16);
        bf16* mine = reinterpret_cast<bf16*>(cp.region + byte0);
        LTX2_PROPAGATE(linear_bf16(ctx_src + row0 * dim, dim, cw.kv, Sl, mine, 2 * cw.inner, false, st));
        void* peers[kMaxCpRanks];
        int np = 0;
        for (int r = 0; r < cp.world; ++r)
          if (r != cp.rank) peers[np++] = cp.peer_base[r] + byte0;
        LTX2_PROPAGATE(peer_broadcast(mine, peers, np, int64_t(Sl) * 2 * cw.inner * sizeof(bf16), st));
      }
    }
    uint32_t* my_flags = reinterpret_cast<uint32_t*>(cp.region + cp.off_flags);
    LTX2_PROPAGATE(cp_barrier(cp.peer_flags_dev, my_flags, cp.rank, cp.world, ++cp.epoch, st));
    AttnV av;
    av.ptr = cp.region + cp.off_v; av.rows = 1;
    av.stride_t = Dh; av.stride_h = int64_t(Nt) * Dh; av.stride_b = int64_t(Hl) * Nt * Dh;
    AttnOutScatter sc;
    for (int r = 0; r < cp.world; ++r) sc.peer[r] = reinterpret_cast<bf16*>(cp.peer_base[r] + cp.off_o);
    sc.rows_per_rank = cp.n_local; sc.pitch = inner; sc.head0 = cp.rank * Hl;
    {
      ProfScope ps(PROF_ATTN, 4.0 * B * Hl * double(Nt) * Nt * Dh, st);
      // attention over all tokens for my heads; the epilogue stores each row into the token owner's buffer
      LTX2_PROPAGATE(attention_bf16_v(cp.region + cp.off_q, cp.region + cp.off_k, av, nullptr, B, Hl, Nt, Nt, Dh,
                                      1.0f / sqrtf((float)Dh), gl, nullptr, st, nullptr, &sc));
    }
    LTX2_PROPAGATE(cp_barrier(cp.peer_flags_dev, my_flags, cp.rank, cp.world, ++cp.epoch, st));
    LTX2_PROPAGATE(linear_residual(reinterpret_cast<const bf16*>(cp.region + cp.off_o), inner, aw.o, M, sb.x, dim,
                                   mod + 2 * dim, ms, sb.row_cls, 1.0f, st));
  } else if (skip_self && e->cp.world > 1 && &sb == &e->vb) {
    // STG skip on sharded ranks: keep one barrier per block so the double-buffered exchange areas stay ordered
    LTX2_PROPAGATE(cp_barrier(e->cp.peer_flags_dev, reinterpret_cast<uint32_t*>(e->cp.region + e->cp.off_flags),
                              e->cp.rank, e->cp.world, ++e->cp.epoch, st));
  } else if (!skip_self) {
    AttnCall a{&w.attn1, sb.xn, dim, M, N, sb.xn, dim, N, sb.cos, sb.sin, sb.cos, sb.sin};
    if (w.attn1.q.w8 != nullptr) {
      LTX2_PROPAGATE(rms_mod_q8(e, sb, w.attn1.gate.w != nullptr ? sb.xn : nullptr, mod, ms, 0, 1, sb.row_cls, st));
      a.xq8 = sb.xq;
      a.xq8_scale = sb.xs;
    } else {
      LTX2_PROPAGATE(rms_mod(e, sb, sb.xn, mod, ms, 0, 1, sb.row_cls, st));
    }
    LTX2_PROPAGATE(run_attention_core(e, sb, a, B, st));
    LTX2_PROPAGATE(linear_residual(sb.attn, w.attn1.inner, w.attn1.o, M, sb.x, dim, mod + 2 * dim, ms, sb.row_cls,
                                   1.0f, st));
  }
  const bf16* ctx = sb.ctx;
  const bool tq8 = w.attn2.q.w8 != nullptr;
  bf16* txn = (!tq8 || w.attn2.gate.w != nullptr) ? sb.xn : nullptr;      // bf16 rows only if somebody reads them
  if (v2) {
    if (tq8) LTX2_PROPAGATE(rms_mod_q8(e, sb, txn, mod, ms, 6, 7, sb.row_cls, st));
    else LTX2_PROPAGATE(rms_mod(e, sb, sb.xn, mod, ms, 6, 7, sb.row_cls, st));
    const bool kv_from_peers = e->cp.world > 1 && &sb == &e->vb && e->cp.ctx_tokens == S && !skip_self;
    if (!kv_from_peers) {       // (context-parallel ranks already formed and projected their share of it above)
      const float* pm = sb.prompt_mod + size_t(layer) * B * 2 * dim;
      LTX2_PROPAGATE(norm_modulate(sb.ctx, 1, dim, sb.ctx_mod, dim, B * S, dim, NORM_NONE, c.norm_eps, pm,
                                   int64_t(2) * dim, 0, dim, sb.ctx_batch, st));
    }
    ctx = sb.ctx_mod;
  } else {
    if (tq8) LTX2_PROPAGATE(rms_mod_q8(e, sb, txn, nullptr, 0, 0, 0, nullptr, st));
    else LTX2_PROPAGATE(rms_mod(e, sb, sb.xn, nullptr, 0, 0, 0, nullptr, st));
  }
  AttnCall a{&w.attn2, sb.xn, dim, M, N, ctx, dim, S, nullptr, nullptr, nullptr, nullptr};
  if (tq8) {
    a.xq8 = sb.xq;
    a.xq8_scale = sb.xs;
  }
  const bool region_kv = e->cp.world > 1 && &sb == &e->vb && e->cp.ctx_tokens == S && !skip_self;
  if (&sb == &e->vb && e->ctx_cache_on) {
    // V1: the context K/V of this block do not depend on sigma -- computed by the first forward of a sample, reused after
    bf16* kvl = e->kvc_kv(layer, B, S);
    bf16* khl = e->kvc_kh(layer, B, S);
    const AttnW& cw = w.attn2;
    if (!e->ctx_cache_hit) {
      if (region_kv) {
        LTX2_CUDA_CHECK(cudaMemcpyAsync(kvl, e->cp.region + e->cp.off_ckv[layer & 1],
                                        size_t(B) * S * 2 * cw.inner * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
      } else {
        LTX2_PROPAGATE(linear_bf16(ctx, dim, cw.kv, B * S, kvl, 2 * cw.inner, false, st));
      }
      LTX2_PROPAGATE(headnorm_rope(kvl, 2 * cw.inner, cw.knorm, nullptr, nullptr, khl, B, S, cw.heads, cw.dh,
                                   c.norm_eps, st));
    }
    a.kv_pre = kvl;
    a.kh_pre = khl;
  } else if (region_kv) {
    a.kv_pre = reinterpret_cast<const bf16*>(e->cp.region + e->cp.off_ckv[layer & 1]);
  }
  LTX2_PROPAGATE(run_attention_core(e, sb, a, B, st));
  return linear_residual(sb.attn, w.attn2.inner, w.attn2.o, M, sb.x, dim, v2 ? mod + 8 * dim : nullptr, ms,
                         sb.row_cls, ca_scale, st);
}

int run_ffn(LtxDit* e, StreamBuf& sb, const BlockStreamW& w, int layer, cudaStream_t st) {
  const int dim = sb.dim, M = sb.B * sb.N, n = e->n_ada;
  const float* mod = sb.mod + size_t(layer) * sb.n_cls * n * dim;
  const int64_t ms = int64_t(n) * dim;
  if (w.ff1.w8 != nullptr) {
    // FP8 FFN: norm -> E4M3 rows (+ their L2 norms) -> up-projection whose epilogue writes gelu(.) as E4M3 against the
    // Cauchy-Schwarz bound |row|_2 max|w_j|_2 + max|b| (no second pass over the 4D-wide hidden) -> FP8 down-projection
    // with the gated residual epilogue
    LTX2_PROPAGATE(rms_mod_q8(e, sb, nullptr, mod, ms, 3, 4, sb.row_cls, st, sb.xl2));
    {
      GemmEpilogue ep;
      ep.mode = GEMM_EPI_E4M3_GELU;
      ep.bias = w.ff1.b;
      ep.out = sb.hq;
      ep.ldo = 4 * dim;
      ep.row_scale = sb.xs;
      ep.col_scale = w.ff1.cscale;
      ep.out_l2 = sb.xl2;
      ep.out_coef = w.ff_coef;
      ProfScope ps(PROF_GEMM8, 2.0 * M * double(4 * dim) * dim, st);
      LTX2_PROPAGATE(gemm_e4m3(sb.xq, dim, w.ff1.w8, dim, M, 4 * dim, dim, ep, st));
    }
    GemmEpilogue ep;
    ep.mode = GEMM_EPI_F32_RESIDUAL;
    ep.bias = w.ff2.b;
    ep.out = sb.x;
    ep.ldo = dim;
    ep.gate = mod + 5 * dim;
    ep.gate_stride = ms;
    ep.row_cls = sb.row_cls;
    ep.alpha = 1.0f;
    ep.max_splits = g_split_k;
    ep.row_scale = sb.xl2;
    ep.row_coef = w.ff_coef;
    ep.col_scale = w.ff2.cscale;
    ProfScope ps(PROF_GEMM8, 2.0 * M * double(dim) * 4 * dim, st);
    return gemm_e4m3(sb.hq, 4 * dim, w.ff2.w8, 4 * dim, M, dim, 4 * dim, ep, st);
  } else {
    LTX2_PROPAGATE(rms_mod(e, sb, sb.xn, mod, ms, 3, 4, sb.row_cls, st));
    LTX2_PROPAGATE(linear_bf16(sb.xn, dim, w.ff1, M, sb.hidden, 4 * dim, true, st));
  }
  return linear_residual(sb.hidden, 4 * dim, w.ff2, M, sb.x, dim, mod + 5 * dim, ms, sb.row_cls, 1.0f, st);
}

int run_head(LtxDit* e, StreamBuf& sb, const StreamW& w, const LtxModalityView& m, int x0, float* out,
             cudaStream_t st) {
  const int dim = sb.dim, M = sb.B * sb.N, C = w.proj_out.out;
  LTX2_PROPAGATE(norm_modulate(sb.x, 0, dim, sb.xn, dim, M, dim, NORM_LAYER, e->cfg.norm_eps, sb.head_mod,
                               int64_t(2) * dim, 0, dim, sb.row_cls, st));
  if (!x0) return linear_f32(sb.xn, dim, w.proj_out, M, out, C, st);
  LTX2_REQUIRE(C == w.patchify.in, "x0 output needs in_channels == out_channels");
  LTX2_PROPAGATE(linear_f32(sb.xn, dim, w.proj_out, M, sb.vel, C, st));
  float* lat32 = sb.x;   // the residual stream is dead after the head norm
  LTX2_PROPAGATE(cast_to_f32(m.latent, m.latent_dtype, lat32, int64_t(M) * C, st));
  return x0_from_velocity(lat32, sb.vel, sb.t_row, out, M, C, st);
}

int check_view(const LtxModalityView* m, const char* name, int n_dims) {
  LTX2_REQUIRE(m->latent && m->context && m->timesteps && m->positions, "%s modality: null pointer", name);
  LTX2_REQUIRE(m->batch >= 1 && m->tokens >= 1 && m->context_tokens >= 1, "%s modality: empty", name);
  if (m->n_cls > 0) {
    LTX2_REQUIRE(m->row_cls != nullptr && m->n_cls <= 64, "%s modality: row_cls is null or n_cls > 64", name);
  } else {
    LTX2_REQUIRE(m->n_t == 1 || m->n_t == m->tokens, "%s modality: timesteps must be (B,) or (B,N)", name);
  }
  LTX2_REQUIRE(m->n_dims == n_dims, "%s modality: positions must have %d axes (got %d)", name, n_dims, m->n_dims);
  return LTX2_OK;
}

}  // namespace
}  // namespace ltx2

extern "C" int ltx2_dit_forward(LtxDit* e, const LtxModalityView* video, const LtxModalityView* audio,
                                const LtxDitSkip* skip, int32_t x0, float* out_video, float* out_audio,
                                void* stream) {
  LTX2_REQUIRE(e != nullptr, "dit_forward: null handle");
  if (video == nullptr) {
    set_error("Video modality required for video-enabled model");   // model.py:824
    return LTX2_ERR_INVALID;
  }
  LTX2_REQUIRE(out_video != nullptr, "dit_forward: null output");
  {
    int missing = 0;
    for (auto& kv : e->slots) missing += kv.second.loaded ? 0 : 1;
    if (missing) {
      set_error("dit_forward: %d weight tensors have not been set", missing);
      return LTX2_ERR_STATE;
    }
  }
  const LtxDitConfig& c = e->cfg;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LTX2_PROPAGATE(check_view(video, "video", 3));
  const bool has_audio = c.audio_enabled && audio != nullptr && audio->tokens > 0;
  if (has_audio) {
    LTX2_PROPAGATE(check_view(audio, "audio", 1));
    LTX2_REQUIRE(audio->batch == video->batch, "audio/video batch mismatch");
    LTX2_REQUIRE(out_audio != nullptr, "dit_forward: audio modality given but out_audio is null");
  }
  if (e->cp.world > 1) {
    LTX2_REQUIRE(e->cp.connected, "context parallel: ltx2_dit_cp_connect has not been called");
    LTX2_REQUIRE(video->tokens == e->cp.n_local && video->batch == e->cp.B,
                 "context parallel: expected the local slice of %d tokens (batch %d), got %d (batch %d)",
                 e->cp.n_local, e->cp.B, video->tokens, video->batch);
  }
  if (c.fp8_linear && e->coef_dirty) {
    for (auto& bw : e->blocks) {
      LTX2_PROPAGATE(e4m3_bound_coef(bw.v.ff1.w8, bw.v.ff1.cscale, bw.v.ff1.b, bw.v.ff1.out, bw.v.ff1.in, bw.v.ff_coef, st));
      if (c.audio_enabled)
        LTX2_PROPAGATE(e4m3_bound_coef(bw.a.ff1.w8, bw.a.ff1.cscale, bw.a.ff1.b, bw.a.ff1.out, bw.a.ff1.in, bw.a.ff_coef, st));
    }
    e->coef_dirty = false;
  }
  LtxDitSkip sk = {0, 0, 0, 0};
  if (skip) sk = *skip;
  g_prof = &e->prof;
  g_split_k = e->cp.world > 1 ? e->cp.split_k : 1;
  e->prof.used = 0;
  e->prof.recs.clear();

  std::vector<float> cls_v, cls_a;
  std::vector<int> rc_v, rc_a;
  int ncv = 0, nca = 0;
  LTX2_PROPAGATE(prepare_classes(e, *video, &ncv, cls_v, rc_v, st));
  if (has_audio) LTX2_PROPAGATE(prepare_classes(e, *audio, &nca, cls_a, rc_a, st));
  LTX2_REQUIRE(ncv <= 64 && nca <= 64, "more than 64 distinct (batch, sigma) classes");
  Shapes sh;
  sh.B = video->batch; sh.N = video->tokens; sh.S = video->context_tokens; sh.n_cls = ncv;
  sh.Na = has_audio ? audio->tokens : 0; sh.Sa = has_audio ? audio->context_tokens : 0;
  sh.n_cls_a = has_audio ? nca : 0;
  LTX2_PROPAGATE(ensure_workspace(e, sh));
  // text-context reuse (ltx2_dit_set_context_tag)
  const int n_layers = (e->layer_limit > 0 && e->layer_limit < c.num_layers) ? e->layer_limit : c.num_layers;
  e->ctx_cache_on = e->ctx_tag != 0 && !c.cross_attention_adaln && n_layers == c.num_layers;
  e->ctx_cache_hit = false;
  if (e->ctx_cache_on) {
    const size_t need = size_t(c.num_layers) * 3 * size_t(sh.B) * sh.S * e->D * sizeof(bf16);
    if (need > e->kvc_bytes) {
      if (e->kvc) cudaFree(e->kvc);
      e->kvc = nullptr; e->kvc_bytes = 0; e->cached_tag = 0;
      if (cudaMalloc(&e->kvc, need) != cudaSuccess) {
        set_error("context K/V cache allocation of %zu bytes failed", need);
        return LTX2_ERR_NOMEM;
      }
      e->kvc_bytes = need;
    }
    e->ctx_cache_hit = e->cached_tag == e->ctx_tag && e->cached_B == sh.B && e->cached_S == sh.S;
    e->cached_tag = 0;              // re-validated only when this forward has been enqueued completely
  }
  StreamBuf& vb = e->vb;
  StreamBuf& ab = e->ab;
  LTX2_PROPAGATE(upload_classes(vb, *video, cls_v, rc_v, st));
  if (has_audio) LTX2_PROPAGATE(upload_classes(ab, *audio, cls_a, rc_a, st));

  LTX2_PROPAGATE(prepare_stream(e, vb, e->vw, *video, false, e->table_arena_v, e->ptable_arena_v, st));
  if (has_audio) {
    LTX2_PROPAGATE(prepare_stream(e, ab, e->aw, *audio, true, e->table_arena_a, e->ptable_arena_a, st));
    // cross-modal RoPE: THIS modality's temporal axis at the audio width (model.py:320-343).  For the audio
    // stream that is exactly its self-attention table (same dim, max_pos and positions).
    const float mp[1] = {c.audio_max_pos};
    LTX2_PROPAGATE(rope_tables_dev(video->positions, sh.B, 3, 1, sh.N, e->Da, mp, e->fg_audio, e->nf_audio, vb.ccos,
                                   vb.csin, st));
    if (e->cp.world > 1) {
      // the v2a keys are ALL video tokens: gather every rank's slice of the table into every rank's copy
      CpState& cp = e->cp;
      cp.fwd_parity ^= 1;
      const size_t half = size_t(e->Da) / 2, slice = size_t(cp.n_local) * half * 4;
      for (int t = 0; t < 2; ++t)
        for (int b = 0; b < sh.B; ++b) {
          const size_t byte0 = cp.off_crope[cp.fwd_parity][t] + (size_t(b) * cp.n_total + size_t(cp.rank) * cp.n_local) * half * 4;
          const float* src = (t == 0 ? vb.ccos : vb.csin) + size_t(b) * cp.n_local * half;
          LTX2_CUDA_CHECK(cudaMemcpyAsync(cp.region + byte0, src, slice, cudaMemcpyDeviceToDevice, st));
          void* peers[kMaxCpRanks];
          int np = 0;
          for (int r = 0; r < cp.world; ++r)
            if (r != cp.rank) peers[np++] = cp.peer_base[r] + byte0;
          LTX2_PROPAGATE(peer_broadcast(cp.region + byte0, peers, np, int64_t(slice), st));
        }
    }
    // cross-attention timestep embeddings use the OTHER modality's sigma (model.py:394-399)
    LTX2_PROPAGATE(prepare_cross_mod(e, vb, e->av_v_ss, e->av_v_gate, e->catable_arena_v, *audio, st));
    LTX2_PROPAGATE(prepare_cross_mod(e, ab, e->av_a_ss, e->av_a_gate, e->catable_arena_a, *video, st));
  }

  const int D = e->D, Da = e->Da, B = sh.B, N = sh.N, Na = sh.Na;
  // fork / join of the audio side stream (LTX2_AUDIO_SIDE_STREAM=0: everything on the caller's stream)
  cudaStream_t sa = st;
  if (has_audio) {
    const char* env = getenv("LTX2_AUDIO_SIDE_STREAM");
    if (!(env && env[0] == '0')) {
      if (e->side == nullptr) {
        LTX2_CUDA_CHECK(cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking));
        LTX2_CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
        LTX2_CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
      }
      sa = e->side;
    }
  }
  auto fork_audio = [&]() -> int {          // the side stream may start once the main stream got here
    if (sa == st) return LTX2_OK;
    LTX2_CUDA_CHECK(cudaEventRecord(e->ev_fork, st));
    LTX2_CUDA_CHECK(cudaStreamWaitEvent(sa, e->ev_fork, 0));
    return LTX2_OK;
  };
  auto join_audio = [&]() -> int {          // the main stream continues once the side stream got here
    if (sa == st) return LTX2_OK;
    LTX2_CUDA_CHECK(cudaEventRecord(e->ev_join, sa));
    LTX2_CUDA_CHECK(cudaStreamWaitEvent(st, e->ev_join, 0));
    return LTX2_OK;
  };
  if (has_audio) LTX2_PROPAGATE(fork_audio());
  for (int l = 0; l < n_layers; ++l) {
    const BlockW& w = e->blocks[l];
    const uint64_t bit = uint64_t(1) << l;
    const float cas = isnan(e->cross_attn_scale[l]) ? 1.0f : e->cross_attn_scale[l];
    // split-K reductions are unordered fp32 atomics: only the sharded video stream uses them; the audio stream is
    // REPLICATED across ranks and must stay bit-identical on all of them
    // LTX2_SHARD_SPLIT_K (diagnostics): the split-K cap of a context-parallel rank on a single GPU, so that a one-GPU
    // run on a shard-sized token grid launches exactly a rank's GEMM kernels (tools/profile_step.py --grid 1 18 24)
    static const int shard_split = getenv("LTX2_SHARD_SPLIT_K") ? atoi(getenv("LTX2_SHARD_SPLIT_K")) : 1;
    const int split_v = e->cp.world > 1 ? e->cp.split_k : (shard_split > 1 ? shard_split : 1);
    g_split_k = split_v;
    LTX2_PROPAGATE(run_self_and_text(e, vb, w.v, l, (sk.video_self_attn & bit) != 0, cas, st));
    if (has_audio) {
      g_split_k = 1;
      LTX2_PROPAGATE(run_self_and_text(e, ab, w.a, l, (sk.audio_self_attn & bit) != 0, 1.0f, sa));
      LTX2_PROPAGATE(join_audio());            // the cross-modal attentions read and write both residual streams
      const bool do_a2v = (sk.a2v_cross_attn & bit) == 0, do_v2a = (sk.v2a_cross_attn & bit) == 0;
      const float* cmv = vb.ca_mod + size_t(l) * B * 5 * D;     // rows: scale_a2v, shift_a2v, scale_v2a, shift_v2a, gate
      const float* cma = ab.ca_mod + size_t(l) * B * 5 * Da;
      // all four modulated inputs are formed from the PRE-update residuals (transformer.py:561-562)
      bf16* vq = vb.xn; bf16* ak = ab.xn; bf16* aq = ab.hidden; bf16* vk = vb.hidden;
      if (do_a2v) {
        LTX2_PROPAGATE(rms_mod(e, vb, vq, cmv, int64_t(5) * D, 1, 0, vb.row_batch, st));
        LTX2_PROPAGATE(rms_mod(e, ab, ak, cma, int64_t(5) * Da, 1, 0, ab.row_batch, st));
      }
      if (do_v2a) {
        LTX2_PROPAGATE(rms_mod(e, ab, aq, cma, int64_t(5) * Da, 3, 2, ab.row_batch, st));
        LTX2_PROPAGATE(rms_mod(e, vb, vk, cmv, int64_t(5) * D, 3, 2, vb.row_batch, st));
      }
      if (do_a2v) {
        AttnCall a{&w.a2v, vq, D, B * N, N, ak, Da, Na, vb.ccos, vb.csin, ab.cos, ab.sin};
        g_split_k = split_v;
        LTX2_PROPAGATE(run_attention_core(e, vb, a, B, st));
        LTX2_PROPAGATE(linear_residual(vb.attn, w.a2v.inner, w.a2v.o, B * N, vb.x, D, cmv + 4 * D, int64_t(5) * D,
                                       vb.row_batch, 1.0f, st));
      }
      g_split_k = 1;
      if (do_v2a && e->cp.world > 1) {
        // audio queries (replicated) attend over ALL video tokens: every rank projects K|V of its token slice into
        // every rank's [B*Nt, 2*Da] buffer, then all ranks run the same (small) attention -- SURVEY.md 8(e) item 2
        CpState& cp = e->cp;
        const int inner = w.v2a.inner, Nt = cp.n_total;
        for (int b = 0; b < B; ++b) {
          const size_t byte0 = cp.off_akv + (size_t(b) * Nt + size_t(cp.rank) * N) * 2 * inner * sizeof(bf16);
          bf16* mine = reinterpret_cast<bf16*>(cp.region + byte0);
          LTX2_PROPAGATE(linear_bf16(vk + size_t(b) * N * D, D, w.v2a.kv, N, mine, 2 * inner, false, st));
          void* peers[kMaxCpRanks];
          int np = 0;
          for (int r = 0; r < cp.world; ++r)
            if (r != cp.rank) peers[np++] = cp.peer_base[r] + byte0;
          LTX2_PROPAGATE(peer_broadcast(mine, peers, np, int64_t(N) * 2 * inner * sizeof(bf16), st));
        }
        LTX2_PROPAGATE(cp_barrier(cp.peer_flags_dev, reinterpret_cast<uint32_t*>(cp.region + cp.off_flags), cp.rank,
                                  cp.world, ++cp.epoch, st));
        AttnCall a{&w.v2a, aq, Da, B * Na, Na, nullptr, D, Nt, ab.cos, ab.sin,
                   reinterpret_cast<const float*>(cp.region + cp.off_crope[cp.fwd_parity][0]),
                   reinterpret_cast<const float*>(cp.region + cp.off_crope[cp.fwd_parity][1])};
        a.kv_pre = reinterpret_cast<const bf16*>(cp.region + cp.off_akv);
        LTX2_PROPAGATE(run_attention_core(e, ab, a, B, st));
        LTX2_PROPAGATE(linear_residual(ab.attn, w.v2a.inner, w.v2a.o, B * Na, ab.x, Da, cma + 4 * Da, int64_t(5) * Da,
                                       ab.row_batch, 1.0f, st));
      } else if (do_v2a) {
        AttnCall a{&w.v2a, aq, Da, B * Na, Na, vk, D, N, ab.cos, ab.sin, vb.ccos, vb.csin};
        LTX2_PROPAGATE(run_attention_core(e, ab, a, B, st));
        LTX2_PROPAGATE(linear_residual(ab.attn, w.v2a.inner, w.v2a.o, B * Na, ab.x, Da, cma + 4 * Da, int64_t(5) * Da,
                                       ab.row_batch, 1.0f, st));
      }
    }
    if (has_audio) LTX2_PROPAGATE(fork_audio());
    g_split_k = split_v;
    LTX2_PROPAGATE(run_ffn(e, vb, w.v, l, st));
    g_split_k = 1;
    if (has_audio) LTX2_PROPAGATE(run_ffn(e, ab, w.a, l, sa));
  }
  LTX2_PROPAGATE(run_head(e, vb, e->vw, *video, x0, out_video, st));
  if (has_audio) {
    LTX2_PROPAGATE(run_head(e, ab, e->aw, *audio, x0, out_audio, sa));
    LTX2_PROPAGATE(join_audio());
  }
  if (e->ctx_cache_on) {
    e->cached_tag = e->ctx_tag;
    e->cached_B = sh.B;
    e->cached_S = sh.S;
  }
  return LTX2_OK;
}

extern "C" int ltx2_dit_set_context_tag(LtxDit* e, uint64_t tag) {
  LTX2_REQUIRE(e != nullptr, "dit_set_context_tag: null handle");
  e->ctx_tag = tag;
  return LTX2_OK;
}

extern "C" int ltx2_dit_set_layer_limit(LtxDit* e, int32_t n) {
  LTX2_REQUIRE(e != nullptr, "dit_set_layer_limit: null handle");
  e->layer_limit = (n > 0 && n < e->cfg.num_layers) ? n : 0;
  return LTX2_OK;
}

extern "C" int ltx2_dit_set_profile(LtxDit* e, int32_t on) {
  LTX2_REQUIRE(e != nullptr, "dit_set_profile: null handle");
  e->prof.on = on != 0;
  return LTX2_OK;
}

// After a profiled forward: synchronises the device and returns, per class (0 = GEMM, 1 = attention), the summed
// kernel time in ms, the summed algorithmic FLOPs and the launch count.
extern "C" int ltx2_dit_profile_read(LtxDit* e, double* ms_out, double* flops_out, int64_t* launches_out,
                                     int32_t n_classes) {
  LTX2_REQUIRE(e && ms_out && flops_out && launches_out && n_classes >= 2, "dit_profile_read: bad argument");
  LTX2_CUDA_CHECK(cudaDeviceSynchronize());
  for (int i = 0; i < n_classes; ++i) { ms_out[i] = 0; flops_out[i] = 0; launches_out[i] = 0; }
  for (auto& r : e->prof.recs) {
    float ms = 0.f;
    LTX2_CUDA_CHECK(cudaEventElapsedTime(&ms, e->prof.events[r.e0], e->prof.events[r.e0 + 1]));
    const int cat = r.cat < n_classes ? r.cat : 0;
    ms_out[cat] += ms;
    flops_out[cat] += r.work;
    launches_out[cat] += 1;
  }
  return LTX2_OK;
}

// Profiled launch i of the last forward (launch order): kernel time in ms, algorithmic FLOPs, class (0 = GEMM,
// 1 = attention, 2 = FP8 GEMM).  LTX2_ERR_INVALID when i is past the last record.
extern "C" int ltx2_dit_profile_launch(LtxDit* e, int32_t i, double* ms_out, double* flops_out, int32_t* class_out) {
  LTX2_REQUIRE(e && ms_out && flops_out && class_out, "dit_profile_launch: null argument");
  LTX2_REQUIRE(i >= 0 && i < static_cast<int32_t>(e->prof.recs.size()), "dit_profile_launch: launch %d of %d", i,
               static_cast<int>(e->prof.recs.size()));
  LTX2_CUDA_CHECK(cudaDeviceSynchronize());
  const auto& r = e->prof.recs[i];
  float ms = 0.f;
  LTX2_CUDA_CHECK(cudaEventElapsedTime(&ms, e->prof.events[r.e0], e->prof.events[r.e0 + 1]));
  *ms_out = ms;
  *flops_out = r.work;
  *class_out = r.cat;
  return LTX2_OK;
}

extern "C" int64_t ltx2_launch_count(void) { return ltx2::launch_count(); }

// =====================================================================================
// context parallel: exchange region + CUDA IPC plumbing
// =====================================================================================
extern "C" int ltx2_dit_cp_init(LtxDit* e, int32_t rank, int32_t world, int32_t batch, int32_t n_total,
                                int32_t context_tokens, char* handle_out /* 64 bytes */) {
  LTX2_REQUIRE(e && handle_out, "dit_cp_init: null argument");
  LTX2_REQUIRE(world >= 1 && world <= kMaxCpRanks && rank >= 0 && rank < world, "dit_cp_init: bad rank %d / world %d",
               rank, world);
  const LtxDitConfig& c = e->cfg;
  LTX2_REQUIRE(c.num_attention_heads % world == 0, "dit_cp_init: %d heads do not split over %d ranks",
               c.num_attention_heads, world);
  LTX2_REQUIRE(n_total % world == 0 && batch >= 1, "dit_cp_init: %d tokens do not split over %d ranks", n_total, world);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CpState& cp = e->cp;
  if (cp.region) {
    for (int r = 0; r < cp.world; ++r)
      if (cp.opened[r]) cudaIpcCloseMemHandle(cp.peer_base[r]);
    cudaFree(cp.region);
    if (cp.peer_flags_dev) cudaFree(cp.peer_flags_dev);
    cp = CpState();
  }
  cp.rank = rank; cp.world = world; cp.B = batch; cp.n_total = n_total; cp.n_local = n_total / world;
  cp.heads_local = c.num_attention_heads / world;
  if (const char* sk = getenv("LTX2_CP_SPLIT_K")) cp.split_k = std::min(16, std::max(1, atoi(sk)));
  const size_t Dh = c.attention_head_dim;
  const size_t qkv_bytes = (size_t(batch) * cp.heads_local * n_total * Dh * 2 + 255) & ~size_t(255);
  const size_t o_bytes = (size_t(batch) * cp.n_local * e->D * 2 + 255) & ~size_t(255);
  cp.off_q = 0; cp.off_k = qkv_bytes; cp.off_v = 2 * qkv_bytes; cp.off_o = 3 * qkv_bytes;
  size_t end = cp.off_o + o_bytes;
  if (context_tokens > 0 && context_tokens % world == 0 && (context_tokens / world) % 8 == 0) {
    cp.ctx_tokens = context_tokens;
    const size_t ckv_bytes = (size_t(batch) * context_tokens * 2 * e->D * 2 + 255) & ~size_t(255);
    cp.off_ckv[0] = end;
    cp.off_ckv[1] = end + ckv_bytes;
    end += 2 * ckv_bytes;
  }
  auto take = [&](size_t bytes) { const size_t o = end; end += (bytes + 255) & ~size_t(255); return o; };
  if (c.apply_gated_attention) cp.off_gate = take(size_t(batch) * n_total * cp.heads_local * 4);
  if (c.audio_enabled) {
    cp.off_akv = take(size_t(batch) * n_total * 2 * e->Da * 2);
    for (int p = 0; p < 2; ++p)
      for (int t = 0; t < 2; ++t) cp.off_crope[p][t] = take(size_t(batch) * n_total * (e->Da / 2) * 4);
  }
  cp.off_flags = end;
  cp.region_bytes = cp.off_flags + 256;
  LTX2_CUDA_CHECK(cudaMalloc(&cp.region, cp.region_bytes));
  LTX2_CUDA_CHECK(cudaMemset(cp.region, 0, cp.region_bytes));
  LTX2_CUDA_CHECK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  LTX2_CUDA_CHECK(cudaIpcGetMemHandle(&h, cp.region));
  memcpy(handle_out, &h, 64);
  return LTX2_OK;
}

// Orderly teardown of the exchange region.  An exported allocation must not be freed while a peer still maps it, so the
// host layer calls phase 0 on every rank (close the imported peer mappings), a host barrier, then phase 1 (free the own
// region; the engine is single-GPU again).
extern "C" int ltx2_dit_cp_shutdown(LtxDit* e, int32_t phase) {
  LTX2_REQUIRE(e != nullptr && (phase == 0 || phase == 1), "dit_cp_shutdown: bad argument");
  CpState& cp = e->cp;
  if (!cp.region) return LTX2_OK;
  LTX2_CUDA_CHECK(cudaDeviceSynchronize());
  if (phase == 0) {
    for (int r = 0; r < cp.world; ++r)
      if (cp.opened[r]) {
        cudaIpcCloseMemHandle(cp.peer_base[r]);
        cp.opened[r] = false;
        cp.peer_base[r] = nullptr;
      }
    cp.connected = false;
    return LTX2_OK;
  }
  cudaFree(cp.region);
  if (cp.peer_flags_dev) cudaFree(cp.peer_flags_dev);
  cp = CpState();
  return LTX2_OK;
}

// split-K cap of the residual GEMMs on sharded ranks: 1 = off (the sharded forward is then bit-identical to the
// single-GPU one), 0 = the default (8, or LTX2_CP_SPLIT_K)
extern "C" int ltx2_dit_cp_set_split_k(LtxDit* e, int32_t max_splits) {
  LTX2_REQUIRE(e != nullptr && max_splits >= 0 && max_splits <= 16, "dit_cp_set_split_k: bad argument");
  int k = max_splits;
  if (k == 0) {
    k = 8;
    if (const char* sk = getenv("LTX2_CP_SPLIT_K")) k = std::min(16, std::max(1, atoi(sk)));
  }
  e->cp.split_k = k;
  return LTX2_OK;
}

extern "C" int ltx2_dit_cp_connect(LtxDit* e, const char* handles /* world x 64 bytes, rank order */) {
  LTX2_REQUIRE(e && handles && e->cp.region, "dit_cp_connect: call ltx2_dit_cp_init first");
  CpState& cp = e->cp;
  std::vector<uint32_t*> flags(cp.world);
  for (int r = 0; r < cp.world; ++r) {
    if (r == cp.rank) {
      cp.peer_base[r] = cp.region;
    } else {
      cudaIpcMemHandle_t h;
      memcpy(&h, handles + size_t(r) * 64, 64);
      void* p = nullptr;
      LTX2_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      cp.peer_base[r] = reinterpret_cast<char*>(p);
      cp.opened[r] = true;
    }
    flags[r] = reinterpret_cast<uint32_t*>(cp.peer_base[r] + cp.off_flags);
  }
  LTX2_CUDA_CHECK(cudaMalloc(&cp.peer_flags_dev, sizeof(uint32_t*) * kMaxCpRanks));
  LTX2_CUDA_CHECK(cudaMemcpy(cp.peer_flags_dev, flags.data(), sizeof(uint32_t*) * cp.world, cudaMemcpyHostToDevice));
  cp.epoch = 0;
  cp.connected = true;
  return LTX2_OK;
}
