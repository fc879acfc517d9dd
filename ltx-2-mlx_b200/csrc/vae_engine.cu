// Video-VAE decoder engine: packed conv weights + activation workspace + the launch sequence of one
// SimpleVideoDecoder forward.
//
// Reference path replaced (LTX_2_MLX/model/video_vae/simple_decoder.py):
//   SimpleVideoDecoder.__init__/__call__ (:364-563), ResBlockGroup/ResBlock3d (:183-240, 316-336),
//   DepthToSpaceUpsample3d (:243-313), TimestepEmbedder (:42-59), load_vae_decoder_weights key names (:566-673).
//
// HBM layout: activations channels-last bf16 [B,T,H,W,C]; every conv input is materialised PADDED
// ([B,T+2,H+2,W+2,C]) by the kernel that also applies pixel-norm/scale-shift/SiLU, so the conv's TMA boxes need no
// boundary logic.  Conv weights are packed once at load time as bf16 [C_out, 27*C_in] (tap-major).
#include "common.cuh"
#include "kernels.h"
#include "../../include/ltx2_b200.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include <string>
#include <unordered_map>
#include <vector>

using namespace ltx2;
typedef __nv_bfloat16 bf16;

namespace {

enum SlotKind { SK_CONV_W, SK_CONV_B, SK_F32, SK_LIN_W, SK_SCALAR };

struct VSlot {
  int kind = SK_F32;
  void* dst = nullptr;
  int64_t n = 0;                 // element count expected (SK_F32 / SK_LIN_W / SK_CONV_B source length)
  int Cout = 0, Cout_pad = 0, Cin = 0, sp = 1;
  int64_t rows = 0, cols = 0;    // SK_LIN_W
  bool loaded = false;
};

struct ConvW {
  bf16* w = nullptr;
  float* b = nullptr;
  int Cin = 0, Cout = 0, Cout_pad = 0, sp = 1;
};

struct MlpW {                    // TimestepEmbedder: Linear(256,h) -> SiLU -> Linear(h,out)
  bf16 *w1 = nullptr, *w2 = nullptr;
  float *b1 = nullptr, *b2 = nullptr;
  int hidden = 0, out = 0;
  bool present = false;
};

struct StageW {
  int kind = 0;                  // 0 res group, 1 upsample
  int C = 0;                     // input channels
  int num_layers = 0;
  std::vector<ConvW> conv1, conv2;
  float* tables = nullptr;       // [num_layers, 4, C]
  MlpW temb;
  ConvW up;
  int ft = 1, fh = 1, fw = 1, multiplier = 1, residual = 0;
};

struct VArena {
  uintptr_t base = 0;
  size_t off = 0;
  void* take(size_t bytes) {
    size_t a = (off + 255) & ~size_t(255);
    off = a + bytes;
    return reinterpret_cast<void*>(base + a);
  }
};

}  // namespace

struct VaeProfiler {                 // CUDA events around every conv launch (bench.py roofline leg); off by default
  bool on = false;
  std::vector<cudaEvent_t> events;
  std::vector<double> flops;
  size_t used = 0;
};

constexpr int kMaxVaeRanks = 8;
// temporal shards: exchange region (two padded conv-input buffers + barrier flags) mapped into every rank by CUDA IPC
struct VaeCp {
  int rank = 0, world = 1;
  char* region = nullptr;
  size_t region_bytes = 0, pad_bytes = 0;
  size_t off_xp[2] = {0, 0}, off_flags = 0;
  size_t off_out = 0, out_bytes = 0;      // the clip being assembled: every rank stores its frames into the receivers' copy
  char* peer_base[kMaxVaeRanks] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool opened[kMaxVaeRanks] = {false, false, false, false, false, false, false, false};
  uint32_t** peer_flags_dev = nullptr;
  uint32_t epoch[4] = {0, 0, 0, 0};       // one barrier domain per group size: world, world/2, world/4, world/8
  bool connected = false;
};
constexpr int kVaeClipSlots = 4;           // clips in flight: two concurrent rank groups x two alternating rounds (a group
                                           // outside the receiver's group has no barrier with it inside a decode, so a
                                           // slot is only reused after one full-world collect in between)

struct LtxVae {
  VaeCp cp;
  VaeProfiler prof;
  LtxVaeConfig cfg;
  std::unordered_map<std::string, VSlot> slots;
  char* arena = nullptr;
  size_t arena_bytes = 0;
  float *mean = nullptr, *stdv = nullptr;
  ConvW conv_in, conv_out;
  std::vector<StageW> stages;
  float* last_table = nullptr;   // [2, Cf]
  MlpW last_temb;
  int Cf = 0;
  float ts_multiplier = 1000.0f;
  float* ts_mult_dev = nullptr;
  // workspace
  char* ws = nullptr;
  size_t ws_bytes = 0;
};

namespace {

int pad32(int c) { return (c + 31) / 32 * 32; }

struct VLayout {
  LtxVae* e;
  VArena* ar;
  bool dry;
  void put(const std::string& key, const VSlot& s) {
    if (!dry) e->slots[key] = s;
  }
  ConvW conv(const std::string& prefix, int Cout, int Cin, int sp = 1, int Cout_pad = -1) {
    ConvW c;
    c.Cin = Cin; c.Cout = Cout; c.sp = sp;
    c.Cout_pad = Cout_pad > 0 ? Cout_pad : pad32(Cout);
    c.w = reinterpret_cast<bf16*>(ar->take(sizeof(bf16) * size_t(c.Cout_pad) * 27 * Cin));
    c.b = reinterpret_cast<float*>(ar->take(sizeof(float) * c.Cout_pad));
    VSlot w;
    w.kind = SK_CONV_W; w.dst = c.w; w.n = int64_t(Cout) * Cin * 27; w.Cout = Cout; w.Cout_pad = c.Cout_pad;
    w.Cin = Cin; w.sp = sp;
    put(prefix + ".weight", w);
    VSlot b;
    b.kind = SK_CONV_B; b.dst = c.b; b.n = Cout; b.Cout = Cout; b.Cout_pad = c.Cout_pad; b.sp = sp;
    put(prefix + ".bias", b);
    return c;
  }
  float* f32(const std::string& key, int64_t n) {
    float* p = reinterpret_cast<float*>(ar->take(sizeof(float) * n));
    VSlot s;
    s.kind = SK_F32; s.dst = p; s.n = n;
    put(key, s);
    return p;
  }
  MlpW mlp(const std::string& prefix, int hidden, int out) {
    MlpW m;
    m.hidden = hidden; m.out = out; m.present = true;
    m.w1 = reinterpret_cast<bf16*>(ar->take(sizeof(bf16) * size_t(hidden) * 256));
    m.b1 = reinterpret_cast<float*>(ar->take(sizeof(float) * hidden));
    m.w2 = reinterpret_cast<bf16*>(ar->take(sizeof(bf16) * size_t(out) * hidden));
    m.b2 = reinterpret_cast<float*>(ar->take(sizeof(float) * out));
    VSlot s;
    s.kind = SK_LIN_W; s.dst = m.w1; s.rows = hidden; s.cols = 256; s.n = int64_t(hidden) * 256;
    put(prefix + ".linear_1.weight", s);
    s.dst = m.w2; s.rows = out; s.cols = hidden; s.n = int64_t(out) * hidden;
    put(prefix + ".linear_2.weight", s);
    VSlot b;
    b.kind = SK_F32; b.dst = m.b1; b.n = hidden;
    put(prefix + ".linear_1.bias", b);
    b.dst = m.b2; b.n = out;
    put(prefix + ".linear_2.bias", b);
    return m;
  }
};

int build_vae_layout(LtxVae* e, VArena* ar, bool dry) {
  VLayout L{e, ar, dry};
  const LtxVaeConfig& c = e->cfg;
  const int Lc = c.latent_channels;
  e->mean = L.f32("vae.per_channel_statistics.mean-of-means", Lc);
  e->stdv = L.f32("vae.per_channel_statistics.std-of-means", Lc);
  int C = c.base_channels * 8;
  e->conv_in = L.conv("vae.decoder.conv_in.conv", C, Lc);
  std::vector<StageW> stages;
  for (int i = 0; i < c.num_stages; ++i) {
    const LtxVaeStage& s = c.stages[i];
    const std::string U = "vae.decoder.up_blocks." + std::to_string(i);
    StageW st;
    st.kind = s.kind; st.C = C;
    if (s.kind == 0) {
      st.num_layers = s.num_layers;
      st.tables = reinterpret_cast<float*>(ar->take(sizeof(float) * size_t(s.num_layers) * 4 * C));
      for (int j = 0; j < s.num_layers; ++j) {
        const std::string R = U + ".res_blocks." + std::to_string(j);
        st.conv1.push_back(L.conv(R + ".conv1.conv", C, C));
        st.conv2.push_back(L.conv(R + ".conv2.conv", C, C));
        VSlot t;
        t.kind = SK_F32; t.n = int64_t(4) * C;
        t.dst = reinterpret_cast<void*>(reinterpret_cast<uintptr_t>(st.tables) + sizeof(float) * size_t(j) * 4 * C);
        L.put(R + ".scale_shift_table", t);
      }
      if (c.timestep_conditioning) st.temb = L.mlp(U + ".time_embedder.timestep_embedder", 4 * C, 4 * C);
    } else {
      st.ft = s.stride_t; st.fh = s.stride_h; st.fw = s.stride_w;
      st.multiplier = s.multiplier > 0 ? s.multiplier : 1;
      st.residual = s.residual;
      const int sp = st.ft * st.fh * st.fw;
      if ((sp * C) % st.multiplier != 0 || C % sp != 0) {
        set_error("vae: stage %d: channels %d incompatible with stride product %d / multiplier %d", i, C, sp,
                  st.multiplier);
        return LTX2_ERR_INVALID;
      }
      st.up = L.conv(U + ".conv.conv", sp * C / st.multiplier, C, sp);
      C = C / st.multiplier;
    }
    stages.push_back(st);
  }
  e->Cf = C;
  e->conv_out = L.conv("vae.decoder.conv_out.conv", 48, C, 1, 64);
  e->last_table = L.f32("vae.decoder.last_scale_shift_table", int64_t(2) * C);
  if (c.timestep_conditioning) {
    VSlot s;
    s.kind = SK_SCALAR; s.n = 1;
    e->ts_mult_dev = reinterpret_cast<float*>(ar->take(16));
    s.dst = e->ts_mult_dev;
    L.put("vae.decoder.timestep_scale_multiplier", s);
    // the reference hard-codes hidden 256 here (simple_decoder.py:663-665)
    e->last_temb = L.mlp("vae.decoder.last_time_embedder.timestep_embedder", 256, 2 * C);
  }
  if (!dry) e->stages = stages;
  return LTX2_OK;
}

struct Dims { int T, H, W, C; };

size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

}  // namespace

extern "C" {

int ltx2_vae_create(const LtxVaeConfig* cfg, LtxVae** out) {
  LTX2_REQUIRE(cfg && out, "vae_create: null argument");
  LTX2_REQUIRE(cfg->num_stages >= 1 && cfg->num_stages <= 16, "vae_create: 1..16 stages");
  LTX2_REQUIRE(cfg->base_channels % 64 == 0, "vae_create: base_channels must be a multiple of 64 (got %d)",
               cfg->base_channels);
  LTX2_REQUIRE(cfg->latent_channels % 64 == 0, "vae_create: latent_channels must be a multiple of 64");
  LtxVae* e = new LtxVae();
  e->cfg = *cfg;
  VArena dry;
  int s = build_vae_layout(e, &dry, true);
  if (s != LTX2_OK) { delete e; return s; }
  e->arena_bytes = dry.off + 1024;
  if (cudaMalloc(&e->arena, e->arena_bytes) != cudaSuccess) {
    set_error("vae weight arena allocation of %zu bytes failed", e->arena_bytes);
    delete e;
    return LTX2_ERR_NOMEM;
  }
  cudaMemset(e->arena, 0, e->arena_bytes);
  VArena real;
  real.base = reinterpret_cast<uintptr_t>(e->arena);
  build_vae_layout(e, &real, false);
  for (auto& st : e->stages) {
    const int sp = st.ft * st.fh * st.fw;
    if (st.kind == 1 && (st.up.Cout / sp) % 32 != 0) {
      set_error("vae: upsample output channels %d must be a multiple of 32", st.up.Cout / sp);
      ltx2_vae_destroy(e);
      return LTX2_ERR_INVALID;
    }
  }
  // tensors the reference treats as optional start as "loaded" with neutral values only where it does so:
  // none -- every slot must be set (the timestep MLPs are optional in the reference loader, see missing_weights).
  *out = e;
  return LTX2_OK;
}

void ltx2_vae_destroy(LtxVae* e) {
  if (!e) return;
  if (e->cp.region) {
    for (int r = 0; r < e->cp.world; ++r)
      if (e->cp.opened[r]) cudaIpcCloseMemHandle(e->cp.peer_base[r]);
    cudaFree(e->cp.region);
    if (e->cp.peer_flags_dev) cudaFree(e->cp.peer_flags_dev);
  }
  if (e->arena) cudaFree(e->arena);
  if (e->ws) cudaFree(e->ws);
  for (auto ev : e->prof.events) cudaEventDestroy(ev);
  delete e;
}

int ltx2_vae_set_weight(LtxVae* e, const char* key, const void* data, int32_t dtype, const int64_t* shape,
                        int32_t ndim, void* stream) {
  LTX2_REQUIRE(e && key && data, "vae_set_weight: null argument");
  auto it = e->slots.find(key);
  if (it == e->slots.end()) {
    set_error("vae_set_weight: unknown key '%s'", key);
    return LTX2_ERR_NOKEY;
  }
  VSlot& s = it->second;
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= shape[i];
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int r = LTX2_OK;
  switch (s.kind) {
    case SK_CONV_W:
      LTX2_REQUIRE(ndim == 5 && shape[0] == s.Cout && shape[1] == s.Cin && shape[2] == 3 && shape[3] == 3 && shape[4] == 3,
                   "vae_set_weight: '%s' expects [%d,%d,3,3,3]", key, s.Cout, s.Cin);
      r = pack_conv_weight(data, dtype, s.dst, s.Cout, s.Cout_pad, s.Cin, s.sp, st);
      break;
    case SK_CONV_B:
      LTX2_REQUIRE(n == s.Cout, "vae_set_weight: '%s' expects %d elements, got %lld", key, s.Cout, (long long)n);
      r = pack_conv_bias(data, dtype, reinterpret_cast<float*>(s.dst), s.Cout, s.Cout_pad, s.sp, st);
      break;
    case SK_LIN_W:
      LTX2_REQUIRE(ndim == 2 && shape[0] == s.rows && shape[1] == s.cols, "vae_set_weight: '%s' expects [%lld,%lld]",
                   key, (long long)s.rows, (long long)s.cols);
      r = cast_to_bf16(data, dtype, s.dst, n, st);
      break;
    case SK_SCALAR: {
      LTX2_REQUIRE(n == 1, "vae_set_weight: '%s' is a scalar", key);
      r = cast_to_f32(data, dtype, reinterpret_cast<float*>(s.dst), 1, st);
      if (r == LTX2_OK) {
        LTX2_CUDA_CHECK(cudaMemcpyAsync(&e->ts_multiplier, s.dst, 4, cudaMemcpyDeviceToHost, st));
        LTX2_CUDA_CHECK(cudaStreamSynchronize(st));
      }
      break;
    }
    default:
      LTX2_REQUIRE(n == s.n, "vae_set_weight: '%s' expects %lld elements, got %lld", key, (long long)s.n, (long long)n);
      r = cast_to_f32(data, dtype, reinterpret_cast<float*>(s.dst), n, st);
  }
  if (r == LTX2_OK) s.loaded = true;
  return r;
}

int ltx2_vae_missing_weights(LtxVae* e, char* names_out, int64_t names_cap) {
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

int ltx2_vae_output_shape(LtxVae* e, const int64_t in_shape[5], int64_t out_shape[5]) {
  LTX2_REQUIRE(e && in_shape && out_shape, "vae_output_shape: null argument");
  int64_t T = in_shape[2], H = in_shape[3], W = in_shape[4];
  for (auto& st : e->stages)
    if (st.kind == 1) {
      T = T * st.ft - (st.ft > 1 ? 1 : 0);
      H *= st.fh;
      W *= st.fw;
    }
  out_shape[0] = in_shape[0]; out_shape[1] = 3; out_shape[2] = T; out_shape[3] = H * 4; out_shape[4] = W * 4;
  return LTX2_OK;
}

}  // extern "C"

namespace {

// contiguous, as-even-as-possible split of T frames over the first min(world, T) ranks; other ranks get an empty range
void split_frames(int T, int world, int r, int* a, int* b) {
  const int active = std::min(world, T);
  if (r >= active) { *a = *b = T; return; }
  const int base = T / active, rem = T % active;
  *a = r * base + std::min(r, rem);
  *b = *a + base + (r < rem ? 1 : 0);
}

// One decoder forward.  sharded == false: the whole clip on this GPU (ltx2_vae_decode).  sharded == true: this rank
// computes the frames [a, b) of every activation (temporal shards, ltx2_vae_decode_sharded); `out` then receives only the
// rank's own output frames and *out_t0 / *out_tn their position in the clip.
int vae_decode_impl(LtxVae* e, const void* latent, int32_t dtype, const int64_t shape[5], float timestep,
                    float noise_scale, const float* noise, int32_t causal, float* out, void* stream, bool sharded,
                    int dst, int grp_first = 0, int grp_size = 0, int slot = 0) {
  LTX2_REQUIRE(e && latent && shape && (out || sharded), "vae_decode: null argument");
  const int B = (int)shape[0], Cl = (int)shape[1];
  LTX2_REQUIRE(Cl == e->cfg.latent_channels, "vae_decode: latent has %d channels, decoder expects %d", Cl,
               e->cfg.latent_channels);
  LTX2_REQUIRE(B >= 1 && shape[2] >= 1 && shape[3] >= 2 && shape[4] >= 2, "vae_decode: latent grid too small");
  const bool use_t = e->cfg.timestep_conditioning && timestep >= 0.f;
  for (auto& kv : e->slots)
    if (!kv.second.loaded) {
      // the reference loader treats the timestep embedders as optional (simple_decoder.py:620-636, 661-671)
      const bool optional = kv.first.find("time_embedder") != std::string::npos ||
                            kv.first.find("timestep_scale_multiplier") != std::string::npos;
      if (!optional || use_t) {
        set_error("vae_decode: weight '%s' has not been set", kv.first.c_str());
        return LTX2_ERR_STATE;
      }
    }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  VaeCp& cp = e->cp;
  int out_T_total = 0, out_T0 = 0;          // set for the sharded conv_out
  // the shard world is the rank GROUP [grp_first, grp_first + grp_size) of the connected ranks (the whole world by
  // default); `rank` is this rank's index inside it
  if (sharded && grp_size <= 0) { grp_first = 0; grp_size = cp.world; }
  const int world = sharded ? grp_size : 1, rank = sharded ? cp.rank - grp_first : 0;
  int dom = 0;                                       // barrier domain: log2(world / group size)
  if (sharded) {
    LTX2_REQUIRE(cp.connected && cp.world > 1, "vae_decode_sharded: ltx2_vae_cp_connect has not been called");
    LTX2_REQUIRE(!causal, "vae_decode_sharded: causal decoding is not sharded (its halo is two frames on one side)");
    LTX2_REQUIRE(grp_size >= 1 && cp.world % grp_size == 0 && grp_first % grp_size == 0 && rank >= 0 && rank < grp_size,
                 "vae_decode_sharded: rank %d is not in the group [%d, %d)", cp.rank, grp_first, grp_first + grp_size);
    for (int g = cp.world / grp_size; g > 1; g >>= 1) ++dom;
    LTX2_REQUIRE(dom < 4 && slot >= 0 && slot < kVaeClipSlots, "vae_decode_sharded: bad group size or clip slot");
  }
  // frame ranges [ra[r], rb[r]) of every rank at the current stage; Tt = frames of the whole clip at that stage
  int ra[kMaxVaeRanks], rb[kMaxVaeRanks];
  int Tt = (int)shape[2];
  for (int r = 0; r < world; ++r) split_frames(Tt, world, r, &ra[r], &rb[r]);
  auto n_of = [&](int r) { return (r >= 0 && r < world) ? rb[r] - ra[r] : 0; };

  // ---- workspace sizing: walk the stages (local frame counts when sharded) ----
  Dims d{n_of(rank), (int)shape[3], (int)shape[4], e->conv_in.Cout};
  size_t max_act = size_t(B) * std::max(d.T, 1) * d.H * d.W * d.C;
  size_t max_pad = size_t(B) * (d.T + 2) * (d.H + 2) * (d.W + 2) * std::max(d.C, Cl);
  size_t max_mod = 0;
  {
    int a = ra[rank], b = rb[rank];
    for (auto& s : e->stages) {
      if (s.kind == 0) {
        max_mod = std::max(max_mod, size_t(s.num_layers) * B * 4 * s.C);
      } else {
        if (s.ft > 1 && b > a) { a = std::max(0, 2 * a - 1); b = 2 * b - 1; }
        d.T = b - a; d.H *= s.fh; d.W *= s.fw; d.C = s.C / s.multiplier;
        max_act = std::max(max_act, size_t(B) * std::max(d.T, 1) * d.H * d.W * d.C);
        max_pad = std::max(max_pad, size_t(B) * (d.T + 2) * (d.H + 2) * (d.W + 2) * d.C);
      }
    }
  }
  max_mod = std::max(max_mod, size_t(B) * 2 * e->Cf);
  const int maxC = e->conv_in.Cout;
  const size_t need = align256(max_act * 2) * 2 + align256(max_pad * 2) * 2 + align256(max_mod * 4) +
                      align256(size_t(B) * 2 * e->Cf * 4) + align256(size_t(B) * (256 + 4 * maxC * 2 + 8) * 4) + 4096;
  if (need > e->ws_bytes) {
    if (e->ws) cudaFree(e->ws);
    e->ws = nullptr; e->ws_bytes = 0;
    if (cudaMalloc(&e->ws, need) != cudaSuccess) {
      set_error("vae workspace allocation of %zu bytes failed", need);
      return LTX2_ERR_NOMEM;
    }
    e->ws_bytes = need;
  }
  VArena ws;
  ws.base = reinterpret_cast<uintptr_t>(e->ws);
  bf16* cur = reinterpret_cast<bf16*>(ws.take(max_act * 2));
  bf16* other = reinterpret_cast<bf16*>(ws.take(max_act * 2));
  bf16* xps[2];
  xps[0] = reinterpret_cast<bf16*>(ws.take(max_pad * 2));
  xps[1] = reinterpret_cast<bf16*>(ws.take(max_pad * 2));
  if (sharded) {
    // the padded conv inputs live in the exchange region: neighbours store their boundary frames into them
    LTX2_REQUIRE(max_pad * 2 <= cp.pad_bytes, "vae_decode_sharded: padded input of %zu bytes exceeds the exchange buffers "
                 "(%zu): call ltx2_vae_cp_init with the largest latent", max_pad * 2, cp.pad_bytes);
    xps[0] = reinterpret_cast<bf16*>(cp.region + cp.off_xp[0]);
    xps[1] = reinterpret_cast<bf16*>(cp.region + cp.off_xp[1]);
  }
  float* mod = reinterpret_cast<float*>(ws.take(max_mod * 4));
  float* mod_final = reinterpret_cast<float*>(ws.take(size_t(B) * 2 * e->Cf * 4));
  float* tdev = reinterpret_cast<float*>(ws.take(size_t(B) * 4 + 16));
  float* sinus = reinterpret_cast<float*>(ws.take(size_t(B) * 256 * 4));
  float* hid = reinterpret_cast<float*>(ws.take(size_t(B) * 4 * maxC * 4));
  float* temb = reinterpret_cast<float*>(ws.take(size_t(B) * 4 * maxC * 4));
  LTX2_REQUIRE(B <= 8, "vae_decode: batch %d > 8 unsupported", B);

  // ---- temporal shards: halo exchange in front of every conv except conv_in ----
  // A conv reads buffer xps[i]; the producer of the NEXT conv's input writes xps[1 - i], so a neighbour that is one conv
  // ahead never stores into a buffer this rank is still reading.
  int last_read = 1;
  auto wbuf = [&]() { return 1 - last_read; };
  auto has_prev = [&]() { return sharded && n_of(rank) > 0 && ra[rank] > 0; };
  auto has_next = [&]() { return sharded && n_of(rank) > 0 && rb[rank] < Tt; };
  auto sync_halo = [&](int buf, const Dims& dd) -> int {
    if (!sharded) return LTX2_OK;
    if (n_of(rank) > 0) {
      const int64_t frame_bytes = int64_t(dd.H + 2) * (dd.W + 2) * dd.C * 2;
      void* prev = has_prev() ? cp.peer_base[cp.rank - 1] + cp.off_xp[buf] : nullptr;
      void* next = has_next() ? cp.peer_base[cp.rank + 1] + cp.off_xp[buf] : nullptr;
      LTX2_PROPAGATE(halo_push(xps[buf], prev, next, B, n_of(rank), n_of(rank - 1), n_of(rank + 1), frame_bytes, st));
    }
    if (grp_size == 1) return LTX2_OK;               // a group of one rank: nothing to wait for
    return cp_barrier_group(cp.peer_flags_dev, reinterpret_cast<uint32_t*>(cp.region + cp.off_flags), cp.rank, grp_first,
                            grp_size, dom * 8, ++cp.epoch[dom], st);
  };
  const bool idle = sharded && n_of(rank) == 0;      // more ranks than latent frames: only keeps the barriers in step

  if (use_t && !idle) {
    std::vector<float> tv(B, timestep);
    LTX2_CUDA_CHECK(cudaMemcpyAsync(tdev, tv.data(), size_t(B) * 4, cudaMemcpyHostToDevice, st));
    LTX2_CUDA_CHECK(cudaStreamSynchronize(st));     // tv is a stack-owned staging buffer
    LTX2_PROPAGATE(timestep_sinusoid(tdev, B, e->ts_multiplier, sinus, st));
  }
  auto run_mlp = [&](const MlpW& m) -> int {
    LTX2_PROPAGATE(small_linear(sinus, B, 256, m.w1, m.b1, hid, m.hidden, 0, st));
    return small_linear(hid, B, m.hidden, m.w2, m.b2, temb, m.out, 1, st);
  };
  struct PadOut {                 // fused producer of the next conv's padded input (conv3d_sm100.cu)
    bf16* dst = nullptr;
    int act = 0;
    const float* mod = nullptr;
    int64_t stride = 0, shift_off = 0, scale_off = 0;
  };
  // conv that reads xps[buf]; in sharded mode it is preceded by the halo exchange of that buffer (except conv_in)
  auto conv = [&](const ConvW& w, int buf, bool halo, const Dims& dd, int mode, bf16* o, const bf16* residual, float* o32,
                  const StageW* up, const PadOut* po = nullptr) -> int {
    if (halo) LTX2_PROPAGATE(sync_halo(buf, Dims{dd.T, dd.H, dd.W, w.Cin}));
    last_read = buf;
    if (idle) return LTX2_OK;
    ConvParams p;
    if (po != nullptr && po->dst != nullptr) {
      p.pad_out = po->dst; p.pad_act = po->act; p.pad_mod = po->mod; p.pad_mod_stride = po->stride;
      p.pad_shift_off = po->shift_off; p.pad_scale_off = po->scale_off; p.pad_eps = 1e-6f; p.pad_causal = causal;
      p.pad_skip_front = has_prev(); p.pad_skip_back = has_next();
    }
    p.B = B; p.T = dd.T; p.H = dd.H; p.W = dd.W;
    p.Cin = w.Cin; p.Cout = w.Cout; p.Cout_pad = w.Cout_pad;
    p.mode = mode; p.bias = w.b; p.out = o; p.out_f32 = o32; p.residual = residual;
    if (up) {
      p.ft = up->ft; p.fh = up->fh; p.fw = up->fw;
      p.c_d2s = up->residual ? w.Cin / (up->ft * up->fh * up->fw) : 0;
      p.d2s_keep_first = has_prev();
    }
    p.out_t_total = out_T_total; p.out_t0 = out_T0;
    VaeProfiler& pf = e->prof;
    if (pf.on) {
      if (pf.used + 2 > pf.events.size()) {
        const size_t old = pf.events.size();
        pf.events.resize(old + 256);
        for (size_t i = old; i < pf.events.size(); ++i) cudaEventCreate(&pf.events[i]);
      }
      pf.flops.push_back(2.0 * w.Cin * double(w.Cout) * 27.0 * B * dd.T * double(dd.H) * dd.W);
      cudaEventRecord(pf.events[pf.used], st);
    }
    const int r = conv3d_bf16(xps[buf], w.w, p, st);
    if (pf.on) {
      cudaEventRecord(pf.events[pf.used + 1], st);
      pf.used += 2;
    }
    return r;
  };
  // separate normalise / activate / pad pass into the write buffer
  auto pad_pass = [&](const bf16* src, const Dims& dd, int C, int act, const float* m, int64_t stride, int64_t sh,
                      int64_t sc) -> int {
    if (idle) return LTX2_OK;
    return norm_act_pad(src, xps[wbuf()], B, dd.T, dd.H, dd.W, C, act, m, stride, sh, sc, 1e-6f, causal, st, has_prev(),
                        has_next());
  };
  e->prof.used = 0;
  e->prof.flops.clear();

  d = Dims{n_of(rank), (int)shape[3], (int)shape[4], Cl};
  const bool inject = use_t && noise != nullptr && noise_scale != 0.f;
  if (!idle)
    LTX2_PROPAGATE(latent_to_padded(latent, dtype, e->stdv, e->mean, inject ? noise : nullptr, noise_scale, xps[0], B, Cl,
                                    Tt, d.H, d.W, causal, st, ra[rank], d.T));
  LTX2_PROPAGATE(conv(e->conv_in, 0, false, d, CONV_EPI_PLAIN, cur, nullptr, nullptr, nullptr));
  d.C = e->conv_in.Cout;

  // Final norm / scale-shift rows (simple_decoder.py:528-542), computed up front: the last conv of the last group
  // produces the activated, padded input of conv_out in its epilogue.
  const int Cf = e->Cf;
  if (!idle) {
    if (use_t && e->last_temb.present) {
      LTX2_PROPAGATE(run_mlp(e->last_temb));
    } else {
      LTX2_CUDA_CHECK(cudaMemsetAsync(temb, 0, size_t(B) * 2 * Cf * 4, st));
    }
    LTX2_PROPAGATE(build_modulation_ex(e->last_table, 0, temb, int64_t(2) * Cf, Cf, mod_final, 0, int64_t(2) * Cf, 1, B,
                                       2, Cf, st));
  }
  // Stages with 128 or 256 channels (83 % of the conv FLOPs, all of the large activations) run FUSED: every conv's
  // epilogue writes the next conv's input already pixel-normalised, modulated, SiLU-activated and padded, so the
  // separate norm_act_pad pass (one read + one write of the activation per conv) only runs once per group.
  // LTX2_VAE_FUSE=0 restores the unfused sequence (A/B and debugging).
  const char* fuse_env = getenv("LTX2_VAE_FUSE");
  const bool fuse_on = !(fuse_env && fuse_env[0] == '0');
  bool xp_ready = false;           // xps[wbuf()] already holds the padded input of the next consumer
  const size_t n_stages = e->stages.size();
  for (size_t si = 0; si < n_stages; ++si) {
    StageW& s = e->stages[si];
    if (s.kind == 0) {
      const int C = s.C;
      if (!idle) {
        if (use_t && s.temb.present) {
          LTX2_PROPAGATE(run_mlp(s.temb));
        } else {
          LTX2_CUDA_CHECK(cudaMemsetAsync(temb, 0, size_t(B) * 4 * C * 4, st));
        }
        LTX2_PROPAGATE(build_modulation_ex(s.tables, int64_t(4) * C, temb, int64_t(4) * C, C, mod, int64_t(B) * 4 * C,
                                           int64_t(4) * C, s.num_layers, B, 4, C, st));
      }
      const bool fuse = fuse_on && (C == 128 || C == 256);
      if (!fuse) {
        for (int j = 0; j < s.num_layers; ++j) {
          const float* mj = mod + size_t(j) * B * 4 * C;
          LTX2_PROPAGATE(pad_pass(cur, d, C, 1, mj, int64_t(4) * C, 0, C));
          LTX2_PROPAGATE(conv(s.conv1[j], wbuf(), true, d, CONV_EPI_PLAIN, other, nullptr, nullptr, nullptr));
          LTX2_PROPAGATE(pad_pass(other, d, C, 1, mj, int64_t(4) * C, int64_t(2) * C, int64_t(3) * C));
          LTX2_PROPAGATE(conv(s.conv2[j], wbuf(), true, d, CONV_EPI_RESIDUAL, cur, cur, nullptr, nullptr));
        }
        xp_ready = false;
        continue;
      }
      LTX2_PROPAGATE(pad_pass(cur, d, C, 1, mod, int64_t(4) * C, 0, C));
      for (int j = 0; j < s.num_layers; ++j) {
        const float* mj = mod + size_t(j) * B * 4 * C;
        // conv1: reads xps[i] -> xps[1 - i] = act(norm2_j(.)); its raw output has no other reader
        int rb_ = wbuf();
        PadOut p1;
        p1.dst = xps[1 - rb_]; p1.act = 1; p1.mod = mj; p1.stride = int64_t(4) * C; p1.shift_off = int64_t(2) * C;
        p1.scale_off = int64_t(3) * C;
        LTX2_PROPAGATE(conv(s.conv1[j], rb_, true, d, CONV_EPI_PLAIN, nullptr, nullptr, nullptr, nullptr, &p1));
        // conv2: -> cur = residual + conv (raw, the next residual) and the padded input of the next consumer
        rb_ = wbuf();
        PadOut p2;
        p2.dst = xps[1 - rb_];
        if (j + 1 < s.num_layers) {
          p2.act = 1; p2.mod = mod + size_t(j + 1) * B * 4 * C; p2.stride = int64_t(4) * C; p2.shift_off = 0;
          p2.scale_off = C;
          xp_ready = true;
        } else if (si + 1 == n_stages) {
          p2.act = 1; p2.mod = mod_final; p2.stride = int64_t(2) * Cf; p2.shift_off = 0; p2.scale_off = Cf;
          xp_ready = true;
        } else if (e->stages[si + 1].kind == 1) {
          p2.act = 0;                       // the depth-to-space conv reads the raw activation, padded
          xp_ready = true;
        } else {
          p2.dst = nullptr;                 // another res group follows: its rows are not built yet
          xp_ready = false;
        }
        LTX2_PROPAGATE(conv(s.conv2[j], rb_, true, d, CONV_EPI_RESIDUAL, cur, cur, nullptr, nullptr, &p2));
      }
    } else {
      if (!xp_ready) LTX2_PROPAGATE(pad_pass(cur, d, s.C, 0, nullptr, 0, 0, 0));
      xp_ready = false;
      LTX2_PROPAGATE(conv(s.up, wbuf(), true, d, CONV_EPI_D2S, other, cur, nullptr, &s));
      std::swap(cur, other);
      if (s.ft > 1) {
        for (int r = 0; r < world; ++r)
          if (rb[r] > ra[r]) { ra[r] = std::max(0, 2 * ra[r] - 1); rb[r] = 2 * rb[r] - 1; }
        Tt = 2 * Tt - 1;
        for (int r = 0; r < world; ++r)
          if (rb[r] <= ra[r]) ra[r] = rb[r] = Tt;
      }
      d.T = n_of(rank); d.H *= s.fh; d.W *= s.fw; d.C = s.C / s.multiplier;
    }
  }
  // final norm + scale/shift + SiLU (unless the last conv already produced it), conv_out, unpatchify (:528-552)
  if (!xp_ready) LTX2_PROPAGATE(pad_pass(cur, d, Cf, 1, mod_final, int64_t(2) * Cf, 0, Cf));
  if (!sharded) return conv(e->conv_out, wbuf(), true, d, CONV_EPI_UNPATCHIFY, nullptr, nullptr, out, nullptr);
  // Temporal shards: conv_out writes this rank's frames at their place in ITS copy of the clip (exchange region, clip
  // slot `slot`), and the spans are then stored into the receiving ranks' copies over NVLink.  ltx2_vae_cp_collect (a
  // barrier over ALL ranks, then a copy out of the region) completes the clip -- no collective library call anywhere.
  const int64_t HW = int64_t(d.H) * 4 * d.W * 4;
  const size_t clip_bytes = size_t(B) * 3 * Tt * HW * 4;
  LTX2_REQUIRE(clip_bytes <= cp.out_bytes, "vae_decode_sharded: clip of %zu bytes exceeds the exchange buffer (%zu)",
               clip_bytes, cp.out_bytes);
  const size_t clip_off = cp.off_out + size_t(slot) * cp.out_bytes;
  float* my_clip = reinterpret_cast<float*>(cp.region + clip_off);
  out_T_total = Tt;
  out_T0 = idle ? 0 : ra[rank];
  LTX2_PROPAGATE(conv(e->conv_out, wbuf(), true, d, CONV_EPI_UNPATCHIFY, nullptr, nullptr, my_clip, nullptr));
  if (!idle) {
    void* peers[kMaxVaeRanks];
    int np = 0;
    for (int r = 0; r < cp.world; ++r)
      if (r != cp.rank && (dst < 0 || r == dst)) peers[np++] = cp.peer_base[r] + clip_off;
    if (np > 0)
      for (int bc = 0; bc < B * 3; ++bc) {
        const size_t off = (size_t(bc) * Tt + ra[rank]) * HW * 4;
        void* pp[kMaxVaeRanks];
        for (int i = 0; i < np; ++i) pp[i] = static_cast<char*>(peers[i]) + off;
        LTX2_PROPAGATE(peer_broadcast(reinterpret_cast<char*>(my_clip) + off, pp, np, int64_t(n_of(rank)) * HW * 4, st));
      }
  }
  (void)out;
  return LTX2_OK;
}

}  // namespace

extern "C" {

int ltx2_vae_decode(LtxVae* e, const void* latent, int32_t dtype, const int64_t shape[5], float timestep,
                    float noise_scale, const float* noise, int32_t causal, float* out, void* stream) {
  LTX2_REQUIRE(out != nullptr, "vae_decode: null output");
  return vae_decode_impl(e, latent, dtype, shape, timestep, noise_scale, noise, causal, out, stream, false, 0);
}

// ---- temporal shards over the GPUs of one NVLink box ---------------------------------------------
// Every rank holds the whole latent and the decoder weights and computes a contiguous range of frames of every activation.
// A 3x3x3 conv needs one frame from each neighbour: the padded conv inputs live in a CUDA-IPC exchange region, and after a
// rank has produced its padded input it stores its first / last frame into the neighbours' pad slots (halo_push) and all
// ranks pass a flag barrier -- one small kernel pair per conv, no collective library on the data path.  Frame t of a
// stage becomes frames 2t-1, 2t of the next (frame 0 -> 0), so ownership follows the depth-to-space without any
// re-distribution, and every rank ends with a contiguous range of output frames.  Bit-identical to the single-GPU decode.
int ltx2_vae_shard_frames(LtxVae* e, int64_t latent_frames, int32_t rank, int32_t world, int64_t* t0, int64_t* tn) {
  LTX2_REQUIRE(e && t0 && tn && world >= 1 && world <= kMaxVaeRanks && rank >= 0 && rank < world && latent_frames >= 1,
               "vae_shard_frames: bad argument");
  int a, b;
  split_frames((int)latent_frames, world, rank, &a, &b);
  for (auto& s : e->stages)
    if (s.kind == 1 && s.ft > 1 && b > a) { a = std::max(0, 2 * a - 1); b = 2 * b - 1; }
  *t0 = b > a ? a : 0;
  *tn = b - a;
  return LTX2_OK;
}

int ltx2_vae_cp_init(LtxVae* e, int32_t rank, int32_t world, const int64_t max_latent_shape[5], char* handle_out) {
  LTX2_REQUIRE(e && handle_out && max_latent_shape, "vae_cp_init: null argument");
  LTX2_REQUIRE(world >= 2 && world <= kMaxVaeRanks && rank >= 0 && rank < world, "vae_cp_init: bad rank %d / world %d", rank,
               world);
  VaeCp& cp = e->cp;
  LTX2_REQUIRE(cp.region == nullptr, "vae_cp_init: already initialised (call ltx2_vae_cp_shutdown first)");
  // largest padded conv input of ANY rank for this latent shape (the region layout must be the same on every rank)
  const int B = (int)max_latent_shape[0], T = (int)max_latent_shape[2];
  size_t max_pad = 0;
  // ... for every group size a decode may run with (the world, its halves, ..., one rank)
  for (int gs = world; gs >= 1; gs >>= 1)
  for (int r = 0; r < gs; ++r) {
    int a, b;
    split_frames(T, gs, r, &a, &b);
    int H = (int)max_latent_shape[3], W = (int)max_latent_shape[4], C = std::max(e->conv_in.Cout, e->cfg.latent_channels);
    max_pad = std::max(max_pad, size_t(B) * (b - a + 2) * (H + 2) * (W + 2) * C);
    for (auto& s : e->stages)
      if (s.kind == 1) {
        if (s.ft > 1 && b > a) { a = std::max(0, 2 * a - 1); b = 2 * b - 1; }
        H *= s.fh; W *= s.fw; C = s.C / s.multiplier;
        max_pad = std::max(max_pad, size_t(B) * (b - a + 2) * (H + 2) * (W + 2) * C);
      }
  }
  cp.rank = rank; cp.world = world;
  cp.pad_bytes = align256(max_pad * 2);
  {
    int64_t os[5];
    ltx2_vae_output_shape(e, max_latent_shape, os);
    cp.out_bytes = align256(size_t(os[0]) * os[1] * os[2] * os[3] * os[4] * 4);
  }
  cp.off_xp[0] = 0; cp.off_xp[1] = cp.pad_bytes; cp.off_out = 2 * cp.pad_bytes;
  cp.off_flags = cp.off_out + kVaeClipSlots * cp.out_bytes;
  cp.region_bytes = cp.off_flags + 256;
  LTX2_CUDA_CHECK(cudaMalloc(&cp.region, cp.region_bytes));
  LTX2_CUDA_CHECK(cudaMemset(cp.region, 0, cp.region_bytes));
  LTX2_CUDA_CHECK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  LTX2_CUDA_CHECK(cudaIpcGetMemHandle(&h, cp.region));
  memcpy(handle_out, &h, 64);
  return LTX2_OK;
}

int ltx2_vae_cp_connect(LtxVae* e, const char* handles) {
  LTX2_REQUIRE(e && handles && e->cp.region, "vae_cp_connect: call ltx2_vae_cp_init first");
  VaeCp& cp = e->cp;
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
  LTX2_CUDA_CHECK(cudaMalloc(&cp.peer_flags_dev, sizeof(uint32_t*) * kMaxVaeRanks));
  LTX2_CUDA_CHECK(cudaMemcpy(cp.peer_flags_dev, flags.data(), sizeof(uint32_t*) * cp.world, cudaMemcpyHostToDevice));
  for (int i = 0; i < 4; ++i) cp.epoch[i] = 0;
  cp.connected = true;
  return LTX2_OK;
}

// phase 0 on every rank (close the imported mappings), host barrier, phase 1 (free the own region)
int ltx2_vae_cp_shutdown(LtxVae* e, int32_t phase) {
  LTX2_REQUIRE(e != nullptr && (phase == 0 || phase == 1), "vae_cp_shutdown: bad argument");
  VaeCp& cp = e->cp;
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
  cp = VaeCp();
  return LTX2_OK;
}

int ltx2_vae_decode_sharded(LtxVae* e, const void* latent, int32_t dtype, const int64_t shape[5], float timestep,
                            float noise_scale, const float* noise, int32_t group_first, int32_t group_size, int32_t slot,
                            int32_t dst, void* stream) {
  LTX2_REQUIRE(e && dst >= -1 && dst < e->cp.world, "vae_decode_sharded: bad destination rank %d", dst);
  return vae_decode_impl(e, latent, dtype, shape, timestep, noise_scale, noise, 0, nullptr, stream, true, dst, group_first,
                         group_size, slot);
}

// Completes clip slot `slot`: a barrier over ALL connected ranks (every group that stored into the slot has finished),
// then the receiving ranks (dst, or all with dst = -1) copy the clip [B,3,T',32H,32W] fp32 out of the exchange region.
int ltx2_vae_cp_collect(LtxVae* e, int32_t slot, const int64_t clip_shape[5], int32_t dst, float* out, void* stream) {
  LTX2_REQUIRE(e && clip_shape && e->cp.connected && slot >= 0 && slot < kVaeClipSlots && dst >= -1 && dst < e->cp.world,
               "vae_cp_collect: bad argument");
  VaeCp& cp = e->cp;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LTX2_PROPAGATE(cp_barrier_group(cp.peer_flags_dev, reinterpret_cast<uint32_t*>(cp.region + cp.off_flags), cp.rank, 0,
                                  cp.world, 0, ++cp.epoch[0], st));
  if (dst < 0 || dst == cp.rank) {
    LTX2_REQUIRE(out != nullptr, "vae_cp_collect: the receiving rank needs an output buffer");
    const size_t bytes = size_t(clip_shape[0]) * clip_shape[1] * clip_shape[2] * clip_shape[3] * clip_shape[4] * 4;
    LTX2_REQUIRE(bytes <= cp.out_bytes, "vae_cp_collect: clip larger than the exchange buffer");
    LTX2_CUDA_CHECK(cudaMemcpyAsync(out, cp.region + cp.off_out + size_t(slot) * cp.out_bytes, bytes,
                                    cudaMemcpyDeviceToDevice, st));
  }
  return LTX2_OK;
}

// Conv3dSimple.__call__ (simple_decoder.py:90-180) as ONE op, for unit parity of the conv kernel at production shapes:
// pad (reflect H/W, replicate T) -> implicit-GEMM conv -> bias.  The workspace holds the padded input, the packed
// weight and the packed bias.
int64_t ltx2_conv3d_workspace_bytes(int32_t B, int32_t T, int32_t H, int32_t W, int32_t Cin, int32_t Cout) {
  const size_t pad = align256(size_t(B) * (T + 2) * (H + 2) * (W + 2) * Cin * 2);
  const size_t wt = align256(size_t(pad32(Cout)) * 27 * Cin * 2);
  return static_cast<int64_t>(pad + wt + align256(size_t(pad32(Cout)) * 4) + 1024);
}

int ltx2_conv3d(const void* x, const void* weight, int32_t w_dtype, const void* bias, int32_t b_dtype, void* out,
                int32_t B, int32_t T, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t causal, void* workspace,
                void* stream) {
  LTX2_REQUIRE(x && weight && bias && out && workspace, "conv3d: null argument");
  LTX2_REQUIRE(B >= 1 && T >= 1 && H >= 2 && W >= 2, "conv3d: grid too small (reflect padding needs H, W >= 2)");
  LTX2_REQUIRE(Cin % 64 == 0 && Cin <= 1024 && Cout % 8 == 0, "conv3d: C_in %% 64 == 0 (<= 1024), C_out %% 8 == 0");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  VArena ws;
  ws.base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
  const int Cp = pad32(Cout);
  bf16* xp = reinterpret_cast<bf16*>(ws.take(size_t(B) * (T + 2) * (H + 2) * (W + 2) * Cin * 2));
  bf16* wp = reinterpret_cast<bf16*>(ws.take(size_t(Cp) * 27 * Cin * 2));
  float* bp = reinterpret_cast<float*>(ws.take(size_t(Cp) * 4));
  LTX2_PROPAGATE(pack_conv_weight(weight, w_dtype, wp, Cout, Cp, Cin, 1, st));
  LTX2_PROPAGATE(pack_conv_bias(bias, b_dtype, bp, Cout, Cp, 1, st));
  LTX2_PROPAGATE(norm_act_pad(x, xp, B, T, H, W, Cin, 0, nullptr, 0, 0, 0, 1e-6f, causal, st));
  ConvParams p;
  p.B = B; p.T = T; p.H = H; p.W = W;
  p.Cin = Cin; p.Cout = Cout; p.Cout_pad = Cp;
  p.mode = CONV_EPI_PLAIN; p.bias = bp; p.out = reinterpret_cast<bf16*>(out);
  return conv3d_bf16(xp, wp, p, st);
}

int ltx2_vae_set_profile(LtxVae* e, int32_t on) {
  LTX2_REQUIRE(e != nullptr, "vae_set_profile: null handle");
  e->prof.on = on != 0;
  return LTX2_OK;
}

// after a profiled decode: device sync, then summed conv kernel time (ms), algorithmic conv FLOPs and launch count
int ltx2_vae_profile_read(LtxVae* e, double* ms_out, double* flops_out, int64_t* launches_out) {
  LTX2_REQUIRE(e && ms_out && flops_out && launches_out, "vae_profile_read: null argument");
  LTX2_CUDA_CHECK(cudaDeviceSynchronize());
  *ms_out = 0; *flops_out = 0; *launches_out = 0;
  for (size_t i = 0; i * 2 < e->prof.used; ++i) {
    float ms = 0.f;
    LTX2_CUDA_CHECK(cudaEventElapsedTime(&ms, e->prof.events[2 * i], e->prof.events[2 * i + 1]));
    *ms_out += ms;
    *flops_out += e->prof.flops[i];
    *launches_out += 1;
  }
  return LTX2_OK;
}

// per-launch detail of the last profiled decode: conv launch i -> its time (ms) and algorithmic FLOPs
int ltx2_vae_profile_launch(LtxVae* e, int32_t i, double* ms_out, double* flops_out) {
  LTX2_REQUIRE(e && ms_out && flops_out && i >= 0 && size_t(i) * 2 < e->prof.used, "vae_profile_launch: bad index");
  LTX2_CUDA_CHECK(cudaDeviceSynchronize());
  float ms = 0.f;
  LTX2_CUDA_CHECK(cudaEventElapsedTime(&ms, e->prof.events[2 * i], e->prof.events[2 * i + 1]));
  *ms_out = ms;
  *flops_out = e->prof.flops[i];
  return LTX2_OK;
}

int ltx2_blend_chunk(float* dst, const float* src, int32_t BC, int32_t T_dst, int32_t T_src, int32_t HW, int32_t t0,
                     int32_t overlap, void* stream) {
  return blend_chunk(dst, src, BC, T_dst, T_src, HW, t0, overlap, reinterpret_cast<cudaStream_t>(stream));
}

int ltx2_tile_accumulate(float* out, float* wsum, const float* tile, int32_t BC, int32_t To, int32_t Ho, int32_t Wo,
                         int32_t dt, int32_t dh, int32_t dw, int32_t t0, int32_t h0, int32_t w0, int32_t tt, int32_t th,
                         int32_t tw, const float* mask_t, const float* mask_h, const float* mask_w, void* stream) {
  return tile_accumulate(out, wsum, tile, BC, To, Ho, Wo, dt, dh, dw, t0, h0, w0, tt, th, tw, mask_t, mask_h, mask_w,
                         reinterpret_cast<cudaStream_t>(stream));
}

int ltx2_tile_normalize(float* out, const float* wsum, int32_t BC, int64_t plane, void* stream) {
  return tile_normalize(out, wsum, BC, plane, reinterpret_cast<cudaStream_t>(stream));
}

int ltx2_video_to_uint8(const float* video, uint8_t* out, int32_t T, int32_t H, int32_t W, void* stream) {
  return video_to_uint8(video, out, T, H, W, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
