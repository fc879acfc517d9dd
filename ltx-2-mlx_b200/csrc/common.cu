// Host-side helpers: last-error string, tensor-map creation through the driver entry
// point (no libcuda link dependency, so the library loads on a GPU-less host), a small
// tensor-map cache and device queries.
#include <stdlib.h>

#include "common.cuh"

#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <tuple>

namespace ltx2 {

static thread_local char g_err[512] = "";

const char* last_error() { return g_err; }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tensor_map_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  return make_tensor_map(out, base, 2, rank, dims, strides_bytes, box, swizzle128);
}

int make_tensor_map(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return LTX2_ERR_CUDA;
  }
  cuuint64_t gdims[5], gstr[4];
  cuuint32_t gbox[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(out, elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdims, gstr, gbox, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): base=%p rank=%d dims=[%llu,%llu] stride=%llu box=[%u,%u]", (int)r,
              base, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0], rank > 1 ? box[1] : 0);
    return LTX2_ERR_INVALID;
  }
  return LTX2_OK;
}

int get_tensor_map_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, int elem_bytes) {
  // Encoding a tensor map costs a driver call, so the maps of the (static) weight / workspace buffers are cached.
  // The descriptor is COPIED out under the lock: callers never hold a pointer into the cache, so the growth guard
  // below (or a second thread) cannot invalidate a map between two fetches of one launch.
  typedef std::tuple<int, const void*, uint64_t, uint64_t, uint64_t, uint32_t, int> Key;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  Key key(dev, base, rows, cols, ld, box_rows, elem_bytes);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return LTX2_OK;
  }
  if (cache.size() > 16384) cache.clear();  // unbounded growth guard for callers that stream fresh buffers
  uint64_t dims[2] = {cols, rows};
  uint64_t strides[1] = {ld * elem_bytes};
  uint32_t box[2] = {static_cast<uint32_t>(128 / elem_bytes), box_rows};      // 128-byte rows (one swizzle atom)
  CUtensorMap m;
  int s = make_tensor_map(&m, base, elem_bytes, 2, dims, strides, box, true);
  if (s != LTX2_OK) return s;
  cache[key] = m;
  *out = m;
  return LTX2_OK;
}

static std::atomic<int64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("LTX2_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
  return dev;
}

int num_sms() {
  static int n[kMaxDevices] = {0};
  const int dev = current_device();
  if (n[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    n[dev] = v;
  }
  return n[dev];
}

}  // namespace ltx2
