// common.cu — error reporting, launch accounting and the TMA descriptor cache.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <unordered_map>

#include "mts_internal.h"

namespace mts {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int set_cuda_error(const char* what, cudaError_t e) {
  return set_error(MTS_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

int check_launch(const char* kernel) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(kernel, e);
  return MTS_OK;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
      sms = v;
    else
      sms = 148;
  }
  return sms;
}

// ------------------------------------------------------------------------------------------
// TMA descriptors.  libcuda is NOT linked (the library must load on GPU-less hosts for the ABI
// tests); cuTensorMapEncodeTiled is resolved through the runtime on first use.
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault,
                                         &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct TmapKey {
  const void* base;
  int64_t k, rows, batch, ld, bstride;
  int box_k, box_rows;
  int64_t elem_bytes;
  bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& key) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&key);
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) {
      h ^= w[i] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    }
    return static_cast<size_t>(h);
  }
};
static_assert(sizeof(TmapKey) % 8 == 0, "TmapKey must hash as whole words");

static std::mutex g_tmap_mu;
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmaps;

int get_tmap_bf16_3d(CUtensorMap* out, const void* base, int64_t k, int64_t rows, int64_t batch,
                     int64_t ld, int64_t batch_stride, int box_k, int box_rows) {
  return get_tmap_3d(out, base, k, rows, batch, ld, batch_stride, box_k, box_rows, 2);
}

int get_tmap_3d(CUtensorMap* out, const void* base, int64_t k, int64_t rows, int64_t batch,
                int64_t ld, int64_t batch_stride, int box_k, int box_rows, int elem_bytes) {
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base; key.k = k; key.rows = rows; key.batch = batch; key.ld = ld;
  key.bstride = batch_stride; key.box_k = box_k; key.box_rows = box_rows; key.elem_bytes = elem_bytes;
  {
    std::lock_guard<std::mutex> lock(g_tmap_mu);
    auto it = g_tmaps.find(key);
    if (it != g_tmaps.end()) { *out = it->second; return MTS_OK; }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(MTS_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * elem_bytes, (cuuint64_t)batch_stride * elem_bytes};
  cuuint32_t box[3] = {(cuuint32_t)box_k, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                  const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(MTS_ERR_CUDA,
                     "cuTensorMapEncodeTiled failed (CUresult %d) base=%p k=%lld rows=%lld "
                     "batch=%lld ld=%lld bstride=%lld box=%dx%d",
                     (int)r, base, (long long)k, (long long)rows, (long long)batch, (long long)ld,
                     (long long)batch_stride, box_k, box_rows);
  {
    std::lock_guard<std::mutex> lock(g_tmap_mu);
    if (g_tmaps.size() > 8192) g_tmaps.clear();
    g_tmaps.emplace(key, *out);
  }
  return MTS_OK;
}

}  // namespace mts

using namespace mts;

extern "C" int mts_version(void) { return MTS_ABI_VERSION; }
extern "C" const char* mts_last_error(void) { return g_err; }
extern "C" int64_t mts_launch_count(void) { return g_launches.load(); }
extern "C" int mts_clear_caches(void) {
  std::lock_guard<std::mutex> lock(g_tmap_mu);
  g_tmaps.clear();
  return MTS_OK;
}
