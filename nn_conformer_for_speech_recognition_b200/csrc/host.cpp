// Host-side plumbing of the C-ABI library: thread-local error string, TMA tensor-map cache, device info.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "b200conformer.h"
#include "common.cuh"

namespace b2c {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return static_cast<int>(e);
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

struct MapKey {
  const void* base;
  uint64_t inner, outer, ld;
  uint32_t bi, bo, eb, swz;
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(MapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return static_cast<size_t>(h);
  }
};
static_assert(sizeof(MapKey) == 48, "MapKey must be padding-free");

// 3-D fp32/bf16 map (dims inner -> outer d0, d1, d2; byte strides of d1 and d2), box {b0, b1, 1}.  Small per-process cache.
int get_tensor_map_3d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2,
                      uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, bool swizzle128) {
  struct Key3 {
    const void* base; uint64_t d0, d1, d2, s1, s2; uint32_t b0, b1, eb, swz;
    bool operator==(const Key3& o) const { return memcmp(this, &o, sizeof(Key3)) == 0; }
  };
  static std::mutex mu;
  static std::vector<std::pair<Key3, CUtensorMap>> cache;
  Key3 key;
  memset(&key, 0, sizeof(key));
  key.base = base; key.d0 = d0; key.d1 = d1; key.d2 = d2; key.s1 = stride1_bytes; key.s2 = stride2_bytes;
  key.b0 = b0; key.b1 = b1; key.eb = static_cast<uint32_t>(elem_bytes); key.swz = swizzle128;
  {
    std::lock_guard<std::mutex> g(mu);
    for (auto& e : cache)
      if (e.first == key) { *out = e.second; return 0; }
  }
  EncodeTiledFn fn = encode_fn();
  B2C_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  B2C_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && stride1_bytes % 16 == 0 && stride2_bytes % 16 == 0,
              "TMA needs 16-byte aligned base and pitches");
  B2C_REQUIRE(b0 * elem_bytes <= 128 || !swizzle128, "TMA box inner extent exceeds the 128B swizzle span");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(out, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B2C_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (3-D) failed with CUresult %d", static_cast<int>(r));
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 64) cache.clear();
    cache.emplace_back(key, *out);
  }
  return 0;
}

int get_tensor_map_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t inner, uint64_t outer,
                      uint64_t ld_bytes, uint32_t box_inner, uint32_t box_outer, bool swizzle128) {
  return get_tensor_map_2d_sw(out, base, elem_bytes, inner, outer, ld_bytes, box_inner, box_outer, swizzle128 ? 1 : 0);
}

int get_tensor_map_2d_sw(CUtensorMap* out, const void* base, int elem_bytes, uint64_t inner, uint64_t outer,
                         uint64_t ld_bytes, uint32_t box_inner, uint32_t box_outer, int swizzle) {
  const bool swizzle128 = swizzle == 1;
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base; key.inner = inner; key.outer = outer; key.ld = ld_bytes;
  key.bi = box_inner; key.bo = box_outer; key.eb = static_cast<uint32_t>(elem_bytes); key.swz = static_cast<uint32_t>(swizzle);
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn fn = encode_fn();
  B2C_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  B2C_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && ld_bytes % 16 == 0,
              "TMA needs 16-byte aligned base and pitch (pitch=%llu)", (unsigned long long)ld_bytes);
  B2C_REQUIRE(box_inner * elem_bytes <= 128 || !swizzle128, "TMA box inner extent exceeds the 128B swizzle span");
  B2C_REQUIRE(swizzle != 2 || box_inner * elem_bytes <= 64, "TMA box inner extent exceeds the 64B swizzle span");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : (swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B2C_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu ld=%llu box=%ux%u)",
              static_cast<int>(r), (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld_bytes,
              box_inner, box_outer);
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > (1u << 16)) cache.clear();
    cache.emplace(key, *out);
  }
  return 0;
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("B2C_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

}  // namespace b2c

extern "C" const char* b2c_last_error(void) { return b2c::g_err; }

extern "C" int b2c_version(void) { return B2C_VERSION; }

extern "C" int b2c_device_check(void) {
  int dev = 0;
  B2C_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  B2C_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  B2C_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  B2C_REQUIRE(major == 10, "this library is built for sm_100a only; device reports sm_%d%d", major, minor);
  return 0;
}

