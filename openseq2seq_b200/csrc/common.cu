#include "common.h"

#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace os2s {

static thread_local std::string g_err;

void set_error(const std::string& msg) { g_err = msg; }
const char* last_error_cstr() { return g_err.c_str(); }

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return OK;
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return ERR_CUDA;
}

int check_launch(const char* what) { return check_cuda(cudaGetLastError(), what); }

int device_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

// ------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e =
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
  }
  return fn;
}

struct TmapKey {
  uint64_t v[12];
  bool operator==(const TmapKey& o) const { return std::memcmp(v, o.v, sizeof(v)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 12; ++i) {
      h ^= k.v[i];
      h *= 1099511628211ull;
    }
    return (size_t)h;
  }
};

const CUtensorMap* get_tmap_bf16(const void* base, int rank, const uint64_t* dims,
                                 const uint64_t* strides_bytes, const uint32_t* box) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap*, TmapKeyHash> cache;
  TmapKey key;
  std::memset(&key, 0, sizeof(key));
  key.v[0] = (uint64_t)base;
  key.v[1] = (uint64_t)rank;
  for (int i = 0; i < rank; ++i) {
    key.v[2 + i] = dims[i];
    key.v[8 + i] = box[i];
  }
  for (int i = 0; i + 1 < rank; ++i) key.v[5 + i] = strides_bytes[i];
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;

  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled not available (no CUDA driver?)");
    return nullptr;
  }
  // 64-byte aligned storage, never freed (maps live as long as the process).
  CUtensorMap* m = nullptr;
  if (posix_memalign((void**)&m, 64, sizeof(CUtensorMap)) != 0) {
    set_error("posix_memalign failed");
    return nullptr;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                  gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    free(m);
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return nullptr;
  }
  cache.emplace(key, m);
  return m;
}

}  // namespace os2s
