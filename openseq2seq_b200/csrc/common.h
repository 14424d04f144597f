// Shared host-side helpers for libos2s_b200: error reporting and tensor-map cache.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace os2s {

// Status codes returned by every extern "C" entry point (see include/os2s.h).
enum : int { OK = 0, ERR_INVALID = -1, ERR_CUDA = -2, ERR_UNSUPPORTED = -3 };

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
int check_cuda(cudaError_t e, const char* what);
int check_launch(const char* what);

int device_sm_count();

// TMA tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point; the library
// never links libcuda directly so it loads on a host without a driver).
// dims/box are innermost-first; strides_bytes[i] is the byte stride of dim i+1.
// Returns nullptr (and sets the error string) on failure. Maps are cached per
// (pointer, geometry) so a steady-state training step does no encoding.
const CUtensorMap* get_tmap_bf16(const void* base, int rank, const uint64_t* dims,
                                 const uint64_t* strides_bytes, const uint32_t* box);

#define OS2S_CUDA(expr)                                         \
  do {                                                          \
    int _s = ::os2s::check_cuda((expr), #expr);                 \
    if (_s != 0) return _s;                                     \
  } while (0)

}  // namespace os2s
