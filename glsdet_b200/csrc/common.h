// Host-side helpers shared by the translation units of libglsdet_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

namespace glsdet {

// thread-local message of the last failing ABI call (glsdet_last_error)
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launch_count;

#define GLSDET_CHECK_CUDA(expr)                                                              \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      glsdet::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

#define GLSDET_REQUIRE(cond, ...)      \
  do {                                 \
    if (!(cond)) {                     \
      glsdet::set_error(__VA_ARGS__);  \
      return 2;                        \
    }                                  \
  } while (0)

inline int count_launch(const char* what) {
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

// cuTensorMapEncodeTiled resolved through the runtime (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();

int device_sm_count();

}  // namespace glsdet
