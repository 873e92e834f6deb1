// Host-side helpers shared by the translation units of libglsdet_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

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

// Programmatic dependent launch for the small kernels of a step: the launch carries programmaticStreamSerialization, the
// kernel starts with pdl_prologue() - it lets ITS dependents start their launch right away (they block in their own
// griddepcontrol.wait until this grid has completed and flushed) and waits for its own prerequisites before touching
// memory.  Launch latency and the kernel-boundary drain of ~35 tiny launches per step overlap instead of adding up.
// GLSDET_NO_PDL=1 launches them plainly.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  static const bool plain = getenv("GLSDET_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = plain ? 0 : 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface in count_launch (cudaGetLastError)
}
#endif

// cuTensorMapEncodeTiled resolved through the runtime (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();

int device_sm_count();

}  // namespace glsdet
