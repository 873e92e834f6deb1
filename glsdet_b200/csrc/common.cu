#include "common.h"

#include <mutex>

#include "../../include/glsdet_b200.h"

namespace glsdet {

static thread_local char t_error[1024] = "";
std::atomic<int64_t> g_launch_count{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int device_sm_count() {
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    sms[dev] = n;
  }
  return sms[dev];
}

}  // namespace glsdet

extern "C" {
int glsdet_abi_version(void) { return GLSDET_ABI_VERSION; }
const char* glsdet_last_error(void) { return glsdet::t_error; }
int64_t glsdet_launch_count(void) { return glsdet::g_launch_count.load(); }
}
