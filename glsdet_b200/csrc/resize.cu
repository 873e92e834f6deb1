// Image.resize(size, Image.BICUBIC) + letterbox paste of the YOLO facade (yolox-drone/models/core/utils.py:21-34,
// called from yolo.py:130) on the device, bit-exact to Pillow's 8-bit resampler (third party, not under /root/reference:
// src/libImaging/Resample.c): per output pixel a window of bicubic weights (a = -0.5, support 2 * max(scale, 1)),
// normalised in double, rounded to 22-bit fixed point; horizontal pass into a uint8 intermediate, then the vertical pass;
// every sum starts at 2^21 and is clipped with (sum >> 22) to 0..255.  The coefficient tables are built on the host
// (glsdet_pil_bicubic_table, plain C double arithmetic like Pillow's precompute_coeffs / normalize_coeffs_8bpc).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/glsdet_b200.h"
#include "common.h"

namespace glsdet {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

inline double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// One pass over a uint8 HWC image: out[y][x][c] = clip8(2^21 + sum_k in[...][c] * kk[.][k]).
// HORIZONTAL: window over columns of the same row.  VERTICAL: window over rows of the same column; the result is written
// at (off_y + y, off_x + x) of a larger canvas (the letterbox paste).
template <bool HORIZONTAL>
__global__ void __launch_bounds__(256) resample_kernel(const uint8_t* __restrict__ in, int in_h, int in_w,
                                                       uint8_t* __restrict__ out, int out_h, int out_w, int can_w,
                                                       int off_y, int off_x, const int32_t* __restrict__ bounds,
                                                       const int32_t* __restrict__ kk, int ksize) {
  pdl_prologue();
  const int64_t total = static_cast<int64_t>(out_h) * out_w;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int y = static_cast<int>(i / out_w), x = static_cast<int>(i - static_cast<int64_t>(y) * out_w);
    const int o = HORIZONTAL ? x : y;
    const int lo = bounds[2 * o], n = bounds[2 * o + 1];
    const int32_t* k = kk + static_cast<int64_t>(o) * ksize;
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    for (int t = 0; t < n; ++t) {
      const uint8_t* p = HORIZONTAL ? in + (static_cast<int64_t>(y) * in_w + lo + t) * 3
                                    : in + (static_cast<int64_t>(lo + t) * in_w + x) * 3;
      const int w = k[t];
      s0 += p[0] * w; s1 += p[1] * w; s2 += p[2] * w;
    }
    uint8_t* q = out + (static_cast<int64_t>(off_y + y) * can_w + off_x + x) * 3;
    q[0] = clip8(s0); q[1] = clip8(s1); q[2] = clip8(s2);
  }
}

__global__ void __launch_bounds__(256) fill_u8_kernel(uint8_t* __restrict__ p, int64_t n, uint8_t v) {
  pdl_prologue();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    p[i] = v;
}

__global__ void __launch_bounds__(256) paste_u8_kernel(const uint8_t* __restrict__ in, int in_h, int in_w, uint8_t* __restrict__ out,
                                                       int can_w, int off_y, int off_x) {
  pdl_prologue();
  const int64_t total = static_cast<int64_t>(in_h) * in_w * 3;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t px = i / 3;
    const int c = static_cast<int>(i - px * 3);
    const int y = static_cast<int>(px / in_w), x = static_cast<int>(px - static_cast<int64_t>(y) * in_w);
    out[(static_cast<int64_t>(off_y + y) * can_w + off_x + x) * 3 + c] = in[i];
  }
}

}  // namespace
}  // namespace glsdet

using namespace glsdet;

extern "C" int glsdet_pil_bicubic_ksize(int32_t in_size, int32_t out_size) {
  if (in_size <= 0 || out_size <= 0) return 0;
  double filterscale = static_cast<double>(in_size) / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  return static_cast<int>(ceil(support)) * 2 + 1;
}

extern "C" int glsdet_pil_bicubic_table(int32_t in_size, int32_t out_size, int32_t* bounds, int32_t* kk) {
  GLSDET_REQUIRE(in_size > 0 && out_size > 0 && bounds && kk, "pil_bicubic_table: bad arguments");
  // Resample.c precompute_coeffs with box = (0, in_size)
  const double in0 = 0.0, in1 = static_cast<double>(in_size);
  double scale, filterscale;
  scale = filterscale = (in1 - in0) / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  const int ksize = static_cast<int>(ceil(support)) * 2 + 1;
  double* k = static_cast<double*>(malloc(sizeof(double) * ksize));
  GLSDET_REQUIRE(k != nullptr, "pil_bicubic_table: out of host memory");
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = in0 + (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    int x = 0;
    for (; x < xmax; ++x) {
      const double w = bicubic_filter((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    for (; x < ksize; ++x) k[x] = 0;
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
    // normalize_coeffs_8bpc
    for (x = 0; x < ksize; ++x) {
      const double v = k[x];
      kk[static_cast<int64_t>(xx) * ksize + x] = v < 0 ? static_cast<int>(-0.5 + v * (1 << kPrecisionBits))
                                                       : static_cast<int>(0.5 + v * (1 << kPrecisionBits));
    }
  }
  free(k);
  return 0;
}

extern "C" int glsdet_resize_bicubic_u8(const uint8_t* image, int32_t in_h, int32_t in_w, uint8_t* canvas, int32_t can_h,
                                        int32_t can_w, int32_t out_h, int32_t out_w, int32_t off_y, int32_t off_x,
                                        int32_t fill, uint8_t* tmp, const int32_t* bounds_h, const int32_t* kk_h,
                                        int32_t ksize_h, const int32_t* bounds_v, const int32_t* kk_v, int32_t ksize_v,
                                        void* stream) {
  GLSDET_REQUIRE(image && canvas && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0, "resize_bicubic_u8: bad arguments");
  GLSDET_REQUIRE(off_y >= 0 && off_x >= 0 && off_y + out_h <= can_h && off_x + out_w <= can_w,
                 "resize_bicubic_u8: the resized image does not fit the canvas");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int sms = device_sm_count();
  if (out_h != can_h || out_w != can_w)   // letterbox: Image.new('RGB', size, (128, 128, 128))
    launch_pdl(fill_u8_kernel, dim3(2 * sms), dim3(256), 0, st, canvas, static_cast<int64_t>(can_h) * can_w * 3,
               static_cast<uint8_t>(fill));
  // Pillow: the horizontal pass runs iff the width changes, then the vertical pass iff the height changes
  const bool need_h = out_w != in_w, need_v = out_h != in_h;
  const uint8_t* src = image;
  int cur_w = in_w;
  if (need_h) {
    GLSDET_REQUIRE(bounds_h && kk_h && ksize_h > 0, "resize_bicubic_u8: missing horizontal table");
    GLSDET_REQUIRE(!need_v || tmp != nullptr, "resize_bicubic_u8: missing intermediate buffer");
    uint8_t* dst = need_v ? tmp : canvas;
    launch_pdl(resample_kernel<true>, dim3(4 * sms), dim3(256), 0, st, src, in_h, in_w, dst, in_h, out_w,
               need_v ? out_w : can_w, need_v ? 0 : off_y, need_v ? 0 : off_x, bounds_h, kk_h, ksize_h);
    count_launch("resample_kernel<horizontal>");
    src = dst;
    cur_w = out_w;
  }
  if (need_v) {
    GLSDET_REQUIRE(bounds_v && kk_v && ksize_v > 0, "resize_bicubic_u8: missing vertical table");
    launch_pdl(resample_kernel<false>, dim3(4 * sms), dim3(256), 0, st, src, in_h, cur_w, canvas, out_h, out_w, can_w, off_y,
               off_x, bounds_v, kk_v, ksize_v);
    return count_launch("resample_kernel<vertical>");
  }
  if (!need_h) {   // same size: Image.resize returns a copy
    launch_pdl(paste_u8_kernel, dim3(4 * sms), dim3(256), 0, st, image, in_h, in_w, canvas, can_w, off_y, off_x);
    return count_launch("paste_u8_kernel");
  }
  return 0;
}
