// CSPDarknet pieces that are not convolutions (yolox-drone/models/ffa/darknet.py):
//   * Focus space-to-depth (darknet.py:15-21) fused with the NCHW fp32 -> NHWC bf16 layout change of the image,
//   * the three stride-1 max pools of SPPBottleneck (darknet.py:28,33-36) as a cascade of separable 5-wide maxima.
// Both are HBM / shared-memory bound byte movers; the convolutions of the backbone run on conv_gemm_kernel.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <string.h>

#include <type_traits>

#include "../../include/glsdet_b200.h"
#include "common.h"
#include "ptx.cuh"

namespace glsdet {

// ---------------------------------------------------------------------------------------------- Focus
// dst[b, y, x, q*3 + c] = img[b, c, 2y + dy(q), 2x + dx(q)],  q = 0: top-left, 1: bottom-left, 2: top-right,
// 3: bottom-right (the torch.cat order of darknet.py:16-20); channels 12..15 are zero so that a pixel is 32 bytes.
// One thread per output pixel: six coalesced float2 loads (3 channels x 2 rows), two 16-byte stores.
__global__ void __launch_bounds__(256) focus_kernel(const float* __restrict__ img, uint16_t* __restrict__ dst, int H, int W,
                                                    int border, int64_t total, int f16) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int Wo = W >> 1, Ho = H >> 1;
  const int x = static_cast<int>(idx % Wo);
  const int64_t t = idx / Wo;
  const int y = static_cast<int>(t % Ho);
  const int64_t b = t / Ho;
  const int64_t plane = static_cast<int64_t>(H) * W;
  const float* p = img + b * 3 * plane + static_cast<int64_t>(2 * y) * W + 2 * x;
  float tl[3], bl[3], tr[3], br[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float2 r0 = __ldg(reinterpret_cast<const float2*>(p + c * plane));
    const float2 r1 = __ldg(reinterpret_cast<const float2*>(p + c * plane + W));
    tl[c] = r0.x; tr[c] = r0.y; bl[c] = r1.x; br[c] = r1.y;
  }
  __align__(16) uint16_t o[16];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    o[c] = to_16(tl[c], f16);
    o[3 + c] = to_16(bl[c], f16);
    o[6 + c] = to_16(tr[c], f16);
    o[9 + c] = to_16(br[c], f16);
  }
#pragma unroll
  for (int c = 12; c < 16; ++c) o[c] = 0;
  // border = 1: rows of Wo + 2 pixels, pixel x lands at x + 1 (zero border pixels left and right)
  uint4* out = reinterpret_cast<uint4*>(dst + ((b * Ho + y) * (Wo + 2 * border) + x + border) * 16);
  out[0] = reinterpret_cast<const uint4*>(o)[0];
  out[1] = reinterpret_cast<const uint4*>(o)[1];
}

// uint8 HWC image (what the camera / decoder delivers) -> the same Focus output, with the reference's normalisation
// (models/core/utils.py:47-51 preprocess_input: float32 x /= 255.0; then -= mean and /= std against float64 arrays, i.e.
// computed in double and rounded to float32 each time) applied on the fly: bit-identical to normalising on the host and
// calling focus_kernel, at a quarter of the upload bytes.
// The input has 256 values per channel, so the host tabulates the three statements once (IEEE float / double arithmetic,
// the same as numpy's) and the kernel only looks the bf16 results up.
struct NormLut {
  uint16_t v[3][256];   // bf16 (or fp16) bits of ((float(u) / 255.0f) - mean[c]) / std[c]
};
__global__ void __launch_bounds__(256) focus_u8_kernel(const uint8_t* __restrict__ img, uint16_t* __restrict__ dst, int H,
                                                       int W, int border, const __grid_constant__ NormLut lut, int64_t total_pairs) {
  // one thread = two horizontally adjacent output pixels: 12 contiguous bytes per input row (three aligned 32-bit loads),
  // 24 table look-ups, 64 contiguous output bytes; persistent CTAs amortise the table copy
  __shared__ uint16_t s_lut[3][256];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) s_lut[i >> 8][i & 255] = lut.v[i >> 8][i & 255];
  __syncthreads();
  const int Wo = W >> 1, Ho = H >> 1, Wp = Wo >> 1;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total_pairs;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int xp = static_cast<int>(idx % Wp);
    const int64_t t = idx / Wp;
    const int y = static_cast<int>(t % Ho);
    const int64_t b = t / Ho;
    const uint8_t* p = img + ((b * H + 2 * y) * W + 4 * xp) * 3;   // [B, H, W, 3]; 12 bytes per row, 4-byte aligned
    __align__(16) uint16_t o[32];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const uint32_t* row = reinterpret_cast<const uint32_t*>(p + static_cast<int64_t>(dy) * W * 3);
      const uint32_t w[3] = {__ldg(row), __ldg(row + 1), __ldg(row + 2)};
#pragma unroll
      for (int k = 0; k < 12; ++k) {   // byte k = input pixel (k / 3) of the row, channel k % 3
        const uint32_t u = (w[k >> 2] >> ((k & 3) * 8)) & 255u;
        const int px = k / 3, c = k % 3;
        // output pixel px >> 1, Focus slot q = (px & 1) * 2 + dy (top-left, bottom-left, top-right, bottom-right)
        o[(px >> 1) * 16 + ((px & 1) * 2 + dy) * 3 + c] = s_lut[c][u];
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int c = 12; c < 16; ++c) o[h * 16 + c] = 0;
    uint4* out = reinterpret_cast<uint4*>(dst + ((b * Ho + y) * (Wo + 2 * border) + 2 * xp + border) * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) out[q] = reinterpret_cast<const uint4*>(o)[q];
  }
}

// fp32 accuracy mode: dst [B, H/2, W/2, 12] fp32 NHWC, same channel order, no padding channels
__global__ void __launch_bounds__(256) focus_f32_kernel(const float* __restrict__ img, float* __restrict__ dst, int H, int W,
                                                        int64_t total) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int Wo = W >> 1, Ho = H >> 1;
  const int x = static_cast<int>(idx % Wo);
  const int64_t t = idx / Wo;
  const int y = static_cast<int>(t % Ho);
  const int64_t b = t / Ho;
  const int64_t plane = static_cast<int64_t>(H) * W;
  const float* p = img + b * 3 * plane + static_cast<int64_t>(2 * y) * W + 2 * x;
  float o[12];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float2 r0 = __ldg(reinterpret_cast<const float2*>(p + c * plane));
    const float2 r1 = __ldg(reinterpret_cast<const float2*>(p + c * plane + W));
    o[c] = r0.x; o[3 + c] = r1.x; o[6 + c] = r0.y; o[9 + c] = r1.y;
  }
  float4* out = reinterpret_cast<float4*>(dst + idx * 12);
#pragma unroll
  for (int q = 0; q < 3; ++q) out[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
}

// ---------------------------------------------------------------------------------------------- SPP max pools
__device__ __forceinline__ uint4 max8(uint4 a, uint4 b) {
  uint4 r;
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}

__device__ __forceinline__ uint4 max4f(uint4 a, uint4 b) {   // four fp32 values
  uint4 r;
  r.x = __float_as_uint(fmaxf(__uint_as_float(a.x), __uint_as_float(b.x)));
  r.y = __float_as_uint(fmaxf(__uint_as_float(a.y), __uint_as_float(b.y)));
  r.z = __float_as_uint(fmaxf(__uint_as_float(a.z), __uint_as_float(b.z)));
  r.w = __float_as_uint(fmaxf(__uint_as_float(a.w), __uint_as_float(b.w)));
  return r;
}
__device__ __forceinline__ uint4 max8h(uint4 a, uint4 b) {   // eight fp16 values
  uint4 r;
  __half2* pr = reinterpret_cast<__half2*>(&r);
  const __half2* pa = reinterpret_cast<const __half2*>(&a);
  const __half2* pb = reinterpret_cast<const __half2*>(&b);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}
// KIND: 0 = bf16, 1 = fp32 (accuracy mode), 2 = fp16
template <int KIND>
__device__ __forceinline__ uint4 vmax(uint4 a, uint4 b) { return KIND == 1 ? max4f(a, b) : KIND == 2 ? max8h(a, b) : max8(a, b); }

// One CTA = one image x 8 channels (bf16; 4 channels in the fp32 accuracy mode: 16-byte vectors either way), the whole h x w map in shared memory (two ping-pong planes of 16-byte vectors).
// MaxPool2d(k, 1, k/2) pads with -inf, i.e. the maximum runs over the in-bounds part of the window, and
// pool9 = pool5(pool5), pool13 = pool5(pool9) exactly (maxima of nested windows), so three rounds of a separable
// 5-wide maximum (row pass, column pass) produce the three outputs.
template <int KIND>
__global__ void __launch_bounds__(1024) spp_pool_kernel(void* __restrict__ buf_v, int h, int w, int ld, int coff_src,
                                                        int coff5, int coff9, int coff13, int groups) {
  using T = typename std::conditional<KIND == 1, float, uint16_t>::type;
  constexpr int kVec = KIND == 1 ? 4 : 8;
  T* buf = reinterpret_cast<T*>(buf_v);
  extern __shared__ uint4 smem_pool[];
  uint4* a = smem_pool;
  uint4* t = smem_pool + h * w;
  const int g = blockIdx.x % groups;
  const int64_t b = blockIdx.x / groups;
  T* base = buf + b * h * w * ld + g * kVec;
  const int n = h * w;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    a[i] = *reinterpret_cast<const uint4*>(base + static_cast<int64_t>(i) * ld + coff_src);
  __syncthreads();
  for (int round = 0; round < 3; ++round) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {   // row pass: a -> t
      const int y = i / w, x = i - y * w;
      uint4 m = a[i];
#pragma unroll
      for (int d = -2; d <= 2; ++d) {
        const int xx = x + d;
        if (d != 0 && xx >= 0 && xx < w) m = vmax<KIND>(m, a[y * w + xx]);
      }
      t[i] = m;
    }
    __syncthreads();
    const int coff = round == 0 ? coff5 : round == 1 ? coff9 : coff13;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {   // column pass: t -> a (+ global)
      const int y = i / w, x = i - y * w;
      uint4 m = t[i];
#pragma unroll
      for (int d = -2; d <= 2; ++d) {
        const int yy = y + d;
        if (d != 0 && yy >= 0 && yy < h) m = vmax<KIND>(m, t[yy * w + x]);
      }
      *reinterpret_cast<uint4*>(base + static_cast<int64_t>(i) * ld + coff) = m;
      // every thread rewrites only its own elements of `a`, and nobody reads `a` during this pass
      a[i] = m;
    }
    __syncthreads();
  }
}

}  // namespace glsdet

extern "C" int glsdet_focus_nchw_f32_to_nhwc_bf16(const float* image, void* dst, int32_t batch, int32_t height,
                                                  int32_t width, int32_t dst_border, void* stream) {
  return glsdet_focus_nchw_f32_to_nhwc_16(image, dst, batch, height, width, dst_border, GLSDET_DT_BF16, stream);
}

extern "C" int glsdet_focus_nchw_f32_to_nhwc_16(const float* image, void* dst, int32_t batch, int32_t height,
                                                int32_t width, int32_t dst_border, int32_t dtype, void* stream) {
  GLSDET_REQUIRE(dtype == GLSDET_DT_BF16 || dtype == GLSDET_DT_F16, "focus: bad storage dtype");
  GLSDET_REQUIRE(image && dst && batch > 0 && height > 0 && width > 0 && (dst_border == 0 || dst_border == 1),
                 "focus: bad arguments");
  GLSDET_REQUIRE((height % 2) == 0 && (width % 2) == 0, "focus: height and width must be even (got %d x %d)", height, width);
  GLSDET_REQUIRE((reinterpret_cast<uintptr_t>(image) & 7) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
                 "focus: image must be 8-byte aligned and dst 16-byte aligned");
  const int64_t total = static_cast<int64_t>(batch) * (height / 2) * (width / 2);
  glsdet::focus_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      image, reinterpret_cast<uint16_t*>(dst), height, width, dst_border, total, static_cast<int>(dtype));
  return glsdet::count_launch("focus_kernel");
}

namespace {
template <int KIND>
int spp_maxpool_impl(void* buf, int32_t batch, int32_t height, int32_t width, int32_t channels, int32_t ld, int32_t src_coff,
                     int32_t coff5, int32_t coff9, int32_t coff13, void* stream) {
  constexpr int kVec = KIND == 1 ? 4 : 8;
  GLSDET_REQUIRE(buf && batch > 0 && height > 0 && width > 0 && channels > 0, "spp_maxpool: bad arguments");
  GLSDET_REQUIRE((channels % kVec) == 0 && (ld % kVec) == 0 && (src_coff % kVec) == 0 && (coff5 % kVec) == 0 &&
                     (coff9 % kVec) == 0 && (coff13 % kVec) == 0,
                 "spp_maxpool: channels, pitch and offsets must be multiples of %d", kVec);
  const int32_t offs[4] = {src_coff, coff5, coff9, coff13};
  for (int i = 0; i < 4; ++i) {
    GLSDET_REQUIRE(offs[i] >= 0 && offs[i] + channels <= ld, "spp_maxpool: window %d exceeds the pitch", i);
    for (int j = 0; j < i; ++j)
      GLSDET_REQUIRE(offs[i] >= offs[j] + channels || offs[j] >= offs[i] + channels, "spp_maxpool: windows %d and %d overlap", j, i);
  }
  const size_t smem = static_cast<size_t>(height) * width * 2 * sizeof(uint4);
  GLSDET_REQUIRE(smem <= 200 * 1024, "spp_maxpool: a %d x %d map does not fit shared memory (max 6400 pixels)", height, width);
  static bool attr_set[64] = {false};
  int dev = 0;
  GLSDET_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !attr_set[dev]) {
    GLSDET_CHECK_CUDA(cudaFuncSetAttribute(glsdet::spp_pool_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set[dev] = true;
  }
  const int groups = channels / kVec;
  const int pixels = height * width;
  const int threads = pixels >= 1024 ? 1024 : ((pixels + 31) / 32) * 32;
  glsdet::spp_pool_kernel<KIND><<<static_cast<unsigned>(batch) * groups, threads, smem, static_cast<cudaStream_t>(stream)>>>(
      buf, height, width, ld, src_coff, coff5, coff9, coff13, groups);
  return glsdet::count_launch("spp_pool_kernel");
}
}  // namespace

extern "C" int glsdet_spp_maxpool(void* buf, int32_t batch, int32_t height, int32_t width, int32_t channels, int32_t ld,
                                  int32_t src_coff, int32_t coff5, int32_t coff9, int32_t coff13, void* stream) {
  return spp_maxpool_impl<0>(buf, batch, height, width, channels, ld, src_coff, coff5, coff9, coff13, stream);
}

extern "C" int glsdet_spp_maxpool_16(void* buf, int32_t batch, int32_t height, int32_t width, int32_t channels, int32_t ld,
                                     int32_t src_coff, int32_t coff5, int32_t coff9, int32_t coff13, int32_t dtype,
                                     void* stream) {
  GLSDET_REQUIRE(dtype == GLSDET_DT_BF16 || dtype == GLSDET_DT_F16, "spp_maxpool: bad storage dtype");
  if (dtype == GLSDET_DT_F16)
    return spp_maxpool_impl<2>(buf, batch, height, width, channels, ld, src_coff, coff5, coff9, coff13, stream);
  return spp_maxpool_impl<0>(buf, batch, height, width, channels, ld, src_coff, coff5, coff9, coff13, stream);
}

extern "C" int glsdet_spp_maxpool_f32(float* buf, int32_t batch, int32_t height, int32_t width, int32_t channels, int32_t ld,
                                      int32_t src_coff, int32_t coff5, int32_t coff9, int32_t coff13, void* stream) {
  return spp_maxpool_impl<1>(buf, batch, height, width, channels, ld, src_coff, coff5, coff9, coff13, stream);
}

extern "C" int glsdet_focus_nchw_f32_to_nhwc_f32(const float* image, float* dst, int32_t batch, int32_t height, int32_t width,
                                                 void* stream) {
  GLSDET_REQUIRE(image && dst && batch > 0 && height > 0 && width > 0, "focus_f32: bad arguments");
  GLSDET_REQUIRE((height % 2) == 0 && (width % 2) == 0, "focus_f32: height and width must be even (got %d x %d)", height, width);
  GLSDET_REQUIRE((reinterpret_cast<uintptr_t>(image) & 7) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
                 "focus_f32: image must be 8-byte aligned and dst 16-byte aligned");
  const int64_t total = static_cast<int64_t>(batch) * (height / 2) * (width / 2);
  glsdet::focus_f32_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      image, dst, height, width, total);
  return glsdet::count_launch("focus_f32_kernel");
}

extern "C" int glsdet_focus_u8_to_nhwc_bf16(const uint8_t* image, void* dst, int32_t batch, int32_t height, int32_t width,
                                            int32_t dst_border, const double* mean, const double* std, void* stream) {
  return glsdet_focus_u8_to_nhwc_16(image, dst, batch, height, width, dst_border, mean, std, GLSDET_DT_BF16, stream);
}

extern "C" int glsdet_focus_u8_to_nhwc_16(const uint8_t* image, void* dst, int32_t batch, int32_t height, int32_t width,
                                          int32_t dst_border, const double* mean, const double* std, int32_t dtype,
                                          void* stream) {
  GLSDET_REQUIRE(dtype == GLSDET_DT_BF16 || dtype == GLSDET_DT_F16, "focus_u8: bad storage dtype");
  GLSDET_REQUIRE(image && dst && mean && std && batch > 0 && height > 0 && width > 0 && (dst_border == 0 || dst_border == 1),
                 "focus_u8: bad arguments");
  GLSDET_REQUIRE((height % 2) == 0 && (width % 2) == 0, "focus_u8: height and width must be even (got %d x %d)", height, width);
  GLSDET_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 15) == 0, "focus_u8: dst must be 16-byte aligned");
  glsdet::NormLut lut;
  for (int c = 0; c < 3; ++c) {
    GLSDET_REQUIRE(std[c] != 0.0, "focus_u8: std[%d] is zero", c);
    for (int u = 0; u < 256; ++u) {
      volatile float v = static_cast<float>(u) / 255.0f;                       // image /= 255.0          (float32)
      v = static_cast<float>(static_cast<double>(v) - mean[c]);                // image -= float64 array  (double, rounded)
      v = static_cast<float>(static_cast<double>(v) / std[c]);                 // image /= float64 array
      const float f = v;
      if (dtype == GLSDET_DT_F16) {
        lut.v[c][u] = __half_as_ushort(__float2half_rn(f));                    // round to nearest even
      } else {
        uint32_t bits;
        memcpy(&bits, &f, 4);
        bits += 0x7FFFu + ((bits >> 16) & 1u);                                 // round to nearest even (finite values)
        lut.v[c][u] = static_cast<uint16_t>(bits >> 16);
      }
    }
  }
  GLSDET_REQUIRE((width % 4) == 0 && (reinterpret_cast<uintptr_t>(image) & 3) == 0,
                 "focus_u8: width must be a multiple of 4 and the image 4-byte aligned");
  const int64_t total_pairs = static_cast<int64_t>(batch) * (height / 2) * (width / 4);
  int64_t blocks = (total_pairs + 255) / 256;
  const int64_t cap = static_cast<int64_t>(glsdet::device_sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  glsdet::focus_u8_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      image, reinterpret_cast<uint16_t*>(dst), height, width, dst_border, lut, total_pairs);
  return glsdet::count_launch("focus_u8_kernel");
}
