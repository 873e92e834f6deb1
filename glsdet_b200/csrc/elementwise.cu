// Bandwidth-bound pieces of the path: layout converters at the module boundary, the SE gate of FFA,
// gate * PixelShuffle, and the stand-alone decode of raw NCHW logits.  All 128-bit vectorised on the
// NHWC side, coalesced on the NCHW side (shared-memory transpose).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/glsdet_b200.h"
#include "common.h"
#include "ptx.cuh"

namespace glsdet {

constexpr int kTC = 64;  // channels per transpose tile (64 bf16 = 128 bytes on the NHWC side)
constexpr int kTP = 32;  // pixels per transpose tile   (32 fp32 = 128 bytes on the NCHW side)

// [B, C, HW] fp32 -> [B, HW, ld] bf16 (channels coff..coff+C)
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst,
                                                           int C, int HW, int ld, int coff, int f16) {
  pdl_prologue();
  __shared__ float tile[kTC][kTP + 1];
  const int p0 = blockIdx.x * kTP, c0 = blockIdx.y * kTC, b = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* s = src + static_cast<int64_t>(b) * C * HW;
#pragma unroll
  for (int i = 0; i < kTC / 8; ++i) {
    const int c = c0 + warp * (kTC / 8) + i;
    const int p = p0 + lane;
    tile[warp * (kTC / 8) + i][lane] = (c < C && p < HW) ? __ldg(s + static_cast<int64_t>(c) * HW + p) : 0.0f;
  }
  __syncthreads();
  const int pl = threadIdx.x >> 3, cv = threadIdx.x & 7;  // 32 pixels x 8 vectors of 8 channels
  const int p = p0 + pl, c = c0 + cv * 8;
  if (p < HW && c < C) {
    uint16_t* o = dst + (static_cast<int64_t>(b) * HW + p) * ld + coff + c;
    if (c + 8 <= C && ((ld | coff) & 7) == 0) {
      uint4 v;
      v.x = pack_16x2(tile[cv * 8 + 0][pl], tile[cv * 8 + 1][pl], f16);
      v.y = pack_16x2(tile[cv * 8 + 2][pl], tile[cv * 8 + 3][pl], f16);
      v.z = pack_16x2(tile[cv * 8 + 4][pl], tile[cv * 8 + 5][pl], f16);
      v.w = pack_16x2(tile[cv * 8 + 6][pl], tile[cv * 8 + 7][pl], f16);
      *reinterpret_cast<uint4*>(o) = v;
    } else {
      for (int j = 0; j < 8 && c + j < C; ++j) o[j] = to_16(tile[cv * 8 + j][pl], f16);
    }
  }
}

// [B, HW, ld] bf16 (channels coff..coff+C) -> [B, C, HW] fp32
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const uint16_t* __restrict__ src, float* __restrict__ dst,
                                                           int C, int HW, int ld, int coff, int f16) {
  __shared__ float tile[kTC][kTP + 1];
  const int p0 = blockIdx.x * kTP, c0 = blockIdx.y * kTC, b = blockIdx.z;
  const int pl = threadIdx.x >> 3, cv = threadIdx.x & 7;
  const int p = p0 + pl, c = c0 + cv * 8;
  if (p < HW && c < C) {
    const uint16_t* s = src + (static_cast<int64_t>(b) * HW + p) * ld + coff + c;
    if (c + 8 <= C && ((ld | coff) & 7) == 0) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(s));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) unpack_16x2(w[q], f16, tile[cv * 8 + 2 * q][pl], tile[cv * 8 + 2 * q + 1][pl]);
    } else {
      for (int j = 0; j < 8; ++j) tile[cv * 8 + j][pl] = (c + j < C) ? from_16(s[j], f16) : 0.0f;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* d = dst + static_cast<int64_t>(b) * C * HW;
#pragma unroll
  for (int i = 0; i < kTC / 8; ++i) {
    const int cc = c0 + warp * (kTC / 8) + i;
    const int pp = p0 + lane;
    if (cc < C && pp < HW) d[static_cast<int64_t>(cc) * HW + pp] = tile[warp * (kTC / 8) + i][lane];
  }
}

// ---------------------------------------------------------------------------------------------- SE gate
// stage 1: per (image, slab of pixels) channel sums, fixed summation order (deterministic)
__global__ void __launch_bounds__(256) se_partial_kernel(const uint16_t* __restrict__ x, float* __restrict__ scratch,
                                                         int HW, int C, int ld, int f16) {
  pdl_prologue();
  extern __shared__ float red[];  // [rows_par][C]
  const int slab = blockIdx.x, b = blockIdx.y;
  const int per = (HW + GLSDET_SE_SLABS - 1) / GLSDET_SE_SLABS;
  const int p_begin = slab * per;
  const int p_end = min(HW, p_begin + per);
  const int nvec = C >> 3;
  const int lanes = min(nvec, 256);
  const int rows_par = 256 / lanes;
  const int row = threadIdx.x / lanes;
  const int vl = threadIdx.x % lanes;
  if (row < rows_par) {
    for (int vec = vl; vec < nvec; vec += lanes) {
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 4
      for (int p = p_begin + row; p < p_end; p += rows_par) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + (static_cast<int64_t>(b) * HW + p) * ld) + vec);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float lo, hi;
          unpack_16x2(w[q], f16, lo, hi);
          acc[2 * q] += lo;
          acc[2 * q + 1] += hi;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) red[row * C + vec * 8 + j] = acc[j];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float s = 0.0f;
    for (int r = 0; r < rows_par; ++r) s += red[r * C + c];
    scratch[(static_cast<int64_t>(b) * GLSDET_SE_SLABS + slab) * C + c] = s;
  }
}

// stage 2: mean -> FC -> ReLU -> FC -> sigmoid; gate = 1 + y  (x + x*y == x*(1+y), ffa.py:77)
__global__ void __launch_bounds__(256) se_fc_kernel(const float* __restrict__ scratch, const float* __restrict__ w1,
                                                    const float* __restrict__ w2, float* __restrict__ gate, int HW,
                                                    int C, int hidden) {
  pdl_prologue();
  extern __shared__ float sm[];  // mean[C] | hid[hidden]
  float* mean = sm;
  float* hid = sm + C;
  const int b = blockIdx.x;
  const float inv = 1.0f / static_cast<float>(HW);
  for (int c = threadIdx.x; c < C; c += 256) {
    float s = 0.0f;
    for (int k = 0; k < GLSDET_SE_SLABS; ++k) s += scratch[(static_cast<int64_t>(b) * GLSDET_SE_SLABS + k) * C + c];
    mean[c] = s * inv;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int h = warp; h < hidden; h += 8) {
    float s = 0.0f;
    for (int c = lane; c < C; c += 32) s += __ldg(w1 + static_cast<int64_t>(h) * C + c) * mean[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) hid[h] = fmaxf(s, 0.0f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float s = 0.0f;
    for (int h = 0; h < hidden; ++h) s += __ldg(w2 + static_cast<int64_t>(c) * hidden + h) * hid[h];
    gate[static_cast<int64_t>(b) * C + c] = 1.0f + 1.0f / (1.0f + expf(-s));
  }
}

// dst[b, 2y+i, 2x+j, coff + c] = x[b, y, x, (2i+j)*Cout + c] * gate[b, (2i+j)*Cout + c]
__global__ void __launch_bounds__(256) scale_shuffle_kernel(const uint16_t* __restrict__ x,
                                                            const float* __restrict__ gate,
                                                            uint16_t* __restrict__ dst, int B, int H, int W,
                                                            int Cout, int ld, int coff, int xf16, int df16) {
  pdl_prologue();
  const int vec_per_pix = (4 * Cout) >> 3;
  const int64_t total = static_cast<int64_t>(B) * H * W * vec_per_pix;
  const int cvn = Cout >> 3;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int vec = static_cast<int>(i % vec_per_pix);
    int64_t pix = i / vec_per_pix;
    const int xx = static_cast<int>(pix % W);
    pix /= W;
    const int yy = static_cast<int>(pix % H);
    const int b = static_cast<int>(pix / H);
    const int ij = vec / cvn, cv = vec % cvn;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x) + i);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gate + static_cast<int64_t>(b) * 4 * Cout + vec * 8));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gate + static_cast<int64_t>(b) * 4 * Cout + vec * 8) + 1);
    uint4 o;
    float a0, a1, a2, a3, a4, a5, a6, a7;
    unpack_16x2(v.x, xf16, a0, a1);
    unpack_16x2(v.y, xf16, a2, a3);
    unpack_16x2(v.z, xf16, a4, a5);
    unpack_16x2(v.w, xf16, a6, a7);
    o.x = pack_16x2(a0 * g0.x, a1 * g0.y, df16);
    o.y = pack_16x2(a2 * g0.z, a3 * g0.w, df16);
    o.z = pack_16x2(a4 * g1.x, a5 * g1.y, df16);
    o.w = pack_16x2(a6 * g1.z, a7 * g1.w, df16);
    const int oy = 2 * yy + (ij >> 1), ox = 2 * xx + (ij & 1);
    uint16_t* d = dst + ((static_cast<int64_t>(b) * 2 * H + oy) * (2 * W) + ox) * ld + coff + cv * 8;
    *reinterpret_cast<uint4*>(d) = o;
  }
}

// ---------------------------------------------------------------------------------------------- decode
struct DecodeLevels {
  const float* ptr[8];
  int h[8], w[8], a_off[8];
  int num;
};

// models/core/utils_bbox.py:254-306 and its variants (:36-251) - one thread per (image, anchor).
// mode bits (GLSDET_DECODE_*): sigmoid on the objectness / on the classes, normalise by the input size, corner form.
__global__ void __launch_bounds__(256) decode_kernel(DecodeLevels lv, int B, int A, int nch, float in_h, float in_w,
                                                     float* __restrict__ pred, int mode) {
  const int64_t total = static_cast<int64_t>(B) * A;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int a = static_cast<int>(i % A);
    const int b = static_cast<int>(i / A);
    int l = 0;
    while (l + 1 < lv.num && a >= lv.a_off[l + 1]) ++l;
    const int hw = lv.h[l] * lv.w[l];
    const int cell = a - lv.a_off[l];
    const int gy = cell / lv.w[l], gx = cell % lv.w[l];
    const float stride = in_h / static_cast<float>(lv.h[l]);  // utils_bbox.py:285 (H-stride on both axes)
    const float* s = lv.ptr[l] + static_cast<int64_t>(b) * nch * hw + cell;
    float* o = pred + i * nch;
    float cx = (__ldg(s) + static_cast<float>(gx)) * stride;
    float cy = (__ldg(s + hw) + static_cast<float>(gy)) * stride;
    float bw = expf(__ldg(s + 2 * static_cast<int64_t>(hw))) * stride;
    float bh = expf(__ldg(s + 3 * static_cast<int64_t>(hw))) * stride;
    if (mode & GLSDET_DECODE_NORMALISE) { cx = cx / in_w; cy = cy / in_h; bw = bw / in_w; bh = bh / in_h; }
    if (mode & GLSDET_DECODE_XYXY) {   // utils_bbox.py:84-89: corners = centre -/+ size / 2 (explicit roundings, no FMA)
      o[0] = __fsub_rn(cx, __fdiv_rn(bw, 2.0f)); o[1] = __fsub_rn(cy, __fdiv_rn(bh, 2.0f));
      o[2] = __fadd_rn(cx, __fdiv_rn(bw, 2.0f)); o[3] = __fadd_rn(cy, __fdiv_rn(bh, 2.0f));
    } else {
      o[0] = cx; o[1] = cy; o[2] = bw; o[3] = bh;
    }
    const float ob = __ldg(s + 4 * static_cast<int64_t>(hw));
    o[4] = (mode & GLSDET_DECODE_SIGMOID_OBJ) ? 1.0f / (1.0f + expf(-ob)) : ob;
    if (mode & GLSDET_DECODE_SIGMOID_CLS) {
      for (int c = 5; c < nch; ++c) o[c] = 1.0f / (1.0f + expf(-__ldg(s + c * static_cast<int64_t>(hw))));
    } else {
      for (int c = 5; c < nch; ++c) o[c] = __ldg(s + c * static_cast<int64_t>(hw));
    }
  }
}

struct MmdetLevels {
  const float* cls[8];
  const float* box[8];
  const float* obj[8];
  long long cls_bs[8], box_bs[8], obj_bs[8];
  int h[8], w[8], a_off[8];
  float stride[8];
  int num;
};

// yolox-ufp/mmdet/models/dense_heads/yolox_head.py:262-281,298-301 - one thread per (image, anchor)
__global__ void __launch_bounds__(256) decode_mmdet_kernel(MmdetLevels lv, int B, int A, int nc, float* __restrict__ pred) {
  const int64_t total = static_cast<int64_t>(B) * A;
  const int nch = 5 + nc;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int a = static_cast<int>(i % A);
    const int b = static_cast<int>(i / A);
    int l = 0;
    while (l + 1 < lv.num && a >= lv.a_off[l + 1]) ++l;
    const int64_t hw = static_cast<int64_t>(lv.h[l]) * lv.w[l];
    const int cell = a - lv.a_off[l];
    const int gy = cell / lv.w[l], gx = cell % lv.w[l];
    const float st = lv.stride[l];
    const float* bx = lv.box[l] + b * lv.box_bs[l] + cell;
    float* o = pred + i * nch;
    o[0] = __fadd_rn(__fmul_rn(__ldg(bx), st), static_cast<float>(gx) * st);
    o[1] = __fadd_rn(__fmul_rn(__ldg(bx + hw), st), static_cast<float>(gy) * st);
    o[2] = __fmul_rn(expf(__ldg(bx + 2 * hw)), st);
    o[3] = __fmul_rn(expf(__ldg(bx + 3 * hw)), st);
    o[4] = 1.0f / (1.0f + expf(-__ldg(lv.obj[l] + b * lv.obj_bs[l] + cell)));
    const float* cl = lv.cls[l] + b * lv.cls_bs[l] + cell;
    for (int c = 0; c < nc; ++c) o[5 + c] = 1.0f / (1.0f + expf(-__ldg(cl + c * hw)));
  }
}

}  // namespace glsdet

using namespace glsdet;

extern "C" int glsdet_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int32_t batch, int32_t channels,
                                            int32_t height, int32_t width, int32_t dst_ld, int32_t dst_coff,
                                            void* stream) {
  return glsdet_nchw_f32_to_nhwc_16(src, dst, batch, channels, height, width, dst_ld, dst_coff, GLSDET_DT_BF16, stream);
}

extern "C" int glsdet_nchw_f32_to_nhwc_16(const float* src, void* dst, int32_t batch, int32_t channels,
                                          int32_t height, int32_t width, int32_t dst_ld, int32_t dst_coff,
                                          int32_t dtype, void* stream) {
  GLSDET_REQUIRE(src && dst && batch > 0 && channels > 0 && height > 0 && width > 0, "nchw_to_nhwc: bad arguments");
  GLSDET_REQUIRE(dtype == GLSDET_DT_BF16 || dtype == GLSDET_DT_F16, "nchw_to_nhwc: bad storage dtype");
  GLSDET_REQUIRE(dst_coff >= 0 && dst_coff + channels <= dst_ld, "nchw_to_nhwc: channel window exceeds pitch");
  const int HW = height * width;
  dim3 grid((HW + kTP - 1) / kTP, (channels + kTC - 1) / kTC, batch);
  GLSDET_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "nchw_to_nhwc: grid too large");
  launch_pdl(nchw_to_nhwc_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), 
      src, reinterpret_cast<uint16_t*>(dst), channels, HW, dst_ld, dst_coff, static_cast<int>(dtype));
  return count_launch("nchw_to_nhwc_kernel");
}

extern "C" int glsdet_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int32_t batch, int32_t channels,
                                            int32_t height, int32_t width, int32_t src_ld, int32_t src_coff,
                                            void* stream) {
  return glsdet_nhwc_16_to_nchw_f32(src, dst, batch, channels, height, width, src_ld, src_coff, GLSDET_DT_BF16, stream);
}

extern "C" int glsdet_nhwc_16_to_nchw_f32(const void* src, float* dst, int32_t batch, int32_t channels,
                                          int32_t height, int32_t width, int32_t src_ld, int32_t src_coff,
                                          int32_t dtype, void* stream) {
  GLSDET_REQUIRE(src && dst && batch > 0 && channels > 0 && height > 0 && width > 0, "nhwc_to_nchw: bad arguments");
  GLSDET_REQUIRE(dtype == GLSDET_DT_BF16 || dtype == GLSDET_DT_F16, "nhwc_to_nchw: bad storage dtype");
  GLSDET_REQUIRE(src_coff >= 0 && src_coff + channels <= src_ld, "nhwc_to_nchw: channel window exceeds pitch");
  const int HW = height * width;
  dim3 grid((HW + kTP - 1) / kTP, (channels + kTC - 1) / kTC, batch);
  GLSDET_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "nhwc_to_nchw: grid too large");
  nhwc_to_nchw_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint16_t*>(src), dst, channels, HW, src_ld, src_coff, static_cast<int>(dtype));
  return count_launch("nhwc_to_nchw_kernel");
}

extern "C" int glsdet_se_gate(const void* x, int32_t batch, int32_t hw, int32_t channels, int32_t x_ld,
                              const float* w1, const float* w2, int32_t hidden, float* scratch, float* gate,
                              void* stream) {
  return glsdet_se_gate_16(x, batch, hw, channels, x_ld, w1, w2, hidden, scratch, gate, GLSDET_DT_BF16, stream);
}

extern "C" int glsdet_se_gate_16(const void* x, int32_t batch, int32_t hw, int32_t channels, int32_t x_ld,
                                 const float* w1, const float* w2, int32_t hidden, float* scratch, float* gate,
                                 int32_t x_dtype, void* stream) {
  GLSDET_REQUIRE(x && w1 && w2 && scratch && gate, "se_gate: null pointer");
  GLSDET_REQUIRE(x_dtype == GLSDET_DT_BF16 || x_dtype == GLSDET_DT_F16, "se_gate: bad storage dtype");
  GLSDET_REQUIRE(batch > 0 && hw > 0 && channels > 0 && hidden > 0, "se_gate: bad sizes");
  GLSDET_REQUIRE((channels % 8) == 0 && (x_ld % 8) == 0 && x_ld >= channels, "se_gate: channels/pitch must be multiples of 8");
  const int nvec = channels / 8;
  const int lanes = nvec < 256 ? nvec : 256;
  const int rows_par = 256 / lanes;
  const size_t smem1 = static_cast<size_t>(rows_par) * channels * sizeof(float);
  GLSDET_REQUIRE(smem1 <= 48 * 1024, "se_gate: too many channels (%d)", channels);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  launch_pdl(se_partial_kernel, dim3(GLSDET_SE_SLABS, batch), 256, smem1, st, reinterpret_cast<const uint16_t*>(x), scratch,
             hw, channels, x_ld, static_cast<int>(x_dtype));
  if (int rc = count_launch("se_partial_kernel")) return rc;
  const size_t smem2 = static_cast<size_t>(channels + hidden) * sizeof(float);
  launch_pdl(se_fc_kernel, batch, 256, smem2, st, scratch, w1, w2, gate, hw, channels, hidden);
  return count_launch("se_fc_kernel");
}

extern "C" int glsdet_se_fc(const float* scratch, const float* w1, const float* w2, int32_t hidden, float* gate,
                            int32_t batch, int32_t hw, int32_t channels, void* stream) {
  GLSDET_REQUIRE(scratch && w1 && w2 && gate && batch > 0 && hw > 0 && channels > 0 && hidden > 0, "se_fc: bad arguments");
  const size_t smem2 = static_cast<size_t>(channels + hidden) * sizeof(float);
  GLSDET_REQUIRE(smem2 <= 48 * 1024, "se_fc: too many channels (%d)", channels);
  launch_pdl(se_fc_kernel, batch, 256, smem2, static_cast<cudaStream_t>(stream), scratch, w1, w2, gate, hw, channels, hidden);
  return count_launch("se_fc_kernel");
}

extern "C" int glsdet_scale_pixel_shuffle(const void* x, const float* gate, void* dst, int32_t batch, int32_t height,
                                          int32_t width, int32_t out_channels, int32_t dst_ld, int32_t dst_coff,
                                          void* stream) {
  return glsdet_scale_pixel_shuffle_16(x, gate, dst, batch, height, width, out_channels, dst_ld, dst_coff, GLSDET_DT_BF16,
                                       GLSDET_DT_BF16, stream);
}

extern "C" int glsdet_scale_pixel_shuffle_16(const void* x, const float* gate, void* dst, int32_t batch, int32_t height,
                                             int32_t width, int32_t out_channels, int32_t dst_ld, int32_t dst_coff,
                                             int32_t x_dtype, int32_t dst_dtype, void* stream) {
  GLSDET_REQUIRE(x && gate && dst, "scale_pixel_shuffle: null pointer");
  GLSDET_REQUIRE((x_dtype == GLSDET_DT_BF16 || x_dtype == GLSDET_DT_F16) && (dst_dtype == GLSDET_DT_BF16 || dst_dtype == GLSDET_DT_F16),
                 "scale_pixel_shuffle: bad storage dtype");
  GLSDET_REQUIRE(batch > 0 && height > 0 && width > 0 && out_channels > 0, "scale_pixel_shuffle: bad sizes");
  GLSDET_REQUIRE((out_channels % 8) == 0 && (dst_ld % 8) == 0 && (dst_coff % 8) == 0 && dst_coff + out_channels <= dst_ld,
                 "scale_pixel_shuffle: channels/pitch/offset must be multiples of 8");
  const int64_t total = static_cast<int64_t>(batch) * height * width * (4 * out_channels / 8);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 32;
  if (blocks > cap) blocks = cap;
  launch_pdl(scale_shuffle_kernel, static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream), 
      reinterpret_cast<const uint16_t*>(x), gate, reinterpret_cast<uint16_t*>(dst), batch, height, width,
      out_channels, dst_ld, dst_coff, static_cast<int>(x_dtype), static_cast<int>(dst_dtype));
  return count_launch("scale_shuffle_kernel");
}

extern "C" int glsdet_decode_outputs(const float* const* levels, const int32_t* heights, const int32_t* widths,
                                     int32_t num_levels, int32_t batch, int32_t num_classes, int32_t in_h,
                                     int32_t in_w, float* pred, void* stream) {
  return glsdet_decode_outputs_mode(levels, heights, widths, num_levels, batch, num_classes, in_h, in_w,
                                    GLSDET_DECODE_SIGMOID_OBJ | GLSDET_DECODE_SIGMOID_CLS | GLSDET_DECODE_NORMALISE, pred, stream);
}

extern "C" int glsdet_decode_outputs_mode(const float* const* levels, const int32_t* heights, const int32_t* widths,
                                          int32_t num_levels, int32_t batch, int32_t num_classes, int32_t in_h,
                                          int32_t in_w, int32_t mode, float* pred, void* stream) {
  GLSDET_REQUIRE(levels && heights && widths && pred, "decode_outputs: null pointer");
  GLSDET_REQUIRE(mode >= 0 && mode < 16, "decode_outputs: bad mode %d", mode);
  GLSDET_REQUIRE(num_levels > 0 && num_levels <= 8, "decode_outputs: 1..8 levels supported (got %d)", num_levels);
  GLSDET_REQUIRE(batch > 0 && num_classes > 0 && in_h > 0 && in_w > 0, "decode_outputs: bad sizes");
  DecodeLevels lv;
  int a = 0;
  for (int l = 0; l < num_levels; ++l) {
    GLSDET_REQUIRE(levels[l] && heights[l] > 0 && widths[l] > 0, "decode_outputs: bad level %d", l);
    lv.ptr[l] = levels[l]; lv.h[l] = heights[l]; lv.w[l] = widths[l]; lv.a_off[l] = a;
    a += heights[l] * widths[l];
  }
  lv.num = num_levels;
  const int64_t total = static_cast<int64_t>(batch) * a;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 32;
  if (blocks > cap) blocks = cap;
  decode_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      lv, batch, a, 5 + num_classes, static_cast<float>(in_h), static_cast<float>(in_w), pred, static_cast<int>(mode));
  return count_launch("decode_kernel");
}

extern "C" int glsdet_decode_mmdet(const float* const* cls, const float* const* box, const float* const* obj,
                                   const int64_t* cls_bs, const int64_t* box_bs, const int64_t* obj_bs,
                                   const int32_t* heights, const int32_t* widths, const int32_t* strides,
                                   int32_t num_levels, int32_t batch, int32_t num_classes, float* pred, void* stream) {
  GLSDET_REQUIRE(cls && box && obj && cls_bs && box_bs && obj_bs && heights && widths && strides && pred,
                 "decode_mmdet: null pointer");
  GLSDET_REQUIRE(num_levels > 0 && num_levels <= 8, "decode_mmdet: 1..8 levels supported (got %d)", num_levels);
  GLSDET_REQUIRE(batch > 0 && num_classes > 0, "decode_mmdet: bad sizes");
  MmdetLevels lv;
  int a = 0;
  for (int l = 0; l < num_levels; ++l) {
    GLSDET_REQUIRE(cls[l] && box[l] && obj[l] && heights[l] > 0 && widths[l] > 0 && strides[l] > 0,
                   "decode_mmdet: bad level %d", l);
    lv.cls[l] = cls[l]; lv.box[l] = box[l]; lv.obj[l] = obj[l];
    lv.cls_bs[l] = cls_bs[l]; lv.box_bs[l] = box_bs[l]; lv.obj_bs[l] = obj_bs[l];
    lv.h[l] = heights[l]; lv.w[l] = widths[l]; lv.a_off[l] = a; lv.stride[l] = static_cast<float>(strides[l]);
    a += heights[l] * widths[l];
  }
  lv.num = num_levels;
  const int64_t total = static_cast<int64_t>(batch) * a;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 32;
  if (blocks > cap) blocks = cap;
  decode_mmdet_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(lv, batch, a,
                                                                                                  num_classes, pred);
  return count_launch("decode_mmdet_kernel");
}

// ---------------------------------------------------------------------------------------------- non-local helpers
namespace glsdet {

// [B, C, H, W] fp32 -> per-patch transposed bf16 [B*4, rows, t_ld]: row c of patch image b' = (b*2 + py)*2 + px holds
// the pixels of patch (py, px) in row-major order (t = y' * w2 + x').  This is the K-major operand of the Gram
// product X^T X (the contraction runs over pixels).  Rows >= C (ones row, zero padding) and columns >= h2*w2 are
// written once at plan build time and never touched here.  PAIR: one thread per two pixels of a row (needs an even
// patch width so that a pair never straddles two patches), else one thread per pixel.
template <bool PAIR>
__global__ void __launch_bounds__(256) patch_transpose_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst,
                                                              int C, int H, int W, int rows, int t_ld, int f16,
                                                              float scale) {
  const int h2 = H >> 1, w2 = W >> 1;
  const int b = blockIdx.y;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int per_row = PAIR ? (W >> 1) : W;
  if (idx >= static_cast<int64_t>(C) * H * per_row) return;
  const int xp = static_cast<int>(idx % per_row);
  const int y = static_cast<int>((idx / per_row) % H);
  const int c = static_cast<int>(idx / (static_cast<int64_t>(per_row) * H));
  const int x = PAIR ? xp * 2 : xp;
  const float* s = src + ((static_cast<int64_t>(b) * C + c) * H + y) * W + x;
  const int py = y >= h2, px = x >= w2;
  const int bp = (b * 2 + py) * 2 + px;
  const int t = (y - py * h2) * w2 + (x - px * w2);
  uint16_t* o = dst + (static_cast<int64_t>(bp) * rows + c) * t_ld + t;
  if (PAIR) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(s));
    *reinterpret_cast<uint32_t*>(o) = pack_16x2(v.x * scale, v.y * scale, f16);
  } else {
    *o = to_16(__ldg(s) * scale, f16);
  }
}

// bias[b'][n] = base[n] + w[b'][n][col]  (the bias column of the per-patch effective weight matrix)
__global__ void __launch_bounds__(256) gather_bias_kernel(const uint16_t* __restrict__ w, const float* __restrict__ base,
                                                          float* __restrict__ bias, int n_rows, int ld, int col,
                                                          int64_t batch_stride, int base_groups, int total, int f16) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int b = i / n_rows, n = i - b * n_rows;
  bias[i] = base[(b % base_groups) * n_rows + n] + from_16(w[static_cast<int64_t>(b) * batch_stride + static_cast<int64_t>(n) * ld + col], f16);
}

}  // namespace glsdet

extern "C" int glsdet_patch_transpose(const float* src, void* dst, int32_t batch, int32_t channels, int32_t height,
                                      int32_t width, int32_t dst_rows, int32_t dst_ld, void* stream) {
  return glsdet_patch_transpose_16(src, dst, batch, channels, height, width, dst_rows, dst_ld, GLSDET_DT_BF16, 1.0f, stream);
}

extern "C" int glsdet_patch_transpose_16(const float* src, void* dst, int32_t batch, int32_t channels, int32_t height,
                                         int32_t width, int32_t dst_rows, int32_t dst_ld, int32_t dtype, float scale,
                                         void* stream) {
  GLSDET_REQUIRE(src && dst && batch > 0 && channels > 0, "patch_transpose: bad arguments");
  GLSDET_REQUIRE(dtype == GLSDET_DT_BF16 || dtype == GLSDET_DT_F16, "patch_transpose: bad storage dtype");
  GLSDET_REQUIRE((height % 2) == 0 && (width % 2) == 0, "patch_transpose: height and width must be even "
                 "(equal 2x2 patches; odd splits of Non_local_family.py:230-233 are not supported)");
  GLSDET_REQUIRE(dst_rows >= channels && dst_ld >= (height / 2) * (width / 2) && (dst_ld % 2) == 0,
                 "patch_transpose: destination too small");
  const bool pair = (width % 4) == 0;
  const int64_t n = static_cast<int64_t>(channels) * height * (pair ? width / 2 : width);
  dim3 grid(static_cast<unsigned>((n + 255) / 256), batch);
  if (pair)
    glsdet::patch_transpose_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, reinterpret_cast<uint16_t*>(dst), channels, height, width, dst_rows, dst_ld, static_cast<int>(dtype), scale);
  else
    glsdet::patch_transpose_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, reinterpret_cast<uint16_t*>(dst), channels, height, width, dst_rows, dst_ld, static_cast<int>(dtype), scale);
  return glsdet::count_launch("patch_transpose_kernel");
}

extern "C" int glsdet_gather_bias(const void* w, const float* base, float* bias, int32_t batch, int32_t n_rows,
                                  int32_t ld, int32_t col, int64_t batch_stride, int32_t base_groups, void* stream) {
  return glsdet_gather_bias_16(w, base, bias, batch, n_rows, ld, col, batch_stride, base_groups, GLSDET_DT_BF16, stream);
}

extern "C" int glsdet_gather_bias_16(const void* w, const float* base, float* bias, int32_t batch, int32_t n_rows,
                                     int32_t ld, int32_t col, int64_t batch_stride, int32_t base_groups, int32_t dtype,
                                     void* stream) {
  GLSDET_REQUIRE(dtype == GLSDET_DT_BF16 || dtype == GLSDET_DT_F16, "gather_bias: bad storage dtype");
  GLSDET_REQUIRE(w && base && bias && batch > 0 && n_rows > 0 && col >= 0 && col < ld && base_groups > 0,
                 "gather_bias: bad arguments");
  const int total = batch * n_rows;
  glsdet::gather_bias_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint16_t*>(w), base, bias, n_rows, ld, col, batch_stride, base_groups, total,
      static_cast<int>(dtype));
  return glsdet::count_launch("gather_bias_kernel");
}

// ---------------------------------------------------------------------------------------------- nearest upsample
namespace glsdet {

// dst[b, 2y+i, 2x+j, dcoff + c] = src[b, y, x, scoff + c]  (nn.Upsample(scale_factor=2, mode="nearest") into a
// channel window of a concat buffer).  One thread per (source pixel, 8-channel vector): one 16-byte load, four stores.
__global__ void __launch_bounds__(256) upsample2x_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                         int H, int W, int C, int sld, int scoff, int dld, int dcoff,
                                                         int64_t total) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int nvec = C >> 3;
  const int v = static_cast<int>(idx % nvec);
  int64_t pix = idx / nvec;
  const int x = static_cast<int>(pix % W);
  pix /= W;
  const int y = static_cast<int>(pix % H);
  const int b = static_cast<int>(pix / H);
  const uint4 val = __ldg(reinterpret_cast<const uint4*>(src + ((static_cast<int64_t>(b) * H + y) * W + x) * sld + scoff) + v);
  __nv_bfloat16* o = dst + ((static_cast<int64_t>(b) * 2 * H + 2 * y) * (2 * W) + 2 * x) * dld + dcoff + v * 8;
  const int64_t row = static_cast<int64_t>(2) * W * dld;
  *reinterpret_cast<uint4*>(o) = val;
  *reinterpret_cast<uint4*>(o + dld) = val;
  *reinterpret_cast<uint4*>(o + row) = val;
  *reinterpret_cast<uint4*>(o + row + dld) = val;
}

}  // namespace glsdet

extern "C" int glsdet_upsample2x(const void* src, void* dst, int32_t batch, int32_t height, int32_t width,
                                 int32_t channels, int32_t src_ld, int32_t src_coff, int32_t dst_ld, int32_t dst_coff,
                                 void* stream) {
  GLSDET_REQUIRE(src && dst && batch > 0 && height > 0 && width > 0 && channels > 0, "upsample2x: bad arguments");
  GLSDET_REQUIRE((channels % 8) == 0 && (src_ld % 8) == 0 && (src_coff % 8) == 0 && (dst_ld % 8) == 0 && (dst_coff % 8) == 0,
                 "upsample2x: channels, pitches and offsets must be multiples of 8");
  GLSDET_REQUIRE(src_coff + channels <= src_ld && dst_coff + channels <= dst_ld, "upsample2x: window exceeds pitch");
  const int64_t total = static_cast<int64_t>(batch) * height * width * (channels / 8);
  glsdet::upsample2x_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), reinterpret_cast<__nv_bfloat16*>(dst), height, width, channels, src_ld,
      src_coff, dst_ld, dst_coff, total);
  return glsdet::count_launch("upsample2x_kernel");
}

// ---------------------------------------------------------------------------------------------- rectangle copies
namespace glsdet {

struct RectList {
  glsdet_rect r[8];
};

// one thread per (rectangle pixel, 8-channel vector); blockIdx.y = rectangle
__global__ void __launch_bounds__(256) rect_copy_kernel(const __nv_bfloat16* __restrict__ src, int sh, int sw, int sld, int scoff,
                                                        __nv_bfloat16* __restrict__ dst, int dh, int dw, int dld, int dcoff,
                                                        int B, int C, RectList rl) {
  const glsdet_rect q = rl.r[blockIdx.y];
  const int nvec = C >> 3;
  const int64_t total = static_cast<int64_t>(B) * q.h * q.w * nvec;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % nvec);
    int64_t pix = i / nvec;
    const int x = static_cast<int>(pix % q.w);
    pix /= q.w;
    const int y = static_cast<int>(pix % q.h);
    const int b = static_cast<int>(pix / q.h);
    const uint4 val = __ldg(reinterpret_cast<const uint4*>(
        src + ((static_cast<int64_t>(q.sb + b) * sh + q.sy + y) * sw + q.sx + x) * sld + scoff) + v);
    *(reinterpret_cast<uint4*>(dst + ((static_cast<int64_t>(q.db + b) * dh + q.dy + y) * dw + q.dx + x) * dld + dcoff) + v) = val;
  }
}

// [B, T, ld] bf16 (channels coff..coff+C) -> [B, rows, t_ld] bf16 with dst[b][c][t] = src[b][t][c]; 32 x 32 smem tiles
// `scaled`: elements are converted, multiplied by `scale` and rounded again (f16: dtype flag); else a plain 16-bit copy
__global__ void __launch_bounds__(256) nhwc_transpose_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst,
                                                             int T, int C, int sld, int scoff, int rows, int t_ld,
                                                             int scaled, int f16, float scale) {
  __shared__ uint16_t tile[32][34];
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32, b = blockIdx.z;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  for (int i = ly; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + lx;
    uint16_t v = (t < T && c < C) ? src[(static_cast<int64_t>(b) * T + t) * sld + scoff + c] : static_cast<uint16_t>(0);
    if (scaled) v = to_16(from_16(v, f16) * scale, f16);
    tile[i][lx] = v;
  }
  __syncthreads();
  for (int i = ly; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + lx;
    if (c < C && t < T) dst[(static_cast<int64_t>(b) * rows + c) * t_ld + t] = tile[lx][i];
  }
}

}  // namespace glsdet

extern "C" int glsdet_rect_copy(const void* src, int32_t src_h, int32_t src_w, int32_t src_ld, int32_t src_coff, void* dst,
                                int32_t dst_h, int32_t dst_w, int32_t dst_ld, int32_t dst_coff, int32_t batch,
                                int32_t channels, const glsdet_rect* rects, int32_t num_rects, void* stream) {
  GLSDET_REQUIRE(src && dst && rects && batch > 0 && channels > 0 && num_rects >= 1 && num_rects <= 8,
                 "rect_copy: bad arguments (1..8 rectangles)");
  GLSDET_REQUIRE(((channels | src_ld | src_coff | dst_ld | dst_coff) % 8) == 0, "rect_copy: channels, pitches and offsets "
                 "must be multiples of 8");
  GLSDET_REQUIRE(src_coff + channels <= src_ld && dst_coff + channels <= dst_ld, "rect_copy: window exceeds pitch");
  glsdet::RectList rl;
  int64_t biggest = 0;
  for (int i = 0; i < num_rects; ++i) {
    const glsdet_rect& q = rects[i];
    GLSDET_REQUIRE(q.h > 0 && q.w > 0 && q.sy >= 0 && q.sx >= 0 && q.dy >= 0 && q.dx >= 0 && q.sb >= 0 && q.db >= 0 &&
                   q.sy + q.h <= src_h && q.sx + q.w <= src_w && q.dy + q.h <= dst_h && q.dx + q.w <= dst_w,
                   "rect_copy: rectangle %d leaves its tensor", i);
    rl.r[i] = q;
    const int64_t n = static_cast<int64_t>(batch) * q.h * q.w * (channels / 8);
    if (n > biggest) biggest = n;
  }
  int64_t gx = (biggest + 255) / 256;
  if (gx > 4096) gx = 4096;
  glsdet::rect_copy_kernel<<<dim3(static_cast<unsigned>(gx), num_rects), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), src_h, src_w, src_ld, src_coff, reinterpret_cast<__nv_bfloat16*>(dst), dst_h,
      dst_w, dst_ld, dst_coff, batch, channels, rl);
  return glsdet::count_launch("rect_copy_kernel");
}

extern "C" int glsdet_nhwc_transpose(const void* src, void* dst, int32_t batch, int32_t pixels, int32_t channels,
                                     int32_t src_ld, int32_t src_coff, int32_t dst_rows, int32_t dst_ld, void* stream) {
  return glsdet_nhwc_transpose_16(src, dst, batch, pixels, channels, src_ld, src_coff, dst_rows, dst_ld, GLSDET_DT_BF16, 1.0f,
                                  stream);
}

extern "C" int glsdet_nhwc_transpose_16(const void* src, void* dst, int32_t batch, int32_t pixels, int32_t channels,
                                        int32_t src_ld, int32_t src_coff, int32_t dst_rows, int32_t dst_ld, int32_t dtype,
                                        float scale, void* stream) {
  GLSDET_REQUIRE(dtype == GLSDET_DT_BF16 || dtype == GLSDET_DT_F16, "nhwc_transpose: bad storage dtype");
  GLSDET_REQUIRE(src && dst && batch > 0 && pixels > 0 && channels > 0, "nhwc_transpose: bad arguments");
  GLSDET_REQUIRE(src_coff + channels <= src_ld && dst_rows >= channels && dst_ld >= pixels, "nhwc_transpose: sizes");
  dim3 grid((pixels + 31) / 32, (channels + 31) / 32, batch);
  GLSDET_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "nhwc_transpose: grid too large");
  glsdet::nhwc_transpose_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint16_t*>(src), reinterpret_cast<uint16_t*>(dst), pixels, channels, src_ld, src_coff,
      dst_rows, dst_ld, scale != 1.0f ? 1 : 0, static_cast<int>(dtype), scale);
  return glsdet::count_launch("nhwc_transpose_kernel");
}

// ---------------------------------------------------------------------------------------------------------------
// Payload of the detection gather (glsdet_b200/dist.py::DetectionGather; replaces the pickle -> uint8 tensor step of
// mmdet's collect_results_gpu, yolox-ufp/mmdet/apis/test.py:161-175): ONE launch instead of ~10 framework kernels.
//   padded (total_rows == 0): out[b][0][0] = min(count[b], rows), out[b][1 + r] = det[b][r] for the rows that exist;
//   packed (total_rows  > 0): out = [hdr rows holding the B counts | every image's rows back to back | ...]; rows that
//                             would land beyond hdr + total_rows are dropped.
namespace glsdet {
__global__ void __launch_bounds__(256) pack_detections_kernel(const float* __restrict__ det, const int32_t* __restrict__ count,
                                                              int det_rows, int rows, int hdr, int total_rows,
                                                              float* __restrict__ out) {
  pdl_prologue();
  const int b = blockIdx.y;
  const int n = min(count[b], rows);
  const float* src = det + static_cast<int64_t>(b) * det_rows * 7;
  if (total_rows == 0) {
    float* dst = out + static_cast<int64_t>(b) * (1 + rows) * 7;
    if (blockIdx.x == 0 && threadIdx.x == 0) dst[0] = static_cast<float>(n);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * 7; i += gridDim.x * blockDim.x) dst[7 + i] = src[i];
    return;
  }
  int off = 0;
  for (int i = 0; i < b; ++i) off += min(count[i], rows);     // B is a few dozen: a serial prefix per CTA is cheapest
  if (blockIdx.x == 0 && threadIdx.x == 0) out[b] = static_cast<float>(n);
  const int room = max(0, min(n, total_rows - off));
  float* dst = out + static_cast<int64_t>(hdr + off) * 7;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < room * 7; i += gridDim.x * blockDim.x) dst[i] = src[i];
}
}  // namespace glsdet

extern "C" int glsdet_pack_detections(const float* det, const int32_t* count, int32_t batch, int32_t det_rows, int32_t rows,
                                      int32_t total_rows, float* out, void* stream) {
  GLSDET_REQUIRE(det && count && out && batch > 0 && det_rows > 0 && rows > 0 && rows <= det_rows && total_rows >= 0,
                 "pack_detections: bad arguments");
  const int hdr = (batch + 6) / 7;
  const int gx = (rows * 7 + 256 * 8 - 1) / (256 * 8);
  glsdet::launch_pdl(glsdet::pack_detections_kernel, dim3(gx < 1 ? 1 : (gx > 64 ? 64 : gx), batch), dim3(256), 0,
                     static_cast<cudaStream_t>(stream), det, count, det_rows, rows, hdr, total_rows, out);
  return glsdet::count_launch("pack_detections_kernel");
}
