// Depthwise k x k convolution + folded BatchNorm + activation on NHWC channel windows: the `dconv` half of the
// reference's DWConv (yolox-drone/models/base/baseConv.py:22-30; phi = 'nano' builds every k > 1 conv of the backbone's
// stages, the neck's bu_convs, the Bottleneck 3x3 convs and the head towers this way, e.g. models/ffa/yolox_ffa.py:15,125,
// models/ffa/darknet.py:48,120).  The `pconv` half is a 1x1 BaseConv and runs on the tcgen05 kernel (conv_gemm.cu).
//
// groups = channels: every output element is a k*k-tap dot product of ONE channel - 2 k^2 FLOP per 2 + 2 bytes of
// 16-bit traffic, i.e. HBM-bound by two orders of magnitude; there is no contraction for the tensor core.  One thread
// owns a 16-byte channel vector (8 x 16-bit or 4 x fp32) of one output pixel: consecutive lanes = consecutive channel
// vectors, then consecutive pixels, so every tap is a coalesced run; the k^2-fold re-reads of a pixel hit L1 / L2 (a tile
// of neighbouring pixels is touched by the same CTA within a few hundred cycles).  fp32 accumulation, weights fp32
// [tap][C] (BatchNorm folded in by the caller), persistent grid of SM-count multiples.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/glsdet_b200.h"
#include "common.h"
#include "ptx.cuh"

namespace glsdet {

struct DwParams {
  const void* src;
  void* dst;
  const float* w;      // [k * k][C]
  const float* bias;   // [C]
  int B, H, W, C, Ho, Wo;
  int sld, scoff, dld, dcoff;
  int k, stride, act;
  int f16;             // 16-bit flavour (VEC == 8 only)
  int64_t total;       // B * Ho * Wo * (C / VEC)
};

__device__ __forceinline__ float dw_act(float v, int act) {
  switch (act) {
    case GLSDET_ACT_SILU: return v / (1.0f + expf(-v));   // exact: the result is stored in a 16-bit (or fp32) tensor
    case GLSDET_ACT_RELU: return fmaxf(v, 0.0f);
    case GLSDET_ACT_LRELU: return v > 0.0f ? v : 0.1f * v;
    default: return v;
  }
}

// VEC = 8: 16-bit storage (bf16 / fp16 by p.f16); VEC = 4: fp32 (accuracy mode)
template <int VEC>
__global__ void __launch_bounds__(256) dwconv_kernel(const DwParams p) {
  pdl_prologue();
  const int nvec = p.C / VEC;
  const int pad = (p.k - 1) >> 1;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % nvec);
    int64_t pix = i / nvec;
    const int ox = static_cast<int>(pix % p.Wo);
    pix /= p.Wo;
    const int oy = static_cast<int>(pix % p.Ho);
    const int b = static_cast<int>(pix / p.Ho);
    const int c0 = v * VEC;
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; j += 4) {
      const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + j));
      acc[j] = bv.x; acc[j + 1] = bv.y; acc[j + 2] = bv.z; acc[j + 3] = bv.w;
    }
    for (int ky = 0; ky < p.k; ++ky) {
      const int iy = oy * p.stride + ky - pad;
      if (iy < 0 || iy >= p.H) continue;
      for (int kx = 0; kx < p.k; ++kx) {
        const int ix = ox * p.stride + kx - pad;
        if (ix < 0 || ix >= p.W) continue;
        const int64_t off = ((static_cast<int64_t>(b) * p.H + iy) * p.W + ix) * p.sld + p.scoff + c0;
        const float* wt = p.w + static_cast<int64_t>(ky * p.k + kx) * p.C + c0;
        float x[VEC];
        if constexpr (VEC == 8) {
          const uint4 t = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.src) + off));
          unpack_16x2(t.x, p.f16, x[0], x[1]);
          unpack_16x2(t.y, p.f16, x[2], x[3]);
          unpack_16x2(t.z, p.f16, x[4], x[5]);
          unpack_16x2(t.w, p.f16, x[6], x[7]);
        } else {
          const float4 t = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.src) + off));
          x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < VEC; j += 4) {
          const float4 wv = __ldg(reinterpret_cast<const float4*>(wt + j));
          acc[j] = fmaf(x[j], wv.x, acc[j]);
          acc[j + 1] = fmaf(x[j + 1], wv.y, acc[j + 1]);
          acc[j + 2] = fmaf(x[j + 2], wv.z, acc[j + 2]);
          acc[j + 3] = fmaf(x[j + 3], wv.w, acc[j + 3]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = dw_act(acc[j], p.act);
    const int64_t ooff = ((static_cast<int64_t>(b) * p.Ho + oy) * p.Wo + ox) * p.dld + p.dcoff + c0;
    if constexpr (VEC == 8) {
      uint4 o;
      o.x = pack_16x2(acc[0], acc[1], p.f16);
      o.y = pack_16x2(acc[2], acc[3], p.f16);
      o.z = pack_16x2(acc[4], acc[5], p.f16);
      o.w = pack_16x2(acc[6], acc[7], p.f16);
      *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.dst) + ooff) = o;
    } else {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.dst) + ooff) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
  }
}

// 3x3 specialisation (every DWConv of the reference is 3x3): the generic kernel above measured 0.07 - 0.30 of the HBM peak -
// nine dependent load -> FMA rounds per thread, 48 bytes through L1 per tap (16 of activations, 32 of weights) and 64-bit
// index arithmetic per tap.  Here a thread owns TX horizontally adjacent output pixels of one channel vector: per kernel row
// it issues all (TX - 1) * S + 3 column loads at once (independent, so their latencies overlap), every loaded pixel feeds up
// to three outputs from registers (S = 1: 4.5 loads per output instead of 9), and the folded weights + bias sit in shared
// memory (read once per TX outputs, bank-conflict free: lanes with the same channel vector broadcast).
template <int VEC, int S, int TX>
__global__ void __launch_bounds__(256) dw3_kernel(const DwParams p) {
  extern __shared__ float s_w[];   // [9][C] weights, then [C] bias
  pdl_prologue();
  for (int i = threadIdx.x; i < 10 * p.C; i += blockDim.x) s_w[i] = (i < 9 * p.C) ? __ldg(p.w + i) : __ldg(p.bias + (i - 9 * p.C));
  __syncthreads();
  constexpr int NC = (TX - 1) * S + 3;   // input columns under TX outputs
  const int nvec = p.C / VEC;
  const int wt = (p.Wo + TX - 1) / TX;
  const int64_t total = static_cast<int64_t>(p.B) * p.Ho * wt * nvec;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % nvec);
    int64_t r = i / nvec;
    const int xt = static_cast<int>(r % wt);
    r /= wt;
    const int oy = static_cast<int>(r % p.Ho);
    const int b = static_cast<int>(r / p.Ho);
    const int c0 = v * VEC;
    const int ox0 = xt * TX;
    const int ix0 = ox0 * S - 1;
    float acc[TX][VEC];
#pragma unroll
    for (int t = 0; t < TX; ++t)
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[t][j] = s_w[9 * p.C + c0 + j];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * S + ky - 1;
      const bool rowok = (iy >= 0) && (iy < p.H);
      const int64_t rowoff = (static_cast<int64_t>(b) * p.H + iy) * p.W * p.sld + p.scoff + c0;
      uint4 raw[NC];
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const int ix = ix0 + j;
        raw[j] = make_uint4(0u, 0u, 0u, 0u);
        if (rowok && ix >= 0 && ix < p.W) {
          const int64_t off = rowoff + static_cast<int64_t>(ix) * p.sld;
          if constexpr (VEC == 8) raw[j] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.src) + off));
          else raw[j] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.src) + off));
        }
      }
      float wk[3][VEC];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int j = 0; j < VEC; j += 4) {
          const float4 t4 = *reinterpret_cast<const float4*>(s_w + (ky * 3 + kx) * p.C + c0 + j);
          wk[kx][j] = t4.x; wk[kx][j + 1] = t4.y; wk[kx][j + 2] = t4.z; wk[kx][j + 3] = t4.w;
        }
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        float x[VEC];
        if constexpr (VEC == 8) {
          unpack_16x2(raw[j].x, p.f16, x[0], x[1]);
          unpack_16x2(raw[j].y, p.f16, x[2], x[3]);
          unpack_16x2(raw[j].z, p.f16, x[4], x[5]);
          unpack_16x2(raw[j].w, p.f16, x[6], x[7]);
        } else {
          x[0] = __uint_as_float(raw[j].x); x[1] = __uint_as_float(raw[j].y);
          x[2] = __uint_as_float(raw[j].z); x[3] = __uint_as_float(raw[j].w);
        }
#pragma unroll
        for (int t = 0; t < TX; ++t) {
          const int kx = j - t * S;   // compile-time after unrolling
          if (kx >= 0 && kx < 3) {
#pragma unroll
            for (int q = 0; q < VEC; ++q) acc[t][q] = fmaf(x[q], wk[kx][q], acc[t][q]);
          }
        }
      }
    }
    const int64_t orow = (static_cast<int64_t>(b) * p.Ho + oy) * p.Wo;
#pragma unroll
    for (int t = 0; t < TX; ++t) {
      const int ox = ox0 + t;
      if (ox >= p.Wo) break;
#pragma unroll
      for (int q = 0; q < VEC; ++q) acc[t][q] = dw_act(acc[t][q], p.act);
      const int64_t ooff = (orow + ox) * p.dld + p.dcoff + c0;
      if constexpr (VEC == 8) {
        uint4 o;
        o.x = pack_16x2(acc[t][0], acc[t][1], p.f16);
        o.y = pack_16x2(acc[t][2], acc[t][3], p.f16);
        o.z = pack_16x2(acc[t][4], acc[t][5], p.f16);
        o.w = pack_16x2(acc[t][6], acc[t][7], p.f16);
        *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.dst) + ooff) = o;
      } else {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.dst) + ooff) =
            make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
      }
    }
  }
}

template <int VEC, int S, int TX>
static void launch_dw3(const DwParams& p, cudaStream_t st) {
  const int wt = (p.Wo + TX - 1) / TX;
  const int64_t total = static_cast<int64_t>(p.B) * p.Ho * wt * (p.C / VEC);
  const int64_t want = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 16;
  const unsigned grid = static_cast<unsigned>(want < cap ? want : cap);
  launch_pdl(dw3_kernel<VEC, S, TX>, dim3(grid), dim3(256), static_cast<size_t>(10 * p.C) * sizeof(float), st, p);
}

}  // namespace glsdet

extern "C" int glsdet_dwconv(const void* src, int32_t src_ld, int32_t src_coff, void* dst, int32_t dst_ld,
                             int32_t dst_coff, int32_t batch, int32_t height, int32_t width, int32_t channels,
                             int32_t ksize, int32_t stride, const float* weight, const float* bias, int32_t act,
                             int32_t dtype, void* stream) {
  using namespace glsdet;
  GLSDET_REQUIRE(src && dst && weight && bias && batch > 0 && height > 0 && width > 0 && channels > 0,
                 "dwconv: bad arguments");
  GLSDET_REQUIRE(ksize >= 1 && ksize <= 7 && (ksize & 1) == 1 && (stride == 1 || stride == 2), "dwconv: ksize 1/3/5/7, stride 1/2");
  GLSDET_REQUIRE(dtype == GLSDET_DT_BF16 || dtype == GLSDET_DT_F16 || dtype == GLSDET_DT_F32, "dwconv: unknown dtype %d", dtype);
  GLSDET_REQUIRE(act == GLSDET_ACT_NONE || act == GLSDET_ACT_SILU || act == GLSDET_ACT_RELU || act == GLSDET_ACT_LRELU,
                 "dwconv: unsupported activation %d", act);
  const int vec = (dtype == GLSDET_DT_F32) ? 4 : 8;
  GLSDET_REQUIRE((channels % vec) == 0 && (src_ld % vec) == 0 && (src_coff % vec) == 0 && (dst_ld % vec) == 0 &&
                     (dst_coff % vec) == 0,
                 "dwconv: channels, pitches and offsets must be multiples of %d (16-byte vectors)", vec);
  GLSDET_REQUIRE(src_coff + channels <= src_ld && dst_coff + channels <= dst_ld, "dwconv: window exceeds pitch");
  GLSDET_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(weight) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0,
                 "dwconv: pointers must be 16-byte aligned");
  DwParams p;
  p.src = src; p.dst = dst; p.w = weight; p.bias = bias;
  p.B = batch; p.H = height; p.W = width; p.C = channels;
  const int pad = (ksize - 1) / 2;
  p.Ho = (height + 2 * pad - ksize) / stride + 1;   // nn.Conv2d(k, stride, pad = (k - 1) // 2): baseConv.py:8-10
  p.Wo = (width + 2 * pad - ksize) / stride + 1;
  p.sld = src_ld; p.scoff = src_coff; p.dld = dst_ld; p.dcoff = dst_coff;
  p.k = ksize; p.stride = stride; p.act = act;
  p.f16 = (dtype == GLSDET_DT_F16) ? 1 : 0;
  p.total = static_cast<int64_t>(batch) * p.Ho * p.Wo * (channels / vec);
  const int64_t want = (p.total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 8;   // 8 resident CTAs of 256 threads per SM
  const unsigned grid = static_cast<unsigned>(want < cap ? want : cap);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool generic_only = getenv("GLSDET_DW_GENERIC") != nullptr;   // tests: force the generic kernel
  if (ksize == 3 && channels <= 1024 && !generic_only) {   // weights + bias of the 3x3 kernel: 40 bytes per channel of shared memory
    if (vec == 8 && stride == 1) launch_dw3<8, 1, 4>(p, st);
    else if (vec == 8) launch_dw3<8, 2, 2>(p, st);
    else if (stride == 1) launch_dw3<4, 1, 4>(p, st);
    else launch_dw3<4, 2, 2>(p, st);
    return count_launch("dw3_kernel");
  }
  if (vec == 8) launch_pdl(dwconv_kernel<8>, dim3(grid), dim3(256), 0, st, p);
  else launch_pdl(dwconv_kernel<4>, dim3(grid), dim3(256), 0, st, p);
  return count_launch("dwconv_kernel");
}
