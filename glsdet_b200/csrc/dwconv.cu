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
  int xt, ry;          // 3x3 kernel: CTA tile = ry output rows x xt strips of TX pixels (x all channel vectors)
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

// 3x3 specialisation (every DWConv of the reference is 3x3).  The generic kernel above measured 0.07 - 0.30 of the HBM peak:
// nine dependent load -> FMA rounds per thread, 48 bytes through L1 per tap (16 of activations, 32 of weights), 64-bit index
// arithmetic and an exact SiLU (expf + IEEE division, ~20 instructions) per element - it is ISSUE-bound, not memory-bound.
// Here:
//   * a thread owns TX horizontally adjacent output pixels of one channel vector: per kernel row it issues all
//     (TX - 1) * S + 3 column loads at once (independent: their latencies overlap) and every loaded pixel feeds up to three
//     outputs from registers (S = 1: 4.5 loads and conversions per output instead of 9);
//   * grid = (strips x channel vectors, output row, image): no 64-bit divisions, 32-bit offsets inside an image;
//   * tiles whose 3 x NC input window lies inside the image take a path without any bounds predicate (zero padding only
//     matters for the border strips);
//   * folded weights + bias sit in shared memory, filled BEFORE griddepcontrol.wait (they are constants of the plan), read
//     once per TX outputs, conflict-free (lanes with the same channel vector broadcast);
//   * CTA tile = several output rows x strips (see the kernel): the input rows shared by neighbouring output rows are L1
//     hits (one output row per CTA: 137 us for 3x3/1 C = 64 at 16 x 256^2, tiled: 115 us);
//   * 16-bit storage: SiLU in the tanh form h + h tanh(h), h = x / 2 (one MUFU, three instructions; 2^-11 relative, the
//     form the conv epilogues use); the fp32 accuracy mode keeps expf + IEEE division.
// Measured (tools/dw_bench.py, L2 flushed, 16 x 1024^2 nano shapes): 0.30 - 0.46 of the HBM copy peak on the large layers
// (generic kernel: 0.12 - 0.30), DRAM traffic = the algorithmic bytes (ncu: 134 MB read + 94 MB written for 134 + 134).
// What is left is instruction issue, not memory: 977 instructions per warp (32 outputs x 8 channels: 288 FFMA, 144
// conversions, 96 for SiLU, the rest addressing / packing) at IPC 2.1 with two CTAs per SM (103 registers).  Strips of two
// pixels (more warps, more loads per output) and all 18 loads hoisted in front of the arithmetic measured the same
// (137 - 141 us before the row tiling): the next step would be a vertical register sliding window (each input converted
// once per strip instead of three times).
template <int VEC, int F16>
__device__ __forceinline__ void dw_unpack(const uint4& raw, float (&x)[VEC]) {
  if constexpr (VEC == 8) {
    unpack_16x2(raw.x, F16, x[0], x[1]);
    unpack_16x2(raw.y, F16, x[2], x[3]);
    unpack_16x2(raw.z, F16, x[4], x[5]);
    unpack_16x2(raw.w, F16, x[6], x[7]);
  } else {
    x[0] = __uint_as_float(raw.x); x[1] = __uint_as_float(raw.y);
    x[2] = __uint_as_float(raw.z); x[3] = __uint_as_float(raw.w);
  }
}

template <int VEC, int S, int TX, int F16, bool INTERIOR, bool PRE>
__device__ __forceinline__ void dw3_rows(const DwParams& p, const float* s_w, const uint8_t* img, int oy, int ix0, int c0,
                                         float (&acc)[TX][VEC]) {
  constexpr int NC = (TX - 1) * S + 3;
  constexpr int ES = (VEC == 8) ? 2 : 4;   // element size in bytes
  constexpr int NR = PRE ? 3 : 1;          // PRE: all 3 x NC loads are issued before any arithmetic (one latency per thread)
  uint4 raw[NR][NC];
  auto load_row = [&](int ky, uint4 (&dst)[NC]) {
    const int iy = oy * S + ky - 1;
    const bool rowok = INTERIOR || ((iy >= 0) && (iy < p.H));
    const uint32_t rowoff = static_cast<uint32_t>(iy * p.W + ix0) * static_cast<uint32_t>(p.sld);   // elements; only used when valid
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      if (INTERIOR || (rowok && (ix0 + j >= 0) && (ix0 + j < p.W)))
        dst[j] = __ldg(reinterpret_cast<const uint4*>(img + static_cast<size_t>(rowoff + static_cast<uint32_t>(j * p.sld)) * ES));
      else
        dst[j] = make_uint4(0u, 0u, 0u, 0u);
    }
  };
  if constexpr (PRE) {
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) load_row(ky, raw[ky]);
  }
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    if constexpr (!PRE) load_row(ky, raw[0]);
    const uint4 (&rw)[NC] = raw[PRE ? ky : 0];
    float wk[3][VEC];
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int j = 0; j < VEC; j += 4) {
        const float4 t4 = *reinterpret_cast<const float4*>(s_w + (ky * 3 + kx) * p.C + ((j >> 2) * (p.C / VEC) + c0 / VEC) * 4);
        wk[kx][j] = t4.x; wk[kx][j + 1] = t4.y; wk[kx][j + 2] = t4.z; wk[kx][j + 3] = t4.w;
      }
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      float x[VEC];
      dw_unpack<VEC, F16>(rw[j], x);
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
}

template <int VEC, int S, int TX, int F16, bool PRE>
__global__ void __launch_bounds__(256) dw3_kernel(const DwParams p) {
  extern __shared__ float s_w[];   // [9][C] weights, then [C] bias
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // shared layout [tap | bias][VEC / 4 quads][channel vector][4 floats]: the channel vectors of a quarter-warp read
  // consecutive 16-byte words (conflict-free; [tap][C] order put vectors 32 bytes apart: 5.8-way conflicts in the ncu capture)
  {
    const int nv = p.C / VEC;
    for (int i = threadIdx.x; i < 10 * p.C; i += blockDim.x) {
      const int tap = i / p.C, c = i - tap * p.C;
      const int v = c / VEC, e = c - v * VEC;
      s_w[tap * p.C + ((e >> 2) * nv + v) * 4 + (e & 3)] = (tap < 9) ? __ldg(p.w + i) : __ldg(p.bias + c);
    }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __syncthreads();
  constexpr int NC = (TX - 1) * S + 3;
  constexpr int ES = (VEC == 8) ? 2 : 4;
  // CTA tile = p.ry output rows x p.xt strips x all channel vectors (lanes: channel vector fastest, then strip, then row):
  // the three input rows an output row needs are shared with its neighbours INSIDE the CTA (L1 hits) - with one output row
  // per CTA every input row travelled L2 -> SM three times
  const uint32_t nvec = static_cast<uint32_t>(p.C / VEC);
  const uint32_t wt = static_cast<uint32_t>((p.Wo + TX - 1) / TX);
  const uint32_t slab = static_cast<uint32_t>(p.xt) * nvec;          // threads of one row of the tile
  const uint32_t r = threadIdx.x / slab;
  const uint32_t in_row = threadIdx.x - r * slab;
  const uint32_t xl = in_row / nvec;
  const int xt = static_cast<int>(blockIdx.x * static_cast<uint32_t>(p.xt) + xl);
  const int c0 = static_cast<int>(in_row - xl * nvec) * VEC;
  const int oy = static_cast<int>(blockIdx.y * static_cast<uint32_t>(p.ry) + r);
  const int b = static_cast<int>(blockIdx.z);
  if (r >= static_cast<uint32_t>(p.ry) || static_cast<uint32_t>(xt) >= wt || oy >= p.Ho) return;
  const int ox0 = xt * TX;
  const int ix0 = ox0 * S - 1;
  const uint8_t* img = reinterpret_cast<const uint8_t*>(p.src) +
                       (static_cast<size_t>(b) * p.H * p.W * p.sld + p.scoff + c0) * ES;   // this image, this channel vector
  float acc[TX][VEC];
#pragma unroll
  for (int t = 0; t < TX; ++t)
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[t][j] = s_w[9 * p.C + ((j >> 2) * (p.C / VEC) + c0 / VEC) * 4 + (j & 3)];
  const bool interior = (oy * S >= 1) && (oy * S + 1 < p.H) && (ix0 >= 0) && (ix0 + NC <= p.W);
  if (interior) dw3_rows<VEC, S, TX, F16, true, PRE>(p, s_w, img, oy, ix0, c0, acc);
  else dw3_rows<VEC, S, TX, F16, false, false>(p, s_w, img, oy, ix0, c0, acc);
  uint8_t* orow = reinterpret_cast<uint8_t*>(p.dst) +
                  ((static_cast<size_t>(b) * p.Ho + oy) * p.Wo * p.dld + p.dcoff + c0) * ES;
#pragma unroll
  for (int t = 0; t < TX; ++t) {
    const int ox = ox0 + t;
    if (ox >= p.Wo) break;
    uint8_t* o = orow + static_cast<size_t>(static_cast<uint32_t>(ox) * static_cast<uint32_t>(p.dld)) * ES;
    if constexpr (VEC == 8) {
#pragma unroll
      for (int q = 0; q < VEC; ++q) acc[t][q] = (p.act == GLSDET_ACT_SILU) ? silu_fast(acc[t][q]) : dw_act(acc[t][q], p.act);
      uint4 v;
      v.x = pack_16x2(acc[t][0], acc[t][1], F16);
      v.y = pack_16x2(acc[t][2], acc[t][3], F16);
      v.z = pack_16x2(acc[t][4], acc[t][5], F16);
      v.w = pack_16x2(acc[t][6], acc[t][7], F16);
      *reinterpret_cast<uint4*>(o) = v;
    } else {
#pragma unroll
      for (int q = 0; q < VEC; ++q) acc[t][q] = dw_act(acc[t][q], p.act);
      *reinterpret_cast<float4*>(o) = make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
    }
  }
}

// Vertical streaming variant of the 3x3 kernel (16-bit storage): a thread owns a strip of TX output pixels of one channel
// vector over a BAND of output rows and walks down the input rows; every input row is loaded and converted ONCE and scattered
// into the (up to three) output rows it contributes to, whose accumulators sit in a rotating set of register slots
// (S = 1: three slots, S = 2: two).  Per 16 outputs x 8 channels: 4 loads, 32 conversions, 18 shared-memory weight reads,
// 144 FFMA - the tiled kernel above converts every input three times (S = 1).  The row loop is unrolled by the slot period so
// that slot indices are compile-time register arrays; a slot is stored (if its row is real) and reset to the bias right
// after its last contribution, which also discards what the rows above the band's first output row scattered into it.
template <int VEC, int S, int TX, int F16>
__global__ void __launch_bounds__(256) dw3s_kernel(const DwParams p) {
  extern __shared__ float s_w[];   // conflict-free layout of dw3_kernel
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  {
    const int nv = p.C / VEC;
    for (int i = threadIdx.x; i < 10 * p.C; i += blockDim.x) {
      const int tap = i / p.C, c = i - tap * p.C;
      const int v = c / VEC, e = c - v * VEC;
      s_w[tap * p.C + ((e >> 2) * nv + v) * 4 + (e & 3)] = (tap < 9) ? __ldg(p.w + i) : __ldg(p.bias + c);
    }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __syncthreads();
  constexpr int NC = (TX - 1) * S + 3;
  constexpr int ES = (VEC == 8) ? 2 : 4;
  constexpr int NSLOT = (S == 1) ? 3 : 2;   // output rows in flight
  constexpr int PERIOD = (S == 1) ? 3 : 4;  // input rows per slot rotation
  const uint32_t nvec = static_cast<uint32_t>(p.C / VEC);
  const uint32_t wt = static_cast<uint32_t>((p.Wo + TX - 1) / TX);
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= wt * nvec) return;
  const int xt = static_cast<int>(idx / nvec);
  const int vch = static_cast<int>(idx - static_cast<uint32_t>(xt) * nvec);
  const int c0 = vch * VEC;
  const int b = static_cast<int>(blockIdx.z);
  const int oy0 = static_cast<int>(blockIdx.y) * p.ry;
  const int nrows = min(p.ry, p.Ho - oy0);
  const int nsteps = (S == 1) ? nrows + 2 : 2 * nrows + 1;   // input rows iy = oy0 * S - 1 + j, j = 0 .. nsteps - 1
  const int ox0 = xt * TX;
  const int ix0 = ox0 * S - 1;
  const uint8_t* img = reinterpret_cast<const uint8_t*>(p.src) +
                       (static_cast<size_t>(b) * p.H * p.W * p.sld + p.scoff + c0) * ES;
  uint8_t* oimg = reinterpret_cast<uint8_t*>(p.dst) + (static_cast<size_t>(b) * p.Ho * p.Wo * p.dld + p.dcoff + c0) * ES;
  const bool cols_in = (ix0 >= 0) && (ix0 + NC <= p.W);
  float bias[VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) bias[q] = s_w[9 * p.C + ((q >> 2) * nvec + vch) * 4 + (q & 3)];
  float acc[NSLOT][TX][VEC];
#pragma unroll
  for (int sl = 0; sl < NSLOT; ++sl)
#pragma unroll
    for (int t = 0; t < TX; ++t)
#pragma unroll
      for (int q = 0; q < VEC; ++q) acc[sl][t][q] = bias[q];

  for (int j0 = 0; j0 < nsteps; j0 += PERIOD) {
#pragma unroll
    for (int r = 0; r < PERIOD; ++r) {
      const int j = j0 + r;
      if (j < nsteps) {
        const int iy = oy0 * S - 1 + j;
        float x[NC][VEC];
        const bool rowok = (iy >= 0) && (iy < p.H);
        if (rowok) {
          const uint32_t rowoff = static_cast<uint32_t>(iy * p.W + ix0) * static_cast<uint32_t>(p.sld);
          uint4 raw[NC];
#pragma unroll
          for (int jc = 0; jc < NC; ++jc) {
            if (cols_in || ((ix0 + jc >= 0) && (ix0 + jc < p.W)))
              raw[jc] = __ldg(reinterpret_cast<const uint4*>(img + static_cast<size_t>(rowoff + static_cast<uint32_t>(jc * p.sld)) * ES));
            else
              raw[jc] = make_uint4(0u, 0u, 0u, 0u);
          }
#pragma unroll
          for (int jc = 0; jc < NC; ++jc) dw_unpack<VEC, F16>(raw[jc], x[jc]);
          // contributions of this input row: (kernel row ky, slot of the output row it goes to), compile-time per r
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            // S = 1: output row q = j - ky;  S = 2: j = 2q + ky  ->  q = (j - ky) / 2 when j - ky is even
            const bool has = (S == 1) ? true : (((r - ky) & 1) == 0);
            if (has) {
              const int sl = (S == 1) ? ((r - ky + 3) % 3) : ((((r - ky + 4) >> 1)) & 1);   // slot of q (q mod NSLOT; j0 is a multiple of the period)
              float wk[3][VEC];
#pragma unroll
              for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int q4 = 0; q4 < VEC; q4 += 4) {
                  const float4 t4 = *reinterpret_cast<const float4*>(s_w + (ky * 3 + kx) * p.C + ((q4 >> 2) * nvec + vch) * 4);
                  wk[kx][q4] = t4.x; wk[kx][q4 + 1] = t4.y; wk[kx][q4 + 2] = t4.z; wk[kx][q4 + 3] = t4.w;
                }
#pragma unroll
              for (int jc = 0; jc < NC; ++jc)
#pragma unroll
                for (int t = 0; t < TX; ++t) {
                  const int kx = jc - t * S;
                  if (kx >= 0 && kx < 3) {
#pragma unroll
                    for (int q = 0; q < VEC; ++q) acc[sl][t][q] = fmaf(x[jc][q], wk[kx][q], acc[sl][t][q]);
                  }
                }
            }
          }
        }
        // the output row that received its LAST contribution (ky = 2) from this input row: S = 1: q = j - 2; S = 2: even j,
        // q = (j - 2) / 2.  Store it if it is a real row of the band, then reset the slot.
        const bool done = (S == 1) ? true : ((r & 1) == 0);
        if (done) {
          const int sl = (S == 1) ? ((r + 1) % 3) : (((r + 2) >> 1) & 1);
          const int q_out = (S == 1) ? (j - 2) : ((j - 2) >> 1);
          if (j >= 2 && q_out < nrows) {
            uint8_t* orow = oimg + static_cast<size_t>(static_cast<uint32_t>((oy0 + q_out) * p.Wo) * static_cast<uint32_t>(p.dld)) * ES;
#pragma unroll
            for (int t = 0; t < TX; ++t) {
              const int ox = ox0 + t;
              if (ox < p.Wo) {
                float o[VEC];
#pragma unroll
                for (int q = 0; q < VEC; ++q)
                  o[q] = (VEC == 8 && p.act == GLSDET_ACT_SILU) ? silu_fast(acc[sl][t][q]) : dw_act(acc[sl][t][q], p.act);
                uint8_t* op = orow + static_cast<size_t>(static_cast<uint32_t>(ox) * static_cast<uint32_t>(p.dld)) * ES;
                if constexpr (VEC == 8) {
                  uint4 v;
                  v.x = pack_16x2(o[0], o[1], F16);
                  v.y = pack_16x2(o[2], o[3], F16);
                  v.z = pack_16x2(o[4], o[5], F16);
                  v.w = pack_16x2(o[6], o[7], F16);
                  *reinterpret_cast<uint4*>(op) = v;
                } else {
                  *reinterpret_cast<float4*>(op) = make_float4(o[0], o[1], o[2], o[3]);
                }
              }
            }
          }
#pragma unroll
          for (int t = 0; t < TX; ++t)
#pragma unroll
            for (int q = 0; q < VEC; ++q) acc[sl][t][q] = bias[q];
        }
      }
    }
  }
}

template <int VEC, int S, int TX, int F16>
static void launch_dw3s(DwParams p, cudaStream_t st) {
  const int wt = (p.Wo + TX - 1) / TX;
  const int per_row = wt * (p.C / VEC);
  const int threads = per_row >= 256 ? 256 : ((per_row + 31) / 32) * 32;
  const int gx = (per_row + threads - 1) / threads;
  // band height: long bands re-read fewer halo rows ((ry + 2) / ry), short ones give the device enough CTAs
  int ry = 16;
  const int64_t want = 4ll * device_sm_count();
  while (ry > 2 && static_cast<int64_t>(gx) * ((p.Ho + ry - 1) / ry) * p.B < want) ry >>= 1;
  p.ry = ry;
  const dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>((p.Ho + ry - 1) / ry), static_cast<unsigned>(p.B));
  launch_pdl(dw3s_kernel<VEC, S, TX, F16>, grid, dim3(threads), static_cast<size_t>(10 * p.C) * sizeof(float), st, p);
}

template <int VEC, int S, int TX, int F16, bool PRE = false>
static void launch_dw3(DwParams p, cudaStream_t st) {
  const int wt = (p.Wo + TX - 1) / TX;
  const int nvec = p.C / VEC;
  int xt = 64 / nvec;                   // about 64 threads per tile row
  if (xt < 1) xt = 1;
  if (xt > wt) xt = wt;
  int ry = 256 / (xt * nvec);           // rows of the tile: 256 threads per CTA
  if (ry < 1) ry = 1;
  if (ry > p.Ho) ry = p.Ho;
  p.xt = xt; p.ry = ry;
  const int threads = ((xt * nvec * ry + 31) / 32) * 32;
  const dim3 grid(static_cast<unsigned>((wt + xt - 1) / xt), static_cast<unsigned>((p.Ho + ry - 1) / ry), static_cast<unsigned>(p.B));
  launch_pdl(dw3_kernel<VEC, S, TX, F16, PRE>, grid, dim3(threads), static_cast<size_t>(10 * p.C) * sizeof(float), st, p);
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
  const bool fits32 = static_cast<int64_t>(height) * width * src_ld < (1ll << 31) && p.Ho <= 65535 && batch <= 65535;
  p.xt = p.ry = 1;
  if (ksize == 3 && channels <= 1024 && fits32 && !generic_only) {   // weights + bias of the 3x3 kernel: 40 bytes per channel of shared memory
    const char* se = getenv("GLSDET_DW_STREAM");   // A/B knob: 0 = the row-tiled kernel for every shape
    const bool stream_rows = !(se && se[0] == '0');
    if (vec == 8 && stream_rows) {
      if (stride == 1) { if (p.f16) launch_dw3s<8, 1, 2, 1>(p, st); else launch_dw3s<8, 1, 2, 0>(p, st); }
      else { if (p.f16) launch_dw3s<8, 2, 2, 1>(p, st); else launch_dw3s<8, 2, 2, 0>(p, st); }
      return count_launch("dw3s_kernel");
    }
    if (vec == 8 && stride == 1) { if (p.f16) launch_dw3<8, 1, 4, 1>(p, st); else launch_dw3<8, 1, 4, 0>(p, st); }
    else if (vec == 8) { if (p.f16) launch_dw3<8, 2, 2, 1>(p, st); else launch_dw3<8, 2, 2, 0>(p, st); }
    else if (stride == 1) launch_dw3<4, 1, 4, 0>(p, st);
    else launch_dw3<4, 2, 2, 0>(p, st);
    return count_launch("dw3_kernel");
  }
  if (vec == 8) launch_pdl(dwconv_kernel<8>, dim3(grid), dim3(256), 0, st, p);
  else launch_pdl(dwconv_kernel<4>, dim3(grid), dim3(256), 0, st, p);
  return count_launch("dwconv_kernel");
}
