// MP-Det head pieces that are not convolutions (yolox-ufp/mmdet/models/dense_heads/mp_head.py, gfl_head.py):
// GroupNorm + ReLU of the shared towers, the proxy classification, the integral box decode and the per-level
// candidate selection of _get_bboxes_single.  The convolutions (FPN, towers, gfl_cls_conv, gfl_reg) run on
// conv_gemm_kernel.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstring>

#include "../../include/glsdet_b200.h"
#include "common.h"
#include "ptx.cuh"

namespace glsdet {

constexpr int kGnSlabs = 64;   // pixel slabs per image: 64 x batch CTAs read the map (16 left a 69 MB map to 128 CTAs: 1.7 TB/s)

// ---------------------------------------------------------------------------------------------- GroupNorm + ReLU
// stage 1: per (image, slab of pixels) channel sums and sums of squares, fixed order (deterministic).
// stage 2 runs in the last CTA of an image to finish (arrival counter in `counters`, left at zero): group statistics ->
// per (image, channel) scale a and shift s with  y = relu(x * a + s).  One thread per channel sums its slab partials in
// fixed order, then the channels of a group are added in fixed order.  (A separate finalize launch cost 11 us x 40.)
__global__ void __launch_bounds__(256, 4) gn_partial_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ scratch,
                                                         int HW, int C, int ld, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float* __restrict__ ab,
                                                         int* __restrict__ counters, int groups, float eps) {
  pdl_prologue();
  const int slab = blockIdx.x, b = blockIdx.y;
  const int per = (HW + kGnSlabs - 1) / kGnSlabs;
  const int p0 = slab * per, p1 = min(HW, p0 + per);
  const int nvec = C >> 3;
  // thread -> (vector of 8 channels, pixel lane); 256 threads = nvec x (256 / nvec) pixel lanes
  const int lanes = 256 / nvec;
  const int v = threadIdx.x % nvec, pl = threadIdx.x / nvec;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (pl < lanes) {
    for (int p = p0 + pl; p < p1; p += lanes) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (static_cast<int64_t>(b) * HW + p) * ld) + v);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float a = __uint_as_float(w[k] << 16), c = __uint_as_float(w[k] & 0xFFFF0000u);
        s[2 * k] += a; q[2 * k] += a * a;
        s[2 * k + 1] += c; q[2 * k + 1] += c * c;
      }
    }
  }
  extern __shared__ __align__(16) unsigned char gn_smem[];
  float* red = reinterpret_cast<float*>(gn_smem);   // [lanes][2][C]
  if (pl < lanes) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[(pl * 2 + 0) * C + v * 8 + j] = s[j];
      red[(pl * 2 + 1) * C + v * 8 + j] = q[j];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += 256) {
    const int which = c / C, ch = c % C;
    float t = 0.0f;
    for (int l = 0; l < lanes; ++l) t += red[(l * 2 + which) * C + ch];
    scratch[((static_cast<int64_t>(b) * kGnSlabs + slab) * 2 + which) * C + ch] = t;
  }
  __shared__ int last_s;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last_s = (atomicAdd(&counters[b], 1) == kGnSlabs - 1);
  __syncthreads();
  if (!last_s) return;
  __threadfence();
  const int cpg = C / groups;
  double* fin = reinterpret_cast<double*>(gn_smem);   // [2][C] channel sums
  __shared__ float mean_s[64], rstd_s[64];
  for (int c = threadIdx.x; c < C; c += 256) {
    // 16 slabs (32 loads) in flight per batch: the loads, not the adds, pace this tail (one dependent load at a time took
    // 11 us), while 64 at once cost the streaming loop above its occupancy (128 registers)
    double sm = 0.0, sq = 0.0;
    for (int k0 = 0; k0 < kGnSlabs; k0 += 16) {
      float ps[16], pq[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        ps[k] = __ldcg(&scratch[((static_cast<int64_t>(b) * kGnSlabs + k0 + k) * 2 + 0) * C + c]);
        pq[k] = __ldcg(&scratch[((static_cast<int64_t>(b) * kGnSlabs + k0 + k) * 2 + 1) * C + c]);
      }
#pragma unroll
      for (int k = 0; k < 16; ++k) { sm += ps[k]; sq += pq[k]; }
    }
    fin[c] = sm;
    fin[C + c] = sq;
  }
  __syncthreads();
  for (int g = threadIdx.x; g < groups; g += 256) {
    double sm = 0.0, sq = 0.0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) { sm += fin[c]; sq += fin[C + c]; }
    const double n = static_cast<double>(HW) * cpg;
    const double m = sm / n;
    const double var = fmax(sq / n - m * m, 0.0);
    mean_s[g] = static_cast<float>(m);
    rstd_s[g] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    const int g = c / cpg;
    const float a = gamma[c] * rstd_s[g];
    ab[(static_cast<int64_t>(b) * 2 + 0) * C + c] = a;
    ab[(static_cast<int64_t>(b) * 2 + 1) * C + c] = beta[c] - mean_s[g] * a;
  }
  if (threadIdx.x == 0) counters[b] = 0;
}

// stage 3: in place  x = relu(x * a[b][c] + s[b][c]).  A CTA owns one pixel slab of one image, so a thread keeps the
// scale / shift of its eight channels in registers and streams 16-byte vectors (round 1 re-read 16 scalars per vector
// through L1: 1.9 TB/s on the 100 x 168 level).
__global__ void __launch_bounds__(256) gn_apply_kernel(__nv_bfloat16* __restrict__ x, const float* __restrict__ ab, int HW, int C,
                                                       int ld) {
  pdl_prologue();
  const int slab = blockIdx.x, b = blockIdx.y;
  const int per = (HW + gridDim.x - 1) / gridDim.x;
  const int p0 = slab * per, p1 = min(HW, p0 + per);
  const int nvec = C >> 3;
  const int lanes = 256 / nvec;                       // pixels in flight per iteration
  const int v = threadIdx.x % nvec, pl = threadIdx.x / nvec;
  if (pl >= lanes) return;
  float a[8], sh[8];
  const float* ap = ab + (static_cast<int64_t>(b) * 2) * C + v * 8;
#pragma unroll
  for (int k = 0; k < 8; ++k) { a[k] = ap[k]; sh[k] = ap[C + k]; }
  for (int p = p0 + pl; p < p1; p += lanes) {
    uint4* ptr = reinterpret_cast<uint4*>(x + (static_cast<int64_t>(b) * HW + p) * ld) + v;
    const uint4 u = *ptr;
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float lo = fmaxf(fmaf(__uint_as_float(w[k] << 16), a[2 * k], sh[2 * k]), 0.0f);
      const float hi = fmaxf(fmaf(__uint_as_float(w[k] & 0xFFFF0000u), a[2 * k + 1], sh[2 * k + 1]), 0.0f);
      o[k] = pack_bf16x2(lo, hi);
    }
    *ptr = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ---------------------------------------------------------------------------------------------- proxy classification
// mp_head.py:105-121: one warp per pixel.  feat fp32 NHWC [P, C]; centers fp32 [n_prox, C] L2-normalised on the host;
// cls_start[c] .. cls_start[c+1] = proxies of class c.  rows[pixel_row0 + p][col0 + c] = gamma * sum_j softmax(gamma s)_j s_j
// (raw class score), with s_j = <feat / |feat|, center_j>.
__global__ void __launch_bounds__(256) proxy_scores_kernel(const float* __restrict__ feat, const float* __restrict__ centers,
                                                           const int* __restrict__ cls_start, int nc, int n_prox, int C, int HW,
                                                           int64_t total_pix, float gamma, float* __restrict__ rows, int rows_ld,
                                                           int64_t rows_bs, int row0) {
  // Persistent CTAs: the normalised proxies (42 x 256 floats = 43 KB) are staged ONCE per CTA - round 1 launched one CTA
  // per 8 pixels, each re-reading all proxies (720 MB of L2 traffic at the 100 x 168 level) - and a pixel's features stay
  // in registers for all dot products.  Summation order (lane-strided partial sums, butterfly reduce) as before.
  extern __shared__ float sc[];   // centers [n_prox][C]
  __shared__ int cls_s[33];       // class -> first proxy (read per pixel and class: from global it was a dependent-load chain)
  pdl_prologue();
  for (int i = threadIdx.x; i < n_prox * C; i += blockDim.x) sc[i] = centers[i];
  if (threadIdx.x <= nc) cls_s[threadIdx.x] = cls_start[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  constexpr int kMaxPer = 16;   // channels per lane held in registers (C <= 512)
  const int per = C >> 5;
  for (int64_t pix = static_cast<int64_t>(blockIdx.x) * wpb + warp; pix < total_pix; pix += static_cast<int64_t>(gridDim.x) * wpb) {
    const float* f = feat + pix * C;
    float fv[kMaxPer];
    float nn = 0.0f;
#pragma unroll
    for (int k = 0; k < kMaxPer; ++k) {
      fv[k] = (k < per) ? __ldg(f + lane + 32 * k) : 0.0f;
      if (k < per) nn += fv[k] * fv[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
    const float inv = 1.0f / fmaxf(sqrtf(nn), 1e-12f);   // F.normalize eps
    float my_sim = 0.0f;                                  // lane j keeps the similarity of proxies j and j + 32
    float my_sim2 = 0.0f;
    // six proxies per pass: six independent dot products and butterfly reductions in flight (the 42 serial
    // shuffle chains of one proxy at a time left the schedulers idle)
    for (int j0 = 0; j0 < n_prox; j0 += 6) {
      float d[6];
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        d[u] = 0.0f;
        const float* cj = sc + min(j0 + u, n_prox - 1) * C + lane;
#pragma unroll
        for (int k = 0; k < kMaxPer; ++k)
          if (k < per) d[u] += fv[k] * cj[32 * k];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int u = 0; u < 6; ++u) d[u] += __shfl_xor_sync(0xffffffffu, d[u], o);
      }
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int j = j0 + u;
        if (j < n_prox) {
          const float v = d[u] * inv;
          if (j == lane) my_sim = v;
          if (j == lane + 32) my_sim2 = v;
        }
      }
    }
    // class aggregation: lane c handles class c
    float out = 0.0f;
    for (int c = 0; c < nc; ++c) {
      const int j0 = cls_s[c], j1 = cls_s[c + 1];
      float mx = -1e30f;
      for (int j = j0; j < j1; ++j) {
        const float s = (j < 32) ? __shfl_sync(0xffffffffu, my_sim, j) : __shfl_sync(0xffffffffu, my_sim2, j - 32);
        mx = fmaxf(mx, s * gamma);
      }
      float den = 0.0f, num = 0.0f;
      for (int j = j0; j < j1; ++j) {
        const float s = (j < 32) ? __shfl_sync(0xffffffffu, my_sim, j) : __shfl_sync(0xffffffffu, my_sim2, j - 32);
        const float e = expf(s * gamma - mx);
        den += e;
        num += e * s;
      }
      if (lane == c) out = gamma * num / den;
    }
    const int b = static_cast<int>(pix / HW);
    const int p = static_cast<int>(pix % HW);
    if (lane < nc) rows[b * rows_bs + static_cast<int64_t>(row0 + p) * rows_ld + lane] = out;
  }
}

// Proxy classification on the tensor core: the similarities are a 1x1 conv of the (bf16) class features with the
// normalised proxies as weights (conv_gemm_kernel, N = proxies padded to 16, fp32 output); what is left per pixel is the
// L2 norm of the features and the per-class softmax-weighted sum - one warp per pixel, ~100 instructions instead of the
// 1800 of the all-SIMT kernel above (42 dot products of 256 with a butterfly reduction each).
__global__ void __launch_bounds__(256) proxy_aggregate_kernel(const __nv_bfloat16* __restrict__ feat, const float* __restrict__ sims,
                                                              int sims_ld, const int* __restrict__ cls_start, int nc, int C, int HW,
                                                              int64_t total_pix, float gamma, float* __restrict__ rows, int rows_ld,
                                                              int64_t rows_bs, int row0) {
  __shared__ int cls_s[33];
  pdl_prologue();
  if (threadIdx.x <= nc) cls_s[threadIdx.x] = cls_start[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const int nvec = C >> 3;   // 16-byte vectors per pixel
  for (int64_t pix = static_cast<int64_t>(blockIdx.x) * wpb + warp; pix < total_pix; pix += static_cast<int64_t>(gridDim.x) * wpb) {
    const uint4* f = reinterpret_cast<const uint4*>(feat + pix * C);
    // the class lanes fetch their similarities first, so that both loads of a pixel are in flight together
    const float* sp = sims + pix * sims_ld;
    const int j0 = lane < nc ? cls_s[lane] : 0, j1 = lane < nc ? cls_s[lane + 1] : 0;
    float sv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) sv[k] = (j0 + k < j1) ? __ldg(sp + j0 + k) : 0.0f;
    float nn = 0.0f;
    for (int v = lane; v < nvec; v += 32) {
      const uint4 u = __ldg(f + v);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float a = __uint_as_float(w[k] << 16), c = __uint_as_float(w[k] & 0xFFFF0000u);
        nn = fmaf(a, a, nn);
        nn = fmaf(c, c, nn);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
    const float inv = 1.0f / fmaxf(sqrtf(nn), 1e-12f);   // F.normalize eps
    if (lane < nc) {   // lane c = class c: softmax(gamma s)-weighted sum of its proxies' similarities (mp_head.py:112-118)
      float mx = -1e30f;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (j0 + k < j1) mx = fmaxf(mx, sv[k] * inv * gamma);
      for (int j = j0 + 8; j < j1; ++j) mx = fmaxf(mx, sp[j] * inv * gamma);
      float den = 0.0f, num = 0.0f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (j0 + k < j1) {
          const float sj = sv[k] * inv;
          const float e = expf(sj * gamma - mx);
          den += e;
          num += e * sj;
        }
      }
      for (int j = j0 + 8; j < j1; ++j) {
        const float sj = sp[j] * inv;
        const float e = expf(sj * gamma - mx);
        den += e;
        num += e * sj;
      }
      const int b = static_cast<int>(pix / HW);
      const int p = static_cast<int>(pix % HW);
      rows[b * rows_bs + static_cast<int64_t>(row0 + p) * rows_ld + lane] = gamma * num / den;
    }
  }
}

// ---------------------------------------------------------------------------------------------- integral box decode
// gfl_head.py:35-49,437-438,456-457: reg fp32 NHWC [P, reg_ld] holds 4 x (reg_max + 1) logits per pixel (Scale already
// folded into the conv); distance = softmax-expectation over the bins * stride; box = point -/+ distance, clamped.
__global__ void __launch_bounds__(256) gfl_decode_kernel(const float* __restrict__ reg, int reg_ld, int bins, int H, int W,
                                                         float stride, float max_x, float max_y, int64_t total,
                                                         float* __restrict__ boxes, int64_t boxes_bs, int row0) {
  pdl_prologue();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // (pixel, side)
  if (i >= total) return;
  const int side = static_cast<int>(i & 3);
  const int64_t pix = i >> 2;
  const float* r = reg + pix * reg_ld + side * bins;
  float mx = -1e30f;
  for (int k = 0; k < bins; ++k) mx = fmaxf(mx, __ldg(r + k));
  float den = 0.0f, num = 0.0f;
  for (int k = 0; k < bins; ++k) {
    const float e = expf(__ldg(r + k) - mx);
    den += e;
    num += e * static_cast<float>(k);
  }
  const float d = num / den * stride;
  const int HW = H * W;
  const int b = static_cast<int>(pix / HW), p = static_cast<int>(pix % HW);
  const float px = static_cast<float>(p % W) * stride, py = static_cast<float>(p / W) * stride;
  float v = (side == 0) ? px - d : (side == 1) ? py - d : (side == 2) ? px + d : py + d;
  v = fminf(fmaxf(v, 0.0f), (side & 1) ? max_y : max_x);
  boxes[b * boxes_bs + (static_cast<int64_t>(row0 + p)) * 4 + side] = v;
}

// ---------------------------------------------------------------------------------------------- candidate selection
// filter_scores_and_topk (core/utils/misc.py:143-165) for one level of one image per CTA: candidates = (anchor, class)
// pairs with sigmoid(score) > thr, best `topk` by score (ties: lower flattened index first), appended to the image's
// candidate list (boxes, scores, labels).  Keys are sorted in global memory by a single-CTA bitonic network whose size
// follows the candidate count.
constexpr int kSelBins = 2048;   // histogram of the score bits above the threshold, 2^15 ulps (2^-8 relative) per bin
constexpr int kSelCap = 4096;    // candidates sorted in shared memory
constexpr int kSelScratch = kSelBins + 8;   // ints per image: histogram + collected-count

// Only the best `topk` of up to anchors x classes candidates are needed, so the full sort of round 1 (a single-CTA bitonic
// network over up to 2^18 keys in global memory: 1.9 ms per level, half of an MP-Det step) is replaced by a selection in
// three small kernels: (1) histogram of the score bits over the whole grid, (2) every CTA finds the bin b* in which the
// topk-th best score lies and collects the candidates of the bins >= b* (topk plus the rest of one bin), (3) one CTA per
// image sorts only those - in shared memory - and appends the best topk to the image's candidate list.
// Same keys as before ((~score bits) << 32 | flattened index), so the result is identical.
// bins of 2^shift ulps above the threshold; the host picks shift so that score 1.0 lands below kSelBins (no clamping:
// sub-bins of a bin stay ordered, see the refinement in the sort kernel)
__device__ __forceinline__ int sel_bin(uint32_t score_bits, uint32_t thr_bits, int shift) {
  return min(kSelBins - 1, static_cast<int>((score_bits - thr_bits) >> shift));
}

__global__ void __launch_bounds__(256) gfl_select_hist_kernel(const float* __restrict__ rows, int rows_ld, int64_t rows_bs, int row0,
                                                              int A_l, int nc, float thr, int shift, int* __restrict__ scratch) {
  pdl_prologue();
  __shared__ int hist[kSelBins];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < kSelBins; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const float* rb = rows + b * rows_bs + static_cast<int64_t>(row0) * rows_ld;
  const int total = A_l * nc;
  const uint32_t thr_bits = __float_as_uint(fmaxf(thr, 0.0f));
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int a = i / nc, c = i - a * nc;
    const float sc = 1.0f / (1.0f + expf(-rb[static_cast<int64_t>(a) * rows_ld + c]));
    if (sc > thr) atomicAdd(&hist[sel_bin(__float_as_uint(sc), thr_bits, shift)], 1);
  }
  __syncthreads();
  int* gh = scratch + static_cast<int64_t>(b) * kSelScratch;
  for (int i = threadIdx.x; i < kSelBins; i += blockDim.x)
    if (hist[i]) atomicAdd(&gh[i], hist[i]);
}

// Whole CTA (blockDim a multiple of 32, at most 32 warps, kSelBins / warps a multiple of 64): b* = the bin in which the
// topk-th best score lies (0 if there are fewer candidates) and the total candidate count, from an image's histogram in
// shared memory.  Warp w owns the bins [kSelBins - (w + 1) * seg, kSelBins - w * seg), best scores first; every warp adds up
// its segment, then the one warp whose segment holds the crossing walks it again (one warp over all 2048 bins: 32
// dependent rounds, 10 us per launch).  Call from all threads; results are valid after the closing barrier.
__device__ __forceinline__ void sel_threshold_bin(const int* hist, int topk, int* bstar_s, int* total_s, int* wsum_s) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int seg = kSelBins / nwarps;
  const int hi0 = kSelBins - warp * seg;
  int mine = 0;
  for (int hi = hi0; hi > hi0 - seg; hi -= 64) {
    const int b0 = hi - 1 - 2 * lane;
    mine += hist[b0] + hist[b0 - 1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if (lane == 0) wsum_s[warp] = mine;
  if (threadIdx.x == 0) *bstar_s = 0;
  __syncthreads();
  int before_w = 0, total = 0;
  for (int w = 0; w < nwarps; ++w) {
    const int v = wsum_s[w];
    if (w < warp) before_w += v;
    total += v;
  }
  if (threadIdx.x == 0) *total_s = total;
  if (before_w < topk && before_w + mine >= topk) {   // warp-uniform: the crossing is in this segment
    int cum = before_w;
    for (int hi = hi0; hi > hi0 - seg; hi -= 64) {
      const int b0 = hi - 1 - 2 * lane;          // lanes own bins (b0, b0 - 1), descending
      const int h0 = hist[b0], h1 = hist[b0 - 1];
      int incl = h0 + h1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const int before = cum + incl - h0 - h1;     // candidates in strictly better bins than b0
      const bool hit0 = before + h0 >= topk, hit1 = before + h0 + h1 >= topk;
      const unsigned m = __ballot_sync(0xffffffffu, hit0 || hit1);
      if (m) {
        const int l = __ffs(m) - 1;
        const int bs = __shfl_sync(0xffffffffu, hit0 ? b0 : b0 - 1, l);
        if (lane == 0) *bstar_s = bs;
        break;
      }
      cum += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) gfl_select_collect_kernel(const float* __restrict__ rows, int rows_ld, int64_t rows_bs, int row0,
                                                                 int A_l, int nc, float thr, int shift, int topk, int* __restrict__ scratch,
                                                                 unsigned long long* __restrict__ keys, int64_t keys_bs) {
  pdl_prologue();
  __shared__ int bstar_s;
  __shared__ int hs[kSelBins];   // the scan walks the bins serially: from shared memory (25 us of dependent global loads otherwise)
  const int b = blockIdx.y;
  int* gh = scratch + static_cast<int64_t>(b) * kSelScratch;
  for (int i = threadIdx.x; i < kSelBins; i += blockDim.x) hs[i] = __ldcg(&gh[i]);
  __syncthreads();
  __shared__ int total_s, wsum_s[32];
  sel_threshold_bin(hs, topk, &bstar_s, &total_s, wsum_s);
  const int bstar = bstar_s;
  unsigned long long* kb = keys + b * keys_bs;
  const float* rb = rows + b * rows_bs + static_cast<int64_t>(row0) * rows_ld;
  const int total = A_l * nc;
  const uint32_t thr_bits = __float_as_uint(fmaxf(thr, 0.0f));
  const int lane = threadIdx.x & 31;
  const int rounded = (total + 31) & ~31;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rounded; i += gridDim.x * blockDim.x) {
    bool hit = false;
    float sc = 0.0f;
    if (i < total) {
      const int a = i / nc, c = i - a * nc;
      sc = 1.0f / (1.0f + expf(-rb[static_cast<int64_t>(a) * rows_ld + c]));
      hit = sc > thr && sel_bin(__float_as_uint(sc), thr_bits, shift) >= bstar;
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (m) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&gh[kSelBins], __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      // descending score, ascending index: key = (~score_bits << 32) | index, sorted ascending
      if (hit) kb[base + __popc(m & ((1u << lane) - 1u))] =
          (static_cast<unsigned long long>(~__float_as_uint(sc)) << 32) | static_cast<unsigned int>(i);
    }
  }
}

__global__ void __launch_bounds__(1024) gfl_select_sort_kernel(const float* __restrict__ boxes, int64_t boxes_bs, int row0, int nc,
                                                               float thr, int shift, int topk, int* __restrict__ scratch,
                                                               unsigned long long* __restrict__ keys, int64_t keys_bs,
                                                               int* __restrict__ cand_count, float* __restrict__ cboxes,
                                                               float* __restrict__ cscores, float* __restrict__ clabels, int cap) {
  pdl_prologue();
  const int b = blockIdx.x;
  __shared__ unsigned long long sk[kSelCap];
  __shared__ int base_s, n_s, bstar_s, wsum_s[32], above_s, sstar_s, sub_total_s, m2_s;
  int* gh = scratch + static_cast<int64_t>(b) * kSelScratch;
  unsigned long long* kb = keys + b * keys_bs;
  int* hs = reinterpret_cast<int*>(sk);   // staged histograms (the sort buffer is not in use yet)
  for (int i = threadIdx.x; i < kSelBins; i += blockDim.x) hs[i] = __ldcg(&gh[i]);
  if (threadIdx.x == 0) { above_s = 0; m2_s = 0; }
  __syncthreads();
  sel_threshold_bin(hs, topk, &bstar_s, &n_s, wsum_s);
  const int n = n_s, bstar = bstar_s;
  int m = __ldcg(&gh[kSelBins]);
  const int take = min(n, topk);
  __syncthreads();
  for (int i = threadIdx.x; i < kSelScratch; i += blockDim.x) gh[i] = 0;   // clean for the next level / call
  const uint32_t thr_bits = __float_as_uint(fmaxf(thr, 0.0f));
  bool staged = false;
  if (m > kSelCap / 2) {
    // Many candidates share the crossing bin (narrow score distributions): refine inside it with 2048 sub-bins, keep
    // everything above the crossing sub-bin, so that the network below sorts ~topk keys instead of thousands.
    const int sub_shift = shift > 11 ? shift - 11 : 0;
    const uint32_t sub_mask = (shift >= 32) ? 0xFFFFFFFFu : ((1u << shift) - 1u);
    for (int i = threadIdx.x; i < kSelBins; i += blockDim.x) hs[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int i0 = 0; i0 < m; i0 += blockDim.x) {
      const int i = i0 + threadIdx.x;
      bool up = false;
      if (i < m) {
        const uint32_t bits = ~static_cast<uint32_t>(kb[i] >> 32);
        const int bin = sel_bin(bits, thr_bits, shift);
        up = bin > bstar;
        if (bin == bstar) atomicAdd(&hs[min(kSelBins - 1, static_cast<int>(((bits - thr_bits) & sub_mask) >> sub_shift))], 1);
      }
      const unsigned mk = __ballot_sync(0xffffffffu, up);
      if (lane == 0 && mk) atomicAdd(&above_s, __popc(mk));
    }
    __syncthreads();
    const int above = above_s;
    sel_threshold_bin(hs, topk - above, &sstar_s, &sub_total_s, wsum_s);
    const int sstar = sstar_s;
    __syncthreads();   // hs (= sk) is free from here
    // survivors: bins above b*, and the sub-bins >= s* of b*
    for (int i0 = 0; i0 < m; i0 += blockDim.x) {
      const int i = i0 + threadIdx.x;
      bool keep = false;
      unsigned long long key = 0;
      if (i < m) {
        key = kb[i];
        const uint32_t bits = ~static_cast<uint32_t>(key >> 32);
        const int bin = sel_bin(bits, thr_bits, shift);
        keep = bin > bstar ||
               (bin == bstar && min(kSelBins - 1, static_cast<int>(((bits - thr_bits) & sub_mask) >> sub_shift)) >= sstar);
      }
      const unsigned mk = __ballot_sync(0xffffffffu, keep);
      int base = 0;
      if (lane == 0 && mk) base = atomicAdd(&m2_s, __popc(mk));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (keep) {
        const int slot = base + __popc(mk & ((1u << lane) - 1u));
        if (slot < kSelCap) sk[slot] = key;
      }
    }
    __syncthreads();
    if (m2_s <= kSelCap) { m = m2_s; staged = true; }   // else (ties beyond the sub-bins): sort everything in global memory
    __syncthreads();
  }
  int P = 1, lp = 0;
  while (P < m) { P <<= 1; ++lp; }
  const bool in_smem = staged || m <= kSelCap;
  unsigned long long* sv = in_smem ? sk : kb;
  if (in_smem && !staged)
    for (int i = threadIdx.x; i < m; i += blockDim.x) sk[i] = kb[i];
  for (int i = m + threadIdx.x; i < P; i += blockDim.x) sv[i] = ~0ull;
  __syncthreads();
  for (int lk = 1; lk <= lp; ++lk) {
    const int k = 1 << lk;
    for (int lj = lk - 1; lj >= 0; --lj) {
      const int j = 1 << lj;
      for (int t = threadIdx.x; t < P / 2; t += blockDim.x) {
        const int i = ((t >> lj) << (lj + 1)) | (t & (j - 1));
        const bool asc = ((i & k) == 0);
        const unsigned long long x = sv[i], y = sv[i + j];
        if ((x > y) == asc) { sv[i] = y; sv[i + j] = x; }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) base_s = atomicAdd(&cand_count[b], take);
  __syncthreads();
  const int base = base_s;
  for (int t = threadIdx.x; t < take; t += blockDim.x) {
    if (base + t >= cap) continue;
    const unsigned long long key = sv[t];
    const unsigned int idx = static_cast<unsigned int>(key & 0xFFFFFFFFull);
    const int a = idx / nc, c = idx - a * nc;
    const float sc = __uint_as_float(~static_cast<unsigned int>(key >> 32));
    const float4 bx = *reinterpret_cast<const float4*>(boxes + b * boxes_bs + (static_cast<int64_t>(row0) + a) * 4);
    *reinterpret_cast<float4*>(cboxes + (static_cast<int64_t>(b) * cap + base + t) * 4) = bx;
    cscores[static_cast<int64_t>(b) * cap + base + t] = sc;
    clabels[static_cast<int64_t>(b) * cap + base + t] = static_cast<float>(c);
  }
}

}  // namespace glsdet

using namespace glsdet;

extern "C" int glsdet_group_norm_relu(void* x, int32_t batch, int32_t hw, int32_t channels, int32_t x_ld, int32_t groups,
                                      const float* gamma, const float* beta, float eps, float* scratch, void* stream) {
  GLSDET_REQUIRE(x && gamma && beta && scratch && batch > 0 && hw > 0, "group_norm_relu: bad arguments");
  GLSDET_REQUIRE(channels > 0 && (channels % 8) == 0 && channels <= 2048 && (256 % (channels / 8)) == 0 && (x_ld % 8) == 0,
                 "group_norm_relu: channels must be 8 * (a divisor of 256)");
  GLSDET_REQUIRE(groups > 0 && groups <= 64 && (channels % groups) == 0, "group_norm_relu: bad group count");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int lanes = 256 / (channels / 8);
  size_t smem = static_cast<size_t>(lanes) * 2 * channels * sizeof(float);
  if (smem < static_cast<size_t>(2 * channels) * sizeof(double)) smem = static_cast<size_t>(2 * channels) * sizeof(double);
  GLSDET_REQUIRE(smem <= 48 * 1024, "group_norm_relu: too many channels for the reduction buffer");
  float* ab = scratch + static_cast<int64_t>(batch) * kGnSlabs * 2 * channels;
  int* counters = reinterpret_cast<int*>(ab + static_cast<int64_t>(batch) * 2 * channels);
  launch_pdl(gn_partial_kernel, dim3(kGnSlabs, batch), dim3(256), smem, st, reinterpret_cast<const __nv_bfloat16*>(x), scratch, hw,
             channels, x_ld, gamma, beta, ab, counters, groups, eps);
  if (int rc = count_launch("gn_partial_kernel")) return rc;
  const int pix_per_cta = 256 / (channels / 8) * 16;     // 16 iterations per thread
  const int slabs = (hw + pix_per_cta - 1) / pix_per_cta;
  launch_pdl(gn_apply_kernel, dim3(slabs < 1 ? 1 : (slabs > 1024 ? 1024 : slabs), batch), dim3(256), 0, st,
             reinterpret_cast<__nv_bfloat16*>(x), ab, hw, channels, x_ld);
  return count_launch("gn_apply_kernel");
}

extern "C" int64_t glsdet_group_norm_scratch_floats(int32_t batch, int32_t channels) {
  return static_cast<int64_t>(batch) * (kGnSlabs * 2 + 2) * channels + batch;   // partials, scale / shift, arrival counters
}

extern "C" int glsdet_proxy_scores(const float* feat, const float* centers, const int32_t* cls_start, int32_t num_classes,
                                   int32_t num_proxies, int32_t channels, int32_t batch, int32_t hw, float gamma, float* rows,
                                   int32_t rows_ld, int64_t rows_batch_stride, int32_t row0, void* stream) {
  GLSDET_REQUIRE(feat && centers && cls_start && rows && batch > 0 && hw > 0, "proxy_scores: bad arguments");
  GLSDET_REQUIRE(num_classes > 0 && num_classes <= 32 && num_proxies > 0 && num_proxies <= 64, "proxy_scores: at most 32 "
                 "classes and 64 proxies");
  const size_t smem = static_cast<size_t>(num_proxies) * channels * sizeof(float);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(proxy_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); attr = true; }
  GLSDET_REQUIRE(smem <= 160 * 1024, "proxy_scores: proxies do not fit shared memory");
  GLSDET_REQUIRE((channels % 32) == 0 && channels <= 512, "proxy_scores: channels must be a multiple of 32, at most 512");
  const int64_t total = static_cast<int64_t>(batch) * hw;
  const int64_t want = (total + 7) / 8;
  const int64_t cap_ctas = 4ll * device_sm_count();
  launch_pdl(proxy_scores_kernel, dim3(static_cast<unsigned>(want < cap_ctas ? want : cap_ctas)), dim3(256), smem,
             static_cast<cudaStream_t>(stream), feat, centers, cls_start, num_classes, num_proxies, channels, hw, total, gamma, rows,
             rows_ld, rows_batch_stride, row0);
  return count_launch("proxy_scores_kernel");
}

extern "C" int glsdet_proxy_aggregate(const void* feat, const float* sims, int32_t sims_ld, const int32_t* cls_start,
                                      int32_t num_classes, int32_t channels, int32_t batch, int32_t hw, float gamma, float* rows,
                                      int32_t rows_ld, int64_t rows_batch_stride, int32_t row0, void* stream) {
  GLSDET_REQUIRE(feat && sims && cls_start && rows && batch > 0 && hw > 0, "proxy_aggregate: bad arguments");
  GLSDET_REQUIRE(num_classes > 0 && num_classes <= 32 && channels > 0 && (channels % 8) == 0, "proxy_aggregate: at most 32 classes, "
                 "channels a multiple of 8");
  const int64_t total = static_cast<int64_t>(batch) * hw;
  const int64_t want = (total + 7) / 8;
  static int per_sm = 0;   // resident CTAs per SM: a persistent grid larger than that runs a second, mostly empty wave
  if (per_sm == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, proxy_aggregate_kernel, 256, 0) != cudaSuccess || n <= 0) n = 4;
    per_sm = n;
  }
  const int64_t cap_ctas = static_cast<int64_t>(per_sm) * device_sm_count();
  launch_pdl(proxy_aggregate_kernel, dim3(static_cast<unsigned>(want < cap_ctas ? want : cap_ctas)), dim3(256), 0,
             static_cast<cudaStream_t>(stream), reinterpret_cast<const __nv_bfloat16*>(feat), sims, sims_ld, cls_start, num_classes,
             channels, hw, total, gamma, rows, rows_ld, rows_batch_stride, row0);
  return count_launch("proxy_aggregate_kernel");
}

extern "C" int glsdet_gfl_decode(const float* reg, int32_t reg_ld, int32_t bins, int32_t batch, int32_t height, int32_t width,
                                 float stride, float max_x, float max_y, float* boxes, int64_t boxes_batch_stride, int32_t row0,
                                 void* stream) {
  GLSDET_REQUIRE(reg && boxes && batch > 0 && height > 0 && width > 0 && bins > 0 && reg_ld >= 4 * bins,
                 "gfl_decode: bad arguments");
  const int64_t total = static_cast<int64_t>(batch) * height * width * 4;
  launch_pdl(gfl_decode_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream),
             reg, reg_ld, bins, height, width, stride, max_x, max_y, total, boxes, boxes_batch_stride, row0);
  return count_launch("gfl_decode_kernel");
}

extern "C" int64_t glsdet_gfl_select_scratch_ints(int32_t batch) { return static_cast<int64_t>(batch) * kSelScratch; }

extern "C" int glsdet_gfl_select(const float* rows, int32_t rows_ld, int64_t rows_batch_stride, const float* boxes,
                                 int64_t boxes_batch_stride, int32_t row0, int32_t level_anchors, int32_t num_classes,
                                 float score_thr, int32_t topk, int32_t batch, void* keys, int64_t keys_batch_stride,
                                 int32_t* cand_count, float* cand_boxes, float* cand_scores, float* cand_labels,
                                 int32_t cand_capacity, int32_t* scratch, void* stream) {
  GLSDET_REQUIRE(rows && boxes && keys && cand_count && cand_boxes && cand_scores && cand_labels && scratch, "gfl_select: null pointer");
  GLSDET_REQUIRE(batch > 0 && level_anchors > 0 && num_classes > 0 && topk > 0 && cand_capacity > 0, "gfl_select: bad sizes");
  int64_t need = 1;
  while (need < static_cast<int64_t>(level_anchors) * num_classes) need <<= 1;
  GLSDET_REQUIRE(keys_batch_stride >= need, "gfl_select: key buffer too small (needs the next power of two of anchors * classes)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t total = static_cast<int64_t>(level_anchors) * num_classes;
  int chunks = static_cast<int>((total + 256 * 8 - 1) / (256 * 8));
  const int cap_chunks = 2 * device_sm_count() / batch + 1;
  if (chunks > cap_chunks) chunks = cap_chunks;
  if (chunks < 1) chunks = 1;
  int shift = 0;   // histogram bin = 2^shift ulps of the score, the range (threshold, 1.0] fits kSelBins bins
  {
    const float t = score_thr > 0.0f ? score_thr : 0.0f;
    uint32_t tb, one = 0x3F800000u;
    memcpy(&tb, &t, sizeof(tb));
    const uint32_t range = tb < one ? one - tb : 1u;
    while ((range >> shift) >= static_cast<uint32_t>(kSelBins)) ++shift;
  }
  launch_pdl(gfl_select_hist_kernel, dim3(chunks, batch), dim3(256), 0, st, rows, rows_ld, rows_batch_stride, row0, level_anchors,
             num_classes, score_thr, shift, scratch);
  if (int rc = count_launch("gfl_select_hist_kernel")) return rc;
  launch_pdl(gfl_select_collect_kernel, dim3(chunks, batch), dim3(256), 0, st, rows, rows_ld, rows_batch_stride, row0, level_anchors,
             num_classes, score_thr, shift, topk, scratch, reinterpret_cast<unsigned long long*>(keys), keys_batch_stride);
  if (int rc = count_launch("gfl_select_collect_kernel")) return rc;
  launch_pdl(gfl_select_sort_kernel, dim3(batch), dim3(1024), 0, st, boxes, boxes_batch_stride, row0, num_classes, score_thr, shift,
             topk, scratch,
             reinterpret_cast<unsigned long long*>(keys), keys_batch_stride, cand_count, cand_boxes, cand_scores, cand_labels,
             cand_capacity);
  return count_launch("gfl_select_sort_kernel");
}
