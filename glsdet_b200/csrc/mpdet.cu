// MP-Det head pieces that are not convolutions (yolox-ufp/mmdet/models/dense_heads/mp_head.py, gfl_head.py):
// GroupNorm + ReLU of the shared towers, the proxy classification, the integral box decode and the per-level
// candidate selection of _get_bboxes_single.  The convolutions (FPN, towers, gfl_cls_conv, gfl_reg) run on
// conv_gemm_kernel.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/glsdet_b200.h"
#include "common.h"
#include "ptx.cuh"

namespace glsdet {

constexpr int kGnSlabs = 16;

// ---------------------------------------------------------------------------------------------- GroupNorm + ReLU
// stage 1: per (image, slab of pixels) channel sums and sums of squares, fixed order (deterministic)
__global__ void __launch_bounds__(256) gn_partial_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ scratch,
                                                         int HW, int C, int ld) {
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
  extern __shared__ float red[];   // [lanes][2][C]
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
}

// stage 2: group statistics -> per (image, channel) scale a and shift s with  y = relu(x * a + s).
// One thread per channel sums its slab partials (16 independent loads; round 1 walked slabs x channels of a group
// serially in one thread: 256 dependent loads, 20 us per launch), then the channels of a group are added in fixed order.
__global__ void __launch_bounds__(256) gn_finalize_kernel(const float* __restrict__ scratch, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float* __restrict__ ab, int HW, int C,
                                                          int groups, float eps) {
  pdl_prologue();
  const int b = blockIdx.x;
  const int cpg = C / groups;
  extern __shared__ double gn_fin[];   // [2][C] channel sums
  __shared__ float mean_s[64], rstd_s[64];
  for (int c = threadIdx.x; c < C; c += 256) {
    double sm = 0.0, sq = 0.0;
#pragma unroll
    for (int k = 0; k < kGnSlabs; ++k) {
      sm += scratch[((static_cast<int64_t>(b) * kGnSlabs + k) * 2 + 0) * C + c];
      sq += scratch[((static_cast<int64_t>(b) * kGnSlabs + k) * 2 + 1) * C + c];
    }
    gn_fin[c] = sm;
    gn_fin[C + c] = sq;
  }
  __syncthreads();
  for (int g = threadIdx.x; g < groups; g += 256) {
    double sm = 0.0, sq = 0.0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) { sm += gn_fin[c]; sq += gn_fin[C + c]; }
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
}

// stage 3: in place  x = relu(x * a[b][c] + s[b][c])
__global__ void __launch_bounds__(256) gn_apply_kernel(__nv_bfloat16* __restrict__ x, const float* __restrict__ ab, int HW, int C,
                                                       int ld, int64_t total) {
  pdl_prologue();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int nvec = C >> 3;
  const int v = static_cast<int>(i % nvec);
  const int64_t pix = i / nvec;
  const int b = static_cast<int>(pix / HW);
  uint4* ptr = reinterpret_cast<uint4*>(x + pix * ld) + v;
  const uint4 u = *ptr;
  const float* a = ab + (static_cast<int64_t>(b) * 2) * C + v * 8;
  const float* s = a + C;
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float lo = fmaxf(fmaf(__uint_as_float(w[k] << 16), a[2 * k], s[2 * k]), 0.0f);
    const float hi = fmaxf(fmaf(__uint_as_float(w[k] & 0xFFFF0000u), a[2 * k + 1], s[2 * k + 1]), 0.0f);
    o[k] = pack_bf16x2(lo, hi);
  }
  *ptr = make_uint4(o[0], o[1], o[2], o[3]);
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
      const float* sp = sims + pix * sims_ld;
      const int j0 = cls_s[lane], j1 = cls_s[lane + 1];
      float mx = -1e30f;
      for (int j = j0; j < j1; ++j) mx = fmaxf(mx, sp[j] * inv * gamma);
      float den = 0.0f, num = 0.0f;
      for (int j = j0; j < j1; ++j) {
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
__global__ void __launch_bounds__(1024) gfl_select_kernel(const float* __restrict__ rows, int rows_ld, int64_t rows_bs,
                                                          const float* __restrict__ boxes, int64_t boxes_bs, int row0, int A_l,
                                                          int nc, float thr, int topk, unsigned long long* __restrict__ keys,
                                                          int64_t keys_bs, int* __restrict__ cand_count, float* __restrict__ cboxes,
                                                          float* __restrict__ cscores, float* __restrict__ clabels, int cap) {
  // Only the best `topk` of up to anchors x classes candidates are needed, so the full sort of round 1 (a single-CTA
  // bitonic network over up to 2^18 keys in global memory: 1.9 ms per level, half of an MP-Det step) is replaced by a
  // selection: pass 1 histograms the score bits, the bin b* in which the topk-th best score lies is found, pass 2 collects
  // the candidates of the bins >= b* (topk plus the rest of one bin), and only those are sorted - in shared memory.
  // Same keys as before ((~score bits) << 32 | flattened index), so the result is identical.
  pdl_prologue();
  const int b = blockIdx.x;
  __shared__ int hist[kSelBins];
  __shared__ unsigned long long sk[kSelCap];
  __shared__ int n_s, base_s, bstar_s, m_s;
  for (int i = threadIdx.x; i < kSelBins; i += blockDim.x) hist[i] = 0;
  if (threadIdx.x == 0) { n_s = 0; m_s = 0; }
  __syncthreads();
  unsigned long long* kb = keys + b * keys_bs;
  const float* rb = rows + b * rows_bs + static_cast<int64_t>(row0) * rows_ld;
  const int total = A_l * nc;
  const uint32_t thr_bits = __float_as_uint(fmaxf(thr, 0.0f));
  auto bin_of = [&](float sc) { return min(kSelBins - 1, static_cast<int>((__float_as_uint(sc) - thr_bits) >> 15)); };
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int a = i / nc, c = i - a * nc;
    const float sc = 1.0f / (1.0f + expf(-rb[static_cast<int64_t>(a) * rows_ld + c]));
    if (sc > thr) atomicAdd(&hist[bin_of(sc)], 1);
  }
  __syncthreads();
  if (threadIdx.x < 32) {   // one warp walks the histogram from the best bin down, 64 bins per step
    int cum = 0, bstar = 0, ntot = 0;
    bool found = false;
    for (int hi = kSelBins; hi > 0; hi -= 64) {
      const int b0 = hi - 1 - 2 * static_cast<int>(threadIdx.x);          // lanes own bins (b0, b0 - 1), descending
      const int h0 = hist[b0], h1 = hist[b0 - 1];
      int incl = h0 + h1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (static_cast<int>(threadIdx.x) >= o) incl += v;
      }
      const int before = cum + incl - h0 - h1;     // candidates in strictly better bins than b0
      if (!found) {
        const bool hit0 = before + h0 >= topk, hit1 = before + h0 + h1 >= topk;
        const unsigned m = __ballot_sync(0xffffffffu, hit0 || hit1);
        if (m) {
          const int l = __ffs(m) - 1;
          const int bs = __shfl_sync(0xffffffffu, hit0 ? b0 : b0 - 1, l);
          bstar = bs;
          found = true;
        }
      }
      cum += __shfl_sync(0xffffffffu, incl, 31);
    }
    ntot = cum;
    if (threadIdx.x == 0) { n_s = ntot; bstar_s = found ? bstar : 0; }
  }
  __syncthreads();
  const int n = n_s, bstar = bstar_s;
  const int take = min(n, topk);
  // pass 2: collect the candidates of the bins >= b*
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int a = i / nc, c = i - a * nc;
    const float sc = 1.0f / (1.0f + expf(-rb[static_cast<int64_t>(a) * rows_ld + c]));
    if (sc > thr && bin_of(sc) >= bstar) {
      const int pos = atomicAdd(&m_s, 1);
      // descending score, ascending index: key = (~score_bits << 32) | index, sorted ascending
      const unsigned long long key = (static_cast<unsigned long long>(~__float_as_uint(sc)) << 32) | static_cast<unsigned int>(i);
      if (pos < kSelCap) sk[pos] = key;
      kb[pos] = key;      // global copy: only read when the collected set does not fit shared memory
    }
  }
  __syncthreads();
  const int m = m_s;
  int P = 1;
  while (P < m) P <<= 1;
  const bool in_smem = m <= kSelCap;
  unsigned long long* sv = in_smem ? sk : kb;
  for (int i = m + threadIdx.x; i < P; i += blockDim.x) sv[i] = ~0ull;
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < P / 2; t += blockDim.x) {
        const int i = 2 * j * (t / j) + (t % j);
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
  const size_t smem = static_cast<size_t>(lanes) * 2 * channels * sizeof(float);
  GLSDET_REQUIRE(smem <= 48 * 1024, "group_norm_relu: too many channels for the reduction buffer");
  float* ab = scratch + static_cast<int64_t>(batch) * kGnSlabs * 2 * channels;
  launch_pdl(gn_partial_kernel, dim3(kGnSlabs, batch), dim3(256), smem, st, reinterpret_cast<const __nv_bfloat16*>(x), scratch, hw,
             channels, x_ld);
  if (int rc = count_launch("gn_partial_kernel")) return rc;
  launch_pdl(gn_finalize_kernel, dim3(batch), dim3(256), static_cast<size_t>(2 * channels) * sizeof(double), st, scratch, gamma, beta,
             ab, hw, channels, groups, eps);
  if (int rc = count_launch("gn_finalize_kernel")) return rc;
  const int64_t total = static_cast<int64_t>(batch) * hw * (channels / 8);
  launch_pdl(gn_apply_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, reinterpret_cast<__nv_bfloat16*>(x),
             ab, hw, channels, x_ld, total);
  return count_launch("gn_apply_kernel");
}

extern "C" int64_t glsdet_group_norm_scratch_floats(int32_t batch, int32_t channels) {
  return static_cast<int64_t>(batch) * (kGnSlabs * 2 + 2) * channels;
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
  const int64_t cap_ctas = 8ll * device_sm_count();
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

extern "C" int glsdet_gfl_select(const float* rows, int32_t rows_ld, int64_t rows_batch_stride, const float* boxes,
                                 int64_t boxes_batch_stride, int32_t row0, int32_t level_anchors, int32_t num_classes,
                                 float score_thr, int32_t topk, int32_t batch, void* keys, int64_t keys_batch_stride,
                                 int32_t* cand_count, float* cand_boxes, float* cand_scores, float* cand_labels,
                                 int32_t cand_capacity, void* stream) {
  GLSDET_REQUIRE(rows && boxes && keys && cand_count && cand_boxes && cand_scores && cand_labels, "gfl_select: null pointer");
  GLSDET_REQUIRE(batch > 0 && level_anchors > 0 && num_classes > 0 && topk > 0 && cand_capacity > 0, "gfl_select: bad sizes");
  int64_t need = 1;
  while (need < static_cast<int64_t>(level_anchors) * num_classes) need <<= 1;
  GLSDET_REQUIRE(keys_batch_stride >= need, "gfl_select: key buffer too small (needs the next power of two of anchors * classes)");
  launch_pdl(gfl_select_kernel, dim3(batch), dim3(1024), 0, static_cast<cudaStream_t>(stream), rows, rows_ld, rows_batch_stride, boxes,
             boxes_batch_stride, row0, level_anchors, num_classes, score_thr, topk, reinterpret_cast<unsigned long long*>(keys),
             keys_batch_stride, cand_count, cand_boxes, cand_scores, cand_labels, cand_capacity);
  return count_launch("gfl_select_kernel");
}
