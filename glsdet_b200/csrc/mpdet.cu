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

// stage 2: group statistics -> per (image, channel) scale a and shift s with  y = relu(x * a + s)
__global__ void __launch_bounds__(256) gn_finalize_kernel(const float* __restrict__ scratch, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float* __restrict__ ab, int HW, int C,
                                                          int groups, float eps) {
  const int b = blockIdx.x;
  const int cpg = C / groups;
  __shared__ float mean_s[64], rstd_s[64];
  for (int g = threadIdx.x; g < groups; g += 256) {
    double s = 0.0, q = 0.0;
    for (int k = 0; k < kGnSlabs; ++k)
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
        s += scratch[((static_cast<int64_t>(b) * kGnSlabs + k) * 2 + 0) * C + c];
        q += scratch[((static_cast<int64_t>(b) * kGnSlabs + k) * 2 + 1) * C + c];
      }
    const double n = static_cast<double>(HW) * cpg;
    const double m = s / n;
    const double var = fmax(q / n - m * m, 0.0);
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
  extern __shared__ float sc[];   // centers [n_prox][C]
  for (int i = threadIdx.x; i < n_prox * C; i += blockDim.x) sc[i] = centers[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t pix = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + warp;
  if (pix >= total_pix) return;
  const float* f = feat + pix * C;
  float nn = 0.0f;
  for (int c = lane; c < C; c += 32) { const float v = __ldg(f + c); nn += v * v; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
  const float inv = 1.0f / fmaxf(sqrtf(nn), 1e-12f);   // F.normalize eps
  float my_sim = 0.0f;                                  // lane j keeps the similarity of proxies j and j + 32
  float my_sim2 = 0.0f;
  for (int j = 0; j < n_prox; ++j) {
    float d = 0.0f;
    const float* cj = sc + j * C;
    for (int c = lane; c < C; c += 32) d += __ldg(f + c) * cj[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    d *= inv;
    if (j == lane) my_sim = d;
    if (j == lane + 32) my_sim2 = d;
  }
  // class aggregation: lane c handles class c
  float out = 0.0f;
  for (int c = 0; c < nc; ++c) {
    const int j0 = cls_start[c], j1 = cls_start[c + 1];
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

// ---------------------------------------------------------------------------------------------- integral box decode
// gfl_head.py:35-49,437-438,456-457: reg fp32 NHWC [P, reg_ld] holds 4 x (reg_max + 1) logits per pixel (Scale already
// folded into the conv); distance = softmax-expectation over the bins * stride; box = point -/+ distance, clamped.
__global__ void __launch_bounds__(256) gfl_decode_kernel(const float* __restrict__ reg, int reg_ld, int bins, int H, int W,
                                                         float stride, float max_x, float max_y, int64_t total,
                                                         float* __restrict__ boxes, int64_t boxes_bs, int row0) {
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
__global__ void __launch_bounds__(1024) gfl_select_kernel(const float* __restrict__ rows, int rows_ld, int64_t rows_bs,
                                                          const float* __restrict__ boxes, int64_t boxes_bs, int row0, int A_l,
                                                          int nc, float thr, int topk, unsigned long long* __restrict__ keys,
                                                          int64_t keys_bs, int* __restrict__ cand_count, float* __restrict__ cboxes,
                                                          float* __restrict__ cscores, float* __restrict__ clabels, int cap) {
  const int b = blockIdx.x;
  __shared__ int n_s, base_s;
  if (threadIdx.x == 0) n_s = 0;
  __syncthreads();
  unsigned long long* kb = keys + b * keys_bs;
  const float* rb = rows + b * rows_bs + static_cast<int64_t>(row0) * rows_ld;
  const int total = A_l * nc;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int a = i / nc, c = i - a * nc;
    const float s = 1.0f / (1.0f + expf(-rb[static_cast<int64_t>(a) * rows_ld + c]));
    if (s > thr) {
      const int pos = atomicAdd(&n_s, 1);
      // descending score, ascending index: key = (~score_bits << 32) | index, sorted ascending
      kb[pos] = (static_cast<unsigned long long>(~__float_as_uint(s)) << 32) | static_cast<unsigned int>(i);
    }
  }
  __syncthreads();
  const int n = n_s;
  int P = 1;
  while (P < n) P <<= 1;
  for (int i = n + threadIdx.x; i < P; i += blockDim.x) kb[i] = ~0ull;
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < P / 2; t += blockDim.x) {
        const int i = 2 * j * (t / j) + (t % j);
        const bool asc = ((i & k) == 0);
        const unsigned long long x = kb[i], y = kb[i + j];
        if ((x > y) == asc) { kb[i] = y; kb[i + j] = x; }
      }
      __syncthreads();
    }
  }
  const int take = min(n, topk);
  if (threadIdx.x == 0) base_s = atomicAdd(&cand_count[b], take);
  __syncthreads();
  const int base = base_s;
  for (int t = threadIdx.x; t < take; t += blockDim.x) {
    if (base + t >= cap) continue;
    const unsigned long long key = kb[t];
    const unsigned int idx = static_cast<unsigned int>(key & 0xFFFFFFFFull);
    const int a = idx / nc, c = idx - a * nc;
    const float s = __uint_as_float(~static_cast<unsigned int>(key >> 32));
    const float4 bx = *reinterpret_cast<const float4*>(boxes + b * boxes_bs + (static_cast<int64_t>(row0) + a) * 4);
    *reinterpret_cast<float4*>(cboxes + (static_cast<int64_t>(b) * cap + base + t) * 4) = bx;
    cscores[static_cast<int64_t>(b) * cap + base + t] = s;
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
  gn_partial_kernel<<<dim3(kGnSlabs, batch), 256, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), scratch, hw, channels, x_ld);
  if (int rc = count_launch("gn_partial_kernel")) return rc;
  gn_finalize_kernel<<<batch, 256, 0, st>>>(scratch, gamma, beta, ab, hw, channels, groups, eps);
  if (int rc = count_launch("gn_finalize_kernel")) return rc;
  const int64_t total = static_cast<int64_t>(batch) * hw * (channels / 8);
  gn_apply_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(x), ab, hw, channels,
                                                                              x_ld, total);
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
  const int64_t total = static_cast<int64_t>(batch) * hw;
  proxy_scores_kernel<<<static_cast<unsigned>((total + 7) / 8), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      feat, centers, cls_start, num_classes, num_proxies, channels, hw, total, gamma, rows, rows_ld, rows_batch_stride, row0);
  return count_launch("proxy_scores_kernel");
}

extern "C" int glsdet_gfl_decode(const float* reg, int32_t reg_ld, int32_t bins, int32_t batch, int32_t height, int32_t width,
                                 float stride, float max_x, float max_y, float* boxes, int64_t boxes_batch_stride, int32_t row0,
                                 void* stream) {
  GLSDET_REQUIRE(reg && boxes && batch > 0 && height > 0 && width > 0 && bins > 0 && reg_ld >= 4 * bins,
                 "gfl_decode: bad arguments");
  const int64_t total = static_cast<int64_t>(batch) * height * width * 4;
  gfl_decode_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
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
  gfl_select_kernel<<<batch, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      rows, rows_ld, rows_batch_stride, boxes, boxes_batch_stride, row0, level_anchors, num_classes, score_thr, topk,
      reinterpret_cast<unsigned long long*>(keys), keys_batch_stride, cand_count, cand_boxes, cand_scores, cand_labels,
      cand_capacity);
  return count_launch("gfl_select_kernel");
}
