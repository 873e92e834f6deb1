// UFP stage of UFPMP-Det: the step between the coarse detector and the MP-Det pass.
//
//   host   glsdet_ufp_pack      scale_boxes + ForegroundRegionGeneration + Packing (strip packing, spp.py) - strictly
//                               sequential, order-dependent algorithms over a few hundred boxes: C++ on the host
//                               (yolox-ufp/mmdet/core/ufp/unified_foreground_packing.py:6-197, spp.py:69-168)
//   device glsdet_ufp_mosaic    crop + integer-factor bilinear resize (cv2.resize INTER_LINEAR, bit-exact) + paste of every
//                               chip into the mosaic canvas (ufpmp_det_eval.py:182-193)
//   device glsdet_ufp_merge     map the second-stage detections back through the chips (intersection over the smaller
//                               area > 0.9, ufpmp_det_eval.py:270-296) and merge them per class with the legacy "+1" NMS
//                               (py_cpu_nms :149-179, threshold 0.6 at :306), one CTA per class
//
// Arithmetic follows the reference's NumPy types: float32 boxes, Python ints for chip rectangles, float64 for the packing.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>
#include <array>
#include <numeric>
#include <vector>

#include "../../include/glsdet_b200.h"
#include "common.h"

namespace glsdet {
namespace {

// ------------------------------------------------------------------------------------------------ host: packing
struct Rect { double x, y, w, h; bool set; };

// spp.py:115-168, D = 0 (rotation is not allowed in phsppog)
void recursive_packing(double x, double y, double w, double h, const std::vector<std::array<double, 2>>& rem,
                       std::vector<int>& indices, std::vector<Rect>& result) {
  int priority = 6, best = -1;
  for (int idx : indices) {
    const double rw = rem[idx][0], rh = rem[idx][1];
    if (priority > 1 && rw == w && rh == h) { priority = 1; best = idx; break; }
    else if (priority > 2 && rw == w && rh < h) { priority = 2; best = idx; }
    else if (priority > 3 && rw < w && rh == h) { priority = 3; best = idx; }
    else if (priority > 4 && rw < w && rh < h) { priority = 4; best = idx; }
    else if (priority > 5) { priority = 5; best = idx; }
  }
  if (priority >= 5) return;
  const double omega = rem[best][0], d = rem[best][1];
  result[best] = {x, y, omega, d, true};
  indices.erase(std::find(indices.begin(), indices.end(), best));
  if (priority == 2) {
    recursive_packing(x, y + d, w, h - d, rem, indices, result);
  } else if (priority == 3) {
    recursive_packing(x + omega, y, w - omega, h, rem, indices, result);
  } else if (priority == 4) {
    double min_w = INFINITY, min_h = INFINITY;
    for (int idx : indices) { min_w = std::min(min_w, rem[idx][0]); min_h = std::min(min_h, rem[idx][1]); }
    min_w = std::min(min_h, min_w);
    min_h = min_w;
    if (w - omega < min_w) {
      recursive_packing(x, y + d, w, h - d, rem, indices, result);
    } else if (h - d < min_h) {
      recursive_packing(x + omega, y, w - omega, h, rem, indices, result);
    } else if (omega < min_w) {
      recursive_packing(x + omega, y, w - omega, d, rem, indices, result);
      recursive_packing(x, y + d, w, h - d, rem, indices, result);
    } else {
      recursive_packing(x, y + d, omega, h - d, rem, indices, result);
      recursive_packing(x + omega, y, w - omega, h, rem, indices, result);
    }
  }
}

// spp.py:69-112 with sorting = 'height' (stable, descending)
double phsppog_height(double width, const std::vector<std::array<double, 2>>& rects, std::vector<Rect>& result) {
  const int n = static_cast<int>(rects.size());
  result.assign(n, Rect{0, 0, 0, 0, false});
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return -rects[a][1] < -rects[b][1]; });
  double H = 0;
  while (!order.empty()) {
    const int idx = order.front();
    order.erase(order.begin());
    const double r0 = rects[idx][0], r1 = rects[idx][1];
    result[idx] = {0, H, r0, r1, true};
    const double x = r0, y = H, w = width - r0, h = r1;
    H = H + r1;
    recursive_packing(x, y, w, h, rects, order, result);
  }
  return H;
}

// ------------------------------------------------------------------------------------------------ device: mosaic
struct Chip { int32_t x1, y1, w, h, nx, ny, sf; };

// cv2.resize(crop, (w * sf, h * sf)) for uint8, INTER_LINEAR (OpenCV imgproc/resize.cpp): 11-bit coefficients; columns
// clamp the COEFFICIENT at the borders (sx < 0 -> fx = 0, sx = 0; sx >= w - 1 -> fx = 0, sx = w - 1), rows clamp the source
// ROW and keep the coefficient; dst = ((b0 * (S0 >> 4) >> 16) + (b1 * (S1 >> 4) >> 16) + 2) >> 2.
__global__ void __launch_bounds__(256) ufp_mosaic_kernel(const uint8_t* __restrict__ img, int img_h, int img_w,
                                                         const Chip* __restrict__ chips, uint8_t* __restrict__ canvas,
                                                         int can_h, int can_w) {
  pdl_prologue();
  const Chip c = chips[blockIdx.y];
  if (c.w <= 0 || c.h <= 0) return;
  const int ow = c.w * c.sf, oh = c.h * c.sf;
  const int total = ow * oh;
  const int two_sf = 2 * c.sf;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int oy = i / ow, ox = i - oy * ow;
    const int cy = c.ny + oy, cx = c.nx + ox;
    if (cy < 0 || cy >= can_h || cx < 0 || cx >= can_w) continue;
    // source coordinate f = (o + 0.5) / sf - 0.5 = (2 o + 1 - sf) / (2 sf): floor and fraction in integers
    int nx = 2 * ox + 1 - c.sf, ny = 2 * oy + 1 - c.sf;
    int sx = (nx >= 0) ? nx / two_sf : -((-nx + two_sf - 1) / two_sf);
    int sy = (ny >= 0) ? ny / two_sf : -((-ny + two_sf - 1) / two_sf);
    int ax = ((nx - sx * two_sf) * 2048) / two_sf;   // exact for sf in {1, 2, 4}
    const int ay = ((ny - sy * two_sf) * 2048) / two_sf;
    if (sx < 0) { sx = 0; ax = 0; }
    if (sx >= c.w - 1) { sx = c.w - 1; ax = 0; }
    const int sx1 = min(sx + 1, c.w - 1);
    const int sy0 = min(max(sy, 0), c.h - 1), sy1 = min(max(sy + 1, 0), c.h - 1);
    const uint8_t* r0 = img + (static_cast<int64_t>(c.y1 + sy0) * img_w + c.x1) * 3;
    const uint8_t* r1 = img + (static_cast<int64_t>(c.y1 + sy1) * img_w + c.x1) * 3;
    uint8_t* o = canvas + (static_cast<int64_t>(cy) * can_w + cx) * 3;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const int s0 = r0[sx * 3 + ch] * (2048 - ax) + r0[sx1 * 3 + ch] * ax;
      const int s1 = r1[sx * 3 + ch] * (2048 - ax) + r1[sx1 * 3 + ch] * ax;
      o[ch] = static_cast<uint8_t>(((((2048 - ay) * (s0 >> 4)) >> 16) + ((ay * (s1 >> 4)) >> 16) + 2) >> 2);
    }
  }
}

// ------------------------------------------------------------------------------------------------ device: merge
// A NumPy scalar of the reference's map-back arithmetic: an exact integer (Python int) or a float32.
struct Num {
  double v;
  bool is_int;
};
__device__ __forceinline__ Num num_i(int x) { return Num{static_cast<double>(x), true}; }
__device__ __forceinline__ Num num_f(float x) { return Num{static_cast<double>(x), false}; }
__device__ __forceinline__ Num num_sub(Num a, Num b) {
  if (a.is_int && b.is_int) return Num{a.v - b.v, true};
  return num_f(__fsub_rn(static_cast<float>(a.v), static_cast<float>(b.v)));
}
__device__ __forceinline__ Num num_mul(Num a, Num b) {
  if (a.is_int && b.is_int) return Num{a.v * b.v, true};
  return num_f(__fmul_rn(static_cast<float>(a.v), static_cast<float>(b.v)));
}
__device__ __forceinline__ Num num_max(Num a, Num b) { return (b.v > a.v) ? b : a; }   // Python max: first maximal element
__device__ __forceinline__ Num num_min(Num a, Num b) { return (b.v < a.v) ? b : a; }

// ufpmp_det_eval.py:36-50 with pos1 = float32 detection, pos2 = integer chip rectangle; returns iof > 0.9
__device__ bool iof_above(float l1, float t1, float r1, float d1, int l2, int t2, int r2, int d2) {
  const Num area1 = num_mul(num_sub(num_f(r1), num_f(l1)), num_sub(num_f(d1), num_f(t1)));
  const Num area2 = num_i((r2 - l2) * (d2 - t2));
  const Num left = num_max(num_f(l1), num_i(l2)), right = num_min(num_f(r1), num_i(r2));
  const Num top = num_max(num_f(t1), num_i(t2)), bottom = num_min(num_f(d1), num_i(d2));
  if (left.v >= right.v || top.v >= bottom.v) return false;
  const Num inter = num_mul(num_sub(right, left), num_sub(bottom, top));
  const Num den = num_min(area1, area2);
  if (inter.is_int && den.is_int) return inter.v / den.v > 0.9;                      // Python float division
  return __fdiv_rn(static_cast<float>(inter.v), static_cast<float>(den.v)) > 0.9f;   // float32 result vs weak 0.9
}

constexpr int kMergeThreads = 256;
constexpr int kMergeCap = 4096;   // mapped detections per class held in shared memory (keys) for the sort

// One CTA per class: ordered map-back (chip-major, then detection order), bitonic sort by (score desc, index asc),
// greedy NMS with the legacy +1 areas (keep while ovr <= thresh), rows out in score order.
__global__ void __launch_bounds__(kMergeThreads) ufp_merge_kernel(const float* __restrict__ dets, const int32_t* __restrict__ cls_off,
                                                                  const Chip* __restrict__ chips, int n_chips, float thresh,
                                                                  float* __restrict__ mapped, int cap, float* __restrict__ out,
                                                                  int32_t* __restrict__ out_count, int32_t* __restrict__ mapped_count) {
  pdl_prologue();
  __shared__ unsigned long long keys[kMergeCap];
  __shared__ uint8_t dead[kMergeCap];
  __shared__ int s_count, s_warp[kMergeThreads / 32], s_kept;
  __shared__ float s_box[5];
  const int c = blockIdx.x;
  const int d0 = cls_off[c], d1 = cls_off[c + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* mp = mapped + static_cast<int64_t>(c) * cap * 5;
  if (tid == 0) s_count = 0;
  __syncthreads();
  // ---- map-back, ordered compaction
  for (int ci = 0; ci < n_chips; ++ci) {
    const Chip ch = chips[ci];
    const int l2 = ch.nx, t2 = ch.ny, r2 = ch.nx + ch.w * ch.sf, b2 = ch.ny + ch.h * ch.sf;
    for (int base = d0; base < d1; base += kMergeThreads) {
      const int i = base + tid;
      bool hit = false;
      float x1 = 0, y1 = 0, x2 = 0, y2 = 0, sc = 0;
      if (i < d1) {
        x1 = dets[i * 5]; y1 = dets[i * 5 + 1]; x2 = dets[i * 5 + 2]; y2 = dets[i * 5 + 3]; sc = dets[i * 5 + 4];
        hit = iof_above(x1, y1, x2, y2, l2, t2, r2, b2);
      }
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (lane == 0) s_warp[warp] = __popc(m);
      __syncthreads();
      int off = s_count;
      for (int w = 0; w < warp; ++w) off += s_warp[w];
      if (hit) {
        const int slot = off + __popc(m & ((1u << lane) - 1u));
        if (slot < cap) {
          const float sf = static_cast<float>(ch.sf);
          const float nw = __fdiv_rn(__fsub_rn(x2, x1), sf), nh = __fdiv_rn(__fsub_rn(y2, y1), sf);
          const float bx = __fadd_rn(__fdiv_rn(__fsub_rn(x1, static_cast<float>(ch.nx)), sf), static_cast<float>(ch.x1));
          const float by = __fadd_rn(__fdiv_rn(__fsub_rn(y1, static_cast<float>(ch.ny)), sf), static_cast<float>(ch.y1));
          float* o = mp + static_cast<int64_t>(slot) * 5;
          o[0] = bx; o[1] = by; o[2] = __fadd_rn(bx, nw); o[3] = __fadd_rn(by, nh); o[4] = sc;
        }
      }
      __syncthreads();
      if (tid == 0) {
        int tot = 0;
        for (int w = 0; w < kMergeThreads / 32; ++w) tot += s_warp[w];
        s_count += tot;
      }
      __syncthreads();
    }
  }
  const int n_all = s_count;
  if (tid == 0) mapped_count[c] = n_all;
  const int n = min(n_all, min(cap, kMergeCap));
  __threadfence_block();
  __syncthreads();
  // ---- sort: key = ~order(score) << 32 | index  (ascending keys = descending score, ties by the lower index)
  int npow = 1;
  while (npow < n) npow <<= 1;
  for (int i = tid; i < npow; i += kMergeThreads) {
    unsigned long long k = ~0ull;
    if (i < n) {
      const uint32_t b = __float_as_uint(mp[i * 5 + 4]);
      const uint32_t ord = (b & 0x80000000u) ? ~b : (b | 0x80000000u);   // total order of floats
      k = (static_cast<unsigned long long>(~ord) << 32) | static_cast<unsigned>(i);
    }
    keys[i] = k;
    if (i < kMergeCap) dead[i] = 0;
  }
  __syncthreads();
  for (int k = 2; k <= npow; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < npow; i += kMergeThreads) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], b = keys[ixj];
          const bool up = ((i & k) == 0);
          if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  // ---- greedy NMS over the sorted order (py_cpu_nms: areas with +1, suppressed when ovr > thresh)
  if (tid == 0) s_kept = 0;
  __syncthreads();
  float* oc = out + static_cast<int64_t>(c) * cap * 5;
  for (int i = 0; i < n; ++i) {
    if (dead[i]) continue;          // uniform: dead[] is only written before a barrier
    const int bi = static_cast<int>(keys[i] & 0xffffffffu);
    if (tid < 5) s_box[tid] = mp[bi * 5 + tid];
    __syncthreads();
    const float ax1 = s_box[0], ay1 = s_box[1], ax2 = s_box[2], ay2 = s_box[3];
    const float area_a = __fmul_rn(__fadd_rn(__fsub_rn(ax2, ax1), 1.0f), __fadd_rn(__fsub_rn(ay2, ay1), 1.0f));
    if (tid == 0) {
      float* o = oc + static_cast<int64_t>(s_kept) * 5;
      o[0] = ax1; o[1] = ay1; o[2] = ax2; o[3] = ay2; o[4] = s_box[4];
      s_kept += 1;
    }
    for (int j = i + 1 + tid; j < n; j += kMergeThreads) {
      if (dead[j]) continue;
      const int bj = static_cast<int>(keys[j] & 0xffffffffu);
      const float bx1 = mp[bj * 5], by1 = mp[bj * 5 + 1], bx2 = mp[bj * 5 + 2], by2 = mp[bj * 5 + 3];
      const float area_b = __fmul_rn(__fadd_rn(__fsub_rn(bx2, bx1), 1.0f), __fadd_rn(__fsub_rn(by2, by1), 1.0f));
      const float w = fmaxf(0.0f, __fadd_rn(__fsub_rn(fminf(ax2, bx2), fmaxf(ax1, bx1)), 1.0f));
      const float h = fmaxf(0.0f, __fadd_rn(__fsub_rn(fminf(ay2, by2), fmaxf(ay1, by1)), 1.0f));
      const float inter = __fmul_rn(w, h);
      const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
      if (!(ovr <= thresh)) dead[j] = 1;
    }
    __syncthreads();
  }
  if (tid == 0) out_count[c] = s_kept;
}

}  // namespace
}  // namespace glsdet

using namespace glsdet;

extern "C" int glsdet_ufp_pack(const float* boxes, int32_t n, float scale, int32_t in_w, int32_t in_h, double* rows,
                               int32_t* n_rows, double* new_w, double* new_h) {
  GLSDET_REQUIRE(n >= 0 && (n == 0 || boxes != nullptr) && rows && n_rows && new_w && new_h, "ufp_pack: bad arguments");
  // ---- scale_boxes (unified_foreground_packing.py:6-32), float32
  std::vector<std::array<float, 4>> sc(n);
  std::vector<float> avg(n);
  std::vector<int64_t> cnt(n, 1);
  std::vector<char> used(n, 1);
  auto clipf = [](float v, float lo, float hi) { return std::min(std::max(v, lo), hi); };
  for (int i = 0; i < n; ++i) {
    const float x0 = boxes[i * 4], y0 = boxes[i * 4 + 1], x1 = boxes[i * 4 + 2], y1 = boxes[i * 4 + 3];
    float w_half = (x1 - x0) * 0.5f, h_half = (y1 - y0) * 0.5f;
    const float xc = (x1 + x0) * 0.5f, yc = (y1 + y0) * 0.5f;
    w_half *= scale;
    h_half *= scale;
    sc[i] = {clipf(xc - w_half, 0.0f, static_cast<float>(in_w - 1)), clipf(yc - h_half, 0.0f, static_cast<float>(in_h - 1)),
             clipf(xc + w_half, 0.0f, static_cast<float>(in_w - 1)), clipf(yc + h_half, 0.0f, static_cast<float>(in_h - 1))};
    avg[i] = (x1 - x0 + 1.0f) * (y1 - y0 + 1.0f);
  }
  // ---- ForegroundRegionGeneration (:68-103): greedy, order-dependent merge of the scaled boxes
  for (int i = 0; i < n; ++i) {
    if (!used[i]) continue;
    std::array<float, 4> a = sc[i];
    for (int j = 0; j < n; ++j) {
      if (!used[j] || i == j) continue;
      const std::array<float, 4>& b = sc[j];
      const float a1 = (a[2] - a[0]) * (a[3] - a[1]);
      const float a2 = (b[2] - b[0]) * (b[3] - b[1]);
      const float mx0 = std::min(a[0], b[0]), my0 = std::min(a[1], b[1]);
      const float mx1 = std::max(a[2], b[2]), my1 = std::max(a[3], b[3]);
      const float merge = (mx1 - mx0) * (my1 - my0);
      const float origin = a1 + a2;
      if (merge < origin) {
        a = {mx0, my0, mx1, my1};
        used[j] = 0;
        avg[i] = avg[i] + avg[j];
        cnt[i] += cnt[j];
      }
    }
    sc[i] = a;
  }
  std::vector<std::array<float, 4>> regions;
  std::vector<int> factor;
  for (int i = 0; i < n; ++i) {
    if (!used[i]) continue;
    const double mean = static_cast<double>(avg[i]) / static_cast<double>(cnt[i]);   // float32 / int64 -> float64
    factor.push_back(mean < 32.0 * 32.0 ? 4 : mean < 96.0 * 96.0 ? 2 : 1);
    regions.push_back(sc[i]);
  }
  // ---- Packing (:140-181): binary search of the strip width; the layout of the LAST probe is the one used
  const int m = static_cast<int>(regions.size());
  std::vector<std::array<double, 2>> rects(m);
  for (int i = 0; i < m; ++i) {
    const float w = regions[i][2] - regions[i][0], h = regions[i][3] - regions[i][1];
    rects[i] = {static_cast<double>(w) * factor[i], static_cast<double>(h) * factor[i]};
  }
  double lo = 300, hi = 2666;
  std::vector<Rect> layout;
  while (lo <= hi) {
    const double mid = (lo + hi) / 2;
    const double height = phsppog_height(mid, rects, layout);
    if (height > mid) lo = mid + 1;
    else hi = mid - 1;
  }
  std::vector<char> flag(m, 1);
  int out_n = 0;
  double nw = 0, nh = 0;
  for (const Rect& r : layout) {
    nw = std::max(nw, r.x + r.w);
    nh = std::max(nh, r.y + r.h);
    for (int i = 0; i < m; ++i) {
      if (!flag[i]) continue;
      const float w = regions[i][2] - regions[i][0], h = regions[i][3] - regions[i][1];
      if (static_cast<double>(w) * factor[i] == r.w && static_cast<double>(h) * factor[i] == r.h) {
        flag[i] = 0;
        double* o = rows + static_cast<int64_t>(out_n) * 7;
        o[0] = regions[i][0]; o[1] = regions[i][1]; o[2] = w; o[3] = h; o[4] = r.x; o[5] = r.y; o[6] = factor[i];
        ++out_n;
      }
    }
  }
  *n_rows = out_n;
  *new_w = nw;
  *new_h = nh;
  return 0;
}

extern "C" int glsdet_ufp_mosaic(const uint8_t* image, int32_t img_h, int32_t img_w, const int32_t* chips, int32_t n_chips,
                                 uint8_t* canvas, int32_t can_h, int32_t can_w, void* stream) {
  GLSDET_REQUIRE(image && canvas && img_h > 0 && img_w > 0 && can_h >= 0 && can_w >= 0 && n_chips >= 0, "ufp_mosaic: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GLSDET_CHECK_CUDA(cudaMemsetAsync(canvas, 0, static_cast<size_t>(can_h) * can_w * 3, st));
  if (n_chips == 0 || can_h == 0 || can_w == 0) return 0;
  GLSDET_REQUIRE(chips != nullptr, "ufp_mosaic: null chip list");
  const int gx = 2 * device_sm_count() / (n_chips < 8 ? 1 : (n_chips < 32 ? 4 : 8));
  launch_pdl(ufp_mosaic_kernel, dim3(gx > 0 ? gx : 1, n_chips), dim3(256), 0, st, image, img_h, img_w,
             reinterpret_cast<const Chip*>(chips), canvas, can_h, can_w);
  return count_launch("ufp_mosaic_kernel");
}

extern "C" int glsdet_ufp_merge(const float* dets, const int32_t* cls_off, int32_t num_classes, const int32_t* chips,
                                int32_t n_chips, float nms_thresh, float* mapped, int32_t cap, float* out,
                                int32_t* out_count, int32_t* mapped_count, void* stream) {
  GLSDET_REQUIRE(cls_off && mapped && out && out_count && mapped_count && num_classes > 0, "ufp_merge: bad arguments");
  GLSDET_REQUIRE(cap > 0 && cap <= kMergeCap, "ufp_merge: cap must be in 1..%d mapped detections per class", kMergeCap);
  GLSDET_REQUIRE(n_chips == 0 || chips != nullptr, "ufp_merge: null chip list");
  launch_pdl(ufp_merge_kernel, dim3(num_classes), dim3(kMergeThreads), 0, static_cast<cudaStream_t>(stream), dets, cls_off,
             reinterpret_cast<const Chip*>(chips), n_chips, nms_thresh, mapped, cap, out, out_count, mapped_count);
  return count_launch("ufp_merge_kernel");
}
