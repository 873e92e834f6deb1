// Post-processing of the GLSDet path on the device, without host synchronisation:
//   score filter (obj * max cls >= conf)  ->  warp-ballot compaction  ->  per-image bitonic sort of
//   (class, score desc, anchor) keys  ->  per-(image, class) greedy NMS  ->  rank of every kept box among all
//   kept boxes of its image (binary searches)  ->  [K,7] rows, score desc.
//
// NMS proper, per (image, class) segment of n score-sorted boxes:
//   1. mask_tiles_kernel  - all 64x64 IoU tiles of the upper triangle, one 64-bit suppression word per (row,
//      column tile), computed by the whole GPU (persistent CTAs pulling tiles from a device-side work list);
//   2. scan_kernel        - one CTA per segment walks the 64-box chunks in order: the diagonal word resolves the
//      chunk greedily, the kept rows' words are OR-ed into the 'removed' bitmap of the later chunks.
//   Segments whose mask does not fit the workspace budget fall back to nms_segment_kernel (blocked greedy against
//   the kept list, no mask memory).  Both are the same greedy algorithm and give identical keep sets.
//
// Replaces non_max_suppression (yolox-drone/models/core/utils_bbox.py:375-484) and the torchvision
// batched_nms it calls (:414-419).  Bit-exactness contract (see include/glsdet_b200.h): same keep set and
// order as the chosen torchvision strategy, all box arithmetic in IEEE binary32 without FMA contraction.
//
// Greedy NMS decomposes exactly by class when boxes of different classes cannot overlap:
//   - per-class strategy: by definition;
//   - coordinate trick: boxes are shifted by label * (max_coord + 1); classes are disjoint whenever the
//     smallest coordinate is >= -0.5 (far inside the guard band of max_coord + 1).  Otherwise the image
//     falls back to one class-agnostic segment on the shifted boxes, which is the literal algorithm.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <new>

#include "../../include/glsdet_b200.h"
#include "common.h"

namespace glsdet {

constexpr int kSortChunk = 4096;    // keys sorted per CTA in shared memory
constexpr int kSortThreads = 512;
constexpr int kNmsThreads = 1024;   // 8 warps per scheduler: the kept-list sweep of a chunk is latency-bound
constexpr int kKeptSmem = 1536;     // kept boxes cached in shared memory per segment
constexpr int kMaxClasses = 256;    // 8 label bits in the sort key
constexpr uint64_t kPadKey = ~0ull;

// key = [63:56] label (0 when the image runs class-agnostic) | [55:24] ~score bits | [23:0] anchor index
__device__ __forceinline__ uint64_t make_key(uint32_t label, float score, uint32_t idx) {
  return (static_cast<uint64_t>(label) << 56) | (static_cast<uint64_t>(~__float_as_uint(score)) << 24) | idx;
}
__device__ __forceinline__ uint32_t key_idx(uint64_t k) { return static_cast<uint32_t>(k & 0xFFFFFFu); }
__device__ __forceinline__ float key_score(uint64_t k) { return __uint_as_float(~static_cast<uint32_t>(k >> 24)); }

__device__ __forceinline__ uint32_t float_order_bits(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float order_bits_float(uint32_t e) {
  return __uint_as_float((e & 0x80000000u) ? (e & 0x7FFFFFFFu) : ~e);
}

struct Source {
  // decoded predictions [B][A][nch] (cx, cy, w, h, obj, cls...), or caller-supplied boxes/scores/labels
  const float* pred;
  int A, nch, nc;
  int cls_logits;        // class columns hold raw logits: sigmoid is applied here (GLSDET_PRED_CLS_LOGITS)
  int64_t sa, sc;        // element strides of the anchor / channel index of pred: (nch, 1) rows, (1, A) planes [B][nch][A]
  const float* boxes;
  const float* scores;
  const float* labels;
  const float* box_div;  // optional [B][4] divisors of the corner boxes (mmdet rescale), pred mode only
};

struct Cand {
  float4 box;   // x1 y1 x2 y2 in network coordinates
  float obj, cls_conf, label_f;
  int label;
};

template <bool kFromPred>
__device__ __forceinline__ Cand load_cand(const Source& s, int b, int idx) {
  Cand c;
  if (kFromPred) {
    const float* r = s.pred + static_cast<int64_t>(b) * s.A * s.nch + idx * s.sa;
    const int64_t sc = s.sc;
    const float cx = __ldg(r), cy = __ldg(r + sc), w = __ldg(r + 2 * sc), h = __ldg(r + 3 * sc);
    // utils_bbox.py:381-386: corner = centre -/+ size / 2
    c.box.x = __fsub_rn(cx, __fdiv_rn(w, 2.0f));
    c.box.y = __fsub_rn(cy, __fdiv_rn(h, 2.0f));
    c.box.z = __fadd_rn(cx, __fdiv_rn(w, 2.0f));
    c.box.w = __fadd_rn(cy, __fdiv_rn(h, 2.0f));
    if (s.box_div != nullptr) {  // yolox_head.py:283-285: flatten_bboxes[..., :4] /= scale_factors
      const float* dv = s.box_div + 4 * b;
      c.box.x = __fdiv_rn(c.box.x, __ldg(dv));
      c.box.y = __fdiv_rn(c.box.y, __ldg(dv + 1));
      c.box.z = __fdiv_rn(c.box.z, __ldg(dv + 2));
      c.box.w = __fdiv_rn(c.box.w, __ldg(dv + 3));
    }
    c.obj = __ldg(r + 4 * sc);
    float best;
    int arg = 0;
    if (s.cls_logits) {
      // The fused prediction conv left the class logits raw; the sigmoid formula is the one of its decoding epilogue, so
      // the values are bit-identical to the decoded ones (utils_bbox.py:268 applies the sigmoid before the max of :398).
      // The sigmoid is monotonic up to a few ulps, so only the classes within a small window below the largest logit
      // (or in the saturated ranges) can attain the maximal probability: the sigmoid is evaluated for those only
      // (normally one per anchor), and the first index attaining the maximum wins, as torch.max does.
      float m = __ldg(r + 5 * sc);
      for (int k = 1; k < s.nc; ++k) m = fmaxf(m, __ldg(r + (5 + k) * sc));
      const float lim = m - 1e-3f * fmaxf(1.0f, fabsf(m));
      const bool all = (m < -80.0f);
      best = -1.0f;
      for (int k = 0; k < s.nc; ++k) {
        const float x = __ldg(r + (5 + k) * sc);
        if (all || x >= lim || x > 15.0f) {
          const float v = 1.0f / (1.0f + expf(-x));
          if (v > best) { best = v; arg = k; }
        }
      }
    } else {
      best = __ldg(r + 5 * sc);
      for (int k = 1; k < s.nc; ++k) {  // utils_bbox.py:398 torch.max: first maximal index
        const float v = __ldg(r + (5 + k) * sc);
        if (v > best) { best = v; arg = k; }
      }
    }
    c.cls_conf = best;
    c.label = arg;
    c.label_f = static_cast<float>(arg);
  } else {
    c.box = __ldg(reinterpret_cast<const float4*>(s.boxes) + idx);
    c.obj = __ldg(s.scores + idx);
    c.cls_conf = 1.0f;
    c.label_f = __ldg(s.labels + idx);
    c.label = static_cast<int>(c.label_f);
  }
  return c;
}

struct Work {
  int B, cap, P, nc;
  int32_t* cand_count;   // [B]
  uint32_t* max_bits;    // [B]
  uint32_t* min_bits;    // [B]
  int32_t* seg_start;    // [B][nc+1]
  int32_t* seg_kept;     // [B][nc]
  float* cand_score;     // [B][cap]
  int32_t* cand_idx;     // [B][cap]
  uint8_t* cand_label;   // [B][cap]
  uint64_t* keys;        // [B][P]
  uint64_t* kept_key;    // [B][cap]
  float4* kept_box;      // [B][cap]   spill of the kept list beyond shared memory (fallback kernel)
  float4* nbox;          // [B][cap]   NMS-space box of every sorted candidate (offset applied for the trick)
  int32_t* seg_tile_off; // [B*nc + 1] exclusive prefix of per-segment tile counts (mask path)
  int64_t* seg_word_off; // [B*nc]     first mask word of the segment, or -1 when it runs on the fallback kernel
  unsigned long long* row_any;  // [B][P/64] bit i set: sorted candidate i suppresses at least one later box
  unsigned long long* mask;  // [mask_words]
  int64_t mask_words;
  int32_t max_scan_tiles;    // segments with more 64-box chunks than this use the fallback kernel
  int32_t topk;              // > 0: only the `topk` best kept boxes per image are wanted (max_det) and the kept list of
                             // a segment fits shared memory: every segment runs on the blocked-greedy kernel, which
                             // stops once it has kept `topk` boxes (a later box of the segment cannot rank in the top
                             // `topk` of its image); no suppression bitmask is built at all
};

struct ImageMode {
  bool use_offsets;   // coordinate trick
  bool per_class;     // class-segmented NMS
  float offset_scale; // max_coord + 1
};

__device__ __forceinline__ ImageMode image_mode(const Work& w, int b, int strategy) {
  ImageMode m;
  const int n = w.cand_count[b];
  bool trick;
  if (strategy == GLSDET_NMS_COORD_TRICK || strategy == GLSDET_NMS_MMCV) trick = true;
  else if (strategy == GLSDET_NMS_PER_CLASS) trick = false;
  else if (strategy == GLSDET_NMS_AUTO_CUDA) trick = (4ll * n <= 100000);
  else trick = (4ll * n <= 4000);
  m.use_offsets = trick;
  const float maxc = order_bits_float(w.max_bits[b]);
  const float minc = order_bits_float(w.min_bits[b]);
  m.offset_scale = __fadd_rn(maxc, 1.0f);
  // class separation needs a positive gap after rounding: min >= -0.5 leaves 0.5, and offsets below 2^21 keep
  // every rounding error under 0.125
  m.per_class = !trick || (n > 0 && minc >= -0.5f && static_cast<float>(w.nc) * m.offset_scale < 2097152.0f);
  // mmcv: from split_thr = 10000 boxes on, NMS runs class by class (on the shifted boxes)
  if (strategy == GLSDET_NMS_MMCV && n >= 10000) m.per_class = true;
  return m;
}

__global__ void reset_kernel(Work w) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < w.B) {
    w.cand_count[i] = 0;
    w.max_bits[i] = 0u;
    w.min_bits[i] = 0xFFFFFFFFu;
  }
}

// ---------------------------------------------------------------------------------------------- filter
template <bool kFromPred>
__global__ void __launch_bounds__(256) filter_kernel(Source s, Work w, float conf_thres) {
  pdl_prologue();
  // Compaction with ONE atomicAdd per CTA and iteration (warp ballots -> shared prefix -> a single slot range), and one
  // atomicMax / atomicMin per CTA for the coordinate range: per-warp atomics on the 16 per-image counters serialised at
  // the L2 (2 700 same-address atomics per counter).  The order of the compacted candidates is irrelevant: the sort keys
  // carry the anchor index.
  __shared__ int s_cnt[8];
  __shared__ int s_base;
  __shared__ float s_hi[8], s_lo[8];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float hi = -INFINITY, lo = INFINITY;
  for (int base = blockIdx.x * blockDim.x; base < s.A; base += gridDim.x * blockDim.x) {
    const int a = base + threadIdx.x;
    bool pass = false;
    Cand c;
    float score = 0.0f;
    if (a < s.A) {
      c = load_cand<kFromPred>(s, b, a);
      score = kFromPred ? __fmul_rn(c.obj, c.cls_conf) : c.obj;
      pass = kFromPred ? (score >= conf_thres) : true;  // utils_bbox.py:403 (>=)
    }
    const uint32_t mask = __ballot_sync(0xffffffffu, pass);
    if (lane == 0) s_cnt[warp] = __popc(mask);
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) { const int n = s_cnt[i]; s_cnt[i] = tot; tot += n; }   // exclusive prefix
      s_base = tot ? atomicAdd(&w.cand_count[b], tot) : 0;
    }
    __syncthreads();
    if (pass) {
      const int slot = s_base + s_cnt[warp] + __popc(mask & ((1u << lane) - 1u));
      const int64_t o = static_cast<int64_t>(b) * w.cap + slot;
      w.cand_score[o] = score;
      w.cand_idx[o] = a;
      w.cand_label[o] = static_cast<uint8_t>(c.label);
      hi = fmaxf(hi, fmaxf(fmaxf(c.box.x, c.box.y), fmaxf(c.box.z, c.box.w)));
      lo = fminf(lo, fminf(fminf(c.box.x, c.box.y), fminf(c.box.z, c.box.w)));
    }
    __syncthreads();   // s_cnt / s_base are rewritten by the next iteration
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
  }
  if (lane == 0) { s_hi[warp] = hi; s_lo[warp] = lo; }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 1; i < 8; ++i) { hi = fmaxf(hi, s_hi[i]); lo = fminf(lo, s_lo[i]); }
    if (hi != -INFINITY) {   // at least one candidate in this CTA
      atomicMax(&w.max_bits[b], float_order_bits(hi));
      atomicMin(&w.min_bits[b], float_order_bits(lo));
    }
  }
}

__device__ __forceinline__ int sort_extent(int n) {
  int p = kSortChunk;
  while (p < n) p <<= 1;
  return p;
}

__global__ void __launch_bounds__(256) build_keys_kernel(Work w, int strategy) {
  pdl_prologue();
  const int b = blockIdx.y;
  const int n = w.cand_count[b];
  const int pb = sort_extent(n);
  const ImageMode m = image_mode(w, b, strategy);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pb; i += gridDim.x * blockDim.x) {
    uint64_t k = kPadKey;
    if (i < n) {
      const int64_t o = static_cast<int64_t>(b) * w.cap + i;
      k = make_key(m.per_class ? w.cand_label[o] : 0u, w.cand_score[o], static_cast<uint32_t>(w.cand_idx[o]));
    }
    w.keys[static_cast<int64_t>(b) * w.P + i] = k;
    if ((i & 63) == 0) w.row_any[static_cast<int64_t>(b) * (w.P >> 6) + (i >> 6)] = 0ull;
  }
}

// ---------------------------------------------------------------------------------------------- bitonic sort
__device__ __forceinline__ void cmp_swap(uint64_t& a, uint64_t& b, bool asc) {
  if ((a > b) == asc) { const uint64_t t = a; a = b; b = t; }
}

// sorts every kSortChunk-sized chunk (all levels k <= kSortChunk), direction alternating by global index
__global__ void __launch_bounds__(kSortThreads) bitonic_local_sort_kernel(Work w) {
  pdl_prologue();
  __shared__ uint64_t sk[kSortChunk];
  const int b = blockIdx.y;
  const int pb = sort_extent(w.cand_count[b]);
  const int base = blockIdx.x * kSortChunk;
  if (base >= pb) return;
  uint64_t* g = w.keys + static_cast<int64_t>(b) * w.P + base;
  for (int i = threadIdx.x; i < kSortChunk; i += kSortThreads) sk[i] = g[i];
  __syncthreads();
  for (int k = 2; k <= kSortChunk; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < kSortChunk / 2; t += kSortThreads) {
        const int i = 2 * j * (t / j) + (t % j);
        const bool asc = (((base + i) & k) == 0);
        cmp_swap(sk[i], sk[i + j], asc);
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < kSortChunk; i += kSortThreads) g[i] = sk[i];
}

// one compare-exchange step (level k, distance j >= kSortChunk) over the whole padded array
__global__ void __launch_bounds__(256) bitonic_global_step_kernel(Work w, int k, int j) {
  pdl_prologue();
  const int b = blockIdx.y;
  const int pb = sort_extent(w.cand_count[b]);
  if (k > pb) return;
  uint64_t* g = w.keys + static_cast<int64_t>(b) * w.P;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < pb / 2; t += gridDim.x * blockDim.x) {
    const int i = 2 * j * (t / j) + (t % j);
    const bool asc = ((i & k) == 0);
    uint64_t a = g[i], c = g[i + j];
    if ((a > c) == asc) { g[i] = c; g[i + j] = a; }
  }
}

// finishes level k inside each chunk (distances kSortChunk/2 .. 1)
__global__ void __launch_bounds__(kSortThreads) bitonic_local_merge_kernel(Work w, int k) {
  pdl_prologue();
  __shared__ uint64_t sk[kSortChunk];
  const int b = blockIdx.y;
  const int pb = sort_extent(w.cand_count[b]);
  const int base = blockIdx.x * kSortChunk;
  if (k > pb || base >= pb) return;
  uint64_t* g = w.keys + static_cast<int64_t>(b) * w.P + base;
  for (int i = threadIdx.x; i < kSortChunk; i += kSortThreads) sk[i] = g[i];
  __syncthreads();
  for (int j = kSortChunk >> 1; j > 0; j >>= 1) {
    for (int t = threadIdx.x; t < kSortChunk / 2; t += kSortThreads) {
      const int i = 2 * j * (t / j) + (t % j);
      const bool asc = (((base + i) & k) == 0);
      cmp_swap(sk[i], sk[i + j], asc);
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < kSortChunk; i += kSortThreads) g[i] = sk[i];
}

// ---------------------------------------------------------------------------------------------- segments
__global__ void segment_bounds_kernel(Work w, int strategy) {
  pdl_prologue();
  const int b = blockIdx.x;
  const int n = w.cand_count[b];
  const ImageMode m = image_mode(w, b, strategy);
  const uint64_t* g = w.keys + static_cast<int64_t>(b) * w.P;
  for (int c = threadIdx.x; c <= w.nc; c += blockDim.x) {
    int pos;
    if (!m.per_class) {
      pos = (c == 0) ? 0 : n;
    } else {
      const uint64_t target = static_cast<uint64_t>(c) << 56;  // first key with label >= c
      int lo = 0, hi = n;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (g[mid] < target) lo = mid + 1; else hi = mid;
      }
      pos = (c == w.nc) ? n : lo;
    }
    w.seg_start[b * (w.nc + 1) + c] = pos;
  }
}

// ---------------------------------------------------------------------------------------------- NMS
// torchvision nms_kernel_impl: ovr = inter / (area_i + area_j - inter); suppress iff ovr > thr
__device__ __forceinline__ bool iou_exceeds(const float4& a, float aarea, const float4& b, float barea, float thr,
                                            bool thr_nonneg) {
  const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  const float ww = fmaxf(0.0f, __fsub_rn(xx2, xx1));
  const float hh = fmaxf(0.0f, __fsub_rn(yy2, yy1));
  const float inter = __fmul_rn(ww, hh);
  if (inter <= 0.0f && thr_nonneg) return false;  // 0/x = 0 or NaN: never > thr when thr >= 0
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aarea, barea), inter));
  return ovr > thr;
}

// NMS-space box of every sorted candidate: raw box, or box + label * (max_coord + 1) for the coordinate trick
template <bool kFromPred>
__global__ void __launch_bounds__(256) box_prep_kernel(Source s, Work w, int strategy) {
  pdl_prologue();
  const int b = blockIdx.y;
  const int n = w.cand_count[b];
  const ImageMode m = image_mode(w, b, strategy);
  const uint64_t* keys = w.keys + static_cast<int64_t>(b) * w.P;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const Cand c = load_cand<kFromPred>(s, b, static_cast<int>(key_idx(keys[i])));
    float4 bx = c.box;
    if (m.use_offsets) {
      // boxes.py _batched_nms_coordinate_trick: offsets = idxs * (max_coordinate + 1); boxes + offsets[:, None]
      const float off = __fmul_rn(c.label_f, m.offset_scale);
      bx.x = __fadd_rn(bx.x, off); bx.y = __fadd_rn(bx.y, off);
      bx.z = __fadd_rn(bx.z, off); bx.w = __fadd_rn(bx.w, off);
    }
    w.nbox[static_cast<int64_t>(b) * w.cap + i] = bx;
  }
}

// Mask layout of a T-tile segment: upper-triangular list of 64x64 tiles, row-tile major; the 64 words of a tile
// (one per row) are contiguous, so a tile is written with one coalesced 512-byte store.
__device__ __forceinline__ int64_t tri_tile(int T, int r, int c) {
  return static_cast<int64_t>(r) * T - (static_cast<int64_t>(r) * (r - 1)) / 2 + (c - r);
}
__device__ __forceinline__ int64_t tri_word(int T, int r, int i, int c) { return tri_tile(T, r, c) * 64 + i; }

__device__ __forceinline__ float box_area(const float4& b) {
  return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}

// One block: decide which segments get a bitmask (in segment order, until the word budget is spent) and build the
// tile work list.  Segment ids are b * nc + c.
__global__ void __launch_bounds__(1024) seg_plan_kernel(Work w) {
  pdl_prologue();
  __shared__ long long s_words[1024];
  __shared__ int s_tiles[1024];
  __shared__ long long carry_words;
  __shared__ int carry_tiles;
  const int total = w.B * w.nc;
  if (threadIdx.x == 0) { carry_words = 0; carry_tiles = 0; }
  __syncthreads();
  for (int base = 0; base < total; base += 1024) {
    const int sgm = base + threadIdx.x;
    long long words = 0;
    int tiles = 0, n = 0, T = 0;
    if (sgm < total) {
      const int b = sgm / w.nc, c = sgm % w.nc;
      n = w.seg_start[b * (w.nc + 1) + c + 1] - w.seg_start[b * (w.nc + 1) + c];
      T = (n + 63) >> 6;
      if (w.topk == 0 && T <= w.max_scan_tiles && T < 46000) {  // T*T must fit an int; larger segments go to the fallback
        tiles = T * (T + 1) / 2;     // upper triangle only (column tile >= row tile)
        words = 64ll * tiles;
      }
    }
    s_words[threadIdx.x] = words;
    s_tiles[threadIdx.x] = tiles;
    __syncthreads();
    // serial prefix by thread 0 over <= 1024 entries: the budget cut-off is order dependent and tiny
    if (threadIdx.x == 0) {
      long long cw = carry_words;
      int ct = carry_tiles;
      for (int i = 0; i < 1024 && base + i < total; ++i) {
        const long long wd = s_words[i];
        const bool fits = wd > 0 && cw + wd <= w.mask_words && s_tiles[i] > 0 && ct + s_tiles[i] > ct;
        w.seg_tile_off[base + i] = ct;
        w.seg_word_off[base + i] = fits ? cw : -1;
        if (fits) { cw += wd; ct += s_tiles[i]; }
      }
      carry_words = cw;
      carry_tiles = ct;
      if (base + 1024 >= total) w.seg_tile_off[total] = ct;
    }
    __syncthreads();
  }
}

// Persistent CTAs of 4 x 64 threads; every 64-thread group owns one 64x64 tile at a time (row tile r, column tile
// c >= r), statically strided over the work list.  Thread i of the group tests row box i against the 64 column
// boxes staged in shared memory and emits one 64-bit word.
constexpr int kMaskGroups = 4;
__global__ void __launch_bounds__(64 * kMaskGroups) mask_tiles_kernel(Work w, float thr) {
  pdl_prologue();
  __shared__ float4 cbox[kMaskGroups][64];
  __shared__ float carea[kMaskGroups][64];
  const int total_seg = w.B * w.nc;
  const int total_tiles = w.seg_tile_off[total_seg];
  const bool thr_nonneg = (thr >= 0.0f);
  const int grp = threadIdx.x >> 6, tid = threadIdx.x & 63;
  int cur_lo = 0, cur_hi = 0, sgm = 0, s0 = 0, n = 0, T = 0, b = 0;
  for (int t = blockIdx.x * kMaskGroups + grp; t < total_tiles; t += gridDim.x * kMaskGroups) {
    if (t >= cur_hi || t < cur_lo) {
      int lo = 0, hi = total_seg;  // seg_tile_off[lo] <= t < seg_tile_off[hi]
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (w.seg_tile_off[mid] <= t) lo = mid; else hi = mid;
      }
      sgm = lo;
      cur_lo = w.seg_tile_off[sgm];
      cur_hi = w.seg_tile_off[sgm + 1];
      b = sgm / w.nc;
      const int c = sgm % w.nc;
      s0 = w.seg_start[b * (w.nc + 1) + c];
      n = w.seg_start[b * (w.nc + 1) + c + 1] - s0;
      T = (n + 63) >> 6;
    }
    // invert local = r*T - r(r-1)/2 + (c - r)
    const int local = t - cur_lo;
    const float fT = static_cast<float>(2 * T + 1);
    int r = static_cast<int>((fT - sqrtf(fmaxf(fT * fT - 8.0f * static_cast<float>(local), 0.0f))) * 0.5f);
    r = max(0, min(r, T - 1));
    while (r > 0 && static_cast<int>(tri_tile(T, r, r)) > local) --r;
    while (r + 1 < T && static_cast<int>(tri_tile(T, r + 1, r + 1)) <= local) ++r;
    const int cc = r + (local - static_cast<int>(tri_tile(T, r, r)));
    const float4* bx = w.nbox + static_cast<int64_t>(b) * w.cap + s0;
    const int cj = cc * 64 + tid;
    if (cj < n) { cbox[grp][tid] = bx[cj]; carea[grp][tid] = box_area(cbox[grp][tid]); }
    asm volatile("bar.sync %0, 64;" ::"r"(grp + 1) : "memory");
    const int ri = r * 64 + tid;
    unsigned long long bits = 0ull;
    if (ri < n) {
      const float4 rb = bx[ri];
      const float ra = box_area(rb);
      const int jn = min(64, n - cc * 64);
      const int j0 = (cc == r) ? tid + 1 : 0;
      if (thr_nonneg) {
        // common case first: two subtractions decide "no overlap" (inter == 0 never exceeds thr >= 0); the exact
        // IEEE division runs only for the few pairs that really intersect
        uint32_t lo = 0u, hi = 0u;
#pragma unroll 4
        for (int j = j0; j < jn; ++j) {
          const float4 cb = cbox[grp][j];
          const float ww = __fsub_rn(fminf(rb.z, cb.z), fmaxf(rb.x, cb.x));
          const float hh = __fsub_rn(fminf(rb.w, cb.w), fmaxf(rb.y, cb.y));
          if (fminf(ww, hh) > 0.0f) {
            if (iou_exceeds(rb, ra, cb, carea[grp][j], thr, true)) {
              if (j < 32) lo |= 1u << j; else hi |= 1u << (j - 32);
            }
          }
        }
        bits = (static_cast<unsigned long long>(hi) << 32) | lo;
      } else {
        for (int j = j0; j < jn; ++j)
          if (iou_exceeds(rb, ra, cbox[grp][j], carea[grp][j], thr, false)) bits |= (1ull << j);
      }
      if (bits) atomicOr(&w.row_any[static_cast<int64_t>(b) * (w.P >> 6) + ((s0 + ri) >> 6)], 1ull << ((s0 + ri) & 63));
    }
    w.mask[w.seg_word_off[sgm] + tri_word(T, r, tid, cc)] = bits;
    asm volatile("bar.sync %0, 64;" ::"r"(grp + 1) : "memory");
  }
}

// One CTA per mask-path segment: sequential over 64-box chunks, parallel inside.  Suppression is sparse (most
// boxes suppress nothing), so only rows flagged in row_any are ever read from the mask, and the serial part of a
// chunk only walks the flagged rows.  The diagonal words of chunk r+1 are prefetched while chunk r is resolved.
constexpr int kScanThreads = 512;
__global__ void __launch_bounds__(kScanThreads) scan_kernel(Work w) {
  pdl_prologue();
  extern __shared__ unsigned long long scan_smem[];  // removed[T] | any[T]
  __shared__ unsigned long long diag[2][64];
  __shared__ unsigned long long kept_bits_s;
  __shared__ int sup_rows[64];
  __shared__ int sup_n;
  __shared__ int kept_n;
  const int seg = blockIdx.x, b = blockIdx.y;
  const int sgm = b * w.nc + seg;
  const int64_t woff = w.seg_word_off[sgm];
  if (woff < 0) return;  // empty, or handled by nms_segment_kernel
  const int s0 = w.seg_start[b * (w.nc + 1) + seg];
  const int n = w.seg_start[b * (w.nc + 1) + seg + 1] - s0;
  const int T = (n + 63) >> 6;
  unsigned long long* removed = scan_smem;
  unsigned long long* any_s = scan_smem + T;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const unsigned long long* mask = w.mask + woff;
  const unsigned long long* any_words = w.row_any + static_cast<int64_t>(b) * (w.P >> 6);
  const uint64_t* keys = w.keys + static_cast<int64_t>(b) * w.P + s0;
  uint64_t* gk_key = w.kept_key + static_cast<int64_t>(b) * w.cap + s0;
  for (int r = tid; r < T; r += kScanThreads) {
    removed[r] = 0ull;
    // 64 'suppresses something' flags of chunk r (the segment need not start on a word boundary)
    const int g0 = s0 + r * 64;
    unsigned long long any = any_words[g0 >> 6] >> (g0 & 63);
    if (g0 & 63) any |= any_words[(g0 >> 6) + 1] << (64 - (g0 & 63));
    const int mc = min(64, n - r * 64);
    if (mc < 64) any &= ((1ull << mc) - 1ull);
    any_s[r] = any;
  }
  if (tid == 0) kept_n = 0;
  __syncthreads();
  if (tid < 64) {
    const unsigned long long a0 = any_s[0];
    diag[0][tid] = ((a0 >> tid) & 1ull) ? mask[tri_word(T, 0, tid, 0)] : 0ull;
  }
  __syncthreads();
  for (int r = 0; r < T; ++r) {
    const int mcnt = min(64, n - r * 64);
    const int cur = r & 1;
    const unsigned long long any = any_s[r];
    // prefetch: keys of this chunk, diagonal words of the next one (both independent of the resolve below)
    uint64_t my_key = 0;
    if (tid < mcnt) my_key = keys[r * 64 + tid];
    unsigned long long next_diag = 0ull;
    if (tid >= 64 && tid < 128 && r + 1 < T) {
      const int i = tid - 64;
      if ((any_s[r + 1] >> i) & 1ull) next_diag = mask[tri_word(T, r + 1, i, r + 1)];
    }
    if (tid == 0) {
      unsigned long long alive = ~removed[r];
      if (mcnt < 64) alive &= ((1ull << mcnt) - 1ull);
      int ns = 0;
      unsigned long long m = alive & any;   // only flagged rows can suppress; walk them in score order
      while (m) {
        const int i = __ffsll(static_cast<long long>(m)) - 1;
        m &= m - 1ull;
        if ((alive >> i) & 1ull) {
          const unsigned long long d = diag[cur][i];
          alive &= ~d;
          m &= ~d;
          sup_rows[ns++] = i;
        }
      }
      kept_bits_s = alive;
      sup_n = ns;
    }
    __syncthreads();
    const unsigned long long kept = kept_bits_s;
    const int ns = sup_n;
    const int kn = kept_n;
    if (tid < mcnt && ((kept >> tid) & 1ull))
      gk_key[kn + __popcll(kept & ((1ull << tid) - 1ull))] = my_key & 0x00FFFFFFFFFFFFFFull;
    if (tid >= 64 && tid < 128) diag[cur ^ 1][tid - 64] = next_diag;
    // OR the kept, suppressing rows into the bitmap of the later chunks: warps take rows, lanes take column tiles
    const int ncol = T - r - 1;
    if (ncol > 0 && ns > 0) {
      const int64_t tbase = tri_tile(T, r, r + 1) * 64;
      for (int k = warp; k < ns; k += kScanThreads / 32) {
        const int i = sup_rows[k];
        for (int c = lane; c < ncol; c += 32) {
          const unsigned long long mm = mask[tbase + static_cast<int64_t>(c) * 64 + i];
          if (mm) atomicOr(&removed[r + 1 + c], mm);
        }
      }
    }
    __syncthreads();
    if (tid == 0) kept_n = kn + __popcll(kept);  // read again only after the next barrier
  }
  if (tid == 0) w.seg_kept[sgm] = kept_n;
}

template <bool kFromPred>
__global__ void __launch_bounds__(kNmsThreads) nms_segment_kernel(Source s, Work w, float thr, int strategy) {
  pdl_prologue();
  __shared__ float4 kbox[kKeptSmem];
  __shared__ float karea[kKeptSmem];
  __shared__ float4 cbox[64];
  __shared__ float carea[64];
  __shared__ uint64_t ckey[64];
  __shared__ unsigned long long cmask[64];
  __shared__ unsigned long long sup_prev;
  __shared__ unsigned long long kept_bits;
  __shared__ int kept_n;

  const int seg = blockIdx.x, b = blockIdx.y;
  const int s0 = w.seg_start[b * (w.nc + 1) + seg];
  const int s1 = w.seg_start[b * (w.nc + 1) + seg + 1];
  const int n = s1 - s0;
  const int tid = threadIdx.x;
  if (n <= 0) {
    if (tid == 0) w.seg_kept[b * w.nc + seg] = 0;
    return;
  }
  if (w.seg_word_off[b * w.nc + seg] >= 0) return;  // this segment runs on the bitmask path
  const bool thr_nonneg = (thr >= 0.0f);
  const uint64_t* keys = w.keys + static_cast<int64_t>(b) * w.P + s0;
  float4* gk_box = w.kept_box + static_cast<int64_t>(b) * w.cap + s0;
  uint64_t* gk_key = w.kept_key + static_cast<int64_t>(b) * w.cap + s0;
  if (tid == 0) kept_n = 0;
  __syncthreads();

  for (int c0 = 0; c0 < n; c0 += 64) {
    if (w.topk > 0 && kept_n >= w.topk) break;   // uniform: kept_n was published before the last barrier
    const int mcnt = min(64, n - c0);
    if (tid < 64) {
      cmask[tid] = 0ull;
      if (tid < mcnt) {
        const float4 bx = w.nbox[static_cast<int64_t>(b) * w.cap + s0 + c0 + tid];
        cbox[tid] = bx;
        carea[tid] = box_area(bx);
        ckey[tid] = keys[c0 + tid];
      }
    }
    if (tid == 0) sup_prev = 0ull;
    __syncthreads();

    // phase 1: which candidates of this chunk are suppressed by an already kept box
    {
      const int kn = kept_n;
      unsigned long long sup = 0ull;
      for (int k = tid; k < kn; k += kNmsThreads) {
        float4 kb;
        float ka;
        if (k < kKeptSmem) { kb = kbox[k]; ka = karea[k]; }
        else {
          kb = gk_box[k];
          ka = __fmul_rn(__fsub_rn(kb.z, kb.x), __fsub_rn(kb.w, kb.y));
        }
        for (int j = 0; j < mcnt; ++j)
          if (iou_exceeds(kb, ka, cbox[j], carea[j], thr, thr_nonneg)) sup |= (1ull << j);
      }
      uint32_t lo = static_cast<uint32_t>(sup), hi = static_cast<uint32_t>(sup >> 32);
      lo = __reduce_or_sync(0xffffffffu, lo);
      hi = __reduce_or_sync(0xffffffffu, hi);
      if ((tid & 31) == 0 && (lo | hi)) atomicOr(&sup_prev, (static_cast<unsigned long long>(hi) << 32) | lo);
    }
    // phase 2: pairwise masks inside the chunk (row i suppresses later column j)
    {
      constexpr int kParts = kNmsThreads / 64, kCols = 64 / kParts;
      const int i = tid & 63, part = tid >> 6;  // kParts parts x kCols columns
      if (i < mcnt) {
        unsigned long long bits = 0ull;
        const float4 bi = cbox[i];
        const float ai = carea[i];
        for (int j = max(i + 1, part * kCols); j < min(mcnt, part * kCols + kCols); ++j)
          if (iou_exceeds(bi, ai, cbox[j], carea[j], thr, thr_nonneg)) bits |= (1ull << j);
        if (bits) atomicOr(&cmask[i], bits);
      }
    }
    __syncthreads();
    // phase 3: serial greedy resolution of the 64 candidates
    if (tid == 0) {
      unsigned long long alive = ~sup_prev;
      if (mcnt < 64) alive &= ((1ull << mcnt) - 1ull);
      unsigned long long kept = 0ull;
      for (int i = 0; i < mcnt; ++i) {
        if ((alive >> i) & 1ull) {
          kept |= (1ull << i);
          alive &= ~cmask[i];
        }
      }
      kept_bits = kept;
    }
    __syncthreads();
    // phase 4: append the survivors to the kept list (order preserved)
    {
      const unsigned long long kept = kept_bits;
      const int kn = kept_n;
      if (tid < mcnt && ((kept >> tid) & 1ull)) {
        const int pos = kn + __popcll(kept & ((1ull << tid) - 1ull));
        if (pos < kKeptSmem) { kbox[pos] = cbox[tid]; karea[pos] = carea[tid]; }
        else gk_box[pos] = cbox[tid];
        gk_key[pos] = ckey[tid] & 0x00FFFFFFFFFFFFFFull;  // (~score, anchor): order across classes
      }
      __syncthreads();
      if (tid == 0) kept_n = kn + __popcll(kept);
    }
    __syncthreads();
  }
  if (tid == 0) w.seg_kept[b * w.nc + seg] = kept_n;
}

// ---------------------------------------------------------------------------------------------- output
template <bool kFromPred>
__global__ void __launch_bounds__(256) rank_gather_kernel(Source s, Work w, int max_det, float* det,
                                                          int32_t* det_count, int32_t* keep_index) {
  pdl_prologue();
  __shared__ int sstart[kMaxClasses + 1];
  __shared__ int skept[kMaxClasses];
  const int b = blockIdx.y;
  for (int c = threadIdx.x; c <= w.nc; c += blockDim.x) {
    sstart[c] = w.seg_start[b * (w.nc + 1) + c];
    if (c < w.nc) skept[c] = w.seg_kept[b * w.nc + c];
  }
  __syncthreads();
  const int n = sstart[w.nc];
  const uint64_t* kk = w.kept_key + static_cast<int64_t>(b) * w.cap;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int total = 0;
    for (int c = 0; c < w.nc; ++c) total += skept[c];
    det_count[b] = total < max_det ? total : max_det;
  }
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    int c = 0;
    while (c + 1 < w.nc && p >= sstart[c + 1]) ++c;
    if (p - sstart[c] >= skept[c]) continue;  // not a kept slot
    const uint64_t key = kk[p];
    int rank = 0;
    for (int c2 = 0; c2 < w.nc; ++c2) {
      const uint64_t* base = kk + sstart[c2];
      int lo = 0, hi = skept[c2];
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (base[mid] < key) lo = mid + 1; else hi = mid;
      }
      rank += lo;
    }
    if (rank >= max_det) continue;
    const int idx = static_cast<int>(key_idx(key));
    if (keep_index) keep_index[static_cast<int64_t>(b) * max_det + rank] = idx;
    if (kFromPred && det) {
      const Cand cd = load_cand<true>(s, b, idx);
      float* o = det + (static_cast<int64_t>(b) * max_det + rank) * 7;
      o[0] = cd.box.x; o[1] = cd.box.y; o[2] = cd.box.z; o[3] = cd.box.w;
      o[4] = cd.obj; o[5] = cd.cls_conf; o[6] = cd.label_f;  // utils_bbox.py:411 row layout
    }
  }
}

// ---------------------------------------------------------------------------------------------- host
inline int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}
inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct Layout {
  int64_t off_count, off_max, off_min, off_segstart, off_segkept, off_score, off_idx, off_label, off_keys,
      off_keptkey, off_keptbox, off_nbox, off_tileoff, off_wordoff, off_rowany, off_mask, mask_words, total;
};

constexpr int kMaxScanTiles = 2816;  // 2 x 22 KB of shared memory for the 'removed' and 'any' bitmaps

inline int64_t mask_budget_words(int B, int cap) {
  int64_t bytes = static_cast<int64_t>(B) * cap * 768;
  if (bytes < (32ll << 20)) bytes = 32ll << 20;
  if (bytes > (3072ll << 20)) bytes = 3072ll << 20;
  return bytes / 8;
}

Layout make_layout(int B, int cap, int P, int nc) {
  Layout l;
  int64_t o = 0;
  l.off_count = o; o = align_up(o + 4ll * B, 256);
  l.off_max = o; o = align_up(o + 4ll * B, 256);
  l.off_min = o; o = align_up(o + 4ll * B, 256);
  l.off_segstart = o; o = align_up(o + 4ll * B * (nc + 1), 256);
  l.off_segkept = o; o = align_up(o + 4ll * B * nc, 256);
  l.off_score = o; o = align_up(o + 4ll * B * cap, 256);
  l.off_idx = o; o = align_up(o + 4ll * B * cap, 256);
  l.off_label = o; o = align_up(o + 1ll * B * cap, 256);
  l.off_keys = o; o = align_up(o + 8ll * B * P, 256);
  l.off_keptkey = o; o = align_up(o + 8ll * B * cap, 256);
  l.off_keptbox = o; o = align_up(o + 16ll * B * cap, 256);
  l.off_nbox = o; o = align_up(o + 16ll * B * cap, 256);
  l.off_tileoff = o; o = align_up(o + 4ll * (static_cast<int64_t>(B) * nc + 1), 256);
  l.off_wordoff = o; o = align_up(o + 8ll * B * nc, 256);
  l.off_rowany = o; o = align_up(o + 8ll * B * (P / 64 + 2), 256);
  l.mask_words = mask_budget_words(B, cap);
  l.off_mask = o; o = align_up(o + 8ll * l.mask_words, 256);
  l.total = o;
  return l;
}

Work make_work(void* ws, int B, int cap, int nc) {
  const int P = pow2_ceil(cap < kSortChunk ? kSortChunk : cap);
  const Layout l = make_layout(B, cap, P, nc);
  uint8_t* p = static_cast<uint8_t*>(ws);
  Work w;
  w.B = B; w.cap = cap; w.P = P; w.nc = nc;
  w.cand_count = reinterpret_cast<int32_t*>(p + l.off_count);
  w.max_bits = reinterpret_cast<uint32_t*>(p + l.off_max);
  w.min_bits = reinterpret_cast<uint32_t*>(p + l.off_min);
  w.seg_start = reinterpret_cast<int32_t*>(p + l.off_segstart);
  w.seg_kept = reinterpret_cast<int32_t*>(p + l.off_segkept);
  w.cand_score = reinterpret_cast<float*>(p + l.off_score);
  w.cand_idx = reinterpret_cast<int32_t*>(p + l.off_idx);
  w.cand_label = p + l.off_label;
  w.keys = reinterpret_cast<uint64_t*>(p + l.off_keys);
  w.kept_key = reinterpret_cast<uint64_t*>(p + l.off_keptkey);
  w.kept_box = reinterpret_cast<float4*>(p + l.off_keptbox);
  w.nbox = reinterpret_cast<float4*>(p + l.off_nbox);
  w.seg_tile_off = reinterpret_cast<int32_t*>(p + l.off_tileoff);
  w.seg_word_off = reinterpret_cast<int64_t*>(p + l.off_wordoff);
  w.row_any = reinterpret_cast<unsigned long long*>(p + l.off_rowany);
  w.mask = reinterpret_cast<unsigned long long*>(p + l.off_mask);
  w.mask_words = l.mask_words;
  if (const char* e = getenv("GLSDET_NMS_MASK_WORDS")) {  // tests: shrink the budget to force the fallback kernel
    const long long v = atoll(e);
    if (v >= 0 && v < w.mask_words) w.mask_words = v;
  }
  w.max_scan_tiles = kMaxScanTiles;
  w.topk = 0;
  return w;
}

int64_t workspace_bytes(int B, int cap, int nc) {
  const int P = pow2_ceil(cap < kSortChunk ? kSortChunk : cap);
  return make_layout(B, cap, P, nc).total;
}

template <bool kFromPred>
int run_pipeline(const Source& s, const Work& w, float conf_thres, float nms_thres, int strategy, int max_det,
                 float* det, int32_t* det_count, int32_t* keep_index, cudaStream_t st) {
  const int B = w.B;
  launch_pdl(reset_kernel, (B + 255) / 256, 256, 0, st, w);
  if (int rc = count_launch("reset_kernel")) return rc;
  {
    int gx = (s.A + 255) / 256;
    if (gx > 4096) gx = 4096;
    launch_pdl(filter_kernel<kFromPred>, dim3(gx, B), 256, 0, st, s, w, conf_thres);
    if (int rc = count_launch("filter_kernel")) return rc;
  }
  {
    int gx = w.P / 256;
    if (gx > 1024) gx = 1024;
    launch_pdl(build_keys_kernel, dim3(gx, B), 256, 0, st, w, strategy);
    if (int rc = count_launch("build_keys_kernel")) return rc;
  }
  const int chunks = w.P / kSortChunk;
  launch_pdl(bitonic_local_sort_kernel, dim3(chunks, B), kSortThreads, 0, st, w);
  if (int rc = count_launch("bitonic_local_sort_kernel")) return rc;
  for (int k = 2 * kSortChunk; k <= w.P; k <<= 1) {
    for (int j = k >> 1; j >= kSortChunk; j >>= 1) {
      int gx = w.P / 2 / 256;
      if (gx > 2048) gx = 2048;
      launch_pdl(bitonic_global_step_kernel, dim3(gx, B), 256, 0, st, w, k, j);
      if (int rc = count_launch("bitonic_global_step_kernel")) return rc;
    }
    launch_pdl(bitonic_local_merge_kernel, dim3(chunks, B), kSortThreads, 0, st, w, k);
    if (int rc = count_launch("bitonic_local_merge_kernel")) return rc;
  }
  launch_pdl(segment_bounds_kernel, B, 256, 0, st, w, strategy);
  if (int rc = count_launch("segment_bounds_kernel")) return rc;
  {
    int gx = (w.cap + 255) / 256;
    if (gx > 1024) gx = 1024;
    launch_pdl(box_prep_kernel<kFromPred>, dim3(gx, B), 256, 0, st, s, w, strategy);
    if (int rc = count_launch("box_prep_kernel")) return rc;
  }
  launch_pdl(seg_plan_kernel, 1, 1024, 0, st, w);
  if (int rc = count_launch("seg_plan_kernel")) return rc;
  if (w.topk == 0) {
    launch_pdl(mask_tiles_kernel, device_sm_count() * 6, 64 * kMaskGroups, 0, st, w, nms_thres);
    if (int rc = count_launch("mask_tiles_kernel")) return rc;
  }
  if (w.topk == 0) {
    int t = (w.cap + 63) / 64;
    if (t > kMaxScanTiles) t = kMaxScanTiles;
    launch_pdl(scan_kernel, dim3(w.nc, B), kScanThreads, static_cast<size_t>(t) * 16, st, w);
    if (int rc = count_launch("scan_kernel")) return rc;
  }
  launch_pdl(nms_segment_kernel<kFromPred>, dim3(w.nc, B), kNmsThreads, 0, st, s, w, nms_thres, strategy);
  if (int rc = count_launch("nms_segment_kernel")) return rc;
  {
    int gx = (w.cap + 255) / 256;
    if (gx > 256) gx = 256;
    launch_pdl(rank_gather_kernel<kFromPred>, dim3(gx, B), 256, 0, st, s, w, max_det, det, det_count, keep_index);
    if (int rc = count_launch("rank_gather_kernel")) return rc;
  }
  return 0;
}

}  // namespace glsdet

using namespace glsdet;

struct glsdet_nms {
  int B, A, nc, max_det;
  Work w;
};

extern "C" int64_t glsdet_nms_workspace_bytes(int32_t batch, int32_t anchors, int32_t num_classes) {
  if (batch <= 0 || anchors <= 0 || num_classes <= 0) return -1;
  return workspace_bytes(batch, anchors, num_classes);
}

extern "C" int glsdet_nms_create(int32_t batch, int32_t anchors, int32_t num_classes, int32_t max_det,
                                 void* workspace, int64_t workspace_bytes_given, glsdet_nms_t** op) {
  GLSDET_REQUIRE(op != nullptr, "nms_create: null output handle");
  *op = nullptr;
  GLSDET_REQUIRE(batch > 0 && anchors > 0 && max_det > 0, "nms_create: bad sizes");
  GLSDET_REQUIRE(num_classes > 0 && num_classes <= kMaxClasses, "nms_create: 1..%d classes supported", kMaxClasses);
  GLSDET_REQUIRE(anchors < (1 << 24), "nms_create: at most 2^24-1 anchors per image");
  GLSDET_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
                 "nms_create: workspace must be 256-byte aligned");
  GLSDET_REQUIRE(workspace_bytes_given >= workspace_bytes(batch, anchors, num_classes), "nms_create: workspace too small");
  glsdet_nms* h = new (std::nothrow) glsdet_nms();
  GLSDET_REQUIRE(h != nullptr, "nms_create: out of host memory");
  h->B = batch; h->A = anchors; h->nc = num_classes; h->max_det = max_det;
  h->w = make_work(workspace, batch, anchors, num_classes);
  if (max_det <= kKeptSmem && max_det < anchors && getenv("GLSDET_NMS_NO_TOPK") == nullptr) h->w.topk = max_det;
  *op = h;
  return 0;
}

extern "C" int glsdet_nms_launch(glsdet_nms_t* op, const float* pred, float conf_thres, float nms_thres,
                                 int32_t strategy, float* det, int32_t* det_count, int32_t* keep_index, void* stream) {
  return glsdet_nms_launch_scaled(op, pred, nullptr, conf_thres, nms_thres, strategy, det, det_count, keep_index, stream);
}

extern "C" int glsdet_nms_launch_scaled(glsdet_nms_t* op, const float* pred, const float* box_div, float conf_thres,
                                        float nms_thres, int32_t strategy, float* det, int32_t* det_count,
                                        int32_t* keep_index, void* stream) {
  return glsdet_nms_launch_layout(op, pred, GLSDET_PRED_ROWS, box_div, conf_thres, nms_thres, strategy, det, det_count,
                                  keep_index, stream);
}

extern "C" int glsdet_nms_launch_layout(glsdet_nms_t* op, const float* pred, int32_t layout, const float* box_div,
                                        float conf_thres, float nms_thres, int32_t strategy, float* det,
                                        int32_t* det_count, int32_t* keep_index, void* stream) {
  GLSDET_REQUIRE(op && pred && det && det_count, "nms_launch: null pointer");
  GLSDET_REQUIRE(strategy >= 0 && strategy <= GLSDET_NMS_MMCV, "nms_launch: bad strategy %d", strategy);
  GLSDET_REQUIRE(layout >= 0 && layout <= (GLSDET_PRED_PLANES | GLSDET_PRED_CLS_LOGITS), "nms_launch: bad prediction layout %d", layout);
  Source s{};
  s.pred = pred; s.A = op->A; s.nch = 5 + op->nc; s.nc = op->nc; s.box_div = box_div;
  s.cls_logits = (layout & GLSDET_PRED_CLS_LOGITS) ? 1 : 0;
  s.sa = (layout & GLSDET_PRED_PLANES) ? 1 : s.nch;
  s.sc = (layout & GLSDET_PRED_PLANES) ? s.A : 1;
  return run_pipeline<true>(s, op->w, conf_thres, nms_thres, strategy, op->max_det, det, det_count, keep_index,
                            static_cast<cudaStream_t>(stream));
}

extern "C" void glsdet_nms_destroy(glsdet_nms_t* op) { delete op; }

extern "C" int64_t glsdet_batched_nms_workspace_bytes(int32_t k) {
  if (k <= 0) return 256;
  return workspace_bytes(1, k, kMaxClasses);
}

extern "C" int glsdet_batched_nms(const float* boxes, const float* scores, const float* labels, int32_t k,
                                  float nms_thres, int32_t strategy, void* workspace, int64_t workspace_bytes_given,
                                  int32_t* keep, int32_t* keep_count, void* stream) {
  GLSDET_REQUIRE(keep_count != nullptr, "batched_nms: null keep_count");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (k <= 0) {
    GLSDET_CHECK_CUDA(cudaMemsetAsync(keep_count, 0, sizeof(int32_t), st));
    return 0;
  }
  GLSDET_REQUIRE(boxes && scores && labels && keep, "batched_nms: null pointer");
  GLSDET_REQUIRE(k < (1 << 24), "batched_nms: at most 2^24-1 boxes");
  GLSDET_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "batched_nms: boxes must be 16-byte aligned");
  GLSDET_REQUIRE(strategy >= 0 && strategy <= GLSDET_NMS_MMCV, "batched_nms: bad strategy %d", strategy);
  GLSDET_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
                 "batched_nms: workspace must be 256-byte aligned");
  GLSDET_REQUIRE(workspace_bytes_given >= workspace_bytes(1, k, kMaxClasses), "batched_nms: workspace too small");
  const Work w = make_work(workspace, 1, k, kMaxClasses);
  Source s{};
  s.A = k; s.nch = 0; s.nc = kMaxClasses;
  s.boxes = boxes; s.scores = scores; s.labels = labels;
  return run_pipeline<false>(s, w, 0.0f, nms_thres, strategy, k, nullptr, keep_count, keep, st);
}
