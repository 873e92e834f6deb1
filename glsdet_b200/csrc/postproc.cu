// Post-processing of the GLSDet path on the device, without host synchronisation:
//   score filter (obj * max cls >= conf)  ->  warp-ballot compaction  ->  per-image bitonic sort of
//   (class, score desc, anchor) keys  ->  per-(image, class) greedy NMS  ->  rank of every kept box among all
//   kept boxes of its image (binary searches)  ->  [K,7] rows, score desc.
//
// NMS proper, per (image, class) segment of n score-sorted boxes:
//   1. mask_tiles_kernel  - all 64x64 IoU tiles of the upper triangle, one 64-bit suppression word per (row,
//      column tile), computed by the whole GPU (persistent CTAs pulling tiles from a device-side work list), plus the
//      in-degree of every box (how many better boxes overlap it beyond the threshold);
//   2. resolve_kernel     - greedy NMS as a propagation over that sparse DAG, one CTA per segment: a box is KEPT once
//      all better overlapping boxes are known to be suppressed (in-degree counted down to zero) and SUPPRESSED as soon as
//      one kept box overlaps it.  Kept boxes mark their targets removed, suppressed boxes release theirs; every round
//      handles the whole frontier in parallel, and the number of rounds is the longest kept/suppressed alternation in a
//      cluster of boxes (a handful), not the number of 64-box chunks of the segment.  Same keep set as the serial walk.
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

#include <mutex>
#include <new>
#include <utility>
#include <vector>

#include "../../include/glsdet_b200.h"
#include "common.h"

namespace glsdet {

constexpr int kSortMin = 1024;        // smallest sort extent (the key array is padded to a power of two)
constexpr int kSortSmemKeys = 16384;  // keys one CTA sorts entirely in shared memory (128 KB)
constexpr int kSortThreads = 1024;
constexpr int kNmsThreads = 1024;   // 8 warps per scheduler: the kept-list sweep of a chunk is latency-bound
constexpr int kKeptSmem = 1536;     // kept boxes cached in shared memory per segment
constexpr int kMaxClasses = 256;    // 8 label bits in the sort key
constexpr uint64_t kPadKey = ~0ull;
constexpr int kGrid = 32;             // spatial grid per segment (kGrid x kGrid cells) for the tile culling
constexpr int kSizeClasses = 4;       // boxes are first split by size (a few large boxes would blow up every chunk's bounds)
constexpr int kCells = kSizeClasses * kGrid * kGrid;

__device__ __forceinline__ uint32_t float_order_bits(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
// key = [63:56] label (0 when the image runs class-agnostic) | [55:24] ~(order-preserving score bits) | [23:0] index.
// float_order_bits is monotonic over ALL floats (caller-supplied scores may be negative, e.g. raw logits), so ascending
// keys = descending scores, ties by ascending index - the stable descending sort of torchvision.
__device__ __forceinline__ uint64_t make_key(uint32_t label, float score, uint32_t idx) {
  return (static_cast<uint64_t>(label) << 56) | (static_cast<uint64_t>(~float_order_bits(score)) << 24) | idx;
}
__device__ __forceinline__ uint32_t key_idx(uint64_t k) { return static_cast<uint32_t>(k & 0xFFFFFFu); }
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
  const int32_t* label_ids;  // optional dense class ids 0..255 of the caller-supplied labels (sort key / segments); the
                             // coordinate-trick offsets always use the original label values
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
      // window below the largest logit m inside which another class can round to the same fp32 probability: the sigmoid's
      // slope is e^-m for large m while one ulp of a probability near 1 is 2^-24, so the window grows like 2^-23 e^m;
      // from m = 9 on (window 1e-3) every logit above 9 is evaluated, below that 1e-3 * max(1, |m|) covers it
      const float lim = m - 1e-3f * fmaxf(1.0f, fabsf(m));
      const bool all = (m < -80.0f);
      best = -1.0f;
      for (int k = 0; k < s.nc; ++k) {
        const float x = __ldg(r + (5 + k) * sc);
        if (all || x >= lim || x > 9.0f) {
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
    const int64_t o = static_cast<int64_t>(b) * s.A + idx;    // caller-supplied arrays are [B][A] (B = 1 for one image)
    c.box = __ldg(reinterpret_cast<const float4*>(s.boxes) + o);
    c.obj = __ldg(s.scores + o);
    c.cls_conf = 1.0f;
    c.label_f = __ldg(s.labels + o);
    c.label = s.label_ids ? __ldg(s.label_ids + o) : static_cast<int>(c.label_f);
  }
  return c;
}

struct Work {
  int B, cap, P, nc;
  int32_t* cand_count;   // [B]
  uint32_t* max_bits;    // [B]
  uint32_t* min_bits;    // [B]
  int32_t* cls_count;    // [B][nc]    candidates per class (score filter)
  int32_t* cls_cursor;   // [B][nc]    scatter cursors of build_keys_kernel
  int32_t* seg_start;    // [B][nc+1]
  int32_t* seg_kept;     // [B][nc]
  float* cand_score;     // [B][cap]
  int32_t* cand_idx;     // [B][cap]
  uint8_t* cand_label;   // [B][cap]
  uint64_t* keys;        // [B][P]     sorted keys (class segments, score desc inside)
  uint64_t* keys_raw;    // [B][P]     keys in arrival order per class, then sorted runs (input of the rank merge)
  uint64_t* kept_key;    // [B][cap]
  float4* kept_box;      // [B][cap]   spill of the kept list beyond shared memory (fallback kernel)
  float4* nbox;          // [B][cap]   NMS-space box of every sorted candidate (offset applied for the trick)
  int32_t* seg_tile_off; // [B*nc + 1] exclusive prefix of per-segment tile counts (mask path)
  int64_t* seg_word_off; // [B*nc]     first mask word of the segment, or -1 when it runs on the fallback kernel
  int32_t* pending;      // [B][cap]   in-degree of every sorted candidate: better boxes of its segment that overlap it beyond thr
  unsigned long long* row_tiles;  // [B][cap]  per sorted candidate: which column tiles of its mask row hold suppression bits
                                  // (bit k covers column tiles [k << shift, (k + 1) << shift), shift = tile_shift(T));
                                  // zero = the box suppresses nothing
  int32_t* cell_hist;        // [B*nc][kCells]  counting sort of a segment's boxes by grid cell: counts, then cursors
  uint16_t* cell_id;         // [B][cap]        grid cell of every sorted candidate
  int32_t* srank;            // [B][cap]        spatial order: position -> index of the box in the score order (image-wide)
  float4* sbox;              // [B][cap]        NMS-space boxes in spatial order
  unsigned long long* tile_list;  // [max_tiles]   (segment << 40 | row chunk << 20 | column chunk) of every IoU tile to compute
  int32_t* counters;         // [0] tiles listed, [2..3] (int64) mask words in use
  int32_t max_tiles;
  unsigned long long* mask;  // [mask_words]
  int64_t mask_words;
  float trick_label_max;     // largest |label| (labels integral), or < 0: the coordinate trick may only be decomposed by
                             // class when distinct labels are at least 1 apart and every offset stays below 2^21
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
  m.per_class = !trick || (n > 0 && minc >= -0.5f && w.trick_label_max >= 0.0f &&
                           (w.trick_label_max + 1.0f) * m.offset_scale < 2097152.0f);
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
  if (i < w.B * w.nc) {
    w.cls_count[i] = 0;
    w.cls_cursor[i] = 0;
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
  __shared__ int s_cls[kMaxClasses];   // candidates of this CTA per class: one global atomic per class and CTA
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < w.nc; i += blockDim.x) s_cls[i] = 0;
  __syncthreads();
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
      atomicAdd(&s_cls[c.label & (kMaxClasses - 1)], 1);
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
  for (int i = threadIdx.x; i < w.nc; i += blockDim.x)
    if (s_cls[i]) atomicAdd(&w.cls_count[b * w.nc + i], s_cls[i]);
}

__device__ __forceinline__ int sort_extent(int n) {
  int p = kSortMin;
  while (p < n) p <<= 1;
  return p;
}

// Class segments of an image: [seg_start[c], seg_start[c+1]) of its sorted key array.  Known from the class counts of the
// score filter, so the keys can be written class by class right away and every segment sorted on its own.
__global__ void segment_bounds_kernel(Work w, int strategy) {
  pdl_prologue();
  const int b = blockIdx.x;
  if (threadIdx.x != 0) return;
  const int n = w.cand_count[b];
  const ImageMode m = image_mode(w, b, strategy);
  int32_t* st = w.seg_start + b * (w.nc + 1);
  int pos = 0;
  for (int c = 0; c < w.nc; ++c) {
    st[c] = m.per_class ? pos : (c == 0 ? 0 : n);
    pos += w.cls_count[b * w.nc + c];
  }
  st[w.nc] = n;
}

__global__ void __launch_bounds__(256) build_keys_kernel(Work w, int strategy) {
  pdl_prologue();
  const int b = blockIdx.y;
  const int n = w.cand_count[b];
  const ImageMode m = image_mode(w, b, strategy);
  const int32_t* st = w.seg_start + b * (w.nc + 1);
  // whole warps iterate (the bound is rounded up to the warp) so that the lanes of a warp holding the same class can share
  // ONE atomicAdd on its cursor: one class usually holds most candidates of an image, and a per-candidate atomic on that
  // single address serialised ~11 k operations per image at the L2
  const int lane = threadIdx.x & 31;
  const int n_warp = (n + 31) & ~31;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_warp; i += gridDim.x * blockDim.x) {
    const bool live = i < n;
    const int64_t o = static_cast<int64_t>(b) * w.cap + (live ? i : 0);
    const uint32_t label = (live && m.per_class) ? w.cand_label[o] : 0u;
    const uint32_t active = __ballot_sync(0xffffffffu, live);
    if (live) {
      // position inside the class segment in arrival order (the segment is sorted afterwards)
      const uint32_t peers = __match_any_sync(active, label);
      const int leader = __ffs(peers) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(&w.cls_cursor[b * w.nc + label], __popc(peers));
      base = __shfl_sync(peers, base, leader);
      const int pos = st[label] + base + __popc(peers & ((1u << lane) - 1u));
      w.keys_raw[static_cast<int64_t>(b) * w.P + pos] = make_key(label, w.cand_score[o], static_cast<uint32_t>(w.cand_idx[o]));
      w.pending[static_cast<int64_t>(b) * w.cap + i] = 0;
      w.row_tiles[static_cast<int64_t>(b) * w.cap + i] = 0ull;
    }
  }
  int32_t* hist = w.cell_hist + static_cast<int64_t>(b) * w.nc * kCells;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < w.nc * kCells; i += gridDim.x * blockDim.x) hist[i] = 0;
  if (b == 0 && blockIdx.x == 0 && threadIdx.x == 0) w.counters[0] = 0;
}

// ---------------------------------------------------------------------------------------------- sort
__device__ __forceinline__ void cmp_swap(uint64_t& a, uint64_t& b, bool asc) {
  if ((a > b) == asc) { const uint64_t t = a; a = b; b = t; }
}

// Sorting the keys of a class segment (ascending = score desc, index asc) in two kernels:
//   sort_runs_kernel  - the segment is cut into runs of kRunKeys keys; one CTA sorts one run in shared memory
//                       (bitonic network over <= 4096 keys: a quarter of the compare-exchanges per key of a 16 k network,
//                       and the runs of all segments spread over the whole GPU);
//   merge_runs_kernel - keys are unique, so the final position of a key is its index in its own run plus the number of
//                       smaller keys in every other run of the segment: a few binary searches per key, no merge passes.
// One class usually holds most candidates of an image (11 k of 11.5 k), which made a one-CTA-per-segment bitonic sort as
// slow as sorting the whole image (190 us); runs + rank merge take ~40 us on the same load and scale to any segment size.
constexpr int kRunKeys = 4096;
constexpr int kRunThreads = 1024;

// run index -> (segment, run inside the segment); returns false when the image has fewer runs
__device__ __forceinline__ bool locate_run(const Work& w, int b, int run, int* seg, int* local) {
  const int32_t* st = w.seg_start + b * (w.nc + 1);
  int acc = 0;
  for (int c = 0; c < w.nc; ++c) {
    const int r = (st[c + 1] - st[c] + kRunKeys - 1) / kRunKeys;
    if (run < acc + r) { *seg = c; *local = run - acc; return true; }
    acc += r;
  }
  return false;
}

__global__ void __launch_bounds__(kRunThreads) sort_runs_kernel(Work w) {
  pdl_prologue();
  __shared__ uint64_t sk[kRunKeys];
  const int b = blockIdx.y;
  int seg, local;
  if (!locate_run(w, b, blockIdx.x, &seg, &local)) return;
  const int s0 = w.seg_start[b * (w.nc + 1) + seg] + local * kRunKeys;
  const int n = min(kRunKeys, w.seg_start[b * (w.nc + 1) + seg + 1] - s0);
  int pb = 2;
  while (pb < n) pb <<= 1;
  uint64_t* g = w.keys_raw + static_cast<int64_t>(b) * w.P + s0;
  for (int i = threadIdx.x; i < pb; i += kRunThreads) sk[i] = i < n ? g[i] : kPadKey;
  __syncthreads();
  for (int k = 2; k <= pb; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < pb / 2; t += kRunThreads) {
        const int i = 2 * j * (t / j) + (t % j);
        cmp_swap(sk[i], sk[i + j], (i & k) == 0);
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < n; i += kRunThreads) g[i] = sk[i];
}

__global__ void __launch_bounds__(256) merge_runs_kernel(Work w) {
  pdl_prologue();
  const int b = blockIdx.y;
  const int n = w.cand_count[b];
  const uint64_t* src = w.keys_raw + static_cast<int64_t>(b) * w.P;
  uint64_t* dst = w.keys + static_cast<int64_t>(b) * w.P;
  const int32_t* st = w.seg_start + b * (w.nc + 1);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint64_t key = src[i];
    int lo_c = 0, hi_c = w.nc;   // class segment of position i: st[lo_c] <= i < st[hi_c]
    while (hi_c - lo_c > 1) {
      const int mid = (lo_c + hi_c) >> 1;
      if (st[mid] <= i) lo_c = mid; else hi_c = mid;
    }
    const int s0 = st[lo_c], s1 = st[lo_c + 1];
    const int own = (i - s0) / kRunKeys;
    int rank = (i - s0) - own * kRunKeys;
    for (int r0 = s0, r = 0; r0 < s1; r0 += kRunKeys, ++r) {
      if (r == own) continue;
      int lo = r0, hi = min(r0 + kRunKeys, s1);   // keys of the other run that are smaller than this key
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (src[mid] < key) lo = mid + 1; else hi = mid;
      }
      rank += lo - r0;
    }
    dst[s0 + rank] = key;
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

// column tiles per bit of a row's tile bitmap: the smallest power of two that maps T tiles onto 64 bits
__device__ __forceinline__ int tile_shift(int T) {
  int s = 0;
  while (((T - 1) >> s) > 63) ++s;
  return s;
}

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
      if (base + 1024 >= total) {
        w.seg_tile_off[total] = ct;
        *reinterpret_cast<long long*>(w.counters + 2) = cw;   // mask words in use (zeroed by cell_count_kernel)
      }
    }
    __syncthreads();
  }
}

// ---- spatial culling of the IoU tiles
// All-pairs IoU inside a class is O(n^2) (330 M pairs per 16-image step at 7 k boxes of the dominant class) although a box
// only overlaps its neighbours.  The boxes of a segment are therefore put into a spatially coherent order - counting
// sort by the grid cell of the box centre, cells in Morton order - and cut into chunks of 64; two chunks are compared only if
// their bounding boxes intersect (exact: boxes that do not intersect cannot exceed an IoU threshold >= 0).  The order
// only decides how many tiles survive, never the result: every surviving tile writes its suppression bits straight into
// the SCORE-ordered triangular mask (row = better box, bit = worse box), so the resolve step is unchanged.
__device__ __forceinline__ int segment_of(const Work& w, int b, int i) {   // class segment of sorted candidate i
  const int32_t* st = w.seg_start + b * (w.nc + 1);
  int lo = 0, hi = w.nc;   // st[lo] <= i < st[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (st[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

// grid cell of every box on the mask path + per-segment cell histogram; also zeroes the mask words in use
__global__ void __launch_bounds__(256) cell_count_kernel(Work w, int strategy) {
  pdl_prologue();
  const int b = blockIdx.y;
  const int n = w.cand_count[b];
  {   // zero the used part of the suppression mask (16-byte stores, all CTAs)
    const long long used = *reinterpret_cast<const long long*>(w.counters + 2);
    uint4* m4 = reinterpret_cast<uint4*>(w.mask);
    const long long n4 = (used + 1) >> 1;
    const long long stride = static_cast<long long>(gridDim.x) * gridDim.y * blockDim.x;
    for (long long i = (static_cast<long long>(b) * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride)
      m4[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (n == 0) return;
  const ImageMode m = image_mode(w, b, strategy);
  // Cells: the raw coordinate range of the image is cut into kGrid x kGrid cells.  Per-class segments of the coordinate
  // trick carry one constant offset each, which only rotates the (periodic) cell index; the literal trick (one segment
  // over all classes) stretches the grid over all offsets instead.  The cell only orders the boxes - the culling itself
  // uses the true chunk bounding boxes - so any assignment is correct.
  const float minc = order_bits_float(w.min_bits[b]), maxc = order_bits_float(w.max_bits[b]);
  const bool literal = m.use_offsets && !m.per_class;
  const float span = fmaxf(maxc - minc, 1e-12f) + (literal ? fmaxf(w.trick_label_max, 0.0f) * m.offset_scale : 0.0f);
  const float sc = static_cast<float>(kGrid) / span;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = segment_of(w, b, i);
    if (w.seg_word_off[b * w.nc + c] < 0) continue;
    const float4 bx = w.nbox[static_cast<int64_t>(b) * w.cap + i];
    float fx = floorf((0.5f * (bx.x + bx.z) - minc) * sc), fy = floorf((0.5f * (bx.y + bx.w) - minc) * sc);
    fx = fminf(fmaxf(fx, -1.0e9f), 1.0e9f);   // NaN -> -1e9 (fmaxf drops the NaN), then an ordinary integer
    fy = fminf(fmaxf(fy, -1.0e9f), 1.0e9f);
    const int cx = static_cast<int>(fx) & (kGrid - 1), cy = static_cast<int>(fy) & (kGrid - 1);
    // size class: 0 = larger than a quarter of the range ... 3 = at most 1/16 of it (NaN sizes land in class 0)
    const float sz = fmaxf(bx.z - bx.x, bx.w - bx.y) * sc;   // in cells
    const int sz_class = sz <= 0.0625f * kGrid ? 3 : sz <= 0.125f * kGrid ? 2 : sz <= 0.25f * kGrid ? 1 : 0;
    // Morton (Z-order) index of the cell instead of row-major: 64 consecutive boxes then cover a compact block of cells
    // rather than a strip of one cell row, and compact chunks have far fewer intersecting chunk bounding boxes
    auto spread = [](uint32_t v) {   // 5 bits -> every other bit
      v = (v | (v << 8)) & 0x00FF00FFu;
      v = (v | (v << 4)) & 0x0F0F0F0Fu;
      v = (v | (v << 2)) & 0x33333333u;
      v = (v | (v << 1)) & 0x55555555u;
      return v;
    };
    const int cid = sz_class * kGrid * kGrid + static_cast<int>(spread(static_cast<uint32_t>(cx)) | (spread(static_cast<uint32_t>(cy)) << 1));
    w.cell_id[static_cast<int64_t>(b) * w.cap + i] = static_cast<uint16_t>(cid);
    atomicAdd(&w.cell_hist[(static_cast<int64_t>(b) * w.nc + c) * kCells + cid], 1);
  }
}

// counts -> first position of every cell (exclusive scan of the kCells counters of a segment; 4 consecutive per thread)
constexpr int kScanThreads = kCells / 4;
__global__ void __launch_bounds__(kScanThreads) cell_scan_kernel(Work w) {
  pdl_prologue();
  __shared__ int wsum[32];
  const int sgm = blockIdx.y * w.nc + blockIdx.x;
  if (w.seg_word_off[sgm] < 0) return;
  int4* h = reinterpret_cast<int4*>(w.cell_hist + static_cast<int64_t>(sgm) * kCells);
  const int tid = threadIdx.x, lane = tid & 31;
  const int4 v = h[tid];
  const int mine = v.x + v.y + v.z + v.w;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) wsum[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    const int x = tid < kScanThreads / 32 ? wsum[tid] : 0;
    int sc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, sc, o);
      if (tid >= o) sc += u;
    }
    wsum[tid] = sc - x;
  }
  __syncthreads();
  const int base = wsum[tid >> 5] + incl - mine;
  h[tid] = make_int4(base, base + v.x, base + v.x + v.y, base + v.x + v.y + v.z);
}

// scatter into spatial order: srank[pos] = index in the score order, sbox[pos] = its NMS-space box
__global__ void __launch_bounds__(256) cell_scatter_kernel(Work w) {
  pdl_prologue();
  const int b = blockIdx.y;
  const int n = w.cand_count[b];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = segment_of(w, b, i);
    if (w.seg_word_off[b * w.nc + c] < 0) continue;
    const int64_t o = static_cast<int64_t>(b) * w.cap;
    const int cid = w.cell_id[o + i];
    const int pos = w.seg_start[b * (w.nc + 1) + c] + atomicAdd(&w.cell_hist[(static_cast<int64_t>(b) * w.nc + c) * kCells + cid], 1);
    w.srank[o + pos] = i;
    w.sbox[o + pos] = w.nbox[o + i];
  }
}

// One CTA per segment: bounding box of every 64-box chunk (spatial order), then the list of chunk pairs (r <= c) whose
// bounding boxes intersect.
constexpr int kTileListThreads = 512;
__global__ void __launch_bounds__(kTileListThreads) tile_list_kernel(Work w, float thr) {
  pdl_prologue();
  extern __shared__ float4 chunk_bb[];   // [T]
  const int seg = blockIdx.x, b = blockIdx.y;
  const int sgm = b * w.nc + seg;
  if (w.seg_word_off[sgm] < 0) return;
  const int s0 = w.seg_start[b * (w.nc + 1) + seg];
  const int n = w.seg_start[b * (w.nc + 1) + seg + 1] - s0;
  const int T = (n + 63) >> 6;
  const float4* sb = w.sbox + static_cast<int64_t>(b) * w.cap + s0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < T; r += kTileListThreads / 32) {
    float x1 = INFINITY, y1 = INFINITY, x2 = -INFINITY, y2 = -INFINITY;
    for (int k = lane; k < 64; k += 32) {
      const int i = r * 64 + k;
      if (i < n) {
        const float4 q = sb[i];   // NaN coordinates drop out of fminf / fmaxf: such a box never intersects anything
        x1 = fminf(x1, q.x); y1 = fminf(y1, q.y); x2 = fmaxf(x2, q.z); y2 = fmaxf(y2, q.w);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      x1 = fminf(x1, __shfl_xor_sync(0xffffffffu, x1, o)); y1 = fminf(y1, __shfl_xor_sync(0xffffffffu, y1, o));
      x2 = fmaxf(x2, __shfl_xor_sync(0xffffffffu, x2, o)); y2 = fmaxf(y2, __shfl_xor_sync(0xffffffffu, y2, o));
    }
    if (lane == 0) chunk_bb[r] = make_float4(x1, y1, x2, y2);
  }
  __syncthreads();
  const bool cull = (thr >= 0.0f);   // a negative threshold suppresses disjoint boxes too: every pair counts
  // warp per row chunk, lanes over the column chunks; one slot range per warp and iteration (ballot + one atomic)
  for (int r = warp; r < T; r += kTileListThreads / 32) {
    const float4 a = chunk_bb[r];
    for (int c0 = r; c0 < T; c0 += 32) {
      const int c = c0 + lane;
      bool hit = c < T;
      if (hit && cull) {
        const float4 q = chunk_bb[c];
        hit = fminf(a.z, q.z) > fmaxf(a.x, q.x) && fminf(a.w, q.w) > fmaxf(a.y, q.y);
      }
      const uint32_t bal = __ballot_sync(0xffffffffu, hit);
      if (bal == 0u) continue;
      int base = 0;
      if (lane == 0) base = atomicAdd(&w.counters[0], __popc(bal));
      base = __shfl_sync(0xffffffffu, base, 0);
      const int slot = base + __popc(bal & ((1u << lane) - 1u));
      if (hit && slot < w.max_tiles)
        w.tile_list[slot] = (static_cast<unsigned long long>(sgm) << 40) | (static_cast<unsigned long long>(r) << 20) |
                            static_cast<unsigned long long>(c);
    }
  }
}

// Persistent CTAs of 4 x 64 threads; every 64-thread group owns one listed tile at a time.  Thread i of the group tests
// row box i (chunk r, spatial order) against the 64 boxes of chunk c staged in shared memory; every pair beyond the
// threshold becomes one bit of the score-ordered mask: row = the better box of the pair, column = the worse one.
constexpr int kMaskGroups = 4;
__global__ void __launch_bounds__(64 * kMaskGroups) mask_tiles_kernel(Work w, float thr) {
  pdl_prologue();
  __shared__ float4 cbox[kMaskGroups][64];
  __shared__ float carea[kMaskGroups][64];
  __shared__ int crank[kMaskGroups][64];
  const int total_tiles = min(w.counters[0], w.max_tiles);
  const bool thr_nonneg = (thr >= 0.0f);
  const int grp = threadIdx.x >> 6, tid = threadIdx.x & 63;
  for (int t = blockIdx.x * kMaskGroups + grp; t < total_tiles; t += gridDim.x * kMaskGroups) {
    const unsigned long long e = w.tile_list[t];
    const int sgm = static_cast<int>(e >> 40), r = static_cast<int>((e >> 20) & 0xFFFFF), cc = static_cast<int>(e & 0xFFFFF);
    const int b = sgm / w.nc, c = sgm - b * w.nc;
    const int s0 = w.seg_start[b * (w.nc + 1) + c];
    const int n = w.seg_start[b * (w.nc + 1) + c + 1] - s0;
    const int T = (n + 63) >> 6;
    const int64_t o = static_cast<int64_t>(b) * w.cap + s0;
    const int cj = cc * 64 + tid;
    if (cj < n) {
      const float4 q = w.sbox[o + cj];
      cbox[grp][tid] = q;
      carea[grp][tid] = box_area(q);
      crank[grp][tid] = w.srank[o + cj] - s0;
    }
    asm volatile("bar.sync %0, 64;" ::"r"(grp + 1) : "memory");
    const int ri = r * 64 + tid;
    if (ri < n) {
      const float4 rb = w.sbox[o + ri];
      const float ra = box_area(rb);
      const int rrank = w.srank[o + ri] - s0;
      const int jn = min(64, n - cc * 64);
      unsigned long long m;
      if (thr_nonneg) {
        // phase 1, branch-free: which column boxes intersect this row box at all (inter == 0 never exceeds thr >= 0);
        // phase 2: the exact IEEE IoU only for those
        // min(x2) - max(x1) > 0  <=>  both right edges lie beyond both left edges: four compares chained into one
        // predicate and one predicated OR per pair (a NaN coordinate fails the compares; such a box has a NaN area and
        // can never exceed the threshold, so dropping it here is exact).  The row box's own "x2 > x1 && y2 > y1" part is
        // hoisted out of the loop.
        const bool row_ok = rb.z > rb.x && rb.w > rb.y;
        uint32_t lo = 0u, hi = 0u;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float4 cb = cbox[grp][j];
          if (rb.z > cb.x && cb.z > rb.x && rb.w > cb.y && cb.w > rb.y) lo |= 1u << j;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float4 cb = cbox[grp][32 + j];
          if (rb.z > cb.x && cb.z > rb.x && rb.w > cb.y && cb.w > rb.y) hi |= 1u << j;
        }
        m = row_ok ? ((static_cast<unsigned long long>(hi) << 32) | lo) : 0ull;
      } else {
        m = ~0ull;
      }
      if (jn < 64) m &= (1ull << jn) - 1ull;                             // stale columns of a ragged chunk
      if (cc == r) m &= (tid < 63) ? ~((2ull << tid) - 1ull) : 0ull;     // diagonal tile: every pair once
      while (m) {
        const int j = __ffsll(static_cast<long long>(m)) - 1;
        m &= m - 1ull;
        if (iou_exceeds(rb, ra, cbox[grp][j], carea[grp][j], thr, thr_nonneg)) {
          const int ja = crank[grp][j];
          const int a = min(rrank, ja), d = max(rrank, ja);              // a: better (suppressor), d: worse (suppressed)
          atomicOr(&w.mask[w.seg_word_off[sgm] + tri_word(T, a >> 6, a & 63, d >> 6)], 1ull << (d & 63));
          atomicOr(&w.row_tiles[o + a], 1ull << ((d >> 6) >> tile_shift(T)));
          atomicAdd(&w.pending[o + d], 1);
        }
      }
    }
    asm volatile("bar.sync %0, 64;" ::"r"(grp + 1) : "memory");
  }
}

// One CTA per mask-path segment: greedy NMS as a propagation over the sparse suppression DAG (edges = mask bits, from
// the better box to the worse one; pending = in-degree).
//   kept(j)       <=>  every better neighbour of j is suppressed   (pending[j] counted down to 0)
//   suppressed(j) <=>  some better neighbour of j is kept          (removed bit set by that neighbour)
// Round 0 keeps every box without a better neighbour.  Each later round takes the boxes decided in the previous one
// (only those with outgoing edges are listed): a kept box marks its targets removed, a suppressed box releases its
// targets (pending - 1; at zero the target is kept).  One thread per listed box: its row's tile bitmap names the few
// mask words that hold its edges.  The rounds needed = the longest kept / suppressed alternation among overlapping
// boxes, independent of the segment length.
// warp-aggregated slot allocation: the threads of a warp that reach this call together take one atomicAdd for all of
// them (thousands of pushes onto one shared-memory counter would otherwise serialise)
__device__ __forceinline__ int agg_slot(int* counter) {
  const unsigned m = __activemask();
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(static_cast<int>(m)) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(m, base, leader);
  return base + __popc(m & ((1u << lane) - 1u));
}

constexpr int kResolveThreads = 1024;
constexpr int kResolveSmemN = 24576;   // segments up to this size keep their in-degree counters in shared memory
__global__ void __launch_bounds__(kResolveThreads) resolve_kernel(Work w) {
  pdl_prologue();
  extern __shared__ uint32_t resolve_smem[];  // removed[W] | keptb[W]
  __shared__ int list_n[2];
  __shared__ int warp_tot[kResolveThreads / 32];
  const int seg = blockIdx.x, b = blockIdx.y;
  const int sgm = b * w.nc + seg;
  const int64_t woff = w.seg_word_off[sgm];
  if (woff < 0) return;  // empty, or handled by nms_segment_kernel
  const int s0 = w.seg_start[b * (w.nc + 1) + seg];
  const int n = w.seg_start[b * (w.nc + 1) + seg + 1] - s0;
  const int T = (n + 63) >> 6;
  const int W = (n + 31) >> 5;
  const int shift = tile_shift(T);
  uint32_t* removed = resolve_smem;
  uint32_t* keptb = resolve_smem + W;
  const int tid = threadIdx.x;
  const unsigned long long* mask = w.mask + woff;
  const unsigned long long* row_tiles = w.row_tiles + static_cast<int64_t>(b) * w.cap + s0;
  int32_t* pend = w.pending + static_cast<int64_t>(b) * w.cap + s0;
  if (n <= kResolveSmemN) {   // shared-memory copy: the count-downs are shared-memory atomics instead of L2 round trips
    int32_t* pend_s = reinterpret_cast<int32_t*>(resolve_smem + 2 * W);
    for (int j = threadIdx.x; j < n; j += kResolveThreads) pend_s[j] = pend[j];
    pend = pend_s;
  }
  // the kept-box spill area (16 bytes per box) is unused on the mask path: two frontier lists of n entries each
  int32_t* lists = reinterpret_cast<int32_t*>(w.kept_box + static_cast<int64_t>(b) * w.cap + s0);
  const uint64_t* keys = w.keys + static_cast<int64_t>(b) * w.P + s0;
  uint64_t* gk_key = w.kept_key + static_cast<int64_t>(b) * w.cap + s0;
  for (int i = tid; i < 2 * W; i += kResolveThreads) resolve_smem[i] = 0u;
  if (tid < 2) list_n[tid] = 0;
  __syncthreads();
  // round 0: boxes nobody better overlaps
  for (int j = tid; j < n; j += kResolveThreads) {
    if (pend[j] == 0) {
      atomicOr(&keptb[j >> 5], 1u << (j & 31));
      if (row_tiles[j] != 0ull) lists[agg_slot(&list_n[0])] = j;
    }
  }
  __syncthreads();
  int cur = 0;
  while (true) {
    const int cnt = list_n[cur];
    if (cnt == 0) break;
    const int nxt = cur ^ 1;
    __syncthreads();               // everybody has read list_n[cur] ...
    if (tid == 0) list_n[nxt] = 0; // ... and nobody pushes before the next barrier
    __syncthreads();
    const int32_t* in = lists + cur * n;
    int32_t* out = lists + nxt * n;
    for (int e = tid; e < cnt; e += kResolveThreads) {
      const int ent = in[e];
      const int i = ent & 0x7FFFFFFF;
      const bool kept_i = ent >= 0;
      const int r = i >> 6;
      const int64_t rbase = tri_tile(T, r, r) * 64 + (i & 63);   // word of column tile c: rbase + (c - r) * 64
      unsigned long long tw = row_tiles[i];
      while (tw) {
        const int k = __ffsll(static_cast<long long>(tw)) - 1;
        tw &= tw - 1ull;
        const int c_lo = max(k << shift, r), c_hi = min((k + 1) << shift, T);
        for (int c = c_lo; c < c_hi; ++c) {
          unsigned long long word = mask[rbase + static_cast<int64_t>(c - r) * 64];
          while (word) {
            const int jb = __ffsll(static_cast<long long>(word)) - 1;
            word &= word - 1ull;
            const int j = c * 64 + jb;
            const uint32_t bit = 1u << (j & 31);
            if (kept_i) {
              const uint32_t old = atomicOr(&removed[j >> 5], bit);
              if (!(old & bit) && row_tiles[j] != 0ull) out[agg_slot(&list_n[nxt])] = j | static_cast<int>(0x80000000u);
            } else if (!(removed[j >> 5] & bit)) {
              // a kept better neighbour never releases, so pending reaches zero only if all of them are suppressed
              if (atomicSub(&pend[j], 1) == 1) {
                atomicOr(&keptb[j >> 5], bit);
                if (row_tiles[j] != 0ull) out[agg_slot(&list_n[nxt])] = j;
              }
            }
          }
        }
      }
    }
    __syncthreads();
    cur = nxt;
  }
  // kept boxes in score order: exclusive prefix of the per-word popcounts (contiguous words per thread)
  const int per = (W + kResolveThreads - 1) / kResolveThreads;
  const int w_lo = min(tid * per, W), w_hi = min(w_lo + per, W);
  int mine = 0;
  for (int i = w_lo; i < w_hi; ++i) mine += __popc(keptb[i]);
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if ((tid & 31) >= o) incl += v;
  }
  if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    int v = tid < kResolveThreads / 32 ? warp_tot[tid] : 0;
    int sc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, sc, o);
      if (tid >= o) sc += u;
    }
    if (tid < kResolveThreads / 32) warp_tot[tid] = sc - v;   // exclusive
    if (tid == 31) w.seg_kept[sgm] = sc;
  }
  __syncthreads();
  int pos = warp_tot[tid >> 5] + incl - mine;
  for (int i = w_lo; i < w_hi; ++i) {
    uint32_t bits = keptb[i];
    while (bits) {
      const int jb = __ffs(static_cast<int>(bits)) - 1;
      bits &= bits - 1u;
      gk_key[pos++] = keys[i * 32 + jb] & 0x00FFFFFFFFFFFFFFull;
    }
  }
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
  int64_t off_count, off_max, off_min, off_clscount, off_clscursor, off_segstart, off_segkept, off_score, off_idx, off_label, off_keys, off_keysraw,
      off_keptkey, off_keptbox, off_nbox, off_tileoff, off_wordoff, off_rowtiles, off_pending, off_hist, off_cellid, off_srank, off_sbox, off_tilelist,
      off_counters, off_mask, mask_words, max_tiles, total;
};

constexpr int kMaxScanTiles = 2816;  // 2 x 22 KB of shared memory for the 'removed' and 'kept' bitmaps of a segment

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
  l.off_clscount = o; o = align_up(o + 4ll * B * nc, 256);
  l.off_clscursor = o; o = align_up(o + 4ll * B * nc, 256);
  l.off_segstart = o; o = align_up(o + 4ll * B * (nc + 1), 256);
  l.off_segkept = o; o = align_up(o + 4ll * B * nc, 256);
  l.off_score = o; o = align_up(o + 4ll * B * cap, 256);
  l.off_idx = o; o = align_up(o + 4ll * B * cap, 256);
  l.off_label = o; o = align_up(o + 1ll * B * cap, 256);
  l.off_keys = o; o = align_up(o + 8ll * B * P, 256);
  l.off_keysraw = o; o = align_up(o + 8ll * B * P, 256);
  l.off_keptkey = o; o = align_up(o + 8ll * B * cap, 256);
  l.off_keptbox = o; o = align_up(o + 16ll * B * cap, 256);
  l.off_nbox = o; o = align_up(o + 16ll * B * cap, 256);
  l.off_tileoff = o; o = align_up(o + 4ll * (static_cast<int64_t>(B) * nc + 1), 256);
  l.off_wordoff = o; o = align_up(o + 8ll * B * nc, 256);
  l.off_rowtiles = o; o = align_up(o + 8ll * B * cap, 256);
  l.off_pending = o; o = align_up(o + 4ll * B * cap, 256);
  l.off_hist = o; o = align_up(o + 4ll * B * nc * kCells, 256);
  l.off_cellid = o; o = align_up(o + 2ll * B * cap, 256);
  l.off_srank = o; o = align_up(o + 4ll * B * cap, 256);
  l.off_sbox = o; o = align_up(o + 16ll * B * cap, 256);
  l.off_counters = o; o = align_up(o + 64, 256);
  l.mask_words = mask_budget_words(B, cap);
  l.max_tiles = l.mask_words / 64;
  l.off_tilelist = o; o = align_up(o + 8ll * l.max_tiles, 256);
  l.off_mask = o; o = align_up(o + 8ll * l.mask_words, 256);
  l.total = o;
  return l;
}

Work make_work(void* ws, int B, int cap, int nc) {
  const int P = pow2_ceil(cap < kSortMin ? kSortMin : cap);
  const Layout l = make_layout(B, cap, P, nc);
  uint8_t* p = static_cast<uint8_t*>(ws);
  Work w;
  w.B = B; w.cap = cap; w.P = P; w.nc = nc;
  w.cand_count = reinterpret_cast<int32_t*>(p + l.off_count);
  w.max_bits = reinterpret_cast<uint32_t*>(p + l.off_max);
  w.min_bits = reinterpret_cast<uint32_t*>(p + l.off_min);
  w.cls_count = reinterpret_cast<int32_t*>(p + l.off_clscount);
  w.cls_cursor = reinterpret_cast<int32_t*>(p + l.off_clscursor);
  w.seg_start = reinterpret_cast<int32_t*>(p + l.off_segstart);
  w.seg_kept = reinterpret_cast<int32_t*>(p + l.off_segkept);
  w.cand_score = reinterpret_cast<float*>(p + l.off_score);
  w.cand_idx = reinterpret_cast<int32_t*>(p + l.off_idx);
  w.cand_label = p + l.off_label;
  w.keys = reinterpret_cast<uint64_t*>(p + l.off_keys);
  w.keys_raw = reinterpret_cast<uint64_t*>(p + l.off_keysraw);
  w.kept_key = reinterpret_cast<uint64_t*>(p + l.off_keptkey);
  w.kept_box = reinterpret_cast<float4*>(p + l.off_keptbox);
  w.nbox = reinterpret_cast<float4*>(p + l.off_nbox);
  w.seg_tile_off = reinterpret_cast<int32_t*>(p + l.off_tileoff);
  w.seg_word_off = reinterpret_cast<int64_t*>(p + l.off_wordoff);
  w.row_tiles = reinterpret_cast<unsigned long long*>(p + l.off_rowtiles);
  w.pending = reinterpret_cast<int32_t*>(p + l.off_pending);
  w.cell_hist = reinterpret_cast<int32_t*>(p + l.off_hist);
  w.cell_id = reinterpret_cast<uint16_t*>(p + l.off_cellid);
  w.srank = reinterpret_cast<int32_t*>(p + l.off_srank);
  w.sbox = reinterpret_cast<float4*>(p + l.off_sbox);
  w.tile_list = reinterpret_cast<unsigned long long*>(p + l.off_tilelist);
  w.counters = reinterpret_cast<int32_t*>(p + l.off_counters);
  w.max_tiles = static_cast<int32_t>(l.max_tiles < INT32_MAX ? l.max_tiles : INT32_MAX);
  w.mask = reinterpret_cast<unsigned long long*>(p + l.off_mask);
  w.mask_words = l.mask_words;
  if (const char* e = getenv("GLSDET_NMS_MASK_WORDS")) {  // tests: shrink the budget to force the fallback kernel
    const long long v = atoll(e);
    if (v >= 0 && v < w.mask_words) w.mask_words = v;
  }
  w.max_scan_tiles = kMaxScanTiles;
  w.topk = 0;
  w.trick_label_max = static_cast<float>(nc - 1);
  return w;
}

int64_t workspace_bytes(int B, int cap, int nc) {
  const int P = pow2_ceil(cap < kSortMin ? kSortMin : cap);
  return make_layout(B, cap, P, nc).total;
}

// opt-in to more than 48 KB of dynamic shared memory, once per device and kernel
int ensure_smem_attr(const void* kernel, int bytes) {
  static std::mutex mu;
  static std::vector<std::pair<const void*, int>> done;
  int dev = 0;
  GLSDET_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  for (const auto& d : done)
    if (d.first == kernel && d.second == dev) return 0;
  GLSDET_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done.emplace_back(kernel, dev);
  return 0;
}

template <bool kFromPred>
int run_pipeline(const Source& s, const Work& w, float conf_thres, float nms_thres, int strategy, int max_det,
                 float* det, int32_t* det_count, int32_t* keep_index, cudaStream_t st) {
  const int B = w.B;
  launch_pdl(reset_kernel, (B * w.nc + 255) / 256, 256, 0, st, w);
  if (int rc = count_launch("reset_kernel")) return rc;
  {
    int gx = (s.A + 255) / 256;
    if (gx > 4096) gx = 4096;
    launch_pdl(filter_kernel<kFromPred>, dim3(gx, B), 256, 0, st, s, w, conf_thres);
    if (int rc = count_launch("filter_kernel")) return rc;
  }
  launch_pdl(segment_bounds_kernel, B, 32, 0, st, w, strategy);
  if (int rc = count_launch("segment_bounds_kernel")) return rc;
  {
    int gx = (w.cap + 255) / 256;
    if (gx > 1024) gx = 1024;
    launch_pdl(build_keys_kernel, dim3(gx, B), 256, 0, st, w, strategy);
    if (int rc = count_launch("build_keys_kernel")) return rc;
  }
  {
    const int runs = (w.cap + kRunKeys - 1) / kRunKeys + w.nc;   // upper bound of the runs of an image
    launch_pdl(sort_runs_kernel, dim3(runs, B), kRunThreads, 0, st, w);
    if (int rc = count_launch("sort_runs_kernel")) return rc;
    int gx = (w.cap + 255) / 256;
    if (gx > 256) gx = 256;
    launch_pdl(merge_runs_kernel, dim3(gx, B), 256, 0, st, w);
    if (int rc = count_launch("merge_runs_kernel")) return rc;
  }
  {
    int gx = (w.cap + 255) / 256;
    if (gx > 1024) gx = 1024;
    launch_pdl(box_prep_kernel<kFromPred>, dim3(gx, B), 256, 0, st, s, w, strategy);
    if (int rc = count_launch("box_prep_kernel")) return rc;
  }
  launch_pdl(seg_plan_kernel, 1, 1024, 0, st, w);
  if (int rc = count_launch("seg_plan_kernel")) return rc;
  if (w.topk == 0) {
    int gx = (w.cap + 255) / 256;
    if (gx > 256) gx = 256;
    if (gx * B < device_sm_count()) gx = (device_sm_count() + B - 1) / B;   // enough CTAs to zero the mask words
    launch_pdl(cell_count_kernel, dim3(gx, B), 256, 0, st, w, strategy);
    if (int rc = count_launch("cell_count_kernel")) return rc;
    launch_pdl(cell_scan_kernel, dim3(w.nc, B), kScanThreads, 0, st, w);
    if (int rc = count_launch("cell_scan_kernel")) return rc;
    launch_pdl(cell_scatter_kernel, dim3(gx, B), 256, 0, st, w);
    if (int rc = count_launch("cell_scatter_kernel")) return rc;
    int t = (w.cap + 63) / 64;
    if (t > kMaxScanTiles) t = kMaxScanTiles;
    launch_pdl(tile_list_kernel, dim3(w.nc, B), kTileListThreads, static_cast<size_t>(t) * 16, st, w, nms_thres);
    if (int rc = count_launch("tile_list_kernel")) return rc;
    launch_pdl(mask_tiles_kernel, device_sm_count() * 6, 64 * kMaskGroups, 0, st, w, nms_thres);
    if (int rc = count_launch("mask_tiles_kernel")) return rc;
  }
  if (w.topk == 0) {
    int t = (w.cap + 63) / 64;
    if (t > kMaxScanTiles) t = kMaxScanTiles;
    const size_t smem = static_cast<size_t>(t) * 16 + static_cast<size_t>(w.cap < kResolveSmemN ? w.cap : kResolveSmemN) * 4;
    if (int rc = ensure_smem_attr(reinterpret_cast<const void*>(resolve_kernel), kMaxScanTiles * 16 + kResolveSmemN * 4)) return rc;
    launch_pdl(resolve_kernel, dim3(w.nc, B), kResolveThreads, smem, st, w);
    if (int rc = count_launch("resolve_kernel")) return rc;
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

extern "C" int64_t glsdet_batched_nms_batch_workspace_bytes(int32_t batch, int32_t k, int32_t num_ids) {
  if (batch <= 0 || k <= 0 || num_ids <= 0 || num_ids > kMaxClasses) return -1;
  return workspace_bytes(batch, k, num_ids);
}

// B images at once: every array is [batch][k]; image b's kept indices (into its own k candidates, score order) land in
// keep[b][0 .. keep_count[b]).  Each image is an independent NMS problem with the semantics of glsdet_batched_nms_ids
// (the coordinate-trick maximum is per image), so the result equals `batch` separate calls - in one pipeline launch
// sequence instead of `batch` of them.
extern "C" int glsdet_batched_nms_ids_batch(const float* boxes, const float* scores, const float* labels,
                                            const int32_t* label_ids, float label_abs_max, int32_t num_ids, int32_t batch,
                                            int32_t k, float nms_thres, int32_t strategy, void* workspace,
                                            int64_t workspace_bytes_given, int32_t* keep, int32_t* keep_count, void* stream) {
  GLSDET_REQUIRE(boxes && scores && labels && label_ids && keep && keep_count, "batched_nms_batch: null pointer");
  GLSDET_REQUIRE(batch > 0 && k > 0 && k < (1 << 24), "batched_nms_batch: bad sizes");
  GLSDET_REQUIRE(num_ids > 0 && num_ids <= kMaxClasses, "batched_nms_batch: 1..%d class ids", kMaxClasses);
  GLSDET_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "batched_nms_batch: boxes must be 16-byte aligned");
  GLSDET_REQUIRE(strategy >= 0 && strategy <= GLSDET_NMS_MMCV, "batched_nms_batch: bad strategy %d", strategy);
  GLSDET_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
                 "batched_nms_batch: workspace must be 256-byte aligned");
  GLSDET_REQUIRE(workspace_bytes_given >= workspace_bytes(batch, k, num_ids), "batched_nms_batch: workspace too small");
  Work w = make_work(workspace, batch, k, num_ids);
  w.trick_label_max = label_abs_max;
  Source s{};
  s.A = k; s.nch = 0; s.nc = num_ids;
  s.boxes = boxes; s.scores = scores; s.labels = labels; s.label_ids = label_ids;
  return run_pipeline<false>(s, w, 0.0f, nms_thres, strategy, k, nullptr, keep_count, keep, static_cast<cudaStream_t>(stream));
}

extern "C" int glsdet_batched_nms(const float* boxes, const float* scores, const float* labels, int32_t k,
                                  float nms_thres, int32_t strategy, void* workspace, int64_t workspace_bytes_given,
                                  int32_t* keep, int32_t* keep_count, void* stream) {
  return glsdet_batched_nms_ids(boxes, scores, labels, nullptr, 255.0f, k, nms_thres, strategy, workspace,
                                workspace_bytes_given, keep, keep_count, stream);
}

extern "C" int glsdet_batched_nms_ids(const float* boxes, const float* scores, const float* labels, const int32_t* label_ids,
                                      float label_abs_max, int32_t k, float nms_thres, int32_t strategy, void* workspace,
                                      int64_t workspace_bytes_given, int32_t* keep, int32_t* keep_count, void* stream) {
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
  Work w = make_work(workspace, 1, k, kMaxClasses);
  w.trick_label_max = label_abs_max;
  Source s{};
  s.A = k; s.nch = 0; s.nc = kMaxClasses;
  s.boxes = boxes; s.scores = scores; s.labels = labels; s.label_ids = label_ids;
  return run_pipeline<false>(s, w, 0.0f, nms_thres, strategy, k, nullptr, keep_count, keep, st);
}
