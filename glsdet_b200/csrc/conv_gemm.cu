// Implicit-GEMM convolution for sm_100a: TMA-fed, tcgen05.mma with the accumulator in TMEM,
// fused bias / residual / activation / decode epilogue.
//
//   out[pixel, n] = epilogue( sum_{src, tap, c} A_src[pixel + tap, c] * W[n, (src, tap, c)] )
//
// GEMM view: M = output pixels (a CTA tile is a tile_h x tile_w rectangle of 128 pixels of one image),
// N = output channels, K = taps * input channels.  Activations are NHWC bf16, so for one tap the A tile
// is a dense box of the input tensor shifted by (ky-1, kx-1): a single 5-D TMA box load whose
// out-of-bounds elements are zero-filled by the hardware (that IS the conv padding).  Stride-2 3x3 convs
// view the input as [B, H/2, 2, W/2, (2), C] so that each tap is again a dense box.  A second source
// tensor gives torch.cat((a, b), 1) in front of the conv for free (its K range simply follows).
//
// 3x3 stride-1 convs reuse A vertically: one TMA box of (tile_h + 2) x tile_w pixels per (kx, channel chunk) serves
// the three ky taps - the MMA descriptor simply starts tile_w rows (a multiple of the 1024-byte swizzle atom)
// further down - which cuts the A traffic from L2 2.5x.  A and B therefore live in separate smem rings.
//
// Warp roles (384 threads, persistent CTAs, static round-robin tile schedule):
//   warp 0 lane 0 : TMA producer   (A box + weight box per 64-wide K chunk -> smem ring, mbarrier tx)
//   warp 1 lane 0 : MMA issuer     (4 x tcgen05.mma M128 x N x K16 per chunk; commit frees the smem slot)
//   warp 2        : TMEM allocator (2 accumulator stages so the epilogue of tile i overlaps tile i+1)
//   warps 4..11   : epilogue       (tcgen05.ld 32 lanes x 16 columns -> registers -> global); a warp may only
//                                   read TMEM lanes 32*(warp%4)..+31, so two warps share each lane quarter and
//                                   split the accumulator columns
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <stdlib.h>

#include <new>

#include "../../include/glsdet_b200.h"
#include "common.h"
#include "ptx.cuh"

namespace glsdet {

constexpr int kBlockM = 128;
constexpr int kChunkK = 64;                 // bf16 elements per K chunk = one 128-byte swizzle row
constexpr int kRowBytes = kChunkK * 2;       // one pixel row of an A/B stage: 128 bytes
constexpr int kMaxStages = 8;
constexpr int kThreads = 384;            // 4 control warps + 8 epilogue warps
constexpr int kEpiWarps = 8;
constexpr int kSmemLimit = 232448;          // 227 KB opt-in limit per CTA

struct alignas(64) ConvKParams {
  CUtensorMap tmA[2];
  CUtensorMap tmB;
  CUtensorMap tmB3[2];       // per source: 3-D view (k, n, ky) of the weights, one box = the three ky taps
  CUtensorMap tmO;           // TMA-store epilogue: the bf16 NHWC output window, box = one 128-pixel x 64-channel tile
  CUtensorMap tmRpost;       // res_tma: the post-activation residual (16-bit), box = 64 channels x the tile's pixels >> post_shift
  CUtensorMap tmRpre;        // res_tma: the fp32 pre-activation residual viewed as 2-byte elements, box = 32 floats x pixels >> pre_shift
  int32_t B, Ho, Wo;
  int32_t tile_w_log2, tile_h;
  int32_t tiles_x, tiles_y, n_blocks, total_tiles;
  int32_t m_tiles, total_pairs;   // 2-CTA mode: a CTA pair takes M tiles (2i, 2i+1) of one N block
  int32_t block_n, N;
  int32_t taps, stride;
  int32_t ksz;               // kernel height (1, 3, 5, 7)
  int32_t kw;                // kernel width: ksz, or 1 when the kx taps are folded into the channel view; taps = ksz * kw
  int32_t chunks0, chunks1;
  int32_t sa, sb;            // A / B ring depths
  int32_t nsub;              // MMA sub-steps per A step: 3 (ky taps of a 3x3 stride-1 conv) or 1
  int32_t bgroup;            // sub-steps served by one B stage: 3 (one 3-D box, N <= 128) or 1
  int32_t a_steps;           // A steps per tile
  int32_t a_bytes;           // bytes of one A stage
  int32_t tmem_cols;
  int32_t mt;                // M tiles (vertically adjacent, one A box) per work item: every B stage feeds mt MMAs
  int32_t nacc;              // accumulator ring depth in work items: 2; 1 when 2 * mt * block_n > 512 columns; 3 for the
                             // prediction-MMA class (its epilogue waits for a tcgen05.mma that queues behind the main-loop
                             // MMAs already in flight, so the main loop needs to run two items ahead)
  int32_t bres;              // all weight chunks stay resident in shared memory (loaded once per CTA)
  int32_t b_chunks;          // number of 64-wide K chunks of the weight matrix
  int32_t pdl;               // launched with programmatic stream serialisation
  int32_t ts;                // bf16 epilogues stage the tile in shared memory and TMA-store it
  int32_t w_batched;         // one weight matrix per image (third coordinate of tmB)
  int32_t a_shared;          // k > 0: image b reads activation image b % k (static matrices used as activations)
  int32_t a_div;             // with a_shared: image b reads activation image b / a_div instead of b % a_shared
  int32_t patch;             // src0 / post_res / out are 2x2 patch views (image b' = (b*2 + py)*2 + px)
  int32_t res_tma;           // residual operands of the TMA-store epilogue staged through shared memory by TMA (see epi_tile_ts)
  int32_t res_post_bytes, res_pre_half_bytes, res_slot_bytes;   // per staging sub-group: post tile | pre channels 0..31 | 32..63
  int32_t res_nbuf;          // residual slots per sub-group (1)
  int32_t egrp;              // 16-warp TMA-store class: two groups of 8 epilogue warps drain ALTERNATE work items (one
                             // accumulator slot each) instead of all 16 sharing every item - the epilogue of a small-K
                             // conv is a latency chain per tile, two tiles in flight nearly double its throughput
  const float* bias;
  int32_t act;
  int32_t act_epi;           // activation code of the fast bf16 epilogues (act, or kActSiluExact)
  const float* pre_res;
  int32_t pre_shift, pre_ld;
  const __nv_bfloat16* post_res;
  int32_t post_shift, post_ld;
  void* out;
  int32_t out_mode, out_ld, out_coff;
  int64_t out_bs;
  int64_t out_plane;   // NCHW output: elements between channel planes (Ho * Wo, or the anchor count of all levels)
  float dec_stride, dec_in_w, dec_in_h;
  int32_t epi;   // EPI_* epilogue specialisation chosen at create time
  int32_t hbias;     // s_bias holds bias / 2 (fast SiLU epilogues: EPI_BF16* and the prediction-MMA class)
  int32_t a_f16;     // sources, weights (and the staged operand of the prediction MMA) are fp16 instead of bf16
  int32_t out_f16;   // 16-bit NHWC output is fp16
  int32_t post_f16;  // post-activation residual is fp16
  const float* pred_w;   // fused prediction conv: fp32 [pred_n][N]
  const float* pred_b;
  int32_t pred_n, pred_act;
  unsigned long long* trace;   // diagnostic (GLSDET_CONV_TRACE=1): %globaltimer stamps of CTA 0, see glsdet_conv_read_trace
  int32_t dbg;                 // diagnostic (GLSDET_CONV_DBG bits; wrong results): 1 = no TMA store, 2 = no epilogue math / staging writes
};

__device__ __forceinline__ void trace_stamp(const ConvKParams& p, int slot) {
  if (p.trace != nullptr && blockIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[slot] = t;
  }
}

enum { EPI_GENERIC = 0, EPI_BF16 = 1, EPI_BF16_PRE = 2, EPI_BF16_POST = 3, EPI_F32_PLAIN = 4, EPI_NCHW_RAW = 5,
       EPI_ROWS_BOX = 6, EPI_ROWS_SIGMOID = 7, EPI_TOWER_PRED = 8, EPI_TOWER_PRED_MMA = 9, EPI_BF16_PREPOST = 10 };

// Epilogue classes: the kernel is instantiated once per class so that each binary only carries its own epilogue
// (smaller instruction footprint - the all-in-one kernel lost 18 % of its epilogue issue slots to instruction-cache
// misses - and registers sized for that epilogue).  EC_ALL keeps every path (2-CTA kernel, generic fallback).
// EC_BF16_TSR = EC_BF16_TS plus the TMA-staged residual operands: a class of its own, so that the plain TMA-store kernel
// (96 registers with 16 epilogue warps, no headroom) does not pay for that code
enum { EC_ALL = 0, EC_BF16 = 1, EC_BF16_TS = 2, EC_F32 = 3, EC_SMALL = 4, EC_PRED_FMA = 5, EC_PRED_MMA = 6, EC_BF16_TSR = 7, EC_COUNT = 8 };
constexpr int kActSiluExact = 100;   // diagnostic (GLSDET_CONV_EXACT_SILU=1): ex2 + rcp SiLU instead of tanh.approx
constexpr int kStageTileBytes = kBlockM * kRowBytes;   // one 128-row x 64-channel K-major SW128 operand tile
constexpr int kPredTileBytes = 16 * kRowBytes;         // prediction weights: 16 rows x 64 channels

struct TileCoord {
  int b, y0, x0, n0;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvKParams& p, int tile) {
  TileCoord t;
  int nb = tile % p.n_blocks;
  int m = tile / p.n_blocks;
  int tx = m % p.tiles_x;
  m /= p.tiles_x;
  int ty = m % p.tiles_y;
  t.b = m / p.tiles_y;
  t.x0 = tx << p.tile_w_log2;
  t.y0 = ty * p.tile_h * p.mt;
  t.n0 = nb * p.block_n;
  return t;
}

// 2-CTA mode: work item = (pair of adjacent M tiles, N block); CTA `rank` of the pair owns M tile 2*pm + rank.
// An odd tile count leaves one dummy tile: it sits below the image, so TMA zero-fills it and no row is stored.
__device__ __forceinline__ TileCoord decode_pair(const ConvKParams& p, int work, int rank) {
  TileCoord t;
  const int nb = work % p.n_blocks;
  int m = 2 * (work / p.n_blocks) + rank;
  t.n0 = nb * p.block_n;
  if (m >= p.m_tiles) {
    t.b = 0; t.x0 = 0; t.y0 = p.tiles_y * p.tile_h;
    return t;
  }
  const int tx = m % p.tiles_x;
  m /= p.tiles_x;
  const int ty = m % p.tiles_y;
  t.b = m / p.tiles_y;
  t.x0 = tx << p.tile_w_log2;
  t.y0 = ty * p.tile_h;
  return t;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case GLSDET_ACT_SILU: return silu_f(v);  // MUFU.EX2 + MUFU.RCP, ~1e-6 relative
    case GLSDET_ACT_RELU: return fmaxf(v, 0.0f);
    case GLSDET_ACT_LRELU: return v > 0.0f ? v : 0.1f * v;
    case GLSDET_ACT_SIGMOID: return 1.0f / (1.0f + expf(-v));
    default: return v;
  }
}

// 16 fp32 values -> 16 stored 16-bit values (two 16-byte vectors); `f16` is warp-uniform (a kernel parameter)
__device__ __forceinline__ void pack16_row(const float (&v)[16], int f16, uint4& a, uint4& c) {
  if (f16) {
    a.x = pack_f16x2(v[0], v[1]);  a.y = pack_f16x2(v[2], v[3]);
    a.z = pack_f16x2(v[4], v[5]);  a.w = pack_f16x2(v[6], v[7]);
    c.x = pack_f16x2(v[8], v[9]);  c.y = pack_f16x2(v[10], v[11]);
    c.z = pack_f16x2(v[12], v[13]); c.w = pack_f16x2(v[14], v[15]);
  } else {
    a.x = pack_bf16x2(v[0], v[1]);  a.y = pack_bf16x2(v[2], v[3]);
    a.z = pack_bf16x2(v[4], v[5]);  a.w = pack_bf16x2(v[6], v[7]);
    c.x = pack_bf16x2(v[8], v[9]);  c.y = pack_bf16x2(v[10], v[11]);
    c.z = pack_bf16x2(v[12], v[13]); c.w = pack_bf16x2(v[14], v[15]);
  }
}

// Epilogue for 16 consecutive output channels [n_g, n_g+16) of one output pixel.
// s_bias: shared-memory copy of the (zero-padded) bias, already offset to channel n_g.
__device__ __forceinline__ void epilogue_store16(const ConvKParams& p, const uint32_t (&raw)[16],
                                                 const float* s_bias, int b, int oy, int ox, int n_g) {
  float v[16];
  const bool full = (n_g + 16 <= p.N);
#pragma unroll
  for (int j = 0; j < 16; j += 4) {
    const float4 bv = *reinterpret_cast<const float4*>(s_bias + j);  // warp-uniform address: smem broadcast
    v[j] = __uint_as_float(raw[j]) + bv.x;
    v[j + 1] = __uint_as_float(raw[j + 1]) + bv.y;
    v[j + 2] = __uint_as_float(raw[j + 2]) + bv.z;
    v[j + 3] = __uint_as_float(raw[j + 3]) + bv.w;
  }
  if (p.pre_res != nullptr) {
    const int hs = p.Ho >> p.pre_shift, ws = p.Wo >> p.pre_shift;
    const float* r =
        p.pre_res + ((static_cast<int64_t>(b) * hs + (oy >> p.pre_shift)) * ws + (ox >> p.pre_shift)) * p.pre_ld + n_g;
    if (full && (p.pre_ld & 3) == 0) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        float4 t = __ldg(reinterpret_cast<const float4*>(r + j));
        v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (n_g + j < p.N) v[j] += __ldg(r + j);
    }
  }

  if (p.act == GLSDET_ACT_YOLOX_BOX) {
    // models/core/utils_bbox.py:270-305: sigmoid on obj, (xy + grid) * stride, exp(wh) * stride, normalise
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int n = n_g + j;
      if (n == 0) v[j] = ((v[j] + static_cast<float>(ox)) * p.dec_stride) / p.dec_in_w;
      else if (n == 1) v[j] = ((v[j] + static_cast<float>(oy)) * p.dec_stride) / p.dec_in_h;
      else if (n == 2) v[j] = (expf(v[j]) * p.dec_stride) / p.dec_in_w;
      else if (n == 3) v[j] = (expf(v[j]) * p.dec_stride) / p.dec_in_h;
      else v[j] = 1.0f / (1.0f + expf(-v[j]));
    }
  } else if (p.act != GLSDET_ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = apply_act(v[j], p.act);
  }

  if (p.post_res != nullptr) {
    const int hs = p.Ho >> p.post_shift, ws = p.Wo >> p.post_shift;
    const __nv_bfloat16* r = p.post_res +
        ((static_cast<int64_t>(b) * hs + (oy >> p.post_shift)) * ws + (ox >> p.post_shift)) * p.post_ld + n_g;
    if (full && (p.post_ld & 7) == 0) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint4 t = __ldg(reinterpret_cast<const uint4*>(r) + h);
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float lo, hi;
          unpack_16x2(w[q], p.post_f16, lo, hi);
          v[h * 8 + q * 2] += lo;
          v[h * 8 + q * 2 + 1] += hi;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (n_g + j < p.N) v[j] += from_16(reinterpret_cast<const uint16_t*>(r)[j], p.post_f16);
    }
  }

  if (p.out_mode == GLSDET_OUT_NHWC_BF16) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<int64_t>(b) * p.out_bs +
                       (static_cast<int64_t>(oy) * p.Wo + ox) * p.out_ld + p.out_coff + n_g;
    if (full && ((p.out_ld | p.out_coff) & 7) == 0) {
      uint4 a, c;
      pack16_row(v, p.out_f16, a, c);
      reinterpret_cast<uint4*>(o)[0] = a;
      reinterpret_cast<uint4*>(o)[1] = c;
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (n_g + j < p.N) reinterpret_cast<uint16_t*>(o)[j] = to_16(v[j], p.out_f16);
    }
  } else if (p.out_mode == GLSDET_OUT_NHWC_F32) {
    float* o = reinterpret_cast<float*>(p.out) + static_cast<int64_t>(b) * p.out_bs +
               (static_cast<int64_t>(oy) * p.Wo + ox) * p.out_ld + p.out_coff + n_g;
    if (full && ((p.out_ld | p.out_coff) & 3) == 0 && (p.out_bs & 3) == 0) {
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (n_g + j < p.N) o[j] = v[j];
    }
  } else {  // NCHW fp32: consecutive lanes are consecutive x -> coalesced per channel
    const int64_t plane = static_cast<int64_t>(p.Ho) * p.Wo;
    float* o = reinterpret_cast<float*>(p.out) + static_cast<int64_t>(b) * p.out_bs +
               (static_cast<int64_t>(p.out_coff) + n_g) * plane + static_cast<int64_t>(oy) * p.Wo + ox;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (n_g + j < p.N) o[j * plane] = v[j];
  }
}

// Specialised epilogues (chosen per op at create time; everything warp-uniform is hoisted out of the element
// loops).  bf16 NHWC output of 16 channels = two 16-byte stores.
// Residual operands staged in shared memory by TMA (128-byte rows, 128-byte swizzle): this pixel's rows and the position
// of the 16-channel chunk inside the 64-channel tile.
struct ResRows {
  const uint8_t* post;    // 64 x 16-bit channels
  const uint8_t* pre0;    // fp32 channels 0..31
  const uint8_t* pre1;    // fp32 channels 32..63
  int sw_post, sw_pre;    // row & 7 of the post / pre rows (swizzle phase)
  int cc;                 // chunk of 16 channels, 0..3
};
__device__ __forceinline__ float4 res_pre4(const ResRows& rr, int q) {      // floats 4q .. 4q+3 of the chunk
  const uint8_t* row = (rr.cc & 2) ? rr.pre1 : rr.pre0;
  return *reinterpret_cast<const float4*>(row + (((((rr.cc & 1) << 2) | q) ^ rr.sw_pre) << 4));
}
__device__ __forceinline__ uint4 res_post8(const ResRows& rr, int h) {      // 16-bit values 8h .. 8h+7 of the chunk
  return *reinterpret_cast<const uint4*>(rr.post + ((((rr.cc << 1) | h) ^ rr.sw_post) << 4));
}

template <bool PRE, bool POST>
__device__ __forceinline__ void epi16_bf16(const uint32_t (&raw)[16], const float* s_bias, int act,
                                           const float* pre, const __nv_bfloat16* post, uint4* o0, uint4* o1,
                                           int dt, const ResRows* rs = nullptr) {   // dt: bit 0 = fp16 output, bit 1 = fp16 post residual
  float v[16];
  if (act == GLSDET_ACT_SILU) {
    // s_bias holds bias / 2 (ConvKParams::hbias): h = x / 2 in one FMA, silu(x) = h + h * tanh(h)
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 bv = *reinterpret_cast<const float4*>(s_bias + j);
      v[j] = fmaf(__uint_as_float(raw[j]), 0.5f, bv.x);
      v[j + 1] = fmaf(__uint_as_float(raw[j + 1]), 0.5f, bv.y);
      v[j + 2] = fmaf(__uint_as_float(raw[j + 2]), 0.5f, bv.z);
      v[j + 3] = fmaf(__uint_as_float(raw[j + 3]), 0.5f, bv.w);
    }
    if (PRE) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 t = rs ? res_pre4(*rs, j >> 2) : __ldg(reinterpret_cast<const float4*>(pre + j));
        v[j] = fmaf(t.x, 0.5f, v[j]); v[j + 1] = fmaf(t.y, 0.5f, v[j + 1]);
        v[j + 2] = fmaf(t.z, 0.5f, v[j + 2]); v[j + 3] = fmaf(t.w, 0.5f, v[j + 3]);
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaf(v[j], tanh_approx(v[j]), v[j]);
  } else {
#pragma unroll
  for (int j = 0; j < 16; j += 4) {
    const float4 bv = *reinterpret_cast<const float4*>(s_bias + j);
    v[j] = __uint_as_float(raw[j]) + bv.x;
    v[j + 1] = __uint_as_float(raw[j + 1]) + bv.y;
    v[j + 2] = __uint_as_float(raw[j + 2]) + bv.z;
    v[j + 3] = __uint_as_float(raw[j + 3]) + bv.w;
  }
  if (PRE) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 t = rs ? res_pre4(*rs, j >> 2) : __ldg(reinterpret_cast<const float4*>(pre + j));
      v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
    }
  }
  }
  if (act == GLSDET_ACT_SILU) {
  } else if (act == kActSiluExact) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = silu_newton(v[j]);
  } else if (act == GLSDET_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.0f);
  }
  if (POST) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint4 t = rs ? res_post8(*rs, h) : __ldg(reinterpret_cast<const uint4*>(post) + h);
      const uint32_t w[4] = {t.x, t.y, t.z, t.w};
      if (dt & 2) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float lo, hi;
          unpack_f16x2(w[q], lo, hi);
          v[h * 8 + q * 2] += lo;
          v[h * 8 + q * 2 + 1] += hi;
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          v[h * 8 + q * 2] += __uint_as_float(w[q] << 16);
          v[h * 8 + q * 2 + 1] += __uint_as_float(w[q] & 0xFFFF0000u);
        }
      }
    }
  }
  uint4 a, c;
  pack16_row(v, dt & 1, a, c);
  *o0 = a;
  *o1 = c;
}
template <bool PRE, bool POST>
__device__ __forceinline__ void epi16_bf16(const uint32_t (&raw)[16], const float* s_bias, int act,
                                           const float* pre, const __nv_bfloat16* post, __nv_bfloat16* o, int dt) {
  epi16_bf16<PRE, POST>(raw, s_bias, act, pre, post, reinterpret_cast<uint4*>(o), reinterpret_cast<uint4*>(o) + 1, dt);
}

// bias only, fp32 NHWC, 16 channels = four 16-byte stores (low-resolution partial sums)
__device__ __forceinline__ void epi16_f32_plain(const uint32_t (&raw)[16], const float* s_bias, float* o) {
#pragma unroll
  for (int j = 0; j < 16; j += 4) {
    const float4 bv = *reinterpret_cast<const float4*>(s_bias + j);
    *reinterpret_cast<float4*>(o + j) = make_float4(__uint_as_float(raw[j]) + bv.x, __uint_as_float(raw[j + 1]) + bv.y,
                                                    __uint_as_float(raw[j + 2]) + bv.z, __uint_as_float(raw[j + 3]) + bv.w);
  }
}

// Fused prediction conv: acc[0..4*NV) += act(v[j]) * Wp[k = k0 + j][0..4*NV) for the 16 channels of one chunk.
// s_pw is [N][16] (prediction channel fastest), so the float4 weight loads are warp-uniform broadcasts.
template <int NV>
__device__ __forceinline__ void pred_accumulate16(const uint32_t (&raw)[16], const float* s_bias, bool silu,
                                                  const float* s_pw, float (&acc)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float v = __uint_as_float(raw[j]) + s_bias[j];
    v = silu ? silu_fast(v) : fmaxf(v, 0.0f);
    const float4* w = reinterpret_cast<const float4*>(s_pw + j * 16);
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      const float4 ww = w[q];
      acc[4 * q] = fmaf(v, ww.x, acc[4 * q]);
      acc[4 * q + 1] = fmaf(v, ww.y, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(v, ww.z, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(v, ww.w, acc[4 * q + 3]);
    }
  }
}

// Output of the fused prediction conv for one pixel: raw logits (NCHW planes or NHWC rows), sigmoid rows, or the
// YOLOX box decode (utils_bbox.py:270-305).
__device__ __forceinline__ void store_pred(const ConvKParams& p, const float (&y)[16], int b, int oy, int ox) {
  const bool planes = (p.out_mode == GLSDET_OUT_NCHW_F32);
  // planes: lanes are consecutive x -> one coalesced run per channel; rows: [.., 5+nc] per pixel
  float* o = planes ? reinterpret_cast<float*>(p.out) + static_cast<int64_t>(b) * p.out_bs +
                          static_cast<int64_t>(p.out_coff) * p.out_plane + static_cast<int64_t>(oy) * p.Wo + ox
                    : reinterpret_cast<float*>(p.out) + static_cast<int64_t>(b) * p.out_bs +
                          (static_cast<int64_t>(oy) * p.Wo + ox) * p.out_ld + p.out_coff;
  const int64_t st = planes ? p.out_plane : 1;
  if (p.pred_act == GLSDET_ACT_YOLOX_BOX) {
    o[0] = ((y[0] + static_cast<float>(ox)) * p.dec_stride) / p.dec_in_w;
    o[st] = ((y[1] + static_cast<float>(oy)) * p.dec_stride) / p.dec_in_h;
    o[2 * st] = (expf(y[2]) * p.dec_stride) / p.dec_in_w;
    o[3 * st] = (expf(y[3]) * p.dec_stride) / p.dec_in_h;
    o[4 * st] = 1.0f / (1.0f + expf(-y[4]));
  } else if (p.pred_act == GLSDET_ACT_MMDET_BOX) {
    // mmdet yolox_head.py:298-301: xy = pred * stride + prior (prior = cell index * stride), wh = exp(pred) * stride;
    // explicit roundings: no FMA contraction, like the PyTorch reference
    o[0] = __fadd_rn(__fmul_rn(y[0], p.dec_stride), static_cast<float>(ox) * p.dec_stride);
    o[st] = __fadd_rn(__fmul_rn(y[1], p.dec_stride), static_cast<float>(oy) * p.dec_stride);
    o[2 * st] = __fmul_rn(expf(y[2]), p.dec_stride);
    o[3 * st] = __fmul_rn(expf(y[3]), p.dec_stride);
    o[4 * st] = 1.0f / (1.0f + expf(-y[4]));
  } else if (p.pred_act == GLSDET_ACT_SIGMOID) {
    // all 16 columns unconditionally (independent chains overlap; a per-column branch on pred_n serialised ten
    // exp + divide latency chains per pixel), then predicated stores
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = 1.0f / (1.0f + expf(-y[j]));
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < p.pred_n) o[j * st] = v[j];
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < p.pred_n) o[j * st] = y[j];
  }
}

// Walks the accumulator columns of one tile: two tcgen05.ld in flight per wait, then f(raw, chunk) on each.
template <typename F>
__device__ __forceinline__ void epi_walk(uint32_t taddr, int c_begin, int c_end, F&& f) {
  int c = c_begin;
  for (; c + 2 <= c_end; c += 2) {
    uint32_t v0[16], v1[16];
    tmem_ld16(taddr + static_cast<uint32_t>(c * 16), v0);
    tmem_ld16(taddr + static_cast<uint32_t>(c * 16 + 16), v1);
    tmem_ld_wait();
    f(v0, c);
    f(v1, c + 1);
  }
  if (c < c_end) {
    uint32_t v0[16];
    tmem_ld16(taddr + static_cast<uint32_t>(c * 16), v0);
    tmem_ld_wait();
    f(v0, c);
  }
}

// TMA-store epilogue of one 128-pixel tile (bf16 NHWC output).  Direct per-thread stores cost one L1 wavefront per
// 16 bytes (every lane owns a different pixel row) - 16 * N cycles per tile, more than the MMAs of a K < 512 conv.
// Here each thread writes its row into a 128-pixel x 64-channel staging tile (128-byte swizzle, conflict-free), and
// one thread per warp group hands the tile to the TMA unit, which also clips ragged tiles at the tensor bounds.
// Staging is double-buffered per group: the issuer waits until its earlier stores have finished reading shared
// memory BEFORE the group barrier, so after barrier i every thread may overwrite the buffer of store i-1.
struct TsCtx {
  uint8_t* gbuf;      // this group's two staging tiles
  int nthr, bar_id;   // group size / named barrier
  uint32_t bufmask;   // 1: two staging tiles (double-buffered); 0: one (the group waits for its previous store first)
  bool issuer;
  bool wide;          // block_n >= 128: each column half owns whole 64-channel tiles; else both halves share one
  int parts;          // column parts = epilogue warps / 4 (2, or 4 in the 16-warp variant: block_n 64 or 128 only)
  // res_tma: this sub-group's residual slot, its arrival barrier and the phase of the next arrival
  const uint8_t* res;
  uint64_t* res_bar;
  uint32_t res_count;  // residual tiles consumed by this sub-group (barrier phase = count & 1)
  int res_row;        // this pixel's row in the post / pre residual tiles
  int res_row_pre;
};
template <bool PRE, bool POST, int RS>   // RS: residual operands staged by TMA - 0 never, 1 decided at run time, 2 always
__device__ __forceinline__ void epi_tile_ts(const ConvKParams& p, TsCtx& g, uint32_t taddr, int r, int half,
                                            int k_tiles, bool valid, int act, const float* sb, const float* pre,
                                            const __nv_bfloat16* post, const TileCoord& t, int y_tile,
                                            uint32_t& sbuf) {
  const int c2 = p.patch ? (t.b & 1) : 0;        // patch views: image b' = (b*2 + py)*2 + px
  const int c4 = p.patch ? (t.b >> 1) : t.b;
  int kc_begin = g.wide ? (half ? (k_tiles + 1) >> 1 : 0) : 0;
  int kc_end = g.wide ? (half ? k_tiles : (k_tiles + 1) >> 1) : 1;
  if (g.parts == 4) {   // one 64-channel tile per group (block_n 64: tile 0 for everybody; 128: tile = half / 2)
    kc_begin = g.wide ? (half >> 1) : 0;
    kc_end = kc_begin + 1;
  }
  for (int kc = kc_begin; kc < kc_end; ++kc, ++sbuf) {
    uint8_t* buf = g.gbuf + (sbuf & g.bufmask) * kStageTileBytes;
    uint8_t* rowp = buf + (r >> 3) * 1024 + (r & 7) * kRowBytes;
    if (g.bufmask == 0u) {   // single staging tile: the previous store must have finished reading it
      if (g.issuer) tma_store_wait_read();
      named_bar_sync(g.bar_id, g.nthr);
    }
    int cb = g.wide ? kc * 4 : half * 2;
    int ce = g.wide ? kc * 4 + 4 : half * 2 + 2;
    const int dt = p.out_f16 | (p.post_f16 << 1);
    if (g.parts == 4) {   // 32 columns per warp (block_n 128) or 16 (block_n 64)
      cb = g.wide ? kc * 4 + (half & 1) * 2 : half;
      ce = g.wide ? cb + 2 : half + 1;
    }
    ResRows rs;
    const bool use_rs = (RS == 2) ? (PRE || POST) : (RS == 1 && (PRE || POST) && g.res_bar != nullptr);
    if (use_rs) {   // the residual tiles of this (work item, 64-channel tile) were requested before the accumulator wait
      mbar_wait(g.res_bar, g.res_count & 1u);
      ++g.res_count;
      rs.post = g.res + g.res_row * kRowBytes;
      rs.pre0 = g.res + p.res_post_bytes + g.res_row_pre * kRowBytes;
      rs.pre1 = rs.pre0 + p.res_pre_half_bytes;
      rs.sw_post = g.res_row & 7;
      rs.sw_pre = g.res_row_pre & 7;
    }
    if (!(p.dbg & 2))
    epi_walk(taddr, cb, ce, [&](const uint32_t (&raw)[16], int c) {
      if (valid) {
        const int cc = (c & 3) * 2;   // 16-byte chunk of the 128-byte staging row
        if (use_rs) {
          ResRows q = rs;
          q.cc = c & 3;
          epi16_bf16<PRE, POST>(raw, sb + c * 16, act, nullptr, nullptr,
                                reinterpret_cast<uint4*>(rowp + ((cc ^ (r & 7)) << 4)),
                                reinterpret_cast<uint4*>(rowp + (((cc + 1) ^ (r & 7)) << 4)), dt, &q);
        } else {
        epi16_bf16<PRE, POST>(raw, sb + c * 16, act, PRE ? pre + c * 16 : nullptr, POST ? post + c * 16 : nullptr,
                              reinterpret_cast<uint4*>(rowp + ((cc ^ (r & 7)) << 4)),
                              reinterpret_cast<uint4*>(rowp + (((cc + 1) ^ (r & 7)) << 4)), dt);
        }
      }
    });
    fence_proxy_async_smem();
    if (g.issuer) tma_store_wait_read();
    named_bar_sync(g.bar_id, g.nthr);
    if (g.issuer && !(p.dbg & 1)) {
      tma_store_5d(&p.tmO, buf, t.n0 + kc * kChunkK, t.x0, c2, y_tile, c4);
      tma_store_commit();
    }
  }
}

// EW = epilogue warps (8, or 16 for the TMA-store class: its epilogue is a chain of short dependent instructions per
// tile, so four warps per scheduler instead of two hide twice the latency; the register budget of that class allows it)
template <bool k2, int EC, int EW = kEpiWarps>
__global__ void __launch_bounds__((4 + EW) * 32, 1) conv_gemm_kernel(const __grid_constant__ ConvKParams p) {
  constexpr int kThreads = (4 + EW) * 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  // 2-CTA mode: every CTA stages half of the B rows, and one stage always carries all nsub taps (single ring)
  const int b_tap_bytes = (k2 ? (p.block_n >> 1) : p.block_n) * kRowBytes;
  const int b_bytes = (k2 ? p.nsub : p.bgroup) * b_tap_bytes;
  const int b_region = p.bres ? p.b_chunks * b_tap_bytes : p.sb * b_bytes;
  const bool pred_mma = (p.epi == EPI_TOWER_PRED_MMA);
  const int k_tiles = p.block_n >> 6;   // 64-channel operand tiles of the activated tile (prediction MMA)
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + p.sa * p.a_bytes;
  uint8_t* smem_stage = smem_b + b_region;                                          // [k_tiles][128 rows][128 B]
  const int ts_groups = (p.block_n >= 128 || (EW == 16 && p.egrp)) ? 2 : 1;   // warp groups with their own staging tiles
  uint8_t* smem_res = smem_stage + (p.ts ? ts_groups * 2 * kStageTileBytes : 0);   // res_tma: 4 residual slots
  uint8_t* smem_pw = pred_mma ? smem_stage + k_tiles * kStageTileBytes                 // [k_tiles][16 rows][128 B]
                              : smem_res + (p.res_tma ? 4 * p.res_slot_bytes : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_pw + (pred_mma ? k_tiles * kPredTileBytes : 0));
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + kMaxStages;
  uint64_t* b_full = bars + 2 * kMaxStages;
  uint64_t* b_empty = bars + 3 * kMaxStages;
  uint64_t* tfull_bar = bars + 4 * kMaxStages;
  uint64_t* tempty_bar = bars + 4 * kMaxStages + 4;   // tfull / tempty: up to 4 accumulator slots each
  uint64_t* bres_full = bars + 4 * kMaxStages + 8;
  uint64_t* pred_bar = bars + 4 * kMaxStages + 9;
  uint64_t* pstage_bar = bars + 4 * kMaxStages + 11;   // 2-CTA: both CTAs have staged their activated tile (leader's copy)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * kMaxStages + 10);
  uint64_t* res_bar = bars + 4 * kMaxStages + 12;      // res_tma: residual tiles of a staging sub-group have landed (4 x 2)
  float* s_bias = reinterpret_cast<float*>(bars + 4 * kMaxStages + 20);  // [n_blocks * block_n], zero padded (x 0.5: hbias)
  float* s_pw = s_bias + p.n_blocks * p.block_n;                         // FMA prediction path: weights [N][16]
  float* s_red = s_pw + p.block_n * 16;                                  // [2][128][16] partial sums of the upper column half

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) trace_stamp(p, 0);
  const uint32_t cta_rank = k2 ? cluster_ctarank() : 0u;
  const int work0 = k2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int work_stride = k2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int total_work = k2 ? p.total_pairs : p.total_tiles;
  auto decode = [&](int work) { return k2 ? decode_pair(p, work, static_cast<int>(cta_rank)) : decode_tile(p, work); };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    if (p.chunks1 > 0 || p.stride == 2) tma_prefetch_desc(&p.tmA[1]);
    tma_prefetch_desc(&p.tmB);
    if (p.bgroup == 3) { tma_prefetch_desc(&p.tmB3[0]); tma_prefetch_desc(&p.tmB3[1]); }
    if (p.ts) tma_prefetch_desc(&p.tmO);
    if (p.res_tma && p.res_post_bytes) tma_prefetch_desc(&p.tmRpost);
    if (p.res_tma && p.res_pre_half_bytes) tma_prefetch_desc(&p.tmRpre);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.sa; ++s) {
      mbar_init(&a_full[s], k2 ? 2 : 1);   // 2-CTA: both producers arrive on the leader's barrier
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < p.sb; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&tfull_bar[s], 1);
      const int ew_arr = (EW == 16 && p.egrp) ? EW / 2 : EW;   // warps that drain one accumulator slot
      mbar_init(&tempty_bar[s], k2 ? 2 * ew_arr : ew_arr);  // 2-CTA: the peer's epilogue warps arrive too
    }
    mbar_init(bres_full, 1);
    for (int i = 0; i < 8; ++i) mbar_init(&res_bar[i], 1);
    mbar_init(pred_bar, 1);
    mbar_init(pstage_bar, 2);
    fence_mbar_init();
  }
  if (warp == 2) {
    if (k2) { tmem2_alloc(tmem_slot, static_cast<uint32_t>(p.tmem_cols)); tmem2_relinquish(); }
    else { tmem_alloc(tmem_slot, static_cast<uint32_t>(p.tmem_cols)); tmem_relinquish(); }
  }
  // Programmatic dependent launch: everything above overlaps the tail of the previous kernel of the stream; global
  // memory (inputs, and outputs a predecessor may still read) is only touched after this point.
  if (p.pdl) {
    griddep_wait();
    griddep_launch_dependents();
  }
  if (threadIdx.x == 0) trace_stamp(p, 1);
  {
    // hbias: the tanh-form SiLU epilogues work on h = x / 2 = fma(acc, 0.5, bias / 2) - one FMA-pipe op less per element
    const float bscale = p.hbias ? 0.5f : 1.0f;
    for (int i = threadIdx.x; i < p.n_blocks * p.block_n; i += kThreads)
      s_bias[i] = (p.bias != nullptr && i < p.N) ? bscale * __ldg(p.bias + i) : 0.0f;
  }
  if (pred_mma) {
    // prediction weights fp32 [pred_n][N] -> bf16 B operand (16 rows, K-major, 128-byte swizzle) per 64-channel tile
    // (2-CTA: the M256 x N16 prediction MMA takes 8 of the 16 rows from each CTA, at the same shared-memory offset)
    const int prow = k2 ? 8 : 16, prow0 = k2 ? static_cast<int>(cta_rank) * 8 : 0;
    for (int i = threadIdx.x; i < prow * p.block_n; i += kThreads) {
      const int n = i / p.block_n, k = i - n * p.block_n;
      const int ng = prow0 + n;
      const float v = (ng < p.pred_n && k < p.N) ? __ldg(p.pred_w + static_cast<int64_t>(ng) * p.N + k) : 0.0f;
      const int kc = k >> 6, kk = k & 63;
      const int off = kc * kPredTileBytes + (n >> 3) * 1024 + (n & 7) * kRowBytes + ((((kk >> 3) ^ (n & 7))) << 4) + (kk & 7) * 2;
      *reinterpret_cast<uint16_t*>(smem_pw + off) = to_16(v, p.a_f16);
    }
    fence_proxy_async_smem();
  }
  if (p.epi == EPI_TOWER_PRED) {
    for (int i = threadIdx.x; i < p.block_n * 16; i += kThreads) {
      const int kk = i >> 4, j = i & 15;
      s_pw[i] = (kk < p.N && j < p.pred_n) ? __ldg(p.pred_w + static_cast<int64_t>(j) * p.N + kk) : 0.0f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (k2) cluster_sync_all();   // the peer's barriers must be initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) trace_stamp(p, 2);

  if (k2 && warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer, 2-CTA mode (single ring)
    int s = 0;
    uint32_t ph = 0;
    const uint32_t pair_bytes = 2u * static_cast<uint32_t>(p.a_bytes + b_bytes);
    const int n_half = p.block_n >> 1;
    for (int work = work0; work < total_work; work += work_stride) {
      const TileCoord t = decode(work);
      const int nrow0 = t.n0 + static_cast<int>(cta_rank) * n_half;
      for (int src = 0; src < 2; ++src) {
        const int nch = src ? p.chunks1 : p.chunks0;
        if (nch == 0) continue;
        const int kbase = src ? p.taps * p.chunks0 : 0;
        const int outer = (p.nsub == 3) ? 3 : p.taps;     // kx for 3x3 with vertical reuse, else the tap itself
        for (int o = 0; o < outer; ++o) {
          for (int ch = 0; ch < nch; ++ch) {
            mbar_wait(&a_empty[s], ph ^ 1u);
            if (cta_rank == 0) mbar_arrive_expect_tx(&a_full[s], pair_bytes);
            else mbar_arrive_cluster_relaxed(&a_full[s], 0);   // counts this producer; the data is tracked by complete_tx
            uint8_t* bdst = smem_b + s * b_bytes;
            if (p.nsub == 3) {
              tma2_load_5d(smem_a + s * p.a_bytes, &p.tmA[src], &a_full[s], ch * kChunkK, t.x0 + o - 1, 0, t.y0 - 1, t.b);
              if (p.bgroup == 3) {
                tma2_load_3d(bdst, &p.tmB3[src], &a_full[s], (kbase + o * nch + ch) * kChunkK, nrow0, 0);
              } else {
                for (int ky = 0; ky < 3; ++ky)
                  tma2_load_2d(bdst + ky * b_tap_bytes, &p.tmB, &a_full[s], (kbase + (ky * 3 + o) * nch + ch) * kChunkK, nrow0);
              }
            } else {
              const int dy = (p.taps == 9) ? o / 3 - 1 : 0;
              const int dx = (p.taps == 9) ? o % 3 - 1 : 0;
              tma2_load_5d(smem_a + s * p.a_bytes, &p.tmA[src], &a_full[s], ch * kChunkK, t.x0 + dx, 0, t.y0 + dy, t.b);
              tma2_load_2d(bdst, &p.tmB, &a_full[s], (kbase + o * nch + ch) * kChunkK, nrow0);
            }
            if (++s == p.sa) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (k2 && warp == 1) {
    // ------------------------------------------------------------ MMA issuer, 2-CTA mode (leader only)
    // The whole warp runs the loop (uniform control flow keeps descriptors in uniform registers); one elected lane
    // issues.  A divergent single-lane loop costs a register->uniform-register move per operand of every MMA.
    if (cta_rank == 0) {
      const uint32_t idesc = umma_idesc_16(2 * kBlockM, static_cast<uint32_t>(p.block_n), p.a_f16 != 0);
      const int a_sub_bytes = kRowBytes << p.tile_w_log2;
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int work = work0; work < total_work; work += work_stride, ++it) {
        const int as = it % p.nacc;
        const uint32_t aph = static_cast<uint32_t>(it / p.nacc) & 1u;
        mbar_wait(&tempty_bar[as], aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * p.block_n);
        uint32_t acc = 0u;
        for (int a = 0; a < p.a_steps; ++a) {
          mbar_wait(&a_full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + s * p.a_bytes);
          const uint32_t b_addr = smem_u32(smem_b + s * b_bytes);
          if (elect_one()) {
            for (int sub = 0; sub < p.nsub; ++sub) {
              const uint64_t da = umma_desc_k_sw128(a_addr + static_cast<uint32_t>(sub * a_sub_bytes));
              const uint64_t db = umma_desc_k_sw128(b_addr + static_cast<uint32_t>(sub * b_tap_bytes));
#pragma unroll
              for (int kk = 0; kk < kChunkK / 16; ++kk) {
                umma2_bf16(d_tmem, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2), idesc,
                           (acc | static_cast<uint32_t>(sub | kk)) ? 1u : 0u);
              }
            }
            umma2_commit_both(&a_empty[s]);   // frees this stage in both CTAs
            if (a == p.a_steps - 1) umma2_commit_both(&tfull_bar[as]);  // both epilogues may read their half
          }
          __syncwarp();
          acc = 1u;
          if (++s == p.sa) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (!k2 && warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer
    int sa = 0, sb = 0;
    uint32_t pha = 0, phb = 0;
    auto load_a = [&](const CUtensorMap* tm, int c0, int c1, int c2, int c3, int c4) {
      mbar_wait(&a_empty[sa], pha ^ 1u);
      mbar_arrive_expect_tx(&a_full[sa], static_cast<uint32_t>(p.a_bytes));
      tma_load_5d(smem_a + sa * p.a_bytes, tm, &a_full[sa], c0, c1, c2, c3, c4);
      if (p.trace != nullptr && p.trace[3] == 0ull) trace_stamp(p, 3);
      if (++sa == p.sa) { sa = 0; pha ^= 1u; }
    };
    int wb = 0;   // image index of the weight matrix (per-image weights)
    auto load_b = [&](int kchunk, int n0) {
      if (p.bres) return;
      mbar_wait(&b_empty[sb], phb ^ 1u);
      mbar_arrive_expect_tx(&b_full[sb], static_cast<uint32_t>(b_bytes));
      if (p.w_batched) tma_load_3d(smem_b + sb * b_bytes, &p.tmB, &b_full[sb], kchunk * kChunkK, n0, wb);
      else tma_load_2d(smem_b + sb * b_bytes, &p.tmB, &b_full[sb], kchunk * kChunkK, n0);
      if (++sb == p.sb) { sb = 0; phb ^= 1u; }
    };
    auto load_b3 = [&](int src, int kchunk, int n0) {
      mbar_wait(&b_empty[sb], phb ^ 1u);
      mbar_arrive_expect_tx(&b_full[sb], static_cast<uint32_t>(b_bytes));
      tma_load_3d(smem_b + sb * b_bytes, &p.tmB3[src], &b_full[sb], kchunk * kChunkK, n0, 0);
      if (++sb == p.sb) { sb = 0; phb ^= 1u; }
    };
    if (p.bres && work0 < total_work) {
      // resident weights (one N block): every K chunk once, at its K position
      mbar_arrive_expect_tx(bres_full, static_cast<uint32_t>(p.b_chunks * b_tap_bytes));
      for (int c = 0; c < p.b_chunks; ++c)
        tma_load_2d(smem_b + c * b_tap_bytes, &p.tmB, bres_full, c * kChunkK, 0);
    }
    for (int tile = work0; tile < total_work; tile += work_stride) {
      const TileCoord t = decode_tile(p, tile);
      wb = t.b;
      const int ac2 = p.patch ? (t.b & 1) : 0;                              // patch views: b' = (b*2 + py)*2 + px
      const int ab = p.a_shared ? (p.a_div ? t.b / p.a_div : t.b % p.a_shared) : (p.patch ? (t.b >> 1) : t.b);
      if (p.stride == 1) {
        for (int src = 0; src < 2; ++src) {
          const int nch = src ? p.chunks1 : p.chunks0;
          if (nch == 0) continue;
          const int kbase = src ? p.taps * p.chunks0 : 0;  // weight K order: (source, tap, chunk)
          const int pad = p.ksz >> 1;
          const int padx = p.kw >> 1;
          if (p.nsub > 1) {   // one A box of (rows + ksz - 1) rows per (kx, chunk) serves all ky taps
            for (int kx = 0; kx < p.kw; ++kx) {
              for (int ch = 0; ch < nch; ++ch) {
                load_a(&p.tmA[src], ch * kChunkK, t.x0 + kx - padx, 0, t.y0 - pad, ab);
                if (p.bres) continue;
                if (p.bgroup == 3) load_b3(src, kbase + kx * nch + ch, t.n0);         // ky = 0,1,2 in one box
                else for (int ky = 0; ky < p.ksz; ++ky) load_b(kbase + (ky * p.kw + kx) * nch + ch, t.n0);
              }
            }
          } else {
            for (int tap = 0; tap < p.taps; ++tap) {
              const int dy = tap / p.kw - pad;
              const int dx = tap % p.kw - padx;
              for (int ch = 0; ch < nch; ++ch) {
                load_a(&p.tmA[src], ch * kChunkK, t.x0 + dx, ac2, t.y0 + dy, ab);
                load_b(kbase + tap * nch + ch, t.n0);
              }
            }
          }
        }
      } else if (p.kw == 2) {
        // stride 2 over pixel pairs (two input pixels = one 2C-channel pixel): input row 2*oy + ky - 1 -> (half row,
        // parity in coordinate 2); columns are dense with taps dx = -1 (the pair left of the output) and 0
        for (int ky = 0; ky < 3; ++ky) {
          const int py = (ky == 1) ? 0 : 1;
          const int yh = t.y0 + (ky == 0 ? -1 : 0);
          for (int kx = 0; kx < 2; ++kx) {
            for (int ch = 0; ch < p.chunks0; ++ch) {
              load_a(&p.tmA[0], ch * kChunkK, t.x0 + kx - 1, py, yh, ab);
              load_b((ky * 2 + kx) * p.chunks0 + ch, t.n0);
            }
          }
        }
      } else {
        // stride 2: input row 2*oy + ky - 1 -> (half row, parity); tmA[0] holds even columns, tmA[1] odd ones
        for (int tap = 0; tap < 9; ++tap) {
          const int ky = tap / 3, kx = tap % 3;
          const int map = (kx == 1) ? 0 : 1;
          const int xh = t.x0 + (kx == 0 ? -1 : 0);
          const int py = (ky == 1) ? 0 : 1;
          const int yh = t.y0 + (ky == 0 ? -1 : 0);
          for (int ch = 0; ch < p.chunks0; ++ch) {
            load_a(&p.tmA[map], ch * kChunkK, xh, py, yh, ab);
            load_b(tap * p.chunks0 + ch, t.n0);
          }
        }
      }
    }
  } else if (!k2 && warp == 1) {
    // ------------------------------------------------------------ MMA issuer (whole warp, one elected lane issues)
    const uint32_t idesc = umma_idesc_16(kBlockM, static_cast<uint32_t>(p.block_n), p.a_f16 != 0);
    const int a_sub_bytes = kRowBytes << p.tile_w_log2;  // tile_w rows
    const int a_m_bytes = p.tile_h * a_sub_bytes;         // distance between the M tiles of a work item
    int sa = 0, sb = 0;
    uint32_t pha = 0, phb = 0;
    int it = 0;
    if (p.bres && work0 < total_work) mbar_wait(bres_full, 0);
    for (int tile = work0; tile < total_work; tile += work_stride, ++it) {
      const int as = it % p.nacc;
      const uint32_t aph = static_cast<uint32_t>(it / p.nacc) & 1u;
      mbar_wait(&tempty_bar[as], aph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * p.mt * p.block_n);
      uint32_t acc = 0u;
      int src = 0, o = 0, ch = 0;   // position of this A step in the producer's (source, kx | tap, chunk) order
      for (int a = 0; a < p.a_steps; ++a) {
        mbar_wait(&a_full[sa], pha);
        if (p.trace != nullptr && lane == 0 && p.trace[4] == 0ull) trace_stamp(p, 4);
        const uint32_t a_addr = smem_u32(smem_a + sa * p.a_bytes);
        const int nch = src ? p.chunks1 : p.chunks0;
        for (int sub = 0; sub < p.nsub; ++sub) {
          const int bsub = (p.bgroup == 3) ? sub : 0;
          uint32_t b_addr;
          bool last_of_b = false;
          if (p.bres) {
            const int kchunk = (p.nsub > 1) ? (src ? p.taps * p.chunks0 : 0) + (sub * p.kw + o) * nch + ch : a;
            b_addr = smem_u32(smem_b + kchunk * b_tap_bytes);
          } else {
            if (bsub == 0) mbar_wait(&b_full[sb], phb);
            b_addr = smem_u32(smem_b + sb * b_bytes + bsub * b_tap_bytes);
            last_of_b = (p.bgroup == 1 || sub == p.nsub - 1);
          }
          tc_fence_after();
          if (elect_one()) {
            // ky tap = sub: the 128 rows of the MMA start sub * tile_w rows into the (mt * tile_h + 2)-row A stage
            const uint64_t db = umma_desc_k_sw128(b_addr);
            for (int m = 0; m < p.mt; ++m) {
              const uint64_t da = umma_desc_k_sw128(a_addr + static_cast<uint32_t>(m * a_m_bytes + sub * a_sub_bytes));
              const uint32_t d_m = d_tmem + static_cast<uint32_t>(m * p.block_n);
#pragma unroll
              for (int kk = 0; kk < kChunkK / 16; ++kk) {
                // +32 bytes per K=16 step inside the 128-byte swizzle row (start-address field is in 16-byte units)
                umma_bf16(d_m, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2), idesc,
                          (acc | static_cast<uint32_t>(kk)) ? 1u : 0u);
              }
            }
            if (last_of_b) umma_commit(&b_empty[sb]);
            if (sub == p.nsub - 1) {
              umma_commit(&a_empty[sa]);
              if (a == p.a_steps - 1) umma_commit(&tfull_bar[as]);
            }
          }
          __syncwarp();
          acc = 1u;
          if (last_of_b) {
            if (++sb == p.sb) { sb = 0; phb ^= 1u; }
          }
        }
        if (++sa == p.sa) { sa = 0; pha ^= 1u; }
        if (++ch == nch) {
          ch = 0;
          if (++o == p.kw) { o = 0; ++src; }
        }
      }
      if (lane == 0) trace_stamp(p, 5);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int e = warp - 4;
    const int q = e & 3;      // TMEM lane quarter this warp may read (== warp % 4)
    const int half = e >> 2;  // which part of the accumulator columns (2 parts with 8 warps, 4 with 16)
    const int r = q * 32 + lane;
    const int py = r >> p.tile_w_log2;
    const int px = r & ((1 << p.tile_w_log2) - 1);
    const int n_chunks = p.block_n >> 4;
    // column chunks of this warp: two parts with 8 epilogue warps, four with 16
    const int c_begin = (EW == 16) ? (n_chunks * half) / 4 : (half ? ((n_chunks + 1) >> 1) : 0);
    const int c_end = (EW == 16) ? (n_chunks * (half + 1)) / 4 : (half ? n_chunks : ((n_chunks + 1) >> 1));
    const bool silu = (p.act == GLSDET_ACT_SILU);
    const int mt = k2 ? 1 : p.mt;
    TsCtx tsg;
    tsg.wide = p.block_n >= 128;
    if (EW == 16) {
      // four column parts: block_n = 64 -> all 16 warps share one 64-channel tile; block_n = 128 -> parts (0,1) own
      // tile 0, parts (2,3) tile 1
      const int grp = tsg.wide ? (half >> 1) : 0;
      tsg.gbuf = smem_stage + grp * 2 * kStageTileBytes;
      tsg.nthr = tsg.wide ? 256 : 512;
      tsg.bar_id = 6 + grp;
      tsg.issuer = (lane == 0) && (q == 0) && (tsg.wide ? ((half & 1) == 0) : (half == 0));
    } else {
      tsg.gbuf = smem_stage + (tsg.wide ? half : 0) * 2 * kStageTileBytes;
      tsg.nthr = tsg.wide ? 128 : 256;
      tsg.bar_id = 6 + (tsg.wide ? half : 0);
      tsg.issuer = (lane == 0) && (tsg.wide ? (q == 0) : (e == 0));
    }
    tsg.parts = EW / 4;
    tsg.bufmask = 1u;
    const bool egrp = (EW == 16) && p.egrp;
    const int my_slot = half >> 1;   // egrp: the accumulator slot (= parity of the work item) this warp group drains
    if (egrp && p.block_n >= 128) {
      // group of 8 warps = the 8-warp wide layout: the four warps with (half & 1) == s own 64-channel tile s of the
      // item, with ONE staging tile each (4 sub-groups x 16 KB; the other group's work hides the store wait)
      const int sub = half & 1;
      tsg.wide = true;
      tsg.gbuf = smem_stage + (my_slot * 2 + sub) * kStageTileBytes;
      tsg.nthr = 128;
      tsg.bar_id = 6 + my_slot * 2 + sub;
      tsg.issuer = (lane == 0) && (q == 0);
      tsg.parts = 2;
      tsg.bufmask = 0u;
    } else if (egrp) {   // group of 8 warps = the 8-warp, one-64-channel-tile layout: column halves by (half & 1)
      tsg.wide = false;
      tsg.gbuf = smem_stage + my_slot * 2 * kStageTileBytes;
      tsg.nthr = 256;
      tsg.bar_id = 6 + my_slot;
      tsg.issuer = (lane == 0) && (q == 0) && ((half & 1) == 0);
      tsg.parts = 2;
    }
    const int half_ts = egrp ? (half & 1) : half;
    tsg.res = nullptr; tsg.res_bar = nullptr; tsg.res_count = 0u; tsg.res_row = 0; tsg.res_row_pre = 0;
    const int res_kc = (egrp && p.block_n >= 128) ? (half & 1) : 0;   // the 64-channel tile this sub-group stores
    if (p.res_tma && egrp) {
      const int sub_idx = my_slot * 2 + res_kc;
      tsg.res = smem_res + sub_idx * p.res_slot_bytes;
      tsg.res_bar = &res_bar[sub_idx];
      tsg.res_row = ((py >> p.post_shift) << (p.tile_w_log2 - p.post_shift)) + (px >> p.post_shift);
      tsg.res_row_pre = ((py >> p.pre_shift) << (p.tile_w_log2 - p.pre_shift)) + (px >> p.pre_shift);
    }
    uint32_t sbuf = 0;     // staging tiles written by this warp group (TMA-store epilogue)
    int it = 0;
    uint32_t tcount = 0;   // tiles processed by this CTA (prediction-MMA barrier phase, FMA scratch slot)
    for (int tile = work0; tile < total_work; tile += work_stride, ++it) {
      if (egrp && (it & 1) != my_slot) continue;
      const TileCoord t = decode(tile);
      const int as = it % p.nacc;
      const uint32_t aph = static_cast<uint32_t>(it / p.nacc) & 1u;
      if constexpr (EC == EC_ALL || EC == EC_BF16_TSR) {
        // Residual operands of the TMA-store epilogue.  Per-thread global loads of this pixel's row cost one L1 tag
        // lookup per 16 bytes (every lane another pixel, the same problem the TMA store solves on the way out): 1x1
        // 128 -> 128 + residual at 256^2 ran 152 us against 100 us without one.  Instead the TMA unit drops the residual
        // tile of this (work item, 64-channel tile) into shared memory, requested here - before the accumulator wait -
        // and read row-wise (swizzled, conflict-free) in epi_tile_ts.
        if (p.res_tma && tsg.issuer && tsg.res_bar != nullptr) {
          // (two slots per sub-group with the request one work item further ahead measured slower: 145 -> 160 us on the
          // 256^2 layer - the extra tile decode and bookkeeping cost the 96-register epilogue more than the latency hid)
          mbar_arrive_expect_tx(tsg.res_bar, static_cast<uint32_t>(p.res_post_bytes + 2 * p.res_pre_half_bytes));
          uint8_t* dst = const_cast<uint8_t*>(tsg.res);
          const int ch0 = t.n0 + res_kc * kChunkK;
          if (p.res_post_bytes)
            tma_load_5d(dst, &p.tmRpost, tsg.res_bar, ch0, t.x0 >> p.post_shift, 0, t.y0 >> p.post_shift, t.b);
          if (p.res_pre_half_bytes) {   // fp32 viewed as 2-byte elements: 32 floats = one 128-byte row
            tma_load_5d(dst + p.res_post_bytes, &p.tmRpre, tsg.res_bar, 2 * ch0, t.x0 >> p.pre_shift, 0, t.y0 >> p.pre_shift, t.b);
            tma_load_5d(dst + p.res_post_bytes + p.res_pre_half_bytes, &p.tmRpre, tsg.res_bar, 2 * ch0 + kChunkK,
                        t.x0 >> p.pre_shift, 0, t.y0 >> p.pre_shift, t.b);
          }
        }
      }
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      if (p.trace != nullptr && warp == 4 && lane == 0 && p.trace[6] == 0ull) trace_stamp(p, 6);
      for (int m = 0; m < mt; ++m, ++tcount) {
      const uint32_t tcol = static_cast<uint32_t>((as * mt + m) * p.block_n);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + tcol;
      const int oy = t.y0 + m * p.tile_h + py, ox = t.x0 + px;
      const bool valid = (oy < p.Ho) && (ox < p.Wo);
      const float* sb = s_bias + t.n0;
      // every lane of the warp must execute the TMEM loads; only valid pixels / channels below N are stored
      switch (p.epi) {
        case EPI_BF16:
        case EPI_BF16_PRE:
        case EPI_BF16_POST:
        case EPI_BF16_PREPOST: { if constexpr (EC == EC_ALL || EC == EC_BF16 || EC == EC_BF16_TS || EC == EC_BF16_TSR) {
          const bool has_pre = (p.epi == EPI_BF16_PRE || p.epi == EPI_BF16_PREPOST);
          const bool has_post = (p.epi == EPI_BF16_POST || p.epi == EPI_BF16_PREPOST);
          const float* pre = nullptr;
          const __nv_bfloat16* post = nullptr;
          if (has_pre) {   // a shift beyond the image size selects one row per image (per-image bias)
            const int hs = max(p.Ho >> p.pre_shift, 1), ws = max(p.Wo >> p.pre_shift, 1);
            pre = p.pre_res + ((static_cast<int64_t>(t.b) * hs + (oy >> p.pre_shift)) * ws + (ox >> p.pre_shift)) * p.pre_ld + t.n0;
          }
          if (has_post) {
            if (p.patch) {   // residual read from the un-split tensor [B, 2 Ho, 2 Wo, ld]
              const int pb = t.b >> 2, ppy = (t.b >> 1) & 1, ppx = t.b & 1;
              post = p.post_res + ((static_cast<int64_t>(pb) * 2 * p.Ho + ppy * p.Ho + oy) * (2 * p.Wo) + ppx * p.Wo + ox) * p.post_ld + t.n0;
            } else {
              const int hs = p.Ho >> p.post_shift, ws = p.Wo >> p.post_shift;
              post = p.post_res + ((static_cast<int64_t>(t.b) * hs + (oy >> p.post_shift)) * ws + (ox >> p.post_shift)) * p.post_ld + t.n0;
            }
          }
          if constexpr (EC == EC_ALL || EC == EC_BF16_TS || EC == EC_BF16_TSR) if (p.ts) {
            constexpr int RS = (EC == EC_BF16_TSR) ? 2 : (EC == EC_ALL) ? 1 : 0;
            const int y_tile = t.y0 + m * p.tile_h;
            if (has_pre && has_post) epi_tile_ts<true, true, RS>(p, tsg, taddr, r, half_ts, k_tiles, valid, p.act_epi, sb, pre, post, t, y_tile, sbuf);
            else if (has_pre) epi_tile_ts<true, false, RS>(p, tsg, taddr, r, half_ts, k_tiles, valid, p.act_epi, sb, pre, post, t, y_tile, sbuf);
            else if (has_post) epi_tile_ts<false, true, RS>(p, tsg, taddr, r, half_ts, k_tiles, valid, p.act_epi, sb, pre, post, t, y_tile, sbuf);
            else if constexpr (EC != EC_BF16_TSR) epi_tile_ts<false, false, 0>(p, tsg, taddr, r, half_ts, k_tiles, valid, p.act_epi, sb, pre, post, t, y_tile, sbuf);
            break;
          }
          if constexpr (EC == EC_ALL || EC == EC_BF16) {
          const int dtf = p.out_f16 | (p.post_f16 << 1);
          __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<int64_t>(t.b) * p.out_bs +
                                (static_cast<int64_t>(oy) * p.Wo + ox) * p.out_ld + p.out_coff + t.n0;
          if (has_pre && has_post) {
            epi_walk(taddr, c_begin, c_end, [&](const uint32_t (&raw)[16], int c) {
              if (valid && t.n0 + c * 16 < p.N) epi16_bf16<true, true>(raw, sb + c * 16, p.act_epi, pre + c * 16, post + c * 16, orow + c * 16, dtf);
            });
          } else if (has_pre) {
            epi_walk(taddr, c_begin, c_end, [&](const uint32_t (&raw)[16], int c) {
              if (valid && t.n0 + c * 16 < p.N) epi16_bf16<true, false>(raw, sb + c * 16, p.act_epi, pre + c * 16, nullptr, orow + c * 16, dtf);
            });
          } else if (has_post) {
            epi_walk(taddr, c_begin, c_end, [&](const uint32_t (&raw)[16], int c) {
              if (valid && t.n0 + c * 16 < p.N) epi16_bf16<false, true>(raw, sb + c * 16, p.act_epi, nullptr, post + c * 16, orow + c * 16, dtf);
            });
          } else {
            epi_walk(taddr, c_begin, c_end, [&](const uint32_t (&raw)[16], int c) {
              if (valid && t.n0 + c * 16 < p.N) epi16_bf16<false, false>(raw, sb + c * 16, p.act_epi, nullptr, nullptr, orow + c * 16, dtf);
            });
          }
          }  // direct stores
          }
          break;
        }
        case EPI_F32_PLAIN: { if constexpr (EC == EC_ALL || EC == EC_F32) {
          float* orow = reinterpret_cast<float*>(p.out) + static_cast<int64_t>(t.b) * p.out_bs +
                        (static_cast<int64_t>(oy) * p.Wo + ox) * p.out_ld + p.out_coff + t.n0;
          epi_walk(taddr, c_begin, c_end, [&](const uint32_t (&raw)[16], int c) {
            if (valid && t.n0 + c * 16 < p.N) epi16_f32_plain(raw, sb + c * 16, orow + c * 16);
          });
          }
          break;
        }
        case EPI_NCHW_RAW: { if constexpr (EC == EC_ALL || EC == EC_SMALL) {  // N <= 16: raw logits, reference layout; lanes are consecutive x -> coalesced per channel
          const int64_t plane = static_cast<int64_t>(p.Ho) * p.Wo;
          float* o = reinterpret_cast<float*>(p.out) + static_cast<int64_t>(t.b) * p.out_bs +
                     static_cast<int64_t>(p.out_coff) * plane + static_cast<int64_t>(oy) * p.Wo + ox;
          epi_walk(taddr, c_begin, c_end, [&](const uint32_t (&raw)[16], int c) {
            if (valid && c == 0) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < p.N) o[j * plane] = __uint_as_float(raw[j]) + sb[j];
            }
          });
          }
          break;
        }
        case EPI_ROWS_BOX: { if constexpr (EC == EC_ALL || EC == EC_SMALL) {  // N == 5: utils_bbox.py:270-305 on (x, y, w, h, obj) -> decoded row prefix
          float* o = reinterpret_cast<float*>(p.out) + static_cast<int64_t>(t.b) * p.out_bs +
                     (static_cast<int64_t>(oy) * p.Wo + ox) * p.out_ld + p.out_coff;
          epi_walk(taddr, c_begin, c_end, [&](const uint32_t (&raw)[16], int c) {
            if (valid && c == 0) {
              o[0] = ((__uint_as_float(raw[0]) + sb[0] + static_cast<float>(ox)) * p.dec_stride) / p.dec_in_w;
              o[1] = ((__uint_as_float(raw[1]) + sb[1] + static_cast<float>(oy)) * p.dec_stride) / p.dec_in_h;
              o[2] = (expf(__uint_as_float(raw[2]) + sb[2]) * p.dec_stride) / p.dec_in_w;
              o[3] = (expf(__uint_as_float(raw[3]) + sb[3]) * p.dec_stride) / p.dec_in_h;
              o[4] = 1.0f / (1.0f + expf(-(__uint_as_float(raw[4]) + sb[4])));
            }
          });
          }
          break;
        }
        case EPI_ROWS_SIGMOID: { if constexpr (EC == EC_ALL || EC == EC_SMALL) {  // N <= 16 class probabilities of a decoded row
          float* o = reinterpret_cast<float*>(p.out) + static_cast<int64_t>(t.b) * p.out_bs +
                     (static_cast<int64_t>(oy) * p.Wo + ox) * p.out_ld + p.out_coff;
          epi_walk(taddr, c_begin, c_end, [&](const uint32_t (&raw)[16], int c) {
            if (valid && c == 0) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < p.N) o[j] = 1.0f / (1.0f + expf(-(__uint_as_float(raw[j]) + sb[j])));
            }
          });
          }
          break;
        }
        case EPI_TOWER_PRED: { if constexpr (EC == EC_ALL || EC == EC_PRED_FMA) {
          // second tower conv + prediction conv: the activated tile never leaves the SM.  Each thread reduces its
          // row over its half of the channels; the two warps of a lane quarter meet at a 64-thread named barrier.
          float acc[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
          if (p.pred_n <= 8) {
            epi_walk(taddr, c_begin, c_end, [&](const uint32_t (&raw)[16], int c) {
              pred_accumulate16<2>(raw, sb + c * 16, silu, s_pw + c * 256, acc);
            });
          } else if (p.pred_n <= 12) {
            epi_walk(taddr, c_begin, c_end, [&](const uint32_t (&raw)[16], int c) {
              pred_accumulate16<3>(raw, sb + c * 16, silu, s_pw + c * 256, acc);
            });
          } else {
            epi_walk(taddr, c_begin, c_end, [&](const uint32_t (&raw)[16], int c) {
              pred_accumulate16<4>(raw, sb + c * 16, silu, s_pw + c * 256, acc);
            });
          }
          float* red = s_red + ((tcount & 1u) * kBlockM + r) * 16;
          if (half == 1) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(red + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
          }
          asm volatile("bar.sync %0, 64;" ::"r"(q + 1) : "memory");
          if (half == 0 && valid) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 o4 = *reinterpret_cast<const float4*>(red + j);
              acc[j] += o4.x; acc[j + 1] += o4.y; acc[j + 2] += o4.z; acc[j + 3] += o4.w;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < p.pred_n) acc[j] += __ldg(p.pred_b + j);
            store_pred(p, acc, t.b, oy, ox);
          }
          }
          break;
        }
        case EPI_TOWER_PRED_MMA: { if constexpr (EC == EC_ALL || EC == EC_PRED_MMA) {
          // second tower conv + prediction conv on the tensor core: the activated tile goes to shared memory as a
          // bf16 K-major operand (row = pixel, 128-byte swizzle), one thread issues M128 x N16 x K(block_n) MMAs into
          // the first 16 columns of this (already drained) accumulator, and the lower-half warps decode + store.
          epi_walk(taddr, c_begin, c_end, [&](const uint32_t (&raw)[16], int c) {
            float v[16];
            const float ascale = p.hbias ? 0.5f : 1.0f;   // hbias: sb holds bias / 2, v = x / 2
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 bv = *reinterpret_cast<const float4*>(sb + c * 16 + j);
              v[j] = fmaf(__uint_as_float(raw[j]), ascale, bv.x);
              v[j + 1] = fmaf(__uint_as_float(raw[j + 1]), ascale, bv.y);
              v[j + 2] = fmaf(__uint_as_float(raw[j + 2]), ascale, bv.z);
              v[j + 3] = fmaf(__uint_as_float(raw[j + 3]), ascale, bv.w);
            }
            if (silu && p.act_epi == kActSiluExact) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = silu_newton(v[j]);
            } else if (silu) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = fmaf(v[j], tanh_approx(v[j]), v[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.0f);
            }
            uint4 lo, hi;
            pack16_row(v, p.a_f16, lo, hi);
            uint8_t* rowp = smem_stage + (c >> 2) * kStageTileBytes + (r >> 3) * 1024 + (r & 7) * kRowBytes;
            const int cc = (c & 3) * 2;   // 16-byte chunk of the 128-byte row
            *reinterpret_cast<uint4*>(rowp + ((cc ^ (r & 7)) << 4)) = lo;
            *reinterpret_cast<uint4*>(rowp + (((cc + 1) ^ (r & 7)) << 4)) = hi;
          });
          fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core (async proxy)
          tc_fence_before();
          named_bar_sync(5, EW * 32);
          if (k2) {
            // both CTAs of the pair stage their tile; the leader issues ONE M256 x N16 MMA chain over both halves and
            // its commit arrives on pred_bar in both CTAs.  The staged data is consumed by each SM's own tensor core
            // after the leader has seen both arrives, so the arrive itself needs no release fence.
            if (e == 0) {
              if (elect_one()) mbar_arrive_cluster_relaxed(pstage_bar, 0);
              __syncwarp();
              if (cta_rank == 0) {
                mbar_wait(pstage_bar, tcount & 1u);
                tc_fence_after();
                if (elect_one()) {
                  const uint32_t idesc_p = umma_idesc_16(2 * kBlockM, 16, p.a_f16 != 0);
                  const uint32_t d_pred = tmem_base + tcol;
                  for (int kc = 0; kc < k_tiles; ++kc) {
                    const uint64_t da = umma_desc_k_sw128(smem_u32(smem_stage + kc * kStageTileBytes));
                    const uint64_t db = umma_desc_k_sw128(smem_u32(smem_pw + kc * kPredTileBytes));
#pragma unroll
                    for (int kk = 0; kk < kChunkK / 16; ++kk)
                      umma2_bf16(d_pred, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2), idesc_p,
                                 (kc | kk) ? 1u : 0u);
                  }
                  umma2_commit_both(pred_bar);
                }
                __syncwarp();
              }
            }
          } else if (e == 0) {
            tc_fence_after();
            if (elect_one()) {
              const uint32_t idesc_p = umma_idesc_16(kBlockM, 16, p.a_f16 != 0);
              const uint32_t d_pred = tmem_base + tcol;
              for (int kc = 0; kc < k_tiles; ++kc) {
                const uint64_t da = umma_desc_k_sw128(smem_u32(smem_stage + kc * kStageTileBytes));
                const uint64_t db = umma_desc_k_sw128(smem_u32(smem_pw + kc * kPredTileBytes));
#pragma unroll
                for (int kk = 0; kk < kChunkK / 16; ++kk)
                  umma_bf16(d_pred, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2), idesc_p,
                            (kc | kk) ? 1u : 0u);
              }
              umma_commit(pred_bar);
            }
            __syncwarp();
          }
          mbar_wait(pred_bar, tcount & 1u);
          tc_fence_after();
          if (half == 0) {
            uint32_t yr[16];
            tmem_ld16(taddr, yr);
            tmem_ld_wait();
            if (valid) {
              float y[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) y[j] = __uint_as_float(yr[j]) + ((j < p.pred_n) ? __ldg(p.pred_b + j) : 0.0f);
              store_pred(p, y, t.b, oy, ox);
            }
          }
          }
          break;
        }
        default: { if constexpr (EC == EC_ALL) {
          for (int c = c_begin; c < c_end; ++c) {
            uint32_t v[16];
            tmem_ld16(taddr + static_cast<uint32_t>(c * 16), v);
            tmem_ld_wait();
            if (valid && t.n0 + c * 16 < p.N)
              epilogue_store16(p, v, s_bias + t.n0 + c * 16, t.b, oy, ox, t.n0 + c * 16);
          }
        } }
      }
      }  // M tiles of the work item
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (k2) mbar_arrive_cluster_relaxed(&tempty_bar[as], 0);   // the leader's MMA thread waits for both epilogues
        else mbar_arrive(&tempty_bar[as]);
      }
    }
    if (warp == 4 && lane == 0) trace_stamp(p, 7);
    if (p.ts && tsg.issuer) tma_store_wait_all();   // shared memory must outlive the bulk stores
    if (warp == 4 && lane == 0) trace_stamp(p, 8);
  }

  tc_fence_before();
  __syncthreads();
  if (k2) cluster_sync_all();   // the peer may still be reading operands / signalling barriers of this CTA
  tc_fence_after();
  if (threadIdx.x == 0) trace_stamp(p, 9);
  if (warp == 2) {
    if (k2) tmem2_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
    else tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
    if (lane == 0) trace_stamp(p, 10);
  }
}

}  // namespace glsdet

// ===================================================================================== host side
using namespace glsdet;

struct glsdet_conv {
  ConvKParams kp;
  int grid;
  int smem_bytes;
  int two_cta;   // launched as clusters of 2 CTAs issuing tcgen05.mma.cta_group::2
  int ec;        // epilogue class = kernel instantiation (EC_*)
  int threads;   // CTA size of that instantiation
};

namespace {
typedef void (*ConvKernelFn)(const ConvKParams);
constexpr int EC_BF16_TS16 = EC_COUNT;   // TMA-store class with 16 epilogue warps (640 threads)
constexpr int EC_PRED_MMA16 = EC_COUNT + 1;   // tensor-core prediction class with 16 epilogue warps
constexpr int EC2_ALL = EC_COUNT + 2;         // 2-CTA kernel, every epilogue
constexpr int EC2_PRED_MMA16 = EC_COUNT + 3;  // 2-CTA kernel, tensor-core prediction class, 16 epilogue warps
constexpr int EC2_BF16_TS16 = EC_COUNT + 4;   // 2-CTA kernel, TMA-store class, 16 epilogue warps
constexpr int EC_BF16_TS16R = EC_COUNT + 5;   // TMA-store class with TMA-staged residual operands, 16 epilogue warps
constexpr int EC2_BF16_TS16R = EC_COUNT + 6;  // the same on the 2-CTA kernel
constexpr int EC_KERNELS = EC_COUNT + 7;
ConvKernelFn conv_kernel_for(int ec) {
  switch (ec) {
    case EC_BF16_TS16R: return conv_gemm_kernel<false, EC_BF16_TSR, 16>;
    case EC2_BF16_TS16R: return conv_gemm_kernel<true, EC_BF16_TSR, 16>;
    case EC_BF16_TSR: return conv_gemm_kernel<false, EC_BF16_TS>;   // (8-warp residual class: not used)
    case EC_BF16_TS16: return conv_gemm_kernel<false, EC_BF16_TS, 16>;
    case EC_PRED_MMA16: return conv_gemm_kernel<false, EC_PRED_MMA, 16>;
    case EC2_ALL: return conv_gemm_kernel<true, EC_ALL>;
    case EC2_PRED_MMA16: return conv_gemm_kernel<true, EC_PRED_MMA, 16>;
    case EC2_BF16_TS16: return conv_gemm_kernel<true, EC_BF16_TS, 16>;
    case EC_BF16: return conv_gemm_kernel<false, EC_BF16>;
    case EC_BF16_TS: return conv_gemm_kernel<false, EC_BF16_TS>;
    case EC_F32: return conv_gemm_kernel<false, EC_F32>;
    case EC_SMALL: return conv_gemm_kernel<false, EC_SMALL>;
    case EC_PRED_FMA: return conv_gemm_kernel<false, EC_PRED_FMA>;
    case EC_PRED_MMA: return conv_gemm_kernel<false, EC_PRED_MMA>;
    default: return conv_gemm_kernel<false, EC_ALL>;
  }
}
}  // namespace

namespace {

struct ConvGeom {
  int taps, kw, chunks0, chunks1, n_blocks, block_n, n_pad, k_pad, Ho, Wo;
};

int conv_geometry(const glsdet_conv_desc* d, ConvGeom* g) {
  GLSDET_REQUIRE(d != nullptr, "conv: null descriptor");
  GLSDET_REQUIRE(d->ksize == 1 || d->ksize == 3 || d->ksize == 5 || d->ksize == 7,
                 "conv: ksize must be 1, 3, 5 or 7 (got %d)", d->ksize);
  GLSDET_REQUIRE(d->ksize <= 3 || d->stride == 1, "conv: 5x5 / 7x7 convs are stride 1");
  GLSDET_REQUIRE(d->stride == 1 || d->stride == 2, "conv: stride must be 1 or 2 (got %d)", d->stride);
  GLSDET_REQUIRE(d->batch > 0 && d->height > 0 && d->width > 0, "conv: bad input size");
  GLSDET_REQUIRE(d->src0_c > 0 && (d->src0_ld >= d->src0_c || d->src0_row_pitch > 0), "conv: bad src0 channels/pitch");
  const bool row_strided = d->ksize_w == 2 && d->stride == 2 && d->ksize == 3 && d->src1 == nullptr && d->patch_mode == 0 &&
                           d->src_shared == 0;   // 3x3 stride-2 conv over PIXEL PAIRS: rows stride 2, columns taps {-1, 0}
  GLSDET_REQUIRE(d->ksize_w == 0 || d->ksize_w == d->ksize || row_strided ||
                     (d->ksize_w == 1 && d->stride == 1 && d->src1 == nullptr && d->patch_mode == 0 && d->src_shared == 0),
                 "conv: ksize_w must be 0, ksize, 1 (kx taps folded into the channel view: stride 1, single source) or 2 "
                 "(3x3 stride-2 conv over pixel pairs)");
  GLSDET_REQUIRE(d->src0_row_pitch >= 0 && d->src0_img_pitch >= 0 && (d->src0_row_pitch % 8) == 0 && (d->src0_img_pitch % 8) == 0 &&
                     (d->src0_row_pitch == 0 || (d->stride == 1 && d->patch_mode == 0)),
                 "conv: explicit src0 row / image pitches must be multiples of 8 elements (stride-1 convs only)");
  GLSDET_REQUIRE((d->src0_ld % 8) == 0, "conv: src0 pitch must be a multiple of 8 elements (TMA 16-byte strides)");
  GLSDET_REQUIRE(d->out_channels > 0, "conv: out_channels must be positive");
  if (d->stride == 2) {
    GLSDET_REQUIRE(d->ksize == 3 && d->src1 == nullptr, "conv: stride 2 needs ksize 3 and a single source");
    GLSDET_REQUIRE((d->height % 2) == 0 && ((d->width % 2) == 0 || row_strided), "conv: stride 2 needs even height/width");
  }
  if (d->src1 != nullptr) {
    GLSDET_REQUIRE(d->src1_c > 0 && d->src1_ld >= d->src1_c && (d->src1_ld % 8) == 0, "conv: bad src1 channels/pitch");
  }
  if (d->patch_mode) {
    GLSDET_REQUIRE(d->ksize == 1 && d->src1 == nullptr && (d->batch % 4) == 0,
                   "conv: patch_mode needs a 1x1 conv, a single source and batch = 4 * images");
    GLSDET_REQUIRE(d->out_mode == GLSDET_OUT_NHWC_BF16 && d->pre_shift >= 0 && d->post_shift == 0,
                   "conv: patch_mode writes bf16 NHWC and reads its post residual at full resolution");
  }
  if (d->weight_batch_stride != 0) {
    GLSDET_REQUIRE(d->ksize == 1, "conv: per-image weights are supported for 1x1 convs");
    GLSDET_REQUIRE((d->weight_batch_stride % 8) == 0 && (d->weight_ld % 8) == 0,
                   "conv: per-image weight strides must be multiples of 8 elements");
  }
  g->kw = d->ksize_w > 0 ? d->ksize_w : d->ksize;
  g->taps = d->ksize * g->kw;
  g->chunks0 = (d->src0_c + kChunkK - 1) / kChunkK;
  g->chunks1 = d->src1 ? (d->src1_c + kChunkK - 1) / kChunkK : 0;
  g->n_blocks = (d->out_channels + 255) / 256;
  int per = (d->out_channels + g->n_blocks - 1) / g->n_blocks;
  g->block_n = (per + 15) / 16 * 16;
  {
    // 1x1 convs with 256+ output channels and a 16-bit output: N blocks of 128 instead of 256.  Their epilogue, not their
    // MMAs, paces them (one 128 x 256 tile: 2 us of MMAs at K = 512 against ~6 us in the 8-warp epilogue); 128-wide blocks
    // run in the 16-warp TMA-store class with two epilogue groups (~1.4 us per 128 x 128 tile).  The A tile is loaded once
    // per N block (consecutive work items share it through L2).  Opt-in (GLSDET_CONV_SPLIT_N=1), see below.
    const char* e = getenv("GLSDET_CONV_SPLIT_N");
    const bool on = (e && e[0] == '1');   // measured slower (2884 vs 2864 us per step): these layers are bound by operand streaming from L2, not by the epilogue
    const int kpad = d->ksize * (d->ksize_w > 0 ? d->ksize_w : d->ksize) * (g->chunks0 + g->chunks1) * kChunkK;
    if (on && d->ksize == 1 && d->stride == 1 && d->out_channels >= 256 && (d->out_channels % 128) == 0 &&
        d->out_mode == GLSDET_OUT_NHWC_BF16 && d->pred_weight == nullptr && d->weight_batch_stride == 0 &&
        d->patch_mode == 0 && kpad <= 640) {
      g->n_blocks = d->out_channels / 128;
      g->block_n = 128;
    }
  }
  {
    // Small 3x3 launches (MP-Det towers on the 25 x 42 .. 7 x 11 levels: 8 .. 96 tiles of K = 2304) leave most SMs idle while
    // each CTA issues 36 k-steps of 128 x 256 MMAs (~10 us at the full tensor rate, 24 us measured per launch).  Narrower N
    // blocks spread the same MMAs over more SMs.  GLSDET_CONV_SMALL_SPLIT=0 disables it.
    const char* e = getenv("GLSDET_CONV_SMALL_SPLIT");
    if (!(e && e[0] == '0') && d->ksize == 3 && d->stride == 1 && d->ksize_w == 0 && d->weight_batch_stride == 0 &&
        d->patch_mode == 0 && d->pred_weight == nullptr && d->src1 == nullptr && (d->out_channels % 64) == 0 &&
        g->n_blocks == 1 && g->block_n >= 128) {
      const int64_t tiles = static_cast<int64_t>(d->batch) * ((d->height + 7) / 8) * ((d->width + 15) / 16);
      const int sms = device_sm_count();
      int bn = g->block_n;
      while (bn > 64 && (bn % 2) == 0 && ((bn / 2) % 32) == 0 && tiles * (d->out_channels / (bn / 2)) <= sms) bn /= 2;
      if (bn != g->block_n && (d->out_channels % bn) == 0) {
        g->block_n = bn;
        g->n_blocks = d->out_channels / bn;
      }
    }
  }
  g->n_pad = g->n_blocks * g->block_n;
  g->k_pad = g->taps * (g->chunks0 + g->chunks1) * kChunkK;
  g->Ho = d->height / d->stride;
  g->Wo = row_strided ? d->width : d->width / d->stride;
  return 0;
}

constexpr int kPatchView = 3;   // encode_act_map `stride` value selecting the 2x2 patch view
constexpr int kRowStride2 = 4;  // rows split by parity (coordinate 2), columns dense: the pixel-pair form of a stride-2 conv

int encode_act_map(CUtensorMap* tm, const void* base, int c_view, int ld, int B, int H, int W, int stride,
                   int tile_w, int box_rows, int64_t row_pitch = 0, int64_t img_pitch = 0) {
  EncodeTiledFn enc = get_encode_tiled();
  GLSDET_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
  GLSDET_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "conv: source pointer must be 16-byte aligned");
  cuuint64_t dims[5];
  cuuint64_t strides[4];
  const cuuint64_t e = 2;  // bf16
  if (stride == 1) {
    // explicit pitches: rows / images of a padded buffer; with ld < c_view the pixels of the view overlap (the kx taps
    // of a few-channel conv folded into the channel dimension: "channel" k of pixel x = element k of the row starting at x)
    const cuuint64_t rp = row_pitch > 0 ? static_cast<cuuint64_t>(row_pitch) : static_cast<cuuint64_t>(W) * ld;
    const cuuint64_t ip = img_pitch > 0 ? static_cast<cuuint64_t>(img_pitch) : static_cast<cuuint64_t>(H) * rp;
    dims[0] = c_view; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = B;
    strides[0] = static_cast<cuuint64_t>(ld) * e;
    strides[1] = rp * e;
    strides[2] = rp * e;
    strides[3] = ip * e;
  } else if (stride == kRowStride2) {
    dims[0] = c_view; dims[1] = W; dims[2] = 2; dims[3] = H / 2; dims[4] = B;
    strides[0] = static_cast<cuuint64_t>(ld) * e;
    strides[1] = static_cast<cuuint64_t>(W) * ld * e;
    strides[2] = static_cast<cuuint64_t>(2) * W * ld * e;
    strides[3] = static_cast<cuuint64_t>(H) * W * ld * e;
  } else if (stride == kPatchView) {
    // 2x2 patch views of a [B/4, 2H, 2W, ld] tensor: coordinate 2 = px, coordinate 4 = image * 2 + py
    dims[0] = c_view; dims[1] = W; dims[2] = 2; dims[3] = H; dims[4] = B / 2;
    strides[0] = static_cast<cuuint64_t>(ld) * e;
    strides[1] = static_cast<cuuint64_t>(W) * ld * e;
    strides[2] = static_cast<cuuint64_t>(2) * W * ld * e;
    strides[3] = static_cast<cuuint64_t>(H) * 2 * W * ld * e;
  } else {
    dims[0] = c_view; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = B;
    strides[0] = static_cast<cuuint64_t>(2) * ld * e;
    strides[1] = static_cast<cuuint64_t>(W) * ld * e;
    strides[2] = static_cast<cuuint64_t>(2) * W * ld * e;
    strides[3] = static_cast<cuuint64_t>(H) * W * ld * e;
  }
  cuuint32_t box[5] = {static_cast<cuuint32_t>(kChunkK), static_cast<cuuint32_t>(tile_w), 1,
                       static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GLSDET_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation) failed with CUresult %d", (int)r);
  return 0;
}

}  // namespace

extern "C" int glsdet_conv_weight_shape(const glsdet_conv_desc* desc, int32_t* n_pad, int32_t* k_pad,
                                        int32_t* block_n) {
  ConvGeom g;
  if (int rc = conv_geometry(desc, &g)) return rc;
  if (n_pad) *n_pad = g.n_pad;
  if (k_pad) *k_pad = g.k_pad;
  if (block_n) *block_n = g.block_n;
  return 0;
}

extern "C" int glsdet_conv_create(const glsdet_conv_desc* d, glsdet_conv_t** out_op) {
  GLSDET_REQUIRE(out_op != nullptr, "conv_create: null output handle");
  *out_op = nullptr;
  ConvGeom g;
  if (int rc = conv_geometry(d, &g)) return rc;
  GLSDET_REQUIRE(d->src0 && d->weight && d->out, "conv_create: null src0/weight/out pointer");
  GLSDET_REQUIRE(d->out_mode >= 0 && d->out_mode <= 2, "conv_create: bad out_mode %d", d->out_mode);
  GLSDET_REQUIRE(d->act >= 0 && d->act <= GLSDET_ACT_YOLOX_BOX, "conv_create: bad act %d", d->act);  // MMDET_BOX: fused preds only
  GLSDET_REQUIRE((d->src_dtype | d->out_dtype | d->post_dtype) >= 0 && d->src_dtype <= GLSDET_DT_F16 &&
                     d->out_dtype <= GLSDET_DT_F16 && d->post_dtype <= GLSDET_DT_F16, "conv_create: bad storage dtype");
  if (d->pred_weight != nullptr) {
    GLSDET_REQUIRE(d->pred_bias != nullptr && d->pred_channels >= 1 && d->pred_channels <= 16,
                   "conv_create: fused prediction conv needs a bias and 1..16 channels");
    GLSDET_REQUIRE(d->out_channels <= 256 && (d->out_channels % 16) == 0,
                   "conv_create: fused prediction conv needs out_channels <= 256 and a multiple of 16");
    GLSDET_REQUIRE(d->act == GLSDET_ACT_SILU || d->act == GLSDET_ACT_RELU,
                   "conv_create: fused prediction conv supports SiLU/ReLU towers");
    GLSDET_REQUIRE(d->out_mode == GLSDET_OUT_NHWC_F32 || d->out_mode == GLSDET_OUT_NCHW_F32,
                   "conv_create: fused prediction conv writes fp32 rows or planes");
    GLSDET_REQUIRE(d->pred_act == GLSDET_ACT_NONE || d->pred_act == GLSDET_ACT_SIGMOID ||
                   ((d->pred_act == GLSDET_ACT_YOLOX_BOX || d->pred_act == GLSDET_ACT_MMDET_BOX) && d->pred_channels == 5),
                   "conv_create: bad pred_act");
    GLSDET_REQUIRE(d->pre_res == nullptr && d->post_res == nullptr, "conv_create: fused prediction conv takes no residual");
  }

  void* mem = nullptr;
  if (posix_memalign(&mem, 64, sizeof(glsdet_conv)) != 0 || mem == nullptr) {
    set_error("conv_create: out of host memory");
    return 3;
  }
  glsdet_conv* op = new (mem) glsdet_conv();
  ConvKParams& k = op->kp;
  k.B = d->batch; k.Ho = g.Ho; k.Wo = g.Wo;

  // 2-CTA mode (cta_group::2, M = 256 per CTA pair): each SM feeds the tensor core its own 128 rows of A and HALF of the
  // B rows, so the shared-memory operand traffic per MMA drops from A + B to A + B/2 (the 1-CTA kernel is bound by
  // it: 50 % tensor-pipe utilisation at N = 128, 67 % at N = 256) and every weight byte is fetched from L2 once per 256
  // pixels.  Measured (profiles/README.md): 3x3 N=256 at 256^2 491 -> 385 us (1.6 PF/s), N=128 299 -> 207 us.  Default
  // for 3x3 stride-1 convs; the small-K 1x1 convs stay on the 1-CTA kernel (resident weights, TMA-store epilogue).
  // GLSDET_CONV_2CTA=0 disables it, =1 also uses it for 1x1 convs.
  bool two_cta = d->stride == 1 && d->ksize == 3 && g.kw == 3 && (g.block_n % 32) == 0;
  // wide 1x1 convs (K >= 256, 256-wide N blocks) stream 384 KB of operands per tile from L2 and are bound by it: the pair
  // kernel halves the weight traffic (512 -> 512 at 64^2 52 -> 42 us, 256 -> 256 at 128^2 62 -> 54 us).  The small-K / narrow
  // 1x1 convs (resident weights, TMA-store epilogue) measured slower in pairs and stay on the 1-CTA kernel.
  if (d->stride == 1 && d->ksize == 1 && g.kw == 1 && g.block_n == 256 && (g.chunks0 + g.chunks1) >= 4 &&
      getenv("GLSDET_CONV_NO_2CTA_1X1") == nullptr)
    two_cta = true;
  if (const char* e = getenv("GLSDET_CONV_2CTA")) {
    if (e[0] == '0') two_cta = false;
    if (e[0] == '1') two_cta = d->stride == 1 && d->ksize <= 3 && g.kw == d->ksize && (g.block_n % 32) == 0;
  }
  if (d->weight_batch_stride != 0 || d->patch_mode != 0) two_cta = false;
  op->two_cta = two_cta ? 1 : 0;
  const int sms = device_sm_count();

  // Tile rectangle: 128 pixels (tile_h x tile_w), minimise padded area, prefer the squarest (best halo reuse for 3x3).
  // A work item is `mt` vertically adjacent tiles covered by ONE A box: every weight stage then feeds mt MMAs, which
  // halves the weight traffic L2 -> SM (the 1-tile kernel re-reads all weights per 128 pixels and saturates L2).
  auto pick_tile = [&](int mt, int* best_w_out) -> int64_t {
    int best_w = 16;
    int64_t best_cost = INT64_MAX;
    const int cand[5] = {16, 8, 32, 64, 128};
    for (int i = 0; i < 5; ++i) {
      const int tw = cand[i], th = (kBlockM / tw) * mt;
      const int64_t cost = static_cast<int64_t>((g.Wo + tw - 1) / tw) * ((g.Ho + th - 1) / th);
      if (cost < best_cost) { best_cost = cost; best_w = tw; }
    }
    *best_w_out = best_w;
    return best_cost;   // work items per image and N block
  };
  int mt = 1, best_w = 16;
  {
    int mt_mode = 0;   // 0 auto, 1 never, 2 whenever legal
    if (const char* e = getenv("GLSDET_CONV_MT")) mt_mode = (e[0] == '1') ? 1 : (e[0] == '2') ? 2 : 0;
    // block_n = 256 would need 2 x 2 x 256 TMEM columns for double-buffered accumulators; single-buffered (exposed
    // epilogue) it measured 15 % slower than one tile per item, so two-tile items stop at block_n = 128 by default
    int mt_max_n = 128;
    if (const char* e = getenv("GLSDET_CONV_MT_MAXN")) mt_max_n = atoi(e);
    int w2 = 16;
    const int64_t items2 = pick_tile(2, &w2) * d->batch * g.n_blocks;
    // resident-weight 1x1 convs gain nothing from sharing a weight stage between two tiles; with one tile per item the
    // 128-column accumulators fit four slots (two per epilogue group, see nacc): 128 -> 128 at 256^2 156 -> 146 us,
    // 256 -> 128 at 128^2 84 -> 80 us.  K = 64 (one A stage per tile) measured slower that way (87 -> 90 us) and keeps
    // two-tile items.  GLSDET_CONV_BRES_MT1=0 disables it.
    const bool bres_likely = g.n_blocks == 1 && d->weight_batch_stride == 0 &&
                             g.taps * (g.chunks0 + g.chunks1) * g.block_n * kRowBytes <= 96 * 1024;
    const char* mt1_env = getenv("GLSDET_CONV_BRES_MT1");
    const bool mt1_bres = !(mt1_env && mt1_env[0] == '0') && bres_likely && d->ksize == 1 && g.block_n == 128 &&
                          (g.chunks0 + g.chunks1) >= 2;
    const bool legal = !two_cta && mt_mode != 1 && !mt1_bres && g.block_n <= mt_max_n && 2 * g.block_n <= 512;
    if (legal && (mt_mode == 2 || items2 >= 2 * static_cast<int64_t>(sms))) { mt = 2; best_w = w2; }
    else pick_tile(1, &best_w);
  }
  k.mt = mt;
  k.nacc = (2 * mt * g.block_n <= 512) ? 2 : 1;
  int tw_log2 = 0;
  while ((1 << tw_log2) < best_w) ++tw_log2;
  k.tile_w_log2 = tw_log2;
  k.tile_h = kBlockM / best_w;
  k.tiles_x = (g.Wo + best_w - 1) / best_w;
  k.tiles_y = (g.Ho + k.tile_h * mt - 1) / (k.tile_h * mt);
  k.n_blocks = g.n_blocks;
  k.block_n = g.block_n;
  k.N = d->out_channels;
  const int64_t total = static_cast<int64_t>(k.tiles_x) * k.tiles_y * k.B * k.n_blocks;
  if (total > INT32_MAX) { free(mem); set_error("conv_create: too many tiles"); return 2; }
  k.total_tiles = static_cast<int32_t>(total);
  k.taps = g.taps; k.stride = d->stride;
  k.chunks0 = g.chunks0; k.chunks1 = g.chunks1;
  k.b_chunks = g.taps * (g.chunks0 + g.chunks1);

  // A/B rings.  3x3 stride-1: one (mt * tile_h + 2)-row A stage feeds three B sub-steps (ky taps).
  const int super_h = k.tile_h * mt;
  const bool vreuse = (d->ksize >= 3 && d->stride == 1 && (super_h + d->ksize - 1) <= 256 &&
                       (getenv("GLSDET_CONV_NO_VREUSE") == nullptr || d->ksize > 3));
  k.ksz = d->ksize;
  k.kw = g.kw;
  k.nsub = vreuse ? d->ksize : 1;
  const int box_rows = vreuse ? super_h + d->ksize - 1 : super_h;
  k.a_bytes = box_rows * best_w * kRowBytes;
  k.a_steps = (vreuse ? g.kw : g.taps) * (g.chunks0 + g.chunks1);
  k.m_tiles = k.tiles_x * k.tiles_y * k.B;
  k.total_pairs = ((k.m_tiles + 1) / 2) * k.n_blocks;

  k.epi = EPI_GENERIC;
  {
    const bool n16 = (d->out_channels % 16) == 0;
    const bool act_sr = (d->act == GLSDET_ACT_SILU || d->act == GLSDET_ACT_RELU || d->act == GLSDET_ACT_NONE);
    const bool bf16_vec = d->out_mode == GLSDET_OUT_NHWC_BF16 && n16 && ((d->out_ld | d->out_coff) % 8) == 0 &&
                          (d->out_batch_stride % 8) == 0;
    const bool small_n = d->out_channels <= 16;
    k.epi = EPI_GENERIC;
    if (bf16_vec && act_sr && !d->pre_res && !d->post_res) k.epi = EPI_BF16;
    else if (bf16_vec && act_sr && d->pre_res && !d->post_res && (d->pre_ld % 4) == 0 &&
             (reinterpret_cast<uintptr_t>(d->pre_res) & 15) == 0) k.epi = EPI_BF16_PRE;
    else if (bf16_vec && act_sr && !d->pre_res && d->post_res && (d->post_ld % 8) == 0 &&
             (reinterpret_cast<uintptr_t>(d->post_res) & 15) == 0) k.epi = EPI_BF16_POST;
    else if (bf16_vec && act_sr && d->pre_res && d->post_res && (d->pre_ld % 4) == 0 && (d->post_ld % 8) == 0 &&
             (reinterpret_cast<uintptr_t>(d->pre_res) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(d->post_res) & 15) == 0) k.epi = EPI_BF16_PREPOST;
    else if (d->out_mode == GLSDET_OUT_NHWC_F32 && d->act == GLSDET_ACT_NONE && n16 && !d->pre_res && !d->post_res &&
             ((d->out_ld | d->out_coff) % 4) == 0 && (d->out_batch_stride % 4) == 0 &&
             (reinterpret_cast<uintptr_t>(d->out) & 15) == 0) k.epi = EPI_F32_PLAIN;
    else if (d->out_mode == GLSDET_OUT_NCHW_F32 && d->act == GLSDET_ACT_NONE && small_n && !d->pre_res && !d->post_res)
      k.epi = EPI_NCHW_RAW;
    else if (d->out_mode == GLSDET_OUT_NHWC_F32 && d->act == GLSDET_ACT_YOLOX_BOX && d->out_channels == 5 &&
             !d->pre_res && !d->post_res) k.epi = EPI_ROWS_BOX;
    else if (d->out_mode == GLSDET_OUT_NHWC_F32 && d->act == GLSDET_ACT_SIGMOID && small_n && !d->pre_res &&
             !d->post_res) k.epi = EPI_ROWS_SIGMOID;
    if (const char* e = getenv("GLSDET_CONV_GENERIC_EPILOGUE")) {  // tests: force the generic epilogue
      if (e[0] == '1') k.epi = EPI_GENERIC;
    }
  }
  // TMA-store epilogue: small-K convs whose epilogue (not the MMAs) paces the kernel
  const bool fused_pred = d->pred_weight != nullptr;
  const bool patch = d->patch_mode != 0;
  const bool w_batched = d->weight_batch_stride != 0;
  bool ts = !fused_pred &&
            (k.epi == EPI_BF16 || k.epi == EPI_BF16_PRE || k.epi == EPI_BF16_POST || k.epi == EPI_BF16_PREPOST) &&
            (d->out_channels % 64) == 0 && (g.block_n % 64) == 0 && (g.k_pad <= 640 || patch) &&
            d->out_batch_stride == static_cast<int64_t>(g.Ho) * g.Wo * d->out_ld &&
            (getenv("GLSDET_CONV_NO_TMA_STORE") == nullptr || patch);
  if (patch && !ts) {
    free(mem);
    set_error("conv_create: patch_mode needs the TMA-store epilogue (bf16 output, channels a multiple of 64)");
    return 2;
  }

  // Fused prediction conv: on the tensor core when the tower width is a multiple of 64 (activated tile staged in
  // shared memory as a bf16 operand), else per-thread FMAs.
  bool pred_mma = fused_pred && g.n_blocks == 1 && (d->out_channels % 64) == 0 &&
                  getenv("GLSDET_CONV_PRED_FMA") == nullptr;
  // two epilogue groups draining alternate work items (16-warp TMA-store class, one 64-channel tile per M tile)
  const bool egrp_want = (g.block_n == 64 || (g.block_n == 128 && getenv("GLSDET_CONV_NO_EGRP128") == nullptr)) && (two_cta || k.nacc == 2) && getenv("GLSDET_CONV_EPI8") == nullptr &&
                         getenv("GLSDET_CONV_ONE_KERNEL") == nullptr && getenv("GLSDET_CONV_NO_EGRP") == nullptr;
  const int b_tap_bytes = g.block_n * kRowBytes;
  bool bres = false;
  k.bgroup = 1;
  // residual operands of the 16-warp TMA-store class through TMA + shared memory (see the epilogue): one work item = one
  // tile, residual resolution = the output's (post) or half of it (pre / post)
  bool res_tma = false;
  int res_post_bytes = 0, res_pre_half = 0;
  if ((d->pre_res || d->post_res) && egrp_want && ts && (g.block_n == 64 || g.block_n == 128) && mt == 1 && !patch &&
      (k.epi == EPI_BF16_PRE || k.epi == EPI_BF16_POST || k.epi == EPI_BF16_PREPOST) &&
      getenv("GLSDET_CONV_NO_RES_TMA") == nullptr) {
    bool ok = true;
    if (d->post_res) {
      const int sh = d->post_shift;
      ok = ok && (sh == 0 || sh == 1) && (best_w >> sh) >= 1 && (k.tile_h >> sh) >= 1 && (g.Ho % (1 << sh)) == 0 &&
           (g.Wo % (1 << sh)) == 0;
      res_post_bytes = kRowBytes * (best_w >> sh) * (k.tile_h >> sh);
      ok = ok && (res_post_bytes % 1024) == 0;
    }
    if (d->pre_res) {
      const int sh = d->pre_shift;
      ok = ok && sh == 1 && (best_w >> sh) >= 1 && (k.tile_h >> sh) >= 1 && (g.Ho % 2) == 0 && (g.Wo % 2) == 0;
      res_pre_half = kRowBytes * (best_w >> sh) * (k.tile_h >> sh);
      ok = ok && (res_pre_half % 1024) == 0;
    }
    res_tma = ok;
    if (!ok) res_post_bytes = res_pre_half = 0;
  }
  const int res_slot_bytes = res_post_bytes + 2 * res_pre_half;
  const int res_nbuf = 1;
  auto size_rings = [&](bool want_bgroup3) -> bool {
    const int pred_smem = !fused_pred ? 0
                          : pred_mma ? (g.block_n / 64) * (kStageTileBytes + kPredTileBytes)
                                     : (g.block_n * 16 + 2 * kBlockM * 16) * 4;
    const int ts_smem = ts ? ((g.block_n >= 128 || egrp_want) ? 4 : 2) * kStageTileBytes : 0;
    const int fixed = 1024 + 512 + g.n_pad * 4 + pred_smem + ts_smem + (ts && res_tma ? 4 * res_nbuf * res_slot_bytes : 0);
    const int budget = kSmemLimit - fixed;
    bres = false;
    k.bgroup = 1;
    if (two_cta) {  // single ring; a stage = A box + all nsub taps of this CTA's half of B
      k.bgroup = want_bgroup3 && vreuse && d->ksize == 3 ? 3 : 1;
      const int b_bytes2 = k.nsub * (g.block_n / 2) * kRowBytes;
      int stages = budget / (k.a_bytes + b_bytes2);
      if (stages > kMaxStages) stages = kMaxStages;
      if (stages < 2) return false;
      k.sa = k.sb = stages;
      op->smem_bytes = stages * (k.a_bytes + b_bytes2) + fixed;
      return true;
    }
    // resident weights: small K x N (one N block) - the weight traffic per tile disappears
    const int b_total = k.b_chunks * b_tap_bytes;
    if (g.n_blocks == 1 && !w_batched && b_total <= 96 * 1024 && getenv("GLSDET_CONV_NO_BRES") == nullptr) {
      int stages = (budget - b_total) / k.a_bytes;
      if (stages > kMaxStages) stages = kMaxStages;
      if (stages >= 3) {
        bres = true;
        k.sa = stages; k.sb = 0;
        op->smem_bytes = stages * k.a_bytes + b_total + fixed;
        return true;
      }
    }
    if (vreuse && d->ksize == 3 && want_bgroup3 && g.block_n <= 128 && !w_batched) {
      const int stages = budget / (k.a_bytes + 3 * b_tap_bytes);
      if (stages >= 3) {
        k.bgroup = 3;
        k.sa = k.sb = stages > 6 ? 6 : stages;
        op->smem_bytes = k.sa * (k.a_bytes + 3 * b_tap_bytes) + fixed;
        return true;
      }
    }
    if (vreuse) {
      k.sa = 3;
      k.sb = (budget - k.sa * k.a_bytes) / b_tap_bytes;
      if (k.sb < 3) { k.sa = 2; k.sb = (budget - k.sa * k.a_bytes) / b_tap_bytes; }
      if (k.sb > kMaxStages) k.sb = kMaxStages;
      if (k.sb < 2) return false;
    } else {
      int stages = budget / (k.a_bytes + b_tap_bytes);
      if (stages > kMaxStages) stages = kMaxStages;
      if (stages > k.a_steps * 2) stages = k.a_steps * 2 > 2 ? k.a_steps * 2 : 2;
      if (stages < 2) return false;
      k.sa = k.sb = stages;
    }
    op->smem_bytes = k.sa * k.a_bytes + k.sb * b_tap_bytes + fixed;
    return true;
  };
  const bool want_b3 = getenv("GLSDET_CONV_NO_BGROUP") == nullptr;
  if (ts && !patch && !size_rings(want_b3)) ts = false;   // no room for the staging tiles: direct stores
  if (!size_rings(want_b3)) {
    if (pred_mma) { pred_mma = false; }   // the staged operand does not fit next to the rings: FMA prediction path
    if (!size_rings(want_b3)) { free(mem); set_error("conv_create: tile does not fit shared memory"); return 2; }
  }
  k.bres = bres ? 1 : 0;
  // prediction-MMA class: three accumulator slots (see ConvKParams::nacc); GLSDET_CONV_NACC=2 restores two
  if (pred_mma && mt == 1 && 3 * g.block_n <= 512 && !(getenv("GLSDET_CONV_NACC") && getenv("GLSDET_CONV_NACC")[0] == '2'))
    k.nacc = 3;
  // two epilogue groups on alternate work items: four accumulator slots give each group two, so the MMAs of a group's
  // next item run while it drains the current one (with two slots the MMA + commit latency is exposed once per item)
  {
    const bool will_egrp = egrp_want && ts && (g.block_n == 64 || g.block_n == 128);
    if (will_egrp && k.nacc == 2 && 4 * mt * g.block_n <= 512 && getenv("GLSDET_CONV_NO_NACC4") == nullptr) k.nacc = 4;
  }
  int cols = 32;
  while (cols < k.nacc * mt * g.block_n) cols <<= 1;
  k.tmem_cols = cols;
  k.pdl = (getenv("GLSDET_CONV_NO_PDL") == nullptr) ? 1 : 0;
  k.ts = ts ? 1 : 0;
  k.w_batched = w_batched ? 1 : 0;
  k.a_shared = d->src_shared > 0 ? d->src_shared : 0;
  k.a_div = (d->src_shared > 0 && d->src_shared_div > 0) ? d->src_shared_div : 0;
  k.patch = patch ? 1 : 0;

  k.bias = d->bias; k.act = d->act;
  // SiLU = h + h * tanh.approx(h) (one MUFU op, relative error of tanh 2^-11) on every 16-bit layer.  Round 2 first ran
  // the fp16-storage layers (strides 8-32, backbone) on an ex2-based ~1e-6 form because the tanh error is as large as the
  // fp16 rounding; measured on the parity report (profiles/r2_parity_report.txt) the network-level error against the fp32
  // oracle is the same to four digits from features (0.03 - 0.23 %) and 0.35 -> 0.38 % / 0.42 -> 0.48 % at the worst
  // image -> logits levels - a tenth of the 2e-2 bound - while the exact form cost 120 us per step (4 %).
  // GLSDET_CONV_EXACT_SILU=1 selects the exact form (silu_newton) everywhere.
  const bool exact_silu = getenv("GLSDET_CONV_EXACT_SILU") != nullptr;
  k.act_epi = (d->act == GLSDET_ACT_SILU && exact_silu) ? kActSiluExact : d->act;
  k.pre_res = d->pre_res; k.pre_shift = d->pre_shift; k.pre_ld = d->pre_ld;
  k.post_res = reinterpret_cast<const __nv_bfloat16*>(d->post_res); k.post_shift = d->post_shift; k.post_ld = d->post_ld;
  k.a_f16 = d->src_dtype == GLSDET_DT_F16; k.out_f16 = d->out_dtype == GLSDET_DT_F16; k.post_f16 = d->post_dtype == GLSDET_DT_F16;
  k.out = d->out; k.out_mode = d->out_mode; k.out_ld = d->out_ld; k.out_coff = d->out_coff; k.out_bs = d->out_batch_stride;
  k.out_plane = d->out_plane_stride > 0 ? d->out_plane_stride : static_cast<int64_t>(g.Ho) * g.Wo;
  k.dec_stride = d->dec_stride; k.dec_in_w = d->dec_in_w; k.dec_in_h = d->dec_in_h;
  {
    k.pred_w = d->pred_weight; k.pred_b = d->pred_bias; k.pred_n = d->pred_channels; k.pred_act = d->pred_act;
    if (fused_pred) k.epi = pred_mma ? EPI_TOWER_PRED_MMA : EPI_TOWER_PRED;
  }

  k.dbg = getenv("GLSDET_CONV_DBG") ? atoi(getenv("GLSDET_CONV_DBG")) : 0;
  k.trace = nullptr;
  if (getenv("GLSDET_CONV_TRACE") != nullptr) {
    void* tp = nullptr;
    if (cudaMalloc(&tp, 16 * sizeof(unsigned long long)) == cudaSuccess) {
      cudaMemset(tp, 0, 16 * sizeof(unsigned long long));
      k.trace = static_cast<unsigned long long*>(tp);
    }
  }
  int rc = 0;
  if (d->stride == 1) {
    rc = encode_act_map(&k.tmA[0], d->src0, d->src0_c, d->src0_ld, d->batch, d->height, d->width,
                        patch ? kPatchView : 1, best_w, box_rows, d->src0_row_pitch, d->src0_img_pitch);
    if (!rc && d->src1)
      rc = encode_act_map(&k.tmA[1], d->src1, d->src1_c, d->src1_ld, d->batch, d->height, d->width, 1, best_w, box_rows);
    else if (!rc) k.tmA[1] = k.tmA[0];
  } else if (g.kw == 2) {   // stride-2 conv over pixel pairs: one map, row parity in coordinate 2
    rc = encode_act_map(&k.tmA[0], d->src0, d->src0_c, d->src0_ld, d->batch, d->height, d->width, kRowStride2, best_w, box_rows);
    k.tmA[1] = k.tmA[0];
  } else {
    const __nv_bfloat16* s = reinterpret_cast<const __nv_bfloat16*>(d->src0);
    rc = encode_act_map(&k.tmA[0], s, d->src0_c, d->src0_ld, d->batch, d->height, d->width, 2, best_w, box_rows);
    if (!rc)
      rc = encode_act_map(&k.tmA[1], s + d->src0_ld, d->src0_c, d->src0_ld, d->batch, d->height, d->width, 2, best_w,
                          box_rows);
  }
  k.res_tma = (k.ts && res_tma) ? 1 : 0;
  k.res_post_bytes = k.res_tma ? res_post_bytes : 0;
  k.res_pre_half_bytes = k.res_tma ? res_pre_half : 0;
  k.res_slot_bytes = k.res_tma ? res_slot_bytes : 0;
  k.res_nbuf = res_nbuf;
  if (!rc && k.res_tma && d->post_res) {
    const int sh = d->post_shift;
    rc = encode_act_map(&k.tmRpost, d->post_res, d->out_channels, d->post_ld, d->batch, g.Ho >> sh, g.Wo >> sh, 1, best_w >> sh,
                        k.tile_h >> sh);
  }
  if (!rc && k.res_tma && d->pre_res) {   // fp32 rows as 2-byte elements: a 64-element box row = 32 floats = 128 bytes
    const int sh = d->pre_shift;
    rc = encode_act_map(&k.tmRpre, d->pre_res, 2 * d->out_channels, 2 * d->pre_ld, d->batch, g.Ho >> sh, g.Wo >> sh, 1,
                        best_w >> sh, k.tile_h >> sh);
  }
  if (!rc && k.ts) {
    const __nv_bfloat16* obase = reinterpret_cast<const __nv_bfloat16*>(d->out) + d->out_coff;
    rc = encode_act_map(&k.tmO, obase, d->out_channels, d->out_ld, d->batch, g.Ho, g.Wo, patch ? kPatchView : 1, best_w,
                        k.tile_h);
  }
  if (!rc) {
    EncodeTiledFn enc = get_encode_tiled();
    const int w_ld = d->weight_ld > 0 ? d->weight_ld : g.k_pad;
    if (w_ld < g.k_pad) { set_error("conv_create: weight_ld %d is smaller than the padded K %d", w_ld, g.k_pad); rc = 2; }
    // shared weights: [n_pad][w_ld]; per-image weights: one more dimension (image), rows beyond N read as zero
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(g.k_pad), static_cast<cuuint64_t>(w_batched ? d->out_channels : g.n_pad),
                          static_cast<cuuint64_t>(d->batch)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(w_ld) * 2, static_cast<cuuint64_t>(d->weight_batch_stride) * 2};
    const cuuint32_t b_rows = static_cast<cuuint32_t>(op->two_cta ? g.block_n / 2 : g.block_n);
    cuuint32_t box[3] = {static_cast<cuuint32_t>(kChunkK), b_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = rc ? CUDA_ERROR_INVALID_VALUE
                    : enc(&k.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, w_batched ? 3 : 2, const_cast<void*>(d->weight), dims,
                          strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(weight) failed with CUresult %d", (int)r);
      rc = 2;
    }
    k.tmB3[0] = k.tmB;
    k.tmB3[1] = k.tmB;
    for (int src = 0; src < 2 && rc == 0 && k.bgroup == 3; ++src) {
      const int nch = src ? g.chunks1 : g.chunks0;
      if (nch == 0) continue;
      // (k, n, ky): ky advances three taps = 3 * nch chunks of 64 columns
      cuuint64_t dims3[3] = {static_cast<cuuint64_t>(g.k_pad), static_cast<cuuint64_t>(g.n_pad), 3};
      cuuint64_t strides3[2] = {static_cast<cuuint64_t>(g.k_pad) * 2, static_cast<cuuint64_t>(g.kw) * nch * kChunkK * 2};
      cuuint32_t box3[3] = {static_cast<cuuint32_t>(kChunkK), b_rows, 3};
      cuuint32_t estr3[3] = {1, 1, 1};
      r = enc(&k.tmB3[src], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(d->weight), dims3, strides3, box3,
              estr3, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {  // not representable: fall back to one 2-D box per ky tap
        k.bgroup = 1;
        if (!size_rings(false)) { set_error("conv_create: tile does not fit shared memory"); rc = 2; }
        break;
      }
    }
  }
  if (rc) { free(mem); return rc; }

  switch (k.epi) {
    case EPI_BF16: case EPI_BF16_PRE: case EPI_BF16_POST: case EPI_BF16_PREPOST: op->ec = k.ts ? EC_BF16_TS : EC_BF16; break;
    case EPI_F32_PLAIN: op->ec = EC_F32; break;
    case EPI_NCHW_RAW: case EPI_ROWS_BOX: case EPI_ROWS_SIGMOID: op->ec = EC_SMALL; break;
    case EPI_TOWER_PRED: op->ec = EC_PRED_FMA; break;
    case EPI_TOWER_PRED_MMA: op->ec = EC_PRED_MMA; break;
    default: op->ec = EC_ALL;
  }
  if (op->ec == EC_BF16_TS && (g.block_n == 64 || g.block_n == 128) && getenv("GLSDET_CONV_EPI8") == nullptr)
    op->ec = k.res_tma ? EC_BF16_TS16R : EC_BF16_TS16;
  // the second tower conv + prediction MMA: its epilogue (activate, stage, prediction MMA round trip, decode) is as long
  // as the MMAs of two tiles with two warps per scheduler; four warps per scheduler halve the per-thread work
  if (op->ec == EC_PRED_MMA && (g.block_n % 64) == 0 && getenv("GLSDET_CONV_EPI8") == nullptr) op->ec = EC_PRED_MMA16;
  if (getenv("GLSDET_CONV_ONE_KERNEL") != nullptr) op->ec = EC_ALL;
  k.hbias = (k.act_epi == GLSDET_ACT_SILU &&
             (k.epi == EPI_BF16 || k.epi == EPI_BF16_PRE || k.epi == EPI_BF16_POST || k.epi == EPI_BF16_PREPOST ||
              k.epi == EPI_TOWER_PRED_MMA)) ? 1 : 0;
  if (op->two_cta)   // the pair kernel: one all-epilogue binary, plus the prediction class with 16 epilogue warps
    op->ec = getenv("GLSDET_CONV_ONE_KERNEL") != nullptr ? EC2_ALL
             : op->ec == EC_PRED_MMA16 ? EC2_PRED_MMA16 : op->ec == EC_BF16_TS16 ? EC2_BF16_TS16
             : op->ec == EC_BF16_TS16R ? EC2_BF16_TS16R : EC2_ALL;
  k.egrp = (egrp_want && (op->ec == EC_BF16_TS16 || op->ec == EC2_BF16_TS16 || op->ec == EC_BF16_TS16R || op->ec == EC2_BF16_TS16R)) ? 1 : 0;
  if (k.res_tma && !k.egrp) { free(mem); set_error("conv_create: internal: residual staging without the two-group epilogue"); return 2; }
  op->threads = (op->ec == EC_BF16_TS16 || op->ec == EC_PRED_MMA16 || op->ec == EC2_PRED_MMA16 || op->ec == EC2_BF16_TS16 ||
                 op->ec == EC_BF16_TS16R || op->ec == EC2_BF16_TS16R)
                    ? (4 + 16) * 32 : kThreads;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaSuccess;
    for (int ec = 0; ec < EC_KERNELS && e == cudaSuccess; ++ec)
      e = cudaFuncSetAttribute(conv_kernel_for(ec), cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (e != cudaSuccess) {
      free(mem);
      set_error("cudaFuncSetAttribute(max dynamic smem) failed: %s", cudaGetErrorString(e));
      return 1;
    }
    attr_set[dev] = true;
  }
  if (op->two_cta) {
    const int clusters = k.total_pairs < sms / 2 ? k.total_pairs : sms / 2;
    op->grid = 2 * clusters;
  } else {
    op->grid = k.total_tiles < sms ? k.total_tiles : sms;
  }
  *out_op = op;
  return 0;
}

extern "C" int glsdet_conv_launch(glsdet_conv_t* op, void* stream) {
  GLSDET_REQUIRE(op != nullptr, "conv_launch: null op");
  if (op->two_cta) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(op->grid);
    cfg.blockDim = dim3(op->threads);
    cfg.dynamicSmemBytes = op->smem_bytes;
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = op->kp.pdl ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_kernel_for(op->ec), op->kp);
    if (e != cudaSuccess) {
      set_error("cudaLaunchKernelEx(conv_gemm_kernel<2cta>) failed: %s", cudaGetErrorString(e));
      return 1;
    }
    return count_launch("conv_gemm_kernel<2cta>");
  }
  if (op->kp.pdl) {
    // programmatic dependent launch: the prologue (barrier init, TMEM allocation, descriptor prefetch) overlaps the
    // tail of the previous kernel of the stream; the kernel itself waits (griddepcontrol.wait) before touching memory
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(op->grid);
    cfg.blockDim = dim3(op->threads);
    cfg.dynamicSmemBytes = op->smem_bytes;
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_kernel_for(op->ec), op->kp);
    if (e != cudaSuccess) {
      set_error("cudaLaunchKernelEx(conv_gemm_kernel, PDL) failed: %s", cudaGetErrorString(e));
      return 1;
    }
    return count_launch("conv_gemm_kernel");
  }
  conv_kernel_for(op->ec)<<<op->grid, op->threads, op->smem_bytes, static_cast<cudaStream_t>(stream)>>>(op->kp);
  return count_launch("conv_gemm_kernel");
}

extern "C" void glsdet_conv_destroy(glsdet_conv_t* op) {
  if (op) {
    if (op->kp.trace) cudaFree(op->kp.trace);
    free(op);
  }
}

extern "C" int glsdet_conv_read_trace(glsdet_conv_t* op, uint64_t* out16) {
  GLSDET_REQUIRE(op != nullptr && out16 != nullptr, "conv_read_trace: null argument");
  GLSDET_REQUIRE(op->kp.trace != nullptr, "conv_read_trace: the op was not created with GLSDET_CONV_TRACE=1");
  GLSDET_CHECK_CUDA(cudaMemcpy(out16, op->kp.trace, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  GLSDET_CHECK_CUDA(cudaMemset(op->kp.trace, 0, 16 * sizeof(uint64_t)));
  return 0;
}
