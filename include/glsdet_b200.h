/*
 * glsdet_b200 — C ABI of the B200-native GLSDet inference hot path
 * (PAFPN neck + FFA fusion + decoupled head + decode + score filter + class-aware NMS).
 *
 * The reference (WUTCM-Lab/GLSDet) is pure Python/PyTorch and has no FFI of its own; the
 * "plugin boundary" it offers is Python-level (yolox-drone/yolo.py:99-111 imports a module path
 * and calls YoloBody(num_classes, phi); mmdet builds NECKS/HEADS from the registry,
 * yolox-ufp/mmdet/models/builder.py:7-45).  This library is what the Python mirrors of those
 * modules (glsdet_b200/*.py) bind through ctypes.  Each entry point below names the reference
 * call it replaces.
 *
 * Conventions
 *   - plain C types only: raw device pointers, sizes, a cudaStream_t passed as void*.
 *   - every function returns 0 on success, non-zero on error; glsdet_last_error() gives the
 *     message of the last failing call of the calling thread.  No exceptions cross the ABI.
 *   - all launches are asynchronous on the caller's stream; no internal device synchronisation.
 *   - the library never allocates or frees caller tensors; op handles own only their descriptors.
 *   - activations inside the path are NHWC with a 16-bit storage type per tensor (GLSDET_DT_BF16 or GLSDET_DT_F16,
 *     channels contiguous), accumulation is fp32.  tcgen05.mma kind::f16 takes either operand type at the same rate;
 *     fp16 storage (3 more mantissa bits) is what keeps the 25-35 layer chains of the deep levels within the 2e-2
 *     parity bar, bf16 stays the type of the stride-4 level, where most of the FLOPs are and the chain is short.
 */
#ifndef GLSDET_B200_H_
#define GLSDET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLSDET_ABI_VERSION 2

/* activation applied in the conv epilogue (reference: models/base/activation.py:9-17, plus the
 * decode of models/core/utils_bbox.py:266-305 when the conv is a prediction conv) */
enum {
  GLSDET_ACT_NONE = 0,
  GLSDET_ACT_SILU = 1,       /* x * sigmoid(x)               activation.py:4-7 */
  GLSDET_ACT_RELU = 2,
  GLSDET_ACT_LRELU = 3,      /* leaky relu 0.1               activation.py:13 */
  GLSDET_ACT_SIGMOID = 4,    /* cls logits -> probabilities  utils_bbox.py:270 */
  GLSDET_ACT_YOLOX_BOX = 5,  /* ch0..3 -> (cx,cy,w,h) normalised, ch4 -> sigmoid(obj); utils_bbox.py:270-305 */
  GLSDET_ACT_MMDET_BOX = 6   /* ch0..3 -> (cx,cy,w,h) in input pixels: xy*stride + prior, exp(wh)*stride; ch4 -> sigmoid
                                (yolox-ufp/mmdet/models/dense_heads/yolox_head.py:298-301, priors of
                                core/anchor/point_generator.py with offset 0) */
};

/* 16-bit storage type of an activation / weight tensor */
enum { GLSDET_DT_BF16 = 0, GLSDET_DT_F16 = 1, GLSDET_DT_F32 = 2 /* accuracy mode; glsdet_dwconv only */ };

/* output layouts of the conv epilogue */
enum {
  GLSDET_OUT_NHWC_BF16 = 0,  /* out[b*bs + (y*W+x)*ld + coff + n]  16-bit (bf16, or fp16 with out_dtype = GLSDET_DT_F16) */
  GLSDET_OUT_NHWC_F32 = 1,   /* same addressing, fp32 */
  GLSDET_OUT_NCHW_F32 = 2    /* out[b*bs + (coff+n)*H*W + y*W + x] fp32 (reference tensor layout) */
};

typedef struct glsdet_conv glsdet_conv_t;

/*
 * One fused  conv(k x k, stride s, pad (k-1)/2) [+ bias] [+ residual] -> activation [+ residual] -> store.
 * Replaces BaseConv.forward (yolox-drone/models/base/baseConv.py:15-16) with the BatchNorm folded into
 * weight/bias, plain nn.Conv2d prediction convs (models/ffa/yolox_ffa.py:43-56), torch.cat of two
 * inputs in front of a conv (yolox_ffa.py:211,228,241,254; ffa.py:79), the residual adds of
 * yolox_ffa.py:73 and ffa.py:83, and - for prediction convs - decode_outputs (utils_bbox.py:254-306).
 */
typedef struct glsdet_conv_desc {
  /* input: up to two NHWC bf16 sources of identical B,H,W concatenated along channels */
  const void* src0;      /* device pointer to channel 0 of the view */
  int32_t src0_c;        /* channels used from src0 */
  int32_t src0_ld;       /* pixel pitch of src0 in elements (>= src0_c) */
  const void* src1;      /* NULL when there is a single source */
  int32_t src1_c;
  int32_t src1_ld;
  int32_t batch, height, width; /* input spatial size */
  int32_t ksize;         /* 1 or 3 */
  int32_t stride;        /* 1 or 2 (2 only with ksize 3, one source, even height/width) */
  /* weights: bf16 [n_pad][k_pad] row-major; K order = (source, tap=ky*3+kx, channel) with every
   * (source, tap) segment zero-padded to a multiple of 64 channels; see glsdet_conv_weight_shape() */
  const void* weight;
  int32_t out_channels;  /* N */
  const float* bias;     /* [N] fp32 or NULL */
  int32_t act;           /* GLSDET_ACT_* */
  /* optional fp32 NHWC tensor added BEFORE the activation, read at (y>>shift, x>>shift): used to add the
   * low-resolution half of a conv over cat(upsample(a), b) (1x1 convs commute with nearest upsampling) */
  const float* pre_res;
  int32_t pre_shift;
  int32_t pre_ld;
  /* optional bf16 NHWC tensor added AFTER the activation, read at (y>>shift, x>>shift) */
  const void* post_res;
  int32_t post_shift;
  int32_t post_ld;
  /* output */
  void* out;
  int32_t out_mode;      /* GLSDET_OUT_* */
  int32_t out_ld;        /* pixel pitch (NHWC modes) or total channel count (NCHW mode) */
  int32_t out_coff;      /* first output channel inside the destination */
  int64_t out_batch_stride; /* elements between images in the destination */
  /* GLSDET_ACT_YOLOX_BOX only: stride of this level and network input size (utils_bbox.py:285,303-304) */
  float dec_stride, dec_in_w, dec_in_h;
  /*
   * Optional fused prediction conv (yolox_ffa.py:88,100,109: cls_preds / reg_preds / obj_preds applied to the
   * output of the second tower conv).  When pred_weight != NULL the activated tile of THIS conv is not stored;
   * instead  y[j] = pred_bias[j] + sum_k act(conv)[k] * pred_weight[j][k]  (j < pred_channels <= 16) is computed in
   * the epilogue, `pred_act` (NONE, SIGMOID, YOLOX_BOX or MMDET_BOX) is applied and y is written through out / out_mode /
   * out_ld / out_coff / out_batch_stride (NHWC_F32 rows or NCHW_F32 planes).  Requires out_channels <= 256.
   */
  const float* pred_weight;  /* fp32 [pred_channels][out_channels] */
  const float* pred_bias;    /* fp32 [pred_channels] */
  int32_t pred_channels;
  int32_t pred_act;
  /*
   * Batched-GEMM extensions.  They carry Non_local_Block (yolox-drone/models/new/Non_local_family.py:6-48) and
   * Patch_Conv_NonLocal_new (:204-250) in the reassociated form  out = x + W_eff(b, patch) x + b_eff(b, patch)
   * (DESIGN.md section 3.4): the Gram matrix of a patch, the two C x C products that turn it into W_eff and the
   * final per-patch 1x1 conv are all this operator with a weight matrix PER IMAGE.
   */
  int64_t weight_batch_stride; /* elements between the weight matrices of consecutive images; 0 = one shared matrix */
  int32_t weight_ld;           /* row pitch of the weight matrix in elements; 0 = k_pad */
  int32_t src_shared;          /* k > 0: image b reads image (b mod k) of src0 (k static matrices used as activations,
                                  e.g. one per patch position) */
  int32_t src_shared_div;      /* with src_shared = k: d > 0 selects image (b / d) instead of (b mod k) (images grouped
                                  by patch position: b = position * images + image) */
  int32_t patch_mode;          /* 1 (1x1 convs): src0, post_res and out are the 2x2 patch views of
                                  [batch/4, 2*height, 2*width, ld] tensors; image b' = (b*2 + py)*2 + px is patch
                                  (py, px) of image b (the split of Non_local_family.py:230-233) */
  /*
   * Few-channel 3x3 convs (the Focus stem of CSPDarknet, yolox-drone/models/ffa/darknet.py:12-21: 12 input channels).
   * ksize_w = 1 folds the kx taps into the channel dimension: src0 is a zero-bordered buffer whose pixels are
   * src0_ld elements apart, the conv reads src0_c (<= 64, > src0_ld) consecutive elements starting at the pixel LEFT of
   * the output position - i.e. "channel" k = kx * src0_ld + c of the (kx, c) pair - and only runs the ky taps
   * (K = ksize * 64 instead of ksize^2 * 64).  src0 then points at padded pixel 0 of row 0 (the left border pixel),
   * src0_row_pitch / src0_img_pitch give the padded row / image pitches in elements, and the weight matrix is packed
   * as a ksize x 1 kernel over src0_c channels.
   */
  int32_t ksize_w;             /* 0 = ksize (square kernel); 1 = kx taps folded into the channel view (stride 1);
                                  2 (ksize 3, stride 2) = 3x3 stride-2 conv over PIXEL PAIRS: src0 is the [B, H, W/2, 2C] view of
                                  a [B, H, W, C] tensor (width = W/2, src0_c = 2C), rows are strided by 2, columns are dense
                                  with taps dx = -1 (only the right pixel of the left pair has non-zero weights) and 0; the
                                  output is [B, H/2, W/2, N].  K = 6 * 2C instead of 9 * 64 for C = 32, dense TMA boxes. */
  int64_t src0_row_pitch;      /* elements between rows of src0; 0 = width * src0_ld */
  int64_t src0_img_pitch;      /* elements between images of src0; 0 = height * row pitch */
  int64_t out_plane_stride;    /* GLSDET_OUT_NCHW_F32: elements between channel planes; 0 = Ho * Wo.  With the fused
                                  prediction conv and a non-zero stride the pred_act (sigmoid / box decode) is applied too:
                                  decoded predictions as planes [B][5+nc][A], `out` pointing at this level's first anchor */
  /* 16-bit storage types (GLSDET_DT_*; 0 = bf16 keeps older callers valid): src_dtype covers src0, src1 AND the packed
   * weights (one MMA operand format per conv), out_dtype the GLSDET_OUT_NHWC_BF16 output, post_dtype the post residual.
   * fp16 stores saturate at +-65504. */
  int32_t src_dtype;
  int32_t out_dtype;
  int32_t post_dtype;
} glsdet_conv_desc;

/* library / device */
int glsdet_abi_version(void);
const char* glsdet_last_error(void);
/* number of kernels launched by this library since load (all threads); bench.py reports it */
int64_t glsdet_launch_count(void);

/* rows/cols (n_pad, k_pad) of the packed weight matrix a descriptor needs, and the N tile used */
int glsdet_conv_weight_shape(const glsdet_conv_desc* desc, int32_t* n_pad, int32_t* k_pad, int32_t* block_n);
int glsdet_conv_create(const glsdet_conv_desc* desc, glsdet_conv_t** op);
int glsdet_conv_launch(glsdet_conv_t* op, void* stream);
void glsdet_conv_destroy(glsdet_conv_t* op);
/* Diagnostic: ops created while GLSDET_CONV_TRACE=1 record %globaltimer stamps (ns) of CTA 0 - 0 kernel entry, 1 after
 * griddepcontrol.wait, 2 set-up done, 3 first TMA load issued, 4 first operand stage landed, 5 last MMA committed,
 * 6 first accumulator ready, 7 epilogue done, 8 bulk stores drained, 9 CTA joined, 10 TMEM released.  Copies the 16
 * slots to the host (synchronous) and clears them.  Not part of the reference's interface. */
int glsdet_conv_read_trace(glsdet_conv_t* op, uint64_t* out16);

/*
 * fp32 evaluation of the same operator (accuracy mode: BASELINE.json configs[0], parity bar 1e-3 relative in fp32).
 * Same semantics as glsdet_conv_desc with every tensor fp32: src NHWC fp32 channel windows, weight fp32 [N][K] with
 * K order (source, tap = ky*k + kx, channel) unpadded (BatchNorm folded in by the caller), fp32 residuals, output
 * GLSDET_OUT_NHWC_F32 or GLSDET_OUT_NCHW_F32.  SIMT fp32 FMA implicit GEMM; no tensor cores, no bf16 anywhere.
 */
typedef struct glsdet_conv_f32_desc {
  const float* src0; int32_t src0_c, src0_ld;
  const float* src1; int32_t src1_c, src1_ld;      /* NULL when there is a single source */
  int32_t batch, height, width, ksize, stride;     /* ksize 1/3/5/7, pad (k-1)/2, stride 1/2 */
  const float* weight; int32_t out_channels;
  const float* bias; int32_t act;                  /* GLSDET_ACT_* (exact expf-based SiLU / sigmoid) */
  const float* pre_res; int32_t pre_shift, pre_ld;   /* added before the activation, read at (y>>shift, x>>shift) */
  const float* post_res; int32_t post_shift, post_ld; /* added after the activation */
  float* out; int32_t out_mode, out_ld, out_coff; int64_t out_batch_stride;
  float dec_stride, dec_in_w, dec_in_h;
} glsdet_conv_f32_desc;
int glsdet_conv_f32(const glsdet_conv_f32_desc* desc, void* stream);
/*
 * Batched fp32 GEMM of the accuracy mode: C[b][m][n] = alpha * sum_k A[b](m, k) * B[b][k][n], A stored [M][K] (a_trans = 0)
 * or [K][M] (a_trans = 1), B [K][N], C [M][N]; leading dimensions and batch strides in elements, so NHWC fp32 channel
 * windows are operands as they are.  The two products of the dot-product non-local block (yolox-drone/models/new/
 * Non_local_family.py:27-31,40-45: P = theta^T phi / T, y = P g, evaluated as theta (phi^T g / T)) in fp32 plans of the P1 /
 * P2 models.  SIMT FMA, fixed summation order.
 */
int glsdet_bgemm_f32(const float* a, int32_t a_trans, int32_t lda, int64_t a_batch_stride, const float* b, int32_t ldb,
                     int64_t b_batch_stride, float* c, int32_t ldc, int64_t c_batch_stride, int32_t m, int32_t n, int32_t k,
                     float alpha, int32_t batch, void* stream);
/* NCHW fp32 <-> channel window of an NHWC fp32 buffer (to_nhwc = 1: src is NCHW; 0: src is the NHWC buffer) */
int glsdet_nchw_nhwc_f32(const float* src, float* dst, int32_t batch, int32_t channels, int32_t height, int32_t width,
                         int32_t nhwc_ld, int32_t nhwc_coff, int32_t to_nhwc, void* stream);
/* fp32 variants of the FFA helpers (SE stage 1 on fp32 NHWC; stage 2 = glsdet_se_fc; gate * PixelShuffle) */
int glsdet_se_partial_f32(const float* x, int32_t batch, int32_t hw, int32_t channels, int32_t x_ld, float* scratch,
                          void* stream);
int glsdet_se_fc(const float* scratch, const float* w1, const float* w2, int32_t hidden, float* gate, int32_t batch,
                 int32_t hw, int32_t channels, void* stream);
int glsdet_scale_pixel_shuffle_f32(const float* x, const float* gate, float* dst, int32_t batch, int32_t height,
                                   int32_t width, int32_t out_channels, int32_t dst_ld, int32_t dst_coff, void* stream);

/* layout converters at the module boundary (reference tensors are NCHW fp32 everywhere) */
int glsdet_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int32_t batch, int32_t channels, int32_t height,
                                 int32_t width, int32_t dst_ld, int32_t dst_coff, void* stream);
int glsdet_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int32_t batch, int32_t channels, int32_t height,
                                 int32_t width, int32_t src_ld, int32_t src_coff, void* stream);

/* Same converters with the 16-bit storage type of the NHWC side given explicitly (GLSDET_DT_*). */
int glsdet_nchw_f32_to_nhwc_16(const float* src, void* dst, int32_t batch, int32_t channels, int32_t height,
                               int32_t width, int32_t dst_ld, int32_t dst_coff, int32_t dtype, void* stream);
int glsdet_nhwc_16_to_nchw_f32(const void* src, float* dst, int32_t batch, int32_t channels, int32_t height,
                               int32_t width, int32_t src_ld, int32_t src_coff, int32_t dtype, void* stream);

/*
 * Non-local helpers (yolox-drone/models/new/Non_local_family.py:32-48, 229-250).
 * glsdet_patch_transpose: the four 2x2 patches of an NCHW fp32 map (split of :230-233, equal halves only) as
 *   per-patch transposed bf16 matrices dst[(b*2+py)*2+px][c][t], t = y'*(W/2) + x', row pitch dst_ld, dst_rows rows
 *   per patch (rows >= channels - the ones row that yields the channel sums, zero padding - are the caller's).
 *   This is the K-major operand of the Gram product X^T X that replaces theta^T phi / T followed by (.) g.
 * glsdet_gather_bias: bias[b][n] = base[b mod base_groups][n] + w[b][n][col] for a bf16 matrix batch (row pitch
 *   ld): the per-patch effective bias W_o M^T b_theta + b_o of conv_out(y) (:45-46); base holds b_o of the
 *   base_groups (= 4) patch positions.
 */
int glsdet_patch_transpose(const float* src, void* dst, int32_t batch, int32_t channels, int32_t height,
                           int32_t width, int32_t dst_rows, int32_t dst_ld, void* stream);
int glsdet_gather_bias(const void* w, const float* base, float* bias, int32_t batch, int32_t n_rows, int32_t ld,
                       int32_t col, int64_t batch_stride, int32_t base_groups, void* stream);
/* explicit 16-bit storage type (GLSDET_DT_*) of dst / w */
/* (every element multiplied by `scale` before rounding: 1 / sqrt(T) on both Gram operands yields the MEAN outer product) */
int glsdet_patch_transpose_16(const float* src, void* dst, int32_t batch, int32_t channels, int32_t height,
                              int32_t width, int32_t dst_rows, int32_t dst_ld, int32_t dtype, float scale, void* stream);
int glsdet_gather_bias_16(const void* w, const float* base, float* bias, int32_t batch, int32_t n_rows, int32_t ld,
                          int32_t col, int64_t batch_stride, int32_t base_groups, int32_t dtype, void* stream);

/*
 * Rectangle copies between NHWC bf16 tensors (up to 8 per launch): the patch split, the seam halves and the re-tiling
 * of Patch_Conv / Patch_Conv_NonLocal (yolox-drone/models/block/non_local/Identity_Conv.py:292-318, 353-384) as data
 * movement:  dst[db + b, dy + y, dx + x, dcoff + c] = src[sb + b, sy + y, sx + x, scoff + c]  for b < batch, y < h,
 * x < w, c < channels (channels, offsets and pitches multiples of 8).
 */
typedef struct glsdet_rect {
  int32_t sb, sy, sx, db, dy, dx, h, w;
} glsdet_rect;
int glsdet_rect_copy(const void* src, int32_t src_h, int32_t src_w, int32_t src_ld, int32_t src_coff, void* dst,
                     int32_t dst_h, int32_t dst_w, int32_t dst_ld, int32_t dst_coff, int32_t batch, int32_t channels,
                     const glsdet_rect* rects, int32_t num_rects, void* stream);
/*
 * NHWC bf16 [B, T pixels, ld] channel window -> per-image transposed matrices dst[b][c][t] (row pitch dst_ld, dst_rows
 * rows per image): the Gram operand of the non-local block when its input is already an NHWC activation
 * (Patch_Conv_NonLocal, Identity_Conv.py:364-367).
 */
int glsdet_nhwc_transpose(const void* src, void* dst, int32_t batch, int32_t pixels, int32_t channels, int32_t src_ld,
                          int32_t src_coff, int32_t dst_rows, int32_t dst_ld, void* stream);
/* explicit 16-bit storage type (GLSDET_DT_*) and a scale applied to every element (1.0 = plain copy) */
int glsdet_nhwc_transpose_16(const void* src, void* dst, int32_t batch, int32_t pixels, int32_t channels, int32_t src_ld,
                             int32_t src_coff, int32_t dst_rows, int32_t dst_ld, int32_t dtype, float scale, void* stream);

/*
 * nn.Upsample(scale_factor=2, mode="nearest") of an NHWC bf16 channel window into a channel window of a concat
 * buffer (yolox-drone/models/new/yolox10.py:95,97: the upsampled neighbour level inside the cls-branch concat, where a
 * 3x3 conv follows, so the upsampling cannot be folded into the conv as it is for the 1x1 convs of the neck).
 */
int glsdet_upsample2x(const void* src, void* dst, int32_t batch, int32_t height, int32_t width, int32_t channels,
                      int32_t src_ld, int32_t src_coff, int32_t dst_ld, int32_t dst_coff, void* stream);

/*
 * Depthwise half of the reference's DWConv (yolox-drone/models/base/baseConv.py:22-30: dconv = BaseConv(C, C, k, stride,
 * groups = C) followed by pconv = 1x1 BaseConv; phi = 'nano' builds every k > 1 conv outside Focus this way -
 * models/ffa/yolox_ffa.py:15,125, models/ffa/darknet.py:48,120).  dst[b, oy, ox, dst_coff + c] =
 * act(bias[c] + sum_{ky,kx} src[b, oy*stride + ky - pad, ox*stride + kx - pad, src_coff + c] * weight[ky*k + kx][c]),
 * pad = (k - 1) / 2, zero padding, output size (H + 2 pad - k) / stride + 1.  src / dst: NHWC channel windows of the
 * storage type `dtype` (GLSDET_DT_BF16 / _F16, or GLSDET_DT_F32 in the accuracy mode); weight fp32 [k*k][channels]
 * with BatchNorm folded in by the caller, bias fp32 [channels]; act GLSDET_ACT_NONE / SILU / RELU / LRELU (exact
 * expf-based SiLU).  HBM-bound SIMT kernel; the pconv half runs on glsdet_conv_*.
 */
int glsdet_dwconv(const void* src, int32_t src_ld, int32_t src_coff, void* dst, int32_t dst_ld, int32_t dst_coff,
                  int32_t batch, int32_t height, int32_t width, int32_t channels, int32_t ksize, int32_t stride,
                  const float* weight, const float* bias, int32_t act, int32_t dtype, void* stream);

/*
 * CSPDarknet backbone pieces that are not convolutions (the backbone's convolutions run on
 * glsdet_conv_*):
 * glsdet_focus_nchw_f32_to_nhwc_bf16: Focus.forward's space-to-depth (yolox-drone/models/ffa/darknet.py:15-21) fused with
 *   the layout change of the image: image [B, 3, H, W] fp32 NCHW -> dst [B, H/2, W/2, 16] bf16 NHWC with channel
 *   q*3 + c = image[b, c, 2y + dy, 2x + dx], q = top-left, bottom-left, top-right, bottom-right (the reference's cat order);
 *   channels 12..15 are written as zeros (32-byte pixels for the TMA loads of the stem conv).  dst_border = 1: dst rows
 *   hold W/2 + 2 pixels and the kernel writes pixels 1..W/2 (the border pixels stay as the caller zeroed them: the
 *   horizontal padding of the kx-folded stem conv, glsdet_conv_desc.ksize_w).
 * glsdet_spp_maxpool: the MaxPool2d(5 / 9 / 13, stride 1, padding k/2) branches of SPPBottleneck.forward
 *   (darknet.py:28,33-36) on an NHWC bf16 buffer [B, H, W, ld]: reads the channel window at src_coff and writes the three
 *   pooled maps into the windows at coff5 / coff9 / coff13 of the same buffer (the concat the following 1x1 conv reads).
 *   H * W <= 6400 (the map of one image x 8 channels is processed in shared memory).
 */
int glsdet_focus_nchw_f32_to_nhwc_bf16(const float* image, void* dst, int32_t batch, int32_t height, int32_t width,
                                       int32_t dst_border, void* stream);
int glsdet_spp_maxpool(void* buf, int32_t batch, int32_t height, int32_t width, int32_t channels, int32_t ld,
                       int32_t src_coff, int32_t coff5, int32_t coff9, int32_t coff13, void* stream);
int glsdet_spp_maxpool_16(void* buf, int32_t batch, int32_t height, int32_t width, int32_t channels, int32_t ld,
                          int32_t src_coff, int32_t coff5, int32_t coff9, int32_t coff13, int32_t dtype, void* stream);
/* glsdet_focus_u8_to_nhwc_bf16: the same Focus output from a uint8 HWC image batch [B, H, W, 3] (host pointers mean[3], std[3]):
 *   the normalisation of yolox-drone/models/core/utils.py:47-51 preprocess_input (x / 255 in float32, then - mean and / std
 *   in double, rounded to float32 each time, as numpy does for the float64 constant arrays) and the HWC -> CHW transpose of
 *   yolo.py:134 are applied on the fly - bit-identical to preprocessing on the host, a quarter of the upload bytes
 *   (the normalisation part of the YOLO facade). */
int glsdet_focus_u8_to_nhwc_bf16(const uint8_t* image, void* dst, int32_t batch, int32_t height, int32_t width,
                                 int32_t dst_border, const double* mean, const double* std, void* stream);
/* the two Focus entry points with an explicit 16-bit storage type (GLSDET_DT_*) of dst */
int glsdet_focus_nchw_f32_to_nhwc_16(const float* image, void* dst, int32_t batch, int32_t height, int32_t width,
                                     int32_t dst_border, int32_t dtype, void* stream);
int glsdet_focus_u8_to_nhwc_16(const uint8_t* image, void* dst, int32_t batch, int32_t height, int32_t width,
                               int32_t dst_border, const double* mean, const double* std, int32_t dtype, void* stream);
/* fp32 accuracy-mode twins (every tensor fp32; Focus output [B, H/2, W/2, 12] without padding channels) */
int glsdet_focus_nchw_f32_to_nhwc_f32(const float* image, float* dst, int32_t batch, int32_t height, int32_t width,
                                      void* stream);
int glsdet_spp_maxpool_f32(float* buf, int32_t batch, int32_t height, int32_t width, int32_t channels, int32_t ld,
                           int32_t src_coff, int32_t coff5, int32_t coff9, int32_t coff13, void* stream);

/*
 * MP-Det head pieces (yolox-ufp/mmdet/models/dense_heads/mp_head.py, gfl_head.py; BASELINE configs[2]).
 * glsdet_group_norm_relu: in place x = relu(GroupNorm(x)) on NHWC bf16 (the norm + activation of mmcv ConvModule in the
 *   shared towers, mp_head.py:46-62); scratch = glsdet_group_norm_scratch_floats(batch, channels) floats,
 *   zero-initialised once by the caller (arrival counters at its end; every call leaves them at zero).
 * glsdet_proxy_scores: MPHead.forward_proxy (mp_head.py:105-121): feat fp32 NHWC [batch*hw, channels], centers = the
 *   L2-normalised proxies fp32 [num_proxies, channels], cls_start[num_classes + 1] = first proxy of each class;
 *   rows[b][row0 + pixel][class] = gamma * sum_j softmax(gamma * sim)_j * sim_j  (raw class scores).
 * glsdet_gfl_decode: Integral (gfl_head.py:35-49) * stride + DistancePointBBoxCoder.decode on the cell points
 *   (x * stride, y * stride) with clamping (core/bbox/transforms.py:136-165): reg fp32 NHWC [.., reg_ld] with
 *   4 * bins logits per pixel (Scale folded into the conv) -> boxes[b][row0 + pixel][4] xyxy.
 * glsdet_gfl_select: filter_scores_and_topk (core/utils/misc.py:143-165) for one level: (anchor, class) pairs with
 *   sigmoid(score) > score_thr, best topk by score, appended to the image's candidate list (cand_count must be zeroed
 *   by the caller before the first level); keys = scratch of 8 * keys_batch_stride bytes per image.
 */
int glsdet_group_norm_relu(void* x, int32_t batch, int32_t hw, int32_t channels, int32_t x_ld, int32_t groups,
                           const float* gamma, const float* beta, float eps, float* scratch, void* stream);
int64_t glsdet_group_norm_scratch_floats(int32_t batch, int32_t channels);
int glsdet_proxy_scores(const float* feat, const float* centers, const int32_t* cls_start, int32_t num_classes,
                        int32_t num_proxies, int32_t channels, int32_t batch, int32_t hw, float gamma, float* rows,
                        int32_t rows_ld, int64_t rows_batch_stride, int32_t row0, void* stream);
/* Second half of MPHead.forward_proxy (mp_head.py:105-121) when the similarities come from a 1x1 conv of the bf16 class
 * features with the normalised proxies (glsdet_conv_*, fp32 output sims[pixel][sims_ld]): L2 norm of the features and the
 * per-class softmax(gamma * s)-weighted sum, raw class scores into rows like glsdet_proxy_scores. */
int glsdet_proxy_aggregate(const void* feat, const float* sims, int32_t sims_ld, const int32_t* cls_start, int32_t num_classes,
                           int32_t channels, int32_t batch, int32_t hw, float gamma, float* rows, int32_t rows_ld,
                           int64_t rows_batch_stride, int32_t row0, void* stream);
int glsdet_gfl_decode(const float* reg, int32_t reg_ld, int32_t bins, int32_t batch, int32_t height, int32_t width,
                      float stride, float max_x, float max_y, float* boxes, int64_t boxes_batch_stride, int32_t row0,
                      void* stream);
/* scratch: glsdet_gfl_select_scratch_ints(batch) int32, zero-initialised once by the caller (the call leaves it zeroed) */
int64_t glsdet_gfl_select_scratch_ints(int32_t batch);
int glsdet_gfl_select(const float* rows, int32_t rows_ld, int64_t rows_batch_stride, const float* boxes,
                      int64_t boxes_batch_stride, int32_t row0, int32_t level_anchors, int32_t num_classes,
                      float score_thr, int32_t topk, int32_t batch, void* keys, int64_t keys_batch_stride,
                      int32_t* cand_count, float* cand_boxes, float* cand_scores, float* cand_labels,
                      int32_t cand_capacity, int32_t* scratch, void* stream);

/*
 * SE gate of FFA (yolox-drone/models/ffa/ffa.py:5-20,77): partial[b][p][c] = sum over a slab of pixels of
 * x[b,:,c]; gate[b][c] = 1 + sigmoid(W2 relu(W1 mean)).  Deterministic two-stage reduction.
 *   x: NHWC bf16 [B, HW, C]; w1: fp32 [C/r][C]; w2: fp32 [C][C/r]; gate: fp32 [B][C];
 *   scratch: fp32 [B][GLSDET_SE_SLABS][C]
 */
#define GLSDET_SE_SLABS 32
int glsdet_se_gate(const void* x, int32_t batch, int32_t hw, int32_t channels, int32_t x_ld, const float* w1,
                   const float* w2, int32_t hidden, float* scratch, float* gate, void* stream);
/*
 * t = PixelShuffle(2)(x * gate)  (ffa.py:77-78) on NHWC bf16 with x's channels pre-permuted to
 * (i, j, c) order (done once when the producing conv's weights are packed):
 *   dst[b, 2y+i, 2x+j, coff + c] = x[b, y, x, (2i+j)*C + c] * gate[b, (2i+j)*C + c]
 */
int glsdet_scale_pixel_shuffle(const void* x, const float* gate, void* dst, int32_t batch, int32_t height,
                               int32_t width, int32_t out_channels, int32_t dst_ld, int32_t dst_coff,
                               void* stream);
/* explicit 16-bit storage types (GLSDET_DT_*): x_dtype of the input, dst_dtype of the shuffled output */
int glsdet_se_gate_16(const void* x, int32_t batch, int32_t hw, int32_t channels, int32_t x_ld, const float* w1,
                      const float* w2, int32_t hidden, float* scratch, float* gate, int32_t x_dtype, void* stream);
int glsdet_scale_pixel_shuffle_16(const void* x, const float* gate, void* dst, int32_t batch, int32_t height,
                                  int32_t width, int32_t out_channels, int32_t dst_ld, int32_t dst_coff,
                                  int32_t x_dtype, int32_t dst_dtype, void* stream);

/*
 * decode_outputs (yolox-drone/models/core/utils_bbox.py:254-306) for callers that hold raw NCHW logits:
 * levels[l] = fp32 [B, 5+nc, h_l, w_l]; pred = fp32 [B, A, 5+nc] contiguous.
 */
int glsdet_decode_outputs(const float* const* levels, const int32_t* heights, const int32_t* widths,
                          int32_t num_levels, int32_t batch, int32_t num_classes, int32_t in_h, int32_t in_w,
                          float* pred, void* stream);

/*
 * The decode variants the YOLO facade dispatches on `decode_mode` (yolox-drone/yolo.py:75-82; utils_bbox.py:36-251):
 *   decode_outputs               SIGMOID_OBJ | SIGMOID_CLS | NORMALISE   (:254-306, 'default')
 *   decode_outputs_no_sigmoid    SIGMOID_OBJ | NORMALISE                 (:149-200, 'obj_sigmoid')
 *   decode_outputs_no_sigmoid_all              NORMALISE                 (:202-251, 'no_sigmoid')
 *   decode_outputs_cls_sigmoid   SIGMOID_CLS | NORMALISE                 (:95-147,  'cls_sigmoid')
 *   decode_outputs_xyxy          XYXY (input pixels, corner form, raw obj / cls logits)   (:36-93)
 */
enum { GLSDET_DECODE_SIGMOID_OBJ = 1, GLSDET_DECODE_SIGMOID_CLS = 2, GLSDET_DECODE_NORMALISE = 4, GLSDET_DECODE_XYXY = 8 };
int glsdet_decode_outputs_mode(const float* const* levels, const int32_t* heights, const int32_t* widths,
                               int32_t num_levels, int32_t batch, int32_t num_classes, int32_t in_h, int32_t in_w,
                               int32_t mode, float* pred, void* stream);

/*
 * mmdet flavour of the decode (yolox-ufp/mmdet/models/dense_heads/yolox_head.py:255-308 with the priors of
 * core/anchor/point_generator.py:148-175, offset 0): three lists of raw NCHW fp32 maps (cls [B,nc,h,w],
 * bbox [B,4,h,w], objectness [B,1,h,w]; pass the same base pointers with channel strides for views) ->
 * pred fp32 [B, A, 5+nc] rows (cx, cy, w, h in input pixels, sigmoid(obj), sigmoid(cls)).
 * cls_bs / box_bs / obj_bs: elements between images of each list entry (lets callers pass channel slices).
 */
int glsdet_decode_mmdet(const float* const* cls, const float* const* box, const float* const* obj,
                        const int64_t* cls_bs, const int64_t* box_bs, const int64_t* obj_bs, const int32_t* heights,
                        const int32_t* widths, const int32_t* strides, int32_t num_levels, int32_t batch,
                        int32_t num_classes, float* pred, void* stream);

/*
 * Post-processing: non_max_suppression (utils_bbox.py:375-484) without the per-image Python loop or host
 * synchronisation.  Works on decoded predictions pred[B][A][5+nc] (cx,cy,w,h,obj,cls...).
 *
 *   strategy: how class awareness is realised, mirroring torchvision.ops.boxes.batched_nms
 *     (third-party, torchvision 0.26.0; call site utils_bbox.py:414-419):
 *       GLSDET_NMS_COORD_TRICK  boxes + label*(max_coord+1) then class-agnostic NMS (_batched_nms_coordinate_trick)
 *       GLSDET_NMS_PER_CLASS    NMS inside each class on the raw boxes (_batched_nms_vanilla)
 *       GLSDET_NMS_AUTO_CUDA    trick iff 4*K <= 100000  (torchvision's dispatch for CUDA tensors)
 *       GLSDET_NMS_AUTO_CPU     trick iff 4*K <= 4000    (torchvision's dispatch for CPU tensors)
 *   Results are bit-identical to the chosen torchvision strategy when fed identical predictions.
 *
 *   det: fp32 [B][max_det][7] rows (x1,y1,x2,y2,obj_conf,class_conf,class_pred) sorted by score desc;
 *   det_count: int32 [B] number of valid rows per image (rows beyond max_det are dropped, count is clamped);
 *   keep_index: optional int32 [B][max_det] anchor index of every kept row (NULL to skip).
 */
enum {
  GLSDET_NMS_COORD_TRICK = 0,
  GLSDET_NMS_PER_CLASS = 1,
  GLSDET_NMS_AUTO_CUDA = 2,
  GLSDET_NMS_AUTO_CPU = 3,
  /* mmcv.ops.nms.batched_nms (mmcv-full 1.3.17 .. 1.5.0, the pin of yolox-ufp/requirements/mminstall.txt; call sites
   * dense_heads/yolox_head.py:321, base_dense_head.py:295): boxes are always shifted by label * (max + 1); one NMS
   * over all boxes when K < split_thr = 10000, otherwise one NMS per class on the shifted boxes */
  GLSDET_NMS_MMCV = 4
};
typedef struct glsdet_nms glsdet_nms_t;
/* bytes of device workspace needed for the given problem size */
int64_t glsdet_nms_workspace_bytes(int32_t batch, int32_t anchors, int32_t num_classes);
int glsdet_nms_create(int32_t batch, int32_t anchors, int32_t num_classes, int32_t max_det, void* workspace,
                      int64_t workspace_bytes, glsdet_nms_t** op);
int glsdet_nms_launch(glsdet_nms_t* op, const float* pred, float conf_thres, float nms_thres, int32_t strategy,
                      float* det, int32_t* det_count, int32_t* keep_index, void* stream);
/* Same with per-image divisors box_div[B][4] applied to the corner boxes (x1/d0, y1/d1, x2/d2, y2/d3) before
 * NMS and in the output rows: mmdet's `rescale` (dense_heads/yolox_head.py:283-285 divides the decoded corners by
 * scale_factor before _bboxes_nms).  box_div == NULL is glsdet_nms_launch. */
int glsdet_nms_launch_scaled(glsdet_nms_t* op, const float* pred, const float* box_div, float conf_thres,
                             float nms_thres, int32_t strategy, float* det, int32_t* det_count, int32_t* keep_index,
                             void* stream);
/* Same with an explicit memory layout of the decoded predictions:
 *   GLSDET_PRED_ROWS   [B][A][5+nc] contiguous rows (what .contiguous() of the reference's decode_outputs gives);
 *   GLSDET_PRED_PLANES [B][5+nc][A] - the PHYSICAL layout of the tensor yolox-drone's decode_outputs returns
 *                      (utils_bbox.py:266,306: cat of the flattened levels along dim 2, then permute(0, 2, 1) as a view),
 *                      and the one the fused prediction convs write (coalesced stores, coalesced filter reads);
 *   | GLSDET_PRED_CLS_LOGITS: the class columns hold raw logits and the score filter applies the sigmoid itself (once per
 *                      anchor and class, in a full-occupancy kernel) - the fused detect path leaves it out of the cls
 *                      prediction conv's epilogue, where ten exp + divide chains per pixel sat on four warps per SM. */
enum { GLSDET_PRED_ROWS = 0, GLSDET_PRED_PLANES = 1, GLSDET_PRED_CLS_LOGITS = 2 };
int glsdet_nms_launch_layout(glsdet_nms_t* op, const float* pred, int32_t layout, const float* box_div, float conf_thres,
                             float nms_thres, int32_t strategy, float* det, int32_t* det_count, int32_t* keep_index,
                             void* stream);
void glsdet_nms_destroy(glsdet_nms_t* op);

/* class-aware NMS on caller-supplied boxes (the exact contract of torchvision batched_nms on one image):
 * boxes fp32 [K][4] xyxy, scores fp32 [K], labels fp32 [K]; keep int32 [K] (first *keep_count valid, score desc) */
int glsdet_batched_nms(const float* boxes, const float* scores, const float* labels, int32_t k, float nms_thres,
                       int32_t strategy, void* workspace, int64_t workspace_bytes, int32_t* keep,
                       int32_t* keep_count, void* stream);
/* Same with dense class ids: label_ids int32 [K] in 0..255 (one id per distinct label value) drive the class segments,
 * while the coordinate-trick offsets keep using the original label values (torchvision: offsets = idxs * (max + 1)).
 * glsdet_batched_nms casts the labels themselves to class ids, i.e. needs integral labels in 0..255; arbitrary idxs
 * (negative, batch * nc + cls, ...) go through this entry point.  label_abs_max: max |label| when every label is
 * integral (the coordinate trick is then evaluated class by class, which is identical as long as the offsets stay below
 * 2^21), negative when they are not (the trick then runs literally, as one class-agnostic NMS on the shifted boxes). */
int glsdet_batched_nms_ids(const float* boxes, const float* scores, const float* labels, const int32_t* label_ids,
                           float label_abs_max, int32_t k, float nms_thres, int32_t strategy, void* workspace, int64_t workspace_bytes,
                           int32_t* keep, int32_t* keep_count, void* stream);
int64_t glsdet_batched_nms_workspace_bytes(int32_t k);
/* `batch` independent problems of glsdet_batched_nms_ids in one launch sequence (mmdet's per-image _bbox_post_process /
 * _bboxes_nms loop, base_dense_head.py:276-301, yolox_head.py:286-294): every array is [batch][k], label_ids in
 * 0..num_ids-1 (num_ids <= 256), keep int32 [batch][k], keep_count int32 [batch]. */
int64_t glsdet_batched_nms_batch_workspace_bytes(int32_t batch, int32_t k, int32_t num_ids);
int glsdet_batched_nms_ids_batch(const float* boxes, const float* scores, const float* labels, const int32_t* label_ids,
                                 float label_abs_max, int32_t num_ids, int32_t batch, int32_t k, float nms_thres,
                                 int32_t strategy, void* workspace, int64_t workspace_bytes, int32_t* keep, int32_t* keep_count,
                                 void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * UFP stage of UFPMP-Det (reference: /root/reference/yolox-ufp).
 *
 * glsdet_ufp_pack (HOST function, no device work): replaces mmdet.core.UnifiedForegroundPacking(bbox_list, scale,
 *   input_shape) (mmdet/core/ufp/unified_foreground_packing.py:185-197 = scale_boxes :6-32 + ForegroundRegionGeneration
 *   :68-103 + Packing :140-181 over spp.py:69-168 phsppog).  boxes: n x 4 float32 (x1, y1, x2, y2) coarse detections;
 *   rows: n x 7 doubles out [x0, y0, w, h, new_x, new_y, factor]; returns the row count and the mosaic extent.
 * glsdet_ufp_mosaic: replaces display_merge_result (ufpmp_det_eval.py:182-193): image uint8 HWC (what cv2.imread
 *   returns) on the device, chips = n x 7 int32 (floor of the rows), canvas uint8 [can_h, can_w, 3] zero-filled, every
 *   chip cropped, resized by its integer factor exactly like cv2.resize(INTER_LINEAR) and pasted.
 * glsdet_ufp_merge: replaces the map-back loop + per-class py_cpu_nms (ufpmp_det_eval.py:270-306, :149-179).  dets:
 *   [K, 5] float32 second-stage detections (x1, y1, x2, y2, score) grouped by class, cls_off[num_classes + 1] their
 *   class offsets; mapped / out: [num_classes, cap, 5] float32 (cap <= 4096); out rows are the kept detections of each
 *   class in score order, out_count[c] their number, mapped_count[c] the number of mapped detections (> cap = overflow). */
int glsdet_ufp_pack(const float* boxes, int32_t n, float scale, int32_t in_w, int32_t in_h, double* rows, int32_t* n_rows,
                    double* new_w, double* new_h);
int glsdet_ufp_mosaic(const uint8_t* image, int32_t img_h, int32_t img_w, const int32_t* chips, int32_t n_chips,
                      uint8_t* canvas, int32_t can_h, int32_t can_w, void* stream);
int glsdet_ufp_merge(const float* dets, const int32_t* cls_off, int32_t num_classes, const int32_t* chips, int32_t n_chips,
                     float nms_thresh, float* mapped, int32_t cap, float* out, int32_t* out_count, int32_t* mapped_count,
                     void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * YOLO facade: resize_image (yolox-drone/models/core/utils.py:21-34; yolo.py:130) = PIL Image.resize(size, BICUBIC),
 * optionally letterboxed onto a (128, 128, 128) canvas, on the device and bit-exact to Pillow's 8-bit resampler.
 *   glsdet_pil_bicubic_ksize / _table (HOST functions): window size and per-output-pixel (first tap, tap count) bounds +
 *     22-bit fixed-point weights [out_size][ksize] for resampling in_size -> out_size samples.
 *   glsdet_resize_bicubic_u8: image uint8 HWC [in_h, in_w, 3] on the device -> canvas uint8 [can_h, can_w, 3]; the
 *     resized image (out_h x out_w) lands at (off_y, off_x), the rest of the canvas is `fill`; tmp: [in_h, out_w, 3]
 *     intermediate of the horizontal pass (needed when both sizes change). */
int glsdet_pil_bicubic_ksize(int32_t in_size, int32_t out_size);
int glsdet_pil_bicubic_table(int32_t in_size, int32_t out_size, int32_t* bounds, int32_t* kk);
int glsdet_resize_bicubic_u8(const uint8_t* image, int32_t in_h, int32_t in_w, uint8_t* canvas, int32_t can_h, int32_t can_w,
                             int32_t out_h, int32_t out_w, int32_t off_y, int32_t off_x, int32_t fill, uint8_t* tmp,
                             const int32_t* bounds_h, const int32_t* kk_h, int32_t ksize_h, const int32_t* bounds_v,
                             const int32_t* kk_v, int32_t ksize_v, void* stream);

/* Payload of the multi-GPU detection gather (replaces the pickle -> uint8 tensor step of mmdet's collect_results_gpu,
 * yolox-ufp/mmdet/apis/test.py:161-175).  det [batch, det_rows, 7] fp32, count [batch] int32.  total_rows == 0: padded
 * layout out[batch][1 + rows][7] (count in element [b][0][0]); total_rows > 0: packed layout
 * out[(batch + 6) / 7 header rows with the counts | rows of all images back to back, at most total_rows][7]. */
int glsdet_pack_detections(const float* det, const int32_t* count, int32_t batch, int32_t det_rows, int32_t rows,
                           int32_t total_rows, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GLSDET_B200_H_ */
