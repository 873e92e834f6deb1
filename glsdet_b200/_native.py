"""ctypes binding of libglsdet_b200.so (the C ABI declared in include/glsdet_b200.h).

There is no fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libglsdet_b200.so"
if os.environ.get("GLSDET_LIB"):   # A/B measurements of two builds on one box (tools/): another in-tree build of the same ABI
    _LIB_PATH = Path(os.environ["GLSDET_LIB"]).resolve()

ACT_NONE, ACT_SILU, ACT_RELU, ACT_LRELU, ACT_SIGMOID, ACT_YOLOX_BOX, ACT_MMDET_BOX = range(7)
OUT_NHWC_BF16, OUT_NHWC_F32, OUT_NCHW_F32 = range(3)
NMS_COORD_TRICK, NMS_PER_CLASS, NMS_AUTO_CUDA, NMS_AUTO_CPU, NMS_MMCV = range(5)
PRED_ROWS, PRED_PLANES, PRED_CLS_LOGITS = 0, 1, 2
DECODE_SIGMOID_OBJ, DECODE_SIGMOID_CLS, DECODE_NORMALISE, DECODE_XYXY = 1, 2, 4, 8
SE_SLABS = 32
DT_BF16, DT_F16, DT_F32 = 0, 1, 2   # GLSDET_DT_*: storage type of an activation tensor (F32: accuracy mode, glsdet_dwconv)


def dt_code(dtype) -> int:
    """GLSDET_DT_* of a torch dtype (the two 16-bit storage types of the tensor-core path)."""
    import torch

    if dtype == torch.bfloat16:
        return DT_BF16
    if dtype == torch.float16:
        return DT_F16
    raise TypeError(f"not a 16-bit storage type of the native path: {dtype}")


def storage_dtype(stride: int, mode: str = None, role: str = "head"):
    """Storage policy of the tensor-core path.  "mixed" (default): the tensors of the HEAD at the
    stride-4 level (head.csp, the level-0 towers: two thirds of the FLOPs behind a chain of fewer than ten layers) are bf16,
    everything else - the backbone, the plan inputs, the neck, FFA and the coarser head levels, which sit behind 25-50
    layers - is fp16.  bf16's 8-bit mantissa alone costs 2-4 % relative error against the fp32 reference along those
    chains (BASELINE.json bounds it by 2e-2), fp16's 11 bits 0.1-0.5 %; tcgen05.mma kind::f16 takes either type at the
    same rate.  fp16 stores saturate at +-65504 and flush below 6e-8: activations of a trained network are O(1), but an
    untrained one whose activations collapse (the reference's own N(0, 0.02) init at width 1.0) needs "bf16".
    GLSDET_STORAGE=bf16 | f16 forces one type everywhere.  `role`: "head" | "input" | "backbone"."""
    import torch

    mode = mode or os.environ.get("GLSDET_STORAGE", "mixed")
    if mode == "bf16":
        return torch.bfloat16
    if mode == "f16":
        return torch.float16
    assert mode == "mixed", mode
    return torch.bfloat16 if (stride <= 4 and role == "head") else torch.float16

ACT_BY_NAME = {"none": ACT_NONE, "silu": ACT_SILU, "relu": ACT_RELU, "lrelu": ACT_LRELU}


class NativeError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    """Mirror of struct glsdet_conv_desc."""

    _fields_ = [
        ("src0", C.c_void_p), ("src0_c", C.c_int32), ("src0_ld", C.c_int32),
        ("src1", C.c_void_p), ("src1_c", C.c_int32), ("src1_ld", C.c_int32),
        ("batch", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
        ("ksize", C.c_int32), ("stride", C.c_int32),
        ("weight", C.c_void_p), ("out_channels", C.c_int32),
        ("bias", C.c_void_p), ("act", C.c_int32),
        ("pre_res", C.c_void_p), ("pre_shift", C.c_int32), ("pre_ld", C.c_int32),
        ("post_res", C.c_void_p), ("post_shift", C.c_int32), ("post_ld", C.c_int32),
        ("out", C.c_void_p), ("out_mode", C.c_int32), ("out_ld", C.c_int32), ("out_coff", C.c_int32),
        ("out_batch_stride", C.c_int64),
        ("dec_stride", C.c_float), ("dec_in_w", C.c_float), ("dec_in_h", C.c_float),
        ("pred_weight", C.c_void_p), ("pred_bias", C.c_void_p), ("pred_channels", C.c_int32), ("pred_act", C.c_int32),
        ("weight_batch_stride", C.c_int64), ("weight_ld", C.c_int32), ("src_shared", C.c_int32),
        ("src_shared_div", C.c_int32), ("patch_mode", C.c_int32),
        ("ksize_w", C.c_int32), ("src0_row_pitch", C.c_int64), ("src0_img_pitch", C.c_int64),
        ("out_plane_stride", C.c_int64),
        ("src_dtype", C.c_int32), ("out_dtype", C.c_int32), ("post_dtype", C.c_int32),
    ]


class Rect(C.Structure):
    """Mirror of struct glsdet_rect."""

    _fields_ = [(n, C.c_int32) for n in ("sb", "sy", "sx", "db", "dy", "dx", "h", "w")]


class ConvF32Desc(C.Structure):
    """Mirror of struct glsdet_conv_f32_desc."""

    _fields_ = [
        ("src0", C.c_void_p), ("src0_c", C.c_int32), ("src0_ld", C.c_int32),
        ("src1", C.c_void_p), ("src1_c", C.c_int32), ("src1_ld", C.c_int32),
        ("batch", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("ksize", C.c_int32), ("stride", C.c_int32),
        ("weight", C.c_void_p), ("out_channels", C.c_int32),
        ("bias", C.c_void_p), ("act", C.c_int32),
        ("pre_res", C.c_void_p), ("pre_shift", C.c_int32), ("pre_ld", C.c_int32),
        ("post_res", C.c_void_p), ("post_shift", C.c_int32), ("post_ld", C.c_int32),
        ("out", C.c_void_p), ("out_mode", C.c_int32), ("out_ld", C.c_int32), ("out_coff", C.c_int32),
        ("out_batch_stride", C.c_int64),
        ("dec_stride", C.c_float), ("dec_in_w", C.c_float), ("dec_in_h", C.c_float),
    ]


_lib = None

# name -> (restype, argtypes); must list every symbol include/glsdet_b200.h declares
SIGNATURES = {
    "glsdet_abi_version": (C.c_int, []),
    "glsdet_last_error": (C.c_char_p, []),
    "glsdet_launch_count": (C.c_int64, []),
    "glsdet_conv_weight_shape": (C.c_int, [C.POINTER(ConvDesc), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                           C.POINTER(C.c_int32)]),
    "glsdet_conv_create": (C.c_int, [C.POINTER(ConvDesc), C.POINTER(C.c_void_p)]),
    "glsdet_conv_launch": (C.c_int, [C.c_void_p, C.c_void_p]),
    "glsdet_conv_destroy": (None, [C.c_void_p]),
    "glsdet_conv_read_trace": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "glsdet_nchw_f32_to_nhwc_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                               C.c_int32, C.c_int32, C.c_void_p]),
    "glsdet_nhwc_bf16_to_nchw_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                               C.c_int32, C.c_int32, C.c_void_p]),
    "glsdet_nchw_f32_to_nhwc_16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                             C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "glsdet_nhwc_16_to_nchw_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                             C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "glsdet_patch_transpose_16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_int32, C.c_int32, C.c_float, C.c_void_p]),
    "glsdet_nhwc_transpose_16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                           C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p]),
    "glsdet_gather_bias_16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "glsdet_spp_maxpool_16": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "glsdet_se_gate_16": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                    C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "glsdet_scale_pixel_shuffle_16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                                C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "glsdet_patch_transpose": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_int32, C.c_void_p]),
    "glsdet_gather_bias": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_int64, C.c_int32, C.c_void_p]),
    "glsdet_conv_f32": (C.c_int, [C.POINTER(ConvF32Desc), C.c_void_p]),
    "glsdet_bgemm_f32": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p,
                                   C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_void_p]),
    "glsdet_nchw_nhwc_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_int32, C.c_void_p]),
    "glsdet_se_partial_f32": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "glsdet_se_fc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                               C.c_int32, C.c_void_p]),
    "glsdet_scale_pixel_shuffle_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                                 C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "glsdet_rect_copy": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                   C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                   C.c_void_p]),
    "glsdet_nhwc_transpose": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_int32, C.c_void_p]),
    "glsdet_group_norm_relu": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                         C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "glsdet_group_norm_scratch_floats": (C.c_int64, [C.c_int32, C.c_int32]),
    "glsdet_proxy_scores": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_float, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p]),
    "glsdet_proxy_aggregate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_float, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p]),
    "glsdet_gfl_decode": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                    C.c_float, C.c_float, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "glsdet_gfl_select": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_float, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "glsdet_gfl_select_scratch_ints": (C.c_int64, [C.c_int32]),
    "glsdet_upsample2x": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "glsdet_dwconv": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                C.c_void_p]),
    "glsdet_focus_nchw_f32_to_nhwc_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                                     C.c_void_p]),
    "glsdet_spp_maxpool": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "glsdet_focus_nchw_f32_to_nhwc_16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                                   C.c_int32, C.c_void_p]),
    "glsdet_focus_u8_to_nhwc_16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                             C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int32, C.c_void_p]),
    "glsdet_focus_u8_to_nhwc_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                               C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p]),
    "glsdet_focus_nchw_f32_to_nhwc_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "glsdet_spp_maxpool_f32": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "glsdet_se_gate": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                 C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "glsdet_scale_pixel_shuffle": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                             C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "glsdet_decode_outputs": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32,
                                        C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "glsdet_decode_outputs_mode": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32,
                                             C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "glsdet_nms_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "glsdet_nms_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                    C.POINTER(C.c_void_p)]),
    "glsdet_nms_launch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int32, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "glsdet_nms_launch_scaled": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int32,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "glsdet_nms_launch_layout": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_float, C.c_float, C.c_int32,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "glsdet_nms_destroy": (None, [C.c_void_p]),
    "glsdet_decode_mmdet": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                      C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                      C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32,
                                      C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "glsdet_batched_nms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_int32, C.c_void_p,
                                     C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "glsdet_batched_nms_ids": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_float, C.c_int32,
                                         C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "glsdet_batched_nms_workspace_bytes": (C.c_int64, [C.c_int32]),
    "glsdet_batched_nms_batch_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "glsdet_batched_nms_ids_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_int32,
                                               C.c_int32, C.c_float, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                               C.c_void_p]),
    "glsdet_pil_bicubic_ksize": (C.c_int, [C.c_int32, C.c_int32]),
    "glsdet_pil_bicubic_table": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "glsdet_resize_bicubic_u8": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                           C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "glsdet_pack_detections": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "glsdet_ufp_pack": (C.c_int, [C.c_void_p, C.c_int32, C.c_float, C.c_int32, C.c_int32, C.c_void_p,
                                  C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "glsdet_ufp_mosaic": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                    C.c_int32, C.c_void_p]),
    "glsdet_ufp_merge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_float, C.c_void_p,
                                   C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}


def lib_path() -> Path:
    return _LIB_PATH


def load():
    """Load the shared library (once) and declare all prototypes. Raises NativeError when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise NativeError(
            f"{_LIB_PATH} is missing: build it with `python -m glsdet_b200._build` (or __graft_entry__.build()). "
            "glsdet_b200 has no CPU or PyTorch fallback path.")
    lib = C.CDLL(os.fspath(_LIB_PATH))
    missing = []
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    if missing and not os.environ.get("GLSDET_ALLOW_PARTIAL_LIB"):
        raise NativeError(f"libglsdet_b200.so lacks symbols declared in include/glsdet_b200.h: {missing}")
    if lib.glsdet_abi_version() != 2:
        raise NativeError("libglsdet_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().glsdet_last_error().decode("utf-8", "replace")
        raise NativeError(f"{what or 'glsdet call'} failed (rc={rc}): {msg}")


def stream_ptr(stream=None) -> C.c_void_p:
    import torch

    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def ptr(t) -> C.c_void_p:
    """Raw device pointer of a torch tensor (None -> NULL)."""
    return C.c_void_p(0 if t is None else t.data_ptr())
