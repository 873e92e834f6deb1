"""Probe: does cuTensorMapEncodeTiled accept an overlapping view (stride of dim 1 smaller than the extent of dim 0)?"""
import torch
from cuda.bindings import driver as drv

torch.cuda.init()
x = torch.zeros(1 << 20, dtype=torch.bfloat16, device="cuda")
e = 2
W, H, B = 64, 64, 2
pitch = (W + 2) * 16
dims = [drv.cuuint64_t(v) for v in (64, W, 1, H, B)]
strides = [drv.cuuint64_t(v) for v in (16 * e, pitch * e, pitch * e, H * pitch * e)]
box = [drv.cuuint32_t(v) for v in (64, 16, 1, 10, 1)]
estr = [drv.cuuint32_t(1)] * 5
res = drv.cuTensorMapEncodeTiled(drv.CUtensorMapDataType.CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, x.data_ptr(), dims, strides, box, estr,
                                 drv.CUtensorMapInterleave.CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 drv.CUtensorMapSwizzle.CU_TENSOR_MAP_SWIZZLE_128B,
                                 drv.CUtensorMapL2promotion.CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 drv.CUtensorMapFloatOOBfill.CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
print("overlapping view:", res[0])
