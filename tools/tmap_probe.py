"""Does cuTensorMapEncodeTiled accept overlapping strides (stride of dim1 < extent of dim0)?"""
import torch
from cuda.bindings import driver as drv
torch.cuda.init()
x = torch.zeros(4 * 230 * 230 * 8, dtype=torch.float16, device="cuda")
def enc(dims, strides, box, estr, swz=drv.CUtensorMapSwizzle.CU_TENSOR_MAP_SWIZZLE_128B):
    r = drv.cuTensorMapEncodeTiled(drv.CUtensorMapDataType.CU_TENSOR_MAP_DATA_TYPE_FLOAT16, len(dims), x.data_ptr(),
                                   [drv.cuuint64_t(d) for d in dims], [drv.cuuint64_t(s) for s in strides],
                                   [drv.cuuint32_t(b) for b in box], [drv.cuuint32_t(e) for e in estr],
                                   drv.CUtensorMapInterleave.CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                   drv.CUtensorMapL2promotion.CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   drv.CUtensorMapFloatOOBfill.CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
    return r[0]
W = 230
print("plain     ", enc([8, W, 230, 4], [16, W * 16, 230 * W * 16], [8, 16, 8, 1], [1, 1, 1, 1], drv.CUtensorMapSwizzle.CU_TENSOR_MAP_SWIZZLE_NONE))
print("overlap32 ", enc([64, 112, 230, 4], [32, W * 16, 230 * W * 16], [64, 16, 8, 1], [1, 1, 1, 1]))
print("overlap16 ", enc([64, 223, 230, 4], [16, W * 16, 230 * W * 16], [64, 16, 8, 1], [1, 1, 1, 1]))
print("overlap32 hstride2", enc([64, 112, 112, 4], [32, 2 * W * 16, 230 * W * 16], [64, 16, 8, 1], [1, 1, 1, 1]))
