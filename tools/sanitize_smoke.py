"""One small launch of every kernel added in round 2 on ragged shapes (written for `compute-sanitizer --tool memcheck`,
which is closed on this GPU pool; as a plain run it still exercises every new launch path with synchronous error checks)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from term_quantization_b200 import conv_codes, inference, tr_cuda  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
# depthwise: both strides, 32- and 64-channel tiles, ragged maps, fp32 + codes out
for (N, H, W, C, stride) in ((1, 9, 11, 32, 1), (2, 7, 7, 64, 1), (1, 13, 10, 96, 2), (1, 17, 19, 128, 2), (1, 5, 3, 8, 1)):
    act = torch.randint(0, 513, (N, H, W, C), device="cuda", generator=g).half()
    w = torch.randint(-30000, 30000, (9, C), device="cuda", generator=g, dtype=torch.int32)
    a = torch.rand(C, device="cuda", generator=g) + 0.5
    b = torch.randn(C, device="cuda", generator=g)
    conv_codes.depthwise3x3_codes(act, w, stride, 1e-7, bn=(a, b), relu="relu6", want_f32=True, next_quant=(0.01, 9, 3), act_unsigned=True)
    conv_codes.depthwise3x3_codes(act, w, stride, 1e-7, relu=False, next_quant=(0.01, 9, 3))
    conv_codes.bn_act_encode(torch.randn(N, H, W, C, device="cuda", generator=g), (a, b), relu=True, want_f32=True, next_quant=(0.01, 9, 3))
# planes, static bound, kind::i8 conv (1 and 2 planes), K-chunk kind::f16 conv, single accumulator stage
for (C, Cout, k, amax, wmax) in ((48, 40, 3, 512, 256), (64, 64, 1, 100, 100), (256, 128, 3, 512, 256)):
    act = torch.randint(0, amax + 1, (2, 9, 10, C), device="cuda", generator=g).half()
    wgt = torch.randint(-wmax, wmax + 1, (k * k, Cout, C), device="cuda", generator=g).half().contiguous()
    plan = conv_codes.plan_weight(wgt, amax, engine="i8")
    conv_codes.conv2d_codes(act, wgt, None, (k, k), 1, k // 2, 1.0, plan=plan)
    a = torch.rand(Cout, device="cuda", generator=g) + 0.5
    b = torch.randn(Cout, device="cuda", generator=g)
    conv_codes.conv2d_codes_fused(act, wgt, (k, k), 1, k // 2, 1e-6, bn=(a, b), relu="relu6", next_quant=(0.01, 9, 3), plan=plan)
for (C, Cout, amp, groups) in ((256, 256, 60, 2), (512, 512, 90, 4)):
    act = torch.randint(0, 513, (3, 7, 7, C), device="cuda", generator=g).half()
    wgt = torch.randint(-amp, amp + 1, (9, Cout, C), device="cuda", generator=g).half().contiguous()
    plan = conv_codes.plan_weight(wgt, 512, engine="f16")
    assert plan.groups == groups, plan
    res = torch.randn(3, 7, 7, Cout, device="cuda", generator=g)
    conv_codes.conv2d_codes_fused(act, wgt, (3, 3), 1, 1, 1e-6, residual=res, relu=True, next_quant=(0.01, 9, 3), plan=plan)
# linear path with K and out padding, uint8 normalisation, BINARY / BOOTH encodings
x = torch.randint(-200, 201, (37, 650), device="cuda", generator=g).half()
w = torch.randint(-128, 129, (30, 650), device="cuda", generator=g).float()
packed, _ = conv_codes.pack_linear_weight(w, 1.0)
conv_codes.linear_codes(x, packed, 1.0, out_features=30, act_max=256)
inference.normalize_u8(torch.randint(0, 256, (2, 6, 10, 3), device="cuda", dtype=torch.uint8, generator=g))
for enc in ("binary", "booth"):
    tr_cuda.tr(torch.randn(4, 64, 3, 3, device="cuda", generator=g), 0.01, 8, 8, 12, encoding=enc)
torch.cuda.synchronize()
print("sanitize smoke ok")
