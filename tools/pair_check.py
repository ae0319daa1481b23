"""Quick exactness check of the CTA-pair (cta_group::2) conv path against an fp64 torch conv on integer codes."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from term_quantization_b200 import conv_codes  # noqa: E402

torch.manual_seed(0)
bad = 0
for (N, H, C, Cout, amp) in ((32, 28, 128, 128, 40), (33, 28, 128, 128, 40), (64, 14, 256, 256, 40), (25, 28, 64, 192, 60), (256, 28, 128, 128, 40),
                            (256, 7, 512, 512, 20), (5, 7, 128, 128, 40), (2, 7, 64, 128, 40), (301, 7, 256, 256, 30), (37, 4, 128, 256, 40),
                            (64, 5, 128, 128, 40),
                            (16, 56, 64, 64, 60), (5, 56, 64, 64, 60), (256, 56, 64, 64, 60), (9, 40, 32, 48, 60), (7, 60, 64, 16, 60)):
    act = (torch.randint(0, 513, (N, H, H, C), device="cuda") * (torch.rand(N, H, H, C, device="cuda") < 0.5)).half()
    wgt = (torch.randn(9, Cout, C, device="cuda") * amp).round().clamp(-256, 256).half().contiguous()
    plan = conv_codes.plan_weight(wgt, 512)
    out = conv_codes.conv2d_codes(act, wgt, None, (3, 3), 1, 1, 1.0, plan=plan)
    torch.cuda.synchronize()
    ref = F.conv2d(act.double().permute(0, 3, 1, 2), wgt.double().view(3, 3, Cout, C).permute(2, 3, 0, 1), None, 1, 1).permute(0, 2, 3, 1)
    diff = (out.double() - ref).abs().max().item()
    print(f"N={N} H={H} C={C} Cout={Cout} engine={plan.engine} x{plan.groups}: max |diff| = {diff}", flush=True)
    bad += diff != 0
sys.exit(1 if bad else 0)
