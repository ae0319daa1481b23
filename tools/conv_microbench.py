"""Per-layer throughput of the tcgen05 code-domain conv at the ResNet-18 / VGG-16 shapes of
BASELINE.json (batch 256 / 128).  CUDA events, 10 timed iterations after 3 warm-ups."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from term_quantization_b200 import conv_codes  # noqa: E402

RESNET18 = [  # (H, W, C, Cout, k, stride, pad, count)
    (56, 56, 64, 64, 3, 1, 1, 4), (56, 56, 64, 128, 3, 2, 1, 1), (56, 56, 64, 128, 1, 2, 0, 1),
    (28, 28, 128, 128, 3, 1, 1, 3), (28, 28, 128, 256, 3, 2, 1, 1), (28, 28, 128, 256, 1, 2, 0, 1),
    (14, 14, 256, 256, 3, 1, 1, 3), (14, 14, 256, 512, 3, 2, 1, 1), (14, 14, 256, 512, 1, 2, 0, 1),
    (7, 7, 512, 512, 3, 1, 1, 3)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--only", type=int, default=-1)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    torch.manual_seed(0)
    total_ms = total_flop = 0.0
    for i, (H, W, C, Cout, k, s, p, cnt) in enumerate(RESNET18):
        if args.only >= 0 and i != args.only:
            continue
        act = (torch.randint(0, 513, (args.batch, H, W, C), device="cuda") *
               (torch.rand(args.batch, H, W, C, device="cuda") < 0.5)).half()
        # codes of the magnitude real term-revealed weights have (He-initialised, 9-bit: mean |code| ~ 40), so that the
        # static exactness proof picks the same engine / K chunks as in the networks
        wgt = (torch.randn(k * k, Cout, C, device="cuda") * 55).round().clamp(-256, 256).half().contiguous()
        plan = conv_codes.plan_weight(wgt, 512)
        Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        out = torch.empty(args.batch, Ho, Wo, Cout, device="cuda")
        for _ in range(3):
            conv_codes.conv2d_codes(act, wgt, None, (k, k), s, p, 1.0, out=out, plan=plan)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            conv_codes.conv2d_codes(act, wgt, None, (k, k), s, p, 1.0, out=out, plan=plan)
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        flop = 2.0 * args.batch * Ho * Wo * Cout * C * k * k
        mb = (act.numel() * 2 + wgt.numel() * 2 + out.numel() * 4) / 1e6
        print(json.dumps({"layer": i, "shape": [H, W, C, Cout, k, s], "ms": round(ms, 4), "TFLOPs": round(flop / ms / 1e9, 1),
                          "min_GBs": round(mb / ms, 1), "count": cnt, "engine": f"{plan.engine} x{plan.groups}"}))
        total_ms += ms * cnt
        total_flop += flop * cnt
    print(json.dumps({"resnet18_wrapped_convs_ms": total_ms, "TFLOPs": total_flop / max(total_ms, 1e-9) / 1e9}))
    if args.only < 0 or args.only == 99:
        import torch.nn.functional as F
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cudnn.benchmark = True
        x = torch.randn(args.batch, 3, 224, 224, device="cuda").contiguous(memory_format=torch.channels_last)
        w = torch.randn(64, 3, 7, 7, device="cuda") * 0.1
        w2 = conv_codes.pack_stem_weight(w)
        xn = x.permute(0, 2, 3, 1)
        _, scratch = conv_codes.stem_conv7x7s2(xn, w2)
        xb = xn.bfloat16().contiguous()
        bn = (torch.rand(64, device="cuda") + 0.5, torch.randn(64, device="cuda"))
        nq = (0.01, 9, 3)
        w2s, bns = conv_codes.stem_pool_operands(w, bn)
        for name, fn in (("stem tcgen05 hi/lo", lambda: conv_codes.stem_conv7x7s2(xn, w2, scratch)),
                         ("stem tcgen05 bf16 images", lambda: conv_codes.stem_conv7x7s2(xb, w2, scratch)),
                         ("stem+bn+relu+pool+encode fused, fp32 images", lambda: conv_codes.stem_conv_pool(xn, w2s, bns, True, nq, scratch)),
                         ("stem+bn+relu+pool+encode fused, bf16 images", lambda: conv_codes.stem_conv_pool(xb, w2s, bns, True, nq, scratch)),
                         ("stem cuDNN fp32", lambda: F.conv2d(x, w, None, 2, 3))):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                fn()
            e1.record()
            e1.synchronize()
            print(json.dumps({"case": name, "ms": e0.elapsed_time(e1) / args.iters}))


if __name__ == "__main__":
    main()
