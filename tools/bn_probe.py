"""Which fp32 formula reproduces torch's eval-mode BatchNorm2d on this GPU bit for bit?"""
import torch
torch.manual_seed(0)
dev = "cuda"
for fmt in (torch.contiguous_format, torch.channels_last):
    for cudnn in (True, False):
        torch.backends.cudnn.enabled = cudnn
        C = 64
        bn = torch.nn.BatchNorm2d(C).to(dev).eval()
        with torch.no_grad():
            bn.weight.normal_(1, 0.3); bn.bias.normal_(0, 0.3)
            bn.running_mean.normal_(0, 1.0); bn.running_var.uniform_(0.3, 3.0)
        x = (torch.randn(8, C, 28, 28, device=dev) * 5).contiguous(memory_format=fmt)
        with torch.no_grad():
            y = bn(x)
        g, b, m, v, eps = bn.weight.view(1, C, 1, 1), bn.bias.view(1, C, 1, 1), bn.running_mean.view(1, C, 1, 1), bn.running_var.view(1, C, 1, 1), bn.eps
        invstd = 1.0 / torch.sqrt(v + eps)
        rs = torch.rsqrt(v + eps)
        cands = {}
        a1 = g * invstd; cands["fma(x, g*invstd, b - m*(g*invstd))"] = torch.addcmul(b - m * a1, x, a1)
        cands["x*a + (b - m*a) separate"] = x * a1 + (b - m * a1)
        a2 = g * rs; cands["fma(x, g*rsqrt, b - m*g*rsqrt)"] = torch.addcmul(b - m * a2, x, a2)
        cands["((x-m)*invstd)*g + b"] = ((x - m) * invstd) * g + b
        cands["fma((x-m)*invstd, g, b)"] = torch.addcmul(b, (x - m) * invstd, g)
        cands["fma(g*(x-m), invstd, b)"] = torch.addcmul(b, g * (x - m), invstd)
        cands["(g*(x-m))*invstd + b"] = (g * (x - m)) * invstd + b
        cands["fma((x-m), g*invstd, b)"] = torch.addcmul(b, (x - m), a1)
        cands["(x-m)*(g*invstd) + b"] = (x - m) * a1 + b
        cands["fma((x-m)*rs, g, b)"] = torch.addcmul(b, (x - m) * rs, g)
        # double precision reference rounded once
        yd = ((x.double() - m.double()) / torch.sqrt(v.double() + eps) * g.double() + b.double()).float()
        cands["fp64 formula rounded"] = yd
        print(f"--- format={fmt} cudnn={cudnn}")
        for k, c in cands.items():
            nz = int((c != y).sum())
            print(f"   {k:45s} mismatches {nz:8d} / {y.numel()}  maxdiff {float((c-y).abs().max()):.3e}")
