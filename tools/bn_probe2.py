import itertools
import torch
torch.manual_seed(0)
dev = "cuda"
C = 64
bn = torch.nn.BatchNorm2d(C).to(dev).eval()
with torch.no_grad():
    bn.weight.normal_(1, 0.3); bn.bias.normal_(0, 0.3)
    bn.running_mean.normal_(0, 1.0); bn.running_var.uniform_(0.3, 3.0)
    x = (torch.randn(8, C, 28, 28, device=dev) * 5).contiguous(memory_format=torch.channels_last)
    y = bn(x)
    g, b, m, v, eps = [t.view(1, C, 1, 1) for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var)] + [bn.eps]
    ve = v + eps
    inv_variants = {"rsqrt": torch.rsqrt(ve), "1/sqrt": 1.0 / torch.sqrt(ve), "rsqrt64": torch.rsqrt(ve.double()).float(),
                    "sqrt_then_rcp64": (1.0 / torch.sqrt(ve.double())).float()}
    def fma(a, bb, c):
        return torch.addcmul(c, a, bb)
    for iname, inv in inv_variants.items():
        a = g * inv
        a64 = (g.double() * torch.rsqrt(ve.double())).float()
        res = {
            "fma(x,a,fma(-m,a,b))": fma(x, a, fma(-m, a, b)),
            "fma(x,a,b-m*a)": fma(x, a, b - m * a),
            "fma(x-m,a,b)": fma(x - m, a, b),
            "fma(x,a64,b-m*a64 in 64)": fma(x, a64, (b.double() - m.double() * g.double() * torch.rsqrt(ve.double())).float()),
            "fma((x-m)*inv,g,b)": fma((x - m) * inv, g, b),
            "fma(fma(x,inv,-m*inv),g,b)": fma(fma(x, inv, -m * inv), g, b),
            "fma(x*inv - m*inv..)": fma(x * inv - m * inv, g, b),
        }
        for k, c in res.items():
            print(f"{iname:16s} {k:32s} mismatches {int((c != y).sum()):8d} maxdiff {float((c - y).abs().max()):.3e}")
