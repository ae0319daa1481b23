"""GPU parity of the tcgen05 convolution on term codes: the fp32 accumulators must equal the
exact integer convolution of the same codes (float64 conv of integer-valued tensors is exact)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [
    # N, H, W, C, Cout, k, stride, pad          (ResNet-18 layer shapes)
    (2, 56, 56, 64, 64, 3, 1, 1),
    (3, 28, 28, 128, 128, 3, 1, 1),
    (2, 56, 56, 64, 128, 3, 2, 1),
    (2, 56, 56, 64, 128, 1, 2, 0),
    (3, 14, 14, 256, 256, 3, 1, 1),
    (2, 28, 28, 128, 256, 3, 2, 1),
    (5, 7, 7, 512, 512, 3, 1, 1),
    (3, 14, 14, 256, 512, 1, 2, 0),
    (1, 9, 13, 72, 68, 3, 1, 1),            # ragged: C % 64 != 0, Cout % 64 != 0, odd image
    (2, 5, 5, 8, 4, 5, 1, 2),
]


@pytest.mark.parametrize("case", CASES)
def test_conv_codes_exact_accumulators(case):
    from term_quantization_b200 import conv_codes
    N, H, W, C, Cout, k, stride, pad = case
    g = torch.Generator(device="cuda").manual_seed(sum(case))
    act = torch.randint(0, 513, (N, H, W, C), device="cuda", generator=g)
    act = act * (torch.rand(N, H, W, C, device="cuda", generator=g) < 0.6)       # post-ReLU sparsity
    wgt = torch.randint(-256, 257, (k * k, Cout, C), device="cuda", generator=g)
    bias = torch.randn(Cout, device="cuda", generator=g)
    out = conv_codes.conv2d_codes(act.half(), wgt.half(), None, (k, k), stride, pad, 1.0)
    w_oihw = wgt.view(k, k, Cout, C).permute(2, 3, 0, 1).double()
    want = F.conv2d(act.permute(0, 3, 1, 2).double(), w_oihw, None, stride, pad).permute(0, 2, 3, 1)
    assert float(want.abs().max()) < 2 ** 24
    assert out.shape == want.shape
    assert torch.equal(out.double(), want), float((out.double() - want).abs().max())
    # scale and bias in the epilogue
    out2 = conv_codes.conv2d_codes(act.half(), wgt.half(), bias, (k, k), stride, pad, 0.00123)
    want2 = (want.float() * np.float32(0.00123) + bias).float()
    assert torch.equal(out2, want2)
