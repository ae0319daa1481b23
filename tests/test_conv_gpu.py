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
    # resident-weight + halo mode (Cout <= 64, one channel block, stride 1): ragged maps, narrow channels
    (1, 9, 13, 64, 64, 3, 1, 1),
    (2, 20, 33, 32, 40, 3, 1, 1),
    (2, 17, 17, 16, 8, 3, 1, 1),
    (2, 30, 31, 64, 64, 2, 1, 0),
    (1, 130, 70, 48, 64, 3, 1, 1),
    # streamed-weight halo mode (Cout > 64 or several channel blocks, stride 1)
    (2, 28, 28, 128, 128, 3, 1, 1),
    (1, 33, 21, 72, 136, 3, 1, 1),
    (2, 14, 14, 256, 256, 3, 1, 1),
    (1, 40, 40, 64, 128, 5, 1, 2),
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


FUSED = [(2, 56, 56, 64, 64, 3, 1, 1), (3, 28, 28, 128, 128, 3, 1, 1), (2, 56, 56, 64, 128, 1, 2, 0),
         (5, 7, 7, 512, 512, 3, 1, 1), (1, 9, 13, 72, 72, 3, 1, 1), (3, 14, 14, 256, 256, 3, 1, 1),
         (2, 20, 33, 32, 40, 3, 1, 1), (1, 57, 29, 64, 64, 3, 1, 1), (2, 30, 27, 128, 192, 3, 1, 1)]


def _random_fused_cases(n, seed=31):
    rng = np.random.default_rng(seed)
    cases = []
    while len(cases) < n:
        k = int(rng.choice([1, 3, 3, 3, 5]))
        stride = int(rng.choice([1, 1, 2]))
        pad = int(rng.integers(0, k // 2 + 1))
        H, W = int(rng.integers(max(k, 4), 61)), int(rng.integers(max(k, 4), 61))
        C = 8 * int(rng.integers(1, 9))                  # accumulators stay below 2^24 with 9-bit codes
        Cout = 8 * int(rng.integers(1, 33))              # code output needs Cout % 8 == 0
        N = int(rng.integers(1, 4))
        if 513 * 256 * C * k * k >= 2 ** 24:
            continue
        cases.append((N, H, W, C, Cout, k, stride, pad))
    return cases


@pytest.mark.parametrize("case", FUSED + _random_fused_cases(16))
def test_conv_fused_epilogue(case):
    """scale -> fma(BN) -> + residual -> ReLU -> fp32 out + fp16 term codes of the result, each
    step against plain torch / the CPU oracle."""
    from oracle import tq_oracle as O
    from term_quantization_b200 import conv_codes
    N, H, W, C, Cout, k, stride, pad = case
    g = torch.Generator(device="cuda").manual_seed(7 + sum(case))
    act = (torch.randint(0, 513, (N, H, W, C), device="cuda", generator=g) *
           (torch.rand(N, H, W, C, device="cuda", generator=g) < 0.5)).half()
    wgt = torch.randint(-256, 257, (k * k, Cout, C), device="cuda", generator=g).half()
    w_oihw = wgt.view(k, k, Cout, C).permute(2, 3, 0, 1).double()
    acc = F.conv2d(act.permute(0, 3, 1, 2).double(), w_oihw, None, stride, pad).permute(0, 2, 3, 1).float()
    scale = np.float32(3.1e-6)
    a = torch.rand(Cout, device="cuda", generator=g) + 0.5
    b = torch.randn(Cout, device="cuda", generator=g)
    res = torch.randn(acc.shape, device="cuda", generator=g)
    t0 = acc * scale
    t1 = (t0.double() * a.double() + b.double()).float()          # fma(t, a, b), one rounding
    t2 = t1 + res
    t3 = torch.relu(t2)
    nq = (float(t3.max()) / 512, 9, 3)

    out, codes = conv_codes.conv2d_codes_fused(act, wgt, (k, k), stride, pad, scale)
    assert codes is None and torch.equal(out, t0)
    out, _ = conv_codes.conv2d_codes_fused(act, wgt, (k, k), stride, pad, scale, bn=(a, b))
    assert torch.equal(out, t1)
    out, _ = conv_codes.conv2d_codes_fused(act, wgt, (k, k), stride, pad, scale, bn=(a, b), residual=res)
    assert torch.equal(out, t2)
    out, codes = conv_codes.conv2d_codes_fused(act, wgt, (k, k), stride, pad, scale, bn=(a, b), residual=res,
                                               relu=True, next_quant=nq)
    assert torch.equal(out, t3)
    _, want = O.tr(t3.cpu().numpy().reshape(1, -1, 1, 1), nq[0], nq[1], 1, nq[2], return_codes=True)
    assert np.array_equal(codes.cpu().numpy().astype(np.int32).reshape(-1), want.reshape(-1))
    # codes only (no fp32 tile), signed values (no ReLU)
    none, codes = conv_codes.conv2d_codes_fused(act, wgt, (k, k), stride, pad, scale, bn=(a, b), want_f32=False,
                                                next_quant=(float(t1.abs().max()) / 256, 8, 2))
    assert none is None
    _, want = O.tr(t1.cpu().numpy().reshape(1, -1, 1, 1), float(t1.abs().max()) / 256, 8, 1, 2, return_codes=True)
    assert np.array_equal(codes.cpu().numpy().astype(np.int32).reshape(-1), want.reshape(-1))


def test_bn_relu_maxpool_encode():
    from oracle import tq_oracle as O
    from term_quantization_b200 import conv_codes
    g = torch.Generator(device="cuda").manual_seed(11)
    for (N, H, W, C) in ((2, 112, 112, 64), (3, 17, 9, 8), (1, 7, 7, 20)):
        x = torch.randn(N, H, W, C, device="cuda", generator=g) * 2
        a = torch.randn(C, device="cuda", generator=g)            # negative slopes included
        b = torch.randn(C, device="cuda", generator=g)
        y = (x.double() * a.double() + b.double()).float()        # fma, one rounding
        want = F.max_pool2d(torch.relu(y).permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1).contiguous()
        nq = (float(want.max()) / 512, 9, 3)
        out, codes = conv_codes.bn_relu_maxpool_encode(x, (a, b), relu=True, next_quant=nq)
        assert torch.equal(out, want)
        _, wc = O.tr(want.cpu().numpy().reshape(1, -1, 1, 1), nq[0], nq[1], 1, nq[2], return_codes=True)
        assert np.array_equal(codes.cpu().numpy().astype(np.int32).reshape(-1), wc.reshape(-1))
        out2, none = conv_codes.bn_relu_maxpool_encode(x, (a, b), relu=False)
        assert none is None
        assert torch.equal(out2, F.max_pool2d(y.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1))


def test_stem_conv_tensor_cores_fp32_accuracy():
    """7x7/s2/p3 stem conv with hi/lo fp16 operand pairs on the tcgen05 kernel: as close to an fp64
    evaluation as cuDNN's fp32 conv is."""
    from term_quantization_b200 import conv_codes
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(5)
    for N, H, W in ((3, 224, 224), (2, 64, 96)):
        x = torch.randn(N, 3, H, W, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
        w = torch.randn(64, 3, 7, 7, device="cuda", generator=g) * 0.1
        want = F.conv2d(x.double(), w.double(), None, 2, 3).permute(0, 2, 3, 1)
        cudnn = F.conv2d(x, w, None, 2, 3).permute(0, 2, 3, 1)
        got, _ = conv_codes.stem_conv7x7s2(x.permute(0, 2, 3, 1), conv_codes.pack_stem_weight(w))
        scale = float(want.abs().max())
        e_mine = float((got.double() - want).abs().max()) / scale
        e_cudnn = float((cudnn.double() - want).abs().max()) / scale
        print(f"stem conv max rel err vs fp64: tcgen05 hi/lo {e_mine:.2e}, cuDNN fp32 {e_cudnn:.2e}")
        assert got.shape == want.shape
        # tensor-core fp32 accumulation truncates (48 accumulate steps): ~2e-6 of the output range,
        # against ~3e-7 for an fp32 FMA chain; TF32 would be ~1e-3
        assert e_mine < 5e-6
        # 16-bit images: the same values reach the MMAs through the hi plane alone
        for dt in (torch.bfloat16, torch.float16):
            x16 = x.permute(0, 2, 3, 1).to(dt).contiguous()
            want16 = F.conv2d(x16.permute(0, 3, 1, 2).double(), w.double(), None, 2, 3).permute(0, 2, 3, 1)
            got16, _ = conv_codes.stem_conv7x7s2(x16, conv_codes.pack_stem_weight(w))
            assert float((got16.double() - want16).abs().max()) / scale < 5e-6


def test_stem_conv_pool_fused_equals_two_pass():
    """conv7x7s2 -> BN -> ReLU -> maxpool -> encode in ONE tensor-core kernel (pooling in the epilogue out of
    shared memory) must equal the two-pass path (stem conv, then bn_relu_maxpool_encode) bit for bit."""
    from term_quantization_b200 import conv_codes
    g = torch.Generator(device="cuda").manual_seed(21)
    for (N, H, W, dt, relu) in ((3, 224, 224, torch.float32, True), (2, 64, 96, torch.bfloat16, True),
                                (2, 50, 38, torch.float32, False), (1, 226, 222, torch.float16, True)):
        x = torch.randn(N, H, W, 3, device="cuda", generator=g).to(dt).contiguous()
        w = torch.randn(64, 3, 7, 7, device="cuda", generator=g) * 0.1
        a = torch.randn(64, device="cuda", generator=g)
        b = torch.randn(64, device="cuda", generator=g)
        w2 = conv_codes.pack_stem_weight(w)
        y, _ = conv_codes.stem_conv7x7s2(x, w2)
        want, _ = conv_codes.bn_relu_maxpool_encode(y, (a, b), relu=relu)
        nq = (float(want.abs().max()) / 512, 9, 3)
        want, want_codes = conv_codes.bn_relu_maxpool_encode(y, (a, b), relu=relu, next_quant=nq)
        w2s, bns = conv_codes.stem_pool_operands(w, (a, b))      # sign of the slope folded into the weights
        got, codes, _ = conv_codes.stem_conv_pool(x, w2s, bns, relu=relu, next_quant=nq)
        assert got.shape == want.shape
        assert torch.equal(got, want), float((got - want).abs().max())
        assert torch.equal(codes, want_codes)
        got2, none, _ = conv_codes.stem_conv_pool(x, w2s, bns, relu=relu)
        assert none is None and torch.equal(got2, want)


def test_conv_codes_random_geometries():
    """Seeded sweep over conv geometries (image sizes, channel counts, filter sizes, strides, paddings, batch) to
    exercise every operand-movement mode of the kernel (streamed, resident weights, halo, streamed-weight halo) and
    their edge tiles: exact accumulators against an fp64 conv of the same integer tensors."""
    from term_quantization_b200 import conv_codes
    rng = np.random.default_rng(2024)
    g = torch.Generator(device="cuda").manual_seed(99)
    done = 0
    while done < 60:
        k = int(rng.choice([1, 2, 3, 3, 3, 5, 7]))
        stride = int(rng.choice([1, 1, 2]))
        pad = int(rng.integers(0, k // 2 + 2))
        H, W = int(rng.integers(max(k, 3), 71)), int(rng.integers(max(k, 3), 71))
        C = 8 * int(rng.integers(1, 26))
        Cout = 4 * int(rng.integers(1, 66))
        N = int(rng.integers(1, 5))
        if (H + 2 * pad - k) // stride + 1 < 1 or (W + 2 * pad - k) // stride + 1 < 1:
            continue
        act = torch.randint(0, 513, (N, H, W, C), device="cuda", generator=g)
        act = act * (torch.rand(N, H, W, C, device="cuda", generator=g) < 0.6)
        amp = max(1, min(256, int((2 ** 24 - 1) // (513 * C * k * k))))      # keep every partial sum below 2^24
        wgt = torch.randint(-amp, amp + 1, (k * k, Cout, C), device="cuda", generator=g)
        out = conv_codes.conv2d_codes(act.half(), wgt.half(), None, (k, k), stride, pad, 1.0)
        w_oihw = wgt.view(k, k, Cout, C).permute(2, 3, 0, 1).double()
        want = F.conv2d(act.permute(0, 3, 1, 2).double(), w_oihw, None, stride, pad).permute(0, 2, 3, 1)
        assert out.shape == want.shape, (N, H, W, C, Cout, k, stride, pad)
        assert torch.equal(out.double(), want), (N, H, W, C, Cout, k, stride, pad, float((out.double() - want).abs().max()))
        done += 1
