"""GPU parity of the tcgen05 convolution on term codes: the accumulators must equal the exact integer
convolution of the same codes (float64 conv of integer-valued tensors is exact) for EVERY input the code
range allows.  No test shrinks its data to stay below 2^24: the engine is chosen by the static proof of
conv_codes.plan_weight (kind::f16 with K-chunk accumulators where provable, kind::i8 planes otherwise)."""
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
    # CTA pairs (cta_group::2, at least one pair of M tiles per TPC): odd tile counts leave a tile without a partner
    (41, 28, 28, 128, 128, 3, 1, 1),
    (41, 14, 14, 256, 256, 3, 1, 1),
    (38, 20, 24, 64, 192, 3, 1, 1),
    # packed halo (small maps, several images per tile): single CTAs, then pairs, odd batches
    (75, 7, 7, 128, 128, 3, 1, 1),
    (151, 7, 7, 256, 512, 3, 1, 1),
    (37, 4, 4, 128, 256, 3, 1, 1),
    (301, 5, 6, 64, 128, 3, 1, 1),
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
    try:
        conv_codes.plan_weight(wgt.half().contiguous(), 512)
    except NotImplementedError:
        # refused by design: not provable on kind::f16 and C % 16 != 0 rules out the plane engine; the same geometry
        # with weights the proof accepts
        assert C % 16 != 0
        wgt = torch.randint(-48, 49, (k * k, Cout, C), device="cuda", generator=g)
    w_oihw = wgt.view(k, k, Cout, C).permute(2, 3, 0, 1).double()
    want = F.conv2d(act.permute(0, 3, 1, 2).double(), w_oihw, None, stride, pad).permute(0, 2, 3, 1)
    wh = wgt.half().contiguous()
    for engine in ("auto", "i8") if C % 16 == 0 else ("auto",):
        out = conv_codes.conv2d_codes(act.half(), wh, None, (k, k), stride, pad, 1.0, engine=engine)
        assert out.shape == want.shape
        # float(int32 accumulator): exact below 2^24, one RN conversion above
        assert torch.equal(out, want.float()), (engine, float((out.double() - want).abs().max()))
        # scale and bias in the epilogue
        out2 = conv_codes.conv2d_codes(act.half(), wh, bias, (k, k), stride, pad, 0.00123, engine=engine)
        want2 = (want.float() * np.float32(0.00123) + bias).float()
        assert torch.equal(out2, want2), engine


def _exact(act, wgt, k, stride, pad):
    Cout, C = wgt.shape[1], wgt.shape[2]
    w_oihw = wgt.view(k, k, Cout, C).permute(2, 3, 0, 1).double()
    return F.conv2d(act.permute(0, 3, 1, 2).double(), w_oihw, None, stride, pad).permute(0, 2, 3, 1)


def test_worst_case_codes_at_k4608_are_exact_or_refused():
    """SURVEY section 7 / VERDICT r1: boundary values +-2^bits at K = 4608 (ResNet-18 layer4: 3x3 x 512).
    Sum |a w| reaches 4608 * 512 * 256 = 6.0e8 >> 2^24: the kind::f16 engine must REFUSE (the static proof fails for
    any number of accumulator groups) and the automatic choice (kind::i8 planes, int32 accumulators) must be exact."""
    from term_quantization_b200 import conv_codes
    N, H, W, C, Cout, k = 3, 7, 7, 512, 512, 3
    g = torch.Generator(device="cuda").manual_seed(5)
    cases = {
        "all max": (torch.full((N, H, W, C), 512, device="cuda"), torch.full((k * k, Cout, C), 256, device="cuda")),
        "all max, negative weights": (torch.full((N, H, W, C), 512, device="cuda"), torch.full((k * k, Cout, C), -256, device="cuda")),
    }
    wr = torch.randint(0, 2, (k * k, Cout, C), device="cuda", generator=g) * 512 - 256           # +-256
    # adversarial activations: 512 exactly where output channel 0's weight is positive -> the largest positive sum
    adv = torch.zeros(N, H, W, C, device="cuda", dtype=torch.long)
    adv[:, 3, 3, :] = (wr[4, 0, :] > 0).long() * 512
    cases["+-256 weights, random 0/512 activations"] = (torch.randint(0, 2, (N, H, W, C), device="cuda", generator=g) * 512, wr)
    cases["adversarial activations"] = (adv, wr)
    for name, (act, wgt) in cases.items():
        want = _exact(act, wgt, k, 1, 1)
        wh = wgt.half().contiguous()
        with pytest.raises(NotImplementedError, match="cannot prove"):
            conv_codes.plan_weight(wh, 512, engine="f16")
        plan = conv_codes.plan_weight(wh, 512, engine="auto")
        assert plan.engine == "i8" and plan.planes_w == 2
        out = conv_codes.conv2d_codes(act.half().contiguous(), wh, None, (k, k), 1, 1, 1.0, plan=plan)
        assert float(want.abs().max()) > 2 ** 24 or name == "adversarial activations"
        assert torch.equal(out, want.float()), name
        assert torch.equal(out.double(), want) or float(want.abs().max()) >= 2 ** 24


@pytest.mark.parametrize("case,amp,groups", [
    ((3, 14, 14, 256, 256, 3, 1, 1), 60, 2), ((2, 14, 14, 256, 512, 3, 1, 1), 110, 4), ((4, 7, 7, 512, 512, 3, 1, 1), 40, 2),
    ((4, 7, 7, 512, 512, 3, 1, 1), 100, 4), ((2, 28, 28, 128, 64, 3, 2, 1), 150, 2), ((2, 14, 14, 512, 256, 1, 2, 0), 250, 2),
    ((2, 20, 20, 128, 128, 3, 1, 1), 180, 2), ((2, 28, 28, 128, 64, 3, 2, 1), 100, 1),
    # several tiles per CTA (more than 148 tiles): 2 accumulator stages of 2 groups / ONE stage of 4 groups
    ((96, 14, 14, 256, 256, 3, 1, 1), 60, 2), ((128, 7, 7, 512, 512, 3, 1, 1), 90, 4), ((40, 14, 14, 512, 256, 3, 2, 1), 75, 4)])
def test_k_chunk_accumulators_exact_beyond_2_24(case, amp, groups):
    """kind::f16 with the K dimension cut into accumulator groups: weights for which ONE fp32 accumulator cannot be
    proven exact but `groups` chunks can.  Activations are driven to the adversarial extreme (512 wherever output
    channel 0's weights are positive) so that real partial sums approach the per-chunk bound, and the int32 sum of
    the chunks exceeds 2^24."""
    from term_quantization_b200 import conv_codes
    N, H, W, C, Cout, k, stride, pad = case
    g = torch.Generator(device="cuda").manual_seed(sum(case) + amp)
    wgt = torch.randint(-amp, amp + 1, (k * k, Cout, C), device="cuda", generator=g)
    wh = wgt.half().contiguous()
    plan = conv_codes.plan_weight(wh, 512, engine="f16")
    assert plan.engine == "f16" and plan.groups == groups and plan.bound < 2 ** 24, plan
    act = torch.randint(0, 513, (N, H, W, C), device="cuda", generator=g)
    w0 = wgt[:, 0, :].view(k, k, C)
    act[0, :k, :k, :] = (w0 > 0).long() * 512                   # one patch maximises output channel 0
    want = _exact(act, wgt, k, stride, pad)
    out = conv_codes.conv2d_codes(act.half().contiguous(), wh, None, (k, k), stride, pad, 1.0, plan=plan)
    assert torch.equal(out, want.float()), float((out.double() - want).abs().max())
    if groups > 1:
        assert plan.bound * groups >= 2 ** 24                      # a single accumulator would not have been provable
    if N >= 40 and stride == 1:
        # the same layer with the whole fused tail (BN, residual, ReLU, fp32 tile + next codes): at this batch the halo
        # modes run on CTA pairs (and 7x7 maps on the packed halo), with 2 / 4 accumulator groups per tile
        from oracle import tq_oracle as O
        scale = np.float32(2.1e-7)
        a = torch.rand(Cout, device="cuda", generator=g) + 0.5
        b = torch.randn(Cout, device="cuda", generator=g)
        res = torch.randn(want.shape, device="cuda", generator=g)
        t = torch.relu(((want.float() * scale).double() * a.double() + b.double()).float() + res)
        nq = (max(float(t.max()), 1e-3) / 512, 9, 3)
        o2, codes = conv_codes.conv2d_codes_fused(act.half().contiguous(), wh, (k, k), stride, pad, scale, bn=(a, b), residual=res,
                                                  relu=True, next_quant=nq, plan=plan)
        assert torch.equal(o2, t)
        _, wc = O.tr(t.cpu().numpy().reshape(1, -1, 1, 1), nq[0], nq[1], 1, nq[2], return_codes=True)
        assert np.array_equal(codes.cpu().numpy().astype(np.int32).reshape(-1), wc.reshape(-1))


def test_i8_plane_engine_plane_counts():
    """kind::i8 engine with 1 or 2 planes per operand (codes within +-127 are a single plane), signed activations,
    fused epilogue; exact against the integer conv."""
    from oracle import tq_oracle as O
    from term_quantization_b200 import conv_codes
    g = torch.Generator(device="cuda").manual_seed(77)
    for (amax, wmax, pa, pw) in ((64, 8, 1, 1), (512, 64, 2, 1), (100, 256, 1, 2), (1024, 512, 2, 2)):
        for (N, H, W, C, Cout, k, stride, pad) in ((2, 9, 11, 48, 40, 3, 1, 1), (1, 1, 300, 656, 512, 1, 1, 0), (3, 14, 14, 256, 256, 3, 2, 1),
                                                   (64, 14, 14, 256, 256, 3, 1, 1)):      # > 148 tiles: several per CTA
            act = torch.randint(-amax, amax + 1, (N, H, W, C), device="cuda", generator=g)
            wgt = torch.randint(-wmax, wmax + 1, (k * k, Cout, C), device="cuda", generator=g)
            wh = wgt.half().contiguous()
            plan = conv_codes.plan_weight(wh, amax, signed_act=True, engine="i8")
            assert plan.engine == "i8" and plan.planes_w == pw
            want = _exact(act, wgt, k, stride, pad)
            out = conv_codes.conv2d_codes(act.half().contiguous(), wh, None, (k, k), stride, pad, 1.0, plan=plan)
            assert torch.equal(out, want.float()), (amax, wmax, N, H, W, C, Cout, k)
            # fused tail on the int32 accumulator
            scale = np.float32(1.7e-6)
            a = torch.rand(Cout, device="cuda", generator=g) + 0.5
            b = torch.randn(Cout, device="cuda", generator=g)
            t = torch.relu((want.float() * scale).double() * a.double() + b.double()).float()
            nq = (max(float(t.max()), 1e-3) / 512, 9, 3)
            out, codes = conv_codes.conv2d_codes_fused(act.half().contiguous(), wh, (k, k), stride, pad, scale, bn=(a, b),
                                                       relu=True, next_quant=nq, plan=plan)
            assert torch.equal(out, t)
            _, wc = O.tr(t.cpu().numpy().reshape(1, -1, 1, 1), nq[0], nq[1], 1, nq[2], return_codes=True)
            assert np.array_equal(codes.cpu().numpy().astype(np.int32).reshape(-1), wc.reshape(-1))


FUSED = [(2, 56, 56, 64, 64, 3, 1, 1), (3, 28, 28, 128, 128, 3, 1, 1), (2, 56, 56, 64, 128, 1, 2, 0),
         (5, 7, 7, 512, 512, 3, 1, 1), (1, 9, 13, 72, 72, 3, 1, 1), (3, 14, 14, 256, 256, 3, 1, 1),
         (2, 20, 33, 32, 40, 3, 1, 1), (1, 57, 29, 64, 64, 3, 1, 1), (2, 30, 27, 128, 192, 3, 1, 1),
         (41, 28, 28, 128, 128, 3, 1, 1), (151, 7, 7, 128, 256, 3, 1, 1), (9, 7, 7, 512, 512, 3, 1, 1)]   # CTA pairs / packed halo


def _random_fused_cases(n, seed=31):
    rng = np.random.default_rng(seed)
    cases = []
    while len(cases) < n:
        k = int(rng.choice([1, 3, 3, 3, 5]))
        stride = int(rng.choice([1, 1, 2]))
        pad = int(rng.integers(0, k // 2 + 1))
        H, W = int(rng.integers(max(k, 4), 61)), int(rng.integers(max(k, 4), 61))
        C = 8 * int(rng.integers(1, 17))
        Cout = 8 * int(rng.integers(1, 33))              # code output needs Cout % 8 == 0
        N = int(rng.integers(1, 4))
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
    if C % 16:
        # C % 16 != 0 cannot fall back to the plane engine: such a layer must be provable on kind::f16 or is refused
        try:
            conv_codes.plan_weight(wgt, 512)
        except NotImplementedError:
            pytest.skip("unprovable on kind::f16 and C % 16 != 0: refused by design")
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
    # (tiled kernel: 64-channel blocks on maps of >= 4 x 7 pooled pixels, ragged edges; the rest: one thread per output)
    for (N, H, W, C) in ((2, 112, 112, 64), (3, 17, 9, 8), (1, 7, 7, 20), (2, 57, 31, 128), (1, 23, 111, 24), (3, 30, 28, 64)):
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
    done, engines = 0, {}
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
        amp = int(rng.choice([256, 256, 64, 16]))            # full-range codes: the plan picks K chunks or the plane engine
        wgt = torch.randint(-amp, amp + 1, (k * k, Cout, C), device="cuda", generator=g)
        wh = wgt.half().contiguous()
        try:
            plan = conv_codes.plan_weight(wh, 512)
        except NotImplementedError:                          # C % 16 != 0 and unprovable on kind::f16: refused by design
            assert C % 16 != 0
            continue
        engines[plan.engine + str(plan.groups)] = engines.get(plan.engine + str(plan.groups), 0) + 1
        out = conv_codes.conv2d_codes(act.half(), wh, None, (k, k), stride, pad, 1.0, plan=plan)
        w_oihw = wgt.view(k, k, Cout, C).permute(2, 3, 0, 1).double()
        want = F.conv2d(act.permute(0, 3, 1, 2).double(), w_oihw, None, stride, pad).permute(0, 2, 3, 1)
        assert out.shape == want.shape, (N, H, W, C, Cout, k, stride, pad)
        assert torch.equal(out, want.float()), (plan, N, H, W, C, Cout, k, stride, pad, float((out.double() - want).abs().max()))
        done += 1
    print("engines used:", engines)
    assert any(e.startswith("i8") for e in engines) and any(e.startswith("f16") for e in engines)
