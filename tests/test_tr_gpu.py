"""GPU parity tests: the CUDA path (through the C ABI, via the drop-in tr_cuda.tr) against the
CPU oracle and the golden vectors generated from the reference kernel body.  Bit-exact."""
import numpy as np
import pytest
import torch

from conftest import bits_equal, golden_cases
from oracle import tq_oracle as O

pytestmark = pytest.mark.gpu


def _tr(x_np, sf, bits, g, alpha, **kw):
    from term_quantization_b200 import tr_cuda
    x = torch.from_numpy(np.ascontiguousarray(x_np)).cuda()
    y = tr_cuda.tr(x, sf, bits, g, alpha, **kw)
    torch.cuda.synchronize()
    return y.cpu().numpy()


def test_library_loaded_and_counts_launches():
    from term_quantization_b200 import _lib
    n0 = _lib.launch_count()
    _tr(np.ones((1, 8), dtype=np.float32), 1.0, 8, 1, 2)
    assert _lib.launch_count() == n0 + 1


def test_golden_vectors():
    for x, sf, bits, g, alpha, y in golden_cases():
        assert bits_equal(_tr(x, sf, bits, g, alpha), y), (x.shape, x.dtype, bits, g, alpha)


SPECIALS = np.array([0.0, -0.0, np.nan, -np.nan, np.inf, -np.inf, 1e-45, -1e-45, 1e-38, 3e38, -3e38,
                     0.49999997, 0.5, 0.50000006, 1.5, 2.5, -0.5, -1.5, 254.5, 255.5, 1e9,
                     127.49999, 127.5, 0.013 * 7.5, 0.013 * 8.4999995], dtype=np.float32)


@pytest.mark.parametrize("n", [1, 3, 4, 5, 1023, 4099, 65536 + 7, (1 << 20) + 13])
def test_elementwise_sizes_and_specials(n):
    rng = np.random.default_rng(n)
    for bits, terms in ((8, 3), (9, 2), (4, 1), (12, 4), (13, 3), (16, 16), (8, 0), (8, 9), (1, 1)):
        x = (rng.standard_normal(n) * rng.uniform(0.01, 40)).astype(np.float32)
        x[: min(n, SPECIALS.size)] = SPECIALS[: min(n, SPECIALS.size)]
        sf = float(np.float32(rng.uniform(0.005, 0.5)))
        x = x.reshape(1, n, 1, 1)
        assert bits_equal(_tr(x, sf, bits, 1, terms), O.tr(x, sf, bits, 1, terms)), (n, bits, terms)


def test_elementwise_rounding_boundaries_dense():
    # every half-integer multiple of sf and its float neighbours: the +0.5-in-double rule
    for sf in (1.0, 0.05, 0.013, 3.0e-4, 0.37):
        sf32 = np.float32(sf)
        k = np.arange(0, 600, dtype=np.float32) * np.float32(0.5)
        base = (k * sf32).astype(np.float32)
        x = np.concatenate([base, np.nextafter(base, np.float32(np.inf)),
                            np.nextafter(base, np.float32(-np.inf)),
                            -base]).astype(np.float32)
        x = np.tile(x, 4)[None, :, None, None]
        for bits, terms in ((8, 8), (8, 2), (9, 3)):
            assert bits_equal(_tr(x, float(sf32), bits, 1, terms), O.tr(x, float(sf32), bits, 1, terms))


def test_elementwise_mse_profile_scale_factors():
    # the calibration sweep's arguments (tr_layer.py:44-48): grid in [-50, 50], sf from 1e-8
    x = np.linspace(-50, 50, 8192, dtype=np.float32).reshape(-1, 1, 1, 1)
    for sf in np.linspace(1e-8, 50, 2048, dtype=np.float32)[[0, 1, 2, 5, 100, 777, 2047]]:
        assert bits_equal(_tr(x, float(sf), 8, 1, 4), O.tr(x, float(sf), 8, 1, 4)), sf


def test_misaligned_and_inplace():
    from term_quantization_b200 import tr_cuda
    rng = np.random.default_rng(5)
    x = rng.standard_normal(10001).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    view = xd[1:].view(1, -1, 1, 1)                       # 4-byte aligned only
    assert view.data_ptr() % 16 != 0
    y = tr_cuda.tr(view, 0.02, 8, 1, 3).cpu().numpy()
    assert bits_equal(y, O.tr(x[1:].reshape(1, -1, 1, 1), 0.02, 8, 1, 3))
    z = xd.view(1, -1, 1, 1).clone()
    tr_cuda.tr(z, 0.02, 8, 1, 3, out=z)
    assert bits_equal(z.cpu().numpy(), O.tr(x.reshape(1, -1, 1, 1), 0.02, 8, 1, 3))
    w = torch.from_numpy(rng.standard_normal((64, 64, 3, 3)).astype(np.float32)).cuda()
    ref = O.tr(w.cpu().numpy(), 0.02, 8, 8, 12)
    tr_cuda.tr(w, 0.02, 8, 8, 12, out=w)
    assert bits_equal(w.cpu().numpy(), ref)


@pytest.mark.parametrize("dtype", [np.float64, "bfloat16", "float16"])
def test_other_dtypes(dtype):
    from term_quantization_b200 import tr_cuda
    rng = np.random.default_rng(11)
    for shape, g, alpha in (((1, 70001, 1, 1), 1, 3), ((8, 64, 3, 3), 8, 12), ((5, 650), 8, 12)):
        x64 = rng.standard_normal(shape) * 2
        sf = float(np.float32(np.abs(x64).max() / 128))
        if dtype is np.float64:
            got = _tr(x64, sf, 8, g, alpha)
            assert bits_equal(got, O.tr(x64, sf, 8, g, alpha))
        else:
            td = getattr(torch, dtype)
            xt = torch.from_numpy(x64).to(td)
            want = torch.from_numpy(O.tr(xt.float().numpy(), sf, 8, g, alpha)).to(td)
            got = tr_cuda.tr(xt.cuda(), sf, 8, g, alpha).cpu()
            assert torch.equal(got.view(torch.int16), want.view(torch.int16))


@pytest.mark.parametrize("g", [2, 3, 4, 5, 8, 12, 16, 32])
def test_grouped_weights_layouts(g):
    rng = np.random.default_rng(100 + g)
    # (O, I, k, k) with groups along I at stride k*k (tr_layer.py:120), (out, in) contiguous
    # (tr_layer.py:148), C % g != 0 tails (LSTM 650 % 8) and one-group tensors
    shapes = [(6, 96, 3, 3), (4, 96, 1, 1), (3, 96, 5, 5), (7, 960), (3, 650), (2, g), (1, 1), (5, 33, 3, 3)]
    for shape in shapes:
        for bits in (4, 8, 9, 15, 16):
            w = (rng.standard_normal(shape) * 0.1).astype(np.float32)
            sf = float(np.float32(np.abs(w).max() / 2 ** (bits - 1)))
            for alpha in (0, 1, g, int(1.5 * g), 3 * g, 100 * g):
                assert bits_equal(_tr(w, sf, bits, g, alpha), O.tr(w, sf, bits, g, alpha)), \
                    (shape, bits, g, alpha)


def test_grouped_ties_and_saturation():
    for g in (4, 8, 32):
        w = np.full((3, 4 * g), 4.0, dtype=np.float32)           # every term ties
        for alpha in range(0, 2 * g + 2):
            assert bits_equal(_tr(w, 1.0, 8, g, alpha), O.tr(w, 1.0, 8, g, alpha))
        w = np.full((2, 2 * g, 3, 3), 255.0, dtype=np.float32)   # 2 terms per value, 2^bits reachable
        w[:, ::3] *= -1
        for alpha in (1, g - 1, g, g + 1, 2 * g):
            assert bits_equal(_tr(w, 1.0, 8, g, alpha), O.tr(w, 1.0, 8, g, alpha))


def test_matches_reference_body_on_host():
    if not O.have_ref():
        pytest.skip("oracle/_ref not shipped")
    rng = np.random.default_rng(2)
    w = (rng.standard_normal((64, 128, 3, 3)) * np.sqrt(2.0 / (128 * 9))).astype(np.float32)
    sf = float(np.float32(np.abs(w).max() / 128))
    assert bits_equal(_tr(w, sf, 8, 8, 12), O.ref_tr(w, sf, 8, 8, 12))
    x = np.maximum(rng.standard_normal((1, 1 << 22, 1, 1)), 0).astype(np.float32)
    sf = float(np.float32(x.max() / 512))
    assert bits_equal(_tr(x, sf, 9, 1, 3), O.ref_tr(x, sf, 9, 1, 3))


def test_full_size_resnet_layer1_activation():
    # BASELINE.json config 2: largest activation 256x64x56x56 = 51,380,224 elements.
    from term_quantization_b200 import tr_cuda
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.relu(torch.randn(1, 256 * 64 * 56 * 56, 1, 1, device="cuda", generator=g))
    sf = float(x.max()) / 512
    y = tr_cuda.tr(x, sf, 9, 1, 3)
    codes = torch.round(y / np.float32(sf)).to(torch.int32)
    # size-independent properties: integer multiples of sf, at most 3 terms, within one
    # dropped-term bound of the plain quantised value, sign preserved
    assert torch.equal(codes.float() * np.float32(sf), y)
    m = codes.abs()
    assert int(m.max()) <= 512
    q = torch.clamp(torch.floor(x / np.float32(sf) + 0.5), max=511).to(torch.int32)
    assert int((m - q).abs().max()) <= 2 ** 6          # 9-bit HESE, 3 terms kept: error < 2^6
    assert bool((y >= 0).all())
    if O.have_ref():                                     # full equality against the reference body
        want = O.ref_tr(x.cpu().numpy(), sf, 9, 1, 3)
        assert bits_equal(y.cpu().numpy(), want)
    else:
        sub = x[:, : 1 << 22].contiguous()
        assert bits_equal(tr_cuda.tr(sub, sf, 9, 1, 3).cpu().numpy(), O.tr(sub.cpu().numpy(), sf, 9, 1, 3))


@pytest.mark.parametrize("enc", ["binary", "booth", "hese"])
def test_encodings_and_relu(enc):
    rng = np.random.default_rng(9)
    code = {"hese": O.ENC_HESE, "binary": O.ENC_BINARY, "booth": O.ENC_BOOTH}[enc]
    x = (rng.standard_normal((1, 50000, 1, 1)) * 3).astype(np.float32)
    w = (rng.standard_normal((16, 64, 3, 3))).astype(np.float32)
    for relu in (False, True):
        assert bits_equal(_tr(x, 0.03, 8, 1, 3, encoding=enc, relu=relu),
                          O.tr(x, 0.03, 8, 1, 3, encoding=code, relu=relu))
        assert bits_equal(_tr(w, 0.03, 8, 8, 12, encoding=enc, relu=relu),
                          O.tr(w, 0.03, 8, 8, 12, encoding=code, relu=relu))
        assert bits_equal(_tr(w, 0.03, 8, 3, 5, encoding=enc, relu=relu),
                          O.tr(w, 0.03, 8, 3, 5, encoding=code, relu=relu))


def test_golden_vectors_through_the_pybind_adapter():
    """The compiled torch extension (csrc/pybind/tr_cuda_pybind.cpp, the reference's `tr_cuda.tr` signature) over the
    same C ABI: every golden vector generated from the reference kernel body, new output tensor, input untouched,
    non-default stream, the reference's error for a non-contiguous input."""
    from conftest import golden_cases
    from term_quantization_b200 import tr_cuda
    m = tr_cuda.pybind()
    for x, sf, bits, g, alpha, y in golden_cases():
        xd = torch.from_numpy(x).cuda()
        keep = xd.clone()
        out = m.tr(xd, sf, bits, g, alpha)
        assert out.data_ptr() != xd.data_ptr() and torch.equal(xd, keep)
        assert bits_equal(out.cpu().numpy(), y), (x.shape, bits, g, alpha)
    s = torch.cuda.Stream()
    x = torch.randn(1, 1 << 20, 1, 1, device="cuda")
    with torch.cuda.stream(s):
        a = m.tr(x, 0.01, 8, 1, 3)
    s.synchronize()
    assert torch.equal(a, tr_cuda.tr(x, 0.01, 8, 1, 3))
    with pytest.raises(RuntimeError, match="input must be contiguous"):
        m.tr(torch.randn(4, 8, 3, 3, device="cuda").permute(0, 1, 3, 2), 0.01, 8, 8, 12)


def test_binary_and_booth_kernels_against_reference_held_fixtures():
    """The device kernels' BINARY / BOOTH paths against tests/golden/enc_golden.npz directly: BINARY = the bit planes
    of bit_utils.expand_binary_bits (bit_utils.py:63-73), BOOTH = verilog/booth_encoder.v:57-78 clocked bit-serially
    (tests/golden/make_golden_encodings.py).  g = 1: keeping k terms keeps the k most significant digits."""
    import os
    from conftest import ROOT
    from term_quantization_b200 import tr_cuda
    z = np.load(os.path.join(ROOT, "tests", "golden", "enc_golden.npz"))
    W, sf, bits, planes = z["bin_W"], float(z["bin_sf"]), int(z["bin_bits"]), z["bin_planes"]
    x = torch.from_numpy(W.reshape(1, -1, 1, 1).copy()).cuda()
    weights = np.arange(bits - 1, -1, -1)
    for k in (1, 2, 3, bits + 1):
        got = tr_cuda.tr_codes(x, sf, bits + 1, 1, k, dtype=torch.int32, encoding="binary").cpu().numpy().reshape(-1)
        first_k = (np.cumsum(planes, axis=1) <= k) & (planes > 0)
        want = (first_k.astype(np.int64) << weights).sum(1) * np.where(W < 0, -1, 1)
        assert np.array_equal(got, want), k
    # Booth: every 12-bit value as an exactly representable input with sf = 1
    q, P, N = z["booth_q"], z["booth_P"], z["booth_N"]
    xq = torch.from_numpy(q.astype(np.float32).reshape(1, -1, 1, 1)).cuda()
    for k in (1, 2, 4, 13):
        got = tr_cuda.tr_codes(xq, 1.0, 12, 1, k, dtype=torch.int32, encoding="booth").cpu().numpy().reshape(-1)
        want = np.zeros_like(q)
        for i in range(len(q)):
            kept, left = 0, k
            for pos in range(12, -1, -1):
                if left and ((int(P[i]) | int(N[i])) >> pos) & 1:
                    kept += (1 << pos) if (int(P[i]) >> pos) & 1 else -(1 << pos)
                    left -= 1
            want[i] = kept
        assert np.array_equal(got, want), k


@pytest.mark.parametrize("cdtype", [torch.int8, torch.uint8, torch.int16, torch.int32])
def test_integer_codes(cdtype):
    from term_quantization_b200 import tr_cuda
    rng = np.random.default_rng(21)
    info = torch.iinfo(cdtype)
    for shape, g, alpha, bits, relu in (((1, 100003, 1, 1), 1, 3, 8, True), ((1, 5000, 1, 1), 1, 2, 7, False),
                                        ((32, 64, 3, 3), 8, 12, 8, False), ((16, 650), 8, 12, 7, False),
                                        ((1, 3000, 1, 1), 1, 1, 8, False)):
        x = (rng.standard_normal(shape) * 2).astype(np.float32)
        sf = float(np.float32(np.abs(x).max() / 2 ** (bits - 1)))
        _, codes = O.tr(x, sf, bits, g, alpha, relu=relu, return_codes=True)
        ovf = torch.zeros(1, dtype=torch.int32, device="cuda")
        got = tr_cuda.tr_codes(torch.from_numpy(x).cuda(), sf, bits, g, alpha, dtype=cdtype,
                               relu=relu, overflow=ovf).cpu().numpy()
        want = np.clip(codes, info.min, info.max)
        assert np.array_equal(got.astype(np.int64), want.astype(np.int64)), (shape, cdtype)
        assert int(ovf.item()) == int((want != codes).any())


def test_argument_errors():
    from term_quantization_b200 import tr_cuda
    x = torch.zeros(2, 8, device="cuda")
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        tr_cuda.tr(torch.zeros(2, 8), 1.0, 8, 8, 12)
    with pytest.raises(RuntimeError, match="must be contiguous"):
        tr_cuda.tr(torch.zeros(8, 2, device="cuda").t(), 1.0, 8, 8, 12)
    with pytest.raises(RuntimeError):
        tr_cuda.tr(x.int(), 1.0, 8, 8, 12)
    for bad in (dict(sf=0.0), dict(sf=-1.0), dict(sf=float("nan")), dict(bits=0), dict(bits=17),
                dict(g=0), dict(g=33), dict(alpha=-1)):
        a = dict(sf=1.0, bits=8, g=8, alpha=12)
        a.update(bad)
        with pytest.raises(ValueError):
            tr_cuda.tr(x, a["sf"], a["bits"], a["g"], a["alpha"])
    # empty tensors are fine
    assert tr_cuda.tr(torch.zeros(0, 8, device="cuda"), 1.0, 8, 8, 12).shape == (0, 8)


def test_hoisted_reciprocal_divide_is_exact():
    """The stream kernels divide with a per-thread refined reciprocal + Markstein correction
    instead of a per-element div.rn.f32; on-device comparison over 2^31 pairs, plus an
    end-to-end comparison of the two kernel variants, plus scale factors outside the fast
    range (which must take the div.rn.f32 variant)."""
    from term_quantization_b200 import _lib
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    for seed in (1, 2):
        _lib.check(_lib.lib().tq_selftest_division(1 << 30, seed, bad.data_ptr(), None))
    torch.cuda.synchronize()
    assert int(bad.item()) == 0
    rng = np.random.default_rng(77)
    x = (rng.standard_normal((1, (1 << 22) + 5, 1, 1)) * 10).astype(np.float32)
    for sf in (0.0123, 1.0, 3.3e-7, 999.0):
        a = _tr(x, sf, 9, 1, 3)
        b = _tr(x, sf, 9, 1, 3, _exact_div=True)
        assert bits_equal(a, b)
    w = x[:, : 1 << 20].reshape(-1, 64, 4, 4)
    assert bits_equal(_tr(w, 0.05, 8, 8, 12), _tr(w, 0.05, 8, 8, 12, _exact_div=True))
    small = (rng.standard_normal((1, 5000, 1, 1)) * 1e-11).astype(np.float32)
    for sf in (1e-12, 1e-38, 3e12):
        xs = small if sf < 1 else x[:, :5000] * np.float32(1e12)
        assert bits_equal(_tr(xs, sf, 8, 1, 3), O.tr(xs, sf, 8, 1, 3)), sf


@pytest.mark.skipif(not O.have_ref_gpu(), reason="oracle/_ref/libtq_ref_gpu.so not built")
def test_matches_reference_kernel_on_the_gpu_at_full_sizes():
    """GPU-vs-GPU parity at BASELINE sizes: the reference's own kernel (kernels/tr_cuda_kernel.cu:58-125, body
    byte-identical, compiled for sm_100a by oracle/Makefile) against this repo's kernels on the same tensors,
    bit for bit: the largest ResNet-18 activation (51,380,224 values, g = 1) and every ResNet-18 conv-weight shape
    in the reference's OIHW layout (groups of 8 input channels, stride kh*kw), plus g in {2, 4, 16, 32}."""
    from term_quantization_b200 import tr_cuda
    gen = torch.Generator(device="cuda").manual_seed(12)
    x = torch.randn(1, 256 * 64 * 56 * 56, 1, 1, device="cuda", generator=gen) * 1.7      # both signs
    for bits, k in ((9, 3), (8, 4), (9, 2)):
        sf = float(x.abs().max()) / 2 ** bits
        assert torch.equal(tr_cuda.tr(x, sf, bits, 1, k).view(torch.int32), O.ref_gpu_tr(x, sf, bits, 1, k).view(torch.int32))
    for (o, i, kk) in ((64, 64, 3), (128, 64, 3), (128, 64, 1), (128, 128, 3), (256, 128, 3), (256, 256, 3),
                       (512, 256, 3), (512, 256, 1), (512, 512, 3)):
        w = torch.randn(o, i, kk, kk, device="cuda", generator=gen) * (2.0 / (i * kk * kk)) ** 0.5
        for bits in (8, 9):
            sf = float(w.abs().max()) / 2 ** (bits - 1)
            got, want = tr_cuda.tr(w, sf, bits, 8, 12), O.ref_gpu_tr(w, sf, bits, 8, 12)
            assert torch.equal(got.view(torch.int32), want.view(torch.int32)), (o, i, kk, bits)
    w2 = torch.randn(4096, 1024, device="cuda", generator=gen) * 0.03
    sf = float(w2.abs().max()) / 128
    for g_, a_ in ((2, 3), (4, 6), (8, 12), (16, 20), (32, 40)):
        assert torch.equal(tr_cuda.tr(w2, sf, 8, g_, a_).view(torch.int32), O.ref_gpu_tr(w2, sf, 8, g_, a_).view(torch.int32))
    # table paths of the grouped kernel: cumulative-count table (g <= 8, bits <= 10, alpha <= 127), 4-byte table else
    for bits_, g_, a_ in ((10, 8, 12), (10, 4, 5), (11, 8, 12), (12, 8, 9), (7, 8, 200), (6, 2, 1), (9, 8, 0), (10, 8, 127)):
        sf_ = float(w2.abs().max()) / 2 ** (bits_ - 1)
        assert torch.equal(tr_cuda.tr(w2, sf_, bits_, g_, a_).view(torch.int32),
                           O.ref_gpu_tr(w2, sf_, bits_, g_, a_).view(torch.int32)), (bits_, g_, a_)
    xd = torch.randn(2048, 512, device="cuda", generator=gen, dtype=torch.float64)
    assert torch.equal(tr_cuda.tr(xd, 0.01, 8, 8, 12).view(torch.int64), O.ref_gpu_tr(xd, 0.01, 8, 8, 12).view(torch.int64))
