"""GPU tests of everything around the TR op: calibration kernels, TR layer wrappers against the
reference stack on the CPU (oracle/ref_stack.py), op-counter known answers from results/*.json,
and the evaluate_* drivers end to end on synthetic data."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import ref_stack
from oracle import tq_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_math():
    # the float path is compared at 1e-5 relative: keep cuDNN/cuBLAS in true fp32 (no TF32)
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_histogram_matches_torch_histc_and_oracle():
    from term_quantization_b200 import tr_layer
    g = torch.Generator(device="cuda").manual_seed(3)
    q = tr_layer.LinearQuantize(8, 3).cuda()
    total = torch.zeros(8192, device="cuda")
    oracle_hist = np.zeros(8192, dtype=np.float32)
    for n, scale in ((100003, 5.0), (7, 1.0), (1 << 20, 30.0), (4096, 0.01)):
        x = torch.randn(n, device="cuda", generator=g) * scale
        x[: min(n, 5)] = torch.tensor([-50.0, 50.0, 0.0, 60.0, float("nan")], device="cuda")[: min(n, 5)]
        q(x.view(-1, 1))                                   # tracking mode: accumulate
        total += torch.histc(x, 8192, -50, 50)
        O.hist(x.cpu().numpy(), oracle_hist, -50, 50)
    assert torch.equal(q.hist_bins, total)
    assert np.array_equal(q.hist_bins.cpu().numpy(), oracle_hist)
    xb = (torch.randn(50001, device="cuda", generator=g) * 3).bfloat16()
    q2 = tr_layer.LinearQuantize(8, 3).cuda()
    q2(xb)
    assert torch.equal(q2.hist_bins, torch.histc(xb.float(), 8192, -50, 50))


def test_mse_profile_matches_oracle_and_reference_loop():
    from term_quantization_b200 import tr_cuda, tr_layer
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.relu(torch.randn(200000, device="cuda", generator=g)) * 2.5
    hist = torch.histc(x, 8192, -50, 50)
    for bits, terms in ((8, 3), (9, 2), (6, 6)):
        sf = tr_layer.mse_profile(hist, -50, 50, bits, terms)
        grid = torch.linspace(-50, 50, 8192)
        sfs = torch.linspace(1e-8, 50, 2048)
        idx, errs = O.mse_profile(hist.cpu().numpy(), grid.numpy(), sfs.numpy(), bits, terms)
        assert sf == sfs.tolist()[idx]
        # the reference's own loop (tr_layer.py:44-54) over a window around the minimum
        xg = grid.cuda()
        lo, hi = max(idx - 8, 0), min(idx + 9, 2048)
        ref = []
        for s in sfs.tolist()[lo:hi]:
            xh = tr_cuda.tr(xg.view(-1, 1, 1, 1), s, bits, 1, terms).view(-1)
            ref.append((hist * (xg - xh) ** 2).sum())
        assert lo + int(torch.argmin(torch.Tensor(ref))) == idx
        np.testing.assert_allclose(torch.Tensor(ref).numpy(), errs[lo:hi], rtol=2e-5)


def test_compute_compressed_hese_matches_python_loop():
    from term_quantization_b200 import tr_cuda, tr_layer
    g = torch.Generator(device="cuda").manual_seed(7)
    w = torch.randn(40, 96, device="cuda", generator=g) * 0.1
    sf = w.abs().max().item() / 128
    wq = tr_cuda.tr(w, sf, 8, 8, 12)
    ints = (wq / sf).int().view(-1).tolist()              # exactly tr_layer.py:60 on the device
    want = (int(np.ceil(np.log2(12))) + 2) * sum(len(tr_layer.hese(v)) for v in ints)
    assert tr_layer.compute_compressed_hese(wq, sf, 12) == want


def _small_cnn():
    torch.manual_seed(0)
    return nn.Sequential(
        nn.Conv2d(3, 16, 3, padding=1), nn.ReLU(),
        nn.Conv2d(16, 32, 3, padding=1, stride=2), nn.BatchNorm2d(32), nn.ReLU(),
        nn.Conv2d(32, 32, 3, padding=1, groups=32), nn.ReLU(),         # depthwise -> (16, 1, 16)
        nn.Conv2d(32, 64, 1), nn.ReLU(), nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(64, 10)).eval()


def test_tr_conv_stack_matches_reference_stack_on_cpu():
    from term_quantization_b200 import cnn_models, tr_layer
    base = _small_cnn()
    with torch.no_grad():
        for m in base.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
    cpu = ref_stack.convert_cnn(base, 8, 8, 12, 8, 3)
    gpu_base = _small_cnn()
    gpu_base.load_state_dict(base.state_dict())
    params = cnn_models.static_conv_layer_settings(gpu_base, 8, 8, 12)
    assert params == [(16, 1, 16), (8, 8, 12), (16, 1, 16), (8, 8, 12)]
    gpu = cnn_models.convert_model(gpu_base.cuda(), params, 8, 3)
    # weights: bit-exact
    cw = [m.conv.weight for m in cpu.modules() if isinstance(m, ref_stack.RefTRConv2d)]
    gw = [m.conv.weight for m in gpu.modules() if isinstance(m, tr_layer.TRConv2dLayer)]
    assert len(cw) == len(gw) == 3
    for a, b in zip(cw, gw):
        assert torch.equal(a, b.cpu())
    x = torch.randn(4, 3, 20, 20, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        # calibration on both sides: identical histograms -> identical scale factors
        cpu(x)
        gpu(x.cuda())
        tr_layer.set_tr_tracking(gpu, False)
        for m in ref_stack.quantizers(cpu):
            m.finish_tracking()
        sg = [m.input_quant.sf for m in gpu.modules() if isinstance(m, tr_layer.TRConv2dLayer)]
        sc = [m.sf for m in ref_stack.quantizers(cpu)]
        assert sg == sc
        want = cpu(x)
        got = gpu(x.cuda()).cpu()
    # floating-point logits: within 1e-5 relative (north star), fp32 conv on both sides
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5 * float(want.abs().max()))


def test_linear_and_lstm_layers_follow_reference_quirks():
    from term_quantization_b200 import tr_cuda, tr_layer
    torch.manual_seed(0)
    lin = nn.Linear(650, 40).cuda()
    w0 = lin.weight.detach().clone()
    layer = tr_layer.TRLinearLayer(lin, 8, 8, 8, 8, 12)
    sf = w0.abs().max().item() / 128
    assert layer.w_sf == sf
    assert torch.equal(layer.linear.weight, tr_cuda.tr(w0, sf, 8, 8, 12))   # 650 % 8 != 0 tail
    assert np.array_equal(layer.linear.weight.detach().cpu().numpy(), O.tr(w0.cpu().numpy(), sf, 8, 8, 12))
    x = torch.randn(5, 650, device="cuda")
    with torch.no_grad():
        y_track = layer(x)
        assert float(layer.input_quant.hist_bins.sum()) == x.numel()
        tr_layer.set_tr_tracking(layer, False)
        y = layer(x)
    assert torch.equal(y, y_track) and torch.equal(y, lin(x))     # xq is discarded (tr_layer.py:152-154)

    lstm = nn.LSTM(32, 32, 2).cuda()
    w_ih1 = lstm.weight_ih_l1.detach().clone()
    tl = tr_layer.TRLSTMLayer(lstm, 8, 8, 8, 8, 12)
    assert torch.equal(tl.lstm.weight_ih_l1, w_ih1)               # only layer 0 is term-revealed
    emb = torch.randn(7, 3, 32, device="cuda")
    hid = (torch.zeros(2, 3, 32, device="cuda"), torch.zeros(2, 3, 32, device="cuda"))
    with torch.no_grad():
        tl(emb, hid)
        assert float(tl.input_quant.hist_bins.sum()) == emb.numel() + 2 * hid[0].numel()
        tr_layer.set_tr_tracking(tl, False)
        out, (h, c) = tl(emb, hid)
    assert out.shape == (7, 3, 32) and torch.isfinite(out).all()


def test_op_counter_known_answers_from_results_json():
    from term_quantization_b200 import cnn_models, evaluate_mlp, profile_model
    from term_quantization_b200.train_mlp import MNISTMLP
    from torchvision.models import resnet18
    x = (torch.zeros(1, 1, 28, 28, device="cuda"),)
    # results/mnist-quant.json idx 0 (evaluate_mlp.sh:3): wb=2 wt=2 db=6 dt=6 gs=1
    m = evaluate_mlp.replace_linear_layers(MNISTMLP().cuda(), [(2, 1, 2)] * 3, 6, 6)
    assert profile_model.get_model_ops(m, x) == (8024064.0, 1337344.0)
    # results/mnist-tr.json idx 0 (evaluate_mlp.sh:4): wb=4 wt=6 gs=16
    m = evaluate_mlp.replace_linear_layers(MNISTMLP().cuda(), [(4, 16, 6)] * 3, 6, 6)
    tmacs, _ = profile_model.get_model_ops(m, x)
    assert tmacs == 1504512.0
    # results/resnet18-group-size-results.json: g=8, avg 1.0 term, data_terms=3 -> 5,086,642,176
    r = resnet18(weights=None).cuda().eval()
    q = cnn_models.convert_model(r, cnn_models.static_conv_layer_settings(r, 9, 8, 8), 9, 3)
    tmacs, params = profile_model.get_model_ops(q, (torch.zeros(1, 3, 224, 224, device="cuda"),))
    assert tmacs == 5086642176.0 and params == 0.0


def test_evaluate_drivers_run_on_synthetic_data(tmp_path):
    from term_quantization_b200 import evaluate_lstm, evaluate_mlp
    out = tmp_path / "mlp.json"
    res = evaluate_mlp.main(["--wb", "4", "8", "--wt", "12", "12", "--db", "6", "8", "--dt", "6", "4",
                             "--gs", "16", "8", "--out-file", str(out), "--samples", "512"])
    assert len(res["accs"]) == 2 and out.exists()
    assert res["tmacs"][0] == float(int(6 * (12 / 16) * 668672))
    res = evaluate_lstm.main(["--wb", "8", "--wt", "12", "--db", "8", "--dt", "8", "--gs", "8",
                              "--emsize", "64", "--nhid", "64", "--tokens", "1500"])
    assert len(res["ppls"]) == 1 and np.isfinite(res["ppls"][0])
    # decoder term pairs: 8 * (12/8) * (35*10*33278*64) accumulated in float32
    assert res["tmacs"][0] == float(np.float32(int(8 * 1.5 * 35 * 10 * 33278 * 64)))


def test_tensor_core_conv_layer_matches_float_path():
    """TRConv2dLayer.use_tensor_cores(): same layer, computed on integer codes by the tcgen05
    kernel.  Integer accumulators are exact (tests/test_conv_gpu.py); here the layer output is
    compared with the reference float path (fp32 cuDNN conv of the dequantised operands)."""
    from term_quantization_b200 import cnn_models, tr_layer
    base = _small_cnn().cuda()
    params = cnn_models.static_conv_layer_settings(base, 9, 8, 12)
    q = cnn_models.convert_model(base, params, 9, 3)
    x = torch.randn(6, 3, 40, 40, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    with torch.no_grad():
        q(x)
        tr_layer.set_tr_tracking(q, False)
        want = q(x)
        layers = [m for m in q.modules() if isinstance(m, tr_layer.TRConv2dLayer)]
        feats = {}
        hooks = [m.register_forward_hook(lambda mod, i, o, k=k: feats.__setitem__(k, (i[0], o)))
                 for k, m in enumerate(layers)]
        q(x)
        ref_feats = dict(feats)
        switched, skipped = tr_layer.use_tensor_cores(q)
        assert len(switched) == 2 and len(skipped) == 1 and "grouped" in skipped[0][1]
        got = q(x)
        for k in (0, 2):
            xin, yref = ref_feats[k]
            yin, ytc = feats[k]
            assert torch.equal(xin, yin)                      # same inputs up to that layer
            assert ytc.is_contiguous(memory_format=torch.channels_last)
            tol = 2e-6 * float(yref.abs().max()) * (layers[k].conv.in_channels * 9) ** 0.5
            assert float((ytc - yref).abs().max()) <= tol
        for h in hooks:
            h.remove()
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-4 * float(want.abs().max()))


def test_fused_resnet_matches_unfused_tensor_core_path():
    """FusedResNet (BN / residual / ReLU / next-layer encode in the conv epilogue) against the
    same model run layer by layer (use_tensor_cores) and against the float reference path."""
    from torchvision.models import resnet18
    from term_quantization_b200 import cnn_models, fused, inference, tr_layer
    torch.manual_seed(0)
    base = resnet18(weights=None).cuda().eval()
    with torch.no_grad():
        for mod in base.modules():
            if isinstance(mod, nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.2)
                mod.running_var.uniform_(0.5, 1.5)
                mod.weight.uniform_(0.5, 1.5)
                mod.bias.normal_(0, 0.2)
    q = cnn_models.convert_model(base, cnn_models.static_conv_layer_settings(base, 9, 8, 12), 9, 3)
    x = torch.randn(8, 3, 224, 224, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    inference.calibrate(q, [x])
    with torch.no_grad():
        ref = q(x)                                              # float path (fp32 cuDNN conv)
        q = q.to(memory_format=torch.channels_last)
        switched, skipped = tr_layer.use_tensor_cores(q)
        assert len(switched) == 19 and not skipped
        unfused = q(x.contiguous(memory_format=torch.channels_last))
        got = fused.FusedResNet(q, stem="cudnn")(x)
        f = fused.FusedResNet(q)                                # stem conv on the tensor cores too
        assert f.stem_w is not None
        got_tc_stem = f(x)
        # the one-kernel stem (pooling in the conv epilogue) is bit-identical to the default two-launch stem
        assert torch.equal(fused.FusedResNet(q, stem="tcgen05_pool")(x), got_tc_stem)
        # 16-bit images are used as they are: same logits as their fp32 upcast
        xb = x.bfloat16()
        assert torch.equal(f(xb), f(xb.float()))
    scale = float(ref.abs().max())
    d_unfused = float((unfused - ref).abs().max()) / scale
    d_fused = float((got - unfused).abs().max()) / scale
    d_stem = float((got_tc_stem - got).abs().max()) / scale
    print(f"rel diff: tensor-core vs float path {d_unfused:.3e}; fused vs unfused {d_fused:.3e}; "
          f"tcgen05 stem vs cuDNN stem {d_stem:.3e}")
    assert d_stem < 2e-2
    assert float((got_tc_stem.argmax(1) == ref.argmax(1)).float().mean()) >= 0.75
    # The float path is cuDNN fp32 (Winograd / FFT algorithms under cudnn.benchmark): its own
    # rounding noise, amplified by 19 re-quantisations, is what separates it from the exact
    # integer path (tools/numerics_probe.py measures both against an fp64 run).
    # The fused path evaluates BatchNorm as fma(x, a, b), 1-2 ulp from cuDNN's BatchNorm: the same
    # class of noise, so the same bound; with default BN statistics the two are bit-identical.
    assert d_unfused < 2e-2
    assert d_fused < 2e-2
    assert float((got.argmax(1) == ref.argmax(1)).float().mean()) >= 0.75


def _randomise_bn(model, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    with torch.no_grad():
        for mod in model.modules():
            if isinstance(mod, nn.BatchNorm2d):
                n = mod.num_features
                mod.running_mean.copy_(torch.randn(n, device="cuda", generator=g) * 0.2)
                mod.running_var.copy_(torch.rand(n, device="cuda", generator=g) + 0.5)
                mod.weight.copy_(torch.rand(n, device="cuda", generator=g) + 0.5)
                mod.bias.copy_(torch.randn(n, device="cuda", generator=g) * 0.2)


@pytest.mark.parametrize("engine", ["auto", "i8"])
def test_fused_resnet18_bit_exact_against_cpu_emulation(engine):
    """BASELINE configs[1] at the benchmarked setting (9-bit, g=8, alpha=12, 3 data terms), batch 8 at 224x224,
    non-trivial BatchNorm statistics: the output of the 20-launch fused chain (19 wrapped convs with BN / residual /
    ReLU / next-layer encode in the epilogue) must EQUAL, bit for bit, the CPU emulation of the same arithmetic
    (oracle/fused_emul.py: exact integer conv -> fl32 -> * scale -> fmaf BN -> + residual -> ReLU -> oracle.tr), and
    the logits must agree with an fp64 classifier on the emulated features within 1e-5 relative (north star).
    Both engines: 'auto' = kind::f16 with proven K-chunk accumulators, 'i8' = s8 planes on kind::i8 everywhere."""
    from torchvision.models import resnet18
    from oracle import fused_emul
    from term_quantization_b200 import cnn_models, fused, inference, tr_layer
    torch.manual_seed(0)
    base = resnet18(weights=None).cuda().eval()
    _randomise_bn(base)
    q = cnn_models.convert_model(base, cnn_models.static_conv_layer_settings(base, 9, 8, 12), 9, 3)
    x = torch.randn(8, 3, 224, 224, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)).bfloat16()
    inference.calibrate(q, [x.float()])
    with torch.no_grad():
        q = q.to(memory_format=torch.channels_last)
        f = fused.FusedResNet(q, engine=engine)
        cap = {}
        logits = f(x, capture=cap)
    desc = f.chain_description()
    used = sorted({(c["engine"], c["groups"]) for blk in desc for c in blk if c is not None})
    print("engines:", used)
    if engine == "auto":
        # the static proof needs 2 / 4 accumulator groups on the K = 2304 / 4608 layers of this random-init model
        assert all(e == "f16" for e, _ in used) and max(gp for _, gp in used) >= 2, used
    else:
        assert used == [("i8", 1)]
    emu = fused_emul.run_resnet_chain(desc, cap["stem"].cpu().numpy())
    got = cap["final"].cpu().numpy()
    assert got.shape == emu.shape
    assert np.array_equal(got.view(np.uint32), emu.view(np.uint32)) or np.array_equal(got, emu), \
        f"{int((got != emu).sum())} of {got.size} block outputs differ, max {float(np.abs(got - emu).max())}"
    # classifier on the emulated features in fp64
    feat = torch.from_numpy(emu).double().mean(dim=(1, 2))
    want = feat @ q.fc.weight.detach().double().cpu().t() + q.fc.bias.detach().double().cpu()
    rel = float((logits.double().cpu() - want).abs().max()) / float(want.abs().max())
    print(f"logits vs fp64 classifier on emulated features: {rel:.2e}")
    assert rel < 1e-5


def test_depthwise_and_bn_act_kernels_against_cpu_emulation():
    """tq_depthwise3x3_codes / tq_bn_act_encode (the memory-bound layers of the depthwise CNNs) bit for bit against
    the CPU emulation: exact int32 accumulator, fl32 * scale, fmaf BN, ReLU / ReLU6, oracle.tr codes."""
    from oracle import fused_emul
    from term_quantization_b200 import conv_codes
    g = torch.Generator(device="cuda").manual_seed(12)
    for (N, H, W, C, stride, relu) in ((2, 14, 14, 96, 1, "relu6"), (3, 17, 9, 32, 2, True), (1, 7, 7, 960, 1, "relu6"),
                                       (2, 112, 112, 32, 1, "relu6"), (2, 5, 4, 8, 2, False), (1, 56, 57, 144, 2, "relu6")):
        act = (torch.randint(0, 513, (N, H, W, C), device="cuda", generator=g) *
               (torch.rand(N, H, W, C, device="cuda", generator=g) < 0.6)).half()
        if not relu:
            act = act * (torch.randint(0, 2, (N, H, W, C), device="cuda", generator=g) * 2 - 1).half()
        w = torch.randint(-32768, 32769, (9, C), device="cuda", generator=g, dtype=torch.int32)
        a = torch.rand(C, device="cuda", generator=g) + 0.5
        b = torch.randn(C, device="cuda", generator=g)
        scale = float(np.float32(2.3e-8))
        dw = {"w": w.cpu().numpy(), "stride": stride, "scale": scale, "bias": None, "bn": (a.cpu().numpy(), b.cpu().numpy())}
        t_ref, _ = fused_emul.fused_depthwise(act.cpu().numpy().astype(np.int32), dw, relu=relu)
        nq = (max(float(np.abs(t_ref).max()), 1e-3) / 512, 9, 3)
        t_ref, c_ref = fused_emul.fused_depthwise(act.cpu().numpy().astype(np.int32), dw, relu=relu, next_quant=nq)
        out, codes = conv_codes.depthwise3x3_codes(act, w, stride, scale, bn=(a, b), relu=relu, want_f32=True, next_quant=nq,
                                                   act_unsigned=bool(relu))
        assert np.array_equal(out.cpu().numpy(), t_ref), (N, H, W, C, stride)
        assert np.array_equal(codes.cpu().numpy().astype(np.int32), c_ref), (N, H, W, C, stride)
        none, codes2 = conv_codes.depthwise3x3_codes(act, w, stride, scale, bn=(a, b), relu=relu, next_quant=nq)
        assert none is None and torch.equal(codes2, codes)
        # bn_act_encode on the fp32 tensor
        x = torch.randn(N, H, W, C, device="cuda", generator=g) * 3
        want = fused_emul._act(O.fma_channels(x.cpu().numpy(), a.cpu().numpy(), b.cpu().numpy()), relu)
        o2, c2 = conv_codes.bn_act_encode(x, (a, b), relu=relu, want_f32=True, next_quant=nq)
        assert np.array_equal(o2.cpu().numpy(), want)
        assert np.array_equal(c2.cpu().numpy().astype(np.int32), fused_emul.encode(want, nq))
        # ... with the producing conv's bias as one fp32 add before the affine (what torch's conv2d(bias=...) computes)
        bias = torch.randn(C, device="cuda", generator=g)
        want_b = fused_emul._act(O.fma_channels((x + bias).cpu().numpy(), a.cpu().numpy(), b.cpu().numpy()), relu)
        o3, c3 = conv_codes.bn_act_encode(x, (a, b), relu=relu, want_f32=True, next_quant=nq, bias=bias)
        assert np.array_equal(o3.cpu().numpy(), want_b)
        assert np.array_equal(c3.cpu().numpy().astype(np.int32), fused_emul.encode(want_b, nq))


def test_first_conv3x3_fused_against_torch():
    """tq_first_conv3x3_fused (the unwrapped first conv of VGG / MobileNet-V2 + bias + BN + ReLU(6) + first encode in one
    kernel): the fp32 result against torch's fp32 conv (different summation order: tolerance 2e-6 of the output range,
    the reference's own TF32 default is ~1e-3), the codes EXACTLY against the oracle encode of the kernel's own fp32
    result; ragged maps, both strides, both channel counts."""
    from oracle import fused_emul
    from term_quantization_b200 import conv_codes
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(21)
    for (N, H, W, Cout, stride, relu, with_bias) in ((2, 224, 224, 64, 1, True, True), (3, 96, 128, 32, 2, "relu6", False),
                                                    (1, 37, 53, 64, 2, True, True), (2, 9, 7, 32, 1, False, True),
                                                    (1, 8, 300, 64, 1, "relu6", False)):
        x = torch.randn(N, H, W, 3, device="cuda", generator=g)
        w = torch.randn(Cout, 3, 3, 3, device="cuda", generator=g) * 0.2
        bias = torch.randn(Cout, device="cuda", generator=g) if with_bias else None
        a = torch.rand(Cout, device="cuda", generator=g) + 0.5
        b = torch.randn(Cout, device="cuda", generator=g)
        ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), None if bias is None else bias.double(),
                                         stride, 1).permute(0, 2, 3, 1)
        ref = ref * a.double() + b.double()
        if relu:
            ref = ref.clamp(min=0)
        if relu == "relu6":
            ref = ref.clamp(max=6)
        nq = (max(float(ref.abs().max()), 1e-3) / 512, 9, 3)
        out, codes = conv_codes.first_conv3x3_fused(x, conv_codes.pack_first_conv_weight(w), stride, bias, (a, b), relu=relu,
                                                    want_f32=True, next_quant=nq)
        assert out.shape == ref.shape
        err = float((out.double() - ref).abs().max()) / max(float(ref.abs().max()), 1e-6)
        assert err < 2e-6, (N, H, W, Cout, stride, err)
        assert np.array_equal(codes.cpu().numpy().astype(np.int32), fused_emul.encode(out.cpu().numpy(), nq)), (N, H, W, Cout, stride)
        none, codes2 = conv_codes.first_conv3x3_fused(x, conv_codes.pack_first_conv_weight(w), stride, bias, (a, b), relu=relu,
                                                      next_quant=nq)
        assert none is None and torch.equal(codes2, codes)


def test_maxpool_on_codes_equals_torch():
    """tq_maxpool2d_f16 (the pools between the wrapped convs of the VGG-style stacks, run on fp16 term codes) against
    nn.MaxPool2d on the same tensor: VGG's 2x2/s2, AlexNet's 3x3/s2, ResNet's 3x3/s2/p1, odd maps, signed values."""
    from term_quantization_b200 import conv_codes
    g = torch.Generator(device="cuda").manual_seed(3)
    for (N, H, W, C, k, s, p) in ((2, 224, 224, 64, 2, 2, 0), (3, 13, 13, 256, 3, 2, 0), (2, 57, 31, 128, 3, 2, 1),
                                  (1, 7, 9, 8, 2, 2, 0), (5, 14, 14, 512, 2, 2, 0), (2, 6, 6, 24, 3, 1, 1)):
        codes = (torch.randint(-512, 513, (N, H, W, C), device="cuda", generator=g)).half()
        got = conv_codes.maxpool_codes(codes, k, s, p)
        want = torch.nn.functional.max_pool2d(codes.permute(0, 3, 1, 2), k, s, p).permute(0, 2, 3, 1)
        assert got.shape == want.shape and torch.equal(got, want.contiguous()), (N, H, W, C, k, s, p)
    with pytest.raises(Exception):
        conv_codes.maxpool_codes(torch.zeros(1, 4, 4, 12, device="cuda").half(), 2)          # C % 8 != 0


def test_fused_mobilenet_v2_bit_exact_against_cpu_emulation():
    """BASELINE configs[3]: fused.FusedMobileNet (1x1 convs on tcgen05 with BN / ReLU6 / residual / next encode in the
    epilogue, depthwise convs code-to-code) -- the tensor the average pool reads must EQUAL the CPU emulation of the
    chain bit for bit; the logits stay close to the reference's float path (cuDNN fp32, re-quantised 52 times)."""
    import torchvision
    from oracle import fused_emul
    from term_quantization_b200 import cnn_models, fused, inference
    torch.manual_seed(0)
    base = torchvision.models.mobilenet_v2(weights=None).cuda().eval()
    _randomise_bn(base, seed=5)
    q = cnn_models.convert_model(base, cnn_models.static_conv_layer_settings(base, 9, 8, 12), 9, 3)
    x = torch.randn(4, 3, 96, 128, device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
    inference.calibrate(q, [x])
    with torch.no_grad():
        ref = q(x)
        f = fused.FusedMobileNet(q)
        cap = {}
        got = f(x, capture=cap)
    emu = fused_emul.run_mobilenet_chain(f.chain_description(), cap["stem_codes"].cpu().numpy())
    final = cap["final"].cpu().numpy()
    assert final.shape == emu.shape
    assert np.array_equal(final, emu), f"{int((final != emu).sum())} of {emu.size} differ"
    rel = float((got - ref).abs().max()) / float(ref.abs().max())
    print(f"mobilenet_v2 fused vs float path: {rel:.2e}")
    assert got.shape == ref.shape and rel < 5e-2


@pytest.mark.parametrize("arch,size,expect_grouped", [("vgg16_bn", 64, 0), ("mobilenet_v2", 96, 17)])
def test_other_cnn_configs_layer_by_layer_on_tensor_cores(arch, size, expect_grouped):
    """BASELINE configs[2] / [3]: every wrapped conv of VGG-16-bn and MobileNet-V2 (random init, reference
    layer settings: depthwise convs get the (16, 1, 16) 'unquantised' setting and stay on the float path) run on
    the tcgen05 code-domain kernel must reproduce the float path's output for the SAME input: the integer
    accumulators are exact, so the two differ by fp32 rounding of the cuDNN conv only."""
    import torchvision
    from term_quantization_b200 import cnn_models, inference, tr_layer
    torch.manual_seed(0)
    base = getattr(torchvision.models, arch)(weights=None).cuda().eval()
    q = cnn_models.convert_model(base, cnn_models.static_conv_layer_settings(base, 9, 8, 12), 9, 3)
    x = torch.randn(2, 3, size, size, device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
    inference.calibrate(q, [x])
    layers = [(n, m) for n, m in q.named_modules() if isinstance(m, tr_layer.TRConv2dLayer)]
    feats = {}
    hooks = [m.register_forward_hook(lambda mod, i, o, k=n: feats.__setitem__(k, (i[0].detach(), o.detach())))
             for n, m in layers]
    with torch.no_grad():
        ref = q(x)
        for h in hooks:
            h.remove()
        switched, skipped = tr_layer.use_tensor_cores(q)
        assert len(skipped) == expect_grouped and all("grouped" in why for _, why in skipped)
        assert len(switched) == len(layers) - expect_grouped
        worst = 0.0
        for n, m in layers:
            if n not in switched:
                continue
            xin, yref = feats[n]
            ytc = m(xin)
            k = m.conv.in_channels * m.conv.kernel_size[0] * m.conv.kernel_size[1]
            tol = 2e-6 * max(float(yref.abs().max()), 1e-30) * k ** 0.5
            err = float((ytc - yref).abs().max())
            worst = max(worst, err / tol)
            assert err <= tol, (n, err, tol)
        got = q(x)
    assert got.shape == ref.shape and bool(torch.isfinite(got).all())
    print(f"{arch}: {len(switched)} convs on tcgen05, {len(skipped)} grouped on the float path, worst err/tol {worst:.2f}, "
          f"logit rel diff {float((got - ref).abs().max()) / max(float(ref.abs().max()), 1e-30):.2e}")


def test_evaluate_cnn_driver_engines(tmp_path):
    """evaluate_cnn (mirror of evaluate_cnn.py) end to end on synthetic data: the reference float path and the
    fused tensor-core engine give the same term-pair counts and a finite accuracy."""
    from term_quantization_b200 import evaluate_cnn
    res = {}
    for engine in ("float", "auto"):
        res[engine] = evaluate_cnn.main(["-a", "resnet18", "-b", "16", "--images", "32", "--quick", "--engine", engine,
                                         "--out-file", str(tmp_path / f"{engine}.json")])
        assert len(res[engine]["tr-data3"]["accs"]) == 1
    # data_terms * (alpha / g) * MACs = 3 * 1.5 * 1,695,547,392, accumulated in float32 (thop/profile.py:72-73)
    assert res["float"]["tr-data3"]["tmacs"] == res["auto"]["tr-data3"]["tmacs"]
    assert abs(res["float"]["tr-data3"]["tmacs"][0] / (3 * 1.5 * 1695547392) - 1) < 1e-6


def test_fused_resnet34_odd_resolution():
    """The fused engine on another BasicBlock ResNet and a ragged shape (batch 3, 160 x 192: 40 x 48 ... 5 x 6 maps,
    partial tiles everywhere): same logits as the layer-by-layer tensor-core path up to the fused BatchNorm's
    1-2 ulp, and the three stem variants agree."""
    from torchvision.models import resnet34
    from term_quantization_b200 import cnn_models, fused, inference, tr_layer
    torch.manual_seed(1)
    base = resnet34(weights=None).cuda().eval()
    q = cnn_models.convert_model(base, cnn_models.static_conv_layer_settings(base, 9, 8, 12), 9, 3)
    x = torch.randn(3, 3, 160, 192, device="cuda", generator=torch.Generator(device="cuda").manual_seed(8))
    inference.calibrate(q, [x])
    with torch.no_grad():
        q = q.to(memory_format=torch.channels_last)
        switched, skipped = tr_layer.use_tensor_cores(q)
        assert len(switched) == 35 and not skipped
        unfused = q(x.contiguous(memory_format=torch.channels_last))
        a = fused.FusedResNet(q)(x)
        b = fused.FusedResNet(q, stem="tcgen05_pool")(x)
        c = fused.FusedResNet(q, stem="cudnn")(x)
    scale = float(unfused.abs().max())
    assert torch.equal(a, b)
    assert float((a - c).abs().max()) / scale < 2e-2
    assert float((a - unfused).abs().max()) / scale < 2e-2
    assert float((a.argmax(1) == unfused.argmax(1)).float().mean()) >= 2 / 3


def test_fused_vgg_pools_on_codes():
    """fused.FusedVGG (BASELINE configs[2]): conv + bias + BN + ReLU + next encode in one launch per conv, max-pool on
    the fp16 codes.  Against the layer-by-layer tensor-core path (same integer contraction, BatchNorm by cuDNN) the
    logits differ only by the fused BatchNorm's 1-2 ulp; max-pooling on codes is exact."""
    import torchvision
    from term_quantization_b200 import cnn_models, fused, inference, tr_layer
    torch.manual_seed(0)
    base = torchvision.models.vgg16_bn(weights=None).cuda().eval()
    q = cnn_models.convert_model(base, cnn_models.static_conv_layer_settings(base, 9, 8, 12), 9, 3)
    x = torch.randn(4, 3, 96, 128, device="cuda", generator=torch.Generator(device="cuda").manual_seed(6))
    inference.calibrate(q, [x])
    with torch.no_grad():
        ref = q(x)
        q = q.to(memory_format=torch.channels_last)
        switched, skipped = tr_layer.use_tensor_cores(q)
        assert len(switched) == 12 and not skipped
        unfused = q(x.contiguous(memory_format=torch.channels_last))
        got = fused.FusedVGG(q)(x)
    scale = float(ref.abs().max())
    print(f"vgg16_bn: fused vs layer-by-layer {float((got - unfused).abs().max()) / scale:.2e}, "
          f"layer-by-layer vs float path {float((unfused - ref).abs().max()) / scale:.2e}")
    assert got.shape == ref.shape
    assert float((got - unfused).abs().max()) / scale < 2e-2
    # exactness of pooling on codes, on one layer's real codes
    codes = torch.randint(0, 513, (2, 8, 10, 64), device="cuda").half()
    vals = codes.float() * 0.0371
    pc = torch.nn.functional.max_pool2d(codes.permute(0, 3, 1, 2), 2, 2)
    pv = torch.nn.functional.max_pool2d(vals.permute(0, 3, 1, 2), 2, 2)
    assert torch.equal(pc.float() * 0.0371, pv)


def test_fused_vgg16_bit_exact_against_cpu_emulation():
    """BASELINE configs[2]: from the codes reaching the first wrapped conv to the map the average pool reads, FusedVGG
    (12 tcgen05 convs with bias / BN / ReLU / next encode in the epilogue, 4 max-pools on fp16 codes in the library's
    kernel, the last pool on the fp32 map) must EQUAL the CPU emulation of the chain bit for bit."""
    import torchvision
    from oracle import fused_emul
    from term_quantization_b200 import cnn_models, fused, inference
    torch.manual_seed(0)
    base = torchvision.models.vgg16_bn(weights=None).cuda().eval()
    _randomise_bn(base, seed=9)
    q = cnn_models.convert_model(base, cnn_models.static_conv_layer_settings(base, 9, 8, 12), 9, 3)
    x = torch.randn(2, 3, 64, 96, device="cuda", generator=torch.Generator(device="cuda").manual_seed(14))
    inference.calibrate(q, [x])
    f = fused.FusedVGG(q.to(memory_format=torch.channels_last))
    cap = {}
    with torch.no_grad():
        f(x, capture=cap)
    stages = f.chain_description()
    assert sum(s[0] == "conv" for s in stages) == 12 and sum(s[0] == "pool" for s in stages) == 5
    want = fused_emul.run_vgg_chain(stages, cap["stem_codes"].cpu().numpy().astype(np.int32))
    got = cap["final"].cpu().numpy()
    assert got.shape == want.shape and np.array_equal(got, want), float(np.abs(got - want).max())


def test_linear_and_lstm_input_projection_on_tensor_cores():
    """BASELINE configs[0] / [4]: the Linear layers of the MLP and the LSTM's layer-0 input projection
    W_ih . q(emb) (T x B = 2,800 tokens, K = 650 zero-padded to 656, 4H = 2,600; evaluate_lstm.py:17-37,
    tr_layer.py:174-195) on the tcgen05 kernel: exact int32 accumulators against an integer GEMM of the same codes, and
    the layers' outputs against the float path on the same quantised inputs."""
    from term_quantization_b200 import conv_codes, tr_cuda, tr_layer
    g = torch.Generator(device="cuda").manual_seed(2)
    # raw op: exact accumulators, ragged K and out
    for (M, K, N, amax, wmax) in ((2800, 650, 2600, 256, 128), (256, 784, 512, 64, 8), (256, 512, 10, 64, 8), (35, 650, 33278, 256, 128)):
        x = torch.randint(-amax, amax + 1, (M, K), device="cuda", generator=g)
        w = torch.randint(-wmax, wmax + 1, (N, K), device="cuda", generator=g)
        packed, _ = conv_codes.pack_linear_weight(w.float(), 1.0)
        out = conv_codes.linear_codes(x.half(), packed, 1.0, out_features=N, act_max=amax, signed_act=True)
        want = (x.double() @ w.double().t())
        assert out.shape == (M, N) and torch.equal(out, want.float()), (M, K, N)
    # TRLinearLayer: linear(q(x)) on codes == the float linear on the dequantised input, up to fp32 summation order
    torch.manual_seed(0)
    lin = nn.Linear(784, 512).cuda()
    layer = tr_layer.TRLinearLayer(lin, 6, 6, 4, 8, 12)
    x = torch.randn(256, 784, device="cuda", generator=g)
    with torch.no_grad():
        layer(x)
        tr_layer.set_tr_tracking(layer, False)
        ref = layer.linear(layer.input_quant(x))
        layer.use_tensor_cores()
        got = layer(x)
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) <= 2e-6 * float(ref.abs().max()) * 784 ** 0.5
    # TRLSTMLayer: integer input projection + step-wise layer 0 + cuDNN upper layer vs the reference forward
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    lstm = nn.LSTM(650, 650, 2).cuda().eval()
    L = tr_layer.TRLSTMLayer(lstm, 8, 8, 8, 8, 12)
    emb = torch.randn(35, 80, 650, device="cuda", generator=g) * 0.1
    hid = (torch.randn(2, 80, 650, device="cuda", generator=g) * 0.1, torch.randn(2, 80, 650, device="cuda", generator=g) * 0.1)
    with torch.no_grad():
        L(emb, hid)
        tr_layer.set_tr_tracking(L, False)
        y_ref, (h_ref, c_ref) = L(emb, hid)
        L.use_tensor_cores()
        # the projection itself: exact against the integer GEMM of the same codes
        q = L.input_quant
        codes = tr_cuda.tr_codes(emb.view(1, -1, 1, 1), q.sf, 8, 1, 8, dtype=torch.int32).view(-1, 650)
        wc = torch.round(lstm.weight_ih_l0 / np.float32(L.w_sf_ih))
        acc = codes.double() @ wc.double().t()
        scale = np.float32(q.sf) * np.float32(L.w_sf_ih)
        assert torch.equal(L.input_projection(emb).view(-1, 2600), acc.float() * scale)
        y, (h, c) = L(emb, hid)
    for a, b in ((y, y_ref), (h, h_ref), (c, c_ref)):
        assert a.shape == b.shape
        assert float((a - b).abs().max()) <= 1e-5 * max(float(b.abs().max()), 1.0), float((a - b).abs().max())


def test_u8_normalize_matches_torchvision_arithmetic():
    """tq_u8_normalize_bf16 = ToTensor + Normalize (util.py:12-27) in fp32, rounded to bf16; U8Frontend feeds a model."""
    from term_quantization_b200 import inference
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randint(0, 256, (3, 20, 12, 3), device="cuda", dtype=torch.uint8, generator=g)
    got = inference.normalize_u8(x)
    mean = torch.tensor(inference.IMAGENET_MEAN, device="cuda")
    std = torch.tensor(inference.IMAGENET_STD, device="cuda")
    want = ((x.float() / 255.0 - mean) / std).bfloat16().permute(0, 3, 1, 2)
    assert got.shape == want.shape and torch.equal(got, want)
    net = torch.nn.Conv2d(3, 4, 3).cuda().bfloat16()
    assert torch.equal(inference.U8Frontend(net)(x), net(want))


def test_fused_resnet_takes_uint8_images_through_the_stem_fold():
    """FusedResNet.forward_u8 (tq_stem_conv7x7s2_u8): ToTensor + Normalize inside the stem's fold pass -- the logits must
    EQUAL those of the same engine fed normalize_u8's bf16 images (same values, one pass and one tensor less)."""
    import torchvision
    from term_quantization_b200 import cnn_models, fused, inference
    torch.manual_seed(0)
    base = torchvision.models.resnet18(weights=None).cuda().eval()
    q = cnn_models.convert_model(base, cnn_models.static_conv_layer_settings(base, 9, 8, 12), 9, 3)
    g = torch.Generator(device="cuda").manual_seed(8)
    x8 = torch.randint(0, 256, (4, 96, 128, 3), device="cuda", dtype=torch.uint8, generator=g)
    inference.calibrate(q, [inference.normalize_u8(x8).float()])
    f = fused.FusedResNet(q.to(memory_format=torch.channels_last))
    with torch.no_grad():
        want = f(inference.normalize_u8(x8))
        got = inference.U8Frontend(f)(x8)
        assert torch.equal(got, want)
        # odd map: falls back to normalising first
        x_odd = x8[:, :95].contiguous()
        assert torch.equal(inference.U8Frontend(f)(x_odd), f(inference.normalize_u8(x_odd)))
