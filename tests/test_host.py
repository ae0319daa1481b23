"""CPU tests of the host-side logic: C ABI exports, drop-in module surface, op counters,
model-surgery rules and the pure-Python HESE helper (no compute calls: there is no GPU here)."""
import os
import re
import sys
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_capi_exports_every_declared_symbol():
    from term_quantization_b200 import _lib
    header = open(os.path.join(ROOT, "include", "tq_b200.h")).read()
    declared = set(re.findall(r"\b(tq_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = _lib.lib()                       # loads, binds every symbol (AttributeError otherwise)
    assert L.tq_version() == 201
    assert isinstance(_lib.launch_count(), int)


def test_library_missing_fails_loudly(monkeypatch):
    from term_quantization_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libtq_b200.so")
    with pytest.raises(_lib.TQError, match="no CPU fallback"):
        _lib.lib()


def test_tr_rejects_cpu_tensors_like_the_reference():
    from term_quantization_b200 import tr_cuda, tr_layer
    with pytest.raises(RuntimeError, match="input must be a CUDA tensor"):
        tr_cuda.tr(torch.zeros(2, 8), 1.0, 8, 8, 12)
    with pytest.raises(RuntimeError, match="CUDA"):
        tr_layer.LinearQuantize(8, 3)(torch.zeros(4))


def test_dropin_aliases():
    sys.path.insert(0, os.path.join(ROOT, "dropin"))
    try:
        import cnn_models
        import lstm_models.model as lm
        import profile_model
        import thop
        import tr_cuda
        import tr_layer
        import train_mlp
        import util
        for name in ("LinearQuantize", "TRConv2dLayer", "TRLinearLayer", "TRLSTMLayer", "set_tr_tracking",
                     "mse_profile", "hese", "compute_compressed_hese"):
            assert hasattr(tr_layer, name)
        assert callable(tr_cuda.tr) and callable(thop.profile) and callable(profile_model.get_model_ops)
        assert cnn_models.model_names() == ['alexnet', 'vgg16_bn', 'resnet18', 'efficientnet_b0', 'mobilenet_v2']
        assert train_mlp.MNISTMLP and lm.RNNModel and util.validate
    finally:
        sys.path.pop(0)


def test_python_hese_matches_reference_function():
    from term_quantization_b200 import tr_layer
    z = np.load(os.path.join(ROOT, "tests", "golden", "tr_golden.npz"))
    flat, off = z["hese_flat"], z["hese_off"]
    for i, v in enumerate(z["hese_in"]):
        assert tr_layer.hese(int(v)) == flat[off[i]:off[i + 1]].tolist(), v


def test_thop_counts_and_format():
    from term_quantization_b200 import thop
    m = nn.Sequential(nn.Conv2d(3, 8, 3, bias=False), nn.BatchNorm2d(8), nn.ReLU(), nn.AdaptiveAvgPool2d(1),
                      nn.Flatten(), nn.Linear(8, 10))
    ops, params = thop.profile(m, (torch.zeros(2, 3, 10, 10),))
    conv = 2 * 8 * 8 * 8 * 27
    bn = 2 * 2 * 8 * 8 * 8
    pool = (64 + 1) * 2 * 8
    assert ops == conv + bn + pool + 2 * 10 * 8 and params == 0
    assert not any("total_ops" in mod._buffers for mod in m.modules())
    assert thop.clever_format(1234567) == "1.23M"
    assert thop.clever_format([999, 2.5e9]) == ("999.00B", "2.50G")
    assert thop.count_hooks.zero_ops


def test_term_pair_formula_known_answers():
    # results/mnist-quant.json tmacs[0], results/mnist-tr.json tmacs[0] (evaluate_mlp.sh:3-4),
    # results/resnet18-group-size-results.json, results/mobilenet_v2-results.json
    from term_quantization_b200.profile_model import _term_pairs
    L = types.SimpleNamespace
    assert int(_term_pairs(L(num_terms=2, weight_bits=2, group_size=1, data_terms=6, data_bits=6), 668672)) == 8024064
    assert int(_term_pairs(L(num_terms=6, weight_bits=4, group_size=16, data_terms=6, data_bits=6), 668672)) == 1504512
    assert int(_term_pairs(L(num_terms=8, weight_bits=9, group_size=8, data_terms=3, data_bits=9), 1695547392)) == 5086642176
    assert int(_term_pairs(L(num_terms=9, weight_bits=6, group_size=1, data_terms=9, data_bits=9), 267939840)) == 14468751360


def test_conv_layer_settings_rules():
    from torchvision.models import mobilenet_v2, resnet18
    from term_quantization_b200 import cnn_models
    r = cnn_models.static_conv_layer_settings(resnet18(weights=None), 9, 8, 12)
    assert len(r) == 20 and r[0] == (16, 1, 16) and all(t == (9, 8, 12) for t in r[1:])
    mb = mobilenet_v2(weights=None)
    s = cnn_models.static_conv_layer_settings(mb, 9, 8, 12)
    convs = [m for m in mb.modules() if isinstance(m, nn.Conv2d)]
    assert len(s) == len(convs) == 52
    for i, (c, t) in enumerate(zip(convs, s)):
        assert t == ((16, 1, 16) if (i == 0 or c.groups > 1) else (9, 8, 12))
    assert sum(c.groups > 1 for c in convs) == 17


def test_wrapped_mac_counts_match_published_results():
    # MACs per image over the wrapped convs: ResNet-18 1,695,547,392 (results/resnet18-*.json)
    from torchvision.models import resnet18
    m = resnet18(weights=None).eval()
    macs = []

    def hook(mod, inp, out):
        kh, kw = mod.weight.shape[2:]
        macs.append(out.nelement() * (mod.in_channels // mod.groups) * kh * kw)
    hs = [c.register_forward_hook(hook) for c in m.modules() if isinstance(c, nn.Conv2d)]
    with torch.no_grad():
        m(torch.zeros(1, 3, 224, 224))
    for h in hs:
        h.remove()
    assert sum(macs[1:]) == 1695547392


def test_shard_bounds():
    from term_quantization_b200.inference import shard_bounds
    for n in (0, 1, 7, 256, 1000):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_models_build_on_cpu():
    from term_quantization_b200.lstm_models.model import RNNModel
    from term_quantization_b200.train_mlp import MNISTMLP
    assert MNISTMLP()(torch.zeros(3, 1, 28, 28)).shape == (3, 10)
    m = RNNModel('LSTM', 100, 16, 16, 2, tie_weights=True).eval()
    out, hid = m(torch.zeros(5, 3, dtype=torch.long), m.init_hidden(3))
    assert out.shape == (15, 100) and m.decoder.weight is m.encoder.weight
    with pytest.raises(ValueError):
        RNNModel('LSTM', 100, 16, 32, 2, tie_weights=True)


def test_bench_line_contract_on_committed_profile():
    """The bench line committed under profiles/ (written by bench.py on a B200) carries every key of the driver's
    contract plus this round's grid / baselines, and the ncu-traffic reader refuses a capture taken from another
    version of the kernel source."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = json.load(open(os.path.join(root, "profiles", "r02_bench_n1_first.json")))
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["unit"] == "images/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["gpu_launches"] > 0
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
    r = d["roofline"]
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(r) and r["bound"] in ("hbm", "tensor")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(c) and c["kind"] in ("reference", "port")
    assert len(d["tr_grid"]) >= 12 and all(set(("case", "elements", "GBs_best", "frac_of_nominal_8000")) <= set(r) for r in d["tr_grid"])
    assert d["kernel_to_beat"]["rows"] and set(d["other_configs"]) == {"vgg16_bn_b128", "mobilenet_v2_b512", "mlp_b256", "lstm_35x80"}
    assert d["e2e"]["h2d_ceiling_gbs"] > 0
    import bench
    traffic, src = bench.ncu_conv_traffic()
    assert traffic is None or traffic > 1e7           # None unless profiles/r02_conv_traffic.json matches csrc/tq_gemm.cu


def test_pybind_adapter_loads_and_keeps_the_reference_checks():
    """csrc/pybind/tr_cuda_pybind.cpp built ahead of time: the reference's module surface (kernels/tr_cuda.cpp:20-28)
    and its precondition messages (kernels/tr_cuda.cpp:12-18).  No compute without a GPU."""
    import pytest
    import torch
    from term_quantization_b200 import _lib, tr_cuda
    m = tr_cuda.pybind()
    assert m.version() == _lib.lib().tq_version()
    assert "Term Revealing (TR) (CUDA)" in m.tr.__doc__
    with pytest.raises(RuntimeError, match="input must be a CUDA tensor"):
        m.tr(torch.zeros(2, 8), 1.0, 8, 1, 3)


def test_fused_executor_host_rules():
    """Host-side rules of the fused executors added in round 2 (no GPU): which unwrapped first convs and max-pools go to
    the library kernels, the first-conv weight layout, the uint8 front end's dispatch, and the loud CPU rejections."""
    import pytest
    import torch
    import torch.nn as nn
    from term_quantization_b200 import conv_codes, fused, inference
    ok = nn.Conv2d(3, 64, 3, 1, 1)
    assert fused._first_conv_ok(ok) and fused._first_conv_ok(nn.Conv2d(3, 32, 3, 2, 1, bias=False))
    for bad in (nn.Conv2d(3, 64, 7, 2, 3), nn.Conv2d(3, 48, 3, 1, 1), nn.Conv2d(4, 64, 3, 1, 1), nn.Conv2d(3, 64, 3, 1, 0),
                nn.Conv2d(3, 64, 3, 1, 1, dilation=2), nn.Conv2d(3, 64, 3, 3, 1)):
        assert not fused._first_conv_ok(bad)
    assert fused._square(nn.MaxPool2d(2, 2)) and fused._square(nn.MaxPool2d(3, 2, 1)) and fused._square(nn.MaxPool2d(3))
    assert not fused._square(nn.MaxPool2d((2, 3), 2)) and not fused._square(nn.MaxPool2d(9, 2))
    # [filter row][filter column][input channel][output channel]
    w = torch.arange(64 * 27, dtype=torch.float32).view(64, 3, 3, 3)
    p = conv_codes.pack_first_conv_weight(w)
    assert p.shape == (3, 3, 3, 64) and p.is_contiguous() and float(p[1, 2, 0, 5]) == float(w[5, 0, 1, 2])
    with pytest.raises(RuntimeError):
        conv_codes.pack_first_conv_weight(torch.zeros(64, 3, 7, 7))
    with pytest.raises(RuntimeError, match="CUDA"):
        conv_codes.maxpool_codes(torch.zeros(1, 4, 4, 8, dtype=torch.float16), 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        conv_codes.first_conv3x3_fused(torch.zeros(1, 8, 8, 3), p, next_quant=(1.0, 8, 3))
    with pytest.raises(RuntimeError, match="uint8"):
        conv_codes.stem_conv7x7s2_u8(torch.zeros(1, 8, 8, 3), (0, 0, 0), (1, 1, 1), None)

    class Engine(nn.Module):                     # an engine that folds the normalisation into its first pass
        def forward_u8(self, x, mean, std):
            return ("folded", tuple(mean), tuple(std))
    out = inference.U8Frontend(Engine(), mean=(0.1, 0.2, 0.3), std=(1.0, 2.0, 3.0))(torch.zeros(1, 2, 2, 3, dtype=torch.uint8))
    assert out == ("folded", (0.1, 0.2, 0.3), (1.0, 2.0, 3.0))
