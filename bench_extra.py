"""Secondary measurements of bench.py, all taken AFTER the headline timed regions (rank 0, N = 1 only):

  tr_grid          BASELINE.md section 3's TR microbench grid: N in {2^20, 2^24, 51,380,224, 2^28} x (g, alpha, bits) in
                   {(1,3,9), (8,12,8), (8,12,9)}, the real conv-weight layout (OIHW, stride-9 groups) and the fused
                   bf16 -> 8-bit-code variant (3 B/element); GB/s against the measured and the nominal HBM peak
  kernel_to_beat   the reference kernel (kernels/tr_cuda_kernel.cu:58-125, byte-identical body recompiled for sm_100a:
                   oracle/_ref/libtq_ref_gpu.so) on the same tensors -- a BASELINE, executed from bench.py's baseline leg
  other_configs    BASELINE.json configs[0], [2], [3], [4]: MLP, VGG-16-bn, MobileNet-V2, LSTM throughput
  h2d_ceiling      copy-only loop over the e2e staging buffers: what the box's PCIe / host memory allows

Every timing: CUDA events on the launching stream around each launch, >= 3 warm-up launches, inputs rotating through
more than the 126 MB L2 (or larger than L2 themselves)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
NOMINAL_HBM = 8000.0
L2_BYTES = 126 << 20


def _time_launches(fn, nbuf, iters, warmup=3):
    """fn(i) launches once on buffer i % nbuf; returns (best ms, median ms) over `iters` event-timed launches."""
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    evs = []
    for i in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(i)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    return ms[0], ms[len(ms) // 2]


def _nbuf(bytes_per_buffer):
    return max(2, min(8, -(-2 * L2_BYTES // max(bytes_per_buffer, 1))))


def tr_grid(dev, peak, with_ref=True):
    """The TR-encode grid.  Returns (rows, kernel_to_beat rows)."""
    from term_quantization_b200 import tr_cuda
    rows, beat = [], []
    gen = torch.Generator(device=dev).manual_seed(0)
    sizes = [1 << 20, 1 << 24, 51380224, 1 << 28]
    settings = [(1, 3, 9), (8, 12, 8), (8, 12, 9)]
    ref_tr = None
    if with_ref:
        try:
            from oracle import tq_oracle as O
            if O.have_ref_gpu():
                ref_tr = O.ref_gpu_tr
        except Exception:
            ref_tr = None

    def row(name, n, bytes_per_elem, best, med, extra=None):
        ach = n * bytes_per_elem / (best * 1e-3) / 1e9
        r = {"case": name, "elements": n, "bytes_per_element": bytes_per_elem, "ms_best": best, "ms_median": med,
             "GBs_best": ach, "GBs_median": n * bytes_per_elem / (med * 1e-3) / 1e9,
             "frac_of_measured": ach / peak, "frac_of_nominal_8000": ach / NOMINAL_HBM}
        if extra:
            r.update(extra)
        return r

    for n in sizes:
        for (g, alpha, bits) in settings:
            nb = _nbuf(n * 4)
            if g == 1:     # activations, tr_layer.py:97-98: relu(randn), view (1, N, 1, 1), sf = max / 2^bits
                xs = [torch.relu(torch.randn(1, n, 1, 1, device=dev, generator=gen)) for _ in range(nb)]
                sf = float(xs[0].max()) / 2 ** bits
            else:          # weights, tr_layer.py:117-120 on a 2-D (out, in) tensor: groups of g consecutive `in` features
                xs = [torch.randn(n // 512, 512, device=dev, generator=gen) * (2.0 / 512) ** 0.5 for _ in range(nb)]
                sf = float(xs[0].abs().max()) / 2 ** (bits - 1)
            out = torch.empty_like(xs[0])
            iters = 20 if n <= (1 << 24) else 10
            best, med = _time_launches(lambda i: tr_cuda.tr(xs[i % nb], sf, bits, g, alpha, out=out), nb, iters)
            rows.append(row(f"fp32->fp32 g={g} alpha={alpha} bits={bits}", n, 8, best, med,
                            {"g": g, "alpha": alpha, "bits": bits, "layout": "contiguous groups"}))
            if ref_tr is not None and n in (1 << 24, 51380224) and (g, alpha, bits) != (8, 12, 9):
                rb, rm = _time_launches(lambda i: ref_tr(xs[i % nb].view(xs[0].shape), sf, bits, g, alpha, out=out), nb, 3, 1)
                beat.append(row(f"REFERENCE kernel, g={g} alpha={alpha} bits={bits}", n, 8, rb, rm,
                                {"speedup_of_this_repo": rb / best}))
            del xs, out
    # the real conv-weight layout: OIHW, groups of 8 input channels at a fixed (o, kh, kw): element stride 9
    for shape in ((512, 512, 3, 3), (2048, 1024, 3, 3)):
        n = 1
        for d in shape:
            n *= d
        nb = _nbuf(n * 4)
        ws = [torch.randn(*shape, device=dev, generator=gen) * (2.0 / (shape[1] * 9)) ** 0.5 for _ in range(nb)]
        sf = float(ws[0].abs().max()) / 256
        out = torch.empty_like(ws[0])
        best, med = _time_launches(lambda i: tr_cuda.tr(ws[i % nb], sf, 9, 8, 12, out=out), nb, 10)
        rows.append(row(f"fp32->fp32 g=8 alpha=12 bits=9 OIHW {shape}", n, 8, best, med, {"layout": "stride-9 groups (conv weight)"}))
        del ws, out
    # fused variant: bf16 activations in, 8-bit codes out (3 B/element); 7-bit quantiser so that every code (<= 128) fits u8
    n = 51380224
    nb = _nbuf(n * 2)
    xs = [torch.relu(torch.randn(1, n, 1, 1, device=dev, generator=gen)).bfloat16() for _ in range(nb)]
    sf = float(xs[0].float().max()) / 128
    out = torch.empty(xs[0].shape, dtype=torch.uint8, device=dev)
    best, med = _time_launches(lambda i: tr_cuda.tr_codes(xs[i % nb], sf, 7, 1, 3, dtype=torch.uint8, relu=True, out=out), nb, 20)
    rows.append(row("bf16->u8 codes, fused ReLU, g=1 terms=3 bits=7", n, 3, best, med, {"layout": "activation"}))
    out16 = torch.empty(xs[0].shape, dtype=torch.float16, device=dev)
    best, med = _time_launches(lambda i: tr_cuda.tr_codes(xs[i % nb], sf * 0.25, 9, 1, 3, dtype=torch.float16, relu=True, out=out16), nb, 20)
    rows.append(row("bf16->fp16 codes (operand format of the conv), fused ReLU, g=1 terms=3 bits=9", n, 4, best, med, {"layout": "activation"}))
    return rows, beat


def _time_model(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def _time_graphed(fn, iters, warmup=3):
    """ms per call of fn() replayed as ONE CUDA graph (like the headline engine); falls back to eager launches when the
    capture fails.  Returns (ms, 'cuda graph' | 'eager')."""
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            fn()
        for _ in range(2):
            graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters, "cuda graph"
    except Exception:
        torch.cuda.synchronize()
        return _time_model(fn, iters, warmup), "eager"


def other_configs(dev):
    """BASELINE.json configs[0], [2], [3], [4] on one GPU: units/s with inputs resident in HBM (synthetic, random init)."""
    import torchvision
    from term_quantization_b200 import _lib, cnn_models, evaluate_lstm, evaluate_mlp, fused, inference, tr_layer
    from term_quantization_b200.lstm_models.model import RNNModel
    from term_quantization_b200.train_mlp import MNISTMLP
    out = {}
    gen = torch.Generator(device=dev).manual_seed(0)

    def cnn(arch, batch, macs_per_image, key):
        torch.manual_seed(0)
        base = getattr(torchvision.models, arch)(weights=None).to(dev).eval()
        q = cnn_models.convert_model(base, cnn_models.static_conv_layer_settings(base, 9, 8, 12), 9, 3)
        xs = [torch.randn(batch, 3, 224, 224, device=dev, generator=gen).bfloat16().float()
              .contiguous(memory_format=torch.channels_last) for _ in range(2)]
        inference.calibrate(q, [xs[0][:16]])
        q = q.to(memory_format=torch.channels_last)
        engine, model = "tensor cores, layer by layer", None
        if arch.startswith("vgg"):
            model, engine = fused.FusedVGG(q), "fused.FusedVGG"
        elif arch.startswith("mobilenet") and hasattr(fused, "FusedMobileNet"):
            model, engine = fused.FusedMobileNet(q), "fused.FusedMobileNet"
        if model is None:
            tr_layer.use_tensor_cores(q)
            model = q
        state = {"i": 0}

        def step():
            state["i"] += 1
            with torch.no_grad():
                return model(xs[state["i"] & 1])
        n0 = _lib.launch_count()
        ms = _time_model(step, 5)
        launches = (_lib.launch_count() - n0) // 8
        out[key] = {"workload": f"{arch} TQ (9-bit, g=8, alpha=12, 3 data terms), batch {batch} at 3x224x224", "engine": engine,
                    "ms_per_step": ms, "value": batch / ms * 1e3, "unit": "images/s", "gpu_launches_per_step": launches,
                    "wrapped_conv_TFLOPs_if_all_time_were_conv": 2 * macs_per_image * batch / ms / 1e9}
        del model, q, base, xs
        torch.cuda.empty_cache()

    cnn("vgg16_bn", 128, 15259926528, "vgg16_bn_b128")
    cnn("mobilenet_v2", 512, 288656256, "mobilenet_v2_b512")

    # configs[0]: MNIST MLP 784-512-512-10, g=8 alpha=12 (wb 4 / db 6 as evaluate_mlp.sh:4), batch 256
    torch.manual_seed(0)
    mlp = MNISTMLP().to(dev).eval()
    mlp = evaluate_mlp.replace_linear_layers(mlp, evaluate_mlp.static_linear_layer_settings(mlp, 4, 8, 12), 6, 6)
    xm = torch.randn(256, 1, 28, 28, device=dev, generator=gen)
    with torch.no_grad():
        mlp(xm)
        tr_layer.set_tr_tracking(mlp, False)
        ms, how = _time_graphed(lambda: mlp(xm), 20)
        ms_eager = _time_model(lambda: mlp(xm), 20)
        tr_layer.use_tensor_cores(mlp)                      # linear(q(x)) on term codes, tcgen05 (non-strict semantics)
        mlp(xm)
        n1 = _lib.launch_count()
        mlp(xm)
        l_tc = _lib.launch_count() - n1
        ms_tc, how_tc = _time_graphed(lambda: mlp(xm), 20)
        ms_tc_eager = _time_model(lambda: mlp(xm), 20)
    out["mlp_b256"] = {"workload": "MNIST MLP 784-512-512-10 TQ (wb 4, g=8, alpha=12, db 6), batch 256",
                       "engine": "tcgen05 on term codes: linear(q(x)), exact int32 accumulators; " + how_tc, "ms_per_step": ms_tc,
                       "value": 256 / ms_tc * 1e3, "unit": "images/s", "gpu_launches_per_step": l_tc, "ms_per_step_eager": ms_tc_eager,
                       "strict_reference_path": {"what": "tr_layer.py:152-154 as shipped: the quantised input is discarded, cuBLAS fp32 "
                                                         "linear on the raw input with term-revealed weights; " + how,
                                                 "ms_per_step": ms, "value": 256 / ms * 1e3, "ms_per_step_eager": ms_eager},
                       "note": "launch-latency bound (0.17 MMAC per image): eager launches cost more than the arithmetic"}

    # configs[4]: LSTM 650/650 tied, vocab 33,278, seq 35 x batch 80, wb 8 / g 8 / alpha 12 / db 8 (evaluate_lstm.sh:4)
    torch.manual_seed(0)
    lstm = RNNModel('LSTM', evaluate_lstm.NTOKENS, 650, 650, 2, 0.5, True).to(dev).eval()
    lstm = evaluate_lstm.convert_model(lstm, evaluate_lstm.static_lstm_layer_settings(lstm, 8, 8, 12), 8, 8)
    tokens = torch.randint(evaluate_lstm.NTOKENS, (35, 80), device=dev, generator=gen)
    hidden = lstm.init_hidden(80)
    with torch.no_grad():
        lstm(tokens, hidden)
        tr_layer.set_tr_tracking(lstm, False)
        n0 = _lib.launch_count()
        lstm(tokens, hidden)
        l_strict = _lib.launch_count() - n0
        ms, how = _time_graphed(lambda: lstm(tokens, hidden), 10)
        tr_layer.use_tensor_cores(lstm)
        lstm(tokens, hidden)
        n1 = _lib.launch_count()
        lstm(tokens, hidden)
        l_tc = _lib.launch_count() - n1
        ms_tc, how_tc = _time_graphed(lambda: lstm(tokens, hidden), 10)
    out["lstm_35x80"] = {"workload": "Wikitext-2-shaped LSTM 650/650 tied, TQ on layer-0 gates and decoder (wb 8, g=8, alpha=12, db 8), "
                                     "seq 35 x batch 80",
                         "engine": "tcgen05 on term codes: layer-0 input projection W_ih.q(emb) and the 650 -> 33,278 decoder "
                                   "linear(q(x)) (60.6 GMAC per step), exact int32 accumulators; recurrence step-wise, layer 1 cuDNN; " + how_tc,
                         "ms_per_step": ms_tc, "value": 35 * 80 / ms_tc * 1e3, "unit": "tokens/s", "gpu_launches_per_step": l_tc,
                         "strict_reference_path": {"what": "the reference forward as shipped: cuDNN LSTM on quantised emb / h0 / c0, "
                                                           "decoder = cuBLAS fp32 linear on the RAW input (tr_layer.py:152-154)",
                                                   "ms_per_step": ms, "value": 35 * 80 / ms * 1e3, "timed_as": how,
                                                   "gpu_launches_per_step": l_strict}}
    return out


def h2d_ceiling(host_buffers, dev_buffers, steps, stream):
    """Copy-only loop over the e2e staging buffers (pinned host -> device, same sizes, same stream): ms per step."""
    for i in range(2):
        with torch.cuda.stream(stream):
            dev_buffers[i % len(dev_buffers)].copy_(host_buffers[i % len(host_buffers)], non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for i in range(steps):
            dev_buffers[i % len(dev_buffers)].copy_(host_buffers[i % len(host_buffers)], non_blocking=True)
        e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
