"""Throughput + agreement of the three ResNet-18 TQ paths (float / tensor-core / fused)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from term_quantization_b200 import fused, inference, tr_layer  # noqa: E402
dev = torch.device("cuda", 0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
model = bench.build_tq_resnet18(dev)
x = torch.randn(256, 3, 224, 224, device=dev)
inference.calibrate(model, [x[:64]])
def timeit(fn, n=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n, out
with torch.no_grad():
    t_ref, y_ref = timeit(lambda: model(x))
    model = model.to(memory_format=torch.channels_last)
    xc = x.contiguous(memory_format=torch.channels_last)
    tr_layer.use_tensor_cores(model)
    t_tc, y_tc = timeit(lambda: model(xc))
    f = fused.FusedResNet(model)
    t_f, y_f = timeit(lambda: f(xc))
    g = torch.cuda.CUDAGraph()
    static_x = xc.clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2): f(static_x)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        static_y = f(static_x)
    t_g, _ = timeit(lambda: g.replay())
    sc = float(y_ref.abs().max())
    print(f"float {t_ref:.3f} ms | tensor-core {t_tc:.3f} ms | fused {t_f:.3f} ms | fused+graph {t_g:.3f} ms")
    print(f"rel diff tc-vs-float {float((y_tc - y_ref).abs().max())/sc:.3e}  fused-vs-tc {float((y_f - y_tc).abs().max())/sc:.3e}  graph-vs-fused {float((static_y - y_f).abs().max())/sc:.3e}")
    print("argmax agreement fused vs float:", float((y_f.argmax(1) == y_ref.argmax(1)).float().mean()))
