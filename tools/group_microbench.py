import sys, os, json, torch
sys.path.insert(0, '/root/repo')
from term_quantization_b200 import tr_cuda
n = 1 << 28
xs = [torch.randn(n // 512, 512, device="cuda") * 0.05 for _ in range(3)]
out = torch.empty_like(xs[0])
sf = float(xs[0].abs().max()) / 128
for bits, g, a in ((8, 8, 12), (9, 8, 12), (8, 16, 24), (8, 4, 6)):
    for i in range(3): tr_cuda.tr(xs[i % 3], sf, bits, g, a, out=out)
    torch.cuda.synchronize()
    ts = []
    for i in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); tr_cuda.tr(xs[i % 3], sf, bits, g, a, out=out); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(json.dumps({"case": f"grouped contiguous b={bits} g={g} a={a}", "n": n, "ms_best": ts[0], "GBs_best": n * 8 / ts[0] / 1e6, "GBs_med": n * 8 / ts[5] / 1e6}))
w = torch.randn(512, 512, 3, 3, device="cuda") * 0.05
ws = [w.clone() for _ in range(3)]
wo = torch.empty_like(w)
for i in range(3): tr_cuda.tr(ws[i % 3], sf, 8, 8, 12, out=wo)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10): tr_cuda.tr(ws[i % 3], sf, 8, 8, 12, out=wo)
e1.record(); e1.synchronize()
print(json.dumps({"case": "OIHW 512x512x3x3 g=8 (stride 9)", "GBs": w.numel() * 8 * 10 / e0.elapsed_time(e1) / 1e6}))
