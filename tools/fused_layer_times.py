"""Per-launch CUDA-event times of one FusedResNet forward (ResNet-18 TQ, batch 256): which launch of
this repo's kernels costs what, next to its MMA and HBM lower bounds.  Not a bench number."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from term_quantization_b200 import conv_codes, fused, inference, tr_layer  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = 5
arch = sys.argv[2] if len(sys.argv) > 2 else "resnet18"          # or vgg16_bn (BASELINE.json configs[2], batch 128)
x = torch.randn(batch, 3, 224, 224, device=dev).contiguous(memory_format=torch.channels_last)
if arch == "resnet18":
    model = bench.build_tq_resnet18(dev)
else:
    import torchvision
    from term_quantization_b200 import cnn_models
    torch.manual_seed(0)
    base = getattr(torchvision.models, arch)(weights=None).to(dev).eval()
    model = cnn_models.convert_model(base, cnn_models.static_conv_layer_settings(base, 9, 8, 12), 9, 3)
inference.calibrate(model, [x[:16]])
model = model.to(memory_format=torch.channels_last)
if arch == "resnet18":
    tr_layer.use_tensor_cores(model)
    f = fused.FusedResNet(model)
else:
    f = fused.FusedVGG(model)

records = []
recording = False


def wrap(name, describe):
    orig = getattr(conv_codes, name)

    def timed(*a, **k):
        if not recording:
            return orig(*a, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig(*a, **k)
        e1.record()
        records.append((name, describe(a, k, out), e0, e1))
        return out
    setattr(conv_codes, name, timed)


def d_conv(a, k, out):
    act, wgt = a[0], a[1]
    o = out[0] if out[0] is not None else out[1]
    flop = 2 * o.numel() * wgt.shape[0] * wgt.shape[2]
    nbytes = act.numel() * 2 + wgt.numel() * 2
    if out[0] is not None:
        nbytes += out[0].numel() * 4
    if out[1] is not None:
        nbytes += out[1].numel() * 2
    if k.get("residual") is not None:
        nbytes += k["residual"].numel() * 4
    return {"in": list(act.shape), "out": list(o.shape), "k": wgt.shape[0], "stride": a[3], "f32": out[0] is not None,
            "codes": out[1] is not None, "res": k.get("residual") is not None, "flop": flop, "bytes": nbytes}


def d_pool(a, k, out):
    return {"in": list(a[0].shape), "flop": 0,
            "bytes": a[0].numel() * 4 + out[0].numel() * 4 + (out[1].numel() * 2 if out[1] is not None else 0)}


def d_stem(a, k, out):
    o = out[0]
    return {"in": list(a[0].shape), "out": list(o.shape), "flop": 2 * o.numel() * 147,
            "bytes": a[0].numel() * 4 + o.numel() * 4}


wrap("conv2d_codes_fused", d_conv)
wrap("bn_relu_maxpool_encode", d_pool)
wrap("stem_conv7x7s2", d_stem)
wrap("stem_conv_pool", lambda a, k, out: {"in": list(a[0].shape), "out": list(out[0].shape), "flop": 2 * out[0].numel() * 4 * 147,
                                         "bytes": a[0].numel() * a[0].element_size() + out[0].numel() * 6})

with torch.no_grad():
    for _ in range(3):
        f(x)
    torch.cuda.synchronize()
    recording = True
    for _ in range(iters):
        f(x)
    torch.cuda.synchronize()
per = len(records) // iters
tot = 0.0
for i in range(per):
    name, d, _, _ = records[i]
    ms = sum(records[i + j * per][2].elapsed_time(records[i + j * per][3]) for j in range(iters)) / iters
    tot += ms
    d = dict(d)
    flop, nbytes = d.pop("flop"), d.pop("bytes")
    print(json.dumps({"i": i, "op": name, **d, "ms": round(ms, 4), "TFLOPs": round(flop / ms / 1e9, 1),
                      "GBs": round(nbytes / ms / 1e6, 1), "mma_floor_ms": round(flop / 1374.5e9, 4),
                      "hbm_floor_ms": round(nbytes / 6537.3e6, 4)}))
print(json.dumps({"sum_ms": tot}))
