"""VGG-16-bn TQ (BASELINE.json configs[2]: g=8, alpha=12, batch 128 at 224x224) on one B200: images/s of the float
path (TR kernels + cuDNN fp32 conv), the layer-by-layer tensor-core path and fused.FusedVGG.  Not the bench metric."""
import json
import os
import sys

import torch
import torchvision

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from term_quantization_b200 import cnn_models, fused, inference, tr_layer  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 128
torch.manual_seed(0)
base = torchvision.models.vgg16_bn(weights=None).cuda().eval()
q = cnn_models.convert_model(base, cnn_models.static_conv_layer_settings(base, 9, 8, 12), 9, 3)
x = torch.randn(batch, 3, 224, 224, device="cuda")
inference.calibrate(q, [x[:16]])


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / n, out


with torch.no_grad():
    t_f, y_f = timeit(lambda: q(x))
    q = q.to(memory_format=torch.channels_last)
    xc = x.contiguous(memory_format=torch.channels_last)
    tr_layer.use_tensor_cores(q)
    t_t, y_t = timeit(lambda: q(xc))
    f = fused.FusedVGG(q)
    t_v, y_v = timeit(lambda: f(xc))
macs = 15259926528                      # wrapped convs, per image
for name, t in (("float path (TR + cuDNN fp32)", t_f), ("tensor cores, layer by layer", t_t), ("fused.FusedVGG", t_v)):
    print(json.dumps({"engine": name, "batch": batch, "ms": round(t, 3), "images_per_s": round(batch / t * 1e3, 1),
                      "wrapped_conv_TFLOPs_if_all_time_were_conv": round(2 * macs * batch / t / 1e9, 1)}))
print(json.dumps({"fused_vs_layerwise_max_rel_diff": float((y_v - y_t).abs().max()) / float(y_t.abs().max())}))
