"""MobileNet-V2 TQ (BASELINE.json configs[3], batch 512) through fused.FusedMobileNet: images/s and the per-kernel split."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchvision
from term_quantization_b200 import cnn_models, fused, inference

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 512
torch.manual_seed(0)
base = torchvision.models.mobilenet_v2(weights=None).cuda().eval()
q = cnn_models.convert_model(base, cnn_models.static_conv_layer_settings(base, 9, 8, 12), 9, 3)
x = torch.randn(batch, 3, 224, 224, device="cuda").bfloat16().float().contiguous(memory_format=torch.channels_last)
inference.calibrate(q, [x[:16]])
f = fused.FusedMobileNet(q)
from torch.profiler import profile, ProfilerActivity
with torch.no_grad():
    for _ in range(3):
        f(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        f(x)
    e1.record()
    torch.cuda.synchronize()
    print("ms/step", e0.elapsed_time(e1) / 5, "img/s", batch * 5 / e0.elapsed_time(e1) * 1e3)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        f(x)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
