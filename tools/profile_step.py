"""Kernel-time breakdown of one bench step with torch.profiler (shares only, not a bench number)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from term_quantization_b200 import inference, tr_layer  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
backend = sys.argv[1] if len(sys.argv) > 1 else "tcgen05"
model = bench.build_tq_resnet18(dev)
x = torch.randn(256, 3, 224, 224, device=dev)
inference.calibrate(model, [x[:64]])
if backend in ("tcgen05", "fused"):
    model = model.to(memory_format=torch.channels_last)
    x = x.contiguous(memory_format=torch.channels_last)
    tr_layer.use_tensor_cores(model)
    if backend == "fused":
        from term_quantization_b200 import fused
        model = fused.FusedResNet(model)
with torch.no_grad():
    for _ in range(3):
        model(x)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            model(x)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=90))
