import os, sys, torch, torchvision
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from term_quantization_b200 import cnn_models, fused, inference
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda", 0)
torch.manual_seed(0)
base = torchvision.models.vgg16_bn(weights=None).to(dev).eval()
q = cnn_models.convert_model(base, cnn_models.static_conv_layer_settings(base, 9, 8, 12), 9, 3)
x = torch.randn(128, 3, 224, 224, device=dev).contiguous(memory_format=torch.channels_last)
inference.calibrate(q, [x[:16]])
q = q.to(memory_format=torch.channels_last)
f = fused.FusedVGG(q)
with torch.no_grad():
    for _ in range(3):
        f(x)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            f(x)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
