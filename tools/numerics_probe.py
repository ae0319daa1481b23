"""How far are the three ResNet-18 TQ paths from an fp64 evaluation of the same network, and how
far is cuDNN's fp32 path from itself under a different algorithm choice?  (Explains why logits of
the exact-integer path and the cuDNN float path differ by ~1e-3 while each conv agrees to 1e-6.)"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from term_quantization_b200 import fused, inference, tr_layer  # noqa: E402
dev = torch.device("cuda", 0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
model = bench.build_tq_resnet18(dev)
x = torch.randn(16, 3, 224, 224, device=dev)
inference.calibrate(model, [x])
sfs = [m.input_quant.sf for m in model.modules() if isinstance(m, tr_layer.TRConv2dLayer)]
with torch.no_grad():
    torch.backends.cudnn.benchmark = False
    y_a = model(x)
    torch.backends.cudnn.benchmark = True
    y_b = model(x)
    torch.backends.cudnn.deterministic = True
    y_c = model(x)
    torch.backends.cudnn.deterministic = False
    # fp64 evaluation: same TR'd weights (exactly representable), same scale factors, fp64 convs,
    # activations term-revealed by the fp64 instantiation of the kernel
    import copy
    m64 = copy.deepcopy(model).double()
    y64 = m64(x.double()).float()
    mt = copy.deepcopy(model).to(memory_format=torch.channels_last)
    tr_layer.use_tensor_cores(mt)
    y_tc = mt(x.contiguous(memory_format=torch.channels_last))
    y_f = fused.FusedResNet(mt)(x)
sc = float(y64.abs().max())
rel = lambda a, b: float((a - b).abs().max()) / sc   # noqa: E731
print(f"cudnn fp32 (benchmark off) vs fp64 : {rel(y_a, y64):.3e}")
print(f"cudnn fp32 (benchmark on)  vs fp64 : {rel(y_b, y64):.3e}")
print(f"cudnn fp32 deterministic   vs fp64 : {rel(y_c, y64):.3e}")
print(f"cudnn benchmark on vs off          : {rel(y_a, y_b):.3e}")
print(f"tcgen05 integer path       vs fp64 : {rel(y_tc, y64):.3e}")
print(f"fused integer path         vs fp64 : {rel(y_f, y64):.3e}")
print(f"fused vs unfused integer path      : {rel(y_f, y_tc):.3e}")
