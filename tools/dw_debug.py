import sys, numpy as np, torch
sys.path.insert(0, ".")
from oracle import fused_emul
from term_quantization_b200 import conv_codes
g = torch.Generator(device="cuda").manual_seed(12)
for (N, H, W, C, stride, relu) in ((2, 14, 14, 96, 1, "relu6"), (3, 17, 9, 32, 2, True), (1, 7, 7, 960, 1, "relu6"),
                                   (2, 112, 112, 32, 1, "relu6"), (2, 5, 4, 8, 2, False), (1, 56, 57, 144, 2, "relu6"), (2, 14, 14, 64, 1, True), (2,14,14,64,2,True)):
    act = (torch.randint(0, 513, (N, H, W, C), device="cuda", generator=g) * (torch.rand(N, H, W, C, device="cuda", generator=g) < 0.6)).half()
    w = torch.randint(-32768, 32769, (9, C), device="cuda", generator=g, dtype=torch.int32)
    scale = float(np.float32(2.3e-8))
    dw = {"w": w.cpu().numpy(), "stride": stride, "scale": scale, "bias": None, "bn": None}
    t_ref, _ = fused_emul.fused_depthwise(act.cpu().numpy().astype(np.int32), dw, relu=False)
    for un in (False, True):
        out, _ = conv_codes.depthwise3x3_codes(act, w, stride, scale, relu=False, want_f32=True, act_unsigned=un)
        o = out.cpu().numpy()
        bad = np.argwhere(o != t_ref)
        print((N, H, W, C, stride), "unsigned", un, "mismatches", len(bad), bad[:4].tolist() if len(bad) else "")
