"""Convolution / linear on term codes with the tcgen05 kernel (csrc/tq_gemm.cu).

`conv2d_codes` is the raw op (fp16 NHWC codes in, fp32 NHWC out).  `pack_conv_weight` turns a
term-revealed OIHW weight into the [R*S][Cout][Cin] fp16 code tensor the kernel reads."""
import torch

from . import _lib
from . import tr_cuda


def pack_conv_weight(w, w_sf, weight_bits, group_size, num_terms):
    """Term-reveal an (O, I, kh, kw) fp32 weight (groups along I, tr_layer.py:117-120) and
    return its integer codes as fp16 [kh*kw, O, I]."""
    codes = tr_cuda.tr_codes(w.detach().contiguous(), w_sf, weight_bits, group_size, num_terms,
                             dtype=torch.int16)
    O, I, kh, kw = codes.shape
    return codes.permute(2, 3, 0, 1).reshape(kh * kw, O, I).to(torch.float16).contiguous()


def conv2d_codes(act, wgt, bias, kernel_size, stride, pad, scale, out=None):
    """act: fp16 [N, H, W, C] codes; wgt: fp16 [R*S, Cout, C] codes; returns fp32 [N, Ho, Wo, Cout]."""
    if not (act.is_cuda and wgt.is_cuda and act.dtype == torch.float16 and wgt.dtype == torch.float16):
        raise RuntimeError("conv2d_codes expects fp16 CUDA code tensors")
    if not (act.is_contiguous() and wgt.is_contiguous()):
        raise RuntimeError("conv2d_codes expects contiguous NHWC activations and packed weights")
    N, H, W, C = act.shape
    R, S = kernel_size
    RS, Cout, C2 = wgt.shape
    if RS != R * S or C2 != C:
        raise RuntimeError("weight / activation shape mismatch")
    Ho = (H + 2 * pad - R) // stride + 1
    Wo = (W + 2 * pad - S) // stride + 1
    if out is None:
        out = torch.empty((N, Ho, Wo, Cout), dtype=torch.float32, device=act.device)
    with torch.cuda.device(act.device):
        rc = _lib.lib().tq_conv2d_codes_f16(
            act.data_ptr(), wgt.data_ptr(), bias.data_ptr() if bias is not None else None,
            out.data_ptr(), N, H, W, C, Cout, R, S, stride, pad, float(scale),
            torch.cuda.current_stream(act.device).cuda_stream)
    _lib.check(rc)
    return out


def conv2d_codes_fused(act, wgt, kernel_size, stride, pad, scale, *, bias=None, bn=None, residual=None,
                       relu=False, want_f32=True, next_quant=None):
    """Conv on codes with the fused tail (see tq_conv2d_codes_fused in include/tq_b200.h).

    bn = (a, b) fp32 [Cout] per-channel affine applied as fma(t, a, b); residual fp32
    [N, Ho, Wo, Cout]; next_quant = (sf, bits, terms) of the consumer's LinearQuantize.
    Returns (out_f32 or None, out_codes or None)."""
    N, H, W, C = act.shape
    R, S = kernel_size
    RS, Cout, C2 = wgt.shape
    if act.dtype != torch.float16 or wgt.dtype != torch.float16 or RS != R * S or C2 != C:
        raise RuntimeError("conv2d_codes_fused: bad operands")
    Ho = (H + 2 * pad - R) // stride + 1
    Wo = (W + 2 * pad - S) // stride + 1
    out = torch.empty((N, Ho, Wo, Cout), dtype=torch.float32, device=act.device) if want_f32 else None
    codes = torch.empty((N, Ho, Wo, Cout), dtype=torch.float16, device=act.device) if next_quant else None
    if residual is not None and (residual.shape != (N, Ho, Wo, Cout) or not residual.is_contiguous()
                                 or residual.dtype != torch.float32):
        raise RuntimeError("residual must be a contiguous fp32 [N, Ho, Wo, Cout] tensor")
    sf, bits, terms = next_quant if next_quant else (1.0, 1, 0)
    ptr = lambda t: t.data_ptr() if t is not None else None   # noqa: E731
    with torch.cuda.device(act.device):
        rc = _lib.lib().tq_conv2d_codes_fused(
            act.data_ptr(), wgt.data_ptr(), ptr(out), ptr(codes), ptr(bias),
            ptr(bn[0]) if bn else None, ptr(bn[1]) if bn else None, ptr(residual),
            N, H, W, C, Cout, R, S, stride, pad, float(scale), int(bool(relu)), float(sf), int(bits),
            int(terms), torch.cuda.current_stream(act.device).cuda_stream)
    _lib.check(rc)
    return out, codes


def bn_relu_maxpool_encode(x_nhwc, bn, relu=True, next_quant=None):
    """maxpool3x3/s2/p1(relu(fma(x, a, b))) on an fp32 [N, H, W, C] tensor, plus the fp16 term
    codes of the result for the first wrapped conv.  Returns (out [N, Ho, Wo, C], codes or None)."""
    if x_nhwc.dtype != torch.float32 or not x_nhwc.is_contiguous() or not x_nhwc.is_cuda:
        raise RuntimeError("bn_relu_maxpool_encode expects a contiguous fp32 CUDA [N, H, W, C] tensor")
    N, H, W, C = x_nhwc.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.empty((N, Ho, Wo, C), dtype=torch.float32, device=x_nhwc.device)
    codes = torch.empty((N, Ho, Wo, C), dtype=torch.float16, device=x_nhwc.device) if next_quant else None
    sf, bits, terms = next_quant if next_quant else (1.0, 1, 0)
    with torch.cuda.device(x_nhwc.device):
        rc = _lib.lib().tq_bn_relu_maxpool_encode(
            x_nhwc.data_ptr(), bn[0].data_ptr(), bn[1].data_ptr(), out.data_ptr(),
            codes.data_ptr() if codes is not None else None, N, H, W, C, int(bool(relu)),
            float(sf), int(bits), int(terms), torch.cuda.current_stream(x_nhwc.device).cuda_stream)
    _lib.check(rc)
    return out, codes


def pack_stem_weight(w, channel_sign=None):
    """(Cout, 3, 7, 7) fp32 -> fp16 [8][Cout][64] for tq_stem_conv7x7s2: the kernel padded to 8x8 and
    folded 2x2 (r = 2R + dr, s = 2S + ds), row R holding (S, dr, ds, c) with c padded to 4, as the
    two operand planes (w_hi, w_lo) with w = w_hi + w_lo in fp16 pairs.  `channel_sign` (+1 / -1 per
    output channel) is folded into the weights (the one-kernel stem pools before the BatchNorm affine and
    needs non-negative slopes: see stem_conv_pool)."""
    Cout, Cin, kh, kw = w.shape
    if (Cin, kh, kw) != (3, 7, 7):
        raise NotImplementedError("stem conv packing expects a (Cout, 3, 7, 7) weight")
    w8 = torch.zeros(Cout, 4, 8, 8, dtype=torch.float32, device=w.device)
    w8[:, :3, :7, :7] = w.detach().float()
    if channel_sign is not None:
        w8 = w8 * channel_sign.view(-1, 1, 1, 1).to(w8)
    # [co, c, R, dr, S, ds] -> [R, co, S, dr, ds, c]
    w2 = w8.view(Cout, 4, 4, 2, 4, 2).permute(2, 0, 4, 3, 5, 1).reshape(4, Cout, 64)
    hi = w2.half()
    lo = (w2 - hi.float()).half()
    return torch.cat([hi, lo], dim=0).contiguous()


_STEM_DTYPES = {torch.float32: _lib.TQ_F32, torch.bfloat16: _lib.TQ_BF16, torch.float16: _lib.TQ_F16}


def stem_conv7x7s2(x_nhwc, w2, scratch=None):
    """fp32 / bf16 / fp16 [N, H, W, 3] -> fp32 [N, H/2, W/2, Cout] (7x7 / stride 2 / pad 3, no bias)."""
    if x_nhwc.dtype not in _STEM_DTYPES or not x_nhwc.is_contiguous() or x_nhwc.shape[-1] != 3:
        raise RuntimeError("stem_conv7x7s2 expects a contiguous fp32 / bf16 / fp16 [N, H, W, 3] tensor")
    N, H, W, _ = x_nhwc.shape
    Cout = w2.shape[1]
    need = (2 if x_nhwc.dtype == torch.float32 else 1) * N * (H // 2 + 3) * (W // 2 + 3) * 16
    if scratch is None or scratch.numel() < need:
        scratch = torch.empty(need, dtype=torch.float16, device=x_nhwc.device)
    out = torch.empty((N, H // 2, W // 2, Cout), dtype=torch.float32, device=x_nhwc.device)
    with torch.cuda.device(x_nhwc.device):
        rc = _lib.lib().tq_stem_conv7x7s2_dt(x_nhwc.data_ptr(), _STEM_DTYPES[x_nhwc.dtype], scratch.data_ptr(),
                                             w2.data_ptr(), out.data_ptr(), N, H, W, Cout,
                                             torch.cuda.current_stream(x_nhwc.device).cuda_stream)
    _lib.check(rc)
    return out, scratch


def stem_pool_operands(w, bn):
    """Operands of the one-kernel stem for a (Cout, 3, 7, 7) weight and BatchNorm affine (a, b): the kernel pools
    the raw conv sums and applies the affine once per pooled value, which needs a >= 0, so sign(a) moves into the
    channel's weights (exact) and |a| is passed.  Returns (w2_signed, (|a|, b))."""
    a, b = bn
    sign = torch.where(a < 0, -torch.ones_like(a), torch.ones_like(a))
    return pack_stem_weight(w, sign), (a.abs().contiguous(), b.contiguous())


def stem_conv_pool(x_nhwc, w2, bn, relu=True, next_quant=None, scratch=None):
    """The whole stem in one launch: maxpool3x3/s2/p1(relu(fma(conv7x7s2(x), a, b))) -> fp32
    [N, Hp, Wp, Cout] plus the fp16 term codes of the result (next_quant = (sf, bits, terms)).
    Same values as stem_conv7x7s2 followed by bn_relu_maxpool_encode; the conv output never reaches HBM.
    `w2`, `bn` must come from stem_pool_operands (sign of the slope folded into the weights, bn[0] >= 0).
    Returns (out, codes or None, scratch)."""
    if x_nhwc.dtype not in _STEM_DTYPES or not x_nhwc.is_contiguous() or x_nhwc.shape[-1] != 3:
        raise RuntimeError("stem_conv_pool expects a contiguous fp32 / bf16 / fp16 [N, H, W, 3] tensor")
    N, H, W, _ = x_nhwc.shape
    Cout = w2.shape[1]
    need = (2 if x_nhwc.dtype == torch.float32 else 1) * N * (H // 2 + 3) * (W // 2 + 3) * 16
    if scratch is None or scratch.numel() < need:
        scratch = torch.empty(need, dtype=torch.float16, device=x_nhwc.device)
    Hp, Wp = (H // 2 - 1) // 2 + 1, (W // 2 - 1) // 2 + 1
    out = torch.empty((N, Hp, Wp, Cout), dtype=torch.float32, device=x_nhwc.device)
    codes = torch.empty((N, Hp, Wp, Cout), dtype=torch.float16, device=x_nhwc.device) if next_quant else None
    sf, bits, terms = next_quant if next_quant else (1.0, 1, 0)
    with torch.cuda.device(x_nhwc.device):
        rc = _lib.lib().tq_stem_conv7x7s2_pool(
            x_nhwc.data_ptr(), _STEM_DTYPES[x_nhwc.dtype], scratch.data_ptr(), w2.data_ptr(),
            bn[0].data_ptr(), bn[1].data_ptr(), int(bool(relu)), out.data_ptr(),
            codes.data_ptr() if codes is not None else None, N, H, W, Cout, float(sf), int(bits), int(terms),
            torch.cuda.current_stream(x_nhwc.device).cuda_stream)
    _lib.check(rc)
    return out, codes, scratch
