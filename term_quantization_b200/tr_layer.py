"""Drop-in for the reference module ``tr_layer`` (tr_layer.py:1-201).

Same public names, constructor signatures and attributes:

    LinearQuantize(data_bits, data_terms)                        tr_layer.py:78-104
    TRConv2dLayer / TRLinearLayer / TRLSTMLayer(layer, data_bits=8, data_terms=4,
        weight_bits=8, group_size=1, num_terms=8)                tr_layer.py:106-201
    set_tr_tracking(model, tracking)                             tr_layer.py:66-76
    mse_profile(hist, minv, maxv, bit_width, terms)              tr_layer.py:43-54
    hese(number), compute_compressed_hese(w, sf, weight_terms)   tr_layer.py:9-41, 57-63

What changed underneath: every ``tr_cuda.tr`` call lands in the sm_100a kernels of
libtq_b200.so; the tracking-mode ``torch.histc`` + add is one fused histogram pass; the
calibration sweep (2048 launches + 2048 syncs per layer in the reference) is one fused kernel;
``compute_compressed_hese`` is a device-side popcount reduction instead of a Python loop over
every weight.  There is no CPU path.

Reference quirks are kept by default (``STRICT_REFERENCE = True``): TRLinearLayer.forward
feeds the *unquantised* input to the linear (tr_layer.py:152-154), TRLSTMLayer only term-reveals
layer 0 and leaves ``w_sf`` at the hh value (tr_layer.py:174-186).  Set
``tr_layer.STRICT_REFERENCE = False`` to have TRLinearLayer use the quantised input.
"""
import math

import torch
import torch.nn as nn

from . import _lib
from . import tr_cuda

STRICT_REFERENCE = True

__all__ = ["hese", "mse_profile", "compute_compressed_hese", "set_tr_tracking", "LinearQuantize",
           "TRConv2dLayer", "TRLinearLayer", "TRLSTMLayer", "tr_cuda", "STRICT_REFERENCE"]


def hese(number):
    """Signed power-of-two terms of an integer under HESE, smallest magnitude first
    (same list as tr_layer.py:9-41): an isolated 1-bit stays +2^i, a run of ones [lo..hi]
    becomes -2^lo, +2^(hi+1); everything is negated for negative numbers."""
    number = int(number)
    sign = -1 if number < 0 else 1
    q = abs(number)
    out, i = [], 0
    while q >> i:
        if not (q >> i) & 1:
            i += 1
            continue
        lo = i
        while (q >> i) & 1:
            i += 1
        if i - lo == 1:
            out.append(sign * (1 << lo))
        else:
            out.append(-sign * (1 << lo))
            out.append(sign * (1 << i))
    return out


def mse_profile(hist, minv, maxv, bit_width, terms):
    """Scale factor minimising the histogram-weighted squared error of g=1 term quantisation
    (tr_layer.py:43-54).  Same grid (linspace(minv, maxv, len(hist)) on the device), same 2048
    candidates, same first-minimum rule; one fused sweep on the device."""
    dev = hist.device if hist.is_cuda else torch.device("cuda")
    hist = hist.to(device=dev, dtype=torch.float32).contiguous()
    x = torch.linspace(minv, maxv, len(hist)).to(dev)
    sfs_t = torch.linspace(1e-8, maxv, 2048)
    sfs = sfs_t.tolist()
    sfs_d = sfs_t.to(dev)
    errs = torch.empty(len(sfs), dtype=torch.float64, device=dev)
    argmin = torch.empty(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().tq_mse_profile(hist.data_ptr(), x.data_ptr(), len(hist), sfs_d.data_ptr(),
                                       len(sfs), int(bit_width), int(terms), errs.data_ptr(),
                                       argmin.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc)
    return sfs[int(argmin.item())]


def compute_compressed_hese(w, sf, weight_terms):
    """Parameter bits of a HESE-compressed weight tensor (tr_layer.py:57-63):
    (ceil(log2(weight_terms)) + 2) bits per surviving term."""
    exp_bits = math.ceil(math.log2(weight_terms))
    bit_width = exp_bits + 2          # 1 for sign, 1 for barrier
    w = w.detach()
    if not w.is_cuda:
        raise RuntimeError("w must be a CUDA tensor")
    w = w.contiguous()
    count = torch.zeros(1, dtype=torch.int64, device=w.device)
    dt = {torch.float32: _lib.TQ_F32, torch.bfloat16: _lib.TQ_BF16}[w.dtype]
    with torch.cuda.device(w.device):
        # `w / sf` with a Python float on a CUDA tensor is w * (1/sf) inside torch
        rc = _lib.lib().tq_hese_term_count(w.data_ptr(), dt, w.numel(), float(sf),
                                           _lib.FLAG_RECIP_DIV, count.data_ptr(),
                                           torch.cuda.current_stream(w.device).cuda_stream)
    _lib.check(rc)
    return bit_width * int(count.item())


def set_tr_tracking(model, tracking):
    """Switch every TR layer between histogram tracking and quantised inference
    (tr_layer.py:66-76); leaving tracking runs the calibration sweep."""
    for layer in list(model.modules()):
        if isinstance(layer, (TRLinearLayer, TRLSTMLayer, TRConv2dLayer)):
            layer.tracking(tracking)
    return model


class LinearQuantize(nn.Module):
    """Activation quantiser (tr_layer.py:78-104).  While ``tracking`` it accumulates an
    8192-bin histogram over [-50, 50] and passes the input through; afterwards it
    term-reveals the flattened activation with g = 1 and ``data_terms`` terms per value."""

    def __init__(self, data_bits, data_terms):
        super().__init__()
        self.sf = 1
        self.num_bins = 8192
        self.minv = -50
        self.maxv = 50
        self.register_buffer("hist_bins", torch.zeros(self.num_bins))
        self.tracking = True
        self.data_bits = data_bits
        self.data_terms = data_terms
        self._scratch = None

    def _track(self, x):
        xc = x.detach().contiguous()
        if self._scratch is None or self._scratch.device != xc.device:
            self._scratch = torch.zeros(self.num_bins, dtype=torch.int32, device=xc.device)
        dt = tr_cuda._DTYPES[xc.dtype]
        with torch.cuda.device(xc.device):
            rc = _lib.lib().tq_hist_accumulate(
                xc.data_ptr(), dt, xc.numel(), self.hist_bins.data_ptr(), self._scratch.data_ptr(),
                self.num_bins, float(self.minv), float(self.maxv),
                torch.cuda.current_stream(xc.device).cuda_stream)
        _lib.check(rc)

    def forward(self, x):
        if self.tracking:
            if not x.is_cuda:
                raise RuntimeError("input must be a CUDA tensor")
            self._track(x)
            return x
        dims = x.shape
        flat = x.contiguous().view(1, -1, 1, 1)          # tr_layer.py:97
        flat = tr_cuda.tr(flat, self.sf, self.data_bits, 1, self.data_terms)
        return flat.view(*dims)

    def finish_tracking(self):
        self.sf = mse_profile(self.hist_bins, self.minv, self.maxv, self.data_bits, self.data_terms)
        self.tracking = False


class _TRBase(nn.Module):
    def _setup(self, device, data_bits, data_terms, weight_bits, group_size, num_terms):
        self.data_bits = data_bits
        self.data_terms = data_terms
        self.input_quant = LinearQuantize(data_bits, data_terms).to(device)
        self.group_size = group_size
        self.num_terms = num_terms
        self.weight_bits = weight_bits

    def _reveal_weight(self, w):
        """w_sf = max|w| / 2^(bits-1) then tr over groups of input channels
        (tr_layer.py:117-121).  Sets self.w_sf, returns the new Parameter."""
        self.w_sf = w.abs().max().item() / 2 ** (self.weight_bits - 1)
        wq = tr_cuda.tr(w.detach().contiguous(), self.w_sf, self.weight_bits, self.group_size,
                        self.num_terms)
        return nn.Parameter(wq)

    def tracking(self, tracking):
        if not tracking:
            self.input_quant.finish_tracking()
        else:
            self.input_quant.tracking = True


class TRConv2dLayer(_TRBase):
    """Conv2d on term-revealed weights and activations (tr_layer.py:106-132)."""

    def __init__(self, conv_layer, data_bits=8, data_terms=4, weight_bits=8, group_size=1,
                 num_terms=8):
        super().__init__()
        self._setup(conv_layer.weight.device, data_bits, data_terms, weight_bits, group_size,
                    num_terms)
        conv_layer.weight = self._reveal_weight(conv_layer.weight)
        self.conv = conv_layer

    def forward(self, x):
        return self.conv(self.input_quant(x))


class TRLinearLayer(_TRBase):
    """Linear on term-revealed weights (tr_layer.py:134-160)."""

    def __init__(self, linear_layer, data_bits=8, data_terms=4, weight_bits=8, group_size=1,
                 num_terms=8):
        super().__init__()
        self._setup(linear_layer.weight.device, data_bits, data_terms, weight_bits, group_size,
                    num_terms)
        linear_layer.weight = self._reveal_weight(linear_layer.weight)
        self.linear = linear_layer

    def forward(self, x):
        if STRICT_REFERENCE:
            # tr_layer.py:152-154 quantises x and then ignores the result; only the histogram
            # side effect of tracking mode is observable, so that is all that is kept.
            if self.input_quant.tracking:
                self.input_quant(x)
            return self.linear(x)
        return self.linear(self.input_quant(x))


class TRLSTMLayer(_TRBase):
    """nn.LSTM whose layer-0 weights are term-revealed; emb, h0 and c0 share one input
    quantiser (tr_layer.py:162-201)."""

    def __init__(self, lstm_layer, data_bits=8, data_terms=4, weight_bits=8, group_size=1,
                 num_terms=8):
        super().__init__()
        self._setup(lstm_layer.weight_ih_l0.device, data_bits, data_terms, weight_bits, group_size,
                    num_terms)
        lstm_layer.weight_ih_l0 = self._reveal_weight(lstm_layer.weight_ih_l0)
        lstm_layer.weight_hh_l0 = self._reveal_weight(lstm_layer.weight_hh_l0)   # w_sf := hh's
        self.lstm = lstm_layer
        self.lstm.flatten_parameters()

    def forward(self, emb, hidden):
        embq = self.input_quant(emb)
        hidden_qs = tuple(self.input_quant(h) for h in hidden)
        return self.lstm(embq, hidden_qs)
