"""The reference's module stack on the CPU (TEST INFRASTRUCTURE ONLY).

The reference has no CPU implementation of its TR op (kernels/tr_cuda.cpp rejects CPU tensors),
so "the reference's CPU path" is assembled here exactly as SURVEY 8(d) prescribes: the
reference kernel body compiled for the host (oracle/_ref/libtq_ref.so, threaded over all host
cores) -- or our restatement when _ref was not shipped -- behind wrappers with the
semantics of tr_layer.py:78-160, feeding PyTorch's CPU fp32 conv/linear.

Used by: tests (logit parity of the CUDA path), bench.py's cpu_baseline leg and
`bench.py --impl reference`.  Never imported by the product package.
"""
from copy import deepcopy

import numpy as np
import torch
import torch.nn as nn

from . import tq_oracle as O


def cpu_tr(x, sf, bits, g, alpha, threads=0):
    """tr_cuda.tr on a CPU tensor via the reference kernel body (or the restatement)."""
    a = x.detach().contiguous().numpy()
    shape = a.shape
    if a.ndim not in (2, 4):
        a = a.reshape(shape[0], shape[1], -1, 1)
    C = a.shape[1]
    if O.have_ref() and C % g == 0:
        y = O.ref_tr(a, sf, bits, g, alpha, threads=threads)
    else:
        y = O.tr(a, sf, bits, g, alpha)
    return torch.from_numpy(y.reshape(shape))


class RefLinearQuantize(nn.Module):
    """tr_layer.py:78-104 on the CPU (histogram tracking via the oracle's histc restatement)."""

    def __init__(self, data_bits, data_terms):
        super().__init__()
        self.sf = 1
        self.num_bins, self.minv, self.maxv = 8192, -50, 50
        self.register_buffer("hist_bins", torch.zeros(self.num_bins))
        self.tracking = True
        self.data_bits, self.data_terms = data_bits, data_terms

    def forward(self, x):
        if self.tracking:
            O.hist(x.detach().contiguous().numpy(), self.hist_bins.numpy(), self.minv, self.maxv)
            return x
        dims = x.shape
        y = cpu_tr(x.contiguous().view(1, -1, 1, 1), self.sf, self.data_bits, 1, self.data_terms)
        return y.view(*dims)

    def finish_tracking(self):
        grid = torch.linspace(self.minv, self.maxv, self.num_bins).numpy()
        sfs = torch.linspace(1e-8, self.maxv, 2048)
        idx, _ = O.mse_profile(self.hist_bins.numpy(), grid, sfs.numpy(), self.data_bits, self.data_terms)
        self.sf = sfs.tolist()[idx]
        self.tracking = False


class RefTRConv2d(nn.Module):
    """tr_layer.py:106-132 on the CPU."""

    def __init__(self, conv, data_bits=8, data_terms=4, weight_bits=8, group_size=1, num_terms=8):
        super().__init__()
        self.input_quant = RefLinearQuantize(data_bits, data_terms)
        self.data_bits, self.data_terms = data_bits, data_terms
        self.group_size, self.num_terms, self.weight_bits = group_size, num_terms, weight_bits
        w = conv.weight
        self.w_sf = w.abs().max().item() / 2 ** (weight_bits - 1)
        conv.weight = nn.Parameter(cpu_tr(w, self.w_sf, weight_bits, group_size, num_terms))
        self.conv = conv

    def forward(self, x):
        return self.conv(self.input_quant(x))


class RefTRLinear(nn.Module):
    """tr_layer.py:134-160 on the CPU (forward uses the unquantised input, :152-154)."""

    def __init__(self, linear, data_bits=8, data_terms=4, weight_bits=8, group_size=1, num_terms=8):
        super().__init__()
        self.input_quant = RefLinearQuantize(data_bits, data_terms)
        self.data_bits, self.data_terms = data_bits, data_terms
        self.group_size, self.num_terms, self.weight_bits = group_size, num_terms, weight_bits
        w = linear.weight
        self.w_sf = w.abs().max().item() / 2 ** (weight_bits - 1)
        linear.weight = nn.Parameter(cpu_tr(w, self.w_sf, weight_bits, group_size, num_terms))
        self.linear = linear

    def forward(self, x):
        self.input_quant(x)
        return self.linear(x)


def convert_cnn(model, weight_bits, group_size, weight_terms, data_bits, data_terms):
    """cnn_models.convert_model with static_conv_layer_settings (cnn_models/__init__.py:30-70):
    first conv untouched, grouped / 'se' convs at (16, 1, 16)."""
    model = deepcopy(model)
    convs = [(n, m) for n, m in model.named_modules() if isinstance(m, nn.Conv2d)]
    for idx, (name, layer) in enumerate(convs):
        if idx == 0:
            continue
        wb, gs, wt = (16, 1, 16) if (layer.groups > 1 or 'se' in name) else (weight_bits, group_size, weight_terms)
        parent = model
        keys = name.split('.')
        for k in keys[:-1]:
            parent = parent._modules[k]
        parent._modules[keys[-1]] = RefTRConv2d(layer, data_bits, data_terms, wb, gs, wt)
    return model


def quantizers(model):
    return [m for m in model.modules() if isinstance(m, RefLinearQuantize)]


def set_scale_factors(model, sfs):
    """Inject fixed activation scale factors (skips the 2048-step sweep) and leave tracking."""
    qs = quantizers(model)
    assert len(qs) == len(sfs)
    for q, sf in zip(qs, sfs):
        q.sf = float(sf)
        q.tracking = False
    return model
