"""``profile(model, inputs, custom_ops)`` with the semantics of thop/profile.py:59-128: hooks
go on leaf modules and on any module whose type is in ``custom_ops``; one forward in eval
mode; totals are summed from 1-element float32 buffers in ``model.modules()`` order (so the
result is float32-rounded like the published results/*.json)."""
import logging

import torch
import torch.nn as nn

from .count_hooks import (count_adap_avgpool, count_avgpool, count_bn, count_convNd,
                          count_linear, count_relu, count_upsample, zero_ops)

logger = logging.getLogger(__name__)

register_hooks = {
    nn.Conv1d: count_convNd, nn.Conv2d: count_convNd, nn.Conv3d: count_convNd,
    nn.ConvTranspose1d: count_convNd, nn.ConvTranspose2d: count_convNd, nn.ConvTranspose3d: count_convNd,
    nn.BatchNorm1d: count_bn, nn.BatchNorm2d: count_bn, nn.BatchNorm3d: count_bn,
    nn.ReLU: zero_ops, nn.ReLU6: zero_ops, nn.LeakyReLU: count_relu,
    nn.MaxPool1d: zero_ops, nn.MaxPool2d: zero_ops, nn.MaxPool3d: zero_ops,
    nn.AdaptiveMaxPool1d: zero_ops, nn.AdaptiveMaxPool2d: zero_ops, nn.AdaptiveMaxPool3d: zero_ops,
    nn.AvgPool1d: count_avgpool, nn.AvgPool2d: count_avgpool, nn.AvgPool3d: count_avgpool,
    nn.AdaptiveAvgPool1d: count_adap_avgpool, nn.AdaptiveAvgPool2d: count_adap_avgpool,
    nn.AdaptiveAvgPool3d: count_adap_avgpool,
    nn.Linear: count_linear, nn.Dropout: zero_ops,
    nn.Upsample: count_upsample, nn.UpsamplingBilinear2d: count_upsample,
    nn.UpsamplingNearest2d: count_upsample,
}


def _counted(m, custom_ops):
    return len(list(m.children())) == 0 or type(m) in custom_ops


def profile(model, inputs, custom_ops=None, verbose=True):
    custom_ops = custom_ops or {}
    handles = []

    def add_hooks(m):
        if not _counted(m, custom_ops):
            return
        m.register_buffer('total_ops', torch.zeros(1))
        m.register_buffer('total_params', torch.zeros(1))
        fn = custom_ops.get(type(m), register_hooks.get(type(m)))
        if fn is None:
            if verbose:
                logger.info("no counting rule for %s", type(m).__name__)
            return
        handles.append(m.register_forward_hook(fn))

    training = model.training
    model.eval()
    model.apply(add_hooks)
    try:
        with torch.no_grad():
            model(*inputs)
        total_ops = 0
        total_params = 0
        for m in model.modules():
            if not _counted(m, custom_ops):
                continue
            total_ops += m.total_ops
            total_params += m.total_params
        total_ops = total_ops.item()
        total_params = total_params.item()
    finally:
        model.train(training)
        for h in handles:
            h.remove()
        for m in model.modules():
            m._buffers.pop("total_ops", None)
            m._buffers.pop("total_params", None)
    return total_ops, total_params
