"""Per-module counting hooks (subset of thop/count_hooks.py that the TQ drivers reach).
Every hook adds into the module's 1-element float32 ``total_ops`` buffer, exactly like the
reference (thop/profile.py:72-73), so large counts round the same way."""
import torch

multiply_adds = 1


def _add(m, value):
    m.total_ops += torch.Tensor([int(value)])


def zero_ops(m, x, y):
    _add(m, 0)


def count_convNd(m, x, y):
    kernel_ops = 1
    for k in m.weight.shape[2:]:
        kernel_ops *= k
    bias_ops = 1 if m.bias is not None else 0
    _add(m, y.nelement() * (m.in_channels // m.groups * kernel_ops + bias_ops))


def count_bn(m, x, y):
    if not m.training:
        _add(m, 2 * x[0].numel())


def count_relu(m, x, y):
    _add(m, x[0].numel())


def count_softmax(m, x, y):
    batch, nfeatures = x[0].shape[0], x[0].shape[1]
    _add(m, batch * (nfeatures + (nfeatures - 1) + nfeatures))


def count_avgpool(m, x, y):
    _add(m, y.numel())


def count_adap_avgpool(m, x, y):
    kernel = torch.div(torch.Tensor([*x[0].shape[2:]]), torch.Tensor(list((m.output_size,))).squeeze(),
                       rounding_mode="floor")
    _add(m, (torch.prod(kernel) + 1) * y.numel())


def count_upsample(m, x, y):
    _add(m, {"nearest": 0, "linear": 5, "bilinear": 11, "bicubic": 259, "trilinear": 31}.get(m.mode, 0) * y.nelement())


def count_linear(m, x, y):
    _add(m, m.in_features * y.numel())
