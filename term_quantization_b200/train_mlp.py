"""MNIST MLP definition and test loop used by evaluate_mlp (mirror of train_mlp.py:10-26,44-64;
the training loop of the reference is out of scope)."""
import torch.nn as nn
import torch.nn.functional as F


class MNISTMLP(nn.Module):
    """784-512-512-10 with ReLU + Dropout(0.2) between layers, log-softmax output."""

    def __init__(self):
        super().__init__()
        layers = []
        for i, (fan_in, fan_out) in enumerate(((784, 512), (512, 512), (512, 10))):
            layers.append(nn.Linear(fan_in, fan_out))
            if i < 2:
                layers += [nn.ReLU(), nn.Dropout(0.2)]
        self.features = nn.Sequential(*layers)

    def forward(self, x):
        return F.log_softmax(self.features(x.flatten(1)), dim=1)


def test(args, model, device, test_loader, pct=1.0):
    """Accuracy over the first `pct` of the loader's dataset (stops early like the reference,
    but still divides by the full dataset size, train_mlp.py:62-64)."""
    import torch
    model.eval()
    n_total = len(test_loader.dataset.targets)
    eval_samples = round(pct * n_total)
    seen = correct = 0
    with torch.no_grad():
        for data, target in test_loader:
            seen += len(target)
            data, target = data.to(device), target.to(device)
            pred = model(data).argmax(dim=1)
            correct += int((pred == target).sum().item())
            if seen >= eval_samples:
                break
    return correct / len(test_loader.dataset)
