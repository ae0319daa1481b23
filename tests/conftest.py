import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_cases():
    """(x, sf, bits, g, alpha, y_ref) tuples generated from the reference kernel body by
    tests/golden/make_golden.py."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "tr_golden.npz"))
    out = []
    for n in range(int(z["n"])):
        bits, g, alpha = (int(v) for v in z[f"p{n}"])
        out.append((z[f"x{n}"], float(z[f"sf{n}"]), bits, g, alpha, z[f"y{n}"]))
    return out


def bits_equal(a, b):
    """Bit-for-bit equality of two float arrays (distinguishes -0.0 from +0.0)."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    u = {2: np.uint16, 4: np.uint32, 8: np.uint64}[a.dtype.itemsize]
    return bool(np.array_equal(a.view(u), b.view(u)))
