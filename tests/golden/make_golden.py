"""Generates tests/golden/tr_golden.npz from the REFERENCE kernel body run on the host
(oracle/_ref/libtq_ref.so, built by oracle/Makefile from /root/reference).

Run only where /root/reference exists:  python tests/golden/make_golden.py
The .npz is committed; tests read it and never need /root/reference.

Cases cover: SURVEY section 4's hand table, the shapes of the reference's call sites
(tr_layer.py:97-98 activations as (1,N,1,1) with g=1; tr_layer.py:120 conv weights
(O,I,k,k) with g=8; tr_layer.py:148 linear weights (out,in)), all group sizes of
evaluate_group_size.py:75-88, float64, clipping, ties and all-zero groups.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import tq_oracle as O  # noqa: E402


def main():
    assert O.have_ref(), "build oracle/_ref first (make -C oracle)"
    rng = np.random.default_rng(20261018)
    cases = []

    def add(x, sf, bits, g, alpha):
        x = np.ascontiguousarray(x)
        y = O.ref_tr(x, sf, bits, g, alpha)
        cases.append((x, np.float32(sf), bits, g, alpha, y))

    # SURVEY section 4 table
    v = np.array([[127, -127, 3, -6, 11, 0, 27, 255]], dtype=np.float32)
    for a in (1, 2, 4, 8, 12, 16):
        add(v, 1.0, 8, 8, a)
    v1 = v.reshape(1, 8, 1, 1)
    add(v1, 1.0, 8, 1, 1)
    add(v1, 1.0, 8, 1, 2)
    add(v1, 1.0, 7, 1, 16)
    add(np.array([[3.2, 0.15, 0.7]], dtype=np.float32), 0.05, 8, 3, 4)
    add(np.array([[0.49999997, 0.5, 1.5, 2.5]], dtype=np.float32).reshape(1, 4, 1, 1), 1.0, 8, 1, 8)
    add(np.array([[4, 4, 4, 4]], dtype=np.float32), 1.0, 8, 4, 2)
    add(np.zeros((2, 16), dtype=np.float32), 0.5, 8, 8, 12)

    # activations, tr_layer.py:97-98: (1, N, 1, 1), g = 1
    for bits, terms in ((8, 4), (9, 3), (9, 2), (6, 6), (8, 1), (16, 16), (8, 8)):
        x = np.maximum(rng.standard_normal((1, 4099, 1, 1)), 0).astype(np.float32)
        add(x, float(x.max()) / 2 ** bits, bits, 1, terms)
        xs = (rng.standard_normal((1, 2053, 1, 1)) * 3).astype(np.float32)
        add(xs, float(np.abs(xs).max()) / 2 ** (bits - 1), bits, 1, terms)

    # conv weights, tr_layer.py:117-120: (O, I, k, k), groups along I
    for (o, i, k), bits, g, alpha in (((8, 64, 3), 8, 8, 12), ((4, 32, 3), 9, 8, 12),
                                      ((16, 64, 1), 8, 8, 16), ((3, 32, 3), 8, 16, 24),
                                      ((2, 64, 3), 8, 32, 32), ((4, 16, 5), 4, 16, 12),
                                      ((5, 8, 3), 8, 2, 3), ((5, 8, 3), 8, 4, 5),
                                      ((2, 8, 7), 8, 8, 0), ((2, 8, 1), 8, 8, 200)):
        w = (rng.standard_normal((o, i, k, k)) * np.sqrt(2.0 / (i * k * k))).astype(np.float32)
        add(w, float(np.abs(w).max()) / 2 ** (bits - 1), bits, g, alpha)

    # linear weights, tr_layer.py:145-148: (out, in)
    for (o, i), bits, g, alpha in (((10, 512), 8, 8, 12), ((7, 784), 4, 16, 12),
                                   ((6, 648), 8, 8, 8), ((5, 96), 5, 3, 4)):
        w = rng.uniform(-0.1, 0.1, size=(o, i)).astype(np.float32)
        add(w, float(np.abs(w).max()) / 2 ** (bits - 1), bits, g, alpha)

    # float64 dispatch (kernels/tr_cuda_kernel.cu:146)
    wd = rng.standard_normal((3, 16, 3, 3))
    add(wd, float(np.abs(wd).max()) / 128, 8, 8, 12)
    add(np.maximum(rng.standard_normal((1, 1001, 1, 1)), 0), 0.013, 8, 1, 3)

    # tr_layer.hese (tr_layer.py:9-41): the module cannot be imported (it JIT-builds the CUDA
    # extension at import time), so only the text of that one function is executed here.
    import re
    src = open("/root/reference/tr_layer.py").read()
    fn_src = re.search(r"^def hese\(number\):.*?^    return keep_exponents\n", src, re.S | re.M).group(0)
    ns = {}
    exec(fn_src, ns)
    hese_in = np.arange(-700, 701, dtype=np.int64)
    hese_flat, hese_off = [], [0]
    for v in hese_in:
        hese_flat += ns["hese"](int(v))
        hese_off.append(len(hese_flat))

    out = {"n": np.int64(len(cases)), "hese_in": hese_in, "hese_flat": np.array(hese_flat, dtype=np.int64),
           "hese_off": np.array(hese_off, dtype=np.int64)}
    for n, (x, sf, bits, g, alpha, y) in enumerate(cases):
        out[f"x{n}"] = x
        out[f"y{n}"] = y
        out[f"p{n}"] = np.array([bits, g, alpha], dtype=np.int64)
        out[f"sf{n}"] = sf
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tr_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(cases), "cases", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
