"""Times the REFERENCE kernel (kernels/tr_cuda_kernel.cu:58-125, recompiled for sm_100a: oracle/_ref/libtq_ref_gpu.so)
next to this repo's kernels on the same tensors and the same B200 -- "the kernel to beat".  Lives under tests/ because
only tests/, smoke() and bench.py's CPU-baseline leg may execute anything under oracle/ (it is not collected by pytest).

    python tests/time_reference_kernel.py [--sizes 51380224 268435456]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import tq_oracle as O  # noqa: E402
from term_quantization_b200 import tr_cuda  # noqa: E402


def time_call(fn, iters, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", type=int, nargs="+", default=[51380224, 1 << 28])
    args = ap.parse_args()
    assert O.have_ref_gpu(), "build oracle/_ref/libtq_ref_gpu.so first (make -C oracle, needs /root/reference)"
    torch.manual_seed(0)
    for n in args.sizes:
        nbuf = max(2, min(8, (1 << 30) // (4 * n) + 1))
        xs = [torch.relu(torch.randn(1, n, 1, 1, device="cuda")) for _ in range(nbuf)]
        out = torch.empty_like(xs[0])
        for label, shape_fn, bits, g, alpha in (("g=1 k=3 b=9", lambda t: t, 9, 1, 3),
                                                ("g=8 a=12 b=8 contiguous", lambda t: t.view(-1, 512), 8, 8, 12)):
            views = [shape_fn(t) for t in xs]
            sf = float(xs[0].max()) / 2 ** bits
            o = out.view(views[0].shape)
            state = {"i": 0}

            def run(fn):
                v = views[state["i"] % nbuf]
                state["i"] += 1
                fn(v)
            t_ref = time_call(lambda: run(lambda v: O.ref_gpu_tr(v, sf, bits, g, alpha, out=o)), 5, 2)
            t_new = time_call(lambda: run(lambda v: tr_cuda.tr(v, sf, bits, g, alpha, out=o)), 20, 5)
            print(json.dumps({"case": label, "n": n, "reference_kernel_GBs": n * 8 / t_ref / 1e6,
                              "this_repo_GBs": n * 8 / t_new / 1e6, "speedup": t_ref / t_new}))


if __name__ == "__main__":
    main()
