"""TR-encode micro-benchmark: GB/s of tq_tr_encode at the BASELINE sizes.
Times with CUDA events on the launching stream; rotates buffers larger than L2."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from term_quantization_b200 import tr_cuda  # noqa: E402


def time_call(fn, iters=20, warm=5):
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
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", type=int, nargs="+", default=[1 << 20, 1 << 24, 51380224, 1 << 28])
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    torch.manual_seed(0)
    rows = []
    for n in args.sizes:
        nbuf = max(2, min(8, (1 << 30) // (4 * n) + 1))       # rotate > 126 MB of L2
        xs = [torch.relu(torch.randn(1, n, 1, 1, device="cuda")) for _ in range(nbuf)]
        out = torch.empty_like(xs[0])
        for label, shape_fn, bits, g, alpha, bytes_per in (
                ("act g=1 k=3 b=9", lambda t: t, 9, 1, 3, 8),
                ("act g=1 k=4 b=8", lambda t: t, 8, 1, 4, 8),
                ("act g=1 k=16 b=16", lambda t: t, 16, 1, 16, 8),
                ("wgt g=8 a=12 b=8 contiguous", lambda t: t.view(-1, 512), 8, 8, 12, 8),
                ("wgt g=8 a=12 b=8 stride9", lambda t: t.view(-1, 64, 3, 3) if n % 576 == 0 else None, 8, 8, 12, 8)):
            views = [shape_fn(t) for t in xs]
            if views[0] is None:
                continue
            sf = float(xs[0].max()) / 2 ** bits
            o = out.view(views[0].shape)
            state = {"i": 0}

            def fn():
                v = views[state["i"] % nbuf]
                state["i"] += 1
                tr_cuda.tr(v, sf, bits, g, alpha, out=o)
            best, med = time_call(fn)
            rows.append({"case": label, "n": n, "ms_best": best, "ms_med": med,
                         "GBs_best": n * bytes_per / best / 1e6, "GBs_med": n * bytes_per / med / 1e6})
            print(json.dumps(rows[-1]))
        # copy baseline for context
        best, med = time_call(lambda: out.copy_(xs[0]))
        rows.append({"case": "torch copy_", "n": n, "ms_best": best, "GBs_best": n * 8 / best / 1e6})
        print(json.dumps(rows[-1]))
        del xs, out
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
