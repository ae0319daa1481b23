"""The TR-encode grid of the bench line on its own (bench_extra.tr_grid): GB/s per case."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_extra  # noqa: E402

peak = 6537.3
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
rows, _ = bench_extra.tr_grid(torch.device("cuda:0"), peak, with_ref=False)
for r in rows:
    print(f"{r['case']:90s} n={r['elements']:>10d}  {r['GBs_best']:8.1f} GB/s best  {r['GBs_median']:8.1f} median  {r['frac_of_measured']:.3f} of measured")
