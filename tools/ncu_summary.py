"""Summarise an .ncu-rep captured with `ncu --set full` (read here, on the CPU box, with `ncu -i ... --page raw --csv`):
one row per launch with the metrics the roofline discussion uses, as a text table for profiles/.

    python tools/ncu_summary.py gpurun_out/r2_prof_conv.ncu-rep | gpurun_out/r2_conv_raw.csv [--traffic profiles/r02_conv_traffic.json]

--traffic additionally writes the mean DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) with the
sha256 of csrc/tq_gemm.cu it was measured on, which is what bench.py's roofline.traffic reads (and refuses when stale)."""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3,
         "nsecond": 1e-3, "second": 1e6}


def main():
    rep = sys.argv[1]
    if rep.endswith(".csv"):                      # `ncu -i x.ncu-rep --page raw --csv` already run (on the GPU box)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    names = [m for m, _ in METRICS if m in col]
    print("# " + os.path.basename(rep) + ": ncu --set full --clock-control none (cold-cache, serialised launches: compare shares, not absolutes)")
    print(" | ".join(["#", "kernel", "grid"] + [f"{short} [{units[col[m]]}]" if units[col[m]] else short for m, short in METRICS if m in col]))
    tot_bytes, n = 0.0, 0
    for i, r in enumerate(data):
        kname = r[col["Kernel Name"]]
        kname = kname[:kname.index("(")] if "(" in kname else kname
        vals = [r[col[m]] for m in names]
        print(" | ".join([str(i), kname.replace("void ", ""), r[col.get("Grid Size", 0)]] + vals))
        try:
            rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * SCALE.get(units[col["dram__bytes_read.sum"]], 1.0)
            wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * SCALE.get(units[col["dram__bytes_write.sum"]], 1.0)
            tot_bytes += rd + wr
            n += 1
        except (KeyError, ValueError):
            pass
    if "--traffic" in sys.argv and n:
        out = sys.argv[sys.argv.index("--traffic") + 1]
        src = os.path.join(ROOT, "term_quantization_b200", "csrc", "tq_gemm.cu")
        rec = {"dram_bytes_per_launch": tot_bytes / n, "launches": n, "tq_gemm_cu_sha256": hashlib.sha256(open(src, "rb").read()).hexdigest(),
               "how": f"ncu --set full --clock-control none, mean over the {n} conv launches of one forward ({os.path.basename(rep)}; "
                      "cold L2 per launch under ncu, so this is an upper bound of the in-step traffic)"}
        json.dump(rec, open(out, "w"), indent=1)
        print(f"# wrote {out}: {tot_bytes / n / 1e6:.1f} MB per launch over {n} launches")


if __name__ == "__main__":
    main()
