"""Top SASS instructions of one launch by stall samples and by executed instructions (ncu --page source --csv export)."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
head = rows[1]
col = {h: i for i, h in enumerate(head)}
data = [r for r in rows[2:] if len(r) == len(head) and r[0] != 'Address']
tot_s = sum(int(r[col['# Samples']] or 0) for r in data)
tot_i = sum(int(r[col['Instructions Executed']] or 0) for r in data)
print(rows[0][1][:90], "| samples", tot_s, "| warp instructions", tot_i)
ops = Counter()
opsamp = Counter()
for r in data:
    op = r[col['Source']].split()[0] if r[col['Source']].split() else '?'
    if op.startswith('@'):
        op = r[col['Source']].split()[1]
    op = op.split('.')[0]
    ops[op] += int(r[col['Instructions Executed']] or 0)
    opsamp[op] += int(r[col['# Samples']] or 0)
print("by opcode (warp instr %, stall-sample %):")
for op, n in ops.most_common(22):
    print(f"  {op:10s} {100 * n / tot_i:5.1f}%  {100 * opsamp[op] / tot_s:5.1f}%")
stalls = [h for h in head if h.startswith('stall_') and 'Not Issued' not in h]
agg = Counter()
for r in data:
    for h in stalls:
        agg[h] += int(r[col[h]] or 0)
print("stall reasons (% of samples):", ", ".join(f"{h[6:]} {100 * v / tot_s:.1f}" for h, v in agg.most_common(8)))
print("top instructions by samples:")
for r in sorted(data, key=lambda r: -int(r[col['# Samples']] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    top = max(stalls, key=lambda h: int(r[col[h]] or 0))
    print(f"  {100 * int(r[col['# Samples']] or 0) / tot_s:5.1f}%  exec {int(r[col['Instructions Executed']] or 0):9d}  {top[6:]:12s} {r[col['Source']][:90]}")
