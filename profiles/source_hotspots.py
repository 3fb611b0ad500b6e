#!/usr/bin/env python
"""Per-instruction view of an ncu source page: executed counts, active lanes and stall samples.
usage: ncu -i X.ncu-rep --page source --csv > src.csv; python profiles/source_hotspots.py src.csv [lo_addr hi_addr]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
data = rows[2:]
tot_inst = sum(float(r[col["Instructions Executed"]] or 0) for r in data)
tot_samp = sum(float(r[col["# Samples"]] or 0) for r in data)
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 62
base = int(data[0][col["Address"]], 16)
stalls = ["stall_math", "stall_wait", "stall_not_selected", "stall_long_sb", "stall_short_sb", "stall_branch_resolving", "stall_dispatch", "stall_selected"]
print(f"total warp instructions {tot_inst:.4g}, samples {tot_samp:.0f}")
print(f"{'off':>6} {'inst%':>6} {'lanes':>5} {'samp%':>6}  " + " ".join(f"{s[6:10]:>5}" for s in stalls) + "  sass")
acc_i = acc_s = 0.0
for r in data:
    off = int(r[col["Address"]], 16) - base
    if not (lo <= off < hi):
        continue
    ie = float(r[col["Instructions Executed"]] or 0)
    sm = float(r[col["# Samples"]] or 0)
    acc_i += ie
    acc_s += sm
    lanes = float(r[col["Avg. Threads Executed"]] or 0)
    st = " ".join(f"{float(r[col[s]] or 0) / max(1.0, tot_samp) * 100:5.2f}" for s in stalls)
    print(f"{off:6x} {ie / tot_inst * 100:6.3f} {lanes:5.1f} {sm / tot_samp * 100:6.3f}  {st}  {r[col['Source']][:70]}")
print(f"range: {acc_i / tot_inst * 100:.2f} % of instructions, {acc_s / tot_samp * 100:.2f} % of samples")
