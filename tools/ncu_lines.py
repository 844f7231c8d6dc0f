#!/usr/bin/env python
"""Per-instruction view of an ncu source page: index, samples, executions, lanes, top stall reasons, SASS.
usage: python tools/ncu_lines.py <rep> <first> <last>    (instruction indices as in ncu_summary.py --blocks)"""
import csv, io, subprocess, sys
rep, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for k, r in enumerate(rows[2:]):
    if k < lo or k > hi or len(r) < len(hdr):
        continue
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    ss = " ".join(f"{n}:{v}" for v, n in st if v)
    print(f"{k:5d} s={r[ix['# Samples']]:>6s} x={int(r[ix['Instructions Executed']])/1e6:7.2f}M l={r[ix['Avg. Threads Executed']]:>4s} {r[ix['Source']].strip():60s} {ss}")
