#!/usr/bin/env python
"""Per-instruction view of an ncu source page: index, stall samples, executions, active lanes, top stall
reasons, SASS.
usage: python tools/ncu_lines.py <rep> [--kernel K] (<first> <last> | --min-exec N)
  --kernel K      the K-th launch of the report (default 0)
  <first> <last>  instruction indices as printed by ncu_summary.py --blocks
  --min-exec N    instead: every instruction executed at least N times (the hot loop)"""
import csv
import io
import subprocess
import sys

args = sys.argv[1:]
rep = args.pop(0)
k = 0
if "--kernel" in args:
    i = args.index("--kernel")
    k = int(args[i + 1])
    del args[i:i + 2]
min_exec = None
if "--min-exec" in args:
    i = args.index("--min-exec")
    min_exec = float(args[i + 1])
    del args[i:i + 2]
lo, hi = (int(args[0]), int(args[1])) if len(args) >= 2 else (0, 10 ** 9)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(k), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print("#", rows[0][1] if len(rows[0]) > 1 else "")
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
seen = set()
n = -1
for r in rows[2:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    if r[0] in seen:
        break
    seen.add(r[0])
    n += 1
    x = int(r[ix["Instructions Executed"]])
    if n < lo or n > hi or (min_exec is not None and x < min_exec):
        continue
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    ss = " ".join(f"{name}:{v}" for v, name in st if v)
    print(f"{n:5d} s={r[ix['# Samples']]:>6s} x={x / 1e6:7.2f}M l={r[ix['Avg. Threads Executed']]:>4s} "
          f"{r[ix['Source']].strip():60s} {ss}")
