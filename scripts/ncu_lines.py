#!/usr/bin/env python
"""Per-source-line instruction counts from an .ncu-rep captured with --import-source on.
Usage: python scripts/ncu_lines.py prof.ncu-rep [top_n]  -> per kernel: lines sorted by warp instructions executed."""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
per = defaultdict(lambda: defaultdict(lambda: [0, 0, 0, ""]))  # kernel -> (file, line) -> [inst, thread_inst, samples, src]
totals = defaultdict(int)
fpath = fn = None
order = []
for row in csv.reader(raw.splitlines()):
    if not row:
        continue
    if row[0] == "File Path":
        fpath = row[1].split("/")[-1]
        continue
    if row[0] == "Function Name":
        fn = row[1]
        if fn not in order:
            order.append(fn)
        continue
    if row[0] == "Line No" or fn is None:
        continue
    if row[0] != "" and row[0].isdigit():  # aggregated source line
        try:
            inst, tinst, samp = int(row[7]), int(row[8]), int(row[6])
        except ValueError:
            continue
        e = per[fn][(fpath, int(row[0]))]
        e[0] += inst
        e[1] += tinst
        e[2] += samp
        e[3] = row[1].strip()[:110]
        totals[fn] += inst
for fn in order:
    print("=" * 20, fn[:100], "total warp-inst", totals[fn])
    items = sorted(per[fn].items(), key=lambda kv: -kv[1][0])[:top]
    for (f, ln), (inst, tinst, samp, src) in items:
        if inst == 0:
            continue
        print(f"{100.0 * inst / max(1, totals[fn]):5.1f}%  act={tinst / max(1, inst):4.1f}  smp={samp:5d}  {f}:{ln}  {src}")
