#!/usr/bin/env python
"""Share of device time per kernel from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: python scripts/launch_shares.py profiles/r1g_launches_bench_step.csv"""
import csv
import sys
from collections import defaultdict

tot = defaultdict(lambda: [0, 0.0])
with open(sys.argv[1]) as f:
    rows = [r for r in csv.reader(l for l in f if l.startswith('"'))]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
for r in rows[1:]:
    t = tot[r[ki]]
    t[0] += 1
    t[1] += float(r[vi]) / 1e6
total = sum(v[1] for v in tot.values())
print("| kernel | launches | total ms | share | avg us / launch |\n|---|---|---|---|---|")
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {ms:.2f} | {ms / total:.3f} | {1e3 * ms / n:.1f} |")
