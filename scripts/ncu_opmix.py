#!/usr/bin/env python
"""Opcode mix (executed warp instructions per SASS opcode) per kernel from an .ncu-rep with source info.
Usage: python scripts/ncu_opmix.py prof.ncu-rep [rays_per_launch]"""
import csv, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
mix = defaultdict(lambda: defaultdict(int)); tot = defaultdict(int); launches = defaultdict(int)
fn = None
for row in csv.reader(raw.splitlines()):
    if not row: continue
    if row[0] == "Kernel Name":
        fn = row[1]; launches[fn] += 1; continue
    if row[0] == "Address" or fn is None: continue
    try: inst = int(row[5])
    except (ValueError, IndexError): continue
    op = row[1].strip().split()
    if op and op[0].startswith("@"): op = op[1:]
    o = op[0].split(".")[0] if op else "?"
    mix[fn][o] += inst; tot[fn] += inst
for fn in mix:
    print("=" * 10, fn[:90], "launches", launches[fn], "warp-inst", tot[fn])
    for o, c in sorted(mix[fn].items(), key=lambda kv: -kv[1])[:28]:
        print(f"  {o:12s} {100.0*c/tot[fn]:5.1f}%  {c}")
