#!/usr/bin/env python
"""Registers / stack / spill bytes per kernel from the build's ptxas log.
Usage: python scripts/ptxas_table.py [ptxas.log] [--all]  (default: kernels that spill + the production fast kernels)"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = [a for a in sys.argv[1:] if not a.startswith("--")]
log = args[0] if args else os.path.join(ROOT, "ice_halo_sim_b200", "csrc", "ptxas.log")
show_all = "--all" in sys.argv
cur, st, rows = None, (0, 0, 0), []
for line in open(log):
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = m.group(1)
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and cur:
        st = tuple(int(v) for v in m.groups())
    m = re.search(r"Used (\d+) registers", line)
    if m and cur:
        rows.append((cur, int(m.group(1)), st))
        cur = None
names = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
print("regs  (stack, spill st, spill ld)  kernel")
for (_, regs, st), d in zip(rows, names):
    d = re.sub(r"\(.*\)$", "", d.replace("hb::", "").replace("(bool)", "").replace("(int)", "").replace("void ", ""))
    if show_all or st[1] or st[2] or d.startswith("bounce_kernel<0") or d.startswith("gen_kernel"):
        print(f"{regs:4d}  {st}  {d}")
