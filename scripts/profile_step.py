#!/usr/bin/env python
"""One small pass of the BASELINE config-2 hot path, for ncu (launch list / --set full captures).
Usage: python scripts/profile_step.py [rays] [sessions]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
from ice_halo_sim_b200 import backend as B  # noqa: E402

rays = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
sessions = int(sys.argv[2]) if len(sys.argv) > 2 else 2
case = parity.CASES[sys.argv[3]] if len(sys.argv) > 3 else parity.CASES["column_config2"]
tables = B.SceneTables(case["scene"](), 7)
be = B.B200TraceBackend(0)
be.SetScene(tables)
be.SetRender(case["render"]())
wl = [B.make_wl_entry(550.0, 1.0)]
for s in range(sessions):
    be.BeginSession(B.SessionSpec(seed=42, wl=wl, ray_num=rays, accumulate=True))
    be.TraceLayer(B.RootRaySource.FromHost(rays), want_stats=False)
    be.EndSession()
be.Synchronize()
img, landed = be.ReadbackXyzAccum()
print("landed", landed, "sum", float(img.sum()))
be.close()
