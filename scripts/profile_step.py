#!/usr/bin/env python
"""A few sessions of one scene's hot path, for ncu (launch list / --set full captures).
Usage: python scripts/profile_step.py [rays] [sessions] [case] [key=value engine options ...]
Default case: BASELINE config 2 (column_config2); all layers of the scene are traced (driver.trace_session)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ice_halo_sim_b200 import backend as B  # noqa: E402
from ice_halo_sim_b200 import driver, scenes  # noqa: E402

rays = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
sessions = int(sys.argv[2]) if len(sys.argv) > 2 else 2
case = scenes.CASES[sys.argv[3]] if len(sys.argv) > 3 else scenes.CASES["column_config2"]
desc = case["scene"]()
tables = B.SceneTables(desc, 7)
be = B.B200TraceBackend(0)
for kv in sys.argv[4:]:
    k, v = kv.split("=")
    be.SetOption(k, int(v))
be.SetScene(tables)
be.SetRender(case["render"]())
wl = [B.make_wl_entry(550.0, 1.0)]
for s in range(sessions):
    driver.trace_session(be, int(desc.layer_cnt), B.SessionSpec(seed=42, wl=wl, ray_num=rays, accumulate=True), rays)
be.Synchronize()
img, landed = be.ReadbackXyzAccum()
print("landed", landed, "sum", float(img.sum()))
be.close()
