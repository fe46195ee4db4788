#!/usr/bin/env python
"""Per-kernel CUDA-event times of one traced session (engine "profile" mode), for A/B runs of engine options.
Usage: python scripts/kernel_times.py [--case column_config2] [--rays N] [--set k=v ...] [--ab key]
  --ab key  : run twice, with option key = 0 and key = 1, and print both lines."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
from ice_halo_sim_b200 import backend as B  # noqa: E402
from ice_halo_sim_b200.driver import trace_session  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--case", default="column_config2")
ap.add_argument("--rays", type=int, default=1 << 24)
ap.add_argument("--set", action="append", default=[])
ap.add_argument("--ab", default=None)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--pool", type=int, default=0)
args = ap.parse_args()

case = parity.CASES[args.case]
desc = case["scene"]()
if args.pool:
    desc.geom_pool_size = args.pool
tables = B.SceneTables(desc, 7)
be = B.B200TraceBackend(0)
be.SetScene(tables)
be.SetRender(case["render"]())
wl = [B.make_wl_entry(case["wl"][0], 1.0)]
for kv in args.set:
    k, v = kv.split("=")
    be.SetOption(k, int(v))


def measure(tag):
    trace_session(be, int(desc.layer_cnt), B.SessionSpec(seed=42, wl=wl, ray_num=args.rays, accumulate=True), args.rays)
    be.Synchronize()
    be.SetOption("profile", 1)
    c0 = be.Counters()
    for _ in range(args.reps):
        trace_session(be, int(desc.layer_cnt), B.SessionSpec(seed=42, wl=wl, ray_num=args.rays, accumulate=True), args.rays)
    be.Synchronize()
    c1 = be.Counters()
    be.SetOption("profile", 0)
    d = lambda f: getattr(c1, f) - getattr(c0, f)  # noqa: E731
    out = {"tag": tag, "case": args.case, "rays": args.rays,
           "optics_ms": d("optics_ms") / max(1, d("optics_launches")), "optics_launches": d("optics_launches") // args.reps,
           "intersect_ms": d("intersect_ms") / max(1, d("intersect_launches")),
           "gen_ms": d("gen_ms") / max(1, d("gen_launches")),
           "session_ms": (d("optics_ms") + d("intersect_ms") + d("gen_ms")) / args.reps}
    out["Mrays_s"] = args.rays / out["session_ms"] / 1e3
    img, landed = be.ReadbackXyzAccum()
    out["landed"] = float(landed)
    out["img_sum"] = float(img.astype("float64").sum())
    print(json.dumps(out), flush=True)


if args.ab:
    for v in (0, 1):
        be.SetOption(args.ab, v)
        measure(f"{args.ab}={v}")
else:
    measure("run")
be.close()
