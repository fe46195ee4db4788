#!/usr/bin/env python
"""Multi-device fan-out BEHIND the seam: the reference's unmodified Simulator::Run (one host thread) drives one
B200TraceBackend instance spread over R devices (oracle/_ref/libhalo_refb200.so, HALOTRACE_B200_DEVICES), config-2
scene, 16 Mi-ray SimBatches. Prints one JSON line per R. Usage: python scripts/seam_scaling.py [batches_per_wl] [R ...]
Times include backend creation (one CUDA context per device) because Simulator::Run creates its backend per Run()."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

batches = int(sys.argv[1]) if len(sys.argv) > 1 else 64
counts = [int(x) for x in sys.argv[2:]] or [1, 2, 4, 8]
base = None
for r in counts:
    res = bench.reference_driver_run("libhalo_refb200.so", batches * bench.SESSION_RAYS, bench.SESSION_RAYS,
                                     devices=list(range(r)))
    res["devices"] = r
    if res.get("mrays_per_s"):
        base = base or res["mrays_per_s"]
        res["vs_one_device"] = res["mrays_per_s"] / base
    print(json.dumps(res), flush=True)
