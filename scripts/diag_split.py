#!/usr/bin/env python
"""Diagnostic: where does |landed(one session) - landed(split sessions)| of test_full_size_invariants come from?
Prints landed weight / image sums for one 16 Mi-ray session, the same in 2 Mi-ray tiles, the three-way split, and
without the pixel cache, plus the per-pixel differences (a ray-set difference shows as isolated pixels, fp32
absorption as a loss concentrated on the hottest pixels)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ice_halo_sim_b200 import backend as B  # noqa: E402
from ice_halo_sim_b200 import scenes  # noqa: E402

case = scenes.CASES["column_config2"]
be = B.B200TraceBackend(0)
be.SetScene(B.SceneTables(case["scene"](), 7))
be.SetRender(case["render"]())
wl = [B.make_wl_entry(550.0, 1.0)]
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 24)


def run(splits, **opts):
    for k, v in opts.items():
        be.SetOption(k, v)
    be.ReadbackXyzAccum()
    base = 0
    for cnt in splits:
        be.BeginSession(B.SessionSpec(seed=99, wl=wl, ray_num=cnt, ray_base=base))
        be.TraceLayer(B.RootRaySource.FromHost(cnt), want_stats=False)
        be.EndSession()
        base += cnt
    img, landed = be.ReadbackXyzAccum()
    return img.astype(np.float64), float(landed)


t8 = max(n >> 3, 4096)
runs = {
    "a  one tile": run([n], tile_rays=1 << 24, pixel_cache=1),
    "a2 n/8 tiles": run([n], tile_rays=t8, pixel_cache=1),
    "b  3 sessions": run([n // 2, n // 2 - 12345, 12345], tile_rays=1 << 24, pixel_cache=1),
    "c  2 halves": run([n // 2, n // 2], tile_rays=1 << 24, pixel_cache=1),
    "a  no cache": run([n], tile_rays=1 << 24, pixel_cache=0),
    "a2 no cache n/8": run([n], tile_rays=t8, pixel_cache=0),
    "a  again": run([n], tile_rays=1 << 24, pixel_cache=1),
}
ref_img, ref_l = runs["a2 n/8 tiles"]
for k, (img, l) in runs.items():
    d = img[..., 1] - ref_img[..., 1]
    hot = np.argsort(ref_img[..., 1].ravel())[-40:]
    print(f"{k:18s} landed {l:14.2f} (vs a2 {l - ref_l:+9.2f})  Ysum {img[..., 1].sum():14.2f}  "
          f"dY on 40 hottest px {d.ravel()[hot].sum():+9.2f}  elsewhere {d.sum() - d.ravel()[hot].sum():+9.2f}  "
          f"max|dY| {np.abs(d).max():.3f}  px with |dY|>1e-3*max(Y,1) {(np.abs(d) > 1e-3 * np.maximum(ref_img[..., 1], 1)).sum()}")
print("hottest pixel Y:", np.sort(ref_img[..., 1].ravel())[-5:])
be.close()
