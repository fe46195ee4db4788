import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, parity
from ice_halo_sim_b200 import backend as B
case = parity.CASES["column_config2"]
be = B.B200TraceBackend(0)
be.SetScene(B.SceneTables(case["scene"](), 7)); be.SetRender(case["render"]())
wl = [B.make_wl_entry(550.0, 1.0)]
for n in (1<<16, 1<<20, 1<<22, 1<<24):
    be.BeginSession(B.SessionSpec(seed=99, wl=wl, ray_num=n, ray_base=0)); be.TraceLayer(B.RootRaySource.FromHost(n), want_stats=False); be.EndSession()
    img, landed = be.ReadbackXyzAccum()
    y = img[...,1].astype(np.float64).sum()
    print(n, "landed", landed, "y", y, "cmf_y*landed", wl[0][3]*landed, "ratio", y/(wl[0][3]*landed), "max pixel", img[...,1].max(), "x/y", img[...,0].astype(np.float64).sum()/y, wl[0][2]/wl[0][3])
