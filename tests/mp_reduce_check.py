"""Two-rank frame-end check, launched by tests/test_gpu_parity.py::test_two_rank_reduce_equals_single_rank as
`python -m torch.distributed.run --nproc-per-node 2 tests/mp_reduce_check.py` (one rank per GPU, NCCL):
  * ranks trace the two halves of an index range; after hb_reduce_image(0) rank 0 holds the image one engine
    accumulates for the whole range (per-pixel tolerance: only the order of the additions differs) and rank 1's
    accumulator is zero; a repeated reduce changes nothing;
  * hb_allreduce_image leaves the sum on both ranks and refuses a second call on the same accumulation;
  * a two-layer scene: the ranks' layer-1 orientation draws differ (stream bases follow the global ray index).
Rank 0 prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ice_halo_sim_b200 import backend as B  # noqa: E402
from ice_halo_sim_b200 import driver, scenes  # noqa: E402
from ice_halo_sim_b200.lib import HaloTraceError  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    be = B.B200TraceBackend(local)
    driver.ensure_comm(be, rank, world)
    case = scenes.CASES["column_config2"]
    tables = B.SceneTables(case["scene"](), 7)
    be.SetScene(tables)
    be.SetRender(case["render"]())
    be.ReadbackXyzAccum()
    wl = [B.make_wl_entry(550.0, 1.0)]
    n = 1 << 21
    out = {}

    def trace(first, count):
        be.BeginSession(B.SessionSpec(seed=77, wl=wl, ray_num=count, ray_base=first))
        be.TraceLayer(B.RootRaySource.FromHost(count), want_stats=False)
        be.EndSession()

    # reference: the whole range on every rank, read back locally
    trace(0, n)
    whole, whole_landed = be.ReadbackXyzAccum()
    # sharded: each rank its half, reduce to rank 0 (twice: idempotent)
    half = n // world
    trace(rank * half, half)
    be.ReduceImage(0)
    be.ReduceImage(0)
    img, landed = be.ReadbackXyzAccum()
    if rank == 0:
        scale = float(np.abs(whole).max())
        out["reduce_landed_rel"] = abs(landed - whole_landed) / whole_landed
        out["reduce_image_ok"] = bool(np.allclose(img, whole, rtol=2e-4, atol=2e-6 * scale))
        out["reduce_sum_rel"] = abs(float(img.astype(np.float64).sum()) / float(whole.astype(np.float64).sum()) - 1.0)
    else:
        out["peer_zero"] = bool(landed == 0.0 and not img.any())
    # all-reduce: the sum on every rank, second call refused
    trace(rank * half, half)
    be.AllReduceImage()
    refused = False
    try:
        be.AllReduceImage()
    except HaloTraceError as e:
        refused = e.status == -4
    img2, landed2 = be.ReadbackXyzAccum()
    out["allreduce_refuses_second_call"] = refused
    out["allreduce_landed_rel"] = abs(landed2 - whole_landed) / whole_landed
    # two layers: layer-1 orientation draws of the two ranks differ
    case2 = scenes.CASES["partial_prob"]
    be.SetScene(B.SceneTables(case2["scene"](), 7))
    m = 4000
    be.BeginSession(B.SessionSpec(seed=3, wl=[B.make_wl_entry(530.0, 1.0)], ray_num=m, record_exits=True, accumulate=False,
                                  ray_base=rank * m))
    h = be.TraceLayer(B.RootRaySource.FromHost(m))
    be.DrainExits()
    be.ExportRoots()
    src = be.Recombine(h, shuffle=True)
    be.TraceLayer(src)
    be.DrainExits()
    r = be.ExportRoots()
    be.EndSession()
    k = 256
    mine = torch.tensor(r["rot"][:k].reshape(-1), device="cuda")
    both = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(both, mine)
    out["layer1_orientations_differ"] = bool((both[0] - both[1]).abs().max().item() > 1e-3)
    # driver.render_config owns the communicator and the frame-end reduce: rank 0 gets the frames, the others nothing
    from ice_halo_sim_b200 import load_config, render_config
    cfg_json = {
        "crystal": [{"id": 1, "type": "prism", "shape": {"height": 1.3},
                     "axis": {"zenith": {"type": "gauss", "mean": 90.0, "std": 0.3},
                              "azimuth": {"type": "uniform", "mean": 0.0, "std": 360.0},
                              "roll": {"type": "uniform", "mean": 0.0, "std": 360.0}}}],
        "filter": [],
        "render": [{"id": 4, "lens": {"type": "fisheye_equal_area", "fov": 120.0}, "resolution": [480, 270],
                    "view": {"azimuth": 0.0, "elevation": 30.0, "roll": 0.0}, "visible": "upper"}],
        "scene": {"id": 1, "light_source": {"type": "sun", "altitude": 20.0, "azimuth": 0.0, "diameter": 0.5,
                                            "spectrum": [{"wavelength": 550.0, "weight": 1.0}]},
                  "ray_num": 1 << 20, "max_hits": 7,
                  "scattering": [{"prob": 0.0, "entries": [{"crystal": 1, "proportion": 1.0}]}]},
    }
    cfg = load_config(cfg_json)
    be2 = B.B200TraceBackend(local)      # a backend WITHOUT a communicator: render_config must create it
    frames = render_config(cfg, be2, seed=5, rank=rank, world=world)
    one = render_config(cfg, be2, seed=5, rank=0, world=1) if rank == 0 else None
    if rank == 0:
        out["driver_frames_on_root"] = sorted(frames) == [4]
        out["driver_landed_rel"] = abs(frames[4].landed_weight - one[4].landed_weight) / one[4].landed_weight
    else:
        out["driver_no_frames_on_peer"] = frames == {}
    be2.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        merged = {}
        for g in gathered:
            merged.update(g)
        print(json.dumps(merged), flush=True)
    dist.destroy_process_group()
    be.close()


if __name__ == "__main__":
    main()
