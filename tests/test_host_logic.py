"""Host-side logic that needs no GPU: Lumice JSON config parsing, scene-table building, ray-index
sharding, and the world_size-2 (gloo) check that sharded traces + a sum all-reduce equal the 1-rank run."""
import ctypes as C
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import harness as H
from ice_halo_sim_b200 import backend as B
from ice_halo_sim_b200 import config as CFG
from ice_halo_sim_b200 import sharding as S

A = H.A

# the parts of examples/config_example.json the trace path reads (crystals 3/5/6, filters 2-6, scene, render 4)
EXAMPLE = {
    "crystal": [
        {"id": 3, "type": "prism", "shape": {"height": 1.3, "face_distance": [1, 1, 1, 1, 1, 1]},
         "axis": {"zenith": {"type": "gauss", "mean": 90, "std": 0.3}, "roll": {"type": "uniform", "mean": 0, "std": 360},
                  "azimuth": {"type": "uniform", "mean": 0, "std": 360}}},
        {"id": 5, "type": "pyramid", "shape": {"upper_h": 0.1, "lower_h": 0.5, "prism_h": 1.2, "upper_indices": [2, 0, 3]},
         "axis": {"zenith": 0}},
        {"id": 6, "type": "prism", "shape": {"height": 0.3}, "axis": {"zenith": {"type": "gauss", "mean": 0, "std": 0.8}}},
    ],
    "filter": [
        {"id": 2, "type": "raypath", "raypath": [3, 1, 5, 7, 4], "symmetry": "PBD"},
        {"id": 3, "type": "raypath", "raypath": [3, 5], "symmetry": "P"},
        {"id": 4, "type": "entry_exit", "entry": 3, "exit": 5, "action": "filter_in"},
        {"id": 5, "type": "direction", "az": 180, "el": 25, "radii": 0.5, "action": "filter_out"},
        {"id": 6, "type": "crystal", "crystal_id": 3},
    ],
    "scene": {
        "light_source": {"type": "sun", "altitude": 20.0, "azimuth": 0, "diameter": 0.5,
                         "spectrum": [{"wavelength": w, "weight": 1.0} for w in range(450, 771, 40)]},
        "ray_num": 450000000, "max_hits": 7,
        "scattering": [{"prob": 0.5, "entries": [{"crystal": 6, "proportion": 10, "filter": 3},
                                                 {"crystal": 5, "proportion": 2}]},
                       {"prob": 0.0, "entries": [{"crystal": 3, "proportion": 10}]}],
    },
    "render": [{"id": 1, "lens": {"type": "linear", "f": 14}, "resolution": [1920, 1080], "lens_shift": [0, 200],
                "view": {"azimuth": -10, "elevation": 20, "roll": 0}},
               {"id": 4, "lens": {"type": "fisheye_equal_area", "fov": 120}, "resolution": [1920, 1080],
                "view": {"elevation": 30}}],
}


def test_config_parsing():
    cfg = CFG.load_config(EXAMPLE)
    d = cfg.desc
    assert d.max_hits == 7 and d.layer_cnt == 2
    assert (d.sun_altitude_deg, d.sun_azimuth_deg, d.sun_diameter_deg) == (20.0, 0.0, 0.5)
    assert len(cfg.spectrum) == 9 and cfg.spectrum[0] == (450.0, 1.0) and cfg.spectrum[-1] == (770.0, 1.0)
    assert cfg.rays_per_wavelength() == 50_000_000          # ceil(450 M / 9), ray_num_semantics.hpp:12-16
    l0 = d.layers[0]
    assert abs(l0.prob - 0.5) < 1e-7 and l0.population_cnt == 2
    plate = l0.populations[0]
    assert plate.crystal.kind == 0 and plate.crystal.id == 6 and abs(plate.crystal.height[0].center - 0.3) < 1e-7
    # zenith gauss(0, 0.8) -> latitude gauss(90, 0.8); azimuth/roll default to uniform(0, 360) (math.cpp:707-713)
    assert (plate.crystal.latitude.type, plate.crystal.latitude.center) == (A.DIST["gauss"], 90.0)
    assert abs(plate.crystal.latitude.spread - 0.8) < 1e-6
    assert (plate.crystal.azimuth.type, plate.crystal.azimuth.spread) == (A.DIST["uniform"], 360.0)
    assert (plate.filter.kind, plate.filter.symmetry, plate.filter.simple.path_len) == (1, 1, 2)
    assert list(plate.filter.simple.path[:2]) == [3, 5]
    pyr = l0.populations[1]
    assert pyr.crystal.kind == 1 and abs(pyr.crystal.height[1].center - 1.2) < 1e-6
    assert abs(pyr.crystal.wedge_upper_deg - np.degrees(np.arctan(0.866025403784 * 3 / 2 / 1.629))) < 1e-4
    assert pyr.crystal.wedge_lower_deg == 28.0 and pyr.filter.kind == 0
    assert pyr.crystal.latitude.center == 90.0 and pyr.crystal.azimuth.type == A.DIST["uniform"]
    col = d.layers[1].populations[0]
    assert col.crystal.id == 3 and col.crystal.latitude.center == 0.0 and abs(col.crystal.latitude.spread - 0.3) < 1e-6
    r4 = cfg.renders[4]
    assert (r4.lens_type, r4.fov_deg, r4.img_w, r4.img_h, r4.view_el_deg) == (1, 120.0, 1920, 1080, 30.0)
    r1 = cfg.renders[1]
    assert r1.lens_type == 0 and abs(r1.fov_deg - 2 * np.degrees(np.arctan2(12.0, 14.0))) < 1e-4
    assert (r1.lens_shift_x, r1.lens_shift_y) == (0, 200)
    # round-trip through a file
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump(EXAMPLE, f)
    assert CFG.load_config(f.name).desc.layer_cnt == 2
    os.unlink(f.name)
    # complex filter (config_example.json filter 7): OR over [2, (3 AND 6), 5]
    ex2 = json.loads(json.dumps(EXAMPLE))
    ex2["filter"].append({"id": 7, "type": "complex", "composition": [2, [3, 6], 5], "symmetry": "P"})
    ex2["scene"]["scattering"][0]["entries"][0]["filter"] = 7
    f7 = CFG.load_config(ex2).desc.layers[0].populations[0].filter
    assert f7.kind == 5 and f7.term_cnt == 3 and list(f7.term_len)[:3] == [1, 2, 1] and f7.symmetry == 1
    assert f7.terms[0][0].kind == 1 and f7.terms[0][0].path_len == 5
    assert f7.terms[1][0].kind == 1 and f7.terms[1][1].kind == 4 and f7.terms[1][1].crystal_id == 3
    assert f7.terms[2][0].kind == 3 and abs(f7.terms[2][0].radii_deg - 0.5) < 1e-7
    with pytest.raises(ValueError):
        CFG.load_config({**EXAMPLE, "filter": [{"id": 3, "type": "complex", "composition": list(range(9))}]})


def test_raypath_color_parsing():
    """raypath_color -> component bits in first-occurrence order with (layer, crystal, predicate, symmetry)
    dedupe (color_gate_table.cpp:47-95) and class bit sets / combine (color_class_table.hpp)."""
    ex = json.loads(json.dumps(EXAMPLE))
    ex["raypath_color"] = [
        {"color": [1, 0, 0], "match": [{"layer": 0, "crystal": 6, "type": "raypath", "raypath": [3, 5], "symmetry": "P"},
                                       {"layer": 1, "crystal": 3}]},
        {"color": [0, 1, 0], "combine": "all",
         "match": [{"layer": 0, "crystal": 6, "type": "raypath", "raypath": [3, 5], "symmetry": "P"},   # duplicate -> bit 0
                   {"layer": 0, "crystal": 6, "type": "raypath", "raypath": [3, 5]},                    # other symmetry -> new bit
                   {"layer": 0, "crystal": 5, "type": "entry_exit", "entry": 1, "exit": 2}]},
    ]
    cfg = CFG.load_config(ex)
    d = cfg.desc
    assert d.color_classes.class_cnt == 2 and d.color_classes.combine_all_mask == 0b10
    assert d.color_classes.bits[0] == 0b0011 and d.color_classes.bits[1] == 0b1101
    plate, pyr, col = d.layers[0].populations[0], d.layers[0].populations[1], d.layers[1].populations[0]
    assert plate.color_pred_cnt == 2 and [plate.color_preds[k].bit for k in range(2)] == [0, 2]
    assert [plate.color_preds[k].symmetry for k in range(2)] == [1, 0]
    assert pyr.color_pred_cnt == 1 and pyr.color_preds[0].bit == 3 and pyr.color_preds[0].pred.kind == 2
    assert col.color_pred_cnt == 1 and col.color_preds[0].bit == 1 and col.color_preds[0].pred.kind == 0
    assert cfg.color_classes[1]["combine"] == "all" and cfg.color_classes[0]["color"] == (1.0, 0.0, 0.0)
    # object form with a composite mode; errors mirror BuildColorGateTable's invalid_argument cases
    ex["raypath_color"] = {"mode": "additive", "classes": ex["raypath_color"]}
    assert CFG.load_config(ex).desc.color_classes.class_cnt == 2
    for bad in ({"layer": 5, "crystal": 6}, {"layer": 0, "crystal": 99}):
        ex["raypath_color"] = [{"color": [1, 1, 1], "match": [bad]}]
        with pytest.raises(ValueError):
            CFG.load_config(ex)
    ex["raypath_color"] = [{"color": [1, 1, 1], "combine": "xor", "match": []}]
    with pytest.raises(ValueError):
        CFG.load_config(ex)
    # the host builder groups predicates by symmetry value in first-occurrence order
    from ice_halo_sim_b200 import backend as B
    ex["raypath_color"] = [{"color": [1, 0, 0], "match": [
        {"layer": 0, "crystal": 6, "type": "raypath", "raypath": [3, 5], "symmetry": "P"},
        {"layer": 0, "crystal": 6, "type": "raypath", "raypath": [1, 3, 2]},
        {"layer": 0, "crystal": 6, "type": "entry_exit", "entry": 3, "exit": 5, "symmetry": "P"}]}]
    tables = B.SceneTables(CFG.load_config(ex).desc, 1)
    gp = tables.scene().layers[0].populations[0]
    assert gp.color_group_cnt == 2
    g0, g1 = gp.color_groups[0], gp.color_groups[1]
    assert (g0.filter.kind, g0.filter.symmetry, g0.filter.term_cnt) == (5, 1, 2) and list(g0.bit)[:2] == [0, 2]
    assert (g1.filter.symmetry, g1.filter.term_cnt) == (0, 1) and g1.bit[0] == 1
    assert tables.scene().color_classes.class_cnt == 1 and tables.scene().color_classes.bits[0] == 0b111


def test_scene_tables_from_config():
    cfg = CFG.load_config(EXAMPLE, geom_pool_size=8)
    t = B.SceneTables(cfg.desc, geometry_seed=3)
    sc = t.scene()
    assert sc.layer_cnt == 2 and sc.max_hits == 7
    assert abs(sc.sun_lon - np.pi) < 1e-6 and abs(sc.sun_lat + np.radians(20.0)) < 1e-6
    assert abs(sc.sun_half_angle - np.radians(0.25)) < 1e-7
    plate = sc.layers[0].populations[0]
    assert plate.shape_cnt == 1 and plate.shapes[0].face_cnt == 8 and plate.shapes[0].subtri_cnt == 20
    assert plate.axis.lat_path == A.LAT_LUT and plate.axis.lut_n == 257
    assert plate.filter.kind == 1 and plate.filter.simple.path_len == 2 and list(plate.filter.simple.path[:2]) == [3, 5]
    pyr = sc.layers[0].populations[1]
    assert pyr.shapes[0].face_cnt > 8 and pyr.axis.lat_path == A.LAT_NO_RANDOM
    # stochastic shapes draw a pool, deterministic ones do not (IsDeterministic, simulator.cpp:453-471)
    d2 = CFG.load_config(EXAMPLE, geom_pool_size=8).desc
    d2.layers[1].populations[0].crystal.face_dist[2] = A.HbDist(A.DIST["gauss"], 1.0, 0.15)
    t2 = B.SceneTables(d2, geometry_seed=3)
    col = t2.scene().layers[1].populations[0]
    assert col.shape_cnt == 8
    d0 = [col.shapes[i].plane[4][3] for i in range(8)]
    assert len(set(d0)) == 8


def test_shard_ranges_cover_disjointly():
    for total in (0, 1, 7, 450_000_000, 10**9 + 7):
        for world in (1, 2, 3, 4, 8):
            r = [S.shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    plan = S.session_plan(100, 1, 2, 16, index_base=1000)
    assert plan[0] == (1050, 16) and sum(n for _, n in plan) == 50 and plan[-1] == (1098, 2)
    with pytest.raises(ValueError):
        S.shard_range(10, 2, 2)


WORKER = textwrap.dedent("""
    import os, sys, ctypes as C
    import numpy as np
    import torch, torch.distributed as dist
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    import harness as H, parity
    from ice_halo_sim_b200 import backend as B, sharding as S
    A = H.A
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
    rank, world = dist.get_rank(), dist.get_world_size()
    case = parity.CASES["column_config2"]
    tables = B.SceneTables(case["scene"](), 7)
    rd = parity.render(res=(240, 135))
    proj = B.make_proj_params(rd)
    wl = [B.make_wl_entry(550.0, 1.0)]
    wl_arr = (A.HbWlEntry * 1)(*[A.HbWlEntry(*e) for e in wl])
    orc = H.oracle()
    total = 6000

    def trace(begin, n):
        r = dict(d=np.zeros((n, 3), np.float32), p=np.zeros((n, 3), np.float32), w=np.zeros(n, np.float32),
                 face=np.zeros(n, np.uint16), rot=np.zeros((n, 9), np.float32), shape=np.zeros(n, np.uint32),
                 wl=np.zeros(n, np.uint32))
        orc.orc_gen_roots(tables.scene_ptr, 0, 0, 0, C.addressof(wl_arr), 1, 42, begin, n, H.ptr(r["d"]), H.ptr(r["p"]),
                          H.ptr(r["w"]), H.ptr(r["face"]), None, H.ptr(r["rot"]), H.ptr(r["shape"]), H.ptr(r["wl"]))
        lp, keep = parity.layer_params(tables.scene(), 0, wl_arr, 42, begin)
        ex, er, _ = parity.oracle_trace(lp, r, n * 9 + 16)
        img, mag, landed, _ = parity.oracle_image(proj, wl_arr, ex)
        return img, landed, len(ex)

    img = np.zeros((135, 240, 3), np.float32); landed = 0.0; exits = 0
    for begin, n in S.session_plan(total, rank, world, 1024):
        i, l, e = trace(begin, n)
        img += i; landed += l; exits += e
    t = torch.from_numpy(np.concatenate([img.ravel().astype(np.float64), [landed, exits]]))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)           # the frame-end image reduction (SURVEY 8(e))
    if rank == 0:
        ref_img, ref_landed, ref_exits = trace(0, total)
        got = t.numpy()
        assert int(got[-1]) == ref_exits, (got[-1], ref_exits)
        assert abs(got[-2] - ref_landed) <= 1e-6 * ref_landed
        assert np.allclose(got[:-2].reshape(ref_img.shape), ref_img, rtol=1e-5, atol=1e-6)
        print("SHARD_OK", ref_exits)
    dist.destroy_process_group()
""")


def test_two_rank_sharding_equals_single_rank(tmp_path):
    """world_size 2 over gloo: each rank traces its own contiguous global ray-index range (counter-based RNG,
    oracle as the tracer on CPU) and the images are sum-all-reduced; the result equals the 1-rank trace of the
    whole range — the property the multi-GPU path relies on."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=H.ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "SHARD_OK" in outs[0]


def test_shape_sync_groups_parse_and_sample():
    """Crystal shape "sync_group" (crystal_config.cpp:47-146,179-205 + SyncGroupSampler, simulator.cpp:341-393):
    singleton groups dissolve, survivors are renumbered by first appearance, members take the leader's
    distribution, and the host pool builder draws ONE value per group and crystal."""
    import ctypes as C
    from ice_halo_sim_b200 import _abi as A
    from ice_halo_sim_b200 import backend as B
    from ice_halo_sim_b200 import config as cfgmod
    g = {"type": "gauss", "mean": 1.0, "std": 0.2}
    c = {"id": 1, "type": "prism",
         "shape": {"height": 1.2, "face_distance": [g, g, {"type": "gauss", "mean": 1.0, "std": 0.05}, g, g, g],
                   "sync_group": {"height": 9, "face_distance": [7, 3, 7, 0, 3, 5]}}}
    d = cfgmod.crystal_desc(c)
    assert list(d.sync_group) == [0, 0, 0, 0, 1, 2, 1, 0, 2, 0]          # height 9 and face 5 are singletons
    assert abs(d.face_dist[2].spread - 0.2) < 1e-7                        # member takes the leader's distribution
    pop = A.HbPopulationDesc()
    pop.proportion = 1.0
    pop.crystal = d
    pop.crystal.latitude = A.HbDist(1, 90.0, 360.0)
    pop.crystal.azimuth = pop.crystal.roll = A.HbDist(1, 0.0, 360.0)
    pop.filter.simple.entry_fn = pop.filter.simple.exit_fn = -1
    sd = A.HbSceneDesc()
    sd.max_hits, sd.layer_cnt, sd.geom_pool_size = 4, 1, 64
    sd.sun_altitude_deg, sd.sun_diameter_deg = 20.0, 0.5
    sd.layers[0].prob, sd.layers[0].population_cnt = 0.0, 1
    sd.layers[0].populations[0] = pop
    tables = B.SceneTables(sd, 11)
    p = tables.scene().layers[0].populations[0]
    assert p.shape_cnt == 64
    same02 = same14 = diff01 = 0
    for k in range(64):
        t = p.shapes[k]
        if t.face_cnt != 8:
            continue
        d0 = {int(t.face_fn[f]): float(t.plane[f][3]) for f in range(8)}
        same02 += d0[3] == d0[5]
        same14 += d0[4] == d0[7]
        diff01 += d0[3] != d0[4]
    assert same02 >= 60 and same14 == same02 and diff01 >= 60
