"""GPU-vs-oracle parity protocol, shared by tests/test_gpu_parity.py, smoke() and bench.py.

For every scattering layer: the engine traces on the GPU with exit records on, the roots it generated are
exported (crystal-local d/p/w/face + the rot9 it uses), the oracle replays exactly those roots on the CPU,
and the two exit lists are compared per root: face-number sequences, world directions and weights
bit-for-bit. The GPU image is compared with the oracle's accumulation of the same exit list under a
stated per-pixel tolerance (float atomics reorder the additions).
"""
import ctypes as C

import numpy as np

import harness as H
from ice_halo_sim_b200 import _abi as A
from ice_halo_sim_b200 import backend as B

# Per-pixel image tolerance. The GPU sums the k contributions of a pixel in fp32 in arbitrary order (atomics,
# per-CTA partial sums), the oracle sequentially; each order is within (k-1) * 2^-24 * sum|x| of the exact sum,
# so   |gpu - oracle| <= (IMG_RTOL + IMG_RTOL_PER_TERM * k) * sum|contributions| + IMG_ATOL
# with IMG_RTOL_PER_TERM = 2 * 2^-24.
IMG_RTOL = 1e-6
IMG_RTOL_PER_TERM = 1.2e-7
IMG_ATOL = 1e-7
IMG_FLIP_FRAC = 5e-4


class OrcLayerParams(C.Structure):
    _fields_ = [("shapes", C.c_void_p), ("shape_pop", C.c_void_p), ("shape_cnt", C.c_uint32),
                ("pops", C.c_void_p), ("pop_cnt", C.c_uint32), ("wl", C.c_void_p), ("wl_cnt", C.c_uint32),
                ("max_hits", C.c_uint32), ("prob", C.c_float), ("layer_idx", C.c_uint32), ("seed", C.c_uint32),
                ("gate_base", C.c_uint64), ("root_mask", C.c_void_p), ("cont_mask", C.c_void_p)]


from ice_halo_sim_b200.scenes import (CASES, color_classes, color_pred, complex_filter, dist, prism_pop,  # noqa: E402,F401
                                      pyramid_pop, raypath_filter, render, scene, simple_spec)


def layer_params(sc: A.HbScene, li, wl_arr, seed, gate_base):
    layer = sc.layers[li]
    shapes, shape_pop = [], []
    for ci in range(layer.population_cnt):
        pop = layer.populations[ci]
        for s in range(pop.shape_cnt):
            shapes.append(pop.shapes[s])
            shape_pop.append(ci)
    arr = (A.HbCrystalTables * len(shapes))(*shapes)
    sp = np.array(shape_pop, np.uint32)
    lp = OrcLayerParams(C.addressof(arr), H.ptr(sp), len(shapes), C.addressof(layer.populations.contents),
                        layer.population_cnt, C.addressof(wl_arr), len(wl_arr), sc.max_hits, layer.prob, li, seed,
                        gate_base, None, None)
    return lp, (arr, sp)


def oracle_trace(lp, roots, cap, root_mask=None):
    orc = H.oracle()
    n = len(roots["w"])
    cm = np.zeros(cap, np.uint64)
    lp.cont_mask = H.ptr(cm)
    lp.root_mask = H.ptr(root_mask) if root_mask is not None and len(root_mask) else None
    ex = np.zeros(cap, H.EXIT_DTYPE)
    er = np.zeros(cap, np.uint32)
    ec = C.c_uint64()
    cd = np.zeros((cap, 3), np.float32)
    cw = np.zeros(cap, np.float32)
    cwl = np.zeros(cap, np.uint32)
    cr = np.zeros(cap, np.uint32)
    cc = C.c_uint64()
    rc = orc.orc_trace_layer(C.byref(lp), n, H.ptr(roots["d"]), H.ptr(roots["p"]), H.ptr(roots["w"]),
                             H.ptr(roots["face"]), H.ptr(roots["rot"]), H.ptr(roots["shape"]), H.ptr(roots["wl"]), cap,
                             H.ptr(ex), H.ptr(er), C.byref(ec), H.ptr(cd), H.ptr(cw), H.ptr(cwl), H.ptr(cr),
                             C.byref(cc))
    assert rc == 0, rc
    return ex[: ec.value], er[: ec.value], (cd[: cc.value], cw[: cc.value], cr[: cc.value], cm[: cc.value])


def compare_exit_lists(g_ex, g_roots, o_ex, o_roots):
    g, gr = H.sort_exits(g_ex, g_roots)
    o, orr = H.sort_exits(o_ex, o_roots)
    res = dict(n_gpu=len(g), n_oracle=len(o))
    same_n = len(g) == len(o)
    res["paths_equal"] = bool(same_n and np.array_equal(gr, orr) and np.array_equal(g["path_len"], o["path_len"]) and
                              np.array_equal(g["path"], o["path"]))
    res["dirs_bit_equal"] = bool(same_n and np.array_equal(g["dir"].view(np.uint32), o["dir"].view(np.uint32)))
    res["weights_bit_equal"] = bool(same_n and np.array_equal(g["weight"].view(np.uint32), o["weight"].view(np.uint32)))
    res["meta_equal"] = bool(same_n and np.array_equal(g["crystal_id"], o["crystal_id"]) and
                             np.array_equal(g["ms_layer_idx"], o["ms_layer_idx"]) and
                             np.array_equal(g["wl_idx"], o["wl_idx"]))
    res["masks_equal"] = bool(same_n and np.array_equal(g["component_mask"], o["component_mask"]))
    return res


def oracle_image(proj, wl_arr, exits):
    orc = H.oracle()
    h, w = proj.img_h, proj.img_w
    img = np.zeros((h, w, 3), np.float32)
    mag = np.zeros((h, w, 3), np.float32)
    landed = C.c_double(0.0)
    d = np.ascontiguousarray(exits["dir"])
    ww = np.ascontiguousarray(exits["weight"])
    wi = np.ascontiguousarray(exits["wl_idx"])
    orc.orc_accumulate(C.byref(proj), C.addressof(wl_arr), len(wl_arr), len(ww), H.ptr(d), H.ptr(ww), H.ptr(wi),
                       H.ptr(img), C.byref(landed))
    dummy = C.c_double(0.0)
    wabs = np.ascontiguousarray(np.abs(ww))
    orc.orc_accumulate(C.byref(proj), C.addressof(wl_arr), len(wl_arr), len(ww), H.ptr(d), H.ptr(wabs),
                       H.ptr(wi), H.ptr(mag), C.byref(dummy))
    # number of contributions per pixel: accumulate unit weights with a unit "CMF"
    ones = np.ones(len(ww), np.float32)
    unit = (A.HbWlEntry * len(wl_arr))(*[A.HbWlEntry(1.0, 1.0, 1.0, 1.0, 1.0) for _ in range(len(wl_arr))])
    cnt = np.zeros((h, w, 3), np.float32)
    orc.orc_accumulate(C.byref(proj), C.addressof(unit), len(wl_arr), len(ww), H.ptr(d), H.ptr(ones), H.ptr(wi),
                       H.ptr(cnt), C.byref(dummy))
    return img, mag, landed.value, cnt


def resample_pools_on_device(be, tables, seed, draw_base=0):
    """Redraw every multi-shape pool of the scene on the device (hb_resample_shapes) and copy the device-built
    tables over the host scene's, so the oracle replays exactly the crystals the engine traces.
    Returns [(layer, population, tables, scalars)]."""
    sc = tables.scene()
    out = []
    for li in range(sc.layer_cnt):
        layer = sc.layers[li]
        for pi in range(layer.population_cnt):
            pop = layer.populations[pi]
            if pop.shape_cnt <= 1:
                continue
            be.ResampleShapes(li, pi, tables.desc.layers[li].populations[pi].crystal, seed, draw_base)
            tb, scal = be.ExportShapes(li, pi)
            assert len(tb) == pop.shape_cnt
            C.memmove(pop.shapes, tb, C.sizeof(A.HbCrystalTables) * len(tb))
            out.append((li, pi, tb, scal))
    return out


def sync_host_scene_with_device_pools(be, tables):
    """Copy the engine's LIVE shape pools over the host scene's (geometry clock: the pool changes per session)."""
    sc = tables.scene()
    for li in range(sc.layer_cnt):
        layer = sc.layers[li]
        for pi in range(layer.population_cnt):
            pop = layer.populations[pi]
            if pop.shape_cnt > 1:
                tb, _ = be.ExportShapes(li, pi)
                C.memmove(pop.shapes, tb, C.sizeof(A.HbCrystalTables) * len(tb))


def run_case(case, n_rays=20000, seed=42, device=0, backend=None, geometry_seed=7, tile_rays=None, device_pool_seed=None,
             geometry_clock_seed=None, fused_bounce=None):
    """Full protocol on one case; returns a dict of comparison results (all layers merged)."""
    desc = case["scene"]()
    rdescs = case["render"]()
    if not isinstance(rdescs, (list, tuple)):
        rdescs = [rdescs]
    tables = B.SceneTables(desc, geometry_seed)
    sc = tables.scene()
    wl = [B.make_wl_entry(x, 1.0) for x in case["wl"]]
    wl_arr = (A.HbWlEntry * len(wl))(*[A.HbWlEntry(*e) for e in wl])
    own = backend is None
    be = backend or B.B200TraceBackend(device)
    if tile_rays:
        be.SetOption("tile_rays", tile_rays)
    if fused_bounce is not None:   # 1: one bounce kernel per interaction (default), 0: split optics + intersect kernels
        be.SetOption("fused_bounce", int(fused_bounce))
    be.SetScene(tables)
    if device_pool_seed is not None:
        resample_pools_on_device(be, tables, device_pool_seed)
    if geometry_clock_seed is not None:   # engine-run geometry clock: pools are swapped in by BeginSession
        for li in range(desc.layer_cnt):
            for pi in range(desc.layers[li].population_cnt):
                if sc.layers[li].populations[pi].shape_cnt > 1:
                    be.AutoResample(li, pi, desc.layers[li].populations[pi].crystal, geometry_clock_seed, 0)
    be.SetOption("stream_base", 0)
    be.SetRenders(rdescs)
    projs = [B.make_proj_params(r) for r in rdescs]
    for r in range(len(rdescs)):
        be.ReadbackXyzAccum(render=r)  # start from zero images
    be.BeginSession(B.SessionSpec(seed=seed, wl=wl, ray_num=n_rays, record_exits=True, accumulate=True))
    if geometry_clock_seed is not None:
        sync_host_scene_with_device_pools(be, tables)
    out = dict(paths_equal=True, dirs_bit_equal=True, weights_bit_equal=True, meta_equal=True, exits=0, layers=[])
    all_exits = []
    roots_src = B.RootRaySource.FromHost(n_rays)
    gate_base = 0
    stats_ok = True
    try:
        for li in range(desc.layer_cnt):
            handle = be.TraceLayer(roots_src)
            g_ex, g_roots = be.DrainExits(with_roots=True)
            roots = be.ExportRoots()
            root_masks = be.ExportRootMasks()
            lp, keep = layer_params(sc, li, wl_arr, seed, gate_base)
            cap = len(roots["w"]) * (desc.max_hits + 2) + 16
            o_ex, o_roots, o_cont = oracle_trace(lp, roots, cap, root_masks)
            r = compare_exit_lists(g_ex, g_roots, o_ex, o_roots)
            out["masks_equal"] = out.get("masks_equal", True) and r["masks_equal"]
            out["mask_bits_seen"] = out.get("mask_bits_seen", 0) | int(np.bitwise_or.reduce(o_ex["component_mask"])) \
                if len(o_ex) else out.get("mask_bits_seen", 0)
            # continuation masks travel with the (shuffled) pool: compare as multisets
            out["cont_masks_oracle"] = np.sort(o_cont[3])
            r["continuations_gpu"] = handle.continuation_count
            r["continuations_oracle"] = len(o_cont[1])
            r["roots"] = len(roots["w"])
            # LayerStats: exit_count / exit_w_sum cover outgoing + continuation (trace_backend.hpp:279-300)
            tot_w = float(o_ex["weight"].astype(np.float64).sum() + o_cont[1].astype(np.float64).sum())
            r["stats_count_equal"] = handle.exit_count == len(o_ex) + len(o_cont[1])
            r["stats_w_rel_err"] = abs(handle.exit_w_sum - tot_w) / max(tot_w, 1e-30)
            stats_ok = stats_ok and r["stats_count_equal"] and r["stats_w_rel_err"] < 1e-6
            for k in ("paths_equal", "dirs_bit_equal", "weights_bit_equal", "meta_equal"):
                out[k] = out[k] and r[k]
            out[k] = out[k] and (r["continuations_gpu"] == r["continuations_oracle"])
            out["exits"] += len(g_ex)
            out["layers"].append(r)
            all_exits.append(g_ex)
            gate_base += len(roots["w"])
            if li + 1 == desc.layer_cnt:
                break
            roots_src = be.Recombine(handle, shuffle=True)
    finally:
        be.EndSession()
    out["stats_ok"] = stats_ok
    ex = np.concatenate(all_exits) if all_exits else np.zeros(0, H.EXIT_DTYPE)
    out["renders"] = []
    for r, proj in enumerate(projs):   # every render of the trace against the oracle's projection of the same exits
        img, landed = be.ReadbackXyzAccum(render=r)
        o_img, o_mag, o_landed, o_cnt = oracle_image(proj, wl_arr, ex)
        err = np.abs(img - o_img)
        tol = (IMG_RTOL + IMG_RTOL_PER_TERM * o_cnt) * o_mag + IMG_ATOL
        q = dict(image_within_tol=bool(np.all(err <= tol)))
        # Lenses whose forward map uses atan2/asin/acos/tan differ between glibc (oracle) and libdevice (GPU) by
        # an ulp, which moves a ray sitting on a pixel boundary into the neighbouring pixel: for those the
        # per-pixel check is replaced by "total |difference| <= IMG_FLIP_FRAC of the image" (a few rays in 1e5).
        total = float(o_mag.astype(np.float64).sum())
        q["image_l1_frac"] = float(err.astype(np.float64).sum()) / max(total, 1e-30)
        q["lens_transcendental"] = int(proj.proj_type) in (2, 3, 5, 6, 7)
        q["image_ok"] = q["image_within_tol"] or (q["lens_transcendental"] and q["image_l1_frac"] <= IMG_FLIP_FRAC)
        q["image_max_rel_err"] = float((err / np.maximum(o_mag, 1e-20)).max()) if o_mag.max() > 0 else 0.0
        q["image_sum"] = float(img.astype(np.float64).sum())
        q["landed_gpu"] = float(landed)
        q["landed_oracle"] = float(o_landed)
        q["landed_rel_err"] = abs(landed - o_landed) / max(abs(o_landed), 1e-30)
        out["renders"].append(q)
    out.update(out["renders"][0])
    out["image_ok"] = all(q["image_ok"] for q in out["renders"])
    out["landed_rel_err"] = max(q["landed_rel_err"] for q in out["renders"])
    # per-class Y lanes of render 0 (ReadbackClassLanes) against the oracle's fan-out of the same exits
    ncls = int(sc.color_classes.class_cnt)
    if ncls:
        lanes = be.ReadbackClassLanes()
        proj = projs[0]
        want = np.zeros((ncls, proj.img_h, proj.img_w), np.float32)
        mag = np.zeros_like(want)
        d = np.ascontiguousarray(ex["dir"]); ww = np.ascontiguousarray(ex["weight"])
        wi = np.ascontiguousarray(ex["wl_idx"]); mk = np.ascontiguousarray(ex["component_mask"])
        H.oracle().orc_accumulate_lanes(C.byref(proj), C.addressof(wl_arr), len(wl_arr), C.byref(sc.color_classes),
                                        len(ww), H.ptr(d), H.ptr(ww), H.ptr(wi), H.ptr(mk), H.ptr(want))
        wabs = np.ascontiguousarray(np.abs(ww))
        H.oracle().orc_accumulate_lanes(C.byref(proj), C.addressof(wl_arr), len(wl_arr), C.byref(sc.color_classes),
                                        len(ww), H.ptr(d), H.ptr(wabs), H.ptr(wi), H.ptr(mk), H.ptr(mag))
        out["lanes_shape_ok"] = lanes is not None and lanes.shape == want.shape
        if out["lanes_shape_ok"]:
            err = np.abs(lanes - want)
            out["lanes_ok"] = bool(np.all(err <= 2e-5 * mag + 1e-7))
            out["lane_sums"] = [float(x) for x in lanes.reshape(ncls, -1).astype(np.float64).sum(axis=1)]
        else:
            out["lanes_ok"] = False
    if own:
        be.close()
    return out
