"""Pins oracle/halo_oracle.cpp against the committed golden fixtures (tests/golden/*.npz), which were
produced by the UNMODIFIED reference CPU core (oracle/make_golden.py) and include the reference's own
known-answer scenarios: Sellmeier KATs (test_optics.cpp:26-55), HitSurface/Propagate vectors
(:137-840), golden rays through the h=1 prism (test_cpu_golden_rays.cpp:144-310), 11-lens projection
anchors (test/golden-analytic/core/test_projection.cpp)."""
import ctypes as C
import os

import numpy as np
import pytest

import harness as H

A = H.A
G = H.GOLDEN


def tables_from_golden(g, name):
    t = A.HbCrystalTables()
    t.face_cnt = int(g[f"{name}.face_cnt"])
    t.subtri_cnt = int(g[f"{name}.subtri_cnt"])
    for fld in ("plane", "tri_v", "tri_n", "tri_area", "tri_face", "face_fn"):
        arr = np.ascontiguousarray(g[f"{name}.{fld}"])
        C.memmove(C.addressof(getattr(t, fld)), arr.ctypes.data, arr.nbytes)
    return t


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_sellmeier_known_answers():
    g = np.load(os.path.join(G, "optics.npz"))
    n = dict(zip(g["sellmeier.wl"], g["sellmeier.n"]))
    # reference KATs (test_optics.cpp:26-55), to the reference's own tolerance
    assert abs(n[589.0] - 1.3097193039) < 1e-6
    assert abs(n[550.0] - 1.3110129171) < 1e-6
    assert abs(n[350.0] - 1.3246528270) < 1e-6
    assert abs(n[900.0] - 1.3031960565) < 1e-6
    assert n[349.0] == 1.0 and n[901.0] == 1.0


@pytest.mark.parametrize("name", ["prism_h1", "column_h1p3", "pyramid_full", "prism_irregular"])
def test_hit_surface_and_propagate_bit_exact(name):
    g = np.load(os.path.join(G, "optics.npz"))
    gt = np.load(os.path.join(G, "crystal_tables.npz"))
    t = tables_from_golden(gt, name)
    orc = H.oracle()
    d, w, face = g[f"{name}.hs_d"], g[f"{name}.hs_w"], g[f"{name}.hs_face"]
    n = len(w)
    d_out = np.zeros((n, 6), np.float32)
    w_out = np.zeros((n, 2), np.float32)
    orc.orc_hit_surface(C.byref(t), 1.31, n, H.ptr(d), H.ptr(w), H.ptr(face), H.ptr(d_out), H.ptr(w_out))
    assert np.array_equal(bits(w_out), bits(g[f"{name}.hs_wout"]))
    tir = g[f"{name}.hs_wout"][:, 1] < 0
    assert tir.any() and (~tir).any()          # both Snell and total-internal-reflection branches exercised
    assert np.array_equal(bits(d_out), bits(g[f"{name}.hs_dout"]))
    # energy: w_refl + w_refr == w (within 1 ulp-ish) where not TIR; R in [0, 1]
    ok = ~tir
    assert np.all(np.abs(w_out[ok].sum(axis=1) - w[ok]) <= 2e-7)
    # Propagate incl. source-face exclusion (second half sits exactly on a face)
    p, ff = g[f"{name}.pr_p"], g[f"{name}.pr_from"]
    p_out = np.zeros((n, 3), np.float32)
    to = np.zeros(n, np.uint16)
    orc.orc_propagate(C.byref(t), n, H.ptr(d), H.ptr(p), H.ptr(w), H.ptr(ff), H.ptr(p_out), H.ptr(to))
    assert np.array_equal(to, g[f"{name}.pr_to"])
    assert np.array_equal(bits(p_out), bits(g[f"{name}.pr_pout"]))
    assert (to == 0xFFFF).any() and (to != 0xFFFF).any()


def test_golden_rays_match_closed_form_physics():
    """Normal incidence on the top face of the h=1 prism: back-reflection (0,0,1) w = R, transmission
    (0,0,-1) w = T^2 with R = ((n-1)/(n+1))^2; Snell 30 deg keeps its direction with w = t_in * t_out."""
    g = np.load(os.path.join(G, "optics.npz"))
    n = float(g["golden_rays.n"])
    ex, roots = g["golden_rays.mh2.exits"], g["golden_rays.mh2.roots"]
    r0 = ex[roots == 0]
    R = ((n - 1) / (n + 1)) ** 2
    back = r0[r0["path_len"] == 1][0]
    thru = r0[r0["path_len"] == 2][0]
    assert np.allclose(back["dir"], [0, 0, 1], atol=1e-6) and abs(back["weight"] - R) < 5e-6
    assert np.allclose(thru["dir"], [0, 0, -1], atol=1e-6) and abs(thru["weight"] - (1 - R) ** 2) < 5e-6
    assert list(back["path"][:1]) == [1] and list(thru["path"][:2]) == [1, 2]
    r1 = ex[roots == 1]
    thru1 = r1[r1["path_len"] == 2][0]
    assert np.allclose(thru1["dir"], g["golden_rays.d"][1], atol=1e-5)   # slab: direction preserved
    for k in (2, 7):
        e, r = g[f"golden_rays.mh{k}.exits"], g[f"golden_rays.mh{k}.roots"]
        for root in (0, 1):
            assert e[r == root]["weight"].sum() <= 1.0 + 1e-6        # energy never created
    # and the oracle reproduces them bit-for-bit
    gt = np.load(os.path.join(G, "crystal_tables.npz"))
    t = tables_from_golden(gt, "prism_h1")
    for k in (2, 7):
        o_ex, o_roots = oracle_trace_single(t, np.float32(n), k, g["golden_rays.d"], g["golden_rays.p"],
                                            np.ones(2, np.float32), np.zeros(2, np.uint16))
        a, ar = H.sort_exits(o_ex, o_roots)
        b, br = H.sort_exits(g[f"golden_rays.mh{k}.exits"], g[f"golden_rays.mh{k}.roots"])
        assert np.array_equal(ar, br) and np.array_equal(a["path"], b["path"])
        assert np.array_equal(bits(a["dir"]), bits(b["dir"])) and np.array_equal(bits(a["weight"]), bits(b["weight"]))


class _Lp(C.Structure):
    _fields_ = [("shapes", C.c_void_p), ("shape_pop", C.c_void_p), ("shape_cnt", C.c_uint32),
                ("pops", C.c_void_p), ("pop_cnt", C.c_uint32), ("wl", C.c_void_p), ("wl_cnt", C.c_uint32),
                ("max_hits", C.c_uint32), ("prob", C.c_float), ("layer_idx", C.c_uint32), ("seed", C.c_uint32),
                ("gate_base", C.c_uint64), ("root_mask", C.c_void_p), ("cont_mask", C.c_void_p)]


def oracle_trace_single(t, n_idx, max_hits, d, p, w, face, pop=None):
    orc = H.oracle()
    wl = A.HbWlEntry(float(n_idx), 1.0, 0.0, 0.0, 0.0)
    lp = _Lp(C.addressof(t), None, 1, C.addressof(pop) if pop is not None else None, 1 if pop is not None else 0,
             C.addressof(wl), 1, max_hits, 0.0, 0, 1, 0, None, None)
    n = len(w)
    cap = n * (max_hits + 2) + 8
    ex = np.zeros(cap, H.EXIT_DTYPE)
    er = np.zeros(cap, np.uint32)
    ec = C.c_uint64()
    d, p, w, face = (np.ascontiguousarray(x) for x in (d, p, w, face))
    rc = orc.orc_trace_layer(C.byref(lp), n, H.ptr(d), H.ptr(p), H.ptr(w), H.ptr(face), None, None, None, cap,
                             H.ptr(ex), H.ptr(er), C.byref(ec), None, None, None, None, None)
    assert rc == 0
    return ex[: ec.value], er[: ec.value]


@pytest.mark.parametrize("key,shape", [("column_h1p3.mh7", "column_h1p3"), ("plate_h0p3.mh7", "plate_h0p3"),
                                       ("pyramid_full.mh8", "pyramid_full"), ("prism_irregular.mh12", "prism_irregular")])
def test_traced_face_sequences_bit_exact(key, shape):
    """Reference CpuTraceBackend (HostRayBatch injection, one ray per session) vs the oracle on the same
    rays: face-number sequences, exit directions and weights identical bit for bit."""
    g = np.load(os.path.join(G, "traces.npz"))
    gt = np.load(os.path.join(G, "crystal_tables.npz"))
    t = tables_from_golden(gt, shape)
    mh = int(key.split("mh")[1])
    ex, er = oracle_trace_single(t, g[f"{key}.n_idx"], mh, g[f"{key}.d"], g[f"{key}.p"], g[f"{key}.w"], g[f"{key}.face"])
    a, ar = H.sort_exits(ex, er)
    assert len(a) == len(g[f"{key}.exit_root"])
    assert np.array_equal(ar, g[f"{key}.exit_root"])
    assert np.array_equal(a["path_len"], g[f"{key}.exit_len"])
    assert np.array_equal(a["path"][:, :16], g[f"{key}.exit_path"])
    assert np.array_equal(bits(a["dir"]), bits(g[f"{key}.exit_dir"]))
    assert np.array_equal(bits(a["weight"]), bits(g[f"{key}.exit_w"]))
    assert a["path_len"].max() <= mh


def test_projection_all_lens_types():
    g = np.load(os.path.join(G, "projection.npz"))
    orc = H.oracle()
    dirs = g["dirs"]
    n = len(dirs)
    names = sorted({k.split(".")[0] for k in g.files if "." in k})
    assert len(names) == 12
    for name in names:
        pp = A.HbProjParams.from_buffer_copy(g[f"{name}.params"].tobytes())
        px = np.zeros((n, 2), np.int32)
        py = np.zeros((n, 2), np.int32)
        cnt = np.zeros(n, np.int32)
        bump = np.zeros((n, 2), np.int32)
        orc.orc_project(C.byref(pp), n, H.ptr(dirs), H.ptr(px), H.ptr(py), H.ptr(cnt), H.ptr(bump))
        assert np.array_equal(cnt, g[f"{name}.cnt"]), name
        assert np.array_equal(px, g[f"{name}.px"]) and np.array_equal(py, g[f"{name}.py"]), name
        assert np.array_equal(bump, g[f"{name}.bump"]), name


def test_post_snapshot_golden():
    """Display sink: orc_post_snapshot == the reference's PostSnapshot loop over its own colour functions
    (fixture from oracle/make_golden.py gen_snapshot), byte for byte."""
    import sys
    sys.path.insert(0, os.path.join(H.ROOT, "oracle"))
    import make_golden as MG
    orc = H.oracle()
    g = np.load(os.path.join(G, "snapshot.npz"))
    xyz = np.ascontiguousarray(g["xyz"])
    for i, (fac, rc, bg) in enumerate(MG.SNAPSHOT_VARIANTS):
        rgb = np.zeros(xyz.shape, np.uint8)
        rc_a, bg_a = np.array(rc, np.float32), np.array(bg, np.float32)  # keep alive across the call
        orc.orc_post_snapshot(H.ptr(xyz), xyz.shape[1], xyz.shape[0], float(g["intensity"]), fac, H.ptr(rc_a),
                              H.ptr(bg_a), H.ptr(rgb))
        assert np.array_equal(rgb, g[f"rgb{i}"]), i
        assert rgb.max() >= 254 and (i != 0 or 0 < (rgb == 0).mean() < 0.9)   # clipping and the dark end
    # zero intensity -> black frame (render.cpp:514-517)
    orc.orc_post_snapshot(H.ptr(xyz), xyz.shape[1], xyz.shape[0], 0.0, 1.0, H.ptr(rc_a), H.ptr(bg_a), H.ptr(rgb))
    assert not rgb.any()


def test_colour_component_masks_golden():
    """Raypath colour: the oracle's colour pass (colour groups built by the product's host builder) reproduces
    the component masks the reference CpuTraceBackend assigned to the same injected rays (fixture)."""
    import sys
    sys.path.insert(0, os.path.join(H.ROOT, "oracle"))
    import make_golden as MG
    import parity
    from ice_halo_sim_b200 import backend as B
    g = np.load(os.path.join(G, "color_masks.npz"))
    pop = MG.color_population()
    tables = B.SceneTables(parity.scene([(0.0, [pop])], 6), 1)
    gpop = tables.scene().layers[0].populations[0]
    t = gpop.shapes[0]
    ex, er = oracle_trace_single(t, np.float32(1.31), 6, g["d"], g["p"], g["w"], g["face"], pop=gpop)
    a, ar = H.sort_exits(ex, er)
    assert np.array_equal(ar, g["exit_root"]) and np.array_equal(a["path_len"], g["exit_len"])
    assert np.array_equal(a["path"][:, :8], g["exit_path"])
    assert np.array_equal(a["weight"].view(np.uint32), g["exit_w"].view(np.uint32))
    assert np.array_equal(a["component_mask"], g["exit_mask"])
    assert bin(int(np.bitwise_or.reduce(a["component_mask"]))).count("1") >= 8


def test_rng_and_feistel_properties():
    orc = H.oracle()
    # pcg_hash known values computed from the definition (pcg_shared.h:193-197)
    def ref_hash(x):
        x = (x * 747796405 + 2891336453) & 0xFFFFFFFF
        x = ((((x >> ((x >> 28) + 4)) ^ x) & 0xFFFFFFFF) * 277803737) & 0xFFFFFFFF
        return ((x >> 22) ^ x) & 0xFFFFFFFF
    for x in (0, 1, 42, 0xDEADBEEF, 0xFFFFFFFF):
        assert orc.orc_pcg_hash(x) == ref_hash(x)
    u = np.array([orc.orc_draw(42, i, 3) for i in range(20000)])
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.01
    # feistel_bijection is a permutation of [0, n) for awkward n (pcg_shared.h:550-603)
    for n in (1, 2, 3, 5, 17, 1000, 4097):
        perm = sorted(orc.orc_feistel(i, n, 12345) for i in range(n))
        assert perm == list(range(n))


def test_entry_sampling_matches_analytic_distribution():
    """Entry triangle / point sampling (two-level face-group categorical) against the analytic law the
    reference samples from (InitRay_p_fid, simulator.cpp:133-192; test/support/incidence_sampling_oracle.hpp):
    P(face f | d) = max(-d.n_f, 0) A_f / sum_g max(-d.n_g, 0) A_g and a uniform point on the face, i.e.
    (a) every entry point lies on its face and inside the crystal, (b) face counts agree with the summed
    per-ray probabilities within 4.5 sigma, (c) the mean entry point of a face is the polygon's centroid."""
    import ctypes as C
    import harness as H
    import parity
    from ice_halo_sim_b200 import backend as B
    A = H.A
    orc = H.oracle()
    for pop in (parity.prism_pop(1.3, zenith=("gauss", 90, 0.3)), parity.pyramid_pop(),
                parity.prism_pop(0.4, zenith=("uniform", 90, 360), face_dist=[parity.dist("none", c) for c in (1.0, 0.7, 1.2, 0.9, 1.1, 0.8)])):
        tables = B.SceneTables(parity.scene([(0.0, [pop])], 3), 1)
        t = tables.scene().layers[0].populations[0].shapes[0]
        nf, nt = t.face_cnt, t.subtri_cnt
        planes = np.ctypeslib.as_array(t.plane)[:nf].astype(np.float64)
        tri_v = np.ctypeslib.as_array(t.tri_v)[:nt].astype(np.float64).reshape(nt, 3, 3)
        tri_area = np.ctypeslib.as_array(t.tri_area)[:nt].astype(np.float64)
        tri_face = np.ctypeslib.as_array(t.tri_face)[:nt]
        n = 200000
        wl = A.HbWlEntry(1.31, 1.0, 0, 0, 0)
        d = np.zeros((n, 3), np.float32); p = np.zeros((n, 3), np.float32); w = np.zeros(n, np.float32)
        f = np.zeros(n, np.uint16)
        orc.orc_gen_roots(tables.scene_ptr, 0, 0, 0, C.byref(wl), 1, 99, 0, n, H.ptr(d), H.ptr(p), H.ptr(w), H.ptr(f),
                          None, None, None, None)
        assert (f < nf).all()
        # (a) on the face, inside all other half-spaces
        s = p.astype(np.float64) @ planes[:, :3].T + planes[:, 3]
        assert np.abs(s[np.arange(n), f]).max() < 2e-6
        assert s.max() < 2e-6
        # (b) face frequencies
        area_f = np.array([tri_area[tri_face == k].sum() for k in range(nf)])
        wgt = np.maximum(-(d.astype(np.float64) @ planes[:, :3].T), 0.0) * area_f
        prob = wgt / wgt.sum(axis=1, keepdims=True)
        exp = prob.sum(axis=0)
        sig = np.sqrt((prob * (1 - prob)).sum(axis=0)) + 1e-9
        cnt = np.bincount(f, minlength=nf)[:nf]
        assert (np.abs(cnt - exp) < 4.5 * sig + 1).all(), (cnt, exp, sig)
        # (c) centroid of the entry points of each well-populated face = area-weighted triangle centroid
        for k in range(nf):
            sel = f == k
            if sel.sum() < 3000:
                continue
            tris = tri_face == k
            cen = (tri_v[tris].mean(axis=1) * tri_area[tris, None]).sum(axis=0) / tri_area[tris].sum()
            got = p[sel].astype(np.float64).mean(axis=0)
            spread = p[sel].astype(np.float64).std(axis=0).max()
            assert np.abs(got - cen).max() < 5.0 * spread / np.sqrt(sel.sum()) + 1e-6, (k, got, cen)


def test_sampler_twin_matches_reference_pcg_fixture():
    """tests/golden/sampler_pin.npz holds outputs of the reference's own counter-based sampler (lm_pcg::*,
    core/shared/pcg_shared.h) on fixed inputs (oracle/make_golden.py::gen_sampler_pin). The oracle's generator
    building blocks reproduce them: integers and uniforms exactly, libm-dependent floats to 2 ulp-scale tolerances
    (the fixture was written on the build container's glibc), the orientation matrix to 1e-6, Feistel exactly."""
    import ctypes as C
    g = np.load(os.path.join(H.GOLDEN, "sampler_pin.npz"))
    axes = []
    for k in g.files:
        if k.startswith("axis_"):
            ax = A.HbAxisSampler.from_buffer_copy(g[k].tobytes())
            axes.append((k[5:], ax))
    order = ["full_sphere", "fixed", "gauss_lut", "gauss_legacy", "laplacian", "zigzag", "uniform_band"]
    axes.sort(key=lambda kv: order.index(kv[0]))
    got = H.sampler_vectors(H.oracle(), "orc_", axes)
    for k, v in got.items():
        want = g[k]
        if v.dtype == np.uint32 or k == "uniforms":
            assert np.array_equal(v, want), k
        elif k.startswith("rot_"):
            assert np.abs(v - want).max() <= 1e-6, k
        else:
            assert np.allclose(v, want, rtol=3e-7, atol=3e-7), k
    orc = H.oracle()
    for k in g.files:
        if k.startswith("feistel_"):
            _, n, seed = k.split("_")
            n, seed = int(n), int(seed)
            assert np.array_equal(np.array([orc.orc_feistel(i, n, seed) for i in range(n)], np.uint32), g[k]), k
