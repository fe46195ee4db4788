"""Live cross-checks of the oracle restatement against the UNMODIFIED reference CPU core
(oracle/_ref/libhalo_ref.so, built from /root/reference by oracle/Makefile). Skipped where that library
is absent; the committed fixtures (test_oracle_golden.py) cover the same ground there."""
import ctypes as C

import numpy as np
import pytest

import harness as H
import parity
from test_oracle_golden import oracle_trace_single

A = H.A
pytestmark = pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built (no /root/reference here)")


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def random_shape(rng):
    dist = tuple(float(x) for x in rng.uniform(0.75, 1.25, 6))
    if rng.random() < 0.5:
        return H.ref_shape(0, (float(rng.uniform(0.2, 2.5)), 0, 0), dist)
    return H.ref_shape(1, (float(rng.uniform(0.1, 1.0)), float(rng.uniform(0.1, 1.5)), float(rng.uniform(0.0, 1.0))),
                       dist, (float(rng.uniform(15, 60)), float(rng.uniform(15, 60))))


def roots_on_crystal(rng, t, n):
    """Random entry rays: a point in a random fan triangle, direction facing into that face."""
    tri = rng.integers(0, t.subtri_cnt, n)
    tv = np.ctypeslib.as_array(t.tri_v)
    tn = np.ctypeslib.as_array(t.tri_n)
    tf = np.ctypeslib.as_array(t.tri_face)
    u, v = rng.random(n), rng.random(n)
    flip = u + v > 1
    u[flip], v[flip] = 1 - u[flip], 1 - v[flip]
    a, b, c = tv[tri, 0:3], tv[tri, 3:6], tv[tri, 6:9]
    p = (a + u[:, None] * (b - a) + v[:, None] * (c - a)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    facing = np.sum(d * tn[tri], axis=1) > 0
    d[facing] *= -1
    return d.astype(np.float32), p, np.ones(n, np.float32), tf[tri].astype(np.uint16)


def test_random_crystals_traced_bit_exact():
    """Random prisms and pyramids (irregular face distances), random entry rays, max_hits up to 12:
    reference CpuTraceBackend vs oracle, everything bit-exact."""
    ref = H.ref()
    rng = np.random.default_rng(2024)
    total = 0
    for trial in range(12):
        sh = random_shape(rng)
        t = A.HbCrystalTables()
        ref.ref_make_tables(C.byref(sh), C.byref(t))
        if t.face_cnt < 4:
            continue
        n = 1500
        mh = int(rng.integers(1, 13))
        d, p, w, f = roots_on_crystal(rng, t, n)
        n_idx = np.float32(rng.uniform(1.30, 1.33))
        cap = n * (mh + 2)
        ex = np.zeros(cap, H.EXIT_DTYPE)
        er = np.zeros(cap, np.uint32)
        ec = C.c_uint64()
        assert ref.ref_trace_injected(C.byref(sh), float(n_idx), mh, n, H.ptr(d), H.ptr(p), H.ptr(w), H.ptr(f), cap,
                                      H.ptr(ex), H.ptr(er), C.byref(ec)) == 0
        o_ex, o_er = oracle_trace_single(t, n_idx, mh, d, p, w, f)
        a, ar = H.sort_exits(ex[: ec.value], er[: ec.value])
        b, br = H.sort_exits(o_ex, o_er)
        assert len(a) == len(b) and np.array_equal(ar, br)
        assert np.array_equal(a["path_len"], b["path_len"]) and np.array_equal(a["path"], b["path"])
        assert np.array_equal(bits(a["dir"]), bits(b["dir"])) and np.array_equal(bits(a["weight"]), bits(b["weight"]))
        total += len(a)
    assert total > 20000


def test_adversarial_edge_and_vertex_rays_bit_exact():
    """Rays entering exactly on vertices / edges and refracted onto vertices / edges of other faces (ties between
    planes, t ~ 0 candidates, near-edge double continuation): reference CpuTraceBackend vs oracle, bit-exact."""
    ref = H.ref()
    rng = np.random.default_rng(77)
    shapes = [H.ref_shape(0, (1.3, 0, 0), (1.0,) * 6), H.ref_shape(0, (0.3, 0, 0), (1.0,) * 6),
              H.ref_shape(1, (0.3, 0.5, 0.4), (1.0,) * 6, (28.0, 28.0))] + [random_shape(rng) for _ in range(5)]
    total = 0
    for sh in shapes:
        t = A.HbCrystalTables()
        ref.ref_make_tables(C.byref(sh), C.byref(t))
        if t.face_cnt < 4:
            continue
        n, mh = 1500, 8
        n_idx = np.float32(1.3110129)
        d, p, w, f = H.adversarial_roots(rng, t, n, n_idx)
        cap = n * (mh + 2) * 2
        ex = np.zeros(cap, H.EXIT_DTYPE)
        er = np.zeros(cap, np.uint32)
        ec = C.c_uint64()
        assert ref.ref_trace_injected(C.byref(sh), float(n_idx), mh, n, H.ptr(d), H.ptr(p), H.ptr(w), H.ptr(f), cap,
                                      H.ptr(ex), H.ptr(er), C.byref(ec)) == 0
        o_ex, o_er = oracle_trace_single(t, n_idx, mh, d, p, w, f)
        a, ar = H.sort_exits(ex[: ec.value], er[: ec.value])
        b, br = H.sort_exits(o_ex, o_er)
        assert len(a) == len(b) and np.array_equal(ar, br)
        assert np.array_equal(a["path_len"], b["path_len"]) and np.array_equal(a["path"], b["path"])
        assert np.array_equal(bits(a["dir"]), bits(b["dir"])) and np.array_equal(bits(a["weight"]), bits(b["weight"]))
        total += len(a)
    assert total > 20000


def test_accumulate_matches_scatter_outgoing_to_xyz():
    """orc_accumulate == ScatterOutgoingToXyz (sequential fp32 adds in the same order => bit-exact)."""
    import sys, os
    sys.path.insert(0, os.path.join(H.ROOT, "oracle"))
    import make_golden as G
    ref, orc = H.ref(), H.oracle()
    rng = np.random.default_rng(5)
    n = 20000
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    w = rng.uniform(0, 1, n).astype(np.float32)
    for name in ("fisheye_ea_1080p", "linear_shift", "dual_ea_overlap", "rectangular_full", "globe"):
        rd = G.render_desc(**G.RENDERS[name])
        pp = A.HbProjParams()
        ref.ref_make_proj_params(C.byref(rd), C.byref(pp))
        wl = A.HbWlEntry()
        ref.ref_wl_entry(610.0, 1.0, C.byref(wl))
        img_r = np.zeros((rd.img_h, rd.img_w, 3), np.float32)
        landed_r = C.c_float(0)
        ref.ref_scatter_xyz(C.byref(rd), 610.0, n, H.ptr(d), H.ptr(w), H.ptr(img_r), C.byref(landed_r))
        img_o = np.zeros_like(img_r)
        landed_o = C.c_double(0)
        orc.orc_accumulate(C.byref(pp), C.byref(wl), 1, n, H.ptr(d), H.ptr(w), None, H.ptr(img_o), C.byref(landed_o))
        assert np.array_equal(bits(img_r), bits(img_o)), name
        assert abs(landed_r.value - landed_o.value) <= 1e-4 * max(1.0, landed_o.value), name


FILTERS = [
    dict(kind=1, path=[3, 5], symmetry=1),                 # examples/config_example.json filter 3
    dict(kind=1, path=[3, 1, 5, 7, 4], symmetry=7),        # filter 2 (PBD)
    dict(kind=1, path=[1, 3, 2], symmetry=0),
    dict(kind=1, path=[4, 2, 6], symmetry=2, action=1),
    dict(kind=2, entry=3, exit=5, symmetry=1),             # filter 4
    dict(kind=2, entry=1, exit=-1, symmetry=2, min_len=2, max_len=4),
    dict(kind=3, lon=180.0, lat=25.0, radii=30.0, action=1),  # filter 5 (wider cone so both outcomes occur)
    dict(kind=4, crystal_id=3),
    dict(kind=4, crystal_id=9),
    # complex: OR of AND-terms (examples/config_example.json filter 7 shape: [1, [2, 6], 5])
    dict(kind=5, sym="P", terms=[[dict(kind=1, path=[3, 5])], [dict(kind=1, path=[1, 3, 2]), dict(kind=4, crystal_id=3)],
                                 [dict(kind=3, lon=180.0, lat=25.0, radii=30.0)]]),
    dict(kind=5, sym="PBD", action=1, terms=[[dict(kind=2, entry=3, exit=5), dict(kind=1, path=[3, 1, 5])],
                                             [dict(kind=4, crystal_id=9)]]),
]


@pytest.mark.parametrize("spec", FILTERS)
@pytest.mark.parametrize("roll", [(1, 0.0, 360.0), (0, 30.0, 0.0), (2, 17.0, 3.0)])
def test_filter_check_matches_filter_spec(spec, roll):
    """BuildDeviceFilterDesc tables + the oracle's matcher == FilterSpec::Check on random raypaths, for every
    symmetry combination and D-applicability (roll mean at / off a multiple of 30 deg)."""
    ref, orc = H.ref(), H.oracle()
    import parity
    pop = parity.prism_pop(1.2, zenith=("gauss", 90, 1.0), roll=("uniform", 0, 360), cid=3)
    pop.crystal.roll = A.HbDist(*roll)
    if spec["kind"] == 5:
        f = parity.complex_filter(spec["terms"], spec.get("sym", ""), spec.get("action", 0))
    else:
        f = A.HbFilterSpecDesc()
        f.kind, f.action, f.symmetry = spec["kind"], spec.get("action", 0), spec.get("symmetry", 0)
        parity.simple_spec(f.simple, spec["kind"], spec.get("path", []), spec.get("entry", -1), spec.get("exit", -1),
                           spec.get("min_len", 1), spec.get("max_len", 0), spec.get("lon", 0.0), spec.get("lat", 0.0),
                           spec.get("radii", 0.0), spec.get("crystal_id", 0))
    pop.filter = f
    sh = H.ref_shape(0, (1.2, 0, 0))
    t = A.HbCrystalTables()
    ref.ref_make_tables(C.byref(sh), C.byref(t))
    # the product's own descriptor (hb_build_scene) must equal the reference's BuildDeviceFilterDesc
    rdesc = A.HbFilterDesc()
    ref.ref_filter_desc(C.byref(pop), C.byref(sh), C.byref(rdesc))
    sd = parity.scene([(0.0, [pop])], 7)
    from ice_halo_sim_b200 import backend as B
    tables = B.SceneTables(sd, 1)
    mine = tables.scene().layers[0].populations[0].filter
    for fld in ("kind", "action", "symmetry", "fn_period", "sigma_a", "d_applicable"):
        assert getattr(mine, fld) == getattr(rdesc, fld), fld
    def same_simple(a, b):
        assert a.kind == b.kind and a.path_len == b.path_len
        assert bytes(a.path)[: a.path_len] == bytes(b.path)[: b.path_len]
        assert (a.entry_fn >= 0) == (b.entry_fn >= 0) and (a.exit_fn >= 0) == (b.exit_fn >= 0)
        assert np.allclose(list(a.dir), list(b.dir), atol=1e-7) and abs(a.cos_radii - b.cos_radii) < 1e-7
        assert a.crystal_id == b.crystal_id
    if spec["kind"] == 5:
        assert mine.term_cnt == rdesc.term_cnt and list(mine.term_len) == list(rdesc.term_len)
        for o in range(mine.term_cnt):
            for a in range(mine.term_len[o]):
                same_simple(mine.terms[o][a], rdesc.terms[o][a])
    else:
        same_simple(mine.simple, rdesc.simple)
    # random paths (compact ids 0..7 == face numbers 1..8 on a prism) incl. the filter's own path and its images
    rng = np.random.default_rng(99)
    n = 4000
    plen = rng.integers(1, 7, n).astype(np.uint8)
    paths = np.zeros((n, 64), np.uint8)
    for i in range(n):
        paths[i, : plen[i]] = rng.integers(0, 8, plen[i])
    seed_paths = [spec["path"]] if spec["kind"] == 1 else \
        [t["path"] for term in spec.get("terms", []) for t in term if t["kind"] == 1]
    for base_fn in seed_paths:
        base = np.array(base_fn) - 1
        for i in range(0, 600):
            k = int(rng.integers(0, 6))
            img = np.where(base >= 2, (base - 2 + k) % 6 + 2, base)
            if rng.random() < 0.5:
                img = np.where(img < 2, 1 - img, img)
            paths[i, : len(img)] = img
            plen[i] = len(img)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    fn_paths = paths + 1   # reference recorder holds face numbers
    pass_r = np.zeros(n, np.uint8)
    ref.ref_filter_check(C.byref(pop), C.byref(sh), n, H.ptr(np.ascontiguousarray(fn_paths)), H.ptr(plen), H.ptr(d),
                         H.ptr(pass_r))
    pass_o = np.zeros(n, np.uint8)
    fn = np.ctypeslib.as_array(t.face_fn).copy()
    orc.orc_filter_check(C.byref(mine), H.ptr(fn), pop.crystal.id, n, H.ptr(paths), H.ptr(plen), H.ptr(d), H.ptr(pass_o))
    assert np.array_equal(pass_r, pass_o)
    assert 0 < pass_o.sum() < n or spec["kind"] == 4
    # ... and so does the matcher the kernels run (csrc/hb_filter.h compiled for the CPU, tests/host_twin)
    pass_t = np.zeros(n, np.uint8)
    H.host_twin().twin_filter_check(C.byref(mine), H.ptr(fn), pop.crystal.id, n, H.ptr(paths), H.ptr(plen), H.ptr(d),
                                    H.ptr(pass_t))
    assert np.array_equal(pass_r, pass_t)


def test_orientation_sampler_statistics_match_cpu_sampler():
    """Oracle root generation (counter-based PCG) vs the reference CPU sampler (mt19937): same orientation
    distribution. Compared through rotation-matrix moments; tolerance 4 sigma of the sample mean."""
    ref, orc = H.ref(), H.oracle()
    import parity
    for zen, az, roll in [(("gauss", 90, 0.3), ("uniform", 0, 360), ("uniform", 0, 360)),
                          (("uniform", 90, 360), ("uniform", 0, 360), ("uniform", 0, 360)),
                          (("laplacian", 10, 3.0), ("uniform", 0, 360), ("none", 30, 0)),
                          (("gauss", 0, 20.0), ("gauss", 40, 10.0), ("uniform", 0, 360))]:
        pop = parity.prism_pop(1.0, zenith=zen, azimuth=az, roll=roll)
        from ice_halo_sim_b200 import backend as B
        tables = B.SceneTables(parity.scene([(0.0, [pop])], 3), 1)
        n = 100000
        wl = A.HbWlEntry(1.31, 1.0, 0, 0, 0)
        d = np.zeros((n, 3), np.float32); p = np.zeros((n, 3), np.float32); w = np.zeros(n, np.float32)
        f = np.zeros(n, np.uint16); rot = np.zeros((n, 9), np.float32)
        orc.orc_gen_roots(tables.scene_ptr, 0, 0, 0, C.byref(wl), 1, 42, 0, n, H.ptr(d), H.ptr(p), H.ptr(w), H.ptr(f),
                          None, H.ptr(rot), None, None)
        c = pop.crystal
        llr = np.zeros((n, 3), np.float32)
        rot_r = np.zeros((n, 9), np.float32)
        ref.ref_sample_orientations(C.byref(c.latitude), C.byref(c.azimuth), C.byref(c.roll), 7, n, H.ptr(llr), H.ptr(rot_r))
        # orthonormal rotations
        m = rot.reshape(n, 3, 3)
        assert np.allclose(np.einsum("nij,nkj->nik", m, m), np.eye(3), atol=2e-6)
        for k in range(9):
            a, b = rot[:, k].astype(np.float64), rot_r[:, k].astype(np.float64)
            se = np.sqrt(a.var() / n + b.var() / n) + 1e-9
            assert abs(a.mean() - b.mean()) < 4.5 * se, (zen, k, a.mean(), b.mean())
            assert abs(np.mean(a * a) - np.mean(b * b)) < 4.5 * np.sqrt(np.var(a * a) / n + np.var(b * b) / n) + 1e-9


def test_post_snapshot_matches_reference_colour_pipeline():
    """orc_post_snapshot == PostSnapshot's loop over the reference's GamutClipXyz / XyzToLinearRgb / LinearToSrgb
    on a large random image, byte for byte (same libm on both sides)."""
    import sys, os
    sys.path.insert(0, os.path.join(H.ROOT, "oracle"))
    import make_golden as MG
    ref, orc = H.ref(), H.oracle()
    xyz, inten = MG.snapshot_input(seed=99, w=640, h=360)
    for fac, rc, bg in MG.SNAPSHOT_VARIANTS + [(0.25, (0.2, 1.0, 0.3), (0.5, 0.0, 0.0))]:
        rc_a, bg_a = np.array(rc, np.float32), np.array(bg, np.float32)
        a = np.zeros(xyz.shape, np.uint8)
        b = np.zeros(xyz.shape, np.uint8)
        ref.ref_post_snapshot(H.ptr(xyz), 640, 360, inten, fac, H.ptr(rc_a), H.ptr(bg_a), H.ptr(a))
        orc.orc_post_snapshot(H.ptr(xyz), 640, 360, inten, fac, H.ptr(rc_a), H.ptr(bg_a), H.ptr(b))
        assert np.array_equal(a, b) and a.any()


@pytest.mark.parametrize("roll", [(1, 0.0, 360.0), (0, 60.0, 0.0)])
def test_colour_component_masks_match_reference(roll):
    """Raypath colour: ExitRayRecord::component_mask from the reference CpuTraceBackend with a
    raypath_color config (one class per predicate => bit k) == the oracle's colour pass over the colour groups
    the product's host builder derives from the same predicates; rays injected, everything else bit-exact too."""
    import parity
    from ice_halo_sim_b200 import backend as B
    ref = H.ref()
    rng = np.random.default_rng(77)
    pop = parity.prism_pop(1.2, zenith=("gauss", 90, 1.0), cid=3)
    pop.crystal.roll = A.HbDist(*roll)
    preds = [("PBD", dict(kind=1, path=[3, 5])), ("PBD", dict(kind=2, entry=1, exit=3)), ("", dict(kind=0)),
             ("", dict(kind=1, path=[1, 3, 2])), ("P", dict(kind=1, path=[3, 1, 5])), ("B", dict(kind=2, entry=3, exit=-1)),
             ("", dict(kind=3, lon=10.0, lat=20.0, radii=60.0)), ("PBD", dict(kind=1, path=[4, 6])),
             ("P", dict(kind=2, entry=-1, exit=2, min_len=2, max_len=3)), ("", dict(kind=4, crystal_id=3))]
    for k, (sym, spec) in enumerate(preds):
        parity.color_pred(pop, k, sym, **spec)
    sh = H.ref_shape(0, (1.2, 0, 0))
    t = A.HbCrystalTables()
    ref.ref_make_tables(C.byref(sh), C.byref(t))
    n, mh = 4000, 6
    d, p, w, f = roots_on_crystal(rng, t, n)
    cap = n * (mh + 2)
    ex = np.zeros(cap, H.EXIT_DTYPE)
    er = np.zeros(cap, np.uint32)
    ec = C.c_uint64()
    assert ref.ref_trace_injected_color(C.byref(sh), 1.31, mh, n, H.ptr(d), H.ptr(p), H.ptr(w), H.ptr(f), C.byref(pop),
                                        cap, H.ptr(ex), H.ptr(er), C.byref(ec)) == 0
    tables = B.SceneTables(parity.scene([(0.0, [pop])], mh), 1)
    gpop = tables.scene().layers[0].populations[0]
    assert gpop.color_group_cnt == 4           # symmetry groups in first-occurrence order: PBD, none, P, B
    o_ex, o_er = oracle_trace_single(t, np.float32(1.31), mh, d, p, w, f, pop=gpop)
    a, ar = H.sort_exits(ex[: ec.value], er[: ec.value])
    b, br = H.sort_exits(o_ex, o_er)
    assert len(a) == len(b) and np.array_equal(ar, br) and np.array_equal(a["path"], b["path"])
    assert np.array_equal(bits(a["weight"]), bits(b["weight"]))
    assert np.array_equal(a["component_mask"], b["component_mask"])
    seen = int(np.bitwise_or.reduce(a["component_mask"]))
    assert seen & 0b100 and bin(seen).count("1") >= 8, bin(seen)      # whole-crystal bit + most predicates fire


def pin_axis_samplers():
    """One axis sampler per latitude path / distribution type, built by the reference's own host code."""
    cases = [("full_sphere", ("uniform", 90, 360), ("uniform", 0, 360), ("uniform", 0, 360)),
             ("fixed", ("none", 35, 0), ("gauss", 40, 10), ("gauss", 10, 5)),
             ("gauss_lut", ("gauss", 90, 0.3), ("uniform", 0, 360), ("uniform", 0, 360)),
             ("gauss_legacy", ("gauss_legacy", 80, 6), ("uniform", 0, 360), ("none", 30, 0)),
             ("laplacian", ("laplacian", 10, 3), ("laplacian", 90, 20), ("zigzag", 0, 15)),
             ("zigzag", ("zigzag", 90, 25), ("uniform", 0, 360), ("uniform", 0, 60)),
             ("uniform_band", ("uniform", 60, 40), ("zigzag", 0, 30), ("uniform", 0, 360))]
    out = []
    for name, zen, az, roll in cases:
        z = parity.dist(*zen)
        lat = A.HbDist(z.type, 90.0 - z.center, z.spread)
        ax = A.HbAxisSampler()
        assert H.ref().ref_make_axis_sampler(C.byref(lat), C.byref(parity.dist(*az)), C.byref(parity.dist(*roll)), C.byref(ax)) == 0
        out.append((name, ax))
    return out


def test_sampler_twin_matches_reference_pcg():
    """The oracle's generator building blocks against the reference's own counter-based sampler (lm_pcg::*,
    core/shared/pcg_shared.h, the code its GPU backends run; host-compiled into oracle/_ref): hashes, seeds, uniforms,
    every distribution type, sample_lat_lon_roll on every latitude path (angles AND the number of stream slots
    consumed), sample_sph_cap, sample_triangle bit-for-bit (same libm); the Feistel bijection and the categorical
    pick exactly; the orientation matrix (the oracle builds it from a quaternion, the reference from three axis
    rotations) to 1e-6."""
    orc, ref = H.oracle(), H.ref()
    axes = pin_axis_samplers()
    a = H.sampler_vectors(orc, "orc_", axes)
    b = H.sampler_vectors(ref, "ref_pcg_", axes)
    assert sorted(a) == sorted(b)
    for k in a:
        if k.startswith("rot_"):
            assert np.abs(a[k] - b[k]).max() <= 1e-6, k
            eye = np.einsum("nij,nkj->nik", b[k].reshape(-1, 3, 3), b[k].reshape(-1, 3, 3))
            assert np.abs(eye - np.eye(3)).max() < 1e-5
        else:
            assert a[k].dtype == b[k].dtype and np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
    assert len(np.unique(a["hash"])) > 4000 and 0.0 <= a["uniforms"].min() and a["uniforms"].max() < 1.0
    # Feistel bijection: same permutation for awkward sizes and seeds
    for n, seed in ((1, 5), (2, 5), (3, 9), (17, 1), (1000, 0xABCDEF), (65536, 42), (100003, 7)):
        want = np.zeros(n, np.uint32)
        ref.ref_pcg_feistel(n, seed, H.ptr(want))
        got = np.array([orc.orc_feistel(i, n, seed) for i in range(n)], np.uint32)
        assert np.array_equal(got, want), (n, seed)
        assert np.array_equal(np.sort(want), np.arange(n, dtype=np.uint32))
    # categorical_sample (the triangle-level entry pick the oracle falls back to) on random weights
    rng = np.random.default_rng(5)
    w = rng.random(20).astype(np.float32)
    w[[3, 7]] = 0.0
    u = np.concatenate([rng.random(2000).astype(np.float32), np.array([0.0, 0.99999994], np.float32)])
    pick = np.zeros(len(u), np.uint32)
    ref.ref_pcg_categorical(H.ptr(w), len(w), H.ptr(u), len(u), H.ptr(pick))
    cum = np.cumsum(w, dtype=np.float32)
    assert pick.max() < len(w) and not np.isin(pick, [3, 7]).any()
    freq = np.bincount(pick, minlength=len(w))[: len(w)] / len(u)
    assert np.abs(freq - w / w.sum()).max() < 0.03
