"""The product's trace arithmetic, compiled for the CPU from the very source nvcc compiles (csrc/hb_device.cuh and
csrc/hb_tables.h with -DHB_HOST_TWIN; tests/host_twin/), against the golden vectors of the UNMODIFIED reference
(tests/golden/optics.npz: HitSurface / Propagate on four crystals) and against itself (every fused one-pass form of the
kernels equals the stand-alone functions it replaces). No GPU: this is the `-m "not gpu"` guard of the device path's
arithmetic; the GPU suite checks the compiled kernels."""
import ctypes as C
import os

import numpy as np
import pytest

import harness as H
from test_oracle_golden import bits, tables_from_golden

A = H.A
G = H.GOLDEN
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
NAMES = ["prism_h1", "column_h1p3", "pyramid_full", "prism_irregular"]


@pytest.fixture(scope="module")
def twin():
    return H.host_twin()


def derive(twin, t):
    planes = np.zeros((20, 4), np.float32)
    fn = np.zeros(20, np.uint8)
    axes = np.zeros((20, 2, 4), np.float32)
    meta = np.zeros(1, np.uint32)
    p4 = twin.twin_derive(C.byref(t), H.ptr(planes), H.ptr(fn), H.ptr(axes), H.ptr(meta))
    return planes, axes, int(meta[0]), bool(p4)


def propagate(twin, planes, axes, meta, mode, d, p, ff):
    n = len(ff)
    p_out = np.zeros((n, 3), np.float32)
    to = np.zeros(n, np.uint16)
    fe = np.zeros(n, np.uint8)
    twin.twin_propagate(H.ptr(planes), H.ptr(axes), meta, mode, n, H.ptr(d), H.ptr(p), H.ptr(ff), H.ptr(p_out), H.ptr(to),
                        H.ptr(fe))
    return p_out, to, fe


@pytest.mark.parametrize("name", NAMES)
def test_device_source_on_the_cpu_equals_the_reference_golden_vectors(twin, name):
    g = np.load(os.path.join(G, "optics.npz"))
    t = tables_from_golden(np.load(os.path.join(G, "crystal_tables.npz")), name)
    planes, axes, meta, p4 = derive(twin, t)
    assert p4 == (name != "pyramid_full")       # every hexagonal prism with its eight faces runs the unrolled forms
    # HitSurface (optics.cpp:18-53): weights of both children and the reflected direction everywhere, the refracted
    # direction wherever there is one
    d, w, face = g[f"{name}.hs_d"], g[f"{name}.hs_w"], g[f"{name}.hs_face"]
    n = len(w)
    d_out = np.zeros((n, 6), np.float32)
    w_out = np.zeros((n, 2), np.float32)
    n_idx = np.float32(1.31)
    twin.twin_hit_surface(H.ptr(planes), n_idx, np.float32(1.0) / n_idx, n, H.ptr(d), H.ptr(w), H.ptr(face), H.ptr(d_out),
                          H.ptr(w_out))
    assert np.array_equal(bits(w_out), bits(g[f"{name}.hs_wout"]))
    assert np.array_equal(bits(d_out[:, :3]), bits(g[f"{name}.hs_dout"][:, :3]))
    refr = g[f"{name}.hs_wout"][:, 1] >= 0
    assert refr.any() and (~refr).any()
    assert np.array_equal(bits(d_out[refr, 3:]), bits(g[f"{name}.hs_dout"][refr, 3:]))
    # Propagate (optics.cpp:64-158): hit face and advanced point, every form of the scan the kernels run
    p, ff = g[f"{name}.pr_p"], g[f"{name}.pr_from"]
    live = w >= 0
    want_to, want_p = g[f"{name}.pr_to"], g[f"{name}.pr_pout"]
    hit = live & (want_to != 0xFFFF)
    modes = [0, 1] + ([2, 3] if p4 else [])
    for mode in modes:
        p_out, to, _ = propagate(twin, planes, axes, meta, mode, d, p, ff)
        assert np.array_equal(to[live], want_to[live]), (name, mode)
        assert np.array_equal(bits(p_out[hit]), bits(want_p[hit])), (name, mode)
    # the one-pass forms of the fused bounce kernel need a source face: the half of the vectors that sits on one
    on_face = live & (ff != 0xFFFF)
    assert on_face.sum() > 100
    for mode in [4] + ([5] if p4 else []):
        p_out, to, _ = propagate(twin, planes, axes, meta, mode, d, p, ff)
        assert np.array_equal(to[on_face], want_to[on_face]), (name, mode)
        sel = on_face & hit
        assert np.array_equal(bits(p_out[sel]), bits(want_p[sel])), (name, mode)


@pytest.mark.parametrize("name", NAMES)
def test_one_pass_forms_equal_the_functions_they_fuse(twin, name):
    """bounce_axes / bounce_axes_p4 / last_axes_p4 against far_child_surely_exits(_p4), near_child_surely_hits and
    slab_exit(_p4) on rays that start ON a face (golden vectors + points pushed to within rounding of edges)."""
    g = np.load(os.path.join(G, "optics.npz"))
    t = tables_from_golden(np.load(os.path.join(G, "crystal_tables.npz")), name)
    planes, axes, meta, p4 = derive(twin, t)
    d, p, ff = g[f"{name}.hs_d"].copy(), g[f"{name}.pr_p"].copy(), g[f"{name}.pr_from"].copy()
    on = ff != 0xFFFF
    d, p, ff = d[on], p[on], ff[on]
    # adversarial copies: slide the start point along the ray to (almost) its exit point, i.e. next to an edge
    p_exit, to, _ = propagate(twin, planes, axes, meta, 0, d, p, ff)
    near_edge = to != 0xFFFF
    mix = np.float32(1.0) - np.float32(1e-6) * np.arange(near_edge.sum(), dtype=np.float32)[:, None] % np.float32(3e-5)
    p2 = (p[near_edge] + (p_exit[near_edge] - p[near_edge]) * mix).astype(np.float32)
    d = np.concatenate([d, d[near_edge]])
    p = np.concatenate([p, p2])
    ff = np.concatenate([ff, ff[near_edge]])
    n = len(ff)
    quick = np.zeros((n, 5), np.uint8)
    twin.twin_quick_tests(H.ptr(planes), H.ptr(axes), meta, n, H.ptr(d), H.ptr(p), H.ptr(ff), H.ptr(quick))
    ref_p, ref_to, _ = propagate(twin, planes, axes, meta, 0, d, p, ff)
    gp, gto, gfe = propagate(twin, planes, axes, meta, 4, d, p, ff)
    assert np.array_equal(gto, ref_to) and np.array_equal(bits(gp), bits(ref_p))
    # the quick far-child verdict may only ever say "leaves" when the full scan of that direction finds no face
    full_p, full_to, _ = propagate(twin, planes, axes, meta, 1, d, p, ff)
    for verdict in (gfe, quick[:, 0]):
        assert not (verdict.astype(bool) & (full_to != 0xFFFF)).any()
    # ... and "surely hits a face" only when it does
    assert not (quick[:, 2].astype(bool) & (ref_to == 0xFFFF)).any()
    if p4:
        pp, pto, pfe = propagate(twin, planes, axes, meta, 5, d, p, ff)
        assert np.array_equal(pto, ref_to) and np.array_equal(bits(pp), bits(ref_p))
        assert np.array_equal(pfe, quick[:, 1])                       # bounce_axes_p4 == far_child_surely_exits_p4
        assert np.array_equal(quick[:, 3], quick[:, 1])               # last_axes_p4 far == the same
        assert np.array_equal(quick[:, 4], quick[:, 2])               # last_axes_p4 near == near_child_surely_hits
        assert not (quick[:, 1].astype(bool) & (full_to != 0xFFFF)).any()
        assert quick[:, 1].any() and quick[:, 2].any()                # the fast verdicts actually fire
    assert gfe.any()


def test_device_projection_source_equals_the_reference_golden_pixels(twin):
    """project_exit as the kernels compile it, all 11 lens types (12 golden render set-ups of the reference's
    ProjectExitToPixel, projection_shared.h:196-375): pixel coordinates, hit counts and landed-weight flags."""
    g = np.load(os.path.join(G, "projection.npz"))
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
        twin.twin_project(C.byref(pp), n, H.ptr(dirs), H.ptr(px), H.ptr(py), H.ptr(cnt), H.ptr(bump))
        assert np.array_equal(cnt, g[f"{name}.cnt"]), name
        assert np.array_equal(px, g[f"{name}.px"]) and np.array_equal(py, g[f"{name}.py"]), name
        assert np.array_equal(bump, g[f"{name}.bump"]), name
        # the emission's early cull never drops a direction the lens would have drawn, and does fire where it applies
        culled = np.zeros(n, np.uint8)
        twin.twin_project_culls(C.byref(pp), n, H.ptr(dirs), H.ptr(culled))
        assert not (culled.astype(bool) & (cnt > 0)).any(), name
        if pp.proj_type in (0, 1, 2, 3, 8) and (cnt == 0).any():   # the single-hemisphere lenses (halotrace_b200.h: HB_LENS_*)
            assert culled.any(), name


def test_device_sampler_source_equals_the_reference_pcg_fixture(twin):
    """The generator's building blocks as the kernels compile them (pcg_hash, seed_with_high, uniforms, get_dist of
    all six types, sample_lon_lat_roll on every latitude path incl. the slots it consumes, sample_sph_cap, the
    point-in-triangle step, feistel) against tests/golden/sampler_pin.npz = outputs of the reference's own lm_pcg::*
    (core/shared/pcg_shared.h): integers and uniforms exactly, libm-dependent floats to ulp-scale tolerances, the
    orientation matrix (quaternion here, three axis rotations there) to 1e-6."""
    g = np.load(os.path.join(G, "sampler_pin.npz"))
    order = ["full_sphere", "fixed", "gauss_lut", "gauss_legacy", "laplacian", "zigzag", "uniform_band"]
    axes = sorted(((k[5:], A.HbAxisSampler.from_buffer_copy(g[k].tobytes())) for k in g.files if k.startswith("axis_")),
                  key=lambda kv: order.index(kv[0]))
    assert len(axes) == len(order)
    got = H.sampler_vectors(twin, "twin_", axes)
    for k, v in got.items():
        want = g[k]
        if v.dtype == np.uint32 or k == "uniforms":
            assert np.array_equal(v, want), k
        elif k.startswith("rot_"):
            assert np.abs(v - want).max() <= 1e-6, k
        else:
            assert np.allclose(v, want, rtol=3e-7, atol=3e-7), k
    for k in g.files:
        if k.startswith("feistel_"):
            _, n, seed = k.split("_")
            n, seed = int(n), int(seed)
            assert np.array_equal(np.array([twin.twin_feistel(i, n, seed) for i in range(n)], np.uint32), g[k]), k


def test_unchecked_division_as_compiled_for_the_host_is_ieee(twin):
    rng = np.random.default_rng(5)
    n = 2_000_000
    a = (rng.standard_normal(n) * np.exp(rng.uniform(-20, 20, n))).astype(np.float32)
    b = np.exp(rng.uniform(-20, 20, n)).astype(np.float32)
    a[::17] = 0.0
    a[1::34] = -0.0
    assert twin.twin_div_mismatches(n, H.ptr(a), H.ptr(b)) == 0
    assert twin.twin_div_mismatches(n, H.ptr(a), H.ptr(-b)) <= (a == 0).sum()   # zero numerators over b < 0: sign only
    x = np.exp(rng.uniform(-40, 40, n)).astype(np.float32)
    assert twin.twin_sqrt_mismatches(n, H.ptr(x)) == 0
