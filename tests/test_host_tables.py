"""Host-side table builders of the product library (hb_make_prism / pyramid / axis sampler / proj params /
wl entry / partition — no GPU) against tables produced by the reference's own host code (fixtures)."""
import ctypes as C
import os

import numpy as np
import pytest

import harness as H
from ice_halo_sim_b200 import lib as L

A = H.A
G = H.GOLDEN


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


SHAPES = ["prism_h1", "column_h1p3", "plate_h0p3", "prism_irregular", "pyramid_full", "pyramid_upper_apex",
          "pyramid_irregular"]


def build(shape_row):
    lib = L.load()
    t = A.HbCrystalTables()
    dist = (C.c_float * 6)(*shape_row[6:12])
    if shape_row[0] == 0:
        rc = lib.hb_make_prism(float(shape_row[3]), dist, C.byref(t))
    else:
        rc = lib.hb_make_pyramid(float(shape_row[1]), float(shape_row[2]), float(shape_row[3]), float(shape_row[4]),
                                 float(shape_row[5]), dist, C.byref(t))
    assert rc == 0
    return t


@pytest.mark.parametrize("name", SHAPES)
def test_crystal_tables(name):
    """Planes (unit normal + d0) and face numbers bit-identical to Crystal::GetPolygonFaceNormal/Dist/GetFn;
    the entry fan table covers the same surface (per-face area + normals) as BuildEntrySubTris."""
    g = np.load(os.path.join(G, "crystal_tables.npz"))
    t = build(g[f"{name}.shape"])
    fc = int(g[f"{name}.face_cnt"])
    assert t.face_cnt == fc
    assert np.array_equal(bits(np.ctypeslib.as_array(t.plane)[:fc]), bits(g[f"{name}.plane"][:fc]))
    assert list(t.face_fn[:fc]) == list(g[f"{name}.face_fn"][:fc])
    tc, rc = t.subtri_cnt, int(g[f"{name}.subtri_cnt"])
    area, face = np.ctypeslib.as_array(t.tri_area)[:tc], np.ctypeslib.as_array(t.tri_face)[:tc]
    r_area, r_face = g[f"{name}.tri_area"][:rc], g[f"{name}.tri_face"][:rc]
    for f in range(fc):
        assert abs(area[face == f].sum() - r_area[r_face == f].sum()) < 2e-6, (name, f)
    # every fan triangle lies in its face plane with the outward winding normal, area = |cross| / 2
    planes = np.ctypeslib.as_array(t.plane)
    tv, tn = np.ctypeslib.as_array(t.tri_v)[:tc], np.ctypeslib.as_array(t.tri_n)[:tc]
    for k in range(tc):
        pl = planes[face[k]]
        for c in range(3):
            assert abs(tv[k, c * 3:c * 3 + 3] @ pl[:3] + pl[3]) < 2e-6
        if area[k] > 1e-7:
            assert tn[k] @ pl[:3] > 0.999
        e1, e2 = tv[k, 3:6] - tv[k, 0:3], tv[k, 6:9] - tv[k, 0:3]
        assert abs(np.linalg.norm(np.cross(e1, e2)) / 2 - area[k]) < 1e-6


def test_degenerate_crystals_are_empty():
    lib = L.load()
    t = A.HbCrystalTables()
    one = (C.c_float * 6)(1, 1, 1, 1, 1, 1)
    assert lib.hb_make_prism(0.0, one, C.byref(t)) == 0 and t.face_cnt == 0 and t.subtri_cnt == 0
    bad = (C.c_float * 6)(1, 1, 1, -2, -2, -2)  # opposite-pair sums <= 0: empty cross-section
    assert lib.hb_make_prism(1.0, bad, C.byref(t)) == 0 and t.face_cnt == 0
    assert lib.hb_make_pyramid(28.0, 28.0, 0.0, 0.0, 0.0, one, C.byref(t)) == 0 and t.face_cnt == 0


def test_axis_samplers_and_lat_lut_bit_exact():
    lib = L.load()
    g = np.load(os.path.join(G, "axis_samplers.npz"))
    names = sorted({k.split(".")[0] for k in g.files})
    assert len(names) == 8
    for name in names:
        a = g[f"{name}.args"]
        s = A.HbAxisSampler()
        assert lib.hb_make_axis_sampler(int(a[0]), float(a[1]), float(a[2]), int(a[3]), float(a[4]), float(a[5]),
                                        int(a[6]), float(a[7]), float(a[8]), C.byref(s)) == 0
        assert [s.lat_path, s.az_type, s.roll_type, s.lut_n] == list(g[f"{name}.scalars"]), name
        fl = np.array([s.lat_mean, s.lat_std, s.az_mean, s.az_std, s.roll_mean, s.roll_std], np.float32)
        assert np.array_equal(bits(fl), bits(g[f"{name}.floats"])), name
        lut = np.stack([np.ctypeslib.as_array(s.lut_theta), np.ctypeslib.as_array(s.lut_cdf),
                        np.ctypeslib.as_array(s.lut_flip)])
        assert np.array_equal(bits(lut), bits(g[f"{name}.lut"])), name
        if s.lut_n:
            assert np.all(np.diff(lut[1]) > 0)   # strictly increasing CDF (binary search contract)


def test_proj_params_bit_exact():
    lib = L.load()
    g = np.load(os.path.join(G, "projection.npz"))
    names = sorted({k.split(".")[0] for k in g.files if "." in k})
    for name in names:
        rd = A.HbRenderDesc.from_buffer_copy(g[f"{name}.desc"].tobytes())
        pp = A.HbProjParams()
        assert lib.hb_build_render(C.byref(rd), C.byref(pp)) == 0
        assert bytes(pp) == g[f"{name}.params"].tobytes(), name


def test_wl_entries_and_sellmeier_bit_exact():
    lib = L.load()
    g = np.load(os.path.join(G, "optics.npz"))
    for wl, n, e in zip(g["sellmeier.wl"], g["sellmeier.n"], g["wl_entries"]):
        assert lib.hb_ice_refractive_index(float(wl)) == n
        m = A.HbWlEntry()
        assert lib.hb_make_wl_entry(float(wl), 1.0, C.byref(m)) == 0
        assert np.array_equal(bits(np.array([m.n_idx, m.spd_weight, m.cmf_x, m.cmf_y, m.cmf_z], np.float32)), bits(e))


def test_illuminant_wl_pools_bit_exact():
    """hb_make_wl_pool_illuminant == ComputeWlPool(illuminant_mode) for D50/D55/D65/D75/A/E (wl_pool.hpp:73-84)."""
    lib = L.load()
    g = np.load(os.path.join(G, "wl_pools.npz"))
    for key in g.files:
        ill, m = int(key[3]), int(key.split("_m")[1])
        pool = (A.HbWlEntry * m)()
        assert lib.hb_make_wl_pool_illuminant(ill, m, pool) == 0
        mine = np.frombuffer(pool, np.float32).reshape(m, 5)
        assert np.array_equal(bits(mine), bits(g[key])), key
    assert lib.hb_make_wl_pool_illuminant(9, 64, pool) != 0 and lib.hb_make_wl_pool_illuminant(2, 0, pool) != 0
    assert lib.hb_illuminant_spd(2, 299.0) == 0.0 and lib.hb_illuminant_spd(5, 500.0) == 1.0
    from ice_halo_sim_b200 import make_wl_pool
    assert len(make_wl_pool("D65")) == 64


def partition(prop, n, carry):
    lib = L.load()
    p = np.array(prop, np.float32)
    c = np.array(carry, np.float64)
    out = np.zeros(len(prop), np.uint64)
    assert lib.hb_partition_rays(p.ctypes.data, len(prop), n, c.ctypes.data, out.ctypes.data) == 0
    return out, c


def test_partition_rays():
    """PartitionCrystalRayNum scenarios (reference test_simulator.cpp:38-233): exact totals, proportionality,
    cross-batch carry, zero / negative proportions, empty inputs."""
    out, carry = partition([1, 1, 1], 100, [0, 0, 0])
    assert out.sum() == 100 and sorted(out) == [33, 33, 34]
    # carry makes the long-run split exact: 3 batches of 100 at 1:1:1 give 100 each
    tot = np.zeros(3, np.uint64)
    carry = [0.0, 0.0, 0.0]
    for _ in range(3):
        o, carry = partition([1, 1, 1], 100, carry)
        assert o.sum() == 100
        tot += o
    assert list(tot) == [100, 100, 100]
    out, _ = partition([10, 0, -5, 30], 1000, [0] * 4)
    assert list(out) == [250, 0, 0, 750]
    out, _ = partition([0, 0], 10, [0, 0])
    assert list(out) == [0, 0]
    out, _ = partition([1, 2], 0, [0, 0])
    assert list(out) == [0, 0]
    out, _ = partition([1], 7, [0])
    assert list(out) == [7]
    # tiny batches with uneven proportions still sum exactly, long run converges to the proportions
    carry = [0.0, 0.0, 0.0]
    tot = np.zeros(3, np.uint64)
    for _ in range(1000):
        o, carry = partition([0.7, 0.2, 0.1], 3, carry)
        assert o.sum() == 3
        tot += o
    assert abs(int(tot[0]) - 2100) <= 2 and abs(int(tot[1]) - 600) <= 2 and abs(int(tot[2]) - 300) <= 2
    if H.have_ref():
        rng = np.random.default_rng(3)
        for _ in range(50):
            k = int(rng.integers(1, 8))
            prop = rng.uniform(-0.2, 1.0, k).astype(np.float32)
            n = int(rng.integers(0, 5000))
            c0 = rng.uniform(-0.5, 0.5, k)
            mine, c_m = partition(prop, n, c0)
            r_out = np.zeros(k, np.uint64)
            c_r = c0.copy()
            H.ref().ref_partition(prop.ctypes.data, k, n, c_r.ctypes.data, r_out.ctypes.data)
            assert np.array_equal(mine, r_out) and np.allclose(c_m, c_r)
