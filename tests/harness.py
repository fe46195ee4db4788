"""Shared test plumbing: loads the oracle (oracle/_build), the reference library (oracle/_ref, only
present where /root/reference was available at build time) and the product library, and wraps their
C entry points for numpy. TEST INFRASTRUCTURE — the product package never imports this."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from ice_halo_sim_b200 import _abi as A  # noqa: E402

ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libhalo_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libhalo_ref.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")

_vp = C.c_void_p


def ptr(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def build_oracle():
    src = os.path.join(ROOT, "oracle", "halo_oracle.cpp")
    if (not os.path.exists(ORACLE_SO)) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle"])


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        build_oracle()
        lib = C.CDLL(ORACLE_SO)
        lib.orc_pcg_hash.restype = C.c_uint32
        lib.orc_pcg_hash.argtypes = [C.c_uint32]
        lib.orc_draw.restype = C.c_float
        lib.orc_draw.argtypes = [C.c_uint32] * 3
        lib.orc_feistel.restype = C.c_uint32
        lib.orc_feistel.argtypes = [C.c_uint32] * 3
        lib.orc_gen_roots.argtypes = [_vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp, C.c_uint32, C.c_uint32,
                                      C.c_uint64, C.c_uint64] + [_vp] * 8
        lib.orc_transit.argtypes = [_vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64] + \
            [_vp] * 7
        lib.orc_trace_layer.argtypes = [_vp, C.c_uint64] + [_vp] * 7 + [C.c_uint64] + [_vp] * 8
        lib.orc_hit_surface.argtypes = [_vp, C.c_float, C.c_uint64] + [_vp] * 5
        lib.orc_propagate.argtypes = [_vp, C.c_uint64] + [_vp] * 6
        lib.orc_project.argtypes = [_vp, C.c_uint64] + [_vp] * 5
        lib.orc_accumulate.argtypes = [_vp, _vp, C.c_uint32, C.c_uint64] + [_vp] * 5
        lib.orc_accumulate_lanes.argtypes = [_vp, _vp, C.c_uint32, _vp, C.c_uint64] + [_vp] * 5
        lib.orc_post_snapshot.argtypes = [_vp, C.c_int, C.c_int, C.c_float, C.c_float, _vp, _vp, _vp]
        lib.orc_filter_check.argtypes = [_vp, _vp, C.c_uint32, C.c_uint64] + [_vp] * 4
        lib.orc_quat_to_rot9.argtypes = [_vp, _vp]
        _sampler_prototypes(lib, "orc_")
        _oracle = lib
    return _oracle


def _sampler_prototypes(lib, prefix):
    """The generator's building blocks, same signatures on both sides (oracle: orc_*, reference: ref_pcg_*)."""
    u32, f32 = C.c_uint32, C.c_float
    getattr(lib, prefix + "seed_with_high").restype = u32
    getattr(lib, prefix + "seed_with_high").argtypes = [u32, u32]
    getattr(lib, prefix + "uniforms").argtypes = [u32, u32, u32, u32, _vp]
    getattr(lib, prefix + "get_dist").argtypes = [u32, u32, u32, u32, f32, f32, _vp]
    getattr(lib, prefix + "lat_lon_roll").argtypes = [_vp, u32, u32, u32, _vp, _vp]
    getattr(lib, prefix + "rotation9").argtypes = [C.c_uint64, _vp, _vp]
    getattr(lib, prefix + "sph_cap").argtypes = [u32, u32, u32, u32, f32, f32, f32, _vp]
    getattr(lib, prefix + "triangle").argtypes = [u32, u32, u32, u32, _vp, _vp]


_twin = None


def host_twin():
    """The product's device arithmetic compiled for the CPU (tests/host_twin/: csrc/hb_device.cuh, hb_tables.h and
    hb_filter.h under g++ with -DHB_HOST_TWIN); built on first use."""
    global _twin
    if _twin is None:
        here = os.path.join(ROOT, "tests", "host_twin")
        out = os.path.join(here, "_build", "libhb_host_twin.so")
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-frounding-math", "-fPIC", "-shared",
                               "-DHB_HOST_TWIN=1", "-I" + os.path.join(ROOT, "include"),
                               "-I" + os.path.join(ROOT, "ice_halo_sim_b200", "csrc"), "-I" + here, "-o", out,
                               os.path.join(here, "hb_host_twin.cpp")])
        lib = C.CDLL(out)
        u32, f32, u64 = C.c_uint32, C.c_float, C.c_uint64
        lib.twin_derive.argtypes = [_vp] * 5
        lib.twin_hit_surface.argtypes = [_vp, f32, f32, u64] + [_vp] * 5
        lib.twin_propagate.argtypes = [_vp, _vp, u32, u32, u64] + [_vp] * 6
        lib.twin_quick_tests.argtypes = [_vp, _vp, u32, u64] + [_vp] * 4
        lib.twin_project.argtypes = [_vp, u64] + [_vp] * 5
        lib.twin_project_culls.argtypes = [_vp, u64, _vp, _vp]
        lib.twin_filter_check.argtypes = [_vp, _vp, u32, u64] + [_vp] * 4
        lib.twin_filter_max_len.restype = u32
        lib.twin_filter_max_len.argtypes = [_vp]
        lib.twin_div_mismatches.argtypes = [u64, _vp, _vp]
        lib.twin_div_mismatches.restype = u64
        lib.twin_sqrt_mismatches.argtypes = [u64, _vp]
        lib.twin_sqrt_mismatches.restype = u64
        lib.twin_pcg_hash.restype = u32
        lib.twin_pcg_hash.argtypes = [u32]
        lib.twin_feistel.restype = u32
        lib.twin_feistel.argtypes = [u32] * 3
        _sampler_prototypes(lib, "twin_")
        _twin = lib
    return _twin


def sampler_vectors(lib, prefix, axis_samplers):
    """Run every building block of the generator on fixed inputs; returns {name: array}. `axis_samplers`:
    [(name, HbAxisSampler)]."""
    out = {}
    n = 4096
    hash_fn = lib.ref_pcg_hash if prefix == "ref_pcg_" else getattr(lib, prefix + "pcg_hash")
    xs = np.concatenate([np.arange(0, 2048, dtype=np.uint64), np.linspace(0, 2 ** 32 - 1, 2048).astype(np.uint64)])
    out["hash"] = np.array([hash_fn(int(x)) for x in xs], np.uint32)
    out["seed_hi"] = np.array([getattr(lib, prefix + "seed_with_high")(0x1234ABCD, h) for h in (0, 1, 2, 77, 2 ** 31)], np.uint32)
    u = np.zeros(256, np.float32)
    getattr(lib, prefix + "uniforms")(0xC0FFEE, 123456789, 3, 256, ptr(u))
    out["uniforms"] = u
    for t in range(6):
        g = np.zeros(n, np.float32)
        getattr(lib, prefix + "get_dist")(99 + t, 5000, n, t, 0.3, 0.7, ptr(g))
        out[f"dist{t}"] = g
    for name, ax in axis_samplers:
        llr = np.zeros((n, 3), np.float32)
        slots = np.zeros(n, np.uint32)
        getattr(lib, prefix + "lat_lon_roll")(C.byref(ax), 0xBEEF, 4_000_000_000, n, ptr(llr), ptr(slots))
        rot = np.zeros((n, 9), np.float32)
        getattr(lib, prefix + "rotation9")(n, ptr(llr), ptr(rot))
        out[f"llr_{name}"], out[f"slots_{name}"], out[f"rot_{name}"] = llr, slots, rot
    d = np.zeros((n, 3), np.float32)
    getattr(lib, prefix + "sph_cap")(7, 100, 5, n, 3.3, -0.35, 0.00436, ptr(d))
    out["sph_cap"] = d
    vtx = np.array([0.1, -0.2, 0.65, 0.9, 0.3, 0.65, -0.4, 0.8, 0.65], np.float32)
    p = np.zeros((n, 3), np.float32)
    getattr(lib, prefix + "triangle")(8, 200, 9, n, ptr(vtx), ptr(p))
    out["triangle"] = p
    return out


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_SO)
        lib.ref_refractive_index.restype = C.c_double
        lib.ref_refractive_index.argtypes = [C.c_double]
        lib.ref_make_tables.argtypes = [_vp, _vp]
        lib.ref_make_axis_sampler.argtypes = [_vp] * 4
        lib.ref_make_proj_params.argtypes = [_vp, _vp]
        lib.ref_wl_entry.argtypes = [C.c_float, C.c_float, _vp]
        lib.ref_wl_pool_illuminant.argtypes = [C.c_int, C.c_uint32, _vp]
        lib.ref_cmf_table.argtypes = [_vp]
        lib.ref_daylight_basis.argtypes = [_vp]
        lib.ref_post_snapshot.argtypes = [_vp, C.c_int, C.c_int, C.c_float, C.c_float, _vp, _vp, _vp]
        lib.ref_filter_desc.argtypes = [_vp] * 3
        lib.ref_hit_surface.argtypes = [_vp, C.c_float, C.c_uint64] + [_vp] * 5
        lib.ref_propagate.argtypes = [_vp, C.c_uint64] + [_vp] * 6
        lib.ref_project.argtypes = [_vp, C.c_uint64] + [_vp] * 5
        lib.ref_scatter_xyz.argtypes = [_vp, C.c_float, C.c_uint64] + [_vp] * 4
        lib.ref_filter_check.argtypes = [_vp, _vp, C.c_uint64] + [_vp] * 4
        lib.ref_sample_orientations.argtypes = [_vp, _vp, _vp, C.c_uint32, C.c_uint64, _vp, _vp]
        lib.ref_partition.argtypes = [_vp, C.c_uint32, C.c_uint64, _vp, _vp]
        lib.ref_trace_injected.argtypes = [_vp, C.c_float, C.c_uint32, C.c_uint64] + [_vp] * 4 + [C.c_uint64] + \
            [_vp] * 3
        lib.ref_trace_injected_color.argtypes = [_vp, C.c_float, C.c_uint32, C.c_uint64] + [_vp] * 5 + [C.c_uint64] + \
            [_vp] * 3
        lib.ref_cpu_backend_run.argtypes = [_vp, _vp, C.c_float, C.c_float, C.c_uint32, C.c_uint64, C.c_uint64] + \
            [_vp] * 4
        lib.ref_legacy_bench.argtypes = [_vp, _vp, _vp, _vp, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32] + \
            [_vp] * 3
        lib.ref_physical_cores.restype = C.c_uint32
        _sampler_prototypes(lib, "ref_pcg_")
        lib.ref_pcg_hash.restype = C.c_uint32
        lib.ref_pcg_hash.argtypes = [C.c_uint32]
        lib.ref_pcg_feistel.argtypes = [C.c_uint32, C.c_uint32, _vp]
        lib.ref_pcg_categorical.argtypes = [_vp, C.c_uint32, _vp, C.c_uint32, _vp]
        _ref = lib
    return _ref


class RefShape(C.Structure):  # oracle/ref_driver.h
    _fields_ = [("kind", C.c_uint32), ("upper_alpha_deg", C.c_float), ("lower_alpha_deg", C.c_float),
                ("h1", C.c_float), ("h2", C.c_float), ("h3", C.c_float), ("dist", C.c_float * 6)]


def ref_shape(kind=0, h=(1.0, 0.0, 0.0), dist=(1, 1, 1, 1, 1, 1), alpha=(28.0, 28.0)):
    return RefShape(kind, alpha[0], alpha[1], h[0], h[1], h[2], (C.c_float * 6)(*dist))


def exits_to_numpy(buf, n):
    """(HbExitRecord * cap) -> structured numpy view of the first n records."""
    dt = np.dtype([("dir", np.float32, 3), ("weight", np.float32), ("path_len", np.uint8), ("path", np.uint8, 64),
                   ("pad0", np.uint8), ("crystal_id", np.uint16), ("ms_layer_idx", np.uint8), ("wl_idx", np.uint8),
                   ("pad1", np.uint8, 2), ("component_mask", np.uint64)])
    assert dt.itemsize == 96
    return np.frombuffer(buf, dtype=dt, count=n)


EXIT_DTYPE = np.dtype([("dir", np.float32, 3), ("weight", np.float32), ("path_len", np.uint8), ("path", np.uint8, 64),
                       ("pad0", np.uint8), ("crystal_id", np.uint16), ("ms_layer_idx", np.uint8), ("wl_idx", np.uint8),
                       ("pad1", np.uint8, 2), ("component_mask", np.uint64)])


def exit_keys(recs, roots):
    """Canonical per-exit sort key: (root, path_len, path bytes) — unique per exit on a convex crystal."""
    keys = []
    for r, root in zip(recs, roots):
        keys.append((int(root), int(r["path_len"]), bytes(r["path"][: r["path_len"]])))
    return keys


def sort_exits(recs, roots):
    """Canonical order of an exit list: (root, path length, path bytes, weight). Vectorised (np.lexsort): the
    tile-scale parity tests sort millions of records."""
    roots = np.asarray(roots)
    if len(recs) == 0:
        return recs, roots
    plen = recs["path_len"]
    width = max(1, int(plen.max()))
    path = recs["path"][:, :width].copy()
    path[np.arange(width)[None, :] >= plen[:, None]] = 0      # only the record's own path bytes take part
    keys = [recs["weight"].astype(np.float64)] + [path[:, k] for k in range(width - 1, -1, -1)] + [plen, roots]
    idx = np.lexsort(keys)                                      # last key is the primary one
    return recs[idx], roots[idx]


def adversarial_roots(rng, t, n, n_idx):
    """Entry rays that stress the rare branches of the exit-face search (ties, near-edge double continuation,
    zero numerators, fast-test fallbacks): entry points exactly on polygon vertices / edges or 1e-6.5 .. 1e-3.5
    inside an edge, and directions whose REFRACTED ray is aimed (to float rounding) at a vertex, an edge point
    or the interior of another face (inverse Snell; directions that cannot be reached from outside are
    injected as they are). Returns crystal-local (d, p, w, face) like roots_on_crystal."""
    nf, nt = int(t.face_cnt), int(t.subtri_cnt)
    planes = np.ctypeslib.as_array(t.plane)[:nf].astype(np.float64)
    tri_v = np.ctypeslib.as_array(t.tri_v)[:nt].astype(np.float64).reshape(nt, 3, 3)
    tri_face = np.ctypeslib.as_array(t.tri_face)[:nt]
    polys = []
    for k in range(nf):
        tris = tri_v[tri_face == k]
        polys.append(np.vstack([tris[0, 0][None], tris[:, 1], tris[-1, 2][None]]))

    def point_on_face(k, kind):
        poly = polys[k]
        m = len(poly)
        i = int(rng.integers(m))
        a, b, cen = poly[i], poly[(i + 1) % m], poly.mean(axis=0)
        if kind == 0:
            return a.copy()
        e = a + rng.random() * (b - a)
        if kind == 1:
            return e
        if kind == 2:
            return e + (cen - e) * 10.0 ** rng.uniform(-6.5, -3.5)
        return e + (cen - e) * rng.uniform(0.05, 1.0)

    d = np.zeros((n, 3), np.float32)
    p = np.zeros((n, 3), np.float32)
    f = np.zeros(n, np.uint16)
    k = 0
    while k < n:
        fi = int(rng.integers(nf))
        gi = int(rng.integers(nf))
        if gi == fi:
            continue
        pp = point_on_face(fi, int(rng.integers(4)))
        qq = point_on_face(gi, int(rng.integers(4)))
        tdir = qq - pp
        nrm = np.linalg.norm(tdir)
        if nrm < 1e-3:
            continue
        tdir /= nrm
        nvec = planes[fi, :3]
        c = float(tdir @ nvec)
        if c > -1e-3:
            continue                     # must head into the crystal through face fi
        tang = float(n_idx) * (tdir - c * nvec)
        s2 = float(tang @ tang)
        if s2 < 0.98 and rng.random() < 0.85:
            din = tang - np.sqrt(1.0 - s2) * nvec     # refracts into tdir
        else:
            din = tdir                                  # taken as the incoming direction itself
        d[k] = (din / np.linalg.norm(din)).astype(np.float32)
        p[k] = pp.astype(np.float32)
        f[k] = fi
        k += 1
    return d, p, np.ones(n, np.float32), f
