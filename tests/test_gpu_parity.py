"""GPU-vs-oracle parity through the C ABI (the parity tests proper). Protocol: tests/parity.py."""
import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("pipeline", ["fused", "split"])
@pytest.mark.parametrize("name", sorted(parity.CASES))
def test_case_bit_exact(backend, name, pipeline):
    """Face-number sequences, world exit directions and weights bit-exact against the oracle replay of
    the engine's own roots; image within IMG_RTOL of the oracle accumulation of the same exits. Both kernel
    pipelines: the fused bounce kernels (default) and the split optics + intersect kernels."""
    try:
        res = parity.run_case(parity.CASES[name], n_rays=30000, seed=42, backend=backend,
                              fused_bounce=pipeline == "fused")
    finally:
        backend.SetOption("fused_bounce", 1)
    assert res["exits"] > 0
    assert res["paths_equal"], res
    assert res["dirs_bit_equal"], res
    assert res["weights_bit_equal"], res
    assert res["meta_equal"], res
    assert res["stats_ok"], res
    assert res["image_ok"], res
    assert res["landed_rel_err"] < 1e-5, res
    assert res["masks_equal"], res
    if name in ("filter_out_raypath", "filter_d_symmetry"):
        assert 0 < res["exits"] < 30000 * 5          # the filter is selective, not empty and not everything
    if name == "color_classes":
        assert res["mask_bits_seen"] == 0b1111111, bin(res["mask_bits_seen"])   # every predicate fired somewhere
        assert res["lanes_shape_ok"] and res["lanes_ok"], res
        assert all(x > 0 for x in res["lane_sums"][:4]) and res["lane_sums"][4] == 0.0, res["lane_sums"]


def test_unchecked_division_is_ieee(backend):
    """dvd_nr / sqrt_nr (csrc/hb_device.cuh: the IEEE fast paths without the range check and its branch) equal
    __fdiv_rn / __fsqrt_rn bit for bit on 2^32 random operand pairs per mode: positive divisors with signed-zero
    numerators mixed in, divisors of either sign, square roots."""
    for mode in (0, 1, 2):
        bad, a, b, got = backend.SelftestArith(mode, 1 << 32, seed=20260 + mode)
        assert bad == 0, (mode, bad, hex(a), hex(b), hex(got))


def test_seed_and_tile_invariance(backend):
    """Counter-based RNG: the same seed gives the same exits whatever the tile size; another seed differs."""
    a = parity.run_case(parity.CASES["column_config2"], n_rays=20000, seed=7, backend=backend, tile_rays=4096)
    b = parity.run_case(parity.CASES["column_config2"], n_rays=20000, seed=7, backend=backend, tile_rays=1 << 22)
    c = parity.run_case(parity.CASES["column_config2"], n_rays=20000, seed=8, backend=backend, tile_rays=1 << 22)
    assert a["paths_equal"] and b["paths_equal"] and c["paths_equal"]
    assert a["exits"] == b["exits"]
    assert abs(a["landed_gpu"] - b["landed_gpu"]) <= 1e-5 * abs(a["landed_gpu"])
    assert a["exits"] != c["exits"] or a["landed_gpu"] != c["landed_gpu"]


def test_wavelength_pool_per_ray_index(backend):
    """Illuminant-style session: 16-entry wavelength pool, per-ray index drawn on the device
    (WlPoolSize() > 0 contract, trace_backend.hpp:516-521); n and CMF follow the ray."""
    case = dict(parity.CASES["column_config2"])
    case["wl"] = [400.0 + 25.0 * k for k in range(16)]
    res = parity.run_case(case, n_rays=30000, seed=11, backend=backend)
    assert res["paths_equal"] and res["dirs_bit_equal"] and res["weights_bit_equal"] and res["meta_equal"], res
    assert res["image_ok"], res


def test_generated_roots_match_oracle_generator(backend):
    """Root generation (statistical parity with the CPU sampler, exact parity with our own stream spec):
    the oracle's restatement of the generator reproduces the engine's roots up to libm-vs-libdevice ulps."""
    import ctypes as C
    import harness as H
    from ice_halo_sim_b200 import backend as B
    A = H.A
    # every latitude path of the sampler (full sphere, fixed, legacy Gauss, inverse-CDF LUT of each distribution
    # type) and the azimuth / roll distribution types, beyond what the named cases use
    extra = {
        "lat_none_roll_gauss": dict(scene=lambda: parity.scene([(0.0, [parity.prism_pop(
            0.4, zenith=("none", 35.0, 0.0), azimuth=("gauss", 40.0, 10.0), roll=("gauss", 10.0, 5.0))])], 5), wl=[550.0]),
        "lat_gauss_legacy": dict(scene=lambda: parity.scene([(0.0, [parity.prism_pop(
            1.5, zenith=("gauss_legacy", 80.0, 6.0), roll=("none", 30.0, 0.0))])], 5), wl=[550.0]),
        "lat_laplacian": dict(scene=lambda: parity.scene([(0.0, [parity.prism_pop(
            0.3, zenith=("laplacian", 10.0, 3.0), azimuth=("laplacian", 90.0, 20.0), roll=("zigzag", 0.0, 15.0))])], 5),
            wl=[550.0]),
        "lat_zigzag": dict(scene=lambda: parity.scene([(0.0, [parity.prism_pop(
            1.0, zenith=("zigzag", 90.0, 25.0), roll=("uniform", 0.0, 60.0))])], 5), wl=[550.0]),
        "lat_uniform_band": dict(scene=lambda: parity.scene([(0.0, [parity.prism_pop(
            1.0, zenith=("uniform", 60.0, 40.0), azimuth=("zigzag", 0.0, 30.0))])], 5), wl=[550.0]),
    }
    for name in ("column_config2", "stoch_config5", "pyramid") + tuple(sorted(extra)):
        case = parity.CASES[name] if name in parity.CASES else extra[name]
        desc = case["scene"]()
        tables = B.SceneTables(desc, 7)
        wl = [B.make_wl_entry(x, 1.0) for x in case["wl"]]
        wl_arr = (A.HbWlEntry * len(wl))(*[A.HbWlEntry(*e) for e in wl])
        backend.SetScene(tables)
        backend.SetOption("stream_base", 0)
        n = 50000
        backend.BeginSession(B.SessionSpec(seed=1234, wl=wl, record_exits=True, accumulate=False, ray_base=5_000_000_000))
        backend.TraceLayer(B.RootRaySource.FromHost(n))
        backend.DrainExits()
        r = backend.ExportRoots()
        backend.EndSession()
        d = np.zeros((n, 3), np.float32); p = np.zeros((n, 3), np.float32); w = np.zeros(n, np.float32)
        f = np.zeros(n, np.uint16); rot = np.zeros((n, 9), np.float32)
        si = np.zeros(n, np.uint32); wi = np.zeros(n, np.uint32)
        H.oracle().orc_gen_roots(tables.scene_ptr, 0, 0, 0, C.addressof(wl_arr), len(wl), 1234, 5_000_000_000, n,
                                 H.ptr(d), H.ptr(p), H.ptr(w), H.ptr(f), None, H.ptr(rot), H.ptr(si), H.ptr(wi))
        assert np.array_equal(r["shape"], si) and np.array_equal(r["wl"], wi), name
        assert np.allclose(r["rot"], rot, atol=2e-6), name
        assert np.allclose(r["d"], d, atol=5e-6), name
        same_face = r["face"] == f
        assert same_face.mean() > 0.999, (name, same_face.mean())      # categorical boundary flips only
        assert np.allclose(r["p"][same_face], p[same_face], atol=2e-5), name
        assert np.array_equal(r["w"], w)


def test_full_size_invariants(backend):
    """BASELINE-size properties that need no oracle: (1) total Y of the image == cmf_y x landed weight
    (no reduction is lost), (2) splitting a session into two index ranges gives the same image
    (counter-based RNG + additive accumulator), (3) the drain zeroes the accumulator."""
    from ice_halo_sim_b200 import backend as B
    case = parity.CASES["column_config2"]
    tables = B.SceneTables(case["scene"](), 7)
    backend.SetScene(tables)
    backend.SetRender(case["render"]())
    wl = [B.make_wl_entry(550.0, 1.0)]
    backend.ReadbackXyzAccum()
    n = 1 << 24

    def run(splits):
        base = 0
        for cnt in splits:
            backend.BeginSession(B.SessionSpec(seed=99, wl=wl, ray_num=cnt, ray_base=base))
            backend.TraceLayer(B.RootRaySource.FromHost(cnt), want_stats=False)
            backend.EndSession()
            base += cnt
        return backend.ReadbackXyzAccum()

    img_a, landed_a = run([n])
    img_b, landed_b = run([n // 2, n // 2 - 12345, 12345])
    assert landed_a > 0.5 * n * 0.5          # most of the energy lands in this view
    y = img_a[..., 1].astype(np.float64).sum()
    assert abs(y - wl[0][3] * landed_a) <= 2e-4 * y
    assert abs(landed_a - landed_b) <= 1e-5 * landed_a
    denom = np.maximum(np.abs(img_a), 1e-3)
    assert np.max(np.abs(img_a - img_b) / denom) < 5e-3      # same rays, different summation order
    assert abs(img_a.astype(np.float64).sum() - img_b.astype(np.float64).sum()) <= 1e-5 * img_a.sum()
    img_c, landed_c = backend.ReadbackXyzAccum()
    assert landed_c == 0.0 and not img_c.any()


def test_state_machine_errors(backend):
    """Calls outside the BeginSession/EndSession bracket fail loudly (trace_backend.hpp:91-116)."""
    from ice_halo_sim_b200 import backend as B
    from ice_halo_sim_b200.lib import HaloTraceError
    case = parity.CASES["column_config2"]
    backend.SetScene(B.SceneTables(case["scene"](), 7))
    wl = [B.make_wl_entry(550.0, 1.0)]
    with pytest.raises(HaloTraceError):
        backend.TraceLayer(B.RootRaySource.FromHost(10))
    backend.BeginSession(B.SessionSpec(seed=1, wl=wl))
    with pytest.raises(HaloTraceError):
        backend.BeginSession(B.SessionSpec(seed=1, wl=wl))
    h = backend.TraceLayer(B.RootRaySource.FromHost(0))
    assert h.continuation_count == 0
    with pytest.raises(HaloTraceError):
        backend.TraceLayer(B.RootRaySource.FromHost(10))        # needs Recombine first
    backend.Recombine(h)
    with pytest.raises(HaloTraceError):
        backend.TraceLayer(B.RootRaySource.FromDevice(0))       # beyond the configured layers
    backend.EndSession()
    with pytest.raises(HaloTraceError):
        backend.EndSession()


def test_reference_driver_through_adapter():
    """Drop-in: the reference's host code drives B200TraceBackend (adapter/) through the seam; the image
    passes the reference's cross-backend battery against its own CpuTraceBackend
    (Pearson >= 0.95 on 4x4 block means, total Y within 5 %)."""
    import json
    import os
    import subprocess
    import harness as H
    exe = os.path.join(H.ROOT, "oracle", "_ref", "adapter_demo")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/adapter_demo not built (needs /root/reference at build time)")
    for scene, rays in ((0, 1500000), (1, 600000), (2, 1500000), (3, 1500000), (4, 1000000), (5, 1000000)):
        out = subprocess.run([exe, str(scene), str(rays)], capture_output=True, text=True, timeout=600)
        lines = [json.loads(x) for x in out.stdout.strip().splitlines()]
        res = lines[-1]
        assert out.returncode == 0 and res["pass"], lines
        assert res["pearson_4x4"] >= 0.95 and abs(res["total_y_ratio"] - 1) <= 0.05
        if scene == 4:   # exit-seam egress vs device-fused consumer of the same backend: the same rays
            assert res["pearson_4x4"] >= 0.9999 and abs(res["total_y_ratio"] - 1) <= 2e-4
        if scene == 2:   # raypath colour: per-class Y lanes vs lanes built from the CPU backend's component masks
            lanes = [x for x in lines if "class" in x]
            assert len(lanes) == 3 and all(abs(x["ratio"] - 1) <= 0.05 and x["pearson_8x8"] >= 0.95 for x in lanes), lanes


def _fused_run(backend, case, renders, n, seed, wl=None):
    """One fused (no exit materialisation) single-layer session over `renders`; leaves the images on the device."""
    from ice_halo_sim_b200 import backend as B
    tables = B.SceneTables(case["scene"](), 7)
    backend.SetScene(tables)
    backend.SetOption("stream_base", 0)
    backend.SetRenders(renders)
    for r in range(len(renders)):
        backend.ReadbackXyzAccum(render=r)
    wl = wl or [B.make_wl_entry(x, 1.0) for x in case["wl"]]
    backend.BeginSession(B.SessionSpec(seed=seed, wl=wl, ray_num=n))
    backend.TraceLayer(B.RootRaySource.FromHost(n), want_stats=False)
    backend.EndSession()


def test_multi_render_equals_separate_single_render_traces(backend):
    """N projections of one trace (fused kernels, no exit records): each render's image equals what a
    single-render trace of the same seed accumulates; only the order of float additions differs."""
    case = parity.CASES["multi_render"]
    renders = case["render"]()
    n = 400000
    # small tiles folded into the fp64 master after every tile: the fp32 working image stays small, so the
    # comparison is not blurred by fp32 absorption in the ~1e5-term sun-disk pixels
    backend.SetOption("tile_rays", 1 << 15)
    backend.SetOption("fold_rays", 1 << 15)
    _fused_run(backend, case, renders, n, seed=5)
    multi = [backend.ReadbackXyzAccum(render=r) for r in range(len(renders))]
    for r, rd in enumerate(renders):
        _fused_run(backend, case, [rd], n, seed=5)
        img, landed = backend.ReadbackXyzAccum()
        m_img, m_landed = multi[r]
        assert img.shape == m_img.shape and img.sum() > 0
        # same rays, same pixels; only the summation order differs (atomics, per-CTA pixel cache)
        scale = float(np.abs(img).max())
        assert np.allclose(img, m_img, rtol=5e-5, atol=2e-6 * scale), r
        assert abs(float(img.astype(np.float64).sum()) / float(m_img.astype(np.float64).sum()) - 1.0) < 3e-6, r
        assert abs(landed - m_landed) <= 1e-5 * landed, r
    backend.SetOption("tile_rays", 1 << 24)
    backend.SetOption("fold_rays", 1 << 21)
    with pytest.raises(Exception):
        backend.ReadbackXyzAccum(render=3)      # single render set now


def test_device_snapshot_matches_post_snapshot(backend):
    """Display sink on the device (hb_snapshot) against the oracle's PostSnapshot restatement applied to
    the snapshot's own XYZ: 8-bit sRGB equal except where powf (libdevice) and pow (glibc) round a value
    across an integer boundary: |difference| <= 1 level on < 0.5 % of the channels. The snapshot does not
    disturb the accumulator."""
    import harness as H
    from ice_halo_sim_b200 import backend as B
    case = parity.CASES["column_config2"]
    rd = case["render"]()
    _fused_run(backend, case, [rd], 2_000_000, seed=3, wl=B.make_wl_pool("D65", 64))
    variants = [(1.0, (-1.0, -1.0, -1.0), (0.0, 0.0, 0.0)), (8.0, (-1.0, -1.0, -1.0), (0.01, 0.02, 0.04)),
                (3.0, (1.0, 0.7, 0.4), (0.0, 0.0, 0.05))]
    xyz0 = None
    for fac, rc, bg in variants:
        rgb, xyz, inten = backend.Snapshot(0, fac, rc, bg, want_xyz=True)
        assert inten > 0 and xyz.shape == (rd.img_h, rd.img_w, 3) and rgb.shape == xyz.shape
        if xyz0 is not None:
            assert np.array_equal(xyz, xyz0)          # non-destructive
        xyz0 = xyz
        want = np.zeros_like(rgb)
        rc_a, bg_a = np.array(rc, np.float32), np.array(bg, np.float32)
        H.oracle().orc_post_snapshot(H.ptr(xyz), rd.img_w, rd.img_h, inten, fac, H.ptr(rc_a), H.ptr(bg_a), H.ptr(want))
        diff = np.abs(rgb.astype(np.int16) - want.astype(np.int16))
        assert diff.max() <= 1, (fac, int(diff.max()))
        assert (diff != 0).mean() < 5e-3, (fac, float((diff != 0).mean()))
        assert rgb.max() > 200 and len(np.unique(rgb)) > 50
    img, landed = backend.ReadbackXyzAccum()
    assert np.array_equal(img, xyz0) and abs(landed - inten) <= 1e-6 * inten
    rgb, _, inten = backend.Snapshot()                # drained accumulator: black frame, zero intensity
    assert inten == 0.0 and not rgb.any()


def test_driver_renders_whole_config_and_shards_by_rank(backend):
    """Host driver (ice_halo_sim_b200.driver): a two-layer, filtered, two-renderer, 9-wavelength Lumice config
    end to end; tracing it as two ranks' shards (summed) lands the same rays as one rank for the first
    layer, so the frames agree statistically (layer-2 gate/transit streams differ per session)."""
    import json
    from test_host_logic import EXAMPLE
    from ice_halo_sim_b200 import load_config, render_config
    ex = json.loads(json.dumps(EXAMPLE))
    ex["scene"]["ray_num"] = 1_800_000
    cfg = load_config(ex)
    one = render_config(cfg, backend, seed=9, session_rays=1 << 17, srgb=True)
    assert sorted(one) == [1, 4]
    for rid, fr in one.items():
        assert fr.xyz.shape == (1080, 1920, 3) and fr.rgb.shape == fr.xyz.shape and fr.rgb.dtype == np.uint8
        assert fr.landed_weight > 0 and fr.xyz.sum() > 0 and fr.rgb.max() > 100
    halves = [render_config(cfg, backend, seed=9, session_rays=1 << 17, rank=r, world=2, allreduce=False)
              for r in range(2)]
    for rid in one:
        both = halves[0][rid].xyz.astype(np.float64) + halves[1][rid].xyz
        tot1, tot2 = one[rid].xyz.astype(np.float64).sum(), both.sum()
        assert abs(tot1 / tot2 - 1.0) < 0.02, (rid, tot1, tot2)
        landed = halves[0][rid].landed_weight + halves[1][rid].landed_weight
        assert abs(landed / one[rid].landed_weight - 1.0) < 0.02
        # block-mean Pearson, the reference battery's statistic (test_cuda_backend_parity.cpp:217-239)
        a = one[rid].xyz[..., 1].reshape(27, 40, 48, 40).mean(axis=(1, 3)).ravel()
        b = both[..., 1].reshape(27, 40, 48, 40).mean(axis=(1, 3)).ravel()
        assert np.corrcoef(a, b)[0, 1] > 0.95, rid
    # illuminant spectrum: one pool session, per-ray wavelength index
    ex["scene"]["light_source"]["spectrum"] = "D65"
    ex["scene"]["ray_num"] = 400_000
    fr = render_config(load_config(ex), backend, seed=9)[4]
    assert fr.landed_weight > 0 and (fr.xyz[..., 0].sum() > 0 and fr.xyz[..., 2].sum() > 0)


def test_device_geometry_pool_tables_equal_host_builder(backend):
    """hb_resample_shapes: the shapes drawn and built on the device are byte-identical to what the
    host builders (hb_make_prism / hb_make_pyramid, themselves pinned to the reference's MakeCrystal tables in
    test_host_tables.py) make of the same scalars; the scalars follow the configured distributions; a second
    draw range gives different shapes, the same range the same ones."""
    import ctypes as C
    import harness as H
    from ice_halo_sim_b200 import backend as B
    A = H.A
    lib = B.load()
    prism = parity.prism_pop(parity.dist("uniform", 1.2, 0.6), zenith=("uniform", 90, 360),
                             face_dist=[parity.dist("gauss", 1.0, 0.15)] * 6)
    prism.crystal.sync_group[4] = prism.crystal.sync_group[6] = 1   # faces 0 and 2 share one draw (SyncGroupSampler)
    pyr = parity.pyramid_pop()
    for k in range(3):
        pyr.crystal.height[k] = parity.dist("uniform", pyr.crystal.height[k].center, 0.2)
    for k in range(6):
        pyr.crystal.face_dist[k] = parity.dist("gauss", 1.0, 0.1)
    for pop, n_pool in ((prism, 512), (pyr, 128)):
        tables = B.SceneTables(parity.scene([(0.0, [pop])], 6, pool=n_pool), 3)
        backend.SetScene(tables)
        assert backend.ResampleShapes(0, 0, pop.crystal, 77, 0) == 0
        tb, sc = backend.ExportShapes(0, 0)
        assert len(tb) == n_pool and (sc[:, 9] == 0).all()
        want = A.HbCrystalTables()
        dist6 = (C.c_float * 6)()
        for k in range(n_pool):
            for i in range(6):
                dist6[i] = float(sc[k, 3 + i])
            if pop.crystal.kind == 0:
                assert lib.hb_make_prism(C.c_float(float(sc[k, 0])), dist6, C.byref(want)) == 0
            else:
                assert lib.hb_make_pyramid(C.c_float(pop.crystal.wedge_upper_deg), C.c_float(pop.crystal.wedge_lower_deg),
                                           C.c_float(float(sc[k, 0])), C.c_float(float(sc[k, 1])), C.c_float(float(sc[k, 2])),
                                           dist6, C.byref(want)) == 0
            assert bytes(tb[k]) == bytes(want), (pop.crystal.kind, k)
            assert tb[k].face_cnt >= 4
        if pop.crystal.kind == 0:   # sync group: members identical, others independent
            assert np.array_equal(sc[:, 3], sc[:, 5]) and not np.array_equal(sc[:, 3], sc[:, 4])
        d = sc[:, 3:9].astype(np.float64)
        spread = pop.crystal.face_dist[0].spread
        assert abs(d.mean() - 1.0) < 4.5 * spread / np.sqrt(d.size)
        assert abs(d.std() - spread) < 0.1 * spread
        backend.ResampleShapes(0, 0, pop.crystal, 77, 0)
        _, sc_same = backend.ExportShapes(0, 0)
        backend.ResampleShapes(0, 0, pop.crystal, 77, n_pool)
        _, sc_next = backend.ExportShapes(0, 0)
        assert np.array_equal(sc, sc_same) and not np.array_equal(sc, sc_next)


@pytest.mark.parametrize("name", ["stoch_config5"])
def test_device_geometry_pool_traces_bit_exact(backend, name):
    """Parity protocol on a pool that was drawn and built on the device (tables, axis tables, entry face groups
    all derived by the device kernels): same bit-exact bar as the host-built pools."""
    res = parity.run_case(parity.CASES[name], n_rays=20000, seed=5, backend=backend, device_pool_seed=1234)
    assert res["paths_equal"] and res["dirs_bit_equal"] and res["weights_bit_equal"] and res["meta_equal"], res
    assert res["image_ok"], res


def test_injected_adversarial_rays_bit_exact(backend):
    """Rays that enter exactly on vertices / edges and are refracted onto vertices / edges of other faces
    (harness.adversarial_roots): ties between planes, t ~ 0 candidates, near-edge double continuation (fork rays),
    the fall-backs of the quick far-child / near-child tests and the reference-order tie scan, in the unrolled
    prism kernels (P4) and the generic axis-loop kernels alike. Same bit-exact bar as the random rays; the oracle
    is pinned to the reference on the same kind of rays in test_oracle_vs_reference.py."""
    import harness as H
    from ice_halo_sim_b200 import backend as B
    from test_oracle_golden import oracle_trace_single
    rng = np.random.default_rng(4242)
    pops = [parity.prism_pop(1.3), parity.prism_pop(0.3),
            parity.prism_pop(0.8, face_dist=[1.0, 0.7, 1.2, 0.9, 1.1, 0.8]),
            parity.prism_pop(1.0, face_dist=[1.0, 2.4, 1.0, 1.0, 1.0, 1.0]),      # one prism face vanishes: generic kernels
            parity.pyramid_pop(), parity.pyramid_pop(h=(0.6, 0.0, 0.6))]
    wl = [B.make_wl_entry(550.0, 1.0)]
    total = 0
    for use_p4 in (1, 0):
        backend.SetOption("prism_fast_path", use_p4)
        for pop in pops[: (len(pops) if use_p4 else 2)]:
            mh = 8
            tables = B.SceneTables(parity.scene([(0.0, [pop])], mh), 1)
            t = tables.scene().layers[0].populations[0].shapes[0]
            assert t.face_cnt >= 4
            n = 4000
            d, p, w, f = H.adversarial_roots(rng, t, n, wl[0][0])
            backend.SetScene(tables)
            backend.SetOption("stream_base", 0)
            backend.BeginSession(B.SessionSpec(seed=3, wl=wl, ray_num=n, record_exits=True, accumulate=False))
            backend.TraceLayer(B.RootRaySource.FromHost(n, d, p, w, f))
            g_ex, g_roots = backend.DrainExits(with_roots=True)
            backend.EndSession()
            o_ex, o_roots = oracle_trace_single(t, np.float32(wl[0][0]), mh, d, p, w, f)
            r = parity.compare_exit_lists(g_ex, g_roots, o_ex, o_roots)
            assert r["paths_equal"] and r["dirs_bit_equal"] and r["weights_bit_equal"], (use_p4, t.face_cnt, r)
            total += len(g_ex)
    backend.SetOption("prism_fast_path", 1)
    assert total > 100000


def test_geometry_clock_prefetch(backend):
    """hb_auto_resample: every BeginSession swaps in a fresh pool that was drawn one session ahead. Session k's pool
    equals an explicit hb_resample_shapes of the stream range [k n, (k + 1) n); traces on the swapped-in pools
    pass the bit-exact protocol; switching the clock off freezes the live pool."""
    from ice_halo_sim_b200 import backend as B
    case = parity.CASES["stoch_config5"]
    desc = case["scene"]()
    n_pool = int(desc.geom_pool_size)
    crystal = desc.layers[0].populations[0].crystal
    tables = B.SceneTables(desc, 7)
    wl = [B.make_wl_entry(550.0, 1.0)]
    backend.SetScene(tables)
    backend.SetRender(case["render"]())
    backend.AutoResample(0, 0, crystal, 31, 0)
    seen = []
    for k in range(3):
        backend.BeginSession(B.SessionSpec(seed=1, wl=wl, ray_num=4096))
        backend.TraceLayer(B.RootRaySource.FromHost(4096), want_stats=False)
        backend.EndSession()
        seen.append(backend.ExportShapes(0, 0)[1].copy())
    backend.AutoResample(0, 0, None)
    backend.BeginSession(B.SessionSpec(seed=1, wl=wl, ray_num=16))
    backend.TraceLayer(B.RootRaySource.FromHost(16), want_stats=False)
    backend.EndSession()
    assert np.array_equal(backend.ExportShapes(0, 0)[1], seen[2])          # clock off: pool frozen
    backend.ReadbackXyzAccum()
    for k in range(3):
        backend.ResampleShapes(0, 0, crystal, 31, k * n_pool)
        assert np.array_equal(backend.ExportShapes(0, 0)[1], seen[k]), k
    assert not np.array_equal(seen[0], seen[1])
    for seed in (5, 6):   # two consecutive sessions on the clock, each against the oracle on its own pool
        res = parity.run_case(case, n_rays=15000, seed=seed, backend=backend, geometry_clock_seed=99)
        assert res["paths_equal"] and res["dirs_bit_equal"] and res["weights_bit_equal"] and res["meta_equal"], res
        assert res["image_ok"], res


def test_driver_stochastic_config_device_geometry(backend):
    """driver.render_config on a bench_config_stoch.json-shaped config (prism, face distances ~ N(1, 0.15), full-sphere
    axis, rectangular full-sky lens): the engine-run geometry clock (fresh device-built pool per session) and the
    host-drawn pool give statistically the same frame."""
    from ice_halo_sim_b200 import load_config, render_config
    g = {"type": "gauss", "mean": 1.0, "std": 0.15}
    u360 = {"type": "uniform", "mean": 0.0, "std": 360.0}
    cfg_json = {
        "crystal": [{"id": 1, "type": "prism", "shape": {"height": 1.0, "face_distance": [g] * 6},
                     "axis": {"zenith": u360, "azimuth": u360, "roll": u360}}],
        "filter": [],
        "render": [{"id": 1, "lens": {"type": "rectangular", "fov": 180.0}, "resolution": [1024, 512],
                    "view": {"azimuth": 0.0, "elevation": 0.0, "roll": 0.0}, "visible": "full"}],
        "scene": {"id": 1, "light_source": {"type": "sun", "altitude": 20.0, "azimuth": 0.0, "diameter": 0.5,
                                            "spectrum": [{"wavelength": 550.0, "weight": 1.0}]},
                  "ray_num": 3_000_000, "max_hits": 8,
                  "scattering": [{"prob": 0.0, "entries": [{"crystal": 1, "proportion": 1.0}]}]},
    }
    cfg = load_config(cfg_json, geom_pool_size=128)
    dev = render_config(cfg, backend, seed=3, session_rays=1 << 18, device_geometry=True)[1]
    host = render_config(cfg, backend, seed=3, session_rays=1 << 18, device_geometry=False)[1]
    assert dev.landed_weight > 0 and abs(dev.landed_weight / host.landed_weight - 1.0) < 0.02
    a = dev.xyz[..., 1].reshape(32, 16, 64, 16).mean(axis=(1, 3)).ravel()
    b = host.xyz[..., 1].reshape(32, 16, 64, 16).mean(axis=(1, 3)).ravel()
    assert np.corrcoef(a, b)[0, 1] > 0.95


def test_two_layer_full_size_invariants(backend):
    """BASELINE config 4 shape (two layers, prob 1.0) at 4 Mi roots, properties that need no oracle: every exit of
    layer 0 continues (LayerStats: continuations == exits, nothing lands from layer 0), layer 1 traces exactly those
    continuations, energy only decreases, and total Y == cmf_y x landed weight after both layers."""
    from ice_halo_sim_b200 import backend as B
    case = parity.CASES["two_layer_config4"]
    backend.SetScene(B.SceneTables(case["scene"](), 7))
    backend.SetRender(case["render"]())
    backend.ReadbackXyzAccum()
    wl = [B.make_wl_entry(550.0, 1.0)]
    n = 1 << 22
    backend.BeginSession(B.SessionSpec(seed=21, wl=wl, ray_num=n, accumulate=True))
    h0 = backend.TraceLayer(B.RootRaySource.FromHost(n), want_stats=True)
    assert h0.root_count == n
    assert h0.continuation_count == h0.exit_count and 4.0 * n < h0.exit_count < 8.0 * n
    assert h0.exit_w_sum <= n * (1 + 1e-6)                      # Fresnel splitting never creates energy
    _, landed_mid = backend.ReadbackXyzAccum()
    assert landed_mid == 0.0                                     # prob 1.0: nothing of layer 0 reaches the image
    roots = backend.Recombine(h0, shuffle=True)
    assert roots.is_device and roots.count == h0.continuation_count
    h1 = backend.TraceLayer(roots, want_stats=True)
    backend.EndSession()
    assert h1.root_count == h0.continuation_count and h1.continuation_count == 0
    assert h1.exit_w_sum <= h0.exit_w_sum * (1 + 1e-6)
    img, landed = backend.ReadbackXyzAccum()
    assert 0 < landed <= h1.exit_w_sum * (1 + 1e-5)              # "upper" view: only part of the sky lands
    y = img[..., 1].astype(np.float64).sum()
    assert abs(y - wl[0][3] * landed) <= 2e-4 * y


ENGINE_DEFAULTS = {"fused_bounce": 1, "fused_gen": 0, "filter_hit_bound": 1}


def _fused_image(backend, case, n, seed, want_stats, opts):
    """Single-layer accumulate-only session under engine options `opts`; returns (image, landed)."""
    from ice_halo_sim_b200 import backend as B
    for k, v in opts.items():
        backend.SetOption(k, v)
    try:
        tables = B.SceneTables(case["scene"](), 7)
        backend.SetScene(tables)
        backend.SetOption("stream_base", 0)
        backend.SetRender(case["render"]())
        backend.ReadbackXyzAccum()
        wl = [B.make_wl_entry(x, 1.0) for x in case["wl"]]
        backend.BeginSession(B.SessionSpec(seed=seed, wl=wl, ray_num=n))
        backend.TraceLayer(B.RootRaySource.FromHost(n), want_stats=want_stats)
        backend.EndSession()
        return backend.ReadbackXyzAccum()
    finally:
        for k in opts:
            backend.SetOption(k, ENGINE_DEFAULTS[k])


@pytest.mark.parametrize("name", ["column_config2", "stoch_config5", "pyramid", "two_populations"])
def test_production_kernels_equal_parity_kernels(backend, name):
    """The kernels a production session runs (no exit records: fast non-general instantiations, root generation
    fused with the entry interaction) trace the same rays as the general kernels the bit-exact protocol runs on:
    same seed => same exits => the same image up to the order of the float additions. Covers every pipeline
    variant: gen fused / separate, bounce fused / split optics + intersect."""
    case = parity.CASES[name]
    n = 300000
    backend.SetOption("tile_rays", 1 << 15)   # small tiles folded into the fp64 master: fp32 absorption cannot blur the comparison
    backend.SetOption("fold_rays", 1 << 15)
    try:
        ref_img, ref_landed = _fused_image(backend, case, n, 17, True, {})   # general kernels (LayerStats)
        assert ref_landed > 0
        scale = float(np.abs(ref_img).max())
        for opts in ({}, {"fused_gen": 1}, {"fused_bounce": 0}):
            img, landed = _fused_image(backend, case, n, 17, False, opts)
            assert abs(landed - ref_landed) <= 1e-5 * ref_landed, (name, opts)
            assert np.allclose(img, ref_img, rtol=5e-5, atol=2e-6 * scale), (name, opts)
            assert abs(float(img.astype(np.float64).sum()) / float(ref_img.astype(np.float64).sum()) - 1.0) < 3e-6
    finally:
        backend.SetOption("tile_rays", 1 << 24)
        backend.SetOption("fold_rays", 1 << 21)


def test_device_overflow_is_reported_not_dropped(backend):
    """Fork-slot and continuation-pool overflow (forced through the `fork_cap` / `cont_cap` options) in the
    production call sequence -- no LayerStats, no hb_synchronize -- surface as HB_ERR_CAPACITY at the next
    synchronising call (ReadbackXyzAccum at the latest) and are reported once; the session after is clean."""
    from ice_halo_sim_b200 import backend as B
    from ice_halo_sim_b200.lib import HaloTraceError
    wl = [B.make_wl_entry(550.0, 1.0)]
    # (a) continuation pool: two layers, prob 1.0, pool capped at 1000 records
    case = parity.CASES["two_layer_config4"]
    backend.SetScene(B.SceneTables(case["scene"](), 7))
    backend.SetRender(case["render"]())
    backend.ReadbackXyzAccum()
    backend.SetOption("cont_cap", 1000)
    try:
        backend.BeginSession(B.SessionSpec(seed=5, wl=wl, ray_num=100000))
        with pytest.raises(HaloTraceError) as ei:
            backend.TraceLayer(B.RootRaySource.FromHost(100000), want_stats=False)   # the gate layer reads its count: sync
        assert ei.value.status == -5 and "continuation" in str(ei.value)
        backend.EndSession()
    finally:
        backend.SetOption("cont_cap", 0)
    backend.ReadbackXyzAccum()       # reported once: the accumulator is readable again
    # (b) fork slots: single layer, one fork slot; ~1e-6 of the ray-bounces fork
    case = parity.CASES["column_config2"]
    backend.SetScene(B.SceneTables(case["scene"](), 7))
    backend.SetRender(case["render"]())
    backend.SetOption("fork_cap", 1)
    try:
        n = 1 << 23
        backend.BeginSession(B.SessionSpec(seed=5, wl=wl, ray_num=n))
        backend.TraceLayer(B.RootRaySource.FromHost(n), want_stats=False)             # no sync, no error yet
        try:
            backend.EndSession()                                                     # may already see the mirror
            with pytest.raises(HaloTraceError) as ei:
                backend.ReadbackXyzAccum()
        except HaloTraceError as e:
            ei = type("E", (), {"value": e})
        assert ei.value.status == -5 and "fork" in str(ei.value)
    finally:
        backend.SetOption("fork_cap", 0)
    backend.ReadbackXyzAccum()
    n = 1 << 20
    backend.BeginSession(B.SessionSpec(seed=5, wl=wl, ray_num=n))
    backend.TraceLayer(B.RootRaySource.FromHost(n), want_stats=False)
    backend.EndSession()
    img, landed = backend.ReadbackXyzAccum()                                          # clean again
    assert landed > 0


def test_sharded_sessions_draw_disjoint_layer_streams(backend):
    """Multi-GPU sharding: sessions that start at different global ray indices (two ranks' shards)
    must draw different transit / gate streams in layers >= 1 too, not only different roots; the same index range
    replays exactly."""
    from ice_halo_sim_b200 import backend as B
    case = parity.CASES["partial_prob"]
    tables = B.SceneTables(case["scene"](), 7)
    wl = [B.make_wl_entry(530.0, 1.0)]
    backend.SetScene(tables)
    n = 4000

    def layer1_roots(ray_base):
        backend.BeginSession(B.SessionSpec(seed=3, wl=wl, ray_num=n, record_exits=True, accumulate=False, ray_base=ray_base))
        h = backend.TraceLayer(B.RootRaySource.FromHost(n))
        backend.DrainExits()
        backend.ExportRoots()
        src = backend.Recombine(h, shuffle=True)
        backend.TraceLayer(src)
        backend.DrainExits()
        r = backend.ExportRoots()
        backend.EndSession()
        return h.continuation_count, r

    c0, r0 = layer1_roots(0)
    c1, r1 = layer1_roots(n)          # the neighbouring rank's shard
    c0b, r0b = layer1_roots(0)
    assert c0 > 100 and c1 > 100
    m = min(c0, c1)
    assert not np.array_equal(r0["rot"][:m], r1["rot"][:m])            # different orientation draws
    assert (np.abs(r0["rot"][:m] - r1["rot"][:m]).max(axis=1) > 1e-3).mean() > 0.99
    # the same index range replays the same draws (the continuation pool is filled by atomics, so WHICH continuation
    # meets draw k may differ between runs: orientations replay exactly, entry points follow the direction they meet)
    assert c0 == c0b and np.array_equal(r0["rot"], r0b["rot"])


@pytest.mark.parametrize("name", ["plate_filter_config3", "complex_filter", "filter_d_symmetry", "filter_out_raypath"])
def test_filter_hit_bound_changes_nothing(backend, name):
    """A layer whose populations all carry a length-bounded filter_in filter ends its hit loop at that length
    (filter_max_len): interactions beyond it can only produce exits the filter rejects. Same image, same landed
    weight and the same LayerStats as tracing all max_hits interactions."""
    from ice_halo_sim_b200 import backend as B
    case = parity.CASES[name]
    n = 400000
    backend.SetOption("tile_rays", 1 << 15)
    backend.SetOption("fold_rays", 1 << 15)
    try:
        full_img, full_landed = _fused_image(backend, case, n, 23, False, {"filter_hit_bound": 0})
        c0 = backend.Counters()
        img, landed = _fused_image(backend, case, n, 23, False, {})
        c1 = backend.Counters()
        assert full_landed > 0 and abs(landed - full_landed) <= 1e-5 * full_landed
        scale = float(np.abs(full_img).max())
        assert np.allclose(img, full_img, rtol=5e-5, atol=2e-6 * scale)
        launches = (c1.bounce_launches - c0.bounce_launches)
        tiles = -(-n // (1 << 15))
        max_hits = int(case["scene"]().max_hits)
        if name == "plate_filter_config3":
            assert launches == 2 * tiles          # raypath [3, 5]: two interactions instead of seven
        if name == "filter_out_raypath":
            assert launches == max_hits * tiles   # filter_out bounds nothing
    finally:
        backend.SetOption("tile_rays", 1 << 24)
        backend.SetOption("fold_rays", 1 << 21)


def _gpu_count():
    import torch
    return torch.cuda.device_count()


def test_multi_device_backend_behind_the_seam():
    """One B200TraceBackend instance fanned out over 2 devices (adapter ctor with a device list): the reference's
    host code drives it through the seam exactly as the single-device one; device 0 gathers the peer's accumulator
    (hb_merge_from_peer, P2P loads) at ReadbackXyzAccum. Same cross-backend battery against CpuTraceBackend, one
    layer and two layers. Needs 2 GPUs (gpurun --gpus 2)."""
    import json
    import os
    import subprocess
    import harness as H
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(H.ROOT, "oracle", "_ref", "adapter_demo")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/adapter_demo not built")
    for mode, rays in ((7, 1500000), (6, 600000)):
        out = subprocess.run([exe, str(mode), str(rays), "2"], capture_output=True, text=True, timeout=600)
        lines = [json.loads(x) for x in out.stdout.strip().splitlines()]
        res = lines[-1]
        assert out.returncode == 0 and res["pass"] and res["devices"] == 2, (lines, out.stderr[-500:])
        assert res["pearson_4x4"] >= 0.95 and abs(res["total_y_ratio"] - 1) <= 0.05


def test_two_engines_merge_equals_one_engine(backend):
    """hb_merge_from_peer: two engines (two devices when the box has them, else the same device twice) trace the two
    halves of an index range; after the merge engine 0 holds the image a single engine accumulates for the whole
    range and the peer's accumulator is zero."""
    from ice_halo_sim_b200 import B200TraceBackend
    from ice_halo_sim_b200 import backend as B
    case = parity.CASES["column_config2"]
    tables = B.SceneTables(case["scene"](), 7)
    wl = [B.make_wl_entry(550.0, 1.0)]
    n = 1 << 20
    other = B200TraceBackend(1 if _gpu_count() >= 2 else 0)
    try:
        for be in (backend, other):
            be.SetScene(tables)
            be.SetRender(case["render"]())
            be.ReadbackXyzAccum()
        backend.BeginSession(B.SessionSpec(seed=31, wl=wl, ray_num=n, ray_base=0))
        backend.TraceLayer(B.RootRaySource.FromHost(n), want_stats=False)
        backend.EndSession()
        whole, whole_landed = backend.ReadbackXyzAccum()
        for be, base in ((backend, 0), (other, n // 2)):
            be.BeginSession(B.SessionSpec(seed=31, wl=wl, ray_num=n // 2, ray_base=base))
            be.TraceLayer(B.RootRaySource.FromHost(n // 2), want_stats=False)
            be.EndSession()
        backend.MergeFromPeer(other)
        merged, merged_landed = backend.ReadbackXyzAccum()
        rest, rest_landed = other.ReadbackXyzAccum()
        assert rest_landed == 0.0 and not rest.any()
        assert abs(merged_landed - whole_landed) <= 1e-5 * whole_landed
        scale = float(np.abs(whole).max())
        assert np.allclose(merged, whole, rtol=2e-4, atol=2e-6 * scale)
        assert abs(float(merged.astype(np.float64).sum()) / float(whole.astype(np.float64).sum()) - 1.0) < 1e-5
    finally:
        other.close()


def test_two_rank_reduce_equals_single_rank():
    """Frame end on hardware (needs 2 GPUs): two ranks' shards reduced with hb_reduce_image == one rank's image of the
    same index range; the reduce is idempotent, the all-reduce refuses a second call, and the ranks draw disjoint
    layer-1 streams. See tests/mp_reduce_check.py."""
    import json
    import os
    import subprocess
    import sys
    import harness as H
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(H.ROOT, "tests", "mp_reduce_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [x for x in p.stdout.splitlines() if x.startswith("{")]
    assert p.returncode == 0 and lines, (p.stdout[-1000:], p.stderr[-2000:])
    res = json.loads(lines[-1])
    assert res["reduce_image_ok"] and res["reduce_landed_rel"] < 1e-5 and res["reduce_sum_rel"] < 1e-5, res
    assert res["peer_zero"] and res["allreduce_refuses_second_call"] and res["allreduce_landed_rel"] < 1e-5, res
    assert res["layer1_orientations_differ"], res
    assert res["driver_frames_on_root"] and res["driver_no_frames_on_peer"] and res["driver_landed_rel"] < 1e-5, res


def test_reference_simulator_run_drives_the_engine():
    """The reference's own driver, unmodified -- Simulator::Run -> CreateBackend -> SimulateOneWavelengthWithBackend ->
    third-clock DrainDeviceXyz (simulator.cpp:959-1117,1409-1694) -- on this engine: oracle/_ref/libhalo_refb200.so is
    the reference core compiled with oracle/shim ahead of its include path, so the backend its CreateBackend
    instantiates IS adapter/b200_trace_backend.hpp. The frames it drains pass the reference's cross-backend battery
    (4x4 block-mean Pearson >= 0.95, total Y and landed weight within 5 %) against CpuTraceBackend on the same scene,
    at the reference's CUDA dispatch size (262144 rays per SimBatch) with two wavelengths."""
    import json
    import os
    import subprocess
    import sys
    import harness as H
    so = os.path.join(H.ROOT, "oracle", "_ref", "libhalo_refb200.so")
    if not os.path.exists(so) or not H.have_ref():
        pytest.skip("oracle/_ref/libhalo_refb200.so not built (needs /root/reference at build time)")
    code = f"""
import ctypes as C, json, sys
sys.path.insert(0, {H.ROOT!r}); sys.path.insert(0, {os.path.join(H.ROOT, 'tests')!r})
import numpy as np
import harness as H
from ice_halo_sim_b200 import scenes
case = scenes.CASES["column_config2"]
desc, rd = case["scene"](), case["render"]()
rd.img_w, rd.img_h = 480, 270
wls = [550.0, 610.0]
n = 4 * 262144
lib = C.CDLL({so!r})
vp = C.c_void_p
lib.ref_backend_bench.argtypes = [vp, vp, vp, vp, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, vp, vp, vp, vp, vp]
wl = np.array(wls, np.float32); ww = np.ones(len(wl), np.float32)
img = np.zeros((270, 480, 3), np.float32)
landed, rate, sec, used = C.c_double(), C.c_double(), C.c_double(), C.c_uint32()
lib.ref_backend_bench(C.byref(desc), C.byref(rd), wl.ctypes.data, ww.ctypes.data, len(wl), n, 262144, 42,
                      img.ctypes.data, C.byref(landed), C.byref(rate), C.byref(sec), C.byref(used))
cpu = np.zeros_like(img); cpu_landed = 0.0
for w in wls:   # CpuTraceBackend (mt19937 sampling), same scene, fewer rays
    one = np.zeros_like(img); l = C.c_float(); ec = C.c_uint64(); ws = C.c_double()
    H.ref().ref_cpu_backend_run(C.byref(desc), C.byref(rd), w, 1.0, 7, 300000, 4096, one.ctypes.data, C.byref(l),
                                C.byref(ec), C.byref(ws))
    cpu += one; cpu_landed += l.value
scale = n / 300000.0
a = img[:268, :, 1].reshape(67, 4, 120, 4).mean(axis=(1, 3)).ravel()
b = cpu[:268, :, 1].reshape(67, 4, 120, 4).mean(axis=(1, 3)).ravel()
print(json.dumps(dict(used=int(used.value), pearson=float(np.corrcoef(a, b)[0, 1]),
                      y_ratio=float(img[..., 1].sum() / (cpu[..., 1].sum() * scale)),
                      landed_ratio=float(landed.value / (cpu_landed * scale)), mrays=rate.value / 1e6)))
"""
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    lines = [x for x in p.stdout.splitlines() if x.startswith("{")]
    assert p.returncode == 0 and lines, (p.stdout[-500:], p.stderr[-1500:])
    res = json.loads(lines[-1])
    assert res["used"] == 1, res                      # the TraceBackend route ran, not the legacy CPU fallback
    assert res["pearson"] >= 0.95 and abs(res["y_ratio"] - 1) <= 0.05 and abs(res["landed_ratio"] - 1) <= 0.05, res


@pytest.mark.parametrize("name,chunks", [("column_config2", 8), ("plate_filter_config3", 8), ("stoch_config5", 8),
                                         ("two_layer_config4", 3)])
def test_baseline_configs_bit_exact_at_tile_scale(backend, name, chunks):
    """The bit-exact protocol on the BASELINE scenes at >= 4 Mi traced rays per config (config 2/3/5: 8 x 2^19 roots;
    config 4: 3 x 2^18 roots = 0.75 Mi roots whose 4.7 continuations each make 3.7 Mi layer-1 rays): every exit's
    face-number path, world direction and weight against the oracle replay, image within the per-pixel tolerance.
    Chunked so host memory stays bounded (96-byte records: 2.4 M exits per chunk); each chunk is its own seed."""
    n = 1 << 18 if name == "two_layer_config4" else 1 << 19
    exits = 0
    for c in range(chunks):
        res = parity.run_case(parity.CASES[name], n_rays=n, seed=1000 + c, backend=backend)
        assert res["paths_equal"] and res["dirs_bit_equal"] and res["weights_bit_equal"] and res["meta_equal"], (name, c, res)
        assert res["stats_ok"] and res["image_ok"], (name, c, res)
        exits += res["exits"]
    assert exits > 0
