"""Scene / render descriptors in the library's own description ABI (HbSceneDesc, HbRenderDesc): small builders and
the named scenes of BASELINE.json's configs plus the parity suite's cases. Used by bench.py (workloads), smoke()
and tests/parity.py (cases); everything here is host-side description, no tracing.
"""
from . import _abi as A


def dist(t, c=0.0, s=0.0):
    return A.HbDist(A.DIST[t] if isinstance(t, str) else t, float(c), float(s))


def prism_pop(h=1.0, zenith=("none", 0.0, 0.0), azimuth=("uniform", 0, 360), roll=("uniform", 0, 360),
              face_dist=None, cid=1, proportion=1.0, filt=None):
    p = A.HbPopulationDesc()
    p.proportion = proportion
    c = p.crystal
    c.kind, c.id = 0, cid
    c.height[0] = h if isinstance(h, A.HbDist) else dist("none", h)
    for i in range(6):
        fd = 1.0 if face_dist is None else face_dist[i]
        c.face_dist[i] = fd if isinstance(fd, A.HbDist) else dist("none", fd)
    z = dist(*zenith)
    c.latitude = A.HbDist(z.type, 90.0 - z.center, z.spread)
    c.azimuth, c.roll = dist(*azimuth), dist(*roll)
    p.filter.simple.entry_fn = p.filter.simple.exit_fn = -1
    if filt is not None:
        p.filter = filt
    return p


def pyramid_pop(h=(0.3, 0.5, 0.4), alpha=(28.0, 28.0), zenith=("uniform", 90, 360), azimuth=("uniform", 0, 360),
                roll=("uniform", 0, 360), face_dist=None, cid=2, proportion=1.0):
    p = A.HbPopulationDesc()
    p.proportion = proportion
    c = p.crystal
    c.kind, c.id = 1, cid
    for i in range(3):
        c.height[i] = dist("none", h[i])
    for i in range(6):
        fd = 1.0 if face_dist is None else face_dist[i]
        c.face_dist[i] = fd if isinstance(fd, A.HbDist) else dist("none", fd)
    c.wedge_upper_deg, c.wedge_lower_deg = alpha
    z = dist(*zenith)
    c.latitude = A.HbDist(z.type, 90.0 - z.center, z.spread)
    c.azimuth, c.roll = dist(*azimuth), dist(*roll)
    p.filter.simple.entry_fn = p.filter.simple.exit_fn = -1
    return p


def raypath_filter(path, symmetry="", action=0):
    f = A.HbFilterSpecDesc()
    f.kind, f.action = 1, action
    f.symmetry = sum({"P": 1, "B": 2, "D": 4}[ch] for ch in symmetry)
    f.simple.kind = 1
    f.simple.path_len = len(path)
    for i, x in enumerate(path):
        f.simple.path[i] = x
    f.simple.entry_fn = f.simple.exit_fn = -1
    return f


def simple_spec(s, kind, path=(), entry=-1, exit=-1, min_len=1, max_len=0, lon=0.0, lat=0.0, radii=0.0, crystal_id=0):
    s.kind = kind
    s.path_len = len(path)
    for i, x in enumerate(path):
        s.path[i] = x
    s.entry_fn, s.exit_fn, s.min_len, s.max_len = entry, exit, min_len, max_len
    s.lon_deg, s.lat_deg, s.radii_deg, s.crystal_id = lon, lat, radii, crystal_id
    return s


def complex_filter(terms, symmetry="", action=0):
    """terms: [[dict(kind=..., ...), ...], ...] — OR over terms, AND inside a term."""
    f = A.HbFilterSpecDesc()
    f.kind, f.action = 5, action
    f.symmetry = sum({"P": 1, "B": 2, "D": 4}[ch] for ch in symmetry)
    f.simple.entry_fn = f.simple.exit_fn = -1
    f.term_cnt = len(terms)
    for o, term in enumerate(terms):
        f.term_len[o] = len(term)
        for a, kw in enumerate(term):
            simple_spec(f.terms[o][a], **kw)
    return f


def color_pred(pop, bit, symmetry="", **spec):
    """Append one raypath-colour predicate (HbColorPredDesc) to a population description."""
    cp = pop.color_preds[pop.color_pred_cnt]
    simple_spec(cp.pred, **spec)
    cp.symmetry = sum({"P": 1, "B": 2, "D": 4}[ch] for ch in symmetry)
    cp.bit = bit
    pop.color_pred_cnt += 1
    return pop


def color_classes(desc, classes):
    """classes: [(bits, 'any' | 'all')]"""
    desc.color_classes.class_cnt = len(classes)
    allm = 0
    for c, (bits, combine) in enumerate(classes):
        desc.color_classes.bits[c] = bits
        if combine == "all":
            allm |= 1 << c
    desc.color_classes.combine_all_mask = allm
    return desc


def scene(layers, max_hits=7, sun=(20.0, 0.0, 0.5), pool=1):
    """layers: [(prob, [HbPopulationDesc, ...]), ...]"""
    d = A.HbSceneDesc()
    d.max_hits = max_hits
    d.layer_cnt = len(layers)
    d.sun_altitude_deg, d.sun_azimuth_deg, d.sun_diameter_deg = sun
    d.geom_pool_size = pool
    for li, (prob, pops) in enumerate(layers):
        d.layers[li].prob = prob
        d.layers[li].population_cnt = len(pops)
        for ci, p in enumerate(pops):
            d.layers[li].populations[ci] = p
    return d


def render(lens="fisheye_equal_area", fov=120.0, res=(1920, 1080), view=(0.0, 30.0, 0.0), visible="upper",
           shift=(0, 0), overlap=0.0):
    return A.HbRenderDesc(A.LENS[lens], fov, res[0], res[1], view[0], view[1], view[2], A.VISIBLE[visible],
                          shift[0], shift[1], overlap)


# BASELINE.json configs, at parity-test sizes
CASES = {
    # config 1/2: examples/config_example.json crystal 3, render id 4
    "column_config2": dict(scene=lambda: scene([(0.0, [prism_pop(1.3, zenith=("gauss", 90, 0.3), cid=3)])], 7),
                           render=lambda: render(), wl=[550.0]),
    # config 3: plate parhelia + raypath filter [3,5] symmetry P
    "plate_filter_config3": dict(
        scene=lambda: scene([(0.0, [prism_pop(0.3, zenith=("gauss", 0, 0.8), cid=6,
                                              filt=raypath_filter([3, 5], "P"))])], 7),
        render=lambda: render(), wl=[550.0]),
    # config 4: two layers, plate (prob 1.0) over random column
    "two_layer_config4": dict(
        scene=lambda: scene([(1.0, [prism_pop(0.3, zenith=("gauss", 0, 0.8), cid=6)]),
                             (0.0, [prism_pop(1.3, zenith=("uniform", 90, 360), cid=3)])], 7),
        render=lambda: render(), wl=[550.0]),
    # config 5: stochastic prism geometry, rectangular full-sky render, max_hits 8
    "stoch_config5": dict(
        scene=lambda: scene([(0.0, [prism_pop(1.0, zenith=("uniform", 90, 360), cid=1,
                                              face_dist=[dist("gauss", 1.0, 0.15)] * 6)])], 8, pool=64),
        render=lambda: render("rectangular", 360.0, (2048, 1024), (0.0, 90.0, 0.0), "full"), wl=[550.0]),
    # config_example.json filter 7: complex composition [raypath | (raypath & crystal) | direction-out]
    "complex_filter": dict(
        scene=lambda: scene([(0.0, [prism_pop(1.3, zenith=("gauss", 90, 5.0), cid=3, filt=complex_filter(
            [[dict(kind=1, path=[3, 5])], [dict(kind=1, path=[1, 3, 2]), dict(kind=4, crystal_id=3)],
             [dict(kind=2, entry=3, exit=6), dict(kind=3, lon=180.0, lat=25.0, radii=40.0)]], "PB"))])], 6),
        render=lambda: render(res=(960, 540)), wl=[570.0]),
    # examples/config_example.json ships several renderers for one scene: N projections of one trace
    "multi_render": dict(
        scene=lambda: scene([(0.0, [prism_pop(1.3, zenith=("gauss", 90, 0.3))])], 7),
        render=lambda: [render(), render("linear", 60.0, (800, 600), view=(0.0, 20.0, 0.0)),
                        render("dual_fisheye_equal_area", 180.0, (1024, 512), visible="full"),
                        render("fisheye_orthographic", 170.0, (640, 640), view=(90.0, 90.0, 0.0))],
        wl=[450.0, 610.0]),
    # raypath colour: two layers so masks are carried through the continuation pool; predicates
    # of every kind, two symmetry groups on the first population; classes with "any" and "all" combines
    "color_classes": dict(
        scene=lambda: color_classes(scene([
            (0.4, [color_pred(color_pred(color_pred(color_pred(
                prism_pop(1.3, zenith=("gauss", 90, 20.0), cid=3),
                0, "PBD", kind=1, path=[3, 5]), 1, "PBD", kind=2, entry=1, exit=3), 2, "", kind=0),
                3, "", kind=3, lon=180.0, lat=-20.0, radii=30.0),
                   color_pred(prism_pop(0.3, zenith=("gauss", 0, 10.0), cid=6), 4, "P", kind=1, path=[1, 3, 2])]),
            (0.0, [color_pred(color_pred(prism_pop(1.0, zenith=("uniform", 90, 360), cid=9),
                                         5, "B", kind=2, entry=3, exit=5), 6, "", kind=4, crystal_id=9)])], 6),
            [(0b0000011, "any"), (0b0100100, "all"), (0b1000000, "any"), (0b0011000, "any"), (0, "any")]),
        render=lambda: render(res=(640, 360)), wl=[550.0]),
    "pyramid": dict(scene=lambda: scene([(0.0, [pyramid_pop()])], 8),
                    render=lambda: render("dual_fisheye_equal_area", 120.0, (1024, 512), visible="full", overlap=0.1),
                    wl=[610.0]),
    "two_populations": dict(
        scene=lambda: scene([(0.0, [prism_pop(1.3, zenith=("gauss", 90, 0.3), cid=3, proportion=10.0),
                                    pyramid_pop(proportion=3.0)])], 6),
        render=lambda: render("linear", 40.0, (640, 480), (-50.0, 30.0, 0.0)), wl=[490.0]),
    # ---- the lens types no other case reaches (projection_shared.h:196-375) ----
    "lens_stereographic": dict(scene=lambda: scene([(0.0, [prism_pop(1.3, zenith=("gauss", 90, 0.3), cid=3)])], 6),
                               render=lambda: render("fisheye_stereographic", 140.0, (800, 800), (0.0, 60.0, 0.0)),
                               wl=[550.0]),
    "lens_dual_equidistant": dict(scene=lambda: scene([(0.0, [prism_pop(0.3, zenith=("gauss", 0, 1.0), cid=6)])], 6),
                                  render=lambda: render("dual_fisheye_equidistant", 180.0, (1024, 512), visible="full",
                                                        overlap=0.05), wl=[490.0]),
    "lens_dual_stereographic": dict(scene=lambda: scene([(0.0, [prism_pop(1.0, zenith=("uniform", 90, 360))])], 6),
                                    render=lambda: render("dual_fisheye_stereographic", 180.0, (1024, 512),
                                                          visible="full", overlap=0.1), wl=[610.0]),
    "lens_dual_orthographic": dict(scene=lambda: scene([(0.0, [prism_pop(1.3, zenith=("gauss", 90, 0.3), cid=3)])], 6),
                                   render=lambda: render("dual_fisheye_orthographic", 180.0, (1024, 512), visible="full",
                                                         overlap=0.08), wl=[570.0]),
    "lens_globe": dict(scene=lambda: scene([(0.0, [prism_pop(1.0, zenith=("uniform", 90, 360))])], 6),
                       render=lambda: render("globe", 60.0, (900, 900), (30.0, 20.0, 10.0), visible="full"), wl=[530.0]),
    # filter_out action (action 1: the filter REMOVES what it matches) on a raypath filter with full symmetry
    "filter_out_raypath": dict(
        scene=lambda: scene([(0.0, [prism_pop(1.3, zenith=("gauss", 90, 5.0), cid=3,
                                              filt=raypath_filter([3, 5], "PBD", action=1))])], 6),
        render=lambda: render(res=(960, 540)), wl=[550.0]),
    # D-symmetry PHYSICAL filters: raypath and entry-exit under the dihedral reduction only / with P and B
    "filter_d_symmetry": dict(
        scene=lambda: scene([(0.0, [prism_pop(1.3, zenith=("gauss", 90, 10.0), cid=3, proportion=2.0,
                                              filt=raypath_filter([3, 1, 5], "D")),
                                    prism_pop(0.4, zenith=("gauss", 0, 10.0), roll=("gauss", 60.0, 20.0), cid=6,
                                              proportion=1.0, filt=complex_filter([[dict(kind=2, entry=1, exit=4)],
                                                                   [dict(kind=1, path=[4, 2, 6])]], "BD"))])], 6),
        render=lambda: render(res=(960, 540)), wl=[550.0]),
    "partial_prob": dict(
        scene=lambda: scene([(0.5, [prism_pop(1.0, zenith=("uniform", 90, 360), cid=1)]),
                             (0.0, [prism_pop(0.5, zenith=("gauss", 0, 2.0), cid=2)])], 5),
        render=lambda: render("fisheye_equidistant", 180.0, (512, 512), (0.0, 90.0, 0.0)), wl=[530.0]),
}
