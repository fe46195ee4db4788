"""Lumice JSON config -> scene / render descriptions (HbSceneDesc, HbRenderDesc).

Mirrors the reference's config front door for the fields the trace path consumes
(reference: src/config/crystal_config.cpp from_json, src/core/math.cpp:590-725 axis parsing,
src/config/filter_config.cpp, src/config/render_config.cpp, src/config/config_manager.cpp).
Complex ("composition") filters are supported up to 8 OR-terms of 4 AND-factors; `raypath_color` classes up to
16 classes / 64 component bits (config/raypath_color_config.cpp, color_gate_table.cpp, color_class_table.hpp).
"""
import json
import math

from . import _abi as A

_SYM = {"P": 1, "B": 2, "D": 4}


def _dist(obj, default=None):
    """JSON number or {type, mean, std} -> HbDist (math.cpp from_json(Distribution))."""
    if obj is None:
        return default
    if isinstance(obj, (int, float)):
        return A.HbDist(A.DIST["none"], float(obj), 0.0)
    t = obj.get("type")
    if t not in A.DIST:
        raise ValueError(f"distribution needs a valid 'type' (got {t!r})")
    return A.HbDist(A.DIST[t], float(obj.get("mean", 0.0)), float(obj.get("std", 0.0)))


def _miller_to_alpha(i1, i4):
    if i1 == 0:
        return 28.0
    return math.degrees(math.atan(0.866025403784 * i4 / i1 / 1.629))


def _sync_groups(d, sg):
    """Optional "sync_group" sub-map of a crystal shape (ReadSyncGroupJson + PrepareSyncGroups,
    crystal_config.cpp:47-146,179-205): scalars sharing a non-zero id share one draw per crystal instance.
    Slots this crystal type lacks and single-member groups are zeroed, survivors renumbered 1..N by first
    appearance, and every member takes its group leader's distribution."""
    slots = [0] * 10
    if sg:
        keys = {"height": 0} if d.kind == 0 else {"upper_h": 1, "prism_h": 2, "lower_h": 3}
        for k, slot in keys.items():
            if k in sg:
                slots[slot] = int(sg[k])
        for i, e in enumerate(list(sg.get("face_distance", []))[:6]):
            slots[4 + i] = int(e)
    owned = ([0] if d.kind == 0 else [1, 2, 3]) + list(range(4, 10))
    slots = [g if i in owned else 0 for i, g in enumerate(slots)]
    slots = [g if g != 0 and slots.count(g) > 1 else 0 for g in slots]
    remap = {}
    for i, g in enumerate(slots):
        if g != 0:
            slots[i] = remap.setdefault(g, len(remap) + 1)

    def dist_of(slot):
        return d.face_dist[slot - 4] if slot >= 4 else d.height[0 if slot == 0 else slot - 1]

    def set_dist(slot, v):
        if slot >= 4:
            d.face_dist[slot - 4] = v
        else:
            d.height[0 if slot == 0 else slot - 1] = v

    leader = {}
    for i, g in enumerate(slots):
        if g == 0:
            continue
        if g in leader:
            src = dist_of(leader[g])
            set_dist(i, A.HbDist(src.type, src.center, src.spread))
        else:
            leader[g] = i
    for i, g in enumerate(slots):
        d.sync_group[i] = g


def crystal_desc(c):
    d = A.HbCrystalDesc()
    d.id = int(c["id"])
    shape = c.get("shape", {})
    ones = A.HbDist(0, 1.0, 0.0)
    for i in range(6):
        d.face_dist[i] = ones
    fd = shape.get("face_distance")
    if fd is not None:
        for i, e in enumerate(fd[:6]):
            d.face_dist[i] = _dist(e)
    if c["type"] == "prism":
        d.kind = 0
        d.height[0] = _dist(shape.get("height"), A.HbDist(0, 1.0, 0.0))
    elif c["type"] == "pyramid":
        d.kind = 1
        d.height[0] = _dist(shape.get("upper_h"), A.HbDist(0, 0.0, 0.0))
        d.height[1] = _dist(shape["prism_h"])
        d.height[2] = _dist(shape.get("lower_h"), A.HbDist(0, 0.0, 0.0))
        for upper in (True, False):
            key = "upper" if upper else "lower"
            alpha = 28.0
            if f"{key}_wedge_angle" in shape:
                alpha = float(shape[f"{key}_wedge_angle"])
            elif f"{key}_indices" in shape and len(shape[f"{key}_indices"]) == 3:
                idx = shape[f"{key}_indices"]
                alpha = _miller_to_alpha(int(idx[0]), int(idx[2]))
            if upper:
                d.wedge_upper_deg = alpha
            else:
                d.wedge_lower_deg = alpha
    else:
        raise ValueError(f"unknown crystal type {c['type']!r}")
    _sync_groups(d, shape.get("sync_group"))
    axis = c.get("axis")
    if axis is None:  # AxisDistribution default: zenith 0 (latitude 90), no randomness (math.cpp:536-538)
        d.latitude = A.HbDist(0, 90.0, 0.0)
        d.azimuth = A.HbDist(0, 0.0, 0.0)
        d.roll = A.HbDist(0, 0.0, 0.0)
    else:
        z = _dist(axis["zenith"])
        d.latitude = A.HbDist(z.type, 90.0 - z.center, z.spread)  # zenith -> latitude
        d.azimuth = _dist(axis.get("azimuth"), A.HbDist(1, 0.0, 360.0))
        d.roll = _dist(axis.get("roll"), A.HbDist(1, 0.0, 360.0))
    return d


def _simple_filter(f, s):
    """Fill an HbSimpleFilterSpec from one simple filter entry of the config."""
    t = f.get("type", "none")
    s.entry_fn = -1
    s.exit_fn = -1
    if t == "none":
        s.kind = 0
    elif t == "raypath":
        rp = f["raypath"]
        if len(rp) > A.HB_MAX_FILTER_PATH:
            raise ValueError("raypath filter longer than 32 faces")
        s.kind = 1
        s.path_len = len(rp)
        for i, x in enumerate(rp):
            s.path[i] = int(x)
    elif t == "entry_exit":
        s.kind = 2
        s.entry_fn = int(f["entry"]) if "entry" in f else -1
        s.exit_fn = int(f["exit"]) if "exit" in f else -1
        s.min_len = int(f.get("min_len", 1))
        s.max_len = int(f.get("max_len", 0))
    elif t == "direction":
        s.kind = 3
        s.lon_deg, s.lat_deg, s.radii_deg = float(f["az"]), float(f["el"]), float(f["radii"])
    elif t == "crystal":
        s.kind = 4
        s.crystal_id = int(f["crystal_id"])
    else:
        raise ValueError(f"filter type {t!r} cannot be used here")


def filter_desc(f, all_filters=None):
    """FilterConfig (filter_config.cpp from_json). A "complex" filter is an OR over its `composition`
    entries, each a filter id or a list of ids that are AND-ed; the sub-filters contribute their parameters
    only (symmetry and action come from the complex filter itself, device_filter_desc.cpp:146-166)."""
    d = A.HbFilterSpecDesc()
    d.simple.entry_fn = d.simple.exit_fn = -1
    if f is None or f.get("type", "none") == "none":
        return d
    d.action = 1 if f.get("action", "filter_in") == "filter_out" else 0
    d.symmetry = sum(_SYM[ch] for ch in f.get("symmetry", "") if ch in _SYM)
    if f["type"] == "complex":
        comp = f["composition"]
        if all_filters is None:
            raise ValueError("complex filter needs the filter table")
        if len(comp) > A.HB_MAX_FILTER_TERMS:
            raise ValueError("complex filter with more than 8 OR-terms is not supported by this backend")
        d.kind = 5
        d.term_cnt = len(comp)
        for o, term in enumerate(comp):
            ids = term if isinstance(term, list) else [term]
            if len(ids) > 4:
                raise ValueError("complex filter term with more than 4 AND-factors is not supported")
            d.term_len[o] = len(ids)
            for a, fid in enumerate(ids):
                sub = all_filters[int(fid)]
                if sub.get("type") == "complex":
                    raise ValueError("complex filters cannot nest")
                _simple_filter(sub, d.terms[o][a])
        return d
    _simple_filter(f, d.simple)
    d.kind = d.simple.kind
    return d


def render_desc(r):
    lens = r.get("lens", {})
    lens_type = A.LENS[lens.get("type", "linear")]
    if "fov" in lens:
        fov = float(lens["fov"])
    elif "f" in lens:  # 35 mm focal length -> fov, half short edge d = 12 mm (render_config.cpp:60-110)
        f, d = float(lens["f"]), 12.0
        name = lens.get("type", "linear")
        if name == "linear":
            fov = 2.0 * math.degrees(math.atan2(d, f))
        elif name.endswith("equal_area"):
            fov = 4.0 * math.degrees(math.asin(d / (2 * f)))
        elif name.endswith("equidistant"):
            fov = math.degrees(d / f)
        elif name.endswith("stereographic"):
            fov = 4.0 * math.degrees(math.atan(d / (2 * f)))
        elif name.endswith("orthographic"):
            fov = 2.0 * math.degrees(math.asin(d / f))
        else:
            fov = 0.0
    else:
        fov = 90.0
    res = r.get("resolution", [1920, 1080])
    view = r.get("view", {})
    shift = r.get("lens_shift", [0, 0])
    return A.HbRenderDesc(lens_type, fov, int(res[0]), int(res[1]), float(view.get("azimuth", 0.0)),
                          float(view.get("elevation", 0.0)), float(view.get("roll", 0.0)),
                          A.VISIBLE[r.get("visible", "upper")], int(shift[0]), int(shift[1]),
                          float(r.get("overlap", 0.0)))


def _pred_key(ref):
    """Structural identity of a colour predicate (SimpleFilterParam operator==, config_compare.hpp)."""
    t = ref.get("type", "none")
    if t == "raypath":
        return (t, tuple(int(x) for x in ref["raypath"]))
    if t == "entry_exit":
        return (t, ref.get("entry"), ref.get("exit"), int(ref.get("min_len", 1)), int(ref.get("max_len", 0)))
    if t == "direction":
        return (t, float(ref["az"]), float(ref["el"]), float(ref["radii"]))
    if t == "crystal":
        return (t, int(ref["crystal_id"]))
    return ("none",)


def apply_raypath_color(rc, d):
    """`raypath_color` -> per-population colour predicates with component bits + the class table.

    BuildColorGateTable (color_gate_table.cpp:47-95): walk classes[].match[] in order; a ref must name exactly
    one population of its layer; (layer, crystal, predicate, symmetry) duplicates share a bit; bits are handed
    out in first-occurrence order, predicates past bit 63 get none. BuildColorClassTable: a class's bit set is
    the union of its refs' bits, `combine` is "any" or "all". Returns the class list (colour, combine, visible)."""
    if rc is None:
        return []
    classes = rc.get("classes", []) if isinstance(rc, dict) else rc
    if len(classes) > A.HB_MAX_COLOR_CLASSES:
        raise ValueError(f"raypath_color: more than {A.HB_MAX_COLOR_CLASSES} colour classes")
    seen = {}
    next_bit = 0
    out = []
    allm = 0
    for ci, cls in enumerate(classes):
        combine = cls.get("combine", "any")
        if combine not in ("any", "all"):
            raise ValueError(f"raypath_color: unknown combine {combine!r}")
        bits = 0
        for ref in cls["match"]:
            layer, cid = int(ref["layer"]), int(ref["crystal"])
            if not 0 <= layer < d.layer_cnt:
                raise ValueError(f"raypath_color: layer index out of range (layer={layer}, crystal_id={cid})")
            ld = d.layers[layer]
            hits = [k for k in range(ld.population_cnt) if ld.populations[k].crystal.id == cid]
            if len(hits) != 1:
                raise ValueError(f"raypath_color: crystal_id {cid} matches {len(hits)} scattering settings on layer {layer}")
            sym = sum(_SYM[ch] for ch in ref.get("symmetry", "") if ch in _SYM)
            key = (layer, cid, _pred_key(ref), sym)
            if key not in seen:
                bit = next_bit if next_bit < 64 else 255
                next_bit += 1 if next_bit < 64 else 0
                seen[key] = bit
                pop = ld.populations[hits[0]]
                if pop.color_pred_cnt >= A.HB_MAX_COLOR_PREDS:
                    raise ValueError("raypath_color: more than 32 predicates on one population")
                cp = pop.color_preds[pop.color_pred_cnt]
                _simple_filter(ref, cp.pred)
                cp.symmetry = sym
                cp.bit = bit
                pop.color_pred_cnt += 1
            if seen[key] < 64:
                bits |= 1 << seen[key]
        d.color_classes.bits[ci] = bits
        if combine == "all":
            allm |= 1 << ci
        out.append(dict(color=tuple(float(x) for x in cls["color"]), combine=combine,
                        visible=bool(cls.get("visible", True)), solo=bool(cls.get("solo", False)), bits=bits))
    d.color_classes.class_cnt = len(classes)
    d.color_classes.combine_all_mask = allm
    return out


class SceneConfig:
    """Parsed config: scene description + renders + spectrum + ray count (per wavelength)."""

    def __init__(self, desc, renders, spectrum, ray_num_total, illuminant=None, color_classes=None):
        self.desc = desc
        self.renders = renders          # {id: HbRenderDesc}
        self.spectrum = spectrum        # [(wavelength_nm, weight)]
        self.ray_num_total = ray_num_total
        self.illuminant = illuminant
        self.color_classes = color_classes or []   # [{color, combine, visible, solo, bits}] (raypath_color)

    def rays_per_wavelength(self):
        """ray_num is the total across wavelengths; per-wavelength = ceil(total / N)
        (reference: src/server/ray_num_semantics.hpp:12-16)."""
        n = max(1, len(self.spectrum))
        return -(-self.ray_num_total // n)


def load_config(path_or_dict, geom_pool_size=1):
    cfg = path_or_dict
    if not isinstance(cfg, dict):
        with open(path_or_dict) as f:
            cfg = json.load(f)
    crystals = {int(c["id"]): c for c in cfg.get("crystal", [])}
    filters = {int(f["id"]): f for f in cfg.get("filter", [])}
    scene = cfg["scene"]
    ls = scene["light_source"]
    d = A.HbSceneDesc()
    d.max_hits = int(scene.get("max_hits", 8))
    d.sun_altitude_deg = float(ls.get("altitude", 0.0))
    d.sun_azimuth_deg = float(ls.get("azimuth", 0.0))
    d.sun_diameter_deg = float(ls.get("diameter", 0.5))
    d.geom_pool_size = int(geom_pool_size)
    layers = scene["scattering"]
    if not 1 <= len(layers) <= A.HB_MAX_LAYERS:
        raise ValueError("scene needs 1..8 scattering layers")
    d.layer_cnt = len(layers)
    for li, layer in enumerate(layers):
        ld = d.layers[li]
        ld.prob = float(layer.get("prob", 0.0))
        entries = layer["entries"]
        if not 1 <= len(entries) <= A.HB_MAX_CRYSTALS:
            raise ValueError("a scattering layer needs 1..16 entries")
        ld.population_cnt = len(entries)
        for ci, e in enumerate(entries):
            pop = ld.populations[ci]
            pop.crystal = crystal_desc(crystals[int(e["crystal"])])
            pop.filter = filter_desc(filters.get(int(e["filter"])) if "filter" in e else None, filters)
            pop.proportion = float(e.get("proportion", 1.0))
    spectrum = ls.get("spectrum", [])
    illuminant = None
    if isinstance(spectrum, str):
        illuminant = spectrum
        spectrum = []
    spec = [(float(s["wavelength"]), float(s.get("weight", 1.0))) for s in spectrum]
    renders = {int(r["id"]): render_desc(r) for r in cfg.get("render", [])}
    classes = apply_raypath_color(cfg.get("raypath_color"), d)
    return SceneConfig(d, renders, spec, int(scene.get("ray_num", 0)), illuminant, classes)
