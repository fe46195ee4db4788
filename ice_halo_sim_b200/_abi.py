"""ctypes mirror of include/halotrace_b200.h (the C ABI of the B200 trace engine).

Field order and sizes must match the header exactly; tests/test_abi.py checks sizeof() of every
struct against the values the C side reports and that every declared symbol is exported.
"""
import ctypes as C

HB_MAX_FACES = 20
HB_MAX_FACE_VTX = 12
HB_MAX_SUBTRIS = 64
HB_MAX_HITS = 64
HB_MAX_LAYERS = 8
HB_MAX_CRYSTALS = 16
HB_LUT_NODES = 257
HB_MAX_WL = 256
HB_INVALID_FACE = 0xFFFF
HB_MAX_FILTER_PATH = 32
HB_MAX_FILTER_TERMS = 8
HB_MAX_RENDERS = 8
HB_MAX_COLOR_GROUPS = 4
HB_MAX_COLOR_CLASSES = 16
HB_MAX_COLOR_PREDS = 32

HB_OK = 0
STATUS_NAMES = {0: "HB_OK", -1: "HB_ERR_INVALID_ARG", -2: "HB_ERR_NO_DEVICE", -3: "HB_ERR_CUDA",
                -4: "HB_ERR_STATE", -5: "HB_ERR_CAPACITY", -6: "HB_ERR_UNSUPPORTED", -7: "HB_ERR_COMM"}

DIST = {"none": 0, "no_random": 0, "uniform": 1, "gauss": 2, "zigzag": 3, "laplacian": 4, "gauss_legacy": 5}
LENS = {"linear": 0, "fisheye_equal_area": 1, "fisheye_equidistant": 2, "fisheye_stereographic": 3,
        "dual_fisheye_equal_area": 4, "dual_fisheye_equidistant": 5, "dual_fisheye_stereographic": 6,
        "rectangular": 7, "fisheye_orthographic": 8, "dual_fisheye_orthographic": 9, "globe": 10}
VISIBLE = {"upper": 0, "lower": 1, "full": 2}
LAT_FULL_SPHERE, LAT_NO_RANDOM, LAT_GAUSS_LEGACY, LAT_LUT = 0, 1, 3, 6

f32, u32, i32, u8, u16, u64, f64 = C.c_float, C.c_uint32, C.c_int32, C.c_uint8, C.c_uint16, C.c_uint64, C.c_double


class HbCrystalTables(C.Structure):
    _fields_ = [("face_cnt", u32), ("subtri_cnt", u32),
                ("plane", (f32 * 4) * HB_MAX_FACES),
                ("tri_v", (f32 * 9) * HB_MAX_SUBTRIS),
                ("tri_n", (f32 * 3) * HB_MAX_SUBTRIS),
                ("tri_area", f32 * HB_MAX_SUBTRIS),
                ("tri_face", u8 * HB_MAX_SUBTRIS),
                ("face_fn", u8 * HB_MAX_FACES),
                ("reserved_", u8 * 12)]


class HbAxisSampler(C.Structure):
    _fields_ = [("lat_path", u32), ("lat_mean", f32), ("lat_std", f32),
                ("az_type", u32), ("az_mean", f32), ("az_std", f32),
                ("roll_type", u32), ("roll_mean", f32), ("roll_std", f32),
                ("lut_n", u32),
                ("lut_theta", f32 * HB_LUT_NODES), ("lut_cdf", f32 * HB_LUT_NODES), ("lut_flip", f32 * HB_LUT_NODES)]


class HbSimpleFilter(C.Structure):
    _fields_ = [("kind", u32), ("path_len", u32), ("path", u8 * HB_MAX_FILTER_PATH),
                ("entry_fn", i32), ("exit_fn", i32), ("min_len", u32), ("max_len", u32),
                ("dir", f32 * 3), ("cos_radii", f32), ("crystal_id", u32)]


class HbFilterDesc(C.Structure):
    _fields_ = [("kind", u32), ("action", u32), ("symmetry", u32), ("fn_period", i32), ("sigma_a", i32),
                ("d_applicable", u32), ("simple", HbSimpleFilter), ("term_cnt", u32),
                ("term_len", u32 * HB_MAX_FILTER_TERMS),
                ("terms", (HbSimpleFilter * 4) * HB_MAX_FILTER_TERMS)]


class HbColorGroup(C.Structure):
    _fields_ = [("filter", HbFilterDesc), ("bit", u8 * HB_MAX_FILTER_TERMS)]


class HbColorClasses(C.Structure):
    _fields_ = [("class_cnt", u32), ("combine_all_mask", u32), ("bits", u64 * HB_MAX_COLOR_CLASSES)]


class HbCrystalPopulation(C.Structure):
    _fields_ = [("proportion", f32), ("crystal_id", u32), ("shape_cnt", u32),
                ("shapes", C.POINTER(HbCrystalTables)),
                ("axis", HbAxisSampler), ("filter", HbFilterDesc),
                ("color_group_cnt", u32), ("reserved_", u32), ("color_groups", HbColorGroup * HB_MAX_COLOR_GROUPS)]


class HbLayer(C.Structure):
    _fields_ = [("prob", f32), ("population_cnt", u32), ("populations", C.POINTER(HbCrystalPopulation))]


class HbScene(C.Structure):
    _fields_ = [("max_hits", u32), ("layer_cnt", u32), ("layers", C.POINTER(HbLayer)),
                ("sun_lon", f32), ("sun_lat", f32), ("sun_half_angle", f32), ("reserved_", u32),
                ("color_classes", HbColorClasses)]


class HbWlEntry(C.Structure):
    _fields_ = [("n_idx", f32), ("spd_weight", f32), ("cmf_x", f32), ("cmf_y", f32), ("cmf_z", f32)]


class HbProjParams(C.Structure):
    _fields_ = [("proj_type", i32), ("img_w", i32), ("img_h", i32), ("visible_range", i32),
                ("lens_shift_x", i32), ("lens_shift_y", i32),
                ("scale", f32), ("az0", f32), ("r_scale", f32), ("max_abs_dz", f32), ("rot", f32 * 9)]


class HbExitRecord(C.Structure):
    _fields_ = [("dir", f32 * 3), ("weight", f32), ("path_len", u8), ("path", u8 * 64), ("pad0_", u8),
                ("crystal_id", u16), ("ms_layer_idx", u8), ("wl_idx", u8), ("pad1_", u8 * 2),
                ("component_mask", u64)]


class HbSessionSpec(C.Structure):
    _fields_ = [("seed", u32), ("wl_cnt", u32), ("wl", C.POINTER(HbWlEntry)), ("ray_num", u64),
                ("record_exits", u32), ("accumulate", u32), ("ray_base", u64), ("use_ray_base", u32),
                ("reserved_", u32)]


class HbLayerStats(C.Structure):
    _fields_ = [("root_count", u64), ("continuation_count", u64), ("exit_count", u64), ("exit_w_sum", f64)]


class HbCounters(C.Structure):
    _fields_ = [("kernel_launches", u64), ("rays_traced", u64), ("last_layer_ms", f64),
                ("intersect_ms", f64), ("optics_ms", f64), ("gen_ms", f64),
                ("intersect_launches", u64), ("optics_launches", u64), ("gen_launches", u64),
                ("intersect_rays", u64), ("optics_rays", u64),
                ("bounce_ms", f64), ("bounce_launches", u64), ("bounce_rays", u64),
                ("genbounce_ms", f64), ("genbounce_launches", u64), ("genbounce_rays", u64)]


class HbDist(C.Structure):
    _fields_ = [("type", u32), ("center", f32), ("spread", f32)]


class HbCrystalDesc(C.Structure):
    _fields_ = [("kind", u32), ("id", u32), ("height", HbDist * 3), ("face_dist", HbDist * 6),
                ("wedge_upper_deg", f32), ("wedge_lower_deg", f32),
                ("latitude", HbDist), ("azimuth", HbDist), ("roll", HbDist), ("sync_group", i32 * 10)]


class HbSimpleFilterSpec(C.Structure):
    _fields_ = [("kind", u32), ("path_len", u32), ("path", u8 * HB_MAX_FILTER_PATH), ("entry_fn", i32),
                ("exit_fn", i32), ("min_len", u32), ("max_len", u32), ("lon_deg", f32), ("lat_deg", f32),
                ("radii_deg", f32), ("crystal_id", u32)]


class HbFilterSpecDesc(C.Structure):
    _fields_ = [("kind", u32), ("action", u32), ("symmetry", u32), ("simple", HbSimpleFilterSpec),
                ("term_cnt", u32), ("term_len", u32 * HB_MAX_FILTER_TERMS),
                ("terms", (HbSimpleFilterSpec * 4) * HB_MAX_FILTER_TERMS)]


class HbColorPredDesc(C.Structure):
    _fields_ = [("pred", HbSimpleFilterSpec), ("symmetry", u32), ("bit", u32)]


class HbPopulationDesc(C.Structure):
    _fields_ = [("crystal", HbCrystalDesc), ("filter", HbFilterSpecDesc), ("proportion", f32),
                ("color_pred_cnt", u32), ("color_preds", HbColorPredDesc * HB_MAX_COLOR_PREDS)]


class HbLayerDesc(C.Structure):
    _fields_ = [("prob", f32), ("population_cnt", u32), ("populations", HbPopulationDesc * HB_MAX_CRYSTALS)]


class HbSceneDesc(C.Structure):
    _fields_ = [("max_hits", u32), ("layer_cnt", u32),
                ("sun_altitude_deg", f32), ("sun_azimuth_deg", f32), ("sun_diameter_deg", f32),
                ("geom_pool_size", u32), ("layers", HbLayerDesc * HB_MAX_LAYERS), ("color_classes", HbColorClasses)]


class HbRenderDesc(C.Structure):
    _fields_ = [("lens_type", i32), ("fov_deg", f32), ("img_w", i32), ("img_h", i32),
                ("view_az_deg", f32), ("view_el_deg", f32), ("view_ro_deg", f32),
                ("visible_range", i32), ("lens_shift_x", i32), ("lens_shift_y", i32), ("overlap", f32)]


class HbSnapshotDesc(C.Structure):
    _fields_ = [("intensity_factor", f32), ("ray_color", f32 * 3), ("background", f32 * 3)]


ALL_STRUCTS = [HbCrystalTables, HbAxisSampler, HbSimpleFilter, HbFilterDesc, HbColorGroup, HbColorClasses,
               HbCrystalPopulation, HbLayer, HbScene,
               HbWlEntry, HbProjParams, HbExitRecord, HbSessionSpec, HbLayerStats, HbCounters, HbDist, HbCrystalDesc,
               HbSimpleFilterSpec, HbFilterSpecDesc, HbColorPredDesc, HbPopulationDesc, HbLayerDesc, HbSceneDesc, HbRenderDesc, HbSnapshotDesc]
