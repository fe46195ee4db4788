/*
 * halotrace_b200.h — C ABI of the B200 (sm_100a) ice-halo trace engine.
 *
 * This is the drop-in boundary. Every entry point mirrors one virtual of the
 * reference's trace-backend seam `lumice::TraceBackend`
 * (reference: src/core/backend/trace_backend.hpp:367-641) or one host-side
 * table builder the reference's GPU backends call before uploading
 * (reference: src/core/backend/cuda_trace_backend.cu:2436-2542).
 * A thin C++ subclass of `lumice::TraceBackend` (see adapter/ and
 * INTEGRATION.md) forwards each virtual to the function named next to it.
 *
 * Conventions
 *   - plain C, POD structs, pointers + sizes; no C++/torch types cross this ABI
 *   - every function returns HB_OK (0) or a negative HbStatus; the message of
 *     the last failure on a handle is available from hb_last_error()
 *   - there is NO CPU fallback: every compute entry point fails with
 *     HB_ERR_NO_DEVICE / HB_ERR_CUDA when no sm_100 device is usable
 *   - angles are radians unless the field name ends in _deg
 *   - rays crossing this ABI are world-space except where a field says
 *     "crystal-local" (parity-only injection/export helpers)
 */
#ifndef HALOTRACE_B200_H_
#define HALOTRACE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_ABI_VERSION 2u

/* Limits (reference: src/core/def.hpp:23-31, crystal.hpp:67,76, pcg_shared.h:71). */
#define HB_MAX_FACES 20u      /* kCrystalGeomMaxFaces */
#define HB_MAX_FACE_VTX 12u   /* kCrystalGeomMaxVtxPerFace */
#define HB_MAX_SUBTRIS 64u    /* kMaxTriPerKernel */
#define HB_MAX_HITS 64u       /* kMaxHits */
#define HB_MAX_LAYERS 8u      /* C API scatter-layer cap, lumice.h:286-293 */
#define HB_MAX_CRYSTALS 16u   /* kMaxCrystalNum (per layer) */
#define HB_LUT_NODES 257u     /* LatLut::kNodes, lat_lut.hpp:31-35 */
#define HB_MAX_WL 256u        /* wavelength pool cap (kWlPoolSizeMax=255) */
#define HB_INVALID_FACE 0xFFFFu /* kInvalidId */
#define HB_MAX_FILTER_PATH 32u  /* C API raypath len cap, lumice.h:286-293 */
#define HB_MAX_FILTER_TERMS 8u
#define HB_MAX_RENDERS 8u
#define HB_MAX_COLOR_GROUPS 4u   /* colour-predicate symmetry groups per population (kColorMaxGroupsPerSlot) */
#define HB_MAX_COLOR_CLASSES 16u /* kMaxColorClassesDevice */
#define HB_MAX_COLOR_PREDS 32u   /* colour predicates per population in a scene description */

typedef enum HbStatus {
  HB_OK = 0,
  HB_ERR_INVALID_ARG = -1,
  HB_ERR_NO_DEVICE = -2,   /* maps to lumice::BackendUnavailableError */
  HB_ERR_CUDA = -3,        /* maps to lumice::BackendUnavailableError */
  HB_ERR_STATE = -4,       /* call outside the BeginSession/EndSession bracket */
  HB_ERR_CAPACITY = -5,
  HB_ERR_UNSUPPORTED = -6,
  HB_ERR_COMM = -7
} HbStatus;

/* Distribution types (reference: src/core/math.hpp DistributionType). */
enum { HB_DIST_NO_RANDOM = 0, HB_DIST_UNIFORM = 1, HB_DIST_GAUSSIAN = 2, HB_DIST_ZIGZAG = 3,
       HB_DIST_LAPLACIAN = 4, HB_DIST_GAUSSIAN_LEGACY = 5 };
/* Latitude sampling paths (reference: pcg_shared.h:56-59, lat_path_selection.hpp:39-46). */
enum { HB_LAT_FULL_SPHERE = 0, HB_LAT_NO_RANDOM = 1, HB_LAT_GAUSS_LEGACY = 3, HB_LAT_LUT = 6 };
/* Lens types (reference: config/render_config.hpp LensParam::LensType). */
enum { HB_LENS_LINEAR = 0, HB_LENS_FISHEYE_EQUAL_AREA = 1, HB_LENS_FISHEYE_EQUIDISTANT = 2,
       HB_LENS_FISHEYE_STEREOGRAPHIC = 3, HB_LENS_DUAL_FISHEYE_EQUAL_AREA = 4,
       HB_LENS_DUAL_FISHEYE_EQUIDISTANT = 5, HB_LENS_DUAL_FISHEYE_STEREOGRAPHIC = 6,
       HB_LENS_RECTANGULAR = 7, HB_LENS_FISHEYE_ORTHOGRAPHIC = 8, HB_LENS_DUAL_FISHEYE_ORTHOGRAPHIC = 9,
       HB_LENS_GLOBE = 10 };
enum { HB_VISIBLE_UPPER = 0, HB_VISIBLE_LOWER = 1, HB_VISIBLE_FULL = 2 };

/* ---------------------------------------------------------------------------
 * Tables (what the reference's GPU backends upload per scene)
 * ------------------------------------------------------------------------- */

/* One convex crystal shape: polygon-face planes + the entry-sampling fan table.
 * Replaces Crystal::GetPolygonFaceNormal/Dist/GetFn (crystal.hpp:227,281-288) and
 * detail::BuildEntrySubTris (simulator.cpp:90-129).
 * plane[f] = (nx,ny,nz,d0), unit outward normal, inside <=> n.x + d0 <= 0. */
typedef struct HbCrystalTables {
  uint32_t face_cnt;                      /* compact present faces, <= HB_MAX_FACES */
  uint32_t subtri_cnt;                    /* <= HB_MAX_SUBTRIS; 0 => degenerate crystal */
  float plane[HB_MAX_FACES][4];
  float tri_v[HB_MAX_SUBTRIS][9];         /* 3 corners x xyz (fan 0,k,k+1) */
  float tri_n[HB_MAX_SUBTRIS][3];         /* raw-winding unit normal */
  float tri_area[HB_MAX_SUBTRIS];
  uint8_t tri_face[HB_MAX_SUBTRIS];       /* compact face id of each sub-triangle */
  uint8_t face_fn[HB_MAX_FACES];          /* GetFn: compact face id -> face number 1..8/13..18/23..28 */
  uint8_t reserved_[12];
} HbCrystalTables;

/* Orientation sampler of one crystal population.
 * Replaces the orientation fields of lm_pcg::GenRootKernelParams (pcg_shared.h:150-189)
 * + the three LatLut arrays (lat_lut.hpp:31-35). */
typedef struct HbAxisSampler {
  uint32_t lat_path;                      /* HB_LAT_* */
  float lat_mean, lat_std;                /* radians */
  uint32_t az_type;                       /* HB_DIST_* */
  float az_mean, az_std;
  uint32_t roll_type;
  float roll_mean, roll_std;
  uint32_t lut_n;                         /* HB_LUT_NODES when lat_path == HB_LAT_LUT else 0 */
  float lut_theta[HB_LUT_NODES];
  float lut_cdf[HB_LUT_NODES];
  float lut_flip[HB_LUT_NODES];
} HbAxisSampler;

/* Raypath filter, flattened. Replaces DeviceFilterDesc (device_filter_desc.hpp:91-123).
 * kind: 0 none, 1 raypath, 2 entry-exit, 3 direction, 4 crystal, 5 complex (OR of AND-terms
 * of simple filters, held in `terms`). action: 0 filter_in, 1 filter_out. */
typedef struct HbSimpleFilter {
  uint32_t kind;
  uint32_t path_len;
  uint8_t path[HB_MAX_FILTER_PATH];       /* raypath: canonical (symmetry-reduced) face numbers */
  int32_t entry_fn, exit_fn;              /* entry-exit: -1 = wildcard */
  uint32_t min_len, max_len;              /* entry-exit: max_len 0 = unbounded */
  float dir[3]; float cos_radii;          /* direction: unit vector + cos(radius) */
  uint32_t crystal_id;                    /* crystal filter */
} HbSimpleFilter;

typedef struct HbFilterDesc {
  uint32_t kind;                          /* as HbSimpleFilter.kind; 5 = complex */
  uint32_t action;                        /* 0 in, 1 out */
  uint32_t symmetry;                      /* bit0 P, bit1 B, bit2 D */
  int32_t fn_period;                      /* Crystal::FnPeriod (6 for hex crystals, -1 none) */
  int32_t sigma_a;                        /* D-mirror parameter 0..5 */
  uint32_t d_applicable;
  HbSimpleFilter simple;                  /* kind 1..4 */
  uint32_t term_cnt;                      /* kind 5: number of OR terms */
  uint32_t term_len[HB_MAX_FILTER_TERMS]; /* number of AND factors in each term */
  HbSimpleFilter terms[HB_MAX_FILTER_TERMS][4];
} HbFilterDesc;

/* Raypath-colour predicates of one population that share a symmetry value (reference: ColorSpecGroup,
 * filter_spec.cpp:389-425): `filter` is a complex descriptor whose OR-term k is predicate k (one AND factor),
 * bit[k] the component bit (< 64) a match sets in the exit's component mask; bit >= 64 = no bit
 * (ColorGateEntry::bit_ overflow, color_gate_table.hpp:25-33). Evaluated after the physical filter admits an
 * exit; never changes survival, weight or the gate draw (simulator.cpp:684-712). */
typedef struct HbColorGroup {
  HbFilterDesc filter;
  uint8_t bit[HB_MAX_FILTER_TERMS];
} HbColorGroup;

/* Colour classes (reference: ColorClassTable / ColorGateParams, cuda_trace_backend.cu:311-318): class c is
 * satisfied by an exit whose mask m has (m & bits[c]) != 0 ("any") or == bits[c] ("all", bit c of
 * combine_all_mask); each satisfied class receives the exit's Y in its own W*H lane of render 0. */
typedef struct HbColorClasses {
  uint32_t class_cnt;                     /* 0 = colour off (no lanes, no mask work) */
  uint32_t combine_all_mask;
  uint64_t bits[HB_MAX_COLOR_CLASSES];
} HbColorClasses;

/* One crystal population of a scattering layer (reference: ScatteringSetting, proj_config.hpp). */
typedef struct HbCrystalPopulation {
  float proportion;                       /* crystal_proportion_ */
  uint32_t crystal_id;                    /* CrystalConfig::id_ (carried into exit records) */
  uint32_t shape_cnt;                     /* geometry pool size (1 = deterministic shape) */
  const HbCrystalTables* shapes;          /* [shape_cnt] */
  HbAxisSampler axis;
  HbFilterDesc filter;
  uint32_t color_group_cnt;               /* 0..HB_MAX_COLOR_GROUPS */
  uint32_t reserved_;
  HbColorGroup color_groups[HB_MAX_COLOR_GROUPS];
} HbCrystalPopulation;

typedef struct HbLayer {
  float prob;                             /* MsInfo::prob_ : continue-to-next-layer probability */
  uint32_t population_cnt;
  const HbCrystalPopulation* populations;
} HbLayer;

/* Whole scene (reference: SceneConfig, proj_config.hpp:27-38). */
typedef struct HbScene {
  uint32_t max_hits;                      /* surface interactions incl. entry (CPU semantics) */
  uint32_t layer_cnt;
  const HbLayer* layers;
  float sun_lon;                          /* (azimuth + 180 deg) in rad, simulator.cpp:194-196 */
  float sun_lat;                          /* (-altitude) in rad */
  float sun_half_angle;                   /* diameter / 2 in rad */
  uint32_t reserved_;
  HbColorClasses color_classes;
} HbScene;

/* Wavelength pool entry (reference: WlEntry, backend/wl_pool.hpp:29-36). */
typedef struct HbWlEntry {
  float n_idx, spd_weight, cmf_x, cmf_y, cmf_z;
} HbWlEntry;

/* Projection parameters; field-for-field lm_proj::ProjParams (projection_shared.h:106-118). */
typedef struct HbProjParams {
  int32_t proj_type, img_w, img_h, visible_range, lens_shift_x, lens_shift_y;
  float scale, az0, r_scale, max_abs_dz;
  float rot[9];
} HbProjParams;

/* Exit ray record; byte-for-byte lumice::ExitRayRecord (exit_seam.hpp:40-53), 96 B. */
typedef struct HbExitRecord {
  float dir[3];
  float weight;
  uint8_t path_len;
  uint8_t path[64];                       /* face numbers (GetFn), entry first */
  uint8_t pad0_;
  uint16_t crystal_id;
  uint8_t ms_layer_idx;
  uint8_t wl_idx;
  uint8_t pad1_[2];
  uint64_t component_mask;
} HbExitRecord;

typedef struct HbSessionSpec {            /* reference: SessionSpec, trace_backend.hpp:197-215 */
  uint32_t seed;                          /* never 0 on the reference's call path */
  uint32_t wl_cnt;                        /* 1 = discrete wavelength session; >1 = per-ray pool draw */
  const HbWlEntry* wl;                    /* [wl_cnt] */
  uint64_t ray_num;                       /* hint (pool sizing) */
  uint32_t record_exits;                  /* 1 => materialise exit records for hb_drain_exits AND keep the generated
                                             roots for hb_export_roots (parity harness); 2 => exit records only: the
                                             exit-seam egress of TraceBackend::DrainExits (trace_backend.hpp:391-446) */
  uint32_t accumulate;                    /* 1 => fused projection + XYZ accumulate (production) */
  uint64_t ray_base;                      /* global index of this session's first root ray when
                                             use_ray_base != 0 (multi-GPU sharding: disjoint index ranges per
                                             rank); otherwise the engine's own monotone counter
                                             (cuda_trace_backend.cu:3717-3754) is used */
  uint32_t use_ray_base;
  uint32_t reserved_;
} HbSessionSpec;

typedef struct HbLayerStats {             /* reference: LayerStats + LayerHandle::ContinuationCount */
  uint64_t root_count;
  uint64_t continuation_count;
  uint64_t exit_count;
  double exit_w_sum;
} HbLayerStats;

typedef struct HbCounters {               /* measurement helpers (bench.py) */
  uint64_t kernel_launches;               /* kernels of this library launched since hb_create */
  uint64_t rays_traced;                   /* root rays over all layers */
  double last_layer_ms;                   /* CUDA-event time of the last hb_trace_layer */
  double intersect_ms, optics_ms, gen_ms; /* accumulated per-kernel-family CUDA-event time (profiling mode) */
  uint64_t intersect_launches, optics_launches, gen_launches;
  uint64_t intersect_rays, optics_rays;   /* ray-bounces processed by each family */
  double bounce_ms;                       /* fused bounce kernels (one launch per interaction) */
  uint64_t bounce_launches, bounce_rays;
  double genbounce_ms;                    /* root generation fused with the entry interaction */
  uint64_t genbounce_launches, genbounce_rays;
} HbCounters;

typedef struct HbEngine HbEngine;         /* opaque; one per TraceBackend instance / GPU */

/* ---------------------------------------------------------------------------
 * Lifetime + the TraceBackend virtuals
 * ------------------------------------------------------------------------- */
uint32_t hb_abi_version(void);
const char* hb_last_error(const HbEngine* h);         /* h may be NULL: last create failure */

/* CudaTraceBackend ctor/probe (cuda_trace_backend.cu:147-193): fails without an sm_100 device. */
int hb_create(int device_ordinal, HbEngine** out);
void hb_destroy(HbEngine* h);

/* Scene + render tables; the engine deep-copies everything (BeginSession's captured
 * spec.scene/spec.render, trace_backend.hpp:118-122). May be called between sessions. */
int hb_set_scene(HbEngine* h, const HbScene* scene);
int hb_set_render(HbEngine* h, const HbProjParams* proj);
/* N projections of ONE trace (examples/config_example.json ships four renderers, which the
 * reference's device backends refuse, server.cpp:402-437): every emitted exit is projected through each of
 * the n <= HB_MAX_RENDERS lenses into its own accumulator. Render 0 is the one hb_readback_xyz drains. */
int hb_set_renders(HbEngine* h, uint32_t n, const HbProjParams* proj);

/* TraceBackend::BeginSession (trace_backend.hpp:374). */
int hb_begin_session(HbEngine* h, const HbSessionSpec* spec);
/* TraceBackend::TraceLayer (trace_backend.hpp:384). First call of a session: n_roots > 0 =
 * RootRaySource::FromHost{count} with null d/p/w/tf (engine generates roots); later calls:
 * n_roots == 0 = RootRaySource::FromDevice (continuations of the preceding hb_recombine). */
int hb_trace_layer(HbEngine* h, uint64_t n_roots, HbLayerStats* stats);
/* TraceBackend::Recombine (trace_backend.hpp:389). */
int hb_recombine(HbEngine* h, int shuffle, uint64_t* continuation_count);
/* TraceBackend::EndSession (trace_backend.hpp:507). */
int hb_end_session(HbEngine* h);
/* TraceBackend::ReadbackXyzAccum (trace_backend.hpp:466): copies W*H*3 floats, ADDS the landed
 * weight to *landed_weight, then zeroes the device accumulators. Legal between sessions. */
int hb_readback_xyz(HbEngine* h, float* xyz_wh3, float* landed_weight);
int hb_readback_xyz_render(HbEngine* h, uint32_t render, float* xyz_wh3, float* landed_weight);
/* TraceBackend::ReadbackClassLanes (trace_backend.hpp:471-493): copies the class_cnt * W*H per-class Y lanes of
 * render 0 (layout lane[c * W*H + py*W + px]) and zeroes them. *class_count receives the class count
 * (0 when the scene has no colour classes; nothing is written then). */
int hb_readback_class_lanes(HbEngine* h, float* lanes, uint64_t cap_floats, uint32_t* class_count);

/* Display sink on the device = RenderConsumer::PrepareSnapshot + PostSnapshot
 * (server/render.cpp:463-495,508-577; util/color_space.cpp:10-52): NON-destructive; the accumulator keeps
 * integrating. Writes (each optional) the 8-bit sRGB frame [H][W][3], the fp32 XYZ snapshot [H][W][3] and the
 * snapshot intensity (total landed weight). Exposure: intensity_factor * 0.08 * W*H / intensity
 * (ExposureScale, render.cpp:96-102). ray_color[0] < 0 selects real colour (gamut clip towards D65 grey);
 * otherwise luminance tinted by ray_color. */
typedef struct HbSnapshotDesc {
  float intensity_factor;   /* RenderConfig::intensity_factor_ (2^EV), default 1 */
  float ray_color[3];       /* RenderConfig::ray_color_, default {-1,-1,-1} = real colour */
  float background[3];      /* RenderConfig::background_, default 0 */
} HbSnapshotDesc;
int hb_snapshot(HbEngine* h, uint32_t render, const HbSnapshotDesc* desc, uint8_t* rgb8_wh3, float* xyz_wh3,
                float* intensity);
/* TraceBackend::DrainExits (trace_backend.hpp:443): destructive, grow-not-clamp. Call with
 * out == NULL to query the count. root_ids (optional) receives the layer-root index of each
 * exit, for per-ray parity association. */
int hb_drain_exits(HbEngine* h, HbExitRecord* out, uint32_t* root_ids, uint64_t cap, uint64_t* count);

/* Parity helpers (HostRayBatch injection, trace_backend.hpp:230-239; cpu_trace_backend.cpp:121-144).
 * hb_inject_rays: trace caller-supplied CRYSTAL-LOCAL rays (population 0, shape 0 of layer 0)
 * instead of generating roots; rot9 (row-major, optional, identity if NULL) is the per-ray
 * crystal->world rotation. Must be called after hb_begin_session and replaces the next
 * hb_trace_layer's generation step. */
int hb_inject_rays(HbEngine* h, uint64_t n, const float* d3, const float* p3, const float* w,
                   const uint16_t* to_face, const float* rot9);
/* hb_export_roots: after a root hb_trace_layer with spec.record_exits, copy out the roots the
 * engine generated (crystal-local d/p, weight, entry face, rot9, shape index, wl index). */
int hb_export_roots(HbEngine* h, uint64_t cap, float* d3, float* p3, float* w, uint16_t* to_face,
                    float* rot9, uint32_t* shape_idx, uint32_t* wl_idx, uint64_t* count);

/* Raypath colour, parity only: component masks the exported roots carry in from earlier layers (all zero on
 * layer 0); same order as hb_export_roots. Empty when the scene has no colour classes. */
int hb_export_root_masks(HbEngine* h, uint64_t cap, uint64_t* masks, uint64_t* count);

/* Tuning + measurement. Keys (value):
 *   tile_rays (1024 .. 2^30)   rays per device tile, default 2^24
 *   fold_rays (>= 1024)        root rays between folds of the fp32 working image into the fp64 master, default 2^21
 *   fused_bounce (0 | 1)       1 (default): one bounce kernel per interaction; 0: split optics + intersect kernels
 *   fused_gen (0 | 1)          1: root generation fused with the entry interaction; default 0 (measured slower)
 *   filter_hit_bound (0 | 1)   1 (default): end a layer's hit loop at the longest path its filters can admit
 *   prism_fast_path (0 | 1)    1 (default): unrolled canonical-prism forms; 0: generic paired-axis loop everywhere
 *   pixel_cache (0 | 1)        1 (default): per-CTA shared-memory pixel cache in front of the image reduction
 *   profile (0 | 1)            CUDA events around every launch (HbCounters *_ms)
 *   blocks_per_sm (1 .. 32)    override the occupancy-sized persistent grids (experiments)
 *   gen_base / stream_base     restart the monotone stream counters (tests: reproducible replays)
 *   fork_cap, cont_cap         cap the fork-ray slots / continuation pool (0 = automatic; tests force overflows) */
int hb_set_option(HbEngine* h, const char* key, int64_t value);
int hb_get_counters(HbEngine* h, HbCounters* out);
int hb_synchronize(HbEngine* h);
/* Device self-test of the engine's unchecked exact division / square root (csrc/hb_device.cuh: dvd_nr, sqrt_nr)
 * against the IEEE intrinsics on n counter-based random operand pairs. mode 0: a / b, b > 0, numerators incl. +-0;
 * mode 1: either sign of b; mode 2: sqrt. out4[0] = number of bitwise mismatches (must be 0), out4[1..3] = operand
 * bits and result of the last mismatch. */
int hb_selftest_arith(HbEngine* h, uint32_t mode, uint64_t n, uint32_t seed, uint64_t* out4);
/* Device pointer of the DOUBLE master accumulator (x,y,z,landed per pixel; all renders back to back; the
 * fp32 working image is folded into it first) for collectives driven from outside (torch.distributed);
 * valid until hb_set_render(s) / hb_destroy. */
int hb_image_device_ptr(HbEngine* h, void** ptr, uint64_t* float_count);
void* hb_stream(HbEngine* h);             /* cudaStream_t the engine launches on */

/* Multi-GPU frame end: in-place NCCL sum all-reduce of the accumulator. */
int hb_comm_init(HbEngine* h, const void* nccl_unique_id_128B, int rank, int nranks);
int hb_comm_unique_id(void* out_128B);
int hb_allreduce_image(HbEngine* h);
/* Frame end, preferred form: ncclReduce of fp32 (X, Y, Z, landed) per pixel (+ the colour-class lanes) to rank
 * `root`; afterwards the root's accumulator holds the sum of all ranks and every other rank's accumulator is
 * zero (its contribution moved), so only the root reads back and a repeated call changes nothing.
 * hb_allreduce_image leaves the sum on EVERY rank and therefore refuses (HB_ERR_STATE) to run twice on the same
 * accumulation. */
int hb_reduce_image(HbEngine* h, int root);
/* Several engines in ONE process (one host thread driving R devices behind the seam): add `src`'s
 * accumulators (all renders + colour lanes) into `dst`'s and zero `src`'s. Different devices: the kernel runs on
 * dst's device and reads src's fp64 master through peer access (NVLink P2P loads); asynchronous, ordered by events
 * on the two engines' streams. Both engines must hold the same renders; legal wherever hb_readback_xyz is
 * (between sessions or inside one, after the layers traced so far). */
int hb_merge_from_peer(HbEngine* dst, HbEngine* src);

/* ---------------------------------------------------------------------------
 * Host-side table builders (no GPU needed). The reference adapter uses the
 * reference's own builders instead; these exist so the library is usable
 * stand-alone and are parity-tested against the reference's tables.
 * ------------------------------------------------------------------------- */
/* Crystal::CreatePrism (crystal.cpp:349-356, geo3d_closedform.cpp ComputeClosedFormPrism). */
int hb_make_prism(float h, const float dist6[6], HbCrystalTables* out);
/* Crystal::CreatePyramid wedge-angle form (crystal.cpp:380-384). */
int hb_make_pyramid(float upper_alpha_deg, float lower_alpha_deg, float h1, float h2, float h3,
                    const float dist6[6], HbCrystalTables* out);
/* BuildGenGpParams + GetSharedLatLut (cuda_trace_backend.cu:336-400, lat_lut.cpp:74-180).
 * type/center/spread triples are in DEGREES as in the JSON config (zenith = 90 - latitude is
 * resolved by the caller: pass LATITUDE center). */
int hb_make_axis_sampler(uint32_t lat_type, float lat_center_deg, float lat_spread_deg,
                         uint32_t az_type, float az_center_deg, float az_spread_deg,
                         uint32_t roll_type, float roll_center_deg, float roll_spread_deg,
                         HbAxisSampler* out);
/* IceRefractiveIndex::Get (optics.cpp:180-198). */
double hb_ice_refractive_index(double wavelength_nm);
/* MakeCameraRotation + BuildProjParams (scatter_accum.hpp:18-26, lens_proj_build.hpp:24-140). */
int hb_make_proj_params(int lens_type, float fov_deg, int img_w, int img_h, float view_az_deg,
                        float view_el_deg, float view_ro_deg, int visible_range, int lens_shift_x,
                        int lens_shift_y, float overlap, HbProjParams* out);
/* PartitionCrystalRayNum (simulator.cpp:519-582). carry is in/out [cnt]. */
int hb_partition_rays(const float* proportions, uint32_t cnt, uint64_t ray_num, double* carry,
                      uint64_t* out_counts);

/* ---------------------------------------------------------------------------
 * Config-level scene description (mirrors SceneConfig / RenderConfig,
 * reference: src/config/proj_config.hpp:14-38, crystal_config.hpp, filter_config.hpp,
 * render_config.hpp:71-94) and the host builder that turns it into tables:
 * MakeCrystal per population (simulator.cpp:405-450, stochastic shapes drawn from a host
 * mt19937 into a geometry pool), BuildEntrySubTris, GetSharedLatLut, BuildDeviceFilterDesc.
 * ------------------------------------------------------------------------- */
typedef struct HbDist { uint32_t type; float center, spread; } HbDist;  /* Distribution, math.hpp */

typedef struct HbCrystalDesc {
  uint32_t kind;                 /* 0 prism, 1 pyramid */
  uint32_t id;                   /* CrystalConfig::id_ */
  HbDist height[3];              /* prism: [0] = h; pyramid: [0] upper h1, [1] prism h2, [2] lower h3 */
  HbDist face_dist[6];
  float wedge_upper_deg, wedge_lower_deg;
  HbDist latitude, azimuth, roll; /* AxisDistribution (latitude center = 90 - zenith), degrees */
  /* Shape-scalar sync groups (crystal_config.hpp:29-43,62-76; SyncGroupSampler, simulator.cpp:341-393), in
   * ShapeScalar order: [0] prism height, [1] upper_h, [2] prism_h, [3] lower_h, [4..9] face distances.
   * 0 = independent; scalars sharing a non-zero id share ONE raw draw per crystal instance (taken with the
   * distribution of the group's first member; heights fold it with |.| at their own use site). */
  int32_t sync_group[10];
} HbCrystalDesc;

typedef struct HbSimpleFilterSpec {
  uint32_t kind;                 /* 0 none, 1 raypath, 2 entry_exit, 3 direction, 4 crystal */
  uint32_t path_len;
  uint8_t path[HB_MAX_FILTER_PATH]; /* raypath: face numbers as written in the config */
  int32_t entry_fn, exit_fn;     /* -1 wildcard */
  uint32_t min_len, max_len;     /* max_len 0 = unbounded */
  float lon_deg, lat_deg, radii_deg;
  uint32_t crystal_id;
} HbSimpleFilterSpec;

typedef struct HbFilterSpecDesc {  /* FilterConfig, filter_config.hpp:70-84 */
  uint32_t kind;                 /* as HbSimpleFilterSpec.kind; 5 = complex: OR over terms of AND-ed simple filters */
  uint32_t action;               /* 0 filter_in, 1 filter_out */
  uint32_t symmetry;             /* bit0 P, bit1 B, bit2 D */
  HbSimpleFilterSpec simple;     /* kind 1..4 */
  uint32_t term_cnt;             /* kind 5: <= HB_MAX_FILTER_TERMS OR-terms */
  uint32_t term_len[HB_MAX_FILTER_TERMS]; /* AND-factors per term, <= 4 */
  HbSimpleFilterSpec terms[HB_MAX_FILTER_TERMS][4];
} HbFilterSpecDesc;

/* One colour predicate of a population (reference: ColorGateEntry, color_gate_table.hpp:25-48). */
typedef struct HbColorPredDesc {
  HbSimpleFilterSpec pred;       /* kind 0 = whole crystal */
  uint32_t symmetry;             /* P=1 | B=2 | D=4 */
  uint32_t bit;                  /* component bit, >= 64 = none */
} HbColorPredDesc;

typedef struct HbPopulationDesc {
  HbCrystalDesc crystal;
  HbFilterSpecDesc filter;
  float proportion;
  uint32_t color_pred_cnt;       /* grouped by symmetry in first-occurrence order (GroupPlacementBySymmetry) */
  HbColorPredDesc color_preds[HB_MAX_COLOR_PREDS];
} HbPopulationDesc;

typedef struct HbLayerDesc {
  float prob;
  uint32_t population_cnt;
  HbPopulationDesc populations[HB_MAX_CRYSTALS];
} HbLayerDesc;

typedef struct HbSceneDesc {
  uint32_t max_hits;
  uint32_t layer_cnt;
  float sun_altitude_deg, sun_azimuth_deg, sun_diameter_deg;
  uint32_t geom_pool_size;       /* shapes drawn per stochastic population (K-shape pool); 0 => 1 */
  HbLayerDesc layers[HB_MAX_LAYERS];
  HbColorClasses color_classes;
} HbSceneDesc;

typedef struct HbRenderDesc {
  int32_t lens_type; float fov_deg;
  int32_t img_w, img_h;
  float view_az_deg, view_el_deg, view_ro_deg;
  int32_t visible_range, lens_shift_x, lens_shift_y;
  float overlap;
} HbRenderDesc;

/* ---------------------------------------------------------------------------
 * Stochastic geometry pool on the device (reference: MakeCrystal per 32-ray batch on the CPU,
 * simulator.cpp:405-450 + simulator.hpp:144-151, and the per-batch / per-K-rays geometry pool of the reference
 * CUDA backend, cuda_trace_backend.cu:3617-3627). hb_resample_shapes redraws ALL shapes of one population's
 * pool on the device: shape scalars (heights, face distances) from `crystal`'s distributions with the
 * counter-based RNG (stream: seed, draw_base + shape index), crystal tables built by the same code as
 * hb_make_prism / hb_make_pyramid (bit-identical for equal scalars), pool overwritten in place (its size was
 * fixed by hb_set_scene). Legal between sessions only. A shape the builder rejects (face with more than 12
 * corners) becomes an empty crystal whose rays are discarded; their number is returned in *rejected.
 * hb_export_shapes copies the current pool back (parity harness): tables and, when the pool was drawn on the
 * device, the ten scalars per shape {h1, h2, h3, d0..d5, builder status}.
 * ------------------------------------------------------------------------- */
int hb_resample_shapes(HbEngine* h, uint32_t layer, uint32_t population, const HbCrystalDesc* crystal,
                       uint32_t seed, uint32_t draw_base, uint32_t* rejected);
int hb_export_shapes(HbEngine* h, uint32_t layer, uint32_t population, uint32_t cap, HbCrystalTables* tables,
                     float* scalars10, uint32_t* count);
/* Geometry clock run by the engine: from now on every hb_begin_session gives the population a FRESH pool
 * (shape-stream indices draw_base, draw_base + pool size, ...), drawn and built on the device ONE SESSION AHEAD
 * into a second copy of the layer's tables, so no session waits for its crystals and the host never drains
 * the stream (only the first session waits for its own pool). crystal == NULL switches the clock off (the live
 * pool stays). Legal between sessions; hb_set_scene resets it. */
int hb_auto_resample(HbEngine* h, uint32_t layer, uint32_t population, const HbCrystalDesc* crystal, uint32_t seed,
                     uint32_t draw_base);
/* (sqrt3/4) / tan(alpha): slope parameter of a pyramidal segment, -1 when alpha is outside [0.1, 89.9] deg
 * (geo3d_closedform.cpp ComputeClosedFormPyramidInner). */
double hb_pyramid_slope(float alpha_deg);

typedef struct HbSceneTables HbSceneTables;  /* owns an HbScene and all its storage */
int hb_build_scene(const HbSceneDesc* desc, uint32_t geometry_seed, HbSceneTables** out);
const HbScene* hb_scene_tables_get(const HbSceneTables* t);
void hb_free_scene(HbSceneTables* t);
int hb_build_render(const HbRenderDesc* desc, HbProjParams* out);
/* ComputeWlPool (backend/wl_pool.hpp:67-95) for a discrete wavelength: n from Sellmeier, CMF from
 * the CIE 1931 2-degree 1-nm table looked up at int(wl + 0.5). */
int hb_make_wl_entry(float wavelength_nm, float weight, HbWlEntry* out);
/* ComputeWlPool in illuminant mode (backend/wl_pool.hpp:73-84): M mid-point wavelengths across
 * [380, 780] nm, spd_weight = GetIlluminantSpd (util/illuminant.cpp:113-135; D-series rebuilt from the CIE
 * S0/S1/S2 daylight basis, A = Planck 2856 K normalised at 560 nm, E = 1). One session traced with this pool
 * draws a per-ray wavelength index (pcg_shared.h wl stream), as the reference's device backends do.
 * `illuminant`: HB_ILLUMINANT_*; 1 <= m <= 255. */
#define HB_ILLUMINANT_D50 0
#define HB_ILLUMINANT_D55 1
#define HB_ILLUMINANT_D65 2
#define HB_ILLUMINANT_D75 3
#define HB_ILLUMINANT_A 4
#define HB_ILLUMINANT_E 5
int hb_make_wl_pool_illuminant(int illuminant, uint32_t m, HbWlEntry* out);
float hb_illuminant_spd(int illuminant, float wavelength_nm);

#ifdef __cplusplus
}
#endif
#endif /* HALOTRACE_B200_H_ */
