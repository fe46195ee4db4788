/*
 * halo_oracle.h — CPU restatement of the reference's per-ray trace path.
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg — never by the product library, which has no CPU path at all.
 *
 * Pinning: tests/test_oracle_vs_reference.py checks every function here against the UNMODIFIED
 * reference CPU core (oracle/_ref, built from /root/reference by oracle/Makefile) when that library
 * is present, and tests/test_oracle_golden.py checks it against the committed fixtures in
 * tests/golden/ (generated from the same reference library by oracle/make_golden.py), which also
 * hold the reference's own known-answer values (test_optics.cpp, test_cpu_golden_rays.cpp).
 */
#ifndef HALO_ORACLE_H_
#define HALO_ORACLE_H_

#include <stdint.h>

#include "halotrace_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* One scattering layer as the engine sees it: shapes flattened over populations. */
typedef struct OrcLayerParams {
  const HbCrystalTables* shapes;        /* [shape_cnt], population 0's pool first */
  const uint32_t* shape_pop;            /* [shape_cnt] population index of each shape */
  uint32_t shape_cnt;
  const HbCrystalPopulation* pops;      /* [pop_cnt] (filter + crystal_id) */
  uint32_t pop_cnt;
  const HbWlEntry* wl;                  /* [wl_cnt] */
  uint32_t wl_cnt;
  uint32_t max_hits;
  float prob;
  uint32_t layer_idx;
  uint32_t seed;                        /* session seed (gate stream = seed ^ gate nonce) */
  uint64_t gate_base;                   /* global index of layer-ray 0 in the gate stream */
  /* raypath colour (simulator.cpp:688-712): pops[].color_groups are evaluated on filter-admitted exits */
  const uint64_t* root_mask;            /* [n] component mask each root carries in (NULL = 0) */
  uint64_t* cont_mask;                  /* [cap] out: mask of each continuation (NULL = not wanted) */
} OrcLayerParams;

uint32_t orc_pcg_hash(uint32_t x);
float orc_draw(uint32_t seed, uint32_t idx, uint32_t slot);
uint32_t orc_feistel(uint32_t i, uint32_t n, uint32_t seed);
/* the generator's building blocks (twins of lm_pcg::*, pcg_shared.h:193-624) */
uint32_t orc_seed_with_high(uint32_t seed, uint32_t hi);
void orc_uniforms(uint32_t seed, uint32_t idx, uint32_t slot0, uint32_t n, float* out);
void orc_get_dist(uint32_t seed, uint32_t idx0, uint32_t n, uint32_t type, float mean, float stdv, float* out);
void orc_lat_lon_roll(const HbAxisSampler* a, uint32_t seed, uint32_t idx0, uint32_t n, float* lon_lat_roll3,
                      uint32_t* slots_used);
void orc_rotation9(uint64_t n, const float* lon_lat_roll3, float* rot9);
void orc_sph_cap(uint32_t seed, uint32_t idx0, uint32_t slot0, uint32_t n, float lon, float lat, float half, float* d3);
void orc_triangle(uint32_t seed, uint32_t idx0, uint32_t slot0, uint32_t n, const float* vtx9, float* p3);

/* Root generation (engine stream spec, DESIGN.md "RNG streams"; samplers restate pcg_shared.h). */
int orc_gen_roots(const HbScene* scene, uint32_t layer, uint32_t pop, uint32_t shape_base, const HbWlEntry* wl,
                  uint32_t wl_cnt, uint32_t seed, uint64_t ray_base, uint64_t n, float* d3, float* p3, float* w,
                  uint16_t* face, float* quat4, float* rot9, uint32_t* shape_idx, uint32_t* wl_idx);
/* Layer transit (InitRayOtherMs, simulator.cpp:309-339): world dir in -> new crystal-local root. */
int orc_transit(const HbScene* scene, uint32_t layer, uint32_t pop, uint32_t shape_base, uint32_t seed,
                uint64_t ray_base, uint64_t n, const float* d_world3, float* d3, float* p3, uint16_t* face,
                float* quat4, float* rot9, uint32_t* shape_idx);

/* Hit loop + emit gate (simulator.cpp:585-762, optics.cpp:18-177). Rays are crystal-local. */
int orc_trace_layer(const OrcLayerParams* lp, uint64_t n, const float* d3, const float* p3, const float* w,
                    const uint16_t* face, const float* rot9, const uint32_t* shape_idx, const uint32_t* wl_idx,
                    uint64_t cap, HbExitRecord* exits, uint32_t* exit_root, uint64_t* exit_cnt, float* cont_d3,
                    float* cont_w, uint32_t* cont_wl, uint32_t* cont_root, uint64_t* cont_cnt);

/* Primitives, one call = n independent evaluations. */
int orc_hit_surface(const HbCrystalTables* t, float n_idx, uint64_t n, const float* d3, const float* w,
                    const uint16_t* face, float* d_out6, float* w_out2);
int orc_propagate(const HbCrystalTables* t, uint64_t n, const float* d3, const float* p3, const float* w,
                  const uint16_t* from_face, float* p_out3, uint16_t* to_face);
int orc_project(const HbProjParams* p, uint64_t n, const float* dir3, int32_t* px2, int32_t* py2, int32_t* cnt,
                int32_t* bump2);
/* ScatterOutgoingToXyz (scatter_accum.hpp:47-110) with per-ray wavelength-pool CMF. */
int orc_accumulate(const HbProjParams* p, const HbWlEntry* wl, uint32_t wl_cnt, uint64_t n, const float* dir3,
                   const float* w, const uint8_t* wl_idx, float* xyz_wh3, double* landed);
int orc_filter_check(const HbFilterDesc* f, const uint8_t* face_fn, uint32_t pop_crystal_id, uint64_t n,
                     const uint8_t* paths64, const uint8_t* path_len, const float* dir3, uint8_t* pass);
void orc_quat_to_rot9(const float* q4, float* rot9);
/* FanColorClassLanes (cuda_trace_backend.cu:538-556) over a list of exits: lane[c*W*H + pix] += cmf_y * w for
 * every class c the exit's mask satisfies; sequential fp32 adds in list order. */
int orc_accumulate_lanes(const HbProjParams* p, const HbWlEntry* wl, uint32_t wl_cnt, const HbColorClasses* classes,
                         uint64_t n, const float* dir3, const float* w, const uint8_t* wl_idx, const uint64_t* mask,
                         float* lanes);
/* Display sink: RenderConsumer::PostSnapshot (server/render.cpp:508-577) on a snapshot XYZ image with
 * ExposureScale (render.cpp:96-102), GamutClipXyz / XyzToLinearRgb / LinearToSrgb (util/color_space.cpp:10-52). */
int orc_post_snapshot(const float* xyz_wh3, int w, int h, float snapshot_intensity, float intensity_factor,
                      const float* ray_color3, const float* background3, uint8_t* rgb8_wh3);

#ifdef __cplusplus
}
#endif
#endif
