/*
 * ref_driver.h — C entry points of oracle/_ref/libhalo_ref*.so.
 *
 * TEST INFRASTRUCTURE ONLY. The library behind this header is the UNMODIFIED reference
 * (LoveDaisy/ice_halo_sim) CPU core compiled from /root/reference by oracle/Makefile plus the thin
 * driver in ref_driver.cpp. It exists to (1) pin our own oracle restatement (halo_oracle.cpp),
 * (2) generate the golden fixtures under tests/golden/ (oracle/make_golden.py), and (3) serve as the
 * `cpu_baseline.kind = "reference"` leg of bench.py. Product code never links or loads it.
 */
#ifndef HALO_REF_DRIVER_H_
#define HALO_REF_DRIVER_H_

#include <stdint.h>

#include "halotrace_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* A concrete crystal shape (already-sampled scalars). kind 0: prism(h1 = height), 1: pyramid. */
typedef struct RefShape {
  uint32_t kind;
  float upper_alpha_deg, lower_alpha_deg;
  float h1, h2, h3;
  float dist[6];
} RefShape;

/* --- tables built by the reference's own host code --- */
int ref_make_tables(const RefShape* shape, HbCrystalTables* out);          /* Crystal + BuildEntrySubTris */
int ref_make_axis_sampler(const HbDist* lat, const HbDist* az, const HbDist* roll, HbAxisSampler* out);
int ref_make_proj_params(const HbRenderDesc* r, HbProjParams* out);        /* MakeCameraRotation + BuildProjParams */
int ref_wl_entry(float wl, float weight, HbWlEntry* out);                  /* ComputeWlPool, discrete */
int ref_wl_pool_illuminant(int illuminant, uint32_t m, HbWlEntry* out);    /* ComputeWlPool, illuminant */
double ref_refractive_index(double wl);                                    /* IceRefractiveIndex::Get */
int ref_post_snapshot(const float* xyz_wh3, int w, int h, float snapshot_intensity, float intensity_factor,
                      const float* ray_color3, const float* background3, uint8_t* rgb8_wh3); /* PostSnapshot */
int ref_daylight_basis(float* s012_107x3);                                /* kDaylightS0/S1/S2, 300..830 nm */
int ref_cmf_table(float* xyz_471x3);                                       /* kCmfX/Y/Z, 360..830 nm */
int ref_filter_desc(const HbPopulationDesc* pop, const RefShape* shape, HbFilterDesc* out); /* BuildDeviceFilterDesc */

/* --- primitives (KAT scenarios of test_optics.cpp) --- */
int ref_hit_surface(const RefShape* shape, float n_idx, uint64_t n, const float* d3, const float* w,
                    const uint16_t* face, float* d_out6, float* w_out2);   /* HitSurface */
int ref_propagate(const RefShape* shape, uint64_t n, const float* d3, const float* p3, const float* w,
                  const uint16_t* from_face, float* p_out3, uint16_t* to_face);  /* Propagate, step=1 */
int ref_project(const HbRenderDesc* r, uint64_t n, const float* dir3, int32_t* px2, int32_t* py2,
                int32_t* cnt, int32_t* bump2);                             /* ProjectExitToPixel */
int ref_scatter_xyz(const HbRenderDesc* r, float wl, uint64_t n, const float* dir3, const float* w,
                    float* xyz_wh3, float* landed);                        /* ScatterOutgoingToXyz */
int ref_filter_check(const HbPopulationDesc* pop, const RefShape* shape, uint64_t n, const uint8_t* paths64,
                     const uint8_t* path_len, const float* dir3, uint8_t* pass);  /* FilterSpec::Check */
int ref_sample_orientations(const HbDist* lat, const HbDist* az, const HbDist* roll, uint32_t seed,
                            uint64_t n, float* lon_lat_roll3, float* rot9);       /* InitRay_rot path */
int ref_partition(const float* proportions, uint32_t cnt, uint64_t ray_num, double* carry, uint64_t* out);

/* --- CpuTraceBackend with HostRayBatch injection, ONE ray per session (SURVEY 8(c) protocol) ---
 * rays are crystal-local; out_ray[k] is the injected-ray index of exit k. */
int ref_trace_injected(const RefShape* shape, float n_idx, uint32_t max_hits, uint64_t n, const float* d3,
                       const float* p3, const float* w, const uint16_t* to_face, uint64_t cap,
                       HbExitRecord* out, uint32_t* out_ray, uint64_t* count);
/* Same with raypath-colour predicates (color_pop->color_preds, one colour class per predicate => bit k for
 * predicate k): exits carry ExitRayRecord::component_mask as CollectData produces it (simulator.cpp:688-712). */
int ref_trace_injected_color(const RefShape* shape, float n_idx, uint32_t max_hits, uint64_t n, const float* d3,
                             const float* p3, const float* w, const uint16_t* to_face,
                             const HbPopulationDesc* color_pop, uint64_t cap, HbExitRecord* out, uint32_t* out_ray,
                             uint64_t* count);

/* --- CpuTraceBackend, self-generated rays (mt19937), whole scene; image via ReadbackImage --- */
int ref_cpu_backend_run(const HbSceneDesc* scene, const HbRenderDesc* render, float wl, float weight,
                        uint32_t seed, uint64_t n_rays, uint64_t session_rays, float* xyz_wh3, float* landed,
                        uint64_t* exit_count, double* exit_w_sum);

/* --- legacy multi-threaded CPU path (Simulator::Run + host projection), the CPU baseline ---
 * returns root rays/s over the timed window (setup excluded). */
int ref_legacy_bench(const HbSceneDesc* scene, const HbRenderDesc* render, const float* wl, const float* wl_weight,
                     uint32_t wl_cnt, uint64_t rays_per_wl, uint32_t threads, uint32_t dispatch_rays,
                     double* rays_per_sec, double* seconds, uint64_t* exits);
uint32_t ref_physical_cores(void);

/* --- the reference's counter-based sampler, core/shared/pcg_shared.h (lm_pcg::*), called as is --- */
uint32_t ref_pcg_hash(uint32_t x);
uint32_t ref_pcg_seed_with_high(uint32_t seed, uint32_t hi);
void ref_pcg_uniforms(uint32_t seed, uint32_t idx, uint32_t slot0, uint32_t n, float* out);
void ref_pcg_get_dist(uint32_t seed, uint32_t idx0, uint32_t n, uint32_t type, float mean, float stdv, float* out);
void ref_pcg_lat_lon_roll(const HbAxisSampler* a, uint32_t seed, uint32_t idx0, uint32_t n, float* lon_lat_roll3,
                          uint32_t* slots_used);
void ref_pcg_rotation9(uint64_t n, const float* lon_lat_roll3, float* rot9);    /* build_crystal_rotation_9 */
void ref_pcg_sph_cap(uint32_t seed, uint32_t idx0, uint32_t slot0, uint32_t n, float lon, float lat, float half, float* d3);
void ref_pcg_triangle(uint32_t seed, uint32_t idx0, uint32_t slot0, uint32_t n, const float* vtx9, float* p3);
void ref_pcg_feistel(uint32_t n, uint32_t seed, uint32_t* out);                /* feistel_bijection(i, n, seed), i < n */
void ref_pcg_categorical(const float* weights, uint32_t n, const float* u, uint32_t m, uint32_t* out);

/* --- the reference driver's TraceBackend route: unmodified Simulator::Run (one thread) + third-clock drain; the
 * backend is fixed by the build (see ref_driver.cpp): legacy CPU fallback / the reference CudaTraceBackend /
 * this repo's B200TraceBackend through oracle/shim. xyz_wh3 receives the sum of all drained images. --- */
int ref_backend_bench(const HbSceneDesc* scene, const HbRenderDesc* render, const float* wl, const float* wl_weight,
                      uint32_t wl_cnt, uint64_t rays_per_wl, uint64_t dispatch_rays, uint32_t seed, float* xyz_wh3,
                      double* landed, double* rays_per_sec, double* seconds, uint32_t* backend_used);

#ifdef __cplusplus
}
#endif
#endif
