// hb_launch.cuh — host-side launch entry points of the trace kernels. The kernel templates (hb_kernels.cuh) are
// instantiated in several translation units (hb_tu.cu compiled with -DHB_TU=n, see the Makefile) so the library
// builds in parallel; hb_engine.cu only sees these declarations.
#ifndef HB_LAUNCH_CUH_
#define HB_LAUNCH_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "hb_kernels.cuh"

namespace hb {

struct LaunchCtx {
  int device;
  int sm_count;
  int blocks_per_sm_override;  // option "blocks_per_sm" (experiments); 0 = occupancy
  cudaStream_t stream;
};

// key = general << 2 | last << 1 | tables_in_smem; p4 = 0 generic axis loop, 1 all hexagonal prisms, 2 per ray.
// Sessions with extra renders or raypath colour run the MULTI instantiations.
void launch_optics(const LaunchCtx& c, bool general, bool last, bool in_smem, int p4, size_t smem, const TraceParams& tp);
void launch_intersect(const LaunchCtx& c, bool general, bool in_smem, int p4, size_t smem, const TraceParams& tp);
void launch_bounce(const LaunchCtx& c, bool general, bool last, bool in_smem, int p4, size_t smem, const TraceParams& tp);

// Root generation fused with the entry interaction (genbounce_kernel); the caller continues the hit loop at hit 1.
void launch_genbounce(const LaunchCtx& c, bool transit, bool general, bool in_smem, int p4, size_t smem, const GenParams& gp,
                      const TraceParams& tp);

// per-TU pieces
void launch_genbounce_p0(const LaunchCtx& c, int key, size_t smem, const GenParams& gp, const TraceParams& tp);
void launch_genbounce_p1(const LaunchCtx& c, int key, size_t smem, const GenParams& gp, const TraceParams& tp);
void launch_genbounce_p2(const LaunchCtx& c, int key, size_t smem, const GenParams& gp, const TraceParams& tp);
void launch_genbounce_multi(const LaunchCtx& c, int key, size_t smem, const GenParams& gp, const TraceParams& tp);
void launch_optics_p0(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp);
void launch_optics_p1(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp);
void launch_optics_p2(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp);
void launch_optics_multi(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp);
void launch_intersect_p0(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp);
void launch_intersect_p1(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp);
void launch_intersect_p2(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp);
void launch_intersect_multi(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp);
void launch_bounce_p0(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp);
void launch_bounce_p1(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp);
void launch_bounce_p2(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp);
void launch_bounce_multi(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp);

// Grid-stride launch of kGridWaves x the co-resident CTAs (occupancy x SM count). A grid of exactly the resident
// CTAs leaves SMs idle while the slowest CTAs finish (exit-heavy ranges, the far die's L2 latency); three waves of
// smaller CTAs let the hardware scheduler even that out at a per-CTA prologue (table staging, pixel-cache init and
// flush) that stays below 1 % of a CTA's work. Measured on config 2 (CTAs per SM for the 4-resident bounce kernels):
// 4: 5.41, 8: 5.54, 12: 5.58, 16: 5.58, 24: 5.54, 32: 5.50, 64: 5.20 G rays/s.
constexpr int kGridWaves = 3;
template <typename K>
uint32_t resident_grid(const LaunchCtx& c, K kernel, size_t smem, uint64_t n) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, smem) != cudaSuccess || per_sm < 1) per_sm = 4;
  per_sm *= kGridWaves;
  if (c.blocks_per_sm_override > 0) per_sm = c.blocks_per_sm_override;
  const uint64_t blocks = (n + 255) / 256;
  const uint64_t cap = static_cast<uint64_t>(c.sm_count) * per_sm;
  return static_cast<uint32_t>(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace hb

#endif  // HB_LAUNCH_CUH_
