// hb_tu.cu — kernel instantiations, one slice per translation unit (-DHB_TU=n; Makefile). Every slice owns the
// launchers of one kernel family x one crystal-form variant, so the slices compile in parallel.
//   0..3  : split pipeline (optics + intersect), P4 = 0 / 1 / 2 / MULTI
//   4..7  : fused bounce,                        P4 = 0 / 1 / 2 / MULTI
//   8     : dispatch (no kernels)
//   9..12 : root generation fused with the entry interaction, P4 = 0 / 1 / 2 / MULTI
#include <string>

namespace hb {
std::string& global_error();
}
#include "hb_launch.cuh"

namespace hb {
namespace {

// Function attributes live in the device's context: set once per device and instantiation.
template <bool G, bool L, bool S, bool M, int P>
void launch_optics_t(const LaunchCtx& c, size_t smem, const TraceParams& tp) {
  static bool attr_set[64] = {};
  if (!attr_set[c.device & 63]) {
    cudaFuncSetAttribute(optics_kernel<G, L, S, M, P>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         static_cast<int>(shared_tables_bytes(kSmemShapes) + kCacheBytes + kStageBytes));
    attr_set[c.device & 63] = true;
  }
  const uint32_t grid = resident_grid(c, optics_kernel<G, L, S, M, P>, smem, tp.cap);
  optics_kernel<G, L, S, M, P><<<grid, 256, smem, c.stream>>>(tp);
}
template <bool G, bool S, bool M, int P>
void launch_intersect_t(const LaunchCtx& c, size_t smem, const TraceParams& tp) {
  const uint32_t grid = resident_grid(c, intersect_kernel<G, S, M, P>, smem, tp.cap);
  intersect_kernel<G, S, M, P><<<grid, 256, smem, c.stream>>>(tp);
}
template <bool G, bool L, bool S, bool M, int P>
void launch_bounce_t(const LaunchCtx& c, size_t smem, const TraceParams& tp) {
  static bool attr_set[64] = {};
  if (!attr_set[c.device & 63]) {
    cudaFuncSetAttribute(bounce_kernel<G, L, S, M, P>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         static_cast<int>(shared_tables_bytes(kSmemShapes) + kCacheBytes + kStage2Bytes + kQueueBytes + kQueue2Bytes));
    attr_set[c.device & 63] = true;
  }
  const uint32_t grid = resident_grid(c, bounce_kernel<G, L, S, M, P>, smem, tp.cap);
  bounce_kernel<G, L, S, M, P><<<grid, 256, smem, c.stream>>>(tp);
}

template <bool T, bool G, bool S, bool M, int P>
void launch_genbounce_t(const LaunchCtx& c, size_t smem, const GenParams& gp, const TraceParams& tp) {
  static bool attr_set[64] = {};
  if (!attr_set[c.device & 63]) {
    cudaFuncSetAttribute(genbounce_kernel<T, G, S, M, P>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         static_cast<int>(kGenSharedBytes + shared_tables_bytes(kSmemShapes) + kCacheBytes + kQueueBytes + kQueue2Bytes));
    attr_set[c.device & 63] = true;
  }
  const uint32_t grid = resident_grid(c, genbounce_kernel<T, G, S, M, P>, smem, gp.count);
  genbounce_kernel<T, G, S, M, P><<<grid, 256, smem, c.stream>>>(gp, tp);
}
// key = transit << 2 | general << 1 | tables_in_smem
template <bool M, int P>
void genbounce_by_key(const LaunchCtx& c, int key, size_t smem, const GenParams& gp, const TraceParams& tp) {
  if (M) key |= 2;
  switch (key) {
    case 0: if constexpr (!M) launch_genbounce_t<false, false, false, false, P>(c, smem, gp, tp); break;
    case 1: if constexpr (!M) launch_genbounce_t<false, false, true, false, P>(c, smem, gp, tp); break;
    case 2: launch_genbounce_t<false, true, false, M, P>(c, smem, gp, tp); break;
    case 3: launch_genbounce_t<false, true, true, M, P>(c, smem, gp, tp); break;
    case 4: if constexpr (!M) launch_genbounce_t<true, false, false, false, P>(c, smem, gp, tp); break;
    case 5: if constexpr (!M) launch_genbounce_t<true, false, true, false, P>(c, smem, gp, tp); break;
    case 6: launch_genbounce_t<true, true, false, M, P>(c, smem, gp, tp); break;
    default: launch_genbounce_t<true, true, true, M, P>(c, smem, gp, tp); break;
  }
}

template <bool M, int P>
void optics_by_key(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp) {
  if (M) key |= 4;  // MULTI kernels are GENERAL
  switch (key) {
    case 0: if constexpr (!M) launch_optics_t<false, false, false, false, P>(c, smem, tp); break;
    case 1: if constexpr (!M) launch_optics_t<false, false, true, false, P>(c, smem, tp); break;
    case 2: if constexpr (!M) launch_optics_t<false, true, false, false, P>(c, smem, tp); break;
    case 3: if constexpr (!M) launch_optics_t<false, true, true, false, P>(c, smem, tp); break;
    case 4: launch_optics_t<true, false, false, M, P>(c, smem, tp); break;
    case 5: launch_optics_t<true, false, true, M, P>(c, smem, tp); break;
    case 6: launch_optics_t<true, true, false, M, P>(c, smem, tp); break;
    default: launch_optics_t<true, true, true, M, P>(c, smem, tp); break;
  }
}
template <bool M, int P>
void intersect_by_key(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp) {
  if (M) key |= 4;
  switch (key & 5) {
    case 0: if constexpr (!M) launch_intersect_t<false, false, false, P>(c, smem, tp); break;
    case 1: if constexpr (!M) launch_intersect_t<false, true, false, P>(c, smem, tp); break;
    case 4: launch_intersect_t<true, false, M, P>(c, smem, tp); break;
    default: launch_intersect_t<true, true, M, P>(c, smem, tp); break;
  }
}
template <bool M, int P>
void bounce_by_key(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp) {
  if (M) key |= 4;
  switch (key) {
    case 0: if constexpr (!M) launch_bounce_t<false, false, false, false, P>(c, smem, tp); break;
    case 1: if constexpr (!M) launch_bounce_t<false, false, true, false, P>(c, smem, tp); break;
    case 2: if constexpr (!M) launch_bounce_t<false, true, false, false, P>(c, smem, tp); break;
    case 3: if constexpr (!M) launch_bounce_t<false, true, true, false, P>(c, smem, tp); break;
    case 4: launch_bounce_t<true, false, false, M, P>(c, smem, tp); break;
    case 5: launch_bounce_t<true, false, true, M, P>(c, smem, tp); break;
    case 6: launch_bounce_t<true, true, false, M, P>(c, smem, tp); break;
    default: launch_bounce_t<true, true, true, M, P>(c, smem, tp); break;
  }
}

}  // namespace

#if HB_TU == 0
void launch_optics_p0(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp) { optics_by_key<false, 0>(c, key, smem, tp); }
void launch_intersect_p0(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp) { intersect_by_key<false, 0>(c, key, smem, tp); }
#elif HB_TU == 1
void launch_optics_p1(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp) { optics_by_key<false, 1>(c, key, smem, tp); }
void launch_intersect_p1(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp) { intersect_by_key<false, 1>(c, key, smem, tp); }
#elif HB_TU == 2
void launch_optics_p2(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp) { optics_by_key<false, 2>(c, key, smem, tp); }
void launch_intersect_p2(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp) { intersect_by_key<false, 2>(c, key, smem, tp); }
#elif HB_TU == 3
void launch_optics_multi(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp) { optics_by_key<true, 2>(c, key, smem, tp); }
void launch_intersect_multi(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp) { intersect_by_key<true, 2>(c, key, smem, tp); }
#elif HB_TU == 4
void launch_bounce_p0(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp) { bounce_by_key<false, 0>(c, key, smem, tp); }
#elif HB_TU == 5
void launch_bounce_p1(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp) { bounce_by_key<false, 1>(c, key, smem, tp); }
#elif HB_TU == 6
void launch_bounce_p2(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp) { bounce_by_key<false, 2>(c, key, smem, tp); }
#elif HB_TU == 7
void launch_bounce_multi(const LaunchCtx& c, int key, size_t smem, const TraceParams& tp) { bounce_by_key<true, 2>(c, key, smem, tp); }
#elif HB_TU == 9
void launch_genbounce_p0(const LaunchCtx& c, int key, size_t smem, const GenParams& gp, const TraceParams& tp) { genbounce_by_key<false, 0>(c, key, smem, gp, tp); }
#elif HB_TU == 10
void launch_genbounce_p1(const LaunchCtx& c, int key, size_t smem, const GenParams& gp, const TraceParams& tp) { genbounce_by_key<false, 1>(c, key, smem, gp, tp); }
#elif HB_TU == 11
void launch_genbounce_p2(const LaunchCtx& c, int key, size_t smem, const GenParams& gp, const TraceParams& tp) { genbounce_by_key<false, 2>(c, key, smem, gp, tp); }
#elif HB_TU == 12
void launch_genbounce_multi(const LaunchCtx& c, int key, size_t smem, const GenParams& gp, const TraceParams& tp) { genbounce_by_key<true, 2>(c, key, smem, gp, tp); }
#elif HB_TU == 8
void launch_genbounce(const LaunchCtx& c, bool transit, bool general, bool in_smem, int p4, size_t smem, const GenParams& gp,
                      const TraceParams& tp) {
  const int key = (transit ? 4 : 0) | (general ? 2 : 0) | (in_smem ? 1 : 0);
  if (tp.extra_cnt != 0u || tp.color_on != 0u) launch_genbounce_multi(c, key, smem, gp, tp);
  else if (p4 == 1) launch_genbounce_p1(c, key, smem, gp, tp);
  else if (p4 == 2) launch_genbounce_p2(c, key, smem, gp, tp);
  else launch_genbounce_p0(c, key, smem, gp, tp);
}
void launch_optics(const LaunchCtx& c, bool general, bool last, bool in_smem, int p4, size_t smem, const TraceParams& tp) {
  const int key = (general ? 4 : 0) | (last ? 2 : 0) | (in_smem ? 1 : 0);
  if (tp.extra_cnt != 0u || tp.color_on != 0u) launch_optics_multi(c, key, smem, tp);
  else if (p4 == 1) launch_optics_p1(c, key, smem, tp);
  else if (p4 == 2) launch_optics_p2(c, key, smem, tp);
  else launch_optics_p0(c, key, smem, tp);
}
void launch_intersect(const LaunchCtx& c, bool general, bool in_smem, int p4, size_t smem, const TraceParams& tp) {
  const int key = (general ? 4 : 0) | (in_smem ? 1 : 0);
  if (tp.extra_cnt != 0u || tp.color_on != 0u) launch_intersect_multi(c, key, smem, tp);
  else if (p4 == 1) launch_intersect_p1(c, key, smem, tp);
  else if (p4 == 2) launch_intersect_p2(c, key, smem, tp);
  else launch_intersect_p0(c, key, smem, tp);
}
void launch_bounce(const LaunchCtx& c, bool general, bool last, bool in_smem, int p4, size_t smem, const TraceParams& tp) {
  const int key = (general ? 4 : 0) | (last ? 2 : 0) | (in_smem ? 1 : 0);
  if (tp.extra_cnt != 0u || tp.color_on != 0u) launch_bounce_multi(c, key, smem, tp);
  else if (p4 == 1) launch_bounce_p1(c, key, smem, tp);
  else if (p4 == 2) launch_bounce_p2(c, key, smem, tp);
  else launch_bounce_p0(c, key, smem, tp);
}
#endif

}  // namespace hb
