// hb_engine.cu — host side of the B200 wavefront trace engine (sm_100a): session state machine, tile loop,
// kernel launches and the compute half of the C ABI declared in include/halotrace_b200.h.
// The kernels themselves are in hb_kernels.cuh.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "halotrace_b200.h"
#include "hb_device.cuh"

#ifdef HB_WITH_NCCL
#include <nccl.h>
#endif

namespace hb {

std::string& global_error();  // hb_host.cpp

}  // namespace hb

#include "hb_kernels.cuh"
#include "hb_launch.cuh"

namespace hb {

// ------------------------------------------------------------------------------------------------
// Engine
// ------------------------------------------------------------------------------------------------
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t ensure(size_t want) {
    if (want <= n) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess) n = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

struct PopHost {
  float proportion;
  uint32_t crystal_id;
  uint32_t shape_base, shape_cnt;
  AxisParams axis;
  bool has_filter;
  bool device_pool;  // pool last drawn on the device (scalars available for export)
  // geometry clock (hb_auto_resample): the pool is redrawn on the device for every session, one session ahead
  bool auto_geom;
  HbCrystalDesc auto_desc;
  uint32_t auto_seed;
  uint32_t auto_draws;  // shape-stream index of the next pool
};

// Everything the kernels read of a layer's crystal shapes. Two copies exist once a population's pool is redrawn
// ahead of time: the trace reads sb[cur] while the next session's pool is built into sb[cur ^ 1].
struct ShapeBufs {
  DevBuf<float4> planes;
  DevBuf<float4> axes;
  DevBuf<uint32_t> shape_meta;
  DevBuf<uint8_t> face_fn;
  DevBuf<HbCrystalTables> shapes;
  DevBuf<EntryFaces> entry_faces;  // [shape] face groups of the entry fan table
  DevBuf<float> geom_scalars;      // [shape][10] scalars of device-drawn shapes (parity export)
  std::vector<uint32_t> pop_p4;    // [pop] number of shapes of the pool that are full hexagonal prisms (kMetaP4)
  std::vector<uint32_t> pop_shapes;  // [pop] pool size
  bool allocated = false;
  // kernel variant of the layer: 1 = every shape is P4 (unrolled forms only), 2 = some are (per-ray choice), 0 = none
  int p4_mode() const {
    uint32_t p4 = 0, all = 0;
    for (size_t i = 0; i < pop_p4.size(); i++) {
      p4 += pop_p4[i];
      all += pop_shapes[i];
    }
    return p4 == all ? 1 : (p4 != 0 ? 2 : 0);
  }
  void release() {
    planes.release();
    axes.release();
    shape_meta.release();
    face_fn.release();
    shapes.release();
    entry_faces.release();
    geom_scalars.release();
    allocated = false;
  }
};

struct LayerDev {
  float prob = 0.0f;
  std::vector<PopHost> pops;
  std::vector<double> carry;  // PartitionCrystalRayNum carry (simulator.cpp:519-582)
  uint32_t shape_cnt = 0;
  bool any_filter = false;
  uint32_t hit_bound = 0xFFFFFFFFu;  // longest path any population's filter can admit (filter_max_len)
  ShapeBufs sb[2];
  int cur = 0;
  ShapeBufs& b() { return sb[cur]; }
  // geometry prefetch state (see prefetch_geometry)
  bool any_auto = false;
  bool shadow_pending = false;
  cudaEvent_t shadow_event = nullptr;
  cudaEvent_t main_event = nullptr;  // "everything enqueued on the trace stream so far" marker for the geometry stream
  DevBuf<uint32_t> geom_flags_dev;   // [pop][2]: shapes that are not P4, shapes the builder rejected
  uint32_t* geom_flags_host = nullptr;  // pinned mirror
  DevBuf<HbFilterDesc> filters;
  DevBuf<uint32_t> pop_crystal_id;
  DevBuf<float> luts;  // [pop][3][257]
  bool any_color = false;              // some population carries colour predicates
  DevBuf<HbColorGroup> color_groups;   // [pop][HB_MAX_COLOR_GROUPS]
  DevBuf<uint32_t> color_group_cnt;    // [pop]
  void release() {
    sb[0].release();
    sb[1].release();
    geom_flags_dev.release();
    if (geom_flags_host) cudaFreeHost(geom_flags_host);
    geom_flags_host = nullptr;
    if (shadow_event) cudaEventDestroy(shadow_event);
    shadow_event = nullptr;
    if (main_event) cudaEventDestroy(main_event);
    main_event = nullptr;
    filters.release();
    pop_crystal_id.release();
    luts.release();
    color_groups.release();
    color_group_cnt.release();
  }
};

struct EventPair {
  cudaEvent_t a, b;
  int family;       // 0 gen, 1 optics, 2 intersect, 3 fused bounce, 4 gen + entry interaction
  uint64_t rays;
};

uint32_t pcg_hash_host(uint32_t x);
uint32_t bits_face_host(const float4& p);

}  // namespace hb

using namespace hb;  // NOLINT

struct HbEngine {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  std::string error;

  // scene / render
  uint32_t max_hits = 0;
  float sun_lon = 0, sun_lat = 0, sun_half = 0;
  std::vector<std::unique_ptr<LayerDev>> layers;
  bool have_scene = false, have_render = false;
  HbProjParams proj{};                 // render 0
  std::vector<HbProjParams> renders;   // all renders of the current trace (1..HB_MAX_RENDERS)
  std::vector<uint32_t> render_off;    // first pixel of each render in the arena
  uint32_t arena_pix = 0;              // pixels of all renders together
  DevBuf<ExtraRender> extra_dev;       // renders 1.. for the emit path
  HbColorClasses classes{};            // raypath colour classes of the scene (class_cnt 0 = off)
  DevBuf<float> lanes;                 // [class_cnt][pixels of render 0] per-class Y lanes
  uint64_t lanes_floats = 0;
  DevBuf<uint64_t> mask_tile;          // [cap] carried component mask per ray slot (layers >= 1)
  DevBuf<uint8_t> rgb_stage;           // 8-bit snapshot staging
  DevBuf<float4> image;
  DevBuf<double4> master;
  uint64_t rays_since_fold = 0;
  uint64_t fold_rays = 1u << 21;
  DevBuf<float> xyz_stage;
  DevBuf<double> landed_dev;
  DevBuf<float4> reduce_stage;         // fp32 (X, Y, Z, landed) of the arena for the NCCL reduce
  bool allreduced = false;             // the master holds an all-reduced sum (a second all-reduce would multiply it)
  cudaEvent_t merge_event = nullptr;   // hb_merge_from_peer hand-shake

  // session
  bool in_session = false;
  HbSessionSpec spec{};
  struct WlSlot {
    std::vector<HbWlEntry> host;
    DevBuf<HbWlEntry> dev;
    std::vector<WlDev> host2;  // derived table the trace kernels read (n, 1/n, CMF)
    DevBuf<WlDev> dev2;
  };
  std::vector<WlSlot> wl_cache;
  size_t wl_cache_next = 0;
  const HbWlEntry* wl_cur = nullptr;
  const WlDev* wl2_cur = nullptr;
  WlDev wl0{};
  uint32_t wl_cnt = 0;
  uint32_t layer_idx = 0;       // next layer to trace
  bool layer_traced = false;    // a TraceLayer result is pending Recombine
  uint64_t cont_n = 0;          // continuation pool size after Recombine
  bool cont_shuffle = false;
  uint32_t shuffle_round = 0;

  // monotone stream counters (never reset per session; cuda_trace_backend.cu:3717-3754)
  uint64_t gen_base = 0, gate_base = 0, transit_base = 0;

  // tile buffers
  uint64_t tile_rays = 1u << 24;
  DevBuf<float4> P, D, Q;
  DevBuf<uint8_t> path;
  DevBuf<uint32_t> fork_root, fork_code;
  DevBuf<uint32_t> counters;    // [0] fork_count [1] fork_snapshot [2] cont_count [3] exit_count [4] error [5] done_count
  DevBuf<unsigned long long> stat_cnt;
  DevBuf<double> stat_sum;
  DevBuf<ContRec> cont[2];      // continuation pools (appended by one layer, gathered by the next)
  int cont_cur = 0;             // pool being appended by the current layer
  DevBuf<HbExitRecord> exits_dev;
  DevBuf<uint32_t> exit_root_dev;
  std::vector<HbExitRecord> exits_host;
  std::vector<uint32_t> exit_roots_host;

  // parity: injected / exported roots
  bool injected = false;
  std::vector<float4> inj_P, inj_D, inj_Q;
  std::vector<float> exp_d, exp_p, exp_w, exp_rot;
  std::vector<uint16_t> exp_face;
  std::vector<uint32_t> exp_shape, exp_wl;
  std::vector<uint64_t> exp_mask;

  // measurement
  HbCounters ctr{};
  bool profile = false;
  std::vector<EventPair> ev_pool;
  size_t ev_used = 0;
  int blocks_per_sm = 16;   // generator / image kernels: CTAs per SM of their grid-stride launches
  int blocks_per_sm_override = 0;
  cudaStream_t geom_stream = nullptr;  // geometry-clock redraws run here, beside the trace stream
  uint64_t geom_rejected = 0;   // shapes the device builder rejected since hb_set_scene (empty crystals)
  bool p4_enable = true;  // option "prism_fast_path": 0 forces the generic axis-loop kernels (A/B tests)
  bool pixel_cache = true;
  bool fused_bounce = true;  // option "fused_bounce": 0 runs the split optics + intersect pipeline
  // option "fused_gen": 1 fuses root generation with the entry interaction (genbounce_kernel). Off by default: the
  // path is instruction-issue bound, the fusion only saves the 80 B/root round trip through HBM and runs the
  // generator at the bounce kernel's lower occupancy (measured 1.06 ms against 0.58 + 0.38 ms per 16 Mi roots).
  bool fused_gen = false;
  // Device error word (fork-slot / continuation-pool / exit-record overflow): sticky on the device until it has
  // been REPORTED to the caller. Every TraceLayer enqueues a copy into this pinned mirror; every entry point that
  // synchronises the stream anyway (readback, snapshot, stats, drain, hb_synchronize) checks it.
  bool filter_hit_bound = true;  // option "filter_hit_bound": 0 traces all max_hits interactions of filtered layers too
  uint32_t* err_host = nullptr;
  uint64_t fork_cap_override = 0;   // option "fork_cap" (tests: force an overflow)
  uint64_t cont_cap_override = 0;   // option "cont_cap"
#ifdef HB_WITH_NCCL
  ncclComm_t comm = nullptr;
#endif
  int nranks = 1, rank = 0;
};

namespace {

std::string g_create_error;

#define HB_CUDA(h, expr)                                                                            \
  do {                                                                                              \
    cudaError_t e_ = (expr);                                                                        \
    if (e_ != cudaSuccess) {                                                                        \
      (h)->error = std::string(#expr) + ": " + cudaGetErrorString(e_);                              \
      return HB_ERR_CUDA;                                                                           \
    }                                                                                               \
  } while (0)

int fail(HbEngine* h, int code, const std::string& msg) {
  h->error = msg;
  return code;
}

// After a stream synchronisation: report a pending device overflow exactly once, then re-arm the flag.
int check_device_error(HbEngine* h) {
  if (h->err_host == nullptr || *h->err_host == 0u) return HB_OK;
  const uint32_t code = *h->err_host;
  *h->err_host = 0u;
  if (h->counters.p) cudaMemsetAsync(h->counters.p + 4, 0, sizeof(uint32_t), h->stream);
  const char* what = code == 1u ? "continuation pool" : code == 2u ? "exit record buffer" : "fork-ray slots";
  return fail(h, HB_ERR_CAPACITY, std::string("device buffer overflow: ") + what + " (code " + std::to_string(code) +
                                      "); rays were dropped, the frame is incomplete");
}

void flush_events(HbEngine* h) {
  for (size_t i = 0; i < h->ev_used; i++) {
    float ms = 0.0f;
    EventPair& e = h->ev_pool[i];
    if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) {
      if (e.family == 0) {
        h->ctr.gen_ms += ms;
      } else if (e.family == 1) {
        h->ctr.optics_ms += ms;
      } else if (e.family == 2) {
        h->ctr.intersect_ms += ms;
      } else if (e.family == 3) {
        h->ctr.bounce_ms += ms;
      } else {
        h->ctr.genbounce_ms += ms;
      }
    }
  }
  h->ev_used = 0;
}

EventPair* begin_event(HbEngine* h, int family, uint64_t rays) {
  if (!h->profile) return nullptr;
  if (h->ev_used == h->ev_pool.size()) {
    if (h->ev_pool.size() >= 8192) {
      cudaStreamSynchronize(h->stream);
      flush_events(h);
    } else {
      EventPair e{};
      cudaEventCreate(&e.a);
      cudaEventCreate(&e.b);
      h->ev_pool.push_back(e);
    }
  }
  EventPair* e = &h->ev_pool[h->ev_used++];
  e->family = family;
  e->rays = rays;
  cudaEventRecord(e->a, h->stream);
  return e;
}
void end_event(HbEngine* h, EventPair* e) {
  if (e) cudaEventRecord(e->b, h->stream);
}

uint32_t grid_for(const HbEngine* h, uint64_t n) {
  uint64_t blocks = (n + 255) / 256;
  uint64_t cap = static_cast<uint64_t>(h->sm_count) * h->blocks_per_sm;
  return static_cast<uint32_t>(std::max<uint64_t>(1, std::min(blocks, cap)));
}

// Fold the fp32 working arena (all renders) into the fp64 master.
void fold_arena(HbEngine* h) {
  if (h->arena_pix == 0) return;
  fold_image_kernel<<<grid_for(h, h->arena_pix), 256, 0, h->stream>>>(h->image.p, h->master.p, h->arena_pix);
  h->ctr.kernel_launches++;
  h->rays_since_fold = 0;
}

LaunchCtx launch_ctx(const HbEngine* h) { return LaunchCtx{ h->device, h->sm_count, h->blocks_per_sm_override, h->stream }; }

size_t trace_smem(const LayerDev& L) { return shared_tables_bytes(L.shape_cnt); }

int upload_layer(HbEngine* h, const HbLayer& src, LayerDev* L) {
  L->prob = src.prob;
  L->pops.clear();
  std::vector<float4> planes;
  std::vector<float4> axes;
  std::vector<uint32_t> meta;
  std::vector<uint8_t> fn;
  std::vector<HbCrystalTables> shapes;
  std::vector<EntryFaces> entry_faces;
  std::vector<HbFilterDesc> filters;
  std::vector<uint32_t> cid;
  std::vector<float> luts;
  std::vector<HbColorGroup> cgroups;
  std::vector<uint32_t> cgroup_cnt;
  L->any_filter = false;
  L->any_color = false;
  ShapeBufs& B0 = L->sb[0];
  L->cur = 0;
  B0.pop_p4.clear();
  B0.pop_shapes.clear();
  if (src.population_cnt == 0 || src.population_cnt > HB_MAX_CRYSTALS || src.populations == nullptr)
    return fail(h, HB_ERR_INVALID_ARG, "layer: population_cnt must be 1..HB_MAX_CRYSTALS");
  for (uint32_t ci = 0; ci < src.population_cnt; ci++) {
    const HbCrystalPopulation& p = src.populations[ci];
    if (p.shape_cnt == 0 || p.shapes == nullptr) return fail(h, HB_ERR_INVALID_ARG, "population without shapes");
    PopHost ph{};
    ph.proportion = p.proportion;
    ph.crystal_id = p.crystal_id;
    ph.shape_base = static_cast<uint32_t>(shapes.size());
    ph.shape_cnt = p.shape_cnt;
    ph.axis = AxisParams{ p.axis.lat_path, p.axis.lat_mean, p.axis.lat_std, p.axis.az_type, p.axis.az_mean, p.axis.az_std,
                          p.axis.roll_type, p.axis.roll_mean, p.axis.roll_std, p.axis.lut_n };
    ph.has_filter = p.filter.kind != 0;
    uint32_t pop_p4 = 0;  // shapes with four axes, all paired (hexagonal prism with all eight faces)
    L->any_filter = L->any_filter || ph.has_filter;
    for (uint32_t s = 0; s < p.shape_cnt; s++) {
      const HbCrystalTables& t = p.shapes[s];
      if (t.face_cnt > HB_MAX_FACES || t.subtri_cnt > HB_MAX_SUBTRIS) return fail(h, HB_ERR_INVALID_ARG, "crystal table too large");
      shapes.push_back(t);
      planes.resize(planes.size() + HB_MAX_FACES);
      fn.resize(fn.size() + HB_MAX_FACES);
      axes.resize(axes.size() + HB_MAX_FACES * 2);
      meta.push_back(0u);
      entry_faces.emplace_back();
      if (derive_shape_tables(t, ci, &planes[planes.size() - HB_MAX_FACES], &fn[fn.size() - HB_MAX_FACES],
                              &axes[axes.size() - HB_MAX_FACES * 2], &meta.back(), &entry_faces.back()))
        pop_p4++;
    }
    B0.pop_p4.push_back(pop_p4);
    B0.pop_shapes.push_back(p.shape_cnt);
    filters.push_back(p.filter);
    {
      const uint32_t b = filter_max_len(p.filter);
      L->hit_bound = ci == 0 ? b : std::max(L->hit_bound, b);
    }
    cid.push_back(p.crystal_id);
    if (p.color_group_cnt > HB_MAX_COLOR_GROUPS) return fail(h, HB_ERR_INVALID_ARG, "population: too many colour groups");
    cgroup_cnt.push_back(p.color_group_cnt);
    cgroups.insert(cgroups.end(), p.color_groups, p.color_groups + HB_MAX_COLOR_GROUPS);
    L->any_color = L->any_color || p.color_group_cnt != 0;
    luts.insert(luts.end(), p.axis.lut_theta, p.axis.lut_theta + HB_LUT_NODES);
    luts.insert(luts.end(), p.axis.lut_cdf, p.axis.lut_cdf + HB_LUT_NODES);
    luts.insert(luts.end(), p.axis.lut_flip, p.axis.lut_flip + HB_LUT_NODES);
    L->pops.push_back(ph);
  }
  if (shapes.size() > 65535) return fail(h, HB_ERR_CAPACITY, "more than 65535 crystal shapes in one layer");
  L->shape_cnt = static_cast<uint32_t>(shapes.size());
  L->carry.assign(L->pops.size(), 0.0);
  HB_CUDA(h, B0.planes.ensure(planes.size()));
  HB_CUDA(h, B0.axes.ensure(axes.size()));
  HB_CUDA(h, B0.shape_meta.ensure(meta.size()));
  HB_CUDA(h, B0.face_fn.ensure(fn.size()));
  HB_CUDA(h, B0.shapes.ensure(shapes.size()));
  HB_CUDA(h, B0.geom_scalars.ensure(shapes.size() * 10));
  B0.allocated = true;
  HB_CUDA(h, L->filters.ensure(filters.size()));
  HB_CUDA(h, L->pop_crystal_id.ensure(cid.size()));
  HB_CUDA(h, L->luts.ensure(luts.size()));
  HB_CUDA(h, cudaMemcpy(B0.planes.p, planes.data(), planes.size() * sizeof(float4), cudaMemcpyHostToDevice));
  HB_CUDA(h, cudaMemcpy(B0.axes.p, axes.data(), axes.size() * sizeof(float4), cudaMemcpyHostToDevice));
  HB_CUDA(h, cudaMemcpy(B0.shape_meta.p, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice));
  HB_CUDA(h, cudaMemcpy(B0.face_fn.p, fn.data(), fn.size(), cudaMemcpyHostToDevice));
  HB_CUDA(h, cudaMemcpy(B0.shapes.p, shapes.data(), shapes.size() * sizeof(HbCrystalTables), cudaMemcpyHostToDevice));
  HB_CUDA(h, B0.entry_faces.ensure(entry_faces.size()));
  HB_CUDA(h, cudaMemcpy(B0.entry_faces.p, entry_faces.data(), entry_faces.size() * sizeof(EntryFaces), cudaMemcpyHostToDevice));
  HB_CUDA(h, cudaMemcpy(L->filters.p, filters.data(), filters.size() * sizeof(HbFilterDesc), cudaMemcpyHostToDevice));
  HB_CUDA(h, cudaMemcpy(L->pop_crystal_id.p, cid.data(), cid.size() * 4, cudaMemcpyHostToDevice));
  HB_CUDA(h, cudaMemcpy(L->luts.p, luts.data(), luts.size() * 4, cudaMemcpyHostToDevice));
  if (L->any_color) {
    HB_CUDA(h, L->color_groups.ensure(cgroups.size()));
    HB_CUDA(h, L->color_group_cnt.ensure(cgroup_cnt.size()));
    HB_CUDA(h, cudaMemcpy(L->color_groups.p, cgroups.data(), cgroups.size() * sizeof(HbColorGroup), cudaMemcpyHostToDevice));
    HB_CUDA(h, cudaMemcpy(L->color_group_cnt.p, cgroup_cnt.data(), cgroup_cnt.size() * 4, cudaMemcpyHostToDevice));
  }
  return HB_OK;
}

uint32_t session_flags(const HbEngine* h, const LayerDev& L, bool last_layer) {
  uint32_t f = 0;
  if (h->spec.accumulate) f |= kFlagAccum;
  if (h->spec.accumulate && h->pixel_cache) f |= kFlagPixelCache;
  if (h->spec.record_exits) f |= kFlagRecord | kFlagPath | kFlagStats;
  if (L.any_filter || L.any_color) f |= kFlagPath;
  if (L.prob > 0.0f) f |= kFlagGate;
  (void)last_layer;
  return f;
}

int ensure_tile(HbEngine* h, uint64_t cap, uint64_t fork_cap, uint32_t flags) {
  HB_CUDA(h, h->P.ensure(cap));
  HB_CUDA(h, h->D.ensure(cap));
  HB_CUDA(h, h->Q.ensure(cap));
  HB_CUDA(h, h->counters.ensure(8));
  HB_CUDA(h, h->stat_cnt.ensure(1));
  HB_CUDA(h, h->stat_sum.ensure(1));
  HB_CUDA(h, h->fork_root.ensure(fork_cap));
  HB_CUDA(h, h->fork_code.ensure(fork_cap));
  if (flags & kFlagPath) HB_CUDA(h, h->path.ensure(cap * h->max_hits));
  if (h->classes.class_cnt != 0) HB_CUDA(h, h->mask_tile.ensure(cap));
  return HB_OK;
}

// Trace one tile [root0, root0 + n) of layer `li`; roots come from gen (li == 0, not injected),
// the injected host batch, or the continuation pool (li > 0).
int trace_tile(HbEngine* h, uint32_t li, uint64_t root0, uint32_t n, const std::vector<uint64_t>& pop_begin,
               uint32_t flags, uint64_t layer_total) {
  LayerDev& L = *h->layers[li];
  const uint32_t fork_cap = h->fork_cap_override ? static_cast<uint32_t>(h->fork_cap_override) : std::max<uint32_t>(4096u, n / 64u);
  const uint32_t cap = n + fork_cap;
  int rc = ensure_tile(h, cap, fork_cap, flags);
  if (rc != HB_OK) return rc;
  const bool color_on = h->classes.class_cnt != 0;
  const bool general = (flags & (kFlagPath | kFlagRecord | kFlagGate | kFlagStats)) != 0 || !(flags & kFlagAccum) ||
                       h->renders.size() > 1 || color_on;
  HB_CUDA(h, cudaMemsetAsync(h->counters.p, 0, 2 * sizeof(uint32_t), h->stream));  // fork_count, fork_snapshot
  if (flags & kFlagRecord) {
    const uint64_t ecap = static_cast<uint64_t>(n) * (h->max_hits + 1) + fork_cap;
    HB_CUDA(h, h->exits_dev.ensure(ecap));
    HB_CUDA(h, h->exit_root_dev.ensure(ecap));
    HB_CUDA(h, cudaMemsetAsync(h->counters.p + 3, 0, sizeof(uint32_t), h->stream));
  }

  // ---- kernel parameters of the hit loop ----
  TraceParams tp{};
  tp.P = h->P.p;
  tp.D = h->D.p;
  tp.Q = h->Q.p;
  tp.path = h->path.p;
  tp.fork_root = h->fork_root.p;
  tp.fork_code = h->fork_code.p;
  tp.fork_count = h->counters.p + 0;
  tp.fork_snapshot = h->counters.p + 1;
  tp.done_count = h->counters.p + 5;
  tp.n_main = n;
  tp.cap = cap;
  tp.fork_cap = fork_cap;
  tp.root_base = static_cast<uint32_t>(root0);
  tp.lt = LayerTables{ L.b().planes.p, L.b().axes.p, L.b().shape_meta.p, L.b().face_fn.p, L.b().shapes.p, L.filters.p, L.pop_crystal_id.p,
                       L.shape_cnt, static_cast<uint32_t>(L.pops.size()), L.any_filter ? 1u : 0u,
                       L.any_color ? L.color_groups.p : nullptr, L.color_group_cnt.p };
  tp.wl = h->wl_cur;
  tp.wl_cnt = h->wl_cnt;
  tp.wl0 = h->wl0;
  tp.wl2 = h->wl2_cur;
  tp.image = h->image.p;
  tp.master = h->master.p;
  tp.proj = h->proj;
  tp.extra = h->extra_dev.p;
  tp.extra_cnt = static_cast<uint32_t>(h->renders.size()) - 1u;
  tp.color_on = color_on ? 1u : 0u;
  tp.M = (color_on && li > 0) ? h->mask_tile.p : nullptr;
  tp.lane = h->lanes.p;
  tp.lane_stride = h->renders.empty() ? 0u : static_cast<uint32_t>(h->renders[0].img_w) * h->renders[0].img_h;
  tp.classes = h->classes;
  tp.max_hits = h->max_hits;
  tp.layer_idx = li;
  tp.flags = flags;
  tp.prob = L.prob;
  tp.gate_seed = h->spec.seed ^ kNonceGate;
  tp.gate_base_lo = static_cast<uint32_t>(h->gate_base);
  tp.gate_base_hi = static_cast<uint32_t>(h->gate_base >> 32);
  tp.cont = h->cont[h->cont_cur].p;
  tp.cont_count = h->counters.p + 2;
  tp.cont_cap = static_cast<uint32_t>(h->cont[h->cont_cur].n);
  if (h->cont_cap_override) tp.cont_cap = static_cast<uint32_t>(std::min<uint64_t>(tp.cont_cap, h->cont_cap_override));
  tp.exits = h->exits_dev.p;
  tp.exit_root = h->exit_root_dev.p;
  tp.exit_count = h->counters.p + 3;
  tp.exit_cap = static_cast<uint32_t>(std::min<size_t>(h->exits_dev.n, 0xFFFFFFFFu));
  tp.stat_exit_count = h->stat_cnt.p;
  tp.stat_w_sum = h->stat_sum.p;
  tp.error_flag = h->counters.p + 4;

  const size_t smem = trace_smem(L);
  const bool in_smem = L.shape_cnt <= kSmemShapes;
  const int p4_mode = h->p4_enable ? L.b().p4_mode() : 0;
  // Root generation fused with the entry interaction (genbounce_kernel). Parity sessions export the roots
  // between generation and hit 0 and injected rays skip generation: both take the separate gen kernel.
  const bool fused_gen = h->fused_bounce && h->fused_gen && h->max_hits > 1 && (!h->filter_hit_bound || L.hit_bound > 1u) && !(li == 0 && h->injected) &&
                         h->spec.record_exits != 1u;

  // ---- roots ----
  if (li == 0 && h->injected) {
    HB_CUDA(h, cudaMemcpyAsync(h->P.p, h->inj_P.data() + root0, n * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
    HB_CUDA(h, cudaMemcpyAsync(h->D.p, h->inj_D.data() + root0, n * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
    HB_CUDA(h, cudaMemcpyAsync(h->Q.p, h->inj_Q.data() + root0, n * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
    if (flags & kFlagPath) {
      std::vector<uint8_t> p0(n);
      for (uint32_t i = 0; i < n; i++) p0[i] = static_cast<uint8_t>(bits_face_host(h->inj_P[root0 + i]));
      HB_CUDA(h, cudaMemcpyAsync(h->path.p, p0.data(), n, cudaMemcpyHostToDevice, h->stream));
      HB_CUDA(h, cudaStreamSynchronize(h->stream));
    }
  } else {
    for (size_t ci = 0; ci < L.pops.size(); ci++) {
      const uint64_t b = std::max<uint64_t>(pop_begin[ci], root0), e = std::min<uint64_t>(pop_begin[ci + 1], root0 + n);
      if (b >= e) continue;
      const PopHost& ph = L.pops[ci];
      GenParams gp{};
      gp.P = h->P.p;
      gp.D = h->D.p;
      gp.Q = h->Q.p;
      gp.path = h->path.p;
      gp.slot0 = static_cast<uint32_t>(b - root0);
      gp.count = static_cast<uint32_t>(e - b);
      gp.cap = cap;
      const uint64_t base = (li == 0 ? h->gen_base : h->transit_base) + b;
      gp.idx_lo = static_cast<uint32_t>(base);
      gp.idx_hi = static_cast<uint32_t>(base >> 32);
      gp.seed = h->spec.seed ^ (li == 0 ? kNonceGen : kNonceTransit);
      gp.axis = ph.axis;
      gp.lut = L.luts.p + ci * 3 * HB_LUT_NODES;
      gp.shapes = L.b().shapes.p + ph.shape_base;
      gp.entry_faces = L.b().entry_faces.p + ph.shape_base;
      gp.shape_base = ph.shape_base;
      gp.shape_cnt = ph.shape_cnt;
      gp.wl = h->wl_cur;
      gp.wl_cnt = h->wl_cnt;
      gp.sun_lon = h->sun_lon;
      gp.sun_lat = h->sun_lat;
      gp.sun_half = h->sun_half;
      gp.sun_c_cap = std::cos(h->sun_half);
      gp.sun_c_lon = std::cos(h->sun_lon);
      gp.sun_s_lon = std::sin(h->sun_lon);
      gp.sun_c_lat = std::cos(h->sun_lat);
      gp.sun_s_lat = std::sin(h->sun_lat);
      gp.flags = flags;
      EventPair* ev = begin_event(h, fused_gen ? 4 : 0, gp.count);
      const size_t gb_smem = kGenSharedBytes + kQueueBytes + kQueue2Bytes + ((flags & kFlagPixelCache) ? kCacheBytes : 0) + smem;
      if (li == 0) {
        if (fused_gen) {
          tp.hit = 0;
          launch_genbounce(launch_ctx(h), false, general, in_smem, p4_mode, gb_smem, gp, tp);
        } else {
          // 16 CTAs per SM (3.2 waves at 5 co-resident CTAs): measured 0.565 ms against 0.584 (8 per SM) and 0.605 (exactly co-resident)
          gen_kernel<false><<<grid_for(h, gp.count), 256, sizeof(GenShared), h->stream>>>(gp);
        }
      } else {
        const int src = h->cont_cur ^ 1;
        gp.cont = h->cont[src].p;
        if (color_on) gp.M = h->mask_tile.p;
        gp.cont_n = static_cast<uint32_t>(layer_total);
        gp.cont_first = static_cast<uint32_t>(b);
        gp.shuffle = h->cont_shuffle ? 1u : 0u;
        gp.shuffle_seed = pcg_hash_host(h->spec.seed ^ kNonceShuffle ^ h->shuffle_round);
        if (fused_gen) {
          tp.hit = 0;
          launch_genbounce(launch_ctx(h), true, general, in_smem, p4_mode, gb_smem, gp, tp);
        } else {
          gen_kernel<true><<<grid_for(h, gp.count), 256, kGenSharedBytes + kTransitSrcBytes, h->stream>>>(gp);
        }
      }
      end_event(h, ev);
      h->ctr.kernel_launches++;
      if (fused_gen) {
        h->ctr.genbounce_launches++;
        h->ctr.genbounce_rays += gp.count;
      } else {
        h->ctr.gen_launches++;
      }
    }
  }
  HB_CUDA(h, cudaGetLastError());

  // ---- parity export of the roots this tile starts from ----
  if (h->spec.record_exits == 1u && !(li == 0 && h->injected)) {  // parity sessions only (2 = plain egress)
    std::vector<float4> hp(n), hd(n), hq(n);
    DevBuf<float> rot;
    HB_CUDA(h, rot.ensure(static_cast<size_t>(n) * 9));
    quat_to_rot_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(h->Q.p, rot.p, n);
    std::vector<float> hr(static_cast<size_t>(n) * 9);
    HB_CUDA(h, cudaMemcpyAsync(hp.data(), h->P.p, n * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
    HB_CUDA(h, cudaMemcpyAsync(hd.data(), h->D.p, n * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
    HB_CUDA(h, cudaMemcpyAsync(hr.data(), rot.p, hr.size() * 4, cudaMemcpyDeviceToHost, h->stream));
    HB_CUDA(h, cudaStreamSynchronize(h->stream));
    rot.release();
    for (uint32_t i = 0; i < n; i++) {
      uint32_t bits;
      std::memcpy(&bits, &hp[i].w, 4);
      h->exp_p.insert(h->exp_p.end(), { hp[i].x, hp[i].y, hp[i].z });
      h->exp_d.insert(h->exp_d.end(), { hd[i].x, hd[i].y, hd[i].z });
      h->exp_w.push_back(hd[i].w);
      const uint32_t f = bits & 63u;
      h->exp_face.push_back(f == kFaceInvalid ? static_cast<uint16_t>(HB_INVALID_FACE) : static_cast<uint16_t>(f));
      h->exp_shape.push_back((bits >> 14) & 65535u);
      h->exp_wl.push_back((bits >> 6) & 255u);
    }
    h->exp_rot.insert(h->exp_rot.end(), hr.begin(), hr.end());
    if (color_on) {  // component masks the roots carry in from earlier layers
      const size_t old = h->exp_mask.size();
      h->exp_mask.resize(old + n, 0ull);
      if (li > 0) HB_CUDA(h, cudaMemcpy(h->exp_mask.data() + old, h->mask_tile.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    }
  }

  // ---- hit loop ----
  // A layer whose every population carries a length-bounded filter_in filter stops at that length: deeper
  // interactions can only produce exits the filter rejects (hb_filter.h, filter_max_len).
  const uint32_t hits = h->filter_hit_bound ? std::min<uint32_t>(h->max_hits, std::max<uint32_t>(1u, L.hit_bound)) : h->max_hits;
  for (uint32_t hit = fused_gen ? 1u : 0u; hit < hits; hit++) {
    tp.hit = hit;
    const bool last = hit + 1 == hits;
    if (h->fused_bounce) {  // one launch per interaction (DESIGN.md "kernels")
      EventPair* ev = begin_event(h, 3, n);
      launch_bounce(launch_ctx(h), general, last, in_smem, p4_mode,
                    smem + kStage2Bytes + kQueueBytes + kQueue2Bytes + ((flags & kFlagPixelCache) ? kCacheBytes : 0), tp);
      end_event(h, ev);
      h->ctr.kernel_launches++;
      h->ctr.bounce_launches++;
      h->ctr.bounce_rays += n;
      continue;
    }
    EventPair* ev = begin_event(h, 1, n);
    launch_optics(launch_ctx(h), general, last, in_smem, p4_mode, smem + kStageBytes + ((flags & kFlagPixelCache) ? kCacheBytes : 0), tp);
    end_event(h, ev);
    h->ctr.kernel_launches++;
    h->ctr.optics_launches++;
    h->ctr.optics_rays += n;
    if (!last) {
      ev = begin_event(h, 2, n);
      launch_intersect(launch_ctx(h), general, in_smem, p4_mode, smem, tp);
      end_event(h, ev);
      h->ctr.kernel_launches++;
      h->ctr.intersect_launches++;
      h->ctr.intersect_rays += n;
    }
  }
  HB_CUDA(h, cudaGetLastError());

  if (flags & kFlagRecord) {  // move this tile's exit records to the host (DrainExits is grow-not-clamp)
    uint32_t cnt = 0;
    HB_CUDA(h, cudaMemcpyAsync(&cnt, h->counters.p + 3, 4, cudaMemcpyDeviceToHost, h->stream));
    HB_CUDA(h, cudaStreamSynchronize(h->stream));
    if (cnt > h->exits_dev.n) return fail(h, HB_ERR_CAPACITY, "exit record buffer overflow");
    const size_t old = h->exits_host.size();
    h->exits_host.resize(old + cnt);
    h->exit_roots_host.resize(old + cnt);
    if (cnt) {
      HB_CUDA(h, cudaMemcpy(h->exits_host.data() + old, h->exits_dev.p, cnt * sizeof(HbExitRecord), cudaMemcpyDeviceToHost));
      HB_CUDA(h, cudaMemcpy(h->exit_roots_host.data() + old, h->exit_root_dev.p, cnt * 4, cudaMemcpyDeviceToHost));
    }
  }
  h->ctr.rays_traced += n;
  h->allreduced = false;
  if (flags & kFlagAccum) {
    h->rays_since_fold += n;
    if (h->rays_since_fold >= h->fold_rays) fold_arena(h);
  }
  return HB_OK;
}

}  // namespace

// host twins of two tiny device helpers (kept out of the header: host code never traces rays)
namespace hb {
uint32_t pcg_hash_host(uint32_t x) {
  x = x * 747796405u + 2891336453u;
  x = ((x >> ((x >> 28u) + 4u)) ^ x) * 277803737u;
  return (x >> 22u) ^ x;
}
uint32_t bits_face_host(const float4& p) {
  uint32_t b;
  std::memcpy(&b, &p.w, 4);
  return b & 63u;
}
}  // namespace hb

extern "C" {

const char* hb_last_error(const HbEngine* h) {
  if (h == nullptr) return g_create_error.empty() ? hb::global_error().c_str() : g_create_error.c_str();
  return h->error.c_str();
}

int hb_create(int device, HbEngine** out) {
  if (out == nullptr) return HB_ERR_INVALID_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this engine has no CPU path)";
    return HB_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= count) {
    g_create_error = "device ordinal out of range";
    return HB_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop{};
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) {
    g_create_error = "device is not sm_100 or newer (this library carries sm_100a code only)";
    return HB_ERR_NO_DEVICE;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    g_create_error = "cudaSetDevice failed";
    return HB_ERR_CUDA;
  }
  auto h = std::make_unique<HbEngine>();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    g_create_error = "cudaStreamCreate failed";
    return HB_ERR_CUDA;
  }
  if (h->counters.ensure(8) != cudaSuccess || cudaMemset(h->counters.p, 0, 8 * sizeof(uint32_t)) != cudaSuccess ||
      cudaMallocHost(&h->err_host, sizeof(uint32_t)) != cudaSuccess) {
    g_create_error = "device counter allocation failed";
    return HB_ERR_CUDA;
  }
  *h->err_host = 0u;
  *out = h.release();
  return HB_OK;
}

void hb_destroy(HbEngine* h) {
  if (h == nullptr) return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  for (auto& e : h->ev_pool) {
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  if (h->geom_stream) {
    cudaStreamSynchronize(h->geom_stream);
    cudaStreamDestroy(h->geom_stream);
  }
  for (auto& L : h->layers) L->release();
  h->image.release();
  h->master.release();
  h->extra_dev.release();
  h->lanes.release();
  h->mask_tile.release();
  h->rgb_stage.release();
  h->xyz_stage.release();
  h->landed_dev.release();
  h->reduce_stage.release();
  if (h->merge_event) cudaEventDestroy(h->merge_event);
  for (auto& s : h->wl_cache) {
    s.dev.release();
    s.dev2.release();
  }
  h->P.release();
  h->D.release();
  h->Q.release();
  h->path.release();
  h->fork_root.release();
  h->fork_code.release();
  h->counters.release();
  h->stat_cnt.release();
  h->stat_sum.release();
  for (int i = 0; i < 2; i++) {
    h->cont[i].release();
  }
  h->exits_dev.release();
  h->exit_root_dev.release();
  if (h->err_host) cudaFreeHost(h->err_host);
#ifdef HB_WITH_NCCL
  if (h->comm) ncclCommDestroy(h->comm);
#endif
  cudaStreamDestroy(h->stream);
  delete h;
}

int hb_set_scene(HbEngine* h, const HbScene* s) {
  if (h == nullptr || s == nullptr) return HB_ERR_INVALID_ARG;
  if (h->in_session) return fail(h, HB_ERR_STATE, "hb_set_scene inside a session");
  if (s->layer_cnt == 0 || s->layer_cnt > HB_MAX_LAYERS || s->max_hits == 0 || s->max_hits > HB_MAX_HITS)
    return fail(h, HB_ERR_INVALID_ARG, "scene: layer_cnt must be 1..8 and max_hits 1..64");
  cudaSetDevice(h->device);
  HB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->geom_stream) HB_CUDA(h, cudaStreamSynchronize(h->geom_stream));
  if (s->color_classes.class_cnt > HB_MAX_COLOR_CLASSES) return fail(h, HB_ERR_INVALID_ARG, "scene: more than 16 colour classes");
  for (auto& L : h->layers) L->release();
  h->layers.clear();
  h->have_scene = false;  // a failed upload leaves the engine without a scene, not with half of the old one
  for (uint32_t li = 0; li < s->layer_cnt; li++) {
    auto L = std::make_unique<LayerDev>();
    int rc = upload_layer(h, s->layers[li], L.get());
    if (rc != HB_OK) {
      L->release();
      for (auto& K : h->layers) K->release();
      h->layers.clear();
      return rc;
    }
    h->layers.push_back(std::move(L));
  }
  h->max_hits = s->max_hits;
  h->sun_lon = s->sun_lon;
  h->sun_lat = s->sun_lat;
  h->sun_half = s->sun_half_angle;
  h->classes = s->color_classes;
  h->have_scene = true;
  return HB_OK;
}

int hb_set_renders(HbEngine* h, uint32_t n, const HbProjParams* p) {
  if (h == nullptr || p == nullptr) return HB_ERR_INVALID_ARG;
  if (h->in_session) return fail(h, HB_ERR_STATE, "hb_set_render inside a session");
  if (n == 0 || n > HB_MAX_RENDERS) return fail(h, HB_ERR_INVALID_ARG, "render count must be 1..HB_MAX_RENDERS");
  uint64_t total = 0;
  for (uint32_t r = 0; r < n; r++) {
    if (p[r].img_w <= 0 || p[r].img_h <= 0) return fail(h, HB_ERR_INVALID_ARG, "render: empty image");
    total += static_cast<uint64_t>(p[r].img_w) * p[r].img_h;
  }
  if (total >= (1ull << 31)) return fail(h, HB_ERR_CAPACITY, "renders: more than 2^31 pixels");
  cudaSetDevice(h->device);
  HB_CUDA(h, cudaStreamSynchronize(h->stream));
  bool realloc = !h->have_render || n != h->renders.size();
  for (uint32_t r = 0; r < n && !realloc; r++)
    realloc = p[r].img_w != h->renders[r].img_w || p[r].img_h != h->renders[r].img_h;
  h->renders.assign(p, p + n);
  h->proj = p[0];
  h->render_off.resize(n);
  std::vector<ExtraRender> extra;
  uint32_t off = 0;
  size_t max_pix = 0;
  for (uint32_t r = 0; r < n; r++) {
    h->render_off[r] = off;
    if (r > 0) extra.push_back(ExtraRender{ p[r], off });
    const size_t pix = static_cast<size_t>(p[r].img_w) * p[r].img_h;
    max_pix = std::max(max_pix, pix);
    off += static_cast<uint32_t>(pix);
  }
  h->arena_pix = off;
  if (!extra.empty()) {
    HB_CUDA(h, h->extra_dev.ensure(extra.size()));
    HB_CUDA(h, cudaMemcpy(h->extra_dev.p, extra.data(), extra.size() * sizeof(ExtraRender), cudaMemcpyHostToDevice));
  }
  if (realloc) {  // resolution change => realloc + zero (cuda_trace_backend.cu:3799-3835)
    HB_CUDA(h, h->image.ensure(h->arena_pix));
    HB_CUDA(h, h->master.ensure(h->arena_pix));
    HB_CUDA(h, cudaMemset(h->master.p, 0, static_cast<size_t>(h->arena_pix) * sizeof(double4)));
    h->rays_since_fold = 0;
    HB_CUDA(h, h->xyz_stage.ensure(max_pix * 3));
    HB_CUDA(h, h->rgb_stage.ensure(max_pix * 3));
    HB_CUDA(h, h->landed_dev.ensure(1));
    HB_CUDA(h, cudaMemset(h->image.p, 0, static_cast<size_t>(h->arena_pix) * sizeof(float4)));
    HB_CUDA(h, cudaMemset(h->landed_dev.p, 0, sizeof(double)));
  }
  h->have_render = true;
  return HB_OK;
}

int hb_set_render(HbEngine* h, const HbProjParams* p) { return hb_set_renders(h, 1, p); }

namespace {
int prefetch_geometry(HbEngine* h, LayerDev& L);  // defined with the geometry-pool API below
}

int hb_begin_session(HbEngine* h, const HbSessionSpec* spec) {
  if (h == nullptr || spec == nullptr) return HB_ERR_INVALID_ARG;
  if (h->in_session) return fail(h, HB_ERR_STATE, "BeginSession on an already-open session");
  if (!h->have_scene) return fail(h, HB_ERR_STATE, "BeginSession before hb_set_scene");
  if (spec->accumulate && !h->have_render) return fail(h, HB_ERR_STATE, "accumulate session before hb_set_render");
  if (spec->wl == nullptr || spec->wl_cnt == 0 || spec->wl_cnt > HB_MAX_WL)
    return fail(h, HB_ERR_INVALID_ARG, "session: wavelength pool must hold 1..256 entries");
  cudaSetDevice(h->device);
  // per-class Y lanes (render 0) follow the scene's class count and the render's resolution
  if (spec->accumulate && h->classes.class_cnt != 0) {
    const uint64_t want = static_cast<uint64_t>(h->classes.class_cnt) * h->renders[0].img_w * h->renders[0].img_h;
    if (want != h->lanes_floats) {
      HB_CUDA(h, h->lanes.ensure(want));
      HB_CUDA(h, cudaMemsetAsync(h->lanes.p, 0, want * sizeof(float), h->stream));
      h->lanes_floats = want;
    }
  }
  h->spec = *spec;
  h->wl_cnt = spec->wl_cnt;
  h->spec.wl = nullptr;  // borrowed only for the call
  // Wavelength pools are tiny and recur (one per discrete wavelength): keep the last few on the device so a
  // steady-state BeginSession issues no copy and no synchronisation.
  {
    const size_t bytes = spec->wl_cnt * sizeof(HbWlEntry);
    int hit = -1;
    for (size_t i = 0; i < h->wl_cache.size(); i++) {
      if (h->wl_cache[i].host.size() == spec->wl_cnt && std::memcmp(h->wl_cache[i].host.data(), spec->wl, bytes) == 0) {
        hit = static_cast<int>(i);
        break;
      }
    }
    if (hit < 0) {
      if (h->wl_cache.size() < 32) {
        h->wl_cache.emplace_back();
        hit = static_cast<int>(h->wl_cache.size()) - 1;
      } else {
        hit = static_cast<int>(h->wl_cache_next++ % 32);
        HB_CUDA(h, cudaStreamSynchronize(h->stream));  // the evicted table may still be in use
      }
      HbEngine::WlSlot& s = h->wl_cache[hit];
      s.host.assign(spec->wl, spec->wl + spec->wl_cnt);
      HB_CUDA(h, s.dev.ensure(spec->wl_cnt));
      HB_CUDA(h, cudaMemcpy(s.dev.p, s.host.data(), bytes, cudaMemcpyHostToDevice));
      s.host2.resize(spec->wl_cnt);
      for (uint32_t i = 0; i < spec->wl_cnt; i++) {
        const HbWlEntry& e = s.host[i];
        const volatile float one = 1.0f;  // a correctly rounded single-precision division, whatever the host flags
        s.host2[i] = WlDev{ e.n_idx, one / e.n_idx, e.spd_weight, 0.0f, e.cmf_x, e.cmf_y, e.cmf_z, 0.0f };
      }
      HB_CUDA(h, s.dev2.ensure(spec->wl_cnt));
      HB_CUDA(h, cudaMemcpy(s.dev2.p, s.host2.data(), spec->wl_cnt * sizeof(WlDev), cudaMemcpyHostToDevice));
    }
    h->wl_cur = h->wl_cache[hit].dev.p;
    h->wl2_cur = h->wl_cache[hit].dev2.p;
    h->wl0 = h->wl_cache[hit].host2[0];
  }
  for (auto& L : h->layers) {  // geometry clock: swap in the pool drawn one session ahead, enqueue the next one
    if (!L->any_auto) continue;
    int rc = prefetch_geometry(h, *L);
    if (rc != HB_OK) return rc;
  }
  if (spec->use_ray_base) {
    // Sharded sessions (multi-GPU): EVERY stream of the session is keyed by the global ray index, not
    // only root generation -- ranks tracing disjoint root ranges must also draw disjoint gate and transit streams,
    // or the second-layer samples would be replicated across ranks. A root spawns at most (max_hits + 1) exits per
    // layer, so `span` stream indices per root are enough for all layers of the session.
    h->gen_base = spec->ray_base;
    uint64_t span = 0, per = 1;
    for (size_t l = 0; l < h->layers.size(); l++) {
      span += per;
      per *= static_cast<uint64_t>(h->max_hits) + 1u;
    }
    h->gate_base = h->transit_base = spec->ray_base * span;
  }
  h->layer_idx = 0;
  h->layer_traced = false;
  h->cont_n = 0;
  h->injected = false;
  h->exits_host.clear();
  h->exit_roots_host.clear();
  h->exp_d.clear();
  h->exp_p.clear();
  h->exp_w.clear();
  h->exp_rot.clear();
  h->exp_face.clear();
  h->exp_shape.clear();
  h->exp_wl.clear();
  h->exp_mask.clear();
  h->in_session = true;
  return HB_OK;
}

int hb_inject_rays(HbEngine* h, uint64_t n, const float* d3, const float* p3, const float* w, const uint16_t* to_face,
                   const float* rot9) {
  if (h == nullptr || d3 == nullptr || p3 == nullptr || w == nullptr || to_face == nullptr) return HB_ERR_INVALID_ARG;
  if (!h->in_session || h->layer_idx != 0 || h->layer_traced) return fail(h, HB_ERR_STATE, "inject_rays: only before the first TraceLayer");
  h->inj_P.resize(n);
  h->inj_D.resize(n);
  h->inj_Q.resize(n);
  for (uint64_t i = 0; i < n; i++) {
    const uint32_t f = to_face[i] == HB_INVALID_FACE ? kFaceInvalid : (to_face[i] & 63u);
    const uint32_t bits = (f & 63u);
    float fb;
    std::memcpy(&fb, &bits, 4);
    h->inj_P[i] = make_float4(p3[i * 3], p3[i * 3 + 1], p3[i * 3 + 2], fb);
    h->inj_D[i] = make_float4(d3[i * 3], d3[i * 3 + 1], d3[i * 3 + 2], w[i]);
    float4 q = make_float4(1.0f, 0.0f, 0.0f, 0.0f);
    if (rot9 != nullptr) {  // rotation matrix -> unit quaternion (approximate inverse of rot_from_quat)
      const float* m = rot9 + i * 9;
      const float tr = m[0] + m[4] + m[8];
      if (tr > 0.0f) {
        float s = std::sqrt(tr + 1.0f) * 2.0f;
        q = make_float4(0.25f * s, (m[7] - m[5]) / s, (m[2] - m[6]) / s, (m[3] - m[1]) / s);
      } else if (m[0] > m[4] && m[0] > m[8]) {
        float s = std::sqrt(1.0f + m[0] - m[4] - m[8]) * 2.0f;
        q = make_float4((m[7] - m[5]) / s, 0.25f * s, (m[1] + m[3]) / s, (m[2] + m[6]) / s);
      } else if (m[4] > m[8]) {
        float s = std::sqrt(1.0f + m[4] - m[0] - m[8]) * 2.0f;
        q = make_float4((m[2] - m[6]) / s, (m[1] + m[3]) / s, 0.25f * s, (m[5] + m[7]) / s);
      } else {
        float s = std::sqrt(1.0f + m[8] - m[0] - m[4]) * 2.0f;
        q = make_float4((m[3] - m[1]) / s, (m[2] + m[6]) / s, (m[5] + m[7]) / s, 0.25f * s);
      }
    }
    h->inj_Q[i] = q;
  }
  h->injected = true;
  return HB_OK;
}

int hb_trace_layer(HbEngine* h, uint64_t n_roots, HbLayerStats* stats) {
  if (h == nullptr) return HB_ERR_INVALID_ARG;
  if (!h->in_session) return fail(h, HB_ERR_STATE, "TraceLayer outside BeginSession/EndSession");
  if (h->layer_traced) return fail(h, HB_ERR_STATE, "TraceLayer called twice without Recombine");
  if (h->layer_idx >= h->layers.size()) return fail(h, HB_ERR_STATE, "TraceLayer beyond the configured layers");
  cudaSetDevice(h->device);
  const uint32_t li = h->layer_idx;
  LayerDev& L = *h->layers[li];
  uint64_t n = 0;
  if (li == 0) {
    n = h->injected ? h->inj_P.size() : n_roots;
  } else {
    if (n_roots != 0) return fail(h, HB_ERR_INVALID_ARG, "continuation layers take n_roots == 0 (RootRaySource::FromDevice)");
    n = h->cont_n;
  }
  if (n >= (1ull << 31)) return fail(h, HB_ERR_CAPACITY, "one TraceLayer call is limited to 2^31 - 1 rays; split the batch");
  const bool last_layer = li + 1 == h->layers.size();
  uint32_t flags = session_flags(h, L, last_layer);
  if (stats != nullptr) flags |= kFlagStats;  // LayerStats asked for: count exits and their weight (general kernels)

  // PartitionCrystalRayNum: contiguous index range per population
  std::vector<uint64_t> counts(L.pops.size(), 0), pop_begin(L.pops.size() + 1, 0);
  if (li == 0 && h->injected) {
    counts[0] = n;
  } else if (n > 0) {
    std::vector<float> prop;
    for (auto& p : L.pops) prop.push_back(p.proportion);
    hb_partition_rays(prop.data(), static_cast<uint32_t>(prop.size()), n, L.carry.data(), counts.data());
  }
  for (size_t i = 0; i < counts.size(); i++) pop_begin[i + 1] = pop_begin[i] + counts[i];
  // All proportions <= 0 assign no ray at all (the reference traces nothing then): trace what was assigned, never
  // the stale state of slots no generator wrote.
  n = pop_begin.back();

  // continuation pool of this layer (only when its exits can continue)
  HB_CUDA(h, h->counters.ensure(8));
  HB_CUDA(h, h->stat_cnt.ensure(1));
  HB_CUDA(h, h->stat_sum.ensure(1));
  HB_CUDA(h, cudaMemsetAsync(h->counters.p + 2, 0, 2 * sizeof(uint32_t), h->stream));  // cont_count, exit_count
  HB_CUDA(h, cudaMemsetAsync(h->counters.p + 5, 0, sizeof(uint32_t), h->stream));      // done_count (the error word [4] stays)
  HB_CUDA(h, cudaMemsetAsync(h->stat_cnt.p, 0, sizeof(unsigned long long), h->stream));
  HB_CUDA(h, cudaMemsetAsync(h->stat_sum.p, 0, sizeof(double), h->stream));
  if (flags & kFlagGate) {
    uint64_t ccap = std::min<uint64_t>(n * (h->max_hits + 1) + 4096, 0xFFFFFFF0ull);
    if (h->cont_cap_override) ccap = std::min<uint64_t>(ccap, h->cont_cap_override);
    HB_CUDA(h, h->cont[h->cont_cur].ensure(ccap));
  }

  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (stats != nullptr) {
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, h->stream);
  }
  uint64_t tile = h->tile_rays;
  if (flags & kFlagRecord) tile = std::min<uint64_t>(tile, h->spec.record_exits == 1u ? 1u << 18 : 1u << 20);
  for (uint64_t r0 = 0; r0 < n; r0 += tile) {
    const uint32_t cnt = static_cast<uint32_t>(std::min<uint64_t>(tile, n - r0));
    int rc = trace_tile(h, li, r0, cnt, pop_begin, flags, n);
    if (rc != HB_OK) return rc;
  }
  if (li == 0 && !h->injected) h->gen_base += n;
  if (li > 0) h->transit_base += n;
  h->gate_base += n;
  h->layer_traced = true;

  if (stats != nullptr || (flags & kFlagGate)) {
    uint32_t c[3] = { 0, 0, 0 };
    unsigned long long ec = 0;
    double ws = 0.0;
    if (stats != nullptr) cudaEventRecord(e1, h->stream);
    HB_CUDA(h, cudaMemcpyAsync(c, h->counters.p + 2, 12, cudaMemcpyDeviceToHost, h->stream));
    HB_CUDA(h, cudaMemcpyAsync(&ec, h->stat_cnt.p, 8, cudaMemcpyDeviceToHost, h->stream));
    HB_CUDA(h, cudaMemcpyAsync(&ws, h->stat_sum.p, 8, cudaMemcpyDeviceToHost, h->stream));
    HB_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->profile) flush_events(h);
    if (c[2] != 0) {
      *h->err_host = c[2];
      return check_device_error(h);
    }
    h->cont_n = c[0];
    if (stats != nullptr) {
      float ms = 0.0f;
      cudaEventElapsedTime(&ms, e0, e1);
      h->ctr.last_layer_ms = ms;
      stats->root_count = n;
      stats->continuation_count = c[0];
      stats->exit_count = ec;
      stats->exit_w_sum = ws;
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
    }
  } else {
    h->cont_n = 0;
    // no synchronisation here: mirror the error word; the next synchronising entry point reports it
    HB_CUDA(h, cudaMemcpyAsync(h->err_host, h->counters.p + 4, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  }
  return HB_OK;
}

int hb_recombine(HbEngine* h, int shuffle, uint64_t* continuation_count) {
  if (h == nullptr) return HB_ERR_INVALID_ARG;
  if (!h->in_session || !h->layer_traced) return fail(h, HB_ERR_STATE, "Recombine must follow a TraceLayer");
  h->layer_traced = false;
  h->layer_idx++;
  h->cont_shuffle = shuffle != 0 && h->cont_n > 1;
  h->shuffle_round++;
  h->cont_cur ^= 1;  // the pool just filled becomes the source of the next layer
  h->injected = false;
  if (continuation_count) *continuation_count = h->cont_n;
  return HB_OK;
}

int hb_end_session(HbEngine* h) {
  if (h == nullptr) return HB_ERR_INVALID_ARG;
  if (!h->in_session) return fail(h, HB_ERR_STATE, "EndSession without BeginSession");
  h->in_session = false;
  h->layer_traced = false;
  h->layer_idx = 0;
  h->injected = false;
  // Non-blocking: an overflow whose mirror copy has already landed is reported here, otherwise by the next
  // synchronising call (ReadbackXyzAccum at the latest); it is never cleared unreported.
  return check_device_error(h);
}

int hb_readback_xyz_render(HbEngine* h, uint32_t render, float* xyz, float* landed) {
  if (h == nullptr || xyz == nullptr || landed == nullptr) return HB_ERR_INVALID_ARG;
  if (!h->have_render) return fail(h, HB_ERR_STATE, "readback before hb_set_render");
  if (render >= h->renders.size()) return fail(h, HB_ERR_INVALID_ARG, "readback: no such render");
  cudaSetDevice(h->device);
  const uint32_t pix = static_cast<uint32_t>(h->renders[render].img_w) * h->renders[render].img_h;
  fold_arena(h);
  drain_image_kernel<<<grid_for(h, pix), 256, 0, h->stream>>>(h->master.p + h->render_off[render], h->xyz_stage.p,
                                                              h->landed_dev.p, pix);
  h->allreduced = false;
  h->ctr.kernel_launches++;
  double l = 0.0;
  HB_CUDA(h, cudaMemcpyAsync(xyz, h->xyz_stage.p, static_cast<size_t>(pix) * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  HB_CUDA(h, cudaMemcpyAsync(&l, h->landed_dev.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  HB_CUDA(h, cudaMemsetAsync(h->landed_dev.p, 0, sizeof(double), h->stream));
  HB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->profile) flush_events(h);
  *landed += static_cast<float>(l);
  return check_device_error(h);
}

int hb_readback_xyz(HbEngine* h, float* xyz, float* landed) { return hb_readback_xyz_render(h, 0, xyz, landed); }

int hb_readback_class_lanes(HbEngine* h, float* lanes, uint64_t cap_floats, uint32_t* class_count) {
  if (h == nullptr || class_count == nullptr) return HB_ERR_INVALID_ARG;
  *class_count = 0;
  if (!h->have_scene || h->classes.class_cnt == 0 || h->lanes_floats == 0) return HB_OK;  // base impl: empty
  if (lanes == nullptr || cap_floats < h->lanes_floats) return fail(h, HB_ERR_CAPACITY, "ReadbackClassLanes: buffer too small");
  cudaSetDevice(h->device);
  HB_CUDA(h, cudaMemcpyAsync(lanes, h->lanes.p, h->lanes_floats * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  HB_CUDA(h, cudaMemsetAsync(h->lanes.p, 0, h->lanes_floats * sizeof(float), h->stream));
  HB_CUDA(h, cudaStreamSynchronize(h->stream));
  *class_count = h->classes.class_cnt;
  return check_device_error(h);
}

int hb_snapshot(HbEngine* h, uint32_t render, const HbSnapshotDesc* desc, uint8_t* rgb8, float* xyz, float* intensity) {
  if (h == nullptr || desc == nullptr) return HB_ERR_INVALID_ARG;
  if (!h->have_render) return fail(h, HB_ERR_STATE, "snapshot before hb_set_render");
  if (render >= h->renders.size()) return fail(h, HB_ERR_INVALID_ARG, "snapshot: no such render");
  cudaSetDevice(h->device);
  const uint32_t pix = static_cast<uint32_t>(h->renders[render].img_w) * h->renders[render].img_h;
  const double4* src = h->master.p + h->render_off[render];
  fold_arena(h);
  // PrepareSnapshot (render.cpp:463-495): running sum -> fp32 snapshot + total landed weight
  HB_CUDA(h, cudaMemsetAsync(h->landed_dev.p, 0, sizeof(double), h->stream));
  peek_image_kernel<<<grid_for(h, pix), 256, 0, h->stream>>>(src, xyz ? h->xyz_stage.p : nullptr, h->landed_dev.p, pix);
  h->ctr.kernel_launches++;
  double l = 0.0;
  HB_CUDA(h, cudaMemcpyAsync(&l, h->landed_dev.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  HB_CUDA(h, cudaMemsetAsync(h->landed_dev.p, 0, sizeof(double), h->stream));
  if (xyz)
    HB_CUDA(h, cudaMemcpyAsync(xyz, h->xyz_stage.p, static_cast<size_t>(pix) * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  HB_CUDA(h, cudaStreamSynchronize(h->stream));
  {
    int rc = check_device_error(h);
    if (rc != HB_OK) return rc;
  }
  const float snapshot_intensity = static_cast<float>(l);
  if (intensity) *intensity = snapshot_intensity;
  if (rgb8 != nullptr) {
    if (snapshot_intensity <= 0.0f) {  // PostSnapshot early-out: black frame (render.cpp:514-517)
      std::memset(rgb8, 0, static_cast<size_t>(pix) * 3);
      return HB_OK;
    }
    SnapshotParams sp{};
    // ExposureScale (render.cpp:96-102): intensity_factor * kNormScale * total_pix / snapshot_intensity
    sp.scale = desc->intensity_factor * 0.08f * static_cast<float>(static_cast<int>(pix)) / snapshot_intensity;
    sp.use_real_color = desc->ray_color[0] < 0.0f ? 1 : 0;
    for (int j = 0; j < 3; j++) {
      sp.ray_color[j] = desc->ray_color[j];
      sp.background[j] = desc->background[j];
    }
    snapshot_srgb_kernel<<<grid_for(h, pix), 256, 0, h->stream>>>(src, h->rgb_stage.p, pix, sp);
    h->ctr.kernel_launches++;
    HB_CUDA(h, cudaMemcpyAsync(rgb8, h->rgb_stage.p, static_cast<size_t>(pix) * 3, cudaMemcpyDeviceToHost, h->stream));
    HB_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return HB_OK;
}

int hb_drain_exits(HbEngine* h, HbExitRecord* out, uint32_t* root_ids, uint64_t cap, uint64_t* count) {
  if (h == nullptr || count == nullptr) return HB_ERR_INVALID_ARG;
  if (!h->in_session) return fail(h, HB_ERR_STATE, "DrainExits outside a session");
  if (h->spec.record_exits) {  // record sessions are synchronous per tile already: make the error word current
    cudaSetDevice(h->device);
    HB_CUDA(h, cudaStreamSynchronize(h->stream));
    int rc = check_device_error(h);
    if (rc != HB_OK) return rc;
  }
  *count = h->exits_host.size();
  if (out == nullptr) return HB_OK;
  if (cap < h->exits_host.size()) return fail(h, HB_ERR_CAPACITY, "DrainExits: caller buffer too small (grow, not clamp)");
  if (!h->exits_host.empty()) {
    std::memcpy(out, h->exits_host.data(), h->exits_host.size() * sizeof(HbExitRecord));
    if (root_ids) std::memcpy(root_ids, h->exit_roots_host.data(), h->exit_roots_host.size() * 4);
  }
  h->exits_host.clear();
  h->exit_roots_host.clear();
  return HB_OK;
}

int hb_export_roots(HbEngine* h, uint64_t cap, float* d3, float* p3, float* w, uint16_t* to_face, float* rot9,
                    uint32_t* shape_idx, uint32_t* wl_idx, uint64_t* count) {
  if (h == nullptr || count == nullptr) return HB_ERR_INVALID_ARG;
  const uint64_t n = h->exp_w.size();
  *count = n;
  if (d3 == nullptr) return HB_OK;
  if (cap < n) return fail(h, HB_ERR_CAPACITY, "export_roots: caller buffer too small");
  std::memcpy(d3, h->exp_d.data(), n * 12);
  std::memcpy(p3, h->exp_p.data(), n * 12);
  std::memcpy(w, h->exp_w.data(), n * 4);
  std::memcpy(to_face, h->exp_face.data(), n * 2);
  if (rot9) std::memcpy(rot9, h->exp_rot.data(), n * 36);
  if (shape_idx) std::memcpy(shape_idx, h->exp_shape.data(), n * 4);
  if (wl_idx) std::memcpy(wl_idx, h->exp_wl.data(), n * 4);
  h->exp_d.clear();
  h->exp_p.clear();
  h->exp_w.clear();
  h->exp_rot.clear();
  h->exp_face.clear();
  h->exp_shape.clear();
  h->exp_wl.clear();
  return HB_OK;
}

int hb_export_root_masks(HbEngine* h, uint64_t cap, uint64_t* masks, uint64_t* count) {
  if (h == nullptr || count == nullptr) return HB_ERR_INVALID_ARG;
  *count = h->exp_mask.size();
  if (masks == nullptr) return HB_OK;
  if (cap < h->exp_mask.size()) return fail(h, HB_ERR_CAPACITY, "export_root_masks: caller buffer too small");
  if (!h->exp_mask.empty()) std::memcpy(masks, h->exp_mask.data(), h->exp_mask.size() * sizeof(uint64_t));
  h->exp_mask.clear();
  return HB_OK;
}

int hb_set_option(HbEngine* h, const char* key, int64_t value) {
  if (h == nullptr || key == nullptr) return HB_ERR_INVALID_ARG;
  const std::string k(key);
  if (k == "tile_rays") {
    if (value < 1024 || value > (1ll << 30)) return fail(h, HB_ERR_INVALID_ARG, "tile_rays out of range");
    h->tile_rays = static_cast<uint64_t>(value);
  } else if (k == "profile") {
    h->profile = value != 0;
  } else if (k == "blocks_per_sm") {
    if (value < 1 || value > 128) return fail(h, HB_ERR_INVALID_ARG, "blocks_per_sm out of range");
    h->blocks_per_sm = static_cast<int>(value);
    h->blocks_per_sm_override = static_cast<int>(value);
  } else if (k == "prism_fast_path") {
    h->p4_enable = value != 0;
  } else if (k == "pixel_cache") {
    h->pixel_cache = value != 0;
  } else if (k == "fused_bounce") {
    h->fused_bounce = value != 0;
  } else if (k == "fused_gen") {
    h->fused_gen = value != 0;
  } else if (k == "filter_hit_bound") {
    h->filter_hit_bound = value != 0;
  } else if (k == "fork_cap") {   // fork-ray slots per tile (0 = automatic: max(4096, n / 64)); tests force overflows
    if (value < 0 || value > (1ll << 30)) return fail(h, HB_ERR_INVALID_ARG, "fork_cap out of range");
    h->fork_cap_override = static_cast<uint64_t>(value);
  } else if (k == "cont_cap") {   // continuation-pool capacity (0 = automatic: n (max_hits + 1) + 4096)
    if (value < 0) return fail(h, HB_ERR_INVALID_ARG, "cont_cap out of range");
    h->cont_cap_override = static_cast<uint64_t>(value);
  } else if (k == "fold_rays") {
    if (value < 1024) return fail(h, HB_ERR_INVALID_ARG, "fold_rays out of range");
    h->fold_rays = static_cast<uint64_t>(value);
  } else if (k == "gen_base") {
    h->gen_base = static_cast<uint64_t>(value);
  } else if (k == "stream_base") {  // restart all monotone stream counters (tests: reproducible replays)
    h->gen_base = h->gate_base = h->transit_base = static_cast<uint64_t>(value);
    h->shuffle_round = 0;
    for (auto& L : h->layers) std::fill(L->carry.begin(), L->carry.end(), 0.0);
  } else {
    return fail(h, HB_ERR_INVALID_ARG, "unknown option " + k);
  }
  return HB_OK;
}

namespace {

int ensure_geom_state(HbEngine* h, LayerDev& L) {
  HB_CUDA(h, L.geom_flags_dev.ensure(L.pops.size() * 2));
  if (L.geom_flags_host == nullptr) HB_CUDA(h, cudaMallocHost(&L.geom_flags_host, HB_MAX_CRYSTALS * 2 * sizeof(uint32_t)));
  if (L.shadow_event == nullptr) HB_CUDA(h, cudaEventCreateWithFlags(&L.shadow_event, cudaEventDisableTiming));
  if (L.main_event == nullptr) HB_CUDA(h, cudaEventCreateWithFlags(&L.main_event, cudaEventDisableTiming));
  if (h->geom_stream == nullptr) HB_CUDA(h, cudaStreamCreateWithFlags(&h->geom_stream, cudaStreamNonBlocking));
  return HB_OK;
}

// Second copy of the layer's shape tables (allocated on first use, filled device-to-device from the live one:
// deterministic populations never change, stochastic ones are overwritten by every redraw).
int ensure_shadow(HbEngine* h, LayerDev& L) {
  ShapeBufs& src = L.sb[L.cur];
  ShapeBufs& dst = L.sb[L.cur ^ 1];
  if (dst.allocated) return HB_OK;
  HB_CUDA(h, dst.planes.ensure(src.planes.n));
  HB_CUDA(h, dst.axes.ensure(src.axes.n));
  HB_CUDA(h, dst.shape_meta.ensure(src.shape_meta.n));
  HB_CUDA(h, dst.face_fn.ensure(src.face_fn.n));
  HB_CUDA(h, dst.shapes.ensure(src.shapes.n));
  HB_CUDA(h, dst.entry_faces.ensure(src.entry_faces.n));
  HB_CUDA(h, dst.geom_scalars.ensure(src.geom_scalars.n));
  const cudaMemcpyKind k = cudaMemcpyDeviceToDevice;
  HB_CUDA(h, cudaMemcpyAsync(dst.planes.p, src.planes.p, src.planes.n * sizeof(float4), k, h->stream));
  HB_CUDA(h, cudaMemcpyAsync(dst.axes.p, src.axes.p, src.axes.n * sizeof(float4), k, h->stream));
  HB_CUDA(h, cudaMemcpyAsync(dst.shape_meta.p, src.shape_meta.p, src.shape_meta.n * sizeof(uint32_t), k, h->stream));
  HB_CUDA(h, cudaMemcpyAsync(dst.face_fn.p, src.face_fn.p, src.face_fn.n, k, h->stream));
  HB_CUDA(h, cudaMemcpyAsync(dst.shapes.p, src.shapes.p, src.shapes.n * sizeof(HbCrystalTables), k, h->stream));
  HB_CUDA(h, cudaMemcpyAsync(dst.entry_faces.p, src.entry_faces.p, src.entry_faces.n * sizeof(EntryFaces), k, h->stream));
  HB_CUDA(h, cudaMemcpyAsync(dst.geom_scalars.p, src.geom_scalars.p, src.geom_scalars.n * sizeof(float), k, h->stream));
  dst.pop_p4 = src.pop_p4;
  dst.pop_shapes = src.pop_shapes;
  dst.allocated = true;
  return HB_OK;
}

// Enqueue the redraw of one population's pool into buffer `buf` of the layer (no synchronisation).
int launch_resample(HbEngine* h, LayerDev& L, uint32_t population, int buf, const HbCrystalDesc& crystal, uint32_t seed,
                    uint32_t draw_base, cudaStream_t stream) {
  const PopHost& ph = L.pops[population];
  ShapeBufs& B = L.sb[buf];
  HB_CUDA(h, cudaMemsetAsync(L.geom_flags_dev.p + population * 2, 0, 2 * sizeof(uint32_t), stream));
  ShapeGenParams sp{};
  sp.desc = crystal;
  sp.a1 = crystal.kind == 1u ? hb_pyramid_slope(crystal.wedge_upper_deg) : -1.0;
  sp.a2 = crystal.kind == 1u ? hb_pyramid_slope(crystal.wedge_lower_deg) : -1.0;
  sp.seed = seed ^ kNonceGeom;
  sp.draw_base = draw_base;
  sp.count = ph.shape_cnt;
  sp.pop = population;
  sp.shapes = B.shapes.p + ph.shape_base;
  sp.planes = B.planes.p + static_cast<size_t>(ph.shape_base) * HB_MAX_FACES;
  sp.axes = B.axes.p + static_cast<size_t>(ph.shape_base) * HB_MAX_FACES * 2;
  sp.meta = B.shape_meta.p + ph.shape_base;
  sp.fn = B.face_fn.p + static_cast<size_t>(ph.shape_base) * HB_MAX_FACES;
  sp.ef = B.entry_faces.p + ph.shape_base;
  sp.scalars = B.geom_scalars.p + static_cast<size_t>(ph.shape_base) * 10;
  sp.flags = L.geom_flags_dev.p + population * 2;
  resample_shapes_kernel<<<(sp.count + 63u) / 64u, 64, 0, stream>>>(sp);
  h->ctr.kernel_launches++;
  HB_CUDA(h, cudaGetLastError());
  return HB_OK;
}

// Geometry clock, one session ahead. Called by hb_begin_session for every layer with hb_auto_resample
// populations: (1) the pool that was enqueued during the PREVIOUS BeginSession -- i.e. before that session's
// kernels, so it finished long ago and waiting for its event does not drain the stream -- becomes the live
// one; (2) the redraw of the next pool is enqueued into the buffer that just became free. The host learns the
// P4 flag of the new live pool from a pinned copy of the builder's counters. Only the very first session
// waits for its own redraw.
int prefetch_geometry(HbEngine* h, LayerDev& L) {
  int rc = ensure_geom_state(h, L);
  if (rc != HB_OK) return rc;
  rc = ensure_shadow(h, L);
  if (rc != HB_OK) return rc;
  // The redraw runs on its own stream, concurrently with the trace kernels of the session that is about to
  // start: it only has to wait for what is ALREADY enqueued on the trace stream (the previous session, which
  // still reads the buffer that is about to be overwritten).
  auto enqueue_shadow = [&]() -> int {
    const int buf = L.cur ^ 1;
    HB_CUDA(h, cudaEventRecord(L.main_event, h->stream));
    HB_CUDA(h, cudaStreamWaitEvent(h->geom_stream, L.main_event, 0));
    for (uint32_t pi = 0; pi < L.pops.size(); pi++) {
      PopHost& ph = L.pops[pi];
      if (!ph.auto_geom) continue;
      int r = launch_resample(h, L, pi, buf, ph.auto_desc, ph.auto_seed, ph.auto_draws, h->geom_stream);
      if (r != HB_OK) return r;
      ph.auto_draws += ph.shape_cnt;
    }
    HB_CUDA(h, cudaMemcpyAsync(L.geom_flags_host, L.geom_flags_dev.p, L.pops.size() * 2 * sizeof(uint32_t),
                               cudaMemcpyDeviceToHost, h->geom_stream));
    HB_CUDA(h, cudaEventRecord(L.shadow_event, h->geom_stream));
    L.shadow_pending = true;
    return HB_OK;
  };
  if (!L.shadow_pending) {
    rc = enqueue_shadow();
    if (rc != HB_OK) return rc;
  }
  HB_CUDA(h, cudaEventSynchronize(L.shadow_event));
  HB_CUDA(h, cudaStreamWaitEvent(h->stream, L.shadow_event, 0));
  ShapeBufs& fresh = L.sb[L.cur ^ 1];
  for (uint32_t pi = 0; pi < L.pops.size(); pi++) {
    if (!L.pops[pi].auto_geom) continue;
    fresh.pop_p4[pi] = L.pops[pi].shape_cnt - std::min(L.pops[pi].shape_cnt, L.geom_flags_host[pi * 2]);
    L.pops[pi].device_pool = true;
    h->geom_rejected += L.geom_flags_host[pi * 2 + 1];
  }
  L.cur ^= 1;
  L.shadow_pending = false;
  return enqueue_shadow();
}

}  // namespace

int hb_resample_shapes(HbEngine* h, uint32_t layer, uint32_t population, const HbCrystalDesc* crystal, uint32_t seed,
                       uint32_t draw_base, uint32_t* rejected) {
  if (h == nullptr || crystal == nullptr) return HB_ERR_INVALID_ARG;
  if (!h->have_scene) return fail(h, HB_ERR_STATE, "resample_shapes before hb_set_scene");
  if (h->in_session) return fail(h, HB_ERR_STATE, "resample_shapes inside a session");
  if (layer >= h->layers.size() || population >= h->layers[layer]->pops.size())
    return fail(h, HB_ERR_INVALID_ARG, "resample_shapes: no such layer / population");
  if (crystal->kind > 1u) return fail(h, HB_ERR_INVALID_ARG, "resample_shapes: crystal kind must be prism (0) or pyramid (1)");
  cudaSetDevice(h->device);
  LayerDev& L = *h->layers[layer];
  if (L.pops[population].auto_geom) return fail(h, HB_ERR_STATE, "resample_shapes: population is on the automatic geometry clock");
  int rc = ensure_geom_state(h, L);
  if (rc != HB_OK) return rc;
  if (L.shadow_pending) {  // a prefetched pool of this layer is in flight: it would carry a stale copy of this population
    HB_CUDA(h, cudaEventSynchronize(L.shadow_event));
    L.shadow_pending = false;
  }
  L.sb[L.cur ^ 1].allocated = false;  // re-copied from the live tables before its next use
  rc = launch_resample(h, L, population, L.cur, *crystal, seed, draw_base, h->stream);
  if (rc != HB_OK) return rc;
  uint32_t flags[2] = { 0, 0 };
  HB_CUDA(h, cudaMemcpyAsync(flags, L.geom_flags_dev.p + population * 2, sizeof(flags), cudaMemcpyDeviceToHost, h->stream));
  HB_CUDA(h, cudaStreamSynchronize(h->stream));
  L.b().pop_p4[population] = L.pops[population].shape_cnt - std::min(L.pops[population].shape_cnt, flags[0]);
  L.pops[population].device_pool = true;
  if (rejected != nullptr) *rejected = flags[1];
  return HB_OK;
}

int hb_auto_resample(HbEngine* h, uint32_t layer, uint32_t population, const HbCrystalDesc* crystal, uint32_t seed,
                     uint32_t draw_base) {
  if (h == nullptr) return HB_ERR_INVALID_ARG;
  if (!h->have_scene) return fail(h, HB_ERR_STATE, "auto_resample before hb_set_scene");
  if (h->in_session) return fail(h, HB_ERR_STATE, "auto_resample inside a session");
  if (layer >= h->layers.size() || population >= h->layers[layer]->pops.size())
    return fail(h, HB_ERR_INVALID_ARG, "auto_resample: no such layer / population");
  cudaSetDevice(h->device);
  LayerDev& L = *h->layers[layer];
  PopHost& ph = L.pops[population];
  if (crystal == nullptr) {  // switch the clock off: the live pool stays
    ph.auto_geom = false;
  } else {
    if (crystal->kind > 1u) return fail(h, HB_ERR_INVALID_ARG, "auto_resample: crystal kind must be prism (0) or pyramid (1)");
    ph.auto_geom = true;
    ph.auto_desc = *crystal;
    ph.auto_seed = seed;
    ph.auto_draws = draw_base;
  }
  if (L.shadow_pending) {  // a pool drawn under the old settings is in flight: drop it
    HB_CUDA(h, cudaEventSynchronize(L.shadow_event));
    L.shadow_pending = false;
  }
  L.any_auto = false;
  for (const PopHost& q : L.pops) L.any_auto = L.any_auto || q.auto_geom;
  return HB_OK;
}

int hb_export_shapes(HbEngine* h, uint32_t layer, uint32_t population, uint32_t cap, HbCrystalTables* tables,
                     float* scalars10, uint32_t* count) {
  if (h == nullptr || count == nullptr) return HB_ERR_INVALID_ARG;
  if (!h->have_scene || layer >= h->layers.size() || population >= h->layers[layer]->pops.size())
    return fail(h, HB_ERR_INVALID_ARG, "export_shapes: no such layer / population");
  cudaSetDevice(h->device);
  LayerDev& L = *h->layers[layer];
  const PopHost& ph = L.pops[population];
  *count = ph.shape_cnt;
  if (tables == nullptr) return HB_OK;
  if (cap < ph.shape_cnt) return fail(h, HB_ERR_CAPACITY, "export_shapes: caller buffer too small");
  HB_CUDA(h, cudaStreamSynchronize(h->stream));
  HB_CUDA(h, cudaMemcpy(tables, L.b().shapes.p + ph.shape_base, ph.shape_cnt * sizeof(HbCrystalTables), cudaMemcpyDeviceToHost));
  if (scalars10 != nullptr) {
    if (ph.device_pool) {
      HB_CUDA(h, cudaMemcpy(scalars10, L.b().geom_scalars.p + static_cast<size_t>(ph.shape_base) * 10,
                            static_cast<size_t>(ph.shape_cnt) * 10 * sizeof(float), cudaMemcpyDeviceToHost));
    } else {
      std::memset(scalars10, 0, static_cast<size_t>(ph.shape_cnt) * 10 * sizeof(float));
    }
  }
  return HB_OK;
}

int hb_get_counters(HbEngine* h, HbCounters* out) {
  if (h == nullptr || out == nullptr) return HB_ERR_INVALID_ARG;
  *out = h->ctr;
  return HB_OK;
}

int hb_synchronize(HbEngine* h) {
  if (h == nullptr) return HB_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  HB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->profile) flush_events(h);
  if (h->counters.p) {
    HB_CUDA(h, cudaMemcpy(h->err_host, h->counters.p + 4, 4, cudaMemcpyDeviceToHost));
    return check_device_error(h);
  }
  return HB_OK;
}

int hb_selftest_arith(HbEngine* h, uint32_t mode, uint64_t n, uint32_t seed, uint64_t* out4) {
  if (h == nullptr || out4 == nullptr || mode > 2u) return HB_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  unsigned long long* bad = nullptr;
  HB_CUDA(h, cudaMalloc(&bad, 4 * sizeof(unsigned long long)));
  cudaMemsetAsync(bad, 0, 4 * sizeof(unsigned long long), h->stream);
  hb::selftest_arith_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(n, seed, mode, bad);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out4, bad, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(bad);
  HB_CUDA(h, e);
  return HB_OK;
}

int hb_image_device_ptr(HbEngine* h, void** ptr, uint64_t* float_count) {
  if (h == nullptr || ptr == nullptr || float_count == nullptr) return HB_ERR_INVALID_ARG;
  if (!h->have_render) return fail(h, HB_ERR_STATE, "no render set");
  cudaSetDevice(h->device);
  fold_arena(h);
  *ptr = h->master.p;
  *float_count = static_cast<uint64_t>(h->arena_pix) * 4;
  return HB_OK;
}

void* hb_stream(HbEngine* h) { return h ? static_cast<void*>(h->stream) : nullptr; }

int hb_comm_unique_id(void* out) {
#ifdef HB_WITH_NCCL
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return HB_ERR_COMM;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  std::memcpy(out, &id, 128);
  return HB_OK;
#else
  (void)out;
  return HB_ERR_UNSUPPORTED;
#endif
}

int hb_comm_init(HbEngine* h, const void* id128, int rank, int nranks) {
  if (h == nullptr || id128 == nullptr) return HB_ERR_INVALID_ARG;
#ifdef HB_WITH_NCCL
  cudaSetDevice(h->device);
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  if (ncclCommInitRank(&h->comm, nranks, id, rank) != ncclSuccess) return fail(h, HB_ERR_COMM, "ncclCommInitRank failed");
  h->rank = rank;
  h->nranks = nranks;
  return HB_OK;
#else
  (void)rank;
  (void)nranks;
  return fail(h, HB_ERR_UNSUPPORTED, "library built without NCCL");
#endif
}

namespace {

#ifdef HB_WITH_NCCL
// Frame-end reduction over the NCCL communicator; root < 0: all-reduce. fp32 (X, Y, Z, landed) per pixel of the
// whole arena + the colour-class lanes.
int reduce_over_comm(HbEngine* h, int root) {
  if (h->comm == nullptr) return fail(h, HB_ERR_STATE, "hb_comm_init not called");
  cudaSetDevice(h->device);
  if (!h->have_render) return fail(h, HB_ERR_STATE, "no render set");
  if (root >= h->nranks) return fail(h, HB_ERR_INVALID_ARG, "reduce: root rank out of range");
  if (root < 0 && h->allreduced) {
    return fail(h, HB_ERR_STATE, "hb_allreduce_image called twice on the same accumulation: every rank already holds the sum "
                                 "(trace or drain first, or use hb_reduce_image)");
  }
  fold_arena(h);
  HB_CUDA(h, h->reduce_stage.ensure(h->arena_pix));
  pack_master_kernel<<<grid_for(h, h->arena_pix), 256, 0, h->stream>>>(h->master.p, h->reduce_stage.p, h->arena_pix);
  h->ctr.kernel_launches++;
  const size_t cnt = static_cast<size_t>(h->arena_pix) * 4;
  const bool lanes = h->classes.class_cnt != 0 && h->lanes_floats != 0;
  ncclResult_t rc = ncclGroupStart();
  if (rc == ncclSuccess) {
    rc = root < 0 ? ncclAllReduce(h->reduce_stage.p, h->reduce_stage.p, cnt, ncclFloat, ncclSum, h->comm, h->stream)
                  : ncclReduce(h->reduce_stage.p, h->reduce_stage.p, cnt, ncclFloat, ncclSum, root, h->comm, h->stream);
  }
  if (rc == ncclSuccess && lanes) {
    rc = root < 0 ? ncclAllReduce(h->lanes.p, h->lanes.p, h->lanes_floats, ncclFloat, ncclSum, h->comm, h->stream)
                  : ncclReduce(h->lanes.p, h->lanes.p, h->lanes_floats, ncclFloat, ncclSum, root, h->comm, h->stream);
  }
  if (rc == ncclSuccess) rc = ncclGroupEnd();
  if (rc != ncclSuccess) return fail(h, HB_ERR_COMM, std::string("NCCL reduce failed: ") + ncclGetErrorString(rc));
  if (root < 0 || root == h->rank) {
    unpack_master_kernel<<<grid_for(h, h->arena_pix), 256, 0, h->stream>>>(h->reduce_stage.p, h->master.p, h->arena_pix);
    h->ctr.kernel_launches++;
  } else {  // this rank's contribution now lives on the root
    HB_CUDA(h, cudaMemsetAsync(h->master.p, 0, static_cast<size_t>(h->arena_pix) * sizeof(double4), h->stream));
    if (lanes) HB_CUDA(h, cudaMemsetAsync(h->lanes.p, 0, h->lanes_floats * sizeof(float), h->stream));
  }
  h->allreduced = root < 0;
  return HB_OK;
}
#endif

}  // namespace

int hb_allreduce_image(HbEngine* h) {
  if (h == nullptr) return HB_ERR_INVALID_ARG;
#ifdef HB_WITH_NCCL
  return reduce_over_comm(h, -1);
#else
  return fail(h, HB_ERR_UNSUPPORTED, "library built without NCCL");
#endif
}

int hb_reduce_image(HbEngine* h, int root) {
  if (h == nullptr || root < 0) return HB_ERR_INVALID_ARG;
#ifdef HB_WITH_NCCL
  return reduce_over_comm(h, root);
#else
  return fail(h, HB_ERR_UNSUPPORTED, "library built without NCCL");
#endif
}

int hb_merge_from_peer(HbEngine* dst, HbEngine* src) {
  if (dst == nullptr || src == nullptr || dst == src) return HB_ERR_INVALID_ARG;
  // colour lanes exist only while the scene has colour classes (a stale allocation of an earlier scene does not count)
  const uint64_t dst_lanes = dst->classes.class_cnt != 0 ? dst->lanes_floats : 0;
  const uint64_t src_lanes = src->classes.class_cnt != 0 ? src->lanes_floats : 0;
  if (!dst->have_render || !src->have_render || dst->arena_pix != src->arena_pix || dst_lanes != src_lanes)
    return fail(dst, HB_ERR_STATE, "merge_from_peer: the two engines do not hold the same renders");
  // Legal wherever hb_readback_xyz is, i.e. also inside a session: the reference driver drains on its third clock from
  // within SimulateOneWavelengthWithBackend, before EndSession (simulator.cpp:1610-1640). The merge is ordered on the
  // engines' streams after everything traced so far.
  if (dst->device != src->device) {
    int can = 0;
    cudaDeviceCanAccessPeer(&can, dst->device, src->device);
    if (!can) return fail(dst, HB_ERR_UNSUPPORTED, "merge_from_peer: no peer access between the two devices");
    cudaSetDevice(dst->device);
    cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
      dst->error = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e);
      return HB_ERR_CUDA;
    }
    (void)cudaGetLastError();
  }
  // src: fold, then signal; dst: wait, add the peer's master (P2P loads), signal; src: wait, zero.
  cudaSetDevice(src->device);
  fold_arena(src);
  if (src->merge_event == nullptr) HB_CUDA(src, cudaEventCreateWithFlags(&src->merge_event, cudaEventDisableTiming));
  HB_CUDA(src, cudaEventRecord(src->merge_event, src->stream));
  cudaSetDevice(dst->device);
  fold_arena(dst);
  if (dst->merge_event == nullptr) HB_CUDA(dst, cudaEventCreateWithFlags(&dst->merge_event, cudaEventDisableTiming));
  HB_CUDA(dst, cudaStreamWaitEvent(dst->stream, src->merge_event, 0));
  merge_peer_kernel<<<grid_for(dst, dst->arena_pix), 256, 0, dst->stream>>>(dst->master.p, src->master.p, dst->arena_pix,
                                                                           dst->lanes.p, src->lanes.p, dst_lanes);
  dst->ctr.kernel_launches++;
  HB_CUDA(dst, cudaGetLastError());
  HB_CUDA(dst, cudaEventRecord(dst->merge_event, dst->stream));
  cudaSetDevice(src->device);
  HB_CUDA(src, cudaStreamWaitEvent(src->stream, dst->merge_event, 0));
  HB_CUDA(src, cudaMemsetAsync(src->master.p, 0, static_cast<size_t>(src->arena_pix) * sizeof(double4), src->stream));
  if (src_lanes != 0) HB_CUDA(src, cudaMemsetAsync(src->lanes.p, 0, src_lanes * sizeof(float), src->stream));
  // the source's device error word travels with its image
  HB_CUDA(src, cudaMemcpyAsync(src->err_host, src->counters.p + 4, sizeof(uint32_t), cudaMemcpyDeviceToHost, src->stream));
  cudaSetDevice(dst->device);
  return HB_OK;
}

}  // extern "C"
