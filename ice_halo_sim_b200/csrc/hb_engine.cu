// hb_engine.cu — wavefront trace engine for NVIDIA B200 (sm_100a): kernels, session state machine and
// the compute half of the C ABI declared in include/halotrace_b200.h.
//
// Pipeline of one scattering layer over one tile of rays (DESIGN.md "kernels"):
//   gen_roots / transit_roots      root state  P{p.xyz,bits} D{d.xyz,w} Q{orientation quaternion}  (SoA, float4)
//   for hit h = 0 .. H-1:
//     optics   : Fresnel split at the face the ray sits on; the child that leaves the crystal is
//                rotated to world space, filtered, gated, projected and reduced into the image;
//                the child that stays inside overwrites D
//     intersect: slab exit-face search for the inside child; overwrites P (new point + hit face)
// Reference behaviour restated: simulator.cpp:1308-1336 (hit loop, max_hits counts the entry
// interaction), optics.cpp:18-177, simulator.cpp:665-762 (emit gate), scatter_accum.hpp:47-110.
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

// ------------------------------------------------------------------------------------------------
// Kernel parameter blocks
// ------------------------------------------------------------------------------------------------
struct LayerTables {            // device pointers, one scattering layer
  const float4* planes;         // [shape_cnt][HB_MAX_FACES]
  const float4* axes;           // [shape_cnt][HB_MAX_FACES][2] paired-plane axis table (see slab_exit)
  const uint32_t* shape_meta;   // [shape_cnt] face_cnt | population << 8 | axis_cnt << 16
  const uint8_t* face_fn;       // [shape_cnt][HB_MAX_FACES]
  const HbCrystalTables* shapes;  // [shape_cnt] full tables (entry sampling)
  const HbFilterDesc* filters;  // [pop_cnt]
  const uint32_t* pop_crystal_id;  // [pop_cnt]
  uint32_t shape_cnt;
  uint32_t pop_cnt;
  uint32_t any_filter;
  const HbColorGroup* color_groups;   // [pop_cnt][HB_MAX_COLOR_GROUPS] (nullptr: no colour predicates in this layer)
  const uint32_t* color_group_cnt;    // [pop_cnt]
};

enum : uint32_t {
  kFlagPath = 1u,     // record the face sequence of every ray (filters / exit records)
  kFlagRecord = 2u,   // materialise HbExitRecord for every outgoing ray
  kFlagAccum = 4u,    // fused projection + image reduction
  kFlagGate = 8u,     // layer prob > 0: draw the continue/outgoing gate, append continuations
  kFlagStats = 16u,   // LayerStats (exit count, weight sum)
  kFlagPixelCache = 32u,  // per-CTA shared-memory pixel cache in the optics kernel
};

// Additional projections of the same trace (N renderers per trace). Render 0 lives in the
// kernel parameters; the others are read from this device table on the (rare) emit path.
struct ExtraRender {
  HbProjParams proj;
  uint32_t pixel_offset;        // first pixel of this render in the image arena
};

struct TraceParams {
  float4* P;
  float4* D;
  float4* Q;
  uint8_t* path;                // [max_hits][cap] compact face ids (kFlagPath)
  uint32_t* fork_root;          // [fork_cap] layer-root index of fork rays
  uint32_t* fork_code;          // [fork_cap] branch code of fork rays
  uint32_t* fork_count;         // rays appended behind the main slots (near-edge double continuation)
  uint32_t* fork_snapshot;      // fork_count as of the last intersect kernel
  uint32_t n_main, cap, fork_cap;
  uint32_t root_base;           // layer-root index of slot 0 of this tile
  LayerTables lt;
  const HbWlEntry* wl;
  uint32_t wl_cnt;
  float4* image;                // image arena: render r occupies [off_r, off_r + W_r*H_r), (X, Y, Z, landed)
  HbProjParams proj;            // render 0 (arena offset 0)
  const ExtraRender* extra;     // renders 1..extra_cnt (device)
  uint32_t extra_cnt;
  // raypath colour (kernels instantiated with MULTI only)
  uint32_t color_on;            // scene has colour classes
  uint64_t* M;                  // [cap + fork_cap] component mask carried in from earlier layers (nullptr on layer 0)
  uint64_t* cont_mask;          // continuation records: component mask
  float* lane;                  // [class_cnt][lane_stride] per-class Y lanes of render 0
  uint32_t lane_stride;
  HbColorClasses classes;
  uint32_t hit, max_hits, layer_idx, flags;
  float prob;
  uint32_t gate_seed;           // session seed ^ gate nonce
  uint32_t gate_base_lo, gate_base_hi;  // global gate index of layer-root 0
  float4* cont_dw;              // continuation records: world dir + weight
  uint32_t* cont_meta;          // wl index | population << 8
  uint32_t* cont_root;          // layer-root index of the parent (record mode)
  uint32_t* cont_count;
  uint32_t cont_cap;
  HbExitRecord* exits;
  uint32_t* exit_root;
  uint32_t* exit_count;
  uint32_t exit_cap;
  unsigned long long* stat_exit_count;
  double* stat_w_sum;
  uint32_t* error_flag;
};

struct GenParams {
  float4* P;
  float4* D;
  float4* Q;
  uint8_t* path;
  uint32_t slot0;               // first tile slot written by this launch
  uint32_t count;
  uint32_t cap;
  uint32_t idx_lo, idx_hi;      // 64-bit stream index of slot0's ray
  uint32_t seed;                // session seed ^ stream nonce
  AxisParams axis;
  const float* lut;             // [3][HB_LUT_NODES] device
  const HbCrystalTables* shapes;  // this population's pool (device)
  uint32_t shape_base, shape_cnt;
  const HbWlEntry* wl;
  uint32_t wl_cnt;
  float sun_lon, sun_lat, sun_half;
  float sun_c_cap, sun_c_lon, sun_s_lon, sun_c_lat, sun_s_lat;  // per-launch constants of sample_sph_cap
  uint32_t flags;
  // transit only
  const float4* cont_dw;
  const uint32_t* cont_meta;
  const uint64_t* cont_mask;    // component masks of the continuations (raypath colour) or nullptr
  uint64_t* M;                  // per-slot carried mask of the next layer
  uint32_t cont_n;              // size of the permuted continuation pool
  uint32_t cont_first;          // pool position of slot0
  uint32_t shuffle_seed;
  uint32_t shuffle;
};

// ------------------------------------------------------------------------------------------------
// Emission (CollectData branch 1, simulator.cpp:678-730 + ScatterOutgoingToXyz)
// ------------------------------------------------------------------------------------------------
struct Tally {
  unsigned long long exits = 0;
  double w_sum = 0.0;
  uint32_t* cache_keys = nullptr;  // per-CTA pixel cache (see PixelCache below); nullptr = reduce straight to L2
  float* cache_vals = nullptr;
};

// Per-CTA pixel cache. Halo images are extremely peaked (the undeviated light through parallel faces lands
// on the ~35 pixels of the sun disk: ~40 % of all exits), and same-address reductions serialise in the L2
// atomic unit while the fp32 accumulator of such a pixel absorbs small addends. Each CTA therefore keeps a
// direct-mapped table of kCacheSlots pixels in shared memory (first come, first claimed): contributions to a
// cached pixel are summed in shared memory and reduced into the global image once, when the CTA retires.
// Everything else goes straight to the L2 with one red.global.add.v4.f32.
constexpr uint32_t kCacheSlots = 512;
constexpr uint32_t kCacheEmpty = 0xFFFFFFFFu;
constexpr size_t kCacheBytes = kCacheSlots * (sizeof(uint32_t) + 4 * sizeof(float));

HB_DEV void accumulate_pixel(const TraceParams& tp, const Tally& tally, uint32_t pix, float x, float y, float z, float lw) {
  if (tally.cache_keys != nullptr) {
    const uint32_t slot = (pix * 2654435761u) >> 23;  // top 9 bits
    uint32_t k = tally.cache_keys[slot];
    if (k == kCacheEmpty) {
      k = atomicCAS(&tally.cache_keys[slot], kCacheEmpty, pix);
      if (k == kCacheEmpty) k = pix;
    }
    if (k == pix) {
      float* v = tally.cache_vals + slot * 4u;
      atomicAdd(v + 0, x);
      atomicAdd(v + 1, y);
      atomicAdd(v + 2, z);
      if (lw != 0.0f) atomicAdd(v + 3, lw);
      return;
    }
  }
  red_add_f4(tp.image + pix, x, y, z, lw);
}

HB_DEV void cache_init(Tally& tally, unsigned char* smem) {
  tally.cache_keys = reinterpret_cast<uint32_t*>(smem);
  tally.cache_vals = reinterpret_cast<float*>(smem + kCacheSlots * sizeof(uint32_t));
  for (uint32_t i = threadIdx.x; i < kCacheSlots; i += blockDim.x) {
    tally.cache_keys[i] = kCacheEmpty;
    tally.cache_vals[4u * i + 0] = 0.0f;
    tally.cache_vals[4u * i + 1] = 0.0f;
    tally.cache_vals[4u * i + 2] = 0.0f;
    tally.cache_vals[4u * i + 3] = 0.0f;
  }
}

HB_DEV void cache_flush(const TraceParams& tp, const Tally& tally) {
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < kCacheSlots; i += blockDim.x) {
    const uint32_t k = tally.cache_keys[i];
    if (k != kCacheEmpty) {
      const float* v = tally.cache_vals + 4u * i;
      red_add_f4(tp.image + k, v[0], v[1], v[2], v[3]);
    }
  }
}

template <bool GENERAL, bool MULTI, typename TablesT>
HB_DEV void emit_exit(const TraceParams& tp, uint32_t slot, uint32_t bits, float4 q, float lx, float ly, float lz,
                      float w, uint32_t role, const TablesT& tb, Tally& tally) {
  const Rot r = rot_from_quat(q);
  float wx, wy, wz;
  rot_apply(r, lx, ly, lz, wx, wy, wz);
  const uint32_t wl_i = bits_wl(bits);
  bool to_next_layer = false;
  uint64_t mask = 0ull;  // component mask (raypath colour)
  if (GENERAL) {
    const uint32_t shape = bits_shape(bits);
    const uint32_t pop = (tp.lt.shape_meta[shape] >> 8) & 255u;
    uint32_t root, code;
    if (slot < tp.n_main) {
      root = tp.root_base + slot;
      code = 0u;
    } else {
      root = tp.fork_root[slot - tp.n_main];
      code = tp.fork_code[slot - tp.n_main];
    }
    uint8_t fn_path[HB_MAX_HITS];
    const uint32_t len = tp.hit + 1u;
    if (tp.flags & kFlagPath) {
      for (uint32_t k = 0; k < len; k++)
        fn_path[k] = static_cast<uint8_t>(tb.face_fn(shape, tp.path[static_cast<size_t>(k) * tp.cap + slot]));
      if (tp.lt.any_filter) {
        const HbFilterDesc& f = tp.lt.filters[pop];
        if (f.kind != 0u) {
          const float dir[3] = { wx, wy, wz };
          if (!filter_check(f, fn_path, len, dir, tp.lt.pop_crystal_id[pop])) return;  // filter-fail terminates
        }
      }
    }
    if constexpr (MULTI) {
      // Non-destructive colour pass on a filter-admitted exit (simulator.cpp:688-712): every matching
      // predicate ORs its component bit into the mask carried from earlier layers.
      if (tp.color_on) {
        mask = tp.M != nullptr ? tp.M[slot] : 0ull;
        if (tp.lt.color_groups != nullptr) {
          const uint32_t gcnt = tp.lt.color_group_cnt[pop];
          const float dir[3] = { wx, wy, wz };
          for (uint32_t g = 0; g < gcnt; g++) {
            const HbColorGroup& cg = tp.lt.color_groups[pop * HB_MAX_COLOR_GROUPS + g];
            for (uint32_t k = 0; k < cg.filter.term_cnt; k++) {
              if (cg.bit[k] < 64u &&
                  filter_match_simple(cg.filter, cg.filter.terms[k][0], fn_path, len, dir, tp.lt.pop_crystal_id[pop]))
                mask |= 1ull << cg.bit[k];
            }
          }
        }
      }
    }
    if (tp.flags & kFlagGate) {
      // one uniform per filter-passing exit (simulator.cpp:719-723), keyed by (layer root, hit, role):
      // role 0 = the child on the far side of the face, role 1 = the child that stays on the near side
      const uint32_t glo = tp.gate_base_lo + root;
      const uint32_t ghi = tp.gate_base_hi + (glo < tp.gate_base_lo ? 1u : 0u);
      uint32_t seed = seed_with_high(tp.gate_seed, ghi);
      if (code != 0u) seed ^= pcg_hash(code);
      to_next_layer = draw(seed, glo, tp.hit * 2u + role) < tp.prob;
    }
    if (tp.flags & kFlagStats) {
      tally.exits++;
      tally.w_sum += static_cast<double>(w);
    }
    if (to_next_layer) {
      // warp-aggregated append into the continuation pool (ballot + one atomic per warp)
      const uint32_t active = __activemask();
      const uint32_t lane = threadIdx.x & 31u;
      const uint32_t leader = __ffs(active) - 1u;
      uint32_t base = 0u;
      if (lane == leader) base = atomicAdd(tp.cont_count, static_cast<uint32_t>(__popc(active)));
      base = __shfl_sync(active, base, leader);
      const uint32_t dst = base + __popc(active & ((1u << lane) - 1u));
      if (dst < tp.cont_cap) {
        tp.cont_dw[dst] = make_float4(wx, wy, wz, w);
        tp.cont_meta[dst] = wl_i | (pop << 8);
        if (tp.cont_root != nullptr) tp.cont_root[dst] = root;
        if constexpr (MULTI) {
          if (tp.cont_mask != nullptr) tp.cont_mask[dst] = mask;
        }
      } else {
        *tp.error_flag = 1u;
      }
      return;
    }
    if (tp.flags & kFlagRecord) {
      const uint32_t dst = atomicAdd(tp.exit_count, 1u);
      if (dst < tp.exit_cap) {
        HbExitRecord& e = tp.exits[dst];
        e.dir[0] = wx;
        e.dir[1] = wy;
        e.dir[2] = wz;
        e.weight = w;
        e.path_len = static_cast<uint8_t>(len);
        for (uint32_t k = 0; k < HB_MAX_HITS; k++) e.path[k] = k < len ? fn_path[k] : 0;
        e.pad0_ = 0;
        e.crystal_id = static_cast<uint16_t>(pop);
        e.ms_layer_idx = static_cast<uint8_t>(tp.layer_idx);
        e.wl_idx = static_cast<uint8_t>(wl_i);
        e.pad1_[0] = e.pad1_[1] = 0;
        e.component_mask = mask;
        tp.exit_root[dst] = root;
      } else {
        *tp.error_flag = 2u;
      }
    }
    if (!(tp.flags & kFlagAccum)) return;
  }
  const PixelHits h = project_exit(tp.proj, wx, wy, wz);
  const HbWlEntry we = tp.wl[wl_i];
#pragma unroll
  for (int k = 0; k < 2; k++) {
    if (k < h.count) {
      const int px = h.px[k], py = h.py[k];
      if (px >= 0 && px < tp.proj.img_w && py >= 0 && py < tp.proj.img_h) {
        const uint32_t pix = static_cast<uint32_t>(py) * static_cast<uint32_t>(tp.proj.img_w) + static_cast<uint32_t>(px);
        accumulate_pixel(tp, tally, pix, mul(we.cmf_x, w), mul(we.cmf_y, w), mul(we.cmf_z, w), h.bump[k] ? w : 0.0f);
        if constexpr (MULTI) {
          // FanColorClassLanes (cuda_trace_backend.cu:538-556): Y into every satisfied class, overlap-ring hits too
          if (tp.color_on && mask != 0ull) {
            const float y = mul(we.cmf_y, w);
            for (uint32_t c = 0; c < tp.classes.class_cnt; c++) {
              const uint64_t cb = tp.classes.bits[c];
              if (cb == 0ull) continue;
              const uint64_t m = mask & cb;
              const bool ok = ((tp.classes.combine_all_mask >> c) & 1u) ? (m == cb) : (m != 0ull);
              if (ok) atomicAdd(tp.lane + static_cast<size_t>(c) * tp.lane_stride + pix, y);
            }
          }
        }
      }
    }
  }
  if constexpr (MULTI) {  // further projections of the same exit (multi-render traces run the GENERAL+MULTI kernels)
    for (uint32_t r = 0; r < tp.extra_cnt; r++) {
      const HbProjParams pr = tp.extra[r].proj;
      const uint32_t off = tp.extra[r].pixel_offset;
      const PixelHits hr = project_exit(pr, wx, wy, wz);
#pragma unroll
      for (int k = 0; k < 2; k++) {
        if (k < hr.count) {
          const int px = hr.px[k], py = hr.py[k];
          if (px >= 0 && px < pr.img_w && py >= 0 && py < pr.img_h) {
            accumulate_pixel(tp, tally, off + static_cast<uint32_t>(py) * static_cast<uint32_t>(pr.img_w) + static_cast<uint32_t>(px),
                             mul(we.cmf_x, w), mul(we.cmf_y, w), mul(we.cmf_z, w), hr.bump[k] ? w : 0.0f);
          }
        }
      }
    }
  }
}

// Shared-memory staging of the per-layer crystal tables.
// Layout in dynamic shared memory, n = shape count:
//   float4 planes[n][20] | float4 axes[n][20][2] | uint32 meta[n] | uint8 face_fn[n][20]
// Pools of more than kSmemShapes shapes do not fit and are read through the read-only L1/L2 path.
constexpr uint32_t kSmemShapes = 40;
constexpr uint32_t kShapeSmemBytes = HB_MAX_FACES * 16u + HB_MAX_FACES * 32u + 4u + HB_MAX_FACES;
__host__ __device__ inline size_t shared_tables_bytes(uint32_t shape_cnt) {
  const uint32_t n = shape_cnt <= kSmemShapes ? shape_cnt : 0u;
  return static_cast<size_t>(n) * kShapeSmemBytes + 16;
}

HB_DEV float4 lds128(uint32_t addr) {  // explicit shared-space load (LDS.128), no generic-address resolution
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

template <bool SMEM>
struct AxisRow;
template <>
struct AxisRow<true> {
  uint32_t addr;
  HB_DEV void load(uint32_t i, float4& a, float4& b) const {
    a = lds128(addr + i * 32u);
    b = lds128(addr + i * 32u + 16u);
  }
};
template <>
struct AxisRow<false> {
  const float4* p;
  HB_DEV void load(uint32_t i, float4& a, float4& b) const {
    a = __ldg(p + 2u * i);
    b = __ldg(p + 2u * i + 1u);
  }
};

template <bool SMEM>
struct Tables;
template <>
struct Tables<true> {
  uint32_t planes_addr, axes_addr, meta_addr, fn_addr;
  HB_DEV float4 plane(uint32_t shape, uint32_t face) const { return lds128(planes_addr + (shape * HB_MAX_FACES + face) * 16u); }
  HB_DEV AxisRow<true> axes(uint32_t shape) const { return AxisRow<true>{ axes_addr + shape * (HB_MAX_FACES * 32u) }; }
  HB_DEV uint32_t meta(uint32_t shape) const {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(meta_addr + shape * 4u));
    return v;
  }
  HB_DEV uint32_t face_fn(uint32_t shape, uint32_t face) const {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(fn_addr + shape * HB_MAX_FACES + face));
    return v;
  }
};
template <>
struct Tables<false> {
  const float4* planes_p;
  const float4* axes_p;
  const uint32_t* meta_p;
  const uint8_t* fn_p;
  HB_DEV float4 plane(uint32_t shape, uint32_t face) const { return __ldg(planes_p + shape * HB_MAX_FACES + face); }
  HB_DEV AxisRow<false> axes(uint32_t shape) const { return AxisRow<false>{ axes_p + shape * (HB_MAX_FACES * 2u) }; }
  HB_DEV uint32_t meta(uint32_t shape) const { return __ldg(meta_p + shape); }
  HB_DEV uint32_t face_fn(uint32_t shape, uint32_t face) const { return __ldg(fn_p + shape * HB_MAX_FACES + face); }
};

template <bool SMEM>
HB_DEV Tables<SMEM> stage_tables(const LayerTables& lt, unsigned char* smem, bool want_fn);
template <>
HB_DEV Tables<false> stage_tables<false>(const LayerTables& lt, unsigned char*, bool) {
  return Tables<false>{ lt.planes, lt.axes, lt.shape_meta, lt.face_fn };
}
template <>
HB_DEV Tables<true> stage_tables<true>(const LayerTables& lt, unsigned char* smem, bool want_fn) {
  const uint32_t n = lt.shape_cnt;
  float4* pl = reinterpret_cast<float4*>(smem);
  float4* ax = pl + n * HB_MAX_FACES;
  uint32_t* meta = reinterpret_cast<uint32_t*>(ax + n * HB_MAX_FACES * 2u);
  uint8_t* fn = reinterpret_cast<uint8_t*>(meta + n);
  for (uint32_t i = threadIdx.x; i < n * HB_MAX_FACES; i += blockDim.x) {
    pl[i] = lt.planes[i];
    ax[2u * i] = lt.axes[2u * i];
    ax[2u * i + 1u] = lt.axes[2u * i + 1u];
    if (want_fn) fn[i] = lt.face_fn[i];
  }
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) meta[i] = lt.shape_meta[i];
  __syncthreads();
  const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const uint32_t ax_off = n * HB_MAX_FACES * 16u, meta_off = ax_off + n * HB_MAX_FACES * 32u;
  return Tables<true>{ base, base + ax_off, base + meta_off, base + meta_off + n * 4u };
}

template <bool GENERAL>
HB_DEV void fork_append(const TraceParams& tp, uint32_t slot, uint32_t bits, float4 q, float px, float py, float pz,
                        float dx, float dy, float dz, float w, uint32_t new_face) {
  const uint32_t k = atomicAdd(tp.fork_count, 1u);
  if (k >= tp.fork_cap) {
    *tp.error_flag = 3u;
    return;
  }
  const uint32_t dst = tp.n_main + k;
  const uint32_t nb = bits_with_face(bits, new_face) | (1u << 30);
  tp.P[dst] = make_float4(px, py, pz, __uint_as_float(nb));
  tp.D[dst] = make_float4(dx, dy, dz, w);
  tp.Q[dst] = q;
  if (GENERAL) {
    uint32_t root, code;
    if (slot < tp.n_main) {
      root = tp.root_base + slot;
      code = 0u;
    } else {
      root = tp.fork_root[slot - tp.n_main];
      code = tp.fork_code[slot - tp.n_main];
    }
    tp.fork_root[k] = root;
    tp.fork_code[k] = code | (1u << (tp.hit & 31u));
    if (tp.M != nullptr) tp.M[dst] = tp.M[slot];
    if (tp.flags & kFlagPath) {
      for (uint32_t h = 0; h <= tp.hit; h++)
        tp.path[static_cast<size_t>(h) * tp.cap + dst] = tp.path[static_cast<size_t>(h) * tp.cap + slot];
      if (tp.hit + 1u < tp.max_hits) tp.path[static_cast<size_t>(tp.hit + 1u) * tp.cap + dst] = static_cast<uint8_t>(new_face);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// optics kernel: one surface interaction per live ray
// ------------------------------------------------------------------------------------------------
#ifndef HB_OPTICS_MINB
#define HB_OPTICS_MINB 4
#endif
#ifndef HB_INTERSECT_MINB
#define HB_INTERSECT_MINB 5
#endif
template <bool GENERAL, bool LAST, bool SMEM, bool MULTI>
__global__ void __launch_bounds__(256, HB_OPTICS_MINB) optics_kernel(const TraceParams tp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Tally tally;
  const bool use_cache = (tp.flags & kFlagPixelCache) != 0u;
  if (use_cache) cache_init(tally, smem_raw);
  const Tables<SMEM> tb = stage_tables<SMEM>(tp.lt, smem_raw + (use_cache ? kCacheBytes : 0), GENERAL);
  if (!SMEM && use_cache) __syncthreads();
  const uint32_t total = tp.n_main + *tp.fork_snapshot;
  const uint32_t stride = gridDim.x * blockDim.x;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  // software pipeline: the next ray's state is in flight while this one is being computed
  float4 d4 = make_float4(0.f, 0.f, 0.f, -1.f), p4 = d4, q = d4;
  if (i < total) {
    d4 = tp.D[i];
    p4 = tp.P[i];
    q = tp.Q[i];
  }
  while (i < total) {
    const uint32_t i_next = i + stride;
    float4 d_n = make_float4(0.f, 0.f, 0.f, -1.f), p_n = d_n, q_n = d_n;
    if (i_next < total) {
      d_n = tp.D[i_next];
      p_n = tp.P[i_next];
      q_n = tp.Q[i_next];
    }
    const uint32_t bits = __float_as_uint(p4.w);
    const uint32_t face = bits_face(bits);
    if (d4.w >= 0.0f && face != kFaceInvalid) {  // else: terminated ray
      const uint32_t shape = bits_shape(bits);
      const uint32_t meta = tb.meta(shape);
      const AxisRow<SMEM> axes = tb.axes(shape);
      const uint32_t axis_cnt = (meta >> 16) & 255u;
      const float n_idx = tp.wl[bits_wl(bits)].n_idx;

      const float4 pl = tb.plane(shape, face);
      const Split s = hit_surface(pl, n_idx, d4.x, d4.y, d4.z, d4.w);
      // The child on the far side of the face normally leaves the crystal: classify it here.
      const uint32_t out_child = s.cos_in > 0.0f ? 1u : 0u;  // internal hit: refracted; entry: reflected
      const float ox = out_child ? s.tx : s.rx, oy = out_child ? s.ty : s.ry, oz = out_child ? s.tz : s.rz;
      const float ow = out_child ? s.tw : s.rw;
      const float ix = out_child ? s.rx : s.tx, iy = out_child ? s.ry : s.ty, iz = out_child ? s.rz : s.tz;
      const float iw = out_child ? s.rw : s.tw;
      if (ow >= 0.0f) {
        float nx = 0.f, ny = 0.f, nz = 0.f;
        uint32_t nf = kFaceInvalid;
        if (!far_child_surely_exits(axes, axis_cnt, face, pl, p4.x, p4.y, p4.z, ox, oy, oz))
          nf = slab_exit<true>(axes, axis_cnt, face, p4.x, p4.y, p4.z, ox, oy, oz, nx, ny, nz);
        if (nf == kFaceInvalid) {
          emit_exit<GENERAL, MULTI>(tp, i, bits, q, ox, oy, oz, ow, /*role=*/0u, tb, tally);
        } else if (!LAST) {
          fork_append<GENERAL>(tp, i, bits, q, nx, ny, nz, ox, oy, oz, ow, nf);  // near-edge leak: both children stay
        }
      }
      if (LAST) {
        // no intersect pass follows the final interaction: classify the inside child here too
        if (iw >= 0.0f) {
          float nx, ny, nz;
          const uint32_t nf = slab_exit<false>(axes, axis_cnt, face, p4.x, p4.y, p4.z, ix, iy, iz, nx, ny, nz);
          if (nf == kFaceInvalid) emit_exit<GENERAL, MULTI>(tp, i, bits, q, ix, iy, iz, iw, /*role=*/1u, tb, tally);
        }
      } else {
        tp.D[i] = make_float4(ix, iy, iz, iw);  // iw < 0 (TIR sentinel) terminates the ray
      }
    }
    d4 = d_n;
    p4 = p_n;
    q = q_n;
    i = i_next;
  }
  if (use_cache) cache_flush(tp, tally);
  if (GENERAL && (tp.flags & kFlagStats) && tally.exits != 0ull) {
    atomicAdd(tp.stat_exit_count, tally.exits);
    atomicAdd(tp.stat_w_sum, tally.w_sum);
  }
}

// ------------------------------------------------------------------------------------------------
// intersect kernel: slab exit-face search for the inside child
// ------------------------------------------------------------------------------------------------
template <bool GENERAL, bool SMEM, bool MULTI>
__global__ void __launch_bounds__(256, HB_INTERSECT_MINB) intersect_kernel(const TraceParams tp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Tables<SMEM> tb = stage_tables<SMEM>(tp.lt, smem_raw, GENERAL);
  const uint32_t forks = *tp.fork_count;
  const uint32_t total = tp.n_main + min(forks, tp.fork_cap);
  if (blockIdx.x == 0 && threadIdx.x == 0) *tp.fork_snapshot = min(forks, tp.fork_cap);
  const uint32_t stride = gridDim.x * blockDim.x;
  Tally tally;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  float4 d4 = make_float4(0.f, 0.f, 0.f, -1.f), p4 = d4;
  if (i < total) {
    d4 = tp.D[i];
    p4 = tp.P[i];
  }
  while (i < total) {
    const uint32_t i_next = i + stride;
    float4 d_n = make_float4(0.f, 0.f, 0.f, -1.f), p_n = d_n;
    if (i_next < total) {
      d_n = tp.D[i_next];
      p_n = tp.P[i_next];
    }
    const uint32_t bits = __float_as_uint(p4.w);
    const uint32_t face = bits_face(bits);
    if (d4.w >= 0.0f && face != kFaceInvalid) {
      if (bits_advanced(bits)) {  // fork ray: advanced when it was created
        tp.P[i] = make_float4(p4.x, p4.y, p4.z, __uint_as_float(bits & ~(1u << 30)));
      } else {
        const uint32_t shape = bits_shape(bits);
        const uint32_t meta = tb.meta(shape);
        float nx, ny, nz;
        const uint32_t nf = slab_exit<false>(tb.axes(shape), (meta >> 16) & 255u, face, p4.x, p4.y, p4.z, d4.x, d4.y, d4.z, nx, ny, nz);
        if (nf == kFaceInvalid) {
          // the inside child found no face: it is outgoing (CollectData branch 1) and the ray ends here
          emit_exit<GENERAL, MULTI>(tp, i, bits, tp.Q[i], d4.x, d4.y, d4.z, d4.w, /*role=*/1u, tb, tally);
          tp.D[i] = make_float4(d4.x, d4.y, d4.z, -1.0f);
        } else {
          tp.P[i] = make_float4(nx, ny, nz, __uint_as_float(bits_with_face(bits, nf)));
          if (GENERAL && (tp.flags & kFlagPath) && tp.hit + 1u < tp.max_hits)
            tp.path[static_cast<size_t>(tp.hit + 1u) * tp.cap + i] = static_cast<uint8_t>(nf);
        }
      }
    }
    d4 = d_n;
    p4 = p_n;
    i = i_next;
  }
  if (GENERAL && (tp.flags & kFlagStats) && tally.exits != 0ull) {
    atomicAdd(tp.stat_exit_count, tally.exits);
    atomicAdd(tp.stat_w_sum, tally.w_sum);
  }
}

// ------------------------------------------------------------------------------------------------
// root generation / layer transit
// ------------------------------------------------------------------------------------------------
struct GenShared {
  float lut[3 * HB_LUT_NODES];
  HbCrystalTables shape0;        // single-shape populations: entry fan table staged on chip
  float4 tri_na[HB_MAX_SUBTRIS];  // (normal, area) per fan triangle: one LDS.128 per categorical term
};

constexpr uint32_t kFastTris = 24;  // prism = 20 fan triangles: weights kept in registers

// Entry sampling, single-shape fast path: identical arithmetic and summation order as sample_entry, but
// the triangle weights live in registers (one pass over shared memory, branch-free pick).
HB_DEV void sample_entry_fast(Stream& s, const GenShared* gs, float dx, float dy, float dz, float& px, float& py,
                              float& pz, uint32_t& face) {
  const uint32_t n = gs->shape0.subtri_cnt;
  float w[kFastTris];
  float total = 0.0f;
#pragma unroll
  for (uint32_t i = 0; i < kFastTris; i++) {
    w[i] = 0.0f;
    if (i < n) {
      const float4 na = gs->tri_na[i];
      const float dt = dx * na.x + dy * na.y + dz * na.z;
      w[i] = fmaxf(-dt * na.w, 0.0f);
      total += w[i];
    }
  }
  const float u_cat = s.next();
  uint32_t tri = 0u;
  if (total > 0.0f) {
    const float target = u_cat * total;
    float cum = 0.0f;
    bool found = false;
    tri = n - 1u;
#pragma unroll
    for (uint32_t i = 0; i < kFastTris; i++) {
      if (i < n) {
        cum += w[i];
        if (!found && cum > target) {
          tri = i;
          found = true;
        }
      }
    }
  }
  float u = s.next(), v = s.next();
  if (u + v > 1.0f) {
    u = 1.0f - u;
    v = 1.0f - v;
  }
  const float* tv = gs->shape0.tri_v[tri];
  px = u * (tv[3] - tv[0]) + v * (tv[6] - tv[0]) + tv[0];
  py = u * (tv[4] - tv[1]) + v * (tv[7] - tv[1]) + tv[1];
  pz = u * (tv[5] - tv[2]) + v * (tv[8] - tv[2]) + tv[2];
  face = gs->shape0.tri_face[tri];
}

template <bool TRANSIT>
__global__ void __launch_bounds__(256) gen_kernel(const GenParams gp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GenShared* gs = reinterpret_cast<GenShared*>(smem_raw);
  if (gp.axis.lat_path == HB_LAT_LUT) {
    for (uint32_t i = threadIdx.x; i < 3 * HB_LUT_NODES; i += blockDim.x) gs->lut[i] = gp.lut[i];
  }
  if (gp.shape_cnt == 1u) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(gp.shapes);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&gs->shape0);
    for (uint32_t i = threadIdx.x; i < sizeof(HbCrystalTables) / 4; i += blockDim.x) dst[i] = src[i];
    for (uint32_t i = threadIdx.x; i < HB_MAX_SUBTRIS; i += blockDim.x)
      gs->tri_na[i] = make_float4(gp.shapes->tri_n[i][0], gp.shapes->tri_n[i][1], gp.shapes->tri_n[i][2], gp.shapes->tri_area[i]);
  }
  __syncthreads();
  const bool fast_entry = gp.shape_cnt == 1u && gs->shape0.subtri_cnt <= kFastTris;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < gp.count; k += gridDim.x * blockDim.x) {
    const uint32_t lo = gp.idx_lo + k;
    const uint32_t hi = gp.idx_hi + (lo < gp.idx_lo ? 1u : 0u);
    const uint32_t s0 = seed_with_high(gp.seed, hi);
    uint32_t wl_i = 0u;
    float wx, wy, wz, weight;
    if (TRANSIT) {
      uint32_t src = gp.cont_first + k;
      if (gp.shuffle) src = feistel(src, gp.cont_n, gp.shuffle_seed);
      const float4 c = gp.cont_dw[src];
      wx = c.x;
      wy = c.y;
      wz = c.z;
      weight = c.w;
      wl_i = gp.cont_meta[src] & 255u;
      if (gp.cont_mask != nullptr) gp.M[gp.slot0 + k] = gp.cont_mask[src];
    } else if (gp.wl_cnt > 1u) {
      wl_i = min(static_cast<uint32_t>(draw(s0 ^ kNonceWl, lo, 0u) * static_cast<float>(gp.wl_cnt)), gp.wl_cnt - 1u);
    }
    Stream s{ s0, lo, 0u };
    float lon, lat, roll;
    sample_lon_lat_roll(s, gp.axis, gs->lut, lon, lat, roll);
    const float4 q = quat_from_angles(lon, lat, roll);
    const Rot r = rot_from_quat(q);
    if (!TRANSIT) {
      // sample_sph_cap (pcg_shared.h:514-529) with the per-launch trigonometry hoisted to the host
      const float u = s.next();
      const float x = u + (1.0f - u) * gp.sun_c_cap;
      const float rr = sqrtf(fmaxf(1.0f - x * x, 0.0f));
      float sp, cp;
      sincosf(s.next() * 2.0f * kPiF, &sp, &cp);
      const float y = cp * rr, z = sp * rr;
      wx = gp.sun_c_lon * gp.sun_c_lat * x - gp.sun_s_lon * y - gp.sun_c_lon * gp.sun_s_lat * z;
      wy = gp.sun_s_lon * gp.sun_c_lat * x + gp.sun_c_lon * y - gp.sun_s_lon * gp.sun_s_lat * z;
      wz = gp.sun_s_lat * x + gp.sun_c_lat * z;
      weight = gp.wl[wl_i].spd_weight;
    }
    float dx, dy, dz;
    rot_apply_t(r.m, wx, wy, wz, dx, dy, dz);
    uint32_t sh = 0u;
    if (gp.shape_cnt > 1u) {
      sh = min(static_cast<uint32_t>(draw(s0 ^ kNonceShape, lo, 0u) * static_cast<float>(gp.shape_cnt)), gp.shape_cnt - 1u);
    }
    const HbCrystalTables* tab = gp.shape_cnt == 1u ? &gs->shape0 : gp.shapes + sh;
    float px = 0.0f, py = 0.0f, pz = 0.0f;
    uint32_t face = kFaceInvalid;
    if (tab->subtri_cnt == 0u) {
      weight = -1.0f;  // degenerate crystal: nothing to trace (zero-weight discard, simulator.cpp:149-159)
    } else {
      if (fast_entry) sample_entry_fast(s, gs, dx, dy, dz, px, py, pz, face);
      else sample_entry(s, tab, dx, dy, dz, px, py, pz, face);
    }
    const uint32_t slot = gp.slot0 + k;
    gp.P[slot] = make_float4(px, py, pz, __uint_as_float(pack_bits(face, wl_i, gp.shape_base + sh, 0u)));
    gp.D[slot] = make_float4(dx, dy, dz, weight);
    gp.Q[slot] = q;
    if ((gp.flags & kFlagPath) && gp.path != nullptr) gp.path[slot] = static_cast<uint8_t>(face);
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 accumulators absorb small addends once a pixel grows (a 2e5 sun pixel has ulp 0.016): every
// `fold_rays` root rays the working float4 image is folded into a double-precision master image and
// zeroed, which bounds the relative loss (measured 2e-6 at 1 Mi rays, 6e-4 at 16 Mi without folding).
// The reference bounds the same error by draining every 64 batches into a host Neumaier sum
// (simulator.hpp:136, accum_shared.h:71-75).
__global__ void __launch_bounds__(256) fold_image_kernel(float4* image, double4* master, uint32_t pixels) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += gridDim.x * blockDim.x) {
    const float4 v = image[i];
    if (v.x != 0.0f || v.y != 0.0f || v.z != 0.0f || v.w != 0.0f) {
      double4 m = master[i];
      m.x += static_cast<double>(v.x);
      m.y += static_cast<double>(v.y);
      m.z += static_cast<double>(v.z);
      m.w += static_cast<double>(v.w);
      master[i] = m;
      image[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
  }
}

// image drain: master (X,Y,Z,landed) double4 -> packed fp32 XYZ + landed-weight sum, then zero
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) drain_image_kernel(double4* master, float* xyz, double* landed, uint32_t pixels) {
  double acc = 0.0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += gridDim.x * blockDim.x) {
    const double4 v = master[i];
    xyz[static_cast<size_t>(i) * 3 + 0] = static_cast<float>(v.x);
    xyz[static_cast<size_t>(i) * 3 + 1] = static_cast<float>(v.y);
    xyz[static_cast<size_t>(i) * 3 + 2] = static_cast<float>(v.z);
    acc += v.w;
    master[i] = make_double4(0.0, 0.0, 0.0, 0.0);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double warp_sum[8];
  if ((threadIdx.x & 31u) == 0u) warp_sum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; w++) t += warp_sum[w];
    atomicAdd(landed, t);
  }
}

// Non-destructive readout of one render: packed fp32 XYZ + landed-weight sum (the master keeps accumulating).
__global__ void __launch_bounds__(256) peek_image_kernel(const double4* master, float* xyz, double* landed, uint32_t pixels) {
  double lsum = 0.0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += gridDim.x * blockDim.x) {
    const double4 v = master[i];
    if (xyz != nullptr) {
      xyz[3u * i + 0] = static_cast<float>(v.x);
      xyz[3u * i + 1] = static_cast<float>(v.y);
      xyz[3u * i + 2] = static_cast<float>(v.z);
    }
    lsum += v.w;
  }
  for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xFFFFFFFFu, lsum, o);
  if ((threadIdx.x & 31u) == 0u && lsum != 0.0) atomicAdd(landed, lsum);
}

// Display sink (RenderConsumer::PostSnapshot, server/render.cpp:508-577, with util/color_space.cpp:10-52):
// XYZ * exposure scale -> gamut clip towards the D65 grey of equal luminance -> linear sRGB -> + background,
// clamp -> sRGB transfer curve -> 8-bit. With a ray colour the luminance-only branch is taken instead.
struct SnapshotParams {
  float scale;
  float ray_color[3];
  float background[3];
  int use_real_color;
};

HB_DEV float linear_to_srgb(float v) {  // color_space.cpp:47-52
  if (v < 0.0031308f) return v * 12.92f;
  return 1.055f * powf(v, 1.0f / 2.4f) - 0.055f;
}

__global__ void __launch_bounds__(256) snapshot_srgb_kernel(const double4* master, uint8_t* rgb8, uint32_t pixels, SnapshotParams sp) {
  const float kWhite[3] = { 0.95047f, 1.00000f, 1.08883f };
  const float kM[9] = { 3.2404542f, -1.5371385f, -0.4985314f, -0.9692660f, 1.8760108f, 0.0415560f,
                        0.0556434f, -0.2040259f, 1.0572252f };
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += gridDim.x * blockDim.x) {
    const double4 m = master[i];
    const float xyz[3] = { mul(static_cast<float>(m.x), sp.scale), mul(static_cast<float>(m.y), sp.scale),
                           mul(static_cast<float>(m.z), sp.scale) };
    float gray[3], rgb[3];
#pragma unroll
    for (int j = 0; j < 3; j++) gray[j] = mul(kWhite[j], xyz[1]);
    if (sp.use_real_color) {
      float s = 1.0f, diff[3];
#pragma unroll
      for (int j = 0; j < 3; j++) diff[j] = sub(xyz[j], gray[j]);
#pragma unroll
      for (int j = 0; j < 3; j++) {
        float a = 0.0f, b = 0.0f;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          a = add(a, mul(-gray[k], kM[j * 3 + k]));
          b = add(b, mul(diff[k], kM[j * 3 + k]));
        }
        if (mul(a, b) > 0.0f && __fdiv_rn(a, b) < s) s = __fdiv_rn(a, b);
      }
#pragma unroll
      for (int j = 0; j < 3; j++) {
        float v = 0.0f;
#pragma unroll
        for (int k = 0; k < 3; k++) v = add(v, mul(add(mul(diff[k], s), gray[k]), kM[j * 3 + k]));
        rgb[j] = fminf(fmaxf(v, 0.0f), 1.0f);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 3; j++) {
        float v = 0.0f;
#pragma unroll
        for (int k = 0; k < 3; k++) v = add(v, mul(gray[k], kM[j * 3 + k]));
        rgb[j] = mul(v, sp.ray_color[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 3; j++) {
      float v = add(rgb[j], sp.background[j]);
      v = fminf(fmaxf(v, 0.0f), 1.0f);
      rgb8[3u * i + j] = static_cast<uint8_t>(mul(linear_to_srgb(v), 255.0f));
    }
  }
}

// Export helper: quaternion -> rot9 with the device's own arithmetic (parity harness).
__global__ void quat_to_rot_kernel(const float4* Q, float* rot9, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Rot r = rot_from_quat(Q[i]);
  for (int k = 0; k < 9; k++) rot9[static_cast<size_t>(i) * 9 + k] = r.m[k];
}

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
};

struct LayerDev {
  float prob = 0.0f;
  std::vector<PopHost> pops;
  std::vector<double> carry;  // PartitionCrystalRayNum carry (simulator.cpp:519-582)
  uint32_t shape_cnt = 0;
  bool any_filter = false;
  DevBuf<float4> planes;
  DevBuf<float4> axes;
  DevBuf<uint32_t> shape_meta;
  DevBuf<uint8_t> face_fn;
  DevBuf<HbCrystalTables> shapes;
  DevBuf<HbFilterDesc> filters;
  DevBuf<uint32_t> pop_crystal_id;
  DevBuf<float> luts;  // [pop][3][257]
  bool any_color = false;              // some population carries colour predicates
  DevBuf<HbColorGroup> color_groups;   // [pop][HB_MAX_COLOR_GROUPS]
  DevBuf<uint32_t> color_group_cnt;    // [pop]
};

struct EventPair {
  cudaEvent_t a, b;
  int family;       // 0 gen, 1 optics, 2 intersect
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
  DevBuf<uint64_t> cont_mask[2];       // component masks of the continuation pools
  DevBuf<uint8_t> rgb_stage;           // 8-bit snapshot staging
  DevBuf<float4> image;
  DevBuf<double4> master;
  uint64_t rays_since_fold = 0;
  uint64_t fold_rays = 1u << 21;
  DevBuf<float> xyz_stage;
  DevBuf<double> landed_dev;

  // session
  bool in_session = false;
  HbSessionSpec spec{};
  struct WlSlot {
    std::vector<HbWlEntry> host;
    DevBuf<HbWlEntry> dev;
  };
  std::vector<WlSlot> wl_cache;
  size_t wl_cache_next = 0;
  const HbWlEntry* wl_cur = nullptr;
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
  DevBuf<uint32_t> counters;    // [0] fork_count [1] fork_snapshot [2] cont_count [3] exit_count [4] error
  DevBuf<unsigned long long> stat_cnt;
  DevBuf<double> stat_sum;
  DevBuf<float4> cont_dw[2];
  DevBuf<uint32_t> cont_meta[2];
  DevBuf<uint32_t> cont_root[2];
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
  int blocks_per_sm = 8;
  int blocks_per_sm_override = 0;
  bool pixel_cache = true;
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

void flush_events(HbEngine* h) {
  for (size_t i = 0; i < h->ev_used; i++) {
    float ms = 0.0f;
    EventPair& e = h->ev_pool[i];
    if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) {
      if (e.family == 0) {
        h->ctr.gen_ms += ms;
      } else if (e.family == 1) {
        h->ctr.optics_ms += ms;
      } else {
        h->ctr.intersect_ms += ms;
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

// Persistent grid-stride launch: exactly as many CTAs as are co-resident (occupancy x SM count), so there is
// no partially filled second wave.
template <typename K>
uint32_t resident_grid(const HbEngine* h, K kernel, size_t smem, uint64_t n) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, smem) != cudaSuccess || per_sm < 1) per_sm = 4;
  if (h->blocks_per_sm_override > 0) per_sm = h->blocks_per_sm_override;
  const uint64_t blocks = (n + 255) / 256;
  return static_cast<uint32_t>(std::max<uint64_t>(1, std::min<uint64_t>(blocks, static_cast<uint64_t>(h->sm_count) * per_sm)));
}

template <bool G, bool L, bool S, bool M = false>
void launch_optics_t(HbEngine* h, size_t smem, const TraceParams& tp) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(optics_kernel<G, L, S, M>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         static_cast<int>(shared_tables_bytes(kSmemShapes) + kCacheBytes));
    attr_set = true;
  }
  const uint32_t grid = resident_grid(h, optics_kernel<G, L, S, M>, smem, tp.cap);
  optics_kernel<G, L, S, M><<<grid, 256, smem, h->stream>>>(tp);
}
void launch_optics(HbEngine* h, bool general, bool last, bool in_smem, size_t smem, const TraceParams& tp) {
  const int key = (general ? 4 : 0) | (last ? 2 : 0) | (in_smem ? 1 : 0);
  if (tp.extra_cnt != 0u || tp.color_on != 0u) {  // N projections per trace / raypath colour: extended GENERAL kernels
    switch (key & 3) {
      case 0: launch_optics_t<true, false, false, true>(h, smem, tp); break;
      case 1: launch_optics_t<true, false, true, true>(h, smem, tp); break;
      case 2: launch_optics_t<true, true, false, true>(h, smem, tp); break;
      default: launch_optics_t<true, true, true, true>(h, smem, tp); break;
    }
    return;
  }
  switch (key) {
    case 0: launch_optics_t<false, false, false>(h, smem, tp); break;
    case 1: launch_optics_t<false, false, true>(h, smem, tp); break;
    case 2: launch_optics_t<false, true, false>(h, smem, tp); break;
    case 3: launch_optics_t<false, true, true>(h, smem, tp); break;
    case 4: launch_optics_t<true, false, false>(h, smem, tp); break;
    case 5: launch_optics_t<true, false, true>(h, smem, tp); break;
    case 6: launch_optics_t<true, true, false>(h, smem, tp); break;
    default: launch_optics_t<true, true, true>(h, smem, tp); break;
  }
}
template <bool G, bool S, bool M = false>
void launch_intersect_t(HbEngine* h, size_t smem, const TraceParams& tp) {
  const uint32_t grid = resident_grid(h, intersect_kernel<G, S, M>, smem, tp.cap);
  intersect_kernel<G, S, M><<<grid, 256, smem, h->stream>>>(tp);
}
void launch_intersect(HbEngine* h, bool general, bool in_smem, size_t smem, const TraceParams& tp) {
  if (tp.extra_cnt != 0u || tp.color_on != 0u) {
    if (in_smem) launch_intersect_t<true, true, true>(h, smem, tp); else launch_intersect_t<true, false, true>(h, smem, tp);
  } else if (general) {
    if (in_smem) launch_intersect_t<true, true>(h, smem, tp); else launch_intersect_t<true, false>(h, smem, tp);
  } else {
    if (in_smem) launch_intersect_t<false, true>(h, smem, tp); else launch_intersect_t<false, false>(h, smem, tp);
  }
}

size_t trace_smem(const LayerDev& L) { return shared_tables_bytes(L.shape_cnt); }

int upload_layer(HbEngine* h, const HbLayer& src, LayerDev* L) {
  L->prob = src.prob;
  L->pops.clear();
  std::vector<float4> planes;
  std::vector<float4> axes;
  std::vector<uint32_t> meta;
  std::vector<uint8_t> fn;
  std::vector<HbCrystalTables> shapes;
  std::vector<HbFilterDesc> filters;
  std::vector<uint32_t> cid;
  std::vector<float> luts;
  std::vector<HbColorGroup> cgroups;
  std::vector<uint32_t> cgroup_cnt;
  L->any_filter = false;
  L->any_color = false;
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
    L->any_filter = L->any_filter || ph.has_filter;
    for (uint32_t s = 0; s < p.shape_cnt; s++) {
      const HbCrystalTables& t = p.shapes[s];
      if (t.face_cnt > HB_MAX_FACES || t.subtri_cnt > HB_MAX_SUBTRIS) return fail(h, HB_ERR_INVALID_ARG, "crystal table too large");
      shapes.push_back(t);
      for (uint32_t f = 0; f < HB_MAX_FACES; f++) {
        planes.push_back(make_float4(t.plane[f][0], t.plane[f][1], t.plane[f][2], t.plane[f][3]));
        fn.push_back(t.face_fn[f]);
      }
      // Axis table: faces whose unit normals are exact negatives share one entry (see slab_exit).
      uint32_t axis_cnt = 0;
      bool used[HB_MAX_FACES] = {};
      const size_t ax0 = axes.size();
      axes.resize(ax0 + HB_MAX_FACES * 2, make_float4(0.f, 0.f, 0.f, 0.f));
      for (uint32_t f = 0; f < t.face_cnt; f++) {
        if (used[f]) continue;
        used[f] = true;
        uint32_t partner = kFaceInvalid;
        for (uint32_t g = f + 1; g < t.face_cnt; g++) {
          if (!used[g] && t.plane[g][0] == -t.plane[f][0] && t.plane[g][1] == -t.plane[f][1] && t.plane[g][2] == -t.plane[f][2]) {
            partner = g;
            used[g] = true;
            break;
          }
        }
        const uint32_t fbits = f | (partner << 8);
        float fb;
        std::memcpy(&fb, &fbits, 4);
        axes[ax0 + 2 * axis_cnt] = make_float4(t.plane[f][0], t.plane[f][1], t.plane[f][2], t.plane[f][3]);
        axes[ax0 + 2 * axis_cnt + 1] = make_float4(partner == kFaceInvalid ? 0.0f : t.plane[partner][3], fb, 0.f, 0.f);
        axis_cnt++;
      }
      meta.push_back(t.face_cnt | (ci << 8) | (axis_cnt << 16));
    }
    filters.push_back(p.filter);
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
  HB_CUDA(h, L->planes.ensure(planes.size()));
  HB_CUDA(h, L->axes.ensure(axes.size()));
  HB_CUDA(h, L->shape_meta.ensure(meta.size()));
  HB_CUDA(h, L->face_fn.ensure(fn.size()));
  HB_CUDA(h, L->shapes.ensure(shapes.size()));
  HB_CUDA(h, L->filters.ensure(filters.size()));
  HB_CUDA(h, L->pop_crystal_id.ensure(cid.size()));
  HB_CUDA(h, L->luts.ensure(luts.size()));
  HB_CUDA(h, cudaMemcpy(L->planes.p, planes.data(), planes.size() * sizeof(float4), cudaMemcpyHostToDevice));
  HB_CUDA(h, cudaMemcpy(L->axes.p, axes.data(), axes.size() * sizeof(float4), cudaMemcpyHostToDevice));
  HB_CUDA(h, cudaMemcpy(L->shape_meta.p, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice));
  HB_CUDA(h, cudaMemcpy(L->face_fn.p, fn.data(), fn.size(), cudaMemcpyHostToDevice));
  HB_CUDA(h, cudaMemcpy(L->shapes.p, shapes.data(), shapes.size() * sizeof(HbCrystalTables), cudaMemcpyHostToDevice));
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
  const uint32_t fork_cap = std::max<uint32_t>(4096u, n / 64u);
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
      gp.shapes = L.shapes.p + ph.shape_base;
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
      EventPair* ev = begin_event(h, 0, gp.count);
      if (li == 0) {
        gen_kernel<false><<<grid_for(h, gp.count), 256, sizeof(GenShared), h->stream>>>(gp);
      } else {
        const int src = h->cont_cur ^ 1;
        gp.cont_dw = h->cont_dw[src].p;
        gp.cont_meta = h->cont_meta[src].p;
        if (color_on) {
          gp.cont_mask = h->cont_mask[src].p;
          gp.M = h->mask_tile.p;
        }
        gp.cont_n = static_cast<uint32_t>(layer_total);
        gp.cont_first = static_cast<uint32_t>(b);
        gp.shuffle = h->cont_shuffle ? 1u : 0u;
        gp.shuffle_seed = pcg_hash_host(h->spec.seed ^ kNonceShuffle ^ h->shuffle_round);
        gen_kernel<true><<<grid_for(h, gp.count), 256, sizeof(GenShared), h->stream>>>(gp);
      }
      end_event(h, ev);
      h->ctr.kernel_launches++;
      h->ctr.gen_launches++;
    }
  }
  HB_CUDA(h, cudaGetLastError());

  // ---- parity export of the roots this tile starts from ----
  if (h->spec.record_exits && !(li == 0 && h->injected)) {
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
  TraceParams tp{};
  tp.P = h->P.p;
  tp.D = h->D.p;
  tp.Q = h->Q.p;
  tp.path = h->path.p;
  tp.fork_root = h->fork_root.p;
  tp.fork_code = h->fork_code.p;
  tp.fork_count = h->counters.p + 0;
  tp.fork_snapshot = h->counters.p + 1;
  tp.n_main = n;
  tp.cap = cap;
  tp.fork_cap = fork_cap;
  tp.root_base = static_cast<uint32_t>(root0);
  tp.lt = LayerTables{ L.planes.p, L.axes.p, L.shape_meta.p, L.face_fn.p, L.shapes.p, L.filters.p, L.pop_crystal_id.p,
                       L.shape_cnt, static_cast<uint32_t>(L.pops.size()), L.any_filter ? 1u : 0u,
                       L.any_color ? L.color_groups.p : nullptr, L.color_group_cnt.p };
  tp.wl = h->wl_cur;
  tp.wl_cnt = h->wl_cnt;
  tp.image = h->image.p;
  tp.proj = h->proj;
  tp.extra = h->extra_dev.p;
  tp.extra_cnt = static_cast<uint32_t>(h->renders.size()) - 1u;
  tp.color_on = color_on ? 1u : 0u;
  tp.M = (color_on && li > 0) ? h->mask_tile.p : nullptr;
  tp.cont_mask = (color_on && (flags & kFlagGate)) ? h->cont_mask[h->cont_cur].p : nullptr;
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
  tp.cont_dw = h->cont_dw[h->cont_cur].p;
  tp.cont_meta = h->cont_meta[h->cont_cur].p;
  tp.cont_root = h->spec.record_exits ? h->cont_root[h->cont_cur].p : nullptr;
  tp.cont_count = h->counters.p + 2;
  tp.cont_cap = static_cast<uint32_t>(h->cont_dw[h->cont_cur].n);
  tp.exits = h->exits_dev.p;
  tp.exit_root = h->exit_root_dev.p;
  tp.exit_count = h->counters.p + 3;
  tp.exit_cap = static_cast<uint32_t>(std::min<size_t>(h->exits_dev.n, 0xFFFFFFFFu));
  tp.stat_exit_count = h->stat_cnt.p;
  tp.stat_w_sum = h->stat_sum.p;
  tp.error_flag = h->counters.p + 4;

  const size_t smem = trace_smem(L);
  const bool in_smem = L.shape_cnt <= kSmemShapes;
  for (uint32_t hit = 0; hit < h->max_hits; hit++) {
    tp.hit = hit;
    const bool last = hit + 1 == h->max_hits;
    EventPair* ev = begin_event(h, 1, n);
    launch_optics(h, general, last, in_smem, smem + ((flags & kFlagPixelCache) ? kCacheBytes : 0), tp);
    end_event(h, ev);
    h->ctr.kernel_launches++;
    h->ctr.optics_launches++;
    h->ctr.optics_rays += n;
    if (!last) {
      ev = begin_event(h, 2, n);
      launch_intersect(h, general, in_smem, smem, tp);
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
  for (auto& L : h->layers) {
    L->planes.release();
    L->axes.release();
    L->shape_meta.release();
    L->face_fn.release();
    L->shapes.release();
    L->filters.release();
    L->pop_crystal_id.release();
    L->luts.release();
  }
  h->image.release();
  h->master.release();
  h->extra_dev.release();
  h->lanes.release();
  h->mask_tile.release();
  h->cont_mask[0].release();
  h->cont_mask[1].release();
  h->rgb_stage.release();
  h->xyz_stage.release();
  h->landed_dev.release();
  for (auto& s : h->wl_cache) s.dev.release();
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
    h->cont_dw[i].release();
    h->cont_meta[i].release();
    h->cont_root[i].release();
  }
  h->exits_dev.release();
  h->exit_root_dev.release();
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
  h->layers.clear();
  for (uint32_t li = 0; li < s->layer_cnt; li++) {
    auto L = std::make_unique<LayerDev>();
    int rc = upload_layer(h, s->layers[li], L.get());
    if (rc != HB_OK) return rc;
    h->layers.push_back(std::move(L));
  }
  h->max_hits = s->max_hits;
  h->sun_lon = s->sun_lon;
  h->sun_lat = s->sun_lat;
  h->sun_half = s->sun_half_angle;
  if (s->color_classes.class_cnt > HB_MAX_COLOR_CLASSES) return fail(h, HB_ERR_INVALID_ARG, "scene: more than 16 colour classes");
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
    }
    h->wl_cur = h->wl_cache[hit].dev.p;
  }
  if (spec->use_ray_base) h->gen_base = spec->ray_base;
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
  const uint32_t flags = session_flags(h, L, last_layer);

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

  // continuation pool of this layer (only when its exits can continue)
  HB_CUDA(h, h->counters.ensure(8));
  HB_CUDA(h, h->stat_cnt.ensure(1));
  HB_CUDA(h, h->stat_sum.ensure(1));
  HB_CUDA(h, cudaMemsetAsync(h->counters.p + 2, 0, 3 * sizeof(uint32_t), h->stream));  // cont_count, exit_count, error
  HB_CUDA(h, cudaMemsetAsync(h->stat_cnt.p, 0, sizeof(unsigned long long), h->stream));
  HB_CUDA(h, cudaMemsetAsync(h->stat_sum.p, 0, sizeof(double), h->stream));
  if (flags & kFlagGate) {
    const uint64_t ccap = std::min<uint64_t>(n * (h->max_hits + 1) + 4096, 0xFFFFFFF0ull);
    HB_CUDA(h, h->cont_dw[h->cont_cur].ensure(ccap));
    HB_CUDA(h, h->cont_meta[h->cont_cur].ensure(ccap));
    if (h->classes.class_cnt != 0) HB_CUDA(h, h->cont_mask[h->cont_cur].ensure(ccap));
    if (h->spec.record_exits) HB_CUDA(h, h->cont_root[h->cont_cur].ensure(ccap));
  }

  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (stats != nullptr) {
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, h->stream);
  }
  uint64_t tile = h->tile_rays;
  if (flags & kFlagRecord) tile = std::min<uint64_t>(tile, 1u << 18);
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
    if (c[2] != 0) return fail(h, HB_ERR_CAPACITY, "device buffer overflow (code " + std::to_string(c[2]) + ")");
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
  return HB_OK;
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
  h->ctr.kernel_launches++;
  double l = 0.0;
  HB_CUDA(h, cudaMemcpyAsync(xyz, h->xyz_stage.p, static_cast<size_t>(pix) * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  HB_CUDA(h, cudaMemcpyAsync(&l, h->landed_dev.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  HB_CUDA(h, cudaMemsetAsync(h->landed_dev.p, 0, sizeof(double), h->stream));
  HB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->profile) flush_events(h);
  *landed += static_cast<float>(l);
  return HB_OK;
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
  return HB_OK;
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
    if (value < 1 || value > 32) return fail(h, HB_ERR_INVALID_ARG, "blocks_per_sm out of range");
    h->blocks_per_sm = static_cast<int>(value);
    h->blocks_per_sm_override = static_cast<int>(value);
  } else if (k == "pixel_cache") {
    h->pixel_cache = value != 0;
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
  uint32_t err = 0;
  if (h->counters.p) {
    HB_CUDA(h, cudaMemcpy(&err, h->counters.p + 4, 4, cudaMemcpyDeviceToHost));
    if (err != 0) return fail(h, HB_ERR_CAPACITY, "device buffer overflow (code " + std::to_string(err) + ")");
  }
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

int hb_allreduce_image(HbEngine* h) {
  if (h == nullptr) return HB_ERR_INVALID_ARG;
#ifdef HB_WITH_NCCL
  if (h->comm == nullptr) return fail(h, HB_ERR_STATE, "hb_comm_init not called");
  cudaSetDevice(h->device);
  if (!h->have_render) return fail(h, HB_ERR_STATE, "no render set");
  const size_t cnt = static_cast<size_t>(h->arena_pix) * 4;
  fold_arena(h);
  if (ncclAllReduce(h->master.p, h->master.p, cnt, ncclDouble, ncclSum, h->comm, h->stream) != ncclSuccess)
    return fail(h, HB_ERR_COMM, "ncclAllReduce failed");
  return HB_OK;
#else
  return fail(h, HB_ERR_UNSUPPORTED, "library built without NCCL");
#endif
}

}  // extern "C"
