// hb_kernels.cuh — device kernels of the B200 wavefront trace engine (sm_100a): kernel parameter blocks,
// emission (filter / gate / projection / image reduction), shared-memory table staging and the
// gen / optics / intersect / image kernels. Host-side session logic lives in hb_engine.cu.
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
#ifndef HB_KERNELS_CUH_
#define HB_KERNELS_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "halotrace_b200.h"
#include "hb_device.cuh"
#include "hb_geometry.h"
#include "hb_tables.h"

namespace hb {

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

// Wavelength-pool entry as the kernels read it: the refractive index with its IEEE reciprocal (the entry-side
// Fresnel ratio 1/n of HitSurface, optics.cpp:24, computed once on the host instead of once per ray) and the CMF.
struct WlDev {
  float n_idx, inv_n, spd_weight, pad0;
  float cmf_x, cmf_y, cmf_z, pad1;
};
HB_DEV WlDev load_wl(const WlDev& wl0, const WlDev* wl2, uint32_t wl_cnt, uint32_t i) {
  if (wl_cnt == 1u) return wl0;
  const float4 a = __ldg(reinterpret_cast<const float4*>(wl2 + i));
  const float4 b = __ldg(reinterpret_cast<const float4*>(wl2 + i) + 1);
  return WlDev{ a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
}

// Continuation record (multi-scatter): everything that carries over to the next layer in ONE 32-byte sector, so the
// Feistel-permuted gather of the transit generator costs one random sector per continuation (it was two -- direction
// + weight and the meta word in separate arrays -- and three with raypath colour).
struct __align__(16) ContRec {
  float4 dw;                    // world direction, weight
  uint32_t meta;                // wavelength index | population << 8
  uint32_t root;                // layer-root index of the parent
  uint64_t mask;                // raypath-colour component mask
};
static_assert(sizeof(ContRec) == 32, "one sector");

struct TraceParams {
  float4* P;
  float4* D;
  float4* Q;
  uint8_t* path;                // [max_hits][cap] compact face ids (kFlagPath)
  uint32_t* fork_root;          // [fork_cap] layer-root index of fork rays
  uint32_t* fork_code;          // [fork_cap] branch code of fork rays
  uint32_t* fork_count;         // rays appended behind the main slots (near-edge double continuation)
  uint32_t* fork_snapshot;      // fork_count as of the last intersect kernel (split) / the last bounce kernel (fused)
  uint32_t* done_count;         // CTAs of the running bounce kernel that have retired (publish_fork_snapshot)
  uint32_t n_main, cap, fork_cap;
  uint32_t root_base;           // layer-root index of slot 0 of this tile
  LayerTables lt;
  const HbWlEntry* wl;
  uint32_t wl_cnt;
  WlDev wl0;                    // entry 0 of the derived table: single-wavelength sessions read it from the parameter bank
  const WlDev* wl2;             // derived table (n, 1/n, CMF), one per pool entry
  float4* image;                // image arena: render r occupies [off_r, off_r + W_r*H_r), (X, Y, Z, landed)
  double4* master;              // fp64 master of the same arena (the pixel cache flushes straight into it)
  HbProjParams proj;            // render 0 (arena offset 0)
  const ExtraRender* extra;     // renders 1..extra_cnt (device)
  uint32_t extra_cnt;
  // raypath colour (kernels instantiated with MULTI only)
  uint32_t color_on;            // scene has colour classes
  uint64_t* M;                  // [cap + fork_cap] component mask carried in from earlier layers (nullptr on layer 0)
  float* lane;                  // [class_cnt][lane_stride] per-class Y lanes of render 0
  uint32_t lane_stride;
  HbColorClasses classes;
  uint32_t hit, max_hits, layer_idx, flags;
  float prob;
  uint32_t gate_seed;           // session seed ^ gate nonce
  uint32_t gate_base_lo, gate_base_hi;  // global gate index of layer-root 0
  ContRec* cont;                // continuation pool being appended (one 32-byte record per continuation)
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

struct EntryFaces;
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
  const EntryFaces* entry_faces;  // face groups of the same shapes
  uint32_t shape_base, shape_cnt;
  const HbWlEntry* wl;
  uint32_t wl_cnt;
  float sun_lon, sun_lat, sun_half;
  float sun_c_cap, sun_c_lon, sun_s_lon, sun_c_lat, sun_s_lat;  // per-launch constants of sample_sph_cap
  uint32_t flags;
  // transit only
  const ContRec* cont;          // permuted source pool
  uint64_t* M;                  // per-slot carried mask of the next layer (raypath colour) or nullptr
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
  uint32_t* cache_keys = nullptr;  // per-CTA pixel cache: [owners | candidates] (below); nullptr = reduce straight to L2
  float* cache_vals = nullptr;
};

// Per-CTA pixel cache. Halo images are extremely peaked (the undeviated light through parallel faces lands
// on the ~35 pixels of the sun disk: ~40 % of all exits), and same-address reductions serialise in the L2
// atomic unit while the fp32 accumulator of such a pixel absorbs small addends. Each CTA therefore keeps a
// direct-mapped table of kCacheSlots pixels in shared memory (claimed on a pixel's second sighting, see
// accumulate_pixel): contributions to a cached pixel are summed in shared memory and added to the fp64 master once,
// when the CTA retires (cache_flush). Everything else goes straight to the L2 with one red.global.add.v4.f32.
//   * The slot is a TILE hash of the pixel coordinates (low bits of x and of y): the hot pixels of a halo image are
//     spatially clustered (sun disk, parhelia), and a 16 x 8 tile maps a 7 x 7 cluster without a single collision
//     where a multiplicative hash of the linear index loses ~4 of 35 to the birthday problem.
//   * Every slot has kCacheWays value cells, chosen by the lane: the lanes of a warp that hit the same hot pixel in
//     the same instruction serialise in the compare-and-swap loop (ncu: 3.7 rounds of ATOMS.CAS.128 per 32 projected
//     exits with one cell); cells per lane class cut the rounds by the number of ways.
#ifndef HB_CACHE_SLOTS_LOG2
#define HB_CACHE_SLOTS_LOG2 7
#endif
#ifndef HB_CACHE_WAYS_LOG2
#define HB_CACHE_WAYS_LOG2 2
#endif
#ifndef HB_EXIT_STAGES
#define HB_EXIT_STAGES 2   // 2: exits are projected from a second, visibility-culled queue (see queue_drain)
#endif
constexpr uint32_t kCacheSlots = 1u << HB_CACHE_SLOTS_LOG2;
constexpr uint32_t kCacheWays = 1u << HB_CACHE_WAYS_LOG2;
constexpr uint32_t kCacheXBits = (HB_CACHE_SLOTS_LOG2 + 1) / 2, kCacheYBits = HB_CACHE_SLOTS_LOG2 - kCacheXBits;
constexpr uint32_t kCacheEmpty = 0xFFFFFFFFu;
constexpr size_t kCacheBytes = kCacheSlots * (2 * sizeof(uint32_t) + kCacheWays * 4 * sizeof(float));  // keys, candidates, cells

// (x, y, z, w) += into one 16-byte shared-memory slot with a single 128-bit compare-and-swap loop
// (ATOMS.CAS.128). Shared memory has no native fp32 add: four scalar atomicAdd calls are four CAS loops.
HB_DEV void smem_add_f4(uint32_t addr, float x, float y, float z, float w) {
  asm volatile(
      "{\n"
      ".reg .b128 oldv, newv, got;\n"
      ".reg .b64 lo, hi, glo, ghi;\n"
      ".reg .f32 a0, a1, a2, a3;\n"
      ".reg .pred p, q;\n"
      "ld.shared.v2.b64 {lo, hi}, [%0];\n"
      "HB_CAS_RETRY:\n"
      "mov.b64 {a0, a1}, lo;\n"
      "mov.b64 {a2, a3}, hi;\n"
      "mov.b128 oldv, {lo, hi};\n"
      "add.rn.f32 a0, a0, %1;\n"
      "add.rn.f32 a1, a1, %2;\n"
      "add.rn.f32 a2, a2, %3;\n"
      "add.rn.f32 a3, a3, %4;\n"
      "mov.b64 glo, {a0, a1};\n"
      "mov.b64 ghi, {a2, a3};\n"
      "mov.b128 newv, {glo, ghi};\n"
      "atom.shared.cas.b128 got, [%0], oldv, newv;\n"
      "mov.b128 {glo, ghi}, got;\n"
      "setp.ne.b64 p, glo, lo;\n"
      "setp.ne.b64 q, ghi, hi;\n"
      "or.pred p, p, q;\n"
      "mov.b64 lo, glo;\n"
      "mov.b64 hi, ghi;\n"
      "@p bra HB_CAS_RETRY;\n"
      "}\n" ::"r"(addr), "f"(x), "f"(y), "f"(z), "f"(w)
      : "memory");
}

// `px`, `py`: the pixel's coordinates inside its render (tile hash); `pix`: its index in the image arena.
// Claiming is by SECOND SIGHTING: a pixel that finds its slot unowned leaves its index in the slot's candidate word
// (plain store: a lost race only delays a claim) and claims the slot (compare-and-swap from empty, so an owner is
// never replaced) only when it meets its own candidate again. A hot pixel -- most of its slot's traffic -- owns the
// slot after two or three exits; a cold one would need two consecutive sightings with no other pixel of the slot in
// between, so cold pixels practically never lock a hot one out (first come, first claimed lost 17-29 % of the hot
// pixels' traffic per CTA that way, which then went to the L2 one by one and into a large fp32 accumulator).
HB_DEV void accumulate_pixel(const TraceParams& tp, const Tally& tally, uint32_t pix, uint32_t px, uint32_t py, float x, float y,
                             float z, float lw) {
  if (tally.cache_keys != nullptr) {
    const uint32_t slot = (px & ((1u << kCacheXBits) - 1u)) | ((py & ((1u << kCacheYBits) - 1u)) << kCacheXBits);
    uint32_t k = tally.cache_keys[slot];
    if (k == kCacheEmpty) {
      uint32_t* cand = tally.cache_keys + kCacheSlots;
      if (cand[slot] == pix) {
        k = atomicCAS(&tally.cache_keys[slot], kCacheEmpty, pix);
        if (k == kCacheEmpty) k = pix;
      } else {
        cand[slot] = pix;
      }
    }
    if (k == pix) {
      const uint32_t cell = slot * kCacheWays + (threadIdx.x & (kCacheWays - 1u));
      smem_add_f4(static_cast<uint32_t>(__cvta_generic_to_shared(tally.cache_vals + cell * 4u)), x, y, z, lw);
      return;
    }
  }
  red_add_f4(tp.image + pix, x, y, z, lw);
}

HB_DEV void cache_init(Tally& tally, unsigned char* smem) {
  tally.cache_keys = reinterpret_cast<uint32_t*>(smem);
  tally.cache_vals = reinterpret_cast<float*>(smem + 2u * kCacheSlots * sizeof(uint32_t));
  for (uint32_t i = threadIdx.x; i < 2u * kCacheSlots; i += blockDim.x) tally.cache_keys[i] = kCacheEmpty;  // keys + candidates
  float4* v = reinterpret_cast<float4*>(tally.cache_vals);
  for (uint32_t i = threadIdx.x; i < kCacheSlots * kCacheWays; i += blockDim.x) v[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

HB_DEV void cache_flush(const TraceParams& tp, const Tally& tally) {
  __syncthreads();
  const float4* v = reinterpret_cast<const float4*>(tally.cache_vals);
  for (uint32_t i = threadIdx.x; i < kCacheSlots; i += blockDim.x) {
    const uint32_t k = tally.cache_keys[i];
    if (k == kCacheEmpty) continue;
    double x = 0.0, y = 0.0, z = 0.0, w = 0.0;
#pragma unroll
    for (uint32_t c = 0; c < kCacheWays; c++) {
      const float4 cell = v[i * kCacheWays + c];
      x += static_cast<double>(cell.x);
      y += static_cast<double>(cell.y);
      z += static_cast<double>(cell.z);
      w += static_cast<double>(cell.w);
    }
    // A CTA's partial sum of a hot pixel goes to the fp64 master, not to the fp32 working image: the working
    // accumulator of such a pixel then only ever holds the few contributions that arrived before the claim, stays
    // small, and does not swallow the ~1e-5-weight exits of long ray paths (a 1e4 fp32 pixel drops addends < 5e-4).
    double* m = reinterpret_cast<double*>(tp.master + k);
    atomicAdd(m + 0, x);
    atomicAdd(m + 1, y);
    atomicAdd(m + 2, z);
    atomicAdd(m + 3, w);
  }
}

// Emission, first half: crystal-local direction -> world, then everything that decides the exit's fate (filter,
// colour predicates, gate, continuation append, exit record, stats). Returns true when the exit goes on to the
// image (second half: emit_project) with its world direction and component mask.
template <bool GENERAL, bool MULTI, typename TablesT>
HB_DEV bool emit_world(const TraceParams& tp, uint32_t slot, uint32_t bits, float4 q, float lx, float ly, float lz,
                       float w, uint32_t role, const TablesT& tb, Tally& tally, float& wx, float& wy, float& wz,
                       uint64_t& mask) {
  const Rot r = rot_from_quat(q);
  rot_apply(r, lx, ly, lz, wx, wy, wz);
  const uint32_t wl_i = bits_wl(bits);
  bool to_next_layer = false;
  mask = 0ull;  // component mask (raypath colour)
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
    if (tp.lt.any_filter) {
      // A filter_in raypath filter only admits paths of its own length: everything else fails before the path
      // is even gathered (len is uniform over the launch, so whole hit levels skip the matcher).
      const HbFilterDesc& f = tp.lt.filters[pop];
      if (f.kind == 1u && f.action == 0u && len != f.simple.path_len) return false;
    }
    if (tp.flags & kFlagPath) {
      for (uint32_t k = 0; k < len; k++)
        fn_path[k] = static_cast<uint8_t>(tb.face_fn(shape, tp.path[static_cast<size_t>(k) * tp.cap + slot]));
      if (tp.lt.any_filter) {
        const HbFilterDesc& f = tp.lt.filters[pop];
        if (f.kind != 0u) {
          const float dir[3] = { wx, wy, wz };
          if (!filter_check(f, fn_path, len, dir, tp.lt.pop_crystal_id[pop])) return false;  // filter-fail terminates
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
        float4* rec = reinterpret_cast<float4*>(tp.cont + dst);
        rec[0] = make_float4(wx, wy, wz, w);
        reinterpret_cast<uint4*>(rec)[1] = make_uint4(wl_i | (pop << 8), root, static_cast<uint32_t>(mask), static_cast<uint32_t>(mask >> 32));
      } else {
        *tp.error_flag = 1u;
      }
      return false;
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
    if (!(tp.flags & kFlagAccum)) return false;
  }
  return true;
}

// Emission, second half: projection through every render of the trace + image reduction (+ colour lanes).
template <bool MULTI>
HB_DEV void emit_project(const TraceParams& tp, uint32_t wl_i, float wx, float wy, float wz, float w, uint64_t mask,
                         Tally& tally) {
  const PixelHits h = project_exit(tp.proj, wx, wy, wz);
  const WlDev we = load_wl(tp.wl0, tp.wl2, tp.wl_cnt, wl_i);
#pragma unroll
  for (int k = 0; k < 2; k++) {
    if (k < h.count) {
      const int px = h.px[k], py = h.py[k];
      if (px >= 0 && px < tp.proj.img_w && py >= 0 && py < tp.proj.img_h) {
        const uint32_t pix = static_cast<uint32_t>(py) * static_cast<uint32_t>(tp.proj.img_w) + static_cast<uint32_t>(px);
        accumulate_pixel(tp, tally, pix, static_cast<uint32_t>(px), static_cast<uint32_t>(py), mul(we.cmf_x, w), mul(we.cmf_y, w),
                         mul(we.cmf_z, w), h.bump[k] ? w : 0.0f);
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
                             static_cast<uint32_t>(px), static_cast<uint32_t>(py), mul(we.cmf_x, w), mul(we.cmf_y, w), mul(we.cmf_z, w),
                             hr.bump[k] ? w : 0.0f);
          }
        }
      }
    }
  }
}

template <bool GENERAL, bool MULTI, typename TablesT>
HB_DEV void emit_exit(const TraceParams& tp, uint32_t slot, uint32_t bits, float4 q, float lx, float ly, float lz,
                      float w, uint32_t role, const TablesT& tb, Tally& tally) {
  float wx, wy, wz;
  uint64_t mask;
  if (emit_world<GENERAL, MULTI>(tp, slot, bits, q, lx, ly, lz, w, role, tb, tally, wx, wy, wz, mask))
    emit_project<MULTI>(tp, bits_wl(bits), wx, wy, wz, w, mask, tally);
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

// `advanced`: the split pipeline marks a fork ray so that the intersect pass that follows skips it (it was advanced
// when it was created); the fused bounce kernel picks fork rays up at the next interaction and needs no mark.
template <bool GENERAL>
HB_DEV void fork_append(const TraceParams& tp, uint32_t slot, uint32_t bits, float4 q, float px, float py, float pz,
                        float dx, float dy, float dz, float w, uint32_t new_face, uint32_t advanced = 1u) {
  const uint32_t k = atomicAdd(tp.fork_count, 1u);
  if (k >= tp.fork_cap) {
    *tp.error_flag = 3u;
    return;
  }
  const uint32_t dst = tp.n_main + k;
  const uint32_t nb = bits_with_face(bits, new_face) | (advanced << 30);
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
#ifndef HB_ASYNC_STAGE
#define HB_ASYNC_STAGE 1
#endif
#if HB_ASYNC_STAGE
constexpr uint32_t kStageBytes = 2u * 3u * 256u * 16u;  // two stages x (D, P, Q) x 256 threads x 16 B
#else
constexpr uint32_t kStageBytes = 0u;
#endif
HB_DEV void cp_async16(uint32_t saddr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gptr) : "memory");
}
HB_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
HB_DEV void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

#ifndef HB_OPTICS_MINB_GENERAL
#define HB_OPTICS_MINB_GENERAL 4
#endif
#ifndef HB_INTERSECT_MINB_GENERAL
#define HB_INTERSECT_MINB_GENERAL 4
#endif
#ifndef HB_OPTICS_MINB
#define HB_OPTICS_MINB 4
#endif
#ifndef HB_INTERSECT_MINB
#define HB_INTERSECT_MINB 5
#endif
// P4: 0 = generic axis loop, 1 = every shape of the layer is a full hexagonal prism (unrolled forms only),
// 2 = mixed layer: chosen per ray from the shape's meta word (kMetaP4). Shapes are uniform over 32-ray blocks
// and populations occupy contiguous index ranges, so the choice is warp-uniform except at block boundaries.
template <bool GENERAL, bool LAST, bool SMEM, bool MULTI, int P4 = 0>
__global__ void __launch_bounds__(256, GENERAL ? HB_OPTICS_MINB_GENERAL : HB_OPTICS_MINB) optics_kernel(const TraceParams tp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Tally tally;
  const bool use_cache = (tp.flags & kFlagPixelCache) != 0u;
  if (use_cache) cache_init(tally, smem_raw);
  // dynamic shared memory: [pixel cache] [ray staging] [crystal tables]
  const uint32_t stage_off = use_cache ? static_cast<uint32_t>(kCacheBytes) : 0u;
  const Tables<SMEM> tb = stage_tables<SMEM>(tp.lt, smem_raw + stage_off + kStageBytes, GENERAL);
  if (!SMEM && use_cache) __syncthreads();
  const uint32_t total = tp.n_main + *tp.fork_snapshot;
  const uint32_t stride = gridDim.x * blockDim.x;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
#if HB_ASYNC_STAGE
  // Software pipeline through shared memory: the next ray's state (D, P, Q) is copied global -> shared with
  // cp.async (LDGSTS, no registers in flight) while this ray is computed; every thread reads back only the
  // 48 bytes it requested itself, so cp.async.wait_group is the only synchronisation. The orientation
  // quaternion stays in shared memory until an exit actually needs it.
  const uint32_t stage0 = static_cast<uint32_t>(__cvta_generic_to_shared(smem_raw + stage_off)) + threadIdx.x * 16u;
  uint32_t stage = 0u;
  if (i < total) {
    cp_async16(stage0, tp.D + i);
    cp_async16(stage0 + 4096u, tp.P + i);
    cp_async16(stage0 + 8192u, tp.Q + i);
  }
  cp_async_commit();
  while (i < total) {
    const uint32_t i_next = i + stride;
    const uint32_t cur = stage0 + stage * 12288u, nxt = stage0 + (stage ^ 1u) * 12288u;
    if (i_next < total) {
      cp_async16(nxt, tp.D + i_next);
      cp_async16(nxt + 4096u, tp.P + i_next);
      cp_async16(nxt + 8192u, tp.Q + i_next);
    }
    cp_async_commit();
    cp_async_wait<1>();
    const float4 d4 = lds128(cur), p4 = lds128(cur + 4096u);
#define HB_LOAD_Q() lds128(cur + 8192u)
#else
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
#define HB_LOAD_Q() q
#endif
    const uint32_t bits = __float_as_uint(p4.w);
    const uint32_t face = bits_face(bits);
    if (d4.w >= 0.0f && face != kFaceInvalid) {  // else: terminated ray
      const uint32_t shape = bits_shape(bits);
      const uint32_t meta = tb.meta(shape);
      const AxisRow<SMEM> axes = tb.axes(shape);
      const uint32_t axis_cnt = (meta >> 16) & 255u;
      const bool shape_p4 = (meta & kMetaP4) != 0u;
      float n_idx = tp.wl0.n_idx, inv_n = tp.wl0.inv_n;
      if (tp.wl_cnt != 1u) {
        const float2 nn = __ldg(reinterpret_cast<const float2*>(tp.wl2 + bits_wl(bits)));
        n_idx = nn.x;
        inv_n = nn.y;
      }

      const float4 pl = tb.plane(shape, face);
      const Split s = hit_surface(pl, n_idx, inv_n, d4.x, d4.y, d4.z, d4.w);
      // The child on the far side of the face normally leaves the crystal: classify it here.
      const uint32_t out_child = s.cos_in > 0.0f ? 1u : 0u;  // internal hit: refracted; entry: reflected
      const float ox = out_child ? s.tx : s.rx, oy = out_child ? s.ty : s.ry, oz = out_child ? s.tz : s.rz;
      const float ow = out_child ? s.tw : s.rw;
      const float ix = out_child ? s.rx : s.tx, iy = out_child ? s.ry : s.ty, iz = out_child ? s.rz : s.tz;
      const float iw = out_child ? s.rw : s.tw;
      if (ow >= 0.0f) {
        float nx = 0.f, ny = 0.f, nz = 0.f;
        uint32_t nf = kFaceInvalid;
        if (P4 == 1 || (P4 == 2 && shape_p4)) {
          if (!far_child_surely_exits_p4(axes, pl, p4.x, p4.y, p4.z, ox, oy, oz))
            nf = slab_exit_p4<true>(axes, face, p4.x, p4.y, p4.z, ox, oy, oz, nx, ny, nz);
        } else {
          if (!far_child_surely_exits(axes, axis_cnt, face, pl, p4.x, p4.y, p4.z, ox, oy, oz))
            nf = slab_exit<true>(axes, axis_cnt, face, p4.x, p4.y, p4.z, ox, oy, oz, nx, ny, nz);
        }
        if (nf == kFaceInvalid) {
          emit_exit<GENERAL, MULTI>(tp, i, bits, HB_LOAD_Q(), ox, oy, oz, ow, /*role=*/0u, tb, tally);
        } else if (!LAST) {
          fork_append<GENERAL>(tp, i, bits, HB_LOAD_Q(), nx, ny, nz, ox, oy, oz, ow, nf);  // near-edge leak: both children stay
        }
      }
      if (LAST) {
        // no intersect pass follows the final interaction: classify the inside child here too
        if (iw >= 0.0f && !near_child_surely_hits(axes, P4 == 1 ? 4u : axis_cnt, p4.x, p4.y, p4.z, ix, iy, iz)) {
          float nx, ny, nz;
          uint32_t nf;
          if (P4 == 1 || (P4 == 2 && shape_p4)) nf = slab_exit_p4<false>(axes, face, p4.x, p4.y, p4.z, ix, iy, iz, nx, ny, nz);
          else nf = slab_exit<false>(axes, axis_cnt, face, p4.x, p4.y, p4.z, ix, iy, iz, nx, ny, nz);
          if (nf == kFaceInvalid) emit_exit<GENERAL, MULTI>(tp, i, bits, HB_LOAD_Q(), ix, iy, iz, iw, /*role=*/1u, tb, tally);
        }
      } else {
        tp.D[i] = make_float4(ix, iy, iz, iw);  // iw < 0 (TIR sentinel) terminates the ray
      }
    }
#if HB_ASYNC_STAGE
    stage ^= 1u;
#else
    d4 = d_n;
    p4 = p_n;
    q = q_n;
#endif
    i = i_next;
  }
#undef HB_LOAD_Q
  if (use_cache) cache_flush(tp, tally);
  if (GENERAL && (tp.flags & kFlagStats) && tally.exits != 0ull) {
    atomicAdd(tp.stat_exit_count, tally.exits);
    atomicAdd(tp.stat_w_sum, tally.w_sum);
  }
}

// ------------------------------------------------------------------------------------------------
// intersect kernel: slab exit-face search for the inside child
// ------------------------------------------------------------------------------------------------
template <bool GENERAL, bool SMEM, bool MULTI, int P4 = 0>
__global__ void __launch_bounds__(256, GENERAL ? HB_INTERSECT_MINB_GENERAL : HB_INTERSECT_MINB) intersect_kernel(const TraceParams tp) {
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
        uint32_t nf;
        if (P4 == 1 || (P4 == 2 && (meta & kMetaP4) != 0u))
          nf = slab_exit_p4<false>(tb.axes(shape), face, p4.x, p4.y, p4.z, d4.x, d4.y, d4.z, nx, ny, nz);
        else
          nf = slab_exit<false>(tb.axes(shape), (meta >> 16) & 255u, face, p4.x, p4.y, p4.z, d4.x, d4.y, d4.z, nx, ny, nz);
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
// Fused bounce kernel: one whole surface interaction per launch (the optics and the intersect pass in one), with
// a per-warp EXIT QUEUE in shared memory.
//
// Why a queue. In the split optics kernel a lane whose far-side child leaves the crystal walks the whole emission
// path (orientation matrix, world rotation, visibility, projection, image reduction: ~130 instructions) while the
// lanes of the same warp whose ray was totally reflected (a third of them after the entry interaction) idle:
// ncu counted 19.5 of 32 lanes active per instruction at hit 1. Here a leaving child is only PUSHED (ballot +
// prefix + two shared-memory stores) and the warp emits 32 queued exits at a time with all lanes busy; whatever
// is left when the warp runs out of rays is emitted at the end of the kernel.
//
// Why fused. The near-side child's slab scan needs the same point-to-plane offsets as the far-side child's quick
// classification (bounce_axes_p4 evaluates them once), the child's direction never travels through HBM between
// the two passes, and one launch replaces two: read P, D (32 B) + Q of the exits, write P, D (32 B) per
// ray-bounce instead of 112 B.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kQueueSlots = 96u;                                  // < 32 pending + up to 2 x 32 new per iteration
// float4 (local dir, w) + u32 (tile slot | role << 31). The ray's packed bits (wavelength index, shape) never change
// during a layer, so the emission re-reads them from P[slot] when it needs them (several wavelengths, GENERAL).
constexpr uint32_t kQueueWarpBytes = kQueueSlots * 16u + kQueueSlots * 4u;
constexpr uint32_t kQueueBytes = 8u * kQueueWarpBytes;                 // 8 warps per CTA
// HB_BULK_STAGE = 1: the ray stage-in is a per-WARP 1-D bulk copy (cp.async.bulk global -> shared, completion on an
// mbarrier: SASS UBLKCP + SYNCS) issued by lane 0 -- a warp's 32 rays are 512 contiguous bytes of D and of P --
// instead of one LDGSTS.128 per thread and array. Per-warp barriers keep the warps of a CTA independent. Measured
// A/B in profiles/README.md (the kernel is issue-bound: the copy engine saves about four issue slots per iteration).
#ifndef HB_BULK_STAGE
#define HB_BULK_STAGE 0
#endif
constexpr uint32_t kStage2Bytes = 2u * 2u * 256u * 16u + (HB_BULK_STAGE ? 128u : 0u);  // two stages x (D, P) x 256 threads x 16 B (+ 8 warps x 2 mbarriers)

HB_DEV void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
HB_DEV void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
HB_DEV void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
HB_DEV void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "HB_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra HB_MBAR_DONE;\n"
      "bra HB_MBAR_WAIT;\n"
      "HB_MBAR_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
// Lane 0 of a warp: copy the D and P entries of the warp's rays [first, first + 32) (clipped to `total`) into stage `s`.
HB_DEV void bulk_stage_in(const TraceParams& tp, uint32_t stage_base, uint32_t bar0, uint32_t s, uint32_t first, uint32_t total) {
  if ((threadIdx.x & 31u) == 0u && first < total) {
    const uint32_t bytes = min(32u, total - first) * 16u;
    const uint32_t dst = stage_base + s * 8192u + (threadIdx.x >> 5) * 512u, bar = bar0 + s * 8u;
#if HB_BULK_STAGE == 2
    // Ordering of the warp's earlier LDS of this buffer before the async-proxy write. Those loads have completed (their
    // values were consumed by the previous iteration, which ended in a warp-wide ballot), so the build without the
    // fence (HB_BULK_STAGE = 1) is the one measured for speed; this one (SYNCS.CCTL.IVALL per copy) costs 7 %.
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
    mbar_expect_tx(bar, 2u * bytes);
    bulk_g2s(dst, tp.D + first, bytes, bar);
    bulk_g2s(dst + 4096u, tp.P + first, bytes, bar);
  }
}

HB_DEV void sts128(uint32_t addr, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
HB_DEV void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
HB_DEV uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}

HB_DEV void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

struct ExitQueue {
  uint32_t addr;    // shared-space address of this warp's queue
  uint32_t count;   // warp-uniform
};

// All 32 lanes call this together. `meta0` = tile slot | role << 31.
HB_DEV void queue_push(ExitQueue& xq, bool has, float x, float y, float z, float w, uint32_t meta0) {
  const uint32_t m = __ballot_sync(0xFFFFFFFFu, has);
  if (m == 0u) return;
  if (has) {
    const uint32_t pos = xq.count + __popc(m & ((1u << (threadIdx.x & 31u)) - 1u));
    sts128(xq.addr + pos * 16u, x, y, z, w);
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(xq.addr + kQueueSlots * 16u + pos * 4u), "r"(meta0) : "memory");
  }
  xq.count += __popc(m);
  __syncwarp();
}

// Both children of one interaction in one go (role 0 entries first, then role 1: the order two queue_push calls
// would leave): two ballots, one early-out, one __syncwarp.
HB_DEV void queue_push2(ExitQueue& xq, bool has0, float4 e0, bool has1, float4 e1, uint32_t slot) {
  const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, has0), m1 = __ballot_sync(0xFFFFFFFFu, has1);
  if ((m0 | m1) == 0u) return;
  const uint32_t below = (1u << (threadIdx.x & 31u)) - 1u;
  const uint32_t n0 = __popc(m0);
  if (has0) {
    const uint32_t pos = xq.count + __popc(m0 & below);
    sts128(xq.addr + pos * 16u, e0.x, e0.y, e0.z, e0.w);
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(xq.addr + kQueueSlots * 16u + pos * 4u), "r"(slot) : "memory");
  }
  if (has1) {
    const uint32_t pos = xq.count + n0 + __popc(m1 & below);
    sts128(xq.addr + pos * 16u, e1.x, e1.y, e1.z, e1.w);
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(xq.addr + kQueueSlots * 16u + pos * 4u), "r"(slot | 0x80000000u) : "memory");
  }
  xq.count += n0 + __popc(m1);
  __syncwarp();
}

// Emit queued exits, 32 at a time (`all`: whatever is left, with the lanes that still have an entry), in two
// converged stages:
//   stage 1  orientation matrix, world rotation, filter / gate / record (emit_world), then the exact visibility
//            cull of render 0 (project_culls) -- for a one-hemisphere view about half of the exits end here;
//   stage 2  the survivors are compacted into the warp's second queue (world direction, weight, wavelength) and
//            projected + reduced into the image 32 at a time (emit_project).
// MULTI kernels (extra renders, colour lanes: per-exit masks and several lenses) project straight from stage 1.
struct ExitQueue2 {
  uint32_t addr;
  uint32_t count;
};
constexpr uint32_t kQueue2Slots = 64u;                                      // < 32 pending + up to 32 new per stage-1 pass
constexpr uint32_t kQueue2WarpBytes = kQueue2Slots * 16u + kQueue2Slots * 4u;  // float4 (world dir, w) + u32 wavelength index
#if HB_EXIT_STAGES == 2
constexpr uint32_t kQueue2Bytes = 8u * kQueue2WarpBytes;
#else
constexpr uint32_t kQueue2Bytes = 0u;
#endif

template <bool MULTI>
HB_DEV void queue2_drain(ExitQueue2& q2, bool all, const TraceParams& tp, Tally& tally) {
  const uint32_t lane = threadIdx.x & 31u;
  while (q2.count >= 32u || (all && q2.count != 0u)) {
    const uint32_t n = min(q2.count, 32u), first = q2.count - n;
    if (lane < n) {
      const float4 e = lds128(q2.addr + (first + lane) * 16u);
      uint32_t wl_i;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wl_i) : "r"(q2.addr + kQueue2Slots * 16u + (first + lane) * 4u));
      emit_project<MULTI>(tp, wl_i, e.x, e.y, e.z, e.w, 0ull, tally);
    }
    q2.count = first;
    __syncwarp();
  }
}

template <bool GENERAL, bool MULTI, typename TablesT>
HB_DEV void queue_drain(ExitQueue& xq, ExitQueue2& q2, bool all, const TraceParams& tp, const TablesT& tb, Tally& tally) {
  const uint32_t lane = threadIdx.x & 31u;
  while (xq.count >= 32u || (all && xq.count != 0u)) {
    const uint32_t n = min(xq.count, 32u), first = xq.count - n;
    bool keep = false;
    float wx = 0.f, wy = 0.f, wz = 0.f, w = 0.f;
    uint32_t wl_i = 0u;
    if (lane < n) {
      const float4 e = lds128(xq.addr + (first + lane) * 16u);
      uint32_t m0;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(m0) : "r"(xq.addr + kQueueSlots * 16u + (first + lane) * 4u));
      const uint32_t slot = m0 & 0x7FFFFFFFu;
      uint32_t bits = 0u;  // wavelength index 0, shape 0: all a single-wavelength, non-general session needs
      if (GENERAL || tp.wl_cnt != 1u) bits = __float_as_uint(tp.P[slot].w);
#if HB_EXIT_STAGES == 2
      if constexpr (!MULTI) {
        uint64_t mask;
        w = e.w;
        wl_i = bits_wl(bits);
        keep = emit_world<GENERAL, MULTI>(tp, slot, bits, tp.Q[slot], e.x, e.y, e.z, e.w, m0 >> 31, tb, tally, wx, wy, wz, mask) &&
               !project_culls(tp.proj, wx, wy, wz);
      } else
#endif
      {
        emit_exit<GENERAL, MULTI>(tp, slot, bits, tp.Q[slot], e.x, e.y, e.z, e.w, m0 >> 31, tb, tally);
      }
    }
    xq.count = first;
    __syncwarp();
#if HB_EXIT_STAGES == 2
    if constexpr (!MULTI) {
      const uint32_t km = __ballot_sync(0xFFFFFFFFu, keep);
      if (km != 0u) {
        if (keep) {
          const uint32_t pos = q2.count + __popc(km & ((1u << lane) - 1u));
          sts128(q2.addr + pos * 16u, wx, wy, wz, w);
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(q2.addr + kQueue2Slots * 16u + pos * 4u), "r"(wl_i) : "memory");
        }
        q2.count += __popc(km);
        __syncwarp();
        queue2_drain<MULTI>(q2, false, tp, tally);
      }
    }
#endif
  }
#if HB_EXIT_STAGES == 2
  if constexpr (!MULTI) {
    if (all) queue2_drain<MULTI>(q2, true, tp, tally);
  }
#endif
}

// The last CTA to retire publishes the number of fork rays appended so far: the next bounce launch traces
// [0, n_main + snapshot). (The split pipeline lets its intersect pass take the snapshot.)
HB_DEV void publish_fork_snapshot(const TraceParams& tp) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t prev = atomicAdd(tp.done_count, 1u);
    if (prev + 1u == gridDim.x) {
      __threadfence();
      *tp.fork_snapshot = min(atomicAdd(tp.fork_count, 0u), tp.fork_cap);
      *tp.done_count = 0u;
    }
  }
}

// One interaction of one live ray in registers. Outputs: up to two leaving children (e0: far side, role 0;
// e1: near side, role 1) and the ray's next state: d_out always (w < 0: the ray ends), p_out when `moved`
// (the near child reached a face). The caller stores them (nothing is stored after the LAST interaction).
template <bool GENERAL, bool LAST, bool SMEM, int P4>
HB_DEV void bounce_ray(const TraceParams& tp, const Tables<SMEM>& tb, uint32_t i, float4 p4, float4 d4, bool& has0,
                       float4& e0, bool& has1, float4& e1, float4& d_out, float4& p_out, bool& moved) {
  const uint32_t bits = __float_as_uint(p4.w);
  const uint32_t face = bits_face(bits);
  const uint32_t shape = bits_shape(bits);
  const uint32_t meta = tb.meta(shape);
  const AxisRow<SMEM> axes = tb.axes(shape);
  const uint32_t axis_cnt = (meta >> 16) & 255u;
  const bool shape_p4 = P4 == 1 || (P4 == 2 && (meta & kMetaP4) != 0u);
  float n_idx = tp.wl0.n_idx, inv_n = tp.wl0.inv_n;
  if (tp.wl_cnt != 1u) {
    const float2 nn = __ldg(reinterpret_cast<const float2*>(tp.wl2 + bits_wl(bits)));
    n_idx = nn.x;
    inv_n = nn.y;
  }
  const float4 pl = tb.plane(shape, face);
  const Split s = hit_surface(pl, n_idx, inv_n, d4.x, d4.y, d4.z, d4.w);
  const bool internal = s.cos_in > 0.0f;  // internal hit: the refracted child leaves; entry: the reflected one
  const float ox = internal ? s.tx : s.rx, oy = internal ? s.ty : s.ry, oz = internal ? s.tz : s.rz;
  const float ow = internal ? s.tw : s.rw;
  const float ix = internal ? s.rx : s.tx, iy = internal ? s.ry : s.ty, iz = internal ? s.rz : s.tz;
  const float iw = internal ? s.rw : s.tw;

  // ---- both children against the crystal ----
  bool far_exits;
  float nx = p4.x, ny = p4.y, nz = p4.z;   // near child's next point
  uint32_t nf = kFaceInvalid;              // ... and face
  bool near_done = false;                  // LAST: the near child provably stays inside, nothing to do
  if (shape_p4) {
    if (LAST) {
      bool near_hits;
      last_axes_p4(axes, pl, p4.x, p4.y, p4.z, ox, oy, oz, ix, iy, iz, far_exits, near_hits);
      near_done = iw < 0.0f || near_hits;
      if (!near_done) nf = slab_exit_p4<false>(axes, face, p4.x, p4.y, p4.z, ix, iy, iz, nx, ny, nz);
    } else {
      nf = bounce_axes_p4(axes, face, pl, p4.x, p4.y, p4.z, ox, oy, oz, ix, iy, iz, far_exits, nx, ny, nz);
    }
  } else {
    if (LAST) {
      far_exits = far_child_surely_exits(axes, axis_cnt, face, pl, p4.x, p4.y, p4.z, ox, oy, oz);
      near_done = iw < 0.0f || near_child_surely_hits(axes, axis_cnt, p4.x, p4.y, p4.z, ix, iy, iz);
      if (!near_done) nf = slab_exit<false>(axes, axis_cnt, face, p4.x, p4.y, p4.z, ix, iy, iz, nx, ny, nz);
    } else {
      nf = bounce_axes(axes, axis_cnt, face, pl, p4.x, p4.y, p4.z, ox, oy, oz, ix, iy, iz, far_exits, nx, ny, nz);
    }
  }
  // far-side child: leaves (the common case), or stays inside within ~1e-7 of an edge (fork ray)
  if (ow >= 0.0f) {
    uint32_t ff = kFaceInvalid;
    float fx = 0.f, fy = 0.f, fz = 0.f;
    if (!far_exits) {
      if (shape_p4) ff = slab_exit_p4<true>(axes, face, p4.x, p4.y, p4.z, ox, oy, oz, fx, fy, fz);
      else ff = slab_exit<true>(axes, axis_cnt, face, p4.x, p4.y, p4.z, ox, oy, oz, fx, fy, fz);
    }
    if (ff == kFaceInvalid) {
      has0 = true;
      e0 = make_float4(ox, oy, oz, ow);
    } else if (!LAST) {
      fork_append<GENERAL>(tp, i, bits, tp.Q[i], fx, fy, fz, ox, oy, oz, ow, ff, /*advanced=*/0u);
    }
  }
  // near-side child
  moved = false;
  d_out = make_float4(ix, iy, iz, iw);  // iw < 0: TIR sentinel, the ray ends (entry-side refraction never does)
  p_out = p4;
  if (LAST) {
    if (!near_done && nf == kFaceInvalid) {
      has1 = true;
      e1 = make_float4(ix, iy, iz, iw);
    }
  } else if (iw >= 0.0f) {
    if (nf == kFaceInvalid) {
      // the inside child found no face: it is outgoing (CollectData branch 1) and the ray ends here
      has1 = true;
      e1 = make_float4(ix, iy, iz, iw);
      d_out.w = -1.0f;
    } else {
      moved = true;
      p_out = make_float4(nx, ny, nz, __uint_as_float(bits_with_face(bits, nf)));
      if (GENERAL && (tp.flags & kFlagPath) && tp.hit + 1u < tp.max_hits)
        tp.path[static_cast<size_t>(tp.hit + 1u) * tp.cap + i] = static_cast<uint8_t>(nf);
    }
  }
}

#ifndef HB_BOUNCE_MINB
#define HB_BOUNCE_MINB 4
#endif
#ifndef HB_PREFETCH_Q
#define HB_PREFETCH_Q 1
#endif
// General (filter / gate / record / colour) instantiations: 3 CTAs per SM = 80 registers, no spills (4 CTAs at 64
// registers spill 30-170 bytes); measured +1.5 % on config 3 and +2.2 % on config 4.
#ifndef HB_BOUNCE_MINB_GENERAL
#define HB_BOUNCE_MINB_GENERAL 3
#endif
// dynamic shared memory: [pixel cache] [ray staging D, P x 2] [exit queues] [crystal tables]
template <bool GENERAL, bool LAST, bool SMEM, bool MULTI, int P4 = 0>
__global__ void __launch_bounds__(256, GENERAL ? HB_BOUNCE_MINB_GENERAL : HB_BOUNCE_MINB) bounce_kernel(const TraceParams tp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Tally tally;
  const bool use_cache = (tp.flags & kFlagPixelCache) != 0u;
  if (use_cache) cache_init(tally, smem_raw);
  const uint32_t stage_off = use_cache ? static_cast<uint32_t>(kCacheBytes) : 0u;
  const Tables<SMEM> tb = stage_tables<SMEM>(tp.lt, smem_raw + stage_off + kStage2Bytes + kQueueBytes + kQueue2Bytes, GENERAL);
  if (!SMEM && use_cache) __syncthreads();
  const uint32_t total = tp.n_main + *tp.fork_snapshot;
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t smem_base = static_cast<uint32_t>(__cvta_generic_to_shared(smem_raw));
  ExitQueue xq{ smem_base + stage_off + kStage2Bytes + (threadIdx.x >> 5) * kQueueWarpBytes, 0u };
  ExitQueue2 q2{ smem_base + stage_off + kStage2Bytes + kQueueBytes + (threadIdx.x >> 5) * kQueue2WarpBytes, 0u };
  const uint32_t stage0 = smem_base + stage_off + threadIdx.x * 16u;
  uint32_t stage = 0u;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t warp_first = i - (threadIdx.x & 31u);  // the loop runs while the WARP has a ray: pushes are warp-wide
#if HB_BULK_STAGE
  const uint32_t stage_base = smem_base + stage_off;
  const uint32_t bar0 = stage_base + 16384u + (threadIdx.x >> 5) * 16u;  // this warp's two mbarriers (one per stage)
  if ((threadIdx.x & 31u) == 0u) {
    mbar_init(bar0, 1u);
    mbar_init(bar0 + 8u, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  bulk_stage_in(tp, stage_base, bar0, 0u, warp_first, total);
  uint32_t iter = 0u;
#else
  if (i < total) {
    cp_async16(stage0, tp.D + i);
    cp_async16(stage0 + 4096u, tp.P + i);
  }
  cp_async_commit();
#endif
  // One emission site: the loop body runs once more after the warp's last ray to flush the queue, so the (large)
  // emission code is instantiated once and nothing of the trace state is live across it.
  for (;;) {
    const bool more = warp_first < total;
    if (more) {
      const uint32_t i_next = i + stride;
      const uint32_t cur = stage0 + stage * 8192u;
#if HB_BULK_STAGE
      bulk_stage_in(tp, stage_base, bar0, stage ^ 1u, warp_first + stride, total);
      mbar_wait(bar0 + stage * 8u, (iter >> 1) & 1u);
      iter++;
#else
      const uint32_t nxt = stage0 + (stage ^ 1u) * 8192u;
      if (i_next < total) {
        cp_async16(nxt, tp.D + i_next);
        cp_async16(nxt + 4096u, tp.P + i_next);
      }
      cp_async_commit();
      cp_async_wait<1>();
#endif
      bool has0 = false, has1 = false;
      float4 e0 = make_float4(0.f, 0.f, 0.f, 0.f), e1 = e0;
      uint32_t bits = 0u;
      if (i < total) {
        const float4 d4 = lds128(cur), p4 = lds128(cur + 4096u);
        bits = __float_as_uint(p4.w);
        if (d4.w >= 0.0f && bits_face(bits) != kFaceInvalid) {  // else: terminated ray
          float4 d_out, p_out;
          bool moved;
          bounce_ray<GENERAL, LAST, SMEM, P4>(tp, tb, i, p4, d4, has0, e0, has1, e1, d_out, p_out, moved);
          if (!LAST) {
            tp.D[i] = d_out;
            if (moved) tp.P[i] = p_out;
          }
        }
      }
#if HB_PREFETCH_Q
      if (has0 || has1) prefetch_l1(tp.Q + i);  // the emission (some iterations later, another lane) reads the orientation
#endif
      queue_push2(xq, has0, e0, has1, e1, i);
      stage ^= 1u;
      i = i_next;
      warp_first += stride;
    }
    if (xq.count >= 32u || !more) queue_drain<GENERAL, MULTI>(xq, q2, !more, tp, tb, tally);
    if (!more) break;
  }
  if (use_cache) cache_flush(tp, tally);
  if (GENERAL && (tp.flags & kFlagStats) && tally.exits != 0ull) {
    atomicAdd(tp.stat_exit_count, tally.exits);
    atomicAdd(tp.stat_w_sum, tally.w_sum);
  }
  if (!LAST) publish_fork_snapshot(tp);
}

// ------------------------------------------------------------------------------------------------
// root generation / layer transit
// ------------------------------------------------------------------------------------------------
struct EntryFaces;
struct GenShared;

constexpr uint32_t kFastGroups = 8;  // prism: 8 faces, weights kept in registers

// Returns the chosen fan triangle. `ef` may point to shared or global memory (warp-uniform address).
// Level 1 walks the cumulative weights c[g] = c[g-1] + w[g] exactly as a sequential categorical would (first g with
// c[g] > target; the residual is target - c[g-1]); because c is non-decreasing "first g with c[g] > target" is the
// NUMBER of g with c[g] <= target, which needs no found-flag chain: per group one compare and three predicated
// moves. Level 2 counts the cumulative area fractions below the residual ratio from one packed 16-byte load.
HB_DEV uint32_t pick_entry_triangle(Stream& s, const EntryFaces* ef, float dx, float dy, float dz) {
  const uint32_t ng = ef->group_cnt;
  const float u_cat = s.next();
  // not found (rounding pushed the target up to the total): last group, residual 0
  uint32_t sel = ng - 1u;
  float resid = 0.0f, w_sel = 0.0f, total = 0.0f;
  if (ng <= kFastGroups) {
    float w[kFastGroups], c[kFastGroups];
#pragma unroll
    for (uint32_t g = 0; g < kFastGroups; g++) {
      w[g] = 0.0f;
      if (g < ng) {
        const float4 na = ef->na[g];
        w[g] = fmaxf(-(dx * na.x + dy * na.y + dz * na.z) * na.w, 0.0f);
        total += w[g];
      }
      c[g] = total;
    }
    if (!(total > 0.0f)) return 0u;
    const float target = u_cat * total;
    uint32_t below = 0u;
    float lo = 0.0f;
#pragma unroll
    for (int g = static_cast<int>(kFastGroups) - 1; g >= 0; g--) {  // descending: the last write of w_sel is the first c[g] > target
      const bool le = c[g] <= target;
      w_sel = le ? w_sel : w[g];
      lo = le ? fmaxf(lo, c[g]) : lo;
      below += le ? 1u : 0u;
    }
    sel = min(below, ng - 1u);   // groups past ng repeat the total: they only count when nothing was found
    resid = below < ng ? target - lo : 0.0f;
    w_sel = below < ng ? w_sel : 0.0f;
  } else {
    for (uint32_t g = 0; g < ng; g++) {
      const float4 na = ef->na[g];
      total += fmaxf(-(dx * na.x + dy * na.y + dz * na.z) * na.w, 0.0f);
    }
    if (!(total > 0.0f)) return 0u;
    const float target = u_cat * total;
    float cum = 0.0f;
    for (uint32_t g = 0; g < ng; g++) {
      const float4 na = ef->na[g];
      const float wg = fmaxf(-(dx * na.x + dy * na.y + dz * na.z) * na.w, 0.0f);
      const float c1 = cum + wg;
      if (c1 > target) {
        sel = g;
        resid = target - cum;
        w_sel = wg;
        break;
      }
      cum = c1;
    }
  }
  const float r = w_sel > 0.0f ? dvd_nr(resid, w_sel) : 0.0f;  // normal-range quotient: == IEEE division
  const float4 pk = ef->pick[sel];
  const uint32_t pbits = __float_as_uint(pk.w);
  const uint32_t t0 = pbits & 255u, cnt = pbits >> 8;
  if (cnt <= 4u) return t0 + (r >= pk.x ? 1u : 0u) + (r >= pk.y ? 1u : 0u) + (r >= pk.z ? 1u : 0u);
  uint32_t tri = t0;
  for (uint32_t j = 0; j + 1u < cnt; j++) {
    if (r >= ef->cum[t0 + j]) tri = t0 + j + 1u;
  }
  return tri;
}

// Uniform point in the chosen triangle (SampleTrianglePoint, geo3d.cpp:114-190; sample_triangle, pcg_shared.h:493-509).
HB_DEV void sample_entry_faces(Stream& s, const EntryFaces* ef, const HbCrystalTables* tab, float dx, float dy, float dz,
                               float& px, float& py, float& pz, uint32_t& face) {
  const uint32_t tri = pick_entry_triangle(s, ef, dx, dy, dz);
  float u = s.next(), v = s.next();
  if (u + v > 1.0f) {
    u = 1.0f - u;
    v = 1.0f - v;
  }
  const float* tv = tab->tri_v[tri];
  px = u * (tv[3] - tv[0]) + v * (tv[6] - tv[0]) + tv[0];
  py = u * (tv[4] - tv[1]) + v * (tv[7] - tv[1]) + tv[1];
  pz = u * (tv[5] - tv[2]) + v * (tv[8] - tv[2]) + tv[2];
  face = tab->tri_face[tri];
}

struct GenShared {
  float lut[3 * HB_LUT_NODES];
  HbCrystalTables shape0;        // single-shape populations: entry fan table staged on chip
  EntryFaces ef0;                // ... and its face groups
};
constexpr uint32_t kGenSharedBytes = (static_cast<uint32_t>(sizeof(GenShared)) + 15u) & ~15u;

HB_DEV void stage_gen_shared(const GenParams& gp, GenShared* gs) {
  if (gp.axis.lat_path == HB_LAT_LUT) {
    for (uint32_t i = threadIdx.x; i < 3 * HB_LUT_NODES; i += blockDim.x) gs->lut[i] = gp.lut[i];
  }
  if (gp.shape_cnt == 1u) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(gp.shapes);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&gs->shape0);
    for (uint32_t i = threadIdx.x; i < sizeof(HbCrystalTables) / 4; i += blockDim.x) dst[i] = src[i];
    const uint4* esrc = reinterpret_cast<const uint4*>(gp.entry_faces);
    uint4* edst = reinterpret_cast<uint4*>(&gs->ef0);
    for (uint32_t i = threadIdx.x; i < sizeof(EntryFaces) / 16; i += blockDim.x) edst[i] = esrc[i];
  }
}

// Root ray k of a launch (InitRay_*, simulator.cpp:133-339 with the counter-based streams of pcg_shared.h): wavelength
// draw / continuation gather, orientation, sun-cone direction, shape pick, entry point. Returns the ray state.
constexpr uint32_t kNoSource = 0xFFFFFFFFu;
template <bool TRANSIT>
HB_DEV void gen_root(const GenParams& gp, const GenShared* gs, uint32_t k, float4& p_out, float4& d_out, float4& q_out,
                     uint32_t src_known = kNoSource) {
  const uint32_t lo = gp.idx_lo + k;
  const uint32_t hi = gp.idx_hi + (lo < gp.idx_lo ? 1u : 0u);
  const uint32_t s0 = seed_with_high(gp.seed, hi);
  uint32_t wl_i = 0u;
  float wx, wy, wz, weight;
  if (TRANSIT) {
    uint32_t src = src_known;
    if (src == kNoSource) {
      src = gp.cont_first + k;
      if (gp.shuffle) src = feistel(src, gp.cont_n, gp.shuffle_seed);
    }
    const float4* rec = reinterpret_cast<const float4*>(gp.cont + src);
    const float4 c = __ldg(rec);
    const uint4 m = __ldg(reinterpret_cast<const uint4*>(rec) + 1);
    wx = c.x;
    wy = c.y;
    wz = c.z;
    weight = c.w;
    wl_i = m.x & 255u;
    if (gp.M != nullptr) gp.M[gp.slot0 + k] = static_cast<uint64_t>(m.z) | (static_cast<uint64_t>(m.w) << 32);
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
    const float rr2 = 1.0f - x * x;
    const float rr = rr2 > 0.0f ? sqrt_nr(rr2) : 0.0f;  // == sqrtf(fmaxf(rr2, 0)): rr2 is 0 or at least an ulp of 1
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
  // Geometry clock: one shape of the pool serves a block of 32 consecutive ray indices, as on the reference's
  // CPU path (kSmallBatchRayNum, simulator.hpp:144-151). A warp therefore reads ONE shape's tables in every
  // kernel of the hit loop (uniform addresses: broadcast loads) instead of 32 different ones. The block's shape is
  // drawn (measured against a sequential walk through the pool in runs of 32 or 256 rays: the drawn assignment is
  // 3 % faster on 256-shape pools -- tables of neighbouring CTAs then do not collide in the same L2 lines).
  uint32_t sh = 0u;
  if (gp.shape_cnt > 1u) {
    sh = min(static_cast<uint32_t>(draw(s0 ^ kNonceShape, lo >> 5, 0u) * static_cast<float>(gp.shape_cnt)), gp.shape_cnt - 1u);
  }
  const HbCrystalTables* tab = gp.shape_cnt == 1u ? &gs->shape0 : gp.shapes + sh;
  const EntryFaces* ef = gp.shape_cnt == 1u ? &gs->ef0 : gp.entry_faces + sh;
  float px = 0.0f, py = 0.0f, pz = 0.0f;
  uint32_t face = kFaceInvalid;
  if (tab->subtri_cnt == 0u) {
    weight = -1.0f;  // degenerate crystal: nothing to trace (zero-weight discard, simulator.cpp:149-159)
  } else {
    if (ef->group_cnt != 0u) sample_entry_faces(s, ef, tab, dx, dy, dz, px, py, pz, face);
    else sample_entry(s, tab, dx, dy, dz, px, py, pz, face);
  }
  p_out = make_float4(px, py, pz, __uint_as_float(pack_bits(face, wl_i, gp.shape_base + sh, 0u)));
  d_out = make_float4(dx, dy, dz, weight);
  q_out = q;
}

// Transit generator: the Feistel-permuted sources of kTransitBatch consecutive iterations of a warp are resolved
// together, as a pool of 32 x kTransitBatch items the lanes draw from. Cycle-walking needs a geometric number of
// passes per item (3.4 on average when the pool is just above a quarter of the 2^bits domain), so a warp that walks
// its 32 items in lock step waits for the slowest lane -- ncu: 9 to 14 of 32 lanes active in 41 % of the kernel's
// instructions. Here a lane whose item has landed below n takes the next unresolved item, and every pass runs with
// (almost) all lanes (512 items per pool: the tail, where the last items finish their walks, stays near a tenth of
// the passes); the sources go through shared memory to the lanes that generate the rays.
constexpr uint32_t kTransitBatch = 16u;
constexpr uint32_t kTransitSrcBytes = 8u * 32u * kTransitBatch * 4u;   // 8 warps per CTA

HB_DEV void resolve_sources(const GenParams& gp, uint32_t k_warp, uint32_t stride, uint32_t* src_out) {
  const uint32_t lane = threadIdx.x & 31u;
  const FeistelDomain fd = feistel_domain(gp.cont_n);
  // item t: iteration t / 32 of the batch, lane t % 32; items past the end of the launch are not issued
  uint32_t total = 0u;
#pragma unroll
  for (uint32_t c = 0; c < kTransitBatch; c++) {
    const uint32_t first = k_warp + c * stride;
    total += first < gp.count ? min(32u, gp.count - first) : 0u;   // only the last iteration of a launch is ragged
  }
  uint32_t next = 0u;              // warp-uniform: items handed out so far
  uint32_t item = 0xFFFFFFFFu, cur = 0u, walks = 0u;
  for (;;) {
    const bool need = item == 0xFFFFFFFFu;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, need);
    const uint32_t avail = total - next;
    if (need) {
      const uint32_t r = __popc(m & ((1u << lane) - 1u));
      if (r < avail) {
        item = next + r;
        cur = gp.cont_first + k_warp + (item >> 5) * stride + (item & 31u);
        walks = 0u;
      }
    }
    next += min(static_cast<uint32_t>(__popc(m)), avail);
    const bool active = item != 0xFFFFFFFFu;
    if (__ballot_sync(0xFFFFFFFFu, active) == 0u) break;
    if (active) {
      cur = feistel_walk(cur, fd, gp.shuffle_seed);
      walks++;
      if (cur < gp.cont_n || walks == kFeistelMaxWalks) {
        src_out[item] = cur < gp.cont_n ? cur : cur % gp.cont_n;
        item = 0xFFFFFFFFu;
      }
    }
  }
  __syncwarp();
}

// Five CTAs per SM (48 registers, no spills) for both generators: the transit form settles at 58 registers without
// the bound, which costs it a fifth of its resident warps.
#ifndef HB_GEN_MINB
#define HB_GEN_MINB 5
#endif
template <bool TRANSIT>
__global__ void __launch_bounds__(256, HB_GEN_MINB) gen_kernel(const GenParams gp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GenShared* gs = reinterpret_cast<GenShared*>(smem_raw);
  stage_gen_shared(gp, gs);
  __syncthreads();
  const uint32_t stride = gridDim.x * blockDim.x;
  if (TRANSIT && gp.shuffle && gp.cont_n > 2u) {
    uint32_t* src_warp = reinterpret_cast<uint32_t*>(smem_raw + kGenSharedBytes) + (threadIdx.x >> 5) * (32u * kTransitBatch);
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t k_warp = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); k_warp < gp.count; k_warp += kTransitBatch * stride) {
      resolve_sources(gp, k_warp, stride, src_warp);
#pragma unroll 1
      for (uint32_t c = 0; c < kTransitBatch; c++) {
        const uint32_t k = k_warp + c * stride + lane;
        if (k < gp.count) {
          float4 p4, d4, q;
          gen_root<TRANSIT>(gp, gs, k, p4, d4, q, src_warp[c * 32u + lane]);
          const uint32_t slot = gp.slot0 + k;
          gp.P[slot] = p4;
          gp.D[slot] = d4;
          gp.Q[slot] = q;
          if ((gp.flags & kFlagPath) && gp.path != nullptr) gp.path[slot] = static_cast<uint8_t>(bits_face(__float_as_uint(p4.w)));
        }
      }
      __syncwarp();
    }
    return;
  }
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < gp.count; k += stride) {
    float4 p4, d4, q;
    gen_root<TRANSIT>(gp, gs, k, p4, d4, q);
    const uint32_t slot = gp.slot0 + k;
    gp.P[slot] = p4;
    gp.D[slot] = d4;
    gp.Q[slot] = q;
    if ((gp.flags & kFlagPath) && gp.path != nullptr) gp.path[slot] = static_cast<uint8_t>(bits_face(__float_as_uint(p4.w)));
  }
}

// ------------------------------------------------------------------------------------------------
// Root generation fused with the ENTRY interaction (InitRay_* + the first pass of the hit loop,
// simulator.cpp:133-259,1308-1336): the root never travels through HBM between the generator and hit 0 --
// the state after the entry interaction is written once (P, D, Q: 48 B per root instead of 48 written +
// 32..48 read + 32 written), and a session needs one launch less. The external reflection goes through the
// same per-warp exit queue as in bounce_kernel.
// dynamic shared memory: [GenShared] [pixel cache] [exit queues] [crystal tables]
// ------------------------------------------------------------------------------------------------
template <bool TRANSIT, bool GENERAL, bool SMEM, bool MULTI, int P4 = 0>
__global__ void __launch_bounds__(256, 4) genbounce_kernel(const GenParams gp, const TraceParams tp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GenShared* gs = reinterpret_cast<GenShared*>(smem_raw);
  stage_gen_shared(gp, gs);
  Tally tally;
  const bool use_cache = (tp.flags & kFlagPixelCache) != 0u;
  if (use_cache) cache_init(tally, smem_raw + kGenSharedBytes);
  const uint32_t q_off = kGenSharedBytes + (use_cache ? static_cast<uint32_t>(kCacheBytes) : 0u);
  const Tables<SMEM> tb = stage_tables<SMEM>(tp.lt, smem_raw + q_off + kQueueBytes + kQueue2Bytes, GENERAL);
  if (!SMEM) __syncthreads();
  const uint32_t smem_base = static_cast<uint32_t>(__cvta_generic_to_shared(smem_raw));
  ExitQueue xq{ smem_base + q_off + (threadIdx.x >> 5) * kQueueWarpBytes, 0u };
  ExitQueue2 q2{ smem_base + q_off + kQueueBytes + (threadIdx.x >> 5) * kQueue2WarpBytes, 0u };
  const uint32_t stride = gridDim.x * blockDim.x;
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t warp_first = k - (threadIdx.x & 31u);
  for (;;) {
    const bool more = warp_first < gp.count;
    if (more) {
      bool has0 = false, has1 = false;
      float4 e0 = make_float4(0.f, 0.f, 0.f, 0.f), e1 = e0;
      uint32_t bits = 0u;
      const uint32_t slot = gp.slot0 + k;
      if (k < gp.count) {
        float4 p4, d4, q;
        gen_root<TRANSIT>(gp, gs, k, p4, d4, q);
        bits = __float_as_uint(p4.w);
        gp.Q[slot] = q;
        if (GENERAL && (gp.flags & kFlagPath) && gp.path != nullptr) gp.path[slot] = static_cast<uint8_t>(bits_face(bits));
        if (d4.w >= 0.0f && bits_face(bits) != kFaceInvalid) {
          float4 d_out, p_out;
          bool moved;
          bounce_ray<GENERAL, false, SMEM, P4>(tp, tb, slot, p4, d4, has0, e0, has1, e1, d_out, p_out, moved);
          d4 = d_out;
          p4 = p_out;
        }
        gp.P[slot] = p4;
        gp.D[slot] = d4;
      }
      queue_push(xq, has0, e0.x, e0.y, e0.z, e0.w, slot);
      queue_push(xq, has1, e1.x, e1.y, e1.z, e1.w, slot | 0x80000000u);
      k += stride;
      warp_first += stride;
    }
    if (xq.count >= 32u || !more) queue_drain<GENERAL, MULTI>(xq, q2, !more, tp, tb, tally);
    if (!more) break;
  }
  if (use_cache) cache_flush(tp, tally);
  if (GENERAL && (tp.flags & kFlagStats) && tally.exits != 0ull) {
    atomicAdd(tp.stat_exit_count, tally.exits);
    atomicAdd(tp.stat_w_sum, tally.w_sum);
  }
  publish_fork_snapshot(tp);
}

#ifndef HB_TU  // non-template kernels: defined once, in the engine translation unit (hb_tu.cu slices skip them)
// ------------------------------------------------------------------------------------------------
// Stochastic geometry pool on the device. One thread per shape: draw the shape scalars of the
// population's CrystalConfig with the counter-based RNG (MakeCrystal's order, simulator.cpp:405-450: heights
// first, then the six face distances), build the crystal tables with the SAME code the host builders run
// (hb_geometry.h) and derive everything the trace kernels read (derive_shape_tables). The pool is overwritten
// in place, so a fresh pool per session costs one small launch and no host work.
// ------------------------------------------------------------------------------------------------
struct ShapeGenParams {
  HbCrystalDesc desc;
  double a1, a2;               // pyramid slopes (sqrt3/4)/tan(alpha) from the host, <= 0: segment absent
  uint32_t seed;               // session-independent geometry seed ^ kNonceGeom
  uint32_t draw_base;          // stream index of shape 0
  uint32_t count;
  uint32_t pop;                // population index inside the layer (meta word)
  HbCrystalTables* shapes;     // the population's slices of the layer tables
  float4* planes;
  float4* axes;
  uint32_t* meta;
  uint8_t* fn;
  EntryFaces* ef;
  float* scalars;              // [count][10] h1, h2, h3, d0..d5, builder status (parity export)
  uint32_t* flags;             // [0]: number of shapes that are NOT P4, [1]: number of builder errors
};

__global__ void __launch_bounds__(64) resample_shapes_kernel(const ShapeGenParams sp) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= sp.count) return;
  Stream s{ sp.seed, sp.draw_base + k, 0u };
  float hgt[3], dist[6];
  hbg::SampleShapeScalars(sp.desc, [&](const HbDist& d) { return get_dist(s, d.type, d.center, d.spread); }, hgt, dist);
  HbCrystalTables* t = sp.shapes + k;
  const int rc = sp.desc.kind == 0u ? hbg::MakePrism(hgt[0], dist, t)
                                    : hbg::MakePyramidFromSlopes(sp.a1, sp.a2, hgt[0], hgt[1], hgt[2], dist, t);
  if (rc != HB_OK) {
    memset(t, 0, sizeof(*t));  // degenerate crystal: rays of this shape are discarded (zero sub-triangles)
    atomicAdd(sp.flags + 1, 1u);
  }
  const bool p4 = derive_shape_tables(*t, sp.pop, sp.planes + k * HB_MAX_FACES, sp.fn + k * HB_MAX_FACES,
                                      sp.axes + k * HB_MAX_FACES * 2u, sp.meta + k, sp.ef + k);
  if (!p4) atomicAdd(sp.flags + 0, 1u);
  if (sp.scalars != nullptr) {
    float* o = sp.scalars + k * 10u;
    o[0] = hgt[0];
    o[1] = hgt[1];
    o[2] = hgt[2];
    for (int i = 0; i < 6; i++) o[3 + i] = dist[i];
    o[9] = static_cast<float>(rc);
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

// ---- multi-GPU frame end ---------------------------------------------------------------------------------
// Cross-process reduce (NCCL): the fp64 master is packed to fp32 (X, Y, Z, landed) -- the precision of the frame the
// caller reads anyway -- so 16 instead of 32 bytes per pixel cross NVLink; after the reduce the root's master is
// REPLACED by the sum and every other rank's master is zero, which makes a repeated reduce harmless.
__global__ void __launch_bounds__(256) pack_master_kernel(const double4* master, float4* out, uint32_t pixels) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += gridDim.x * blockDim.x) {
    const double4 v = master[i];
    out[i] = make_float4(static_cast<float>(v.x), static_cast<float>(v.y), static_cast<float>(v.z), static_cast<float>(v.w));
  }
}
__global__ void __launch_bounds__(256) unpack_master_kernel(const float4* in, double4* master, uint32_t pixels) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += gridDim.x * blockDim.x) {
    const float4 v = in[i];
    master[i] = make_double4(static_cast<double>(v.x), static_cast<double>(v.y), static_cast<double>(v.z), static_cast<double>(v.w));
  }
}
// In-process merge (one host thread driving several devices behind the seam): the destination device adds a peer's
// fp64 master straight out of the peer's memory -- P2P loads over NVLink inside the kernel, no staging copy, no
// packing -- and the peer's colour lanes likewise; the source zeroes its accumulators afterwards.
__global__ void __launch_bounds__(256) merge_peer_kernel(double4* master, const double4* peer_master, uint32_t pixels,
                                                         float* lanes, const float* peer_lanes, uint64_t lane_floats) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += stride) {
    const double4 v = peer_master[i];
    if (v.x != 0.0 || v.y != 0.0 || v.z != 0.0 || v.w != 0.0) {
      double4 m = master[i];
      m.x += v.x;
      m.y += v.y;
      m.z += v.z;
      m.w += v.w;
      master[i] = m;
    }
  }
  for (uint64_t i = blockIdx.x * blockDim.x + threadIdx.x; i < lane_floats; i += stride) {
    const float v = peer_lanes[i];
    if (v != 0.0f) lanes[i] += v;
  }
}

// Self-test of the unchecked exact division / square root (dvd_nr, sqrt_nr, hb_device.cuh) against the IEEE
// intrinsics over counter-based random operands. mode 0: a / b with |a|, |b| in [2^-40, 2^40], either sign of a,
// b > 0, one numerator in 64 an exact +-0 (bitwise comparison, so the signed zeros count); mode 1: the same with
// either sign of b and non-zero a; mode 2: sqrt(x), x in [2^-60, 2^60]. bad[0] counts mismatches, bad[1..3] keep
// the operand bits and the unchecked result of the last one.
__global__ void __launch_bounds__(256) selftest_arith_kernel(uint64_t n, uint32_t seed, uint32_t mode, unsigned long long* bad) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t k = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += stride) {
    const uint32_t lo = static_cast<uint32_t>(k), hi = static_cast<uint32_t>(k >> 32);
    const uint32_t s0 = seed_with_high(seed, hi);
    const uint32_t h0 = pcg_hash(s0 ^ pcg_hash(lo * 3u)), h1 = pcg_hash(s0 ^ pcg_hash(lo * 3u + 1u)),
                   h2 = pcg_hash(s0 ^ pcg_hash(lo * 3u + 2u));
    // mantissa from one hash, exponent (and signs) from another
    const uint32_t ea = 127u - 40u + (h2 & 0xFFFFu) % 81u, eb = 127u - 40u + (h2 >> 16) % 81u;
    uint32_t abits = (h0 & 0x807FFFFFu) | (ea << 23);
    uint32_t bbits = (h1 & 0x007FFFFFu) | (eb << 23);
    bool ok;
    uint32_t got;
    if (mode == 2u) {
      const uint32_t ex = 127u - 60u + (h2 & 0xFFFFu) % 121u;
      const float x = __uint_as_float((h0 & 0x007FFFFFu) | (ex << 23));
      abits = __float_as_uint(x);
      got = __float_as_uint(sqrt_nr(x));
      ok = got == __float_as_uint(__fsqrt_rn(x));
    } else {
      if (mode == 0u && (h1 >> 26) == 0u) abits &= 0x80000000u;  // +-0 numerator
      if (mode == 1u) bbits |= h1 & 0x80000000u;
      const float a = __uint_as_float(abits), b = __uint_as_float(bbits);
      got = __float_as_uint(dvd_nr(a, b));
      ok = got == __float_as_uint(__fdiv_rn(a, b));
    }
    if (!ok) {
      atomicAdd(bad, 1ull);
      bad[1] = abits;
      bad[2] = bbits;
      bad[3] = got;
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


#endif  // HB_TU

}  // namespace hb

#endif  // HB_KERNELS_CUH_
