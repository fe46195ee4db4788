// hb_tables.h — per-shape tables derived from an HbCrystalTables on the host (upload_layer) and on the device
// (resample_shapes_kernel) by the SAME code: the paired-axis table of the slab scan, the meta word and the entry
// face groups. Split out of hb_kernels.cuh so that the CPU checks of the device arithmetic (tests/host_twin, which
// compiles hb_device.cuh and this file with g++) build their axis tables with the code the engine runs.
#ifndef HB_TABLES_H_
#define HB_TABLES_H_

#ifdef HB_HOST_TWIN
#include "hb_host_twin_shim.h"
#else
#include <cuda_runtime.h>
#endif
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "halotrace_b200.h"

namespace hb {

constexpr uint32_t kMetaP4 = 1u << 24;  // shape meta word: the shape is a full hexagonal prism (four paired axes)

// Entry sampling by face groups. The reference draws the entry triangle from a categorical over ALL fan
// triangles with weights max(-d.n_t, 0) * area_t (InitRay_p_fid, simulator.cpp:133-192). The triangles of one
// face share their normal, so the same distribution is drawn in two levels: a categorical over the faces
// (weight max(-d.n_f, 0) * A_f, A_f = fan area) and, inside the chosen face, the triangle whose cumulative
// area fraction brackets the residual of the same uniform. A prism needs 8 dot products instead of 20.
// Host-built per shape (build_entry_faces, hb_engine.cu); a group is a run of consecutive triangles with the
// same face id and normal.
struct EntryFaces {
  float4 na[HB_MAX_FACES];     // group normal, group area
  float4 pick[HB_MAX_FACES];   // cum[first], cum[first + 1], cum[first + 2] (+inf beyond the group), bits: first | cnt << 8
  float cum[HB_MAX_SUBTRIS];   // per triangle: cumulative area fraction inside its group
  uint8_t first[HB_MAX_FACES];
  uint8_t cnt[HB_MAX_FACES];
  uint32_t group_cnt;          // 0: more than HB_MAX_FACES groups -> triangle-level sampler (sample_entry)
  uint32_t pad_[1];
};
static_assert(sizeof(EntryFaces) % 16 == 0, "EntryFaces is copied as uint4 words");

// Face groups of one shape's entry fan table: runs of consecutive triangles that share the face id and the
// normal; per triangle the cumulative area fraction inside its run. Host (upload_layer) and device
// (derive_shapes_kernel) run this same code.
__host__ __device__ inline float bits_to_float_hd(int v) {
  float f;
  memcpy(&f, &v, 4);
  return f;
}
__host__ __device__ inline void build_entry_faces(const HbCrystalTables& t, EntryFaces* out) {
  EntryFaces& ef = *out;
  memset(&ef, 0, sizeof(ef));
  uint32_t g = 0;
  for (uint32_t i = 0; i < t.subtri_cnt;) {
    uint32_t j = i + 1;
    while (j < t.subtri_cnt && t.tri_face[j] == t.tri_face[i] && fabsf(t.tri_n[j][0] - t.tri_n[i][0]) <= 1e-4f &&
           fabsf(t.tri_n[j][1] - t.tri_n[i][1]) <= 1e-4f && fabsf(t.tri_n[j][2] - t.tri_n[i][2]) <= 1e-4f)
      j++;
    if (g == HB_MAX_FACES) {  // cannot happen for the reference's crystals; keep the triangle-level sampler
      ef.group_cnt = 0;
      return;
    }
    float area = 0.0f;
    for (uint32_t k = i; k < j; k++) area += t.tri_area[k];
    float run = 0.0f;
    for (uint32_t k = i; k < j; k++) {
      run += t.tri_area[k];
      ef.cum[k] = area > 0.0f ? run / area : 1.0f;
    }
    ef.na[g] = make_float4(t.tri_n[i][0], t.tri_n[i][1], t.tri_n[i][2], area);
    ef.first[g] = static_cast<uint8_t>(i);
    ef.cnt[g] = static_cast<uint8_t>(j - i);
    {
      // the same thresholds once more, packed for one 16-byte load (groups of up to 4 triangles: every face of a
      // hexagonal prism or pyramid); r >= +inf never holds, so missing thresholds count nothing
      const uint32_t pbits = i | ((j - i) << 8);
      float pb;
      memcpy(&pb, &pbits, 4);
      const float inf = bits_to_float_hd(0x7f800000);
      ef.pick[g] = make_float4(j - i > 1u ? ef.cum[i] : inf, j - i > 2u ? ef.cum[i + 1u] : inf, j - i > 3u ? ef.cum[i + 2u] : inf, pb);
    }
    g++;
    i = j;
  }
  ef.group_cnt = g;
}

// Everything the trace kernels read of one shape, derived from its HbCrystalTables: plane table, face numbers,
// the paired-axis table of the slab scan (faces whose unit normals are exact negatives share one entry, see
// slab_exit), meta word (face_cnt | population << 8 | axis_cnt << 16 | kMetaP4) and the entry face groups.
// Returns true when the shape qualifies for the P4 kernels (exactly four axes, all paired).
__host__ __device__ inline bool derive_shape_tables(const HbCrystalTables& t, uint32_t pop, float4* planes, uint8_t* fn,
                                                    float4* axes, uint32_t* meta, EntryFaces* ef) {
  for (uint32_t f = 0; f < HB_MAX_FACES; f++) {
    planes[f] = make_float4(t.plane[f][0], t.plane[f][1], t.plane[f][2], t.plane[f][3]);
    fn[f] = t.face_fn[f];
    axes[2u * f] = make_float4(0.f, 0.f, 0.f, 0.f);
    axes[2u * f + 1u] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  uint32_t axis_cnt = 0;
  bool used[HB_MAX_FACES];
  for (uint32_t f = 0; f < HB_MAX_FACES; f++) used[f] = false;
  bool all_paired = true;
  for (uint32_t f = 0; f < t.face_cnt && f < HB_MAX_FACES; f++) {
    if (used[f]) continue;
    used[f] = true;
    uint32_t partner = 63u;  // kFaceInvalid
    for (uint32_t g = f + 1; g < t.face_cnt && g < HB_MAX_FACES; g++) {
      if (!used[g] && t.plane[g][0] == -t.plane[f][0] && t.plane[g][1] == -t.plane[f][1] && t.plane[g][2] == -t.plane[f][2]) {
        partner = g;
        used[g] = true;
        break;
      }
    }
    const uint32_t fbits = f | (partner << 8);
    float fb;
    memcpy(&fb, &fbits, 4);
    axes[2u * axis_cnt] = make_float4(t.plane[f][0], t.plane[f][1], t.plane[f][2], t.plane[f][3]);
    axes[2u * axis_cnt + 1u] = make_float4(partner == 63u ? 0.0f : t.plane[partner][3], fb, 0.f, 0.f);
    axis_cnt++;
    if (partner == 63u) all_paired = false;
  }
  // P4 = full hexagonal prism in MakeCrystal's canonical frame: four paired axes, axis 0 = (0, 0, 1), axis 1 = (1, 0, 0),
  // axes 2 and 3 in the xy-plane -- exactly (dot_axis_p4 relies on the zeros and ones being exact).
  const bool canonical = axis_cnt == 4u && axes[0].x == 0.0f && axes[0].y == 0.0f && axes[0].z == 1.0f &&
                         axes[2].x == 1.0f && axes[2].y == 0.0f && axes[2].z == 0.0f && axes[4].z == 0.0f && axes[6].z == 0.0f;
  const bool p4 = all_paired && axis_cnt == 4u && canonical;
  *meta = t.face_cnt | (pop << 8) | (axis_cnt << 16) | (p4 ? kMetaP4 : 0u);
  build_entry_faces(t, ef);
  return p4;
}

}  // namespace hb

#endif  // HB_TABLES_H_
