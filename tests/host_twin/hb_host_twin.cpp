// TEST INFRASTRUCTURE: the product's device arithmetic (csrc/hb_device.cuh) and table derivation (csrc/hb_tables.h)
// compiled for the CPU, behind a small C interface for tests/test_host_twin.py. Nothing here is shipped.
#define HB_HOST_TWIN 1
#include "hb_device.cuh"
#include "hb_tables.h"

using namespace hb;

namespace {
struct HostAxisRow {  // the AxisRow interface of the kernels over a plain array: axis i = two float4
  const float4* p;
  void load(uint32_t i, float4& a, float4& b) const {
    a = p[2u * i];
    b = p[2u * i + 1u];
  }
};
uint32_t face_in(uint16_t f) { return f == HB_INVALID_FACE ? kFaceInvalid : f; }
uint16_t face_out(uint32_t f) { return f == kFaceInvalid ? static_cast<uint16_t>(HB_INVALID_FACE) : static_cast<uint16_t>(f); }
}  // namespace

extern "C" {

// derive_shape_tables as the engine runs it; returns 1 when the shape qualifies for the P4 (hexagonal prism) forms
int twin_derive(const HbCrystalTables* t, float* planes, uint8_t* fn, float* axes, uint32_t* meta) {
  EntryFaces ef;
  return derive_shape_tables(*t, 0u, reinterpret_cast<float4*>(planes), fn, reinterpret_cast<float4*>(axes), meta, &ef) ? 1 : 0;
}

// hit_surface; d_out6 = reflected xyz, refracted xyz; w_out2 = reflected, refracted (-1: total internal reflection)
void twin_hit_surface(const float* planes, float n_idx, float inv_n, uint64_t n, const float* d3, const float* w,
                      const uint16_t* face, float* d_out6, float* w_out2) {
  const float4* pl = reinterpret_cast<const float4*>(planes);
  for (uint64_t i = 0; i < n; i++) {
    if (face[i] == HB_INVALID_FACE) continue;
    const Split s = hit_surface(pl[face[i]], n_idx, inv_n, d3[3 * i], d3[3 * i + 1], d3[3 * i + 2], w[i]);
    float* o = d_out6 + 6 * i;
    o[0] = s.rx, o[1] = s.ry, o[2] = s.rz, o[3] = s.tx, o[4] = s.ty, o[5] = s.tz;
    w_out2[2 * i] = s.rw;
    w_out2[2 * i + 1] = s.tw;
  }
}

// Slab search of one child. mode 0: slab_exit<false>, 1: slab_exit<true>, 2: slab_exit_p4<false>, 3: slab_exit_p4<true>,
// 4: the near child of bounce_axes (generic one-pass form), 5: the near child of bounce_axes_p4. Modes 4 / 5 need a
// source face (its plane feeds the far child's test) and also return that test's verdict for the SAME direction.
void twin_propagate(const float* planes, const float* axes, uint32_t meta, uint32_t mode, uint64_t n, const float* d3,
                    const float* p3, const uint16_t* from_face, float* p_out3, uint16_t* to_face, uint8_t* far_exits) {
  const HostAxisRow ax{ reinterpret_cast<const float4*>(axes) };
  const float4* pl = reinterpret_cast<const float4*>(planes);
  const uint32_t axis_cnt = (meta >> 16) & 255u;
  for (uint64_t i = 0; i < n; i++) {
    const float px = p3[3 * i], py = p3[3 * i + 1], pz = p3[3 * i + 2];
    const float dx = d3[3 * i], dy = d3[3 * i + 1], dz = d3[3 * i + 2];
    const uint32_t src = face_in(from_face[i]);
    float ox = px, oy = py, oz = pz;
    uint32_t f = kFaceInvalid;
    bool fe = false;
    if (mode == 0u) f = slab_exit<false>(ax, axis_cnt, src, px, py, pz, dx, dy, dz, ox, oy, oz);
    else if (mode == 1u) f = slab_exit<true>(ax, axis_cnt, src, px, py, pz, dx, dy, dz, ox, oy, oz);
    else if (mode == 2u) f = slab_exit_p4<false>(ax, src, px, py, pz, dx, dy, dz, ox, oy, oz);
    else if (mode == 3u) f = slab_exit_p4<true>(ax, src, px, py, pz, dx, dy, dz, ox, oy, oz);
    else if (src != kFaceInvalid) {
      if (mode == 4u) f = bounce_axes(ax, axis_cnt, src, pl[src], px, py, pz, dx, dy, dz, dx, dy, dz, fe, ox, oy, oz);
      else f = bounce_axes_p4(ax, src, pl[src], px, py, pz, dx, dy, dz, dx, dy, dz, fe, ox, oy, oz);
    }
    p_out3[3 * i] = ox, p_out3[3 * i + 1] = oy, p_out3[3 * i + 2] = oz;
    to_face[i] = face_out(f);
    if (far_exits != nullptr) far_exits[i] = fe ? 1 : 0;
  }
}

// The quick classifications on their own: out[0] far_child_surely_exits (generic), out[1] far_child_surely_exits_p4,
// out[2] near_child_surely_hits, and from last_axes_p4: out[3] far_exits, out[4] near_hits (p4 shapes only).
void twin_quick_tests(const float* planes, const float* axes, uint32_t meta, uint64_t n, const float* d3, const float* p3,
                      const uint16_t* from_face, uint8_t* out5) {
  const HostAxisRow ax{ reinterpret_cast<const float4*>(axes) };
  const float4* pl = reinterpret_cast<const float4*>(planes);
  const uint32_t axis_cnt = (meta >> 16) & 255u;
  const bool p4 = (meta & kMetaP4) != 0u;
  for (uint64_t i = 0; i < n; i++) {
    uint8_t* o = out5 + 5 * i;
    o[0] = o[1] = o[2] = o[3] = o[4] = 0;
    const uint32_t src = face_in(from_face[i]);
    if (src == kFaceInvalid) continue;
    const float px = p3[3 * i], py = p3[3 * i + 1], pz = p3[3 * i + 2];
    const float dx = d3[3 * i], dy = d3[3 * i + 1], dz = d3[3 * i + 2];
    o[0] = far_child_surely_exits(ax, axis_cnt, src, pl[src], px, py, pz, dx, dy, dz) ? 1 : 0;
    o[2] = near_child_surely_hits(ax, axis_cnt, px, py, pz, dx, dy, dz) ? 1 : 0;
    if (p4) {
      o[1] = far_child_surely_exits_p4(ax, pl[src], px, py, pz, dx, dy, dz) ? 1 : 0;
      bool fe, nh;
      last_axes_p4(ax, pl[src], px, py, pz, dx, dy, dz, dx, dy, dz, fe, nh);
      o[3] = fe ? 1 : 0;
      o[4] = nh ? 1 : 0;
    }
  }
}

// project_exit (all 11 lens types): up to two pixel hits per direction, -1 where there is none
void twin_project(const HbProjParams* p, uint64_t n, const float* dir3, int32_t* px2, int32_t* py2, int32_t* cnt,
                  int32_t* bump2) {
  for (uint64_t i = 0; i < n; i++) {
    const PixelHits h = project_exit(*p, dir3[3 * i], dir3[3 * i + 1], dir3[3 * i + 2]);
    cnt[i] = h.count;
    for (int k = 0; k < 2; k++) {
      px2[2 * i + k] = k < h.count ? h.px[k] : -1;
      py2[2 * i + k] = k < h.count ? h.py[k] : -1;
      bump2[2 * i + k] = k < h.count ? (h.bump[k] ? 1 : 0) : 0;
    }
  }
}

// project_culls: the exact early "reaches no pixel of render 0" test of the emission's first stage
void twin_project_culls(const HbProjParams* p, uint64_t n, const float* dir3, uint8_t* culled) {
  for (uint64_t i = 0; i < n; i++) culled[i] = project_culls(*p, dir3[3 * i], dir3[3 * i + 1], dir3[3 * i + 2]) ? 1 : 0;
}

// filter_check (csrc/hb_filter.h, the matcher the emission runs), interface of the oracle's orc_filter_check
int twin_filter_check(const HbFilterDesc* f, const uint8_t* face_fn, uint32_t crystal_id, uint64_t n, const uint8_t* paths64,
                      const uint8_t* path_len, const float* dir3, uint8_t* pass) {
  for (uint64_t i = 0; i < n; i++) {
    uint8_t fn[64];
    for (uint32_t k = 0; k < path_len[i]; k++) fn[k] = face_fn[paths64[i * 64 + k]];
    pass[i] = filter_check(*f, fn, path_len[i], dir3 + 3 * i, crystal_id) ? 1 : 0;
  }
  return 0;
}
uint32_t twin_filter_max_len(const HbFilterDesc* f) { return filter_max_len(*f); }

// The generator's building blocks, same interface as the oracle's orc_* / the reference pin's ref_* wrappers
// (tests/harness.py: sampler_vectors), running the device functions of hb_device.cuh.
uint32_t twin_pcg_hash(uint32_t x) { return pcg_hash(x); }
uint32_t twin_seed_with_high(uint32_t seed, uint32_t hi) { return seed_with_high(seed, hi); }
uint32_t twin_feistel(uint32_t i, uint32_t n, uint32_t seed) { return feistel(i, n, seed); }
void twin_uniforms(uint32_t seed, uint32_t idx, uint32_t slot0, uint32_t n, float* out) {
  Stream s{ seed, idx, slot0 };
  for (uint32_t i = 0; i < n; i++) out[i] = s.next();
}
void twin_get_dist(uint32_t seed, uint32_t idx0, uint32_t n, uint32_t type, float mean, float stdv, float* out) {
  for (uint32_t i = 0; i < n; i++) {
    Stream s{ seed, idx0 + i, 0u };
    out[i] = get_dist(s, type, mean, stdv);
  }
}
void twin_lat_lon_roll(const HbAxisSampler* a, uint32_t seed, uint32_t idx0, uint32_t n, float* lon_lat_roll3,
                       uint32_t* slots_used) {
  // what upload_layer makes of the sampler: the scalar block + [theta | cdf | flip] (hb_engine.cu)
  const AxisParams ap{ a->lat_path, a->lat_mean, a->lat_std, a->az_type, a->az_mean, a->az_std,
                       a->roll_type, a->roll_mean, a->roll_std, a->lut_n };
  static float lut[3 * HB_LUT_NODES];
  memcpy(lut, a->lut_theta, sizeof(float) * HB_LUT_NODES);
  memcpy(lut + HB_LUT_NODES, a->lut_cdf, sizeof(float) * HB_LUT_NODES);
  memcpy(lut + 2 * HB_LUT_NODES, a->lut_flip, sizeof(float) * HB_LUT_NODES);
  for (uint32_t i = 0; i < n; i++) {
    Stream s{ seed, idx0 + i, 0u };
    sample_lon_lat_roll(s, ap, lut, lon_lat_roll3[3 * i], lon_lat_roll3[3 * i + 1], lon_lat_roll3[3 * i + 2]);
    if (slots_used != nullptr) slots_used[i] = s.slot;
  }
}
void twin_rotation9(uint64_t n, const float* lon_lat_roll3, float* rot9) {
  for (uint64_t i = 0; i < n; i++) {
    const Rot r = rot_from_quat(quat_from_angles(lon_lat_roll3[3 * i], lon_lat_roll3[3 * i + 1], lon_lat_roll3[3 * i + 2]));
    memcpy(rot9 + 9 * i, r.m, sizeof(r.m));
  }
}
void twin_sph_cap(uint32_t seed, uint32_t idx0, uint32_t slot0, uint32_t n, float lon, float lat, float half, float* d3) {
  for (uint32_t i = 0; i < n; i++) {
    Stream s{ seed, idx0 + i, slot0 };
    sample_sph_cap(s, lon, lat, half, d3[3 * i], d3[3 * i + 1], d3[3 * i + 2]);
  }
}
void twin_triangle(uint32_t seed, uint32_t idx0, uint32_t slot0, uint32_t n, const float* vtx9, float* p3) {
  // sample_entry's point-in-triangle step on a one-triangle fan table (weights irrelevant: a single candidate)
  static HbCrystalTables t;
  memset(&t, 0, sizeof(t));
  t.subtri_cnt = 1;
  memcpy(t.tri_v[0], vtx9, 9 * sizeof(float));
  t.tri_n[0][2] = 1.0f;
  t.tri_area[0] = 1.0f;
  for (uint32_t i = 0; i < n; i++) {
    Stream s{ seed, idx0 + i, slot0 - 1u };   // sample_entry draws the categorical uniform first
    uint32_t face;
    sample_entry(s, &t, 0.0f, 0.0f, -1.0f, p3[3 * i], p3[3 * i + 1], p3[3 * i + 2], face);
  }
}

// dvd_nr / sqrt_nr as compiled here (host reciprocal) against the host's IEEE division / square root
uint64_t twin_div_mismatches(uint64_t n, const float* a, const float* b) {
  uint64_t bad = 0;
  for (uint64_t i = 0; i < n; i++) {
    const float q = dvd_nr(a[i], b[i]), want = a[i] / b[i];
    bad += __float_as_uint(q) != __float_as_uint(want) ? 1u : 0u;
  }
  return bad;
}
uint64_t twin_sqrt_mismatches(uint64_t n, const float* x) {
  uint64_t bad = 0;
  for (uint64_t i = 0; i < n; i++) bad += __float_as_uint(sqrt_nr(x[i])) != __float_as_uint(sqrtf(x[i])) ? 1u : 0u;
  return bad;
}

}  // extern "C"
