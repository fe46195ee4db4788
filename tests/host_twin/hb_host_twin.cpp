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
