// hb_geometry.h — crystal geometry shared by the host table builders (hb_host.cpp) and the device-side
// stochastic geometry pool (resample_shapes_kernel, hb_kernels.cuh). One source, two compilers: all arithmetic
// is IEEE double / float add, mul, div and sqrt without contraction (the library is built with -fmad=false and
// the host pass has no FMA target), so the device builds bit-identical tables to the host for the same
// shape scalars (tests/test_gpu_parity.py::test_device_geometry_pool...).
//
// What is built (reference: MakeCrystal simulator.cpp:405-450, ComputeClosedFormPrism/Pyramid
// geo3d_closedform.hpp:195,290, PopulateFromCfGeom crystal.cpp:304-347, BuildEntrySubTris simulator.cpp:61-129):
// the compact present-face plane table, face numbers and the entry fan table of a convex crystal given by up
// to 20 half-spaces.
#ifndef HB_GEOMETRY_H_
#define HB_GEOMETRY_H_

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "halotrace_b200.h"

#if defined(__CUDACC__)
#define HBG_HD __host__ __device__ inline
#else
#define HBG_HD inline
#endif

namespace hbg {

constexpr float kFloatEps = 1e-5f;               // math::kFloatEps
constexpr float kSqrt3F = 1.73205080757f;        // math::kSqrt3
constexpr float kPiF = 3.14159265359f;           // reference math::kPi (src/core/math.hpp:20)
constexpr float kDeg2RadF = kPiF / 180.0f;       // math::kDegreeToRad
constexpr int kMaxPoly = 32;                     // clip polygon capacity (4 + one corner per cutting plane)

// Hexagon face-normal directions (theta_i = i*60 deg) and corner directions (i*60 - 30 deg),
// reference: geo3d_closedform.hpp:12-19.
HBG_HD double FaceCos(int i) {
  const double v[6] = { 1.0, 0.5, -0.5, -1.0, -0.5, 0.5 };
  return v[i];
}
HBG_HD double FaceSin(int i) {
  const double h = 0.86602540378443864676;
  const double v[6] = { 0.0, h, h, 0.0, -h, -h };
  return v[i];
}
HBG_HD double VtxCos(int i) {
  const double h = 0.86602540378443864676;
  const double v[6] = { h, h, 0.0, -h, -h, 0.0 };
  return v[i];
}
HBG_HD double VtxSin(int i) {
  const double v[6] = { -0.5, 0.5, 1.0, 0.5, -0.5, -1.0 };
  return v[i];
}

struct V3 {
  double x, y, z;
};
HBG_HD V3 Add(V3 a, V3 b) { return V3{ a.x + b.x, a.y + b.y, a.z + b.z }; }
HBG_HD V3 Sub(V3 a, V3 b) { return V3{ a.x - b.x, a.y - b.y, a.z - b.z }; }
HBG_HD V3 Mul(V3 a, double s) { return V3{ a.x * s, a.y * s, a.z * s }; }
HBG_HD double Dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
HBG_HD V3 Cross(V3 a, V3 b) { return V3{ a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }

// A convex crystal given as up to 20 half-spaces coef.(x,y,z,1) <= 0 in the reference's slot order
// (0,1 basal; 2-7 prism; 8-13 upper pyramid; 14-19 lower pyramid; crystal.cpp:326-346).
struct PlaneSet {
  int slot_cnt;
  float coef[HB_MAX_FACES][4];
  float unit_n[HB_MAX_FACES][3];
  int face_number[HB_MAX_FACES];
  bool candidate[HB_MAX_FACES];
};
HBG_HD void ClearPlaneSet(PlaneSet* ps) { memset(ps, 0, sizeof(*ps)); }

struct Poly {
  V3 p[kMaxPoly];
  int n;
};

// Clip a convex polygon (3-D points on a plane) by the half-space c.(p,1) <= 0 (Sutherland-Hodgman).
// Returns false when the capacity is exceeded.
HBG_HD bool ClipPolygon(Poly* poly, const double c[4], double tol) {
  if (poly->n == 0) return true;
  Poly out;
  out.n = 0;
  const int n = poly->n;
  for (int i = 0; i < n; i++) {
    const V3 a = poly->p[i], b = poly->p[(i + 1) % n];
    const double fa = c[0] * a.x + c[1] * a.y + c[2] * a.z + c[3];
    const double fb = c[0] * b.x + c[1] * b.y + c[2] * b.z + c[3];
    const bool ina = fa <= tol, inb = fb <= tol;
    if (ina) {
      if (out.n == kMaxPoly) return false;
      out.p[out.n++] = a;
    }
    if (ina != inb) {
      if (out.n == kMaxPoly) return false;
      const double t = fa / (fa - fb);
      out.p[out.n++] = Add(a, Mul(Sub(b, a), t));
    }
  }
  *poly = out;
  return true;
}

HBG_HD void DedupRing(Poly* poly, double tol) {
  Poly out;
  out.n = 0;
  for (int i = 0; i < poly->n; i++) {
    const V3 p = poly->p[i];
    if (out.n > 0) {
      const V3 d = Sub(p, out.p[out.n - 1]);
      if (sqrt(Dot(d, d)) < tol) continue;
    }
    out.p[out.n++] = p;
  }
  while (out.n > 1) {
    const V3 d = Sub(out.p[0], out.p[out.n - 1]);
    if (sqrt(Dot(d, d)) < tol) out.n--; else break;
  }
  *poly = out;
}

// Face polygons of the intersection of the candidate half-spaces, each CCW seen from outside.
// Produces the compact present-face tables + the entry fan table (BuildEntrySubTris convention:
// fan (0,k,k+1), raw-winding normal, area = |cross|/2; simulator.cpp:90-129).
// Returns HB_OK or HB_ERR_CAPACITY (face with more than 12 corners / more than 64 fan triangles).
HBG_HD int BuildTablesFromPlanes(const PlaneSet& ps, HbCrystalTables* out) {
  memset(out, 0, sizeof(*out));
  uint32_t face = 0, tri = 0;
  for (int s = 0; s < ps.slot_cnt; s++) {
    if (!ps.candidate[s]) continue;
    const V3 n{ ps.coef[s][0], ps.coef[s][1], ps.coef[s][2] };
    const double mag = sqrt(Dot(n, n));
    if (!(mag > 0)) continue;
    const V3 nu = Mul(n, 1.0 / mag);
    const double d0 = ps.coef[s][3] / mag;
    const V3 origin = Mul(nu, -d0);
    const V3 helper = fabs(nu.z) < 0.9 ? V3{ 0, 0, 1 } : V3{ 1, 0, 0 };
    V3 t1 = Cross(helper, nu);
    t1 = Mul(t1, 1.0 / sqrt(Dot(t1, t1)));
    const V3 t2 = Cross(nu, t1);  // t1 x t2 = nu  => (t1,t2)-CCW is CCW seen from outside
    const double big = 1.0e3;
    Poly poly;
    poly.n = 4;
    poly.p[0] = Add(Add(origin, Mul(t1, -big)), Mul(t2, -big));
    poly.p[1] = Add(Add(origin, Mul(t1, big)), Mul(t2, -big));
    poly.p[2] = Add(Add(origin, Mul(t1, big)), Mul(t2, big));
    poly.p[3] = Add(Add(origin, Mul(t1, -big)), Mul(t2, big));
    for (int j = 0; j < ps.slot_cnt && poly.n != 0; j++) {
      if (j == s || !ps.candidate[j]) continue;
      double c[4] = { ps.coef[j][0], ps.coef[j][1], ps.coef[j][2], ps.coef[j][3] };
      const double cm = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
      if (!(cm > 0)) continue;
      for (int q = 0; q < 4; q++) c[q] /= cm;
      if (!ClipPolygon(&poly, c, 1e-12)) return HB_ERR_CAPACITY;
    }
    DedupRing(&poly, 1e-7);
    if (poly.n < 3) continue;
    double area2 = 0;
    for (int k = 1; k + 1 < poly.n; k++) {
      const V3 cr = Cross(Sub(poly.p[k], poly.p[0]), Sub(poly.p[k + 1], poly.p[0]));
      area2 += Dot(cr, nu);
    }
    if (!(area2 > 1e-10)) continue;
    if (poly.n > static_cast<int>(HB_MAX_FACE_VTX)) return HB_ERR_CAPACITY;
    // --- present face: plane entry (PopulateFromCfGeom, crystal.cpp:304-347) ---
    out->plane[face][0] = ps.unit_n[s][0];
    out->plane[face][1] = ps.unit_n[s][1];
    out->plane[face][2] = ps.unit_n[s][2];
    const float* cf = ps.coef[s];
    const float nrm = sqrtf(cf[0] * cf[0] + cf[1] * cf[1] + cf[2] * cf[2]);
    out->plane[face][3] = nrm > kFloatEps ? cf[3] / nrm : 0.0f;
    out->face_fn[face] = static_cast<uint8_t>(ps.face_number[s]);
    // --- fan sub-triangles ---
    float v[HB_MAX_FACE_VTX * 3];
    for (int k = 0; k < poly.n; k++) {
      v[k * 3 + 0] = static_cast<float>(poly.p[k].x);
      v[k * 3 + 1] = static_cast<float>(poly.p[k].y);
      v[k * 3 + 2] = static_cast<float>(poly.p[k].z);
    }
    for (int k = 1; k + 1 < poly.n; k++) {
      if (tri >= HB_MAX_SUBTRIS) return HB_ERR_CAPACITY;
      float* tv = out->tri_v[tri];
      for (int q = 0; q < 3; q++) {
        tv[q] = v[q];
        tv[3 + q] = v[k * 3 + q];
        tv[6 + q] = v[(k + 1) * 3 + q];
      }
      const float e1[3] = { tv[3] - tv[0], tv[4] - tv[1], tv[5] - tv[2] };
      const float e2[3] = { tv[6] - tv[0], tv[7] - tv[1], tv[8] - tv[2] };
      const float cr[3] = { -e2[1] * e1[2] + e1[1] * e2[2], e2[0] * e1[2] - e1[0] * e2[2], -e2[0] * e1[1] + e1[0] * e2[1] };
      const float len = sqrtf(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
      out->tri_area[tri] = len / 2.0f;
      for (int q = 0; q < 3; q++) out->tri_n[tri][q] = len > 0.0f ? cr[q] / len : 0.0f;
      out->tri_face[tri] = static_cast<uint8_t>(face);
      tri++;
    }
    face++;
  }
  out->face_cnt = face;
  out->subtri_cnt = tri;
  if (face < 4) {  // not a solid: the reference returns an empty Crystal (crystal.cpp:80-100)
    memset(out, 0, sizeof(*out));
  }
  return HB_OK;
}

// Largest inset m (in face-distance units) for which the hexagonal cross-section
// { n_i . x <= (sqrt3/4)(dist_i - m) } is non-empty: a 3-variable LP solved by vertex enumeration.
// This is the apex height parameter of a pyramidal segment (geo3d_closedform.cpp MaxFeasibleInsetLP).
HBG_HD double ApexInset(const float dist[6]) {
  const double k = 0.25 * 1.7320508075688772935;
  double best = 0.0;
  bool found = false;
  for (int a = 0; a < 6; a++)
    for (int b = a + 1; b < 6; b++)
      for (int c = b + 1; c < 6; c++) {
        const int id[3] = { a, b, c };
        double M[3][4];
        for (int r = 0; r < 3; r++) {
          M[r][0] = FaceCos(id[r]);
          M[r][1] = FaceSin(id[r]);
          M[r][2] = k;
          M[r][3] = k * dist[id[r]];
        }
        const double det = M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
                           M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
        if (fabs(det) < 1e-12) continue;
        double sol[3];
        for (int col = 0; col < 3; col++) {
          double T[3][3];
          for (int r = 0; r < 3; r++)
            for (int q = 0; q < 3; q++) T[r][q] = q == col ? M[r][3] : M[r][q];
          sol[col] = (T[0][0] * (T[1][1] * T[2][2] - T[1][2] * T[2][1]) - T[0][1] * (T[1][0] * T[2][2] - T[1][2] * T[2][0]) +
                      T[0][2] * (T[1][0] * T[2][1] - T[1][1] * T[2][0])) / det;
        }
        const double u = sol[0], v = sol[1], m = sol[2];
        bool ok = true;
        for (int i = 0; i < 6 && ok; i++) ok = FaceCos(i) * u + FaceSin(i) * v + k * m <= k * dist[i] + 1e-9;
        if (ok && (!found || m > best)) {
          best = m;
          found = true;
        }
      }
  return found ? (best > 0.0 ? best : 0.0) : 0.0;
}

// geo3d_closedform.cpp ComputeClosedFormPrism: basal (0,0,+-1,-h/2); side i 0.5(cos,sin,0), d = -dist*sqrt3/8.
HBG_HD void PrismPlanes(float h, const float dist[6], PlaneSet* ps) {
  ClearPlaneSet(ps);
  ps->slot_cnt = 8;
  const float hh = 0.5f * h;
  const float basal[2][4] = { { 0, 0, 1, -hh }, { 0, 0, -1, -hh } };
  for (int s = 0; s < 2; s++) {
    for (int q = 0; q < 4; q++) ps->coef[s][q] = basal[s][q];
    ps->unit_n[s][0] = 0;
    ps->unit_n[s][1] = 0;
    ps->unit_n[s][2] = basal[s][2];
    ps->face_number[s] = s + 1;
    ps->candidate[s] = true;
  }
  const double kd = static_cast<double>(kSqrt3F) / 8.0;
  for (int i = 0; i < 6; i++) {
    const int s = 2 + i;
    ps->coef[s][0] = 0.5f * static_cast<float>(FaceCos(i));
    ps->coef[s][1] = 0.5f * static_cast<float>(FaceSin(i));
    ps->coef[s][2] = 0.0f;
    ps->coef[s][3] = -static_cast<float>(kd * static_cast<double>(dist[i]));
    ps->unit_n[s][0] = static_cast<float>(FaceCos(i));
    ps->unit_n[s][1] = static_cast<float>(FaceSin(i));
    ps->unit_n[s][2] = 0.0f;
    ps->face_number[s] = 3 + i;
    ps->candidate[s] = true;
  }
}

// Crystal::CreatePrism (crystal.cpp:349-356).
HBG_HD int MakePrism(float h, const float dist6[6], HbCrystalTables* out) {
  memset(out, 0, sizeof(*out));
  if (!(h > kFloatEps)) return HB_OK;  // zero-volume: empty crystal (crystal.cpp:80-82)
  PlaneSet ps;
  PrismPlanes(h, dist6, &ps);
  return BuildTablesFromPlanes(ps, out);
}

// Crystal::CreatePyramid, wedge-angle form (crystal.cpp:380-384; geo3d_closedform.cpp ComputeClosedFormPyramid +
// ComputeClosedFormPyramidInner). a1 / a2 = (sqrt3/4) / tan(alpha) of the upper / lower segment when it exists
// (h > eps and alpha in [0.1, 89.9] deg), else any value <= 0: the tangent is taken by the caller (host libm),
// because the wedge angles are fixed per population and tan() is not bit-identical across host and device.
HBG_HD int MakePyramidFromSlopes(double a1, double a2, float h1, float h2, float h3, const float dist[6], HbCrystalTables* out) {
  memset(out, 0, sizeof(*out));
  const bool has_upper = a1 > 0 && h1 > kFloatEps, has_lower = a2 > 0 && h3 > kFloatEps;
  const double h2_2 = 0.5 * static_cast<double>(h2);
  if (!has_upper && !has_lower && h2 < kFloatEps) return HB_OK;

  PlaneSet ps;
  ClearPlaneSet(&ps);
  ps.slot_cnt = 20;
  ps.face_number[0] = 1;
  ps.face_number[1] = 2;
  for (int i = 0; i < 6; i++) {
    ps.face_number[2 + i] = 3 + i;
    ps.face_number[8 + i] = 13 + i;
    ps.face_number[14 + i] = 23 + i;
    const int i2 = (i + 1) % 6;
    const double x1 = 0.5 * VtxCos(i), x2 = 0.5 * VtxCos(i2), y1 = 0.5 * VtxSin(i), y2 = 0.5 * VtxSin(i2);
    const double det = x1 * y2 - x2 * y1;
    float* c = ps.coef[2 + i];
    c[0] = static_cast<float>(y2 - y1);
    c[1] = static_cast<float>(x1 - x2);
    c[2] = 0;
    c[3] = static_cast<float>(-static_cast<double>(dist[i]) * det);
    ps.candidate[2 + i] = true;
    if (has_upper) {
      float* u = ps.coef[8 + i];
      u[0] = static_cast<float>(a1 * (y2 - y1));
      u[1] = static_cast<float>(a1 * (x1 - x2));
      u[2] = static_cast<float>(det);
      u[3] = static_cast<float>(-(h2_2 + a1 * static_cast<double>(dist[i])) * det);
      ps.candidate[8 + i] = true;
    }
    if (has_lower) {
      float* l = ps.coef[14 + i];
      l[0] = static_cast<float>(a2 * (y2 - y1));
      l[1] = static_cast<float>(a2 * (x1 - x2));
      l[2] = static_cast<float>(-det);
      l[3] = static_cast<float>(-(h2_2 + a2 * static_cast<double>(dist[i])) * det);
      ps.candidate[14 + i] = true;
    }
  }
  const double m_apex = (has_upper || has_lower) ? ApexInset(dist) : 0.0;
  const double mt = static_cast<double>(h1) * m_apex, mb = static_cast<double>(h3) * m_apex;
  const double m_top = has_upper ? (mt < m_apex ? mt : m_apex) : 0.0;
  const double m_bot = has_lower ? (mb < m_apex ? mb : m_apex) : 0.0;
  const double z_top = has_upper ? h2_2 + a1 * m_top : h2_2;
  const double z_bot = has_lower ? -h2_2 - a2 * m_bot : -h2_2;
  const float basal[2][4] = { { 0, 0, 1, static_cast<float>(-z_top) }, { 0, 0, -1, static_cast<float>(z_bot) } };
  for (int s = 0; s < 2; s++) {
    for (int q = 0; q < 4; q++) ps.coef[s][q] = basal[s][q];
    ps.candidate[s] = true;
  }
  ps.unit_n[0][2] = 1.0f;
  ps.unit_n[1][2] = -1.0f;
  for (int s = 2; s < 20; s++) {
    const double nx = ps.coef[s][0], ny = ps.coef[s][1], nz = ps.coef[s][2];
    const double mag = sqrt(nx * nx + ny * ny + nz * nz);
    if (mag > 0) {
      ps.unit_n[s][0] = static_cast<float>(nx / mag);
      ps.unit_n[s][1] = static_cast<float>(ny / mag);
      ps.unit_n[s][2] = static_cast<float>(nz / mag);
    }
  }
  return BuildTablesFromPlanes(ps, out);
}

// Shape scalars of one crystal instance in MakeCrystal's draw order (simulator.cpp:405-450) with the reference's
// sync groups (SyncGroupSampler, simulator.cpp:341-393): a slot in group g != 0 reuses the RAW draw of the first
// slot of g that was reached (no RNG consumed), group 0 draws independently. `draw(dist)` is the caller's RNG:
// mt19937 on the host, the counter-based PCG stream on the device.
// Outputs: hgt[3] = {h} (prism) or {upper_h, prism_h, lower_h} (pyramid), folded with |.|; dist[6] signed.
template <typename DrawFn>
HBG_HD void SampleShapeScalars(const HbCrystalDesc& c, DrawFn draw, float hgt[3], float dist[6]) {
  int cached_group[10];
  float cached_value[10];
  int cached_cnt = 0;
  auto sync_draw = [&](int slot, const HbDist& d) -> float {
    const int group = c.sync_group[slot];
    if (group == 0) return draw(d);
    for (int i = 0; i < cached_cnt; i++)
      if (cached_group[i] == group) return cached_value[i];
    const float v = draw(d);
    cached_group[cached_cnt] = group;
    cached_value[cached_cnt] = v;
    cached_cnt++;
    return v;
  };
  hgt[0] = hgt[1] = hgt[2] = 0.0f;
  if (c.kind == 0u) {
    hgt[0] = fabsf(sync_draw(0, c.height[0]));
  } else {
    hgt[0] = fabsf(sync_draw(1, c.height[0]));
    hgt[1] = fabsf(sync_draw(2, c.height[1]));
    hgt[2] = fabsf(sync_draw(3, c.height[2]));
  }
  for (int i = 0; i < 6; i++) dist[i] = sync_draw(4 + i, c.face_dist[i]);
}

}  // namespace hbg

#endif  // HB_GEOMETRY_H_
