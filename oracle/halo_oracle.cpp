// halo_oracle.cpp — CPU restatement of the reference trace path. TEST INFRASTRUCTURE ONLY.
// See halo_oracle.h for the pinning statement. Build: g++ -O2 -ffp-contract=off (no FMA contraction,
// IEEE single arithmetic) so +,-,*,/,sqrt agree bit-for-bit with the CUDA kernels compiled -fmad=false.
#include "halo_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

constexpr float kPi = 3.14159265358979323846f;   // LM_PI_F, lm_shims.h:74
constexpr float kPi2 = 1.5707963267948966f;      // LM_PI_2F
constexpr float kEps = 1e-5f;                    // math::kFloatEps == kSlabEps (traversal_shared.h:46)
constexpr uint32_t kNonceGen = 0x3C9A7F11u;      // stream nonces: pcg_shared.h:97-116 clearinghouse
constexpr uint32_t kNonceWl = 0x9E3779B9u;
constexpr uint32_t kNonceShape = 0x94D049BBu;
constexpr uint32_t kNonceGate = 0x5A5A5A5Au;
constexpr uint32_t kNonceTransit = 0xA5A5A5A5u;

// ---- counter-based RNG: pcg_shared.h:193-274 ---------------------------------------------------
uint32_t PcgHash(uint32_t x) {
  x = x * 747796405u + 2891336453u;
  x = ((x >> ((x >> 28u) + 4u)) ^ x) * 277803737u;
  return (x >> 22u) ^ x;
}
float U01(uint32_t h) { return static_cast<float>(h >> 8) * (1.0f / 16777216.0f); }
float Draw(uint32_t seed, uint32_t idx, uint32_t slot) { return U01(PcgHash(seed ^ PcgHash(idx * 1000003u + slot))); }
uint32_t SeedWithHigh(uint32_t seed, uint32_t hi) { return hi == 0u ? seed : seed ^ PcgHash(hi); }

struct Stream {
  uint32_t seed, idx, slot;
  float Next() { return Draw(seed, idx, slot++); }
};

// pcg_gaussian, pcg_shared.h:277-281
float Gaussian(Stream& s) {
  float u1 = std::fmax(s.Next(), 1e-7f);
  float u2 = s.Next();
  return std::sqrt(-2.0f * std::log(u1)) * std::cos(2.0f * kPi * u2);
}
// pcg_get_dist, pcg_shared.h:290-308
float GetDist(Stream& s, uint32_t type, float mean, float stdv) {
  if (type == HB_DIST_NO_RANDOM) return mean;
  if (type == HB_DIST_UNIFORM) return (s.Next() - 0.5f) * stdv + mean;
  if (type == HB_DIST_GAUSSIAN || type == HB_DIST_GAUSSIAN_LEGACY) return Gaussian(s) * stdv + mean;
  if (type == HB_DIST_ZIGZAG) return std::fabs(stdv * std::sin(s.Next() * 2.0f * kPi) + mean);
  float u = s.Next();
  float sgn = (u < 0.5f) ? -1.0f : 1.0f;
  float arg = std::fmax(1.0f - 2.0f * std::fabs(u - 0.5f), 1e-30f);
  return mean - stdv * sgn * std::log(arg);
}
// normalize_latitude, pcg_shared.h:311-322
void NormalizeLatitude(float phi, float& phi_out, bool& flip) {
  float theta = kPi2 - phi;
  theta = std::fmod(theta, 2.0f * kPi);
  if (theta < 0.0f) theta += 2.0f * kPi;
  flip = theta > kPi;
  if (flip) theta = 2.0f * kPi - theta;
  phi_out = kPi2 - theta;
}
// invert_lat_lut, pcg_shared.h:345-363
float InvertLut(float xi, const float* theta, const float* cdf, uint32_t n) {
  xi = std::fmin(std::fmax(xi, cdf[0]), cdf[n - 1]);
  uint32_t lo = 0, hi = n - 1;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (cdf[mid] <= xi) lo = mid; else hi = mid;
  }
  float c0 = cdf[lo], c1 = cdf[lo + 1];
  float denom = c1 - c0;
  float w = denom > 0.0f ? (xi - c0) / denom : 0.0f;
  return theta[lo] + w * (theta[lo + 1] - theta[lo]);
}
// lat_lut_bin, pcg_shared.h:370-378
uint32_t LutBin(float th, const float* theta, uint32_t n) {
  float span = theta[n - 1] - theta[0];
  float t = span > 0.0f ? (th - theta[0]) / span : 0.0f;
  int idx = static_cast<int>(t * static_cast<float>(n - 1));
  idx = idx < 0 ? 0 : idx;
  int last = static_cast<int>(n) - 2;
  return static_cast<uint32_t>(idx > last ? last : idx);
}
// sample_lat_lon_roll, pcg_shared.h:392-437
void SampleLonLatRoll(Stream& s, const HbAxisSampler& a, float& lon, float& lat, float& roll) {
  float phi = 0.0f;
  bool flip = false;
  lon = 0.0f;
  if (a.lat_path == HB_LAT_FULL_SPHERE) {
    float u = s.Next() * 2.0f - 1.0f;
    u = std::fmin(std::fmax(u, -1.0f), 1.0f);
    phi = std::asin(u);
    lon = s.Next() * 2.0f * kPi;
  } else if (a.lat_path == HB_LAT_NO_RANDOM) {
    phi = a.lat_mean;
  } else if (a.lat_path == HB_LAT_GAUSS_LEGACY) {
    float raw = GetDist(s, HB_DIST_GAUSSIAN_LEGACY, a.lat_mean, a.lat_std);
    NormalizeLatitude(raw, phi, flip);
  } else if (a.lat_path == HB_LAT_LUT) {
    float xi = s.Next();
    float colat = InvertLut(xi, a.lut_theta, a.lut_cdf, a.lut_n);
    phi = kPi2 - colat;
    uint32_t bin = LutBin(colat, a.lut_theta, a.lut_n);
    flip = s.Next() < a.lut_flip[bin];
  }
  if (a.lat_path != HB_LAT_FULL_SPHERE) lon = GetDist(s, a.az_type, a.az_mean, a.az_std);
  roll = GetDist(s, a.roll_type, a.roll_mean, a.roll_std);
  if (flip) {
    lon += kPi;
    roll += kPi;
  }
  lat = phi;
}

// Orientation R = Rz(lon - pi) * Ry(lat - pi/2) * Rz(roll) (BuildCrystalRotation, simulator.cpp:224-231),
// held by the engine as the unit quaternion of that product (DESIGN.md "orientation record"):
//   a = lon - pi, b = lat - pi/2, c = roll
//   q = (cos(b/2) cos((a+c)/2), sin(b/2) sin((c-a)/2), sin(b/2) cos((c-a)/2), cos(b/2) sin((a+c)/2))
void QuatFromAngles(float lon, float lat, float roll, float* q) {
  float a = lon - kPi, b = lat - kPi2, c = roll;
  float hb = 0.5f * b, hp = 0.5f * (a + c), hm = 0.5f * (c - a);
  float cb = std::cos(hb), sb = std::sin(hb);
  q[0] = cb * std::cos(hp);
  q[1] = sb * std::sin(hm);
  q[2] = sb * std::cos(hm);
  q[3] = cb * std::sin(hp);
}
void QuatToRot(const float* q, float* m) {
  float w = q[0], x = q[1], y = q[2], z = q[3];
  float xx = x * x, yy = y * y, zz = z * z;
  float xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
  m[0] = 1.0f - 2.0f * (yy + zz);
  m[1] = 2.0f * (xy - wz);
  m[2] = 2.0f * (xz + wy);
  m[3] = 2.0f * (xy + wz);
  m[4] = 1.0f - 2.0f * (xx + zz);
  m[5] = 2.0f * (yz - wx);
  m[6] = 2.0f * (xz - wy);
  m[7] = 2.0f * (yz + wx);
  m[8] = 1.0f - 2.0f * (xx + yy);
}
// Rotation::Apply (geo3d.cpp:69-77): world = M v, row dot products left to right.
void ApplyRot(const float* m, const float* v, float* o) {
  o[0] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  o[1] = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  o[2] = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
}
// apply_inverse_mat9, pcg_shared.h:487-491
void ApplyRotT(const float* m, const float* v, float* o) {
  o[0] = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
  o[1] = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
  o[2] = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
}
// sample_sph_cap, pcg_shared.h:514-529
void SampleSphCap(Stream& s, float lon, float lat, float half, float* d) {
  float c_cap = std::cos(half);
  float u = s.Next();
  float x = u + (1.0f - u) * c_cap;
  float r = std::sqrt(std::fmax(1.0f - x * x, 0.0f));
  float phi = s.Next() * 2.0f * kPi;
  float y = std::cos(phi) * r;
  float z = std::sin(phi) * r;
  float cl = std::cos(lon), sl = std::sin(lon), ca = std::cos(lat), sa = std::sin(lat);
  d[0] = cl * ca * x - sl * y - cl * sa * z;
  d[1] = sl * ca * x + cl * y - sl * sa * z;
  d[2] = sa * x + ca * z;
}
// Entry point: area x facing categorical pick over the fan table + uniform point in triangle
// (InitRay_p_fid simulator.cpp:133-192; device twins categorical_sample / sample_triangle
// pcg_shared.h:496-509,607-624).
// The engine draws the triangle in two levels (face group, then triangle inside the group by cumulative area
// fraction of the same uniform): the same distribution as the reference's categorical over all fan triangles,
// because the triangles of one face share their normal. Groups = runs of consecutive triangles with the same
// face id and normal; more than HB_MAX_FACES groups -> triangle-level categorical.
struct EntryGroups {
  float n[HB_MAX_FACES][3], area[HB_MAX_FACES];
  float cum[HB_MAX_SUBTRIS];
  uint32_t first[HB_MAX_FACES], cnt[HB_MAX_FACES], group_cnt;
};
EntryGroups BuildEntryGroups(const HbCrystalTables& t) {
  EntryGroups eg;
  std::memset(&eg, 0, sizeof(eg));
  uint32_t g = 0;
  for (uint32_t i = 0; i < t.subtri_cnt;) {
    uint32_t j = i + 1;
    while (j < t.subtri_cnt && t.tri_face[j] == t.tri_face[i] && std::fabs(t.tri_n[j][0] - t.tri_n[i][0]) <= 1e-4f &&
           std::fabs(t.tri_n[j][1] - t.tri_n[i][1]) <= 1e-4f && std::fabs(t.tri_n[j][2] - t.tri_n[i][2]) <= 1e-4f)
      j++;
    if (g == HB_MAX_FACES) {
      eg.group_cnt = 0;
      return eg;
    }
    float area = 0.0f;
    for (uint32_t k = i; k < j; k++) area += t.tri_area[k];
    float run = 0.0f;
    for (uint32_t k = i; k < j; k++) {
      run += t.tri_area[k];
      eg.cum[k] = area > 0.0f ? run / area : 1.0f;
    }
    for (int q = 0; q < 3; q++) eg.n[g][q] = t.tri_n[i][q];
    eg.area[g] = area;
    eg.first[g] = i;
    eg.cnt[g] = j - i;
    g++;
    i = j;
  }
  eg.group_cnt = g;
  return eg;
}

uint32_t PickEntryTriangle(Stream& s, const HbCrystalTables& t, const float* d, bool* grouped) {
  const EntryGroups eg = BuildEntryGroups(t);
  *grouped = eg.group_cnt != 0;
  if (eg.group_cnt == 0) {  // triangle-level categorical (InitRay_p_fid, simulator.cpp:133-192)
    float prob[HB_MAX_SUBTRIS];
    uint32_t n = t.subtri_cnt;
    float total = 0.0f;
    for (uint32_t i = 0; i < n; i++) {
      float dot = d[0] * t.tri_n[i][0] + d[1] * t.tri_n[i][1] + d[2] * t.tri_n[i][2];
      prob[i] = std::fmax(-dot * t.tri_area[i], 0.0f);
      total += prob[i];
    }
    float u_cat = s.Next();
    uint32_t tri = 0;
    if (total > 0.0f) {
      float target = u_cat * total;
      float cum = 0.0f;
      tri = n - 1;
      for (uint32_t i = 0; i < n; i++) {
        cum += prob[i];
        if (cum > target) {
          tri = i;
          break;
        }
      }
    }
    return tri;
  }
  const uint32_t ng = eg.group_cnt;
  float w[HB_MAX_FACES];
  float total = 0.0f;
  for (uint32_t g = 0; g < ng; g++) {
    float dot = d[0] * eg.n[g][0] + d[1] * eg.n[g][1] + d[2] * eg.n[g][2];
    w[g] = std::fmax(-dot * eg.area[g], 0.0f);
    total += w[g];
  }
  float u_cat = s.Next();
  if (!(total > 0.0f)) return 0;
  float target = u_cat * total;
  uint32_t sel = ng - 1;
  float resid = 0.0f, w_sel = 0.0f, cum = 0.0f;
  for (uint32_t g = 0; g < ng; g++) {
    float c1 = cum + w[g];
    if (c1 > target) {
      sel = g;
      resid = target - cum;
      w_sel = w[g];
      break;
    }
    cum = c1;
  }
  float r = w_sel > 0.0f ? resid / w_sel : 0.0f;
  uint32_t t0 = eg.first[sel], c = eg.cnt[sel], tri = t0;
  for (uint32_t j = 0; j + 1 < c; j++)
    if (r >= eg.cum[t0 + j]) tri = t0 + j + 1;
  return tri;
}

void SampleEntry(Stream& s, const HbCrystalTables& t, const float* d, float* p, uint16_t* face) {
  bool grouped = false;
  uint32_t tri = PickEntryTriangle(s, t, d, &grouped);
  float u = s.Next();
  float v = s.Next();
  if (u + v > 1.0f) {
    u = 1.0f - u;
    v = 1.0f - v;
  }
  const float* vt = t.tri_v[tri];
  for (int k = 0; k < 3; k++) {
    float a = vt[k], b = vt[3 + k], c = vt[6 + k];
    p[k] = u * (b - a) + v * (c - a) + a;
  }
  *face = t.tri_face[tri];
}

// ---- optics.cpp:18-53 HitSurface (one ray) ------------------------------------------------------
void HitSurface(const float* nrm, float n_idx, const float* d, float w, float* d_refl, float* w_refl, float* d_refr,
                float* w_refr) {
  float c = d[0] * nrm[0] + d[1] * nrm[1] + d[2] * nrm[2];
  float rr = c > 0 ? n_idx : 1.0f / n_idx;
  float delta = (1.0f - rr * rr) / (c * c) + rr * rr;
  bool tir = delta <= 0.0f;
  // lm_optics::GetReflectRatio, optics_shared.h:17-24
  float dd = std::max(delta, 0.0f);
  float ds = std::sqrt(dd);
  float rs = (rr - ds) / (rr + ds);
  rs *= rs;
  float rp = (1.0f - rr * ds) / (1.0f + rr * ds);
  rp *= rp;
  float ratio = (rs + rp) * 0.5f;
  *w_refl = ratio * w;
  *w_refr = tir ? -1.0f : w - *w_refl;
  float sq = std::sqrt(delta);
  for (int j = 0; j < 3; j++) {
    d_refl[j] = d[j] - 2 * c * nrm[j];
    d_refr[j] = tir ? d_refl[j] : rr * d[j] - (rr - sq) * c * nrm[j];
  }
}
// ---- optics.cpp:64-158 PropagateSlab + traversal_shared.h:60-69 SlabFaceT (one ray) -------------
void Propagate(const HbCrystalTables& t, const float* d, const float* p, int src, float* p_out, uint16_t* face_out) {
  float t_far = 1e30f;
  int far = -1;
  for (uint32_t fi = 0; fi < t.face_cnt; fi++) {
    const float* pl = t.plane[fi];
    float denom = d[0] * pl[0] + d[1] * pl[1] + d[2] * pl[2];
    float tt = 1.0e30f;
    if (!(denom <= kEps)) tt = -(p[0] * pl[0] + p[1] * pl[1] + p[2] * pl[2] + pl[3]) / denom;
    if (tt < t_far) {
      t_far = tt;
      far = static_cast<int>(fi);
    }
  }
  float thr = (src >= 0 && far != src) ? -kEps : kEps;
  if (far >= 0 && t_far > thr) {
    p_out[0] = p[0] + t_far * d[0];
    p_out[1] = p[1] + t_far * d[1];
    p_out[2] = p[2] + t_far * d[2];
    *face_out = static_cast<uint16_t>(far);
  } else {
    p_out[0] = p[0];
    p_out[1] = p[1];
    p_out[2] = p[2];
    *face_out = HB_INVALID_FACE;
  }
}

// ---- filter_shared.h:53-315 -----------------------------------------------------------------------
void PShift(uint8_t* data, uint32_t size) {
  int first = -1;
  for (uint32_t i = 0; i < size; i++) {
    uint8_t x = data[i];
    if (x < 3u) continue;
    uint8_t pyr = static_cast<uint8_t>(x / 10u);
    int pri = static_cast<int>(x % 10u);
    if (first < 0) first = pri;
    pri = (pri + 6 - first) % 6 + 3;
    data[i] = static_cast<uint8_t>(pyr * 10u + static_cast<uint32_t>(pri));
  }
}
bool LexLess(const uint8_t* a, const uint8_t* b, uint32_t n) {
  for (uint32_t i = 0; i < n; i++)
    if (a[i] != b[i]) return a[i] < b[i];
  return false;
}
void ReduceBuffer(uint8_t* data, uint32_t size, uint32_t sym, int sigma_a, bool d_app) {
  if (sym == 0) return;
  if (sym & 1u) PShift(data, size);
  if ((sym & 4u) && d_app) {
    uint8_t sc[64];
    for (uint32_t i = 0; i < size; i++) {
      uint8_t x = data[i];
      if (x < 3u) {
        sc[i] = x;
        continue;
      }
      uint8_t pyr = static_cast<uint8_t>(x / 10u);
      int pri0 = static_cast<int>(x % 10u) - 3;
      int np = ((sigma_a - pri0) % 6 + 6) % 6;
      sc[i] = static_cast<uint8_t>(pyr * 10u + static_cast<uint32_t>(np + 3));
    }
    if (sym & 1u) PShift(sc, size);
    if (LexLess(sc, data, size)) std::memcpy(data, sc, size);
  }
  if (sym & 2u) {
    uint8_t sc[64];
    bool changed = false;
    for (uint32_t i = 0; i < size; i++) {
      uint8_t x = data[i];
      if (x <= 2u) {
        sc[i] = static_cast<uint8_t>(3u - x);
        changed = true;
      } else if (x >= 13u && x <= 18u) {
        sc[i] = static_cast<uint8_t>(x + 10u);
        changed = true;
      } else if (x >= 23u && x <= 28u) {
        sc[i] = static_cast<uint8_t>(x - 10u);
        changed = true;
      } else {
        sc[i] = x;
      }
    }
    if (changed && LexLess(sc, data, size)) std::memcpy(data, sc, size);
  }
}
bool MatchSimple(const HbFilterDesc& f, const HbSimpleFilter& s, const uint8_t* fn_path, uint32_t len,
                 const float* dir, uint32_t crystal_id) {
  bool reduce = !(f.fn_period < 0 || f.symmetry == 0u);
  switch (s.kind) {
    case 0:
      return true;
    case 1: {
      if (len != s.path_len) return false;
      uint8_t buf[64];
      std::memcpy(buf, fn_path, len);
      if (reduce) ReduceBuffer(buf, len, f.symmetry, f.sigma_a, f.d_applicable != 0);
      return std::memcmp(buf, s.path, len) == 0;
    }
    case 2: {
      if (len == 0 || len < s.min_len) return false;
      if (s.max_len != 0 && len > s.max_len) return false;
      bool he = s.entry_fn >= 0, hx = s.exit_fn >= 0;
      if (!he && !hx) return true;
      uint8_t ee[2];
      uint32_t n = 0;
      if (he) ee[n++] = fn_path[0];
      if (hx) ee[n++] = fn_path[len - 1];
      if (reduce) {
        ReduceBuffer(ee, n, f.symmetry, f.sigma_a, f.d_applicable != 0);
        if (n != s.path_len) return false;
      }
      return std::memcmp(ee, s.path, n) == 0;
    }
    case 3:
      return s.dir[0] * dir[0] + s.dir[1] * dir[1] + s.dir[2] * dir[2] > s.cos_radii;
    case 4:
      return crystal_id == s.crystal_id;
    default:
      return false;
  }
}
bool FilterCheck(const HbFilterDesc& f, const uint8_t* fn_path, uint32_t len, const float* dir, uint32_t crystal_id) {
  bool m;
  if (f.kind == 5) {
    m = false;
    for (uint32_t o = 0; o < f.term_cnt && !m; o++) {
      bool ok = true;
      for (uint32_t a = 0; a < f.term_len[o] && ok; a++) ok = MatchSimple(f, f.terms[o][a], fn_path, len, dir, crystal_id);
      m = ok;
    }
  } else {
    m = MatchSimple(f, f.simple, fn_path, len, dir, crystal_id);
  }
  return f.action == 0u ? m : !m;
}

// ---- projection_shared.h:34-375 -------------------------------------------------------------------
struct Hits {
  int px[2], py[2];
  bool bump[2];
  int count;
};
void EqualArea(float dx, float dy, float dz, float rs, float& x, float& y, bool& ok) {
  float k = rs / std::sqrt(1.0f + std::fmin(std::fmax(dz, -1.0f + 1e-6f), 1.0f));
  x = k * dx;
  y = k * dy;
  ok = true;
}
void Equidistant(float dx, float dy, float dz, float rs, float& x, float& y, bool& ok) {
  ok = true;
  float rho = std::sqrt(dx * dx + dy * dy);
  if (rho < 1e-10f) {
    x = y = 0.0f;
    return;
  }
  float theta = std::acos(std::fmin(std::fmax(dz, -1.0f), 1.0f));
  float sc = rs * theta / (kPi2 * rho);
  x = sc * dx;
  y = sc * dy;
}
void Stereographic(float dx, float dy, float dz, float rs, float& x, float& y, bool& ok) {
  ok = true;
  float rho = std::sqrt(dx * dx + dy * dy);
  if (rho < 1e-10f) {
    x = y = 0.0f;
    return;
  }
  float theta = std::acos(std::fmin(std::fmax(dz, -1.0f), 1.0f));
  float sc = rs * std::tan(theta / 2.0f) / rho;
  x = sc * dx;
  y = sc * dy;
}
void Orthographic(float dx, float dy, float dz, float rs, float& x, float& y, bool& ok) {
  if (dz < 0.0f) {
    x = y = 0.0f;
    ok = false;
    return;
  }
  x = rs * dx;
  y = rs * dy;
  ok = true;
}
void FisheyeByType(int base, float dx, float dy, float dz, float rs, float& x, float& y, bool& ok) {
  if (base == 0) EqualArea(dx, dy, dz, rs, x, y, ok);
  else if (base == 1) Equidistant(dx, dy, dz, rs, x, y, ok);
  else if (base == 2) Stereographic(dx, dy, dz, rs, x, y, ok);
  else Orthographic(dx, dy, dz, rs, x, y, ok);
}
void DualToPixel(float xn, float yn, bool upper, int w, int h, float& fx, float& fy) {
  int short_res = std::min(w / 2, h);
  float r = static_cast<float>(short_res) / 2.0f;
  float cy = static_cast<float>(h) / 2.0f;
  if (upper) {
    float cx = static_cast<float>(w) / 2.0f - r;
    fx = -yn * r + cx;
    fy = xn * r + cy;
  } else {
    float cx = static_cast<float>(w) / 2.0f + r;
    fx = yn * r + cx;
    fy = xn * r + cy;
  }
}
Hits Project(const HbProjParams& p, float wx, float wy, float wz) {
  Hits r{};
  r.count = 0;
  int t = p.proj_type;
  bool single = t == 0 || t == 1 || t == 2 || t == 3 || t == 8;
  if (single) {
    if ((p.visible_range == 0 && wz > 0.0f) || (p.visible_range == 1 && wz < 0.0f)) return r;
    float in[3] = { -wx, -wy, -wz }, c[3];
    ApplyRotT(p.rot, in, c);
    float x = 0, y = 0;
    bool ok = false;
    if (t == 0) {
      if (c[2] <= 0.0f) return r;
      x = c[0] / c[2];
      y = c[1] / c[2];
      ok = true;
    } else {
      if (c[2] <= 0.0f) return r;
      int base = t == 1 ? 0 : t == 2 ? 1 : t == 3 ? 2 : 3;
      FisheyeByType(base, c[0], c[1], c[2], 1.0f, x, y, ok);
    }
    if (!ok) return r;
    x = -x;
    r.px[0] = static_cast<int>(std::floor(x * p.scale + static_cast<float>(p.img_w) / 2.0f + 0.5f +
                                          static_cast<float>(p.lens_shift_x)));
    r.py[0] = static_cast<int>(std::floor(y * p.scale + static_cast<float>(p.img_h) / 2.0f + 0.5f +
                                          static_cast<float>(p.lens_shift_y)));
    r.bump[0] = true;
    r.count = 1;
    return r;
  }
  if (t == 7) {
    float lon = std::atan2(-wy, -wx);
    float lat = std::asin(std::fmin(std::fmax(-wz, -1.0f), 1.0f));
    lon = lon - p.az0;
    while (lon < -kPi) lon += 2.0f * kPi;
    while (lon > kPi) lon -= 2.0f * kPi;
    int raw_x = static_cast<int>(std::floor(lon * p.scale + static_cast<float>(p.img_w) / 2.0f + 0.5f));
    r.px[0] = ((raw_x % p.img_w) + p.img_w) % p.img_w;
    r.py[0] = static_cast<int>(std::floor(-lat * p.scale + static_cast<float>(p.img_h) / 2.0f + 0.5f));
    r.bump[0] = true;
    r.count = 1;
    return r;
  }
  if (t == 4 || t == 5 || t == 6 || t == 9) {
    int base = t == 4 ? 0 : t == 5 ? 1 : t == 6 ? 2 : 3;
    float sx = -wx, sy = -wy, sz = -wz;
    bool upper = sz >= 0.0f;
    float zh = upper ? sz : -sz;
    float x, y;
    bool ok;
    FisheyeByType(base, sx, sy, zh, p.r_scale, x, y, ok);
    float fx, fy;
    DualToPixel(x, y, upper, p.img_w, p.img_h, fx, fy);
    r.px[0] = static_cast<int>(std::floor(fx + 0.5f));
    r.py[0] = static_cast<int>(std::floor(fy + 0.5f));
    r.bump[0] = true;
    r.count = 1;
    if (p.max_abs_dz > 0.0f && std::fabs(sz) < p.max_abs_dz) {
      FisheyeByType(base, sx, sy, -zh, p.r_scale, x, y, ok);
      DualToPixel(x, y, !upper, p.img_w, p.img_h, fx, fy);
      r.px[1] = static_cast<int>(std::floor(fx + 0.5f));
      r.py[1] = static_cast<int>(std::floor(fy + 0.5f));
      r.bump[1] = false;
      r.count = 2;
    }
    return r;
  }
  if (t == 10) {
    float in[3] = { -wx, -wy, -wz }, c[3];
    ApplyRotT(p.rot, in, c);
    const float kD = 4.0f;
    if (c[2] >= -1.0f / kD) return r;
    float denom = kD + c[2];
    r.px[0] = static_cast<int>(std::floor(-c[0] / denom * p.scale + static_cast<float>(p.img_w) / 2.0f + 0.5f +
                                          static_cast<float>(p.lens_shift_x)));
    r.py[0] = static_cast<int>(std::floor(c[1] / denom * p.scale + static_cast<float>(p.img_h) / 2.0f + 0.5f +
                                          static_cast<float>(p.lens_shift_y)));
    r.bump[0] = true;
    r.count = 1;
    return r;
  }
  return r;
}

struct Ray {
  float p[3], d[3], w;
  uint16_t face;  // face the ray currently sits on / hits next
  uint32_t root;
  uint32_t code;  // branch history: bit h set when this ray is the far-side (role 0) child that kept tracing at hit h
  uint8_t len;
  uint8_t path[64];  // compact face ids, entry first
};

}  // namespace

extern "C" {

uint32_t orc_pcg_hash(uint32_t x) { return PcgHash(x); }
float orc_draw(uint32_t seed, uint32_t idx, uint32_t slot) { return Draw(seed, idx, slot); }
void orc_quat_to_rot9(const float* q4, float* rot9) { QuatToRot(q4, rot9); }

// feistel_bijection, pcg_shared.h:550-603
uint32_t orc_feistel(uint32_t i, uint32_t n, uint32_t seed) {
  if (n <= 1u) return i;
  if (n == 2u) return i ^ 1u;
  uint32_t bits = 0;
  while (bits < 30u && (1u << bits) < n) bits++;
  if (bits & 1u) bits++;
  uint32_t hb = bits >> 1, hm = (1u << hb) - 1u;
  const uint32_t rc[4] = { 0x9E3779B9u, 0x85EBCA6Bu, 0xC2B2AE35u, 0x27D4EB2Fu };
  uint32_t cur = i;
  for (uint32_t g = 0; g < 64u; g++) {
    uint32_t L = (cur >> hb) & hm, R = cur & hm;
    for (uint32_t k = 0; k < 4u; k++) {
      uint32_t f = PcgHash(seed ^ R ^ rc[k]) & hm;
      uint32_t nr = L ^ f;
      L = R;
      R = nr;
    }
    uint32_t out = (L << hb) | R;
    if (out < n) return out;
    cur = out;
  }
  return cur % n;
}

// ---- the generator's building blocks one by one (pinned against lm_pcg::* of the reference through
// oracle/ref_driver.cpp: tests/test_oracle_vs_reference.py::test_sampler_twin_matches_reference_pcg) ----
uint32_t orc_seed_with_high(uint32_t seed, uint32_t hi) { return SeedWithHigh(seed, hi); }
void orc_uniforms(uint32_t seed, uint32_t idx, uint32_t slot0, uint32_t n, float* out) {
  Stream s{ seed, idx, slot0 };
  for (uint32_t i = 0; i < n; i++) out[i] = s.Next();
}
void orc_get_dist(uint32_t seed, uint32_t idx0, uint32_t n, uint32_t type, float mean, float stdv, float* out) {
  for (uint32_t i = 0; i < n; i++) {
    Stream s{ seed, idx0 + i, 0u };
    out[i] = GetDist(s, type, mean, stdv);
  }
}
void orc_lat_lon_roll(const HbAxisSampler* a, uint32_t seed, uint32_t idx0, uint32_t n, float* lon_lat_roll3,
                      uint32_t* slots_used) {
  for (uint32_t i = 0; i < n; i++) {
    Stream s{ seed, idx0 + i, 0u };
    SampleLonLatRoll(s, *a, lon_lat_roll3[i * 3], lon_lat_roll3[i * 3 + 1], lon_lat_roll3[i * 3 + 2]);
    if (slots_used != nullptr) slots_used[i] = s.slot;
  }
}
void orc_rotation9(uint64_t n, const float* lon_lat_roll3, float* rot9) {  // quaternion form of BuildCrystalRotation
  for (uint64_t i = 0; i < n; i++) {
    float q[4];
    QuatFromAngles(lon_lat_roll3[i * 3], lon_lat_roll3[i * 3 + 1], lon_lat_roll3[i * 3 + 2], q);
    QuatToRot(q, rot9 + i * 9);
  }
}
void orc_sph_cap(uint32_t seed, uint32_t idx0, uint32_t slot0, uint32_t n, float lon, float lat, float half, float* d3) {
  for (uint32_t i = 0; i < n; i++) {
    Stream s{ seed, idx0 + i, slot0 };
    SampleSphCap(s, lon, lat, half, d3 + i * 3);
  }
}
void orc_triangle(uint32_t seed, uint32_t idx0, uint32_t slot0, uint32_t n, const float* vtx9, float* p3) {
  for (uint32_t i = 0; i < n; i++) {
    Stream s{ seed, idx0 + i, slot0 };
    float u = s.Next();
    float v = s.Next();
    if (u + v > 1.0f) {
      u = 1.0f - u;
      v = 1.0f - v;
    }
    for (int k = 0; k < 3; k++) {
      float a = vtx9[k], b = vtx9[3 + k], c = vtx9[6 + k];
      p3[i * 3 + k] = u * (b - a) + v * (c - a) + a;
    }
  }
}

int orc_gen_roots(const HbScene* scene, uint32_t layer, uint32_t pop_i, uint32_t shape_base, const HbWlEntry* wl,
                  uint32_t wl_cnt, uint32_t seed, uint64_t ray_base, uint64_t n, float* d3, float* p3, float* w,
                  uint16_t* face, float* quat4, float* rot9, uint32_t* shape_idx, uint32_t* wl_idx) {
  const HbCrystalPopulation& pop = scene->layers[layer].populations[pop_i];
  for (uint64_t i = 0; i < n; i++) {
    uint64_t g = ray_base + i;
    uint32_t lo = static_cast<uint32_t>(g), hi = static_cast<uint32_t>(g >> 32);
    uint32_t s0 = SeedWithHigh(seed ^ kNonceGen, hi);
    uint32_t wi = 0;
    if (wl_cnt > 1) {
      wi = static_cast<uint32_t>(Draw(s0 ^ kNonceWl, lo, 0) * static_cast<float>(wl_cnt));
      if (wi >= wl_cnt) wi = wl_cnt - 1;
    }
    Stream s{ s0, lo, 0 };
    float lon, lat, roll;
    SampleLonLatRoll(s, pop.axis, lon, lat, roll);
    float q[4], m[9];
    QuatFromAngles(lon, lat, roll, q);
    QuatToRot(q, m);
    float dw[3], dl[3];
    SampleSphCap(s, scene->sun_lon, scene->sun_lat, scene->sun_half_angle, dw);
    ApplyRotT(m, dw, dl);
    uint32_t sh = 0;
    if (pop.shape_cnt > 1) {
      sh = static_cast<uint32_t>(Draw(s0 ^ kNonceShape, lo >> 5, 0)  /* geometry clock: 32 rays per shape */ * static_cast<float>(pop.shape_cnt));
      if (sh >= pop.shape_cnt) sh = pop.shape_cnt - 1;
    }
    const HbCrystalTables& t = pop.shapes[sh];
    float p[3] = { 0, 0, 0 };
    uint16_t f = HB_INVALID_FACE;
    float weight = wl[wi].spd_weight;
    if (t.subtri_cnt == 0) {
      weight = 0.0f;
    } else {
      SampleEntry(s, t, dl, p, &f);
    }
    std::memcpy(d3 + i * 3, dl, 12);
    std::memcpy(p3 + i * 3, p, 12);
    w[i] = weight;
    face[i] = f;
    if (quat4) std::memcpy(quat4 + i * 4, q, 16);
    if (rot9) std::memcpy(rot9 + i * 9, m, 36);
    if (shape_idx) shape_idx[i] = shape_base + sh;
    if (wl_idx) wl_idx[i] = wi;
  }
  return 0;
}

int orc_transit(const HbScene* scene, uint32_t layer, uint32_t pop_i, uint32_t shape_base, uint32_t seed,
                uint64_t ray_base, uint64_t n, const float* d_world3, float* d3, float* p3, uint16_t* face, float* quat4,
                float* rot9, uint32_t* shape_idx) {
  const HbCrystalPopulation& pop = scene->layers[layer].populations[pop_i];
  for (uint64_t i = 0; i < n; i++) {
    uint64_t g = ray_base + i;
    uint32_t lo = static_cast<uint32_t>(g), hi = static_cast<uint32_t>(g >> 32);
    uint32_t s0 = SeedWithHigh(seed ^ kNonceTransit, hi);
    Stream s{ s0, lo, 0 };
    float lon, lat, roll;
    SampleLonLatRoll(s, pop.axis, lon, lat, roll);
    float q[4], m[9];
    QuatFromAngles(lon, lat, roll, q);
    QuatToRot(q, m);
    float dl[3];
    ApplyRotT(m, d_world3 + i * 3, dl);
    uint32_t sh = 0;
    if (pop.shape_cnt > 1) {
      sh = static_cast<uint32_t>(Draw(s0 ^ kNonceShape, lo >> 5, 0)  /* geometry clock: 32 rays per shape */ * static_cast<float>(pop.shape_cnt));
      if (sh >= pop.shape_cnt) sh = pop.shape_cnt - 1;
    }
    const HbCrystalTables& t = pop.shapes[sh];
    float p[3] = { 0, 0, 0 };
    uint16_t f = HB_INVALID_FACE;
    if (t.subtri_cnt != 0) SampleEntry(s, t, dl, p, &f);
    std::memcpy(d3 + i * 3, dl, 12);
    std::memcpy(p3 + i * 3, p, 12);
    face[i] = f;
    if (quat4) std::memcpy(quat4 + i * 4, q, 16);
    if (rot9) std::memcpy(rot9 + i * 9, m, 36);
    if (shape_idx) shape_idx[i] = shape_base + sh;
  }
  return 0;
}

int orc_hit_surface(const HbCrystalTables* t, float n_idx, uint64_t n, const float* d3, const float* w,
                    const uint16_t* face, float* d_out6, float* w_out2) {
  for (uint64_t i = 0; i < n; i++) {
    if (face[i] == HB_INVALID_FACE) {
      w_out2[2 * i] = w_out2[2 * i + 1] = 0.0f;
      continue;
    }
    HitSurface(t->plane[face[i]], n_idx, d3 + i * 3, w[i], d_out6 + i * 6, w_out2 + 2 * i, d_out6 + i * 6 + 3,
               w_out2 + 2 * i + 1);
  }
  return 0;
}

int orc_propagate(const HbCrystalTables* t, uint64_t n, const float* d3, const float* p3, const float* w,
                  const uint16_t* from_face, float* p_out3, uint16_t* to_face) {
  for (uint64_t i = 0; i < n; i++) {
    if (w[i] < 0) continue;  // optics.cpp:135-137
    int src = from_face[i] != HB_INVALID_FACE ? static_cast<int>(from_face[i]) : -1;
    Propagate(*t, d3 + i * 3, p3 + i * 3, src, p_out3 + i * 3, to_face + i);
  }
  return 0;
}

int orc_project(const HbProjParams* p, uint64_t n, const float* dir3, int32_t* px2, int32_t* py2, int32_t* cnt,
                int32_t* bump2) {
  for (uint64_t i = 0; i < n; i++) {
    Hits h = Project(*p, dir3[i * 3], dir3[i * 3 + 1], dir3[i * 3 + 2]);
    cnt[i] = h.count;
    for (int k = 0; k < 2; k++) {
      px2[i * 2 + k] = k < h.count ? h.px[k] : -1;
      py2[i * 2 + k] = k < h.count ? h.py[k] : -1;
      bump2[i * 2 + k] = k < h.count ? (h.bump[k] ? 1 : 0) : 0;
    }
  }
  return 0;
}

int orc_accumulate(const HbProjParams* p, const HbWlEntry* wl, uint32_t wl_cnt, uint64_t n, const float* dir3,
                   const float* w, const uint8_t* wl_idx, float* xyz, double* landed) {
  double land = 0.0;
  for (uint64_t i = 0; i < n; i++) {
    Hits h = Project(*p, dir3[i * 3], dir3[i * 3 + 1], dir3[i * 3 + 2]);
    uint32_t wi = wl_idx ? wl_idx[i] : 0;
    if (wi >= wl_cnt) wi = 0;
    for (int k = 0; k < h.count; k++) {
      int px = h.px[k], py = h.py[k];
      if (px < 0 || px >= p->img_w || py < 0 || py >= p->img_h) continue;
      size_t base = (static_cast<size_t>(py) * p->img_w + px) * 3;
      // AccumXyzToPixel, accum_shared.h:66-72
      xyz[base + 0] += wl[wi].cmf_x * w[i];
      xyz[base + 1] += wl[wi].cmf_y * w[i];
      xyz[base + 2] += wl[wi].cmf_z * w[i];
      if (h.bump[k]) land += w[i];
    }
  }
  *landed += land;
  return 0;
}

int orc_accumulate_lanes(const HbProjParams* p, const HbWlEntry* wl, uint32_t wl_cnt, const HbColorClasses* classes,
                         uint64_t n, const float* dir3, const float* w, const uint8_t* wl_idx, const uint64_t* mask,
                         float* lanes) {
  const size_t stride = static_cast<size_t>(p->img_w) * p->img_h;
  for (uint64_t i = 0; i < n; i++) {
    if (mask[i] == 0) continue;
    Hits h = Project(*p, dir3[i * 3], dir3[i * 3 + 1], dir3[i * 3 + 2]);
    uint32_t wi = wl_idx ? wl_idx[i] : 0;
    if (wi >= wl_cnt) wi = 0;
    const float y = wl[wi].cmf_y * w[i];
    for (int k = 0; k < h.count; k++) {
      int px = h.px[k], py = h.py[k];
      if (px < 0 || px >= p->img_w || py < 0 || py >= p->img_h) continue;
      size_t pix = static_cast<size_t>(py) * p->img_w + px;
      for (uint32_t c = 0; c < classes->class_cnt; c++) {
        uint64_t bits = classes->bits[c];
        if (bits == 0) continue;
        uint64_t m = mask[i] & bits;
        bool ok = ((classes->combine_all_mask >> c) & 1u) ? (m == bits) : (m != 0);
        if (ok) lanes[c * stride + pix] += y;
      }
    }
  }
  return 0;
}

int orc_filter_check(const HbFilterDesc* f, const uint8_t* face_fn, uint32_t crystal_id, uint64_t n,
                     const uint8_t* paths64, const uint8_t* path_len, const float* dir3, uint8_t* pass) {
  for (uint64_t i = 0; i < n; i++) {
    uint8_t fn[64];
    for (uint32_t k = 0; k < path_len[i]; k++) fn[k] = face_fn[paths64[i * 64 + k]];
    pass[i] = FilterCheck(*f, fn, path_len[i], dir3 + i * 3, crystal_id) ? 1 : 0;
  }
  return 0;
}

// The hit loop: simulator.cpp:1308-1336 (legacy) == cpu_trace_backend.cpp:155-202; per hit
// TraceRayBasicInfo (simulator.cpp:585-642) + FillRayOtherInfo (:645-652) + CollectData (:665-762).
int orc_trace_layer(const OrcLayerParams* lp, uint64_t n, const float* d3, const float* p3, const float* w,
                    const uint16_t* face, const float* rot9, const uint32_t* shape_idx, const uint32_t* wl_idx,
                    uint64_t cap, HbExitRecord* exits, uint32_t* exit_root, uint64_t* exit_cnt, float* cont_d3,
                    float* cont_w, uint32_t* cont_wl, uint32_t* cont_root, uint64_t* cont_cnt) {
  uint64_t ne = 0, nc = 0;
  const float ident[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
  std::vector<Ray> cur, nxt;
  for (uint64_t i = 0; i < n; i++) {
    uint32_t sh = shape_idx ? shape_idx[i] : 0;
    uint32_t wi = wl_idx ? wl_idx[i] : 0;
    const HbCrystalTables& t = lp->shapes[sh];
    uint32_t pop_i = lp->shape_pop ? lp->shape_pop[sh] : 0;
    const HbCrystalPopulation* pop = lp->pops ? &lp->pops[pop_i] : nullptr;
    float n_idx = lp->wl[wi < lp->wl_cnt ? wi : 0].n_idx;
    const float* rot = rot9 ? rot9 + i * 9 : ident;
    uint64_t g = lp->gate_base + i;
    uint32_t glo = static_cast<uint32_t>(g), ghi = static_cast<uint32_t>(g >> 32);
    uint32_t gseed = SeedWithHigh(lp->seed ^ kNonceGate, ghi);

    cur.clear();
    Ray r0{};
    std::memcpy(r0.p, p3 + i * 3, 12);
    std::memcpy(r0.d, d3 + i * 3, 12);
    r0.w = w[i];
    r0.face = face[i];
    r0.root = static_cast<uint32_t>(i);
    r0.code = 0;
    r0.len = 0;
    if (r0.face != HB_INVALID_FACE) r0.path[r0.len++] = static_cast<uint8_t>(r0.face);  // InitRay_other_info :270
    if (r0.face == HB_INVALID_FACE || !(r0.w >= 0)) continue;  // degenerate entry: nothing traced (w = 0 root)
    cur.push_back(r0);

    for (uint32_t hit = 0; hit < lp->max_hits && !cur.empty(); hit++) {
      nxt.clear();
      for (const Ray& r : cur) {
        float dc[2][3], wc[2];
        HitSurface(t.plane[r.face], n_idx, r.d, r.w, dc[0], &wc[0], dc[1], &wc[1]);
        // Engine bookkeeping (no effect on the traced geometry): the child on the far side of the face is
        // "role 0" (reflection at the entry hit, refraction at an internal hit), the other one "role 1";
        // gate draws are keyed by role, and a role-0 child that keeps tracing is a fork (branch-code bit).
        const float* pn0 = t.plane[r.face];
        const float cos_in = r.d[0] * pn0[0] + r.d[1] * pn0[1] + r.d[2] * pn0[2];
        const int out_child = cos_in > 0.0f ? 1 : 0;
        for (int c = 0; c < 2; c++) {
          if (wc[c] < 0) continue;  // TIR sentinel: Propagate skips it, CollectData drops it
          float pn[3];
          uint16_t fn;
          Propagate(t, dc[c], r.p, static_cast<int>(r.face), pn, &fn);
          if (fn != HB_INVALID_FACE) {
            Ray ch = r;
            std::memcpy(ch.p, pn, 12);
            std::memcpy(ch.d, dc[c], 12);
            ch.w = wc[c];
            ch.face = fn;
            if (ch.len < 64) ch.path[ch.len++] = static_cast<uint8_t>(fn);  // FillRayOtherInfo
            if (c == out_child) ch.code = r.code | (1u << (hit & 31u));
            nxt.push_back(ch);
            continue;
          }
          // Outgoing candidate (CollectData branch 1): rotate to world, filter, colour (n/a), prob gate.
          float dw[3];
          ApplyRot(rot, dc[c], dw);
          uint8_t fnp[64];
          for (uint32_t k = 0; k < r.len; k++) fnp[k] = t.face_fn[r.path[k]];
          bool pass = true;
          if (pop != nullptr && pop->filter.kind != 0) pass = FilterCheck(pop->filter, fnp, r.len, dw, pop->crystal_id);
          if (!pass) continue;  // filter-fail terminates
          // Colour pass (simulator.cpp:688-712): non-destructive, after the physical filter, before the gate
          uint64_t mask = lp->root_mask ? lp->root_mask[i] : 0ull;
          if (pop != nullptr) {
            for (uint32_t g = 0; g < pop->color_group_cnt; g++) {
              const HbColorGroup& cg = pop->color_groups[g];
              for (uint32_t k = 0; k < cg.filter.term_cnt; k++) {
                if (cg.bit[k] < 64 && MatchSimple(cg.filter, cg.filter.terms[k][0], fnp, r.len, dw, pop->crystal_id))
                  mask |= 1ull << cg.bit[k];
              }
            }
          }
          uint32_t es = r.code == 0 ? gseed : gseed ^ PcgHash(r.code);
          float u = Draw(es, glo, hit * 2u + (c == out_child ? 0u : 1u));
          if (u < lp->prob) {
            if (cont_d3 != nullptr && nc < cap) {
              std::memcpy(cont_d3 + nc * 3, dw, 12);
              cont_w[nc] = wc[c];
              cont_wl[nc] = wi;
              cont_root[nc] = static_cast<uint32_t>(i);
              if (lp->cont_mask != nullptr) lp->cont_mask[nc] = mask;
            }
            nc++;
          } else {
            if (exits != nullptr && ne < cap) {
              HbExitRecord& e = exits[ne];
              std::memset(&e, 0, sizeof(e));
              std::memcpy(e.dir, dw, 12);
              e.weight = wc[c];
              e.path_len = r.len;
              std::memcpy(e.path, fnp, r.len);
              e.crystal_id = static_cast<uint16_t>(pop_i);
              e.ms_layer_idx = static_cast<uint8_t>(lp->layer_idx);
              e.wl_idx = static_cast<uint8_t>(wi);
              e.component_mask = mask;
              exit_root[ne] = static_cast<uint32_t>(i);
            }
            ne++;
          }
        }
      }
      cur.swap(nxt);
    }
  }
  *exit_cnt = ne;
  if (cont_cnt) *cont_cnt = nc;
  return (ne > cap || nc > cap) ? HB_ERR_CAPACITY : 0;
}

// ---- display sink (server/render.cpp:508-577; util/color_space.cpp:10-52; util/color_data.hpp:6-13) ----
static const float kOrcWhiteD65[3] = { 0.95047f, 1.00000f, 1.08883f };
static const float kOrcXyzToRgb[9] = { 3.2404542f, -1.5371385f, -0.4985314f, -0.9692660f, 1.8760108f,
                                       0.0415560f, 0.0556434f,  -0.2040259f, 1.0572252f };

static float orc_linear_to_srgb(float v) {
  if (v < 0.0031308f) return v * 12.92f;
  return 1.055f * std::pow(v, 1.0f / 2.4f) - 0.055f;
}

int orc_post_snapshot(const float* xyz_img, int w, int h, float snapshot_intensity, float intensity_factor,
                      const float* ray_color, const float* background, uint8_t* out) {
  const int total_pix = w * h;
  if (total_pix <= 0 || snapshot_intensity <= 0) {
    std::memset(out, 0, static_cast<size_t>(std::max(total_pix, 0)) * 3);
    return 0;
  }
  const float scale = intensity_factor * 0.08f * total_pix / snapshot_intensity;
  const bool real_color = ray_color[0] < 0;
  for (int i = 0; i < total_pix; i++) {
    float xyz[3], gray[3], rgb[3];
    for (int j = 0; j < 3; j++) xyz[j] = xyz_img[i * 3 + j] * scale;
    for (int j = 0; j < 3; j++) gray[j] = kOrcWhiteD65[j] * xyz[1];
    if (real_color) {
      float s = 1.0f, diff[3], clipped[3];
      for (int j = 0; j < 3; j++) diff[j] = xyz[j] - gray[j];
      for (int j = 0; j < 3; j++) {
        float a = 0, b = 0;
        for (int k = 0; k < 3; k++) {
          a += -gray[k] * kOrcXyzToRgb[j * 3 + k];
          b += diff[k] * kOrcXyzToRgb[j * 3 + k];
        }
        if (a * b > 0 && a / b < s) s = a / b;
      }
      for (int j = 0; j < 3; j++) clipped[j] = diff[j] * s + gray[j];
      for (int j = 0; j < 3; j++) {
        float v = 0;
        for (int k = 0; k < 3; k++) v += clipped[k] * kOrcXyzToRgb[j * 3 + k];
        rgb[j] = std::min(std::max(v, 0.0f), 1.0f);
      }
    } else {
      for (int j = 0; j < 3; j++) {
        float v = 0;
        for (int k = 0; k < 3; k++) v += gray[k] * kOrcXyzToRgb[j * 3 + k];
        rgb[j] = v * ray_color[j];
      }
    }
    for (int j = 0; j < 3; j++) {
      float v = rgb[j] + background[j];
      v = std::min(std::max(v, 0.0f), 1.0f);
      out[i * 3 + j] = static_cast<uint8_t>(orc_linear_to_srgb(v) * 255);
    }
  }
  return 0;
}

}  // extern "C"
