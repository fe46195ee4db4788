// hb_host.cpp — host-side table builders of the B200 trace engine (no CUDA in this file).
//
// These turn a config-level scene description into the POD tables the kernels consume. When the
// engine runs behind the reference's TraceBackend seam the adapter fills the same tables with the
// reference's own host code (MakeCrystal, BuildEntrySubTris, GetSharedLatLut, BuildProjParams,
// ComputeWlPool); the builders here make the library usable stand-alone and are parity-tested
// against the reference's tables (tests/test_host_tables.py, tests/golden/).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "halotrace_b200.h"
#include "hb_filter.h"
#include "hb_geometry.h"

namespace hb {

std::string& global_error() {
  static thread_local std::string e;
  return e;
}

namespace {

constexpr double kPiD = 3.14159265358979323846;
constexpr float kPiF = 3.14159265359f;           // reference math::kPi (src/core/math.hpp:20)
constexpr float kDeg2RadF = kPiF / 180.0f;       // math::kDegreeToRad
constexpr float kSqrt3F = 1.73205080757f;        // math::kSqrt3
constexpr float kFloatEps = 1e-5f;               // math::kFloatEps
}  // namespace
}  // namespace hb

using namespace hb;  // NOLINT

extern "C" {

uint32_t hb_abi_version(void) { return HB_ABI_VERSION; }

namespace {
int GeometryStatus(int rc) {
  if (rc == HB_ERR_CAPACITY) global_error() = "crystal face has more than 12 corners or the crystal needs more than 64 entry sub-triangles";
  return rc;
}
}  // namespace

int hb_make_prism(float h, const float dist6[6], HbCrystalTables* out) {
  if (out == nullptr || dist6 == nullptr) return HB_ERR_INVALID_ARG;
  return GeometryStatus(hbg::MakePrism(h, dist6, out));
}

double hb_pyramid_slope(float alpha_deg) {
  // geo3d_closedform.cpp ComputeClosedFormPyramidInner: a = (sqrt3/4)/tan(alpha) for alpha in [0.1, 89.9] deg
  const float sqrt3_4 = kSqrt3F / 4.0f;
  if (!(alpha_deg >= 0.1f && alpha_deg <= 89.9f)) return -1.0;
  return static_cast<double>(sqrt3_4) / std::tan(static_cast<double>(alpha_deg) * static_cast<double>(kDeg2RadF));
}

int hb_make_pyramid(float upper_alpha, float lower_alpha, float h1, float h2, float h3, const float dist[6],
                    HbCrystalTables* out) {
  if (out == nullptr || dist == nullptr) return HB_ERR_INVALID_ARG;
  return GeometryStatus(hbg::MakePyramidFromSlopes(hb_pyramid_slope(upper_alpha), hb_pyramid_slope(lower_alpha), h1, h2, h3, dist, out));
}

// IceRefractiveIndex::Get, optics.cpp:180-198 (Sellmeier fit, valid 350..900 nm, float coefficients).
double hb_ice_refractive_index(double wl) {
  const float coef[4] = { 0.701777f, 1.091144f, 0.884400f, 0.796950f };
  if (wl < 350.0f || wl > 900.0f) return 1.0f;
  wl /= 1e3;
  double n = 1.0;
  n += coef[0] / (1 - coef[2] * 1e-2f / wl / wl);
  n += coef[1] / (1 - coef[3] * 1e2f / wl / wl);
  return std::sqrt(n);
}

}  // extern "C"

// ---- CIE 1931 2-degree observer, 1 nm, 360..830 nm -------------------------------------------------
namespace hb {
namespace {
const float kCie1931[471][3] = {
#include "cie1931_2deg_1nm.inc"
};
// CIE daylight basis S0/S1/S2, 5 nm, 300..830 nm
const float kDaylightBasis[107][3] = {
#include "cie_daylight_basis_5nm.inc"
};

// --- latitude inverse-CDF LUT: area-measure mass of the latitude proposal over colatitude,
// resampled to 257 uniform-theta nodes (reference: lat_lut.cpp:74-180). ---
void FoldLatitude(float phi, float& phi_out, bool& flip) {  // pcg_shared.h:311-322
  const float pi = 3.14159265358979323846f, pi2 = 1.5707963267948966f;
  float theta = pi2 - phi;
  theta = std::fmod(theta, 2.0f * pi);
  if (theta < 0.0f) theta += 2.0f * pi;
  flip = theta > pi;
  if (flip) theta = 2.0f * pi - theta;
  phi_out = pi2 - theta;
}

void DegenerateLut(double colat, HbAxisSampler* out) {
  float c = static_cast<float>(std::min(std::max(colat, 0.0), kPiD));
  for (uint32_t i = 0; i < HB_LUT_NODES; i++) {
    out->lut_theta[i] = c;
    out->lut_cdf[i] = static_cast<float>(i) / static_cast<float>(HB_LUT_NODES - 1);
    out->lut_flip[i] = 0.0f;
  }
}

void BuildLatLut(uint32_t type, float center_deg, float spread_deg, HbAxisSampler* out) {
  const int kFine = 4096, kQuad = 1 << 16;
  const double mean = static_cast<double>(center_deg) * (kPiD / 180.0);
  const double scale = static_cast<double>(spread_deg) * (kPiD / 180.0);
  const double dth = kPiD / kFine;
  std::vector<double> mass(kFine, 0.0), fmass(kFine, 0.0);
  auto add = [&](double lat, double weight) {
    float phi = 0;
    bool flip = false;
    FoldLatitude(static_cast<float>(lat), phi, flip);
    double th = kPiD / 2.0 - static_cast<double>(phi);
    double w = weight * std::sin(th);
    if (w <= 0.0) return;
    int bin = std::min(std::max(static_cast<int>(th / dth), 0), kFine - 1);
    mass[bin] += w;
    if (flip) fmass[bin] += w;
  };
  if (type == HB_DIST_GAUSSIAN) {
    double lo = mean - 12.0 * scale, hi = mean + 12.0 * scale, dl = (hi - lo) / kQuad;
    double inv = scale > 0.0 ? 1.0 / (2.0 * scale * scale) : 0.0;
    for (int i = 0; i < kQuad; i++) {
      double l = lo + (i + 0.5) * dl, d = l - mean;
      add(l, std::exp(-d * d * inv) * dl);
    }
  } else {
    double du = 1.0 / kQuad;
    for (int i = 0; i < kQuad; i++) {
      double u = (i + 0.5) * du, l = mean;
      if (type == HB_DIST_UNIFORM) {
        l = (u - 0.5) * scale + mean;
      } else if (type == HB_DIST_ZIGZAG) {
        l = std::fabs(scale * std::sin(u * 2.0 * kPiD) + mean);
      } else if (type == HB_DIST_LAPLACIAN) {
        double sg = u < 0.5 ? -1.0 : 1.0;
        l = mean - scale * sg * std::log(std::max(1.0 - 2.0 * std::fabs(u - 0.5), 1e-30));
      }
      add(l, du);
    }
  }
  std::vector<double> cm(kFine + 1, 0.0), cf(kFine + 1, 0.0);
  for (int i = 0; i < kFine; i++) {
    cm[i + 1] = cm[i] + mass[i];
    cf[i + 1] = cf[i] + fmass[i];
  }
  auto lerp = [&](const std::vector<double>& c, double th) {
    double x = th / dth;
    int i = static_cast<int>(x);
    if (i < 0) return c.front();
    if (i >= kFine) return c.back();
    double f = x - i;
    return c[i] * (1.0 - f) + c[i + 1] * f;
  };
  double total = cm[kFine];
  if (!(total > 0.0)) {
    float phi = 0;
    bool flip = false;
    FoldLatitude(static_cast<float>(mean), phi, flip);
    DegenerateLut(kPiD / 2.0 - static_cast<double>(phi), out);
    return;
  }
  double tlo = 0.0, thi = kPiD;
  for (int i = 0; i <= kFine; i++)
    if (cm[i] / total >= 1e-7) {
      tlo = i * dth;
      break;
    }
  for (int i = kFine; i >= 0; i--)
    if (cm[i] / total <= 1.0 - 1e-7) {
      thi = i * dth;
      break;
    }
  if (!(thi > tlo)) {
    DegenerateLut(0.5 * (tlo + thi), out);
    return;
  }
  const uint32_t N = HB_LUT_NODES;
  for (uint32_t n = 0; n < N; n++) {
    double t = tlo + (thi - tlo) * n / (N - 1);
    out->lut_theta[n] = static_cast<float>(t);
    out->lut_cdf[n] = static_cast<float>(lerp(cm, t) / total);
  }
  for (uint32_t n = 1; n < N; n++)
    if (out->lut_cdf[n] <= out->lut_cdf[n - 1])
      out->lut_cdf[n] = std::nextafter(out->lut_cdf[n - 1], std::numeric_limits<float>::infinity());
  for (uint32_t n = 0; n + 1 < N; n++) {
    double t0 = out->lut_theta[n], t1 = out->lut_theta[n + 1];
    double m = lerp(cm, t1) - lerp(cm, t0), fm = lerp(cf, t1) - lerp(cf, t0);
    out->lut_flip[n] = m > 0.0 ? static_cast<float>(std::min(std::max(fm / m, 0.0), 1.0)) : 0.0f;
  }
  out->lut_flip[N - 1] = out->lut_flip[N - 2];
}

bool FloatEq(float a, float b) { return std::fabs(a - b) < kFloatEps; }

// Rotation::FillMat / Chain, geo3d.cpp:35-111 (float arithmetic, left-multiply).
void AxisAngle(const float* ax, float th, float* m) {
  float c = std::cos(th), s = std::sin(th), cc = 1 - c;
  m[0] = ax[0] * ax[0] * cc + c;
  m[1] = ax[0] * ax[1] * cc - ax[2] * s;
  m[2] = ax[0] * ax[2] * cc + ax[1] * s;
  m[3] = ax[0] * ax[1] * cc + ax[2] * s;
  m[4] = ax[1] * ax[1] * cc + c;
  m[5] = ax[1] * ax[2] * cc - ax[0] * s;
  m[6] = ax[0] * ax[2] * cc - ax[1] * s;
  m[7] = ax[1] * ax[2] * cc + ax[0] * s;
  m[8] = ax[2] * ax[2] * cc + c;
}
void ChainLeft(float* m, const float* r) {
  float m0[9];
  std::memcpy(m0, m, 36);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      m[i * 3 + j] = 0;
      for (int k = 0; k < 3; k++) m[i * 3 + j] += r[i * 3 + k] * m0[k * 3 + j];
    }
}

struct SceneStorage {
  HbScene scene{};
  std::vector<HbLayer> layers;
  std::vector<std::vector<HbCrystalPopulation>> pops;
  std::vector<std::unique_ptr<std::vector<HbCrystalTables>>> shapes;
};

float DrawDist(std::mt19937& gen, const HbDist& d) {  // RandomNumberGenerator::Get, math.cpp:419-444
  std::uniform_real_distribution<float> uni(0.0f, 1.0f);
  std::normal_distribution<float> gau(0.0f, 1.0f);
  switch (d.type) {
    case HB_DIST_UNIFORM:
      return (uni(gen) - 0.5f) * d.spread + d.center;
    case HB_DIST_GAUSSIAN:
    case HB_DIST_GAUSSIAN_LEGACY:
      return gau(gen) * d.spread + d.center;
    case HB_DIST_ZIGZAG:
      return std::fabs(d.spread * std::sin(uni(gen) * 2.0f * kPiF) + d.center);
    case HB_DIST_LAPLACIAN: {
      float u = uni(gen), sg = u < 0.5f ? -1.0f : 1.0f;
      float arg = std::max(1.0f - 2.0f * std::fabs(u - 0.5f), std::numeric_limits<float>::min());
      return d.center - d.spread * sg * std::log(arg);
    }
    default:
      return d.center;
  }
}

bool ShapeIsDeterministic(const HbCrystalDesc& c) {  // IsDeterministic, simulator.cpp:453-471
  int hn = c.kind == 0 ? 1 : 3;
  for (int i = 0; i < hn; i++)
    if (c.height[i].type != HB_DIST_NO_RANDOM) return false;
  for (int i = 0; i < 6; i++)
    if (c.face_dist[i].type != HB_DIST_NO_RANDOM) return false;
  return true;
}

int MakeShape(std::mt19937& gen, const HbCrystalDesc& c, HbCrystalTables* out) {  // MakeCrystal, simulator.cpp:405-450
  float hgt[3], dist[6];
  hbg::SampleShapeScalars(c, [&](const HbDist& d) { return DrawDist(gen, d); }, hgt, dist);
  if (c.kind == 0) return hb_make_prism(hgt[0], dist, out);
  return hb_make_pyramid(c.wedge_upper_deg, c.wedge_lower_deg, hgt[0], hgt[1], hgt[2], dist, out);
}

// BuildDeviceFilterDesc, device_filter_desc.cpp:130-143 (+ crystal.cpp:710-730 D-symmetry helpers).
void BuildFilter(const HbPopulationDesc& p, const HbFilterSpecDesc& f, HbFilterDesc* out) {
  std::memset(out, 0, sizeof(*out));
  out->kind = f.kind;
  out->action = f.action;
  out->symmetry = f.symmetry;
  out->fn_period = 6;
  const HbDist& az = p.crystal.azimuth;
  bool az_sym = az.type == HB_DIST_UNIFORM && FloatEq(az.spread, 360.0f);
  float rem = std::fmod(std::fmod(p.crystal.roll.center, 30.0f) + 30.0f, 30.0f);
  bool roll30 = FloatEq(rem, 0.0f) || FloatEq(rem, 30.0f);
  bool d_app = az_sym && roll30;
  out->d_applicable = d_app ? 1u : 0u;
  int sigma = 0;
  if (d_app && !(std::fabs(p.crystal.roll.center) > 1e6f)) {
    int n = (static_cast<int>(std::round(p.crystal.roll.center / 30.0f)) % 6 + 6) % 6;
    sigma = (6 - n) % 6;
  }
  out->sigma_a = sigma;
  auto fill = [&](const HbSimpleFilterSpec& f, HbSimpleFilter& s) {
    std::memset(&s, 0, sizeof(s));
    s.kind = f.kind;
    s.entry_fn = -1;
    s.exit_fn = -1;
    if (f.kind == 1) {
      s.path_len = std::min<uint32_t>(f.path_len, HB_MAX_FILTER_PATH);
      std::memcpy(s.path, f.path, s.path_len);
      filter_reduce(s.path, s.path_len, out->symmetry, sigma, d_app);
    } else if (f.kind == 2) {
      s.entry_fn = f.entry_fn >= 0 ? 1 : -1;
      s.exit_fn = f.exit_fn >= 0 ? 1 : -1;
      s.min_len = f.min_len == 0 ? 1 : f.min_len;
      s.max_len = f.max_len;
      uint32_t n = 0;
      if (f.entry_fn >= 0) s.path[n++] = static_cast<uint8_t>(f.entry_fn);
      if (f.exit_fn >= 0) s.path[n++] = static_cast<uint8_t>(f.exit_fn);
      s.path_len = n;
      if (n > 0) filter_reduce(s.path, n, out->symmetry, sigma, d_app);
    } else if (f.kind == 3) {
      float lon = f.lon_deg * kDeg2RadF, lat = f.lat_deg * kDeg2RadF;
      s.dir[0] = std::cos(lat) * std::cos(lon);
      s.dir[1] = std::cos(lat) * std::sin(lon);
      s.dir[2] = std::sin(lat);
      s.cos_radii = std::cos(f.radii_deg * kDeg2RadF);
    } else if (f.kind == 4) {
      s.crystal_id = f.crystal_id;
    }
  };
  if (f.kind == 5) {  // BuildComplexSubDescs, device_filter_desc.cpp:146-166: sub-filters inherit symmetry
    out->term_cnt = std::min<uint32_t>(f.term_cnt, HB_MAX_FILTER_TERMS);
    for (uint32_t o = 0; o < out->term_cnt; o++) {
      out->term_len[o] = std::min<uint32_t>(f.term_len[o], 4u);
      for (uint32_t a = 0; a < out->term_len[o]; a++) fill(f.terms[o][a], out->terms[o][a]);
    }
  } else {
    fill(f.simple, out->simple);
  }
}

// BuildColorSpecGroups (filter_spec.cpp:389-425) over GroupPlacementBySymmetry (color_gate_table.hpp): the
// population's colour predicates grouped by symmetry value in first-occurrence order; each group becomes one
// complex descriptor with a single-factor OR-term per predicate.
int BuildColorGroups(const HbPopulationDesc& p, HbCrystalPopulation* out) {
  out->color_group_cnt = 0;
  if (p.color_pred_cnt > HB_MAX_COLOR_PREDS) return HB_ERR_INVALID_ARG;
  std::vector<uint32_t> group_sym;
  std::vector<HbFilterSpecDesc> specs;
  for (uint32_t k = 0; k < p.color_pred_cnt; k++) {
    const HbColorPredDesc& cp = p.color_preds[k];
    if (cp.pred.kind > 4u) return HB_ERR_INVALID_ARG;
    size_t gi = 0;
    while (gi < group_sym.size() && group_sym[gi] != cp.symmetry) gi++;
    if (gi == group_sym.size()) {
      if (gi == HB_MAX_COLOR_GROUPS) return HB_ERR_UNSUPPORTED;  // kColorMaxGroupsPerSlot
      group_sym.push_back(cp.symmetry);
      HbFilterSpecDesc fs;
      std::memset(&fs, 0, sizeof(fs));
      fs.kind = 5;
      fs.symmetry = cp.symmetry;
      specs.push_back(fs);
      std::memset(out->color_groups[gi].bit, 0xFF, sizeof(out->color_groups[gi].bit));
    }
    HbFilterSpecDesc& fs = specs[gi];
    if (fs.term_cnt == HB_MAX_FILTER_TERMS) return HB_ERR_UNSUPPORTED;  // kDeviceFilterMaxOrClauses
    fs.term_len[fs.term_cnt] = 1;
    fs.terms[fs.term_cnt][0] = cp.pred;
    out->color_groups[gi].bit[fs.term_cnt] = static_cast<uint8_t>(std::min<uint32_t>(cp.bit, 255u));
    fs.term_cnt++;
  }
  for (size_t gi = 0; gi < specs.size(); gi++) BuildFilter(p, specs[gi], &out->color_groups[gi].filter);
  out->color_group_cnt = static_cast<uint32_t>(specs.size());
  return HB_OK;
}

}  // namespace
}  // namespace hb

struct HbSceneTables {
  hb::SceneStorage st;
};

extern "C" {

int hb_make_wl_entry(float wl, float weight, HbWlEntry* out) {
  if (out == nullptr) return HB_ERR_INVALID_ARG;
  out->n_idx = static_cast<float>(hb_ice_refractive_index(wl));  // Crystal::GetRefractiveIndex, crystal.cpp:703
  out->spd_weight = weight;
  int key = static_cast<int>(wl + 0.5f);  // ComputeCmf, wl_pool.hpp:47-58
  if (key < 360 || key > 830) {
    out->cmf_x = out->cmf_y = out->cmf_z = 0.0f;
  } else {
    out->cmf_x = kCie1931[key - 360][0];
    out->cmf_y = kCie1931[key - 360][1];
    out->cmf_z = kCie1931[key - 360][2];
  }
  return HB_OK;
}

float hb_illuminant_spd(int illuminant, float wl) {
  if (illuminant < 0 || illuminant > HB_ILLUMINANT_E) return 0.0f;
  if (wl < 300.0f || wl > 830.0f) return 0.0f;  // illuminant.cpp:62-64,122-131
  if (illuminant == HB_ILLUMINANT_E) return 1.0f;
  if (illuminant == HB_ILLUMINANT_A) {  // Planck 2856 K relative to 560 nm, illuminant.cpp:95-104
    const float ref_wl = 560.0f, temp = 2856.0f, c2 = 1.4388e7f;
    float ratio = ref_wl / wl;
    float ratio5 = ratio * ratio * ratio * ratio * ratio;
    float exp_ref = std::exp(c2 / (temp * ref_wl));
    float exp_lam = std::exp(c2 / (temp * wl));
    return 100.0f * ratio5 * (exp_ref - 1.0f) / (exp_lam - 1.0f);
  }
  // D-series: chromaticity from the correlated colour temperature, then S0 + M1*S1 + M2*S2
  // (CIE 015; illuminant.cpp:14-42,62-90).
  const float cct_of[4] = { 5003.0f, 5503.0f, 6504.0f, 7504.0f };
  float cct = cct_of[illuminant];
  float ti = 1.0f / cct, ti2 = ti * ti, ti3 = ti2 * ti;
  float xd = cct <= 7000.0f ? 0.244063f + 0.09911e3f * ti + 2.9678e6f * ti2 - 4.6070e9f * ti3
                            : 0.237040f + 0.24748e3f * ti + 1.9018e6f * ti2 - 2.0064e9f * ti3;
  float yd = -3.000f * xd * xd + 2.870f * xd - 0.275f;
  float denom = 0.0241f + 0.2562f * xd - 0.7341f * yd;
  float m1 = (-1.3515f - 1.7703f * xd + 5.9114f * yd) / denom;
  float m2 = (0.0300f - 31.4424f * xd + 30.0717f * yd) / denom;
  float fi = (wl - 300.0f) / 5.0f;
  int i0 = static_cast<int>(fi);
  float frac = fi - static_cast<float>(i0);
  if (i0 >= 106) {
    i0 = 106;
    frac = 0.0f;
  }
  int i1 = i0 + (i0 < 106 ? 1 : 0);
  float basis[3];
  for (int k = 0; k < 3; k++) basis[k] = kDaylightBasis[i0][k] + frac * (kDaylightBasis[i1][k] - kDaylightBasis[i0][k]);
  return basis[0] + m1 * basis[1] + m2 * basis[2];
}

int hb_make_wl_pool_illuminant(int illuminant, uint32_t m, HbWlEntry* out) {
  if (out == nullptr || m == 0 || m > 255 || illuminant < 0 || illuminant > HB_ILLUMINANT_E) return HB_ERR_INVALID_ARG;
  for (uint32_t i = 0; i < m; i++) {
    float wl = 380.0f + (static_cast<float>(i) + 0.5f) * 400.0f / static_cast<float>(m);
    hb_make_wl_entry(wl, hb_illuminant_spd(illuminant, wl), &out[i]);
  }
  return HB_OK;
}

int hb_make_axis_sampler(uint32_t lat_type, float lat_c, float lat_s, uint32_t az_type, float az_c, float az_s,
                         uint32_t roll_type, float roll_c, float roll_s, HbAxisSampler* out) {
  if (out == nullptr) return HB_ERR_INVALID_ARG;
  std::memset(out, 0, sizeof(*out));
  // SelectLatPath, lat_path_selection.hpp:67-81 + AxisDistribution::IsFullSphereUniform, math.cpp:555-559
  bool full = az_type == HB_DIST_UNIFORM && FloatEq(az_c, 0.0f) && FloatEq(az_s, 360.0f) &&
              lat_type == HB_DIST_UNIFORM && FloatEq(lat_c, 90.0f) && FloatEq(lat_s, 360.0f);
  if (full) out->lat_path = HB_LAT_FULL_SPHERE;
  else if (lat_type == HB_DIST_NO_RANDOM) out->lat_path = HB_LAT_NO_RANDOM;
  else if (lat_type == HB_DIST_GAUSSIAN_LEGACY) out->lat_path = HB_LAT_GAUSS_LEGACY;
  else out->lat_path = HB_LAT_LUT;
  out->lat_mean = lat_c * kDeg2RadF;
  out->lat_std = lat_s * kDeg2RadF;
  out->az_type = az_type;
  out->az_mean = az_c * kDeg2RadF;
  out->az_std = az_s * kDeg2RadF;
  out->roll_type = roll_type;
  out->roll_mean = roll_c * kDeg2RadF;
  out->roll_std = roll_s * kDeg2RadF;
  if (out->lat_path == HB_LAT_LUT) {
    out->lut_n = HB_LUT_NODES;
    BuildLatLut(lat_type, lat_c, lat_s, out);
  }
  return HB_OK;
}

int hb_make_proj_params(int lens_type, float fov_deg, int img_w, int img_h, float az, float el, float ro,
                        int visible, int shift_x, int shift_y, float overlap, HbProjParams* out) {
  if (out == nullptr || img_w <= 0 || img_h <= 0 || lens_type < 0 || lens_type > 10) return HB_ERR_INVALID_ARG;
  std::memset(out, 0, sizeof(*out));
  // MakeCameraRotation, scatter_accum.hpp:18-26
  float m[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 }, r[9];
  const float ez[3] = { 0, 0, 1 }, ey[3] = { 0, 1, 0 };
  AxisAngle(ez, (-90.0f + ro) * kDeg2RadF, r);
  ChainLeft(m, r);
  AxisAngle(ey, (90.0f - el) * kDeg2RadF, r);
  ChainLeft(m, r);
  AxisAngle(ez, az * kDeg2RadF, r);
  ChainLeft(m, r);
  // BuildProjParams + ComputeScaleAz0, lens_proj_build.hpp:24-140
  out->proj_type = lens_type;
  out->img_w = img_w;
  out->img_h = img_h;
  out->visible_range = visible;
  out->lens_shift_x = shift_x;
  out->lens_shift_y = shift_y;
  out->r_scale = 1.0f;
  out->max_abs_dz = 0.0f;
  std::memcpy(out->rot, m, 36);
  float short_pix = static_cast<float>(std::min(img_w, img_h));
  float fov = fov_deg * kDeg2RadF;
  const float pi_2 = kPiF / 2.0f;
  out->scale = 1.0f;
  out->az0 = 0.0f;
  switch (lens_type) {
    case HB_LENS_LINEAR:
    case HB_LENS_GLOBE:
      out->scale = short_pix / 2.0f / std::tan(fov / 2.0f);
      break;
    case HB_LENS_FISHEYE_EQUAL_AREA:
      out->scale = short_pix / 2.0f / std::sqrt(2.0f) / std::sin(fov / 4.0f);
      break;
    case HB_LENS_FISHEYE_EQUIDISTANT:
      out->scale = short_pix * pi_2 / fov;
      break;
    case HB_LENS_FISHEYE_STEREOGRAPHIC:
      out->scale = short_pix / 2.0f / std::tan(fov / 4.0f);
      break;
    case HB_LENS_FISHEYE_ORTHOGRAPHIC:
      out->scale = short_pix / 2.0f / std::sin(fov / 2.0f);
      break;
    case HB_LENS_RECTANGULAR: {
      int short_res = std::min(img_w / 2, img_h);
      out->scale = static_cast<float>(short_res) / kPiF;
      float zx = m[2], zy = m[5];  // rot.Apply((0,0,1)) = third column
      out->az0 = std::atan2(zy, zx);
      break;
    }
    default:
      break;
  }
  if (overlap > 0) {  // projection.cpp:194-204
    if (lens_type == HB_LENS_DUAL_FISHEYE_EQUAL_AREA) {
      out->max_abs_dz = overlap;
      out->r_scale = 1.0f / std::sqrt(1.0f + overlap);
    } else if (lens_type == HB_LENS_DUAL_FISHEYE_EQUIDISTANT) {
      out->max_abs_dz = overlap;
      out->r_scale = pi_2 / (pi_2 + std::asin(overlap));
    } else if (lens_type == HB_LENS_DUAL_FISHEYE_STEREOGRAPHIC) {
      out->max_abs_dz = overlap;
      out->r_scale = 1.0f / std::tan((pi_2 + std::asin(overlap)) / 2.0f);
    }
  }
  return HB_OK;
}

int hb_build_render(const HbRenderDesc* d, HbProjParams* out) {
  if (d == nullptr) return HB_ERR_INVALID_ARG;
  return hb_make_proj_params(d->lens_type, d->fov_deg, d->img_w, d->img_h, d->view_az_deg, d->view_el_deg,
                             d->view_ro_deg, d->visible_range, d->lens_shift_x, d->lens_shift_y, d->overlap, out);
}

// PartitionCrystalRayNum, simulator.cpp:519-582: floor of the ideal share + per-population carry, then a
// largest-remainder correction so the counts sum to ray_num exactly.
int hb_partition_rays(const float* prop, uint32_t cnt, uint64_t ray_num, double* carry, uint64_t* out) {
  if (prop == nullptr || carry == nullptr || out == nullptr) return HB_ERR_INVALID_ARG;
  for (uint32_t i = 0; i < cnt; i++) out[i] = 0;
  if (cnt == 0 || ray_num == 0) return HB_OK;
  float total = 0.0f;
  for (uint32_t i = 0; i < cnt; i++) total += std::max(0.0f, prop[i]);
  if (total <= 0.0f) return HB_OK;
  uint64_t assigned = 0;
  for (uint32_t i = 0; i < cnt; i++) {
    double ideal = carry[i] + (static_cast<double>(std::max(0.0f, prop[i])) / total) * ray_num;
    uint64_t a = static_cast<uint64_t>(std::max(0.0, ideal));
    carry[i] = ideal - static_cast<double>(a);
    out[i] = a;
    assigned += a;
  }
  std::vector<uint32_t> idx(cnt);
  for (uint32_t i = 0; i < cnt; i++) idx[i] = i;
  if (assigned < ray_num) {
    uint64_t deficit = ray_num - assigned;
    std::partial_sort(idx.begin(), idx.begin() + std::min<uint64_t>(deficit, cnt), idx.end(),
                      [&](uint32_t a, uint32_t b) { return carry[a] > carry[b]; });
    for (uint64_t i = 0; i < deficit && i < cnt; i++) {
      out[idx[i]]++;
      carry[idx[i]] -= 1.0;
    }
  } else if (assigned > ray_num) {
    uint64_t surplus = assigned - ray_num;
    auto end = std::partition(idx.begin(), idx.end(), [&](uint32_t i) { return out[i] > 0; });
    uint64_t m = std::min<uint64_t>(surplus, static_cast<uint64_t>(end - idx.begin()));
    std::partial_sort(idx.begin(), idx.begin() + m, end, [&](uint32_t a, uint32_t b) { return carry[a] < carry[b]; });
    for (uint64_t i = 0; i < m; i++) {
      out[idx[i]]--;
      carry[idx[i]] += 1.0;
    }
  }
  return HB_OK;
}

int hb_build_scene(const HbSceneDesc* d, uint32_t geometry_seed, HbSceneTables** out) {
  if (d == nullptr || out == nullptr) return HB_ERR_INVALID_ARG;
  if (d->layer_cnt == 0 || d->layer_cnt > HB_MAX_LAYERS || d->max_hits == 0 || d->max_hits > HB_MAX_HITS) {
    global_error() = "scene: layer_cnt must be 1..8 and max_hits 1..64";
    return HB_ERR_INVALID_ARG;
  }
  auto t = std::make_unique<HbSceneTables>();
  hb::SceneStorage& st = t->st;
  std::mt19937 gen(geometry_seed);
  st.layers.resize(d->layer_cnt);
  st.pops.resize(d->layer_cnt);
  for (uint32_t li = 0; li < d->layer_cnt; li++) {
    const HbLayerDesc& ld = d->layers[li];
    if (ld.population_cnt == 0 || ld.population_cnt > HB_MAX_CRYSTALS) {
      global_error() = "scene: a layer needs 1..16 crystal populations";
      return HB_ERR_INVALID_ARG;
    }
    st.pops[li].resize(ld.population_cnt);
    for (uint32_t ci = 0; ci < ld.population_cnt; ci++) {
      const HbPopulationDesc& pd = ld.populations[ci];
      HbCrystalPopulation& p = st.pops[li][ci];
      std::memset(&p, 0, sizeof(p));
      p.proportion = pd.proportion;
      p.crystal_id = pd.crystal.id;
      uint32_t pool = ShapeIsDeterministic(pd.crystal) ? 1u : std::max(1u, d->geom_pool_size);
      auto shapes = std::make_unique<std::vector<HbCrystalTables>>(pool);
      for (uint32_t s = 0; s < pool; s++) {
        int rc = MakeShape(gen, pd.crystal, &(*shapes)[s]);
        if (rc != HB_OK) return rc;
      }
      p.shape_cnt = pool;
      p.shapes = shapes->data();
      st.shapes.push_back(std::move(shapes));
      int rc = hb_make_axis_sampler(pd.crystal.latitude.type, pd.crystal.latitude.center, pd.crystal.latitude.spread,
                                    pd.crystal.azimuth.type, pd.crystal.azimuth.center, pd.crystal.azimuth.spread,
                                    pd.crystal.roll.type, pd.crystal.roll.center, pd.crystal.roll.spread, &p.axis);
      if (rc != HB_OK) return rc;
      BuildFilter(pd, pd.filter, &p.filter);
      rc = BuildColorGroups(pd, &p);
      if (rc != HB_OK) {
        global_error() = "scene: more than 4 colour symmetry groups or 8 predicates per group in a population";
        return rc;
      }
    }
    st.layers[li].prob = ld.prob;
    st.layers[li].population_cnt = ld.population_cnt;
    st.layers[li].populations = st.pops[li].data();
  }
  st.scene.max_hits = d->max_hits;
  st.scene.layer_cnt = d->layer_cnt;
  st.scene.layers = st.layers.data();
  st.scene.sun_lon = (d->sun_azimuth_deg + 180.0f) * kDeg2RadF;  // cuda_trace_backend.cu:395-397
  st.scene.sun_lat = -d->sun_altitude_deg * kDeg2RadF;
  st.scene.sun_half_angle = (d->sun_diameter_deg * 0.5f) * kDeg2RadF;
  if (d->color_classes.class_cnt > HB_MAX_COLOR_CLASSES) {
    global_error() = "scene: more than 16 colour classes";
    return HB_ERR_UNSUPPORTED;
  }
  st.scene.color_classes = d->color_classes;
  *out = t.release();
  return HB_OK;
}

const HbScene* hb_scene_tables_get(const HbSceneTables* t) { return t ? &t->st.scene : nullptr; }
void hb_free_scene(HbSceneTables* t) { delete t; }

}  // extern "C"
