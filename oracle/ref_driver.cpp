// ref_driver.cpp — thin C driver around the UNMODIFIED reference CPU core (TEST INFRASTRUCTURE).
// See ref_driver.h. Everything here calls reference functions; no reference code is restated.
#include "ref_driver.h"

#include <atomic>
#include <chrono>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "config/proj_config.hpp"
#include "config/raypath_color_config.hpp"
#include "config/render_config.hpp"
#include "config/sim_data.hpp"
#include "core/backend/cpu_trace_backend.hpp"
#include "core/backend/wl_pool.hpp"
#include "core/crystal.hpp"
#include "core/device_filter_desc.hpp"
#include "core/filter_spec.hpp"
#include "core/lat_lut.hpp"
#include "core/lens_proj_build.hpp"
#include "core/optics.hpp"
#include "core/scatter_accum.hpp"
#include "core/shared/lat_path_selection.hpp"
#include "core/shared/pcg_shared.h"
#include "core/shared/projection_shared.h"
#include "core/simulator.hpp"
#include "core/trace_ops.hpp"
#include "core/color_util.hpp"
#include "util/color_data.hpp"
#include "util/color_space.hpp"
#include "util/cpu_info.hpp"
#include "util/illuminant.hpp"
#include "util/queue.hpp"

using namespace lumice;  // NOLINT

namespace {

Crystal MakeShape(const RefShape& s) {
  if (s.kind == 0) {
    return Crystal::CreatePrism(s.h1, s.dist);
  }
  return Crystal::CreatePyramid(s.upper_alpha_deg, s.lower_alpha_deg, s.h1, s.h2, s.h3, s.dist);
}

Distribution ToDist(const HbDist& d) {
  return Distribution{ static_cast<DistributionType>(d.type), d.center, d.spread };
}

AxisDistribution ToAxis(const HbDist& lat, const HbDist& az, const HbDist& roll) {
  AxisDistribution a;
  a.latitude_dist = ToDist(lat);
  a.azimuth_dist = ToDist(az);
  a.roll_dist = ToDist(roll);
  return a;
}

RenderConfig ToRender(const HbRenderDesc& r) {
  RenderConfig cfg;
  cfg.id_ = 1;
  cfg.lens_.type_ = static_cast<LensParam::LensType>(r.lens_type);
  cfg.lens_.fov_ = r.fov_deg;
  cfg.resolution_[0] = r.img_w;
  cfg.resolution_[1] = r.img_h;
  cfg.view_.az_ = r.view_az_deg;
  cfg.view_.el_ = r.view_el_deg;
  cfg.view_.ro_ = r.view_ro_deg;
  cfg.visible_ = static_cast<RenderConfig::VisibleRange>(r.visible_range);
  cfg.lens_shift_[0] = r.lens_shift_x;
  cfg.lens_shift_[1] = r.lens_shift_y;
  cfg.overlap_ = r.overlap;
  return cfg;
}

SimpleFilterParam ToSimple(const HbSimpleFilterSpec& f) {
  switch (f.kind) {
    case 1: {
      RaypathFilterParam p;
      for (uint32_t i = 0; i < f.path_len; i++) {
        p.raypath_.push_back(static_cast<IdType>(f.path[i]));
      }
      return SimpleFilterParam{ p };
    }
    case 2: {
      EntryExitFilterParam p;
      if (f.entry_fn >= 0) {
        p.entry_ = static_cast<IdType>(f.entry_fn);
      }
      if (f.exit_fn >= 0) {
        p.exit_ = static_cast<IdType>(f.exit_fn);
      }
      p.min_len_ = f.min_len == 0 ? 1 : f.min_len;
      if (f.max_len != 0) {
        p.max_len_ = f.max_len;
      }
      return SimpleFilterParam{ p };
    }
    case 3:
      return SimpleFilterParam{ DirectionFilterParam{ f.lon_deg, f.lat_deg, f.radii_deg } };
    case 4:
      return SimpleFilterParam{ CrystalFilterParam{ static_cast<IdType>(f.crystal_id) } };
    default:
      return SimpleFilterParam{ NoneFilterParam{} };
  }
}

FilterConfig ToFilter(const HbFilterSpecDesc& f) {
  FilterConfig c{};
  c.id_ = 1;
  c.symmetry_ = static_cast<uint8_t>(f.symmetry);
  c.action_ = f.action == 0 ? FilterConfig::kFilterIn : FilterConfig::kFilterOut;
  if (f.kind == 5) {
    ComplexFilterParam cp;
    for (uint32_t o = 0; o < f.term_cnt; o++) {
      std::vector<std::pair<IdType, SimpleFilterParam>> term;
      for (uint32_t a = 0; a < f.term_len[o]; a++) {
        term.emplace_back(static_cast<IdType>(o * 4 + a + 1), ToSimple(f.terms[o][a]));
      }
      cp.filters_.push_back(std::move(term));
    }
    c.param_ = cp;
  } else {
    c.param_ = ToSimple(f.simple);
  }
  return c;
}

CrystalConfig ToCrystalConfig(const HbCrystalDesc& c) {
  CrystalConfig cfg;
  cfg.id_ = static_cast<IdType>(c.id);
  if (c.kind == 0) {
    PrismCrystalParam p;
    p.h_ = ToDist(c.height[0]);
    for (int i = 0; i < 6; i++) {
      p.d_[i] = ToDist(c.face_dist[i]);
    }
    cfg.param_ = p;
  } else {
    PyramidCrystalParam p;
    p.h_pyr_u_ = ToDist(c.height[0]);
    p.h_prs_ = ToDist(c.height[1]);
    p.h_pyr_l_ = ToDist(c.height[2]);
    for (int i = 0; i < 6; i++) {
      p.d_[i] = ToDist(c.face_dist[i]);
    }
    p.wedge_angle_u_ = c.wedge_upper_deg;
    p.wedge_angle_l_ = c.wedge_lower_deg;
    cfg.param_ = p;
  }
  cfg.axis_ = ToAxis(c.latitude, c.azimuth, c.roll);
  return cfg;
}

SceneConfig ToScene(const HbSceneDesc& s, const std::vector<WlParam>& spectrum) {
  SceneConfig scene;
  scene.ray_num_ = 0;
  scene.max_hits_ = s.max_hits;
  scene.light_source_.param_ = SunParam{ s.sun_altitude_deg, s.sun_azimuth_deg, s.sun_diameter_deg };
  scene.light_source_.spectrum_ = spectrum;
  for (uint32_t li = 0; li < s.layer_cnt; li++) {
    MsInfo ms;
    ms.prob_ = s.layers[li].prob;
    for (uint32_t ci = 0; ci < s.layers[li].population_cnt; ci++) {
      const auto& pop = s.layers[li].populations[ci];
      ScatteringSetting st;
      st.crystal_ = ToCrystalConfig(pop.crystal);
      st.filter_ = ToFilter(pop.filter);
      st.crystal_proportion_ = pop.proportion;
      ms.setting_.push_back(std::move(st));
    }
    scene.ms_.push_back(std::move(ms));
  }
  return scene;
}

void FillTables(const Crystal& c, HbCrystalTables* out) {
  std::memset(out, 0, sizeof(*out));
  size_t fc = c.PolygonFaceCount();
  out->face_cnt = static_cast<uint32_t>(fc);
  const float* n = c.GetPolygonFaceNormal();
  const float* d = c.GetPolygonFaceDist();
  for (size_t i = 0; i < fc && i < HB_MAX_FACES; i++) {
    out->plane[i][0] = n[i * 3 + 0];
    out->plane[i][1] = n[i * 3 + 1];
    out->plane[i][2] = n[i * 3 + 2];
    out->plane[i][3] = d[i];
    out->face_fn[i] = static_cast<uint8_t>(c.GetFn(static_cast<IdType>(i)) & 0xFF);
  }
  size_t tc = detail::CountEntrySubTris(c.CfGeom());
  if (tc > HB_MAX_SUBTRIS) {
    tc = HB_MAX_SUBTRIS;
  }
  std::vector<detail::EntrySubTri> sub(detail::CountEntrySubTris(c.CfGeom()));
  if (!sub.empty()) {
    detail::BuildEntrySubTris(c.CfGeom(), sub.data());
  }
  out->subtri_cnt = static_cast<uint32_t>(tc);
  for (size_t t = 0; t < tc; t++) {
    std::memcpy(out->tri_v[t], sub[t].v, sizeof(float) * 9);
    std::memcpy(out->tri_n[t], sub[t].n, sizeof(float) * 3);
    out->tri_area[t] = sub[t].area;
    out->tri_face[t] = static_cast<uint8_t>(sub[t].face_id);
  }
}

}  // namespace

extern "C" {

int ref_make_tables(const RefShape* shape, HbCrystalTables* out) {
  Crystal c = MakeShape(*shape);
  FillTables(c, out);
  return 0;
}

int ref_make_axis_sampler(const HbDist* lat, const HbDist* az, const HbDist* roll, HbAxisSampler* out) {
  std::memset(out, 0, sizeof(*out));
  AxisDistribution axis = ToAxis(*lat, *az, *roll);
  auto decision = lat_path::SelectLatPath(axis);
  out->lat_path = lat_path::ToWireValue(decision.kind);
  out->lat_mean = axis.latitude_dist.center * math::kDegreeToRad;
  out->lat_std = axis.latitude_dist.spread * math::kDegreeToRad;
  out->az_type = static_cast<uint32_t>(axis.azimuth_dist.type);
  out->az_mean = axis.azimuth_dist.center * math::kDegreeToRad;
  out->az_std = axis.azimuth_dist.spread * math::kDegreeToRad;
  out->roll_type = static_cast<uint32_t>(axis.roll_dist.type);
  out->roll_mean = axis.roll_dist.center * math::kDegreeToRad;
  out->roll_std = axis.roll_dist.spread * math::kDegreeToRad;
  if (decision.kind == lat_path::LatPathKind::kLutInverseCdf) {
    const LatLut* lut = GetSharedLatLut(axis.latitude_dist);
    out->lut_n = LatLut::kNodes;
    std::memcpy(out->lut_theta, lut->theta.data(), sizeof(float) * LatLut::kNodes);
    std::memcpy(out->lut_cdf, lut->cdf.data(), sizeof(float) * LatLut::kNodes);
    std::memcpy(out->lut_flip, lut->flip_prob.data(), sizeof(float) * LatLut::kNodes);
  }
  return 0;
}

int ref_make_proj_params(const HbRenderDesc* r, HbProjParams* out) {
  RenderConfig cfg = ToRender(*r);
  Rotation rot = MakeCameraRotation(cfg);
  auto short_pix = static_cast<float>(std::min(cfg.resolution_[0], cfg.resolution_[1]));
  lm_proj::ProjParams p = BuildProjParams(cfg, rot, short_pix);
  static_assert(sizeof(lm_proj::ProjParams) == sizeof(HbProjParams), "ProjParams layout");
  std::memcpy(out, &p, sizeof(p));
  return 0;
}

int ref_wl_entry(float wl, float weight, HbWlEntry* out) {
  Crystal c = Crystal::CreatePrism(1.0f);
  std::vector<WlEntry> pool;
  ComputeWlPool(c, false, IlluminantType::kD65, wl, weight, 1, pool);
  std::memcpy(out, pool.data(), sizeof(WlEntry));
  return 0;
}

int ref_wl_pool_illuminant(int illuminant, uint32_t m, HbWlEntry* out) {
  Crystal c = Crystal::CreatePrism(1.0f);
  std::vector<WlEntry> pool;
  ComputeWlPool(c, true, static_cast<IlluminantType>(illuminant), 0.0f, 0.0f, m, pool);
  std::memcpy(out, pool.data(), sizeof(WlEntry) * m);
  return 0;
}

double ref_refractive_index(double wl) {
  return IceRefractiveIndex::Get(wl);
}

// RenderConsumer::PostSnapshot's per-pixel loop (server/render.cpp:537-577) driven with the reference's own
// colour functions (util/color_space.cpp). render.cpp itself pulls in the server layer and is not built here.
int ref_post_snapshot(const float* xyz_img, int w, int h, float snapshot_intensity, float intensity_factor,
                      const float* ray_color, const float* background, uint8_t* out) {
  int total_pix = w * h;
  if (total_pix <= 0 || snapshot_intensity <= 0) {
    std::memset(out, 0, static_cast<size_t>(std::max(total_pix, 0)) * 3);
    return 0;
  }
  float scale = intensity_factor * kNormScale * total_pix / snapshot_intensity;
  bool use_real_color = ray_color[0] < 0;
  for (int i = 0; i < total_pix; i++) {
    float xyz[3];
    for (int j = 0; j < 3; j++) {
      xyz[j] = xyz_img[i * 3 + j] * scale;
    }
    float rgb[3];
    if (use_real_color) {
      float clipped[3];
      GamutClipXyz(xyz, clipped);
      XyzToLinearRgb(clipped, rgb);
    } else {
      float gray[3];
      for (int j = 0; j < 3; j++) {
        gray[j] = kWhitePointD65[j] * xyz[1];
      }
      for (int j = 0; j < 3; j++) {
        float v = 0;
        for (int k = 0; k < 3; k++) {
          v += gray[k] * kXyzToRgb[j * 3 + k];
        }
        rgb[j] = v * ray_color[j];
      }
    }
    for (int j = 0; j < 3; j++) {
      rgb[j] += background[j];
      rgb[j] = std::clamp(rgb[j], 0.0f, 1.0f);
      rgb[j] = LinearToSrgb(rgb[j]);
      out[i * 3 + j] = static_cast<uint8_t>(rgb[j] * 255);
    }
  }
  return 0;
}

int ref_daylight_basis(float* s012) {
  for (int i = 0; i < kDaylightNumPoints; i++) {
    s012[i * 3 + 0] = kDaylightS0[i];
    s012[i * 3 + 1] = kDaylightS1[i];
    s012[i * 3 + 2] = kDaylightS2[i];
  }
  return kDaylightNumPoints;
}

int ref_cmf_table(float* xyz) {
  int n = kCmfMaxWavelength - kCmfMinWavelength + 1;
  for (int i = 0; i < n; i++) {
    xyz[i * 3 + 0] = kCmfX[i];
    xyz[i * 3 + 1] = kCmfY[i];
    xyz[i * 3 + 2] = kCmfZ[i];
  }
  return n;
}

static void FillSimpleFromDevice(const DeviceFilterDesc& d, HbSimpleFilter* s) {
  std::memset(s, 0, sizeof(*s));
  s->kind = d.type;
  s->path_len = d.canonical_len;
  std::memcpy(s->path, d.canonical_bytes, HB_MAX_FILTER_PATH);
  s->entry_fn = d.has_entry ? 1 : -1;  // presence flags only; canonical bytes carry the values
  s->exit_fn = d.has_exit ? 1 : -1;
  s->min_len = d.min_len;
  s->max_len = d.max_len;
  std::memcpy(s->dir, d.dir, sizeof(float) * 3);
  s->cos_radii = d.radii_c;
  s->crystal_id = d.crystal_id;
}

int ref_filter_desc(const HbPopulationDesc* pop, const RefShape* shape, HbFilterDesc* out) {
  std::memset(out, 0, sizeof(*out));
  Crystal c = MakeShape(*shape);
  FilterConfig fc = ToFilter(pop->filter);
  AxisDistribution axis = ToAxis(pop->crystal.latitude, pop->crystal.azimuth, pop->crystal.roll);
  DeviceFilterDesc d = detail::BuildDeviceFilterDesc(fc, c, axis);
  out->kind = d.type;
  out->action = d.action;
  out->symmetry = d.symmetry;
  out->fn_period = d.fn_period;
  out->sigma_a = d.sigma_a;
  out->d_applicable = d.d_applicable;
  if (d.type == kDeviceFilterTypeComplex) {
    std::vector<DeviceFilterDesc> subs;
    std::vector<uint8_t> counts;
    detail::BuildComplexSubDescs(std::get<ComplexFilterParam>(fc.param_), c, d.symmetry, d.sigma_a, d.d_applicable != 0,
                                 subs, counts);
    out->term_cnt = static_cast<uint32_t>(counts.size());
    size_t k = 0;
    for (size_t o = 0; o < counts.size() && o < HB_MAX_FILTER_TERMS; o++) {
      out->term_len[o] = counts[o];
      for (uint8_t a = 0; a < counts[o]; a++, k++) {
        if (a < 4) {
          FillSimpleFromDevice(subs[k], &out->terms[o][a]);
        }
      }
    }
  } else {
    FillSimpleFromDevice(d, &out->simple);
  }
  return 0;
}

int ref_hit_surface(const RefShape* shape, float n_idx, uint64_t n, const float* d3, const float* w,
                    const uint16_t* face, float* d_out6, float* w_out2) {
  Crystal c = MakeShape(*shape);
  std::vector<float> d(d3, d3 + n * 3);
  std::vector<float> ww(w, w + n);
  std::vector<IdType> f(face, face + n);
  HitSurface(c, n_idx, n, float_bf_t{ d.data(), 3 * sizeof(float) }, float_bf_t{ ww.data(), sizeof(float) },
             id_bf_t{ f.data(), sizeof(IdType) }, float_bf_t{ d_out6, 3 * sizeof(float) },
             float_bf_t{ w_out2, sizeof(float) });
  return 0;
}

int ref_propagate(const RefShape* shape, uint64_t n, const float* d3, const float* p3, const float* w,
                  const uint16_t* from_face, float* p_out3, uint16_t* to_face) {
  Crystal c = MakeShape(*shape);
  std::vector<float> d(d3, d3 + n * 3);
  std::vector<float> p(p3, p3 + n * 3);
  std::vector<float> ww(w, w + n);
  std::vector<IdType> f(from_face, from_face + n);
  std::vector<IdType> tf(n, kInvalidId);
  Propagate(c, n, 1, float_bf_t{ d.data(), 3 * sizeof(float) }, float_bf_t{ p.data(), 3 * sizeof(float) },
            float_bf_t{ ww.data(), sizeof(float) }, id_bf_t{ f.data(), sizeof(IdType) },
            float_bf_t{ p_out3, 3 * sizeof(float) }, id_bf_t{ tf.data(), sizeof(IdType) });
  for (uint64_t i = 0; i < n; i++) {
    to_face[i] = tf[i];
  }
  return 0;
}

int ref_project(const HbRenderDesc* r, uint64_t n, const float* dir3, int32_t* px2, int32_t* py2, int32_t* cnt,
                int32_t* bump2) {
  HbProjParams hp;
  ref_make_proj_params(r, &hp);
  lm_proj::ProjParams p;
  std::memcpy(&p, &hp, sizeof(p));
  for (uint64_t i = 0; i < n; i++) {
    auto res = lm_proj::ProjectExitToPixel(p, dir3[i * 3], dir3[i * 3 + 1], dir3[i * 3 + 2]);
    cnt[i] = res.count;
    for (int k = 0; k < 2; k++) {
      px2[i * 2 + k] = k < res.count ? res.hits[k].px : -1;
      py2[i * 2 + k] = k < res.count ? res.hits[k].py : -1;
      bump2[i * 2 + k] = k < res.count ? (res.hits[k].bump_landed ? 1 : 0) : 0;
    }
  }
  return 0;
}

int ref_scatter_xyz(const HbRenderDesc* r, float wl, uint64_t n, const float* dir3, const float* w, float* xyz_wh3,
                    float* landed) {
  RenderConfig cfg = ToRender(*r);
  Rotation rot = MakeCameraRotation(cfg);
  ScatterOutgoingToXyz(dir3, w, n, cfg, rot, wl, xyz_wh3, landed);
  return 0;
}

int ref_filter_check(const HbPopulationDesc* pop, const RefShape* shape, uint64_t n, const uint8_t* paths64,
                     const uint8_t* path_len, const float* dir3, uint8_t* pass) {
  Crystal c = MakeShape(*shape);
  c.config_id_ = static_cast<IdType>(pop->crystal.id);
  FilterConfig fc = ToFilter(pop->filter);
  AxisDistribution axis = ToAxis(pop->crystal.latitude, pop->crystal.azimuth, pop->crystal.roll);
  auto spec = FilterSpec::Create(fc, c, axis);
  RayBuffer buf(2);
  for (uint64_t i = 0; i < n; i++) {
    buf.size_ = 1;
    RaySeg& r = buf[0];
    r = RaySeg{};
    r.d_[0] = dir3[i * 3];
    r.d_[1] = dir3[i * 3 + 1];
    r.d_[2] = dir3[i * 3 + 2];
    r.w_ = 1.0f;
    r.to_face_ = kInvalidId;
    r.crystal_config_id_ = c.config_id_;
    r.crystal_idx_ = 0;
    buf.RecorderClear(0);
    for (uint32_t k = 0; k < path_len[i]; k++) {
      buf.RecorderAppend(0, static_cast<IdType>(paths64[i * 64 + k]));  // face numbers
    }
    pass[i] = (spec == nullptr || spec->Check(r, buf.RecorderAt(0), buf.OverflowArena())) ? 1 : 0;
  }
  return 0;
}

int ref_sample_orientations(const HbDist* lat, const HbDist* az, const HbDist* roll, uint32_t seed, uint64_t n,
                            float* lon_lat_roll3, float* rot9) {
  AxisDistribution axis = ToAxis(*lat, *az, *roll);
  RandomNumberGenerator rng(seed);
  RandomNumberGenerator::GetInstance().SetSeed(seed);
  const bool full_sphere = axis.IsFullSphereUniform();
  const LatLut* lut = nullptr;
  if (!full_sphere && lat_path::SelectLatPath(axis).kind == lat_path::LatPathKind::kLutInverseCdf) {
    lut = GetSharedLatLut(axis.latitude_dist);
  }
  for (uint64_t i = 0; i < n; i++) {
    float llr[3]{};
    if (!full_sphere) {
      RandomSampler::SampleSphericalPointsSph(axis, llr, 1, lut);
    } else {
      RandomSampler::SampleSphericalPointsSph(llr);
      llr[2] = rng.Get(axis.roll_dist) * math::kDegreeToRad;
    }
    std::memcpy(lon_lat_roll3 + i * 3, llr, sizeof(llr));
    if (rot9 != nullptr) {
      Rotation r = BuildCrystalRotation(llr[0], llr[1], llr[2]);
      std::memcpy(rot9 + i * 9, r.GetMat(), sizeof(float) * 9);
    }
  }
  return 0;
}

int ref_partition(const float* proportions, uint32_t cnt, uint64_t ray_num, double* carry, uint64_t* out) {
  std::vector<float> p(proportions, proportions + cnt);
  std::vector<double> c(carry, carry + cnt);
  auto res = PartitionCrystalRayNum(p, ray_num, c);
  for (uint32_t i = 0; i < cnt; i++) {
    out[i] = res[i];
    carry[i] = c[i];
  }
  return 0;
}

static int TraceInjected(const RefShape* shape, float n_idx, uint32_t max_hits, uint64_t n, const float* d3,
                         const float* p3, const float* w, const uint16_t* to_face, const HbPopulationDesc* color_pop,
                         uint64_t cap, HbExitRecord* out, uint32_t* out_ray, uint64_t* count) {
  static_assert(sizeof(HbExitRecord) == sizeof(ExitRayRecord), "exit record layout");
  Crystal crystal = MakeShape(*shape);
  crystal.config_id_ = 0;

  HbSceneDesc sd{};
  sd.max_hits = max_hits;
  sd.layer_cnt = 1;
  sd.sun_altitude_deg = 20.0f;
  sd.sun_diameter_deg = 0.5f;
  sd.layers[0].prob = 0.0f;
  sd.layers[0].population_cnt = 1;
  auto& pop = sd.layers[0].populations[0];
  pop.proportion = 1.0f;
  pop.crystal.kind = 0;
  pop.crystal.height[0] = HbDist{ 0, 1.0f, 0.0f };
  for (auto& d : pop.crystal.face_dist) {
    d = HbDist{ 0, 1.0f, 0.0f };
  }
  pop.crystal.latitude = HbDist{ 0, 90.0f, 0.0f };
  // Raypath colour: one class per predicate => BuildColorGateTable assigns bit k to predicate k
  // (insertion order within the single (layer 0, crystal id) placement, color_gate_table.hpp:25-33).
  RaypathColorConfig color_cfg;
  if (color_pop != nullptr) {
    pop.crystal.id = color_pop->crystal.id;
    crystal.config_id_ = static_cast<IdType>(color_pop->crystal.id);  // what CrystalSpec::Check compares
    pop.crystal.azimuth = color_pop->crystal.azimuth;   // decide D-symmetry applicability only (injected rays)
    pop.crystal.roll = color_pop->crystal.roll;
    for (uint32_t k = 0; k < color_pop->color_pred_cnt; k++) {
      const HbColorPredDesc& cp = color_pop->color_preds[k];
      ColorClassConfig cls;
      RaypathColorRef ref;
      ref.layer_ = 0;
      ref.crystal_ = static_cast<IdType>(color_pop->crystal.id);
      ref.predicate_ = ToSimple(cp.pred);
      ref.symmetry_ = static_cast<uint8_t>(cp.symmetry);
      cls.match_.push_back(ref);
      color_cfg.classes_.push_back(cls);
    }
  }
  std::vector<WlParam> spectrum{ { 550.0f, 1.0f } };
  SceneConfig scene = ToScene(sd, spectrum);

  RenderConfig render;
  render.lens_.type_ = LensParam::kRectangular;
  render.lens_.fov_ = 360.0f;
  render.resolution_[0] = 16;
  render.resolution_[1] = 8;
  render.view_.el_ = 90.0f;
  render.visible_ = RenderConfig::kFull;

  CpuTraceBackend backend;
  uint64_t k = 0;
  std::vector<ExitRayRecord> recs;
  for (uint64_t i = 0; i < n; i++) {
    SessionSpec spec{};
    spec.scene = &scene;
    spec.render = &render;
    spec.wl = WlParam{ 550.0f, 1.0f };
    spec.seed = 1;
    if (color_pop != nullptr) {
      spec.raypath_color = std::make_shared<const RaypathColorConfig>(color_cfg);
    }
    backend.BeginSession(spec);
    // The session carries the ray under test plus kPad zero-weight copies. The copies only size the
    // reference's per-session workspace (workspace[0] holds 2 x session rays, cpu_trace_backend.cpp:118-119) the
    // way a production 32-ray small batch does, so a near-edge double continuation is not dropped for lack of
    // room; their exits have weight exactly 0 and are discarded below.
    constexpr size_t kPad = 7;
    float bd[(kPad + 1) * 3], bp[(kPad + 1) * 3], bw[kPad + 1];
    IdType btf[kPad + 1];
    for (size_t q = 0; q <= kPad; q++) {
      std::memcpy(bd + q * 3, d3 + i * 3, 12);
      std::memcpy(bp + q * 3, p3 + i * 3, 12);
      bw[q] = q == 0 ? w[i] : 0.0f;
      btf[q] = to_face[i];
    }
    HostRayBatch hb;
    hb.count = kPad + 1;
    hb.d = bd;
    hb.p = bp;
    hb.w = bw;
    hb.tf = btf;
    hb.crystal = &crystal;
    hb.refractive_index = n_idx;
    hb.crystal_id = color_pop != nullptr ? static_cast<IdType>(color_pop->crystal.id) : 0;
    auto handle = backend.TraceLayer(RootRaySource::FromHost(hb));
    backend.DrainExits(recs);
    backend.EndSession();
    for (const auto& r : recs) {
      if (r.weight == 0.0f && w[i] != 0.0f) {
        continue;  // padding ray
      }
      if (k < cap) {
        std::memcpy(&out[k], &r, sizeof(r));
        out_ray[k] = static_cast<uint32_t>(i);
      }
      k++;
    }
  }
  *count = k;
  return k > cap ? -5 : 0;
}

int ref_trace_injected(const RefShape* shape, float n_idx, uint32_t max_hits, uint64_t n, const float* d3,
                       const float* p3, const float* w, const uint16_t* to_face, uint64_t cap, HbExitRecord* out,
                       uint32_t* out_ray, uint64_t* count) {
  return TraceInjected(shape, n_idx, max_hits, n, d3, p3, w, to_face, nullptr, cap, out, out_ray, count);
}

int ref_trace_injected_color(const RefShape* shape, float n_idx, uint32_t max_hits, uint64_t n, const float* d3,
                             const float* p3, const float* w, const uint16_t* to_face,
                             const HbPopulationDesc* color_pop, uint64_t cap, HbExitRecord* out, uint32_t* out_ray,
                             uint64_t* count) {
  return TraceInjected(shape, n_idx, max_hits, n, d3, p3, w, to_face, color_pop, cap, out, out_ray, count);
}

int ref_cpu_backend_run(const HbSceneDesc* sd, const HbRenderDesc* rd, float wl, float weight, uint32_t seed,
                        uint64_t n_rays, uint64_t session_rays, float* xyz_wh3, float* landed, uint64_t* exit_count,
                        double* exit_w_sum) {
  std::vector<WlParam> spectrum{ { wl, weight } };
  SceneConfig scene = ToScene(*sd, spectrum);
  RenderConfig render = ToRender(*rd);
  size_t pix = static_cast<size_t>(render.resolution_[0]) * render.resolution_[1];
  std::vector<float> img(pix * 3);
  Rotation cam = MakeCameraRotation(render);
  CpuTraceBackend backend;
  uint64_t exits = 0;
  double wsum = 0.0;
  float landed_total = 0.0f;
  std::vector<ExitRayRecord> recs;
  for (uint64_t done = 0; done < n_rays; done += session_rays) {
    uint64_t cnt = std::min<uint64_t>(session_rays, n_rays - done);
    SessionSpec spec{};
    spec.scene = &scene;
    spec.render = &render;
    spec.wl = WlParam{ wl, weight };
    spec.seed = seed;
    backend.BeginSession(spec);
    HostRayBatch hb;
    hb.count = cnt;
    RootRaySource roots = RootRaySource::FromHost(hb);
    for (size_t mi = 0; mi < scene.ms_.size(); mi++) {
      auto handle = backend.TraceLayer(roots);
      backend.DrainExits(recs);
      // Same consumer the reference's exit-seam path feeds: ScatterOutgoingToXyz on (dir, weight).
      std::vector<float> d(recs.size() * 3);
      std::vector<float> ww(recs.size());
      for (size_t i = 0; i < recs.size(); i++) {
        d[i * 3] = recs[i].dir[0];
        d[i * 3 + 1] = recs[i].dir[1];
        d[i * 3 + 2] = recs[i].dir[2];
        ww[i] = recs[i].weight;
        wsum += recs[i].weight;
      }
      exits += recs.size();
      ScatterOutgoingToXyz(d.data(), ww.data(), ww.size(), render, cam, wl, xyz_wh3, &landed_total);
      if (mi + 1 == scene.ms_.size()) {
        break;
      }
      roots = backend.Recombine(std::move(handle), RecombineSpec{ true });
    }
    backend.EndSession();
  }
  *landed = landed_total;
  *exit_count = exits;
  *exit_w_sum = wsum;
  return 0;
}

uint32_t ref_physical_cores(void) {
  return static_cast<uint32_t>(PhysicalCoreCount());
}

int ref_legacy_bench(const HbSceneDesc* sd, const HbRenderDesc* rd, const float* wl, const float* wl_weight,
                     uint32_t wl_cnt, uint64_t rays_per_wl, uint32_t threads, uint32_t dispatch_rays,
                     double* rays_per_sec, double* seconds, uint64_t* exits_out) {
  std::vector<WlParam> spectrum;
  for (uint32_t i = 0; i < wl_cnt; i++) {
    spectrum.push_back(WlParam{ wl[i], wl_weight[i] });
  }
  auto scene = std::make_shared<const SceneConfig>(ToScene(*sd, spectrum));
  RenderConfig render = ToRender(*rd);
  Rotation cam = MakeCameraRotation(render);
  size_t pix = static_cast<size_t>(render.resolution_[0]) * render.resolution_[1];
  std::vector<float> img(pix * 3, 0.0f);

  auto scene_q = std::make_shared<Queue<SimBatch>>();
  auto data_q = std::make_shared<Queue<SimData>>();
  scene_q->Start();
  data_q->Start();

  // Producer: the same 128-ray (dispatch_rays) SimBatches ServerImpl::GenerateScene enqueues
  // (server.cpp:1493-1521); each SimBatch runs every wavelength of the discrete spectrum.
  uint64_t n_batches = 0;
  for (uint64_t committed = 0; committed < rays_per_wl; committed += dispatch_rays) {
    uint64_t b = std::min<uint64_t>(dispatch_rays, rays_per_wl - committed);
    SimBatch sb;
    sb.ray_num_ = b;
    sb.scene_ = scene;
    scene_q->Emplace(std::move(sb));
    n_batches++;
  }
  for (uint32_t t = 0; t < threads; t++) {
    scene_q->Emplace(SimBatch{});  // termination signal, one per worker
  }

  std::vector<std::unique_ptr<Simulator>> sims;
  for (uint32_t t = 0; t < threads; t++) {
    sims.push_back(std::make_unique<Simulator>(scene_q, data_q, 0));
  }

  std::atomic<uint64_t> roots{ 0 };
  std::atomic<uint64_t> exits{ 0 };
  std::atomic<bool> done{ false };
  float landed = 0.0f;
  // Consumer: host projection + XYZ accumulate with the reference's own ScatterOutgoingToXyz
  // (the arithmetic RenderConsumer::Consume runs on outgoing_d_/w_, render.cpp:204-...).
  std::thread consumer([&]() {
    while (true) {
      SimData data = data_q->Get();
      if (data.outgoing_w_.empty() && data.root_ray_count_ == 0) {
        if (done.load()) {
          break;
        }
        continue;
      }
      ScatterOutgoingToXyz(data.outgoing_d_.data(), data.outgoing_w_.data(), data.outgoing_w_.size(), render, cam,
                           data.curr_wl_, img.data(), &landed);
      roots += data.root_ray_count_;
      exits += data.outgoing_w_.size();
    }
  });

  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> workers;
  for (uint32_t t = 0; t < threads; t++) {
    workers.emplace_back([&, t]() { sims[t]->Run(); });
  }
  for (auto& th : workers) {
    th.join();
  }
  // Drain: wait until the consumer has folded every SimData.
  uint64_t expect = rays_per_wl * wl_cnt;
  while (roots.load() < expect) {
    std::this_thread::sleep_for(std::chrono::milliseconds(1));
    auto el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (el > 3600.0) {
      break;
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  done = true;
  data_q->Shutdown();
  consumer.join();
  double sec = std::chrono::duration<double>(t1 - t0).count();
  *seconds = sec;
  *rays_per_sec = static_cast<double>(roots.load()) / sec;
  *exits_out = exits.load();
  return 0;
}

// --- the reference's counter-based sampler (core/shared/pcg_shared.h, host-compilable: the code its GPU backends run) ---
// Every function below CALLS lm_pcg::*; the oracle's twin generator is pinned against them bit-for-bit on integers
// and to libm-equal floats (tests/test_oracle_vs_reference.py::test_sampler_twin_matches_reference_pcg).
uint32_t ref_pcg_hash(uint32_t x) { return lm_pcg::pcg_hash(x); }
uint32_t ref_pcg_seed_with_high(uint32_t seed, uint32_t hi) { return lm_pcg::pcg_seed_with_high(seed, hi); }
void ref_pcg_uniforms(uint32_t seed, uint32_t idx, uint32_t slot0, uint32_t n, float* out) {
  lm_pcg::PcgStream s{ seed, idx, slot0 };
  for (uint32_t i = 0; i < n; i++) out[i] = lm_pcg::pcg_uniform(s);
}
void ref_pcg_get_dist(uint32_t seed, uint32_t idx0, uint32_t n, uint32_t type, float mean, float stdv, float* out) {
  for (uint32_t i = 0; i < n; i++) {
    lm_pcg::PcgStream s{ seed, idx0 + i, 0u };
    out[i] = lm_pcg::pcg_get_dist(s, type, mean, stdv);
  }
}
// sample_lat_lon_roll with the wire struct filled from an HbAxisSampler (same fields the backends upload)
void ref_pcg_lat_lon_roll(const HbAxisSampler* a, uint32_t seed, uint32_t idx0, uint32_t n, float* lon_lat_roll3,
                          uint32_t* slots_used) {
  lm_pcg::GenRootKernelParams gp{};
  gp.lat_path = a->lat_path;
  gp.lat_mean_rad = a->lat_mean;
  gp.lat_std_rad = a->lat_std;
  gp.lat_rejection_m = 1.0f;
  gp.lat_lut_n = a->lut_n;
  gp.az_type = a->az_type;
  gp.az_mean_rad = a->az_mean;
  gp.az_std_rad = a->az_std;
  gp.roll_type = a->roll_type;
  gp.roll_mean_rad = a->roll_mean;
  gp.roll_std_rad = a->roll_std;
  for (uint32_t i = 0; i < n; i++) {
    lm_pcg::PcgStream s{ seed, idx0 + i, 0u };
    float lon, lat, roll;
    lm_pcg::sample_lat_lon_roll(s, gp, a->lut_theta, a->lut_cdf, a->lut_flip, lon, lat, roll, nullptr);
    lon_lat_roll3[i * 3 + 0] = lon;
    lon_lat_roll3[i * 3 + 1] = lat;
    lon_lat_roll3[i * 3 + 2] = roll;
    if (slots_used != nullptr) slots_used[i] = s.slot;
  }
}
void ref_pcg_rotation9(uint64_t n, const float* lon_lat_roll3, float* rot9) {
  for (uint64_t i = 0; i < n; i++) {
    lm_pcg::build_crystal_rotation_9(lon_lat_roll3[i * 3], lon_lat_roll3[i * 3 + 1], lon_lat_roll3[i * 3 + 2], rot9 + i * 9);
  }
}
void ref_pcg_sph_cap(uint32_t seed, uint32_t idx0, uint32_t slot0, uint32_t n, float lon, float lat, float half, float* d3) {
  for (uint32_t i = 0; i < n; i++) {
    lm_pcg::PcgStream s{ seed, idx0 + i, slot0 };
    lm_pcg::sample_sph_cap(s, lon, lat, half, d3 + i * 3);
  }
}
void ref_pcg_triangle(uint32_t seed, uint32_t idx0, uint32_t slot0, uint32_t n, const float* vtx9, float* p3) {
  for (uint32_t i = 0; i < n; i++) {
    lm_pcg::PcgStream s{ seed, idx0 + i, slot0 };
    lm_pcg::sample_triangle(s, vtx9, p3 + i * 3);
  }
}
void ref_pcg_feistel(uint32_t n, uint32_t seed, uint32_t* out) {
  for (uint32_t i = 0; i < n; i++) out[i] = lm_pcg::feistel_bijection(i, n, seed);
}
void ref_pcg_categorical(const float* weights, uint32_t n, const float* u, uint32_t m, uint32_t* out) {
  for (uint32_t i = 0; i < m; i++) out[i] = lm_pcg::categorical_sample(weights, n, u[i]);
}

// Simulator::Run on its TraceBackend route (SimulateOneWavelengthWithBackend + third-clock drain), unmodified.
// Which backend CreateBackend() hands out is decided by how this library was BUILT (oracle/Makefile):
//   libhalo_ref*.so      no LUMICE_CUDA_ENABLED: kCuda falls back to the legacy CPU path (backend_used = 0)
//   libhalo_refcuda.so   the reference's own CudaTraceBackend (cuda_trace_backend.cu compiled for sm_100a)
//   libhalo_refb200.so   oracle/shim: this repo's B200TraceBackend behind the same name
// One Simulator thread, as the reference's GPU route runs (server.cpp:451-454); `dispatch_rays` per SimBatch
// (kDefaultCudaDispatchRayNum = 262144 is the reference's CUDA default). The consumer sums the drained device images.
int ref_backend_bench(const HbSceneDesc* sd, const HbRenderDesc* rd, const float* wl, const float* wl_weight,
                      uint32_t wl_cnt, uint64_t rays_per_wl, uint64_t dispatch_rays, uint32_t seed, float* xyz_wh3,
                      double* landed_out, double* rays_per_sec, double* seconds, uint32_t* backend_used) {
  std::vector<WlParam> spectrum;
  for (uint32_t i = 0; i < wl_cnt; i++) {
    spectrum.push_back(WlParam{ wl[i], wl_weight[i] });
  }
  auto scene = std::make_shared<const SceneConfig>(ToScene(*sd, spectrum));
  auto renders = std::make_shared<const std::vector<RenderConfig>>(std::vector<RenderConfig>{ ToRender(*rd) });
  const RenderConfig& render = (*renders)[0];
  Rotation cam = MakeCameraRotation(render);
  const size_t pix = static_cast<size_t>(render.resolution_[0]) * render.resolution_[1];
  std::vector<double> img(pix * 3, 0.0);
  std::vector<float> host_img(pix * 3, 0.0f);

  auto scene_q = std::make_shared<Queue<SimBatch>>();
  auto data_q = std::make_shared<Queue<SimData>>();
  scene_q->Start();
  data_q->Start();
  for (uint64_t committed = 0; committed < rays_per_wl; committed += dispatch_rays) {
    SimBatch sb;
    sb.ray_num_ = std::min<uint64_t>(dispatch_rays, rays_per_wl - committed);
    sb.scene_ = scene;
    sb.renders_ = renders;
    sb.generation_ = 1;
    scene_q->Emplace(std::move(sb));
  }
  scene_q->Emplace(SimBatch{});  // termination signal

  Simulator sim(scene_q, data_q, seed);
  sim.SetPreferredBackend(BackendKind::kCuda);

  std::atomic<uint64_t> roots{ 0 };
  std::atomic<uint64_t> fused{ 0 };
  std::atomic<bool> done{ false };
  double landed = 0.0;
  float host_landed = 0.0f;
  std::thread consumer([&]() {
    while (true) {
      SimData data = data_q->Get();
      if (data.root_ray_count_ == 0 && data.outgoing_w_.empty() && data.xyz_pixel_data_.empty()) {
        if (done.load()) {
          break;
        }
        continue;
      }
      if (!data.xyz_pixel_data_.empty()) {  // device-fused window (ConsumeDeviceFused's payload)
        for (size_t i = 0; i < data.xyz_pixel_data_.size() && i < img.size(); i++) {
          img[i] += data.xyz_pixel_data_[i];
        }
        landed += data.xyz_landed_weight_;
        fused++;
      } else if (!data.outgoing_w_.empty()) {  // legacy / exit-seam payload: host projection
        ScatterOutgoingToXyz(data.outgoing_d_.data(), data.outgoing_w_.data(), data.outgoing_w_.size(), render, cam,
                             data.curr_wl_, host_img.data(), &host_landed);
      }
      roots += data.root_ray_count_;
    }
  });

  auto t0 = std::chrono::steady_clock::now();
  sim.Run();
  const uint64_t expect = rays_per_wl * wl_cnt;
  while (roots.load() < expect) {
    std::this_thread::sleep_for(std::chrono::microseconds(200));
    if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 3600.0) {
      break;
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  done = true;
  data_q->Shutdown();
  consumer.join();
  const double sec = std::chrono::duration<double>(t1 - t0).count();
  if (xyz_wh3 != nullptr) {
    for (size_t i = 0; i < img.size(); i++) {
      xyz_wh3[i] = static_cast<float>(img[i] + host_img[i]);
    }
  }
  if (landed_out != nullptr) *landed_out = landed + host_landed;
  if (seconds != nullptr) *seconds = sec;
  if (rays_per_sec != nullptr) *rays_per_sec = static_cast<double>(roots.load()) / sec;
  if (backend_used != nullptr) *backend_used = fused.load() != 0 ? 1u : 0u;
  return 0;
}

}  // extern "C"
