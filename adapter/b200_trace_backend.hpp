// b200_trace_backend.hpp — the reference-side binding: a `lumice::TraceBackend` that forwards every
// virtual to the C ABI of libhalotrace_b200.so (include/halotrace_b200.h).
//
// This file is compiled INSIDE the reference tree (it includes the reference's own headers and uses its
// host geometry code to build the tables, as the reference's CUDA backend does,
// src/core/backend/cuda_trace_backend.cu:2436-2542,3689-3704). It is what a Lumice maintainer adds next
// to cuda_trace_backend.{hpp,cu}; INTEGRATION.md lists the five small touch points that register it.
// `make -C oracle adapter_check` compiles it against /root/reference as a syntax/ABI check.
#ifndef ADAPTER_B200_TRACE_BACKEND_HPP_
#define ADAPTER_B200_TRACE_BACKEND_HPP_

#include <algorithm>
#include <cstring>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "config/color_class_table.hpp"
#include "config/color_gate_table.hpp"
#include "config/raypath_color_config.hpp"
#include "core/backend/trace_backend.hpp"
#include "core/backend/wl_pool.hpp"
#include "core/device_filter_desc.hpp"
#include "core/lat_lut.hpp"
#include "core/lens_proj_build.hpp"
#include "core/scatter_accum.hpp"
#include "core/shared/lat_path_selection.hpp"
#include "core/simulator.hpp"
#include "core/trace_ops.hpp"
#include "halotrace_b200.h"

namespace lumice {

class B200LayerHandle : public LayerHandle {
 public:
  size_t ContinuationCount() const override { return static_cast<size_t>(stats_.continuation_count); }
  LayerStats GetLayerStats() const override {
    return LayerStats{ static_cast<size_t>(stats_.exit_count), static_cast<float>(stats_.exit_w_sum) };
  }
  HbLayerStats stats_{};
};

class B200TraceBackend : public TraceBackend {
 public:
  explicit B200TraceBackend(int device_ordinal = 0) : B200TraceBackend(std::vector<int>{ device_ordinal }) {}
  // Several devices behind ONE backend instance: the reference's GPU route runs a single Simulator
  // thread (server.cpp:451-454), so scaling over the GPUs of a node has to happen behind the seam. Every session is
  // split into contiguous shares of the global ray-index range, one per device; the engines' calls are asynchronous,
  // so this one host thread keeps all devices busy; at drain time device 0 adds the other devices' accumulators out
  // of their memory (hb_merge_from_peer: P2P loads over NVLink) and is the only one read back.
  explicit B200TraceBackend(const std::vector<int>& devices) : rng_(0) {
    if (devices.empty()) {
      throw BackendUnavailableError("B200TraceBackend: empty device list");
    }
    for (int d : devices) {
      HbEngine* h = nullptr;
      if (hb_create(d, &h) != HB_OK) {
        for (HbEngine* e : hs_) hb_destroy(e);
        throw BackendUnavailableError(std::string("B200TraceBackend: ") + hb_last_error(nullptr));
      }
      hs_.push_back(h);
    }
    h_ = hs_[0];
    share_.assign(hs_.size(), 0);
  }
  ~B200TraceBackend() override {
    for (HbEngine* e : hs_) hb_destroy(e);
  }
  size_t DeviceCount() const { return hs_.size(); }

  // Exit-seam egress (trace_backend.hpp:391-446). Default: the device-fused consumer (projection + XYZ accumulate
  // on the GPU, DrainExits returns nothing). With SetExitEgress(true) the backend behaves like the CPU backend at
  // the seam instead: no device image, every outgoing ray of every layer is materialised as a 96-byte
  // ExitRayRecord {dir, weight, path (face numbers), crystal_id, ms_layer_idx, wl_idx, component_mask} and handed
  // to the driver by DrainExits (destructive, grow-not-clamp) for its host consumers (show_rays, raypath
  // statistics, GUI filters). Switch between sessions only.
  void SetExitEgress(bool on) { egress_ = on; }
  bool SupportsDeviceXyzAccum() const override { return !egress_; }
  bool SupportsThirdClockDrain() const override { return true; }
  uint32_t WlPoolSize() const override { return kWlPoolSizeDefault; }
  bool IsCompatible(const RenderConfig&) const override { return true; }  // all 11 lens types

  void BeginSession(const SessionSpec& spec) override {
    ThrowDeferred();
    try {
      raypath_color_ = spec.raypath_color;
      if (spec.scene != scene_ || raypath_color_ != color_uploaded_) {
        UploadScene(*spec.scene);
        // Stochastic geometry: hand the populations' shape distributions to the engine's geometry clock. Every
        // session then gets a fresh pool of kPoolShapes crystals per population, drawn and built on the device one
        // session ahead (hb_auto_resample) -- the host no longer calls MakeCrystal per session.
        StartGeometryClock(*spec.scene, spec.seed);
      }
      stochastic_shapes_last_upload_ = stochastic_population_cnt_ * kPoolShapes;
      if (spec.render != render_ || render_snapshot_dirty_) {
        UploadRender(*spec.render);
      }
      // Wavelength pool: one entry for a discrete wavelength, M midpoint samples for an illuminant
      // (ComputeWlPool, backend/wl_pool.hpp:67-95).
      const bool illuminant = std::holds_alternative<IlluminantType>(spec.scene->light_source_.spectrum_);
      std::vector<WlEntry> pool;
      Crystal probe = Crystal::CreatePrism(1.0f);
      ComputeWlPool(probe, illuminant,
                    illuminant ? std::get<IlluminantType>(spec.scene->light_source_.spectrum_) : IlluminantType::kD65,
                    spec.wl.wl_, spec.wl.weight_, illuminant ? kWlPoolSizeDefault : 1u, pool);
      static_assert(sizeof(WlEntry) == sizeof(HbWlEntry), "WlEntry layout");
      HbSessionSpec s{};
      s.seed = spec.seed;
      s.wl_cnt = static_cast<uint32_t>(pool.size());
      s.wl = reinterpret_cast<const HbWlEntry*>(pool.data());
      s.record_exits = egress_ ? 2u : 0u;
      s.accumulate = egress_ ? 0u : 1u;
      // contiguous shares of the session's global ray-index range [ray_base_, ray_base_ + ray_num)
      const size_t R = hs_.size();
      size_t first = 0;
      for (size_t r = 0; r < R; r++) {
        share_[r] = spec.ray_num / R + (r < spec.ray_num % R ? 1 : 0);
        s.ray_num = share_[r];
        if (R > 1) {  // every stream of the share is keyed by its global index (hb_begin_session)
          s.use_ray_base = 1;
          s.ray_base = ray_base_ + first;
        }
        Check(hb_begin_session(hs_[r], &s), "BeginSession", hs_[r]);
        first += share_[r];
      }
      ray_base_ += spec.ray_num;
      layer_cnt_ = spec.scene->ms_.size();
      layer_idx_ = 0;
      layer_roots_ = spec.ray_num;
      orientation_draws_ = 0;
    } catch (...) {
      for (HbEngine* e : hs_) hb_end_session(e);  // leave the instance un-sessioned (trace_backend.hpp:149-153)
      throw;
    }
  }

  LayerHandlePtr TraceLayer(const RootRaySource& roots) override {
    ThrowDeferred();
    auto handle = std::make_unique<B200LayerHandle>();
    const bool last = layer_idx_ + 1 == layer_cnt_;
    if (layer_idx_ < layer_axis_stochastic_.size() && layer_axis_stochastic_[layer_idx_]) {
      orientation_draws_ += layer_roots_;
    }
    // The final layer needs no host-visible counter: launch and return (no synchronisation), so the devices of a
    // multi-device backend all run at once; a gated layer reads each device's continuation count in turn.
    for (size_t r = 0; r < hs_.size(); r++) {
      HbLayerStats st{};
      Check(hb_trace_layer(hs_[r], roots.is_device ? 0 : share_[r], last ? nullptr : &st), "TraceLayer", hs_[r]);
      handle->stats_.root_count += st.root_count;
      handle->stats_.continuation_count += st.continuation_count;
      handle->stats_.exit_count += st.exit_count;
      handle->stats_.exit_w_sum += st.exit_w_sum;
    }
    return handle;
  }

  RootRaySource Recombine(LayerHandlePtr handle, const RecombineSpec& spec) override {
    (void)handle;
    uint64_t n = 0;
    for (HbEngine* e : hs_) {  // continuations stay on the device that produced them
      uint64_t k = 0;
      Check(hb_recombine(e, spec.shuffle ? 1 : 0, &k), "Recombine", e);
      n += k;
    }
    layer_idx_++;
    layer_roots_ = static_cast<size_t>(n);
    DeviceRayBatch dev;
    dev.backend_ptr = h_;
    dev.count = static_cast<size_t>(n);
    return RootRaySource::FromDevice(dev);
  }

  size_t DrainExits(std::vector<ExitRayRecord>& out) override {
    out.clear();
    if (!egress_) {
      return 0;  // device-fused path: exits are reduced into the image, never materialised
    }
    static_assert(sizeof(ExitRayRecord) == sizeof(HbExitRecord), "ExitRayRecord layout (exit_seam.hpp:40-53)");
    for (HbEngine* e : hs_) {
      uint64_t n = 0;
      Check(hb_drain_exits(e, nullptr, nullptr, 0, &n), "DrainExits", e);
      const size_t old = out.size();
      out.resize(old + static_cast<size_t>(n));
      if (n != 0) {
        Check(hb_drain_exits(e, reinterpret_cast<HbExitRecord*>(out.data() + old), nullptr, n, &n), "DrainExits", e);
      }
    }
    return out.size();
  }
  // Session-level egress of the older contract: here the same destructive drain (the driver calls one or the other).
  size_t ReadbackExitRays(std::vector<ExitRayRecord>& out) override { return DrainExits(out); }

  void ReadbackXyzAccum(XyzImageData& xyz, float& landed_weight) override {
    ThrowDeferred();
    MergeDevices();
    landed_weight = 0.0f;  // the seam's contract is "copies the running scalar out" (trace_backend.hpp), the C ABI adds
    Check(hb_readback_xyz(h_, xyz.data, &landed_weight), "ReadbackXyzAccum");
  }

  void ReadbackClassLanes(std::vector<float>& lane_data, size_t& class_count) override {
    lane_data.clear();
    class_count = 0;
    if (render_ == nullptr || !raypath_color_ || raypath_color_->classes_.empty()) {
      return;  // base behaviour: no colour config, nothing to drain
    }
    const size_t pix = static_cast<size_t>(render_->resolution_[0]) * static_cast<size_t>(render_->resolution_[1]);
    lane_data.resize(raypath_color_->classes_.size() * pix);
    uint32_t n = 0;
    MergeDevices();
    Check(hb_readback_class_lanes(h_, lane_data.data(), lane_data.size(), &n), "ReadbackClassLanes");
    class_count = n;
    lane_data.resize(static_cast<size_t>(n) * pix);
  }

  // Never throws: the reference driver ends sessions from a noexcept scope guard (simulator.cpp:1509-1515). A device
  // overflow reported here is kept and thrown by the next BeginSession / TraceLayer / ReadbackXyzAccum instead.
  void EndSession() override {
    for (HbEngine* e : hs_) {  // close every device's session even when one reports an error
      const int rc = hb_end_session(e);
      if (rc != HB_OK && deferred_error_.empty()) {
        deferred_error_ = std::string("B200TraceBackend::EndSession: ") + hb_last_error(e);
      }
    }
  }

  size_t GetLastBatchStochasticCrystalSampleCount() const override { return stochastic_shapes_last_upload_; }
  // Orientations are drawn per ray on the device: every root of layer 0 and every continuation that enters a later
  // layer draws one, so the count is the number of rays traced through layers whose axis is not deterministic
  // (AxisDistribution::IsAxisDeterministic, the seam's single predicate, trace_backend.hpp:589-625).
  size_t GetLastBatchStochasticOrientationSampleCount() const override { return orientation_draws_; }

 private:
  void ThrowDeferred() {
    if (!deferred_error_.empty()) {
      std::string msg;
      msg.swap(deferred_error_);
      throw std::runtime_error(msg);
    }
  }

  // Device 0 gathers the other devices' accumulators (asynchronous; ordered by events on the engines' streams).
  void MergeDevices() {
    for (size_t r = 1; r < hs_.size(); r++) {
      Check(hb_merge_from_peer(hs_[0], hs_[r]), "MergeDevices", hs_[0]);
    }
  }

  void Check(int status, const char* what, HbEngine* which = nullptr) {
    if (status == HB_OK) {
      return;
    }
    std::string msg = std::string("B200TraceBackend::") + what + ": " + hb_last_error(which ? which : h_);
    if (status == HB_ERR_NO_DEVICE || status == HB_ERR_CUDA) {
      throw BackendUnavailableError(msg);
    }
    throw std::runtime_error(msg);
  }

  static bool SceneHasStochasticShapes(const SceneConfig& scene) {
    for (const auto& ms : scene.ms_) {
      for (const auto& st : ms.setting_) {
        if (!IsDeterministic(st.crystal_.param_)) {
          return true;
        }
      }
    }
    return false;
  }

  static HbDist ToHbDist(const Distribution& d) {
    static_assert(static_cast<int>(DistributionType::kNoRandom) == HB_DIST_NO_RANDOM &&
                      static_cast<int>(DistributionType::kUniform) == HB_DIST_UNIFORM &&
                      static_cast<int>(DistributionType::kGaussian) == HB_DIST_GAUSSIAN &&
                      static_cast<int>(DistributionType::kZigzag) == HB_DIST_ZIGZAG &&
                      static_cast<int>(DistributionType::kLaplacian) == HB_DIST_LAPLACIAN &&
                      static_cast<int>(DistributionType::kGaussianLegacy) == HB_DIST_GAUSSIAN_LEGACY,
                  "DistributionType order");
    return HbDist{ static_cast<uint32_t>(d.type), d.center, d.spread };
  }

  // CrystalParam -> the shape part of HbCrystalDesc (distributions of the shape scalars + sync groups,
  // crystal_config.hpp:62-76); the axis part is not needed for resampling.
  static HbCrystalDesc ToCrystalDesc(const CrystalConfig& cfg) {
    HbCrystalDesc d{};
    d.id = cfg.id_;
    static_assert(kShapeScalarCount == 10, "HbCrystalDesc::sync_group follows the ShapeScalar order");
    if (const auto* prism = std::get_if<PrismCrystalParam>(&cfg.param_)) {
      d.kind = 0;
      d.height[0] = ToHbDist(prism->h_);
      for (int i = 0; i < 6; i++) d.face_dist[i] = ToHbDist(prism->d_[i]);
      for (int i = 0; i < kShapeScalarCount; i++) d.sync_group[i] = prism->sync_group_[i];
    } else {
      const auto& pyr = std::get<PyramidCrystalParam>(cfg.param_);
      d.kind = 1;
      d.height[0] = ToHbDist(pyr.h_pyr_u_);
      d.height[1] = ToHbDist(pyr.h_prs_);
      d.height[2] = ToHbDist(pyr.h_pyr_l_);
      for (int i = 0; i < 6; i++) d.face_dist[i] = ToHbDist(pyr.d_[i]);
      for (int i = 0; i < kShapeScalarCount; i++) d.sync_group[i] = pyr.sync_group_[i];
      d.wedge_upper_deg = pyr.wedge_angle_u_;
      d.wedge_lower_deg = pyr.wedge_angle_l_;
    }
    return d;
  }

  void StartGeometryClock(const SceneConfig& scene, uint32_t seed) {
    stochastic_population_cnt_ = 0;
    for (size_t li = 0; li < scene.ms_.size(); li++) {
      for (size_t ci = 0; ci < scene.ms_[li].setting_.size(); ci++) {
        const ScatteringSetting& st = scene.ms_[li].setting_[ci];
        if (IsDeterministic(st.crystal_.param_)) {
          continue;
        }
        const HbCrystalDesc desc = ToCrystalDesc(st.crystal_);
        for (HbEngine* e : hs_) {
          Check(hb_auto_resample(e, static_cast<uint32_t>(li), static_cast<uint32_t>(ci), &desc, seed, geom_draws_), "AutoResample", e);
          geom_draws_ += 1u << 24;  // disjoint shape-stream ranges per population, device and upload
        }
        stochastic_population_cnt_++;
      }
    }
  }

  static void FillTables(const Crystal& c, HbCrystalTables* out) {
    std::memset(out, 0, sizeof(*out));
    const size_t fc = c.PolygonFaceCount();
    out->face_cnt = static_cast<uint32_t>(fc);
    for (size_t i = 0; i < fc; i++) {
      std::memcpy(out->plane[i], c.GetPolygonFaceNormal() + i * 3, 3 * sizeof(float));
      out->plane[i][3] = c.GetPolygonFaceDist()[i];
      out->face_fn[i] = static_cast<uint8_t>(c.GetFn(static_cast<IdType>(i)) & 0xFF);
    }
    const size_t tc = detail::CountEntrySubTris(c.CfGeom());
    if (tc > HB_MAX_SUBTRIS) {
      throw BackendUnavailableError("B200TraceBackend: crystal needs more than 64 entry sub-triangles");
    }
    std::vector<detail::EntrySubTri> sub(tc);
    if (tc > 0) {
      detail::BuildEntrySubTris(c.CfGeom(), sub.data());
    }
    out->subtri_cnt = static_cast<uint32_t>(tc);
    for (size_t t = 0; t < tc; t++) {
      std::memcpy(out->tri_v[t], sub[t].v, 9 * sizeof(float));
      std::memcpy(out->tri_n[t], sub[t].n, 3 * sizeof(float));
      out->tri_area[t] = sub[t].area;
      out->tri_face[t] = static_cast<uint8_t>(sub[t].face_id);
    }
  }

  static void FillAxis(const AxisDistribution& axis, HbAxisSampler* out) {
    std::memset(out, 0, sizeof(*out));
    const auto decision = lat_path::SelectLatPath(axis);
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
      std::memcpy(out->lut_theta, lut->theta.data(), sizeof(out->lut_theta));
      std::memcpy(out->lut_cdf, lut->cdf.data(), sizeof(out->lut_cdf));
      std::memcpy(out->lut_flip, lut->flip_prob.data(), sizeof(out->lut_flip));
    }
  }

  static void FillSimple(const DeviceFilterDesc& d, HbSimpleFilter* s) {
    std::memset(s, 0, sizeof(*s));
    s->kind = d.type;
    s->path_len = d.canonical_len;
    std::memcpy(s->path, d.canonical_bytes, std::min<size_t>(d.canonical_len, HB_MAX_FILTER_PATH));
    s->entry_fn = d.has_entry ? 1 : -1;
    s->exit_fn = d.has_exit ? 1 : -1;
    s->min_len = d.min_len;
    s->max_len = d.max_len;
    std::memcpy(s->dir, d.dir, sizeof(d.dir));
    s->cos_radii = d.radii_c;
    s->crystal_id = d.crystal_id;
  }

  static void FillFilter(const FilterConfig& cfg, const Crystal& crystal, const AxisDistribution& axis, HbFilterDesc* out) {
    std::memset(out, 0, sizeof(*out));
    const DeviceFilterDesc d = detail::BuildDeviceFilterDesc(cfg, crystal, axis);
    out->kind = d.type;
    out->action = d.action;
    out->symmetry = d.symmetry;
    out->fn_period = d.fn_period;
    out->sigma_a = d.sigma_a;
    out->d_applicable = d.d_applicable;
    if (d.type == kDeviceFilterTypeComplex) {
      std::vector<DeviceFilterDesc> subs;
      std::vector<uint8_t> counts;
      detail::BuildComplexSubDescs(std::get<ComplexFilterParam>(cfg.param_), crystal, d.symmetry, d.sigma_a,
                                   d.d_applicable != 0, subs, counts);
      if (counts.size() > HB_MAX_FILTER_TERMS) {
        throw BackendUnavailableError("B200TraceBackend: complex filter with more than 8 OR-clauses");
      }
      out->term_cnt = static_cast<uint32_t>(counts.size());
      size_t k = 0;
      for (size_t o = 0; o < counts.size(); o++) {
        if (counts[o] > 4) {
          throw BackendUnavailableError("B200TraceBackend: complex filter clause with more than 4 AND-terms");
        }
        out->term_len[o] = counts[o];
        for (uint8_t a = 0; a < counts[o]; a++, k++) {
          FillSimple(subs[k], &out->terms[o][a]);
        }
      }
    } else {
      FillSimple(d, &out->simple);
    }
  }

  // One HbColorGroup per symmetry value of the placement's colour predicates (BuildColorSpecGroups,
  // filter_spec.cpp:389-425, in device-descriptor form).
  static void FillColorGroups(const ColorGateTable& gate_table, IdType layer, const ScatteringSetting& st,
                              const Crystal& crystal, HbCrystalPopulation* out) {
    const ColorGatePlacement placement = ColorGatePlacementFor(gate_table, layer, st.crystal_.id_);
    if (placement.predicates_.empty()) {
      return;
    }
    const ColorPlacementGrouping grouping = GroupPlacementBySymmetry(placement);
    const size_t group_cnt = grouping.group_symmetry_.size();
    if (group_cnt > HB_MAX_COLOR_GROUPS) {
      throw BackendUnavailableError("B200TraceBackend: more than 4 colour symmetry groups on one placement");
    }
    std::vector<ComplexFilterParam> cfps(group_cnt);
    std::vector<std::vector<uint8_t>> bits(group_cnt);
    for (size_t k = 0; k < placement.predicates_.size(); k++) {
      const size_t gi = grouping.group_of_[k];
      cfps[gi].filters_.push_back({ { kInvalidId, placement.predicates_[k] } });
      bits[gi].push_back(placement.bits_[k]);
    }
    for (size_t gi = 0; gi < group_cnt; gi++) {
      if (bits[gi].size() > HB_MAX_FILTER_TERMS) {
        throw BackendUnavailableError("B200TraceBackend: more than 8 colour predicates in one symmetry group");
      }
      FilterConfig fc{};
      fc.id_ = kInvalidId;
      fc.symmetry_ = grouping.group_symmetry_[gi];
      fc.action_ = FilterConfig::kFilterIn;
      fc.param_ = FilterParam{ cfps[gi] };
      FillFilter(fc, crystal, st.crystal_.axis_, &out->color_groups[gi].filter);
      std::memset(out->color_groups[gi].bit, 0xFF, sizeof(out->color_groups[gi].bit));
      std::copy(bits[gi].begin(), bits[gi].end(), out->color_groups[gi].bit);
    }
    out->color_group_cnt = static_cast<uint32_t>(group_cnt);
  }

  void UploadScene(const SceneConfig& scene) {
    // Geometry pool per stochastic population: kPoolShapes slots (filled here once with MakeCrystal so the
    // filter / colour descriptors have a crystal to probe; redrawn on the device at every BeginSession), one
    // shape per 32 consecutive rays (the CPU path's geometry clock, simulator.hpp:144-151).
    std::vector<HbLayer> layers(scene.ms_.size());
    std::vector<std::vector<HbCrystalPopulation>> pops(scene.ms_.size());
    std::vector<std::unique_ptr<std::vector<HbCrystalTables>>> shapes;
    stochastic_shapes_last_upload_ = 0;
    // Raypath colour (Design 2): the same tables the CPU gate uses (cpu_trace_backend.cpp:264-270,371-384)
    const RaypathColorConfig empty_color;
    const RaypathColorConfig& color_cfg = raypath_color_ ? *raypath_color_ : empty_color;
    const ColorGateTable gate_table = BuildColorGateTable(color_cfg, scene);
    const ColorClassTable class_table = BuildColorClassTable(color_cfg, scene, gate_table);
    for (size_t li = 0; li < scene.ms_.size(); li++) {
      const MsInfo& ms = scene.ms_[li];
      pops[li].resize(ms.setting_.size());
      for (size_t ci = 0; ci < ms.setting_.size(); ci++) {
        const ScatteringSetting& st = ms.setting_[ci];
        HbCrystalPopulation& p = pops[li][ci];
        std::memset(&p, 0, sizeof(p));
        p.proportion = st.crystal_proportion_;
        p.crystal_id = st.crystal_.id_;
        const bool deterministic = IsDeterministic(st.crystal_.param_);
        const uint32_t n = deterministic ? 1u : kPoolShapes;
        auto pool = std::make_unique<std::vector<HbCrystalTables>>(n);
        Crystal first;
        for (uint32_t s = 0; s < n; s++) {
          Crystal c = MakeCrystal(rng_, st.crystal_.param_);
          FillTables(c, &(*pool)[s]);
          if (s == 0) {
            first = c;
          }
        }
        if (!deterministic) {
          stochastic_shapes_last_upload_ += n;
        }
        p.shape_cnt = n;
        p.shapes = pool->data();
        shapes.push_back(std::move(pool));
        FillAxis(st.crystal_.axis_, &p.axis);
        FillFilter(st.filter_, first, st.crystal_.axis_, &p.filter);
        FillColorGroups(gate_table, static_cast<IdType>(li), st, first, &p);
      }
      layers[li].prob = ms.prob_;
      layers[li].population_cnt = static_cast<uint32_t>(pops[li].size());
      layers[li].populations = pops[li].data();
    }
    HbScene s{};
    s.max_hits = static_cast<uint32_t>(scene.max_hits_);
    s.layer_cnt = static_cast<uint32_t>(layers.size());
    s.layers = layers.data();
    const SunParam& sun = scene.light_source_.param_;
    s.sun_lon = (sun.azimuth_ + 180.0f) * math::kDegreeToRad;
    s.sun_lat = -sun.altitude_ * math::kDegreeToRad;
    s.sun_half_angle = (sun.diameter_ * 0.5f) * math::kDegreeToRad;
    if (class_table.classes_.size() > HB_MAX_COLOR_CLASSES) {
      throw BackendUnavailableError("B200TraceBackend: more than 16 colour classes");
    }
    s.color_classes.class_cnt = static_cast<uint32_t>(class_table.classes_.size());
    for (size_t c = 0; c < class_table.classes_.size(); c++) {
      s.color_classes.bits[c] = class_table.classes_[c].member_bits_;
      if (class_table.classes_[c].combine_ == ColorClassCombine::kAll) {
        s.color_classes.combine_all_mask |= 1u << c;
      }
    }
    for (HbEngine* e : hs_) {
      Check(hb_set_scene(e, &s), "UploadScene", e);
    }
    layer_axis_stochastic_.assign(scene.ms_.size(), false);
    for (size_t li = 0; li < scene.ms_.size(); li++) {
      for (const auto& st : scene.ms_[li].setting_) {
        if (!st.crystal_.axis_.IsAxisDeterministic()) {
          layer_axis_stochastic_[li] = true;
        }
      }
    }
    scene_ = &scene;
    color_uploaded_ = raypath_color_;
  }

  static HbProjParams ToProj(const RenderConfig& render) {
    const Rotation rot = MakeCameraRotation(render);
    const auto short_pix = static_cast<float>(std::min(render.resolution_[0], render.resolution_[1]));
    const lm_proj::ProjParams p = BuildProjParams(render, rot, short_pix);
    static_assert(sizeof(lm_proj::ProjParams) == sizeof(HbProjParams), "ProjParams layout");
    HbProjParams hp;
    std::memcpy(&hp, &p, sizeof(hp));
    return hp;
  }

  void UploadRender(const RenderConfig& render) {
    std::vector<HbProjParams> all{ ToProj(render) };  // render 0 = the session's renderer
    for (const RenderConfig* extra : extra_renders_) {
      all.push_back(ToProj(*extra));
    }
    for (HbEngine* e : hs_) {
      Check(hb_set_renders(e, static_cast<uint32_t>(all.size()), all.data()), "UploadRender", e);
    }
    render_ = &render;
    render_snapshot_dirty_ = false;
  }

 public:
  // ---- extensions beyond the seam -------------------------------------------------------
  // Additional renderers projected from the SAME trace (the reference refuses multi-renderer configs on its
  // device route, server.cpp:402-437). Pointers must outlive the sessions; call between sessions.
  void SetExtraRenders(std::vector<const RenderConfig*> extra) {
    if (extra.size() + 1 > HB_MAX_RENDERS) {
      throw BackendUnavailableError("B200TraceBackend: more than 8 renderers");
    }
    extra_renders_ = std::move(extra);
    render_snapshot_dirty_ = true;
  }
  // ReadbackXyzAccum for renderer `index` (0 = the session's own, 1.. = SetExtraRenders order).
  void ReadbackXyzAccumOf(uint32_t index, XyzImageData& xyz, float& landed_weight) {
    MergeDevices();
    landed_weight = 0.0f;
    Check(hb_readback_xyz_render(h_, index, xyz.data, &landed_weight), "ReadbackXyzAccumOf");
  }
  // RenderConsumer::PrepareSnapshot + PostSnapshot on the device: 8-bit sRGB frame of renderer `index`,
  // without draining the accumulator. Returns the snapshot intensity.
  float SnapshotSrgb(uint32_t index, const RenderConfig& cfg, uint8_t* rgb8) {
    HbSnapshotDesc d{};
    d.intensity_factor = cfg.intensity_factor_;
    for (int j = 0; j < 3; j++) {
      d.ray_color[j] = cfg.ray_color_[j];
      d.background[j] = cfg.background_[j];
    }
    float intensity = 0.0f;
    MergeDevices();
    Check(hb_snapshot(h_, index, &d, rgb8, nullptr, &intensity), "SnapshotSrgb");
    return intensity;
  }

 private:

  std::string deferred_error_;         // an error EndSession could not throw (see EndSession)
  HbEngine* h_ = nullptr;              // device 0 of hs_: the one that is read back
  std::vector<HbEngine*> hs_;          // one engine per device
  std::vector<size_t> share_;          // root rays of the current session per device
  uint64_t ray_base_ = 0;              // global index of the next session's first root (multi-device sessions)
  std::vector<bool> layer_axis_stochastic_;  // per layer: some population draws its orientation
  size_t layer_roots_ = 0;             // rays entering the layer about to be traced
  size_t orientation_draws_ = 0;       // GetLastBatchStochasticOrientationSampleCount of the current session
  RandomNumberGenerator rng_;
  const SceneConfig* scene_ = nullptr;
  const RenderConfig* render_ = nullptr;
  std::shared_ptr<const RaypathColorConfig> raypath_color_;   // of the current session
  std::shared_ptr<const RaypathColorConfig> color_uploaded_;  // the one the device tables were built from
  std::vector<const RenderConfig*> extra_renders_;
  bool render_snapshot_dirty_ = false;
  size_t layer_cnt_ = 0;
  size_t layer_idx_ = 0;
  size_t stochastic_shapes_last_upload_ = 0;
  bool egress_ = false;
  static constexpr uint32_t kPoolShapes = 256;
  uint32_t geom_draws_ = 0;  // shape-stream base handed to the engine's geometry clock
  size_t stochastic_population_cnt_ = 0;
};

}  // namespace lumice

#endif  // ADAPTER_B200_TRACE_BACKEND_HPP_
