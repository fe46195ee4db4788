// hb_filter.h — raypath symmetry reduction + filter match, shared by the host scene builder
// (canonical bytes of a configured raypath) and the optics kernel's emit gate.
//
// Behavioural spec: reference src/core/shared/filter_shared.h:53-315 (PCanonicalShiftInPlace_dev,
// ReduceBuffer_dev, DeviceFilterMatch*), and Crystal::ReduceRaypath via FillCanonicalBytes
// (src/core/device_filter_desc.cpp:22-30).
#ifndef HB_FILTER_H_
#define HB_FILTER_H_

#include <stdint.h>

#include "halotrace_b200.h"

#if defined(__CUDACC__)
#define HB_HD __host__ __device__ __forceinline__
#else
#define HB_HD inline
#endif

namespace hb {

// Face numbers: 1,2 basal; 3..8 prism; 13..18 upper pyramid; 23..28 lower pyramid.
// P symmetry: rotate the prism index so the first lateral face of the path becomes 3.
HB_HD void filter_p_shift(uint8_t* data, uint32_t size) {
  int first = -1;
  for (uint32_t i = 0; i < size; ++i) {
    uint32_t x = data[i];
    if (x < 3u) continue;
    uint32_t ring = x / 10u;
    int idx = static_cast<int>(x % 10u);
    if (first < 0) first = idx;
    idx = (idx + 6 - first) % 6 + 3;
    data[i] = static_cast<uint8_t>(ring * 10u + static_cast<uint32_t>(idx));
  }
}

HB_HD bool filter_lex_less(const uint8_t* a, const uint8_t* b, uint32_t size) {
  for (uint32_t i = 0; i < size; ++i) {
    if (a[i] != b[i]) return a[i] < b[i];
  }
  return false;
}

// Canonical representative under the enabled symmetries (bit0 P, bit1 B, bit2 D).
HB_HD void filter_reduce(uint8_t* data, uint32_t size, uint32_t symmetry, int sigma_a, bool d_applicable) {
  if (symmetry == 0u) return;
  if (symmetry & 1u) filter_p_shift(data, size);
  if ((symmetry & 4u) && d_applicable) {
    uint8_t alt[HB_MAX_HITS];
    for (uint32_t i = 0; i < size; ++i) {
      uint32_t x = data[i];
      if (x < 3u) {
        alt[i] = static_cast<uint8_t>(x);
        continue;
      }
      uint32_t ring = x / 10u;
      int k = static_cast<int>(x % 10u) - 3;
      int mk = ((sigma_a - k) % 6 + 6) % 6;
      alt[i] = static_cast<uint8_t>(ring * 10u + static_cast<uint32_t>(mk + 3));
    }
    if (symmetry & 1u) filter_p_shift(alt, size);
    if (filter_lex_less(alt, data, size)) {
      for (uint32_t i = 0; i < size; ++i) data[i] = alt[i];
    }
  }
  if (symmetry & 2u) {
    uint8_t alt[HB_MAX_HITS];
    bool changed = false;
    for (uint32_t i = 0; i < size; ++i) {
      uint32_t x = data[i];
      if (x <= 2u) {
        alt[i] = static_cast<uint8_t>(3u - x);
        changed = true;
      } else if (x >= 13u && x <= 18u) {
        alt[i] = static_cast<uint8_t>(x + 10u);
        changed = true;
      } else if (x >= 23u && x <= 28u) {
        alt[i] = static_cast<uint8_t>(x - 10u);
        changed = true;
      } else {
        alt[i] = static_cast<uint8_t>(x);
      }
    }
    if (changed && filter_lex_less(alt, data, size)) {
      for (uint32_t i = 0; i < size; ++i) data[i] = alt[i];
    }
  }
}

// fn_path: the exit's path as FACE NUMBERS. dir: world-space exit direction.
HB_HD bool filter_match_simple(const HbFilterDesc& f, const HbSimpleFilter& s, const uint8_t* fn_path, uint32_t len,
                               const float* dir, uint32_t crystal_id) {
  const bool reduce = !(f.fn_period < 0 || f.symmetry == 0u);
  if (s.kind == 0u) return true;
  if (s.kind == 1u) {
    if (len != s.path_len || len > HB_MAX_FILTER_PATH) return false;
    uint8_t buf[HB_MAX_FILTER_PATH];
    for (uint32_t i = 0; i < len; ++i) buf[i] = fn_path[i];
    if (reduce) filter_reduce(buf, len, f.symmetry, f.sigma_a, f.d_applicable != 0u);
    for (uint32_t i = 0; i < len; ++i) {
      if (buf[i] != s.path[i]) return false;
    }
    return true;
  }
  if (s.kind == 2u) {
    if (len == 0u || len < s.min_len) return false;
    if (s.max_len != 0u && len > s.max_len) return false;
    const bool has_entry = s.entry_fn >= 0, has_exit = s.exit_fn >= 0;
    if (!has_entry && !has_exit) return true;
    uint8_t ee[2];
    uint32_t n = 0;
    if (has_entry) ee[n++] = fn_path[0];
    if (has_exit) ee[n++] = fn_path[len - 1u];
    if (reduce) {
      filter_reduce(ee, n, f.symmetry, f.sigma_a, f.d_applicable != 0u);
      if (n != s.path_len) return false;
    }
    for (uint32_t i = 0; i < n; ++i) {
      if (ee[i] != s.path[i]) return false;
    }
    return true;
  }
  if (s.kind == 3u) return s.dir[0] * dir[0] + s.dir[1] * dir[1] + s.dir[2] * dir[2] > s.cos_radii;
  if (s.kind == 4u) return crystal_id == s.crystal_id;
  return false;
}

HB_HD bool filter_check(const HbFilterDesc& f, const uint8_t* fn_path, uint32_t len, const float* dir,
                        uint32_t crystal_id) {
  bool m = false;
  if (f.kind == 5u) {
    for (uint32_t o = 0; o < f.term_cnt && !m; ++o) {
      bool all = true;
      for (uint32_t a = 0; a < f.term_len[o] && all; ++a) {
        all = filter_match_simple(f, f.terms[o][a], fn_path, len, dir, crystal_id);
      }
      m = all;
    }
  } else {
    m = filter_match_simple(f, f.simple, fn_path, len, dir, crystal_id);
  }
  return f.action == 0u ? m : !m;
}

// Longest path (number of interactions) an exit may have and still be ADMITTED by filter `f`; 0xFFFFFFFF = unbounded.
// A filter_in raypath filter admits only paths of its own length, an entry-exit filter with max_len only paths up to
// it; a complex filter admits what any OR-term admits, a term what all of its AND-factors admit. filter_out and the
// direction / crystal filters bound nothing. The engine ends a layer's hit loop at the largest bound of its
// populations: later interactions can only produce exits the filter rejects (filter-fail terminates,
// simulator.cpp:678-730), so results are unchanged.
HB_HD uint32_t filter_factor_max_len(const HbSimpleFilter& s) {
  if (s.kind == 1u) return s.path_len;
  if (s.kind == 2u && s.max_len != 0u) return s.max_len;
  return 0xFFFFFFFFu;
}
HB_HD uint32_t filter_max_len(const HbFilterDesc& f) {
  if (f.kind == 0u || f.action != 0u) return 0xFFFFFFFFu;
  if (f.kind != 5u) return filter_factor_max_len(f.simple);
  uint32_t best = 0u;
  for (uint32_t o = 0; o < f.term_cnt; ++o) {
    uint32_t term = 0xFFFFFFFFu;
    for (uint32_t a = 0; a < f.term_len[o]; ++a) {
      const uint32_t b = filter_factor_max_len(f.terms[o][a]);
      term = b < term ? b : term;
    }
    best = term > best ? term : best;
  }
  return best;
}

}  // namespace hb

#endif  // HB_FILTER_H_
