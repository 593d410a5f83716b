// oracle_common.h — TEST INFRASTRUCTURE. Shared by the two CPU oracles (oracle_port.cpp, ref_glue.cpp).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load what
// is built from this directory. The product (raytracingdiffusioncurves_b200/) never includes or links it.
//
// What lives here is the part of the path the reference delegates to closed code: the
// chord list of every spline segment and a brute-force closest hit over ALL chords — no tree, so it checks
// the product's LBVH independently. The arithmetic comes from the product's rdc_math.h on purpose: hit
// indices are compared bit-for-bit, which needs one definition of the intersection.
#ifndef ORACLE_COMMON_H
#define ORACLE_COMMON_H

#include <cstdint>
#include <limits>
#include <vector>

#include "../include/rdc_b200.h"
#include "../raytracingdiffusioncurves_b200/csrc/rdc_math.h"

namespace oracle {

struct Chord {
  float ax, ay, bx, by;
  uint32_t seg;
  int k, K;
};

struct ChordSet {
  std::vector<Chord> chords;           // original order: segment by segment, k ascending
  std::vector<uint32_t> seg_base;      // [n_segments+1]
};

inline void control_points(const rdc_scene_arrays& a, uint32_t seg, rdc_f2 v[4]) {
  const float* p = a.vertices + 3 * (size_t)a.segment_indices[seg];
  for (int i = 0; i < 4; ++i) v[i] = rdc_f2{p[3 * i], p[3 * i + 1]};
}

inline ChordSet build_chords(const rdc_scene_arrays& a, const rdc_accel_options& o) {
  ChordSet cs;
  cs.seg_base.resize(a.n_segments + 1);
  for (uint32_t s = 0; s < a.n_segments; ++s) {
    rdc_f2 v[4];
    control_points(a, s, v);
    const int K = rdc_chord_count(v[0], v[1], v[2], v[3], o.flatness_tolerance, o.max_chords_per_segment);
    cs.seg_base[s] = (uint32_t)cs.chords.size();
    for (int k = 0; k < K; ++k) {
      rdc_f2 p = rdc_spline_point(rdc_chord_u(k, K), v[0], v[1], v[2], v[3]);
      rdc_f2 q = rdc_spline_point(rdc_chord_u(k + 1, K), v[0], v[1], v[2], v[3]);
      cs.chords.push_back(Chord{p.x, p.y, q.x, q.y, s, k, K});
    }
  }
  cs.seg_base[a.n_segments] = (uint32_t)cs.chords.size();
  return cs;
}

struct Hit {
  float t = std::numeric_limits<float>::infinity();
  float s = 0.0f;
  uint32_t id = 0xFFFFFFFFu;
  bool valid() const { return id != 0xFFFFFFFFu; }
};

// every chord, no acceleration structure; [skip_lo,skip_hi] (inclusive chord ids) are invisible
// `primary`: see rdc_inv_dd (rdc_math.h)
inline Hit closest_hit(const ChordSet& cs, float ox, float oy, float dx, float dy, bool primary, uint32_t skip_lo,
                       uint32_t skip_hi) {
  Hit h;
  const float inv_dd = rdc_inv_dd(dx, dy, primary);
  const uint32_t n = (uint32_t)cs.chords.size();
  for (uint32_t c = 0; c < n; ++c) {
    if (c >= skip_lo && c <= skip_hi) continue;
    const Chord& ch = cs.chords[c];
    float t, s;
    if (!rdc_ray_chord(ox, oy, dx, dy, inv_dd, ch.ax, ch.ay, ch.bx, ch.by, &t, &s)) continue;
    if (rdc_hit_closer(t, c, h.t, h.id)) {
      h.t = t;
      h.s = s;
      h.id = c;
    }
  }
  return h;
}

}  // namespace oracle

#endif
