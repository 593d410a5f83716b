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

#include <algorithm>
#include <cmath>
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

// Uniform grid over the chords' boxes: lets the oracle answer closest-hit queries on scenes where testing every
// chord per ray is out of reach (the 100 k-curve scene has 1.3 M chords). Deliberately unlike the product's
// structure (Morton LBVH over runs of chords): fixed cells, CSR lists, a double-precision cell walk. It only
// narrows WHICH chords are tested; the test and the (t, id) ordering are rdc_math.h's, as in the brute force.
// tests/test_oracle_cpu.py checks grid == brute force, bit for bit.
struct Grid {
  double x0 = 0, y0 = 0, cell = 1, inv_cell = 1;
  int nx = 0, ny = 0;
  std::vector<uint32_t> start;  // [nx*ny+1]
  std::vector<uint32_t> items;  // chord ids, ascending inside a cell
  bool empty() const { return nx == 0; }
};

struct ChordSet {
  std::vector<Chord> chords;           // original order: segment by segment, k ascending
  std::vector<uint32_t> seg_base;      // [n_segments+1]
  Grid grid;                           // filled by build_grid; empty -> brute force
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

inline void build_grid(ChordSet& cs) {
  Grid& g = cs.grid;
  const size_t n = cs.chords.size();
  if (n == 0) return;
  double xmin = 1e300, ymin = 1e300, xmax = -1e300, ymax = -1e300;
  for (const Chord& c : cs.chords) {
    xmin = std::min({xmin, (double)c.ax, (double)c.bx}); xmax = std::max({xmax, (double)c.ax, (double)c.bx});
    ymin = std::min({ymin, (double)c.ay, (double)c.by}); ymax = std::max({ymax, (double)c.ay, (double)c.by});
  }
  const double extent = std::max(xmax - xmin, ymax - ymin);
  const int side = (int)std::min(2048.0, std::max(1.0, std::ceil(std::sqrt((double)n / 2.0))));
  g.cell = std::max(extent / side, 1e-6);
  g.inv_cell = 1.0 / g.cell;
  // one cell of margin all round: hits land strictly inside, whatever the rounding of a hit point
  g.x0 = xmin - g.cell;
  g.y0 = ymin - g.cell;
  g.nx = (int)std::floor((xmax - g.x0) * g.inv_cell) + 2;
  g.ny = (int)std::floor((ymax - g.y0) * g.inv_cell) + 2;
  // a chord is listed in every cell its box, grown by a hundredth of a cell, touches: a hit point that rounding
  // moves across a cell border is still found from either side
  const double grow = 0.01 * g.cell;
  auto span = [&](const Chord& c, int& ix0, int& ix1, int& iy0, int& iy1) {
    ix0 = std::max(0, (int)std::floor((std::min(c.ax, c.bx) - grow - g.x0) * g.inv_cell));
    ix1 = std::min(g.nx - 1, (int)std::floor((std::max(c.ax, c.bx) + grow - g.x0) * g.inv_cell));
    iy0 = std::max(0, (int)std::floor((std::min(c.ay, c.by) - grow - g.y0) * g.inv_cell));
    iy1 = std::min(g.ny - 1, (int)std::floor((std::max(c.ay, c.by) + grow - g.y0) * g.inv_cell));
  };
  g.start.assign((size_t)g.nx * g.ny + 1, 0u);
  for (const Chord& c : cs.chords) {
    int ix0, ix1, iy0, iy1;
    span(c, ix0, ix1, iy0, iy1);
    for (int iy = iy0; iy <= iy1; ++iy)
      for (int ix = ix0; ix <= ix1; ++ix) g.start[(size_t)iy * g.nx + ix + 1]++;
  }
  for (size_t i = 1; i < g.start.size(); ++i) g.start[i] += g.start[i - 1];
  g.items.resize(g.start.back());
  std::vector<uint32_t> fill(g.start.begin(), g.start.end() - 1);
  for (uint32_t id = 0; id < (uint32_t)n; ++id) {
    int ix0, ix1, iy0, iy1;
    span(cs.chords[id], ix0, ix1, iy0, iy1);
    for (int iy = iy0; iy <= iy1; ++iy)
      for (int ix = ix0; ix <= ix1; ++ix) g.items[fill[(size_t)iy * g.nx + ix]++] = id;
  }
}

struct Hit {
  float t = std::numeric_limits<float>::infinity();
  float s = 0.0f;
  uint32_t id = 0xFFFFFFFFu;
  bool valid() const { return id != 0xFFFFFFFFu; }
};

// every chord, no acceleration structure; [skip_lo,skip_hi] (inclusive chord ids) are invisible
// `primary`: see rdc_inv_dd (rdc_math.h)
inline Hit closest_hit_grid(const ChordSet& cs, float ox, float oy, float dx, float dy, bool primary, uint32_t skip_lo,
                            uint32_t skip_hi);

inline Hit closest_hit(const ChordSet& cs, float ox, float oy, float dx, float dy, bool primary, uint32_t skip_lo,
                       uint32_t skip_hi) {
  if (!cs.grid.empty()) return closest_hit_grid(cs, ox, oy, dx, dy, primary, skip_lo, skip_hi);
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

// The same answer through the grid: cells are visited in the order the ray crosses them (cell walk in double
// precision), every chord listed in a visited cell gets the shared test, and the walk stops once the best hit lies
// clearly in front of the border to the next cell. `t` of a hit is measured in units of |D| like the walk's own
// parameter ((P - O).D / D.D), so the two compare directly.
inline Hit closest_hit_grid(const ChordSet& cs, float ox, float oy, float dx, float dy, bool primary, uint32_t skip_lo,
                            uint32_t skip_hi) {
  const Grid& g = cs.grid;
  Hit h;
  const float inv_dd = rdc_inv_dd(dx, dy, primary);
  const double Ox = ox, Oy = oy, Dx = dx, Dy = dy;
  const double gx1 = g.x0 + g.nx * g.cell, gy1 = g.y0 + g.ny * g.cell;
  // where the ray is inside the grid's rectangle
  double t0 = 0.0, t1 = std::numeric_limits<double>::infinity();
  auto clip = [&](double o, double d, double lo, double hi) {
    if (d == 0.0) return o >= lo && o <= hi;
    double a = (lo - o) / d, b = (hi - o) / d;
    if (a > b) std::swap(a, b);
    t0 = std::max(t0, a);
    t1 = std::min(t1, b);
    return t0 <= t1;
  };
  if (!clip(Ox, Dx, g.x0, gx1) || !clip(Oy, Dy, g.y0, gy1)) return h;
  const double px = Ox + t0 * Dx, py = Oy + t0 * Dy;
  int ix = std::min(g.nx - 1, std::max(0, (int)std::floor((px - g.x0) * g.inv_cell)));
  int iy = std::min(g.ny - 1, std::max(0, (int)std::floor((py - g.y0) * g.inv_cell)));
  const int sx = Dx > 0 ? 1 : -1, sy = Dy > 0 ? 1 : -1;
  const double inf = std::numeric_limits<double>::infinity();
  // ray parameter at which the walk leaves the current cell along each axis
  double tx = Dx != 0.0 ? (g.x0 + (ix + (sx > 0 ? 1 : 0)) * g.cell - Ox) / Dx : inf;
  double ty = Dy != 0.0 ? (g.y0 + (iy + (sy > 0 ? 1 : 0)) * g.cell - Oy) / Dy : inf;
  const double dtx = Dx != 0.0 ? g.cell / std::fabs(Dx) : inf, dty = Dy != 0.0 ? g.cell / std::fabs(Dy) : inf;
  for (;;) {
    const size_t cell = (size_t)iy * g.nx + ix;
    for (uint32_t k = g.start[cell]; k < g.start[cell + 1]; ++k) {
      const uint32_t c = g.items[k];
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
    const double t_leave = std::min(tx, ty);
    // every chord with a hit beyond this cell is listed in a later cell; fp32 rounding of t: 1e-4 relative + absolute is generous
    if (h.valid() && (double)h.t < t_leave - 1e-4 * (1.0 + std::fabs(t_leave))) break;
    if (tx <= ty) {
      ix += sx;
      tx += dtx;
    } else {
      iy += sy;
      ty += dty;
    }
    if (ix < 0 || iy < 0 || ix >= g.nx || iy >= g.ny) break;
  }
  return h;
}

}  // namespace oracle

#endif
