// ingest.cpp — diffusion-curve XML -> structure-of-arrays scene.
//
// Behavioural restatement of the reference's loader (optixHello.cpp:212-515 and helpers :1302-1386),
// organised differently: a curve is first read into a small record (control points, the five stop
// families), then emitted. The arithmetic (operand types, order of float operations, the bit-trick
// reciprocal square root used for end caps) is kept, because the arrays must be identical.
//
//   Bezier -> B-spline matrix            optixHello.cpp:76-79, correctControlPoints :1335-1343
//   vertex orientation / centring        push4Points :1314-1332  (integer W/2, H/2)
//   end-cap teardrop                     :229-274, :290-329, getBezierTangent :1354-1357,
//                                        getEndcapPoints :1360-1369, invSqrt :1372-1386
//   colour stops (B,G,R swap, closing    pushColor :1302-1311, :333-410
//     stop, end-cap fill)
//   blur / weight / exponent stops       pushSingle :1346-1351, :414-511
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <stdexcept>

#include "host_scene.h"
#include "xml_dom.h"

namespace {

struct P2 {
  float x, y;
};

const rdc::XmlElement& need_child(const rdc::XmlElement& e, const char* name, size_t curve_no) {
  const rdc::XmlElement* c = e.child(name);
  if (!c) throw std::runtime_error("ingest: curve " + std::to_string(curve_no) + " has no <" + name + ">");
  return *c;
}

const char* need_attr(const rdc::XmlElement& e, const char* name) {
  const char* a = e.attr(name);
  if (!a) throw std::runtime_error(std::string("ingest: <") + e.name + "> lacks attribute " + name);
  return a;
}

// atof() of the reference (optixHello.cpp:226-227 and the stop loops), four million times for the 100 000-curve scene.
// std::from_chars is correctly rounded like strtod and several times faster; whatever it does not take in full
// (leading blanks or '+', hex floats, trailing text, out-of-range values) goes to atof itself.
double parse_number(const char* s) {
  const char* e = s + std::strlen(s);
  double v = 0.0;
  const auto r = std::from_chars(s, e, v);
  if (r.ec == std::errc() && r.ptr == e) return v;
  return std::atof(s);
}

// optixHello.cpp:1372-1386 — one Newton step on the 0x5f3759df seed (about 0.2 % error). Kept because the
// end-cap control points, hence the rendered teardrop, depend on it.
float inv_sqrt_bits(float number) {
  float x2 = number * 0.5F;
  uint32_t i;
  std::memcpy(&i, &number, 4);
  i = 0x5f3759df - (i >> 1);
  float y;
  std::memcpy(&y, &i, 4);
  return y * (1.5F - (x2 * y * y));
}

// derivative of a cubic Bezier (optixHello.cpp:1354-1357), float arithmetic
P2 bezier_tangent(float t, const P2* v) {
  float a3 = 3 * t * t;
  float a0 = -3 * t * t + 6 * t - 3;
  float a1 = 9 * t * t - 12 * t + 3;
  float a2 = -9 * t * t + 6 * t;
  return {a3 * v[3].x + v[0].x * a0 + v[1].x * a1 + v[2].x * a2, a3 * v[3].y + v[0].y * a0 + v[1].y * a1 + v[2].y * a2};
}

struct Emitter {
  rdc_host_scene& s;
  uint32_t next_vertex = 0;

  // Bezier control polygon -> uniform cubic B-spline control points with the same curve:
  // rows of {6,-7,2,0 / 0,2,-1,0 / 0,-1,2,0 / 0,2,-7,6}
  void push_segment(const P2* b, uint32_t curve, uint32_t ordinal) {
    static const float M[16] = {6, -7, 2, 0, 0, 2, -1, 0, 0, -1, 2, 0, 0, 2, -7, 6};
    for (int i = 0; i < 4; ++i) {
      const float* m = M + 4 * i;
      s.vertices.push_back(b[0].x * m[0] + b[1].x * m[1] + b[2].x * m[2] + b[3].x * m[3]);
      s.vertices.push_back(b[0].y * m[0] + b[1].y * m[1] + b[2].y * m[2] + b[3].y * m[3]);
      s.vertices.push_back(0.0f);
    }
    s.segment_indices.push_back(next_vertex);
    next_vertex += 4;
    s.curve_map.push_back(curve);
    s.curve_index.push_back(ordinal);
  }

  // teardrop through `end` whose axis is `tangent`
  void push_endcap(P2 end, P2 tangent, int size, uint32_t curve, uint32_t ordinal) {
    float inv = inv_sqrt_bits(tangent.x * tangent.x + tangent.y * tangent.y);
    float c = tangent.y * inv;
    float sn = -tangent.x * inv;
    P2 cap[4];
    cap[0] = end;
    cap[1] = {(-c - sn) * size + end.x, (-sn + c) * size + end.y};
    cap[2] = {(c - sn) * size + end.x, (sn + c) * size + end.y};
    cap[3] = end;
    push_segment(cap, curve, ordinal);
  }
};

void set_stop(rdc_host_scene::StopList& l, uint32_t i, const float* v) {
  for (int k = 0; k < l.stride; ++k) l.value[(size_t)i * l.stride + k] = v[k];
}
const float* get_stop(const rdc_host_scene::StopList& l, uint32_t i) { return &l.value[(size_t)i * l.stride]; }

float stop_u(const rdc::XmlElement& e, bool endcap) {
  // double arithmetic, rounded once: atof(..)/10.0f + (endcap ? 1.0f : 0.0f)
  return (float)(parse_number(need_attr(e, "globalID")) / 10.0f + (endcap ? 1.0f : 0.0f));
}

// blur / weight / exponent family. `set` may be null only when a default exists (has_default).
void scalar_family(rdc_host_scene::StopList& l, const rdc::XmlElement* set, const char* attr_name, bool endcap,
                   uint32_t n_curve_segments, bool has_default, float default_value, float endcap_placeholder) {
  l.begin_curve(l.size());
  if (!set) {
    if (!has_default) throw std::runtime_error("ingest: a curve lacks its blur_points_set");
    float us[2] = {0.0f, (float)n_curve_segments};
    for (float u : us) {
      l.push(&default_value, u);
      l.count()++;
    }
    return;
  }
  if (endcap) {
    l.push(&endcap_placeholder, 0.0f);
    l.count()++;
  }
  for (const rdc::XmlElement* e = set->first_child; e; e = e->next_sibling) {
    float v = (float)parse_number(need_attr(*e, attr_name));
    l.push(&v, stop_u(*e, endcap));
    l.count()++;
  }
  if (endcap) {
    if (l.count() < 2) throw std::runtime_error("ingest: end-capped curve with an empty stop list");
    set_stop(l, l.start(), get_stop(l, l.start() + 1));
    float last = l.value.back();
    l.push(&last, (float)n_curve_segments);
    l.count()++;
  }
}

}  // namespace

void rdc_host_scene::seal() {
  if (sealed) return;
  const float inf = std::numeric_limits<float>::infinity();
  // The portal colour filter indexes the LEFT arrays with the RIGHT list's {start,count}
  // (DeviceCode.cu:297), so both colour families are padded to a common length.
  size_t colour_len = (color_left.u.size() > color_right.u.size() ? color_left.u.size() : color_right.u.size()) + 2;
  auto pad = [&](StopList& l, size_t len) {
    l.u.resize(len, inf);
    l.value.resize(len * l.stride, 0.0f);
  };
  n_true[0] = color_left.size();
  n_true[1] = color_right.size();
  n_true[2] = blur.size();
  n_true[3] = weight.size();
  n_true[4] = weight_degree.size();
  pad(color_left, colour_len);
  pad(color_right, colour_len);
  pad(blur, blur.u.size() + 2);
  pad(weight, weight.u.size() + 2);
  pad(weight_degree, weight_degree.u.size() + 2);
  sealed = true;
}

void rdc_host_scene::view(rdc_scene_arrays* o) const {
  std::memset(o, 0, sizeof *o);
  o->image_width = image_width;
  o->image_height = image_height;
  o->n_vertices = (uint32_t)(vertices.size() / 3);
  o->n_segments = (uint32_t)segment_indices.size();
  o->n_curves = (uint32_t)curve_connect.size();
  o->vertices = vertices.data();
  o->segment_indices = segment_indices.data();
  o->curve_map = curve_map.data();
  o->curve_index = curve_index.data();
  o->curve_connect = curve_connect.data();
  o->curve_map_inverse = curve_map_inverse.data();
  o->n_color_left = n_true[0];
  o->n_color_right = n_true[1];
  o->n_blur = n_true[2];
  o->n_weight = n_true[3];
  o->n_weight_degree = n_true[4];
  o->color_left_index = color_left.index.data();
  o->color_left = color_left.value.data();
  o->color_left_u = color_left.u.data();
  o->color_right_index = color_right.index.data();
  o->color_right = color_right.value.data();
  o->color_right_u = color_right.u.data();
  o->blur_index = blur.index.data();
  o->blur = blur.value.data();
  o->blur_u = blur.u.data();
  o->weight_index = weight.index.data();
  o->weight = weight.value.data();
  o->weight_u = weight.u.data();
  o->weight_degree_index = weight_degree.index.data();
  o->weight_degree = weight_degree.value.data();
  o->weight_degree_u = weight_degree.u.data();
}

namespace rdc {

static void ingest_tree(const XmlElement& root, const rdc_ingest_options& opts, rdc_host_scene& s) {
  const bool orzan = opts.use_diffusion_curve_save != 0;
  s.image_width = std::atoi(need_attr(root, "image_width"));
  s.image_height = std::atoi(need_attr(root, "image_height"));
  const int half_w = s.image_width / 2, half_h = s.image_height / 2;
  const char* first_axis = orzan ? "y" : "x";
  const char* second_axis = orzan ? "x" : "y";
  const int cap_size = (int)opts.endcap_size;

  Emitter emit{s};
  uint32_t n_segments = 0;

  size_t ci = 0;
  for (const XmlElement* curve_el = root.first_child; curve_el; curve_el = curve_el->next_sibling, ++ci) {
    const XmlElement& curve = *curve_el;
    const uint32_t curve_no = (uint32_t)ci;

    // ---- control points ------------------------------------------------------------------------
    const XmlElement& cps = need_child(curve, "control_points_set", ci);
    std::vector<P2> pts;
    pts.reserve(cps.n_children);
    for (const XmlElement* cp = cps.first_child; cp; cp = cp->next_sibling) {
      pts.push_back({(float)parse_number(need_attr(*cp, first_axis)) - half_w,
                     (float)parse_number(need_attr(*cp, second_axis)) - half_h});
    }
    if (pts.size() < 4 || (pts.size() - 1) % 3 != 0)
      throw std::runtime_error("ingest: curve " + std::to_string(ci) + " needs 3k+1 control points, has " +
                               std::to_string(pts.size()));
    const char* cap_attr = curve.attr("use_endcap");
    const bool endcap = cap_attr && std::strcmp(cap_attr, "true") == 0;
    const char* connects = curve.attr("connects");
    s.curve_connect.push_back(connects ? std::stoi(std::string(connects)) : -1);
    s.curve_map_inverse.push_back(n_segments);

    uint32_t ordinal = 0;
    if (endcap) {
      P2 t = bezier_tangent(1e-3, &pts[0]);
      emit.push_endcap(pts[0], {-t.x, -t.y}, cap_size, curve_no, ordinal++);
    }
    for (size_t i = 0; i + 3 < pts.size(); i += 3) emit.push_segment(&pts[i], curve_no, ordinal++);
    if (endcap) {
      P2 t = bezier_tangent(1 - 1e-3, &pts[pts.size() - 4]);
      emit.push_endcap(pts.back(), t, cap_size, curve_no, ordinal++);
    }
    const uint32_t n_curve_segments = ordinal;

    // ---- colours -------------------------------------------------------------------------------
    rdc_host_scene::StopList& L = s.color_left;
    rdc_host_scene::StopList& R = s.color_right;
    L.begin_curve(L.size());
    R.begin_curve(R.size());
    const float zero3[3] = {0, 0, 0};
    if (endcap) {  // two placeholder stops (u = 0, 1) in both lists, filled below
      L.push(zero3, 0.0f); L.push(zero3, 1.0f);
      R.push(zero3, 0.0f); R.push(zero3, 1.0f);
    }
    auto read_colours = [&](rdc_host_scene::StopList& l, const char* set_name) {
      const XmlElement& set = need_child(curve, set_name, ci);
      for (const XmlElement* e = set.first_child; e; e = e->next_sibling) {
        float c[3] = {std::atoi(need_attr(*e, orzan ? "B" : "R")) / 255.0f,
                      std::atoi(need_attr(*e, "G")) / 255.0f,
                      std::atoi(need_attr(*e, orzan ? "R" : "B")) / 255.0f};
        l.push(c, stop_u(*e, endcap));
        l.count()++;
      }
      if (l.count() == 0) throw std::runtime_error("ingest: curve " + std::to_string(ci) + " has an empty " + set_name);
    };
    read_colours(L, "left_colors_set");
    read_colours(R, "right_colors_set");
    if (orzan) {  // closing stop at the curve's last parameter value (:370-378), right list first
      float u_close = (float)(int)(n_curve_segments - (endcap ? 1 : 0));
      float c[3];
      std::memcpy(c, get_stop(R, R.size() - 1), sizeof c);
      R.push(c, u_close); R.count()++;
      std::memcpy(c, get_stop(L, L.size() - 1), sizeof c);
      L.push(c, u_close); L.count()++;
    }
    if (endcap) {
      // leading cap: both lists get {first left, first right} on stops 0 and 1 (:384-390). The right
      // list is filled after the left one, i.e. from the already-updated left list's first real stop.
      set_stop(L, L.start(), get_stop(L, L.start() + 2));
      set_stop(L, L.start() + 1, get_stop(R, R.start() + 2));
      L.count() += 2;
      set_stop(R, R.start(), get_stop(L, L.start() + 2));
      set_stop(R, R.start() + 1, get_stop(R, R.start() + 2));
      R.count() += 2;
      // trailing cap: both lists get {last right, last left} at u = n-1, n (:394-405)
      float last_right[3], last_left[3];
      std::memcpy(last_right, get_stop(R, R.size() - 1), sizeof last_right);
      std::memcpy(last_left, get_stop(L, L.size() - 1), sizeof last_left);
      float u1 = (float)(int)(n_curve_segments - 1), u2 = (float)(int)n_curve_segments;
      L.push(last_right, u1); L.push(last_left, u2); L.count() += 2;
      R.push(last_right, u1); R.push(last_left, u2); R.count() += 2;
    }

    // ---- blur, weight, exponent ----------------------------------------------------------------
    scalar_family(s.blur, &need_child(curve, "blur_points_set", ci), "value", endcap, n_curve_segments, false, 0.0f, 0.0f);
    scalar_family(s.weight, curve.child("weight_set"), "w", endcap, n_curve_segments, true, 1.0f, 0.0f);
    scalar_family(s.weight_degree, curve.child("weight_degree_set"), "w", endcap, n_curve_segments, true,
                  opts.default_weight_degree, opts.default_weight_degree);

    n_segments += n_curve_segments;
  }
  if (s.curve_connect.empty()) throw std::runtime_error("ingest: the curve set holds no curves");
  // portal targets must exist and have at least as many segments as their source (DeviceCode.cu:228
  // indexes the target curve with the source segment's ordinal)
  const uint32_t n_curves = (uint32_t)s.curve_connect.size();
  for (uint32_t c = 0; c < n_curves; ++c) {
    int32_t t = s.curve_connect[c];
    if (t < 0) continue;
    if ((uint32_t)t >= n_curves) throw std::runtime_error("ingest: curve " + std::to_string(c) + " connects to a missing curve");
    auto segs = [&](uint32_t k) { return (k + 1 < n_curves ? s.curve_map_inverse[k + 1] : n_segments) - s.curve_map_inverse[k]; };
    if (segs((uint32_t)t) < segs(c)) throw std::runtime_error("ingest: portal target of curve " + std::to_string(c) + " has fewer segments");
  }
  s.seal();
}

void ingest_xml_text(const char* text, size_t len, const rdc_ingest_options& opts, rdc_host_scene& scene) {
  auto doc = xml_parse(text, len);
  ingest_tree(*doc->root, opts, scene);
}

void ingest_xml_file(const std::string& path, const rdc_ingest_options& opts, rdc_host_scene& scene) {
  auto doc = xml_parse_file(path);
  ingest_tree(*doc->root, opts, scene);
}

}  // namespace rdc
