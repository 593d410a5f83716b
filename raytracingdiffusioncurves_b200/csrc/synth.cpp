// synth.cpp — deterministic synthetic diffusion-curve sets. Nothing like it
// ships with the reference; the scene is emitted as XML in the reference's schema (Appendix B.1) so it goes
// through the same loader as the bundled files.
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/rdc_b200.h"

namespace rdc {
void set_error(const char* fmt, ...);
}

namespace {

struct SplitMix64 {
  uint64_t s;
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
  int below(int n) { return (int)(uniform() * n); }
};

void append(std::string& s, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
void append(std::string& s, const char* fmt, ...) {
  char buf[256];
  va_list ap;
  va_start(ap, fmt);
  int n = vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  s.append(buf, (size_t)n);
}

}  // namespace

extern "C" int rdc_synth_xml(uint32_t n_curves, uint32_t width, uint32_t height, uint64_t seed, char** out_text,
                             size_t* out_len) {
  if (!out_text || n_curves == 0 || width == 0 || height == 0) {
    rdc::set_error("synth: bad argument");
    return RDC_E_INVALID;
  }
  SplitMix64 rng{seed};
  std::string xml;
  xml.reserve((size_t)n_curves * 700 + 256);
  append(xml, "<!DOCTYPE CurveSetXML>\n<curve_set image_width=\"%u\" image_height=\"%u\" nb_curves=\"%u\">\n", width, height, n_curves);
  const double two_pi = 6.283185307179586;
  for (uint32_t c = 0; c < n_curves; ++c) {
    double x = rng.uniform() * width, y = rng.uniform() * height;
    double theta = rng.uniform() * two_pi;
    double len = 16.0 * std::exp(rng.uniform() * std::log(256.0 / 16.0));  // log-uniform in [16,256]
    xml += " <curve nb_control_points=\"4\" nb_left_colors=\"2\" nb_right_colors=\"2\" nb_blur_points=\"2\">\n  <control_points_set>\n";
    for (int p = 0; p < 4; ++p) {
      if (p > 0) {
        double a = theta + (rng.uniform() - 0.5);
        x += len / 3.0 * std::cos(a);
        y += len / 3.0 * std::sin(a);
      }
      double cx = std::fmin(std::fmax(x, 0.0), (double)width - 1.0);
      double cy = std::fmin(std::fmax(y, 0.0), (double)height - 1.0);
      append(xml, "   <control_point x=\"%.3f\" y=\"%.3f\" />\n", cx, cy);
    }
    xml += "  </control_points_set>\n";
    const char* sides[2] = {"left", "right"};
    for (const char* side : sides) {
      append(xml, "  <%s_colors_set>\n", side);
      for (int id = 0; id <= 10; id += 10) {
        int r = rng.below(256), g = rng.below(256), b = rng.below(256);
        append(xml, "   <%s_color R=\"%d\" G=\"%d\" B=\"%d\" globalID=\"%d\" />\n", side, r, g, b, id);
      }
      append(xml, "  </%s_colors_set>\n", side);
    }
    xml += "  <blur_points_set>\n";
    for (int id = 0; id <= 10; id += 10) {
      int v = rng.uniform() < 0.75 ? 0 : 1 + rng.below(4);
      append(xml, "   <best_scale value=\"%d\" globalID=\"%d\" />\n", v, id);
    }
    xml += "  </blur_points_set>\n </curve>\n";
  }
  xml += "</curve_set>\n";
  char* buf = (char*)std::malloc(xml.size() + 1);
  if (!buf) {
    rdc::set_error("synth: out of memory");
    return RDC_E_LIMIT;
  }
  std::memcpy(buf, xml.c_str(), xml.size() + 1);
  *out_text = buf;
  if (out_len) *out_len = xml.size();
  return 0;
}
