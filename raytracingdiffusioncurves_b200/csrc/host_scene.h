// host_scene.h — the structure-of-arrays scene the ingest produces (host memory).
// One std::vector per array of struct Params (reference params.h:60-92).
#ifndef RDC_HOST_SCENE_H
#define RDC_HOST_SCENE_H

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/rdc_b200.h"

struct rdc_host_scene {
  int image_width = 0, image_height = 0;
  std::vector<float> vertices;  // x,y,z per control point
  std::vector<uint32_t> segment_indices, curve_map, curve_index, curve_map_inverse;
  std::vector<int32_t> curve_connect;

  // a family of stops: {start,count} per curve, values (stride floats per stop), parameter u per stop
  struct StopList {
    int stride = 1;
    std::vector<uint32_t> index;
    std::vector<float> value;
    std::vector<float> u;
    uint32_t size() const { return (uint32_t)u.size(); }
    uint32_t& start() { return index[index.size() - 2]; }
    uint32_t& count() { return index[index.size() - 1]; }
    void begin_curve(uint32_t running_total) {
      index.push_back(running_total);
      index.push_back(0);
    }
    void push(const float* v, float uu) {
      value.insert(value.end(), v, v + stride);
      u.push_back(uu);
    }
  };
  StopList color_left, color_right, blur, weight, weight_degree;

  rdc_host_scene() {
    color_left.stride = 3;
    color_right.stride = 3;
  }
  // appends the +INF sentinels; call once after the last curve
  void seal();
  void view(rdc_scene_arrays* out) const;
  bool sealed = false;
  uint32_t n_true[5] = {0, 0, 0, 0, 0};  // stop counts without sentinels: left, right, blur, weight, exponent
};

namespace rdc {
// Throws std::runtime_error (XmlError included) on malformed input.
void ingest_xml_text(const char* text, size_t len, const rdc_ingest_options& opts, rdc_host_scene& scene);
void ingest_xml_file(const std::string& path, const rdc_ingest_options& opts, rdc_host_scene& scene);
}  // namespace rdc

#endif
