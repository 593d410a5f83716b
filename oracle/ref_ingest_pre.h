// ref_ingest_pre.h — TEST INFRASTRUCTURE: what stands in front of the reference's own ingest lines in
// oracle/_ref/ref_ingest_gen.cpp (see ref_extract.sh). Supplies the types the lines use (float3, uint2: shim/optix.h),
// the reference's params.h (USE_DIFFUSION_CURVE_SAVE) and optixHello.h (the helpers' prototypes, CALL_CHECK), rapidxml
// from the reference's support/ directory, and the container the epilogue copies the loop's locals into.
#ifndef REF_INGEST_PRE_H
#define REF_INGEST_PRE_H
#include <cstdlib>
#include <filesystem>
#include <string>
#include <vector>

#include "shim/optix.h"
#include "params.h"      // the reference's: -I$REF/optixHello
#include "optixHello.h"  // the reference's: prototypes of pushColor ... invSqrt, includes rapidxml_utils.hpp
#include <rapidxml/rapidxml.hpp>

struct RefScene {
  int width = 0, height = 0;
  std::vector<float3> vertices, color_left, color_right;
  std::vector<unsigned int> segmentIndices, curve_map, curve_map_inverse, curve_index;
  std::vector<int> curve_connect;
  std::vector<uint2> color_left_index, color_right_index, blur_index, weight_index, weight_degree_index;
  std::vector<float> color_left_u, color_right_u, blur, blur_u, weight, weight_u, weight_degree, weight_degree_u;
};
#endif
