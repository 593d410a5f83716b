// device_scene.h — device-resident scene layout shared by the accel build, the render kernels and the
// C ABI glue. Everything a kernel needs travels by value in DevScene (kernel parameter space).
#ifndef RDC_DEVICE_SCENE_H
#define RDC_DEVICE_SCENE_H

#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/rdc_b200.h"

// One LBVH node, 48 bytes = three 128-bit loads. A node stores the padded boxes of BOTH children, so a
// ray decides which children to visit from one node fetch. child >= 0: inner node index; child < 0:
// leaf, ~child is the chord's position in Morton order.
struct __align__(16) BvhNode {
  float4 lbox;  // xmin, ymin, xmax, ymax of the left child
  float4 rbox;
  int left, right;
  int parent;  // -1 for the root
  int pad;
};

struct DevStops {
  const uint2* index;  // {start,count} per curve
  const float* u;      // flat stop parameters (+INF sentinels at the end)
  const float* value;  // scalar families
  const float4* rgb;   // colour families (rgb, 0)
};

struct DevScene {
  // spline segments (reference Params: vertices, segmentIndices, curve_map, curve_index, curve_connect,
  // curve_map_inverse). Control points are float2 (z == 0 always); segment s owns points [4s,4s+4).
  const float2* vertices;
  const uint32_t* segment_indices;
  const uint32_t* curve_map;
  const uint32_t* curve_index;
  const int32_t* curve_connect;
  const uint32_t* curve_map_inverse;
  DevStops color_left, color_right, blur, weight, weight_degree;
  // acceleration structure
  const float4* chord_geom;        // [n_chords] Morton order: ax, ay, bx, by
  const uint4* chord_ids;          // [n_chords] Morton order: original chord id, segment, k, K
  const uint32_t* seg_chord_base;  // [n_segments+1] original id of chord 0 of each segment
  const uint32_t* seg_chord_count; // [n_segments]   K
  const BvhNode* nodes;            // [max(n_chords-1,1)]
  uint32_t n_segments, n_curves, n_chords, n_nodes;
};

struct rdc_scene {
  int device = 0;
  DevScene dev{};
  rdc_scene_info info{};
  std::vector<void*> allocations;  // everything to cudaFree on destroy
  // per-N table of the iterated base directions (DeviceCode.cu:110-112,167-171)
  float2* base_dirs = nullptr;
  float base_dirs_n = -1.0f;
  uint32_t base_dirs_capacity = 0;
  float* zero_sigma = nullptr;  // device float used when the caller passes no max_sigma
};

namespace rdc {
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
// accel.cu
int build_scene(const rdc_scene_arrays& a, const rdc_accel_options& o, cudaStream_t stream, rdc_scene** out);
void destroy_scene(rdc_scene* s);
int download_chords(const rdc_scene* s, float* geom, uint32_t* ids);
// render.cu
int render(rdc_scene* s, const rdc_frame_params& p, float4* image, float* blur_map, cudaStream_t stream);
// blur.cu
int gaussian_blur(float4* dest, const float4* src, const float* sigma, float4* scratch, int width, int height,
                  int row_begin, int row_end, const float* max_sigma, cudaStream_t stream);
int set_float(float* dest, unsigned n, float v, cudaStream_t stream);
}  // namespace rdc

#define RDC_CUDA(call)                                            \
  do {                                                            \
    cudaError_t e__ = (call);                                     \
    if (e__ != cudaSuccess) return rdc::cuda_fail(e__, #call);    \
  } while (0)

#endif
