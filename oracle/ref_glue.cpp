// ref_glue.cpp — TEST INFRASTRUCTURE: runs the reference's OWN device programs on the host.
//
// `#include "DeviceCode.cu"` below pulls in /root/reference/optixHello/DeviceCode.cu (and, through it, the
// reference's params.h) unmodified, from where they lie — the build recipe (oracle/Makefile, target _ref)
// passes -I/root/reference/optixHello and puts oracle/shim/ first on the include path so that <optix.h>,
// <cuda_runtime.h>, <curand_kernel.h> ... resolve to the host stand-ins. The result, oracle/_ref/
// libref_oracle.so, is git-ignored and never part of the product.
//
// So __raygen__rg, __closesthit__ch, __miss__ms, interpolate, calculateSpline, calculateSplineNormal and
// isRayRight executed here ARE the reference's code, compiled with the switches it ships with
// (USE_DIFFUSION_CURVE_SAVE true, USE_AA true, MAX_TRACE_DEPTH 2). This file supplies only what the
// reference takes from OptiX and cuRAND: the launch loop, optixTrace, the accessors and the random stream.
#include <math.h>
#include <omp.h>

#include <vector>

#include "oracle_common.h"
#include "shim/optix.h"

extern "C" void __closesthit__ch();
extern "C" void __miss__ms();

namespace {

struct HitRecord {
  float u = 0, t = 0;
  unsigned int prim = 0;
  float3 origin{}, direction{};
  unsigned int payload[6] = {0, 0, 0, 0, 0, 0};
};

struct ThreadState {
  uint3 launch_index{};
  unsigned int draws = 0;  // curand_uniform calls since the pixel started
  int level = 0;           // optixTrace nesting
  HitRecord cur;
  uint32_t first_hit = 0xFFFFFFFFu;
};

thread_local ThreadState tls;

struct Global {
  const oracle::ChordSet* chords = nullptr;
  uint32_t seed = 0, frame = 0;
  uint32_t* hit_ids = nullptr;
  int n_iter = 0;
  uint32_t row_begin = 0, width = 0;
} g;

}  // namespace

#include "DeviceCode.cu"  // the reference's file: defines `params` and the three programs

void sincospif(float x, float* s, float* c) { rdc_sincospi(x, s, c); }

float curand_uniform(curandState_t* state) {
  const uint32_t pixel = (uint32_t)(state - params.curandStates);
  const unsigned int n = tls.draws++;
  const rdc_u4 r = rdc_philox4x32_10(pixel, n / 3, 0u, 0u, g.seed, g.frame);  // 3 draws per ray (USE_AA true)
  const unsigned int lane = n % 3;
  return rdc_u01(lane == 0 ? r.x : lane == 1 ? r.y : r.z);
}

uint3 optixGetLaunchIndex() { return tls.launch_index; }
float optixGetCurveParameter() { return tls.cur.u; }
float optixGetRayTmax() { return tls.cur.t; }
unsigned int optixGetPrimitiveIndex() { return tls.cur.prim; }
float3 optixGetWorldRayDirection() { return tls.cur.direction; }
float3 optixGetWorldRayOrigin() { return tls.cur.origin; }
unsigned int optixGetPayload_5() { return tls.cur.payload[5]; }
void optixSetPayload_0(unsigned int v) { tls.cur.payload[0] = v; }
void optixSetPayload_1(unsigned int v) { tls.cur.payload[1] = v; }
void optixSetPayload_2(unsigned int v) { tls.cur.payload[2] = v; }
void optixSetPayload_3(unsigned int v) { tls.cur.payload[3] = v; }
void optixSetPayload_4(unsigned int v) { tls.cur.payload[4] = v; }

void optixTrace(OptixTraversableHandle, float3 o, float3 d, float, float, float, OptixVisibilityMask, unsigned int,
                unsigned int, unsigned int, unsigned int, unsigned int& p0, unsigned int& p1, unsigned int& p2,
                unsigned int& p3, unsigned int& p4, unsigned int& p5) {
  const oracle::ChordSet& cs = *g.chords;
  uint32_t skip_lo = 1, skip_hi = 0;
  if (tls.level > 0) {
    // a continuation ray leaves the portal's target segment at the parameter of the enclosing hit
    const unsigned int seg = tls.cur.prim;
    const uint32_t target = params.curve_map_inverse[params.curve_connect[params.curve_map[seg]]] + params.curve_index[seg];
    int klo, khi;
    rdc_portal_skip(tls.cur.u, (int)(cs.seg_base[target + 1] - cs.seg_base[target]), &klo, &khi);
    skip_lo = cs.seg_base[target] + (uint32_t)klo;
    skip_hi = cs.seg_base[target] + (uint32_t)khi;
  }
  const oracle::Hit h = oracle::closest_hit(cs, o.x, o.y, d.x, d.y, tls.level == 0, skip_lo, skip_hi);
  if (tls.level == 0) tls.first_hit = h.id;

  const HitRecord saved = tls.cur;
  tls.cur.payload[0] = p0; tls.cur.payload[1] = p1; tls.cur.payload[2] = p2;
  tls.cur.payload[3] = p3; tls.cur.payload[4] = p4; tls.cur.payload[5] = p5;
  tls.cur.origin = o;
  tls.cur.direction = d;
  tls.level++;
  if (h.valid()) {
    const oracle::Chord& ch = cs.chords[h.id];
    tls.cur.u = rdc_hit_u(ch.k, ch.K, h.s);
    tls.cur.t = h.t;
    tls.cur.prim = ch.seg;
    __closesthit__ch();
  } else {
    __miss__ms();
  }
  tls.level--;
  p0 = tls.cur.payload[0]; p1 = tls.cur.payload[1]; p2 = tls.cur.payload[2];
  p3 = tls.cur.payload[3]; p4 = tls.cur.payload[4];
  tls.cur = saved;

  if (tls.level == 0 && g.hit_ids) {
    const unsigned int ray = tls.draws / 3 - 1;
    const size_t local = (size_t)(tls.launch_index.y - g.row_begin) * g.width + tls.launch_index.x;
    g.hit_ids[local * g.n_iter + ray] = tls.first_hit;
  }
}

extern "C" {

int ref_threads(void) { return omp_get_max_threads(); }
static int g_search = 0;  // 0: every chord per ray; 1: through the uniform grid of oracle_common.h
void ref_set_search(int mode) { g_search = mode; }

// the switches the reference was compiled with
void ref_switches(int* use_diffusion_curve_save, int* use_aa, int* max_trace_depth) {
  *use_diffusion_curve_save = USE_DIFFUSION_CURVE_SAVE;
  *use_aa = USE_AA;
  *max_trace_depth = MAX_TRACE_DEPTH;
}

// Same contract as oracle_render (oracle_port.cpp). Returns -1 when the requested switches are not the
// ones the reference is compiled with.
int ref_render(const rdc_scene_arrays* a, const rdc_accel_options* o, const rdc_frame_params* p, float* image,
               float* blur_map, uint32_t* hit_ids, int threads) {
  if ((p->use_diffusion_curve_save != 0) != (bool)USE_DIFFUSION_CURVE_SAVE || (p->use_aa != 0) != (bool)USE_AA ||
      p->max_trace_depth != MAX_TRACE_DEPTH || p->strip_stride > 1)
    return -1;
  oracle::ChordSet cs = oracle::build_chords(*a, *o);
  if (g_search == 1) oracle::build_grid(cs);
  const size_t n_vertices = a->n_vertices;
  std::vector<float3> vertices(n_vertices);
  for (size_t i = 0; i < n_vertices; ++i) vertices[i] = float3{a->vertices[3 * i], a->vertices[3 * i + 1], a->vertices[3 * i + 2]};
  std::vector<int> segment_indices(a->segment_indices, a->segment_indices + a->n_segments);
  static curandState_t state_base[1];

  g.chords = &cs;
  g.seed = p->seed;
  g.frame = p->frame;
  g.hit_ids = hit_ids;
  g.n_iter = (int)ceilf(p->number_of_rays_per_pixel);
  g.row_begin = p->row_begin;
  g.width = p->image_width;

  params = Params{};
  // band-local output buffers: the programs index them with the full-image pixel number
  params.image = reinterpret_cast<float4*>(image) - (ptrdiff_t)p->row_begin * p->image_width;
  params.blur_map = blur_map - (ptrdiff_t)p->row_begin * p->image_width;
  params.image_width = p->image_width;
  params.image_height = p->image_height;
  params.curandStates = state_base;  // never dereferenced: curand_uniform() reads the pixel off the address
  params.number_of_rays_per_pixel = p->number_of_rays_per_pixel;
  params.vertices = vertices.data();
  params.segmentIndices = segment_indices.data();
  params.curve_map = const_cast<unsigned int*>(a->curve_map);
  params.curve_index = const_cast<unsigned int*>(a->curve_index);
  params.curve_connect = const_cast<int*>(a->curve_connect);
  params.curve_map_inverse = const_cast<unsigned int*>(a->curve_map_inverse);
  params.color_left_index = reinterpret_cast<uint2*>(const_cast<uint32_t*>(a->color_left_index));
  params.color_left = reinterpret_cast<float3*>(const_cast<float*>(a->color_left));
  params.color_left_u = const_cast<float*>(a->color_left_u);
  params.color_right_index = reinterpret_cast<uint2*>(const_cast<uint32_t*>(a->color_right_index));
  params.color_right = reinterpret_cast<float3*>(const_cast<float*>(a->color_right));
  params.color_right_u = const_cast<float*>(a->color_right_u);
  params.blur_index = reinterpret_cast<uint2*>(const_cast<uint32_t*>(a->blur_index));
  params.blur = const_cast<float*>(a->blur);
  params.blur_u = const_cast<float*>(a->blur_u);
  params.weight_index = reinterpret_cast<uint2*>(const_cast<uint32_t*>(a->weight_index));
  params.weight = const_cast<float*>(a->weight);
  params.weight_u = const_cast<float*>(a->weight_u);
  params.weight_degree_index = reinterpret_cast<uint2*>(const_cast<uint32_t*>(a->weight_degree_index));
  params.weight_degree = const_cast<float*>(a->weight_degree);
  params.weight_degree_u = const_cast<float*>(a->weight_degree_u);
  params.zoom_factor = p->zoom_factor;
  params.offset_x = p->offset_x;
  params.offset_y = p->offset_y;
  params.frame = p->frame;

  if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (uint32_t iy = p->row_begin; iy < p->row_end; ++iy)
    for (uint32_t ix = 0; ix < p->image_width; ++ix) {
      tls.launch_index = uint3{ix, iy, 0};
      tls.draws = 0;
      tls.level = 0;
      __raygen__rg();
      // the reference never writes image.w (DeviceCode.cu:176-178); give it the product's convention
      params.image[(size_t)iy * p->image_width + ix].w = 1.0f;
    }
  g.chords = nullptr;
  return 0;
}

}  // extern "C"
