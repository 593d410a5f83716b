// microbench.cu — measurement aid, not part of the render path: a dependent-FFMA kernel whose rate is the
// FP32 roofline denominator bench.py reports against (MEASURED_PEAKS.json carries no FP32 figure).
#include "device_scene.h"

namespace {

constexpr int kChains = 8;

__global__ void __launch_bounds__(256) k_ffma(float* sink, int iters, float a, float b) {
  float x[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) x[c] = (float)(threadIdx.x + c) * 1e-3f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < kChains; ++c) x[c] = fmaf(x[c], a, b);
  }
  float s = 0.0f;
#pragma unroll
  for (int c = 0; c < kChains; ++c) s += x[c];
  if (s == 123.456f) sink[0] = s;  // never true; keeps the chains alive
}

// L2 read rate: every block streams 128-bit loads that bypass L1 (ld.global.cg) over a buffer small enough to stay in
// L2 (the caller picks the size: 32 MB fits either half of the 126 MB L2), several passes; the first pass warms it.
__global__ void __launch_bounds__(256) k_l2_read(const uint4* __restrict__ buf, size_t n_vec, int passes, unsigned int* sink) {
  unsigned int acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int p = 0; p < passes; ++p)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
      const uint4 v = __ldcg(buf + i);
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
  if (acc == 0x12345678u) sink[0] = acc;  // practically never; keeps the loads alive
}

}  // namespace

// L2 read microbenchmark: the denominator of the L2-side roofline of large scenes (rays * B_ray / t / L2 peak).
// Launches `launches` kernels that each read `bytes` (a multiple of 16, resident in L2) `passes` times from `buffer`;
// *bytes_per_launch = bytes * passes. Enqueue-only.
extern "C" int rdc_microbench_l2(const void* buffer, size_t bytes, int passes, int launches, float* sink, double* bytes_per_launch,
                                 rdc_stream stream) {
  if (!buffer || bytes < 16 || passes <= 0 || launches <= 0 || !sink) {
    rdc::set_error("microbench: bad argument");
    return RDC_E_INVALID;
  }
  const size_t n_vec = bytes / 16;
  for (int l = 0; l < launches; ++l)
    k_l2_read<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(static_cast<const uint4*>(buffer), n_vec, passes, reinterpret_cast<unsigned int*>(sink));
  RDC_CUDA(cudaGetLastError());
  if (bytes_per_launch) *bytes_per_launch = (double)(n_vec * 16) * passes;
  return 0;
}

extern "C" int rdc_microbench_fp32(int iters, int launches, float* sink, double* flops_per_launch, rdc_stream stream) {
  if (iters <= 0 || launches <= 0 || !sink) {
    rdc::set_error("microbench: bad argument");
    return RDC_E_INVALID;
  }
  const int blocks = 148 * 8, threads = 256;
  for (int l = 0; l < launches; ++l) k_ffma<<<blocks, threads, 0, (cudaStream_t)stream>>>(sink, iters, 0.999f, 1e-4f);
  RDC_CUDA(cudaGetLastError());
  if (flops_per_launch) *flops_per_launch = 2.0 * 64.0 * (double)iters * blocks * threads;
  return 0;
}
