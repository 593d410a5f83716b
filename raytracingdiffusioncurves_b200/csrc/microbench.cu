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

}  // namespace

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
