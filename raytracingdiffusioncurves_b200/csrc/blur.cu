// blur.cu — separable gather blur with a per-destination-pixel sigma.
//
// Same result as helperKernels.cu:48-148 (gaussHorizontal, gaussVertical, gaussianBlur): taps
// k in [-ceil(3 sigma), +ceil(3 sigma)], weight expf(-(k*k)/(sigma+1e-6)^2) (no factor 2), clamp-to-edge,
// renormalised, all four channels, the vertical pass reads the SAME un-blurred sigma map.
// What differs is the execution: a grid sized to the GPU (148 SMs x 16 blocks) instead of a fixed 512x256
// launch, accumulation in registers instead of read-modify-write on dest[i], scratch supplied
// by the caller (or the stream-ordered allocator) instead of cudaMalloc/cudaFree per frame, and an
// optional device flag that turns both passes into (at most) one copy when every sigma is zero.
#include "device_scene.h"

namespace rdc {
namespace {

constexpr int kThreads = 256;
constexpr unsigned kMaxBlocks = 148 * 16;  // 148 SMs x 16 resident blocks of 256 threads at most
constexpr float kMinSigma = 1e-6f;   // helperKernels.cu:27
constexpr float kMaxTaps = 16384.0f; // |k| bound for non-finite or absurd sigmas (edge pixels repeat anyway)

struct Taps {
  int first, last;
  float sig_square;
};

__device__ __forceinline__ Taps taps_for(float sigma) {
  // k_size = 2*ceilf(3*sigma)+1; for (int k = -k_size/2; k <= k_size/2; k++)   (helperKernels.cu:65,74)
  float k_size = 2 * ceilf(3 * sigma) + 1;
  Taps t;
  float lo = -k_size / 2, hi = k_size / 2;
  if (!(k_size == k_size)) {  // NaN sigma: the reference's loop condition is false at once -> 0/0
    t.first = 0;
    t.last = -1;
  } else {
    lo = fmaxf(lo, -kMaxTaps);
    hi = fminf(hi, kMaxTaps);
    t.first = (int)lo;               // truncation toward zero, like the reference's int conversion
    t.last = (int)floorf(hi);        // largest int k with (float)k <= hi
  }
  t.sig_square = (sigma + kMinSigma) * (sigma + kMinSigma);
  return t;
}

__device__ __forceinline__ bool all_sigma_zero(const float* max_sigma) {
  return max_sigma != nullptr && __ldg(max_sigma) == 0.0f;
}

// rows [h_begin,h_end): scratch = horizontal blur of source. With an all-zero sigma map the pass is the
// identity, so it writes `dest` directly (nothing at all when dest == source).
__global__ void k_blur_horizontal(const float4* __restrict__ source, float4* scratch, float4* dest,
                                  const float* __restrict__ sigma, int width, int h_begin, int h_end, const float* max_sigma) {
  const size_t count = (size_t)width * (h_end - h_begin), first = (size_t)h_begin * width;
  const size_t stride = (size_t)blockDim.x * gridDim.x;
  const bool identity = all_sigma_zero(max_sigma);
  if (identity && dest == source) return;
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += stride) {
    const size_t i = first + j;
    if (identity) {
      dest[i] = source[i];
      continue;
    }
    const int x = (int)(i % width);
    const size_t row = i - x;
    const Taps t = taps_for(sigma[i]);
    float accum = 0.0f;
    float4 d = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    for (int k = t.first; k <= t.last; ++k) {
      const int sx = max(0, min(x + k, width - 1));
      const float g = expf(-(k * k) / t.sig_square);
      const float4 s = source[row + sx];
      accum += g;
      d.x += s.x * g;
      d.y += s.y * g;
      d.z += s.z * g;
      d.w += s.w * g;
    }
    scratch[i] = make_float4(d.x / accum, d.y / accum, d.z / accum, d.w / accum);
  }
}

// rows [row_begin,row_end): dest = vertical blur of scratch, rows clamped to [0,height-1]
__global__ void k_blur_vertical(const float4* __restrict__ scratch, float4* __restrict__ dest,
                                const float* __restrict__ sigma, int width, int height, int row_begin, int row_end,
                                const float* max_sigma) {
  if (all_sigma_zero(max_sigma)) return;
  const size_t count = (size_t)width * (row_end - row_begin), first = (size_t)row_begin * width;
  const size_t stride = (size_t)blockDim.x * gridDim.x;
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += stride) {
    const size_t i = first + j;
    const int x = (int)(i % width);
    const int y = (int)(i / width);
    const Taps t = taps_for(sigma[i]);
    float accum = 0.0f;
    float4 d = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    for (int k = t.first; k <= t.last; ++k) {
      const int sy = max(0, min(y + k, height - 1));
      const float g = expf(-(k * k) / t.sig_square);
      const float4 s = scratch[(size_t)sy * width + x];
      accum += g;
      d.x += s.x * g;
      d.y += s.y * g;
      d.z += s.z * g;
      d.w += s.w * g;
    }
    dest[i] = make_float4(d.x / accum, d.y / accum, d.z / accum, d.w / accum);
  }
}

inline unsigned grid_for(size_t n) {
  const size_t blocks = (n + kThreads - 1) / kThreads;
  return (unsigned)(blocks < kMaxBlocks ? blocks : kMaxBlocks);
}

}  // namespace

// forces the blur kernels' module to load now (a later first launch would otherwise do it, and loading may have to
// wait for the device to go idle — not something to meet between two barrier kernels)
int blur_preload() {
  cudaFuncAttributes fa;
  RDC_CUDA(cudaFuncGetAttributes(&fa, k_blur_horizontal));
  RDC_CUDA(cudaFuncGetAttributes(&fa, k_blur_vertical));
  return 0;
}

int gaussian_blur(float4* dest, const float4* src, const float* sigma, float4* scratch, int width, int height,
                  int row_begin, int row_end, const float* max_sigma, cudaStream_t stream, int halo_rows) {
  if (!dest || !src || !sigma || !scratch || width <= 0 || height <= 0 || row_begin < 0 || row_end > height ||
      row_begin >= row_end) {
    set_error("blur: bad argument");
    return RDC_E_INVALID;
  }
  // rows the vertical pass of the band can read
  const int h_begin = halo_rows < 0 ? 0 : (row_begin - halo_rows > 0 ? row_begin - halo_rows : 0);
  const int h_end = halo_rows < 0 ? height : (row_end + halo_rows < height ? row_end + halo_rows : height);
  const size_t n_all = (size_t)width * (h_end - h_begin);
  const size_t n_band = (size_t)width * (row_end - row_begin);
  k_blur_horizontal<<<grid_for(n_all), kThreads, 0, stream>>>(src, scratch, dest, sigma, width, h_begin, h_end, max_sigma);
  RDC_CUDA(cudaGetLastError());
  k_blur_vertical<<<grid_for(n_band), kThreads, 0, stream>>>(scratch, dest, sigma, width, height, row_begin, row_end, max_sigma);
  RDC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace rdc
