// ref_blur_pre.h — TEST INFRASTRUCTURE: what stands in front of the reference's own blur kernels
// (helperKernels.cu:48-134) in oracle/_ref/ref_blur_gen.cpp (see ref_extract.sh): the CUDA words they use, as host code.
// A "launch" is an OpenMP loop over T host threads, each playing thread t of one block of T (blockDim.x = T,
// gridDim.x = 1): the kernels' grid-stride loops then cover every pixel exactly once, as on the device.
#ifndef REF_BLUR_PRE_H
#define REF_BLUR_PRE_H
#include <algorithm>
#include <cmath>

#include "shim/optix.h"  // float4, __global__ -> nothing
using std::max;
using std::min;
struct RefDim {
  int x = 1, y = 1, z = 1;
};
static thread_local RefDim threadIdx, blockIdx;
static RefDim blockDim, gridDim;
#endif
