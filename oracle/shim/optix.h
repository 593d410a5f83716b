// oracle/shim/optix.h — TEST INFRASTRUCTURE. Host stand-in for the OptiX 7.3 / CUDA device headers, just
// large enough to compile the reference's optixHello/DeviceCode.cu and params.h UNMODIFIED with g++
// (they are #included from /root/reference where they lie; nothing of them is copied).
//
// What the shim supplies is exactly what the reference gets from closed code:
//   * optixTrace(): closest hit by brute force over the chord list (oracle_common.h), then a call to the
//     reference's own __closesthit__ch / __miss__ms;
//   * optixGet*/optixSet* accessors over a per-thread record of the current hit and payload;
//   * curand_uniform(): the Philox stream the product uses, in the reference's draw order;
//   * sincospif(): the product's deterministic polynomial.
#ifndef ORACLE_SHIM_OPTIX_H
#define ORACLE_SHIM_OPTIX_H

#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __global__
#define __host__
#define __constant__
#define __forceinline__ inline __attribute__((always_inline))

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct uint2 { unsigned int x, y; };
struct uint3 { unsigned int x, y, z; };

typedef void* CUstream;
typedef unsigned long long OptixTraversableHandle;
typedef unsigned int OptixVisibilityMask;
enum { OPTIX_RAY_FLAG_NONE = 0 };

// the per-pixel generator state shrinks to nothing: the pixel is recovered from the state's ADDRESS
struct curandState_t { char unused; };

inline int float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
inline float int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
inline float int_as_float(unsigned int i) { float f; std::memcpy(&f, &i, 4); return f; }

void sincospif(float x, float* s, float* c);
float curand_uniform(curandState_t* state);

uint3 optixGetLaunchIndex();
float optixGetCurveParameter();
float optixGetRayTmax();
unsigned int optixGetPrimitiveIndex();
float3 optixGetWorldRayDirection();
float3 optixGetWorldRayOrigin();
unsigned int optixGetPayload_5();
void optixSetPayload_0(unsigned int v);
void optixSetPayload_1(unsigned int v);
void optixSetPayload_2(unsigned int v);
void optixSetPayload_3(unsigned int v);
void optixSetPayload_4(unsigned int v);
inline void optixSetPayload_0(int v) { optixSetPayload_0((unsigned int)v); }
inline void optixSetPayload_1(int v) { optixSetPayload_1((unsigned int)v); }
inline void optixSetPayload_2(int v) { optixSetPayload_2((unsigned int)v); }
inline void optixSetPayload_3(int v) { optixSetPayload_3((unsigned int)v); }
inline void optixSetPayload_4(int v) { optixSetPayload_4((unsigned int)v); }

void optixTrace(OptixTraversableHandle handle, float3 rayOrigin, float3 rayDirection, float tmin, float tmax, float rayTime,
                OptixVisibilityMask visibilityMask, unsigned int rayFlags, unsigned int SBToffset, unsigned int SBTstride,
                unsigned int missSBTIndex, unsigned int& p0, unsigned int& p1, unsigned int& p2, unsigned int& p3,
                unsigned int& p4, unsigned int& p5);

#endif
