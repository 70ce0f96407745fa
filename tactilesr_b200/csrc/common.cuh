// Shared helpers for the tactilesr_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define TSR_OK 0
#define TSR_ERR_BAD_ARG 1
#define TSR_ERR_CUDA 2
#define TSR_ERR_UNSUPPORTED 3
#define TSR_ERR_WORKSPACE 4

void tsr_set_error(const char* fmt, ...);
extern "C" long long tsr_launch_count_inc(int n);   // internal: counts kernel launches

#define TSR_REQUIRE(cond, ...)                                                   \
  do {                                                                           \
    if (!(cond)) {                                                               \
      tsr_set_error(__VA_ARGS__);                                                \
      return TSR_ERR_BAD_ARG;                                                    \
    }                                                                            \
  } while (0)

#define TSR_CHECK_LAUNCH(name)                                                   \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      tsr_set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e__)); \
      return TSR_ERR_CUDA;                                                       \
    }                                                                            \
    tsr_launch_count_inc(1);                                                     \
  } while (0)

#define TSR_CUDA(call)                                                           \
  do {                                                                           \
    cudaError_t e__ = (call);                                                    \
    if (e__ != cudaSuccess) {                                                    \
      tsr_set_error("%s failed: %s", #call, cudaGetErrorString(e__));            \
      return TSR_ERR_CUDA;                                                       \
    }                                                                            \
  } while (0)

static inline int tsr_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// element load/store as float for the two storage types of the activation buffers
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// 4-wide vector access (16 B for fp32, 8 B for bf16)
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x), b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
