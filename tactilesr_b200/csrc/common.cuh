// Shared helpers for the tactilesr_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
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
int* tsr_f16_overflow_ptr();                        // internal: device flag of the fp16 overflow guard, or null

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

// Two independent IEEE fp32 FMAs in one instruction (FFMA2): d.x = a.x b.x + d.x, d.y = a.y b.y + d.y.  On sm_100 a scalar
// three-register FFMA issues every other cycle per scheduler, FFMA2 at the same rate: the packed form is what reaches the
// fp32 peak.  Each half rounds exactly like fmaf, so kernels converted to it stay bit-identical.  ptxas folds a pair built
// as (s, s) into the instruction's scalar-broadcast operand form.
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
}
__device__ __forceinline__ void ffma2s(float2& d, const float s, const float2 b) { ffma2(d, make_float2(s, s), b); }
__device__ __forceinline__ float2 fmul2s(const float2 a, const float s) {          // (a.x s, a.y s) in one FMUL2
  float2 d;
  const float2 ss = make_float2(s, s);
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(ss)));
  return d;
}

// 16-byte asynchronous copy global -> shared without staging registers (LDGSTS); !valid writes zeros (src-size 0: nothing
// is read, the pointer only has to be a mapped address)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// element load/store as float for the two storage types of the activation buffers
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float ldf(const __half* p) { return __half2float(*p); }
__device__ __forceinline__ void stf(__half* p, float v) { *p = __float2half_rn(v); }

// storage-type codes of the `*_bf16` / dtype arguments of the C ABI
#define TSR_DT_F32 0
#define TSR_DT_BF16 1
#define TSR_DT_F16 2

// run BODY with T bound to the storage type selected by CODE
#define TSR_DISPATCH_T(CODE, T, ...)                              \
  do {                                                            \
    if ((CODE) == TSR_DT_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
    else if ((CODE) == TSR_DT_F16) { using T = __half; __VA_ARGS__; }    \
    else { using T = float; __VA_ARGS__; }                        \
  } while (0)

// 4-wide vector access (16 B for fp32, 8 B for bf16 / fp16)
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

__device__ __forceinline__ float4 ld4(const __half* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __half2 a = *reinterpret_cast<__half2*>(&u.x), b = *reinterpret_cast<__half2*>(&u.y);
  float2 fa = __half22float2(a), fb = __half22float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(__half* p, float4 v) {
  __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// 16-byte vector access of an activation row: 4 floats or 8 bf16, always presented as floats
template <typename T> struct V16;
template <> struct V16<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void load(const float* p, float (&f)[4]) {
    float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  __device__ static __forceinline__ void store(float* p, const float (&f)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};
template <> struct V16<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
      float2 t = __bfloat1622float2(h);
      f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
  }
  __device__ static __forceinline__ void store(__nv_bfloat16* p, const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <> struct V16<__half> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void load(const __half* p, float (&f)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      float2 t = __half22float2(h);
      f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
  }
  __device__ static __forceinline__ void store(__half* p, const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};
