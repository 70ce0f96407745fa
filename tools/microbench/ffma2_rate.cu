// FFMA vs FFMA2 issue rate on sm_100a (one B200): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(float2& d, float2 a, float2 b) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
}
template <int MODE>   // 0: scalar FFMA 3-reg, 1: FFMA2 vector x vector, 2: FFMA2 scalar-broadcast x vector
__global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
  float2 acc[16];
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  float2 b0 = make_float2(s, s * 0.5f), b1 = make_float2(s * 0.25f, s * 0.125f);
  float a0 = s * 1.0001f, a1 = s * 0.9999f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (MODE == 0) {
          asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i].x) : "f"(a0), "f"(b0.x));
          asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i].y) : "f"(a1), "f"(b1.y));
        } else if (MODE == 1) {
          ffma2(acc[i], (i & 1) ? b0 : b1, (i & 2) ? b1 : b0);
        } else {
          ffma2(acc[i], make_float2((i & 1) ? a0 : a1, (i & 1) ? a0 : a1), (i & 2) ? b1 : b0);
        }
      }
    }
  }
  float r = 0.f;
  for (int i = 0; i < 16; ++i) r += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, float* out, int ctas_per_sm) {
  const int iters = 4096, grid = 148 * ctas_per_sm;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, 256>>>(out, 16, 1.0f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(out, iters, 1.0f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma = (double)grid * 256 * iters * 4 * 16 * 2;   // fp32 FMAs
  printf("%-28s %d CTAs/SM: %.3f ms  %.1f TFLOP/s  (%.1f FMA/clk/SM at 1.9 GHz)\n", name, ctas_per_sm, ms, 2 * fma / ms / 1e9,
         fma / (ms * 1e-3) / 148 / 1.9e9);
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  for (int c : {2, 4}) {
    run<0>("FFMA (3-reg scalar)", out, c);
    run<1>("FFMA2 (vector x vector)", out, c);
    run<2>("FFMA2 (scalar bcast x vector)", out, c);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
