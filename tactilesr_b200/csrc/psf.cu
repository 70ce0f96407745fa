// tPSFNet point-spread-function forward model, fused per sample (fp32).
//
// Replaces the python `for i in range(B)` loop of reference model/tPSFNet.py:118-125, i.e.
// tactilePSF (:78-83), depth2tactile (:85-100, dense 99x99 F.conv2d + second-max fill) and
// degradation_process (:129-141), and its autograd backward.
//
// Both Gaussians are exactly separable:
//   psf[u][v]   = alpha * e(u) e(v),            e(t)   = exp(-cp2 (t-49)^2 / beta^2),  cp2 = 100/4802
//   conv        = alpha * E D E,                E[m][k] = e(k-m+49) for |k-m| <= 49      (D = depth, 100x100)
//   mask_ij[x,y]= (Ex_i(x) Ey_j(y) - m)/(1-m),  E_k(t) = exp(-cm2 (t-12-25k)^2 / gamma), cm2 = 100/15138,
//                                               m = exp(-100/gamma)
//   LRd[i][j]   = 1e-4 (sum HR Ex_i Ey_j - m sum HR) / (1 - m)
// One CTA owns one whole sample: depth plane, intermediate and HR live in shared memory; global
// traffic is the compulsory 119 KB/sample (depth in; HR, psf, LRd out).
#include "common.cuh"

namespace {

constexpr int N = 100;          // high-res plane
constexpr int PITCH = 101;      // smem pitch (conflict-free transposed stores)
constexpr int PLANE = N * PITCH;
constexpr int TABN = 200;       // padded tap table: index i+50 for i in [-50, 149)
constexpr int NT = 512;         // threads per CTA
constexpr float CP2 = 100.0f / 4802.0f;
constexpr float CM2 = 100.0f / 15138.0f;

__device__ __forceinline__ float block_sum(float v, float* scratch) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < NT / 32; ++i) t += scratch[i];
  return t;
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = scratch[0];
#pragma unroll
  for (int i = 1; i < NT / 32; ++i) t = fmaxf(t, scratch[i]);
  return t;
}

// dstA[n][m] = sum_k tabA[k-m+49] * src[k][n]      (dst^T = KerA * src), and optionally the same with tabB -> dstB.
// Thread item = (group of 10 consecutive m, column n); the 10 taps live in a rotating register window.
template <bool DUAL>
__device__ __forceinline__ void colpass(const float* __restrict__ src, const float* __restrict__ tabA,
                                        const float* __restrict__ tabB, float* __restrict__ dstA,
                                        float* __restrict__ dstB) {
  for (int item = threadIdx.x; item < 10 * N; item += NT) {
    const int mg = item / N, n = item - mg * N;
    const int m0 = mg * 10;
    const int base = 50 + 49 - m0;            // tab index of (k - m0) = 0
    float accA[10], accB[10], cA[10], cB[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) { accA[j] = 0.f; accB[j] = 0.f; }
#pragma unroll
    for (int r = 1; r < 10; ++r) {            // entries for relative positions -9..-1 -> slot r (= (r-10) mod 10)
      cA[r] = tabA[base + r - 10];
      if (DUAL) cB[r] = tabB[base + r - 10];
    }
    cA[0] = 0.f; cB[0] = 0.f;
    for (int k0 = 0; k0 < N; k0 += 10) {
#pragma unroll
      for (int kk = 0; kk < 10; ++kk) {
        const int k = k0 + kk;
        const float s = src[k * PITCH + n];
        cA[kk] = tabA[base + k];
        if (DUAL) cB[kk] = tabB[base + k];
#pragma unroll
        for (int j = 0; j < 10; ++j) {
          accA[j] = fmaf(cA[(kk - j + 10) % 10], s, accA[j]);
          if (DUAL) accB[j] = fmaf(cB[(kk - j + 10) % 10], s, accB[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      dstA[n * PITCH + m0 + j] = accA[j];
      if (DUAL) dstB[n * PITCH + m0 + j] = accB[j];
    }
  }
}

// dst[n][m] = sum_k (tabA[k-m+49]*srcA[k][n] + tabB[k-m+49]*srcB[k][n])
__device__ __forceinline__ void colpass_sum2(const float* __restrict__ srcA, const float* __restrict__ tabA,
                                             const float* __restrict__ srcB, const float* __restrict__ tabB,
                                             float* __restrict__ dst) {
  for (int item = threadIdx.x; item < 10 * N; item += NT) {
    const int mg = item / N, n = item - mg * N;
    const int m0 = mg * 10;
    const int base = 50 + 49 - m0;
    float acc[10], cA[10], cB[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) acc[j] = 0.f;
#pragma unroll
    for (int r = 1; r < 10; ++r) { cA[r] = tabA[base + r - 10]; cB[r] = tabB[base + r - 10]; }
    cA[0] = 0.f; cB[0] = 0.f;
    for (int k0 = 0; k0 < N; k0 += 10) {
#pragma unroll
      for (int kk = 0; kk < 10; ++kk) {
        const int k = k0 + kk;
        const float sa = srcA[k * PITCH + n], sb = srcB[k * PITCH + n];
        cA[kk] = tabA[base + k];
        cB[kk] = tabB[base + k];
#pragma unroll
        for (int j = 0; j < 10; ++j) {
          acc[j] = fmaf(cA[(kk - j + 10) % 10], sa, acc[j]);
          acc[j] = fmaf(cB[(kk - j + 10) % 10], sb, acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 10; ++j) dst[n * PITCH + m0 + j] = acc[j];
  }
}

__device__ __forceinline__ void fill_tabs(float beta, float* tabE, float* tabE2) {
  const float inv_b2 = 1.0f / (beta * beta);
  for (int i = threadIdx.x; i < TABN; i += NT) {
    int t = i - 50;                       // tap index
    float e = 0.f, e2 = 0.f;
    if (t >= 0 && t < 99) {
      float r2 = (float)((t - 49) * (t - 49));
      e = expf(-(CP2 * r2) * inv_b2);
      e2 = e * r2;
    }
    tabE[i] = e;
    if (tabE2) tabE2[i] = e2;
  }
}

// Ex[i][t] = exp(-cm2 (t - 12 - 25 i)^2 / gamma);  Ex2 = Ex * (t - c_i)^2
__device__ __forceinline__ void fill_mask_tabs(float gamma, float* ex, float* ex2) {
  const float inv_g = 1.0f / gamma;
  for (int i = threadIdx.x; i < 4 * N; i += NT) {
    int k = i / N, t = i - k * N;
    float d = (float)(t - 12 - 25 * k);
    float v = expf(-(CM2 * d * d) * inv_g);
    ex[i] = v;
    if (ex2) ex2[i] = v * d * d;
  }
}

__global__ void __launch_bounds__(NT, 1)
psf_fwd_kernel(const float* __restrict__ ab, const float* __restrict__ depth, float* __restrict__ HR,
               float* __restrict__ LRd, float* __restrict__ psf, int B) {
  extern __shared__ float smem[];
  float* P0 = smem;                 // depth
  float* P1 = P0 + PLANE;           // (E D)^T
  float* P2 = P1 + PLANE;           // E D E -> HR
  float* tabE = P2 + PLANE;         // TABN
  float* ex = tabE + TABN;          // 4*N
  float* R = ex + 4 * N;            // 4*N  R[i][col]
  float* colsum = R + 4 * N;        // N
  float* scratch = colsum + N;      // 32

  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    const float alpha = ab[b * 3 + 0], beta = ab[b * 3 + 1], gamma = ab[b * 3 + 2];
    const float* dsrc = depth + (long long)b * N * N;
    float lmax = -INFINITY;
    for (int i = threadIdx.x; i < N * N; i += NT) {
      float v = dsrc[i];
      P0[(i / N) * PITCH + i % N] = v;
      lmax = fmaxf(lmax, v);
    }
    fill_tabs(beta, tabE, nullptr);
    fill_mask_tabs(gamma, ex, nullptr);
    const float dmax = block_max(lmax, scratch);   // (syncs: tables and P0 visible)
    if (psf) {
      float* pdst = psf + (long long)b * 99 * 99;
      for (int i = threadIdx.x; i < 99 * 99; i += NT) pdst[i] = (alpha * tabE[50 + i / 99]) * tabE[50 + i % 99];
    }
    colpass<false>(P0, tabE, nullptr, P1, nullptr);
    __syncthreads();
    colpass<false>(P1, tabE, nullptr, P2, nullptr);
    __syncthreads();
    // second-max fill (tPSFNet.py:95-97): contact pixels <- max over the conv result with contact zeroed
    const float thr = dmax - 1e-3f;
    float m2 = 0.f;   // the zeroed contact pixels take part in the max; the contact set is never empty
    for (int i = threadIdx.x; i < N * N; i += NT) {
      int o = (i / N) * PITCH + i % N;
      float c = alpha * P2[o];
      P2[o] = c;
      if (!(P0[o] > thr)) m2 = fmaxf(m2, c);
    }
    m2 = block_max(m2, scratch);
    float* hdst = HR + (long long)b * N * N;
    for (int i = threadIdx.x; i < N * N; i += NT) {
      int o = (i / N) * PITCH + i % N;
      float h = (P0[o] > thr) ? m2 : P2[o];
      P2[o] = h;
      hdst[i] = h;
    }
    __syncthreads();
    // degradation: R[i][col] = sum_row HR[row][col] Ex_i(row); colsum[col] = sum_row HR[row][col]
    for (int it = threadIdx.x; it < 5 * N; it += NT) {
      int i = it / N, col = it - i * N;
      float s = 0.f;
      if (i < 4) {
        for (int row = 0; row < N; ++row) s = fmaf(P2[row * PITCH + col], ex[i * N + row], s);
        R[i * N + col] = s;
      } else {
        for (int row = 0; row < N; ++row) s += P2[row * PITCH + col];
        colsum[col] = s;
      }
    }
    __syncthreads();
    {
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;   // 16 warps <-> 16 taxels
      const int i = warp >> 2, j = warp & 3;
      float s = 0.f, tot = 0.f;
      for (int col = lane; col < N; col += 32) {
        s = fmaf(R[i * N + col], ex[j * N + col], s);
        tot += colsum[col];
      }
      s = warp_sum(s);
      tot = warp_sum(tot);
      if (lane == 0) {
        float m = expf(-100.0f / gamma);
        LRd[b * 16 + warp] = 1e-4f * (s - m * tot) / (1.0f - m);
      }
    }
  }
}

// backward: d(alpha, beta, gamma) given dLRd (B,16), optional dHR (B,100,100) and dpsf (B,99,99)
__global__ void __launch_bounds__(NT, 1)
psf_bwd_kernel(const float* __restrict__ ab, const float* __restrict__ depth, const float* __restrict__ HR,
               const float* __restrict__ dLRd, const float* __restrict__ dHR, const float* __restrict__ dpsf,
               float* __restrict__ dab, int B) {
  extern __shared__ float smem[];
  float* P0 = smem;                 // depth, then HR
  float* P1 = P0 + PLANE;           // (E D)^T
  float* P2 = P1 + PLANE;           // (E2 D)^T
  float* P3 = P2 + PLANE;           // A = E2 D E + E D E2
  float* tabE = P3 + PLANE;
  float* tabE2 = tabE + TABN;
  float* ex = tabE2 + TABN;         // 4*N
  float* ex2 = ex + 4 * N;          // 4*N
  float* Q = ex2 + 4 * N;           // 4*N   Q_i(col) = sum_j g_ij Ey_j(col)
  float* R = Q + 4 * N;             // 4*N
  float* R2 = R + 4 * N;            // 4*N
  float* colsum = R2 + 4 * N;       // N
  float* scratch = colsum + N;      // 32
  float* gsh = scratch + 32;        // 16
  float* S = gsh + 16;              // 16
  float* S1 = S + 16;               // 16
  uint32_t* maskbits = reinterpret_cast<uint32_t*>(S1 + 16);   // 320 words

  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    const float alpha = ab[b * 3 + 0], beta = ab[b * 3 + 1], gamma = ab[b * 3 + 2];
    const float* dsrc = depth + (long long)b * N * N;
    float lmax = -INFINITY;
    for (int i = threadIdx.x; i < N * N; i += NT) {
      float v = dsrc[i];
      P0[(i / N) * PITCH + i % N] = v;
      lmax = fmaxf(lmax, v);
    }
    fill_tabs(beta, tabE, tabE2);
    fill_mask_tabs(gamma, ex, ex2);
    if (threadIdx.x < 16) gsh[threadIdx.x] = dLRd ? dLRd[b * 16 + threadIdx.x] : 0.f;
    for (int i = threadIdx.x; i < 320; i += NT) maskbits[i] = 0u;
    const float dmax = block_max(lmax, scratch);
    const float thr = dmax - 1e-3f;
    colpass<true>(P0, tabE, tabE2, P1, P2);
    // contact mask bits (one word per 32 flattened pixels), built from the depth plane before it is overwritten
    for (int w = threadIdx.x; w < 313; w += NT) {
      uint32_t bits = 0u;
      for (int k = 0; k < 32; ++k) {
        int i = w * 32 + k;
        if (i < N * N && P0[(i / N) * PITCH + i % N] > thr) bits |= 1u << k;
      }
      maskbits[w] = bits;
    }
    __syncthreads();
    const float* hsrc = HR + (long long)b * N * N;
    for (int i = threadIdx.x; i < N * N; i += NT) P0[(i / N) * PITCH + i % N] = hsrc[i];
    colpass_sum2(P1, tabE2, P2, tabE, P3);
    // Q_i(col) = sum_j g_ij Ey_j(col)
    for (int it = threadIdx.x; it < 4 * N; it += NT) {
      int i = it / N, col = it - i * N;
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) s = fmaf(gsh[i * 4 + j], ex[j * N + col], s);
      Q[it] = s;
    }
    __syncthreads();
    const float m = expf(-100.0f / gamma);
    const float kk = 1e-4f / (1.0f - m);
    float gsum = 0.f;
#pragma unroll
    for (int t = 0; t < 16; ++t) gsum += gsh[t];
    // d alpha, d beta through the non-contact HR pixels
    const float* dh = dHR ? dHR + (long long)b * N * N : nullptr;
    float da = 0.f, db = 0.f;
    for (int i = threadIdx.x; i < N * N; i += NT) {
      if ((maskbits[i >> 5] >> (i & 31)) & 1u) continue;
      int row = i / N, col = i - row * N;
      float w = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) w = fmaf(ex[q * N + row], Q[q * N + col], w);
      w = kk * (w - m * gsum);
      if (dh) w += dh[i];
      int o = row * PITCH + col;
      da = fmaf(w, P0[o], da);
      db = fmaf(w, P3[o], db);
    }
    const float dbeta_scale = alpha * 2.0f * CP2 / (beta * beta * beta);
    if (dpsf) {
      const float* dp = dpsf + (long long)b * 99 * 99;
      float pa = 0.f, pb = 0.f;
      for (int i = threadIdx.x; i < 99 * 99; i += NT) {
        int u = i / 99, v = i - u * 99;
        float ee = tabE[50 + u] * tabE[50 + v];
        float r2 = (float)((u - 49) * (u - 49) + (v - 49) * (v - 49));
        float g = dp[i];
        pa = fmaf(g, ee, pa);
        pb = fmaf(g, ee * r2, pb);
      }
      da = da / alpha + pa;      // (da so far is sum w*HR; /alpha below otherwise)
      db = db + pb;
      da = block_sum(da, scratch);
    } else {
      da = block_sum(da, scratch) / alpha;
    }
    db = block_sum(db, scratch) * dbeta_scale;
    // d gamma: R, R2 (row-weighted column profiles of HR), then the 16 taxel sums
    for (int it = threadIdx.x; it < 9 * N; it += NT) {
      int i = it / N, col = it - i * N;
      float s = 0.f;
      if (i < 4) {
        for (int row = 0; row < N; ++row) s = fmaf(P0[row * PITCH + col], ex[i * N + row], s);
        R[i * N + col] = s;
      } else if (i < 8) {
        for (int row = 0; row < N; ++row) s = fmaf(P0[row * PITCH + col], ex2[(i - 4) * N + row], s);
        R2[(i - 4) * N + col] = s;
      } else {
        for (int row = 0; row < N; ++row) s += P0[row * PITCH + col];
        colsum[col] = s;
      }
    }
    __syncthreads();
    float tot_all = 0.f;
    {
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      const int i = warp >> 2, j = warp & 3;
      float s = 0.f, s1 = 0.f, tot = 0.f;
      for (int col = lane; col < N; col += 32) {
        s = fmaf(R[i * N + col], ex[j * N + col], s);
        s1 = fmaf(R2[i * N + col], ex[j * N + col], s1);
        s1 = fmaf(R[i * N + col], ex2[j * N + col], s1);
        tot += colsum[col];
      }
      s = warp_sum(s);
      s1 = warp_sum(s1);
      tot = warp_sum(tot);
      if (lane == 0) { S[warp] = s; S1[warp] = s1; }
      tot_all = tot;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const float inv_g2 = 1.0f / (gamma * gamma);
      const float mp = m * 100.0f * inv_g2;
      const float om = 1.0f - m;
      float dg = 0.f;
      for (int t = 0; t < 16; ++t) {
        float sp = CM2 * inv_g2 * S1[t];
        float d = 1e-4f * ((sp - mp * tot_all) / om + (S[t] - m * tot_all) * mp / (om * om));
        dg = fmaf(gsh[t], d, dg);
      }
      dab[b * 3 + 0] = da;
      dab[b * 3 + 1] = db;
      dab[b * 3 + 2] = dg;
    }
  }
}

constexpr size_t FWD_SMEM = (size_t)(3 * PLANE + TABN + 4 * N + 4 * N + N + 32) * sizeof(float);
constexpr size_t BWD_SMEM = (size_t)(4 * PLANE + 2 * TABN + 5 * 4 * N + N + 32 + 48 + 320) * sizeof(float);

}  // namespace

extern "C" {

// FFMA version of tsr_psf_forward (the default is the tensor-core kernel of psf_tc.cu; tsr_set_psf_mode(1) selects
// this one): one CTA per sample, register-window separable passes.
int tsr_psf_forward_ffma(const float* alphaBeta, const float* depth, float* HR, float* LRd, float* psf, int B,
                    cudaStream_t stream) {
  TSR_REQUIRE(alphaBeta && depth && HR && LRd && B > 0, "psf_forward: bad argument");
  TSR_CUDA(cudaFuncSetAttribute(psf_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM));
  int grid = B < 148 ? B : 148;
  psf_fwd_kernel<<<grid, NT, FWD_SMEM, stream>>>(alphaBeta, depth, HR, LRd, psf, B);
  TSR_CHECK_LAUNCH("psf_forward");
  return TSR_OK;
}

// d alphaBeta (B,3) from dLRd (B,16) [NULL = 0], dHR (B,100,100) [NULL = 0], dpsf (B,99,99) [NULL = 0].
int tsr_psf_backward(const float* alphaBeta, const float* depth, const float* HR, const float* dLRd,
                     const float* dHR, const float* dpsf, float* dalphaBeta, int B, cudaStream_t stream) {
  TSR_REQUIRE(alphaBeta && depth && HR && dalphaBeta && B > 0, "psf_backward: bad argument");
  TSR_CUDA(cudaFuncSetAttribute(psf_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
  int grid = B < 148 ? B : 148;
  psf_bwd_kernel<<<grid, NT, BWD_SMEM, stream>>>(alphaBeta, depth, HR, dLRd, dHR, dpsf, dalphaBeta, B);
  TSR_CHECK_LAUNCH("psf_backward");
  return TSR_OK;
}

}  // extern "C"
