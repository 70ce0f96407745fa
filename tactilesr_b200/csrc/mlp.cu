// Small dense layers of tPSFNet's MLP (48-256-1024-256-3, reference model/tPSFNet.py:26-36) and their
// backward.  M = batch rows; the matrices are tiny (0.54 MMAC/sample) so a plain shared-memory tiled
// fp32 GEMM with generic strides serves forward (X W^T), data gradient (dY W) and weight gradient
// (dY^T X).  One CTA per output tile walks K sequentially => deterministic.
#include "common.cuh"

namespace {

// C[m][n] = epilogue( sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn] )
// epilogue: + bias[n]; act 0 none / 1 relu / 2 softplus(beta=1, threshold=20); accumulate adds to C.
__global__ void __launch_bounds__(256)
sgemm_strided_kernel(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ Bm,
                     long long sbk, long long sbn, float* __restrict__ C, long long ldc, int M, int N, int K,
                     const float* __restrict__ bias, int act, int accumulate) {
  __shared__ float As[16][64 + 1];
  __shared__ float Bs[16][64 + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int idx = tid + r * 256;          // 1024 elements per operand tile
      // choose the fastest-varying tile index to follow the operand's unit stride
      int ka, ma;
      if (sak == 1) { ka = idx & 15; ma = idx >> 4; } else { ma = idx & 63; ka = idx >> 6; }
      int gm = m0 + ma, gk = k0 + ka;
      As[ka][ma] = (gm < M && gk < K) ? A[gm * sam + gk * sak] : 0.f;
      int kb, nb;
      if (sbk == 1) { kb = idx & 15; nb = idx >> 4; } else { nb = idx & 63; kb = idx >> 6; }
      int gn = n0 + nb;
      gk = k0 + kb;
      Bs[kb][nb] = (gn < N && gk < K) ? Bm[gk * sbk + gn * sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      if (act == 1) v = fmaxf(v, 0.f);
      else if (act == 2) v = v > 20.f ? v : log1pf(expf(v));
      float* c = C + m * ldc + n;
      *c = accumulate ? *c + v : v;
    }
  }
}

// dpre = dY * act'(pre), expressed through the stored output: relu: [out > 0]; softplus: 1 - exp(-out)
__global__ void act_backward_kernel(const float* __restrict__ dy, const float* __restrict__ out,
                                    float* __restrict__ dpre, long long n, int act) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float g = dy[i], o = out[i];
    if (act == 1) g = o > 0.f ? g : 0.f;
    else if (act == 2) g = o > 20.f ? g : g * (1.f - expf(-o));
    dpre[i] = g;
  }
}

// deterministic column sum of a small [M][N] matrix: one thread per column
__global__ void colsum_small_kernel(const float* __restrict__ x, int M, int N, float* __restrict__ out, int accumulate) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int m = 0; m < M; ++m) s += x[(long long)m * N + n];
  out[n] = accumulate ? out[n] + s : s;
}

}  // namespace

extern "C" {

int tsr_sgemm_strided(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn,
                      float* C, long long ldc, int M, int N, int K, const float* bias, int act, int accumulate,
                      cudaStream_t stream) {
  TSR_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "sgemm_strided: bad argument");
  dim3 grid(tsr_cdiv(N, 64), tsr_cdiv(M, 64));
  sgemm_strided_kernel<<<grid, 256, 0, stream>>>(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, bias, act, accumulate);
  TSR_CHECK_LAUNCH("sgemm_strided");
  return TSR_OK;
}

// y[M][N] = act(x[M][K] W[N][K]^T + b)          (nn.Linear forward)
int tsr_linear_fwd(const float* x, const float* w, const float* b, float* y, int M, int N, int K, int act,
                   cudaStream_t stream) {
  return tsr_sgemm_strided(x, K, 1, w, 1, K, y, N, M, N, K, b, act, 0, stream);
}

// given dy and the stored output `out` of the layer: dpre (scratch [M][N]), dW[N][K] (+)=, db[N] (+)=, dx[M][K] (optional)
int tsr_linear_bwd(const float* dy, const float* out, const float* x, const float* w, float* dpre, float* dw,
                   float* db, float* dx, int M, int N, int K, int act, int accumulate, cudaStream_t stream) {
  TSR_REQUIRE(dy && out && x && w && dpre && dw && db, "linear_bwd: null pointer");
  long long n = (long long)M * N;
  int grid = (int)((n + 255) / 256);
  if (grid > 2048) grid = 2048;
  act_backward_kernel<<<grid, 256, 0, stream>>>(dy, out, dpre, n, act);
  TSR_CHECK_LAUNCH("act_backward");
  // dW[n][k] = sum_m dpre[m][n] x[m][k]:  A(n, m) = dpre[m*N + n], B(m, k) = x[m*K + k]
  int rc = tsr_sgemm_strided(dpre, 1, N, x, K, 1, dw, K, N, K, M, nullptr, 0, accumulate, stream);
  if (rc) return rc;
  colsum_small_kernel<<<tsr_cdiv(N, 128), 128, 0, stream>>>(dpre, M, N, db, accumulate);
  TSR_CHECK_LAUNCH("colsum_small");
  if (dx) {
    // dx[m][k] = sum_n dpre[m][n] w[n][k]
    rc = tsr_sgemm_strided(dpre, N, 1, w, K, 1, dx, K, M, K, N, nullptr, 0, 0, stream);
    if (rc) return rc;
  }
  return TSR_OK;
}

}  // extern "C"
