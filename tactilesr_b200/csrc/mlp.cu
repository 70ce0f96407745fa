// Small dense layers of tPSFNet's MLP (48-256-1024-256-3, reference model/tPSFNet.py:26-36) and their
// backward.  M = batch rows; the matrices are tiny (0.54 MMAC/sample) so a plain shared-memory tiled
// fp32 GEMM with generic strides serves forward (X W^T), data gradient (dY W) and weight gradient
// (dY^T X).  One CTA per output tile walks K sequentially => deterministic.
#include "common.cuh"

namespace {

// C[m][n] = epilogue( sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn] )
// epilogue: + bias[n]; act 0 none / 1 relu / 2 softplus(beta=1, threshold=20); accumulate adds to C.
// 128 x 64 x 16 tiles, 256 threads, 8 x 4 outputs per thread; operands are staged k-major in shared memory so that the
// inner loop is three 16-byte shared loads per 32 FMAs.  gridDim.z > 1 splits K into contiguous ranges whose partial
// products go to C + z * split_stride (no epilogue); sgemm_split_reduce_kernel sums them in a fixed order.
constexpr int TM = 128, TN = 64, TK = 16;

__global__ void __launch_bounds__(256)
sgemm_strided_kernel(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ Bm,
                     long long sbk, long long sbn, float* __restrict__ C, long long ldc, int M, int N, int K,
                     const float* __restrict__ bias, int act, int accumulate, int k_per_split, long long split_stride) {
  __shared__ __align__(16) float As[TK][TM + 4];
  __shared__ __align__(16) float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;          // thread tile: rows ty*8.., columns tx*4..
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int kbeg = blockIdx.z * k_per_split, kend = min(K, kbeg + k_per_split);
  float2 acc2[8][2];                               // column pairs: the inner product runs as FFMA2
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc2[i][j] = make_float2(0.f, 0.f);
  for (int k0 = kbeg; k0 < kend; k0 += TK) {
#pragma unroll
    for (int r = 0; r < TM * TK / 256; ++r) {
      const int idx = tid + r * 256;
      // the fastest-varying tile index follows the operand's unit stride
      int ka, ma;
      if (sak == 1) { ka = idx & (TK - 1); ma = idx / TK; } else { ma = idx & (TM - 1); ka = idx / TM; }
      const int gm = m0 + ma, gk = k0 + ka;
      As[ka][ma] = (gm < M && gk < kend) ? A[gm * sam + gk * sak] : 0.f;
    }
#pragma unroll
    for (int r = 0; r < TN * TK / 256; ++r) {
      const int idx = tid + r * 256;
      int kb, nb;
      if (sbk == 1) { kb = idx & (TK - 1); nb = idx / TK; } else { nb = idx & (TN - 1); kb = idx / TN; }
      const int gn = n0 + nb, gk = k0 + kb;
      Bs[kb][nb] = (gn < N && gk < kend) ? Bm[gk * sbk + gn * sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float2 bb[2] = {make_float2(b4.x, b4.y), make_float2(b4.z, b4.w)};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) ffma2s(acc2[i][j], a[i], bb[j]);
    }
    __syncthreads();
  }
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i][0] = acc2[i][0].x; acc[i][1] = acc2[i][0].y; acc[i][2] = acc2[i][1].x; acc[i][3] = acc2[i][1].y;
  }
  float* Cz = C + (long long)blockIdx.z * split_stride;
  const bool partial = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      float* c = Cz + m * ldc + n;
      if (partial) { *c = v; continue; }
      if (bias) v += bias[n];
      if (act == 1) v = fmaxf(v, 0.f);
      else if (act == 2) v = v > 20.f ? v : log1pf(expf(v));
      *c = accumulate ? *c + v : v;
    }
  }
}

// C[i] (+)= sum_z partial[z][i], fixed order
__global__ void sgemm_split_reduce_kernel(const float* __restrict__ partial, int S, long long n, float* __restrict__ C,
                                          int accumulate) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < S; ++z) s += partial[(long long)z * n + i];
    C[i] = accumulate ? C[i] + s : s;
  }
}

// deterministic column sums of an [M][N] matrix: one block per column, fixed-order tree
__global__ void __launch_bounds__(256)
colsum_block_kernel(const float* __restrict__ x, int M, int N, float* __restrict__ out, int accumulate) {
  __shared__ float sh[256];
  const int n = blockIdx.x;
  float s = 0.f;
  for (int m = threadIdx.x; m < M; m += 256) s += x[(long long)m * N + n];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = accumulate ? out[n] + sh[0] : sh[0];
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

// ---- thin layers (N <= 4 outputs: the 256 -> 3 softplus head of tPSFNet, reference model/tPSFNet.py:33-34) -------------------
// A 128 x 64 GEMM tile would compute 61 dead columns; these kernels keep the N weight rows in registers instead.
constexpr int THIN_MAX_N = 4;

// y[m][n] = act(sum_k x[m][k] w[n][k] + b[n]): one warp per row, lane l owns k = 4 l + 128 j (K % 128 == 0, K <= 512)
template <int KJ>
__global__ void __launch_bounds__(256)
thin_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ y,
                int M, int N, int act) {
  const int lane = threadIdx.x & 31;
  const int K = KJ * 128;
  float4 wr[THIN_MAX_N][KJ];
#pragma unroll
  for (int n = 0; n < THIN_MAX_N; ++n)
#pragma unroll
    for (int j = 0; j < KJ; ++j)
      wr[n][j] = n < N ? *reinterpret_cast<const float4*>(w + (size_t)n * K + j * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  const int warps = gridDim.x * (blockDim.x >> 5);
  for (int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); m < M; m += warps) {
    float s[THIN_MAX_N] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < KJ; ++j) {
      const float4 v = *reinterpret_cast<const float4*>(x + (size_t)m * K + j * 128 + lane * 4);
#pragma unroll
      for (int n = 0; n < THIN_MAX_N; ++n)
        s[n] = fmaf(v.x, wr[n][j].x, fmaf(v.y, wr[n][j].y, fmaf(v.z, wr[n][j].z, fmaf(v.w, wr[n][j].w, s[n]))));
    }
#pragma unroll
    for (int n = 0; n < THIN_MAX_N; ++n) s[n] = warp_sum(s[n]);
    if (lane < N) {
      float v = (lane == 0 ? s[0] : lane == 1 ? s[1] : lane == 2 ? s[2] : s[3]) + (b ? b[lane] : 0.f);
      if (act == 1) v = fmaxf(v, 0.f);
      else if (act == 2) v = v > 20.f ? v : log1pf(expf(v));
      y[(size_t)m * N + lane] = v;
    }
  }
}

// dx[m][k] = sum_n dpre[m][n] w[n][k]: one thread per (m, 4 k)
__global__ void thin_dx_kernel(const float* __restrict__ dpre, const float* __restrict__ w, float* __restrict__ dx, int M, int N,
                               int K) {
  const int k4 = K >> 2;
  const long long total = (long long)M * k4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % k4);
    const long long m = i / k4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int n = 0; n < N; ++n) {
      const float g = dpre[m * N + n];
      const float4 wv = *reinterpret_cast<const float4*>(w + (size_t)n * K + q * 4);
      a.x = fmaf(g, wv.x, a.x); a.y = fmaf(g, wv.y, a.y); a.z = fmaf(g, wv.z, a.z); a.w = fmaf(g, wv.w, a.w);
    }
    *reinterpret_cast<float4*>(dx + m * K + q * 4) = a;
  }
}

// partial[slab][n][k] = sum_{m in slab} dpre[m][n] x[m][k]: block = one slab of rows, thread = k (fixed order => deterministic)
__global__ void __launch_bounds__(256)
thin_dw_kernel(const float* __restrict__ dpre, const float* __restrict__ x, float* __restrict__ partial, int M, int N, int K,
               int rows_per_slab) {
  const int m0 = blockIdx.x * rows_per_slab, m1 = min(M, m0 + rows_per_slab);
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float a[THIN_MAX_N] = {0.f, 0.f, 0.f, 0.f};
    for (int m = m0; m < m1; ++m) {
      const float xv = x[(size_t)m * K + k];
#pragma unroll
      for (int n = 0; n < THIN_MAX_N; ++n)
        if (n < N) a[n] = fmaf(dpre[(size_t)m * N + n], xv, a[n]);
    }
#pragma unroll
    for (int n = 0; n < THIN_MAX_N; ++n)
      if (n < N) partial[((size_t)blockIdx.x * N + n) * K + k] = a[n];
  }
}

constexpr int THIN_SLAB = 32;
inline bool thin_shape(int N, int K) { return N <= THIN_MAX_N && K % 128 == 0 && K <= 512; }
inline int thin_slabs(int M) { return tsr_cdiv(M, THIN_SLAB); }

}  // namespace

extern "C" {

int tsr_sgemm_strided(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn,
                      float* C, long long ldc, int M, int N, int K, const float* bias, int act, int accumulate,
                      cudaStream_t stream) {
  TSR_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "sgemm_strided: bad argument");
  dim3 grid(tsr_cdiv(N, TN), tsr_cdiv(M, TM));
  sgemm_strided_kernel<<<grid, 256, 0, stream>>>(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, bias, act, accumulate, K, 0);
  TSR_CHECK_LAUNCH("sgemm_strided");
  return TSR_OK;
}

// y[M][N] = act(x[M][K] W[N][K]^T + b)          (nn.Linear forward)
int tsr_linear_fwd(const float* x, const float* w, const float* b, float* y, int M, int N, int K, int act,
                   cudaStream_t stream) {
  if (thin_shape(N, K)) {
    TSR_REQUIRE(x && w && y && M > 0, "linear_fwd: bad argument");
    int grid = tsr_cdiv(M, 8);
    if (grid > 4 * 148) grid = 4 * 148;
    switch (K / 128) {
      case 1: thin_fwd_kernel<1><<<grid, 256, 0, stream>>>(x, w, b, y, M, N, act); break;
      case 2: thin_fwd_kernel<2><<<grid, 256, 0, stream>>>(x, w, b, y, M, N, act); break;
      case 3: thin_fwd_kernel<3><<<grid, 256, 0, stream>>>(x, w, b, y, M, N, act); break;
      default: thin_fwd_kernel<4><<<grid, 256, 0, stream>>>(x, w, b, y, M, N, act); break;
    }
    TSR_CHECK_LAUNCH("thin_fwd");
    return TSR_OK;
  }
  return tsr_sgemm_strided(x, K, 1, w, 1, K, y, N, M, N, K, b, act, 0, stream);
}

// splits of the batch dimension for the weight gradient: enough CTAs to fill the GPU, at least 256 rows each
static int linear_wgrad_splits(int M, int N, int K) {
  int tiles = tsr_cdiv(K, TN) * tsr_cdiv(N, TM);
  int s = (2 * 148 + tiles - 1) / tiles;
  int maxs = M / 256;
  if (s > maxs) s = maxs;
  if (s > 64) s = 64;
  return s < 1 ? 1 : s;
}

size_t tsr_linear_bwd_workspace(int M, int N, int K) {
  if (thin_shape(N, K)) return (size_t)thin_slabs(M) * N * K * sizeof(float);
  int s = linear_wgrad_splits(M, N, K);
  return s > 1 ? (size_t)s * N * K * sizeof(float) : 0;
}

// given dy and the stored output `out` of the layer: dpre (scratch [M][N]), dW[N][K] (+)=, db[N] (+)=, dx[M][K] (optional).
// workspace: tsr_linear_bwd_workspace(M, N, K) bytes (split-batch partial weight gradients, summed in a fixed order).
int tsr_linear_bwd(const float* dy, const float* out, const float* x, const float* w, float* dpre, float* dw,
                   float* db, float* dx, int M, int N, int K, int act, int accumulate, void* workspace, size_t ws_bytes,
                   cudaStream_t stream) {
  TSR_REQUIRE(dy && out && x && w && dpre && dw && db, "linear_bwd: null pointer");
  long long n = (long long)M * N;
  int grid = (int)((n + 255) / 256);
  if (grid > 2048) grid = 2048;
  act_backward_kernel<<<grid, 256, 0, stream>>>(dy, out, dpre, n, act);
  TSR_CHECK_LAUNCH("act_backward");
  if (thin_shape(N, K)) {
    const int S = thin_slabs(M);
    TSR_REQUIRE(workspace && ws_bytes >= (size_t)S * N * K * sizeof(float), "linear_bwd: workspace too small");
    thin_dw_kernel<<<S, 256, 0, stream>>>(dpre, x, (float*)workspace, M, N, K, THIN_SLAB);
    TSR_CHECK_LAUNCH("thin_dw");
    long long nk = (long long)N * K;
    sgemm_split_reduce_kernel<<<(int)((nk + 255) / 256), 256, 0, stream>>>((const float*)workspace, S, nk, dw, accumulate);
    TSR_CHECK_LAUNCH("linear_wgrad_reduce");
    colsum_block_kernel<<<N, 256, 0, stream>>>(dpre, M, N, db, accumulate);
    TSR_CHECK_LAUNCH("colsum_block");
    if (dx) {
      long long t = (long long)M * (K / 4);
      int g = (int)((t + 255) / 256);
      if (g > 8 * 148) g = 8 * 148;
      thin_dx_kernel<<<g, 256, 0, stream>>>(dpre, w, dx, M, N, K);
      TSR_CHECK_LAUNCH("thin_dx");
    }
    return TSR_OK;
  }
  // dW[n][k] = sum_m dpre[m][n] x[m][k]:  A(n, m) = dpre[m*N + n], B(m, k) = x[m*K + k]
  const int S = linear_wgrad_splits(M, N, K);
  if (S > 1) {
    TSR_REQUIRE(workspace && ws_bytes >= (size_t)S * N * K * sizeof(float), "linear_bwd: workspace too small");
    const int kps = tsr_cdiv(tsr_cdiv(M, S), TK) * TK;
    dim3 g3(tsr_cdiv(K, TN), tsr_cdiv(N, TM), tsr_cdiv(M, kps));
    sgemm_strided_kernel<<<g3, 256, 0, stream>>>(dpre, 1, N, x, K, 1, (float*)workspace, K, N, K, M, nullptr, 0, 0, kps,
                                                 (long long)N * K);
    TSR_CHECK_LAUNCH("linear_wgrad_split");
    long long nk = (long long)N * K;
    sgemm_split_reduce_kernel<<<(int)((nk + 255) / 256), 256, 0, stream>>>((const float*)workspace, (int)g3.z, nk, dw, accumulate);
    TSR_CHECK_LAUNCH("linear_wgrad_reduce");
  } else {
    int rc = tsr_sgemm_strided(dpre, 1, N, x, K, 1, dw, K, N, K, M, nullptr, 0, accumulate, stream);
    if (rc) return rc;
  }
  colsum_block_kernel<<<N, 256, 0, stream>>>(dpre, M, N, db, accumulate);
  TSR_CHECK_LAUNCH("colsum_block");
  if (dx) {
    // dx[m][k] = sum_n dpre[m][n] w[n][k]
    int rc = tsr_sgemm_strided(dpre, N, 1, w, K, 1, dx, K, M, K, N, nullptr, 0, 0, stream);
    if (rc) return rc;
  }
  return TSR_OK;
}

}  // extern "C"
