// tPSFNet point-spread-function forward model on the tensor cores (tcgen05 / TMEM), fp32-accurate.
//
// Same contract as psf_fwd_kernel (psf.cu): replaces the python per-sample loop of reference model/tPSFNet.py:118-125
// (tactilePSF :78-83, depth2tactile :85-100, degradation_process :129-141).  The 99x99 correlation is separable
// (SURVEY.md Appendix B):  conv = alpha * E D E,  E[m][k] = e(|k-m|) for |k-m| <= 49 else 0,  e(t) = exp(-cp2 t^2 / beta^2)
// (100x100, symmetric banded Toeplitz, different for every sample because beta is), i.e. two dense 100^3 contractions per sample = 4 MFLOP against
// 119 KB of compulsory HBM traffic: on FFMA pipes that is 3x over the HBM time, so the two products run as tcgen05.mma.
//
// fp32 accuracy from 16-bit operands: every operand is split x = hi + lo into two fp16 numbers (power-of-two pre-scaling
// keeps both halves in fp16's normal range) and each product is three MMAs  hi*hi + lo*hi + hi*lo  accumulated in fp32
// in TMEM (the dropped lo*lo term is 2^-22 relative).  Measured against the fp64 reference run: ~1e-6 rel-L2.
//
// Per sample (one CTA, 256 threads; two CTAs per SM overlap each other's phases):
//   A  tables e(t), Ex_i(t);  depth max (contact threshold);  E -> smem (K-major SWIZZLE_128B, hi and lo tiles);
//      depth^T -> smem (same layout; 4-byte transposing loads, 16-byte swizzled stores) + contact-mask bytes
//   B  GEMM1  T = E * D      (M=128, N=112, K=7x16; 21 MMAs)      -> TMEM columns [0,112)
//   X  T: TMEM -> registers -> hi/lo fp16 -> smem, over the dead depth tiles (all 8 warps: 4 lane quarters x 2 halves)
//   C  GEMM2  HR = T * E     (E symmetric: the same E tiles are the B operand)   -> TMEM columns [128,240)
//      (the psf output, a pure function of alpha / beta, is written to HBM while GEMM2 runs)
//   E  epilogue: second-max fill (tPSFNet.py:95-97), HR store, LRd = 1e-4 (Ex HR Ex^T - m sum HR) / (1 - m)
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int N = 100;
constexpr int NT = 256;
constexpr float CP2 = 100.0f / 4802.0f;
constexpr float CM2 = 100.0f / 15138.0f;

constexpr uint32_t ROWS = 104;                 // allocated rows of a tile (13 groups of 8); MMAs over-read up to row 127
constexpr uint32_t ATOM = ROWS * 128u;         // one 64-wide K atom: rows x 128 B, SWIZZLE_128B
constexpr uint32_t TILE = 2u * ATOM;           // K = 112 = 64 + 48
constexpr uint32_t OFF_E_HI = 0, OFF_E_LO = TILE, OFF_X_HI = 2 * TILE, OFF_X_LO = 3 * TILE;
constexpr uint32_t OFF_TAB = 4 * TILE;                       // float e(t), t = 0..99  (+ pad)
constexpr uint32_t OFF_TAB2 = OFF_TAB + 128 * 4;             // uint32 (hi | lo << 16) of 2^14 e(|j - 99|), j = 0..198 (+ pad)
constexpr uint32_t OFF_EX = OFF_TAB2 + 208 * 4;              // float Ex_i(t), 4 x 100
constexpr uint32_t OFF_MASK = OFF_EX + 400 * 4;              // contact bytes [13][112]: bit j of [cg][n] <-> depth[8 cg + j][n]
constexpr uint32_t OFF_RED = OFF_MASK + 13 * 112;            // float scratch [8][20]
constexpr uint32_t OFF_BAR = (OFF_RED + 8 * 20 * 4 + 15u) & ~15u;   // 2 mbarriers + tmem slot
constexpr uint32_t SMEM_USED = OFF_BAR + 32;
// the last A tile (X_lo) is over-read by (128 - 104) rows = 3 KB: the tables behind it cover that
static_assert(SMEM_USED - 4 * TILE >= 3072, "the tables must cover the over-read of the last tile");
constexpr size_t SMEM_BYTES = 1024 + ((SMEM_USED + 15) & ~15u);
static_assert(2 * (SMEM_BYTES + 1024) <= 228 * 1024, "two CTAs per SM");

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

// byte offset of the 16-byte chunk holding k = 8 cg .. 8 cg + 7 of row r in a K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t chunk_off(int r, int cg) {
  return (uint32_t)(cg >> 3) * ATOM + (uint32_t)r * 128u + (uint32_t)(((cg & 7) ^ (r & 7)) << 4);
}

__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// x (already scaled) -> fp16 hi, fp16 lo with hi + lo = x to ~22 bits
__device__ __forceinline__ void split_h(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}

__device__ __forceinline__ float block_max256(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = red[0];
#pragma unroll
  for (int i = 1; i < NT / 32; ++i) t = fmaxf(t, red[i]);
  return t;
}

// one GEMM: D[128 x 112] (TMEM) = sum of three hi/lo products of K-major tiles A (rows = M) and B (rows = N)
__device__ __forceinline__ void issue_gemm3(uint32_t tmem_d, uint32_t a_hi_addr, uint32_t a_lo_addr, uint32_t b_hi_addr,
                                            uint32_t b_lo_addr, uint32_t idesc) {
  const uint32_t hi_word = desc_hi(1024u);
  uint32_t acc = 0u;
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const uint32_t a0 = pass == 1 ? a_lo_addr : a_hi_addr;
    const uint32_t b0 = pass == 2 ? b_lo_addr : b_hi_addr;
#pragma unroll
    for (int ks = 0; ks < 7; ++ks) {
      const uint32_t off = (uint32_t)(ks >> 2) * ATOM + (uint32_t)(ks & 3) * 32u;
      umma_f16(tmem_d, desc_join(desc_lo(a0 + off, 16u), hi_word), desc_join(desc_lo(b0 + off, 16u), hi_word), idesc, acc);
      acc = 1u;
    }
  }
}

__global__ void __launch_bounds__(NT, 2)
psf_fwd_tc_kernel(const float* __restrict__ ab, const float* __restrict__ depth, float* __restrict__ HR,
                  float* __restrict__ LRd, float* __restrict__ psf, int B) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  float* tab = reinterpret_cast<float*>(sm + OFF_TAB);
  uint32_t* tab2 = reinterpret_cast<uint32_t*>(sm + OFF_TAB2);
  float* ex = reinterpret_cast<float*>(sm + OFF_EX);
  uint8_t* maskb = sm + OFF_MASK;
  float* red = reinterpret_cast<float*>(sm + OFF_RED);
  const uint32_t bar1 = base + OFF_BAR, bar2 = bar1 + 8u, tmem_slot = bar1 + 16u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + OFF_BAR + 16);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // one-time: zero all four tiles (K padding columns, rows 100..103), barriers, TMEM
  for (uint32_t i = tid; i < 4 * TILE / 16; i += NT) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    mbar_init(bar1, 1);
    mbar_init(bar2, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t idesc = make_idesc(128, 112, 0, 0, 0, 0);    // fp16 x fp16 -> fp32, both K-major

  // epilogue geometry: TMEM lane quarter q = warp % 4, row m = 32 q + lane; column half h = warp / 4 (56 columns each)
  const int q = warp & 3, half = warp >> 2;
  const int m = q * 32 + lane;
  const uint32_t lane_addr = (uint32_t)(q * 32) << 16;

  int it = 0;
  for (int b = blockIdx.x; b < B; b += gridDim.x, ++it) {
    const uint32_t ph = (uint32_t)(it & 1);
    const float alpha = ab[b * 3 + 0], beta = ab[b * 3 + 1], gamma = ab[b * 3 + 2];
    const float* dsrc = depth + (size_t)b * N * N;

    // ---- phase A1: tables, depth max / abs-max ----
    {
      const float inv_b2 = 1.0f / (beta * beta);
      if (tid < N) tab[tid] = expf(-(CP2 * (float)(tid * tid)) * inv_b2);
      const float inv_g = 1.0f / gamma;
      for (int i = tid; i < 4 * N; i += NT) {
        const int k = i / N, t = i - k * N;
        const float d = (float)(t - 12 - 25 * k);
        ex[i] = expf(-(CM2 * d * d) * inv_g);
      }
    }
    float lmax = -INFINITY, lamax = 0.f;
    for (int i = tid; i < N * N / 4; i += NT) {
      const float4 v = reinterpret_cast<const float4*>(dsrc)[i];
      lmax = fmaxf(fmaxf(lmax, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
      lamax = fmaxf(fmaxf(lamax, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    const float dmax = block_max256(lmax, red);        // (syncs: tables visible)
    const float amax = block_max256(lamax, red);
    const float thr = dmax - 1e-3f;
    int dexp = 0;
    if (amax > 0.f) (void)frexpf(amax, &dexp);         // amax = f * 2^dexp, f in [0.5, 1)
    const float sD = ldexpf(1.0f, 14 - dexp);          // |depth| * sD < 2^14
    // packed (hi, lo) of 2^14 e(|j - 99|)
    if (tid < 199) {
      const int t = tid < 99 ? 99 - tid : tid - 99;
      __half h, l;
      split_h(t <= 49 ? tab[t] * 16384.0f : 0.f, h, l);      // the PSF has 99 taps: E is banded, |k - m| <= 49
      tab2[tid] = pack_h2(h, l);
    }
    __syncthreads();

    // ---- phase A2: E tiles, transposed depth tiles + contact bytes ----
    for (int item = tid; item < 13 * N; item += NT) {
      const int cg = item / N, r = item - cg * N;      // row r (= m of E, = n of depth^T), k = 8 cg .. 8 cg + 7
      const int k0 = cg * 8;
      const int kv = cg == 12 ? 4 : 8;                 // valid k (k < 100)
      // E[r][k] = e(|k - r|)
      uint32_t eh[4], el[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t p0 = (2 * j < kv) ? tab2[k0 + 2 * j - r + 99] : 0u;
        const uint32_t p1 = (2 * j + 1 < kv) ? tab2[k0 + 2 * j + 1 - r + 99] : 0u;
        eh[j] = (p0 & 0xFFFFu) | (p1 << 16);
        el[j] = (p0 >> 16) | (p1 & 0xFFFF0000u);
      }
      const uint32_t off = chunk_off(r, cg);
      *reinterpret_cast<uint4*>(sm + OFF_E_HI + off) = make_uint4(eh[0], eh[1], eh[2], eh[3]);
      *reinterpret_cast<uint4*>(sm + OFF_E_LO + off) = make_uint4(el[0], el[1], el[2], el[3]);
      // depth^T[n = r][k]: lanes are consecutive n => every k is one coalesced 4-byte load per warp
      float dv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) dv[j] = j < kv ? dsrc[(k0 + j) * N + r] : 0.f;
      uint32_t dh[4], dl[4], bits = 0u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __half h0, l0, h1, l1;
        split_h(dv[2 * j] * sD, h0, l0);
        split_h(dv[2 * j + 1] * sD, h1, l1);
        dh[j] = pack_h2(h0, h1);
        dl[j] = pack_h2(l0, l1);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < kv && dv[j] > thr) bits |= 1u << j;
      *reinterpret_cast<uint4*>(sm + OFF_X_HI + off) = make_uint4(dh[0], dh[1], dh[2], dh[3]);
      *reinterpret_cast<uint4*>(sm + OFF_X_LO + off) = make_uint4(dl[0], dl[1], dl[2], dl[3]);
      maskb[cg * 112 + r] = (uint8_t)bits;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();

    // ---- phase B: GEMM1  T = E * D ----
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        issue_gemm3(tmem_base, base + OFF_E_HI, base + OFF_E_LO, base + OFF_X_HI, base + OFF_X_LO, idesc);
        umma_commit(bar1);
      }
      __syncwarp();
    }
    mbar_wait(bar1, ph);
    tc_fence_after();

    // ---- phase X: T -> hi/lo fp16 -> smem (over the depth tiles, which GEMM1 has finished reading) ----
    // accumulator = 2^14 sD T;  written as 2^(8 - dexp) T  (|T| <= 100 |depth|max  =>  < 2^15)
    {
      const float sx = 1.0f / 1048576.0f;              // 2^-20 = 2^(8 - dexp) / (2^14 sD)
#pragma unroll 1
      for (int g = 0; g < 7; ++g) {
        const int cg = half * 7 + g;
        uint32_t v[8];
        tmem_ld8(tmem_base + lane_addr + (uint32_t)(cg * 8), v);
        tmem_ld_wait();
        if (m < (int)ROWS) {
          uint32_t th[4], tl[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const bool ok0 = cg * 8 + 2 * j < N, ok1 = cg * 8 + 2 * j + 1 < N;     // columns >= 100 are padding: force 0
            __half h0, l0, h1, l1;
            split_h(ok0 ? __uint_as_float(v[2 * j]) * sx : 0.f, h0, l0);
            split_h(ok1 ? __uint_as_float(v[2 * j + 1]) * sx : 0.f, h1, l1);
            th[j] = pack_h2(h0, h1);
            tl[j] = pack_h2(l0, l1);
          }
          const uint32_t off = chunk_off(m, cg);
          *reinterpret_cast<uint4*>(sm + OFF_X_HI + off) = make_uint4(th[0], th[1], th[2], th[3]);
          *reinterpret_cast<uint4*>(sm + OFF_X_LO + off) = make_uint4(tl[0], tl[1], tl[2], tl[3]);
        }
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();

    // ---- phase C: GEMM2  HR = T * E  (E symmetric: its K-major tile is also the B operand) ----
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        issue_gemm3(tmem_base + 128u, base + OFF_X_HI, base + OFF_X_LO, base + OFF_E_HI, base + OFF_E_LO, idesc);
        umma_commit(bar2);
      }
      __syncwarp();
    }
    // psf = alpha e(u) e(v)  (tPSFNet.py:83), written while GEMM2 runs
    if (psf) {
      float* pdst = psf + (size_t)b * 99 * 99;
      for (int i = tid; i < 99 * 99; i += NT) {
        const int u = i / 99, v = i - u * 99;
        pdst[i] = alpha * (tab[u < 49 ? 49 - u : u - 49] * tab[v < 49 ? 49 - v : v - 49]);
      }
    }
    mbar_wait(bar2, ph);
    tc_fence_after();

    // ---- phase E: epilogue.  accumulator = 2^(8 - dexp) 2^14 (E D E) ----
    const float cs = alpha * ldexpf(1.0f, dexp - 22);
    const uint32_t acc2 = tmem_base + 128u + lane_addr;
    // pass 1: second max = max over the conv result with the contact pixels zeroed (tPSFNet.py:95-97)
    float m2 = 0.f;
#pragma unroll 1
    for (int g = 0; g < 7; ++g) {
      const int cg = half * 7 + g;
      uint32_t v[8];
      tmem_ld8(acc2 + (uint32_t)(cg * 8), v);
      tmem_ld_wait();
      if (m < N) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int n = cg * 8 + j;
          if (n < N && !((maskb[(m >> 3) * 112 + n] >> (m & 7)) & 1)) m2 = fmaxf(m2, __uint_as_float(v[j]) * cs);
        }
      }
    }
    m2 = block_max256(m2, red);
    // pass 2: fill, store HR, accumulate the degradation sums of this thread's row segment
    float rj[4] = {0.f, 0.f, 0.f, 0.f}, rs = 0.f;
    float* hdst = HR + (size_t)b * N * N + (size_t)m * N;
#pragma unroll 1
    for (int g = 0; g < 7; ++g) {
      const int cg = half * 7 + g;
      uint32_t v[8];
      tmem_ld8(acc2 + (uint32_t)(cg * 8), v);
      tmem_ld_wait();
      if (m < N && cg * 8 < N) {
        float h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int n = cg * 8 + j;
          const bool contact = n < N && ((maskb[(m >> 3) * 112 + n] >> (m & 7)) & 1);
          h[j] = n < N ? (contact ? m2 : __uint_as_float(v[j]) * cs) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int n = cg * 8 + j;
          if (n < N) {
            rs += h[j];
#pragma unroll
            for (int i = 0; i < 4; ++i) rj[i] = fmaf(h[j], ex[i * N + n], rj[i]);
          }
        }
        *reinterpret_cast<float4*>(hdst + cg * 8) = make_float4(h[0], h[1], h[2], h[3]);
        if (cg * 8 + 4 < N) *reinterpret_cast<float4*>(hdst + cg * 8 + 4) = make_float4(h[4], h[5], h[6], h[7]);
      }
    }
    // LRd[i][j] = 1e-4 (sum_m Ex_i(m) R_j(m) - mm sum HR) / (1 - mm),  R_j(m) = sum_n HR[m][n] Ex_j(n)
    {
      float p[17];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float e = m < N ? ex[i * N + m] : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) p[i * 4 + j] = e * rj[j];
      }
      p[16] = m < N ? rs : 0.f;
#pragma unroll
      for (int k = 0; k < 17; ++k) p[k] = warp_sum(p[k]);
      __syncthreads();                                  // (red was last read by block_max256)
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 17; ++k) red[warp * 20 + k] = p[k];
      }
      __syncthreads();
      if (tid < 16) {
        float s = 0.f, tot = 0.f;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) { s += red[w * 20 + tid]; tot += red[w * 20 + 16]; }
        const float mm = expf(-100.0f / gamma);
        LRd[b * 16 + tid] = 1e-4f * (s - mm * tot) / (1.0f - mm);
      }
    }
    // all TMEM reads and shared-memory reads of this sample are complete before the next sample overwrites them
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
  }
}

}  // namespace

int g_psf_mode = 0;     // 0 = tensor-core forward (default), 1 = FFMA forward (psf.cu)

extern "C" {

void tsr_set_psf_mode(int mode) { g_psf_mode = mode; }
int tsr_get_psf_mode(void) { return g_psf_mode; }

int tsr_psf_forward_ffma(const float* alphaBeta, const float* depth, float* HR, float* LRd, float* psf, int B,
                         cudaStream_t stream);

int tsr_psf_forward_tc(const float* alphaBeta, const float* depth, float* HR, float* LRd, float* psf, int B,
                       cudaStream_t stream) {
  TSR_REQUIRE(alphaBeta && depth && HR && LRd && B > 0, "psf_forward_tc: bad argument");
  TSR_REQUIRE(((uintptr_t)depth & 15) == 0 && ((uintptr_t)HR & 15) == 0, "psf_forward_tc: depth / HR must be 16-byte aligned");
  TSR_CUDA(cudaFuncSetAttribute(psf_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  int grid = B < 2 * sms ? B : 2 * sms;
  psf_fwd_tc_kernel<<<grid, NT, SMEM_BYTES, stream>>>(alphaBeta, depth, HR, LRd, psf, B);
  TSR_CHECK_LAUNCH("psf_forward_tc");
  return TSR_OK;
}

// (HR, LRd, psf) = PSF forward model of `depth` (B,100,100) under alphaBeta (B,3).  psf may be NULL.
int tsr_psf_forward(const float* alphaBeta, const float* depth, float* HR, float* LRd, float* psf, int B,
                    cudaStream_t stream) {
  if (g_psf_mode == 1) return tsr_psf_forward_ffma(alphaBeta, depth, HR, LRd, psf, B, stream);
  return tsr_psf_forward_tc(alphaBeta, depth, HR, LRd, psf, B, stream);
}

}  // extern "C"
