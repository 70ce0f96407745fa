// fp32-accurate convolution kernels (NHWC activations, implicit GEMM on the FFMA pipe).
//
// This is the "fp32 mode" of the hot path (north_star tolerance <= 1e-5 relative): every
// product and sum is fp32, the reduction order is fixed (deterministic), weight gradients
// use a two-level split-K reduction.  The tensor-core (tcgen05) path lives in conv_tc.cu.
//
// Replaces the ATen/cuDNN calls behind nn.Conv2d at reference
// model/tactileSR_model.py:37,41,47,53,55,61,168,174,180,186,191,219,220 and their autograd
// backward (cpu/trainer.py:353).
#include "common.cuh"

namespace {

constexpr int FLAG_RELU = 1;

// ---------------------------------------------------------------------------------------------
// weight packing: OIHW fp32 -> [tap][ci][co] (forward) and [tap'][co][ci] with flipped taps (dgrad)
// ---------------------------------------------------------------------------------------------
__global__ void pack_weight_f32_kernel(const float* __restrict__ w, float* __restrict__ wf,
                                       float* __restrict__ wd, int Cout, int Cin, int KS) {
  int taps = KS * KS;
  long long n = (long long)Cout * Cin * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int t = (int)(i % taps);
    long long r = i / taps;
    int ci = (int)(r % Cin);
    int co = (int)(r / Cin);
    float v = w[i];
    if (wf) wf[((long long)t * Cin + ci) * Cout + co] = v;
    if (wd) wd[((long long)(taps - 1 - t) * Cout + co) * Cin + ci] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// forward / data-gradient: out[p][co] = sum_{tap,ci} in[p+shift(tap)][ci] * wp[tap][ci][co]
// tile 128 pixels x (16 TN) couts x 16 channels, 256 threads, 8 x TN outputs per thread: TN = 4 (64 couts) or, when the
// grid still fills the machine, TN = 8 (128 couts: 4 LDS.128 per 64 FFMA instead of 3 per 32).  Every output is one serial
// fmaf chain over (tap, ci) in both variants, so they are bit-identical.
// ---------------------------------------------------------------------------------------------
constexpr int BM = 128, BK = 16, APAD = 4;

template <int TN>
__global__ void __launch_bounds__(256, TN == 8 ? 2 : 3)
conv2d_f32_kernel(const float* __restrict__ in, int in_ld, const float* __restrict__ wp,
                  const float* __restrict__ bias, const float* residual, int res_ld,
                  float* out, int out_ld, int Mtotal, int H, int W, int Cin, int Cout, int KS,
                  int flags) {
  constexpr int BN = 16 * TN;
  constexpr int NB4 = TN / 4;            // float4 groups of 4 couts per thread: columns g*64 + tn*4 (conflict-free LDS.128)
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int tn = tid & 15, tm = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int pad = KS >> 1;
  const int HW = H * W;

  // the two A-load slots of this thread: pixel (tid>>2) and 64 + (tid>>2), channel quad tid&3.  NHWC with (b, y, x)
  // flattened: the neighbour (y + dy, x + dx) of pixel p is pixel p + dy W + dx, so a slot keeps ONE pointer to its centre
  // pixel and the (tap, channel chunk) walk only adds CTA-uniform offsets (no division or 64-bit multiply per K step: in the
  // first version that address arithmetic was half of all executed instructions).
  int ay[2], ax[2];
  const float* acen[2];
  bool avalid[2];
  const int ac4 = (tid & 3) * 4;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    int m = (tid >> 2) + s * 64;
    int p = m0 + m;
    avalid[s] = p < Mtotal;
    int pp = avalid[s] ? p : 0;
    int b = pp / HW, rem = pp - b * HW;
    ay[s] = rem / W;
    ax[s] = rem - ay[s] * W;
    acen[s] = in + (long long)pp * in_ld + ac4;
  }
  const int bk = tid >> 4, bn4 = (tid & 15) * 4;
  const int cchunks = Cin / BK;
  const int nk = KS * KS * cchunks;
  const float* wptr = wp + (long long)bk * Cout + n0 + bn4;      // advances by BK rows of the [tap][ci][co] matrix per K step
  const int wstep = BK * Cout;
  int tdy = -pad, tdx = -pad, c0 = 0;                            // the K step about to be loaded: tap offset, channel chunk
  int tapoff = (tdy * W + tdx) * in_ld;

  // A (transposed on its way into shared memory) is staged in registers across the compute block; the weight tile is a
  // straight copy and goes global -> shared asynchronously (staging both in registers spilled under the 128-register cap,
  // and the spill store waited for the load right where it was meant to be hidden)
  float4 ra[2];
  auto load_global = [&](int buf) {                              // loads the next K step and advances the walk
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const bool ok = avalid[s] && (unsigned)(ay[s] + tdy) < (unsigned)H && (unsigned)(ax[s] + tdx) < (unsigned)W;
      ra[s] = ok ? *reinterpret_cast<const float4*>(acen[s] + tapoff + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int g = 0; g < NB4; ++g) cp_async16(&Bs[buf][bk][g * 64 + bn4], wptr + g * 64, true);
    wptr += wstep;
    c0 += BK;
    if (c0 == Cin) {
      c0 = 0;
      if (++tdx > pad) { tdx = -pad; ++tdy; }
      tapoff = (tdy * W + tdx) * in_ld;
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      int m = (tid >> 2) + s * 64;
      As[buf][ac4 + 0][m] = ra[s].x;
      As[buf][ac4 + 1][m] = ra[s].y;
      As[buf][ac4 + 2][m] = ra[s].z;
      As[buf][ac4 + 3][m] = ra[s].w;
    }
    cp_async_wait_all();
  };

  float2 acc[8][TN / 2];               // pairs of couts: the inner product runs as FFMA2 (pixel broadcast x cout pair)
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN / 2; ++j) acc[i][j] = make_float2(0.f, 0.f);

  load_global(0);
  store_smem(0);
  __syncthreads();
  for (int kc = 0; kc < nk; ++kc) {
    int buf = kc & 1;
    if (kc + 1 < nk) load_global(buf ^ 1);      // (buffer buf^1 was last read before the barrier that ended step kc-1)
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][tm * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][tm * 8 + 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float2 bb[TN / 2];
#pragma unroll
      for (int g = 0; g < NB4; ++g) {
        float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][g * 64 + tn * 4]);
        bb[2 * g] = make_float2(b.x, b.y); bb[2 * g + 1] = make_float2(b.z, b.w);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN / 2; ++j) ffma2s(acc[i][j], a[i], bb[j]);
    }
    if (kc + 1 < nk) store_smem(buf ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int g = 0; g < NB4; ++g) {
    const int n = n0 + g * 64 + tn * 4;
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) bv = *reinterpret_cast<const float4*>(bias + n);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int p = m0 + tm * 8 + i;
      if (p >= Mtotal) continue;
      float4 v = make_float4(acc[i][2 * g].x + bv.x, acc[i][2 * g].y + bv.y, acc[i][2 * g + 1].x + bv.z, acc[i][2 * g + 1].y + bv.w);
      if (residual) {
        float4 r = *reinterpret_cast<const float4*>(residual + (long long)p * res_ld + n);
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
      }
      if (flags & FLAG_RELU) {
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      }
      *reinterpret_cast<float4*>(out + (long long)p * out_ld + n) = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// weight gradient, level 1: partial[s][tap][ci][co] = sum_{p in split s} in[p+shift][ci]*dout[p][co]
// A GEMM over the pixels of the split with both operands pixel-major as they lie in HBM (NHWC): rows = (tap, ci), columns =
// co.  CTA tile 128 rows x 16 TN columns, 16 pixels per K step, 256 threads, 8 x TN outputs per thread (FFMA2 on column
// pairs; 3 or 4 LDS.128 per 16 / 32 FFMA2 -- the 4 x 4 tiles of round 1 were bound by their shared-memory reads).  The 128
// rows are two half-tiles of 64 input channels, each with its own tap: two channel chunks of one tap (Cin >= 128) or the
// same 64 channels under two consecutive taps (Cin = 64), so the same kernel serves every layer width.
// ---------------------------------------------------------------------------------------------
template <int TN>
__global__ void __launch_bounds__(256, 2)
conv2d_wgrad_f32_kernel(const float* __restrict__ in, int in_ld, const float* __restrict__ dout,
                        int dout_ld, float* __restrict__ partial, int Mtotal, int H, int W, int Cin,
                        int Cout, int KS, int pix_per_split) {
  constexpr int BN = 16 * TN, NB4 = TN / 4;
  __shared__ __align__(16) float As[2][16][128];
  __shared__ __align__(16) float Bs[2][16][BN];
  const int tid = threadIdx.x;
  const int tn = tid & 15, tm = tid >> 4;          // columns g*64 + tn*4.., rows tm*8..
  const int otiles = Cout / BN, cchunks = Cin / 64;
  const int ot = blockIdx.x % otiles, rt = blockIdx.x / otiles;
  const int nhalves = KS * KS * cchunks;
  const int pad = KS >> 1;
  const int co0 = ot * BN;
  // the two half-tiles of this CTA
  int hdy[2], hdx[2], hci[2], htap[2];
  bool hok[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int hh = 2 * rt + h;
    hok[h] = hh < nhalves;
    htap[h] = hok[h] ? hh / cchunks : 0;
    hci[h] = hok[h] ? (hh - htap[h] * cchunks) * 64 : 0;
    hdy[h] = htap[h] / KS - pad;
    hdx[h] = htap[h] % KS - pad;
  }
  const int pbeg = blockIdx.y * pix_per_split;
  const int pend = min(pbeg + pix_per_split, Mtotal);
  const int lk = tid >> 4, l4 = (tid & 15) * 4;    // load slot: pixel lk of the step, 4 channels from l4

  // this thread's load pixel, advanced by 16 per step without divisions; the shifted pixel of a tap is the linear pixel
  // index plus dy W + dx (NHWC with (b, y, x) flattened), so each half-tile keeps one running pointer
  int lp = pbeg + lk;
  int ly, lx;
  {
    const int HW = H * W;
    const int pp = lp < Mtotal ? lp : 0;
    const int rem = pp - (pp / HW) * HW;
    ly = rem / W;
    lx = rem - ly * W;
  }
  const float* aptr[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) aptr[h] = in + ((long long)lp + hdy[h] * W + hdx[h]) * in_ld + hci[h] + l4;
  const float* bptr = dout + (long long)lp * dout_ld + co0 + l4;
  const int astep = 16 * in_ld, bstep = 16 * dout_ld;
  // both tiles are straight 16-byte copies: global -> shared asynchronously (zero fill outside the image / the split), no
  // staging registers
  auto load_tiles = [&](int buf) {
    const bool pv = lp < pend;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const bool ok = pv && hok[h] && (unsigned)(ly + hdy[h]) < (unsigned)H && (unsigned)(lx + hdx[h]) < (unsigned)W;
      cp_async16(&As[buf][lk][h * 64 + l4], ok ? aptr[h] : in, ok);
      aptr[h] += astep;
    }
#pragma unroll
    for (int g = 0; g < NB4; ++g) cp_async16(&Bs[buf][lk][g * 64 + l4], pv ? bptr + g * 64 : dout, pv);
    bptr += bstep;
    lp += 16;
    lx += 16;
    while (lx >= W) { lx -= W; ++ly; }
    while (ly >= H) ly -= H;
  };

  float2 acc[8][TN / 2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN / 2; ++j) acc[i][j] = make_float2(0.f, 0.f);

  const int nsteps = (pend - pbeg + 15) / 16;
  if (nsteps > 0) load_tiles(0);
  cp_async_wait_all();
  __syncthreads();
  for (int s = 0; s < nsteps; ++s) {
    const int buf = s & 1;
    if (s + 1 < nsteps) load_tiles(buf ^ 1);     // (buffer buf^1 was last read before the barrier that ended step s-1)
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][tm * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][tm * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float2 bb[TN / 2];
#pragma unroll
      for (int g = 0; g < NB4; ++g) {
        const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][g * 64 + tn * 4]);
        bb[2 * g] = make_float2(b.x, b.y); bb[2 * g + 1] = make_float2(b.z, b.w);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN / 2; ++j) ffma2s(acc[i][j], a[i], bb[j]);
    }
    cp_async_wait_all();
    __syncthreads();
  }
  const bool h1 = tm >= 8;                          // rows tm*8 .. tm*8+7 lie in one half-tile
  if (!(h1 ? hok[1] : hok[0])) return;
  float* dst = partial + (((long long)blockIdx.y * KS * KS + (h1 ? htap[1] : htap[0])) * Cin + (h1 ? hci[1] : hci[0]) +
                          (tm & 7) * 8) * Cout + co0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int g = 0; g < NB4; ++g)
      *reinterpret_cast<float4*>(dst + (long long)i * Cout + g * 64 + tn * 4) =
          make_float4(acc[i][2 * g].x, acc[i][2 * g].y, acc[i][2 * g + 1].x, acc[i][2 * g + 1].y);
}

// level 2: dw_oihw[co][ci][tap] (+)= sum_s partial[s][tap][ci][co]   (fixed order => deterministic)
// block = 32 consecutive outputs x 8 split lanes (warp j sums the splits j, j + 8, ...; lane sums combined in lane order
// through shared memory): the dependent-load chain is S / 8 long instead of S
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int S,
                    int taps, int Cin, int Cout, int accumulate) {
  __shared__ float sh[8][32];
  const long long n = (long long)taps * Cin * Cout;
  const int lane = threadIdx.x & 31, j = threadIdx.x >> 5;
  const long long i = (long long)blockIdx.x * 32 + lane;
  const bool ok = i < n;
  float s = 0.f;
  if (ok) {
#pragma unroll 4
    for (int k = j; k < S; k += 8) s += partial[(long long)k * n + i];
  }
  sh[j][lane] = s;
  __syncthreads();
  if (j != 0 || !ok) return;
  float t = sh[0][lane];
#pragma unroll
  for (int q = 1; q < 8; ++q) t += sh[q][lane];
  const int co = (int)(i % Cout);
  const long long r = i / Cout;
  const int ci = (int)(r % Cin);
  const int tap = (int)(r / Cin);
  const long long o = ((long long)co * Cin + ci) * taps + tap;
  dw[o] = accumulate ? dw[o] + t : t;
}

// ---------------------------------------------------------------------------------------------
// per-channel column sums (bias gradients): two-level, deterministic
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void colsum_partial_kernel(const T* __restrict__ x, int ld, int npix, int C,
                                      float* __restrict__ partial, int rows_per_block) {
  // blockDim = (C/4 threads in x... ) generic: thread handles channel quad q = tid % (C/4), row lane = tid / (C/4)
  int q4 = C / 4;
  int lanes = blockDim.x / q4;
  int q = threadIdx.x % q4, lane = threadIdx.x / q4;
  int r0 = blockIdx.x * rows_per_block;
  int r1 = min(r0 + rows_per_block, npix);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane < lanes)
    for (int r = r0 + lane; r < r1; r += lanes) {
      float4 v = ld4(x + (long long)r * ld + q * 4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  extern __shared__ float4 sm[];
  sm[threadIdx.x] = s;
  __syncthreads();
  if (lane == 0) {
    for (int l = 1; l < lanes; ++l) {
      float4 v = sm[l * q4 + q];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(partial + (long long)blockIdx.x * C + q * 4) = s;
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ partial, int nblocks, int C,
                                    float* __restrict__ out, int accumulate) {
  // block = (32 channels, 32 partial-row lanes); fixed-order double accumulation
  __shared__ double sh[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  double a = 0.0;
  if (c < C)
    for (int r = threadIdx.y; r < nblocks; r += 32) a += (double)partial[(long long)r * C + c];
  sh[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    double s = 0.0;
    for (int r = 0; r < 32; ++r) s += sh[r][threadIdx.x];
    out[c] = accumulate ? out[c] + (float)s : (float)s;
  }
}

// ---------------------------------------------------------------------------------------------
// head: bilinear x sf upsample of a (3,4,4) taxel frame fused with the 3x3 conv 3->64 (no bias)
// reference: nn.Upsample + nn.Conv2d at tactileSR_model.py:35-37, 60-61 (+ReLU :62), :107,122
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void build_upsampled(const float* __restrict__ xs /*3x16*/, float* up, int sf) {
  // up: (H+2) x (W+2) x 3 with a zero ring; H = W = 4*sf.  ATen: src = (d+0.5)/sf - 0.5 clamped at 0.
  const int H = 4 * sf, P = H + 2;
  const float scale = 1.0f / (float)sf;
  for (int i = threadIdx.x; i < P * P; i += blockDim.x) {
    int yy = i / P - 1, xx = i % P - 1;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
    if (yy >= 0 && yy < H && xx >= 0 && xx < H) {
      float sy = fmaxf(scale * (yy + 0.5f) - 0.5f, 0.f), sx = fmaxf(scale * (xx + 0.5f) - 0.5f, 0.f);
      int y0 = (int)sy, x0 = (int)sx;
      int y1 = min(y0 + 1, 3), x1 = min(x0 + 1, 3);
      float ly = sy - y0, lx = sx - x0;
      float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
#define TSR_UP(c) (w00 * xs[c * 16 + y0 * 4 + x0] + w01 * xs[c * 16 + y0 * 4 + x1] + \
                   w10 * xs[c * 16 + y1 * 4 + x0] + w11 * xs[c * 16 + y1 * 4 + x1])
      v0 = TSR_UP(0); v1 = TSR_UP(1); v2 = TSR_UP(2);
#undef TSR_UP
    }
    *reinterpret_cast<float4*>(up + i * 4) = make_float4(v0, v1, v2, 0.f);      // 16 bytes per pixel: one LDS.128
  }
}

// Persistent over samples; a thread owns one channel quad (16 quads) with its 27 x 4 weights in registers and strides over
// the pixels (16 pixel lanes): per pixel 27 shared-memory reads of the upsampled frame (broadcast within the half-warp that
// shares the pixel) feed 108 FMAs, and the 16 quads of a pixel store 64 contiguous channels.
template <typename OutT>
__global__ void __launch_bounds__(256, 1)
head_fwd_kernel(const float* __restrict__ x, long long x_bstride, const float* __restrict__ w /*64,3,3,3*/,
                OutT* __restrict__ out, int out_ld, int B, int sf, int relu) {
  extern __shared__ float smem[];
  float* xs = smem;              // 48
  float* up = xs + 48;           // (H+2)^2*3
  const int H = 4 * sf, P = H + 2;
  const int g = threadIdx.x & 15, pl = threadIdx.x >> 4;
  float2 wr[27][2];              // [tap*3 + c][channel pair of the quad]
#pragma unroll
  for (int q = 0; q < 27; ++q)
#pragma unroll
    for (int j = 0; j < 2; ++j)
      wr[q][j] = make_float2(w[((g * 4 + 2 * j) * 3 + q % 3) * 9 + q / 3], w[((g * 4 + 2 * j + 1) * 3 + q % 3) * 9 + q / 3]);
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    if (threadIdx.x < 48) xs[threadIdx.x] = x[(long long)b * x_bstride + threadIdx.x];
    __syncthreads();
    build_upsampled(xs, up, sf);
    __syncthreads();
    int y = pl / H, xx = pl - y * H;                 // advanced incrementally: an integer division per pixel costs as
    for (int p = pl; p < H * H; p += 16) {           // much as a quarter of the pixel's FMAs
      const float4* u0 = reinterpret_cast<const float4*>(up) + (y * P + xx);
      float2 a01 = make_float2(0.f, 0.f), a23 = a01;         // channel pairs: FFMA2
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const float4 u4 = u0[(tap / 3) * P + tap % 3];       // the three axes of one neighbour pixel
        const float uv[3] = {u4.x, u4.y, u4.z};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int q = tap * 3 + c;
          ffma2s(a01, uv[c], wr[q][0]);
          ffma2s(a23, uv[c], wr[q][1]);
        }
      }
      float4 acc = make_float4(a01.x, a01.y, a23.x, a23.y);
      if (relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
      st4(out + ((long long)b * H * H + p) * out_ld + g * 4, acc);
      xx += 16;
      while (xx >= H) { xx -= H; ++y; }
    }
  }
}

// four consecutive channels as the raw words of their storage type (conversion deferred to the point of use)
template <typename T> struct Raw4;
template <> struct Raw4<float> {
  typedef float4 type;
  __device__ static __forceinline__ type load(const float* p) { return *reinterpret_cast<const float4*>(p); }
  __device__ static __forceinline__ float4 cvt(type r) { return r; }
};
template <> struct Raw4<__nv_bfloat16> {
  typedef uint2 type;
  __device__ static __forceinline__ type load(const __nv_bfloat16* p) { return *reinterpret_cast<const uint2*>(p); }
  __device__ static __forceinline__ float4 cvt(type r) {
    const float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&r.x)), b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
};
template <> struct Raw4<__half> {
  typedef uint2 type;
  __device__ static __forceinline__ type load(const __half* p) { return *reinterpret_cast<const uint2*>(p); }
  __device__ static __forceinline__ float4 cvt(type r) {
    const float2 a = __half22float2(*reinterpret_cast<__half2*>(&r.x)), b = __half22float2(*reinterpret_cast<__half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
};

// head weight gradient, level 1: one partial [27][64] per CTA, CTAs stride over samples.
// HEAD_WG_NT threads = 16 channel quads x 16 pixel lanes; a thread keeps all 27 x 4 (tap, axis, channel) sums in registers
// (as 54 channel pairs: FFMA2; 256 threads per CTA: the paired accumulators spill under the 128- / 168-register caps of
// 512 / 384 threads), so a pixel costs 9 shared-memory reads of the upsampled frame, one of the gradient and 54 FFMA2.
// The gradient rows are staged in shared memory by 16-byte asynchronous copies, `chpix` pixels (<= 40 KB) per chunk, double
// buffered across chunks AND samples: with one 8-warp CTA per SM, register prefetch kept only 6 KB per SM in flight and the
// kernel ran at 0.57 TB/s of HBM latency; a staged chunk keeps 40 KB in flight.
constexpr int HEAD_WG_NT = 256, HEAD_WG_LANES = HEAD_WG_NT / 16;
template <typename GT> struct HeadWg {
  static constexpr int ROWB = 64 * (int)sizeof(GT);            // bytes of one pixel's 64 gradients
  static constexpr int CH_MAX = 40960 / ROWB;                  // pixels per staged chunk (320 / 160)
};
template <typename GT>
__global__ void __launch_bounds__(HEAD_WG_NT, 1)
head_wgrad_kernel(const float* __restrict__ x, long long x_bstride, const GT* __restrict__ dout, int dout_ld,
                  float* __restrict__ partial, int B, int sf, int up_floats, int chpix) {
  constexpr int LN = HEAD_WG_LANES, ROWB = HeadWg<GT>::ROWB, PIECES = ROWB / 16;
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                       // 48
  float* up = xs + 48;                    // (H+2)^2 * 4 (>= 2048), later reused for the lane reduction (LN x 64 floats)
  uint8_t* stage = reinterpret_cast<uint8_t*>(up + up_floats);   // 2 x chpix x ROWB
  const int H = 4 * sf, P = H + 2, HH = H * H;
  const int tid = threadIdx.x;
  const int g = tid & 15, pl = tid >> 4;  // channel quad, pixel lane 0..LN-1
  float2 acc[27][2];                      // channel pairs: FFMA2
#pragma unroll
  for (int q = 0; q < 27; ++q)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[q][j] = make_float2(0.f, 0.f);

  const int nch = (HH + chpix - 1) / chpix;
  const int nsamples = (int)blockIdx.x < B ? (B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int nitems = nsamples * nch;      // work items (sample, chunk) of this CTA, in order
  auto issue = [&](int it, int buf) {     // asynchronous copy of item it's gradient rows into stage buffer buf
    const int si = it / nch, c0 = (it - si * nch) * chpix;
    const int b = blockIdx.x + si * gridDim.x;
    const int n = min(chpix, HH - c0);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(dout + ((long long)b * HH + c0) * dout_ld);
    uint8_t* dst = stage + (size_t)buf * chpix * ROWB;
    const long long rstride = (long long)dout_ld * (long long)sizeof(GT);
    for (int i = tid; i < n * PIECES; i += HEAD_WG_NT) {
      const int row = i / PIECES, q = i - row * PIECES;
      cp_async16(dst + row * ROWB + q * 16, src + row * rstride + q * 16, true);
    }
    cp_async_commit();
  };
  if (nitems > 0) issue(0, 0);
  for (int it = 0; it < nitems; ++it) {
    const int buf = it & 1;
    const int si = it / nch, c0 = (it - si * nch) * chpix;
    if (c0 == 0) {                        // a new sample: its upsampled frame (the previous sample's pixels are all consumed:
      const int b = blockIdx.x + si * gridDim.x;      //  barrier at the end of the previous item)
      if (tid < 48) xs[tid] = x[(long long)b * x_bstride + tid];
      __syncthreads();
      build_upsampled(xs, up, sf);
    }
    if (it + 1 < nitems) {
      issue(it + 1, buf ^ 1);             // (buffer buf^1 was read in item it-1, which ended with a barrier)
      cp_async_wait_group<1>();
    } else {
      cp_async_wait_group<0>();
    }
    __syncthreads();                      // this item's rows (every thread's copies) and `up` are visible
    const int n = min(chpix, HH - c0);
    const uint8_t* rows = stage + (size_t)buf * chpix * ROWB + g * 4 * sizeof(GT);
    int y = (c0 + pl) / H, xx = (c0 + pl) - y * H;
    for (int p = pl; p < n; p += LN) {
      const float4 gv = Raw4<GT>::cvt(Raw4<GT>::load(reinterpret_cast<const GT*>(rows + (size_t)p * ROWB)));
      const float2 g01 = make_float2(gv.x, gv.y), g23 = make_float2(gv.z, gv.w);
      const float4* u = reinterpret_cast<const float4*>(up) + (y * P + xx);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const float4 u4 = u[(tap / 3) * P + tap % 3];
        const float uv[3] = {u4.x, u4.y, u4.z};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int q = tap * 3 + c;
          ffma2s(acc[q][0], uv[c], g01);
          ffma2s(acc[q][1], uv[c], g23);
        }
      }
      xx += LN;
      while (xx >= H) { xx -= H; ++y; }
    }
    __syncthreads();                      // stage[buf] and (at a sample's last chunk) `up` may be overwritten
  }
  // fixed-order reduction over the pixel lanes, one (tap, axis) at a time through shared memory
  float4* red = reinterpret_cast<float4*>(up);      // [LN lanes][16 quads]
#pragma unroll
  for (int q = 0; q < 27; ++q) {
    red[pl * 16 + g] = make_float4(acc[q][0].x, acc[q][0].y, acc[q][1].x, acc[q][1].y);
    __syncthreads();
    if (pl == 0) {
      float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int l = 0; l < LN; ++l) {
        const float4 v = red[l * 16 + g];
        sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
      }
      *reinterpret_cast<float4*>(partial + ((long long)blockIdx.x * 27 + q) * 64 + g * 4) = sum;
    }
    __syncthreads();
  }
}
__global__ void head_wgrad_reduce_kernel(const float* __restrict__ partial, int nparts, float* __restrict__ dw,
                                         int accumulate) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;   // i = q*64 + co
  if (i >= 27 * 64) return;
  int co = i & 63, q = i >> 6;
  int tap = q / 3, c = q % 3;
  float s = 0.f;
  for (int k = 0; k < nparts; ++k) s += partial[(long long)k * 27 * 64 + i];
  int o = (co * 3 + c) * 9 + tap;
  dw[o] = accumulate ? dw[o] + s : s;
}

// ---------------------------------------------------------------------------------------------
// tail: 3x3 conv Cin -> 1 (no bias) + ReLU.   reference tactileSR_model.py:55-56, :125-126
// ---------------------------------------------------------------------------------------------
// A CTA owns TR = 4 rows x W columns of one sample: the (TR+2) x (W+2) x Cin halo tile is staged once in shared memory
// (coalesced 16-byte loads, zero fill = padding), so every input element is read from HBM/L2 once instead of 9 times;
// then one warp per output pixel: lane = 4 channels, the 9x4 weights of the lane stay in registers, warp-shuffle sum.
constexpr int TAIL_TR = 4;
template <typename InT>
__global__ void __launch_bounds__(256)
tail_fwd_kernel(const InT* __restrict__ in, int in_ld, const float* __restrict__ w /*1,Cin,3,3*/,
                float* __restrict__ out, int B, int H, int W, int Cin, int relu) {
  extern __shared__ __align__(16) uint8_t tail_smem[];
  constexpr int VEC = 16 / sizeof(InT);                 // elements per 16-byte access
  const int pitch = Cin + VEC;                          // padded pixel pitch (elements): conflict-free lane access
  InT* tile = reinterpret_cast<InT*>(tail_smem);
  const int strips = (H + TAIL_TR - 1) / TAIL_TR;
  const int b = blockIdx.x / strips, y0 = (blockIdx.x % strips) * TAIL_TR;
  const int PW = W + 2, PH = TAIL_TR + 2;
  const int vpp = Cin / VEC;                            // vectors per pixel
  // thread -> (vector v of a pixel, pixel lane); the pixel coordinates advance incrementally (no integer divisions in the
  // loop: they cost more than the copy itself), 4 independent 16-byte loads in flight per thread
  {
    const int v = threadIdx.x % vpp, lanes = blockDim.x / vpp;      // blockDim.x is a multiple of vpp (Cin <= 1024)
    const int npix = PH * PW;
    int pix = threadIdx.x / vpp;
    int py = pix / PW, px = pix - py * PW;
    while (pix < npix) {
      uint4 val[4];
      int pixk[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        pixk[k] = pix;
        val[k] = make_uint4(0u, 0u, 0u, 0u);
        const int gx = px - 1, gy = y0 + py - 1;
        if (pix < npix && gx >= 0 && gx < W && gy >= 0 && gy < H)
          val[k] = *reinterpret_cast<const uint4*>(in + ((long long)(b * H + gy) * W + gx) * in_ld + v * VEC);
        pix += lanes;
        px += lanes;
        while (px >= PW) { px -= PW; ++py; }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (pixk[k] < npix) *reinterpret_cast<uint4*>(tile + (long long)pixk[k] * pitch + v * VEC) = val[k];
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // this lane's weights: channels c = lane*4 + 128*k (k = 0 for Cin <= 128)
  float2 wr[9][2];                                      // channel pairs (FFMA2)
  const int c0 = lane * 4;
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 2; ++j)
      wr[t][j] = make_float2((c0 + 2 * j < Cin) ? w[(c0 + 2 * j) * 9 + t] : 0.f, (c0 + 2 * j + 1 < Cin) ? w[(c0 + 2 * j + 1) * 9 + t] : 0.f);
  __syncthreads();
  // 4 pixels per warp iteration (independent accumulation chains), then one transposing reduction of the 4 sums
  const int npx = min(TAIL_TR, H - y0) * W;
  for (int p0 = warp * 4; p0 < npx; p0 += nw * 4) {
    float sum4[4] = {0.f, 0.f, 0.f, 0.f};
    if (c0 < Cin) {
      const int ty0 = p0 / W, x0 = p0 - ty0 * W;      // one division per 4 pixels
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int ty = ty0, x = x0 + k;
        if (x >= W) { x -= W; ++ty; }
        if (p0 + k >= npx) { ty = ty0; x = x0; }       // (padding lanes of the last group recompute pixel p0)
        float2 s01 = make_float2(0.f, 0.f), s23 = s01; // even / odd channel partial sums: two 9-long FFMA2 chains per pixel
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float4 v = ld4(tile + (long long)((ty + t / 3) * PW + x + t % 3) * pitch + c0);
          ffma2(s01, make_float2(v.x, v.y), wr[t][0]);
          ffma2(s23, make_float2(v.z, v.w), wr[t][1]);
        }
        sum4[k] = (s01.x + s01.y) + (s23.x + s23.y);
      }
    }
    if (Cin > 128) {   // generic tail for wider inputs (not used by the reference networks)
      for (int k = 0; k < 4; ++k) {
        const int p = min(p0 + k, npx - 1);
        const int ty = p / W, x = p - ty * W;
        for (int c = c0 + 128; c < Cin; c += 128)
          for (int t = 0; t < 9; ++t) {
            const float4 v = ld4(tile + (long long)((ty + t / 3) * PW + x + t % 3) * pitch + c);
            sum4[k] += v.x * w[c * 9 + t] + v.y * w[(c + 1) * 9 + t] + v.z * w[(c + 2) * 9 + t] + v.w * w[(c + 3) * 9 + t];
          }
      }
    }
    // lanes exchange halves: after two exchanges lane l carries pixel ((l >> 4) & 1) * 2 + ((l >> 3) & 1)
    {
      const bool up16 = (lane & 16) != 0;
      const float k0 = up16 ? sum4[2] : sum4[0], s0 = up16 ? sum4[0] : sum4[2];
      const float k1 = up16 ? sum4[3] : sum4[1], s1 = up16 ? sum4[1] : sum4[3];
      const float a0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 16), a1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
      const bool up8 = (lane & 8) != 0;
      float v = (up8 ? a1 : a0) + __shfl_xor_sync(0xffffffffu, up8 ? a0 : a1, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      const int k = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
      if ((lane & 7) == 0 && p0 + k < npx) {
        const int p = p0 + k, ty = p / W, x = p - ty * W;
        out[((long long)b * H + y0 + ty) * W + x] = relu ? fmaxf(v, 0.f) : v;
      }
    }
  }
}

// Register-blocked tail forward for Cin <= 128 (every reference network): same strip tile as above, staged by 16-byte
// asynchronous copies (zero fill = padding), then ONE WARP PER 4 x 5 BLOCK OF OUTPUT PIXELS (8 blocks = 8 warps for W = 40):
// lane = 4 channels, its 9 x 4 weights in registers; each of the block's 6 x 7 input pixels is read from shared memory ONCE
// and feeds up to nine outputs (2.1 shared-memory reads and conversions per output instead of 9: the one-pixel-per-warp
// kernel was bound by exactly those -- 398 us at B = 1024 against 64 us of HBM time), accumulators as (even, odd) channel
// pairs (FFMA2), then a transposing warp reduction of the 20 sums (22 shuffles).
constexpr int TAIL_BW = 5;
template <typename InT>
__global__ void __launch_bounds__(512, 1)
tail_fwd_blocked_kernel(const InT* __restrict__ in, int in_ld, const float* __restrict__ w /*1,Cin,3,3*/,
                        float* __restrict__ out, int B, int H, int W, int Cin, int relu, int nbuf) {
  // nbuf = 2: persistent CTAs (one per SM), the next strip's tile is copied while this one is computed (16-bit inputs: two
  // tiles fit); nbuf = 1: one strip per CTA (fp32 inputs), the CTAs of an SM overlap each other instead.
  extern __shared__ __align__(16) uint8_t tail_smem[];
  constexpr int VEC = 16 / sizeof(InT);
  const int pitch = Cin + VEC;
  const int strips = (H + TAIL_TR - 1) / TAIL_TR, total = B * strips;
  const int PW = W + 2, PH = TAIL_TR + 2;
  const size_t tile_elems = (size_t)PH * PW * pitch;
  const int vpp = Cin / VEC;
  // Tile copy: thread = (16-byte vector v of a pixel, pixel lane).  All address arithmetic is hoisted: the shared-memory
  // address advances by a constant, the global one is a per-strip base plus a 32-bit (row, column) offset (in the first
  // version the 64-bit products and generic-to-shared conversions made this loop 46 % of the kernel's instructions).
  const int cv = threadIdx.x % vpp, clanes = blockDim.x / vpp;      // blockDim.x is a multiple of vpp
  const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(tail_smem);
  const uint32_t tile_bytes = (uint32_t)(tile_elems * sizeof(InT)), pix_bytes = (uint32_t)(pitch * sizeof(InT));
  auto issue = [&](int sidx, int buf) {
    const int b = sidx / strips, y0 = (sidx - b * strips) * TAIL_TR;
    const InT* g0 = in + ((long long)(b * H + y0 - 1) * W - 1) * in_ld + cv * VEC;     // pixel (py, px) = (0, 0) of the tile
    const int npix = PH * PW;
    int pix = threadIdx.x / vpp;
    int py = pix / PW, px = pix - py * PW;
    uint32_t dst = smem0 + (uint32_t)buf * tile_bytes + (uint32_t)pix * pix_bytes + (uint32_t)(cv * 16);
    const uint32_t dstep = (uint32_t)clanes * pix_bytes;
    while (pix < npix) {
      const bool ok = (unsigned)(px - 1) < (unsigned)W && (unsigned)(y0 + py - 1) < (unsigned)H;
      const InT* src = ok ? g0 + (py * W + px) * in_ld : in;
      const int sz = ok ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
      dst += dstep;
      pix += clanes;
      px += clanes;
      while (px >= PW) { px -= PW; ++py; }
    }
    cp_async_commit();
  };
  if ((int)blockIdx.x < total) issue(blockIdx.x, 0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int c0 = lane * 4;
  float2 wr[9][2];                                      // this lane's weights as channel pairs
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 2; ++j)
      wr[t][j] = make_float2((c0 + 2 * j < Cin) ? w[(c0 + 2 * j) * 9 + t] : 0.f, (c0 + 2 * j + 1 < Cin) ? w[(c0 + 2 * j + 1) * 9 + t] : 0.f);
  const int nblk = (W + TAIL_BW - 1) / TAIL_BW;
  int it = 0;
  for (int sidx = blockIdx.x; sidx < total; sidx += gridDim.x, ++it) {
    const int buf = nbuf == 2 ? (it & 1) : 0;
    const int next = sidx + gridDim.x;
    if (nbuf == 2 && next < total) {
      issue(next, buf ^ 1);                             // (tile buf^1 was read in the previous iteration, which ended with a barrier)
      cp_async_wait_group<1>();
    } else {
      cp_async_wait_group<0>();
    }
    __syncthreads();
    const int b = sidx / strips, y0 = (sidx - b * strips) * TAIL_TR;
    const InT* tile = reinterpret_cast<const InT*>(tail_smem) + buf * tile_elems;
    for (int blk = warp; blk < nblk; blk += nw) {
      const int x0 = blk * TAIL_BW;
      float2 acc[TAIL_TR][TAIL_BW];
#pragma unroll
      for (int oy = 0; oy < TAIL_TR; ++oy)
#pragma unroll
        for (int ox = 0; ox < TAIL_BW; ++ox) acc[oy][ox] = make_float2(0.f, 0.f);
      if (c0 < Cin) {
#pragma unroll
        for (int iy = 0; iy < TAIL_TR + 2; ++iy) {
#pragma unroll
          for (int ix = 0; ix < TAIL_BW + 2; ++ix) {
            const int col = min(x0 + ix, PW - 1);        // (a partial last block re-reads the right padding column: masked)
            const float4 v = ld4(tile + (long long)(iy * PW + col) * pitch + c0);
            const float2 v01 = make_float2(v.x, v.y), v23 = make_float2(v.z, v.w);
#pragma unroll
            for (int oy = 0; oy < TAIL_TR; ++oy) {
              const int ty = iy - oy;
              if (ty < 0 || ty > 2) continue;
#pragma unroll
              for (int ox = 0; ox < TAIL_BW; ++ox) {
                const int tx = ix - ox;
                if (tx < 0 || tx > 2) continue;
                ffma2(acc[oy][ox], v01, wr[ty * 3 + tx][0]);
                ffma2(acc[oy][ox], v23, wr[ty * 3 + tx][1]);
              }
            }
          }
        }
      }
      // 20 per-lane sums -> warp sums.  Values 0..15: every exchange halves what a lane carries (lane l ends with value
      // l >> 1); values 16..19: two halving exchanges, then three plain ones (lane l ends with value 16 + (l >> 3)).
      float p[16], q4[4];
#pragma unroll
      for (int i = 0; i < 16; ++i) p[i] = acc[i / TAIL_BW][i % TAIL_BW].x + acc[i / TAIL_BW][i % TAIL_BW].y;
#pragma unroll
      for (int i = 0; i < 4; ++i) q4[i] = acc[(16 + i) / TAIL_BW][(16 + i) % TAIL_BW].x + acc[(16 + i) / TAIL_BW][(16 + i) % TAIL_BW].y;
#pragma unroll
      for (int wdt = 8, bit = 16; wdt >= 1; wdt >>= 1, bit >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int k = 0; k < wdt; ++k) {
          const float keep = up ? p[k + wdt] : p[k], send = up ? p[k] : p[k + wdt];
          p[k] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
      }
      p[0] += __shfl_xor_sync(0xffffffffu, p[0], 1);
#pragma unroll
      for (int wdt = 2, bit = 16; wdt >= 1; wdt >>= 1, bit >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int k = 0; k < wdt; ++k) {
          const float keep = up ? q4[k + wdt] : q4[k], send = up ? q4[k] : q4[k + wdt];
          q4[k] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
      }
      q4[0] += __shfl_xor_sync(0xffffffffu, q4[0], 4);
      q4[0] += __shfl_xor_sync(0xffffffffu, q4[0], 2);
      q4[0] += __shfl_xor_sync(0xffffffffu, q4[0], 1);
      int idx = -1;
      float val = 0.f;
      if ((lane & 1) == 0) { idx = lane >> 1; val = p[0]; }
      if ((lane & 7) == 1) { idx = 16 + (lane >> 3); val = q4[0]; }     // (odd lanes 1, 9, 17, 25: free in the first group)
      if (idx >= 0) {
        const int oy = idx / TAIL_BW, ox = idx - oy * TAIL_BW;
        if (y0 + oy < H && x0 + ox < W) out[((long long)b * H + y0 + oy) * W + x0 + ox] = relu ? fmaxf(val, 0.f) : val;
      }
    }
    __syncthreads();                                    // every warp is done with tile buf before it is refilled
  }
}

// tail data gradient: din[p][ci] = sum_tap dz[p - shift(tap)] * w[tap][ci],  dz = dout * (out > 0)
// A CTA owns TAIL_TR rows of one sample: the masked output gradient of the strip (+ halo) is staged in shared memory once;
// a thread owns a group of 8 channels (its 9 x 8 weights in registers) and strides over the strip's pixels: 9 shared
// reads (broadcast to the 16 threads that share the pixel), 72 FMAs and one 16/32-byte store, a pixel's 16 groups
// forming one contiguous channel row.
template <typename GT>
__global__ void __launch_bounds__(256)
tail_dgrad_kernel(const float* __restrict__ dout, const float* __restrict__ out_act,
                  const float* __restrict__ w, GT* __restrict__ din, int din_ld, int B, int H, int W,
                  int Cin, int relu, const void* __restrict__ in_act, int in_ld, int in_f16) {
  extern __shared__ float dz[];                     // (TAIL_TR + 2) x (W + 2), zero ring
  const int strips = (H + TAIL_TR - 1) / TAIL_TR;
  const int PW = W + 2, PH = TAIL_TR + 2;
  const int ngroups = Cin / 8;
  // Every thread runs every pass and strip (the staging loop and its barriers need the whole CTA); a thread whose channel
  // group does not exist (Cin < 128: TactileSRCNN's 64 -> 1 tail) only skips the arithmetic and the store.
  for (int g0 = 0; g0 < ngroups; g0 += 16) {                   // (one pass for Cin <= 128)
    const int g = g0 + (threadIdx.x & 15);
    const bool active = g < ngroups;
    float2 wr[9][4];                                           // (channel pairs: FFMA2) loaded once: the CTA is persistent over strips
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        wr[t][j] = active ? make_float2(w[(g * 8 + 2 * j) * 9 + t], w[(g * 8 + 2 * j + 1) * 9 + t]) : make_float2(0.f, 0.f);
    for (int sidx = blockIdx.x; sidx < B * strips; sidx += gridDim.x) {
      const int b = sidx / strips, y0 = (sidx - b * strips) * TAIL_TR;
      __syncthreads();
      for (int i = threadIdx.x; i < PH * PW; i += blockDim.x) {
        const int px = i % PW - 1, py = y0 + i / PW - 1;
        float gv = 0.f;
        if (px >= 0 && px < W && py >= 0 && py < H) {
          const long long o = ((long long)b * H + py) * W + px;
          gv = dout[o];
          if (relu && !(out_act[o] > 0.f)) gv = 0.f;
        }
        dz[i] = gv;
      }
      __syncthreads();
      const int npx = min(TAIL_TR, H - y0) * W;
      int ty = (threadIdx.x >> 4) / W, x = (threadIdx.x >> 4) - ty * W;
      // the saved activation of the NEXT pixel is requested before this pixel's arithmetic (raw 16-byte word): loaded at
      // its point of use, every iteration waited a full HBM latency with 32 pixels in flight per SM (350 us at B = 1024)
      const long long strip0 = ((long long)b * H + y0) * W;
      const int esz = 2;                                   // fp16 / bf16 activations
      auto act_ptr = [&](int p) {
        return reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(in_act) + ((strip0 + p) * in_ld + g * 8) * esz);
      };
      uint4 araw = make_uint4(0u, 0u, 0u, 0u);
      if (in_act && active && (int)(threadIdx.x >> 4) < npx) araw = *act_ptr(threadIdx.x >> 4);
      for (int p = threadIdx.x >> 4; active && p < npx; p += 16) {
        const uint4 acur = araw;
        if (in_act && p + 16 < npx) araw = *act_ptr(p + 16);
        float2 acc2[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          // output pixel o = p - shift(tap) used input p with weight tap
          const float gv = dz[(ty + 1 - (t / 3 - 1)) * PW + x + 1 - (t % 3 - 1)];
#pragma unroll
          for (int j = 0; j < 4; ++j) ffma2s(acc2[j], gv, wr[t][j]);
        }
        float acc[8] = {acc2[0].x, acc2[0].y, acc2[1].x, acc2[1].y, acc2[2].x, acc2[2].y, acc2[3].x, acc2[3].y};
        const long long pixel = ((long long)b * H + y0 + ty) * W + x;
        if (in_act) {        // ReLU backward of the layer that produced the tail's input: zero where its activation is <= 0
          // (`a > 0` on the raw 16-bit words: 0 < bits <= +inf; negative numbers, both zeros and NaN compare false)
          const uint32_t wds[4] = {acur.x, acur.y, acur.z, acur.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t h = (wds[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu;
            const bool pos = in_f16 ? (h != 0u && h < 0x7C01u) : (h != 0u && h < 0x7F81u);   // 0 < a <= +inf
            acc[j] = pos ? acc[j] : 0.f;
          }
        }
        GT* dst = din + pixel * din_ld + g * 8;
        st4(dst, make_float4(acc[0], acc[1], acc[2], acc[3]));
        st4(dst + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
        x += 16;
        while (x >= W) { x -= W; ++ty; }
      }
    }
  }
}

// tail weight gradient, level 1: partial[blk][9][Cin]; thread = (channel quad, pixel lane)
template <typename InT>
__global__ void __launch_bounds__(256)
tail_wgrad_kernel(const InT* __restrict__ in, int in_ld, const float* __restrict__ dout,
                  const float* __restrict__ out_act, float* __restrict__ partial, int Mtotal, int H, int W,
                  int Cin, int relu, int pix_per_block) {
  extern __shared__ float4 red[];   // [lanes][q4] reused per tap
  const int q4 = Cin / 4, lanes = blockDim.x / q4;
  const int q = threadIdx.x % q4, lane = threadIdx.x / q4;
  const int HW = H * W;
  const int p0 = blockIdx.x * pix_per_block, p1 = min(p0 + pix_per_block, Mtotal);
  float4 acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = p0 + lane; p < p1; p += lanes) {
    float g = dout[p];
    if (relu && !(out_act[p] > 0.f)) g = 0.f;
    if (g == 0.f) continue;
    int b = p / HW, rem = p - b * HW;
    int y = rem / W, x = rem - y * W;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
      if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
      float4 v = ld4(in + ((long long)b * HW + (long long)yy * W + xx) * in_ld + q * 4);
      acc[tap].x = fmaf(g, v.x, acc[tap].x); acc[tap].y = fmaf(g, v.y, acc[tap].y);
      acc[tap].z = fmaf(g, v.z, acc[tap].z); acc[tap].w = fmaf(g, v.w, acc[tap].w);
    }
  }
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    __syncthreads();
    red[lane * q4 + q] = acc[tap];
    __syncthreads();
    if (lane == 0) {
      float4 s = acc[tap];
      for (int l = 1; l < lanes; ++l) {
        float4 v = red[l * q4 + q];
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      *reinterpret_cast<float4*>(partial + ((long long)blockIdx.x * 9 + tap) * Cin + q * 4) = s;
    }
  }
}
// (32 outputs x 8 split lanes per block, lane sums combined in lane order: the chain of 592 dependent double additions per
//  output becomes 74)
__global__ void __launch_bounds__(256)
tail_wgrad_reduce_kernel(const float* __restrict__ partial, int nparts, int Cin,
                         float* __restrict__ dw, int accumulate) {
  __shared__ double sh[8][32];
  const int lane = threadIdx.x & 31, j = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;   // i = tap*Cin + ci
  const bool ok = i < 9 * Cin;
  double s = 0.0;
  if (ok)
    for (int k = j; k < nparts; k += 8) s += (double)partial[(long long)k * 9 * Cin + i];
  sh[j][lane] = s;
  __syncthreads();
  if (j != 0 || !ok) return;
  double t = sh[0][lane];
#pragma unroll
  for (int q = 1; q < 8; ++q) t += sh[q][lane];
  const int ci = i % Cin, tap = i / Cin;
  const int o = ci * 9 + tap;
  dw[o] = accumulate ? dw[o] + (float)t : (float)t;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int tsr_pack_conv_weight_f32(const float* w_oihw, float* w_fwd, float* w_dgrad, int Cout, int Cin, int KS,
                             cudaStream_t stream) {
  TSR_REQUIRE(w_oihw && (w_fwd || w_dgrad), "pack_conv_weight_f32: null pointer");
  long long n = (long long)Cout * Cin * KS * KS;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 2048) blocks = 2048;
  pack_weight_f32_kernel<<<blocks, 256, 0, stream>>>(w_oihw, w_fwd, w_dgrad, Cout, Cin, KS);
  TSR_CHECK_LAUNCH("pack_conv_weight_f32");
  return TSR_OK;
}

int tsr_conv2d_f32(const float* in, int in_ld, const float* w_packed, const float* bias,
                   const float* residual, int res_ld, float* out, int out_ld, int B, int H, int W, int Cin,
                   int Cout, int KS, int flags, cudaStream_t stream) {
  TSR_REQUIRE(in && w_packed && out, "conv2d_f32: null pointer");
  TSR_REQUIRE(Cin % 16 == 0 && Cout % 64 == 0, "conv2d_f32: need Cin %% 16 == 0 and Cout %% 64 == 0 (got %d, %d)", Cin, Cout);
  TSR_REQUIRE(KS == 1 || KS == 3 || KS == 5, "conv2d_f32: kernel size %d unsupported", KS);
  TSR_REQUIRE(in_ld % 4 == 0 && out_ld % 4 == 0 && (!residual || res_ld % 4 == 0), "conv2d_f32: row strides must be multiples of 4");
  long long M = (long long)B * H * W;
  TSR_REQUIRE(M > 0 && M < (1ll << 31), "conv2d_f32: bad pixel count");
  // 128-wide cout tiles once they still give every SM its two resident CTAs
  if (Cout % 128 == 0 && (long long)tsr_cdiv(M, BM) * (Cout / 128) >= 2 * 148) {
    dim3 grid(tsr_cdiv(M, BM), Cout / 128);
    conv2d_f32_kernel<8><<<grid, 256, 0, stream>>>(in, in_ld, w_packed, bias, residual, res_ld, out, out_ld, (int)M,
                                                   H, W, Cin, Cout, KS, flags);
  } else {
    dim3 grid(tsr_cdiv(M, BM), Cout / 64);
    conv2d_f32_kernel<4><<<grid, 256, 0, stream>>>(in, in_ld, w_packed, bias, residual, res_ld, out, out_ld, (int)M,
                                                   H, W, Cin, Cout, KS, flags);
  }
  TSR_CHECK_LAUNCH("conv2d_f32");
  return TSR_OK;
}

static int wgrad_splits(long long M, int tiles) {
  int s = (4 * 148) / tiles;       // two full waves at 2 CTAs per SM, never a third (rounding up cost 20 % as a 2-CTA tail wave)
  int maxs = (int)((M + 255) / 256);
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  if (s > 256) s = 256;
  return s;
}

// level-1 CTAs of one split: pairs of (tap, 64-channel chunk) half-tiles x column tiles of 128 (or 64) output channels
static int wgrad_f32_tiles(int Cin, int Cout, int KS) {
  const int nhalves = KS * KS * (Cin / 64);
  return ((nhalves + 1) / 2) * (Cout % 128 == 0 ? Cout / 128 : Cout / 64);
}

size_t tsr_conv2d_wgrad_f32_workspace(int B, int H, int W, int Cin, int Cout, int KS) {
  long long M = (long long)B * H * W;
  int tiles = wgrad_f32_tiles(Cin, Cout, KS);
  if (tiles <= 0) return 0;
  return (size_t)wgrad_splits(M, tiles) * KS * KS * Cin * Cout * sizeof(float);
}

int tsr_conv2d_wgrad_f32(const float* in, int in_ld, const float* dout, int dout_ld, float* dw_oihw,
                         void* workspace, size_t ws_bytes, int B, int H, int W, int Cin, int Cout, int KS,
                         int accumulate, cudaStream_t stream) {
  TSR_REQUIRE(in && dout && dw_oihw && workspace, "conv2d_wgrad_f32: null pointer");
  TSR_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "conv2d_wgrad_f32: need Cin, Cout %% 64 == 0 (got %d, %d)", Cin, Cout);
  TSR_REQUIRE(in_ld % 4 == 0 && dout_ld % 4 == 0, "conv2d_wgrad_f32: row strides must be multiples of 4");
  long long M = (long long)B * H * W;
  int taps = KS * KS;
  int tiles = wgrad_f32_tiles(Cin, Cout, KS);
  int S = wgrad_splits(M, tiles);
  size_t need = (size_t)S * taps * Cin * Cout * sizeof(float);
  if (ws_bytes < need) {
    tsr_set_error("conv2d_wgrad_f32: workspace too small (%zu < %zu)", ws_bytes, need);
    return TSR_ERR_WORKSPACE;
  }
  int pps = (int)(((M + S - 1) / S + 15) / 16 * 16);
  if (Cout % 128 == 0)
    conv2d_wgrad_f32_kernel<8><<<dim3(tiles, S), 256, 0, stream>>>(in, in_ld, dout, dout_ld, (float*)workspace, (int)M, H, W,
                                                                   Cin, Cout, KS, pps);
  else
    conv2d_wgrad_f32_kernel<4><<<dim3(tiles, S), 256, 0, stream>>>(in, in_ld, dout, dout_ld, (float*)workspace, (int)M, H, W,
                                                                   Cin, Cout, KS, pps);
  TSR_CHECK_LAUNCH("conv2d_wgrad_f32");
  long long n = (long long)taps * Cin * Cout;
  wgrad_reduce_kernel<<<(int)((n + 31) / 32), 256, 0, stream>>>((const float*)workspace, dw_oihw, S, taps, Cin,
                                                                Cout, accumulate);
  TSR_CHECK_LAUNCH("wgrad_reduce");
  return TSR_OK;
}

// level-1 blocks: 1024 rows each for big tensors (<= 1024 blocks), but never fewer than ~4 CTAs per SM worth of blocks
// when the tensor is short and wide (the MLP bias gradients: 8192 x 1024), down to 16 rows per block
static int colsum_blocks(long long npix) {
  int nb = tsr_cdiv(npix, 1024);
  int want = tsr_cdiv(npix, 16);
  if (want > 592) want = 592;
  if (nb < want) nb = want;
  if (nb > 1024) nb = 1024;
  return nb;
}

size_t tsr_colsum_workspace(long long npix, int C) { return (size_t)colsum_blocks(npix) * C * sizeof(float); }

int tsr_colsum(const void* x, int ld, int x_bf16, long long npix, int C, float* out, void* workspace, size_t ws_bytes,
               int accumulate, cudaStream_t stream) {
  TSR_REQUIRE(x && out && workspace, "colsum: null pointer");
  TSR_REQUIRE(C % 4 == 0 && C <= 1024 && ld % 4 == 0, "colsum: C must be a multiple of 4 and <= 1024");
  int nb = colsum_blocks(npix);
  int rpb = tsr_cdiv(npix, nb);
  nb = tsr_cdiv(npix, rpb);
  TSR_REQUIRE(ws_bytes >= (size_t)nb * C * sizeof(float), "colsum: workspace too small");
  int q4 = C / 4;
  int threads = (256 / q4) * q4;
  if (threads < q4) threads = q4;
  TSR_DISPATCH_T(x_bf16, T, colsum_partial_kernel<T><<<nb, threads, threads * sizeof(float4), stream>>>((const T*)x, ld, (int)npix, C, (float*)workspace, rpb));
  TSR_CHECK_LAUNCH("colsum_partial");
  colsum_final_kernel<<<tsr_cdiv(C, 32), dim3(32, 32), 0, stream>>>((const float*)workspace, nb, C, out, accumulate);
  TSR_CHECK_LAUNCH("colsum_final");
  return TSR_OK;
}

static size_t head_smem(int sf) { return (size_t)(48 + (4 * sf + 2) * (4 * sf + 2) * 4) * sizeof(float); }

int tsr_head_fwd(const float* x, long long x_bstride, const float* w_oihw, void* out, int out_ld, int out_bf16,
                 int B, int sf, int relu, cudaStream_t stream) {
  TSR_REQUIRE(x && w_oihw && out, "head_fwd: null pointer");
  TSR_REQUIRE(sf >= 1 && sf <= 24, "head_fwd: scale_factor %d unsupported (1..24)", sf);
  TSR_REQUIRE(out_ld % 4 == 0, "head_fwd: out_ld must be a multiple of 4");
  size_t smem = head_smem(sf);
  int grid = B < 148 ? B : 148;
  TSR_DISPATCH_T(out_bf16, T,
                 TSR_CUDA(cudaFuncSetAttribute(head_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                 head_fwd_kernel<T><<<grid, 256, smem, stream>>>(x, x_bstride, w_oihw, (T*)out, out_ld, B, sf, relu));
  TSR_CHECK_LAUNCH("head_fwd");
  return TSR_OK;
}

size_t tsr_head_wgrad_workspace(int B) {
  int grid = B < 296 ? B : 296;
  return (size_t)grid * 27 * 64 * sizeof(float);
}

int tsr_head_wgrad(const float* x, long long x_bstride, const void* dout, int dout_ld, int dout_bf16,
                   float* dw_oihw, void* workspace, size_t ws_bytes, int B, int sf, int accumulate,
                   cudaStream_t stream) {
  TSR_REQUIRE(x && dout && dw_oihw && workspace, "head_wgrad: null pointer");
  TSR_REQUIRE(sf >= 1 && sf <= 24, "head_wgrad: scale_factor %d unsupported", sf);
  int grid = B < 296 ? B : 296;
  TSR_REQUIRE(ws_bytes >= (size_t)grid * 27 * 64 * sizeof(float), "head_wgrad: workspace too small");
  size_t up_floats = (size_t)(4 * sf + 2) * (4 * sf + 2) * 4;
  if (up_floats < 2048) up_floats = 2048;        // the lane reduction reuses it: <= 32 lanes x 64 channels
  TSR_REQUIRE(dout_ld % 8 == 0, "head_wgrad: dout row stride must be a multiple of 8 elements (16-byte copies)");
  TSR_DISPATCH_T(dout_bf16, T,
                 // staged chunk: up to 40 KB, less when a large scale factor's upsampled frame leaves less shared memory
                 const long long room = 227LL * 1024 - (long long)(48 + up_floats) * 4;
                 long long chpix = room / (2 * HeadWg<T>::ROWB);
                 if (chpix > HeadWg<T>::CH_MAX) chpix = HeadWg<T>::CH_MAX;
                 TSR_REQUIRE(chpix >= 16, "head_wgrad: scale_factor %d leaves no shared memory for the gradient stage", sf);
                 size_t smem = (48 + up_floats) * sizeof(float) + (size_t)2 * chpix * HeadWg<T>::ROWB;
                 TSR_CUDA(cudaFuncSetAttribute(head_wgrad_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                 head_wgrad_kernel<T><<<grid, HEAD_WG_NT, smem, stream>>>(x, x_bstride, (const T*)dout, dout_ld, (float*)workspace, B, sf,
                                                                          (int)up_floats, (int)chpix));
  TSR_CHECK_LAUNCH("head_wgrad");
  head_wgrad_reduce_kernel<<<tsr_cdiv(27 * 64, 256), 256, 0, stream>>>((const float*)workspace, grid, dw_oihw, accumulate);
  TSR_CHECK_LAUNCH("head_wgrad_reduce");
  return TSR_OK;
}

int tsr_tail_fwd(const void* in, int in_ld, int in_bf16, const float* w_oihw, float* out, int B, int H, int W,
                 int Cin, int relu, cudaStream_t stream) {
  TSR_REQUIRE(in && w_oihw && out, "tail_fwd: null pointer");
  TSR_REQUIRE(Cin % 4 == 0 && Cin <= 1024 && in_ld % 4 == 0, "tail_fwd: Cin must be a multiple of 4");
  TSR_REQUIRE(Cin % 8 == 0, "tail_fwd: Cin must be a multiple of 8");
  TSR_REQUIRE(256 % (Cin / 8) == 0 && 256 % (Cin / 4) == 0, "tail_fwd: Cin / 4 must divide 256 (got Cin = %d)", Cin);
  const int strips = (H + TAIL_TR - 1) / TAIL_TR;
  const int grid = B * strips;
  TSR_DISPATCH_T(in_bf16, T,
                 size_t smem = (size_t)(TAIL_TR + 2) * (W + 2) * (Cin + 16 / sizeof(T)) * sizeof(T);
                 TSR_REQUIRE(smem <= 227 * 1024, "tail_fwd: tile does not fit in shared memory");
                 if (Cin <= 128 && ((uintptr_t)in & 15) == 0 && (in_ld * sizeof(T)) % 16 == 0) {
                   // two tiles and one persistent CTA per SM when they fit (16-bit inputs), else one strip per CTA
                   const int nbuf = 2 * smem <= 227 * 1024 ? 2 : 1;
                   const int g2 = nbuf == 2 ? (grid < 148 ? grid : 148) : grid;
                   TSR_CUDA(cudaFuncSetAttribute(tail_fwd_blocked_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(nbuf * smem)));
                   // persistent form: 16 warps copy, the first W / 5 of them compute; one-strip form: 8 warps, 2-3 CTAs per SM
                   tail_fwd_blocked_kernel<T><<<g2, nbuf == 2 ? 512 : 256, nbuf * smem, stream>>>((const T*)in, in_ld, w_oihw, out, B, H, W, Cin, relu, nbuf);
                 } else {
                   TSR_CUDA(cudaFuncSetAttribute(tail_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                   tail_fwd_kernel<T><<<grid, 256, smem, stream>>>((const T*)in, in_ld, w_oihw, out, B, H, W, Cin, relu);
                 });
  TSR_CHECK_LAUNCH("tail_fwd");
  return TSR_OK;
}

static int tail_dgrad_impl(const float* dout, const float* out_act, const float* w_oihw, void* din, int din_ld,
                           int din_bf16, int B, int H, int W, int Cin, int relu, const void* in_act, int in_ld, int in_dtype,
                           cudaStream_t stream);

int tsr_tail_dgrad(const float* dout, const float* out_act, const float* w_oihw, void* din, int din_ld,
                   int din_bf16, int B, int H, int W, int Cin, int relu, cudaStream_t stream) {
  return tail_dgrad_impl(dout, out_act, w_oihw, din, din_ld, din_bf16, B, H, W, Cin, relu, nullptr, 0, 0, stream);
}

// tsr_tail_dgrad that also applies the ReLU backward of the layer that PRODUCED the tail's input (in_act: its stored 16-bit
// activation, in_dtype 1 = bf16 / 2 = fp16): din = dgrad * [in_act > 0] -- saves a separate tsr_relu_backward pass
int tsr_tail_dgrad_masked(const float* dout, const float* out_act, const float* w_oihw, void* din, int din_ld, int din_bf16,
                          int B, int H, int W, int Cin, int relu, const void* in_act, int in_ld, int in_dtype,
                          cudaStream_t stream) {
  TSR_REQUIRE(in_act && (in_dtype == TSR_DT_BF16 || in_dtype == TSR_DT_F16) && in_ld % 8 == 0 && ((uintptr_t)in_act & 15) == 0,
              "tail_dgrad_masked: needs a 16-byte aligned 16-bit activation");
  return tail_dgrad_impl(dout, out_act, w_oihw, din, din_ld, din_bf16, B, H, W, Cin, relu, in_act, in_ld, in_dtype, stream);
}

static int tail_dgrad_impl(const float* dout, const float* out_act, const float* w_oihw, void* din, int din_ld,
                           int din_bf16, int B, int H, int W, int Cin, int relu, const void* in_act, int in_ld, int in_dtype,
                           cudaStream_t stream) {
  TSR_REQUIRE(dout && w_oihw && din && (!relu || out_act), "tail_dgrad: null pointer");
  TSR_REQUIRE(Cin % 4 == 0 && din_ld % 4 == 0, "tail_dgrad: Cin must be a multiple of 4");
  TSR_REQUIRE(Cin % 8 == 0, "tail_dgrad: Cin must be a multiple of 8");
  const int strips = (H + TAIL_TR - 1) / TAIL_TR;
  const size_t smem = (size_t)(TAIL_TR + 2) * (W + 2) * sizeof(float);
  TSR_DISPATCH_T(din_bf16, T, tail_dgrad_kernel<T><<<(B * strips < 148 * 6 ? B * strips : 148 * 6), 256, smem, stream>>>(dout, out_act, w_oihw, (T*)din, din_ld, B, H, W, Cin, relu, in_act, in_ld, in_dtype == TSR_DT_F16));
  TSR_CHECK_LAUNCH("tail_dgrad");
  return TSR_OK;
}

static int tail_wgrad_blocks(long long M) {
  int nb = tsr_cdiv(M, 512);
  if (nb > 592) nb = 592;
  return nb;
}
size_t tsr_tail_wgrad_workspace(int B, int H, int W, int Cin) {
  return (size_t)tail_wgrad_blocks((long long)B * H * W) * 9 * Cin * sizeof(float);
}

int tsr_tail_wgrad(const void* in, int in_ld, int in_bf16, const float* dout, const float* out_act,
                   float* dw_oihw, void* workspace, size_t ws_bytes, int B, int H, int W, int Cin, int relu,
                   int accumulate, cudaStream_t stream) {
  TSR_REQUIRE(in && dout && dw_oihw && workspace && (!relu || out_act), "tail_wgrad: null pointer");
  TSR_REQUIRE(Cin % 4 == 0 && Cin <= 1024 && in_ld % 4 == 0, "tail_wgrad: Cin must be a multiple of 4");
  long long M = (long long)B * H * W;
  int nb = tail_wgrad_blocks(M);
  int ppb = tsr_cdiv(M, nb);
  nb = tsr_cdiv(M, ppb);
  TSR_REQUIRE(ws_bytes >= (size_t)nb * 9 * Cin * sizeof(float), "tail_wgrad: workspace too small");
  int q4 = Cin / 4;
  int threads = (256 / q4) * q4;
  if (threads < q4) threads = q4;
  size_t smem = (size_t)threads * sizeof(float4);
  TSR_DISPATCH_T(in_bf16, T, tail_wgrad_kernel<T><<<nb, threads, smem, stream>>>((const T*)in, in_ld, dout, out_act, (float*)workspace, (int)M, H, W, Cin, relu, ppb));
  TSR_CHECK_LAUNCH("tail_wgrad");
  tail_wgrad_reduce_kernel<<<tsr_cdiv(9 * Cin, 32), 256, 0, stream>>>((const float*)workspace, nb, Cin, dw_oihw, accumulate);
  TSR_CHECK_LAUNCH("tail_wgrad_reduce");
  return TSR_OK;
}

}  // extern "C"
