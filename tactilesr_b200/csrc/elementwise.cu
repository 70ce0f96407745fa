// HBM-bound kernels of the SR path: BatchNorm (train/eval, forward/backward), ReLU backward,
// fused MSE loss + HR-label preparation, fused Adam, layout conversion.
// All take NHWC activations [pixel][ld] with C channels starting at the given pointer; storage is
// fp32 (fp32 mode) or bf16 (tensor-core mode).  Loads/stores are 4 channels wide, reductions are
// two-level (per-block partials in a fixed order, then one finalising block) => deterministic.
//
// Replaces nn.BatchNorm2d / nn.ReLU / nn.MSELoss / F.interpolate / optim.Adam at reference
// model/tactileSR_model.py:38-39,42-43,48-49,169-170,..., train/tactileSR_train.py:39,44-45,49,212.
#include "common.cuh"

namespace {

// Second-level reduction shared by the finalising kernels: block = (FIN_CH channels, FIN_LANES partial-row lanes), i.e. few
// channels and many row lanes per block, so that the <= 592 partial rows are two or three loads per thread and the kernel is
// not one long dependent load chain; returns, for threadIdx.y == 0, the fixed-order double sums over all level-1
// partial rows of quantities 0 and 1 of channel c (fixed-order tree => deterministic).
constexpr int FIN_CH = 8, FIN_LANES = 128;
__device__ __forceinline__ void reduce_partials2(const float* __restrict__ partial, int nblocks, int C, int c,
                                                 double& s0, double& s1, int ld = 0) {
  __shared__ double sh[2][FIN_LANES][FIN_CH + 1];
  double a = 0.0, b = 0.0;
  if (ld <= 0) ld = C;            // row pitch of the table (a channel slice of a wider table: ld > C)
  if (c < C)
    for (int r = threadIdx.y; r < nblocks; r += FIN_LANES) {
      a += (double)partial[((long long)r * 2 + 0) * ld + c];
      b += (double)partial[((long long)r * 2 + 1) * ld + c];
    }
  sh[0][threadIdx.y][threadIdx.x] = a;
  sh[1][threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  for (int o = FIN_LANES / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.y < o) {
      sh[0][threadIdx.y][threadIdx.x] += sh[0][threadIdx.y + o][threadIdx.x];
      sh[1][threadIdx.y][threadIdx.x] += sh[1][threadIdx.y + o][threadIdx.x];
    }
    __syncthreads();
  }
  s0 = sh[0][0][threadIdx.x];
  s1 = sh[1][0][threadIdx.x];
}

// ------------------------------------------------------------------------------------------
// BatchNorm statistics: partial[blk][0][C] = sum x, partial[blk][1][C] = sum x^2
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void bn_stats_partial_kernel(const T* __restrict__ x, int ld, int npix, int C,
                                        float* __restrict__ partial, int rows_per_block) {
  extern __shared__ float4 sm[];
  const int q4 = C / 4, lanes = blockDim.x / q4;
  const int q = threadIdx.x % q4, lane = threadIdx.x / q4;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, npix);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), ss = s;
  for (int r = r0 + lane; r < r1; r += lanes) {
    float4 v = ld4(x + (long long)r * ld + q * 4);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    ss.x = fmaf(v.x, v.x, ss.x); ss.y = fmaf(v.y, v.y, ss.y); ss.z = fmaf(v.z, v.z, ss.z); ss.w = fmaf(v.w, v.w, ss.w);
  }
  sm[threadIdx.x] = s;
  sm[blockDim.x + threadIdx.x] = ss;
  __syncthreads();
  if (lane == 0) {
    for (int l = 1; l < lanes; ++l) {
      float4 a = sm[l * q4 + q], b = sm[blockDim.x + l * q4 + q];
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
      ss.x += b.x; ss.y += b.y; ss.z += b.z; ss.w += b.w;
    }
    *reinterpret_cast<float4*>(partial + ((long long)blockIdx.x * 2 + 0) * C + q * 4) = s;
    *reinterpret_cast<float4*>(partial + ((long long)blockIdx.x * 2 + 1) * C + q * 4) = ss;
  }
}

// finalise: batch mean / biased var -> scale, shift, saved mean / invstd; running-stat update with
// momentum and the unbiased variance (nn.BatchNorm2d semantics), num_batches_tracked += 1.
__global__ void bn_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, double n,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   long long* __restrict__ nbt, float momentum, float eps,
                                   float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ save_mean, float* __restrict__ save_invstd, int part_ld = 0) {
  const int c = blockIdx.x * FIN_CH + threadIdx.x;
  if (c == 0 && threadIdx.y == 0 && nbt) *nbt += 1;
  double s, ss;
  reduce_partials2(partial, nblocks, C, c, s, ss, part_ld);
  if (c >= C || threadIdx.y != 0) return;
  double mean = s / n;
  double var = ss / n - mean * mean;
  if (var < 0.0) var = 0.0;
  double invstd = 1.0 / sqrt(var + (double)eps);
  float sc = (float)((double)gamma[c] * invstd);
  scale[c] = sc;
  shift[c] = (float)((double)beta[c] - mean * (double)gamma[c] * invstd);
  save_mean[c] = (float)mean;
  save_invstd[c] = (float)invstd;
  if (running_mean) {
    double unbiased = n > 1.0 ? var * (n / (n - 1.0)) : var;
    running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mean);
    running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unbiased);
  }
}

// eval mode: scale/shift from the running statistics
__global__ void bn_eval_coeffs_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                      float eps, float* __restrict__ scale, float* __restrict__ shift,
                                      float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float invstd = 1.0f / sqrtf(running_var[c] + eps);
  float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - running_mean[c] * sc;
  save_mean[c] = running_mean[c];
  save_invstd[c] = invstd;
}

template <typename Tin, typename Tout>
__global__ void bn_apply_kernel(const Tin* __restrict__ x, int x_ld, const float* __restrict__ scale,
                                const float* __restrict__ shift, Tout* __restrict__ out, int out_ld,
                                long long npix, int C, int relu) {
  const int q4 = C / 4;
  long long total = npix * q4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int q = (int)(i % q4);
    long long p = i / q4;
    float4 v = ld4(x + p * x_ld + q * 4);
    float4 sc = *reinterpret_cast<const float4*>(scale + q * 4);
    float4 sh = *reinterpret_cast<const float4*>(shift + q * 4);
    v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    st4(out + p * out_ld + q * 4, v);
  }
}

// ------------------------------------------------------------------------------------------
// BatchNorm backward.  z = y*scale + shift, a = relu(z).  Given da:
//   dz = da * [z > 0],  xhat = (y - mean) * invstd
//   level 1: partial sums of dz and dz*xhat;  finalise: dgamma, dbeta, c1 = sum dz / n, c2 = sum dz*xhat / n
//   apply:   dy = gamma*invstd * (dz - c1 - xhat*c2)           (train)   |   dy = dz * scale  (eval)
// ------------------------------------------------------------------------------------------
template <typename Tg, typename Ty>
__global__ void bn_bwd_partial_kernel(const Tg* __restrict__ da, int da_ld, const Ty* __restrict__ y, int y_ld,
                                      const float* __restrict__ scale, const float* __restrict__ shift,
                                      const float* __restrict__ mean, const float* __restrict__ invstd,
                                      int npix, int C, int relu, float* __restrict__ partial, int rows_per_block) {
  extern __shared__ float4 sm[];
  const int q4 = C / 4, lanes = blockDim.x / q4;
  const int q = threadIdx.x % q4, lane = threadIdx.x / q4;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, npix);
  const float4 sc = *reinterpret_cast<const float4*>(scale + q * 4);
  const float4 sh = *reinterpret_cast<const float4*>(shift + q * 4);
  const float4 mu = *reinterpret_cast<const float4*>(mean + q * 4);
  const float4 is = *reinterpret_cast<const float4*>(invstd + q * 4);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), sx = s;
  for (int r = r0 + lane; r < r1; r += lanes) {
    float4 g = ld4(da + (long long)r * da_ld + q * 4);
    float4 v = ld4(y + (long long)r * y_ld + q * 4);
    if (relu) {
      if (!(fmaf(v.x, sc.x, sh.x) > 0.f)) g.x = 0.f;
      if (!(fmaf(v.y, sc.y, sh.y) > 0.f)) g.y = 0.f;
      if (!(fmaf(v.z, sc.z, sh.z) > 0.f)) g.z = 0.f;
      if (!(fmaf(v.w, sc.w, sh.w) > 0.f)) g.w = 0.f;
    }
    s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
    sx.x = fmaf(g.x, (v.x - mu.x) * is.x, sx.x); sx.y = fmaf(g.y, (v.y - mu.y) * is.y, sx.y);
    sx.z = fmaf(g.z, (v.z - mu.z) * is.z, sx.z); sx.w = fmaf(g.w, (v.w - mu.w) * is.w, sx.w);
  }
  sm[threadIdx.x] = s;
  sm[blockDim.x + threadIdx.x] = sx;
  __syncthreads();
  if (lane == 0) {
    for (int l = 1; l < lanes; ++l) {
      float4 a = sm[l * q4 + q], b = sm[blockDim.x + l * q4 + q];
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
      sx.x += b.x; sx.y += b.y; sx.z += b.z; sx.w += b.w;
    }
    *reinterpret_cast<float4*>(partial + ((long long)blockIdx.x * 2 + 0) * C + q * 4) = s;
    *reinterpret_cast<float4*>(partial + ((long long)blockIdx.x * 2 + 1) * C + q * 4) = sx;
  }
}

__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, double n,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate,
                                       float* __restrict__ c1, float* __restrict__ c2, int training) {
  const int c = blockIdx.x * FIN_CH + threadIdx.x;
  double s, sx;
  reduce_partials2(partial, nblocks, C, c, s, sx);
  if (c >= C || threadIdx.y != 0) return;
  if (dgamma) dgamma[c] = accumulate ? dgamma[c] + (float)sx : (float)sx;
  if (dbeta) dbeta[c] = accumulate ? dbeta[c] + (float)s : (float)s;
  c1[c] = training ? (float)(s / n) : 0.f;
  c2[c] = training ? (float)(sx / n) : 0.f;
}

// level 2 of the BatchNorm backward when level 1 came out of a data-gradient epilogue (conv_tc2.cu): the table holds
// sum g and sum g*y (y = the pre-normalisation tensor), so sum g*xhat = invstd * (sum g*y - mean * sum g)
__global__ void bn_bwd_finalize_gy_kernel(const float* __restrict__ partial, int part_ld, int nblocks, int C, double n,
                                          const float* __restrict__ mean, const float* __restrict__ invstd,
                                          float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate,
                                          float* __restrict__ c1, float* __restrict__ c2, int training) {
  const int c = blockIdx.x * FIN_CH + threadIdx.x;
  double s, sy;
  reduce_partials2(partial, nblocks, C, c, s, sy, part_ld);
  if (c >= C || threadIdx.y != 0) return;
  const double sx = (double)invstd[c] * (sy - (double)mean[c] * s);
  if (dgamma) dgamma[c] = accumulate ? dgamma[c] + (float)sx : (float)sx;
  if (dbeta) dbeta[c] = accumulate ? dbeta[c] + (float)s : (float)s;
  c1[c] = training ? (float)(s / n) : 0.f;
  c2[c] = training ? (float)(sx / n) : 0.f;
}

template <typename Tg, typename Ty, typename To>
__global__ void bn_bwd_apply_kernel(const Tg* __restrict__ da, int da_ld, const Ty* __restrict__ y, int y_ld,
                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ c1, const float* __restrict__ c2,
                                    To* __restrict__ dy, int dy_ld, long long npix, int C, int relu) {
  const int q4 = C / 4;
  long long total = npix * q4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int q = (int)(i % q4);
    long long p = i / q4;
    float4 g = ld4(da + p * da_ld + q * 4);
    float4 v = ld4(y + p * y_ld + q * 4);
    const float4 sc = *reinterpret_cast<const float4*>(scale + q * 4);
    const float4 sh = *reinterpret_cast<const float4*>(shift + q * 4);
    const float4 mu = *reinterpret_cast<const float4*>(mean + q * 4);
    const float4 is = *reinterpret_cast<const float4*>(invstd + q * 4);
    const float4 k1 = *reinterpret_cast<const float4*>(c1 + q * 4);
    const float4 k2 = *reinterpret_cast<const float4*>(c2 + q * 4);
    if (relu) {
      if (!(fmaf(v.x, sc.x, sh.x) > 0.f)) g.x = 0.f;
      if (!(fmaf(v.y, sc.y, sh.y) > 0.f)) g.y = 0.f;
      if (!(fmaf(v.z, sc.z, sh.z) > 0.f)) g.z = 0.f;
      if (!(fmaf(v.w, sc.w, sh.w) > 0.f)) g.w = 0.f;
    }
    float4 o;
    o.x = sc.x * (g.x - k1.x - (v.x - mu.x) * is.x * k2.x);
    o.y = sc.y * (g.y - k1.y - (v.y - mu.y) * is.y * k2.y);
    o.z = sc.z * (g.z - k1.z - (v.z - mu.z) * is.z * k2.z);
    o.w = sc.w * (g.w - k1.w - (v.w - mu.w) * is.w * k2.w);
    st4(dy + p * dy_ld + q * 4, o);
  }
}

// dz = da * [a > 0]   (ReLU fused into a conv epilogue: a is the stored activation)
template <typename Tg, typename Ta, typename To>
__global__ void relu_bwd_kernel(const Tg* __restrict__ da, int da_ld, const Ta* __restrict__ a, int a_ld,
                                To* __restrict__ dz, int dz_ld, long long npix, int C) {
  const int q4 = C / 4;
  long long total = npix * q4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int q = (int)(i % q4);
    long long p = i / q4;
    float4 g = ld4(da + p * da_ld + q * 4);
    float4 v = ld4(a + p * a_ld + q * 4);
    if (!(v.x > 0.f)) g.x = 0.f;
    if (!(v.y > 0.f)) g.y = 0.f;
    if (!(v.z > 0.f)) g.z = 0.f;
    if (!(v.w > 0.f)) g.w = 0.f;
    st4(dz + p * dz_ld + q * 4, g);
  }
}

// strided copy / dtype conversion of channel slices (NHWC), also used for fp32 <-> bf16
template <typename Ti, typename To>
__global__ void copy_kernel(const Ti* __restrict__ x, int x_ld, To* __restrict__ out, int out_ld, long long npix, int C) {
  const int q4 = C / 4;
  long long total = npix * q4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int q = (int)(i % q4);
    long long p = i / q4;
    st4(out + p * out_ld + q * 4, ld4(x + p * x_ld + q * 4));
  }
}

// NCHW fp32 <-> NHWC (fp32 / bf16); small generic kernels for the module boundary
template <typename To>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, To* __restrict__ out, int out_ld, int B, int C, int HW) {
  long long total = (long long)B * C * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long r = i / C;
    int p = (int)(r % HW);
    int b = (int)(r / HW);
    stf(out + ((long long)b * HW + p) * out_ld + c, x[((long long)b * C + c) * HW + p]);
  }
}
template <typename Ti>
__global__ void nhwc_to_nchw_kernel(const Ti* __restrict__ x, int x_ld, float* __restrict__ out, int B, int C, int HW) {
  long long total = (long long)B * C * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int p = (int)(i % HW);
    long long r = i / HW;
    int c = (int)(r % C);
    int b = (int)(r / C);
    out[i] = ldf(x + ((long long)b * HW + p) * x_ld + c);
  }
}

// ------------------------------------------------------------------------------------------
// fused MSE loss + HR label preparation (train/tactileSR_train.py:44-45,49):
//   HR = bilinear_resize(HR_raw / scale_num, (H, W));  loss = mean((out - HR)^2);  dout = 2 (out - HR) / N
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilinear_src(int d, float scale, int n_in, int& i0, int& i1, float& l1) {
  float s = fmaxf(scale * (d + 0.5f) - 0.5f, 0.f);
  i0 = min((int)s, n_in - 1);
  i1 = min(i0 + 1, n_in - 1);
  l1 = s - i0;
}

__global__ void mse_hr_kernel(const float* __restrict__ out, const float* __restrict__ hr_raw, float inv_scale_num,
                              int B, int H, int W, int Hin, int Win, float* __restrict__ dout, float grad_scale,
                              float* __restrict__ partial) {
  const long long total = (long long)B * H * W;
  const float sy = (float)Hin / (float)H, sx = (float)Win / (float)W;
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int x = (int)(i % W);
    long long r = i / W;
    int y = (int)(r % H);
    int b = (int)(r / H);
    float hr;
    if (Hin == H && Win == W) {
      hr = hr_raw[i] * inv_scale_num;
    } else {
      int y0, y1, x0, x1;
      float ly, lx;
      bilinear_src(y, sy, Hin, y0, y1, ly);
      bilinear_src(x, sx, Win, x0, x1, lx);
      const float* src = hr_raw + (long long)b * Hin * Win;
      float v00 = src[y0 * Win + x0] * inv_scale_num, v01 = src[y0 * Win + x1] * inv_scale_num;
      float v10 = src[y1 * Win + x0] * inv_scale_num, v11 = src[y1 * Win + x1] * inv_scale_num;
      hr = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
    }
    float d = out[i] - hr;
    acc = fmaf(d, d, acc);
    if (dout) dout[i] = d * grad_scale;
  }
  __shared__ float red[32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
  }
}
// ------------------------------------------------------------------------------------------
// evaluation metrics of eval_func (reference train/tactileSR_train.py:76-94, utility/tools.py:49-81), one block per
// sample, HR label preparation (HR / scale_num, bilinear resize) fused as in mse_hr_kernel:
//   sqerr = sum (out - HR)^2;   PSNR = 10 log10(max^2 / (sqerr / psnr_div));
//   SSIM  = ((2 mu1 mu2 + C1)(2 s12 + C2)) / ((mu1^2 + mu2^2 + C1)(s1 + s2 + C2))  with whole-image statistics
// (psnr_div is the reference's `pattern.shape[0] * pattern.shape[1]` of the (1, H, W) slice it passes, i.e. H, not H*W)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
eval_metrics_kernel(const float* __restrict__ out, const float* __restrict__ hr_raw, float inv_scale_num, int H, int W,
                    int Hin, int Win, float max_value, float psnr_div, float c1, float c2, float* __restrict__ sqerr,
                    float* __restrict__ psnr, float* __restrict__ ssim) {
  const int b = blockIdx.x;
  const int npx = H * W;
  const float sy = (float)Hin / (float)H, sx = (float)Win / (float)W;
  const float* o = out + (long long)b * npx;
  const float* src = hr_raw + (long long)b * Hin * Win;
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};     // d^2, o, h, o^2, h^2, o h
  for (int i = threadIdx.x; i < npx; i += blockDim.x) {
    const int x = i % W, y = i / W;
    float hr;
    if (Hin == H && Win == W) {
      hr = src[i] * inv_scale_num;
    } else {
      int y0, y1, x0, x1;
      float ly, lx;
      bilinear_src(y, sy, Hin, y0, y1, ly);
      bilinear_src(x, sx, Win, x0, x1, lx);
      const float v00 = src[y0 * Win + x0] * inv_scale_num, v01 = src[y0 * Win + x1] * inv_scale_num;
      const float v10 = src[y1 * Win + x0] * inv_scale_num, v11 = src[y1 * Win + x1] * inv_scale_num;
      hr = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
    }
    const float ov = o[i], d = ov - hr;
    acc[0] = fmaf(d, d, acc[0]); acc[1] += ov; acc[2] += hr;
    acc[3] = fmaf(ov, ov, acc[3]); acc[4] = fmaf(hr, hr, acc[4]); acc[5] = fmaf(ov, hr, acc[5]);
  }
  __shared__ float red[6][8];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const float v = warp_sum(acc[k]);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[6];
    for (int k = 0; k < 6; ++k) {
      double a = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += (double)red[k][w];
      t[k] = a;
    }
    const double n = (double)npx;
    sqerr[b] = (float)t[0];
    psnr[b] = (float)(10.0 * log10((double)max_value * (double)max_value / (t[0] / (double)psnr_div)));
    const double mu1 = t[1] / n, mu2 = t[2] / n;
    const double s1 = t[3] / n - mu1 * mu1, s2 = t[4] / n - mu2 * mu2, s12 = t[5] / n - mu1 * mu2;
    ssim[b] = (float)(((2.0 * mu1 * mu2 + c1) * (2.0 * s12 + c2)) / ((mu1 * mu1 + mu2 * mu2 + c1) * (s1 + s2 + c2)));
  }
}

__global__ void sum_final_kernel(const float* __restrict__ partial, int n, double mul, float* __restrict__ out) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += (double)partial[i];
    *out = (float)(s * mul);
  }
}

// ------------------------------------------------------------------------------------------
// fused Adam (torch.optim.Adam, coupled L2 weight decay, no amsgrad) over a flat fp32 buffer
// ------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float wd,
                            float inv_bc1, float inv_sqrt_bc2, float grad_scale, const float* __restrict__ hyper) {
  // hyper (optional, device memory): {lr / (1 - beta1^t), 1 / sqrt(1 - beta2^t)} of this step -- lets a captured CUDA
  // graph replay the launch while the host advances the step count and the learning-rate schedule
  if (hyper) { lr = hyper[0]; inv_bc1 = 1.f; inv_sqrt_bc2 = hyper[1]; }
  long long i4 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4;
  for (; i4 < n; i4 += (long long)gridDim.x * blockDim.x * 4) {
    if (i4 + 4 <= n) {
      float4 pp = *reinterpret_cast<float4*>(p + i4), gg = *reinterpret_cast<const float4*>(g + i4);
      float4 mm = *reinterpret_cast<float4*>(m + i4), vv = *reinterpret_cast<float4*>(v + i4);
      float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float gr = fmaf(wd, pa[k], ga[k] * grad_scale);
        ma[k] = b1 * ma[k] + (1.f - b1) * gr;
        va[k] = b2 * va[k] + (1.f - b2) * gr * gr;
        float denom = sqrtf(va[k]) * inv_sqrt_bc2 + eps;
        pa[k] -= (lr * inv_bc1) * (ma[k] / denom);
      }
      *reinterpret_cast<float4*>(p + i4) = pp;
      *reinterpret_cast<float4*>(m + i4) = mm;
      *reinterpret_cast<float4*>(v + i4) = vv;
    } else {
      for (long long i = i4; i < n; ++i) {
        float gr = fmaf(wd, p[i], g[i] * grad_scale);
        m[i] = b1 * m[i] + (1.f - b1) * gr;
        v[i] = b2 * v[i] + (1.f - b2) * gr * gr;
        float denom = sqrtf(v[i]) * inv_sqrt_bc2 + eps;
        p[i] -= (lr * inv_bc1) * (m[i] / denom);
      }
    }
  }
}


// ------------------------------------------------------------------------------------------
// 16-byte-per-thread variants used when every activation operand has the same storage type
// (the engine's case): 8 bf16 / 4 fp32 channels per access.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
bn_stats_partial_v_kernel(const T* __restrict__ x, int ld, int npix, int C, float* __restrict__ partial,
                          int rows_per_block) {
  constexpr int N = V16<T>::N;
  extern __shared__ float smf[];
  const int qn = C / N, lanes = blockDim.x / qn;
  const int q = threadIdx.x % qn, lane = threadIdx.x / qn;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, npix);
  float s[N], ss[N];
#pragma unroll
  for (int i = 0; i < N; ++i) { s[i] = 0.f; ss[i] = 0.f; }
  if (lane < lanes) {
    int r = r0 + lane;
    for (; r + 3 * lanes < r1; r += 4 * lanes) {       // 4 independent 16-byte loads in flight per thread
      float v0[N], v1[N], v2[N], v3[N];
      V16<T>::load(x + (long long)r * ld + q * N, v0);
      V16<T>::load(x + (long long)(r + lanes) * ld + q * N, v1);
      V16<T>::load(x + (long long)(r + 2 * lanes) * ld + q * N, v2);
      V16<T>::load(x + (long long)(r + 3 * lanes) * ld + q * N, v3);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        s[i] += (v0[i] + v1[i]) + (v2[i] + v3[i]);
        ss[i] = fmaf(v0[i], v0[i], fmaf(v1[i], v1[i], fmaf(v2[i], v2[i], fmaf(v3[i], v3[i], ss[i]))));
      }
    }
    for (; r < r1; r += lanes) {
      float v[N];
      V16<T>::load(x + (long long)r * ld + q * N, v);
#pragma unroll
      for (int i = 0; i < N; ++i) { s[i] += v[i]; ss[i] = fmaf(v[i], v[i], ss[i]); }
    }
  }
  // smem layout: [lane][2][C]
  if (lane < lanes) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      smf[(lane * 2 + 0) * C + q * N + i] = s[i];
      smf[(lane * 2 + 1) * C + q * N + i] = ss[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    float a = 0.f;
    for (int l = 0; l < lanes; ++l) a += smf[l * 2 * C + c];
    partial[(long long)blockIdx.x * 2 * C + c] = a;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_apply_v_kernel(const T* __restrict__ x, int x_ld, const float* __restrict__ scale, const float* __restrict__ shift,
                  T* __restrict__ out, int out_ld, long long npix, int C, int relu, __nv_bfloat16* __restrict__ out2,
                  int out2_ld, int* __restrict__ ovf = nullptr) {
  // each thread owns one 16-byte channel vector (fixed q) and strides over pixels: coefficients live in registers.
  // out2 (optional, 16-bit T only): a second, bf16 copy of the result -- the weight-gradient operand of the "fp16" mode
  constexpr int N = V16<T>::N;
  const int qn = C / N;                       // divides 256
  const int q = threadIdx.x % qn;
  float sc[N], sh[N];
#pragma unroll
  for (int k = 0; k < N; ++k) { sc[k] = scale[q * N + k]; sh[k] = shift[q * N + k]; }
  const long long pstride = (long long)gridDim.x * (256 / qn);
  for (long long p = (long long)blockIdx.x * (256 / qn) + threadIdx.x / qn; p < npix; p += pstride) {
    float v[N];
    V16<T>::load(x + p * x_ld + q * N, v);
#pragma unroll
    for (int k = 0; k < N; ++k) {
      v[k] = fmaf(v[k], sc[k], sh[k]);
      if (relu) v[k] = fmaxf(v[k], 0.f);
    }
    V16<T>::store(out + p * out_ld + q * N, v);
    if constexpr (N == 8) {
      if (out2) V16<__nv_bfloat16>::store(out2 + p * out2_ld + q * N, v);
      if (ovf) {             // fp16 output: sticky overflow guard
        float m = 0.f;
        bool bad = false;
#pragma unroll
        for (int k = 0; k < N; ++k) { m = fmaxf(m, fabsf(v[k])); bad |= (v[k] != v[k]); }
        if (bad || !(m <= 65504.f)) *ovf = 1;
      }
    }
  }
}

template <typename T, typename Ty = T>
__global__ void __launch_bounds__(256)
bn_bwd_partial_v_kernel(const T* __restrict__ da, int da_ld, const Ty* __restrict__ y, int y_ld,
                        const float* __restrict__ scale, const float* __restrict__ shift,
                        const float* __restrict__ mean, const float* __restrict__ invstd, int npix, int C, int relu,
                        float* __restrict__ partial, int rows_per_block) {
  constexpr int N = V16<T>::N;
  extern __shared__ float smf[];
  const int qn = C / N, lanes = blockDim.x / qn;
  const int q = threadIdx.x % qn, lane = threadIdx.x / qn;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, npix);
  float sc[N], sh[N], mu[N], is[N], s[N], sx[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    sc[i] = scale[q * N + i]; sh[i] = shift[q * N + i]; mu[i] = mean[q * N + i]; is[i] = invstd[q * N + i];
    s[i] = 0.f; sx[i] = 0.f;
  }
  if (lane < lanes) {
    int r = r0 + lane;
    for (; r + lanes < r1; r += 2 * lanes) {            // 4 independent 16-byte loads in flight per thread
      float g0[N], v0[N], g1[N], v1[N];
      V16<T>::load(da + (long long)r * da_ld + q * N, g0);
      V16<Ty>::load(y + (long long)r * y_ld + q * N, v0);
      V16<T>::load(da + (long long)(r + lanes) * da_ld + q * N, g1);
      V16<Ty>::load(y + (long long)(r + lanes) * y_ld + q * N, v1);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        if (relu && !(fmaf(v0[i], sc[i], sh[i]) > 0.f)) g0[i] = 0.f;
        if (relu && !(fmaf(v1[i], sc[i], sh[i]) > 0.f)) g1[i] = 0.f;
        s[i] += g0[i] + g1[i];
        sx[i] = fmaf(g0[i], (v0[i] - mu[i]) * is[i], fmaf(g1[i], (v1[i] - mu[i]) * is[i], sx[i]));
      }
    }
    for (; r < r1; r += lanes) {
      float g[N], v[N];
      V16<T>::load(da + (long long)r * da_ld + q * N, g);
      V16<Ty>::load(y + (long long)r * y_ld + q * N, v);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        if (relu && !(fmaf(v[i], sc[i], sh[i]) > 0.f)) g[i] = 0.f;
        s[i] += g[i];
        sx[i] = fmaf(g[i], (v[i] - mu[i]) * is[i], sx[i]);
      }
    }
  }
  if (lane < lanes) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      smf[(lane * 2 + 0) * C + q * N + i] = s[i];
      smf[(lane * 2 + 1) * C + q * N + i] = sx[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    float a = 0.f;
    for (int l = 0; l < lanes; ++l) a += smf[l * 2 * C + c];
    partial[(long long)blockIdx.x * 2 * C + c] = a;
  }
}

template <typename T, typename Ty = T>
__global__ void __launch_bounds__(256)
bn_bwd_apply_v_kernel(const T* __restrict__ da, int da_ld, const Ty* __restrict__ y, int y_ld,
                      const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                      const float* __restrict__ invstd, const float* __restrict__ c1, const float* __restrict__ c2,
                      T* __restrict__ dy, int dy_ld, long long npix, int C, int relu) {
  // dy = sc*(g - c1 - (v - mu)*is*c2) = sc*g + kb*v + kc with per-channel constants held in registers
  constexpr int N = V16<T>::N;
  const int qn = C / N;
  const int q = threadIdx.x % qn;
  float sc[N], sh[N], kb[N], kc[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const int c = q * N + k;
    sc[k] = scale[c]; sh[k] = shift[c];
    const float t = sc[k] * invstd[c] * c2[c];
    kb[k] = -t;
    kc[k] = fmaf(t, mean[c], -sc[k] * c1[c]);
  }
  const long long pstride = (long long)gridDim.x * (256 / qn);
  for (long long p = (long long)blockIdx.x * (256 / qn) + threadIdx.x / qn; p < npix; p += pstride) {
    float g[N], v[N], o[N];
    V16<T>::load(da + p * da_ld + q * N, g);
    V16<Ty>::load(y + p * y_ld + q * N, v);
#pragma unroll
    for (int k = 0; k < N; ++k) {
      if (relu && !(fmaf(v[k], sc[k], sh[k]) > 0.f)) g[k] = 0.f;
      o[k] = fmaf(sc[k], g[k], fmaf(kb[k], v[k], kc[k]));
    }
    V16<T>::store(dy + p * dy_ld + q * N, o);
  }
}

template <typename T, typename Ta = T>
__global__ void __launch_bounds__(256)
relu_bwd_v_kernel(const T* __restrict__ da, int da_ld, const Ta* __restrict__ a, int a_ld, T* __restrict__ dz, int dz_ld,
                  long long npix, int C) {
  constexpr int N = V16<T>::N;
  const int qn = C / N;
  const long long total = npix * qn;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % qn);
    const long long p = i / qn;
    float g[N], v[N];
    V16<T>::load(da + p * da_ld + q * N, g);
    V16<Ta>::load(a + p * a_ld + q * N, v);
#pragma unroll
    for (int k = 0; k < N; ++k)
      if (!(v[k] > 0.f)) g[k] = 0.f;
    V16<T>::store(dz + p * dz_ld + q * N, g);
  }
}

inline int redv_threads(int C, int N) {
  int qn = C / N;
  int t = (256 / qn) * qn;
  return t < qn ? qn : t;
}

inline int ew_grid(long long total) {
  long long g = (total + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  if (g < 1) g = 1;
  return (int)g;
}
inline int red_blocks(long long npix, int& rows_per_block) {
  int nb = tsr_cdiv(npix, 512);
  if (nb > 296) nb = 296;
  if (nb < 1) nb = 1;
  rows_per_block = tsr_cdiv(npix, nb);
  return tsr_cdiv(npix, rows_per_block);
}
inline int red_threads(int C) {
  int q4 = C / 4;
  int t = (256 / q4) * q4;
  return t < q4 ? q4 : t;
}

}  // namespace

extern "C" {

size_t tsr_bn_workspace(long long npix, int C) {
  int rpb;
  return (size_t)red_blocks(npix, rpb) * 2 * C * sizeof(float);
}

// training-mode statistics of y -> scale/shift (+ saved mean/invstd), running stats updated in place.
int tsr_bn_train_stats(const void* y, int y_ld, int y_bf16, long long npix, int C, const float* gamma,
                       const float* beta, float* running_mean, float* running_var, long long* num_batches_tracked,
                       float momentum, float eps, float* scale, float* shift, float* save_mean, float* save_invstd,
                       void* workspace, size_t ws_bytes, cudaStream_t stream) {
  TSR_REQUIRE(y && gamma && beta && scale && shift && save_mean && save_invstd && workspace, "bn_train_stats: null pointer");
  TSR_REQUIRE(C % 4 == 0 && C <= 1024 && y_ld % 4 == 0, "bn_train_stats: C must be a multiple of 4 and <= 1024");
  TSR_REQUIRE(npix > 0 && npix < (1ll << 31), "bn_train_stats: bad pixel count");
  int rpb, nb = red_blocks(npix, rpb), th = red_threads(C);
  TSR_REQUIRE(ws_bytes >= (size_t)nb * 2 * C * sizeof(float), "bn_train_stats: workspace too small");
  if (y_bf16 && C % 8 == 0 && y_ld % 8 == 0) {
    int tv = redv_threads(C, 8);
    size_t smv = (size_t)(tv / (C / 8)) * 2 * C * sizeof(float);
    if (y_bf16 == TSR_DT_F16)
      bn_stats_partial_v_kernel<__half><<<nb, tv, smv, stream>>>((const __half*)y, y_ld, (int)npix, C, (float*)workspace, rpb);
    else
      bn_stats_partial_v_kernel<__nv_bfloat16><<<nb, tv, smv, stream>>>((const __nv_bfloat16*)y, y_ld, (int)npix, C, (float*)workspace, rpb);
  } else if (!y_bf16) {
    int tv = redv_threads(C, 4);
    size_t smv = (size_t)(tv / (C / 4)) * 2 * C * sizeof(float);
    bn_stats_partial_v_kernel<float><<<nb, tv, smv, stream>>>((const float*)y, y_ld, (int)npix, C, (float*)workspace, rpb);
  } else {
    size_t smem = (size_t)2 * th * sizeof(float4);
    TSR_DISPATCH_T(y_bf16, T, bn_stats_partial_kernel<T><<<nb, th, smem, stream>>>((const T*)y, y_ld, (int)npix, C, (float*)workspace, rpb));
  }
  TSR_CHECK_LAUNCH("bn_stats_partial");
  bn_finalize_kernel<<<tsr_cdiv(C, FIN_CH), dim3(FIN_CH, FIN_LANES), 0, stream>>>((const float*)workspace, nb, C, (double)npix, gamma, beta,
                                                           running_mean, running_var, num_batches_tracked, momentum,
                                                           eps, scale, shift, save_mean, save_invstd);
  TSR_CHECK_LAUNCH("bn_finalize");
  return TSR_OK;
}

// finalise batch statistics from partial sums some other kernel produced (the tensor-core conv epilogue):
// partial[nrows][2][C] (sum, sum of squares) -> scale / shift / saved mean / invstd, running statistics updated.
int tsr_bn_finalize_partials(const float* partial, int part_ld, int nrows, long long npix, int C, const float* gamma,
                             const float* beta, float* running_mean, float* running_var, long long* num_batches_tracked,
                             float momentum, float eps, float* scale, float* shift, float* save_mean, float* save_invstd,
                             cudaStream_t stream) {
  TSR_REQUIRE(partial && gamma && beta && scale && shift && save_mean && save_invstd && nrows > 0 && npix > 0 && part_ld >= C,
              "bn_finalize_partials: bad argument");
  bn_finalize_kernel<<<tsr_cdiv(C, FIN_CH), dim3(FIN_CH, FIN_LANES), 0, stream>>>(partial, nrows, C, (double)npix, gamma, beta, running_mean,
                                                           running_var, num_batches_tracked, momentum, eps, scale, shift,
                                                           save_mean, save_invstd, part_ld);
  TSR_CHECK_LAUNCH("bn_finalize");
  return TSR_OK;
}

int tsr_bn_eval_coeffs(int C, const float* gamma, const float* beta, const float* running_mean,
                       const float* running_var, float eps, float* scale, float* shift, float* save_mean,
                       float* save_invstd, cudaStream_t stream) {
  TSR_REQUIRE(gamma && beta && running_mean && running_var && scale && shift && save_mean && save_invstd, "bn_eval_coeffs: null pointer");
  bn_eval_coeffs_kernel<<<tsr_cdiv(C, 128), 128, 0, stream>>>(C, gamma, beta, running_mean, running_var, eps, scale, shift, save_mean, save_invstd);
  TSR_CHECK_LAUNCH("bn_eval_coeffs");
  return TSR_OK;
}

#define TSR_DISPATCH2(A_BF16, B_BF16, KERNEL, GRID, BLOCK, SMEM, ...)                                   \
  do {                                                                                                  \
    if (A_BF16) {                                                                                       \
      if (B_BF16) KERNEL<__nv_bfloat16, __nv_bfloat16><<<GRID, BLOCK, SMEM, stream>>>(__VA_ARGS__);     \
      else KERNEL<__nv_bfloat16, float><<<GRID, BLOCK, SMEM, stream>>>(__VA_ARGS__);                    \
    } else {                                                                                            \
      if (B_BF16) KERNEL<float, __nv_bfloat16><<<GRID, BLOCK, SMEM, stream>>>(__VA_ARGS__);             \
      else KERNEL<float, float><<<GRID, BLOCK, SMEM, stream>>>(__VA_ARGS__);                            \
    }                                                                                                   \
  } while (0)

int tsr_bn_apply(const void* y, int y_ld, int y_bf16, const float* scale, const float* shift, void* out, int out_ld,
                 int out_bf16, long long npix, int C, int relu, void* out2_bf16, int out2_ld, cudaStream_t stream) {
  TSR_REQUIRE(y && scale && shift && out, "bn_apply: null pointer");
  TSR_REQUIRE(C % 4 == 0 && y_ld % 4 == 0 && out_ld % 4 == 0, "bn_apply: C and strides must be multiples of 4");
  int grid = ew_grid(npix * (C / 4));
  __nv_bfloat16* o2 = (__nv_bfloat16*)out2_bf16;
#define ARGS(Ti, To) (const Ti*)y, y_ld, scale, shift, (To*)out, out_ld, npix, C, relu
  if (y_bf16 && y_bf16 == out_bf16 && C % 8 == 0 && 256 % (C / 8) == 0 && y_ld % 8 == 0 && out_ld % 8 == 0 &&
      (!o2 || out2_ld % 8 == 0)) {
    if (y_bf16 == TSR_DT_F16) bn_apply_v_kernel<__half><<<ew_grid(npix * (C / 8)), 256, 0, stream>>>(ARGS(__half, __half), o2, out2_ld, tsr_f16_overflow_ptr());
    else bn_apply_v_kernel<__nv_bfloat16><<<ew_grid(npix * (C / 8)), 256, 0, stream>>>(ARGS(__nv_bfloat16, __nv_bfloat16), o2, out2_ld);
  } else {
    TSR_REQUIRE(!o2, "bn_apply: the second (bf16) output needs 16-bit operands with C %% 8 == 0");
    if (!y_bf16 && !out_bf16 && 256 % (C / 4) == 0) {
      bn_apply_v_kernel<float><<<grid, 256, 0, stream>>>(ARGS(float, float), nullptr, 0);
    } else {
      TSR_DISPATCH_T(y_bf16, Ti, TSR_DISPATCH_T(out_bf16, To, bn_apply_kernel<Ti, To><<<grid, 256, 0, stream>>>(ARGS(Ti, To))));
    }
  }
#undef ARGS
  TSR_CHECK_LAUNCH("bn_apply");
  return TSR_OK;
}

// BatchNorm(+ReLU) backward: da -> dy, dgamma, dbeta.  act_bf16 applies to da, y and dy alike.
int tsr_bn_backward(const void* da, int da_ld, const void* y, int y_ld, void* dy, int dy_ld, int act_bf16,
                    const float* scale, const float* shift, const float* save_mean, const float* save_invstd,
                    float* dgamma, float* dbeta, int accumulate, long long npix, int C, int relu, int training,
                    void* workspace, size_t ws_bytes, cudaStream_t stream) {
  TSR_REQUIRE(da && y && dy && scale && shift && save_mean && save_invstd && workspace, "bn_backward: null pointer");
  TSR_REQUIRE(C % 4 == 0 && C <= 1024 && da_ld % 4 == 0 && y_ld % 4 == 0 && dy_ld % 4 == 0, "bn_backward: C and strides must be multiples of 4");
  int rpb, nb = red_blocks(npix, rpb), th = red_threads(C);
  size_t need = (size_t)nb * 2 * C * sizeof(float) + 2 * (size_t)C * sizeof(float);
  TSR_REQUIRE(ws_bytes >= need, "bn_backward: workspace too small");
  float* partial = (float*)workspace;
  float* c1 = partial + (size_t)nb * 2 * C;
  float* c2 = c1 + C;
  // act_bf16: 0 = everything fp32, 1 = everything bf16, 2 = gradients (da, dy) bf16 + saved activation y fp16
  const bool v8 = act_bf16 && C % 8 == 0 && 256 % (C / 8) == 0 && da_ld % 8 == 0 && y_ld % 8 == 0 && dy_ld % 8 == 0;
  const bool v4 = !act_bf16 && 256 % (C / 4) == 0;
  typedef __nv_bfloat16 bf;
  if (v8) {
    int tv = redv_threads(C, 8);
    size_t smv = (size_t)(tv / (C / 8)) * 2 * C * sizeof(float);
    if (act_bf16 == TSR_DT_F16)
      bn_bwd_partial_v_kernel<bf, __half><<<nb, tv, smv, stream>>>((const bf*)da, da_ld, (const __half*)y, y_ld, scale, shift, save_mean, save_invstd, (int)npix, C, relu, partial, rpb);
    else
      bn_bwd_partial_v_kernel<bf, bf><<<nb, tv, smv, stream>>>((const bf*)da, da_ld, (const bf*)y, y_ld, scale, shift, save_mean, save_invstd, (int)npix, C, relu, partial, rpb);
  } else if (v4) {
    int tv = redv_threads(C, 4);
    size_t smv = (size_t)(tv / (C / 4)) * 2 * C * sizeof(float);
    bn_bwd_partial_v_kernel<float><<<nb, tv, smv, stream>>>((const float*)da, da_ld, (const float*)y, y_ld, scale, shift, save_mean, save_invstd, (int)npix, C, relu, partial, rpb);
  } else {
    size_t smem = (size_t)2 * th * sizeof(float4);
    if (act_bf16 == TSR_DT_F16)
      bn_bwd_partial_kernel<bf, __half><<<nb, th, smem, stream>>>((const bf*)da, da_ld, (const __half*)y, y_ld, scale, shift, save_mean, save_invstd, (int)npix, C, relu, partial, rpb);
    else if (act_bf16)
      bn_bwd_partial_kernel<bf, bf><<<nb, th, smem, stream>>>((const bf*)da, da_ld, (const bf*)y, y_ld, scale, shift, save_mean, save_invstd, (int)npix, C, relu, partial, rpb);
    else
      bn_bwd_partial_kernel<float, float><<<nb, th, smem, stream>>>((const float*)da, da_ld, (const float*)y, y_ld, scale, shift, save_mean, save_invstd, (int)npix, C, relu, partial, rpb);
  }
  TSR_CHECK_LAUNCH("bn_bwd_partial");
  bn_bwd_finalize_kernel<<<tsr_cdiv(C, FIN_CH), dim3(FIN_CH, FIN_LANES), 0, stream>>>(partial, nb, C, (double)npix, dgamma, dbeta, accumulate, c1, c2, training);
  TSR_CHECK_LAUNCH("bn_bwd_finalize");
  int grid = ew_grid(npix * (C / 4));
  if (v8 && act_bf16 == TSR_DT_F16)
    bn_bwd_apply_v_kernel<bf, __half><<<ew_grid(npix * (C / 8)), 256, 0, stream>>>((const bf*)da, da_ld, (const __half*)y, y_ld, scale, shift, save_mean, save_invstd, c1, c2, (bf*)dy, dy_ld, npix, C, relu);
  else if (v8)
    bn_bwd_apply_v_kernel<bf, bf><<<ew_grid(npix * (C / 8)), 256, 0, stream>>>((const bf*)da, da_ld, (const bf*)y, y_ld, scale, shift, save_mean, save_invstd, c1, c2, (bf*)dy, dy_ld, npix, C, relu);
  else if (v4)
    bn_bwd_apply_v_kernel<float><<<grid, 256, 0, stream>>>((const float*)da, da_ld, (const float*)y, y_ld, scale, shift, save_mean, save_invstd, c1, c2, (float*)dy, dy_ld, npix, C, relu);
  else if (act_bf16 == TSR_DT_F16)
    bn_bwd_apply_kernel<bf, __half, bf><<<grid, 256, 0, stream>>>((const bf*)da, da_ld, (const __half*)y, y_ld, scale, shift, save_mean, save_invstd, c1, c2, (bf*)dy, dy_ld, npix, C, relu);
  else if (act_bf16)
    bn_bwd_apply_kernel<bf, bf, bf><<<grid, 256, 0, stream>>>((const bf*)da, da_ld, (const bf*)y, y_ld, scale, shift, save_mean, save_invstd, c1, c2, (bf*)dy, dy_ld, npix, C, relu);
  else
    bn_bwd_apply_kernel<float, float, float><<<grid, 256, 0, stream>>>((const float*)da, da_ld, (const float*)y, y_ld, scale, shift, save_mean, save_invstd, c1, c2, (float*)dy, dy_ld, npix, C, relu);
  TSR_CHECK_LAUNCH("bn_bwd_apply");
  return TSR_OK;
}

size_t tsr_bn_backward_workspace(long long npix, int C) {
  int rpb;
  return (size_t)red_blocks(npix, rpb) * 2 * C * sizeof(float) + 2 * (size_t)C * sizeof(float);
}

// BatchNorm backward, level 2, from the partial table a data-gradient epilogue produced (tsr_conv2d_tc2 with
// TSR_TC2_BNB): partial[nrows][2][part_ld] = (sum g, sum g*y) with g already masked by the ReLU.  Writes dgamma / dbeta
// and the two per-channel constants c1 = sum g / n, c2 = sum g*xhat / n that tsr_bn_backward_apply consumes.
int tsr_bn_bwd_finalize_partials(const float* partial, int part_ld, int nrows, long long npix, int C, const float* save_mean,
                                 const float* save_invstd, float* dgamma, float* dbeta, int accumulate, float* c1, float* c2,
                                 int training, cudaStream_t stream) {
  TSR_REQUIRE(partial && save_mean && save_invstd && c1 && c2 && nrows > 0 && npix > 0 && part_ld >= C,
              "bn_bwd_finalize_partials: bad argument");
  bn_bwd_finalize_gy_kernel<<<tsr_cdiv(C, FIN_CH), dim3(FIN_CH, FIN_LANES), 0, stream>>>(partial, part_ld, nrows, C, (double)npix, save_mean,
                                                                                        save_invstd, dgamma, dbeta, accumulate, c1, c2, training);
  TSR_CHECK_LAUNCH("bn_bwd_finalize_gy");
  return TSR_OK;
}

// BatchNorm backward, level 3 alone: dy = scale * (g - c1 - xhat * c2) with g = da (relu = 0: already masked by the
// producer's epilogue) or da * [scale*y + shift > 0] (relu = 1).  act_bf16 as in tsr_bn_backward (1 / 2: 16-bit operands).
int tsr_bn_backward_apply(const void* da, int da_ld, const void* y, int y_ld, void* dy, int dy_ld, int act_bf16,
                          const float* scale, const float* shift, const float* save_mean, const float* save_invstd,
                          const float* c1, const float* c2, long long npix, int C, int relu, cudaStream_t stream) {
  TSR_REQUIRE(da && y && dy && scale && shift && save_mean && save_invstd && c1 && c2, "bn_backward_apply: null pointer");
  TSR_REQUIRE(act_bf16 == TSR_DT_BF16 || act_bf16 == TSR_DT_F16, "bn_backward_apply: 16-bit operands only");
  TSR_REQUIRE(C % 8 == 0 && 256 % (C / 8) == 0 && da_ld % 8 == 0 && y_ld % 8 == 0 && dy_ld % 8 == 0,
              "bn_backward_apply: C must be 8 * (a divisor of 256), strides multiples of 8");
  typedef __nv_bfloat16 bf;
  if (act_bf16 == TSR_DT_F16)
    bn_bwd_apply_v_kernel<bf, __half><<<ew_grid(npix * (C / 8)), 256, 0, stream>>>((const bf*)da, da_ld, (const __half*)y, y_ld, scale, shift, save_mean, save_invstd, c1, c2, (bf*)dy, dy_ld, npix, C, relu);
  else
    bn_bwd_apply_v_kernel<bf, bf><<<ew_grid(npix * (C / 8)), 256, 0, stream>>>((const bf*)da, da_ld, (const bf*)y, y_ld, scale, shift, save_mean, save_invstd, c1, c2, (bf*)dy, dy_ld, npix, C, relu);
  TSR_CHECK_LAUNCH("bn_bwd_apply");
  return TSR_OK;
}

int tsr_relu_backward(const void* da, int da_ld, const void* a, int a_ld, void* dz, int dz_ld, int act_bf16,
                      long long npix, int C, cudaStream_t stream) {
  TSR_REQUIRE(da && a && dz, "relu_backward: null pointer");
  TSR_REQUIRE(C % 4 == 0 && da_ld % 4 == 0 && a_ld % 4 == 0 && dz_ld % 4 == 0, "relu_backward: C and strides must be multiples of 4");
  int grid = ew_grid(npix * (C / 4));
  // act_bf16: 0 = fp32, 1 = bf16, 2 = gradients (da, dz) bf16 + stored activation a fp16
  typedef __nv_bfloat16 bf;
  const bool v8 = act_bf16 && C % 8 == 0 && da_ld % 8 == 0 && a_ld % 8 == 0 && dz_ld % 8 == 0;
  if (v8 && act_bf16 == TSR_DT_F16)
    relu_bwd_v_kernel<bf, __half><<<ew_grid(npix * (C / 8)), 256, 0, stream>>>((const bf*)da, da_ld, (const __half*)a, a_ld, (bf*)dz, dz_ld, npix, C);
  else if (v8)
    relu_bwd_v_kernel<bf, bf><<<ew_grid(npix * (C / 8)), 256, 0, stream>>>((const bf*)da, da_ld, (const bf*)a, a_ld, (bf*)dz, dz_ld, npix, C);
  else if (act_bf16 == TSR_DT_F16)
    relu_bwd_kernel<bf, __half, bf><<<grid, 256, 0, stream>>>((const bf*)da, da_ld, (const __half*)a, a_ld, (bf*)dz, dz_ld, npix, C);
  else if (act_bf16)
    relu_bwd_kernel<bf, bf, bf><<<grid, 256, 0, stream>>>((const bf*)da, da_ld, (const bf*)a, a_ld, (bf*)dz, dz_ld, npix, C);
  else
    relu_bwd_v_kernel<float><<<grid, 256, 0, stream>>>((const float*)da, da_ld, (const float*)a, a_ld, (float*)dz, dz_ld, npix, C);
  TSR_CHECK_LAUNCH("relu_backward");
  return TSR_OK;
}

int tsr_copy_channels(const void* x, int x_ld, int x_bf16, void* out, int out_ld, int out_bf16, long long npix, int C,
                      cudaStream_t stream) {
  TSR_REQUIRE(x && out, "copy_channels: null pointer");
  TSR_REQUIRE(C % 4 == 0 && x_ld % 4 == 0 && out_ld % 4 == 0, "copy_channels: C and strides must be multiples of 4");
  int grid = ew_grid(npix * (C / 4));
  TSR_DISPATCH_T(x_bf16, Ti, TSR_DISPATCH_T(out_bf16, To, copy_kernel<Ti, To><<<grid, 256, 0, stream>>>((const Ti*)x, x_ld, (To*)out, out_ld, npix, C)));
  TSR_CHECK_LAUNCH("copy_channels");
  return TSR_OK;
}

int tsr_nchw_to_nhwc(const float* x, void* out, int out_ld, int out_bf16, int B, int C, int HW, cudaStream_t stream) {
  TSR_REQUIRE(x && out, "nchw_to_nhwc: null pointer");
  int grid = ew_grid((long long)B * C * HW);
  TSR_DISPATCH_T(out_bf16, T, nchw_to_nhwc_kernel<T><<<grid, 256, 0, stream>>>(x, (T*)out, out_ld, B, C, HW));
  TSR_CHECK_LAUNCH("nchw_to_nhwc");
  return TSR_OK;
}

int tsr_nhwc_to_nchw(const void* x, int x_ld, int x_bf16, float* out, int B, int C, int HW, cudaStream_t stream) {
  TSR_REQUIRE(x && out, "nhwc_to_nchw: null pointer");
  int grid = ew_grid((long long)B * C * HW);
  TSR_DISPATCH_T(x_bf16, T, nhwc_to_nchw_kernel<T><<<grid, 256, 0, stream>>>((const T*)x, x_ld, out, B, C, HW));
  TSR_CHECK_LAUNCH("nhwc_to_nchw");
  return TSR_OK;
}

size_t tsr_mse_hr_workspace(void) { return 1024 * sizeof(float); }

// loss = mean((out - resize(hr_raw / scale_num))^2) -> *loss (device scalar); dout = 2 (out - HR) / N * grad_mul
int tsr_mse_hr_loss(const float* out, const float* hr_raw, float scale_num, int B, int H, int W, int Hin, int Win,
                    float* loss, float* dout, float grad_mul, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  TSR_REQUIRE(out && hr_raw && loss && workspace, "mse_hr_loss: null pointer");
  TSR_REQUIRE(ws_bytes >= 1024 * sizeof(float), "mse_hr_loss: workspace too small");
  long long N = (long long)B * H * W;
  int grid = (int)((N + 255) / 256);
  if (grid > 1024) grid = 1024;
  mse_hr_kernel<<<grid, 256, 0, stream>>>(out, hr_raw, 1.0f / scale_num, B, H, W, Hin, Win, dout,
                                          (float)(2.0 / (double)N) * grad_mul, (float*)workspace);
  TSR_CHECK_LAUNCH("mse_hr");
  sum_final_kernel<<<1, 32, 0, stream>>>((const float*)workspace, grid, 1.0 / (double)N, loss);
  TSR_CHECK_LAUNCH("sum_final");
  return TSR_OK;
}

// per-sample evaluation metrics (eval_func): out (B,1,H,W) fp32, hr_raw (B,1,Hin,Win) fp32 -> sqerr[B], psnr[B], ssim[B]
int tsr_eval_metrics(const float* out, const float* hr_raw, float scale_num, int B, int H, int W, int Hin, int Win,
                     float max_value, float psnr_div, float c1, float c2, float* sqerr, float* psnr, float* ssim,
                     cudaStream_t stream) {
  TSR_REQUIRE(out && hr_raw && sqerr && psnr && ssim && B > 0, "eval_metrics: bad argument");
  eval_metrics_kernel<<<B, 256, 0, stream>>>(out, hr_raw, 1.0f / scale_num, H, W, Hin, Win, max_value, psnr_div, c1, c2,
                                             sqerr, psnr, ssim);
  TSR_CHECK_LAUNCH("eval_metrics");
  return TSR_OK;
}

int tsr_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, long long step, float grad_scale, cudaStream_t stream) {
  TSR_REQUIRE(p && g && m && v && n >= 0 && step >= 1, "adam_step: bad argument");
  if (n == 0) return TSR_OK;
  double bc1 = 1.0 - pow((double)beta1, (double)step);
  double bc2 = 1.0 - pow((double)beta2, (double)step);
  int grid = ew_grid((n + 3) / 4);
  adam_kernel<<<grid, 256, 0, stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, (float)(1.0 / bc1),
                                        (float)(1.0 / sqrt(bc2)), grad_scale, nullptr);
  TSR_CHECK_LAUNCH("adam_step");
  return TSR_OK;
}

// the same step with the step-dependent scalars read from device memory: hyper = {lr / (1 - beta1^t), 1 / sqrt(1 - beta2^t)}
int tsr_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, const float* hyper, float beta1,
                      float beta2, float eps, float weight_decay, float grad_scale, cudaStream_t stream) {
  TSR_REQUIRE(p && g && m && v && hyper && n >= 0, "adam_step_dev: bad argument");
  if (n == 0) return TSR_OK;
  int grid = ew_grid((n + 3) / 4);
  adam_kernel<<<grid, 256, 0, stream>>>(p, g, m, v, n, 0.f, beta1, beta2, eps, weight_decay, 1.f, 1.f, grad_scale, hyper);
  TSR_CHECK_LAUNCH("adam_step_dev");
  return TSR_OK;
}

}  // extern "C"
