// Tensor-core support kernels for sm_100a: weight packing into the SWIZZLE_128B shared-memory image and the weight gradient
// (wgrad_tc3_kernel: tcgen05.mma with both operands MN-major, accumulators in TMEM over the CTA's whole pixel range,
// deterministic two-level reduction).  The forward / data-gradient implicit GEMM lives in conv_tc2.cu.
//
// Replaces the cuDNN dispatch behind the weight gradient of nn.Conv2d (reference model/tactileSR_model.py:41,47,53,168,
// 174,180,186,191,219,220 through autograd, cpu/trainer.py:353).
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace {

constexpr int NUM_THREADS = 224;

// experiment switches (env TSR_TC_MODE or tsr_set_tc_desc_mode), read by the engine: bit 7 = BatchNorm statistics in a separate
// pass instead of the conv epilogue, bit 8 = no fused gradient sinks, bit 9 = no dual-branch forward, bit 10 = no bf16 shadows
// in the "fp16" mode (the weight gradient converts fp16 x tiles in shared memory instead).
static int env_mode() { const char* e = getenv("TSR_TC_MODE"); return e ? atoi(e) : 0; }
int g_desc_mode = env_mode();

// ------------------------------------------------------------------------------------------------
// weight packing: OIHW fp32 -> bf16 [chunk][tap][N rows][64] in the SWIZZLE_128B shared-memory image
//   forward: rows = co, K = ci;   dgrad: rows = ci, K = co, taps flipped
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void pack_weight_bf16_kernel(const float* __restrict__ w, T* __restrict__ wf, T* __restrict__ wd, int Cout,
                                        int Cin, int KS, const float* __restrict__ co_scale = nullptr) {
  const int taps = KS * KS;
  const long long n = (long long)Cout * Cin * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % taps);
    const long long r = i / taps;
    const int ci = (int)(r % Cin);
    const int co = (int)(r / Cin);
    T v;
    stf(&v, co_scale ? w[i] * co_scale[co] : w[i]);      // (inference: an eval-mode BatchNorm folded into the weights)
    if (wf) {   // tile (chunk = ci/64, tap t): row co, k = ci%64
      const int c = ci >> 6, k = ci & 63;
      const long long tile = ((long long)c * taps + t) * Cout * 64;
      wf[tile + (long long)co * 64 + (((k >> 3) ^ (co & 7)) << 3) + (k & 7)] = v;
    }
    if (wd) {   // tile (chunk = co/64, tap taps-1-t): row ci, k = co%64
      const int c = co >> 6, k = co & 63;
      const long long tile = ((long long)c * taps + (taps - 1 - t)) * Cin * 64;
      wd[tile + (long long)ci * 64 + (((k >> 3) ^ (ci & 7)) << 3) + (k & 7)] = v;
    }
  }
}


// All conv weights of a model in ONE launch (they all change at every optimizer step): blockIdx.y = table entry.
struct PackDesc {          // mirrors TsrPackDesc of include/tactilesr_b200.h
  const float* w;          // OIHW fp32
  void* wf;                // forward image or null
  void* wd;                // data-gradient image or null
  int Cout, Cin, KS;
  int dt_f, dt_d;          // storage codes of wf / wd: 1 = bf16, 2 = fp16
  int mode;                // 0: wf = standard forward image; 1 / 2: wf = dual-branch image (tc_ptx.cuh), this weight is its
};                         // 3x3 / 5x5 branch (Cout = 64)
static_assert(sizeof(PackDesc) == 48, "TsrPackDesc layout");

__device__ __forceinline__ void store16(void* base, long long idx, float v, int dt) {
  if (dt == TSR_DT_F16) reinterpret_cast<__half*>(base)[idx] = __float2half_rn(v);
  else reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
}

__global__ void pack_weights_multi_kernel(const PackDesc* __restrict__ table) {
  const PackDesc e = table[blockIdx.y];
  const int taps = e.KS * e.KS;
  const long long n = (long long)e.Cout * e.Cin * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % taps);
    const long long r = i / taps;
    const int ci = (int)(r % e.Cin);
    const int co = (int)(r / e.Cin);
    const float v = e.w[i];
    if (e.wf && e.mode) {
      const int c = ci >> 6, k = ci & 63;
      const int row = c * DUAL_CHUNK_ROWS + dual_image_row(e.mode == 2, t / e.KS, t % e.KS, co);
      store16(e.wf, (long long)row * 64 + (((k >> 3) ^ (row & 7)) << 3) + (k & 7), v, e.dt_f);
    } else if (e.wf) {
      const int c = ci >> 6, k = ci & 63;
      const long long tile = ((long long)c * taps + t) * e.Cout * 64;
      store16(e.wf, tile + (long long)co * 64 + (((k >> 3) ^ (co & 7)) << 3) + (k & 7), v, e.dt_f);
    }
    if (e.wd) {
      const int c = co >> 6, k = co & 63;
      const long long tile = ((long long)c * taps + (taps - 1 - t)) * e.Cin * 64;
      store16(e.wd, tile + (long long)ci * 64 + (((k >> 3) ^ (ci & 7)) << 3) + (k & 7), v, e.dt_d);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient on the tensor cores:
//   dW[tap][ci][co] = sum_pix x[pix + shift(tap)][ci] * dy[pix][co]
// with BOTH operands MN-major: a shared-memory row is one pixel's 64 channels (128 B, SWIZZLE_128B) -- exactly the image
// the forward kernel's TMA boxes produce -- so one x halo tile (8x8 pixels + halo of ONE sample, one dense TMA box)
// serves every tap through a shifted descriptor start address, and K = 16 pixels per MMA are two 8-pixel core groups one
// halo row apart.  A CTA owns a set of M-groups x (a contiguous range of pixel tiles); its accumulators stay in TMEM
// across the whole range and are written once, as a partial [split][tap][ci][co], which wgrad_tc_reduce_kernel sums in a
// fixed order (deterministic two-level reduction, same second level as the fp32 path).
// ------------------------------------------------------------------------------------------------
constexpr int WG2_MAX_STAGES = 8;

// Operand roles chosen so that no accumulator row is wasted:
// D[M = 128 rows of ci][N = Cout] += A(x window, MN-major) * B(dy, MN-major)^T.  The two 64-row atoms of the A operand
// are LBO bytes apart, and LBO is free: for Cin % 128 == 0 they are the two channel chunks of one tap (LBO = halo tile
// size); otherwise they are TWO DIFFERENT TAPS of the same chunk (LBO = byte distance of the two windows inside the one
// halo tile), so a 64-channel layer fills all 128 rows with useful work (v2 duplicated 64 rows: half the MMAs wasted).
// One accumulator of Cout columns per M-group; a CTA owns up to 512/Cout M-groups over its whole pixel range.
struct Wgrad3Params {
  float* partial;               // [output group = blockIdx.z][nsplit][taps][Cin][Cout]
  long long group_stride;       // floats between the partials of consecutive output groups (wide layers: Cout_total > 128)
  int B, H, W, KS, pad, P;
  int Cin, Cout, cochunks;
  int tap_mode;                 // 0: atoms = 2 chunks of one tap, 1: atoms = 2 taps of one chunk
  int ngroups_total;            // M-groups per chunk unit: taps (mode 0) or ceil(taps/2) (mode 1)
  int gpc;                      // M-groups per CTA (<= 512 / Cout)
  int ncta_groups, nunits;      // CTA grid.x = ncta_groups * nunits (unit = chunk pair / chunk)
  int tiles_x, tiles_per_sample, nblocks, blocks_per_split;
  int dy_stage_bytes, xtile_bytes, stage_bytes, nstages;
  int x_f16;                    // the x operand arrives as fp16 (the "fp16" precision mode's activations) and is converted to
};                              // bf16 in shared memory before the MMAs read it (kind::f16 rejects fp16 x bf16 operands)

__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_tc3_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                 const Wgrad3Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int NS = p.nstages;
  const uint32_t bar_base = base + NS * (uint32_t)p.stage_bytes;
  auto full = [&](int i) { return bar_base + 8u * i; };
  auto empty = [&](int i) { return bar_base + 8u * (WG2_MAX_STAGES + i); };
  const uint32_t tmem_full = bar_base + 8u * (2 * WG2_MAX_STAGES);
  const uint32_t tmem_slot = tmem_full + 8u;
  auto ready = [&](int i) { return bar_base + 8u * (2 * WG2_MAX_STAGES + 2 + i); };     // x tile converted (x_f16 only)
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg = blockIdx.x % p.ncta_groups, unit = blockIdx.x / p.ncta_groups;
  const int taps = p.KS * p.KS;
  const int gbase = p.ngroups_total / p.ncta_groups, grem = p.ngroups_total % p.ncta_groups;   // balanced
  const int g0 = cg * gbase + min(cg, grem);
  const int ng = gbase + (cg < grem ? 1 : 0);
  const int nxt = p.tap_mode ? 1 : 2;              // x halo tiles (chunks) per stage
  const int blk0 = blockIdx.y * p.blocks_per_split;
  const int blk1 = min(blk0 + p.blocks_per_split, p.nblocks);
  const int nblk = max(blk1 - blk0, 0);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(full(i), 1); mbar_init(empty(i), 2); }     // two MMA-issuing warps release a stage
    for (int i = 0; i < NS; ++i) mbar_init(ready(i), 4);                                // one arrival per converting warp
    mbar_init(tmem_full, 2);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx = (uint32_t)p.P * p.P * 128u * nxt + 8192u * p.cochunks;
      int st = 0;
      uint32_t ph = 1;
      int n = blk0 / p.tiles_per_sample, t = blk0 - n * p.tiles_per_sample;     // running (sample, tile) position
      int ty = t / p.tiles_x, tx_ = t - ty * p.tiles_x;
      for (int i = 0; i < nblk; ++i) {
        mbar_wait(empty(st), ph);
        mbar_expect_tx(full(st), tx);
        const int y0 = ty * 8, x0 = tx_ * 8;
        const uint32_t dy0 = base + st * (uint32_t)p.stage_bytes;
        for (int cc = 0; cc < p.cochunks; ++cc)
          tma_load_4d(dy0 + (uint32_t)cc * 8192u, &tmap_dy, (int)blockIdx.z * p.Cout + cc * 64, x0, y0, n, full(st));
        for (int h = 0; h < nxt; ++h)
          tma_load_4d(dy0 + (uint32_t)p.dy_stage_bytes + (uint32_t)h * p.xtile_bytes, &tmap_x, (unit * nxt + h) * 64,
                      x0 - p.pad, y0 - p.pad, n, full(st));
        if (++st == NS) { st = 0; ph ^= 1u; }
        if (++tx_ == p.tiles_x) { tx_ = 0; ++ty; }
        if (++t == p.tiles_per_sample) { t = 0; ty = 0; tx_ = 0; ++n; }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // TWO issuing warps (on different scheduler partitions): the single issuing lane is the critical path of the 64-channel
    // layers (32 N = 64 MMAs per 64-pixel stage, ~30 tensor cycles each); warp 1 takes the even accumulator groups, warp 2
    // the odd ones.  Groups are independent accumulators, both warps wait on the same "full" barrier and each commits its
    // own MMAs to the stage's "empty" barrier (count 2).
    const int par = warp - 1;
    const uint32_t idesc = make_idesc(128, p.Cout, 1, 1);
    const uint32_t a_hi = desc_hi((uint32_t)p.P * 128u), b_hi = desc_hi(1024u);
    const uint32_t row_units = (uint32_t)p.P * 8u;
    const uint32_t lbo_b = p.cochunks == 2 ? 8192u : 0u;
    // per-M-group descriptor low-word increments (window offset | LBO << 16), computed once: the per-stage loop of
    // the single issuing lane must stay free of integer divisions
    uint32_t gword[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      uint32_t w = 0u;
      if (g < ng) {
        if (p.tap_mode) {
          const int t1 = 2 * (g0 + g), t2 = min(t1 + 1, taps - 1);
          const uint32_t o1 = (uint32_t)((t1 / p.KS) * p.P + t1 % p.KS) * 8u;
          const uint32_t o2 = (uint32_t)((t2 / p.KS) * p.P + t2 % p.KS) * 8u;
          w = o1 | ((o2 - o1) << 16);
        } else {
          const int t1 = g0 + g;
          w = ((uint32_t)((t1 / p.KS) * p.P + t1 % p.KS) * 8u) | (((uint32_t)p.xtile_bytes >> 4) << 16);
        }
      }
      gword[g] = w;
    }
    // (the single issuing lane is the critical path for the 64-channel layers -- 32 N = 64 MMAs per 64-pixel stage: ring
    // position, barrier addresses and descriptor bases advance by adds, no division in the loop)
    uint32_t first = 0u;
    const uint32_t stage_units = (uint32_t)p.stage_bytes >> 4, row2 = 2u * row_units, Cout = (uint32_t)p.Cout;
    const uint32_t xs_lo0 = desc_lo(base + (uint32_t)p.dy_stage_bytes, 0u), b_lo0 = desc_lo(base, lbo_b);
    const uint32_t wait0 = p.x_f16 ? ready(0) : full(0);       // fp16 x: wait for the converted tile, not for the raw TMA data
    uint32_t st = 0, ph = 0, xs_lo = xs_lo0, b_lo = b_lo0, full_bar = wait0, empty_bar = empty(0);
    for (int i = 0; i < nblk; ++i) {
      mbar_wait(full_bar, ph);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          if (g < ng && (g & 1) == par) {
            const uint32_t a_lo = xs_lo + gword[g];
#pragma unroll
            for (int s = 0; s < 4; ++s)
              umma_bf16(tmem_base + g * Cout, desc_join(a_lo + s * row2, a_hi),
                        desc_join(b_lo + s * 128u, b_hi), idesc, (first | (uint32_t)s) ? 1u : 0u);
          }
        }
        umma_commit(empty_bar);
      }
      __syncwarp();
      first = 1u;
      xs_lo += stage_units; b_lo += stage_units; full_bar += 8u; empty_bar += 8u;
      if (++st == (uint32_t)NS) { st = 0; ph ^= 1u; xs_lo = xs_lo0; b_lo = b_lo0; full_bar = wait0; empty_bar = empty(0); }
    }
    if (elect_one()) umma_commit(tmem_full);
    __syncwarp();
  } else if (warp >= 3) {
    if (p.x_f16) {
      // ===== fp16 -> bf16 conversion of the x halo tile(s) of every stage, in place (element-wise: the swizzled layout is
      // untouched), by the four warps that otherwise idle until the epilogue.  The weight gradient can then read the forward
      // activations directly, without a bf16 "shadow" copy in HBM (9 MB per sample).  Measured at B = 1024: the 64-channel
      // layers pay +20 %, the 128-channel layers +60..95 % -- their MMAs (4 KB of A + 4 KB of B per 64 cycles) already take
      // the whole shared-memory bandwidth and the conversion adds 72 B/clk on top -- so the engine keeps the shadows by
      // default and uses this path only when memory matters (TSR_TC_MODE bit 10). =====
      const int ct = threadIdx.x - 96;             // 0..127
      const uint32_t xbytes = (uint32_t)p.P * p.P * 128u;
      int st = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nblk; ++i) {
        mbar_wait(full(st), ph);
        uint8_t* xs = smem_raw + (base - smem_u32(smem_raw)) + (size_t)st * p.stage_bytes + p.dy_stage_bytes;
        for (int h = 0; h < nxt; ++h) {
          uint4* v = reinterpret_cast<uint4*>(xs + (size_t)h * p.xtile_bytes);
          const uint32_t nvec = xbytes / 16u;
          // six 16-byte vectors in flight per thread (loads first, then conversions, then stores): an element-at-a-time
          // loop is a chain of shared-memory latencies and made this warp group, not the tensor pipe, the stage's pace
          for (uint32_t k0 = ct; k0 < nvec; k0 += 6u * 128u) {
            uint4 u[6];
#pragma unroll
            for (int e = 0; e < 6; ++e)
              if (k0 + e * 128u < nvec) u[e] = v[k0 + e * 128u];
#pragma unroll
            for (int e = 0; e < 6; ++e) {
              uint32_t w[4] = {u[e].x, u[e].y, u[e].z, u[e].w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
                const __nv_bfloat162 b = __floats2bfloat162_rn(f.x, f.y);
                w[j] = *reinterpret_cast<const uint32_t*>(&b);
              }
              u[e] = make_uint4(w[0], w[1], w[2], w[3]);
            }
#pragma unroll
            for (int e = 0; e < 6; ++e)
              if (k0 + e * 128u < nvec) v[k0 + e * 128u] = u[e];
          }
        }
        fence_proxy_async();                       // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(ready(st));
        if (++st == NS) { st = 0; ph ^= 1u; }
      }
    }
    const int q = warp & 3;
    const int r = q * 32 + lane;                 // accumulator row: atom = r / 64, ci_local = r % 64
    const int atom = r >> 6, cil = r & 63;
    if (nblk > 0) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
    }
    for (int g = 0; g < ng; ++g) {
      int tap, ci;
      bool valid = true;
      if (p.tap_mode) {
        const int t1 = 2 * (g0 + g);
        tap = t1 + atom;
        valid = tap < taps;                      // odd tap count: the last group's second atom is a duplicate
        ci = unit * 64 + cil;
      } else {
        tap = g0 + g;
        ci = (unit * 2 + atom) * 64 + cil;
      }
      float* dst = p.partial + (size_t)blockIdx.z * p.group_stride +
                   (((size_t)blockIdx.y * taps + (valid ? tap : 0)) * p.Cin + ci) * p.Cout;
#pragma unroll 1
      for (int j = 0; j < p.Cout / 16; ++j) {
        uint32_t v[16];
        if (nblk > 0) {
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + g * p.Cout + j * 16, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = 0u;
        }
        if (valid) {
          float4* d4 = reinterpret_cast<float4*>(dst + j * 16);
          d4[0] = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
          d4[1] = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
          d4[2] = make_float4(__uint_as_float(v[8]), __uint_as_float(v[9]), __uint_as_float(v[10]), __uint_as_float(v[11]));
          d4[3] = make_float4(__uint_as_float(v[12]), __uint_as_float(v[13]), __uint_as_float(v[14]), __uint_as_float(v[15]));
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// level 2 (shared with the fp32 path's layout): dw_oihw[g Cout + co][ci][tap] (+)= sum_s partial[g][s][tap][ci][co]
// A block = 32 consecutive outputs x 8 split lanes: warp j sums the splits j, j + 8, ... of its 32 outputs (coalesced
// loads), the eight lane sums are combined through shared memory in lane order.  Fixed order => still bit-deterministic,
// and the dependent-load chain is S / 8 long instead of S (S = 148 for the 64-channel layers: 31 us per launch, 35
// launches per step, all of it latency).
__global__ void __launch_bounds__(256)
wgrad_tc_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int S, int taps,
                       int Cin, int Cout, int accumulate, int G) {
  __shared__ float sh[8][32];
  const long long n = (long long)taps * Cin * Cout;
  const int lane = threadIdx.x & 31, j = threadIdx.x >> 5;
  long long i = (long long)blockIdx.x * 32 + lane;
  const bool ok = i < n * G;
  int g = 0;
  float s = 0.f;
  if (ok) {
    g = (int)(i / n);
    i -= g * n;
    const float* pg = partial + (long long)g * S * n + i;
#pragma unroll 4
    for (int k = j; k < S; k += 8) s += pg[(long long)k * n];
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
  const long long o = (((long long)g * Cout + co) * Cin + ci) * taps + tap;
  dw[o] = accumulate ? dw[o] + t : t;
}

}  // namespace

extern "C" {

void tsr_set_tc_desc_mode(int mode) { g_desc_mode = mode; }
int tsr_get_tc_desc_mode(void) { return g_desc_mode; }

int tsr_pack_conv_weight_bf16(const float* w_oihw, void* w_fwd, void* w_dgrad, int Cout, int Cin, int KS,
                              cudaStream_t stream) {
  TSR_REQUIRE(w_oihw && (w_fwd || w_dgrad), "pack_conv_weight_bf16: null pointer");
  TSR_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "pack_conv_weight_bf16: Cin, Cout must be multiples of 64");
  long long n = (long long)Cout * Cin * KS * KS;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 2048) blocks = 2048;
  pack_weight_bf16_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(w_oihw, (__nv_bfloat16*)w_fwd, (__nv_bfloat16*)w_dgrad, Cout, Cin, KS);
  TSR_CHECK_LAUNCH("pack_conv_weight_bf16");
  return TSR_OK;
}

// `n` table entries (TsrPackDesc, device memory) packed by one launch; max_elems = the largest Cout*Cin*KS*KS among them.
int tsr_pack_conv_weights_multi(const void* table_dev, int n, long long max_elems, cudaStream_t stream) {
  TSR_REQUIRE(table_dev && n > 0 && max_elems > 0, "pack_conv_weights_multi: bad argument");
  int bx = (int)((max_elems + 255) / 256);
  if (bx > 64) bx = 64;
  pack_weights_multi_kernel<<<dim3(bx, n), 256, 0, stream>>>((const PackDesc*)table_dev);
  TSR_CHECK_LAUNCH("pack_conv_weights_multi");
  return TSR_OK;
}

// same tile image with fp16 elements (forward weights of the "fp16" precision mode)
int tsr_pack_conv_weight_f16(const float* w_oihw, void* w_fwd, void* w_dgrad, int Cout, int Cin, int KS,
                             cudaStream_t stream) {
  TSR_REQUIRE(w_oihw && (w_fwd || w_dgrad), "pack_conv_weight_f16: null pointer");
  TSR_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "pack_conv_weight_f16: Cin, Cout must be multiples of 64");
  long long n = (long long)Cout * Cin * KS * KS;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 2048) blocks = 2048;
  pack_weight_bf16_kernel<__half><<<blocks, 256, 0, stream>>>(w_oihw, (__half*)w_fwd, (__half*)w_dgrad, Cout, Cin, KS);
  TSR_CHECK_LAUNCH("pack_conv_weight_f16");
  return TSR_OK;
}

// Inference-time folding of an eval-mode BatchNorm that follows the convolution (y = scale * conv(x) + shift, scale / shift
// from tsr_bn_eval_coeffs): forward weights scaled per output channel (dtype 1 = bf16, 2 = fp16) and the folded bias
// bias_out[co] = scale[co] * bias[co] + shift[co]  (bias may be NULL).
__global__ void fold_bias_kernel(const float* __restrict__ bias, const float* __restrict__ scale,
                                 const float* __restrict__ shift, float* __restrict__ out, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) out[c] = fmaf(scale[c], bias ? bias[c] : 0.f, shift[c]);
}

int tsr_pack_conv_weight_folded(const float* w_oihw, const float* bias, const float* scale, const float* shift, void* w_fwd,
                                float* bias_out, int Cout, int Cin, int KS, int dtype, cudaStream_t stream) {
  TSR_REQUIRE(w_oihw && scale && shift && w_fwd && bias_out, "pack_conv_weight_folded: null pointer");
  TSR_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "pack_conv_weight_folded: Cin, Cout must be multiples of 64");
  TSR_REQUIRE(dtype == TSR_DT_BF16 || dtype == TSR_DT_F16, "pack_conv_weight_folded: dtype must be 1 (bf16) or 2 (fp16)");
  long long n = (long long)Cout * Cin * KS * KS;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 2048) blocks = 2048;
  if (dtype == TSR_DT_F16)
    pack_weight_bf16_kernel<__half><<<blocks, 256, 0, stream>>>(w_oihw, (__half*)w_fwd, (__half*)nullptr, Cout, Cin, KS, scale);
  else
    pack_weight_bf16_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(w_oihw, (__nv_bfloat16*)w_fwd, (__nv_bfloat16*)nullptr, Cout, Cin, KS, scale);
  TSR_CHECK_LAUNCH("pack_conv_weight_folded");
  fold_bias_kernel<<<tsr_cdiv(Cout, 128), 128, 0, stream>>>(bias, scale, shift, bias_out, Cout);
  TSR_CHECK_LAUNCH("fold_bias");
  return TSR_OK;
}

struct Wgrad3Plan { int tap_mode, ngroups_total, gpc, ncta_groups, nunits; };
static Wgrad3Plan wgrad3_plan(int Cin, int Cout, int KS) {
  Wgrad3Plan q;
  const int taps = KS * KS;
  q.tap_mode = (Cin % 128 == 0) ? 0 : 1;
  q.ngroups_total = q.tap_mode ? (taps + 1) / 2 : taps;
  const int max_g = 512 / Cout;
  q.ncta_groups = (q.ngroups_total + max_g - 1) / max_g;
  q.gpc = (q.ngroups_total + q.ncta_groups - 1) / q.ncta_groups;
  q.nunits = q.tap_mode ? Cin / 64 : Cin / 128;
  return q;
}

// pixel-tile split: a single wave, never more CTAs than SMs
static void wgrad_plan(int B, int H, int W, int Cin, int Cout, int KS, int* nsplit, int* nblocks, int* bps, int G = 1) {
  Wgrad3Plan q = wgrad3_plan(Cin, Cout, KS);
  *nblocks = B * (H / 8) * (W / 8);
  const int ctas = q.ncta_groups * q.nunits * G;
  int s = num_sms() / ctas;
  if (s > *nblocks) s = *nblocks;
  if (s < 1) s = 1;
  *bps = tsr_cdiv(*nblocks, s);
  *nsplit = tsr_cdiv(*nblocks, *bps);
}

size_t tsr_conv2d_wgrad_tc_workspace(int B, int H, int W, int Cin, int Cout, int KS) {
  int ns, nb, bps;
  const int G = Cout > 128 ? Cout / 128 : 1;      // wider layers: groups of 128 output channels (grid.z) in one launch
  if (Cout > 128) Cout = 128;
  wgrad_plan(B, H, W, Cin, Cout, KS, &ns, &nb, &bps, G);
  return (size_t)G * ns * KS * KS * Cin * Cout * sizeof(float);
}

// dw_oihw (fp32, [Cout][Cin][KS][KS]) (+)= wgrad of the conv;  in / dout are bf16 NHWC views.
// Cout = 64, 128, or a multiple of 128 (the MLP layers of tPSFNet): groups of 128 output channels are grid.z of one launch.
static int wgrad_tc_impl(const void* in, int in_ld, int in_f16, const void* dout, int dout_ld, float* dw_oihw, void* workspace,
                         size_t ws_bytes, int B, int H, int W, int Cin, int Cout_total, int KS, int accumulate,
                         cudaStream_t stream);

int tsr_conv2d_wgrad_tc(const void* in, int in_ld, const void* dout, int dout_ld, float* dw_oihw, void* workspace,
                        size_t ws_bytes, int B, int H, int W, int Cin, int Cout_total, int KS, int accumulate,
                        cudaStream_t stream) {
  return wgrad_tc_impl(in, in_ld, 0, dout, dout_ld, dw_oihw, workspace, ws_bytes, B, H, W, Cin, Cout_total, KS, accumulate, stream);
}

// the same with the storage type of `in` given: in_dtype 1 = bf16, 2 = fp16 (converted to bf16 inside the kernel; dout
// stays bf16) -- what the "fp16" precision mode calls with its forward activations
int tsr_conv2d_wgrad_tc_x(const void* in, int in_ld, int in_dtype, const void* dout, int dout_ld, float* dw_oihw, void* workspace,
                          size_t ws_bytes, int B, int H, int W, int Cin, int Cout_total, int KS, int accumulate,
                          cudaStream_t stream) {
  TSR_REQUIRE(in_dtype == TSR_DT_BF16 || in_dtype == TSR_DT_F16, "conv2d_wgrad_tc_x: in_dtype must be 1 (bf16) or 2 (fp16)");
  return wgrad_tc_impl(in, in_ld, in_dtype == TSR_DT_F16, dout, dout_ld, dw_oihw, workspace, ws_bytes, B, H, W, Cin, Cout_total, KS,
                       accumulate, stream);
}

static int wgrad_tc_impl(const void* in, int in_ld, int in_f16, const void* dout, int dout_ld, float* dw_oihw, void* workspace,
                         size_t ws_bytes, int B, int H, int W, int Cin, int Cout_total, int KS, int accumulate,
                         cudaStream_t stream) {
  TSR_REQUIRE(in && dout && dw_oihw && workspace, "conv2d_wgrad_tc: null pointer");
  TSR_REQUIRE(Cout_total == 64 || (Cout_total > 0 && Cout_total % 128 == 0), "conv2d_wgrad_tc: Cout must be 64 or a multiple of 128 (got %d)", Cout_total);
  const int G = Cout_total > 128 ? Cout_total / 128 : 1;
  const int Cout = Cout_total > 128 ? 128 : Cout_total;
  TSR_REQUIRE(Cin % 64 == 0 && Cin > 0, "conv2d_wgrad_tc: Cin must be a multiple of 64 (got %d)", Cin);
  TSR_REQUIRE(KS == 1 || KS == 3 || KS == 5, "conv2d_wgrad_tc: kernel size %d unsupported", KS);
  TSR_REQUIRE(W % 8 == 0 && H % 8 == 0, "conv2d_wgrad_tc: H and W must be multiples of 8");
  TSR_REQUIRE(in_ld % 8 == 0 && dout_ld % 8 == 0, "conv2d_wgrad_tc: row strides must be multiples of 8");
  EncodeTiledFn enc = get_encode();
  if (!enc) { tsr_set_error("conv2d_wgrad_tc: cuTensorMapEncodeTiled unavailable"); return TSR_ERR_CUDA; }
  const int pad = KS / 2, taps = KS * KS;
  int nsplit = 1, nblocks = 0, bps = 0;
  wgrad_plan(B, H, W, Cin, Cout, KS, &nsplit, &nblocks, &bps, G);
  size_t need = (size_t)G * nsplit * taps * Cin * Cout * sizeof(float);
  if (ws_bytes < need) { tsr_set_error("conv2d_wgrad_tc: workspace too small (%zu < %zu)", ws_bytes, need); return TSR_ERR_WORKSPACE; }
  CUtensorMap tmx, tmdy;
  cuuint32_t estr[4] = {1, 1, 1, 1};
  {
    cuuint64_t gdim[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)in_ld * 2, (cuuint64_t)W * in_ld * 2, (cuuint64_t)H * W * in_ld * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(8 + 2 * pad), (cuuint32_t)(8 + 2 * pad), 1};
    CUresult r = enc(&tmx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { tsr_set_error("conv2d_wgrad_tc: tensor map (x) failed (%d)", (int)r); return TSR_ERR_CUDA; }
  }
  {
    cuuint64_t gdim[4] = {(cuuint64_t)Cout_total, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)dout_ld * 2, (cuuint64_t)W * dout_ld * 2, (cuuint64_t)H * W * dout_ld * 2};
    cuuint32_t box[4] = {64, 8, 8, 1};
    CUresult r = enc(&tmdy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(dout), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { tsr_set_error("conv2d_wgrad_tc: tensor map (dy) failed (%d)", (int)r); return TSR_ERR_CUDA; }
  }
  Wgrad3Plan pl = wgrad3_plan(Cin, Cout, KS);
  Wgrad3Params q;
  q.partial = (float*)workspace;
  q.group_stride = (long long)nsplit * taps * Cin * Cout;
  q.B = B; q.H = H; q.W = W; q.KS = KS; q.pad = pad; q.P = 8 + 2 * pad;
  q.Cin = Cin; q.Cout = Cout; q.cochunks = Cout / 64;
  q.tap_mode = pl.tap_mode; q.ngroups_total = pl.ngroups_total; q.gpc = pl.gpc;
  q.ncta_groups = pl.ncta_groups; q.nunits = pl.nunits;
  q.tiles_x = W / 8; q.tiles_per_sample = (H / 8) * (W / 8);
  q.nblocks = nblocks; q.blocks_per_split = bps;
  q.dy_stage_bytes = 8192 * q.cochunks;
  q.xtile_bytes = (q.P * q.P * 128 + 1023) & ~1023;
  q.stage_bytes = q.dy_stage_bytes + (pl.tap_mode ? 1 : 2) * q.xtile_bytes;
  int ns = (int)((SMEM_LIMIT - 1024 - 512) / (size_t)q.stage_bytes);
  if (ns > WG2_MAX_STAGES) ns = WG2_MAX_STAGES;
  q.nstages = ns;
  q.x_f16 = in_f16;
  size_t smem = 1024 + (size_t)ns * q.stage_bytes + 512;
  TSR_CUDA(cudaFuncSetAttribute(wgrad_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(q.ncta_groups * q.nunits, nsplit, G);
  wgrad_tc3_kernel<<<grid, NUM_THREADS, smem, stream>>>(tmx, tmdy, q);
  TSR_CHECK_LAUNCH("conv2d_wgrad_tc3");
  long long n = (long long)taps * Cin * Cout * G;
  wgrad_tc_reduce_kernel<<<(int)((n + 31) / 32), 256, 0, stream>>>((const float*)workspace, dw_oihw, nsplit, taps, Cin,
                                                                  Cout, accumulate, G);
  TSR_CHECK_LAUNCH("wgrad_tc_reduce");
  return TSR_OK;
}


}  // extern "C"
