// Tensor-core convolution for sm_100a: implicit GEMM on tcgen05.mma with TMEM accumulators, activation
// halo tiles staged once per CTA by TMA (zero-filled padding), weights streamed through an mbarrier ring
// by bulk copies.  bf16 operands, fp32 accumulation.
//
//   out[p][co] = epilogue( sum_{tap, ci} in[p + shift(tap)][ci] * w[tap][ci][co] )
//
// Work decomposition ("tall plane"): the B samples are stacked vertically with `pad` virtual zero rows
// after each sample (pitch Hp = H + pad), so a CTA block is simply TM/8 consecutive virtual rows x 8
// columns = T M-tiles of 128 pixels, and every tap is a pure (row, column) shift inside one shared-memory
// halo tile.  The A operand of each MMA is a *window* of that halo tile: the UMMA shared-memory descriptor
// starts at row (ky*P + kx) of the tile, 8-row core groups are P*128 B apart (P = 16-pixel pitch), so no
// im2col copy is ever materialised and each activation byte is read from L2 once per CTA instead of once
// per tap.  The B operand (weights, pre-swizzled on the host side into the SWIZZLE_128B image) is streamed
// per (channel chunk, tap) with cp.async.bulk.
//
// Warp roles (7 warps): 0 = A/TMA producer, 1 = MMA issuer + TMEM owner, 2 = B producer, 3..6 = epilogue
// (TMEM -> registers -> bias / residual / ReLU -> bf16 global stores).
//
// Replaces the cuDNN dispatch behind nn.Conv2d for the channel-heavy convs of reference
// model/tactileSR_model.py:41,47,53,168,174,180,186,191,219,220 (forward) and their data gradients.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace {

constexpr int FLAG_RELU = 1;
constexpr int FLAG_F16 = 2;          // activations / weights / residual / output are fp16 instead of bf16
constexpr int T_TILES = 2;          // M-tiles (128 pixels each) per CTA
constexpr int MAX_NA = 8;           // activation chunk slots (2 for 3x3 / 5x5, more for the bandwidth-bound 1x1)
constexpr int NUM_THREADS = 224;
constexpr int STAT_ROWS = 148 * 4;  // rows of the fused BatchNorm-statistics partials: (CTA, epilogue warp)

// experiment switches (env TSR_TC_MODE or tsr_set_tc_desc_mode): bit1 = 16-pixel halo pitch in the forward kernel
// instead of the dense TMA-box pitch, bit4 = single-CTA (cta_group::1) forward kernel also for N = 128, bit5 = the same
// for N = 64,
// bit7 = (engine) BatchNorm statistics in a separate pass instead of the conv epilogue.
static int env_mode() { const char* e = getenv("TSR_TC_MODE"); return e ? atoi(e) : 0; }
int g_desc_mode = env_mode();

struct ConvParams {
  const __nv_bfloat16* w;        // pre-swizzled [chunk][tap][Cout_total][64]
  const float* bias;             // [N] or null
  const __nv_bfloat16* residual; // [pix][res_ld] or null
  __nv_bfloat16* out;            // [pix][out_ld]
  int res_ld, out_ld;
  int B, H, W, Hp, Vtotal;       // Hp = H + pad, Vtotal = B * Hp
  int KS, pad, P, rows;          // P = smem pixel pitch of a halo row, rows = 16*T + 2*pad
  int nchunks, nxg, flags;
  int w_tile_elems;              // elements between consecutive (chunk, tap) weight tiles = Cout_total * 64
  int nblocks, nb_stages;        // CTA blocks (persistent loop), depth of the weight ring (<= MAX_NB)
  int na_slots;                  // activation chunk slots (<= MAX_NA)
  __nv_bfloat16* out2;           // optional second, bf16 copy of an fp16 output [pix][out2_ld] (weight-gradient operand of
  int out2_ld;                   // the "fp16" precision mode), else null
  float* bn_partial;             // optional [STAT_ROWS][2][bn_C] per-(CTA, epilogue warp) sums / sums of squares of the
  int bn_C;                      // stored output (BatchNorm batch statistics fused into the epilogue), else null
  int ngroups;                   // pair kernel: output-channel groups of N handled by ONE launch (wide layers: the blocks
                                 // of all groups share the grid instead of ngroups launches of a few blocks each)
};

constexpr int MAX_NB = 8;


// Sum over the 32 lanes of a warp of 16 values per lane with 16 shuffles: every exchange halves the values a lane
// carries; lane l ends with the warp sum of value (l >> 1).
__device__ __forceinline__ float transpose_reduce16(float (&p)[16], int lane) {
#pragma unroll
  for (int w = 8, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int k = 0; k < w; ++k) {
      const float keep = up ? p[k + w] : p[k], send = up ? p[k] : p[k + w];
      p[k] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  return p[0] + __shfl_xor_sync(0xffffffffu, p[0], 1);
}

// BatchNorm batch statistics of the stored (rounded) outputs: every epilogue lane adds the 16 channels o[8] of its pixel
// into per-lane sums (sv) and sums of squares (sq); once per block and channel chunk the warp reduces them over its 32
// pixels and adds the result to row `row` of the partial table.  Every (row, channel) address is updated by exactly one
// lane of one warp, in program order => deterministic although it is a reduction instruction.
__device__ __forceinline__ void bn_stats_add(const uint32_t (&o)[8], bool valid, bool f16, float (&sv)[16], float (&sq)[16]) {
  if (!valid) return;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float2 t;
    if (f16) t = __half22float2(*reinterpret_cast<const __half2*>(&o[k]));
    else t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&o[k]));
    sv[2 * k] += t.x; sv[2 * k + 1] += t.y;
    sq[2 * k] = fmaf(t.x, t.x, sq[2 * k]); sq[2 * k + 1] = fmaf(t.y, t.y, sq[2 * k + 1]);
  }
}
__device__ __forceinline__ void bn_stats_flush(float (&sv)[16], float (&sq)[16], float* __restrict__ part, int C, int row,
                                               int col0, int lane) {
  const float s = transpose_reduce16(sv, lane), ss = transpose_reduce16(sq, lane);
  if ((lane & 1) == 0) {
    const int c = col0 + (lane >> 1);
    atomicAdd(part + ((size_t)row * 2 + 0) * C + c, s);
    atomicAdd(part + ((size_t)row * 2 + 1) * C + c, ss);
  }
}

// Epilogue of one block (T_TILES M-tiles of 128 pixels, N output channels) for the epilogue warp that owns TMEM lanes
// acc0 >> 16 ..+31: TMEM -> registers -> + bias, + residual, ReLU -> 16-bit -> 16-byte stores (+ the optional bf16 copy).
// Channel chunks are the outer loop so that the BatchNorm statistics of a chunk are reduced across the warp once per
// block, not once per M-tile (the reduction is as long as the rest of the chunk's epilogue).
template <int N>
__device__ __forceinline__ void epilogue_block(const ConvParams& p, uint32_t acc0, const bool (&valid)[T_TILES],
                                               const long long (&pix)[T_TILES], int stat_row, int lane, int nofs = 0) {
  const bool f16 = (p.flags & FLAG_F16) != 0;
#pragma unroll 1
  for (int j = 0; j < N / 16; ++j) {
    float sv[16], sq[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { sv[k] = 0.f; sq[k] = 0.f; }
#pragma unroll
    for (int mt = 0; mt < T_TILES; ++mt) {
      uint32_t v[16];
      tmem_ld16(acc0 + mt * N + j * 16, v);
      tmem_ld_wait();
      if (valid[mt]) {
        float f[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) f[k] = __uint_as_float(v[k]);
        if (p.bias) {
#pragma unroll
          for (int k = 0; k < 16; k += 4) {
            float4 bv = *reinterpret_cast<const float4*>(p.bias + nofs + j * 16 + k);
            f[k] += bv.x; f[k + 1] += bv.y; f[k + 2] += bv.z; f[k + 3] += bv.w;
          }
        }
        if (p.residual) {
          const __nv_bfloat16* rp = p.residual + pix[mt] * p.res_ld + nofs + j * 16;
#pragma unroll
          for (int k = 0; k < 16; k += 4) {
            float4 rv = f16 ? ld4(reinterpret_cast<const __half*>(rp) + k) : ld4(rp + k);
            f[k] += rv.x; f[k + 1] += rv.y; f[k + 2] += rv.z; f[k + 3] += rv.w;
          }
        }
        if (p.flags & FLAG_RELU) {
#pragma unroll
          for (int k = 0; k < 16; ++k) f[k] = fmaxf(f[k], 0.f);
        }
        uint32_t o[8];
        if (f16) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            __half2 h = __floats2half2_rn(f[2 * k], f[2 * k + 1]);
            o[k] = *reinterpret_cast<uint32_t*>(&h);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
            o[k] = *reinterpret_cast<uint32_t*>(&h);
          }
        }
        uint4* op = reinterpret_cast<uint4*>(p.out + pix[mt] * p.out_ld + nofs + j * 16);
        op[0] = make_uint4(o[0], o[1], o[2], o[3]);
        op[1] = make_uint4(o[4], o[5], o[6], o[7]);
        if (p.out2) {
          uint32_t o2[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
            o2[k] = *reinterpret_cast<uint32_t*>(&h);
          }
          uint4* op2 = reinterpret_cast<uint4*>(p.out2 + pix[mt] * p.out2_ld + nofs + j * 16);
          op2[0] = make_uint4(o2[0], o2[1], o2[2], o2[3]);
          op2[1] = make_uint4(o2[4], o2[5], o2[6], o2[7]);
        }
        if (p.bn_partial) bn_stats_add(o, true, f16, sv, sq);
      }
    }
    if (p.bn_partial)      // (warp-uniform: all 32 lanes take part in the shuffles)
      bn_stats_flush(sv, sq, p.bn_partial, p.bn_C, stat_row, nofs + j * 16, lane);
  }
}

// Persistent: gridDim.x CTAs (one per SM) stride over the blocks.  The activation-chunk ring, the weight ring and the
// two TMEM accumulator buffers all run on global counters, so the TMA producers prefetch the next block's halo while
// the current block is still in its main loop and the epilogue of block i overlaps the MMAs of block i+1.
template <int N>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap, const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_slot_bytes = ((uint32_t)p.rows * p.P * 128u + 1023u) & ~1023u;
  const uint32_t a_base = base;
  const int NA_SLOTS = p.na_slots;
  const uint32_t b_base = a_base + NA_SLOTS * a_slot_bytes;
  constexpr uint32_t B_STAGE = N * 128u;
  const int NB = p.nb_stages;
  const uint32_t bar_base = b_base + NB * B_STAGE;
  auto a_full = [&](int i) { return bar_base + 8u * i; };
  auto a_empty = [&](int i) { return bar_base + 8u * (MAX_NA + i); };
  auto b_full = [&](int i) { return bar_base + 8u * (2 * MAX_NA + i); };
  auto b_empty = [&](int i) { return bar_base + 8u * (2 * MAX_NA + MAX_NB + i); };
  auto t_full = [&](int i) { return bar_base + 8u * (2 * MAX_NA + 2 * MAX_NB + i); };
  auto t_empty = [&](int i) { return bar_base + 8u * (2 * MAX_NA + 2 * MAX_NB + 2 + i); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_NA + 2 * MAX_NB + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int taps = p.KS * p.KS;
  constexpr uint32_t ACC_COLS = T_TILES * N;      // one accumulator buffer
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;    // 256 or 512

  if (threadIdx.x == 0) {
    for (int i = 0; i < NA_SLOTS; ++i) { mbar_init(a_full(i), 1); mbar_init(a_empty(i), 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(b_full(i), 1); mbar_init(b_empty(i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(t_full(i), 1); mbar_init(t_empty(i), 4); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===== A producer: one TMA box (64 ch x (8+2pad) px x 1 row) per virtual halo row, issued by all 32 lanes in
    // parallel (a single lane issuing 36 boxes per chunk was the bottleneck of the short 64-channel main loops);
    // the 1x1 conv has no halo and its rows are contiguous across samples, so one 32-row box per chunk suffices =====
    {
      const uint32_t row_bytes = (uint32_t)(8 + 2 * p.pad) * 128u;
      int ac = 0;
      for (int blk = blockIdx.x; blk < p.nblocks; blk += gridDim.x) {
        const int xg = blk % p.nxg, vb = blk / p.nxg;
        const int x0 = xg * 8, v0 = vb * (16 * T_TILES);
        for (int c = 0; c < p.nchunks; ++c, ++ac) {
          const int slot = ac % NA_SLOTS;
          const uint32_t dst0 = a_base + slot * a_slot_bytes;
          if (lane == 0) {
            mbar_wait(a_empty(slot), ((ac / NA_SLOTS) & 1) ^ 1);
            mbar_expect_tx(a_full(slot), row_bytes * p.rows);
          }
          __syncwarp();
          if (p.pad == 0) {
            if (lane == 0) tma_load_4d(dst0, &tmap, c * 64, x0, v0, 0, a_full(slot));
          } else {
            for (int r = lane; r < p.rows; r += 32) {
              const int vr = v0 - p.pad + r;
              int n = 0, y = p.H;   // out-of-bounds row => TMA zero fill
              if (vr >= 0 && vr < p.Vtotal) { n = vr / p.Hp; y = vr - n * p.Hp; }
              tma_load_4d(dst0 + (uint32_t)r * p.P * 128u, &tmap, c * 64, x0 - p.pad, y, n, a_full(slot));
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 2) {
    // ===== B producer: one pre-swizzled [N][64] weight tile per (chunk, tap), same sequence for every block =====
    if (lane == 0) {
      const int per_block = p.nchunks * taps;
      int it = 0;
      for (int blk = blockIdx.x; blk < p.nblocks; blk += gridDim.x) {
        for (int k = 0; k < per_block; ++k, ++it) {
          const int st = it % NB;
          mbar_wait(b_empty(st), ((it / NB) & 1) ^ 1);
          mbar_expect_tx(b_full(st), B_STAGE);
          bulk_load(b_base + st * B_STAGE, p.w + (size_t)k * p.w_tile_elems, B_STAGE, b_full(st));
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the loop converged (keeps the address math in uniform registers);
    // one elected lane issues the tcgen05 instructions =====
    {
      const int bf = (p.flags & FLAG_F16) ? 0 : 1;
      const uint32_t idesc = make_idesc(128, N, 0, 0, bf, bf);
      const uint32_t a_hi = desc_hi((uint32_t)p.P * 128u), b_hi = desc_hi(1024u);
      const uint32_t row_units = (uint32_t)p.P * 8u;            // one halo row, in 16-byte descriptor units
      const uint32_t mt_units = 16u * row_units;                 // next M-tile = 16 halo rows further
      const uint32_t wrap_units = row_units - (uint32_t)p.KS * 8u;
      int it = 0, ac = 0, lb = 0;
      for (int blk = blockIdx.x; blk < p.nblocks; blk += gridDim.x, ++lb) {
        const int buf = lb & 1;
        mbar_wait(t_empty(buf), ((lb >> 1) & 1) ^ 1);     // epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t acc0 = tmem_base + buf * ACC_COLS;
        uint32_t first = 0u;                              // 0 only for the very first MMA of the block
        for (int c = 0; c < p.nchunks; ++c, ++ac) {
          const int slot = ac % NA_SLOTS;
          mbar_wait(a_full(slot), (ac / NA_SLOTS) & 1);
          tc_fence_after();
          uint32_t a_lo = desc_lo(a_base + slot * a_slot_bytes, 16u);   // window start of tap (0,0), M-tile 0
          int kx = 0;
          for (int t = 0; t < taps; ++t, ++it) {
            const int st = it % NB;
            mbar_wait(b_full(st), (it / NB) & 1);
            tc_fence_after();
            const uint32_t b_lo = desc_lo(b_base + st * B_STAGE, 16u);
            if (elect_one()) {
#pragma unroll
              for (int mt = 0; mt < T_TILES; ++mt) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                  umma_bf16(acc0 + mt * N, desc_join(a_lo + mt * mt_units + kk * 2u, a_hi), desc_join(b_lo + kk * 2u, b_hi),
                            idesc, (first | (uint32_t)kk) ? 1u : 0u);
                }
              }
              umma_commit(b_empty(st));
            }
            __syncwarp();
            first = 1u;
            // next tap: one pixel to the right, or wrap to the start of the next halo row
            a_lo += 8u;
            if (++kx == p.KS) { kx = 0; a_lo += wrap_units; }
          }
          if (elect_one()) umma_commit(a_empty(slot));
          __syncwarp();
        }
        if (elect_one()) umma_commit(t_full(buf));
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue: warps 3..6 own TMEM lanes 32*(warp%4).. =====
    const int q = warp & 3;
    const int r = q * 32 + lane;          // accumulator row = pixel within the M-tile
    const int wx = r & 7, vrow = r >> 3;
    int lb = 0;
    for (int blk = blockIdx.x; blk < p.nblocks; blk += gridDim.x, ++lb) {
      const int buf = lb & 1;
      const int xg = blk % p.nxg, vb = blk / p.nxg;
      const int x0 = xg * 8, v0 = vb * (16 * T_TILES);
      mbar_wait(t_full(buf), (lb >> 1) & 1);
      tc_fence_after();
      const uint32_t acc0 = tmem_base + buf * ACC_COLS + ((uint32_t)(q * 32) << 16);
      bool valid[T_TILES];
      long long pix[T_TILES];
#pragma unroll
      for (int mt = 0; mt < T_TILES; ++mt) {
        const int vr = v0 + mt * 16 + vrow;
        const int n = vr / p.Hp, y = vr - n * p.Hp;
        valid[mt] = vr < p.Vtotal && y < p.H;
        pix[mt] = ((long long)n * p.H + y) * p.W + x0 + wx;
      }
      epilogue_block<N>(p, acc0, valid, pix, (int)blockIdx.x * 4 + q, lane);
      // all TMEM reads of this warp are complete (wait::ld above): hand the buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty(buf));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2) of the forward kernel for N = 128: two CTAs of one cluster (adjacent SMs of a TPC)
// each own one block (their own halo tile and their own 128 TMEM lanes per M-tile); the 128x64 weight tile is split
// between them (64 rows each, same shared-memory offset) and ONE tcgen05.mma.cta_group::2 of the leader CTA (M = 256)
// drives both tensor cores.  Per SM and MMA this reads 4 KB of A + 2 KB of B from shared memory (96 B/clk instead of
// the 128 B/clk that saturate the shared-memory port with cta_group::1) and halves the weight traffic from L2.
// Protocol (as in the canonical 2-SM GEMM): both CTAs' TMA loads complete on the LEADER's "full" barriers
// (.cta_group::2, peer bit of the barrier address cleared; leader arrives with expect_tx of both halves, the peer
// arrives remotely without tx); tcgen05.commit multicasts "empty"/"accumulator ready" to both CTAs; both epilogues
// arrive (remotely for the peer) on the leader's "accumulator drained" barrier.
// ------------------------------------------------------------------------------------------------
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
conv_tc_pair_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap wmap,
                    const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_slot_bytes = ((uint32_t)p.rows * p.P * 128u + 1023u) & ~1023u;
  const int NA_SLOTS = p.na_slots;
  const uint32_t a_base = base;
  const uint32_t b_base = a_base + NA_SLOTS * a_slot_bytes;
  constexpr uint32_t B_HALF = (N / 2) * 128u;     // this CTA's N/2 rows of the N x 64 weight tile
  const int NB = p.nb_stages;
  const uint32_t bar_base = b_base + NB * B_HALF;
  auto a_full = [&](int i) { return bar_base + 8u * i; };
  auto a_empty = [&](int i) { return bar_base + 8u * (MAX_NA + i); };
  auto b_full = [&](int i) { return bar_base + 8u * (2 * MAX_NA + i); };
  auto b_empty = [&](int i) { return bar_base + 8u * (2 * MAX_NA + MAX_NB + i); };
  auto t_full = [&](int i) { return bar_base + 8u * (2 * MAX_NA + 2 * MAX_NB + i); };
  auto t_empty = [&](int i) { return bar_base + 8u * (2 * MAX_NA + 2 * MAX_NB + 2 + i); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_NA + 2 * MAX_NB + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int taps = p.KS * p.KS;
  constexpr uint32_t ACC_COLS = T_TILES * N;
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;    // 512 or 256
  const int npairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  const int npb_pix = (p.nblocks + 1) >> 1;       // pair-blocks of one output-channel group
  const int npb = npb_pix * p.ngroups;            // pair-block pb = (group pb / npb_pix, pixel pair-block pb % npb_pix)

  if (threadIdx.x == 0) {
    for (int i = 0; i < NA_SLOTS; ++i) { mbar_init(a_full(i), 2); mbar_init(a_empty(i), 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(b_full(i), 2); mbar_init(b_empty(i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(t_full(i), 1); mbar_init(t_empty(i), 8); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===== A producer (both CTAs): own block's halo rows, completing on the leader's a_full =====
    const uint32_t row_bytes = (uint32_t)(8 + 2 * p.pad) * 128u;
    int ac = 0;
    for (int pb = pair; pb < npb; pb += npairs) {
      const int blk = 2 * (pb % npb_pix) + (int)rank;   // may be == nblocks for the last odd block: all rows out of range
      const int xg = blk % p.nxg, vb = blk / p.nxg;
      const int x0 = xg * 8, v0 = vb * (16 * T_TILES);
      for (int c = 0; c < p.nchunks; ++c, ++ac) {
        const int slot = ac % NA_SLOTS;
        const uint32_t dst0 = a_base + slot * a_slot_bytes;
        const uint32_t full_leader = a_full(slot) & PEER_MASK;
        if (lane == 0) mbar_wait(a_empty(slot), ((ac / NA_SLOTS) & 1) ^ 1);
        __syncwarp();
        if (p.pad == 0) {
          if (lane == 0) tma_load_4d_2sm(dst0, &tmap, c * 64, x0, v0, 0, full_leader);
        } else {
          for (int r = lane; r < p.rows; r += 32) {
            const int vr = v0 - p.pad + r;
            int n = 0, y = p.H;
            if (vr >= 0 && vr < p.Vtotal) { n = vr / p.Hp; y = vr - n * p.Hp; }
            tma_load_4d_2sm(dst0 + (uint32_t)r * p.P * 128u, &tmap, c * 64, x0 - p.pad, y, n, full_leader);
          }
        }
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_expect_tx(a_full(slot), 2u * row_bytes * p.rows);
          else mbar_arrive_cluster(full_leader);
        }
      }
    }
  } else if (warp == 2) {
    // ===== B producer (both CTAs): rows rank*64.. of each [128][64] weight tile =====
    if (lane == 0) {
      const int per_block = p.nchunks * taps;
      const int rows_per_tile = p.w_tile_elems / 64;     // Cout_total
      int it = 0;
      for (int pb = pair; pb < npb; pb += npairs) {
        const int row0 = (pb / npb_pix) * N + (int)rank * (N / 2);
        for (int k = 0; k < per_block; ++k, ++it) {
          const int st = it % NB;
          const uint32_t full_leader = b_full(st) & PEER_MASK;
          mbar_wait(b_empty(st), ((it / NB) & 1) ^ 1);
          tma_load_2d_2sm(b_base + st * B_HALF, &wmap, 0, k * rows_per_tile + row0, full_leader);
          if (leader) mbar_expect_tx(b_full(st), 2u * B_HALF);
          else mbar_arrive_cluster(full_leader);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: leader CTA only =====
    if (leader) {
      const int bf = (p.flags & FLAG_F16) ? 0 : 1;
      const uint32_t idesc = make_idesc(256, N, 0, 0, bf, bf);
      const uint32_t a_hi = desc_hi((uint32_t)p.P * 128u), b_hi = desc_hi(1024u);
      const uint32_t row_units = (uint32_t)p.P * 8u;
      const uint32_t mt_units = 16u * row_units;
      const uint32_t wrap_units = row_units - (uint32_t)p.KS * 8u;
      int it = 0, ac = 0, lb = 0;
      for (int pb = pair; pb < npb; pb += npairs, ++lb) {
        const int buf = lb & 1;
        mbar_wait(t_empty(buf), ((lb >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc0 = tmem_base + buf * ACC_COLS;
        uint32_t first = 0u;
        for (int c = 0; c < p.nchunks; ++c, ++ac) {
          const int slot = ac % NA_SLOTS;
          mbar_wait(a_full(slot), (ac / NA_SLOTS) & 1);
          tc_fence_after();
          uint32_t a_lo = desc_lo(a_base + slot * a_slot_bytes, 16u);
          int kx = 0;
          for (int t = 0; t < taps; ++t, ++it) {
            const int st = it % NB;
            mbar_wait(b_full(st), (it / NB) & 1);
            tc_fence_after();
            const uint32_t b_lo = desc_lo(b_base + st * B_HALF, 16u);
            if (elect_one()) {
#pragma unroll
              for (int mt = 0; mt < T_TILES; ++mt) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                  umma_bf16_2sm(acc0 + mt * N, desc_join(a_lo + mt * mt_units + kk * 2u, a_hi),
                                desc_join(b_lo + kk * 2u, b_hi), idesc, (first | (uint32_t)kk) ? 1u : 0u);
                }
              }
              umma_commit_2sm(b_empty(st));
            }
            __syncwarp();
            first = 1u;
            a_lo += 8u;
            if (++kx == p.KS) { kx = 0; a_lo += wrap_units; }
          }
          if (elect_one()) umma_commit_2sm(a_empty(slot));
          __syncwarp();
        }
        if (elect_one()) umma_commit_2sm(t_full(buf));
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue (both CTAs): own 128 TMEM lanes =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int wx = r & 7, vrow = r >> 3;
    int lb = 0;
    for (int pb = pair; pb < npb; pb += npairs, ++lb) {
      const int buf = lb & 1;
      const int blk = 2 * (pb % npb_pix) + (int)rank;
      const int xg = blk % p.nxg, vb = blk / p.nxg;
      const int x0 = xg * 8, v0 = vb * (16 * T_TILES);
      mbar_wait(t_full(buf), (lb >> 1) & 1);
      tc_fence_after();
      const uint32_t acc0 = tmem_base + buf * ACC_COLS + ((uint32_t)(q * 32) << 16);
      bool valid[T_TILES];
      long long pix[T_TILES];
#pragma unroll
      for (int mt = 0; mt < T_TILES; ++mt) {
        const int vr = v0 + mt * 16 + vrow;
        const int n = vr / p.Hp, y = vr - n * p.Hp;
        valid[mt] = blk < p.nblocks && vr < p.Vtotal && y < p.H;
        pix[mt] = ((long long)n * p.H + y) * p.W + x0 + wx;
      }
      epilogue_block<N>(p, acc0, valid, pix, (int)blockIdx.x * 4 + q, lane, (pb / npb_pix) * N);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(t_empty(buf) & PEER_MASK);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

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
};

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
    for (int i = 0; i < NS; ++i) { mbar_init(full(i), 1); mbar_init(empty(i), 1); }
    mbar_init(tmem_full, 1);
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
  } else if (warp == 1) {
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
    uint32_t st = 0, ph = 0, xs_lo = xs_lo0, b_lo = b_lo0, full_bar = full(0), empty_bar = empty(0);
    for (int i = 0; i < nblk; ++i) {
      mbar_wait(full_bar, ph);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          if (g < ng) {
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
      if (++st == (uint32_t)NS) { st = 0; ph ^= 1u; xs_lo = xs_lo0; b_lo = b_lo0; full_bar = full(0); empty_bar = empty(0); }
    }
    if (elect_one()) umma_commit(tmem_full);
    __syncwarp();
  } else if (warp >= 3) {
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
__global__ void wgrad_tc_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int S, int taps,
                                       int Cin, int Cout, int accumulate, int G) {
  long long n = (long long)taps * Cin * Cout;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n * G) return;
  const int g = (int)(i / n);
  i -= g * n;
  int co = (int)(i % Cout);
  long long r = i / Cout;
  int ci = (int)(r % Cin);
  int tap = (int)(r / Cin);
  const float* pg = partial + (long long)g * S * n;
  float s = 0.f;
  for (int k = 0; k < S; ++k) s += pg[(long long)k * n + i];
  long long o = (((long long)g * Cout + co) * Cin + ci) * taps + tap;
  dw[o] = accumulate ? dw[o] + s : s;
}

// shared-memory plan of the forward kernel: A slots fixed by the geometry, the weight ring takes what is left
void conv_smem_plan(int N, int rows, int P, int na, size_t* smem, int* nb) {
  size_t a = ((size_t)rows * P * 128 + 1023) & ~(size_t)1023;
  size_t fixed = 1024 + na * a + 512;
  int stages = (int)((SMEM_LIMIT - fixed) / ((size_t)N * 128));
  if (stages > MAX_NB) stages = MAX_NB;
  *nb = stages;
  *smem = fixed + (size_t)stages * N * 128;
}

template <int N>
int launch_conv_pair(const CUtensorMap& tmap, ConvParams p, int cout_total, int n0, const void* w_packed,
                     cudaStream_t stream) {
  EncodeTiledFn enc = get_encode();
  // weights as a plain 2D byte image [nchunks*taps*Cout_total rows][128 B]; already in the swizzled layout
  CUtensorMap wmap;
  const int taps = p.KS * p.KS;
  cuuint64_t gdim[2] = {64, (cuuint64_t)p.nchunks * taps * cout_total};
  cuuint64_t gstr[1] = {128};
  cuuint32_t box[2] = {64, N / 2};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&wmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { tsr_set_error("conv2d_tc: weight tensor map failed (%d)", (int)r); return TSR_ERR_CUDA; }
  (void)n0;
  p.na_slots = p.KS == 1 ? 5 : 2;
  size_t a = ((size_t)p.rows * p.P * 128 + 1023) & ~(size_t)1023;
  size_t fixed = 1024 + p.na_slots * a + 512;
  constexpr size_t half = (size_t)(N / 2) * 128;
  int stages = (int)((SMEM_LIMIT - fixed) / half);
  if (stages > MAX_NB) stages = MAX_NB;
  p.nb_stages = stages;
  size_t smem = fixed + (size_t)stages * half;
  TSR_CUDA(cudaFuncSetAttribute(conv_tc_pair_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int npb = (p.nblocks + 1) / 2 * p.ngroups;
  int pairs = npb < num_sms() / 2 ? npb : num_sms() / 2;
  conv_tc_pair_kernel<N><<<2 * pairs, NUM_THREADS, smem, stream>>>(tmap, wmap, p);
  TSR_CHECK_LAUNCH("conv2d_tc_pair");
  return TSR_OK;
}

template <int N>
int launch_conv(const CUtensorMap& tmap, ConvParams p, cudaStream_t stream) {
  size_t smem;
  p.na_slots = p.KS == 1 ? 5 : 2;
  conv_smem_plan(N, p.rows, p.P, p.na_slots, &smem, &p.nb_stages);
  if (p.nb_stages < 2) { tsr_set_error("conv2d_tc: shared memory plan infeasible"); return TSR_ERR_UNSUPPORTED; }
  TSR_CUDA(cudaFuncSetAttribute(conv_tc_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = p.nblocks < num_sms() ? p.nblocks : num_sms();
  conv_tc_kernel<N><<<grid, NUM_THREADS, smem, stream>>>(tmap, p);
  TSR_CHECK_LAUNCH("conv2d_tc");
  return TSR_OK;
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

size_t tsr_conv2d_tc_workspace(int, int, int, int, int, int) { return 0; }

// bf16 NHWC convolution on the tensor cores.  in: [B*H*W][in_ld] (Cin channels from `in`), w_packed from
// tsr_pack_conv_weight_bf16, bias fp32 [Cout] or NULL, residual bf16 [pix][res_ld] or NULL, out bf16.
// out2_bf16 (may be NULL): a second copy of the result rounded to bf16, row stride out2_ld (the "fp16" precision mode keeps
// it as the weight-gradient operand of the next convolution).
// bn_partial (may be NULL): [tsr_conv2d_tc_stat_rows()][2][Cout] floats that receive per-(CTA, warp) partial sums and sums of
// squares of the stored output -- the batch statistics of a BatchNorm that consumes this convolution, finished by
// tsr_bn_finalize_partials without another pass over the tensor.
int tsr_conv2d_tc_stat_rows(void) { return STAT_ROWS; }

int tsr_conv2d_tc(const void* in, int in_ld, const void* w_packed, const float* bias, const void* residual,
                  int res_ld, void* out, int out_ld, int B, int H, int W, int Cin, int Cout, int KS, int flags,
                  void* workspace, size_t ws_bytes, float* bn_partial, void* out2_bf16, int out2_ld, cudaStream_t stream) {
  (void)workspace; (void)ws_bytes;
  TSR_REQUIRE(!out2_bf16 || (out2_ld % 8 == 0 && ((uintptr_t)out2_bf16 & 15) == 0), "conv2d_tc: second output must be 16-byte aligned with a row stride that is a multiple of 8");
  if (bn_partial) TSR_CUDA(cudaMemsetAsync(bn_partial, 0, (size_t)STAT_ROWS * 2 * Cout * sizeof(float), stream));
  TSR_REQUIRE(in && w_packed && out, "conv2d_tc: null pointer");
  TSR_REQUIRE(Cout % 64 == 0 && Cout > 0, "conv2d_tc: Cout must be a multiple of 64 (got %d)", Cout);
  TSR_REQUIRE(Cin % 64 == 0 && Cin > 0, "conv2d_tc: Cin must be a multiple of 64 (got %d)", Cin);
  TSR_REQUIRE(KS == 1 || KS == 3 || KS == 5, "conv2d_tc: kernel size %d unsupported", KS);
  TSR_REQUIRE(W % 8 == 0, "conv2d_tc: W must be a multiple of 8 (got %d)", W);
  TSR_REQUIRE(in_ld % 8 == 0 && out_ld % 8 == 0 && (!residual || res_ld % 8 == 0), "conv2d_tc: row strides must be multiples of 8");
  TSR_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)w_packed & 15) == 0, "conv2d_tc: pointers must be 16-byte aligned");
  EncodeTiledFn enc = get_encode();
  if (!enc) { tsr_set_error("conv2d_tc: cuTensorMapEncodeTiled unavailable"); return TSR_ERR_CUDA; }
  const int pad = KS / 2;
  CUtensorMap tmap;
  cuuint64_t gdim[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)in_ld * 2, (cuuint64_t)W * in_ld * 2, (cuuint64_t)H * W * in_ld * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)(8 + 2 * pad), 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (pad == 0) {   // no halo: rows of consecutive samples are contiguous => one (B*H)-row dimension, 32-row boxes
    gdim[2] = (cuuint64_t)H * B; gdim[3] = 1;
    box[2] = 16 * T_TILES;
  }
  CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { tsr_set_error("conv2d_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return TSR_ERR_CUDA; }
  ConvParams p;
  p.res_ld = res_ld; p.out_ld = out_ld;
  p.B = B; p.H = H; p.W = W; p.Hp = H + pad; p.Vtotal = B * (H + pad);
  // halo pitch: dense (8 + 2*pad pixels, the TMA box width) unless desc-mode bit 1 asks for the 16-pixel pitch
  p.KS = KS; p.pad = pad; p.P = (g_desc_mode & 2) ? (pad ? 16 : 8) : 8 + 2 * pad; p.rows = 16 * T_TILES + 2 * pad;
  p.nchunks = Cin / 64; p.nxg = W / 8; p.flags = flags;
  p.w_tile_elems = Cout * 64;
  const int nvb = tsr_cdiv(p.Vtotal, 16 * T_TILES);
  p.nblocks = nvb * p.nxg;
  // output channels are produced in groups of 128 (or a trailing 64): rows n0.. of every pre-swizzled weight tile
  for (int n0 = 0; n0 < Cout;) {
    int nt = (Cout - n0) >= 128 ? 128 : 64;
    p.w = (const __nv_bfloat16*)w_packed + (size_t)n0 * 64;
    p.bias = bias ? bias + n0 : nullptr;
    p.residual = residual ? (const __nv_bfloat16*)residual + n0 : nullptr;
    p.out = (__nv_bfloat16*)out + n0;
    p.out2 = out2_bf16 ? (__nv_bfloat16*)out2_bf16 + n0 : nullptr;
    p.out2_ld = out2_ld;
    p.bn_partial = bn_partial ? bn_partial + n0 : nullptr;
    p.bn_C = Cout;
    p.ngroups = 1;
    int rc;
    if (nt == 128 && !(g_desc_mode & 16)) { // bit 4 set = force the single-CTA kernel
      p.ngroups = (Cout - n0) / 128;        // all remaining 128-wide groups in this one launch
      rc = launch_conv_pair<128>(tmap, p, Cout, n0, (const __nv_bfloat16*)w_packed + (size_t)n0 * 64, stream);
      nt = 128 * p.ngroups;
    } else if (nt == 64 && !(g_desc_mode & 16) && !(g_desc_mode & 32))
      rc = launch_conv_pair<64>(tmap, p, Cout, n0, (const __nv_bfloat16*)w_packed + (size_t)n0 * 64, stream);
    else
      rc = nt == 128 ? launch_conv<128>(tmap, p, stream) : launch_conv<64>(tmap, p, stream);
    if (rc) return rc;
    n0 += nt;
  }
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
int tsr_conv2d_wgrad_tc(const void* in, int in_ld, const void* dout, int dout_ld, float* dw_oihw, void* workspace,
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
  size_t smem = 1024 + (size_t)ns * q.stage_bytes + 512;
  TSR_CUDA(cudaFuncSetAttribute(wgrad_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(q.ncta_groups * q.nunits, nsplit, G);
  wgrad_tc3_kernel<<<grid, NUM_THREADS, smem, stream>>>(tmx, tmdy, q);
  TSR_CHECK_LAUNCH("conv2d_wgrad_tc3");
  long long n = (long long)taps * Cin * Cout * G;
  wgrad_tc_reduce_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>((const float*)workspace, dw_oihw, nsplit, taps, Cin,
                                                                    Cout, accumulate, G);
  TSR_CHECK_LAUNCH("wgrad_tc_reduce");
  return TSR_OK;
}


}  // extern "C"
