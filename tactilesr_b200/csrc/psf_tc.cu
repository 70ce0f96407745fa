// tPSFNet point-spread-function forward model on the tensor cores (tcgen05 / TMEM).
//
// Same contract as psf_fwd_kernel (psf.cu): replaces the python per-sample loop of reference model/tPSFNet.py:118-125
// (tactilePSF :78-83, depth2tactile :85-100, degradation_process :129-141).  The 99x99 correlation is separable
// (SURVEY.md Appendix B):  conv = alpha * E D E,  E[m][k] = e(|k-m|) for |k-m| <= 49 else 0,  e(t) = exp(-cp2 t^2 / beta^2)
// (100x100, symmetric banded Toeplitz, different for every sample because beta is), i.e. two dense 100^3 contractions per
// sample = 4 MFLOP against 119 KB of compulsory HBM traffic: on FFMA pipes that is 3x over the HBM time, so the two products
// run as tcgen05.mma.
//
// Two arithmetic variants (template PASSES):
//   3  fp32-accurate (the "fp32" precision mode): every operand is split x = hi + lo into two fp16 numbers (power-of-two
//      pre-scaling keeps both halves in fp16's normal range) and each product is three MMAs  hi*hi + lo*hi + hi*lo
//      accumulated in fp32 in TMEM (the dropped lo*lo term is 2^-22 relative).  ~1e-6 rel-L2 against the fp64 reference run.
//   1  one fp16 pass (the 16-bit tensor-core precision modes, tolerance 1e-2): a third of the MMAs, half of the operand
//      tiles and conversions.  ~3e-4 rel-L2.
//
// Per sample (one CTA, 256 threads; two CTAs per SM overlap each other's phases; 6 block barriers per sample):
//   A  the depth plane arrives in shared memory by ONE bulk copy issued a sample ahead (over the dead depth tiles, or
//      -- PASSES 1 -- into its own buffer); plane -> registers; depth max (contact threshold) and |max| (scaling) in one
//      block reduction; tables e(t), Ex_i(t);  depth -> smem as it lies in HBM (row = k: the MN-major B operand) +
//      contact-mask bytes;  E -> smem (K-major SWIZZLE_128B)
//   B  GEMM1  T = E * D      (M=128, N=112, K=7x16)   -> TMEM columns [0,112)
//      (the psf output, a pure function of alpha / beta, is written to HBM while the MMAs run)
//   X  T: TMEM -> registers -> fp16 (hi, lo) -> TMEM columns [128,240) (tcgen05.st): GEMM2 takes its A operand straight
//      from tensor memory, so T never touches shared memory
//   C  GEMM2  HR = T * E     (E symmetric: the same E tiles are the B operand)   -> TMEM columns [0,112)
//   E  epilogue, ONE pass over the accumulator held in registers: second-max fill (tPSFNet.py:95-97), HR rows staged in
//      shared memory (over the dead E tiles) and stored by one bulk copy, LRd = 1e-4 (Ex HR Ex^T - m sum HR) / (1 - m)
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int N = 100;
constexpr int PSF_THREADS = 256;   // (512 = four threads per accumulator row was measured 5..25 % slower: the kernels are
                                   //  bound by shared-memory wavefronts, not by latency hiding)
constexpr float CP2 = 100.0f / 4802.0f;
constexpr float CM2 = 100.0f / 15138.0f;

constexpr uint32_t ROWS = 104;                 // allocated rows of a tile (13 groups of 8); MMAs over-read up to row 127
constexpr uint32_t ATOM = ROWS * 128u;         // one 64-wide atom: rows x 128 B, SWIZZLE_128B
constexpr uint32_t TILE = 2u * ATOM;           // 112 = 64 + 48 elements along the atom direction
constexpr uint32_t PLANE_BYTES = N * N * 4u;   // 40 000
constexpr uint32_t TAB2_LEN = 216;             // per shifted copy (>= 204; 216: a quarter-warp of consecutive rows reads 8 distinct bank groups)

// forward -> backward hand-over, AUX_STRIDE floats per sample: rows m = 0..99 hold U_j(m) = sum_n HR[m][n] Ex_j(n),
// U2_j(m) = sum_n HR[m][n] Ex_j(n) (n - 12 - 25 j)^2 (j = 0..3), the row sum and 3 pad floats; row 100 = {second max}
constexpr int AUX_ROW = 12;
constexpr int AUX_STRIDE = 101 * AUX_ROW;

// Shared-memory map.  Depth tiles first: GEMM1 reads K rows 96..111 of the (MN-major) depth tiles, i.e. 1 KB past a tile's
// 104 rows -- for the last depth tile that lands in E_hi (always finite; it meets the zero K-padding of E).  Over-reads of
// M / N rows >= 104 only produce accumulator rows / columns that are never used.
// NT threads per CTA = 128 x SEGS: SEGS threads share an accumulator row (NT = 256: two threads, 56 columns each)
template <int PASSES, int NT>
struct Lay {
  static constexpr int NW = NT / 32;                                     // warps
  static constexpr int SEGS = NT / 128;                                  // column segments per accumulator row
  static constexpr int COLS = 112 / SEGS;                                // accumulator columns per thread
  static constexpr int NV_LAST = N - (SEGS - 1) * COLS;                  // valid (< 100) columns of the last segment
  static constexpr uint32_t NH = PASSES == 3 ? 2u : 1u;                  // tiles per operand: hi [, lo]
  static constexpr uint32_t OFF_D = 0;                                   // D_hi [, D_lo]
  static constexpr uint32_t OFF_E = NH * TILE;                           // E_hi [, E_lo]
  static constexpr uint32_t TILES_END = 2u * NH * TILE;
  // the raw fp32 depth plane of the NEXT sample (bulk copy): over the depth tiles once GEMM1 has read them (PASSES 3: no
  // shared memory to spare), or in a buffer of its own, loaded a whole sample ahead (PASSES 1)
  static constexpr uint32_t OFF_RAW = PASSES == 3 ? 0u : TILES_END;
  static constexpr uint32_t OFF_TAB = PASSES == 3 ? TILES_END : TILES_END + PLANE_BYTES;   // float e(t), t = 0..99 (+ pad)
  static constexpr uint32_t OFF_TAB2 = OFF_TAB + 128 * 4;                // 4 shifted copies of uint32 (hi | lo << 16) of 16 e(|j - 99|)
  static constexpr uint32_t OFF_EX = OFF_TAB2 + 4 * TAB2_LEN * 4;        // float4 (Ex_0..Ex_3)(t), t = 0..99
  static constexpr uint32_t OFF_MASK = OFF_EX + 100 * 16;                // contact bytes [13][104]: bit j of [cg][k] <-> depth[k][8 cg + j]
  static constexpr uint32_t OFF_RED = OFF_MASK + 104 * 16;               // float scratch: 2 x [NW] block-max partials, [NW][20] sums
  static constexpr uint32_t OFF_BAR = OFF_RED + (2 * NW + NW * 20) * 4;  // 3 mbarriers + tmem slot
  static constexpr uint32_t OFF_X24 = PASSES == 3 ? 40064u : OFF_BAR + 32u;   // float4 Ex_j(t) (t - 12 - 25 j)^2 (training hand-over)
  // row-segment exchange of the hand-over statistics: per writer segment [100] float4 U, [100] float4 U2, [100] float sum
  static constexpr uint32_t OFF_XCH = OFF_X24 + 1600u;
  static constexpr uint32_t XCH_SEG = 100 * 36;
  static constexpr uint32_t XCH_END = OFF_XCH + (SEGS - 1) * XCH_SEG;
  static constexpr uint32_t USED = PASSES == 3 ? OFF_BAR + 32u : XCH_END;
  // the finished HR plane (dense rows) before its bulk store: over tiles that are dead after GEMM2
  static constexpr uint32_t OFF_STAGE = PASSES == 3 ? OFF_E : 0u;
  static constexpr size_t BYTES = (USED + 15u) & ~15u;
  static_assert(OFF_BAR % 8 == 0 && OFF_TAB % 16 == 0 && OFF_RAW % 16 == 0, "alignment");
  static_assert(USED - TILES_END >= 3072, "the last tile's over-read must stay inside the allocation");
  static_assert(PASSES != 3 || XCH_END <= OFF_E, "hand-over scratch inside the dead depth tiles");
  static_assert(OFF_STAGE + PLANE_BYTES <= TILES_END, "HR staging inside the tiles");
  static_assert(2 * (BYTES + 1024) <= 228 * 1024, "two CTAs per SM");
};

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// A operand in tensor memory (lane = row m, one 32-bit column = two consecutive K elements), B in shared memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
      "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
        "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
        "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// byte offset of the 16-byte chunk cg (elements 8 cg .. 8 cg + 7 along the contiguous direction) of row r of a tile
__device__ __forceinline__ uint32_t chunk_off(int r, int cg) {
  return (uint32_t)(cg >> 3) * ATOM + (uint32_t)r * 128u + (uint32_t)(((cg & 7) ^ (r & 7)) << 4);
}

__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// (x0, x1) -> packed fp16 hi pair and lo pair with hi + lo = x to ~22 bits
__device__ __forceinline__ void split_h2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 f = __half22float2(h);
  hi = h2_bits(h);
  lo = h2_bits(__floats2half2_rn(x0 - f.x, x1 - f.y));
}

// block maximum of two values with ONE barrier: partials in red[0..NW) / red[NW..2 NW); the caller has another barrier
// before the next call
template <int NT>
__device__ __forceinline__ void block_max2(float& a, float& b, float* red) {
  constexpr int NW = NT / 32;
  a = warp_max(a);
  b = warp_max(b);
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = a; red[NW + (threadIdx.x >> 5)] = b; }
  __syncthreads();
  const float4* r4 = reinterpret_cast<const float4*>(red);
  a = -INFINITY; b = -INFINITY;
#pragma unroll
  for (int i = 0; i < NW / 4; ++i) {
    const float4 x = r4[i], y = r4[NW / 4 + i];
    a = fmaxf(a, fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w)));
    b = fmaxf(b, fmaxf(fmaxf(y.x, y.y), fmaxf(y.z, y.w)));
  }
}

// D[128 x 112] (TMEM) = A B over PASSES operand pairs (hi hi [+ lo hi + hi lo]).  A: K-major smem tile.  B: MN-major smem
// tile (rows = K; the two 64-wide N atoms are ATOM bytes apart).
template <int PASSES, int ACCUMULATE = 0>
__device__ __forceinline__ void issue_gemm_kmn(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                               uint32_t idesc) {
  const uint32_t hi_word = desc_hi(1024u);
  uint32_t acc = ACCUMULATE;
#pragma unroll
  for (int pass = 0; pass < PASSES; ++pass) {
    const uint32_t a0 = pass == 1 ? a_lo : a_hi;
    const uint32_t b0 = pass == 2 ? b_lo : b_hi;
#pragma unroll
    for (int ks = 0; ks < 7; ++ks) {
      const uint32_t koff = (uint32_t)(ks >> 2) * ATOM + (uint32_t)(ks & 3) * 32u;     // 16 elements along K, K-major
      umma_f16(tmem_d, desc_join(desc_lo(a0 + koff, 16u), hi_word),
               desc_join(desc_lo(b0 + (uint32_t)ks * 2048u, ATOM), hi_word), idesc, acc);     // B: 16 K rows further
      acc = 1u;
    }
  }
}
// both operands K-major smem tiles
template <int PASSES, int ACCUMULATE = 0>
__device__ __forceinline__ void issue_gemm_kk(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                              uint32_t idesc) {
  const uint32_t hi_word = desc_hi(1024u);
  uint32_t acc = ACCUMULATE;
#pragma unroll
  for (int pass = 0; pass < PASSES; ++pass) {
    const uint32_t a0 = pass == 1 ? a_lo : a_hi;
    const uint32_t b0 = pass == 2 ? b_lo : b_hi;
#pragma unroll
    for (int ks = 0; ks < 7; ++ks) {
      const uint32_t koff = (uint32_t)(ks >> 2) * ATOM + (uint32_t)(ks & 3) * 32u;
      umma_f16(tmem_d, desc_join(desc_lo(a0 + koff, 16u), hi_word), desc_join(desc_lo(b0 + koff, 16u), hi_word), idesc, acc);
      acc = 1u;
    }
  }
}
// A in tensor memory (packed fp16 pairs: 8 columns per K step), B a K-major smem tile
template <int PASSES>
__device__ __forceinline__ void issue_gemm_tk(uint32_t tmem_d, uint32_t t_hi, uint32_t t_lo, uint32_t b_hi, uint32_t b_lo,
                                              uint32_t idesc) {
  const uint32_t hi_word = desc_hi(1024u);
  uint32_t acc = 0u;
#pragma unroll
  for (int pass = 0; pass < PASSES; ++pass) {
    const uint32_t a0 = pass == 1 ? t_lo : t_hi;
    const uint32_t b0 = pass == 2 ? b_lo : b_hi;
#pragma unroll
    for (int ks = 0; ks < 7; ++ks) {
      const uint32_t koff = (uint32_t)(ks >> 2) * ATOM + (uint32_t)(ks & 3) * 32u;
      umma_f16_ts(tmem_d, a0 + (uint32_t)ks * 8u, desc_join(desc_lo(b0 + koff, 16u), hi_word), idesc, acc);
      acc = 1u;
    }
  }
}

// (row, 8-element chunk) work items of a 100 x 100 plane / of a Toeplitz tile incl. its zero K-padding chunk 13.
// RMAJOR: consecutive lanes take consecutive ROWS of one chunk column (rows padded to 104, so 8-row groups stay inside a
// quarter-warp): every 16-byte shared-memory access of a quarter-warp -- row-major fp32 plane (row pitch 400 B), swizzled
// tile, shifted table -- then falls into 8 distinct bank groups.  Chunk-fastest order (consecutive lanes = consecutive
// 32-byte pieces of a row) is kept for planes read straight from global memory, where it coalesces.
constexpr int ITEMS = 13 * 104;
constexpr int ITEMS_E = 14 * 104;
template <int NT> __host__ __device__ constexpr int ipt() { return (ITEMS + NT - 1) / NT; }        // 6 rounds per thread
template <int NT> __host__ __device__ constexpr int ipt_e() { return (ITEMS_E + NT - 1) / NT; }
template <bool RMAJOR>
__device__ __forceinline__ bool item_rc(int item, int nchunks, int& r, int& cg) {
  if (RMAJOR) {
    cg = item / 104;
    r = item - cg * 104;
    return cg < nchunks && r < N;
  }
  r = item / nchunks;
  cg = item - r * nchunks;
  return r < N;
}

// this thread's chunks of one depth plane -> registers: item = (row k, columns 8 cg .. 8 cg + 7), consecutive lanes read
// consecutive 32-byte pieces; chunk 12 of a row holds columns 96..99 only (the rest reads as 0).  SMEM: the plane lies in
// shared memory (bulk copy), else in global memory.
template <bool SMEM, int NT>
__device__ __forceinline__ void load_plane(const float* __restrict__ dsrc, int tid, float (&dreg)[ipt<NT>()][8]) {
#pragma unroll
  for (int i = 0; i < ipt<NT>(); ++i) {
    int k, cg;
    const bool ok = item_rc<SMEM>(tid + i * NT, 13, k, cg);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
    if (ok) {
      const float4* src = reinterpret_cast<const float4*>(dsrc + k * N + cg * 8);
      if (SMEM) {
        a = src[0];
        if (cg < 12) c = src[1];
      } else {
        a = __ldg(src);
        if (cg < 12) c = __ldg(src + 1);
      }
    }
    dreg[i][0] = a.x; dreg[i][1] = a.y; dreg[i][2] = a.z; dreg[i][3] = a.w;
    dreg[i][4] = c.x; dreg[i][5] = c.y; dreg[i][6] = c.z; dreg[i][7] = c.w;
  }
}

// 4 shifted copies of the packed (hi | lo << 16) fp16 table of the banded Toeplitz generator, so that any 8 consecutive
// entries are two aligned 16-byte loads.  kind 0: 16 e(t);  1: 16 e(t) t^2;  2 / 3: 32768 e(t)   (t = |j - 99| <= 49, else 0:
// the PSF has 99 taps); kind 3 carries 16 e(t) t^2 as a single fp16 in the lo half instead of the residual
template <int KIND, int NT>
__device__ __forceinline__ void build_tab2(const float* tab, uint32_t* tab2, int tid) {
  static_assert(NT >= (int)TAB2_LEN, "one thread per table entry, the four shifted copies unrolled (no index division)");
  if (tid >= (int)TAB2_LEN) return;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int qn = tid, i = s * (int)TAB2_LEN + qn;
    const int j = qn + s;
    const int t = j < 99 ? 99 - j : j - 99;
    float x = 0.f;
    if (t <= 49) x = KIND == 0 ? tab[t] * 16.0f : (KIND == 1 ? tab[t] * (float)(16 * t * t) : tab[t] * 32768.0f);
    uint32_t hi, lo;
    if (KIND == 3) {          // two single-fp16 generators in one entry: 32768 e(t) | 16 e(t) t^2 << 16
      const float y = t <= 49 ? tab[t] * (float)(16 * t * t) : 0.f;
      tab2[i] = h2_bits(__floats2half2_rn(x, y));
      continue;
    }
    split_h2(x, 0.f, hi, lo);
    tab2[i] = (hi & 0xFFFFu) | (lo << 16);
  }
}

// K-major hi [/ lo] tiles of the Toeplitz matrix from the table: M[r][8 cg + j] = gen(|8 cg + j - r|); chunk 13 (the K
// padding 104..111, which the MMAs read) is rewritten as zeros: HR staging / raw planes lie over the tiles between samples
template <int PASSES, int NT>
__device__ __forceinline__ void build_toeplitz_tiles(const uint32_t* tab2, uint8_t* t_hi, uint8_t* t_lo, int tid) {
#pragma unroll
  for (int i = 0; i < ipt_e<NT>(); ++i) {
    int r, cg;
    if (item_rc<true>(tid + i * NT, 14, r, cg)) {
      const uint32_t off = chunk_off(r, cg);
      uint4 p0 = make_uint4(0u, 0u, 0u, 0u), p1 = p0;
      if (cg < 13) {
        const int start = 8 * cg - r + 99;
        const int s = start & 3;
        const uint4* tp = reinterpret_cast<const uint4*>(tab2 + s * (int)TAB2_LEN + (start - s));
        p0 = tp[0];
        if (cg < 12) p1 = tp[1];                             // chunk 12: k = 100..103 are K padding
      }
      *reinterpret_cast<uint4*>(t_hi + off) = make_uint4(__byte_perm(p0.x, p0.y, 0x5410), __byte_perm(p0.z, p0.w, 0x5410),
                                                         __byte_perm(p1.x, p1.y, 0x5410), __byte_perm(p1.z, p1.w, 0x5410));
      if (PASSES == 3)
        *reinterpret_cast<uint4*>(t_lo + off) = make_uint4(__byte_perm(p0.x, p0.y, 0x7632), __byte_perm(p0.z, p0.w, 0x7632),
                                                           __byte_perm(p1.x, p1.y, 0x7632), __byte_perm(p1.z, p1.w, 0x7632));
    }
  }
}

// the register-held depth plane -> hi [/ lo] tiles as it lies in HBM (row = k: the MN-major B operand) + contact bytes.
// Everything GEMM1 reads as K rows 100..111 meets the zero K-padding of E and must be finite, and other data has been lying
// over the tiles: rows 100..103 of every atom are zeroed, and -- rows 104..111 of an atom-0 are the first 8 rows of the
// following atom-1 -- the column chunks 5..7 of those rows, which no depth store covers (columns 104..127)
template <int PASSES, int NT, bool RMAJOR>
__device__ __forceinline__ void store_plane_tiles(const float (&dreg)[ipt<NT>()][8], float sD, float thr, uint8_t* x_hi,
                                                  uint8_t* x_lo, uint8_t* maskb, int tid) {
#pragma unroll
  for (int i = 0; i < ipt<NT>(); ++i) {
    int r, cg;
    if (item_rc<RMAJOR>(tid + i * NT, 13, r, cg)) {
      const uint32_t off = chunk_off(r, cg);
      uint32_t dh[4], dl[4], bits = 0u;
#pragma unroll
      for (int j = 0; j < 4; ++j) split_h2(dreg[i][2 * j] * sD, dreg[i][2 * j + 1] * sD, dh[j], dl[j]);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (dreg[i][j] > thr) bits |= 1u << j;
      if (cg == 12) bits &= 0x0Fu;                       // chunk 12 holds columns 96..99 only
      *reinterpret_cast<uint4*>(x_hi + off) = make_uint4(dh[0], dh[1], dh[2], dh[3]);
      if (PASSES == 3) *reinterpret_cast<uint4*>(x_lo + off) = make_uint4(dl[0], dl[1], dl[2], dl[3]);
      maskb[cg * 104 + r] = (uint8_t)bits;
    }
  }
  constexpr int NATOM = PASSES == 3 ? 4 : 2;
  if (tid < NATOM * 32) {
    const int a = tid >> 5, r = 100 + ((tid >> 3) & 3), c = tid & 7;      // atom x row x 16-byte chunk
    *reinterpret_cast<uint4*>(x_hi + a * ATOM + r * 128 + c * 16) = make_uint4(0u, 0u, 0u, 0u);
  } else if (tid < NATOM * 32 + NATOM * 12) {
    const int t = tid - NATOM * 32, tile = t / 24, r = (t % 24) / 3, c = 5 + t % 3;
    *reinterpret_cast<uint4*>(x_hi + tile * TILE + ATOM + r * 128 + ((c ^ (r & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
  }
}

__device__ __forceinline__ void tmem_st2(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(r[0]), "r"(r[1]) : "memory");
}
// COLS (56 or 28) consecutive accumulator columns of this thread's lane -> registers (loads in flight: call tmem_ld_wait)
template <int COLS>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t* v) {
  if (COLS == 56) {
    tmem_ld32(taddr, v);
    tmem_ld16(taddr + 32u, v + 32);
    tmem_ld8(taddr + 48u, v + 48);
  } else {
    tmem_ld16(taddr, v);
    tmem_ld8(taddr + 16u, v + 16);
    tmem_ld4(taddr + 24u, v + 24);
  }
}
// NP (28 or 14) consecutive columns of packed fp16 pairs
template <int NP>
__device__ __forceinline__ void tmem_st_pairs(uint32_t taddr, const uint32_t* r) {
  if (NP == 28) {
    tmem_st16(taddr, r);
    tmem_st8(taddr + 16u, r + 16);
    tmem_st4(taddr + 24u, r + 24);
  } else {
    tmem_st8(taddr, r);
    tmem_st4(taddr + 8u, r + 8);
    tmem_st2(taddr + 12u, r + 12);
  }
}

// accumulator (TMEM, this thread's 56 columns from column seg * 56 of `acc`) * scale -> K-major hi [/ lo] smem tiles (whole
// 16-byte chunks: 4 + 3 chunks in two register batches); thread = (row m, column segment)
template <int PASSES, int COLS>
__device__ __forceinline__ void acc_to_tiles(uint32_t acc, float scale, uint8_t* x_hi, uint8_t* x_lo, int m, int seg) {
  static_assert(COLS == 56, "seven chunks per thread");
#pragma unroll
  for (int part = 0; part < 2; ++part) {
    uint32_t v[32];
    const int cbase = seg * COLS + part * 32;
    if (part == 0) {
      tmem_ld32(acc + (uint32_t)cbase, v);
    } else {
      tmem_ld16(acc + (uint32_t)cbase, v);
      tmem_ld8(acc + (uint32_t)cbase + 16u, v + 16);
    }
    tmem_ld_wait();
    if (m < (int)ROWS) {
#pragma unroll
      for (int g = 0; g < (part == 0 ? 4 : 3); ++g) {
        const int col0 = cbase + g * 8;
        uint32_t th[4], tl[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // columns >= 100 are padding (depth-tile garbage): force 0
          const float x0 = col0 + 2 * j < N ? __uint_as_float(v[g * 8 + 2 * j]) * scale : 0.f;
          const float x1 = col0 + 2 * j + 1 < N ? __uint_as_float(v[g * 8 + 2 * j + 1]) * scale : 0.f;
          split_h2(x0, x1, th[j], tl[j]);
        }
        const uint32_t off = chunk_off(m, col0 >> 3);
        *reinterpret_cast<uint4*>(x_hi + off) = make_uint4(th[0], th[1], th[2], th[3]);
        if (PASSES == 3) *reinterpret_cast<uint4*>(x_lo + off) = make_uint4(tl[0], tl[1], tl[2], tl[3]);
      }
    }
  }
}

// accumulator (TMEM, this thread's COLS columns) -> packed fp16 hi [/ lo] pairs in tensor memory at t_hi / t_lo (this
// thread's COLS / 2 columns of each): the A operand of the next product.  K columns >= 100 (depth-tile garbage) are zeroed.
template <int PASSES, int COLS>
__device__ __forceinline__ void acc_to_tmem(uint32_t acc, uint32_t t_hi, uint32_t t_lo, int seg) {
  constexpr int NP = COLS / 2;
  uint32_t v[COLS];
  tmem_ld_cols<COLS>(acc + (uint32_t)(seg * COLS), v);
  tmem_ld_wait();
  uint32_t th[NP], tl[NP];
  constexpr int SEGS_ = 112 / COLS, JV = (N - (SEGS_ - 1) * COLS) / 2;     // pairs of the last segment below k = 100
  const bool lastseg = seg == SEGS_ - 1;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    split_h2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]), th[j], tl[j]);
    if (j >= JV && lastseg) { th[j] = 0u; tl[j] = 0u; }                 // k >= 100
  }
  const uint32_t p0 = (uint32_t)(seg * NP);
  tmem_st_pairs<NP>(t_hi + p0, th);
  if (PASSES == 3) tmem_st_pairs<NP>(t_lo + p0, tl);
  tmem_st_wait();
}

// depth max / abs-max over this thread's register-held chunks
template <int NT, bool RMAJOR>
__device__ __forceinline__ void plane_max(const float (&dreg)[ipt<NT>()][8], int tid, float& lmax, float& lamax) {
  lmax = -INFINITY; lamax = 0.f;
#pragma unroll
  for (int i = 0; i < ipt<NT>(); ++i) {
    int r, cg;
    if (item_rc<RMAJOR>(tid + i * NT, 13, r, cg)) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < 4 || cg < 12) lmax = fmaxf(lmax, dreg[i][j]);      // chunk 12 holds columns 96..99 only
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) lamax = fmaxf(lamax, fabsf(dreg[i][j]));
  }
}

// e(t) and the (Ex_0..Ex_3)(t) table of one sample
template <int NT>
__device__ __forceinline__ void build_tables(float beta, float gamma, float* tab, float4* ex4, int tid) {
  const float inv_b2 = 1.0f / (beta * beta);
  if (tid < N) tab[tid] = expf(-(CP2 * (float)(tid * tid)) * inv_b2);
  const float inv_g = 1.0f / gamma;
  for (int i = tid; i < 4 * N; i += NT) {
    const int k = i / N, t = i - k * N;
    const float d = (float)(t - 12 - 25 * k);
    reinterpret_cast<float*>(ex4)[t * 4 + k] = expf(-(CM2 * d * d) * inv_g);
  }
}

// this thread's 56 contact bits: bit j <-> column seg * 56 + j of row `row`.  Contact bytes [13][104]: bit j of [cg][k] <->
// depth[k][8 cg + j] (chunk-major: consecutive rows = consecutive bytes for the row-per-lane writers and readers)
template <int COLS>
__device__ __forceinline__ uint64_t contact_bits(const uint8_t* maskb, int row, int seg) {
  static_assert(COLS == 56, "seven mask bytes per thread");
  uint64_t b = 0ull;
#pragma unroll
  for (int c = 0; c < 7; ++c) {
    const int cg = seg * 7 + c;
    if (cg < 13) b |= (uint64_t)maskb[cg * 104 + row] << (8 * c);
  }
  return b;
}
__device__ __forceinline__ bool bitof(uint64_t bits, int j) { return ((bits >> j) & 1ull) != 0ull; }

// AUX: also leave the backward hand-over in `aux` (a compile-time switch: as a run-time one its 5 instructions per pixel
// were issued predicated-off in every inference launch)
template <int PASSES, bool AUX, int NT>
__global__ void __launch_bounds__(NT, 2)
psf_fwd_tc_kernel(const float* __restrict__ ab, const float* __restrict__ depth, float* __restrict__ HR,
                  float* __restrict__ LRd, float* __restrict__ psf, float* __restrict__ aux, int B) {
  using L = Lay<PASSES, NT>;
  constexpr int NW = L::NW, SEGS = L::SEGS, COLS = L::COLS, NV_LAST = L::NV_LAST, IPT = ipt<NT>();
  extern __shared__ __align__(1024) uint8_t sm[];
  const uint32_t base = smem_u32(sm);
  float* tab = reinterpret_cast<float*>(sm + L::OFF_TAB);
  uint32_t* tab2 = reinterpret_cast<uint32_t*>(sm + L::OFF_TAB2);
  float4* ex4 = reinterpret_cast<float4*>(sm + L::OFF_EX);
  uint8_t* maskb = sm + L::OFF_MASK;
  float* red = reinterpret_cast<float*>(sm + L::OFF_RED);           // [0, 2 NW) block-max partials, then [NW][20] sums
  float* reds = red + 2 * NW;
  const uint32_t bar1 = base + L::OFF_BAR, bar2 = bar1 + 8u, bar_raw = bar1 + 16u, tmem_slot = bar1 + 24u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + L::OFF_BAR + 24);
  uint8_t* const d_hi = sm + L::OFF_D; uint8_t* const d_lo = sm + L::OFF_D + TILE;
  uint8_t* const e_hi = sm + L::OFF_E; uint8_t* const e_lo = sm + L::OFF_E + TILE;
  const uint32_t a_d_hi = base + L::OFF_D, a_d_lo = a_d_hi + TILE, a_e_hi = base + L::OFF_E, a_e_lo = a_e_hi + TILE;
  const float* raw = reinterpret_cast<const float*>(sm + L::OFF_RAW);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if ((base & 1023u) != 0u) {                  // SWIZZLE_128B atoms are addressed by absolute shared-memory address bits
    if (tid == 0) printf("tactilesr_b200 psf_tc: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }

  // one-time: zero the tiles (padding chunks) and the mask table, barriers, TMEM
  for (uint32_t i = tid; i < L::TILES_END / 16; i += NT) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (uint32_t i = tid; i < 104 * 16 / 4; i += NT) reinterpret_cast<uint32_t*>(maskb)[i] = 0u;
  if (tid == 0) {
    mbar_init(bar1, 1);
    mbar_init(bar2, 1);
    mbar_init(bar_raw, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_proxy_async();                         // the zeroing above -> ordered before the first bulk copy into the tiles
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t idesc1 = make_idesc(128, 112, 0, 1, 0, 0);   // fp16 x fp16 -> fp32; A K-major, B MN-major
  const uint32_t idesc2 = make_idesc(128, 112, 0, 0, 0, 0);   // A from tensor memory, B K-major

  // epilogue geometry: TMEM lane quarter q = warp % 4, row m = 32 q + lane; column segment seg = warp / 4 (COLS columns)
  const int q = warp & 3, seg = warp >> 2;
  const int m = q * 32 + lane;
  const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
  const uint32_t acc = tmem_base + lane_addr;                  // accumulator of both products: columns [0,112)
  const uint32_t t_hi = tmem_base + 128u, t_lo = tmem_base + 184u;     // packed T: 56 columns each

  auto load_raw = [&](int bn) {                // one elected thread: bulk copy of sample bn's depth plane
    mbar_expect_tx(bar_raw, PLANE_BYTES);
    bulk_load(base + L::OFF_RAW, depth + (size_t)bn * N * N, PLANE_BYTES, bar_raw);
  };
  if (tid == 0 && (int)blockIdx.x < B) load_raw(blockIdx.x);

  int it = 0;
  for (int b = blockIdx.x; b < B; b += gridDim.x, ++it) {
    const uint32_t ph = (uint32_t)(it & 1);
    const bool has_next = b + (int)gridDim.x < B;
    const float alpha = ab[b * 3 + 0], beta = ab[b * 3 + 1], gamma = ab[b * 3 + 2];

    // ---- phase A1: plane -> registers, depth max / |max|, tables ----
    float dreg[IPT][8];
    mbar_wait(bar_raw, ph);
    load_plane<true, NT>(raw, tid, dreg);
    float dmax, amax;
    plane_max<NT, true>(dreg, tid, dmax, amax);
    build_tables<NT>(beta, gamma, tab, ex4, tid);
    // the previous sample's HR plane has left shared memory before the tiles under it are rebuilt (the barrier inside
    // block_max2 orders this wait against every other thread)
    if (tid == 32) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    block_max2<NT>(dmax, amax, red);                   // (syncs: tab / ex4 visible, every thread holds its part of the plane)
    if (PASSES == 1 && tid == 0 && has_next) load_raw(b + gridDim.x);      // own buffer: a whole sample ahead
    const float thr = dmax - 1e-3f;
    int dexp = 0;
    if (amax > 0.f) (void)frexpf(amax, &dexp);         // amax = f * 2^dexp, f in [0.5, 1)
    const float sD = ldexpf(1.0f, 4 - dexp);           // |depth| sD < 16
    build_tab2<0, NT>(tab, tab2, tid);
    // ---- phase A2: depth tiles (MN-major: row = k, as in HBM) + contact bytes; then the E tiles (K-major) ----
    store_plane_tiles<PASSES, NT, true>(dreg, sD, thr, d_hi, d_lo, maskb, tid);
    __syncthreads();
    build_toeplitz_tiles<PASSES, NT>(tab2, e_hi, e_lo, tid);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();

    // ---- phase B: GEMM1  T = E * D ----
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        issue_gemm_kmn<PASSES>(tmem_base, a_e_hi, a_e_lo, a_d_hi, a_d_lo, idesc1);
        umma_commit(bar1);
      }
      __syncwarp();
    }
    // psf = alpha e(u) e(v)  (tPSFNet.py:83): warp w owns rows u = u0 + w, u0 + w + NW, ...; rows 0..49 are written while
    // GEMM1 runs, rows 50..98 while GEMM2 runs
    auto write_psf = [&](int u0, int u1) {
      if (!psf) return;
      float* pdst = psf + (size_t)b * 99 * 99;
      float ev[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int v = lane + 32 * c;
        ev[c] = v < 99 ? tab[v < 49 ? 49 - v : v - 49] : 0.f;
      }
      for (int u = u0 + warp; u < u1; u += NW) {
        const float aeu = alpha * tab[u < 49 ? 49 - u : u - 49];
        float* row = pdst + u * 99;
#pragma unroll
        for (int c = 0; c < 3; ++c) row[lane + 32 * c] = aeu * ev[c];
        if (lane < 3) row[lane + 96] = aeu * ev[3];
      }
    };
    write_psf(0, 50);
    mbar_wait(bar1, ph);
    tc_fence_after();
    if (PASSES == 3 && tid == 0 && has_next) load_raw(b + gridDim.x);      // over the depth tiles GEMM1 has finished reading

    // ---- phase X: T -> fp16 hi [/ lo] -> tensor memory (the A operand of GEMM2) ----
    // accumulator = 16 sD T = 2^(8 - dexp) T,  |T| <= 99 |depth|max  =>  < 2^15: no rescaling needed
    acc_to_tmem<PASSES, COLS>(acc, t_hi + lane_addr, t_lo + lane_addr, seg);
    tc_fence_before();
    __syncthreads();

    // ---- phase C: GEMM2  HR = T * E  (E symmetric: its K-major tile is also the B operand) ----
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        issue_gemm_tk<PASSES>(tmem_base, t_hi, t_lo, a_e_hi, a_e_lo, idesc2);
        umma_commit(bar2);
      }
      __syncwarp();
    }
    write_psf(50, 99);
    // (Ex_j(t) (t - 12 - 25 j)^2) and the row-segment hand-over of the backward statistics (AUX only) live in memory that
    // is dead by now
    float4* ex24 = reinterpret_cast<float4*>(sm + L::OFF_X24);
    if (AUX) {
      for (int i = tid; i < 4 * N; i += NT) {
        const int t = i >> 2, k = i & 3;
        const float d = (float)(t - 12 - 25 * k);
        reinterpret_cast<float*>(ex24)[i] = reinterpret_cast<const float*>(ex4)[i] * (d * d);
      }
    }
    const uint64_t bits = contact_bits<COLS>(maskb, m < N ? m : 0, seg);
    mbar_wait(bar2, ph);
    tc_fence_after();

    // ---- phase E: epilogue.  accumulator = 2^(8 - dexp) 16 (E D E); this thread: row m, columns c0 .. c0 + COLS - 1 ----
    const float cs = alpha * ldexpf(1.0f, dexp - 12);
    const int c0 = seg * COLS;
    const bool last = seg == SEGS - 1;                  // the last segment has NV_LAST valid columns, the rest is padding
    float h[COLS];
    {
      uint32_t v[COLS];
      tmem_ld_cols<COLS>(acc + (uint32_t)c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < COLS / 2; ++j) {
        const float2 t = fmul2s(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), cs);
        h[2 * j] = t.x; h[2 * j + 1] = t.y;
      }
    }
    tc_fence_before();                                  // (all TMEM reads of this sample are complete)
    // second max = max over the conv result with the contact pixels zeroed (tPSFNet.py:95-97)
    const bool any_contact = bits != 0ull;
    float m2 = 0.f, unused = 0.f;
    if (m < N) {
      if (!any_contact) {
#pragma unroll
        for (int j = 0; j < NV_LAST; ++j) m2 = fmaxf(m2, h[j]);
        if (!last) {
#pragma unroll
          for (int j = NV_LAST; j < COLS; ++j) m2 = fmaxf(m2, h[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < NV_LAST; ++j)
          if (!bitof(bits, j)) m2 = fmaxf(m2, h[j]);
        if (!last) {
#pragma unroll
          for (int j = NV_LAST; j < COLS; ++j)
            if (!bitof(bits, j)) m2 = fmaxf(m2, h[j]);
        }
      }
    }
    block_max2<NT>(m2, unused, red);                    // (the tile barriers since the depth-max call protect the scratch)
    // fill, stage HR, accumulate the degradation sums of this thread's row segment
    float rj[4] = {0.f, 0.f, 0.f, 0.f}, r2[4] = {0.f, 0.f, 0.f, 0.f}, rs = 0.f;
    float* hstage = reinterpret_cast<float*>(sm + L::OFF_STAGE);
    if (m < N) {
      if (any_contact) {
#pragma unroll
        for (int j = 0; j < COLS; ++j)
          if (bitof(bits, j)) h[j] = m2;
      }
      float2 rj01 = make_float2(0.f, 0.f), rj23 = rj01, r201 = rj01, r223 = rj01;   // (FFMA2: two sums per instruction)
      auto accumulate = [&](int j) {
        const float4 e4 = ex4[c0 + j];
        rs += h[j];
        ffma2s(rj01, h[j], make_float2(e4.x, e4.y));
        ffma2s(rj23, h[j], make_float2(e4.z, e4.w));
        if (AUX) {
          const float4 f4 = ex24[c0 + j];
          ffma2s(r201, h[j], make_float2(f4.x, f4.y));
          ffma2s(r223, h[j], make_float2(f4.z, f4.w));
        }
      };
      float4* hdst = reinterpret_cast<float4*>(hstage + m * N + c0);
#pragma unroll
      for (int j = 0; j < NV_LAST; ++j) accumulate(j);
#pragma unroll
      for (int j4 = 0; j4 < NV_LAST / 4; ++j4) hdst[j4] = make_float4(h[4 * j4], h[4 * j4 + 1], h[4 * j4 + 2], h[4 * j4 + 3]);
      if (!last) {
#pragma unroll
        for (int j = NV_LAST; j < COLS; ++j) accumulate(j);
#pragma unroll
        for (int j4 = NV_LAST / 4; j4 < COLS / 4; ++j4) hdst[j4] = make_float4(h[4 * j4], h[4 * j4 + 1], h[4 * j4 + 2], h[4 * j4 + 3]);
      }
      rj[0] = rj01.x; rj[1] = rj01.y; rj[2] = rj23.x; rj[3] = rj23.y;
      r2[0] = r201.x; r2[1] = r201.y; r2[2] = r223.x; r2[3] = r223.y;
    }
    // LRd[i][j] = 1e-4 (sum_m Ex_i(m) R_j(m) - mm sum HR) / (1 - mm),  R_j(m) = sum_n HR[m][n] Ex_j(n)
    {
      float p[16];
      const float4 em = m < N ? ex4[m] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float ei[4] = {em.x, em.y, em.z, em.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) p[i * 4 + j] = ei[i] * rj[j];
      }
      // transposing warp reduction: every exchange halves the values a lane carries (16 shuffles instead of 80);
      // lane l ends with the warp sum of value (l >> 1)
#pragma unroll
      for (int w = 8, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int k = 0; k < w; ++k) {
          const float keep = up ? p[k + w] : p[k], send = up ? p[k] : p[k + w];
          p[k] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
      }
      p[0] += __shfl_xor_sync(0xffffffffu, p[0], 1);
      const float tot_w = warp_sum(m < N ? rs : 0.f);
      fence_proxy_async();                              // the staged HR rows -> visible to the bulk-copy engine
      if (AUX && seg > 0 && m < N) {                    // upper column segments -> the segment-0 thread of the same row
        uint8_t* xs = sm + L::OFF_XCH + (uint32_t)(seg - 1) * L::XCH_SEG;
        reinterpret_cast<float4*>(xs)[m] = make_float4(rj[0], rj[1], rj[2], rj[3]);
        reinterpret_cast<float4*>(xs + 1600)[m] = make_float4(r2[0], r2[1], r2[2], r2[3]);
        reinterpret_cast<float*>(xs + 3200)[m] = rs;
      }
      if ((lane & 1) == 0) reds[warp * 20 + (lane >> 1)] = p[0];
      if (lane == 0) reds[warp * 20 + 16] = tot_w;
      __syncthreads();
      if (tid == 32) {                                  // (all rows are staged)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(HR + (size_t)b * N * N),
                     "r"(base + L::OFF_STAGE), "r"(PLANE_BYTES)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (tid < 16) {
        float s = 0.f, tot = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) { s += reds[w * 20 + tid]; tot += reds[w * 20 + 16]; }
        const float mm = expf(-100.0f / gamma);
        LRd[b * 16 + tid] = 1e-4f * (s - mm * tot) / (1.0f - mm);
      }
      if (AUX && seg == 0 && m < N) {
        float a[9] = {rj[0], rj[1], rj[2], rj[3], r2[0], r2[1], r2[2], r2[3], rs};
#pragma unroll
        for (int sg = 0; sg < SEGS - 1; ++sg) {         // fixed order: deterministic
          const uint8_t* xs = sm + L::OFF_XCH + (uint32_t)sg * L::XCH_SEG;
          const float4 a0 = reinterpret_cast<const float4*>(xs)[m], a1 = reinterpret_cast<const float4*>(xs + 1600)[m];
          a[0] += a0.x; a[1] += a0.y; a[2] += a0.z; a[3] += a0.w;
          a[4] += a1.x; a[5] += a1.y; a[6] += a1.z; a[7] += a1.w;
          a[8] += reinterpret_cast<const float*>(xs + 3200)[m];
        }
        float4* dst = reinterpret_cast<float4*>(aux + (size_t)b * AUX_STRIDE + m * AUX_ROW);
        dst[0] = make_float4(a[0], a[1], a[2], a[3]);
        dst[1] = make_float4(a[4], a[5], a[6], a[7]);
        dst[2] = make_float4(a[8], 0.f, 0.f, 0.f);
      }
      if (AUX && tid == 0) aux[(size_t)b * AUX_STRIDE + 100 * AUX_ROW] = m2;
    }
    // the next sample's barriers (block_max2 and the two tile barriers) order everything above before its first MMA
  }

  if (tid == 32) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
  }
}


// ------------------------------------------------------------------------------------------------------------------
// backward of the PSF model for the training case (gradient arrives through LR_degrade only, train/tPSFNet_train.py:
// 186-189): d(alpha, beta, gamma) from dLRd, the depth plane and the forward's hand-over `aux`.
//   w(m,n)  = kk (sum_j Ex_j(n) Qt_j(m) - mm gsum),  Qt_j(m) = sum_i g_ij Ex_i(m),  kk = 1e-4 / (1 - mm), mm = exp(-100/gamma)
//   d alpha = sum_{non-contact} w HR / alpha = [sum_all w HR - m2 sum_contact w] / alpha   (HR = m2 on the contact set)
//   d beta  = alpha 2 cp2 / beta^3 * sum_{non-contact} w P3,   P3 = E2 D E + E D E2,  E2[m][k] = e(|k-m|) (k-m)^2
//   d gamma = closed form in G0 = sum g_ij S_ij, G1 = sum g_ij S1_ij, sum HR (psf.cu psf_bwd_kernel), all three linear in
//             the per-row statistics U, U2, row sum that the forward left in `aux`
// P3 takes four 100^3 contractions; they run as tcgen05.mma on fp16 (PASSES 3: hi/lo split) operands like the forward:
//   T = E D -> acc0;  T2 = E2 D -> acc1;  acc0' = T E2  (+)=  T2 E      (the two products share one accumulator: the
//   operand scales are chosen so that both carry 2^(12 - dexp))
// ------------------------------------------------------------------------------------------------------------------
template <int PASSES, int NT>
__global__ void __launch_bounds__(NT, 2)
psf_bwd_tc_kernel(const float* __restrict__ ab, const float* __restrict__ depth, const float* __restrict__ aux,
                  const float* __restrict__ dLRd, float* __restrict__ dab, int B) {
  using L = Lay<PASSES, NT>;
  constexpr int NW = L::NW, SEGS = L::SEGS, COLS = L::COLS, IPT = ipt<NT>();
  extern __shared__ __align__(1024) uint8_t sm[];
  const uint32_t base = smem_u32(sm);
  float* tab = reinterpret_cast<float*>(sm + L::OFF_TAB);
  uint32_t* tab2 = reinterpret_cast<uint32_t*>(sm + L::OFF_TAB2);
  float4* ex4 = reinterpret_cast<float4*>(sm + L::OFF_EX);
  uint8_t* maskb = sm + L::OFF_MASK;
  float* red = reinterpret_cast<float*>(sm + L::OFF_RED);
  float* reds = red + 2 * NW;
  const uint32_t bar1 = base + L::OFF_BAR, tmem_slot = bar1 + 24u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + L::OFF_BAR + 24);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if ((base & 1023u) != 0u) {
    if (tid == 0) printf("tactilesr_b200 psf_tc: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  for (uint32_t i = tid; i < L::TILES_END / 16; i += NT) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (uint32_t i = tid; i < 104 * 16 / 4; i += NT) reinterpret_cast<uint32_t*>(maskb)[i] = 0u;
  if (tid == 0) {
    mbar_init(bar1, 1);
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
  const uint32_t idesc1 = make_idesc(128, 112, 0, 1, 0, 0);   // A K-major, B MN-major (depth as it lies in HBM)
  const uint32_t idesc2 = make_idesc(128, 112, 0, 0, 0, 0);   // both K-major

  const int q = warp & 3, seg = warp >> 2;
  const int m = q * 32 + lane;
  const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
  const uint32_t acc0 = tmem_base + lane_addr, acc1 = tmem_base + 128u + lane_addr;
  uint8_t* const x_hi = sm + L::OFF_D; uint8_t* const x_lo = sm + L::OFF_D + TILE;
  uint8_t* const e_hi = sm + L::OFF_E; uint8_t* const e_lo = sm + L::OFF_E + TILE;
  const uint32_t a_x_hi = base + L::OFF_D, a_x_lo = a_x_hi + TILE, a_e_hi = base + L::OFF_E, a_e_lo = a_e_hi + TILE;

  float dreg[IPT][8];
  if ((int)blockIdx.x < B) load_plane<false, NT>(depth + (size_t)blockIdx.x * N * N, tid, dreg);
  uint32_t nph = 0;                       // completed phases of bar1 (4 per sample)
  // ONE warp polls the MMA-completion mbarrier, the others wait at the block barrier behind it: 256 polling threads
  // cost 9 % of the kernel's issue slots, which the co-resident CTA needs
  auto wait_mma = [&]() {
    if (warp == 1) {
      mbar_wait(bar1, nph & 1u);
      tc_fence_before();
    }
    ++nph;
    __syncthreads();
    tc_fence_after();
  };
  auto publish = [&]() {                  // generic-proxy smem writes -> visible to the MMA; all TMEM reads retired
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
  };

  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const float alpha = ab[b * 3 + 0], beta = ab[b * 3 + 1], gamma = ab[b * 3 + 2];

    // ---- tables, depth max, E and depth tiles ----
    float dmax, amax;
    plane_max<NT, false>(dreg, tid, dmax, amax);
    build_tables<NT>(beta, gamma, tab, ex4, tid);
    block_max2<NT>(dmax, amax, red);
    const float thr = dmax - 1e-3f;
    int dexp = 0;
    if (amax > 0.f) (void)frexpf(amax, &dexp);
    const float sD = ldexpf(1.0f, 4 - dexp);
    build_tab2<0, NT>(tab, tab2, tid);
    store_plane_tiles<PASSES, NT, false>(dreg, sD, thr, x_hi, x_lo, maskb, tid);
    __syncthreads();
    build_toeplitz_tiles<PASSES, NT>(tab2, e_hi, e_lo, tid);
    publish();

    // ---- GEMM 1: T = E D -> acc0  (2^(8 - dexp) T) ----
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        issue_gemm_kmn<PASSES>(tmem_base, a_e_hi, a_e_lo, a_x_hi, a_x_lo, idesc1);
        umma_commit(bar1);
      }
      __syncwarp();
    }
    if (b + (int)gridDim.x < B) load_plane<false, NT>(depth + (size_t)(b + gridDim.x) * N * N, tid, dreg);
    build_tab2<1, NT>(tab, tab2, tid);    // (all reads of the E table finished before publish())
    wait_mma();                           // (its barrier also publishes the table)
    build_toeplitz_tiles<PASSES, NT>(tab2, e_hi, e_lo, tid);      // E2 over E
    publish();

    // ---- GEMM 2: T2 = E2 D -> acc1  (2^(8 - dexp) T2) ----
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        issue_gemm_kmn<PASSES>(tmem_base + 128u, a_e_hi, a_e_lo, a_x_hi, a_x_lo, idesc1);
        umma_commit(bar1);
      }
      __syncwarp();
    }
    build_tab2<2, NT>(tab, tab2, tid);    // 2^15 e(t) for the last product
    wait_mma();
    acc_to_tiles<PASSES, COLS>(acc0, 1.0f, x_hi, x_lo, m, seg);   // T over the depth tiles
    publish();

    // ---- GEMM 3: acc0 = T E2  (2^(12 - dexp)) ----
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        issue_gemm_kk<PASSES>(tmem_base, a_x_hi, a_x_lo, a_e_hi, a_e_lo, idesc2);
        umma_commit(bar1);
      }
      __syncwarp();
    }
    wait_mma();
    acc_to_tiles<PASSES, COLS>(acc1, 1.0f / 2048.0f, x_hi, x_lo, m, seg);   // 2^(-3 - dexp) T2 over T;  |T2| < 2^(18 + dexp)
    build_toeplitz_tiles<PASSES, NT>(tab2, e_hi, e_lo, tid);                // 2^15 E over E2
    publish();

    // ---- GEMM 4: acc0 += T2 E ----
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        issue_gemm_kk<PASSES, 1>(tmem_base, a_x_hi, a_x_lo, a_e_hi, a_e_lo, idesc2);
        umma_commit(bar1);
      }
      __syncwarp();
    }
    // per-row factors while the MMAs run
    float g[16], gsum = 0.f;
#pragma unroll
    for (int t = 0; t < 16; ++t) { g[t] = dLRd[b * 16 + t]; gsum += g[t]; }
    const float mm = expf(-100.0f / gamma);
    const float mg = mm * gsum;
    const float4 em = m < N ? ex4[m] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float ei[4] = {em.x, em.y, em.z, em.w};
    float qt[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) qt[j] = ei[0] * g[j] + ei[1] * g[4 + j] + ei[2] * g[8 + j] + ei[3] * g[12 + j];
    float G0 = 0.f, G1 = 0.f, tot = 0.f;
    if (seg == 0 && m < N) {              // d gamma / d alpha statistics from the forward's per-row sums
      const float4* a4 = reinterpret_cast<const float4*>(aux + (size_t)b * AUX_STRIDE + m * AUX_ROW);
      const float4 u = a4[0], u2 = a4[1];
      tot = a4[2].x;
      const float uu[4] = {u.x, u.y, u.z, u.w}, vv[4] = {u2.x, u2.y, u2.z, u2.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float ai = 0.f, bi = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) { ai = fmaf(g[i * 4 + j], uu[j], ai); bi = fmaf(g[i * 4 + j], vv[j], bi); }
        const float d = (float)(m - 12 - 25 * i);
        G0 = fmaf(ei[i], ai, G0);
        G1 = fmaf(ei[i] * (d * d), ai, fmaf(ei[i], bi, G1));
      }
    }
    const uint64_t bits = contact_bits<COLS>(maskb, m < N ? m : 0, seg);
    wait_mma();

    // ---- epilogue over acc0 = 2^(12 - dexp) P3: this thread's COLS columns, 28 at a time ----
    float dbs = 0.f, cw = 0.f;
    const int c0 = seg * COLS;
#pragma unroll
    for (int part = 0; part < COLS / 28; ++part) {
      uint32_t v[28];
      tmem_ld_cols<28>(acc0 + (uint32_t)(c0 + part * 28), v);
      tmem_ld_wait();
      if (m < N) {
#pragma unroll
        for (int jj = 0; jj < 28; ++jj) {
          const int j = part * 28 + jj;
          if (j < L::NV_LAST || seg < SEGS - 1) {
            const float4 e4 = ex4[c0 + j];
            const float wc = fmaf(e4.x, qt[0], fmaf(e4.y, qt[1], fmaf(e4.z, qt[2], e4.w * qt[3])));
            const bool contact = bitof(bits, j);
            dbs = fmaf(contact ? 0.f : wc - mg, __uint_as_float(v[jj]), dbs);
            cw += contact ? wc : 0.f;
          }
        }
      }
    }
    const float cnt = m < N ? (float)__popcll(bits) : 0.f;
    {
      float p[6] = {dbs, cw, cnt, G0, G1, tot};
#pragma unroll
      for (int k = 0; k < 6; ++k) p[k] = warp_sum(p[k]);
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) reds[warp * 20 + k] = p[k];
      }
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        float r[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int w = 0; w < NW; ++w)
#pragma unroll
          for (int k = 0; k < 6; ++k) r[k] += reds[w * 20 + k];
        const float om = 1.0f - mm, kk = 1e-4f / om;
        const float m2 = aux[(size_t)b * AUX_STRIDE + 100 * AUX_ROW];
        const float s_all = kk * (r[3] - mg * r[5]);                 // sum over all pixels of w HR
        const float s_c = m2 * kk * (r[1] - mg * r[2]);              // its contact part (HR = m2 there)
        const float inv_g2 = 1.0f / (gamma * gamma);
        const float mp = mm * 100.0f * inv_g2;
        dab[b * 3 + 0] = (s_all - s_c) / alpha;
        dab[b * 3 + 1] = alpha * 2.0f * CP2 / (beta * beta * beta) * kk * ldexpf(1.0f, dexp - 12) * r[0];
        dab[b * 3 + 2] = 1e-4f * ((CM2 * inv_g2 * r[4] - mp * r[5] * gsum) / om + (r[3] - mm * r[5] * gsum) * mp / (om * om));
      }
    }
    // (the barriers of the next sample's tile phase order everything above before its first MMA)
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
  }
}


// ------------------------------------------------------------------------------------------------------------------
// The same backward for the 16-bit precision modes: single fp16 operands halve every tile, so E and E2 AND T and T2 fit
// in shared memory at once (the four tile slots of the split layout).  The products then run as two groups of two --
// [T = E D | T2 = E2 D] and [P3 = T E2 + T2 E] -- with one completion wait each, E / E2 come out of ONE table pass, and a
// sample takes 7 block barriers instead of 11.  Scales: E tile 2^15 e, E2 tile 16 e t^2, both accumulators are converted
// with 2^-11, so both halves of P3 carry 2^(12 - dexp).
// ------------------------------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT, 2)
psf_bwd_f16_kernel(const float* __restrict__ ab, const float* __restrict__ depth, const float* __restrict__ aux,
                   const float* __restrict__ dLRd, float* __restrict__ dab, int B) {
  using L = Lay<3, NT>;                    // slots: [D -> T | T2 | E | E2]
  constexpr int NW = L::NW, SEGS = L::SEGS, COLS = L::COLS, IPT = ipt<NT>();
  extern __shared__ __align__(1024) uint8_t sm[];
  const uint32_t base = smem_u32(sm);
  float* tab = reinterpret_cast<float*>(sm + L::OFF_TAB);
  uint32_t* tab2 = reinterpret_cast<uint32_t*>(sm + L::OFF_TAB2);
  float4* ex4 = reinterpret_cast<float4*>(sm + L::OFF_EX);
  uint8_t* maskb = sm + L::OFF_MASK;
  float* red = reinterpret_cast<float*>(sm + L::OFF_RED);
  float* reds = red + 2 * NW;
  const uint32_t bar1 = base + L::OFF_BAR, tmem_slot = bar1 + 24u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + L::OFF_BAR + 24);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if ((base & 1023u) != 0u) {
    if (tid == 0) printf("tactilesr_b200 psf_tc: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  for (uint32_t i = tid; i < L::TILES_END / 16; i += NT) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (uint32_t i = tid; i < 104 * 16 / 4; i += NT) reinterpret_cast<uint32_t*>(maskb)[i] = 0u;
  if (tid == 0) {
    mbar_init(bar1, 1);
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
  const uint32_t idesc1 = make_idesc(128, 112, 0, 1, 0, 0);   // A K-major, B MN-major (depth as it lies in HBM)
  const uint32_t idesc2 = make_idesc(128, 112, 0, 0, 0, 0);   // both K-major

  const int q = warp & 3, seg = warp >> 2;
  const int m = q * 32 + lane;
  const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
  const uint32_t acc0 = tmem_base + lane_addr, acc1 = tmem_base + 128u + lane_addr;
  uint8_t* const x1 = sm + L::OFF_D; uint8_t* const x2 = sm + L::OFF_D + TILE;
  uint8_t* const e1 = sm + L::OFF_E; uint8_t* const e2 = sm + L::OFF_E + TILE;
  const uint32_t a_x1 = base + L::OFF_D, a_x2 = a_x1 + TILE, a_e1 = base + L::OFF_E, a_e2 = a_e1 + TILE;

  float dreg[IPT][8];
  if ((int)blockIdx.x < B) load_plane<false, NT>(depth + (size_t)blockIdx.x * N * N, tid, dreg);
  uint32_t nph = 0;                       // completed phases of bar1 (2 per sample)
  auto wait_mma = [&]() {                 // one warp polls, the others wait at the block barrier behind it
    if (warp == 1) {
      mbar_wait(bar1, nph & 1u);
      tc_fence_before();
    }
    ++nph;
    __syncthreads();
    tc_fence_after();
  };
  auto publish = [&]() {                  // generic-proxy smem writes -> visible to the MMA; all TMEM reads retired
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
  };

  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const float alpha = ab[b * 3 + 0], beta = ab[b * 3 + 1], gamma = ab[b * 3 + 2];

    // ---- tables, depth max, depth tile, E and E2 tiles ----
    float dmax, amax;
    plane_max<NT, false>(dreg, tid, dmax, amax);
    build_tables<NT>(beta, gamma, tab, ex4, tid);
    block_max2<NT>(dmax, amax, red);
    const float thr = dmax - 1e-3f;
    int dexp = 0;
    if (amax > 0.f) (void)frexpf(amax, &dexp);
    const float sD = ldexpf(1.0f, 4 - dexp);
    build_tab2<3, NT>(tab, tab2, tid);
    store_plane_tiles<1, NT, false>(dreg, sD, thr, x1, x1, maskb, tid);
    __syncthreads();
    build_toeplitz_tiles<3, NT>(tab2, e1, e2, tid);           // "hi" halves -> 2^15 E, "lo" halves -> 16 E2
    publish();

    // ---- T = E D -> acc0 (2^(19 - dexp) T),  T2 = E2 D -> acc1 (2^(8 - dexp) T2) ----
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        issue_gemm_kmn<1>(tmem_base, a_e1, a_e1, a_x1, a_x1, idesc1);
        issue_gemm_kmn<1>(tmem_base + 128u, a_e2, a_e2, a_x1, a_x1, idesc1);
        umma_commit(bar1);
      }
      __syncwarp();
    }
    if (b + (int)gridDim.x < B) load_plane<false, NT>(depth + (size_t)(b + gridDim.x) * N * N, tid, dreg);
    wait_mma();
    acc_to_tiles<1, COLS>(acc0, 1.0f / 2048.0f, x1, x1, m, seg);     // 2^(8 - dexp) T over the depth tile
    acc_to_tiles<1, COLS>(acc1, 1.0f / 2048.0f, x2, x2, m, seg);     // 2^(-3 - dexp) T2;  |T2| < 2^(18 + dexp)
    publish();

    // ---- acc0 = T E2 + T2 E  (2^(12 - dexp) P3) ----
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        issue_gemm_kk<1>(tmem_base, a_x1, a_x1, a_e2, a_e2, idesc2);
        issue_gemm_kk<1, 1>(tmem_base, a_x2, a_x2, a_e1, a_e1, idesc2);
        umma_commit(bar1);
      }
      __syncwarp();
    }
    // per-row factors while the MMAs run
    float g[16], gsum = 0.f;
#pragma unroll
    for (int t = 0; t < 16; ++t) { g[t] = dLRd[b * 16 + t]; gsum += g[t]; }
    const float mm = expf(-100.0f / gamma);
    const float mg = mm * gsum;
    const float4 em = m < N ? ex4[m] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float ei[4] = {em.x, em.y, em.z, em.w};
    float qt[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) qt[j] = ei[0] * g[j] + ei[1] * g[4 + j] + ei[2] * g[8 + j] + ei[3] * g[12 + j];
    float G0 = 0.f, G1 = 0.f, tot = 0.f;
    if (seg == 0 && m < N) {              // d gamma / d alpha statistics from the forward's per-row sums
      const float4* a4 = reinterpret_cast<const float4*>(aux + (size_t)b * AUX_STRIDE + m * AUX_ROW);
      const float4 u = a4[0], u2 = a4[1];
      tot = a4[2].x;
      const float uu[4] = {u.x, u.y, u.z, u.w}, vv[4] = {u2.x, u2.y, u2.z, u2.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float ai = 0.f, bi = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) { ai = fmaf(g[i * 4 + j], uu[j], ai); bi = fmaf(g[i * 4 + j], vv[j], bi); }
        const float d = (float)(m - 12 - 25 * i);
        G0 = fmaf(ei[i], ai, G0);
        G1 = fmaf(ei[i] * (d * d), ai, fmaf(ei[i], bi, G1));
      }
    }
    const uint64_t bits = contact_bits<COLS>(maskb, m < N ? m : 0, seg);
    wait_mma();

    // ---- epilogue over acc0 = 2^(12 - dexp) P3: this thread's COLS columns, 28 at a time ----
    float dbs = 0.f, cw = 0.f;
    const int c0 = seg * COLS;
#pragma unroll
    for (int part = 0; part < COLS / 28; ++part) {
      uint32_t v[28];
      tmem_ld_cols<28>(acc0 + (uint32_t)(c0 + part * 28), v);
      tmem_ld_wait();
      if (m < N) {
#pragma unroll
        for (int jj = 0; jj < 28; ++jj) {
          const int j = part * 28 + jj;
          if (j < L::NV_LAST || seg < SEGS - 1) {
            const float4 e4 = ex4[c0 + j];
            const float wc = fmaf(e4.x, qt[0], fmaf(e4.y, qt[1], fmaf(e4.z, qt[2], e4.w * qt[3])));
            const bool contact = bitof(bits, j);
            dbs = fmaf(contact ? 0.f : wc - mg, __uint_as_float(v[jj]), dbs);
            cw += contact ? wc : 0.f;
          }
        }
      }
    }
    const float cnt = m < N ? (float)__popcll(bits) : 0.f;
    {
      float p[6] = {dbs, cw, cnt, G0, G1, tot};
#pragma unroll
      for (int k = 0; k < 6; ++k) p[k] = warp_sum(p[k]);
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) reds[warp * 20 + k] = p[k];
      }
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        float r[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int w = 0; w < NW; ++w)
#pragma unroll
          for (int k = 0; k < 6; ++k) r[k] += reds[w * 20 + k];
        const float om = 1.0f - mm, kk = 1e-4f / om;
        const float m2 = aux[(size_t)b * AUX_STRIDE + 100 * AUX_ROW];
        const float s_all = kk * (r[3] - mg * r[5]);                 // sum over all pixels of w HR
        const float s_c = m2 * kk * (r[1] - mg * r[2]);              // its contact part (HR = m2 there)
        const float inv_g2 = 1.0f / (gamma * gamma);
        const float mp = mm * 100.0f * inv_g2;
        dab[b * 3 + 0] = (s_all - s_c) / alpha;
        dab[b * 3 + 1] = alpha * 2.0f * CP2 / (beta * beta * beta) * kk * ldexpf(1.0f, dexp - 12) * r[0];
        dab[b * 3 + 2] = 1e-4f * ((CM2 * inv_g2 * r[4] - mp * r[5] * gsum) / om + (r[3] - mm * r[5] * gsum) * mp / (om * om));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
  }
}

}  // namespace

int g_psf_mode = 0;     // 0 = tensor-core forward, fp32-accurate (default), 1 = FFMA forward (psf.cu), 2 = tensor cores, one fp16 pass

static int psf_tc_grid(int B) {
  const int sms = num_sms();
  return B < 2 * sms ? B : 2 * sms;
}

template <int PASSES, bool AUX, int NT>
static int psf_forward_tc_launch(const float* alphaBeta, const float* depth, float* HR, float* LRd, float* psf, float* aux, int B,
                                 cudaStream_t stream) {
  const size_t smem = Lay<PASSES, NT>::BYTES;
  TSR_CUDA(cudaFuncSetAttribute(psf_fwd_tc_kernel<PASSES, AUX, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  psf_fwd_tc_kernel<PASSES, AUX, NT><<<psf_tc_grid(B), NT, smem, stream>>>(alphaBeta, depth, HR, LRd, psf, aux, B);
  TSR_CHECK_LAUNCH("psf_forward_tc");
  return TSR_OK;
}

template <int PASSES>
static int psf_forward_tc_impl(const float* alphaBeta, const float* depth, float* HR, float* LRd, float* psf, float* aux, int B,
                               cudaStream_t stream) {
  TSR_REQUIRE(alphaBeta && depth && HR && LRd && B > 0, "psf_forward_tc: bad argument");
  TSR_REQUIRE(((uintptr_t)depth & 15) == 0 && ((uintptr_t)HR & 15) == 0 && ((uintptr_t)aux & 15) == 0,
              "psf_forward_tc: depth / HR / aux must be 16-byte aligned");
  return aux ? psf_forward_tc_launch<PASSES, true, PSF_THREADS>(alphaBeta, depth, HR, LRd, psf, aux, B, stream)
             : psf_forward_tc_launch<PASSES, false, PSF_THREADS>(alphaBeta, depth, HR, LRd, psf, aux, B, stream);
}

template <int PASSES, int NT>
static int psf_backward_tc_launch(const float* alphaBeta, const float* depth, const float* aux, const float* dLRd,
                                  float* dalphaBeta, int B, cudaStream_t stream) {
  const size_t smem = Lay<PASSES, NT>::BYTES;
  TSR_CUDA(cudaFuncSetAttribute(psf_bwd_tc_kernel<PASSES, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  psf_bwd_tc_kernel<PASSES, NT><<<psf_tc_grid(B), NT, smem, stream>>>(alphaBeta, depth, aux, dLRd, dalphaBeta, B);
  TSR_CHECK_LAUNCH("psf_backward_tc");
  return TSR_OK;
}

template <int PASSES>
static int psf_backward_tc_impl(const float* alphaBeta, const float* depth, const float* aux, const float* dLRd,
                                float* dalphaBeta, int B, cudaStream_t stream) {
  TSR_REQUIRE(alphaBeta && depth && aux && dLRd && dalphaBeta && B > 0, "psf_backward_tc: bad argument");
  TSR_REQUIRE(((uintptr_t)depth & 15) == 0 && ((uintptr_t)aux & 15) == 0, "psf_backward_tc: depth / aux must be 16-byte aligned");
  return psf_backward_tc_launch<PASSES, PSF_THREADS>(alphaBeta, depth, aux, dLRd, dalphaBeta, B, stream);
}

extern "C" {

void tsr_set_psf_mode(int mode) { g_psf_mode = mode; }
int tsr_get_psf_mode(void) { return g_psf_mode; }

int tsr_psf_forward_ffma(const float* alphaBeta, const float* depth, float* HR, float* LRd, float* psf, int B,
                         cudaStream_t stream);

size_t tsr_psf_aux_floats(void) { return (size_t)AUX_STRIDE; }

// fp32-accurate (fp16 hi/lo split operands).  aux (B x tsr_psf_aux_floats() floats, or NULL): per-row statistics of HR for
// tsr_psf_backward_tc
int tsr_psf_forward_tc(const float* alphaBeta, const float* depth, float* HR, float* LRd, float* psf, float* aux, int B,
                       cudaStream_t stream) {
  return psf_forward_tc_impl<3>(alphaBeta, depth, HR, LRd, psf, aux, B, stream);
}
// one fp16 pass (the 16-bit tensor-core precision modes)
int tsr_psf_forward_tc_f16(const float* alphaBeta, const float* depth, float* HR, float* LRd, float* psf, float* aux, int B,
                           cudaStream_t stream) {
  return psf_forward_tc_impl<1>(alphaBeta, depth, HR, LRd, psf, aux, B, stream);
}

// d alphaBeta (B,3) from dLRd (B,16) alone (the training case), from the depth planes and the forward's aux
int tsr_psf_backward_tc(const float* alphaBeta, const float* depth, const float* aux, const float* dLRd,
                        float* dalphaBeta, int B, cudaStream_t stream) {
  return psf_backward_tc_impl<3>(alphaBeta, depth, aux, dLRd, dalphaBeta, B, stream);
}
int tsr_psf_backward_tc_f16(const float* alphaBeta, const float* depth, const float* aux, const float* dLRd,
                            float* dalphaBeta, int B, cudaStream_t stream) {
  TSR_REQUIRE(alphaBeta && depth && aux && dLRd && dalphaBeta && B > 0, "psf_backward_tc_f16: bad argument");
  TSR_REQUIRE(((uintptr_t)depth & 15) == 0 && ((uintptr_t)aux & 15) == 0, "psf_backward_tc_f16: depth / aux must be 16-byte aligned");
  const size_t smem = Lay<3, PSF_THREADS>::BYTES;
  TSR_CUDA(cudaFuncSetAttribute(psf_bwd_f16_kernel<PSF_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  psf_bwd_f16_kernel<PSF_THREADS><<<psf_tc_grid(B), PSF_THREADS, smem, stream>>>(alphaBeta, depth, aux, dLRd, dalphaBeta, B);
  TSR_CHECK_LAUNCH("psf_backward_tc_f16");
  return TSR_OK;
}

// (HR, LRd, psf) = PSF forward model of `depth` (B,100,100) under alphaBeta (B,3).  psf may be NULL.
int tsr_psf_forward(const float* alphaBeta, const float* depth, float* HR, float* LRd, float* psf, int B,
                    cudaStream_t stream) {
  if (g_psf_mode == 1) return tsr_psf_forward_ffma(alphaBeta, depth, HR, LRd, psf, B, stream);
  if (g_psf_mode == 2) return psf_forward_tc_impl<1>(alphaBeta, depth, HR, LRd, psf, nullptr, B, stream);
  return psf_forward_tc_impl<3>(alphaBeta, depth, HR, LRd, psf, nullptr, B, stream);
}

}  // extern "C"
