// Fused CP convolution on channels-last rows (bf16, 3x3, padding 1, 56-wide maps): the WHOLE chain
//     y = F_out . dw3x3( F_in^T . x )            (reference: tltorch/functional/convolution.py:231-283, cp_conv)
// in one launch per direction, R-channel intermediates never leave the SM.
//
// Layout of the work on an SM (one persistent CTA per SM, 512 threads):
//   * activations are (B, H, W, C) rows; a CHUNK is two image rows = 112 positions x C channels, one 4-D TMA box (rows outside
//     the image are zero-filled by TMA: that IS the convolution's zero padding, the first 1x1 stage has no bias);
//   * stage 1 runs TRANSPOSED on tcgen05: D1[r, pos] = A[r, c] . X[pos, c]^T with the rank on the 128 TMEM lanes and the 112
//     positions on the columns, into a 3-slot TMEM ring (3 x 112 fp32 columns);
//   * the depthwise stage is thread-local: a thread owns ONE rank channel (its TMEM lane, its six separable taps in registers),
//     reads its row of positions with tcgen05.ld, runs the H pass and the W pass in fp32 registers and stores bf16 z2 as the
//     MN-major A operand of stage 3 -- no shared-memory round trip of z1, no bf16 unpacking, no halo exchange between threads;
//   * stage 3: D2[pos, o] = Z2[pos, r] . B[o, r]^T (positions on the lanes), epilogue warps pack bf16 rows and TMA-store them.
// Three modes share the skeleton:
//   FWD         in = x,  A = F_in^T,  B = F_out                         -> y
//   BWD_DATA    in = dy, A = F_out^T, B = F_in, flipped taps            -> dx, and dF_in  += dz1^T . x   (x centre chunks)
//   BWD_WEIGHT  in = x,  A = F_in^T (z1 ring), dz2 centre = F_out^T . dy -> dF_out += z2^T . dy, dT[9] += z1(+tap) * dz2
// The weight-gradient accumulators live in TMEM for the whole kernel ([rank lane][channel column]) and are flushed once per CTA
// with fp32 reductions; the nine tap gradients are per-thread registers (thread = rank channel).
// Stride 2 (FWD only): a unit is one output row (28 positions), four units fill one stage-3 tile.
#include "cprows_common.cuh"

namespace tlt {
namespace cpr {

constexpr int kThreads = 512;          // warps: 0 TMA | 1 stage-1 MMA | 2 TMEM alloc | 3 stage-3 MMA | 4-11 depthwise (2 groups x 4) | 12-15 epilogue
constexpr int NS = 3;                  // TMEM ring slots
constexpr int XS_MAX = 8;               // input-chunk smem stages: per configuration (Cfg::XS), as deep as shared memory allows
constexpr int OS_MAX = 4;
constexpr int kZBytes = 2 * 128 * 128; // z2 tile: 2 blocks of 64 positions x 128 rank rows x 128 B

enum { MODE_FWD = 0, MODE_BWD_DATA = 1, MODE_BWD_WEIGHT = 2 };

struct Params {
  int batch, H, W, Ho, Wo;
  int R, KR;                 // rank, rank rounded up to a multiple of 16
  int gpi;                   // stage-3 tiles (groups) per image
  int groups_total;
  const float* taps;         // [6][128] fp32: kh0 kh1 kh2 kw0 kw1 kw2 (lambda folded into kw; reversed for BWD_DATA)
  const float* bias;         // fp32 [C_OUT] or nullptr (FWD)
  float* dF;                 // BWD_*: fp32 [C_OTHER][128] accumulators
  float* dT;                 // BWD_WEIGHT: fp32 [9][128]
};

template <int MODE, int STRIDE, int C_IN, int C_OUT> struct Cfg {
  static constexpr int KB_IN = C_IN / 64;
  static constexpr int KB_OUT = C_OUT / 64;
  static constexpr int G = STRIDE == 1 ? 1 : 4;                    // units per group
  static constexpr bool HAS_OUT = MODE != MODE_BWD_WEIGHT;         // stage 3 + store
  static constexpr bool HAS_OTHER = MODE != MODE_FWD;              // centre chunks of the other activation tensor
  static constexpr int C_OTHER = MODE == MODE_BWD_DATA ? C_OUT : (MODE == MODE_BWD_WEIGHT ? C_OUT : 0);
  // In BWD_DATA the kernel's "in" tensor is dy (C_IN = layer C_out) and its output is dx (C_OUT = layer C_in); the other tensor is
  // x (C_OUT channels).  In BWD_WEIGHT "in" is x (C_IN), the other tensor is dy (C_OUT channels).
  static constexpr int KB_OTHER = C_OTHER / 64;
  static constexpr int ND = (MODE == MODE_FWD && NS * NP + 2 * C_OUT <= 512) ? 2 : 1;
  // TMA ring depths: HBM latency x the SM's share of the bandwidth is ~60 KB in flight; three 14 KB stages held the forward
  // kernel at 2.1 TB/s (the depthwise warps waited on stage 1), so every configuration takes what its shared memory allows
  static constexpr int XS = MODE == MODE_FWD ? (C_OUT == 64 ? 6 : 4) : 4;
  static constexpr int OS = 3;
  // TMEM columns
  static constexpr int COL_D = NS * NP;                            // stage-3 accumulator(s)
  static constexpr int COL_C = NS * NP;                            // BWD_WEIGHT: dz2 centre
  static constexpr int COL_F = MODE == MODE_BWD_DATA ? COL_D + C_OUT : COL_C + NP;   // weight-gradient accumulator
  // shared memory
  static constexpr int OFF_A = 0;
  static constexpr int OFF_A2 = OFF_A + KB_IN * 16384;                                   // BWD_WEIGHT: F_out^T
  static constexpr int OFF_B = OFF_A2 + (MODE == MODE_BWD_WEIGHT ? KB_OUT * 16384 : 0);  // [2 rank blocks][C_OUT rows x 128 B]
  static constexpr int OFF_X = OFF_B + (HAS_OUT ? 2 * C_OUT * 128 : 0);
  static constexpr int X_STAGE = KB_IN * kChunkBytes;
  static constexpr int OFF_Z = OFF_X + XS * X_STAGE;
  static constexpr int OFF_Y = OFF_Z + 2 * kZBytes;
  static constexpr int Y_STAGE = KB_OUT * kChunkBytes;
  static constexpr int OFF_O = OFF_Y + (HAS_OUT ? 2 * Y_STAGE : 0);
  static constexpr int O_STAGE = KB_OTHER * kChunkBytes;
  static constexpr int OFF_BAR = OFF_O + (HAS_OTHER ? OS * O_STAGE : 0);
  static constexpr int SMEM = OFF_BAR + 512 + 1024;
};

// ---- barrier slots ---------------------------------------------------------------------------------------------------------
enum { B_XFULL = 0, B_XEMPTY = B_XFULL + XS_MAX, B_SFULL = B_XEMPTY + XS_MAX, B_SEMPTY = B_SFULL + NS, B_ZFULL = B_SEMPTY + NS,
       B_ZEMPTY = B_ZFULL + 2, B_DFULL = B_ZEMPTY + 2, B_DEMPTY = B_DFULL + 2, B_OFULL = B_DEMPTY + 2, B_OEMPTY = B_OFULL + OS_MAX,
       B_WFULL = B_OEMPTY + OS_MAX, B_CFULL, B_CEMPTY, B_FFULL, B_COUNT };

// --------------------------------------------------------------------------------------------------------------------------------
// Depthwise stage of one output row, stride 1.  t0/t1/t2: TMEM addresses (lane bits included) of column 0 of input rows h-1, h,
// h+1; the thread's channel is its lane.  z2[w] = sum_j kw[j] v[w + j - 1],  v[w] = sum_i kh[i] z1[h + i - 1][w].
// Column pieces of 8: the loads of piece p+1 are in flight while piece p is finished; outputs lag one column behind the loads,
// so a 16-byte group of 8 outputs completes when the first column of the next piece arrives.  Column pairs are packed (f2):
// H pass = 3 packed ops per pair; W pass = 2 packed ops (outer taps, both operands even-aligned pairs) + 2 scalar FMAs (centre tap).
// WGRAD (BWD_WEIGHT): tc = dz2 of row h (same columns); dT[3 i + j] += dz2[w] * z1[h + i - 1][w + j - 1].  The z1 operand is
// always an even-aligned pair; dz2 comes even-aligned (j = 1) and odd-aligned (a second tcgen05.ld one column further; j = 0, 2).
// kk: taps duplicated into pairs (kh0 kh1 kh2 kw0 kw1 kw2); k: the scalar taps.
// --------------------------------------------------------------------------------------------------------------------------------
template <bool WGRAD>
__device__ __forceinline__ void dw_row_s1(uint32_t t0, uint32_t t1, uint32_t t2, uint32_t tc, const f2 (&kk)[6], const float (&k)[6],
                                          uint32_t zbuf, int m_base, int r, f2 (&dT)[9]) {
  uint32_t a[8], b[8], c[8], d[8], e[8];
  float pend[7];
  f2 vp = 0ull;                                    // (v[8p-2], v[8p-1])
  f2 dprev = 0ull;                                 // WGRAD: (dz2[8p-1], dz2[8p])
  tmem_ld8(t0, a); tmem_ld8(t1, b); tmem_ld8(t2, c);
  if (WGRAD) { tmem_ld8(tc, d); tmem_ld8(tc + 1, e); }
#pragma unroll
  for (int p = 0; p < 7; ++p) {
    tmem_wait8(a); tmem_wait8(b); tmem_wait8(c);
    if (WGRAD) { tmem_wait8(d); tmem_wait8(e); }
    f2 za[4], zb[4], zc[4], v[4], de[4], dod[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      za[i] = pku(a[2 * i], a[2 * i + 1]); zb[i] = pku(b[2 * i], b[2 * i + 1]); zc[i] = pku(c[2 * i], c[2 * i + 1]);
      v[i] = fma2(kk[2], zc[i], fma2(kk[1], zb[i], mul2(kk[0], za[i])));
      if (WGRAD) {
        de[i] = pku(d[2 * i], d[2 * i + 1]);
        dod[i] = pku(e[2 * i], (p == 6 && i == 3) ? 0u : e[2 * i + 1]);      // column 56 does not exist
      }
    }
    if (WGRAD && p == 0) dprev = pku(0u, d[0]);
    if (p + 1 < 7) {
      tmem_ld8(t0 + 8 * (p + 1), a); tmem_ld8(t1 + 8 * (p + 1), b); tmem_ld8(t2 + 8 * (p + 1), c);
      if (WGRAD) { tmem_ld8(tc + 8 * (p + 1), d); tmem_ld8(tc + 8 * (p + 1) + 1, e); }
    }
    if (WGRAD) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const f2 dl = i == 0 ? dprev : dod[i - 1];     // centres one column left of the z pair  (tap j = 2)
        dT[0] = fma2(dod[i], za[i], dT[0]); dT[1] = fma2(de[i], za[i], dT[1]); dT[2] = fma2(dl, za[i], dT[2]);
        dT[3] = fma2(dod[i], zb[i], dT[3]); dT[4] = fma2(de[i], zb[i], dT[4]); dT[5] = fma2(dl, zb[i], dT[5]);
        dT[6] = fma2(dod[i], zc[i], dT[6]); dT[7] = fma2(de[i], zc[i], dT[7]); dT[8] = fma2(dl, zc[i], dT[8]);
      }
      dprev = dod[3];
    }
    // outputs 8p-1 .. 8p+6 as pairs (o[8p+2i-1], o[8p+2i]) = kw0 (v[2i-2], v[2i-1]) + kw2 (v[2i], v[2i+1]) + kw1 (v[2i-1], v[2i])
    float o[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const f2 vl = i == 0 ? vp : v[i - 1];
      const f2 t = fma2(kk[5], v[i], mul2(kk[3], vl));
      o[2 * i] = fmaf(k[4], hi_of(vl), lo_of(t));
      o[2 * i + 1] = fmaf(k[4], lo_of(v[i]), hi_of(t));
    }
    if (p > 0)
      st_shared_v4(z_addr(zbuf, m_base + 8 * (p - 1), r), pack_bf16x2(pend[0], pend[1]), pack_bf16x2(pend[2], pend[3]),
                   pack_bf16x2(pend[4], pend[5]), pack_bf16x2(pend[6], o[0]));
#pragma unroll
    for (int i = 0; i < 7; ++i) pend[i] = o[i + 1];
    vp = v[3];
  }
  const float o_last = fmaf(k[4], hi_of(vp), k[3] * lo_of(vp));               // column 55
  st_shared_v4(z_addr(zbuf, m_base + 48, r), pack_bf16x2(pend[0], pend[1]), pack_bf16x2(pend[2], pend[3]),
               pack_bf16x2(pend[4], pend[5]), pack_bf16x2(pend[6], o_last));
}

// --------------------------------------------------------------------------------------------------------------------------------
// Depthwise stage of one output row, stride 2 (28 outputs from 56 inputs): out[wo] = sum_j kw[j] v[2 wo + j - 1].  Output pieces
// of 8 = input pieces of 16; pieces [p_begin, p_end) of 4 (the last one holds 4 outputs from 8 input columns).
// --------------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dw_row_s2(uint32_t t0, uint32_t t1, uint32_t t2, const f2 (&kk)[6], const float (&k)[6], uint32_t zbuf,
                                          int m_base, int r, int p_begin, int p_end) {
  float vl = 0.f;                                   // v[16p - 1]
  if (p_begin > 0) {
    uint32_t a[1], b[1], c[1];
    tmem_ld1(t0 + 16 * p_begin - 1, a); tmem_ld1(t1 + 16 * p_begin - 1, b); tmem_ld1(t2 + 16 * p_begin - 1, c);
    tmem_wait1(a); tmem_wait1(b); tmem_wait1(c);
    vl = fmaf(k[2], F(c[0]), fmaf(k[1], F(b[0]), k[0] * F(a[0])));
  }
#pragma unroll 1
  for (int p = p_begin; p < p_end; ++p) {
    if (p < 3) {
      uint32_t a[16], b[16], c[16];
      tmem_ld16(t0 + 16 * p, a); tmem_ld16(t1 + 16 * p, b); tmem_ld16(t2 + 16 * p, c);
      tmem_wait16(a); tmem_wait16(b); tmem_wait16(c);
      f2 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        v[i] = fma2(kk[2], pku(c[2 * i], c[2 * i + 1]), fma2(kk[1], pku(b[2 * i], b[2 * i + 1]), mul2(kk[0], pku(a[2 * i], a[2 * i + 1]))));
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf(k[5], hi_of(v[i]), fmaf(k[4], lo_of(v[i]), k[3] * (i == 0 ? vl : hi_of(v[i - 1]))));
      vl = hi_of(v[7]);
      z_store8(zbuf, m_base + 8 * p, r, o);
    } else {
      uint32_t a[8], b[8], c[8];
      tmem_ld8(t0 + 48, a); tmem_ld8(t1 + 48, b); tmem_ld8(t2 + 48, c);
      tmem_wait8(a); tmem_wait8(b); tmem_wait8(c);
      f2 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        v[i] = fma2(kk[2], pku(c[2 * i], c[2 * i + 1]), fma2(kk[1], pku(b[2 * i], b[2 * i + 1]), mul2(kk[0], pku(a[2 * i], a[2 * i + 1]))));
      float o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = fmaf(k[5], hi_of(v[i]), fmaf(k[4], lo_of(v[i]), k[3] * (i == 0 ? vl : hi_of(v[i - 1]))));
      st_shared_v2(z_addr4(zbuf, m_base + 24, r), pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
    }
  }
}

// --------------------------------------------------------------------------------------------------------------------------------
// The kernel.  Units, chunks and groups (every role walks the same sequence):
//   stride 1: unit j of an image = output rows (2j, 2j+1); it reads chunks j and j+1, chunk c = input rows (2c-1, 2c); group = unit.
//   stride 2: unit j = output row j; it reads chunk j (rows 2j-1, 2j) and the first row of chunk j+1; group = 4 units.
// A CTA owns a contiguous range of the batch's groups; the first unit of a range / of an image starts a fresh chunk pair.
// --------------------------------------------------------------------------------------------------------------------------------
template <int MODE, int STRIDE, int C_IN, int C_OUT, bool BIAS>
__global__ void __launch_bounds__(kThreads, 1)
cpconv_rows_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out,
                   const __grid_constant__ CUtensorMap map_other, const __grid_constant__ CUtensorMap map_a,
                   const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_b, const Params p) {
  using C = Cfg<MODE, STRIDE, C_IN, C_OUT>;
  constexpr int G = C::G, XS = C::XS, OS = C::OS;
  static_assert(XS <= XS_MAX && OS <= OS_MAX, "ring depth");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + C::OFF_BAR;
  auto bar = [&](int i) { return bar_base + 8u * i; };
  const uint32_t tmem_slot = bar_base + 8u * B_COUNT;
  volatile uint32_t* tmem_slot_gen = (volatile uint32_t*)(smem_gen + C::OFF_BAR + 8 * B_COUNT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int g_begin = (int)(((long long)blockIdx.x * p.groups_total) / gridDim.x);
  const int g_end = (int)(((long long)(blockIdx.x + 1) * p.groups_total) / gridDim.x);
  const bool has_work = g_end > g_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_in); tma_prefetch_desc(&map_a);
    if (C::HAS_OUT) { tma_prefetch_desc(&map_out); tma_prefetch_desc(&map_b); }
    if (C::HAS_OTHER) tma_prefetch_desc(&map_other);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < XS; ++s) { mbar_init(bar(B_XFULL + s), 1); mbar_init(bar(B_XEMPTY + s), 1); }
    for (int s = 0; s < NS; ++s) { mbar_init(bar(B_SFULL + s), 1); mbar_init(bar(B_SEMPTY + s), kDwWarps); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(B_ZFULL + s), kDwWarps); mbar_init(bar(B_ZEMPTY + s), 1);
      mbar_init(bar(B_DFULL + s), 1); mbar_init(bar(B_DEMPTY + s), 4);
    }
    for (int s = 0; s < OS; ++s) { mbar_init(bar(B_OFULL + s), 1); mbar_init(bar(B_OEMPTY + s), 1); }
    mbar_init(bar(B_WFULL), 1); mbar_init(bar(B_CFULL), 1); mbar_init(bar(B_CEMPTY), kDwWarps); mbar_init(bar(B_FFULL), 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  // the shared walk over (group, unit): f(n, u, uu, gi, fresh, last_of_segment)
  auto walk = [&](auto&& f) {
    for (int g = g_begin; g < g_end; ++g) {
      const int n = g / p.gpi, gi = g - n * p.gpi;
#pragma unroll 1
      for (int uu = 0; uu < G; ++uu) {
        const bool fresh = uu == 0 && (g == g_begin || gi == 0);
        const bool last = uu == G - 1 && (g == g_end - 1 || gi == p.gpi - 1);
        f(n, gi * G + uu, uu, gi, fresh, last);
      }
    }
  };

  if (warp == 0) {
    // ===================================================================================== TMA producer
    if (has_work) {
      if (elect_one()) {
        const uint32_t wbytes = C::KB_IN * 16384 + (C::HAS_OUT ? 2 * C_OUT * 128 : 0) + (MODE == MODE_BWD_WEIGHT ? C::KB_OUT * 16384 : 0);
        mbar_expect_tx(bar(B_WFULL), wbytes);
        for (int kb = 0; kb < C::KB_IN; ++kb) tma_load_2d(smem_base + C::OFF_A + kb * 16384, &map_a, bar(B_WFULL), kb * 64, 0);
        if (MODE == MODE_BWD_WEIGHT)
          for (int kb = 0; kb < C::KB_OUT; ++kb) tma_load_2d(smem_base + C::OFF_A2 + kb * 16384, &map_a2, bar(B_WFULL), kb * 64, 0);
        if (C::HAS_OUT)
          for (int rb = 0; rb < 2; ++rb) tma_load_2d(smem_base + C::OFF_B + rb * (C_OUT * 128), &map_b, bar(B_WFULL), rb * 64, 0);
      }
      __syncwarp();
      uint32_t xq = 0, oq = 0;
      auto load_chunk = [&](int n, int c) {
        const uint32_t st = xq % XS;
        mbar_wait(bar(B_XEMPTY + st), ((xq / XS) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(bar(B_XFULL + st), C::X_STAGE);
#pragma unroll
          for (int kb = 0; kb < C::KB_IN; ++kb)
            tma_load_4d(smem_base + C::OFF_X + st * C::X_STAGE + kb * kChunkBytes, &map_in, bar(B_XFULL + st), kb * 64, 0, 2 * c - 1, n);
        }
        __syncwarp();
        ++xq;
      };
      walk([&](int n, int u, int uu, int gi, bool fresh, bool last) {
        if (fresh) load_chunk(n, u);
        load_chunk(n, u + 1);
        if (C::HAS_OTHER) {
          const uint32_t st = oq % OS;
          mbar_wait(bar(B_OEMPTY + st), ((oq / OS) & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx(bar(B_OFULL + st), C::O_STAGE);
#pragma unroll
            for (int kb = 0; kb < C::KB_OTHER; ++kb)
              tma_load_4d(smem_base + C::OFF_O + st * C::O_STAGE + kb * kChunkBytes, &map_other, bar(B_OFULL + st), kb * 64, 0, 2 * u, n);
          }
          __syncwarp();
          ++oq;
        }
      });
    }
  } else if (warp == 1) {
    // ===================================================================================== stage-1 issuer (z1 ring, dz2 centre)
    // Two issuing warps (this one and warp 3): a single thread walking every wait and every descriptor of a unit took ~3.5 k
    // cycles per unit and was the critical path of the first version; stage 1 now runs ahead as far as the TMEM ring allows.
    if (has_work) {
      constexpr uint32_t idesc1 = make_idesc(128, NP, 0, 0);            // D1[r][pos] : A K-major, chunk K-major
      mbar_wait(bar(B_WFULL), 0);
      const uint32_t a_lo = desc_lo(smem_base + C::OFF_A, 16), a2_lo = desc_lo(smem_base + C::OFF_A2, 16);
      const uint32_t x_lo = desc_lo(smem_base + C::OFF_X, 16), o_lo = desc_lo(smem_base + C::OFF_O, 16);
      uint32_t xq = 0, q = 0, oq = 0;
      auto mma1 = [&]() {
        const uint32_t st = xq % XS, slot = q % NS;
        mbar_wait(bar(B_XFULL + st), (xq / XS) & 1);
        mbar_wait(bar(B_SEMPTY + slot), ((q / NS) & 1) ^ 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < C::KB_IN * 4; ++kk)
            umma_lo(tmem_base + slot * NP, a_lo + (((kk >> 2) * 16384 + (kk & 3) * 32) >> 4),
                    x_lo + ((st * C::X_STAGE + (kk >> 2) * kChunkBytes + (kk & 3) * 32) >> 4), idesc1, kk > 0 ? 1u : 0u);
          umma_commit(bar(B_XEMPTY + st));
          umma_commit(bar(B_SFULL + slot));
        }
        __syncwarp();
        ++xq; ++q;
      };
      walk([&](int n, int u, int uu, int gi, bool fresh, bool last) {
        if (fresh) mma1();
        mma1();
        if (MODE == MODE_BWD_WEIGHT) {       // dz2 of the unit's own rows: D[r][pos] = F_out^T . dy chunk
          const uint32_t ost = oq % OS;
          mbar_wait(bar(B_OFULL + ost), (oq / OS) & 1);
          mbar_wait(bar(B_CEMPTY), (oq & 1) ^ 1);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < C::KB_OUT * 4; ++kk)
              umma_lo(tmem_base + C::COL_C, a2_lo + (((kk >> 2) * 16384 + (kk & 3) * 32) >> 4),
                      o_lo + ((ost * C::O_STAGE + (kk >> 2) * kChunkBytes + (kk & 3) * 32) >> 4), idesc1, kk > 0 ? 1u : 0u);
            umma_commit(bar(B_CFULL));
          }
          __syncwarp();
          ++oq;
        }
      });
    }
  } else if (warp == 3) {
    // ===================================================================================== stage-3 / weight-gradient issuer
    if (has_work) {
      constexpr uint32_t idesc2 = make_idesc(128, C_OUT, 1, 0);         // D2[pos][o] : z2 MN-major, B K-major
      constexpr uint32_t idesc3 = make_idesc(128, C::HAS_OTHER ? C::C_OTHER : 64, 0, 1);   // dF[r][c] : z K-major (rows = rank), chunk MN-major
      const int ksteps = p.KR / 16;
      mbar_wait(bar(B_WFULL), 0);
      const uint32_t zm_lo = desc_lo(smem_base + C::OFF_Z, 16384), zk_lo = desc_lo(smem_base + C::OFF_Z, 16);
      const uint32_t b_lo = desc_lo(smem_base + C::OFF_B, 16), o_lo = desc_lo(smem_base + C::OFF_O, kChunkBytes);
      bool first_w = true;
      for (uint32_t zq = 0; zq < (uint32_t)(g_end - g_begin); ++zq) {      // stage 3 and / or the weight-gradient product of group zq
        const uint32_t zb = zq & 1;
        mbar_wait(bar(B_ZFULL + zb), (zq >> 1) & 1);
        if (C::HAS_OUT) {
          const uint32_t dbuf = zq % C::ND;
          mbar_wait(bar(B_DEMPTY + dbuf), ((zq / C::ND) & 1) ^ 1);
          tc_fence_after();
          if (elect_one()) {
            for (int kk = 0; kk < ksteps; ++kk)
              umma_lo(tmem_base + C::COL_D + dbuf * C_OUT, zm_lo + ((zb * kZBytes + kk * 2048) >> 4),
                      b_lo + (((kk >> 2) * (C_OUT * 128) + (kk & 3) * 32) >> 4), idesc2, kk > 0 ? 1u : 0u);
            umma_commit(bar(B_DFULL + dbuf));
          }
          __syncwarp();
        }
        if (C::HAS_OTHER) {
          const uint32_t ost = zq % OS;      // stride 1: group == unit, the other-tensor stage of this unit
          mbar_wait(bar(B_OFULL + ost), (zq / OS) & 1);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int s = 0; s < NP / 16; ++s)
              umma_lo(tmem_base + C::COL_F, zk_lo + ((zb * kZBytes + (s >> 2) * 16384 + (s & 3) * 32) >> 4),
                      o_lo + ((ost * C::O_STAGE + s * 2048) >> 4), idesc3, (first_w && s == 0) ? 0u : 1u);
            umma_commit(bar(B_OEMPTY + ost));
          }
          __syncwarp();
          first_w = false;
        }
        if (elect_one()) umma_commit(bar(B_ZEMPTY + zb));
        __syncwarp();
      }
      if (elect_one()) umma_commit(bar(B_FFULL));
      __syncwarp();
    }
  } else if (warp >= 4 && warp < 4 + kDwWarps) {
    // ===================================================================================== depthwise stage
    const int wg = (warp - 4) >> 2, quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const bool active = quarter * 32 < p.KR;
    const uint32_t lane_bits = (uint32_t)(quarter * 32) << 16;
    float k[6];
    f2 kk[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { k[i] = __ldg(p.taps + i * 128 + r); kk[i] = pk(k[i], k[i]); }
    f2 dT[9];                                       // (sum over even columns, sum over odd columns)
#pragma unroll
    for (int i = 0; i < 9; ++i) dT[i] = 0ull;
    uint32_t q = 0, zq = 0, uq = 0, qb_prev = 0;
    walk([&](int n, int u, int uu, int gi, bool fresh, bool last) {
      uint32_t qa, qb;
      if (fresh) { qa = q; qb = q + 1; q += 2; } else { qa = qb_prev; qb = q; q += 1; }
      qb_prev = qb;
      mbar_wait(bar(B_SFULL + qb % NS), (qb / NS) & 1);
      const uint32_t zb = zq & 1;
      if (uu == 0) mbar_wait(bar(B_ZEMPTY + zb), ((zq >> 1) & 1) ^ 1);
      if (MODE == MODE_BWD_WEIGHT) mbar_wait(bar(B_CFULL), uq & 1);
      tc_fence_after();
      if (active) {
        const uint32_t ta = tmem_base + lane_bits + (qa % NS) * NP, tb = tmem_base + lane_bits + (qb % NS) * NP;
        const uint32_t zbuf = smem_base + C::OFF_Z + zb * kZBytes;
        if (STRIDE == 1) {
          // rows 2u-1, 2u | 2u+1, 2u+2 ; this warp group's output row is 2u + wg
          const uint32_t t0 = wg == 0 ? ta : ta + 56, t1 = wg == 0 ? ta + 56 : tb, t2 = wg == 0 ? tb : tb + 56;
          const uint32_t tc = tmem_base + lane_bits + C::COL_C + wg * 56;
          dw_row_s1<MODE == MODE_BWD_WEIGHT>(t0, t1, t2, tc, kk, k, zbuf, wg * 56, r, dT);
        } else {
          // rows 2u-1, 2u (chunk a) and 2u+1 (first row of chunk b); warp group 0 takes outputs 0..15, group 1 outputs 16..27
          dw_row_s2(ta, ta + 56, tb, kk, k, zbuf, uu * 28, r, wg * 2, wg * 2 + 2);
        }
      }
      tc_fence_before();
      if (uu == G - 1) fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(B_SEMPTY + qa % NS));
        if (last) mbar_arrive(bar(B_SEMPTY + qb % NS));
        if (MODE == MODE_BWD_WEIGHT) mbar_arrive(bar(B_CEMPTY));
        if (uu == G - 1) mbar_arrive(bar(B_ZFULL + zb));
      }
      if (uu == G - 1) ++zq;
      ++uq;
    });
    if (MODE == MODE_BWD_WEIGHT && active && r < p.R) {
#pragma unroll
      for (int i = 0; i < 9; ++i) red_add_f32(p.dT + i * 128 + r, lo_of(dT[i]) + hi_of(dT[i]));
    }
  } else if (warp >= 12) {
    // ===================================================================================== epilogue
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;                      // position of the tile = TMEM lane
    const uint32_t lane_bits = (uint32_t)(quarter * 32) << 16;
    if (C::HAS_OUT) {
      uint32_t gq = 0;
      for (int g = g_begin; g < g_end; ++g, ++gq) {
        const int n = g / p.gpi, gi = g - n * p.gpi;
        const uint32_t dbuf = gq % C::ND, sb = gq & 1;
        mbar_wait(bar(B_DFULL + dbuf), (gq / C::ND) & 1);
        tc_fence_after();
        if (warp == 12 && lane == 0) bulk_wait_read1();       // the store that last used staging buffer sb has read it
        named_bar_sync(1, kEpiThreads);
        const uint32_t ybuf = smem_base + C::OFF_Y + sb * C::Y_STAGE;
#pragma unroll
        for (int cb = 0; cb < C_OUT / 32; ++cb) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + lane_bits + C::COL_D + dbuf * C_OUT + cb * 32, v);
          tmem_ld_wait();
          if (cb == C_OUT / 32 - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_DEMPTY + dbuf));
          }
          if (m < NP) {
            const uint32_t row = ybuf + (cb >> 1) * kChunkBytes + m * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = F(v[8 * j + e]);
              if (BIAS) {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] += __ldg(p.bias + cb * 32 + 8 * j + e);
              }
              st_shared_v4(row + ((((uint32_t)((cb & 1) * 4 + j)) ^ (uint32_t)(m & 7)) << 4), pack_bf16x2(f[0], f[1]),
                           pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
            }
          }
        }
        fence_proxy_async();
        named_bar_sync(1, kEpiThreads);
        if (warp == 12 && lane == 0) {
          const int h0 = STRIDE == 1 ? 2 * gi : 4 * gi;
          for (int cb = 0; cb < C::KB_OUT; ++cb) tma_store_4d(&map_out, ybuf + cb * kChunkBytes, cb * 64, 0, h0, n);
          bulk_commit();
        }
      }
      if (warp == 12 && lane == 0) bulk_wait_read0();
    }
    if (C::HAS_OTHER && has_work) {
      // flush the weight-gradient accumulator: dF[c][r] += D[r lane][c column]
      mbar_wait(bar(B_FFULL), 0);
      tc_fence_after();
      if (quarter * 32 < p.KR) {
#pragma unroll 1
        for (int cb = 0; cb < C::C_OTHER / 32; ++cb) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + lane_bits + C::COL_F + cb * 32, v);
          tmem_ld_wait();
          if (m < p.R) {
#pragma unroll
            for (int e = 0; e < 32; ++e) red_add_f32(p.dF + (cb * 32 + e) * 128 + m, F(v[e]));
          }
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// --------------------------------------------------------------------------------------------------------------------------------
// Operand packing: one launch builds every matrix the three kernels read, zero-padded to the tcgen05 tile shapes.
//   ws (bf16): A_in [128][C_in] = F_in^T | B_in [C_in][128] = F_in | A_out [128][C_out] = F_out^T | B_out [C_out][128] = F_out
//   then fp32 taps_fwd [6][128] (kh0 kh1 kh2, lambda kw0, lambda kw1, lambda kw2) and taps_bwd [6][128] (both reversed),
//   then (bf16) the tap-folded stage-1 operands of the second-generation kernels (cpconv_rows2.cu):
//   A3_fwd [3][128][C_in] = kh[i, r] F_in[c, r]   and   A3_bwd [3][128][C_out] = lambda_r kw[2 - i, r] F_out[o, r].
// --------------------------------------------------------------------------------------------------------------------------------
__global__ void pack_kernel(const __nv_bfloat16* __restrict__ f_in, const __nv_bfloat16* __restrict__ f_out,
                            const __nv_bfloat16* __restrict__ kh, const __nv_bfloat16* __restrict__ kw,
                            const __nv_bfloat16* __restrict__ lam, int c_in, int c_out, int R, __nv_bfloat16* __restrict__ ws,
                            float* __restrict__ taps, __nv_bfloat16* __restrict__ a3) {
  const int n_in = 128 * c_in, n_out = 128 * c_out;
  const int total = 2 * n_in + 2 * n_out + 2 * 6 * 128;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * (n_in + n_out); i += gridDim.x * blockDim.x) {
    float v = 0.f;
    if (i < 3 * n_in) {
      const int t = i / n_in, j = i - t * n_in, r = j / c_in, c = j - r * c_in;
      if (r < R) v = __bfloat162float(kh[t * R + r]) * __bfloat162float(f_in[c * R + r]);
    } else {
      const int i2 = i - 3 * n_in, t = i2 / n_out, j = i2 - t * n_out, r = j / c_out, c = j - r * c_out;
      if (r < R) v = __bfloat162float(lam[r]) * __bfloat162float(kw[(2 - t) * R + r]) * __bfloat162float(f_out[c * R + r]);
    }
    a3[i] = __float2bfloat16_rn(v);
  }
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int j = i;
    if (j < n_in) { const int r = j / c_in, c = j - r * c_in; ws[i] = r < R ? f_in[c * R + r] : zero; continue; }
    j -= n_in;
    if (j < n_in) { const int c = j >> 7, r = j & 127; ws[i] = r < R ? f_in[c * R + r] : zero; continue; }
    j -= n_in;
    if (j < n_out) { const int r = j / c_out, c = j - r * c_out; ws[i] = r < R ? f_out[c * R + r] : zero; continue; }
    j -= n_out;
    if (j < n_out) { const int c = j >> 7, r = j & 127; ws[i] = r < R ? f_out[c * R + r] : zero; continue; }
    j -= n_out;
    const int dir = j / 768, t = (j % 768) >> 7, r = j & 127;
    float v = 0.f;
    if (r < R) {
      const int tt = dir == 0 ? t : (t < 3 ? 2 - t : 3 + (5 - t));
      v = tt < 3 ? __bfloat162float(kh[tt * R + r]) : __bfloat162float(kw[(tt - 3) * R + r]) * __bfloat162float(lam[r]);
    }
    taps[j] = v;
  }
}

// Gradient finalisation: fp32 accumulators -> the five parameter gradients (bf16).
//   acc: dF_in [C_in][128] | dF_out [C_out][128] | dT [9][128]   (dT[3 i + j][r] = sum dz2 * z1(+tap i, j))
__global__ void finalize_kernel(const float* __restrict__ acc, const __nv_bfloat16* __restrict__ kh, const __nv_bfloat16* __restrict__ kw,
                                const __nv_bfloat16* __restrict__ lam, int c_in, int c_out, int R, int folded, __nv_bfloat16* __restrict__ g_fout,
                                __nv_bfloat16* __restrict__ g_fin, __nv_bfloat16* __restrict__ g_kh, __nv_bfloat16* __restrict__ g_kw,
                                __nv_bfloat16* __restrict__ g_lam) {
  const int n_in = c_in * R, n_out = c_out * R;
  const int total = n_in + n_out + R;
  const float* dT = acc + (c_in + c_out) * 128;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (i < n_in) { const int c = i / R, r = i - c * R; g_fin[i] = __float2bfloat16_rn(acc[c * 128 + r]); continue; }
    int j = i - n_in;
    if (j < n_out) { const int c = j / R, r = j - c * R; g_fout[j] = __float2bfloat16_rn(acc[(c_in + c) * 128 + r]); continue; }
    const int r = j - n_out;
    float h[3], w[3], t[9];
    const float l = __bfloat162float(lam[r]);
#pragma unroll
    for (int a = 0; a < 3; ++a) { h[a] = __bfloat162float(kh[a * R + r]); w[a] = __bfloat162float(kw[a * R + r]); }
#pragma unroll
    for (int a = 0; a < 9; ++a) t[a] = dT[a * 128 + r];
    float gl = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      // folded kernels: t[0..2] = d (lambda kw[a]) from BWD_WEIGHT, t[3..5] = d kh[2 - a] from the column-order BWD_DATA
      const float gh = folded ? t[3 + (2 - a)] : l * (w[0] * t[3 * a] + w[1] * t[3 * a + 1] + w[2] * t[3 * a + 2]);      // d kh[a]
      const float gwl = folded ? t[a] : h[0] * t[a] + h[1] * t[3 + a] + h[2] * t[6 + a];                                  // d (lambda kw[a])
      g_kh[a * R + r] = __float2bfloat16_rn(gh);
      g_kw[a * R + r] = __float2bfloat16_rn(l * gwl);
      gl += w[a] * gwl;
    }
    g_lam[r] = __float2bfloat16_rn(gl);
  }
}

// --------------------------------------------------------------------------------------------------------------------------------
// Host side
// --------------------------------------------------------------------------------------------------------------------------------
static int make_map4d(CUtensorMap* map, const void* ptr, int64_t c, int64_t w, int64_t h, int64_t n, int box_w, int box_h) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)"); return TLT_ERR_CUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)c * w * 2, (cuuint64_t)c * w * h * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = CUDA_SUCCESS;
  for (int attempt = 0; attempt < 2; ++attempt) {
    r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_ERROR_INVALID_CONTEXT) break;       // see make_map
    cudaFree(0);
  }
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(4d) failed (%d) for dims {%lld,%lld,%lld,%lld} box {64,%d,%d,1}", (int)r, (long long)c,
              (long long)w, (long long)h, (long long)n, box_w, box_h);
    return TLT_ERR_CUDA;
  }
  return TLT_OK;
}

struct Ws {        // byte offsets into the packed-operand workspace
  size_t a_in, b_in, a_out, b_out, taps_fwd, taps_bwd, a3_fwd, a3_bwd, total;
};
static Ws ws_layout(int c_in, int c_out) {
  Ws w;
  w.a_in = 0;
  w.b_in = w.a_in + (size_t)128 * c_in * 2;
  w.a_out = w.b_in + (size_t)128 * c_in * 2;
  w.b_out = w.a_out + (size_t)128 * c_out * 2;
  w.taps_fwd = w.b_out + (size_t)128 * c_out * 2;
  w.taps_bwd = w.taps_fwd + 6 * 128 * 4;
  w.a3_fwd = w.taps_bwd + 6 * 128 * 4;
  w.a3_bwd = w.a3_fwd + (size_t)3 * 128 * c_in * 2;
  w.total = w.a3_bwd + (size_t)3 * 128 * c_out * 2;
  return w;
}

static bool shape_ok(const tlt_cpconv_rows_desc* d) {
  if (d->width != 56 || d->height % 2 != 0 || d->height < 2 || d->rank < 1 || d->rank > 128 || d->batch < 1) return false;
  if (d->stride == 1) return d->in_channels == 64 && d->out_channels == 64;
  if (d->stride == 2) return d->in_channels == 64 && d->out_channels == 128 && d->height % 8 == 0;
  return false;
}

template <int MODE, int STRIDE, int C_IN, int C_OUT>
static int launch(const tlt_cpconv_rows_desc* d, const void* in, void* out, const void* other, const void* a, const void* a2,
                  const void* b, const float* taps, const float* bias, float* dF, float* dT, cudaStream_t st) {
  using C = Cfg<MODE, STRIDE, C_IN, C_OUT>;
  static_assert(C::SMEM <= 227 * 1024, "shared memory budget");
  static_assert(C::COL_D + (C::HAS_OUT ? C::ND * C_OUT : 0) <= 512, "TMEM budget");
  static_assert(!C::HAS_OTHER || C::COL_F + C::C_OTHER <= 512, "TMEM budget (weight gradient)");
  const int H = (int)d->height, W = (int)d->width, B = (int)d->batch;
  const int Ho = H / STRIDE, Wo = W / STRIDE;
  CUtensorMap m_in, m_out, m_other, m_a, m_a2, m_b;
  int rc;
  if ((rc = make_map4d(&m_in, in, C_IN, W, H, B, W, 2)) != TLT_OK) return rc;
  m_out = m_in; m_other = m_in; m_a2 = m_in; m_b = m_in;
  if (C::HAS_OUT && (rc = make_map4d(&m_out, out, C_OUT, Wo, Ho, B, Wo, NP / Wo)) != TLT_OK) return rc;
  if (C::HAS_OTHER && (rc = make_map4d(&m_other, other, C::C_OTHER, W, H, B, W, 2)) != TLT_OK) return rc;
  if ((rc = make_map(&m_a, a, C_IN, 128, C_IN, 64, 128)) != TLT_OK) return rc;
  if (MODE == MODE_BWD_WEIGHT && (rc = make_map(&m_a2, a2, C_OUT, 128, C_OUT, 64, 128)) != TLT_OK) return rc;
  if (C::HAS_OUT && (rc = make_map(&m_b, b, 128, C_OUT, 128, 64, C_OUT)) != TLT_OK) return rc;
  Params p;
  p.batch = B; p.H = H; p.W = W; p.Ho = Ho; p.Wo = Wo;
  p.R = (int)d->rank; p.KR = (p.R + 15) / 16 * 16;
  p.gpi = Ho * Wo / NP;
  p.groups_total = B * p.gpi;
  p.taps = taps; p.bias = bias; p.dF = dF; p.dT = dT;
  int grid = num_sms();
  if (grid > p.groups_total) grid = p.groups_total;
  if (MODE == MODE_FWD && bias != nullptr) {
    auto kern = cpconv_rows_kernel<MODE, STRIDE, C_IN, C_OUT, MODE == MODE_FWD>;
    TLT_ENSURE_SMEM(kern, C::SMEM);
    kern<<<grid, kThreads, C::SMEM, st>>>(m_in, m_out, m_other, m_a, m_a2, m_b, p);
  } else {
    auto kern = cpconv_rows_kernel<MODE, STRIDE, C_IN, C_OUT, false>;
    TLT_ENSURE_SMEM(kern, C::SMEM);
    kern<<<grid, kThreads, C::SMEM, st>>>(m_in, m_out, m_other, m_a, m_a2, m_b, p);
  }
  return check_launch("cpconv_rows_kernel");
}

}  // namespace cpr
}  // namespace tlt

namespace tlt {
namespace cpr2 {
int fwd_64(int batch, int rank, const void* x, void* y, const void* a3, const void* b, const float* taps, const float* bias, cudaStream_t st);
int bwd_data_64(int batch, int rank, const void* dy, void* dx, const void* x, const void* a3, const void* a2, const void* b,
                const float* taps, float* dF, float* dG, cudaStream_t st);
int bwd_weight_64(int batch, int rank, const void* x, const void* dy, const void* a3, const void* a2, const float* taps, float* dF,
                  float* dG, cudaStream_t st);
}  // namespace cpr2
}  // namespace tlt

using namespace tlt;
using namespace tlt::cpr;

// the second-generation kernels (first separable pass folded into the stage-1 MMA): square 56 x 56 maps, 64 -> 64, rank <= 80
static bool folded_ok(const tlt_cpconv_rows_desc* d) {
  return d->stride == 1 && d->height == 56 && d->width == 56 && d->in_channels == 64 && d->out_channels == 64 && d->rank <= 80;
}

extern "C" int tlt_cpconv_rows_supported(const tlt_cpconv_rows_desc* d) { return d && shape_ok(d) ? 1 : 0; }

extern "C" size_t tlt_cpconv_rows_workspace_size(const tlt_cpconv_rows_desc* d) {
  return d ? ws_layout((int)d->in_channels, (int)d->out_channels).total : 0;
}
extern "C" size_t tlt_cpconv_rows_grad_workspace_size(const tlt_cpconv_rows_desc* d) {
  return d ? (size_t)(d->in_channels + d->out_channels + 9) * 128 * 4 : 0;
}

extern "C" int tlt_cpconv_rows_pack(const tlt_cpconv_rows_desc* d, const void* f_in, const void* f_out, const void* kh, const void* kw,
                                    const void* lam, void* ws, void* stream) {
  TLT_REQUIRE(d && shape_ok(d), "tlt_cpconv_rows_pack: unsupported shape");
  const Ws w = ws_layout((int)d->in_channels, (int)d->out_channels);
  pack_kernel<<<64, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)f_in, (const __nv_bfloat16*)f_out, (const __nv_bfloat16*)kh,
                                                    (const __nv_bfloat16*)kw, (const __nv_bfloat16*)lam, (int)d->in_channels,
                                                    (int)d->out_channels, (int)d->rank, (__nv_bfloat16*)ws,
                                                    (float*)((uint8_t*)ws + w.taps_fwd), (__nv_bfloat16*)((uint8_t*)ws + w.a3_fwd));
  return check_launch("cpconv_rows pack_kernel");
}

extern "C" int tlt_cpconv_rows_fwd(const tlt_cpconv_rows_desc* d, const void* x, const void* ws, const float* bias, void* y, void* stream) {
  TLT_REQUIRE(d && shape_ok(d), "tlt_cpconv_rows_fwd: unsupported shape (W = 56; 64 -> 64 stride 1 or 64 -> 128 stride 2; rank <= 128)");
  const Ws w = ws_layout((int)d->in_channels, (int)d->out_channels);
  const uint8_t* base = (const uint8_t*)ws;
  cudaStream_t st = (cudaStream_t)stream;
  if (folded_ok(d))
    return cpr2::fwd_64((int)d->batch, (int)d->rank, x, y, base + w.a3_fwd, base + w.b_out, (const float*)(base + w.taps_fwd) + 3 * 128, bias, st);
  if (d->stride == 1)
    return launch<MODE_FWD, 1, 64, 64>(d, x, y, nullptr, base + w.a_in, nullptr, base + w.b_out, (const float*)(base + w.taps_fwd), bias,
                                       nullptr, nullptr, st);
  return launch<MODE_FWD, 2, 64, 128>(d, x, y, nullptr, base + w.a_in, nullptr, base + w.b_out, (const float*)(base + w.taps_fwd), bias,
                                      nullptr, nullptr, st);
}

extern "C" int tlt_cpconv_rows_bwd(const tlt_cpconv_rows_desc* d, const void* x, const void* dy, const void* ws, void* dx, float* acc,
                                   void* stream) {
  TLT_REQUIRE(d && shape_ok(d) && d->stride == 1, "tlt_cpconv_rows_bwd: unsupported shape (stride 1, 64 -> 64, W = 56, rank <= 128)");
  const Ws w = ws_layout((int)d->in_channels, (int)d->out_channels);
  const uint8_t* base = (const uint8_t*)ws;
  cudaStream_t st = (cudaStream_t)stream;
  TLT_CHECK_CUDA(cudaMemsetAsync(acc, 0, tlt_cpconv_rows_grad_workspace_size(d), st));
  float* dF_in = acc;
  float* dF_out = acc + (size_t)d->in_channels * 128;
  float* dT = dF_out + (size_t)d->out_channels * 128;
  if (folded_ok(d)) {
    // column-order data gradient: "in" = dy, fold lambda kw (flipped) into F_out^T, pass = kh (flipped); centre z1 = F_in^T x  -> dx, dF_in, d kh
    int rc = cpr2::bwd_data_64((int)d->batch, (int)d->rank, dy, dx, x, base + w.a3_bwd, base + w.a_in, base + w.b_in,
                               (const float*)(base + w.taps_bwd), dF_in, dT + 3 * 128, st);
    if (rc != TLT_OK) return rc;
    // row-order weight gradient: "in" = x, fold kh into F_in^T, pass = lambda kw; centre dz2 = F_out^T dy  -> dF_out, d (lambda kw)
    return cpr2::bwd_weight_64((int)d->batch, (int)d->rank, x, dy, base + w.a3_fwd, base + w.a_out,
                               (const float*)(base + w.taps_fwd) + 3 * 128, dF_out, dT, st);
  }
  // data gradient (+ dF_in): "in" = dy, A = F_out^T, B = F_in, other = x
  int rc = launch<MODE_BWD_DATA, 1, 64, 64>(d, dy, dx, x, base + w.a_out, nullptr, base + w.b_in, (const float*)(base + w.taps_bwd),
                                            nullptr, dF_in, nullptr, st);
  if (rc != TLT_OK) return rc;
  // weight gradients of F_out and of the taps: "in" = x, A = F_in^T, A2 = F_out^T, other = dy
  return launch<MODE_BWD_WEIGHT, 1, 64, 64>(d, x, nullptr, dy, base + w.a_in, base + w.a_out, nullptr, (const float*)(base + w.taps_fwd),
                                            nullptr, dF_out, dT, st);
}

extern "C" int tlt_cpconv_rows_finalize(const tlt_cpconv_rows_desc* d, const float* acc, const void* kh, const void* kw, const void* lam,
                                        void* g_fout, void* g_fin, void* g_kh, void* g_kw, void* g_lam, void* stream) {
  TLT_REQUIRE(d && shape_ok(d), "tlt_cpconv_rows_finalize: unsupported shape");
  finalize_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(acc, (const __nv_bfloat16*)kh, (const __nv_bfloat16*)kw, (const __nv_bfloat16*)lam,
                                                        (int)d->in_channels, (int)d->out_channels, (int)d->rank, folded_ok(d) ? 1 : 0,
                                                        (__nv_bfloat16*)g_fout,
                                                        (__nv_bfloat16*)g_fin, (__nv_bfloat16*)g_kh, (__nv_bfloat16*)g_kw,
                                                        (__nv_bfloat16*)g_lam);
  return check_launch("cpconv_rows finalize_kernel");
}
