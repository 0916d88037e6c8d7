// Fused CP convolution on channels-last rows, second generation: the FIRST separable depthwise pass is folded into the stage-1
// tensor-core product, so the CUDA cores only run the second pass.
//
//     v[r, h, w] = sum_i kh[i, r] z1[r, h+i-1, w] = sum_i sum_c (kh[i, r] F_in[c, r]) x[h+i-1, w, c]
//
// is ONE accumulation of three MMAs whose B operands are the same shared-memory box of four image rows shifted by one row
// (56 positions = 7 swizzle atoms of 8 rows, so every shifted start is atom-aligned) and whose A operands are the three
// tap-scaled copies of F_in^T the packing launch writes.  What that buys over the first kernel (cpconv_rows.cu, profiles/r02_*):
//   * every output row needs ONE TMEM row instead of three (tcgen05.ld waits were the top stall), no vertical halo in TMEM, units
//     are independent (a unit = two output rows of one image = one 112-position stage-3 tile);
//   * the depthwise warps run 3 taps instead of 3 + 3 (packed fp32 pairs: 2 fma.rn.f32x2 + 2 scalar FMAs per output pair);
//   * the weight gradients split symmetrically over TWO kernels of the same shape, each owning one tap gradient:
//       BWD_WEIGHT  (row order)     in = x,  fold kh into F_in^T,  pass = kw :  z2 -> dF_out += z2^T dy,  g[j] = sum dz2 * v(+j-1) = d kw'
//       BWD_DATA    (column order)  in = dy, fold kw' (flipped) into F_out^T, pass = kh (flipped) : dz1 -> dx, dF_in += dz1^T x,
//                                   g[j] = sum z1 * gw(+j-1) = d kh (flipped)
//     "column order" is the same kernel on tensor maps whose W and H strides are swapped (the maps are square, 56 x 56): a shift
//     of one position is then a shift along H.  dz2 / z1 of the unit's own rows come from a fourth stage-1 product (centre).
// Modes: FWD (y), BWD_DATA (dx + dF_in + one tap gradient), BWD_WEIGHT (dF_out + the other tap gradient).
// Reference: cp_conv (tltorch/functional/convolution.py:231-283) and its autograd.
#include "cprows_common.cuh"

namespace tlt {
namespace cpr2 {
using namespace tlt::cpr;

constexpr int kThreads = 512;          // warps: 0 TMA | 1 stage-1 MMA | 2 TMEM alloc | 3 stage-3 MMA | 4-11 depthwise (2 x 4) | 12-15 epilogue
constexpr int kBoxBytes = 4 * 56 * 128;  // four image rows of one 64-channel block

enum { MODE_FWD = 0, MODE_BWD_DATA = 1, MODE_BWD_WEIGHT = 2 };

struct Params {
  int units_total;           // batch * 28
  int R, KR;                 // rank, rank rounded up to a multiple of 16
  const float* taps;         // [3][128] fp32: the taps of the pass the CUDA cores run
  const float* bias;         // fp32 [C_OUT] or nullptr (FWD)
  float* dF;                 // BWD_*: fp32 [C_OTHER][128]
  float* dG;                 // BWD_*: fp32 [3][128] tap gradient
};

// KRT: rank rows kept in shared memory per operand block (a multiple of 16 >= KR; the MMAs read 128 rows, rows >= KRT are
// whatever follows in shared memory and only feed TMEM lanes nobody reads)
template <int MODE, int C_IN, int C_OUT, int KRT> struct Cfg {
  static constexpr int KB_IN = C_IN / 64, KB_OUT = C_OUT / 64;
  static constexpr bool HAS_OUT = MODE != MODE_BWD_WEIGHT;
  static constexpr bool HAS_OTHER = MODE != MODE_FWD;
  static constexpr int C_OTHER = C_OUT;        // BWD_DATA: "in" = dy, out = dx, other = x (C_OUT channels); BWD_WEIGHT: other = dy
  static constexpr int KB_OTHER = C_OTHER / 64;
  static constexpr int A_BLOCK = KRT * 128;                          // one 64-channel block of one A matrix
  static constexpr int Z_BLOCK = KRT * 128;                          // one 64-position block of a z tile
  static constexpr int Z_BYTES = 2 * Z_BLOCK;
  static constexpr int NS = 2;                                       // TMEM ring slots (units are independent)
  static constexpr int NC = MODE == MODE_BWD_WEIGHT ? 2 : 1;         // centre buffers
  static constexpr int ND = MODE == MODE_FWD ? 2 : 1;
  static constexpr int XS = MODE == MODE_FWD ? 3 : (MODE == MODE_BWD_WEIGHT ? 3 : 2);
  static constexpr int OS = 3;
  static constexpr int YS = MODE == MODE_FWD ? 2 : 1;
  // TMEM columns
  static constexpr int COL_C = NS * NP;
  static constexpr int COL_D = COL_C + (HAS_OTHER ? NC * NP : 0);
  static constexpr int COL_F = COL_D + (HAS_OUT ? ND * C_OUT : 0);
  static constexpr int COL_END = COL_F + (HAS_OTHER ? C_OTHER : 0);
  // shared memory
  static constexpr int OFF_A = 0;                                                        // 3 tap-scaled A matrices
  static constexpr int OFF_A2 = OFF_A + 3 * KB_IN * A_BLOCK;                             // centre factor
  static constexpr int OFF_B = OFF_A2 + (HAS_OTHER ? KB_OTHER * A_BLOCK : 0);            // [2 rank blocks][C_OUT rows x 128 B]
  static constexpr int OFF_X = (OFF_B + (HAS_OUT ? 2 * C_OUT * 128 : 0) + 1023) / 1024 * 1024;
  static constexpr int X_STAGE = KB_IN * kBoxBytes;
  static constexpr int OFF_Z = OFF_X + XS * X_STAGE;
  static constexpr int OFF_Y = (OFF_Z + 2 * Z_BYTES + 1023) / 1024 * 1024;
  static constexpr int Y_STAGE = KB_OUT * kChunkBytes;
  static constexpr int OFF_O = OFF_Y + (HAS_OUT ? YS * Y_STAGE : 0);
  static constexpr int O_STAGE = KB_OTHER * kChunkBytes;
  static constexpr int OFF_BAR = OFF_O + (HAS_OTHER ? OS * O_STAGE : 0);
  static constexpr int SMEM = OFF_BAR + 512 + 1024;
};

enum { B_XFULL = 0, B_XEMPTY = 4, B_SFULL = 8, B_SEMPTY = 10, B_ZFULL = 12, B_ZEMPTY = 14, B_DFULL = 16, B_DEMPTY = 18,
       B_OFULL = 20, B_OEMPTY = 24, B_CFULL = 28, B_CEMPTY = 30, B_WFULL = 32, B_FFULL = 33, B_COUNT = 34 };

// z tile address of (position m, rank row r) with Z_BLOCK bytes per 64-position block
template <int Z_BLOCK>
__device__ __forceinline__ uint32_t z2_addr(uint32_t zbuf, int m, int r) {
  return zbuf + (uint32_t)((m >> 6) * Z_BLOCK + r * 128) + ((uint32_t)((((m & 63) >> 3) ^ (r & 7))) << 4);
}

// --------------------------------------------------------------------------------------------------------------------------------
// The CUDA-core pass over one row of 56 positions: z[w] = k0 y[w-1] + k1 y[w] + k2 y[w+1]  (zero outside the row).
// ty: TMEM address of the row in the ring; column pieces of 8, next piece's loads in flight, outputs lag one column (see
// cpconv_rows.cu).  GRAD: tc = the centre row (dz2 or z1 of the same positions); g[j] += sum_w c[w] y[w + j - 1], accumulated as
// (even columns, odd columns) pairs: the y operand is always an even-aligned pair, the centre comes even-aligned (j = 1) and
// odd-aligned (a second tcgen05.ld one column further; j = 0 and j = 2).
// --------------------------------------------------------------------------------------------------------------------------------
template <bool GRAD, int Z_BLOCK>
__device__ __forceinline__ void pass_row(uint32_t ty, uint32_t tc, const f2 (&kk)[3], float k1, uint32_t zbuf, int m_base, int r,
                                         bool store, f2 (&g)[3]) {
  uint32_t a[8], d[8], e[8];
  float pend[7];
  f2 yp = 0ull;                                    // (y[8p-2], y[8p-1])
  f2 dprev = 0ull;                                 // GRAD: (c[8p-1], c[8p])
  tmem_ld8(ty, a);
  if (GRAD) { tmem_ld8(tc, d); tmem_ld8(tc + 1, e); }
#pragma unroll
  for (int p = 0; p < 7; ++p) {
    tmem_wait8(a);
    if (GRAD) { tmem_wait8(d); tmem_wait8(e); }
    f2 y[4], de[4], dod[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      y[i] = pku(a[2 * i], a[2 * i + 1]);
      if (GRAD) {
        de[i] = pku(d[2 * i], d[2 * i + 1]);
        dod[i] = pku(e[2 * i], (p == 6 && i == 3) ? 0u : e[2 * i + 1]);      // column 56 does not exist
      }
    }
    if (GRAD && p == 0) dprev = pku(0u, d[0]);
    if (p + 1 < 7) {
      tmem_ld8(ty + 8 * (p + 1), a);
      if (GRAD) { tmem_ld8(tc + 8 * (p + 1), d); tmem_ld8(tc + 8 * (p + 1) + 1, e); }
    }
    if (GRAD) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const f2 dl = i == 0 ? dprev : dod[i - 1];     // centres one column left of the y pair (tap j = 2)
        g[0] = fma2(dod[i], y[i], g[0]); g[1] = fma2(de[i], y[i], g[1]); g[2] = fma2(dl, y[i], g[2]);
      }
      dprev = dod[3];
    }
    // outputs 8p-1 .. 8p+6 as pairs (o[2i-1], o[2i]) = k0 (y[2i-2], y[2i-1]) + k2 (y[2i], y[2i+1]) + k1 (y[2i-1], y[2i])
    float o[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const f2 yl = i == 0 ? yp : y[i - 1];
      const f2 t = fma2(kk[2], y[i], mul2(kk[0], yl));
      o[2 * i] = fmaf(k1, hi_of(yl), lo_of(t));
      o[2 * i + 1] = fmaf(k1, lo_of(y[i]), hi_of(t));
    }
    if (p > 0 && store)
      st_shared_v4(z2_addr<Z_BLOCK>(zbuf, m_base + 8 * (p - 1), r), pack_bf16x2(pend[0], pend[1]), pack_bf16x2(pend[2], pend[3]),
                   pack_bf16x2(pend[4], pend[5]), pack_bf16x2(pend[6], o[0]));
#pragma unroll
    for (int i = 0; i < 7; ++i) pend[i] = o[i + 1];
    yp = y[3];
  }
  const float o_last = fmaf(k1, hi_of(yp), lo_of(kk[0]) * lo_of(yp));          // column 55
  if (store)     // lanes >= KR hold whatever followed the operand block in shared memory: their z rows do not exist
    st_shared_v4(z2_addr<Z_BLOCK>(zbuf, m_base + 48, r), pack_bf16x2(pend[0], pend[1]), pack_bf16x2(pend[2], pend[3]),
                 pack_bf16x2(pend[4], pend[5]), pack_bf16x2(pend[6], o_last));
}

// --------------------------------------------------------------------------------------------------------------------------------
// Unit u of the batch: image n = u / 28, rows (2j, 2j+1), j = u % 28 (columns when the tensor maps are transposed).  Boxes: the
// "in" tensor rows 2j-1 .. 2j+2 (zero-filled outside the image), the other tensor / the output rows 2j, 2j+1.
// --------------------------------------------------------------------------------------------------------------------------------
template <int MODE, int C_IN, int C_OUT, int KRT, bool BIAS>
__global__ void __launch_bounds__(kThreads, 1)
cpconv_rows2_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out,
                    const __grid_constant__ CUtensorMap map_other, const __grid_constant__ CUtensorMap map_a,
                    const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_b, const Params p) {
  using C = Cfg<MODE, C_IN, C_OUT, KRT>;
  constexpr int XS = C::XS, OS = C::OS, NS = C::NS, NC = C::NC, ND = C::ND, YS = C::YS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + C::OFF_BAR;
  auto bar = [&](int i) { return bar_base + 8u * i; };
  const uint32_t tmem_slot = bar_base + 8u * B_COUNT;
  volatile uint32_t* tmem_slot_gen = (volatile uint32_t*)(smem_gen + C::OFF_BAR + 8 * B_COUNT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int u_begin = (int)(((long long)blockIdx.x * p.units_total) / gridDim.x);
  const int u_end = (int)(((long long)(blockIdx.x + 1) * p.units_total) / gridDim.x);
  const uint32_t n_units = (uint32_t)(u_end - u_begin);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_in); tma_prefetch_desc(&map_a);
    if (C::HAS_OUT) { tma_prefetch_desc(&map_out); tma_prefetch_desc(&map_b); }
    if (C::HAS_OTHER) { tma_prefetch_desc(&map_other); tma_prefetch_desc(&map_a2); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(bar(B_XFULL + s), 1); mbar_init(bar(B_XEMPTY + s), 1);
      mbar_init(bar(B_OFULL + s), 1); mbar_init(bar(B_OEMPTY + s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(B_SFULL + s), 1); mbar_init(bar(B_SEMPTY + s), kDwWarps);
      mbar_init(bar(B_ZFULL + s), kDwWarps); mbar_init(bar(B_ZEMPTY + s), 1);
      mbar_init(bar(B_DFULL + s), 1); mbar_init(bar(B_DEMPTY + s), 4);
      mbar_init(bar(B_CFULL + s), 1); mbar_init(bar(B_CEMPTY + s), kDwWarps);
    }
    mbar_init(bar(B_WFULL), 1); mbar_init(bar(B_FFULL), 1);
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

  if (warp == 0) {
    // ===================================================================================== TMA producer
    if (n_units) {
      if (elect_one()) {
        const uint32_t wbytes = 3 * C::KB_IN * C::A_BLOCK + (C::HAS_OTHER ? C::KB_OTHER * C::A_BLOCK : 0) + (C::HAS_OUT ? 2 * C_OUT * 128 : 0);
        mbar_expect_tx(bar(B_WFULL), wbytes);
        for (int i = 0; i < 3; ++i)
          for (int kb = 0; kb < C::KB_IN; ++kb)
            tma_load_3d(smem_base + C::OFF_A + (i * C::KB_IN + kb) * C::A_BLOCK, &map_a, bar(B_WFULL), kb * 64, 0, i);
        if (C::HAS_OTHER)
          for (int kb = 0; kb < C::KB_OTHER; ++kb) tma_load_2d(smem_base + C::OFF_A2 + kb * C::A_BLOCK, &map_a2, bar(B_WFULL), kb * 64, 0);
        if (C::HAS_OUT)
          for (int rb = 0; rb < 2; ++rb) tma_load_2d(smem_base + C::OFF_B + rb * (C_OUT * 128), &map_b, bar(B_WFULL), rb * 64, 0);
      }
      __syncwarp();
      for (uint32_t q = 0; q < n_units; ++q) {
        const int u = u_begin + (int)q, n = u / 28, j = u - n * 28;
        const uint32_t st = q % XS;
        mbar_wait(bar(B_XEMPTY + st), ((q / XS) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(bar(B_XFULL + st), C::X_STAGE);
#pragma unroll
          for (int kb = 0; kb < C::KB_IN; ++kb)
            tma_load_4d(smem_base + C::OFF_X + st * C::X_STAGE + kb * kBoxBytes, &map_in, bar(B_XFULL + st), kb * 64, 0, 2 * j - 1, n);
        }
        __syncwarp();
        if (C::HAS_OTHER) {
          const uint32_t ost = q % OS;
          mbar_wait(bar(B_OEMPTY + ost), ((q / OS) & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx(bar(B_OFULL + ost), C::O_STAGE);
#pragma unroll
            for (int kb = 0; kb < C::KB_OTHER; ++kb)
              tma_load_4d(smem_base + C::OFF_O + ost * C::O_STAGE + kb * kChunkBytes, &map_other, bar(B_OFULL + ost), kb * 64, 0, 2 * j, n);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================================== stage-1 issuer: folded pass + centre
    if (n_units) {
      constexpr uint32_t idesc1 = make_idesc(128, NP, 0, 0);            // D[r][pos] : A K-major, positions K-major
      mbar_wait(bar(B_WFULL), 0);
      const uint32_t a_lo = desc_lo(smem_base + C::OFF_A, 16), a2_lo = desc_lo(smem_base + C::OFF_A2, 16);
      const uint32_t x_lo = desc_lo(smem_base + C::OFF_X, 16), o_lo = desc_lo(smem_base + C::OFF_O, 16);
      for (uint32_t q = 0; q < n_units; ++q) {
        const uint32_t st = q % XS, slot = q % NS;
        mbar_wait(bar(B_XFULL + st), (q / XS) & 1);
        mbar_wait(bar(B_SEMPTY + slot), ((q / NS) & 1) ^ 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int i = 0; i < 3; ++i)                                   // tap i: the box shifted by i rows (56 positions = 7168 bytes)
#pragma unroll
            for (int kk = 0; kk < C::KB_IN * 4; ++kk)
              umma_lo(tmem_base + slot * NP, a_lo + (((i * C::KB_IN + (kk >> 2)) * C::A_BLOCK + (kk & 3) * 32) >> 4),
                      x_lo + ((st * C::X_STAGE + (kk >> 2) * kBoxBytes + i * 7168 + (kk & 3) * 32) >> 4), idesc1, (i | kk) ? 1u : 0u);
          umma_commit(bar(B_XEMPTY + st));
          umma_commit(bar(B_SFULL + slot));
        }
        __syncwarp();
        if (C::HAS_OTHER) {                                             // centre: D[r][pos] = A2[r][c] . other[pos][c]^T
          const uint32_t ost = q % OS, cb = q % NC;
          mbar_wait(bar(B_OFULL + ost), (q / OS) & 1);
          mbar_wait(bar(B_CEMPTY + cb), ((q / NC) & 1) ^ 1);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < C::KB_OTHER * 4; ++kk)
              umma_lo(tmem_base + C::COL_C + cb * NP, a2_lo + (((kk >> 2) * C::A_BLOCK + (kk & 3) * 32) >> 4),
                      o_lo + ((ost * C::O_STAGE + (kk >> 2) * kChunkBytes + (kk & 3) * 32) >> 4), idesc1, kk > 0 ? 1u : 0u);
            umma_commit(bar(B_CFULL + cb));
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 3) {
    // ===================================================================================== stage-3 / weight-gradient issuer
    if (n_units) {
      constexpr uint32_t idesc2 = make_idesc(128, C_OUT, 1, 0);         // D2[pos][o] : z MN-major, B K-major
      constexpr uint32_t idesc3 = make_idesc(128, C::C_OTHER, 0, 1);    // dF[r][c]   : z K-major (rows = rank), other MN-major
      const int ksteps = p.KR / 16;
      mbar_wait(bar(B_WFULL), 0);
      const uint32_t zm_lo = desc_lo(smem_base + C::OFF_Z, C::Z_BLOCK), zk_lo = desc_lo(smem_base + C::OFF_Z, 16);
      const uint32_t b_lo = desc_lo(smem_base + C::OFF_B, 16), o_lo = desc_lo(smem_base + C::OFF_O, kChunkBytes);
      for (uint32_t q = 0; q < n_units; ++q) {
        const uint32_t zb = q & 1;
        mbar_wait(bar(B_ZFULL + zb), (q >> 1) & 1);
        if (C::HAS_OUT) {
          const uint32_t dbuf = q % ND;
          mbar_wait(bar(B_DEMPTY + dbuf), ((q / ND) & 1) ^ 1);
          tc_fence_after();
          if (elect_one()) {
            for (int kk = 0; kk < ksteps; ++kk)
              umma_lo(tmem_base + C::COL_D + dbuf * C_OUT, zm_lo + ((zb * C::Z_BYTES + kk * 2048) >> 4),
                      b_lo + (((kk >> 2) * (C_OUT * 128) + (kk & 3) * 32) >> 4), idesc2, kk > 0 ? 1u : 0u);
            umma_commit(bar(B_DFULL + dbuf));
          }
          __syncwarp();
        }
        if (C::HAS_OTHER) {
          const uint32_t ost = q % OS;
          mbar_wait(bar(B_OFULL + ost), (q / OS) & 1);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int s = 0; s < NP / 16; ++s)
              umma_lo(tmem_base + C::COL_F, zk_lo + ((zb * C::Z_BYTES + (s >> 2) * C::Z_BLOCK + (s & 3) * 32) >> 4),
                      o_lo + ((ost * C::O_STAGE + s * 2048) >> 4), idesc3, (q | (uint32_t)s) ? 1u : 0u);
            umma_commit(bar(B_OEMPTY + ost));
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(bar(B_ZEMPTY + zb));
        __syncwarp();
      }
      if (elect_one()) umma_commit(bar(B_FFULL));
      __syncwarp();
    }
  } else if (warp >= 4 && warp < 4 + kDwWarps) {
    // ===================================================================================== second pass (CUDA cores)
    const int wg = (warp - 4) >> 2, quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const bool active = quarter * 32 < p.KR;
    const uint32_t lane_bits = (uint32_t)(quarter * 32) << 16;
    const float k0 = __ldg(p.taps + r), k1 = __ldg(p.taps + 128 + r), k2 = __ldg(p.taps + 256 + r);
    f2 kk[3] = {pk(k0, k0), pk(k1, k1), pk(k2, k2)};
    f2 g[3] = {0ull, 0ull, 0ull};
    for (uint32_t q = 0; q < n_units; ++q) {
      const uint32_t slot = q % NS, zb = q & 1, cb = q % NC;
      mbar_wait(bar(B_SFULL + slot), (q / NS) & 1);
      mbar_wait(bar(B_ZEMPTY + zb), ((q >> 1) & 1) ^ 1);
      if (C::HAS_OTHER) mbar_wait(bar(B_CFULL + cb), (q / NC) & 1);
      tc_fence_after();
      if (active)
        pass_row<C::HAS_OTHER, C::Z_BLOCK>(tmem_base + lane_bits + slot * NP + wg * 56, tmem_base + lane_bits + C::COL_C + cb * NP + wg * 56,
                                           kk, k1, smem_base + C::OFF_Z + zb * C::Z_BYTES, wg * 56, r, r < p.KR, g);
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(B_SEMPTY + slot));
        if (C::HAS_OTHER) mbar_arrive(bar(B_CEMPTY + cb));
        mbar_arrive(bar(B_ZFULL + zb));
      }
    }
    if (C::HAS_OTHER && active && r < p.R) {
#pragma unroll
      for (int i = 0; i < 3; ++i) red_add_f32(p.dG + i * 128 + r, lo_of(g[i]) + hi_of(g[i]));
    }
  } else if (warp >= 12) {
    // ===================================================================================== epilogue
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;                      // position of the tile = TMEM lane
    const uint32_t lane_bits = (uint32_t)(quarter * 32) << 16;
    if (C::HAS_OUT) {
      for (uint32_t q = 0; q < n_units; ++q) {
        const int u = u_begin + (int)q, n = u / 28, j = u - n * 28;
        const uint32_t dbuf = q % ND, sb = q % YS;
        mbar_wait(bar(B_DFULL + dbuf), (q / ND) & 1);
        tc_fence_after();
        if (warp == 12 && lane == 0) { if (YS == 2) bulk_wait_read1(); else bulk_wait_read0(); }   // staging buffer sb has been read
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
            for (int jj = 0; jj < 4; ++jj) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = F(v[8 * jj + e]);
              if (BIAS) {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] += __ldg(p.bias + cb * 32 + 8 * jj + e);
              }
              st_shared_v4(row + ((((uint32_t)((cb & 1) * 4 + jj)) ^ (uint32_t)(m & 7)) << 4), pack_bf16x2(f[0], f[1]),
                           pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
            }
          }
        }
        fence_proxy_async();
        named_bar_sync(1, kEpiThreads);
        if (warp == 12 && lane == 0) {
          for (int cb = 0; cb < C::KB_OUT; ++cb) tma_store_4d(&map_out, ybuf + cb * kChunkBytes, cb * 64, 0, 2 * j, n);
          bulk_commit();
        }
      }
      if (warp == 12 && lane == 0) bulk_wait_read0();
    }
    if (C::HAS_OTHER && n_units) {
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
// Host side
// --------------------------------------------------------------------------------------------------------------------------------
// 4-D map of a (B, H, W, C) bf16 tensor, box {64, 56, box_rows, 1}.  transposed: dims (C, H, W, B) -- the second coordinate walks
// H, the third W -- so that "a row of 56 positions" is an image column.
static int make_map4d(CUtensorMap* map, const void* ptr, int64_t c, int64_t w, int64_t h, int64_t n, int box_rows, bool transposed) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)"); return TLT_ERR_CUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)(transposed ? h : w), (cuuint64_t)(transposed ? w : h), (cuuint64_t)n};
  const cuuint64_t sw = (cuuint64_t)c * 2, sh = (cuuint64_t)c * w * 2;
  cuuint64_t strides[3] = {transposed ? sh : sw, transposed ? sw : sh, (cuuint64_t)c * w * h * 2};
  cuuint32_t box[4] = {64, 56, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = CUDA_SUCCESS;
  for (int attempt = 0; attempt < 2; ++attempt) {
    r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_ERROR_INVALID_CONTEXT) break;       // see make_map
    cudaFree(0);
  }
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(4d) failed (%d) for dims {%lld,%lld,%lld,%lld} box {64,56,%d,1}", (int)r, (long long)c, (long long)w,
              (long long)h, (long long)n, box_rows);
    return TLT_ERR_CUDA;
  }
  return TLT_OK;
}

template <int MODE, int C_IN, int C_OUT, int KRT, bool BIAS>
static int launch_t(int batch, int rank, bool transposed, const void* in, void* out, const void* other, const void* a3, const void* a2,
                    const void* b, const float* taps, const float* bias, float* dF, float* dG, cudaStream_t st) {
  using C = Cfg<MODE, C_IN, C_OUT, KRT>;
  static_assert(C::SMEM <= 227 * 1024, "shared memory budget");
  static_assert(C::COL_END <= 512, "TMEM budget");
  CUtensorMap m_in, m_out, m_other, m_a, m_a2, m_b;
  int rc;
  if ((rc = make_map4d(&m_in, in, C_IN, 56, 56, batch, 4, transposed)) != TLT_OK) return rc;
  m_out = m_in; m_other = m_in; m_a2 = m_in; m_b = m_in;
  if (C::HAS_OUT && (rc = make_map4d(&m_out, out, C_OUT, 56, 56, batch, 2, transposed)) != TLT_OK) return rc;
  if (C::HAS_OTHER && (rc = make_map4d(&m_other, other, C::C_OTHER, 56, 56, batch, 2, transposed)) != TLT_OK) return rc;
  if ((rc = make_map3d(&m_a, a3, C_IN, 128, 3, C_IN, (int64_t)128 * C_IN, 64, KRT, true)) != TLT_OK) return rc;      // [3][128][C_IN]
  if (C::HAS_OTHER && (rc = make_map(&m_a2, a2, C::C_OTHER, 128, C::C_OTHER, 64, KRT)) != TLT_OK) return rc;           // [128][C_OTHER]
  if (C::HAS_OUT && (rc = make_map(&m_b, b, 128, C_OUT, 128, 64, C_OUT)) != TLT_OK) return rc;                         // [C_OUT][128]
  Params p;
  p.units_total = batch * 28;
  p.R = rank; p.KR = (rank + 15) / 16 * 16;
  p.taps = taps; p.bias = bias; p.dF = dF; p.dG = dG;
  auto kern = cpconv_rows2_kernel<MODE, C_IN, C_OUT, KRT, BIAS>;
  TLT_ENSURE_SMEM(kern, C::SMEM);
  int grid = num_sms();
  if (grid > p.units_total) grid = p.units_total;
  kern<<<grid, kThreads, C::SMEM, st>>>(m_in, m_out, m_other, m_a, m_a2, m_b, p);
  return check_launch("cpconv_rows2_kernel");
}

template <int MODE, int C_IN, int C_OUT, bool BIAS>
static int launch_k(int batch, int rank, bool transposed, const void* in, void* out, const void* other, const void* a3, const void* a2,
                    const void* b, const float* taps, const float* bias, float* dF, float* dG, cudaStream_t st) {
  // one instantiation: 80 rank rows per shared-memory operand block (rank <= 80; larger ranks take the first-generation kernels)
  return launch_t<MODE, C_IN, C_OUT, 80, BIAS>(batch, rank, transposed, in, out, other, a3, a2, b, taps, bias, dF, dG, st);
}

// entry points used by cpconv_rows.cu (64 -> 64 channels, 56 x 56 maps, stride 1)
int fwd_64(int batch, int rank, const void* x, void* y, const void* a3, const void* b, const float* taps, const float* bias, cudaStream_t st) {
  if (bias) return launch_k<MODE_FWD, 64, 64, true>(batch, rank, false, x, y, nullptr, a3, nullptr, b, taps, bias, nullptr, nullptr, st);
  return launch_k<MODE_FWD, 64, 64, false>(batch, rank, false, x, y, nullptr, a3, nullptr, b, taps, nullptr, nullptr, nullptr, st);
}
int bwd_data_64(int batch, int rank, const void* dy, void* dx, const void* x, const void* a3, const void* a2, const void* b,
                const float* taps, float* dF, float* dG, cudaStream_t st) {
  return launch_k<MODE_BWD_DATA, 64, 64, false>(batch, rank, true, dy, dx, x, a3, a2, b, taps, nullptr, dF, dG, st);
}
int bwd_weight_64(int batch, int rank, const void* x, const void* dy, const void* a3, const void* a2, const float* taps, float* dF,
                  float* dG, cudaStream_t st) {
  return launch_k<MODE_BWD_WEIGHT, 64, 64, false>(batch, rank, false, x, nullptr, dy, a3, a2, nullptr, taps, nullptr, dF, dG, st);
}

}  // namespace cpr2
}  // namespace tlt
