// Fused CP convolution (3x3, stride 1, padding 1) on NCHW bf16 activations: the whole chain
//
//        y = F_out . dw3x3( F_in^T . x ) (+ bias)          x (B, C, H, W) -> y (B, O, H, W),   R-channel intermediates on chip
//
// in ONE launch per direction, the R-channel intermediates never leaving the SM.  The data gradient is the same chain with
// the roles of the two channel factors swapped and the taps flipped (flip_taps = 1), so one kernel serves forward and dgrad.
//
// Reference: cp_conv / cp_conv_mobilenet, tltorch/functional/convolution.py:231-345 -- F.conv1d(k=1) -> d depthwise
// F.conv1d passes -> F.conv1d(k=1): three-plus full HBM round trips of the R-channel tensor per direction.  The d separable
// depthwise passes are merged into one 3x3 depthwise kernel (taps[t, r] = lambda_r prod_j F_{2+j}[k_j, r], built by
// kr_ops.khatri_rao), exactly as the unfused route of this library does.
//
// Design (BASELINE config 3: ResNet-18's 56^2 x 64 and 28^2 x 128 layers are HBM-bound, 2 flop-bytes apart from the tensor
// roofline, so legacy HMMA is enough and the kernel is judged on HBM GB/s):
//   * persistent CTAs; a tile = (image, strip of TH output rows, all W columns, all channels).  The input strip with a
//     one-row halo -- C runs of (TH+2) W contiguous bf16 -- is brought in with cp.async (16-byte chunks; 8-byte when W % 8 != 0)
//     into XS[c][p] (p = strip position), rows outside the image zero-filled;
//   * the rank is walked in chunks of 16:   Z1c[16][P_in]  = W1[chunk] . XS          (mma.sync m16n8k16, B operand = XS via
//     ldmatrix.trans: positions are contiguous in NCHW memory, so the activation tile is the N-contiguous operand),
//     Z2c[16][P_out] = depthwise 3x3 of Z1c (fp32 FMA from shared memory, bf16 out -- the reference rounds here too),
//     Yacc[O][P_out] += W3[:, chunk] . Z2c  (accumulators live in registers across the whole rank loop);
//   * both channel factors (F (channels, R) row-major, as the reference stores them) and the taps are staged ONCE per CTA
//     into zero-padded shared-memory operands (rank padded to a multiple of 16);
//   * epilogue: (+ bias) -> bf16 -> staged in the dead XS buffer -> 16-byte coalesced stores of O contiguous runs.
//   * optional z1_out / z2_out (B, R, H, W): the two intermediates of the OWNED rows, for the unfused weight-gradient
//     kernels (pointwise_wgrad_tc, dwconv_wgrad) -- written from shared memory as contiguous runs.
// Every row pitch in shared memory is an odd multiple of 16 bytes: all ldmatrix phases and fragment stores are conflict-free.
#include "tc_common.cuh"

namespace tlt {
namespace cpf {

// ---------------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t addr, uint2 v) { asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_bf16(uint32_t addr) {          // one bf16 element -> fp32
  uint32_t v;
  asm volatile("{.reg .b16 t; ld.shared.b16 t, [%1]; cvt.u32.u16 %0, t;}" : "=r"(v) : "r"(addr));
  return __uint_as_float(v << 16);
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
template <int BYTES> __device__ __forceinline__ void cp_async(uint32_t dst, const void* src) {
  if constexpr (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

__host__ __device__ constexpr int odd16(int bytes) {                 // smallest odd multiple of 16 that is >= bytes
  int n = (bytes + 15) / 16;
  if (n % 2 == 0) ++n;
  return n * 16;
}

struct Params {
  const __nv_bfloat16* x;        // (B, C, H, W)
  const __nv_bfloat16* f_in;     // (C, R) row-major: W1[r][c] = f_in[c * R + r]
  const __nv_bfloat16* taps_t;   // (9, R) taps-major (lambda folded in)
  const __nv_bfloat16* f_out;    // (O, R) row-major: W3[o][r] = f_out[o * R + r]
  const void* bias;              // (O) or null
  int bias_f32;
  __nv_bfloat16* y;              // (B, O, H, W)
  __nv_bfloat16* z1_out;         // (B, R, H, W) or null
  __nv_bfloat16* z2_out;         // (B, R, H, W) or null
  int B, H, R, RP;               // RP = R rounded up to 16
  int flip;                      // 1: use taps[8 - t] (data gradient)
  int w3p;                       // row pitch in bytes of operands whose rows run over the rank = odd16(RP * 2)
  long long tiles;
};

// ---------------------------------------------------------------------------------------------------------------------
// Building blocks shared by the forward / dgrad kernel and the one-pass backward kernel.  All operate on a chunk of up to
// 32 rank channels (`two` = the second 16 are live).
// ---------------------------------------------------------------------------------------------------------------------

// Strip of TH + 2 image rows (one-row halo) of CH channels -> S[c][p], rows outside the image zero-filled.
//   (W * 2) % 16 == 0: one TMA bulk copy per channel (its rows are one contiguous run of NCHW memory), mbarrier-signalled;
//   otherwise 8-byte cp.async chunks.
template <int CH, int W, int TH, int SP, int NT>
__device__ __forceinline__ void load_strip(uint32_t dst, const __nv_bfloat16* src_img, int H, int h0, uint32_t bar, uint32_t parity,
                                           int tid) {
  constexpr int RB = W * 2;                                           // bytes per image row
  if constexpr (RB % 16 == 0) {
    const int lo = h0 == 0 ? 1 : 0;                                   // first / one-past-last strip row that exists
    const int hi = h0 + TH + 1 > H ? TH + 1 : TH + 2;
    const uint32_t bytes = (uint32_t)(hi - lo) * RB;
    if (tid < 32) {
      fence_proxy_async();                                            // S was last touched through the generic proxy
      if (tid == 0) mbar_expect_tx(bar, bytes * CH);
      __syncwarp();
      const __nv_bfloat16* s0 = src_img + (size_t)(h0 - 1 + lo) * W;
      for (int c = tid; c < CH; c += 32)
        bulk_g2s(dst + c * SP + lo * RB, s0 + (size_t)c * H * W, bytes, bar);
    }
    if (lo == 1)
      for (int e = tid; e < CH * (RB / 16); e += NT) sts128(dst + (e / (RB / 16)) * SP + (e % (RB / 16)) * 16, make_uint4(0, 0, 0, 0));
    if (hi == TH + 1)
      for (int e = tid; e < CH * (RB / 16); e += NT)
        sts128(dst + (e / (RB / 16)) * SP + (TH + 1) * RB + (e % (RB / 16)) * 16, make_uint4(0, 0, 0, 0));
    mbar_wait(bar, parity);
  } else {
    constexpr int CPR = (TH + 2) * RB / 8, CPW = RB / 8;
    for (int e = tid; e < CH * CPR; e += NT) {
      const int c = e / CPR, j = e % CPR;
      const int h = h0 - 1 + j / CPW;
      const uint32_t d = dst + c * SP + j * 8;
      if (h >= 0 && h < H) cp_async<8>(d, src_img + ((size_t)c * H + h) * W + (j % CPW) * 4);
      else sts64(d, make_uint2(0, 0));
    }
    cp_async_wait_all();
  }
}

// D[32][16 NG] = A[32][KC] . S[KC][16 NG]      A rows K-contiguous (pitch AP), S rows = channels, positions contiguous (pitch SP),
// D rows = rank channels of the chunk, bf16 (pitch SP).  Warp `warp` of NWARPS owns position groups warp, warp + NWARPS, ...
template <int KC, int NG, int NWARPS, int SP, int AP>
__device__ __forceinline__ void stage1_chunk(uint32_t a_base, uint32_t src, uint32_t dst, bool two, int warp, int lane) {
  constexpr int GP = (NG + NWARPS - 1) / NWARPS;
  const int gid = lane >> 2, tq = lane & 3, l15 = lane & 15, lhi = lane >> 4;
  float acc[GP][2][2][4];
#pragma unroll
  for (int gi = 0; gi < GP; ++gi)
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[gi][m][n][i] = 0.f;
#pragma unroll
  for (int kk = 0; kk < KC / 16; ++kk) {
    uint32_t a0[4], a1[4] = {0u, 0u, 0u, 0u};
    ldsm4(a0, a_base + l15 * AP + (kk * 16 + lhi * 8) * 2);
    if (two) ldsm4(a1, a_base + (16 + l15) * AP + (kk * 16 + lhi * 8) * 2);
#pragma unroll
    for (int gi = 0; gi < GP; ++gi) {
      const int g = warp + gi * NWARPS;
      if (g < NG) {
        uint32_t bf[4];
        ldsm4t(bf, src + (kk * 16 + l15) * SP + (g * 16 + lhi * 8) * 2);
        mma16816(acc[gi][0][0], a0, bf[0], bf[1]);
        mma16816(acc[gi][0][1], a0, bf[2], bf[3]);
        if (two) {
          mma16816(acc[gi][1][0], a1, bf[0], bf[1]);
          mma16816(acc[gi][1][1], a1, bf[2], bf[3]);
        }
      }
    }
  }
#pragma unroll
  for (int gi = 0; gi < GP; ++gi) {
    const int g = warp + gi * NWARPS;
    if (g < NG) {
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        if (m == 0 || two) {
#pragma unroll
          for (int n = 0; n < 2; ++n) {
            const uint32_t col = (g * 16 + n * 8 + 2 * tq) * 2;
            sts32(dst + (m * 16 + gid) * SP + col, pack2(acc[gi][m][n][0], acc[gi][m][n][1]));
            sts32(dst + (m * 16 + gid + 8) * SP + col, pack2(acc[gi][m][n][2], acc[gi][m][n][3]));
          }
        }
      }
    }
  }
}

// Depthwise 3x3 of `nrows` rank channels: Z2[r][h][w] = sum_{i,j} taps[r][3 i + j] Z1[r][h + i][w + j - 1], fp32 accumulate.
// A work item = (channel, block of RB output rows, strip of 4 columns): RB + 2 input rows of one 8-byte word each, the two
// halo columns come from the neighbouring lanes' words (consecutive lanes own consecutive strips).
template <int W, int TH, int RB, int NT, int SP, int DP>
__device__ __forceinline__ void dw_chunk(uint32_t z1c, uint32_t z2c, uint32_t taps, int nrows, int tid) {
  constexpr int NS = W / 4, NRB = TH / RB, IPR = NS * NRB;
  static_assert((16 * IPR) % 32 == 0 && W % 4 == 0 && TH % RB == 0, "whole warps of depthwise work items");
  const int lane = tid & 31;
  const int items = nrows * IPR;
  for (int i = tid; i - lane < items; i += NT) {                      // warp-uniform trip count (items % 32 == 0)
    const int rr = i / IPR, rem = i % IPR;
    const int rb = rem / NS, s = rem % NS;
    float t[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) t[k] = lds_f32(taps + rr * 36 + k * 4);
    const uint32_t src = z1c + rr * SP + (rb * RB * W + s * 4) * 2;
    const uint32_t dst = z2c + rr * DP + (rb * RB * W + s * 4) * 2;
    float acc[RB][4];
#pragma unroll
    for (int h = 0; h < RB; ++h)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[h][c] = 0.f;
#pragma unroll
    for (int j = 0; j < RB + 2; ++j) {
      const uint2 q = lds64(src + j * W * 2);
      const uint32_t ql = __shfl_up_sync(0xffffffffu, q.y, 1), qr = __shfl_down_sync(0xffffffffu, q.x, 1);
      float v[6];
      v[0] = s == 0 ? 0.f : (lane == 0 ? lds_bf16(src + j * W * 2 - 2) : bf_hi(ql));
      v[5] = s == NS - 1 ? 0.f : (lane == 31 ? lds_bf16(src + j * W * 2 + 8) : bf_lo(qr));
      v[1] = bf_lo(q.x); v[2] = bf_hi(q.x); v[3] = bf_lo(q.y); v[4] = bf_hi(q.y);
#pragma unroll
      for (int dh = 0; dh < 3; ++dh) {
        const int h = j - dh;
        if (h >= 0 && h < RB) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            acc[h][c] = fmaf(t[dh * 3], v[c], fmaf(t[dh * 3 + 1], v[c + 1], fmaf(t[dh * 3 + 2], v[c + 2], acc[h][c])));
        }
      }
    }
#pragma unroll
    for (int h = 0; h < RB; ++h) sts64(dst + h * W * 2, make_uint2(pack2(acc[h][0], acc[h][1]), pack2(acc[h][2], acc[h][3])));
  }
}

// Yacc[16 MT][16 NG3] += A[16 MT][32] . Z[32][16 NG3]   A rows K-contiguous at runtime pitch AP (a_base already points at the
// warp's first row and the chunk's first column), Z rows = rank channels of the chunk (pitch ZP).
template <int MT, int GPW, int NGW, int NG3, int ZP>
__device__ __forceinline__ void stage3_chunk(float (&yacc)[GPW][MT][2][4], uint32_t a_base, int AP, uint32_t zc, bool two, int gw,
                                             int lane) {
  const int l15 = lane & 15, lhi = lane >> 4;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    if (ks == 0 || two) {
      uint32_t a3[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) ldsm4(a3[mt], a_base + (mt * 16 + l15) * AP + (ks * 16 + lhi * 8) * 2);
#pragma unroll
      for (int gi = 0; gi < GPW; ++gi) {
        const int g = gw + gi * NGW;
        if (g < NG3) {
          uint32_t bf[4];
          ldsm4t(bf, zc + (ks * 16 + l15) * ZP + (g * 16 + lhi * 8) * 2);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            mma16816(yacc[gi][mt][0], a3[mt], bf[0], bf[1]);
            mma16816(yacc[gi][mt][1], a3[mt], bf[2], bf[3]);
          }
        }
      }
    }
  }
}

// Rows [row0, row0 + nrows) of a chunk buffer (owned positions only) -> global (B, R, H, W) tensor, CB-byte chunks.
template <int P_OUT, int CB, int NT>
__device__ __forceinline__ void save_chunk(uint32_t src, int pitch, int skip_bytes, __nv_bfloat16* out_img, size_t plane, int rc, int R,
                                           int nrows, int tid) {
  constexpr int CPO = P_OUT * 2 / CB;
  for (int e = tid; e < nrows * CPO; e += NT) {
    const int rr = e / CPO, j = e % CPO;
    if (rc + rr < R) {
      unsigned char* dst = reinterpret_cast<unsigned char*>(out_img + (size_t)(rc + rr) * plane) + j * CB;
      const uint32_t s = src + rr * pitch + skip_bytes + j * CB;
      if constexpr (CB == 16) *reinterpret_cast<uint4*>(dst) = lds128(s);
      else *reinterpret_cast<uint2*>(dst) = lds64(s);
    }
  }
}

// Zero-padded operand staging: dst[row][col] (pitch bytes) = src[row * s_row + col * s_col] for row < rows, col < cols, else 0
template <int NT>
__device__ __forceinline__ void stage_matrix(unsigned char* dst, int pitch, int rows_p, int cols_p, const __nv_bfloat16* src, int rows,
                                             int cols, int s_row, int s_col, bool col_fastest_in_src, int tid) {
  for (int e = tid; e < rows_p * cols_p; e += NT) {
    const int row = col_fastest_in_src ? e / cols_p : e % rows_p;
    const int col = col_fastest_in_src ? e % cols_p : e / rows_p;
    const __nv_bfloat16 v = (row < rows && col < cols) ? src[(size_t)row * s_row + (size_t)col * s_col] : __float2bfloat16_rn(0.f);
    *reinterpret_cast<__nv_bfloat16*>(dst + (size_t)row * pitch + col * 2) = v;
  }
}

template <int C_, int O_, int W_, int TH_, int NWARPS_, int MSPLIT_, int MINB_, int RB_>
struct Cfg {
  static constexpr int C = C_, O = O_, W = W_, TH = TH_, NWARPS = NWARPS_, MSPLIT = MSPLIT_, MINB = MINB_, RB = RB_;
  static constexpr int NT = NWARPS * 32;
  static constexpr int P_IN = (TH + 2) * W, P_OUT = TH * W;
  static constexpr int NG1 = (P_IN + 15) / 16;           // n16 position groups of stage 1
  static constexpr int NG3 = P_OUT / 16;                 // n16 position groups of stage 3
  static constexpr int XP = odd16(NG1 * 32);             // XS / Z1C row pitch (bytes)
  static constexpr int Z2P = odd16(P_OUT * 2);           // Z2C / YS row pitch
  static constexpr int W1P = odd16(C * 2);
  static constexpr int CB = (W % 8 == 0) ? 16 : 8;       // bytes per global <-> shared chunk (never straddles an image row)
  static constexpr int MT = O / 16 / MSPLIT;             // m16 tiles per warp in stage 3
  static constexpr int NGW = NWARPS / MSPLIT;            // warps (slots) that own position groups
  static constexpr int GPW = (NG3 + NGW - 1) / NGW;      // position groups per warp
  static_assert(P_OUT % 16 == 0, "strip must hold a multiple of 16 output positions");
  static_assert(C % 16 == 0 && O % (16 * MSPLIT) == 0 && NWARPS % MSPLIT == 0, "channel tiling");
  static_assert(O * Z2P <= C * XP, "output staging must fit into the input strip buffer");
  static size_t smem_bytes(int RP, int w3p) {
    return (size_t)C * XP + (size_t)(RP + 16) * W1P + (size_t)O * w3p + (size_t)(RP + 16) * 9 * 4 + 32 * XP + 32 * Z2P + 64;
  }
};

template <class K>
__global__ void __launch_bounds__(K::NT, K::MINB) cpconv3x3_fused_kernel(const Params p) {
  constexpr int C = K::C, O = K::O, W = K::W, TH = K::TH, NT = K::NT;
  constexpr int P_OUT = K::P_OUT, XP = K::XP, Z2P = K::Z2P, W1P = K::W1P, CB = K::CB;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar_store;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int RP = p.RP, R = p.R, W3P = p.w3p;
  unsigned char* xs_g = smem;
  unsigned char* w1s_g = xs_g + (size_t)C * XP;                       // [RP + 16][C]   (16 spare zero rows: chunks are 32 wide)
  unsigned char* w3s_g = w1s_g + (size_t)(RP + 16) * W1P;             // [O][RP (+16)]  (pitch covers RP + 16 columns)
  float* taps_g = reinterpret_cast<float*>(w3s_g + (size_t)O * W3P);  // [RP + 16][9]
  unsigned char* z1c_g = reinterpret_cast<unsigned char*>(taps_g + (RP + 16) * 9);
  unsigned char* z2c_g = z1c_g + 32 * XP;
  const uint32_t xs = s_u32(xs_g), w1s = s_u32(w1s_g), w3s = s_u32(w3s_g), tps = s_u32(taps_g), z1c = s_u32(z1c_g),
                 z2c = s_u32(z2c_g), bar = s_u32(&bar_store);

  // ---- stage the operands once per CTA (zero-padded in the rank) ----
  stage_matrix<NT>(w1s_g, W1P, RP + 16, C, p.f_in, R, C, 1, R, false, tid);          // W1S[r][c] = f_in[c][r]
  stage_matrix<NT>(w3s_g, W3P, O, RP + 16, p.f_out, O, R, R, 1, true, tid);          // W3S[o][r] = f_out[o][r]
  for (int e = tid; e < (RP + 16) * 9; e += NT) {
    const int r = e % (RP + 16), t = e / (RP + 16);
    const int ts = p.flip ? 8 - t : t;
    taps_g[r * 9 + t] = r < R ? __bfloat162float(p.taps_t[(size_t)ts * R + r]) : 0.f;
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncthreads();

  const int nstrip = p.H / TH;
  const int gid = lane >> 2, tq = lane & 3;
  const int msub = warp % K::MSPLIT, gw = warp / K::MSPLIT;
  uint32_t parity = 0;

  for (long long tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
    const int b = (int)(tile / nstrip), strip = (int)(tile % nstrip);
    const int h0 = strip * TH;

    load_strip<C, W, TH, XP, NT>(xs, p.x + (size_t)b * C * p.H * W, p.H, h0, bar, parity, tid);
    parity ^= 1;
    __syncthreads();

    float yacc[K::GPW][K::MT][2][4];
#pragma unroll
    for (int gi = 0; gi < K::GPW; ++gi)
#pragma unroll
      for (int mt = 0; mt < K::MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) yacc[gi][mt][nt][i] = 0.f;

    for (int rc = 0; rc < RP; rc += 32) {
      const bool two = rc + 16 < RP;
      const int nrows = two ? 32 : 16;
      // stage 1: Z1c = W1[rc .. rc+32) . XS
      stage1_chunk<C, K::NG1, K::NWARPS, XP, W1P>(w1s + rc * W1P, xs, z1c, two, warp, lane);
      __syncthreads();
      if (p.z1_out != nullptr)
        save_chunk<P_OUT, CB, NT>(z1c, XP, W * 2, p.z1_out + ((size_t)b * R * p.H + h0) * W, (size_t)p.H * W, rc, R, nrows, tid);
      // stage 2: depthwise 3x3
      dw_chunk<W, TH, K::RB, NT, XP, Z2P>(z1c, z2c, tps + rc * 36, nrows, tid);
      __syncthreads();
      if (p.z2_out != nullptr)
        save_chunk<P_OUT, CB, NT>(z2c, Z2P, 0, p.z2_out + ((size_t)b * R * p.H + h0) * W, (size_t)p.H * W, rc, R, nrows, tid);
      // stage 3: Yacc += W3[:, rc .. rc+32) . Z2c
      stage3_chunk<K::MT, K::GPW, K::NGW, K::NG3, Z2P>(yacc, w3s + (msub * K::MT * 16) * W3P + rc * 2, W3P, z2c, two, gw, lane);
      // no barrier here: the next chunk's stage 1 only writes Z1C (whose readers passed the barrier above) and its
      // depthwise stage -- the next writer of Z2C -- sits behind the barrier that follows stage 1
    }
    __syncthreads();                                                  // every warp is done with XS (stage 1) and Z2C

    // ---- epilogue: (+ bias) -> bf16 -> YS (in the XS buffer) -> coalesced stores ----
#pragma unroll
    for (int gi = 0; gi < K::GPW; ++gi) {
      const int g = gw + gi * K::NGW;
      if (g < K::NG3) {
#pragma unroll
        for (int mt = 0; mt < K::MT; ++mt) {
          const int o0 = (msub * K::MT + mt) * 16 + gid;
          float b0 = 0.f, b1 = 0.f;
          if (p.bias != nullptr) {
            if (p.bias_f32) { b0 = reinterpret_cast<const float*>(p.bias)[o0]; b1 = reinterpret_cast<const float*>(p.bias)[o0 + 8]; }
            else { b0 = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.bias)[o0]); b1 = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.bias)[o0 + 8]); }
          }
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            const uint32_t col = (g * 16 + nt * 8 + 2 * tq) * 2;
            sts32(xs + o0 * Z2P + col, pack2(yacc[gi][mt][nt][0] + b0, yacc[gi][mt][nt][1] + b0));
            sts32(xs + (o0 + 8) * Z2P + col, pack2(yacc[gi][mt][nt][2] + b1, yacc[gi][mt][nt][3] + b1));
          }
        }
      }
    }
    __syncthreads();
    {
      constexpr int CPO = P_OUT * 2 / CB;
      __nv_bfloat16* yb = p.y + ((size_t)b * O * p.H + h0) * W;
      for (int e = tid; e < O * CPO; e += NT) {
        const int o = e / CPO, j = e % CPO;
        unsigned char* dst = reinterpret_cast<unsigned char*>(yb + (size_t)o * p.H * W) + j * CB;
        const uint32_t src = xs + o * Z2P + j * CB;
        if constexpr (CB == 16) *reinterpret_cast<uint4*>(dst) = lds128(src);
        else *reinterpret_cast<uint2*>(dst) = lds64(src);
      }
    }
    fence_proxy_async();                                              // the next strip is written through the async proxy
    __syncthreads();                                                  // YS is read out before the next strip lands in XS
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Streaming 1x1 channel mix on NCHW activations with the same building blocks:  y[b, n, s] = sum_k wt[n, k] x[b, k, s].
// A tile = (image, PT consecutive positions): every channel's run is one contiguous span, landed by a 1-D TMA bulk copy into
// a double buffer (the next tile always in flight); the rank is walked in chunks of 32 output channels with stage1_chunk and
// each chunk leaves as 16-byte coalesced stores of contiguous runs.  The tcgen05 kernel (pointwise_tc.cu) feeds its MN-major
// operand with tiled-mode TMA boxes of 128-byte rows and transposes every 32 x 32 block in its epilogue; at config 3's 56^2 /
// 28^2 shapes it moves 2.6 - 3.3 TB/s (tools/pointwise_bench.py), this kernel is the HBM-bound alternative for small K, N.
// ---------------------------------------------------------------------------------------------------------------------
struct PwhParams {
  const __nv_bfloat16* x;        // (B, C, S)
  const __nv_bfloat16* wt;       // (N, w_ld) K-major
  __nv_bfloat16* y;              // (B, N, S)
  int C, N, S, w_ld, RP;         // RP = N rounded up to 16
  int tpi;                       // tiles per image
  long long tiles;
};

template <int KC, int PT, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, 1) pointwise_hmma_kernel(const PwhParams p) {
  constexpr int NT = NWARPS * 32, NG = (PT + 15) / 16, XP = odd16(NG * 32), W1P = odd16(KC * 2);
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.C, N = p.N, S = p.S, RP = p.RP;
  unsigned char* xs_g = smem;                                         // [2][KC][XP]
  unsigned char* zc_g = xs_g + (size_t)2 * KC * XP;                   // [32][XP]
  unsigned char* w1s_g = zc_g + 32 * XP;                              // [RP + 16][W1P]
  const uint32_t xs = s_u32(xs_g), zc = s_u32(zc_g), w1s = s_u32(w1s_g), bar0 = s_u32(&bars[0]);

  for (int e = tid; e < 2 * (KC - C) * (XP / 4); e += NT) {           // channel rows C .. KC stay zero (never landed on)
    const int slot = e / ((KC - C) * (XP / 4)), r = e % ((KC - C) * (XP / 4));
    reinterpret_cast<uint32_t*>(xs_g + ((size_t)slot * KC + C) * XP)[r] = 0u;
  }
  stage_matrix<NT>(w1s_g, W1P, RP + 16, KC, p.wt, N, C, p.w_ld, 1, true, tid);
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](long long tile, int slot) {                        // executed by warp 0
    const long long b = tile / p.tpi;
    const int p0 = (int)(tile % p.tpi) * PT;
    const uint32_t bytes = (uint32_t)min(PT, S - p0) * 2;
    fence_proxy_async();
    if (lane == 0) mbar_expect_tx(bar0 + 8 * slot, bytes * C);
    __syncwarp();
    const __nv_bfloat16* src = p.x + ((size_t)b * C) * S + p0;
    for (int c = lane; c < C; c += 32) bulk_g2s(xs + ((uint32_t)slot * KC + c) * XP, src + (size_t)c * S, bytes, bar0 + 8 * slot);
  };
  const long long first = blockIdx.x, step = gridDim.x;
  if (warp == 0 && first < p.tiles) issue(first, 0);
  int it = 0;
  for (long long tile = first; tile < p.tiles; tile += step, ++it) {
    const int slot = it & 1;
    if (warp == 0 && tile + step < p.tiles) issue(tile + step, slot ^ 1);   // its readers passed the barrier that ended it - 1
    mbar_wait(bar0 + 8 * slot, (it >> 1) & 1);
    const long long b = tile / p.tpi;
    const int p0 = (int)(tile % p.tpi) * PT;
    const int plen = min(PT, S - p0), cpr = plen / 8;                 // 16-byte chunks per output run
    const uint32_t xt = xs + (uint32_t)slot * KC * XP;
    for (int rc = 0; rc < RP; rc += 32) {
      const bool two = rc + 16 < RP;
      const int nrows = min(two ? 32 : 16, N - rc);
      stage1_chunk<KC, NG, NWARPS, XP, W1P>(w1s + rc * W1P, xt, zc, two, warp, lane);
      __syncthreads();
      unsigned char* yb = reinterpret_cast<unsigned char*>(p.y + (((size_t)b * N + rc) * S + p0));
      for (int e = tid; e < nrows * cpr; e += NT) {
        const int rr = e / cpr, j = e - rr * cpr;
        *reinterpret_cast<uint4*>(yb + (size_t)rr * S * 2 + j * 16) = lds128(zc + rr * XP + j * 16);
      }
      __syncthreads();
    }
  }
}

template <int KC, int PT>
static int pwh_launch(const PwhParams& p0, cudaStream_t st) {
  constexpr int NWARPS = 8, NG = (PT + 15) / 16, XP = odd16(NG * 32), W1P = odd16(KC * 2);
  PwhParams p = p0;
  p.tpi = (p.S + PT - 1) / PT;
  const size_t smem = (size_t)2 * KC * XP + 32 * XP + (size_t)(p.RP + 16) * W1P + 64;
  if (smem > 227 * 1024) return 1;
  TLT_ENSURE_SMEM((pointwise_hmma_kernel<KC, PT, NWARPS>), smem);
  const int sms = device_sms();
  long long tiles = p.tiles = (p.tiles /* batch */) * p.tpi;
  const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
  pointwise_hmma_kernel<KC, PT, NWARPS><<<grid, NWARPS * 32, smem, st>>>(p);
  return check_launch("pointwise_hmma_kernel");
}

}  // namespace cpf

// returns 1 when the shape is not taken (caller falls back to the tcgen05 kernel), else a tlt_status
int pointwise_hmma(const void* x, const void* wt, void* y, long long batch, long long C, long long N, long long S, long long w_ld,
                   cudaStream_t st) {
  if (S % 8 != 0 || C < 16 || C > 144 || N < 1 || N > 160 || batch < 1) return 1;
  cpf::PwhParams p{};
  p.x = (const __nv_bfloat16*)x; p.wt = (const __nv_bfloat16*)wt; p.y = (__nv_bfloat16*)y;
  p.C = (int)C; p.N = (int)N; p.S = (int)S; p.w_ld = (int)w_ld; p.RP = (int)((N + 15) / 16 * 16);
  p.tiles = batch;
  const int kc = (int)((C + 15) / 16 * 16);
  if (kc <= 64) return cpf::pwh_launch<64, 448>(p, st);
  if (kc <= 80) return cpf::pwh_launch<80, 448>(p, st);
  if (kc <= 128) return cpf::pwh_launch<128, 224>(p, st);
  return cpf::pwh_launch<144, 224>(p, st);
}

namespace cpf {

template <class K>
static int launch(const Params& p, cudaStream_t st) {
  const size_t smem = K::smem_bytes(p.RP, p.w3p);
  if (smem > 227 * 1024) {
    set_error("tlt_cpconv3x3_fused: rank %d needs %zu bytes of shared memory", p.R, smem);
    return TLT_ERR_UNSUPPORTED;
  }
  TLT_ENSURE_SMEM(cpconv3x3_fused_kernel<K>, smem);
  int bps = 0;
  TLT_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, cpconv3x3_fused_kernel<K>, K::NT, smem));
  const int cached_grid = device_sms() * (bps < 1 ? 1 : bps);
  long long grid = cached_grid;
  if (grid > p.tiles) grid = p.tiles;
  cpconv3x3_fused_kernel<K><<<(unsigned)grid, K::NT, smem, st>>>(p);
  return check_launch("cpconv3x3_fused_kernel");
}

using CfgA = Cfg<64, 64, 56, 4, 7, 1, 2, 4>;      // 2 CTAs / SM, 4-row strips
using CfgB = Cfg<64, 64, 56, 8, 14, 1, 1, 4>;     // 1 CTA / SM, 8-row strips
using CfgC = Cfg<128, 128, 28, 4, 14, 2, 1, 2>;

}  // namespace cpf
}  // namespace tlt

using namespace tlt;

namespace {
// variant table: (C, O, W) -> tiling.  variant 0 = default, 1 = alternative strip height (development switch)
int pick(const tlt_cpconv_desc* d) {
  if (d->in_channels == 64 && d->out_channels == 64 && d->width == 56) return d->height % 8 == 0 && d->variant == 1 ? 2 : (d->height % 4 == 0 ? 1 : 0);
  if (d->in_channels == 128 && d->out_channels == 128 && d->width == 28) return d->height % 4 == 0 ? 3 : 0;
  return 0;
}
}  // namespace

extern "C" int tlt_cpconv3x3_fused_supported(const tlt_cpconv_desc* d) {
  if (!d || d->batch < 1 || d->rank < 1 || d->height < 1) return 0;
  const int v = pick(d);
  if (!v) return 0;
  const int RP = (int)((d->rank + 15) / 16 * 16), w3p = cpf::odd16((RP + 16) * 2);
  size_t smem = 0;
  if (v == 1) smem = cpf::CfgA::smem_bytes(RP, w3p);
  if (v == 2) smem = cpf::CfgB::smem_bytes(RP, w3p);
  if (v == 3) smem = cpf::CfgC::smem_bytes(RP, w3p);
  return smem <= 227 * 1024 ? 1 : 0;
}

extern "C" int tlt_cpconv3x3_fused(const tlt_cpconv_desc* d, const void* x, const void* f_in, const void* taps_t, const void* f_out,
                                   const void* bias, void* y, void* z1_out, void* z2_out, void* stream) {
  TLT_REQUIRE(d && x && f_in && taps_t && f_out && y, "tlt_cpconv3x3_fused: null argument");
  const int v = pick(d);
  if (!v || !tlt_cpconv3x3_fused_supported(d)) {
    set_error("tlt_cpconv3x3_fused: unsupported shape C=%lld O=%lld H=%lld W=%lld R=%lld", (long long)d->in_channels,
              (long long)d->out_channels, (long long)d->height, (long long)d->width, (long long)d->rank);
    return TLT_ERR_UNSUPPORTED;
  }
  cpf::Params p;
  p.x = (const __nv_bfloat16*)x;
  p.f_in = (const __nv_bfloat16*)f_in;
  p.taps_t = (const __nv_bfloat16*)taps_t;
  p.f_out = (const __nv_bfloat16*)f_out;
  p.bias = bias;
  p.bias_f32 = d->dtype_bias == TLT_F32;
  p.y = (__nv_bfloat16*)y;
  p.z1_out = (__nv_bfloat16*)z1_out;
  p.z2_out = (__nv_bfloat16*)z2_out;
  p.B = (int)d->batch;
  p.H = (int)d->height;
  p.R = (int)d->rank;
  p.RP = (p.R + 15) / 16 * 16;
  p.flip = d->flip_taps;
  p.w3p = cpf::odd16((p.RP + 16) * 2);
  cudaStream_t st = (cudaStream_t)stream;
  if (v == 1) { p.tiles = (long long)p.B * (p.H / 4); return cpf::launch<cpf::CfgA>(p, st); }
  if (v == 2) { p.tiles = (long long)p.B * (p.H / 8); return cpf::launch<cpf::CfgB>(p, st); }
  p.tiles = (long long)p.B * (p.H / 4);
  return cpf::launch<cpf::CfgC>(p, st);
}
