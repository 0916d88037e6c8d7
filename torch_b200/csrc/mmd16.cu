// Fused three-mode product chains for tensorized samples whose modes are all <= 16 (bf16), warp-level tensor cores.
//
//   forward   y[b] = x[b] x_0 M0 x_1 M1 x_2 M2                       x (batch, d0,d1,d2) -> y (batch, r0,r1,r2)
//   backward  du[b] = g[b] x_0 M0^T x_1 M1^T x_2 M2^T,  dM_k = sum_b <g[b], u[b] x_{j != k} M_j>_k   in ONE pass over (u, g)
//
// Reference call sites: the project-in / expand-out stages of linear_tucker (tltorch/functional/factorized_linear.py:7-21,
// factorized_tensordot.py:10-66 evaluated in FLOP-sane order) and their autograd; tenalg.multi_mode_dot generally.
// At BASELINE config 5 (16^3 <-> 15^3 per token, 8192 tokens) the chain is HBM-bound: 4 bytes of traffic per element per
// pass, ~0.35 MFLOP per token.  The first version (modedot.cu: SIMT FFMA from fp32 shared memory, one launch for the data
// gradient plus, per factor, one more fused launch and a long-reduction contraction) spent 0.3-0.4 ms per launch and 14
// launches per step on it.  Here:
//   * every mode is zero-padded to 16, so a sample is a 16x16x16 bf16 tile (8 KB) and every mode product is a set of
//     m16n8k16 HMMA (mma.sync) with K = 16 -- one k-step -- whose second operand comes straight out of shared memory with
//     ldmatrix(.trans); the fp32 accumulators are rounded to bf16 and stored for the next stage (what the reference's
//     bf16 mode_dot chain does as well);
//   * a CTA of 4 warps owns a sample; each stage's 16 independent (ldmatrix, 2 HMMA, 4 stores) iterations are dealt 4 per
//     warp with one __syncthreads between stages.  (v1 gave each WARP a sample: 32 KB of tiles per warp capped the SM at 6
//     warps and the kernel was latency-bound at 0.29 IPC -- profiles/r01_mmd16_v1.md; a 4-warp team needs the same 32 KB
//     for 4 warps, so 20-24 warps are resident and another CTA's stage covers each barrier);
//   * tiles live in a padded linear layout (row pitch 48 B, slab pitch 784 B) in which the 8-row ldmatrix phases are
//     conflict-free whether the rows run along mode 0 or mode 1, and every address is a per-thread base + immediate;
//   * backward recomputes the two forward intermediates, then alternates "matrix gradient" (K = the other two modes,
//     accumulated in registers across all samples of the warp) and "data gradient" stages through 4 tile buffers;
//     factor gradients leave the kernel as one fp32 atomicAdd per entry per CTA.
//   * the NEXT sample of the CTA is always in flight: one thread issues a TMA bulk copy (cp.async.bulk, one instruction per
//     8 KB sample) into a linear landing buffer as soon as the previous landing has been re-tiled; an mbarrier with a
//     transaction count signals arrival.  (v2 loaded each sample synchronously with 16-byte ld.global: with <= 40 KB of
//     loads in flight per SM for less than half of the time the kernel was bound by memory-level parallelism at 2.4 TB/s
//     whatever the occupancy.)
// Compact global tensors whose innermost mode is < 16 (e.g. 15^3 rank-space tensors at a pitch of 3376) are staged
// linearly with 16-byte loads and re-rowed inside shared memory.
#include "tc_common.cuh"
#include <mutex>

namespace tlt {
namespace m16 {

constexpr int kMatPitchB = 48;              // bytes per staged matrix row (16 bf16 + 8 pad: conflict-free ldmatrix)
constexpr int kMatBytes = 16 * kMatPitchB;  // 768
constexpr int kWarps = 4, kThreads = kWarps * 32;   // one team = one CTA

struct Params {
  int d[3], r[3];                           // forward input / output extents of the three modes (1..16)
  int trans[3];                             // matrix k stored (d_k, r_k) instead of (r_k, d_k)
  unsigned magic_d1, magic_r1;              // ceil(65536 / n1): row -> (a, b) split without a division
  long long batch;
  long long ld_x, ld_y;                     // sample pitches (elements) of the d-shaped and of the r-shaped tensor
};

// ---------------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s_u32(const void* p) { return smem_u32(p); }
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
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v) {
  asm volatile("{.reg .b16 t; cvt.u16.u32 t, %1; st.shared.b16 [%0], t;}" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
  uint32_t v;
  asm volatile("{.reg .b16 t; ld.shared.b16 t, [%1]; cvt.u32.u16 %0, t;}" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 ldg128(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Byte offset of the 16-byte chunk (a, b, ch) of a 16x16x16 bf16 tile in shared memory.  Rows (a, b) are 32 bytes of data at a
// pitch of 48 bytes and the 16 rows of one `a` are followed by 16 bytes of padding (slab pitch 784 = 6 * 128 + 16), so 8
// consecutive b AND 8 consecutive a land in 8 different 16-byte bank groups: every ldmatrix phase and every fragment
// store is conflict-free whichever mode the rows run along.  The layout is LINEAR on purpose: with the iteration indices
// known at compile time every address is one per-thread base plus an immediate (an XOR swizzle packs the tile into 8 KB
// but cost ~8 integer instructions per access -- half of the instruction stream of v2).
constexpr int kRowPitch = 48, kSlabPitch = 16 * kRowPitch + 16;     // 784
constexpr int kTile = 16 * kSlabPitch;                              // 12544 bytes
constexpr int kLand = 8192;                                         // linear landing / staging buffer of one sample
__device__ __forceinline__ uint32_t sw_off(int a, int b, int ch) { return (uint32_t)(a * kSlabPitch + b * kRowPitch + ch * 16); }

// lane -> (row, chunk) patterns of the four ldmatrix.x4 uses
//   P16 : row = lane % 16, chunk = lane / 16      A fragment of a K-contiguous tile / B fragment (.trans) of an N-contiguous tile
//   P8  : row = lane % 8 + 8 * (lane / 16), chunk = (lane / 8) % 2      B fragment of a K-contiguous ("col") tile / A fragment
//                                                                        (.trans) of an M-contiguous tile
#define P16_ROW(l) ((l) & 15)
#define P16_CH(l) ((l) >> 4)
#define P8_ROW(l) (((l) & 7) + (((l) >> 4) << 3))
#define P8_CH(l) (((l) >> 3) & 1)

// ---------------------------------------------------------------------------------------------------------------------
// Staged matrices: Mat_k[j][i] (j = output index < r_k, i = input index < d_k), zero-padded to 16x16, row pitch 48 bytes
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_matrices(unsigned char* ms, const Params& p, const __nv_bfloat16* m0,
                                               const __nv_bfloat16* m1, const __nv_bfloat16* m2) {
  const __nv_bfloat16* mats[3] = {m0, m1, m2};
  for (int e = threadIdx.x; e < 3 * 256; e += blockDim.x) {
    const int k = e >> 8, j = (e >> 4) & 15, i = e & 15;
    __nv_bfloat16 v = __float2bfloat16_rn(0.f);
    if (mats[k] == nullptr) {
      if (i == j) v = __float2bfloat16_rn(1.f);                       // untouched mode: identity
    } else if (j < p.r[k] && i < p.d[k]) {
      v = p.trans[k] ? mats[k][i * p.r[k] + j] : mats[k][j * p.d[k] + i];
    }
    *reinterpret_cast<__nv_bfloat16*>(ms + k * kMatBytes + j * kMatPitchB + i * 2) = v;
  }
}
// A[m = j][k = i] = Mat[j][i]
__device__ __forceinline__ void mat_a(uint32_t (&f)[4], uint32_t mb, int lane) { ldsm4(f, mb + P16_ROW(lane) * kMatPitchB + P16_CH(lane) * 16); }
// A[m = i][k = j] = Mat[j][i]   (transposed matrix as the left operand)
__device__ __forceinline__ void mat_at(uint32_t (&f)[4], uint32_t mb, int lane) { ldsm4t(f, mb + P8_ROW(lane) * kMatPitchB + P8_CH(lane) * 16); }
// B[k = i][n = j] = Mat[j][i]   ({f0,f1}: n 0-7, {f2,f3}: n 8-15)
__device__ __forceinline__ void mat_b(uint32_t (&f)[4], uint32_t mb, int lane) { ldsm4(f, mb + P8_ROW(lane) * kMatPitchB + P8_CH(lane) * 16); }
// B[k = j][n = i] = Mat[j][i]   (transposed matrix as the right operand)
__device__ __forceinline__ void mat_bt(uint32_t (&f)[4], uint32_t mb, int lane) { ldsm4t(f, mb + P16_ROW(lane) * kMatPitchB + P16_CH(lane) * 16); }

// ---------------------------------------------------------------------------------------------------------------------
// Stages.  Tiles are addressed T[a][b][c] (c innermost).
// ---------------------------------------------------------------------------------------------------------------------
// Every stage has 16 independent iterations; warp w of the team owns iterations 4w .. 4w+3.  The PTX wrappers are
// `asm volatile`, so program order is issue order: all ldmatrix first, then the HMMAs, then the stores.
//
// MODE 0: out[j][o][c] = sum_k A[j][k] in[k][o][c]        MODE 1: out[o][j][c] = sum_k A[j][k] in[o][k][c]
template <int MODE>
__device__ __forceinline__ void stage_x(uint32_t in, uint32_t out, const uint32_t (&am)[4], int lane, int w) {
  const int g = lane >> 2, q4 = (lane & 3) * 4;
  const int kr = P16_ROW(lane), kc = P16_CH(lane);
  uint32_t b[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int o = 4 * w + u;
    ldsm4t(b[u], in + (MODE == 0 ? sw_off(kr, o, kc) : sw_off(o, kr, kc)));
  }
  float c[4][2][4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
#pragma unroll
    for (int i = 0; i < 4; ++i) c[u][0][i] = c[u][1][i] = 0.f;
    mma16816(c[u][0], am, b[u][0], b[u][1]);
    mma16816(c[u][1], am, b[u][2], b[u][3]);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int o = 4 * w + u;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = g + 8 * h;
      sts32(out + (MODE == 0 ? sw_off(j, o, 0) : sw_off(o, j, 0)) + q4, pack2(c[u][0][2 * h], c[u][0][2 * h + 1]));
      sts32(out + (MODE == 0 ? sw_off(j, o, 1) : sw_off(o, j, 1)) + q4, pack2(c[u][1][2 * h], c[u][1][2 * h + 1]));
    }
  }
}
// out[a][b][j] = sum_c in[a][b][c] B[c][j]
__device__ __forceinline__ void stage_y(uint32_t in, uint32_t out, const uint32_t (&bm)[4], int lane, int w) {
  const int g = lane >> 2, q4 = (lane & 3) * 4;
  const int ar = P16_ROW(lane), ac = P16_CH(lane);
  uint32_t af[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u) ldsm4(af[u], in + sw_off(4 * w + u, ar, ac));
  float c[4][2][4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
#pragma unroll
    for (int i = 0; i < 4; ++i) c[u][0][i] = c[u][1][i] = 0.f;
    mma16816(c[u][0], af[u], bm[0], bm[1]);
    mma16816(c[u][1], af[u], bm[2], bm[3]);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int a = 4 * w + u;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int bb = g + 8 * h;
      sts32(out + sw_off(a, bb, 0) + q4, pack2(c[u][0][2 * h], c[u][0][2 * h + 1]));
      sts32(out + sw_off(a, bb, 1) + q4, pack2(c[u][1][2 * h], c[u][1][2 * h + 1]));
    }
  }
}
// matrix gradient of an X stage:  MODE 0: acc[j][i] += sum_{o,c} gout[j][o][c] in[i][o][c]     MODE 1: ... gout[o][j][c] in[o][i][c]
// (this warp's share of the reduction: o = 4w .. 4w+3)
template <int MODE>
__device__ __forceinline__ void grad_x(uint32_t gout, uint32_t in, float (&acc)[2][4], int lane, int w) {
  const int ar = P16_ROW(lane), ac = P16_CH(lane), br = P8_ROW(lane), bc = P8_CH(lane);
  uint32_t af[4][4], bf[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int o = 4 * w + u;
    ldsm4(af[u], gout + (MODE == 0 ? sw_off(ar, o, ac) : sw_off(o, ar, ac)));
    ldsm4(bf[u], in + (MODE == 0 ? sw_off(br, o, bc) : sw_off(o, br, bc)));
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    mma16816(acc[0], af[u], bf[u][0], bf[u][1]);
    mma16816(acc[1], af[u], bf[u][2], bf[u][3]);
  }
}
// matrix gradient of the Y stage:  acc[j][i] += sum_{a,b} gout[a][b][j] in[a][b][i]      (a = 4w .. 4w+3)
__device__ __forceinline__ void grad_y(uint32_t gout, uint32_t in, float (&acc)[2][4], int lane, int w) {
  const int ar = P8_ROW(lane), ac = P8_CH(lane), br = P16_ROW(lane), bc = P16_CH(lane);
  uint32_t af[4][4], bf[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    ldsm4t(af[u], gout + sw_off(4 * w + u, ar, ac));
    ldsm4t(bf[u], in + sw_off(4 * w + u, br, bc));
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    mma16816(acc[0], af[u], bf[u][0], bf[u][1]);
    mma16816(acc[1], af[u], bf[u][2], bf[u][3]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Global <-> tile (all kThreads threads of the team; callers put a __syncthreads after load_tile / before reusing buffers)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void zero_tile(uint32_t buf, int tid) {
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
  for (int i = 0; i < (kTile / 16 + kThreads - 1) / kThreads; ++i)
    if (i * kThreads + tid < kTile / 16) sts128(buf + (i * kThreads + tid) * 16, z);
}

// TMA bulk copy of one sample (bytes = round8(elements) * 2, a multiple of 16) into a linear landing buffer
__device__ __forceinline__ void issue_sample(uint32_t dst, const __nv_bfloat16* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// linear landing buffer `stg` holding a compact (n0, n1, n2) sample -> swizzled tile `dst`.  Only the valid rows are written
// (each as a full 32-byte row, zero past n2); rows with a >= n0 or b >= n1 keep the zeros put there by zero_tile.
__device__ __forceinline__ void unpack_tile(uint32_t stg, int n0, int n1, int n2, unsigned magic1, uint32_t dst, int tid) {
  const int rows = n0 * n1;
  if (n2 == 16) {
    const int nch = rows * 2;                      // <= 512: at most 4 chunks per thread
    uint4 val[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int v = u * kThreads + tid;
      if (v < nch) val[u] = lds128(stg + v * 16);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int v = u * kThreads + tid;
      if (v < nch) {
        const int r = v >> 1;
        const int a = (int)((unsigned)r * magic1 >> 16), b = r - a * n1;
        sts128(dst + sw_off(a, b, v & 1), val[u]);
      }
    }
  } else {
    for (int r = tid; r < rows; r += kThreads) {
      const int a = (int)((unsigned)r * magic1 >> 16), b = r - a * n1;
      const uint32_t sb = stg + (uint32_t)(r * n2) * 2;
      uint32_t w[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t lo = 2 * k < n2 ? lds16(sb + 4 * k) : 0u;
        const uint32_t hi = 2 * k + 1 < n2 ? lds16(sb + 4 * k + 2) : 0u;
        w[k] = lo | (hi << 16);
      }
      sts128(dst + sw_off(a, b, 0), make_uint4(w[0], w[1], w[2], w[3]));
      sts128(dst + sw_off(a, b, 1), make_uint4(w[4], w[5], w[6], w[7]));
    }
  }
}

__device__ __forceinline__ uint32_t add_bf16x2(uint32_t v, uint32_t a) {
  const float2 x = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
  const float2 y = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&a));
  return pack2(x.x + y.x, x.y + y.y);
}

// swizzled tile `src` -> compact (n0, n1, n2) sample at `dst`; elements [n0 n1 n2, round8(n0 n1 n2)) are written as zeros.
// `addend` (n0 n1 n2 elements) is added when not null.  The caller has synchronised the team on `src`; ends with a
// __syncthreads (src and stg are free afterwards).
__device__ __forceinline__ void store_tile(uint32_t src, __nv_bfloat16* __restrict__ dst, int n0, int n1, int n2, unsigned magic1,
                                           uint32_t stg, const __nv_bfloat16* __restrict__ addend, int tid) {
  const int rows = n0 * n1;
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  const uint4* a4 = reinterpret_cast<const uint4*>(addend);
  if (n2 == 16) {
    const int nch = rows * 2;
    for (int v = tid; v < nch; v += kThreads) {
      const int r = v >> 1;
      const int a = (int)((unsigned)r * magic1 >> 16), b = r - a * n1;
      uint4 val = lds128(src + sw_off(a, b, v & 1));
      if (addend) {
        const uint4 ad = __ldg(a4 + v);
        val.x = add_bf16x2(val.x, ad.x); val.y = add_bf16x2(val.y, ad.y);
        val.z = add_bf16x2(val.z, ad.z); val.w = add_bf16x2(val.w, ad.w);
      }
      d4[v] = val;
    }
  } else {
    const int total = rows * n2, nvec = (total + 7) >> 3;
    if (tid == 0) sts128(stg + (nvec - 1) * 16, make_uint4(0u, 0u, 0u, 0u));
    __syncthreads();
    for (int r = tid; r < rows; r += kThreads) {
      const int a = (int)((unsigned)r * magic1 >> 16), b = r - a * n1;
      const uint4 lo = lds128(src + sw_off(a, b, 0)), hi = lds128(src + sw_off(a, b, 1));
      const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
      const uint32_t sb = stg + (uint32_t)(r * n2) * 2;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (2 * k < n2) sts16(sb + 4 * k, w[k] & 0xffffu);
        if (2 * k + 1 < n2) sts16(sb + 4 * k + 2, w[k] >> 16);
      }
    }
    __syncthreads();
    for (int v = tid; v < nvec; v += kThreads) {
      uint4 val = lds128(stg + v * 16);
      if (addend) {
        if (v * 8 + 8 <= total) {
          const uint4 ad = __ldg(a4 + v);
          val.x = add_bf16x2(val.x, ad.x); val.y = add_bf16x2(val.y, ad.y);
          val.z = add_bf16x2(val.z, ad.z); val.w = add_bf16x2(val.w, ad.w);
        } else {
          uint32_t w[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int e = v * 8 + 2 * k;
            const uint32_t alo = e < total ? (uint32_t)__bfloat16_as_ushort(addend[e]) : 0u;
            const uint32_t ahi = e + 1 < total ? (uint32_t)__bfloat16_as_ushort(addend[e + 1]) : 0u;
            w[k] = add_bf16x2(w[k], alo | (ahi << 16));
          }
          val = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      d4[v] = val;
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------------
// Kernels
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kHdr = 3 * kMatBytes + 128;                                        // staged matrices + mbarrier slot
constexpr int kFwdSmem = 128 + kHdr + 2 * kTile + 2 * kLand;                     // 43.1 KB: A, B, landing, output staging
constexpr int kBwdSmem = 128 + kHdr + 3 * 256 * 4 + 4 * kTile + 2 * kLand;       // 71.1 KB: U, G, T1, T2, two landings
constexpr int kFwdCtasPerSm = 5, kBwdCtasPerSm = 3;

__global__ void __launch_bounds__(kThreads, kFwdCtasPerSm) mmd16_fwd_kernel(const Params p, const __nv_bfloat16* __restrict__ x,
                                                                            const __nv_bfloat16* __restrict__ m0,
                                                                            const __nv_bfloat16* __restrict__ m1,
                                                                            const __nv_bfloat16* __restrict__ m2,
                                                                            const __nv_bfloat16* __restrict__ addend,
                                                                            __nv_bfloat16* __restrict__ y) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const uint32_t mb = s_u32(sm);
  const uint32_t bar = mb + 3 * kMatBytes;
  const uint32_t A = mb + kHdr, B = A + kTile, L = B + kTile, O = L + kLand;
  const uint32_t x_bytes = (uint32_t)((p.d[0] * p.d[1] * p.d[2] + 7) & ~7) * 2;
  const bool full_in = p.d[0] * p.d[1] == 256 && p.d[2] == 16;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
    if ((long long)blockIdx.x < p.batch) {
      mbar_expect_tx(bar, x_bytes);
      issue_sample(L, x + (long long)blockIdx.x * p.ld_x, x_bytes, bar);
    }
  }
  stage_matrices(sm, p, m0, m1, m2);
  zero_tile(A, tid);
  __syncthreads();
  uint32_t am0[4], am1[4], bm2[4];
  mat_a(am0, mb, lane);
  mat_a(am1, mb + kMatBytes, lane);
  mat_b(bm2, mb + 2 * kMatBytes, lane);
  uint32_t phase = 0;
  for (long long s = blockIdx.x; s < p.batch; s += gridDim.x) {
    mbar_wait(bar, phase);
    phase ^= 1;
    unpack_tile(L, p.d[0], p.d[1], p.d[2], p.magic_d1, A, tid);
    __syncthreads();                                     // landing buffer consumed, A complete
    if (tid == 0 && s + gridDim.x < p.batch) {
      mbar_expect_tx(bar, x_bytes);
      issue_sample(L, x + (s + gridDim.x) * p.ld_x, x_bytes, bar);
    }
    stage_x<0>(A, B, am0, lane, w);
    __syncthreads();
    stage_x<1>(B, A, am1, lane, w);
    __syncthreads();
    stage_y(A, B, bm2, lane, w);
    __syncthreads();
    if (!full_in) zero_tile(A, tid);                     // pad rows back to exact zeros for the next sample
    store_tile(B, y + s * p.ld_y, p.r[0], p.r[1], p.r[2], p.magic_r1, O, addend, tid);
  }
}

// u: the forward input (d-shaped, pitch ld_x); g: gradient w.r.t. the forward output (r-shaped, pitch ld_y);
// du: d-shaped, pitch ld_x; dM: fp32 [3][16][16] (j = output index, i = input index), ACCUMULATED into (caller zeroes).
__global__ void __launch_bounds__(kThreads, kBwdCtasPerSm) mmd16_bwd_kernel(const Params p, const __nv_bfloat16* __restrict__ u,
                                                                            const __nv_bfloat16* __restrict__ g,
                                                                            const __nv_bfloat16* __restrict__ m0,
                                                                            const __nv_bfloat16* __restrict__ m1,
                                                                            const __nv_bfloat16* __restrict__ m2,
                                                                            __nv_bfloat16* __restrict__ du, float* __restrict__ dM) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  float* red = reinterpret_cast<float*>(sm + kHdr);
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const uint32_t mb = s_u32(sm);
  const uint32_t bar = mb + 3 * kMatBytes;
  const uint32_t U = mb + kHdr + 3 * 256 * 4, G = U + kTile, T1 = G + kTile, T2 = T1 + kTile, LU = T2 + kTile, LG = LU + kLand;
  const bool need_dm = dM != nullptr;
  const uint32_t u_bytes = (uint32_t)((p.d[0] * p.d[1] * p.d[2] + 7) & ~7) * 2;
  const uint32_t g_bytes = (uint32_t)((p.r[0] * p.r[1] * p.r[2] + 7) & ~7) * 2;
  const uint32_t tx_bytes = g_bytes + (need_dm ? u_bytes : 0u);
  const bool full_u = p.d[0] * p.d[1] == 256 && p.d[2] == 16, full_g = p.r[0] * p.r[1] == 256 && p.r[2] == 16;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
    if ((long long)blockIdx.x < p.batch) {
      mbar_expect_tx(bar, tx_bytes);
      issue_sample(LG, g + (long long)blockIdx.x * p.ld_y, g_bytes, bar);
      if (need_dm) issue_sample(LU, u + (long long)blockIdx.x * p.ld_x, u_bytes, bar);
    }
  }
  stage_matrices(sm, p, m0, m1, m2);
  for (int e = tid; e < 3 * 256; e += kThreads) red[e] = 0.f;
  zero_tile(U, tid);
  zero_tile(G, tid);
  __syncthreads();
  float acc[3][2][4];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[k][t][i] = 0.f;
  uint32_t phase = 0;
  for (long long s = blockIdx.x; s < p.batch; s += gridDim.x) {
    uint32_t f[4];
    mbar_wait(bar, phase);
    phase ^= 1;
    unpack_tile(LG, p.r[0], p.r[1], p.r[2], p.magic_r1, G, tid);
    if (need_dm) unpack_tile(LU, p.d[0], p.d[1], p.d[2], p.magic_d1, U, tid);
    __syncthreads();                                     // landings consumed; U, G complete
    if (tid == 0 && s + gridDim.x < p.batch) {
      mbar_expect_tx(bar, tx_bytes);
      issue_sample(LG, g + (s + gridDim.x) * p.ld_y, g_bytes, bar);
      if (need_dm) issue_sample(LU, u + (s + gridDim.x) * p.ld_x, u_bytes, bar);
    }
    if (need_dm) {
      mat_a(f, mb, lane);
      stage_x<0>(U, T1, f, lane, w);                     // t1[j0][i1][i2]
      __syncthreads();
      mat_a(f, mb + kMatBytes, lane);
      stage_x<1>(T1, T2, f, lane, w);                    // t2[j0][j1][i2]
      __syncthreads();
      grad_y(G, T2, acc[2], lane, w);                    // dM2[j2][i2]
      __syncthreads();
    }
    mat_bt(f, mb + 2 * kMatBytes, lane);
    stage_y(G, T2, f, lane, w);                          // dt2[j0][j1][i2]
    __syncthreads();
    if (need_dm) grad_x<1>(T2, T1, acc[1], lane, w);     // dM1[j1][i1]
    mat_at(f, mb + kMatBytes, lane);
    stage_x<1>(T2, G, f, lane, w);                       // dt1[j0][i1][i2]
    __syncthreads();
    if (need_dm) grad_x<0>(G, U, acc[0], lane, w);       // dM0[j0][i0]
    if (du != nullptr) {
      mat_at(f, mb, lane);
      stage_x<0>(G, T1, f, lane, w);                     // du[i0][i1][i2]
    }
    __syncthreads();
    if (!full_g) zero_tile(G, tid);                      // pad rows back to exact zeros for the next sample
    if (!full_u && need_dm) zero_tile(U, tid);
    if (du != nullptr) store_tile(T1, du + s * p.ld_x, p.d[0], p.d[1], p.d[2], p.magic_d1, T2, nullptr, tid);
    else __syncthreads();
  }
  if (need_dm) {
    const int gq = lane >> 2, q = lane & 3;
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = gq + 8 * (i >> 1), c = 2 * q + (i & 1) + 8 * t;
          atomicAdd(red + k * 256 + j * 16 + c, acc[k][t][i]);
        }
    __syncthreads();
    for (int e = tid; e < 3 * 256; e += kThreads) {
      const float v = red[e];
      if (v != 0.f) atomicAdd(dM + e, v);
    }
  }
}

static unsigned magic16(int n) { return (unsigned)((65536 + n - 1) / n); }

static int fill(const tlt_mmd16_desc* d, Params* p) {
  TLT_REQUIRE(d, "tlt_mmd16: null descriptor");
  TLT_REQUIRE(d->batch >= 1, "tlt_mmd16: empty batch");
  long long in_e = 1, out_e = 1;
  for (int k = 0; k < 3; ++k) {
    TLT_REQUIRE(d->in_dims[k] >= 1 && d->in_dims[k] <= 16 && d->out_dims[k] >= 1 && d->out_dims[k] <= 16,
                "tlt_mmd16: every mode extent must lie in 1..16");
    p->d[k] = d->in_dims[k]; p->r[k] = d->out_dims[k]; p->trans[k] = d->transpose[k];
    in_e *= p->d[k]; out_e *= p->r[k];
  }
  TLT_REQUIRE(d->ld_in % 8 == 0 && d->ld_out % 8 == 0 && d->ld_in >= ((in_e + 7) & ~7LL) && d->ld_out >= ((out_e + 7) & ~7LL),
              "tlt_mmd16: sample pitches must be multiples of 8 elements and cover round8(sample)");
  p->magic_d1 = magic16(p->d[1]); p->magic_r1 = magic16(p->r[1]);
  p->batch = d->batch; p->ld_x = d->ld_in; p->ld_y = d->ld_out;
  return TLT_OK;
}

static int num_sms16() { return device_sms(); }

}  // namespace m16
}  // namespace tlt

using namespace tlt;
using namespace tlt::m16;

#define ALIGNED16(p) ((p) == nullptr || ((uintptr_t)(p) % 16) == 0)

extern "C" int tlt_mmd16_fwd(const tlt_mmd16_desc* d, const void* x, const void* m0, const void* m1, const void* m2,
                             const void* addend, void* y, void* stream) {
  Params p{};
  int rc = fill(d, &p);
  if (rc) return rc;
  TLT_REQUIRE(x && y, "tlt_mmd16_fwd: null tensor");
  TLT_REQUIRE(ALIGNED16(x) && ALIGNED16(y) && ALIGNED16(addend), "tlt_mmd16_fwd: tensors must be 16-byte aligned");
  TLT_ENSURE_SMEM(mmd16_fwd_kernel, kFwdSmem);
  long long blocks = d->batch;
  const long long cap = (long long)num_sms16() * kFwdCtasPerSm;
  if (blocks > cap) blocks = cap;
  mmd16_fwd_kernel<<<(unsigned)blocks, kThreads, kFwdSmem, (cudaStream_t)stream>>>(
      p, (const __nv_bfloat16*)x, (const __nv_bfloat16*)m0, (const __nv_bfloat16*)m1, (const __nv_bfloat16*)m2,
      (const __nv_bfloat16*)addend, (__nv_bfloat16*)y);
  return check_launch("mmd16_fwd_kernel");
}

extern "C" int tlt_mmd16_bwd(const tlt_mmd16_desc* d, const void* u, const void* g, const void* m0, const void* m1,
                             const void* m2, void* du, float* dM, void* stream) {
  Params p{};
  int rc = fill(d, &p);
  if (rc) return rc;
  TLT_REQUIRE(g && (du || dM), "tlt_mmd16_bwd: null tensor");
  TLT_REQUIRE(!dM || u, "tlt_mmd16_bwd: the matrix gradients need the forward input");
  TLT_REQUIRE(ALIGNED16(u) && ALIGNED16(g) && ALIGNED16(du), "tlt_mmd16_bwd: tensors must be 16-byte aligned");
  TLT_ENSURE_SMEM(mmd16_bwd_kernel, kBwdSmem);
  long long blocks = d->batch;
  const long long cap = (long long)num_sms16() * kBwdCtasPerSm;
  if (blocks > cap) blocks = cap;
  mmd16_bwd_kernel<<<(unsigned)blocks, kThreads, kBwdSmem, (cudaStream_t)stream>>>(
      p, (const __nv_bfloat16*)u, (const __nv_bfloat16*)g, (const __nv_bfloat16*)m0, (const __nv_bfloat16*)m1,
      (const __nv_bfloat16*)m2, (__nv_bfloat16*)du, dM);
  return check_launch("mmd16_bwd_kernel");
}
