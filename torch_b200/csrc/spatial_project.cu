// Spatial-modes-first input projection of the tensor regression / contraction layers on channels-first feature maps.
//
//   forward   t[b, j, c] = sum_s x[b, c, s] k[s, j]                                 x (B, C, S) -> t (B, J, C)
//   backward  dx[b, c, s] = sum_j gt[b, j, c] k[s, j],   dk[s, j] += sum_{b,c} x[b, c, s] gt[b, j, c]
//
// with S = prod(spatial) <= 64 (7 x 7 = 49 at BASELINE config 4), J = prod(spatial ranks) <= 16 (4 x 4), k = the Kronecker
// product of the spatial factors.  Reference call site: tucker_trl's input projection, tenalg.multi_mode_dot(x, factors[:n],
// transpose=True) (tltorch/functional/tensor_regression.py:40) and TCL on the same maps (tensor_contraction_layers.py:75),
// evaluated spatial modes first: the activation is read ONCE where it lies in NCHW memory (no transposition to channels-last
// rows -- the 98-byte rows of a 7 x 7 map break TMA's and ldmatrix's 16-byte rule, which is why the previous plan paid two
// transposes of the 103 MB activation around a 30 us channel GEMM), shrinks 49 -> 16 per channel, and leaves t channels-last
// (K-major) for the channel-mode GEMM on tcgen05.
//
// HBM-bound by design: forward moves |x| + |t| bytes, backward 2|x| + |t|.  Per CTA (4 warps, persistent, several per SM):
//   * a tile = 128 consecutive channels of one image = ONE contiguous span of 128 S bf16; it is brought in by a single TMA bulk
//     copy (cp.async.bulk + mbarrier tx count) into a double buffer, the next tile always in flight;
//   * the products run on HMMA (mma.sync m16n8k16, fp32 accumulate): forward D[j][c] = K^T[j][s] X^T[s][c]; the activation
//     fragments are gathered with 2-byte shared loads (rows are only 2-byte aligned), the zero-padded K^T fragments stay in
//     registers for the CTA's whole life; out-of-row reads (s >= S) land on the next row and meet zeros of K^T;
//   * backward: dX^T[s][c] = K[s][j] gT[j][c] (gT tile via ldmatrix.trans) staged linearly in shared memory and written back
//     with one TMA bulk store per tile; dK^T[j][s] = gT[j][c] X[c][s] accumulates in registers across all tiles of the CTA,
//     is reduced across the 4 warps in shared memory and leaves as one fp32 atomicAdd per entry per CTA.
#include "tc_common.cuh"

namespace tlt {
namespace sp {

constexpr int kTileC = 128;                 // channels per tile
constexpr int kThreads = 128;
constexpr int kKP = 144;                    // K^T row pitch: 64 s * 2 B + 16 (odd multiple of 16: conflict-free ldmatrix)
constexpr int kOP = 272;                    // 128 c * 2 B + 16
constexpr int kK2P = 48;                    // K row pitch in the backward: 16 j * 2 B + 16

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
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
  uint32_t v;
  asm volatile("{.reg .b16 t; ld.shared.b16 t, [%1]; cvt.u32.u16 %0, t;}" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v) {
  asm volatile("{.reg .b16 t; cvt.u16.u32 t, %1; st.shared.b16 [%0], t;}" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t bf16_bits(float v) {
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  return (uint32_t)(*reinterpret_cast<unsigned short*>(&h));
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct Params {
  const __nv_bfloat16* x;     // (B, C, S)
  const __nv_bfloat16* k;     // (S, J)
  const __nv_bfloat16* gt;    // (B, J, C)     backward
  __nv_bfloat16* t;           // (B, J, C)     forward
  __nv_bfloat16* dx;          // (B, C, S)     backward
  float* dk;                  // (S, J) fp32   backward, accumulated
  int C, S, J;
  long long tiles;            // B * C / 128
};

// Activation fragment (the "col" operand of m16n8k16): b0 = {X[row0][col0], X[row0 + rstep][col0 + cstep]} ... gathered with
// 2-byte loads from rows of S bf16 (2-byte aligned only).
__device__ __forceinline__ uint32_t gather2(uint32_t a0, uint32_t a1) { return lds16(a0) | (lds16(a1) << 16); }

// ---------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------
template <int KSTEPS>
__global__ void __launch_bounds__(kThreads) spatial_project_fwd_kernel(const Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tq = lane & 3, l15 = lane & 15, lhi = lane >> 4;
  const int S = p.S, J = p.J, C = p.C;
  const uint32_t tile_bytes = (uint32_t)kTileC * S * 2;
  const uint32_t buf_stride = (tile_bytes + 32 + 127) / 128 * 128;
  unsigned char* ks_g = smem;                                   // K^T [16][kKP]
  unsigned char* out_g = ks_g + 16 * kKP;                       // OUT [16][kOP]
  unsigned char* xb_g = out_g + 16 * kOP;                       // 2 x (tile + 32 bytes of zeros)
  const uint32_t ks = smem_u32(ks_g), out = smem_u32(out_g), xb = smem_u32(xb_g), bar0 = smem_u32(&bars[0]);

  for (int e = tid; e < 16 * 64; e += kThreads) {
    const int j = e >> 6, s = e & 63;
    const __nv_bfloat16 v = (j < J && s < S) ? p.k[s * J + j] : __float2bfloat16_rn(0.f);
    *reinterpret_cast<__nv_bfloat16*>(ks_g + j * kKP + s * 2) = v;
  }
  for (int e = tid; e < 2 * 8; e += kThreads)                   // the 32 bytes behind each landing buffer: finite zeros
    *reinterpret_cast<uint32_t*>(xb_g + (e >> 3) * buf_stride + tile_bytes + (e & 7) * 4) = 0u;
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_barrier_init();
  }
  __syncthreads();
  uint32_t a[KSTEPS][4];
#pragma unroll
  for (int kk = 0; kk < KSTEPS; ++kk) ldsm4(a[kk], ks + l15 * kKP + (kk * 16 + lhi * 8) * 2);

  const long long first = blockIdx.x, step = gridDim.x;
  if (tid == 0 && first < p.tiles) {
    mbar_expect_tx(bar0, tile_bytes);
    bulk_g2s(xb, p.x + (size_t)first * kTileC * S, tile_bytes, bar0);
  }
  int it = 0;
  for (long long tile = first; tile < p.tiles; tile += step, ++it) {
    const int slot = it & 1;
    const uint32_t parity = (it >> 1) & 1;
    if (tid == 0 && tile + step < p.tiles) {                    // the other buffer was released by the barrier that ended it - 1
      mbar_expect_tx(bar0 + 8 * (slot ^ 1), tile_bytes);
      bulk_g2s(xb + (slot ^ 1) * buf_stride, p.x + (size_t)(tile + step) * kTileC * S, tile_bytes, bar0 + 8 * (slot ^ 1));
    }
    mbar_wait(bar0 + 8 * slot, parity);
    const uint32_t xt = xb + slot * buf_stride;
    // warp w: channels 32 w .. 32 w + 31 of the tile = 4 n-tiles
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int cl = warp * 32 + nt * 8;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t row = xt + (uint32_t)(cl + gid) * S * 2 + tq * 4;
#pragma unroll
      for (int kk = 0; kk < KSTEPS; ++kk) {
        const uint32_t a0 = row + kk * 32;
        const uint32_t b0 = gather2(a0, a0 + 2), b1 = gather2(a0 + 16, a0 + 18);
        mma16816(acc, a[kk], b0, b1);
      }
      sts32(out + gid * kOP + (cl + 2 * tq) * 2, pack2(acc[0], acc[1]));
      sts32(out + (gid + 8) * kOP + (cl + 2 * tq) * 2, pack2(acc[2], acc[3]));
    }
    __syncthreads();
    {
      const long long row0 = tile * kTileC;
      const long long b = row0 / C;
      const int c0 = (int)(row0 % C);
      __nv_bfloat16* tb = p.t + ((size_t)b * J) * C + c0;
      for (int e = tid; e < J * 16; e += kThreads) {
        const int j = e >> 4, q = e & 15;
        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(tb + (size_t)j * C) + q * 16) = lds128(out + j * kOP + q * 16);
      }
    }
    __syncthreads();                                            // OUT and this landing buffer are free again
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------------
template <int KSTEPS>                                            // ceil(S / 16): m16 tiles of s in dX, and 2 KSTEPS n8 tiles in dK
__global__ void __launch_bounds__(kThreads) spatial_project_bwd_kernel(const Params p) {
  constexpr int NTS = 2 * KSTEPS;                               // n8 tiles over s in the dK product
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tq = lane & 3, l15 = lane & 15, lhi = lane >> 4;
  const int S = p.S, J = p.J, C = p.C;
  const uint32_t tile_bytes = (uint32_t)kTileC * S * 2;
  const uint32_t buf_stride = (tile_bytes + 32 + 127) / 128 * 128;
  unsigned char* k2_g = smem;                                   // K [64][kK2P]  (rows s, cols j)
  unsigned char* gb_g = k2_g + 64 * kK2P;                       // 2 x gT tile [16][kOP]
  float* red_g = reinterpret_cast<float*>(gb_g + 2 * 16 * kOP); // dK^T partial [16][64] fp32
  unsigned char* dxs_g = reinterpret_cast<unsigned char*>(red_g + 16 * 64);   // dX staging, linear [128][S]
  unsigned char* xb_g = dxs_g + buf_stride;                     // 2 x X tile
  const uint32_t k2 = smem_u32(k2_g), gb = smem_u32(gb_g), dxs = smem_u32(dxs_g), xb = smem_u32(xb_g), bar0 = smem_u32(&bars[0]);

  for (int e = tid; e < 64 * 16; e += kThreads) {
    const int s = e >> 4, j = e & 15;
    const __nv_bfloat16 v = (j < J && s < S) ? p.k[s * J + j] : __float2bfloat16_rn(0.f);
    *reinterpret_cast<__nv_bfloat16*>(k2_g + s * kK2P + j * 2) = v;
  }
  for (int e = tid; e < 2 * 16 * kOP / 4; e += kThreads) reinterpret_cast<uint32_t*>(gb_g)[e] = 0u;     // rows j >= J stay zero
  for (int e = tid; e < 16 * 64; e += kThreads) red_g[e] = 0.f;
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  __syncthreads();
  uint32_t a2[KSTEPS][4];                                       // A[m = s][k = j]
#pragma unroll
  for (int mt = 0; mt < KSTEPS; ++mt) ldsm4(a2[mt], k2 + (mt * 16 + l15) * kK2P + lhi * 16);
  float dk[NTS][4];
#pragma unroll
  for (int n = 0; n < NTS; ++n)
#pragma unroll
    for (int i = 0; i < 4; ++i) dk[n][i] = 0.f;

  const long long first = blockIdx.x, step = gridDim.x;
  const uint32_t load_bytes = tile_bytes + (uint32_t)J * kTileC * 2;
  auto issue = [&](long long tile, int slot) {
    const long long row0 = tile * kTileC;
    const long long b = row0 / C;
    const int c0 = (int)(row0 % C);
    const uint32_t bar = bar0 + 8 * slot;
    mbar_expect_tx(bar, load_bytes);
    bulk_g2s(xb + slot * buf_stride, p.x + (size_t)row0 * S, tile_bytes, bar);
    for (int j = 0; j < J; ++j)
      bulk_g2s(gb + slot * 16 * kOP + j * kOP, p.gt + ((size_t)b * J + j) * C + c0, kTileC * 2, bar);
  };
  if (tid == 0 && first < p.tiles) issue(first, 0);
  int it = 0;
  for (long long tile = first; tile < p.tiles; tile += step, ++it) {
    const int slot = it & 1;
    const uint32_t parity = (it >> 1) & 1;
    if (tid == 0 && tile + step < p.tiles) issue(tile + step, slot ^ 1);
    mbar_wait(bar0 + 8 * slot, parity);
    const uint32_t xt = xb + slot * buf_stride, gt = gb + slot * 16 * kOP;
    if (tid == 0) bulk_wait_read0();                            // the previous tile's bulk store has finished reading DXS
    __syncthreads();

    // ---- dX^T[s][c] = K[s][j] gT[j][c]: warp w owns channels 32 w .. 32 w + 31 (two n16 pairs) ----
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      const int cl = warp * 32 + np * 16;
      uint32_t bf[4];
      ldsm4t(bf, gt + l15 * kOP + (cl + lhi * 8) * 2);
#pragma unroll
      for (int mt = 0; mt < KSTEPS; ++mt) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          mma16816(acc, a2[mt], bf[2 * h], bf[2 * h + 1]);
          const int c = cl + h * 8 + 2 * tq;
          const int s0 = mt * 16 + gid;
          if (s0 < S) {
            sts16(dxs + (uint32_t)(c * S + s0) * 2, bf16_bits(acc[0]));
            sts16(dxs + (uint32_t)((c + 1) * S + s0) * 2, bf16_bits(acc[1]));
          }
          if (s0 + 8 < S) {
            sts16(dxs + (uint32_t)(c * S + s0 + 8) * 2, bf16_bits(acc[2]));
            sts16(dxs + (uint32_t)((c + 1) * S + s0 + 8) * 2, bf16_bits(acc[3]));
          }
        }
      }
    }
    // ---- dK^T[j][s] += gT[j][c] X[c][s]: warp w owns the k-steps (channels) 32 w .. 32 w + 31 ----
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const int cl = warp * 32 + kk * 16;
      uint32_t ag[4];
      ldsm4(ag, gt + l15 * kOP + (cl + lhi * 8) * 2);
      const uint32_t r0 = xt + (uint32_t)(cl + 2 * tq) * S * 2 + gid * 2;
#pragma unroll
      for (int n = 0; n < NTS; ++n) {
        const uint32_t q = r0 + n * 16;
        const uint32_t b0 = gather2(q, q + S * 2), b1 = gather2(q + 8 * S * 2, q + 9 * S * 2);
        mma16816(dk[n], ag, b0, b1);
      }
    }
    fence_proxy_async();                                        // DXS was written through the generic proxy
    __syncthreads();
    if (tid == 0) {
      bulk_s2g(p.dx + (size_t)tile * kTileC * S, dxs, tile_bytes);
      bulk_commit();
    }
    // the landing buffers of this slot are free once every warp passed the barrier above
  }
  if (tid == 0) bulk_wait0();
  // ---- dK: reduce the 4 warps in shared memory, then one atomic per entry per CTA ----
#pragma unroll
  for (int n = 0; n < NTS; ++n) {
    const int s = n * 8 + 2 * tq;
    atomicAdd(&red_g[gid * 64 + s], dk[n][0]);
    atomicAdd(&red_g[gid * 64 + s + 1], dk[n][1]);
    atomicAdd(&red_g[(gid + 8) * 64 + s], dk[n][2]);
    atomicAdd(&red_g[(gid + 8) * 64 + s + 1], dk[n][3]);
  }
  __syncthreads();
  if (first < p.tiles)
    for (int e = tid; e < S * J; e += kThreads) {
      const int s = e / J, j = e % J;
      atomicAdd(&p.dk[e], red_g[j * 64 + s]);
    }
}

static int check_shape(int64_t batch, int64_t channels, int32_t S, int32_t J) {
  TLT_REQUIRE(batch >= 1 && channels >= kTileC && channels % kTileC == 0, "tlt_spatial_project: channels must be a multiple of 128");
  TLT_REQUIRE(S >= 1 && S <= 64 && J >= 1 && J <= 16, "tlt_spatial_project: needs prod(spatial) <= 64 and prod(ranks) <= 16");
  return TLT_OK;
}

template <class Kern>
static int grid_for(Kern kern, size_t smem, long long tiles, int* grid) {
  int dev = 0, sms = 0, bps = 0;
  TLT_ENSURE_SMEM(kern, smem);
  TLT_CHECK_CUDA(cudaGetDevice(&dev));
  TLT_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  TLT_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, kThreads, smem));
  if (bps < 1) bps = 1;
  if (bps > 6) bps = 6;
  long long g = (long long)sms * bps;
  if (g > tiles) g = tiles;
  *grid = (int)g;
  return TLT_OK;
}

}  // namespace sp
}  // namespace tlt

using namespace tlt;

extern "C" int tlt_spatial_project_supported(int64_t batch, int64_t channels, int32_t S, int32_t J) {
  return batch >= 1 && channels >= sp::kTileC && channels % sp::kTileC == 0 && S >= 1 && S <= 64 && J >= 1 && J <= 16 ? 1 : 0;
}

extern "C" int tlt_spatial_project_fwd(int64_t batch, int64_t channels, int32_t S, int32_t J, const void* x, const void* k, void* t,
                                       void* stream) {
  int rc = sp::check_shape(batch, channels, S, J);
  if (rc) return rc;
  TLT_REQUIRE(x && k && t, "tlt_spatial_project_fwd: null argument");
  sp::Params p{};
  p.x = (const __nv_bfloat16*)x; p.k = (const __nv_bfloat16*)k; p.t = (__nv_bfloat16*)t;
  p.C = (int)channels; p.S = S; p.J = J;
  p.tiles = batch * channels / sp::kTileC;
  const size_t tile = (size_t)sp::kTileC * S * 2, stride = (tile + 32 + 127) / 128 * 128;
  const size_t smem = 16 * sp::kKP + 16 * sp::kOP + 2 * stride;
  const int ksteps = (S + 15) / 16;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = 0;
#define TLT_SP_FWD(KS)                                                                            \
  case KS:                                                                                        \
    rc = sp::grid_for(sp::spatial_project_fwd_kernel<KS>, smem, p.tiles, &grid);                  \
    if (rc) return rc;                                                                            \
    sp::spatial_project_fwd_kernel<KS><<<grid, sp::kThreads, smem, st>>>(p);                      \
    break;
  switch (ksteps) { TLT_SP_FWD(1) TLT_SP_FWD(2) TLT_SP_FWD(3) TLT_SP_FWD(4) }
#undef TLT_SP_FWD
  return check_launch("spatial_project_fwd_kernel");
}

extern "C" int tlt_spatial_project_bwd(int64_t batch, int64_t channels, int32_t S, int32_t J, const void* x, const void* k,
                                       const void* gt, void* dx, float* dk, void* stream) {
  int rc = sp::check_shape(batch, channels, S, J);
  if (rc) return rc;
  TLT_REQUIRE(x && k && gt && dx && dk, "tlt_spatial_project_bwd: null argument");
  sp::Params p{};
  p.x = (const __nv_bfloat16*)x; p.k = (const __nv_bfloat16*)k; p.gt = (const __nv_bfloat16*)gt;
  p.dx = (__nv_bfloat16*)dx; p.dk = dk;
  p.C = (int)channels; p.S = S; p.J = J;
  p.tiles = batch * channels / sp::kTileC;
  const size_t tile = (size_t)sp::kTileC * S * 2, stride = (tile + 32 + 127) / 128 * 128;
  const size_t smem = 64 * sp::kK2P + 2 * 16 * sp::kOP + 16 * 64 * 4 + 3 * stride;
  const int ksteps = (S + 15) / 16;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = 0;
#define TLT_SP_BWD(KS)                                                                            \
  case KS:                                                                                        \
    rc = sp::grid_for(sp::spatial_project_bwd_kernel<KS>, smem, p.tiles, &grid);                  \
    if (rc) return rc;                                                                            \
    sp::spatial_project_bwd_kernel<KS><<<grid, sp::kThreads, smem, st>>>(p);                      \
    break;
  switch (ksteps) { TLT_SP_BWD(1) TLT_SP_BWD(2) TLT_SP_BWD(3) TLT_SP_BWD(4) }
#undef TLT_SP_BWD
  return check_launch("spatial_project_bwd_kernel");
}
