// Depthwise 3x3 (stride 1, pad 1) on NCHW bf16 planes, fed by TMA bulk copies -- the middle stage of the CP convolution on
// the large maps of BASELINE config 3 (56 x 56 and 28 x 28, R = 69 .. 141 channels, batch 256: 111 MB in, 111 MB out).
//
// The previous fast path (dwconv.cu: dwconv3x3_fast) staged each group of planes with ld.global + convert-to-fp32 +
// st.shared, synchronised, and only then computed: ~136 us per pass at 56 x 56 (4x the HBM time), with the load latency
// exposed once per block.  Here
//   * a persistent block always has its NEXT group of planes in flight: consecutive (image, channel) planes are one
//     contiguous span of NCHW memory, so one thread issues ONE cp.async.bulk per group into the other half of a double
//     buffer and an mbarrier transaction count signals arrival -- no thread touches the data on its way in;
//   * the planes stay bf16 in shared memory: a thread owns a strip of 4 consecutive outputs and reads, per input row, one
//     aligned 8-byte word (the 4 centre inputs) and two 2-byte halo elements; rows / columns outside the plane are
//     predicated to zero instead of being stored as padding;
//   * weight gradient: the same landing scheme over the planes of ONE channel across images (one bulk copy per plane for
//     x and for dy), nine tap sums per thread in registers, one atomicAdd per (channel, tap) per block.
// Requirements (otherwise the callers stay on dwconv.cu): bf16, W % 4 == 0, (H W) % 8 == 0, one plane <= 16 KB.
// Reference call sites: as dwconv.cu (tltorch/functional/convolution.py:268-271, 321-332 and their autograd).
#include "tc_common.cuh"

namespace tlt {
namespace dwt {

struct Params {
  int H, W, strips;                 // strips per row = W / SW (SW = 8 when W % 8 == 0, else 4)
  int channels, PB;                 // planes per group
  long long planes;                 // batch * channels
  int spp;                          // strips per plane
  uint32_t m_spp, m_strips;         // magic multipliers: floor(n / d) = umulhi(n, m) for n < 2^31 / ... (d > 1)
};

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ int fdiv(int n, uint32_t m, int d) { return d > 1 ? (int)__umulhi((uint32_t)n, m) : n; }

// SW + 2 inputs (columns SW sg - 1 .. SW sg + SW) of row `r` of a bf16 plane in shared memory, zero outside the plane
template <int SW>
__device__ __forceinline__ void load_row(const __nv_bfloat16* plane, int r, int sg, const Params& q, float (&in)[SW + 2]) {
  if (r < 0 || r >= q.H) {
#pragma unroll
    for (int i = 0; i < SW + 2; ++i) in[i] = 0.f;
    return;
  }
  const __nv_bfloat16* rp = plane + r * q.W + SW * sg;
  in[0] = sg > 0 ? __bfloat162float(rp[-1]) : 0.f;
#pragma unroll
  for (int h = 0; h < SW / 4; ++h) {
    const uint2 mid = *reinterpret_cast<const uint2*>(rp + 4 * h);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&mid.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&mid.y));
    in[1 + 4 * h] = a.x; in[2 + 4 * h] = a.y; in[3 + 4 * h] = b.x; in[4 + 4 * h] = b.y;
  }
  in[SW + 1] = sg < q.strips - 1 ? __bfloat162float(rp[SW]) : 0.f;
}

// FLIP = false: y = corr(x, w);  true: dx = corr(dy, reversed w)  (the adjoint at stride 1, pad 1)
template <bool FLIP, int SW>
__global__ void __launch_bounds__(256) dwconv3x3_tma_kernel(const Params q, const __nv_bfloat16* __restrict__ x,
                                                            const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ y,
                                                            long long n_groups) {
  extern __shared__ __align__(128) unsigned char dwt_smem[];
  const int psz = q.H * q.W;
  const uint32_t buf_bytes = (uint32_t)q.PB * psz * 2;
  __nv_bfloat16* bufs[2] = {reinterpret_cast<__nv_bfloat16*>(dwt_smem), reinterpret_cast<__nv_bfloat16*>(dwt_smem + buf_bytes)};
  float* ws = reinterpret_cast<float*>(dwt_smem + 2 * buf_bytes);                 // [2][PB][9]
  const uint32_t bar0 = smem_u32(dwt_smem + 2 * buf_bytes + 2 * q.PB * 9 * 4);    // two mbarriers (8 bytes each)
  auto issue = [&](long long grp, int slot) {
    const long long p0 = grp * q.PB;
    const int npl = (int)min((long long)q.PB, q.planes - p0);
    const uint32_t bytes = (uint32_t)npl * psz * 2;
    mbar_expect_tx(bar0 + 8 * slot, bytes);
    bulk_g2s(smem_u32(bufs[slot]), x + p0 * psz, bytes, bar0 + 8 * slot);
  };
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_barrier_init();
    if ((long long)blockIdx.x < n_groups) issue(blockIdx.x, 0);
  }
  __syncthreads();
  uint32_t phase[2] = {0, 0};
  int slot = 0;
  for (long long grp = blockIdx.x; grp < n_groups; grp += gridDim.x, slot ^= 1) {
    const long long p0 = grp * q.PB;
    const int npl = (int)min((long long)q.PB, q.planes - p0);
    if (threadIdx.x == 0 && grp + gridDim.x < n_groups) issue(grp + gridDim.x, slot ^ 1);     // its readers finished last iteration
    float* wsl = ws + slot * q.PB * 9;
    for (int e = threadIdx.x; e < npl * 9; e += blockDim.x) {
      const int j = e / 9, t = e - j * 9;
      const int c = (int)((p0 + j) % q.channels);
      wsl[j * 9 + (FLIP ? 8 - t : t)] = __bfloat162float(w[(long long)c * 9 + t]);
    }
    mbar_wait(bar0 + 8 * slot, phase[slot]);
    phase[slot] ^= 1;
    __syncthreads();                                                 // weights staged
    const __nv_bfloat16* buf = bufs[slot];
    const int total = npl * q.spp;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
      const int j = fdiv(e, q.m_spp, q.spp), r = e - j * q.spp;
      const int row = fdiv(r, q.m_strips, q.strips), sg = r - row * q.strips;
      const __nv_bfloat16* plane = buf + j * psz;
      const float* wp = wsl + j * 9;
      float wk[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) wk[t] = wp[t];
      float acc[SW];
#pragma unroll
      for (int o = 0; o < SW; ++o) acc[o] = 0.f;
#pragma unroll
      for (int t1 = 0; t1 < 3; ++t1) {
        float in[SW + 2];
        load_row<SW>(plane, row + t1 - 1, sg, q, in);
#pragma unroll
        for (int t2 = 0; t2 < 3; ++t2)
#pragma unroll
          for (int o = 0; o < SW; ++o) acc[o] = fmaf(in[o + t2], wk[t1 * 3 + t2], acc[o]);
      }
      uint32_t packed[SW / 2];
#pragma unroll
      for (int o = 0; o < SW / 2; ++o) {
        __nv_bfloat162 a = __floats2bfloat162_rn(acc[2 * o], acc[2 * o + 1]);
        packed[o] = *reinterpret_cast<uint32_t*>(&a);
      }
      __nv_bfloat16* dptr = y + (p0 + j) * psz + row * q.W + SW * sg;
      if (SW == 8) *reinterpret_cast<uint4*>(dptr) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
      else *reinterpret_cast<uint2*>(dptr) = make_uint2(packed[0], packed[1]);
    }
    __syncthreads();                                                 // everyone is done with this slot (and its weights)
  }
}

// dw[c, t1, t2] += sum_{b, h, w} dy[b,c,h,w] * x[b,c,h + t1 - 1, w + t2 - 1];  grid = (channels, batch chunks)
template <int SW>
__global__ void __launch_bounds__(256) dwconv3x3_wgrad_tma_kernel(const Params q, int batch, int batch_per_block,
                                                                  const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                                  float* __restrict__ dw) {
  extern __shared__ __align__(128) unsigned char dwt_smem[];
  __shared__ float red[9 * 8];
  const int psz = q.H * q.W;
  const uint32_t half = (uint32_t)q.PB * psz * 2;                   // one operand of one slot
  // layout: slot 0: xs | gs, slot 1: xs | gs, then two mbarriers
  auto xs_of = [&](int slot) { return reinterpret_cast<__nv_bfloat16*>(dwt_smem + (size_t)slot * 2 * half); };
  auto gs_of = [&](int slot) { return reinterpret_cast<__nv_bfloat16*>(dwt_smem + (size_t)slot * 2 * half + half); };
  const uint32_t bar0 = smem_u32(dwt_smem + 4 * (size_t)half);
  const int c = blockIdx.x;
  const int b_begin = blockIdx.y * batch_per_block;
  const int b_end = min(batch, b_begin + batch_per_block);
  auto issue = [&](int b0, int slot) {
    const int npl = min(q.PB, b_end - b0);
    mbar_expect_tx(bar0 + 8 * slot, (uint32_t)npl * psz * 2 * 2);
    for (int j = 0; j < npl; ++j) {
      const long long off = ((long long)(b0 + j) * q.channels + c) * psz;
      bulk_g2s(smem_u32(xs_of(slot) + j * psz), x + off, (uint32_t)psz * 2, bar0 + 8 * slot);
      bulk_g2s(smem_u32(gs_of(slot) + j * psz), dy + off, (uint32_t)psz * 2, bar0 + 8 * slot);
    }
  };
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_barrier_init();
    if (b_begin < b_end) issue(b_begin, 0);
  }
  __syncthreads();
  float acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = 0.f;
  uint32_t phase[2] = {0, 0};
  int slot = 0;
  for (int b0 = b_begin; b0 < b_end; b0 += q.PB, slot ^= 1) {
    const int npl = min(q.PB, b_end - b0);
    if (threadIdx.x == 0 && b0 + q.PB < b_end) issue(b0 + q.PB, slot ^ 1);
    mbar_wait(bar0 + 8 * slot, phase[slot]);
    phase[slot] ^= 1;
    const __nv_bfloat16* xs = xs_of(slot);
    const __nv_bfloat16* gs = gs_of(slot);
    const int total = npl * q.spp;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
      const int j = fdiv(e, q.m_spp, q.spp), r = e - j * q.spp;
      const int row = fdiv(r, q.m_strips, q.strips), sg = r - row * q.strips;
      float g[SW];
#pragma unroll
      for (int h = 0; h < SW / 4; ++h) {
        const uint2 g2 = *reinterpret_cast<const uint2*>(gs + j * psz + row * q.W + SW * sg + 4 * h);
        const float2 ga = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&g2.x));
        const float2 gb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&g2.y));
        g[4 * h] = ga.x; g[4 * h + 1] = ga.y; g[4 * h + 2] = gb.x; g[4 * h + 3] = gb.y;
      }
      const __nv_bfloat16* plane = xs + j * psz;
#pragma unroll
      for (int t1 = 0; t1 < 3; ++t1) {
        float in[SW + 2];
        load_row<SW>(plane, row + t1 - 1, sg, q, in);
#pragma unroll
        for (int t2 = 0; t2 < 3; ++t2)
#pragma unroll
          for (int o = 0; o < SW; ++o) acc[t1 * 3 + t2] = fmaf(in[o + t2], g[o], acc[t1 * 3 + t2]);
      }
    }
    __syncthreads();
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    float v = acc[t];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if (lane == 0) red[t * 8 + warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    float v = 0.f;
    for (int wv = 0; wv < 8; ++wv) v += red[threadIdx.x * 8 + wv];
    atomicAdd(dw + (long long)c * 9 + threadIdx.x, v);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Tensor-core version of the forward / data-gradient kernel.  The FFMA kernel above is issue-bound (72 % issue-active, 59 M warp
// instructions per pass at 56 x 56: 9 FMAs + unpack per output).  A depthwise 3x3 over a plane is
//      Y = sum_dh  shift_dh(Z) . T_dh          T_dh[w'][w] = tap[dh][w' - w + 1]  (banded Toeplitz, W x W)
// so per 16-row slab of a plane the vertical taps are three row-shifted ldmatrix reads of the SAME shared-memory plane (a
// one-row shift keeps the 16-byte alignment because W % 8 == 0; rows outside the plane read a zero row) and the horizontal
// taps live in the B operand: for the 8-column output tile j only the k16 step that contains it and ONE neighbouring step (for
// the column just left / right of the tile) are non-zero, i.e. 2 m16n8k16 HMMAs per tile per dh.  B fragments are five register
// patterns per dh built from the channel's nine taps.  ~0.2 warp instructions per output instead of ~1.1.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm4_(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816_(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma1688_(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ uint32_t bf16_bits_(float v) { return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v)); }

template <bool FLIP, int W>
__global__ void __launch_bounds__(256) dwconv3x3_mma_kernel(const Params q, const __nv_bfloat16* __restrict__ x,
                                                            const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ y,
                                                            long long n_groups) {
  // W % 8 != 0 (28-wide maps): the 56-byte rows of the landed planes break ldmatrix's 16-byte row alignment, so each group is
  // re-pitched once inside shared memory (8-byte moves) into rows of ROWB bytes -- an odd multiple of 16, zero beyond column W
  constexpr bool REPITCH = W % 8 != 0;
  constexpr int NKS = (W + 15) / 16, NNT = (W + 7) / 8;
  constexpr int ROWB = REPITCH ? ((((W + 7) / 8) * 16 / 16) % 2 == 0 ? ((W + 7) / 8) * 16 + 16 : ((W + 7) / 8) * 16) : W * 2;
  static_assert(W % 4 == 0 && W <= 64, "rows of whole 8-byte chunks, at most four k16 steps");
  static_assert(NKS * 32 + 16 <= 128 + 16, "zero row covers every k step");
  extern __shared__ __align__(128) unsigned char dwt_smem[];
  const int psz = q.H * W, H = q.H;
  const uint32_t buf_bytes = (uint32_t)q.PB * psz * 2;
  const uint32_t slot_stride = (buf_bytes + 32 + 127) / 128 * 128;               // 32 zero bytes behind each landing buffer
  const uint32_t work_bytes = REPITCH ? ((uint32_t)q.PB * H * ROWB + 32 + 127) / 128 * 128 : 0u;
  unsigned char* zrow_g = dwt_smem;                                              // 128 bytes of zeros
  unsigned char* buf_g = dwt_smem + 128;
  unsigned char* work_g = buf_g + 2 * slot_stride;                               // re-pitched planes of the current group
  float* ws = reinterpret_cast<float*>(work_g + work_bytes);                     // [2][PB][9]
  const uint32_t bar0 = smem_u32(reinterpret_cast<unsigned char*>(ws) + 2 * q.PB * 9 * 4);
  const uint32_t zrow = smem_u32(zrow_g), buf0 = smem_u32(buf_g);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tq = lane & 3;
  auto issue = [&](long long grp, int slot) {
    const long long p0 = grp * q.PB;
    const int npl = (int)min((long long)q.PB, q.planes - p0);
    const uint32_t bytes = (uint32_t)npl * psz * 2;
    mbar_expect_tx(bar0 + 8 * slot, bytes);
    bulk_g2s(buf0 + slot * slot_stride, x + p0 * psz, bytes, bar0 + 8 * slot);
  };
  for (int e = threadIdx.x; e < 32; e += blockDim.x) reinterpret_cast<uint32_t*>(zrow_g)[e] = 0u;
  for (int e = threadIdx.x; e < 16; e += blockDim.x)
    reinterpret_cast<uint32_t*>(buf_g + (e >> 3) * slot_stride + buf_bytes)[e & 7] = 0u;
  if constexpr (REPITCH) {
    const int words = (int)(work_bytes / 4);
    for (int e = threadIdx.x; e < words; e += blockDim.x) reinterpret_cast<uint32_t*>(work_g)[e] = 0u;
  }
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_barrier_init();
    if ((long long)blockIdx.x < n_groups) issue(blockIdx.x, 0);
  }
  __syncthreads();
  const int nmt = (H + 15) / 16;
  const int i0 = 2 * tq - gid + 1;                                   // tap index of the band entry (k = 2 tq, n = gid); +1 for k + 1
  const bool f07 = tq == 0 && gid == 7, f30 = tq == 3 && gid == 0;
  uint32_t phase[2] = {0, 0};
  int slot = 0;
  for (long long grp = blockIdx.x; grp < n_groups; grp += gridDim.x, slot ^= 1) {
    const long long p0 = grp * q.PB;
    const int npl = (int)min((long long)q.PB, q.planes - p0);
    if (threadIdx.x == 0 && grp + gridDim.x < n_groups) issue(grp + gridDim.x, slot ^ 1);
    float* wsl = ws + slot * q.PB * 9;
    for (int e = threadIdx.x; e < npl * 9; e += blockDim.x) {
      const int j = e / 9, t = e - j * 9;
      const int c = (int)((p0 + j) % q.channels);
      wsl[j * 9 + (FLIP ? 8 - t : t)] = __bfloat162float(w[(long long)c * 9 + t]);
    }
    mbar_wait(bar0 + 8 * slot, phase[slot]);
    phase[slot] ^= 1;
    if constexpr (REPITCH) {
      constexpr int CPR = W / 4;                                      // 8-byte chunks per row
      const unsigned char* land = buf_g + slot * slot_stride;
      for (int e = threadIdx.x; e < npl * H * CPR; e += blockDim.x) {
        const int r = e / CPR, c = e - r * CPR;                       // r = (plane, row)
        *reinterpret_cast<uint2*>(work_g + (size_t)r * ROWB + c * 8) = *reinterpret_cast<const uint2*>(land + (size_t)r * (W * 2) + c * 8);
      }
    }
    __syncthreads();                                                 // weights staged (and planes re-pitched)
    for (int task = warp; task < npl * nmt; task += 8) {
      const int j = task / nmt, mt = task - j * nmt;
      const uint32_t plane = REPITCH ? smem_u32(work_g) + (uint32_t)j * H * ROWB : buf0 + slot * slot_stride + (uint32_t)j * psz * 2;
      float acc[NNT][4];
#pragma unroll
      for (int n = 0; n < NNT; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[n][i] = 0.f;
#pragma unroll
      for (int dh = 0; dh < 3; ++dh) {
        // B fragments of T_dh: one band fragment (the same for an even tile's first half and an odd tile's second half) and
        // two single-lane corner entries -- a handful of selects per dh
        const uint32_t tb0 = bf16_bits_(wsl[j * 9 + dh * 3]), tb1 = bf16_bits_(wsl[j * 9 + dh * 3 + 1]), tb2 = bf16_bits_(wsl[j * 9 + dh * 3 + 2]);
        const uint32_t lo = i0 == 0 ? tb0 : (i0 == 1 ? tb1 : (i0 == 2 ? tb2 : 0u));
        const uint32_t hi = i0 == -1 ? tb0 : (i0 == 0 ? tb1 : (i0 == 1 ? tb2 : 0u));
        const uint32_t band = lo | (hi << 16);
        const uint32_t c07 = f07 ? tb2 : 0u, c30 = f30 ? (tb0 << 16) : 0u;
        const uint32_t pe[2] = {band, c07}, plast[2] = {band, 0u}, po[2] = {c30, band}, pl[2] = {0u, c30}, pr[2] = {c07, 0u};
        const int row = mt * 16 + (lane & 15) + dh - 1;
        const uint32_t ra = (row >= 0 && row < H ? plane + (uint32_t)row * ROWB : zrow) + (lane >> 4) * 16;
        uint32_t a[NKS][4];
#pragma unroll
        for (int s = 0; s < NKS; ++s) ldsm4_(a[s], ra + s * 32);
#pragma unroll
        for (int n = 0; n < NNT; ++n) {
          const int s = n >> 1;
          if ((n & 1) == 0) {
            if (n == NNT - 1 && !REPITCH) mma16816_(acc[n], a[s], plast[0], plast[1]);
            else mma16816_(acc[n], a[s], pe[0], pe[1]);
            if (n > 0) mma1688_(acc[n], a[s - 1][2], a[s - 1][3], pl[1]);      // column n0 - 1: second k8 half of the step before
          } else {
            mma16816_(acc[n], a[s], po[0], po[1]);
            if (n < NNT - 1) mma1688_(acc[n], a[s + 1][0], a[s + 1][1], pr[0]);  // column n0 + 8: first k8 half of the next step
          }
        }
      }
      const int r0 = mt * 16 + gid, r1 = r0 + 8;
      __nv_bfloat16* yp = y + (p0 + j) * psz + 2 * tq;
#pragma unroll
      for (int n = 0; n < NNT; ++n) {
        const bool colok = !REPITCH || n * 8 + 2 * tq < W;
        if (r0 < H && colok) *reinterpret_cast<__nv_bfloat162*>(yp + r0 * W + n * 8) = __floats2bfloat162_rn(acc[n][0], acc[n][1]);
        if (r1 < H && colok) *reinterpret_cast<__nv_bfloat162*>(yp + r1 * W + n * 8) = __floats2bfloat162_rn(acc[n][2], acc[n][3]);
      }
    }
    __syncthreads();                                                 // everyone is done with this slot (and its weights)
  }
}

// Weight gradient on the tensor cores: for one channel plane pair (z, g) and one vertical tap dh
//      D_dh[w'][w] = sum_h z[h + dh - 1][w'] g[h][w]            (m16n8k16, both operands through ldmatrix.trans, k = image row)
// and dw[dh][dw] is the sum of the diagonal w' - w = dw - 1 of D_dh.  Only the tiles that touch the three diagonals are computed
// (per 8-column tile of g one 16-row tile of z plus one neighbour for the column just outside), accumulated in registers over all
// images of the block, and the diagonals are picked out once at the end.  A warp owns one dh; two warps per dh walk alternate planes.
__device__ __forceinline__ void ldsm4t_(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm2t_(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}

template <int W>
__global__ void __launch_bounds__(192) dwconv3x3_wgrad_mma_kernel(const Params q, int batch, int batch_per_block,
                                                                  const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                                  float* __restrict__ dw) {
  constexpr int NMT = (W + 15) / 16, NNT = W / 8, ROWB = W * 2;
  static_assert(W % 8 == 0 && W <= 64, "rows of whole 16-byte chunks");
  extern __shared__ __align__(128) unsigned char dwt_smem[];
  const int psz = q.H * W, H = q.H;
  const uint32_t half = ((uint32_t)q.PB * psz * 2 + 32 + 127) / 128 * 128;       // one operand of one slot (+ slack for over-reads)
  const uint32_t zrow = smem_u32(dwt_smem);                                      // 128 bytes of zeros
  const uint32_t base = zrow + 128;
  auto xs_of = [&](int slot) { return base + (uint32_t)slot * 2 * half; };
  auto gs_of = [&](int slot) { return base + (uint32_t)slot * 2 * half + half; };
  const uint32_t bar0 = base + 4 * half;
  const int c = blockIdx.x;
  const int b_begin = blockIdx.y * batch_per_block;
  const int b_end = min(batch, b_begin + batch_per_block);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tq = lane & 3;
  const int dh = warp % 3, pl = warp / 3;                                        // vertical tap of this warp, plane lane (0 / 1)
  auto issue = [&](int b0, int slot) {
    const int npl = min(q.PB, b_end - b0);
    mbar_expect_tx(bar0 + 8 * slot, (uint32_t)npl * psz * 2 * 2);
    for (int j = 0; j < npl; ++j) {
      const long long off = ((long long)(b0 + j) * q.channels + c) * psz;
      bulk_g2s(xs_of(slot) + j * psz * 2, x + off, (uint32_t)psz * 2, bar0 + 8 * slot);
      bulk_g2s(gs_of(slot) + j * psz * 2, dy + off, (uint32_t)psz * 2, bar0 + 8 * slot);
    }
  };
  for (int e = threadIdx.x; e < 32; e += blockDim.x) reinterpret_cast<uint32_t*>(dwt_smem)[e] = 0u;
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_barrier_init();
    if (b_begin < b_end) issue(b_begin, 0);
  }
  __syncthreads();
  float accm[NNT][4], accc[NNT][4];                                              // main tiles / neighbour (corner) tiles per n tile
#pragma unroll
  for (int n = 0; n < NNT; ++n)
#pragma unroll
    for (int i = 0; i < 4; ++i) { accm[n][i] = 0.f; accc[n][i] = 0.f; }
  const int ak = (lane & 7) + ((lane >> 4) << 3);                               // A operand: lane -> k row of the step, and
  const uint32_t am = ((lane >> 3) & 1) * 16;                                    //            which 8-column half of the m16 tile
  const int nks = (H + 15) / 16;
  uint32_t phase[2] = {0, 0};
  int slot = 0;
  for (int b0 = b_begin; b0 < b_end; b0 += q.PB, slot ^= 1) {
    const int npl = min(q.PB, b_end - b0);
    if (threadIdx.x == 0 && b0 + q.PB < b_end) issue(b0 + q.PB, slot ^ 1);
    mbar_wait(bar0 + 8 * slot, phase[slot]);
    phase[slot] ^= 1;
    for (int j = pl; j < npl; j += 2) {
      const uint32_t zp = xs_of(slot) + (uint32_t)j * psz * 2, gp = gs_of(slot) + (uint32_t)j * psz * 2;
      for (int ks = 0; ks < nks; ++ks) {
        const int hb = ks * 16 + (lane & 15);                                    // B operand row (image row of g)
        const uint32_t gaddr = hb < H ? gp + (uint32_t)hb * ROWB : zrow;
        const int ha = ks * 16 + ak + dh - 1;                                    // A operand row (image row of z, shifted by the tap)
        const uint32_t zaddr = (ha >= 0 && ha < H ? zp + (uint32_t)ha * ROWB : zrow) + am;
        uint32_t a[NMT][4];
#pragma unroll
        for (int i = 0; i < NMT; ++i) ldsm4t_(a[i], zaddr + i * 32);
#pragma unroll
        for (int n = 0; n < NNT; ++n) {
          uint32_t b0r, b1r;
          ldsm2t_(b0r, b1r, gaddr + n * 16);
          const int i = n >> 1;
          mma16816_(accm[n], a[i], b0r, b1r);
          if ((n & 1) == 0) { if (n > 0) mma16816_(accc[n], a[i - 1], b0r, b1r); }
          else if (i + 1 < NMT) mma16816_(accc[n], a[i + 1], b0r, b1r);
        }
      }
    }
    __syncthreads();
  }
  // diagonals w' - w = dw - 1 of the accumulated tiles
  float sum[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int n = 0; n < NNT; ++n) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int ml = gid + (e >> 1) * 8, nl = 2 * tq + (e & 1);
      const int wn = n * 8 + nl;
      {
        const int wm = (n >> 1) * 16 + ml, d = wm - wn + 1;
        if (d >= 0 && d <= 2 && wm < W) sum[d] += accm[n][e];
      }
      {
        const int ic = (n & 1) == 0 ? (n >> 1) - 1 : (n >> 1) + 1;
        const int wm = ic * 16 + ml, d = wm - wn + 1;
        if (ic >= 0 && ic < NMT && d >= 0 && d <= 2 && wm < W) sum[d] += accc[n][e];
      }
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float v = sum[d];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if (lane == 0) atomicAdd(dw + (long long)c * 9 + dh * 3 + d, v);
  }
}

static uint32_t magic(int d) { return d > 1 ? (uint32_t)((1ull << 32) / (uint32_t)d) + 1u : 0u; }

}  // namespace dwt

// Entry points used by dwconv.cu's dispatch.  Return TLT_ERR_UNSUPPORTED when the shape does not qualify.
bool dwconv3x3_tma_ok(int H, int W, const void* a, const void* b) {
  return W % 4 == 0 && (H * W) % 8 == 0 && H * W * 2 <= 16 * 1024 && (((uintptr_t)a) & 15) == 0 && (((uintptr_t)b) & 15) == 0;
}

template <int W>
static int dwconv3x3_mma_launch(bool flip, int H, int channels, long long planes, const void* x, const void* w, void* y, cudaStream_t st) {
  dwt::Params q{};
  q.H = H; q.W = W; q.channels = channels; q.planes = planes;
  const int psz = H * W;
  constexpr bool REPITCH = W % 8 != 0;
  constexpr int ROWB = REPITCH ? ((((W + 7) / 8)) % 2 == 0 ? ((W + 7) / 8) * 16 + 16 : ((W + 7) / 8) * 16) : W * 2;
  int pb = (32 * 1024) / (psz * 2);                                  // planes per landing buffer
  if (pb > (REPITCH ? 8 : 4)) pb = REPITCH ? 8 : 4;
  if (pb < 1) pb = 1;
  if ((long long)pb > planes) pb = (int)planes;
  q.PB = pb;
  const long long groups = (planes + pb - 1) / pb;
  const size_t slot = ((size_t)pb * psz * 2 + 32 + 127) / 128 * 128;
  const size_t work = REPITCH ? ((size_t)pb * H * ROWB + 32 + 127) / 128 * 128 : 0;
  const size_t smem = 128 + 2 * slot + work + (size_t)2 * pb * 9 * 4 + 16;
  TLT_ENSURE_SMEM((dwt::dwconv3x3_mma_kernel<true, W>), smem);
  TLT_ENSURE_SMEM((dwt::dwconv3x3_mma_kernel<false, W>), smem);
  const long long cap = 148LL * 4;
  const unsigned blocks = (unsigned)(groups < cap ? groups : cap);
  const __nv_bfloat16 *xp = (const __nv_bfloat16*)x, *wp = (const __nv_bfloat16*)w;
  if (flip) dwt::dwconv3x3_mma_kernel<true, W><<<blocks, 256, smem, st>>>(q, xp, wp, (__nv_bfloat16*)y, groups);
  else dwt::dwconv3x3_mma_kernel<false, W><<<blocks, 256, smem, st>>>(q, xp, wp, (__nv_bfloat16*)y, groups);
  return check_launch("dwconv3x3_mma_kernel");
}

int dwconv3x3_tma(bool flip, int H, int W, int channels, long long planes, const void* x, const void* w, void* y, cudaStream_t st) {
  {
    const char* env = getenv("TLT_DW_MMA");
    const bool off = env && env[0] == 'o' && env[1] == 'f';
    if (!off && H <= 64 && H * W * 2 <= 16 * 1024) {
      if (W == 56) return dwconv3x3_mma_launch<56>(flip, H, channels, planes, x, w, y, st);
      if (W == 64) return dwconv3x3_mma_launch<64>(flip, H, channels, planes, x, w, y, st);
      if (W == 32) return dwconv3x3_mma_launch<32>(flip, H, channels, planes, x, w, y, st);
      if (W == 16) return dwconv3x3_mma_launch<16>(flip, H, channels, planes, x, w, y, st);
      if (W == 28) return dwconv3x3_mma_launch<28>(flip, H, channels, planes, x, w, y, st);
    }
  }
  dwt::Params q;
  const int SWd = W % 8 == 0 ? 8 : 4;
  q.H = H; q.W = W; q.strips = W / SWd; q.channels = channels; q.planes = planes;
  q.spp = H * q.strips;
  q.m_spp = dwt::magic(q.spp); q.m_strips = dwt::magic(q.strips);
  const int psz = H * W;
  // planes per group: ~16 KB per slot so that 3 blocks (2 slots each) fit one SM, at least 1 plane
  int pb = (16 * 1024) / (psz * 2);
  if (pb < 1) pb = 1;
  if ((long long)pb > planes) pb = (int)planes;
  q.PB = pb;
  const long long groups = (planes + pb - 1) / pb;
  const size_t smem = (size_t)2 * pb * psz * 2 + (size_t)2 * pb * 9 * 4 + 16;
  TLT_ENSURE_SMEM((dwt::dwconv3x3_tma_kernel<true, 8>), smem);
  TLT_ENSURE_SMEM((dwt::dwconv3x3_tma_kernel<false, 8>), smem);
  TLT_ENSURE_SMEM((dwt::dwconv3x3_tma_kernel<true, 4>), smem);
  TLT_ENSURE_SMEM((dwt::dwconv3x3_tma_kernel<false, 4>), smem);
  const long long cap = 148LL * 4;
  const unsigned blocks = (unsigned)(groups < cap ? groups : cap);
  const __nv_bfloat16 *xp = (const __nv_bfloat16*)x, *wp = (const __nv_bfloat16*)w;
  __nv_bfloat16* yp = (__nv_bfloat16*)y;
  if (SWd == 8) {
    if (flip) dwt::dwconv3x3_tma_kernel<true, 8><<<blocks, 256, smem, st>>>(q, xp, wp, yp, groups);
    else dwt::dwconv3x3_tma_kernel<false, 8><<<blocks, 256, smem, st>>>(q, xp, wp, yp, groups);
  } else {
    if (flip) dwt::dwconv3x3_tma_kernel<true, 4><<<blocks, 256, smem, st>>>(q, xp, wp, yp, groups);
    else dwt::dwconv3x3_tma_kernel<false, 4><<<blocks, 256, smem, st>>>(q, xp, wp, yp, groups);
  }
  return check_launch("dwconv3x3_tma_kernel");
}

template <int W>
static int dwconv3x3_wgrad_mma_launch(int H, int channels, int batch, const void* x, const void* dy, float* dw, cudaStream_t st) {
  dwt::Params q{};
  q.H = H; q.W = W; q.channels = channels; q.planes = (long long)batch * channels;
  const int psz = H * W;
  long long chunks = (148LL * 6 + channels - 1) / channels;
  if (chunks > batch) chunks = batch;
  if (chunks < 1) chunks = 1;
  if (chunks > 65535) chunks = 65535;
  const int bpb = (int)((batch + chunks - 1) / chunks);
  chunks = (batch + bpb - 1) / bpb;
  int pb = (16 * 1024) / (psz * 2);
  if (pb < 2) pb = 2;
  if (pb > 8) pb = 8;
  if (pb > bpb) pb = bpb;
  q.PB = pb;
  const size_t half = ((size_t)pb * psz * 2 + 32 + 127) / 128 * 128;
  const size_t smem = 128 + 4 * half + 16;
  TLT_ENSURE_SMEM((dwt::dwconv3x3_wgrad_mma_kernel<W>), smem);
  dim3 grid((unsigned)channels, (unsigned)chunks);
  dwt::dwconv3x3_wgrad_mma_kernel<W><<<grid, 192, smem, st>>>(q, batch, bpb, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, dw);
  return check_launch("dwconv3x3_wgrad_mma_kernel");
}

int dwconv3x3_wgrad_tma(int H, int W, int channels, int batch, const void* x, const void* dy, float* dw, cudaStream_t st) {
  {
    const char* env = getenv("TLT_DW_MMA");
    const bool off = env && env[0] == 'o' && env[1] == 'f';
    if (!off && H <= 64 && H * W * 2 <= 16 * 1024) {
      if (W == 56) return dwconv3x3_wgrad_mma_launch<56>(H, channels, batch, x, dy, dw, st);
      if (W == 64) return dwconv3x3_wgrad_mma_launch<64>(H, channels, batch, x, dy, dw, st);
      if (W == 32) return dwconv3x3_wgrad_mma_launch<32>(H, channels, batch, x, dy, dw, st);
      if (W == 16) return dwconv3x3_wgrad_mma_launch<16>(H, channels, batch, x, dy, dw, st);
    }
  }
  dwt::Params q;
  const int SWd = W % 8 == 0 ? 8 : 4;
  q.H = H; q.W = W; q.strips = W / SWd; q.channels = channels; q.planes = (long long)batch * channels;
  q.spp = H * q.strips;
  q.m_spp = dwt::magic(q.spp); q.m_strips = dwt::magic(q.strips);
  const int psz = H * W;
  long long chunks = (148LL * 6 + channels - 1) / channels;
  if (chunks > batch) chunks = batch;
  if (chunks < 1) chunks = 1;
  if (chunks > 65535) chunks = 65535;
  const int bpb = (int)((batch + chunks - 1) / chunks);
  chunks = (batch + bpb - 1) / bpb;
  int pb = (8 * 1024) / (psz * 2);                                  // ~8 KB per operand per slot: 32 KB per block
  if (pb < 1) pb = 1;
  if (pb > bpb) pb = bpb;
  q.PB = pb;
  const size_t smem = (size_t)4 * pb * psz * 2 + 16;
  TLT_ENSURE_SMEM((dwt::dwconv3x3_wgrad_tma_kernel<8>), smem);
  TLT_ENSURE_SMEM((dwt::dwconv3x3_wgrad_tma_kernel<4>), smem);
  dim3 grid((unsigned)channels, (unsigned)chunks);
  if (SWd == 8) dwt::dwconv3x3_wgrad_tma_kernel<8><<<grid, 256, smem, st>>>(q, batch, bpb, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, dw);
  else dwt::dwconv3x3_wgrad_tma_kernel<4><<<grid, 256, smem, st>>>(q, batch, bpb, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, dw);
  return check_launch("dwconv3x3_wgrad_tma_kernel");
}

}  // namespace tlt
