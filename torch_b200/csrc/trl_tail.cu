// Tail of the Tucker tensor-regression layer after the input projection (bf16): everything tucker_trl does behind
// "x x_1 F_c^T x_2 F_h^T x_3 F_w^T" (tltorch/functional/tensor_regression.py:37-49) in ONE launch forward and TWO backward:
//
//     v[b, r] = sum_{j, c} zp[b, j, c] core[c, j, r]              (zp: the projected activation, channels-last, pitch ldz >= rc)
//     y[b, o] = sum_r v[b, r] F_o[o, r] + bias[o]
//
// The reference expands the core onto the output factor first (a (rc J) x O regression matrix, 1 M entries at config 4) and takes
// the inner product with it; contracting with the rank-ro core first is the same sum, 100x fewer MACs, and keeps every operand
// of the tail (core 20 KB, F_o 20 KB) on chip.  The generic route ran this tail as ~20 launches (core expansion, a packed composite-k
// GEMM with pack / unpack passes, split-K finalisations, casts): 65 % of the layer's step at config 4 (profiles/r01_launches_trl.md).
//   trl_tail_fwd_kernel    y, v                      block = 4 batch rows
//   trl_tail_bwd_rows      dv, dzp                   block = 4 batch rows       dv = dy F_o,  dzp = dv core^T
//   trl_tail_bwd_params    dF_o, dbias, dcore        one thread per gradient entry, loop over the batch
#include "common.cuh"

namespace tlt {
namespace trlt {

constexpr int TB = 4;          // batch rows per block
constexpr int RO_MAX = 16;     // output-side rank the register tiles are sized for

struct Geom {
  int B, J, rc, ldz, ro, O;
  int P;                       // J * ldz: elements of one zp row
};

__device__ __forceinline__ float bf(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// block-wide sum of acc[TB][RO_MAX] -> red[TB][RO_MAX] (shared); 256 threads
__device__ __forceinline__ void block_sum(float (&acc)[TB][RO_MAX], int ro, float* red, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < TB; ++t)
#pragma unroll
    for (int r = 0; r < RO_MAX; ++r) {
      if (r < ro) {
        float v = acc[t][r];
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        if (lane == 0) scratch[(warp * TB + t) * RO_MAX + r] = v;
      }
    }
  __syncthreads();
  if (threadIdx.x < TB * RO_MAX) {
    const int t = threadIdx.x / RO_MAX, r = threadIdx.x % RO_MAX;
    float v = 0.f;
    if (r < ro)
      for (int w = 0; w < 8; ++w) v += scratch[(w * TB + t) * RO_MAX + r];
    red[t * RO_MAX + r] = v;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) trl_tail_fwd_kernel(const Geom g, const __nv_bfloat16* __restrict__ zp, const __nv_bfloat16* __restrict__ core,
                                                           const __nv_bfloat16* __restrict__ fo, const __nv_bfloat16* __restrict__ bias,
                                                           __nv_bfloat16* __restrict__ y, float* __restrict__ v) {
  __shared__ float red[TB * RO_MAX], scratch[8 * TB * RO_MAX];
  const int b0 = blockIdx.x * TB;
  float acc[TB][RO_MAX];
#pragma unroll
  for (int t = 0; t < TB; ++t)
#pragma unroll
    for (int r = 0; r < RO_MAX; ++r) acc[t][r] = 0.f;
  for (int p = threadIdx.x; p < g.P; p += 256) {
    const int j = p / g.ldz, c = p - j * g.ldz;
    if (c >= g.rc) continue;
    const __nv_bfloat16* cr = core + ((size_t)c * g.J + j) * g.ro;
    float z[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) z[t] = b0 + t < g.B ? bf(zp + (size_t)(b0 + t) * g.P + p) : 0.f;
#pragma unroll
    for (int r = 0; r < RO_MAX; ++r)
      if (r < g.ro) {
        const float w = bf(cr + r);
#pragma unroll
        for (int t = 0; t < TB; ++t) acc[t][r] = fmaf(z[t], w, acc[t][r]);
      }
  }
  block_sum(acc, g.ro, red, scratch);
  if (threadIdx.x < TB * RO_MAX) {
    const int t = threadIdx.x / RO_MAX, r = threadIdx.x % RO_MAX;
    if (r < g.ro && b0 + t < g.B) v[(size_t)(b0 + t) * g.ro + r] = red[threadIdx.x];
  }
  for (int o = threadIdx.x; o < g.O; o += 256) {
    float f[RO_MAX];
#pragma unroll
    for (int r = 0; r < RO_MAX; ++r) f[r] = r < g.ro ? bf(fo + (size_t)o * g.ro + r) : 0.f;
    const float bv = bias ? bf(bias + o) : 0.f;
#pragma unroll
    for (int t = 0; t < TB; ++t) {
      if (b0 + t < g.B) {
        float s = bv;
#pragma unroll
        for (int r = 0; r < RO_MAX; ++r) s = fmaf(red[t * RO_MAX + r], f[r], s);
        y[(size_t)(b0 + t) * g.O + o] = __float2bfloat16_rn(s);
      }
    }
  }
}

// dv[b, r] = sum_o gy[b, o] F_o[o, r];   dzp[b, j, c] = sum_r dv[b, r] core[c, j, r]  (pad channels: 0)
__global__ void __launch_bounds__(256) trl_tail_bwd_rows(const Geom g, const __nv_bfloat16* __restrict__ gy, const __nv_bfloat16* __restrict__ core,
                                                         const __nv_bfloat16* __restrict__ fo, float* __restrict__ dv,
                                                         __nv_bfloat16* __restrict__ dzp) {
  __shared__ float red[TB * RO_MAX], scratch[8 * TB * RO_MAX];
  const int b0 = blockIdx.x * TB;
  float acc[TB][RO_MAX];
#pragma unroll
  for (int t = 0; t < TB; ++t)
#pragma unroll
    for (int r = 0; r < RO_MAX; ++r) acc[t][r] = 0.f;
  for (int o = threadIdx.x; o < g.O; o += 256) {
    float q[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) q[t] = b0 + t < g.B ? bf(gy + (size_t)(b0 + t) * g.O + o) : 0.f;
#pragma unroll
    for (int r = 0; r < RO_MAX; ++r)
      if (r < g.ro) {
        const float w = bf(fo + (size_t)o * g.ro + r);
#pragma unroll
        for (int t = 0; t < TB; ++t) acc[t][r] = fmaf(q[t], w, acc[t][r]);
      }
  }
  block_sum(acc, g.ro, red, scratch);
  if (threadIdx.x < TB * RO_MAX) {
    const int t = threadIdx.x / RO_MAX, r = threadIdx.x % RO_MAX;
    if (r < g.ro && b0 + t < g.B) dv[(size_t)(b0 + t) * g.ro + r] = red[threadIdx.x];
  }
  for (int p = threadIdx.x; p < g.P; p += 256) {
    const int j = p / g.ldz, c = p - j * g.ldz;
    float s[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) s[t] = 0.f;
    if (c < g.rc) {
      const __nv_bfloat16* cr = core + ((size_t)c * g.J + j) * g.ro;
#pragma unroll
      for (int r = 0; r < RO_MAX; ++r)
        if (r < g.ro) {
          const float w = bf(cr + r);
#pragma unroll
          for (int t = 0; t < TB; ++t) s[t] = fmaf(red[t * RO_MAX + r], w, s[t]);
        }
    }
#pragma unroll
    for (int t = 0; t < TB; ++t)
      if (b0 + t < g.B) dzp[(size_t)(b0 + t) * g.P + p] = __float2bfloat16_rn(s[t]);
  }
}

// Parameter gradients of the tail, both reductions over the batch:
//   family 0: dF_o[o, r] = sum_b gy[b, o] v[b, r]   (+ dbias[o] = sum_b gy[b, o]);   family 1: dcore[c, j, r] = sum_b zp[b, j, c] dv[b, r]
// A block owns 64 consecutive columns (o, or the position p = j ldz + c of a zp row) and all ranks: 32 batch rows x 64 columns are
// staged in shared memory per step (16-byte global loads, the next tile's loads in flight in registers), a thread = (column, rank
// group of 4).  The first version walked the batch with one 2-byte global load per thread and step: 117 us of pure L2 latency.
constexpr int PC = 64, PB = 32;
__global__ void __launch_bounds__(256) trl_tail_bwd_params(const Geom g, const __nv_bfloat16* __restrict__ gy, const __nv_bfloat16* __restrict__ zp,
                                                           const float* __restrict__ v, const float* __restrict__ dv,
                                                           __nv_bfloat16* __restrict__ dfo, __nv_bfloat16* __restrict__ dbias,
                                                           __nv_bfloat16* __restrict__ dcore) {
  __shared__ float tile[PB][PC + 1];
  __shared__ float sv[PB][RO_MAX];
  const int fo_blocks = (g.O + PC - 1) / PC;
  const bool is_fo = (int)blockIdx.x < fo_blocks;
  const int c0 = (is_fo ? blockIdx.x : blockIdx.x - fo_blocks) * PC;
  const int ncol = is_fo ? g.O : g.P;
  const __nv_bfloat16* src = is_fo ? gy : zp;
  const float* vec = is_fo ? v : dv;
  const int tid = threadIdx.x;
  const int cl = tid & 63, rg = tid >> 6;                 // column of the tile; ranks rg, rg + 4, rg + 8, rg + 12
  const int lrow = tid >> 3, lcol = (tid & 7) * 8;         // this thread's 8 elements of a staged tile
  const bool vec_ok = (ncol % 8) == 0 && ((uintptr_t)src % 16) == 0;
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, colsum = 0.f;

  uint4 q = make_uint4(0u, 0u, 0u, 0u);
  auto fetch = [&](int b0) {                               // 8 consecutive columns of row b0 + lrow (zero outside)
    q = make_uint4(0u, 0u, 0u, 0u);
    const int b = b0 + lrow;
    if (b >= g.B) return;
    const __nv_bfloat16* row = src + (size_t)b * ncol + c0 + lcol;
    if (vec_ok && c0 + lcol + 8 <= ncol) { q = __ldg((const uint4*)row); return; }
    unsigned short e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = c0 + lcol + i < ncol ? *(const unsigned short*)(row + i) : (unsigned short)0;
    q = make_uint4(e[0] | ((uint32_t)e[1] << 16), e[2] | ((uint32_t)e[3] << 16), e[4] | ((uint32_t)e[5] << 16), e[6] | ((uint32_t)e[7] << 16));
  };
  fetch(0);
  for (int b0 = 0; b0 < g.B; b0 += PB) {
    __syncthreads();
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      tile[lrow][lcol + 2 * i] = __uint_as_float(w[i] << 16);
      tile[lrow][lcol + 2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
    for (int i = tid; i < PB * g.ro; i += 256) {
      const int b = i / g.ro, r = i - b * g.ro;
      sv[b][r] = b0 + b < g.B ? vec[(size_t)(b0 + b) * g.ro + r] : 0.f;
    }
    if (b0 + PB < g.B) fetch(b0 + PB);
    __syncthreads();
#pragma unroll 8
    for (int b = 0; b < PB; ++b) {
      const float a = tile[b][cl];
      colsum += a;
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] = fmaf(a, sv[b][rg + 4 * k], acc[k]);     // ranks >= ro hold stale shared memory: never stored
    }
  }
  const int col = c0 + cl;
  if (col >= ncol) return;
  if (is_fo) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (rg + 4 * k < g.ro) dfo[(size_t)col * g.ro + rg + 4 * k] = __float2bfloat16_rn(acc[k]);
    if (rg == 0 && dbias) dbias[col] = __float2bfloat16_rn(colsum);
  } else {
    const int j = col / g.ldz, c = col - j * g.ldz;
    if (c < g.rc) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (rg + 4 * k < g.ro) dcore[((size_t)c * g.J + j) * g.ro + rg + 4 * k] = __float2bfloat16_rn(acc[k]);
    }
  }
}

// Kronecker product of two small factor matrices (the spatial factors of the projection): k[(a, b), (p, q)] = A[a, p] B[b, q],
// and its backward dA[a, p] = sum_{b, q} dk[..] B[b, q], dB[b, q] = sum_{a, p} dk[..] A[a, p].  A generic contraction launch per
// product (three per step, ~15 us each for a 49 x 16 result) was 20 % of the layer's step.
__global__ void kron_fwd_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ Bm, int da, int pa, int db, int pb,
                                __nv_bfloat16* __restrict__ k) {
  const int n = da * db * pa * pb;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int col = i % (pa * pb), row = i / (pa * pb);
    const int a = row / db, b = row - a * db, p = col / pb, q = col - p * pb;
    k[i] = __float2bfloat16_rn(bf(A + a * pa + p) * bf(Bm + b * pb + q));
  }
}
__global__ void kron_bwd_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ Bm, const __nv_bfloat16* __restrict__ dk,
                                int da, int pa, int db, int pb, __nv_bfloat16* __restrict__ dA, __nv_bfloat16* __restrict__ dB) {
  const int nA = da * pa, nB = db * pb, ldk = pa * pb;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nA + nB; i += gridDim.x * blockDim.x) {
    float s = 0.f;
    if (i < nA) {
      const int a = i / pa, p = i - a * pa;
      for (int b = 0; b < db; ++b)
        for (int q = 0; q < pb; ++q) s = fmaf(bf(dk + (size_t)(a * db + b) * ldk + p * pb + q), bf(Bm + b * pb + q), s);
      dA[i] = __float2bfloat16_rn(s);
    } else {
      const int e = i - nA, b = e / pb, q = e - b * pb;
      for (int a = 0; a < da; ++a)
        for (int p = 0; p < pa; ++p) s = fmaf(bf(dk + (size_t)(a * db + b) * ldk + p * pb + q), bf(A + a * pa + p), s);
      dB[e] = __float2bfloat16_rn(s);
    }
  }
}

static int fill(const tlt_trl_tail_desc* d, Geom* g) {
  TLT_REQUIRE(d && d->batch >= 1 && d->j >= 1 && d->rc >= 1 && d->ldz >= d->rc && d->ro >= 1 && d->ro <= RO_MAX && d->out >= 1,
              "tlt_trl_tail: bad descriptor (output rank 1..16)");
  TLT_REQUIRE(d->batch < (1LL << 30) && d->j * d->ldz < (1LL << 24) && d->out < (1LL << 24), "tlt_trl_tail: extents too large");
  g->B = (int)d->batch; g->J = (int)d->j; g->rc = (int)d->rc; g->ldz = (int)d->ldz; g->ro = (int)d->ro; g->O = (int)d->out;
  g->P = g->J * g->ldz;
  return TLT_OK;
}

}  // namespace trlt
}  // namespace tlt

using namespace tlt;
using namespace tlt::trlt;

extern "C" int tlt_trl_tail_fwd(const tlt_trl_tail_desc* d, const void* zp, const void* core, const void* fo, const void* bias, void* y,
                                float* v, void* stream) {
  Geom g;
  int rc = fill(d, &g);
  if (rc) return rc;
  TLT_REQUIRE(zp && core && fo && y && v, "tlt_trl_tail_fwd: null pointer");
  trl_tail_fwd_kernel<<<(g.B + TB - 1) / TB, 256, 0, (cudaStream_t)stream>>>(g, (const __nv_bfloat16*)zp, (const __nv_bfloat16*)core,
                                                                            (const __nv_bfloat16*)fo, (const __nv_bfloat16*)bias,
                                                                            (__nv_bfloat16*)y, v);
  return check_launch("trl_tail_fwd_kernel");
}

// phases: 1 = the per-row launch (dv, dzp), 2 = the parameter-gradient launch (dcore, dF_o, dbias; reads dv), 3 = both.  A caller that
// overlaps the parameter gradients with the rest of its backward pass issues the two phases on different streams.
extern "C" int tlt_trl_tail_bwd(const tlt_trl_tail_desc* d, int32_t phases, const void* zp, const void* gy, const void* core, const void* fo,
                                const float* v, float* dv, void* dzp, void* dcore, void* dfo, void* dbias, void* stream) {
  Geom g;
  int rc = fill(d, &g);
  if (rc) return rc;
  TLT_REQUIRE(phases >= 1 && phases <= 3 && zp && gy && core && fo && v && dv, "tlt_trl_tail_bwd: bad arguments");
  TLT_REQUIRE(!(phases & 1) || dzp, "tlt_trl_tail_bwd: phase 1 needs dzp");
  TLT_REQUIRE(!(phases & 2) || (dcore && dfo), "tlt_trl_tail_bwd: phase 2 needs dcore and df_out");
  cudaStream_t st = (cudaStream_t)stream;
  if (phases & 1) {
    trl_tail_bwd_rows<<<(g.B + TB - 1) / TB, 256, 0, st>>>(g, (const __nv_bfloat16*)gy, (const __nv_bfloat16*)core, (const __nv_bfloat16*)fo, dv,
                                                          (__nv_bfloat16*)dzp);
    rc = check_launch("trl_tail_bwd_rows");
    if (rc) return rc;
  }
  if (!(phases & 2)) return TLT_OK;
  const int blocks = (g.O + PC - 1) / PC + (g.P + PC - 1) / PC;
  trl_tail_bwd_params<<<blocks, 256, 0, st>>>(g, (const __nv_bfloat16*)gy, (const __nv_bfloat16*)zp, v, dv, (__nv_bfloat16*)dfo,
                                              (__nv_bfloat16*)dbias, (__nv_bfloat16*)dcore);
  return check_launch("trl_tail_bwd_params");
}

extern "C" int tlt_kron2_fwd(const void* a, const void* b, int32_t da, int32_t pa, int32_t db, int32_t pb, void* k, void* stream) {
  TLT_REQUIRE(a && b && k && da >= 1 && pa >= 1 && db >= 1 && pb >= 1 && (long long)da * db * pa * pb < (1LL << 24), "tlt_kron2_fwd: bad arguments");
  const int n = da * db * pa * pb;
  kron_fwd_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, da, pa, db, pb,
                                                                    (__nv_bfloat16*)k);
  return check_launch("kron_fwd_kernel");
}

extern "C" int tlt_kron2_bwd(const void* a, const void* b, const void* dk, int32_t da, int32_t pa, int32_t db, int32_t pb, void* d_a, void* d_b,
                             void* stream) {
  TLT_REQUIRE(a && b && dk && d_a && d_b && da >= 1 && pa >= 1 && db >= 1 && pb >= 1 && (long long)da * db * pa * pb < (1LL << 24),
              "tlt_kron2_bwd: bad arguments");
  const int n = da * pa + db * pb;
  kron_bwd_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (const __nv_bfloat16*)dk,
                                                                    da, pa, db, pb, (__nv_bfloat16*)d_a, (__nv_bfloat16*)d_b);
  return check_launch("kron_bwd_kernel");
}
