// Fused chain of mode products:  y[b] = x[b] x_1 M_1 x_2 M_2 x_3 M_3   for every sample b, intermediates in shared memory.
//
//   x (batch, d1, d2, d3) -> y (batch, r1, r2, r3),   y[b, j1, j2, j3] = sum_{i1,i2,i3} x[b, i1, i2, i3] M1[j1,i1] M2[j2,i2] M3[j3,i3]
//
// Replaces the tenalg.multi_mode_dot chains of the reference: TCL.forward (factorized_layers/tensor_contraction_layers.py:75),
// tucker_trl's input projection (functional/tensor_regression.py:40), and the project-in / expand-out stages of
// linear_tucker (functional/factorized_tensordot.py:10-66 evaluated in FLOP-sane order).  The reference (and the first
// version here) runs one matmul per mode with the partially-projected tensor making a round trip through HBM each
// time; at BASELINE config 5 (16^3 -> 15^3 per token, 8192 tokens) that is 3 x (67 MB read + 63 MB write) plus, on the
// tensor-core bridge, a pack and an unpack pass per mode.  Here a block keeps one sample on chip: x[b] is staged once,
// the three small contractions ping-pong between two shared-memory buffers, and only y[b] is written -- 4 bytes of HBM
// traffic per element of x instead of ~24.
//
// Each step is  out[o, j, n] = sum_i in[o, i, n] * M[j, i]  (o = already-processed outer modes, n = not-yet-processed inner
// modes).  A thread owns one (o, n) pair and a tile of 16 outputs j in registers; inputs are read with consecutive
// threads on consecutive n (conflict-free), the matrix row M[j, i..i+3] as a broadcast float4.  Modes are processed
// outermost first so that the last step writes y[b] to global memory with consecutive threads on consecutive rows.
#include "common.cuh"
#include <mutex>
#include <unordered_map>

namespace tlt {

constexpr int kJT = 16;       // outputs of one (o, n) pair held in registers

struct MmdParams {
  int d[3], r[3];             // contracted / produced extents per mode
  int skip[3];                // 1: mode untouched (d == r, no matrix)
  int mp[3];                  // pitch of the staged matrices: round4(d)
  int moff[3];                // offsets of the staged matrices in shared memory (floats)
  int off_a, off_b;           // the two activation buffers
  int P2;                     // pitch of the innermost rows (length d[2]) in shared memory: odd when mode 2 is contracted
  long long batch;
  long long in_elems, out_elems;
};

// One mode product with the innermost mode (extent d2, pitch P2 in shared memory) untouched:
//   out[(o * r + j), nh, i2] = sum_i in[(o * d + i), nh, i2] * M[j, i]          nh in [0, inner_hi)
// A thread owns one (o, nh, i2) and kJT outputs j.  TO_GLOBAL: the last step writes y (compact, type T).
template <typename T, bool TO_GLOBAL>
__device__ __forceinline__ void mmd_step(const float* __restrict__ in, float* __restrict__ out_s, T* __restrict__ out_g,
                                         const T* __restrict__ addend, const float* __restrict__ M, int outer, int d, int r,
                                         int inner_hi, int d2, int P2, int mp) {
  const int inner = inner_hi * d2;                 // logical inner extent
  const int slab = inner_hi * P2;                  // physical size of one (o, i) slab in shared memory
  const int pairs = outer * inner;
  for (int e = threadIdx.x; e < pairs; e += blockDim.x) {
    const int o = e / inner, n = e - o * inner;
    const int nh = n / d2, i2 = n - nh * d2;
    const float* ip = in + o * d * slab + nh * P2 + i2;
    for (int j0 = 0; j0 < r; j0 += kJT) {
      float acc[kJT];
#pragma unroll
      for (int j = 0; j < kJT; ++j) acc[j] = 0.f;
      for (int i = 0; i < d; i += 4) {
        const float x0 = ip[i * slab];
        const float x1 = i + 1 < d ? ip[(i + 1) * slab] : 0.f;
        const float x2 = i + 2 < d ? ip[(i + 2) * slab] : 0.f;
        const float x3 = i + 3 < d ? ip[(i + 3) * slab] : 0.f;
#pragma unroll
        for (int j = 0; j < kJT; ++j) {
          if (j0 + j < r) {
            const float4 m = *(const float4*)(M + (j0 + j) * mp + i);       // zero-padded past d
            acc[j] = fmaf(x0, m.x, acc[j]);
            acc[j] = fmaf(x1, m.y, acc[j]);
            acc[j] = fmaf(x2, m.z, acc[j]);
            acc[j] = fmaf(x3, m.w, acc[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < kJT; ++j) {
        if (j0 + j < r) {
          if (TO_GLOBAL) {
            const long long idx = (long long)(o * r + j0 + j) * inner + n;
            out_g[idx] = from_f32<T>(addend ? acc[j] + to_f32<T>(addend[idx]) : acc[j]);
          } else out_s[(o * r + j0 + j) * slab + nh * P2 + i2] = acc[j];
        }
      }
    }
  }
}

// The innermost mode:  y[o, j] = sum_i in[o * P2 + i] * M[j, i]     (rows at odd pitch P2: conflict-free strided reads)
template <typename T>
__device__ __forceinline__ void mmd_last(const float* __restrict__ in, T* __restrict__ out_g, const T* __restrict__ addend,
                                         const float* __restrict__ M, int rows, int d, int r, int P2, int mp) {
  for (int o = threadIdx.x; o < rows; o += blockDim.x) {
    const float* ip = in + o * P2;
    for (int j0 = 0; j0 < r; j0 += kJT) {
      float acc[kJT];
#pragma unroll
      for (int j = 0; j < kJT; ++j) acc[j] = 0.f;
      for (int i = 0; i < d; i += 4) {
        const float x0 = ip[i];
        const float x1 = i + 1 < d ? ip[i + 1] : 0.f;
        const float x2 = i + 2 < d ? ip[i + 2] : 0.f;
        const float x3 = i + 3 < d ? ip[i + 3] : 0.f;
#pragma unroll
        for (int j = 0; j < kJT; ++j) {
          if (j0 + j < r) {
            const float4 m = *(const float4*)(M + (j0 + j) * mp + i);
            acc[j] = fmaf(x0, m.x, acc[j]);
            acc[j] = fmaf(x1, m.y, acc[j]);
            acc[j] = fmaf(x2, m.z, acc[j]);
            acc[j] = fmaf(x3, m.w, acc[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < kJT; ++j)
        if (j0 + j < r) {
          const long long idx = (long long)o * r + j0 + j;
          out_g[idx] = from_f32<T>(addend ? acc[j] + to_f32<T>(addend[idx]) : acc[j]);
        }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) multi_mode_dot_kernel(const MmdParams p, const T* __restrict__ x,
                                                             const T* __restrict__ m0, const T* __restrict__ m1,
                                                             const T* __restrict__ m2, int t0, int t1, int t2,
                                                             const T* __restrict__ addend, T* __restrict__ y) {
  extern __shared__ __align__(16) float msm[];
  // stage the matrices once per block as M[j][i] (row pitch mp, zero padded), whatever their storage order
  const T* mats[3] = {m0, m1, m2};
  const int trans[3] = {t0, t1, t2};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (p.skip[k]) continue;
    float* Ms = msm + p.moff[k];
    for (int e = threadIdx.x; e < p.r[k] * p.mp[k]; e += blockDim.x) {
      const int j = e / p.mp[k], i = e - j * p.mp[k];
      float v = 0.f;
      if (i < p.d[k]) v = to_f32<T>(trans[k] ? mats[k][(long long)i * p.r[k] + j] : mats[k][(long long)j * p.d[k] + i]);
      Ms[e] = v;
    }
  }
  const int last = !p.skip[2] ? 2 : (!p.skip[1] ? 1 : 0);     // at least one mode is contracted (host-checked)
  const int d2 = p.d[2], P2 = p.P2;
  for (long long b = blockIdx.x; b < p.batch; b += gridDim.x) {
    __syncthreads();                                           // previous sample's readers are done; matrices are staged
    const T* xb = x + b * p.in_elems;
    T* yb = y + b * p.out_elems;
    float* cur = msm + p.off_a;
    float* nxt = msm + p.off_b;
    const int n_in = (int)p.in_elems;
    if (P2 == d2) {
      for (int e = threadIdx.x; e < n_in; e += blockDim.x) cur[e] = to_f32<T>(xb[e]);
    } else {
      for (int e = threadIdx.x; e < n_in; e += blockDim.x) {
        const int row = e / d2, c = e - row * d2;
        cur[row * P2 + c] = to_f32<T>(xb[e]);
      }
    }
    __syncthreads();
    int e0 = p.d[0], e1 = p.d[1];
    if (!p.skip[0]) {
      if (last == 0) mmd_step<T, true>(cur, nullptr, yb, addend, msm + p.moff[0], 1, p.d[0], p.r[0], p.d[1], d2, P2, p.mp[0]);
      else mmd_step<T, false>(cur, nxt, nullptr, nullptr, msm + p.moff[0], 1, p.d[0], p.r[0], p.d[1], d2, P2, p.mp[0]);
      __syncthreads();
      float* t = cur; cur = nxt; nxt = t;
      e0 = p.r[0];
    }
    if (!p.skip[1]) {
      if (last == 1) mmd_step<T, true>(cur, nullptr, yb, addend, msm + p.moff[1], e0, p.d[1], p.r[1], 1, d2, P2, p.mp[1]);
      else mmd_step<T, false>(cur, nxt, nullptr, nullptr, msm + p.moff[1], e0, p.d[1], p.r[1], 1, d2, P2, p.mp[1]);
      __syncthreads();
      float* t = cur; cur = nxt; nxt = t;
      e1 = p.r[1];
    }
    if (!p.skip[2]) mmd_last<T>(cur, yb, addend, msm + p.moff[2], e0 * e1, d2, p.r[2], P2, p.mp[2]);
  }
}

static int mmd_round4(int v) { return (v + 3) & ~3; }

}  // namespace tlt

using namespace tlt;

extern "C" int tlt_multi_mode_dot(const tlt_mmd_desc* d, const void* x, const void* m0, const void* m1, const void* m2,
                                  const void* addend, void* y, void* stream) {
  TLT_REQUIRE(d && x && y, "tlt_multi_mode_dot: null argument");
  TLT_REQUIRE(d->dtype == TLT_F32 || d->dtype == TLT_BF16, "tlt_multi_mode_dot: dtype must be f32 or bf16");
  TLT_REQUIRE(d->batch >= 1, "tlt_multi_mode_dot: empty batch");
  MmdParams p{};
  const void* mats[3] = {m0, m1, m2};
  long long in_e = 1, out_e = 1;
  int any = 0;
  for (int k = 0; k < 3; ++k) {
    TLT_REQUIRE(d->in_dims[k] >= 1 && d->out_dims[k] >= 1 && d->in_dims[k] <= 4096 && d->out_dims[k] <= 4096,
                "tlt_multi_mode_dot: mode sizes out of range");
    p.d[k] = d->in_dims[k]; p.r[k] = d->out_dims[k];
    p.skip[k] = mats[k] == nullptr;
    if (p.skip[k]) TLT_REQUIRE(p.d[k] == p.r[k], "tlt_multi_mode_dot: a skipped mode must keep its size");
    any |= !p.skip[k];
    p.mp[k] = mmd_round4(p.d[k]);
    in_e *= p.d[k]; out_e *= p.r[k];
  }
  TLT_REQUIRE(any, "tlt_multi_mode_dot: no mode to contract");
  p.batch = d->batch; p.in_elems = in_e; p.out_elems = out_e;
  p.P2 = p.skip[2] ? p.d[2] : (p.d[2] | 1);
  // shared memory: matrices, then two activation buffers sized for the largest intermediate (innermost rows at pitch P2)
  int off = 0;
  for (int k = 0; k < 3; ++k) { p.moff[k] = off; if (!p.skip[k]) off += mmd_round4(p.r[k] * p.mp[k]); }
  const long long e0 = p.skip[0] ? p.d[0] : p.r[0], e1 = p.skip[1] ? p.d[1] : p.r[1];
  const long long s_in = (long long)p.d[0] * p.d[1] * p.P2;
  const long long s_0 = e0 * p.d[1] * p.P2;
  const long long s_1 = e0 * e1 * p.P2;
  long long buf = s_in > s_0 ? s_in : s_0;
  if (s_1 > buf) buf = s_1;
  buf = (buf + 3) & ~3LL;
  const long long total = (long long)off + 2 * buf;
  if (total * 4 > 200 * 1024 || in_e >= (1LL << 24) || out_e >= (1LL << 24)) {
    set_error("tlt_multi_mode_dot: one sample (%lld -> %lld elements) does not fit shared memory; use the per-mode chain",
              in_e, out_e);
    return TLT_ERR_UNSUPPORTED;
  }
  p.off_a = off; p.off_b = off + (int)buf;
  const size_t smem = (size_t)total * 4;
  cudaStream_t st = (cudaStream_t)stream;
  long long blocks = d->batch < 148LL * 8 ? d->batch : 148LL * 8;
  const int ti = d->dtype == TLT_BF16;
  if (ti) TLT_ENSURE_SMEM(multi_mode_dot_kernel<__nv_bfloat16>, smem);
  else TLT_ENSURE_SMEM(multi_mode_dot_kernel<float>, smem);
  if (ti)
    multi_mode_dot_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, smem, st>>>(p, (const __nv_bfloat16*)x, (const __nv_bfloat16*)m0,
        (const __nv_bfloat16*)m1, (const __nv_bfloat16*)m2, d->transpose[0], d->transpose[1], d->transpose[2], (const __nv_bfloat16*)addend, (__nv_bfloat16*)y);
  else
    multi_mode_dot_kernel<float><<<(unsigned)blocks, 256, smem, st>>>(p, (const float*)x, (const float*)m0, (const float*)m1,
        (const float*)m2, d->transpose[0], d->transpose[1], d->transpose[2], (const float*)addend, (float*)y);
  return check_launch("multi_mode_dot_kernel");
}
