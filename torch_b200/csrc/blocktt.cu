// Fused BlockTT (TT-matrix) kernels for the 3-core case: reconstruction of the dense matrix from the cores, and the
// chain rule from the dense-matrix gradient back onto the three cores.
//
//   W[(o1,o2,o3),(i1,i2,i3)] = sum_{r1,r2} C1[0,o1,i1,r1] * C2[r1,o2,i2,r2] * C3[r2,o3,i3,0]
//
// Reference: BlockTT.to_tensor, tltorch/factorized_tensors/tensorized_matrices.py:354-384 (pairwise einsum loop
// 'adeb,bfgc->adfegc' followed by a permute), and its autograd.  At BASELINE config 2 the cores hold 31 744 parameters
// and W 2.36 M elements: the whole rebuild is ~75 MFLOP and 4.7 MB of writes, so one launch keeps the rank-indexed
// intermediate T12[o1,i1,o2,i2,r2] in shared memory (it never exists in HBM) and writes W in full rows.  The generic
// pairwise chain (tlt_contract) needed 2 launches forward and 8 backward (~50 + ~100 us, latency-bound grids).
//
// With only R2 (= 16) multiply-adds per output element the kernels are issue-bound, not bandwidth-bound, so every inner
// loop is register-tiled: a thread owns a strip of outputs and keeps the operand they share in registers (8 x 4 outputs
// per T12 value in the reconstruction, 16 ranks per dW value in the gradient kernels), which cuts the instruction count
// per multiply-add from ~6 (scalar version: 33 / 44 us per launch) to ~1.4.
#include "common.cuh"
#include <mutex>
#include <unordered_map>

namespace tlt {

constexpr int kG3Max = 8;     // o3 rows per block (accumulator strip height of the reconstruction)
constexpr int kRT = 16;       // rank tile held in registers by the gradient kernels

// Division by a launch-time constant: q = umulhi(n, m) with m = floor(2^32 / d) + 1, exact while n * d < 2^32 (every index
// here is far below that).  Integer division by a runtime value costs ~25 instructions on the SM; these kernels do one or
// two per multiply-add strip, and ncu showed them executing 8x the instructions the arithmetic needs.
struct FastDiv {
  uint32_t d, m;
  __host__ __device__ FastDiv() : d(1), m(0) {}
  __host__ __device__ explicit FastDiv(int dd) : d((uint32_t)dd), m(dd > 1 ? (uint32_t)((1ull << 32) / (uint32_t)dd) + 1u : 0u) {}
  __device__ __forceinline__ int div(int n) const { return d > 1 ? (int)__umulhi((uint32_t)n, m) : n; }
  __device__ __forceinline__ void divmod(int n, int& q, int& r) const { q = div(n); r = n - q * (int)d; }
};

struct Btt3Params {
  int O1, O2, O3, I1, I2, I3, R1, R2;
  FastDiv fI3, fI3p, fI3v, fG3, fR2r, fR2, fI2, fI1, fO2, fR1, fRow, fRun, fRun4;
  int G3, n_g3;     // o3 rows per block / number of o3 groups
  int IC, n_ic;     // (i1,i2) pairs per block / number of such chunks
  int R2r, I3p;     // R2 and I3 rounded up to a multiple of 4 (float4 rows in shared memory)
  int I3s;          // pitch of one (o3, i12) run of the staged dW tile (odd: conflict-free across i12)
  long long I, O;   // I1*I2*I3, O1*O2*O3
};

// Shared-memory carve-up (floats):  C1s [I1][R1] | C2s [R1][I2][R2] | T12 [IC][R2r] | C3 slice | dWs [G3][IC][I3s]
struct Btt3Smem {
  int c1s, c2s, t12, c3, dws, total;
};
__host__ __device__ inline int btt3_round4(int v) { return (v + 3) & ~3; }
__host__ __device__ inline Btt3Smem btt3_carve(const Btt3Params& p, bool grads) {
  Btt3Smem m;
  m.c1s = 0;
  m.c2s = m.c1s + btt3_round4(p.I1 * p.R1);
  m.t12 = m.c2s + btt3_round4(p.R1 * p.I2 * p.R2);
  m.c3 = m.t12 + btt3_round4(p.IC * p.R2r);
  // reconstruction: C3s [R2][G3][I3p];  gradients: C3t [G3*I3][R2r]
  m.dws = m.c3 + (grads ? btt3_round4(p.G3 * p.I3 * p.R2r) : btt3_round4(p.R2 * p.G3 * p.I3p));
  m.total = m.dws + (grads ? p.G3 * p.IC * p.I3s : 0);
  return m;
}


// Cooperative global -> shared staging of `rows` runs of `run` contiguous elements (row pitch `pitch` elements).  When the
// runs are 16-byte aligned multiples of a 16-byte vector, each thread issues up to kStageBatch independent 16-byte loads
// before consuming any of them: these kernels are latency-bound (0.65 waves of blocks march through the phases in
// lockstep), so memory-level parallelism per thread is what shortens a phase.  sink(row, col, value) writes shared memory.
constexpr int kStageBatch = 4;
template <typename T> __device__ __forceinline__ void unpack16(const uint4& raw, float* out);
template <> __device__ __forceinline__ void unpack16<float>(const uint4& raw, float* out) {
  out[0] = __uint_as_float(raw.x); out[1] = __uint_as_float(raw.y); out[2] = __uint_as_float(raw.z); out[3] = __uint_as_float(raw.w);
}
template <> __device__ __forceinline__ void unpack16<__nv_bfloat16>(const uint4& raw, float* out) {
  const __nv_bfloat162* h = (const __nv_bfloat162*)&raw;
#pragma unroll
  for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); out[2 * j] = f.x; out[2 * j + 1] = f.y; }
}
template <typename T, typename F>
__device__ __forceinline__ void stage_rows(const T* __restrict__ src, long long pitch, int rows, int run, F&& sink) {
  constexpr int V = 16 / (int)sizeof(T);
  const bool vec = (run % V) == 0 && (pitch % V) == 0 && (((uintptr_t)src) & 15) == 0;
  if (vec) {
    const int vpr = run / V, total = rows * vpr;
    const FastDiv fv(vpr);
    for (int base = threadIdx.x; base < total; base += blockDim.x * kStageBatch) {
      uint4 raw[kStageBatch];
#pragma unroll
      for (int b = 0; b < kStageBatch; ++b) {
        const int e = base + b * blockDim.x;
        if (e < total) {
          int row, q;
          fv.divmod(e, row, q);
          raw[b] = __ldg((const uint4*)(src + (long long)row * pitch) + q);
        }
      }
#pragma unroll
      for (int b = 0; b < kStageBatch; ++b) {
        const int e = base + b * blockDim.x;
        if (e < total) {
          int row, q;
          fv.divmod(e, row, q);
          float f[V];
          unpack16<T>(raw[b], f);
#pragma unroll
          for (int j = 0; j < V; ++j) sink(row, q * V + j, f[j]);
        }
      }
    }
  } else {
    const int total = rows * run;
    const FastDiv fr(run);
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
      int row, q;
      fr.divmod(e, row, q);
      sink(row, q, to_f32<T>(src[(long long)row * pitch + q]));
    }
  }
}

// T12[i12 - ic0][r2] (row pitch R2r) = sum_r1 C1[o1,i1,r1] * C2[r1,o2,i2,r2] for the block's chunk of (i1,i2) pairs.  The
// two core slices are first staged in shared memory with coalesced, independent loads.
template <typename T>
__device__ __forceinline__ void btt3_build_t12(const Btt3Params& p, const Btt3Smem& m, float* smem_f, int o1, int o2,
                                                int ic0, int icn, const T* __restrict__ c1, const T* __restrict__ c2) {
  float* C1s = smem_f + m.c1s;
  float* C2s = smem_f + m.c2s;
  float* T12 = smem_f + m.t12;
  const int n1 = p.I1 * p.R1;                                       // C1[0, o1, :, :] is contiguous
  stage_rows<T>(c1 + (long long)o1 * n1, n1, 1, n1, [&](int, int q, float v) { C1s[q] = v; });
  const int row = p.I2 * p.R2;                                      // C2[r1, o2, :, :] is contiguous
  stage_rows<T>(c2 + (long long)o2 * row, (long long)p.O2 * row, p.R1, row, [&](int r1, int q, float v) { C2s[r1 * row + q] = v; });
  __syncthreads();
  for (int e = threadIdx.x; e < icn * p.R2r; e += blockDim.x) {
    int il, r2;
    p.fR2r.divmod(e, il, r2);
    float acc = 0.f;
    if (r2 < p.R2) {
      int i1, i2;
      p.fI2.divmod(ic0 + il, i1, i2);
      const float* a = C1s + i1 * p.R1;
      const float* b = C2s + i2 * p.R2 + r2;
#pragma unroll 4
      for (int r1 = 0; r1 < p.R1; ++r1) acc = fmaf(a[r1], b[r1 * row], acc);
    }
    T12[e] = acc;                                                   // ranks R2..R2r-1 are zero padding
  }
}

template <typename T, int VEC> struct Pack;
template <> struct Pack<float, 4> {
  static __device__ __forceinline__ void store(float* dst, const float* v) { *(float4*)dst = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Pack<__nv_bfloat16, 4> {
  static __device__ __forceinline__ void store(__nv_bfloat16* dst, const float* v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 o; o.x = *(uint32_t*)&a; o.y = *(uint32_t*)&b;
    *(uint2*)dst = o;
  }
};
template <typename T> struct Pack<T, 1> {
  static __device__ __forceinline__ void store(T* dst, const float* v) { dst[0] = from_f32<T>(v[0]); }
};

// grid (O1*O2, n_g3, n_ic), 256 threads.  A thread owns VEC consecutive columns (same (i1,i2), consecutive i3) of up to
// kG3Max consecutive rows: the T12 value is loaded once per rank and reused for all of them.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) btt3_reconstruct_kernel(const Btt3Params p, const T* __restrict__ c1,
                                                               const T* __restrict__ c2, const T* __restrict__ c3,
                                                               T* __restrict__ W) {
  extern __shared__ __align__(16) float smem_f[];
  const Btt3Smem m = btt3_carve(p, false);
  const float* T12 = smem_f + m.t12;
  float* C3s = smem_f + m.c3;
  const int o12 = blockIdx.x, o1 = o12 / p.O2, o2 = o12 % p.O2;
  const int o3_begin = blockIdx.y * p.G3;
  const int g3 = min(p.G3, p.O3 - o3_begin);
  const int ic0 = blockIdx.z * p.IC;
  const int icn = min(p.IC, p.I1 * p.I2 - ic0);
  if (g3 < p.G3 || p.I3p != p.I3) {                                 // ragged group / padded rows: zero the slice first
    for (int e = threadIdx.x; e < p.R2 * p.G3 * p.I3p; e += blockDim.x) C3s[e] = 0.f;
    __syncthreads();
  }
  // C3[r2, o3_begin .. o3_begin+g3, :] is one contiguous run per rank
  stage_rows<T>(c3 + (long long)o3_begin * p.I3, (long long)p.O3 * p.I3, p.R2, g3 * p.I3, [&](int r2, int q, float v) {
    int o3l, i3;
    p.fI3.divmod(q, o3l, i3);
    C3s[(r2 * p.G3 + o3l) * p.I3p + i3] = v;
  });
  btt3_build_t12<T>(p, m, smem_f, o1, o2, ic0, icn, c1, c2);
  __syncthreads();
  const int per_row = p.I3 / VEC;                 // VEC == 4 requires I3 % 4 == 0 (host-checked)
  const int items = icn * per_row;
  const int cstride = p.G3 * p.I3p;
  T* Wbase = W + (((long long)o1 * p.O2 + o2) * p.O3 + o3_begin) * p.I + (long long)ic0 * p.I3;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    int il, i3;
    (VEC == 4 ? p.fI3v : p.fI3).divmod(it, il, i3);
    i3 *= VEC;
    const float* t = T12 + il * p.R2r;
    const float* c = C3s + i3;
    float acc[kG3Max][VEC];
#pragma unroll
    for (int g = 0; g < kG3Max; ++g)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[g][v] = 0.f;
    for (int r2 = 0; r2 < p.R2; ++r2) {
      const float tv = t[r2];
      const float* cr = c + r2 * cstride;
#pragma unroll
      for (int g = 0; g < kG3Max; ++g) {
        if (g < p.G3) {
          if (VEC == 4) {
            const float4 cv = *(const float4*)(cr + g * p.I3p);
            acc[g][0] = fmaf(tv, cv.x, acc[g][0]);
            acc[g][VEC > 1 ? 1 : 0] = fmaf(tv, cv.y, acc[g][VEC > 1 ? 1 : 0]);
            acc[g][VEC > 2 ? 2 : 0] = fmaf(tv, cv.z, acc[g][VEC > 2 ? 2 : 0]);
            acc[g][VEC > 3 ? 3 : 0] = fmaf(tv, cv.w, acc[g][VEC > 3 ? 3 : 0]);
          } else {
            acc[g][0] = fmaf(tv, cr[g * p.I3p], acc[g][0]);
          }
        }
      }
    }
    T* dst = Wbase + il * p.I3 + i3;
#pragma unroll
    for (int g = 0; g < kG3Max; ++g)
      if (g < g3) Pack<T, VEC>::store(dst + (long long)g * p.I, acc[g]);
  }
}

// Core-gradient stage A.  grid (O1*O2, n_g3, n_ic), 256 threads.
//   dT12[o12][i12][r2] += sum_{o3 in group, i3} dW[o3][i12,i3] * C3[r2,o3,i3]                 (atomic, n_g3-way contention)
//   dC3p[o12,ic][r2][o3][i3] = sum_{i12 in chunk} T12[i12][r2] * dW[o3][i12,i3]                 (partials, reduced in stage B)
// Both loops keep a tile of kRT ranks in registers per staged dW value.
template <typename T>
__global__ void __launch_bounds__(256) btt3_grad_a_kernel(const Btt3Params p, int KQ, int KH, const T* __restrict__ c1,
                                                          const T* __restrict__ c2, const T* __restrict__ c3,
                                                          const float* __restrict__ dW, float* __restrict__ dT12,
                                                          float* __restrict__ dC3p) {
  extern __shared__ __align__(16) float smem_f[];
  const Btt3Smem m = btt3_carve(p, true);
  const int I12 = p.I1 * p.I2;
  const float* T12 = smem_f + m.t12;
  float* C3t = smem_f + m.c3;
  float* dWs = smem_f + m.dws;
  const int o12 = blockIdx.x, o1 = o12 / p.O2, o2 = o12 % p.O2;
  const int o3_begin = blockIdx.y * p.G3;
  const int g3 = min(p.G3, p.O3 - o3_begin);
  const int ic0 = blockIdx.z * p.IC;
  const int icn = min(p.IC, I12 - ic0);
  // the dW tile first: the longest-latency loads go out before anything waits.  dWs[o3l][il][i3], pitch I3s per (o3l, il)
  const float* dW_rows = dW + (((long long)o1 * p.O2 + o2) * p.O3 + o3_begin) * p.I + (long long)ic0 * p.I3;
  stage_rows<float>(dW_rows, p.I, g3, icn * p.I3, [&](int o3l, int q, float v) {
    int il, i3;
    p.fI3.divmod(q, il, i3);
    dWs[(o3l * p.IC + il) * p.I3s + i3] = v;
  });
  if (p.R2r != p.R2) {                                              // zero the rank padding of C3t
    for (int e = threadIdx.x; e < g3 * p.I3 * p.R2r; e += blockDim.x) C3t[e] = 0.f;
    __syncthreads();
  }
  // C3t[k = o3l*I3 + i3][r2]: one contiguous source run per rank, scattered with r2 fastest
  stage_rows<T>(c3 + (long long)o3_begin * p.I3, (long long)p.O3 * p.I3, p.R2, g3 * p.I3,
                [&](int r2, int k, float v) { C3t[k * p.R2r + r2] = v; });
  btt3_build_t12<T>(p, m, smem_f, o1, o2, ic0, icn, c1, c2);
  __syncthreads();

  const int kq_shift = 31 - __clz(KQ), kh_shift = 31 - __clz(KH);      // KQ, KH are powers of two
  for (int rt0 = 0; rt0 < p.R2; rt0 += kRT) {
    // (1) dT12: item = (il, kq); the KQ lanes of an item split the (o3l, i3) range and meet through shuffles
    {
      const int items = icn * KQ;
      const int K = g3 * p.I3;
      for (int base = 0; base < items; base += blockDim.x) {
        const int it = base + threadIdx.x;
        const bool live = it < items;
        const int il = live ? it >> kq_shift : 0, kq = it & (KQ - 1);
        float acc[kRT];
#pragma unroll
        for (int r = 0; r < kRT; ++r) acc[r] = 0.f;
        if (live) {
          for (int k = kq; k < K; k += KQ) {
            int o3l, i3;
            p.fI3.divmod(k, o3l, i3);
            const float d = dWs[(o3l * p.IC + il) * p.I3s + i3];
            const float4* cr = (const float4*)(C3t + k * p.R2r + rt0);
#pragma unroll
            for (int q = 0; q < kRT / 4; ++q) {
              if (rt0 + 4 * q < p.R2r) {
                const float4 c = cr[q];
                acc[4 * q] = fmaf(d, c.x, acc[4 * q]);
                acc[4 * q + 1] = fmaf(d, c.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(d, c.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(d, c.w, acc[4 * q + 3]);
              }
            }
          }
        }
        for (int s = 1; s < KQ; s <<= 1) {
#pragma unroll
          for (int r = 0; r < kRT; ++r) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], s);
        }
        if (live && kq == 0) {
          float* dst = dT12 + ((long long)o12 * I12 + ic0 + il) * p.R2 + rt0;
#pragma unroll
          for (int r = 0; r < kRT; ++r)
            if (rt0 + r < p.R2) atomicAdd(dst + r, acc[r]);
        }
      }
    }
    // (2) dC3 partials: item = (o3l, i3, kh); the KH lanes of an item split the chunk's (i1,i2) pairs
    {
      const int N = g3 * p.I3;
      const int items = N * KH;
      for (int base = 0; base < items; base += blockDim.x) {
        const int it = base + threadIdx.x;
        const bool live = it < items;
        const int n = live ? it >> kh_shift : 0, kh = it & (KH - 1);
        int o3l, i3;
        p.fI3.divmod(n, o3l, i3);
        float acc[kRT];
#pragma unroll
        for (int r = 0; r < kRT; ++r) acc[r] = 0.f;
        if (live) {
          const float* dcol = dWs + o3l * p.IC * p.I3s + i3;
          for (int il = kh; il < icn; il += KH) {
            const float d = dcol[il * p.I3s];
            const float4* tr = (const float4*)(T12 + il * p.R2r + rt0);
#pragma unroll
            for (int q = 0; q < kRT / 4; ++q) {
              if (rt0 + 4 * q < p.R2r) {
                const float4 t = tr[q];
                acc[4 * q] = fmaf(d, t.x, acc[4 * q]);
                acc[4 * q + 1] = fmaf(d, t.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(d, t.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(d, t.w, acc[4 * q + 3]);
              }
            }
          }
        }
        for (int s = 1; s < KH; s <<= 1) {
#pragma unroll
          for (int r = 0; r < kRT; ++r) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], s);
        }
        if (live && kh == 0) {
          float* dst = dC3p + ((long long)(o12 * p.n_ic + blockIdx.z) * p.R2 + rt0) * p.O3 * p.I3 +
                       (long long)(o3_begin + o3l) * p.I3 + i3;
#pragma unroll
          for (int r = 0; r < kRT; ++r)
            if (rt0 + r < p.R2) dst[(long long)r * p.O3 * p.I3] = acc[r];
        }
      }
    }
  }
}

// Core-gradient stage B: three roles by block range; every output is summed by a group of lanes (the terms are
// independent loads, so splitting them across lanes hides the L2 latency that a per-thread serial sum exposes).
//   dC3[r2,o3,i3]    = sum_{o12,ic} dC3p[o12,ic][r2,o3,i3]                                     8 partial sums per output
//   dC2[r1,o2,i2,r2] = sum_{o1,i1} C1[o1,i1,r1] * dT12[o1,o2][i1,i2][r2]                       4 partial sums per output
//   dC1[o1,i1,r1]    = sum_{o2,i2,r2} dT12[o1,o2][i1,i2][r2] * C2[r1,o2,i2,r2]                 one warp per output
template <typename T>
__global__ void __launch_bounds__(256) btt3_grad_b_kernel(const Btt3Params p, int blocks_c3, int blocks_c2,
                                                          const T* __restrict__ c1, const T* __restrict__ c2,
                                                          const float* __restrict__ dT12, const float* __restrict__ dC3p,
                                                          T* __restrict__ d1, T* __restrict__ d2, T* __restrict__ d3) {
  const int I12 = p.I1 * p.I2;
  __shared__ float red[8][65];
  int b = blockIdx.x;
  if (b < blocks_c3) {
    const long long n3 = (long long)p.R2 * p.O3 * p.I3;
    const int parts = p.O1 * p.O2 * p.n_ic;
    const int sub = threadIdx.x >> 5;                                   // 8 warps = 8 partial sums per output
    const int lane = threadIdx.x & 31;
    const long long e = (long long)b * 32 + lane;                       // consecutive lanes -> consecutive outputs
    float acc = 0.f;
    if (e < n3) {
#pragma unroll 16
      for (int q = sub; q < parts; q += 8) acc += dC3p[(long long)q * n3 + e];
    }
    red[sub][lane] = acc;
    __syncthreads();
    if (sub == 0 && e < n3) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][lane];
      d3[e] = from_f32<T>(s);
    }
    return;
  }
  b -= blocks_c3;
  if (b < blocks_c2) {
    const long long n2 = (long long)p.R1 * p.O2 * p.I2 * p.R2;
    const int sub = threadIdx.x >> 6;                                   // 4 groups of 64 threads
    const int idx = threadIdx.x & 63;
    const long long e = (long long)b * 64 + idx;
    float acc = 0.f;
    if (e < n2) {
      int t, r2, i2, o2, r1;
      p.fR2.divmod((int)e, t, r2);
      p.fI2.divmod(t, t, i2);
      p.fO2.divmod(t, r1, o2);
      const int terms = p.O1 * p.I1;
#pragma unroll 16
      for (int q = sub; q < terms; q += 4) {
        int o1, i1;
        p.fI1.divmod(q, o1, i1);
        const float a = to_f32<T>(c1[(long long)q * p.R1 + r1]);
        acc = fmaf(a, dT12[(((long long)(o1 * p.O2 + o2)) * I12 + i1 * p.I2 + i2) * p.R2 + r2], acc);
      }
    }
    red[sub][idx] = acc;
    __syncthreads();
    if (sub == 0 && e < n2) d2[e] = from_f32<T>(red[0][idx] + red[1][idx] + red[2][idx] + red[3][idx]);
    return;
  }
  b -= blocks_c2;
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n1 = (long long)p.O1 * p.I1 * p.R1;
    const long long e = (long long)b * (blockDim.x >> 5) + warp;
    if (e >= n1) return;
    int t, r1, i1, o1;
    p.fR1.divmod((int)e, t, r1);
    p.fI1.divmod(t, o1, i1);
    const int terms = p.O2 * p.I2 * p.R2;
    float acc = 0.f;
#pragma unroll 16
    for (int q = lane; q < terms; q += 32) {
      int u, r2, i2, o2;
      p.fR2.divmod(q, u, r2);
      p.fI2.divmod(u, o2, i2);
      const float g = dT12[(((long long)(o1 * p.O2 + o2)) * I12 + i1 * p.I2 + i2) * p.R2 + r2];
      const float c = to_f32<T>(c2[(((long long)r1 * p.O2 + o2) * p.I2 + i2) * p.R2 + r2]);
      acc = fmaf(g, c, acc);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) d1[e] = from_f32<T>(acc);
  }
}

static int fill_btt3(const tlt_blocktt3_desc* d, Btt3Params* p, bool grads, size_t* smem_bytes) {
  TLT_REQUIRE(d != nullptr, "blocktt3: null descriptor");
  TLT_REQUIRE(d->dtype == TLT_F32 || d->dtype == TLT_BF16, "blocktt3: dtype must be f32 or bf16");
  for (int k = 0; k < 3; ++k) TLT_REQUIRE(d->rows[k] >= 1 && d->cols[k] >= 1, "blocktt3: mode sizes must be >= 1");
  TLT_REQUIRE(d->rank1 >= 1 && d->rank2 >= 1, "blocktt3: ranks must be >= 1");
  p->O1 = d->rows[0]; p->O2 = d->rows[1]; p->O3 = d->rows[2];
  p->I1 = d->cols[0]; p->I2 = d->cols[1]; p->I3 = d->cols[2];
  p->R1 = d->rank1; p->R2 = d->rank2;
  p->I = (long long)p->I1 * p->I2 * p->I3;
  p->O = (long long)p->O1 * p->O2 * p->O3;
  TLT_REQUIRE(p->I < (1LL << 31) && p->O < (1LL << 31), "blocktt3: matrix too large");
  p->R2r = btt3_round4(p->R2);
  p->I3p = btt3_round4(p->I3);
  p->I3s = (p->I3 & 1) ? p->I3 : p->I3 + 1;
  const long long o12 = (long long)p->O1 * p->O2;
  const int i12 = p->I1 * p->I2;
  // balanced o3 groups of at most kG3Max rows, then as many (i1,i2) chunks as it takes to reach ~2 blocks per SM
  p->n_g3 = (p->O3 + kG3Max - 1) / kG3Max;
  p->G3 = (p->O3 + p->n_g3 - 1) / p->n_g3;
  p->n_g3 = (p->O3 + p->G3 - 1) / p->G3;
  long long want = (296 + o12 * p->n_g3 - 1) / (o12 * p->n_g3);
  if (want < 1) want = 1;
  if (want > i12) want = i12;
  p->IC = (int)((i12 + want - 1) / want);
  const size_t limit = 200 * 1024;
  while (p->IC > 1 && (size_t)btt3_carve(*p, grads).total * 4 > limit) p->IC = (p->IC + 1) / 2;
  p->n_ic = (i12 + p->IC - 1) / p->IC;
  const size_t need = (size_t)btt3_carve(*p, grads).total * 4;
  if (need > limit) {
    set_error("blocktt3: core slices (%d x %d x %d x %d) do not fit shared memory; use the generic chain", p->R1, p->I2,
              p->R2, p->I1);
    return TLT_ERR_UNSUPPORTED;
  }
  TLT_REQUIRE(o12 < (1LL << 31) && p->n_g3 <= 65535 && p->n_ic <= 65535, "blocktt3: grid too large");
  TLT_REQUIRE(p->I * 8 < (1LL << 31) && (long long)p->R1 * p->O2 * p->I2 * p->R2 < (1LL << 24) &&
              (long long)p->O1 * p->I1 * p->R1 < (1LL << 24), "blocktt3: cores too large for the fused kernels");
  p->fI3 = FastDiv(p->I3); p->fI3p = FastDiv(p->I3p); p->fI3v = FastDiv(p->I3 / 4 > 0 ? p->I3 / 4 : 1);
  p->fG3 = FastDiv(p->G3); p->fR2r = FastDiv(p->R2r); p->fR2 = FastDiv(p->R2); p->fI2 = FastDiv(p->I2);
  p->fI1 = FastDiv(p->I1); p->fO2 = FastDiv(p->O2); p->fR1 = FastDiv(p->R1); p->fRow = FastDiv(p->I2 * p->R2);
  p->fRun = FastDiv(p->IC * p->I3); p->fRun4 = FastDiv(p->IC * p->I3 / 4 > 0 ? p->IC * p->I3 / 4 : 1);
  *smem_bytes = need;
  return TLT_OK;
}

// opt in to > 48 KB of dynamic shared memory, once per (kernel, high-water mark)
template <typename K>
static int set_smem(K kernel, size_t bytes) {
  return ensure_smem((const void*)kernel, bytes);          // per (device, function), see common.cuh
}

static int pow2_floor(int v) {
  int r = 1;
  while (r * 2 <= v) r *= 2;
  return r;
}

template <typename T>
static int launch_reconstruct(const Btt3Params& p, size_t smem, const void* c1, const void* c2, const void* c3, void* W,
                              cudaStream_t st) {
  dim3 grid(p.O1 * p.O2, p.n_g3, p.n_ic);
  int rc;
  const bool vec = (p.I3 % 4) == 0 && (((uintptr_t)W) % 16) == 0;
  if (vec) {
    if ((rc = set_smem(btt3_reconstruct_kernel<T, 4>, smem))) return rc;
    btt3_reconstruct_kernel<T, 4><<<grid, 256, smem, st>>>(p, (const T*)c1, (const T*)c2, (const T*)c3, (T*)W);
  } else {
    if ((rc = set_smem(btt3_reconstruct_kernel<T, 1>, smem))) return rc;
    btt3_reconstruct_kernel<T, 1><<<grid, 256, smem, st>>>(p, (const T*)c1, (const T*)c2, (const T*)c3, (T*)W);
  }
  return check_launch("btt3_reconstruct_kernel");
}

template <typename T>
static int launch_core_grads(const Btt3Params& p, size_t smem, const void* c1, const void* c2, const void* c3,
                             const float* dW, void* d1, void* d2, void* d3, float* dT12, float* dC3p, cudaStream_t st) {
  int rc;
  dim3 grid(p.O1 * p.O2, p.n_g3, p.n_ic);
  int KQ = pow2_floor(256 / (p.IC < 256 ? p.IC : 256));
  if (KQ > 8) KQ = 8;
  int KH = pow2_floor(256 / (p.G3 * p.I3 < 256 ? p.G3 * p.I3 : 256));
  if (KH > 4) KH = 4;
  if ((rc = set_smem(btt3_grad_a_kernel<T>, smem))) return rc;
  btt3_grad_a_kernel<T><<<grid, 256, smem, st>>>(p, KQ, KH, (const T*)c1, (const T*)c2, (const T*)c3, dW, dT12, dC3p);
  if ((rc = check_launch("btt3_grad_a_kernel"))) return rc;
  const long long n3 = (long long)p.R2 * p.O3 * p.I3, n2 = (long long)p.R1 * p.O2 * p.I2 * p.R2, n1 = (long long)p.O1 * p.I1 * p.R1;
  const int blocks_c3 = (int)((n3 + 31) / 32), blocks_c2 = (int)((n2 + 63) / 64), blocks_c1 = (int)((n1 + 7) / 8);
  btt3_grad_b_kernel<T><<<blocks_c3 + blocks_c2 + blocks_c1, 256, 0, st>>>(p, blocks_c3, blocks_c2, (const T*)c1, (const T*)c2,
                                                                           dT12, dC3p, (T*)d1, (T*)d2, (T*)d3);
  return check_launch("btt3_grad_b_kernel");
}

}  // namespace tlt

using namespace tlt;

extern "C" int tlt_blocktt3_reconstruct(const tlt_blocktt3_desc* d, const void* c1, const void* c2, const void* c3,
                                        void* W, void* stream) {
  Btt3Params p;
  size_t smem = 0;
  int rc = fill_btt3(d, &p, false, &smem);
  if (rc) return rc;
  TLT_REQUIRE(c1 && c2 && c3 && W, "tlt_blocktt3_reconstruct: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->dtype == TLT_BF16) return launch_reconstruct<__nv_bfloat16>(p, smem, c1, c2, c3, W, st);
  return launch_reconstruct<float>(p, smem, c1, c2, c3, W, st);
}

extern "C" size_t tlt_blocktt3_core_grads_workspace_size(const tlt_blocktt3_desc* d) {
  Btt3Params p;
  size_t smem = 0;
  if (fill_btt3(d, &p, true, &smem)) return 0;
  const size_t o12 = (size_t)p.O1 * p.O2;
  const size_t dt12 = o12 * p.I1 * p.I2 * p.R2;
  const size_t dc3p = o12 * p.n_ic * p.R2 * p.O3 * p.I3;
  return (dt12 + dc3p) * sizeof(float);
}

extern "C" int tlt_blocktt3_core_grads(const tlt_blocktt3_desc* d, const void* c1, const void* c2, const void* c3,
                                       const float* dW, void* d1, void* d2, void* d3, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  Btt3Params p;
  size_t smem = 0;
  int rc = fill_btt3(d, &p, true, &smem);
  if (rc) return rc;
  TLT_REQUIRE(c1 && c2 && c3 && dW && d1 && d2 && d3 && workspace, "tlt_blocktt3_core_grads: null pointer");
  TLT_REQUIRE(((uintptr_t)dW % 16) == 0, "tlt_blocktt3_core_grads: dW must be 16-byte aligned");
  if (workspace_bytes < tlt_blocktt3_core_grads_workspace_size(d)) {
    set_error("tlt_blocktt3_core_grads: workspace too small (%zu < %zu)", workspace_bytes, tlt_blocktt3_core_grads_workspace_size(d));
    return TLT_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n_dt12 = (size_t)p.O1 * p.O2 * p.I1 * p.I2 * p.R2;
  float* dT12 = (float*)workspace;
  float* dC3p = dT12 + n_dt12;
  TLT_CHECK_CUDA(cudaMemsetAsync(dT12, 0, n_dt12 * sizeof(float), st));
  if (d->dtype == TLT_BF16) return launch_core_grads<__nv_bfloat16>(p, smem, c1, c2, c3, dW, d1, d2, d3, dT12, dC3p, st);
  return launch_core_grads<float>(p, smem, c1, c2, c3, dW, d1, d2, d3, dT12, dC3p, st);
}
