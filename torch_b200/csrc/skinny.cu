// Skinny row transform  y[r, j] = sum_i a[r, i] K[j, i]   for many rows and small I, J (<= 64), and its gradients.
//
// Two reference call sites reduce to it:
//  * the core expansion of tucker_conv (tltorch/functional/convolution.py:160, tenalg.multi_mode_dot(core, factors[2:])):
//    (r0 r1) rows of prod(kernel ranks) values times the Kronecker product of the (k_j, rank_j) factors -- 150 544 rows
//    of 4 -> 9 values at BASELINE config 4;
//  * the spatial half of tucker_trl's input projection (functional/tensor_regression.py:40): (batch, channel) rows of
//    7 x 7 = 49 values times the Kronecker product of the two spatial factors (49 -> 16) -- 1 048 576 rows at config 4,
//    the one pass that reads the 103 MB activation.  Doing it FIRST shrinks the tensor 3x before the channel mode goes
//    to the tensor cores (pointwise_tc needs a spatial extent that is a multiple of 8: 16 is, 49 is not).
// Rows are short and odd-sized (98 bytes), so a block stages 128 consecutive rows -- one contiguous, 16-byte-aligned span of
// global memory -- in shared memory with 16-byte copies, a thread owns a row (matrix in shared memory as fp32, read as
// broadcast float4), and the outputs leave through shared memory the same way.  The (J x I) matrix gradient is reduced
// per tile by 4 x 4 register blocks over a quarter of the tile's rows each, then shared-memory and global atomics.
#include "common.cuh"

namespace tlt {

constexpr int kSkMax = 64;
constexpr int kSkRows = 128;

template <typename T>
__device__ __forceinline__ void copy_span(T* __restrict__ dst, const T* __restrict__ src, long long n_el, int tid, int nthreads) {
  // both 16-byte aligned; n_el elements
  const long long n16 = (n_el * (long long)sizeof(T)) >> 4;
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  for (long long v = tid; v < n16; v += nthreads) d4[v] = s4[v];
  const long long done = (n16 << 4) / (long long)sizeof(T);
  for (long long e = done + tid; e < n_el; e += nthreads) dst[e] = src[e];
}

// k_trans = 0: K given as (J, I);  1: K given as (I, J)  (used for the data gradient: da = g . K)
template <typename T>
__global__ void __launch_bounds__(kSkRows) skinny_fwd_kernel(const T* __restrict__ a, const T* __restrict__ k, T* __restrict__ y,
                                                             long long rows, int I, int J, int k_trans) {
  extern __shared__ __align__(16) unsigned char sk_smem[];
  const int IP = (I + 3) & ~3;
  float* ks = reinterpret_cast<float*>(sk_smem);                                  // [J][IP]
  T* xin = reinterpret_cast<T*>(sk_smem + (size_t)J * IP * 4);                   // [128][I]
  T* yout = xin + (((size_t)kSkRows * I + 7) & ~(size_t)7);                      // [128][J]
  for (int e = threadIdx.x; e < J * IP; e += blockDim.x) {
    const int j = e / IP, i = e - j * IP;
    ks[e] = i < I ? to_f32<T>(k_trans ? k[(long long)i * J + j] : k[(long long)j * I + i]) : 0.f;
  }
  const long long tiles = (rows + kSkRows - 1) / kSkRows;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long base = tile * kSkRows;
    const int nrows = (int)(rows - base < kSkRows ? rows - base : kSkRows);
    __syncthreads();                                                             // previous tile's stores have read yout; ks staged
    copy_span(xin, a + base * I, (long long)nrows * I, threadIdx.x, blockDim.x);
    __syncthreads();
    if ((int)threadIdx.x < nrows) {
      float av[kSkMax];
#pragma unroll
      for (int i = 0; i < kSkMax; ++i) av[i] = i < I ? to_f32<T>(xin[threadIdx.x * I + i]) : 0.f;
      for (int j = 0; j < J; ++j) {
        const float4* kr = reinterpret_cast<const float4*>(ks + j * IP);
        float acc = 0.f;
#pragma unroll
        for (int i4 = 0; i4 < kSkMax / 4; ++i4) {
          if (i4 * 4 < I) {
            const float4 kv = kr[i4];
            acc = fmaf(av[4 * i4], kv.x, acc);
            acc = fmaf(av[4 * i4 + 1], kv.y, acc);
            acc = fmaf(av[4 * i4 + 2], kv.z, acc);
            acc = fmaf(av[4 * i4 + 3], kv.w, acc);
          }
        }
        yout[threadIdx.x * J + j] = from_f32<T>(acc);
      }
    }
    __syncthreads();
    copy_span(y + base * J, yout, (long long)nrows * J, threadIdx.x, blockDim.x);
  }
}

// dk[j, i] += sum_r g[r, j] a[r, i]
template <typename T>
__global__ void __launch_bounds__(256) skinny_dk_kernel(const T* __restrict__ a, const T* __restrict__ g, float* __restrict__ dk,
                                                        long long rows, int I, int J) {
  extern __shared__ __align__(16) unsigned char sk_smem[];
  float* red = reinterpret_cast<float*>(sk_smem);                                 // [J][I]
  T* as_ = reinterpret_cast<T*>(sk_smem + (((size_t)J * I * 4 + 15) & ~(size_t)15));     // [128][I]
  T* gs_ = as_ + (((size_t)kSkRows * I + 7) & ~(size_t)7);                       // [128][J]
  for (int e = threadIdx.x; e < I * J; e += blockDim.x) red[e] = 0.f;
  const int nbi = (I + 3) >> 2, nbj = (J + 3) >> 2, nblk = nbi * nbj;            // <= 256 register blocks of 4 x 4
  const int grp = threadIdx.x >> 6, p = threadIdx.x & 63;
  float acc[4][16];
#pragma unroll
  for (int n = 0; n < 4; ++n)
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[n][e] = 0.f;
  const long long tiles = (rows + kSkRows - 1) / kSkRows;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long base = tile * kSkRows;
    const int nrows = (int)(rows - base < kSkRows ? rows - base : kSkRows);
    __syncthreads();
    copy_span(as_, a + base * I, (long long)nrows * I, threadIdx.x, blockDim.x);
    copy_span(gs_, g + base * J, (long long)nrows * J, threadIdx.x, blockDim.x);
    __syncthreads();
    const int r0 = grp * 32, r1 = min(r0 + 32, nrows);
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int blk = p + 64 * n;
      if (blk < nblk) {
        const int j0 = (blk / nbi) * 4, i0 = (blk - (blk / nbi) * nbi) * 4;
        for (int r = r0; r < r1; ++r) {
          float gv[4], av[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            gv[e] = j0 + e < J ? to_f32<T>(gs_[r * J + j0 + e]) : 0.f;
            av[e] = i0 + e < I ? to_f32<T>(as_[r * I + i0 + e]) : 0.f;
          }
#pragma unroll
          for (int e = 0; e < 16; ++e) acc[n][e] = fmaf(gv[e >> 2], av[e & 3], acc[n][e]);
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    const int blk = p + 64 * n;
    if (blk < nblk) {
      const int j0 = (blk / nbi) * 4, i0 = (blk - (blk / nbi) * nbi) * 4;
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int j = j0 + (e >> 2), i = i0 + (e & 3);
        if (j < J && i < I) atomicAdd(red + j * I + i, acc[n][e]);
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < I * J; e += blockDim.x) {
    const float v = red[e];
    if (v != 0.f) atomicAdd(dk + e, v);
  }
}

template <typename T>
static int launch_skinny_fwd(const void* a, const void* k, void* y, long long rows, int I, int J, int k_trans, cudaStream_t st) {
  const int IP = (I + 3) & ~3;
  const size_t smem = (size_t)J * IP * 4 + ((((size_t)kSkRows * I + 7) & ~(size_t)7) + (size_t)kSkRows * J + 8) * sizeof(T);
  TLT_ENSURE_SMEM((skinny_fwd_kernel<T>), smem);
  long long tiles = (rows + kSkRows - 1) / kSkRows;
  const long long cap = 148LL * 6;
  skinny_fwd_kernel<T><<<(unsigned)(tiles < cap ? tiles : cap), kSkRows, smem, st>>>((const T*)a, (const T*)k, (T*)y, rows, I, J, k_trans);
  return check_launch("skinny_fwd_kernel");
}

template <typename T>
static int launch_skinny_dk(const void* a, const void* g, float* dk, long long rows, int I, int J, cudaStream_t st) {
  const size_t smem = (((size_t)J * I * 4 + 15) & ~(size_t)15) + ((((size_t)kSkRows * I + 7) & ~(size_t)7) + (size_t)kSkRows * J + 8) * sizeof(T);
  TLT_ENSURE_SMEM((skinny_dk_kernel<T>), smem);
  long long tiles = (rows + kSkRows - 1) / kSkRows;
  const long long cap = 148LL * 2;
  skinny_dk_kernel<T><<<(unsigned)(tiles < cap ? tiles : cap), 256, smem, st>>>((const T*)a, (const T*)g, dk, rows, I, J);
  return check_launch("skinny_dk_kernel");
}

}  // namespace tlt

using namespace tlt;

extern "C" int tlt_skinny_fwd(int32_t dtype, int64_t rows, int32_t I, int32_t J, const void* a, const void* k, void* y, void* stream) {
  TLT_REQUIRE(a && k && y && rows >= 1, "tlt_skinny_fwd: bad arguments");
  TLT_REQUIRE(I >= 1 && I <= kSkMax && J >= 1 && J <= kSkMax, "tlt_skinny_fwd: I and J must lie in 1..64");
  TLT_REQUIRE(((uintptr_t)a % 16) == 0 && ((uintptr_t)y % 16) == 0, "tlt_skinny_fwd: tensors must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TLT_BF16) return launch_skinny_fwd<__nv_bfloat16>(a, k, y, rows, I, J, 0, st);
  if (dtype == TLT_F32) return launch_skinny_fwd<float>(a, k, y, rows, I, J, 0, st);
  set_error("tlt_skinny_fwd: unknown dtype");
  return TLT_ERR_INVALID;
}

extern "C" int tlt_skinny_bwd(int32_t dtype, int64_t rows, int32_t I, int32_t J, const void* a, const void* g, const void* k,
                              void* da, float* dk, void* stream) {
  TLT_REQUIRE(g && k && rows >= 1 && (da || dk), "tlt_skinny_bwd: bad arguments");
  TLT_REQUIRE(!dk || a, "tlt_skinny_bwd: the matrix gradient needs the forward input");
  TLT_REQUIRE(I >= 1 && I <= kSkMax && J >= 1 && J <= kSkMax, "tlt_skinny_bwd: I and J must lie in 1..64");
  TLT_REQUIRE(((uintptr_t)a % 16) == 0 && ((uintptr_t)g % 16) == 0 && ((uintptr_t)da % 16) == 0, "tlt_skinny_bwd: tensors must be 16-byte aligned");
  TLT_REQUIRE(dtype == TLT_BF16 || dtype == TLT_F32, "tlt_skinny_bwd: unknown dtype");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = TLT_OK;
  if (da) {          // da[r, i] = sum_j g[r, j] k[j, i]: the forward kernel on g with the matrix read transposed
    rc = dtype == TLT_BF16 ? launch_skinny_fwd<__nv_bfloat16>(g, k, da, rows, J, I, 1, st) : launch_skinny_fwd<float>(g, k, da, rows, J, I, 1, st);
    if (rc) return rc;
  }
  if (dk) rc = dtype == TLT_BF16 ? launch_skinny_dk<__nv_bfloat16>(a, g, dk, rows, I, J, st) : launch_skinny_dk<float>(a, g, dk, rows, I, J, st);
  return rc;
}

// =====================================================================================================================
// tlt_scale_modes: y[i_0..i_{n-1}] = alpha * prod_k s_k[i_k] * x[i_0..i_{n-1}]   (s_k == NULL: mode untouched)
// The mask application of tensor dropout in its fixed-shape form (tltorch/tensor_hooks/_tensor_dropout.py:56-142: the 0/1
// keep vectors of the dropped rank modes and the 1/(1-p)^ndim rescale, multiplied into the Tucker core / CP weights / TT
// cores in ONE pass instead of index_select gathers).
// =====================================================================================================================
namespace tlt {
struct ScaleModesParams {
  int ndim;
  int dims[6];
  const void* s[6];
  float alpha;
  long long total;
};

template <typename T>
__global__ void __launch_bounds__(256) scale_modes_kernel(const ScaleModesParams p, const T* __restrict__ x, T* __restrict__ y) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < p.total; e += (long long)gridDim.x * blockDim.x) {
    long long r = e;
    float f = p.alpha;
#pragma unroll
    for (int d = 5; d >= 0; --d) {
      if (d < p.ndim) {
        const int i = (int)(r % p.dims[d]);
        r /= p.dims[d];
        if (p.s[d] != nullptr) f *= to_f32<T>(static_cast<const T*>(p.s[d])[i]);
      }
    }
    y[e] = from_f32<T>(f * to_f32<T>(x[e]));
  }
}
}  // namespace tlt

extern "C" int tlt_scale_modes(int32_t dtype, int32_t ndim, const int64_t* dims, const void* x, const void* const* scales, float alpha,
                               void* y, void* stream) {
  TLT_REQUIRE(x && y && dims && scales && ndim >= 1 && ndim <= 6, "tlt_scale_modes: bad arguments (1..6 modes)");
  ScaleModesParams p{};
  p.ndim = ndim; p.alpha = alpha; p.total = 1;
  for (int k = 0; k < ndim; ++k) {
    TLT_REQUIRE(dims[k] >= 1 && dims[k] < (1LL << 31), "tlt_scale_modes: bad extent");
    p.dims[k] = (int)dims[k]; p.s[k] = scales[k]; p.total *= dims[k];
  }
  const long long blocks = (p.total + 255) / 256;
  const unsigned grid = (unsigned)(blocks < 148LL * 8 ? blocks : 148LL * 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TLT_BF16) scale_modes_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p, (const __nv_bfloat16*)x, (__nv_bfloat16*)y);
  else if (dtype == TLT_F32) scale_modes_kernel<float><<<grid, 256, 0, st>>>(p, (const float*)x, (float*)y);
  else { set_error("tlt_scale_modes: unknown dtype"); return TLT_ERR_INVALID; }
  return check_launch("scale_modes_kernel");
}
