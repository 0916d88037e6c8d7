// Skinny row transform  y[r, j] = sum_i a[r, i] K[j, i]   for many rows and tiny I, J (<= 32), and its gradients.
//
// This is the core expansion of the factorized convolutions: tucker_conv contracts the Tucker core with its kernel-mode
// factors before the core convolution (tltorch/functional/convolution.py:160, tenalg.multi_mode_dot(core, factors[2:])),
// i.e. (r0 r1) rows of prod(kernel ranks) values times the Kronecker product of the (k_j, rank_j) factors -- 150 544 rows
// of 4 -> 9 values at BASELINE config 4.  On the generic contraction kernel the two mode products and (worse) their
// factor gradients -- 36 outputs reduced over 150 544 rows -- took ~0.95 ms per step on 128-block grids; here a thread
// owns a row (the matrix sits in shared memory as fp32) and the (J x I) matrix gradient is reduced tile by tile through
// shared memory with one atomicAdd per entry per block.
#include "common.cuh"

namespace tlt {

constexpr int kSkinnyMax = 32;

template <typename T>
__global__ void __launch_bounds__(256) skinny_fwd_kernel(const T* __restrict__ a, const T* __restrict__ k, T* __restrict__ y,
                                                         long long rows, int I, int J) {
  __shared__ float ks[kSkinnyMax * kSkinnyMax];
  for (int e = threadIdx.x; e < I * J; e += blockDim.x) ks[e] = to_f32<T>(k[e]);
  __syncthreads();
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    float av[kSkinnyMax];
#pragma unroll
    for (int i = 0; i < kSkinnyMax; ++i) av[i] = i < I ? to_f32<T>(a[r * I + i]) : 0.f;
    for (int j = 0; j < J; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < kSkinnyMax; ++i)
        if (i < I) acc = fmaf(av[i], ks[j * I + i], acc);
      y[r * J + j] = from_f32<T>(acc);
    }
  }
}

// da[r, i] = sum_j g[r, j] K[j, i]      dk[j, i] += sum_r g[r, j] a[r, i]
template <typename T>
__global__ void __launch_bounds__(256) skinny_bwd_kernel(const T* __restrict__ a, const T* __restrict__ g, const T* __restrict__ k,
                                                         T* __restrict__ da, float* __restrict__ dk, long long rows, int I, int J) {
  extern __shared__ float sm[];
  float* ks = sm;                                  // [J][I]
  float* as_ = ks + kSkinnyMax * kSkinnyMax;       // [256][I + 1]
  float* gs_ = as_ + 256 * (I + 1);                // [256][J + 1]
  for (int e = threadIdx.x; e < I * J; e += blockDim.x) ks[e] = to_f32<T>(k[e]);
  const int pairs = I * J;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};             // this thread's entries e = tid + 256 * n of dk
  __syncthreads();
  const long long tiles = (rows + 255) / 256;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long r = tile * 256 + threadIdx.x;
    const bool live = r < rows;
    float gv[kSkinnyMax];
#pragma unroll
    for (int j = 0; j < kSkinnyMax; ++j) {
      gv[j] = (live && j < J) ? to_f32<T>(g[r * J + j]) : 0.f;
      if (j < J) gs_[threadIdx.x * (J + 1) + j] = gv[j];
    }
    for (int i = 0; i < I; ++i) {
      if (dk != nullptr) as_[threadIdx.x * (I + 1) + i] = live ? to_f32<T>(a[r * I + i]) : 0.f;
      if (da != nullptr && live) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < kSkinnyMax; ++j)
          if (j < J) s = fmaf(gv[j], ks[j * I + i], s);
        da[r * I + i] = from_f32<T>(s);
      }
    }
    if (dk != nullptr) {
      __syncthreads();
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const int e = threadIdx.x + 256 * n;
        if (e < pairs) {
          const int j = e / I, i = e - j * I;
          float s = 0.f;
          for (int rr = 0; rr < 256; ++rr) s = fmaf(gs_[rr * (J + 1) + j], as_[rr * (I + 1) + i], s);
          acc[n] += s;
        }
      }
      __syncthreads();
    }
  }
  if (dk != nullptr) {
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int e = threadIdx.x + 256 * n;
      if (e < pairs && acc[n] != 0.f) atomicAdd(dk + e, acc[n]);
    }
  }
}

static unsigned skinny_grid(long long rows) {
  long long b = (rows + 255) / 256;
  const long long cap = 148LL * 4;
  return (unsigned)(b < 1 ? 1 : (b < cap ? b : cap));
}

}  // namespace tlt

using namespace tlt;

extern "C" int tlt_skinny_fwd(int32_t dtype, int64_t rows, int32_t I, int32_t J, const void* a, const void* k, void* y, void* stream) {
  TLT_REQUIRE(a && k && y && rows >= 1, "tlt_skinny_fwd: bad arguments");
  TLT_REQUIRE(I >= 1 && I <= kSkinnyMax && J >= 1 && J <= kSkinnyMax, "tlt_skinny_fwd: I and J must lie in 1..32");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TLT_BF16)
    skinny_fwd_kernel<__nv_bfloat16><<<skinny_grid(rows), 256, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)k, (__nv_bfloat16*)y, rows, I, J);
  else if (dtype == TLT_F32)
    skinny_fwd_kernel<float><<<skinny_grid(rows), 256, 0, st>>>((const float*)a, (const float*)k, (float*)y, rows, I, J);
  else { set_error("tlt_skinny_fwd: unknown dtype"); return TLT_ERR_INVALID; }
  return check_launch("skinny_fwd_kernel");
}

extern "C" int tlt_skinny_bwd(int32_t dtype, int64_t rows, int32_t I, int32_t J, const void* a, const void* g, const void* k,
                              void* da, float* dk, void* stream) {
  TLT_REQUIRE(g && k && rows >= 1 && (da || dk), "tlt_skinny_bwd: bad arguments");
  TLT_REQUIRE(!dk || a, "tlt_skinny_bwd: the matrix gradient needs the forward input");
  TLT_REQUIRE(I >= 1 && I <= kSkinnyMax && J >= 1 && J <= kSkinnyMax, "tlt_skinny_bwd: I and J must lie in 1..32");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)(kSkinnyMax * kSkinnyMax + 256 * (I + 1) + 256 * (J + 1)) * sizeof(float);
  if (dtype == TLT_BF16) {
    if (smem > 48 * 1024) TLT_CHECK_CUDA(cudaFuncSetAttribute(skinny_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    skinny_bwd_kernel<__nv_bfloat16><<<skinny_grid(rows), 256, smem, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)g, (const __nv_bfloat16*)k,
                                                                         (__nv_bfloat16*)da, dk, rows, I, J);
  } else if (dtype == TLT_F32) {
    if (smem > 48 * 1024) TLT_CHECK_CUDA(cudaFuncSetAttribute(skinny_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    skinny_bwd_kernel<float><<<skinny_grid(rows), 256, smem, st>>>((const float*)a, (const float*)g, (const float*)k, (float*)da, dk, rows, I, J);
  } else { set_error("tlt_skinny_bwd: unknown dtype"); return TLT_ERR_INVALID; }
  return check_launch("skinny_bwd_kernel");
}
