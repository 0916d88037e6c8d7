// fp32 -> 3-way bf16 split operands for "fp32 GEMMs on the bf16 tensor cores".
//
//   x = hi + mid + lo   with hi = bf16(x), mid = bf16(x - hi), lo = bf16(x - hi - mid)      (24 mantissa bits in three 8-bit pieces)
//   a . b ~= hi.hi + hi.mid + mid.hi + hi.lo + mid.mid + lo.hi                               (dropped terms <= 2^-24 |a||b|)
//
// The six partial products become ONE tcgen05 GEMM whose reduction dimension is six blocks long (fp32 accumulation in TMEM): the
// A-side operand repeats its pieces as [hi hi mid hi mid lo], the B-side operand as [hi mid hi lo mid hi].  Measured against fp64
// on a B200 (tools/scratch/split_gemm_test.py): 5e-7 .. 1.5e-6 relative Frobenius error for K = 128 .. 2344, the same class as an
// fp32 FFMA GEMM (2e-7 .. 4e-7) and 7x inside the layer's 1e-5 tolerance; three blocks (hi.hi + hi.mid + mid.hi) give 4.5e-6.
// This is how the CP FactorizedLinear of BASELINE config 1 (fp32, tltorch/functional/factorized_linear.py:23-36) reaches the tensor
// cores.  One launch writes the split operand in either or both layouts a GEMM may want:
//   "cols": (rows, 6 * colsP) -- the blocks side by side along the reduction index of a K-major operand
//   "rows": (6 * rows, colsP) -- the blocks stacked along the reduction index of an MN-major operand
// colsP = cols rounded up to 8 (16-byte TMA pitch); pad columns are written as zeros.
#include "common.cuh"

namespace tlt {
namespace sp3 {

__device__ __forceinline__ void split(float x, __nv_bfloat16 (&p)[3]) {
  p[0] = __float2bfloat16_rn(x);
  float r = x - __bfloat162float(p[0]);
  p[1] = __float2bfloat16_rn(r);
  r -= __bfloat162float(p[1]);
  p[2] = __float2bfloat16_rn(r);
}
// piece index of block b: pattern 0 (A side) hi hi mid hi mid lo, pattern 1 (B side) hi mid hi lo mid hi
__device__ __forceinline__ int piece(int pattern, int b) {
  const int pa = (0) | (0 << 2) | (1 << 4) | (0 << 6) | (1 << 8) | (2 << 10);
  const int pb = (0) | (1 << 2) | (0 << 4) | (2 << 6) | (1 << 8) | (0 << 10);
  return ((pattern ? pb : pa) >> (2 * b)) & 3;
}

__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ src, long long rows, int cols, long long ld, int colsP,
                                                     int pat_cols, __nv_bfloat162* __restrict__ out_cols, int pat_rows,
                                                     __nv_bfloat162* __restrict__ out_rows) {
  const int half = colsP / 2;
  const long long total = rows * half;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / half;
    const int c = (int)(e - r * half) * 2;
    const float x0 = c < cols ? __ldg(src + r * ld + c) : 0.f;
    const float x1 = c + 1 < cols ? __ldg(src + r * ld + c + 1) : 0.f;
    __nv_bfloat16 p0[3], p1[3];
    split(x0, p0);
    split(x1, p1);
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      if (out_cols) { const int q = piece(pat_cols, b); out_cols[(r * 6 * colsP + (long long)b * colsP + c) / 2] = __halves2bfloat162(p0[q], p1[q]); }
      if (out_rows) { const int q = piece(pat_rows, b); out_rows[(((long long)b * rows + r) * colsP + c) / 2] = __halves2bfloat162(p0[q], p1[q]); }
    }
  }
}

}  // namespace sp3
}  // namespace tlt

using namespace tlt;

extern "C" int tlt_split3_bf16(const float* src, int64_t rows, int64_t cols, int64_t ld, int32_t pattern_cols, void* out_cols,
                               int32_t pattern_rows, void* out_rows, void* stream) {
  TLT_REQUIRE(src && rows >= 1 && cols >= 1 && ld >= cols && cols < (1LL << 28) && rows < (1LL << 31) && (out_cols || out_rows),
              "tlt_split3_bf16: bad arguments");
  TLT_REQUIRE((pattern_cols == 0 || pattern_cols == 1) && (pattern_rows == 0 || pattern_rows == 1), "tlt_split3_bf16: pattern is 0 (A side) or 1 (B side)");
  const int colsP = (int)((cols + 7) & ~(int64_t)7);
  const long long total = rows * (colsP / 2);
  const long long blocks = (total + 255) / 256;
  const unsigned grid = (unsigned)(blocks < (long long)device_sms() * 16 ? blocks : (long long)device_sms() * 16);
  sp3::split3_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, rows, (int)cols, ld, colsP, pattern_cols, (__nv_bfloat162*)out_cols, pattern_rows,
                                                            (__nv_bfloat162*)out_rows);
  return check_launch("split3_kernel");
}
