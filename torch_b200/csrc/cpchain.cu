// Operand preparation and gradient finalisation of the UNFUSED CP-convolution chain on channels-last rows
//     x -> [GEMM: C_in -> R] -> separable depthwise sweep (dwsep_rows.cu) -> [GEMM: R -> C_out] -> y
// (cp_conv, tltorch/functional/convolution.py:231-283).  The chain used to be assembled from generic ops, each of which re-pitched
// its weight, zero-filled pad rows, cast fp32 weight gradients back to bf16 and sliced pad columns off with ATen kernels:
// ~10 small launches per layer and direction, 0.5 ms of config 3's step.  Here ONE launch writes every operand the chain's GEMMs
// and sweeps read, and ONE launch turns every fp32 partial result into the five parameter gradients.
#include "common.cuh"

namespace tlt {
namespace cpc {

// wp_in [ld][cin_p] = F_in^T, wp_out [cout_p][ld] = F_out (zero pad rows / columns), taps_f / taps_d [6][ld] fp32
// (kh0 kh1 kh2, lambda kw0 .. kw2; taps_d reversed when the data gradient is the same sweep, i.e. stride 1)
__global__ void prep_kernel(const __nv_bfloat16* __restrict__ f_in, const __nv_bfloat16* __restrict__ f_out, const __nv_bfloat16* __restrict__ kh,
                            const __nv_bfloat16* __restrict__ kw, const __nv_bfloat16* __restrict__ lam, int c_in, int c_out, int R,
                            int cin_p, int cout_p, int ld, int flip_d, __nv_bfloat16* __restrict__ wp_in, __nv_bfloat16* __restrict__ wp_out,
                            float* __restrict__ taps_f, float* __restrict__ taps_d) {
  const int n_in = ld * cin_p, n_out = cout_p * ld, n_t = 6 * ld;
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_in + n_out + 2 * n_t; i += gridDim.x * blockDim.x) {
    int j = i;
    if (j < n_in) { const int r = j / cin_p, c = j - r * cin_p; wp_in[j] = (r < R && c < c_in) ? f_in[c * R + r] : zero; continue; }
    j -= n_in;
    if (j < n_out) { const int c = j / ld, r = j - c * ld; wp_out[j] = (r < R && c < c_out) ? f_out[c * R + r] : zero; continue; }
    j -= n_out;
    const int dir = j / n_t, k = j - dir * n_t, t = k / ld, r = k - t * ld;
    float v = 0.f;
    if (r < R) {
      const int tt = (dir == 1 && flip_d) ? (t < 3 ? 2 - t : 8 - t) : t;
      v = tt < 3 ? __bfloat162float(kh[tt * R + r]) : __bfloat162float(kw[(tt - 3) * R + r]) * __bfloat162float(lam[r]);
    }
    (dir == 0 ? taps_f : taps_d)[k] = v;
  }
}

// blocks [0, tap_blocks): 8 rank channels x 32 slab lanes -> gradients of kh, kw, lambda from the sweep's partial sums
// blocks [tap_blocks, ..): element-wise  g_fout (C_out, R) <- dW_out [C_out][ld] fp32,  g_fin (C_in, R) <- dW_in [C_in][ld] fp32
__global__ void __launch_bounds__(256) post_kernel(const float* __restrict__ dw_out, const float* __restrict__ dw_in, const float* __restrict__ dT,
                                                   int parts, const __nv_bfloat16* __restrict__ kh, const __nv_bfloat16* __restrict__ kw,
                                                   const __nv_bfloat16* __restrict__ lam, int c_in, int c_out, int R, int cin_p, int ld,
                                                   int tap_blocks, __nv_bfloat16* __restrict__ g_fout, __nv_bfloat16* __restrict__ g_fin,
                                                   __nv_bfloat16* __restrict__ g_kh, __nv_bfloat16* __restrict__ g_kw,
                                                   __nv_bfloat16* __restrict__ g_lam) {
  __shared__ float red[9][32][8];
  if ((int)blockIdx.x >= tap_blocks) {
    // a thread converts 8 consecutive rank entries of one row: two 16-byte loads, eight 2-byte stores (rows of the (C, R) gradients
    // start at odd element offsets when R is odd); neighbouring threads cover neighbouring 16-byte spans
    const int cpr = ld >> 3, total = (c_out + c_in) * cpr;
    for (int i = (blockIdx.x - tap_blocks) * blockDim.x + threadIdx.x; i < total; i += (gridDim.x - tap_blocks) * blockDim.x) {
      int row = i / cpr;
      const int r0 = (i - row * cpr) * 8;
      const float* src;
      __nv_bfloat16* dst;
      if (row < c_out) { src = dw_out + (size_t)row * ld + r0; dst = g_fout + (size_t)row * R + r0; }
      else { row -= c_out; src = dw_in + (size_t)row * ld + r0; dst = g_fin + (size_t)row * R + r0; }
      const float4 a = __ldg((const float4*)src), b = __ldg((const float4*)src + 1);
      const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (r0 + e < R) dst[e] = __float2bfloat16_rn(v[e]);
    }
    return;
  }
  const int ch = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int r = blockIdx.x * 8 + ch;
  float t[9];
#pragma unroll
  for (int a = 0; a < 9; ++a) t[a] = 0.f;
  // 5 x 9 independent loads in flight per thread (148 slabs = one trip): walked one slab at a time the block spent ~1 us of L2 latency
  // per trip and the launch took 11 us inside the graph (gpurun_out/r4b/timeline.json)
  if (r < R)
    for (int p0 = pl; p0 < parts; p0 += 32 * 5) {
      float v[5][9];
#pragma unroll
      for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int a = 0; a < 9; ++a) v[i][a] = p0 + 32 * i < parts ? __ldg(dT + ((size_t)(p0 + 32 * i) * 9 + a) * ld + r) : 0.f;
#pragma unroll
      for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int a = 0; a < 9; ++a) t[a] += v[i][a];
    }
#pragma unroll
  for (int a = 0; a < 9; ++a) red[a][pl][ch] = t[a];
  __syncthreads();
  for (int s = 16; s >= 1; s >>= 1) {
    if (pl < s)
#pragma unroll
      for (int a = 0; a < 9; ++a) red[a][pl][ch] += red[a][pl + s][ch];
    __syncthreads();
  }
  if (pl == 0 && r < R) {
    float h[3], w[3];
    const float l = __bfloat162float(lam[r]);
#pragma unroll
    for (int a = 0; a < 3; ++a) { h[a] = __bfloat162float(kh[a * R + r]); w[a] = __bfloat162float(kw[a * R + r]); }
#pragma unroll
    for (int a = 0; a < 9; ++a) t[a] = red[a][0][ch];
    float gl = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float gh = l * (w[0] * t[3 * a] + w[1] * t[3 * a + 1] + w[2] * t[3 * a + 2]);
      const float gwl = h[0] * t[a] + h[1] * t[3 + a] + h[2] * t[6 + a];
      g_kh[a * R + r] = __float2bfloat16_rn(gh);
      g_kw[a * R + r] = __float2bfloat16_rn(l * gwl);
      gl += w[a] * gwl;
    }
    g_lam[r] = __float2bfloat16_rn(gl);
  }
}

}  // namespace cpc
}  // namespace tlt

using namespace tlt;

static inline int r8(int64_t v) { return (int)((v + 7) & ~(int64_t)7); }

extern "C" int tlt_cpchain_prep(int64_t c_in, int64_t c_out, int64_t rank, int32_t flip_dgrad_taps, const void* f_in, const void* f_out,
                                const void* kh, const void* kw, const void* lam, void* wp_in, void* wp_out, float* taps_f, float* taps_d,
                                void* stream) {
  TLT_REQUIRE(c_in >= 1 && c_out >= 1 && rank >= 1 && f_in && f_out && kh && kw && lam && wp_in && wp_out && taps_f && taps_d,
              "tlt_cpchain_prep: bad arguments");
  const int cin_p = r8(c_in), cout_p = r8(c_out), ld = r8(rank);
  const long long total = (long long)ld * cin_p + (long long)cout_p * ld + 12LL * ld;
  const unsigned grid = (unsigned)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  cpc::prep_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)f_in, (const __nv_bfloat16*)f_out, (const __nv_bfloat16*)kh,
                                                          (const __nv_bfloat16*)kw, (const __nv_bfloat16*)lam, (int)c_in, (int)c_out, (int)rank,
                                                          cin_p, cout_p, ld, flip_dgrad_taps, (__nv_bfloat16*)wp_in, (__nv_bfloat16*)wp_out,
                                                          taps_f, taps_d);
  return check_launch("cpchain prep_kernel");
}

extern "C" int tlt_cpchain_post(int64_t c_in, int64_t c_out, int64_t rank, const float* dw_out, const float* dw_in, const float* acc,
                                int32_t parts, const void* kh, const void* kw, const void* lam, void* g_fout, void* g_fin, void* g_kh,
                                void* g_kw, void* g_lam, void* stream) {
  TLT_REQUIRE(c_in >= 1 && c_out >= 1 && rank >= 1 && parts >= 1 && dw_out && dw_in && acc && kh && kw && lam && g_fout && g_fin && g_kh &&
              g_kw && g_lam, "tlt_cpchain_post: bad arguments");
  const int cin_p = r8(c_in), ld = r8(rank);
  const int tap_blocks = (int)((rank + 7) / 8);
  const long long vecs = (long long)(c_in + c_out) * (ld / 8);
  const int ew_blocks = (int)((vecs + 255) / 256 < 592 ? (vecs + 255) / 256 : 592);
  cpc::post_kernel<<<tap_blocks + ew_blocks, 256, 0, (cudaStream_t)stream>>>(dw_out, dw_in, acc, parts, (const __nv_bfloat16*)kh,
                                                                            (const __nv_bfloat16*)kw, (const __nv_bfloat16*)lam, (int)c_in,
                                                                            (int)c_out, (int)rank, cin_p, ld, tap_blocks,
                                                                            (__nv_bfloat16*)g_fout, (__nv_bfloat16*)g_fin,
                                                                            (__nv_bfloat16*)g_kh, (__nv_bfloat16*)g_kw, (__nv_bfloat16*)g_lam);
  return check_launch("cpchain post_kernel");
}
