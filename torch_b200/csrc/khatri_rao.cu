// Weighted Khatri-Rao product of up to 4 factor matrices and its gradients.
//
//   KR[(i_1 .. i_n), r] = w_r * prod_k F_k[i_k, r]          F_k (d_k, R) row-major, w (R) optional, KR (prod d_k, R)
//
// This is how the CP paths meet their dense operand: tensor_dot_cp / linear_cp contract the activation with
// KhatriRao(in-mode factors) and expand with lambda * KhatriRao(out-mode factors) (tltorch/functional/
// factorized_tensordot.py:70-103, the n-ary einsum evaluated in FLOP-sane order), and
// CPTensor.to_tensor builds the same product (factorized_tensors.py:129-130).  As pairwise contractions the rank index is a
// BATCH index: every step launched R = 2341 blocks (one 64 x 64 tile each for 4 x 8 useful outputs) and cost 43-57 us, 16
// such launches per step at BASELINE config 1; the lambda scaling and its gradient added 140 / 387 us more.  Element-wise
// the whole product is 1.2 M multiplies: one launch forward, one launch per factor backward (thread per (i_k, r), reads
// coalesced along r, no atomics except the n-term lambda reduction).
#include "common.cuh"

namespace tlt {

struct KrParams {
  int n, R;
  int dims[4];
  const void* f[4];
  const void* w;
  long long rows;
  long long ldg;            // backward: leading dimension of g (R when dense)
};

template <typename T>
__global__ void __launch_bounds__(256) khatri_rao_fwd_kernel(const KrParams p, T* __restrict__ out) {
  const long long total = p.rows * p.R;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e % p.R);
    long long row = e / p.R;
    float v = p.w ? to_f32<T>(static_cast<const T*>(p.w)[r]) : 1.f;
#pragma unroll
    for (int k = 3; k >= 0; --k) {
      if (k < p.n) {
        const int i = (int)(row % p.dims[k]);
        row /= p.dims[k];
        v *= to_f32<T>(static_cast<const T*>(p.f[k])[(long long)i * p.R + r]);
      }
    }
    out[e] = from_f32<T>(v);
  }
}

// dF_k[i, r] += w_r * sum_{rows with i_k = i} g[row, r] * prod_{j != k} F_j[i_j, r];   k == 0 also accumulates
// dw_r += sum_i F_0[i, r] * (the unweighted sum).   grid = (ceil(R / 128), d_k, splits of the other indices); dF zeroed.
template <typename T>
__global__ void __launch_bounds__(128) khatri_rao_bwd_kernel(const KrParams p, int k, const T* __restrict__ g, float* __restrict__ dF,
                                                             float* __restrict__ dw, int others, int per_split) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int ik = blockIdx.y;
  if (r >= p.R) return;
  const int o0 = blockIdx.z * per_split, o1 = min(others, o0 + per_split);
  float acc = 0.f;
  for (int o = o0; o < o1; ++o) {
    // decode the other indices (last factor fastest) and rebuild the row index
    int rem = o, row = 0, mul = 1;
    float v = 1.f;
#pragma unroll
    for (int j = 3; j >= 0; --j) {
      if (j < p.n) {
        int i;
        if (j == k) i = ik;
        else {
          const int q = rem / p.dims[j];
          i = rem - q * p.dims[j];
          rem = q;
          v *= to_f32<T>(static_cast<const T*>(p.f[j])[(long long)i * p.R + r]);
        }
        row += i * mul;
        mul *= p.dims[j];
      }
    }
    acc = fmaf(to_f32<T>(g[(long long)row * p.ldg + r]), v, acc);
  }
  const float wr = p.w ? to_f32<T>(static_cast<const T*>(p.w)[r]) : 1.f;
  atomicAdd(dF + (long long)ik * p.R + r, wr * acc);
  if (k == 0 && dw != nullptr) atomicAdd(dw + r, acc * to_f32<T>(static_cast<const T*>(p.f[0])[(long long)ik * p.R + r]));
}

// All factor gradients of one Khatri-Rao product in ONE launch: blockIdx.y walks the rows of every factor in turn (the launch per
// factor re-read g once each: 3 launches of ~12 us per side at BASELINE config 1).  dF[k] zeroed by the caller.
struct KrGrads { float* dF[4]; };
template <typename T>
__global__ void __launch_bounds__(128) khatri_rao_bwd_all_kernel(const KrParams p, const T* __restrict__ g, const KrGrads out,
                                                                 float* __restrict__ dw) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  int ik = blockIdx.y, k = 0;
  while (k < p.n - 1 && ik >= p.dims[k]) { ik -= p.dims[k]; ++k; }
  if (r >= p.R || out.dF[k] == nullptr) return;
  const int others = (int)(p.rows / p.dims[k]);
  const int per_split = (others + gridDim.z - 1) / gridDim.z;
  const int o0 = blockIdx.z * per_split, o1 = min(others, o0 + per_split);
  if (o0 >= o1) return;
  // the other indices of row o0 are decoded once and then counted up (last factor fastest); a factor entry is re-read only when
  // its index moves.  Decoding every row with two runtime divisions made this launch 30 us at BASELINE config 1 (3.6 M terms).
  int idx[4], mul[4];
  float fv[4];
  {
    int rem = o0, m = 1;
#pragma unroll
    for (int j = 3; j >= 0; --j) {
      idx[j] = 0; mul[j] = 0; fv[j] = 1.f;
      if (j < p.n) {
        mul[j] = m;
        m *= p.dims[j];
        if (j == k) idx[j] = ik;
        else {
          const int q = rem / p.dims[j];
          idx[j] = rem - q * p.dims[j];
          rem = q;
          fv[j] = to_f32<T>(static_cast<const T*>(p.f[j])[(long long)idx[j] * p.R + r]);
        }
      }
    }
  }
  float acc = 0.f;
  for (int o = o0; o < o1; ++o) {
    const int row = idx[0] * mul[0] + idx[1] * mul[1] + idx[2] * mul[2] + idx[3] * mul[3];
    acc = fmaf(to_f32<T>(g[(long long)row * p.ldg + r]), (fv[0] * fv[1]) * (fv[2] * fv[3]), acc);
    bool carry = true;
#pragma unroll
    for (int j = 3; j >= 0; --j) {
      if (j < p.n && j != k && carry) {
        if (++idx[j] == p.dims[j]) idx[j] = 0; else carry = false;
        fv[j] = to_f32<T>(static_cast<const T*>(p.f[j])[(long long)idx[j] * p.R + r]);
      }
    }
  }
  const float wr = p.w ? to_f32<T>(static_cast<const T*>(p.w)[r]) : 1.f;
  atomicAdd(out.dF[k] + (long long)ik * p.R + r, wr * acc);
  if (k == 0 && dw != nullptr) atomicAdd(dw + r, acc * to_f32<T>(static_cast<const T*>(p.f[0])[(long long)ik * p.R + r]));
}

// The same gradients in ONE pass over g: a thread owns a rank column r and a run of PREFIXES (the index tuple of all factors but the
// last); per prefix it loads the d_last entries of g that share it (independent loads, all in flight), keeps the last factor's
// column and its gradient in registers, and carries one running sum per earlier factor that is flushed (one atomic) when that
// factor's index moves.  g is read once instead of once per factor, and nothing is decoded by division inside the loop: 30 -> ~5 us
// per launch at BASELINE config 1 (dims (4, 8, 16), R = 2341).  Last dimensions above MAXD keep the kernel above.
template <typename T, int MAXD>
__global__ void __launch_bounds__(128) khatri_rao_bwd_fused_kernel(const KrParams p, const T* __restrict__ g, const KrGrads out,
                                                                   float* __restrict__ dw, int prefixes, int per_split) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= p.R) return;
  const int n = p.n, dl = p.dims[n - 1];
  const int o0 = blockIdx.y * per_split, o1 = min(prefixes, o0 + per_split);
  if (o0 >= o1) return;
  const float wr = p.w ? to_f32<T>(static_cast<const T*>(p.w)[r]) : 1.f;
  const T* fl_ptr = static_cast<const T*>(p.f[n - 1]) + r;
  float fl[MAXD], dlast[MAXD];
#pragma unroll
  for (int i = 0; i < MAXD; ++i) { fl[i] = i < dl ? to_f32<T>(fl_ptr[(long long)i * p.R]) : 0.f; dlast[i] = 0.f; }
  int idx[3] = {0, 0, 0};
  float fv[3] = {1.f, 1.f, 1.f}, acc[3] = {0.f, 0.f, 0.f};
  {
    int rem = o0;
#pragma unroll
    for (int j = 2; j >= 0; --j)
      if (j < n - 1) {
        const int q = rem / p.dims[j];
        idx[j] = rem - q * p.dims[j];
        rem = q;
        fv[j] = to_f32<T>(static_cast<const T*>(p.f[j])[(long long)idx[j] * p.R + r]);
      }
  }
  auto flush = [&](int j) {
    if (out.dF[j]) atomicAdd(out.dF[j] + (long long)idx[j] * p.R + r, wr * acc[j]);
    if (j == 0 && dw) atomicAdd(dw + r, acc[0] * fv[0]);
    acc[j] = 0.f;
  };
  for (int o = o0; o < o1; ++o) {
    const T* gp = g + (long long)o * dl * p.ldg + r;
    float gv[MAXD];
#pragma unroll
    for (int i = 0; i < MAXD; ++i) gv[i] = i < dl ? to_f32<T>(gp[(long long)i * p.ldg]) : 0.f;
    const float pre = fv[0] * fv[1] * fv[2];
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < MAXD; ++i) { t = fmaf(gv[i], fl[i], t); dlast[i] = fmaf(gv[i], pre, dlast[i]); }
    acc[0] = fmaf(t, fv[1] * fv[2], acc[0]);
    acc[1] = fmaf(t, fv[0] * fv[2], acc[1]);
    acc[2] = fmaf(t, fv[0] * fv[1], acc[2]);
    const bool last = o + 1 == o1;
    bool carry = true;
#pragma unroll
    for (int j = 2; j >= 0; --j)
      if (j < n - 1 && (carry || last)) {
        flush(j);
        if (carry) {
          if (++idx[j] == p.dims[j]) idx[j] = 0; else carry = false;
          if (!last) fv[j] = to_f32<T>(static_cast<const T*>(p.f[j])[(long long)idx[j] * p.R + r]);
        }
      }
  }
  if (out.dF[n - 1]) {
#pragma unroll
    for (int i = 0; i < MAXD; ++i)
      if (i < dl) atomicAdd(out.dF[n - 1] + (long long)i * p.R + r, wr * dlast[i]);
  }
  if (n == 1 && dw) {
    float sacc = 0.f;
#pragma unroll
    for (int i = 0; i < MAXD; ++i) sacc = fmaf(dlast[i], fl[i], sacc);
    atomicAdd(dw + r, sacc);
  }
}

// Khatri-Rao product written directly as 3-way bf16 split operands (see split3.cu): no fp32 matrix in between.
__device__ __forceinline__ void kr_split3(float x, __nv_bfloat16 (&pz)[3]) {
  pz[0] = __float2bfloat16_rn(x);
  float r = x - __bfloat162float(pz[0]);
  pz[1] = __float2bfloat16_rn(r);
  r -= __bfloat162float(pz[1]);
  pz[2] = __float2bfloat16_rn(r);
}
__global__ void __launch_bounds__(256) khatri_rao_split3_kernel(const KrParams p, int colsP, int pat_cols, __nv_bfloat162* __restrict__ out_cols,
                                                                int pat_rows, __nv_bfloat162* __restrict__ out_rows) {
  const int half = colsP / 2;
  const long long total = p.rows * half;
  const int pa = (0) | (0 << 2) | (1 << 4) | (0 << 6) | (1 << 8) | (2 << 10), pb = (0) | (1 << 2) | (0 << 4) | (2 << 6) | (1 << 8) | (0 << 10);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long row = e / half;
    const int c = (int)(e - row * half) * 2;
    float v0 = 0.f, v1 = 0.f;
    if (c < p.R) {
      const bool two = c + 1 < p.R;
      v0 = p.w ? static_cast<const float*>(p.w)[c] : 1.f;
      v1 = two ? (p.w ? static_cast<const float*>(p.w)[c + 1] : 1.f) : 0.f;
      long long rem = row;
#pragma unroll
      for (int k = 3; k >= 0; --k) {
        if (k < p.n) {
          const int i = (int)(rem % p.dims[k]);
          rem /= p.dims[k];
          const float* f = static_cast<const float*>(p.f[k]) + (long long)i * p.R + c;
          v0 *= f[0];
          if (two) v1 *= f[1];
        }
      }
    }
    __nv_bfloat16 p0[3], p1[3];
    kr_split3(v0, p0);
    kr_split3(v1, p1);
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      if (out_cols) { const int q = ((pat_cols ? pb : pa) >> (2 * b)) & 3; out_cols[(row * 6 * colsP + (long long)b * colsP + c) / 2] = __halves2bfloat162(p0[q], p1[q]); }
      if (out_rows) { const int q = ((pat_rows ? pb : pa) >> (2 * b)) & 3; out_rows[(((long long)b * p.rows + row) * colsP + c) / 2] = __halves2bfloat162(p0[q], p1[q]); }
    }
  }
}

static int fill_kr(int32_t dtype, int32_t n, const int64_t* dims, int64_t R, const void* const* factors, const void* w, KrParams* p) {
  TLT_REQUIRE(dtype == TLT_F32 || dtype == TLT_BF16, "khatri_rao: dtype must be f32 or bf16");
  TLT_REQUIRE(n >= 1 && n <= 4 && dims && factors && R >= 1 && R < (1LL << 31), "khatri_rao: 1..4 factors");
  p->n = n; p->R = (int)R; p->w = w; p->rows = 1; p->ldg = R;
  for (int k = 0; k < 4; ++k) { p->dims[k] = 1; p->f[k] = nullptr; }
  for (int k = 0; k < n; ++k) {
    TLT_REQUIRE(dims[k] >= 1 && dims[k] <= 65535 && factors[k], "khatri_rao: bad factor");
    p->dims[k] = (int)dims[k]; p->f[k] = factors[k]; p->rows *= dims[k];
  }
  return TLT_OK;
}

}  // namespace tlt

using namespace tlt;

extern "C" int tlt_khatri_rao_fwd(int32_t dtype, int32_t n, const int64_t* dims, int64_t R, const void* const* factors,
                                  const void* weights, void* out, void* stream) {
  KrParams p;
  int rc = fill_kr(dtype, n, dims, R, factors, weights, &p);
  if (rc) return rc;
  TLT_REQUIRE(out, "tlt_khatri_rao_fwd: null output");
  const long long total = p.rows * p.R;
  const long long blocks = (total + 255) / 256;
  const unsigned grid = (unsigned)(blocks < 148LL * 8 ? blocks : 148LL * 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TLT_BF16) khatri_rao_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p, (__nv_bfloat16*)out);
  else khatri_rao_fwd_kernel<float><<<grid, 256, 0, st>>>(p, (float*)out);
  return check_launch("khatri_rao_fwd_kernel");
}

extern "C" int tlt_khatri_rao_bwd(int32_t dtype, int32_t n, const int64_t* dims, int64_t R, const void* const* factors,
                                  const void* weights, const void* g, float* const* dF, float* dweights, void* stream) {
  return tlt_khatri_rao_bwd_ld(dtype, n, dims, R, factors, weights, g, R, dF, dweights, stream);
}

// same with g given at a leading dimension ldg >= R (the fp32 output of a GEMM whose rows are padded to a 16-byte pitch)
extern "C" int tlt_khatri_rao_bwd_ld(int32_t dtype, int32_t n, const int64_t* dims, int64_t R, const void* const* factors,
                                     const void* weights, const void* g, int64_t ldg, float* const* dF, float* dweights, void* stream) {
  KrParams p;
  int rc = fill_kr(dtype, n, dims, R, factors, weights, &p);
  if (rc) return rc;
  TLT_REQUIRE(g && dF && ldg >= R, "tlt_khatri_rao_bwd: null argument");
  p.ldg = ldg;
  cudaStream_t st = (cudaStream_t)stream;
  for (int k = 0; k < n; ++k) {
    if (dF[k] == nullptr && !(k == 0 && dweights)) continue;
    TLT_REQUIRE(dF[k] != nullptr, "tlt_khatri_rao_bwd: the weight gradient needs dF[0]");
    const long long others = p.rows / p.dims[k];
    TLT_REQUIRE(others < (1LL << 31) && p.rows < (1LL << 31), "tlt_khatri_rao_bwd: product of the factor extents out of range");
    const long long bx = (p.R + 127) / 128;
    long long splits = (148LL * 16 + bx * p.dims[k] - 1) / (bx * p.dims[k]);          // ~16 blocks of 128 threads per SM
    if (splits > others) splits = others;
    if (splits > 65535) splits = 65535;
    if (splits < 1) splits = 1;
    const int per_split = (int)((others + splits - 1) / splits);
    splits = (others + per_split - 1) / per_split;
    TLT_CHECK_CUDA(cudaMemsetAsync(dF[k], 0, (size_t)p.dims[k] * p.R * sizeof(float), st));
    dim3 grid((unsigned)bx, (unsigned)p.dims[k], (unsigned)splits);
    if (dtype == TLT_BF16) khatri_rao_bwd_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(p, k, (const __nv_bfloat16*)g, dF[k], k == 0 ? dweights : nullptr, (int)others, per_split);
    else khatri_rao_bwd_kernel<float><<<grid, 128, 0, st>>>(p, k, (const float*)g, dF[k], k == 0 ? dweights : nullptr, (int)others, per_split);
    rc = check_launch("khatri_rao_bwd_kernel");
    if (rc) return rc;
  }
  return TLT_OK;
}

// fp32 factors -> the (weighted) Khatri-Rao matrix as 3-way bf16 split operands: out_cols (rows, 6 round8(R)) and / or out_rows
// (6 rows, round8(R)), patterns as in tlt_split3_bf16
extern "C" int tlt_khatri_rao_split3(int32_t n, const int64_t* dims, int64_t R, const void* const* factors, const void* weights,
                                     int32_t pattern_cols, void* out_cols, int32_t pattern_rows, void* out_rows, void* stream) {
  KrParams p;
  int rc = fill_kr(TLT_F32, n, dims, R, factors, weights, &p);
  if (rc) return rc;
  TLT_REQUIRE((out_cols || out_rows) && (pattern_cols == 0 || pattern_cols == 1) && (pattern_rows == 0 || pattern_rows == 1) &&
              p.rows < (1LL << 31), "tlt_khatri_rao_split3: bad arguments");
  const int colsP = (int)((R + 7) & ~(int64_t)7);
  const long long total = p.rows * (colsP / 2), blocks = (total + 255) / 256;
  const unsigned grid = (unsigned)(blocks < 148LL * 16 ? blocks : 148LL * 16);
  khatri_rao_split3_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, colsP, pattern_cols, (__nv_bfloat162*)out_cols, pattern_rows,
                                                                  (__nv_bfloat162*)out_rows);
  return check_launch("khatri_rao_split3_kernel");
}

// every factor gradient (and the weight gradient) in one launch; dF[k] may be NULL to skip a factor (dF[0] is needed for dweights)
extern "C" int tlt_khatri_rao_bwd_all(int32_t dtype, int32_t n, const int64_t* dims, int64_t R, const void* const* factors,
                                      const void* weights, const void* g, int64_t ldg, float* const* dF, float* dweights, void* stream) {
  KrParams p;
  int rc = fill_kr(dtype, n, dims, R, factors, weights, &p);
  if (rc) return rc;
  TLT_REQUIRE(g && dF && ldg >= R && p.rows < (1LL << 31), "tlt_khatri_rao_bwd_all: bad arguments");
  TLT_REQUIRE(!dweights || dF[0], "tlt_khatri_rao_bwd_all: the weight gradient needs dF[0]");
  p.ldg = ldg;
  cudaStream_t st = (cudaStream_t)stream;
  KrGrads out;
  long long ysum = 0, min_others = p.rows;
  for (int k = 0; k < 4; ++k) {
    out.dF[k] = k < n ? dF[k] : nullptr;
    if (k < n) {
      ysum += p.dims[k];
      if (p.rows / p.dims[k] < min_others) min_others = p.rows / p.dims[k];
      if (dF[k]) TLT_CHECK_CUDA(cudaMemsetAsync(dF[k], 0, (size_t)p.dims[k] * p.R * sizeof(float), st));
    }
  }
  const long long bxf = (p.R + 127) / 128;
  if (p.dims[n - 1] <= 32) {      // one pass over g (khatri_rao_bwd_fused_kernel)
    const long long prefixes = p.rows / p.dims[n - 1];
    long long fs = 148 / bxf;                                          // about one 128-thread block per SM: runs of prefixes stay
    if (fs < 1) fs = 1;                                               // long, so the atomics per loaded entry stay few
    if (fs > prefixes) fs = prefixes;
    const int per_split = (int)((prefixes + fs - 1) / fs);
    dim3 gridf((unsigned)bxf, (unsigned)((prefixes + per_split - 1) / per_split));
    const bool small = p.dims[n - 1] <= 16;
    if (dtype == TLT_BF16) {
      if (small) khatri_rao_bwd_fused_kernel<__nv_bfloat16, 16><<<gridf, 128, 0, st>>>(p, (const __nv_bfloat16*)g, out, dweights, (int)prefixes, per_split);
      else khatri_rao_bwd_fused_kernel<__nv_bfloat16, 32><<<gridf, 128, 0, st>>>(p, (const __nv_bfloat16*)g, out, dweights, (int)prefixes, per_split);
    } else {
      if (small) khatri_rao_bwd_fused_kernel<float, 16><<<gridf, 128, 0, st>>>(p, (const float*)g, out, dweights, (int)prefixes, per_split);
      else khatri_rao_bwd_fused_kernel<float, 32><<<gridf, 128, 0, st>>>(p, (const float*)g, out, dweights, (int)prefixes, per_split);
    }
    return check_launch("khatri_rao_bwd_fused_kernel");
  }
  TLT_REQUIRE(ysum <= 65535, "tlt_khatri_rao_bwd_all: too many factor rows");
  const long long bx = (p.R + 127) / 128;
  long long splits = (148LL * 16 + bx * ysum - 1) / (bx * ysum);                      // ~16 blocks of 128 threads per SM
  if (splits > min_others) splits = min_others;
  if (splits > 65535) splits = 65535;
  if (splits < 1) splits = 1;
  dim3 grid((unsigned)bx, (unsigned)ysum, (unsigned)splits);
  if (dtype == TLT_BF16) khatri_rao_bwd_all_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(p, (const __nv_bfloat16*)g, out, dweights);
  else khatri_rao_bwd_all_kernel<float><<<grid, 128, 0, st>>>(p, (const float*)g, out, dweights);
  return check_launch("khatri_rao_bwd_all_kernel");
}
