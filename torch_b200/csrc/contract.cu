// tlt_contract: generic strided pairwise contraction  C[b,m,n] = alpha * sum_k A[b,m,k] B[b,k,n] (+ D[b,m,n]).
//
// SIMT (FFMA, fp32 accumulate) kernel used for (a) every fp32 path -- north_star's 1e-5 relative-Frobenius bound rules
// out plain TF32/bf16 tensor-core math for fp32 inputs -- and (b) the skinny / awkwardly-strided stages of the bf16
// paths that the tcgen05 GEMM (gemm_tc.cu) cannot take.  Every index class (b, m, n, k) is a multi-index with
// per-operand strides, so TT sweeps, Tucker mode products, Khatri-Rao builds and NCHW channel mixing all run on the
// tensors where they lie, with no permute / contiguous copies in HBM.
#include "common.cuh"

namespace tlt {

constexpr int kBK = 16;
constexpr int kThreads = 256;

struct ContractParams {
  int nb, nm, nn, nk;
  int size_b[TLT_MAX_SUBDIMS], size_m[TLT_MAX_SUBDIMS], size_n[TLT_MAX_SUBDIMS], size_k[TLT_MAX_SUBDIMS];
  long long a_b[TLT_MAX_SUBDIMS], a_m[TLT_MAX_SUBDIMS], a_k[TLT_MAX_SUBDIMS];
  long long b_b[TLT_MAX_SUBDIMS], b_k[TLT_MAX_SUBDIMS], b_n[TLT_MAX_SUBDIMS];
  long long c_b[TLT_MAX_SUBDIMS], c_m[TLT_MAX_SUBDIMS], c_n[TLT_MAX_SUBDIMS];
  long long d_b[TLT_MAX_SUBDIMS], d_m[TLT_MAX_SUBDIMS], d_n[TLT_MAX_SUBDIMS];
  int M, N, K, Bt;
  int tiles_m, tiles_n, split_k, k_per_split;
  float alpha;
  int accumulate;   // C += (fp32 C)
  int atomic_out;   // partial sums -> atomicAdd
  int dense_c;      // C is the dense fp32 workspace [Bt][M][N]
  int has_d, d_bf16;
  int a_k_fast, b_n_fast;
};

// offset of flattened index `idx` of one class: sub-dim 0 is outermost.
__device__ __forceinline__ long long decomp(int idx, const int* size, const long long* stride, int n) {
  long long off = 0;
#pragma unroll
  for (int d = TLT_MAX_SUBDIMS - 1; d >= 1; --d) {
    if (d < n) {
      int s = size[d];
      int q = idx / s;
      off += (long long)(idx - q * s) * stride[d];
      idx = q;
    }
  }
  if (n >= 1) off += (long long)idx * stride[0];
  return off;
}

template <typename TA, typename TB, typename TC, int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(kThreads) contract_kernel(const ContractParams p, const TA* __restrict__ A,
                                                            const TB* __restrict__ B, const void* __restrict__ D,
                                                            TC* __restrict__ C) {
  static_assert((BM / TM) * (BN / TN) == kThreads, "thread tiling must cover the block tile");
  constexpr int A_PER_T = BM * kBK / kThreads;   // = BM / 16
  constexpr int B_PER_T = BN * kBK / kThreads;   // = BN / 16
  __shared__ __align__(16) float As[kBK][BM + 4];
  __shared__ __align__(16) float Bs[kBK][BN + 4];
  // composite reduction index (nk > 1): offsets of the 16 k's of a tile are decomposed ONCE per tile by 32 threads
  // into these double-buffered tables instead of once per load by every thread
  __shared__ long long koffA[2][kBK], koffB[2][kBK];

  const int tid = threadIdx.x;
  // block id -> (z, tile_m, tile_n);  z -> (batch, split)
  unsigned bid = blockIdx.x;
  const int tn = bid % p.tiles_n; bid /= p.tiles_n;
  const int tm = bid % p.tiles_m; bid /= p.tiles_m;
  const int split = bid % p.split_k;
  const int bidx = bid / p.split_k;
  const int m0 = tm * BM, n0 = tn * BN;
  const int k_begin = split * p.k_per_split;
  const int k_end = min(p.K, k_begin + p.k_per_split);

  A += decomp(bidx, p.size_b, p.a_b, p.nb);
  B += decomp(bidx, p.size_b, p.b_b, p.nb);

  // ---- per-thread load coordinates -------------------------------------------------------------------------
  // A tile (BM x BK): k-fast mapping (kk = tid%16) when the contiguous dim of A is a K dim, else m-fast.
  int a_mm[A_PER_T], a_kk[A_PER_T];
  long long a_moff[A_PER_T];
#pragma unroll
  for (int j = 0; j < A_PER_T; ++j) {
    if (p.a_k_fast) { a_kk[j] = tid % kBK; a_mm[j] = tid / kBK + (kThreads / kBK) * j; }
    else            { a_mm[j] = tid % BM;  a_kk[j] = tid / BM + (kThreads / BM) * j; }
    int m = m0 + a_mm[j];
    a_moff[j] = (m < p.M) ? decomp(m, p.size_m, p.a_m, p.nm) : -1;
  }
  int b_nn[B_PER_T], b_kk[B_PER_T];
  long long b_noff[B_PER_T];
#pragma unroll
  for (int j = 0; j < B_PER_T; ++j) {
    if (p.b_n_fast) { b_nn[j] = tid % BN;  b_kk[j] = tid / BN + (kThreads / BN) * j; }
    else            { b_kk[j] = tid % kBK; b_nn[j] = tid / kBK + (kThreads / kBK) * j; }
    int n = n0 + b_nn[j];
    b_noff[j] = (n < p.N) ? decomp(n, p.size_n, p.b_n, p.nn) : -1;
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  float ra[A_PER_T], rb[B_PER_T];

  const bool ktab = p.nk > 1;
  auto fill_ktab = [&](int k0, int buf) {
    if (tid < kBK) { int k = k0 + tid; if (k < k_end) koffA[buf][tid] = decomp(k, p.size_k, p.a_k, p.nk); }
    else if (tid >= 32 && tid < 32 + kBK) { int k = k0 + tid - 32; if (k < k_end) koffB[buf][tid - 32] = decomp(k, p.size_k, p.b_k, p.nk); }
  };
  auto load_tile = [&](int k0, int buf) {
#pragma unroll
    for (int j = 0; j < A_PER_T; ++j) {
      int k = k0 + a_kk[j];
      float v = 0.f;
      if (a_moff[j] >= 0 && k < k_end) {
        long long ko = ktab ? koffA[buf][a_kk[j]] : (long long)k * p.a_k[0];
        v = to_f32<TA>(A[a_moff[j] + ko]);
      }
      ra[j] = v;
    }
#pragma unroll
    for (int j = 0; j < B_PER_T; ++j) {
      int k = k0 + b_kk[j];
      float v = 0.f;
      if (b_noff[j] >= 0 && k < k_end) {
        long long ko = ktab ? koffB[buf][b_kk[j]] : (long long)k * p.b_k[0];
        v = to_f32<TB>(B[b_noff[j] + ko]);
      }
      rb[j] = v;
    }
  };
  auto store_tile = [&]() {
#pragma unroll
    for (int j = 0; j < A_PER_T; ++j) As[a_kk[j]][a_mm[j]] = ra[j];
#pragma unroll
    for (int j = 0; j < B_PER_T; ++j) Bs[b_kk[j]][b_nn[j]] = rb[j];
  };

  if (k_begin < k_end) {
    if (ktab) {
      fill_ktab(k_begin, 0);
      fill_ktab(k_begin + kBK, 1);
      __syncthreads();
    }
    load_tile(k_begin, 0);
    store_tile();
    __syncthreads();
    int it = 0;
    for (int k0 = k_begin; k0 < k_end; k0 += kBK, ++it) {
      const bool more = (k0 + kBK) < k_end;
      if (more) load_tile(k0 + kBK, (it + 1) & 1);   // global loads in flight while the FMAs below run
      if (ktab && k0 + 2 * kBK < k_end) fill_ktab(k0 + 2 * kBK, it & 1);   // table of tile it+2 (buffer last read for tile it)
#pragma unroll
      for (int kk = 0; kk < kBK; ++kk) {
        float av[TM], bv[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) av[i] = As[kk][ty * TM + i];
#pragma unroll
        for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
      if (more) {
        store_tile();
        __syncthreads();
      }
    }
  }

  // ---- epilogue ----------------------------------------------------------------------------------------------
  long long cb = p.dense_c ? (long long)bidx * p.M * p.N : decomp(bidx, p.size_b, p.c_b, p.nb);
  long long db = (p.has_d && !p.dense_c) ? decomp(bidx, p.size_b, p.d_b, p.nb) : 0;
  long long cn[TN], dn[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    int n = n0 + tx * TN + j;
    cn[j] = dn[j] = 0;
    if (n < p.N) {
      cn[j] = p.dense_c ? (long long)n : decomp(n, p.size_n, p.c_n, p.nn);
      if (p.has_d && !p.dense_c) dn[j] = decomp(n, p.size_n, p.d_n, p.nn);
    }
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= p.M) continue;
    long long cm = p.dense_c ? (long long)m * p.N : decomp(m, p.size_m, p.c_m, p.nm);
    long long dm = (p.has_d && !p.dense_c) ? decomp(m, p.size_m, p.d_m, p.nm) : 0;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n >= p.N) continue;
      float v = p.alpha * acc[i][j];
      if (p.has_d && !p.dense_c && split == 0) {
        long long doff = db + dm + dn[j];
        v += p.d_bf16 ? __bfloat162float(((const __nv_bfloat16*)D)[doff]) : ((const float*)D)[doff];
      }
      TC* dst = C + cb + cm + cn[j];
      if constexpr (sizeof(TC) == 4) {
        if (p.atomic_out) atomicAdd((float*)dst, v);
        else if (p.accumulate) *dst = from_f32<TC>(to_f32<TC>(*dst) + v);
        else *dst = from_f32<TC>(v);
      } else {
        *dst = from_f32<TC>(v);
      }
    }
  }
}

// workspace [Bt][M][N] fp32 -> strided C (+D)
template <typename TC>
__global__ void contract_finalize_kernel(const ContractParams p, const float* __restrict__ ws,
                                         const void* __restrict__ D, TC* __restrict__ C) {
  long long total = (long long)p.Bt * p.M * p.N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int n = (int)(i % p.N);
    long long r = i / p.N;
    int m = (int)(r % p.M);
    int b = (int)(r / p.M);
    float v = ws[i];
    if (p.has_d) {
      long long doff = decomp(b, p.size_b, p.d_b, p.nb) + decomp(m, p.size_m, p.d_m, p.nm) +
                       decomp(n, p.size_n, p.d_n, p.nn);
      v += p.d_bf16 ? __bfloat162float(((const __nv_bfloat16*)D)[doff]) : ((const float*)D)[doff];
    }
    long long coff = decomp(b, p.size_b, p.c_b, p.nb) + decomp(m, p.size_m, p.c_m, p.nm) +
                     decomp(n, p.size_n, p.c_n, p.nn);
    if (p.accumulate) C[coff] = from_f32<TC>(to_f32<TC>(C[coff]) + v);
    else C[coff] = from_f32<TC>(v);
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Vector-matrix special case (M == 1 or N == 1): y[j] = alpha * sum_k v[k] X[k, j].  Bias-gradient reductions
// (v = stride-0 ones), lambda folds and projections onto rank-1 modes land here.  HBM-bound: every element of X is
// read once, coalesced, by the variant matching X's contiguous dimension; K is split across blockIdx.y and partial
// sums meet in the fp32 workspace.
// ---------------------------------------------------------------------------------------------------------------
struct MatvecParams {
  ContractParams p;
  int vec_is_a, J, k_per_split;
};

template <typename TV, typename TX, bool J_FAST>
__global__ void __launch_bounds__(256) matvec_kernel(const MatvecParams q, const TV* __restrict__ vec,
                                                     const TX* __restrict__ mat, float* __restrict__ ws) {
  const ContractParams& p = q.p;
  const int b = blockIdx.z;
  const int* size_j = q.vec_is_a ? p.size_n : p.size_m;
  const long long* mat_j = q.vec_is_a ? p.b_n : p.a_m;
  const int nj = q.vec_is_a ? p.nn : p.nm;
  const long long* vec_k = q.vec_is_a ? p.a_k : p.b_k;
  const long long* mat_k = q.vec_is_a ? p.b_k : p.a_k;
  vec += decomp(b, p.size_b, q.vec_is_a ? p.a_b : p.b_b, p.nb);
  mat += decomp(b, p.size_b, q.vec_is_a ? p.b_b : p.a_b, p.nb);
  const int k_begin = blockIdx.y * q.k_per_split;
  const int k_end = min(p.K, k_begin + q.k_per_split);
  if (J_FAST) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= q.J) return;
    const TX* col = mat + decomp(j, size_j, mat_j, nj);
    float acc = 0.f;
    if (p.nk == 1) {
      const long long sv = vec_k[0], sx = mat_k[0];
      int k = k_begin;
      for (; k + 8 <= k_end; k += 8) {
        float xv[8], vv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { xv[u] = to_f32<TX>(col[(long long)(k + u) * sx]); vv[u] = to_f32<TV>(vec[(long long)(k + u) * sv]); }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc = fmaf(xv[u], vv[u], acc);
      }
      for (; k < k_end; ++k) acc = fmaf(to_f32<TX>(col[(long long)k * sx]), to_f32<TV>(vec[(long long)k * sv]), acc);
    } else {
      for (int k = k_begin; k < k_end; ++k)
        acc = fmaf(to_f32<TX>(col[decomp(k, p.size_k, mat_k, p.nk)]), to_f32<TV>(vec[decomp(k, p.size_k, vec_k, p.nk)]), acc);
    }
    atomicAdd(ws + (long long)b * q.J + j, p.alpha * acc);
  } else {
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= q.J) return;
    const int lane = threadIdx.x & 31;
    const TX* col = mat + decomp(j, size_j, mat_j, nj);
    float acc = 0.f;
    for (int k = k_begin + lane; k < k_end; k += 32) {
      long long ox = p.nk == 1 ? (long long)k * mat_k[0] : decomp(k, p.size_k, mat_k, p.nk);
      long long ov = p.nk == 1 ? (long long)k * vec_k[0] : decomp(k, p.size_k, vec_k, p.nk);
      acc = fmaf(to_f32<TX>(col[ox]), to_f32<TV>(vec[ov]), acc);
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
    if (lane == 0) atomicAdd(ws + (long long)b * q.J + j, p.alpha * acc);
  }
}

// Vectorised j-fast variant: the matrix is unit-stride along a single j dim; each thread owns VEC adjacent outputs and
// reads them with one 16-byte load per k (bias-gradient column sums run at HBM speed through this path).
template <typename TV, typename TX, int VEC>
__global__ void __launch_bounds__(256) matvec_vec_kernel(const MatvecParams q, const TV* __restrict__ vec,
                                                         const TX* __restrict__ mat, float* __restrict__ ws) {
  // block = 32 j-vectors (32*VEC adjacent outputs: one 512-byte row segment per warp load) x 8 k-lanes;
  // k-lanes meet in shared memory, then ONE atomic per output per block (few-way contention only).
  const ContractParams& p = q.p;
  const long long* vec_k = q.vec_is_a ? p.a_k : p.b_k;
  const long long* mat_k = q.vec_is_a ? p.b_k : p.a_k;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int k_begin = blockIdx.y * q.k_per_split;
  const int k_end = min(p.K, k_begin + q.k_per_split);
  const int j = (blockIdx.x * 32 + tx) * VEC;
  const bool active = j < q.J;
  const long long sv = vec_k[0], sx = mat_k[0];
  float acc[VEC];
#pragma unroll
  for (int u = 0; u < VEC; ++u) acc[u] = 0.f;
  if (active) {
    const TX* col = mat + j;
    constexpr int UNROLL = 4;
    int k = k_begin + ty;
    for (; k + 8 * (UNROLL - 1) < k_end; k += 8 * UNROLL) {
      uint4 raw[UNROLL];
      float vv[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        raw[u] = *(const uint4*)(col + (long long)(k + 8 * u) * sx);
        vv[u] = to_f32<TV>(vec[(long long)(k + 8 * u) * sv]);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const TX* e = (const TX*)&raw[u];
#pragma unroll
        for (int w = 0; w < VEC; ++w) acc[w] = fmaf(to_f32<TX>(e[w]), vv[u], acc[w]);
      }
    }
    for (; k < k_end; k += 8) {
      uint4 raw = *(const uint4*)(col + (long long)k * sx);
      float vv = to_f32<TV>(vec[(long long)k * sv]);
      const TX* e = (const TX*)&raw;
#pragma unroll
      for (int w = 0; w < VEC; ++w) acc[w] = fmaf(to_f32<TX>(e[w]), vv, acc[w]);
    }
  }
  __shared__ float red[8][32 * VEC + 1];
#pragma unroll
  for (int w = 0; w < VEC; ++w) red[ty][tx * VEC + w] = acc[w];
  __syncthreads();
  for (int o = threadIdx.x; o < 32 * VEC; o += 256) {
    const int jj = blockIdx.x * 32 * VEC + o;
    if (jj < q.J) {
      float sum = 0.f;
#pragma unroll
      for (int l = 0; l < 8; ++l) sum += red[l][o];
      atomicAdd(ws + jj, p.alpha * sum);
    }
  }
}

template <typename TV, typename TX>
static int launch_matvec_vec(const MatvecParams& q, int splits, const void* vec, const void* mat, float* ws, cudaStream_t st) {
  constexpr int VEC = 16 / sizeof(TX);
  dim3 grid((q.J + 32 * VEC - 1) / (32 * VEC), splits, 1);
  matvec_vec_kernel<TV, TX, VEC><<<grid, 256, 0, st>>>(q, (const TV*)vec, (const TX*)mat, ws);
  return check_launch("matvec_vec_kernel");
}

template <typename TV, typename TX>
static int launch_matvec(const MatvecParams& q, bool j_fast, int splits, const void* vec, const void* mat, float* ws, cudaStream_t st) {
  if (j_fast) {
    dim3 grid((q.J + 255) / 256, splits, q.p.Bt);
    matvec_kernel<TV, TX, true><<<grid, 256, 0, st>>>(q, (const TV*)vec, (const TX*)mat, ws);
  } else {
    dim3 grid((q.J + 7) / 8, splits, q.p.Bt);
    matvec_kernel<TV, TX, false><<<grid, 256, 0, st>>>(q, (const TV*)vec, (const TX*)mat, ws);
  }
  return check_launch("matvec_kernel");
}

// Tiny output, long reduction (M, N <= 8, no batch): factor gradients of the small spatial / kernel-mode matrices, e.g.
// dM[q, i] = sum_{b,p,r} g[b,p,q,r] z[b,p,i,r] -- 4 x 7 outputs over 229 376 terms in tucker_trl's backward.  On the tiled
// kernel a 64 x 64 tile holds 28 useful outputs and the split-K grid is 128 blocks (120-130 us).  Here a thread walks the
// reduction index (consecutive threads on consecutive k), keeps the whole M x N outer-product sum in registers, and the
// block leaves one atomicAdd per output.
constexpr int kSmallOut = 8;
static bool use_small_out(const ContractParams& p) {
  return p.Bt == 1 && p.M > 1 && p.N > 1 && p.M <= kSmallOut && p.N <= kSmallOut && p.K >= 8192;
}

template <typename TA, typename TB>
__global__ void __launch_bounds__(256) small_out_kernel(const ContractParams p, const TA* __restrict__ A, const TB* __restrict__ B,
                                                        float* __restrict__ ws) {
  __shared__ float red[kSmallOut * kSmallOut];
  long long moff[kSmallOut], noff[kSmallOut];
#pragma unroll
  for (int m = 0; m < kSmallOut; ++m) moff[m] = m < p.M ? decomp(m, p.size_m, p.a_m, p.nm) : 0;
#pragma unroll
  for (int n = 0; n < kSmallOut; ++n) noff[n] = n < p.N ? decomp(n, p.size_n, p.b_n, p.nn) : 0;
  float acc[kSmallOut][kSmallOut];
#pragma unroll
  for (int m = 0; m < kSmallOut; ++m)
#pragma unroll
    for (int n = 0; n < kSmallOut; ++n) acc[m][n] = 0.f;
  if (threadIdx.x < kSmallOut * kSmallOut) red[threadIdx.x] = 0.f;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < p.K; k += (long long)gridDim.x * blockDim.x) {
    const long long ao = decomp((int)k, p.size_k, p.a_k, p.nk), bo = decomp((int)k, p.size_k, p.b_k, p.nk);
    float av[kSmallOut], bv[kSmallOut];
#pragma unroll
    for (int m = 0; m < kSmallOut; ++m) av[m] = m < p.M ? to_f32<TA>(A[ao + moff[m]]) : 0.f;
#pragma unroll
    for (int n = 0; n < kSmallOut; ++n) bv[n] = n < p.N ? to_f32<TB>(B[bo + noff[n]]) : 0.f;
#pragma unroll
    for (int m = 0; m < kSmallOut; ++m)
#pragma unroll
      for (int n = 0; n < kSmallOut; ++n) acc[m][n] = fmaf(av[m], bv[n], acc[m][n]);
  }
  __syncthreads();
#pragma unroll
  for (int m = 0; m < kSmallOut; ++m)
#pragma unroll
    for (int n = 0; n < kSmallOut; ++n) {
      if (m < p.M && n < p.N) {
        float v = acc[m][n];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(red + m * kSmallOut + n, v);
      }
    }
  __syncthreads();
  if (threadIdx.x < kSmallOut * kSmallOut) {
    const int m = threadIdx.x / kSmallOut, n = threadIdx.x % kSmallOut;
    if (m < p.M && n < p.N) atomicAdd(ws + m * p.N + n, p.alpha * red[threadIdx.x]);
  }
}

static bool use_matvec(const ContractParams& p) {
  if (p.Bt > 65535) return false;
  long long J = p.M == 1 ? p.N : (p.N == 1 ? p.M : 0);
  return J > 0 && (long long)p.K * J >= 4096;
}

static long long prod(const int64_t* s, int n) {
  long long r = 1;
  for (int i = 0; i < n; ++i) r *= s[i];
  return r;
}

static int choose_split(const tlt_contract_desc* d, long long M, long long N, long long K, long long Bt, int bm, int bn) {
  if (d->split_k > 0) return d->split_k;
  long long tiles = ((M + bm - 1) / bm) * ((N + bn - 1) / bn) * Bt;
  if (tiles >= 2 * 148 || K < 128) return 1;
  long long want = (2 * 148 + tiles - 1) / tiles;       // aim at ~2 blocks per SM
  long long max_by_k = K / 64;                          // but keep >= 4 k-tiles per split
  long long s = want < max_by_k ? want : max_by_k;
  if (s > 128) s = 128;
  return s < 1 ? 1 : (int)s;
}

static void tile_shape(long long M, long long N, int* bm, int* bn) {
  if (N <= 16 && M > 64) { *bm = 128; *bn = 16; }
  else if (M <= 16 && N > 64) { *bm = 16; *bn = 128; }
  else { *bm = 64; *bn = 64; }
}

static bool needs_workspace(const tlt_contract_desc* d, int split) {
  // split-K partials are combined with fp32 atomics in a dense fp32 workspace, then written (cast / +D / +=C) by the
  // finalize kernel, so the caller never has to pre-zero C.
  (void)d;
  return split > 1;
}

template <typename TA, typename TB, typename TC>
static int launch_typed(const ContractParams& p, int bm, int bn, const void* A, const void* B, const void* D, void* C,
                        cudaStream_t st) {
  long long blocks = (long long)p.tiles_m * p.tiles_n * p.split_k * p.Bt;
  if (blocks <= 0 || blocks > 0x7fffffffLL) {
    set_error("tlt_contract: grid of %lld blocks is out of range", blocks);
    return TLT_ERR_INVALID;
  }
  dim3 grid((unsigned)blocks), block(kThreads);
  if (bm == 64) contract_kernel<TA, TB, TC, 64, 64, 4, 4><<<grid, block, 0, st>>>(p, (const TA*)A, (const TB*)B, D, (TC*)C);
  else if (bm == 128) contract_kernel<TA, TB, TC, 128, 16, 4, 2><<<grid, block, 0, st>>>(p, (const TA*)A, (const TB*)B, D, (TC*)C);
  else contract_kernel<TA, TB, TC, 16, 128, 2, 4><<<grid, block, 0, st>>>(p, (const TA*)A, (const TB*)B, D, (TC*)C);
  return check_launch("contract_kernel");
}

template <typename TC>
static int launch_ab(const ContractParams& p, int da, int db, int bm, int bn, const void* A, const void* B,
                     const void* D, void* C, cudaStream_t st) {
  typedef __nv_bfloat16 bf;
  if (da == TLT_F32 && db == TLT_F32) return launch_typed<float, float, TC>(p, bm, bn, A, B, D, C, st);
  if (da == TLT_BF16 && db == TLT_BF16) return launch_typed<bf, bf, TC>(p, bm, bn, A, B, D, C, st);
  if (da == TLT_F32 && db == TLT_BF16) return launch_typed<float, bf, TC>(p, bm, bn, A, B, D, C, st);
  return launch_typed<bf, float, TC>(p, bm, bn, A, B, D, C, st);
}

static int fill_params(const tlt_contract_desc* d, ContractParams* p) {
  TLT_REQUIRE(d != nullptr, "tlt_contract: null descriptor");
  TLT_REQUIRE(d->nb >= 0 && d->nb <= TLT_MAX_SUBDIMS && d->nm >= 0 && d->nm <= TLT_MAX_SUBDIMS && d->nn >= 0 &&
                  d->nn <= TLT_MAX_SUBDIMS && d->nk >= 0 && d->nk <= TLT_MAX_SUBDIMS,
              "tlt_contract: sub-dimension counts must be in [0, %d]", TLT_MAX_SUBDIMS);
  long long M = prod(d->size_m, d->nm), N = prod(d->size_n, d->nn), K = prod(d->size_k, d->nk), Bt = prod(d->size_b, d->nb);
  TLT_REQUIRE(M >= 0 && N >= 0 && K >= 0 && Bt >= 0, "tlt_contract: negative extent");
  TLT_REQUIRE(M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31) && Bt < (1LL << 31),
              "tlt_contract: a flattened index class exceeds 2^31-1");
  memset(p, 0, sizeof(*p));
  p->nb = d->nb; p->nm = d->nm; p->nn = d->nn; p->nk = d->nk;
  for (int i = 0; i < TLT_MAX_SUBDIMS; ++i) {
    p->size_b[i] = i < d->nb ? (int)d->size_b[i] : 1;
    p->size_m[i] = i < d->nm ? (int)d->size_m[i] : 1;
    p->size_n[i] = i < d->nn ? (int)d->size_n[i] : 1;
    p->size_k[i] = i < d->nk ? (int)d->size_k[i] : 1;
    p->a_b[i] = d->a_b[i]; p->a_m[i] = d->a_m[i]; p->a_k[i] = d->a_k[i];
    p->b_b[i] = d->b_b[i]; p->b_k[i] = d->b_k[i]; p->b_n[i] = d->b_n[i];
    p->c_b[i] = d->c_b[i]; p->c_m[i] = d->c_m[i]; p->c_n[i] = d->c_n[i];
    p->d_b[i] = d->d_b[i]; p->d_m[i] = d->d_m[i]; p->d_n[i] = d->d_n[i];
  }
  p->M = (int)M; p->N = (int)N; p->K = (int)K; p->Bt = (int)Bt;
  p->alpha = d->alpha;
  p->accumulate = d->accumulate;
  // choose the coalescing-friendly load mapping: is the unit-stride dim of A (of B) a K dim?
  p->a_k_fast = 0;
  for (int i = 0; i < d->nk; ++i) if (d->a_k[i] == 1 && d->size_k[i] > 1) p->a_k_fast = 1;
  for (int i = 0; i < d->nm; ++i) if (d->a_m[i] == 1 && d->size_m[i] > 1) p->a_k_fast = 0;
  p->b_n_fast = 1;
  for (int i = 0; i < d->nk; ++i) if (d->b_k[i] == 1 && d->size_k[i] > 1) p->b_n_fast = 0;
  for (int i = 0; i < d->nn; ++i) if (d->b_n[i] == 1 && d->size_n[i] > 1) p->b_n_fast = 1;
  return TLT_OK;
}

}  // namespace tlt

using namespace tlt;

extern "C" size_t tlt_contract_workspace_size(const tlt_contract_desc* d) {
  ContractParams p;
  if (fill_params(d, &p) != TLT_OK) return 0;
  if (p.M == 0 || p.N == 0 || p.Bt == 0) return 0;
  if (use_matvec(p) || use_small_out(p)) return (size_t)p.Bt * p.M * p.N * sizeof(float);
  int bm, bn;
  tile_shape(p.M, p.N, &bm, &bn);
  int split = choose_split(d, p.M, p.N, p.K, p.Bt, bm, bn);
  if (!needs_workspace(d, split)) return 0;
  return (size_t)p.Bt * p.M * p.N * sizeof(float);
}

extern "C" int tlt_contract(const tlt_contract_desc* d, const void* A, const void* B, const void* D, void* C,
                            void* workspace, size_t workspace_bytes, void* stream) {
  ContractParams p;
  int rc = fill_params(d, &p);
  if (rc != TLT_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (p.M == 0 || p.N == 0 || p.Bt == 0) return TLT_OK;   // empty output: nothing to do
  TLT_REQUIRE(C && (p.K == 0 || (A && B)), "tlt_contract: null operand pointer");   // empty K: operands are never read
  TLT_REQUIRE(!(d->accumulate && d->dtype_c != TLT_F32), "tlt_contract: accumulate needs an fp32 C");
  int bm, bn;
  tile_shape(p.M, p.N, &bm, &bn);
  p.tiles_m = (p.M + bm - 1) / bm;
  p.tiles_n = (p.N + bn - 1) / bn;
  p.split_k = choose_split(d, p.M, p.N, p.K, p.Bt, bm, bn);
  p.k_per_split = ((p.K + p.split_k - 1) / p.split_k + kBK - 1) / kBK * kBK;
  if (p.k_per_split < kBK) p.k_per_split = kBK;
  p.has_d = D != nullptr;
  p.d_bf16 = d->dtype_d == TLT_BF16;

  if (use_matvec(p)) {
    MatvecParams q;
    q.p = p;
    q.vec_is_a = p.M == 1;
    q.J = q.vec_is_a ? p.N : p.M;
    size_t need = (size_t)p.Bt * q.J * sizeof(float);
    if (!workspace || workspace_bytes < need) {
      set_error("tlt_contract: vector-matrix path needs %zu bytes of workspace, got %zu", need, workspace_bytes);
      return TLT_ERR_WORKSPACE;
    }
    TLT_CHECK_CUDA(cudaMemsetAsync(workspace, 0, need, st));
    // does the matrix operand run contiguously along j (output index) or along k?
    const int64_t* mj = q.vec_is_a ? d->b_n : d->a_m;
    const int64_t* sj = q.vec_is_a ? d->size_n : d->size_m;
    int nj = q.vec_is_a ? d->nn : d->nm;
    bool j_fast = false;
    for (int i = 0; i < nj; ++i) if (mj[i] == 1 && sj[i] > 1) j_fast = true;
    // 16-byte vector path: single unit-stride j dim, single k dim, everything 16-byte aligned
    {
      int dxm = q.vec_is_a ? d->dtype_b : d->dtype_a;
      int vecw = dxm == TLT_BF16 ? 8 : 4;
      const int64_t* mk = q.vec_is_a ? d->b_k : d->a_k;
      const void* matp = q.vec_is_a ? B : A;
      bool ok = nj == 1 && mj[0] == 1 && d->nk == 1 && d->nb == 0 && q.J % vecw == 0 && mk[0] % vecw == 0 &&
                ((uintptr_t)matp % 16) == 0 && q.J >= vecw * 32;
      if (ok) {
        long long blocks_jv = (q.J + 32 * vecw - 1) / (32 * vecw);
        long long want = (148LL * 4 + blocks_jv - 1) / blocks_jv;
        long long max_by_k = p.K / 64;
        long long sp = want < max_by_k ? want : max_by_k;
        if (sp < 1) sp = 1;
        if (sp > 65535) sp = 65535;
        q.k_per_split = (int)((p.K + sp - 1) / sp);
        sp = (p.K + q.k_per_split - 1) / q.k_per_split;
        const void* vecp = q.vec_is_a ? A : B;
        int dvv = q.vec_is_a ? d->dtype_a : d->dtype_b;
        typedef __nv_bfloat16 bf;
        if (dvv == TLT_F32 && dxm == TLT_F32) rc = launch_matvec_vec<float, float>(q, (int)sp, vecp, matp, (float*)workspace, st);
        else if (dvv == TLT_BF16 && dxm == TLT_BF16) rc = launch_matvec_vec<bf, bf>(q, (int)sp, vecp, matp, (float*)workspace, st);
        else if (dvv == TLT_F32 && dxm == TLT_BF16) rc = launch_matvec_vec<float, bf>(q, (int)sp, vecp, matp, (float*)workspace, st);
        else rc = launch_matvec_vec<bf, float>(q, (int)sp, vecp, matp, (float*)workspace, st);
        if (rc != TLT_OK) return rc;
        ContractParams f = p;
        f.alpha = 1.f;
        long long total = (long long)q.J;
        int blocks = (int)((total + 255) / 256 > 148 * 8 ? 148 * 8 : (total + 255) / 256);
        if (d->dtype_c == TLT_F32) contract_finalize_kernel<float><<<blocks, 256, 0, st>>>(f, (const float*)workspace, D, (float*)C);
        else contract_finalize_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(f, (const float*)workspace, D, (__nv_bfloat16*)C);
        return check_launch("contract_finalize_kernel");
      }
    }
    long long blocks_j = j_fast ? (q.J + 255) / 256 : (q.J + 7) / 8;
    long long want = (148LL * 8 + blocks_j * p.Bt - 1) / (blocks_j * p.Bt);
    long long max_by_k = p.K / (j_fast ? 64 : 256);
    long long splits = want < max_by_k ? want : max_by_k;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    q.k_per_split = (int)((p.K + splits - 1) / splits);
    if (q.k_per_split < 1) q.k_per_split = 1;
    splits = (p.K + q.k_per_split - 1) / q.k_per_split;
    if (splits < 1) splits = 1;
    const void* vec = q.vec_is_a ? A : B;
    const void* mat = q.vec_is_a ? B : A;
    int dv = q.vec_is_a ? d->dtype_a : d->dtype_b, dx = q.vec_is_a ? d->dtype_b : d->dtype_a;
    typedef __nv_bfloat16 bf;
    if (dv == TLT_F32 && dx == TLT_F32) rc = launch_matvec<float, float>(q, j_fast, (int)splits, vec, mat, (float*)workspace, st);
    else if (dv == TLT_BF16 && dx == TLT_BF16) rc = launch_matvec<bf, bf>(q, j_fast, (int)splits, vec, mat, (float*)workspace, st);
    else if (dv == TLT_F32 && dx == TLT_BF16) rc = launch_matvec<float, bf>(q, j_fast, (int)splits, vec, mat, (float*)workspace, st);
    else rc = launch_matvec<bf, float>(q, j_fast, (int)splits, vec, mat, (float*)workspace, st);
    if (rc != TLT_OK) return rc;
    ContractParams f = p;
    f.alpha = 1.f;
    long long total = (long long)p.Bt * q.J;
    int blocks = (int)((total + 255) / 256 > 148 * 8 ? 148 * 8 : (total + 255) / 256);
    if (d->dtype_c == TLT_F32) contract_finalize_kernel<float><<<blocks, 256, 0, st>>>(f, (const float*)workspace, D, (float*)C);
    else contract_finalize_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(f, (const float*)workspace, D, (__nv_bfloat16*)C);
    return check_launch("contract_finalize_kernel");
  }

  if (use_small_out(p)) {
    const size_t need = (size_t)p.M * p.N * sizeof(float);
    if (!workspace || workspace_bytes < need) {
      set_error("tlt_contract: small-output path needs %zu bytes of workspace, got %zu", need, workspace_bytes);
      return TLT_ERR_WORKSPACE;
    }
    TLT_CHECK_CUDA(cudaMemsetAsync(workspace, 0, need, st));
    long long blocks = (p.K + 255) / 256 / 4;
    if (blocks < 1) blocks = 1;
    if (blocks > 148LL * 4) blocks = 148LL * 4;
    typedef __nv_bfloat16 bf;
    if (d->dtype_a == TLT_F32 && d->dtype_b == TLT_F32) small_out_kernel<float, float><<<(unsigned)blocks, 256, 0, st>>>(p, (const float*)A, (const float*)B, (float*)workspace);
    else if (d->dtype_a == TLT_BF16 && d->dtype_b == TLT_BF16) small_out_kernel<bf, bf><<<(unsigned)blocks, 256, 0, st>>>(p, (const bf*)A, (const bf*)B, (float*)workspace);
    else if (d->dtype_a == TLT_F32) small_out_kernel<float, bf><<<(unsigned)blocks, 256, 0, st>>>(p, (const float*)A, (const bf*)B, (float*)workspace);
    else small_out_kernel<bf, float><<<(unsigned)blocks, 256, 0, st>>>(p, (const bf*)A, (const float*)B, (float*)workspace);
    rc = check_launch("small_out_kernel");
    if (rc != TLT_OK) return rc;
    ContractParams f = p;
    f.alpha = 1.f;
    if (d->dtype_c == TLT_F32) contract_finalize_kernel<float><<<1, 64, 0, st>>>(f, (const float*)workspace, D, (float*)C);
    else contract_finalize_kernel<__nv_bfloat16><<<1, 64, 0, st>>>(f, (const float*)workspace, D, (__nv_bfloat16*)C);
    return check_launch("contract_finalize_kernel");
  }

  if (needs_workspace(d, p.split_k)) {
    size_t need = (size_t)p.Bt * p.M * p.N * sizeof(float);
    if (!workspace || workspace_bytes < need) {
      set_error("tlt_contract: split-K needs %zu bytes of workspace, got %zu", need, workspace_bytes);
      return TLT_ERR_WORKSPACE;
    }
    TLT_CHECK_CUDA(cudaMemsetAsync(workspace, 0, need, st));
    ContractParams q = p;
    q.dense_c = 1; q.atomic_out = 1; q.accumulate = 0; q.has_d = 0;
    rc = launch_ab<float>(q, d->dtype_a, d->dtype_b, bm, bn, A, B, nullptr, workspace, st);
    if (rc != TLT_OK) return rc;
    long long total = (long long)p.Bt * p.M * p.N;
    int blocks = (int)((total + 255) / 256 > 148 * 8 ? 148 * 8 : (total + 255) / 256);
    if (d->dtype_c == TLT_F32) contract_finalize_kernel<float><<<blocks, 256, 0, st>>>(p, (const float*)workspace, D, (float*)C);
    else contract_finalize_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(p, (const float*)workspace, D, (__nv_bfloat16*)C);
    return check_launch("contract_finalize_kernel");
  }
  p.split_k = 1;
  p.k_per_split = (p.K + kBK - 1) / kBK * kBK;
  if (p.k_per_split < kBK) p.k_per_split = kBK;
  if (d->dtype_c == TLT_F32) return launch_ab<float>(p, d->dtype_a, d->dtype_b, bm, bn, A, B, D, C, st);
  return launch_ab<__nv_bfloat16>(p, d->dtype_a, d->dtype_b, bm, bn, A, B, D, C, st);
}

// =====================================================================================================================
// Pack / unpack: the bridge that lets ANY large bf16 contraction run on the tcgen05 GEMM (gemm_tc.cu).
//
//   tlt_contract_pack   : operand (read through the descriptor's multi-index strides) -> dense K-major bf16 matrix
//                         rows x K with pitch ld (rows = the m class for operand A, the n class for operand B)
//   tlt_contract_unpack : dense M x N result (+ strided addend D) -> the strided output tensor of the descriptor
//
// Mode products whose reduction index is not contiguous, ranks such as 15 / 69 / 388 that break TMA's 16-byte pitch rule,
// unfolded convolutions with K = 9 * 388 ... all ran on the SIMT kernel at 7-15 TFLOP/s.  Two extra streaming passes
// (tile-transposed through shared memory so that both the strided side and the dense side are coalesced) cost a few
// microseconds per 10 MB and buy a 50-100x faster inner product.
// =====================================================================================================================
namespace tlt {

struct PackParams {
  int nr, nk;
  int size_r[TLT_MAX_SUBDIMS], size_k[TLT_MAX_SUBDIMS];
  long long stride_r[TLT_MAX_SUBDIMS], stride_k[TLT_MAX_SUBDIMS];
  long long dstride_r[TLT_MAX_SUBDIMS], dstride_k[TLT_MAX_SUBDIMS];   // unpack: strides of the addend D
  int R, K;
  long long ld;
  int fast_is_k;       // the strided tensor's unit-stride index belongs to the column (k / n) class
  int has_d, d_bf16;
};

constexpr int kPT = 64;   // tile edge

// dst[r][k] = (bf16) src[off_r(r) + off_k(k)]
template <typename TS>
__global__ void __launch_bounds__(256) pack_kernel(const PackParams p, const TS* __restrict__ src, __nv_bfloat16* __restrict__ dst) {
  __shared__ __nv_bfloat16 tile[kPT][kPT + 2];
  __shared__ long long roff[kPT], koff[kPT];
  const int r0 = blockIdx.y * kPT, k0 = blockIdx.x * kPT;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (threadIdx.x < kPT) {
    const int r = r0 + threadIdx.x;
    roff[threadIdx.x] = r < p.R ? decomp(r, p.size_r, p.stride_r, p.nr) : -1;
  } else if (threadIdx.x < 2 * kPT) {
    const int k = k0 + threadIdx.x - kPT;
    koff[threadIdx.x - kPT] = k < p.K ? decomp(k, p.size_k, p.stride_k, p.nk) : -1;
  }
  __syncthreads();
  if (p.fast_is_k) {
    for (int rr = ty; rr < kPT; rr += 8)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int kk = tx + 32 * h;
        const long long ro = roff[rr], ko = koff[kk];
        tile[rr][kk] = (ro >= 0 && ko >= 0) ? __float2bfloat16_rn(to_f32<TS>(src[ro + ko])) : __float2bfloat16_rn(0.f);
      }
  } else {
    for (int kk = ty; kk < kPT; kk += 8)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int rr = tx + 32 * h;
        const long long ro = roff[rr], ko = koff[kk];
        tile[rr][kk] = (ro >= 0 && ko >= 0) ? __float2bfloat16_rn(to_f32<TS>(src[ro + ko])) : __float2bfloat16_rn(0.f);
      }
  }
  __syncthreads();
  // dense side: 4-byte stores, 128 bytes per warp (ld and k0 are even)
  for (int rr = ty; rr < kPT; rr += 8) {
    const int r = r0 + rr, k = k0 + 2 * tx;
    if (r < p.R && k < p.ld) {
      __nv_bfloat162 v;
      v.x = tile[rr][2 * tx];
      v.y = tile[rr][2 * tx + 1];
      *(__nv_bfloat162*)(dst + (long long)r * p.ld + k) = v;
    }
  }
}

// C[off_m(m) + off_n(n)] = Cp[m][n] (+ D[doff_m(m) + doff_n(n)])        Cp of type TP (fp32 for split-K results), C of type TC
template <typename TP, typename TC>
__global__ void __launch_bounds__(256) unpack_kernel(const PackParams p, const TP* __restrict__ Cp, const void* __restrict__ D,
                                                     TC* __restrict__ C) {
  __shared__ float tile[kPT][kPT + 1];
  __shared__ long long roff[kPT], koff[kPT], droff[kPT], dkoff[kPT];
  const int r0 = blockIdx.y * kPT, k0 = blockIdx.x * kPT;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (threadIdx.x < kPT) {
    const int r = r0 + threadIdx.x;
    roff[threadIdx.x] = r < p.R ? decomp(r, p.size_r, p.stride_r, p.nr) : -1;
    droff[threadIdx.x] = (r < p.R && p.has_d) ? decomp(r, p.size_r, p.dstride_r, p.nr) : 0;
  } else if (threadIdx.x < 2 * kPT) {
    const int k = k0 + threadIdx.x - kPT;
    koff[threadIdx.x - kPT] = k < p.K ? decomp(k, p.size_k, p.stride_k, p.nk) : -1;
    dkoff[threadIdx.x - kPT] = (k < p.K && p.has_d) ? decomp(k, p.size_k, p.dstride_k, p.nk) : 0;
  }
  for (int rr = ty; rr < kPT; rr += 8)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kk = tx + 32 * h;
      const int r = r0 + rr, k = k0 + kk;
      tile[rr][kk] = (r < p.R && k < p.K) ? to_f32<TP>(Cp[(long long)r * p.ld + k]) : 0.f;
    }
  __syncthreads();
  auto emit = [&](int rr, int kk) {
    const long long ro = roff[rr], ko = koff[kk];
    if (ro < 0 || ko < 0) return;
    float v = tile[rr][kk];
    if (p.has_d) {
      const long long doff = droff[rr] + dkoff[kk];
      v += p.d_bf16 ? __bfloat162float(((const __nv_bfloat16*)D)[doff]) : ((const float*)D)[doff];
    }
    C[ro + ko] = from_f32<TC>(v);
  };
  if (p.fast_is_k) {
    for (int rr = ty; rr < kPT; rr += 8) { emit(rr, tx); emit(rr, tx + 32); }
  } else {
    for (int kk = ty; kk < kPT; kk += 8) { emit(tx, kk); emit(tx + 32, kk); }
  }
}

static long long min_abs_stride(const int64_t* sizes, const int64_t* strides, int n) {
  long long best = (1LL << 62);
  for (int i = 0; i < n; ++i)
    if (sizes[i] > 1) { long long s = strides[i] < 0 ? -strides[i] : strides[i]; if (s < best) best = s; }
  return best;
}

static int fill_pack(PackParams* p, int nr, const int64_t* size_r, const int64_t* stride_r, int nk, const int64_t* size_k,
                     const int64_t* stride_k, long long ld) {
  TLT_REQUIRE(nr >= 0 && nr <= TLT_MAX_SUBDIMS && nk >= 0 && nk <= TLT_MAX_SUBDIMS, "pack: bad sub-dim counts");
  long long R = 1, K = 1;
  p->nr = nr; p->nk = nk;
  for (int i = 0; i < TLT_MAX_SUBDIMS; ++i) {
    p->size_r[i] = i < nr ? (int)size_r[i] : 1; p->stride_r[i] = i < nr ? stride_r[i] : 0;
    p->size_k[i] = i < nk ? (int)size_k[i] : 1; p->stride_k[i] = i < nk ? stride_k[i] : 0;
    p->dstride_r[i] = 0; p->dstride_k[i] = 0;
    if (i < nr) R *= size_r[i];
    if (i < nk) K *= size_k[i];
  }
  TLT_REQUIRE(R >= 1 && K >= 1 && R < (1LL << 31) && K < (1LL << 31), "pack: extents out of range");
  p->R = (int)R; p->K = (int)K; p->ld = ld;
  p->fast_is_k = min_abs_stride(size_k, stride_k, nk) <= min_abs_stride(size_r, stride_r, nr);
  p->has_d = 0; p->d_bf16 = 0;
  return TLT_OK;
}

}  // namespace tlt

extern "C" int tlt_contract_pack(const tlt_contract_desc* d, int32_t operand, const void* src, void* dst, int64_t dst_ld,
                                 void* stream) {
  TLT_REQUIRE(d && src && dst, "tlt_contract_pack: null argument");
  TLT_REQUIRE(operand == 0 || operand == 1, "tlt_contract_pack: operand must be 0 (A) or 1 (B)");
  TLT_REQUIRE(d->nb == 0, "tlt_contract_pack: batched contractions are not packed");
  TLT_REQUIRE(dst_ld % 2 == 0, "tlt_contract_pack: dst_ld must be even");
  PackParams p;
  int rc = operand == 0 ? fill_pack(&p, d->nm, d->size_m, d->a_m, d->nk, d->size_k, d->a_k, dst_ld)
                        : fill_pack(&p, d->nn, d->size_n, d->b_n, d->nk, d->size_k, d->b_k, dst_ld);
  if (rc) return rc;
  TLT_REQUIRE(dst_ld >= p.K, "tlt_contract_pack: dst_ld smaller than the reduction extent");
  const int dt = operand == 0 ? d->dtype_a : d->dtype_b;
  dim3 grid((unsigned)((dst_ld + kPT - 1) / kPT), (unsigned)((p.R + kPT - 1) / kPT));
  TLT_REQUIRE(grid.y <= 65535, "tlt_contract_pack: too many rows");
  cudaStream_t st = (cudaStream_t)stream;
  if (dt == TLT_BF16) pack_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p, (const __nv_bfloat16*)src, (__nv_bfloat16*)dst);
  else if (dt == TLT_F32) pack_kernel<float><<<grid, 256, 0, st>>>(p, (const float*)src, (__nv_bfloat16*)dst);
  else { set_error("tlt_contract_pack: unknown dtype"); return TLT_ERR_INVALID; }
  return check_launch("pack_kernel");
}

extern "C" int tlt_contract_unpack(const tlt_contract_desc* d, const void* Cp, int32_t dtype_cp, int64_t ldc, const void* D,
                                   void* C, void* stream) {
  TLT_REQUIRE(d && Cp && C, "tlt_contract_unpack: null argument");
  TLT_REQUIRE(d->nb == 0, "tlt_contract_unpack: batched contractions are not packed");
  PackParams p;
  int rc = fill_pack(&p, d->nm, d->size_m, d->c_m, d->nn, d->size_n, d->c_n, ldc);
  if (rc) return rc;
  if (D) {
    p.has_d = 1; p.d_bf16 = d->dtype_d == TLT_BF16;
    for (int i = 0; i < d->nm; ++i) p.dstride_r[i] = d->d_m[i];
    for (int i = 0; i < d->nn; ++i) p.dstride_k[i] = d->d_n[i];
  }
  dim3 grid((unsigned)((p.K + kPT - 1) / kPT), (unsigned)((p.R + kPT - 1) / kPT));
  TLT_REQUIRE(grid.y <= 65535, "tlt_contract_unpack: too many rows");
  cudaStream_t st = (cudaStream_t)stream;
  typedef __nv_bfloat16 bf;
  if (d->dtype_c == TLT_BF16 && dtype_cp == TLT_BF16) unpack_kernel<bf, bf><<<grid, 256, 0, st>>>(p, (const bf*)Cp, D, (bf*)C);
  else if (d->dtype_c == TLT_BF16 && dtype_cp == TLT_F32) unpack_kernel<float, bf><<<grid, 256, 0, st>>>(p, (const float*)Cp, D, (bf*)C);
  else if (d->dtype_c == TLT_F32 && dtype_cp == TLT_F32) unpack_kernel<float, float><<<grid, 256, 0, st>>>(p, (const float*)Cp, D, (float*)C);
  else { set_error("tlt_contract_unpack: unsupported dtype combination"); return TLT_ERR_INVALID; }
  return check_launch("unpack_kernel");
}
