// Fused CP-factorized linear layer (fp32): the whole chain of linear_cp / tensor_dot_cp
//     z[b, r] = sum_i x[b, i] KR_in[i, r]          KR_in[i, r]  = prod_m U_m[i_m, r]     (i = (i_1 .. i_n) row-major)
//     y[b, o] = sum_r z[b, r] lambda_r KR_out[o, r] + bias[o]          KR_out[o, r] = prod_m V_m[o_m, r]
// (tltorch/functional/factorized_linear.py:23-36, factorized_tensordot.py:70-103) in TWO launches forward and FOUR backward; the
// Khatri-Rao matrices are never materialised -- every kernel rebuilds the tile it needs from the (d_m, R) factors in shared memory.
// The unfused path (kr_ops + ops.contract) ran the same step as 20 launches whose fp32 GEMMs (128 x 2341 x 512 on 64 x 64 tiles)
// occupied 16-74 of the 148 SMs.  fp32 FFMA throughout: the layer's 1e-5 tolerance rules out TF32 / bf16 tensor-core products.
//
//   cpl_in_kernel   out[b, r] = sum_k in[b, k] KR[k, r]                       z  (in = x,  U)      G  (in = dy, V)
//   cpl_out_kernel  out[b, n] = sum_r in[b, r] w_r KR[n, r] (+ bias[n])       y  (in = z,  V)      dx (in = G,  U)
//   cpl_grad_kernel dF_m[j, r] = sum_{k: k_m = j} (s_r sum_b A[b, k] W[b, r]) prod_{m' != m} F_m'[k_m', r]
//                                                                             dU (A = x, W = G)    dV (A = dy, W = z)
//                   + dlambda_r = sum_b z[b, r] G[b, r]  and  dbias[k] = sum_b A[b, k]  riding in the dV call
// with G[b, r] = sum_o dy[b, o] KR_out[o, r], dz = lambda G.  A block of cpl_in / cpl_grad owns a tile of 16 rank columns, so the
// factor gradients need no reduction across blocks.
#include "common.cuh"

namespace tlt {
namespace cpl {

struct Modes {
  int n;
  int dim[4];
  const float* f[4];        // (dim[m], R) row-major
};

// digits of a row-major multi-index, last mode fastest
__device__ __forceinline__ void digits(int k, const Modes& md, int (&d)[4]) {
#pragma unroll
  for (int m = 3; m >= 0; --m) {
    if (m < md.n) { d[m] = k % md.dim[m]; k /= md.dim[m]; }
    else d[m] = 0;
  }
}
__device__ __forceinline__ float kr_value(const Modes& md, const int (&d)[4], int r, int R) {
  float v = 1.f;
#pragma unroll
  for (int m = 0; m < 4; ++m)
    if (m < md.n) v *= __ldg(md.f[m] + (size_t)d[m] * R + r);
  return v;
}
__device__ __forceinline__ void cp_async16(void* smem, const void* g, bool ok) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  const int bytes = ok ? 16 : 0;                 // zero-fill when out of range
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(g), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------------------------------------------------------------------------
// out[b, r] = sum_k in[b, k] KR[k, r].   Block: 128 rows x 16 rank columns, 256 threads (4 rows x 2 columns each), k chunks of 64
// double-buffered with cp.async.  in: (B, K) with leading dimension ldi (a multiple of 4 floats); out: (B, ldo).
// ---------------------------------------------------------------------------------------------------------------------------------
constexpr int IN_BM = 128, IN_BN = 16, IN_BK = 64, IN_PITCH = IN_BK + 4;

__global__ void __launch_bounds__(256) cpl_in_kernel(const float* __restrict__ in, int B, int K, int ldi, int R, Modes md,
                                                     float* __restrict__ out, int ldo) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;                                   // [2][128][IN_PITCH]
  float* ks = sm + 2 * IN_BM * IN_PITCH;            // [2][64][16]
  float* fs = ks + 2 * IN_BK * IN_BN;               // [sum dims][16]: this rank tile's columns of every factor (read once)
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * IN_BN, b0 = blockIdx.y * IN_BM;
  const int tb = tid >> 3, tr = tid & 7;            // rows tb, tb + 32, tb + 64, tb + 96; columns 2 tr, 2 tr + 1
  const int nchunk = (K + IN_BK - 1) / IN_BK;
  int sumd = 0, off[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) { off[m] = sumd; if (m < md.n) sumd += md.dim[m]; }
  for (int i = tid; i < sumd * IN_BN; i += 256) {
    const int row = i / IN_BN, rr = i - row * IN_BN;
    int m = 0;
#pragma unroll
    for (int q = 1; q < 4; ++q) if (q < md.n && row >= off[q]) m = q;
    fs[i] = r0 + rr < R ? __ldg(md.f[m] + (size_t)(row - off[m]) * R + r0 + rr) : 0.f;
  }
  __syncthreads();

  auto load = [&](int c, int buf) {
    const int k0 = c * IN_BK;
    float* xd = xs + buf * IN_BM * IN_PITCH;
#pragma unroll
    for (int i = 0; i < 8; ++i) {                   // 128 rows x 16 chunks of 16 bytes
      const int q = tid + 256 * i, row = q >> 4, ch = q & 15;
      const bool ok = b0 + row < B && k0 + ch * 4 < K;
      cp_async16(xd + row * IN_PITCH + ch * 4, in + (size_t)(ok ? b0 + row : 0) * ldi + (ok ? k0 + ch * 4 : 0), ok);
    }
    cp_async_commit();
    float* kd = ks + buf * IN_BK * IN_BN;
#pragma unroll
    for (int i = 0; i < 4; ++i) {                   // 64 x 16 Khatri-Rao entries
      const int q = tid + 256 * i, kk = q >> 4, rr = q & 15;
      float v = 0.f;
      if (k0 + kk < K) {
        int d[4];
        digits(k0 + kk, md, d);
        v = 1.f;
#pragma unroll
        for (int m = 0; m < 4; ++m) if (m < md.n) v *= fs[(off[m] + d[m]) * IN_BN + rr];
      }
      kd[kk * IN_BN + rr] = v;
    }
  };

  float acc[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.f;
  load(0, 0);
  for (int c = 0; c < nchunk; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunk) { load(c + 1, buf ^ 1); cp_async_wait<1>(); } else cp_async_wait<0>();
    __syncthreads();
    const float* xb = xs + buf * IN_BM * IN_PITCH + tb * IN_PITCH;
    const float* kb = ks + buf * IN_BK * IN_BN + 2 * tr;
#pragma unroll 4
    for (int k = 0; k < IN_BK; k += 4) {
      float4 xv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i] = *(const float4*)(xb + i * 32 * IN_PITCH + k);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 kv = *(const float2*)(kb + (k + q) * IN_BN);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float x = q == 0 ? xv[i].x : q == 1 ? xv[i].y : q == 2 ? xv[i].z : xv[i].w;
          acc[i][0] = fmaf(x, kv.x, acc[i][0]);
          acc[i][1] = fmaf(x, kv.y, acc[i][1]);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = b0 + tb + 32 * i;
    if (b < B) {
      // columns [R, ldo) are the row's 16-byte padding: their Khatri-Rao entries are zero, so the sums written there are zeros
      if (r0 + 2 * tr < ldo) out[(size_t)b * ldo + r0 + 2 * tr] = acc[i][0];
      if (r0 + 2 * tr + 1 < ldo) out[(size_t)b * ldo + r0 + 2 * tr + 1] = acc[i][1];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// out[b, n] = sum_r in[b, r] w_r KR[n, r] (+ bias[n]).   Block: 32 rows x 16 columns, 128 threads (4 rows x 1 column each), rank
// chunks of 32 double-buffered.  in: (B, ldi) with ldi a multiple of 4 and columns [R, ldi) zero or ignored (KR is zero there).
// ---------------------------------------------------------------------------------------------------------------------------------
constexpr int OUT_BM = 32, OUT_BN = 16, OUT_BK = 32, OUT_PITCH = OUT_BK + 4;

__device__ __forceinline__ void cp_async4(void* smem, const void* g, bool ok) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  const int bytes = ok ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(s), "l"(g), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(128) cpl_out_kernel(const float* __restrict__ in, int B, int R, int ldi, int N, Modes md,
                                                      const float* __restrict__ w, const float* __restrict__ bias,
                                                      float* __restrict__ out, int ldo) {
  __shared__ __align__(16) float xs[2][OUT_BM][OUT_PITCH];
  // rank chunk of the factor rows this tile's 16 columns use: [column][mode][32 ranks] (+ the weights' chunk); the factors' rows
  // are R floats apart (any R), so they arrive as 4-byte asynchronous copies riding in the same group as the activation chunk
  __shared__ __align__(16) float fr[2][4][OUT_BN][OUT_PITCH];     // column stride 36 floats: 16 columns x float4 conflict-free
  __shared__ __align__(16) float wsm[2][OUT_BK];
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * OUT_BN, b0 = blockIdx.y * OUT_BM;
  const int tb = tid >> 4, tn = tid & 15;           // rows 4 tb .. 4 tb + 3, column tn
  const int nchunk = (R + OUT_BK - 1) / OUT_BK;
  // staging slots of this thread: (column sc, mode sm) pairs tid >> 5 + 4 j, rank lane tid & 31
  const int lane = tid & 31;
  const float* srow[16];                            // 64 (column, slot) rows / 4 warps = 16 rows per thread
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int q = (tid >> 5) + 4 * j, sc = q >> 2, smode = q & 3;
    int dn[4];
    digits(min(n0 + sc, N - 1), md, dn);
    srow[j] = smode < md.n ? md.f[smode] + (size_t)dn[smode] * R : nullptr;
  }

  auto load = [&](int c, int buf) {
    const int rr0 = c * OUT_BK;
#pragma unroll
    for (int i = 0; i < 2; ++i) {                   // 32 rows x 8 chunks of 16 bytes
      const int q = tid + 128 * i, row = q >> 3, ch = q & 7;
      const bool ok = b0 + row < B && rr0 + ch * 4 < ldi;
      cp_async16(&xs[buf][row][ch * 4], in + (size_t)(ok ? b0 + row : 0) * ldi + (ok ? rr0 + ch * 4 : 0), ok);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int q = (tid >> 5) + 4 * j, sc = q >> 2, smode = q & 3;
      if (srow[j]) cp_async4(&fr[buf][smode][sc][lane], srow[j] + (rr0 + lane < R ? rr0 + lane : 0), rr0 + lane < R);
    }
    if (w && tid < OUT_BK) cp_async4(&wsm[buf][tid], w + (rr0 + tid < R ? rr0 + tid : 0), rr0 + tid < R);
    cp_async_commit();
  };
  // slots nobody copies into (missing modes, no weights) hold ones
  for (int i = tid; i < 2 * 4 * OUT_BN * OUT_PITCH; i += 128) {
    const int smode = (i / (OUT_BN * OUT_PITCH)) & 3;
    if (smode >= md.n) (&fr[0][0][0][0])[i] = 1.f;
  }
  if (!w && tid < 2 * OUT_BK) (&wsm[0][0])[tid] = 1.f;
  __syncthreads();

  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  load(0, 0);
  for (int c = 0; c < nchunk; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunk) { load(c + 1, buf ^ 1); cp_async_wait<1>(); } else cp_async_wait<0>();
    __syncthreads();
#pragma unroll
    for (int k = 0; k < OUT_BK; k += 4) {
      const float4 f0 = *(const float4*)&fr[buf][0][tn][k], f1 = *(const float4*)&fr[buf][1][tn][k];
      const float4 f2 = *(const float4*)&fr[buf][2][tn][k], f3 = *(const float4*)&fr[buf][3][tn][k];
      const float4 wv = *(const float4*)&wsm[buf][k];
      float4 kv;
      kv.x = (f0.x * f1.x) * (f2.x * f3.x) * wv.x; kv.y = (f0.y * f1.y) * (f2.y * f3.y) * wv.y;
      kv.z = (f0.z * f1.z) * (f2.z * f3.z) * wv.z; kv.w = (f0.w * f1.w) * (f2.w * f3.w) * wv.w;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 xv = *(const float4*)&xs[buf][4 * tb + i][k];
        acc[i] = fmaf(xv.x, kv.x, acc[i]); acc[i] = fmaf(xv.y, kv.y, acc[i]);
        acc[i] = fmaf(xv.z, kv.z, acc[i]); acc[i] = fmaf(xv.w, kv.w, acc[i]);
      }
    }
    __syncthreads();
  }
  if (n0 + tn < N) {
    const float bv = bias ? __ldg(bias + n0 + tn) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = b0 + 4 * tb + i;
      if (b < B) out[(size_t)b * ldo + n0 + tn] = acc[i] + bv;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Factor gradients of one side.  Block: 16 rank columns, all k, all b; 256 threads, T[64 k][16 r] per k chunk (2 k x 2 r per thread).
//   T[k, r] = s_r sum_b A[b, k] W[b, r];   dF_m[j, r] = sum_{k: k_m = j} T[k, r] prod_{m' != m} F_m'[k_m', r]
// dF_m: (dim[m], R) fp32, written completely.  Optional riders: dlam[r] = sum_b W[b, r] W2[b, r];  dcol[k] = sum_b A[b, k] (block 0).
// ---------------------------------------------------------------------------------------------------------------------------------
constexpr int G_BK = 64, G_BN = 16, G_BB = 128, G_PITCH = G_BK + 4, G_TP = G_BN + 1;

__global__ void __launch_bounds__(256) cpl_grad_kernel(const float* __restrict__ A, int B, int K, int lda, const float* __restrict__ W, int ldw,
                                                       int R, Modes md, const float* __restrict__ s, float* dF0, float* dF1, float* dF2,
                                                       float* dF3, const float* __restrict__ W2, float* __restrict__ dlam,
                                                       float* __restrict__ dcol) {
  extern __shared__ __align__(16) float sm[];
  float* as = sm;                                   // [2][128][G_PITCH]
  float* ws = as + 2 * G_BB * G_PITCH;              // [2][128][16]
  float* ts = ws + 2 * G_BB * G_BN;                 // [64][G_TP]   T of the finished k chunk, scaled
  float* fs = ts + G_BK * G_TP;                     // [sum dims][16]  factor values of this rank tile
  int sumd = 0, off[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) { off[m] = sumd; if (m < md.n) sumd += md.dim[m]; }
  float* gs = fs + sumd * G_BN;                     // [sum dims][16]  gradient accumulators
  float* rs = gs + sumd * G_BN;                     // [16] dlam sums + [64] column sums of the current k chunk
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * G_BN;

  if (tid < 80) rs[tid] = 0.f;
  for (int i = tid; i < sumd * G_BN; i += 256) {
    const int row = i / G_BN, rr = i - row * G_BN;
    int m = 0;
#pragma unroll
    for (int q = 1; q < 4; ++q) if (q < md.n && row >= off[q]) m = q;
    fs[i] = r0 + rr < R ? __ldg(md.f[m] + (size_t)(row - off[m]) * R + r0 + rr) : 0.f;
    gs[i] = 0.f;
  }
  const int tk = tid >> 3, tr = tid & 7;            // k = 2 tk, 2 tk + 1 of the chunk;  r = 2 tr, 2 tr + 1
  const float s0 = (s && r0 + 2 * tr < R) ? __ldg(s + r0 + 2 * tr) : 1.f;
  const float s1 = (s && r0 + 2 * tr + 1 < R) ? __ldg(s + r0 + 2 * tr + 1) : 1.f;
  const int nk = (K + G_BK - 1) / G_BK, nb = (B + G_BB - 1) / G_BB, total = nk * nb;
  const bool do_col = dcol && blockIdx.x == 0;

  auto issue = [&](int it, int buf) {
    const int kc = it / nb, bc = it - kc * nb, k0 = kc * G_BK, b0 = bc * G_BB;
    float* ad = as + buf * G_BB * G_PITCH;
#pragma unroll
    for (int i = 0; i < 8; ++i) {                   // A chunk: 128 rows x 16 chunks of 16 bytes
      const int q = tid + 256 * i, row = q >> 4, ch = q & 15;
      const bool ok = b0 + row < B && k0 + ch * 4 < K;
      cp_async16(ad + row * G_PITCH + ch * 4, A + (size_t)(ok ? b0 + row : 0) * lda + (ok ? k0 + ch * 4 : 0), ok);
    }
    if (nb > 1 || it == 0) {                        // W tile of this batch block (one batch block: loaded once, slot 0)
      float* wd = ws + (nb > 1 ? buf : 0) * G_BB * G_BN;
#pragma unroll
      for (int i = 0; i < 2; ++i) {                 // 128 rows x 4 chunks (ldw is a multiple of 4, pad columns are zero)
        const int q = tid + 256 * i, row = q >> 2, ch = q & 3;
        const bool ok = b0 + row < B && r0 + ch * 4 < ldw;
        cp_async16(wd + row * G_BN + ch * 4, W + (size_t)(ok ? b0 + row : 0) * ldw + (ok ? r0 + ch * 4 : 0), ok);
      }
    }
    cp_async_commit();
  };

  float t[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  issue(0, 0);
  for (int it = 0; it < total; ++it) {
    const int buf = it & 1, kc = it / nb, bc = it - kc * nb, k0 = kc * G_BK, b0 = bc * G_BB;
    if (it + 1 < total) { issue(it + 1, buf ^ 1); cp_async_wait<1>(); } else cp_async_wait<0>();
    __syncthreads();
    const float* ab = as + buf * G_BB * G_PITCH;
    const float* wb = ws + (nb > 1 ? buf : 0) * G_BB * G_BN;
    const int rows = min(G_BB, B - b0);
    if (dlam && kc == 0) {                          // riders: every thread sums a strip of rows, shared-memory reduction
      const int rr = tid & 15;
      float p = 0.f;
      if (r0 + rr < R)
        for (int b = tid >> 4; b < rows; b += 16) p += wb[b * G_BN + rr] * __ldg(W2 + (size_t)(b0 + b) * ldw + r0 + rr);
      atomicAdd(rs + rr, p);
    }
    if (do_col) {
      const int cc = tid & 63;
      float p = 0.f;
      for (int b = tid >> 6; b < rows; b += 4) p += ab[b * G_PITCH + cc];
      atomicAdd(rs + 16 + cc, p);
    }
#pragma unroll 8
    for (int b = 0; b < G_BB; ++b) {                // rows past the batch are zero-filled
      const float2 a = *(const float2*)(ab + b * G_PITCH + 2 * tk);
      const float2 w = *(const float2*)(wb + b * G_BN + 2 * tr);
      t[0][0] = fmaf(a.x, w.x, t[0][0]); t[0][1] = fmaf(a.x, w.y, t[0][1]);
      t[1][0] = fmaf(a.y, w.x, t[1][0]); t[1][1] = fmaf(a.y, w.y, t[1][1]);
    }
    if (bc == nb - 1) {
      // T of this k chunk is complete: through shared memory so that the lanes of a warp span the RANK columns of the fold
      // (distinct accumulator addresses) instead of the k rows (one address for every mode that is constant over the chunk)
      ts[(2 * tk) * G_TP + 2 * tr] = t[0][0] * s0;     ts[(2 * tk) * G_TP + 2 * tr + 1] = t[0][1] * s1;
      ts[(2 * tk + 1) * G_TP + 2 * tr] = t[1][0] * s0; ts[(2 * tk + 1) * G_TP + 2 * tr + 1] = t[1][1] * s1;
      t[0][0] = t[0][1] = t[1][0] = t[1][1] = 0.f;
      __syncthreads();
      if (do_col && tid < 64) {
        if (k0 + tid < K) dcol[k0 + tid] = rs[16 + tid];
        rs[16 + tid] = 0.f;                         // (next chunk's sums start after the loop's closing barrier)
      }
      const int rr = tid & 15, kq = tid >> 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int kl = kq + 16 * q, k = k0 + kl;
        if (k < K) {
          int d[4];
          digits(k, md, d);
          const float tv = ts[kl * G_TP + rr];
          float fv[4];
#pragma unroll
          for (int m = 0; m < 4; ++m) fv[m] = m < md.n ? fs[(off[m] + d[m]) * G_BN + rr] : 1.f;
#pragma unroll
          for (int m = 0; m < 4; ++m)
            if (m < md.n) {
              float p = tv;
#pragma unroll
              for (int o = 0; o < 4; ++o) if (o != m && o < md.n) p *= fv[o];
              atomicAdd(gs + (off[m] + d[m]) * G_BN + rr, p);
            }
        }
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < sumd * G_BN; i += 256) {
    const int row = i / G_BN, rr = i - row * G_BN;
    int m = 0;
#pragma unroll
    for (int q = 1; q < 4; ++q) if (q < md.n && row >= off[q]) m = q;
    float* dst = m == 0 ? dF0 : m == 1 ? dF1 : m == 2 ? dF2 : dF3;
    if (r0 + rr < R) dst[(size_t)(row - off[m]) * R + r0 + rr] = gs[i];
  }
  if (dlam && tid < G_BN && r0 + tid < R) dlam[r0 + tid] = rs[tid];
}

static int fill_modes(int n, const int64_t* dims, const void* const* ptrs, Modes* md, long long* prod) {
  TLT_REQUIRE(n >= 1 && n <= 4, "tlt_cplinear: 1 to 4 tensorized modes per side");
  md->n = n;
  *prod = 1;
  for (int m = 0; m < 4; ++m) {
    md->dim[m] = m < n ? (int)dims[m] : 1;
    md->f[m] = m < n ? (const float*)ptrs[m] : nullptr;
    if (m < n) {
      TLT_REQUIRE(dims[m] >= 1 && ptrs[m], "tlt_cplinear: bad factor");
      *prod *= dims[m];
    }
  }
  TLT_REQUIRE(*prod < (1LL << 30), "tlt_cplinear: side too large");
  return TLT_OK;
}

}  // namespace cpl
}  // namespace tlt

using namespace tlt;
using namespace tlt::cpl;

static inline int64_t r4(int64_t v) { return (v + 3) & ~(int64_t)3; }

extern "C" size_t tlt_cplinear_workspace_size(const tlt_cplinear_desc* d) {
  if (!d || d->batch < 0 || d->rank < 1) return 0;
  return (size_t)d->batch * (size_t)r4(d->rank) * sizeof(float);      // one (B, round4(R)) fp32 matrix: z (forward) / G (backward)
}

static int check_desc(const tlt_cplinear_desc* d, Modes* in, Modes* out, const void* const* f_in, const void* const* f_out, long long* I,
                      long long* O) {
  TLT_REQUIRE(d && d->batch >= 1 && d->rank >= 1 && d->batch < (1LL << 30) && d->rank < (1LL << 30), "tlt_cplinear: bad descriptor");
  int rc = fill_modes(d->n_in, d->in_dims, f_in, in, I);
  if (rc) return rc;
  rc = fill_modes(d->n_out, d->out_dims, f_out, out, O);
  if (rc) return rc;
  long long sd = 0;
  for (int m = 0; m < d->n_in; ++m) sd += d->in_dims[m];
  long long so = 0;
  for (int m = 0; m < d->n_out; ++m) so += d->out_dims[m];
  TLT_REQUIRE((sd > so ? sd : so) <= 512, "tlt_cplinear: sum of a side's mode sizes above 512");
  TLT_REQUIRE(*I % 4 == 0, "tlt_cplinear: input features must be a multiple of 4 (16-byte rows)");
  return TLT_OK;
}

static size_t in_smem(const Modes& md) {
  int sd = 0;
  for (int m = 0; m < md.n; ++m) sd += md.dim[m];
  return (size_t)(2 * IN_BM * IN_PITCH + 2 * IN_BK * IN_BN + sd * IN_BN) * sizeof(float);
}
static size_t grad_smem(const Modes& md) {
  int sd = 0;
  for (int m = 0; m < md.n; ++m) sd += md.dim[m];
  return (size_t)(2 * G_BB * G_PITCH + 2 * G_BB * G_BN + G_BK * G_TP + 2 * sd * G_BN + 80) * sizeof(float);
}

// y (B, O) <- x (B, I); z (B, round4(R)) is written for the backward call
extern "C" int tlt_cplinear_fwd(const tlt_cplinear_desc* d, const float* x, const void* const* f_in, const void* const* f_out,
                                const float* lambda, const float* bias, float* z, float* y, void* stream) {
  Modes in, out;
  long long I, O;
  int rc = check_desc(d, &in, &out, f_in, f_out, &I, &O);
  if (rc) return rc;
  TLT_REQUIRE(x && lambda && z && y, "tlt_cplinear_fwd: null pointer");
  TLT_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)z % 16) == 0, "tlt_cplinear_fwd: 16-byte aligned x and z");
  cudaStream_t st = (cudaStream_t)stream;
  const int B = (int)d->batch, R = (int)d->rank, ldz = (int)r4(R);
  const size_t smem_in = in_smem(in);
  TLT_ENSURE_SMEM(cpl_in_kernel, smem_in);
  cpl_in_kernel<<<dim3((R + IN_BN - 1) / IN_BN, (B + IN_BM - 1) / IN_BM), 256, smem_in, st>>>(x, B, (int)I, (int)I, R, in, z, ldz);
  rc = check_launch("cpl_in_kernel");
  if (rc) return rc;
  cpl_out_kernel<<<dim3((unsigned)((O + OUT_BN - 1) / OUT_BN), (B + OUT_BM - 1) / OUT_BM), 128, 0, st>>>(z, B, R, ldz, (int)O, out, lambda,
                                                                                                        bias, y, (int)O);
  return check_launch("cpl_out_kernel");
}

// dx (B, I), dF_in[m] (d_m, R), dF_out[m] (e_m, R), dlambda (R), dbias (O) or NULL  <- x, dy, z (from the forward call); g: workspace
extern "C" int tlt_cplinear_bwd(const tlt_cplinear_desc* d, const float* x, const float* dy, const float* z, const void* const* f_in,
                                const void* const* f_out, const float* lambda, float* g, float* dx, void* const* df_in, void* const* df_out,
                                float* dlambda, float* dbias, void* stream) {
  Modes in, out;
  long long I, O;
  int rc = check_desc(d, &in, &out, f_in, f_out, &I, &O);
  if (rc) return rc;
  TLT_REQUIRE(x && dy && z && lambda && g && dx && df_in && df_out && dlambda, "tlt_cplinear_bwd: null pointer");
  TLT_REQUIRE(O % 4 == 0, "tlt_cplinear_bwd: output features must be a multiple of 4 (16-byte rows)");
  TLT_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)dy % 16) == 0 && ((uintptr_t)z % 16) == 0 && ((uintptr_t)g % 16) == 0,
              "tlt_cplinear_bwd: 16-byte aligned activations");
  cudaStream_t st = (cudaStream_t)stream;
  const int B = (int)d->batch, R = (int)d->rank, ldz = (int)r4(R);
  const size_t smem_in = in_smem(out);
  TLT_ENSURE_SMEM(cpl_in_kernel, smem_in);
  const size_t sg_in = grad_smem(in), sg_out = grad_smem(out);
  TLT_ENSURE_SMEM(cpl_grad_kernel, sg_in > sg_out ? sg_in : sg_out);
  const dim3 grid_r((R + IN_BN - 1) / IN_BN, (B + IN_BM - 1) / IN_BM);
  // G[b, r] = sum_o dy[b, o] KR_out[o, r]
  cpl_in_kernel<<<grid_r, 256, smem_in, st>>>(dy, B, (int)O, (int)O, R, out, g, ldz);
  rc = check_launch("cpl_in_kernel (G)");
  if (rc) return rc;
  // dx[b, i] = sum_r G[b, r] lambda_r KR_in[i, r]
  cpl_out_kernel<<<dim3((unsigned)((I + OUT_BN - 1) / OUT_BN), (B + OUT_BM - 1) / OUT_BM), 128, 0, st>>>(g, B, R, ldz, (int)I, in, lambda, nullptr,
                                                                                                        dx, (int)I);
  rc = check_launch("cpl_out_kernel (dx)");
  if (rc) return rc;
  const unsigned gr = (unsigned)((R + G_BN - 1) / G_BN);
  // dU_m from T = lambda (x^T G)
  cpl_grad_kernel<<<gr, 256, sg_in, st>>>(x, B, (int)I, (int)I, g, ldz, R, in, lambda, (float*)df_in[0], in.n > 1 ? (float*)df_in[1] : nullptr,
                                          in.n > 2 ? (float*)df_in[2] : nullptr, in.n > 3 ? (float*)df_in[3] : nullptr, nullptr, nullptr,
                                          nullptr);
  rc = check_launch("cpl_grad_kernel (in)");
  if (rc) return rc;
  // dV_m from T = lambda (dy^T z);  dlambda = sum_b z G;  dbias = column sums of dy
  cpl_grad_kernel<<<gr, 256, sg_out, st>>>(dy, B, (int)O, (int)O, z, ldz, R, out, lambda, (float*)df_out[0],
                                           out.n > 1 ? (float*)df_out[1] : nullptr, out.n > 2 ? (float*)df_out[2] : nullptr,
                                           out.n > 3 ? (float*)df_out[3] : nullptr, g, dlambda, dbias);
  return check_launch("cpl_grad_kernel (out)");
}
