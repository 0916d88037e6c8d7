// Op-level entry points: a whole factorized layer per call, for hosts that are not Python (tlt_<op>_{fwd,bwd} +
// *_workspace_size).  Each one only sequences the kernels the Python ops of torch_b200 sequence themselves (blocktt_ops.py,
// trl_ops.py) -- same launches, same order, same results -- on the caller's stream, with every intermediate in a workspace the
// caller allocates and keeps between the forward and the backward call.
//
//   tlt_blocktt3_linear_*   FactorizedLinear with a 3-core BlockTT weight, bf16: rebuild W on chip-resident cores, one tcgen05 GEMM
//                           (+ bias); backward: dX GEMM, fp32 split-K dW GEMM, the two core-gradient launches, bias column sums
//                           (linear_blocktt, tltorch/functional/factorized_linear.py:39-60, in its dense plan)
//   tlt_tucker_trl_*        Tucker tensor-regression layer on (B, C, s_1..s_n) maps, bf16, one output mode: Kronecker product of the
//                           spatial factors, spatial projection (one pass over the activation), channel GEMM, fused tail;
//                           backward with the parameter gradients on an optional second stream
//                           (tucker_trl, tltorch/functional/tensor_regression.py:37-49)
#include "common.cuh"

namespace tlt {
namespace opl {

static inline size_t al(size_t v) { return (v + 255) & ~(size_t)255; }
static inline int64_t r8(int64_t v) { return (v + 7) & ~(int64_t)7; }

// column sums of a (rows, cols) bf16 matrix -> bf16 / fp32 vector (bias gradient); block = 64 columns x 4 row lanes
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ a, long long rows, int cols, int out_f32, void* __restrict__ out) {
  __shared__ float red[4][64];
  const int cl = threadIdx.x & 63, lane = threadIdx.x >> 6;
  const int c = blockIdx.x * 64 + cl;
  float s = 0.f;
  if (c < cols)
    for (long long r = lane; r < rows; r += 4) s += __bfloat162float(a[r * cols + c]);
  red[lane][cl] = s;
  __syncthreads();
  if (lane == 0 && c < cols) {
    s = red[0][cl] + red[1][cl] + red[2][cl] + red[3][cl];
    if (out_f32) ((float*)out)[c] = s; else ((__nv_bfloat16*)out)[c] = __float2bfloat16_rn(s);
  }
}

// dst (cols, rows) bf16 <- src (rows, ld) fp32, transposed (the channel-factor gradient of the TRL: (rc, C) fp32 -> (C, rc))
__global__ void transpose_cast_kernel(const float* __restrict__ src, int rows, int cols, int ld, __nv_bfloat16* __restrict__ dst) {
  const long long n = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i / rows), r = (int)(i - (long long)c * rows);
    dst[i] = __float2bfloat16_rn(src[(size_t)r * ld + c]);
  }
}

struct BttGeom { long long B, R, Cn, fin, fout; };
static int btt_geom(const tlt_blocktt3_linear_desc* d, BttGeom* g) {
  TLT_REQUIRE(d && d->tt.dtype == TLT_BF16 && d->batch >= 1 && d->batch < (1LL << 31), "tlt_blocktt3_linear: bf16 cores, batch >= 1");
  g->B = d->batch;
  g->R = (long long)d->tt.rows[0] * d->tt.rows[1] * d->tt.rows[2];
  g->Cn = (long long)d->tt.cols[0] * d->tt.cols[1] * d->tt.cols[2];
  TLT_REQUIRE(g->R >= 8 && g->Cn >= 8 && g->R % 8 == 0 && g->Cn % 8 == 0, "tlt_blocktt3_linear: feature counts must be multiples of 8");
  g->fin = d->transpose ? g->Cn : g->R;
  g->fout = d->transpose ? g->R : g->Cn;
  return TLT_OK;
}
static void gemm_desc(tlt_gemm_desc* q, long long M, long long N, long long K, long long lda, long long ldb, long long ldc, int am, int bm, int dtc) {
  memset(q, 0, sizeof(*q));
  q->M = M; q->N = N; q->K = K; q->lda = lda; q->ldb = ldb; q->ldc = ldc; q->a_major = am; q->b_major = bm; q->dtype_c = dtc; q->alpha = 1.f;
}

}  // namespace opl
}  // namespace tlt

using namespace tlt;
using namespace tlt::opl;

// ================================================================================================================ BlockTT linear
extern "C" size_t tlt_blocktt3_linear_workspace_size(const tlt_blocktt3_linear_desc* d) {
  BttGeom g;
  if (btt_geom(d, &g)) return 0;
  return al((size_t)g.R * g.Cn * 2);                                      // W, kept for the backward call
}
extern "C" size_t tlt_blocktt3_linear_grad_workspace_size(const tlt_blocktt3_linear_desc* d) {
  BttGeom g;
  if (btt_geom(d, &g)) return 0;
  return al((size_t)g.R * g.Cn * 4) + al(tlt_blocktt3_core_grads_workspace_size(&d->tt));
}

extern "C" int tlt_blocktt3_linear_fwd(const tlt_blocktt3_linear_desc* d, const void* x, const void* c1, const void* c2, const void* c3,
                                       const void* bias, void* ws, void* y, void* stream) {
  BttGeom g;
  int rc = btt_geom(d, &g);
  if (rc) return rc;
  TLT_REQUIRE(x && c1 && c2 && c3 && ws && y, "tlt_blocktt3_linear_fwd: null pointer");
  rc = tlt_blocktt3_reconstruct(&d->tt, c1, c2, c3, ws, stream);            // W (R, Cn) row-major
  if (rc) return rc;
  tlt_gemm_desc q;
  // transpose: y = x W^T (W rows are K-major);  else y = x W (W is the (k, n) matrix: MN-major B operand)
  if (d->transpose) gemm_desc(&q, g.B, g.fout, g.fin, g.fin, g.fin, g.fout, 0, 0, TLT_BF16);
  else gemm_desc(&q, g.B, g.fout, g.fin, g.fin, g.fout, g.fout, 0, 1, TLT_BF16);
  q.dtype_bias = d->dtype_bias;
  return tlt_gemm_bf16_tc(&q, x, ws, bias, y, stream);
}

// dx (B, in) | NULL; d1..d3 (core gradients, bf16) | all NULL; dbias (out) of dtype dtype_bias | NULL.  gws: grad workspace.
extern "C" int tlt_blocktt3_linear_bwd(const tlt_blocktt3_linear_desc* d, const void* x, const void* dy, const void* ws, const void* c1,
                                       const void* c2, const void* c3, void* gws, void* dx, void* d1, void* d2, void* d3, void* dbias,
                                       void* stream) {
  BttGeom g;
  int rc = btt_geom(d, &g);
  if (rc) return rc;
  TLT_REQUIRE(x && dy && ws && c1 && c2 && c3 && gws, "tlt_blocktt3_linear_bwd: null pointer");
  TLT_REQUIRE((d1 && d2 && d3) || (!d1 && !d2 && !d3), "tlt_blocktt3_linear_bwd: all three core gradients or none");
  tlt_gemm_desc q;
  if (dx) {
    if (d->transpose) gemm_desc(&q, g.B, g.fin, g.fout, g.fout, g.fin, g.fin, 0, 1, TLT_BF16);   // dx = dy W
    else gemm_desc(&q, g.B, g.fin, g.fout, g.fout, g.fout, g.fin, 0, 0, TLT_BF16);               // dx = dy W^T
    rc = tlt_gemm_bf16_tc(&q, dy, ws, nullptr, dx, stream);
    if (rc) return rc;
  }
  if (d1) {
    float* dW = (float*)gws;
    void* cg = (char*)gws + al((size_t)g.R * g.Cn * 4);
    // dW (R, Cn) fp32 (split-K partials meet there): transpose: dW = dy^T x, else dW = x^T dy; both operands MN-major
    if (d->transpose) gemm_desc(&q, g.R, g.Cn, g.B, g.fout, g.fin, g.Cn, 1, 1, TLT_F32);
    else gemm_desc(&q, g.R, g.Cn, g.B, g.fin, g.fout, g.Cn, 1, 1, TLT_F32);
    rc = tlt_gemm_bf16_tc(&q, d->transpose ? dy : x, d->transpose ? x : dy, nullptr, dW, stream);
    if (rc) return rc;
    rc = tlt_blocktt3_core_grads(&d->tt, c1, c2, c3, dW, d1, d2, d3, cg, tlt_blocktt3_core_grads_workspace_size(&d->tt), stream);
    if (rc) return rc;
  }
  if (dbias) {
    colsum_kernel<<<(unsigned)((g.fout + 63) / 64), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy, g.B, (int)g.fout,
                                                                                  d->dtype_bias == TLT_F32, dbias);
    rc = check_launch("colsum_kernel");
  }
  return rc;
}

// ================================================================================================================ Tucker TRL
namespace {
struct TrlGeom {
  long long B, C, S, J, rc, ldz, cp, ro, O;
  int ns;
  size_t off_k, off_t, off_wp, off_zp, off_v, off_kmid, fwd_bytes;          // forward workspace
  size_t off_dv, off_dzp, off_dt, off_dwp, off_dk32, off_dk16, off_dmid, bwd_bytes;
};
int trl_geom(const tlt_tucker_trl_desc* d, TrlGeom* g) {
  TLT_REQUIRE(d && d->batch >= 1 && d->channels >= 8 && d->channels % 8 == 0 && d->n_spatial >= 1 && d->n_spatial <= 3 && d->rc >= 1 &&
              d->ro >= 1 && d->ro <= 16 && d->out >= 1, "tlt_tucker_trl: bad descriptor");
  g->B = d->batch; g->C = d->channels; g->ns = d->n_spatial; g->rc = d->rc; g->ro = d->ro; g->O = d->out;
  g->S = 1; g->J = 1;
  for (int i = 0; i < g->ns; ++i) {
    TLT_REQUIRE(d->spatial[i] >= 1 && d->spatial_rank[i] >= 1, "tlt_tucker_trl: bad spatial mode");
    g->S *= d->spatial[i]; g->J *= d->spatial_rank[i];
  }
  TLT_REQUIRE(tlt_spatial_project_supported(g->B, g->C, (int)g->S, (int)g->J), "tlt_tucker_trl: shape outside the spatial projection kernel");
  g->ldz = r8(g->rc); g->cp = r8(g->C);
  size_t o = 0;
  g->off_k = o;    o += al((size_t)g->S * g->J * 2);
  g->off_kmid = o; o += al((size_t)g->S * g->J * 2);                        // Kronecker intermediate (three spatial modes)
  g->off_t = o;    o += al((size_t)g->B * g->J * g->C * 2);
  g->off_wp = o;   o += al((size_t)g->ldz * g->cp * 2);
  g->off_zp = o;   o += al((size_t)g->B * g->J * g->ldz * 2);
  g->off_v = o;    o += al((size_t)g->B * g->ro * 4);
  g->fwd_bytes = o;
  o = 0;
  g->off_dv = o;   o += al((size_t)g->B * g->ro * 4);
  g->off_dzp = o;  o += al((size_t)g->B * g->J * g->ldz * 2);
  g->off_dt = o;   o += al((size_t)g->B * g->J * g->C * 2);
  g->off_dwp = o;  o += al((size_t)g->rc * g->cp * 4);
  g->off_dk32 = o; o += al((size_t)g->S * g->J * 4);
  g->off_dk16 = o; o += al((size_t)g->S * g->J * 2);
  g->off_dmid = o; o += al((size_t)g->S * g->J * 2);
  g->bwd_bytes = o;
  return TLT_OK;
}
void tail_desc(const TrlGeom& g, tlt_trl_tail_desc* t) {
  t->batch = g.B; t->j = g.J; t->rc = g.rc; t->ldz = g.ldz; t->ro = g.ro; t->out = g.O;
}
}  // namespace

extern "C" size_t tlt_tucker_trl_workspace_size(const tlt_tucker_trl_desc* d) {
  TrlGeom g;
  return trl_geom(d, &g) ? 0 : g.fwd_bytes;
}
extern "C" size_t tlt_tucker_trl_grad_workspace_size(const tlt_tucker_trl_desc* d) {
  TrlGeom g;
  return trl_geom(d, &g) ? 0 : g.bwd_bytes;
}

// factors: [F_c (C, rc), F_s1 (s_1, q_1) .. F_sn (s_n, q_n), F_o (O, ro)], all bf16 contiguous; core (rc, q_1..q_n, ro); x (B, C, s_1..s_n)
extern "C" int tlt_tucker_trl_fwd(const tlt_tucker_trl_desc* d, const void* x, const void* core, const void* const* factors, const void* bias,
                                  void* ws, void* y, void* stream) {
  TrlGeom g;
  int rc = trl_geom(d, &g);
  if (rc) return rc;
  TLT_REQUIRE(x && core && factors && ws && y, "tlt_tucker_trl_fwd: null pointer");
  char* w = (char*)ws;
  const void* k = factors[1];
  if (g.ns >= 2) {                                                           // k = F_s1 (x) F_s2 (x) F_s3
    void* dst = g.ns == 2 ? w + g.off_k : w + g.off_kmid;
    rc = tlt_kron2_fwd(factors[1], factors[2], (int)d->spatial[0], (int)d->spatial_rank[0], (int)d->spatial[1], (int)d->spatial_rank[1], dst, stream);
    if (rc) return rc;
    if (g.ns == 3) {
      rc = tlt_kron2_fwd(dst, factors[3], (int)(d->spatial[0] * d->spatial[1]), (int)(d->spatial_rank[0] * d->spatial_rank[1]),
                         (int)d->spatial[2], (int)d->spatial_rank[2], w + g.off_k, stream);
      if (rc) return rc;
    }
    k = w + g.off_k;
  } else {
    TLT_CHECK_CUDA(cudaMemcpyAsync(w + g.off_k, k, (size_t)g.S * g.J * 2, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    k = w + g.off_k;
  }
  rc = tlt_spatial_project_fwd(g.B, g.C, (int)g.S, (int)g.J, x, k, w + g.off_t, stream);
  if (rc) return rc;
  // F_c^T (rc, C) re-pitched K-major with zero pad rows
  if (g.ldz != g.rc) TLT_CHECK_CUDA(cudaMemsetAsync(w + g.off_wp, 0, (size_t)g.ldz * g.cp * 2, (cudaStream_t)stream));
  rc = tlt_pack_kmajor(factors[0], TLT_BF16, g.rc, g.C, 1, g.rc, w + g.off_wp, g.cp, stream);   // element (r, c) of F_c^T at F_c[c * rc + r]
  if (rc) return rc;
  tlt_gemm_desc q;
  gemm_desc(&q, g.B * g.J, g.ldz, g.C, g.C, g.cp, g.ldz, 0, 0, TLT_BF16);
  rc = tlt_gemm_bf16_tc(&q, w + g.off_t, w + g.off_wp, nullptr, w + g.off_zp, stream);
  if (rc) return rc;
  tlt_trl_tail_desc t;
  tail_desc(g, &t);
  return tlt_trl_tail_fwd(&t, w + g.off_zp, core, factors[g.ns + 1], bias, y, (float*)(w + g.off_v), stream);
}

// d_factors: same order as factors, bf16; dbias (O) bf16 | NULL.  side_stream (or NULL): the tail's parameter gradients and the
// channel-factor gradient run there beside dt -> spatial backward -> Kronecker backward (fork / join by events: capturable).
extern "C" int tlt_tucker_trl_bwd(const tlt_tucker_trl_desc* d, const void* x, const void* dy, const void* core, const void* const* factors,
                                  const void* ws, void* gws, void* dx, void* dcore, void* const* d_factors, void* dbias, void* stream,
                                  void* side_stream) {
  TrlGeom g;
  int rc = trl_geom(d, &g);
  if (rc) return rc;
  TLT_REQUIRE(x && dy && core && factors && ws && gws && dx && dcore && d_factors, "tlt_tucker_trl_bwd: null pointer");
  const char* w = (const char*)ws;
  char* gw = (char*)gws;
  cudaStream_t st = (cudaStream_t)stream, sd = side_stream ? (cudaStream_t)side_stream : st;
  tlt_trl_tail_desc t;
  tail_desc(g, &t);
  const void* fo = factors[g.ns + 1];
  float* dv = (float*)(gw + g.off_dv);
  rc = tlt_trl_tail_bwd(&t, 1, w + g.off_zp, dy, core, fo, (const float*)(w + g.off_v), dv, gw + g.off_dzp, nullptr, nullptr, nullptr, st);
  if (rc) return rc;
  cudaEvent_t fork = nullptr, join = nullptr;
  if (sd != st) {
    TLT_CHECK_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
    TLT_CHECK_CUDA(cudaEventCreateWithFlags(&join, cudaEventDisableTiming));
    TLT_CHECK_CUDA(cudaEventRecord(fork, st));
    TLT_CHECK_CUDA(cudaStreamWaitEvent(sd, fork, 0));
  }
  // ---- parameter branch
  rc = tlt_trl_tail_bwd(&t, 2, w + g.off_zp, dy, core, fo, (const float*)(w + g.off_v), dv, nullptr, dcore, d_factors[g.ns + 1], dbias, sd);
  tlt_gemm_desc q;
  if (!rc) {
    gemm_desc(&q, g.rc, g.C, g.B * g.J, g.ldz, g.C, g.cp, 1, 1, TLT_F32);                       // dF_c^T (rc, C) = dzp^T t
    rc = tlt_gemm_bf16_tc(&q, gw + g.off_dzp, w + g.off_t, nullptr, gw + g.off_dwp, sd);
  }
  if (!rc) {
    const long long n = g.rc * g.C;
    transpose_cast_kernel<<<(unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, sd>>>((const float*)(gw + g.off_dwp), (int)g.rc, (int)g.C,
                                                                                                       (int)g.cp, (__nv_bfloat16*)d_factors[0]);
    rc = check_launch("transpose_cast_kernel");
  }
  // ---- data branch
  if (!rc) {
    gemm_desc(&q, g.B * g.J, g.C, g.ldz, g.ldz, g.cp, g.C, 0, 1, TLT_BF16);                     // dt = dzp F_c^T
    rc = tlt_gemm_bf16_tc(&q, gw + g.off_dzp, w + g.off_wp, nullptr, gw + g.off_dt, st);
  }
  if (!rc) {
    cudaMemsetAsync(gw + g.off_dk32, 0, (size_t)g.S * g.J * 4, st);
    rc = tlt_spatial_project_bwd(g.B, g.C, (int)g.S, (int)g.J, x, w + g.off_k, gw + g.off_dt, dx, (float*)(gw + g.off_dk32), st);
  }
  if (!rc) {
    if (g.ns == 1) {
      rc = tlt_cast(gw + g.off_dk32, TLT_F32, d_factors[1], TLT_BF16, g.S * g.J, st);
    } else {
      rc = tlt_cast(gw + g.off_dk32, TLT_F32, gw + g.off_dk16, TLT_BF16, g.S * g.J, st);
      if (!rc && g.ns == 2)
        rc = tlt_kron2_bwd(factors[1], factors[2], gw + g.off_dk16, (int)d->spatial[0], (int)d->spatial_rank[0], (int)d->spatial[1],
                           (int)d->spatial_rank[1], d_factors[1], d_factors[2], st);
      if (!rc && g.ns == 3) {
        const int s01 = (int)(d->spatial[0] * d->spatial[1]), q01 = (int)(d->spatial_rank[0] * d->spatial_rank[1]);
        rc = tlt_kron2_bwd(w + g.off_kmid, factors[3], gw + g.off_dk16, s01, q01, (int)d->spatial[2], (int)d->spatial_rank[2], gw + g.off_dmid,
                           d_factors[3], st);
        if (!rc)
          rc = tlt_kron2_bwd(factors[1], factors[2], gw + g.off_dmid, (int)d->spatial[0], (int)d->spatial_rank[0], (int)d->spatial[1],
                             (int)d->spatial_rank[1], d_factors[1], d_factors[2], st);
      }
    }
  }
  if (sd != st) {
    cudaEventRecord(join, sd);
    cudaStreamWaitEvent(st, join, 0);
    cudaEventDestroy(fork);
    cudaEventDestroy(join);
  }
  return rc;
}
