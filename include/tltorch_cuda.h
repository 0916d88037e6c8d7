/*
 * tltorch_cuda.h -- C ABI of the B200-native TensorLy-Torch factorized-layer hot path (libtltorch_cuda.so).
 *
 * The reference (tensorly/torch 0.5.0) has no FFI layer: its hot path bottoms out in library calls issued from
 * tltorch/functional/*.py (torch.einsum through tensorly, F.conv{1,2,3}d, torch.matmul).  Each entry point below
 * replaces one family of those call sites; the reference file:line it stands in for is cited with it.  The
 * reference-side binding (a ctypes stub) is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C: raw device pointers, sizes and strides in ELEMENTS, no torch types;
 *   - every function returns 0 on success or a negative tlt_status; tlt_last_error_string() describes the failure
 *     (thread-local); nothing ever throws across the ABI;
 *   - the caller allocates every input / output / workspace buffer and owns it; the library keeps no per-call
 *     state (only immutable caches of kernel attributes and TMA descriptors keyed by shape);
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *   - dtype codes: TLT_F32 / TLT_BF16; accumulation is always fp32.
 */
#ifndef TLTORCH_CUDA_H_
#define TLTORCH_CUDA_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TLT_ABI_VERSION 1
#define TLT_MAX_SUBDIMS 4

typedef enum { TLT_F32 = 0, TLT_BF16 = 1 } tlt_dtype;

typedef enum {
  TLT_OK = 0,
  TLT_ERR_INVALID = -1,     /* bad descriptor / unsupported shape          */
  TLT_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed         */
  TLT_ERR_WORKSPACE = -3,   /* workspace too small                         */
  TLT_ERR_UNSUPPORTED = -4  /* valid request the kernel does not implement */
} tlt_status;

int tlt_version(void);
const char* tlt_last_error_string(void);
/* Number of kernel launches issued through this library since load (bench.py's "gpu_launches" evidence). */
uint64_t tlt_launch_count(void);

/* --------------------------------------------------------------------------------------------------------------
 * tlt_contract: generic pairwise tensor contraction  C[b,m,n] = alpha * sum_k A[b,m,k] * B[b,k,n] (+ D[b,m,n])
 *
 * Each logical index class (batch b, rows m, cols n, reduction k) is a multi-index of up to TLT_MAX_SUBDIMS
 * sub-dimensions, each with its own stride per operand, so mode-n products and TT / Tucker / CP sweeps run on the
 * tensors where they lie -- no permute/contiguous copies.  Sub-dimension 0 is the OUTERMOST of its class.
 * Replaces: every pairwise step of tl.einsum in tltorch/functional/factorized_linear.py:59 and
 * factorized_tensordot.py:66,102; tenalg.mode_dot / multi_mode_dot (tensor_regression.py:40-42, convolution.py:160,
 * factorized_layers/tensor_contraction_layers.py:75); tenalg.inner (tensor_regression.py:32-49); the einsum loop of
 * BlockTT.to_tensor (factorized_tensors/tensorized_matrices.py:382); tl.{cp,tucker,tt}_to_tensor
 * (factorized_tensors/factorized_tensors.py:130,308,415); F.conv1d(k=1) channel mixing (convolution.py:154,168,207,
 * 223,262,278,311,340); the lambda scaling x*weights (convolution.py:278,340) and the "+ bias" elementwise adds
 * (linear.py:32-43).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t dtype_a, dtype_b, dtype_c, dtype_d;           /* tlt_dtype; dtype_d ignored when D == NULL               */
  int32_t nb, nm, nn, nk;                               /* number of sub-dims per class (0 => class has extent 1)  */
  int64_t size_b[TLT_MAX_SUBDIMS], size_m[TLT_MAX_SUBDIMS], size_n[TLT_MAX_SUBDIMS], size_k[TLT_MAX_SUBDIMS];
  int64_t a_b[TLT_MAX_SUBDIMS], a_m[TLT_MAX_SUBDIMS], a_k[TLT_MAX_SUBDIMS];    /* strides of A                      */
  int64_t b_b[TLT_MAX_SUBDIMS], b_k[TLT_MAX_SUBDIMS], b_n[TLT_MAX_SUBDIMS];    /* strides of B                      */
  int64_t c_b[TLT_MAX_SUBDIMS], c_m[TLT_MAX_SUBDIMS], c_n[TLT_MAX_SUBDIMS];    /* strides of C                      */
  int64_t d_b[TLT_MAX_SUBDIMS], d_m[TLT_MAX_SUBDIMS], d_n[TLT_MAX_SUBDIMS];    /* strides of D (0 = broadcast)      */
  float alpha;
  int32_t accumulate;                                   /* 1: C += result (C must be fp32)                         */
  int32_t split_k;                                      /* 0 = library chooses; >1 needs fp32 C or a workspace     */
} tlt_contract_desc;

/* Bytes of fp32 workspace tlt_contract needs for this descriptor (0 when none). */
size_t tlt_contract_workspace_size(const tlt_contract_desc* desc);
int tlt_contract(const tlt_contract_desc* desc, const void* A, const void* B, const void* D, void* C,
                 void* workspace, size_t workspace_bytes, void* stream);

/* Bridge to the tensor cores for contractions whose operands are strided or break TMA's 16-byte pitch rule:
 *   tlt_contract_pack   copies operand A (operand = 0: rows = the m class) or B (operand = 1: rows = the n class), read
 *                       through the descriptor's strides, into a dense K-major bf16 matrix rows x K with pitch dst_ld;
 *   tlt_contract_unpack writes a dense M x N result of dtype_cp (pitch ldc; fp32 when the GEMM ran split-K) (+ the strided
 *                       addend D) to the strided output of dtype_c.
 * Used as pack -> tlt_gemm_bf16_tc -> unpack; only for descriptors with nb == 0.  Same reference call sites as tlt_contract. */
int tlt_contract_pack(const tlt_contract_desc* desc, int32_t operand, const void* src, void* dst, int64_t dst_ld, void* stream);
int tlt_contract_unpack(const tlt_contract_desc* desc, const void* Cp, int32_t dtype_cp, int64_t ldc, const void* D, void* C,
                        void* stream);

/* --------------------------------------------------------------------------------------------------------------
 * tlt_gemm_bf16_tc: dense bf16 GEMM on the tcgen05 tensor cores (TMA-staged operands, TMEM fp32 accumulators).
 *   C[M,N] = A[M,K] * B[N,K]^T (+ bias[N])      A, B bf16;  C bf16 or fp32, row-major with leading dim ldc.
 * a_major / b_major: 0 = K-major (element (i,k) at i*ld + k), 1 = MN-major (element (i,k) at k*ld + i).
 * Requirements: the contiguous extent and ld of A and B are multiples of 8 elements; base pointers 16-byte aligned.
 * Replaces: F.linear / torch.matmul on reconstructed or dense-regime weights (functional/linear.py:50,
 * tensor_regression.py:47-49), and the large Tucker-core / TT stage GEMMs inside the einsums cited above.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t M, N, K;
  int64_t lda, ldb, ldc;
  int32_t a_major, b_major;
  int32_t dtype_c;                                      /* TLT_BF16 or TLT_F32                                     */
  int32_t accumulate;                                   /* 1: C += A*B^T (fp32 C only)                             */
  int32_t dtype_bias;                                   /* dtype of bias[N] (ignored when bias == NULL)            */
  float alpha;
  int32_t split_k;                                      /* 0 = library chooses; > 1 only honoured for fp32 C (partials
                                                           meet in C through TMA reduce-add; C is zeroed by the call
                                                           unless accumulate)                                       */
} tlt_gemm_desc;

int tlt_gemm_bf16_tc_supported(const tlt_gemm_desc* desc);      /* 1 if the tensor-core kernel accepts this problem */
int tlt_gemm_bf16_tc(const tlt_gemm_desc* desc, const void* A, const void* B, const void* bias, void* C, void* stream);

/* --------------------------------------------------------------------------------------------------------------
 * Fused 3-core BlockTT (TT-matrix) kernels.
 *   W[(r1,r2,r3),(c1,c2,c3)] = sum_{a,b} C1[0,r1,c1,a] * C2[a,r2,c2,b] * C3[b,r3,c3,0]      W row-major (prod rows, prod cols)
 * cores are contiguous (rank_k, rows_k, cols_k, rank_{k+1}) tensors of dtype `dtype`; accumulation is fp32.
 * Replaces: BlockTT.to_tensor / to_matrix (tltorch/factorized_tensors/tensorized_matrices.py:354-384, core.py:536-543)
 * -- the pairwise einsum loop 'adeb,bfgc->adfegc' + permute -- in ONE launch with the rank-indexed intermediate kept in
 * shared memory, and its autograd (tlt_blocktt3_core_grads: the dense-matrix gradient contracted back onto the cores).
 * Returns TLT_ERR_UNSUPPORTED when the intermediate does not fit shared memory (callers fall back to tlt_contract).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t dtype;                                        /* dtype of the cores and of W / the core gradients        */
  int32_t rows[3], cols[3];                             /* tensorized shape of the matrix                          */
  int32_t rank1, rank2;                                 /* interior TT ranks (boundary ranks are 1)                */
} tlt_blocktt3_desc;

int tlt_blocktt3_reconstruct(const tlt_blocktt3_desc* desc, const void* c1, const void* c2, const void* c3, void* W,
                             void* stream);
size_t tlt_blocktt3_core_grads_workspace_size(const tlt_blocktt3_desc* desc);
/* dW: fp32 (prod rows, prod cols) row-major, 16-byte aligned; d1..d3: gradients of the cores, dtype `dtype` */
int tlt_blocktt3_core_grads(const tlt_blocktt3_desc* desc, const void* c1, const void* c2, const void* c3,
                            const float* dW, void* d1, void* d2, void* d3, void* workspace, size_t workspace_bytes,
                            void* stream);

/* --------------------------------------------------------------------------------------------------------------
 * Fused chain of mode products on the trailing (up to) three modes of a contiguous tensor, intermediates in shared memory:
 *   x (batch, d0, d1, d2) -> y (batch, r0, r1, r2),  y[b] = x[b] x_0 M0 x_1 M1 x_2 M2
 * M_k is (r_k, d_k) row-major, or (d_k, r_k) when transpose[k] != 0 (multiply by M_k^T); a NULL matrix leaves its mode
 * untouched (in_dims[k] == out_dims[k]).  Tensors with fewer modes pad LEADING modes with size 1 and NULL matrices.
 * Returns TLT_ERR_UNSUPPORTED when one sample does not fit shared memory (callers fall back to per-mode tlt_contract).
 * Replaces: tenalg.multi_mode_dot chains -- TCL.forward (factorized_layers/tensor_contraction_layers.py:75), the input
 * projection of tucker_trl (functional/tensor_regression.py:40) and the project / expand stages of linear_tucker
 * (functional/factorized_tensordot.py:10-66) -- and their data gradients (same call, transposed matrices).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t dtype;
  int64_t batch;
  int32_t in_dims[3], out_dims[3];
  int32_t transpose[3];
} tlt_mmd_desc;

/* addend: optional (r0, r1, r2) tensor of the same dtype added to every sample (a bias), or NULL */
int tlt_multi_mode_dot(const tlt_mmd_desc* desc, const void* x, const void* m0, const void* m1, const void* m2,
                       const void* addend, void* y, void* stream);

/* --------------------------------------------------------------------------------------------------------------
 * Tensor-core variant of the fused mode-product chain for bf16 samples whose three modes are all <= 16 (tensorized
 * linear layers: 4096 = 16^3, Tucker ranks 15^3, ...).  One warp keeps one sample on chip as a 16x16x16 tile; every mode
 * product is an m16n8k16 HMMA fed by ldmatrix; no block-level barriers.
 *   tlt_mmd16_fwd : y[b] = x[b] x_0 M0 x_1 M1 x_2 M2 (+ addend)       x (batch, in_dims) pitch ld_in -> y (batch, out_dims) pitch ld_out
 *   tlt_mmd16_bwd : du[b] = g[b] x_0 M0^T x_1 M1^T x_2 M2^T  and  dM[k][j][i] += d<g, y>/dM_k[j][i]   in one pass over (u, g);
 *                   dM is fp32 [3][16][16] (j = out index, i = in index), accumulated into (the caller zeroes it); du or dM
 *                   may be NULL to skip that half.
 * M_k is (out_dims[k], in_dims[k]) row-major, or (in_dims[k], out_dims[k]) when transpose[k] != 0; NULL = identity.
 * Sample pitches are in elements, multiples of 8, >= round8(prod dims); the pad elements [prod, round8(prod)) of every
 * written sample are set to zero, so the tensors can feed tlt_gemm_bf16_tc as K-major operands directly.
 * Replaces: the project-in / expand-out multi_mode_dot chains of linear_tucker (tltorch/functional/factorized_linear.py:7-21,
 * factorized_tensordot.py:10-66) and their autograd (data gradient + all three factor gradients in one launch).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t batch;
  int32_t in_dims[3], out_dims[3];
  int32_t transpose[3];
  int64_t ld_in, ld_out;
} tlt_mmd16_desc;

int tlt_mmd16_fwd(const tlt_mmd16_desc* desc, const void* x, const void* m0, const void* m1, const void* m2,
                  const void* addend, void* y, void* stream);
int tlt_mmd16_bwd(const tlt_mmd16_desc* desc, const void* u, const void* g, const void* m0, const void* m1, const void* m2,
                  void* du, float* dM, void* stream);

/* --------------------------------------------------------------------------------------------------------------
 * Pointwise (1x1) channel mixing of channels-first bf16 activations on the tcgen05 tensor cores, from NCHW memory.
 *   tlt_pointwise_tc       : y[b, n, s] = sum_k wt[n, k] * x[b, k, s] (+ bias[n])      x (batch, in, spatial), y (batch, out, spatial)
 *   tlt_pointwise_wgrad_tc : dw[n, k]   = sum_{b, s} dy[b, n, s] * x[b, k, s]          dw fp32, row pitch w_ld
 * wt is a K-major (out_channels, w_ld) bf16 matrix (tlt_pack_kmajor builds it from any strided view; the columns past
 * in_channels are zero).  spatial must be a multiple of 8 (TMA pitch): tlt_pointwise_tc_supported() tells; other shapes
 * stay on tlt_contract.  The data gradient is the same call with the transposed weight.
 * Replaces: F.conv1d(kernel_size=1) in tltorch/functional/convolution.py (:154,168 tucker_conv, :207,223 tt_conv,
 * :262,278 cp_conv, :311,340 cp_conv_mobilenet) and its autograd.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t batch, in_channels, out_channels, spatial;
  int64_t w_ld;                                         /* pitch (elements) of wt (bf16) / of dw (fp32)             */
  int32_t dtype_bias;                                   /* tlt_dtype of bias[out_channels] (ignored when NULL)      */
} tlt_pointwise_desc;

int tlt_pointwise_tc_supported(const tlt_pointwise_desc* desc);
int tlt_pointwise_tc(const tlt_pointwise_desc* desc, const void* x, const void* wt, const void* bias, void* y, void* stream);
int tlt_pointwise_wgrad_tc(const tlt_pointwise_desc* desc, const void* dy, const void* x, float* dw, void* stream);
/* dst[r, c] = (bf16) src[r * row_stride + c * col_stride] for c < cols, 0 for cols <= c < dst_ld   (rows x dst_ld, bf16) */
int tlt_pack_kmajor(const void* src, int32_t dtype_src, int64_t rows, int64_t cols, int64_t row_stride, int64_t col_stride,
                    void* dst, int64_t dst_ld, void* stream);

/* --------------------------------------------------------------------------------------------------------------
 * Convolution building blocks (channels-first, 1 to 3 spatial dims).
 * Replaces: F.conv{1,2,3}d call sites of tltorch/functional/convolution.py -- :133 (general_conv1d, depthwise for CP
 * :268-271 and full for TT :213-216), :161 (Tucker core conv), :321-332 (mobilenet depthwise N-D), and their autograd.
 * Dense (groups == 1) convolutions run as unfold (im2col) + tlt_contract / tlt_gemm_bf16_tc; depthwise
 * (groups == channels) ones run in the dedicated kernels.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t dtype;
  int32_t ndim;                                         /* spatial dims, 1..3                                      */
  int64_t batch, channels;
  int64_t in_size[3], out_size[3];
  int32_t kernel[3], stride[3], padding[3], dilation[3];
} tlt_conv_desc;

/* cols[(b, o...), (c, t...)] = x[b, c, o*stride - pad + t*dil]  (zero outside)  -- row-major (B*prod(out), C*prod(kernel)) */
int tlt_im2col(const tlt_conv_desc* desc, const void* x, void* cols, void* stream);
/* dx[b, c, i...] = sum over (o,t) with o*stride - pad + t*dil == i of dcols[(b,o),(c,t)]   (gather form, no atomics) */
int tlt_col2im(const tlt_conv_desc* desc, const void* dcols, void* dx, void* stream);
/* depthwise: y[b,c,o] = sum_t x[b,c,o*s-p+t*d] * w[c,t]; w is (channels, prod(kernel)) row-major */
int tlt_dwconv_fwd(const tlt_conv_desc* desc, const void* x, const void* w, void* y, void* stream);
int tlt_dwconv_dgrad(const tlt_conv_desc* desc, const void* dy, const void* w, void* dx, void* stream);
/* dw must be fp32 (channels, prod(kernel)), zero-initialised by the caller; the kernel accumulates into it */
int tlt_dwconv_wgrad(const tlt_conv_desc* desc, const void* x, const void* dy, float* dw, void* stream);

/* --------------------------------------------------------------------------------------------------------------
 * Channels-last ("rows") staging of bf16 activations for dense k x k convolutions on tlt_gemm_bf16_tc.
 *   tlt_nchw_to_rows : rows[(b, s), c] = x[b, c, s], c in [channels, ld) zero-filled       rows (batch * spatial, ld)
 *   tlt_rows_to_nchw : y[b, c, s] = rows[(b, s), c]
 *   tlt_im2col_rows  : cols[(b, o...), (t..., c)] = z[(b, o*stride - pad + t*dil), c] (zero outside); desc->channels is the
 *                      row length (a multiple of 8) of z (batch * prod(in_size), channels); cols is
 *                      (batch * prod(out_size), prod(kernel) * channels) -- 16-byte chunks, coalesced on both sides
 *   tlt_col2im_rows  : dz[(b, i...), c] = sum over (o, t) with o*stride - pad + t*dil == i of dcols[(b, o), (t, c)]   (gather)
 * Replaces: the unfold / fold halves of F.conv{1,2,3}d with a dense kernel (tltorch/functional/convolution.py:161 Tucker
 * core conv, :213-216 TT cores, :16-53 reconstructed weights) when the surrounding 1x1 stages run as row GEMMs.
 * ------------------------------------------------------------------------------------------------------------ */
int tlt_nchw_to_rows(const void* x, int64_t batch, int64_t channels, int64_t spatial, void* rows, int64_t ld, void* stream);
int tlt_rows_to_nchw(const void* rows, int64_t batch, int64_t channels, int64_t spatial, int64_t ld, void* y, void* stream);
int tlt_im2col_rows(const tlt_conv_desc* desc, const void* z, void* cols, void* stream);
int tlt_col2im_rows(const tlt_conv_desc* desc, const void* dcols, void* dz, void* stream);

/* --------------------------------------------------------------------------------------------------------------
 * Depthwise convolution on channels-last rows (bf16; desc->channels = row length, a multiple of 8; 1..3 spatial dims).
 * w is taps-major: (prod(kernel), channels) bf16.  dw is fp32 (prod(kernel), channels), accumulated into (caller zeroes);
 * the weight gradient supports up to 9 taps (TLT_ERR_UNSUPPORTED otherwise).
 * Replaces: the depthwise F.conv1d chain of cp_conv (tltorch/functional/convolution.py:268-271; merged into one N-D
 * depthwise conv like cp_conv_mobilenet :321-332) and its autograd, when the surrounding 1x1 stages run as row GEMMs.
 * ------------------------------------------------------------------------------------------------------------ */
int tlt_dwconv_rows_fwd(const tlt_conv_desc* desc, const void* z, const void* w, void* y, void* stream);
int tlt_dwconv_rows_dgrad(const tlt_conv_desc* desc, const void* gy, const void* w, void* dz, void* stream);
int tlt_dwconv_rows_wgrad(const tlt_conv_desc* desc, const void* z, const void* gy, float* dw, void* stream);

/* --------------------------------------------------------------------------------------------------------------
 * Skinny row transform for many rows and tiny extents (I, J <= 32): the core expansion of the factorized convolutions.
 *   tlt_skinny_fwd : y[r, j] = sum_i a[r, i] k[j, i]                     a (rows, I), k (J, I), y (rows, J), dtype f32 / bf16
 *   tlt_skinny_bwd : da[r, i] = sum_j g[r, j] k[j, i];  dk[j, i] += sum_r g[r, j] a[r, i]   (dk fp32, caller zeroes; da or dk
 *                    may be NULL)
 * Replaces: tenalg.multi_mode_dot(core, factors[2:]) of tucker_conv (tltorch/functional/convolution.py:160) with k the
 * Kronecker product of the kernel-mode factors, and its autograd.
 * ------------------------------------------------------------------------------------------------------------ */
int tlt_skinny_fwd(int32_t dtype, int64_t rows, int32_t I, int32_t J, const void* a, const void* k, void* y, void* stream);
int tlt_skinny_bwd(int32_t dtype, int64_t rows, int32_t I, int32_t J, const void* a, const void* g, const void* k, void* da,
                   float* dk, void* stream);

/* --------------------------------------------------------------------------------------------------------------
 * tlt_scale_modes: y[i_0..i_{n-1}] = alpha * prod_k s_k[i_k] * x[i_0..i_{n-1}]  for a contiguous tensor of 1..6 modes;
 * scales[k] is a vector of dims[k] elements of the same dtype, or NULL to leave mode k untouched.
 * Replaces: the mask application of tensor dropout (tltorch/tensor_hooks/_tensor_dropout.py:56-142) in fixed-shape form:
 * the 0/1 keep vectors of the dropped rank modes and the 1/(1-p) rescale multiplied into the core / weights / TT cores
 * in one pass, so no index vector reaches the host and the training step stays CUDA-graph capturable.
 * ------------------------------------------------------------------------------------------------------------ */
int tlt_scale_modes(int32_t dtype, int32_t ndim, const int64_t* dims, const void* x, const void* const* scales, float alpha,
                    void* y, void* stream);

/* --------------------------------------------------------------------------------------------------------------
 * Weighted Khatri-Rao product of 1..4 factor matrices:  KR[(i_1..i_n), r] = w_r * prod_k F_k[i_k, r]
 * F_k (dims[k], R) row-major, weights (R) or NULL, out (prod dims, R); dtype f32 / bf16.
 *   tlt_khatri_rao_bwd: dF[k] (dims[k], R) fp32 (written, NULL to skip; dF[0] is required when dweights != NULL) and
 *   dweights (R) fp32 (ACCUMULATED: caller zeroes) from g (prod dims, R).
 * Replaces: the Khatri-Rao operands of tensor_dot_cp / linear_cp (tltorch/functional/factorized_tensordot.py:70-103) and of
 * CPTensor.to_tensor (factorized_tensors/factorized_tensors.py:129-130), including the lambda scaling and its gradient.
 * ------------------------------------------------------------------------------------------------------------ */
int tlt_khatri_rao_fwd(int32_t dtype, int32_t n, const int64_t* dims, int64_t R, const void* const* factors, const void* weights,
                       void* out, void* stream);
int tlt_khatri_rao_bwd(int32_t dtype, int32_t n, const int64_t* dims, int64_t R, const void* const* factors, const void* weights,
                       const void* g, float* const* dF, float* dweights, void* stream);
/* same, with g at a leading dimension ldg >= R (e.g. the fp32 output of a GEMM with rows padded to a 16-byte pitch) */
int tlt_khatri_rao_bwd_ld(int32_t dtype, int32_t n, const int64_t* dims, int64_t R, const void* const* factors, const void* weights,
                          const void* g, int64_t ldg, float* const* dF, float* dweights, void* stream);
/* all factor gradients (and the weight gradient) of one Khatri-Rao product in ONE launch; dF[k] may be NULL */
int tlt_khatri_rao_bwd_all(int32_t dtype, int32_t n, const int64_t* dims, int64_t R, const void* const* factors, const void* weights,
                           const void* g, int64_t ldg, float* const* dF, float* dweights, void* stream);
/* fp32 factors -> the (weighted) Khatri-Rao matrix written directly as 3-way bf16 split operands (layouts / patterns of
 * tlt_split3_bf16 below), no fp32 matrix in between */
int tlt_khatri_rao_split3(int32_t n, const int64_t* dims, int64_t R, const void* const* factors, const void* weights,
                          int32_t pattern_cols, void* out_cols, int32_t pattern_rows, void* out_rows, void* stream);
/* fp32 -> 3-way bf16 split operand (hi / mid / lo pieces repeated over six reduction blocks: pattern 0 = A side, 1 = B side) so that
 * an fp32 product runs as ONE tlt_gemm_bf16_tc launch with K six blocks long, fp32 accumulation (~1e-6 relative error).  out_cols:
 * (rows, 6 * round8(cols)) for a K-major operand, out_rows: (6 * rows, round8(cols)) for an MN-major operand; either may be NULL.
 * Replaces: the fp32 torch.einsum / matmul products of linear_cp (tltorch/functional/factorized_linear.py:23-36,
 * factorized_tensordot.py:70-103) at BASELINE config 1, whose 1e-5 tolerance rules plain bf16 / TF32 products out. */
int tlt_split3_bf16(const float* src, int64_t rows, int64_t cols, int64_t ld, int32_t pattern_cols, void* out_cols,
                    int32_t pattern_rows, void* out_rows, void* stream);

/* --------------------------------------------------------------------------------------------------------------
 * Separable 3x3 depthwise convolution on channels-last rows (bf16, padding 1, stride 1 or 2): rows (B H W, ld), ld = round8(R),
 *   z2[n, ho, wo, r] = sum_{i,j} kh[i, r] (lambda_r kw[j, r]) z1[n, s ho + i - 1, s wo + j - 1, r]
 * Both separable passes in ONE sweep (csrc/dwsep_rows.cu): the middle stage of cp_conv where the chain is not fused.
 *   tlt_dwsep_rows_taps      taps [6][ld] fp32 <- kh (3, R), kw (3, R), lambda (R) bf16 (flip = 1: both triples reversed)
 *   tlt_dwsep_rows_fwd       z2 <- z1, taps
 *   tlt_dwsep_rows_dgrad     dz1 <- dz2, taps (stride 1: the FLIPPED taps; stride 2: the unflipped ones)
 *   tlt_dwsep_rows_wgrad     acc [parts][9][ld] fp32 <- z1, dz2: per-block partial sums of dz2 * z1(+i-1, +j-1) at [3 i + j][r];
 *                            parts = tlt_dwsep_rows_wgrad_parts() (the call writes every element)
 *   tlt_dwsep_rows_finalize  bf16 gradients of kh, kw, lambda <- the partial sums
 * Replaces: the per-mode depthwise convolutions of cp_conv (tltorch/functional/convolution.py:251-269) and their autograd, and
 * the Khatri-Rao product of the kernel-mode factors that the merged-tap kernels needed.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t batch, height, width, ld;
  int32_t stride;
} tlt_dwsep_desc;
int tlt_dwsep_rows_taps(const void* kh, const void* kw, const void* lambda, int64_t rank, int64_t ld, int32_t flip, float* taps, void* stream);
int tlt_dwsep_rows_fwd(const tlt_dwsep_desc* desc, const void* z, const float* taps, void* y, void* stream);
int tlt_dwsep_rows_dgrad(const tlt_dwsep_desc* desc, const void* gy, const float* taps, void* dz, void* stream);
int tlt_dwsep_rows_wgrad_parts(void);
int tlt_dwsep_rows_wgrad(const tlt_dwsep_desc* desc, const void* z, const void* gy, float* acc, void* stream);
int tlt_dwsep_rows_finalize(const float* acc, int32_t parts, const void* kh, const void* kw, const void* lambda, int64_t rank, int64_t ld,
                            void* g_kh, void* g_kw, void* g_lambda, void* stream);

/* --------------------------------------------------------------------------------------------------------------
 * Operand preparation / gradient finalisation of the unfused CP-convolution chain on rows (csrc/cpchain.cu):
 *   tlt_cpchain_prep  wp_in [round8(R)][round8(C_in)] = F_in^T, wp_out [round8(C_out)][round8(R)] = F_out (bf16, zero pads),
 *                     taps_f / taps_d [6][round8(R)] fp32 for tlt_dwsep_rows_fwd / _dgrad  -- one launch
 *   tlt_cpchain_post  bf16 gradients of F_out (C_out, R), F_in (C_in, R), kh, kw, lambda from the fp32 results of the two
 *                     weight-gradient GEMMs (dw_out [C_out][round8(R)], dw_in [C_in][round8(R)]) and the sweep's slabs -- one launch
 * Replaces the per-op weight re-pitching, pad zero-fills, fp32 -> bf16 casts and slicing copies around the chain's GEMMs.
 * ------------------------------------------------------------------------------------------------------------ */
int tlt_cpchain_prep(int64_t c_in, int64_t c_out, int64_t rank, int32_t flip_dgrad_taps, const void* f_in, const void* f_out,
                     const void* kh, const void* kw, const void* lambda, void* wp_in, void* wp_out, float* taps_f, float* taps_d,
                     void* stream);
int tlt_cpchain_post(int64_t c_in, int64_t c_out, int64_t rank, const float* dw_out, const float* dw_in, const float* acc,
                     int32_t parts, const void* kh, const void* kw, const void* lambda, void* g_fout, void* g_fin, void* g_kh,
                     void* g_kw, void* g_lambda, void* stream);

/* --------------------------------------------------------------------------------------------------------------
 * Op-level fused CP convolution on channels-last rows (bf16): x (B, H, W, C_in) -> y (B, H/s, W/s, C_out),
 *     y = F_out . dw3x3( F_in^T . x )  with separable taps kh (3, R), kw (3, R) and CP weights lambda (R), padding 1.
 * ONE launch per direction; the R-channel intermediates stay in TMEM / shared memory (csrc/cpconv_rows.cu).
 *   tlt_cpconv_rows_workspace_size      bytes of the packed-operand workspace `ws`
 *   tlt_cpconv_rows_grad_workspace_size bytes of the fp32 gradient accumulators `acc`
 *   tlt_cpconv_rows_pack      ws <- (F_in (C_in, R), F_out (C_out, R), kh, kw, lambda), all bf16: tile-padded operand matrices
 *   tlt_cpconv_rows_fwd       y <- x, ws (+ fp32 bias (C_out) or NULL)
 *   tlt_cpconv_rows_bwd       dx <- x, dy, ws; acc <- fp32 sums for every parameter gradient (zeroed by the call)   [stride 1]
 *   tlt_cpconv_rows_finalize  bf16 gradients of F_out, F_in, kh, kw, lambda <- acc
 * Supported: W = 56, H even; (C_in, C_out, stride) = (64, 64, 1) forward + backward, (64, 128, 2) forward; R <= 128.
 * Replaces: cp_conv / cp_conv_mobilenet (tltorch/functional/convolution.py:231-345: 1x1 conv -> one depthwise 1-D conv per
 * kernel mode -> 1x1 conv, each a full HBM round trip of the R-channel tensor) and their autograd, for BASELINE config 3's
 * HBM-bound 56^2 layers.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t batch, in_channels, out_channels, rank, height, width;
  int32_t stride;
} tlt_cpconv_rows_desc;
int tlt_cpconv_rows_supported(const tlt_cpconv_rows_desc* desc);
size_t tlt_cpconv_rows_workspace_size(const tlt_cpconv_rows_desc* desc);
size_t tlt_cpconv_rows_grad_workspace_size(const tlt_cpconv_rows_desc* desc);
int tlt_cpconv_rows_pack(const tlt_cpconv_rows_desc* desc, const void* f_in, const void* f_out, const void* kh, const void* kw,
                         const void* lambda, void* ws, void* stream);
int tlt_cpconv_rows_fwd(const tlt_cpconv_rows_desc* desc, const void* x, const void* ws, const float* bias, void* y, void* stream);
int tlt_cpconv_rows_bwd(const tlt_cpconv_rows_desc* desc, const void* x, const void* dy, const void* ws, void* dx, float* acc,
                        void* stream);
int tlt_cpconv_rows_finalize(const tlt_cpconv_rows_desc* desc, const float* acc, const void* kh, const void* kw, const void* lambda,
                             void* g_fout, void* g_fin, void* g_kh, void* g_kw, void* g_lambda, void* stream);

/* --------------------------------------------------------------------------------------------------------------
 * Spatial-modes-first projection of channels-first feature maps (bf16):
 *   tlt_spatial_project_fwd : t[b, j, c] = sum_s x[b, c, s] k[s, j]            x (B, C, S), k (S, J) -> t (B, J, C)
 *   tlt_spatial_project_bwd : dx[b, c, s] = sum_j gt[b, j, c] k[s, j];   dk[s, j] += sum_{b,c} x[b, c, s] gt[b, j, c]
 *                             (dk fp32 (S, J), ACCUMULATED: caller zeroes)
 * S = prod(spatial extents) <= 64, J = prod(spatial ranks) <= 16, C % 128 == 0; k is the Kronecker product of the spatial
 * factors.  One pass over the activation where it lies (TMA bulk tiles of 128 channels, HMMA), t comes out channels-last for
 * the channel-mode GEMM.  Replaces: the spatial modes of tenalg.multi_mode_dot(x, factors[:n], transpose=True) in tucker_trl
 * (tltorch/functional/tensor_regression.py:40) and of TCL.forward (factorized_layers/tensor_contraction_layers.py:75) on
 * (B, 2048, 7, 7)-like maps, and their autograd.
 * ------------------------------------------------------------------------------------------------------------ */
int tlt_spatial_project_supported(int64_t batch, int64_t channels, int32_t S, int32_t J);
int tlt_spatial_project_fwd(int64_t batch, int64_t channels, int32_t S, int32_t J, const void* x, const void* k, void* t,
                            void* stream);
int tlt_spatial_project_bwd(int64_t batch, int64_t channels, int32_t S, int32_t J, const void* x, const void* k, const void* gt,
                            void* dx, float* dk, void* stream);

/* --------------------------------------------------------------------------------------------------------------
 * Small utilities used by the host side.
 * ------------------------------------------------------------------------------------------------------------ */
/* dst[i] = (dst dtype) src[i], n elements */
int tlt_cast(const void* src, int32_t dtype_src, void* dst, int32_t dtype_dst, int64_t n, void* stream);
/* Writes `bytes` bytes of zeros/garbage through L2 (bench.py's L2 flush between timed iterations). */
int tlt_l2_flush(void* scratch, size_t bytes, void* stream);

/* --------------------------------------------------------------------------------------------------------------
 * tlt_cplinear_*: the CP-factorized linear layer as ONE op per direction (fp32)
 *     y[b, o] = sum_r ( sum_i x[b, i] prod_m U_m[i_m, r] ) lambda_r prod_m V_m[o_m, r] + bias[o]
 * Replaces: linear_cp (tltorch/functional/factorized_linear.py:23-36) -> tensor_dot_cp (factorized_tensordot.py:70-103),
 * i.e. the n-ary tl.einsum of the activation with the CP factors, and its autograd.  The Khatri-Rao matrices are rebuilt
 * tile by tile in shared memory, never materialised.  Factors are (d_m, R) row-major fp32, in the order of the
 * tensorized shape (first mode slowest); x is (B, I = prod in_dims), y is (B, O = prod out_dims), I and O multiples of 4.
 *   tlt_cplinear_workspace_size  bytes of ONE (B, round4(R)) fp32 matrix: `z` for the forward call (kept for the backward
 *                                call), `g` (scratch) for the backward call
 *   tlt_cplinear_fwd   y, z <- x, factors, lambda (+ bias or NULL)                                       2 launches
 *   tlt_cplinear_bwd   dx, dU_m (d_m, R), dV_m (e_m, R), dlambda (R), dbias (O) or NULL <- x, dy, z      4 launches
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t batch, rank;
  int32_t n_in, n_out;             /* tensorized modes per side, 1..4 */
  int64_t in_dims[4], out_dims[4];
} tlt_cplinear_desc;
size_t tlt_cplinear_workspace_size(const tlt_cplinear_desc* desc);
int tlt_cplinear_fwd(const tlt_cplinear_desc* desc, const float* x, const void* const* factors_in, const void* const* factors_out,
                     const float* lambda, const float* bias, float* z, float* y, void* stream);
int tlt_cplinear_bwd(const tlt_cplinear_desc* desc, const float* x, const float* dy, const float* z, const void* const* factors_in,
                     const void* const* factors_out, const float* lambda, float* g, float* dx, void* const* d_factors_in,
                     void* const* d_factors_out, float* dlambda, float* dbias, void* stream);

/* --------------------------------------------------------------------------------------------------------------
 * tlt_trl_tail_*: the Tucker tensor-regression layer behind its input projection (bf16), one op per direction
 *     v[b, r] = sum_{j, c} zp[b, j, c] core[c, j, r];      y[b, o] = sum_r v[b, r] F_o[o, r] + bias[o]
 * zp (B, J, ldz) is the projected activation, channels-last (rc valid channels per row of pitch ldz, J = product of the
 * spatial ranks); core (rc, J, ro); F_o (out, ro), ro <= 16.  Replaces: the core expansion tenalg.multi_mode_dot(core,
 * factors[n_input:]) and tenalg.inner(x, regression_weights) of tucker_trl (tltorch/functional/tensor_regression.py:42-49)
 * and their autograd.  v / dv: fp32 (B, ro), v written by the forward call and read by the backward call.
 *   tlt_trl_tail_fwd   y (B, out), v <- zp, core, F_o (+ bias or NULL)                                  1 launch
 *   tlt_trl_tail_bwd   dzp, dcore, dF_o, dbias (or NULL), all bf16 <- zp, dy, core, F_o, v; dv scratch  2 launches
 *                      (phases: 1 = dv / dzp, 2 = parameter gradients, 3 = both -- the two launches may go to different streams)
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t batch, j, rc, ldz, ro, out;
} tlt_trl_tail_desc;
int tlt_trl_tail_fwd(const tlt_trl_tail_desc* desc, const void* zp, const void* core, const void* f_out, const void* bias, void* y,
                     float* v, void* stream);
int tlt_trl_tail_bwd(const tlt_trl_tail_desc* desc, int32_t phases, const void* zp, const void* dy, const void* core, const void* f_out,
                     const float* v, float* dv, void* dzp, void* dcore, void* df_out, void* dbias, void* stream);
/* Kronecker product of two small bf16 matrices A (da, pa), B (db, pb): k[(a, b), (p, q)] = A[a, p] B[b, q], and its backward
 * (dA, dB <- dk).  Replaces: the merged spatial factor of the input projection of tucker_trl / TCL (tensor_regression.py:40,
 * tensor_contraction_layers.py:75 contract one mode at a time; the spatial modes are applied here as ONE matrix). */
int tlt_kron2_fwd(const void* a, const void* b, int32_t da, int32_t pa, int32_t db, int32_t pb, void* k, void* stream);
int tlt_kron2_bwd(const void* a, const void* b, const void* dk, int32_t da, int32_t pa, int32_t db, int32_t pb, void* d_a, void* d_b,
                  void* stream);

/* --------------------------------------------------------------------------------------------------------------
 * Op-level entry points (a whole layer per call; csrc/op_level.cu).  They sequence the same launches as the Python ops on the
 * caller's stream; every intermediate lives in caller-allocated workspaces (`ws`: written by fwd, read by bwd; `gws`: scratch of bwd).
 *
 * tlt_blocktt3_linear_*: FactorizedLinear with a 3-core BlockTT weight (bf16), dense plan: W rebuilt from the cores, one GEMM (+ bias);
 *   backward: dx, the three core gradients (through an fp32 split-K dW) and the bias gradient.  transpose = 1: y = x W^T (the layer's
 *   orientation, x has prod(cols) features), 0: y = x W.  Replaces: linear_blocktt (tltorch/functional/factorized_linear.py:39-60)
 *   and its autograd.
 * tlt_tucker_trl_*: Tucker tensor-regression layer with one output mode on (B, C, s_1..s_n) bf16 maps (n = 1..3; shapes accepted by
 *   tlt_spatial_project_supported).  factors = [F_c (C, rc), F_s1 (s_1, q_1) .. F_sn, F_o (out, ro)], core (rc, q_1..q_n, ro).
 *   Replaces: tucker_trl (tltorch/functional/tensor_regression.py:37-49) and its autograd.  side_stream (or NULL): a second stream
 *   for the parameter-gradient branch of the backward pass.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  tlt_blocktt3_desc tt;
  int64_t batch;
  int32_t transpose;
  int32_t dtype_bias;              /* tlt_dtype of bias / dbias (ignored when NULL) */
} tlt_blocktt3_linear_desc;
size_t tlt_blocktt3_linear_workspace_size(const tlt_blocktt3_linear_desc* desc);
size_t tlt_blocktt3_linear_grad_workspace_size(const tlt_blocktt3_linear_desc* desc);
int tlt_blocktt3_linear_fwd(const tlt_blocktt3_linear_desc* desc, const void* x, const void* c1, const void* c2, const void* c3,
                            const void* bias, void* ws, void* y, void* stream);
int tlt_blocktt3_linear_bwd(const tlt_blocktt3_linear_desc* desc, const void* x, const void* dy, const void* ws, const void* c1,
                            const void* c2, const void* c3, void* gws, void* dx, void* d1, void* d2, void* d3, void* dbias, void* stream);

typedef struct {
  int64_t batch, channels;
  int32_t n_spatial;
  int64_t spatial[3], spatial_rank[3];
  int64_t rc, ro, out;             /* channel rank, output rank (<= 16), output features */
} tlt_tucker_trl_desc;
size_t tlt_tucker_trl_workspace_size(const tlt_tucker_trl_desc* desc);
size_t tlt_tucker_trl_grad_workspace_size(const tlt_tucker_trl_desc* desc);
int tlt_tucker_trl_fwd(const tlt_tucker_trl_desc* desc, const void* x, const void* core, const void* const* factors, const void* bias,
                       void* ws, void* y, void* stream);
int tlt_tucker_trl_bwd(const tlt_tucker_trl_desc* desc, const void* x, const void* dy, const void* core, const void* const* factors,
                       const void* ws, void* gws, void* dx, void* dcore, void* const* d_factors, void* dbias, void* stream,
                       void* side_stream);

#ifdef __cplusplus
}
#endif
#endif /* TLTORCH_CUDA_H_ */
