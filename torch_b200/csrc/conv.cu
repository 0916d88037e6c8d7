// Convolution building blocks: unfold (im2col / col2im) for dense convs and direct depthwise kernels.
// Channels-first layout (B, C, S1[, S2[, S3]]), 1..3 spatial dims, arbitrary stride / padding / dilation.
#include "common.cuh"

namespace tlt {

// element-wise fallbacks of the depthwise entry points (dwconv.cu dispatches to them when a plane does not fit smem)
int dwconv_fwd_generic(const tlt_conv_desc* d, const void* x, const void* w, void* y, void* stream);
int dwconv_dgrad_generic(const tlt_conv_desc* d, const void* dy, const void* w, void* dx, void* stream);
int dwconv_wgrad_generic(const tlt_conv_desc* d, const void* x, const void* dy, float* dw, void* stream);
int im2col_generic(const tlt_conv_desc* d, const void* x, void* cols, void* stream);
int col2im_generic(const tlt_conv_desc* d, const void* dcols, void* dx, void* stream);

struct ConvParams {
  int ndim;
  long long batch, channels;
  int in[3], out[3], ker[3], stride[3], pad[3], dil[3];
  long long in_sp, out_sp;   // prod(in), prod(out)
  int taps;                  // prod(ker)
};

static int fill_conv(const tlt_conv_desc* d, ConvParams* p) {
  TLT_REQUIRE(d != nullptr, "conv: null descriptor");
  TLT_REQUIRE(d->ndim >= 1 && d->ndim <= 3, "conv: ndim must be 1..3 (got %d)", d->ndim);
  p->ndim = d->ndim;
  p->batch = d->batch;
  p->channels = d->channels;
  p->in_sp = 1; p->out_sp = 1; p->taps = 1;
  // normalise to 3 dims by padding LEADING unit dims
  int off = 3 - d->ndim;
  for (int i = 0; i < 3; ++i) { p->in[i] = 1; p->out[i] = 1; p->ker[i] = 1; p->stride[i] = 1; p->pad[i] = 0; p->dil[i] = 1; }
  for (int i = 0; i < d->ndim; ++i) {
    TLT_REQUIRE(d->kernel[i] >= 1 && d->stride[i] >= 1 && d->dilation[i] >= 1 && d->padding[i] >= 0,
                "conv: bad kernel/stride/dilation/padding on dim %d", i);
    long long expect = ((long long)d->in_size[i] + 2LL * d->padding[i] - (long long)d->dilation[i] * (d->kernel[i] - 1) - 1) / d->stride[i] + 1;
    TLT_REQUIRE(expect == d->out_size[i], "conv: out_size[%d]=%lld inconsistent (expected %lld)", i,
                (long long)d->out_size[i], expect);
    p->in[off + i] = (int)d->in_size[i]; p->out[off + i] = (int)d->out_size[i]; p->ker[off + i] = d->kernel[i];
    p->stride[off + i] = d->stride[i]; p->pad[off + i] = d->padding[i]; p->dil[off + i] = d->dilation[i];
    p->in_sp *= d->in_size[i]; p->out_sp *= d->out_size[i]; p->taps *= d->kernel[i];
  }
  return TLT_OK;
}

// cols[(b, o), (c, t)] : one thread per cols element, o fastest within a (b, c, t) group would be coalesced on the
// READ side; we pick the WRITE side (t fastest, then c) because cols is the larger tensor.
template <typename T>
__global__ void im2col_kernel(const ConvParams p, const T* __restrict__ x, T* __restrict__ cols) {
  const long long kdim = p.channels * p.taps;
  const long long total = p.batch * p.out_sp * kdim;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long kk = i % kdim;
    long long row = i / kdim;
    int t = (int)(kk % p.taps);
    long long c = kk / p.taps;
    long long o = row % p.out_sp;
    long long b = row / p.out_sp;
    int t2 = t % p.ker[2], t1 = (t / p.ker[2]) % p.ker[1], t0 = t / (p.ker[2] * p.ker[1]);
    int o2 = (int)(o % p.out[2]), o1 = (int)((o / p.out[2]) % p.out[1]), o0 = (int)(o / ((long long)p.out[2] * p.out[1]));
    int i0 = o0 * p.stride[0] - p.pad[0] + t0 * p.dil[0];
    int i1 = o1 * p.stride[1] - p.pad[1] + t1 * p.dil[1];
    int i2 = o2 * p.stride[2] - p.pad[2] + t2 * p.dil[2];
    T v = from_f32<T>(0.f);
    if (i0 >= 0 && i0 < p.in[0] && i1 >= 0 && i1 < p.in[1] && i2 >= 0 && i2 < p.in[2])
      v = x[((b * p.channels + c) * p.in[0] + i0) * (long long)p.in[1] * p.in[2] + (long long)i1 * p.in[2] + i2];
    cols[i] = v;
  }
}

// dx[b,c,i] = sum_{t : (i + pad - t*dil) % stride == 0, o in range} dcols[(b,o),(c,t)]
template <typename T>
__global__ void col2im_kernel(const ConvParams p, const T* __restrict__ dcols, T* __restrict__ dx) {
  const long long kdim = p.channels * p.taps;
  const long long total = p.batch * p.channels * p.in_sp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long sp = i % p.in_sp;
    long long bc = i / p.in_sp;
    long long c = bc % p.channels, b = bc / p.channels;
    int i2 = (int)(sp % p.in[2]), i1 = (int)((sp / p.in[2]) % p.in[1]), i0 = (int)(sp / ((long long)p.in[2] * p.in[1]));
    float acc = 0.f;
    for (int t0 = 0; t0 < p.ker[0]; ++t0) {
      int n0 = i0 + p.pad[0] - t0 * p.dil[0];
      if (n0 < 0 || n0 % p.stride[0]) continue;
      int o0 = n0 / p.stride[0];
      if (o0 >= p.out[0]) continue;
      for (int t1 = 0; t1 < p.ker[1]; ++t1) {
        int n1 = i1 + p.pad[1] - t1 * p.dil[1];
        if (n1 < 0 || n1 % p.stride[1]) continue;
        int o1 = n1 / p.stride[1];
        if (o1 >= p.out[1]) continue;
        for (int t2 = 0; t2 < p.ker[2]; ++t2) {
          int n2 = i2 + p.pad[2] - t2 * p.dil[2];
          if (n2 < 0 || n2 % p.stride[2]) continue;
          int o2 = n2 / p.stride[2];
          if (o2 >= p.out[2]) continue;
          long long o = ((long long)o0 * p.out[1] + o1) * p.out[2] + o2;
          int t = (t0 * p.ker[1] + t1) * p.ker[2] + t2;
          acc += to_f32<T>(dcols[(b * p.out_sp + o) * kdim + c * p.taps + t]);
        }
      }
    }
    dx[i] = from_f32<T>(acc);
  }
}

// depthwise forward: one thread per output element (innermost spatial dim fastest -> coalesced reads and writes)
template <typename T>
__global__ void dwconv_fwd_kernel(const ConvParams p, const T* __restrict__ x, const T* __restrict__ w, T* __restrict__ y) {
  const long long total = p.batch * p.channels * p.out_sp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long o = i % p.out_sp;
    long long bc = i / p.out_sp;
    long long c = bc % p.channels;
    int o2 = (int)(o % p.out[2]), o1 = (int)((o / p.out[2]) % p.out[1]), o0 = (int)(o / ((long long)p.out[2] * p.out[1]));
    const T* xb = x + bc * p.in_sp;
    const T* wc = w + c * p.taps;
    float acc = 0.f;
    for (int t0 = 0; t0 < p.ker[0]; ++t0) {
      int i0 = o0 * p.stride[0] - p.pad[0] + t0 * p.dil[0];
      if (i0 < 0 || i0 >= p.in[0]) continue;
      for (int t1 = 0; t1 < p.ker[1]; ++t1) {
        int i1 = o1 * p.stride[1] - p.pad[1] + t1 * p.dil[1];
        if (i1 < 0 || i1 >= p.in[1]) continue;
        for (int t2 = 0; t2 < p.ker[2]; ++t2) {
          int i2 = o2 * p.stride[2] - p.pad[2] + t2 * p.dil[2];
          if (i2 < 0 || i2 >= p.in[2]) continue;
          acc = fmaf(to_f32<T>(xb[((long long)i0 * p.in[1] + i1) * p.in[2] + i2]),
                     to_f32<T>(wc[(t0 * p.ker[1] + t1) * p.ker[2] + t2]), acc);
        }
      }
    }
    y[i] = from_f32<T>(acc);
  }
}

template <typename T>
__global__ void dwconv_dgrad_kernel(const ConvParams p, const T* __restrict__ dy, const T* __restrict__ w, T* __restrict__ dx) {
  const long long total = p.batch * p.channels * p.in_sp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long sp = i % p.in_sp;
    long long bc = i / p.in_sp;
    long long c = bc % p.channels;
    int i2 = (int)(sp % p.in[2]), i1 = (int)((sp / p.in[2]) % p.in[1]), i0 = (int)(sp / ((long long)p.in[2] * p.in[1]));
    const T* dyb = dy + bc * p.out_sp;
    const T* wc = w + c * p.taps;
    float acc = 0.f;
    for (int t0 = 0; t0 < p.ker[0]; ++t0) {
      int n0 = i0 + p.pad[0] - t0 * p.dil[0];
      if (n0 < 0 || n0 % p.stride[0]) continue;
      int o0 = n0 / p.stride[0];
      if (o0 >= p.out[0]) continue;
      for (int t1 = 0; t1 < p.ker[1]; ++t1) {
        int n1 = i1 + p.pad[1] - t1 * p.dil[1];
        if (n1 < 0 || n1 % p.stride[1]) continue;
        int o1 = n1 / p.stride[1];
        if (o1 >= p.out[1]) continue;
        for (int t2 = 0; t2 < p.ker[2]; ++t2) {
          int n2 = i2 + p.pad[2] - t2 * p.dil[2];
          if (n2 < 0 || n2 % p.stride[2]) continue;
          int o2 = n2 / p.stride[2];
          if (o2 >= p.out[2]) continue;
          acc = fmaf(to_f32<T>(dyb[((long long)o0 * p.out[1] + o1) * p.out[2] + o2]),
                     to_f32<T>(wc[(t0 * p.ker[1] + t1) * p.ker[2] + t2]), acc);
        }
      }
    }
    dx[i] = from_f32<T>(acc);
  }
}

// dw[c,t] += sum_{b,o} dy[b,c,o] x[b,c,o*s-p+t*d].  grid = (channels, batch chunks); block reduces then one atomic
// per (c,t) per block.  taps <= 32 handled per block in registers (max 27 for 3x3x3); larger kernels loop.
template <typename T>
__global__ void dwconv_wgrad_kernel(const ConvParams p, const T* __restrict__ x, const T* __restrict__ dy,
                                    float* __restrict__ dw, int batch_per_block) {
  const long long c = blockIdx.x;
  const long long b_begin = (long long)blockIdx.y * batch_per_block;
  const long long b_end = min((long long)p.batch, b_begin + batch_per_block);
  __shared__ float red[32];
  for (int t = 0; t < p.taps; ++t) {
    int t2 = t % p.ker[2], t1 = (t / p.ker[2]) % p.ker[1], t0 = t / (p.ker[2] * p.ker[1]);
    float acc = 0.f;
    for (long long b = b_begin; b < b_end; ++b) {
      const T* xb = x + (b * p.channels + c) * p.in_sp;
      const T* dyb = dy + (b * p.channels + c) * p.out_sp;
      for (long long o = threadIdx.x; o < p.out_sp; o += blockDim.x) {
        int o2 = (int)(o % p.out[2]), o1 = (int)((o / p.out[2]) % p.out[1]), o0 = (int)(o / ((long long)p.out[2] * p.out[1]));
        int i0 = o0 * p.stride[0] - p.pad[0] + t0 * p.dil[0];
        int i1 = o1 * p.stride[1] - p.pad[1] + t1 * p.dil[1];
        int i2 = o2 * p.stride[2] - p.pad[2] + t2 * p.dil[2];
        if (i0 < 0 || i0 >= p.in[0] || i1 < 0 || i1 >= p.in[1] || i2 < 0 || i2 >= p.in[2]) continue;
        acc = fmaf(to_f32<T>(dyb[o]), to_f32<T>(xb[((long long)i0 * p.in[1] + i1) * p.in[2] + i2]), acc);
      }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
      float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
      if (threadIdx.x == 0) atomicAdd(dw + c * p.taps + t, v);
    }
    __syncthreads();
  }
}

template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ src, TD* __restrict__ dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = from_f32<TD>(to_f32<TS>(src[i]));
}

__global__ void l2_flush_kernel(float4* __restrict__ buf, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    buf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

static inline int grid_for(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  long long cap = 148LL * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace tlt

using namespace tlt;

#define DISPATCH_DT(dt, CALL_F32, CALL_BF16)            \
  do {                                                  \
    if ((dt) == TLT_F32) { CALL_F32; }                  \
    else if ((dt) == TLT_BF16) { CALL_BF16; }           \
    else { set_error("unknown dtype %d", (int)(dt)); return TLT_ERR_INVALID; } \
  } while (0)

int tlt::im2col_generic(const tlt_conv_desc* d, const void* x, void* cols, void* stream) {
  ConvParams p; int rc = fill_conv(d, &p); if (rc) return rc;
  long long total = p.batch * p.out_sp * p.channels * p.taps;
  if (total == 0) return TLT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_DT(d->dtype, (im2col_kernel<float><<<grid_for(total, 256), 256, 0, st>>>(p, (const float*)x, (float*)cols)),
              (im2col_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, st>>>(p, (const __nv_bfloat16*)x, (__nv_bfloat16*)cols)));
  return check_launch("im2col_kernel");
}

int tlt::col2im_generic(const tlt_conv_desc* d, const void* dcols, void* dx, void* stream) {
  ConvParams p; int rc = fill_conv(d, &p); if (rc) return rc;
  long long total = p.batch * p.channels * p.in_sp;
  if (total == 0) return TLT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_DT(d->dtype, (col2im_kernel<float><<<grid_for(total, 256), 256, 0, st>>>(p, (const float*)dcols, (float*)dx)),
              (col2im_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, st>>>(p, (const __nv_bfloat16*)dcols, (__nv_bfloat16*)dx)));
  return check_launch("col2im_kernel");
}

int tlt::dwconv_fwd_generic(const tlt_conv_desc* d, const void* x, const void* w, void* y, void* stream) {
  ConvParams p; int rc = fill_conv(d, &p); if (rc) return rc;
  long long total = p.batch * p.channels * p.out_sp;
  if (total == 0) return TLT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_DT(d->dtype, (dwconv_fwd_kernel<float><<<grid_for(total, 256), 256, 0, st>>>(p, (const float*)x, (const float*)w, (float*)y)),
              (dwconv_fwd_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, st>>>(p, (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, (__nv_bfloat16*)y)));
  return check_launch("dwconv_fwd_kernel");
}

int tlt::dwconv_dgrad_generic(const tlt_conv_desc* d, const void* dy, const void* w, void* dx, void* stream) {
  ConvParams p; int rc = fill_conv(d, &p); if (rc) return rc;
  long long total = p.batch * p.channels * p.in_sp;
  if (total == 0) return TLT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_DT(d->dtype, (dwconv_dgrad_kernel<float><<<grid_for(total, 256), 256, 0, st>>>(p, (const float*)dy, (const float*)w, (float*)dx)),
              (dwconv_dgrad_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, st>>>(p, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)w, (__nv_bfloat16*)dx)));
  return check_launch("dwconv_dgrad_kernel");
}

int tlt::dwconv_wgrad_generic(const tlt_conv_desc* d, const void* x, const void* dy, float* dw, void* stream) {
  ConvParams p; int rc = fill_conv(d, &p); if (rc) return rc;
  if (p.batch == 0 || p.channels == 0 || p.out_sp == 0) return TLT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // enough blocks to fill the machine: channels x batch-chunks ~ 4 waves of 148 SMs
  long long want_chunks = (148LL * 4 + p.channels - 1) / p.channels;
  if (want_chunks > p.batch) want_chunks = p.batch;
  if (want_chunks < 1) want_chunks = 1;
  if (want_chunks > 65535) want_chunks = 65535;
  int bpb = (int)((p.batch + want_chunks - 1) / want_chunks);
  int chunks = (int)((p.batch + bpb - 1) / bpb);
  TLT_REQUIRE(p.channels <= 0x7fffffffLL, "dwconv_wgrad: too many channels");
  dim3 grid((unsigned)p.channels, (unsigned)chunks);
  int threads = p.out_sp >= 256 ? 256 : (p.out_sp >= 128 ? 128 : 64);
  DISPATCH_DT(d->dtype, (dwconv_wgrad_kernel<float><<<grid, threads, 0, st>>>(p, (const float*)x, (const float*)dy, dw, bpb)),
              (dwconv_wgrad_kernel<__nv_bfloat16><<<grid, threads, 0, st>>>(p, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, dw, bpb)));
  return check_launch("dwconv_wgrad_kernel");
}

extern "C" int tlt_cast(const void* src, int32_t ds, void* dst, int32_t dd, int64_t n, void* stream) {
  if (n <= 0) return TLT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int g = grid_for(n, 256);
  typedef __nv_bfloat16 bf;
  if (ds == TLT_F32 && dd == TLT_BF16) cast_kernel<float, bf><<<g, 256, 0, st>>>((const float*)src, (bf*)dst, n);
  else if (ds == TLT_BF16 && dd == TLT_F32) cast_kernel<bf, float><<<g, 256, 0, st>>>((const bf*)src, (float*)dst, n);
  else if (ds == TLT_F32 && dd == TLT_F32) cast_kernel<float, float><<<g, 256, 0, st>>>((const float*)src, (float*)dst, n);
  else if (ds == TLT_BF16 && dd == TLT_BF16) cast_kernel<bf, bf><<<g, 256, 0, st>>>((const bf*)src, (bf*)dst, n);
  else { set_error("tlt_cast: unknown dtype"); return TLT_ERR_INVALID; }
  return check_launch("cast_kernel");
}

extern "C" int tlt_l2_flush(void* scratch, size_t bytes, void* stream) {
  if (!scratch || bytes < 16) return TLT_OK;
  l2_flush_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>((float4*)scratch, (long long)(bytes / 16));
  return check_launch("l2_flush_kernel");
}
