// Channels-last ("rows") staging for dense k x k convolutions on the tcgen05 GEMM.
//
// A dense conv whose channel counts are in the hundreds (the Tucker core conv 388 -> 388 of BASELINE config 4, TT cores,
// reconstructed kernels) is a GEMM over (tap, channel); what it needs from the activation is rows of contiguous channels.
// The first version unfolded in NCHW (cols[(b,s),(c,t)]: 2-byte gathers, taps innermost) and then re-packed for TMA
// (pack_kernel) -- 233 us + 122 us forward and an 828 us col2im backward at config 4.  Here the activation is transposed
// ONCE to (batch * spatial, channels padded to 8) rows; unfold / fold then move whole 16-byte channel chunks (coalesced on
// both sides, no division per element) and every 1x1 stage around the conv is a plain K-major GEMM on the same rows.
//
//   tlt_nchw_to_rows : rows[(b, s), c] = x[b, c, s]           (c >= C zero-filled up to ld)
//   tlt_rows_to_nchw : y[b, c, s] = rows[(b, s), c]
//   tlt_im2col_rows  : cols[(b, o...), (t..., c)] = z[(b, o*stride - pad + t*dil), c]   (zero outside)
//   tlt_col2im_rows  : dz[(b, i...), c] = sum over (o, t) hitting i of dcols[(b, o), (t, c)]      (gather form, fp32 sum)
//
// Reference call sites: F.conv{1,2,3}d with a dense kernel in tltorch/functional/convolution.py (:161 Tucker core conv,
// :213-216 TT cores, :16-53 reconstructed weights) and their autograd.
#include "common.cuh"

namespace tlt {

// General tiles: block (32, 8) moves 64 channels x 32 positions of one image through shared memory.
__global__ void __launch_bounds__(256) nchw_to_rows_kernel(const __nv_bfloat16* __restrict__ x, long long C, long long S, long long ld,
                                                           __nv_bfloat16* __restrict__ rows) {
  __shared__ unsigned short tile[64][34];
  const long long b = blockIdx.z;
  const long long c0 = (long long)blockIdx.y * 64, s0 = (long long)blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const unsigned short* xs = reinterpret_cast<const unsigned short*>(x);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = ty + 8 * i;
    unsigned short v = 0;
    if (c0 + c < C && s0 + tx < S) v = xs[(b * C + c0 + c) * S + s0 + tx];
    tile[c][tx] = v;
  }
  __syncthreads();
  // store: thread tx owns channels (2 tx, 2 tx + 1); 4-byte stores, 128 bytes per warp
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int s = ty + 8 * i;
    const long long c = c0 + 2 * tx;
    if (s0 + s < S && c < ld) {
      const uint32_t v = (uint32_t)tile[2 * tx][s] | ((uint32_t)tile[2 * tx + 1][s] << 16);
      *reinterpret_cast<uint32_t*>(rows + (b * S + s0 + s) * ld + c) = v;
    }
  }
}

__global__ void __launch_bounds__(256) rows_to_nchw_kernel(const __nv_bfloat16* __restrict__ rows, long long C, long long S, long long ld,
                                                           __nv_bfloat16* __restrict__ y) {
  __shared__ unsigned short tile[64][34];
  const long long b = blockIdx.z;
  const long long c0 = (long long)blockIdx.y * 64, s0 = (long long)blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int s = ty + 8 * i;
    const long long c = c0 + 2 * tx;
    uint32_t v = 0;
    if (s0 + s < S && c < ld) v = *reinterpret_cast<const uint32_t*>(rows + (b * S + s0 + s) * ld + c);
    tile[2 * tx][s] = (unsigned short)(v & 0xffffu);
    tile[2 * tx + 1][s] = (unsigned short)(v >> 16);
  }
  __syncthreads();
  unsigned short* ys = reinterpret_cast<unsigned short*>(y);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = ty + 8 * i;
    if (c0 + c < C && s0 + tx < S) ys[(b * C + c0 + c) * S + s0 + tx] = tile[c][tx];
  }
}

// Small feature maps (S <= 256, C % 64 == 0, 64 * S * 2 bytes a multiple of 16): the 64 channels x S positions of one
// image are ONE contiguous span of NCHW memory.  It moves as 16-byte chunks on the NCHW side and as 128-byte row segments
// (a warp = 64 channels of one position, 4 bytes per lane) on the rows side; shared memory holds the span linearly.
constexpr int kSpanMaxS = 256;
__global__ void __launch_bounds__(256) nchw_to_rows_span_kernel(const __nv_bfloat16* __restrict__ x, int C, int S, long long ld,
                                                                __nv_bfloat16* __restrict__ rows, long long n_units) {
  extern __shared__ __align__(16) unsigned char span_smem[];
  unsigned short* sp = reinterpret_cast<unsigned short*>(span_smem);           // [64][S]
  const int cblocks = C >> 6, n16 = (64 * S * 2) >> 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const long long b = unit / cblocks;
    const int c0 = (int)(unit - b * cblocks) << 6;
    const uint4* src = reinterpret_cast<const uint4*>(x + (b * C + c0) * (long long)S);
    __syncthreads();
    for (int v = threadIdx.x; v < n16; v += 256) reinterpret_cast<uint4*>(sp)[v] = src[v];
    __syncthreads();
    for (int s = warp; s < S; s += 8) {
      const uint32_t v = (uint32_t)sp[(2 * lane) * S + s] | ((uint32_t)sp[(2 * lane + 1) * S + s] << 16);
      *reinterpret_cast<uint32_t*>(rows + (b * S + s) * ld + c0 + 2 * lane) = v;
    }
  }
}

__global__ void __launch_bounds__(256) rows_to_nchw_span_kernel(const __nv_bfloat16* __restrict__ rows, int C, int S, long long ld,
                                                                __nv_bfloat16* __restrict__ y, long long n_units) {
  extern __shared__ __align__(16) unsigned char span_smem[];
  unsigned short* sp = reinterpret_cast<unsigned short*>(span_smem);           // [64][S]
  const int cblocks = C >> 6, n16 = (64 * S * 2) >> 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const long long b = unit / cblocks;
    const int c0 = (int)(unit - b * cblocks) << 6;
    __syncthreads();
    for (int s = warp; s < S; s += 8) {
      const uint32_t v = *reinterpret_cast<const uint32_t*>(rows + (b * S + s) * ld + c0 + 2 * lane);
      sp[(2 * lane) * S + s] = (unsigned short)(v & 0xffffu);
      sp[(2 * lane + 1) * S + s] = (unsigned short)(v >> 16);
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(y + (b * C + c0) * (long long)S);
    for (int v = threadIdx.x; v < n16; v += 256) dst[v] = reinterpret_cast<const uint4*>(sp)[v];
  }
}

static bool span_ok(long long C, long long S, long long ld, const void* nchw) {
  return S <= kSpanMaxS && C % 64 == 0 && ld == C && (64 * S * 2) % 16 == 0 && ((uintptr_t)nchw % 16) == 0;
}

struct RowsGeom {
  int ndim;
  long long batch, ld;                       // ld = channels per row (multiple of 8)
  int in_size[3], out_size[3], kernel[3], stride[3], padding[3], dilation[3];
  long long in_sp, out_sp;
  int taps;
};

// one thread per (output row, 16-byte chunk): the row's position is decoded once, then every tap is one add + one copy
__global__ void __launch_bounds__(256) im2col_rows_kernel(const RowsGeom g, const uint4* __restrict__ z, uint4* __restrict__ cols) {
  const int cpr = (int)(g.ld >> 3);                                  // chunks per (row, tap)
  const long long total = g.batch * g.out_sp * cpr;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(e % cpr);
    const long long row = e / cpr;
    int o = (int)(row % g.out_sp);
    const long long b = row / g.out_sp;
    int base[3] = {0, 0, 0};                                         // input coordinate of tap 0 per dim
    for (int d = g.ndim - 1; d >= 0; --d) {
      const int od = o % g.out_size[d];
      o /= g.out_size[d];
      base[d] = od * g.stride[d] - g.padding[d];
    }
    const uint4* zb = z + b * g.in_sp * cpr + ch;
    uint4* cb = cols + row * g.taps * cpr + ch;
    const int k1 = g.ndim > 1 ? g.kernel[g.ndim - 1] : 1, k2 = g.ndim > 2 ? g.kernel[g.ndim - 2] : 1;
    for (int tap = 0; tap < g.taps; ++tap) {
      int t[3];
      if (g.ndim == 1) { t[0] = tap; }
      else if (g.ndim == 2) { t[0] = tap / k1; t[1] = tap - t[0] * k1; }
      else { t[2] = tap % k1; const int q = tap / k1; t[1] = q % k2; t[0] = q / k2; }
      int off = 0;
      bool ok = true;
      for (int d = 0; d < g.ndim; ++d) {
        const int i = base[d] + t[d] * g.dilation[d];
        ok = ok && i >= 0 && i < g.in_size[d];
        off = off * g.in_size[d] + i;
      }
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (ok) v = zb[(long long)off * cpr];
      cb[(long long)tap * cpr] = v;
    }
  }
}

__device__ __forceinline__ void acc_bf16x2(uint32_t w, float& a, float& b) {
  const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
  a += f.x;
  b += f.y;
}

// one thread per 16-byte chunk of dz: (b, position, chunk); gathers every (output position, tap) that reads this input
__global__ void __launch_bounds__(256) col2im_rows_kernel(const RowsGeom g, const uint4* __restrict__ dcols, uint4* __restrict__ dz) {
  const long long cpr = g.ld >> 3;
  const long long total = g.batch * g.in_sp * cpr;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long ch = e % cpr;
    long long pos = e / cpr;
    long long ip = pos % g.in_sp;
    const long long b = pos / g.in_sp;
    int idx[3] = {0, 0, 0};
    for (int d = g.ndim - 1; d >= 0; --d) {
      idx[d] = (int)(ip % g.in_size[d]);
      ip /= g.in_size[d];
    }
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int k1 = g.ndim > 1 ? g.kernel[g.ndim - 1] : 1, k2 = g.ndim > 2 ? g.kernel[g.ndim - 2] : 1;
    for (int tap = 0; tap < g.taps; ++tap) {
      int t[3];
      if (g.ndim == 1) { t[0] = tap; }
      else if (g.ndim == 2) { t[0] = tap / k1; t[1] = tap - t[0] * k1; }
      else { t[2] = tap % k1; const int q = tap / k1; t[1] = q % k2; t[0] = q / k2; }
      int o_off = 0;
      bool ok = true;
      for (int d = 0; d < g.ndim; ++d) {
        const int num = idx[d] + g.padding[d] - t[d] * g.dilation[d];
        const int od = g.stride[d] == 1 ? num : num / g.stride[d];
        ok = ok && num >= 0 && od * g.stride[d] == num && od < g.out_size[d];
        o_off = o_off * g.out_size[d] + od;
      }
      if (!ok) continue;
      const uint4 v = dcols[((b * g.out_sp + o_off) * g.taps + tap) * cpr + ch];
      acc_bf16x2(v.x, acc[0], acc[1]);
      acc_bf16x2(v.y, acc[2], acc[3]);
      acc_bf16x2(v.z, acc[4], acc[5]);
      acc_bf16x2(v.w, acc[6], acc[7]);
    }
    uint4 out;
    __nv_bfloat162 p;
    p = __floats2bfloat162_rn(acc[0], acc[1]); out.x = *reinterpret_cast<uint32_t*>(&p);
    p = __floats2bfloat162_rn(acc[2], acc[3]); out.y = *reinterpret_cast<uint32_t*>(&p);
    p = __floats2bfloat162_rn(acc[4], acc[5]); out.z = *reinterpret_cast<uint32_t*>(&p);
    p = __floats2bfloat162_rn(acc[6], acc[7]); out.w = *reinterpret_cast<uint32_t*>(&p);
    dz[e] = out;
  }
}

// 2-D specialisations with compile-time taps (the 3 x 3 core of the Tucker / TT convolutions): the position is decoded once, the tap
// loops are unrolled and ALL tap loads of a thread are issued before the first store / add.  The generic kernels above decode the tap
// with divisions inside a runtime loop and keep one load in flight per thread: 63 us (im2col) / 83 us (col2im) for a 177 MB cols
// tensor at BASELINE config 4 -- latency-bound at 2.1 - 3.1 TB/s.
template <int KH, int KW>
__global__ void __launch_bounds__(256) im2col_rows2d_kernel(const RowsGeom g, const uint4* __restrict__ z, uint4* __restrict__ cols) {
  const int cpr = (int)(g.ld >> 3);
  const int OW = g.out_size[1], IH = g.in_size[0], IW = g.in_size[1];
  const long long total = g.batch * g.out_sp * cpr;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(e % cpr);
    const long long row = e / cpr;
    const int o = (int)(row % g.out_sp);
    const long long b = row / g.out_sp;
    const int oy = o / OW, ox = o - oy * OW;
    const int by = oy * g.stride[0] - g.padding[0], bx = ox * g.stride[1] - g.padding[1];
    const uint4* zb = z + b * g.in_sp * cpr + ch;
    uint4* cb = cols + row * (KH * KW) * cpr + ch;
    uint4 v[KH * KW];
#pragma unroll
    for (int kh = 0; kh < KH; ++kh) {
      const int iy = by + kh * g.dilation[0];
      const bool oky = iy >= 0 && iy < IH;
#pragma unroll
      for (int kw = 0; kw < KW; ++kw) {
        const int ix = bx + kw * g.dilation[1];
        const bool ok = oky && ix >= 0 && ix < IW;
        v[kh * KW + kw] = ok ? zb[(long long)(iy * IW + ix) * cpr] : make_uint4(0u, 0u, 0u, 0u);
      }
    }
#pragma unroll
    for (int t = 0; t < KH * KW; ++t) cb[(long long)t * cpr] = v[t];
  }
}

template <int KH, int KW>
__global__ void __launch_bounds__(256) col2im_rows2d_kernel(const RowsGeom g, const uint4* __restrict__ dcols, uint4* __restrict__ dz) {
  const long long cpr = g.ld >> 3;
  const int OH = g.out_size[0], OW = g.out_size[1], IW = g.in_size[1];
  const long long total = g.batch * g.in_sp * cpr;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long ch = e % cpr;
    const long long pos = e / cpr;
    const int ip = (int)(pos % g.in_sp);
    const long long b = pos / g.in_sp;
    const int iy = ip / IW, ix = ip - iy * IW;
    uint4 v[KH * KW];
#pragma unroll
    for (int kh = 0; kh < KH; ++kh) {
      const int ny = iy + g.padding[0] - kh * g.dilation[0];
      const int oy = g.stride[0] == 1 ? ny : ny / g.stride[0];
      const bool oky = ny >= 0 && oy * g.stride[0] == ny && oy < OH;
#pragma unroll
      for (int kw = 0; kw < KW; ++kw) {
        const int nx = ix + g.padding[1] - kw * g.dilation[1];
        const int ox = g.stride[1] == 1 ? nx : nx / g.stride[1];
        const bool ok = oky && nx >= 0 && ox * g.stride[1] == nx && ox < OW;
        v[kh * KW + kw] = ok ? dcols[((b * g.out_sp + (long long)(oy * OW + ox)) * (KH * KW) + kh * KW + kw) * cpr + ch] : make_uint4(0u, 0u, 0u, 0u);
      }
    }
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < KH * KW; ++t) {
      acc_bf16x2(v[t].x, acc[0], acc[1]);
      acc_bf16x2(v[t].y, acc[2], acc[3]);
      acc_bf16x2(v[t].z, acc[4], acc[5]);
      acc_bf16x2(v[t].w, acc[6], acc[7]);
    }
    uint4 out;
    __nv_bfloat162 p;
    p = __floats2bfloat162_rn(acc[0], acc[1]); out.x = *reinterpret_cast<uint32_t*>(&p);
    p = __floats2bfloat162_rn(acc[2], acc[3]); out.y = *reinterpret_cast<uint32_t*>(&p);
    p = __floats2bfloat162_rn(acc[4], acc[5]); out.z = *reinterpret_cast<uint32_t*>(&p);
    p = __floats2bfloat162_rn(acc[6], acc[7]); out.w = *reinterpret_cast<uint32_t*>(&p);
    dz[e] = out;
  }
}

static int fill_rows(const tlt_conv_desc* d, RowsGeom* g) {
  TLT_REQUIRE(d && d->dtype == TLT_BF16, "rows unfold: bf16 only");
  TLT_REQUIRE(d->ndim >= 1 && d->ndim <= 3, "rows unfold: 1..3 spatial dims");
  TLT_REQUIRE(d->batch >= 1 && d->channels >= 8 && d->channels % 8 == 0, "rows unfold: channels per row must be a multiple of 8");
  g->ndim = d->ndim; g->batch = d->batch; g->ld = d->channels;
  g->in_sp = 1; g->out_sp = 1; g->taps = 1;
  for (int i = 0; i < d->ndim; ++i) {
    TLT_REQUIRE(d->in_size[i] >= 1 && d->out_size[i] >= 1 && d->kernel[i] >= 1 && d->stride[i] >= 1 && d->dilation[i] >= 1 && d->padding[i] >= 0,
                "rows unfold: bad geometry");
    g->in_size[i] = (int)d->in_size[i]; g->out_size[i] = (int)d->out_size[i];
    g->kernel[i] = d->kernel[i]; g->stride[i] = d->stride[i]; g->padding[i] = d->padding[i]; g->dilation[i] = d->dilation[i];
    g->in_sp *= d->in_size[i]; g->out_sp *= d->out_size[i]; g->taps *= d->kernel[i];
  }
  return TLT_OK;
}

static unsigned grid_for(long long total) {
  long long blocks = (total + 255) / 256;
  const long long cap = 148LL * 16;
  return (unsigned)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace tlt

using namespace tlt;

extern "C" int tlt_nchw_to_rows(const void* x, int64_t batch, int64_t channels, int64_t spatial, void* rows, int64_t ld, void* stream) {
  TLT_REQUIRE(x && rows && batch >= 1 && channels >= 1 && spatial >= 1, "tlt_nchw_to_rows: bad arguments");
  TLT_REQUIRE(ld >= channels && ld % 2 == 0 && batch <= 65535, "tlt_nchw_to_rows: ld must be even and >= channels; batch <= 65535");
  if (span_ok(channels, spatial, ld, x)) {
    const long long units = batch * (channels >> 6);
    const unsigned blocks = (unsigned)(units < 148LL * 8 ? units : 148LL * 8);
    nchw_to_rows_span_kernel<<<blocks, 256, (size_t)64 * spatial * 2, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (int)channels, (int)spatial, ld,
                                                                                            (__nv_bfloat16*)rows, units);
    return check_launch("nchw_to_rows_span_kernel");
  }
  dim3 grid((unsigned)((spatial + 31) / 32), (unsigned)((ld + 63) / 64), (unsigned)batch);
  nchw_to_rows_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, channels, spatial, ld, (__nv_bfloat16*)rows);
  return check_launch("nchw_to_rows_kernel");
}

extern "C" int tlt_rows_to_nchw(const void* rows, int64_t batch, int64_t channels, int64_t spatial, int64_t ld, void* y, void* stream) {
  TLT_REQUIRE(y && rows && batch >= 1 && channels >= 1 && spatial >= 1, "tlt_rows_to_nchw: bad arguments");
  TLT_REQUIRE(ld >= channels && ld % 2 == 0 && batch <= 65535, "tlt_rows_to_nchw: ld must be even and >= channels; batch <= 65535");
  if (span_ok(channels, spatial, ld, y)) {
    const long long units = batch * (channels >> 6);
    const unsigned blocks = (unsigned)(units < 148LL * 8 ? units : 148LL * 8);
    rows_to_nchw_span_kernel<<<blocks, 256, (size_t)64 * spatial * 2, (cudaStream_t)stream>>>((const __nv_bfloat16*)rows, (int)channels, (int)spatial, ld,
                                                                                            (__nv_bfloat16*)y, units);
    return check_launch("rows_to_nchw_span_kernel");
  }
  dim3 grid((unsigned)((spatial + 31) / 32), (unsigned)((channels + 63) / 64), (unsigned)batch);
  rows_to_nchw_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)rows, channels, spatial, ld, (__nv_bfloat16*)y);
  return check_launch("rows_to_nchw_kernel");
}

extern "C" int tlt_im2col_rows(const tlt_conv_desc* d, const void* z, void* cols, void* stream) {
  RowsGeom g;
  int rc = fill_rows(d, &g);
  if (rc) return rc;
  TLT_REQUIRE(z && cols && ((uintptr_t)z % 16) == 0 && ((uintptr_t)cols % 16) == 0, "tlt_im2col_rows: tensors must be 16-byte aligned");
  const long long total = g.batch * g.out_sp * (g.ld >> 3);
  if (g.ndim == 2 && g.kernel[0] == 3 && g.kernel[1] == 3 && g.in_sp < (1 << 24)) {
    im2col_rows2d_kernel<3, 3><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(g, (const uint4*)z, (uint4*)cols);
    return check_launch("im2col_rows2d_kernel");
  }
  im2col_rows_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(g, (const uint4*)z, (uint4*)cols);
  return check_launch("im2col_rows_kernel");
}

extern "C" int tlt_col2im_rows(const tlt_conv_desc* d, const void* dcols, void* dz, void* stream) {
  RowsGeom g;
  int rc = fill_rows(d, &g);
  if (rc) return rc;
  TLT_REQUIRE(dz && dcols && ((uintptr_t)dz % 16) == 0 && ((uintptr_t)dcols % 16) == 0, "tlt_col2im_rows: tensors must be 16-byte aligned");
  const long long total = g.batch * g.in_sp * (g.ld >> 3);
  if (g.ndim == 2 && g.kernel[0] == 3 && g.kernel[1] == 3 && g.in_sp < (1 << 24) && g.out_sp < (1 << 24)) {
    col2im_rows2d_kernel<3, 3><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(g, (const uint4*)dcols, (uint4*)dz);
    return check_launch("col2im_rows2d_kernel");
  }
  col2im_rows_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(g, (const uint4*)dcols, (uint4*)dz);
  return check_launch("col2im_rows_kernel");
}
