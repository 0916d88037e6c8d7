// Depthwise N-D convolution (forward, data gradient, weight gradient) with the (batch, channel) planes staged in shared
// memory.  Replaces the F.conv{1,2,3}d(groups=R) calls of tltorch/functional/convolution.py: the 1-D depthwise convs of
// cp_conv (:268-271 through general_conv1d :99-137) and the N-D depthwise conv of cp_conv_mobilenet (:321-332), and
// their autograd.
//
// These ops are pure streaming work -- 3..27 multiply-adds per element against 4 bytes of HBM traffic -- so the design
// goal is one coalesced pass over HBM: a block copies PB consecutive (b, c) planes (one contiguous chunk of the NCHW
// tensor) into shared memory as fp32 with 16-byte loads, every thread then produces strips of 8 consecutive outputs of one
// row from shared memory (index arithmetic amortised over the strip, divisions by launch-time constants done with
// multiply-high), and writes them with one 16-byte store.  The first version (one thread per element, 64-bit div/mod per
// element, global-memory taps) ran at 2 % of HBM speed (1.8 ms for a 222 MB pass at ResNet 56x56 shapes).
// Planes that do not fit shared memory fall back to the element-wise kernels in conv.cu.
#include "common.cuh"
#include <mutex>

namespace tlt {

struct FastDivU {
  uint32_t d, m;
  __host__ __device__ FastDivU() : d(1), m(0) {}
  __host__ __device__ explicit FastDivU(int dd) : d((uint32_t)dd), m(dd > 1 ? (uint32_t)((1ull << 32) / (uint32_t)dd) + 1u : 0u) {}
  __device__ __forceinline__ int div(int n) const { return d > 1 ? (int)__umulhi((uint32_t)n, m) : n; }
  __device__ __forceinline__ void divmod(int n, int& q, int& r) const { q = div(n); r = n - q * (int)d; }
};

constexpr int kSeg = 8;            // outputs per thread strip (one 16-byte bf16 store)
constexpr int kMaxTaps = 27;

struct DwParams {
  int in[3], out[3], ker[3], stride[3], pad[3], dil[3];
  int in_sp, out_sp, taps;
  int channels;
  long long planes;              // batch * channels
  int PB;                        // planes per block
  int segs;                      // strips per output row = ceil(out[2] / kSeg)   (dgrad: per input row)
  FastDivU f_segs, f_rows1, f_items_per_plane, f_ch;
};

// global (contiguous, n elements) -> shared fp32, 16-byte loads when aligned
template <typename T>
__device__ __forceinline__ void stage_chunk(float* dst, const T* __restrict__ src, int n) {
  constexpr int V = 16 / (int)sizeof(T);
  if ((((uintptr_t)src) & 15) == 0 && (sizeof(T) == 2 || (((uintptr_t)dst) & 15) == 0)) {
    const int nv = n / V;
    for (int e = threadIdx.x; e < nv; e += blockDim.x) {
      const uint4 raw = __ldg((const uint4*)src + e);
      if (sizeof(T) == 2) {
        const __nv_bfloat162* h = (const __nv_bfloat162*)&raw;
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); dst[e * V + 2 * j] = f.x; dst[e * V + 2 * j + 1] = f.y; }
      } else {
        *(float4*)(dst + e * V) = make_float4(__uint_as_float(raw.x), __uint_as_float(raw.y), __uint_as_float(raw.z), __uint_as_float(raw.w));
      }
    }
    for (int e = nv * V + threadIdx.x; e < n; e += blockDim.x) dst[e] = to_f32<T>(src[e]);
  } else {
    for (int e = threadIdx.x; e < n; e += blockDim.x) dst[e] = to_f32<T>(src[e]);
  }
}

template <typename T>
__device__ __forceinline__ void store_strip(T* dst, const float* v, int n_valid) {
  if (n_valid == kSeg && (((uintptr_t)dst) & (sizeof(T) * kSeg - 1)) == 0) {
    if (sizeof(T) == 2) {
      uint4 o;
      __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
      o.x = *(uint32_t*)&h0; o.y = *(uint32_t*)&h1; o.z = *(uint32_t*)&h2; o.w = *(uint32_t*)&h3;
      *(uint4*)dst = o;
    } else {
      *(float4*)dst = make_float4(v[0], v[1], v[2], v[3]);
      *(float4*)((float*)dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
  } else {
    for (int j = 0; j < n_valid; ++j) dst[j] = from_f32<T>(v[j]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// forward:  y[bc, o0,o1,o2] = sum_t x[bc, o*s - p + t*d] * w[c, t]
// grid = ceil(planes / PB); smem: xs [PB][in_sp] | ws [PB][taps]
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) dwconv_fwd_staged(const DwParams p, const T* __restrict__ x, const T* __restrict__ w,
                                                         T* __restrict__ y) {
  extern __shared__ __align__(16) float dsm[];
  const long long bc0 = (long long)blockIdx.x * p.PB;
  const int npl = (int)min((long long)p.PB, p.planes - bc0);
  float* xs = dsm;
  float* ws = dsm + p.PB * p.in_sp;
  stage_chunk<T>(xs, x + bc0 * p.in_sp, npl * p.in_sp);
  for (int e = threadIdx.x; e < npl * p.taps; e += blockDim.x) {
    const int j = e / p.taps, t = e - j * p.taps;
    const int c = (int)((bc0 + j) % p.channels);
    ws[e] = to_f32<T>(w[(long long)c * p.taps + t]);
  }
  __syncthreads();
  const int items_per_plane = p.out[0] * p.out[1] * p.segs;
  const int items = npl * items_per_plane;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    int j, r, row, sg, o0, o1;
    p.f_items_per_plane.divmod(it, j, r);
    p.f_segs.divmod(r, row, sg);
    p.f_rows1.divmod(row, o0, o1);
    const int o2b = sg * kSeg;
    const int nv = min(kSeg, p.out[2] - o2b);
    const float* xp = xs + j * p.in_sp;
    const float* wp = ws + j * p.taps;
    float acc[kSeg];
#pragma unroll
    for (int q = 0; q < kSeg; ++q) acc[q] = 0.f;
    int t = 0;
    for (int t0 = 0; t0 < p.ker[0]; ++t0) {
      const int i0 = o0 * p.stride[0] - p.pad[0] + t0 * p.dil[0];
      for (int t1 = 0; t1 < p.ker[1]; ++t1) {
        const int i1 = o1 * p.stride[1] - p.pad[1] + t1 * p.dil[1];
        const bool row_ok = i0 >= 0 && i0 < p.in[0] && i1 >= 0 && i1 < p.in[1];
        const float* xr = xp + (i0 * p.in[1] + i1) * p.in[2];
        for (int t2 = 0; t2 < p.ker[2]; ++t2, ++t) {
          if (!row_ok) continue;
          const float wv = wp[t];
          const int ib = o2b * p.stride[2] - p.pad[2] + t2 * p.dil[2];
#pragma unroll
          for (int q = 0; q < kSeg; ++q) {
            const int i2 = ib + q * p.stride[2];
            if (i2 >= 0 && i2 < p.in[2]) acc[q] = fmaf(xr[i2], wv, acc[q]);
          }
        }
      }
    }
    store_strip<T>(y + (bc0 + j) * p.out_sp + (long long)row * p.out[2] + o2b, acc, nv);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// data gradient:  dx[bc, i] = sum_{t : (i + p - t*d) % s == 0} dy[bc, (i + p - t*d) / s] * w[c, t]
// grid = ceil(planes / PB); smem: dys [PB][out_sp] | ws [PB][taps].  Strips run along the input row.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) dwconv_dgrad_staged(const DwParams p, const T* __restrict__ dy, const T* __restrict__ w,
                                                           T* __restrict__ dx) {
  extern __shared__ __align__(16) float dsm[];
  const long long bc0 = (long long)blockIdx.x * p.PB;
  const int npl = (int)min((long long)p.PB, p.planes - bc0);
  float* dys = dsm;
  float* ws = dsm + p.PB * p.out_sp;
  stage_chunk<T>(dys, dy + bc0 * p.out_sp, npl * p.out_sp);
  for (int e = threadIdx.x; e < npl * p.taps; e += blockDim.x) {
    const int j = e / p.taps, t = e - j * p.taps;
    const int c = (int)((bc0 + j) % p.channels);
    ws[e] = to_f32<T>(w[(long long)c * p.taps + t]);
  }
  __syncthreads();
  const int items_per_plane = p.in[0] * p.in[1] * p.segs;
  const int items = npl * items_per_plane;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    int j, r, row, sg, i0, i1;
    p.f_items_per_plane.divmod(it, j, r);
    p.f_segs.divmod(r, row, sg);
    p.f_rows1.divmod(row, i0, i1);
    const int i2b = sg * kSeg;
    const int nv = min(kSeg, p.in[2] - i2b);
    const float* dp = dys + j * p.out_sp;
    const float* wp = ws + j * p.taps;
    float acc[kSeg];
#pragma unroll
    for (int q = 0; q < kSeg; ++q) acc[q] = 0.f;
    int t = 0;
    for (int t0 = 0; t0 < p.ker[0]; ++t0) {
      const int n0 = i0 + p.pad[0] - t0 * p.dil[0];
      const int o0 = n0 / p.stride[0];
      const bool ok0 = n0 >= 0 && n0 == o0 * p.stride[0] && o0 < p.out[0];
      for (int t1 = 0; t1 < p.ker[1]; ++t1) {
        const int n1 = i1 + p.pad[1] - t1 * p.dil[1];
        const int o1 = n1 / p.stride[1];
        const bool row_ok = ok0 && n1 >= 0 && n1 == o1 * p.stride[1] && o1 < p.out[1];
        const float* dr = dp + (o0 * p.out[1] + o1) * p.out[2];
        for (int t2 = 0; t2 < p.ker[2]; ++t2, ++t) {
          if (!row_ok) continue;
          const float wv = wp[t];
          const int nb = i2b + p.pad[2] - t2 * p.dil[2];
          if (p.stride[2] == 1) {
#pragma unroll
            for (int q = 0; q < kSeg; ++q) {
              const int o2 = nb + q;
              if (o2 >= 0 && o2 < p.out[2]) acc[q] = fmaf(dr[o2], wv, acc[q]);
            }
          } else {
#pragma unroll
            for (int q = 0; q < kSeg; ++q) {
              const int n2 = nb + q;
              const int o2 = n2 / p.stride[2];
              if (n2 >= 0 && n2 == o2 * p.stride[2] && o2 < p.out[2]) acc[q] = fmaf(dr[o2], wv, acc[q]);
            }
          }
        }
      }
    }
    store_strip<T>(dx + (bc0 + j) * p.in_sp + (long long)row * p.in[2] + i2b, acc, nv);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// weight gradient:  dw[c, t] += sum_{b, o} dy[b,c,o] * x[b,c, o*s - p + t*d]
// grid = (channels, batch chunks); a block walks its batch range PBW planes at a time (planes of one channel are
// in_sp * channels apart, so each is staged separately); every thread keeps all tap sums in registers.
// smem: xs [PBW][in_sp] | dys [PBW][out_sp] | red [taps][8]
// ---------------------------------------------------------------------------------------------------------------------
template <typename T, int MAXT>
__global__ void __launch_bounds__(256) dwconv_wgrad_staged(const DwParams p, int batch, int batch_per_block, int PBW,
                                                           const T* __restrict__ x, const T* __restrict__ dy,
                                                           float* __restrict__ dw) {
  extern __shared__ __align__(16) float dsm[];
  float* xs = dsm;
  float* dys = xs + PBW * p.in_sp;
  float* red = dys + PBW * p.out_sp;
  const int c = blockIdx.x;
  const int b_begin = blockIdx.y * batch_per_block;
  const int b_end = min(batch, b_begin + batch_per_block);
  float acc[MAXT];
#pragma unroll
  for (int t = 0; t < MAXT; ++t) acc[t] = 0.f;
  const int items_per_plane = p.out[0] * p.out[1] * p.segs;
  for (int b0 = b_begin; b0 < b_end; b0 += PBW) {
    const int npl = min(PBW, b_end - b0);
    __syncthreads();                                   // previous round's readers are done with xs / dys
    for (int j = 0; j < npl; ++j) {
      const long long bc = (long long)(b0 + j) * p.channels + c;
      stage_chunk<T>(xs + j * p.in_sp, x + bc * p.in_sp, p.in_sp);
      stage_chunk<T>(dys + j * p.out_sp, dy + bc * p.out_sp, p.out_sp);
    }
    __syncthreads();
    const int items = npl * items_per_plane;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
      int j, r, row, sg, o0, o1;
      p.f_items_per_plane.divmod(it, j, r);
      p.f_segs.divmod(r, row, sg);
      p.f_rows1.divmod(row, o0, o1);
      const int o2b = sg * kSeg;
      const float* xp = xs + j * p.in_sp;
      const float* gp = dys + j * p.out_sp + row * p.out[2] + o2b;
      float g[kSeg];
#pragma unroll
      for (int q = 0; q < kSeg; ++q) g[q] = (o2b + q < p.out[2]) ? gp[q] : 0.f;
      int t = 0;
      for (int t0 = 0; t0 < p.ker[0]; ++t0) {
        const int i0 = o0 * p.stride[0] - p.pad[0] + t0 * p.dil[0];
        for (int t1 = 0; t1 < p.ker[1]; ++t1) {
          const int i1 = o1 * p.stride[1] - p.pad[1] + t1 * p.dil[1];
          const bool row_ok = i0 >= 0 && i0 < p.in[0] && i1 >= 0 && i1 < p.in[1];
          const float* xr = xp + (i0 * p.in[1] + i1) * p.in[2];
          for (int t2 = 0; t2 < p.ker[2]; ++t2, ++t) {
            if (!row_ok) continue;
            const int ib = o2b * p.stride[2] - p.pad[2] + t2 * p.dil[2];
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < kSeg; ++q) {
              const int i2 = ib + q * p.stride[2];
              if (i2 >= 0 && i2 < p.in[2]) s = fmaf(xr[i2], g[q], s);
            }
            // acc[t] += s with a compile-time register index
#pragma unroll
            for (int u = 0; u < MAXT; ++u)
              if (u == t) acc[u] += s;
          }
        }
      }
    }
  }
  // block reduction of every tap sum, then one atomic per (c, t) per block
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int t = 0; t < MAXT; ++t) {
    if (t < p.taps) {
      float v = acc[t];
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
      if (lane == 0) red[t * 8 + warp] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < p.taps) {
    float v = 0.f;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) v += red[threadIdx.x * 8 + wv];
    atomicAdd(dw + (long long)c * p.taps + threadIdx.x, v);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
int dwconv_fwd_generic(const tlt_conv_desc* d, const void* x, const void* w, void* y, void* stream);
int dwconv_dgrad_generic(const tlt_conv_desc* d, const void* dy, const void* w, void* dx, void* stream);
int dwconv_wgrad_generic(const tlt_conv_desc* d, const void* x, const void* dy, float* dw, void* stream);

constexpr size_t kDwSmemLimit = 160 * 1024;

static bool fill_dw(const tlt_conv_desc* d, DwParams* p) {
  if (!d || d->ndim < 1 || d->ndim > 3 || d->batch <= 0 || d->channels <= 0) return false;
  const int off = 3 - d->ndim;
  for (int i = 0; i < 3; ++i) { p->in[i] = 1; p->out[i] = 1; p->ker[i] = 1; p->stride[i] = 1; p->pad[i] = 0; p->dil[i] = 1; }
  long long in_sp = 1, out_sp = 1, taps = 1;
  for (int i = 0; i < d->ndim; ++i) {
    if (d->kernel[i] < 1 || d->stride[i] < 1 || d->dilation[i] < 1 || d->padding[i] < 0) return false;
    if (d->in_size[i] < 1 || d->out_size[i] < 1) return false;
    p->in[off + i] = (int)d->in_size[i]; p->out[off + i] = (int)d->out_size[i]; p->ker[off + i] = d->kernel[i];
    p->stride[off + i] = d->stride[i]; p->pad[off + i] = d->padding[i]; p->dil[off + i] = d->dilation[i];
    in_sp *= d->in_size[i]; out_sp *= d->out_size[i]; taps *= d->kernel[i];
  }
  if (in_sp > (1 << 20) || out_sp > (1 << 20) || taps > kMaxTaps || d->channels > (1 << 24)) return false;
  if (d->batch * d->channels >= (1LL << 31)) return false;
  p->in_sp = (int)in_sp; p->out_sp = (int)out_sp; p->taps = (int)taps;
  p->channels = (int)d->channels;
  p->planes = d->batch * d->channels;
  return true;
}

// planes per block: enough strips for ~4 rounds of 256 threads, bounded by shared memory and by >= ~2 blocks per SM
static int choose_pb(const DwParams& p, int plane_floats, int items_per_plane) {
  long long pb = (4 * 256 + items_per_plane - 1) / items_per_plane;
  const long long by_smem = (long long)(kDwSmemLimit / 4) / (plane_floats + p.taps);
  if (pb > by_smem) pb = by_smem;
  const long long by_grid = p.planes / (2 * 148);
  if (by_grid >= 1 && pb > by_grid) pb = by_grid;
  if (pb < 1) pb = 1;
  return (int)pb;
}

template <typename K>
static int dw_set_smem(K kernel, size_t bytes) {
  static std::mutex mu;
  static size_t high = 0;            // one instance per kernel type K
  if (bytes <= 48 * 1024) return TLT_OK;
  std::lock_guard<std::mutex> lock(mu);
  if (bytes > high) {
    TLT_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    high = bytes;
  }
  return TLT_OK;
}

template <typename T>
static int launch_fwd(DwParams p, const void* x, const void* w, void* y, cudaStream_t st) {
  p.segs = (p.out[2] + kSeg - 1) / kSeg;
  const int ipp = p.out[0] * p.out[1] * p.segs;
  p.PB = choose_pb(p, p.in_sp, ipp);
  p.f_segs = FastDivU(p.segs); p.f_rows1 = FastDivU(p.out[1]); p.f_items_per_plane = FastDivU(ipp);
  const size_t smem = (size_t)p.PB * (p.in_sp + p.taps) * 4;
  int rc = dw_set_smem(dwconv_fwd_staged<T>, smem);
  if (rc) return rc;
  const long long blocks = (p.planes + p.PB - 1) / p.PB;
  dwconv_fwd_staged<T><<<(unsigned)blocks, 256, smem, st>>>(p, (const T*)x, (const T*)w, (T*)y);
  return check_launch("dwconv_fwd_staged");
}

template <typename T>
static int launch_dgrad(DwParams p, const void* dy, const void* w, void* dx, cudaStream_t st) {
  p.segs = (p.in[2] + kSeg - 1) / kSeg;
  const int ipp = p.in[0] * p.in[1] * p.segs;
  p.PB = choose_pb(p, p.out_sp, ipp);
  p.f_segs = FastDivU(p.segs); p.f_rows1 = FastDivU(p.in[1]); p.f_items_per_plane = FastDivU(ipp);
  const size_t smem = (size_t)p.PB * (p.out_sp + p.taps) * 4;
  int rc = dw_set_smem(dwconv_dgrad_staged<T>, smem);
  if (rc) return rc;
  const long long blocks = (p.planes + p.PB - 1) / p.PB;
  dwconv_dgrad_staged<T><<<(unsigned)blocks, 256, smem, st>>>(p, (const T*)dy, (const T*)w, (T*)dx);
  return check_launch("dwconv_dgrad_staged");
}

template <typename T, int MAXT>
static int launch_wgrad_t(DwParams p, int batch, const void* x, const void* dy, float* dw, cudaStream_t st) {
  p.segs = (p.out[2] + kSeg - 1) / kSeg;
  const int ipp = p.out[0] * p.out[1] * p.segs;
  p.f_segs = FastDivU(p.segs); p.f_rows1 = FastDivU(p.out[1]); p.f_items_per_plane = FastDivU(ipp);
  // batch chunks: ~4 blocks per SM in total
  long long chunks = (148LL * 4 + p.channels - 1) / p.channels;
  if (chunks > batch) chunks = batch;
  if (chunks < 1) chunks = 1;
  if (chunks > 65535) chunks = 65535;
  const int bpb = (int)((batch + chunks - 1) / chunks);
  chunks = (batch + bpb - 1) / bpb;
  long long pbw = (4 * 256 + ipp - 1) / ipp;
  const long long by_smem = (long long)((kDwSmemLimit - 4 * 8 * kMaxTaps) / 4) / (p.in_sp + p.out_sp);
  if (pbw > by_smem) pbw = by_smem;
  if (pbw > bpb) pbw = bpb;
  if (pbw < 1) pbw = 1;
  const size_t smem = ((size_t)pbw * (p.in_sp + p.out_sp) + 8 * p.taps) * 4;
  int rc = dw_set_smem(dwconv_wgrad_staged<T, MAXT>, smem);
  if (rc) return rc;
  dim3 grid((unsigned)p.channels, (unsigned)chunks);
  dwconv_wgrad_staged<T, MAXT><<<grid, 256, smem, st>>>(p, batch, bpb, (int)pbw, (const T*)x, (const T*)dy, dw);
  return check_launch("dwconv_wgrad_staged");
}

template <typename T>
static int launch_wgrad(const DwParams& p, int batch, const void* x, const void* dy, float* dw, cudaStream_t st) {
  if (p.taps <= 3) return launch_wgrad_t<T, 3>(p, batch, x, dy, dw, st);
  if (p.taps <= 9) return launch_wgrad_t<T, 9>(p, batch, x, dy, dw, st);
  return launch_wgrad_t<T, kMaxTaps>(p, batch, x, dy, dw, st);
}

}  // namespace tlt

using namespace tlt;

extern "C" int tlt_dwconv_fwd(const tlt_conv_desc* d, const void* x, const void* w, void* y, void* stream) {
  DwParams p;
  if (!fill_dw(d, &p) || (size_t)(p.in_sp + p.taps) * 4 > kDwSmemLimit || (d->dtype != TLT_F32 && d->dtype != TLT_BF16))
    return dwconv_fwd_generic(d, x, w, y, stream);
  TLT_REQUIRE(x && w && y, "tlt_dwconv_fwd: null pointer");
  for (int i = 0; i < d->ndim; ++i) {
    long long expect = ((long long)d->in_size[i] + 2LL * d->padding[i] - (long long)d->dilation[i] * (d->kernel[i] - 1) - 1) / d->stride[i] + 1;
    TLT_REQUIRE(expect == d->out_size[i], "conv: out_size[%d]=%lld inconsistent (expected %lld)", i, (long long)d->out_size[i], expect);
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (d->dtype == TLT_BF16) return launch_fwd<__nv_bfloat16>(p, x, w, y, st);
  return launch_fwd<float>(p, x, w, y, st);
}

extern "C" int tlt_dwconv_dgrad(const tlt_conv_desc* d, const void* dy, const void* w, void* dx, void* stream) {
  DwParams p;
  if (!fill_dw(d, &p) || (size_t)(p.out_sp + p.taps) * 4 > kDwSmemLimit || (d->dtype != TLT_F32 && d->dtype != TLT_BF16))
    return dwconv_dgrad_generic(d, dy, w, dx, stream);
  TLT_REQUIRE(dy && w && dx, "tlt_dwconv_dgrad: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->dtype == TLT_BF16) return launch_dgrad<__nv_bfloat16>(p, dy, w, dx, st);
  return launch_dgrad<float>(p, dy, w, dx, st);
}

extern "C" int tlt_dwconv_wgrad(const tlt_conv_desc* d, const void* x, const void* dy, float* dw, void* stream) {
  DwParams p;
  if (!fill_dw(d, &p) || (size_t)(p.in_sp + p.out_sp + 8 * kMaxTaps) * 4 > kDwSmemLimit || d->batch >= (1LL << 31) ||
      (d->dtype != TLT_F32 && d->dtype != TLT_BF16))
    return dwconv_wgrad_generic(d, x, dy, dw, stream);
  TLT_REQUIRE(x && dy && dw, "tlt_dwconv_wgrad: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->dtype == TLT_BF16) return launch_wgrad<__nv_bfloat16>(p, (int)d->batch, x, dy, dw, st);
  return launch_wgrad<float>(p, (int)d->batch, x, dy, dw, st);
}
