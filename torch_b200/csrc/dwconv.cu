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
#include <stdlib.h>
#include <mutex>
#include <unordered_map>

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
  FastDivU f_insp, f_outsp, f_i12, f_i2, f_o12, f_o2, f_taps;
  int PPh;                       // halo-padded plane size (floats)
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

// Planes are staged with a zero halo of `pad` elements on every side, so the tap loops need no bounds checks: for a
// valid output o the index o*s - p + t*d lies in [-p, in + p).  Consecutive threads own consecutive output elements of
// the flattened plane, so shared-memory reads and global stores of a warp touch consecutive addresses (the first
// version gave each thread a strip of 8 outputs: 8-way bank conflicts on every tap, 460 us per 222 MB pass).
struct Halo {
  int P1, P2, PP;               // padded extents of dims 1, 2 and padded plane size
};
__device__ __forceinline__ Halo halo_of(const DwParams& p) {
  Halo h;
  h.P1 = p.in[1] + 2 * p.pad[1];
  h.P2 = p.in[2] + 2 * p.pad[2];
  h.PP = (p.in[0] + 2 * p.pad[0]) * h.P1 * h.P2;
  return h;
}

// global (npl contiguous planes) -> shared fp32 with halo; the halo was zeroed beforehand
template <typename T>
__device__ __forceinline__ void stage_planes_halo(float* dst, const T* __restrict__ src, int npl, const DwParams& p, const Halo& h) {
  const int n = npl * p.in_sp;
#pragma unroll 4
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    int j, sp, i0, r, i1, i2;
    p.f_insp.divmod(e, j, sp);
    p.f_i12.divmod(sp, i0, r);
    p.f_i2.divmod(r, i1, i2);
    dst[j * h.PP + ((i0 + p.pad[0]) * h.P1 + i1 + p.pad[1]) * h.P2 + i2 + p.pad[2]] = to_f32<T>(src[e]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// forward:  y[bc, o0,o1,o2] = sum_t x[bc, o*s - p + t*d] * w[c, t]
// grid = ceil(planes / PB); smem: xs [PB][PP] (halo-padded) | ws [PB][taps]
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) dwconv_fwd_staged(const DwParams p, const T* __restrict__ x, const T* __restrict__ w,
                                                         T* __restrict__ y) {
  extern __shared__ __align__(16) float dsm[];
  const Halo h = halo_of(p);
  const long long bc0 = (long long)blockIdx.x * p.PB;
  const int npl = (int)min((long long)p.PB, p.planes - bc0);
  float* xs = dsm;
  float* ws = dsm + p.PB * h.PP;
  for (int e = threadIdx.x; e < npl * h.PP; e += blockDim.x) xs[e] = 0.f;
  for (int e = threadIdx.x; e < npl * p.taps; e += blockDim.x) {
    int j, t;
    p.f_taps.divmod(e, j, t);
    const int c = (int)((bc0 + j) % p.channels);
    ws[e] = to_f32<T>(w[(long long)c * p.taps + t]);
  }
  __syncthreads();
  stage_planes_halo<T>(xs, x + bc0 * p.in_sp, npl, p, h);
  __syncthreads();
  const int n_out = npl * p.out_sp;
  T* yb = y + bc0 * p.out_sp;
  for (int e = threadIdx.x; e < n_out; e += blockDim.x) {
    int j, sp, o0, r, o1, o2;
    p.f_outsp.divmod(e, j, sp);
    p.f_o12.divmod(sp, o0, r);
    p.f_o2.divmod(r, o1, o2);
    // halo coordinates of tap (0,0,0): (o*s - p) + p = o*s
    const float* xp = xs + j * h.PP + (o0 * p.stride[0] * h.P1 + o1 * p.stride[1]) * h.P2 + o2 * p.stride[2];
    const float* wp = ws + j * p.taps;
    float acc = 0.f;
    int t = 0;
    for (int t0 = 0; t0 < p.ker[0]; ++t0)
      for (int t1 = 0; t1 < p.ker[1]; ++t1) {
        const float* xr = xp + (t0 * p.dil[0] * h.P1 + t1 * p.dil[1]) * h.P2;
        for (int t2 = 0; t2 < p.ker[2]; ++t2, ++t) acc = fmaf(xr[t2 * p.dil[2]], wp[t], acc);
      }
    yb[e] = from_f32<T>(acc);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// data gradient:  dx[bc, i] = sum_{t : (i + p - t*d) % s == 0} dy[bc, (i + p - t*d) / s] * w[c, t]
// grid = ceil(planes / PB); smem: dys [PB][out_sp] | ws [PB][taps]
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
    int j, t;
    p.f_taps.divmod(e, j, t);
    const int c = (int)((bc0 + j) % p.channels);
    ws[e] = to_f32<T>(w[(long long)c * p.taps + t]);
  }
  __syncthreads();
  const int n_in = npl * p.in_sp;
  T* dxb = dx + bc0 * p.in_sp;
  const bool unit = p.stride[0] == 1 && p.stride[1] == 1 && p.stride[2] == 1;
  for (int e = threadIdx.x; e < n_in; e += blockDim.x) {
    int j, sp, i0, r, i1, i2;
    p.f_insp.divmod(e, j, sp);
    p.f_i12.divmod(sp, i0, r);
    p.f_i2.divmod(r, i1, i2);
    const float* dp = dys + j * p.out_sp;
    const float* wp = ws + j * p.taps;
    float acc = 0.f;
    int t = 0;
    for (int t0 = 0; t0 < p.ker[0]; ++t0) {
      const int n0 = i0 + p.pad[0] - t0 * p.dil[0];
      const int o0 = unit ? n0 : n0 / p.stride[0];
      const bool ok0 = n0 >= 0 && (unit || n0 == o0 * p.stride[0]) && o0 < p.out[0];
      for (int t1 = 0; t1 < p.ker[1]; ++t1) {
        const int n1 = i1 + p.pad[1] - t1 * p.dil[1];
        const int o1 = unit ? n1 : n1 / p.stride[1];
        const bool ok1 = ok0 && n1 >= 0 && (unit || n1 == o1 * p.stride[1]) && o1 < p.out[1];
        const float* dr = dp + (o0 * p.out[1] + o1) * p.out[2];
        for (int t2 = 0; t2 < p.ker[2]; ++t2, ++t) {
          const int n2 = i2 + p.pad[2] - t2 * p.dil[2];
          const int o2 = unit ? n2 : n2 / p.stride[2];
          if (ok1 && n2 >= 0 && (unit || n2 == o2 * p.stride[2]) && o2 < p.out[2]) acc = fmaf(dr[o2], wp[t], acc);
        }
      }
    }
    dxb[e] = from_f32<T>(acc);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// weight gradient:  dw[c, t] += sum_{b, o} dy[b,c,o] * x[b,c, o*s - p + t*d]
// grid = (channels, batch chunks); a block walks its batch range PBW planes at a time (planes of one channel are
// in_sp * channels apart, so each is staged separately); every thread keeps all tap sums in registers.
// smem: xs [PBW][PP] (halo-padded) | dys [PBW][out_sp] | red [taps][8]
// ---------------------------------------------------------------------------------------------------------------------
template <typename T, int MAXT>
__global__ void __launch_bounds__(256) dwconv_wgrad_staged(const DwParams p, int batch, int batch_per_block, int PBW,
                                                           const T* __restrict__ x, const T* __restrict__ dy,
                                                           float* __restrict__ dw) {
  extern __shared__ __align__(16) float dsm[];
  const Halo h = halo_of(p);
  float* xs = dsm;
  float* dys = xs + PBW * h.PP;
  float* red = dys + PBW * p.out_sp;
  const int c = blockIdx.x;
  const int b_begin = blockIdx.y * batch_per_block;
  const int b_end = min(batch, b_begin + batch_per_block);
  float acc[MAXT];
#pragma unroll
  for (int t = 0; t < MAXT; ++t) acc[t] = 0.f;
  for (int e = threadIdx.x; e < PBW * h.PP; e += blockDim.x) xs[e] = 0.f;     // halo stays zero for the whole kernel
  for (int b0 = b_begin; b0 < b_end; b0 += PBW) {
    const int npl = min(PBW, b_end - b0);
    __syncthreads();                                   // previous round's readers are done with xs / dys (and the zero fill landed)
    for (int j = 0; j < npl; ++j) {
      const long long bc = (long long)(b0 + j) * p.channels + c;
      stage_planes_halo<T>(xs + j * h.PP, x + bc * p.in_sp, 1, p, h);
      stage_chunk<T>(dys + j * p.out_sp, dy + bc * p.out_sp, p.out_sp);
    }
    __syncthreads();
    const int n_out = npl * p.out_sp;
    for (int e = threadIdx.x; e < n_out; e += blockDim.x) {
      int j, sp, o0, r, o1, o2;
      p.f_outsp.divmod(e, j, sp);
      p.f_o12.divmod(sp, o0, r);
      p.f_o2.divmod(r, o1, o2);
      const float g = dys[e];
      const float* xp = xs + j * h.PP + (o0 * p.stride[0] * h.P1 + o1 * p.stride[1]) * h.P2 + o2 * p.stride[2];
      if (MAXT <= 9) {
        // taps fully unrolled against compile-time register indices (tap index -> offsets through small loops)
        int t = 0;
        for (int t0 = 0; t0 < p.ker[0]; ++t0)
          for (int t1 = 0; t1 < p.ker[1]; ++t1) {
            const float* xr = xp + (t0 * p.dil[0] * h.P1 + t1 * p.dil[1]) * h.P2;
            for (int t2 = 0; t2 < p.ker[2]; ++t2, ++t) {
              const float s = xr[t2 * p.dil[2]] * g;
#pragma unroll
              for (int u = 0; u < MAXT; ++u)
                if (u == t) acc[u] += s;
            }
          }
      } else {
        int t = 0;
        for (int t0 = 0; t0 < p.ker[0]; ++t0)
          for (int t1 = 0; t1 < p.ker[1]; ++t1) {
            const float* xr = xp + (t0 * p.dil[0] * h.P1 + t1 * p.dil[1]) * h.P2;
            for (int t2 = 0; t2 < p.ker[2]; ++t2, ++t) {
              const float s = xr[t2 * p.dil[2]] * g;
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
// Fast path: 2-D 3x3 taps, stride 1, padding 1, dilation 1 (every stride-1 3x3 conv of a ResNet -- 13 of the 16 CP layers of
// BASELINE config 3 -- and, with the taps reversed, its data gradient).  The generic staged kernels above spend ~80
// instructions per output (runtime tap loops, per-element index decoding while staging); here
//   * planes are staged with 16-/8-byte loads into rows of pitch RP whose interior starts at float offset 4, so the halo
//     column sits at offset 3 and every 4-output strip reads one aligned float4 plus two scalars per input row;
//   * a thread owns a strip of 4 consecutive outputs: 3 rows x 6 inputs in registers feed 36 FMAs with compile-time taps;
//   * consecutive lanes own consecutive strips: float4 reads of a quarter-warp cover 128 contiguous bytes (no conflicts),
//     and the 8-byte bf16 stores of a warp cover 256 contiguous bytes.
// ~17 instructions per output, i.e. within reach of the HBM roofline of these 4-bytes-per-output passes.
//   FLIP = false: y = corr(x, w)            FLIP = true : dx = corr(dy, reversed w)   (the adjoint at stride 1, pad 1)
//   smem: xs [PB][H + 2][RP] | ws [PB][12]
// ---------------------------------------------------------------------------------------------------------------------
struct Fast3Params {
  int H, W, RP, PB, strips;        // rows, cols, row pitch (floats), planes per block, strips per row
  int channels;
  long long planes;
  FastDivU f_spp, f_strips, f_rowv, f_rowsv;   // strips per plane, strips per row, vectors per row, vectors per plane
};

// stages npl planes whose starts are `plane_stride` elements apart (H*W for consecutive planes, channels*H*W for the
// planes of one channel across images)
template <typename T, int V>
__device__ __forceinline__ void fast3_stage(float* xs, const T* __restrict__ src, int npl, long long plane_stride,
                                            const Fast3Params& q) {
  // zero the halo rows / columns (interior is fully overwritten below); plane = (H + 2) rows of RP floats
  const int plane = (q.H + 2) * q.RP;
  for (int e = threadIdx.x; e < npl * (q.H + 2); e += blockDim.x) {      // one staged row per thread
    const int j = e / (q.H + 2), row = e - j * (q.H + 2);
    float* rp = xs + j * plane + row * q.RP;
    if (row == 0 || row == q.H + 1) {
      for (int c = 0; c < q.RP; c += 4) *(float4*)(rp + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      *(float4*)rp = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = 4 + q.W; c < q.RP; ++c) rp[c] = 0.f;
    }
  }
  const int vec_per_row = q.W / V, vec_per_plane = q.H * vec_per_row;
  const int total = npl * vec_per_plane;
#pragma unroll 2
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    int j, r, row, cv;
    q.f_rowsv.divmod(e, j, r);
    q.f_rowv.divmod(r, row, cv);
    float* dst = xs + j * plane + (row + 1) * q.RP + 4 + cv * V;
    const T* sp = src + (long long)j * plane_stride + row * q.W + cv * V;
    if (V == 1) {
      dst[0] = to_f32<T>(sp[0]);
    } else if (sizeof(T) == 2) {
      if (V == 8) {
        const uint4 raw = __ldg((const uint4*)sp);
        const __nv_bfloat162* h = (const __nv_bfloat162*)&raw;
        const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]), c = __bfloat1622float2(h[2]), d = __bfloat1622float2(h[3]);
        *(float4*)dst = make_float4(a.x, a.y, b.x, b.y);
        *(float4*)(dst + 4) = make_float4(c.x, c.y, d.x, d.y);
      } else {
        const uint2 raw = __ldg((const uint2*)sp);
        const __nv_bfloat162* h = (const __nv_bfloat162*)&raw;
        const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
        *(float4*)dst = make_float4(a.x, a.y, b.x, b.y);
      }
    } else {
      *(float4*)dst = __ldg((const float4*)sp);
      if (V == 8) *(float4*)(dst + 4) = __ldg((const float4*)sp + 1);
    }
  }
}

template <typename T, int V, bool FLIP>
__global__ void __launch_bounds__(256) dwconv3x3_fast(const Fast3Params q, const T* __restrict__ x, const T* __restrict__ w,
                                                      T* __restrict__ y) {
  extern __shared__ __align__(16) float dsm[];
  const int plane = (q.H + 2) * q.RP;
  const long long bc0 = (long long)blockIdx.x * q.PB;
  const int npl = (int)min((long long)q.PB, q.planes - bc0);
  float* xs = dsm;
  float* ws = dsm + q.PB * plane;
  for (int e = threadIdx.x; e < npl * 9; e += blockDim.x) {
    const int j = e / 9, t = e - j * 9;
    const int c = (int)((bc0 + j) % q.channels);
    ws[j * 12 + (FLIP ? 8 - t : t)] = to_f32<T>(w[(long long)c * 9 + t]);
  }
  fast3_stage<T, V>(xs, x + bc0 * q.H * q.W, npl, (long long)q.H * q.W, q);
  __syncthreads();
  const int spp = q.H * q.strips;
  const int total = npl * spp;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    int j, r, row, sg;
    q.f_spp.divmod(e, j, r);
    q.f_strips.divmod(r, row, sg);
    const float* xr = xs + j * plane + row * q.RP + 4 + 4 * sg;      // output row `row` reads staged rows row .. row+2
    const float* wp = ws + j * 12;
    float wk[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) wk[t] = wp[t];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t1 = 0; t1 < 3; ++t1) {
      const float* rr = xr + t1 * q.RP;
      const float4 mid = *(const float4*)rr;
      const float in[6] = {rr[-1], mid.x, mid.y, mid.z, mid.w, rr[4]};
#pragma unroll
      for (int t2 = 0; t2 < 3; ++t2)
#pragma unroll
        for (int o = 0; o < 4; ++o) acc[o] = fmaf(in[o + t2], wk[t1 * 3 + t2], acc[o]);
    }
    T* dst = y + (bc0 + j) * (long long)q.H * q.W + row * q.W + 4 * sg;
    const int nv = min(4, q.W - 4 * sg);
    if (nv == 4 && (((uintptr_t)dst) & (sizeof(T) * 4 - 1)) == 0) {
      if (sizeof(T) == 2) {
        __nv_bfloat162 a = __floats2bfloat162_rn(acc[0], acc[1]), b = __floats2bfloat162_rn(acc[2], acc[3]);
        uint2 o; o.x = *(uint32_t*)&a; o.y = *(uint32_t*)&b;
        *(uint2*)dst = o;
      } else {
        *(float4*)dst = make_float4(acc[0], acc[1], acc[2], acc[3]);
      }
    } else {
      for (int o = 0; o < nv; ++o) dst[o] = from_f32<T>(acc[o]);
    }
  }
}

// weight gradient of the same configuration:  dw[c, t1, t2] += sum_{b, h, w} dy[b,c,h,w] * x[b,c,h + t1 - 1, w + t2 - 1]
// grid = (channels, batch chunks); smem: xs [PBW][H + 2][RP] | dys [PBW][H + 2][RP] (same pitched layout) | red [9][8]
template <typename T, int V>
__global__ void __launch_bounds__(256) dwconv3x3_wgrad_fast(const Fast3Params q, int batch, int batch_per_block, int PBW,
                                                            const T* __restrict__ x, const T* __restrict__ dy,
                                                            float* __restrict__ dw) {
  extern __shared__ __align__(16) float dsm[];
  const int plane = (q.H + 2) * q.RP;
  float* xs = dsm;
  float* gs = xs + PBW * plane;
  float* red = gs + PBW * plane;
  const int c = blockIdx.x;
  const int b_begin = blockIdx.y * batch_per_block;
  const int b_end = min(batch, b_begin + batch_per_block);
  float acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = 0.f;
  const int spp = q.H * q.strips;
  const long long psz = (long long)q.H * q.W;
  for (int b0 = b_begin; b0 < b_end; b0 += PBW) {
    const int npl = min(PBW, b_end - b0);
    __syncthreads();
    {
      const long long first = ((long long)b0 * q.channels + c) * psz;
      fast3_stage<T, V>(xs, x + first, npl, (long long)q.channels * psz, q);
      fast3_stage<T, V>(gs, dy + first, npl, (long long)q.channels * psz, q);
    }
    __syncthreads();
    const int total = npl * spp;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
      int j, r, row, sg;
      q.f_spp.divmod(e, j, r);
      q.f_strips.divmod(r, row, sg);
      const float4 g4 = *(const float4*)(gs + j * plane + (row + 1) * q.RP + 4 + 4 * sg);
      const int nv = q.W - 4 * sg;                                     // strips past the row end carry staged garbage: mask
      const float g[4] = {g4.x, nv > 1 ? g4.y : 0.f, nv > 2 ? g4.z : 0.f, nv > 3 ? g4.w : 0.f};
      const float* xr = xs + j * plane + row * q.RP + 4 + 4 * sg;
#pragma unroll
      for (int t1 = 0; t1 < 3; ++t1) {
        const float* rr = xr + t1 * q.RP;
        const float4 mid = *(const float4*)rr;
        const float in[6] = {rr[-1], mid.x, mid.y, mid.z, mid.w, rr[4]};
#pragma unroll
        for (int t2 = 0; t2 < 3; ++t2)
#pragma unroll
          for (int o = 0; o < 4; ++o) acc[t1 * 3 + t2] = fmaf(in[o + t2], g[o], acc[t1 * 3 + t2]);
      }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    float v = acc[t];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if (lane == 0) red[t * 8 + warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    float v = 0.f;
    for (int wv = 0; wv < 8; ++wv) v += red[threadIdx.x * 8 + wv];
    atomicAdd(dw + (long long)c * 9 + threadIdx.x, v);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Unfold for dense (non-grouped) convolutions:  cols[(b, o), (c, t)] = x[b, c, o*s - p + t*d]   (zero outside)
// and its adjoint  dx[b, c, i] = sum_{(o,t) -> i} dcols[(b, o), (c, t)].
// Replaces the unfold half of F.conv{1,2,3}d for the Tucker core conv (functional/convolution.py:161) and the TT per-mode
// convs (:213-216).  Block = (image, chunk of CC channels): the chunk's planes are staged in shared memory once, after
// which every row (b, o) of cols gets one contiguous run of CC * taps elements written with coalesced stores (im2col),
// or -- for the adjoint -- the (out_sp x CC*taps) slab of dcols is staged with coalesced loads and each dx plane is
// gathered from shared memory.  The element-wise first versions spent 1.0 / 0.37 ms on the 175 MB cols tensor of config 4.
// ---------------------------------------------------------------------------------------------------------------------
struct UnfoldParams {
  DwParams g;                    // geometry (channels = C)
  int batch, CC, n_cc;           // images, channels per block, chunks per image
  FastDivU f_taps, f_k12, f_k2, f_o12, f_o2, f_i12, f_i2, f_run;
};

template <typename T>
__global__ void __launch_bounds__(256) im2col_staged(const UnfoldParams u, const T* __restrict__ x, T* __restrict__ cols) {
  extern __shared__ __align__(16) float dsm[];
  const DwParams& p = u.g;
  const int b = blockIdx.x / u.n_cc, cc = blockIdx.x - b * u.n_cc;
  const int c0 = cc * u.CC, cn = min(u.CC, p.channels - c0);
  stage_chunk<T>(dsm, x + ((long long)b * p.channels + c0) * p.in_sp, cn * p.in_sp);
  __syncthreads();
  const int run = cn * p.taps;                                   // contiguous run of one cols row owned by this block
  const long long kdim = (long long)p.channels * p.taps;
  for (int o = 0; o < p.out_sp; ++o) {
    int o0, r, o1, o2;
    u.f_o12.divmod(o, o0, r);
    u.f_o2.divmod(r, o1, o2);
    T* dst = cols + ((long long)b * p.out_sp + o) * kdim + (long long)c0 * p.taps;
    for (int e = threadIdx.x; e < run; e += blockDim.x) {
      int c, t, t0, t1, t2, q;
      u.f_taps.divmod(e, c, t);
      u.f_k12.divmod(t, t0, q);
      u.f_k2.divmod(q, t1, t2);
      const int i0 = o0 * p.stride[0] - p.pad[0] + t0 * p.dil[0];
      const int i1 = o1 * p.stride[1] - p.pad[1] + t1 * p.dil[1];
      const int i2 = o2 * p.stride[2] - p.pad[2] + t2 * p.dil[2];
      float v = 0.f;
      if (i0 >= 0 && i0 < p.in[0] && i1 >= 0 && i1 < p.in[1] && i2 >= 0 && i2 < p.in[2])
        v = dsm[c * p.in_sp + (i0 * p.in[1] + i1) * p.in[2] + i2];
      dst[e] = from_f32<T>(v);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) col2im_staged(const UnfoldParams u, const T* __restrict__ dcols, T* __restrict__ dx) {
  extern __shared__ __align__(16) float dsm[];                   // slab [out_sp][cn * taps]
  const DwParams& p = u.g;
  const int b = blockIdx.x / u.n_cc, cc = blockIdx.x - b * u.n_cc;
  const int c0 = cc * u.CC, cn = min(u.CC, p.channels - c0);
  const int run = cn * p.taps;
  const long long kdim = (long long)p.channels * p.taps;
  const T* src = dcols + (long long)b * p.out_sp * kdim + (long long)c0 * p.taps;
  const FastDivU f_run(run);                                    // the last chunk of an image may be ragged
  for (int e = threadIdx.x; e < p.out_sp * run; e += blockDim.x) {
    int o, q;
    f_run.divmod(e, o, q);
    dsm[e] = to_f32<T>(src[(long long)o * kdim + q]);
  }
  __syncthreads();
  T* dst = dx + ((long long)b * p.channels + c0) * p.in_sp;
  const bool unit = p.stride[0] == 1 && p.stride[1] == 1 && p.stride[2] == 1;
  for (int e = threadIdx.x; e < cn * p.in_sp; e += blockDim.x) {
    int c, sp;
    p.f_insp.divmod(e, c, sp);
    int i0, r, i1, i2;
    u.f_i12.divmod(sp, i0, r);
    u.f_i2.divmod(r, i1, i2);
    float acc = 0.f;
    int t = 0;
    for (int t0 = 0; t0 < p.ker[0]; ++t0) {
      const int n0 = i0 + p.pad[0] - t0 * p.dil[0];
      const int o0 = unit ? n0 : n0 / p.stride[0];
      const bool ok0 = n0 >= 0 && (unit || n0 == o0 * p.stride[0]) && o0 < p.out[0];
      for (int t1 = 0; t1 < p.ker[1]; ++t1) {
        const int n1 = i1 + p.pad[1] - t1 * p.dil[1];
        const int o1 = unit ? n1 : n1 / p.stride[1];
        const bool ok1 = ok0 && n1 >= 0 && (unit || n1 == o1 * p.stride[1]) && o1 < p.out[1];
        for (int t2 = 0; t2 < p.ker[2]; ++t2, ++t) {
          const int n2 = i2 + p.pad[2] - t2 * p.dil[2];
          const int o2 = unit ? n2 : n2 / p.stride[2];
          if (ok1 && n2 >= 0 && (unit || n2 == o2 * p.stride[2]) && o2 < p.out[2])
            acc += dsm[((o0 * p.out[1] + o1) * p.out[2] + o2) * run + c * p.taps + t];
        }
      }
    }
    dst[e] = from_f32<T>(acc);
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
  p->f_insp = FastDivU(p->in_sp); p->f_outsp = FastDivU(p->out_sp);
  p->f_i12 = FastDivU(p->in[1] * p->in[2]); p->f_i2 = FastDivU(p->in[2]);
  p->f_o12 = FastDivU(p->out[1] * p->out[2]); p->f_o2 = FastDivU(p->out[2]);
  p->f_taps = FastDivU(p->taps);
  long long pph = 1;
  for (int i = 0; i < 3; ++i) pph *= p->in[i] + 2 * p->pad[i];
  if (pph > (1 << 22)) return false;
  p->PPh = (int)pph;
  return true;
}

// planes per block: enough work for ~4 rounds of 256 threads, bounded by shared memory and by >= ~2 blocks per SM
static int choose_pb(const DwParams& p, int plane_floats, int items_per_plane) {
  long long pb = (4 * 256 + items_per_plane - 1) / items_per_plane;
  const long long by_smem = (long long)(kDwSmemLimit / 4) / (plane_floats + p.taps);
  if (pb > by_smem) pb = by_smem;
  const long long by_grid = p.planes / (2 * 148);
  if (by_grid >= 1 && pb > by_grid) pb = by_grid;
  if (pb < 1) pb = 1;
  return (int)pb;
}

// opt in to > 48 KB of dynamic shared memory, once per (kernel, high-water mark).  Keyed by the function POINTER: kernels
// with identical signatures (fwd / dgrad, FLIP = false / true) share one template instantiation of this helper.
template <typename K>
static int dw_set_smem(K kernel, size_t bytes) {
  return ensure_smem((const void*)kernel, bytes);          // per (device, function), see common.cuh
}

template <typename T>
static int launch_fwd(DwParams p, const void* x, const void* w, void* y, cudaStream_t st) {
  p.PB = choose_pb(p, p.PPh, (p.out_sp + 3) / 4);            // ~16 outputs per thread per block
  const size_t smem = (size_t)p.PB * (p.PPh + p.taps) * 4;
  int rc = dw_set_smem(dwconv_fwd_staged<T>, smem);
  if (rc) return rc;
  const long long blocks = (p.planes + p.PB - 1) / p.PB;
  dwconv_fwd_staged<T><<<(unsigned)blocks, 256, smem, st>>>(p, (const T*)x, (const T*)w, (T*)y);
  return check_launch("dwconv_fwd_staged");
}

template <typename T>
static int launch_dgrad(DwParams p, const void* dy, const void* w, void* dx, cudaStream_t st) {
  p.PB = choose_pb(p, p.out_sp, (p.in_sp + 3) / 4);
  const size_t smem = (size_t)p.PB * (p.out_sp + p.taps) * 4;
  int rc = dw_set_smem(dwconv_dgrad_staged<T>, smem);
  if (rc) return rc;
  const long long blocks = (p.planes + p.PB - 1) / p.PB;
  dwconv_dgrad_staged<T><<<(unsigned)blocks, 256, smem, st>>>(p, (const T*)dy, (const T*)w, (T*)dx);
  return check_launch("dwconv_dgrad_staged");
}

template <typename T, int MAXT>
static int launch_wgrad_t(DwParams p, int batch, const void* x, const void* dy, float* dw, cudaStream_t st) {
  // batch chunks: ~4 blocks per SM in total
  long long chunks = (148LL * 4 + p.channels - 1) / p.channels;
  if (chunks > batch) chunks = batch;
  if (chunks < 1) chunks = 1;
  if (chunks > 65535) chunks = 65535;
  const int bpb = (int)((batch + chunks - 1) / chunks);
  chunks = (batch + bpb - 1) / bpb;
  long long pbw = (16 * 256 + p.out_sp - 1) / p.out_sp;
  const long long by_smem = (long long)((kDwSmemLimit - 4 * 8 * kMaxTaps) / 4) / (p.PPh + p.out_sp);
  if (pbw > by_smem) pbw = by_smem;
  if (pbw > bpb) pbw = bpb;
  if (pbw < 1) pbw = 1;
  const size_t smem = ((size_t)pbw * (p.PPh + p.out_sp) + 8 * p.taps) * 4;
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

// TMA-fed bf16 variants (dwconv_tma.cu)
bool dwconv3x3_tma_ok(int H, int W, const void* a, const void* b);
int dwconv3x3_tma(bool flip, int H, int W, int channels, long long planes, const void* x, const void* w, void* y, cudaStream_t st);
int dwconv3x3_wgrad_tma(int H, int W, int channels, int batch, const void* x, const void* dy, float* dw, cudaStream_t st);
static bool tma_variant_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TLT_DWCONV_TMA"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

static bool fast3_applicable(const DwParams& p) {
  return p.in[0] == 1 && p.ker[0] == 1 && p.ker[1] == 3 && p.ker[2] == 3 && p.stride[1] == 1 && p.stride[2] == 1 &&
         p.pad[1] == 1 && p.pad[2] == 1 && p.dil[1] == 1 && p.dil[2] == 1 && p.out[1] == p.in[1] && p.out[2] == p.in[2] &&
         p.in[2] >= 4;
}

static void fast3_fill(const DwParams& p, Fast3Params* q, int planes_needed_per_item) {
  q->H = p.in[1]; q->W = p.in[2];
  q->strips = (q->W + 3) / 4;
  q->RP = ((4 + 4 * q->strips + 1) + 3) & ~3;                 // 4 lead floats, interior rounded to strips, >= 1 trailing halo
  q->channels = p.channels; q->planes = p.planes;
  q->f_spp = FastDivU(q->H * q->strips); q->f_strips = FastDivU(q->strips);
  (void)planes_needed_per_item;
}

template <typename T, bool FLIP>
static int launch_fast3(const DwParams& p, const void* x, const void* w, void* y, cudaStream_t st) {
  Fast3Params q;
  fast3_fill(p, &q, 1);
  const int plane = (q.H + 2) * q.RP;
  long long pb = (8 * 256 + q.H * q.strips - 1) / (q.H * q.strips);       // ~8 strips per thread per block
  const long long by_smem = (long long)(kDwSmemLimit / 4) / (plane + 12);
  if (pb > by_smem) pb = by_smem;
  const long long by_grid = p.planes / (2 * 148);
  if (by_grid >= 1 && pb > by_grid) pb = by_grid;
  if (pb < 1) pb = 1;
  q.PB = (int)pb;
  const size_t smem = (size_t)q.PB * (plane + 12) * 4;
  const unsigned blocks = (unsigned)((p.planes + q.PB - 1) / q.PB);
  const bool a16 = (((uintptr_t)x) & 15) == 0;
  const int esz = (int)sizeof(T);
  int V = 1;
  if (a16 && q.W % (16 / esz) == 0) V = 8 / (esz / 2) / 1;               // bf16: 8 elements, fp32: 4 elements (16 bytes)
  else if (a16 && esz == 2 && q.W % 4 == 0) V = 4;                        // bf16: 8-byte loads
  // for fp32 the 16-byte vector holds 4 elements: reuse the V == 4 code path (one float4)
  int rc;
  if (esz == 4 && V == 4) {
    q.f_rowv = FastDivU(q.W / 4); q.f_rowsv = FastDivU(q.H * (q.W / 4));
    if ((rc = dw_set_smem(dwconv3x3_fast<T, 4, FLIP>, smem))) return rc;
    dwconv3x3_fast<T, 4, FLIP><<<blocks, 256, smem, st>>>(q, (const T*)x, (const T*)w, (T*)y);
  } else if (esz == 2 && V == 8) {
    q.f_rowv = FastDivU(q.W / 8); q.f_rowsv = FastDivU(q.H * (q.W / 8));
    if ((rc = dw_set_smem(dwconv3x3_fast<T, 8, FLIP>, smem))) return rc;
    dwconv3x3_fast<T, 8, FLIP><<<blocks, 256, smem, st>>>(q, (const T*)x, (const T*)w, (T*)y);
  } else if (esz == 2 && V == 4) {
    q.f_rowv = FastDivU(q.W / 4); q.f_rowsv = FastDivU(q.H * (q.W / 4));
    if ((rc = dw_set_smem(dwconv3x3_fast<T, 4, FLIP>, smem))) return rc;
    dwconv3x3_fast<T, 4, FLIP><<<blocks, 256, smem, st>>>(q, (const T*)x, (const T*)w, (T*)y);
  } else {
    q.f_rowv = FastDivU(q.W); q.f_rowsv = FastDivU(q.H * q.W);
    if ((rc = dw_set_smem(dwconv3x3_fast<T, 1, FLIP>, smem))) return rc;
    dwconv3x3_fast<T, 1, FLIP><<<blocks, 256, smem, st>>>(q, (const T*)x, (const T*)w, (T*)y);
  }
  return check_launch("dwconv3x3_fast");
}

template <typename T>
static int launch_fast3_wgrad(const DwParams& p, int batch, const void* x, const void* dy, float* dw, cudaStream_t st) {
  Fast3Params q;
  fast3_fill(p, &q, 2);
  const int plane = (q.H + 2) * q.RP;
  long long chunks = (148LL * 8 + p.channels - 1) / p.channels;
  if (chunks > batch) chunks = batch;
  if (chunks < 1) chunks = 1;
  if (chunks > 65535) chunks = 65535;
  const int bpb = (int)((batch + chunks - 1) / chunks);
  chunks = (batch + bpb - 1) / bpb;
  long long pbw = (4 * 256 + q.H * q.strips - 1) / (q.H * q.strips);
  const long long by_smem = (long long)((64 * 1024 - 4 * 72) / 4) / (2 * plane) > 0 ? (long long)((64 * 1024 - 4 * 72) / 4) / (2 * plane) : 1;
  if (pbw > by_smem) pbw = by_smem;
  if (pbw > bpb) pbw = bpb;
  if (pbw < 1) pbw = 1;
  q.PB = (int)pbw;
  const size_t smem = ((size_t)pbw * 2 * plane + 72) * 4;
  dim3 grid((unsigned)p.channels, (unsigned)chunks);
  const int esz = (int)sizeof(T);
  const bool a16 = (((uintptr_t)x) & 15) == 0 && (((uintptr_t)dy) & 15) == 0 && ((long long)q.H * q.W * esz) % 16 == 0;
  int rc;
  if (a16 && esz == 2 && q.W % 8 == 0) {
    q.f_rowv = FastDivU(q.W / 8); q.f_rowsv = FastDivU(q.H * (q.W / 8));
    if ((rc = dw_set_smem(dwconv3x3_wgrad_fast<T, 8>, smem))) return rc;
    dwconv3x3_wgrad_fast<T, 8><<<grid, 256, smem, st>>>(q, batch, bpb, (int)pbw, (const T*)x, (const T*)dy, dw);
  } else if (a16 && q.W % 4 == 0 && ((long long)q.H * q.W * esz) % 16 == 0) {
    q.f_rowv = FastDivU(q.W / 4); q.f_rowsv = FastDivU(q.H * (q.W / 4));
    if ((rc = dw_set_smem(dwconv3x3_wgrad_fast<T, 4>, smem))) return rc;
    dwconv3x3_wgrad_fast<T, 4><<<grid, 256, smem, st>>>(q, batch, bpb, (int)pbw, (const T*)x, (const T*)dy, dw);
  } else {
    q.f_rowv = FastDivU(q.W); q.f_rowsv = FastDivU(q.H * q.W);
    if ((rc = dw_set_smem(dwconv3x3_wgrad_fast<T, 1>, smem))) return rc;
    dwconv3x3_wgrad_fast<T, 1><<<grid, 256, smem, st>>>(q, batch, bpb, (int)pbw, (const T*)x, (const T*)dy, dw);
  }
  return check_launch("dwconv3x3_wgrad_fast");
}

int im2col_generic(const tlt_conv_desc* d, const void* x, void* cols, void* stream);
int col2im_generic(const tlt_conv_desc* d, const void* dcols, void* dx, void* stream);

static bool fill_unfold(const tlt_conv_desc* d, UnfoldParams* u, bool adjoint) {
  if (!fill_dw(d, &u->g)) return false;
  DwParams& p = u->g;
  if (d->batch >= (1LL << 24)) return false;
  u->batch = (int)d->batch;
  // channels per block: as many as fit ~96 KB of fp32 staging, but keep >= ~2 blocks per SM
  const long long per_ch = adjoint ? (long long)p.out_sp * p.taps : p.in_sp;
  long long cc = (96 * 1024 / 4) / per_ch;
  if (cc < 1) return false;
  if (cc > p.channels) cc = p.channels;
  while (cc > 8 && (long long)u->batch * ((p.channels + cc - 1) / cc) < 296) cc = (cc + 1) / 2;
  u->CC = (int)cc;
  u->n_cc = (p.channels + u->CC - 1) / u->CC;
  if ((long long)u->batch * u->n_cc >= (1LL << 31)) return false;
  u->f_taps = FastDivU(p.taps); u->f_k12 = FastDivU(p.ker[1] * p.ker[2]); u->f_k2 = FastDivU(p.ker[2]);
  u->f_o12 = FastDivU(p.out[1] * p.out[2]); u->f_o2 = FastDivU(p.out[2]);
  u->f_i12 = FastDivU(p.in[1] * p.in[2]); u->f_i2 = FastDivU(p.in[2]);
  u->f_run = FastDivU(u->CC * p.taps);
  return true;
}

}  // namespace tlt

using namespace tlt;

extern "C" int tlt_im2col(const tlt_conv_desc* d, const void* x, void* cols, void* stream) {
  UnfoldParams u;
  if (!fill_unfold(d, &u, false) || (d->dtype != TLT_F32 && d->dtype != TLT_BF16)) return im2col_generic(d, x, cols, stream);
  TLT_REQUIRE(x && cols, "tlt_im2col: null pointer");
  const size_t smem = (size_t)u.CC * u.g.in_sp * 4;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)(u.batch * u.n_cc);
  int rc;
  if (d->dtype == TLT_BF16) {
    if ((rc = dw_set_smem(im2col_staged<__nv_bfloat16>, smem))) return rc;
    im2col_staged<__nv_bfloat16><<<blocks, 256, smem, st>>>(u, (const __nv_bfloat16*)x, (__nv_bfloat16*)cols);
  } else {
    if ((rc = dw_set_smem(im2col_staged<float>, smem))) return rc;
    im2col_staged<float><<<blocks, 256, smem, st>>>(u, (const float*)x, (float*)cols);
  }
  return check_launch("im2col_staged");
}

extern "C" int tlt_col2im(const tlt_conv_desc* d, const void* dcols, void* dx, void* stream) {
  UnfoldParams u;
  if (!fill_unfold(d, &u, true) || (d->dtype != TLT_F32 && d->dtype != TLT_BF16)) return col2im_generic(d, dcols, dx, stream);
  TLT_REQUIRE(dcols && dx, "tlt_col2im: null pointer");
  // the adjoint stages the (out_sp x CC*taps) slab: u.f_run must divide by the ACTUAL run of a full chunk
  const size_t smem = (size_t)u.g.out_sp * u.CC * u.g.taps * 4;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)(u.batch * u.n_cc);
  int rc;
  if (d->dtype == TLT_BF16) {
    if ((rc = dw_set_smem(col2im_staged<__nv_bfloat16>, smem))) return rc;
    col2im_staged<__nv_bfloat16><<<blocks, 256, smem, st>>>(u, (const __nv_bfloat16*)dcols, (__nv_bfloat16*)dx);
  } else {
    if ((rc = dw_set_smem(col2im_staged<float>, smem))) return rc;
    col2im_staged<float><<<blocks, 256, smem, st>>>(u, (const float*)dcols, (float*)dx);
  }
  return check_launch("col2im_staged");
}

extern "C" int tlt_dwconv_fwd(const tlt_conv_desc* d, const void* x, const void* w, void* y, void* stream) {
  DwParams p;
  if (!fill_dw(d, &p) || (size_t)(p.PPh + p.taps) * 4 > kDwSmemLimit || (d->dtype != TLT_F32 && d->dtype != TLT_BF16))
    return dwconv_fwd_generic(d, x, w, y, stream);
  TLT_REQUIRE(x && w && y, "tlt_dwconv_fwd: null pointer");
  for (int i = 0; i < d->ndim; ++i) {
    long long expect = ((long long)d->in_size[i] + 2LL * d->padding[i] - (long long)d->dilation[i] * (d->kernel[i] - 1) - 1) / d->stride[i] + 1;
    TLT_REQUIRE(expect == d->out_size[i], "conv: out_size[%d]=%lld inconsistent (expected %lld)", i, (long long)d->out_size[i], expect);
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (fast3_applicable(p)) {
    if (d->dtype == TLT_BF16 && tma_variant_enabled() && dwconv3x3_tma_ok(p.in[1], p.in[2], x, y))
      return dwconv3x3_tma(false, p.in[1], p.in[2], p.channels, p.planes, x, w, y, st);
    if (d->dtype == TLT_BF16) return launch_fast3<__nv_bfloat16, false>(p, x, w, y, st);
    return launch_fast3<float, false>(p, x, w, y, st);
  }
  if (d->dtype == TLT_BF16) return launch_fwd<__nv_bfloat16>(p, x, w, y, st);
  return launch_fwd<float>(p, x, w, y, st);
}

extern "C" int tlt_dwconv_dgrad(const tlt_conv_desc* d, const void* dy, const void* w, void* dx, void* stream) {
  DwParams p;
  if (!fill_dw(d, &p) || (size_t)(p.out_sp + p.taps) * 4 > kDwSmemLimit || (d->dtype != TLT_F32 && d->dtype != TLT_BF16))
    return dwconv_dgrad_generic(d, dy, w, dx, stream);
  TLT_REQUIRE(dy && w && dx, "tlt_dwconv_dgrad: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (fast3_applicable(p)) {                       // stride 1, pad 1: the adjoint is the same correlation with reversed taps
    if (d->dtype == TLT_BF16 && tma_variant_enabled() && dwconv3x3_tma_ok(p.in[1], p.in[2], dy, dx))
      return dwconv3x3_tma(true, p.in[1], p.in[2], p.channels, p.planes, dy, w, dx, st);
    if (d->dtype == TLT_BF16) return launch_fast3<__nv_bfloat16, true>(p, dy, w, dx, st);
    return launch_fast3<float, true>(p, dy, w, dx, st);
  }
  if (d->dtype == TLT_BF16) return launch_dgrad<__nv_bfloat16>(p, dy, w, dx, st);
  return launch_dgrad<float>(p, dy, w, dx, st);
}

extern "C" int tlt_dwconv_wgrad(const tlt_conv_desc* d, const void* x, const void* dy, float* dw, void* stream) {
  DwParams p;
  if (!fill_dw(d, &p) || (size_t)(p.PPh + p.out_sp + 8 * kMaxTaps) * 4 > kDwSmemLimit || d->batch >= (1LL << 31) ||
      (d->dtype != TLT_F32 && d->dtype != TLT_BF16))
    return dwconv_wgrad_generic(d, x, dy, dw, stream);
  TLT_REQUIRE(x && dy && dw, "tlt_dwconv_wgrad: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (fast3_applicable(p)) {
    if (d->dtype == TLT_BF16 && tma_variant_enabled() && dwconv3x3_tma_ok(p.in[1], p.in[2], x, dy))
      return dwconv3x3_wgrad_tma(p.in[1], p.in[2], p.channels, (int)d->batch, x, dy, dw, st);
    if (d->dtype == TLT_BF16) return launch_fast3_wgrad<__nv_bfloat16>(p, (int)d->batch, x, dy, dw, st);
    return launch_fast3_wgrad<float>(p, (int)d->batch, x, dy, dw, st);
  }
  if (d->dtype == TLT_BF16) return launch_wgrad<__nv_bfloat16>(p, (int)d->batch, x, dy, dw, st);
  return launch_wgrad<float>(p, (int)d->batch, x, dy, dw, st);
}
