// Depthwise convolution on channels-last rows: the middle stage of the CP convolution when the feature map is small and
// odd-sized (14 x 14 = 196 and 7 x 7 = 49 positions are not multiples of 8, so the NCHW tensor-core 1x1 kernel cannot take
// the layers around it; BASELINE config 3's layer3 / layer4 blocks ran on the generic kernels at 3-5 % of their roofline).
// On rows every thread owns 8 consecutive channels of one position: each tap is one 16-byte load of the activation and one
// of the tap weights (taps-major, so neighbouring threads read neighbouring chunks), i.e. the access pattern is coalesced
// for any spatial size, stride, padding or dilation, 1 to 3 spatial dims.
//
//   tlt_dwconv_rows_fwd   : y[(b, o), c] = sum_t z[(b, o*s - p + t*d), c] * w[t, c]
//   tlt_dwconv_rows_dgrad : dz[(b, i), c] = sum over (o, t) hitting i of g[(b, o), c] * w[t, c]          (gather form)
//   tlt_dwconv_rows_wgrad : dw[t, c] += sum_{b, o} z[(b, o*s - p + t*d), c] * g[(b, o), c]                  (fp32, caller zeroes)
//
// Reference call sites: the depthwise F.conv1d chain of cp_conv (tltorch/functional/convolution.py:268-271, merged into one
// N-D depthwise conv as cp_conv_mobilenet does at :321-332) and its autograd.
#include "common.cuh"

namespace tlt {

struct DwRowsGeom {
  int ndim;
  long long batch;
  int cpr;                                   // 16-byte chunks per row (channels / 8)
  int in_size[3], out_size[3], kernel[3], stride[3], padding[3], dilation[3];
  int in_sp, out_sp, taps;
};

__device__ __forceinline__ void dw_tap(const DwRowsGeom& g, int tap, int (&t)[3]) {
  const int k1 = g.ndim > 1 ? g.kernel[g.ndim - 1] : 1, k2 = g.ndim > 2 ? g.kernel[g.ndim - 2] : 1;
  if (g.ndim == 1) { t[0] = tap; }
  else if (g.ndim == 2) { t[0] = tap / k1; t[1] = tap - t[0] * k1; }
  else { t[2] = tap % k1; const int q = tap / k1; t[1] = q % k2; t[0] = q / k2; }
}

__device__ __forceinline__ void fma8(float (&acc)[8], const uint4& a, const uint4& b) {
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[i]));
    const float2 fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&bw[i]));
    acc[2 * i] = fmaf(fa.x, fb.x, acc[2 * i]);
    acc[2 * i + 1] = fmaf(fa.y, fb.y, acc[2 * i + 1]);
  }
}

__device__ __forceinline__ uint4 pack8(const float (&acc)[8]) {
  uint4 out;
  __nv_bfloat162 p;
  p = __floats2bfloat162_rn(acc[0], acc[1]); out.x = *reinterpret_cast<uint32_t*>(&p);
  p = __floats2bfloat162_rn(acc[2], acc[3]); out.y = *reinterpret_cast<uint32_t*>(&p);
  p = __floats2bfloat162_rn(acc[4], acc[5]); out.z = *reinterpret_cast<uint32_t*>(&p);
  p = __floats2bfloat162_rn(acc[6], acc[7]); out.w = *reinterpret_cast<uint32_t*>(&p);
  return out;
}

// DGRAD = false: forward (thread per output position);  true: data gradient (thread per input position, gather)
template <bool DGRAD>
__global__ void __launch_bounds__(256) dwconv_rows_kernel(const DwRowsGeom g, const uint4* __restrict__ src, const uint4* __restrict__ w,
                                                          uint4* __restrict__ dst) {
  const int n_pos = DGRAD ? g.in_sp : g.out_sp, n_src = DGRAD ? g.out_sp : g.in_sp;
  const long long total = g.batch * n_pos * g.cpr;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(e % g.cpr);
    const long long row = e / g.cpr;
    int pos = (int)(row % n_pos);
    const long long b = row / n_pos;
    int idx[3] = {0, 0, 0};
    for (int d = g.ndim - 1; d >= 0; --d) {
      const int ext = DGRAD ? g.in_size[d] : g.out_size[d];
      idx[d] = pos % ext;
      pos /= ext;
    }
    const uint4* sb = src + b * n_src * g.cpr + ch;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int tap = 0; tap < g.taps; ++tap) {
      int t[3];
      dw_tap(g, tap, t);
      int off = 0;
      bool ok = true;
      for (int d = 0; d < g.ndim; ++d) {
        int j;
        if (!DGRAD) {
          j = idx[d] * g.stride[d] - g.padding[d] + t[d] * g.dilation[d];
          ok = ok && j >= 0 && j < g.in_size[d];
          off = off * g.in_size[d] + j;
        } else {
          const int num = idx[d] + g.padding[d] - t[d] * g.dilation[d];
          j = g.stride[d] == 1 ? num : num / g.stride[d];
          ok = ok && num >= 0 && j * g.stride[d] == num && j < g.out_size[d];
          off = off * g.out_size[d] + j;
        }
      }
      if (!ok) continue;
      fma8(acc, sb[(long long)off * g.cpr], w[tap * g.cpr + ch]);
    }
    dst[e] = pack8(acc);
  }
}

// block = 256 threads = (cpr_blk chunks) x (256 / cpr_blk position lanes); each block owns a slab of output positions.
// A thread keeps the 8 channel sums of every tap in registers (TAPS_MAX x 8 fp32) over its positions; position lanes are
// then summed through shared memory and the block issues one atomicAdd per (tap, channel).
constexpr int kDwTapsMax = 9;
__global__ void __launch_bounds__(256) dwconv_rows_wgrad_kernel(const DwRowsGeom g, const uint4* __restrict__ z, const uint4* __restrict__ gy,
                                                                float* __restrict__ dw, int pos_per_block, int chunks_x) {
  extern __shared__ float wg_smem[];                               // [lanes][taps][8] per chunk column -> reduced in place
  const int lanes = 256 / chunks_x;
  const int cx = threadIdx.x % chunks_x, lane = threadIdx.x / chunks_x;
  const long long rows_total = g.batch * g.out_sp;
  const long long row0 = (long long)blockIdx.x * pos_per_block;
  const long long row1 = row0 + pos_per_block < rows_total ? row0 + pos_per_block : rows_total;
  for (int c0 = 0; c0 < g.cpr; c0 += chunks_x) {
    const int ch = c0 + cx;
    float acc[kDwTapsMax][8];
#pragma unroll
    for (int t = 0; t < kDwTapsMax; ++t)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[t][i] = 0.f;
    if (ch < g.cpr && lane < lanes) {
      for (long long row = row0 + lane; row < row1; row += lanes) {
        int pos = (int)(row % g.out_sp);
        const long long b = row / g.out_sp;
        int base[3] = {0, 0, 0};
        for (int d = g.ndim - 1; d >= 0; --d) {
          base[d] = (pos % g.out_size[d]) * g.stride[d] - g.padding[d];
          pos /= g.out_size[d];
        }
        const uint4 gv = gy[row * g.cpr + ch];
        const uint4* zb = z + b * g.in_sp * g.cpr + ch;
#pragma unroll
        for (int tap = 0; tap < kDwTapsMax; ++tap) {
          if (tap < g.taps) {
            int t[3];
            dw_tap(g, tap, t);
            int off = 0;
            bool ok = true;
            for (int d = 0; d < g.ndim; ++d) {
              const int j = base[d] + t[d] * g.dilation[d];
              ok = ok && j >= 0 && j < g.in_size[d];
              off = off * g.in_size[d] + j;
            }
            if (ok) fma8(acc[tap], zb[(long long)off * g.cpr], gv);
          }
        }
      }
    }
    // reduce the position lanes: wg_smem[lane][tap][cx][8]
    __syncthreads();
    if (lane < lanes) {
#pragma unroll
      for (int tap = 0; tap < kDwTapsMax; ++tap)
        if (tap < g.taps) {
#pragma unroll
          for (int i = 0; i < 8; ++i) wg_smem[((lane * g.taps + tap) * chunks_x + cx) * 8 + i] = acc[tap][i];
        }
    }
    __syncthreads();
    const int n_out = g.taps * chunks_x * 8;
    for (int o = threadIdx.x; o < n_out; o += 256) {
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += wg_smem[l * n_out + o];
      const int tap = o / (chunks_x * 8), rem = o - tap * chunks_x * 8;
      const int c = (c0 + rem / 8) * 8 + (rem & 7);
      if (c < g.cpr * 8 && s != 0.f) atomicAdd(dw + (long long)tap * g.cpr * 8 + c, s);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// 2-D specialisations (the CP convolutions of every ResNet block): the position is decoded once per thread, the tap loops
// are two plain nested loops with the row test hoisted, so a tap costs two 16-byte loads and 8 FMAs plus ~6 integer ops
// (the generic N-D kernels above spend ~100 instructions per tap on the index arithmetic: 96-103 us per pass over a 29 MB
// tensor at config 3's layer3 shapes).
// ---------------------------------------------------------------------------------------------------------------------
// v2: a thread owns 8 channels of a whole output ROW (b, y, :).  The 9 x 8 tap weights are unpacked to fp32 registers once,
// the three source-row base pointers once per output row, so an output costs 9 16-byte loads + their unpacking + 72 FMAs
// and a handful of integer ops (v1 decoded the position per output with 64-bit arithmetic: 700 instructions per output,
// issue-bound at 54-58 us per pass; profiles/r01_dwconv_rows.md).
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

template <bool DGRAD>
__global__ void __launch_bounds__(128) dwconv_rows2d_kernel(const DwRowsGeom g, const uint4* __restrict__ src, const uint4* __restrict__ w,
                                                            uint4* __restrict__ dst, int segs, int seg_len) {
  const int H = DGRAD ? g.in_size[0] : g.out_size[0], W = DGRAD ? g.in_size[1] : g.out_size[1];      // extents of dst
  const int SH = DGRAD ? g.out_size[0] : g.in_size[0], SW = DGRAD ? g.out_size[1] : g.in_size[1];    // extents of src
  const int KH = g.kernel[0], KW = g.kernel[1], cpr = g.cpr;
  const int sh = g.stride[0], sw = g.stride[1], ph = g.padding[0], pw = g.padding[1], dh = g.dilation[0], dwl = g.dilation[1];
  const long long total = g.batch * H * segs * cpr;                  // a row is cut into `segs` segments when rows alone
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {   // cannot fill the GPU
    const int ch = (int)(e % cpr);
    long long row = e / cpr;
    const int seg = (int)(row % segs);
    row /= segs;
    const int y = (int)(row % H);
    const long long b = row / H;
    const int x_begin = seg * seg_len, x_end = min(W, x_begin + seg_len);
    float wk[3][3][8];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        if (kh < KH && kw < KW) unpack8(w[(kh * KW + kw) * cpr + ch], wk[kh][kw]);
      }
    const uint4* srow[3];
    bool rok[3];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      int sy = 0;
      bool ok = kh < KH;
      if (!DGRAD) {
        sy = y * sh - ph + kh * dh;
        ok = ok && sy >= 0 && sy < SH;
      } else {
        const int num = y + ph - kh * dh;
        sy = sh == 1 ? num : num / sh;
        ok = ok && num >= 0 && sy * sh == num && sy < SH;
      }
      rok[kh] = ok;
      srow[kh] = src + ((b * SH + (ok ? sy : 0)) * SW) * cpr + ch;
    }
    uint4* drow = dst + ((b * H + y) * (long long)W) * cpr + ch;
    for (int x = x_begin; x < x_end; ++x) {
      // all nine loads are issued before any arithmetic (clamped addresses, invalid taps zeroed afterwards): with the
      // loads behind per-tap branches a thread had ONE 16-byte load in flight and the kernel sat on the long scoreboard
      // at 14 % of the DRAM bandwidth (profiles/r01_dwconv_rows.md)
      uint4 raw[3][3];
      bool cok[3];
      int sxs[3];
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        int sx = 0;
        bool ok = kw < KW;
        if (!DGRAD) {
          sx = x * sw - pw + kw * dwl;
          ok = ok && sx >= 0 && sx < SW;
        } else {
          const int num = x + pw - kw * dwl;
          sx = sw == 1 ? num : num / sw;
          ok = ok && num >= 0 && sx * sw == num && sx < SW;
        }
        cok[kw] = ok;
        sxs[kw] = ok ? sx : 0;
      }
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) raw[kh][kw] = srow[kh][sxs[kw] * cpr];
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          if (kh < 3 && rok[kh] && cok[kw]) {
            float v[8];
            unpack8(raw[kh][kw], v);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(v[i], wk[kh][kw][i], acc[i]);
          }
        }
      drow[x * cpr] = pack8(acc);
    }
  }
}

// generic 2-D kernel sizes (> 3 in a dimension): thread per output, taps as runtime loops
template <bool DGRAD>
__global__ void __launch_bounds__(256) dwconv_rows2d_big_kernel(const DwRowsGeom g, const uint4* __restrict__ src, const uint4* __restrict__ w,
                                                                uint4* __restrict__ dst) {
  const int H = DGRAD ? g.in_size[0] : g.out_size[0], W = DGRAD ? g.in_size[1] : g.out_size[1];
  const int SH = DGRAD ? g.out_size[0] : g.in_size[0], SW = DGRAD ? g.out_size[1] : g.in_size[1];
  const int KH = g.kernel[0], KW = g.kernel[1], cpr = g.cpr;
  const int sh = g.stride[0], sw = g.stride[1], ph = g.padding[0], pw = g.padding[1], dh = g.dilation[0], dwl = g.dilation[1];
  const long long total = g.batch * H * W * cpr;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(e % cpr);
    long long row = e / cpr;
    const int x = (int)(row % W);
    row /= W;
    const int y = (int)(row % H);
    const long long b = row / H;
    const uint4* sb = src + b * SH * SW * cpr + ch;
    const uint4* wp = w + ch;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int kh = 0; kh < KH; ++kh) {
      int sy;
      if (!DGRAD) {
        sy = y * sh - ph + kh * dh;
        if (sy < 0 || sy >= SH) continue;
      } else {
        const int num = y + ph - kh * dh;
        sy = sh == 1 ? num : num / sh;
        if (num < 0 || sy * sh != num || sy >= SH) continue;
      }
      const uint4* srow = sb + (long long)sy * SW * cpr;
      for (int kw = 0; kw < KW; ++kw) {
        int sx;
        if (!DGRAD) {
          sx = x * sw - pw + kw * dwl;
          if (sx < 0 || sx >= SW) continue;
        } else {
          const int num = x + pw - kw * dwl;
          sx = sw == 1 ? num : num / sw;
          if (num < 0 || sx * sw != num || sx >= SW) continue;
        }
        fma8(acc, srow[(long long)sx * cpr], wp[(kh * KW + kw) * cpr]);
      }
    }
    dst[e] = pack8(acc);
  }
}

// 2-D weight gradient: same block organisation as dwconv_rows_wgrad_kernel, position decoded with two divisions, taps as two
// nested loops over a register array indexed at compile time (KH, KW <= 3).
__global__ void __launch_bounds__(256) dwconv_rows2d_wgrad_kernel(const DwRowsGeom g, const uint4* __restrict__ z, const uint4* __restrict__ gy,
                                                                  float* __restrict__ dw, int pos_per_block, int chunks_x) {
  extern __shared__ float wg_smem[];
  const int lanes = 256 / chunks_x;
  const int cx = threadIdx.x % chunks_x, lane = threadIdx.x / chunks_x;
  const int OH = g.out_size[0], OW = g.out_size[1], IH = g.in_size[0], IW = g.in_size[1], cpr = g.cpr;
  const int KH = g.kernel[0], KW = g.kernel[1];
  const int sh = g.stride[0], sw = g.stride[1], ph = g.padding[0], pw = g.padding[1], dh = g.dilation[0], dwl = g.dilation[1];
  const long long rows_total = g.batch * g.out_sp;
  const long long row0 = (long long)blockIdx.x * pos_per_block;
  const long long row1 = row0 + pos_per_block < rows_total ? row0 + pos_per_block : rows_total;
  for (int c0 = 0; c0 < cpr; c0 += chunks_x) {
    const int ch = c0 + cx;
    float acc[3][3][8];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[a][c][i] = 0.f;
    if (ch < cpr && lane < lanes) {
      for (long long row = row0 + lane; row < row1; row += lanes) {
        const int ox = (int)(row % OW);
        const long long r2 = row / OW;
        const int oy = (int)(r2 % OH);
        const long long b = r2 / OH;
        const uint4 gv = gy[row * cpr + ch];
        const uint4* zb = z + b * IH * IW * cpr + ch;
        const int y0 = oy * sh - ph, x0 = ox * sw - pw;
        uint4 raw[3][3];
        bool okh[3], okw[3];
        int iys[3], ixs[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int iy = y0 + k * dh, ix = x0 + k * dwl;
          okh[k] = k < KH && iy >= 0 && iy < IH;
          okw[k] = k < KW && ix >= 0 && ix < IW;
          iys[k] = okh[k] ? iy : 0;
          ixs[k] = okw[k] ? ix : 0;
        }
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) raw[kh][kw] = zb[((long long)iys[kh] * IW + ixs[kw]) * cpr];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
            if (okh[kh] && okw[kw]) fma8(acc[kh][kw], raw[kh][kw], gv);
      }
    }
    __syncthreads();
    if (lane < lanes) {
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
          if (kh < KH && kw < KW) {
            const int tap = kh * KW + kw;
#pragma unroll
            for (int i = 0; i < 8; ++i) wg_smem[((lane * g.taps + tap) * chunks_x + cx) * 8 + i] = acc[kh][kw][i];
          }
    }
    __syncthreads();
    const int n_out = g.taps * chunks_x * 8;
    for (int o = threadIdx.x; o < n_out; o += 256) {
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += wg_smem[l * n_out + o];
      const int tap = o / (chunks_x * 8), rem = o - tap * chunks_x * 8;
      const int c = (c0 + rem / 8) * 8 + (rem & 7);
      if (c < cpr * 8 && s != 0.f) atomicAdd(dw + (long long)tap * cpr * 8 + c, s);
    }
  }
}

static int fill_dw(const tlt_conv_desc* d, DwRowsGeom* g) {
  TLT_REQUIRE(d && d->dtype == TLT_BF16, "dwconv_rows: bf16 only");
  TLT_REQUIRE(d->ndim >= 1 && d->ndim <= 3, "dwconv_rows: 1..3 spatial dims");
  TLT_REQUIRE(d->batch >= 1 && d->channels >= 8 && d->channels % 8 == 0, "dwconv_rows: channels per row must be a multiple of 8");
  g->ndim = d->ndim; g->batch = d->batch; g->cpr = (int)(d->channels >> 3);
  long long in_sp = 1, out_sp = 1, taps = 1;
  for (int i = 0; i < d->ndim; ++i) {
    TLT_REQUIRE(d->in_size[i] >= 1 && d->out_size[i] >= 1 && d->kernel[i] >= 1 && d->stride[i] >= 1 && d->dilation[i] >= 1 && d->padding[i] >= 0,
                "dwconv_rows: bad geometry");
    g->in_size[i] = (int)d->in_size[i]; g->out_size[i] = (int)d->out_size[i];
    g->kernel[i] = d->kernel[i]; g->stride[i] = d->stride[i]; g->padding[i] = d->padding[i]; g->dilation[i] = d->dilation[i];
    in_sp *= d->in_size[i]; out_sp *= d->out_size[i]; taps *= d->kernel[i];
  }
  TLT_REQUIRE(in_sp < (1LL << 30) && out_sp < (1LL << 30) && taps <= 64, "dwconv_rows: extents out of range");
  g->in_sp = (int)in_sp; g->out_sp = (int)out_sp; g->taps = (int)taps;
  return TLT_OK;
}


// ---------------------------------------------------------------------------------------------------------------------
// Shared-memory tiled 3x3 kernels (padding 1, dilation 1, stride 1 or 2): the rows-route layers of BASELINE config 3.
// The register-window kernels above re-fetch every source chunk for each of the taps that touch it through L1 / L2 with ~12
// resident warps per SM (146 registers) and sit on the long scoreboard: 50 / 54 / 77 us per pass over a 29 MB tensor (ncu:
// warps active 17 %, issue active 40 %).  Here a block stages the source rows of a strip of destination rows -- CW chunks
// (CW x 8 channels) of every position, one 16-byte cp.async per chunk, all in flight at once -- and the nine taps read
// shared memory.  A staged position is CW + 1 chunks wide (odd pitch: conflict-free 16-byte reads along positions and along
// chunks).
//   forward / data gradient: thread = (chunk c = tid % CW, position lane tid / CW); 192-byte coalesced runs on both sides
//   weight gradient: warp = chunk, lane = position; 9 x 8 tap sums per thread in registers across ALL tiles of the block,
//   one shuffle reduction and one atomicAdd per (tap, channel) per block at the end
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ uint4 lds128_(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// The nine taps themselves run on the tensor cores: a depthwise tap is a product with a DIAGONAL channel matrix, so for a
// tile of 16 destination positions x 8 channels  D[p][c] = sum_t A_t[p][c'] diag(w_t)[c'][c]  is nine m16n8k8 HMMAs whose A
// operand is gathered by ldmatrix straight from the staged rows (row addresses are free per lane: the tap shift, the stride
// and the zero padding -- a shared all-zero row -- cost nothing), and whose B operand is one register per lane from a
// precomputed table.  7/8 of the multiplies hit zeros, which is irrelevant next to what it replaces: the FFMA version spent
// ~340 thread instructions per 8-channel output (unpack + FMA + predicates; 19 M warp instructions per pass, issue-bound at
// 41 us), this one ~3 warp instructions per output.  The weight gradient is the transposed product
// D[c'][c] = sum_p Z_t[p][c'] G[p][c] (m16n8k16, both operands via ldmatrix.trans), of which only the diagonal is kept.
__device__ __forceinline__ void ldsm2_(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm2t_(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t_(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma1688_(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void mma16816_(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t lds32_(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

constexpr int kDwtThreads = 192;

template <bool DGRAD, int S, int CW>
__global__ void __launch_bounds__(kDwtThreads) dwrows_tile_kernel(const DwRowsGeom g, const uint4* __restrict__ src,
                                                                  const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ dst, int TR,
                                                                  int nstrips, long long tiles) {
  constexpr int NT = kDwtThreads, NW = NT / 32, PITCH = (CW + 1) * 16, HC = CW / 2;
  extern __shared__ __align__(16) unsigned char dwt_smem[];
  const uint32_t zero = (uint32_t)__cvta_generic_to_shared(dwt_smem);          // PITCH bytes of zeros
  const uint32_t diag = zero + PITCH;                                          // [9][CW][32] b32
  const uint32_t tile = diag + 9 * CW * 128;
  const int DH = DGRAD ? g.in_size[0] : g.out_size[0], DW = DGRAD ? g.in_size[1] : g.out_size[1];
  const int SH = DGRAD ? g.out_size[0] : g.in_size[0], SWd = DGRAD ? g.out_size[1] : g.in_size[1];
  const int cpr = g.cpr, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tq = lane & 3;
  const int cbase = blockIdx.x * CW;
  for (int e = tid; e < PITCH / 4; e += NT) reinterpret_cast<uint32_t*>(dwt_smem)[e] = 0u;
  // tap table: one 16-byte load of the 8 channel weights per (tap, chunk), expanded to the 32 per-lane B registers
  for (int e = tid; e < 9 * CW; e += NT) {
    const int cc = e % CW, t = e / CW;
    const uint4 q = *reinterpret_cast<const uint4*>(w + ((size_t)t * cpr + cbase + cc) * 8);
    const uint32_t w8[8] = {q.x & 0xffffu, q.x >> 16, q.y & 0xffffu, q.y >> 16, q.z & 0xffffu, q.z >> 16, q.w & 0xffffu, q.w >> 16};
    uint32_t* dst = reinterpret_cast<uint32_t*>(dwt_smem + PITCH) + e * 32;
#pragma unroll
    for (int l = 0; l < 32; ++l) {
      const int gg = l >> 2, tt = l & 3;
      dst[l] = gg == 2 * tt ? w8[gg] : (gg == 2 * tt + 1 ? (w8[gg] << 16) : 0u);
    }
  }
  for (long long tix = blockIdx.y; tix < tiles; tix += gridDim.y) {
    const long long b = tix / nstrips;
    const int y0 = (int)(tix % nstrips) * TR, y1 = min(DH, y0 + TR);
    int lo, hi;
    if (!DGRAD) { lo = y0 * S - 1; hi = (y1 - 1) * S + 1; }
    else { lo = y0 - 1 <= 0 ? 0 : (y0 - 1 + S - 1) / S; hi = y1 / S; }
    lo = max(lo, 0);
    hi = min(hi, SH - 1);
    const int npos = (hi - lo + 1) * SWd;
    const uint4* sb = src + ((b * SH + lo) * (long long)SWd) * cpr + cbase;
    for (int idx = tid; idx < npos * CW; idx += NT) {
      const int cc = idx % CW, pos = idx / CW;
      cp_async16(tile + pos * PITCH + cc * 16, sb + (long long)pos * cpr + cc);
    }
    cp_async_wait();
    __syncthreads();
    const int nd = (y1 - y0) * DW, nr = y1 - y0;
    // Tiles of 16 destination positions.  Strided data gradient: a destination position only receives the taps whose parity
    // matches it (1, 2, 2 or 4 of the 9), so positions are grouped by (row parity, column parity) class -- every tile is
    // parity-uniform and skips the other taps warp-uniformly instead of multiplying zero rows (2.25 instead of 9 HMMAs per chunk).
    constexpr bool PAR = DGRAD && S == 2;
    int cls_tiles[4] = {(nd + 15) / 16, 0, 0, 0}, cls_nr[4] = {nr, 0, 0, 0}, cls_nc[4] = {DW, 0, 0, 0}, cls_fy[4] = {0, 0, 0, 0};
    int total_tiles = cls_tiles[0];
    if (PAR) {
      total_tiles = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int a = k >> 1, bq = k & 1;
        const int fy = (a - (y0 & 1)) & 1;
        cls_fy[k] = fy;
        cls_nr[k] = nr > fy ? (nr - fy + 1) / 2 : 0;
        cls_nc[k] = (DW - bq + 1) / 2;
        cls_tiles[k] = (cls_nr[k] * cls_nc[k] + 15) / 16;
        total_tiles += cls_tiles[k];
      }
    }
    __nv_bfloat16* db = dst + ((b * DH + y0) * (long long)DW) * cpr * 8 + (size_t)cbase * 8;
    for (int u = warp; u < total_tiles * 2; u += NW) {
      int pt = u >> 1;
      const int half = u & 1;
      int k = 0;
      if (PAR) {
        while (k < 3 && pt >= cls_tiles[k]) { pt -= cls_tiles[k]; ++k; }
      }
      const int ca = k >> 1, cb = k & 1;                              // class parities (PAR only)
      const int ncx = PAR ? cls_nc[k] : DW, count = PAR ? cls_nr[k] * cls_nc[k] : nd, fy = PAR ? cls_fy[k] : 0;
      // tile row r -> destination (local row, column)
      auto locate = [&](int r, int& yy, int& x) {
        const int q = pt * 16 + r;
        const int ry = q / ncx, cx = q - ry * ncx;
        if (PAR) { yy = fy + 2 * ry; x = cb + 2 * cx; }
        else { yy = ry; x = cx; }
        return q < count;
      };
      int yy, x;
      const bool okp = locate(lane & 15, yy, x);
      const int y = y0 + yy;
      uint32_t addr[9];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        int sy;
        bool oky;
        if (!DGRAD) { sy = y * S - 1 + kh; oky = sy >= 0 && sy < SH; }
        else { const int ny = y + 1 - kh; sy = ny / S; oky = ny >= 0 && (S == 1 || (ny % S) == 0) && sy < SH; }
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          int sx;
          bool okx;
          if (!DGRAD) { sx = x * S - 1 + kw; okx = sx >= 0 && sx < SWd; }
          else { const int nx = x + 1 - kw; sx = nx / S; okx = nx >= 0 && (S == 1 || (nx % S) == 0) && sx < SWd; }
          addr[kh * 3 + kw] = (okp && oky && okx) ? tile + ((sy - lo) * SWd + sx) * PITCH : zero;
        }
      }
      int yl, xl, yh, xh;
      const bool ok_lo = locate(gid, yl, xl), ok_hi = locate(gid + 8, yh, xh);
      const size_t off_lo = (size_t)(yl * DW + xl) * cpr * 8 + 2 * tq, off_hi = (size_t)(yh * DW + xh) * cpr * 8 + 2 * tq;
#pragma unroll 2
      for (int cc = 0; cc < HC; ++cc) {
        const int c = half * HC + cc;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          if (PAR && ((ca + 1 - kh) & 1)) continue;                   // warp-uniform
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            if (PAR && ((cb + 1 - kw) & 1)) continue;
            const int t = kh * 3 + kw;
            uint32_t a0, a1;
            ldsm2_(a0, a1, addr[t] + c * 16);
            mma1688_(acc, a0, a1, lds32_(diag + ((t * CW + c) * 32 + lane) * 4));
          }
        }
        if (ok_lo) *reinterpret_cast<__nv_bfloat162*>(db + off_lo + c * 8) = __floats2bfloat162_rn(acc[0], acc[1]);
        if (ok_hi) *reinterpret_cast<__nv_bfloat162*>(db + off_hi + c * 8) = __floats2bfloat162_rn(acc[2], acc[3]);
      }
    }
    __syncthreads();
  }
}

template <int S, int CW>
__global__ void __launch_bounds__(16 * CW) dwrows_tile_wgrad_kernel(const DwRowsGeom g, const uint4* __restrict__ z, const uint4* __restrict__ gy,
                                                                    float* __restrict__ dw, int TR, int nstrips, long long tiles, int zmax) {
  constexpr int NT = 16 * CW, PITCH = (CW + 1) * 16;                   // one warp per pair of chunks (16 channels)
  extern __shared__ __align__(16) unsigned char dwt_smem[];
  const uint32_t zero = (uint32_t)__cvta_generic_to_shared(dwt_smem);
  const uint32_t ztile = zero + PITCH;
  const uint32_t gtile = ztile + (uint32_t)zmax * PITCH;
  const int IH = g.in_size[0], IW = g.in_size[1], OH = g.out_size[0], OW = g.out_size[1];
  const int cpr = g.cpr, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tq = lane & 3;
  const int cbase = blockIdx.x * CW;
  for (int e = tid; e < PITCH / 4; e += NT) reinterpret_cast<uint32_t*>(dwt_smem)[e] = 0u;
  float acc[9][2][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[t][h][i] = 0.f;
  const int ia = (lane & 7) + ((lane >> 4) << 3);                      // A operand: lane -> position of the k16 step
  const uint32_t ca = (uint32_t)(2 * warp + ((lane >> 3) & 1)) * 16;   //            and chunk of the pair
  const int ib = lane & 15;                                            // B operand: lanes 0..15 -> positions
  for (long long tix = blockIdx.y; tix < tiles; tix += gridDim.y) {
    const long long b = tix / nstrips;
    const int y0 = (int)(tix % nstrips) * TR, y1 = min(OH, y0 + TR);
    const int lo = max(y0 * S - 1, 0), hi = min((y1 - 1) * S + 1, IH - 1);
    const int nz = (hi - lo + 1) * IW, ng = (y1 - y0) * OW;
    const uint4* zb = z + ((b * IH + lo) * (long long)IW) * cpr + cbase;
    const uint4* gb = gy + ((b * OH + y0) * (long long)OW) * cpr + cbase;
    for (int idx = tid; idx < nz * CW; idx += NT) {
      const int cc = idx % CW, pos = idx / CW;
      cp_async16(ztile + pos * PITCH + cc * 16, zb + (long long)pos * cpr + cc);
    }
    for (int idx = tid; idx < ng * CW; idx += NT) {
      const int cc = idx % CW, pos = idx / CW;
      cp_async16(gtile + pos * PITCH + cc * 16, gb + (long long)pos * cpr + cc);
    }
    cp_async_wait();
    __syncthreads();
    for (int ks = 0; ks < (ng + 15) / 16; ++ks) {
      const int pb = ks * 16 + ib;
      const uint32_t gaddr = pb < ng ? gtile + pb * PITCH + 2 * warp * 16 : zero;
      uint32_t b0[2], b1[2];
      ldsm2t_(b0[0], b0[1], gaddr);
      ldsm2t_(b1[0], b1[1], gaddr + 16);
      const int pa = ks * 16 + ia;
      const int yy = pa / OW, x = pa - yy * OW, y = y0 + yy;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int sy = y * S - 1 + kh;
        const bool oky = pa < ng && sy >= 0 && sy < IH;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int sx = x * S - 1 + kw;
          const uint32_t zaddr = (oky && sx >= 0 && sx < IW) ? ztile + ((sy - lo) * IW + sx) * PITCH + ca : zero;
          uint32_t a[4];
          ldsm4t_(a, zaddr);
          mma16816_(acc[kh * 3 + kw][0], a, b0[0], b0[1]);
          mma16816_(acc[kh * 3 + kw][1], a, b1[0], b1[1]);
        }
      }
    }
    __syncthreads();
  }
  // diagonal of D0 (rows 0-7 x chunk 0) and of D1 (rows 8-15 x chunk 1)
  if (gid == 2 * tq || gid == 2 * tq + 1) {
    const int odd = gid == 2 * tq + 1;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float* d = dw + (long long)t * cpr * 8 + (cbase + 2 * warp) * 8 + gid;
      atomicAdd(d, odd ? acc[t][0][1] : acc[t][0][0]);
      atomicAdd(d + 8, odd ? acc[t][1][3] : acc[t][1][2]);
    }
  }
}

// 0 when the tiled kernels do not take this geometry, else the chunk-group width (12 or 8)
static int dwt_eligible(const DwRowsGeom& g) {
  const char* env = getenv("TLT_DWROWS_TILE");
  if (env && env[0] == 'o' && env[1] == 'f') return 0;
  if (g.ndim != 2 || g.kernel[0] != 3 || g.kernel[1] != 3 || g.padding[0] != 1 || g.padding[1] != 1 || g.dilation[0] != 1 || g.dilation[1] != 1)
    return 0;
  if (g.stride[0] != g.stride[1] || (g.stride[0] != 1 && g.stride[0] != 2)) return 0;
  if (g.in_size[1] > 64 || g.out_size[1] > 64) return 0;
  return g.cpr % 12 == 0 ? 12 : (g.cpr % 8 == 0 ? 8 : 0);
}

template <class Kern>
static int dwt_smem_attr(Kern kern, size_t smem) {
  return ensure_smem((const void*)kern, smem);
}

// strip height: as many destination rows as fit `budget` bytes of staged source rows, but enough strips to fill the GPU
template <class RowsFn>
static int dwt_pick_tr(int DH, long long groups, int pitch, long long budget, long long min_tiles, RowsFn staged_positions) {
  int TR = DH;
  while (TR > 1 && ((long long)staged_positions(TR) * pitch > budget || groups * ((DH + TR - 1) / TR) < min_tiles)) --TR;
  return TR;
}

template <bool DGRAD, int S, int CW>
static int dwt_launch(const DwRowsGeom& g, const void* src, const void* w, void* dst, cudaStream_t st) {
  constexpr int PITCH = (CW + 1) * 16;
  const int DH = DGRAD ? g.in_size[0] : g.out_size[0];
  const int SH = DGRAD ? g.out_size[0] : g.in_size[0], SWd = DGRAD ? g.out_size[1] : g.in_size[1];
  auto rows = [&](int tr) { int n = DGRAD ? tr / S + 2 : (tr - 1) * S + 3; return (long long)(n < SH ? n : SH) * SWd; };
  const int TR = dwt_pick_tr(DH, g.batch * (g.cpr / CW), PITCH, 28 * 1024, 148LL * 12, rows);
  const size_t smem = (size_t)rows(TR) * PITCH + PITCH + 9 * CW * 128;
  if (smem > 200 * 1024) return 1;                                    // not launched: caller falls back
  int rc = dwt_smem_attr(dwrows_tile_kernel<DGRAD, S, CW>, smem);
  if (rc) return rc;
  const int nstrips = (DH + TR - 1) / TR;
  const long long tiles = g.batch * nstrips;
  // persistent: exactly one resident wave (occupancy x SMs blocks), each block walking several strips -- the tap table is built
  // once per block, and a second partial wave would pay it (and the tail) again
  int occ = 0;
  TLT_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dwrows_tile_kernel<DGRAD, S, CW>, kDwtThreads, smem));
  if (occ < 1) occ = 1;
  long long by = (148LL * occ) / (g.cpr / CW);
  if (by < 1) by = 1;
  if (by > tiles) by = tiles;
  dim3 grid((unsigned)(g.cpr / CW), (unsigned)by);
  dwrows_tile_kernel<DGRAD, S, CW><<<grid, kDwtThreads, smem, st>>>(g, (const uint4*)src, (const __nv_bfloat16*)w, (__nv_bfloat16*)dst, TR, nstrips, tiles);
  return check_launch(DGRAD ? "dwrows_tile_kernel<dgrad>" : "dwrows_tile_kernel<fwd>");
}

template <int S, int CW>
static int dwt_launch_wgrad(const DwRowsGeom& g, const void* z, const void* gy, float* dw, cudaStream_t st) {
  constexpr int PITCH = (CW + 1) * 16;
  const int IH = g.in_size[0], IW = g.in_size[1], OH = g.out_size[0], OW = g.out_size[1];
  auto zrows = [&](int tr) { int n = (tr - 1) * S + 3; return (long long)(n < IH ? n : IH) * IW; };
  auto both = [&](int tr) { return zrows(tr) + (long long)tr * OW; };
  const int TR = dwt_pick_tr(OH, g.batch * (g.cpr / CW), PITCH, 88 * 1024, 148LL * 3, both);
  const size_t smem = (size_t)both(TR) * PITCH + PITCH;
  if (smem > 200 * 1024) return 1;
  int rc = dwt_smem_attr(dwrows_tile_wgrad_kernel<S, CW>, smem);
  if (rc) return rc;
  const int nstrips = (OH + TR - 1) / TR;
  const long long tiles = g.batch * nstrips;
  long long by = 148LL * 4 / (g.cpr / CW);
  if (by < 1) by = 1;
  if (by > tiles) by = tiles;
  dim3 grid((unsigned)(g.cpr / CW), (unsigned)by);
  dwrows_tile_wgrad_kernel<S, CW><<<grid, 16 * CW, smem, st>>>(g, (const uint4*)z, (const uint4*)gy, dw, TR, nstrips, tiles, (int)zrows(TR));
  return check_launch("dwrows_tile_wgrad_kernel");
}

// returns TLT_OK / error when handled, 1 when the caller should fall back to the register-window kernels
template <bool DGRAD>
static int dwt_dispatch(const DwRowsGeom& g, const void* src, const void* w, void* dst, cudaStream_t st) {
  const int cw = dwt_eligible(g);
  if (!cw) return 1;
  const int s = g.stride[0];
  if (cw == 12) return s == 1 ? dwt_launch<DGRAD, 1, 12>(g, src, w, dst, st) : dwt_launch<DGRAD, 2, 12>(g, src, w, dst, st);
  return s == 1 ? dwt_launch<DGRAD, 1, 8>(g, src, w, dst, st) : dwt_launch<DGRAD, 2, 8>(g, src, w, dst, st);
}
static int dwt_dispatch_wgrad(const DwRowsGeom& g, const void* z, const void* gy, float* dw, cudaStream_t st) {
  const int cw = dwt_eligible(g);
  if (!cw) return 1;
  const int s = g.stride[0];
  if (cw == 12) return s == 1 ? dwt_launch_wgrad<1, 12>(g, z, gy, dw, st) : dwt_launch_wgrad<2, 12>(g, z, gy, dw, st);
  return s == 1 ? dwt_launch_wgrad<1, 8>(g, z, gy, dw, st) : dwt_launch_wgrad<2, 8>(g, z, gy, dw, st);
}

}  // namespace tlt

using namespace tlt;

// rows alone give row_threads threads; cut each row of W outputs into segments when there are fewer than ~148 x 768 of them
static void dw_segments(long long row_threads, int W, int* segs, int* seg_len) {
  long long want = row_threads >= 148LL * 768 ? 1 : (148LL * 1024 + row_threads - 1) / row_threads;
  if (want < 1) want = 1;
  if (want > W) want = W;
  *seg_len = (int)((W + want - 1) / want);
  if (*seg_len < 4 && W >= 4) *seg_len = 4;                       // keep a few outputs per thread: the unpacked weights amortise
  *segs = (W + *seg_len - 1) / *seg_len;
}

static unsigned dw_grid(long long total) {
  long long blocks = (total + 255) / 256;
  const long long cap = 148LL * 16;
  return (unsigned)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

extern "C" int tlt_dwconv_rows_fwd(const tlt_conv_desc* d, const void* z, const void* w, void* y, void* stream) {
  DwRowsGeom g;
  int rc = fill_dw(d, &g);
  if (rc) return rc;
  TLT_REQUIRE(z && w && y && ((uintptr_t)z % 16) == 0 && ((uintptr_t)w % 16) == 0 && ((uintptr_t)y % 16) == 0, "tlt_dwconv_rows_fwd: 16-byte aligned tensors");
  rc = dwt_dispatch<false>(g, z, w, y, (cudaStream_t)stream);
  if (rc != 1) return rc;
  if (g.ndim == 2 && g.kernel[0] <= 3 && g.kernel[1] <= 3) {
    int segs, seg_len;
    dw_segments(g.batch * g.out_size[0] * g.cpr, g.out_size[1], &segs, &seg_len);
    const long long threads = g.batch * g.out_size[0] * g.cpr * segs;
    const unsigned blocks = (unsigned)((threads + 127) / 128 < 148LL * 32 ? (threads + 127) / 128 : 148LL * 32);
    dwconv_rows2d_kernel<false><<<blocks, 128, 0, (cudaStream_t)stream>>>(g, (const uint4*)z, (const uint4*)w, (uint4*)y, segs, seg_len);
    return check_launch("dwconv_rows2d_kernel<fwd>");
  }
  if (g.ndim == 2) {
    dwconv_rows2d_big_kernel<false><<<dw_grid(g.batch * g.out_sp * g.cpr), 256, 0, (cudaStream_t)stream>>>(g, (const uint4*)z, (const uint4*)w, (uint4*)y);
    return check_launch("dwconv_rows2d_big_kernel<fwd>");
  }
  dwconv_rows_kernel<false><<<dw_grid(g.batch * g.out_sp * g.cpr), 256, 0, (cudaStream_t)stream>>>(g, (const uint4*)z, (const uint4*)w, (uint4*)y);
  return check_launch("dwconv_rows_kernel<fwd>");
}

extern "C" int tlt_dwconv_rows_dgrad(const tlt_conv_desc* d, const void* gy, const void* w, void* dz, void* stream) {
  DwRowsGeom g;
  int rc = fill_dw(d, &g);
  if (rc) return rc;
  TLT_REQUIRE(gy && w && dz && ((uintptr_t)gy % 16) == 0 && ((uintptr_t)w % 16) == 0 && ((uintptr_t)dz % 16) == 0, "tlt_dwconv_rows_dgrad: 16-byte aligned tensors");
  rc = dwt_dispatch<true>(g, gy, w, dz, (cudaStream_t)stream);
  if (rc != 1) return rc;
  if (g.ndim == 2 && g.kernel[0] <= 3 && g.kernel[1] <= 3) {
    int segs, seg_len;
    dw_segments(g.batch * g.in_size[0] * g.cpr, g.in_size[1], &segs, &seg_len);
    const long long threads = g.batch * g.in_size[0] * g.cpr * segs;
    const unsigned blocks = (unsigned)((threads + 127) / 128 < 148LL * 32 ? (threads + 127) / 128 : 148LL * 32);
    dwconv_rows2d_kernel<true><<<blocks, 128, 0, (cudaStream_t)stream>>>(g, (const uint4*)gy, (const uint4*)w, (uint4*)dz, segs, seg_len);
    return check_launch("dwconv_rows2d_kernel<dgrad>");
  }
  if (g.ndim == 2) {
    dwconv_rows2d_big_kernel<true><<<dw_grid(g.batch * g.in_sp * g.cpr), 256, 0, (cudaStream_t)stream>>>(g, (const uint4*)gy, (const uint4*)w, (uint4*)dz);
    return check_launch("dwconv_rows2d_big_kernel<dgrad>");
  }
  dwconv_rows_kernel<true><<<dw_grid(g.batch * g.in_sp * g.cpr), 256, 0, (cudaStream_t)stream>>>(g, (const uint4*)gy, (const uint4*)w, (uint4*)dz);
  return check_launch("dwconv_rows_kernel<dgrad>");
}

extern "C" int tlt_dwconv_rows_wgrad(const tlt_conv_desc* d, const void* z, const void* gy, float* dw, void* stream) {
  DwRowsGeom g;
  int rc = fill_dw(d, &g);
  if (rc) return rc;
  TLT_REQUIRE(z && gy && dw && ((uintptr_t)z % 16) == 0 && ((uintptr_t)gy % 16) == 0, "tlt_dwconv_rows_wgrad: 16-byte aligned tensors");
  rc = dwt_dispatch_wgrad(g, z, gy, dw, (cudaStream_t)stream);
  if (rc != 1) return rc;
  if (g.taps > kDwTapsMax) { set_error("tlt_dwconv_rows_wgrad: more than %d taps", kDwTapsMax); return TLT_ERR_UNSUPPORTED; }
  const long long rows_total = g.batch * g.out_sp;
  long long blocks = 148LL * 4;
  int ppb = (int)((rows_total + blocks - 1) / blocks);
  if (ppb < 64) ppb = 64;
  blocks = (rows_total + ppb - 1) / ppb;
  const int passes = (g.cpr + 31) / 32;                              // chunk columns per pass: balanced, <= 32
  const int chunks_x = (g.cpr + passes - 1) / passes;
  const int lanes = 256 / chunks_x;
  const size_t smem = (size_t)lanes * g.taps * chunks_x * 8 * sizeof(float);
  TLT_ENSURE_SMEM(dwconv_rows_wgrad_kernel, smem);
  if (g.ndim == 2 && g.kernel[0] <= 3 && g.kernel[1] <= 3) {
    TLT_ENSURE_SMEM(dwconv_rows2d_wgrad_kernel, smem);
    dwconv_rows2d_wgrad_kernel<<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(g, (const uint4*)z, (const uint4*)gy, dw, ppb, chunks_x);
    return check_launch("dwconv_rows2d_wgrad_kernel");
  }
  dwconv_rows_wgrad_kernel<<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(g, (const uint4*)z, (const uint4*)gy, dw, ppb, chunks_x);
  return check_launch("dwconv_rows_wgrad_kernel");
}
