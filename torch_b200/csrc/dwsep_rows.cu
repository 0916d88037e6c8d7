// Separable 3x3 depthwise convolution on channels-last rows (bf16, padding 1, stride 1 or 2) -- the middle stage of the CP
// convolution chain when it is not fused into the 1x1 stages (cp_conv, tltorch/functional/convolution.py:251-269: one
// depthwise 1-D convolution per kernel mode, each its own pass over the R-channel tensor; here both passes in one sweep).
//
//   z2[n, ho, wo, r] = sum_{i,j} kh[i, r] kw'[j, r] z1[n, s ho + i - 1, s wo + j - 1, r]          kw' = lambda kw
//
// Rows are (B H W, ld) with ld = round8(R); a thread owns one 16-byte chunk (8 channels) of one output row of one image and
// WALKS along W: the H pass of a column is computed once when the column arrives, the W pass runs on the three most recent
// columns in registers -- 3 loads, 24 packed FMAs and one store per 8-channel output, all channel PAIRS as fma.rn.f32x2
// (a bf16x2 word unpacks into exactly such a pair).  The generic tap-loop kernels this replaces (dwconv_rows.cu) were
// instruction-issue bound at 93-116 us per pass on config 3's 28^2 layers (HBM floor 18 us).
//   forward / data gradient (stride 1: the same sweep with both tap vectors reversed; stride 2: the transposed sweep)
//   weight gradient: dT[3 i + j][r] = sum dz2[n, ho, wo, r] z1[n, s ho + i - 1, s wo + j - 1, r] in registers (a 3 x 3 window of
//   unpacked columns rotating through an unrolled loop), shared-memory then global fp32 reductions; tlt_dwsep_rows_finalize
//   turns dT into the gradients of kh, kw and lambda.
#include "common.cuh"
#include "f32x2.cuh"

namespace tlt {
namespace dws {

using namespace tlt::f32x2;

struct Geom {
  int batch, H, W, Ho, Wo, cpr;      // cpr = ld / 8 chunks per row
  long long items;                   // threads' work items: (image, row, chunk)
};

// 8 bf16 channels -> 4 fp32 pairs (channels 2i, 2i+1)
__device__ __forceinline__ void unpack(const uint4& v, f2 (&o)[4]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i] = pku(w[i] << 16, w[i] & 0xffff0000u);
}
__device__ __forceinline__ uint4 pack(const f2 (&o)[4]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo_of(o[i]), hi_of(o[i]));
    w[i] = *(uint32_t*)&h;
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ uint4 ldg16(const uint4* p, bool ok) { return ok ? __ldg(p) : make_uint4(0u, 0u, 0u, 0u); }

// taps [6][ld] fp32 (kh0 kh1 kh2 kw0 kw1 kw2) -> this chunk's 8 channels as pairs
__device__ __forceinline__ void load_taps(const float* __restrict__ taps, int ld, int c, f2 (&k)[6][4]) {
#pragma unroll
  for (int t = 0; t < 6; ++t) {
    const float4 a = __ldg((const float4*)(taps + (size_t)t * ld + c * 8));
    const float4 b = __ldg((const float4*)(taps + (size_t)t * ld + c * 8) + 1);
    k[t][0] = pk(a.x, a.y); k[t][1] = pk(a.z, a.w); k[t][2] = pk(b.x, b.y); k[t][3] = pk(b.z, b.w);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Forward sweep (and the stride-1 data gradient, called with reversed taps): item = (n, ho, chunk), walk over the input columns.
// ---------------------------------------------------------------------------------------------------------------------
template <int STRIDE>
__global__ void __launch_bounds__(128) dwsep_fwd_kernel(const Geom g, const uint4* __restrict__ src, const float* __restrict__ taps,
                                                        uint4* __restrict__ dst) {
  for (long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x; item < g.items; item += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(item % g.cpr);
    const long long t = item / g.cpr;
    const int ho = (int)(t % g.Ho), n = (int)(t / g.Ho);
    f2 k[6][4];
    load_taps(taps, g.cpr * 8, c, k);
    const int h1 = ho * STRIDE;
    const bool ok0 = h1 >= 1, ok2 = h1 + 1 < g.H;
    const uint4* r1 = src + ((long long)(n * g.H + h1) * g.W) * g.cpr + c;
    const uint4* r0 = r1 - (long long)g.W * g.cpr;
    const uint4* r2 = r1 + (long long)g.W * g.cpr;
    uint4* out = dst + ((long long)(n * g.Ho + ho) * g.Wo) * g.cpr + c;
    f2 va[4], vb[4];                                   // H-pass results of columns w-2, w-1
#pragma unroll
    for (int i = 0; i < 4; ++i) { va[i] = 0ull; vb[i] = 0ull; }
    // two columns per trip; the next two are in flight while these are finished (the sweep is latency-bound otherwise)
    uint4 q[2][3];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const bool okw = u < g.W;
      q[u][0] = ldg16(r0 + (long long)u * g.cpr, ok0 && okw); q[u][1] = ldg16(r1 + (long long)u * g.cpr, okw);
      q[u][2] = ldg16(r2 + (long long)u * g.cpr, ok2 && okw);
    }
    for (int w0 = 0; w0 < g.W; w0 += 2) {
      f2 a[2][4], b[2][4], d[2][4];
#pragma unroll
      for (int u = 0; u < 2; ++u) { unpack(q[u][0], a[u]); unpack(q[u][1], b[u]); unpack(q[u][2], d[u]); }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int x = w0 + 2 + u;
        const bool okw = x < g.W;
        const long long off = (long long)x * g.cpr;
        q[u][0] = ldg16(r0 + off, ok0 && okw); q[u][1] = ldg16(r1 + off, okw); q[u][2] = ldg16(r2 + off, ok2 && okw);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int w = w0 + u;
        if (w < g.W) {
          f2 v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = fma2(k[2][i], d[u][i], fma2(k[1][i], b[u][i], mul2(k[0][i], a[u][i])));
          // output column w-1 (stride 1) or (w-1)/2 (stride 2, w odd) : kw0 v[w-2] + kw1 v[w-1] + kw2 v[w]
          if (w >= 1 && (STRIDE == 1 || u == 1)) {
            f2 o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = fma2(k[5][i], v[i], fma2(k[4][i], vb[i], mul2(k[3][i], va[i])));
            out[(long long)((w - 1) / STRIDE) * g.cpr] = pack(o);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) { va[i] = vb[i]; vb[i] = v[i]; }
        }
      }
    }
    if (STRIDE == 1) {                                 // last column: right neighbour is the zero padding
      f2 o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = fma2(k[4][i], vb[i], mul2(k[3][i], va[i]));
      out[(long long)(g.W - 1) * g.cpr] = pack(o);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Stride-2 data gradient: item = (n, hi, chunk) at INPUT resolution, walk over the output-resolution columns b.
//   g[b]        = kh1 dz2[a][b]                         (hi = 2a)      |  kh2 dz2[a][b] + kh0 dz2[a+1][b]   (hi = 2a + 1)
//   dz1[hi][2b] = kw1 g[b],   dz1[hi][2b-1] = kw2 g[b-1] + kw0 g[b]
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) dwsep_dgrad2_kernel(const Geom g, const uint4* __restrict__ gy, const float* __restrict__ taps,
                                                           uint4* __restrict__ dz) {
  for (long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x; item < g.items; item += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(item % g.cpr);
    const long long t = item / g.cpr;
    const int hi = (int)(t % g.H), n = (int)(t / g.H);
    f2 k[6][4];
    load_taps(taps, g.cpr * 8, c, k);
    const int a = hi >> 1;
    const bool odd = hi & 1, ok1 = odd && a + 1 < g.Ho;
    const uint4* ra = gy + ((long long)(n * g.Ho + a) * g.Wo) * g.cpr + c;
    const uint4* rb = ra + (long long)g.Wo * g.cpr;
    uint4* out = dz + ((long long)(n * g.H + hi) * g.W) * g.cpr + c;
    f2 gp[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) gp[i] = 0ull;
    uint4 qa = __ldg(ra), qb = ldg16(rb, ok1);
    for (int b = 0; b < g.Wo; ++b) {
      f2 ya[4], yb[4], gb[4], o[4];
      unpack(qa, ya); unpack(qb, yb);
      if (b + 1 < g.Wo) { qa = __ldg(ra + (long long)(b + 1) * g.cpr); qb = ldg16(rb + (long long)(b + 1) * g.cpr, ok1); }
#pragma unroll
      for (int i = 0; i < 4; ++i) gb[i] = odd ? fma2(k[0][i], yb[i], mul2(k[2][i], ya[i])) : mul2(k[1][i], ya[i]);
      if (b >= 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = fma2(k[3][i], gb[i], mul2(k[5][i], gp[i]));
        out[(long long)(2 * b - 1) * g.cpr] = pack(o);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = mul2(k[4][i], gb[i]);
      out[(long long)(2 * b) * g.cpr] = pack(o);
#pragma unroll
      for (int i = 0; i < 4; ++i) gp[i] = gb[i];
    }
    f2 o[4];                                           // column W-1 = 2 (Wo-1) + 1: g[Wo] is the zero padding
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = mul2(k[5][i], gp[i]);
    out[(long long)(g.W - 1) * g.cpr] = pack(o);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Weight gradient: item = (n, ho, chunk), walk over the output columns with a 3 x 3 window of unpacked z1 columns.
// acc [9][ld] fp32 += per-block partial sums (shared-memory reduction over the block's rows first).
// ---------------------------------------------------------------------------------------------------------------------
template <int STRIDE>
__global__ void __launch_bounds__(256) dwsep_wgrad_kernel(const Geom g, const uint4* __restrict__ z, const uint4* __restrict__ gy,
                                                          float* __restrict__ part) {
  extern __shared__ float sacc[];                      // [9][ld]
  const int ld = g.cpr * 8;
  for (int i = threadIdx.x; i < 9 * ld; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  // a thread keeps ONE channel chunk for the whole kernel and walks many (image, row) items: its 72 sums stay in registers and
  // are reduced once -- shared memory across the block's row lanes, then one plain store per block into part[block][9][ld]
  const int lanes = blockDim.x / g.cpr;
  const int c = threadIdx.x % g.cpr, lane = threadIdx.x / g.cpr;
  const long long rows_total = (long long)g.batch * g.Ho;
  f2 dT[9][4];
#pragma unroll
  for (int a = 0; a < 9; ++a)
#pragma unroll
    for (int i = 0; i < 4; ++i) dT[a][i] = 0ull;
  if (lane < lanes) {
    for (long long t = (long long)blockIdx.x * lanes + lane; t < rows_total; t += (long long)gridDim.x * lanes) {
      const int ho = (int)(t % g.Ho), n = (int)(t / g.Ho);
      const int h1 = ho * STRIDE;
      const bool ok0 = h1 >= 1, ok2 = h1 + 1 < g.H;
      const uint4* r1 = z + ((long long)(n * g.H + h1) * g.W) * g.cpr + c;
      const uint4* r0 = r1 - (long long)g.W * g.cpr;
      const uint4* r2 = r1 + (long long)g.W * g.cpr;
      const uint4* rg = gy + ((long long)(n * g.Ho + ho) * g.Wo) * g.cpr + c;
      f2 win[3][3][4];                                 // [column slot][row][pair]
#pragma unroll
      for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int i = 0; i < 4; ++i) win[s][r][i] = 0ull;
      // raw loads of the NEXT centre's new columns are in flight while the current centre is accumulated
      constexpr int NC = STRIDE == 1 ? 1 : 2;          // new input columns per centre
      // two centres ahead: qz / qg hold the centre being consumed next, nz / ng the one after it
      uint4 qz[NC][3], qg, nz[NC][3], ng;
      auto fetch = [&](int wo, uint4 (&tz)[NC][3], uint4& tg) {   // centre wo needs columns wo+1 (stride 1) / 2wo, 2wo+1 (stride 2)
#pragma unroll
        for (int u = 0; u < NC; ++u) {
          const int x = STRIDE == 1 ? wo + 1 : 2 * wo + u;
          const bool okx = wo < g.Wo && x < g.W;
          const long long off = (long long)x * g.cpr;
          tz[u][0] = ldg16(r0 + off, okx && ok0); tz[u][1] = ldg16(r1 + off, okx); tz[u][2] = ldg16(r2 + off, okx && ok2);
        }
        tg = ldg16(rg + (long long)wo * g.cpr, wo < g.Wo);
      };
      auto issue = [&](int wo) {                       // called once the current centre's raw registers are unpacked
#pragma unroll
        for (int u = 0; u < NC; ++u)
#pragma unroll
          for (int r = 0; r < 3; ++r) qz[u][r] = nz[u][r];
        qg = ng;
        fetch(wo + 1, nz, ng);
      };
      if (STRIDE == 1) {                               // column 0 -> slot 1 (column -1 = slot 0 stays zero)
        unpack(ldg16(r0, ok0), win[1][0]); unpack(__ldg(r1), win[1][1]); unpack(ldg16(r2, ok2), win[1][2]);
      }
      fetch(0, qz, qg);
      fetch(1, nz, ng);
      for (int w0 = 0; w0 < g.Wo; w0 += 3) {
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int wo = w0 + u;
          if (wo < g.Wo) {
            f2 d[4];
            unpack(qg, d);
            if (STRIDE == 1) {
              // slots: column wo-1 -> u % 3, wo -> (u+1) % 3, wo+1 -> (u+2) % 3 (w0 is a multiple of 3)
#pragma unroll
              for (int r = 0; r < 3; ++r) unpack(qz[0][r], win[(u + 2) % 3][r]);
              issue(wo + 1);
#pragma unroll
              for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                  for (int i = 0; i < 4; ++i) dT[3 * r + j][i] = fma2(d[i], win[(u + j) % 3][r][i], dT[3 * r + j][i]);
            } else {
              // columns 2wo-1 (kept from the previous centre in slot 0; zero for wo = 0), 2wo -> slot 1, 2wo+1 -> slot 2
#pragma unroll
              for (int r = 0; r < 3; ++r) { unpack(qz[0][r], win[1][r]); unpack(qz[1][r], win[2][r]); }
              issue(wo + 1);
#pragma unroll
              for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                  for (int i = 0; i < 4; ++i) dT[3 * r + j][i] = fma2(d[i], win[j][r][i], dT[3 * r + j][i]);
#pragma unroll
              for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int i = 0; i < 4; ++i) win[0][r][i] = win[2][r][i];
            }
          }
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 9; ++a)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        atomicAdd(sacc + a * ld + c * 8 + 2 * i, lo_of(dT[a][i]));
        atomicAdd(sacc + a * ld + c * 8 + 2 * i + 1, hi_of(dT[a][i]));
      }
  }
  __syncthreads();
  float* out = part + (size_t)blockIdx.x * 9 * ld;
  for (int i = threadIdx.x; i < 9 * ld; i += blockDim.x) out[i] = sacc[i];
}

// taps [6][ld] fp32 = kh0 kh1 kh2, lambda kw0, lambda kw1, lambda kw2 (flip: both triples reversed), pad channels zero
__global__ void dwsep_taps_kernel(const __nv_bfloat16* __restrict__ kh, const __nv_bfloat16* __restrict__ kw,
                                  const __nv_bfloat16* __restrict__ lam, int R, int ld, int flip, float* __restrict__ taps) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 6 * ld; i += gridDim.x * blockDim.x) {
    const int t = i / ld, r = i - t * ld;
    float v = 0.f;
    if (r < R) {
      const int tt = flip ? (t < 3 ? 2 - t : 8 - t) : t;
      v = tt < 3 ? __bfloat162float(kh[tt * R + r]) : __bfloat162float(kw[(tt - 3) * R + r]) * __bfloat162float(lam[r]);
    }
    taps[i] = v;
  }
}

// partial sums [parts][9][ld] -> bf16 gradients of kh (3, R), kw (3, R), lambda (R).  Block = 8 channels x 32 part lanes: the slabs
// are summed 32 at a time (a serial walk over all of them by one thread per channel cost 64 us), shared-memory tree, 8 writers.
__global__ void __launch_bounds__(256) dwsep_finalize_kernel(const float* __restrict__ dT, int parts, const __nv_bfloat16* __restrict__ kh,
                                                             const __nv_bfloat16* __restrict__ kw, const __nv_bfloat16* __restrict__ lam, int R,
                                                             int ld, __nv_bfloat16* __restrict__ g_kh, __nv_bfloat16* __restrict__ g_kw,
                                                             __nv_bfloat16* __restrict__ g_lam) {
  __shared__ float red[9][32][8];
  const int ch = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int r = blockIdx.x * 8 + ch;
  float t[9];
#pragma unroll
  for (int a = 0; a < 9; ++a) t[a] = 0.f;
  // 5 x 9 independent loads in flight per thread (see cpc::post_kernel)
  if (r < R)
    for (int p0 = pl; p0 < parts; p0 += 32 * 5) {
      float v[5][9];
#pragma unroll
      for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int a = 0; a < 9; ++a) v[i][a] = p0 + 32 * i < parts ? __ldg(dT + ((size_t)(p0 + 32 * i) * 9 + a) * ld + r) : 0.f;
#pragma unroll
      for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int a = 0; a < 9; ++a) t[a] += v[i][a];
    }
#pragma unroll
  for (int a = 0; a < 9; ++a) red[a][pl][ch] = t[a];
  __syncthreads();
  for (int s = 16; s >= 1; s >>= 1) {
    if (pl < s)
#pragma unroll
      for (int a = 0; a < 9; ++a) red[a][pl][ch] += red[a][pl + s][ch];
    __syncthreads();
  }
  if (pl == 0 && r < R) {
    float h[3], w[3];
    const float l = __bfloat162float(lam[r]);
#pragma unroll
    for (int a = 0; a < 3; ++a) { h[a] = __bfloat162float(kh[a * R + r]); w[a] = __bfloat162float(kw[a * R + r]); }
#pragma unroll
    for (int a = 0; a < 9; ++a) t[a] = red[a][0][ch];
    float gl = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float gh = l * (w[0] * t[3 * a] + w[1] * t[3 * a + 1] + w[2] * t[3 * a + 2]);
      const float gwl = h[0] * t[a] + h[1] * t[3 + a] + h[2] * t[6 + a];
      g_kh[a * R + r] = __float2bfloat16_rn(gh);
      g_kw[a * R + r] = __float2bfloat16_rn(l * gwl);
      gl += w[a] * gwl;
    }
    g_lam[r] = __float2bfloat16_rn(gl);
  }
}

static int fill(const tlt_dwsep_desc* d, Geom* g, bool input_rows) {
  TLT_REQUIRE(d && d->batch >= 1 && d->height >= 1 && d->width >= 2 && d->ld >= 8 && d->ld % 8 == 0 && (d->stride == 1 || d->stride == 2),
              "tlt_dwsep_rows: bad descriptor");
  TLT_REQUIRE(d->stride == 1 || (d->height % 2 == 0 && d->width % 2 == 0), "tlt_dwsep_rows: stride 2 needs even map extents");
  g->batch = (int)d->batch; g->H = (int)d->height; g->W = (int)d->width;
  g->Ho = g->H / d->stride; g->Wo = g->W / d->stride; g->cpr = (int)(d->ld / 8);
  g->items = (long long)g->batch * (input_rows ? g->H : g->Ho) * g->cpr;
  TLT_REQUIRE((long long)g->batch * g->H * g->W * g->cpr < (1LL << 40), "tlt_dwsep_rows: tensor too large");
  return TLT_OK;
}
static unsigned grid_for(long long items) {
  const long long blocks = (items + 127) / 128, cap = (long long)device_sms() * 16;
  return (unsigned)(blocks < cap ? blocks : cap);
}

}  // namespace dws
}  // namespace tlt

namespace tlt {
namespace dwst {   // dwsep_staged.cu: the same sweeps from TMA-landed bands in shared memory (stride 1); 1 launched, 0 not covered
int staged_fwd(int batch, int H, int W, int ld, int stride, const void* z, const float* taps, void* y, cudaStream_t st);
int staged_wgrad(int batch, int H, int W, int ld, int stride, const void* z, const void* gy, float* acc, cudaStream_t st);
}  // namespace dwst
}  // namespace tlt

using namespace tlt;
using namespace tlt::dws;

static bool staged_enabled() {
  static const bool on = [] { const char* e = getenv("TLT_DWSEP_STAGED"); return !(e && e[0] == '0'); }();
  return on;
}

extern "C" int tlt_dwsep_rows_taps(const void* kh, const void* kw, const void* lam, int64_t rank, int64_t ld, int32_t flip, float* taps,
                                   void* stream) {
  TLT_REQUIRE(kh && kw && lam && taps && rank >= 1 && ld >= rank && ld % 8 == 0, "tlt_dwsep_rows_taps: bad arguments");
  dwsep_taps_kernel<<<(unsigned)((6 * ld + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)kh, (const __nv_bfloat16*)kw,
                                                                                       (const __nv_bfloat16*)lam, (int)rank, (int)ld, flip, taps);
  return check_launch("dwsep_taps_kernel");
}

extern "C" int tlt_dwsep_rows_fwd(const tlt_dwsep_desc* d, const void* z, const float* taps, void* y, void* stream) {
  Geom g;
  int rc = fill(d, &g, false);
  if (rc) return rc;
  TLT_REQUIRE(z && taps && y && ((uintptr_t)z % 16) == 0 && ((uintptr_t)y % 16) == 0 && ((uintptr_t)taps % 16) == 0, "tlt_dwsep_rows_fwd: 16-byte aligned tensors");
  if (staged_enabled() && dwst::staged_fwd(g.batch, g.H, g.W, (int)d->ld, (int)d->stride, z, taps, y, (cudaStream_t)stream) == 1)
    return check_launch("dwsep_staged_kernel<fwd>");
  if (d->stride == 1) dwsep_fwd_kernel<1><<<grid_for(g.items), 128, 0, (cudaStream_t)stream>>>(g, (const uint4*)z, taps, (uint4*)y);
  else dwsep_fwd_kernel<2><<<grid_for(g.items), 128, 0, (cudaStream_t)stream>>>(g, (const uint4*)z, taps, (uint4*)y);
  return check_launch("dwsep_fwd_kernel");
}

// taps_flipped: tlt_dwsep_rows_taps(..., flip = 1) for stride 1; the UNFLIPPED taps for stride 2 (the transposed sweep indexes them itself)
extern "C" int tlt_dwsep_rows_dgrad(const tlt_dwsep_desc* d, const void* gy, const float* taps, void* dz, void* stream) {
  Geom g;
  int rc = fill(d, &g, true);
  if (rc) return rc;
  TLT_REQUIRE(gy && taps && dz && ((uintptr_t)gy % 16) == 0 && ((uintptr_t)dz % 16) == 0 && ((uintptr_t)taps % 16) == 0, "tlt_dwsep_rows_dgrad: 16-byte aligned tensors");
  if (d->stride == 1 && staged_enabled() && dwst::staged_fwd(g.batch, g.H, g.W, (int)d->ld, 1, gy, taps, dz, (cudaStream_t)stream) == 1)
    return check_launch("dwsep_staged_kernel<dgrad>");
  if (d->stride == 1) dwsep_fwd_kernel<1><<<grid_for(g.items), 128, 0, (cudaStream_t)stream>>>(g, (const uint4*)gy, taps, (uint4*)dz);
  else dwsep_dgrad2_kernel<<<grid_for(g.items), 128, 0, (cudaStream_t)stream>>>(g, (const uint4*)gy, taps, (uint4*)dz);
  return check_launch("dwsep dgrad kernel");
}

extern "C" int tlt_dwsep_rows_wgrad_parts(void) { return device_sms(); }   // one 256-thread block per SM (the kernel is register-heavy)

// acc: fp32 [tlt_dwsep_rows_wgrad_parts()][9][ld], fully written by the call (no zeroing needed)
extern "C" int tlt_dwsep_rows_wgrad(const tlt_dwsep_desc* d, const void* z, const void* gy, float* acc, void* stream) {
  Geom g;
  int rc = fill(d, &g, false);
  if (rc) return rc;
  TLT_REQUIRE(z && gy && acc && ((uintptr_t)z % 16) == 0 && ((uintptr_t)gy % 16) == 0, "tlt_dwsep_rows_wgrad: 16-byte aligned tensors");
  const size_t smem = (size_t)9 * d->ld * sizeof(float);
  TLT_REQUIRE(smem <= 48 * 1024 && g.cpr <= 256, "tlt_dwsep_rows_wgrad: ld too large");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)tlt_dwsep_rows_wgrad_parts();
  if (staged_enabled() && dwst::staged_wgrad(g.batch, g.H, g.W, (int)d->ld, (int)d->stride, z, gy, acc, st) == 1)
    return check_launch("dwsep_staged_kernel<wgrad>");
  if (d->stride == 1) dwsep_wgrad_kernel<1><<<grid, 256, smem, st>>>(g, (const uint4*)z, (const uint4*)gy, acc);
  else dwsep_wgrad_kernel<2><<<grid, 256, smem, st>>>(g, (const uint4*)z, (const uint4*)gy, acc);
  return check_launch("dwsep_wgrad_kernel");
}

extern "C" int tlt_dwsep_rows_finalize(const float* acc, int32_t parts, const void* kh, const void* kw, const void* lam, int64_t rank,
                                       int64_t ld, void* g_kh, void* g_kw, void* g_lam, void* stream) {
  TLT_REQUIRE(acc && parts >= 1 && kh && kw && lam && g_kh && g_kw && g_lam && rank >= 1 && ld >= rank, "tlt_dwsep_rows_finalize: bad arguments");
  dwsep_finalize_kernel<<<(unsigned)((rank + 7) / 8), 256, 0, (cudaStream_t)stream>>>(acc, parts, (const __nv_bfloat16*)kh,
                                                                                       (const __nv_bfloat16*)kw, (const __nv_bfloat16*)lam,
                                                                                       (int)rank, (int)ld, (__nv_bfloat16*)g_kh,
                                                                                       (__nv_bfloat16*)g_kw, (__nv_bfloat16*)g_lam);
  return check_launch("dwsep_finalize_kernel");
}
