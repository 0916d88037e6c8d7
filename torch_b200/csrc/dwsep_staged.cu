// Separable 3x3 depthwise sweep on channels-last rows, stride 1, STAGED: the band of image rows a block works on is landed in
// shared memory by TMA bulk copies (cp.async.bulk + mbarrier, double-buffered, the next band always in flight) and every thread
// computes from shared memory.  Same math and the same entry points as dwsep_rows.cu (cp_conv's depthwise stage,
// tltorch/functional/convolution.py:251-269), which walks rows with per-thread global loads: on config 3's 28^2 / 14^2 / 7^2 layers
// those kernels are latency-bound (ncu, profiles/r02_dwsep.md: 8-12 resident warps per SM each waiting on its own 16-byte loads --
// weight gradient 94 us for 116 MB, forward 38 us) because a 234-register weight-gradient thread caps the SM at 8 warps.  Here the
// memory system is driven by the copy engine (80 KB in flight per SM regardless of occupancy) and the register-heavy threads only
// ever wait on shared memory.
//
//   rows are (B H W, ld), ld = 8 cpr;  a BAND = BR output rows of one image (+ one halo row above and below for the input);
//   a row's bytes are contiguous in memory, so a band is one contiguous span: one bulk copy per image row;
//   thread = (16-byte channel chunk c, row r of the band, segment of SL columns): it keeps c for the whole kernel, so the nine
//   weight-gradient sums of its 8 channels stay in registers across all bands of the block.
#include "tc_common.cuh"
#include "f32x2.cuh"
#include <stdio.h>
#include <stdlib.h>

namespace tlt {
namespace dwst {

using namespace tlt::f32x2;

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint4 lds16(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void unpack(const uint4& v, f2 (&o)[4]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i] = pku(w[i] << 16, w[i] & 0xffff0000u);
}
__device__ __forceinline__ uint4 pack(const f2 (&o)[4]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = pack_bf16x2(lo_of(o[i]), hi_of(o[i]));
  return make_uint4(w[0], w[1], w[2], w[3]);
}

struct SGeom {
  int batch, H, W, cpr;           // H, W: the INPUT map (the output map is H / S x W / S; bands, rows r and segments are output rows / columns)
  int S, Ho, Wo;                  // stride (2: weight gradient only)
  uint32_t grow_bytes;            // bytes of a row of the second operand (Wo * ld * 2)
  int BR, SL, NSEG, bands;        // band rows, segment length, segments per row, bands per image
  int units;                      // batch * bands
  int nst;                        // stages of the band ring (2..4)
  int red;                        // weight gradient: the ring is large enough for the [slots][9][ld] block reduction
  uint32_t row_bytes;             // W * ld * 2
  uint32_t z_stage, g_stage;      // bytes of one stage of the input slab ((BR + 2) rows) / of the second operand's slab (BR rows)
};

constexpr int kWgradThreads = 384, kFwdThreads = 512;

// Lands unit u's band: input rows [h0 - 1, h0 + BR] that exist (and, for the weight gradient, the BR rows of the other operand)
template <bool WGRAD>
__device__ __forceinline__ void issue_unit(const SGeom& g, int u, const char* src, const char* gy, uint32_t zbuf, uint32_t gbuf, uint32_t bar) {
  const int n = u / g.bands, h0 = (u - n * g.bands) * g.BR;                      // h0: first OUTPUT row of the band
  const int zfirst = g.S * h0 - 1;                                               // image row held by slab row 0
  const int hz0 = zfirst >= 0 ? zfirst : 0, hz1 = min(g.S * (min(h0 + g.BR, g.Ho) - 1) + 2, g.H);   // input rows [hz0, hz1)
  const int hg1 = min(h0 + g.BR, g.Ho);
  const uint32_t bytes = (uint32_t)(hz1 - hz0) * g.row_bytes + (WGRAD ? (uint32_t)(hg1 - h0) * g.grow_bytes : 0u);
  mbar_expect_tx(bar, bytes);
  const char* base = src + (size_t)n * g.H * g.row_bytes;
  for (int h = hz0; h < hz1; ++h) bulk_g2s(zbuf + (uint32_t)(h - zfirst) * g.row_bytes, base + (size_t)h * g.row_bytes, g.row_bytes, bar);
  if (WGRAD) {
    const char* gb = gy + (size_t)n * g.Ho * g.grow_bytes;
    for (int h = h0; h < hg1; ++h) bulk_g2s(gbuf + (uint32_t)(h - h0) * g.grow_bytes, gb + (size_t)h * g.grow_bytes, g.grow_bytes, bar);
  }
}

// FWD (also the stride-1 data gradient, called with reversed taps):  dst[h][w] = sum_j kw'[j] v[w + j - 1],  v[x] = sum_i kh[i] src[h + i - 1][x]
// WGRAD:  dT[3 i + j] += gy[h][w] * src[h + i - 1][w + j - 1]   -> part[block][9][ld]
template <bool WGRAD>
__global__ void __launch_bounds__(WGRAD ? kWgradThreads : kFwdThreads, 1)
dwsep_staged_kernel(const SGeom g, const char* __restrict__ src, const char* __restrict__ gy, const float* __restrict__ taps,
                    uint4* __restrict__ dst, float* __restrict__ part) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t stage_bytes = g.z_stage + (WGRAD ? g.g_stage : 0u);
  const uint32_t bar0 = sbase + (uint32_t)g.nst * stage_bytes;         // one mbarrier per stage
  float* sacc = (float*)(smem + (size_t)g.nst * stage_bytes + 32);     // WGRAD: [9][ld]
  const int ld = g.cpr * 8;
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < g.nst; ++s) mbar_init(bar0 + 8 * s, 1);
    fence_barrier_init();
  }
  if (WGRAD) for (int i = tid; i < 9 * ld; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();

  const int c = tid % g.cpr, slot = tid / g.cpr;
  const int r = slot / g.NSEG, sg = slot - r * g.NSEG;
  const bool lane_ok = slot < g.BR * g.NSEG;
  const int w_begin = sg * g.SL, w_end = min(g.Wo, w_begin + g.SL);      // output columns of this thread's segment

  f2 k[6][4];                                                          // FWD: taps of this chunk's 8 channels
  if (!WGRAD) {
#pragma unroll
    for (int t = 0; t < 6; ++t) {
      const float4 a = __ldg((const float4*)(taps + (size_t)t * ld + c * 8));
      const float4 b = __ldg((const float4*)(taps + (size_t)t * ld + c * 8) + 1);
      k[t][0] = pk(a.x, a.y); k[t][1] = pk(a.z, a.w); k[t][2] = pk(b.x, b.y); k[t][3] = pk(b.z, b.w);
    }
  }
  f2 dT[9][4];
  if (WGRAD) {
#pragma unroll
    for (int a = 0; a < 9; ++a)
#pragma unroll
      for (int i = 0; i < 4; ++i) dT[a][i] = 0ull;
  }

  // ring of nst band slabs: nst - 1 bands are in flight while one is consumed (a band takes 2-3 us to arrive when every SM pulls
  // its share of the HBM bandwidth; with ONE band in flight the sweep ran at 2.9 TB/s however fast the threads were)
  int u = blockIdx.x;
  if (tid == 0)
    for (int s = 0; s < g.nst - 1; ++s) {
      const long long us = (long long)u + (long long)s * gridDim.x;
      if (us < g.units) issue_unit<WGRAD>(g, (int)us, src, gy, sbase + (uint32_t)s * stage_bytes, sbase + (uint32_t)s * stage_bytes + g.z_stage, bar0 + 8 * s);
    }
  int st = 0, ph = 0;
  for (int it = 0; u < g.units; ++it, u += gridDim.x, st = (st + 1 == g.nst ? 0 : st + 1), ph ^= (st == 0)) {
    __syncthreads();                                                   // everybody is done with unit it - 1: its stage is free again
    const long long un = (long long)u + (long long)(g.nst - 1) * gridDim.x;
    if (tid == 0 && un < g.units) {
      const int sn = st == 0 ? g.nst - 1 : st - 1;
      const uint32_t zb = sbase + (uint32_t)sn * stage_bytes;
      issue_unit<WGRAD>(g, (int)un, src, gy, zb, zb + g.z_stage, bar0 + 8 * sn);
    }
    mbar_wait(bar0 + 8 * st, (uint32_t)ph);
    const int n = u / g.bands, h0 = (u - n * g.bands) * g.BR;
    const int h = h0 + r;                                              // output row
    if (!lane_ok || h >= g.Ho) continue;
    const bool ok0 = g.S * h >= 1, ok2 = g.S * h + 1 < g.H;
    const uint32_t zbuf = sbase + (uint32_t)st * stage_bytes;
    // slab row j holds image row S h0 - 1 + j: input rows S h - 1, S h, S h + 1 are j = S r, S r + 1, S r + 2
    const uint32_t a0 = zbuf + (uint32_t)(g.S * r) * g.row_bytes + (uint32_t)c * 16, a1 = a0 + g.row_bytes, a2 = a1 + g.row_bytes;
    const uint32_t colb = (uint32_t)g.cpr * 16;                        // bytes per position
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);

    if (!WGRAD) {
      auto hpass = [&](int x, f2 (&v)[4]) {                           // v[x], zero outside the row
        if (x < 0 || x >= g.W) {
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = 0ull;
          return;
        }
        f2 a[4], b[4], d[4];
        unpack(ok0 ? lds16(a0 + (uint32_t)x * colb) : zero, a);
        unpack(lds16(a1 + (uint32_t)x * colb), b);
        unpack(ok2 ? lds16(a2 + (uint32_t)x * colb) : zero, d);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = fma2(k[2][i], d[i], fma2(k[1][i], b[i], mul2(k[0][i], a[i])));
      };
      f2 va[4], vb[4], vc[4];
      uint4* out = dst + ((size_t)(n * g.Ho + h) * g.Wo) * g.cpr + c;
      if (g.S == 2) {
        // stride 2: output column w reads v at input columns 2 w - 1, 2 w, 2 w + 1; the last one is the next output's first
        hpass(2 * w_begin - 1, va);
        for (int w = w_begin; w < w_end; ++w) {
          hpass(2 * w, vb);
          hpass(2 * w + 1, vc);
          f2 o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] = fma2(k[5][i], vc[i], fma2(k[4][i], vb[i], mul2(k[3][i], va[i])));
          out[(size_t)w * g.cpr] = pack(o);
#pragma unroll
          for (int i = 0; i < 4; ++i) va[i] = vc[i];
        }
        continue;
      }
      hpass(w_begin - 1, va);
      hpass(w_begin, vb);
#pragma unroll 2
      for (int w = w_begin; w < w_end; ++w) {
        hpass(w + 1, vc);
        f2 o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = fma2(k[5][i], vc[i], fma2(k[4][i], vb[i], mul2(k[3][i], va[i])));
        out[(size_t)w * g.cpr] = pack(o);
#pragma unroll
        for (int i = 0; i < 4; ++i) { va[i] = vb[i]; vb[i] = vc[i]; }
      }
    } else {
      const uint32_t gb = zbuf + g.z_stage + (uint32_t)r * g.grow_bytes + (uint32_t)c * 16;
      f2 win[3][3][4];                                                 // [column slot][row][pair]
      auto col = [&](int x, f2 (&dstw)[3][4]) {
        const bool okx = x >= 0 && x < g.W;
        unpack(okx && ok0 ? lds16(a0 + (uint32_t)x * colb) : zero, dstw[0]);
        unpack(okx ? lds16(a1 + (uint32_t)x * colb) : zero, dstw[1]);
        unpack(okx && ok2 ? lds16(a2 + (uint32_t)x * colb) : zero, dstw[2]);
      };
      if (g.S == 2) {
        // stride 2: centre w reads input columns 2 w - 1, 2 w, 2 w + 1; the last one is the next centre's first
        col(2 * w_begin - 1, win[0]);
        for (int w = w_begin; w < w_end; ++w) {
          col(2 * w, win[1]);
          col(2 * w + 1, win[2]);
          f2 d[4];
          unpack(lds16(gb + (uint32_t)w * colb), d);
#pragma unroll
          for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int rr = 0; rr < 3; ++rr)
#pragma unroll
              for (int i = 0; i < 4; ++i) dT[3 * rr + j][i] = fma2(d[i], win[j][rr][i], dT[3 * rr + j][i]);
#pragma unroll
          for (int rr = 0; rr < 3; ++rr)
#pragma unroll
            for (int i = 0; i < 4; ++i) win[0][rr][i] = win[2][rr][i];
        }
        continue;
      }
      col(w_begin - 1, win[0]);
      col(w_begin, win[1]);
      for (int w0 = w_begin; w0 < w_end; w0 += 3) {
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const int w = w0 + t;
          if (w < w_end) {
            // slots: column w - 1 -> t % 3, w -> (t + 1) % 3, w + 1 -> (t + 2) % 3   (w0 - w_begin is a multiple of 3)
            col(w + 1, win[(t + 2) % 3]);
            f2 d[4];
            unpack(lds16(gb + (uint32_t)w * colb), d);
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
              for (int rr = 0; rr < 3; ++rr)
#pragma unroll
                for (int i = 0; i < 4; ++i) dT[3 * rr + j][i] = fma2(d[i], win[(t + j) % 3][rr][i], dT[3 * rr + j][i]);
          }
        }
      }
    }
  }
  if (WGRAD) {
    float* out = part + (size_t)blockIdx.x * 9 * ld;
    if (g.red) {
      // block reduction through the (now idle) band ring: every thread stores its 72 sums as plain 16-byte stores, [slot][9][ld], then
      // thread i adds the slots of entry i.  The 72 fp32 shared-memory atomics per thread this replaces are compare-and-swap loops:
      // they were most of the ~17 us this launch cost at ANY batch size (tools/sweep_bench.py --batch 32 .. 512)
      __syncthreads();                                                 // the last band has been consumed by everybody
      float4* red = reinterpret_cast<float4*>(smem);
      if (lane_ok) {
#pragma unroll
        for (int a = 0; a < 9; ++a) {
          float4* dst = red + ((size_t)(slot * 9 + a) * ld + c * 8) / 4;
          dst[0] = make_float4(lo_of(dT[a][0]), hi_of(dT[a][0]), lo_of(dT[a][1]), hi_of(dT[a][1]));
          dst[1] = make_float4(lo_of(dT[a][2]), hi_of(dT[a][2]), lo_of(dT[a][3]), hi_of(dT[a][3]));
        }
      }
      __syncthreads();
      const float* rf = reinterpret_cast<const float*>(smem);
      const int slots = g.BR * g.NSEG, n = 9 * ld;
      for (int i = tid; i < n; i += blockDim.x) {
        float sum = 0.f;
        for (int sl = 0; sl < slots; ++sl) sum += rf[(size_t)sl * n + i];
        out[i] = sum;
      }
    } else {
      if (lane_ok) {
#pragma unroll
        for (int a = 0; a < 9; ++a)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            atomicAdd(sacc + a * ld + c * 8 + 2 * i, lo_of(dT[a][i]));
            atomicAdd(sacc + a * ld + c * 8 + 2 * i + 1, hi_of(dT[a][i]));
          }
      }
      __syncthreads();
      for (int i = tid; i < 9 * ld; i += blockDim.x) out[i] = sacc[i];
    }
  }
}

// Band / thread geometry; false when the shape does not fit (the caller keeps the walking kernels of dwsep_rows.cu)
static bool plan(int batch, int H, int W, int ld, int S, bool wgrad, SGeom* g, int* threads, size_t* smem) {
  const int cpr = ld / 8;
  const int cap = wgrad ? kWgradThreads : kFwdThreads;
  const int Ho = H / S, Wo = W / S;
  const long long row_bytes = (long long)W * ld * 2, grow_bytes = (long long)Wo * ld * 2;
  if (cpr > cap || row_bytes > 64 * 1024 || row_bytes % 16 || grow_bytes % 16) return false;
  const long long budget = 200 * 1024 - (wgrad ? 9LL * ld * 4 : 0) - 64;
  // The tallest band with at least two ring stages wins, measured over config 3's shapes (tools/sweep_bench.py, gpurun_out/r4c, r4d):
  // a taller band re-reads fewer halo rows ((S (BR - 1) + 3) input rows per S BR), gives every thread a longer run of columns (two
  // columns of a segment are recomputed halo) and keeps the block at the thread cap on narrow maps -- a 1-row band of a stride-2
  // sweep ran 120 threads per SM (75 us at 56^2 against 41 us with 4-row bands).  Ragged last bands idle their threads, so bands that
  // divide the map are preferred (7 rows at 28^2 / 14^2 / 7^2).  Then the deepest ring that still fits (2..4 stages).
  int BR = 0, nst = 0;
  double best = 1e30;
  for (int br = Ho < 8 ? Ho : 8; br >= 1; --br) {
    const long long stage = (S * (br - 1) + 3) * row_bytes + (wgrad ? br * grow_bytes : 0);
    if (2 * stage > budget || cpr * br > cap) continue;
    const int nb = (Ho + br - 1) / br;
    const double score = (double)(nb * br) / Ho * (double)(S * (br - 1) + 3) / (S * br);
    if (score < best - 1e-9) { best = score; BR = br; nst = (int)(budget / stage < 4 ? budget / stage : 4); }
  }
  if (!BR) return false;
  // development switches (tools/sweep_bench.py): band height / ring depth / minimum segment length of the plan
  int min_sl = 3;
  if (const char* e = getenv("TLT_DWST_PLAN")) {
    int br = 0, ns = 0, sl = 0;
    if (sscanf(e, "%d,%d,%d", &br, &ns, &sl) == 3 && br >= 1 && ns >= 2 && ns <= 4 && sl >= 1 && cpr * br <= cap &&
        (long long)ns * ((S * (br - 1) + 3) * row_bytes + (wgrad ? br * grow_bytes : 0)) <= budget) {
      BR = br; nst = ns; min_sl = sl;
    }
  }
  // segments: as many threads as the cap allows, at least 3 columns per segment (two columns of every segment are recomputed halo)
  const int max_seg = cap / (cpr * BR);
  int SL = (Wo + max_seg - 1) / max_seg;
  if (SL < min_sl) SL = Wo < min_sl ? Wo : min_sl;
  const int nseg = (Wo + SL - 1) / SL;
  g->batch = batch; g->H = H; g->W = W; g->cpr = cpr;
  g->S = S; g->Ho = Ho; g->Wo = Wo; g->grow_bytes = (uint32_t)grow_bytes;
  g->BR = BR; g->NSEG = nseg; g->SL = SL; g->nst = nst;
  g->bands = (Ho + BR - 1) / BR;
  if ((long long)batch * g->bands > 0x7fffffffLL) return false;
  g->units = batch * g->bands;
  g->row_bytes = (uint32_t)row_bytes;
  g->z_stage = (uint32_t)((S * (BR - 1) + 3) * row_bytes);
  g->g_stage = (uint32_t)(BR * grow_bytes);
  g->red = wgrad && (long long)BR * nseg * 9 * ld * 4 <= (long long)nst * ((long long)g->z_stage + g->g_stage) ? 1 : 0;
  *threads = (cpr * BR * nseg + 31) / 32 * 32;
  *smem = (size_t)nst * ((size_t)g->z_stage + (wgrad ? g->g_stage : 0)) + 32 + (wgrad ? (size_t)9 * ld * 4 : 0);
  return true;
}

// -> 1 launched, 0 shape not covered (caller falls back), < 0 error
int staged_fwd(int batch, int H, int W, int ld, int stride, const void* z, const float* taps, void* y, cudaStream_t st) {
  SGeom g; int threads; size_t smem;
  if (!plan(batch, H, W, ld, stride, false, &g, &threads, &smem)) return 0;
  if (ensure_smem((const void*)dwsep_staged_kernel<false>, smem) != TLT_OK) return 0;
  const int resident = (smem * 2 + 2048 <= 227 * 1024 ? 2 : 1) * device_sms();     // two blocks per SM when their slabs fit side by side
  const int grid = g.units < resident ? g.units : resident;
  dwsep_staged_kernel<false><<<grid, threads, smem, st>>>(g, (const char*)z, nullptr, taps, (uint4*)y, nullptr);
  return 1;
}
int staged_wgrad(int batch, int H, int W, int ld, int stride, const void* z, const void* gy, float* acc, cudaStream_t st) {
  SGeom g; int threads; size_t smem;
  if (!plan(batch, H, W, ld, stride, true, &g, &threads, &smem)) return 0;
  if (ensure_smem((const void*)dwsep_staged_kernel<true>, smem) != TLT_OK) return 0;
  dwsep_staged_kernel<true><<<device_sms(), threads, smem, st>>>(g, (const char*)z, (const char*)gy, nullptr, nullptr, acc);
  return 1;
}

}  // namespace dwst
}  // namespace tlt
