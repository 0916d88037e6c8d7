// Pointwise (1x1) channel mixing of channels-first activations on the tcgen05 tensor cores, straight from NCHW memory.
//
//   forward / data-gradient :  Y[b, n, s] = sum_k Wt[n, k] * X[b, k, s] (+ bias[n])          X (B, K, S), Y (B, N, S) bf16
//   weight gradient         :  dW[n, k]  = sum_{b, s} dY[b, n, s] * X[b, k, s]                dW fp32 (N, K)
//
// Replaces the F.conv1d(kernel_size=1) calls of tltorch/functional/convolution.py (:154,168 tucker_conv, :207,223 tt_conv,
// :262,278 cp_conv, :311,340 cp_conv_mobilenet) and their autograd.  The generic strided contraction (SIMT FFMA) ran these
// at ~7 TFLOP/s -- 1 ms for a 214 MB pass at ResNet 56x56 shapes whose HBM floor is 33 us -- because the spatial index, not
// the channel, is the contiguous one.  Here the activation tile is the MN-MAJOR A operand of the MMA (rows = spatial
// positions, exactly as they lie in memory): 3-D TMA boxes {64 s, 64 channels, 1 image} land in SWIZZLE_128B shared
// memory, tcgen05.mma.cta_group::2 (M = 256 positions per CTA pair) accumulates in TMEM, and the epilogue transposes each
// 32 x 32 block through shared memory so that a 3-D TMA store writes Y rows that are contiguous in s.  Out-of-range
// rows / channels / positions are zero-filled on load and clipped on store by the TMA unit, so ranks such as 69 or 573 need
// no padding passes over the activations (only the tiny weight matrix is repacked K-major by the caller).
// The weight gradient is the same pipeline with the reduction running over (image, position) blocks, split across all
// SM pairs, partial sums meeting in dW through TMA reduce-add.
// Requirement: S % 8 == 0 (TMA global strides are multiples of 16 bytes); other shapes stay on tlt_contract.
#include "tc_common.cuh"

namespace tlt {

constexpr int kPwThreads = 384;
constexpr int kPwEpiWarps = 8;
constexpr int kPwEpiBuf = 4096;
// the pair kernel runs 16 epilogue warps (4 per TMEM lane quarter, each a quarter of the tile's columns): its per-unit epilogue --
// tcgen05.ld, convert, transposing stores -- is instruction-issue bound, and 8 warps left two of them per scheduler
constexpr int kPw2EpiWarps = 16;
constexpr int kPw2Threads = 128 + 32 * kPw2EpiWarps;

struct PwParams {
  int S, K, N, batch;            // positions per image, input channels, output channels, images
  int tiles_m, tiles_n;          // fwd: position tiles per image / channel tiles;  wgrad: N tiles / K tiles
  int kpi;                       // wgrad: 64-position blocks per image
  int splits, kb_per_split;      // wgrad
  int has_bias, bias_bf16;
  int c_reduce;
  int w_res, nst;                // forward: weight tile resident in the memory of the last ring stages; ring stages in use
  int tail16;                    // forward: k16 steps of the LAST k block (4 = a full block)
  int dbg;                       // development only: 1 = epilogue drains TMEM but stores nothing, 2 = no TMA store (always 0 in the product)
};

template <int BN, int kStages>
struct PwSmem {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = (BN / 2) * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_OFFSET = kStages * STAGE_BYTES;
  static constexpr int BAR_OFFSET = EPI_OFFSET + kPw2EpiWarps * kPwEpiBuf;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
};

template <int BN, int kStages, bool WGRAD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPw2Threads, 1)
pointwise_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_at, const PwParams p,
                    const void* __restrict__ bias) {
  using L = PwSmem<BN, kStages>;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  constexpr int HALF_N = BN / 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + L::BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_slot_gen = (volatile uint32_t*)(smem_gen + L::BAR_OFFSET + 8 * (2 * kStages + 4));
  const uint32_t wfull_bar = bar_base + 8u * (2 * kStages + 5);
  // forward only: the weight tile is the same for every unit when there is one channel tile, so it is loaded ONCE into the
  // memory of the last ring stages instead of riding in every stage (it was a third of the TMA traffic of a K = 64 layer)
  const bool w_res = !WGRAD && p.w_res;
  const int nst = w_res ? p.nst : kStages;
  const uint32_t wres = smem_base + (uint32_t)nst * L::STAGE_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_units = WGRAD ? p.tiles_m * p.tiles_n * p.splits : p.batch * p.tiles_m * p.tiles_n;
  const int kb_total = WGRAD ? p.batch * p.kpi : (p.K + BLOCK_K - 1) / BLOCK_K;

  // unit -> (image / split, row tile, column tile) and its k-block range
  auto decode = [&](int u, int& z, int& tm, int& tn, int& kb0, int& kb1) {
    if (WGRAD) {
      const int tile = u / p.splits;
      z = u - tile * p.splits;
      tm = tile / p.tiles_n; tn = tile - tm * p.tiles_n;
      kb0 = z * p.kb_per_split; kb1 = min(kb_total, kb0 + p.kb_per_split);
    } else {
      tn = u % p.tiles_n;
      const int r = u / p.tiles_n;
      tm = r % p.tiles_m; z = r / p.tiles_m;
      kb0 = 0; kb1 = kb_total;
    }
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_c);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * kPw2EpiWarps); }
    mbar_init(wfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ===================================== TMA producer (both CTAs) =====================================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      if (w_res) {
        if (leader) mbar_expect_tx(wfull_bar, 2 * kb_total * L::B_BYTES);
        for (int kb = 0; kb < kb_total; ++kb)
          tma_load_2d_2sm(wres + kb * L::B_BYTES, &map_b, wfull_bar, kb * BLOCK_K, (int)cta_rank * HALF_N);
      }
      for (int u = cluster_id; u < num_units; u += num_clusters) {
        int z, tm, tn, kb0, kb1;
        decode(u, z, tm, tn, kb0, kb1);
        const int m0 = tm * (2 * BLOCK_M) + (int)cta_rank * BLOCK_M;
        const int n0 = tn * BN + (int)cta_rank * HALF_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * L::STAGE_BYTES, sb = sa + L::A_BYTES;
          // the last k block of a ragged K (69, 141 channels) only brings the k16 steps that exist: {64 s, 16 k} boxes
          const int n16 = (!WGRAD && kb == kb_total - 1) ? p.tail16 : BLOCK_K / UMMA_K;
          if (leader) {
            if (WGRAD) mbar_expect_tx(full_bar(stage), 2 * L::STAGE_BYTES);
            else mbar_expect_tx(full_bar(stage), 2 * ((BLOCK_M / 64) * n16 * (UMMA_K * 128) + (w_res ? 0 : L::B_BYTES)));
          }
          if (WGRAD) {
            const int img = kb / p.kpi, s0 = (kb - img * p.kpi) * BLOCK_K;
            tma_load_3d_2sm(sa, &map_a, full_bar(stage), s0, m0, img);            // dY: {64 s, 128 out-channels}
            tma_load_3d_2sm(sb, &map_b, full_bar(stage), s0, n0, img);            // X : {64 s, BN/2 in-channels}
          } else {
            const int k0 = kb * BLOCK_K;
            if (n16 == BLOCK_K / UMMA_K) {
#pragma unroll
              for (int c = 0; c < BLOCK_M / 64; ++c)                               // X^T: {64 s, 64 channels} per 64-row chunk
                tma_load_3d_2sm(sa + c * (BLOCK_K * 128), &map_a, full_bar(stage), m0 + 64 * c, k0, z);
            } else {
              for (int c = 0; c < BLOCK_M / 64; ++c)
                for (int j = 0; j < n16; ++j)
                  tma_load_3d_2sm(sa + c * (BLOCK_K * 128) + j * (UMMA_K * 128), &map_at, full_bar(stage), m0 + 64 * c, k0 + UMMA_K * j, z);
            }
            if (!w_res) tma_load_2d_2sm(sb, &map_b, full_bar(stage), k0, n0);      // Wt: {64 k, BN/2 n}
          }
          if (++stage == nst) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA) ======================================
    if (leader && lane == 0) {
      constexpr int a_mn = WGRAD ? 0 : 1;
      const uint32_t idesc = make_idesc(2 * BLOCK_M, BN, a_mn, 0);
      const uint32_t a_lbo = a_mn ? BLOCK_K * 128 : 16;
      const uint32_t a_kstep = a_mn ? (UMMA_K * 128) : (UMMA_K * 2);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      if (w_res) {
        mbar_wait(wfull_bar, 0);
        tc_fence_after();
      }
      for (int u = cluster_id; u < num_units; u += num_clusters) {
        int z, tm, tn, kb0, kb1;
        decode(u, z, tm, tn, kb0, kb1);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * L::STAGE_BYTES;
          const uint32_t sb = w_res ? wres + kb * L::B_BYTES : sa + L::A_BYTES;
          const int n16 = (!WGRAD && kb == kb_total - 1) ? p.tail16 : BLOCK_K / UMMA_K;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            if (k < n16) {
              const uint64_t da = make_smem_desc(sa + k * a_kstep, a_lbo, 1024);
              const uint64_t db = make_smem_desc(sb + k * (UMMA_K * 2), 16, 1024);
              umma_bf16_2sm(tmem_d, da, db, idesc, (kb > kb0 || k != 0) ? 1u : 0u);
            }
          }
          umma_commit_2sm(empty_bar(stage));
          if (kb == kb1 - 1) umma_commit_2sm(tfull_bar(acc));
          if (++stage == nst) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================================== epilogue (both CTAs, 8 warps) ================================
    const int quarter = warp & 3;                  // TMEM lanes [32*quarter, +32)
    constexpr int EPI_N = BN / (kPw2EpiWarps / 4);  // accumulator columns per epilogue warp
    const int col_half = (warp - 4) >> 2;          // accumulator columns [EPI_N*col_half, +EPI_N)
    const uint32_t leader_tempty0 = mapa_u32(tempty_bar(0), 0);
    const uint32_t ebuf = smem_base + L::EPI_OFFSET + (warp - 4) * kPwEpiBuf;
    constexpr int n_chunks = EPI_N / 32;
    int acc = 0; uint32_t acc_phase = 0;
    uint32_t ehalf = 0;                             // bf16 path: the 4 KB staging buffer is two 2 KB halves used alternately
    for (int u = cluster_id; u < num_units; u += num_clusters) {
      int z, tm, tn, kb0, kb1;
      decode(u, z, tm, tn, kb0, kb1);
      const int m0 = tm * (2 * BLOCK_M) + (int)cta_rank * BLOCK_M + quarter * 32;     // first accumulator row of this warp
      const int n0 = tn * BN + col_half * EPI_N;
      const int m_lim = WGRAD ? p.N : p.S, n_lim = WGRAD ? p.K : p.N;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + col_half * EPI_N);
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c) {
        const int col0 = n0 + c * 32;
        const bool live = m0 < m_lim && col0 < n_lim;            // warp-uniform
        uint32_t v[32];
        if (live) {
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
        }
        if (c == n_chunks - 1) {                                  // accumulator fully read: hand the TMEM buffer back early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(leader_tempty0 + 8u * acc);
        }
        if (!live || (p.dbg & 1)) continue;
        // WGRAD: the previous store of this warp has finished reading ebuf.  Forward: the store issued TWO chunks ago (same
        // half) has; the latest one may still be in flight, so the transposition of this chunk overlaps its read
        if (lane == 0) { if (WGRAD) bulk_wait_read0(); else bulk_wait_read1(); }
        __syncwarp();
        if (WGRAD) {
          // fp32 [32 rows][32 cols] block, SWIZZLE_128B, reduce-added into dW
          const uint32_t erow = ebuf + lane * 128;
          const uint32_t sw = (uint32_t)(lane & 7);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_v4(erow + ((((uint32_t)j) ^ sw) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if (p.c_reduce) tma_reduce_add_2d(&map_c, ebuf, col0, m0);
            else tma_store_2d(&map_c, ebuf, col0, m0);
            bulk_commit();
          }
        } else {
          // transpose: lane holds position s = m0 + lane, v[i] is channel col0 + i  ->  smem [32 channels][32 positions] bf16
          if (p.has_bias) {                                       // warp-uniform; the common no-bias launch skips 64 instructions
            float bv[32];
            load_bias32(bias, p.bias_bf16, col0, p.N, bv);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + bv[i]);
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const __nv_bfloat16 h = __float2bfloat16_rn(__uint_as_float(v[i]));
            st_shared_b16(ebuf + ehalf + i * 64 + lane * 2, *(const uint16_t*)&h);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if (!(p.dbg & 2)) tma_store_3d(&map_c, ebuf + ehalf, m0, col0, z);
            bulk_commit();
          }
          ehalf ^= 2048u;
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) bulk_wait_read0();
    __syncwarp();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Transposed-role forward / data-gradient kernel: D[out channel][position] = Wt[out channel][k] . X[k][position].
// The kernel above puts positions on the TMEM lanes, so a thread of the epilogue holds ONE position and 32 channels and has to
// transpose every 32 x 32 block through shared memory (32 two-byte stores + a TMA store per block: 22.7 M warp instructions,
// 63 us per launch at 56^2 x 64 -> 69, 2.6 TB/s, epilogue-bound).  Here the WEIGHT tile is the A operand (M = 128 output
// channels on the lanes) and the activation tile is the MN-major B operand (N = 256 positions, 3-D TMA boxes {64 s, 64 k}
// straight from NCHW memory): a thread of the epilogue holds one channel and 32 consecutive positions = 64 contiguous bytes
// of y, which leave as four 16-byte global stores -- no staging, no fences, no TMA store.  One CTA per SM (cta_group::1),
// persistent over (image, position tile, channel tile) units, 4-stage TMA ring, double-buffered TMEM accumulator (2 x 256 cols).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kTctStages = 4;
constexpr int kTctX = 4 * BLOCK_K * 128;                 // 256 positions x 64 channels bf16 = 32 KB
constexpr int kTctW = BLOCK_M * BLOCK_K * 2;             // 128 output channels x 64 k = 16 KB
constexpr int kTctStage = kTctX + kTctW;
constexpr int kTctEpi = kTctStages * kTctStage;                      // 8 epilogue warps x 4 KB staging for the TMA stores
constexpr int kTctBar = kTctEpi + kPwEpiWarps * kPwEpiBuf;
constexpr int kTctSmem = kTctBar + 256 + 1024;

__global__ void __launch_bounds__(kPwThreads, 1)
pointwise_tct_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                     const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_xt, const PwParams p,
                     const void* __restrict__ bias) {
  constexpr int kStages = kTctStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + kTctBar;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  const uint32_t wfull_bar = bar_base + 8u * (2 * kStages + 5);
  // one channel tile and K <= 64 kStages: every unit uses the same weight blocks, so block kb is loaded ONCE into the weight slot of
  // ring stage kb (unused otherwise in this mode) instead of riding in every stage
  const bool w_res = p.w_res != 0;
  volatile uint32_t* tmem_slot_gen = (volatile uint32_t*)(smem_gen + kTctBar + 8 * (2 * kStages + 4));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_units = p.batch * p.tiles_m * p.tiles_n;
  const int kb_total = (p.K + BLOCK_K - 1) / BLOCK_K;
  auto decode = [&](int u, int& z, int& ts, int& tn) {
    tn = u % p.tiles_n;
    const int r = u / p.tiles_n;
    ts = r % p.tiles_m; z = r / p.tiles_m;
  };
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_w); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), kPwEpiWarps); }
    mbar_init(wfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      if (w_res) {
        mbar_expect_tx(wfull_bar, kb_total * kTctW);
        for (int kb = 0; kb < kb_total; ++kb)
          tma_load_2d(smem_base + kb * kTctStage + kTctX, &map_w, wfull_bar, kb * BLOCK_K, 0);
      }
      for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
        int z, ts, tn;
        decode(u, z, ts, tn);
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sx = smem_base + stage * kTctStage, sw = sx + kTctX;
          const int n16 = kb == kb_total - 1 ? p.tail16 : BLOCK_K / UMMA_K;        // ragged K: only the k16 steps that exist
          mbar_expect_tx(full_bar(stage), 4 * n16 * (UMMA_K * 128) + (w_res ? 0 : kTctW));
          if (n16 == BLOCK_K / UMMA_K) {
#pragma unroll
            for (int c = 0; c < 4; ++c)                                            // X: {64 s, 64 k} per 64-position chunk
              tma_load_3d(sx + c * (BLOCK_K * 128), &map_x, full_bar(stage), ts * 256 + 64 * c, kb * BLOCK_K, z);
          } else {
            for (int c = 0; c < 4; ++c)
              for (int j = 0; j < n16; ++j)                                        // {64 s, 16 k} boxes
                tma_load_3d(sx + c * (BLOCK_K * 128) + j * (UMMA_K * 128), &map_xt, full_bar(stage), ts * 256 + 64 * c,
                            kb * BLOCK_K + UMMA_K * j, z);
          }
          if (!w_res) tma_load_2d(sw, &map_w, full_bar(stage), kb * BLOCK_K, tn * BLOCK_M);    // Wt: {64 k, 128 n}
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BLOCK_M, 256, 0, 1);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      if (w_res) {
        mbar_wait(wfull_bar, 0);
        tc_fence_after();
      }
      for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 256);
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sx = smem_base + stage * kTctStage;
          const uint32_t sw = w_res ? smem_base + kb * kTctStage + kTctX : sx + kTctX;
          const int n16 = kb == kb_total - 1 ? p.tail16 : BLOCK_K / UMMA_K;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            if (k < n16) {
              const uint64_t da = make_smem_desc(sw + k * (UMMA_K * 2), 16, 1024);             // Wt, K-major
              const uint64_t db = make_smem_desc(sx + k * (UMMA_K * 128), BLOCK_K * 128, 1024); // X, MN-major, 4 chunks of 64
              umma_bf16(tmem_d, da, db, idesc, (kb > 0 || k != 0) ? 1u : 0u);
            }
          }
          umma_commit(empty_bar(stage));
          if (kb == kb_total - 1) umma_commit(tfull_bar(acc));
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int quarter = warp & 3;                  // TMEM lanes [32 quarter, +32) = output channels
    const int half = (warp - 4) >> 2;              // accumulator columns [128 half, +128) = positions
    const uint32_t ebuf = smem_base + kTctEpi + (warp - 4) * kPwEpiBuf;
    int acc = 0; uint32_t acc_phase = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      int z, ts, tn;
      decode(u, z, ts, tn);
      const int n = tn * BLOCK_M + quarter * 32 + lane;
      const bool warp_live = tn * BLOCK_M + quarter * 32 < p.N;
      float bv = 0.f;
      if (p.has_bias && n < p.N) bv = p.bias_bf16 ? __bfloat162float(((const __nv_bfloat16*)bias)[n]) : ((const float*)bias)[n];
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 256 + half * 128);
      const int n0 = tn * BLOCK_M + quarter * 32;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int s0 = ts * 256 + half * 128 + c * 32;
        const bool live = warp_live && s0 < p.S;                 // warp-uniform
        uint32_t v[32];
        if (live) {
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
        }
        if (c == 3) {                                            // accumulator fully read: hand the TMEM buffer back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
        }
        if (!live) continue;
        // this lane's channel row: 32 positions = 64 bytes = half of a 128-byte row of the [32 channels][64 positions] staging block
        // (SWIZZLE_128B: 16-byte chunk index ^ (row & 7), conflict-free for the 32 lanes); two chunks leave as ONE TMA store
        if ((c & 1) == 0) {
          if (lane == 0) bulk_wait_read0();                      // the previous store has read the block
          __syncwarp();
        }
        const uint32_t erow = ebuf + lane * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_shared_v4(erow + ((((uint32_t)((c & 1) * 4 + j)) ^ (uint32_t)(lane & 7)) << 4),
                       pack_bf16x2(__uint_as_float(v[8 * j + 0]) + bv, __uint_as_float(v[8 * j + 1]) + bv),
                       pack_bf16x2(__uint_as_float(v[8 * j + 2]) + bv, __uint_as_float(v[8 * j + 3]) + bv),
                       pack_bf16x2(__uint_as_float(v[8 * j + 4]) + bv, __uint_as_float(v[8 * j + 5]) + bv),
                       pack_bf16x2(__uint_as_float(v[8 * j + 6]) + bv, __uint_as_float(v[8 * j + 7]) + bv));
        if ((c & 1) == 1 || s0 + 32 >= p.S) {
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&map_c, ebuf, s0 - (c & 1) * 32, n0, z);  // rows >= N and positions >= S are clipped by the tensor map
            bulk_commit();
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) bulk_wait_read0();
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Single-CTA weight gradient for small channel counts (N <= 128 per tile, K <= 128):  dW[n][k] += sum_{b,s} dY[b,n,s] X[b,k,s].
// In the pair kernel above the two CTAs split the M (output-channel) and N (input-channel) extents of the tile; with 69 x 64
// channels (config 3's 56^2 layers) everything the second CTA loads is out of bounds, so only 74 of the 148 SMs pull data
// (214 MB at 3.2 TB/s).  Here every CTA is its own cta_group::1 MMA (M = 128 lanes of output channels, N = NB input channels),
// owns a slice of the (image, 64-position block) reduction range, accumulates it in ONE TMEM buffer and adds its fp32 partial into
// dW with TMA reduce-add -- all 148 SMs stream.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kW1Stages = 6;
template <int NB> struct W1Smem {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;      // dY {64 s, 128 n}
  static constexpr int B_BYTES = NB * BLOCK_K * 2;           // X  {64 s, NB k}
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int EPI_OFFSET = kW1Stages * STAGE;
  static constexpr int BAR_OFFSET = EPI_OFFSET + kPwEpiWarps * 4096;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
};

template <int NB>
__global__ void __launch_bounds__(kPwThreads, 1)
pointwise_wgrad1_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                        const __grid_constant__ CUtensorMap map_c, const PwParams p) {
  using L = W1Smem<NB>;
  constexpr int kStages = kW1Stages;
  constexpr uint32_t TMEM_COLS = NB < 32 ? 32 : NB;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + L::BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * kStages), tempty_bar = bar_base + 8u * (2 * kStages + 1);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 2);
  volatile uint32_t* tmem_slot_gen = (volatile uint32_t*)(smem_gen + L::BAR_OFFSET + 8 * (2 * kStages + 2));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_units = p.tiles_m * p.splits;
  const int kb_total = p.batch * p.kpi;
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_b); tma_prefetch_desc(&map_c); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, kPwEpiWarps);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
        const int tm = u / p.splits, z = u - tm * p.splits;
        const int kb0 = z * p.kb_per_split, kb1 = min(kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * L::STAGE, sb = sa + L::A_BYTES;
          mbar_expect_tx(full_bar(stage), L::STAGE);
          const int img = kb / p.kpi, s0 = (kb - img * p.kpi) * BLOCK_K;
          tma_load_3d(sa, &map_a, full_bar(stage), s0, tm * BLOCK_M, img);          // dY: {64 s, 128 out-channels}
          tma_load_3d(sb, &map_b, full_bar(stage), s0, 0, img);                     // X : {64 s, NB in-channels}
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BLOCK_M, NB, 0, 0);
      int stage = 0; uint32_t phase = 0, acc_phase = 0;
      for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
        const int tm = u / p.splits, z = u - tm * p.splits;
        const int kb0 = z * p.kb_per_split, kb1 = min(kb_total, kb0 + p.kb_per_split);
        mbar_wait(tempty_bar, acc_phase ^ 1);
        tc_fence_after();
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * L::STAGE, sb = sa + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc(sa + k * (UMMA_K * 2), 16, 1024);
            const uint64_t db = make_smem_desc(sb + k * (UMMA_K * 2), 16, 1024);
            umma_bf16(tmem_base, da, db, idesc, (kb > kb0 || k != 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          if (kb == kb1 - 1) umma_commit(tfull_bar);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int quarter = warp & 3;                  // TMEM lanes [32 quarter, +32) = output channels
    const int half = (warp - 4) >> 2;              // accumulator columns [NB / 2 * half, +NB / 2) = input channels
    const uint32_t ebuf = smem_base + L::EPI_OFFSET + (warp - 4) * 4096;
    constexpr int n_chunks = NB / 2 / 32;
    uint32_t acc_phase = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      const int tm = u / p.splits;
      const int m0 = tm * BLOCK_M + quarter * 32;
      mbar_wait(tfull_bar, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * (NB / 2));
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c) {
        const int col0 = half * (NB / 2) + c * 32;
        const bool live = m0 < p.N && col0 < p.K;                // warp-uniform
        uint32_t v[32];
        if (live) {
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
        }
        if (c == n_chunks - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar);
        }
        if (!live) continue;
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
        const uint32_t erow = ebuf + lane * 128;
        const uint32_t sw = (uint32_t)(lane & 7);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(erow + ((((uint32_t)j) ^ sw) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (p.c_reduce) tma_reduce_add_2d(&map_c, ebuf, col0, m0);
          else tma_store_2d(&map_c, ebuf, col0, m0);
          bulk_commit();
        }
      }
      acc_phase ^= 1;
    }
    if (lane == 0) bulk_wait_read0();
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int NB>
static int w1_launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const PwParams& p, int units, cudaStream_t st) {
  using L = W1Smem<NB>;
  TLT_ENSURE_SMEM((pointwise_wgrad1_kernel<NB>), L::TOTAL);
  const int sms = num_sms();
  pointwise_wgrad1_kernel<NB><<<units < sms ? units : sms, kPwThreads, L::TOTAL, st>>>(ma, mb, mc, p);
  return check_launch("pointwise_wgrad1_kernel");
}

template <int BN, int kStages, bool WGRAD>
static int pw_launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mat, const PwParams& p,
                     const void* bias, int units, cudaStream_t st) {
  using L = PwSmem<BN, kStages>;
  TLT_ENSURE_SMEM((pointwise_tc_kernel<BN, kStages, WGRAD>), L::TOTAL);
  const int pairs = num_sms() / 2;
  const int clusters = units < pairs ? units : pairs;
  pointwise_tc_kernel<BN, kStages, WGRAD><<<2 * clusters, kPw2Threads, L::TOTAL, st>>>(ma, mb, mc, mat, p, bias);
  return check_launch("pointwise_tc_kernel");
}

template <typename T>
__global__ void pack_kmajor_kernel(const T* __restrict__ src, long long rows, long long cols, long long rs, long long cs,
                                   __nv_bfloat16* __restrict__ dst, long long ld) {
  const long long total = rows * ld;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ld, c = i - r * ld;
    dst[i] = c < cols ? __float2bfloat16_rn(to_f32<T>(src[r * rs + c * cs])) : __float2bfloat16_rn(0.f);
  }
}

// Row-contiguous sources (col_stride == 1, e.g. the 3375 x 3375 Tucker core re-pitched to 3376): one block row per matrix
// row, a thread assembles 8 consecutive elements and writes one 16-byte chunk -- no 64-bit division per element (the
// generic kernel above took 35 us for the 23 MB core of BASELINE config 5).
template <typename T>
__global__ void __launch_bounds__(256) pack_kmajor_rows_kernel(const T* __restrict__ src, long long rows, int cols, long long rs,
                                                               __nv_bfloat16* __restrict__ dst, int ld) {
  const int chunks = ld >> 3;
  for (long long r = blockIdx.y; r < rows; r += gridDim.y) {
    const T* sp = src + r * rs;
    uint4* dp = reinterpret_cast<uint4*>(dst + r * (long long)ld);
    for (int c8 = blockIdx.x * blockDim.x + threadIdx.x; c8 < chunks; c8 += gridDim.x * blockDim.x) {
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c8 * 8 + 2 * i;
        const float lo = c < cols ? to_f32<T>(sp[c]) : 0.f, hi = c + 1 < cols ? to_f32<T>(sp[c + 1]) : 0.f;
        __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
        w[i] = *reinterpret_cast<uint32_t*>(&h);
      }
      dp[c8] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

static int pw_check(const tlt_pointwise_desc* d, const char* who) {
  TLT_REQUIRE(d != nullptr, "%s: null descriptor", who);
  TLT_REQUIRE(d->batch >= 1 && d->in_channels >= 1 && d->out_channels >= 1 && d->spatial >= 1, "%s: sizes must be >= 1", who);
  TLT_REQUIRE(d->batch < (1LL << 24) && d->in_channels < (1LL << 24) && d->out_channels < (1LL << 24) && d->spatial < (1LL << 28),
              "%s: sizes too large", who);
  if (d->spatial % 8) {
    set_error("%s: spatial size %lld is not a multiple of 8 (TMA pitch); use tlt_contract", who, (long long)d->spatial);
    return TLT_ERR_UNSUPPORTED;
  }
  return TLT_OK;
}

}  // namespace tlt

using namespace tlt;

extern "C" int tlt_pointwise_tc_supported(const tlt_pointwise_desc* d) {
  return d && d->batch >= 1 && d->in_channels >= 1 && d->out_channels >= 1 && d->spatial >= 8 && d->spatial % 8 == 0 &&
         d->w_ld % 8 == 0 && d->w_ld >= d->in_channels;
}

extern "C" int tlt_pointwise_tc(const tlt_pointwise_desc* d, const void* x, const void* wt, const void* bias, void* y, void* stream) {
  int rc = pw_check(d, "tlt_pointwise_tc");
  if (rc) return rc;
  TLT_REQUIRE(x && wt && y, "tlt_pointwise_tc: null pointer");
  TLT_REQUIRE(d->w_ld % 8 == 0 && d->w_ld >= d->in_channels, "tlt_pointwise_tc: w_ld must be a multiple of 8 and >= in_channels");
  TLT_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)wt % 16) == 0 && ((uintptr_t)y % 16) == 0, "tlt_pointwise_tc: operands must be 16-byte aligned");
  const int64_t S = d->spatial, K = d->in_channels, N = d->out_channels, B = d->batch;
  CUtensorMap ma, mb, mc;
  if ((rc = make_map3d(&ma, x, S, K, B, S, K * S, 64, BLOCK_K, true))) return rc;               // X^T chunks {64 s, 64 k}
  const bool wide = N > 128;
  if ((rc = make_map(&mb, wt, K, N, d->w_ld, BLOCK_K, wide ? 128 : 64))) return rc;             // Wt {64 k, BN/2 n}
  if ((rc = make_map3d(&mc, y, S, N, B, S, N * S, 32, 32, false))) return rc;                   // Y {32 s, 32 n}
  PwParams p{};
  p.S = (int)S; p.K = (int)K; p.N = (int)N; p.batch = (int)B;
  p.tiles_m = (int)((S + 2 * BLOCK_M - 1) / (2 * BLOCK_M));
  p.tiles_n = (int)((N + (wide ? 256 : 128) - 1) / (wide ? 256 : 128));
  p.has_bias = bias != nullptr; p.bias_bf16 = d->dtype_bias == TLT_BF16;
  p.dbg = 0;
  const long long units = (long long)p.batch * p.tiles_m * p.tiles_n;
  TLT_REQUIRE(units < (1LL << 31), "tlt_pointwise_tc: too many tiles");
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap mat;
  if ((rc = make_map3d(&mat, x, S, K, B, S, K * S, 64, UMMA_K, true))) return rc;                // X^T tail chunks {64 s, 16 k}
  {
    const int kb_total = (int)((K + BLOCK_K - 1) / BLOCK_K);
    const int ktail = (int)(K - (int64_t)(kb_total - 1) * BLOCK_K);
    p.tail16 = (ktail + UMMA_K - 1) / UMMA_K;
    const char* env = getenv("TLT_PW_RES");
    const bool res_off = env && env[0] == 'o' && env[1] == 'f';
    if (res_off) p.tail16 = BLOCK_K / UMMA_K;
    const int stage_bytes = BLOCK_M * BLOCK_K * 2 + (wide ? 128 : 64) * BLOCK_K * 2, b_bytes = (wide ? 128 : 64) * BLOCK_K * 2;
    const int res_stages = (kb_total * b_bytes + stage_bytes - 1) / stage_bytes;
    const int stages = wide ? 5 : 6;
    p.w_res = (!res_off && p.tiles_n == 1 && stages - res_stages >= 4) ? 1 : 0;
    p.nst = stages - res_stages;
  }
  {
    // Transposed roles (pointwise_tct_kernel): output channels on the TMEM lanes, 256 positions per tile, so a thread of the
    // epilogue owns 32 consecutive positions of ONE channel and y leaves as TMA stores of 128-byte rows ([32 channels][64 positions]
    // blocks) with 16 convert + 4 store instructions per chunk instead of 64 + a 64-byte-row store.  Measured at 56^2 (tools/
    // pointwise_bench.py): 64 -> 69 channels 49 us vs 55 us for the pair kernel; with two channel tiles (N > 128) the activation
    // tile is read twice and the pair kernel wins, on small maps the two are within noise -- hence the shape test.
    // TLT_PW_T = on | off forces the choice.
    const char* env = getenv("TLT_PW_T");
    const bool force_on = env && env[0] == 'o' && env[1] == 'n', force_off = env && env[0] == 'o' && env[1] == 'f';
    if (force_on || (!force_off && N <= BLOCK_M && S >= 2048)) {
      CUtensorMap mw, mc64;
      if ((rc = make_map(&mw, wt, K, N, d->w_ld, BLOCK_K, BLOCK_M))) return rc;                  // Wt {64 k, 128 n}
      if ((rc = make_map3d(&mc64, y, S, N, B, S, N * S, 64, 32, true))) return rc;              // Y {64 s, 32 n}, 128-byte rows
      p.tiles_m = (int)((S + 255) / 256);
      p.tiles_n = (int)((N + BLOCK_M - 1) / BLOCK_M);
      {
        const char* er = getenv("TLT_PW_RES");
        p.w_res = (!(er && er[0] == 'o' && er[1] == 'f') && p.tiles_n == 1 && (K + BLOCK_K - 1) / BLOCK_K <= kTctStages) ? 1 : 0;
      }
      const long long un = (long long)p.batch * p.tiles_m * p.tiles_n;
      TLT_REQUIRE(un < (1LL << 31), "tlt_pointwise_tc: too many tiles");
      TLT_ENSURE_SMEM(pointwise_tct_kernel, kTctSmem);
      const int sms = num_sms();
      pointwise_tct_kernel<<<(unsigned)(un < sms ? un : sms), kPwThreads, kTctSmem, st>>>(ma, mw, mc64, mat, p, bias);
      return check_launch("pointwise_tct_kernel");
    }
  }
  if (wide) return pw_launch<256, 5, false>(ma, mb, mc, mat, p, bias, (int)units, st);
  return pw_launch<128, 6, false>(ma, mb, mc, mat, p, bias, (int)units, st);
}

extern "C" int tlt_pointwise_wgrad_tc(const tlt_pointwise_desc* d, const void* dy, const void* x, float* dw, void* stream) {
  int rc = pw_check(d, "tlt_pointwise_wgrad_tc");
  if (rc) return rc;
  TLT_REQUIRE(dy && x && dw, "tlt_pointwise_wgrad_tc: null pointer");
  TLT_REQUIRE(d->w_ld % 4 == 0 && d->w_ld >= d->in_channels, "tlt_pointwise_wgrad_tc: w_ld (pitch of dW) must be a multiple of 4 and >= in_channels");
  TLT_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)dy % 16) == 0 && ((uintptr_t)dw % 16) == 0, "tlt_pointwise_wgrad_tc: operands must be 16-byte aligned");
  const int64_t S = d->spatial, K = d->in_channels, N = d->out_channels, B = d->batch;
  CUtensorMap ma, mb, mc;
  {
    const char* env = getenv("TLT_PW_W1");
    const bool off = env && env[0] == 'o' && env[1] == 'f';
    if (!off && K <= 128 && N <= 256) {
      // small channel counts: one cta_group::1 MMA per SM, the reduction range split over all SMs
      const int nb = K <= 64 ? 64 : 128;
      if ((rc = make_map3d(&ma, dy, S, N, B, S, N * S, BLOCK_K, BLOCK_M, true))) return rc;      // dY {64 s, 128 n}
      if ((rc = make_map3d(&mb, x, S, K, B, S, K * S, BLOCK_K, nb, true))) return rc;            // X  {64 s, NB k}
      if ((rc = make_map(&mc, dw, K, N, d->w_ld, 32, 32, 1))) return rc;                          // dW fp32 {32 k, 32 n}
      PwParams p{};
      p.S = (int)S; p.K = (int)K; p.N = (int)N; p.batch = (int)B;
      p.tiles_m = (int)((N + BLOCK_M - 1) / BLOCK_M);
      p.tiles_n = 1;
      p.kpi = (int)((S + BLOCK_K - 1) / BLOCK_K);
      const long long kb_total = (long long)p.batch * p.kpi;
      TLT_REQUIRE(kb_total < (1LL << 31), "tlt_pointwise_wgrad_tc: reduction too long");
      long long splits = num_sms() / p.tiles_m;
      if (splits < 1) splits = 1;
      if (splits > kb_total / 4) splits = kb_total / 4 > 0 ? kb_total / 4 : 1;
      p.kb_per_split = (int)((kb_total + splits - 1) / splits);
      p.splits = (int)((kb_total + p.kb_per_split - 1) / p.kb_per_split);
      p.c_reduce = p.splits > 1;
      cudaStream_t st1 = (cudaStream_t)stream;
      if (p.c_reduce) TLT_CHECK_CUDA(cudaMemset2DAsync(dw, (size_t)d->w_ld * 4, 0, (size_t)K * 4, (size_t)N, st1));
      const int units1 = p.tiles_m * p.splits;
      return nb == 64 ? w1_launch<64>(ma, mb, mc, p, units1, st1) : w1_launch<128>(ma, mb, mc, p, units1, st1);
    }
  }
  const bool wide = K > 128;
  if ((rc = make_map3d(&ma, dy, S, N, B, S, N * S, BLOCK_K, BLOCK_M, true))) return rc;          // dY {64 s, 128 n}
  if ((rc = make_map3d(&mb, x, S, K, B, S, K * S, BLOCK_K, wide ? 128 : 64, true))) return rc;   // X  {64 s, BN/2 k}
  if ((rc = make_map(&mc, dw, K, N, d->w_ld, 32, 32, 1))) return rc;                              // dW fp32 {32 k, 32 n}
  PwParams p{};
  p.S = (int)S; p.K = (int)K; p.N = (int)N; p.batch = (int)B;
  p.tiles_m = (int)((N + 2 * BLOCK_M - 1) / (2 * BLOCK_M));
  p.tiles_n = (int)((K + (wide ? 256 : 128) - 1) / (wide ? 256 : 128));
  p.kpi = (int)((S + BLOCK_K - 1) / BLOCK_K);
  const long long kb_total = (long long)p.batch * p.kpi;
  TLT_REQUIRE(kb_total < (1LL << 31), "tlt_pointwise_wgrad_tc: reduction too long");
  const int pairs = num_sms() / 2;
  long long splits = pairs / ((long long)p.tiles_m * p.tiles_n);
  if (splits < 1) splits = 1;
  if (splits > kb_total / 4) splits = kb_total / 4 > 0 ? kb_total / 4 : 1;
  p.kb_per_split = (int)((kb_total + splits - 1) / splits);
  p.splits = (int)((kb_total + p.kb_per_split - 1) / p.kb_per_split);
  p.c_reduce = p.splits > 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (p.c_reduce) TLT_CHECK_CUDA(cudaMemset2DAsync(dw, (size_t)d->w_ld * 4, 0, (size_t)K * 4, (size_t)N, st));
  const int units = p.tiles_m * p.tiles_n * p.splits;
  if (wide) return pw_launch<256, 5, true>(ma, mb, mc, mc, p, nullptr, units, st);
  return pw_launch<128, 6, true>(ma, mb, mc, mc, p, nullptr, units, st);
}

extern "C" int tlt_pack_kmajor(const void* src, int32_t dtype_src, int64_t rows, int64_t cols, int64_t row_stride,
                               int64_t col_stride, void* dst, int64_t dst_ld, void* stream) {
  TLT_REQUIRE(src && dst && rows >= 1 && cols >= 1 && dst_ld >= cols, "tlt_pack_kmajor: bad arguments");
  const long long total = rows * dst_ld;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (col_stride == 1 && dst_ld % 8 == 0 && dst_ld < (1LL << 31) && ((uintptr_t)dst % 16) == 0 && rows * dst_ld >= (1LL << 16) &&
      (dtype_src == TLT_BF16 || dtype_src == TLT_F32)) {
    const int chunks = (int)(dst_ld >> 3);
    dim3 grid((unsigned)((chunks + 255) / 256), (unsigned)(rows < 65535 ? rows : 65535));
    if (dtype_src == TLT_BF16)
      pack_kmajor_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, rows, (int)cols, row_stride, (__nv_bfloat16*)dst, (int)dst_ld);
    else
      pack_kmajor_rows_kernel<float><<<grid, 256, 0, st>>>((const float*)src, rows, (int)cols, row_stride, (__nv_bfloat16*)dst, (int)dst_ld);
    return check_launch("pack_kmajor_rows_kernel");
  }
  if (dtype_src == TLT_BF16)
    pack_kmajor_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)src, rows, cols, row_stride, col_stride, (__nv_bfloat16*)dst, dst_ld);
  else if (dtype_src == TLT_F32)
    pack_kmajor_kernel<float><<<blocks, 256, 0, st>>>((const float*)src, rows, cols, row_stride, col_stride, (__nv_bfloat16*)dst, dst_ld);
  else { set_error("tlt_pack_kmajor: unknown dtype"); return TLT_ERR_INVALID; }
  return check_launch("pack_kmajor_kernel");
}
