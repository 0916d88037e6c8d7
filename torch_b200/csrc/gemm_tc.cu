// tlt_gemm_bf16_tc: dense bf16 GEMM on the Blackwell tcgen05 tensor cores.
//
//   C[M,N] = alpha * A[M,K] * B[N,K]^T (+ bias[N]) (+ C)       A, B bf16 (K-major or MN-major), C bf16 / fp32 row-major
//
// Structure (one persistent CTA per SM, 256 threads, warp-specialised):
//   warp 0     TMA producer   cp.async.bulk.tensor (SWIZZLE_128B boxes) into a kStages-deep smem ring, mbarrier tx
//   warp 1     MMA issuer     one lane issues tcgen05.mma.cta_group::1.kind::f16 (128 x BN x 16), fp32 accumulators in TMEM;
//                             tcgen05.commit releases smem stages / publishes the accumulator
//   warp 2     TMEM allocator tcgen05.alloc / dealloc (2 accumulator buffers so the epilogue of tile i overlaps the
//                             mainloop of tile i+1)
//   warps 4-7  epilogue       tcgen05.ld (32 lanes x 32 columns per warp per step) -> alpha / bias / cast -> global
//
// The hot path uses it for the dense-regime contractions: BlockTT / TT-matrix linear at BERT-FFN shapes (config 2:
// reconstruct W from the cores once per step, then X*W^T, dY*W, dY^T*X), large-rank Tucker core / TT stages (config 5),
// the TRL inner product and the GEMM behind unfolded Tucker-core convolutions.
#include "tc_common.cuh"

namespace tlt {

struct GemmParams {
  int M, N, K;
  long long ldc;
  int a_mn, b_mn;
  int c_f32, accumulate, has_bias, bias_bf16;
  float alpha;
  int tiles_m, tiles_n;
  int splits, kb_per_split, c_reduce;
};


// =====================================================================================================================
// CTA-pair kernel.  A cluster of two CTAs (one TPC) owns a 256 x BN output tile: CTA r stages rows [128r, 128r+128) of
// A and columns [BN/2 r, BN/2 (r+1)) of B, the leader issues tcgen05.mma.cta_group::2 (M = 256) which reads both
// CTAs' shared memory, and each CTA's TMEM receives its own 128 accumulator rows.  Per FLOP this halves the shared-memory
// traffic of a single-CTA kernel (which saturates the 128 B/clk smem port at ~2/3 of tensor peak).
//   warp 0: TMA producer   warp 1: MMA issuer (leader CTA only)   warp 2: TMEM alloc   warps 4-11: epilogue
// =====================================================================================================================
constexpr int kGemm2Threads = 384;
// Epilogue: shared-memory staged -- TMEM -> registers -> SWIZZLE_128B smem -> cp.async.bulk.tensor store (full 128-byte lines, OOB
// rows / columns clipped by the TMA unit); per-thread 16-byte global stores kept l1tex 72 % busy and the tensor pipe 44 % active on
// the N = 3072 / K = 768 launches (profiles/r01_gemm_tc2_full.md: the kernel version this one replaced).  Split-K work units meet
// in C through cp.reduce.async.bulk.tensor (.add), so the K = tokens weight-gradient GEMMs (36 output tiles) fill all 74 SM pairs.
constexpr int kEpiWarps = 8;
constexpr int kEpiBufBytes = 4096;          // 32 rows x 128 bytes per epilogue warp

template <int BN, int kStages>
struct SmemLayout3 {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;        // 16 KB: this CTA's 128 rows
  static constexpr int B_BYTES = (BN / 2) * BLOCK_K * 2;       // this CTA's half of the N extent
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_OFFSET = kStages * STAGE_BYTES;
  static constexpr int BAR_OFFSET = EPI_OFFSET + kEpiWarps * kEpiBufBytes;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
};

template <int BN, int kStages>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemm2Threads, 1)
gemm_bf16_tc3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const __grid_constant__ CUtensorMap map_c, const GemmParams p, const void* __restrict__ bias) {
  using L = SmemLayout3<BN, kStages>;
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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int num_units = p.tiles_m * p.tiles_n * p.splits;
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_c);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * kEpiWarps); }
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
      for (int u = cluster_id; u < num_units; u += num_clusters) {
        const int tile = u / p.splits, split = u - tile * p.splits;
        const int m0 = (tile / p.tiles_n) * (2 * BLOCK_M) + (int)cta_rank * BLOCK_M;
        const int n0 = (tile % p.tiles_n) * BN + (int)cta_rank * HALF_N;
        const int kb0 = split * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * L::STAGE_BYTES, sb = sa + L::A_BYTES;
          if (leader) mbar_expect_tx(full_bar(stage), 2 * L::STAGE_BYTES);
          const int k0 = kb * BLOCK_K;
          if (!p.a_mn) {
            tma_load_2d_2sm(sa, &map_a, full_bar(stage), k0, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BLOCK_M / 64; ++c)
              tma_load_2d_2sm(sa + c * (BLOCK_K * 128), &map_a, full_bar(stage), m0 + 64 * c, k0);
          }
          if (!p.b_mn) {
            tma_load_2d_2sm(sb, &map_b, full_bar(stage), k0, n0);
          } else {
#pragma unroll
            for (int c = 0; c < HALF_N / 64; ++c)
              tma_load_2d_2sm(sb + c * (BLOCK_K * 128), &map_b, full_bar(stage), n0 + 64 * c, k0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA) ======================================
    if (leader && lane == 0) {
      const uint32_t idesc = make_idesc(2 * BLOCK_M, BN, p.a_mn, p.b_mn);
      const uint32_t a_lbo = p.a_mn ? BLOCK_K * 128 : 16, b_lbo = p.b_mn ? BLOCK_K * 128 : 16;
      const uint32_t a_kstep = p.a_mn ? (UMMA_K * 128) : (UMMA_K * 2);
      const uint32_t b_kstep = p.b_mn ? (UMMA_K * 128) : (UMMA_K * 2);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int u = cluster_id; u < num_units; u += num_clusters) {
        const int tile = u / p.splits, split = u - tile * p.splits;
        const int kb0 = split * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * L::STAGE_BYTES, sb = sa + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc(sa + k * a_kstep, a_lbo, 1024);
            const uint64_t db = make_smem_desc(sb + k * b_kstep, b_lbo, 1024);
            umma_bf16_2sm(tmem_d, da, db, idesc, (kb > kb0 || k != 0) ? 1u : 0u);
          }
          umma_commit_2sm(empty_bar(stage));
          if (kb == kb1 - 1) umma_commit_2sm(tfull_bar(acc));
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================================== epilogue (both CTAs, 8 warps) ================================
    const int quarter = warp & 3;                  // TMEM lanes [32*quarter, +32)
    const int col_half = (warp - 4) >> 2;          // columns [HALF_N*col_half, +HALF_N)
    const uint32_t leader_tempty0 = mapa_u32(tempty_bar(0), 0);
    const uint32_t ebuf = smem_base + L::EPI_OFFSET + (warp - 4) * kEpiBufBytes;
    const uint32_t erow = ebuf + lane * 128;
    const uint32_t sw = (uint32_t)(lane & 7);
    const int chunk_cols = p.c_f32 ? 32 : 64;      // 128 bytes of C per row per TMA store box
    const int n_chunks = HALF_N / chunk_cols;
    int acc = 0; uint32_t acc_phase = 0;
    for (int u = cluster_id; u < num_units; u += num_clusters) {
      const int tile = u / p.splits, split = u - tile * p.splits;
      const int m0 = (tile / p.tiles_n) * (2 * BLOCK_M) + (int)cta_rank * BLOCK_M + quarter * 32;
      const int n0 = (tile % p.tiles_n) * BN + col_half * HALF_N;
      const bool add_bias = p.has_bias && split == 0;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + col_half * HALF_N);
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c) {
        const int col0 = n0 + c * chunk_cols;
        uint32_t w[32];                              // the 128 output bytes of this thread's row
        if (p.c_f32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          float bv[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) bv[i] = 0.f;
          if (add_bias && col0 < p.N) load_bias32(bias, p.bias_bf16, col0, p.N, bv);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) w[i] = __float_as_uint(fmaf(p.alpha, __uint_as_float(v[i]), bv[i]));
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t v[32];
            tmem_ld_32x32(taddr + c * 64 + h * 32, v);
            float bv[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) bv[i] = 0.f;
            if (add_bias && col0 + 32 * h < p.N) load_bias32(bias, p.bias_bf16, col0 + 32 * h, p.N, bv);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i)
              w[16 * h + i] = pack_bf16x2(fmaf(p.alpha, __uint_as_float(v[2 * i]), bv[2 * i]),
                                          fmaf(p.alpha, __uint_as_float(v[2 * i + 1]), bv[2 * i + 1]));
          }
        }
        if (c == n_chunks - 1) {                      // accumulator fully read: hand the TMEM buffer back early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(leader_tempty0 + 8u * acc);
        }
        if (lane == 0) bulk_wait_read0();             // the previous store of this warp has finished reading ebuf
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(erow + ((((uint32_t)j) ^ sw) << 4), w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && m0 < p.M && col0 < p.N) {
          if (p.c_reduce) tma_reduce_add_2d(&map_c, ebuf, col0, m0);
          else tma_store_2d(&map_c, ebuf, col0, m0);
          bulk_commit();
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) bulk_wait_read0();
    __syncwarp();
  }

  tc_fence_before();
  cluster_sync_all();          // both CTAs are done with TMEM and with each other's shared memory
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------------------
template <int BN, int kStages>
static int launch_gemm3(const tlt_gemm_desc* d, int splits, const void* A, const void* B, const void* bias, void* C, cudaStream_t st) {
  using L = SmemLayout3<BN, kStages>;
  CUtensorMap ma, mb, mc;
  int rc;
  if (!d->a_major) rc = make_map(&ma, A, d->K, d->M, d->lda, BLOCK_K, BLOCK_M);
  else rc = make_map(&ma, A, d->M, d->K, d->lda, 64, BLOCK_K);
  if (rc) return rc;
  if (!d->b_major) rc = make_map(&mb, B, d->K, d->N, d->ldb, BLOCK_K, BN / 2);
  else rc = make_map(&mb, B, d->N, d->K, d->ldb, 64, BLOCK_K);
  if (rc) return rc;
  const int c_f32 = d->dtype_c == TLT_F32;
  rc = make_map(&mc, C, d->N, d->M, d->ldc, c_f32 ? 32 : 64, 32, c_f32);     // 128-byte x 32-row store boxes
  if (rc) return rc;
  GemmParams p;
  p.M = (int)d->M; p.N = (int)d->N; p.K = (int)d->K; p.ldc = d->ldc;
  p.a_mn = d->a_major; p.b_mn = d->b_major;
  p.c_f32 = c_f32; p.accumulate = d->accumulate; p.has_bias = bias != nullptr; p.bias_bf16 = d->dtype_bias == TLT_BF16;
  p.alpha = d->alpha;
  p.tiles_m = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  p.tiles_n = (p.N + BN - 1) / BN;
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
  if (splits > num_kb) splits = num_kb;
  if (splits < 1) splits = 1;
  p.kb_per_split = (num_kb + splits - 1) / splits;
  p.splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;          // every split owns >= 1 k-block
  p.c_reduce = (p.accumulate || p.splits > 1) ? 1 : 0;
  if (p.splits > 1 && !p.accumulate)                                   // split-K partials are reduce-added into zeros
    TLT_CHECK_CUDA(cudaMemset2DAsync(C, (size_t)d->ldc * 4, 0, (size_t)d->N * 4, (size_t)d->M, st));
  TLT_ENSURE_SMEM((gemm_bf16_tc3_kernel<BN, kStages>), L::TOTAL);
  int units = p.tiles_m * p.tiles_n * p.splits;
  int pairs = num_sms() / 2;
  int clusters = units < pairs ? units : pairs;
  gemm_bf16_tc3_kernel<BN, kStages><<<2 * clusters, kGemm2Threads, L::TOTAL, st>>>(ma, mb, mc, p, bias);
  return check_launch("gemm_bf16_tc3_kernel");
}

// Tile width and split-K factor by a small cost model: units are dealt round-robin to the 74 SM pairs; a unit costs its
// k-blocks (a 128-wide tile moves 1.5x the operand bytes per FLOP of a 256-wide one, so it runs at ~0.6 of the time of a
// 256-wide tile rather than 0.5) plus a fixed pipeline-fill / epilogue term.
static void choose_tiling(const tlt_gemm_desc* d, int* bn_out, int* splits_out) {
  const int pairs = num_sms() / 2;
  const long long tm = (d->M + 255) / 256;
  const long long num_kb = (d->K + BLOCK_K - 1) / BLOCK_K;
  const bool can_split = d->dtype_c == TLT_F32;
  double best = 1e300;
  int best_bn = 256, best_s = 1;
  const int cand_s[] = {1, 2, 3, 4, 6, 8, 12, 16, 24, 37, 74, 148};
  for (int bn = 256; bn >= 128; bn -= 128) {
    if (bn == 256 && d->N <= 128) continue;
    const long long tiles = tm * ((d->N + bn - 1) / bn);
    for (int s : cand_s) {
      if (s > 1 && (!can_split || num_kb / s < 8)) continue;
      const long long units = tiles * s;
      const long long waves = (units + pairs - 1) / pairs;
      const double unit_cost = (double)((num_kb + s - 1) / s) * (bn == 256 ? 1.0 : 0.6) + 4.0;
      const double cost = (double)waves * unit_cost + (s > 1 ? 2.0 : 0.0);
      if (cost < best) { best = cost; best_bn = bn; best_s = s; }
    }
  }
  *bn_out = best_bn;
  *splits_out = best_s;
}

}  // namespace tlt

using namespace tlt;

extern "C" int tlt_gemm_bf16_tc_supported(const tlt_gemm_desc* d) {
  if (!d) return 0;
  if (d->M <= 0 || d->N <= 0 || d->K <= 0) return 0;
  if (d->M >= (1LL << 31) || d->N >= (1LL << 31) || d->K >= (1LL << 31)) return 0;
  if (d->lda % 8 || d->ldb % 8) return 0;                                  // TMA: 16-byte global strides
  if (d->lda < (d->a_major ? d->M : d->K) || d->ldb < (d->b_major ? d->N : d->K)) return 0;
  if (d->dtype_c != TLT_F32 && d->dtype_c != TLT_BF16) return 0;
  if (d->ldc % (d->dtype_c == TLT_F32 ? 4 : 8) || d->ldc < d->N) return 0;  // 16-byte vector stores in the epilogue
  if (d->accumulate && d->dtype_c != TLT_F32) return 0;
  return 1;
}

extern "C" int tlt_gemm_bf16_tc(const tlt_gemm_desc* d, const void* A, const void* B, const void* bias, void* C, void* stream) {
  TLT_REQUIRE(d && A && B && C, "tlt_gemm_bf16_tc: null argument");
  if (!tlt_gemm_bf16_tc_supported(d)) {
    set_error("tlt_gemm_bf16_tc: unsupported problem M=%lld N=%lld K=%lld lda=%lld ldb=%lld ldc=%lld", (long long)d->M,
              (long long)d->N, (long long)d->K, (long long)d->lda, (long long)d->ldb, (long long)d->ldc);
    return TLT_ERR_UNSUPPORTED;
  }
  TLT_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)B % 16) == 0 && ((uintptr_t)C % 16) == 0,
              "tlt_gemm_bf16_tc: operands must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  int bn, splits;
  choose_tiling(d, &bn, &splits);
  if (d->split_k > 0) splits = d->dtype_c == TLT_F32 ? d->split_k : 1;
  if (bn == 256) return launch_gemm3<256, 6>(d, splits, A, B, bias, C, st);
  return launch_gemm3<128, 8>(d, splits, A, B, bias, C, st);
}
