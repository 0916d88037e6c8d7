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
#include "common.cuh"
#include <cuda.h>
#include <mutex>

namespace tlt {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;          // 64 bf16 = 128 bytes = one SWIZZLE_128B span
constexpr int UMMA_K = 16;
constexpr int kGemmThreads = 256;
constexpr uint32_t kSpinLimit = 20u * 1000u * 1000u;   // bounded mbarrier spin: trap instead of hanging the box

// ------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets lane (base_lane + t), v[i] = column (base_col + i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (sm_100 "version 1"), SWIZZLE_128B.  Fields in 16-byte units:
//   [0,14) start address, [16,30) leading byte offset, [32,46) stride byte offset, [46,48) version = 1,
//   [61,64) layout type (2 = SWIZZLE_128B).
//   K-major  : rows of 128 B, 8-row swizzle atoms 1024 B apart  -> SBO = 1024, LBO unused (1).
//   MN-major : 64-element (128 B) MN spans, 8 K-rows per atom    -> SBO = 1024 (next 8 K rows), LBO = bytes between
//              consecutive 64-wide MN chunks (= BLOCK_K * 128 here: one TMA box per chunk).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32 (cute::UMMA::InstrDescriptor bit layout)
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4)                        // c_format  = F32
         | (1u << 7)                      // a_format  = BF16
         | (1u << 10)                     // b_format  = BF16
         | ((uint32_t)a_mn_major << 15)   // a_major
         | ((uint32_t)b_mn_major << 16)   // b_major
         | ((uint32_t)(n >> 3) << 17)     // n_dim
         | ((uint32_t)(m >> 4) << 24);    // m_dim
}

struct GemmParams {
  int M, N, K;
  long long ldc;
  int a_mn, b_mn;
  int c_f32, accumulate, has_bias, bias_bf16;
  float alpha;
  int tiles_m, tiles_n;
};

template <int BN, int kStages>
struct SmemLayout {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;      // 16 KB
  static constexpr int B_BYTES = BN * BLOCK_K * 2;           // 32 KB @ BN=256
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = kStages * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;      // barriers + tmem slot + alignment slack
};

template <int BN, int kStages>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const GemmParams p, const void* __restrict__ bias, void* __restrict__ Cout) {
  using L = SmemLayout<BN, kStages>;
  constexpr uint32_t TMEM_COLS = 2 * BN;                     // two accumulator buffers (power of two: 256 / 512)
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
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128); }
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
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.tiles_n) * BLOCK_M, n0 = (tile % p.tiles_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * L::STAGE_BYTES, sb = sa + L::A_BYTES;
          mbar_expect_tx(full_bar(stage), L::STAGE_BYTES);
          const int k0 = kb * BLOCK_K;
          if (!p.a_mn) {
            tma_load_2d(sa, &map_a, full_bar(stage), k0, m0);                       // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int c = 0; c < BLOCK_M / 64; ++c)                                   // box {64 m, 64 k} per chunk
              tma_load_2d(sa + c * (BLOCK_K * 128), &map_a, full_bar(stage), m0 + 64 * c, k0);
          }
          if (!p.b_mn) {
            tma_load_2d(sb, &map_b, full_bar(stage), k0, n0);                       // box {64 k, BN n}
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_2d(sb + c * (BLOCK_K * 128), &map_b, full_bar(stage), n0 + 64 * c, k0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BLOCK_M, BN, p.a_mn, p.b_mn);
      const uint32_t a_lbo = p.a_mn ? BLOCK_K * 128 : 16, b_lbo = p.b_mn ? BLOCK_K * 128 : 16;
      const uint32_t a_kstep = p.a_mn ? (UMMA_K * 128) : (UMMA_K * 2);               // bytes per UMMA_K step
      const uint32_t b_kstep = p.b_mn ? (UMMA_K * 128) : (UMMA_K * 2);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * L::STAGE_BYTES, sb = sa + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc(sa + k * a_kstep, a_lbo, 1024);
            const uint64_t db = make_smem_desc(sb + k * b_kstep, b_lbo, 1024);
            umma_bf16(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));            // frees this smem stage once the MMAs above have read it
          if (kb == num_kb - 1) umma_commit(tfull_bar(acc));
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================================== epilogue =========================================
    const int ew = warp - 4;                       // == warp % 4 -> TMEM lanes [32*ew, 32*ew + 32)
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / p.tiles_n) * BLOCK_M, n0 = (tile % p.tiles_n) * BN;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int row = m0 + ew * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;
        if (row < p.M && col0 < p.N) {
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = p.alpha * __uint_as_float(v[i]);
          if (p.has_bias) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col0 + i < p.N)
                f[i] += p.bias_bf16 ? __bfloat162float(((const __nv_bfloat16*)bias)[col0 + i]) : __ldg((const float*)bias + col0 + i);
          }
          if (p.c_f32) {
            float* dst = (float*)Cout + (long long)row * p.ldc + col0;
            if (col0 + 32 <= p.N) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                float4 o = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
                if (p.accumulate) { float4 q = *(float4*)(dst + i); o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w; }
                *(float4*)(dst + i) = o;
              }
            } else {
              for (int i = 0; i < 32 && col0 + i < p.N; ++i) dst[i] = p.accumulate ? dst[i] + f[i] : f[i];
            }
          } else {
            __nv_bfloat16* dst = (__nv_bfloat16*)Cout + (long long)row * p.ldc + col0;
            if (col0 + 32 <= p.N) {
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                uint4 o;
                __nv_bfloat162 h0 = __floats2bfloat162_rn(f[i], f[i + 1]), h1 = __floats2bfloat162_rn(f[i + 2], f[i + 3]);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(f[i + 4], f[i + 5]), h3 = __floats2bfloat162_rn(f[i + 6], f[i + 7]);
                o.x = *(uint32_t*)&h0; o.y = *(uint32_t*)&h1; o.z = *(uint32_t*)&h2; o.w = *(uint32_t*)&h3;
                *(uint4*)(dst + i) = o;
              }
            } else {
              for (int i = 0; i < 32 && col0 + i < p.N; ++i) dst[i] = __float2bfloat16_rn(f[i]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  });
  return fn;
}

// 2-D bf16 tensor map: inner (contiguous) extent d0, outer extent d1 with row pitch ld elements.
static int make_map(CUtensorMap* map, const void* ptr, int64_t d0, int64_t d1, int64_t ld, int box0, int box1) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)"); return TLT_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)d0, (cuuint64_t)d1};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for dims {%lld,%lld} ld %lld box {%d,%d}", (int)r, (long long)d0,
              (long long)d1, (long long)ld, box0, box1);
    return TLT_ERR_CUDA;
  }
  return TLT_OK;
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN, int kStages>
static int launch_gemm(const tlt_gemm_desc* d, const void* A, const void* B, const void* bias, void* C, cudaStream_t st) {
  using L = SmemLayout<BN, kStages>;
  CUtensorMap ma, mb;
  int rc;
  if (!d->a_major) rc = make_map(&ma, A, d->K, d->M, d->lda, BLOCK_K, BLOCK_M);
  else rc = make_map(&ma, A, d->M, d->K, d->lda, 64, BLOCK_K);
  if (rc) return rc;
  if (!d->b_major) rc = make_map(&mb, B, d->K, d->N, d->ldb, BLOCK_K, BN);
  else rc = make_map(&mb, B, d->N, d->K, d->ldb, 64, BLOCK_K);
  if (rc) return rc;

  GemmParams p;
  p.M = (int)d->M; p.N = (int)d->N; p.K = (int)d->K; p.ldc = d->ldc;
  p.a_mn = d->a_major; p.b_mn = d->b_major;
  p.c_f32 = d->dtype_c == TLT_F32; p.accumulate = d->accumulate; p.has_bias = bias != nullptr; p.bias_bf16 = d->dtype_bias == TLT_BF16;
  p.alpha = d->alpha;
  p.tiles_m = (p.M + BLOCK_M - 1) / BLOCK_M;
  p.tiles_n = (p.N + BN - 1) / BN;
  static bool attr_set = false;
  if (!attr_set) {
    TLT_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tc_kernel<BN, kStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_set = true;
  }
  int tiles = p.tiles_m * p.tiles_n;
  int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_bf16_tc_kernel<BN, kStages><<<grid, kGemmThreads, L::TOTAL, st>>>(ma, mb, p, bias, C);
  return check_launch("gemm_bf16_tc_kernel");
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
  // Tile choice: 128x256 tiles halve the smem traffic per FLOP, 128x128 tiles fill the 148 SMs better on small outputs.
  long long t256 = ((d->M + 127) / 128) * ((d->N + 255) / 256);
  long long t128 = ((d->M + 127) / 128) * ((d->N + 127) / 128);
  int sms = num_sms();
  auto waves_eff = [&](long long tiles) { long long w = (tiles + sms - 1) / sms; return (double)tiles / (double)(w * sms); };
  bool use256 = d->N > 128 && (waves_eff(t256) >= 0.8 * waves_eff(t128));
  if (use256) return launch_gemm<256, 4>(d, A, B, bias, C, st);
  return launch_gemm<128, 6>(d, A, B, bias, C, st);
}
