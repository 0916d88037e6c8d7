// Device helpers shared by the fused CP-conv rows kernels (cpconv_rows.cu: TMEM-resident z1 with halo, stride 1 / 2;
// cpconv_rows2.cu: the first separable pass folded into the stage-1 MMA).  sm_100a only.
#pragma once
#include "tc_common.cuh"
#include "f32x2.cuh"

namespace tlt {
namespace cpr {

constexpr int NP = 112;                // positions per chunk = per stage-3 tile
constexpr int kChunkBytes = NP * 128;  // one chunk (two 56-position rows), one 64-channel block
constexpr int kDwWarps = 8;
constexpr int kEpiThreads = 128;

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v[0]) : "r"(taddr) : "memory");
}
// tcgen05.wait::ld that also "touches" the loaded registers, so no consumer of them can be scheduled above the wait
__device__ __forceinline__ void tmem_wait8(uint32_t* v) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]) :: "memory");
}
__device__ __forceinline__ void tmem_wait16(uint32_t* v) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                 "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]) :: "memory");
}
__device__ __forceinline__ void tmem_wait1(uint32_t* v) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v[0]) :: "memory");
}
__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
#define F(u) __uint_as_float(u)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
// shared-memory matrix descriptors as (lo, hi) words: hi is the same for every operand here (SBO = 1024, version 1, SWIZZLE_128B),
// lo = start address | leading byte offset, both in 16-byte units -- an operand's k-step / stage offset is one 32-bit add
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr & 0x3FFFF) >> 4) | ((lbo_bytes >> 4) << 16); }
__device__ __forceinline__ void umma_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accum), "r"(kDescHi)
      : "memory");
}

// z2 tile address of (position m, rank row r): [m / 64][r][128 B], 16-byte chunk index XOR (r & 7)  (SWIZZLE_128B)
__device__ __forceinline__ uint32_t z_addr(uint32_t zbuf, int m, int r) {
  return zbuf + (uint32_t)((m >> 6) * 16384 + r * 128) + ((uint32_t)((((m & 63) >> 3) ^ (r & 7))) << 4);
}

// 8-byte slot of positions m .. m+3 (m a multiple of 4) of rank row r
__device__ __forceinline__ uint32_t z_addr4(uint32_t zbuf, int m, int r) { return z_addr(zbuf, m, r) | (uint32_t)((m & 4) << 1); }
// eight consecutive outputs starting at position m (a multiple of 4): one 16-byte store when aligned, two 8-byte stores otherwise
__device__ __forceinline__ void z_store8(uint32_t zbuf, int m, int r, const float (&o)[8]) {
  const uint32_t w0 = pack_bf16x2(o[0], o[1]), w1 = pack_bf16x2(o[2], o[3]), w2 = pack_bf16x2(o[4], o[5]), w3 = pack_bf16x2(o[6], o[7]);
  if ((m & 7) == 0) {
    st_shared_v4(z_addr(zbuf, m, r), w0, w1, w2, w3);
  } else {
    st_shared_v2(z_addr4(zbuf, m, r), w0, w1);
    st_shared_v2(z_addr4(zbuf, m + 4, r), w2, w3);
  }
}

// ---- packed fp32 pairs -------------------------------------------------------------------------------------------------------
// A three-register FFMA issues every other cycle per SM sub-partition on this part; fma.rn.f32x2 does two FMAs in the same slot.
// Adjacent TMEM columns land in adjacent registers (tcgen05.ld .x8), so a pair of columns IS an aligned 64-bit register pair.
using namespace tlt::f32x2;


}  // namespace cpr
}  // namespace tlt
