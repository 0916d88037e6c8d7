// Packed fp32 pairs (fma.rn.f32x2 / mul.rn.f32x2, sm_100a): two FMAs per issue slot -- a three-register FFMA issues every other
// cycle per SM sub-partition on this part, the packed form does two in the same slot.  A pair is an aligned 64-bit register pair;
// a bf16x2 word unpacks into exactly such a pair (low half << 16, high half masked), and adjacent TMEM columns land in adjacent
// registers (tcgen05.ld .x8).  Shared by the depthwise stages of the fused CP-conv kernels and by the separable sweeps.
#pragma once
#include <stdint.h>

namespace tlt {
namespace f32x2 {

typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2 pku(uint32_t lo, uint32_t hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ float lo_of(f2 v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi_of(f2 v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

}  // namespace f32x2
}  // namespace tlt
