// Packed fp32 pairs (device only): Blackwell issues two fp32 operations per instruction on a 64-bit register pair (PTX add / sub /
// mul / fma .f32x2 -> SASS FADD2 / FMUL2 / FFMA2; operand negation, an immediate and a scalar broadcast are folded into the instruction
// by ptxas).  For issue-bound code whose arithmetic comes in structurally identical pairs: half the issue slots, each half the IEEE
// operation the scalar instruction does.
#pragma once

namespace ape {

struct F2 { unsigned long long v; };

__device__ __forceinline__ F2 pk(float lo, float hi) { F2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ F2 splat(float x) { return pk(x, x); }
__device__ __forceinline__ float lo(F2 a) { return __uint_as_float((unsigned)a.v); }            // the pair's registers: no instruction
__device__ __forceinline__ float hi(F2 a) { return __uint_as_float((unsigned)(a.v >> 32)); }
__device__ __forceinline__ F2 operator+(F2 a, F2 b) { F2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 operator-(F2 a, F2 b) { F2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 operator*(F2 a, F2 b) { F2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 operator-(F2 a) { return pk(-lo(a), -hi(a)); }          // folded into the consumer's operand modifier
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) {
    F2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r;
}
}  // namespace ape
