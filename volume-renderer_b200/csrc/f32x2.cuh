// f32x2.cuh -- packed Blackwell f32x2 arithmetic (FADD2 / FMUL2 / FFMA2) as value types.
//
// Each operation is one correctly rounded binary32 operation per lane, so packing changes
// the number of issue slots, never a result bit.  ptxas (12.9) contracts a mul.rn.f32x2 that
// feeds an add/sub.rn.f32x2 into FFMA2 even under -fmad=false (it honours the explicit .rn
// only on scalar ops), so wherever the shader has an UNFUSED product feeding a sum the sum
// is issued as scalar add.rn.f32 (callers use __fadd_rn on lo()/hi()).
#pragma once

#include <cuda_runtime.h>

namespace vr {

typedef unsigned long long u64;

struct f2 { u64 v; };
__device__ __forceinline__ f2 mk2(float a, float b) { f2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f2 splat2(float a) { return mk2(a, a); }
__device__ __forceinline__ float lo(f2 p) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v)); return a; }
__device__ __forceinline__ float hi(f2 p) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v)); return b; }
__device__ __forceinline__ f2 fadd(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 fsub(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 fmul(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 ffma(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }

}  // namespace vr
