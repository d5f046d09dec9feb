// aep_pack.cuh -- packed fp32x2 arithmetic of sm_100 (FFMA2 / FMUL2 / FADD2).
//
// A B200 SMSP issues one warp instruction per clock and its FP32 pipe retires 32 lanes per clock, so scalar FFMA code is
// bound by the issue port as soon as anything else (LDS, address arithmetic, moves) shares it: the particle kernels sat at
// 65-76 % issue-slot utilisation with the FMA pipe at 43-50 % (profiles/r1_v8_*).  fma.rn.f32x2 performs two IEEE fp32 FMAs
// on an aligned register pair in ONE issue slot (measured: same lane throughput as FFMA, half the instructions --
// profiles/r1_ubench_fma_pipes.txt), which frees issue slots for the loads and bookkeeping.  Each lane is rounded exactly
// like the scalar instruction, so results do not depend on which form the compiler or the author picked.
#pragma once
#ifndef AEP_HOST_MATH_TEST   // tests/cpu_math_harness.cpp compiles this header for the host: same lane-wise IEEE arithmetic, no PTX
#include <cuda_runtime.h>
#endif

namespace aep {

typedef unsigned long long f32x2;      // (lo, hi) = two fp32 values in an aligned 64-bit register pair

#ifndef AEP_HOST_MATH_TEST
__device__ __forceinline__ f32x2 pk(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
#else
inline f32x2 pk(float lo, float hi) { unsigned a, b; std::memcpy(&a, &lo, 4); std::memcpy(&b, &hi, 4); return (f32x2)a | ((f32x2)b << 32); }
inline void upk(f32x2 v, float& lo, float& hi) { const unsigned a = (unsigned)v, b = (unsigned)(v >> 32); std::memcpy(&lo, &a, 4); std::memcpy(&hi, &b, 4); }
inline f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { float al, ah, bl, bh, cl, ch; upk(a, al, ah); upk(b, bl, bh); upk(c, cl, ch); return pk(std::fmaf(al, bl, cl), std::fmaf(ah, bh, ch)); }
inline f32x2 mul2(f32x2 a, f32x2 b) { float al, ah, bl, bh; upk(a, al, ah); upk(b, bl, bh); return pk(al * bl, ah * bh); }
inline f32x2 add2(f32x2 a, f32x2 b) { float al, ah, bl, bh; upk(a, al, ah); upk(b, bl, bh); return pk(al + bl, ah + bh); }
#endif
__device__ __forceinline__ f32x2 pk1(float v) { return pk(v, v); }
__device__ __forceinline__ float lo_of(f32x2 v) { float a, b; upk(v, a, b); return a; }
__device__ __forceinline__ float hi_of(f32x2 v) { float a, b; upk(v, a, b); return b; }

// a float4 in shared / global memory seen as two packed pairs: .x = (x, y), .y = (z, w)   (one LDS.128 / LDG.128)
__device__ __forceinline__ ulonglong2 ld_pairs(const float4* p) { return *reinterpret_cast<const ulonglong2*>(p); }
__device__ __forceinline__ float4 quad(f32x2 xy, f32x2 zw) { float4 r; upk(xy, r.x, r.y); upk(zw, r.z, r.w); return r; }

}  // namespace aep
