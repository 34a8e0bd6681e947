// packed_f32.cuh — sm_100a packed FP32: one instruction (FFMA2 / FMUL2 / FADD2; PTX fma/mul/add.rn.f32x2)
// does two individually rounded IEEE binary32 operations per lane on a 64-bit register pair, in ONE issue
// slot and at the scalar datapath rate of 128 lanes/clk/SM (profiles/experiments/ffma2_bench.cu).  Operands
// may be a pair, one register broadcast to both halves (`R.F32`), a swapped pair (`.LO_HI`) or an immediate,
// so `F2(s, s)` and `F2(c)` cost nothing.  Kernels that are bound by instruction issue, not by the FP32
// datapath, get cheaper by pairing independent operations.
#pragma once
#include <cuda_runtime.h>

namespace rmdf {

typedef float2 f2;
__device__ __forceinline__ f2 F2(float a) { return make_float2(a, a); }
__device__ __forceinline__ f2 F2(float a, float b) { return make_float2(a, b); }
// Negation as PTX `neg.f32` (never neg.ftz): ptxas folds it into the negate modifier of the packed instruction's
// operand; the C++ `-a` of a --use_fast_math unit becomes neg.ftz, which it materialises as two scalar FADDs.
__device__ __forceinline__ float neg1(float a) { float r; asm("neg.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ f2 operator-(f2 a) { return make_float2(neg1(a.x), neg1(a.y)); }
__device__ __forceinline__ f2 operator*(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 operator+(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 operator-(f2 a, f2 b) { return __fadd2_rn(a, -b); }
__device__ __forceinline__ f2 operator*(float a, f2 b) { return __fmul2_rn(F2(a), b); }    // scalar, broadcast
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 fma2(float a, f2 b, f2 c) { return __ffma2_rn(F2(a), b, c); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, float c) { return __ffma2_rn(a, b, F2(c)); }

// ---- IEEE-exact use (the -fmad=false units) -----------------------------------------------------------
// HAZARD: ptxas 12.9 contracts `mul.rn.f32x2` + `add.rn.f32x2` into FFMA2 even under --fmad=false (the scalar
// .rn forms are never contracted) when the product has a single use and feeds the add directly; that changes
// roundings.  julia.cu is safe by construction — every product there has two uses or feeds another product —
// and tests/test_host_cpu.py disassembles the built -fmad=false objects and fails if any FFMA2 appears in
// them.  Where a single-use product must feed a sum, write a + b as fma(−a, −1, b): (−a)·(−1) is exact, the
// result is RN(a + b) bit for bit, and the product sits in the ADDEND, where there is nothing to contract
// (with an immediate addend ptxas folds the −1 away again: use two scalar __fadd_rn there).  This was used
// for a paired evaluation of the literal power-8 estimator in the parity build (bit-identical, but no faster:
// profiles/r1_packed_fp32.md).

}  // namespace rmdf
