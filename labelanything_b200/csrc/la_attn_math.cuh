// labelanything_b200 — arithmetic helpers shared by the attention kernels (la_attention.cu, la_attention_win.cu):
// approximate exponentials, packed fp32x2 forms, named barriers.
#pragma once

#include <cuda_fp16.h>

#include "la_common.cuh"

namespace la {

__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// named barriers 1..15 (0 is __syncthreads): `count` threads in total, bar_sync-ers and bar_arrive-rs together
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// sm_100 packed-pair / three-input forms: half the issue slots of the scalar instructions (FMNMX3, FFMA2, FADD2)
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
// (d0, d1) = (a0, a1) * (b, b) + (c, c)
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b, float c) {
  asm("{\n\t"
      ".reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %4};\n\t"
      "mov.b64 rc, {%5, %5};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t"
      "}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
// 2^a for two arguments on the FMA pipe instead of the MUFU: a = n + f, n = round(a), |f| <= 1/2 (magic-number rounding),
// 2^f by a degree-3 minimax polynomial (relative error 7.5e-5 -- far below the bf16 rounding of P), 2^n by adding
// n to the exponent field.  MUFU.EX2 issues at a quarter of the FMA rate, and the exponentials are what bounds the
// softmax warps at head_dim 64, so a fraction of every score row takes this path (the FlashAttention-4 trick).
// Arguments are clamped at -126 (result ~1e-38 instead of 0 for masked keys); the caller guarantees a <= 8.
__device__ __forceinline__ void exp2_poly_x2(float& a0, float& a1) {
  const float x0 = fmaxf(a0, -126.0f), x1 = fmaxf(a1, -126.0f);
  uint32_t t0, t1, p0, p1;
  asm("{\n\t"
      ".reg .b64 x, t, r, f, p, c;\n\t"
      "mov.b64 x, {%4, %5};\n\t"
      "mov.b64 c, {%6, %6};\n\t"
      "add.rn.f32x2 t, x, c;\n\t"          // t = x + 1.5 * 2^23: round(x) in the low mantissa bits
      "mov.b64 c, {%7, %7};\n\t"
      "add.rn.f32x2 r, t, c;\n\t"          // r = round(x)
      "mov.b64 c, {%8, %8};\n\t"
      "fma.rn.f32x2 f, r, c, x;\n\t"       // f = x - r
      "mov.b64 p, {%9, %9};\n\t"
      "mov.b64 c, {%10, %10};\n\t"
      "fma.rn.f32x2 p, p, f, c;\n\t"
      "mov.b64 c, {%11, %11};\n\t"
      "fma.rn.f32x2 p, p, f, c;\n\t"
      "mov.b64 c, {%12, %12};\n\t"
      "fma.rn.f32x2 p, p, f, c;\n\t"
      "mov.b64 {%0, %1}, t;\n\t"
      "mov.b64 {%2, %3}, p;\n\t"
      "}"
      : "=r"(t0), "=r"(t1), "=r"(p0), "=r"(p1)
      : "f"(x0), "f"(x1), "f"(12582912.0f), "f"(-12582912.0f), "f"(-1.0f), "f"(0.0551716685f), "f"(0.2426111251f),
        "f"(0.6932609677f), "f"(0.9999280572f));
  a0 = __uint_as_float(p0 + (t0 << 23));
  a1 = __uint_as_float(p1 + (t1 << 23));
}
// (d0, d1) = (a0, a1) * (b, b) + (c0, c1)
__device__ __forceinline__ void ffma2v(float& d0, float& d1, float a0, float a1, float b, float c0, float c1) {
  asm("{\n\t"
      ".reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %4};\n\t"
      "mov.b64 rc, {%5, %6};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t"
      "}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b), "f"(c0), "f"(c1));
}
// (d0, d1) = (a0, a1) + (b, b)
__device__ __forceinline__ void fadd2s(float& d0, float& d1, float a0, float a1, float b) {
  asm("{\n\t"
      ".reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %4};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t"
      "}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b));
}
// (d0, d1) = (a0, a1) * (b, b)
__device__ __forceinline__ void fmul2s(float& d0, float& d1, float a0, float a1, float b) {
  asm("{\n\t"
      ".reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %4};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t"
      "}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b));
}
// (d0, d1) += (a0, a1)
__device__ __forceinline__ void fadd2_acc(float& d0, float& d1, float a0, float a1) {
  asm("{\n\t"
      ".reg .b64 ra, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rd, {%0, %1};\n\t"
      "add.rn.f32x2 rd, rd, ra;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t"
      "}"
      : "+f"(d0), "+f"(d1)
      : "f"(a0), "f"(a1));
}


}  // namespace la
