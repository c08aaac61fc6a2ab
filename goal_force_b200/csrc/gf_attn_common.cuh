// gf_attn_common.cuh -- softmax building blocks shared by the attention kernels (gf_attn.cu, gf_attn80.cu).
#pragma once
#include "gf_ptx.cuh"

namespace gf {

// ------------------------------------------------------------------ packed f32x2 helpers (FFMA2 / FADD2 on sm_100)
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t pack2u(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// 2^x for two lanes on the FMA pipe: x = n + r, n = rint(x), r in [-0.5, 0.5];  2^r by a degree-3 minimax
// polynomial (max relative error 7.5e-5), 2^n by adding n to the exponent field.  x is clamped to >= -125.
__device__ __forceinline__ void exp2_poly2(uint64_t x2, float& p0, float& p1) {
  constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23: low mantissa bits of (x + kMagic) hold rint(x)
  float x0, x1;
  unpack2(x2, x0, x1);
  x0 = fmaxf(x0, -125.0f);
  x1 = fmaxf(x1, -125.0f);
  x2 = pack2(x0, x1);
  const uint64_t t2 = fadd2(x2, pack2(kMagic, kMagic));
  const uint64_t n2 = fadd2(t2, pack2(-kMagic, -kMagic));
  const uint64_t r2 = ffma2(n2, pack2(-1.0f, -1.0f), x2);
  uint64_t q2 = ffma2(pack2(0.0551716685295105f, 0.0551716685295105f), r2, pack2(0.2426111251115799f, 0.2426111251115799f));
  q2 = ffma2(q2, r2, pack2(0.6932609677314758f, 0.6932609677314758f));
  q2 = ffma2(q2, r2, pack2(0.9999280571937561f, 0.9999280571937561f));
  float q0, q1, t0, t1;
  unpack2(q2, q0, q1);
  unpack2(t2, t0, t1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

// Exponentiate kCols columns of a score row: p = 2^(s*scale - m), accumulate the row sum, pack to bf16.
// kEmuPairs of every 16 column pairs go through the polynomial, spread evenly between the MUFU pairs.
template <int kEmuPairs, int kCols>
__device__ __forceinline__ void exp_cols(const uint32_t (&s)[kCols], uint64_t scale2, uint64_t negm2,
                                         uint64_t (&acc)[2], uint32_t (&pk)[kCols / 2]) {
#pragma unroll
  for (int k = 0; k < kCols / 2; ++k) {
    const uint64_t x2 = ffma2(pack2u(s[2 * k], s[2 * k + 1]), scale2, negm2);
    float p0, p1;
    const bool emulate = ((k + 1) * kEmuPairs) / 16 != (k * kEmuPairs) / 16;
    if (emulate) {
      exp2_poly2(x2, p0, p1);
    } else {
      float x0, x1;
      unpack2(x2, x0, x1);
      p0 = ex2_approx(x0);
      p1 = ex2_approx(x1);
    }
    acc[k & 1] = fadd2(acc[k & 1], pack2(p0, p1));
    pk[k] = pack_bf16x2(p0, p1);
  }
}
template <int kEmuPairs>
__device__ __forceinline__ void exp_chunk(const uint32_t (&s)[32], uint64_t scale2, uint64_t negm2, uint64_t (&acc)[2],
                                          uint32_t (&pk)[16]) {
  exp_cols<kEmuPairs, 32>(s, scale2, negm2, acc, pk);
}

template <int kCols>
__device__ __forceinline__ float cols_max(const uint32_t (&s)[kCols]) {
  float m = fmaxf(__uint_as_float(s[0]), __uint_as_float(s[1]));
#pragma unroll
  for (int k = 1; k < kCols / 2; ++k) m = fmax3(m, __uint_as_float(s[2 * k]), __uint_as_float(s[2 * k + 1]));
  return m;
}
// same maximum through four independent chains (depth kCols/8 + 2 instead of kCols/2)
template <int kCols>
__device__ __forceinline__ float cols_max4(const uint32_t (&s)[kCols]) {
  static_assert(kCols % 8 == 0, "cols_max4 needs a multiple of 8 columns");
  float m[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) m[c] = fmaxf(__uint_as_float(s[2 * c]), __uint_as_float(s[2 * c + 1]));
#pragma unroll
  for (int k = 4; k < kCols / 2; ++k) m[k & 3] = fmax3(m[k & 3], __uint_as_float(s[2 * k]), __uint_as_float(s[2 * k + 1]));
  return fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3]));
}
__device__ __forceinline__ float chunk_max(const uint32_t (&s)[32]) { return cols_max<32>(s); }

}  // namespace gf
