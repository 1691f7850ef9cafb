// gc_math.cuh — Gaussian likelihood arithmetic shared by the Gaussian-conditional kernels
// (plain and STanH): packed f32x2 helpers, exact shared-reciprocal division, the custom erfc.
#pragma once
#include "common.cuh"

namespace reslic {

// ------------------------------------------------------------------ per-element math
// Mirrors the reference op order: every line is one IEEE fp32 op there.
//   a = (0.5 - v)/s, b = (-0.5 - v)/s   (true divides)
//   L = 0.5*erfc(-2^-0.5 * a) - 0.5*erfc(-2^-0.5 * b)
//
// All FP32 add/mul/fma work is issued as Blackwell packed f32x2 instructions (FFMA2 / FMUL2 /
// FADD2: two IEEE-rounded fp32 lanes per issue slot, sm_100+), two elements per lane pair.
// Results are bit-identical to the scalar ops; what halves is the number of issue slots, which
// is what bounds this kernel next to HBM.
typedef float2 F2;
__device__ __forceinline__ F2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ F2 neg2(F2 a) { return make_float2(-a.x, -a.y); }   // folds into an operand modifier
__device__ __forceinline__ F2 abs2(F2 a) { return make_float2(fabsf(a.x), fabsf(a.y)); }
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ F2 mul2(F2 a, F2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ F2 add2(F2 a, F2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ F2 rcp2(F2 a) { return make_float2(rcp_approx(a.x), rcp_approx(a.y)); }
__device__ __forceinline__ F2 min_nan2(F2 a, float b) { return make_float2(min_nan(a.x, b), min_nan(a.y, b)); }
__device__ __forceinline__ F2 max_nan2(F2 a, float b) { return make_float2(max_nan(a.x, b), max_nan(a.y, b)); }

// The two divides of an element share one reciprocal and use the Markstein residual
// correction, which returns the correctly rounded quotient for the operand range left after
// the clamps (cross-checked against the MIRROR build and the oracle in tests/test_gc_parity.py).
__device__ __forceinline__ void div2_rn(F2 n1, F2 n2, F2 s, F2& q1, F2& q2) {
  F2 r = rcp2(s);
  const F2 ns = neg2(s);
  const F2 e = fma2(ns, r, f2(1.0f));
  r = fma2(r, e, r);
  F2 q = mul2(n1, r);
  F2 rem = fma2(ns, q, n1);
  q1 = fma2(rem, r, q);
  q = mul2(n2, r);
  rem = fma2(ns, q, n2);
  q2 = fma2(rem, r, q);
}

// erfc(x) for 0 <= x <= 12 as exp(-x^2) * (1-u) * Q(u), u = x/(x+2.5): per element one
// MUFU.RCP, one MUFU.EX2 and 19 FP32 ops (CUDA's erfcf: 3 MUFU + FRND + ~44).  Q is a
// degree-9 near-minimax fit (|rel err| < 1.4e-8 in exact arithmetic); u = x*r keeps small x
// free of cancellation and (1-u) carries the 1/x decay so Horner stays well conditioned.
// exp(-x^2) gets the rounding errors of x*x and of the log2(e) product back as a first-order
// correction, so the relative error stays ~3e-7 out to the likelihood floor (x^2 ~ 20).
__device__ __forceinline__ F2 erfc_pos_fast(F2 x) {
  const F2 r = rcp2(add2(x, f2(2.5f)));
  const F2 u = mul2(x, r);
  const F2 w = fma2(neg2(x), r, f2(1.0f));
  F2 q = f2(2.651532926e-02f);
  q = fma2(q, u, f2(-5.741734803e-02f));
  q = fma2(q, u, f2(-3.597635776e-02f));
  q = fma2(q, u, f2(1.268966794e-01f));
  q = fma2(q, u, f2(1.091585010e-01f));
  q = fma2(q, u, f2(-2.630832791e-01f));
  q = fma2(q, u, f2(-4.676126838e-01f));
  q = fma2(q, u, f2(1.608165503e+00f));
  q = fma2(q, u, f2(-1.820949554e+00f));
  q = fma2(q, u, f2(1.0f));
  const F2 L2E = f2(1.44269502162933349609375f);      // fp32(log2 e)
  const F2 s2 = mul2(x, x);
  const F2 e = fma2(x, x, neg2(s2));                  // exact low part of x*x
  const F2 t = mul2(s2, L2E);
  F2 tl = fma2(s2, L2E, neg2(t));                     // exact low part of s2*L2E
  tl = fma2(e, L2E, tl);                              // (s2*L2E_LO <= 4e-7 at the floor: dropped)
  const F2 E0 = make_float2(ex2_approx(-t.x), ex2_approx(-t.y));
  const F2 E = fma2(mul2(E0, tl), f2(-0.693147182464599609375f), E0);   // 2^-(t+tl) ~ E0*(1 - ln2*tl)
  return mul2(E, mul2(w, q));
}

// L for operands known to be finite and moderate: 0 <= v <= 2^23, scale_bound <= s < 2^22 with
// scale_bound >= 1e-3 (the vector kernel's per-group range check).  |x| then stays below 1e10, x*x
// is finite, ex2 flushes to zero beyond x ~ 9.4 and 0 * finite = 0: no clamp is needed and the
// value equals gc_likelihood's bit for bit.
template <bool FAST>
__device__ __forceinline__ F2 gc_likelihood_finite(F2 v, F2 s) {
  F2 a, b;
  div2_rn(add2(f2(0.5f), neg2(v)), add2(f2(-0.5f), neg2(v)), s, a, b);
  const F2 c = f2(-0.70710678118654752440f);
  const F2 xa = mul2(c, a), xb = mul2(c, b);  // xb > 0 always; xa < 0 iff v < 0.5
  if (FAST) {
    F2 ea = erfc_pos_fast(abs2(xa));
    const F2 eb = erfc_pos_fast(xb);
    ea.x = (xa.x < 0.0f) ? 2.0f - ea.x : ea.x;
    ea.y = (xa.y < 0.0f) ? 2.0f - ea.y : ea.y;
    return fma2(f2(0.5f), ea, mul2(f2(-0.5f), eb));
  }
  const F2 upper = make_float2(0.5f * erfcf(xa.x), 0.5f * erfcf(xa.y));
  const F2 lower = make_float2(0.5f * erfcf(xb.x), 0.5f * erfcf(xb.y));
  return add2(upper, neg2(lower));
}

template <bool FAST>
__device__ __forceinline__ F2 gc_likelihood(F2 v, F2 s) {
  // clamps keep every intermediate finite (inf/inf, 0*inf); they change no result for
  // |y-mu|, sigma <= 1e30 and give the reference's limit values (L -> 0 -> bound) beyond.
  const F2 vc = min_nan2(v, 1e30f);
  const F2 sc = min_nan2(s, 1e30f);
  F2 a, b;
  div2_rn(add2(f2(0.5f), neg2(vc)), add2(f2(-0.5f), neg2(vc)), sc, a, b);
  const F2 c = f2(-0.70710678118654752440f);  // float(-(2 ** -0.5)) cast to fp32
  const F2 xa = mul2(c, a), xb = mul2(c, b);  // xb > 0 always; xa < 0 iff v < 0.5
  if (FAST) {
    F2 ea = erfc_pos_fast(min_nan2(abs2(xa), 12.0f));
    const F2 eb = erfc_pos_fast(min_nan2(xb, 12.0f));
    ea.x = (xa.x < 0.0f) ? 2.0f - ea.x : ea.x;
    ea.y = (xa.y < 0.0f) ? 2.0f - ea.y : ea.y;
    return fma2(f2(0.5f), ea, mul2(f2(-0.5f), eb));
  }
  const F2 upper = make_float2(0.5f * erfcf(xa.x), 0.5f * erfcf(xa.y));
  const F2 lower = make_float2(0.5f * erfcf(xb.x), 0.5f * erfcf(xb.y));
  return add2(upper, neg2(lower));
}


// Scalar conveniences for kernels that work one element at a time (STanH): the same packed
// code path with a dummy second lane, so both kernels share one arithmetic definition.
// L = Phi(n1/s) - Phi(n2/s) with Phi(a) = 0.5*erfc(-a/sqrt2); requires n1 >= n2.
template <bool FAST>
__device__ __forceinline__ float gauss_interval_mass(float n1, float n2, float s) {
  const F2 sc = min_nan2(f2(s), 1e30f);
  F2 a, b;
  div2_rn(f2(n1), f2(n2), sc, a, b);
  const F2 c = f2(-0.70710678118654752440f);
  const F2 xa = mul2(c, a), xb = mul2(c, b);
  if (FAST) {
    // pack (xa, xb) of the one element into the two lanes of a single erfc evaluation
    const F2 x = make_float2(min_nan(fabsf(xa.x), 12.0f), min_nan(fabsf(xb.x), 12.0f));
    F2 e = erfc_pos_fast(x);
    e.x = (xa.x < 0.0f) ? 2.0f - e.x : e.x;
    e.y = (xb.x < 0.0f) ? 2.0f - e.y : e.y;
    return fmaf(0.5f, e.x, -0.5f * e.y);
  }
  return 0.5f * erfcf(xa.x) - 0.5f * erfcf(xb.x);
}

}  // namespace reslic
