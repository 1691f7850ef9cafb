// eb_math.cuh — per-channel cumulative-logit MLP of the factorized bottleneck, shared by the
// forward and backward kernels (src/entropy_models/adaptive_entropy_bottleneck.py:525-543, 658-666).
#pragma once
#include "common.cuh"

namespace reslic {

constexpr int kEbMaxCh = 64;     // channels whose parameters fit the per-tile staging area
constexpr int kEbStride = 59;    // 33 matrix + 13 bias + 12 factor + 1 median (odd: bank-conflict free)


__device__ __forceinline__ float softplus_ref(float x) {  // F.softplus(beta=1, threshold=20)
  return x > 20.0f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float sigmoid_ref(float x) { return 1.0f / (1.0f + expf(-x)); }

// offsets inside one channel's staged parameter block
//   M0[3] B0[3] F0[3] | M1[9] B1[3] F1[3] | M2[9] B2[3] F2[3] | M3[9] B3[3] F3[3] | M4[3] B4[1] | med
constexpr int oM0 = 0, oB0 = 3, oF0 = 6, oM1 = 9, oM4 = 54, oB4 = 57, oMed = 58;

// adaptive_entropy_bottleneck.py:525-543 for one scalar input, filters (3,3,3,3).
// matmul rows are fma chains, "+= bias" and "+= tanh(f)*tanh(x)" keep the reference's
// separate roundings.
__device__ __forceinline__ float logits_cumulative(const float* __restrict__ P, float x) {
  float h[3], g[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float t = __fadd_rn(__fmul_rn(P[oM0 + j], x), P[oB0 + j]);
    h[j] = __fadd_rn(t, __fmul_rn(P[oF0 + j], tanhf(t)));
  }
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    const float* M = P + oM1 + l * 15;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float t = fmaf(M[3 * j + 2], h[2], fmaf(M[3 * j + 1], h[1], __fmul_rn(M[3 * j], h[0])));
      t = __fadd_rn(t, M[9 + j]);
      g[j] = __fadd_rn(t, __fmul_rn(M[12 + j], tanhf(t)));
    }
    h[0] = g[0]; h[1] = g[1]; h[2] = g[2];
  }
  const float t = fmaf(P[oM4 + 2], h[2], fmaf(P[oM4 + 1], h[1], __fmul_rn(P[oM4], h[0])));
  return __fadd_rn(t, P[oB4]);
}

// ---- fast form: both cumulative logits of an element (x - 1/2, x + 1/2) as the two lanes of packed
// f32x2 operations, so every staged parameter is read once for the pair and every fp32 op serves both.
// tanh(t) = sign(t) (1 - E) / (1 + E), E = 2^(-2 log2e |t|) <= 1: the three tanh of a layer share ONE
// reciprocal of (1+E0)(1+E1)(1+E2) in [1, 8]  (MUFU per element: 24 EX2 + 8 RCP instead of 24 + 24; the
// MUFU pipe is what bounds this kernel).  Absolute error of each tanh <= 2e-7; bias adds are folded into
// the fma chains.  Used when the math mode is not RESLIC_MATH_MIRROR.
__device__ __forceinline__ void tanh3_pair(const float2 t[3], float2 th[3]) {
  float2 n[3], d[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float ex = ex2_approx(-2.8853900817779268f * fabsf(t[j].x));
    const float ey = ex2_approx(-2.8853900817779268f * fabsf(t[j].y));
    n[j] = make_float2(1.0f - ex, 1.0f - ey);
    d[j] = make_float2(1.0f + ex, 1.0f + ey);
  }
  const float2 d01 = __fmul2_rn(d[0], d[1]), d12 = __fmul2_rn(d[1], d[2]), d02 = __fmul2_rn(d[0], d[2]);
  const float2 d012 = __fmul2_rn(d01, d[2]);
  const float2 r = make_float2(rcp_approx(d012.x), rcp_approx(d012.y));
  const float2 m0 = __fmul2_rn(__fmul2_rn(n[0], d12), r);
  const float2 m1 = __fmul2_rn(__fmul2_rn(n[1], d02), r);
  const float2 m2 = __fmul2_rn(__fmul2_rn(n[2], d01), r);
  th[0] = make_float2(copysignf(m0.x, t[0].x), copysignf(m0.y, t[0].y));
  th[1] = make_float2(copysignf(m1.x, t[1].x), copysignf(m1.y, t[1].y));
  th[2] = make_float2(copysignf(m2.x, t[2].x), copysignf(m2.y, t[2].y));
}

// (logits_cumulative(P, x.x), logits_cumulative(P, x.y)) up to the fast-math differences above
__device__ __forceinline__ float2 logits_cumulative_pair(const float* __restrict__ P, float2 x) {
  float2 h[3], t[3], th[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) t[j] = __ffma2_rn(make_float2(P[oM0 + j], P[oM0 + j]), x, make_float2(P[oB0 + j], P[oB0 + j]));
  tanh3_pair(t, th);
#pragma unroll
  for (int j = 0; j < 3; ++j) h[j] = __ffma2_rn(make_float2(P[oF0 + j], P[oF0 + j]), th[j], t[j]);
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    const float* M = P + oM1 + l * 15;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float2 a = __ffma2_rn(make_float2(M[3 * j], M[3 * j]), h[0], make_float2(M[9 + j], M[9 + j]));
      a = __ffma2_rn(make_float2(M[3 * j + 1], M[3 * j + 1]), h[1], a);
      t[j] = __ffma2_rn(make_float2(M[3 * j + 2], M[3 * j + 2]), h[2], a);
    }
    tanh3_pair(t, th);
#pragma unroll
    for (int j = 0; j < 3; ++j) h[j] = __ffma2_rn(make_float2(M[12 + j], M[12 + j]), th[j], t[j]);
  }
  float2 a = __ffma2_rn(make_float2(P[oM4], P[oM4]), h[0], make_float2(P[oB4], P[oB4]));
  a = __ffma2_rn(make_float2(P[oM4 + 1], P[oM4 + 1]), h[1], a);
  return __ffma2_rn(make_float2(P[oM4 + 2], P[oM4 + 2]), h[2], a);
}

// sign-trick likelihood of x from the pair (lower, upper) = logits at (x - 1/2, x + 1/2); sigmoid as
// 1 / (1 + 2^(-a log2e)) with the approximate MUFU forms (overflow -> 1/inf = 0, as the exact form)
__device__ __forceinline__ float eb_likelihood_fast(const float* __restrict__ P, float x, float lik_floor) {
  const float2 lu = logits_cumulative_pair(P, make_float2(x - 0.5f, x + 0.5f));
  const float sum = lu.x + lu.y;
  const float sg = (sum < 0.0f) ? 1.0f : ((sum > 0.0f) ? -1.0f : 0.0f);  // -torch.sign(sum); NaN -> 0
  const float su = rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * (sg * lu.y)));
  const float sl = rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * (sg * lu.x)));
  return max_nan(fabsf(su - sl), lik_floor);
}

// transformed parameter j of channel c (softplus on matrices, tanh on factors, median last)
template <typename P>
__device__ __forceinline__ float eb_staged_param(const P& p, int c, int j) {
  if (j < oB0) return softplus_ref(p.matrix[0][c * 3 + j]);
  if (j < oF0) return p.bias[0][c * 3 + (j - oB0)];
  if (j < oM1) return tanhf(p.factor[0][c * 3 + (j - oF0)]);
  if (j < oM4) {
    const int l = (j - oM1) / 15, r = (j - oM1) - l * 15;
    if (r < 9) return softplus_ref(p.matrix[1 + l][c * 9 + r]);
    if (r < 12) return p.bias[1 + l][c * 3 + (r - 9)];
    return tanhf(p.factor[1 + l][c * 3 + (r - 12)]);
  }
  if (j < oB4) return softplus_ref(p.matrix[4][c * 3 + (j - oM4)]);
  if (j == oB4) return p.bias[4][c];
  return p.medians[c];
}

// sign-trick likelihood from the two cumulative logits, adaptive_entropy_bottleneck.py:658-666
__device__ __forceinline__ float eb_combine(float lower, float upper, float lik_bound) {
  const float sum = lower + upper;
  const float sg = (sum < 0.0f) ? 1.0f : ((sum > 0.0f) ? -1.0f : 0.0f);  // -torch.sign(sum); NaN -> 0
  float L = fabsf(sigmoid_ref(sg * upper) - sigmoid_ref(sg * lower));
  if (lik_bound > 0.0f) L = max_nan(L, lik_bound);
  return L;
}
__device__ __forceinline__ float eb_likelihood(const float* __restrict__ P, float x, float lik_bound) {
  return eb_combine(logits_cumulative(P, x - 0.5f), logits_cumulative(P, x + 0.5f), lik_bound);
}


}  // namespace reslic
