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
