// eb_bwd.cu — backward of the fused factorized-bottleneck forward (sm_100a).
//
// What autograd does in the reference's training step through EntropyBottleneck.forward
// (src/entropy_models/adaptive_entropy_bottleneck.py:525-543 _logits_cumulative x2, :658-666 sign
// trick, LowerBound): gradient w.r.t. z and w.r.t. the 58 parameters of every channel.
//
// One CTA per channel.  Each thread walks its share of the channel's B*hw elements, re-runs the
// forward with a register tape and accumulates the element's contribution to all 58 parameter
// gradients in registers; the CTA then reduces them with a fixed shuffle/shared-memory tree and
// writes the channel's gradients once — no atomics, bit-reproducible.  FP32-issue bound (four MLP
// evaluations + two MLP backwards per element); z is 3.75 % of y's elements.
#include "common.cuh"
#include <cstdlib>
#include "eb_math.cuh"
#include "reslic_internal.h"

namespace reslic {

struct EbBwdParams {
  const float* z; const float* noise; int64_t z_bs, noise_bs;
  const float* matrix[5]; const float* bias[5]; const float* factor[4]; const float* medians;
  const float* g_zhat; const float* g_lik; int64_t g_zhat_bs, g_lik_bs;
  float* g_z; int64_t g_z_bs;
  float* g_matrix[5]; float* g_bias[5]; float* g_factor[4]; float* g_medians;
  int64_t B; int hw, C, noise_mode;
  float lik_bound;
  uint32_t seed_lo, seed_hi, off_lo, off_hi;
  const float* half_lo; const float* half_up; const int32_t* cell; int64_t half_lo_bs, half_up_bs, cell_bs;   // variable bins
  double* g_dist; int n_dist;    // optional: dLoss / d distance_points (summed per level gap)
  int identity;                  // mode RESLIC_Q_IDENTITY: z is the quantizer's output
  int splits;                    // CTAs per channel
  unsigned int* counters;        // [C] arrival tickets (zero between launches)
  float* partials;               // [C][splits][kPartStride]
};

constexpr int kPartStride = 64;
constexpr int kEbBwdMaxSplits = 8;
#ifndef RESLIC_EBB_THREADS
#define RESLIC_EBB_THREADS 128
#endif
#ifndef RESLIC_EBB_MINB
#define RESLIC_EBB_MINB 3
#endif
constexpr int kBT = RESLIC_EBB_THREADS;      // threads per CTA of eb_bwd_kernel
constexpr int kBMinB = RESLIC_EBB_MINB;      // CTAs per SM the register allocation must allow
constexpr int kNP = 58;   // transformed parameters per channel (median excluded)

// Forward with tape, then backward: adds g_out * d logits / d P[j] to gP[j], returns d logits / d x * g_out.
__device__ __forceinline__ float logits_backward(const float* __restrict__ P, float x, float g_out, float* gP) {
  float hin[4][3];      // input of layers 1..4 (output of layers 0..3)
  float th[4][3];       // tanh(t) of layers 0..3
  {
    float h[3], g[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float t = __fadd_rn(__fmul_rn(P[oM0 + j], x), P[oB0 + j]);
      th[0][j] = tanhf(t);
      h[j] = __fadd_rn(t, __fmul_rn(P[oF0 + j], th[0][j]));
      hin[0][j] = h[j];
    }
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      const float* M = P + oM1 + l * 15;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float t = fmaf(M[3 * j + 2], h[2], fmaf(M[3 * j + 1], h[1], __fmul_rn(M[3 * j], h[0])));
        t = __fadd_rn(t, M[9 + j]);
        th[l + 1][j] = tanhf(t);
        g[j] = __fadd_rn(t, __fmul_rn(M[12 + j], th[l + 1][j]));
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) { h[j] = g[j]; hin[l + 1][j] = g[j]; }
    }
  }
  // layer 4: out = M4 . h3 + b4
  float gh[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    gP[oM4 + k] = fmaf(g_out, hin[3][k], gP[oM4 + k]);
    gh[k] = P[oM4 + k] * g_out;
  }
  gP[oB4] += g_out;
  // layers 3..1
#pragma unroll
  for (int l = 2; l >= 0; --l) {
    const float* M = P + oM1 + l * 15;
    float* gM = gP + oM1 + l * 15;
    float gt[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float tj = th[l + 1][j];
      gM[12 + j] = fmaf(gh[j], tj, gM[12 + j]);                      // d/d a_j  (a = tanh(factor))
      gt[j] = gh[j] * fmaf(M[12 + j], 1.0f - tj * tj, 1.0f);         // h = t + a*tanh(t)
      gM[9 + j] += gt[j];                                            // bias
#pragma unroll
      for (int k = 0; k < 3; ++k) gM[3 * j + k] = fmaf(gt[j], hin[l][k], gM[3 * j + k]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) gh[k] = fmaf(M[6 + k], gt[2], fmaf(M[3 + k], gt[1], M[k] * gt[0]));
  }
  // layer 0: t = M0*x + b0
  float gx = 0.0f;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float tj = th[0][j];
    gP[oF0 + j] = fmaf(gh[j], tj, gP[oF0 + j]);
    const float gt = gh[j] * fmaf(P[oF0 + j], 1.0f - tj * tj, 1.0f);
    gP[oB0 + j] += gt;
    gP[oM0 + j] = fmaf(gt, x, gP[oM0 + j]);
    gx = fmaf(P[oM0 + j], gt, gx);
  }
  return gx;
}

// Fast form (math mode != MIRROR): ONE taped forward of both cumulative logits (x - 1/2, x + 1/2) as the two
// lanes of packed f32x2 operations (tanh3_pair of eb_math.cuh), then the backward of both lanes, packed where
// the lanes stay apart and as two scalar FMAs where they meet in a parameter-gradient accumulator.  The
// reference form above evaluates the forward four times per element (two for the value, one inside each
// logits_backward) with library tanhf.  g = (d loss / d lower, d loss / d upper); returns d loss / d x.
__device__ __forceinline__ float2 logits_pair_tape(const float* __restrict__ P, float2 x, float2 hin[4][3], float2 th[4][3]) {
  float2 h[3], t[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) t[j] = __ffma2_rn(make_float2(P[oM0 + j], P[oM0 + j]), x, make_float2(P[oB0 + j], P[oB0 + j]));
  tanh3_pair(t, th[0]);
#pragma unroll
  for (int j = 0; j < 3; ++j) { h[j] = __ffma2_rn(make_float2(P[oF0 + j], P[oF0 + j]), th[0][j], t[j]); hin[0][j] = h[j]; }
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    const float* M = P + oM1 + l * 15;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float2 a = __ffma2_rn(make_float2(M[3 * j], M[3 * j]), h[0], make_float2(M[9 + j], M[9 + j]));
      a = __ffma2_rn(make_float2(M[3 * j + 1], M[3 * j + 1]), h[1], a);
      t[j] = __ffma2_rn(make_float2(M[3 * j + 2], M[3 * j + 2]), h[2], a);
    }
    tanh3_pair(t, th[l + 1]);
#pragma unroll
    for (int j = 0; j < 3; ++j) { h[j] = __ffma2_rn(make_float2(M[12 + j], M[12 + j]), th[l + 1][j], t[j]); hin[l + 1][j] = h[j]; }
  }
  float2 a = __ffma2_rn(make_float2(P[oM4], P[oM4]), h[0], make_float2(P[oB4], P[oB4]));
  a = __ffma2_rn(make_float2(P[oM4 + 1], P[oM4 + 1]), h[1], a);
  return __ffma2_rn(make_float2(P[oM4 + 2], P[oM4 + 2]), h[2], a);
}

__device__ __forceinline__ float2 logits_backward_pair(const float* __restrict__ P, float2 x, float2 g, const float2 hin[4][3],
                                                       const float2 th[4][3], float* gP) {
  auto dot2 = [](float2 a, float2 b, float acc) { return fmaf(a.x, b.x, fmaf(a.y, b.y, acc)); };
  const float2 one = make_float2(1.0f, 1.0f);
  float2 gh[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    gP[oM4 + k] = dot2(g, hin[3][k], gP[oM4 + k]);
    gh[k] = __fmul2_rn(make_float2(P[oM4 + k], P[oM4 + k]), g);
  }
  gP[oB4] += g.x + g.y;
#pragma unroll
  for (int l = 2; l >= 0; --l) {
    const float* M = P + oM1 + l * 15;
    float* gM = gP + oM1 + l * 15;
    float2 gt[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float2 tj = th[l + 1][j];
      gM[12 + j] = dot2(gh[j], tj, gM[12 + j]);                                  // d/d a_j  (a = tanh(factor))
      const float2 sech2 = __ffma2_rn(make_float2(-tj.x, -tj.y), tj, one);
      gt[j] = __fmul2_rn(gh[j], __ffma2_rn(make_float2(M[12 + j], M[12 + j]), sech2, one));   // h = t + a*tanh(t)
      gM[9 + j] += gt[j].x + gt[j].y;                                            // bias
#pragma unroll
      for (int k = 0; k < 3; ++k) gM[3 * j + k] = dot2(gt[j], hin[l][k], gM[3 * j + k]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float2 a = __fmul2_rn(make_float2(M[k], M[k]), gt[0]);
      a = __ffma2_rn(make_float2(M[3 + k], M[3 + k]), gt[1], a);
      gh[k] = __ffma2_rn(make_float2(M[6 + k], M[6 + k]), gt[2], a);
    }
  }
  float2 gx = make_float2(0.0f, 0.0f);       // per lane: d loss / d (x - lo), d loss / d (x + up)
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float2 tj = th[0][j];
    gP[oF0 + j] = dot2(gh[j], tj, gP[oF0 + j]);
    const float2 sech2 = __ffma2_rn(make_float2(-tj.x, -tj.y), tj, one);
    const float2 gt = __fmul2_rn(gh[j], __ffma2_rn(make_float2(P[oF0 + j], P[oF0 + j]), sech2, one));
    gP[oB0 + j] += gt.x + gt.y;
    gP[oM0 + j] = dot2(gt, x, gP[oM0 + j]);
    gx = __ffma2_rn(make_float2(P[oM0 + j], P[oM0 + j]), gt, gx);
  }
  return gx;
}

// d(transformed)/d(raw) of staged slot j: sigmoid(m) for matrices, 1 - tanh(f)^2 for factors, 1 for biases
template <typename PP>
__device__ __forceinline__ float eb_param_chain(const PP& p, int c, int j, float** dst) {
  auto sig = [](float v) { return 1.0f / (1.0f + expf(-v)); };
  auto dth = [](float v) { const float t = tanhf(v); return 1.0f - t * t; };
  if (j < oB0) { *dst = p.g_matrix[0] + c * 3 + j; return sig(p.matrix[0][c * 3 + j]); }
  if (j < oF0) { *dst = p.g_bias[0] + c * 3 + (j - oB0); return 1.0f; }
  if (j < oM1) { *dst = p.g_factor[0] + c * 3 + (j - oF0); return dth(p.factor[0][c * 3 + (j - oF0)]); }
  if (j < oM4) {
    const int l = (j - oM1) / 15, r = (j - oM1) - l * 15;
    if (r < 9) { *dst = p.g_matrix[1 + l] + c * 9 + r; return sig(p.matrix[1 + l][c * 9 + r]); }
    if (r < 12) { *dst = p.g_bias[1 + l] + c * 3 + (r - 9); return 1.0f; }
    *dst = p.g_factor[1 + l] + c * 3 + (r - 12);
    return dth(p.factor[1 + l][c * 3 + (r - 12)]);
  }
  if (j < oB4) { *dst = p.g_matrix[4] + c * 3 + (j - oM4); return sig(p.matrix[4][c * 3 + (j - oM4)]); }
  *dst = p.g_bias[4] + c;
  return 1.0f;
}

// QUAD (fast math, in-kernel noise, hw % 4 == 0): one Philox call per four elements as in eb_fwd_fast_kernel — the four
// lanes of a quad sit on the four elements of one counter, each evaluates the counter of a different one of its next
// four loop iterations and a 4 x 4 transpose across the quad hands out the words (same field as the per-element form).
template <bool FAST, bool QUAD = false>
__global__ void __launch_bounds__(kBT, kBMinB) eb_bwd_kernel(const EbBwdParams p) {
  __shared__ float s_par[kEbStride + 1];
  __shared__ float s_red[kBT / 32][kNP + 1];
  __shared__ bool s_last;
  __shared__ double s_dist[1024];            // per-CTA sums of the half-width gradients (variable bins only)
  if (p.g_dist)
    for (int i = threadIdx.x; i < p.n_dist; i += kBT) s_dist[i] = 0.0;
  const int c = blockIdx.x / p.splits;
  const int split = blockIdx.x - c * p.splits;
  if (threadIdx.x < kEbStride) s_par[threadIdx.x] = eb_staged_param(p, c, threadIdx.x);
  __syncthreads();
  const float med = s_par[oMed];
  float gP[kNP];
#pragma unroll
  for (int j = 0; j < kNP; ++j) gP[j] = 0.0f;
  float gmed = 0.0f;
  const int64_t total = p.B * p.hw;
  const int64_t base_c = static_cast<int64_t>(c) * p.hw;
  // The inputs of the NEXT element are loaded before the current one is evaluated: with one CTA of this kernel
  // per SM (8 warps) nothing else hides the ~0.8 us of a global load, and an element row is only ~0.8 us of math.
  struct Ld { float zv, u, gl, gzh, lo, up; int cell; int64_t b, e; };
  const int64_t step = static_cast<int64_t>(p.splits) * kBT;
  auto fetch = [&](int64_t idx, Ld& r) {
    r.b = idx / p.hw;
    r.e = base_c + (idx - r.b * p.hw);                                   // offset inside image b
    r.zv = p.z[r.b * p.z_bs + r.e];
    r.u = (p.noise_mode && p.noise) ? p.noise[r.b * p.noise_bs + r.e] : 0.0f;
    r.gl = p.g_lik ? p.g_lik[r.b * p.g_lik_bs + r.e] : 0.0f;
    r.gzh = p.g_zhat ? p.g_zhat[r.b * p.g_zhat_bs + r.e] : 0.0f;
    r.lo = p.half_lo ? p.half_lo[r.b * p.half_lo_bs + r.e] : 0.5f;
    r.up = p.half_up ? p.half_up[r.b * p.half_up_bs + r.e] : 0.5f;
    r.cell = p.cell ? p.cell[r.b * p.cell_bs + r.e] : -1;
  };
  Ld nxt{};
  int64_t idx = static_cast<int64_t>(split) * kBT + threadIdx.x;
  if (idx < total) fetch(idx, nxt);
  float uq[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  int64_t idx_w = idx - (threadIdx.x & 31);            // lane 0's index: the QUAD loop runs warp-uniformly (it shuffles)
  for (int nq = 0; (QUAD ? idx_w : idx) < total; idx += step, idx_w += step, ++nq) {
    if (QUAD && (nq & 3) == 0) {
      const int k = threadIdx.x & 3;
      int64_t ik = idx + k * step;
      if (ik >= total) ik = idx < total ? idx : 0;      // past the end: any counter, the word is not used
      const int64_t bk = ik / p.hw;
      const uint64_t eid = static_cast<uint64_t>(bk) * static_cast<uint64_t>(static_cast<int64_t>(p.C) * p.hw) +
                           static_cast<uint64_t>(base_c + (ik - bk * p.hw));
      const uint64_t gid = eid >> 2;
      const Philox4 r = philox4x32_10(static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32), p.off_lo, p.off_hi, p.seed_lo, p.seed_hi);
      float a[4] = {u32_to_centered_uniform(r.x), u32_to_centered_uniform(r.y), u32_to_centered_uniform(r.z), u32_to_centered_uniform(r.w)};
      const bool odd = (k & 1) != 0, hi = (k & 2) != 0;
      float r0 = __shfl_xor_sync(0xffffffffu, odd ? a[0] : a[1], 1), r1 = __shfl_xor_sync(0xffffffffu, odd ? a[2] : a[3], 1);
      if (odd) { a[0] = r0; a[2] = r1; } else { a[1] = r0; a[3] = r1; }
      r0 = __shfl_xor_sync(0xffffffffu, hi ? a[0] : a[2], 2); r1 = __shfl_xor_sync(0xffffffffu, hi ? a[1] : a[3], 2);
      if (hi) { a[0] = r0; a[1] = r1; } else { a[2] = r0; a[3] = r1; }
      uq[0] = a[0]; uq[1] = a[1]; uq[2] = a[2]; uq[3] = a[3];
    }
    if (QUAD && idx >= total) continue;
    const Ld cur = nxt;
    if (idx + step < total) fetch(idx + step, nxt);
    const int64_t b = cur.b, e = cur.e;
    const float zv = cur.zv;
    float x;
    if (p.identity) {
      x = zv;
    } else if (p.noise_mode) {
      float u = cur.u;
      if (QUAD) {
        u = uq[0]; uq[0] = uq[1]; uq[1] = uq[2]; uq[2] = uq[3];
      } else if (!p.noise) {
        const uint64_t eid = static_cast<uint64_t>(b) * static_cast<uint64_t>(static_cast<int64_t>(p.C) * p.hw) +
                             static_cast<uint64_t>(e);
        const uint64_t gid = eid >> 2;
        const Philox4 r = philox4x32_10(static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32),
                                        p.off_lo, p.off_hi, p.seed_lo, p.seed_hi);
        const int k = static_cast<int>(eid & 3);
        u = u32_to_centered_uniform(k == 0 ? r.x : k == 1 ? r.y : k == 2 ? r.z : r.w);
      }
      x = zv + u;
    } else {
      x = rintf(zv - med) + med;
    }
    float gx = 0.0f;
    if (p.g_lik) {
      const float gl = cur.gl;
      float2 hin[4][3], th[4][3];
      const float2 xx = make_float2(x - cur.lo, x + cur.up);
      float lower, upper;
      if (FAST) { const float2 lu = logits_pair_tape(s_par, xx, hin, th); lower = lu.x; upper = lu.y; }
      else { lower = logits_cumulative(s_par, xx.x); upper = logits_cumulative(s_par, xx.y); }
      const float sum = lower + upper;
      const float sg = (sum < 0.0f) ? 1.0f : ((sum > 0.0f) ? -1.0f : 0.0f);
      float su, sl;
      if (FAST) {
        su = rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * (sg * upper)));
        sl = rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * (sg * lower)));
      } else { su = sigmoid_ref(sg * upper); sl = sigmoid_ref(sg * lower); }
      const float D = su - sl;
      const float L = fabsf(D);
      const bool pass = !(p.lik_bound > 0.0f) || (L >= p.lik_bound) || (gl < 0.0f);
      const float g = pass ? gl : 0.0f;
      const float sd = (D > 0.0f) ? 1.0f : ((D < 0.0f) ? -1.0f : 0.0f);
      const float du = g * sd * sg * su * (1.0f - su);
      const float dl = -g * sd * sg * sl * (1.0f - sl);
      if (du != 0.0f || dl != 0.0f) {
        float2 g2;
        if (FAST) g2 = logits_backward_pair(s_par, xx, make_float2(dl, du), hin, th, gP);
        else {
          g2.y = logits_backward(s_par, xx.y, du, gP);
          g2.x = logits_backward(s_par, xx.x, dl, gP);
        }
        gx = g2.x + g2.y;
        if (p.g_dist && cur.cell >= 0) {       // x - lo and x + up: d/d lo = -lane 0, d/d up = +lane 1
          if (cur.cell > 0 && cur.cell - 1 < p.n_dist) atomicAdd(&s_dist[cur.cell - 1], -static_cast<double>(g2.x));
          if (cur.cell < p.n_dist) atomicAdd(&s_dist[cur.cell], static_cast<double>(g2.y));
        }
      }
    }
    const float gzh = cur.gzh;
    if (p.noise_mode || p.identity) {
      if (p.g_z) p.g_z[b * p.g_z_bs + e] = gzh + gx;
    } else {
      if (p.g_z) p.g_z[b * p.g_z_bs + e] = 0.0f;                    // round() has zero gradient
      gmed += gzh + gx;                                              // z_hat = round(z - med) + med
    }
  }
  if (p.g_dist) {        // K-sized: a double atomic per touched level gap and CTA
    __syncthreads();
    for (int i = threadIdx.x; i < p.n_dist; i += kBT)
      if (s_dist[i] != 0.0) atomicAdd(&p.g_dist[i], s_dist[i]);
  }
  // deterministic CTA reduction of the 58 (+1) per-thread sums
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < kNP; ++j) {
    const float v = warp_sum_f32(gP[j]);
    if (lane == 0) s_red[warp][j] = v;
  }
  {
    const float v = warp_sum_f32(gmed);
    if (lane == 0) s_red[warp][kNP] = v;
  }
  __syncthreads();
  float v = 0.0f;
  if (threadIdx.x <= kNP)
    for (int w = 0; w < kBT / 32; ++w) v += s_red[w][threadIdx.x];
  if (p.splits > 1) {
    // the channel's CTAs leave their sums in the workspace; the one that arrives last adds them in split order
    // (a fixed order whoever is last: bit-reproducible) and writes the gradients
    float* part = p.partials + (static_cast<int64_t>(c) * p.splits) * kPartStride;
    if (threadIdx.x <= kNP) part[split * kPartStride + threadIdx.x] = v;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&p.counters[c], 1u) == static_cast<unsigned int>(p.splits - 1));
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x <= kNP) {
      v = 0.0f;
      for (int sp = 0; sp < p.splits; ++sp) v += __ldcg(part + sp * kPartStride + threadIdx.x);
    }
    if (threadIdx.x == 0) p.counters[c] = 0u;
  }
  if (threadIdx.x <= kNP) {
    if (threadIdx.x < kNP) {
      if (p.g_matrix[0]) {
        float* dst;
        const float chain = eb_param_chain(p, c, threadIdx.x, &dst);
        *dst = v * chain;
      }
    } else if (p.g_medians) {
      p.g_medians[c] = v;
    }
  }
}

static int64_t eb_bwd_counter_bytes(int64_t C) { return ((C * 4 + 255) / 256) * 256; }
int64_t eb_bwd_workspace_bytes(int64_t C) {
  return eb_bwd_counter_bytes(C) + C * kEbBwdMaxSplits * kPartStride * static_cast<int64_t>(sizeof(float));
}

int eb_bwd_launch(const reslic_eb_bwd_desc* d, cudaStream_t st) {
  if (!d) return set_error(RESLIC_ERR_ARG, "eb_bwd: null descriptor");
  if (d->struct_size != sizeof(reslic_eb_bwd_desc))
    return set_error(RESLIC_ERR_ARG, "eb_bwd: struct_size != sizeof(reslic_eb_bwd_desc) (binding built against another ABI revision)");
  if (d->B < 0 || d->C < 0 || d->hw < 0) return set_error(RESLIC_ERR_ARG, "eb_bwd: negative size");
  if (d->C == 0) return RESLIC_OK;
  if (d->C > (1 << 20) || d->hw > (1LL << 30)) return set_error(RESLIC_ERR_ARG, "eb_bwd: size too large");
  if (d->mode != RESLIC_Q_DEQUANTIZE && d->mode != RESLIC_Q_NOISE && d->mode != RESLIC_Q_IDENTITY)
    return set_error(RESLIC_ERR_ARG, "eb_bwd: invalid quantization mode");
  if (d->g_dist && (!d->cell || d->n_dist < 1 || d->n_dist > 1024))
    return set_error(RESLIC_ERR_ARG, "eb_bwd: g_dist needs `cell` and 1 <= n_dist <= 1024");
  if ((d->B > 0 && d->hw > 0 && !d->z) || !d->medians) return set_error(RESLIC_ERR_ARG, "eb_bwd: z or medians is null");
  bool any_g = false, all_g = true;
  for (int i = 0; i < 5; ++i) {
    if (!d->matrix[i] || !d->bias[i] || (i < 4 && !d->factor[i]))
      return set_error(RESLIC_ERR_ARG, "eb_bwd: a parameter pointer is null");
    const bool have = d->g_matrix[i] && d->g_bias[i] && (i == 4 || d->g_factor[i]);
    any_g |= (d->g_matrix[i] || d->g_bias[i] || (i < 4 && d->g_factor[i]));
    all_g &= have;
  }
  if (any_g && !all_g) return set_error(RESLIC_ERR_ARG, "eb_bwd: parameter gradients must be requested as a full set");
  if (!any_g && !d->g_z && !d->g_medians) return set_error(RESLIC_ERR_ARG, "eb_bwd: no output requested");
  EbBwdParams p{};
  p.z = d->z; p.z_bs = d->z_bs; p.noise = d->noise; p.noise_bs = d->noise_bs;
  for (int i = 0; i < 5; ++i) { p.matrix[i] = d->matrix[i]; p.bias[i] = d->bias[i]; p.g_matrix[i] = d->g_matrix[i]; p.g_bias[i] = d->g_bias[i]; }
  for (int i = 0; i < 4; ++i) { p.factor[i] = d->factor[i]; p.g_factor[i] = d->g_factor[i]; }
  p.medians = d->medians; p.g_medians = d->g_medians;
  p.g_zhat = d->g_zhat; p.g_zhat_bs = d->g_zhat_bs; p.g_lik = d->g_lik; p.g_lik_bs = d->g_lik_bs;
  p.g_z = d->g_z; p.g_z_bs = d->g_z_bs;
  p.B = d->B; p.hw = static_cast<int>(d->hw); p.C = static_cast<int>(d->C);
  p.noise_mode = d->mode == RESLIC_Q_NOISE; p.identity = d->mode == RESLIC_Q_IDENTITY; p.lik_bound = d->likelihood_bound;
  p.half_lo = d->half_lo; p.half_up = d->half_up; p.cell = d->cell;
  p.half_lo_bs = d->half_lo_bs; p.half_up_bs = d->half_up_bs; p.cell_bs = d->cell_bs;
  p.g_dist = d->g_dist; p.n_dist = static_cast<int>(d->n_dist);
  p.seed_lo = static_cast<uint32_t>(d->philox_seed); p.seed_hi = static_cast<uint32_t>(d->philox_seed >> 32);
  p.off_lo = static_cast<uint32_t>(d->philox_offset); p.off_hi = static_cast<uint32_t>(d->philox_offset >> 32);
  // CTAs per channel: as many as keep the whole grid inside ONE resident wave (a CTA is a channel's slice of the batch,
  // so a partial second wave costs a full CTA time: 576 CTAs on 444 slots ran 50.7 us, 384 run 41.7 us on 192 channels)
  int64_t splits = 1;
  if (d->workspace) {
    if (d->workspace_bytes < eb_bwd_workspace_bytes(d->C) || (reinterpret_cast<uintptr_t>(d->workspace) & 15u))
      return set_error(RESLIC_ERR_WORKSPACE, "eb_bwd: workspace misaligned or smaller than reslic_eb_bwd_workspace_bytes(C)");
    static const int resident[2] = {
      [] { int n = 0; return (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, eb_bwd_kernel<false>, kBT, 0) == cudaSuccess && n > 0) ? n : 1; }(),
      [] { int n = 0; return (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, eb_bwd_kernel<true>, kBT, 0) == cudaSuccess && n > 0) ? n : 1; }()};
    splits = static_cast<int64_t>(sm_count()) * resident[math_mode() == RESLIC_MATH_MIRROR ? 0 : 1] / d->C;
    static const long forced = [] { const char* e = std::getenv("RESLIC_EBB_SPLITS"); return e ? std::atol(e) : 0L; }();
    if (forced >= 1) splits = forced;
    const int64_t per_thread = (d->B * d->hw + kBT - 1) / kBT;     // elements per thread of an unsplit channel
    if (splits > per_thread) splits = per_thread;
    if (splits > kEbBwdMaxSplits) splits = kEbBwdMaxSplits;
    if (splits < 1) splits = 1;
    p.counters = static_cast<unsigned int*>(d->workspace);
    p.partials = reinterpret_cast<float*>(static_cast<char*>(d->workspace) + eb_bwd_counter_bytes(d->C));
  }
  p.splits = static_cast<int>(splits);
  if (d->C * splits > 0x7fffffffLL) return set_error(RESLIC_ERR_ARG, "eb_bwd: grid too large");
  const int grid = static_cast<int>(d->C * splits);
  if (math_mode() == RESLIC_MATH_MIRROR) eb_bwd_kernel<false><<<grid, kBT, 0, st>>>(p);
  else if (p.noise_mode && !p.noise && !p.identity && d->hw % 4 == 0) eb_bwd_kernel<true, true><<<grid, kBT, 0, st>>>(p);
  else eb_bwd_kernel<true><<<grid, kBT, 0, st>>>(p);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_cuda_error(err, "eb_bwd launch");
  return RESLIC_OK;
}

}  // namespace reslic
