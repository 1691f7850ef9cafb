// stanh_fused.cu — fused STanH ("sum of tanh") quantizer + variable-bin Gaussian likelihood
// (sm_100a).  Replaces the reference's GaussianConditionalStanh.forward / quantize / _likelihood
// (src/entropy_models/adaptive_gaussian_conditional.py:95-157, 495-603), the STanH activation
// itself (src/quantization/activation.py:135-150 NonSym, :294-304 Sym) and compute_gap
// (src/models/stanh/tcm_stanh.py:465-478).
//
// The reference evaluates the quantizer as a dense [1, K, N] broadcast (K = 2*extrema = 160
// thresholds per element: O(K*N) flops AND memory) and finds the likelihood bin with two
// [N, K+1] one-hot matrices.  Here every element does
//   * hard form (beta = -1): one count "how many sorted thresholds lie below x"
//     (value = level[#{k : x > b_k}]);
//   * soft form (finite beta): tanh(beta*(x-b_k)) is exactly +-1 in fp32 once |beta*(x-b_k)| > 9.1
//     (SURVEY.md §7.4 H1), so only the thresholds inside that window are evaluated and the
//     saturated rest comes from the prefix sums the module already keeps (cum_w);
//   * likelihood bin: one more count over the level mid-points (average_points), then the
//     same exact-division + erfc arithmetic as the plain Gaussian-conditional kernel.
// 12 B read + 8 B written per element, no intermediate tensors.
//
// Two generations of kernels live here.  The 128-bit ones (stanh_gc_vec_kernel, stanh_act_kernel,
// stanh_gc_bwd_kernel) answer the counts through the cell grids of stanh_tables.cuh (one LDS.U16 + two
// compares) and are what aligned tensors take; the scalar ones (stanh_gc_fwd_kernel, eb_stanh_fwd_kernel:
// binary searches, one element per thread) serve unaligned tensors, n % 4 != 0, a beta <= 0 other than
// -1, and the factorized STanH bottleneck.  DESIGN.md section 3.3 has the measurements.
#include "common.cuh"
#include "eb_math.cuh"
#include "gc_math.cuh"
#include "reslic_internal.h"
#include "stanh_tables.cuh"
#include <limits>
#include <type_traits>

namespace reslic {

constexpr int kStanhMaxK = 1024;          // thresholds (extrema <= 512)
constexpr float kSatT = 9.1f;             // |t| beyond which 2*sigmoid(2t)-1 == +-1 in fp32

struct StanhParams {
  const float* y; const float* mu; const float* sigma;
  int64_t y_bs, mu_bs, sigma_bs;
  float* yhat; float* lik; int32_t* sym; float* ste;
  int64_t yhat_bs, lik_bs, sym_bs, ste_bs;
  double* bits; unsigned long long* workspace; int bits_accumulate;
  const float* b; const float* w; const float* cum_w; const float* avg; const float* dist;
  int K, steps, symmetric, removing_mean, training, sym_offset;
  float beta, scale_bound, lik_bound;
  int64_t n, B, tiles_per_image;
  double* gap;       // activation/gap kernel: [0] sum (x - soft)^2, [1] sum (x - hard)^2
};

struct StanhTables {
  const float* b; const float* w; const float* cw; const float* avg; const float* dist;
  int K, steps;
};

// #{k < K : x > t[k]} for ascending t (NaN -> 0)
__device__ __forceinline__ int count_gt(float x, const float* t, int K, int steps) {
  int lo = 0;
  for (int step = 1 << (steps - 1); step > 0; step >>= 1) {
    const int i = lo + step;
    if (i <= K && x > t[i - 1]) lo = i;
  }
  return lo;
}
// #{k < K : x >= t[k]}
__device__ __forceinline__ int count_ge(float x, const float* t, int K, int steps) {
  int lo = 0;
  for (int step = 1 << (steps - 1); step > 0; step >>= 1) {
    const int i = lo + step;
    if (i <= K && x >= t[i - 1]) lo = i;
  }
  return lo;
}

// activation.py:143 (NonSym: sum_k w_k*relu(sign(x-b_k)) - w_k/2) and :298 (Sym: sum_k w_k/2*sign(x-b_k)).
// Both equal the level cum_w[c] with c = #{k : x > b_k}; the symmetric form gives the mean of the two
// adjacent levels at an exact tie x == b_k (sign(0) = 0).  `c_out` is the level index ("symbols").
__device__ __forceinline__ float stanh_hard(float x, const StanhTables& T, bool symmetric, int& c_out) {
  const int c = count_gt(x, T.b, T.K, T.steps);
  c_out = c;
  float v = T.cw[c];
  if (symmetric) {
    const int ce = count_ge(x, T.b, T.K, T.steps);
    v = 0.5f * (v + T.cw[ce]);
  }
  // NaN input: torch.sign(NaN) == 0, so the reference lands on the lowest level (NonSym: every
  // relu(sign) term is 0) or on 0 (Sym: every sign term is 0) — not on NaN.
  return (symmetric && x != x) ? 0.0f : v;
}

// activation.py:146-149 / :301-304: sum_k (w_k/2) * (2*sigmoid(2*beta*(x-b_k)) - 1)
__device__ __forceinline__ float stanh_soft(float x, float beta, const StanhTables& T) {
  int lo = 0, hi = T.K;
  float sat = 0.0f;
  if (beta > 0.0f) {
    const float r = kSatT / beta;
    lo = count_ge(x - r, T.b, T.K, T.steps);      // k < lo : beta*(x-b_k) >= T  -> +1
    hi = count_gt(x + r, T.b, T.K, T.steps);      // k >= hi: beta*(x-b_k) <= -T -> -1
    if (hi < lo) hi = lo;
    // sum_{k<lo} w_k/2 - sum_{k>=hi} w_k/2 from the prefix sums W(c) = cum_w[c] - cum_w[0]
    sat = 0.5f * ((T.cw[lo] - T.cw[0]) - (T.cw[T.K] - T.cw[hi]));
  }
  float acc = 0.0f;
  for (int k = lo; k < hi; ++k) {
    const float t = beta * (x - T.b[k]);
    const float sg = 1.0f / (1.0f + expf(-(2.0f * t)));
    acc = fmaf(T.w[k] * 0.5f, 2.0f * sg - 1.0f, acc);
  }
  return (x != x) ? x : sat + acc;
}

// adaptive_gaussian_conditional.py:495-580: half-widths (low, up) of the level cell that contains v,
// then the mass of [level-low, level+up] under N(0, s) in the sign-dependent form of :564-567.
template <bool FAST>
__device__ __forceinline__ float stanh_likelihood(float v, float s, const StanhTables& T) {
  const int j = count_gt(v, T.avg, T.K, T.steps);          // avg_left[j] < v <= avg_right[j]
  const bool inside = (v > -1000.0f) && (v <= 1000.0f);    // the reference's +-1000 sentinels (:506,:511)
  const float low = (inside && j > 0) ? T.dist[j - 1] : 0.0f;
  const float up = (inside && j < T.K) ? T.dist[j] : 0.0f;
  float n1, n2;
  if (v >= 0.0f) { n1 = low - v; n2 = -up - v; }
  else { n1 = v + up; n2 = v - low; }                       // NaN v lands here and stays NaN
  n1 = max_nan(min_nan(n1, 1e30f), -1e30f);
  n2 = max_nan(min_nan(n2, 1e30f), -1e30f);
  return gauss_interval_mass<FAST>(n1, n2, s);
}

__device__ __forceinline__ void stage_tables(const StanhParams& p, float* sm, StanhTables& T) {
  float* sb = sm; float* sw = sb + p.K; float* scw = sw + p.K; float* savg = scw + p.K + 1; float* sdist = savg + p.K;
  for (int i = threadIdx.x; i < p.K; i += blockDim.x) {
    sb[i] = p.b[i]; sw[i] = p.w[i]; savg[i] = p.avg[i]; sdist[i] = p.dist[i];
  }
  for (int i = threadIdx.x; i <= p.K; i += blockDim.x) scw[i] = p.cum_w[i];
  __syncthreads();
  T.b = sb; T.w = sw; T.cw = scw; T.avg = savg; T.dist = sdist; T.K = p.K; T.steps = p.steps;
}

template <bool FAST>
__global__ void __launch_bounds__(kThreads) stanh_gc_fwd_kernel(const StanhParams p) {
  extern __shared__ float sm[];
  StanhTables T;
  stage_tables(p, sm, T);
  const bool need_lik = p.lik || p.bits;
  const int64_t total = p.tiles_per_image * p.B;
  int64_t t = (static_cast<int64_t>(blockIdx.x) * total) / gridDim.x;
  const int64_t t_end = (static_cast<int64_t>(blockIdx.x + 1) * total) / gridDim.x;
  while (t < t_end) {
    const int image = static_cast<int>(t / p.tiles_per_image);
    const int64_t seg_end = (static_cast<int64_t>(image) + 1) * p.tiles_per_image;
    const int64_t stop = seg_end < t_end ? seg_end : t_end;
    float acc = 0.0f;
    for (; t < stop; ++t) {
      const int64_t e = (t - image * p.tiles_per_image) * kThreads + threadIdx.x;
      if (e >= p.n) continue;
      const float y = ld_stream1(p.y + image * p.y_bs + e);
      const float mu = p.mu ? ld_stream1(p.mu + image * p.mu_bs + e) : 0.0f;
      float yhat; int level = 0;
      if (p.training == 2) {
        yhat = y;                                            // likelihood of given values (_likelihood)
      } else if (p.training) {
        // quantize(..., "training"): honours removing_mean (:108-117)
        const float x = (p.mu && p.removing_mean) ? y - mu : y;
        const float q = (p.beta == -1.0f) ? stanh_hard(x, T, p.symmetric, level) : stanh_soft(x, p.beta, T);
        yhat = (p.mu && p.removing_mean) ? q + mu : q;
        if (p.sym) { int c; stanh_hard(y - mu, T, p.symmetric, c); level = c; }
      } else {
        // "dequantize" / "symbols": always about the mean, hard levels (:119-137)
        const float q = stanh_hard(y - mu, T, p.symmetric, level);
        yhat = p.mu ? q + mu : q;
      }
      if (p.yhat) st_stream1(p.yhat + image * p.yhat_bs + e, yhat);
      if (p.ste) st_stream1(p.ste + image * p.ste_bs + e, rintf(y - mu) + mu);     // ste_round(y - mu) + mu, tcm_stanh.py:434
      if (p.sym) st_stream1(p.sym + image * p.sym_bs + e, level + p.sym_offset);
      if (need_lik) {
        const float s = max_nan(ld_stream1(p.sigma + image * p.sigma_bs + e), p.scale_bound);
        const float v = p.mu ? yhat - mu : yhat;             // :547-550
        float L = stanh_likelihood<FAST>(v, s, T);
        if (p.lik_bound > 0.0f) L = max_nan(L, p.lik_bound);
        if (p.lik) st_stream1(p.lik + image * p.lik_bs + e, L);
        acc += FAST ? lg2_approx(L) : log2f(L);
      }
    }
    if (p.bits) {
      const int64_t G = gridDim.x;
      const int64_t first = image * p.tiles_per_image, last = first + p.tiles_per_image - 1;
      const int64_t c_lo = ((first + 1) * G + total - 1) / total - 1;
      const int64_t c_hi = ((last + 1) * G + total - 1) / total - 1;
      rate_commit(acc, image, static_cast<unsigned int>((c_hi - c_lo + 1) * (kThreads / 32)), p.B, p.workspace,
                  p.bits, p.bits_accumulate);
    }
  }
}

// ------------------------------------------------------------------ 128-bit forward
// Same outputs as stanh_gc_fwd_kernel for 16-byte aligned tensors with n % 4 == 0: four elements per
// thread (LDG.128 / STG.128), persistent CTAs over contiguous tile ranges with a two-tile register
// ping-pong (the launch structure of gc_fwd_kernel), table lookups through the cell grids of
// stanh_tables.cuh, the soft terms as (1 - E) / (1 + E), E = 2^(2 beta log2e (b_k - x)) — one MUFU.EX2
// and one MUFU.RCP per threshold of the window — and the likelihood two elements per packed f32x2 op.
struct StanhVecParams {
  StanhParams s;
  unsigned int tpi, q_tiles, r_tiles;
  float lik_floor;        // the bound, or -inf when disabled
  float sat_r, c2;        // 9.1 / beta and 2 * beta * log2(e)
  int rm;                 // quantize about the mean
  int sym_same;           // the level index of the quantizer input is the requested symbol
};

template <bool FAST>
__device__ __forceinline__ F2 stanh_mass2(F2 n1, F2 n2, F2 s) {
  n1 = max_nan2(min_nan2(n1, 1e30f), -1e30f);
  n2 = max_nan2(min_nan2(n2, 1e30f), -1e30f);
  const F2 sc = min_nan2(s, 1e30f);
  F2 a, b;
  div2_rn(n1, n2, sc, a, b);
  const F2 c = f2(-0.70710678118654752440f);
  const F2 xa = mul2(c, a), xb = mul2(c, b);
  if (FAST) {
    F2 ea = erfc_pos_fast(min_nan2(abs2(xa), 12.0f));
    F2 eb = erfc_pos_fast(min_nan2(abs2(xb), 12.0f));
    ea.x = (xa.x < 0.0f) ? 2.0f - ea.x : ea.x;
    ea.y = (xa.y < 0.0f) ? 2.0f - ea.y : ea.y;
    eb.x = (xb.x < 0.0f) ? 2.0f - eb.x : eb.x;
    eb.y = (xb.y < 0.0f) ? 2.0f - eb.y : eb.y;
    return fma2(f2(0.5f), ea, mul2(f2(-0.5f), eb));
  }
  return make_float2(0.5f * erfcf(xa.x) - 0.5f * erfcf(xb.x), 0.5f * erfcf(xa.y) - 0.5f * erfcf(xb.y));
}

// soft quantizer value; beta > 0
// COUNT: also #{k : x > b_k} (the hard level's index) — every threshold below the window is below x, none above it is,
// so the walk only has to count inside the window (compute_gap needs both values of every element)
template <bool FAST, int KMAX, bool COUNT = false>
__device__ __forceinline__ float stanh_soft_sm(float x, float beta, float sat_r, float c2, const StanhSm<KMAX>& T, int* below = nullptr) {
  const int lo = stanh_count_ge_b(x - sat_r, T);         // k <  lo: tanh == +1
  const float xr = x + sat_r;
  float acc = 0.0f;
  int k = lo, cnt = lo;
  for (;; ++k) {
    const float2 e = T.bw()[k];                            // (b_k, w_k / 2); NaN pad ends the loop
    if (!(e.x < xr)) break;
    if (COUNT) cnt += (x > e.x) ? 1 : 0;
    if (FAST) {
      const float E = ex2_approx(c2 * (e.x - x));        // exp(-2 beta (x - b_k))
      acc = fmaf(e.y, (1.0f - E) * rcp_approx(1.0f + E), acc);
    } else {
      const float t = beta * (x - e.x);
      const float sg = 1.0f / (1.0f + expf(-(2.0f * t)));
      acc = fmaf(e.y, 2.0f * sg - 1.0f, acc);
    }
  }                                                      // k >= hi (= k now): tanh == -1
  const float sat = 0.5f * ((T.cw()[lo] - T.cw0) - (T.cwK - T.cw()[k]));
  if (COUNT) *below = cnt;
  return (x != x) ? x : sat + acc;
}

// hard level of x given c = #{k : x > b_k}
template <int KMAX>
__device__ __forceinline__ float stanh_hard_at(float x, int c, const StanhSm<KMAX>& T, bool symmetric) {
  float v = T.cw()[c];
  if (symmetric) {
    int ce = c;
    if (x == T.bw()[c].x) { do ++ce; while (x == T.bw()[ce].x); }   // sign(0) = 0: mean of the two adjacent levels at a tie
    v = (x != x) ? 0.0f : 0.5f * (v + T.cw()[ce]);
  }
  return v;
}

// hard level of x; c_out = #{k : x > b_k}
template <int KMAX>
__device__ __forceinline__ float stanh_hard_sm(float x, const StanhSm<KMAX>& T, bool symmetric, int& c_out) {
  const int c = stanh_count_gt_b(x, T);
  c_out = c;
  float v = T.cw()[c];
  if (symmetric) {
    int ce = c;
    if (x == T.bw()[c].x) { do ++ce; while (x == T.bw()[ce].x); }   // sign(0) = 0: mean of the two adjacent levels at a tie
    v = (x != x) ? 0.0f : 0.5f * (v + T.cw()[ce]);
  }
  return v;
}

// numerators (n1 >= n2) of the level cell that contains v; `guess` < 0: look the cell up, else try it first
template <int KMAX>
__device__ __forceinline__ void stanh_cell_bounds(float v, int guess, const StanhSm<KMAX>& T, float& n1, float& n2) {
  int j;
  if (guess < 0) j = stanh_count_gt_avg(v, T);
  else {
    j = guess;                                              // the level's own cell unless mu pushed v across a mid-point
    const float a0 = T.avgp()[j], a1 = T.avgp()[j + 1];
    if (!(v > a0) || (v > a1)) j = stanh_count_gt_avg(v, T);
  }
  const bool inside = (v > -1000.0f) && (v <= 1000.0f);    // the reference's +-1000 sentinels (:506,:511)
  const float2 lu = T.lowup()[j];
  const float low = inside ? lu.x : 0.0f, up = inside ? lu.y : 0.0f;
  if (v >= 0.0f) { n1 = low - v; n2 = -up - v; }
  else { n1 = v + up; n2 = v - low; }                     // NaN v lands here and stays NaN
}

struct StanhIn { float4 y, m, s; };

template <int MODE, bool FAST, int KMAX>      // MODE 0: hard levels, 1: soft form (beta > 0), 2: likelihood of the given values
__global__ void __launch_bounds__(kThreads, 5) stanh_gc_vec_kernel(const StanhVecParams q) {
  const StanhParams& p = q.s;
  StanhSm<KMAX> T;
  stage_stanh_sm(p.b, p.w, p.cum_w, p.avg, p.dist, p.K, T);
  const bool need_lik = p.lik || p.bits;
  const bool use_mu = p.mu != nullptr;
  const bool rm = q.rm != 0;
  const bool symmetric = p.symmetric != 0;
  constexpr unsigned int kTileBytes = kThreads * 16;
  const int groups = static_cast<int>(p.n >> 2);
  const unsigned int cta = blockIdx.x;
  unsigned int t = cta * q.q_tiles + min(cta, q.r_tiles);
  const unsigned int t_end = t + q.q_tiles + (cta < q.r_tiles ? 1u : 0u);
  const unsigned int tb = threadIdx.x * 16;

  StanhIn a, b;
  a.y = a.m = make_float4(0.f, 0.f, 0.f, 0.f); a.s = make_float4(1.f, 1.f, 1.f, 1.f);
  b = a;
  while (t < t_end) {
    const int image = static_cast<int>(t / q.tpi);
    const int chunk0 = static_cast<int>(t - image * q.tpi);
    const unsigned int seg_end = (static_cast<unsigned int>(image) + 1u) * q.tpi;
    const int ntiles = static_cast<int>(min(seg_end, t_end) - t);
    t += ntiles;
    const int64_t seg = static_cast<int64_t>(chunk0) * (kThreads * 4);
    auto at_seg = [&](auto* base, int64_t bs) { return base + (static_cast<int64_t>(image) * bs + seg); };
    const float* y = at_seg(p.y, p.y_bs);
    const float* mu = use_mu ? at_seg(p.mu, p.mu_bs) : nullptr;
    const float* sg = (need_lik && p.sigma) ? at_seg(p.sigma, p.sigma_bs) : nullptr;
    float* o_yhat = p.yhat ? at_seg(p.yhat, p.yhat_bs) : nullptr;
    float* o_lik = p.lik ? at_seg(p.lik, p.lik_bs) : nullptr;
    int32_t* o_sym = p.sym ? at_seg(p.sym, p.sym_bs) : nullptr;
    float* o_ste = p.ste ? at_seg(p.ste, p.ste_bs) : nullptr;
    const int g = chunk0 * kThreads + threadIdx.x;

    auto load = [&](StanhIn& r, int k) {
      if (g + k * kThreads >= groups) return;
      const unsigned int bo = tb + static_cast<unsigned int>(k) * kTileBytes;
      auto at = [bo](const float* ptr) { return reinterpret_cast<const float*>(reinterpret_cast<const char*>(ptr) + bo); };
      r.y = ld_stream4(at(y));
      if (mu) r.m = ld_stream4(at(mu));
      if (sg) r.s = ld_stream4(at(sg));
    };
    float acc = 0.0f;
    auto compute = [&](const StanhIn& r, int k) {
      if (g + k * kThreads >= groups) return;
      const unsigned int bo = tb + static_cast<unsigned int>(k) * kTileBytes;
      auto at = [bo](auto* ptr) { return reinterpret_cast<decltype(ptr)>(reinterpret_cast<char*>(ptr) + bo); };
      const float yy[4] = {r.y.x, r.y.y, r.y.z, r.y.w};
      const float mm[4] = {r.m.x, r.m.y, r.m.z, r.m.w};                 // zeros without means
      // the mean the quantizer works about: mu, or 0 (y - 0 and q + 0 are exact)
      const float mq[4] = {rm ? mm[0] : 0.0f, rm ? mm[1] : 0.0f, rm ? mm[2] : 0.0f, rm ? mm[3] : 0.0f};
      float yh[4], n1[4], n2[4];
      int lv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float x = yy[i] - mq[i];
        int c = -1;
        lv[i] = 0;
        if (MODE == 2) yh[i] = yy[i];
        else {
          const float qv = (MODE == 1) ? stanh_soft_sm<FAST>(x, p.beta, q.sat_r, q.c2, T) : stanh_hard_sm(x, T, symmetric, c);
          yh[i] = qv + mq[i];
          if (o_sym) {
            if (MODE == 0 && q.sym_same) lv[i] = c;
            else lv[i] = stanh_count_gt_b(yy[i] - mm[i], T);
          }
        }
        if (need_lik) {
          const float v = yh[i] - mm[i];                               // :547-550 (mm = 0 without means)
          stanh_cell_bounds(v, (MODE == 0) ? c : -1, T, n1[i], n2[i]);
        }
      }
      if (o_yhat) st_stream4(at(o_yhat), make_float4(yh[0], yh[1], yh[2], yh[3]));
      if (o_ste)     // ste_round(y - mu) + mu: the value the slice loop carries on when the STanH is frozen (tcm_stanh.py:433-434)
        st_stream4(at(o_ste), make_float4(rintf(yy[0] - mm[0]) + mm[0], rintf(yy[1] - mm[1]) + mm[1], rintf(yy[2] - mm[2]) + mm[2],
                                          rintf(yy[3] - mm[3]) + mm[3]));
      if (o_sym) st_stream4(at(o_sym), make_int4(lv[0] + p.sym_offset, lv[1] + p.sym_offset, lv[2] + p.sym_offset, lv[3] + p.sym_offset));
      if (need_lik) {
        const F2 s0 = max_nan2(make_float2(r.s.x, r.s.y), p.scale_bound), s1 = max_nan2(make_float2(r.s.z, r.s.w), p.scale_bound);
        const F2 L0 = max_nan2(stanh_mass2<FAST>(make_float2(n1[0], n1[1]), make_float2(n2[0], n2[1]), s0), q.lik_floor);
        const F2 L1 = max_nan2(stanh_mass2<FAST>(make_float2(n1[2], n1[3]), make_float2(n2[2], n2[3]), s1), q.lik_floor);
        if (o_lik) st_stream4(at(o_lik), make_float4(L0.x, L0.y, L1.x, L1.y));
        if (FAST) {
          // one MUFU.LG2 per group while the product of four bounded likelihoods stays normal
          if (q.lik_floor >= 1e-9f) acc += lg2_approx((L0.x * L0.y) * (L1.x * L1.y));
          else acc += (lg2_approx(L0.x) + lg2_approx(L0.y)) + (lg2_approx(L1.x) + lg2_approx(L1.y));
        } else {
          acc += (log2f(L0.x) + log2f(L0.y)) + (log2f(L1.x) + log2f(L1.y));
        }
      }
    };

    load(a, 0);
    for (int k = 0; k < ntiles; k += 2) {
      if (k + 1 < ntiles) load(b, k + 1);
      compute(a, k);
      if (k + 2 < ntiles) load(a, k + 2);
      if (k + 1 < ntiles) compute(b, k + 1);
    }
    if (p.bits) {
      auto owner = [&](unsigned int tau) -> unsigned int {
        const unsigned int big = q.r_tiles * (q.q_tiles + 1u);
        return tau < big ? tau / (q.q_tiles + 1u) : q.r_tiles + (tau - big) / q.q_tiles;
      };
      const unsigned int first = static_cast<unsigned int>(image) * q.tpi;
      const unsigned int n_ctas = owner(first + q.tpi - 1u) - owner(first) + 1u;
      rate_commit(acc, image, n_ctas * (kThreads / 32), p.B, p.workspace, p.bits, p.bits_accumulate);
    }
  }
}

// Activation alone (module forward) and the two squared-error sums compute_gap needs, in one pass:
// out_soft / out_hard nullable; gap[0] += sum (x - soft)^2, gap[1] += sum (x - hard)^2.
template <bool FAST, int KMAX>
__global__ void __launch_bounds__(kThreads) stanh_act_kernel(const StanhParams p, float* out_soft, float* out_hard,
                                                             double* partials, unsigned int* counter, int vec,
                                                             float sat_r, float c2) {
  __shared__ double s_red[2][kThreads / 32];
  __shared__ bool s_last;
  StanhSm<KMAX> T;
  stage_stanh_sm(p.b, p.w, p.cum_w, p.avg, p.dist, p.K, T);
  const bool want_soft = out_soft || p.gap, want_hard = out_hard || p.gap;
  const bool soft_is_hard = p.beta == -1.0f;
  const bool symmetric = p.symmetric != 0;
  double se_soft = 0.0, se_hard = 0.0;
  // soft value of one element; a beta <= 0 other than -1 has no saturation window: all K thresholds
  auto soft_of = [&](float x) {
    int c;
    if (soft_is_hard) return stanh_hard_sm(x, T, symmetric, c);
    if (p.beta > 0.0f) return stanh_soft_sm<FAST>(x, p.beta, sat_r, c2, T);
    float acc = 0.0f;
    for (int k = 0; k < T.K; ++k) {
      const float2 e = T.bw()[k];
      const float sg = 1.0f / (1.0f + expf(-(2.0f * (p.beta * (x - e.x)))));
      acc = fmaf(e.y, 2.0f * sg - 1.0f, acc);
    }
    return (x != x) ? x : acc;
  };
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  const int64_t first = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x;
  const int64_t n_vec = vec ? (p.n >> 2) : 0;
  for (int64_t i = first; i < n_vec; i += stride) {
    const float4 x4 = ld_stream4(p.y + 4 * i);
    const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
    float qs[4], qh[4], ds = 0.0f, dh = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c;
      if (want_soft && want_hard && p.beta > 0.0f) {        // compute_gap: the soft walk also yields the hard level's index
        qs[j] = stanh_soft_sm<FAST, KMAX, true>(xs[j], p.beta, sat_r, c2, T, &c);
        qh[j] = stanh_hard_at(xs[j], c, T, symmetric);
        const float d0 = xs[j] - qs[j], d1 = xs[j] - qh[j];
        ds = fmaf(d0, d0, ds); dh = fmaf(d1, d1, dh);
        continue;
      }
      if (want_soft) { qs[j] = soft_of(xs[j]); const float d = xs[j] - qs[j]; ds = fmaf(d, d, ds); }
      if (want_hard) { qh[j] = stanh_hard_sm(xs[j], T, symmetric, c); const float d = xs[j] - qh[j]; dh = fmaf(d, d, dh); }
    }
    if (out_soft) st_stream4(out_soft + 4 * i, make_float4(qs[0], qs[1], qs[2], qs[3]));
    if (out_hard) st_stream4(out_hard + 4 * i, make_float4(qh[0], qh[1], qh[2], qh[3]));
    se_soft += static_cast<double>(ds); se_hard += static_cast<double>(dh);
  }
  for (int64_t i = 4 * n_vec + first; i < p.n; i += stride) {     // everything (scalar launch) or the n % 4 tail
    const float x = p.y[i];
    int c;
    if (want_soft) {
      const float qv = soft_of(x);
      if (out_soft) out_soft[i] = qv;
      const float d = x - qv;
      se_soft += static_cast<double>(d * d);
    }
    if (want_hard) {
      const float qv = stanh_hard_sm(x, T, symmetric, c);
      if (out_hard) out_hard[i] = qv;
      const float d = x - qv;
      se_hard += static_cast<double>(d * d);
    }
  }
  if (!p.gap) return;
  // deterministic: fixed tree inside the CTA, one partial pair per CTA, last CTA adds them in order
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    se_soft += __shfl_xor_sync(0xffffffffu, se_soft, o);
    se_hard += __shfl_xor_sync(0xffffffffu, se_hard, o);
  }
  if (lane == 0) { s_red[0][warp] = se_soft; s_red[1][warp] = se_hard; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < kThreads / 32; ++i) { a += s_red[0][i]; b += s_red[1][i]; }
    partials[2 * blockIdx.x] = a; partials[2 * blockIdx.x + 1] = b;
    __threadfence();
    s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    // the last CTA adds the per-CTA partials: thread i takes partials i, i + 256, ... in order, then the same
    // fixed tree — the loads of a thread are independent, so this costs one L2 round trip, not gridDim.x
    __threadfence();
    double a = 0.0, b = 0.0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += kThreads) { a += __ldcg(partials + 2 * i); b += __ldcg(partials + 2 * i + 1); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    __syncthreads();
    if (lane == 0) { s_red[0][warp] = a; s_red[1][warp] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
      a = 0.0; b = 0.0;
      for (int i = 0; i < kThreads / 32; ++i) { a += s_red[0][i]; b += s_red[1][i]; }
      p.gap[0] = a; p.gap[1] = b;
      *counter = 0u;
    }
  }
}

// ------------------------------------------------------------------ factorized bottleneck + STanH
// EntropyBottleneckStanh.forward (src/entropy_models/adaptive_entropy_bottleneck.py:679-708): z is
// quantized by the STanH activation WITHOUT medians (:113-177 with means=None), the likelihood is the
// sign-trick sigmoid difference of the cumulative-logit MLP over the level cell [x - low, x + up]
// (:551-603, :643-666).  Same tiling as eb_fwd_kernel: the parameters of the few channels a
// 256-element tile touches are staged per tile; the STanH tables are staged once per CTA.
struct EbStanhParams {
  const float* z; int64_t z_bs;
  const float* matrix[5]; const float* bias[5]; const float* factor[4]; const float* medians;   // medians unused (zeros)
  float* zhat; float* lik; int32_t* sym; int64_t zhat_bs, lik_bs, sym_bs;
  float* half_lo; float* half_up; int32_t* cell; int64_t half_lo_bs, half_up_bs, cell_bs;   // optional (for the backward)
  double* bits; unsigned long long* workspace; int bits_accumulate;
  StanhParams st;            // tables + beta/symmetric (tensor fields unused)
  int64_t B, ne; int hw, C, tile, bpi, training;
  float lik_bound;
};

__global__ void __launch_bounds__(kThreads) eb_stanh_fwd_kernel(const EbStanhParams p) {
  extern __shared__ float sm[];
  __shared__ float s_par[kEbMaxCh * kEbStride];
  StanhTables T;
  stage_tables(p.st, sm, T);
  const int image = blockIdx.x / p.bpi;
  const int chunk = blockIdx.x - image * p.bpi;
  const bool need_lik = p.lik || p.bits;
  float acc = 0.0f;
  const int64_t ntiles = (p.ne + p.tile - 1) / p.tile;
  for (int64_t t = chunk; t < ntiles; t += p.bpi) {
    const int64_t e0 = t * p.tile;
    const int64_t e1 = min(p.ne, e0 + static_cast<int64_t>(p.tile));
    const int c_lo = static_cast<int>(e0 / p.hw);
    const int nch = static_cast<int>((e1 - 1) / p.hw) - c_lo + 1;
    __syncthreads();
    if (need_lik)
      for (int i = threadIdx.x; i < nch * kEbStride; i += kThreads) {
        const int cl = i / kEbStride, j = i - cl * kEbStride;
        s_par[i] = (j == oMed) ? 0.0f : eb_staged_param(p, c_lo + cl, j);
      }
    __syncthreads();
    const int64_t e = e0 + threadIdx.x;
    if (threadIdx.x < p.tile && e < e1) {
      const float zv = ld_stream1(p.z + image * p.z_bs + e);
      int level = 0;
      float x;
      if (p.training && p.st.beta != -1.0f) {
        x = stanh_soft(zv, p.st.beta, T);
        if (p.sym) { int c2; stanh_hard(zv, T, p.st.symmetric, c2); level = c2; }
      } else {
        x = stanh_hard(zv, T, p.st.symmetric, level);
      }
      if (p.zhat) st_stream1(p.zhat + image * p.zhat_bs + e, x);
      if (p.sym) st_stream1(p.sym + image * p.sym_bs + e, level + p.st.sym_offset);
      if (need_lik || p.cell || p.half_lo || p.half_up) {
        const float* P = s_par + (static_cast<int>(e / p.hw) - c_lo) * kEbStride;
        const int j = count_gt(x, T.avg, T.K, T.steps);
        const bool inside = (x > -1000.0f) && (x <= 1000.0f);
        const float low = (inside && j > 0) ? T.dist[j - 1] : 0.0f;
        const float up = (inside && j < T.K) ? T.dist[j] : 0.0f;
        if (p.half_lo) p.half_lo[image * p.half_lo_bs + e] = low;
        if (p.half_up) p.half_up[image * p.half_up_bs + e] = up;
        if (p.cell) p.cell[image * p.cell_bs + e] = inside ? j : -1;
        if (need_lik) {
          const float L = eb_combine(logits_cumulative(P, x - low), logits_cumulative(P, x + up), p.lik_bound);
          if (p.lik) st_stream1(p.lik + image * p.lik_bs + e, L);
          acc += log2f(L);
        }
      }
    }
  }
  if (p.bits) rate_commit(acc, image, static_cast<unsigned int>(p.bpi * (kThreads / 32)), p.B, p.workspace, p.bits,
                          p.bits_accumulate);
}

static int steps_for(int K) { int s = 1; while ((1 << s) <= K) ++s; return s; }

static int check_tables(const reslic_stanh_tables* t, const char* who) {
  if (!t || !t->b || !t->w || !t->cum_w || !t->average_points || !t->distance_points) {
    set_error(RESLIC_ERR_ARG, "stanh: a table pointer is null");
    return RESLIC_ERR_ARG;
  }
  if (t->K < 1 || t->K > kStanhMaxK) {
    set_error(RESLIC_ERR_ARG, "stanh: K outside 1..1024");
    return RESLIC_ERR_ARG;
  }
  (void)who;
  return RESLIC_OK;
}

static void fill_tables(StanhParams& p, const reslic_stanh_tables* t) {
  p.b = t->b; p.w = t->w; p.cum_w = t->cum_w; p.avg = t->average_points; p.dist = t->distance_points;
  p.K = t->K; p.steps = steps_for(t->K); p.symmetric = t->symmetric; p.beta = t->beta;
  p.sym_offset = t->symmetric ? -(t->K / 2) : 0;
}

static size_t tables_smem(int K) { return static_cast<size_t>(5 * K + 1) * sizeof(float); }

int stanh_gc_fwd_launch(const reslic_stanh_gc_desc* d, cudaStream_t st) {
  if (!d) return set_error(RESLIC_ERR_ARG, "stanh_gc_fwd: null descriptor");
  if (d->struct_size != sizeof(reslic_stanh_gc_desc))
    return set_error(RESLIC_ERR_ARG, "stanh_gc_fwd: struct_size != sizeof(reslic_stanh_gc_desc) (binding built against another ABI revision)");
  if (d->B < 0 || d->n < 0) return set_error(RESLIC_ERR_ARG, "stanh_gc_fwd: negative size");
  if (d->B == 0 || d->n == 0) return RESLIC_OK;
  if (int rc = check_tables(&d->tables, "stanh_gc_fwd")) return rc;
  if (!d->y) return set_error(RESLIC_ERR_ARG, "stanh_gc_fwd: y is null");
  const bool need_lik = d->lik || rate_requested(d->bits, d->bits_accumulate);
  if (need_lik && !d->sigma) return set_error(RESLIC_ERR_ARG, "stanh_gc_fwd: sigma is null");
  if (!d->yhat && !d->sym && !d->ste && !need_lik) return set_error(RESLIC_ERR_ARG, "stanh_gc_fwd: no output requested");
  if (need_lik && !(d->scale_bound > 0.0f)) return set_error(RESLIC_ERR_ARG, "stanh_gc_fwd: scale_bound must be > 0");
  if (d->n >= (1LL << 31)) return set_error(RESLIC_ERR_ARG, "stanh_gc_fwd: more than 2^31 elements per image");
  StanhParams p{};
  p.y = d->y; p.mu = d->mu; p.sigma = d->sigma; p.y_bs = d->y_bs; p.mu_bs = d->mu_bs; p.sigma_bs = d->sigma_bs;
  p.yhat = d->yhat; p.lik = d->lik; p.sym = d->sym; p.yhat_bs = d->yhat_bs; p.lik_bs = d->lik_bs; p.sym_bs = d->sym_bs;
  p.ste = d->ste; p.ste_bs = d->ste_bs;
  fill_tables(p, &d->tables);
  p.removing_mean = d->removing_mean; p.training = d->training;
  p.scale_bound = d->scale_bound; p.lik_bound = d->likelihood_bound;
  p.n = d->n; p.B = d->B; p.tiles_per_image = (d->n + kThreads - 1) / kThreads;
  if (rate_requested(d->bits, d->bits_accumulate)) {
    const int rc = rate_setup("stanh_gc_fwd", d->bits, d->bits_accumulate, d->workspace, d->workspace_bytes, d->B,
                              &p.bits, &p.bits_accumulate, &p.workspace);
    if (rc != RESLIC_OK) return rc;
  }
  const bool fast = math_mode() != RESLIC_MATH_MIRROR;
  // 128-bit path: every tensor 16-byte aligned, batch strides and n multiples of 4, a soft form with beta > 0
  const bool soft = d->training == 1 && p.beta != -1.0f;
  bool vec = (d->n % 4 == 0) && !(soft && !(p.beta > 0.0f)) && (d->n / 4 + kThreads - 1) / kThreads * d->B < (1LL << 31);
  auto chk = [&](const void* ptr, int64_t bs) {
    if (ptr && ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0 || (bs % 4) != 0)) vec = false;
  };
  chk(d->y, d->y_bs); chk(d->mu, d->mu_bs); if (need_lik) chk(d->sigma, d->sigma_bs);
  chk(d->yhat, d->yhat_bs); chk(d->lik, d->lik_bs); chk(d->sym, d->sym_bs); chk(d->ste, d->ste_bs);
  cudaError_t err = cudaSuccess;
  if (vec) {
    StanhVecParams q{};
    q.s = p;
    const int64_t tpi = (d->n / 4 + kThreads - 1) / kThreads;
    const int64_t total = tpi * d->B;
    const int mode = d->training == 2 ? 2 : (soft ? 1 : 0);
    q.lik_floor = d->likelihood_bound > 0.0f ? d->likelihood_bound : -std::numeric_limits<float>::infinity();
    q.sat_r = soft ? kSatT / p.beta : 0.0f;
    q.c2 = soft ? 2.0f * p.beta * 1.44269504088896340736f : 0.0f;
    q.rm = d->mu && (d->training == 1 ? d->removing_mean != 0 : true);
    q.sym_same = q.rm || !d->mu;
    auto go = [&](auto kernel, int* occ, size_t smem) -> cudaError_t {
      if (*occ == 0) {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kThreads, smem) != cudaSuccess || nb < 1) nb = 1;
        *occ = nb;
      }
      // This kernel is issue-bound: an SM's CTAs share its issue slots, so the launch ends when the SM with the most
      // tiles is done.  With k CTAs on every SM that is k * ceil(tiles / (148 k)) tile times: take the k (>= 3 for
      // latency hiding, <= resident) that minimises it, the larger k on a tie — 1,024 tiles run as 592 CTAs of <= 2
      // tiles (8 per SM) instead of 740 of which 284 ran two (10 per SM); 4,096 as 592 x 7 (28) instead of 740 x 6 (30).
      int64_t grid = static_cast<int64_t>(*occ) * sm_count();
      if (grid > total) grid = total;
      else {
        int64_t best = -1;
        for (int k = *occ; k >= 3; --k) {
          const int64_t g = static_cast<int64_t>(k) * sm_count();
          const int64_t cost = k * ((total + g - 1) / g);
          if (best < 0 || cost < best) { best = cost; grid = g; }
        }
      }
      q.tpi = static_cast<unsigned int>(tpi);
      q.q_tiles = static_cast<unsigned int>(total / grid);
      q.r_tiles = static_cast<unsigned int>(total % grid);
      if (static_cast<int64_t>(q.q_tiles) + 1 >= (1LL << 32) / (kThreads * 16)) return cudaErrorInvalidValue;
      kernel<<<static_cast<int>(grid), kThreads, smem, st>>>(q);
      return cudaGetLastError();
    };
    static int occ[12] = {0};
    auto pick = [&](auto kmax) -> cudaError_t {
      constexpr int KM = decltype(kmax)::value;
      constexpr size_t smem = StanhSm<KM>::kBytes;
      int* o = occ + (KM == 256 ? 0 : 6);
      if (fast) {
        if (mode == 0) return go(stanh_gc_vec_kernel<0, true, KM>, o + 0, smem);
        if (mode == 1) return go(stanh_gc_vec_kernel<1, true, KM>, o + 1, smem);
        return go(stanh_gc_vec_kernel<2, true, KM>, o + 2, smem);
      }
      if (mode == 0) return go(stanh_gc_vec_kernel<0, false, KM>, o + 3, smem);
      if (mode == 1) return go(stanh_gc_vec_kernel<1, false, KM>, o + 4, smem);
      return go(stanh_gc_vec_kernel<2, false, KM>, o + 5, smem);
    };
    err = p.K <= 256 ? pick(std::integral_constant<int, 256>{}) : pick(std::integral_constant<int, 1024>{});
  } else {
    const int64_t total = p.tiles_per_image * p.B;
    int64_t grid = static_cast<int64_t>(sm_count()) * 8;
    if (grid > total) grid = total;
    const size_t smem = tables_smem(p.K);
    if (!fast) stanh_gc_fwd_kernel<false><<<static_cast<int>(grid), kThreads, smem, st>>>(p);
    else stanh_gc_fwd_kernel<true><<<static_cast<int>(grid), kThreads, smem, st>>>(p);
    err = cudaGetLastError();
  }
  if (err != cudaSuccess) return set_cuda_error(err, "stanh_gc_fwd launch");
  return RESLIC_OK;
}

int stanh_act_launch(const float* x, int64_t n, const reslic_stanh_tables* t, float* out_soft, float* out_hard,
                     double* gap, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  if (n < 0) return set_error(RESLIC_ERR_ARG, "stanh_act: negative size");
  if (int rc = check_tables(t, "stanh_act")) return rc;
  if (n == 0) {
    if (gap) cudaMemsetAsync(gap, 0, 2 * sizeof(double), st);
    return RESLIC_OK;
  }
  if (!x) return set_error(RESLIC_ERR_ARG, "stanh_act: x is null");
  if (!out_soft && !out_hard && !gap) return set_error(RESLIC_ERR_ARG, "stanh_act: no output requested");
  StanhParams p{};
  p.y = x; p.n = n; p.gap = gap;
  fill_tables(p, t);
  int64_t grid = ((n + 3) / 4 + kThreads - 1) / kThreads;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (grid > cap) grid = cap;
  double* partials = nullptr; unsigned int* counter = nullptr;
  if (gap) {
    const int64_t need = 16 + static_cast<int64_t>(grid) * 2 * sizeof(double);
    if (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 7u))
      return set_error(RESLIC_ERR_WORKSPACE, "stanh_act: workspace missing, misaligned or too small (need 16 + 16*8*SMs bytes)");
    counter = static_cast<unsigned int*>(workspace);
    partials = reinterpret_cast<double*>(static_cast<char*>(workspace) + 16);
  }
  auto al16 = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  const int vec = (n >= 4 && al16(x) && al16(out_soft) && al16(out_hard)) ? 1 : 0;
  const bool soft = p.beta > 0.0f;
  const float sat_r = soft ? kSatT / p.beta : 0.0f, c2 = soft ? 2.0f * p.beta * 1.44269504088896340736f : 0.0f;
  const bool mirror = math_mode() == RESLIC_MATH_MIRROR;
  // grid-stride over float4s: never more CTAs than are resident at once (1184 CTAs on 740 slots ran as two waves)
  static int resident[4] = {0, 0, 0, 0};             // per instantiation, filled on first use
  auto launch = [&](auto kernel, size_t smem, int which) {
    if (resident[which] == 0) {
      int nb = 0;
      resident[which] = (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kThreads, smem) == cudaSuccess && nb > 0) ? nb : 1;
    }
    const int64_t wave = static_cast<int64_t>(sm_count()) * resident[which];
    kernel<<<static_cast<int>(grid < wave ? grid : wave), kThreads, smem, st>>>(p, out_soft, out_hard, partials, counter, vec, sat_r, c2);
  };
  if (p.K <= 256) {
    if (mirror) launch(stanh_act_kernel<false, 256>, StanhSm<256>::kBytes, 0); else launch(stanh_act_kernel<true, 256>, StanhSm<256>::kBytes, 1);
  } else {
    if (mirror) launch(stanh_act_kernel<false, 1024>, StanhSm<1024>::kBytes, 2); else launch(stanh_act_kernel<true, 1024>, StanhSm<1024>::kBytes, 3);
  }
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_cuda_error(err, "stanh_act launch");
  return RESLIC_OK;
}

int eb_stanh_fwd_launch(const reslic_eb_stanh_desc* d, cudaStream_t st) {
  if (!d) return set_error(RESLIC_ERR_ARG, "eb_stanh_fwd: null descriptor");
  if (d->struct_size != sizeof(reslic_eb_stanh_desc))
    return set_error(RESLIC_ERR_ARG, "eb_stanh_fwd: struct_size != sizeof(reslic_eb_stanh_desc) (binding built against another ABI revision)");
  if (d->B < 0 || d->C < 0 || d->hw < 0) return set_error(RESLIC_ERR_ARG, "eb_stanh_fwd: negative size");
  if (d->B == 0 || d->C == 0 || d->hw == 0) return RESLIC_OK;
  if (int rc = check_tables(&d->tables, "eb_stanh_fwd")) return rc;
  if (!d->z) return set_error(RESLIC_ERR_ARG, "eb_stanh_fwd: z is null");
  for (int i = 0; i < 5; ++i)
    if (!d->matrix[i] || !d->bias[i] || (i < 4 && !d->factor[i]))
      return set_error(RESLIC_ERR_ARG, "eb_stanh_fwd: a parameter pointer is null");
  if (!d->zhat && !d->lik && !d->sym && !d->cell && !d->half_lo && !d->half_up && !rate_requested(d->bits, d->bits_accumulate))
    return set_error(RESLIC_ERR_ARG, "eb_stanh_fwd: no output requested");
  EbStanhParams p{};
  p.z = d->z; p.z_bs = d->z_bs;
  for (int i = 0; i < 5; ++i) { p.matrix[i] = d->matrix[i]; p.bias[i] = d->bias[i]; }
  for (int i = 0; i < 4; ++i) p.factor[i] = d->factor[i];
  p.medians = nullptr;
  p.zhat = d->zhat; p.lik = d->lik; p.sym = d->sym; p.zhat_bs = d->zhat_bs; p.lik_bs = d->lik_bs; p.sym_bs = d->sym_bs;
  p.half_lo = d->half_lo; p.half_up = d->half_up; p.cell = d->cell;
  p.half_lo_bs = d->half_lo_bs; p.half_up_bs = d->half_up_bs; p.cell_bs = d->cell_bs;
  fill_tables(p.st, &d->tables);
  p.B = d->B; p.ne = d->C * d->hw; p.hw = static_cast<int>(d->hw); p.C = static_cast<int>(d->C);
  p.training = d->training; p.lik_bound = d->likelihood_bound;
  int64_t tile = kThreads;
  if ((kEbMaxCh - 1) * d->hw < tile) tile = (kEbMaxCh - 1) * d->hw;
  p.tile = static_cast<int>(tile);
  int64_t bpi = (p.ne + tile - 1) / tile;
  const int64_t max_ctas = static_cast<int64_t>(sm_count()) * 64;
  if (bpi * d->B > max_ctas) bpi = (max_ctas + d->B - 1) / d->B;
  if (bpi < 1) bpi = 1;
  if (bpi * (kThreads / 32) > 60000) bpi = 60000 / (kThreads / 32);     // arrival count field is 16 bits
  p.bpi = static_cast<int>(bpi);
  if (rate_requested(d->bits, d->bits_accumulate)) {
    const int rc = rate_setup("eb_stanh_fwd", d->bits, d->bits_accumulate, d->workspace, d->workspace_bytes, d->B,
                              &p.bits, &p.bits_accumulate, &p.workspace);
    if (rc != RESLIC_OK) return rc;
  }
  eb_stanh_fwd_kernel<<<static_cast<int>(bpi * d->B), kThreads, tables_smem(p.st.K), st>>>(p);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_cuda_error(err, "eb_stanh_fwd launch");
  return RESLIC_OK;
}

// ------------------------------------------------------------------ backward (non-trainable STanH)
// What autograd does through GaussianConditionalStanh.forward for fixed w, b
// (gaussian_configuration["trainable"] = False, the reference default, src/utils/parser.py):
//   soft form:  d stanh / dx = sum_k (w_k/2) * beta * (1 - tanh^2(beta (x - b_k)))   (window only)
//   hard form:  zero (sign / relu)
//   likelihood: bins (low, up) are piecewise constant in v; dL/dv = (phi(a1) - phi(a2)) * dir / s,
//               dL/ds = -(a1 phi(a1) - a2 phi(a2)) / s, both LowerBound gradient rules.
struct StanhBwdParams {
  const float* y; const float* mu; const float* sigma; const float* g_yhat; const float* g_lik;
  float* g_y; float* g_mu; float* g_sigma;
  int64_t y_bs, mu_bs, sigma_bs, g_yhat_bs, g_lik_bs, g_y_bs, g_mu_bs, g_sigma_bs;
  StanhParams st;
  int64_t n, B, tiles_per_image;
  int training, removing_mean;
  float scale_bound, lik_bound;
  double* g_par;       // optional [5K+2]: A[K+1] | Bq[K+1] | Ww[K] | Wb[K] | Hd[K] (see reslic_stanh_gc_bwd_desc)
};

// Parameter-gradient accumulators of one CTA (doubles in shared memory, flushed with one global atomic per
// entry at the end).  For an element with upstream gradient G on the soft quantizer output q and
// saturation window [lo, hi):  dq/dw_k = +1/2 for k < lo, -1/2 for k >= hi — recorded as ONE add to A[lo]
// and ONE to Bq[hi] (the host turns the two histograms into per-k sums with a prefix sum) — and the exact
// derivative inside the window, where also dq/db_k lives.
struct StanhParAcc {
  double* A; double* Bq; double* Ww; double* Wb; double* Hd;
};

// Soft quantizer value, its derivative and (PAR) the parameter-gradient sums of one element in ONE walk over
// the window: tanh = (1 - E) / (1 + E), d tanh / dt = 1 - tanh^2.  beta <= 0 (other than -1) walks all K.
template <bool PAR, int KMAX>
__device__ __forceinline__ void stanh_soft_bwd(float x, float beta, float sat_r, float c2, const StanhSm<KMAX>& T, float G,
                                               const StanhParAcc& acc, float& q, float& dq) {
  int lo = 0;
  float xr = __int_as_float(0x7f800000);
  if (beta > 0.0f) { lo = stanh_count_ge_b(x - sat_r, T); xr = x + sat_r; }
  float aq = 0.0f, ad = 0.0f;
  int k = lo;
  for (; k < T.K; ++k) {
    const float2 e = T.bw()[k];
    if (!(e.x < xr)) break;
    const float E = ex2_approx(c2 * (e.x - x));
    const float th = (1.0f - E) * rcp_approx(1.0f + E);
    const float sech2 = fmaf(-th, th, 1.0f);
    aq = fmaf(e.y, th, aq);
    ad = fmaf(e.y * beta, sech2, ad);
    if (PAR) {
      atomicAdd(&acc.Ww[k], static_cast<double>(G * 0.5f * th));
      atomicAdd(&acc.Wb[k], static_cast<double>(-G * e.y * beta * sech2));
    }
  }
  if (PAR) {
    atomicAdd(&acc.A[lo], static_cast<double>(G));
    atomicAdd(&acc.Bq[k], static_cast<double>(G));
  }
  const float sat = 0.5f * ((T.cw()[lo] - T.cw0) - (T.cwK - T.cw()[k]));
  q = (x != x) ? x : sat + aq;
  dq = ad;
}

// One element of the backward.  Returns gy, gmu, gs.
template <bool PAR, int KMAX>
__device__ __forceinline__ void stanh_bwd_elem(const StanhBwdParams& p, const StanhSm<KMAX>& T, const StanhParAcc& acc,
                                               float sat_r, float c2, float y, float mu, float sg_in, float gyh, float gl,
                                               float& gy, float& gmu, float& gs_out) {
  const bool use_mu = p.mu != nullptr;
  const bool rm = use_mu && (p.training ? p.removing_mean != 0 : true);
  const float x = rm ? y - mu : y;
  const bool soft = p.training && p.st.beta != -1.0f;
  const bool symmetric = p.st.symmetric != 0;
  float qv, dq = 0.0f, gv = 0.0f, gs = 0.0f;
  int lvl = -1;
  // the soft walk needs the total upstream gradient on q for the parameter sums, which includes the
  // likelihood's share gv: without PAR one walk suffices, with PAR the value comes first and the sums later
  StanhParAcc none{};
  if (soft) stanh_soft_bwd<false>(x, p.st.beta, sat_r, c2, T, 0.0f, none, qv, dq);
  else qv = stanh_hard_sm(x, T, symmetric, lvl);
  const float yhat = rm ? qv + mu : qv;
  if (p.g_lik) {
    const float s = max_nan(sg_in, p.scale_bound);
    const float v = use_mu ? yhat - mu : yhat;
    int j;
    if (soft) j = stanh_count_gt_avg(v, T);
    else { j = lvl; if (!(v > T.avgp()[j]) || (v > T.avgp()[j + 1])) j = stanh_count_gt_avg(v, T); }
    const bool inside = (v > -1000.0f) && (v <= 1000.0f);
    const float2 lu = T.lowup()[j];
    const float low = inside ? lu.x : 0.0f, up = inside ? lu.y : 0.0f;
    float n1, n2, dir;
    if (v >= 0.0f) { n1 = low - v; n2 = -up - v; dir = -1.0f; }
    else { n1 = v + up; n2 = v - low; dir = 1.0f; }
    const float rs = rcp_approx(s);
    const float a1 = n1 * rs, a2 = n2 * rs;
    bool pass_l = true;
    if (p.lik_bound > 0.0f && !(gl < 0.0f)) {
      // L >= 1.6e-8 whenever the cell is at least 1e-3 sigma wide and starts within 3.5 sigma of the mean
      const bool near = (a1 * a2 <= 0.0f) || (fminf(fabsf(a1), fabsf(a2)) <= 3.5f);
      const bool surely_above = near && (a1 - a2 >= 1e-3f) && (p.lik_bound <= 1e-8f) && (s <= 1e30f);
      if (!surely_above) {
        const float L = gauss_interval_mass<true>(max_nan(min_nan(n1, 1e30f), -1e30f), max_nan(min_nan(n2, 1e30f), -1e30f), s);
        pass_l = L >= p.lik_bound;
      }
    }
    const float g = pass_l ? gl : 0.0f;
    const float kk = 0.3989422804014327f, ch = -0.72134752044448170368f;     // 1/sqrt(2 pi), -log2(e)/2
    const float p1 = kk * ex2_approx(ch * a1 * a1), p2 = kk * ex2_approx(ch * a2 * a2);
    gv = g * (p1 - p2) * dir * rs;
    gs = -g * (a1 * p1 - a2 * p2) * rs;
    if (PAR && g != 0.0f) {
      // L = Phi(a1) - Phi(a2): v >= 0: a1 = (low - v)/s, a2 = (-up - v)/s;  v < 0: a1 = (v + up)/s, a2 = (v - low)/s;
      // low = dist[j-1], up = dist[j] (the half-widths depend on the weights)
      const float dlow = (v >= 0.0f ? p1 : p2) * rs, dup = (v >= 0.0f ? p2 : p1) * rs;
      if (inside && j > 0) atomicAdd(&acc.Hd[j - 1], static_cast<double>(g * dlow));
      if (inside && j < T.K) atomicAdd(&acc.Hd[j], static_cast<double>(g * dup));
    }
    const bool pass_s = (sg_in >= p.scale_bound) || (gs < 0.0f);
    gs = pass_s ? gs : 0.0f;
  }
  // v = yhat - mu:  dv/dy = dq,  dv/dmu = (rm ? 1 - dq : 0) - (mu given ? 1 : 0)
  const float dyhat_dmu = rm ? 1.0f - dq : 0.0f;
  gy = (gyh + gv) * dq;
  gmu = gyh * dyhat_dmu + gv * (dyhat_dmu - (use_mu ? 1.0f : 0.0f));
  gs_out = gs;
  if (PAR && (gyh + gv) != 0.0f && x == x) {
    if (soft) {
      float q2, d2;
      stanh_soft_bwd<true>(x, p.st.beta, sat_r, c2, T, gyh + gv, acc, q2, d2);
    } else {
      // hard form: the level is still linear in the weights — dq/dw_k = +1/2 where x > b_k, -1/2 where
      // x < b_k; at an exact tie the non-symmetric form (relu(sign(0)) = 0) gives -1/2, the symmetric one 0
      int c_ge = lvl;
      if (symmetric) while (x == T.bw()[c_ge].x) ++c_ge;
      atomicAdd(&acc.A[lvl], static_cast<double>(gyh + gv));
      atomicAdd(&acc.Bq[c_ge], static_cast<double>(gyh + gv));
    }
  }
}

// W = 4: one 128-bit group per thread and tile (aligned tensors, n % 4 == 0); W = 1: any layout.
template <int W, bool PAR, int KMAX>
__global__ void __launch_bounds__(kThreads) stanh_gc_bwd_kernel(const StanhBwdParams p, float sat_r, float c2) {
  StanhSm<KMAX> T;
  stage_stanh_sm(p.st.b, p.st.w, p.st.cum_w, p.st.avg, p.st.dist, p.st.K, T);
  StanhParAcc acc{};
  const int n_acc = 5 * p.st.K + 2;
  if (PAR) {
    double* base = reinterpret_cast<double*>(stanh_smem + StanhSm<KMAX>::kBytes);
    for (int i = threadIdx.x; i < n_acc; i += blockDim.x) base[i] = 0.0;
    acc.A = base; acc.Bq = acc.A + p.st.K + 1; acc.Ww = acc.Bq + p.st.K + 1; acc.Wb = acc.Ww + p.st.K; acc.Hd = acc.Wb + p.st.K;
    __syncthreads();
  }
  const int64_t total = p.tiles_per_image * p.B;
  for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
    const int64_t image = t / p.tiles_per_image;
    const int64_t e = ((t - image * p.tiles_per_image) * kThreads + threadIdx.x) * W;
    if (e >= p.n) continue;
    if (W == 4) {
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 y = ld_stream4(p.y + image * p.y_bs + e);
      const float4 mu = p.mu ? ld_stream4(p.mu + image * p.mu_bs + e) : zero;
      const float4 sg = p.g_lik ? ld_stream4(p.sigma + image * p.sigma_bs + e) : zero;
      const float4 gyh = p.g_yhat ? ld_stream4(p.g_yhat + image * p.g_yhat_bs + e) : zero;
      const float4 gl = p.g_lik ? ld_stream4(p.g_lik + image * p.g_lik_bs + e) : zero;
      float4 gy, gmu, gs;
      stanh_bwd_elem<PAR>(p, T, acc, sat_r, c2, y.x, mu.x, sg.x, gyh.x, gl.x, gy.x, gmu.x, gs.x);
      stanh_bwd_elem<PAR>(p, T, acc, sat_r, c2, y.y, mu.y, sg.y, gyh.y, gl.y, gy.y, gmu.y, gs.y);
      stanh_bwd_elem<PAR>(p, T, acc, sat_r, c2, y.z, mu.z, sg.z, gyh.z, gl.z, gy.z, gmu.z, gs.z);
      stanh_bwd_elem<PAR>(p, T, acc, sat_r, c2, y.w, mu.w, sg.w, gyh.w, gl.w, gy.w, gmu.w, gs.w);
      if (p.g_y) st_stream4(p.g_y + image * p.g_y_bs + e, gy);
      if (p.g_mu) st_stream4(p.g_mu + image * p.g_mu_bs + e, gmu);
      if (p.g_sigma) st_stream4(p.g_sigma + image * p.g_sigma_bs + e, gs);
    } else {
      const float y = p.y[image * p.y_bs + e];
      const float mu = p.mu ? p.mu[image * p.mu_bs + e] : 0.0f;
      const float sg = p.g_lik ? p.sigma[image * p.sigma_bs + e] : 0.0f;
      const float gyh = p.g_yhat ? p.g_yhat[image * p.g_yhat_bs + e] : 0.0f;
      const float gl = p.g_lik ? p.g_lik[image * p.g_lik_bs + e] : 0.0f;
      float gy, gmu, gs;
      stanh_bwd_elem<PAR>(p, T, acc, sat_r, c2, y, mu, sg, gyh, gl, gy, gmu, gs);
      if (p.g_y) p.g_y[image * p.g_y_bs + e] = gy;
      if (p.g_mu) p.g_mu[image * p.g_mu_bs + e] = gmu;
      if (p.g_sigma) p.g_sigma[image * p.g_sigma_bs + e] = gs;
    }
  }
  if (PAR) {
    __syncthreads();
    for (int i = threadIdx.x; i < n_acc; i += blockDim.x) {
      const double v = acc.A[i];
      if (v != 0.0) atomicAdd(&p.g_par[i], v);
    }
  }
}

int stanh_gc_bwd_launch(const reslic_stanh_gc_bwd_desc* d, cudaStream_t st) {
  if (!d) return set_error(RESLIC_ERR_ARG, "stanh_gc_bwd: null descriptor");
  if (d->struct_size != sizeof(reslic_stanh_gc_bwd_desc))
    return set_error(RESLIC_ERR_ARG, "stanh_gc_bwd: struct_size != sizeof(reslic_stanh_gc_bwd_desc) (binding built against another ABI revision)");
  if (d->B < 0 || d->n < 0) return set_error(RESLIC_ERR_ARG, "stanh_gc_bwd: negative size");
  if (d->B == 0 || d->n == 0) return RESLIC_OK;
  if (int rc = check_tables(&d->tables, "stanh_gc_bwd")) return rc;
  if (!d->y) return set_error(RESLIC_ERR_ARG, "stanh_gc_bwd: y is null");
  if (d->g_lik && !d->sigma) return set_error(RESLIC_ERR_ARG, "stanh_gc_bwd: sigma is null");
  if (!d->g_y && !d->g_mu && !d->g_sigma && !d->g_params)
    return set_error(RESLIC_ERR_ARG, "stanh_gc_bwd: no output requested");
  if (d->g_params && d->g_params_len < 5 * static_cast<int64_t>(d->tables.K) + 2)
    return set_error(RESLIC_ERR_ARG, "stanh_gc_bwd: g_params needs 5*K+2 doubles");
  if (d->g_params && (reinterpret_cast<uintptr_t>(d->g_params) & 7u))
    return set_error(RESLIC_ERR_ARG, "stanh_gc_bwd: g_params must be 8-byte aligned");
  if (!(d->scale_bound > 0.0f)) return set_error(RESLIC_ERR_ARG, "stanh_gc_bwd: scale_bound must be > 0");
  StanhBwdParams p{};
  p.y = d->y; p.mu = d->mu; p.sigma = d->sigma; p.g_yhat = d->g_yhat; p.g_lik = d->g_lik;
  p.g_y = d->g_y; p.g_mu = d->g_mu; p.g_sigma = d->g_sigma;
  p.y_bs = d->y_bs; p.mu_bs = d->mu_bs; p.sigma_bs = d->sigma_bs; p.g_yhat_bs = d->g_yhat_bs; p.g_lik_bs = d->g_lik_bs;
  p.g_y_bs = d->g_y_bs; p.g_mu_bs = d->g_mu_bs; p.g_sigma_bs = d->g_sigma_bs;
  fill_tables(p.st, &d->tables);
  p.n = d->n; p.B = d->B;
  p.training = d->training; p.removing_mean = d->removing_mean;
  p.scale_bound = d->scale_bound; p.lik_bound = d->likelihood_bound;
  bool vec = (d->n % 4 == 0);
  auto chk = [&](const void* ptr, int64_t bs) {
    if (ptr && ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0 || (bs % 4) != 0)) vec = false;
  };
  chk(d->y, d->y_bs); chk(d->mu, d->mu_bs); if (d->g_lik) chk(d->sigma, d->sigma_bs);
  chk(d->g_yhat, d->g_yhat_bs); chk(d->g_lik, d->g_lik_bs);
  chk(d->g_y, d->g_y_bs); chk(d->g_mu, d->g_mu_bs); chk(d->g_sigma, d->g_sigma_bs);
  const int64_t per_tile = static_cast<int64_t>(kThreads) * (vec ? 4 : 1);
  p.tiles_per_image = (d->n + per_tile - 1) / per_tile;
  int64_t grid = p.tiles_per_image * p.B;
  p.g_par = d->g_params;
  // without parameter sums there is no per-CTA state beyond the tables: many short CTAs (the shape that
  // reaches the copy bandwidth); with them, few persistent CTAs so that the shared accumulators are flushed rarely
  const int64_t cap = static_cast<int64_t>(sm_count()) * (p.g_par ? 16 : 64);
  if (grid > cap) grid = cap;
  const bool soft = p.st.beta > 0.0f;
  const float sat_r = soft ? kSatT / p.st.beta : 0.0f, c2 = 2.0f * p.st.beta * 1.44269504088896340736f;
  if (p.g_par) {
    cudaError_t e = cudaMemsetAsync(p.g_par, 0, static_cast<size_t>(5 * p.st.K + 2) * sizeof(double), st);
    if (e != cudaSuccess) return set_cuda_error(e, "stanh_gc_bwd setup");
  }
  auto launch = [&](auto kernel, size_t smem) -> cudaError_t {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return e;
    }
    kernel<<<static_cast<int>(grid), kThreads, smem, st>>>(p, sat_r, c2);
    return cudaSuccess;
  };
  auto pick = [&](auto kmax) -> cudaError_t {
    constexpr int KM = decltype(kmax)::value;
    const size_t par = static_cast<size_t>(5 * p.st.K + 2) * sizeof(double);
    if (vec) return p.g_par ? launch(stanh_gc_bwd_kernel<4, true, KM>, StanhSm<KM>::kBytes + par)
                            : launch(stanh_gc_bwd_kernel<4, false, KM>, StanhSm<KM>::kBytes);
    return p.g_par ? launch(stanh_gc_bwd_kernel<1, true, KM>, StanhSm<KM>::kBytes + par)
                   : launch(stanh_gc_bwd_kernel<1, false, KM>, StanhSm<KM>::kBytes);
  };
  {
    const cudaError_t e = p.st.K <= 256 ? pick(std::integral_constant<int, 256>{}) : pick(std::integral_constant<int, 1024>{});
    if (e != cudaSuccess) return set_cuda_error(e, "stanh_gc_bwd setup");
  }
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_cuda_error(err, "stanh_gc_bwd launch");
  return RESLIC_OK;
}

}  // namespace reslic
