// gc_fused.cu — fused Gaussian-conditional forward for sm_100a.
//
// One pass over a latent slice produces any subset of
//   {quantize() output, ste_round output, bounded likelihood, int32 symbols,
//    int32 scale-table indexes, per-image -sum log2 L}
// from y, mu, sigma (+ optional noise).  Replaces the ~19+5+192+3 stock-PyTorch launches of
// compressai GaussianConditional.forward / ste_round / build_indexes / log-sum
// (reference call sites src/models/reference/tcm.py:455,457,544,548; likelihood twin
// tcm.py:570-588; build_indexes src/entropy_models/adaptive_gaussian_conditional.py:606-617;
// rate src/training/loss.py:24-27).
//
// The path is elementwise and HBM-bound (12 B read + 8..16 B written per element), so the
// design rules are: 128-bit coalesced streaming loads/stores, scale table staged once per
// CTA in shared memory, per-image rate reduced by warp shuffles + one fp64 partial per CTA,
// and an instruction budget small enough (~100 issue slots/element) that FP32/MUFU issue
// does not become the bound.  No tensor cores (nothing GEMM-shaped here).
#include "common.cuh"
#include "reslic_internal.h"

namespace reslic {

struct GcParams {
  const float* y; const float* mu; const float* sigma; const float* noise;
  int64_t y_bs, mu_bs, sigma_bs, noise_bs;
  float* yhat; float* ste; float* lik; int32_t* sym; int32_t* idx;
  int64_t yhat_bs, ste_bs, lik_bs, sym_bs, idx_bs;
  double* bits; double* workspace;
  const float* table; int table_len;
  int64_t n;          // elements per image
  int bpi;            // CTAs per image
  float scale_bound, lik_bound;
  uint32_t seed_lo, seed_hi, off_lo, off_hi;
};

// ------------------------------------------------------------------ per-element math
// Mirrors the reference op order: every line is one IEEE fp32 op there.
//   a = (0.5 - v)/s, b = (-0.5 - v)/s   (true divides)
//   L = 0.5*erfc(-2^-0.5 * a) - 0.5*erfc(-2^-0.5 * b)
// The two divides share one reciprocal and use the Markstein residual correction, which
// returns the correctly rounded quotient for the operand range left after the clamps
// (checked against __fdiv_rn in tests/test_gc_parity.py).
__device__ __forceinline__ void div2_rn(float n1, float n2, float s, float& q1, float& q2) {
  float r = rcp_approx(s);
  const float e = fmaf(-s, r, 1.0f);
  r = fmaf(r, e, r);
  float q = n1 * r;
  float rem = fmaf(-s, q, n1);
  q1 = fmaf(rem, r, q);
  q = n2 * r;
  rem = fmaf(-s, q, n2);
  q2 = fmaf(rem, r, q);
}

// erfc(x) for 0 <= x <= 12 as exp(-x^2) * (1-u) * Q(u), u = x/(x+2.5): one MUFU.RCP, one
// MUFU.EX2, 21 FP32 ops (CUDA's erfcf: 3 MUFU + FRND + ~44).  Q is a degree-9 near-minimax
// fit (|rel err| < 1.4e-8 in exact arithmetic); u = x*r keeps small x free of cancellation
// and (1-u) carries the 1/x decay so Horner stays well conditioned.  exp(-x^2) gets the
// rounding errors of x*x and of the log2(e) product back as a first-order correction, so
// the relative error stays ~3e-7 out to the likelihood floor (x^2 ~ 20).
__device__ __forceinline__ float erfc_pos_fast(float x) {
  const float r = rcp_approx(x + 2.5f);
  const float u = x * r;
  const float w = fmaf(-x, r, 1.0f);
  float q = 2.651532926e-02f;
  q = fmaf(q, u, -5.741734803e-02f);
  q = fmaf(q, u, -3.597635776e-02f);
  q = fmaf(q, u, 1.268966794e-01f);
  q = fmaf(q, u, 1.091585010e-01f);
  q = fmaf(q, u, -2.630832791e-01f);
  q = fmaf(q, u, -4.676126838e-01f);
  q = fmaf(q, u, 1.608165503e+00f);
  q = fmaf(q, u, -1.820949554e+00f);
  q = fmaf(q, u, 1.0f);
  const float L2E = 1.44269502162933349609375f;       // fp32(log2 e)
  const float L2E_LO = 1.925963033500011e-08f;        // log2 e - fp32(log2 e)
  const float s2 = x * x;
  const float e = fmaf(x, x, -s2);                    // exact low part of x*x
  const float t = s2 * L2E;
  float tl = fmaf(s2, L2E, -t);                       // exact low part of s2*L2E
  tl = fmaf(e, L2E, tl);
  tl = fmaf(s2, L2E_LO, tl);
  const float E0 = ex2_approx(-t);
  const float E = fmaf(E0 * tl, -0.693147182464599609375f, E0);   // 2^-(t+tl) ~ E0*(1 - ln2*tl)
  return E * (w * q);
}

template <bool FAST>
__device__ __forceinline__ float gc_likelihood(float v, float s) {
  // clamps keep every intermediate finite (inf/inf, 0*inf); they change no result for
  // |y-mu|, sigma <= 1e30 and give the reference's limit values (L -> 0 -> bound) beyond.
  const float vc = min_nan(v, 1e30f);
  const float sc = min_nan(s, 1e30f);
  float a, b;
  div2_rn(0.5f - vc, -0.5f - vc, sc, a, b);
  const float c = -0.70710678118654752440f;  // float(-(2 ** -0.5)) cast to fp32
  const float xa = c * a, xb = c * b;        // xb > 0 always; xa < 0 iff v < 0.5
  if (FAST) {
    float ea = erfc_pos_fast(min_nan(fabsf(xa), 12.0f));
    const float eb = erfc_pos_fast(min_nan(xb, 12.0f));
    ea = (xa < 0.0f) ? 2.0f - ea : ea;
    return fmaf(0.5f, ea, -0.5f * eb);
  }
  const float upper = 0.5f * erfcf(xa);
  const float lower = 0.5f * erfcf(xb);
  return upper - lower;
}

// build_indexes: idx = #{ j < len-1 : !(s <= table[j]) } = (len-1) - sum_j [s <= table[j]].
// Shared-memory layout `pad`: pad[0] = -inf, pad[1+j] = table[j] (j < len-1), +inf beyond, so
// idx = g  <=>  pad[g] < s <= pad[g+1].  The scale table is log-spaced
// (exp(linspace(ln .11, ln 256, 64)), tcm.py:26-34), so g is GUESSED from one MUFU.LG2 and
// a fused multiply-add, then PROVEN with the two exact comparisons the reference makes; a
// failed proof (s within rounding of a table value, a table that is not log-spaced, NaN)
// falls back to the branch-free binary search.  The result is bit-exact for any table.
constexpr int kPadLen = 260;

template <int STEPS>
__device__ __forceinline__ int scale_index_search(float s, const float* pad) {
  int lo = 0;
#pragma unroll
  for (int step = 1 << (STEPS - 1); step > 0; step >>= 1) {
    const float t = pad[lo + step];            // == table[lo + step - 1]
    lo += (!(s <= t)) ? step : 0;
  }
  return lo;  // NaN walks to 2^STEPS - 1; the caller clamps to table_len - 1
}

template <int STEPS>
__device__ __forceinline__ int scale_index(float s, const float* pad, float g_scale, float g_off, int last) {
  int g = __float2int_ru(fmaf(lg2_approx(s), g_scale, g_off));
  g = max(0, min(g, last));
  const float lo = pad[g], hi = pad[g + 1];
  if (!((lo < s) && (s <= hi))) g = min(scale_index_search<STEPS>(s, pad), last);
  return g;
}

template <bool NEED_LIK, bool NEED_IDX, bool NOISE, int STEPS, bool FAST>
struct GcElem {
  __device__ __forceinline__ static void run(const GcParams& p, const float* pad, float g_scale, float g_off,
                                             float y, float mu, float sg, float u, float& yhat,
                                             float& ste, float& lik, int& sym, int& idx, float& acc) {
    const float d = y - mu;
    const float q = rintf(d);            // torch.round: half to even
    sym = __float2int_rn(q);
    ste = q + mu;
    yhat = NOISE ? (y + u) : ste;
    const float s = max_nan(sg, p.scale_bound);
    if (NEED_LIK) {
      const float v = fabsf(yhat - mu);  // reference re-subtracts mu from the quantized value
      float L = gc_likelihood<FAST>(v, s);
      if (p.lik_bound > 0.0f) L = max_nan(L, p.lik_bound);
      lik = L;
      // FAST: MUFU.LG2 (abs err <= 2^-22 for L in (0.5,2), rel err 2^-22 elsewhere)
      acc += FAST ? lg2_approx(L) : log2f(L);
    }
    if (NEED_IDX) idx = scale_index<STEPS>(s, pad, g_scale, g_off, p.table_len - 1);
  }
};

template <bool NEED_LIK, bool NEED_IDX, bool NOISE, bool VEC, int STEPS, bool FAST>
__global__ void __launch_bounds__(kThreads)
gc_fwd_kernel(const GcParams p) {
  __shared__ float pad[NEED_IDX ? kPadLen : 1];
  __shared__ float s_guess[2];
  float g_scale = 0.0f, g_off = 0.0f;
  if (NEED_IDX) {
    const float inf = __int_as_float(0x7f800000);
    for (int i = threadIdx.x; i < kPadLen; i += kThreads)
      pad[i] = (i == 0) ? -inf : ((i <= p.table_len - 1) ? p.table[i - 1] : inf);
    if (threadIdx.x == 0) {
      // guess(s) = ceil((log2 s - log2 t_0) * (len-2) / (log2 t_{len-2} - log2 t_0))
      float sc = 0.0f, off = 0.0f;
      if (p.table_len >= 3) {
        const float l0 = log2f(p.table[0]), l1 = log2f(p.table[p.table_len - 2]);
        // the small downward bias makes sigma == table[j] (notably the 0.11 bound itself) guess j
        if (l1 > l0) { sc = static_cast<float>(p.table_len - 2) / (l1 - l0); off = -l0 * sc - 2.44140625e-4f; }
      }
      s_guess[0] = sc; s_guess[1] = off;
    }
    __syncthreads();
    g_scale = s_guess[0]; g_off = s_guess[1];
  }
  const int image = blockIdx.x / p.bpi;
  const int chunk = blockIdx.x - image * p.bpi;
  const float* __restrict__ y = p.y ? p.y + image * p.y_bs : nullptr;
  const float* __restrict__ mu = p.mu ? p.mu + image * p.mu_bs : nullptr;
  const float* __restrict__ sg = p.sigma ? p.sigma + image * p.sigma_bs : nullptr;
  const float* __restrict__ nz = (NOISE && p.noise) ? p.noise + image * p.noise_bs : nullptr;
  float* yhat = p.yhat ? p.yhat + image * p.yhat_bs : nullptr;
  float* ste = p.ste ? p.ste + image * p.ste_bs : nullptr;
  float* lik = p.lik ? p.lik + image * p.lik_bs : nullptr;
  int32_t* sym = p.sym ? p.sym + image * p.sym_bs : nullptr;
  int32_t* idx = p.idx ? p.idx + image * p.idx_bs : nullptr;

  float acc = 0.0f;
  constexpr int W = VEC ? 4 : 1;
  const int64_t groups = VEC ? (p.n >> 2) : p.n;
  const int64_t stride = static_cast<int64_t>(p.bpi) * kThreads;
  for (int64_t g = static_cast<int64_t>(chunk) * kThreads + threadIdx.x; g < groups; g += stride) {
    const int64_t e = g * W;
    float yv[4] = {0.f, 0.f, 0.f, 0.f}, mv[4] = {0.f, 0.f, 0.f, 0.f}, sv[4] = {1.f, 1.f, 1.f, 1.f}, uv[4] = {0.f, 0.f, 0.f, 0.f};
    if (VEC) {
      if (y) { const float4 t = ld_stream4(y + e); yv[0] = t.x; yv[1] = t.y; yv[2] = t.z; yv[3] = t.w; }
      if (mu) { const float4 m = ld_stream4(mu + e); mv[0] = m.x; mv[1] = m.y; mv[2] = m.z; mv[3] = m.w; }
      if (sg) { const float4 s4 = ld_stream4(sg + e); sv[0] = s4.x; sv[1] = s4.y; sv[2] = s4.z; sv[3] = s4.w; }
      if (NOISE && nz) { const float4 n4 = ld_stream4(nz + e); uv[0] = n4.x; uv[1] = n4.y; uv[2] = n4.z; uv[3] = n4.w; }
    } else {
      if (y) yv[0] = ld_stream1(y + e);
      if (mu) mv[0] = ld_stream1(mu + e);
      if (sg) sv[0] = ld_stream1(sg + e);
      if (NOISE && nz) uv[0] = ld_stream1(nz + e);
    }
    if (NOISE && !nz) {
      // counter = global element-group id; one Philox call feeds the W elements of a group
      const uint64_t gid = static_cast<uint64_t>(image) * static_cast<uint64_t>((p.n + 3) >> 2) +
                           static_cast<uint64_t>(VEC ? g : (g >> 2));
      const Philox4 r = philox4x32_10(static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32),
                                      p.off_lo, p.off_hi, p.seed_lo, p.seed_hi);
      if (VEC) {
        uv[0] = u32_to_centered_uniform(r.x); uv[1] = u32_to_centered_uniform(r.y);
        uv[2] = u32_to_centered_uniform(r.z); uv[3] = u32_to_centered_uniform(r.w);
      } else {
        const int k = static_cast<int>(g & 3);
        uv[0] = u32_to_centered_uniform(k == 0 ? r.x : k == 1 ? r.y : k == 2 ? r.z : r.w);
      }
    }
    float oy[4], os[4], ol[4]; int osym[4], oidx[4];
#pragma unroll
    for (int k = 0; k < W; ++k)
      GcElem<NEED_LIK, NEED_IDX, NOISE, STEPS, FAST>::run(p, pad, g_scale, g_off, yv[k], mv[k], sv[k], uv[k], oy[k], os[k],
                                                    ol[k], osym[k], oidx[k], acc);
    if (VEC) {
      if (yhat) st_stream4(yhat + e, make_float4(oy[0], oy[1], oy[2], oy[3]));
      if (ste) st_stream4(ste + e, make_float4(os[0], os[1], os[2], os[3]));
      if (NEED_LIK && lik) st_stream4(lik + e, make_float4(ol[0], ol[1], ol[2], ol[3]));
      if (sym) st_stream4(sym + e, make_int4(osym[0], osym[1], osym[2], osym[3]));
      if (NEED_IDX && idx) st_stream4(idx + e, make_int4(oidx[0], oidx[1], oidx[2], oidx[3]));
    } else {
      if (yhat) st_stream1(yhat + e, oy[0]);
      if (ste) st_stream1(ste + e, os[0]);
      if (NEED_LIK && lik) st_stream1(lik + e, ol[0]);
      if (sym) st_stream1(sym + e, osym[0]);
      if (NEED_IDX && idx) st_stream1(idx + e, oidx[0]);
    }
  }
  if (NEED_LIK && p.bits)
    image_sum_finish(acc, image, chunk, p.bpi, p.workspace, p.bits);
}

// ------------------------------------------------------------------ host-side launch
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <bool NEED_LIK, bool NEED_IDX, bool NOISE, bool VEC>
static cudaError_t launch_steps(const GcParams& p, int grid, int steps, bool fast, cudaStream_t st) {
  // the math policy only matters when a likelihood is computed
  if (NEED_LIK && !fast) {
    if (steps <= 6) gc_fwd_kernel<NEED_LIK, NEED_IDX, NOISE, VEC, 6, false><<<grid, kThreads, 0, st>>>(p);
    else gc_fwd_kernel<NEED_LIK, NEED_IDX, NOISE, VEC, 8, false><<<grid, kThreads, 0, st>>>(p);
  } else {
    if (steps <= 6) gc_fwd_kernel<NEED_LIK, NEED_IDX, NOISE, VEC, 6, true><<<grid, kThreads, 0, st>>>(p);
    else gc_fwd_kernel<NEED_LIK, NEED_IDX, NOISE, VEC, 8, true><<<grid, kThreads, 0, st>>>(p);
  }
  return cudaGetLastError();
}
template <bool NEED_LIK, bool NEED_IDX, bool NOISE>
static cudaError_t launch_vec(const GcParams& p, int grid, int steps, bool vec, bool fast, cudaStream_t st) {
  return vec ? launch_steps<NEED_LIK, NEED_IDX, NOISE, true>(p, grid, steps, fast, st)
             : launch_steps<NEED_LIK, NEED_IDX, NOISE, false>(p, grid, steps, fast, st);
}
template <bool NEED_LIK, bool NEED_IDX>
static cudaError_t launch_noise(const GcParams& p, int grid, int steps, bool vec, bool noise, bool fast,
                                cudaStream_t st) {
  return noise ? launch_vec<NEED_LIK, NEED_IDX, true>(p, grid, steps, vec, fast, st)
               : launch_vec<NEED_LIK, NEED_IDX, false>(p, grid, steps, vec, fast, st);
}

int gc_fwd_launch(const reslic_gc_desc* d, cudaStream_t st) {
  if (!d) return set_error(RESLIC_ERR_ARG, "gc_fwd: null descriptor");
  if (d->B < 0 || d->n < 0) return set_error(RESLIC_ERR_ARG, "gc_fwd: negative size");
  if (d->B == 0 || d->n == 0) return RESLIC_OK;  // empty input: nothing to do
  if (d->B > (1 << 20)) return set_error(RESLIC_ERR_ARG, "gc_fwd: B too large");
  if (d->mode != RESLIC_Q_DEQUANTIZE && d->mode != RESLIC_Q_NOISE)
    return set_error(RESLIC_ERR_ARG, "gc_fwd: invalid quantization mode");
  const bool need_lik = d->lik || d->bits;
  const bool need_idx = d->idx != nullptr;
  const bool need_y = d->yhat || d->ste || d->sym || need_lik;
  if (need_y && !d->y) return set_error(RESLIC_ERR_ARG, "gc_fwd: y is null");
  if ((need_lik || need_idx) && !d->sigma) return set_error(RESLIC_ERR_ARG, "gc_fwd: sigma is null");
  if (!need_y && !need_idx) return set_error(RESLIC_ERR_ARG, "gc_fwd: no output requested");
  if (need_idx && (!d->scale_table || d->table_len < 1 || d->table_len > 256))
    return set_error(RESLIC_ERR_ARG, "gc_fwd: scale_table missing or table_len outside 1..256");
  if (!(d->scale_bound > 0.0f)) return set_error(RESLIC_ERR_ARG, "gc_fwd: scale_bound must be > 0");

  GcParams p{};
  p.y = d->y; p.y_bs = d->y_bs;
  p.mu = d->mu; p.mu_bs = d->mu_bs; p.sigma = d->sigma; p.sigma_bs = d->sigma_bs;
  p.noise = d->noise; p.noise_bs = d->noise_bs;
  p.yhat = d->yhat; p.ste = d->ste; p.lik = d->lik; p.sym = d->sym; p.idx = d->idx;
  p.yhat_bs = d->yhat_bs; p.ste_bs = d->ste_bs; p.lik_bs = d->lik_bs; p.sym_bs = d->sym_bs; p.idx_bs = d->idx_bs;
  p.table = d->scale_table; p.table_len = d->table_len;
  p.n = d->n; p.scale_bound = d->scale_bound; p.lik_bound = d->likelihood_bound;
  p.seed_lo = static_cast<uint32_t>(d->philox_seed); p.seed_hi = static_cast<uint32_t>(d->philox_seed >> 32);
  p.off_lo = static_cast<uint32_t>(d->philox_offset); p.off_hi = static_cast<uint32_t>(d->philox_offset >> 32);

  bool vec = (d->n % 4 == 0);
  auto chk = [&](const void* ptr, int64_t bs) {
    if (ptr && (!aligned16(ptr) || (bs % 4) != 0)) vec = false;
  };
  chk(d->y, d->y_bs); chk(d->mu, d->mu_bs); chk(d->sigma, d->sigma_bs); chk(d->noise, d->noise_bs);
  chk(d->yhat, d->yhat_bs); chk(d->ste, d->ste_bs); chk(d->lik, d->lik_bs); chk(d->sym, d->sym_bs);
  chk(d->idx, d->idx_bs);

  // CTAs per image: each CTA walks `iters` grid-stride steps of kThreads groups so that the
  // prologue (table staging) and the fp64 reduction epilogue are amortised, while the grid
  // still fills every SM a few times over.
  const int64_t groups = vec ? d->n / 4 : d->n;
  const int64_t tiles = (groups + kThreads - 1) / kThreads;       // per image
  int64_t iters = gc_iters_target();
  const int64_t min_ctas = static_cast<int64_t>(sm_count()) * 4;
  while (iters > 1 && ((tiles + iters - 1) / iters) * d->B < min_ctas) --iters;
  int64_t bpi = (tiles + iters - 1) / iters;
  if (bpi > kMaxBpi) bpi = kMaxBpi;
  if (bpi < 1) bpi = 1;
  p.bpi = static_cast<int>(bpi);
  if (d->bits) {
    if (!d->workspace || d->workspace_bytes < reslic_workspace_bytes(d->B))
      return set_error(RESLIC_ERR_WORKSPACE, "gc_fwd: workspace missing or too small for `bits`");
    p.bits = d->bits;
    p.workspace = static_cast<double*>(d->workspace);
  }
  const int64_t grid64 = bpi * d->B;
  if (grid64 > 0x7fffffffLL) return set_error(RESLIC_ERR_ARG, "gc_fwd: grid too large");
  int steps = 6;
  if (need_idx && d->table_len - 1 > 63) steps = 8;
  const bool noise = d->mode == RESLIC_Q_NOISE;
  const int grid = static_cast<int>(grid64);
  cudaError_t err;
  const bool fast = math_mode() != RESLIC_MATH_MIRROR;
  if (need_lik) err = need_idx ? launch_noise<true, true>(p, grid, steps, vec, noise, fast, st)
                               : launch_noise<true, false>(p, grid, steps, vec, noise, fast, st);
  else err = need_idx ? launch_noise<false, true>(p, grid, steps, vec, noise, fast, st)
                      : launch_noise<false, false>(p, grid, steps, vec, noise, fast, st);
  if (err != cudaSuccess) return set_cuda_error(err, "gc_fwd launch");
  return RESLIC_OK;
}

}  // namespace reslic
