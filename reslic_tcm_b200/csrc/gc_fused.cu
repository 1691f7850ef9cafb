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
#include "gc_math.cuh"
#include "reslic_internal.h"

namespace reslic {

struct GcParams {
  const float* y; const float* mu; const float* sigma; const float* noise;
  int64_t y_bs, mu_bs, sigma_bs, noise_bs;
  float* yhat; float* ste; float* lik; int32_t* sym; int32_t* idx;
  int64_t yhat_bs, ste_bs, lik_bs, sym_bs, idx_bs;
  double* bits; unsigned long long* workspace; int bits_accumulate;
  const float* table; int table_len;
  int64_t n;          // elements per image
  int64_t B;          // images
  int64_t tiles_per_image;
  unsigned int tpi, total_tiles, q_tiles, r_tiles;   // 32-bit partition: CTA c owns q (+1 if c < r) tiles
  float scale_bound, lik_bound;
  uint32_t seed_lo, seed_hi, off_lo, off_hi;
};

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
  // round-to-nearest of (x + 0.5 - bias) ~ ceil(x - bias), read off the mantissa of x + 1.5*2^23
  // (no F2I: the XU pipe is the scarce one here); any wrong or wild guess is caught by the proof.
  int g = __float_as_int(fmaf(lg2_approx(s), g_scale, g_off) + 12582912.0f) - 0x4B400000;
  g = max(0, min(g, last));
  const float lo = pad[g], hi = pad[g + 1];
  if (!((lo < s) && (s <= hi))) g = min(scale_index_search<STEPS>(s, pad), last);
  return g;
}

// One pair of elements.  `prod` accumulates the product of the pair's (bounded) likelihoods
// so that the rate needs one MUFU.LG2 per four elements (see the caller); `acc` is used
// directly when the bound is too small for that.
template <bool NEED_LIK, bool NEED_IDX, bool NOISE, int STEPS, bool FAST, bool PAIR>
struct GcPair {
  __device__ __forceinline__ static void run(const GcParams& p, const float* pad, float g_scale, float g_off,
                                             F2 y, F2 mu, F2 sg, F2 u, F2& yhat, F2& ste, F2& lik,
                                             int2& sym, int2& idx, float& prod, float& acc, bool use_prod) {
    const F2 d = add2(y, neg2(mu));
    // torch.round (half to even) and the int32 symbol: for |d| < 2^22 adding 1.5*2^23 rounds d to
    // an integer in the FMA pipe and leaves that integer in the low mantissa bits; larger |d|
    // (never seen in latents, exercised by the edge-case tests) takes the FRND/F2I path.
    const F2 tq = add2(d, f2(12582912.0f));
    F2 q = add2(tq, f2(-12582912.0f));
    sym.x = __float_as_int(tq.x) - 0x4B400000;
    sym.y = __float_as_int(tq.y) - 0x4B400000;
    if (!(fmaxf(fabsf(d.x), fabsf(d.y)) < 4194304.0f) || d.x != d.x) {
      q.x = rintf(d.x); q.y = rintf(d.y);
      sym.x = __float2int_rn(q.x); sym.y = __float2int_rn(q.y);
    }
    ste = add2(q, mu);
    yhat = NOISE ? add2(y, u) : ste;
    const F2 s = max_nan2(sg, p.scale_bound);
    if (NEED_LIK) {
      const F2 v = abs2(add2(yhat, neg2(mu)));   // reference re-subtracts mu from the quantized value
      F2 L = gc_likelihood<FAST>(v, s);
      if (p.lik_bound > 0.0f) L = max_nan2(L, p.lik_bound);
      lik = L;
      if (!PAIR) {                       // scalar path: lane y is a dummy
        acc += FAST ? lg2_approx(L.x) : log2f(L.x);
      } else if (FAST) {
        if (use_prod) prod *= L.x * L.y;
        else acc += lg2_approx(L.x) + lg2_approx(L.y);
      } else {
        acc += log2f(L.x) + log2f(L.y);
      }
    }
    if (NEED_IDX) {
      idx.x = scale_index<STEPS>(s.x, pad, g_scale, g_off, p.table_len - 1);
      idx.y = PAIR ? scale_index<STEPS>(s.y, pad, g_scale, g_off, p.table_len - 1) : 0;
    }
  }
};

template <bool NEED_LIK, bool NEED_IDX, bool NOISE, bool VEC, int STEPS, bool FAST, bool PF>
__global__ void __launch_bounds__(kThreads, PF ? 4 : 5)
gc_fwd_kernel(const GcParams p) {
  __shared__ float pad[NEED_IDX ? kPadLen : 1];
  __shared__ float s_guess[2];
  float g_scale = 0.0f, g_off = 0.0f;
  // Programmatic dependent launch: this grid may have been scheduled while its predecessor
  // in the stream was still draining.  Nothing is read from global memory before the wait;
  // dependents are released at once so that THEIR launch overlaps this grid's execution.
  griddep_wait();
  griddep_launch_dependents();
  bool table_staged = !NEED_IDX;
  auto stage_table = [&]() {
    const float inf = __int_as_float(0x7f800000);
    for (int i = threadIdx.x; i < kPadLen; i += kThreads)
      pad[i] = (i == 0) ? -inf : ((i <= p.table_len - 1) ? p.table[i - 1] : inf);
    if (threadIdx.x == 0) {
      // guess(s) = ceil((log2 s - log2 t_0) * (len-2) / (log2 t_{len-2} - log2 t_0))
      float sc = 0.0f, off = 0.0f;
      if (p.table_len >= 3) {
        const float l0 = log2f(p.table[0]), l1 = log2f(p.table[p.table_len - 2]);
        // the small downward bias makes sigma == table[j] (notably the 0.11 bound itself) guess j
        if (l1 > l0) { sc = static_cast<float>(p.table_len - 2) / (l1 - l0); off = -l0 * sc - 2.44140625e-4f + 0.5f; }
      }
      s_guess[0] = sc; s_guess[1] = off;
    }
    __syncthreads();
    g_scale = s_guess[0]; g_off = s_guess[1];
    table_staged = true;
  };

  // Persistent CTAs: the slice is cut into tiles of kThreads element groups (a tile never
  // straddles two images); CTA c owns the contiguous tile range [c*T/G, (c+1)*T/G), so all
  // CTAs get the same work to within one tile and stay resident for the whole launch.  The
  // range is walked image segment by image segment: inside a segment every tensor pointer
  // just advances by one tile, and the next tile's loads are issued before the current tile
  // is computed (register double buffer), so a CTA never alternates between an all-loads and
  // an all-math phase.
  constexpr int W = VEC ? 4 : 1;
  constexpr int kTileElems = kThreads * W;
  const int groups = static_cast<int>(VEC ? (p.n >> 2) : p.n);      // per image (n < 2^31 checked on host)
  // CTA c owns tiles [c*q + min(c,r), ...) — q = T / G, r = T % G precomputed on the host, all in
  // 32-bit arithmetic (64-bit divisions here sat on the critical path of short launches)
  const unsigned int cta = blockIdx.x;
  unsigned int t = cta * p.q_tiles + min(cta, p.r_tiles);
  const unsigned int t_end = t + p.q_tiles + (cta < p.r_tiles ? 1u : 0u);

  struct In { float y[4], m[4], s[4], u[4]; };
  while (t < t_end) {
    const int image = static_cast<int>(t / p.tpi);
    const int chunk0 = static_cast<int>(t - image * p.tpi);
    const unsigned int seg_end = (static_cast<unsigned int>(image) + 1u) * p.tpi;
    const int ntiles = static_cast<int>(min(seg_end, t_end) - t);
    t += ntiles;
    int g = chunk0 * kThreads + threadIdx.x;                       // element group inside the image
    const int64_t e0 = static_cast<int64_t>(g) * W;
    const float* __restrict__ y = p.y ? p.y + image * p.y_bs + e0 : nullptr;
    const float* __restrict__ mu = p.mu ? p.mu + image * p.mu_bs + e0 : nullptr;
    const float* __restrict__ sg = p.sigma ? p.sigma + image * p.sigma_bs + e0 : nullptr;
    const float* __restrict__ nz = (NOISE && p.noise) ? p.noise + image * p.noise_bs + e0 : nullptr;
    float* yhat = p.yhat ? p.yhat + image * p.yhat_bs + e0 : nullptr;
    float* ste = p.ste ? p.ste + image * p.ste_bs + e0 : nullptr;
    float* lik = (NEED_LIK && p.lik) ? p.lik + image * p.lik_bs + e0 : nullptr;
    int32_t* sym = p.sym ? p.sym + image * p.sym_bs + e0 : nullptr;
    int32_t* idx = (NEED_IDX && p.idx) ? p.idx + image * p.idx_bs + e0 : nullptr;

    auto load = [&](In& r, int k) {          // tile k of this segment (k*kTileElems ahead of the base)
      const int off = k * kTileElems;
#pragma unroll
      for (int j = 0; j < 4; ++j) { r.y[j] = 0.f; r.m[j] = 0.f; r.s[j] = 1.f; r.u[j] = 0.f; }
      if (g + k * kThreads >= groups) return;
      if (VEC) {
        if (y) { const float4 v = ld_stream4(y + off); r.y[0] = v.x; r.y[1] = v.y; r.y[2] = v.z; r.y[3] = v.w; }
        if (mu) { const float4 v = ld_stream4(mu + off); r.m[0] = v.x; r.m[1] = v.y; r.m[2] = v.z; r.m[3] = v.w; }
        if (sg) { const float4 v = ld_stream4(sg + off); r.s[0] = v.x; r.s[1] = v.y; r.s[2] = v.z; r.s[3] = v.w; }
        if (NOISE && nz) { const float4 v = ld_stream4(nz + off); r.u[0] = v.x; r.u[1] = v.y; r.u[2] = v.z; r.u[3] = v.w; }
      } else {
        if (y) r.y[0] = ld_stream1(y + off);
        if (mu) r.m[0] = ld_stream1(mu + off);
        if (sg) r.s[0] = ld_stream1(sg + off);
        if (NOISE && nz) r.u[0] = ld_stream1(nz + off);
      }
    };

    float acc = 0.0f;
    In cur;
    load(cur, 0);
    if (!table_staged) stage_table();      // the first tile's loads are already in flight
    for (int k = 0; k < ntiles; ++k) {
      In nxt;
      if (PF && k + 1 < ntiles) load(nxt, k + 1);
      const int gk = g + k * kThreads;
      if (gk < groups) {
        const int off = k * kTileElems;
        if (NOISE && !nz) {
          // counter = global element-group id; one Philox call feeds 4 consecutive elements
          const uint64_t gq = static_cast<uint64_t>(VEC ? gk : (gk >> 2));
          const uint64_t gid = static_cast<uint64_t>(image) * static_cast<uint64_t>((p.n + 3) >> 2) + gq;
          const Philox4 r = philox4x32_10(static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32),
                                          p.off_lo, p.off_hi, p.seed_lo, p.seed_hi);
          if (VEC) {
            cur.u[0] = u32_to_centered_uniform(r.x); cur.u[1] = u32_to_centered_uniform(r.y);
            cur.u[2] = u32_to_centered_uniform(r.z); cur.u[3] = u32_to_centered_uniform(r.w);
          } else {
            const int q4 = gk & 3;
            cur.u[0] = u32_to_centered_uniform(q4 == 0 ? r.x : q4 == 1 ? r.y : q4 == 2 ? r.z : r.w);
          }
        }
        float oy[4], os[4], ol[4]; int osym[4], oidx[4];
        // rate: FAST mode multiplies the four bounded likelihoods of a group and takes ONE
        // MUFU.LG2 (abs err <= 2^-22 for arguments in (0.5,2), rel err 2^-22 elsewhere); the
        // product cannot underflow while the bound is >= 1e-9 (1e-36 > FLT_MIN).
        const bool use_prod = FAST && p.lik_bound >= 1e-9f;
        float prod = 1.0f;
#pragma unroll
        for (int j = 0; j < (VEC ? 2 : 1); ++j) {
          const int j0 = 2 * j, j1 = 2 * j + 1;
          F2 yh, st, lk; int2 sy, ix;
          GcPair<NEED_LIK, NEED_IDX, NOISE, STEPS, FAST, VEC>::run(
              p, pad, g_scale, g_off, make_float2(cur.y[j0], cur.y[j1]), make_float2(cur.m[j0], cur.m[j1]),
              make_float2(cur.s[j0], cur.s[j1]), make_float2(cur.u[j0], cur.u[j1]), yh, st, lk, sy, ix, prod, acc,
              use_prod && VEC);
          oy[j0] = yh.x; os[j0] = st.x; ol[j0] = lk.x; osym[j0] = sy.x; oidx[j0] = ix.x;
          if (VEC) { oy[j1] = yh.y; os[j1] = st.y; ol[j1] = lk.y; osym[j1] = sy.y; oidx[j1] = ix.y; }
        }
        if (NEED_LIK && use_prod && VEC) acc += lg2_approx(prod);
        if (VEC) {
          if (yhat) st_stream4(yhat + off, make_float4(oy[0], oy[1], oy[2], oy[3]));
          if (ste) st_stream4(ste + off, make_float4(os[0], os[1], os[2], os[3]));
          if (lik) st_stream4(lik + off, make_float4(ol[0], ol[1], ol[2], ol[3]));
          if (sym) st_stream4(sym + off, make_int4(osym[0], osym[1], osym[2], osym[3]));
          if (idx) st_stream4(idx + off, make_int4(oidx[0], oidx[1], oidx[2], oidx[3]));
        } else {
          if (yhat) st_stream1(yhat + off, oy[0]);
          if (ste) st_stream1(ste + off, os[0]);
          if (lik) st_stream1(lik + off, ol[0]);
          if (sym) st_stream1(sym + off, osym[0]);
          if (idx) st_stream1(idx + off, oidx[0]);
        }
      }
      if (PF) cur = nxt;
      else if (k + 1 < ntiles) load(cur, k + 1);
    }
    if (NEED_LIK && p.bits) {
      // warps committing to this image = (CTAs whose tile range meets the image) * warps per CTA;
      // tile tau belongs to CTA ceil((tau+1)*G/T) - 1.
      // owner of tile tau: the first r CTAs hold q+1 tiles each, the rest q
      auto owner = [&](unsigned int tau) -> unsigned int {
        const unsigned int big = p.r_tiles * (p.q_tiles + 1u);
        return tau < big ? tau / (p.q_tiles + 1u) : p.r_tiles + (tau - big) / p.q_tiles;
      };
      const unsigned int first = static_cast<unsigned int>(image) * p.tpi;
      const unsigned int n_ctas = owner(first + p.tpi - 1u) - owner(first) + 1u;
      rate_commit(acc, image, n_ctas * (kThreads / 32), p.B, p.workspace, p.bits, p.bits_accumulate != 0);
    }
  }
}

// ------------------------------------------------------------------ host-side launch
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// resident CTAs per SM of one kernel instantiation (queried once per instantiation)
template <typename K>
static int resident_ctas(K kernel, int* cache) {
  if (*cache == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kThreads, 0) != cudaSuccess || n < 1) n = 1;
    *cache = n;
  }
  return *cache;
}

template <bool NEED_LIK, bool NEED_IDX, bool NOISE, bool VEC, int STEPS, bool FAST, bool PF>
static cudaError_t launch_pf(GcParams& p, cudaStream_t st) {
  static int occ = 0;
  auto kernel = gc_fwd_kernel<NEED_LIK, NEED_IDX, NOISE, VEC, STEPS, FAST, PF>;
  const int64_t total = p.tiles_per_image * p.B;
  const int waves = gc_tuning().ctas_per_sm;   // 0: one tile per CTA; k>0: k CTAs per SM; <0: resident count
  int64_t grid = total;
  if (waves > 0) grid = static_cast<int64_t>(waves) * sm_count();
  else if (waves < 0) grid = static_cast<int64_t>(resident_ctas(kernel, &occ)) * sm_count();
  if (grid > total) grid = total;
  p.tpi = static_cast<unsigned int>(p.tiles_per_image);
  p.total_tiles = static_cast<unsigned int>(total);
  p.q_tiles = static_cast<unsigned int>(total / grid);
  p.r_tiles = static_cast<unsigned int>(total % grid);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = gc_tuning().pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, p);
}
template <bool NEED_LIK, bool NEED_IDX, bool NOISE, bool VEC, int STEPS, bool FAST>
static cudaError_t launch_one(GcParams& p, cudaStream_t st) {
  return gc_tuning().prefetch ? launch_pf<NEED_LIK, NEED_IDX, NOISE, VEC, STEPS, FAST, true>(p, st)
                              : launch_pf<NEED_LIK, NEED_IDX, NOISE, VEC, STEPS, FAST, false>(p, st);
}
template <bool NEED_LIK, bool NEED_IDX, bool NOISE, bool VEC>
static cudaError_t launch_steps(GcParams& p, int steps, bool fast, cudaStream_t st) {
  // the math policy only matters when a likelihood is computed
  if (NEED_LIK && !fast)
    return steps <= 6 ? launch_one<NEED_LIK, NEED_IDX, NOISE, VEC, 6, false>(p, st)
                      : launch_one<NEED_LIK, NEED_IDX, NOISE, VEC, 8, false>(p, st);
  return steps <= 6 ? launch_one<NEED_LIK, NEED_IDX, NOISE, VEC, 6, true>(p, st)
                    : launch_one<NEED_LIK, NEED_IDX, NOISE, VEC, 8, true>(p, st);
}
template <bool NEED_LIK, bool NEED_IDX, bool NOISE>
static cudaError_t launch_vec(GcParams& p, int steps, bool vec, bool fast, cudaStream_t st) {
  return vec ? launch_steps<NEED_LIK, NEED_IDX, NOISE, true>(p, steps, fast, st)
             : launch_steps<NEED_LIK, NEED_IDX, NOISE, false>(p, steps, fast, st);
}
template <bool NEED_LIK, bool NEED_IDX>
static cudaError_t launch_noise(GcParams& p, int steps, bool vec, bool noise, bool fast, cudaStream_t st) {
  return noise ? launch_vec<NEED_LIK, NEED_IDX, true>(p, steps, vec, fast, st)
               : launch_vec<NEED_LIK, NEED_IDX, false>(p, steps, vec, fast, st);
}

int gc_fwd_launch(const reslic_gc_desc* d, cudaStream_t st) {
  if (!d) return set_error(RESLIC_ERR_ARG, "gc_fwd: null descriptor");
  if (d->B < 0 || d->n < 0) return set_error(RESLIC_ERR_ARG, "gc_fwd: negative size");
  if (d->B == 0 || d->n == 0) return RESLIC_OK;  // empty input: nothing to do
  if (d->B > (1 << 24)) return set_error(RESLIC_ERR_ARG, "gc_fwd: B too large");
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
  p.n = d->n; p.B = d->B; p.scale_bound = d->scale_bound; p.lik_bound = d->likelihood_bound;
  p.seed_lo = static_cast<uint32_t>(d->philox_seed); p.seed_hi = static_cast<uint32_t>(d->philox_seed >> 32);
  p.off_lo = static_cast<uint32_t>(d->philox_offset); p.off_hi = static_cast<uint32_t>(d->philox_offset >> 32);

  bool vec = (d->n % 4 == 0);
  auto chk = [&](const void* ptr, int64_t bs) {
    if (ptr && (!aligned16(ptr) || (bs % 4) != 0)) vec = false;
  };
  chk(d->y, d->y_bs); chk(d->mu, d->mu_bs); chk(d->sigma, d->sigma_bs); chk(d->noise, d->noise_bs);
  chk(d->yhat, d->yhat_bs); chk(d->ste, d->ste_bs); chk(d->lik, d->lik_bs); chk(d->sym, d->sym_bs);
  chk(d->idx, d->idx_bs);

  const int64_t groups = vec ? d->n / 4 : d->n;
  p.tiles_per_image = (groups + kThreads - 1) / kThreads;
  if (d->n >= (1LL << 31)) return set_error(RESLIC_ERR_ARG, "gc_fwd: more than 2^31 elements per image");
  if (p.tiles_per_image * d->B >= (1LL << 31)) return set_error(RESLIC_ERR_ARG, "gc_fwd: input too large (>= 2^31 tiles)");
  if (d->bits) {
    if (!d->workspace || d->workspace_bytes < reslic_workspace_bytes(d->B))
      return set_error(RESLIC_ERR_WORKSPACE, "gc_fwd: workspace missing or too small for `bits`");
    if (reinterpret_cast<uintptr_t>(d->workspace) & 7u)
      return set_error(RESLIC_ERR_WORKSPACE, "gc_fwd: workspace must be 8-byte aligned");
    p.bits = d->bits;
    p.bits_accumulate = d->bits_accumulate;
    p.workspace = static_cast<unsigned long long*>(d->workspace);
  }
  int steps = 6;
  if (need_idx && d->table_len - 1 > 63) steps = 8;
  const bool noise = d->mode == RESLIC_Q_NOISE;
  const bool fast = math_mode() != RESLIC_MATH_MIRROR;
  cudaError_t err;
  if (need_lik) err = need_idx ? launch_noise<true, true>(p, steps, vec, noise, fast, st)
                               : launch_noise<true, false>(p, steps, vec, noise, fast, st);
  else err = need_idx ? launch_noise<false, true>(p, steps, vec, noise, fast, st)
                      : launch_noise<false, false>(p, steps, vec, noise, fast, st);
  if (err != cudaSuccess) return set_cuda_error(err, "gc_fwd launch");
  return RESLIC_OK;
}

}  // namespace reslic
