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
// CTA in shared memory, per-image rate reduced by warp shuffles + one 64-bit integer atomic
// per warp and image, and an instruction budget (~73 issue slots per element in the loop)
// small enough that FP32/MUFU issue stays under the HBM time.  No tensor cores (nothing
// GEMM-shaped here).  DESIGN.md section 3.1 has the measurements behind each choice.
#include "common.cuh"
#include "gc_math.cuh"
#include "reslic_internal.h"
#include <limits>

namespace reslic {

struct GcParams {
  const float* y; const float* mu; const float* sigma; const float* noise;
  uint32_t y_bs, mu_bs, sigma_bs, noise_bs;   // batch strides in elements (< 2^32, checked on the host)
  float* yhat; float* ste; float* lik; int32_t* sym; int32_t* idx;
  uint32_t yhat_bs, ste_bs, lik_bs, sym_bs, idx_bs;
  double* bits; unsigned long long* workspace; int bits_accumulate;
  const float* table; int table_len;
  int64_t n;          // elements per image
  int64_t B;          // images
  int64_t tiles_per_image;
  unsigned int tpi, total_tiles, q_tiles, r_tiles;   // 32-bit partition: CTA c owns q (+1 if c < r) tiles
  float scale_bound;
  float lik_floor;    // the likelihood bound, or -inf when the bound is disabled (max with it is then a no-op)
  float d_limit;      // groups with max(|y-mu|, sigma) < d_limit take the clamp-free path; 0 sends every group to the general one
  uint32_t seed_lo, seed_hi, off_lo, off_hi;
  const float* next_y; uint32_t next_y_bs;   // optional: the y the NEXT launch will read (same shape), prefetched into L2
  RateEx ex;                                 // multi-GPU rate exchange of the collecting launch; world == 0: none
#ifdef RESLIC_TRACE
  unsigned long long* trace;                 // development builds (tools/dev): per-CTA %globaltimer stamps
#endif
};
#ifdef RESLIC_TRACE
static unsigned long long* g_trace_next = nullptr;
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned int smid() { unsigned int r; asm volatile("mov.u32 %0, %smid;" : "=r"(r)); return r; }
#define TRACE(j) do { if (p.trace && threadIdx.x == 0) p.trace[blockIdx.x * 6 + (j)] = gtime(); } while (0)
#else
#define TRACE(j) do {} while (0)
#endif

// build_indexes: idx = #{ j < len-1 : !(s <= table[j]) } = (len-1) - sum_j [s <= table[j]].
// Shared-memory layout: pad2[g] = (lo_g, hi_g) with lo_0 = -inf, lo_g = table[g-1], hi_g = lo_{g+1}
// and +inf beyond table[len-2], so   idx = g  <=>  lo_g < s <= hi_g.   The scale table is log-spaced
// (exp(linspace(ln .11, ln 256, 64)), tcm.py:26-34), so g is GUESSED from one MUFU.LG2 and a fused
// multiply-add, then PROVEN with the two exact comparisons the reference makes (one 64-bit shared
// load); a failed proof (s within rounding of a table value, a table that is not log-spaced, NaN)
// falls back to the branch-free binary search.  The result is bit-exact for any table.
constexpr int kPadLen = 256;

__device__ __forceinline__ int scale_index_search(float s, const float2* pad2) {
  int lo = 0;
#pragma unroll
  for (int step = 128; step > 0; step >>= 1) {
    const float t = pad2[lo + step].x;         // == table[lo + step - 1], +inf past the table
    lo += (!(s <= t)) ? step : 0;
  }
  return lo;  // NaN walks to 255; the caller clamps to table_len - 1
}
__device__ __noinline__ int scale_index_cold(float s, const float2* pad2, int last) {
  return min(scale_index_search(s, pad2), last);
}
// the guess: round-to-nearest of (x + 0.5 - bias) ~ ceil(x - bias), read off the mantissa of
// x + 1.5*2^23 (no F2I: the XU pipe is the scarce one here); a wrong or wild guess fails the proof.
__device__ __forceinline__ int scale_index_guess(float s, float g_scale, float g_off, int last) {
  const int g = __float_as_int(fmaf(lg2_approx(s), g_scale, g_off) + 12582912.0f) - 0x4B400000;
  return max(0, min(g, last));
}

__device__ __forceinline__ float max3_nan(float a, float b, float c) {
  float r;
  asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// ------------------------------------------------------------------ one element, any input
// The general path: every operand class the reference accepts (|y-mu| >= 2^22, inf, NaN, huge or
// denormal sigma) — used by the scalar kernel (unaligned tensors, n % 4 != 0) and, out of line,
// by the vector kernel for the rare group that fails its range check.
template <bool NEED_LIK, bool NEED_IDX, bool NOISE, bool FAST>
__device__ __forceinline__ void gc_elem(float y, float mu, float sg, float u, float scale_bound, float lik_floor,
                                        const float2* pad2, int last, float& yhat, float& ste, float& lik,
                                        int& sym, int& idx, float& acc) {
  const float q = rintf(y - mu);               // torch.round: half to even
  sym = __float2int_rn(q);
  ste = q + mu;
  yhat = NOISE ? y + u : ste;
  const float s = max_nan(sg, scale_bound);
  lik = 0.0f; idx = 0;
  if (NEED_LIK) {
    const float v = fabsf(yhat - mu);          // the reference re-subtracts mu from the quantized value
    const float L = max_nan(gc_likelihood<FAST>(f2(v), f2(s)).x, lik_floor);
    lik = L;
    acc += FAST ? lg2_approx(L) : log2f(L);
  }
  if (NEED_IDX) idx = min(scale_index_search(s, pad2), last);
}

struct GcOutPtrs { float* yhat; float* ste; float* lik; int32_t* sym; int32_t* idx; };

template <bool NEED_LIK, bool NEED_IDX, bool NOISE, bool FAST>
__device__ __noinline__ float gc_group_cold(float4 y, float4 m, float4 s, float4 u, float scale_bound, float lik_floor,
                                            const float2* pad2, int last, GcOutPtrs o) {
  float acc = 0.0f;
  float4 oy, os, ol; int4 osym, oidx;
  gc_elem<NEED_LIK, NEED_IDX, NOISE, FAST>(y.x, m.x, s.x, u.x, scale_bound, lik_floor, pad2, last, oy.x, os.x, ol.x, osym.x, oidx.x, acc);
  gc_elem<NEED_LIK, NEED_IDX, NOISE, FAST>(y.y, m.y, s.y, u.y, scale_bound, lik_floor, pad2, last, oy.y, os.y, ol.y, osym.y, oidx.y, acc);
  gc_elem<NEED_LIK, NEED_IDX, NOISE, FAST>(y.z, m.z, s.z, u.z, scale_bound, lik_floor, pad2, last, oy.z, os.z, ol.z, osym.z, oidx.z, acc);
  gc_elem<NEED_LIK, NEED_IDX, NOISE, FAST>(y.w, m.w, s.w, u.w, scale_bound, lik_floor, pad2, last, oy.w, os.w, ol.w, osym.w, oidx.w, acc);
  if (o.yhat) st_stream4(o.yhat, oy);
  if (o.ste) st_stream4(o.ste, os);
  if (NEED_LIK && o.lik) st_stream4(o.lik, ol);
  if (o.sym) st_stream4(o.sym, osym);
  if (NEED_IDX && o.idx) st_stream4(o.idx, oidx);
  return acc;
}

// ------------------------------------------------------------------ four elements, in range
// One range check per group — max(|d_0..3|, s_0..3) < d_limit (2^22), NaN fails — proves that every
// intermediate below is finite, so the hot path carries no clamp: round and symbol come from one
// magic-number add (adding 1.5*2^23 rounds d to an integer, half to even, in the FMA pipe and
// leaves that integer in the low mantissa bits), the likelihood from gc_likelihood_finite, and
// the rate from ONE MUFU.LG2 of the product of the four bounded likelihoods (abs err <= 2^-22
// for arguments in (0.5,2), rel err 2^-22 elsewhere; the host only enables this path while the
// bound is >= 1e-9, so the product stays above 1e-36 > FLT_MIN).
struct GcIn { float4 y, m, s, u; };

template <bool NEED_LIK, bool NEED_IDX, bool NOISE, bool FAST>
__device__ __forceinline__ void gc_group_vec(const GcParams& p, const float2* pad2, float g_scale, float g_off,
                                             const GcIn& in, const GcOutPtrs& base, unsigned int bo, float& acc) {
  auto at = [bo](auto* ptr) { return reinterpret_cast<decltype(ptr)>(reinterpret_cast<char*>(ptr) + bo); };
  const F2 y0 = make_float2(in.y.x, in.y.y), y1 = make_float2(in.y.z, in.y.w);
  const F2 m0 = make_float2(in.m.x, in.m.y), m1 = make_float2(in.m.z, in.m.w);
  const F2 d0 = add2(y0, neg2(m0)), d1 = add2(y1, neg2(m1));
  const F2 s0 = max_nan2(make_float2(in.s.x, in.s.y), p.scale_bound);
  const F2 s1 = max_nan2(make_float2(in.s.z, in.s.w), p.scale_bound);
  float mx = max3_nan(fabsf(d0.x), fabsf(d0.y), fabsf(d1.x));
  if (NEED_LIK || NEED_IDX) {
    mx = max3_nan(mx, fabsf(d1.y), s0.x);
    mx = max3_nan(mx, s0.y, s1.x);
    mx = max_nan(mx, s1.y);
  } else {
    mx = max_nan(mx, fabsf(d1.y));
  }
  if (NOISE) {   // caller-supplied noise is data too
    mx = max3_nan(mx, fabsf(in.u.x), fabsf(in.u.y));
    mx = max3_nan(mx, fabsf(in.u.z), fabsf(in.u.w));
  }
  const int last = p.table_len - 1;
  if (!(mx < p.d_limit)) {
    GcOutPtrs o;
    o.yhat = base.yhat ? at(base.yhat) : nullptr; o.ste = base.ste ? at(base.ste) : nullptr;
    o.lik = base.lik ? at(base.lik) : nullptr; o.sym = base.sym ? at(base.sym) : nullptr;
    o.idx = base.idx ? at(base.idx) : nullptr;
    acc += gc_group_cold<NEED_LIK, NEED_IDX, NOISE, FAST>(in.y, in.m, in.s, in.u, p.scale_bound, p.lik_floor, pad2, last, o);
    return;
  }
  const F2 magic = f2(12582912.0f);
  const F2 t0 = add2(d0, magic), t1 = add2(d1, magic);
  const F2 q0 = add2(t0, neg2(magic)), q1 = add2(t1, neg2(magic));
  const F2 e0 = add2(q0, m0), e1 = add2(q1, m1);                 // ste_round output
  F2 h0 = e0, h1 = e1;                                           // quantize() output
  if (NOISE) { h0 = add2(y0, make_float2(in.u.x, in.u.y)); h1 = add2(y1, make_float2(in.u.z, in.u.w)); }
  if (base.yhat) st_stream4(at(base.yhat), make_float4(h0.x, h0.y, h1.x, h1.y));
  if (base.ste) st_stream4(at(base.ste), make_float4(e0.x, e0.y, e1.x, e1.y));
  if (base.sym)
    st_stream4(at(base.sym), make_int4(__float_as_int(t0.x) - 0x4B400000, __float_as_int(t0.y) - 0x4B400000,
                                       __float_as_int(t1.x) - 0x4B400000, __float_as_int(t1.y) - 0x4B400000));
  if (NEED_LIK) {
    const F2 v0 = abs2(add2(h0, neg2(m0))), v1 = abs2(add2(h1, neg2(m1)));   // the reference re-subtracts mu
    const F2 L0 = max_nan2(gc_likelihood_finite<FAST>(v0, s0), p.lik_floor);
    const F2 L1 = max_nan2(gc_likelihood_finite<FAST>(v1, s1), p.lik_floor);
    if (base.lik) st_stream4(at(base.lik), make_float4(L0.x, L0.y, L1.x, L1.y));
    if (FAST) acc += lg2_approx((L0.x * L0.y) * (L1.x * L1.y));
    else acc += (log2f(L0.x) + log2f(L0.y)) + (log2f(L1.x) + log2f(L1.y));
  }
  if (NEED_IDX) {
    int g0 = scale_index_guess(s0.x, g_scale, g_off, last), g1 = scale_index_guess(s0.y, g_scale, g_off, last);
    int g2 = scale_index_guess(s1.x, g_scale, g_off, last), g3 = scale_index_guess(s1.y, g_scale, g_off, last);
    const float2 b0 = pad2[g0], b1 = pad2[g1], b2 = pad2[g2], b3 = pad2[g3];
    const bool proven = (b0.x < s0.x) && (s0.x <= b0.y) && (b1.x < s0.y) && (s0.y <= b1.y) &&
                        (b2.x < s1.x) && (s1.x <= b2.y) && (b3.x < s1.y) && (s1.y <= b3.y);
    if (!proven) {
      g0 = scale_index_cold(s0.x, pad2, last); g1 = scale_index_cold(s0.y, pad2, last);
      g2 = scale_index_cold(s1.x, pad2, last); g3 = scale_index_cold(s1.y, pad2, last);
    }
    if (base.idx) st_stream4(at(base.idx), make_int4(g0, g1, g2, g3));
  }
}

// MINB = resident CTAs per SM the register budget is cut for: 4 (64 registers, no spill in the loop) is
// the faster build for long launches; 5 (48 registers, a few spilled scalars, 25 % more warps to hide
// the first-tile latency) wins on short ones (a TCM slice: < 4 tiles per CTA) — measured 14.1 vs 14.5 us
// on 24 x 98304 and 66.8 vs 57.3 us on 24 x 491520, so the host picks by launch size.
// EXCH: the launch also publishes the batch's rate to every rank (reslic_rate_exchange).  A template flag, not a
// run-time test: with the publish path merely PRESENT in the kernel the 48-register build spills three times as much
// (117 LDL against 39 in the loop), which costs every launch 30 % — so only the collecting launch of a multi-GPU
// step runs the instantiation that contains it.
template <bool NEED_LIK, bool NEED_IDX, bool NOISE, bool VEC, bool FAST, int MINB, bool EXCH>
__global__ void __launch_bounds__(kThreads, MINB)
gc_fwd_kernel(const GcParams p) {
  __shared__ float2 pad2[NEED_IDX ? kPadLen : 1];
  __shared__ float s_guess[2];
  float g_scale = 0.0f, g_off = 0.0f;
  TRACE(0);
  // Programmatic dependent launch: this grid may have been scheduled while its predecessor
  // in the stream was still draining.  Nothing is read from global memory before the wait;
  // dependents are released at once so that THEIR launch overlaps this grid's execution.
  griddep_wait();
  griddep_launch_dependents();
  TRACE(1);
  bool table_staged = !NEED_IDX;
  auto stage_table = [&]() {
    const float inf = __int_as_float(0x7f800000);
    const int last = p.table_len - 1;
    auto entry = [&](int i) { return (i == 0) ? -inf : ((i <= last) ? p.table[i - 1] : inf); };
    for (int i = threadIdx.x; i < kPadLen; i += kThreads) pad2[i] = make_float2(entry(i), entry(i + 1));
    if (threadIdx.x == 0) {
      // guess(s) = ceil((log2 s - log2 t_0) * (len-2) / (log2 t_{len-2} - log2 t_0))
      float sc = 0.0f, off = 0.0f;
      if (p.table_len >= 3) {
        // approximate log/divide are enough: this only seeds the guess, the proof is exact
        const float l0 = lg2_approx(p.table[0]), l1 = lg2_approx(p.table[p.table_len - 2]);
        // the small downward bias makes sigma == table[j] (notably the 0.11 bound itself) guess j
        if (l1 > l0) { sc = __fdividef(static_cast<float>(p.table_len - 2), l1 - l0); off = -l0 * sc - 2.44140625e-4f + 0.5f; }
      }
      s_guess[0] = sc; s_guess[1] = off;
    }
    __syncthreads();
    g_scale = s_guess[0]; g_off = s_guess[1];
    table_staged = true;
  };

  // Persistent CTAs: the slice is cut into tiles of kThreads element groups (a tile never
  // straddles two images); CTA c owns the contiguous tile range [c*q + min(c,r), ...) — q = T / G,
  // r = T % G precomputed on the host in 32-bit arithmetic — so all CTAs get the same work to
  // within one tile and stay resident for the whole launch.  The range is walked image segment by
  // image segment.  Inside a segment every tensor is addressed as (segment base, uniform) +
  // (32-bit byte offset, per thread): no per-thread pointer registers, two integer adds per access.
  // Tiles are double-buffered in registers by a two-tile ping-pong (A computes while B loads and
  // vice versa), so a CTA never alternates between an all-loads and an all-math phase and no
  // register is ever copied from a "next" to a "current" buffer.
  constexpr int W = VEC ? 4 : 1;
  constexpr unsigned int kTileBytes = kThreads * W * 4;
  const int groups = static_cast<int>(VEC ? (p.n >> 2) : p.n);      // per image (n < 2^31 checked on host)
  const unsigned int cta = blockIdx.x;
  unsigned int t = cta * p.q_tiles + min(cta, p.r_tiles);
  const unsigned int t_end = t + p.q_tiles + (cta < p.r_tiles ? 1u : 0u);
  const unsigned int tb = threadIdx.x * (W * 4);

  GcIn a, b;
  a.y = a.m = a.u = make_float4(0.f, 0.f, 0.f, 0.f); a.s = make_float4(1.f, 1.f, 1.f, 1.f);
  b = a;                                     // absent inputs keep these defaults for the whole launch
  while (t < t_end) {
    const int image = static_cast<int>(t / p.tpi);
    const int chunk0 = static_cast<int>(t - image * p.tpi);
    const unsigned int seg_end = (static_cast<unsigned int>(image) + 1u) * p.tpi;
    const int ntiles = static_cast<int>(min(seg_end, t_end) - t);
    t += ntiles;
    const int64_t seg = static_cast<int64_t>(chunk0) * (kThreads * W);          // element offset of tile 0 in the image
    // image base + segment start: one 32x32->64 multiply-add per tensor on the uniform datapath
    auto at_seg = [&](auto* base, uint32_t bs) {
      return base + (static_cast<uint64_t>(static_cast<uint32_t>(image)) * bs + static_cast<uint64_t>(seg));
    };
    const float* y = p.y ? at_seg(p.y, p.y_bs) : nullptr;
    const float* mu = p.mu ? at_seg(p.mu, p.mu_bs) : nullptr;
    const float* sg = p.sigma ? at_seg(p.sigma, p.sigma_bs) : nullptr;
    const float* nz = (NOISE && p.noise) ? at_seg(p.noise, p.noise_bs) : nullptr;
    GcOutPtrs o;
    o.yhat = p.yhat ? at_seg(p.yhat, p.yhat_bs) : nullptr;
    o.ste = p.ste ? at_seg(p.ste, p.ste_bs) : nullptr;
    o.lik = (NEED_LIK && p.lik) ? at_seg(p.lik, p.lik_bs) : nullptr;
    o.sym = p.sym ? at_seg(p.sym, p.sym_bs) : nullptr;
    o.idx = (NEED_IDX && p.idx) ? at_seg(p.idx, p.idx_bs) : nullptr;
    const int g = chunk0 * kThreads + threadIdx.x;                 // this thread's element group in tile 0

    auto load = [&](GcIn& r, int k) {        // tile k of this segment
      if (g + k * kThreads >= groups) return;
      const unsigned int bo = tb + static_cast<unsigned int>(k) * kTileBytes;
      auto at = [bo](const float* ptr) { return reinterpret_cast<const float*>(reinterpret_cast<const char*>(ptr) + bo); };
      if (VEC) {
        if (y) r.y = ld_stream4(at(y));
        if (mu) r.m = ld_stream4(at(mu));
        if (sg) r.s = ld_stream4(at(sg));
        if (NOISE && nz) r.u = ld_stream4(at(nz));
      } else {
        if (y) r.y.x = ld_stream1(at(y));
        if (mu) r.m.x = ld_stream1(at(mu));
        if (sg) r.s.x = ld_stream1(at(sg));
        if (NOISE && nz) r.u.x = ld_stream1(at(nz));
      }
    };
    float acc = 0.0f;
    auto compute = [&](GcIn& r, int k) {
      const int gk = g + k * kThreads;
      if (gk >= groups) return;
      const unsigned int bo = tb + static_cast<unsigned int>(k) * kTileBytes;
      if (NOISE && !nz) {
        // counter = global element-group id; one Philox call feeds 4 consecutive elements
        const uint64_t gq = static_cast<uint64_t>(VEC ? gk : (gk >> 2));
        const uint64_t gid = static_cast<uint64_t>(image) * static_cast<uint64_t>((p.n + 3) >> 2) + gq;
        const Philox4 x = philox4x32_10(static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32),
                                        p.off_lo, p.off_hi, p.seed_lo, p.seed_hi);
        if (VEC) {
          r.u = make_float4(u32_to_centered_uniform(x.x), u32_to_centered_uniform(x.y),
                            u32_to_centered_uniform(x.z), u32_to_centered_uniform(x.w));
        } else {
          const int q4 = gk & 3;
          r.u.x = u32_to_centered_uniform(q4 == 0 ? x.x : q4 == 1 ? x.y : q4 == 2 ? x.z : x.w);
        }
      }
      if (VEC) {
        gc_group_vec<NEED_LIK, NEED_IDX, NOISE, FAST>(p, pad2, g_scale, g_off, r, o, bo, acc);
      } else {
        float yh, st, lk; int sy, ix;
        gc_elem<NEED_LIK, NEED_IDX, NOISE, FAST>(r.y.x, r.m.x, r.s.x, r.u.x, p.scale_bound, p.lik_floor, pad2,
                                                 p.table_len - 1, yh, st, lk, sy, ix, acc);
        auto at = [bo](auto* ptr) { return reinterpret_cast<decltype(ptr)>(reinterpret_cast<char*>(ptr) + bo); };
        if (o.yhat) st_stream1(at(o.yhat), yh);
        if (o.ste) st_stream1(at(o.ste), st);
        if (o.lik) st_stream1(at(o.lik), lk);
        if (o.sym) st_stream1(at(o.sym), sy);
        if (o.idx) st_stream1(at(o.idx), ix);
      }
    };

    load(a, 0);
    if (!table_staged) stage_table();      // the first tile's loads are already in flight
    for (int k = 0; k < ntiles; k += 2) {
      if (k + 1 < ntiles) load(b, k + 1);
      compute(a, k);
      if (k == 0) TRACE(2);
      if (k + 2 < ntiles) load(a, k + 2);
      if (k + 1 < ntiles) compute(b, k + 1);
    }
    if (NEED_LIK && p.bits) {
      if (p.bits_accumulate == RESLIC_RATE_DEFERRED) {
        rate_defer(acc, image, p.B, p.workspace);      // fire and forget: the warp retires without a round trip
      } else {
        // warps committing to this image = (CTAs whose tile range meets the image) * warps per CTA;
        // owner of tile tau: the first r CTAs hold q+1 tiles each, the rest q
        auto owner = [&](unsigned int tau) -> unsigned int {
          const unsigned int big = p.r_tiles * (p.q_tiles + 1u);
          return tau < big ? tau / (p.q_tiles + 1u) : p.r_tiles + (tau - big) / p.q_tiles;
        };
        const unsigned int first = static_cast<unsigned int>(image) * p.tpi;
        const unsigned int n_ctas = owner(first + p.tpi - 1u) - owner(first) + 1u;
        rate_commit(acc, image, n_ctas * (kThreads / 32), p.B, p.workspace, p.bits, p.bits_accumulate, EXCH ? &p.ex : nullptr);
      }
    }
  }
  // L2 prefetch of the next launch's y over this CTA's own tile range (TCM walks the five channel slices of ONE y
  // tensor, tcm.py:438-457: the next slice's y exists long before its mu / sigma do).  A slice launch is short —
  // about 3 tiles per CTA — so its first and last microsecond leave HBM under-used; CTAs that are done early pull
  // the next launch's first third of reads into that gap with one bulk instruction per image segment.
  TRACE(3);
#ifdef RESLIC_TRACE
  if (p.trace && threadIdx.x == 0) { p.trace[blockIdx.x * 6 + 4] = smid(); p.trace[blockIdx.x * 6 + 5] = t_end - (cta * p.q_tiles + min(cta, p.r_tiles)); }
#endif
  if (VEC && p.next_y != nullptr && threadIdx.x == 0) {
    unsigned int t2 = cta * p.q_tiles + min(cta, p.r_tiles);
    const unsigned int image_bytes = static_cast<unsigned int>(p.n) * 4u;
    while (t2 < t_end) {
      const unsigned int image = t2 / p.tpi;
      const unsigned int chunk0 = t2 - image * p.tpi;
      const unsigned int ntiles = min((image + 1u) * p.tpi, t_end) - t2;
      t2 += ntiles;
      const unsigned int off = chunk0 * kTileBytes;
      const unsigned int bytes = min(ntiles * kTileBytes, image_bytes - off);
      const char* addr = reinterpret_cast<const char*>(p.next_y + static_cast<uint64_t>(image) * p.next_y_bs) + off;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(addr), "r"(bytes) : "memory");
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

template <bool NEED_LIK, bool NEED_IDX, bool NOISE, bool VEC, bool FAST, int MINB, bool EXCH = false>
static cudaError_t launch_one(GcParams& p, cudaStream_t st) {
  static int occ = 0;
  auto kernel = gc_fwd_kernel<NEED_LIK, NEED_IDX, NOISE, VEC, FAST, MINB, EXCH>;
  const int64_t total = p.tiles_per_image * p.B;
  const int waves = gc_tuning().ctas_per_sm;   // 0: one tile per CTA; k>0: k CTAs per SM; <0: resident count
  int64_t grid = total;
  if (waves > 0) grid = static_cast<int64_t>(waves) * sm_count();
  else if (waves < 0) {
    // one resident wave; long launches (>= 16 tiles per CTA) are cut into two waves of half-length
    // ranges instead, so that the hardware scheduler evens out the spread in CTA completion times
    // (measured: 61.5 -> 57.3 us on 24 x 491520; shorter launches lose more to the second prologue)
    grid = static_cast<int64_t>(resident_ctas(kernel, &occ)) * sm_count();
    if (total >= 16 * grid) grid *= 2;
    else if (total > 3LL * sm_count() && total <= 2 * grid && gc_tuning().balance) {
      // Between three tiles per SM and two waves of tiles (an 8-GPU shard of a slice: 768 tiles on 740 slots for 8 Kodak
      // images, 512 for 32 training patches): two tiles per CTA.  EVERY CTA then has both of its tiles in flight together
      // — instead of a full wave of one-tile CTAs plus, above one wave, a few stragglers that run a second tile alone —
      // and the CTA slots left free host the next launch of an overlapping chain.  Measured with several batches in
      // flight: 8 x 98304 19.8 -> 15.7 us per step, 32 x 16384 (noise) 18.5 -> 16.2; below three tiles per SM the halved
      // CTA count costs more parallelism than it buys (3 x 98304: 10.5 -> 11.1), and mid-size launches lose
      // (24 x 98304 as 576 x 4 tiles: 71 -> 80 us single chain), so the rule covers exactly this band.
      grid = (total + 1) / 2;
    }
  }
  if (grid > total) grid = total;
  // the next-launch L2 prefetch pays on short launches only (measured: 24 x 98304, 3.1 tiles per CTA, 14.1 -> 13.2 us;
  // 256 x 16384 at 5.5 tiles per CTA 21.9 -> 22.0; 64 x 98304 24.2 -> 24.5; whole y 58.2 -> 60.4: a long launch
  // has no idle HBM time to fill and the prefetch only competes with its own streams)
  if (total > 4 * grid) p.next_y = nullptr;     // (issuing it before the CTA's last tile pair instead of after its range: 14.0 vs 13.7 us)
  p.tpi = static_cast<unsigned int>(p.tiles_per_image);
  p.total_tiles = static_cast<unsigned int>(total);
  p.q_tiles = static_cast<unsigned int>(total / grid);
  p.r_tiles = static_cast<unsigned int>(total % grid);
  // per-thread byte offsets inside a CTA's tile range are 32-bit
  if (static_cast<int64_t>(p.q_tiles) + 1 >= (1LL << 32) / (kThreads * 16)) return cudaErrorInvalidValue;
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
#ifdef RESLIC_TRACE
  p.trace = g_trace_next;
#endif
  return cudaLaunchKernelEx(&cfg, kernel, p);
}
template <bool NEED_LIK, bool NEED_IDX, bool NOISE, bool VEC>
static cudaError_t launch_math(GcParams& p, bool fast, cudaStream_t st) {
  if (p.ex.world > 0) {   // the collecting launch of a multi-GPU step (host checked: NEED_LIK and the 128-bit path)
    constexpr bool kEx = NEED_LIK && VEC;
    if (!fast) return launch_one<NEED_LIK, NEED_IDX, NOISE, VEC, false, 4, kEx>(p, st);
    const bool short_ex = p.tiles_per_image * p.B <= 8LL * 4 * sm_count() && gc_tuning().min_ctas != 4;
    if (short_ex || gc_tuning().min_ctas == 5) return launch_one<NEED_LIK, NEED_IDX, NOISE, VEC, true, VEC ? 5 : 4, kEx>(p, st);
    return launch_one<NEED_LIK, NEED_IDX, NOISE, VEC, true, 4, kEx>(p, st);
  }
  // the math policy only matters when a likelihood is computed
  if (NEED_LIK && !fast) return launch_one<NEED_LIK, NEED_IDX, NOISE, VEC, false, 4>(p, st);
  // short launches (at most 8 tiles per CTA slot of the 4-per-SM build) take the 5-per-SM build
  const bool short_launch = VEC && p.tiles_per_image * p.B <= 8LL * 4 * sm_count() && gc_tuning().min_ctas != 4;
  if (short_launch || gc_tuning().min_ctas == 5) return launch_one<NEED_LIK, NEED_IDX, NOISE, VEC, true, VEC ? 5 : 4>(p, st);
  return launch_one<NEED_LIK, NEED_IDX, NOISE, VEC, true, 4>(p, st);
}
template <bool NEED_LIK, bool NEED_IDX, bool NOISE>
static cudaError_t launch_vec(GcParams& p, bool vec, bool fast, cudaStream_t st) {
  return vec ? launch_math<NEED_LIK, NEED_IDX, NOISE, true>(p, fast, st)
             : launch_math<NEED_LIK, NEED_IDX, NOISE, false>(p, fast, st);
}
template <bool NEED_LIK, bool NEED_IDX>
static cudaError_t launch_noise(GcParams& p, bool vec, bool noise, bool fast, cudaStream_t st) {
  return noise ? launch_vec<NEED_LIK, NEED_IDX, true>(p, vec, fast, st)
               : launch_vec<NEED_LIK, NEED_IDX, false>(p, vec, fast, st);
}

int gc_fwd_launch(const reslic_gc_desc* d, cudaStream_t st) {
  if (!d) return set_error(RESLIC_ERR_ARG, "gc_fwd: null descriptor");
  if (d->struct_size != sizeof(reslic_gc_desc))
    return set_error(RESLIC_ERR_ARG, "gc_fwd: struct_size != sizeof(reslic_gc_desc) (binding built against another ABI revision)");
  if (d->B < 0 || d->n < 0) return set_error(RESLIC_ERR_ARG, "gc_fwd: negative size");
  if (d->B == 0 || d->n == 0) return RESLIC_OK;  // empty input: nothing to do
  if (d->B > (1 << 24)) return set_error(RESLIC_ERR_ARG, "gc_fwd: B too large");
  if (d->mode != RESLIC_Q_DEQUANTIZE && d->mode != RESLIC_Q_NOISE)
    return set_error(RESLIC_ERR_ARG, "gc_fwd: invalid quantization mode");
  const bool want_rate = rate_requested(d->bits, d->bits_accumulate);
  const bool need_lik = d->lik || want_rate;
  const bool need_idx = d->idx != nullptr;
  const bool need_y = d->yhat || d->ste || d->sym || need_lik;
  if (need_y && !d->y) return set_error(RESLIC_ERR_ARG, "gc_fwd: y is null");
  if ((need_lik || need_idx) && !d->sigma) return set_error(RESLIC_ERR_ARG, "gc_fwd: sigma is null");
  if (!need_y && !need_idx) return set_error(RESLIC_ERR_ARG, "gc_fwd: no output requested");
  if (need_idx && (!d->scale_table || d->table_len < 1 || d->table_len > 256))
    return set_error(RESLIC_ERR_ARG, "gc_fwd: scale_table missing or table_len outside 1..256");
  if (!(d->scale_bound > 0.0f)) return set_error(RESLIC_ERR_ARG, "gc_fwd: scale_bound must be > 0");

  GcParams p{};
  bool stride_ok = true;
  auto bs32 = [&](const void* ptr, int64_t bs) -> uint32_t {
    if (ptr && d->B > 1 && (bs < 0 || bs >= (1LL << 32))) stride_ok = false;
    return static_cast<uint32_t>(bs);
  };
  p.y = d->y; p.y_bs = bs32(d->y, d->y_bs);
  p.mu = d->mu; p.mu_bs = bs32(d->mu, d->mu_bs); p.sigma = d->sigma; p.sigma_bs = bs32(d->sigma, d->sigma_bs);
  p.noise = d->noise; p.noise_bs = bs32(d->noise, d->noise_bs);
  p.next_y = d->next_y; p.next_y_bs = bs32(d->next_y, d->next_y_bs);
  p.yhat = d->yhat; p.ste = d->ste; p.lik = d->lik; p.sym = d->sym; p.idx = d->idx;
  p.yhat_bs = bs32(d->yhat, d->yhat_bs); p.ste_bs = bs32(d->ste, d->ste_bs); p.lik_bs = bs32(d->lik, d->lik_bs);
  p.sym_bs = bs32(d->sym, d->sym_bs); p.idx_bs = bs32(d->idx, d->idx_bs);
  if (!stride_ok) return set_error(RESLIC_ERR_ARG, "gc_fwd: batch stride outside [0, 2^32) elements");
  p.table = d->scale_table; p.table_len = d->table_len;
  p.n = d->n; p.B = d->B; p.scale_bound = d->scale_bound;
  // compressai applies the likelihood bound only when it is > 0; max(L, -inf) is the branch-free "no bound"
  p.lik_floor = d->likelihood_bound > 0.0f ? d->likelihood_bound : -std::numeric_limits<float>::infinity();
  p.seed_lo = static_cast<uint32_t>(d->philox_seed); p.seed_hi = static_cast<uint32_t>(d->philox_seed >> 32);
  p.off_lo = static_cast<uint32_t>(d->philox_offset); p.off_hi = static_cast<uint32_t>(d->philox_offset >> 32);

  bool vec = (d->n % 4 == 0);
  auto chk = [&](const void* ptr, int64_t bs) {
    if (ptr && (!aligned16(ptr) || (bs % 4) != 0)) vec = false;
  };
  chk(d->y, d->y_bs); chk(d->mu, d->mu_bs); chk(d->sigma, d->sigma_bs); chk(d->noise, d->noise_bs);
  chk(d->yhat, d->yhat_bs); chk(d->ste, d->ste_bs); chk(d->lik, d->lik_bs); chk(d->sym, d->sym_bs);
  chk(d->idx, d->idx_bs);
  if (d->next_y && (!aligned16(d->next_y) || (d->next_y_bs % 4) != 0)) p.next_y = nullptr;   // a hint: dropped, never an error

  const int64_t groups = vec ? d->n / 4 : d->n;
  p.tiles_per_image = (groups + kThreads - 1) / kThreads;
  if (d->n >= (1LL << 31)) return set_error(RESLIC_ERR_ARG, "gc_fwd: more than 2^31 elements per image");
  if (p.tiles_per_image * d->B >= (1LL << 31)) return set_error(RESLIC_ERR_ARG, "gc_fwd: input too large (>= 2^31 tiles)");
  if (want_rate) {
    const int rc = rate_setup("gc_fwd", d->bits, d->bits_accumulate, d->workspace, d->workspace_bytes, d->B,
                              &p.bits, &p.bits_accumulate, &p.workspace);
    if (rc != RESLIC_OK) return rc;
  }
  if (d->exchange) {
    const reslic_rate_exchange* x = d->exchange;
    if (x->struct_size != sizeof(reslic_rate_exchange))
      return set_error(RESLIC_ERR_ARG, "gc_fwd: exchange->struct_size != sizeof(reslic_rate_exchange) (ABI mismatch)");
    if (!want_rate || (d->bits_accumulate != 0 && d->bits_accumulate != RESLIC_RATE_COLLECT))
      return set_error(RESLIC_ERR_ARG, "gc_fwd: exchange needs a rate output in mode 0 or RESLIC_RATE_COLLECT");
    if (d->B >= 65536) return set_error(RESLIC_ERR_UNSUPPORTED, "gc_fwd: exchange supports fewer than 65536 images per launch");
    if (x->world < 1 || x->world > 64 || x->rank < 0 || x->rank >= x->world || x->ring < 1 || !x->peer_base || x->step < 0)
      return set_error(RESLIC_ERR_ARG, "gc_fwd: exchange: bad world/rank/ring/step or null peer_base");
    p.ex.peer = x->peer_base; p.ex.cursor = x->cursor; p.ex.step_rel = x->step; p.ex.extra = x->extra; p.ex.pixels = x->pixels; p.ex.images = x->images;
    p.ex.world = x->world; p.ex.rank = x->rank; p.ex.ring = x->ring;
  }
  if (p.ex.world > 0 && !vec)
    return set_error(RESLIC_ERR_UNSUPPORTED, "gc_fwd: exchange needs the 128-bit path (16-byte aligned tensors, strides and n multiples of 4)");
  const bool noise = d->mode == RESLIC_Q_NOISE;
  const bool fast = math_mode() != RESLIC_MATH_MIRROR;
  // The clamp-free path needs (2^22 + 1) / scale_bound far from overflow and, in FAST mode, a bound
  // that keeps the product of four likelihoods normal; exotic settings run the general path throughout.
  const bool product_ok = !need_lik || !fast || d->likelihood_bound >= 1e-9f;
  p.d_limit = (d->scale_bound >= 1e-3f && product_ok) ? 4194304.0f : 0.0f;
  cudaError_t err;
  if (need_lik) err = need_idx ? launch_noise<true, true>(p, vec, noise, fast, st)
                               : launch_noise<true, false>(p, vec, noise, fast, st);
  else err = need_idx ? launch_noise<false, true>(p, vec, noise, fast, st)
                      : launch_noise<false, false>(p, vec, noise, fast, st);
  if (err != cudaSuccess) return set_cuda_error(err, "gc_fwd launch");
  return RESLIC_OK;
}

}  // namespace reslic

#ifdef RESLIC_TRACE
extern "C" void reslic_debug_set_trace(unsigned long long* ptr) { reslic::g_trace_next = ptr; }
#endif
