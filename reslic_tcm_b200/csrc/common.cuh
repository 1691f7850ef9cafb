// common.cuh — shared device helpers for the sm_100a entropy-model kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace reslic {

constexpr int kThreads = 256;   // threads per CTA for the streaming kernels

// ---- streaming global access (every byte is touched once: keep it out of L1, evict-first in L2)
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  return __ldcs(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ float ld_stream1(const float* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
  __stcs(reinterpret_cast<float4*>(p), v);
}
__device__ __forceinline__ void st_stream4(int32_t* p, int4 v) {
  __stcs(reinterpret_cast<int4*>(p), v);
}
__device__ __forceinline__ void st_stream1(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream1(int32_t* p, int32_t v) { __stcs(p, v); }

// ---- programmatic dependent launch (PDL) controls; no-ops when launched without the attribute
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- NaN-propagating max/min (torch.max(x, bound) propagates NaN; fmaxf does not)
__device__ __forceinline__ float max_nan(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float min_nan(float a, float b) {
  float r;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// ---- Philox4x32-10 (Salmon et al. 2011), counter = 128 bit, key = 64 bit.
struct Philox4 {
  uint32_t x, y, z, w;
};
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return {c0, c1, c2, c3};
}
// 24 random bits -> uniform in the open interval (-1/2, 1/2), symmetric about 0.
__device__ __forceinline__ float u32_to_centered_uniform(uint32_t r) {
  return fmaf(static_cast<float>(r >> 8), 5.9604644775390625e-08f, -0.5f + 2.98023223876953125e-08f);
}

// ---- deterministic per-image rate sum, independent of how tiles are assigned to CTAs.
// Each warp reduces its fp32 partial with a fixed shuffle tree and commits it with ONE 64-bit
// integer atomicAdd to the image's word:
//     [63:48] number of warps that have committed      [47:0] sum of (bits * 2^16 + 2^30)
// Integer adds commute, so the total does not depend on arrival order (bit-reproducible), and
// because the arrival count travels in the same word as the sum, the warp whose add brings the
// count to `expected` holds the complete sum in the atomic's return value: it writes bits[image]
// and re-zeroes the word.  No fence, no second atomic and no grid-wide barrier sit on the
// kernel's critical path.  Non-finite partials (NaN likelihood, L == 0 with the bound disabled)
// are carried by a per-image flag word (rare path, fenced).
//
// Workspace layout (64-bit words): [0, B) sum/arrival words, [B, 2B) flags; all words are zero
// between launches, so launches with different B may share one zero-initialised workspace.
constexpr unsigned long long kArrOne = 1ull << 48;
constexpr long long kRateBias = 1ll << 30;

__device__ __forceinline__ float warp_sum_f32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Split form: `rate_commit_issue` performs the warp reduction and lane 0's atomicAdd and returns
// the post-add word (meaningful in lane 0 only); `rate_commit_finish` inspects it.  Calling finish
// one tile later keeps the atomic's round trip off the warp's critical path.
__device__ __forceinline__ unsigned long long rate_commit_issue(float acc, int image, int64_t B,
                                                               unsigned long long* ws) {
  const float v = warp_sum_f32(acc);
  unsigned long long now = 0ull;
  if ((threadIdx.x & 31) == 0) {
    const bool finite = fabsf(v) <= 1e30f;                       // false for NaN
    long long q = 0;
    if (finite) q = __float2ll_rn(-v * 65536.0f);
    else {
      atomicOr(&ws[B + image], (v != v) ? 1ull : 2ull);
      __threadfence();
    }
    const unsigned long long add = kArrOne + static_cast<unsigned long long>(q + kRateBias);
    now = atomicAdd(&ws[image], add) + add;
  }
  return now;
}

// ---- multi-GPU rate exchange (reslic_rate_exchange): kernel-side copy of the descriptor; world == 0 = none
struct RateEx {
  void* const* peer;              // DEVICE array [world] of exchange-buffer bases as mapped in this process
  const unsigned long long* cursor;   // DEVICE word (nullable): base of the step numbers, advanced by the caller
  long long step_rel;             // this batch's number relative to *cursor
  const double* extra;            // nullable
  double pixels, images;
  int world, rank, ring;
};
__device__ __forceinline__ double rate_fixed_to_bits(long long sum, unsigned long long flag) {
  double bits = static_cast<double>(sum) * (1.0 / 65536.0);
  if (flag & 1ull) bits = __longlong_as_double(0x7ff8000000000000LL);
  else if (flag & 2ull) bits = __longlong_as_double(0x7ff0000000000000LL);
  return bits;
}
// What one lane has won so far: the fixed-point rates of the images whose LAST arriver it was.  Kept in registers across
// the kernel's tile loop and handed to rate_publish once, after the loop, where nothing else is live (with the publish
// inside the loop the 48-register build of the Gaussian-conditional kernel spills three times as much).
struct RateWin { long long sum; unsigned long long flag; unsigned int count; };

// Called after the tile loop by every lane that completed at least one image.  One 64-bit atomicAdd per such lane on
// workspace word [4B] — [63:48] images done, [47:0] sum of (image rate + 2^20) in 48.16 fixed point, so the lane whose
// add completes the count holds the whole batch total in the atomic's return value (no fence, no second atomic; integer
// adds, so the total is bit-reproducible); non-finite images go through the flag word [4B+2] (rare, fenced).  That lane
// publishes the row {bits, extra, pixels, images} into slot ((*cursor + step) % ring, rank) of EVERY rank's exchange
// buffer as four self-validating 16-byte cells {value, step + 1}: a 16-byte store is one transaction, so a reader that
// sees the tag sees the value — no fence, no flag word, nothing to wait for: the lane issues 4 * world peer stores over
// NVLink and retires.  Everything the publisher needs from memory (cursor, extra, flags) is loaded BEFORE the atomic, so
// the launch ends one L2 round trip after its last image completes.  (B < 65536 is checked on the host; 2^20 per image
// keeps a rounding-negative rate from borrowing out of the sum field.)
constexpr unsigned long long kBatchBias = 1ull << 20;
__device__ __forceinline__ void st_cell(void* p, double v, unsigned long long tag) {
  asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(__double_as_longlong(v)), "l"(tag) : "memory");
}
__device__ __forceinline__ void rate_publish(const RateWin win, int64_t B, unsigned long long* ws, const RateEx& ex) {
  unsigned long long* bw = ws + 4 * B;
  if (win.flag) { atomicOr(&bw[2], win.flag); __threadfence(); }
  const unsigned long long base = ex.cursor ? *reinterpret_cast<const volatile unsigned long long*>(ex.cursor) : 0ull;
  const double extra = ex.extra ? *reinterpret_cast<const volatile double*>(ex.extra) : 0.0;
  const unsigned long long add = (static_cast<unsigned long long>(win.count) << 48) +
                                 static_cast<unsigned long long>(win.sum + static_cast<long long>(win.count * kBatchBias));
  const unsigned long long now = atomicAdd(&bw[0], add) + add;
  if ((now >> 48) != static_cast<unsigned long long>(B)) return;
  const long long total = static_cast<long long>(now & (kArrOne - 1ull)) - static_cast<long long>(B) * static_cast<long long>(kBatchBias);
  // (a flagged image fenced its atomicOr before its own add, and this lane's add came last: the flags are complete)
  const unsigned long long f = *reinterpret_cast<volatile unsigned long long*>(&bw[2]);
  bw[0] = 0ull;
  if (f) bw[2] = 0ull;
  const unsigned long long step = static_cast<unsigned long long>(ex.step_rel) + base;
  const size_t cell = (static_cast<size_t>(step % static_cast<unsigned long long>(ex.ring)) * ex.world + ex.rank) * 64;   // bytes
  const double bits = rate_fixed_to_bits(total, f);
  for (int p = 0; p < ex.world; ++p) {
    char* row = static_cast<char*>(ex.peer[p]) + cell;
    st_cell(row, bits, step + 1ull);
    st_cell(row + 16, extra, step + 1ull);
    st_cell(row + 32, ex.pixels, step + 1ull);
    st_cell(row + 48, ex.images, step + 1ull);
  }
}

// (`win`: where the winning lane notes what it completed, for rate_publish after the loop; nullptr = no exchange)
__device__ __forceinline__ void rate_commit_finish(unsigned long long now, int image, unsigned int expected,
                                                   int64_t B, unsigned long long* ws, double* bits_out,
                                                   bool accumulate, bool collect = false, const RateEx* ex = nullptr) {
  if ((threadIdx.x & 31) == 0 && (now >> 48) == expected) {
    __threadfence();
    long long sum = static_cast<long long>(now & (kArrOne - 1ull)) - static_cast<long long>(expected) * kRateBias;
    unsigned long long flag = *reinterpret_cast<volatile unsigned long long*>(&ws[B + image]);
    if (collect) {   // fold in what earlier RESLIC_RATE_DEFERRED launches of this stream left behind
      sum += static_cast<long long>(*reinterpret_cast<volatile unsigned long long*>(&ws[2 * B + image]));
      const unsigned long long dflag = *reinterpret_cast<volatile unsigned long long*>(&ws[3 * B + image]);
      ws[2 * B + image] = 0ull;
      if (dflag) ws[3 * B + image] = 0ull;
      flag |= dflag;
    }
    const double bits = rate_fixed_to_bits(sum, flag);
    bits_out[image] = accumulate ? bits_out[image] + bits : bits;   // single writer per image
    ws[image] = 0ull;
    if (flag) ws[B + image] = 0ull;
    if (ex != nullptr) rate_publish(RateWin{sum, flag, 1u}, B, ws, *ex);
  }
}

// Deferred form (mode RESLIC_RATE_DEFERRED): the warp's sum goes to the image's DEFERRED word with a
// fire-and-forget reduction (no return value, so the warp retires without waiting for the round
// trip to L2, which is what the immediate form pays to learn whether it was the last arriver);
// reslic_rate_finalize_f64 later turns the words into bits and re-zeroes them.  Integer addition
// commutes, so the result is still bit-reproducible and may accumulate over any number of launches.
// Workspace layout (64-bit words): [0,B) immediate sum/arrival, [B,2B) immediate flags,
// [2B,3B) deferred sums (signed fixed point, bits * 2^16), [3B,4B) deferred flags,
// [4B,4B+4) batch total + images done / spare / batch flags / spare (rate_publish).
__device__ __forceinline__ void rate_defer(float acc, int image, int64_t B, unsigned long long* ws) {
  const float v = warp_sum_f32(acc);
  if ((threadIdx.x & 31) == 0) {
    if (fabsf(v) <= 1e30f) atomicAdd(&ws[2 * B + image], static_cast<unsigned long long>(__float2ll_rn(-v * 65536.0f)));
    else atomicOr(&ws[3 * B + image], (v != v) ? 1ull : 2ull);
  }
}

// All 32 lanes of a warp must call.  acc = this warp's sum of log2(L) over elements of `image`;
// `expected` = number of warps (over the whole grid) that commit to `image`;
// mode = the descriptor's bits_accumulate (0 write, 1 accumulate, 2 deferred, 3 write + collect deferred).
__device__ __forceinline__ void rate_commit(float acc, int image, unsigned int expected, int64_t B,
                                            unsigned long long* ws, double* bits_out, int mode,
                                            const RateEx* ex = nullptr) {
  if (mode == 2) { rate_defer(acc, image, B, ws); return; }
  const unsigned long long now = rate_commit_issue(acc, image, B, ws);
  rate_commit_finish(now, image, expected, B, ws, bits_out, mode == 1, mode == 3, ex);
}

}  // namespace reslic
