// common.cuh — shared device helpers for the sm_100a entropy-model kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace reslic {

constexpr int kThreads = 256;   // threads per CTA for the streaming kernels
constexpr int kMaxBpi = 512;    // max CTAs cooperating on one image's rate sum

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

// ---- deterministic per-image sum: thread fp32 partials -> fp64 warp/CTA tree -> one fp64
// partial per CTA -> the last CTA of the image adds the partials in index order.
// `counters` are left at zero, so the workspace needs zero-filling only once.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Workspace layout: one row of (kMaxBpi partials + 1 arrival counter) doubles per image.  A
// row's layout does not depend on B, so a counter slot is never reused as a partial by a
// later launch with a different batch size — the "zero once" contract stays valid.
constexpr int kWsRow = kMaxBpi + 1;

// Returns through bits_out[image] = -(sum of all CTAs' acc).  Must be called by all threads.
__device__ __forceinline__ void image_sum_finish(float acc, int image, int chunk, int bpi,
                                                 double* workspace, double* bits_out) {
  double* partials = workspace + static_cast<int64_t>(image) * kWsRow;
  unsigned int* counter = reinterpret_cast<unsigned int*>(partials + kMaxBpi);
  __shared__ double s_warp[kThreads / 32];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double v = warp_sum(static_cast<double>(acc));
  if (lane == 0) s_warp[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = (lane < kThreads / 32) ? s_warp[lane] : 0.0;
    v = warp_sum(v);
    if (lane == 0) {
      if (bpi == 1) {
        bits_out[image] = -v;
        s_last = false;
      } else {
        partials[chunk] = v;
        __threadfence();
        const unsigned int prev = atomicAdd(counter, 1u);
        s_last = (prev == static_cast<unsigned int>(bpi - 1));
      }
    }
  }
  __syncthreads();
  if (s_last && warp == 0) {
    __threadfence();
    const volatile double* pp = partials;
    double t = 0.0;
    for (int i = lane; i < bpi; i += 32) t += pp[i];
    t = warp_sum(t);
    if (lane == 0) {
      bits_out[image] = -t;
      *counter = 0u;
    }
  }
}

}  // namespace reslic
