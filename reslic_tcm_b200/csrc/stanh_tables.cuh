// stanh_tables.cuh — shared-memory form of the STanH tables for the 128-bit kernels.
//
// The quantizer and the likelihood both start with "how many sorted table entries lie below x"
// (thresholds b for the level / the saturation window, level mid-points for the likelihood cell;
// src/quantization/activation.py:135-150, adaptive_gaussian_conditional.py:495-537).  A binary search
// is 8-10 dependent shared loads per lookup and three lookups per element; here each table gets a
// uniform grid of G = 4K cells over its range with, per cell g,
//     start[g] = #{k : cell(t_k) < g}
// computed with the very cell() function the lookup uses.  cell() is monotone in x, so every entry
// counted by start[cell(x)] is strictly below x whatever the rounding inside cell(): the lookup reads
// start[] and only has to look at the entries of its own cell (0 or 1 of them for near-uniform
// thresholds; exact for ANY ascending table — trained, non-uniform thresholds included).
//
// The layout is a compile-time function of KMAX (256 covers the reference default extrema = 80 -> K = 160;
// 1024 is the ABI limit), so every table access is one LDS with an immediate offset.
#pragma once
#include "common.cuh"

namespace reslic {

constexpr int kStanhPad = 4;          // NaN entries behind each table: every ordered comparison with them is false
extern __shared__ __align__(16) unsigned char stanh_smem[];

template <int KMAX>
struct StanhSm {
  static constexpr int kCells = 4 * KMAX;                       // grid cells (at most; 4K are used)
  static constexpr int oBw = 0;                                 // float2 [KMAX + pad]   (b_k, w_k / 2)
  static constexpr int oLowUp = oBw + (KMAX + kStanhPad) * 8;   // float2 [KMAX + 1]     half-widths (low, up) of level cell j
  static constexpr int oCw = oLowUp + (KMAX + 1) * 8;           // float  [KMAX + 1]     levels
  static constexpr int oAvg = oCw + (KMAX + 1) * 4;             // float  [KMAX + 1 + pad]  -inf, average_points, NaN...
  static constexpr int oSb = oAvg + (KMAX + 1 + kStanhPad) * 4; // uint16 [kCells + 2]
  static constexpr int oSa = oSb + (kCells + 2) * 2;            // uint16 [kCells + 2]
  static constexpr int kBytes = (oSa + (kCells + 2) * 2 + 15) & ~15;

  float b_lo, b_inv, a_lo, a_inv, gmax;   // cell(x) = RN(clamp((x - lo) * inv, 0, gmax)), NaN -> 0
  float cw0, cwK;
  int K;

  // the tables sit at the start of the CTA's dynamic shared memory; naming the array (instead of carrying a
  // pointer) lets every access be one LDS [index * size + constant]
  __device__ __forceinline__ const float2* bw() const { return reinterpret_cast<const float2*>(stanh_smem + oBw); }
  __device__ __forceinline__ const float2* lowup() const { return reinterpret_cast<const float2*>(stanh_smem + oLowUp); }
  __device__ __forceinline__ const float* cw() const { return reinterpret_cast<const float*>(stanh_smem + oCw); }
  __device__ __forceinline__ const float* avgp() const { return reinterpret_cast<const float*>(stanh_smem + oAvg); }
  __device__ __forceinline__ const uint16_t* sb() const { return reinterpret_cast<const uint16_t*>(stanh_smem + oSb); }
  __device__ __forceinline__ const uint16_t* sa() const { return reinterpret_cast<const uint16_t*>(stanh_smem + oSa); }
};

__device__ __forceinline__ int stanh_cell(float x, float lo, float inv, float gmax) {
  float u = (x - lo) * inv;                             // two roundings, relative: never off by a cell near lo
  u = fminf(fmaxf(u, 0.0f), gmax);                      // NaN -> 0
  return __float_as_int(u + 12582912.0f) - 0x4B400000;  // round to nearest in the FMA pipe (no F2I)
}

// #{k : x > b_k}
template <int KMAX>
__device__ __forceinline__ int stanh_count_gt_b(float x, const StanhSm<KMAX>& T) {
  int c = T.sb()[stanh_cell(x, T.b_lo, T.b_inv, T.gmax)];
  const float2* bw = T.bw();
  const float t0 = bw[c].x, t1 = bw[c + 1].x;           // the cell rarely holds more than one entry
  c += (x > t0) ? 1 : 0;
  c += (x > t1) ? 1 : 0;
  if (x > t1) while (x > bw[c].x) ++c;
  return c;
}
// #{k : x >= b_k}
template <int KMAX>
__device__ __forceinline__ int stanh_count_ge_b(float x, const StanhSm<KMAX>& T) {
  int c = T.sb()[stanh_cell(x, T.b_lo, T.b_inv, T.gmax)];
  const float2* bw = T.bw();
  const float t0 = bw[c].x, t1 = bw[c + 1].x;
  c += (x >= t0) ? 1 : 0;
  c += (x >= t1) ? 1 : 0;
  if (x >= t1) while (x >= bw[c].x) ++c;
  return c;
}
// #{j : v > average_points[j]}
template <int KMAX>
__device__ __forceinline__ int stanh_count_gt_avg(float v, const StanhSm<KMAX>& T) {
  int c = T.sa()[stanh_cell(v, T.a_lo, T.a_inv, T.gmax)];
  const float* ap = T.avgp() + 1;
  const float t0 = ap[c], t1 = ap[c + 1];
  c += (v > t0) ? 1 : 0;
  c += (v > t1) ? 1 : 0;
  if (v > t1) while (v > ap[c]) ++c;
  return c;
}

// All threads of the CTA must call; ends with a barrier.  K <= KMAX.
template <int KMAX>
__device__ __forceinline__ void stage_stanh_sm(const float* b, const float* w, const float* cum_w, const float* avg,
                                               const float* dist, int K, StanhSm<KMAX>& T) {
  using S = StanhSm<KMAX>;
  unsigned char* raw = stanh_smem;
  const int G = 4 * K;
  float2* bw = reinterpret_cast<float2*>(raw + S::oBw);
  float2* lowup = reinterpret_cast<float2*>(raw + S::oLowUp);
  float* cw = reinterpret_cast<float*>(raw + S::oCw);
  float* avgp = reinterpret_cast<float*>(raw + S::oAvg);
  uint16_t* sb = reinterpret_cast<uint16_t*>(raw + S::oSb);
  uint16_t* sa = reinterpret_cast<uint16_t*>(raw + S::oSa);
  const float nan = __int_as_float(0x7fc00000);
  T.K = K; T.gmax = static_cast<float>(G);
  auto inv_of = [&](float span) {
    float inv = (span > 0.0f) ? static_cast<float>(G) / span : 0.0f;
    return (inv <= 3.0e38f) ? inv : 0.0f;
  };
  T.b_lo = b[0]; T.b_inv = inv_of(b[K - 1] - b[0]);
  T.a_lo = avg[0]; T.a_inv = inv_of(avg[K - 1] - avg[0]);
  T.cw0 = cum_w[0]; T.cwK = cum_w[K];
  // (zero first: a table that is not ascending — negative trained weights — then still yields in-range starts)
  for (int g = threadIdx.x; g <= G + 1; g += blockDim.x) { sb[g] = 0; sa[g] = 0; }
  __syncthreads();
  // start[g] = #{k : cell(t_k) < g}: entry k is the first one NOT counted for every g in (cell(t_{k-1}), cell(t_k)]
  // (monotone cells of an ascending table), and everything beyond cell(t_{K-1}) counts all K
  for (int k = threadIdx.x; k < K + kStanhPad; k += blockDim.x) {
    if (k < K) {
      const float bk = b[k], ak = avg[k];
      bw[k] = make_float2(bk, 0.5f * w[k]);
      avgp[k + 1] = ak;
      const int cb1 = stanh_cell(bk, T.b_lo, T.b_inv, T.gmax), ca1 = stanh_cell(ak, T.a_lo, T.a_inv, T.gmax);
      const int cb0 = k > 0 ? stanh_cell(b[k - 1], T.b_lo, T.b_inv, T.gmax) : -1;
      const int ca0 = k > 0 ? stanh_cell(avg[k - 1], T.a_lo, T.a_inv, T.gmax) : -1;
      for (int g = cb0 + 1; g <= cb1; ++g) sb[g] = static_cast<uint16_t>(k);
      for (int g = ca0 + 1; g <= ca1; ++g) sa[g] = static_cast<uint16_t>(k);
      if (k == K - 1) {
        for (int g = cb1 + 1; g <= G + 1; ++g) sb[g] = static_cast<uint16_t>(K);
        for (int g = ca1 + 1; g <= G + 1; ++g) sa[g] = static_cast<uint16_t>(K);
      }
    } else {
      bw[k] = make_float2(nan, 0.0f);
      avgp[k + 1] = nan;
    }
  }
  for (int i = threadIdx.x; i <= K; i += blockDim.x) {
    cw[i] = cum_w[i];
    lowup[i] = make_float2(i > 0 ? dist[i - 1] : 0.0f, i < K ? dist[i] : 0.0f);
  }
  if (threadIdx.x == 0) avgp[0] = __int_as_float(0xff800000);
  __syncthreads();
}

}  // namespace reslic
