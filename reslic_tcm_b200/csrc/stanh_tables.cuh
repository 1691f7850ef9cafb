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
// start[] and walks forward over the entries of its own cell only (one comparison on average, exact
// for ANY ascending table — trained, non-uniform thresholds included).
#pragma once
#include "common.cuh"

namespace reslic {

constexpr int kStanhPad = 4;          // NaN entries behind each table: every ordered comparison with them is false
constexpr int kStanhMaxCells = 4096;

struct StanhGrid {
  float lo, inv, gmax;                // cell(x) = RN(clamp((x - lo) * inv, 0, gmax)), NaN -> 0
  const uint16_t* start;              // [gmax + 1]
};

__device__ __forceinline__ int stanh_cell(float x, const StanhGrid& g) {
  float u = (x - g.lo) * g.inv;                       // two roundings, relative: never off by a cell near lo
  u = fminf(fmaxf(u, 0.0f), g.gmax);                  // NaN -> 0
  return __float_as_int(u + 12582912.0f) - 0x4B400000;  // round to nearest in the FMA pipe (no F2I)
}

struct StanhSm {
  const float2* bw;       // [K + pad] (b_k, w_k / 2) ascending in b
  const float2* lowup;    // [K + 1]   half-widths (low, up) of level cell j: (dist[j-1] | 0, dist[j] | 0)
  const float* cw;        // [K + 1]   levels
  const float* avgp;      // [K + 1 + pad]  -inf, average_points[0..K-1], NaN...:  cell j <=> avgp[j] < v <= avgp[j+1]
  StanhGrid gb, ga;       // grids over b and over average_points
  int K;
  float cw0, cwK;
};

__host__ __device__ inline int stanh_cells(int K) { return 4 * K < kStanhMaxCells ? 4 * K : kStanhMaxCells; }
// dynamic shared memory of stage_stanh_sm (bytes)
__host__ __device__ inline size_t stanh_sm_bytes(int K) {
  const size_t G = static_cast<size_t>(stanh_cells(K));
  return static_cast<size_t>(K + kStanhPad) * 8 + static_cast<size_t>(K + 1) * 8 + static_cast<size_t>(K + 1) * 4 +
         static_cast<size_t>(K + 1 + kStanhPad) * 4 + 2 * ((G + 2) & ~size_t(1)) * 2 + 16;
}

// #{k : x > b_k}
__device__ __forceinline__ int stanh_count_gt_b(float x, const StanhSm& T) {
  int c = T.gb.start[stanh_cell(x, T.gb)];
  while (x > T.bw[c].x) ++c;
  return c;
}
// #{k : x >= b_k}
__device__ __forceinline__ int stanh_count_ge_b(float x, const StanhSm& T) {
  int c = T.gb.start[stanh_cell(x, T.gb)];
  while (x >= T.bw[c].x) ++c;
  return c;
}
// #{j : v > average_points[j]}
__device__ __forceinline__ int stanh_count_gt_avg(float v, const StanhSm& T) {
  int c = T.ga.start[stanh_cell(v, T.ga)];
  while (v > T.avgp[c + 1]) ++c;
  return c;
}

// All threads of the CTA must call; ends with a barrier.
__device__ __forceinline__ void stage_stanh_sm(const float* b, const float* w, const float* cum_w, const float* avg,
                                               const float* dist, int K, unsigned char* raw, StanhSm& T) {
  const int G = stanh_cells(K);
  float2* bw = reinterpret_cast<float2*>(raw);
  float2* lowup = bw + (K + kStanhPad);
  float* cw = reinterpret_cast<float*>(lowup + (K + 1));
  float* avgp = cw + (K + 1);
  uint16_t* sb = reinterpret_cast<uint16_t*>(avgp + (K + 1 + kStanhPad));
  uint16_t* sa = sb + ((G + 2) & ~1);
  const float nan = __int_as_float(0x7fc00000);
  for (int i = threadIdx.x; i < K + kStanhPad; i += blockDim.x) {
    bw[i] = (i < K) ? make_float2(b[i], 0.5f * w[i]) : make_float2(nan, 0.0f);
    avgp[i + 1] = (i < K) ? avg[i] : nan;
  }
  for (int i = threadIdx.x; i <= K; i += blockDim.x) {
    cw[i] = cum_w[i];
    lowup[i] = make_float2(i > 0 ? dist[i - 1] : 0.0f, i < K ? dist[i] : 0.0f);
  }
  if (threadIdx.x == 0) avgp[0] = __int_as_float(0xff800000);
  auto grid_of = [&](const float* t, uint16_t* start) {
    StanhGrid g;
    g.lo = t[0];
    const float span = t[K - 1] - t[0];
    float inv = (span > 0.0f) ? static_cast<float>(G) / span : 0.0f;
    if (!(inv <= 3.0e38f)) inv = 0.0f;
    g.inv = inv; g.gmax = static_cast<float>(G); g.start = start;
    return g;
  };
  T.gb = grid_of(b, sb);
  T.ga = grid_of(avg, sa);
  int steps = 1;
  while ((1 << steps) <= K) ++steps;
  for (int g = threadIdx.x; g <= G; g += blockDim.x) {
    int lb = 0, la = 0;
    for (int step = 1 << (steps - 1); step > 0; step >>= 1) {
      const int ib = lb + step, ia = la + step;
      if (ib <= K && stanh_cell(b[ib - 1], T.gb) < g) lb = ib;
      if (ia <= K && stanh_cell(avg[ia - 1], T.ga) < g) la = ia;
    }
    sb[g] = static_cast<uint16_t>(lb);
    sa[g] = static_cast<uint16_t>(la);
  }
  T.bw = bw; T.lowup = lowup; T.cw = cw; T.avgp = avgp; T.K = K;
  T.cw0 = cum_w[0]; T.cwK = cum_w[K];
  __syncthreads();
}

}  // namespace reslic
