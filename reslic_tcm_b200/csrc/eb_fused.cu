// eb_fused.cu — fused factorized EntropyBottleneck forward on the hyper-latent z (sm_100a).
//
// Replaces compressai EntropyBottleneck.forward (reference call site
// src/models/reference/tcm.py:429): the two permute+contiguous copies of the wrapper
// (src/entropy_models/adaptive_entropy_bottleneck.py:679-708), quantize about the medians,
// two passes of the 5-layer per-channel cumulative-logit MLP (:525-543), the sign-trick
// sigmoid difference (:658-666), the 1e-9 LowerBound, ste_round(z - med) + med
// (tcm.py:431-433), the compress-path symbols (tcm.py:507) and the per-image rate sum
// (training/loss.py:24-27) — ~75 launches and 3 HBM round trips in the reference.
//
// z stays in its native [B, C, h, w] layout: channel c of image b is a contiguous run of
// hw floats, so the "permute" is just index arithmetic.  Each CTA walks 256-element tiles
// of one image; the (softplus / tanh transformed) parameters of the few channels a tile
// touches are staged in shared memory once per tile.  The kernel is FP32/MUFU-issue bound
// (24 tanh + 2 sigmoid per element), not HBM bound; z is 3.75 % of y's elements.
#include "common.cuh"
#include "reslic_internal.h"

namespace reslic {

constexpr int kEbMaxCh = 64;     // channels whose parameters fit the per-tile staging area
constexpr int kEbStride = 59;    // 33 matrix + 13 bias + 12 factor + 1 median (odd: bank-conflict free)

struct EbParams {
  const float* z; const float* noise; int64_t z_bs, noise_bs;
  const float* matrix[5]; const float* bias[5]; const float* factor[4]; const float* medians;
  float* zhat; float* ste; float* lik; int32_t* sym;
  int64_t zhat_bs, ste_bs, lik_bs, sym_bs;
  double* bits; unsigned long long* workspace; int64_t B;
  int64_t ne;       // elements per image = C*hw
  int hw, C, tile, bpi, noise_mode;
  float lik_bound;
  uint32_t seed_lo, seed_hi, off_lo, off_hi;
};

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

__global__ void __launch_bounds__(kThreads) eb_fwd_kernel(const EbParams p) {
  __shared__ float s_par[kEbMaxCh * kEbStride];
  const int image = blockIdx.x / p.bpi;
  const int chunk = blockIdx.x - image * p.bpi;
  const float* __restrict__ z = p.z + image * p.z_bs;
  const float* __restrict__ nz = (p.noise_mode && p.noise) ? p.noise + image * p.noise_bs : nullptr;
  float* zhat = p.zhat ? p.zhat + image * p.zhat_bs : nullptr;
  float* ste = p.ste ? p.ste + image * p.ste_bs : nullptr;
  float* lik = p.lik ? p.lik + image * p.lik_bs : nullptr;
  int32_t* sym = p.sym ? p.sym + image * p.sym_bs : nullptr;
  const bool need_lik = p.lik || p.bits;

  float acc = 0.0f;
  const int64_t ntiles = (p.ne + p.tile - 1) / p.tile;
  for (int64_t t = chunk; t < ntiles; t += p.bpi) {
    const int64_t e0 = t * p.tile;
    const int64_t e1 = min(p.ne, e0 + static_cast<int64_t>(p.tile));
    const int c_lo = static_cast<int>(e0 / p.hw);
    const int c_hi = static_cast<int>((e1 - 1) / p.hw);
    const int nch = c_hi - c_lo + 1;
    __syncthreads();  // previous tile's readers are done with s_par
    for (int i = threadIdx.x; i < nch * kEbStride; i += kThreads) {
      const int cl = i / kEbStride, j = i - cl * kEbStride, c = c_lo + cl;
      float v;
      if (j < oB0) v = softplus_ref(p.matrix[0][c * 3 + j]);
      else if (j < oF0) v = p.bias[0][c * 3 + (j - oB0)];
      else if (j < oM1) v = tanhf(p.factor[0][c * 3 + (j - oF0)]);
      else if (j < oM4) {
        const int l = (j - oM1) / 15, r = (j - oM1) - l * 15;
        if (r < 9) v = softplus_ref(p.matrix[1 + l][c * 9 + r]);
        else if (r < 12) v = p.bias[1 + l][c * 3 + (r - 9)];
        else v = tanhf(p.factor[1 + l][c * 3 + (r - 12)]);
      } else if (j < oB4) v = softplus_ref(p.matrix[4][c * 3 + (j - oM4)]);
      else if (j == oB4) v = p.bias[4][c];
      else v = p.medians[c];
      s_par[i] = v;
    }
    __syncthreads();
    const int64_t e = e0 + threadIdx.x;
    if (threadIdx.x < p.tile && e < e1) {
      const int c = static_cast<int>(e / p.hw);
      const float* P = s_par + (c - c_lo) * kEbStride;
      const float med = P[oMed];
      const float zv = ld_stream1(z + e);
      const float q = rintf(zv - med);
      const float s = q + med;                       // "dequantize" / ste_round(z - med) + med
      float out = s;
      if (p.noise_mode) {
        float u;
        if (nz) u = ld_stream1(nz + e);
        else {
          const uint64_t eid = static_cast<uint64_t>(image) * static_cast<uint64_t>(p.ne) + static_cast<uint64_t>(e);
          const uint64_t gid = eid >> 2;
          const Philox4 r = philox4x32_10(static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32),
                                          p.off_lo, p.off_hi, p.seed_lo, p.seed_hi);
          const int k = static_cast<int>(eid & 3);
          u = u32_to_centered_uniform(k == 0 ? r.x : k == 1 ? r.y : k == 2 ? r.z : r.w);
        }
        out = zv + u;                                // "noise": medians ignored
      }
      if (zhat) st_stream1(zhat + e, out);
      if (ste) st_stream1(ste + e, s);
      if (sym) st_stream1(sym + e, __float2int_rn(q));
      if (need_lik) {
        const float lower = logits_cumulative(P, out - 0.5f);
        const float upper = logits_cumulative(P, out + 0.5f);
        const float sum = lower + upper;
        const float sg = (sum < 0.0f) ? 1.0f : ((sum > 0.0f) ? -1.0f : 0.0f);  // -torch.sign(sum); NaN -> 0
        float L = fabsf(sigmoid_ref(sg * upper) - sigmoid_ref(sg * lower));
        if (p.lik_bound > 0.0f) L = max_nan(L, p.lik_bound);
        if (lik) st_stream1(lik + e, L);
        acc += log2f(L);
      }
    }
  }
  if (p.bits) rate_commit(acc, image, static_cast<unsigned int>(p.bpi * (kThreads / 32)), p.B, p.workspace, p.bits);
}

int eb_fwd_launch(const reslic_eb_desc* d, cudaStream_t st) {
  if (!d) return set_error(RESLIC_ERR_ARG, "eb_fwd: null descriptor");
  if (d->B < 0 || d->C < 0 || d->hw < 0) return set_error(RESLIC_ERR_ARG, "eb_fwd: negative size");
  if (d->B == 0 || d->C == 0 || d->hw == 0) return RESLIC_OK;
  if (d->B > (1 << 24) || d->C > (1 << 20) || d->hw > (1LL << 30))
    return set_error(RESLIC_ERR_ARG, "eb_fwd: size too large");
  if (d->mode != RESLIC_Q_DEQUANTIZE && d->mode != RESLIC_Q_NOISE)
    return set_error(RESLIC_ERR_ARG, "eb_fwd: invalid quantization mode");
  if (!d->z || !d->medians) return set_error(RESLIC_ERR_ARG, "eb_fwd: z or medians is null");
  for (int i = 0; i < 5; ++i)
    if (!d->matrix[i] || !d->bias[i] || (i < 4 && !d->factor[i]))
      return set_error(RESLIC_ERR_ARG, "eb_fwd: a parameter pointer is null");
  if (!d->zhat && !d->ste && !d->lik && !d->sym && !d->bits)
    return set_error(RESLIC_ERR_ARG, "eb_fwd: no output requested");

  EbParams p{};
  p.z = d->z; p.z_bs = d->z_bs; p.noise = d->noise; p.noise_bs = d->noise_bs;
  for (int i = 0; i < 5; ++i) { p.matrix[i] = d->matrix[i]; p.bias[i] = d->bias[i]; }
  for (int i = 0; i < 4; ++i) p.factor[i] = d->factor[i];
  p.medians = d->medians;
  p.zhat = d->zhat; p.ste = d->ste; p.lik = d->lik; p.sym = d->sym;
  p.zhat_bs = d->zhat_bs; p.ste_bs = d->ste_bs; p.lik_bs = d->lik_bs; p.sym_bs = d->sym_bs;
  p.B = d->B;
  p.ne = d->C * d->hw; p.hw = static_cast<int>(d->hw); p.C = static_cast<int>(d->C);
  p.noise_mode = d->mode == RESLIC_Q_NOISE; p.lik_bound = d->likelihood_bound;
  p.seed_lo = static_cast<uint32_t>(d->philox_seed); p.seed_hi = static_cast<uint32_t>(d->philox_seed >> 32);
  p.off_lo = static_cast<uint32_t>(d->philox_offset); p.off_hi = static_cast<uint32_t>(d->philox_offset >> 32);
  // a tile may touch at most kEbMaxCh channels: tile <= (kEbMaxCh - 1) * hw
  int64_t tile = kThreads;
  if ((kEbMaxCh - 1) * d->hw < tile) tile = (kEbMaxCh - 1) * d->hw;
  p.tile = static_cast<int>(tile);
  const int64_t ntiles = (p.ne + tile - 1) / tile;
  int64_t bpi = ntiles;
  const int64_t max_ctas = static_cast<int64_t>(sm_count()) * 64;
  if (bpi * d->B > max_ctas) bpi = (max_ctas + d->B - 1) / d->B;
  if (bpi < 1) bpi = 1;
  p.bpi = static_cast<int>(bpi);
  if (d->bits) {
    if (!d->workspace || d->workspace_bytes < reslic_workspace_bytes(d->B))
      return set_error(RESLIC_ERR_WORKSPACE, "eb_fwd: workspace missing or too small for `bits`");
    p.bits = d->bits;
    if (reinterpret_cast<uintptr_t>(d->workspace) & 7u)
      return set_error(RESLIC_ERR_WORKSPACE, "eb_fwd: workspace must be 8-byte aligned");
    p.workspace = static_cast<unsigned long long*>(d->workspace);
  }
  const int64_t grid64 = bpi * d->B;
  if (grid64 > 0x7fffffffLL) return set_error(RESLIC_ERR_ARG, "eb_fwd: grid too large");
  eb_fwd_kernel<<<static_cast<int>(grid64), kThreads, 0, st>>>(p);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_cuda_error(err, "eb_fwd launch");
  return RESLIC_OK;
}

}  // namespace reslic
