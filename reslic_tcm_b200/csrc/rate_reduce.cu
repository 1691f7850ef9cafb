// rate_reduce.cu — per-image rate of a likelihood tensor (sm_100a).
//
// The reference's loss does, per likelihood tensor, `torch.log(likelihoods).sum() / (-math.log(2) * num_pixels)`
// (src/training/loss.py:22-25, src/eval.py:27-31): a log kernel that writes a tensor the size of its input and a
// reduction that reads it back.  Here one pass reads L (4 B/element, 128-bit loads) and commits
// bits[b] = -sum log2 L per image through the same deterministic integer commit as the fused forward kernels
// (common.cuh), so the result is bit-reproducible and can share a workspace / the deferred and collect modes
// with them.  This is the entry a caller uses when it only HAS likelihood tensors (the reference's
// RateDistortionLoss API); the fused forward kernels emit the same sums without ever re-reading L.
#include "common.cuh"
#include "reslic_internal.h"

namespace reslic {

struct RateParams {
  const float* lik; int64_t lik_bs; int64_t n, B;
  int cpi;                        // CTAs per image
  int vec, fast;
  double* bits; unsigned long long* workspace; int bits_accumulate;
};

__global__ void __launch_bounds__(kThreads) rate_from_lik_kernel(const RateParams p) {
  __shared__ float s_red[kThreads / 32];
  const int image = blockIdx.x / p.cpi;
  const int chunk = blockIdx.x - image * p.cpi;
  const float* __restrict__ L = p.lik + image * p.lik_bs;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc = 0.0f;
  const int64_t stride = static_cast<int64_t>(p.cpi) * kThreads;
  const int64_t first = static_cast<int64_t>(chunk) * kThreads + threadIdx.x;
  const int64_t n_vec = p.vec ? (p.n >> 2) : 0;
  for (int64_t g = first; g < n_vec; g += stride) {
    const float4 v = ld_stream4(L + 4 * g);
    if (p.fast) acc += (lg2_approx(v.x) + lg2_approx(v.y)) + (lg2_approx(v.z) + lg2_approx(v.w));
    else acc += (log2f(v.x) + log2f(v.y)) + (log2f(v.z) + log2f(v.w));
  }
  for (int64_t i = 4 * n_vec + first; i < p.n; i += stride) acc += p.fast ? lg2_approx(L[i]) : log2f(L[i]);
  // one commit per (CTA, image): the eight warps' sums meet in shared memory in a fixed order
  const float v = warp_sum_f32(acc);
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float total = 0.0f;
    if (lane == 0) {
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) total += s_red[w];
    }
    rate_commit(total, image, static_cast<unsigned int>(p.cpi), p.B, p.workspace, p.bits, p.bits_accumulate);
  }
}

int rate_from_lik_launch(const float* lik, int64_t lik_bs, int64_t B, int64_t n, double* bits, int32_t bits_accumulate,
                         void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  if (B < 0 || n < 0) return set_error(RESLIC_ERR_ARG, "rate_from_likelihood: negative size");
  if (B == 0) return RESLIC_OK;
  if (B > (1 << 24)) return set_error(RESLIC_ERR_ARG, "rate_from_likelihood: B too large");
  if (n > 0 && !lik) return set_error(RESLIC_ERR_ARG, "rate_from_likelihood: likelihood tensor is null");
  if (!rate_requested(bits, bits_accumulate)) return set_error(RESLIC_ERR_ARG, "rate_from_likelihood: bits is null");
  RateParams p{};
  p.lik = lik; p.lik_bs = lik_bs; p.n = n; p.B = B;
  const int rc = rate_setup("rate_from_likelihood", bits, bits_accumulate, workspace, workspace_bytes, B, &p.bits,
                            &p.bits_accumulate, &p.workspace);
  if (rc != RESLIC_OK) return rc;
  p.vec = ((reinterpret_cast<uintptr_t>(lik) & 15u) == 0 && (lik_bs % 4) == 0) ? 1 : 0;
  p.fast = math_mode() != RESLIC_MATH_MIRROR;
  // CTAs per image: about eight groups per thread, at most what fills the machine 16 times, arrival count < 2^16
  int64_t cpi = ((n + 3) / 4 + 8LL * kThreads - 1) / (8LL * kThreads);
  const int64_t cap = (16LL * sm_count() + B - 1) / B;
  if (cpi > cap) cpi = cap;
  if (cpi > 60000) cpi = 60000;
  if (cpi < 1) cpi = 1;
  p.cpi = static_cast<int>(cpi);
  if (cpi * B > 0x7fffffffLL) return set_error(RESLIC_ERR_ARG, "rate_from_likelihood: grid too large");
  rate_from_lik_kernel<<<static_cast<int>(cpi * B), kThreads, 0, st>>>(p);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_cuda_error(err, "rate_from_likelihood launch");
  return RESLIC_OK;
}

}  // namespace reslic
