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
// hw floats, so the "permute" is just index arithmetic.  Eval mode is a per-channel table of the
// 65 likelihoods about the median (eb_lut8_kernel / eb_lut_kernel); noise mode evaluates the MLP per
// element — eb_fwd_fast_kernel: one CTA per (channel group, batch split), parameters staged once per
// CTA, both cumulative logits as packed f32x2 lanes; eb_fwd_kernel (per-tile staging, library tanhf)
// is the RESLIC_MATH_MIRROR form.  Noise mode is FP32/MUFU-issue bound (24 tanh + 2 sigmoid per
// element), not HBM bound; z is 3.75 % of y's elements.  DESIGN.md section 3.2.
#include <cstdlib>
#include "common.cuh"
#include "eb_math.cuh"
#include "reslic_internal.h"

namespace reslic {

struct EbParams {
  const float* z; const float* noise; int64_t z_bs, noise_bs;
  const float* matrix[5]; const float* bias[5]; const float* factor[4]; const float* medians;
  float* zhat; float* ste; float* lik; int32_t* sym;
  int64_t zhat_bs, ste_bs, lik_bs, sym_bs;
  double* bits; unsigned long long* workspace; int bits_accumulate; int64_t B;
  int64_t ne;       // elements per image = C*hw
  int hw, C, tile, bpi, noise_mode, splits;
  int ipc;          // eb_lut8_kernel: consecutive images per CTA (its 8 tables are staged once for all of them)
  float lik_bound;
  uint32_t seed_lo, seed_hi, off_lo, off_hi;
  const float* next_y; int64_t next_y_bs, next_y_n;   // optional L2 prefetch hint (table mode)
  const float* lut;     // optional prebuilt [C, 2*kLutN] table (eval mode)
  float* lut_out;       // eb_build_lut_kernel's destination
};

// ------------------------------------------------------------------ eval mode: per-channel LUT
// In "dequantize" mode z_hat = k + median with k = round(z - median) an integer, so the
// likelihood is a function of (channel, k) only.  One CTA serves one channel (x a split of the
// batch): it stages the channel's parameters, evaluates the 2*kLutK+1 likelihoods of
// k = -kLutK..kLutK ONCE with exactly the per-element code path (bit-identical results), then
// every element is a rounding, a table lookup and a log2 — ~15 instructions instead of ~800.
// |k| > kLutK (and NaN) fall back to the direct evaluation.  Each (image, channel) run of hw
// contiguous floats is walked by one warp, so the rate commit stays per (warp, image).
constexpr int kLutK = 32;

constexpr int kLutN = 2 * kLutK + 1;
constexpr int kEbChunk = 4;      // values per lane held in registers: one chunk = 128 contiguous floats

// The channel's table: 2*kLutN independent cumulative-logit evaluations (one per thread) with exactly
// the per-element code path, combined into kLutN bounded likelihoods and their log2.  All threads call.
__device__ __forceinline__ void eb_build_table(const EbParams& p, const float* s_par, float med, float* s_half,
                                               float* s_lut, float* s_lg) {
  if (threadIdx.x < 2 * kLutN) {
    const int k = static_cast<int>(threadIdx.x % kLutN) - kLutK;
    const float x = static_cast<float>(k) + med;             // == round(z - med) + med for that symbol
    s_half[threadIdx.x] = logits_cumulative(s_par, threadIdx.x < kLutN ? x - 0.5f : x + 0.5f);
  }
  __syncthreads();
  if (threadIdx.x < kLutN) {
    const float lower = s_half[threadIdx.x], upper = s_half[kLutN + threadIdx.x];
    const float L = eb_combine(lower, upper, p.lik_bound);
    s_lut[threadIdx.x] = L;
    s_lg[threadIdx.x] = log2f(L);
  }
  __syncthreads();
}

// One CTA per channel: the table of eb_lut_kernel written to global memory (reslic_eb_build_lut_f32).
__global__ void __launch_bounds__(kThreads) eb_build_lut_kernel(const EbParams p) {
  __shared__ float s_par[kEbStride + 1];
  __shared__ float s_half[2 * kLutN];
  __shared__ float s_lut[kLutN];
  __shared__ float s_lg[kLutN];
  const int c = blockIdx.x;
  if (threadIdx.x < kEbStride) s_par[threadIdx.x] = eb_staged_param(p, c, threadIdx.x);
  __syncthreads();
  eb_build_table(p, s_par, s_par[oMed], s_half, s_lut, s_lg);
  if (threadIdx.x < kLutN) {
    p.lut_out[static_cast<int64_t>(c) * (2 * kLutN) + threadIdx.x] = s_lut[threadIdx.x];
    p.lut_out[static_cast<int64_t>(c) * (2 * kLutN) + kLutN + threadIdx.x] = s_lg[threadIdx.x];
  }
}

__global__ void __launch_bounds__(kThreads) eb_lut_kernel(const EbParams p) {
  __shared__ float s_par[kEbStride + 1];
  __shared__ float s_half[2 * kLutN];     // lower / upper cumulative logits of the table symbols
  __shared__ float s_lut[kLutN];
  __shared__ float s_lg[kLutN];           // log2 of the table likelihoods
  griddep_wait();
  griddep_launch_dependents();
  const int c = blockIdx.x / p.splits;
  const int split = blockIdx.x - c * p.splits;
  const bool need_lik = p.lik || p.bits;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wstride = p.splits * (kThreads / 32);
  const int64_t base_c = static_cast<int64_t>(c) * p.hw;
  // work items of a warp: (image b, chunk of 128 floats of the (b, c) run), b strided over warps
  const int chunks = (p.hw + 32 * kEbChunk - 1) / (32 * kEbChunk);
  int64_t b = split * (kThreads / 32) + warp;
  int ch = 0;
  float zv[kEbChunk];
  auto load_chunk = [&](int64_t bb, int cc) {
    const float* __restrict__ z = p.z + bb * p.z_bs + base_c + cc * (32 * kEbChunk);
    const int rem = p.hw - cc * (32 * kEbChunk);
#pragma unroll
    for (int j = 0; j < kEbChunk; ++j) zv[j] = (lane + 32 * j < rem) ? ld_stream1(z + lane + 32 * j) : 0.0f;
  };
  if (b < p.B) load_chunk(b, 0);                 // in flight while the table is built

  if (threadIdx.x < kEbStride) s_par[threadIdx.x] = eb_staged_param(p, c, threadIdx.x);
  const bool have_table = need_lik && p.lut != nullptr;    // built once by eb_build_lut_kernel for these parameters
  if (have_table && threadIdx.x >= kThreads - kLutN) {     // the upper warps: their loads overlap the parameter staging
    const int t = threadIdx.x - (kThreads - kLutN);
    s_lut[t] = p.lut[static_cast<int64_t>(c) * (2 * kLutN) + t];
    s_lg[t] = p.lut[static_cast<int64_t>(c) * (2 * kLutN) + kLutN + t];
  }
  __syncthreads();
  const float med = s_par[oMed];
  if (need_lik && !have_table) eb_build_table(p, s_par, med, s_half, s_lut, s_lg);
  while (b < p.B) {
    float acc = 0.0f;
    for (;;) {
      const int64_t off = base_c + ch * (32 * kEbChunk);
      const int rem = p.hw - ch * (32 * kEbChunk);
      float cur[kEbChunk];
#pragma unroll
      for (int j = 0; j < kEbChunk; ++j) cur[j] = zv[j];
      // next work item's loads before this one's math
      const bool last_chunk = (ch + 1 >= chunks);
      const int64_t nb = last_chunk ? b + wstride : b;
      const int nch = last_chunk ? 0 : ch + 1;
      if (nb < p.B) load_chunk(nb, nch);
#pragma unroll
      for (int j = 0; j < kEbChunk; ++j) {
        const int i = lane + 32 * j;
        if (i < rem) {
          const float q = rintf(cur[j] - med);
          const float st = q + med;                   // "dequantize" == ste_round(z - med) + med
          if (p.zhat) st_stream1(p.zhat + b * p.zhat_bs + off + i, st);
          if (p.ste) st_stream1(p.ste + b * p.ste_bs + off + i, st);
          if (p.sym) st_stream1(p.sym + b * p.sym_bs + off + i, __float2int_rn(q));
          if (need_lik) {
            float L, lg;
            if (fabsf(q) <= static_cast<float>(kLutK)) {
              const int k = __float2int_rn(q) + kLutK;
              L = s_lut[k]; lg = s_lg[k];
            } else {
              L = eb_likelihood(s_par, st, p.lik_bound);
              lg = log2f(L);
            }
            if (p.lik) st_stream1(p.lik + b * p.lik_bs + off + i, L);
            acc += lg;
          }
        }
      }
      ch = nch;
      if (last_chunk) break;
    }
    if (p.bits) rate_commit(acc, static_cast<int>(b), static_cast<unsigned int>(p.C), p.B, p.workspace, p.bits,
                            p.bits_accumulate);
    b += wstride;
  }
}

// ------------------------------------------------------------------ eval mode with a prebuilt table
// With the table built once (reslic_eb_build_lut_f32) nothing ties a CTA to one channel any more, so
// the work is cut the other way: one CTA per (image, 8 consecutive channels), one warp per channel run.
// A warp copies its channel's 2 x 65 table entries to shared memory behind a __syncwarp only; the
// parameters are staged lazily, by the warp that meets an out-of-table symbol (|k| > 32 or NaN).  The
// eight warps' partial rates meet in shared memory and ONE thread commits for the CTA: C/8 arrivals
// per image instead of C (192 same-word atomics per image arriving together made this launch's commit
// cost more than its arithmetic).
constexpr int kEbWarps = kThreads / 32;
__global__ void __launch_bounds__(kThreads) eb_lut8_kernel(const EbParams p) {
  __shared__ float s_lut[kEbWarps][kLutN + 1];
  __shared__ float s_lg[kEbWarps][kLutN + 1];
  __shared__ float s_par[kEbWarps][kEbStride + 1];
  __shared__ float s_red[kEbWarps];
  griddep_wait();
  griddep_launch_dependents();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int octets = (p.C + kEbWarps - 1) / kEbWarps;
  const int group = blockIdx.x / octets;                    // CTA = (group of p.ipc consecutive images, 8 channels)
  const int octet = blockIdx.x - group * octets;
  const int c = octet * kEbWarps + warp;
  const int image0 = group * p.ipc;
  const int image1 = min(static_cast<int>(p.B), image0 + p.ipc);
  const bool need_lik = p.lik || p.bits;
  float med = 0.0f;
  if (c < p.C) {
    med = p.medians[c];
    if (need_lik) {         // the channel's table: once per CTA, whatever the number of images it walks
      for (int t = lane; t < kLutN; t += 32) {
        s_lut[warp][t] = p.lut[static_cast<int64_t>(c) * (2 * kLutN) + t];
        s_lg[warp][t] = p.lut[static_cast<int64_t>(c) * (2 * kLutN) + kLutN + t];
      }
      __syncwarp();
    }
  }
  bool staged = false;
  // (image, chunk) pairs are walked as one flat sequence, the next pair's values loaded before the current pair is
  // processed: with hw = 96 an image is ONE chunk per warp, and without the look-ahead a CTA would pay a full DRAM
  // round trip per image
  const int chunks = (p.hw + 32 * kEbChunk - 1) / (32 * kEbChunk);
  const int iters = (image1 - image0) * chunks;
  const int64_t base_c = static_cast<int64_t>(c) * p.hw;
  float zv[kEbChunk];
  auto load_iter = [&](int it) {
    const int im = image0 + it / chunks, cc = it - (it / chunks) * chunks;
    const float* __restrict__ z = p.z + im * p.z_bs + base_c;
    const int rem = p.hw - cc * (32 * kEbChunk);
#pragma unroll
    for (int j = 0; j < kEbChunk; ++j) zv[j] = (lane + 32 * j < rem) ? ld_stream1(z + cc * (32 * kEbChunk) + lane + 32 * j) : 0.0f;
  };
  if (c < p.C && iters > 0) load_iter(0);
  float acc = 0.0f;
  for (int it = 0; it < iters; ++it) {
    const int image = image0 + it / chunks, ch = it - (it / chunks) * chunks;
    if (ch == 0 && p.next_y != nullptr && threadIdx.x == 0) {
      // this CTA's share of the image's next_y bytes, one bulk L2 prefetch (this launch moves 12 bytes per z element:
      // HBM is idle, and the slice launch that follows is short enough to feel a third of its reads arriving early)
      const unsigned int total = static_cast<unsigned int>(p.next_y_n) * 4u;
      const unsigned int share = ((total + octets - 1) / octets + 15u) & ~15u;
      const unsigned int off = static_cast<unsigned int>(octet) * share;
      if (off < total) {
        const unsigned int bytes = min(share, total - off);
        const char* addr = reinterpret_cast<const char*>(p.next_y + static_cast<int64_t>(image) * p.next_y_bs) + off;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(addr), "r"(bytes) : "memory");
      }
    }
    if (c < p.C) {
      const int64_t off = base_c + ch * (32 * kEbChunk);
      const int rem = p.hw - ch * (32 * kEbChunk);
      float cur[kEbChunk];
#pragma unroll
      for (int j = 0; j < kEbChunk; ++j) cur[j] = zv[j];
      if (it + 1 < iters) load_iter(it + 1);
#pragma unroll
      for (int j = 0; j < kEbChunk; ++j) {
        const int i = lane + 32 * j;
        const bool valid = i < rem;
        const float q = rintf(cur[j] - med);
        const float st = q + med;                   // "dequantize" == ste_round(z - med) + med
        const bool far = need_lik && valid && !(fabsf(q) <= static_cast<float>(kLutK));
        if (__any_sync(0xffffffffu, far) && !staged) {
          for (int t = lane; t < kEbStride; t += 32) s_par[warp][t] = eb_staged_param(p, c, t);
          __syncwarp();
          staged = true;
        }
        if (valid) {
          if (p.zhat) st_stream1(p.zhat + image * p.zhat_bs + off + i, st);
          if (p.ste) st_stream1(p.ste + image * p.ste_bs + off + i, st);
          if (p.sym) st_stream1(p.sym + image * p.sym_bs + off + i, __float2int_rn(q));
          if (need_lik) {
            float L, lg;
            if (!far) {
              const int k = __float2int_rn(q) + kLutK;
              L = s_lut[warp][k]; lg = s_lg[warp][k];
            } else {
              L = eb_likelihood(s_par[warp], st, p.lik_bound);
              lg = log2f(L);
            }
            if (p.lik) st_stream1(p.lik + image * p.lik_bs + off + i, L);
            acc += lg;
          }
        }
      }
    }
    if (ch + 1 == chunks && p.bits) {                 // the image is complete: one commit per (CTA, image)
      const float v = warp_sum_f32(acc);
      acc = 0.0f;
      __syncthreads();                                // the previous image's reader is done with s_red
      if (lane == 0) s_red[warp] = v;
      __syncthreads();
      if (warp == 0) {
        float total = 0.0f;
        if (lane == 0) {
#pragma unroll
          for (int w = 0; w < kEbWarps; ++w) total += s_red[w];     // fixed order: reproducible
        }
        rate_commit(total, image, static_cast<unsigned int>(octets), p.B, p.workspace, p.bits, p.bits_accumulate);
      }
    }
  }
}

// ------------------------------------------------------------------ noise mode: direct evaluation
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
      const int cl = i / kEbStride, j = i - cl * kEbStride;
      const float v = eb_staged_param(p, c_lo + cl, j);
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
        const float L = eb_likelihood(P, out, p.lik_bound);
        if (lik) st_stream1(lik + e, L);
        acc += log2f(L);
      }
    }
  }
  if (p.bits) rate_commit(acc, image, static_cast<unsigned int>(p.bpi * (kThreads / 32)), p.B, p.workspace, p.bits, p.bits_accumulate);
}

// ------------------------------------------------------------------ noise / direct mode, fast math
// One CTA per (group of `cpc` consecutive channels, split of the batch): the transformed parameters of the
// group are staged ONCE per CTA and reused for every image of the split (eb_fwd_kernel re-stages them —
// softplus and tanh included — for every 256-element tile).  The group's elements of one image are one
// contiguous run of cpc*hw floats; the next image's values are loaded before the current ones are
// evaluated.  Both cumulative logits of an element run as the two lanes of packed f32x2 operations
// (eb_math.cuh).  Every warp commits its own rate per image (fixed point, so the order does not matter).
//
// QUAD (in-kernel noise, hw % 4 == 0, at most one element per thread and image): the four lanes of a quad hold the four
// elements of ONE Philox counter (element group eid >> 2), so instead of four identical calls per image each lane
// evaluates the counter of a different one of the CTA's next four images and a 4 x 4 transpose across the quad (four
// shuffles) hands every lane its own word of each — one Philox call per four elements, same noise field.
constexpr int kEbRunMax = 4;      // elements per thread and image: runs of up to 1024 floats
template <bool QUAD>
__global__ void __launch_bounds__(kThreads) eb_fwd_fast_kernel(const EbParams p, int cpc, int groups, int splits, float lik_floor) {
  __shared__ float s_par[kEbMaxCh * kEbStride];
  const int grp = blockIdx.x % groups;
  const int split = blockIdx.x / groups;
  const int c0 = grp * cpc;
  const int nch = min(cpc, p.C - c0);
  const int run = nch * p.hw;                       // <= kEbRunMax * kThreads (host)
  const int64_t base = static_cast<int64_t>(c0) * p.hw;
  const bool need_lik = p.lik || p.bits;
  float zv[kEbRunMax], nv[kEbRunMax];
  auto load_image = [&](int64_t b) {
    const float* __restrict__ z = p.z + b * p.z_bs + base;
    const float* __restrict__ nz = (p.noise_mode && p.noise) ? p.noise + b * p.noise_bs + base : nullptr;
#pragma unroll
    for (int j = 0; j < kEbRunMax; ++j) {
      const int i = threadIdx.x + j * kThreads;
      zv[j] = (i < run) ? ld_stream1(z + i) : 0.0f;
      nv[j] = (nz && i < run) ? ld_stream1(nz + i) : 0.0f;
    }
  };
  int64_t b = split;
  if (b < p.B) load_image(b);                       // in flight while the parameters are transformed
  if (need_lik) {
    // thread -> (parameter slot j, channel quarter): the slot's source array, stride and transform are resolved
    // once, the loop over the group's channels is then load / transform / store
    const int j = threadIdx.x & 63, cq = threadIdx.x >> 6;
    if (j < kEbStride) {
      const float* src; int stride, kind;                 // kind 0: copy, 1: softplus, 2: tanh
      if (j < oB0) { src = p.matrix[0] + j; stride = 3; kind = 1; }
      else if (j < oF0) { src = p.bias[0] + (j - oB0); stride = 3; kind = 0; }
      else if (j < oM1) { src = p.factor[0] + (j - oF0); stride = 3; kind = 2; }
      else if (j < oM4) {
        const int l = (j - oM1) / 15, r = (j - oM1) - l * 15;
        if (r < 9) { src = p.matrix[1 + l] + r; stride = 9; kind = 1; }
        else if (r < 12) { src = p.bias[1 + l] + (r - 9); stride = 3; kind = 0; }
        else { src = p.factor[1 + l] + (r - 12); stride = 3; kind = 2; }
      }
      else if (j < oB4) { src = p.matrix[4] + (j - oM4); stride = 3; kind = 1; }
      else if (j == oB4) { src = p.bias[4]; stride = 1; kind = 0; }
      else { src = p.medians; stride = 1; kind = 0; }
      for (int cl = cq; cl < nch; cl += kThreads / 64) {
        float v = src[static_cast<int64_t>(c0 + cl) * stride];
        if (kind == 1) v = v > 15.0f ? v : 0.6931471805599453f * lg2_approx(1.0f + ex2_approx(1.4426950408889634f * v));
        else if (kind == 2) {
          const float e = ex2_approx(-2.8853900817779268f * fabsf(v));
          v = copysignf((1.0f - e) * rcp_approx(1.0f + e), v);
        }
        s_par[cl * kEbStride + j] = v;
      }
    }
  } else {
    for (int i = threadIdx.x; i < nch; i += kThreads) s_par[i * kEbStride + oMed] = p.medians[c0 + i];
  }
  __syncthreads();
  // channel of this thread's j-th element: fixed for the whole launch
  int cl[kEbRunMax];
#pragma unroll
  for (int j = 0; j < kEbRunMax; ++j) cl[j] = (j * kThreads < run) ? min((threadIdx.x + j * kThreads) / p.hw, nch - 1) : 0;
  const unsigned int expected = static_cast<unsigned int>(groups) * kEbWarps;   // warps committing to one image
  unsigned long long pend_now = 0ull;
  int pend_image = -1;
  float uq[4] = {0.0f, 0.0f, 0.0f, 0.0f};             // QUAD: this element's noise in the next four images
  for (int nq = 0; b < p.B; b += splits, ++nq) {
    float cz[kEbRunMax], cn[kEbRunMax];
#pragma unroll
    for (int j = 0; j < kEbRunMax; ++j) { cz[j] = zv[j]; cn[j] = nv[j]; }
    if (b + splits < p.B) load_image(b + splits);
    if (QUAD && (nq & 3) == 0) {
      const int k = threadIdx.x & 3;
      int64_t bk = b + static_cast<int64_t>(k) * splits;
      if (bk >= p.B) bk = b;                            // past the CTA's last image: any counter, the word is not used
      const uint64_t gid = (static_cast<uint64_t>(bk) * static_cast<uint64_t>(p.ne) + static_cast<uint64_t>(base + threadIdx.x)) >> 2;
      const Philox4 r = philox4x32_10(static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32), p.off_lo, p.off_hi, p.seed_lo, p.seed_hi);
      float a[4] = {u32_to_centered_uniform(r.x), u32_to_centered_uniform(r.y), u32_to_centered_uniform(r.z), u32_to_centered_uniform(r.w)};
      // rows = lanes (images), columns = words (elements): transpose in two exchange steps
      const bool odd = (k & 1) != 0, hi = (k & 2) != 0;
      float r0 = __shfl_xor_sync(0xffffffffu, odd ? a[0] : a[1], 1), r1 = __shfl_xor_sync(0xffffffffu, odd ? a[2] : a[3], 1);
      if (odd) { a[0] = r0; a[2] = r1; } else { a[1] = r0; a[3] = r1; }
      r0 = __shfl_xor_sync(0xffffffffu, hi ? a[0] : a[2], 2); r1 = __shfl_xor_sync(0xffffffffu, hi ? a[1] : a[3], 2);
      if (hi) { a[0] = r0; a[1] = r1; } else { a[2] = r0; a[3] = r1; }
      uq[0] = a[0]; uq[1] = a[1]; uq[2] = a[2]; uq[3] = a[3];
    }
    float acc = 0.0f;
#pragma unroll
    for (int j = 0; j < kEbRunMax; ++j) {
      const int i = threadIdx.x + j * kThreads;
      if (i < run) {
        const float* P = s_par + cl[j] * kEbStride;
        const float med = P[oMed];
        const float q = rintf(cz[j] - med);
        const float s = q + med;                       // "dequantize" / ste_round(z - med) + med
        float out = s;
        if (p.noise_mode) {
          float u = cn[j];
          if (QUAD) {
            u = uq[0]; uq[0] = uq[1]; uq[1] = uq[2]; uq[2] = uq[3];
          } else if (!p.noise) {
            const uint64_t eid = static_cast<uint64_t>(b) * static_cast<uint64_t>(p.ne) + static_cast<uint64_t>(base + i);
            const uint64_t gid = eid >> 2;
            const Philox4 r = philox4x32_10(static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32),
                                            p.off_lo, p.off_hi, p.seed_lo, p.seed_hi);
            const int k = static_cast<int>(eid & 3);
            u = u32_to_centered_uniform(k == 0 ? r.x : k == 1 ? r.y : k == 2 ? r.z : r.w);
          }
          out = cz[j] + u;                             // "noise": medians ignored
        }
        const int64_t e = base + i;
        if (p.zhat) st_stream1(p.zhat + b * p.zhat_bs + e, out);
        if (p.ste) st_stream1(p.ste + b * p.ste_bs + e, s);
        if (p.sym) st_stream1(p.sym + b * p.sym_bs + e, __float2int_rn(q));
        if (need_lik) {
          const float L = eb_likelihood_fast(P, out, lik_floor);
          if (p.lik) st_stream1(p.lik + b * p.lik_bs + e, L);
          acc += lg2_approx(L);
        }
      }
    }
    if (p.bits) {
      // every warp commits its own sum (integer fixed point: order-free, reproducible) — no CTA barrier per image;
      // the immediate forms look at the atomic's answer one image later, off the critical path
      if (p.bits_accumulate == 2) rate_defer(acc, static_cast<int>(b), p.B, p.workspace);
      else {
        if (pend_image >= 0) rate_commit_finish(pend_now, pend_image, expected, p.B, p.workspace, p.bits, p.bits_accumulate == 1, p.bits_accumulate == 3);
        pend_now = rate_commit_issue(acc, static_cast<int>(b), p.B, p.workspace);
        pend_image = static_cast<int>(b);
      }
    }
  }
  if (pend_image >= 0) rate_commit_finish(pend_now, pend_image, expected, p.B, p.workspace, p.bits, p.bits_accumulate == 1, p.bits_accumulate == 3);
}

int eb_fwd_launch(const reslic_eb_desc* d, cudaStream_t st) {
  if (!d) return set_error(RESLIC_ERR_ARG, "eb_fwd: null descriptor");
  if (d->struct_size != sizeof(reslic_eb_desc))
    return set_error(RESLIC_ERR_ARG, "eb_fwd: struct_size != sizeof(reslic_eb_desc) (binding built against another ABI revision)");
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
  if (!d->zhat && !d->ste && !d->lik && !d->sym && !rate_requested(d->bits, d->bits_accumulate))
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
  p.lut = p.noise_mode ? nullptr : d->lut;
  // prefetch hint: 16-byte aligned images of a multiple of 4 floats, < 4 GB each; anything else is dropped
  // (and only up to 12 MB: beyond that the launch that follows is long enough to have no idle HBM time to fill,
  // see gc_fused.cu)
  if (d->next_y && d->next_y_n > 0 && d->next_y_n < (1LL << 30) && (d->next_y_n % 4) == 0 && (d->next_y_bs % 4) == 0 &&
      (reinterpret_cast<uintptr_t>(d->next_y) & 15u) == 0 && d->next_y_n * d->B * 4 <= (12LL << 20)) {
    p.next_y = d->next_y; p.next_y_bs = d->next_y_bs; p.next_y_n = d->next_y_n;
  }
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
  if (rate_requested(d->bits, d->bits_accumulate)) {
    const int rc = rate_setup("eb_fwd", d->bits, d->bits_accumulate, d->workspace, d->workspace_bytes, d->B,
                              &p.bits, &p.bits_accumulate, &p.workspace);
    if (rc != RESLIC_OK) return rc;
  }
  cudaError_t err;
  if (!p.noise_mode && p.lut) {
    // eval with a prebuilt table: one CTA per (image, 8 channels)
    const int64_t octets = (d->C + kEbWarps - 1) / kEbWarps;
    if (octets * d->B > 0x7fffffffLL) return set_error(RESLIC_ERR_ARG, "eb_fwd: grid too large");
    if (octets > 60000) return set_error(RESLIC_ERR_ARG, "eb_fwd: too many channels for the rate arrival count");
    // a CTA walks `ipc` consecutive images with its eight tables staged once: about 1.3 CTAs per SM (a launch of one CTA
    // per (image, octet) spends more on launching CTAs and copying tables than on z: 64 x 192 x 12 x 8 as 1536 CTAs took
    // 11.2 us under ncu for 14 MB of traffic; measured in the config-3 step, images per CTA 1 / 3 / 8: 120.7 / 117.2 /
    // 115.7 us)
    int64_t ipc = (octets * d->B + 13LL * sm_count() / 10 - 1) / (13LL * sm_count() / 10);
    if (ipc < 1) ipc = 1;
    if (const char* e = std::getenv("RESLIC_EB_IPC")) { const long v = std::atol(e); if (v >= 1) ipc = v; }
    if (ipc > d->B) ipc = d->B;
    p.ipc = static_cast<int>(ipc);
    const int64_t groups_of_images = (d->B + ipc - 1) / ipc;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(octets * groups_of_images));
    cfg.blockDim = dim3(kThreads);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = gc_tuning().pdl ? 1 : 0;
    err = cudaLaunchKernelEx(&cfg, eb_lut8_kernel, p);
  } else if (!p.noise_mode) {
    // eval: per-channel LUT kernel, one CTA per (channel, batch split)
    // one (image, channel) run per warp where the machine has room for it (<= 8 CTAs per SM)
    int64_t splits = (d->B + (kThreads / 32) - 1) / (kThreads / 32);
    const int64_t cap = (8 * static_cast<int64_t>(sm_count()) + d->C - 1) / d->C;
    if (splits > cap) splits = cap;
    if (splits < 1) splits = 1;
    p.splits = static_cast<int>(splits);
    if (d->C * splits > 0x7fffffffLL) return set_error(RESLIC_ERR_ARG, "eb_fwd: grid too large");
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(d->C * splits));
    cfg.blockDim = dim3(kThreads);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = gc_tuning().pdl ? 1 : 0;
    err = cudaLaunchKernelEx(&cfg, eb_lut_kernel, p);
  } else {
    // direct evaluation.  Fast math: channel groups of ~256 elements per image, parameters staged once per CTA
    int64_t cpc = (kThreads + d->hw - 1) / d->hw;
    if (cpc > kEbMaxCh) cpc = kEbMaxCh;
    if (cpc > d->C) cpc = d->C;
    if (cpc < 1) cpc = 1;
    const int64_t groups = (d->C + cpc - 1) / cpc;
    const bool fast_ok = math_mode() != RESLIC_MATH_MIRROR && cpc * d->hw <= kEbRunMax * kThreads && groups * kEbWarps < 65536;   // 16-bit arrival count
    if (fast_ok) {
      // one resident wave, never more: a CTA past the resident set starts when the first ones finish and
      // runs its whole image list behind them (600 CTAs on 592 slots cost a second pass over the batch)
      static const int resident = [] {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, eb_fwd_fast_kernel<true>, kThreads, 0) != cudaSuccess || n < 1) n = 1;
        return n;
      }();
      static const long forced = [] { const char* e = std::getenv("RESLIC_EB_SPLITS"); return e ? std::atol(e) : 0L; }();
      int64_t splits = forced >= 1 ? forced : static_cast<int64_t>(sm_count()) * resident / groups;
      if (splits > d->B) splits = d->B;
      if (splits < 1) splits = 1;
      const float lik_floor = d->likelihood_bound > 0.0f ? d->likelihood_bound : -__builtin_huge_valf();
      // (the quad's word order needs element groups aligned to lanes: hw % 4 == 0 makes base and ne multiples of 4)
      const bool quad = p.noise_mode && !p.noise && d->hw % 4 == 0 && cpc * d->hw <= kThreads;
      if (quad) eb_fwd_fast_kernel<true><<<static_cast<int>(groups * splits), kThreads, 0, st>>>(p, static_cast<int>(cpc), static_cast<int>(groups),
                                                                                         static_cast<int>(splits), lik_floor);
      else eb_fwd_fast_kernel<false><<<static_cast<int>(groups * splits), kThreads, 0, st>>>(p, static_cast<int>(cpc), static_cast<int>(groups),
                                                                                          static_cast<int>(splits), lik_floor);
    } else {
      const int64_t grid64 = bpi * d->B;
      if (grid64 > 0x7fffffffLL) return set_error(RESLIC_ERR_ARG, "eb_fwd: grid too large");
      eb_fwd_kernel<<<static_cast<int>(grid64), kThreads, 0, st>>>(p);
    }
    err = cudaGetLastError();
  }
  if (err != cudaSuccess) return set_cuda_error(err, "eb_fwd launch");
  return RESLIC_OK;
}

int eb_build_lut_launch(const reslic_eb_desc* d, float* lut, cudaStream_t st) {
  if (!d || !lut) return set_error(RESLIC_ERR_ARG, "eb_build_lut: null descriptor or table");
  if (d->struct_size != sizeof(reslic_eb_desc))
    return set_error(RESLIC_ERR_ARG, "eb_build_lut: struct_size != sizeof(reslic_eb_desc) (binding built against another ABI revision)");
  if (d->C < 0 || d->C > (1 << 20)) return set_error(RESLIC_ERR_ARG, "eb_build_lut: C out of range");
  if (d->C == 0) return RESLIC_OK;
  if (!d->medians) return set_error(RESLIC_ERR_ARG, "eb_build_lut: medians is null");
  for (int i = 0; i < 5; ++i)
    if (!d->matrix[i] || !d->bias[i] || (i < 4 && !d->factor[i]))
      return set_error(RESLIC_ERR_ARG, "eb_build_lut: a parameter pointer is null");
  EbParams p{};
  for (int i = 0; i < 5; ++i) { p.matrix[i] = d->matrix[i]; p.bias[i] = d->bias[i]; }
  for (int i = 0; i < 4; ++i) p.factor[i] = d->factor[i];
  p.medians = d->medians;
  p.C = static_cast<int>(d->C);
  p.lik_bound = d->likelihood_bound;
  p.lut_out = lut;
  eb_build_lut_kernel<<<static_cast<int>(d->C), kThreads, 0, st>>>(p);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_cuda_error(err, "eb_build_lut launch");
  return RESLIC_OK;
}

}  // namespace reslic
