// cabi.cu — the extern "C" boundary declared in include/reslic_b200.h.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "reslic_internal.h"

namespace reslic {

static thread_local char g_err[512] = "";
static int g_math_mode = RESLIC_MATH_FAST;
int math_mode() { return g_math_mode; }
const GcTuning& gc_tuning() {
  static GcTuning t = [] {
    GcTuning v{-1, 1, 0, 1};
    if (const char* e = std::getenv("RESLIC_GC_CTAS_PER_SM")) v.ctas_per_sm = std::atoi(e);
    if (const char* e = std::getenv("RESLIC_PDL")) v.pdl = std::atoi(e) != 0;
    if (const char* e = std::getenv("RESLIC_GC_BALANCE")) v.balance = std::atoi(e);
    if (const char* e = std::getenv("RESLIC_GC_MIN_CTAS")) v.min_ctas = std::atoi(e);   // 4 / 5: force one build; 0: by launch size
    return v;
  }();
  return t;
}
int set_error(int code, const char* msg) {
  std::snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
int set_cuda_error(cudaError_t err, const char* where) {
  std::snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorName(err), cudaGetErrorString(err));
  return static_cast<int>(err);
}
int rate_setup(const char* who, double* bits, int32_t mode, void* workspace, int64_t workspace_bytes, int64_t B,
               double** bits_out, int* mode_out, unsigned long long** ws_out) {
  char msg[160];
  if (mode < 0 || mode > RESLIC_RATE_COLLECT) {
    std::snprintf(msg, sizeof(msg), "%s: bits_accumulate must be 0, 1, RESLIC_RATE_DEFERRED or RESLIC_RATE_COLLECT", who);
    return set_error(RESLIC_ERR_ARG, msg);
  }
  if (mode != RESLIC_RATE_DEFERRED && !bits) {
    std::snprintf(msg, sizeof(msg), "%s: bits is null", who);
    return set_error(RESLIC_ERR_ARG, msg);
  }
  if (!workspace || workspace_bytes < reslic_workspace_bytes(B) || (reinterpret_cast<uintptr_t>(workspace) & 7u)) {
    std::snprintf(msg, sizeof(msg), "%s: workspace missing, misaligned or too small for the rate output", who);
    return set_error(RESLIC_ERR_WORKSPACE, msg);
  }
  *ws_out = static_cast<unsigned long long*>(workspace);
  *mode_out = mode;
  *bits_out = bits ? bits : reinterpret_cast<double*>(workspace);   // deferred: sentinel, never dereferenced
  return RESLIC_OK;
}

// bits[b] (= or +=) deferred sum of image b; leaves the deferred words zero.
__global__ void rate_finalize_kernel(unsigned long long* ws, int64_t B, double* bits, int accumulate) {
  griddep_wait();
  griddep_launch_dependents();
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= B) return;
  const long long sum = static_cast<long long>(ws[2 * B + i]);
  const unsigned long long flag = ws[3 * B + i];
  double v = static_cast<double>(sum) * (1.0 / 65536.0);
  if (flag & 1ull) v = __longlong_as_double(0x7ff8000000000000LL);
  else if (flag & 2ull) v = __longlong_as_double(0x7ff0000000000000LL);
  bits[i] = accumulate ? bits[i] + v : v;
  ws[2 * B + i] = 0ull;
  if (flag) ws[3 * B + i] = 0ull;
}
int rate_finalize_launch(void* workspace, int64_t workspace_bytes, int64_t B, double* bits, int32_t accumulate,
                         cudaStream_t st) {
  if (B < 0) return set_error(RESLIC_ERR_ARG, "rate_finalize: negative B");
  if (B == 0) return RESLIC_OK;
  if (!bits) return set_error(RESLIC_ERR_ARG, "rate_finalize: bits is null");
  if (!workspace || workspace_bytes < reslic_workspace_bytes(B) || (reinterpret_cast<uintptr_t>(workspace) & 7u))
    return set_error(RESLIC_ERR_WORKSPACE, "rate_finalize: workspace missing, misaligned or too small");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>((B + 127) / 128));
  cfg.blockDim = dim3(128);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = gc_tuning().pdl ? 1 : 0;
  cudaError_t err = cudaLaunchKernelEx(&cfg, rate_finalize_kernel, static_cast<unsigned long long*>(workspace), B, bits,
                                       static_cast<int>(accumulate != 0));
  if (err != cudaSuccess) return set_cuda_error(err, "rate_finalize launch");
  return RESLIC_OK;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

__global__ void __launch_bounds__(kThreads) dequantize_kernel(const int32_t* __restrict__ sym,
                                                             const float* __restrict__ mu, int64_t n,
                                                             float* __restrict__ out) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * kThreads) {
    const float v = static_cast<float>(sym[i]);
    out[i] = mu ? v + mu[i] : v;
  }
}

__global__ void __launch_bounds__(kThreads) lrp_tail_kernel(float* __restrict__ y_hat, int64_t y_bs,
                                                           const float* __restrict__ lrp, int64_t l_bs, int64_t n,
                                                           int64_t tiles_per_image, int64_t total) {
  for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
    const int64_t image = t / tiles_per_image;
    const int64_t e = (t - image * tiles_per_image) * kThreads + threadIdx.x;
    if (e < n) y_hat[image * y_bs + e] += 0.5f * tanhf(lrp[image * l_bs + e]);
  }
}

int lrp_tail_launch(float* y_hat, int64_t y_bs, const float* lrp, int64_t l_bs, int64_t B, int64_t n, cudaStream_t st) {
  if (B < 0 || n < 0) return set_error(RESLIC_ERR_ARG, "lrp_tail: negative size");
  if (B == 0 || n == 0) return RESLIC_OK;
  if (!y_hat || !lrp) return set_error(RESLIC_ERR_ARG, "lrp_tail: null pointer");
  const int64_t tpi = (n + kThreads - 1) / kThreads, total = tpi * B;
  int64_t grid = total;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 32;
  if (grid > cap) grid = cap;
  lrp_tail_kernel<<<static_cast<int>(grid), kThreads, 0, st>>>(y_hat, y_bs, lrp, l_bs, n, tpi, total);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_cuda_error(err, "lrp_tail launch");
  return RESLIC_OK;
}

int dequantize_launch(const int32_t* sym, const float* mu, int64_t n, float* out, cudaStream_t st) {
  if (n < 0) return set_error(RESLIC_ERR_ARG, "dequantize: negative size");
  if (n == 0) return RESLIC_OK;
  if (!sym || !out) return set_error(RESLIC_ERR_ARG, "dequantize: null pointer");
  int64_t blocks = (n + kThreads - 1) / kThreads;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 32;
  if (blocks > cap) blocks = cap;
  dequantize_kernel<<<static_cast<int>(blocks), kThreads, 0, st>>>(sym, mu, n, out);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_cuda_error(err, "dequantize launch");
  return RESLIC_OK;
}

}  // namespace reslic

extern "C" {

int reslic_abi_version(void) { return RESLIC_ABI_VERSION; }
const char* reslic_last_error(void) { return reslic::g_err; }
int reslic_device_sm_count(void) { return reslic::sm_count(); }
int reslic_set_math_mode(int mode) {
  if (mode != RESLIC_MATH_FAST && mode != RESLIC_MATH_MIRROR)
    return reslic::set_error(RESLIC_ERR_ARG, "set_math_mode: unknown mode");
  reslic::g_math_mode = mode;
  return RESLIC_OK;
}
int reslic_get_math_mode(void) { return reslic::g_math_mode; }
int64_t reslic_sizeof_gc_desc(void) { return sizeof(reslic_gc_desc); }
int64_t reslic_sizeof_gc_bwd_desc(void) { return sizeof(reslic_gc_bwd_desc); }
int64_t reslic_sizeof_eb_desc(void) { return sizeof(reslic_eb_desc); }
int64_t reslic_sizeof_eb_bwd_desc(void) { return sizeof(reslic_eb_bwd_desc); }
int64_t reslic_sizeof_stanh_tables(void) { return sizeof(reslic_stanh_tables); }
int64_t reslic_sizeof_stanh_gc_desc(void) { return sizeof(reslic_stanh_gc_desc); }
int64_t reslic_sizeof_stanh_gc_bwd_desc(void) { return sizeof(reslic_stanh_gc_bwd_desc); }
int64_t reslic_sizeof_eb_stanh_desc(void) { return sizeof(reslic_eb_stanh_desc); }
int64_t reslic_sizeof_rate_exchange(void) { return sizeof(reslic_rate_exchange); }
int64_t reslic_workspace_bytes(int64_t B) {
  if (B <= 0) return 0;
  return (4 * B + 4) * static_cast<int64_t>(sizeof(unsigned long long));   // + batch total / count / flags / spare (rate_publish)
}
int reslic_rate_finalize_f64(void* workspace, int64_t workspace_bytes, int64_t B, double* bits, int32_t accumulate,
                             void* stream) {
  return reslic::rate_finalize_launch(workspace, workspace_bytes, B, bits, accumulate, static_cast<cudaStream_t>(stream));
}

int reslic_gc_fwd_f32(const reslic_gc_desc* d, void* stream) {
  return reslic::gc_fwd_launch(d, static_cast<cudaStream_t>(stream));
}

int reslic_gc_bwd_f32(const reslic_gc_bwd_desc* d, void* stream) {
  return reslic::gc_bwd_launch(d, static_cast<cudaStream_t>(stream));
}

int reslic_build_indexes_f32(const float* sigma, int64_t n, float scale_bound, const float* scale_table,
                             int32_t table_len, int32_t* idx, void* stream) {
  reslic_gc_desc d;
  std::memset(&d, 0, sizeof(d));
  d.struct_size = sizeof(d);
  d.sigma = sigma; d.sigma_bs = n; d.B = 1; d.n = n; d.mode = RESLIC_Q_DEQUANTIZE;
  d.scale_bound = scale_bound; d.scale_table = scale_table; d.table_len = table_len;
  d.idx = idx; d.idx_bs = n;
  if (n > 0 && !idx) return reslic::set_error(RESLIC_ERR_ARG, "build_indexes: idx is null");
  return reslic::gc_fwd_launch(&d, static_cast<cudaStream_t>(stream));
}

int reslic_lrp_tail_f32(float* y_hat, int64_t y_hat_bs, const float* lrp, int64_t lrp_bs, int64_t B, int64_t n,
                        void* stream) {
  return reslic::lrp_tail_launch(y_hat, y_hat_bs, lrp, lrp_bs, B, n, static_cast<cudaStream_t>(stream));
}

int reslic_dequantize_f32(const int32_t* sym, const float* mu, int64_t n, float* out, void* stream) {
  return reslic::dequantize_launch(sym, mu, n, out, static_cast<cudaStream_t>(stream));
}

int reslic_stanh_gc_fwd_f32(const reslic_stanh_gc_desc* d, void* stream) {
  return reslic::stanh_gc_fwd_launch(d, static_cast<cudaStream_t>(stream));
}
int reslic_eb_stanh_fwd_f32(const reslic_eb_stanh_desc* d, void* stream) {
  return reslic::eb_stanh_fwd_launch(d, static_cast<cudaStream_t>(stream));
}
int reslic_stanh_gc_bwd_f32(const reslic_stanh_gc_bwd_desc* d, void* stream) {
  return reslic::stanh_gc_bwd_launch(d, static_cast<cudaStream_t>(stream));
}
int64_t reslic_stanh_gap_workspace_bytes(void) { return 16 + static_cast<int64_t>(reslic::sm_count()) * 8 * 16; }
int reslic_stanh_act_f32(const float* x, int64_t n, const reslic_stanh_tables* t, float* out_soft, float* out_hard,
                         double* gap2, void* gap_workspace, int64_t gap_workspace_bytes, void* stream) {
  return reslic::stanh_act_launch(x, n, t, out_soft, out_hard, gap2, gap_workspace, gap_workspace_bytes,
                                  static_cast<cudaStream_t>(stream));
}

int reslic_eb_bwd_f32(const reslic_eb_bwd_desc* d, void* stream) {
  return reslic::eb_bwd_launch(d, static_cast<cudaStream_t>(stream));
}
int64_t reslic_eb_bwd_workspace_bytes(int64_t C) { return C > 0 ? reslic::eb_bwd_workspace_bytes(C) : 0; }

int reslic_rans_slots_u32(const int32_t* symbols, const int32_t* indexes, int64_t n, const int32_t* cdfs,
                          int32_t n_cdfs, int32_t cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets,
                          uint32_t* slots, int32_t* esc_pos, int64_t* esc_raw, int64_t esc_capacity,
                          int32_t* status, void* stream) {
  return reslic::rans_slots_launch(symbols, indexes, n, cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets, slots, esc_pos,
                                   esc_raw, esc_capacity, status, static_cast<cudaStream_t>(stream));
}

int reslic_rate_from_likelihood_f32(const float* lik, int64_t lik_bs, int64_t B, int64_t n, double* bits,
                                    int32_t bits_accumulate, void* workspace, int64_t workspace_bytes, void* stream) {
  return reslic::rate_from_lik_launch(lik, lik_bs, B, n, bits, bits_accumulate, workspace, workspace_bytes,
                                      static_cast<cudaStream_t>(stream));
}

int reslic_eb_fwd_f32(const reslic_eb_desc* d, void* stream) {
  return reslic::eb_fwd_launch(d, static_cast<cudaStream_t>(stream));
}
int reslic_eb_build_lut_f32(const reslic_eb_desc* d, float* lut, void* stream) {
  return reslic::eb_build_lut_launch(d, lut, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
