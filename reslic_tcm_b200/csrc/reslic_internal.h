// reslic_internal.h — host-side glue shared by the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/reslic_b200.h"

namespace reslic {

int set_error(int code, const char* msg);                 // returns code
int set_cuda_error(cudaError_t err, const char* where);   // returns (int)err
int sm_count();
struct GcTuning { int ctas_per_sm; int pdl; int min_ctas; int balance; };
const GcTuning& gc_tuning();                              // launch-shape knobs (env overridable)
// Rate outputs of a *_fwd descriptor: requested when `bits` is given or the mode is RESLIC_RATE_DEFERRED.
inline bool rate_requested(const double* bits, int32_t mode) { return bits != nullptr || mode == 2; }
// Validates mode/workspace and fills the kernel-side fields; in deferred mode *bits_out is a non-null
// sentinel (never dereferenced) so that kernels keep one `if (p.bits)` gate.  Returns RESLIC_OK or an error.
int rate_setup(const char* who, double* bits, int32_t mode, void* workspace, int64_t workspace_bytes, int64_t B,
               double** bits_out, int* mode_out, unsigned long long** ws_out);
int rate_finalize_launch(void* workspace, int64_t workspace_bytes, int64_t B, double* bits, int32_t accumulate,
                         cudaStream_t st);
int math_mode();                                          // RESLIC_MATH_*                                           // SMs of the current device (cached per device)

int gc_fwd_launch(const reslic_gc_desc* d, cudaStream_t st);
int gc_bwd_launch(const reslic_gc_bwd_desc* d, cudaStream_t st);
int eb_bwd_launch(const reslic_eb_bwd_desc* d, cudaStream_t st);
int64_t eb_bwd_workspace_bytes(int64_t C);
int eb_fwd_launch(const reslic_eb_desc* d, cudaStream_t st);
int eb_build_lut_launch(const reslic_eb_desc* d, float* lut, cudaStream_t st);
int stanh_gc_fwd_launch(const reslic_stanh_gc_desc* d, cudaStream_t st);
int eb_stanh_fwd_launch(const reslic_eb_stanh_desc* d, cudaStream_t st);
int stanh_gc_bwd_launch(const reslic_stanh_gc_bwd_desc* d, cudaStream_t st);
int stanh_act_launch(const float* x, int64_t n, const reslic_stanh_tables* t, float* out_soft, float* out_hard,
                     double* gap, void* workspace, int64_t workspace_bytes, cudaStream_t st);
int lrp_tail_launch(float* y_hat, int64_t y_bs, const float* lrp, int64_t l_bs, int64_t B, int64_t n, cudaStream_t st);
int rans_slots_launch(const int32_t* symbols, const int32_t* indexes, int64_t n, const int32_t* cdfs, int32_t n_cdfs,
                      int32_t cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets, uint32_t* slots,
                      int32_t* esc_pos, int64_t* esc_raw, int64_t esc_capacity, int32_t* status, cudaStream_t st);
int rate_from_lik_launch(const float* lik, int64_t lik_bs, int64_t B, int64_t n, double* bits, int32_t bits_accumulate,
                         void* workspace, int64_t workspace_bytes, cudaStream_t st);
int dequantize_launch(const int32_t* sym, const float* mu, int64_t n, float* out, cudaStream_t st);

}  // namespace reslic
