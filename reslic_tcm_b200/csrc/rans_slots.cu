// rans_slots.cu — device-side front end of the rANS coder (sm_100a).
//
// compressai.ans.encode_with_indexes (reference call sites src/models/reference/tcm.py:522,564-565;
// src/entropy_models/adaptive_gaussian_conditional.py:291-299) starts, per symbol, with
//     value = symbol - offsets[index];  clamp to the escape slot cdf_sizes[index] - 2;
//     (start, range) = (cdf[value], cdf[value + 1] - cdf[value])
// and only then runs the sequential state update.  That lookup is elementwise: here it runs on the GPU right
// behind the fused kernel that produced symbols and indexes, and what crosses PCIe is ONE packed 32-bit slot
// (start << 16 | range, both < 2^16 at 16-bit precision) per symbol instead of two int32 — the host coder
// (rans.cpp, reslic_rans_encoder_push_slots) is left with the state update.  Out-of-range symbols take the
// escape slot and are listed (position, raw bypass value) through one atomic counter; they are rare by
// construction of the tables (tail mass 1e-9) and the host sorts the list by position.
#include "common.cuh"
#include "reslic_internal.h"

namespace reslic {

struct SlotParams {
  const int32_t* sym; const int32_t* idx; int64_t n;
  const int32_t* cdfs; int n_cdfs, stride; const int32_t* sizes; const int32_t* offsets;
  uint32_t* slots; int32_t* esc_pos; long long* esc_raw; long long cap;
  int32_t* status;      // [0] number of escapes (may exceed cap: the list is then incomplete), [1] error bits
};

constexpr int kSlotMaxTables = 1024;     // sizes/offsets staged in shared memory up to this many CDFs

__device__ __forceinline__ uint32_t slot_of(const SlotParams& p, const int32_t* s_sizes, const int32_t* s_offsets,
                                            int32_t sy, int32_t ci, int64_t pos, int& err) {
  if (ci < 0 || ci >= p.n_cdfs) { err |= 1; return 0u; }
  const int32_t* __restrict__ cdf = p.cdfs + static_cast<int64_t>(ci) * p.stride;
  const int32_t max_value = s_sizes[ci] - 2;
  if (max_value < 1 || max_value + 2 > p.stride) { err |= 2; return 0u; }      // cdf length outside 3..stride: never read
  long long value = static_cast<long long>(sy) - s_offsets[ci];
  long long raw = -1;
  if (value < 0) { raw = -2 * value - 1; value = max_value; }
  else if (value >= max_value) { raw = 2 * (value - max_value); value = max_value; }
  const int32_t start = __ldg(cdf + value), next = __ldg(cdf + value + 1);
  const int32_t range = next - start;
  if (range <= 0 || range > 0xffff || start < 0 || start > 0xffff) { err |= 2; return 0u; }
  if (raw >= 0) {
    const int at = atomicAdd(p.status, 1);
    if (at < p.cap) { p.esc_pos[at] = static_cast<int32_t>(pos); p.esc_raw[at] = raw; }
  }
  return (static_cast<uint32_t>(start) << 16) | static_cast<uint32_t>(range);
}

template <bool VEC>
__global__ void __launch_bounds__(kThreads) rans_slots_kernel(const SlotParams p) {
  __shared__ int32_t s_sizes[kSlotMaxTables];
  __shared__ int32_t s_offsets[kSlotMaxTables];
  for (int i = threadIdx.x; i < p.n_cdfs; i += kThreads) { s_sizes[i] = p.sizes[i]; s_offsets[i] = p.offsets[i]; }
  __syncthreads();
  int err = 0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  const int64_t first = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x;
  const int64_t n_vec = VEC ? (p.n >> 2) : 0;
  for (int64_t g = first; g < n_vec; g += stride) {
    const int4 s = __ldcs(reinterpret_cast<const int4*>(p.sym) + g);
    const int4 c = __ldcs(reinterpret_cast<const int4*>(p.idx) + g);
    uint4 o;
    o.x = slot_of(p, s_sizes, s_offsets, s.x, c.x, 4 * g, err);
    o.y = slot_of(p, s_sizes, s_offsets, s.y, c.y, 4 * g + 1, err);
    o.z = slot_of(p, s_sizes, s_offsets, s.z, c.z, 4 * g + 2, err);
    o.w = slot_of(p, s_sizes, s_offsets, s.w, c.w, 4 * g + 3, err);
    __stcs(reinterpret_cast<uint4*>(p.slots) + g, o);
  }
  for (int64_t i = 4 * n_vec + first; i < p.n; i += stride)
    p.slots[i] = slot_of(p, s_sizes, s_offsets, p.sym[i], p.idx[i], i, err);
  if (err) atomicOr(p.status + 1, err);
}

int rans_slots_launch(const int32_t* symbols, const int32_t* indexes, int64_t n, const int32_t* cdfs, int32_t n_cdfs,
                      int32_t cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets, uint32_t* slots,
                      int32_t* esc_pos, int64_t* esc_raw, int64_t esc_capacity, int32_t* status, cudaStream_t st) {
  if (n < 0 || esc_capacity < 0) return set_error(RESLIC_ERR_ARG, "rans_slots: negative size");
  if (!status) return set_error(RESLIC_ERR_ARG, "rans_slots: status is null");
  cudaError_t e = cudaMemsetAsync(status, 0, 2 * sizeof(int32_t), st);
  if (e != cudaSuccess) return set_cuda_error(e, "rans_slots memset");
  if (n == 0) return RESLIC_OK;
  if (n >= (1LL << 31)) return set_error(RESLIC_ERR_ARG, "rans_slots: more than 2^31 symbols per call");
  if (!symbols || !indexes || !slots) return set_error(RESLIC_ERR_ARG, "rans_slots: symbols, indexes or slots is null");
  if (!cdfs || !cdf_sizes || !offsets || n_cdfs < 1 || n_cdfs > kSlotMaxTables || cdf_stride < 3)
    return set_error(RESLIC_ERR_ARG, "rans_slots: CDF tables missing or more than 1024 of them");
  if (esc_capacity > 0 && (!esc_pos || !esc_raw)) return set_error(RESLIC_ERR_ARG, "rans_slots: escape list is null");
  SlotParams p{symbols, indexes, n, cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets, slots, esc_pos,
               reinterpret_cast<long long*>(esc_raw), static_cast<long long>(esc_capacity), status};
  const bool vec = ((reinterpret_cast<uintptr_t>(symbols) | reinterpret_cast<uintptr_t>(indexes) |
                     reinterpret_cast<uintptr_t>(slots)) & 15u) == 0;
  int64_t grid = ((vec ? (n + 3) / 4 : n) + kThreads - 1) / kThreads;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (grid > cap) grid = cap;
  if (vec) rans_slots_kernel<true><<<static_cast<int>(grid), kThreads, 0, st>>>(p);
  else rans_slots_kernel<false><<<static_cast<int>(grid), kThreads, 0, st>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "rans_slots launch");
  return RESLIC_OK;
}

}  // namespace reslic
