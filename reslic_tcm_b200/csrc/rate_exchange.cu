// rate_exchange.cu — reader side of the multi-GPU rate exchange and the peer-accessible buffers it lives in.
//
// The WRITER is not a kernel of its own: the launch that collects a batch's rate publishes the packed row into every
// rank's buffer (rate_publish, common.cuh).  Here: the tiny kernel that adds the `world` rows of a step in rank order
// once their tagged cells have arrived, and the CUDA-IPC plumbing that maps the ranks' buffers into each other's processes.
// Replaces the gather of whole likelihood tensors to GPU 0 that the reference's nn.DataParallel performs before its
// loss reduces them (src/utils/helper.py:106-113, src/train.py:168-169, src/training/loss.py:24-27).
#include <cstring>
#include "common.cuh"
#include "reslic_internal.h"

namespace reslic {

// one 16-byte cell {value, tag}, read in one transaction
__device__ __forceinline__ void ld_cell(const void* p, double& v, unsigned long long& tag) {
  long long a;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(tag) : "l"(p) : "memory");
  v = __longlong_as_double(a);
}
__device__ __forceinline__ unsigned long long gtimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// one thread per step; the waits are on LOCAL memory (the peers store into this rank's buffer)
__global__ void rate_exchange_read_kernel(const char* base, int world, int ring, const unsigned long long* cursor,
                                          long long first_step, int n_steps, double* out, int* status) {
  griddep_wait();
  griddep_launch_dependents();
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_steps) return;
  const long long abs_step = first_step + s + (cursor ? static_cast<long long>(*cursor) : 0ll);
  if (abs_step < 0) {                                   // before the first step (first replay of a graph that reads behind)
    for (int j = 0; j < 4; ++j) out[static_cast<size_t>(s) * 4 + j] = 0.0;
    return;
  }
  const unsigned long long step = static_cast<unsigned long long>(abs_step);
  const char* slot = base + static_cast<size_t>(step % static_cast<unsigned long long>(ring)) * world * 64;
  const unsigned long long t0 = gtimer_ns();
  int err = 0;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int r = 0; r < world && !err; ++r) {            // rank order: every rank adds in the same order
    for (int j = 0; j < 4 && !err; ++j) {
      double v; unsigned long long tag;
      for (;;) {
        ld_cell(slot + r * 64 + j * 16, v, tag);
        if (tag >= step + 1ull) break;
        if (gtimer_ns() - t0 > 2000000000ull) { err = 1; break; }      // a rank never published: report, do not hang
        __nanosleep(200);
      }
      if (!err && tag != step + 1ull) err = 2;                         // overwritten: a rank ran >= ring steps ahead
      acc[j] += v;
    }
  }
  if (err) {
    atomicOr(status, err);
    for (int j = 0; j < 4; ++j) acc[j] = __longlong_as_double(0x7ff8000000000000LL);
  }
  for (int j = 0; j < 4; ++j) out[static_cast<size_t>(s) * 4 + j] = acc[j];
}

// one CTA: sum of bits[0..B) (exact: multiples of 2^-16), then the row's 4 * world peer stores
__global__ void __launch_bounds__(128) rate_exchange_publish_kernel(const double* bits, long long B, const RateEx ex) {
  __shared__ double s_part[4];
  griddep_wait();
  griddep_launch_dependents();
  double v = 0.0;
  for (long long i = threadIdx.x; i < B; i += blockDim.x) v += bits[i];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x != 0) return;
  const double total = (s_part[0] + s_part[1]) + (s_part[2] + s_part[3]);
  const unsigned long long step = static_cast<unsigned long long>(ex.step_rel) +
                                  (ex.cursor ? *reinterpret_cast<const volatile unsigned long long*>(ex.cursor) : 0ull);
  const double extra = ex.extra ? *reinterpret_cast<const volatile double*>(ex.extra) : 0.0;
  const size_t cell = (static_cast<size_t>(step % static_cast<unsigned long long>(ex.ring)) * ex.world + ex.rank) * 64;
  for (int p = 0; p < ex.world; ++p) {
    char* row = static_cast<char*>(ex.peer[p]) + cell;
    st_cell(row, total, step + 1ull);
    st_cell(row + 16, extra, step + 1ull);
    st_cell(row + 32, ex.pixels, step + 1ull);
    st_cell(row + 48, ex.images, step + 1ull);
  }
}

}  // namespace reslic

extern "C" {

int64_t reslic_rate_exchange_bytes(int32_t world, int32_t ring) {
  if (world < 1 || ring < 1) return 0;
  return static_cast<int64_t>(ring) * world * 64;      // four 16-byte cells {value, step + 1} per (slot, rank)
}

int reslic_rate_exchange_read_f64(const void* own_base, int32_t world, int32_t ring, const unsigned long long* cursor,
                                  int64_t first_step, int32_t n_steps, double* out, int32_t* status, void* stream) {
  using namespace reslic;
  if (n_steps < 0 || (first_step < 0 && !cursor)) return set_error(RESLIC_ERR_ARG, "rate_exchange_read: negative step");
  if (n_steps == 0) return RESLIC_OK;
  if (!own_base || !out || !status) return set_error(RESLIC_ERR_ARG, "rate_exchange_read: null pointer");
  if (world < 1 || world > 64 || ring < 1 || n_steps > ring)
    return set_error(RESLIC_ERR_ARG, "rate_exchange_read: world outside 1..64, ring < 1 or more steps than ring slots");
  rate_exchange_read_kernel<<<(n_steps + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const char*>(own_base), world, ring, cursor, first_step, n_steps, out, status);
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_cuda_error(err, "rate_exchange_read launch");
  return RESLIC_OK;
}

int reslic_rate_exchange_publish_f64(const reslic_rate_exchange* x, const double* bits, int64_t B, void* stream) {
  using namespace reslic;
  if (!x || !bits || B < 1) return set_error(RESLIC_ERR_ARG, "rate_exchange_publish: null argument or B < 1");
  if (x->struct_size != sizeof(reslic_rate_exchange))
    return set_error(RESLIC_ERR_ARG, "rate_exchange_publish: struct_size != sizeof(reslic_rate_exchange) (ABI mismatch)");
  if (x->world < 1 || x->world > 64 || x->rank < 0 || x->rank >= x->world || x->ring < 1 || !x->peer_base || x->step < 0)
    return set_error(RESLIC_ERR_ARG, "rate_exchange_publish: bad world/rank/ring/step or null peer_base");
  RateEx ex{};
  ex.peer = x->peer_base; ex.cursor = x->cursor; ex.step_rel = x->step; ex.extra = x->extra; ex.pixels = x->pixels;
  ex.images = x->images; ex.world = x->world; ex.rank = x->rank; ex.ring = x->ring;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(1);
  cfg.blockDim = dim3(128);
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = gc_tuning().pdl ? 1 : 0;
  const cudaError_t err = cudaLaunchKernelEx(&cfg, rate_exchange_publish_kernel, bits, static_cast<long long>(B), ex);
  if (err != cudaSuccess) return set_cuda_error(err, "rate_exchange_publish launch");
  return RESLIC_OK;
}

int reslic_peer_buffer_create(int64_t bytes, void** dptr, uint8_t* handle) {
  using namespace reslic;
  static_assert(sizeof(cudaIpcMemHandle_t) == RESLIC_PEER_HANDLE_BYTES, "handle size");
  if (bytes <= 0 || !dptr || !handle) return set_error(RESLIC_ERR_ARG, "peer_buffer_create: bad argument");
  void* p = nullptr;
  cudaError_t err = cudaMalloc(&p, static_cast<size_t>(bytes));
  if (err != cudaSuccess) return set_cuda_error(err, "peer_buffer_create: cudaMalloc");
  err = cudaMemset(p, 0, static_cast<size_t>(bytes));
  if (err == cudaSuccess) err = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (err == cudaSuccess) err = cudaIpcGetMemHandle(&h, p);
  if (err != cudaSuccess) { cudaFree(p); return set_cuda_error(err, "peer_buffer_create: cudaIpcGetMemHandle"); }
  std::memcpy(handle, &h, sizeof(h));
  *dptr = p;
  return RESLIC_OK;
}
int reslic_peer_buffer_open(const uint8_t* handle, void** dptr) {
  using namespace reslic;
  if (!handle || !dptr) return set_error(RESLIC_ERR_ARG, "peer_buffer_open: null argument");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  const cudaError_t err = cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (err != cudaSuccess) return set_cuda_error(err, "peer_buffer_open: cudaIpcOpenMemHandle");
  return RESLIC_OK;
}
int reslic_peer_buffer_close(void* dptr) {
  using namespace reslic;
  if (!dptr) return RESLIC_OK;
  const cudaError_t err = cudaIpcCloseMemHandle(dptr);
  if (err != cudaSuccess) return set_cuda_error(err, "peer_buffer_close");
  return RESLIC_OK;
}
int reslic_peer_buffer_destroy(void* dptr) {
  using namespace reslic;
  if (!dptr) return RESLIC_OK;
  const cudaError_t err = cudaFree(dptr);
  if (err != cudaSuccess) return set_cuda_error(err, "peer_buffer_destroy");
  return RESLIC_OK;
}

}  // extern "C"
