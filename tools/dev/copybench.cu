// Ceiling probe (development tool): what does a kernel with the GC kernel's traffic shape
// (3 streaming reads + 4 streaming writes of 16 B per thread-group, same launch size, PDL, graph
// replay over rotating buffer sets) reach when it does no math at all?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
struct P { const float4* a; const float4* b; const float4* c; float4* o0; float4* o1; float4* o2; float4* o3; unsigned groups, q, r; };
__device__ __forceinline__ void pdl() { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void emit(const P& p, unsigned g, float4 x, float4 y, float4 z) {
  __stcs(p.o0 + g, make_float4(x.x + y.x, x.y, x.z, x.w)); __stcs(p.o1 + g, make_float4(y.x + z.x, y.y, y.z, y.w));
  __stcs(p.o2 + g, make_float4(z.x + x.x, z.y, z.z, z.w)); __stcs(p.o3 + g, make_float4(x.w + y.w, y.z, z.y, x.x));
}
// 1: one group per thread
__global__ void __launch_bounds__(256) k_simple(const P p) {
  pdl();
  unsigned g = blockIdx.x * 256u + threadIdx.x;
  if (g < p.groups) emit(p, g, __ldcs(p.a + g), __ldcs(p.b + g), __ldcs(p.c + g));
}
// 2: persistent, contiguous tile range, D tiles of loads in flight per thread (register ring, unrolled)
template <int D>
__global__ void __launch_bounds__(256, 4) k_persist(const P p) {
  pdl();
  unsigned t = blockIdx.x * p.q + min(blockIdx.x, p.r);
  const unsigned t_end = t + p.q + (blockIdx.x < p.r ? 1u : 0u);
  float4 x[D], y[D], z[D];
#pragma unroll
  for (int d = 0; d < D; ++d) { unsigned g = (t + d) * 256u + threadIdx.x; if (t + d < t_end && g < p.groups) { x[d] = __ldcs(p.a + g); y[d] = __ldcs(p.b + g); z[d] = __ldcs(p.c + g); } }
  for (; t < t_end; t += D) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      unsigned g = (t + d) * 256u + threadIdx.x;
      if (t + d < t_end && g < p.groups) {
        float4 cx = x[d], cy = y[d], cz = z[d];
        unsigned gn = (t + d + D) * 256u + threadIdx.x;
        if (t + d + D < t_end && gn < p.groups) { x[d] = __ldcs(p.a + gn); y[d] = __ldcs(p.b + gn); z[d] = __ldcs(p.c + gn); }
        emit(p, g, cx, cy, cz);
      }
    }
  }
}
// 2b: persistent, grid-stride tile order (concurrent CTAs share one contiguous window), D-deep register ring
template <int D>
__global__ void __launch_bounds__(256, 4) k_stride(const P p) {
  extern __shared__ float4 dummy[];
  pdl();
  const unsigned G = gridDim.x, tiles = (p.groups + 255u) / 256u;
  unsigned t = blockIdx.x;
  float4 x[D], y[D], z[D];
#pragma unroll
  for (int d = 0; d < D; ++d) { unsigned tt = t + d * G; unsigned g = tt * 256u + threadIdx.x; if (tt < tiles && g < p.groups) { x[d] = __ldcs(p.a + g); y[d] = __ldcs(p.b + g); z[d] = __ldcs(p.c + g); } }
  for (; t < tiles; t += D * G) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      unsigned tt = t + d * G; unsigned g = tt * 256u + threadIdx.x;
      if (tt < tiles && g < p.groups) {
        float4 cx = x[d], cy = y[d], cz = z[d];
        unsigned tn = tt + D * G; unsigned gn = tn * 256u + threadIdx.x;
        if (tn < tiles && gn < p.groups) { x[d] = __ldcs(p.a + gn); y[d] = __ldcs(p.b + gn); z[d] = __ldcs(p.c + gn); }
        emit(p, g, cx, cy, cz);
      }
    }
  }
}
__global__ void __launch_bounds__(256) k_simple_smem(const P p) {
  extern __shared__ float4 dummy[];
  pdl();
  unsigned g = blockIdx.x * 256u + threadIdx.x;
  if (g < p.groups) emit(p, g, __ldcs(p.a + g), __ldcs(p.b + g), __ldcs(p.c + g));
}
// 3: per-thread cp.async ring in shared memory, S stages
template <int S>
__global__ void __launch_bounds__(256, 4) k_cpasync(const P p) {
  extern __shared__ float4 ring[];   // [S][3][256]
  pdl();
  unsigned t0 = blockIdx.x * p.q + min(blockIdx.x, p.r);
  const unsigned nt = p.q + (blockIdx.x < p.r ? 1u : 0u);
  auto issue = [&](unsigned k) {
    if (k < nt) {
      unsigned g = (t0 + k) * 256u + threadIdx.x;
      if (g < p.groups) {
        float4* dst = ring + (k % S) * 768 + threadIdx.x;
        unsigned s0 = (unsigned)__cvta_generic_to_shared(dst);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s0), "l"(p.a + g));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s0 + 4096u), "l"(p.b + g));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s0 + 8192u), "l"(p.c + g));
      }
    }
    asm volatile("cp.async.commit_group;");
  };
#pragma unroll
  for (int d = 0; d < S - 1; ++d) issue(d);
  for (unsigned k = 0; k < nt; ++k) {
    issue(k + S - 1);
    asm volatile("cp.async.wait_group %0;" :: "n"(S - 1));
    unsigned g = (t0 + k) * 256u + threadIdx.x;
    if (g < p.groups) {
      const float4* src = ring + (k % S) * 768 + threadIdx.x;
      emit(p, g, src[0], src[256], src[512]);
    }
  }
}
int main(int argc, char** argv) {
  int B = argc > 1 ? atoi(argv[1]) : 24; long n = argc > 2 ? atol(argv[2]) : 98304; int reps = argc > 3 ? atoi(argv[3]) : 40;
  size_t N = (size_t)B * n; unsigned groups = (unsigned)(N / 4);
  const int nset = 3;
  std::vector<float*> buf(nset * 7);
  for (auto& b : buf) { CK(cudaMalloc(&b, N * 4)); CK(cudaMemset(b, 0, N * 4)); }
  cudaStream_t st; CK(cudaStreamCreate(&st));
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  auto run = [&](const char* name, auto kernel, int mode, size_t smem) {
    unsigned tiles = (groups + 255) / 256;
    if (smem) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, smem));
    unsigned grid = mode == 0 ? tiles : std::min<unsigned>(tiles, occ * sms);
    auto launch = [&](int s) {
      P p{(float4*)buf[s * 7], (float4*)buf[s * 7 + 1], (float4*)buf[s * 7 + 2], (float4*)buf[s * 7 + 3], (float4*)buf[s * 7 + 4],
          (float4*)buf[s * 7 + 5], (float4*)buf[s * 7 + 6], groups, tiles / grid, tiles % grid};
      cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = st;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      CK(cudaLaunchKernelEx(&cfg, kernel, p));
    };
    for (int i = 0; i < 6; ++i) launch(i % nset);
    CK(cudaStreamSynchronize(st));
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
    for (int i = 0; i < 5 * nset; ++i) launch(i % nset);
    CK(cudaStreamEndCapture(st, &g)); CK(cudaGraphInstantiate(&ge, g, 0));
    CK(cudaGraphLaunch(ge, st)); CK(cudaStreamSynchronize(st));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    CK(cudaEventRecord(a, st));
    for (int r = 0; r < reps; ++r) CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(b, st)); CK(cudaStreamSynchronize(st));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    double us = ms * 1e3 / (reps * 5.0 * nset), gbs = N * 28.0 / us * 1e-3;
    printf("%-18s grid=%5u occ=%d : %7.2f us/launch  %7.1f GB/s  %.1f%% of 6537.6\n", name, grid, occ, us, gbs, gbs / 65.376);
    CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g));
  };
  printf("B=%d n=%ld (%.1f MB per launch)\n", B, n, N * 28e-6);
  run("simple", k_simple, 0, 0);
  run("simple occ4", k_simple_smem, 0, 50 * 1024);
  run("simple occ3", k_simple_smem, 0, 70 * 1024);
  run("stride D=1", k_stride<1>, 1, 50 * 1024);
  run("stride D=2", k_stride<2>, 1, 50 * 1024);
  run("stride D=2 occ3", k_stride<2>, 1, 70 * 1024);
  run("stride D=1 occ8", k_stride<1>, 1, 0);
  run("persist D=1", k_persist<1>, 1, 0);
  run("persist D=2", k_persist<2>, 1, 0);
  run("persist D=4", k_persist<4>, 1, 0);
  run("cpasync S=2", k_cpasync<2>, 1, 2 * 12288);
  run("cpasync S=4", k_cpasync<4>, 1, 4 * 12288);
  run("cpasync S=6", k_cpasync<6>, 1, 6 * 12288);
  return 0;
}
