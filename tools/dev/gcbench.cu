// Micro-benchmark harness for the fused GC kernel through the C ABI (development tool).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include "../include/reslic_b200.h"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__global__ void fill(float* y, float* mu, float* sg, size_t n, unsigned seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned h = (unsigned)i * 2654435761u + seed; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    float u1 = (h & 0xffff) / 65536.0f, u2 = ((h >> 16) & 0xffff) / 65536.0f;
    mu[i] = 4.0f * u1 - 2.0f;
    sg[i] = expf(-3.0f + 7.16f * u2);
    h = h * 3266489917u + 7u; h ^= h >> 16;
    float u3 = (h & 0xffffff) / 16777216.0f - 0.5f;
    y[i] = mu[i] + sg[i] * 3.0f * u3;
  }
}
int main(int argc, char** argv) {
  int B = argc > 1 ? atoi(argv[1]) : 24; long n = argc > 2 ? atol(argv[2]) : 64 * 48 * 32;
  int with_idx = argc > 3 ? atoi(argv[3]) : 1; int noise = argc > 4 ? atoi(argv[4]) : 0;
  int nset = 3, reps = argc > 5 ? atoi(argv[5]) : 50; int deferred = argc > 6 ? atoi(argv[6]) : 0; int smul = argc > 7 ? atoi(argv[7]) : 1; int prefetch = argc > 8 ? atoi(argv[8]) : 0;  // smul: batch stride = smul * n (a channel slice of a wider tensor)
  size_t N = (size_t)B * n; size_t NA = N * smul; long bs = n * smul;
  std::vector<float*> y(nset), mu(nset), sg(nset), yh(nset), lk(nset), nz(nset); std::vector<int*> sym(nset), idx(nset);
  float tabh[64]; for (int i = 0; i < 64; ++i) tabh[i] = expf(logf(0.11f) + i * (logf(256.f) - logf(0.11f)) / 63.f);
  float* tab; CK(cudaMalloc(&tab, 256)); CK(cudaMemcpy(tab, tabh, 256, cudaMemcpyHostToDevice));
  double* bits; CK(cudaMalloc(&bits, B * 8));
  void* ws; size_t wsb = reslic_workspace_bytes(B); CK(cudaMalloc(&ws, wsb)); CK(cudaMemset(ws, 0, wsb));
  for (int s = 0; s < nset; ++s) {
    CK(cudaMalloc(&y[s], NA * 4)); CK(cudaMalloc(&mu[s], NA * 4)); CK(cudaMalloc(&sg[s], NA * 4));
    CK(cudaMalloc(&yh[s], NA * 4)); CK(cudaMalloc(&lk[s], NA * 4)); CK(cudaMalloc(&sym[s], NA * 4)); CK(cudaMalloc(&idx[s], NA * 4));
    CK(cudaMalloc(&nz[s], NA * 4));
    fill<<<1024, 256>>>(y[s], mu[s], sg[s], NA, 17u + s);
  }
  CK(cudaDeviceSynchronize());
  cudaStream_t st; CK(cudaStreamCreate(&st));
  long slice_ctr = 0;
  auto launch = [&](int s) {
    reslic_gc_desc d; memset(&d, 0, sizeof(d));
    const long so = (long)(smul > 1 ? ((slice_ctr++) % smul) : 0) * n;   // walk the slices of the wide tensor like TCM does
    d.y = y[s] + so; d.y_bs = bs; d.mu = mu[s] + so; d.mu_bs = bs; d.sigma = sg[s] + so; d.sigma_bs = bs; d.B = B; d.n = n;
    d.mode = noise ? RESLIC_Q_NOISE : RESLIC_Q_DEQUANTIZE; d.scale_bound = 0.11f; d.likelihood_bound = 1e-9f;
    d.ste = yh[s] + so; d.ste_bs = bs; d.lik = lk[s] + so; d.lik_bs = bs;
    if (noise) { d.yhat = nz[s] + so; d.yhat_bs = bs; }
    if (with_idx) { d.scale_table = tab; d.table_len = 64; d.sym = sym[s] + so; d.sym_bs = bs; d.idx = idx[s] + so; d.idx_bs = bs; }
    if (prefetch) {   // the y the next launch of this sequence reads
      const long so_next = (long)(smul > 1 ? (slice_ctr % smul) : 0) * n;
      d.next_y = y[(s + 1) % nset] + so_next; d.next_y_bs = bs;
    }
    d.bits = deferred ? nullptr : bits; d.bits_accumulate = deferred ? RESLIC_RATE_DEFERRED : 0; d.workspace = ws; d.workspace_bytes = wsb; d.philox_seed = 1;
    int rc = reslic_gc_fwd_f32(&d, st);
    if (rc) { printf("launch failed %d %s\n", rc, reslic_last_error()); exit(1); }
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
  double us = ms * 1e3 / (reps * 5 * nset);
  int bpe = 12 + 8 + (with_idx ? 8 : 0) + (noise ? 4 : 0);
  double gbs = (double)N * bpe / (us * 1e-6) / 1e9;
  if (deferred) {  // drain what the timed launches accumulated, then one clean launch
    if (reslic_rate_finalize_f64(ws, wsb, B, bits, 0, st)) { printf("finalize failed %s\n", reslic_last_error()); exit(1); }
    launch(0);
    if (reslic_rate_finalize_f64(ws, wsb, B, bits, 0, st)) { printf("finalize failed %s\n", reslic_last_error()); exit(1); }
    CK(cudaStreamSynchronize(st));
  }
  double hb[4]; CK(cudaMemcpy(hb, bits, sizeof(double) * (B < 4 ? B : 4), cudaMemcpyDeviceToHost));
  printf("B=%d n=%ld idx=%d noise=%d def=%d : %.2f us/launch  %.1f GB/s  %.1f%% of 6537.6  (%.1f Gelem/s) bits0=%.3f\n", B, n, with_idx, noise, deferred, us, gbs, 100 * gbs / 6537.6, N / us / 1e3, hb[0]);
  return 0;
}
