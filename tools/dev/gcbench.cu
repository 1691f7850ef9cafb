// Micro-benchmark harness for the fused GC kernel through the C ABI (development tool, not product code).
//
//   gcbench B=24 n=98304 idx=1 noise=0 reps=40 rate=2 smul=5 prefetch=1 chains=1 steps=6 nset=3
//
// A "step" is TCM's slice loop over one buffer set: `smul` launches (one per channel slice of a tensor whose
// batch stride is smul*n), rate mode `rate` (0 immediate, 2 deferred with the last launch of the step
// collecting).  `steps` consecutive steps are captured in ONE graph; with chains > 1 the steps are dealt
// round-robin onto that many forked capture streams (independent batches in flight at once, each chain
// still a dependent PDL sequence), which needs nset >= chains buffer sets (one rate workspace per set).
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/reslic_b200.h"
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
#ifdef RESLIC_TRACE
unsigned long long* g_trace_buf = nullptr; int g_trace_on = 0;
#endif
static long arg(int argc, char** argv, const char* key, long dflt) {
  const size_t k = strlen(key);
  for (int i = 1; i < argc; ++i)
    if (!strncmp(argv[i], key, k) && argv[i][k] == '=') return atol(argv[i] + k + 1);
  return dflt;
}
#ifdef RESLIC_TRACE
extern "C" void reslic_debug_set_trace(unsigned long long* ptr);
constexpr int kTraceCtas = 4096;
#include <algorithm>
#endif
int main(int argc, char** argv) {
  const int B = (int)arg(argc, argv, "B", 24); const long n = arg(argc, argv, "n", 64 * 48 * 32);
  const int with_idx = (int)arg(argc, argv, "idx", 1), noise = (int)arg(argc, argv, "noise", 0);
  const int reps = (int)arg(argc, argv, "reps", 40), rate = (int)arg(argc, argv, "rate", 2);
  const int smul = (int)arg(argc, argv, "smul", 5), prefetch = (int)arg(argc, argv, "prefetch", 1);
  const int chains = (int)arg(argc, argv, "chains", 1), nset = (int)arg(argc, argv, "nset", chains > 3 ? chains : 3);
  const int steps = (int)arg(argc, argv, "steps", 2 * nset);
  if (nset < chains) { printf("nset must be >= chains\n"); return 1; }
  size_t N = (size_t)B * n; size_t NA = N * smul; long bs = n * smul;
  std::vector<float*> y(nset), mu(nset), sg(nset), yh(nset), lk(nset), nz(nset); std::vector<int*> sym(nset), idx(nset);
  std::vector<double*> bits(nset); std::vector<void*> ws(nset);
  float tabh[64]; for (int i = 0; i < 64; ++i) tabh[i] = expf(logf(0.11f) + i * (logf(256.f) - logf(0.11f)) / 63.f);
  float* tab; CK(cudaMalloc(&tab, 256)); CK(cudaMemcpy(tab, tabh, 256, cudaMemcpyHostToDevice));
  size_t wsb = reslic_workspace_bytes(B);
  for (int s = 0; s < nset; ++s) {
    CK(cudaMalloc(&y[s], NA * 4)); CK(cudaMalloc(&mu[s], NA * 4)); CK(cudaMalloc(&sg[s], NA * 4));
    CK(cudaMalloc(&yh[s], NA * 4)); CK(cudaMalloc(&lk[s], NA * 4)); CK(cudaMalloc(&sym[s], NA * 4)); CK(cudaMalloc(&idx[s], NA * 4));
    CK(cudaMalloc(&nz[s], NA * 4)); CK(cudaMalloc(&bits[s], B * 8)); CK(cudaMalloc(&ws[s], wsb)); CK(cudaMemset(ws[s], 0, wsb));
    fill<<<1024, 256>>>(y[s], mu[s], sg[s], NA, 17u + s);
  }
  CK(cudaDeviceSynchronize());
  std::vector<cudaStream_t> st(chains);
  for (auto& s : st) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  auto launch = [&](int s, int k, cudaStream_t stream) {      // slice k of set s
    reslic_gc_desc d; memset(&d, 0, sizeof(d));
    d.struct_size = sizeof(d);
    const long so = (long)k * n;
    d.y = y[s] + so; d.y_bs = bs; d.mu = mu[s] + so; d.mu_bs = bs; d.sigma = sg[s] + so; d.sigma_bs = bs; d.B = B; d.n = n;
    d.mode = noise ? RESLIC_Q_NOISE : RESLIC_Q_DEQUANTIZE; d.scale_bound = 0.11f; d.likelihood_bound = 1e-9f;
    d.ste = yh[s] + so; d.ste_bs = bs; d.lik = lk[s] + so; d.lik_bs = bs;
    if (noise) { d.yhat = nz[s] + so; d.yhat_bs = bs; }
    if (with_idx) { d.scale_table = tab; d.table_len = 64; d.sym = sym[s] + so; d.sym_bs = bs; d.idx = idx[s] + so; d.idx_bs = bs; }
    if (prefetch && k + 1 < smul) { d.next_y = y[s] + so + n; d.next_y_bs = bs; }
    const bool last = k + 1 == smul;
    if (rate == 2) { d.bits = last ? bits[s] : nullptr; d.bits_accumulate = last ? RESLIC_RATE_COLLECT : RESLIC_RATE_DEFERRED; }
    else { d.bits = bits[s]; d.bits_accumulate = k == 0 ? 0 : 1; }
    d.workspace = ws[s]; d.workspace_bytes = wsb; d.philox_seed = 1; d.philox_offset = k;
#ifdef RESLIC_TRACE
    static int trace_launch = 0;
    extern unsigned long long* g_trace_buf; extern int g_trace_on;
    if (g_trace_on) reslic_debug_set_trace(g_trace_buf + (size_t)(trace_launch++) * kTraceCtas * 6);
#endif
    int rc = reslic_gc_fwd_f32(&d, stream);
    if (rc) { printf("launch failed %d %s\n", rc, reslic_last_error()); exit(1); }
  };
  auto step = [&](int i, cudaStream_t stream) { for (int k = 0; k < smul; ++k) launch(i % nset, k, stream); };
  for (int i = 0; i < nset; ++i) step(i, st[0]);
  CK(cudaStreamSynchronize(st[0]));
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaEvent_t fork, join[16];
  CK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
  for (auto& ev : join) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
#ifdef RESLIC_TRACE
  extern unsigned long long* g_trace_buf; extern int g_trace_on;
  const size_t trace_words = (size_t)steps * smul * kTraceCtas * 6;
  CK(cudaMalloc(&g_trace_buf, trace_words * 8)); CK(cudaMemset(g_trace_buf, 0, trace_words * 8));
  g_trace_on = 1;
#endif
  CK(cudaStreamBeginCapture(st[0], cudaStreamCaptureModeGlobal));
  CK(cudaEventRecord(fork, st[0]));
  for (int c = 1; c < chains; ++c) CK(cudaStreamWaitEvent(st[c], fork, 0));
  // set s is always walked by chain s % chains, so two steps over the same buffers are stream-ordered
  for (int i = 0; i < steps; ++i) step(i, st[(i % nset) % chains]);
  for (int c = 1; c < chains; ++c) { CK(cudaEventRecord(join[c], st[c])); CK(cudaStreamWaitEvent(st[0], join[c], 0)); }
  CK(cudaStreamEndCapture(st[0], &g)); CK(cudaGraphInstantiate(&ge, g, 0));
  CK(cudaGraphLaunch(ge, st[0])); CK(cudaStreamSynchronize(st[0]));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  CK(cudaEventRecord(a, st[0]));
  for (int r = 0; r < reps; ++r) CK(cudaGraphLaunch(ge, st[0]));
  CK(cudaEventRecord(b, st[0])); CK(cudaStreamSynchronize(st[0]));
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  const double us_step = ms * 1e3 / ((double)reps * steps), us = us_step / smul;
  int bpe = 12 + 8 + (with_idx ? 8 : 0) + (noise ? 4 : 0);
  double gbs = (double)N * bpe / (us * 1e-6) / 1e9;
#ifdef RESLIC_TRACE
  {   // one more replay from a quiet GPU, then the per-launch timeline (ns, relative to the first CTA of the graph)
    CK(cudaDeviceSynchronize()); CK(cudaMemset(g_trace_buf, 0, trace_words * 8));
    CK(cudaGraphLaunch(ge, st[0])); CK(cudaGraphLaunch(ge, st[0])); CK(cudaStreamSynchronize(st[0]));
    std::vector<unsigned long long> h(trace_words);
    CK(cudaMemcpy(h.data(), g_trace_buf, trace_words * 8, cudaMemcpyDeviceToHost));
    unsigned long long t00 = ~0ull;
    for (size_t i = 0; i < trace_words; i += 6) if (h[i]) t00 = std::min(t00, h[i]);
    printf("launch: ctas | first/median/last CTA entry | median wait-exit(t1-t0) | first-tile done (t2-t1) med/max | exit (t3) first/median/last | tiles max\n");
    for (int l = 0; l < steps * smul; ++l) {
      std::vector<long> t0, w, c, t3; long tmax = 0;
      for (int k = 0; k < kTraceCtas; ++k) {
        const unsigned long long* r = &h[((size_t)l * kTraceCtas + k) * 6];
        if (!r[0]) continue;
        t0.push_back((long)(r[0] - t00)); w.push_back((long)(r[1] - r[0])); c.push_back((long)(r[2] - r[1])); t3.push_back((long)(r[3] - t00));
        tmax = std::max(tmax, (long)r[5]);
      }
      if (t0.empty()) continue;
      for (auto* v : {&t0, &w, &c, &t3}) std::sort(v->begin(), v->end());
      auto med = [](std::vector<long>& v) { return v[v.size() / 2]; };
      printf("%3d: %4zu | %6ld %6ld %6ld | %5ld | %5ld %5ld | %6ld %6ld %6ld | %ld\n", l, t0.size(), t0.front(), med(t0), t0.back(), med(w),
             med(c), c.back(), t3.front(), med(t3), t3.back(), tmax);
    }
  }
#endif
  double hb[4]; CK(cudaMemcpy(hb, bits[0], sizeof(double) * (B < 4 ? B : 4), cudaMemcpyDeviceToHost));
  printf("B=%d n=%ld idx=%d noise=%d rate=%d smul=%d pf=%d chains=%d steps=%d : %.2f us/step  %.2f us/launch  %.1f GB/s  %.1f%% of 6537.6  (%.1f Gelem/s) bits0=%.3f\n",
         B, n, with_idx, noise, rate, smul, prefetch, chains, steps, us_step, us, gbs, 100 * gbs / 6537.6, N / us / 1e3, hb[0]);
  return 0;
}
