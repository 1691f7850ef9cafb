/* reslic_b200.h — C ABI of the B200-native entropy-model hot path for ResLIC_TCM.
 *
 * The reference (AlbertoPresta/ResLIC_TCM) has no FFI layer of its own: the operator
 * API for this path is the CompressAI `EntropyModel` nn.Module family that
 * `TCM.forward/compress/decompress` call per slice.  This header is the
 * C boundary a drop-in module binds instead (ctypes stub: reslic_tcm_b200/_cabi.py;
 * reference-side binding shown in INTEGRATION.md).  Each entry cites the reference
 * interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *  - every function returns 0 on success, a negative RESLIC_ERR_* on a bad argument, or
 *    a positive cudaError_t; `reslic_last_error()` gives a thread-local message.
 *    No exception crosses the ABI.
 *  - all data pointers are DEVICE pointers owned by the caller; the library never
 *    allocates, frees or retains them.  Work is enqueued on the caller's stream
 *    (`stream` = a cudaStream_t cast to void*; NULL = legacy default stream) and is
 *    asynchronous; the caller synchronises.
 *  - tensors are fp32 (integer outputs int32), "image-major": image b of a tensor starts
 *    at base + b*batch_stride (in ELEMENTS) and is one contiguous run of `n` elements.
 *    This lets a channel slice of y[B,320,h,w] (tcm.py:438 `y.chunk`) be passed without a
 *    copy.  16-byte aligned bases and strides%4==0 and n%4==0 take the 128-bit path; any
 *    other layout takes a scalar CUDA path (never a CPU fallback).
 *  - every descriptor struct starts with `struct_size`, which the caller sets to sizeof(the struct) as ITS
 *    header declares it; a call whose struct_size differs from the library's sizeof is rejected with
 *    RESLIC_ERR_ARG before any field is read (a binding built against another ABI revision cannot make the
 *    library read past its struct).  reslic_sizeof_*() export the library's sizes for bindings that cannot
 *    include this header (ctypes, cgo, JNI).
 *  - `workspace`: device scratch of at least reslic_workspace_bytes() bytes, zero-filled
 *    ONCE by the caller before first use (kernels leave it zeroed); one workspace per
 *    stream in flight.
 */
#ifndef RESLIC_B200_H
#define RESLIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RESLIC_ABI_VERSION 17

enum {
  RESLIC_OK = 0,
  RESLIC_ERR_ARG = -1,       /* null/negative/inconsistent argument */
  RESLIC_ERR_WORKSPACE = -2, /* workspace missing or too small      */
  RESLIC_ERR_UNSUPPORTED = -3
};

/* CompressAI EntropyModel.quantize modes. */
enum {
  RESLIC_Q_DEQUANTIZE = 0, /* "dequantize": round(x - mu) + mu  (eval)            */
  RESLIC_Q_NOISE = 1,      /* "noise":      x + U(-1/2, 1/2)    (training)        */
  RESLIC_Q_IDENTITY = 2    /* reslic_eb_bwd_f32 only: z IS the quantizer output (STanH bottleneck) */
};

/* Likelihood arithmetic of the Gaussian-conditional kernel (process-wide switch).
 *  FAST   (default): custom erfc (1 RCP + 1 EX2 + 21 FP32 ops) and MUFU log2 for the rate;
 *                    same operand path (true fp32 divides, same constants) as the reference.
 *  MIRROR          : CUDA erfcf / log2f, i.e. the instruction-for-instruction arithmetic the
 *                    reference's torch CUDA kernels execute; slower, kept for cross-checks.
 * Integers (symbols, indexes) and y_hat are identical in both modes. */
enum { RESLIC_MATH_FAST = 0, RESLIC_MATH_MIRROR = 1 };
int reslic_set_math_mode(int mode);
int reslic_get_math_mode(void);

int reslic_abi_version(void);
const char* reslic_last_error(void);
/* Number of SMs of the current device (grid sizing is done inside the library). */
int reslic_device_sm_count(void);
/* Scratch bytes needed by any *_fwd call with a rate output for B images (4 words per image + 4). */
int64_t reslic_workspace_bytes(int64_t B);
/* sizeof of every struct of this header as the LIBRARY was compiled (see `struct_size` above). */
int64_t reslic_sizeof_gc_desc(void);
int64_t reslic_sizeof_gc_bwd_desc(void);
int64_t reslic_sizeof_eb_desc(void);
int64_t reslic_sizeof_eb_bwd_desc(void);
int64_t reslic_sizeof_stanh_tables(void);
int64_t reslic_sizeof_stanh_gc_desc(void);
int64_t reslic_sizeof_stanh_gc_bwd_desc(void);
int64_t reslic_sizeof_eb_stanh_desc(void);
int64_t reslic_sizeof_rate_exchange(void);

/* Rate output modes (the `bits_accumulate` field of every *_fwd descriptor):
 *   0  bits[b]  = -sum log2 L of this launch       (written by the launch's last-arriving warp)
 *   1  bits[b] += -sum log2 L of this launch
 *   RESLIC_RATE_DEFERRED  the sum is added to the workspace only (`bits` may be NULL and is not
 *      touched); any number of launches — the z launch and the five slice launches of one batch,
 *      training/loss.py:24-27 — accumulate there, and ONE reslic_rate_finalize_f64 call turns the
 *      workspace into bits and re-zeroes it.  The launches then end without the round trip to L2
 *      that detecting the last arriver costs (about 1.3 us of a 15 us slice launch on B200). */
#define RESLIC_RATE_DEFERRED 2
/*   RESLIC_RATE_COLLECT   bits[b] = this launch's sum + everything earlier RESLIC_RATE_DEFERRED launches
 *      left in the workspace (which is re-zeroed): the LAST launch of a batch collects, so the batch
 *      needs no finalize launch and pays the last-arriver round trip once instead of per launch. */
#define RESLIC_RATE_COLLECT 3
/* bits[b] (accumulate ? += : =) the deferred rate of image b, for b < B; same B and workspace as the
 * launches that accumulated.  Stream-ordered after them. */
int reslic_rate_finalize_f64(void* workspace, int64_t workspace_bytes, int64_t B, double* bits,
                             int32_t accumulate, void* stream);

/* Per-image rate of a likelihood tensor that already exists: bits[b] (mode as above) -sum log2 lik[b, 0..n),
 * image b at lik + b * lik_bs.  Replaces `torch.log(likelihoods).sum() / -math.log(2)` of the reference's
 * RateDistortionLoss (src/training/loss.py:22-25; src/eval.py:27-31) — one read pass instead of a log tensor
 * and a reduction over it; same deterministic commit and workspace as the *_fwd kernels. */
int reslic_rate_from_likelihood_f32(const float* lik, int64_t lik_bs, int64_t B, int64_t n, double* bits,
                                    int32_t bits_accumulate, void* workspace, int64_t workspace_bytes, void* stream);

/* ----------------------------------------------------------------------------------
 * Multi-GPU rate exchange without a collective kernel.
 * The path shards by images; the only coupling between ranks is the scalar rate (and distortion) sum that
 * the reference gets by gathering whole likelihood tensors to GPU 0 (nn.DataParallel: src/utils/helper.py:106-113,
 * src/train.py:168-169) before src/training/loss.py:24-27 reduces them.  Here the launch that COLLECTS a batch's
 * rate also publishes it: the warp that completes the batch stores one packed row
 *     { sum_b bits[b], extra[0] (e.g. the squared-error sum), pixels, images }      (4 doubles)
 * straight into every rank's exchange buffer over NVLink (plain peer stores, each cell tagged with its step) — no
 * separate kernel, no NCCL call on the step.  reslic_rate_exchange_read_f64 later adds the `world` rows of a
 * step in rank order (bit-identical on every rank).
 *
 * Exchange buffer (one per rank, zero-filled once, 16-byte aligned; `peer_base[r]` is rank r's buffer as mapped in THIS
 * process — reslic_peer_buffer_* below, or any other peer-accessible allocation):
 *     cells  {double value; uint64 tag}  [ring][world][4]       row (slot, r) is written by rank r only
 * tag = step + 1 once the value of row (step % ring, r) is there: every cell is ONE 16-byte store, so value and tag
 * arrive together and the writer needs no fence.
 * A launch publishes step number `*cursor + step`: `step` is the caller's number of the batch relative to `cursor`, a
 * zero-initialised DEVICE word local to this rank that the CALLER advances (stream-ordered) — after every batch in
 * eager use, or once per CUDA-graph replay by the number of batches the graph holds, whose launches then carry
 * step = 0, 1, 2, ... as constants.  The slot of a batch is therefore fixed by WHICH batch it is, never by the order
 * in which concurrently running batches (several streams, graph branches) happen to finish, and every rank files the
 * same batch under the same step.  A rank may run at most `ring` steps ahead of the slowest reader.  */
typedef struct reslic_rate_exchange {
  uint64_t struct_size;                    /* = sizeof(reslic_rate_exchange)                                  */
  int32_t world, rank;                     /* 1 <= world <= 64                                                */
  int32_t ring;                            /* slots per buffer                                                */
  int32_t reserved;
  int64_t step;                            /* this batch's number relative to *cursor (>= 0)                  */
  void* const* peer_base;                  /* DEVICE array [world] of buffer bases (entry `rank` = own buffer) */
  const unsigned long long* cursor;        /* DEVICE word, see above; NULL = 0                                */
  const double* extra;                     /* DEVICE, nullable: value published as the row's second field,    */
                                           /* read when the batch completes                                   */
  double pixels, images;                   /* the row's third and fourth field                                */
} reslic_rate_exchange;
int64_t reslic_rate_exchange_bytes(int32_t world, int32_t ring);
/* out[s][0..3] = sum over ranks (in rank order) of the rows of steps base + first_step + s, s < n_steps, read from THIS
 * rank's buffer `own_base`; base = *cursor (the DEVICE word of reslic_rate_exchange) or 0 when cursor is NULL.  With
 * the cursor a read is relative to "now" — first_step = -n reads the n steps published last — so it can be captured
 * in the same CUDA graph as the launches that publish; steps before 0 give a zero row.  Waits (bounded: about 2 s, then
 * the step's row is NaN and status[0] |= 1) until every rank's row for the step has arrived; status[0] |= 2 if a row
 * was overwritten (a rank ran `ring` steps ahead).  status: DEVICE int32[1], zeroed by the caller. */
int reslic_rate_exchange_read_f64(const void* own_base, int32_t world, int32_t ring, const unsigned long long* cursor,
                                  int64_t first_step, int32_t n_steps, double* out, int32_t* status, void* stream);

/* Stand-alone publisher: the same row a collecting launch publishes — {sum_b bits[b], extra, pixels, images} as step
 * *cursor + ex->step — from a bits[B] vector that already exists (stream-ordered behind the launch that wrote it).  For
 * callers that want the collecting launch itself untouched: in a CUDA graph this one-CTA kernel sits on a branch of its
 * own behind the batch's last launch, off the next batch's critical path.  bits[b] are multiples of 2^-16 (the kernels'
 * fixed-point commit), so the sum is exact and equals the fused form's bit for bit. */
int reslic_rate_exchange_publish_f64(const reslic_rate_exchange* ex, const double* bits, int64_t B, void* stream);

/* Peer-accessible device memory for the exchange buffers (CUDA IPC): `create` allocates `bytes` zero-filled bytes on
 * the current device and fills the 64-byte handle that another PROCESS on the same node passes to `open` (which
 * maps the buffer, enabling peer access) — torch.distributed moves the handles.  `close` unmaps, `destroy` frees. */
#define RESLIC_PEER_HANDLE_BYTES 64
int reslic_peer_buffer_create(int64_t bytes, void** dptr, uint8_t* handle);
int reslic_peer_buffer_open(const uint8_t* handle, void** dptr);
int reslic_peer_buffer_close(void* dptr);
int reslic_peer_buffer_destroy(void* dptr);

/* ----------------------------------------------------------------------------------
 * Gaussian conditional, fused forward.
 * Replaces, in ONE pass over a slice (any subset of outputs):
 *   compressai GaussianConditional.forward            (call site tcm.py:455)
 *     = EntropyModel.quantize "noise"/"dequantize"     (copy: adaptive_gaussian_conditional.py:95-141)
 *     + _likelihood + both LowerBounds                 (twin: tcm.py:570-588)
 *   ste_round(y - mu) + mu                             (tcm.py:36-37, 457)
 *   quantize(y, "symbols", mu)                         (tcm.py:548)
 *   build_indexes(sigma)                               (adaptive_gaussian_conditional.py:606-617; tcm.py:544,619)
 *   sum log(L) / -ln2  per image                       (training/loss.py:24-27; eval.py:27-31)
 * -------------------------------------------------------------------------------- */
typedef struct reslic_gc_desc {
  uint64_t struct_size;                    /* = sizeof(reslic_gc_desc); checked by the library           */
  /* inputs */
  const float* y;       int64_t y_bs;      /* latent slice                                */
  const float* mu;      int64_t mu_bs;     /* means; NULL = no means (values = inputs)    */
  const float* sigma;   int64_t sigma_bs;  /* scales (pre-LowerBound)                     */
  const float* noise;   int64_t noise_bs;  /* NOISE mode: explicit U(-.5,.5) draw; NULL = */
                                           /* in-kernel Philox4x32-10 (seed, offset)      */
  int64_t B;                               /* images                                      */
  int64_t n;                               /* elements per image (C*h*w of the slice)     */
  int32_t mode;                            /* RESLIC_Q_*                                  */
  float scale_bound;                       /* LowerBound on sigma (0.11)                  */
  float likelihood_bound;                  /* LowerBound on L (1e-9); <= 0 disables       */
  const float* scale_table;                /* [table_len] ascending; needed iff idx != 0  */
  int32_t table_len;                       /* <= 256                                      */
  /* outputs, each nullable */
  float* yhat;   int64_t yhat_bs;          /* quantize() output: y+u or round(y-mu)+mu    */
  float* ste;    int64_t ste_bs;           /* round(y-mu)+mu regardless of mode           */
  float* lik;    int64_t lik_bs;           /* bounded likelihood                          */
  int32_t* sym;  int64_t sym_bs;           /* int32(round(y-mu))                          */
  int32_t* idx;  int64_t idx_bs;           /* scale-table / CDF index, 0..table_len-1     */
  double* bits;                            /* [B] -sum_i log2 L per image                 */
  int32_t bits_accumulate;                 /* 0: bits[b] = sum;  1: bits[b] += sum (lets  */
                                           /* the z and the 5 slice launches of one batch */
                                           /* share one rate vector, loss.py:24-27);      */
                                           /* RESLIC_RATE_DEFERRED: see above             */
  void* workspace; int64_t workspace_bytes;/* required iff a rate output is requested     */
  uint64_t philox_seed, philox_offset;
  /* Optional hint: the y (same B, n) that the NEXT launch on this stream will read — TCM's next channel slice of
   * the same latent (tcm.py:438-457).  CTAs that finish early prefetch it into L2 (cp.async.bulk.prefetch.L2), which
   * moves a third of the next launch's reads into this launch's under-used tail.  Never dereferenced by the math;
   * NULL or a misaligned pointer = no prefetch. */
  const float* next_y;  int64_t next_y_bs;
  /* Optional (HOST pointer, read during the call): publish the batch's rate to every rank, see reslic_rate_exchange.
   * Only with bits_accumulate 0 or RESLIC_RATE_COLLECT — the modes in which this launch completes bits[]. */
  const reslic_rate_exchange* exchange;
} reslic_gc_desc;

int reslic_gc_fwd_f32(const reslic_gc_desc* d, void* stream);

/* ----------------------------------------------------------------------------------
 * Gaussian conditional, backward of reslic_gc_fwd_f32 — what autograd does
 * through compressai GaussianConditional.forward + ste_round in the reference's training step
 * (src/training/step.py:38-43): erfc/abs/divide chain, both LowerBound gradient rules
 * (pass where x >= bound or grad < 0, App. A.4), identity through the additive noise,
 * zero through round(), identity through ste_round.
 * Inputs are the forward's inputs (the noise is regenerated from the same seed/offset or re-read)
 * plus the upstream gradients (each nullable = zero).  Outputs nullable.
 *   mode NOISE      : g_y = g_yhat + g_ste + gv,  g_mu = -gv,          g_sigma = gs*pass
 *   mode DEQUANTIZE : g_y = g_ste,                g_mu = g_yhat,       g_sigma = gs*pass
 *   with gv = gL*pass_L * dL/dv * sign(yhat - mu), gs = gL*pass_L * dL/ds. */
typedef struct reslic_gc_bwd_desc {
  uint64_t struct_size;                    /* = sizeof(reslic_gc_bwd_desc); checked by the library           */
  const float* y;       int64_t y_bs;
  const float* mu;      int64_t mu_bs;     /* NULL = no means                              */
  const float* sigma;   int64_t sigma_bs;
  const float* noise;   int64_t noise_bs;  /* NULL = Philox(seed, offset) as in the forward */
  int64_t B, n;
  int32_t mode;
  float scale_bound, likelihood_bound;
  const float* g_yhat;  int64_t g_yhat_bs; /* d loss / d (quantize output)                  */
  const float* g_ste;   int64_t g_ste_bs;  /* d loss / d (ste_round output)                 */
  const float* g_lik;   int64_t g_lik_bs;  /* d loss / d (bounded likelihood)               */
  float* g_y;     int64_t g_y_bs;
  float* g_mu;    int64_t g_mu_bs;
  float* g_sigma; int64_t g_sigma_bs;
  uint64_t philox_seed, philox_offset;
} reslic_gc_bwd_desc;

int reslic_gc_bwd_f32(const reslic_gc_bwd_desc* d, void* stream);

/* Latent-residual-prediction tail of the slice loop (tcm.py:461-464):
 *   y_hat += 0.5 * tanh(lrp)   in place, image-major like reslic_gc_desc (3 launches -> 1). */
int reslic_lrp_tail_f32(float* y_hat, int64_t y_hat_bs, const float* lrp, int64_t lrp_bs, int64_t B, int64_t n,
                        void* stream);

/* build_indexes alone (adaptive_gaussian_conditional.py:606-617): flat, n elements. */
int reslic_build_indexes_f32(const float* sigma, int64_t n, float scale_bound,
                             const float* scale_table, int32_t table_len,
                             int32_t* idx, void* stream);

/* EntropyModel.dequantize (App. A.1; tcm.py:623): out = float(sym) + mu (mu nullable). */
int reslic_dequantize_f32(const int32_t* sym, const float* mu, int64_t n, float* out,
                          void* stream);

/* ----------------------------------------------------------------------------------
 * Factorized entropy bottleneck on z, fused forward.
 * Replaces compressai EntropyBottleneck.forward (call site tcm.py:429) incl. the
 * permute/reshape wrapper (adaptive_entropy_bottleneck.py:679-708), quantize about the
 * medians, 2x _logits_cumulative (:525-543), sign-trick likelihood (:658-666), the
 * LowerBound, ste_round(z - med) + med (tcm.py:431-433), the compress-path symbols
 * (tcm.py:507) and the per-image rate sum.  filters must be (3,3,3,3).
 * z is [B, C, hw] with batch stride z_bs; parameters are the module's raw tensors.
 * -------------------------------------------------------------------------------- */
typedef struct reslic_eb_desc {
  uint64_t struct_size;                    /* = sizeof(reslic_eb_desc); checked by the library           */
  const float* z;      int64_t z_bs;
  const float* noise;  int64_t noise_bs;   /* as in reslic_gc_desc                        */
  int64_t B, C, hw;
  int32_t mode;                            /* RESLIC_Q_*                                  */
  float likelihood_bound;
  const float* matrix[5];                  /* _matrix{i} [C,f_{i+1},f_i] raw (softplus in-kernel) */
  const float* bias[5];                    /* _bias{i}   [C,f_{i+1},1]                    */
  const float* factor[4];                  /* _factor{i} [C,f_{i+1},1] raw (tanh in-kernel) */
  const float* medians;                    /* [C] = quantiles[:,0,1]                      */
  float* zhat;  int64_t zhat_bs;           /* quantize() output                           */
  float* ste;   int64_t ste_bs;            /* round(z-med)+med regardless of mode         */
  float* lik;   int64_t lik_bs;
  int32_t* sym; int64_t sym_bs;            /* int32(round(z-med))                         */
  double* bits;                            /* [B]                                         */
  int32_t bits_accumulate;                 /* as in reslic_gc_desc                        */
  void* workspace; int64_t workspace_bytes;
  uint64_t philox_seed, philox_offset;
  const float* lut;                        /* optional, DEQUANTIZE mode: [C, RESLIC_EB_LUT_STRIDE] table  */
                                           /* from reslic_eb_build_lut_f32 for THESE parameters and bound */
  /* Optional hint as reslic_gc_desc.next_y: next_y_n floats per image (image b at next_y + b * next_y_bs) that the
   * NEXT launch will read — the first channel slice of y, which exists before z does (tcm.py:427-443).  Prefetched
   * into L2 by the table-mode launch, whose own traffic leaves HBM idle. */
  const float* next_y;  int64_t next_y_bs, next_y_n;
} reslic_eb_desc;

int reslic_eb_fwd_f32(const reslic_eb_desc* d, void* stream);

/* In DEQUANTIZE mode z_hat - median is an integer k, so the likelihood is a function of (channel, k).
 * reslic_eb_fwd_f32 builds that table (|k| <= 32) in every launch; a caller whose parameters do not
 * change between launches (evaluation, compress) builds it ONCE here and passes it as `lut`, which
 * takes the 5-layer cumulative-logit evaluation off the launch's latency chain.  The table holds, per
 * channel, 65 bounded likelihoods followed by their 65 log2 values — bit-identical to what the launch
 * computes itself.  Uses d->C, matrix/bias/factor/medians and likelihood_bound only. */
#define RESLIC_EB_LUT_STRIDE 130
int reslic_eb_build_lut_f32(const reslic_eb_desc* d, float* lut, void* stream);

/* Backward of reslic_eb_fwd_f32: gradients of the bounded likelihood (and of the
 * quantize output) w.r.t. z and w.r.t. every bottleneck parameter — what autograd does through
 * _logits_cumulative x2 + the sign trick + LowerBound (adaptive_entropy_bottleneck.py:525-543,
 * 658-666) in the reference's training step.  Parameter gradients are per-channel reductions over
 * all B*hw elements of the channel; they are OVERWRITTEN (raw-parameter space: softplus' / tanh'
 * already applied).  g_medians is non-zero only in DEQUANTIZE mode (z_hat = round(z-med)+med).
 * With the variable-bin fields at the end of the descriptor the same kernel is the backward of
 * reslic_eb_stanh_fwd_f32's likelihood (the STanH quantizer itself is differentiated by reslic_stanh_gc_bwd_f32). */
typedef struct reslic_eb_bwd_desc {
  uint64_t struct_size;                    /* = sizeof(reslic_eb_bwd_desc); checked by the library           */
  const float* z;      int64_t z_bs;
  const float* noise;  int64_t noise_bs;   /* NULL = Philox(seed, offset) as in the forward   */
  int64_t B, C, hw;
  int32_t mode;
  float likelihood_bound;
  const float* matrix[5]; const float* bias[5]; const float* factor[4]; const float* medians;
  const float* g_zhat; int64_t g_zhat_bs;  /* d loss / d (quantize output), nullable          */
  const float* g_lik;  int64_t g_lik_bs;   /* d loss / d (bounded likelihood), nullable       */
  float* g_z;  int64_t g_z_bs;             /* nullable                                        */
  float* g_matrix[5]; float* g_bias[5]; float* g_factor[4];   /* same shapes as the parameters, nullable as a set */
  float* g_medians;                        /* [C], nullable                                   */
  uint64_t philox_seed, philox_offset;
  /* Optional scratch (device, zero-initialised once, left zeroed by every launch; one per stream): with it
   * a channel's B*hw elements are cut over several CTAs, whose partial sums the channel's last CTA adds in
   * a fixed order (bit-reproducible); without it one CTA serves a whole channel.  Size:
   * reslic_eb_bwd_workspace_bytes(C). */
  void* workspace; int64_t workspace_bytes;
  /* Variable bins (EntropyBottleneckStanh, adaptive_entropy_bottleneck.py:551-603,643-666), all optional: with
   * half_lo / half_up the likelihood is |sigmoid(s f(x + up)) - sigmoid(s f(x - lo))| with per-element half-widths (as
   * reslic_eb_stanh_fwd_f32 wrote them) instead of 1/2; with `cell` and `g_dist` the gradients w.r.t. the half-widths
   * are summed per STanH level gap — g_dist[j-1] += dLoss/d lo, g_dist[j] += dLoss/d up for an element in cell j
   * (cell < 0: outside every cell, no half-width gradient) — into g_dist[n_dist] (DEVICE doubles, zeroed by the
   * caller; = dLoss / d distance_points).  Use with mode RESLIC_Q_IDENTITY and z = the quantizer's output. */
  const float* half_lo;  int64_t half_lo_bs;
  const float* half_up;  int64_t half_up_bs;
  const int32_t* cell;   int64_t cell_bs;
  double* g_dist;        int64_t n_dist;
} reslic_eb_bwd_desc;

int reslic_eb_bwd_f32(const reslic_eb_bwd_desc* d, void* stream);
int64_t reslic_eb_bwd_workspace_bytes(int64_t C);

/* ----------------------------------------------------------------------------------
 * STanH ("sum of tanh") quantizer family.
 * Tables are the state the reference's NonSymStanH / SymStanH modules keep
 * (src/quantization/activation.py:57-98, 214-234): K thresholds, K weights, K+1 levels,
 * K level mid-points and K half-gaps, all fp32 DEVICE arrays.
 * -------------------------------------------------------------------------------- */
typedef struct reslic_stanh_tables {
  const float* b;                /* [K] thresholds ASCENDING: sort(stanh.b) / sort(stanh.sym_b)        */
  const float* w;                /* [K] weights as the module pairs them with the sorted thresholds   */
                                 /*     (stanh.w, or stanh.sym_w for the symmetric form)              */
  const float* cum_w;            /* [K+1] levels (stanh.cum_w)                                        */
  const float* average_points;   /* [K] (stanh.average_points)                                        */
  const float* distance_points;  /* [K] (stanh.distance_points)                                       */
  int32_t K;                     /* 1..1024                                                           */
  int32_t symmetric;             /* 0: NonSymStanH (strict >), 1: SymStanH (sign(), half level at ties)*/
  float beta;                    /* soft-form temperature; -1 = hard form (activation.py:142-143)     */
} reslic_stanh_tables;

/* Fused GaussianConditionalStanh.forward / quantize / _likelihood
 * (src/entropy_models/adaptive_gaussian_conditional.py:588-603, 95-157, 541-580; call sites
 * src/models/stanh/tcm_stanh.py:432, wacnn_stanh.py:305, balle18_stanh.py:126):
 *   training != 0 : y_hat = stanh_beta(y - mu*[removing_mean]) + mu*[removing_mean]   (:108-117)
 *   training == 0 : y_hat = stanh_hard(y - mu) + mu                                   (:119-137)
 *   training == 2 : y_hat = y (no quantization: _likelihood of already-quantised values, :541)
 *   L = mass of the level cell of (y_hat - mu) under N(0, max(sigma, scale_bound)), bounded below
 *   sym = level index of stanh_hard(y - mu): 0..K (non-symmetric) or -K/2..K/2 (symmetric),
 *         i.e. stanh.map_sos_cdf without the per-element Python loop (:144-157)
 * Layout conventions as reslic_gc_desc. */
typedef struct reslic_stanh_gc_desc {
  uint64_t struct_size;                    /* = sizeof(reslic_stanh_gc_desc); checked by the library           */
  const float* y;      int64_t y_bs;
  const float* mu;     int64_t mu_bs;      /* NULL = no means                              */
  const float* sigma;  int64_t sigma_bs;
  int64_t B, n;
  int32_t training;                        /* forward(..., training=...)                   */
  int32_t removing_mean;                   /* gaussian_configuration["removing_mean"]      */
  float scale_bound, likelihood_bound;
  reslic_stanh_tables tables;
  float* yhat;   int64_t yhat_bs;
  float* lik;    int64_t lik_bs;
  int32_t* sym;  int64_t sym_bs;
  double* bits;  int32_t bits_accumulate;
  void* workspace; int64_t workspace_bytes;  /* as reslic_gc_desc                           */
  float* ste;    int64_t ste_bs;           /* nullable: ste_round(y - mu) + mu, what TCM's slice loop carries on when the
                                            * STanH is frozen (src/models/stanh/tcm_stanh.py:432-434) (ABI 17) */
} reslic_stanh_gc_desc;

int reslic_stanh_gc_fwd_f32(const reslic_stanh_gc_desc* d, void* stream);

/* Backward of reslic_stanh_gc_fwd_f32: gradients w.r.t. y, mu, sigma of the quantize output and of the
 * bounded likelihood, and — when `g_params` is given (gaussian_configuration["trainable"] = True) — the
 * raw material of the gradients w.r.t. the STanH parameters, which autograd gets in the reference through
 * the dense [1,K,N] soft quantizer (activation.py:146-149 / 301-304) and through distance_points inside
 * _likelihood (adaptive_gaussian_conditional.py:544-567; update_state runs under grad, tcm_stanh.py:399).
 * g_params (5K+2 doubles, overwritten) holds, over all elements with upstream gradient G on the quantizer
 * output and saturation window [lo, hi) of the soft form:
 *   A[K+1]  sum of G by lo          (every k <  lo has dq/dw_k = +1/2)
 *   Bq[K+1] sum of G by hi          (every k >= hi has dq/dw_k = -1/2)
 *   Ww[K]   sum of G * dq/dw_k  for lo <= k < hi
 *   Wb[K]   sum of G * dq/db_k  for lo <= k < hi          (b = the SORTED thresholds of `tables`)
 *   Hd[K]   sum of g_lik * dL/d distance_points[m]
 * so that  dLoss/dw_k = (sum_{m>k} A[m] - sum_{m<=k} Bq[m]) / 2 + Ww[k]  (w as paired with the sorted
 * thresholds), dLoss/db_k = Wb[k], dLoss/d distance_points[m] = Hd[m]; mapping those to the module's raw
 * w / b (mirroring of the symmetric form, distance_points = diff(cum_w)/2) is K-sized host work. */
typedef struct reslic_stanh_gc_bwd_desc {
  uint64_t struct_size;                    /* = sizeof(reslic_stanh_gc_bwd_desc); checked by the library           */
  const float* y;      int64_t y_bs;
  const float* mu;     int64_t mu_bs;
  const float* sigma;  int64_t sigma_bs;
  int64_t B, n;
  int32_t training, removing_mean;
  float scale_bound, likelihood_bound;
  reslic_stanh_tables tables;
  const float* g_yhat; int64_t g_yhat_bs;
  const float* g_lik;  int64_t g_lik_bs;
  float* g_y;     int64_t g_y_bs;
  float* g_mu;    int64_t g_mu_bs;
  float* g_sigma; int64_t g_sigma_bs;
  double* g_params; int64_t g_params_len;   /* optional, see above */
} reslic_stanh_gc_bwd_desc;

int reslic_stanh_gc_bwd_f32(const reslic_stanh_gc_bwd_desc* d, void* stream);

/* Fused EntropyBottleneckStanh.forward (src/entropy_models/adaptive_entropy_bottleneck.py:679-708;
 * call sites src/models/stanh/wacnn_stanh.py:160-161, balle18_stanh.py:26,124): STanH quantization of
 * z without medians (training != 0: soft, beta; == 0: hard), variable-bin sign-trick likelihood
 * (:551-603, :643-666), level index, per-image rate.  z is [B, C, hw]; filters (3,3,3,3). */
typedef struct reslic_eb_stanh_desc {
  uint64_t struct_size;                    /* = sizeof(reslic_eb_stanh_desc); checked by the library           */
  const float* z;  int64_t z_bs;
  int64_t B, C, hw;
  int32_t training;
  float likelihood_bound;
  const float* matrix[5]; const float* bias[5]; const float* factor[4];
  reslic_stanh_tables tables;
  float* zhat;  int64_t zhat_bs;
  float* lik;   int64_t lik_bs;
  int32_t* sym; int64_t sym_bs;
  double* bits; int32_t bits_accumulate;
  void* workspace; int64_t workspace_bytes;
  /* optional outputs for the backward (reslic_eb_bwd_f32 with variable bins): the half-widths of every element's
   * level cell and the cell index j (0..K; -1 = outside every cell, both half-widths 0) */
  float* half_lo;  int64_t half_lo_bs;
  float* half_up;  int64_t half_up_bs;
  int32_t* cell;   int64_t cell_bs;
} reslic_eb_stanh_desc;

int reslic_eb_stanh_fwd_f32(const reslic_eb_stanh_desc* d, void* stream);

/* The activation alone, flat over n elements (NonSymStanH/SymStanH.forward,
 * activation.py:135-150, 294-304) and the two sums compute_gap needs
 * (src/models/stanh/tcm_stanh.py:465-478):
 *   out_soft[i] = stanh_beta(x[i])   (nullable)      out_hard[i] = stanh_hard(x[i])  (nullable)
 *   gap2[0] = sum_i (x - stanh_beta(x))^2,  gap2[1] = sum_i (x - stanh_hard(x))^2   (nullable, fp64;
 *   gap = |gap2[0] - gap2[1]| / n).  `gap_workspace`: >= reslic_stanh_gap_workspace_bytes() bytes,
 *   zero-filled once, NOT shared with the rate workspace. */
int64_t reslic_stanh_gap_workspace_bytes(void);
int reslic_stanh_act_f32(const float* x, int64_t n, const reslic_stanh_tables* t, float* out_soft,
                         float* out_hard, double* gap2, void* gap_workspace, int64_t gap_workspace_bytes,
                         void* stream);

/* ----------------------------------------------------------------------------------
 * Host-side setup helper (no GPU work): pmf[n] -> cdf[n+1] with 2^precision total mass
 * and every symbol's frequency >= 1.  Replaces compressai._CXX.pmf_to_quantized_cdf
 * (src/entropy_models/coder.py:53-56; adaptive_gaussian_conditional.py:197-205).
 * HOST pointers. */
int reslic_pmf_to_quantized_cdf(const float* pmf, int32_t n, int32_t precision, uint32_t* cdf);

/* ----------------------------------------------------------------------------------
 * Host-side rANS coder with CDF indexes.  HOST pointers, no GPU work.
 * Replaces compressai.ans.{RansEncoder, BufferedRansEncoder, RansDecoder}
 * (src/models/reference/tcm.py:522,564-565,604-605,621;
 *  src/entropy_models/adaptive_gaussian_conditional.py:291-299,711-721).
 * cdfs: int32 [n_cdfs, cdf_stride] row-major (= module._quantized_cdf), cdf_sizes / offsets
 * [n_cdfs] (= _cdf_length / _offset).  value = symbol - offset[idx]; values outside
 * [0, cdf_size-2) are escaped through 4-bit bypass groups. */
void* reslic_rans_encoder_create(void);
/* The encoder divides the state by a symbol's frequency through a table of 64-bit reciprocals (one multiply instead of
 * a 64-bit divide per symbol).  This compares the table with the divide for every frequency 1..65535 at the edges of the
 * admissible state range and at `samples_per_freq` pseudo-random states each; returns the number of mismatches (0). */
int64_t reslic_rans_check_reciprocals(int64_t samples_per_freq, uint64_t seed);
void reslic_rans_encoder_destroy(void* enc);
int reslic_rans_encoder_push(void* enc, const int32_t* symbols, const int32_t* indexes, int64_t n,
                             const int32_t* cdfs, int32_t n_cdfs, int32_t cdf_stride,
                             const int32_t* cdf_sizes, const int32_t* offsets);
/* Device-side front end of reslic_rans_encoder_push: the per-symbol table lookup
 *   value = symbol - offsets[index], clamped to the escape slot cdf_sizes[index] - 2;
 *   slots[i] = cdf[value] << 16 | (cdf[value + 1] - cdf[value])
 * for n symbols, on the GPU behind the kernel that produced symbols and indexes, so that ONE packed 32-bit
 * slot per symbol crosses PCIe instead of two int32.  All pointers are DEVICE pointers; at most 1024 CDFs.
 * Symbols outside their table take the escape slot and are listed, in no particular order, as
 * (esc_pos[k], esc_raw[k]) = (position, bypass value) for k < min(status[0], esc_capacity).
 * status (int32[2], written by the call): [0] number of escaped symbols — if it exceeds esc_capacity the
 * list is incomplete and the caller must fall back to reslic_rans_encoder_push; [1] error bits
 * (1: an index outside [0, n_cdfs), 2: an invalid cdf entry — zero/negative frequency or >= 2^16). */
int reslic_rans_slots_u32(const int32_t* symbols, const int32_t* indexes, int64_t n, const int32_t* cdfs,
                          int32_t n_cdfs, int32_t cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets,
                          uint32_t* slots, int32_t* esc_pos, int64_t* esc_raw, int64_t esc_capacity,
                          int32_t* status, void* stream);
/* HOST: push n symbols whose lookups reslic_rans_slots_u32 has done; the escapes that belong to these n
 * symbols sorted by position (relative to slots[0]).  Produces the same stream as reslic_rans_encoder_push. */
int reslic_rans_encoder_push_slots(void* enc, const uint32_t* slots, int64_t n, const int32_t* esc_pos,
                                   const int64_t* esc_raw, int64_t n_esc);
/* returns the byte count (or -1); *data stays owned by the encoder until its next call */
int64_t reslic_rans_encoder_flush(void* enc, const uint8_t** data);
void* reslic_rans_decoder_create(const uint8_t* data, int64_t nbytes);
void reslic_rans_decoder_destroy(void* dec);
/* decodes the NEXT n symbols of the stream (decode_stream semantics) */
int reslic_rans_decoder_decode(void* dec, const int32_t* indexes, int64_t n, const int32_t* cdfs,
                               int32_t n_cdfs, int32_t cdf_stride, const int32_t* cdf_sizes,
                               const int32_t* offsets, int32_t* out_symbols);

#ifdef __cplusplus
}
#endif
#endif /* RESLIC_B200_H */
