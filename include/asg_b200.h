/*
 * asg_b200.h -- C ABI of libasg_b200.so: the ASG-loss hot path of zh217/torch-asg rebuilt as
 * hand-written sm_100a CUDA.  Plain pointers and sizes only; no torch types cross this boundary.
 *
 * What each entry point replaces in the reference (/root/reference/torch_asg/native):
 *
 *   asg_scan()   the four alpha/beta recursions and everything they drag along:
 *                  fully_connected_alpha_recursion   fully_connected_lattice.cpp:9-29
 *                  fully_connected_beta_recursion    fully_connected_lattice.cpp:32-47
 *                  make_aligned_inputs_{cpu,gpu}     force_aligned_lattice.cpp:15-46, _kernel.cu:9-105
 *                  make_aligned_transition_{cpu,gpu} force_aligned_lattice.cpp:48-82, _kernel.cu:136-226
 *                  force_aligned_alpha_recursion     force_aligned_lattice.cpp:84-111
 *                  force_aligned_beta_recursion      force_aligned_lattice.cpp:113-154
 *                  roll_to_end / should_roll_to_end  utils.cpp:22-66   (replaced by in-kernel length predicates)
 *                  the score read-outs               fully_connected_lattice.cpp:89, force_aligned_lattice.cpp:316
 *                and the stream choreography of fast_asg_gpu_forward{,_only}
 *                  (streamlined_fast_gpu.cpp:24-94, :104-230): the "roles" mask selects which of the four
 *                  scans run; each (utterance, scan) pair is its own CTA and the FCC and FAC launches overlap on two streams.
 *   asg_grad()   the no-recursion gradient assembly:
 *                  fully_connected_derivative/backward  fully_connected_lattice.cpp:50-63, :94-104
 *                  force_aligned_derivative/backward    force_aligned_lattice.cpp:156-183, :321-356
 *                  collect_input_grad_{cpu,gpu}         force_aligned_lattice.cpp:234-264, _kernel.cu:370-470
 *                  collect_transition_grad_{cpu,gpu}    force_aligned_lattice.cpp:204-231, _kernel.cu:255-344
 *                  fast_asg_gpu_backward                streamlined_fast_gpu.cpp:236-297
 *   asg_host_*() the whole ASGLoss forward+backward (asg.py:109-142) for callers that hold HOST buffers
 *                (the end-to-end path bench.py times; H2D/D2H copies are inside the call).
 *
 * Conventions (BASELINE.json letters): emissions x[T,N,L] (batch is dim 1), transition A[L,L] with
 * A[i,j] = score of going from j to i, targets y[N,S] int64, lengths il[N], tl[N] int64.
 * All strides are in ELEMENTS.  Every function returns 0 on success, a negative ASG_E_* code on a bad
 * argument, or a positive cudaError_t; asg_last_error() describes the last failure of the calling thread.
 * All work is enqueued on `stream`; nothing synchronises the host unless stated.
 *
 * Intermediate state is NOT the reference's log-domain alpha/beta: see DESIGN.md "Numerics".  The FCC
 * states are real-domain vectors rescaled by one power of two per frame; the FAC states are block
 * floating point (one power-of-two exponent per 8 target positions per frame).  All of it is kept in one
 * array of per-frame records (asg_layout_t).
 */
#ifndef ASG_B200_H
#define ASG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *asg_stream_t; /* a cudaStream_t */

enum { ASG_F32 = 0, ASG_F64 = 1 };

/* asg_scan(): which recursions to run */
enum {
    ASG_ROLE_FCC_ALPHA = 1,
    ASG_ROLE_FCC_BETA = 2,
    ASG_ROLE_FAC_ALPHA = 4,
    ASG_ROLE_FAC_BETA = 8,
    ASG_ROLE_ALL = 15,
    ASG_ROLE_FORWARD_ONLY = 10 /* the two beta scans: fast_asg_gpu_forward_only */
};

/* asg_grad(): which lattice's gradients to assemble */
enum { ASG_PART_FCC = 1, ASG_PART_FAC = 2, ASG_PART_BOTH = 3 };

enum {
    ASG_OK = 0,
    ASG_E_DTYPE = -1,   /* dtype is not ASG_F32 / ASG_F64 */
    ASG_E_SHAPE = -2,   /* non-positive T/N/L/S, or S > T (asg.py:119-122 truncates before the boundary) */
    ASG_E_NULL = -3,    /* a required pointer is NULL */
    ASG_E_UNSUPPORTED = -4, /* shape outside what this build covers: L > 2048, S > 4096 (when L > 64: S > 768 fp32 / S > 384 fp64), fp64 S > 2048 */
    ASG_E_NOGPU = -5    /* no CUDA device / driver: there is no CPU fallback */
};

/*
 * Sizes and offsets of the opaque buffers a caller allocates for one (dtype,T,N,L,S) problem.
 *
 * All per-frame state lives in ONE array of FRAME RECORDS, utterance-major: state[N][tp] records of rec_bytes each,
 *   record(n,t) = | FAC alpha[sp] | FAC beta[sp] | FAC alpha exponents[nbp] i32 | FAC beta exponents[nbp] i32 |
 *                 | FCC alpha[lpi] | FCC beta[lpi] | FCC u*beta[lpi] | FCC scale |            (every part 32-byte aligned)
 * so that (a) each of the four serial scans writes one sequential, sector-aligned stream, and (b) the gradient assembly
 * fetches a (utterance, frame-range) tile with ONE contiguous bulk (TMA) copy of >= 16 KB -- measured on B200, 1-D bulk
 * copies cost ~0.37 us apiece and are processed serially per SM, so eight small per-array copies per tile cap a kernel
 * at ~2.7 TB/s while one 32 KB copy per tile streams at the HBM peak (tools/tma_probe.cu).
 */
typedef struct asg_layout {
    int64_t sp;               /* FAC row length (elements): S rounded up to 8 */
    int64_t nblk;             /* block exponents per (frame, utterance): sp / 8 */
    int64_t nbp;              /* exponent row length: nblk rounded up to 8 */
    int64_t lpi;              /* FCC row length: L rounded up to 8 */
    int64_t tp;               /* frame pitch: T rounded up to 4 */
    size_t elem_bytes;        /* 4 or 8 */
    size_t rec_bytes;         /* one frame record (multiple of 32) */
    size_t state_bytes;       /* the state array: N * tp * rec_bytes */
    size_t ctx_bytes;         /* the context blob handed from asg_scan() to asg_grad() */
    size_t grad_scratch_bytes;/* scratch for asg_grad() (per-CTA partial sums) */
    /* byte offsets inside a frame record */
    size_t rec_fac_alpha, rec_fac_beta, rec_fac_ea, rec_fac_eb, rec_fcc_alpha, rec_fcc_beta, rec_fcc_w, rec_fcc_scale;
    /* byte offsets inside the context blob (all 256-byte aligned) */
    size_t off_ehat;          /* [L,L]       exp(A - R)                              */
    size_t off_fac_hd;        /* [2,N,sp]    exp(H - Rb), exp(D - Rb)                */
    size_t off_lengths;       /* [2,N] int64 copies of il, tl                        */
    size_t off_scalars;       /* [N,8] float64: per-utterance log-scale bookkeeping (scores in f64 etc.) */
    size_t off_ehat_t;        /* [L,L]       transpose of ehat           (L > 64 only, else = off_ehat)   */
    size_t off_fmax;          /* [N,tp]      per-frame emission maxima   (L > 64 only)                   */
    size_t off_rmax;          /* [1]         max A                       (L > 64 only)                   */
    size_t off_large_ws;      /* state exchange, maxima, barrier counters of the large-alphabet FCC scan (L > 64 only) */
} asg_layout_t;

/* Fills *out.  Returns ASG_E_UNSUPPORTED for shapes this build does not cover. */
int asg_layout(int dtype, int64_t T, int64_t N, int64_t L, int64_t S, asg_layout_t *out);

/*
 * Forward recursions.  `state` (state_bytes, device, 256-byte aligned) receives the frame records of the selected
 * roles; it may be NULL when only beta roles are selected: the beta scans then only produce the scores
 * (forward-only mode, nothing is stored -- fast_asg_gpu_forward_only, streamlined_fast_gpu.cpp:24-94).
 *   ctx                 : ctx_bytes device (always required)
 *   full_scores         : [N] dtype, written when ASG_ROLE_FCC_BETA is selected
 *   aligned_scores      : [N] dtype, written when ASG_ROLE_FAC_BETA is selected
 * il/tl are DEVICE pointers (NULL = full lengths T / S).
 */
int asg_scan(int dtype, int roles,
             const void *x, int64_t x_st, int64_t x_sn, int64_t x_sl,
             const void *A, int64_t A_s0, int64_t A_s1,
             const int64_t *y, int64_t y_s0, int64_t y_s1,
             const int64_t *il, const int64_t *tl,
             int64_t T, int64_t N, int64_t L, int64_t S,
             void *state, void *ctx,
             void *full_scores, void *aligned_scores,
             asg_stream_t stream);

/*
 * Gradient assembly.  grad_x [T,N,L] (contiguous) and grad_A [L,L] are fully overwritten:
 *   grad_x = softmax-posterior(FCC) * g_full  +  scatter(softmax-posterior(FAC) * g_aligned)
 *   grad_A = likewise for the transition matrix
 * i.e. the two upstream vectors are applied independently and the FAC part is ADDED, exactly like
 * fast_asg_gpu_backward (streamlined_fast_gpu.cpp:261-287); autograd supplies g_aligned = -g_full.
 * g_full/g_aligned: [N] dtype device, element strides g*_stride (0 = broadcast scalar).
 * `parts` selects FCC only / FAC only / both (the records must hold the matching roles); unused inputs may be NULL.
 */
int asg_grad(int dtype, int parts,
             const void *g_full, int64_t g_full_stride,
             const void *g_aligned, int64_t g_aligned_stride,
             const void *state, const void *ctx,
             const int64_t *y, int64_t y_s0, int64_t y_s1,
             int64_t T, int64_t N, int64_t L, int64_t S,
             void *grad_x, void *grad_A, void *scratch,
             asg_stream_t stream);

/*
 * One whole training step of the loss on DEVICE buffers, in one call: the four scans, the gradient assembly and the
 * reductions that the reference spreads over ASGGPUFast.forward, `full - aligned`, `.mean()` / `.sum()` and
 * ASGGPUFast.backward (asg.py:71-97, :128-142; streamlined_fast_gpu.cpp:104-297).  It allocates nothing, creates
 * nothing and never synchronises the host -- 4 kernel launches and one fork/join of the side stream -- so the call can
 * be stream-captured into a CUDA graph (together with the all-reduce that follows it on `fused`) and costs a few
 * microseconds of host time.  All buffers are the caller's, sized by asg_layout():
 *   state (state_bytes), ctx (ctx_bytes), scratch (grad_scratch_bytes)   opaque
 *   full_scores, aligned_scores, losses   [N] dtype      losses[n] = full[n] - aligned[n]
 *   grad_x                                [T,N,L] dtype  d(scale * sum_n losses[n]) / dx   (contiguous, fully overwritten)
 *   fused                                 [L*L + 2] dtype = grad_A (row-major) | sum_n losses[n] | N
 *                                         -- the buffer a batch-sharded job all-reduces (SUM) across ranks
 * `scale` is the upstream gradient of every utterance's loss: 1 for reduction 'sum', 1/global_batch for 'mean'.
 */
int asg_step(int dtype,
             const void *x, int64_t x_st, int64_t x_sn, int64_t x_sl,
             const void *A, int64_t A_s0, int64_t A_s1,
             const int64_t *y, int64_t y_s0, int64_t y_s1,
             const int64_t *il, const int64_t *tl,
             int64_t T, int64_t N, int64_t L, int64_t S,
             double scale,
             void *state, void *ctx, void *scratch,
             void *full_scores, void *aligned_scores, void *losses,
             void *grad_x, void *fused,
             asg_stream_t stream);

/*
 * The same step split where autograd splits it (ASGGPUFast.forward / .backward, asg.py:73-97), still with caller-owned
 * buffers and no allocation, so each half can sit in its own CUDA graph behind an autograd.Function:
 *   asg_forward():  the scans (forward_only != 0: the two beta scans alone, `state` unused -- ASGGPUFastForwardOnly,
 *                   asg.py:58-68), losses[n] = full[n] - aligned[n], loss_tail[0] = sum_n losses[n], loss_tail[1] = N,
 *                   loss_tail[2] = scale * sum_n losses[n] (the reduced loss: scale 1 = 'sum', 1/batch = 'mean', asg.py:137-142).
 *   asg_backward(): the gradient assembly for the upstream gradient  scale * g[n * g_stride]  of utterance n's loss
 *                   (g == NULL reads as 1; g_stride 0 = one device scalar, e.g. the grad_output of a reduced loss);
 *                   the FAC part is applied with the opposite sign, as autograd does for `full - aligned`.
 */
int asg_forward(int dtype, int forward_only,
                const void *x, int64_t x_st, int64_t x_sn, int64_t x_sl,
                const void *A, int64_t A_s0, int64_t A_s1,
                const int64_t *y, int64_t y_s0, int64_t y_s1,
                const int64_t *il, const int64_t *tl,
                int64_t T, int64_t N, int64_t L, int64_t S,
                double scale,
                void *state, void *ctx,
                void *full_scores, void *aligned_scores, void *losses, void *loss_tail /* [3] */,
                asg_stream_t stream);
int asg_backward(int dtype, const void *g, int64_t g_stride, double scale,
                 const void *state, const void *ctx,
                 const int64_t *y, int64_t y_s0, int64_t y_s1,
                 int64_t T, int64_t N, int64_t L, int64_t S,
                 void *grad_x, void *grad_A, void *scratch,
                 asg_stream_t stream);

/*
 * Fused log_softmax + ASG (the reference's users feed log_softmax'ed activations, test/test_asg.py:111-113).
 * The loss needs no separate normalisation pass at all: loss(log_softmax(z)) == loss(z), because the fully-connected score IS the
 * per-frame normaliser (both scores shift by the same sum of log-partition values), so asg_scan / asg_forward / asg_step are simply
 * run on the raw activations z.  The gradient is d loss / d z = gx - softmax(z) * sum_l gx, and sum_l gx[t,n,:] is exactly zero for
 * every utterance that has a valid alignment; only an infeasible utterance (aligned score -inf, loss +inf) needs the softmax term.
 * This call applies it in place to grad_x [T,N,L] (as written by asg_grad / asg_backward / asg_step for the same z): one tiny launch
 * whose CTAs leave at once for feasible utterances.  g / g_stride / scale: the upstream gradient of utterance n's loss,
 * scale * g[n * g_stride] (g == NULL reads as 1).  Saves, against log_softmax + ASG, one read and one write of [T,N,L] in the
 * forward pass and two reads and one write in the backward pass.
 */
int asg_logits_backward_fix(int dtype,
                            const void *z, int64_t z_st, int64_t z_sn, int64_t z_sl,
                            const void *aligned_scores,
                            const void *g, int64_t g_stride, double scale,
                            const int64_t *il,
                            int64_t T, int64_t N, int64_t L,
                            void *grad_x,
                            asg_stream_t stream);

/*
 * Range report.  The recursions run in the real domain (FCC) / in block floating point (FAC), not in the reference's log domain:
 * inside one frame (FCC: against the frame's maximum emission; FAC: inside one block of 8 target positions) entries more than
 * the format's range apart (fp32: 2^126 ~ 87 nats, fp64: 2^1022 ~ 708 nats) cannot coexist, and the smaller one is flushed to
 * ZERO (never clamped to the smallest normal number).  For normalised inputs (log-probabilities of a network) that never
 * matters; for extremely peaked inputs it can, and the kernels say so instead of returning a silently wrong gradient:
 * flags[n] (device int32 [N]) is the OR of
 *   1  a finite emission underflowed against its frame's maximum (exp(x - max) == 0),
 *   2  in some frame a block of 8 target positions could hold flushed entries whose alpha*beta product is not negligible (below
 *      2^-24) against the frame's largest product: posterior mass may have been lost,
 *   4  the aligned score of a feasible utterance came out as -inf,
 *   8  a target label of the utterance lies outside [0, L) (not a range matter, reported through the same word): the reference
 *      gathers out of bounds there (force_aligned_lattice.cpp:36-60); these kernels clamp the label to [0, L-1] and say so here.
 * Bits 1 and 4 are final after asg_scan / asg_forward, bit 2 after asg_grad / asg_backward / asg_step on the same ctx.  A caller
 * that sees a flag should rerun that batch in ASG_F64 (same API, 8x the range).
 */
int asg_range_flags(int dtype, const void *state, const void *ctx, int64_t T, int64_t N, int64_t L, int64_t S, int32_t *flags,
                    asg_stream_t stream);

/*
 * Viterbi forced alignment ("Viterbi decoders" is an open TODO of the reference, README.md:33): the
 * force-aligned lattice (force_aligned_lattice.cpp:84-111, doc/tech_report.tex:221-229) in the tropical semiring
 * (tech_report.tex:84-87, max instead of log-sum-exp).  scores[n] = the best path's score (dtype; -inf when the target does
 * not fit), path[n,t] = the target position aligned to frame t (int64, -1 on padding frames and infeasible utterances);
 * ties prefer staying on the current position.  scratch: asg_viterbi_scratch_bytes(T, N, S) bytes (back-pointer bits).
 */
size_t asg_viterbi_scratch_bytes(int64_t T, int64_t N, int64_t S);
int asg_viterbi(int dtype,
                const void *x, int64_t x_st, int64_t x_sn, int64_t x_sl,
                const void *A, int64_t A_s0, int64_t A_s1,
                const int64_t *y, int64_t y_s0, int64_t y_s1,
                const int64_t *il, const int64_t *tl,
                int64_t T, int64_t N, int64_t L, int64_t S,
                void *scores, int64_t *path, void *scratch,
                asg_stream_t stream);

/* ---- host-buffer convenience API (end-to-end path) ------------------------------------------- */
typedef struct asg_host_ctx asg_host_ctx_t;

/* Allocates device + pinned staging buffers for problems up to (T,N,L,S) on `device`. */
int asg_host_create(int dtype, int device, int64_t T, int64_t N, int64_t L, int64_t S, asg_host_ctx_t **out);
void asg_host_destroy(asg_host_ctx_t *h);

/*
 * One ASGLoss forward+backward from HOST memory: copies x, A, y, il, tl to the device, runs the four
 * scans and the gradient assembly, copies loss[N] (= full - aligned), grad_x[T,N,L] and grad_A[L,L]
 * back, and waits for completion.  g[N] is the upstream gradient on the per-utterance loss
 * (NULL = all ones, i.e. reduction 'sum').  grad_x / grad_A may be NULL (forward only: the two beta
 * scans alone run, like ASGLoss(forward_only=True)).  x, y contiguous.
 */
int asg_host_loss(asg_host_ctx_t *h,
                  const void *x, const void *A, const int64_t *y, const int64_t *il, const int64_t *tl,
                  const void *g, int64_t T, int64_t N, int64_t L, int64_t S,
                  void *loss, void *grad_x, void *grad_A);

/*
 * The same, split for pipelining: asg_host_submit() enqueues the copies on the context's private stream and the kernels
 * on the device's one compute stream (events hand over in both directions) and returns; asg_host_wait() blocks until that
 * context's results are in host memory.  Several contexts may be in flight at once (one submit per context between waits):
 * the kernels of all contexts then run back to back in submission order while the H2D of later steps and the D2H of earlier
 * ones overlap them.  Host buffers must be pinned for the overlap.  Thread-safe across contexts.
 */
int asg_host_submit(asg_host_ctx_t *h,
                    const void *x, const void *A, const int64_t *y, const int64_t *il, const int64_t *tl,
                    const void *g, int64_t T, int64_t N, int64_t L, int64_t S,
                    void *loss, void *grad_x, void *grad_A);
int asg_host_wait(asg_host_ctx_t *h);

/*
 * Target preprocessing on the device (the reference leaves it to the user, README.md:93-94): replabels.
 * ASG transcripts cannot repeat a label (doc/tech_report.tex:126), so runs are written with "repeat" tokens, wav2letter++
 * convention (flashlight lib/text packReplabels / unpackReplabels): token first_rep_label + k - 1 (1 <= k <= max_reps) means "the
 * previous label k MORE times"; a run longer than max_reps + 1 starts over (a a a a, max_reps 2 -> a r2 a).  max_reps = 0 copies.
 *   pack:   y[N,S] (strides), tl[N] or NULL -> out[N,S] contiguous (never longer than the input), out_len[N]; tail = fill
 *   unpack: the inverse, into out[N,S_out]; out_len[n] is the FULL unpacked length even when it exceeds S_out (then truncated)
 */
int asg_pack_replabels(const int64_t *y, int64_t y_s0, int64_t y_s1, const int64_t *tl, int64_t N, int64_t S, int64_t first_rep_label,
                       int64_t max_reps, int64_t fill, int64_t *out, int64_t *out_len, asg_stream_t stream);
int asg_unpack_replabels(const int64_t *y, int64_t y_s0, int64_t y_s1, const int64_t *tl, int64_t N, int64_t S, int64_t first_rep_label,
                         int64_t max_reps, int64_t fill, int64_t *out, int64_t S_out, int64_t *out_len, asg_stream_t stream);

/*
 * The exchange step of a batch-sharded job (one process per GPU, utterances sharded over ranks): the sum over ranks
 * of the fused vector [L*L transition gradient | loss sum | utterance count] that asg_backward / asg_step leave in `fused`.
 * The reference has no multi-GPU path of its own (its users wrap the criterion in DistributedDataParallel, which all-reduces
 * transition.grad through NCCL); this is that exchange written over NVLink peer memory, as ONE kernel per rank per step:
 * every rank stores its vector into every rank's exchange buffer, raises a flag there, waits for all flags to arrive locally
 * and adds the slots in rank order (bit-identical result on every rank).  No library collective, no host involvement, no
 * per-call argument: it replays inside a CUDA graph.
 *
 *   asg_peer_create   allocates this rank's exchange buffer (cudaMalloc on the current device) and returns its CUDA IPC handle
 *                     (ASG_PEER_HANDLE_BYTES bytes) for the caller to all-gather over whatever channel it has;
 *   asg_peer_connect  opens the other ranks' buffers (handles: world * ASG_PEER_HANDLE_BYTES bytes in rank order); fails with
 *                     a message when a peer is not reachable by peer access -- the caller then keeps its library collective;
 *   asg_peer_allreduce  in-place sum of buf[count] (dtype) over all ranks, stream-ordered; count * sizeof <= max_bytes;
 *                     every rank must issue the same sequence of calls;
 *   asg_peer_status   (synchronising) the number of exchanges done and the epoch of the first one that timed out waiting
 *                     for a peer (0 = none; a rank gives up after 4 s instead of hanging the GPU and leaves buf untouched).
 * world <= 8 (one NVSwitch domain).
 */
#define ASG_PEER_HANDLE_BYTES 64
typedef struct asg_peer asg_peer_t;
int asg_peer_create(int rank, int world, size_t max_bytes, asg_peer_t **out, void *handle);
int asg_peer_connect(asg_peer_t *p, const void *handles);
int asg_peer_allreduce(asg_peer_t *p, int dtype, void *buf, int64_t count, asg_stream_t stream);
int asg_peer_status(asg_peer_t *p, uint64_t *epoch, uint64_t *timed_out_epoch);
int asg_peer_destroy(asg_peer_t *p);

/* Destroys the side streams and events the library created for asg_scan()'s FCC/FAC overlap (one set per
 * (device, caller stream) pair) and the per-device compute stream of the asg_host_* path.  Optional; call when no work of
 * this library is in flight. */
void asg_release(void);

const char *asg_last_error(void);
const char *asg_version(void);
/* kernels launched by this library since load (bench.py's gpu_launches) */
uint64_t asg_launch_count(void);

/*
 * Live per-kernel timing (bench.py's roofline leg).  While enabled, every kernel this library launches is
 * bracketed by a pair of CUDA events recorded on the stream it is launched on; asg_profile_read() waits for
 * the recorded events and returns the accumulated device time and launch count of one kernel id
 * (0 <= id < asg_kernel_count(); asg_kernel_name(id) names it).  No reference counterpart: the reference
 * has no timers.
 */
int asg_profile_enable(int on);
/* restrict the timing to the kernel ids whose bit is set (default: all).  Event pairs on the two scan streams cost
 * far more than the kernels' launch gaps, so a headline run times only the kernel whose roofline it reports. */
int asg_profile_select(unsigned mask);
/* time only every stride-th launch of each selected kernel (default 1 = every launch): an event pair around a kernel
 * still costs the stream ~20 us of launch overlap, so a headline run samples the roofline kernel instead of timing
 * every launch. */
int asg_profile_stride(unsigned stride);
int asg_profile_reset(void);
int asg_profile_read(int kernel_id, double *total_ms, uint64_t *launches);
const char *asg_kernel_name(int kernel_id);
int asg_kernel_count(void);

#ifdef __cplusplus
}
#endif
#endif /* ASG_B200_H */
