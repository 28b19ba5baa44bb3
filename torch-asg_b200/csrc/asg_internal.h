// asg_internal.h -- host/device structs shared between the translation units of libasg_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>
#include <stdint.h>
#include "../../include/asg_b200.h"

namespace asg {

// per-utterance float64 bookkeeping slots in Ctx::scalars
enum { SC_R = 0, SC_FULL_BETA = 1, SC_FULL_ALPHA = 2, SC_RB = 3, SC_AL_BETA = 4, SC_AL_ALPHA = 5,
       SC_FCC_LOGSCALE_END = 6, SC_FLAGS = 7 };

// Range flags of an utterance (an int32 in the SC_FLAGS slot; cleared by asg_scan, OR-ed by the kernels, read by asg_range_flags):
// the real-domain / block-floating-point arithmetic covers ~87 nats (fp32; 708 in fp64) of dynamic range between the entries
// that share a frame maximum or a block exponent; beyond that mass is flushed to zero (never clamped) and these bits say so.
enum { RF_EMISSION_UNDERFLOW = 1,   // a finite emission lies more than the format's range below its frame's maximum: exp(x - max) = 0
       RF_POSTERIOR_THIN = 2,       // in some frame a block of target positions could hold FLUSHED entries (more than 2^126 below the
                                    // block's maxima) whose alpha*beta product is not negligible against the frame's largest one
       RF_SCORE_LOST = 4,           // the aligned score of a feasible utterance came out as -inf
       RF_LABEL_RANGE = 8 };        // a target label outside [0, L): the reference indexes out of bounds there; here the label was clamped

// typed view of the frame-record array + the context blob (see asg_layout_t)
template <typename T> struct Ctx {
    char *rec;        // [N,tp] frame records
    int64_t R;        // record size in bytes
    int o_fa, o_fb, o_ea, o_eb, o_ca, o_cb, o_cw, o_sc;   // byte offsets of the parts inside a record
    T *ehat;          // [L,L]
    T *ehat_t;        // [L,L]  transpose (large alphabets only)
    T *fmax;          // [N,tp] per-frame emission maxima (large alphabets only)
    T *rmax;          // [1]    max A (large alphabets only)
    char *large_ws;   // workspace of the resident-tile FCC scan (large alphabets only, asg_large.cu)
    T *fac_hd;        // [2,N,sp]   h_s, d_s (edge s -> s+1)
    int64_t *il;      // [N]
    int64_t *tl;      // [N]
    double *scalars;  // [N,8]
    __host__ __device__ __forceinline__ int *flags(int64_t n) const { return reinterpret_cast<int *>(scalars + n * 8 + SC_FLAGS); }
    int64_t sp, nblk, nbp, lpi, tp;
    __host__ __device__ __forceinline__ char *record(int64_t row) const { return rec + row * R; }
    __host__ __device__ __forceinline__ T *fac_alpha(int64_t row) const { return (T *)(rec + row * R + o_fa); }
    __host__ __device__ __forceinline__ T *fac_beta(int64_t row) const { return (T *)(rec + row * R + o_fb); }
    __host__ __device__ __forceinline__ int *fac_ea(int64_t row) const { return (int *)(rec + row * R + o_ea); }
    __host__ __device__ __forceinline__ int *fac_eb(int64_t row) const { return (int *)(rec + row * R + o_eb); }
    __host__ __device__ __forceinline__ T *fcc_alpha(int64_t row) const { return (T *)(rec + row * R + o_ca); }
    __host__ __device__ __forceinline__ T *fcc_beta(int64_t row) const { return (T *)(rec + row * R + o_cb); }
    __host__ __device__ __forceinline__ T *fcc_w(int64_t row) const { return (T *)(rec + row * R + o_cw); }
    __host__ __device__ __forceinline__ T *fcc_scale(int64_t row) const { return (T *)(rec + row * R + o_sc); }
};

template <typename T> inline Ctx<T> make_ctx(void *state, void *blob, const asg_layout_t &lay, int64_t N) {
    char *b = static_cast<char *>(blob);
    Ctx<T> c;
    c.rec = static_cast<char *>(state);
    c.R = (int64_t)lay.rec_bytes;
    c.o_fa = (int)lay.rec_fac_alpha; c.o_fb = (int)lay.rec_fac_beta; c.o_ea = (int)lay.rec_fac_ea; c.o_eb = (int)lay.rec_fac_eb;
    c.o_ca = (int)lay.rec_fcc_alpha; c.o_cb = (int)lay.rec_fcc_beta; c.o_cw = (int)lay.rec_fcc_w; c.o_sc = (int)lay.rec_fcc_scale;
    c.ehat = reinterpret_cast<T *>(b + lay.off_ehat);
    c.ehat_t = reinterpret_cast<T *>(b + lay.off_ehat_t);
    c.fmax = reinterpret_cast<T *>(b + lay.off_fmax);
    c.rmax = reinterpret_cast<T *>(b + lay.off_rmax);
    c.large_ws = lay.off_large_ws ? b + lay.off_large_ws : nullptr;
    c.fac_hd = reinterpret_cast<T *>(b + lay.off_fac_hd);
    c.il = reinterpret_cast<int64_t *>(b + lay.off_lengths);
    c.tl = c.il + N;
    c.scalars = reinterpret_cast<double *>(b + lay.off_scalars);
    c.sp = lay.sp;
    c.nblk = lay.nblk;
    c.nbp = lay.nbp;
    c.lpi = lay.lpi;
    c.tp = lay.tp;
    return c;
}

struct Dims {
    int T, N, L, S;
};

template <typename T> struct ScanArgs {
    const T *x; int64_t x_st, x_sn, x_sl;
    const T *A; int64_t A_s0, A_s1;
    const int64_t *y; int64_t y_s0, y_s1;
    const int64_t *il, *tl;     // may be NULL
    Dims d;
    Ctx<T> ctx;                 // ctx.rec == nullptr: forward-only, nothing is stored
    T *full_scores, *aligned_scores;
    const T *fmax;              // large alphabets: per-frame emission maxima [N,tp] from the prepass (NULL otherwise)
    int dir0, dir1;             // direction (0 = alpha, 1 = beta) served by blockIdx.y = 0 / 1
    int fcc_dirs, fac_dirs;     // bit d set: that lattice runs direction d
};

template <typename T> struct GradArgs {
    // upstream gradient of utterance n: gf_val * g_full[n * stride] (a NULL pointer reads as 1), likewise ga_val * g_aligned[..]
    const T *g_full; int64_t g_full_stride;
    const T *g_aligned; int64_t g_aligned_stride;
    T gf_val, ga_val;
    __device__ __forceinline__ T gfull(int64_t n) const { return g_full ? gf_val * g_full[n * g_full_stride] : gf_val; }
    __device__ __forceinline__ T galigned(int64_t n) const { return g_aligned ? ga_val * g_aligned[n * g_aligned_stride] : ga_val; }
    // asg_step(): the finalize kernel also writes loss[n] = full[n] - aligned[n] and the fused tail (sum, count)
    const T *full_scores, *aligned_scores;
    T *losses, *tail;
    Ctx<T> ctx;
    const int64_t *y; int64_t y_s0, y_s1;
    Dims d;
    T *grad_x, *grad_A;
    T *partials;                // [grid, L*L] per-CTA partial transition gradients
    int tb;                     // frames per tile
    int nstage;                 // depth of the shared-memory tile ring
    int nwarps;                 // consumer warps per CTA
    int dry;                    // diagnostics: move the data, skip the arithmetic (ASG_GRAD_DRY)
    int nopark;                 // diagnostics: serial per-warp reduction of the edge sums (ASG_GRAD_NOPARK)
    int tiles_t;                // ceil(T / tb)
    int num_tiles;              // tiles_t * N
};

// launchers (asg_scan.cu / asg_grad.cu); return cudaError_t as int
template <typename T> int launch_scan(const ScanArgs<T> &a, int roles, cudaStream_t stream);
template <typename T> int launch_grad(const GradArgs<T> &a, int parts, cudaStream_t stream);
// large-alphabet path (asg_large.cu), L > ASG_SMALL_L
constexpr int ASG_SMALL_L = 64;
template <typename T> int large_prepare(const ScanArgs<T> &a, cudaStream_t stream);
template <typename T> int launch_fcc_large(const ScanArgs<T> &a, int ndirs, cudaStream_t stream);
template <typename T> int launch_grad_large(const GradArgs<T> &a, int parts, cudaStream_t stream);
int grad_grid_size(int device);   // number of persistent CTAs asg_grad launches
int large_splitk_host(int L);
size_t large_ws_bytes(size_t elem_bytes, int64_t N, int64_t L);   // asg_large.cu     // K splits of the large-alphabet transition-gradient contraction (asg_large.cu)
void count_launch(int n = 1);
int replabel_dispatch(bool pack, const int64_t *y, int64_t y_s0, int64_t y_s1, const int64_t *tl, int64_t N, int64_t S, int64_t first_rep,
                      int64_t max_reps, int64_t fill, int64_t *out, int64_t S_out, int64_t *out_len, cudaStream_t st);
int fail(int code, const char *fmt, ...);            // records asg_last_error(), returns code
int cuda_fail(int e, const char *where);             // 0 stays 0
void release_side_slots();          // asg_scan.cu: asg_release()
// asg_viterbi.cu
int viterbi_words(int64_t S);
int viterbi_dispatch(int dtype, const void *x, int64_t x_st, int64_t x_sn, int64_t x_sl, const void *A, int64_t A_s0, int64_t A_s1,
                     const int64_t *y, int64_t y_s0, int64_t y_s1, const int64_t *il, const int64_t *tl, int64_t T, int64_t N, int64_t L,
                     int64_t S, void *scores, int64_t *path, void *scratch, cudaStream_t st);

// kernel ids for the launch counter / live profiler (asg_profile_* in include/asg_b200.h)
enum { K_FCC_SCAN = 0, K_FAC_SCAN = 1, K_GRAD = 2, K_GRAD_FINALIZE = 3, K_HOST_MISC = 4, K_GRAD_GEMM = 5, K_PEER_ALLREDUCE = 6, K_COUNT = 7 };

// RAII: counts one launch of kernel `kid` and, while profiling is enabled, brackets it with a pair
// of CUDA events recorded on the launching stream (read back by asg_profile_read()).
struct LaunchScope {
    LaunchScope(int kid, cudaStream_t st);
    ~LaunchScope();
    int kid_; cudaStream_t st_; int slot_;
};


}  // namespace asg
