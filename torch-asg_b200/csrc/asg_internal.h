// asg_internal.h -- host/device structs shared between the translation units of libasg_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/asg_b200.h"

namespace asg {

// typed view of the context blob (see asg_layout_t)
template <typename T> struct Ctx {
    T *ehat;          // [L,L]
    T *fcc_scale;     // [N,tp]
    T *fcc_w;         // [N,tp,lpi]
    T *fac_hd;        // [2,N,sp]   h_s, d_s (edge s -> s+1)
    int *fac_ea;      // [N,tp,nbp]
    int *fac_eb;      // [N,tp,nbp]
    int64_t *il;      // [N]
    int64_t *tl;      // [N]
    double *scalars;  // [N,8]
    int64_t sp, nblk, nbp, lpi, tp;
};

// per-utterance float64 bookkeeping slots in Ctx::scalars
enum { SC_R = 0, SC_FULL_BETA = 1, SC_FULL_ALPHA = 2, SC_RB = 3, SC_AL_BETA = 4, SC_AL_ALPHA = 5,
       SC_FCC_LOGSCALE_END = 6, SC_SPARE = 7 };

template <typename T> inline Ctx<T> make_ctx(void *blob, const asg_layout_t &lay, int64_t N) {
    char *b = static_cast<char *>(blob);
    Ctx<T> c;
    c.ehat = reinterpret_cast<T *>(b + lay.off_ehat);
    c.fcc_scale = reinterpret_cast<T *>(b + lay.off_fcc_scale);
    c.fcc_w = reinterpret_cast<T *>(b + lay.off_fcc_w);
    c.fac_hd = reinterpret_cast<T *>(b + lay.off_fac_hd);
    c.fac_ea = reinterpret_cast<int *>(b + lay.off_fac_ea);
    c.fac_eb = reinterpret_cast<int *>(b + lay.off_fac_eb);
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
    T *fcc_alpha, *fcc_beta;    // [N,tp,lpi]
    T *fac_alpha, *fac_beta;    // [N,tp,sp]
    Ctx<T> ctx;
    T *full_scores, *aligned_scores;
    int dir0, dir1;             // direction (0 = alpha, 1 = beta) served by blockIdx.y = 0 / 1
    int fcc_dirs, fac_dirs;     // bit d set: that lattice runs direction d
};

template <typename T> struct GradArgs {
    const T *g_full; int64_t g_full_stride;
    const T *g_aligned; int64_t g_aligned_stride;
    const T *fcc_alpha, *fcc_beta, *fac_alpha, *fac_beta;
    Ctx<T> ctx;
    const int64_t *y; int64_t y_s0, y_s1;
    Dims d;
    T *grad_x, *grad_A;
    T *partials;                // [grid, L*L] per-CTA partial transition gradients
    int tb;                     // frames per tile
    int nstage;                 // depth of the shared-memory tile ring
    int nwarps;                 // consumer warps per CTA
    int dry;                    // diagnostics: move the data, skip the arithmetic (ASG_GRAD_DRY)
    int tiles_t;                // ceil(T / tb)
    int num_tiles;              // tiles_t * N
};

// launchers (asg_scan.cu / asg_grad.cu); return cudaError_t as int
template <typename T> int launch_scan(const ScanArgs<T> &a, int roles, cudaStream_t stream);
template <typename T> int launch_grad(const GradArgs<T> &a, int parts, cudaStream_t stream);
int grad_grid_size(int device);   // number of persistent CTAs asg_grad launches
void count_launch(int n = 1);

// kernel ids for the launch counter / live profiler (asg_profile_* in include/asg_b200.h)
enum { K_FCC_SCAN = 0, K_FAC_SCAN = 1, K_GRAD = 2, K_GRAD_FINALIZE = 3, K_HOST_MISC = 4, K_COUNT = 5 };

// RAII: counts one launch of kernel `kid` and, while profiling is enabled, brackets it with a pair
// of CUDA events recorded on the launching stream (read back by asg_profile_read()).
struct LaunchScope {
    LaunchScope(int kid, cudaStream_t st);
    ~LaunchScope();
    int kid_; cudaStream_t st_; int slot_;
};

}  // namespace asg
