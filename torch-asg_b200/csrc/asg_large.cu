// asg_large.cu -- the large-alphabet path (64 < L <= 2048), BASELINE.json's "transition tile > SMEM" config.
//
// Same arithmetic, same frame records, same C ABI as the L <= 64 kernels (asg_scan.cu / asg_grad.cu); what changes
// is where the L x L matrix lives.  E = exp(A - max A) no longer fits registers or shared memory (4 MB at L = 1024),
// so it stays in global memory -- L2-resident, read through coalesced rows of E (beta) or E^T (alpha) -- and one
// CTA advances LG_UG utterances at once so that every element of E fetched from L2 feeds LG_UG FMAs.
// The reference cannot run this configuration with a backward pass at all: its path_contrib[T-1,N,L,L] is 268 GB
// (fully_connected_lattice.cpp:26,77-78).
//
//   large_prepare        R = max A, E, E^T; per-frame emission maxima m[n,t] (the scans then form u = exp(x - m))
//   fcc_large_kernel     FCC alpha / beta scan (replaces fully_connected_lattice.cpp:9-47)
//   (FAC scans)          asg_scan.cu's FAC kernel in gather mode (ring rows hold u_t[y_s], S entries, not L)
//   grad_large_frame     per frame: FCC posterior (dense over L), FAC posterior + fixed-point label bins + edge sums,
//                        grad_x row; scales u*beta in place by the frame's 1/Z so that ...
//   cuBLAS GEMM          ... the FCC transition gradient is one plain [L x F] x [F x L] GEMM over the frame records
//   grad_large_finish    grad_A = E o G + deterministic scatter of the FAC edge sums (one thread per destination row)
#include <cublas_v2.h>
#include <climits>
#include "asg_common.cuh"
#include "asg_internal.h"

namespace asg {

constexpr int LG_THREADS = 1024;
constexpr int LG_UG = 4;          // utterances per CTA in the FCC scan
constexpr int LG_RPT = 2;         // rows per thread: L <= LG_THREADS * LG_RPT

template <typename T> struct FixL;
template <> struct FixL<float> {
    typedef int bin_t;
    static constexpr float kPost = 1073741824.0f;
    __device__ static int to_fix(float q) { return __float2int_rn(q); }
    __device__ static float from_fix(int q) { return (float)q * (1.0f / 1073741824.0f); }
    __device__ static void add(int *b, int q) { atomicAdd(b, q); }
};
template <> struct FixL<double> {
    typedef unsigned long long bin_t;
    static constexpr double kPost = 4611686018427387904.0;
    __device__ static unsigned long long to_fix(double q) { return (unsigned long long)__double2ll_rn(q); }
    __device__ static double from_fix(unsigned long long q) { return (double)(long long)q * (1.0 / 4611686018427387904.0); }
    __device__ static void add(unsigned long long *b, unsigned long long q) { atomicAdd(b, q); }
};

// block-wide reductions over LG_THREADS (or fewer) threads; `red` has >= 32 entries
template <typename T> __device__ T block_max(T v, T *red, int nwarps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    v = Num<T>::warp_max(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    T r = (lane < nwarps) ? red[lane] : Num<T>::neg_inf();
    return Num<T>::warp_max(r);
}
template <typename T> __device__ T block_sum(T v, T *red, int nwarps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    T r = (lane < nwarps) ? red[lane] : T(0);
    return warp_sum(r);
}

// ---------------------------------------------------------------------------------------------
// prepare
// ---------------------------------------------------------------------------------------------
template <typename T> __global__ void large_max_kernel(const T *A, int64_t s0, int64_t s1, int L, T *Rout) {
    __shared__ T red[32];
    T m = Num<T>::neg_inf();
    for (int64_t e = threadIdx.x; e < (int64_t)L * L; e += blockDim.x) {
        const T a = A[(e / L) * s0 + (e % L) * s1];
        if (a < -Num<T>::neg_inf()) m = tmax(m, a);
    }
    m = block_max(m, red, blockDim.x / 32);
    if (threadIdx.x == 0) Rout[0] = (m > Num<T>::neg_inf()) ? m : T(0);
}
template <typename T> __global__ void large_E_kernel(const T *A, int64_t s0, int64_t s1, int L, const T *R, T *E, T *ET) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)L * L) return;
    const int i = (int)(e / L), j = (int)(e % L);
    const T a = A[i * s0 + j * s1];
    const T v = (a == Num<T>::neg_inf()) ? T(0) : Num<T>::exp_(a - R[0]);
    E[e] = v;
    ET[(int64_t)j * L + i] = v;
}
// one warp per (utterance, frame): m[n,t] = max_l x[t,n,l] (0 for an all -inf frame)
template <typename T> __global__ void large_frame_max_kernel(ScanArgs<T> a) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int Tt = a.d.T, N = a.d.N, L = a.d.L;
    if (warp >= Tt * N) return;
    const int n = warp / Tt, t = warp % Tt;
    const T *row = a.x + (int64_t)t * a.x_st + (int64_t)n * a.x_sn;
    T m = Num<T>::neg_inf();
    for (int l = lane; l < L; l += 32) m = tmax(m, row[(int64_t)l * a.x_sl]);
    m = Num<T>::warp_max(m);
    if (lane == 0) a.ctx.fmax[(int64_t)n * a.ctx.tp + t] = (m > Num<T>::neg_inf()) ? m : T(0);
}

template <typename T> int large_prepare(const ScanArgs<T> &a, cudaStream_t st) {
    const int L = a.d.L;
    { LaunchScope ls(K_HOST_MISC, st); large_max_kernel<T><<<1, 1024, 0, st>>>(a.A, a.A_s0, a.A_s1, L, a.ctx.rmax); }
    { LaunchScope ls(K_HOST_MISC, st); large_E_kernel<T><<<(unsigned)(((int64_t)L * L + 255) / 256), 256, 0, st>>>(a.A, a.A_s0, a.A_s1, L, a.ctx.rmax, a.ctx.ehat, a.ctx.ehat_t); }
    const int64_t warps = (int64_t)a.d.T * a.d.N;
    { LaunchScope ls(K_HOST_MISC, st); large_frame_max_kernel<T><<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(a); }
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// FCC scan: one CTA = LG_UG utterances x one direction; thread = row(s) of the state vector
// ---------------------------------------------------------------------------------------------
template <typename T, bool BETA>
__device__ void fcc_large_dir(const ScanArgs<T> &a, T *pbuf /*[2][UG][L]*/, T *red /*[UG][32]*/, T *scl /*[UG]*/, int *sce /*[UG]*/) {
    const int L = a.d.L, N = a.d.N, lpi = (int)a.ctx.lpi;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nwarps = blockDim.x >> 5;
    const int n0 = blockIdx.x * LG_UG;
    const T *M = BETA ? a.ctx.ehat : a.ctx.ehat_t;      // M[j*L + i]: coalesced over the row index i
    int Tn[LG_UG];
    int Tmax = 0;
#pragma unroll
    for (int u = 0; u < LG_UG; ++u) {
        const int n = n0 + u;
        Tn[u] = 0;
        if (n < N) { Tn[u] = a.il ? (int)a.il[n] : a.d.T; Tn[u] = max(0, min(Tn[u], a.d.T)); }
        Tmax = max(Tmax, Tn[u]);
        if (!BETA && tid == 0 && n < N) a.ctx.il[n] = Tn[u];
    }
    T last[LG_RPT][LG_UG];      // the feeding vector of each utterance at its last frame (for the score)
    int esum[LG_UG];
#pragma unroll
    for (int u = 0; u < LG_UG; ++u) { esum[u] = 0; for (int r = 0; r < LG_RPT; ++r) last[r][u] = T(0); }
    if (tid < LG_UG) { scl[tid] = T(1); sce[tid] = 0; }
    __syncthreads();

    T *pin = pbuf, *pout = pbuf + (size_t)L * LG_UG;      // ping-pong, swapped at the end of every step
    for (int k = 0; k < Tmax; ++k) {
        // ---- z = M^T p for this thread's rows, all utterances of the group ----
        T z[LG_RPT][LG_UG];
#pragma unroll
        for (int r = 0; r < LG_RPT; ++r)
#pragma unroll
            for (int u = 0; u < LG_UG; ++u) z[r][u] = T(0);
        if (k > 0) {
            for (int j = 0; j < L; ++j) {
                T pj[LG_UG];
#pragma unroll
                for (int u = 0; u < LG_UG; ++u) pj[u] = pin[u * L + j];
#pragma unroll
                for (int r = 0; r < LG_RPT; ++r) {
                    const int i = tid + r * LG_THREADS;
                    if (i < L) {
                        const T e = M[(size_t)j * L + i];
#pragma unroll
                        for (int u = 0; u < LG_UG; ++u) z[r][u] = fma(e, pj[u], z[r][u]);
                    }
                }
            }
        }
        // ---- emissions, scaling, stores ----
        T lmax[LG_UG];
#pragma unroll
        for (int u = 0; u < LG_UG; ++u) lmax[u] = T(0);
#pragma unroll
        for (int u = 0; u < LG_UG; ++u) {
            const int n = n0 + u;
            const bool active = (n < N) && (k < Tn[u]);
            const int t = BETA ? (Tn[u] - 1 - k) : k;
            const T sc = scl[u];
            const int64_t row = (int64_t)n * a.ctx.tp + t;
            const T m = active ? a.ctx.fmax[row] : T(0);
#pragma unroll
            for (int r = 0; r < LG_RPT; ++r) {
                const int i = tid + r * LG_THREADS;
                T v = T(0), w = T(0);
                if (active && i < L) {
                    const T xv = a.x[(int64_t)t * a.x_st + (int64_t)n * a.x_sn + (int64_t)i * a.x_sl];
                    const T uu = Num<T>::exp_(xv - m);
                    if (k == 0) { v = BETA ? T(1) : uu; w = BETA ? uu : v; }
                    else if (BETA) { v = z[r][u] * sc; w = v * uu; }
                    else { v = z[r][u] * uu * sc; w = v; }
                }
                if (active && i < lpi && a.ctx.rec != nullptr) {
                    if (BETA) { a.ctx.fcc_beta(row)[i] = v; a.ctx.fcc_w(row)[i] = w; }
                    else a.ctx.fcc_alpha(row)[i] = v;
                }
                if (i < L) pout[u * L + i] = active ? w : T(0);
                if (active) { last[r][u] = w; lmax[u] = tmax(lmax[u], w); }
            }
            if (!BETA && active && tid == 0 && a.ctx.rec != nullptr) *a.ctx.fcc_scale(row) = (k == 0) ? T(1) : sc;
            if (active && k > 0) esum[u] += sce[u];
        }
        // ---- next step's scale per utterance: max over the block (lagged by one step) ----
#pragma unroll
        for (int u = 0; u < LG_UG; ++u) {
            const T wm = Num<T>::warp_max(lmax[u]);
            if (lane == 0) red[u * 32 + warp] = wm;
        }
        __syncthreads();
        if (warp < LG_UG) {
            T mm = (lane < nwarps) ? red[warp * 32 + lane] : T(0);
            mm = Num<T>::warp_max(mm);
            T s; int e;
            Num<T>::pow2_norm(mm, s, e);
            if (lane == 0) { scl[warp] = s; sce[warp] = e; }
        }
        __syncthreads();
        T *sw = pin; pin = pout; pout = sw;
    }
    // ---- scores ----
    const double ln2 = 0.6931471805599453094;
#pragma unroll
    for (int u = 0; u < LG_UG; ++u) {
        T part = T(0);
#pragma unroll
        for (int r = 0; r < LG_RPT; ++r) part += last[r][u];
        const T tot = block_sum(part, red, nwarps);
        const int n = n0 + u;
        if (tid == 0 && n < N) {
            double score = 0.0;
            if (Tn[u] > 0) {
                double msum = 0.0;
                for (int t = 0; t < Tn[u]; ++t) msum += (double)a.ctx.fmax[(int64_t)n * a.ctx.tp + t];
                score = (double)Num<T>::log_(tot) + msum + (double)(Tn[u] - 1) * (double)a.ctx.rmax[0] + (double)esum[u] * ln2;
            }
            a.ctx.scalars[(int64_t)n * 8 + (BETA ? SC_FULL_BETA : SC_FULL_ALPHA)] = score;
            if (BETA && a.full_scores) a.full_scores[n] = (T)score;
        }
        __syncthreads();
    }
}

template <typename T>
__global__ void __launch_bounds__(LG_THREADS) fcc_large_kernel(const ScanArgs<T> a) {
    extern __shared__ __align__(16) char lg_smem[];
    const int L = a.d.L;
    T *pbuf = (T *)lg_smem;
    T *red = pbuf + (size_t)2 * L * LG_UG;
    T *scl = red + LG_UG * 32;
    int *sce = (int *)(scl + LG_UG);
    for (int i = threadIdx.x; i < 2 * L * LG_UG; i += blockDim.x) pbuf[i] = T(0);
    __syncthreads();
    const int dir = blockIdx.y == 0 ? a.dir0 : a.dir1;
    if (dir == 0) fcc_large_dir<T, false>(a, pbuf, red, scl, sce);
    else fcc_large_dir<T, true>(a, pbuf, red, scl, sce);
}

template <typename T> int launch_fcc_large(const ScanArgs<T> &a, int ndirs, cudaStream_t st) {
    const int L = a.d.L;
    const size_t bytes = sizeof(T) * ((size_t)2 * L * LG_UG + LG_UG * 32 + LG_UG) + sizeof(int) * LG_UG + 64;
    cudaError_t e = cudaFuncSetAttribute(fcc_large_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((a.d.N + LG_UG - 1) / LG_UG, ndirs);
    {
        LaunchScope ls(K_FCC_SCAN, st);
        fcc_large_kernel<T><<<grid, LG_THREADS, bytes, st>>>(a);
    }
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// gradient assembly
// ---------------------------------------------------------------------------------------------
constexpr int GL_THREADS = 256;
constexpr int GL_FRAMES = 8;      // frames per CTA

// One CTA = (utterance, GL_FRAMES consecutive frames).  Per frame: FAC posterior by warp 0 (lane = 8-position block,
// S <= 256) -> fixed-point label bins in shared memory + edge sums; FCC posterior dense over L by all threads;
// grad_x row written once; u*beta scaled in place by gf * scale_t / Z_t for the GEMM.
template <typename T, bool FCC, bool FAC>
__global__ void __launch_bounds__(GL_THREADS) grad_large_frame_kernel(const GradArgs<T> a, unsigned long long *edge /*[2][N][sp]*/) {
    extern __shared__ __align__(16) char gl_smem[];
    typedef FixL<T> FX;
    typedef typename FX::bin_t bin_t;
    const int L = a.d.L, N = a.d.N, Tt = a.d.T;
    const int sp = (int)a.ctx.sp, nblk = (int)a.ctx.nblk, lpi = (int)a.ctx.lpi;
    const int64_t tp = a.ctx.tp;
    bin_t *bins = (bin_t *)gl_smem;                         // [L]
    T *red = (T *)(bins + L);                               // [32]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = blockIdx.y;
    const int t_begin = blockIdx.x * GL_FRAMES;
    const int Tn = (int)a.ctx.il[n];
    const T gf = FCC ? a.gfull(n) : T(0);
    const T ga = FAC ? a.galigned(n) : T(0);

    // FAC constants of this lane's block (warp 0)
    T hreg[8], dreg[8], accH[8], accQ[8];
    int lab[8];
    int Sn = 0;
    if (FAC && warp == 0) {
        Sn = (int)a.ctx.tl[n];
        const int64_t *yn = a.y + (int64_t)n * a.y_s0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int s = lane * 8 + j;
            lab[j] = -1; hreg[j] = dreg[j] = accH[j] = accQ[j] = T(0);
            if (s < Sn) {
                lab[j] = max(0, min(L - 1, (int)yn[(int64_t)s * a.y_s1]));
                hreg[j] = a.ctx.fac_hd[(int64_t)n * sp + s];
                if (s >= 1) dreg[j] = a.ctx.fac_hd[((int64_t)N + n) * sp + s - 1];
            }
        }
    }
    for (int l = tid; l < L; l += GL_THREADS) bins[l] = 0;
    __syncthreads();

    for (int t = t_begin; t < min(t_begin + GL_FRAMES, (int)tp); ++t) {
        const int64_t row = (int64_t)n * tp + t;
        T *gx_row = a.grad_x + ((int64_t)t * N + n) * L;
        if (t >= Tn) {                                      // padding frame / pitch padding: exact zeros, and keep the GEMM operands clean
            if (t < Tt) for (int l = tid; l < L; l += GL_THREADS) gx_row[l] = T(0);
            if (FCC) for (int l = tid; l < lpi; l += GL_THREADS) { a.ctx.fcc_w(row)[l] = T(0); a.ctx.fcc_alpha(row)[l] = T(0); }
            continue;
        }
        // ---- FAC (warp 0): posterior -> bins, edge sums ----
        if (FAC && warp == 0) {
            const bool act = lane < nblk;
            const int b8 = min(lane, nblk - 1) * 8, be = min(lane, nblk - 1);
            T av[8], bv[8];
            const T *pa = a.ctx.fac_alpha(row) + b8, *pb = a.ctx.fac_beta(row) + b8;
#pragma unroll
            for (int j = 0; j < 8; ++j) { av[j] = pa[j]; bv[j] = pb[j]; }
            int ea, eb, da, db;
            fac_unpack(a.ctx.fac_ea(row)[be], ea, da);
            fac_unpack(a.ctx.fac_eb(row)[be], eb, db);
            const T fa = Num<T>::exp2i(da + Num<T>::kHalfProd), fb = Num<T>::exp2i(db + Num<T>::kHalfProd);
            T pm = T(0), ps = T(0);
#pragma unroll
            for (int j = 0; j < 8; ++j) { bv[j] = (av[j] * fa) * (bv[j] * fb); pm = tmax(pm, bv[j]); ps += bv[j]; }
            const int c = ea + eb;
            const int ce = (act && pm > T(0)) ? c + Num<T>::pexp(pm) : INT_MIN;
            const int cmax = __reduce_max_sync(ASG_FULL_MASK, ce);
            const int sh = Num<T>::block_shift(c, cmax);
            const T Z = warp_sum((ce != INT_MIN) ? Num<T>::ldexp_pos(ps, sh) : T(0));
            const T kq = (Z > Num<T>::tiny() && ce != INT_MIN) ? Num<T>::ldexp_pos(T(1) / Z, sh + Num<T>::kPostLog2) : T(0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                bv[j] = bv[j] * kq;
                if (lab[j] >= 0) FX::add(&bins[lab[j]], FX::to_fix(bv[j]));
            }
            if (t >= 1) {
                T pv[8];
                const T *pp = a.ctx.fac_alpha(row - 1) + b8;
#pragma unroll
                for (int j = 0; j < 8; ++j) pv[j] = pp[j];
                int ep, dp_;
                fac_unpack(a.ctx.fac_ea(row - 1)[be], ep, dp_);
                ep += dp_;
                T left = __shfl_up_sync(ASG_FULL_MASK, pv[7], 1);
                const int el = __shfl_up_sync(ASG_FULL_MASK, ep, 1);
                left = (lane == 0) ? T(0) : left * Num<T>::exp2i(min(el - ep, Num<T>::kMaxShift - 8));
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const T stay = hreg[j] * pv[j];
                    const T den = fma(dreg[j], j ? pv[j ? j - 1 : 0] : left, stay);
                    const T fs = stay / tmax(den, Num<T>::tiny());
                    accH[j] = fma(bv[j], fs, accH[j]);
                    accQ[j] += bv[j];
                }
            }
        }
        // ---- FCC: Z = sum alpha*beta (dense), posterior, scale u*beta in place ----
        T inv = T(0);
        if (FCC) {
            T part = T(0);
            const T *al = a.ctx.fcc_alpha(row), *be = a.ctx.fcc_beta(row);
            for (int l = tid; l < L; l += GL_THREADS) part += al[l] * be[l];
            const T Z = block_sum(part, red, GL_THREADS / 32);
            inv = (Z > T(0)) ? gf / Z : T(0);
        } else {
            __syncthreads();
        }
        __syncthreads();                                     // bins complete
        {
            const T gs = ga * (T(1) / FX::kPost);
            const T *al = a.ctx.fcc_alpha(row), *be = a.ctx.fcc_beta(row);
            T *w = a.ctx.fcc_w(row);
            const T winv = (FCC && t >= 1) ? inv * *a.ctx.fcc_scale(row) : T(0);
            for (int l = tid; l < L; l += GL_THREADS) {
                T g = T(0);
                if (FCC) { g = al[l] * be[l] * inv; w[l] = w[l] * winv; }
                if (FAC) { g = fma(gs, (T)(long long)bins[l], g); bins[l] = 0; }
                gx_row[l] = g;
            }
        }
        __syncthreads();
    }
    // ---- FAC edge sums of this CTA's frames -> global fixed point (integer atomics: order independent) ----
    if (FAC && warp == 0 && lane < nblk) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int s = lane * 8 + j;
            if (lab[j] >= 0) {
                const T sc = (sizeof(T) == 4) ? T(4) : T(1.0 / 16777216.0);      // kPost units -> 2^32 (2^38) per unit mass
                atomicAdd(&edge[(int64_t)n * sp + s], (unsigned long long)(long long)llrint((double)(accH[j] * sc)));
                atomicAdd(&edge[((int64_t)N + n) * sp + s], (unsigned long long)(long long)llrint((double)((accQ[j] - accH[j]) * sc)));
            }
        }
    }
}

// grad_A = E o G (+ FAC scatter).  One thread per destination row i walks all target positions in a fixed order.
template <typename T, bool FCC, bool FAC>
__global__ void grad_large_finish_kernel(const GradArgs<T> a, const T *G, const unsigned long long *edge) {
    const int L = a.d.L, N = a.d.N, sp = (int)a.ctx.sp;
    if (a.losses != nullptr && blockIdx.x == gridDim.x - 1) {      // asg_step(): the extra CTA writes the losses and the fused tail
        __shared__ T red8[8];
        write_losses_block<T>(a.full_scores, a.aligned_scores, a.losses, N, a.tail, red8);
        return;
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L) return;
    T *out = a.grad_A + (int64_t)i * L;
    for (int j = 0; j < L; ++j) out[j] = FCC ? a.ctx.ehat[(int64_t)i * L + j] * G[(int64_t)i * L + j] : T(0);
    if (FAC) {
        const double unfix = (sizeof(T) == 4) ? (1.0 / 4294967296.0) : (1.0 / 274877906944.0);
        for (int n = 0; n < N; ++n) {
            const int Sn = (int)a.ctx.tl[n];
            const T ga = a.galigned(n);
            const int64_t *yn = a.y + (int64_t)n * a.y_s0;
            int prev = -1;
            for (int s = 0; s < Sn; ++s) {
                const int l = max(0, min(L - 1, (int)yn[(int64_t)s * a.y_s1]));
                if (l == i) {
                    out[i] += ga * (T)((double)(long long)edge[(int64_t)n * sp + s] * unfix);
                    if (s >= 1) out[prev] += ga * (T)((double)(long long)edge[((int64_t)N + n) * sp + s] * unfix);
                }
                prev = l;
            }
        }
    }
}

static cublasHandle_t cublas_handle(int device) {
    static cublasHandle_t h[64] = {0};
    if (device < 0 || device >= 64) return nullptr;
    if (!h[device] && cublasCreate(&h[device]) != CUBLAS_STATUS_SUCCESS) h[device] = nullptr;
    return h[device];
}
static cublasStatus_t gemm_nt(cublasHandle_t h, int L, int F, const float *Ap, int lda, const float *Bp, int ldb, float *C) {
    const float one = 1.f, zero = 0.f;
    return cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, L, L, F, &one, Ap, lda, Bp, ldb, &zero, C, L);
}
static cublasStatus_t gemm_nt(cublasHandle_t h, int L, int F, const double *Ap, int lda, const double *Bp, int ldb, double *C) {
    const double one = 1.0, zero = 0.0;
    return cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, L, L, F, &one, Ap, lda, Bp, ldb, &zero, C, L);
}

template <typename T, bool FCC, bool FAC>
static int launch_grad_large_t(const GradArgs<T> &a, cudaStream_t st) {
    const int L = a.d.L, N = a.d.N;
    int device = 0;
    cudaGetDevice(&device);
    // scratch: G [L,L] | edge sums [2,N,sp] int64
    T *G = a.partials;
    unsigned long long *edge = (unsigned long long *)((char *)a.partials + align_up(sizeof(T) * (size_t)L * L, 256));
    cudaError_t e = cudaMemsetAsync(edge, 0, sizeof(unsigned long long) * 2 * (size_t)N * a.ctx.sp, st);
    if (e != cudaSuccess) return (int)e;
    const size_t smem = (size_t)L * 8 + 32 * sizeof(T) + 64;
    if ((e = cudaFuncSetAttribute(grad_large_frame_kernel<T, FCC, FAC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
    dim3 grid((unsigned)((a.ctx.tp + GL_FRAMES - 1) / GL_FRAMES), N);
    {
        LaunchScope ls(K_GRAD, st);
        grad_large_frame_kernel<T, FCC, FAC><<<grid, GL_THREADS, smem, st>>>(a, edge);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
    if (FCC) {
        // G[i][j] = sum_f w'[f][i] * alpha[f-1][j] over all frame records f >= 1 (w' is zero at t = 0 and on padding):
        // a plain GEMM straight on the records (leading dimension = record stride)
        cublasHandle_t h = cublas_handle(device);
        if (!h) return (int)cudaErrorUnknown;
        cublasSetStream(h, st);
        const int ld = (int)(a.ctx.R / (int64_t)sizeof(T));
        const int64_t F = (int64_t)N * a.ctx.tp - 1;
        count_launch();
        if (gemm_nt(h, L, (int)F, a.ctx.fcc_alpha(0), ld, a.ctx.fcc_w(1), ld, G) != CUBLAS_STATUS_SUCCESS) return (int)cudaErrorUnknown;
    }
    {
        LaunchScope ls(K_GRAD_FINALIZE, st);
        grad_large_finish_kernel<T, FCC, FAC><<<(L + 255) / 256 + (a.losses ? 1 : 0), 256, 0, st>>>(a, G, edge);
    }
    return (int)cudaGetLastError();
}

template <typename T> int launch_grad_large(const GradArgs<T> &a, int parts, cudaStream_t st) {
    if (parts == ASG_PART_BOTH) return launch_grad_large_t<T, true, true>(a, st);
    if (parts == ASG_PART_FCC) return launch_grad_large_t<T, true, false>(a, st);
    return launch_grad_large_t<T, false, true>(a, st);
}

template int large_prepare<float>(const ScanArgs<float> &, cudaStream_t);
template int large_prepare<double>(const ScanArgs<double> &, cudaStream_t);
template int launch_fcc_large<float>(const ScanArgs<float> &, int, cudaStream_t);
template int launch_fcc_large<double>(const ScanArgs<double> &, int, cudaStream_t);
template int launch_grad_large<float>(const GradArgs<float> &, int, cudaStream_t);
template int launch_grad_large<double>(const GradArgs<double> &, int, cudaStream_t);

}  // namespace asg
