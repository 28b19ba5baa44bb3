// asg_viterbi.cu -- Viterbi forced alignment: the force-aligned lattice of the ASG criterion in the TROPICAL semiring.
//
// The reference lists "Viterbi decoders" as not implemented (README.md:33); its technical report gives the semiring limit
// (doc/tech_report.tex:84-87: x (+)_T y = max(x, y)) under which the force-aligned alpha recursion (tech_report.tex:221-229,
// code: force_aligned_lattice.cpp:84-111) turns into the best-path recursion
//     delta[0,0] = x[0,y_0];   delta[0,s>0] = -inf
//     delta[t,s] = x[t,y_s] + max( delta[t-1,s] + A[y_s,y_s],  delta[t-1,s-1] + A[y_s,y_{s-1}] )
// with score delta[T_n-1, S_n-1] and the alignment read back through one back-pointer bit per (t, s).
// max-plus needs no rescaling (nothing underflows), so this runs directly in the log domain in the inputs' dtype.
//
// One CTA per utterance: thread = target positions s = tid, tid + 128, ...; delta ping-pongs in shared memory with one block
// barrier per frame; the emissions of frame t+1 are fetched (label gather straight from x, L2-resident) before the barrier of
// frame t; the back-pointer bits of a frame leave as warp ballots (bit s%32 of word s/32).  Thread 0 then walks the bits back
// from (T_n-1, S_n-1).  Ties prefer staying (move only if strictly better), like the oracle.
#include "asg_common.cuh"
#include "asg_internal.h"

namespace asg {

constexpr int VT_THREADS = 128;

template <typename T, int MAXSPT>
__global__ void __launch_bounds__(VT_THREADS) viterbi_kernel(const T *x, int64_t x_st, int64_t x_sn, int64_t x_sl, const T *A, int64_t A_s0,
                                                             int64_t A_s1, const int64_t *y, int64_t y_s0, int64_t y_s1, const int64_t *il,
                                                             const int64_t *tl, int Tt, int N, int L, int S, T *scores, int64_t *path,
                                                             unsigned *bits, int words) {
    extern __shared__ __align__(16) char vt_smem[];
    const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int SPT = (S + VT_THREADS - 1) / VT_THREADS;          // positions per thread
    T *delta = reinterpret_cast<T *>(vt_smem);                  // [2][SPT * VT_THREADS]
    const int SP = SPT * VT_THREADS;
    int Tn = il ? (int)il[n] : Tt;
    Tn = max(0, min(Tn, Tt));
    int Sn = tl ? (int)tl[n] : S;
    Sn = max(1, min(Sn, S));
    const T ninf = Num<T>::neg_inf();
    const int64_t *yn = y + (int64_t)n * y_s0;
    const T *xn = x + (int64_t)n * x_sn;
    unsigned *bn = bits + (size_t)n * Tt * words;
    int64_t *pn = path + (int64_t)n * Tt;
    for (int t = tid; t < Tt; t += VT_THREADS) pn[t] = -1;      // padding frames / infeasible utterances
    // MAXSPT = compile-time bound of SPT (1, 2, 4, 8: fully unrolled, everything in registers; 32 covers S <= 4096)
    if (Tn == 0 || Sn > Tn || SPT > MAXSPT) {
        if (tid == 0) scores[n] = (Tn == 0) ? T(0) : ninf;
        return;
    }
    // per position: label offset into a row of x, stay / move transition scores
    int64_t xo[MAXSPT];
    T Hs[MAXSPT], Ds[MAXSPT];
#pragma unroll
    for (int i = 0; i < MAXSPT; ++i) if (i < SPT) {
        const int s = i * VT_THREADS + tid;
        xo[i] = -1; Hs[i] = ninf; Ds[i] = ninf;
        if (s < Sn) {
            const int c = max(0, min(L - 1, (int)yn[(int64_t)s * y_s1]));
            xo[i] = (int64_t)c * x_sl;
            Hs[i] = A[c * A_s0 + c * A_s1];
            if (s >= 1) { const int pv = max(0, min(L - 1, (int)yn[(int64_t)(s - 1) * y_s1])); Ds[i] = A[c * A_s0 + pv * A_s1]; }
        }
    }
    // frame 0
    T xe[MAXSPT];
#pragma unroll
    for (int i = 0; i < MAXSPT; ++i) if (i < SPT) {
        const int s = i * VT_THREADS + tid;
        delta[s] = (s == 0 && xo[i] >= 0) ? xn[xo[i]] : ninf;
        xe[i] = (xo[i] >= 0 && Tn > 1) ? xn[x_st + xo[i]] : ninf;      // emissions of frame 1
    }
    __syncthreads();
    int cur = 0;
    for (int t = 1; t < Tn; ++t) {
        const T *din = delta + cur * SP;
        T *dout = delta + (cur ^ 1) * SP;
        T nxt[MAXSPT];
#pragma unroll
        for (int i = 0; i < MAXSPT; ++i) if (i < SPT) {
            const int s = i * VT_THREADS + tid;
            const T stay = din[s] + Hs[i];
            const T move = (s >= 1) ? din[s - 1] + Ds[i] : ninf;
            const bool mv = move > stay;                       // ties stay
            dout[s] = xe[i] + (mv ? move : stay);
            const unsigned b = __ballot_sync(ASG_FULL_MASK, mv);
            if (lane == 0) bn[(size_t)t * words + i * (VT_THREADS / 32) + warp] = b;
            nxt[i] = (xo[i] >= 0 && t + 1 < Tn) ? xn[(int64_t)(t + 1) * x_st + xo[i]] : ninf;
        }
#pragma unroll
        for (int i = 0; i < MAXSPT; ++i) if (i < SPT) xe[i] = nxt[i];
        cur ^= 1;
        __syncthreads();
    }
    __threadfence_block();
    if (tid == 0) {
        const T sc = delta[cur * SP + Sn - 1];
        scores[n] = sc;
        if (sc > ninf) {
            int s = Sn - 1;
            for (int t = Tn - 1; t >= 1; --t) {
                pn[t] = s;
                const unsigned wv = bn[(size_t)t * words + (s >> 5)];       // written by this CTA, ordered by the barriers above
                if ((wv >> (s & 31)) & 1u) --s;
            }
            pn[0] = s;
        }
    }
}

template <typename T>
static int launch_viterbi(const T *x, int64_t x_st, int64_t x_sn, int64_t x_sl, const T *A, int64_t A_s0, int64_t A_s1, const int64_t *y,
                          int64_t y_s0, int64_t y_s1, const int64_t *il, const int64_t *tl, int Tt, int N, int L, int S, T *scores,
                          int64_t *path, unsigned *bits, cudaStream_t st) {
    const int SPT = (S + VT_THREADS - 1) / VT_THREADS;
    const int words = SPT * (VT_THREADS / 32);
    const size_t smem = sizeof(T) * 2 * (size_t)SPT * VT_THREADS;
    auto go = [&](auto kern) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        {
            LaunchScope ls(K_HOST_MISC, st);
            kern<<<N, VT_THREADS, smem, st>>>(x, x_st, x_sn, x_sl, A, A_s0, A_s1, y, y_s0, y_s1, il, tl, Tt, N, L, S, scores, path, bits, words);
        }
        return (int)cudaGetLastError();
    };
    if (SPT <= 1) return go(viterbi_kernel<T, 1>);
    if (SPT <= 2) return go(viterbi_kernel<T, 2>);
    if (SPT <= 4) return go(viterbi_kernel<T, 4>);
    if (SPT <= 8) return go(viterbi_kernel<T, 8>);
    return go(viterbi_kernel<T, 32>);
}

int viterbi_words(int64_t S) { return (int)((S + VT_THREADS - 1) / VT_THREADS) * (VT_THREADS / 32); }

int viterbi_dispatch(int dtype, const void *x, int64_t x_st, int64_t x_sn, int64_t x_sl, const void *A, int64_t A_s0, int64_t A_s1,
                     const int64_t *y, int64_t y_s0, int64_t y_s1, const int64_t *il, const int64_t *tl, int64_t T, int64_t N, int64_t L,
                     int64_t S, void *scores, int64_t *path, void *scratch, cudaStream_t st) {
    if (dtype == ASG_F32)
        return launch_viterbi<float>((const float *)x, x_st, x_sn, x_sl, (const float *)A, A_s0, A_s1, y, y_s0, y_s1, il, tl, (int)T, (int)N, (int)L,
                                     (int)S, (float *)scores, path, (unsigned *)scratch, st);
    return launch_viterbi<double>((const double *)x, x_st, x_sn, x_sl, (const double *)A, A_s0, A_s1, y, y_s0, y_s1, il, tl, (int)T, (int)N, (int)L,
                                  (int)S, (double *)scores, path, (unsigned *)scratch, st);
}

}  // namespace asg
