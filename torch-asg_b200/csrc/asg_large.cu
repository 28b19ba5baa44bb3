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
//   grad_large_gemm      ... the FCC transition gradient is one [L x F] x [F x L] contraction over the frame records
//                        (hand-written: 128 x 128 tiles, 8 x 8 register tiles, cp.async ring, split-K)
//   grad_large_finish    grad_A = E o G + deterministic scatter of the FAC edge sums (one thread per destination row)
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

// ---------------------------------------------------------------------------------------------
// FCC scan, large alphabets, fast variant: the step of ALL utterances of a slab at once is a dense product
//     Z[L x NU] = M[L x L] * P[L x NU]        (M = E for alpha, E^T for beta; P = last step's state vectors, one column each)
// so the grid tiles it like a GEMM and keeps its tile of M RESIDENT in shared memory for the whole scan: CTA (rs, us, dir)
// owns LG2_RS rows of M (128 KB at L = 1024) and LG2_NU utterances, streams P from L2 through a 3-stage cp.async ring,
// 4 x 4 register tile per thread, and finishes its rows (emission, lagged power-of-two scale per utterance, frame records)
// in the epilogue.  The CTAs that share a (slab of utterances, direction) meet once per frame in a spin barrier on a global
// counter -- every CTA is resident (one per SM, grid <= SM count is checked at launch) -- and exchange the new state
// through L2 (cp.async.cg / ld.cg: never through L1).  The per-utterance maximum that sets the next scale is an integer
// atomicMax per (direction, utterance) in a 3-deep rotation (written at step k, read at k+1, cleared at k+2).
// Roofline: fp32 FMA (2 L^2 N flops per frame and direction: 268 GFLOP per forward at cfg-4), not HBM.
// ---------------------------------------------------------------------------------------------
constexpr int LG2_RS = 32;         // rows of M per CTA
constexpr int LG2_NU = 64;         // utterances per slab
constexpr int LG2_KC = 32;         // k per stage
constexpr int LG2_ST = 3;
constexpr int LG2_THREADS = 128;   // (RS/4) x (NU/4) threads, 4 x 4 outputs each

struct Large2Ws {                  // carved out of the context blob (asg_layout_t::off_large_ws)
    size_t o_P, o_max, o_bar, o_part, bytes;
    int npad, nrs;
};
__host__ __device__ inline Large2Ws large2_ws(size_t eb, int64_t N, int64_t L) {
    Large2Ws w;
    w.npad = (int)((N + LG2_NU - 1) / LG2_NU * LG2_NU);
    w.nrs = (int)((L + LG2_RS - 1) / LG2_RS);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    w.o_P = take(eb * 2 * 2 * (size_t)L * w.npad);                 // [dir][parity][L][npad]
    w.o_max = take(sizeof(unsigned long long) * 2 * 3 * (size_t)w.npad);     // [dir][3][npad] maxima (bit patterns)
    w.o_bar = take(sizeof(unsigned) * 2 * 64);                       // [dir][slab lane] arrival counters
    w.o_part = take(eb * 2 * (size_t)w.nrs * w.npad);              // [dir][rs][npad] partial sums for the score
    w.bytes = off;
    return w;
}
size_t large_ws_bytes(size_t eb, int64_t N, int64_t L) { return large2_ws(eb, N, L).bytes; }

template <typename T> struct UBits;
template <> struct UBits<float> {
    typedef unsigned type;
    __device__ static unsigned of(float v) { return __float_as_uint(v); }
    __device__ static float back(unsigned b) { return __uint_as_float(b); }
};
template <> struct UBits<double> {
    typedef unsigned long long type;
    __device__ static unsigned long long of(double v) { return (unsigned long long)__double_as_longlong(v); }
    __device__ static double back(unsigned long long b) { return __longlong_as_double((long long)b); }
};

template <typename T>
__global__ void __launch_bounds__(LG2_THREADS, 1) fcc_large2_kernel(const ScanArgs<T> a, char *ws_base, int nus_par, int dir_first) {
    extern __shared__ __align__(16) char l2_smem[];
    typedef typename UBits<T>::type ubits;
    const int L = a.d.L, N = a.d.N, Tt = a.d.T, lpi = (int)a.ctx.lpi;
    const Large2Ws w = large2_ws(sizeof(T), N, L);
    const int npad = w.npad, nus = npad / LG2_NU;
    const int rs = blockIdx.x, usl = blockIdx.y, dirz = blockIdx.z;
    const bool BETA = (dirz == 0 ? dir_first : 1) == 1;
    const int tid = threadIdx.x;
    const int i0 = rs * LG2_RS;
    const int MP = (L + LG2_KC - 1) / LG2_KC * LG2_KC + 4;           // pitch of the resident M tile (elements); columns >= L are zero
    T *Ms = reinterpret_cast<T *>(l2_smem);                          // [LG2_RS][MP]
    T *Ps = Ms + (size_t)LG2_RS * MP;                                // [LG2_ST][LG2_KC][LG2_NU]
    T *sc_s = Ps + (size_t)LG2_ST * LG2_KC * LG2_NU;                 // [LG2_NU] scale applied at this step
    int *e_s = reinterpret_cast<int *>(sc_s + LG2_NU);               // [LG2_NU] its exponent
    int *esum_s = e_s + LG2_NU;                                      // [LG2_NU] running exponent sum
    int *tn_s = esum_s + LG2_NU;                                     // [LG2_NU] utterance lengths
    ubits *mx_s = reinterpret_cast<ubits *>(tn_s + LG2_NU);          // [LG2_NU] this CTA's maxima of the new vector

    T *Pg = reinterpret_cast<T *>(ws_base + w.o_P) + (size_t)dirz * 2 * L * npad;                 // [parity][L][npad]
    unsigned long long *maxg = reinterpret_cast<unsigned long long *>(ws_base + w.o_max) + (size_t)dirz * 3 * npad;
    unsigned *bar = reinterpret_cast<unsigned *>(ws_base + w.o_bar) + dirz * 64 + usl;
    T *partg = reinterpret_cast<T *>(ws_base + w.o_part) + ((size_t)dirz * w.nrs + rs) * npad;

    // the resident tile of M: rows i0 .. i0+RS-1 of E (alpha) / E^T (beta); rows beyond L are zero
    const T *Msrc = BETA ? a.ctx.ehat_t : a.ctx.ehat;
    for (int e = tid; e < LG2_RS * MP; e += LG2_THREADS) {
        const int r = e / MP, c = e - r * MP;
        Ms[e] = (i0 + r < L && c < L) ? Msrc[(size_t)(i0 + r) * L + c] : T(0);
    }
    const int rg = tid >> 4, ug = tid & 15;                          // 8 row groups x 16 utterance groups
    const int r0 = rg * 4, u0 = ug * 4;
    const double ln2 = 0.6931471805599453094;
    unsigned arrivals = 0;                                           // barriers this CTA has taken part in

    for (int us = usl; us < nus; us += nus_par) {                    // the slabs this CTA serves, one after the other
        const int n0 = us * LG2_NU;
        int Tmax = 0;
        __syncthreads();
        if (tid < LG2_NU) {
            const int n = n0 + tid;
            int tn = 0;
            if (n < N) { tn = a.il ? (int)a.il[n] : Tt; tn = max(0, min(tn, Tt)); }
            tn_s[tid] = tn;
            esum_s[tid] = 0;
            if (!BETA && rs == 0 && n < N) a.ctx.il[n] = tn;
        }
        __syncthreads();
        for (int u = 0; u < LG2_NU; ++u) Tmax = max(Tmax, tn_s[u]);
        T last[4][4];                                               // the feeding vector at each utterance's last frame
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int u = 0; u < 4; ++u) last[r][u] = T(0);

        for (int k = 0; k < Tmax; ++k) {
            const T *Pin = Pg + (size_t)((k + 1) & 1) * L * npad + n0;      // written at step k-1
            T *Pout = Pg + (size_t)(k & 1) * L * npad + n0;
            unsigned long long *mx_w = maxg + (size_t)(k % 3) * npad + n0, *mx_r = maxg + (size_t)((k + 2) % 3) * npad + n0,
                               *mx_c = maxg + (size_t)((k + 1) % 3) * npad + n0;
            // ---- wait for step k-1 of every CTA of this (slab, direction) ----
            if (k > 0) {
                if (tid == 0) {
                    const unsigned target = arrivals * gridDim.x;
                    while (true) {
                        unsigned v;
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(bar) : "memory");
                        if (v >= target) break;
                        __nanosleep(40);
                    }
                }
                __syncthreads();
            }
            if (tid < LG2_NU) {
                T sc = T(1); int e = 0;
                if (k > 0) {
                    unsigned long long mb;
                    asm volatile("ld.global.cg.u64 %0, [%1];\n" : "=l"(mb) : "l"(mx_r + tid) : "memory");
                    Num<T>::pow2_norm(UBits<T>::back((ubits)mb), sc, e);
                }
                sc_s[tid] = sc; e_s[tid] = e;
                if (k > 0 && k < tn_s[tid]) esum_s[tid] += e;
                mx_s[tid] = 0;
                if (rs == 0) mx_c[tid] = 0ull;                       // the rotation slot written at step k+1
            }
            // ---- Z tile = M tile * P slab ----
            T acc[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int u = 0; u < 4; ++u) acc[r][u] = T(0);
            if (k > 0) {
                const int nch = (L + LG2_KC - 1) / LG2_KC;
                constexpr int VE = 16 / (int)sizeof(T);
                auto issue = [&](int c) {
                    if (c < nch) {
                        T *dst = Ps + (size_t)(c % LG2_ST) * LG2_KC * LG2_NU;
                        for (int v = tid; v < LG2_KC * LG2_NU / VE; v += LG2_THREADS) {
                            const int kk = v / (LG2_NU / VE), col = (v - kk * (LG2_NU / VE)) * VE;
                            const int krow = c * LG2_KC + kk;
                            const T *src = Pin + (size_t)krow * npad + col;
                            const int bytes = krow < L ? 16 : 0;
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(dst + kk * LG2_NU + col)), "l"(bytes ? src : Pin), "r"(bytes) : "memory");
                        }
                    }
                    cp_async_commit();
                };
                for (int c = 0; c < LG2_ST - 1; ++c) issue(c);
                for (int c = 0; c < nch; ++c) {
                    cp_async_wait<LG2_ST - 2>();
                    __syncthreads();
                    issue(c + LG2_ST - 1);
                    const T *Pt = Ps + (size_t)(c % LG2_ST) * LG2_KC * LG2_NU;
                    const T *Mt = Ms + (size_t)r0 * MP + c * LG2_KC;
#pragma unroll
                    for (int kq = 0; kq < LG2_KC; kq += 4) {
                        T mv[4][4], pv[4][4];
#pragma unroll
                        for (int r = 0; r < 4; ++r)
#pragma unroll
                            for (int q = 0; q < 4; ++q) mv[r][q] = Mt[(size_t)r * MP + kq + q];
#pragma unroll
                        for (int q = 0; q < 4; ++q)
#pragma unroll
                            for (int u = 0; u < 4; ++u) pv[q][u] = Pt[(kq + q) * LG2_NU + u0 + u];
#pragma unroll
                        for (int q = 0; q < 4; ++q)
#pragma unroll
                            for (int r = 0; r < 4; ++r)
#pragma unroll
                                for (int u = 0; u < 4; ++u) acc[r][u] = fma(mv[r][q], pv[q][u], acc[r][u]);
                    }
                }
                cp_async_wait<0>();
            }
            __syncthreads();                                         // sc_s / mx_s ready; the P ring is free again
            // ---- epilogue: emissions, scale, records, the new state vector ----
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int ul = u0 + u, n = n0 + ul;
                const int tn = tn_s[ul];
                const bool active = (n < N) && (k < tn);
                const int t = BETA ? (tn - 1 - k) : k;
                const int64_t row = (int64_t)n * a.ctx.tp + t;
                const T sc = sc_s[ul];
                const T m = active ? a.ctx.fmax[row] : T(0);
                T lm = T(0);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int i = i0 + r0 + r;
                    T v = T(0), fd = T(0);
                    if (active && i < L) {
                        const T xv = a.x[(int64_t)t * a.x_st + (int64_t)n * a.x_sn + (int64_t)i * a.x_sl];
                        const T uu = Num<T>::exp_(xv - m);
                        if (k == 0) { v = BETA ? T(1) : uu; fd = BETA ? uu : v; }
                        else if (BETA) { v = acc[r][u] * sc; fd = v * uu; }
                        else { v = acc[r][u] * uu * sc; fd = v; }
                    }
                    if (active && i < lpi && a.ctx.rec != nullptr) {
                        if (BETA) { a.ctx.fcc_beta(row)[i] = v; a.ctx.fcc_w(row)[i] = fd; }
                        else a.ctx.fcc_alpha(row)[i] = v;
                    }
                    if (i < L) __stcg(Pout + (size_t)i * npad + ul, fd);
                    if (active) { last[r][u] = fd; lm = tmax(lm, fd); }
                }
                if (!BETA && active && rs == 0 && rg == 0 && a.ctx.rec != nullptr) *a.ctx.fcc_scale(row) = (k == 0) ? T(1) : sc;
                atomicMax(&mx_s[ul], UBits<T>::of(lm));                  // non-negative values: bit order = value order
            }
            __syncthreads();
            if (tid < LG2_NU) atomicMax(mx_w + tid, (unsigned long long)mx_s[tid]);
            // ---- arrive ----
            __threadfence();
            __syncthreads();
            ++arrivals;
            if (tid == 0) atomicAdd(bar, 1u);
        }
        // ---- scores: sum of the last feeding vector over all rows, fixed order over the row slabs ----
        __syncthreads();
        T *red = Ps;                                                 // [8][LG2_NU] partial sums of the row groups
#pragma unroll
        for (int u = 0; u < 4; ++u) red[rg * LG2_NU + u0 + u] = (last[0][u] + last[1][u]) + (last[2][u] + last[3][u]);
        __syncthreads();
        if (tid < LG2_NU) {
            T sum = T(0);
            for (int g = 0; g < 8; ++g) sum += red[g * LG2_NU + tid];
            __stcg(partg + n0 + tid, sum);
        }
        __threadfence();
        __syncthreads();
        ++arrivals;
        if (tid == 0) atomicAdd(bar, 1u);
        if (rs == 0) {
            if (tid == 0) {
                const unsigned target = arrivals * gridDim.x;
                while (true) {
                    unsigned v;
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(bar) : "memory");
                    if (v >= target) break;
                    __nanosleep(40);
                }
            }
            __syncthreads();
            if (tid < LG2_NU && n0 + tid < N) {
                const int n = n0 + tid, tn = tn_s[tid];
                double score = 0.0;
                if (tn > 0) {
                    T tot = T(0);
                    const T *pp = reinterpret_cast<const T *>(ws_base + w.o_part) + (size_t)dirz * w.nrs * npad + n;
                    for (int q = 0; q < (int)gridDim.x; ++q) tot += __ldcg(pp + (size_t)q * npad);
                    double msum = 0.0;
                    for (int t = 0; t < tn; ++t) msum += (double)a.ctx.fmax[(int64_t)n * a.ctx.tp + t];
                    score = (double)Num<T>::log_(tot) + msum + (double)(tn - 1) * (double)a.ctx.rmax[0] + (double)esum_s[tid] * ln2;
                }
                a.ctx.scalars[(int64_t)n * 8 + (BETA ? SC_FULL_BETA : SC_FULL_ALPHA)] = score;
                if (BETA && a.full_scores) a.full_scores[n] = (T)score;
            }
        }
    }
}

template <typename T> static bool launch_fcc_large2(const ScanArgs<T> &a, int ndirs, cudaStream_t st, int *rc) {
    const int L = a.d.L, N = a.d.N;
    const Large2Ws w = large2_ws(sizeof(T), N, L);
    const size_t smem = sizeof(T) * ((size_t)LG2_RS * ((L + LG2_KC - 1) / LG2_KC * LG2_KC + 4) + (size_t)LG2_ST * LG2_KC * LG2_NU + LG2_NU) + sizeof(int) * 3 * LG2_NU +
                        sizeof(typename UBits<T>::type) * LG2_NU + 64;
    int device = 0, sms = 0;
    cudaGetDevice(&device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int nus = w.npad / LG2_NU;
    int nus_par = sms / (ndirs * w.nrs);
    if (getenv("ASG_LARGE_V1") || smem > 200 * 1024 || nus_par < 1 || a.ctx.large_ws == nullptr) return false;
    if (nus_par > nus) nus_par = nus;
    if (nus_par > 64) nus_par = 64;
    cudaError_t e = cudaFuncSetAttribute(fcc_large2_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { *rc = (int)e; return true; }
    // counters and maxima start at zero (stream-ordered; the state vectors are written before they are read)
    if ((e = cudaMemsetAsync(a.ctx.large_ws + w.o_max, 0, w.o_part - w.o_max, st)) != cudaSuccess) { *rc = (int)e; return true; }
    // Its CTAs spin on each other, so two instances must never share the GPU half-resident each (they would wait for SMs
    // the other one holds): launches from different streams are chained through one event per device.  (Inside a stream
    // capture the launches of a graph are ordered by the graph itself.)
    static cudaEvent_t chain[64] = {nullptr};
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    const bool chained = cap == cudaStreamCaptureStatusNone && device >= 0 && device < 64;
    if (chained) {
        if (!chain[device]) cudaEventCreateWithFlags(&chain[device], cudaEventDisableTiming);
        else cudaStreamWaitEvent(st, chain[device], 0);
    }
    dim3 grid(w.nrs, nus_par, ndirs);
    {
        LaunchScope ls(K_FCC_SCAN, st);
        fcc_large2_kernel<T><<<grid, LG2_THREADS, smem, st>>>(a, a.ctx.large_ws, nus_par, a.dir0);
    }
    *rc = (int)cudaGetLastError();
    if (chained && chain[device]) cudaEventRecord(chain[device], st);
    return true;
}

template <typename T> int launch_fcc_large(const ScanArgs<T> &a, int ndirs, cudaStream_t st) {
    {
        int rc = 0;
        if (launch_fcc_large2<T>(a, ndirs, st, &rc)) return rc;      // else: the resident-tile variant does not fit; fall through
    }

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
// S <= 256; WIDE: by every thread, S <= 2048) -> fixed-point label bins in shared memory + edge sums; FCC posterior dense over L by all threads;
// grad_x row written once; u*beta scaled in place by gf * scale_t / Z_t for the GEMM.
// WIDE (S > 256, up to GL_THREADS blocks = 2048 positions): the FAC part runs on ALL threads, one 8-position block per thread; the
// three warp-level steps (largest product of the frame, its total, the boundary value from the previous block) become CTA-level ones.
template <typename T, bool FCC, bool FAC, bool WIDE>
__global__ void __launch_bounds__(GL_THREADS) grad_large_frame_kernel(const GradArgs<T> a, unsigned long long *edge /*[2][N][sp]*/) {
    extern __shared__ __align__(16) char gl_smem[];
    typedef FixL<T> FX;
    typedef typename FX::bin_t bin_t;
    const int L = a.d.L, N = a.d.N, Tt = a.d.T;
    const int sp = (int)a.ctx.sp, nblk = (int)a.ctx.nblk, lpi = (int)a.ctx.lpi;
    const int64_t tp = a.ctx.tp;
    bin_t *bins = (bin_t *)gl_smem;                         // [L]
    T *red = (T *)(bins + L);                               // [32]
    int *redi = (int *)(red + 32);                          // [8]     (WIDE)
    T *xv = (T *)(redi + 8);                                // [GL_THREADS] boundary values of the previous frame (WIDE)
    int *xe = (int *)(xv + GL_THREADS);                     // [GL_THREADS] and their exponents
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int blk = WIDE ? tid : lane;                      // this thread's 8-position block
    const bool fac_thread = FAC && (WIDE || warp == 0);
    const int n = blockIdx.y;
    const int t_begin = blockIdx.x * GL_FRAMES;
    const int Tn = (int)a.ctx.il[n];
    const T gf = FCC ? a.gfull(n) : T(0);
    const T ga = FAC ? a.galigned(n) : T(0);

    // FAC constants of this lane's block (warp 0)
    T hreg[8], dreg[8], accH[8], accQ[8];
    int lab[8];
    int Sn = 0;
    if (fac_thread) {
        Sn = (int)a.ctx.tl[n];
        const int64_t *yn = a.y + (int64_t)n * a.y_s0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int s = blk * 8 + j;
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
        // ---- FAC (warp 0, or every thread when WIDE): posterior -> bins, edge sums ----
        if (fac_thread) {
            const bool act = blk < nblk;
            const int b8 = min(blk, nblk - 1) * 8, be = min(blk, nblk - 1);
            T av[8], bv[8];
            const T *pa = a.ctx.fac_alpha(row) + b8, *pb = a.ctx.fac_beta(row) + b8;
#pragma unroll
            for (int j = 0; j < 8; ++j) { av[j] = pa[j]; bv[j] = pb[j]; }
            int ea, eb, da, db;
            const int wa = a.ctx.fac_ea(row)[be], wb = a.ctx.fac_eb(row)[be];
            T fa, fb;
            fac_unpack_norm<T>(wa, ea, da, fa);
            fac_unpack_norm<T>(wb, eb, db, fb);
            T pm = T(0), ps = T(0);
#pragma unroll
            for (int j = 0; j < 8; ++j) { bv[j] = (av[j] * fa) * (bv[j] * fb); pm = tmax(pm, bv[j]); ps += bv[j]; }
            const int c = ea + eb;
            const int ce = (act && pm > T(0)) ? c + Num<T>::pexp(pm) : INT_MIN;
            int cmax = __reduce_max_sync(ASG_FULL_MASK, ce);
            if (WIDE) {
                __syncthreads();
                if (lane == 0) redi[warp] = cmax;
                __syncthreads();
                cmax = redi[0];
#pragma unroll
                for (int w = 1; w < GL_THREADS / 32; ++w) cmax = max(cmax, redi[w]);
            }
            const int sh = Num<T>::block_shift(c, cmax);
            const T zpart = (ce != INT_MIN) ? Num<T>::ldexp_pos(ps, sh) : T(0);
            const T Z = WIDE ? block_sum(zpart, red, GL_THREADS / 32) : warp_sum(zpart);
            const T kq = (Z > Num<T>::tiny() && ce != INT_MIN) ? Num<T>::ldexp_pos(T(1) / Z, sh + Num<T>::kPostLog2) : T(0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                bv[j] = bv[j] * kq;
                if (act && lab[j] >= 0) FX::add(&bins[lab[j]], FX::to_fix(bv[j]));
            }
            if (t >= 1) {
                T pv[8];
                const T *pp = a.ctx.fac_alpha(row - 1) + b8;
#pragma unroll
                for (int j = 0; j < 8; ++j) pv[j] = pp[j];
                int ep, dp_;
                fac_unpack(a.ctx.fac_ea(row - 1)[be], ep, dp_);
                ep += dp_;
                T left;
                int el;
                if (WIDE) {
                    __syncthreads();                          // (the previous frame's exchange has been read)
                    xv[tid] = pv[7]; xe[tid] = ep;
                    __syncthreads();
                    left = tid ? xv[tid - 1] : T(0);
                    el = tid ? xe[tid - 1] : ep;
                } else {
                    left = __shfl_up_sync(ASG_FULL_MASK, pv[7], 1);
                    el = __shfl_up_sync(ASG_FULL_MASK, ep, 1);
                }
                left = (blk == 0) ? T(0) : left * Num<T>::exp2i_flush(min(el - ep, Num<T>::kMaxShift - 8));
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
    if (fac_thread && blk < nblk) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int s = blk * 8 + j;
            if (lab[j] >= 0) {
                const T sc = (sizeof(T) == 4) ? T(4) : T(1.0 / 16777216.0);      // kPost units -> 2^32 (2^38) per unit mass
                atomicAdd(&edge[(int64_t)n * sp + s], (unsigned long long)(long long)llrint((double)(accH[j] * sc)));
                atomicAdd(&edge[((int64_t)N + n) * sp + s], (unsigned long long)(long long)llrint((double)((accQ[j] - accH[j]) * sc)));
            }
        }
    }
}

// FAC part of grad_A: gA[y_s,y_s] += ga_n * gH[n,s], gA[y_s,y_{s-1}] += ga_n * gD[n,s] (force_aligned_lattice.cpp:204-231).
// One thread per (utterance, position); the products go into a 64-bit fixed-point image of the matrix by integer atomics
// (order independent: deterministic), which grad_large_finish_kernel converts back.
template <typename T> struct LargeFix;
template <> struct LargeFix<float> { static constexpr double kEdgeUnfix = 1.0 / 4294967296.0, kOut = 4294967296.0; };             // 2^32
template <> struct LargeFix<double> { static constexpr double kEdgeUnfix = 1.0 / 274877906944.0, kOut = 1099511627776.0; };       // 2^38, 2^40
template <typename T>
__global__ void __launch_bounds__(128) grad_large_scatter_kernel(const GradArgs<T> a, const unsigned long long *edge, unsigned long long *galfix) {
    const int L = a.d.L, N = a.d.N, sp = (int)a.ctx.sp;
    const int n = blockIdx.y, s = blockIdx.x * blockDim.x + threadIdx.x;
    const int Sn = (int)a.ctx.tl[n];
    if (s >= Sn) return;
    const int64_t *yn = a.y + (int64_t)n * a.y_s0;
    const int l = max(0, min(L - 1, (int)yn[(int64_t)s * a.y_s1]));
    const T ga = a.galigned(n);
    const T stay = ga * (T)((double)(long long)edge[(int64_t)n * sp + s] * LargeFix<T>::kEdgeUnfix);
    atomicAdd(&galfix[(size_t)l * L + l], (unsigned long long)__double2ll_rn((double)stay * LargeFix<T>::kOut));
    if (s >= 1) {
        const int prev = max(0, min(L - 1, (int)yn[(int64_t)(s - 1) * a.y_s1]));
        const T mv = ga * (T)((double)(long long)edge[((int64_t)N + n) * sp + s] * LargeFix<T>::kEdgeUnfix);
        atomicAdd(&galfix[(size_t)l * L + prev], (unsigned long long)__double2ll_rn((double)mv * LargeFix<T>::kOut));
    }
}

// grad_A = E o (sum of the K splits of G, fixed order) + the FAC image.  Element-wise, coalesced.
template <typename T, bool FCC, bool FAC>
__global__ void __launch_bounds__(256) grad_large_finish_kernel(const GradArgs<T> a, const T *G, const unsigned long long *galfix, int splitk) {
    const int L = a.d.L, N = a.d.N;
    if (a.losses != nullptr && blockIdx.x == gridDim.x - 1) {      // asg_step(): the extra CTA writes the losses and the fused tail
        __shared__ T red8[8];
        write_losses_block<T>(a.full_scores, a.aligned_scores, a.losses, N, a.tail, red8);
        return;
    }
    const size_t LL = (size_t)L * L;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= LL) return;
    T v = T(0);
    if (FCC) {
        T g = T(0);
        for (int sk = 0; sk < splitk; ++sk) g += G[(size_t)sk * LL + e];          // fixed order
        v = a.ctx.ehat[e] * g;
    }
    if (FAC) v += (T)((double)(long long)galfix[e] * (1.0 / LargeFix<T>::kOut));
    a.grad_A[e] = v;
}

// ---------------------------------------------------------------------------------------------
// The FCC transition gradient of a large alphabet IS a dense contraction over the frame records,
//     G[i][j] = sum_f  w'[f+1][i] * alpha[f][j]        (w' = u o beta scaled by gf * scale_t / Z_t by the frame kernel),
// M = N = L, K = N * tp frames (2 x 1024^2 x 64000 = 134 GFLOP at cfg-4): compute bound, fp32 FMA roofline (no tensor
// cores: the path is specified in fp32 and a TF32 product would need a 3-way split).  Hand-written here: 128 x 128 tile
// per CTA, 8 x 8 register tile per thread as packed FFMA2, both operands streamed straight from the frame records
// (K-major: a frame's row is contiguous) through a 3-stage cp.async ring, split over K so that every SM has a tile;
// the splits are summed in fixed order by grad_large_finish_kernel (deterministic).
// ---------------------------------------------------------------------------------------------
constexpr int GM_T = 128;          // tile edge (rows and columns of G)
constexpr int GM_K = 16;           // frames per stage
constexpr int GM_ST = 3;           // stages
constexpr int GM_THREADS = 256;

__host__ __device__ inline int large_splitk(int L) {
    const int tiles = ((L + GM_T - 1) / GM_T) * ((L + GM_T - 1) / GM_T);
    int s = 148 / tiles;
    return s < 1 ? 1 : (s > 8 ? 8 : s);
}

template <typename T>
__global__ void __launch_bounds__(GM_THREADS) grad_large_gemm_kernel(const char *rec, int64_t R, int o_w, int o_a, int lpi, int L,
                                                                     int64_t F, int splitk, T *Gpart) {
    extern __shared__ __align__(16) char gm_smem[];
    constexpr int VE = 16 / (int)sizeof(T);                    // elements per 16-byte vector
    constexpr int VPR = GM_T / VE;                             // vectors per tile row
    T *As = reinterpret_cast<T *>(gm_smem);                    // [GM_ST][GM_K][GM_T]  w' rows (i)
    T *Bs = As + GM_ST * GM_K * GM_T;                          // [GM_ST][GM_K][GM_T]  alpha rows (j)
    const int tid = threadIdx.x;
    const int i0 = blockIdx.x * GM_T, j0 = blockIdx.y * GM_T;
    const int64_t f_begin = F * blockIdx.z / splitk, f_end = F * (blockIdx.z + 1) / splitk;
    const int nchunks = (int)((f_end - f_begin + GM_K - 1) / GM_K);

    auto issue = [&](int c) {
        if (c < nchunks) {
            const int st = c % GM_ST;
            for (int v = tid; v < 2 * GM_K * VPR; v += GM_THREADS) {
                const int which = v / (GM_K * VPR), w = v - which * GM_K * VPR;
                const int kk = w / VPR, col = (w - kk * VPR) * VE;
                const int64_t f = f_begin + (int64_t)c * GM_K + kk;
                const int e0 = (which ? j0 : i0) + col;        // first element of this vector inside the frame's row
                const char *src = rec + (f + (which ? 0 : 1)) * R + (which ? o_a : o_w) + (int64_t)e0 * sizeof(T);
                T *dst = (which ? Bs : As) + ((size_t)st * GM_K + kk) * GM_T + col;
                // bytes of this vector that exist: rows hold lpi (>= L, multiple of 8) elements; frames past the split's end read as 0
                int bytes = (f < f_end) ? (lpi - e0) * (int)sizeof(T) : 0;
                bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(dst)), "l"(bytes ? src : rec), "r"(bytes) : "memory");
            }
        }
        cp_async_commit();
    };
    // register tile: rows  ty*4 + {0..3}  and  64 + ty*4 + {0..3},  columns likewise with tx (conflict-free 16-byte loads)
    const int ty = tid >> 4, tx = tid & 15;
    T acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = T(0);

    for (int c = 0; c < GM_ST - 1; ++c) issue(c);
    for (int c = 0; c < nchunks; ++c) {
        cp_async_wait<GM_ST - 2>();
        __syncthreads();
        issue(c + GM_ST - 1);                                  // refills the stage consumed in the previous iteration
        const T *At = As + (size_t)(c % GM_ST) * GM_K * GM_T, *Bt = Bs + (size_t)(c % GM_ST) * GM_K * GM_T;
#pragma unroll
        for (int kk = 0; kk < GM_K; ++kk) {
            T av[8], bv[8];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    av[4 * h + q] = At[kk * GM_T + 64 * h + ty * 4 + q];
                    bv[4 * h + q] = Bt[kk * GM_T + 64 * h + tx * 4 + q];
                }
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) acc[r][cc] = fma(av[r], bv[cc], acc[r][cc]);
        }
    }
    cp_async_wait<0>();
    T *out = Gpart + (size_t)blockIdx.z * L * L;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int i = i0 + 64 * (r >> 2) + ty * 4 + (r & 3);
        if (i >= L) continue;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
            const int j = j0 + 64 * (cc >> 2) + tx * 4 + (cc & 3);
            if (j < L) out[(size_t)i * L + j] = acc[r][cc];
        }
    }
}

template <typename T, bool FCC, bool FAC>
static int launch_grad_large_t(const GradArgs<T> &a, cudaStream_t st) {
    const int L = a.d.L, N = a.d.N;
    int device = 0;
    cudaGetDevice(&device);
    // scratch: splitk x G [L,L] | edge sums [2,N,sp] int64 | FAC image of grad_A [L,L] int64
    T *G = a.partials;
    unsigned long long *edge = (unsigned long long *)((char *)a.partials + align_up(sizeof(T) * (size_t)large_splitk(L) * L * L, 256));
    unsigned long long *galfix = (unsigned long long *)((char *)edge + align_up(16 * (size_t)N * a.ctx.sp, 256));
    cudaError_t e = cudaMemsetAsync(edge, 0, align_up(16 * (size_t)N * a.ctx.sp, 256) + (FAC ? sizeof(unsigned long long) * (size_t)L * L : 0), st);
    if (e != cudaSuccess) return (int)e;
    const size_t smem = (size_t)L * 8 + 32 * sizeof(T) + 8 * sizeof(int) + GL_THREADS * (sizeof(T) + sizeof(int)) + 64;
    const bool wide = FAC && a.ctx.nblk > 32;
    auto kern = wide ? grad_large_frame_kernel<T, FCC, FAC, true> : grad_large_frame_kernel<T, FCC, FAC, false>;
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
    dim3 grid((unsigned)((a.ctx.tp + GL_FRAMES - 1) / GL_FRAMES), N);
    {
        LaunchScope ls(K_GRAD, st);
        kern<<<grid, GL_THREADS, smem, st>>>(a, edge);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
    const int splitk = large_splitk(L);
    if (FCC) {
        // G[i][j] = sum_f w'[f][i] * alpha[f-1][j] over all frame records f >= 1 (w' is zero at t = 0 and on padding)
        const int64_t F = (int64_t)N * a.ctx.tp - 1;
        const size_t gsm = (size_t)2 * GM_ST * GM_K * GM_T * sizeof(T);
        if ((e = cudaFuncSetAttribute(grad_large_gemm_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm)) != cudaSuccess) return (int)e;
        dim3 ggrid((L + GM_T - 1) / GM_T, (L + GM_T - 1) / GM_T, splitk);
        {
            LaunchScope ls(K_GRAD_GEMM, st);
            grad_large_gemm_kernel<T><<<ggrid, GM_THREADS, gsm, st>>>(a.ctx.rec, a.ctx.R, a.ctx.o_cw, a.ctx.o_ca, (int)a.ctx.lpi, L, F, splitk, G);
        }
        if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
    }
    if (FAC) {
        LaunchScope ls(K_GRAD_FINALIZE, st);
        dim3 sgrid((unsigned)((a.d.S + 127) / 128), N);
        grad_large_scatter_kernel<T><<<sgrid, 128, 0, st>>>(a, edge, galfix);
    }
    {
        LaunchScope ls(K_GRAD_FINALIZE, st);
        const unsigned nb = (unsigned)(((size_t)L * L + 255) / 256);
        grad_large_finish_kernel<T, FCC, FAC><<<nb + (a.losses ? 1 : 0), 256, 0, st>>>(a, G, galfix, splitk);
    }
    return (int)cudaGetLastError();
}

int large_splitk_host(int L) { return large_splitk(L); }

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
