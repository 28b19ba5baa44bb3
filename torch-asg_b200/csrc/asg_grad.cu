// asg_grad.cu -- the "no-recursion" gradient assembly as ONE bandwidth-bound persistent kernel.
//
// Replaces (reference, /root/reference/torch_asg/native):
//   fully_connected_derivative / _backward     fully_connected_lattice.cpp:50-63, :94-104
//   force_aligned_derivative / _backward       force_aligned_lattice.cpp:156-183, :321-356
//   collect_input_grad_{cpu,gpu}               force_aligned_lattice.cpp:234-264, _kernel.cu:370-470
//   collect_transition_grad_{cpu,gpu}          force_aligned_lattice.cpp:204-231, _kernel.cu:255-344
//   fast_asg_gpu_backward                      streamlined_fast_gpu.cpp:236-297
//
// What it streams (per frame t, utterance n): the scaled FCC states alpha[L], beta[L], u*beta[L], the FAC
// block-floating-point states alpha[sp], beta[sp] with their block exponents, and writes grad_x[L] exactly
// once.  Nothing like the reference's path_contrib[T-1,N,L,L] exists: the edge posteriors are rebuilt from
// alpha[t-1] and u*beta[t] on the fly and contracted into a per-CTA L x L register tile
// (G += w_t (x) alpha_{t-1}); the FAC label scatter is a per-utterance CSR gather in shared memory instead of
// T*N*S global atomics; every reduction has a fixed order, so results are run-to-run deterministic.
//
// Work decomposition: tile = (utterance n, TB consecutive frames); tiles are dealt in n-major order to
// a persistent grid so a CTA mostly stays on one utterance (its CSR / transition rows stay in smem).
#include <climits>
#include <cstdlib>
#include "asg_common.cuh"
#include "asg_internal.h"

namespace asg {

constexpr int GRAD_THREADS = 256;
constexpr int GRAD_WARPS = GRAD_THREADS / 32;

template <typename T> struct GradSmem {
    // FAC
    T *sA;      // [(TB+1)][spp]   alpha mantissas, frames t0-1 .. t0+TB-1
    T *sP;      // [TB][spp]       unnormalised posteriors alpha*beta*2^(C-Cmax)
    int *sEa;   // [(TB+1)][nblk]
    int *sC;    // [TB][nblk]      Ea+Eb of live blocks
    T *sZb;     // [TB]            1/Zbar * g_aligned (0 for an empty frame)
    T *hh;      // [sp]
    T *din;     // [sp]            edge (s-1 -> s)
    T *tH;      // [sp]            running sum_t of the stay-edge posteriors
    T *tD;      // [sp]            running sum_t of the move-edge posteriors (edge s-1 -> s)
    int *ylab;  // [sp]
    int *order; // [sp]            positions sorted by label (CSR payload)
    int *lstart;// [L+1]
    T *Gal;     // [L][L]          FAC transition gradient of this CTA
    // FCC
    T *sAl;     // [(TB+1)][LP]
    T *sW;      // [TB][LP]
    // shared
    T *sG;      // [TB][L+1]       grad_x rows of the tile
};

template <typename T>
__host__ __device__ inline size_t grad_smem_carve(GradSmem<T> *v, char *base, int L, int LP, int sp, int nblk, int TB,
                                                  bool fcc, bool fac) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 16); return o; };
    const int spp = sp | 1;
    size_t o_sA = take(fac ? sizeof(T) * (size_t)(TB + 1) * spp : 0);
    size_t o_sP = take(fac ? sizeof(T) * (size_t)TB * spp : 0);
    size_t o_sEa = take(fac ? sizeof(int) * (size_t)(TB + 1) * nblk : 0);
    size_t o_sC = take(fac ? sizeof(int) * (size_t)TB * nblk : 0);
    size_t o_sZb = take(fac ? sizeof(T) * TB : 0);
    size_t o_hh = take(fac ? sizeof(T) * sp : 0);
    size_t o_din = take(fac ? sizeof(T) * sp : 0);
    size_t o_tH = take(fac ? sizeof(T) * sp : 0);
    size_t o_tD = take(fac ? sizeof(T) * sp : 0);
    size_t o_ylab = take(fac ? sizeof(int) * sp : 0);
    size_t o_order = take(fac ? sizeof(int) * sp : 0);
    size_t o_lstart = take(fac ? sizeof(int) * (L + 1) : 0);
    size_t o_Gal = take(fac ? sizeof(T) * (size_t)L * L : 0);
    size_t o_sAl = take(fcc ? sizeof(T) * (size_t)(TB + 1) * LP : 0);
    size_t o_sW = take(fcc ? sizeof(T) * (size_t)TB * LP : 0);
    size_t o_sG = take(sizeof(T) * (size_t)TB * (L + 1));
    if (v) {
        v->sA = (T *)(base + o_sA); v->sP = (T *)(base + o_sP); v->sEa = (int *)(base + o_sEa); v->sC = (int *)(base + o_sC);
        v->sZb = (T *)(base + o_sZb); v->hh = (T *)(base + o_hh); v->din = (T *)(base + o_din);
        v->tH = (T *)(base + o_tH); v->tD = (T *)(base + o_tD); v->ylab = (int *)(base + o_ylab);
        v->order = (int *)(base + o_order); v->lstart = (int *)(base + o_lstart); v->Gal = (T *)(base + o_Gal);
        v->sAl = (T *)(base + o_sAl); v->sW = (T *)(base + o_sW); v->sG = (T *)(base + o_sG);
    }
    return off;
}

// 8 consecutive elements (32-byte aligned for float) through the streaming read path
__device__ __forceinline__ void load8(const float *p, float (&v)[8]) {
    const float4 a = __ldcs(reinterpret_cast<const float4 *>(p));
    const float4 b = __ldcs(reinterpret_cast<const float4 *>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const double *p, double (&v)[8]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const double2 a = __ldcs(reinterpret_cast<const double2 *>(p) + q);
        v[2 * q] = a.x; v[2 * q + 1] = a.y;
    }
}

template <typename T, int LP, bool FCC, bool FAC>
__global__ void __launch_bounds__(GRAD_THREADS) grad_kernel(const GradArgs<T> a) {
    extern __shared__ __align__(16) char smem_raw[];
    const int L = a.d.L, N = a.d.N, Tt = a.d.T;
    const int sp = (int)a.ctx.sp, nblk = (int)a.ctx.nblk, nbp = (int)a.ctx.nbp, lpi = (int)a.ctx.lpi, spp = sp | 1;
    const int64_t tp = a.ctx.tp;
    const int TB = a.tb;
    GradSmem<T> sm;
    grad_smem_carve<T>(&sm, smem_raw, L, LP, sp, nblk, TB, FCC, FAC);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- FCC register tile: thread owns rows i0..i0+1, columns j0..j0+3 of G (looped if LP*LP/8 > 256) ----
    constexpr int TCOLS = LP / 4;                 // tiles per row of G
    constexpr int NT = (LP / 2) * TCOLS;          // number of 2x4 tiles
    constexpr int TPT = (NT + GRAD_THREADS - 1) / GRAD_THREADS;
    T G[TPT][8];
#pragma unroll
    for (int q = 0; q < TPT; ++q)
#pragma unroll
        for (int e = 0; e < 8; ++e) G[q][e] = T(0);

    if (FAC) {
        for (int e = tid; e < L * L; e += GRAD_THREADS) sm.Gal[e] = T(0);
    }

    const int per = (a.num_tiles + gridDim.x - 1) / gridDim.x;
    const int q_begin = blockIdx.x * per;
    const int q_end = min(a.num_tiles, q_begin + per);
    int cur_n = -1;
    int Tn = 0, Sn = 0;
    T gf = T(0), ga = T(0);

    // scatter the running per-position edge sums of utterance cur_n into the CTA's L x L matrix.
    // Thread l owns row l of Gal (destination label), so the order of additions is fixed.
    auto flush_fac = [&]() {
        if (!FAC || cur_n < 0) return;
        __syncthreads();
        if (tid < L) {
            for (int k = sm.lstart[tid]; k < sm.lstart[tid + 1]; ++k) {
                const int s = sm.order[k];
                sm.Gal[tid * L + tid] += sm.tH[s];
                if (s >= 1) sm.Gal[tid * L + sm.ylab[s - 1]] += sm.tD[s];
            }
        }
        __syncthreads();
    };

    for (int q = q_begin; q < q_end; ++q) {
        const int n = q / a.tiles_t;
        const int t0 = (q % a.tiles_t) * TB;
        const int TBt = min(TB, Tt - t0);         // frames in this tile
        if (n != cur_n) {
            flush_fac();
            cur_n = n;
            Tn = (int)a.ctx.il[n];
            gf = FCC ? a.g_full[(int64_t)n * a.g_full_stride] : T(0);
            ga = FAC ? a.g_aligned[(int64_t)n * a.g_aligned_stride] : T(0);
            if (FAC) {
                Sn = (int)a.ctx.tl[n];
                const int64_t *yn = a.y + (int64_t)n * a.y_s0;
                for (int s = tid; s < sp; s += GRAD_THREADS) {
                    int lab = -1;
                    if (s < Sn) { lab = (int)yn[(int64_t)s * a.y_s1]; lab = max(0, min(L - 1, lab)); }
                    sm.ylab[s] = lab;
                    sm.hh[s] = a.ctx.fac_hd[(int64_t)n * sp + s];
                    sm.din[s] = (s >= 1) ? a.ctx.fac_hd[((int64_t)N + n) * sp + s - 1] : T(0);
                    sm.tH[s] = T(0);
                    sm.tD[s] = T(0);
                }
                __syncthreads();
                // CSR by label; positions ascending inside a label (fixed summation order).
                // Thread l counts, thread 0 scans the L counts, thread l fills its own segment.
                int cnt = 0;
                if (tid < L)
                    for (int s = 0; s < Sn; ++s) cnt += (sm.ylab[s] == tid);
                if (tid < L) sm.lstart[tid + 1] = cnt;
                __syncthreads();
                if (tid == 0) {
                    int acc = 0;
                    sm.lstart[0] = 0;
                    for (int l = 0; l < L; ++l) { acc += sm.lstart[l + 1]; sm.lstart[l + 1] = acc; }
                }
                __syncthreads();
                if (tid < L) {
                    int pos = sm.lstart[tid];
                    for (int s = 0; s < Sn; ++s) if (sm.ylab[s] == tid) sm.order[pos++] = s;
                }
            }
            __syncthreads();
        }
        const int rows_valid = max(0, min(TBt, Tn - t0));   // frames t0 .. t0+rows_valid-1 are inside the utterance

        // ================= phase 1: stream the tile in, form posteriors ==========================
        for (int r = warp; r <= TBt; r += GRAD_WARPS) {
            const int t = t0 - 1 + r;
            const bool have_prev = (t >= 0) && (t < Tn);          // alpha row t exists
            const bool cur = (r >= 1) && (t < Tn);                // frame t is a live frame of this tile
            const int64_t row = (int64_t)n * tp + t;
            if (FAC) {
                T zsum = T(0);
                int cmax = INT_MIN;
                // pass 1: load, products, block liveness, exponent max (one 8-position block per lane;
                // sp > 256 loops).  Raw products and the block exponent sums are parked in smem.
                for (int b0 = 0; b0 < nblk; b0 += 32) {
                    const int b = b0 + lane;
                    if (b >= nblk) continue;
                    T av[8], bv[8];
                    int ea = 0, eb = 0;
                    if (have_prev) {
                        load8(a.fac_alpha + row * sp + (int64_t)b * 8, av);
                        ea = __ldcs(a.ctx.fac_ea + row * nbp + b);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) av[j] = T(0);
                    }
                    if (cur) {
                        load8(a.fac_beta + row * sp + (int64_t)b * 8, bv);
                        eb = __ldcs(a.ctx.fac_eb + row * nbp + b);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) sm.sA[r * spp + b * 8 + j] = av[j];
                    sm.sEa[r * nblk + b] = ea;
                    if (r >= 1) {
                        T pm = T(0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            bv[j] = cur ? bv[j] * av[j] : T(0);
                            pm = tmax(pm, bv[j]);
                            sm.sP[(r - 1) * spp + b * 8 + j] = bv[j];
                        }
                        const int c = (pm > T(0)) ? ea + eb : INT_MIN;   // a block without mass is ignored
                        sm.sC[(r - 1) * nblk + b] = c;
                        cmax = max(cmax, c);
                    }
                }
                if (r >= 1) {
                    cmax = __reduce_max_sync(ASG_FULL_MASK, cmax);
                    __syncwarp();
                    // pass 2: bring every block to the frame's common exponent, sum the frame
                    for (int b0 = 0; b0 < nblk; b0 += 32) {
                        const int b = b0 + lane;
                        if (b >= nblk) continue;
                        const int c = sm.sC[(r - 1) * nblk + b];
                        T wgt = T(0);
                        if (c != INT_MIN && c - cmax >= -Num<T>::kMaxShift) wgt = Num<T>::exp2i(c - cmax);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const T pv = sm.sP[(r - 1) * spp + b * 8 + j] * wgt;
                            sm.sP[(r - 1) * spp + b * 8 + j] = pv;
                            zsum += pv;
                        }
                    }
                    zsum = warp_sum(zsum);
                    if (lane == 0) sm.sZb[r - 1] = (cur && zsum > T(0)) ? ga / zsum : T(0);
                }
            }
            if (FCC) {
                T al0 = T(0), al1 = T(0);
                if (have_prev) {
                    if (lane < L) al0 = __ldcs(a.fcc_alpha + row * lpi + lane);
                    if (LP > 32 && lane + 32 < L) al1 = __ldcs(a.fcc_alpha + row * lpi + lane + 32);
                }
                if (lane < LP) sm.sAl[r * LP + lane] = al0;
                if (LP > 32 && lane + 32 < LP) sm.sAl[r * LP + lane + 32] = al1;
                if (r >= 1) {
                    T be0 = T(0), be1 = T(0), w0 = T(0), w1 = T(0), sc = T(0);
                    if (cur) {
                        if (lane < L) { be0 = __ldcs(a.fcc_beta + row * lpi + lane); w0 = __ldcs(a.ctx.fcc_w + row * lpi + lane); }
                        if (LP > 32 && lane + 32 < L) {
                            be1 = __ldcs(a.fcc_beta + row * lpi + lane + 32);
                            w1 = __ldcs(a.ctx.fcc_w + row * lpi + lane + 32);
                        }
                        sc = __ldcs(a.ctx.fcc_scale + row);
                    }
                    const T g0 = al0 * be0, g1 = al1 * be1;
                    const T Z = warp_sum(g0 + g1);
                    const T inv = (cur && Z > T(0)) ? gf / Z : T(0);
                    if (lane < L) sm.sG[(r - 1) * (L + 1) + lane] = g0 * inv;
                    if (LP > 32 && lane + 32 < L) sm.sG[(r - 1) * (L + 1) + lane + 32] = g1 * inv;
                    const T winv = (t >= 1) ? inv * sc : T(0);          // no incoming edge at frame 0
                    if (lane < LP) sm.sW[(r - 1) * LP + lane] = w0 * winv;
                    if (LP > 32 && lane + 32 < LP) sm.sW[(r - 1) * LP + lane + 32] = w1 * winv;
                }
            } else if (r >= 1) {
                for (int l = lane; l < L; l += 32) sm.sG[(r - 1) * (L + 1) + l] = T(0);
            }
        }
        __syncthreads();

        // ================= phase 2: label gather, edge sums, outer products =======================
        if (FAC && rows_valid > 0) {
            // (a) grad_x[t, y_s] += post[t, s]: lanes = frames, one label per (warp, sub-group)
            const int tl_ = lane % TB, sub = lane / TB, nsub = (TB >= 32) ? 1 : 32 / TB;
            if (tl_ < rows_valid && sub < nsub) {
                for (int l = warp * nsub + sub; l < L; l += GRAD_WARPS * nsub) {
                    T acc = T(0);
                    for (int k = sm.lstart[l]; k < sm.lstart[l + 1]; ++k) acc += sm.sP[tl_ * spp + sm.order[k]];
                    sm.sG[tl_ * (L + 1) + l] += acc * sm.sZb[tl_];
                }
            }
            // (b) per-position edge posteriors, summed over the tile's frames
            for (int s = tid; s < Sn; s += GRAD_THREADS) {
                T accH = T(0), accD = T(0);
                const T hs = sm.hh[s], ds = sm.din[s];
                const bool edge = (s & 7) == 0 && s > 0;
                for (int r = 1; r <= rows_valid; ++r) {
                    if (t0 - 1 + r < 1) continue;                       // frame 0 has no incoming edge
                    const T post = sm.sP[(r - 1) * spp + s] * sm.sZb[r - 1];
                    const T stay = hs * sm.sA[(r - 1) * spp + s];
                    T lft = (s > 0) ? sm.sA[(r - 1) * spp + s - 1] : T(0);
                    if (edge) lft *= Num<T>::exp2i(sm.sEa[(r - 1) * nblk + (s >> 3) - 1] - sm.sEa[(r - 1) * nblk + (s >> 3)]);
                    const T move = ds * lft;
                    const T den = stay + move;
                    if (den > T(0)) {
                        const T pd = post / den;
                        accH = fma(pd, stay, accH);
                        accD = fma(pd, move, accD);
                    }
                }
                sm.tH[s] += accH;
                sm.tD[s] += accD;
            }
        }
        if (FCC && rows_valid > 0) {
#pragma unroll
            for (int qq = 0; qq < TPT; ++qq) {
                const int tile = tid + qq * GRAD_THREADS;
                if (tile < NT) {
                    const int i0 = (tile / TCOLS) * 2, j0 = (tile % TCOLS) * 4;
                    for (int r = 1; r <= rows_valid; ++r) {
                        const T wa = sm.sW[(r - 1) * LP + i0], wb = sm.sW[(r - 1) * LP + i0 + 1];
                        const T *pp = sm.sAl + (r - 1) * LP + j0;
                        const T p0 = pp[0], p1 = pp[1], p2 = pp[2], p3 = pp[3];
                        G[qq][0] = fma(wa, p0, G[qq][0]); G[qq][1] = fma(wa, p1, G[qq][1]);
                        G[qq][2] = fma(wa, p2, G[qq][2]); G[qq][3] = fma(wa, p3, G[qq][3]);
                        G[qq][4] = fma(wb, p0, G[qq][4]); G[qq][5] = fma(wb, p1, G[qq][5]);
                        G[qq][6] = fma(wb, p2, G[qq][6]); G[qq][7] = fma(wb, p3, G[qq][7]);
                    }
                }
            }
        }
        __syncthreads();

        // ================= phase 3: write grad_x rows (each element exactly once) ================
        for (int e = tid; e < TBt * L; e += GRAD_THREADS) {
            const int r = e / L, l = e % L;
            a.grad_x[((int64_t)(t0 + r) * N + n) * L + l] = sm.sG[r * (L + 1) + l];
        }
        __syncthreads();
    }
    flush_fac();
    __syncthreads();

    // ---- per-CTA partial of grad_A = E o G (FCC) + Gal (FAC) -------------------------------------
    T *part = a.partials + (size_t)blockIdx.x * L * L;
#pragma unroll
    for (int qq = 0; qq < TPT; ++qq) {
        const int tile = tid + qq * GRAD_THREADS;
        if (tile < NT) {
            const int i0 = (tile / TCOLS) * 2, j0 = (tile % TCOLS) * 4;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int i = i0 + (e >> 2), j = j0 + (e & 3);
                if (i < L && j < L) {
                    T v = T(0);
                    if (FCC) v = a.ctx.ehat[i * L + j] * G[qq][e];
                    if (FAC) v += sm.Gal[i * L + j];
                    part[i * L + j] = v;
                }
            }
        }
    }
}

// grad_A[e] = sum over CTAs of partial[c][e], fixed order
template <typename T>
__global__ void grad_finalize_kernel(const T *partials, int nparts, int LL, T *grad_A) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= LL) return;
    T acc0 = T(0), acc1 = T(0), acc2 = T(0), acc3 = T(0);
    int c = 0;
    for (; c + 3 < nparts; c += 4) {
        acc0 += partials[(size_t)c * LL + e];
        acc1 += partials[(size_t)(c + 1) * LL + e];
        acc2 += partials[(size_t)(c + 2) * LL + e];
        acc3 += partials[(size_t)(c + 3) * LL + e];
    }
    for (; c < nparts; ++c) acc0 += partials[(size_t)c * LL + e];
    grad_A[e] = (acc0 + acc1) + (acc2 + acc3);
}

// =============================================================================================
int grad_grid_size(int device) {
    static int sms[64] = {0};
    if (device < 0 || device >= 64) return 148 * 2;
    if (!sms[device]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
        sms[device] = v;
    }
    return sms[device] * 2;
}

template <typename T, int LP, bool FCC, bool FAC>
static cudaError_t launch_grad_lp(const GradArgs<T> &a_in, cudaStream_t st) {
    GradArgs<T> a = a_in;
    const int L = a.d.L, sp = (int)a.ctx.sp, nblk = (int)a.ctx.nblk;
    int device = 0;
    cudaGetDevice(&device);
    int tb = 16;
    if (const char *env = getenv("ASG_GRAD_TB")) { int v = atoi(env); if (v == 8 || v == 16 || v == 32) tb = v; }
    size_t bytes = grad_smem_carve<T>(nullptr, nullptr, L, LP, sp, nblk, tb, FCC, FAC);
    while (bytes > 100 * 1024 && tb > 1) { tb >>= 1; bytes = grad_smem_carve<T>(nullptr, nullptr, L, LP, sp, nblk, tb, FCC, FAC); }
    if (bytes > 227 * 1024) return cudaErrorInvalidValue;
    a.tb = tb;
    a.tiles_t = (a.d.T + tb - 1) / tb;
    a.num_tiles = a.tiles_t * a.d.N;
    cudaError_t e = cudaFuncSetAttribute(grad_kernel<T, LP, FCC, FAC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    int grid = grad_grid_size(device);
    if (grid > a.num_tiles) grid = a.num_tiles;
    {
        LaunchScope ls(K_GRAD, st);
        grad_kernel<T, LP, FCC, FAC><<<grid, GRAD_THREADS, bytes, st>>>(a);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    const int LL = L * L;
    {
        LaunchScope ls(K_GRAD_FINALIZE, st);
        grad_finalize_kernel<T><<<(LL + 255) / 256, 256, 0, st>>>(a.partials, grid, LL, a.grad_A);
    }
    return cudaGetLastError();
}

template <typename T, int LP>
static cudaError_t launch_grad_parts(const GradArgs<T> &a, int parts, cudaStream_t st) {
    if (parts == ASG_PART_BOTH) return launch_grad_lp<T, LP, true, true>(a, st);
    if (parts == ASG_PART_FCC) return launch_grad_lp<T, LP, true, false>(a, st);
    return launch_grad_lp<T, LP, false, true>(a, st);
}

template <typename T> int launch_grad(const GradArgs<T> &a, int parts, cudaStream_t st) {
    const int L = a.d.L;
    cudaError_t e;
    if (L <= 8) e = launch_grad_parts<T, 8>(a, parts, st);
    else if (L <= 16) e = launch_grad_parts<T, 16>(a, parts, st);
    else if (L <= 24) e = launch_grad_parts<T, 24>(a, parts, st);
    else if (L <= 32) e = launch_grad_parts<T, 32>(a, parts, st);
    else if (L <= 40) e = launch_grad_parts<T, 40>(a, parts, st);
    else if (L <= 48) e = launch_grad_parts<T, 48>(a, parts, st);
    else if (L <= 56) e = launch_grad_parts<T, 56>(a, parts, st);
    else if (L <= 64) e = launch_grad_parts<T, 64>(a, parts, st);
    else e = cudaErrorInvalidValue;
    return (int)e;
}

template int launch_grad<float>(const GradArgs<float> &, int, cudaStream_t);
template int launch_grad<double>(const GradArgs<double> &, int, cudaStream_t);

}  // namespace asg
