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
// Structure: tile = (utterance n, TB consecutive frames).  State is utterance-major, so a tile of every state
// array is ONE contiguous chunk: a producer warp fetches it with ~7 bulk (TMA) copies (cp.async.bulk ->
// UBLKCP) into a two-stage shared-memory ring guarded by mbarriers, while eight consumer warps work on the
// previous tile.  Tiles are dealt in n-major order to a persistent grid so a CTA mostly stays on one utterance
// (its CSR / transition rows stay in shared memory).
#include <climits>
#include <cstdlib>
#include "asg_common.cuh"
#include "asg_internal.h"

namespace asg {

constexpr int GRAD_CWARPS = 8;                        // consumer warps
constexpr int GRAD_CONSUMERS = GRAD_CWARPS * 32;
constexpr int GRAD_THREADS = GRAD_CONSUMERS + 32;     // + one producer warp

// shared-memory map: [per-CTA block][stage 0][stage 1]...
struct GradMap {
    // per stage, byte offsets from the stage base
    size_t oA, oB, oEa, oEb, oBnd, oZb, oFa, oFb, oFw, stage_bytes;
    // per CTA, byte offsets from the base
    size_t o_hh, o_din, o_tH, o_tD, o_ylab, o_order, o_lstart, o_Gal, o_bars, o_stages, total;
};

__host__ __device__ inline GradMap grad_map(size_t eb, int L, int lpi, int sp, int nbp, int TB, int nstage, bool fcc, bool fac) {
    GradMap m;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 16); return o; };
    m.oA = take(fac ? eb * (size_t)(TB + 1) * sp : 0);
    m.oB = take(fac ? eb * (size_t)TB * sp : 0);
    m.oEa = take(fac ? sizeof(int) * (size_t)(TB + 1) * nbp : 0);
    m.oEb = take(fac ? sizeof(int) * (size_t)TB * nbp : 0);
    m.oBnd = take(fac ? eb * (size_t)(TB + 1) * nbp : 0);
    m.oZb = take(fac ? eb * (size_t)TB : 0);
    m.oFa = take(fcc ? eb * (size_t)(TB + 1) * lpi : 0);
    m.oFb = take(fcc ? eb * (size_t)TB * lpi : 0);
    m.oFw = take(fcc ? eb * (size_t)TB * lpi : 0);
    m.stage_bytes = align_up(off, 128);
    off = 0;
    m.o_hh = take(fac ? eb * sp : 0);
    m.o_din = take(fac ? eb * sp : 0);
    m.o_tH = take(fac ? eb * sp : 0);
    m.o_tD = take(fac ? eb * sp : 0);
    m.o_ylab = take(fac ? sizeof(int) * sp : 0);
    m.o_order = take(fac ? sizeof(int) * sp : 0);
    m.o_lstart = take(fac ? sizeof(int) * (L + 1) : 0);
    m.o_Gal = take(fac ? eb * (size_t)L * L : 0);
    m.o_bars = take(sizeof(uint64_t) * 2 * 4);
    m.o_stages = align_up(off, 128);
    m.total = m.o_stages + m.stage_bytes * nstage;
    return m;
}

// 8 consecutive elements (32-byte aligned for float) from / to shared memory
__device__ __forceinline__ void lds8(const float *p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4 *>(p);
    const float4 b = *(reinterpret_cast<const float4 *>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void lds8(const double *p, double (&v)[8]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const double2 a = *(reinterpret_cast<const double2 *>(p) + q);
        v[2 * q] = a.x; v[2 * q + 1] = a.y;
    }
}
__device__ __forceinline__ void sts8(float *p, const float (&v)[8]) {
    *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *(reinterpret_cast<float4 *>(p) + 1) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void sts8(double *p, const double (&v)[8]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) *(reinterpret_cast<double2 *>(p) + q) = make_double2(v[2 * q], v[2 * q + 1]);
}

template <typename T, int LP, bool FCC, bool FAC>
__global__ void __launch_bounds__(GRAD_THREADS) grad_kernel(const GradArgs<T> a) {
    extern __shared__ __align__(128) char smem_raw[];
    const int L = a.d.L, N = a.d.N, Tt = a.d.T;
    const int sp = (int)a.ctx.sp, nblk = (int)a.ctx.nblk, nbp = (int)a.ctx.nbp, lpi = (int)a.ctx.lpi;
    const int64_t tp = a.ctx.tp;
    const int TB = a.tb, NS = a.nstage;
    const GradMap mp = grad_map(sizeof(T), L, lpi, sp, nbp, TB, NS, FCC, FAC);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    T *hh = (T *)(smem_raw + mp.o_hh), *din = (T *)(smem_raw + mp.o_din);
    T *tH = (T *)(smem_raw + mp.o_tH), *tD = (T *)(smem_raw + mp.o_tD), *Gal = (T *)(smem_raw + mp.o_Gal);
    int *ylab = (int *)(smem_raw + mp.o_ylab), *order = (int *)(smem_raw + mp.o_order), *lstart = (int *)(smem_raw + mp.o_lstart);
    uint64_t *full = (uint64_t *)(smem_raw + mp.o_bars), *empty = full + 4;
    char *stages = smem_raw + mp.o_stages;

    const int per = (a.num_tiles + gridDim.x - 1) / gridDim.x;
    const int q_begin = blockIdx.x * per;
    const int q_end = min(a.num_tiles, q_begin + per);

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], GRAD_CWARPS); }
        mbar_fence_init();
    }
    __syncthreads();

    // =========================== producer warp: bulk copies into the ring ===========================
    if (warp == GRAD_CWARPS) {
        if (lane == 0) {
            int cur_n = -1, Tn = 0;
            int it = 0;
            for (int q = q_begin; q < q_end; ++q, ++it) {
                const int stage = it % NS, use = it / NS;
                const int n = q / a.tiles_t;
                const int t0 = (q % a.tiles_t) * TB;
                const int TBt = min(TB, Tt - t0);
                if (n != cur_n) { cur_n = n; Tn = (int)a.ctx.il[n]; }
                const int rows = max(0, min(TBt, Tn - t0));
                mbar_wait(&empty[stage], (use & 1) ^ 1);
                if (rows == 0) { mbar_arrive(&full[stage]); continue; }
                char *sb = stages + (size_t)stage * mp.stage_bytes;
                const int prev = t0 > 0 ? 1 : 0;                   // the halo row t0-1 exists
                const int64_t rowc = (int64_t)n * tp + t0;           // first current row
                const int64_t rowa = rowc - prev;                    // first alpha row fetched
                const unsigned na = (unsigned)(rows + prev);
                unsigned bytes = 0;
                if (FAC) bytes += (unsigned)(sizeof(T) * (na + rows) * sp + sizeof(int) * (na + rows) * nbp);
                if (FCC) bytes += (unsigned)(sizeof(T) * (na + 2 * rows) * lpi);
                mbar_arrive_expect_tx(&full[stage], bytes);
                if (FAC) {
                    bulk_g2s(sb + mp.oA + sizeof(T) * (size_t)(1 - prev) * sp, a.fac_alpha + rowa * sp, (unsigned)(sizeof(T) * na * sp), &full[stage]);
                    bulk_g2s(sb + mp.oB, a.fac_beta + rowc * sp, (unsigned)(sizeof(T) * rows * sp), &full[stage]);
                    bulk_g2s(sb + mp.oEa + sizeof(int) * (size_t)(1 - prev) * nbp, a.ctx.fac_ea + rowa * nbp, (unsigned)(sizeof(int) * na * nbp), &full[stage]);
                    bulk_g2s(sb + mp.oEb, a.ctx.fac_eb + rowc * nbp, (unsigned)(sizeof(int) * rows * nbp), &full[stage]);
                }
                if (FCC) {
                    bulk_g2s(sb + mp.oFa + sizeof(T) * (size_t)(1 - prev) * lpi, a.fcc_alpha + rowa * lpi, (unsigned)(sizeof(T) * na * lpi), &full[stage]);
                    bulk_g2s(sb + mp.oFb, a.fcc_beta + rowc * lpi, (unsigned)(sizeof(T) * rows * lpi), &full[stage]);
                    bulk_g2s(sb + mp.oFw, a.ctx.fcc_w + rowc * lpi, (unsigned)(sizeof(T) * rows * lpi), &full[stage]);
                }
            }
        }
        return;
    }

    // =========================== consumer warps =====================================================
    auto cbar = [&]() { named_bar_sync(1, GRAD_CONSUMERS); };

    // FCC register tile: thread owns rows i0..i0+1, columns j0..j0+3 of G (looped if LP*LP/8 > 256)
    constexpr int TCOLS = LP / 4;                 // tiles per row of G
    constexpr int NT = (LP / 2) * TCOLS;          // number of 2x4 tiles
    constexpr int TPT = (NT + GRAD_CONSUMERS - 1) / GRAD_CONSUMERS;
    T G[TPT][8];
#pragma unroll
    for (int q = 0; q < TPT; ++q)
#pragma unroll
        for (int e = 0; e < 8; ++e) G[q][e] = T(0);

    if (FAC) {
        for (int e = tid; e < L * L; e += GRAD_CONSUMERS) Gal[e] = T(0);
    }

    int cur_n = -1;
    int Tn = 0, Sn = 0;
    T gf = T(0), ga = T(0);

    // scatter the running per-position edge sums of utterance cur_n into the CTA's L x L matrix.
    // Thread l owns row l of Gal (destination label), so the order of additions is fixed.
    auto flush_fac = [&]() {
        if (!FAC || cur_n < 0) return;
        if (tid < L) {
            for (int k = lstart[tid]; k < lstart[tid + 1]; ++k) {
                const int s = order[k];
                Gal[tid * L + tid] += tH[s];
                if (s >= 1) Gal[tid * L + ylab[s - 1]] += tD[s];
            }
        }
    };

    int it = 0;
    for (int q = q_begin; q < q_end; ++q, ++it) {
        const int stage = it % NS, use = it / NS;
        const int n = q / a.tiles_t;
        const int t0 = (q % a.tiles_t) * TB;
        const int TBt = min(TB, Tt - t0);         // frames in this tile
        if (n != cur_n) {
            cbar();                                // every warp is done with the previous utterance's tiles
            flush_fac();
            cbar();
            cur_n = n;
            Tn = (int)a.ctx.il[n];
            gf = FCC ? a.g_full[(int64_t)n * a.g_full_stride] : T(0);
            ga = FAC ? a.g_aligned[(int64_t)n * a.g_aligned_stride] : T(0);
            if (FAC) {
                Sn = (int)a.ctx.tl[n];
                const int64_t *yn = a.y + (int64_t)n * a.y_s0;
                for (int s = tid; s < sp; s += GRAD_CONSUMERS) {
                    int lab = -1;
                    if (s < Sn) { lab = (int)yn[(int64_t)s * a.y_s1]; lab = max(0, min(L - 1, lab)); }
                    ylab[s] = lab;
                    hh[s] = a.ctx.fac_hd[(int64_t)n * sp + s];
                    din[s] = (s >= 1) ? a.ctx.fac_hd[((int64_t)N + n) * sp + s - 1] : T(0);
                    tH[s] = T(0);
                    tD[s] = T(0);
                }
                cbar();
                // CSR by label; positions ascending inside a label (fixed summation order).
                // Thread l counts, thread 0 scans the L counts, thread l fills its own segment.
                int cnt = 0;
                if (tid < L)
                    for (int s = 0; s < Sn; ++s) cnt += (ylab[s] == tid);
                if (tid < L) lstart[tid + 1] = cnt;
                cbar();
                if (tid == 0) {
                    int acc = 0;
                    lstart[0] = 0;
                    for (int l = 0; l < L; ++l) { acc += lstart[l + 1]; lstart[l + 1] = acc; }
                }
                cbar();
                if (tid < L) {
                    int pos = lstart[tid];
                    for (int s = 0; s < Sn; ++s) if (ylab[s] == tid) order[pos++] = s;
                }
            }
            cbar();
        }
        const int rows_valid = max(0, min(TBt, Tn - t0));   // frames t0 .. t0+rows_valid-1 are inside the utterance
        char *sb = stages + (size_t)stage * mp.stage_bytes;
        T *sA = (T *)(sb + mp.oA), *sB = (T *)(sb + mp.oB), *sBnd = (T *)(sb + mp.oBnd), *sZb = (T *)(sb + mp.oZb);
        int *sEa = (int *)(sb + mp.oEa), *sEb = (int *)(sb + mp.oEb);
        T *sFa = (T *)(sb + mp.oFa), *sFb = (T *)(sb + mp.oFb), *sFw = (T *)(sb + mp.oFw);

        mbar_wait(&full[stage], use & 1);

        // ================= phase 1: posteriors, one warp per frame ================================
        // smem row r of the alpha arrays = frame t0-1+r (r = 0 is the halo); beta arrays: row r-1 = frame t0-1+r
        if (rows_valid > 0) {
            const int r_lo = t0 > 0 ? 0 : 1;
            for (int r = r_lo + warp; r <= rows_valid; r += GRAD_CWARPS) {
                const int t = t0 - 1 + r;
                const bool iscur = r >= 1;
                const bool isprev = r < rows_valid;
                if (FAC) {
                    const bool single = nblk <= 32;
                    T zsum = T(0);
                    int cmax = INT_MIN;
                    T keep[8];
                    int keepc = INT_MIN;
                    for (int b0 = 0; b0 < nblk; b0 += 32) {
                        const int b = b0 + lane;
                        if (b >= nblk) continue;
                        T av[8];
                        lds8(sA + (size_t)r * sp + b * 8, av);
                        const int ea = sEa[r * nbp + b];
                        if (isprev) {
                            // the block's last value expressed in the NEXT block's exponent (what position 8(b+1) sees on its left)
                            const int ean = (b + 1 < nblk) ? sEa[r * nbp + b + 1] : ea;
                            sBnd[r * nbp + b] = av[7] * Num<T>::exp2i(ea - ean);
                        }
                        if (iscur) {
                            T bv[8];
                            lds8(sB + (size_t)(r - 1) * sp + b * 8, bv);
                            const int eb = sEb[(r - 1) * nbp + b];
                            T pm = T(0);
#pragma unroll
                            for (int j = 0; j < 8; ++j) { bv[j] *= av[j]; pm = tmax(pm, bv[j]); }
                            const int c = (pm > T(0)) ? ea + eb : INT_MIN;     // a block without mass is ignored
                            cmax = max(cmax, c);
                            if (single) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) keep[j] = bv[j];
                                keepc = c;
                            } else {
                                sts8(sB + (size_t)(r - 1) * sp + b * 8, bv);
                                sEb[(r - 1) * nbp + b] = c;
                            }
                        }
                    }
                    if (iscur) {
                        cmax = __reduce_max_sync(ASG_FULL_MASK, cmax);
                        // bring every block to the frame's common exponent, sum the frame
                        if (single) {
                            if (lane < nblk) {
                                T wgt = T(0);
                                if (keepc != INT_MIN && keepc - cmax >= -Num<T>::kMaxShift) wgt = Num<T>::exp2i(keepc - cmax);
#pragma unroll
                                for (int j = 0; j < 8; ++j) { keep[j] *= wgt; zsum += keep[j]; }
                                sts8(sB + (size_t)(r - 1) * sp + lane * 8, keep);
                            }
                        } else {
                            __syncwarp();
                            for (int b0 = 0; b0 < nblk; b0 += 32) {
                                const int b = b0 + lane;
                                if (b >= nblk) continue;
                                const int c = sEb[(r - 1) * nbp + b];
                                T wgt = T(0);
                                if (c != INT_MIN && c - cmax >= -Num<T>::kMaxShift) wgt = Num<T>::exp2i(c - cmax);
                                T pv[8];
                                lds8(sB + (size_t)(r - 1) * sp + b * 8, pv);
#pragma unroll
                                for (int j = 0; j < 8; ++j) { pv[j] *= wgt; zsum += pv[j]; }
                                sts8(sB + (size_t)(r - 1) * sp + b * 8, pv);
                            }
                        }
                        zsum = warp_sum(zsum);
                        if (lane == 0) sZb[r - 1] = (zsum > T(0)) ? ga / zsum : T(0);
                    }
                }
                if (FCC && iscur) {
                    T al0 = T(0), al1 = T(0), be0 = T(0), be1 = T(0), w0 = T(0), w1 = T(0);
                    if (lane < lpi) { al0 = sFa[r * lpi + lane]; be0 = sFb[(r - 1) * lpi + lane]; w0 = sFw[(r - 1) * lpi + lane]; }
                    if (LP > 32 && lane + 32 < lpi) {
                        al1 = sFa[r * lpi + lane + 32]; be1 = sFb[(r - 1) * lpi + lane + 32]; w1 = sFw[(r - 1) * lpi + lane + 32];
                    }
                    const T sc = __ldg(a.ctx.fcc_scale + (int64_t)n * tp + t);
                    const T g0 = al0 * be0, g1 = al1 * be1;
                    const T Z = warp_sum(g0 + g1);
                    const T inv = (Z > T(0)) ? gf / Z : T(0);
                    const T winv = (t >= 1) ? inv * sc : T(0);          // no incoming edge at frame 0
                    if (lane < lpi) { sFb[(r - 1) * lpi + lane] = g0 * inv; sFw[(r - 1) * lpi + lane] = w0 * winv; }
                    if (LP > 32 && lane + 32 < lpi) { sFb[(r - 1) * lpi + lane + 32] = g1 * inv; sFw[(r - 1) * lpi + lane + 32] = w1 * winv; }
                }
            }
        }
        cbar();

        // ================= phase 2 =================================================================
        // (a) grad_x rows: FCC posterior + label gather of the FAC posterior; each element written exactly once
        for (int p = tid; p < TBt * L; p += GRAD_CONSUMERS) {
            const int r = p / L, l = p - r * L;
            T val = T(0);
            if (r < rows_valid) {
                if (FCC) val = sFb[r * lpi + l];
                if (FAC) {
                    T acc = T(0);
                    const T *row = sB + (size_t)r * sp;
                    for (int k = lstart[l]; k < lstart[l + 1]; ++k) acc += row[order[k]];
                    val += acc * sZb[r];
                }
            }
            a.grad_x[((int64_t)(t0 + r) * N + n) * L + l] = val;
        }
        if (rows_valid > 0) {
            // (b) per-position edge posteriors, summed over the tile's frames (thread <-> position: fixed order)
            if (FAC) {
                for (int s = tid; s < Sn; s += GRAD_CONSUMERS) {
                    T accH = T(0), accD = T(0);
                    const T hs = hh[s], ds = din[s];
                    const bool edge = (s & 7) == 0;
                    const int lo = (s > 0) ? (edge ? (s >> 3) - 1 : s - 1) : 0;
                    for (int r = 1; r <= rows_valid; ++r) {
                        if (t0 - 1 + r < 1) continue;                       // frame 0 has no incoming edge
                        const T post = sB[(size_t)(r - 1) * sp + s] * sZb[r - 1];
                        const T stay = hs * sA[(size_t)(r - 1) * sp + s];
                        const T lft = (edge && s > 0) ? sBnd[(r - 1) * nbp + lo] : sA[(size_t)(r - 1) * sp + lo];
                        const T move = ds * lft;
                        const T den = stay + move;
                        if (den > T(0)) {
                            if (den < -Num<T>::neg_inf()) {
                                const T pd = post / den;                    // post is proportional to den: no overflow for tiny den
                                accH = fma(pd, stay, accH);
                                accD = fma(pd, move, accD);
                            } else {
                                accD += post;                               // the aligned neighbour overflowed: it dominates
                            }
                        }
                    }
                    tH[s] += accH;
                    tD[s] += accD;
                }
            }
            // (c) FCC transition gradient: G += w_t (x) alpha_{t-1}
            if (FCC) {
#pragma unroll
                for (int qq = 0; qq < TPT; ++qq) {
                    const int tile = tid + qq * GRAD_CONSUMERS;
                    if (tile < NT) {
                        const int i0 = (tile / TCOLS) * 2, j0 = (tile % TCOLS) * 4;
                        if (i0 < lpi && j0 < lpi) {
                            for (int r = 1; r <= rows_valid; ++r) {
                                if (t0 - 1 + r < 1) continue;
                                const T wa = sFw[(r - 1) * lpi + i0], wb = sFw[(r - 1) * lpi + i0 + 1];
                                const T *pp = sFa + (r - 1) * lpi + j0;
                                const T p0 = pp[0], p1 = pp[1], p2 = pp[2], p3 = pp[3];
                                G[qq][0] = fma(wa, p0, G[qq][0]); G[qq][1] = fma(wa, p1, G[qq][1]);
                                G[qq][2] = fma(wa, p2, G[qq][2]); G[qq][3] = fma(wa, p3, G[qq][3]);
                                G[qq][4] = fma(wb, p0, G[qq][4]); G[qq][5] = fma(wb, p1, G[qq][5]);
                                G[qq][6] = fma(wb, p2, G[qq][6]); G[qq][7] = fma(wb, p3, G[qq][7]);
                            }
                        }
                    }
                }
            }
        }
        // release the stage: this thread's generic-proxy accesses are ordered before the next bulk copy
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
    }
    cbar();
    flush_fac();
    cbar();

    // ---- per-CTA partial of grad_A = E o G (FCC) + Gal (FAC) -------------------------------------
    T *part = a.partials + (size_t)blockIdx.x * L * L;
#pragma unroll
    for (int qq = 0; qq < TPT; ++qq) {
        const int tile = tid + qq * GRAD_CONSUMERS;
        if (tile < NT) {
            const int i0 = (tile / TCOLS) * 2, j0 = (tile % TCOLS) * 4;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int i = i0 + (e >> 2), j = j0 + (e & 3);
                if (i < L && j < L) {
                    T v = T(0);
                    if (FCC) v = a.ctx.ehat[i * L + j] * G[qq][e];
                    if (FAC) v += Gal[i * L + j];
                    part[i * L + j] = v;
                }
            }
        }
    }
}

// grad_A[e] = sum over CTAs of partial[c][e].  Fixed order: warp w sums parts w, w+8, ...; warp 0 adds the 8 sums.
template <typename T>
__global__ void __launch_bounds__(256) grad_finalize_kernel(const T *partials, int nparts, int LL, T *grad_A) {
    __shared__ T red[8][33];
    const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5;
    const int e = blockIdx.x * 32 + lane;
    T acc0 = T(0), acc1 = T(0);
    if (e < LL) {
        int c = seg;
        for (; c + 8 < nparts; c += 16) {
            acc0 += partials[(size_t)c * LL + e];
            acc1 += partials[(size_t)(c + 8) * LL + e];
        }
        if (c < nparts) acc0 += partials[(size_t)c * LL + e];
    }
    red[seg][lane] = acc0 + acc1;
    __syncthreads();
    if (seg == 0 && e < LL) {
        T s = T(0);
#pragma unroll
        for (int q = 0; q < 8; ++q) s += red[q][lane];
        grad_A[e] = s;
    }
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
    const int L = a.d.L, sp = (int)a.ctx.sp, nbp = (int)a.ctx.nbp, lpi = (int)a.ctx.lpi;
    int device = 0;
    cudaGetDevice(&device);
    // frames per tile / ring depth: the largest that lets two CTAs share an SM, else whatever fits
    static const int cand[][2] = {{16, 2}, {8, 2}, {4, 2}, {2, 2}, {1, 2}, {4, 1}, {2, 1}, {1, 1}};
    int tb = 0, ns = 0;
    size_t bytes = 0;
    int force_tb = 0;
    if (const char *env = getenv("ASG_GRAD_TB")) force_tb = atoi(env);
    for (int pass = 0; pass < 2 && !tb; ++pass) {
        const size_t limit = pass == 0 ? 112 * 1024 : 227 * 1024;
        for (auto &c : cand) {
            if (force_tb && c[0] != force_tb) continue;
            const size_t b = grad_map(sizeof(T), L, lpi, sp, nbp, c[0], c[1], FCC, FAC).total;
            if (b <= limit) { tb = c[0]; ns = c[1]; bytes = b; break; }
        }
    }
    if (!tb) return cudaErrorInvalidValue;
    a.tb = tb;
    a.nstage = ns;
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
        grad_finalize_kernel<T><<<(LL + 31) / 32, 256, 0, st>>>(a.partials, grid, LL, a.grad_A);
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
