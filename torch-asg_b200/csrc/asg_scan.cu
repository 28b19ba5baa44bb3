// asg_scan.cu -- the four ASG recursions as persistent, warp-specialised sm_100a scan kernels.
//
// Replaces (reference, /root/reference/torch_asg/native):
//   fully_connected_alpha_recursion / _beta_recursion   fully_connected_lattice.cpp:9-29, :32-47
//   force_aligned_alpha_recursion / _beta_recursion     force_aligned_lattice.cpp:84-111, :113-154
//   make_aligned_inputs / make_aligned_transition       force_aligned_lattice.cpp:15-82 (+ _kernel.cu)
//   roll_to_end                                         utils.cpp:40-66 (length predicates instead)
//   score read-outs                                     fully_connected_lattice.cpp:89, force_aligned_lattice.cpp:316
//
// Shape of the computation: the T axis is strictly serial, so every (utterance, recursion) pair is
// one latency-bound chain.  Each chain gets its own CTA: one (FCC) or W (FAC) *consumer* warps that do
// nothing but the recurrence, fed by one *producer* warp that streams the utterance's emission rows
// from HBM FB frames at a time (cp.async ring, NRAW batches in flight), turns each batch into
// per-frame-normalised exponentials u_t = exp(x_t - max x_t) with four lanes per frame, and publishes
// them through a shared-memory ring.  The FAC consumers gather u_t[y_s] for their own target positions
// straight from that ring (labels live in registers), so nothing on the producer side scales with S.
// The recurrences run in the real domain -- no exp/log on the critical path:
//   FCC: P_t = s_t * u_t o (E P_{t-1}),  E = exp(A - max A) held in REGISTERS (one row per lane, the
//        rows beyond 32 split four-ways by column), s_t = exact power of two from a single REDUX max.
//   FAC: bidiagonal; 8 consecutive target positions per lane, block floating point (one power-of-two
//        exponent per lane), neighbours exchanged with one shuffle; no warp-wide reduction at all.
// State is written utterance-major ([N,tp,*], see asg_layout_t): every chain writes one sequential stream.
// See DESIGN.md for the numerics and tests/scaled_model.py for the executable specification.
#include <mutex>
#include <vector>

#include "asg_common.cuh"
#include "asg_internal.h"

namespace asg {

constexpr int FB = 8;             // frames per producer batch (4 lanes per frame)
constexpr int NRAW = 4;           // raw batches in flight (cp.async groups)
constexpr int RING = NRAW * FB;   // depth of the u ring, in frames
constexpr int SCAN_CARVEOUT_PCT = 50;   // preferred shared-memory carve-out of both scan kernels (percent of the maximum)

// ---------------------------------------------------------------------------------------------
// producer: x rows -> u ring.  Shared by the FCC and FAC kernels.
//   raw   [NRAW][FB][rp]   landing zone of the cp.async copies
//   uring [RING][up]       u_t = exp(x_t - m_t); entries l >= L of a row are 0
//   consumed_min()         frames every consumer has finished with
// ---------------------------------------------------------------------------------------------
template <typename T, bool BETA, typename ConsumedFn, typename FlushFn>
__device__ __forceinline__ void scan_producer(const ScanArgs<T> &a, T *raw, int rp, T *uring, int up, int *produced,
                                              double *msum_out, int *msum_ready, ConsumedFn consumed_min, FlushFn flush,
                                              int ngroups, int n, int Tn, int lane, const int *glab = nullptr, int glen = 0) {
    // glab != NULL (large alphabets): gather mode -- a ring row holds u_t[y_s] for the target positions s < glen
    // (labels glab[s], -1 = padding) instead of all L labels; the frame maximum comes from the prepass (a.fmax)
    const bool gather = glab != nullptr;
    const int L = gather ? glen : a.d.L;
    const int f = lane >> 2, sub = lane & 3;
    const T *xn = a.x + (int64_t)n * a.x_sn;
    const T ninf = Num<T>::neg_inf();
    auto issue = [&](int b) {
        const int k = b * FB + f;
        if (k < Tn) {
            const int t = BETA ? (Tn - 1 - k) : k;
            const T *row = xn + (int64_t)t * a.x_st;
            T *dst = raw + ((size_t)(b % NRAW) * FB + f) * rp;
            if (gather) {
                for (int l = sub; l < L; l += 4) { const int lb = glab[l]; if (lb >= 0) cp_async<sizeof(T)>(dst + l, row + (int64_t)lb * a.x_sl); }
                if (sub == 0) cp_async<sizeof(T)>(dst + rp - 1, a.fmax + (int64_t)n * a.ctx.tp + t);
            } else {
                for (int l = sub; l < L; l += 4) cp_async<sizeof(T)>(dst + l, row + (int64_t)l * a.x_sl);
            }
        }
    };
    const int nb = (Tn + FB - 1) / FB;
    // Fast path (contiguous, 16-byte aligned rows; L a multiple of 4): 16-byte copies, the row length walked in four
    // predicated, unrolled vector steps, pointers advanced by one batch stride -- ~150 instructions per 8-frame batch
    // instead of ~650 (measured: with the element-wise path the PRODUCER, not the recurrence, set the pace of both scan
    // kernels at 176 ns per frame).
    const bool vec = !gather && a.x_sl == 1 && (L & 3) == 0 && (a.x_st & 3) == 0 && (a.x_sn & 3) == 0 &&
                     ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0) && sizeof(T) == 4 && L <= 64 && (rp & 3) == 0 && (up & 3) == 0;
    if (vec) {
        const int V = L >> 2;                                   // 16-byte vectors per row
        const int64_t bstep = (BETA ? -(int64_t)FB : (int64_t)FB) * a.x_st;
        // this lane: frame f of a batch, vectors sub, sub+4, sub+8, sub+12
        const T *src0 = xn + (int64_t)(BETA ? (Tn - 1 - f) : f) * a.x_st + 4 * sub;     // batch 0
        const unsigned raw0 = smem_u32(raw) + (unsigned)sizeof(T) * (unsigned)(f * rp + 4 * sub);
        const unsigned ur0 = smem_u32(uring) + (unsigned)sizeof(T) * (unsigned)(4 * sub);
        auto issue_v = [&](int b) {
            if (b * FB + f < Tn) {
                const char *src = reinterpret_cast<const char *>(src0 + (int64_t)b * bstep);
                const unsigned dst = raw0 + (unsigned)sizeof(T) * (unsigned)((b % NRAW) * FB * rp);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (sub + 4 * i < V)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst + 64u * i), "l"(src + 64 * i) : "memory");
            }
        };
        for (int b = 0; b < NRAW; ++b) {
            if (b < nb) issue_v(b);
            cp_async_commit();
        }
        double msum = 0.0;
        bool uflow = false;                                     // a finite emission underflowed against its frame's maximum
        int b = 0, g = 0;
        while (b < nb || g < ngroups) {
            const int c = consumed_min();
            bool progressed = false;
            if (g < ngroups && c >= min((g + 1) * FB, Tn)) {
                flush(g);
                ++g;
                progressed = true;
            }
            const int hi = min((b + 1) * FB, Tn);
            if (b < nb && hi - RING <= c) {
                cp_async_wait<NRAW - 1>();
                const int k = b * FB + f;
                const bool valid = k < Tn;
                const unsigned src = raw0 + (unsigned)sizeof(T) * (unsigned)((b % NRAW) * FB * rp);
                float4 xv[4];
                float m = Num<float>::neg_inf();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (valid && sub + 4 * i < V) {
                        xv[i] = lds_v4f32(src + 64u * i);
                        m = fmaxf(fmaxf(m, fmaxf(xv[i].x, xv[i].y)), fmaxf(xv[i].z, xv[i].w));
                    }
                }
                m = fmaxf(m, __shfl_xor_sync(ASG_FULL_MASK, m, 1));
                m = fmaxf(m, __shfl_xor_sync(ASG_FULL_MASK, m, 2));
                if (!(m > Num<float>::neg_inf())) m = 0.f;       // an all -inf frame
                const unsigned dst = ur0 + (unsigned)sizeof(T) * (unsigned)((k % RING) * up);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (valid && sub + 4 * i < V) {
                        float4 u;
                        u.x = expf(xv[i].x - m); u.y = expf(xv[i].y - m); u.z = expf(xv[i].z - m); u.w = expf(xv[i].w - m);
                        const float ninf_ = Num<float>::neg_inf();
                        uflow |= ((u.x == 0.f) & (xv[i].x > ninf_)) | ((u.y == 0.f) & (xv[i].y > ninf_)) | ((u.z == 0.f) & (xv[i].z > ninf_)) |
                                 ((u.w == 0.f) & (xv[i].w > ninf_));
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(dst + 64u * i), "f"(u.x), "f"(u.y), "f"(u.z), "f"(u.w) : "memory");
                    }
                }
                if (valid && sub == 0) msum += (double)m;
                if (b + NRAW < nb) issue_v(b + NRAW);
                cp_async_commit();
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    st_volatile_s32(produced, hi);
                }
                ++b;
                progressed = true;
                if (b == nb) {
                    cp_async_wait<0>();
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) msum += __shfl_xor_sync(ASG_FULL_MASK, msum, o);
                    if (__any_sync(ASG_FULL_MASK, uflow) && lane == 0) atomicOr(a.ctx.flags(n), RF_EMISSION_UNDERFLOW);
                    if (lane == 0) {
                        *msum_out = msum;
                        __threadfence_block();
                        st_volatile_s32(msum_ready, 1);
                    }
                }
            }
            if (!progressed) __nanosleep(32);
        }
        if (nb == 0 && lane == 0) {
            *msum_out = 0.0;
            __threadfence_block();
            st_volatile_s32(msum_ready, 1);
        }
        return;
    }
    for (int b = 0; b < NRAW; ++b) {
        if (b < nb) issue(b);
        cp_async_commit();
    }
    double msum = 0.0;
    int b = 0, g = 0;       // next batch to produce, next group of finished state rows to flush
    while (b < nb || g < ngroups) {
        const int c = consumed_min();
        bool progressed = false;
        if (g < ngroups && c >= min((g + 1) * FB, Tn)) {         // the consumer has finished frames of group g
            flush(g);
            ++g;
            progressed = true;
        }
        const int hi = min((b + 1) * FB, Tn);
        if (b < nb && hi - RING <= c) {                          // the ring slots of batch b have been released
            cp_async_wait<NRAW - 1>();
            const int k = b * FB + f;
            const bool valid = k < Tn;
            const T *src = raw + ((size_t)(b % NRAW) * FB + f) * rp;
            T m = ninf;
            if (valid && !gather)
                for (int l = sub; l < L; l += 4) m = tmax(m, src[l]);
            if (valid && gather) m = src[rp - 1];
            m = tmax(m, __shfl_xor_sync(ASG_FULL_MASK, m, 1));
            m = tmax(m, __shfl_xor_sync(ASG_FULL_MASK, m, 2));
            if (!(m > ninf)) m = T(0);                     // an all -inf frame
            T *dst = uring + (size_t)(k % RING) * up;
            if (valid)                                     // (slots of frames >= Tn belong to frames not yet consumed)
                for (int l = sub; l < up; l += 4) {
                    const bool on = l < L && (!gather || glab[l] >= 0);
                    const T uu = on ? Num<T>::exp_(src[l] - m) : T(0);
                    if (on && uu == T(0) && src[l] > ninf) atomicOr(a.ctx.flags(n), RF_EMISSION_UNDERFLOW);     // rare
                    dst[l] = uu;
                }
            if (valid && sub == 0) msum += (double)m;
            if (b + NRAW < nb) issue(b + NRAW);            // each lane re-fills only elements it has consumed itself
            cp_async_commit();
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                st_volatile_s32(produced, hi);
            }
            ++b;
            progressed = true;
            if (b == nb) {                                 // all frames published: hand over the sum of the maxima
                cp_async_wait<0>();
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) msum += __shfl_xor_sync(ASG_FULL_MASK, msum, o);
                if (lane == 0) {
                    *msum_out = msum;
                    __threadfence_block();
                    st_volatile_s32(msum_ready, 1);
                }
            }
        }
        if (!progressed) __nanosleep(32);
    }
    if (nb == 0 && lane == 0) {
        *msum_out = 0.0;
        __threadfence_block();
        st_volatile_s32(msum_ready, 1);
    }
}

__device__ __forceinline__ double wait_msum(const double *msum, const int *msum_ready) {
    while (ld_volatile_s32(msum_ready) == 0) __nanosleep(32);
    __threadfence_block();
    return *reinterpret_cast<const volatile double *>(msum);
}

// =============================================================================================
// FCC
// =============================================================================================
template <typename T>
__device__ __forceinline__ T ehat_of(const T *A, int64_t s0, int64_t s1, int i, int j, int L, T R) {
    if (i >= L || j >= L) return T(0);
    T a = A[i * s0 + j * s1];
    return (a == Num<T>::neg_inf()) ? T(0) : Num<T>::exp_(a - R);
}

// max over the finite entries of A (0 if there are none), computed by one warp
template <typename T>
__device__ __forceinline__ T transition_max(const T *A, int64_t s0, int64_t s1, int L, int lane) {
    T m = Num<T>::neg_inf();
    for (int e = lane; e < L * L; e += 32) {
        T a = A[(e / L) * s0 + (e % L) * s1];
        if (a < -Num<T>::neg_inf()) m = tmax(m, a);   // skips +inf and NaN
    }
    m = Num<T>::warp_max(m);
    return (m > Num<T>::neg_inf()) ? m : T(0);
}

template <typename T, int NE> struct RowRegs;
// dot product of a register-resident row of E with the state vector in shared memory.
// float: packed FFMA2 (two fp32 FMAs per instruction, sm_100) on register pairs; double: scalar FMAs.
template <int NE> struct RowRegs<float, NE> {
    float2 e[NE / 2];
    __device__ __forceinline__ void set(int j, float v) { if (j & 1) e[j >> 1].y = v; else e[j >> 1].x = v; }
    // p: NE contiguous floats, 8-byte aligned (16-byte aligned when VEC4)
    template <bool VEC4> __device__ __forceinline__ float dot(const float *p) const {
        float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
        if (VEC4) {
#pragma unroll
            for (int i = 0; i < NE / 4; ++i) {
                const float4 q = *reinterpret_cast<const float4 *>(p + 4 * i);
                if (i & 1) { a2 = __ffma2_rn(e[2 * i], make_float2(q.x, q.y), a2); a3 = __ffma2_rn(e[2 * i + 1], make_float2(q.z, q.w), a3); }
                else { a0 = __ffma2_rn(e[2 * i], make_float2(q.x, q.y), a0); a1 = __ffma2_rn(e[2 * i + 1], make_float2(q.z, q.w), a1); }
            }
        } else {
#pragma unroll
            for (int i = 0; i < NE / 2; ++i) {
                const float2 q = *reinterpret_cast<const float2 *>(p + 2 * i);
                if (i & 1) a1 = __ffma2_rn(e[i], q, a1); else a0 = __ffma2_rn(e[i], q, a0);
            }
        }
        a0 = __fadd2_rn(__fadd2_rn(a0, a1), __fadd2_rn(a2, a3));
        return a0.x + a0.y;
    }
};
template <int NE> struct RowRegs<double, NE> {
    double e[NE];
    __device__ __forceinline__ void set(int j, double v) { e[j] = v; }
    template <bool VEC4> __device__ __forceinline__ double dot(const double *p) const {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
        for (int j = 0; j < NE; j += 4) {
            a0 = fma(e[j], p[j], a0); a1 = fma(e[j + 1], p[j + 1], a1);
            if (j + 2 < NE) { a2 = fma(e[j + 2], p[j + 2], a2); a3 = fma(e[j + 3], p[j + 3], a3); }
        }
        return (a0 + a1) + (a2 + a3);
    }
};

constexpr int PR = 32;            // FCC state ring, frames (4 groups of FB): finished rows leave by bulk (TMA) stores
constexpr int PRG = PR / FB;

template <typename T, int LP>
struct FccSmem {
    alignas(128) T raw[NRAW * FB * LP];
    alignas(128) T uring[RING * LP];
    alignas(128) T pring[PR * LP];     // the vector that feeds the next step: alpha P (= the state), beta u o Q
    alignas(128) T qring[PR * LP];     // beta: the state Q
    alignas(16) T scring[PR];          // alpha: power of two applied at each frame
    double msum;
    int msum_ready;
    int produced;
    int consumed;
    int flushed;                       // groups whose shared-memory rows may be overwritten
};

// ring slot of step k.  beta walks frames backwards: inside a group the slots are reversed so that a group is a
// run of ascending frames, i.e. ONE contiguous chunk of the utterance-major state arrays.
template <bool BETA> __device__ __forceinline__ int pslot(int k) {
    return BETA ? ((k & (PR - 1)) ^ (FB - 1)) : (k & (PR - 1));
}

template <typename T, int LP, bool BETA>
__device__ void fcc_consumer(const ScanArgs<T> &a, FccSmem<T, LP> &sm, int n, int Tn, int lane) {
    constexpr bool FULL = (LP >= 32);
    constexpr int REM = FULL ? LP - 32 : LP;                      // rows handled with a column split
    constexpr int RP = REM <= 8 ? 8 : (REM <= 16 ? 16 : 32);
    constexpr int CH = 32 / RP;                                   // column chunks
    constexpr int CL = LP / CH;                                   // columns per chunk (even)
    constexpr int RBASE = FULL ? 32 : 0;
    const int L = a.d.L;
    const int rr = lane % RP, cc = lane / RP;
    const int row2 = RBASE + rr;
    const bool has2 = (REM > 0) && (rr < REM);

    // ---- transition matrix -> registers (real domain, max-normalised) -------------------------
    const T R = transition_max(a.A, a.A_s0, a.A_s1, L, lane);
    constexpr int NE1 = FULL ? LP : 4, NE2 = (REM > 0) ? CL : 4;
    RowRegs<T, NE1> E1;
    RowRegs<T, NE2> E2;
    if (FULL) {
#pragma unroll
        for (int j = 0; j < LP; ++j)
            E1.set(j, BETA ? ehat_of(a.A, a.A_s0, a.A_s1, j, lane, L, R) : ehat_of(a.A, a.A_s0, a.A_s1, lane, j, L, R));
    }
    if (REM > 0) {
#pragma unroll
        for (int c = 0; c < CL; ++c) {
            const int j = cc * CL + c;
            E2.set(c, !has2 ? T(0)
                            : (BETA ? ehat_of(a.A, a.A_s0, a.A_s1, j, row2, L, R) : ehat_of(a.A, a.A_s0, a.A_s1, row2, j, L, R)));
        }
    }
    if (!BETA && n == 0) {   // publish E for the gradient assembly
        for (int e = lane; e < L * L; e += 32) a.ctx.ehat[e] = ehat_of(a.A, a.A_s0, a.A_s1, e / L, e % L, L, R);
    }
    if (!BETA && lane == 0) a.ctx.il[n] = Tn;
    if (Tn <= 0) {
        if (lane == 0 && BETA && a.full_scores) a.full_scores[n] = T(0);
        return;
    }
    for (int l = lane; l < PR * LP; l += 32) { sm.pring[l] = T(0); if (BETA) sm.qring[l] = T(0); }
    __syncwarp();

    int esum = 0;
    T sc = T(1);                // power of two applied at the NEXT step (from the maximum of the vector that feeds it)
    int e_next = 0;
    T f1 = T(0), f2 = T(0);     // this lane's rows of the feeding vector
    const bool wlo = has2 && cc == 0;

    // group bookkeeping at k % FB == 0: frames < k are finished (their u slots may be refilled, their state rows may
    // leave); frames k .. k+FB-1 must be in the u ring; the state slots about to be overwritten must have left.
    auto group_sync = [&](int k) {
        __threadfence_block();                                  // this lane's ring rows are visible before `consumed` moves
        __syncwarp();
        if (lane == 0) st_volatile_s32(&sm.consumed, k);
        const int need = min(k + FB, Tn);
        while (ld_volatile_s32(&sm.produced) < need) __nanosleep(20);
        const int gneed = k / FB - PRG + 1;
        while (ld_volatile_s32(&sm.flushed) < gneed) __nanosleep(20);
        __threadfence_block();
    };
    auto publish = [&](int k, T v1, T v2, T w1, T w2, T applied) {
        const int slot = pslot<BETA>(k) * LP;
        if (FULL) { sm.pring[slot + lane] = BETA ? w1 : v1; if (BETA) sm.qring[slot + lane] = v1; }
        if (wlo) { sm.pring[slot + row2] = BETA ? w2 : v2; if (BETA) sm.qring[slot + row2] = v2; }
        if (!BETA && lane == 0) sm.scring[pslot<BETA>(k)] = applied;
        // the next step's scale: exact power of two from one REDUX over the feeding vector (off the dependent chain)
        const T M = Num<T>::warp_max_nonneg(tmax(BETA ? w1 : v1, BETA ? w2 : v2));
        Num<T>::pow2_norm(M, sc, e_next);
    };

    // ---- frame 0 --------------------------------------------------------------------------------
    group_sync(0);
    {
        const T *urow = sm.uring;
        const T u1 = FULL ? urow[lane] : T(0);
        const T u2 = has2 ? urow[row2] : T(0);
        T v1, v2;
        if (BETA) { v1 = (FULL && lane < L) ? T(1) : T(0); v2 = (has2 && row2 < L) ? T(1) : T(0); }
        else { v1 = u1; v2 = u2; }
        f1 = BETA ? v1 * u1 : v1; f2 = BETA ? v2 * u2 : v2;
        publish(0, v1, v2, f1, f2, T(1));
        __syncwarp();
    }
    // ---- frames 1 .. Tn-1 -------------------------------------------------------------------------
    for (int k = 1; k < Tn; ++k) {
        if ((k & (FB - 1)) == 0) group_sync(k);
        const T *urow = sm.uring + (k & (RING - 1)) * LP;
        const T u1 = FULL ? urow[lane] : T(0);
        const T u2 = has2 ? urow[row2] : T(0);
        const T *pin = sm.pring + pslot<BETA>(k - 1) * LP;
        const T applied = sc;
        esum += e_next;
        T z2 = T(0);
        if (REM > 0) {
            z2 = E2.template dot<false>(pin + cc * CL);
#pragma unroll
            for (int o = RP; o < 32; o <<= 1) z2 += __shfl_xor_sync(ASG_FULL_MASK, z2, o);
        }
        T z1 = T(0);
        if (FULL) z1 = E1.template dot<true>(pin);
        T v1, v2, w1, w2;
        if (BETA) { v1 = z1 * applied; v2 = z2 * applied; w1 = v1 * u1; w2 = v2 * u2; }
        else { v1 = z1 * (u1 * applied); v2 = z2 * (u2 * applied); w1 = v1; w2 = v2; }
        f1 = w1; f2 = w2;
        publish(k, v1, v2, w1, w2, applied);
        __syncwarp();
    }
    __threadfence_block();
    __syncwarp();
    if (lane == 0) st_volatile_s32(&sm.consumed, Tn);

    // ---- score: log(sum of the final weighted state) + bookkeeping -----------------------------
    T tot = warp_sum((FULL ? f1 : T(0)) + (wlo ? f2 : T(0)));
    const double msum = wait_msum(&sm.msum, &sm.msum_ready);
    if (lane == 0) {
        const double score = (double)Num<T>::log_(tot) + msum + (double)(Tn - 1) * (double)R +
                             (double)esum * 0.6931471805599453094;
        a.ctx.scalars[(int64_t)n * 8 + (BETA ? SC_FULL_BETA : SC_FULL_ALPHA)] = score;
        a.ctx.scalars[(int64_t)n * 8 + SC_R] = (double)R;
        if (BETA && a.full_scores) a.full_scores[n] = (T)score;
    }
}

// ---------------------------------------------------------------------------------------------
// float specialisation of the FCC consumer: same arithmetic as fcc_consumer<float,...>, written for a lone warp --
// all shared-memory traffic through 32-bit shared addresses (no per-step generic->shared conversion), the whole state
// vector loaded up front (the loads' latency overlaps instead of being paid once per FFMA2), packed FFMA2.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 lds_v4(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 lds_v2(unsigned addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];\n" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}

// Work split of one step z = E p (LP x LP, LP a multiple of 8) over the 32 lanes of the lone consumer warp:
//   lane = (g = lane / 4, q = lane % 4) owns the RG = LP/8 rows g*RG .. and the CQ = LP/4 columns q*CQ .. of E (in registers),
// so a lane loads only ITS quarter of the feeding vector (2-5 vector LDS instead of the whole vector), runs RG independent
// FFMA2 chains of CQ/2 links each (no long dependent chain: the issue rate, not the FMA latency, bounds the step), and the four
// column-quarters of a row meet in two shuffle rounds.  Row r of a group is then finished (emission, scale, publish) by the
// lane with q == r % 4.  Against the previous split (lane = one full row, the rows beyond 32 column-split): 25 FFMA2 + 7
// loads + 10 shuffles per step instead of 25 FFMA2 + 19 loads + 4 shuffles in two serial halves.
template <int LP, bool BETA>
__device__ void fcc_consumer_f32(const ScanArgs<float> &a, FccSmem<float, LP> &sm, int n, int Tn, int lane) {
    typedef float T;
    static_assert(LP % 8 == 0 && LP >= 8 && LP <= 64, "LP is a multiple of 8 up to 64");
    constexpr int RG = LP / 8, CQ = LP / 4, CQ2 = CQ / 2;
    constexpr bool V4 = (CQ % 4) == 0;                            // the quarter is 16-byte aligned: LDS.128, else LDS.64
    const int L = a.d.L;
    const int g = lane >> 2, q = lane & 3;
    const int rbase = g * RG, cbase = q * CQ;

    const T R = transition_max(a.A, a.A_s0, a.A_s1, L, lane);
    float2 E[RG][CQ2];
#pragma unroll
    for (int r = 0; r < RG; ++r)
#pragma unroll
        for (int k = 0; k < CQ2; ++k) {
            const int i = rbase + r, j = cbase + 2 * k;
            E[r][k].x = BETA ? ehat_of(a.A, a.A_s0, a.A_s1, j, i, L, R) : ehat_of(a.A, a.A_s0, a.A_s1, i, j, L, R);
            E[r][k].y = BETA ? ehat_of(a.A, a.A_s0, a.A_s1, j + 1, i, L, R) : ehat_of(a.A, a.A_s0, a.A_s1, i, j + 1, L, R);
        }
    if (!BETA && n == 0) {   // publish E for the gradient assembly
        for (int e = lane; e < L * L; e += 32) a.ctx.ehat[e] = ehat_of(a.A, a.A_s0, a.A_s1, e / L, e % L, L, R);
    }
    if (!BETA && lane == 0) a.ctx.il[n] = Tn;
    if (Tn <= 0) {
        if (lane == 0 && BETA && a.full_scores) a.full_scores[n] = T(0);
        return;
    }
    for (int l = lane; l < PR * LP; l += 32) { sm.pring[l] = T(0); if (BETA) sm.qring[l] = T(0); }
    __syncwarp();

    // rows this lane finishes: rbase + q and (RG > 4) rbase + q + 4
    const bool has0 = q < RG, has1 = q + 4 < RG;
    const int row0 = rbase + q, row1 = rbase + q + 4;
    const unsigned pr0 = smem_u32(sm.pring), qr0 = smem_u32(sm.qring), sc0 = smem_u32(sm.scring), ur0 = smem_u32(sm.uring);
    constexpr unsigned ROWB = 4u * LP;
    const unsigned o0 = 4u * (unsigned)(has0 ? row0 : 0), o1 = 4u * (unsigned)(has1 ? row1 : 0), oc = 4u * (unsigned)cbase;

    int esum = 0;
    T sc = T(1);
    int e_next = 0;
    T f0 = T(0), f1 = T(0);

    auto group_sync = [&](int k) {
        __threadfence_block();
        __syncwarp();
        if (lane == 0) st_volatile_s32(&sm.consumed, k);
        const int need = min(k + FB, Tn);
        while (ld_volatile_s32(&sm.produced) < need) __nanosleep(20);
        const int gneed = k / FB - PRG + 1;
        while (ld_volatile_s32(&sm.flushed) < gneed) __nanosleep(20);
        __threadfence_block();
    };
    // feed = the vector the next step multiplies (alpha: the state, beta: u o state); st = the state (beta only).
    // All stores are predicated (no divergent branch).  The maximum of the feeding vector goes through one REDUX whose
    // result (Mred) is only READ in the next step, after that step's loads have been issued: a lone warp issues in order,
    // so a result consumed right behind its producer would expose the full REDUX latency on every step.
    T Mred = T(0);
    const bool lane0 = lane == 0;
    unsigned prev_off = 0;                  // byte offset (inside pring) of the row the previous step published
    // One step.  uoff: byte offset of this frame's row in the u ring; slot: ring slot this frame's state goes to.
    // In the main loop both are compile-time offsets from a per-group base (the loop is unrolled over the 8 frames of a
    // group), so a step carries no address arithmetic, no loop counter and no group test.
    auto step = [&](unsigned uoff, unsigned slot) {
        const unsigned pin = pr0 + prev_off + oc;
        float2 p[CQ2];
        if (V4) {
#pragma unroll
            for (int i2 = 0; i2 < CQ / 4; ++i2) {
                const float4 v = lds_v4(pin + 16u * i2);
                p[2 * i2] = make_float2(v.x, v.y); p[2 * i2 + 1] = make_float2(v.z, v.w);
            }
        } else {
#pragma unroll
            for (int i2 = 0; i2 < CQ2; ++i2) p[i2] = lds_v2(pin + 8u * i2);
        }
        const T u0 = lds_f32(ur0 + uoff + o0);               // lanes without a row read row 0 of their group (weight 0 below)
        const T u1 = (RG > 4) ? lds_f32(ur0 + uoff + o1) : T(0);
        // this step's scale from the previous step's REDUX, behind the loads
        pow2_fast(Mred, sc, e_next);
        const T applied = sc;
        esum += e_next;
        float2 acc[RG];                                      // RG independent chains
#pragma unroll
        for (int r = 0; r < RG; ++r) acc[r] = __fmul2_rn(E[r][0], p[0]);
#pragma unroll
        for (int k2 = 1; k2 < CQ2; ++k2)
#pragma unroll
            for (int r = 0; r < RG; ++r) acc[r] = __ffma2_rn(E[r][k2], p[k2], acc[r]);
        // The four column-quarters of a row meet by REDUCE-SCATTER, not all-reduce: lane q only needs rows q and q + 4 of its
        // group.  Round one (lanes q ^ 2) halves the rows a lane carries, round two (q ^ 1) leaves one per quad: 3 shuffles +
        // 3 adds per quad of rows instead of 8 + 8.
        T zs[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) zs[r] = (r < RG) ? acc[r < RG ? r : 0].x + acc[r < RG ? r : 0].y : T(0);
        const T ua0 = BETA ? applied : u0 * applied, ua1 = BETA ? applied : u1 * applied;
        T zq[2] = {T(0), T(0)};
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            if (4 * m >= RG) continue;
            const bool hi2 = (q & 2) != 0, hi1 = (q & 1) != 0;
            // round one: keep the rows whose owner shares this lane's bit 1
            T am = hi2 ? zs[4 * m + 2] : zs[4 * m];
            am += __shfl_xor_sync(ASG_FULL_MASK, hi2 ? zs[4 * m] : zs[4 * m + 2], 2);
            T bm = T(0);
            if (4 * m + 1 < RG) {
                bm = hi2 ? zs[4 * m + 3] : zs[4 * m + 1];
                bm += __shfl_xor_sync(ASG_FULL_MASK, hi2 ? zs[4 * m + 1] : zs[4 * m + 3], 2);
            }
            // round two: keep the row this lane owns
            zq[m] = (hi1 ? bm : am) + __shfl_xor_sync(ASG_FULL_MASK, hi1 ? am : bm, 1);
        }
        const T z0 = zq[0], z1 = zq[1];
        T v0 = z0 * ua0, v1 = z1 * ua1;
        if (!has0) v0 = T(0);
        if (!(RG > 4 && has1)) v1 = T(0);
        const T w0 = BETA ? v0 * u0 : v0, w1 = BETA ? v1 * u1 : v1;
        f0 = w0; f1 = w1;
        const unsigned rowoff = slot * ROWB;
        sts_f32_if(pr0 + rowoff + o0, w0, has0);
        if (BETA) sts_f32_if(qr0 + rowoff + o0, v0, has0);
        if (RG > 4) { sts_f32_if(pr0 + rowoff + o1, w1, has1); if (BETA) sts_f32_if(qr0 + rowoff + o1, v1, has1); }
        if (!BETA) sts_f32_if(sc0 + 4u * slot, applied, lane0);
        Mred = Num<T>::warp_max_nonneg(fmaxf(w0, w1));
        prev_off = rowoff;
        __syncwarp();
    };

    // ---- frame 0 --------------------------------------------------------------------------------
    group_sync(0);
    {
        const T u0 = has0 ? lds_f32(ur0 + o0) : T(0);
        const T u1 = (RG > 4 && has1) ? lds_f32(ur0 + o1) : T(0);
        T v0, v1;
        if (BETA) { v0 = (has0 && row0 < L) ? T(1) : T(0); v1 = (has1 && row1 < L) ? T(1) : T(0); }
        else { v0 = u0; v1 = u1; }
        f0 = BETA ? v0 * u0 : v0; f1 = BETA ? v1 * u1 : v1;
        const unsigned rowoff = (unsigned)pslot<BETA>(0) * ROWB;
        sts_f32_if(pr0 + rowoff + o0, f0, has0);
        if (BETA) sts_f32_if(qr0 + rowoff + o0, v0, has0);
        if (RG > 4) { sts_f32_if(pr0 + rowoff + o1, f1, has1); if (BETA) sts_f32_if(qr0 + rowoff + o1, v1, has1); }
        if (!BETA) sts_f32_if(sc0 + 4u * (unsigned)pslot<BETA>(0), T(1), lane0);
        Mred = Num<T>::warp_max_nonneg(fmaxf(f0, f1));
        prev_off = rowoff;
        __syncwarp();
    }
    // ---- frames 1 .. Tn-1: the rest of group 0, whole groups (unrolled), the last partial group ------
    int k = 1;
    for (; k < min(FB, Tn); ++k) step((unsigned)(k & (RING - 1)) * ROWB, (unsigned)pslot<BETA>(k));
    for (; k + FB <= Tn; k += FB) {
        group_sync(k);
        const unsigned ubase = (unsigned)(k & (RING - 1)) * ROWB, sbase = (unsigned)(k & (PR - 1));
#pragma unroll
        for (int j = 0; j < FB; ++j) step(ubase + (unsigned)j * ROWB, sbase + (unsigned)(BETA ? FB - 1 - j : j));
    }
    if (k < Tn) {
        group_sync(k);
        for (; k < Tn; ++k) step((unsigned)(k & (RING - 1)) * ROWB, (unsigned)pslot<BETA>(k));
    }
    __threadfence_block();
    __syncwarp();
    if (lane == 0) st_volatile_s32(&sm.consumed, Tn);

    T tot = warp_sum(f0 + f1);
    const double msum = wait_msum(&sm.msum, &sm.msum_ready);
    if (lane == 0) {
        const double score = (double)Num<T>::log_(tot) + msum + (double)(Tn - 1) * (double)R +
                             (double)esum * 0.6931471805599453094;
        a.ctx.scalars[(int64_t)n * 8 + (BETA ? SC_FULL_BETA : SC_FULL_ALPHA)] = score;
        a.ctx.scalars[(int64_t)n * 8 + SC_R] = (double)R;
        if (BETA && a.full_scores) a.full_scores[n] = (T)score;
    }
}

// dispatch: float -> the tiled specialisation, double -> generic
template <typename T, int LP, bool BETA> struct FccConsumer {
    __device__ static void run(const ScanArgs<T> &a, FccSmem<T, LP> &sm, int n, int Tn, int lane) { fcc_consumer<T, LP, BETA>(a, sm, n, Tn, lane); }
};
template <int LP, bool BETA> struct FccConsumer<float, LP, BETA> {
    __device__ static void run(const ScanArgs<float> &a, FccSmem<float, LP> &sm, int n, int Tn, int lane) {
        fcc_consumer_f32<LP, BETA>(a, sm, n, Tn, lane);
    }
};

// producer-side: group g of finished state rows leaves the shared-memory rings for the frame records
// (16-byte vector stores by the whole producer warp; a record part is one contiguous, sector-aligned run)
template <typename T, int LP, bool BETA>
__device__ __forceinline__ void fcc_flush_group(const ScanArgs<T> &a, FccSmem<T, LP> &sm, int n, int Tn, int g, int lane) {
    const int rows = min(FB, Tn - g * FB);
    if (a.ctx.rec != nullptr) {
        const int s0 = (g % PRG) * FB + (BETA ? FB - rows : 0);           // first ring slot of the run (ascending frames)
        const int t0 = BETA ? (Tn - g * FB - rows) : g * FB;                // first frame of the run
        const int64_t grow = (int64_t)n * a.ctx.tp + t0;
        constexpr int VPR = (int)(LP * sizeof(T) / 16);                     // 16-byte vectors per row
        const int o_feed = BETA ? a.ctx.o_cw : a.ctx.o_ca;                  // beta: the feed ring is u o Q, qring the state
        for (int idx = lane; idx < rows * VPR; idx += 32) {
            const int r = idx / VPR, v = idx - r * VPR;
            char *rec = a.ctx.record(grow + r);
            const uint4 f = *reinterpret_cast<const uint4 *>(reinterpret_cast<const char *>(sm.pring + (s0 + r) * LP) + 16 * v);
            __stcs(reinterpret_cast<uint4 *>(rec + o_feed + 16 * v), f);
            if (BETA) {
                const uint4 q = *reinterpret_cast<const uint4 *>(reinterpret_cast<const char *>(sm.qring + (s0 + r) * LP) + 16 * v);
                __stcs(reinterpret_cast<uint4 *>(rec + a.ctx.o_cb + 16 * v), q);
            }
        }
        if (!BETA)
            for (int r = lane; r < rows; r += 32) *a.ctx.fcc_scale(grow + r) = sm.scring[s0 + r];
    }
    __syncwarp();
    if (lane == 0) st_volatile_s32(&sm.flushed, g + 1);                     // the rows are in registers / on their way: slots free
}

template <typename T, int LP>
__global__ void __launch_bounds__(64, 1) fcc_scan_kernel(const ScanArgs<T> a) {
    extern __shared__ __align__(128) char fcc_smem_raw[];
    FccSmem<T, LP> &sm = *reinterpret_cast<FccSmem<T, LP> *>(fcc_smem_raw);
    const int n = blockIdx.x;
    const int dir = blockIdx.y == 0 ? a.dir0 : a.dir1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { sm.produced = 0; sm.consumed = 0; sm.msum_ready = 0; sm.flushed = 0; }
    for (int l = threadIdx.x; l < RING * LP; l += blockDim.x) sm.uring[l] = T(0);   // the producer's vector path only writes l < L
    __syncthreads();
    int Tn = a.il ? (int)a.il[n] : a.d.T;
    Tn = max(0, min(Tn, a.d.T));
    auto cons = [&]() { return ld_volatile_s32(&sm.consumed); };
    const int ngroups = (Tn + FB - 1) / FB;
    if (dir == 0) {
        if (warp == 0) FccConsumer<T, LP, false>::run(a, sm, n, Tn, lane);
        else scan_producer<T, false>(a, sm.raw, LP, sm.uring, LP, &sm.produced, &sm.msum, &sm.msum_ready, cons,
                                     [&](int g) { fcc_flush_group<T, LP, false>(a, sm, n, Tn, g, lane); }, ngroups, n, Tn, lane);
    } else {
        if (warp == 0) FccConsumer<T, LP, true>::run(a, sm, n, Tn, lane);
        else scan_producer<T, true>(a, sm.raw, LP, sm.uring, LP, &sm.produced, &sm.msum, &sm.msum_ready, cons,
                                    [&](int g) { fcc_flush_group<T, LP, true>(a, sm, n, Tn, g, lane); }, ngroups, n, Tn, lane);
    }
}

// =============================================================================================
// FAC
// =============================================================================================
template <typename T> struct FacSmemView {
    T *raw;        // [NRAW*FB][lpi]
    T *uring;      // [RING][up]      up = lpi + 4: entries [lpi, up) are the always-zero slots
    T *pring;      // [RING][sp]      (one float consumer warp only) u by target position, pair-interleaved per block
    T *xval;       // [2][W+1]  boundary mantissas between consumer warps
    int *xexp;     // [2][W+1]
    int *ylab;     // [sp]
    T *red;        // [W]
    double *msum;  // [1]
    int *flags;    // [0] produced, [1] msum_ready
    int *consumed; // [W]
};

template <typename T>
__host__ __device__ inline size_t fac_smem_carve(FacSmemView<T> *v, char *base, int lpi, int sp, int W, bool posring) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 16); return o; };
    const int up = lpi + 4;
    size_t o_raw = take(sizeof(T) * NRAW * FB * up);
    size_t o_uring = take(sizeof(T) * RING * up);
    size_t o_pring = take(posring ? sizeof(T) * RING * sp : 0);
    size_t o_xval = take(sizeof(T) * 2 * (W + 1));
    size_t o_xexp = take(sizeof(int) * 2 * (W + 1));
    size_t o_ylab = take(sizeof(int) * sp);
    size_t o_red = take(sizeof(T) * W);
    size_t o_msum = take(sizeof(double));
    size_t o_flags = take(sizeof(int) * 4);
    size_t o_cons = take(sizeof(int) * W);
    if (v) {
        v->raw = (T *)(base + o_raw); v->uring = (T *)(base + o_uring); v->pring = posring ? (T *)(base + o_pring) : nullptr;
        v->xval = (T *)(base + o_xval); v->xexp = (int *)(base + o_xexp);
        v->ylab = (int *)(base + o_ylab); v->red = (T *)(base + o_red); v->msum = (double *)(base + o_msum);
        v->flags = (int *)(base + o_flags); v->consumed = (int *)(base + o_cons);
    }
    return off;
}

__device__ __forceinline__ void fac_consumer_sync(int W) {
    if (W == 1) __syncwarp();
    else named_bar_sync(1, 32 * W);
}

template <typename T, bool BETA>
__device__ void fac_consumer(const ScanArgs<T> &a, FacSmemView<T> &sm, int W, int n, int Tn, int Sn, T Rb,
                             const T (&h)[ASG_BLK], const T (&dd)[ASG_BLK], const int (&lab)[ASG_BLK], int warp, int lane) {
    // dd[j]: alpha -> edge (s-1 -> s) into own state j;  beta -> edge (s -> s+1) out of own state j
    const int sp = (int)a.ctx.sp, nblk = (int)a.ctx.nblk, nbp = (int)a.ctx.nbp;
    const int up = (a.fmax ? (int)a.ctx.sp : (int)a.ctx.lpi) + 4;
    const int64_t tp = a.ctx.tp;
    const int gl = warp * 32 + lane;            // global lane = block index
    const int s0 = gl * ASG_BLK;
    const bool owns = s0 < sp;
    constexpr int KNB = 64;                     // a block never sits more than 2^KNB below the neighbour that feeds it

    // state rows of this utterance are contiguous: walk them with pointers
    // this utterance's frame records are contiguous: walk them with pointers (record stride)
    const int64_t row0 = (int64_t)n * tp + (BETA ? Tn - 1 : 0);
    const bool store = owns && a.ctx.rec != nullptr;
    T *stp = (BETA ? a.ctx.fac_beta(row0) : a.ctx.fac_alpha(row0)) + s0;
    int *esp = (BETA ? a.ctx.fac_eb(row0) : a.ctx.fac_ea(row0)) + gl;
    const int64_t sstep = (BETA ? -a.ctx.R : a.ctx.R) / (int64_t)sizeof(T);
    const int64_t estep = (BETA ? -a.ctx.R : a.ctx.R) / (int64_t)sizeof(int);
    const bool store_e = store && gl < nblk;

    T v[ASG_BLK];        // alpha: P ; beta: Q (mantissas)
    T w[ASG_BLK];        // beta only: emission-weighted state fed to the next step
    T sv[ASG_BLK];       // what is stored for the frame: the state before normalisation, in units of 2^sE, maximum ~2^-sdrift
    int Eme = 0, sE = 0, sdrift = 0;
    bool drop_seen = false, coarse = false;
    T mymax = T(0);      // max of what feeds the next step (alpha: v, beta: w)
    int cur = 0;

    for (int k = 0; k < Tn; ++k) {
        const int t = BETA ? (Tn - 1 - k) : k;
        if ((k & (FB - 1)) == 0) {
            __threadfence_block();                                   // this warp's reads of the ring rows of frames < k come first
            __syncwarp();
            if (lane == 0) st_volatile_s32(&sm.consumed[warp], k);   // frames < k are done with
            const int need = min(k + FB, Tn);
            while (ld_volatile_s32(&sm.flags[0]) < need) __nanosleep(20);
            __threadfence_block();
        }
        const T *urow = sm.uring + (size_t)(k % RING) * up;
        T ub[ASG_BLK];
#pragma unroll
        for (int j = 0; j < ASG_BLK; ++j) ub[j] = urow[lab[j]];

        if (k == 0) {
#pragma unroll
            for (int j = 0; j < ASG_BLK; ++j) {
                if (BETA) v[j] = (s0 + j == Sn - 1) ? T(1) : T(0);
                else v[j] = (s0 + j == 0) ? ub[j] : T(0);
                sv[j] = v[j];
            }
        } else {
            // neighbour's boundary value and exponent from the previous step
            T nv; int ne;
            const T src = BETA ? w[0] : v[ASG_BLK - 1];
            if (BETA) { nv = __shfl_down_sync(ASG_FULL_MASK, src, 1); ne = __shfl_down_sync(ASG_FULL_MASK, Eme, 1); }
            else { nv = __shfl_up_sync(ASG_FULL_MASK, src, 1); ne = __shfl_up_sync(ASG_FULL_MASK, Eme, 1); }
            if (!BETA && lane == 0) {
                if (W > 1 && warp > 0) { nv = sm.xval[cur * (W + 1) + warp - 1]; ne = sm.xexp[cur * (W + 1) + warp - 1]; }
                else { nv = T(0); ne = Eme; }
            }
            if (BETA && lane == 31) {
                if (W > 1 && warp < W - 1) { nv = sm.xval[cur * (W + 1) + warp + 1]; ne = sm.xexp[cur * (W + 1) + warp + 1]; }
                else { nv = T(0); ne = Eme; }
            }
            // Common exponent of this step.  An empty block adopts its neighbour's exponent; a block that sits far below
            // the neighbour feeding it is pulled up (its own mantissas shrink by `fs`, harmlessly) so the inflow is
            // always representable -- no clamping of real mass.
            int Enew = Eme;
            if (mymax == T(0)) Enew = ne;
            else if (nv > T(0)) Enew = max(Eme, ne - KNB);
            const T fs = Num<T>::exp2i_flush(Eme - Enew);          // <= 1 (1 in the common case)
            nv *= Num<T>::exp2i_flush(ne - Enew);                  // more than the format's range below: flushed to zero, never clamped
            Eme = Enew;
            T r[ASG_BLK];
            if (BETA) {
                r[ASG_BLK - 1] = fma(h[ASG_BLK - 1], w[ASG_BLK - 1] * fs, dd[ASG_BLK - 1] * nv);
#pragma unroll
                for (int j = 0; j < ASG_BLK - 1; ++j) r[j] = fs * fma(h[j], w[j], dd[j] * w[j + 1]);
                if (s0 + ASG_BLK - 1 > t) {                    // band: position s is reachable by frame t only if s <= t
#pragma unroll
                    for (int j = 0; j < ASG_BLK; ++j) if (s0 + j > t) r[j] = T(0);
                }
            } else {
                r[0] = ub[0] * fma(h[0], v[0] * fs, dd[0] * nv);
#pragma unroll
                for (int j = 1; j < ASG_BLK; ++j) r[j] = (ub[j] * fs) * fma(h[j], v[j], dd[j] * v[j - 1]);
                const int lb = (Sn - 1) - (Tn - 1 - t);        // band: positions below lb can no longer finish
                if (s0 < lb) {
#pragma unroll
                    for (int j = 0; j < ASG_BLK; ++j) if (s0 + j < lb) r[j] = T(0);
                }
            }
            const T m = tmax(tmax(tmax(r[0], r[1]), tmax(r[2], r[3])), tmax(tmax(r[4], r[5]), tmax(r[6], r[7])));
            T sc; int e;
            Num<T>::pow2_norm(m, sc, e);
            // What is STORED for this frame is the state before normalisation together with the drift of its maximum, so
            // the gradient assembly knows how far below the block maximum entries may have been flushed in this step (range
            // report).  A drift that does not fit the exponent word: stored normalised (fac_pack_coarse), a drop beyond 2^900 is reported here.
            sdrift = -e;
            sE = Eme;
            const bool raw = sdrift < kFacDriftCoarse && sdrift >= -7;   // (a maximum above 2^7 is inflow from a far larger neighbour)
            if (!raw) {
                coarse = sdrift >= kFacDriftCoarse;                      // stored normalised under the reserved drift code
                drop_seen |= sdrift > kFacDriftCoarseValue;
                sdrift = 0;
                sE = Eme + e;
            } else coarse = false;
#pragma unroll
            for (int j = 0; j < ASG_BLK; ++j) { v[j] = r[j] * sc; sv[j] = raw ? r[j] : v[j]; }
            Eme += e;
        }
        // what feeds the next step, and its max
        if (BETA) {
#pragma unroll
            for (int j = 0; j < ASG_BLK; ++j) w[j] = v[j] * ub[j];
            mymax = tmax(tmax(tmax(w[0], w[1]), tmax(w[2], w[3])), tmax(tmax(w[4], w[5]), tmax(w[6], w[7])));
        } else {
            mymax = tmax(tmax(tmax(v[0], v[1]), tmax(v[2], v[3])), tmax(tmax(v[4], v[5]), tmax(v[6], v[7])));
        }

        // store mantissas + exponent
        if (store) {
            store8_stream(stp, sv);
            if (store_e) *esp = coarse ? fac_pack_coarse(sE) : fac_pack(sE, sdrift, tmax(tmax(tmax(sv[0], sv[1]), tmax(sv[2], sv[3])), tmax(tmax(sv[4], sv[5]), tmax(sv[6], sv[7]))) > T(0));
        }
        stp += sstep;
        esp += estep;
        if (W > 1) {
            if (!BETA && lane == 31) { sm.xval[(cur ^ 1) * (W + 1) + warp] = v[ASG_BLK - 1]; sm.xexp[(cur ^ 1) * (W + 1) + warp] = Eme; }
            if (BETA && lane == 0) { sm.xval[(cur ^ 1) * (W + 1) + warp] = w[0]; sm.xexp[(cur ^ 1) * (W + 1) + warp] = Eme; }
            cur ^= 1;
            named_bar_sync(1, 32 * W);
        }
    }
    if (lane == 0) st_volatile_s32(&sm.consumed[warp], Tn);
    if (__any_sync(ASG_FULL_MASK, drop_seen) && lane == 0) atomicOr(a.ctx.flags(n), RF_POSTERIOR_THIN);
    // ---- score ------------------------------------------------------------------------------------
    const double msum = (Tn > 0) ? wait_msum(sm.msum, &sm.flags[1]) : 0.0;
    if (gl == 0) {
        double score = 0.0;
        if (Tn > 0) {
            // beta: w[0] = u_0[y_0] * Q_0[0];  alpha: total mass in the final position S_n-1 is read below
            const T fin = BETA ? w[0] : T(0);
            score = (double)Num<T>::log_(fin) + msum + (double)(Tn - 1) * (double)Rb + (double)Eme * 0.6931471805599453094;
            if (BETA && !(fin > T(0)) && Sn <= Tn) atomicOr(a.ctx.flags(n), RF_SCORE_LOST);
        }
        if (BETA) {
            a.ctx.scalars[(int64_t)n * 8 + SC_AL_BETA] = score;
            a.ctx.scalars[(int64_t)n * 8 + SC_RB] = (double)Rb;
            if (a.aligned_scores) a.aligned_scores[n] = (T)score;
        }
    }
    if (!BETA && Tn > 0) {      // alpha-side score (cross-check): mass at (T_n-1, S_n-1)
        const int sl = Sn - 1;
        if (sl >= s0 && sl < s0 + ASG_BLK) {
            T fin = T(0);
#pragma unroll
            for (int j = 0; j < ASG_BLK; ++j) if (s0 + j == sl) fin = v[j];
            a.ctx.scalars[(int64_t)n * 8 + SC_AL_ALPHA] =
                (double)Num<T>::log_(fin) + msum + (double)(Tn - 1) * (double)Rb + (double)Eme * 0.6931471805599453094;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// float specialisation of the FAC consumer: the same arithmetic as fac_consumer<float,...>, written for the issue
// rate of a lone warp -- packed FMUL2/FFMA2 on position pairs (j, j+4) (the shifted neighbour pair of (j, j+4) is the
// aligned pair (j-1, j+3)), 32-bit shared-memory addressing for the label gather, pointer-walked stores.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float max8(const float2 (&p)[4]) {
    return fmaxf(fmaxf(fmaxf(p[0].x, p[0].y), fmaxf(p[1].x, p[1].y)), fmaxf(fmaxf(p[2].x, p[2].y), fmaxf(p[3].x, p[3].y)));
}

// One consumer warp (S <= 256).  Lane boundaries need no special cases: shfl_up/down hand the edge lane its OWN
// value, which a per-lane 0/1 mask turns into "no neighbour".  The boundary value travels UN-normalised, together
// with the exponent it is expressed in, as soon as it exists -- the shuffle overlaps the normalisation of the step.
template <bool BETA, bool VEC>
__device__ void fac_consumer_f32(const ScanArgs<float> &a, FacSmemView<float> &sm, int n, int Tn, int Sn, float Rb,
                                 const float (&h)[ASG_BLK], const float (&dd)[ASG_BLK], const int (&lab)[ASG_BLK], int lane) {
    typedef float T;
    const int sp = (int)a.ctx.sp, nblk = (int)a.ctx.nblk, nbp = (int)a.ctx.nbp;
    const int up = (a.fmax ? (int)a.ctx.sp : (int)a.ctx.lpi) + 4;
    const int64_t tp = a.ctx.tp;
    const int s0 = lane * ASG_BLK;
    const bool owns = s0 < sp;
    constexpr int KNB = 64;

    const int64_t row0 = (int64_t)n * tp + (BETA ? Tn - 1 : 0);
    const bool store = owns && a.ctx.rec != nullptr;
    T *stp = (BETA ? a.ctx.fac_beta(row0) : a.ctx.fac_alpha(row0)) + s0;
    int *esp = (BETA ? a.ctx.fac_eb(row0) : a.ctx.fac_ea(row0)) + lane;
    const int64_t sstep = (BETA ? -a.ctx.R : a.ctx.R) / (int64_t)sizeof(T);
    const int64_t estep = (BETA ? -a.ctx.R : a.ctx.R) / (int64_t)sizeof(int);
    const bool store_e = store && lane < nblk;

    // pairs: index q holds positions (q, q+4) of this lane's block
    float2 h2[4], d2[4];
    // u by position: sm.pring (two LDS.128 per step, already pair-interleaved) when the producer fills it, else the
    // label gather from the u ring (gather mode of the large alphabets: rows are indexed by position there too)
    constexpr bool vec = VEC;
    unsigned la[8];                                   // shared-memory byte address of u[lab[j]] in ring slot 0
    const unsigned ring0 = smem_u32(sm.uring);
#pragma unroll
    for (int q = 0; q < 4; ++q) { h2[q] = make_float2(h[q], h[q + 4]); d2[q] = make_float2(dd[q], dd[q + 4]); }
#pragma unroll
    for (int j = 0; j < 8; ++j) la[j] = ring0 + 4u * (unsigned)lab[j];
    const unsigned pa = vec ? smem_u32(sm.pring) + 32u * (unsigned)min(lane, nblk - 1) : 0u;
    const unsigned slot_bytes = vec ? 4u * (unsigned)sp : 4u * (unsigned)up;

    // The vector carried from step to step: alpha: the state P; beta: the feed W = Q o u (the state Q itself is only stored).
    // It ping-pongs between X and Y (a step reads one and writes the other), so no register copies are needed.
    float2 X[4], Y[4];
    int Eme = 0;         // the block's exponent: the mantissas are in units of 2^Eme
    float mcar = 0.f;    // maximum of the block's state as stored by the previous step
    float nv = 0.f;      // neighbour's boundary value of the previous step, in units of 2^ne
    int ne = 0;
    float nva = 0.f;     // the same value in THIS block's units (2^Eme)
    bool fix = false;    // some block of the warp needs the slow path before its next step
#pragma unroll
    for (int q = 0; q < 4; ++q) { X[q] = make_float2(0.f, 0.f); Y[q] = X[q]; }

    // LAZY block floating point.  A step is only the recurrence itself -- 12 packed multiply-adds on the block's eight
    // mantissas plus the neighbour's boundary value, which arrives already expressed in this block's units (`nva`; flushed
    // to zero when it is more than 2^126 below: real mass is never clamped), one shuffle of (boundary value, exponent)
    // -- and a range check.  The expensive part (maximum -> power of two -> rescale -> exponent update; re-basing the
    // block on the exponent of what flows in when mass first reaches an empty block or when the inflow would swamp the
    // block's own scale) runs only when some block of the warp asks for it: the block maximum has drifted outside
    // [2^-40, 4), or the inflow is >= 4 in the block's units.  That is a warp-uniform branch taken every few frames
    // instead of ~60 instructions on every frame.  Boundary values always travel WITH the exponent they are expressed in,
    // so blocks renormalise independently of each other.  Every stored exponent word also records the drift of the block
    // maximum (fac_pack), so the gradient assembly still sees normalised mantissas.
    auto to_mine = [&]() {                                         // nv * 2^(ne - Eme), flushed below 2^-126
        const int dl = min(max(ne - Eme, -127), 127);
        return nv * __int_as_float((dl + 127) << 23);
    };
    auto fix_block = [&](float2 (&V)[4]) {
        const float m = mcar;
        if (m > 0.f) {
            float sc; int e;
            Num<T>::pow2_norm(m, sc, e);
            const float2 sc2 = make_float2(sc, sc);
#pragma unroll
            for (int q = 0; q < 4; ++q) V[q] = __fmul2_rn(V[q], sc2);
            Eme += e;
        }
        nva = to_mine();
        if (nv > 0.f && (!(m > 0.f) || !(nva < 4.f))) {
            // re-base on the inflow: afterwards it is in [1,2) and this block's own mantissas have shrunk accordingly
            float sn; int en;
            Num<T>::pow2_norm(nv, sn, en);
            const int En = ne + en;
            const int sh = min(max(Eme - En, -127), 0);
            const float f = __int_as_float((sh + 127) << 23);
            const float2 f2 = make_float2(f, f);
#pragma unroll
            for (int q = 0; q < 4; ++q) V[q] = __fmul2_rn(V[q], f2);
            Eme = En;
            nva = nv * sn;
        } else if (!(m > 0.f)) {
            Eme = ne;                                               // empty and nothing flows in: follow the neighbour
            nva = 0.f;
        }
    };

    unsigned slot_off = 0;                            // byte offset of ring slot (k % RING)
    char *sptr = reinterpret_cast<char *>(stp), *eptr = reinterpret_cast<char *>(esp);
    const int64_t rstep = BETA ? -a.ctx.R : a.ctx.R;
    const bool lane0 = lane == 0;
    const bool real = lane < nblk;                    // lanes beyond the last block hold zeros and never ask for the slow path
    const bool edge = BETA ? lane == 31 : lane == 0;

    auto ring_sync = [&](int k) {                     // frames < k are done with; frames k .. k+FB-1 must be in the ring
        __threadfence_block();                        // this warp's reads of the ring rows of frames < k come first
        __syncwarp();
        if (lane0) st_volatile_s32(&sm.consumed[0], k);
        const int need = min(k + FB, Tn);
        while (ld_volatile_s32(&sm.flags[vec ? 2 : 0]) < need) __nanosleep(20);
        __threadfence_block();
        if ((k & (RING - 1)) == 0) slot_off = 0u;
    };
    auto load_u = [&](float2 (&u2)[4]) {
        if (vec) {
            const float4 ua = lds_v4f32(pa + slot_off), ub = lds_v4f32(pa + slot_off + 16u);
            u2[0] = make_float2(ua.x, ua.y); u2[1] = make_float2(ua.z, ua.w);
            u2[2] = make_float2(ub.x, ub.y); u2[3] = make_float2(ub.z, ub.w);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) { u2[q].x = lds_f32(la[q] + slot_off); u2[q].y = lds_f32(la[q + 4] + slot_off); }
        }
        slot_off += slot_bytes;
    };
    // exchange, range check, stores: the common tail of every step (St: the state to store, snd: the boundary value).
    // No lane-divergent branch anywhere in a step: the stores are predicated.
    auto finish = [&](const float2 (&St)[4], float snd) {
        nv = BETA ? __shfl_down_sync(ASG_FULL_MASK, snd, 1) : __shfl_up_sync(ASG_FULL_MASK, snd, 1);
        ne = BETA ? __shfl_down_sync(ASG_FULL_MASK, Eme, 1) : __shfl_up_sync(ASG_FULL_MASK, Eme, 1);
        nv = edge ? 0.f : nv;                                      // an edge lane gets its own value back: no neighbour
        nva = to_mine();
        const float m = max8(St);
        mcar = m;
        const unsigned me = (unsigned)__float_as_int(m) >> 23;     // m >= 0
        // block maximum out of [2^-40, 4) (also inf/NaN) | mass reaching an empty block | inflow beyond the block's scale
        const bool live = me != 0u;
        const bool flag = real & ((live & ((me - 87u) > 41u)) | ((nv > 0.f) & ((!live) | !(nva < 4.f))));
        fix = __any_sync(ASG_FULL_MASK, flag);
        stg_cs_v4_if(sptr, St[0].x, St[1].x, St[2].x, St[3].x, store);
        stg_cs_v4_if(sptr + 16, St[0].y, St[1].y, St[2].y, St[3].y, store);
        stg_s32_if(eptr, fac_pack(Eme, 127 - (int)me, live), store_e);          // E, the drift of the block maximum, live / empty
        sptr += rstep;
        eptr += rstep;
    };
    // No band masks here: a position that cannot be reached from the start (s > t) never receives mass, and one that can
    // no longer reach the end has beta = 0 exactly, so its alpha never enters a posterior (the generic consumer zeroes both).
    auto step = [&](float2 (&src)[4], float2 (&dst)[4]) {
        float2 u2[4];
        load_u(u2);
        if (fix) fix_block(src);                               // warp-uniform
        if (BETA) {
            // q[j] = h[j] w[j] + d[j] w[j+1];  pair (j, j+4) needs (w[j+1], w[j+5]) = pair j+1; j = 3: (w[4], inflow)
            float2 Q[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 nx = (q < 3) ? src[(q + 1) & 3] : make_float2(src[0].y, nva);
                Q[q] = __ffma2_rn(h2[q], src[q], __fmul2_rn(d2[q], nx));
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[q] = __fmul2_rn(Q[q], u2[q]);
            finish(Q, dst[0].x);
        } else {
            // p[j] = u[j] (h[j] v[j] + d[j] v[j-1]);  pair (j, j+4) needs (v[j-1], v[j+3]) = pair j-1; j = 0: (inflow, v[3])
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 pv = (q > 0) ? src[(q + 3) & 3] : make_float2(nva, src[3].x);
                dst[q] = __fmul2_rn(u2[q], __ffma2_rn(h2[q], src[q], __fmul2_rn(d2[q], pv)));
            }
            finish(dst, dst[3].y);
        }
    };
    auto copy_back = [&]() {
#pragma unroll
        for (int q = 0; q < 4; ++q) X[q] = Y[q];
    };

    if (Tn > 0) {
        // ---- frame 0 (into X) ----
        ring_sync(0);
        {
            float2 u2[4], Q[4];
            load_u(u2);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (BETA) { Q[q].x = (s0 + q == Sn - 1) ? 1.f : 0.f; Q[q].y = (s0 + q + 4 == Sn - 1) ? 1.f : 0.f; }
                else { Q[q].x = (s0 + q == 0) ? u2[q].x : 0.f; Q[q].y = 0.f; }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) X[q] = BETA ? __fmul2_rn(Q[q], u2[q]) : Q[q];
            finish(Q, BETA ? X[0].x : X[3].y);
        }
        // ---- frame 1 alone, then pairs of frames (X -> Y -> X); group boundaries fall on even frames ----
        int k = 1;
        if (k < Tn) { step(X, Y); copy_back(); ++k; }
        for (; k + 1 < Tn; k += 2) {
            if ((k & (FB - 1)) == 0) ring_sync(k);
            step(X, Y);
            step(Y, X);
        }
        if (k < Tn) {
            if ((k & (FB - 1)) == 0) ring_sync(k);
            step(X, Y);
            copy_back();
        }
    }
    // what the score read-out below looks at
    float2 P[4], Wf[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { P[q] = X[q]; Wf[q] = X[q]; }
    if (lane == 0) st_volatile_s32(&sm.consumed[0], Tn);
    // ---- score ------------------------------------------------------------------------------------
    const double msum = (Tn > 0) ? wait_msum(sm.msum, &sm.flags[1]) : 0.0;
    if (lane == 0) {
        double score = 0.0;
        if (Tn > 0) {
            const T fin = BETA ? Wf[0].x : T(0);
            score = (double)Num<T>::log_(fin) + msum + (double)(Tn - 1) * (double)Rb + (double)Eme * 0.6931471805599453094;
            if (BETA && !(fin > T(0)) && Sn <= Tn) atomicOr(a.ctx.flags(n), RF_SCORE_LOST);
        }
        if (BETA) {
            a.ctx.scalars[(int64_t)n * 8 + SC_AL_BETA] = score;
            a.ctx.scalars[(int64_t)n * 8 + SC_RB] = (double)Rb;
            if (a.aligned_scores) a.aligned_scores[n] = (T)score;
        }
    }
    if (!BETA && Tn > 0) {      // alpha-side score (cross-check): mass at (T_n-1, S_n-1)
        const int sl = Sn - 1;
        if (sl >= s0 && sl < s0 + ASG_BLK) {
            T fin = T(0);
#pragma unroll
            for (int q = 0; q < 4; ++q) { if (s0 + q == sl) fin = P[q].x; if (s0 + q + 4 == sl) fin = P[q].y; }
            a.ctx.scalars[(int64_t)n * 8 + SC_AL_ALPHA] =
                (double)Num<T>::log_(fin) + msum + (double)(Tn - 1) * (double)Rb + (double)Eme * 0.6931471805599453094;
        }
    }
}

// dispatch: float -> packed specialisation, double -> generic
template <typename T, bool BETA> struct FacConsumer {
    __device__ static void run(const ScanArgs<T> &a, FacSmemView<T> &sm, int W, int n, int Tn, int Sn, T Rb, const T (&h)[ASG_BLK],
                               const T (&dd)[ASG_BLK], const int (&lab)[ASG_BLK], int warp, int lane) {
        fac_consumer<T, BETA>(a, sm, W, n, Tn, Sn, Rb, h, dd, lab, warp, lane);
    }
};
template <bool BETA> struct FacConsumer<float, BETA> {
    __device__ static void run(const ScanArgs<float> &a, FacSmemView<float> &sm, int W, int n, int Tn, int Sn, float Rb,
                               const float (&h)[ASG_BLK], const float (&dd)[ASG_BLK], const int (&lab)[ASG_BLK], int warp, int lane) {
        if (W == 1 && sm.pring) fac_consumer_f32<BETA, true>(a, sm, n, Tn, Sn, Rb, h, dd, lab, lane);
        else if (W == 1) fac_consumer_f32<BETA, false>(a, sm, n, Tn, Sn, Rb, h, dd, lab, lane);
        else fac_consumer<float, BETA>(a, sm, W, n, Tn, Sn, Rb, h, dd, lab, warp, lane);
    }
};

// ---------------------------------------------------------------------------------------------
// FAC, one float consumer warp: a third warp lays every published u row out BY TARGET POSITION,
//   pring[slot][8b + {0,2,4,6,1,3,5,7}[j]] = u[y_{8b+j}],
// so the consumer reads its block of eight as two LDS.128 that land directly in its packed (j, j+4) register
// pairs, instead of eight scalar gathers with their address arithmetic on the serial chain.  The gather costs
// ~70 instructions per frame: too much to add to the producer warp (it would become the slowest of the three).
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ void fac_gatherer(FacSmemView<T> &sm, int up, int pp, int Tn, int lane) {
    unsigned psrc[8], pdst[8];                   // byte offsets inside a u row / a position row of this lane's 8 positions
    unsigned pvalid = 0;
    const unsigned ur0 = smem_u32(sm.uring), pr0 = smem_u32(sm.pring);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int sx = lane + 32 * i, j = sx & 7;
        psrc[i] = pdst[i] = 0;
        if (sx < pp) {
            pvalid |= 1u << i;
            psrc[i] = (unsigned)sizeof(T) * (unsigned)(sm.ylab[sx] >= 0 ? sm.ylab[sx] : up - 1);
            pdst[i] = (unsigned)sizeof(T) * (unsigned)((sx & ~7) + ((j & 3) << 1) + (j >> 2));
        }
    }
    const int nb = (Tn + FB - 1) / FB;
    for (int b = 0; b < nb; ++b) {
        const int kb = b * FB, hi = min(kb + FB, Tn), nf = hi - kb;
        // the u rows of the batch are published, and the position rows they overwrite have been consumed
        while (ld_volatile_s32(&sm.flags[0]) < hi || ld_volatile_s32(&sm.consumed[0]) < hi - RING) __nanosleep(20);
        __threadfence_block();
        const unsigned slot0 = (unsigned)(kb % RING);          // FB divides RING: a batch never wraps
        for (int ff = 0; ff < nf; ++ff) {
            const unsigned ua = ur0 + (slot0 + ff) * (unsigned)(up * sizeof(T));
            const unsigned pa = pr0 + (slot0 + ff) * (unsigned)(pp * sizeof(T));
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = lds_f32(ua + psrc[i]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (pvalid & (1u << i)) sts_f32(pa + pdst[i], v[i]);
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            st_volatile_s32(&sm.flags[2], hi);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(544) fac_scan_kernel(const ScanArgs<T> a, int W) {
    extern __shared__ __align__(16) char smem_raw[];
    FacSmemView<T> sm;
    const int L = a.d.L, S = a.d.S;
    const int sp = (int)a.ctx.sp;
    const int lpi = a.fmax ? sp : (int)a.ctx.lpi;     // entries per ring row: all labels, or (large alphabets) the target positions
    const int up = lpi + 4;
    const bool posring = W == 1 && sizeof(T) == 4 && a.fmax == nullptr;
    fac_smem_carve<T>(&sm, smem_raw, lpi, sp, W, posring);
    const int n = blockIdx.x;
    const int dir = blockIdx.y == 0 ? a.dir0 : a.dir1;
    // roles by warp index: [0, W) consumers, W producer, W+1 position-ring warp (rotating the roles so that the FCC and
    // FAC consumers of two CTAs sharing an SM sit in different warp slots was measured: 241-246 us vs 222 us as is)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nthreads = 32 * W + 32 + (posring ? 32 : 0);

    int Tn = a.il ? (int)a.il[n] : a.d.T;
    Tn = max(0, min(Tn, a.d.T));
    int Sn = a.tl ? (int)a.tl[n] : S;
    Sn = max(1, min(Sn, S));

    if (threadIdx.x == 0) { sm.flags[0] = 0; sm.flags[1] = 0; sm.flags[2] = 0; }
    if (threadIdx.x < W) sm.consumed[threadIdx.x] = 0;
    for (int l = threadIdx.x; l < RING * up; l += nthreads) sm.uring[l] = T(0);     // incl. the always-zero slots of every row
    const int64_t *yn = a.y + (int64_t)n * a.y_s0;
    bool bad_label = false;
    for (int s = threadIdx.x; s < sp; s += nthreads) {
        int lab = -1;
        if (s < Sn) {
            const int64_t raw = yn[(int64_t)s * a.y_s1];
            bad_label |= raw < 0 || raw >= L;
            lab = (int)max((int64_t)0, min((int64_t)(L - 1), raw));
        }
        sm.ylab[s] = lab;
    }
    if (bad_label) atomicOr(a.ctx.flags(n), RF_LABEL_RANGE);       // rare; reported by asg_range_flags
    __syncthreads();

    // ---- per-position transition scores: H_s = A[y_s,y_s], D_s = A[y_{s+1},y_s]; Rb = their max ----
    const bool consumer = warp < W;
    const int gl = warp * 32 + lane;
    const int s0 = gl * ASG_BLK;
    const T ninf = Num<T>::neg_inf();
    T Hs[ASG_BLK], Ds[ASG_BLK];   // log-domain; Ds: edge s -> s+1
    int lab[ASG_BLK];             // label of each own position (lpi = the always-zero slot of a u row)
    T lmax = ninf;
    if (consumer) {
#pragma unroll
        for (int j = 0; j < ASG_BLK; ++j) {
            const int s = s0 + j;
            Hs[j] = ninf; Ds[j] = ninf; lab[j] = lpi;         // lpi = the always-zero slot of a ring row
            if (s < Sn) {
                const int c = sm.ylab[s];
                lab[j] = a.fmax ? s : c;                       // gather mode: the ring row is indexed by position
                Hs[j] = a.A[c * a.A_s0 + c * a.A_s1];
                if (s + 1 < Sn) { const int nx = sm.ylab[s + 1]; Ds[j] = a.A[nx * a.A_s0 + c * a.A_s1]; }
                if (Hs[j] < -ninf) lmax = tmax(lmax, Hs[j]);
                if (Ds[j] < -ninf) lmax = tmax(lmax, Ds[j]);
            }
        }
        lmax = Num<T>::warp_max(lmax);
        if (lane == 0) sm.red[warp] = lmax;
    }
    __syncthreads();
    T Rb = ninf;
    for (int q = 0; q < W; ++q) Rb = tmax(Rb, sm.red[q]);
    if (!(Rb > ninf)) Rb = T(0);

    if (warp == W + 1) {                 // position-ring warp (only launched with posring)
        fac_gatherer<T>(sm, up, sp, Tn, lane);
        return;
    }
    if (!consumer) {
        auto cons = [&]() {
            int c = 0x7fffffff;
            for (int q = 0; q < W; ++q) c = min(c, ld_volatile_s32(&sm.consumed[q]));
            return c;
        };
        if (dir == 0) scan_producer<T, false>(a, sm.raw, up, sm.uring, up, &sm.flags[0], sm.msum, &sm.flags[1], cons, [](int) {}, 0, n, Tn, lane, a.fmax ? sm.ylab : nullptr, sp);
        else scan_producer<T, true>(a, sm.raw, up, sm.uring, up, &sm.flags[0], sm.msum, &sm.flags[1], cons, [](int) {}, 0, n, Tn, lane, a.fmax ? sm.ylab : nullptr, sp);
        return;
    }
    T h[ASG_BLK], dout[ASG_BLK];
#pragma unroll
    for (int j = 0; j < ASG_BLK; ++j) {
        h[j] = (Hs[j] == ninf) ? T(0) : Num<T>::exp_(Hs[j] - Rb);
        dout[j] = (Ds[j] == ninf) ? T(0) : Num<T>::exp_(Ds[j] - Rb);
    }
    if (dir == 0) {
        // publish h, d for the gradient assembly, and the lengths
        if (s0 < sp) {
#pragma unroll
            for (int j = 0; j < ASG_BLK; ++j) {
                a.ctx.fac_hd[(int64_t)n * sp + s0 + j] = h[j];
                a.ctx.fac_hd[((int64_t)a.d.N + n) * sp + s0 + j] = dout[j];
            }
        }
        if (gl == 0) { a.ctx.il[n] = Tn; a.ctx.tl[n] = Sn; }
        // alpha needs the edge INTO each own position: d_{s-1}
        T din[ASG_BLK];
        T left = __shfl_up_sync(ASG_FULL_MASK, dout[ASG_BLK - 1], 1);
        if (lane == 0) {
            // previous warp's last lane: recompute from A directly
            left = T(0);
            if (s0 > 0 && s0 < Sn) {
                const int c = sm.ylab[s0 - 1], nx = sm.ylab[s0];
                const T dv = a.A[nx * a.A_s0 + c * a.A_s1];
                left = (dv == ninf) ? T(0) : Num<T>::exp_(dv - Rb);
            }
        }
#pragma unroll
        for (int j = 0; j < ASG_BLK; ++j) din[j] = (j == 0) ? left : dout[(j + ASG_BLK - 1) % ASG_BLK];
        FacConsumer<T, false>::run(a, sm, W, n, Tn, Sn, Rb, h, din, lab, warp, lane);
    } else {
        FacConsumer<T, true>::run(a, sm, W, n, Tn, Sn, Rb, h, dout, lab, warp, lane);
    }
}

// =============================================================================================
// launchers
// =============================================================================================
// The FAC scans run beside the FCC scans on a side stream.  Every (device, caller stream) pair owns its side
// stream and its fork/join events, so independent callers (several host contexts in flight, several host
// threads) neither serialise on one shared side stream nor re-record each other's events.
struct SideSlot { int device; cudaStream_t key, side; cudaEvent_t fork, join; };
static std::mutex g_side_mu;
static std::vector<SideSlot> g_side_slots;
void release_side_slots() {
    std::lock_guard<std::mutex> lock(g_side_mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (const SideSlot &s : g_side_slots) {
        cudaSetDevice(s.device);
        cudaEventDestroy(s.fork);
        cudaEventDestroy(s.join);
        cudaStreamDestroy(s.side);
    }
    g_side_slots.clear();
    cudaSetDevice(cur);
}
static bool side_slot(int device, cudaStream_t key, SideSlot &out) {
    std::mutex &mu = g_side_mu;
    std::vector<SideSlot> &slots = g_side_slots;
    std::lock_guard<std::mutex> lock(mu);
    for (const SideSlot &s : slots)
        if (s.device == device && s.key == key) { out = s; return true; }
    if (slots.size() >= 256) return false;          // beyond that: the caller's stream runs both (still correct)
    SideSlot s{device, key, nullptr, nullptr, nullptr};
    if (cudaStreamCreateWithFlags(&s.side, cudaStreamNonBlocking) != cudaSuccess) return false;
    if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) return false;
    slots.push_back(s);
    out = s;
    return true;
}

template <typename T, int LP>
static cudaError_t launch_fcc_lp(const ScanArgs<T> &a, int ndirs, cudaStream_t st) {
    dim3 grid(a.d.N, ndirs);
    {
        const size_t bytes = sizeof(FccSmem<T, LP>);
        cudaError_t e = cudaFuncSetAttribute(fcc_scan_kernel<T, LP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        // the FCC and FAC scan kernels run side by side on the same SMs: both ask for the same L1/shared split, or the
        // second one can only use SMs the first has left empty (an SM changes its carve-out only when idle)
        cudaFuncSetAttribute(fcc_scan_kernel<T, LP>, cudaFuncAttributePreferredSharedMemoryCarveout, SCAN_CARVEOUT_PCT);
        LaunchScope ls(K_FCC_SCAN, st);
        fcc_scan_kernel<T, LP><<<grid, 64, bytes, st>>>(a);
    }
    return cudaGetLastError();
}

template <typename T> static cudaError_t launch_fcc(const ScanArgs<T> &a, int ndirs, cudaStream_t st) {
    const int L = a.d.L;
    if (L <= 8) return launch_fcc_lp<T, 8>(a, ndirs, st);
    if (L <= 16) return launch_fcc_lp<T, 16>(a, ndirs, st);
    if (L <= 24) return launch_fcc_lp<T, 24>(a, ndirs, st);
    if (L <= 32) return launch_fcc_lp<T, 32>(a, ndirs, st);
    if (L <= 40) return launch_fcc_lp<T, 40>(a, ndirs, st);
    if (L <= 48) return launch_fcc_lp<T, 48>(a, ndirs, st);
    if (L <= 56) return launch_fcc_lp<T, 56>(a, ndirs, st);
    if (L <= 64) return launch_fcc_lp<T, 64>(a, ndirs, st);
    return cudaErrorInvalidValue;
}

template <typename T> static cudaError_t launch_fac(const ScanArgs<T> &a, int ndirs, cudaStream_t st) {
    const int lanes = (int)a.ctx.nblk;
    int W = 1;
    while (W * 32 < lanes) W <<= 1;
    if (W > 16) return cudaErrorInvalidValue;
    const size_t bytes = fac_smem_carve<T>(nullptr, nullptr, a.fmax ? (int)a.ctx.sp : (int)a.ctx.lpi, (int)a.ctx.sp, W,
                                           W == 1 && sizeof(T) == 4 && a.fmax == nullptr);
    if (bytes > 227 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(fac_scan_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    dim3 grid(a.d.N, ndirs);
    {
        cudaFuncSetAttribute(fac_scan_kernel<T>, cudaFuncAttributePreferredSharedMemoryCarveout, SCAN_CARVEOUT_PCT);
        LaunchScope ls(K_FAC_SCAN, st);
        const bool posring = W == 1 && sizeof(T) == 4 && a.fmax == nullptr;
        fac_scan_kernel<T><<<grid, 32 * W + 32 + (posring ? 32 : 0), bytes, st>>>(a, W);
    }
    return cudaGetLastError();
}

template <typename T> int launch_scan(const ScanArgs<T> &a_in, int roles, cudaStream_t stream) {
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return (int)e;
    const bool fcc_a = roles & ASG_ROLE_FCC_ALPHA, fcc_b = roles & ASG_ROLE_FCC_BETA;
    const bool fac_a = roles & ASG_ROLE_FAC_ALPHA, fac_b = roles & ASG_ROLE_FAC_BETA;
    const bool any_fcc = fcc_a || fcc_b, any_fac = fac_a || fac_b;
    if (a_in.d.L > ASG_SMALL_L) {      // E, E^T, max A and the per-frame emission maxima, before the streams fork
        int rc = large_prepare<T>(a_in, stream);
        if (rc) return rc;
    }
    // range flags of this call start clear (the scalars block is rewritten by every scan anyway)
    if ((e = cudaMemsetAsync(a_in.ctx.scalars, 0, sizeof(double) * 8 * (size_t)a_in.d.N, stream)) != cudaSuccess) return (int)e;
    cudaStream_t side = nullptr;
    SideSlot slot{};
    if (any_fcc && any_fac && side_slot(device, stream, slot)) {
        // fork the side stream off the caller's stream, join after
        side = slot.side;
        if ((e = cudaEventRecord(slot.fork, stream)) != cudaSuccess) return (int)e;
        if ((e = cudaStreamWaitEvent(side, slot.fork, 0)) != cudaSuccess) return (int)e;
    }
    const bool large = a_in.d.L > ASG_SMALL_L;
    if (any_fcc) {
        ScanArgs<T> a = a_in;
        a.dir0 = fcc_a ? 0 : 1;
        a.dir1 = 1;
        const int nd = (fcc_a && fcc_b) ? 2 : 1;
        if (large) { int rc = launch_fcc_large<T>(a, nd, stream); if (rc) return rc; }
        else if ((e = launch_fcc<T>(a, nd, stream)) != cudaSuccess) return (int)e;
    }
    if (any_fac) {
        ScanArgs<T> a = a_in;
        a.dir0 = fac_a ? 0 : 1;
        a.dir1 = 1;
        if (large) a.fmax = a.ctx.fmax;        // gather mode
        cudaStream_t st = side ? side : stream;
        if ((e = launch_fac<T>(a, (fac_a && fac_b) ? 2 : 1, st)) != cudaSuccess) return (int)e;
        if (side) {
            if ((e = cudaEventRecord(slot.join, side)) != cudaSuccess) return (int)e;
            if ((e = cudaStreamWaitEvent(stream, slot.join, 0)) != cudaSuccess) return (int)e;
        }
    }
    return 0;
}

template int launch_scan<float>(const ScanArgs<float> &, int, cudaStream_t);
template int launch_scan<double>(const ScanArgs<double> &, int, cudaStream_t);

}  // namespace asg
