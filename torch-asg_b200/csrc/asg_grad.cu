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
// alpha[t-1] and u*beta[t] on the fly.
//
// Structure.  tile = (utterance n, TB consecutive frames).  State is utterance-major, so a tile of every state
// array is ONE contiguous chunk: the producer fetches it with one bulk (TMA) copy (cp.async.bulk -> UBLKCP)
// into an NS-stage shared-memory ring guarded by mbarriers.  Each of the NW consumer warps then owns whole
// frames of the tile -- one warp per (utterance, frame) slice, lane = one 8-position block of the target /
// one or two labels -- and does everything for its frame in registers with no block-wide barrier:
//   FAC posterior (block-floating-point alignment, one REDUX + one shuffle tree), label scatter into the
//   warp's private fixed-point bins (integer shared-memory atomics: order-independent, so deterministic),
//   edge posteriors into per-lane register accumulators, FCC posterior, the FCC transition-gradient outer
//   product into a per-lane register slice of G, and the one coalesced store of the grad_x row.
// Each warp has two frames in flight (their dependency chains interleave in one instruction stream).  Warps only
// meet at utterance boundaries (the per-position edge sums go, as 64-bit fixed point, into the CTA's L x L edge
// matrix) and at the end (G slices parked in the idle tile ring, added in warp order).  Three helper warps in the
// producer's warpgroup stage the next utterance's labels / h / d / scalars in shared memory one utterance ahead.
// Every reduction has a fixed order or is integer: results are run-to-run deterministic.
#include <climits>
#include <cstdlib>
#include "asg_common.cuh"
#include "asg_internal.h"

namespace asg {

constexpr int GRAD_MAX_CWARPS = 12;    // consumer warps (+ the producer's warpgroup)
constexpr int GRAD_LBINS = 72;         // label bins per warp: L <= 64, bin L = trash (padding positions)

// 64-bit add into shared memory as two native 32-bit atomics (a 64-bit shared atomicAdd compiles to a compare-and-swap
// spin loop).  Exact: the thread whose add wraps the low word is the one -- and the only one -- that carries into the high
// word, so the final 64-bit value is the modular sum in any order.  Readers must be separated by a barrier.
__device__ __forceinline__ void add64_shared(unsigned long long *addr, unsigned long long v) {
    unsigned *w = reinterpret_cast<unsigned *>(addr);
    const unsigned vl = (unsigned)v, vh = (unsigned)(v >> 32);
    const unsigned old = atomicAdd(w, vl);
    const unsigned hi = vh + ((unsigned)(old + vl) < old ? 1u : 0u);
    if (hi) atomicAdd(w + 1, hi);
}

// Fixed-point images (integers add in any order with the same result, so shared-memory atomics stay deterministic):
//   posterior p of one position in one frame, 0 <= p <= 1, a frame sums to 1  -> p * kPost  (int32 / int64 bins)
//   edge sums: posteriors (already in kPost units) summed over <= 2^24 frames -> 64-bit, kEdge units per unit mass
template <typename T> struct Fix;
template <> struct Fix<float> {
    typedef int bin_t;
    static constexpr float kPost = 1073741824.0f;                 // 2^30
    static constexpr float kInvPost = 1.0f / 1073741824.0f;
    __device__ __forceinline__ static int to_fix(float q) { return __float2int_rn(q); }
    __device__ __forceinline__ static float from_fix(int q) { return (float)q; }
    __device__ __forceinline__ static void add(int *bin, int q) { atomicAdd(bin, q); }
    // accumulators hold kPost * mass (mass <= 2^24): 2^32 units per unit mass
    __device__ __forceinline__ static unsigned long long edge_fix(float acc) { return (unsigned long long)__float2ll_rn(acc * 4.0f); }
    __device__ __forceinline__ static float edge_unfix(unsigned long long q) { return (float)((double)(long long)q * (1.0 / 4294967296.0)); }
};
template <> struct Fix<double> {
    typedef unsigned long long bin_t;
    static constexpr double kPost = 4611686018427387904.0;       // 2^62
    static constexpr double kInvPost = 1.0 / 4611686018427387904.0;
    __device__ __forceinline__ static unsigned long long to_fix(double q) { return (unsigned long long)__double2ll_rn(q); }
    __device__ __forceinline__ static double from_fix(unsigned long long q) { return (double)(long long)q; }
    __device__ __forceinline__ static void add(unsigned long long *bin, unsigned long long q) { add64_shared(bin, q); }
    // 2^38 units per unit mass
    __device__ __forceinline__ static unsigned long long edge_fix(double acc) { return (unsigned long long)__double2ll_rn(acc * (1.0 / 16777216.0)); }
    __device__ __forceinline__ static double edge_unfix(unsigned long long q) { return (double)(long long)q * (1.0 / 274877906944.0); }
};

// pairs of accumulators: float pairs use the packed two-FMA instruction of sm_100 (FFMA2), double pairs are two FMAs
template <typename T> struct Pk;
template <> struct Pk<float> {
    typedef float2 v2;
    __device__ __forceinline__ static v2 mk(float a, float b) { return make_float2(a, b); }
    __device__ __forceinline__ static v2 fma2(v2 a, v2 b, v2 c) { return __ffma2_rn(a, b, c); }
};
template <> struct Pk<double> {
    typedef double2 v2;
    __device__ __forceinline__ static v2 mk(double a, double b) { return make_double2(a, b); }
    __device__ __forceinline__ static v2 fma2(v2 a, v2 b, v2 c) { return make_double2(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y)); }
};

// approximate (float) / exact (double) quotient for the edge posteriors
// float: ONE MUFU.RCP (1 ulp; __fdividef adds a range-scaling sequence and __frcp_rn a refinement with a slow-path branch, both
// of which the callers' operand ranges make unnecessary: every divisor is floored at Num::tiny() and no quotient exceeds 2^100)
__device__ __forceinline__ float rcp_(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ double rcp_(double x) { return 1.0 / x; }
__device__ __forceinline__ float fast_div(float a, float b) { return a * rcp_(b); }
__device__ __forceinline__ double fast_div(double a, double b) { return a / b; }
// posterior bins of the fast path: a shared-memory reduction at a precomputed 32-bit address (no return value, no per-element
// address arithmetic); float only -- the double bins are 64-bit (add64_shared)
// eight of them under ONE predicate (idle lanes stay out), at base addresses + a compile-time offset
template <int OFF>
__device__ __forceinline__ void red8_shared_s32_if(const unsigned (&a)[8], const int (&v)[8], bool pred) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %16, 0;\n"
                 "@p red.shared.add.s32 [%0+%17], %8;\n@p red.shared.add.s32 [%1+%17], %9;\n@p red.shared.add.s32 [%2+%17], %10;\n"
                 "@p red.shared.add.s32 [%3+%17], %11;\n@p red.shared.add.s32 [%4+%17], %12;\n@p red.shared.add.s32 [%5+%17], %13;\n"
                 "@p red.shared.add.s32 [%6+%17], %14;\n@p red.shared.add.s32 [%7+%17], %15;\n}\n"
                 ::"r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
                   "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"((unsigned)pred), "n"(OFF) : "memory");
}

// shared-memory map: [per-CTA block][stage 0][stage 1]...   A stage holds TB+1 frame records (the halo record of
// frame t0-1 first) exactly as they lie in global memory, plus the tile's winv scalars.
struct GradMap {
    size_t oRec, oWinv, stage_bytes;                                   // per stage, byte offsets from the stage base
    size_t o_hh, o_din, o_ylab, o_cscal, o_Gal, o_Gf, o_galfix, o_bins, o_acc, o_bars, o_stages, total;   // per CTA
};

__host__ __device__ inline GradMap grad_map(size_t eb, size_t rec_bytes, int L, int sp, int TB, int nstage, int nw, bool fcc, bool fac) {
    GradMap m;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 16); return o; };
    m.oRec = take(rec_bytes * (size_t)(TB + 1));
    m.oWinv = take(fcc ? eb * (size_t)((TB + 3) / 4 * 4) : 0);
    m.stage_bytes = align_up(off, 128);
    off = 0;
    const size_t nb = sp <= 256 ? 2 : 1;         // narrow mode: the utterance constants are double-buffered (prefetch warp)
    m.o_hh = take(fac ? nb * eb * sp : 0);
    m.o_din = take(fac ? nb * eb * sp : 0);
    m.o_ylab = take(fac ? nb * sizeof(int) * sp : 0);
    m.o_cscal = take(2 * 2 * eb + 2 * 2 * sizeof(int));      // [2]{gf, ga} then [2]{Tn, Sn}
    m.o_Gal = take(eb * (size_t)L * L);
    m.o_Gf = take(fcc ? eb * (size_t)L * L : 0);
    m.o_galfix = take(fac ? 8 * (size_t)L * L : 0);
    m.o_bins = take(fac ? eb * (size_t)nw * 2 * GRAD_LBINS : 0);      // int32 bins for float, int64 for double
    m.o_acc = take((fac && sp > 256) ? eb * (size_t)nw * 2 * sp : 0);
    m.o_bars = take(sizeof(uint64_t) * 3 * 8);
    m.o_stages = align_up(off, 128);
    m.total = m.o_stages + m.stage_bytes * nstage;
    return m;
}

// 8 consecutive elements (32-byte aligned for float) from shared memory
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

// Eight consecutive elements per lane at a 32-byte lane pitch make every 128-bit shared-memory load a 2-way bank
// conflict (lanes l and l+4 of a quarter-warp hit the same banks).  Lanes with bit 2 set therefore fetch their two halves
// in the opposite order (sw = 4): the load is conflict-free, and such a lane simply keeps its block PERMUTED in
// registers -- register r holds position r ^ 4.  Everything per position is loaded the same way (states, labels, h, d),
// so the element-wise arithmetic never notices; only the three places that need a neighbour do (prev8, last8).
__device__ __forceinline__ void lds8sw(const float *p, int sw, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4 *>(p + sw);
    const float4 b = *reinterpret_cast<const float4 *>(p + (4 - sw));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void lds8sw(const double *p, int sw, double (&v)[8]) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const double2 a = *(reinterpret_cast<const double2 *>(p + sw) + q);
        const double2 b = *(reinterpret_cast<const double2 *>(p + (4 - sw)) + q);
        v[2 * q] = a.x; v[2 * q + 1] = a.y; v[4 + 2 * q] = b.x; v[4 + 2 * q + 1] = b.y;
    }
}
// The same loads by 32-bit shared address (fast path: one integer add per load instead of 64-bit generic-pointer arithmetic).
// a_lo / a_hi: the addresses of the block's two 16-byte halves, already swapped for permuted lanes.
__device__ __forceinline__ void lds8_at(unsigned a_lo, unsigned a_hi, float (&v)[8]) {
    const float4 a = lds_v4f32(a_lo), b = lds_v4f32(a_hi);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void lds8_at(unsigned a_lo, unsigned a_hi, double (&v)[8]) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[2 * q]), "=d"(v[2 * q + 1]) : "r"(a_lo + 16u * q));
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[4 + 2 * q]), "=d"(v[4 + 2 * q + 1]) : "r"(a_hi + 16u * q));
    }
}
__device__ __forceinline__ float lds_at(unsigned a, float) { return lds_f32(a); }
__device__ __forceinline__ double lds_at(unsigned a, double) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int lds_s32_at(unsigned a) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 lds_v2_at(unsigned a, float) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ double2 lds_v2_at(unsigned a, double) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
// value at the position before the one register j holds (`left` = the previous lane's last position)
template <typename V> __device__ __forceinline__ V prev8(const V (&v)[8], int j, bool fl, V left) {
    if (j == 0) return fl ? v[7] : left;
    if (j == 4) return fl ? left : v[3];
    return v[j - 1];
}
// value at the lane's last position (7)
template <typename V> __device__ __forceinline__ V last8(const V (&v)[8], bool fl) { return fl ? v[3] : v[7]; }

template <typename T, int LP, bool FCC, bool FAC>
__global__ void __launch_bounds__((GRAD_MAX_CWARPS + 4) * 32, 1) grad_kernel(const GradArgs<T> a) {
    extern __shared__ __align__(128) char smem_raw[];
    typedef typename Fix<T>::bin_t bin_t;
    const int L = a.d.L, N = a.d.N, Tt = a.d.T;
    const int sp = (int)a.ctx.sp, nblk = (int)a.ctx.nblk, nbp = (int)a.ctx.nbp, lpi = (int)a.ctx.lpi;
    const int64_t tp = a.ctx.tp;
    const int TB = a.tb, NS = a.nstage, NW = a.nwarps;
    const GradMap mp = grad_map(sizeof(T), (size_t)a.ctx.R, L, sp, TB, NS, NW, FCC, FAC);
    const int R = (int)a.ctx.R;                    // frame record size in bytes
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ncons = NW * 32;

    T *hh = (T *)(smem_raw + mp.o_hh), *din = (T *)(smem_raw + mp.o_din);
    T *Gal = (T *)(smem_raw + mp.o_Gal), *Gf = (T *)(smem_raw + mp.o_Gf);
    int *ylab = (int *)(smem_raw + mp.o_ylab);
    uint64_t *full = (uint64_t *)(smem_raw + mp.o_bars), *empty = full + 8, *cfull = full + 16, *cempty = full + 18;
    T *cscalT = (T *)(smem_raw + mp.o_cscal);
    int *cscalI = (int *)(cscalT + 4);
    char *stages = smem_raw + mp.o_stages;

    // balanced contiguous tile ranges
    const int q_begin = (int)(((int64_t)a.num_tiles * blockIdx.x) / gridDim.x);
    const int q_end = (int)(((int64_t)a.num_tiles * (blockIdx.x + 1)) / gridDim.x);

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NW); }
        for (int b = 0; b < 2; ++b) { mbar_init(&cfull[b], 96); mbar_init(&cempty[b], NW); }
        mbar_fence_init();
    }
    __syncthreads();

    // Register budget: the consumer warps fill whole warpgroups (idle warps of a partly used group just leave), the
    // producer sits in a warpgroup of its own and hands its registers back (setmaxnreg), so the frame warps keep all
    // their accumulators in registers: 16 warps launch with 128 registers each, the producer group keeps 32, the
    // twelve frame warps get 160.
    const int PW = (NW + 3) & ~3;                  // first warp of the producer's warpgroup
    // =========================== producer warp: ONE bulk copy per tile ================================
    // records t0-1 .. t0+rows-1 of the utterance are one contiguous chunk of the state array
    const bool prefetch = (nblk <= 32) || !FAC;      // narrow targets: a helper warp stages the next utterance's constants
    if (warp >= PW) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;\n");
        if (warp > PW && prefetch && q_begin < q_end) {
            // ---- utterance constants one utterance ahead, into the buffer the frame warps released: no global-memory
            //      latency at an utterance switch.  Three helper warps: labels | h | d (+ the scalars), every lane's loads
            //      issued together (sp <= 256: at most 8 per lane).
            const int role = warp - PW;             // 1: labels, 2: h, 3: d and scalars
            const int n_first = q_begin / a.tiles_t, n_last = (q_end - 1) / a.tiles_t;
            for (int n = n_first, k = 0; n <= n_last; ++n, ++k) {
                const int buf = k & 1;
                mbar_wait(&cempty[buf], ((k >> 1) & 1) ^ 1);
                const int Sn_ = FAC ? (int)a.ctx.tl[n] : 0;
                if (FAC && role == 1) {
                    const int64_t *yn = a.y + (int64_t)n * a.y_s0;
                    int64_t v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { const int s = lane + 32 * i; v[i] = (s < Sn_) ? yn[(int64_t)s * a.y_s1] : (int64_t)L; }
#pragma unroll
                    for (int i = 0; i < 8; ++i) { const int s = lane + 32 * i; if (s < sp) ylab[buf * sp + s] = (s < Sn_) ? max(0, min(L - 1, (int)v[i])) : L; }
                } else if (FAC) {
                    const T *src = a.ctx.fac_hd + (role == 2 ? (int64_t)n * sp : ((int64_t)N + n) * sp - 1);
                    T *dst = (role == 2 ? hh : din) + buf * sp;
                    const int lo = role == 2 ? 0 : 1;
                    T v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { const int s = lane + 32 * i; v[i] = (s >= lo && s < Sn_) ? src[s] : T(0); }
#pragma unroll
                    for (int i = 0; i < 8; ++i) { const int s = lane + 32 * i; if (s < sp) dst[s] = v[i]; }
                }
                if (role == 3 && lane == 0) {
                    cscalT[buf * 2 + 0] = FCC ? a.gfull(n) : T(0);
                    cscalT[buf * 2 + 1] = FAC ? a.galigned(n) : T(0);
                    cscalI[buf * 2 + 0] = (int)a.ctx.il[n];
                    cscalI[buf * 2 + 1] = Sn_;
                }
                mbar_arrive(&cfull[buf]);        // every lane of the three warps: its own stores are released
            }
        }
        if (warp == PW && lane == 0) {
            int n = q_begin / a.tiles_t, tt = q_begin - n * a.tiles_t;     // walked, no per-tile divisions
            int Tn = (q_begin < q_end) ? (int)a.ctx.il[n] : 0;
            int stage = 0, use = 0;
            for (int q = q_begin; q < q_end; ++q) {
                const int t0 = tt * TB;
                const int TBt = min(TB, Tt - t0);
                const int rows = max(0, min(TBt, Tn - t0));
                mbar_wait(&empty[stage], (use & 1) ^ 1);
                if (rows == 0) {
                    mbar_arrive(&full[stage]);
                } else {
                    char *sb = stages + (size_t)stage * mp.stage_bytes;
                    const int prev = t0 > 0 ? 1 : 0;                   // the halo record t0-1 exists
                    const int64_t rowa = (int64_t)n * tp + t0 - prev;    // first record fetched
                    const unsigned bytes = (unsigned)(rows + prev) * (unsigned)R;
                    mbar_arrive_expect_tx(&full[stage], bytes);
                    bulk_g2s(sb + mp.oRec + (size_t)(1 - prev) * R, a.ctx.record(rowa), bytes, &full[stage]);
                }
                if (++stage == NS) { stage = 0; ++use; }
                if (++tt == a.tiles_t) { tt = 0; ++n; if (q + 1 < q_end) Tn = (int)a.ctx.il[n]; }
            }
        }
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 160;\n");
    if (warp >= NW) return;

    // =========================== consumer warps =====================================================
    auto cbar = [&]() { named_bar_sync(1, ncons); };
    const bool narrow = nblk <= 32;                // one 8-position block per lane: everything stays in registers
    const bool fl = (lane >> 2) & 1;               // this lane keeps its 8-position block permuted (see lds8sw)
    const int sw = fl ? 4 : 0;
    typedef Fix<T> FX;

    // FCC: per-lane register slice of this warp's G: rows ib*RI.., columns jb*CJ..
    constexpr int RI = LP / 8, CJ = LP / 4;
    const int ib = lane >> 2, jb = lane & 3;
    typedef typename Pk<T>::v2 V2;
    constexpr int CJ2 = CJ / 2;                    // LP is a multiple of 8: CJ is even
    V2 G[RI][CJ2];                                 // columns (2c, 2c+1) of the slice travel as one pair
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
        for (int j = 0; j < CJ2; ++j) G[i][j] = Pk<T>::mk(T(0), T(0));

    unsigned long long *galfix = (unsigned long long *)(smem_raw + mp.o_galfix);     // [L*L] fixed-point edge sums
    for (int e = tid; e < L * L; e += ncons) { Gal[e] = T(0); if (FCC) Gf[e] = T(0); if (FAC) galfix[e] = 0; }

    bin_t *bins = (bin_t *)(smem_raw + mp.o_bins) + (size_t)warp * 2 * GRAD_LBINS;     // two sets: one per frame in flight
    const unsigned bins_a = smem_u32(bins);
    T *acc = (T *)(smem_raw + mp.o_acc) + (size_t)warp * 2 * sp;     // wide mode only: [2][sp] stay / total sums of this warp
    if (FAC) {
        for (int l = lane; l < 2 * GRAD_LBINS; l += 32) bins[l] = 0;
        if (!narrow)
            for (int s = lane; s < 2 * sp; s += 32) acc[s] = T(0);
    }
    cbar();

    int cur_n = -1, nutt = 0, useq = 0;
    int Tn = 0, Sn = 0;
    T gf = T(0), ga = T(0);
    // narrow mode: this lane's block of the current utterance.  accH: posterior mass that STAYED into position j,
    // accQ: all posterior mass at position j (frames >= 1); the moved mass is their difference.
    T hreg[8], dreg[8], accH[8], accQ[8];
    // lab[j]: the shared-memory ADDRESS of the posterior bin of position j's label (set 0 of this warp's bins): the fast path adds
    // into it without any per-frame address arithmetic; LAB() recovers the label where it is needed (utterance boundaries)
    unsigned lab[8], labm1 = 0;
    auto LAB = [&](unsigned v) { return (int)((v - bins_a) / (unsigned)sizeof(bin_t)); };
    auto LADDR = [&](int l) { return bins_a + (unsigned)l * (unsigned)sizeof(bin_t); };
#pragma unroll
    for (int j = 0; j < 8; ++j) { hreg[j] = dreg[j] = accH[j] = accQ[j] = T(0); lab[j] = LADDR(L); }

    // Utterance boundary: the per-position edge sums go, as 64-bit fixed point (order-independent), into the CTA's
    // L x L edge matrix of this utterance; after one barrier they are scaled by the utterance's upstream gradient and
    // added to Gal by the thread that owns the element.  A second barrier after the owners have zeroed the matrix lets ONE matrix serve every boundary.
    auto flush_fac = [&]() {
        if (!FAC || cur_n < 0) return;
        unsigned long long *gfix = galfix;
        if (narrow) {
            if (lane < nblk) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int lj = LAB(lab[j]);
                    if (lj < L) {
                        const int prevlab = LAB(prev8(lab, j, fl, labm1));
                        add64_shared(&gfix[lj * L + lj], FX::edge_fix(accH[j]));
                        if (lane * 8 + (j ^ sw) >= 1) add64_shared(&gfix[lj * L + prevlab], FX::edge_fix(accQ[j] - accH[j]));
                    }
                    accH[j] = accQ[j] = T(0);
                }
            }
        } else {
            for (int s = lane; s < Sn; s += 32) {
                const int l = ylab[s];
                add64_shared(&gfix[l * L + l], FX::edge_fix(acc[s]));
                if (s >= 1) add64_shared(&gfix[l * L + ylab[s - 1]], FX::edge_fix(acc[sp + s] - acc[s]));
                acc[s] = T(0); acc[sp + s] = T(0);
            }
        }
        cbar();
        for (int e = tid; e < L * L; e += ncons) {
            Gal[e] += ga * FX::edge_unfix(gfix[e]);
            gfix[e] = 0;
        }
        cbar();                                 // one matrix (not two alternating ones: its 12.8 kB are worth more as L1), so the
        ++nutt;                                 // zeroing must be complete before any warp adds the next utterance's sums
    };

    const int64_t gx_pair_stride = (int64_t)NW * N * L;
    const bool has0 = lane < lpi, has1 = LP > 32 && lane + 32 < lpi;           // this lane's two FCC labels exist
    // (utterance, tile-in-utterance, ring stage, ring pass) are walked: no per-tile divisions
    int wn = q_begin / a.tiles_t, wt = q_begin - wn * a.tiles_t, stage = 0, use = 0;
    for (int q = q_begin; q < q_end; ++q) {
        const int n = wn;
        const int t0 = wt * TB;
        const int TBt = min(TB, Tt - t0);         // frames in this tile
        if (n != cur_n) {
            flush_fac();
            cur_n = n;
            if (prefetch) {
                const int buf = useq & 1;
                mbar_wait(&cfull[buf], (useq >> 1) & 1);
                gf = cscalT[buf * 2 + 0]; ga = cscalT[buf * 2 + 1];
                Tn = cscalI[buf * 2 + 0]; Sn = cscalI[buf * 2 + 1];
                if (FAC) {
                    const int o = buf * sp + min(lane, nblk - 1) * 8;
                    const int4 l0 = *reinterpret_cast<const int4 *>(ylab + o + sw), l1 = *reinterpret_cast<const int4 *>(ylab + o + 4 - sw);
                    lab[0] = LADDR(l0.x); lab[1] = LADDR(l0.y); lab[2] = LADDR(l0.z); lab[3] = LADDR(l0.w);
                    lab[4] = LADDR(l1.x); lab[5] = LADDR(l1.y); lab[6] = LADDR(l1.z); lab[7] = LADDR(l1.w);
                    lds8sw(hh + o, sw, hreg);
                    lds8sw(din + o, sw, dreg);
                    if (lane >= nblk) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) { lab[j] = LADDR(L); hreg[j] = T(0); dreg[j] = T(0); }
                    }
                    labm1 = __shfl_up_sync(ASG_FULL_MASK, last8(lab, fl), 1);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&cempty[buf]);      // the constants now live in registers
                ++useq;
            } else {
                Tn = (int)a.ctx.il[n];
                gf = FCC ? a.gfull(n) : T(0);
                ga = FAC ? a.galigned(n) : T(0);
                Sn = (int)a.ctx.tl[n];
                const int64_t *yn = a.y + (int64_t)n * a.y_s0;
                cbar();                        // the previous utterance's frames are done everywhere
                for (int s = tid; s < sp; s += ncons) {
                    int lb = L;
                    if (s < Sn) lb = max(0, min(L - 1, (int)yn[(int64_t)s * a.y_s1]));
                    ylab[s] = lb;
                    hh[s] = a.ctx.fac_hd[(int64_t)n * sp + s];
                    din[s] = (s >= 1) ? a.ctx.fac_hd[((int64_t)N + n) * sp + s - 1] : T(0);
                }
                cbar();
            }
        }
        const int rows_valid = max(0, min(TBt, Tn - t0));   // frames t0 .. t0+rows_valid-1 are inside the utterance
        char *sb = stages + (size_t)stage * mp.stage_bytes;
        // record r of the stage = frame t0-1+r (r = 0 is the halo); every part of a frame sits in its own record
        const char *srec = sb + mp.oRec;
        auto A_ = [&](int r) { return (const T *)(srec + (size_t)r * R + a.ctx.o_fa); };
        auto B_ = [&](int r) { return (const T *)(srec + (size_t)r * R + a.ctx.o_fb); };
        auto Ea_ = [&](int r) { return (const int *)(srec + (size_t)r * R + a.ctx.o_ea); };
        auto Eb_ = [&](int r) { return (const int *)(srec + (size_t)r * R + a.ctx.o_eb); };
        auto Fa_ = [&](int r) { return (const T *)(srec + (size_t)r * R + a.ctx.o_ca); };
        auto Fb_ = [&](int r) { return (const T *)(srec + (size_t)r * R + a.ctx.o_cb); };
        auto Fw_ = [&](int r) { return (const T *)(srec + (size_t)r * R + a.ctx.o_cw); };
        auto Sc_ = [&](int r) { return *(const T *)(srec + (size_t)r * R + a.ctx.o_sc); };

        mbar_wait(&full[stage], use & 1);

        // smem row r of the alpha arrays = frame t0-1+r (r = 0 is the halo); beta arrays: row r-1 = frame t0-1+r
        // ---- fast path: both of this warp's frames are live interior frames -> do them together, so that the two
        //      independent dependency chains (loads, REDUX, shuffle trees, divisions) overlap in one instruction stream
        if (narrow && TB == 2 * NW && rows_valid == TB && t0 > 0 && !a.dry) {
            const int ra = 1 + warp, rb = 1 + warp + NW;
            const int rr[2] = {ra, rb};
            const bool act = lane < nblk;
            const unsigned rowa[2] = {smem_u32(srec) + (unsigned)ra * (unsigned)R, smem_u32(srec) + (unsigned)rb * (unsigned)R};   // this warp's two frame records
            // The per-lane offsets into a record are RECOMPUTED here for every tile (a dozen integer instructions from `lane`
            // and the layout constants in the parameter bank) instead of living in ~14 registers across the whole kernel: kept
            // alive they were the registers ptxas spilled, and the spill reloads sat at the head of every tile.  The empty asm
            // makes the lane index opaque so that the compiler does not hoist them back out of the loop.
            int lv = lane;
            asm volatile("" : "+r"(lv));
            const int lbk_ = min(lv, max(nblk, 1) - 1);
            const unsigned oA_lo = (unsigned)(a.ctx.o_fa + lbk_ * 8 * (int)sizeof(T) + sw * (int)sizeof(T)), oA_hi = oA_lo + (unsigned)((4 - 2 * sw) * (int)sizeof(T));
            const unsigned oB_lo = (unsigned)(a.ctx.o_fb + lbk_ * 8 * (int)sizeof(T) + sw * (int)sizeof(T)), oB_hi = oB_lo + (unsigned)((4 - 2 * sw) * (int)sizeof(T));
            const int offEa = a.ctx.o_ea + lbk_ * 4, offEb = a.ctx.o_eb + lbk_ * 4;
            const int l0_ = has0 ? lv : 0, l1_ = has1 ? lv + 32 : l0_;
            const int offCa0 = a.ctx.o_ca + l0_ * (int)sizeof(T), offCb0 = a.ctx.o_cb + l0_ * (int)sizeof(T);
            const int offCa1 = a.ctx.o_ca + l1_ * (int)sizeof(T), offCb1 = a.ctx.o_cb + l1_ * (int)sizeof(T);
            const int offCw = a.ctx.o_cw + (lv >> 2) * RI * (int)sizeof(T), offPw = a.ctx.o_ca + (lv & 3) * CJ * (int)sizeof(T);
            T bv[2][8];
            if (FAC) {
                T av[2][8];
                int ea[2], eb[2], da[2], db[2];
                T fa[2], fb[2];                             // normalising factors of the stored mantissas (x half the product scale)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    lds8_at(rowa[u] + oA_lo, rowa[u] + oA_hi, av[u]);
                    lds8_at(rowa[u] + oB_lo, rowa[u] + oB_hi, bv[u]);
                    const int wa = lds_s32_at(rowa[u] + (unsigned)offEa), wb = lds_s32_at(rowa[u] + (unsigned)offEb);
                    fac_unpack_norm<T>(wa, ea[u], da[u], fa[u]);
                    fac_unpack_norm<T>(wb, eb[u], db[u], fb[u]);
                }
                T pv[2][8];
                int ep[2];                                  // exponent of the RAW mantissas of frame t-1 (E + d)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    lds8_at(rowa[u] - (unsigned)R + oA_lo, rowa[u] - (unsigned)R + oA_hi, pv[u]);
                    int e_, d_;
                    fac_unpack(lds_s32_at(rowa[u] - (unsigned)R + (unsigned)offEa), e_, d_);
                    ep[u] = e_ + d_;
                }
                T ps[2];
                int c[2], ce[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    // normalised mantissas (block maximum in [1,2)), each carrying half of the product scale
#pragma unroll
                    for (int j = 0; j < 8; ++j) bv[u][j] = (av[u][j] * fa[u]) * (bv[u][j] * fb[u]);
                    const T pm = tmax(tmax(tmax(bv[u][0], bv[u][1]), tmax(bv[u][2], bv[u][3])), tmax(tmax(bv[u][4], bv[u][5]), tmax(bv[u][6], bv[u][7])));
                    ps[u] = ((bv[u][0] + bv[u][1]) + (bv[u][2] + bv[u][3])) + ((bv[u][4] + bv[u][5]) + (bv[u][6] + bv[u][7]));
                    c[u] = ea[u] + eb[u];
                    ce[u] = (act && pm > T(0)) ? c[u] + Num<T>::pexp(pm) : INT_MIN;
                }
                const int cmax0 = __reduce_max_sync(ASG_FULL_MASK, ce[0]);
                const int cmax1 = __reduce_max_sync(ASG_FULL_MASK, ce[1]);
                const int sh[2] = {Num<T>::block_shift(c[0], cmax0), Num<T>::block_shift(c[1], cmax1)};
                T z0 = (ce[0] != INT_MIN) ? Num<T>::ldexp_pos(ps[0], sh[0]) : T(0);
                T z1 = (ce[1] != INT_MIN) ? Num<T>::ldexp_pos(ps[1], sh[1]) : T(0);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    z0 += __shfl_xor_sync(ASG_FULL_MASK, z0, o);
                    z1 += __shfl_xor_sync(ASG_FULL_MASK, z1, o);
                }
                const T Z[2] = {z0, z1};
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    // 1/Z and the lane's weight in one factor; the floor on Z keeps the factor finite (a frame whose total
                    // product mass is below it has underflowed anyway: its posterior is dropped)
                    const T c2 = (Z[u] > Num<T>::tiny() * FX::kPost * T(4) && ce[u] != INT_MIN)
                                     ? Num<T>::ldexp_pos(rcp_(Z[u]), sh[u] + Num<T>::kPostLog2) : T(0);
                    bin_t *bu = bins + u * GRAD_LBINS;
#pragma unroll
                    for (int j = 0; j < 8; ++j) bv[u][j] = bv[u][j] * c2;
                    // idle lanes stay out: they would all hit ONE trash bin (same-address serialisation; 7.75 wavefronts per
                    // reduction were measured with them in, the shared-memory pipe's largest single cost)
                    if constexpr (sizeof(T) == 4) {
                        int bq[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) bq[j] = FX::to_fix(bv[u][j]);
                        if (u == 0) red8_shared_s32_if<0>(lab, bq, act);
                        else red8_shared_s32_if<GRAD_LBINS * 4>(lab, bq, act);
                    } else if (act) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) FX::add(&bu[LAB(lab[j])], FX::to_fix(bv[u][j]));
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    T left = __shfl_up_sync(ASG_FULL_MASK, last8(pv[u], fl), 1);
                    const int el = __shfl_up_sync(ASG_FULL_MASK, ep[u], 1);
                    left = (lane == 0) ? T(0) : left * Num<T>::exp2i_flush(min(el - ep[u], Num<T>::kMaxShift - 8));
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const T stay = hreg[j] * pv[u][j];
                        const T den = fma(dreg[j], prev8(pv[u], j, fl, left), stay);
                        const T fs = fast_div(stay, tmax(den, Num<T>::tiny()));
                        accH[j] = fma(bv[u][j], fs, accH[j]);
                        accQ[j] += bv[u][j];
                    }
                }
                __syncwarp();
            }
            T g0[2] = {T(0), T(0)}, g1[2] = {T(0), T(0)};
            if (FCC) {
                T inv[2];
                {
                    T al0[2], al1[2], be0[2], be1[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        // every lane loads (idle lanes re-read element 0 and drop it): no divergent branch around the loads
                        al0[u] = lds_at(rowa[u] + (unsigned)offCa0, T()); be0[u] = lds_at(rowa[u] + (unsigned)offCb0, T());
                        al1[u] = be1[u] = T(0);
                        if (LP > 32) { al1[u] = lds_at(rowa[u] + (unsigned)offCa1, T()); be1[u] = lds_at(rowa[u] + (unsigned)offCb1, T()); }
                        g0[u] = has0 ? al0[u] * be0[u] : T(0); g1[u] = has1 ? al1[u] * be1[u] : T(0);
                    }
                    T z0 = g0[0] + g1[0], z1 = g0[1] + g1[1];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        z0 += __shfl_xor_sync(ASG_FULL_MASK, z0, o);
                        z1 += __shfl_xor_sync(ASG_FULL_MASK, z1, o);
                    }
                    inv[0] = (z0 > T(0)) ? gf * rcp_(z0) : T(0);
                    inv[1] = (z1 > T(0)) ? gf * rcp_(z1) : T(0);
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    g0[u] *= inv[u]; g1[u] *= inv[u];
                    const T winv = inv[u] * lds_at(rowa[u] + (unsigned)a.ctx.o_sc, T());
                    T wv[RI];
                    V2 pw[CJ2];
#pragma unroll
                    for (int i = 0; i < RI; ++i) wv[i] = lds_at(rowa[u] + (unsigned)offCw + (unsigned)(i * sizeof(T)), T()) * winv;
#pragma unroll
                    for (int j = 0; j < CJ2; ++j) pw[j] = lds_v2_at(rowa[u] - (unsigned)R + (unsigned)offPw + (unsigned)(j * 2 * sizeof(T)), T());
#pragma unroll
                    for (int i = 0; i < RI; ++i) {
                        const V2 w2 = Pk<T>::mk(wv[i], wv[i]);
#pragma unroll
                        for (int j = 0; j < CJ2; ++j) G[i][j] = Pk<T>::fma2(w2, pw[j], G[i][j]);
                    }
                }
            }
            T *const gx_tile = a.grad_x + ((int64_t)(t0 + warp) * N + n) * L;      // frame t0-1+ra; the second frame is NW frames later
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                T *gx_row = gx_tile + (u ? gx_pair_stride : (int64_t)0);
                T o0 = g0[u], o1 = g1[u];
                if (FAC) {
                    bin_t *bu = bins + u * GRAD_LBINS;
                    const T gs = ga * FX::kInvPost;
                    if (lane < L) { o0 = fma(gs, FX::from_fix(bu[lane]), o0); bu[lane] = 0; }
                    if (LP > 32 && lane + 32 < L) { o1 = fma(gs, FX::from_fix(bu[lane + 32]), o1); bu[lane + 32] = 0; }
                }
                if (lane < L) gx_row[lane] = o0;
                if (LP > 32 && lane + 32 < L) gx_row[lane + 32] = o1;
            }
            __syncwarp();
        } else
        for (int r = 1 + warp; r <= TBt; r += NW) {
            const int t = t0 - 1 + r;
            T *gx_row = a.grad_x + ((int64_t)t * N + n) * L;
            if (r > rows_valid || a.dry) {              // padding frame: exact zeros
                for (int l = lane; l < L; l += 32) gx_row[l] = T(0);
                continue;
            }
            // ---------------- FAC: posterior, label scatter, edge posteriors ----------------------
            if (FAC) {
                if (narrow) {
                    const int lb8 = min(lane, nblk - 1) * 8;          // idle lanes re-read the last block (weight 0)
                    const int le = min(lane, nblk - 1);
                    const bool act = lane < nblk;
                    T av[8], bv[8];
                    lds8sw(A_(r) + lb8, sw, av);
                    lds8sw(B_(r) + lb8, sw, bv);
                    int ea, eb, da, db;
                    const int wa = Ea_(r)[le], wb = Eb_(r)[le];
                    T fa, fb;
                    fac_unpack_norm<T>(wa, ea, da, fa);
                    fac_unpack_norm<T>(wb, eb, db, fb);
                    // alpha and beta mantissas are normalised separately (block maxima in [1,2) after the drift correction), so
                    // where one peaks the other can be ~2^-100: the products carry a constant 2^110 (2^960) to stay normal numbers
#pragma unroll
                    for (int j = 0; j < 8; ++j) bv[j] = (av[j] * fa) * (bv[j] * fb);
                    const T pm = tmax(tmax(tmax(bv[0], bv[1]), tmax(bv[2], bv[3])), tmax(tmax(bv[4], bv[5]), tmax(bv[6], bv[7])));
                    const T ps = ((bv[0] + bv[1]) + (bv[2] + bv[3])) + ((bv[4] + bv[5]) + (bv[6] + bv[7]));
                    const int c = ea + eb;
                    const int ce = (act && pm > T(0)) ? c + Num<T>::pexp(pm) : INT_MIN;    // a block without mass is ignored
                    const int cmax = __reduce_max_sync(ASG_FULL_MASK, ce);
                    const int sh = Num<T>::block_shift(c, cmax);
                    const T Z = warp_sum((ce != INT_MIN) ? Num<T>::ldexp_pos(ps, sh) : T(0));
                    const T c2 = (Z > Num<T>::tiny() * FX::kPost * T(4) && ce != INT_MIN) ? Num<T>::ldexp_pos(rcp_(Z), sh + Num<T>::kPostLog2) : T(0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        bv[j] = bv[j] * c2;                            // posterior of position 8*lane+j, in fixed-point units
                        if (act) FX::add(&bins[LAB(lab[j])], FX::to_fix(bv[j]));
                    }
                    if (t >= 1) {
                        T pv[8];
                        lds8sw(A_(r - 1) + lb8, sw, pv);
                        int ep, dp_;
                        fac_unpack(Ea_(r - 1)[le], ep, dp_);
                        ep += dp_;                                 // exponent of the raw mantissas
                        // the left neighbour's last value, expressed in this block's exponent (bounded so den stays finite)
                        T left = __shfl_up_sync(ASG_FULL_MASK, last8(pv, fl), 1);
                        const int el = __shfl_up_sync(ASG_FULL_MASK, ep, 1);
                        left = (lane == 0) ? T(0) : left * Num<T>::exp2i_flush(min(el - ep, Num<T>::kMaxShift - 8));
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const T stay = hreg[j] * pv[j];
                            const T den = fma(dreg[j], prev8(pv, j, fl, left), stay);
                            // den == 0 implies a zero posterior, so the floor only guards 0/0
                            const T fs = fast_div(stay, tmax(den, Num<T>::tiny()));     // share of the mass that stayed (<= 1)
                            accH[j] = fma(bv[j], fs, accH[j]);
                            accQ[j] += bv[j];
                        }
                    }
                } else {
                    // wide targets (S > 256): several blocks per lane; three passes over the frame, sums parked in shared memory
                    int cmax = INT_MIN;
                    for (int b = lane; b < nblk; b += 32) {
                        T av[8], bv[8];
                        lds8(A_(r) + b * 8, av);
                        lds8(B_(r) + b * 8, bv);
                        T pm = T(0);
                        int ea, eb, da, db;
                        T fa, fb;
                        fac_unpack_norm<T>(Ea_(r)[b], ea, da, fa);
                        fac_unpack_norm<T>(Eb_(r)[b], eb, db, fb);
#pragma unroll
                        for (int j = 0; j < 8; ++j) pm = tmax(pm, (av[j] * fa) * (bv[j] * fb));
                        if (pm > T(0)) cmax = max(cmax, ea + eb + Num<T>::pexp(pm));
                    }
                    cmax = __reduce_max_sync(ASG_FULL_MASK, cmax);
                    T zs = T(0);
                    for (int b = lane; b < nblk; b += 32) {
                        T av[8], bv[8];
                        lds8(A_(r) + b * 8, av);
                        lds8(B_(r) + b * 8, bv);
                        int ea, eb, da, db;
                        const int wa = Ea_(r)[b], wb = Eb_(r)[b];
                        T fa, fb;
                        fac_unpack_norm<T>(wa, ea, da, fa);
                        fac_unpack_norm<T>(wb, eb, db, fb);
                        const int c = ea + eb;
                        T pm = T(0), ps = T(0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) { const T p = (av[j] * fa) * (bv[j] * fb); pm = tmax(pm, p); ps += p; }
                        if (pm > T(0)) zs += Num<T>::ldexp_pos(ps, Num<T>::block_shift(c, cmax));
                    }
                    const T Z = warp_sum(zs);
                    const T invZ = (Z > Num<T>::tiny()) ? T(1) / Z : T(0);
                    for (int b = lane; b < nblk; b += 32) {
                        T av[8], bv[8];
                        lds8(A_(r) + b * 8, av);
                        lds8(B_(r) + b * 8, bv);
                        int ea, eb, da, db;
                        T fa, fb;
                        fac_unpack_norm<T>(Ea_(r)[b], ea, da, fa);
                        fac_unpack_norm<T>(Eb_(r)[b], eb, db, fb);
                        const int c = ea + eb;
                        T pm = T(0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) { bv[j] = (av[j] * fa) * (bv[j] * fb); pm = tmax(pm, bv[j]); }
                        T wgt = T(0);
                        if (pm > T(0)) wgt = Num<T>::ldexp_pos(invZ, Num<T>::block_shift(c, cmax) + Num<T>::kPostLog2);
                        if (wgt == T(0)) continue;
                        T pv[8], left = T(0);
                        if (t >= 1) {
                            lds8(A_(r - 1) + b * 8, pv);
                            if (b > 0)
                            {
                                int e0, d0, e1, d1;
                                fac_unpack(Ea_(r - 1)[b - 1], e0, d0);
                                fac_unpack(Ea_(r - 1)[b], e1, d1);
                                left = A_(r - 1)[b * 8 - 1] * Num<T>::exp2i(min((e0 + d0) - (e1 + d1), Num<T>::kMaxShift - 8));
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const T post = bv[j] * wgt;
                            const int s = b * 8 + j;
                            FX::add(&bins[ylab[s]], FX::to_fix(post));
                            if (t >= 1) {
                                const T stay = hh[s] * pv[j];
                                const T den = fma(din[s], j ? pv[j ? j - 1 : 0] : left, stay);
                                const T fs = fast_div(stay, tmax(den, Num<T>::tiny()));
                                acc[s] = fma(post, fs, acc[s]);
                                acc[sp + s] += post;
                            }
                        }
                    }
                }
                __syncwarp();
            }
            // ---------------- FCC: posterior, outer product; grad_x row -----------------------------
            T g0 = T(0), g1 = T(0);
            if (FCC) {
                T al0 = T(0), al1 = T(0), be0 = T(0), be1 = T(0);
                if (lane < lpi) { al0 = Fa_(r)[lane]; be0 = Fb_(r)[lane]; }
                if (LP > 32 && lane + 32 < lpi) { al1 = Fa_(r)[lane + 32]; be1 = Fb_(r)[lane + 32]; }
                g0 = al0 * be0; g1 = al1 * be1;
                const T Z = warp_sum(g0 + g1);
                const T inv = (Z > T(0)) ? gf * rcp_(Z) : T(0);
                g0 *= inv; g1 *= inv;
                if (t >= 1) {                                  // no incoming edge at frame 0
                    const T winv = inv * Sc_(r);
                    T wv[RI];
                    V2 pv[CJ2];
#pragma unroll
                    for (int i = 0; i < RI; ++i) wv[i] = Fw_(r)[ib * RI + i] * winv;
#pragma unroll
                    for (int j = 0; j < CJ2; ++j) pv[j] = ((const V2 *)(Fa_(r - 1) + jb * CJ))[j];
#pragma unroll
                    for (int i = 0; i < RI; ++i) {
                        const V2 w2 = Pk<T>::mk(wv[i], wv[i]);
#pragma unroll
                        for (int j = 0; j < CJ2; ++j) G[i][j] = Pk<T>::fma2(w2, pv[j], G[i][j]);
                    }
                }
            }
            if (FAC) {
                const T gs = ga * FX::kInvPost;
                if (lane < L) { g0 = fma(gs, FX::from_fix(bins[lane]), g0); bins[lane] = 0; }
                if (LP > 32 && lane + 32 < L) { g1 = fma(gs, FX::from_fix(bins[lane + 32]), g1); bins[lane + 32] = 0; }
                __syncwarp();
            }
            if (lane < L) gx_row[lane] = g0;
            if (LP > 32 && lane + 32 < L) gx_row[lane + 32] = g1;
        }
        // release the stage (read-only for the consumers)
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == NS) { stage = 0; ++use; }
        if (++wt == a.tiles_t) { wt = 0; ++wn; }
    }
    flush_fac();
    cbar();

    // ---- per-CTA partial of grad_A = E o G (FCC) + Gal (FAC) ----
    // Every tile has been consumed, so the tile ring is idle: each warp parks its G slice there and, after ONE
    // barrier, the element owners add the NW slices in warp order (fixed order: deterministic).  When the ring is
    // too small for NW slices (tiny tiles) the warps take turns instead.
    const bool park = !a.nopark && (size_t)NW * L * L * sizeof(T) <= (size_t)NS * mp.stage_bytes;
    if (FCC && park) {
        T *slice = (T *)stages + (size_t)warp * L * L;
#pragma unroll
        for (int i = 0; i < RI; ++i)
#pragma unroll
            for (int j = 0; j < CJ; ++j) {
                const int gi = ib * RI + i, gj = jb * CJ + j;
                if (gi < L && gj < L) slice[gi * L + gj] = (j & 1) ? G[i][j >> 1].y : G[i][j >> 1].x;
            }
        cbar();
        for (int e = tid; e < L * L; e += ncons) {
            T v = T(0);
            for (int w = 0; w < NW; ++w) v += ((const T *)stages)[(size_t)w * L * L + e];
            Gf[e] = v;
        }
    } else if (FCC) {
        for (int w = 0; w < NW; ++w) {
            if (warp == w) {
#pragma unroll
                for (int i = 0; i < RI; ++i)
#pragma unroll
                    for (int j = 0; j < CJ; ++j) {
                        const int gi = ib * RI + i, gj = jb * CJ + j;
                        if (gi < L && gj < L) Gf[gi * L + gj] += (j & 1) ? G[i][j >> 1].y : G[i][j >> 1].x;
                    }
            }
            cbar();
        }
    }
    T *part = a.partials + (size_t)blockIdx.x * L * L;
    for (int e = tid; e < L * L; e += ncons) {
        T v = Gal[e];
        if (FCC) v += a.ctx.ehat[e] * Gf[e];
        part[e] = v;
    }
}

// grad_A[e] = sum over CTAs of partial[c][e].  Fixed order: warp w sums parts w, w+8, ...; warp 0 adds the 8 sums.
// With `losses` set (asg_step) one extra CTA also writes loss[n] = full[n] - aligned[n] and the fused tail (sum, count).
template <typename T>
__global__ void __launch_bounds__(256) grad_finalize_kernel(const T *partials, int nparts, int LL, T *grad_A, const T *full,
                                                            const T *aligned, T *losses, int N, T *tail) {
    __shared__ T red[8][33];
    if (losses != nullptr && blockIdx.x == gridDim.x - 1) {
        write_losses_block<T>(full, aligned, losses, N, tail, &red[0][0]);
        return;
    }
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
    if (device < 0 || device >= 64) return 148;
    if (!sms[device]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
        sms[device] = v;
    }
    return sms[device];
}

template <typename T, int LP, bool FCC, bool FAC>
static cudaError_t launch_grad_lp(const GradArgs<T> &a_in, cudaStream_t st) {
    GradArgs<T> a = a_in;
    const int L = a.d.L, sp = (int)a.ctx.sp;
    int device = 0;
    cudaGetDevice(&device);
    // (frames per tile, ring depth, consumer warps): one persistent CTA per SM.  Two stages, not three: measured 64.5 vs 67.7 us
    // at cfg-2 -- the third stage hides nothing more (tiles are consumed in 2.6 us, a bulk copy lands in ~1.5) and its 59 kB
    // are worth more as L1 (register spills, per-utterance constants)
    static const int cand[][3] = {{24, 2, 12}, {24, 3, 12}, {16, 5, 8}, {16, 4, 8}, {16, 3, 8}, {8, 4, 8}, {8, 3, 8}, {8, 2, 8}, {4, 3, 4}, {4, 2, 4},
                                  {2, 2, 2}, {1, 2, 2}, {1, 2, 1}, {1, 1, 1}};
    int tb = 0, ns = 0, nw = 0;
    size_t bytes = 0;
    int force_tb = 0, force_ns = 0;
    if (const char *env = getenv("ASG_GRAD_TB")) force_tb = atoi(env);
    if (const char *env = getenv("ASG_GRAD_NS")) force_ns = atoi(env);
    for (auto &c : cand) {
        if (force_tb && c[0] != force_tb) continue;
        if (force_ns && c[1] != force_ns) continue;
        const size_t b = grad_map(sizeof(T), (size_t)a.ctx.R, L, sp, c[0], c[1], c[2], FCC, FAC).total;
        if (b <= 227 * 1024) { tb = c[0]; ns = c[1]; nw = c[2]; bytes = b; break; }
    }
    if (!tb) return cudaErrorInvalidValue;
    a.tb = tb;
    a.nstage = ns;
    a.nwarps = nw;
    a.dry = getenv("ASG_GRAD_DRY") ? 1 : 0;
    a.nopark = getenv("ASG_GRAD_NOPARK") ? 1 : 0;
    a.tiles_t = (a.d.T + tb - 1) / tb;
    a.num_tiles = a.tiles_t * a.d.N;
    cudaError_t e = cudaFuncSetAttribute(grad_kernel<T, LP, FCC, FAC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    // developer knob: a carve-out of 100 (%) reproduces the L1-starved timing (+3 us at cfg-2: the register spills of 512 threads
    // are 45 kB of local memory, which 28 kB of L1 cannot hold)
    if (const char *env = getenv("ASG_GRAD_CARVE")) cudaFuncSetAttribute(grad_kernel<T, LP, FCC, FAC>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(env));
    int grid = grad_grid_size(device);
    if (grid > a.num_tiles) grid = a.num_tiles;
    {
        LaunchScope ls(K_GRAD, st);
        grad_kernel<T, LP, FCC, FAC><<<grid, (((nw + 3) & ~3) + 4) * 32, bytes, st>>>(a);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    const int LL = L * L;
    {
        LaunchScope ls(K_GRAD_FINALIZE, st);
        grad_finalize_kernel<T><<<(LL + 31) / 32 + (a.losses ? 1 : 0), 256, 0, st>>>(a.partials, grid, LL, a.grad_A, a.full_scores,
                                                                                   a.aligned_scores, a.losses, a.d.N, a.tail);
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
