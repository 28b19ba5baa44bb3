// asg_common.cuh -- device helpers shared by the scan and gradient-assembly kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define ASG_FULL_MASK 0xffffffffu
#define ASG_BLK 8                 // FAC states per lane = per block-floating-point exponent

namespace asg {

// ---------------------------------------------------------------------------------------------
// scalar traits
// ---------------------------------------------------------------------------------------------
template <typename T> struct Num;

template <> struct Num<float> {
    static constexpr int kMaxShift = 126;
    __device__ __forceinline__ static float exp_(float v) { return expf(v); }
    __device__ __forceinline__ static float log_(float v) { return logf(v); }
    __device__ __forceinline__ static float neg_inf() { return __int_as_float(0xff800000); }
    __device__ __forceinline__ static float tiny() { return 1.0e-37f; }
    static constexpr int kHalfProd = 55;       // alpha and beta mantissas each carry 2^55 (their normalised products 2^110)
    __device__ __forceinline__ static float prod_scale() { return 1.2980742146337069e33f; }     // 2^110
    // 2^k for k in [-126,127] (clamped)
    __device__ __forceinline__ static float exp2i(int k) {
        k = max(-126, min(127, k));
        return __int_as_float((k + 127) << 23);
    }
    // 2^k, exactly zero for k below the normal range (mass that far below is flushed, never clamped to the smallest normal)
    __device__ __forceinline__ static float exp2i_flush(int k) { return k < -126 ? 0.0f : __int_as_float((min(k, 127) + 127) << 23); }
    // Range report of the gradient assembly: a block b whose alpha and beta are both non-empty may hold entries flushed below
    // 2^-126 of its maxima, i.e. lost products up to 2^(c_b - 121); the frame's largest product is 2^(cmax_eff - 237).  The loss is
    // below 2^-24 of it iff cmax_eff - c_b >= kSafeGap.
    static constexpr int kSafeGap = 127 + 110 - 126 + 5 + 24;      // = 140 (coarse test; fac_block_at_risk() is the per-position one)
    static constexpr int kMantBits = 24;
    __device__ __forceinline__ static float min_normal() { return 1.17549435e-38f; }
    // Posterior weights of the gradient assembly.  A block's products alpha_j * beta_j * prod_scale live in units of
    // 2^(ea + eb); pexp() is the biased exponent of the block's largest product, so ea + eb + pexp ranks the blocks of a
    // frame by their TRUE largest product (the mantissas are not normalised to [1,2): the scans renormalise lazily, and
    // inside a block alpha and beta can peak at opposite ends).  block_shift() is the power of two that brings a block to
    // the units in which the frame's largest product is ~2^100; it can exceed the exponent range of a float factor (a
    // dominant block may hold products as small as 2^-126), so it is APPLIED with ldexp_pos(), by exponent arithmetic.
    __device__ __forceinline__ static int pexp(float pm) { return (__float_as_int(pm) >> 23) & 0xff; }
    __device__ __forceinline__ static int block_shift(int c, int cmax_eff) { return max(c - cmax_eff, -100000) + (127 + 100); }
    static constexpr int kPostLog2 = 30;
    // v * 2^s for v >= 0 finite: flushed to zero when it underflows, saturated just below the overflow threshold
    __device__ __forceinline__ static float ldexp_pos(float v, int s) {
        const int b = __float_as_int(v), e = b >> 23, en = min(e + s, 253);
        return (e <= 0 || en <= 0) ? 0.0f : __int_as_float((b & 0x007fffff) | (en << 23));
    }
    // For M >= 0: power-of-two `scale` with M*scale in [1,2) and e = -log2(scale).  M == 0, inf or
    // NaN leave the value alone.  A denormal M is only pulled up by 2^126 (the next frame finishes).
    __device__ __forceinline__ static void pow2_norm(float M, float &scale, int &e) {
        int biased = (__float_as_int(M) >> 23) & 0xff;
        bool live = (M > 0.0f) && (biased != 0xff);
        biased = max(1, min(253, biased));
        e = live ? biased - 127 : 0;
        scale = live ? __int_as_float((254 - biased) << 23) : 1.0f;
    }
    // max over the warp of a NON-NEGATIVE value: IEEE order == unsigned integer order -> one REDUX
    __device__ __forceinline__ static float warp_max_nonneg(float v) {
        return __uint_as_float(__reduce_max_sync(ASG_FULL_MASK, __float_as_uint(v)));
    }
    // max over the warp of any (non-NaN) value via an order-preserving signed key
    __device__ __forceinline__ static float warp_max(float v) {
        int b = __float_as_int(v);
        int key = b ^ ((b >> 31) & 0x7fffffff);
        key = __reduce_max_sync(ASG_FULL_MASK, key);
        return __int_as_float(key ^ ((key >> 31) & 0x7fffffff));
    }
};

template <> struct Num<double> {
    static constexpr int kMaxShift = 1022;
    __device__ __forceinline__ static double exp_(double v) { return exp(v); }
    __device__ __forceinline__ static double log_(double v) { return log(v); }
    __device__ __forceinline__ static double neg_inf() { return __longlong_as_double(0xfff0000000000000LL); }
    __device__ __forceinline__ static double tiny() { return 1.0e-300; }
    static constexpr int kHalfProd = 480;
    __device__ __forceinline__ static double prod_scale() { return 9.745314011399999e288; }     // 2^960
    __device__ __forceinline__ static double exp2i(int k) {
        k = max(-1022, min(1023, k));
        return __hiloint2double((k + 1023) << 20, 0);
    }
    __device__ __forceinline__ static double exp2i_flush(int k) { return k < -1022 ? 0.0 : __hiloint2double((min(k, 1023) + 1023) << 20, 0); }
    static constexpr int kMantBits = 53;
    __device__ __forceinline__ static double min_normal() { return 2.2250738585072014e-308; }
    static constexpr int kSafeGap = 1023 + 960 - 1022 + 5 + 53;    // same argument: bias + product scale - min exponent + 5 + 53-bit margin
    __device__ __forceinline__ static int pexp(double pm) { return (__double2hiint(pm) >> 20) & 0x7ff; }
    __device__ __forceinline__ static int block_shift(int c, int cmax_eff) { return max(c - cmax_eff, -100000) + (1023 + 900); }
    static constexpr int kPostLog2 = 62;
    __device__ __forceinline__ static double ldexp_pos(double v, int s) {
        const int hi = __double2hiint(v), e = hi >> 20, en = min(e + s, 2045);
        return (e <= 0 || en <= 0) ? 0.0 : __hiloint2double((hi & 0x000fffff) | (en << 20), __double2loint(v));
    }
    __device__ __forceinline__ static void pow2_norm(double M, double &scale, int &e) {
        int biased = (__double2hiint(M) >> 20) & 0x7ff;
        bool live = (M > 0.0) && (biased != 0x7ff);
        biased = max(1, min(2045, biased));
        e = live ? biased - 1023 : 0;
        scale = live ? __hiloint2double((2046 - biased) << 20, 0) : 1.0;
    }
    __device__ __forceinline__ static double warp_max_nonneg(double v) { return warp_max(v); }
    __device__ __forceinline__ static double warp_max(double v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(ASG_FULL_MASK, v, o));
        return v;
    }
};

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(ASG_FULL_MASK, v, o);
    return v;
}

template <typename T> __device__ __forceinline__ T tmax(T a, T b) { return a > b ? a : b; }
template <> __device__ __forceinline__ float tmax<float>(float a, float b) { return fmaxf(a, b); }
template <> __device__ __forceinline__ double tmax<double>(double a, double b) { return fmax(a, b); }

// ---------------------------------------------------------------------------------------------
// Exponent word of one FAC block in a frame record: (E << 8) | (d + 8).  The block's eight stored mantissas m_j stand for
// m_j * 2^(E + d); d = -floor(log2 max_j m_j), so m_j * 2^d has its block maximum in [1,2) and E is the exponent of that
// normalised block.  The fp32 scans renormalise lazily (their stored mantissas drift inside [2^-40, 8)) and record the
// drift d so that the gradient assembly still multiplies NORMALISED mantissas (full fp32 range for the in-block spread);
// consumers that normalise on every frame store d = 0.  An EMPTY block (all mantissas zero) stores d = -8, i.e. a zero low byte.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int fac_pack(int E_raw, int d, bool live) { return live ? (((E_raw - d) << 8) | ((d + 8) & 0xff)) : (E_raw << 8); }
__device__ __forceinline__ void fac_unpack(int w, int &E, int &d) { E = w >> 8; d = (w & 0xff) - 8; }
__device__ __forceinline__ bool fac_live(int w) { return (w & 0xff) != 0; }
// The drift field holds -8..247.  A generic (fp64) step whose block maximum fell by more than that stores NORMALISED mantissas
// under the reserved code 247 = "drift up to 2^kFacDriftCoarseValue" (beyond that the scan itself raises the range flag).
constexpr int kFacDriftCoarse = 247, kFacDriftCoarseValue = 900;
__device__ __forceinline__ int fac_pack_coarse(int E_norm) { return ((E_norm - kFacDriftCoarse) << 8) | 0xff; }
// For the gradient assembly: E = exponent of the block's NORMALISED mantissas, f = the factor that normalises the stored
// ones times half of the product scale.  A drift so large that the factor would leave the format (a block whose maximum fell
// by more than 2^72 in one frame) is normalised as far as representable and the rest stays in E.
template <typename T> __device__ __forceinline__ void fac_unpack_norm(int w, int &E, int &d, T &f) {
    fac_unpack(w, E, d);
    if (sizeof(T) == 8 && d == kFacDriftCoarse) {       // fp64 only: normalised mantissas, "the maximum fell by up to 2^900"
        E += d;
        f = Num<T>::exp2i(Num<T>::kHalfProd);
        d = kFacDriftCoarseValue;
        return;
    }
    const int dc = min(d, Num<T>::kMaxShift + 1 - Num<T>::kHalfProd);
    E += d - dc;
    f = Num<T>::exp2i(dc + Num<T>::kHalfProd);
}
// Range report, per position (asg_range_flags, run on demand over the stored states; not on the step's path).  An entry of a
// block that is zero (below the normal range) may be a FLUSHED one: up to 2^(-kMaxShift + drift) of its block maximum.  Paired
// with the other factor AT THE SAME POSITION that bounds the product that may have been lost; the block is at risk when that bound
// is not 2^-(mantissa bits) below the frame's largest product (ce units: exponent sum + biased exponent of the scaled product).
template <typename T>
__device__ __forceinline__ bool fac_block_at_risk(const T (&al)[8], const T (&be)[8], T fa, T fb, int da, int db, int c, int cmax) {
    T a0 = T(0), b0 = T(0);          // largest normalised alpha where beta is zero, beta where alpha is zero
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (!(al[j] >= Num<T>::min_normal())) b0 = tmax(b0, be[j] * fb);
        if (!(be[j] >= Num<T>::min_normal())) a0 = tmax(a0, al[j] * fa);
    }
    const int k = Num<T>::kHalfProd - Num<T>::kMaxShift + 3 + Num<T>::kMantBits;
    bool risk = false;
    if (b0 > T(0)) risk |= cmax - c < Num<T>::pexp(b0) + k + max(da, 0);
    if (a0 > T(0)) risk |= cmax - c < Num<T>::pexp(a0) + k + max(db, 0);
    return risk;
}

// ---------------------------------------------------------------------------------------------
// cp.async (LDGSTS) -- element-granular so arbitrary (element-aligned) strides are accepted
// ---------------------------------------------------------------------------------------------
template <int BYTES> __device__ __forceinline__ void cp_async(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(s), "l"(gmem), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------------------------
// mbarrier + bulk (TMA, non-tensor) copies: cp.async.bulk -> SASS UBLKCP, expect_tx -> SYNCS.ARRIVE.TRANS64
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    unsigned ok;
    const unsigned addr = smem_u32(bar);
    // a long suspend-time hint: the thread sleeps in hardware until the phase completes instead of re-polling
    // (polling warps steal issue slots from the warps they are waiting for)
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(addr), "r"(parity), "r"(0x989680u) : "memory");
    } while (!ok);
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned; completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global bulk copy (TMA store), tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
// waits until at most N of this thread's bulk groups are still READING their shared-memory source
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;\n" ::"n"(N) : "memory"); }

// orders this thread's generic-proxy accesses to shared memory before later async-proxy (bulk copy) writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// shared-memory loads through 32-bit shared addresses (no generic->shared conversion in hot loops)
__device__ __forceinline__ void sts_f32(unsigned addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// predicated store: no branch, no reconvergence point (a lone warp issues in order: a divergent `if` around a store costs a
// BSSY/BSYNC pair and ~20 cycles of branch resolution)
__device__ __forceinline__ void sts_f32_if(unsigned addr, float v, bool pred) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p st.shared.f32 [%0], %1;\n}\n" ::"r"(addr), "f"(v), "r"((unsigned)pred) : "memory");
}
// predicated streaming stores to global memory (no branch around them: see sts_f32_if)
__device__ __forceinline__ void stg_cs_v4_if(void *p, float a, float b, float c, float d, bool pred) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %5, 0;\n@p st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};\n}\n"
                 ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "r"((unsigned)pred) : "memory");
}
__device__ __forceinline__ void stg_s32_if(void *p, int v, bool pred) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p st.global.s32 [%0], %1;\n}\n" ::"l"(p), "r"(v), "r"((unsigned)pred) : "memory");
}
__device__ __forceinline__ float4 lds_v4f32(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds_f32(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(addr));
    return v;
}

// exact power of two that brings m (>= 0, finite) into [1,2): scale = 2^-e.  m == 0 or denormal leaves the block alone.
__device__ __forceinline__ void pow2_fast(float m, float &scale, int &e) {
    const int eb = __float_as_int(m) >> 23;                // m >= 0: no sign bit
    const bool live = eb != 0;
    scale = live ? __int_as_float((254 - eb) << 23) : 1.0f;
    e = live ? eb - 127 : 0;
}

// volatile shared-memory progress counters (producer/consumer rings)
__device__ __forceinline__ int ld_volatile_s32(const int *p) {
    int v;
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.volatile.shared.s32 %0, [%1];\n" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_s32(int *p, int v) {
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("st.volatile.shared.s32 [%0], %1;\n" ::"r"(a), "r"(v) : "memory");
}

// named barrier over `nthreads` threads (ids 1.. ; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

// streaming (evict-first) global stores for state arrays that are written once and read once later
template <typename T> __device__ __forceinline__ void st_stream(T *p, T v) { __stcs(p, v); }

// 8 consecutive elements (32-byte aligned for float) through the streaming store path
__device__ __forceinline__ void store8_stream(float *p, const float (&v)[8]) {
    __stcs(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3]));
    __stcs(reinterpret_cast<float4 *>(p) + 1, make_float4(v[4], v[5], v[6], v[7]));
}
__device__ __forceinline__ void store8_stream(double *p, const double (&v)[8]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) __stcs(reinterpret_cast<double2 *>(p) + q, make_double2(v[2 * q], v[2 * q + 1]));
}

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// One CTA of 256 threads: loss[n] = full[n] - aligned[n] (asg.py:128,136) and the fused tail of asg_step():
// tail[0] = sum of the losses in a FIXED order (thread strides, shuffle tree, 8 partials: run-to-run deterministic),
// tail[1] = the number of utterances, and (asg_forward only) tail[2] = scale * sum, the reduced loss.  An infeasible
// utterance (aligned = -inf) makes the sum +inf, as the reference's own .sum() / .mean() does.
template <typename T>
__device__ __forceinline__ void write_losses_block(const T *full, const T *aligned, T *losses, int N, T *tail, T *red8,
                                                   bool with_reduced = false, T scale = T(1)) {
    T acc = T(0);
    for (int n = threadIdx.x; n < N; n += 256) {
        const T l = full[n] - aligned[n];
        losses[n] = l;
        acc += l;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red8[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        T s = T(0);
#pragma unroll
        for (int q = 0; q < 8; ++q) s += red8[q];
        tail[0] = s;
        tail[1] = (T)N;
        if (with_reduced) tail[2] = scale * s;
    }
}

}  // namespace asg
