// asg_api.cu -- the extern "C" surface of libasg_b200.so (see include/asg_b200.h).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>
#include "asg_common.cuh"
#include "asg_internal.h"

namespace asg {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// ---- live profiler: CUDA-event pairs around every launch, on the launching stream ---------------
namespace {
constexpr int PROF_SLOTS = 8192;
struct ProfRec { cudaEvent_t a, b; int kid; };
std::mutex g_prof_mu;
std::atomic<int> g_prof_on{0};
std::atomic<unsigned> g_prof_mask{0xffffffffu};
std::atomic<unsigned> g_prof_stride{1};
std::vector<ProfRec> g_prof_recs;       // recorded since the last read
std::vector<ProfRec> g_prof_free;       // recycled event pairs
double g_prof_ms[K_COUNT] = {0};
uint64_t g_prof_n[K_COUNT] = {0};
std::atomic<uint64_t> g_kid_launches[K_COUNT];
}  // namespace

LaunchScope::LaunchScope(int kid, cudaStream_t st) : kid_(kid), st_(st), slot_(-1) {
    count_launch();
    g_kid_launches[kid].fetch_add(1, std::memory_order_relaxed);
    const uint64_t seq = g_kid_launches[kid].load(std::memory_order_relaxed);
    if (!g_prof_on.load(std::memory_order_relaxed) || !((g_prof_mask.load(std::memory_order_relaxed) >> kid) & 1u)) return;
    const unsigned stride = g_prof_stride.load(std::memory_order_relaxed);
    if (stride > 1 && (seq % stride) != 0) return;          // time every stride-th launch of a kernel only
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if ((int)g_prof_recs.size() >= PROF_SLOTS) return;
    ProfRec r;
    if (!g_prof_free.empty()) { r = g_prof_free.back(); g_prof_free.pop_back(); }
    else if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    r.kid = kid;
    cudaEventRecord(r.a, st);
    g_prof_recs.push_back(r);
    slot_ = (int)g_prof_recs.size() - 1;
}
LaunchScope::~LaunchScope() {
    if (slot_ < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (slot_ < (int)g_prof_recs.size()) cudaEventRecord(g_prof_recs[slot_].b, st_);
}

static void prof_drain() {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto &r : g_prof_recs) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            g_prof_ms[r.kid] += ms;
            g_prof_n[r.kid] += 1;
        }
        g_prof_free.push_back(r);
    }
    g_prof_recs.clear();
}

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int cuda_fail(int e, const char *where) {
    if (e == 0) return 0;
    return fail(e, "%s: CUDA error %d (%s)", where, e, cudaGetErrorString((cudaError_t)e));
}

static int check_dims(int dtype, int64_t T, int64_t N, int64_t L, int64_t S) {
    if (dtype != ASG_F32 && dtype != ASG_F64) return fail(ASG_E_DTYPE, "dtype must be ASG_F32 or ASG_F64, got %d", dtype);
    if (T <= 0 || N <= 0 || L <= 0 || S <= 0)
        return fail(ASG_E_SHAPE, "T, N, L, S must be positive (got %lld, %lld, %lld, %lld)", (long long)T, (long long)N,
                    (long long)L, (long long)S);
    if (S > T) return fail(ASG_E_SHAPE, "S (%lld) > T (%lld): truncate targets first (asg.py:119-122)", (long long)S, (long long)T);
    if (L > 2048) return fail(ASG_E_UNSUPPORTED, "this build covers L <= 2048 labels (got %lld)", (long long)L);
    // large alphabets run the force-aligned scans in gather mode: the emission ring holds S values per frame, 64 frames deep
    const int64_t s_large = dtype == ASG_F32 ? 768 : 384;
    if (L > ASG_SMALL_L && S > s_large)
        return fail(ASG_E_UNSUPPORTED, "large alphabets (L > 64) are covered for S <= %lld target positions in this dtype (got %lld)",
                    (long long)s_large, (long long)S);
    if (S > 4096) return fail(ASG_E_UNSUPPORTED, "this build covers S <= 4096 target positions (got %lld)", (long long)S);
    if (T > (1 << 24) || N > (1 << 20)) return fail(ASG_E_UNSUPPORTED, "T or N too large");
    return 0;
}

}  // namespace asg

using namespace asg;

template <typename T>
static int scan_typed(int roles, const void *x, int64_t x_st, int64_t x_sn, int64_t x_sl, const void *A, int64_t A_s0,
                      int64_t A_s1, const int64_t *y, int64_t y_s0, int64_t y_s1, const int64_t *il, const int64_t *tl,
                      const asg_layout_t &lay, int64_t Tt, int64_t N, int64_t L, int64_t S, void *state, void *ctx,
                      void *full_scores, void *aligned_scores, cudaStream_t stream) {
    ScanArgs<T> a;
    a.x = (const T *)x; a.x_st = x_st; a.x_sn = x_sn; a.x_sl = x_sl;
    a.A = (const T *)A; a.A_s0 = A_s0; a.A_s1 = A_s1;
    a.y = y; a.y_s0 = y_s0; a.y_s1 = y_s1;
    a.il = il; a.tl = tl;
    a.d = Dims{(int)Tt, (int)N, (int)L, (int)S};
    a.ctx = make_ctx<T>(state, ctx, lay, N);
    a.full_scores = (T *)full_scores; a.aligned_scores = (T *)aligned_scores;
    a.fmax = nullptr;
    a.dir0 = a.dir1 = 0; a.fcc_dirs = a.fac_dirs = 0;
    return launch_scan<T>(a, roles, stream);
}

// what asg_step() adds to a gradient assembly: a uniform upstream gradient by value, and the loss / fused-tail outputs
struct StepExtras {
    double scale = 1.0;
    const void *full = nullptr, *aligned = nullptr;
    void *losses = nullptr, *tail = nullptr;
};

template <typename T>
static int grad_typed(int parts, const void *g_full, int64_t gfs, const void *g_al, int64_t gas, const void *state,
                      const void *ctx, const int64_t *y, int64_t y_s0, int64_t y_s1, const asg_layout_t &lay, int64_t Tt,
                      int64_t N, int64_t L, int64_t S, void *grad_x, void *grad_A, void *scratch, cudaStream_t stream,
                      const StepExtras *ex = nullptr) {
    GradArgs<T> a;
    a.g_full = (const T *)g_full; a.g_full_stride = gfs;
    a.g_aligned = (const T *)g_al; a.g_aligned_stride = gas;
    a.gf_val = ex ? (T)ex->scale : T(1); a.ga_val = ex ? (T)(-ex->scale) : T(1);
    a.full_scores = ex ? (const T *)ex->full : nullptr; a.aligned_scores = ex ? (const T *)ex->aligned : nullptr;
    a.losses = ex ? (T *)ex->losses : nullptr; a.tail = ex ? (T *)ex->tail : nullptr;
    a.ctx = make_ctx<T>(const_cast<void *>(state), const_cast<void *>(ctx), lay, N);
    a.y = y; a.y_s0 = y_s0; a.y_s1 = y_s1;
    a.d = Dims{(int)Tt, (int)N, (int)L, (int)S};
    a.grad_x = (T *)grad_x; a.grad_A = (T *)grad_A;
    a.partials = (T *)scratch;
    a.tb = 0; a.nstage = 0; a.nwarps = 0; a.dry = 0; a.tiles_t = 0; a.num_tiles = 0;
    if (L > ASG_SMALL_L) return launch_grad_large<T>(a, parts, stream);
    return launch_grad<T>(a, parts, stream);
}

struct asg_host_ctx {
    int dtype, device;
    int64_t T, N, L, S;
    asg_layout_t lay;
    cudaStream_t stream;                // this context's copies (H2D in, D2H out) and what asg_host_wait synchronises
    cudaEvent_t ev_in, ev_out;          // inputs landed / results ready: the hand-over to and from the device's compute stream
    char *dev;          // one device slab
    size_t dev_bytes;
    // offsets into the slab
    size_t o_x, o_A, o_y, o_il, o_tl, o_g, o_gneg, o_state, o_ctx, o_full, o_al, o_loss, o_gx, o_gA, o_scr;
};

// The kernels of ALL host contexts of a device run on ONE compute stream, in submission order; only the copies use the
// context's own stream.  Kernels of different contexts barely co-reside anyway (the gradient kernel takes every SM's shared
// memory, the scans want another carve-out), and left to eight independent streams they interleave badly (0.30 ms per step
// at cfg-2 against 0.245 ms for the same kernels back to back); serialised here, the copies of neighbouring steps hide behind
// them.  The mutex covers the enqueue only (asg_scan's fork/join events belong to the (device, stream) pair).
static std::mutex g_compute_mu[64];
static cudaStream_t g_compute_stream[64] = {nullptr};
static cudaStream_t compute_stream_locked(int device) {          // call with g_compute_mu[device] held
    if (!g_compute_stream[device]) cudaStreamCreateWithFlags(&g_compute_stream[device], cudaStreamNonBlocking);
    return g_compute_stream[device];
}

template <typename T> __global__ void host_loss_kernel(const T *full, const T *aligned, T *loss, int N) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N) loss[n] = full[n] - aligned[n];
}
template <typename T> __global__ void host_g_kernel(const T *g, T *gpos, T *gneg, int N, int have_g) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N) { T v = have_g ? g[n] : T(1); gpos[n] = v; gneg[n] = -v; }
}

template <typename T> __global__ void __launch_bounds__(256) losses_kernel(const T *full, const T *aligned, T *losses, int N, T *tail, T scale) {
    __shared__ T red8[8];
    write_losses_block<T>(full, aligned, losses, N, tail, red8, true, scale);
}

// Fused log_softmax + ASG, backward side.  loss(log_softmax(z)) == loss(z) (the FCC term IS the per-frame normaliser: both scores
// shift by the same sum of log-partition values) and d loss / d z = gx - softmax(z) * sum_l gx, where sum_l gx[t,n,:] is exactly 0
// for every utterance with a valid alignment (the two posteriors cancel) and g_n for an infeasible one (aligned = -inf: its FAC
// gradients are exactly 0, force_aligned_lattice.cpp:164-168 masked softmax).  So only infeasible utterances need a correction:
// one CTA per utterance, which leaves at once unless its aligned score is -inf; one warp per frame otherwise.
template <typename T>
__global__ void __launch_bounds__(256) logits_fix_kernel(const T *z, int64_t z_st, int64_t z_sn, int64_t z_sl, const T *aligned, const T *g,
                                                         int64_t g_stride, T scale, const int64_t *il, int Tt, int N, int L, T *grad_x) {
    const int n = blockIdx.x;
    const T al = aligned[n];
    if (al > Num<T>::neg_inf()) return;
    const T gn = g ? scale * g[(int64_t)n * g_stride] : scale;
    int Tn = il ? (int)il[n] : Tt;
    Tn = max(0, min(Tn, Tt));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int t = warp; t < Tn; t += 8) {
        const T *row = z + (int64_t)t * z_st + (int64_t)n * z_sn;
        T m = Num<T>::neg_inf();
        for (int l = lane; l < L; l += 32) m = tmax(m, row[(int64_t)l * z_sl]);
        m = Num<T>::warp_max(m);
        if (!(m > Num<T>::neg_inf())) continue;
        T sum = T(0);
        for (int l = lane; l < L; l += 32) sum += Num<T>::exp_(row[(int64_t)l * z_sl] - m);
        sum = warp_sum(sum);
        T *gx = grad_x + ((int64_t)t * N + n) * L;
        for (int l = lane; l < L; l += 32) gx[l] -= gn * (Num<T>::exp_(row[(int64_t)l * z_sl] - m) / sum);
    }
}

__global__ void range_flags_kernel(const double *scalars, int N, int *out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N) out[n] = *reinterpret_cast<const int *>(scalars + (size_t)n * 8 + SC_FLAGS);
}

// The gradient assembly's view of the stored force-aligned states, re-read on demand: flag the utterances in which a block of
// target positions may hold flushed entries that matter (RF_POSTERIOR_THIN; the criterion is fac_block_at_risk()).  One warp per
// frame, the frame's blocks strided over the lanes, two passes (largest product of the frame, then the check).
template <typename T>
__global__ void __launch_bounds__(128) range_check_kernel(Ctx<T> ctx, int N, int *out) {
    const int n = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Tn = (int)ctx.il[n], Sn = (int)ctx.tl[n];
    if (Sn > Tn || Tn <= 0) return;                         // infeasible: no posterior
    const int nblk = (Sn + 7) / 8;
    bool risk = false;
    for (int t = blockIdx.x * 4 + warp; t < Tn; t += gridDim.x * 4) {
        const int64_t row = (int64_t)n * ctx.tp + t;
        const T *pa = ctx.fac_alpha(row), *pb = ctx.fac_beta(row);
        const int *wa_ = ctx.fac_ea(row), *wb_ = ctx.fac_eb(row);
        int cmax = INT_MIN;
        for (int b = lane; b < nblk; b += 32) {
            int ea, eb, da, db;
            T fa, fb;
            fac_unpack_norm<T>(wa_[b], ea, da, fa);
            fac_unpack_norm<T>(wb_[b], eb, db, fb);
            T pm = T(0);
#pragma unroll
            for (int j = 0; j < 8; ++j) pm = tmax(pm, (pa[b * 8 + j] * fa) * (pb[b * 8 + j] * fb));
            if (pm > T(0)) cmax = max(cmax, ea + eb + Num<T>::pexp(pm));
        }
        cmax = __reduce_max_sync(ASG_FULL_MASK, cmax);
        if (cmax == INT_MIN) continue;
        for (int b = lane; b < nblk; b += 32) {
            const int wa = wa_[b], wb = wb_[b];
            if (!fac_live(wa) || !fac_live(wb)) continue;
            int ea, eb, da, db;
            T fa, fb;
            fac_unpack_norm<T>(wa, ea, da, fa);
            fac_unpack_norm<T>(wb, eb, db, fb);
            const int c = ea + eb;
            if (cmax - c >= Num<T>::kSafeGap + max(max(da, db), 0)) continue;        // coarse test clears almost every block
            T al[8], be[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { al[j] = pa[b * 8 + j]; be[j] = pb[b * 8 + j]; }
            risk |= fac_block_at_risk<T>(al, be, fa, fb, da, db, c, cmax);
        }
    }
    if (__any_sync(ASG_FULL_MASK, risk) && lane == 0) atomicOr(out + n, RF_POSTERIOR_THIN);
}

extern "C" {

int asg_layout(int dtype, int64_t T, int64_t N, int64_t L, int64_t S, asg_layout_t *out) {
    if (!out) return fail(ASG_E_NULL, "asg_layout: out is NULL");
    int rc = check_dims(dtype, T, N, L, S);
    if (rc) return rc;
    memset(out, 0, sizeof(*out));
    const size_t eb = dtype == ASG_F32 ? 4 : 8;
    out->elem_bytes = eb;
    out->sp = (S + ASG_BLK - 1) / ASG_BLK * ASG_BLK;
    out->nblk = out->sp / ASG_BLK;
    out->nbp = (out->nblk + 7) / 8 * 8;
    out->lpi = (L + 7) / 8 * 8;
    out->tp = (T + 3) / 4 * 4;
    {   // the frame record: every part starts on a 32-byte boundary
        size_t r = 0;
        auto part = [&](size_t bytes) { size_t o = r; r = align_up(r + bytes, 32); return o; };
        out->rec_fac_alpha = part(eb * (size_t)out->sp);
        out->rec_fac_beta = part(eb * (size_t)out->sp);
        out->rec_fac_ea = part(sizeof(int) * (size_t)out->nbp);
        out->rec_fac_eb = part(sizeof(int) * (size_t)out->nbp);
        out->rec_fcc_alpha = part(eb * (size_t)out->lpi);
        out->rec_fcc_beta = part(eb * (size_t)out->lpi);
        out->rec_fcc_w = part(eb * (size_t)out->lpi);
        out->rec_fcc_scale = part(eb);
        out->rec_bytes = r;
    }
    out->state_bytes = (size_t)N * out->tp * out->rec_bytes;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    out->off_ehat = take(eb * (size_t)L * L);
    out->off_fac_hd = take(eb * 2 * (size_t)N * out->sp);
    out->off_lengths = take(sizeof(int64_t) * 2 * (size_t)N);
    out->off_scalars = take(sizeof(double) * 8 * (size_t)N);
    out->off_ehat_t = out->off_ehat;
    if (L > ASG_SMALL_L) {
        out->off_ehat_t = take(eb * (size_t)L * L);
        out->off_fmax = take(eb * (size_t)N * out->tp);
        out->off_rmax = take(eb);
        out->off_large_ws = take(large_ws_bytes(eb, N, L));
    }
    out->ctx_bytes = off;
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) device = -1;
    out->grad_scratch_bytes = align_up(eb * (size_t)grad_grid_size(device) * L * L, 256);
    if (L > ASG_SMALL_L)   // G [L,L] + the FAC edge sums [2,N,sp] in 64-bit fixed point
        out->grad_scratch_bytes = align_up(eb * (size_t)large_splitk_host((int)L) * L * L, 256) + align_up(16 * (size_t)N * out->sp, 256) +
                                  align_up(8 * (size_t)L * L, 256);
    return 0;
}

int asg_scan(int dtype, int roles, const void *x, int64_t x_st, int64_t x_sn, int64_t x_sl, const void *A, int64_t A_s0,
             int64_t A_s1, const int64_t *y, int64_t y_s0, int64_t y_s1, const int64_t *il, const int64_t *tl, int64_t T,
             int64_t N, int64_t L, int64_t S, void *state, void *ctx, void *full_scores, void *aligned_scores,
             asg_stream_t stream) {
    asg_layout_t lay;
    int rc = asg_layout(dtype, T, N, L, S, &lay);
    if (rc) return rc;
    if ((roles & ~ASG_ROLE_ALL) || roles == 0) return fail(ASG_E_SHAPE, "asg_scan: bad roles mask %d", roles);
    if (!x || !A || !ctx) return fail(ASG_E_NULL, "asg_scan: x, A and ctx are required");
    if ((roles & (ASG_ROLE_FCC_ALPHA | ASG_ROLE_FAC_ALPHA)) && !state) return fail(ASG_E_NULL, "asg_scan: state is NULL (only the beta scans can run without it)");
    if ((roles & (ASG_ROLE_FAC_ALPHA | ASG_ROLE_FAC_BETA)) && !y) return fail(ASG_E_NULL, "asg_scan: y is NULL");
    int e;
    if (dtype == ASG_F32)
        e = scan_typed<float>(roles, x, x_st, x_sn, x_sl, A, A_s0, A_s1, y, y_s0, y_s1, il, tl, lay, T, N, L, S, state, ctx,
                              full_scores, aligned_scores, (cudaStream_t)stream);
    else
        e = scan_typed<double>(roles, x, x_st, x_sn, x_sl, A, A_s0, A_s1, y, y_s0, y_s1, il, tl, lay, T, N, L, S, state, ctx,
                               full_scores, aligned_scores, (cudaStream_t)stream);
    return cuda_fail(e, "asg_scan");
}

int asg_grad(int dtype, int parts, const void *g_full, int64_t g_full_stride, const void *g_aligned,
             int64_t g_aligned_stride, const void *state, const void *ctx, const int64_t *y, int64_t y_s0, int64_t y_s1,
             int64_t T, int64_t N, int64_t L, int64_t S, void *grad_x, void *grad_A, void *scratch, asg_stream_t stream) {
    asg_layout_t lay;
    int rc = asg_layout(dtype, T, N, L, S, &lay);
    if (rc) return rc;
    if (parts < 1 || parts > 3) return fail(ASG_E_SHAPE, "asg_grad: bad parts mask %d", parts);
    if (!state || !ctx || !grad_x || !grad_A || !scratch) return fail(ASG_E_NULL, "asg_grad: state, ctx, grad_x, grad_A, scratch are required");
    if ((parts & ASG_PART_FCC) && !g_full) return fail(ASG_E_NULL, "asg_grad: g_full is NULL");
    if ((parts & ASG_PART_FAC) && (!g_aligned || !y)) return fail(ASG_E_NULL, "asg_grad: FAC inputs missing");
    int e;
    if (dtype == ASG_F32)
        e = grad_typed<float>(parts, g_full, g_full_stride, g_aligned, g_aligned_stride, state, ctx, y, y_s0, y_s1, lay, T, N, L, S,
                              grad_x, grad_A, scratch, (cudaStream_t)stream);
    else
        e = grad_typed<double>(parts, g_full, g_full_stride, g_aligned, g_aligned_stride, state, ctx, y, y_s0, y_s1, lay, T, N, L, S,
                               grad_x, grad_A, scratch, (cudaStream_t)stream);
    return cuda_fail(e, "asg_grad");
}

int asg_forward(int dtype, int forward_only, const void *x, int64_t x_st, int64_t x_sn, int64_t x_sl, const void *A, int64_t A_s0,
                int64_t A_s1, const int64_t *y, int64_t y_s0, int64_t y_s1, const int64_t *il, const int64_t *tl, int64_t T,
                int64_t N, int64_t L, int64_t S, double scale, void *state, void *ctx, void *full_scores, void *aligned_scores,
                void *losses, void *loss_tail, asg_stream_t stream) {
    if (!full_scores || !aligned_scores || !losses || !loss_tail)
        return fail(ASG_E_NULL, "asg_forward: full_scores, aligned_scores, losses and loss_tail are required");
    int rc = asg_scan(dtype, forward_only ? ASG_ROLE_FORWARD_ONLY : ASG_ROLE_ALL, x, x_st, x_sn, x_sl, A, A_s0, A_s1, y, y_s0, y_s1,
                      il, tl, T, N, L, S, forward_only ? nullptr : state, ctx, full_scores, aligned_scores, stream);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    {
        LaunchScope ls(K_HOST_MISC, st);
        if (dtype == ASG_F32) losses_kernel<float><<<1, 256, 0, st>>>((const float *)full_scores, (const float *)aligned_scores, (float *)losses, (int)N, (float *)loss_tail, (float)scale);
        else losses_kernel<double><<<1, 256, 0, st>>>((const double *)full_scores, (const double *)aligned_scores, (double *)losses, (int)N, (double *)loss_tail, scale);
    }
    return cuda_fail((int)cudaGetLastError(), "asg_forward");
}

int asg_backward(int dtype, const void *g, int64_t g_stride, double scale, const void *state, const void *ctx, const int64_t *y,
                 int64_t y_s0, int64_t y_s1, int64_t T, int64_t N, int64_t L, int64_t S, void *grad_x, void *grad_A, void *scratch,
                 asg_stream_t stream) {
    asg_layout_t lay;
    int rc = asg_layout(dtype, T, N, L, S, &lay);
    if (rc) return rc;
    if (!state || !ctx || !y || !grad_x || !grad_A || !scratch) return fail(ASG_E_NULL, "asg_backward: state, ctx, y, grad_x, grad_A, scratch are required");
    StepExtras ex;
    ex.scale = scale;
    int e;
    if (dtype == ASG_F32)
        e = grad_typed<float>(ASG_PART_BOTH, g, g_stride, g, g_stride, state, ctx, y, y_s0, y_s1, lay, T, N, L, S, grad_x, grad_A, scratch,
                              (cudaStream_t)stream, &ex);
    else
        e = grad_typed<double>(ASG_PART_BOTH, g, g_stride, g, g_stride, state, ctx, y, y_s0, y_s1, lay, T, N, L, S, grad_x, grad_A, scratch,
                               (cudaStream_t)stream, &ex);
    return cuda_fail(e, "asg_backward");
}

int asg_logits_backward_fix(int dtype, const void *z, int64_t z_st, int64_t z_sn, int64_t z_sl, const void *aligned_scores, const void *g,
                            int64_t g_stride, double scale, const int64_t *il, int64_t T, int64_t N, int64_t L, void *grad_x,
                            asg_stream_t stream) {
    if (dtype != ASG_F32 && dtype != ASG_F64) return fail(ASG_E_DTYPE, "dtype must be ASG_F32 or ASG_F64, got %d", dtype);
    if (!z || !aligned_scores || !grad_x) return fail(ASG_E_NULL, "asg_logits_backward_fix: z, aligned_scores and grad_x are required");
    if (T <= 0 || N <= 0 || L <= 0) return fail(ASG_E_SHAPE, "asg_logits_backward_fix: T, N, L must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    {
        LaunchScope ls(K_HOST_MISC, st);
        if (dtype == ASG_F32)
            logits_fix_kernel<float><<<(unsigned)N, 256, 0, st>>>((const float *)z, z_st, z_sn, z_sl, (const float *)aligned_scores, (const float *)g,
                                                                   g_stride, (float)scale, il, (int)T, (int)N, (int)L, (float *)grad_x);
        else
            logits_fix_kernel<double><<<(unsigned)N, 256, 0, st>>>((const double *)z, z_st, z_sn, z_sl, (const double *)aligned_scores, (const double *)g,
                                                                    g_stride, scale, il, (int)T, (int)N, (int)L, (double *)grad_x);
    }
    return cuda_fail((int)cudaGetLastError(), "asg_logits_backward_fix");
}

int asg_range_flags(int dtype, const void *state, const void *ctx, int64_t T, int64_t N, int64_t L, int64_t S, int32_t *flags,
                    asg_stream_t stream) {
    asg_layout_t lay;
    int rc = asg_layout(dtype, T, N, L, S, &lay);
    if (rc) return rc;
    if (!ctx || !flags) return fail(ASG_E_NULL, "asg_range_flags: ctx and flags are required");
    const double *sc = reinterpret_cast<const double *>(static_cast<const char *>(ctx) + lay.off_scalars);
    range_flags_kernel<<<(unsigned)((N + 127) / 128), 128, 0, (cudaStream_t)stream>>>(sc, (int)N, flags);
    if (state) {
        const dim3 grid((unsigned)std::min<int64_t>((T + 3) / 4, 64), (unsigned)N);
        if (dtype == ASG_F32)
            range_check_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>(make_ctx<float>(const_cast<void *>(state), const_cast<void *>(ctx), lay, N), (int)N, flags);
        else
            range_check_kernel<double><<<grid, 128, 0, (cudaStream_t)stream>>>(make_ctx<double>(const_cast<void *>(state), const_cast<void *>(ctx), lay, N), (int)N, flags);
    }
    return cuda_fail((int)cudaGetLastError(), "asg_range_flags");
}

int asg_pack_replabels(const int64_t *y, int64_t y_s0, int64_t y_s1, const int64_t *tl, int64_t N, int64_t S, int64_t first_rep_label,
                       int64_t max_reps, int64_t fill, int64_t *out, int64_t *out_len, asg_stream_t stream) {
    if (!y || !out || !out_len) return fail(ASG_E_NULL, "asg_pack_replabels: y, out and out_len are required");
    if (N <= 0 || S <= 0 || max_reps < 0) return fail(ASG_E_SHAPE, "asg_pack_replabels: N, S must be positive and max_reps >= 0");
    if (S > 8192) return fail(ASG_E_UNSUPPORTED, "asg_pack_replabels covers S <= 8192 (got %lld)", (long long)S);
    return cuda_fail(replabel_dispatch(true, y, y_s0, y_s1, tl, N, S, first_rep_label, max_reps, fill, out, S, out_len, (cudaStream_t)stream),
                     "asg_pack_replabels");
}

int asg_unpack_replabels(const int64_t *y, int64_t y_s0, int64_t y_s1, const int64_t *tl, int64_t N, int64_t S, int64_t first_rep_label,
                         int64_t max_reps, int64_t fill, int64_t *out, int64_t S_out, int64_t *out_len, asg_stream_t stream) {
    if (!y || !out || !out_len) return fail(ASG_E_NULL, "asg_unpack_replabels: y, out and out_len are required");
    if (N <= 0 || S <= 0 || S_out <= 0 || max_reps < 0) return fail(ASG_E_SHAPE, "asg_unpack_replabels: N, S, S_out must be positive and max_reps >= 0");
    if (S + S_out > 16384) return fail(ASG_E_UNSUPPORTED, "asg_unpack_replabels covers S + S_out <= 16384 (got %lld)", (long long)(S + S_out));
    return cuda_fail(replabel_dispatch(false, y, y_s0, y_s1, tl, N, S, first_rep_label, max_reps, fill, out, S_out, out_len, (cudaStream_t)stream),
                     "asg_unpack_replabels");
}

size_t asg_viterbi_scratch_bytes(int64_t T, int64_t N, int64_t S) {
    if (T <= 0 || N <= 0 || S <= 0) return 0;
    return align_up(sizeof(unsigned) * (size_t)N * T * viterbi_words(S), 256);
}

int asg_viterbi(int dtype, const void *x, int64_t x_st, int64_t x_sn, int64_t x_sl, const void *A, int64_t A_s0, int64_t A_s1, const int64_t *y,
                int64_t y_s0, int64_t y_s1, const int64_t *il, const int64_t *tl, int64_t T, int64_t N, int64_t L, int64_t S, void *scores,
                int64_t *path, void *scratch, asg_stream_t stream) {
    if (dtype != ASG_F32 && dtype != ASG_F64) return fail(ASG_E_DTYPE, "dtype must be ASG_F32 or ASG_F64, got %d", dtype);
    if (T <= 0 || N <= 0 || L <= 0 || S <= 0) return fail(ASG_E_SHAPE, "asg_viterbi: T, N, L, S must be positive");
    if (S > 4096) return fail(ASG_E_UNSUPPORTED, "asg_viterbi covers S <= 4096 target positions (got %lld)", (long long)S);
    if (!x || !A || !y || !scores || !path || !scratch) return fail(ASG_E_NULL, "asg_viterbi: x, A, y, scores, path and scratch are required");
    return cuda_fail(viterbi_dispatch(dtype, x, x_st, x_sn, x_sl, A, A_s0, A_s1, y, y_s0, y_s1, il, tl, T, N, L, S, scores, path, scratch,
                                      (cudaStream_t)stream), "asg_viterbi");
}

int asg_step(int dtype, const void *x, int64_t x_st, int64_t x_sn, int64_t x_sl, const void *A, int64_t A_s0, int64_t A_s1,
             const int64_t *y, int64_t y_s0, int64_t y_s1, const int64_t *il, const int64_t *tl, int64_t T, int64_t N, int64_t L,
             int64_t S, double scale, void *state, void *ctx, void *scratch, void *full_scores, void *aligned_scores,
             void *losses, void *grad_x, void *fused, asg_stream_t stream) {
    asg_layout_t lay;
    int rc = asg_layout(dtype, T, N, L, S, &lay);
    if (rc) return rc;
    if (!x || !A || !y || !state || !ctx || !scratch || !full_scores || !aligned_scores || !losses || !grad_x || !fused)
        return fail(ASG_E_NULL, "asg_step: every buffer is required (only il / tl may be NULL)");
    rc = asg_scan(dtype, ASG_ROLE_ALL, x, x_st, x_sn, x_sl, A, A_s0, A_s1, y, y_s0, y_s1, il, tl, T, N, L, S, state, ctx,
                  full_scores, aligned_scores, stream);
    if (rc) return rc;
    StepExtras ex;
    ex.scale = scale; ex.full = full_scores; ex.aligned = aligned_scores; ex.losses = losses;
    ex.tail = (char *)fused + lay.elem_bytes * (size_t)L * L;
    int e;
    if (dtype == ASG_F32)
        e = grad_typed<float>(ASG_PART_BOTH, nullptr, 0, nullptr, 0, state, ctx, y, y_s0, y_s1, lay, T, N, L, S, grad_x, fused, scratch,
                              (cudaStream_t)stream, &ex);
    else
        e = grad_typed<double>(ASG_PART_BOTH, nullptr, 0, nullptr, 0, state, ctx, y, y_s0, y_s1, lay, T, N, L, S, grad_x, fused, scratch,
                               (cudaStream_t)stream, &ex);
    return cuda_fail(e, "asg_step");
}

// ---------------------------------------------------------------------------------------------
// host-buffer API
// ---------------------------------------------------------------------------------------------
int asg_host_create(int dtype, int device, int64_t T, int64_t N, int64_t L, int64_t S, asg_host_ctx_t **out) {
    if (!out) return fail(ASG_E_NULL, "asg_host_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev)
        return fail(ASG_E_NOGPU, "asg_host_create: CUDA device %d is not available (no CPU fallback exists)", device);
    int e = cudaSetDevice(device);
    if (e) return cuda_fail(e, "cudaSetDevice");
    asg_host_ctx *h = new (std::nothrow) asg_host_ctx();
    if (!h) return fail(ASG_E_NULL, "out of host memory");
    h->dtype = dtype; h->device = device; h->T = T; h->N = N; h->L = L; h->S = S;
    int rc = asg_layout(dtype, T, N, L, S, &h->lay);
    if (rc) { delete h; return rc; }
    const size_t eb = h->lay.elem_bytes;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    h->o_x = take(eb * (size_t)T * N * L);
    h->o_A = take(eb * (size_t)L * L);
    h->o_y = take(8 * (size_t)N * S);
    h->o_il = take(8 * (size_t)N);
    h->o_tl = take(8 * (size_t)N);
    h->o_g = take(eb * (size_t)N);
    h->o_gneg = take(eb * (size_t)N);
    h->o_state = take(h->lay.state_bytes);
    h->o_ctx = take(h->lay.ctx_bytes);
    h->o_full = take(eb * (size_t)N);
    h->o_al = take(eb * (size_t)N);
    h->o_loss = take(eb * (size_t)N);
    h->o_gx = take(eb * (size_t)T * N * L);
    h->o_gA = take(eb * (size_t)L * L);
    h->o_scr = take(h->lay.grad_scratch_bytes);
    h->dev_bytes = off;
    if ((e = cudaMalloc((void **)&h->dev, off))) { delete h; return cuda_fail(e, "cudaMalloc"); }
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking))) { cudaFree(h->dev); delete h; return cuda_fail(e, "cudaStreamCreate"); }
    if ((e = cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming)) || (e = cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming))) {
        cudaStreamDestroy(h->stream); cudaFree(h->dev); delete h;
        return cuda_fail(e, "cudaEventCreate");
    }
    *out = h;
    return 0;
}

void asg_host_destroy(asg_host_ctx_t *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    cudaEventDestroy(h->ev_in);
    cudaEventDestroy(h->ev_out);
    cudaStreamDestroy(h->stream);
    cudaFree(h->dev);
    delete h;
}

int asg_host_submit(asg_host_ctx_t *h, const void *x, const void *A, const int64_t *y, const int64_t *il, const int64_t *tl,
                    const void *g, int64_t T, int64_t N, int64_t L, int64_t S, void *loss, void *grad_x, void *grad_A) {
    if (!h || !x || !A || !y || !loss) return fail(ASG_E_NULL, "asg_host_submit: h, x, A, y, loss are required");
    if (T != h->T || N != h->N || L != h->L || S != h->S)
        return fail(ASG_E_SHAPE, "asg_host_submit: shape differs from the one the context was created for");
    int e = cudaSetDevice(h->device);
    if (e) return cuda_fail(e, "cudaSetDevice");
    const size_t eb = h->lay.elem_bytes;
    cudaStream_t cs = h->stream;
    char *d = h->dev;
    const bool want_grad = grad_x != nullptr || grad_A != nullptr;
    if (h->device < 0 || h->device >= 64) return fail(ASG_E_UNSUPPORTED, "asg_host_submit: device index %d out of range", h->device);
#define ASG_TRY(call, what) do { if ((e = (int)(call))) return cuda_fail(e, what); } while (0)
    ASG_TRY(cudaMemcpyAsync(d + h->o_x, x, eb * (size_t)T * N * L, cudaMemcpyHostToDevice, cs), "H2D x");
    ASG_TRY(cudaMemcpyAsync(d + h->o_A, A, eb * (size_t)L * L, cudaMemcpyHostToDevice, cs), "H2D A");
    ASG_TRY(cudaMemcpyAsync(d + h->o_y, y, 8 * (size_t)N * S, cudaMemcpyHostToDevice, cs), "H2D y");
    if (il) ASG_TRY(cudaMemcpyAsync(d + h->o_il, il, 8 * (size_t)N, cudaMemcpyHostToDevice, cs), "H2D il");
    if (tl) ASG_TRY(cudaMemcpyAsync(d + h->o_tl, tl, 8 * (size_t)N, cudaMemcpyHostToDevice, cs), "H2D tl");
    if (g && want_grad) ASG_TRY(cudaMemcpyAsync(d + h->o_g, g, eb * (size_t)N, cudaMemcpyHostToDevice, cs), "H2D g");
    ASG_TRY(cudaEventRecord(h->ev_in, cs), "cudaEventRecord");
    {
        std::lock_guard<std::mutex> lock(g_compute_mu[h->device]);
        cudaStream_t st = compute_stream_locked(h->device);
        ASG_TRY(cudaStreamWaitEvent(st, h->ev_in, 0), "cudaStreamWaitEvent");
        const int roles = want_grad ? ASG_ROLE_ALL : ASG_ROLE_FORWARD_ONLY;
        int rc = asg_scan(h->dtype, roles, d + h->o_x, N * L, L, 1, d + h->o_A, L, 1, (const int64_t *)(d + h->o_y), S, 1,
                          il ? (const int64_t *)(d + h->o_il) : nullptr, tl ? (const int64_t *)(d + h->o_tl) : nullptr, T, N, L, S,
                          want_grad ? d + h->o_state : nullptr, d + h->o_ctx, d + h->o_full, d + h->o_al, st);
        if (rc) return rc;
        const int nb = (int)((N + 127) / 128);
        {
            LaunchScope ls(K_HOST_MISC, st);
            if (h->dtype == ASG_F32) host_loss_kernel<float><<<nb, 128, 0, st>>>((float *)(d + h->o_full), (float *)(d + h->o_al), (float *)(d + h->o_loss), (int)N);
            else host_loss_kernel<double><<<nb, 128, 0, st>>>((double *)(d + h->o_full), (double *)(d + h->o_al), (double *)(d + h->o_loss), (int)N);
        }
        ASG_TRY(cudaGetLastError(), "host_loss_kernel");
        if (want_grad) {
            {
                LaunchScope ls(K_HOST_MISC, st);
                if (h->dtype == ASG_F32) host_g_kernel<float><<<nb, 128, 0, st>>>((float *)(d + h->o_g), (float *)(d + h->o_g), (float *)(d + h->o_gneg), (int)N, g != nullptr);
                else host_g_kernel<double><<<nb, 128, 0, st>>>((double *)(d + h->o_g), (double *)(d + h->o_g), (double *)(d + h->o_gneg), (int)N, g != nullptr);
            }
            ASG_TRY(cudaGetLastError(), "host_g_kernel");
            rc = asg_grad(h->dtype, ASG_PART_BOTH, d + h->o_g, 1, d + h->o_gneg, 1, d + h->o_state, d + h->o_ctx,
                          (const int64_t *)(d + h->o_y), S, 1, T, N, L, S, d + h->o_gx, d + h->o_gA, d + h->o_scr, st);
            if (rc) return rc;
        }
        ASG_TRY(cudaEventRecord(h->ev_out, st), "cudaEventRecord");
    }
    ASG_TRY(cudaStreamWaitEvent(cs, h->ev_out, 0), "cudaStreamWaitEvent");
    ASG_TRY(cudaMemcpyAsync(loss, d + h->o_loss, eb * (size_t)N, cudaMemcpyDeviceToHost, cs), "D2H loss");
    if (want_grad) {
        if (grad_x) ASG_TRY(cudaMemcpyAsync(grad_x, d + h->o_gx, eb * (size_t)T * N * L, cudaMemcpyDeviceToHost, cs), "D2H grad_x");
        if (grad_A) ASG_TRY(cudaMemcpyAsync(grad_A, d + h->o_gA, eb * (size_t)L * L, cudaMemcpyDeviceToHost, cs), "D2H grad_A");
    }
#undef ASG_TRY
    return 0;
}

int asg_host_wait(asg_host_ctx_t *h) {
    if (!h) return fail(ASG_E_NULL, "asg_host_wait: h is NULL");
    int e = cudaSetDevice(h->device);
    if (e) return cuda_fail(e, "cudaSetDevice");
    return cuda_fail((int)cudaStreamSynchronize(h->stream), "cudaStreamSynchronize");
}

int asg_host_loss(asg_host_ctx_t *h, const void *x, const void *A, const int64_t *y, const int64_t *il, const int64_t *tl,
                  const void *g, int64_t T, int64_t N, int64_t L, int64_t S, void *loss, void *grad_x, void *grad_A) {
    int rc = asg_host_submit(h, x, A, y, il, tl, g, T, N, L, S, loss, grad_x, grad_A);
    if (rc) return rc;
    return asg_host_wait(h);
}

void asg_release(void) {
    release_side_slots();
    for (int dev = 0; dev < 64; ++dev) {
        std::lock_guard<std::mutex> lock(g_compute_mu[dev]);
        if (g_compute_stream[dev]) {
            int cur = 0;
            cudaGetDevice(&cur);
            cudaSetDevice(dev);
            cudaStreamDestroy(g_compute_stream[dev]);
            g_compute_stream[dev] = nullptr;
            cudaSetDevice(cur);
        }
    }
}
const char *asg_last_error(void) { return g_err; }
const char *asg_version(void) { return "asg_b200 0.1 (sm_100a)"; }
uint64_t asg_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int asg_profile_enable(int on) {
    if (!on) prof_drain();
    if (on) {   // create the event pairs up front: cudaEventCreate costs tens of microseconds and must stay out of timed loops
        std::lock_guard<std::mutex> lk(g_prof_mu);
        while ((int)(g_prof_free.size() + g_prof_recs.size()) < PROF_SLOTS) {
            ProfRec r;
            if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) break;
            r.kid = 0;
            g_prof_free.push_back(r);
        }
    }
    g_prof_on.store(on ? 1 : 0);
    return 0;
}
int asg_profile_select(unsigned mask) {
    g_prof_mask.store(mask);
    return 0;
}
int asg_profile_stride(unsigned stride) { g_prof_stride.store(stride ? stride : 1, std::memory_order_relaxed); return ASG_OK; }
int asg_profile_reset(void) {
    prof_drain();
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int k = 0; k < K_COUNT; ++k) { g_prof_ms[k] = 0; g_prof_n[k] = 0; }
    return 0;
}
int asg_profile_read(int kernel_id, double *total_ms, uint64_t *launches) {
    if (kernel_id < 0 || kernel_id >= K_COUNT) return fail(ASG_E_SHAPE, "asg_profile_read: bad kernel id %d", kernel_id);
    prof_drain();
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (total_ms) *total_ms = g_prof_ms[kernel_id];
    if (launches) *launches = g_prof_n[kernel_id];
    return 0;
}
const char *asg_kernel_name(int kernel_id) {
    static const char *names[K_COUNT] = {"fcc_scan_kernel", "fac_scan_kernel", "grad_kernel", "grad_finalize_kernel", "host_misc", "grad_large_gemm_kernel", "peer_allreduce_kernel"};
    return (kernel_id >= 0 && kernel_id < K_COUNT) ? names[kernel_id] : "";
}
int asg_kernel_count(void) { return K_COUNT; }

}  // extern "C"
