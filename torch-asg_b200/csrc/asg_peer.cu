// asg_peer.cu -- the one exchange step of the batch-sharded path written over NVLink peer memory.
//
// A batch-sharded ASG step has exactly one exchange: the sum over ranks of the fused vector [L*L transition gradient |
// loss sum | utterance count] -- 6.4 KB at L = 40.  At that size a collective is pure latency, and a library collective
// captured in a CUDA graph also costs the HOST ~0.1 ms per replay (measured, profiles/), which is what capped the 2..8 GPU
// scaling of round 1.  Here the exchange is ONE single-CTA kernel per rank:
//
//   push   every rank stores its vector into slot [parity][rank] of EVERY rank's exchange buffer (plain stores over
//          NVLink / NVSwitch; the buffers are cudaMalloc'ed by this library and opened on the peers through CUDA IPC),
//          fences at system scope, then raises flag [parity][rank] = epoch on every rank (st.release.sys);
//   wait   `world` threads spin (ld.acquire.sys) on the LOCAL flags until every rank's epoch has arrived;
//   sum    the `world` slots are added in rank order -- every rank computes the bit-identical sum -- into the caller's
//          buffer, in place.
//
// The epoch is a counter in the rank's own buffer, advanced by the kernel itself, so the kernel takes no per-call argument
// and replays inside a CUDA graph.  Slots are double-buffered by epoch parity: a rank can only be one exchange ahead of
// the slowest rank (it cannot pass the wait of exchange k+1 before everyone pushed k+1, i.e. finished reading k).
// A rank that waits longer than ASG_PEER_TIMEOUT_NS gives up, records the fact in its header (asg_peer_status) and
// leaves the buffer untouched rather than hanging the GPU.
#include <cstdio>
#include <cstring>
#include <new>
#include "asg_common.cuh"
#include "asg_internal.h"

namespace asg {

constexpr int PEER_MAX = 8;
constexpr size_t PEER_HDR = 256;            // [0] epoch counter, [8] status, [64 + (parity*8 + rank)*8] arrival flags
constexpr unsigned long long ASG_PEER_TIMEOUT_NS = 4000000000ull;

struct PeerPtrs { char *base[PEER_MAX]; };

__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <typename T>
__global__ void __launch_bounds__(1024) peer_allreduce_kernel(PeerPtrs pp, int rank, int world, size_t slot_bytes, T *buf, int count) {
    __shared__ unsigned long long s_epoch;
    __shared__ int s_ok;
    char *me = pp.base[rank];
    const int tid = threadIdx.x;
    if (tid == 0) {
        unsigned long long *ctr = reinterpret_cast<unsigned long long *>(me);
        s_epoch = *ctr + 1;
        *ctr = s_epoch;
        s_ok = 1;
    }
    __syncthreads();
    const unsigned long long epoch = s_epoch;
    const int par = (int)(epoch & 1ull);
    // ---- push: my vector into slot [par][rank] of every rank (mine included)
    const size_t my_slot = PEER_HDR + (size_t)(par * world + rank) * slot_bytes;
    for (int i = tid; i < count; i += blockDim.x) {
        const T v = buf[i];
        for (int p = 0; p < world; ++p) reinterpret_cast<T *>(pp.base[p] + my_slot)[i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (tid < world) st_release_sys_u64(reinterpret_cast<unsigned long long *>(pp.base[tid] + 64 + (par * PEER_MAX + rank) * 8), epoch);
    // ---- wait: every rank's epoch has arrived in MY flags (local polling)
    if (tid < world) {
        const unsigned long long *f = reinterpret_cast<const unsigned long long *>(me + 64 + (par * PEER_MAX + tid) * 8);
        const unsigned long long t0 = globaltimer_ns();
        while (ld_acquire_sys_u64(f) < epoch) {
            if (globaltimer_ns() - t0 > ASG_PEER_TIMEOUT_NS) { s_ok = 0; reinterpret_cast<unsigned long long *>(me)[1] = epoch; break; }
        }
    }
    __syncthreads();
    if (!s_ok) return;
    // ---- sum in rank order (bit-identical on every rank); the slots were written by other GPUs: read them past L1
    const char *slots = me + PEER_HDR + (size_t)par * world * slot_bytes;
    for (int i = tid; i < count; i += blockDim.x) {
        T s = T(0);
        for (int r = 0; r < world; ++r) s += __ldcg(reinterpret_cast<const T *>(slots + (size_t)r * slot_bytes) + i);
        buf[i] = s;
    }
}

}  // namespace asg

using namespace asg;

struct asg_peer {
    int rank, world, device;
    size_t max_bytes, slot_bytes, total_bytes;
    char *local;
    PeerPtrs ptrs;
    bool opened[PEER_MAX];
    bool connected;
};

extern "C" {

int asg_peer_create(int rank, int world, size_t max_bytes, asg_peer_t **out, void *handle) {
    if (!out || !handle) return fail(ASG_E_NULL, "asg_peer_create: out and handle are required");
    if (world < 1 || world > PEER_MAX || rank < 0 || rank >= world)
        return fail(ASG_E_SHAPE, "asg_peer_create: 1 <= world <= %d and 0 <= rank < world (got rank %d, world %d)", PEER_MAX, rank, world);
    if (max_bytes == 0) return fail(ASG_E_SHAPE, "asg_peer_create: max_bytes must be positive");
    asg_peer *p = new (std::nothrow) asg_peer();
    if (!p) return fail(ASG_E_UNSUPPORTED, "asg_peer_create: out of host memory");
    p->rank = rank; p->world = world; p->max_bytes = max_bytes;
    p->slot_bytes = (max_bytes + 127) / 128 * 128;
    p->total_bytes = PEER_HDR + 2 * (size_t)world * p->slot_bytes;
    p->connected = false;
    memset(p->opened, 0, sizeof(p->opened));
    memset(&p->ptrs, 0, sizeof(p->ptrs));
    cudaGetDevice(&p->device);
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&p->local), p->total_bytes);     // cudaMalloc: IPC needs a real allocation
    if (e == cudaSuccess) e = cudaMemset(p->local, 0, p->total_bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p->local);
    if (e != cudaSuccess) {
        if (p->local) cudaFree(p->local);
        delete p;
        return cuda_fail((int)e, "asg_peer_create");
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == ASG_PEER_HANDLE_BYTES, "handle size");
    memcpy(handle, &h, sizeof(h));
    p->ptrs.base[rank] = p->local;
    if (world == 1) p->connected = true;
    *out = p;
    return ASG_OK;
}

int asg_peer_connect(asg_peer_t *p, const void *handles) {
    if (!p || !handles) return fail(ASG_E_NULL, "asg_peer_connect: peer and handles are required");
    if (p->connected) return ASG_OK;
    for (int r = 0; r < p->world; ++r) {
        if (r == p->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(handles) + (size_t)r * ASG_PEER_HANDLE_BYTES, sizeof(h));
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(ASG_E_UNSUPPORTED, "asg_peer_connect: cannot open rank %d's exchange buffer from rank %d (%s): no peer access?", r, p->rank,
                        cudaGetErrorString(e));
        }
        p->ptrs.base[r] = static_cast<char *>(ptr);
        p->opened[r] = true;
    }
    p->connected = true;
    return ASG_OK;
}

int asg_peer_allreduce(asg_peer_t *p, int dtype, void *buf, int64_t count, asg_stream_t stream) {
    if (!p || !buf) return fail(ASG_E_NULL, "asg_peer_allreduce: peer and buf are required");
    if (!p->connected) return fail(ASG_E_UNSUPPORTED, "asg_peer_allreduce: asg_peer_connect has not succeeded on this rank");
    if (dtype != ASG_F32 && dtype != ASG_F64) return fail(ASG_E_DTYPE, "dtype must be ASG_F32 or ASG_F64, got %d", dtype);
    const size_t eb = dtype == ASG_F32 ? 4 : 8;
    if (count <= 0 || (size_t)count * eb > p->max_bytes)
        return fail(ASG_E_SHAPE, "asg_peer_allreduce: %lld elements do not fit the exchange slots (%zu bytes)", (long long)count, p->max_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    {
        LaunchScope ls(K_PEER_ALLREDUCE, st);
        if (dtype == ASG_F32)
            peer_allreduce_kernel<float><<<1, 1024, 0, st>>>(p->ptrs, p->rank, p->world, p->slot_bytes, static_cast<float *>(buf), (int)count);
        else
            peer_allreduce_kernel<double><<<1, 1024, 0, st>>>(p->ptrs, p->rank, p->world, p->slot_bytes, static_cast<double *>(buf), (int)count);
    }
    return cuda_fail((int)cudaGetLastError(), "asg_peer_allreduce");
}

int asg_peer_status(asg_peer_t *p, uint64_t *epoch, uint64_t *timed_out_epoch) {
    if (!p) return fail(ASG_E_NULL, "asg_peer_status: peer is required");
    unsigned long long h[2] = {0, 0};
    cudaError_t e = cudaMemcpy(h, p->local, sizeof(h), cudaMemcpyDeviceToHost);      // synchronises with the device
    if (e != cudaSuccess) return cuda_fail((int)e, "asg_peer_status");
    if (epoch) *epoch = h[0];
    if (timed_out_epoch) *timed_out_epoch = h[1];
    return ASG_OK;
}

int asg_peer_destroy(asg_peer_t *p) {
    if (!p) return ASG_OK;
    for (int r = 0; r < p->world; ++r)
        if (p->opened[r]) cudaIpcCloseMemHandle(p->ptrs.base[r]);
    if (p->local) cudaFree(p->local);
    delete p;
    return ASG_OK;
}

}  // extern "C"
