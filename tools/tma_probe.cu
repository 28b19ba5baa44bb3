// tma_probe.cu -- developer micro-benchmark (B200): HBM read throughput of 1-D bulk (TMA) copies issued by one thread
// per CTA, as a function of copy size and copies in flight, next to a plain LDG.128 streaming read.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tools/tma_probe.cu && ./tma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity), "r"(0x989680u) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk(void *dst, const void *src, unsigned bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

// each CTA streams its contiguous slice of `src`; `depth` copies of `csize` bytes in flight; `split` bulk copies per chunk
__global__ void tma_stream(const char *src, size_t bytes_per_cta, int csize, int depth, int split, float *sink) {
    extern __shared__ __align__(128) char smem[];
    __shared__ uint64_t bar[32];
    if (threadIdx.x == 0) { for (int i = 0; i < depth; ++i) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const char *base = src + (size_t)blockIdx.x * bytes_per_cta;
    const int nchunks = (int)(bytes_per_cta / csize);
    float acc = 0.f;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const int sub = csize / split;
        for (int i = 0; i < nchunks + depth; ++i) {
            if (i >= depth) {               // consume chunk i-depth
                const int s = (i - depth) % depth;
                if (lane == 0) mbar_wait(&bar[s], ((i - depth) / depth) & 1);
                __syncwarp();
                acc += *(float *)(smem + (size_t)s * csize + lane * 4);
                __syncwarp();
            }
            if (i < nchunks) {
                const int s = i % depth;
                if (lane == 0) mbar_expect(&bar[s], (unsigned)csize);
                __syncwarp();
                if (lane < split) bulk(smem + (size_t)s * csize + (size_t)lane * sub, base + (size_t)i * csize + (size_t)lane * sub, (unsigned)sub, &bar[s]);
            }
        }
    }
    if (acc == 123.456f) sink[0] = acc;
}

__global__ void ldg_stream(const float4 *src, size_t n4_per_cta, float *sink) {
    const float4 *base = src + (size_t)blockIdx.x * n4_per_cta;
    float acc = 0.f;
    for (size_t i = threadIdx.x; i + 3 * blockDim.x < n4_per_cta; i += 4 * blockDim.x) {
        float4 a = __ldcs(base + i), b = __ldcs(base + i + blockDim.x), c = __ldcs(base + i + 2 * blockDim.x), d = __ldcs(base + i + 3 * blockDim.x);
        acc += a.x + b.y + c.z + d.w;
    }
    if (acc == 123.456f) sink[0] = acc;
}

// the gradient kernel's protocol: one producer warp + NW consumer warps, full/empty mbarrier pairs, chunks of `csize`
// bytes advancing by `adv` bytes (overlapping halo), consumers touch one word and release
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__global__ void ring_stream(const char *src, size_t bytes_per_cta, int csize, int adv, int depth, int nw, float *sink) {
    extern __shared__ __align__(128) char smem[];
    __shared__ uint64_t full[16], empty[16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int i = 0; i < depth; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], nw); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const char *base = src + (size_t)blockIdx.x * bytes_per_cta;
    const int nchunks = (int)((bytes_per_cta - csize) / adv);
    const size_t sstride = ((size_t)csize + 127) / 128 * 128;
    float acc = 0.f;
    if (warp == nw) {
        if (lane == 0)
            for (int i = 0; i < nchunks; ++i) {
                const int s = i % depth, use = i / depth;
                mbar_wait(&empty[s], (use & 1) ^ 1);
                mbar_expect(&full[s], (unsigned)csize);
                bulk(smem + s * sstride, base + (size_t)i * adv, (unsigned)csize, &full[s]);
            }
    } else {
        for (int i = 0; i < nchunks; ++i) {
            const int s = i % depth, use = i / depth;
            mbar_wait(&full[s], use & 1);
            acc += *(float *)(smem + s * sstride + (warp * 32 + lane) * 4);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
    }
    if (acc == 123.456f) sink[0] = acc;
}

int main() {
    const size_t total = (size_t)592 << 20;   // 592 MiB > L2
    char *src; float *sink;
    cudaMalloc(&src, total); cudaMalloc(&sink, 4);
    cudaMemset(src, 1, total);
    int sms = 148;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t per = total / sms / 65536 * 65536;
    for (int rep = 0; rep < 2; ++rep) {
        ldg_stream<<<sms, 1024>>>((const float4 *)src, per / 16, sink);
    }
    cudaEventRecord(e0); ldg_stream<<<sms, 1024>>>((const float4 *)src, per / 16, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("ldg.128 x4 unroll, 148x1024 thr: %.0f GB/s\n", per * sms / ms / 1e6);
    cudaEventRecord(e0); ldg_stream<<<sms * 2, 512>>>((const float4 *)src, per / 32, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("ldg.128 x4 unroll, 296x512 thr:  %.0f GB/s\n", per * sms / ms / 1e6);
    cudaFuncSetAttribute(tma_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int sizes[] = {32768};
    const int depths[] = {1, 4};
    const int splits[] = {1};
    for (int split : splits)
        for (int cs : sizes)
            for (int d : depths) {
                if ((size_t)cs * d > 192 * 1024) continue;
                tma_stream<<<sms, 64, (size_t)cs * d>>>(src, per, cs, d, split, sink);
                cudaEventRecord(e0);
                tma_stream<<<sms, 64, (size_t)cs * d>>>(src, per, cs, d, split, sink);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1);
                printf("bulk copy %6d B x depth %2d x split %d (in flight %6d B/SM): %6.0f GB/s  %s\n", cs, d, split, cs * d, per * sms / ms / 1e6,
                       cudaGetLastError() == cudaSuccess ? "" : "ERR");
            }
    cudaFuncSetAttribute(ring_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int rec = 2368;
    const int cfgs[][4] = {{13 * rec, 12 * rec, 5, 17}, {13 * rec, 12 * rec, 5, 1}, {13 * rec, 13 * rec, 5, 17}, {32768, 32768, 5, 17}, {32768, 32768, 5, 1},
                           {17 * rec, 16 * rec, 4, 17}, {25 * rec, 24 * rec, 3, 17}};
    for (auto &c : cfgs) {
        const size_t smem = ((size_t)c[0] + 127) / 128 * 128 * c[2];
        ring_stream<<<sms, 32 * (c[3] + 1), smem>>>(src, per, c[0], c[1], c[2], c[3], sink);
        cudaEventRecord(e0);
        ring_stream<<<sms, 32 * (c[3] + 1), smem>>>(src, per, c[0], c[1], c[2], c[3], sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const size_t nch = (per - c[0]) / c[1];
        printf("ring: chunk %6d B advance %6d B depth %d consumers %2d warps: %6.0f GB/s moved (%6.0f GB/s useful)  %s\n", c[0], c[1], c[2], c[3],
               (double)nch * c[0] * sms / ms / 1e6, (double)nch * c[1] * sms / ms / 1e6, cudaGetLastError() == cudaSuccess ? "" : cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
