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
    const int sizes[] = {2048, 8192, 16384, 32768};
    const int depths[] = {1, 2, 4, 6, 12};
    const int splits[] = {1, 8};
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
    return 0;
}
