// Developer probe: event-timed duration of (nearly) empty kernels vs block size / dynamic shared memory / setmaxnreg.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/launch_probe tools/launch_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 1) k_empty(int *out) {
    extern __shared__ char sm[];
    if (threadIdx.x == 0 && out == (int *)1) out[0] = sm[0];
}
__global__ void __launch_bounds__(512, 1) k_regs(int *out) {
    extern __shared__ char sm[];
    if (threadIdx.x >= 384) { asm volatile("setmaxnreg.dec.sync.aligned.u32 32;\n"); return; }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 160;\n");
    if (threadIdx.x == 0 && out == (int *)1) out[0] = sm[0];
}
__global__ void k_small(int *out) { if (out == (int *)1) out[0] = 1; }
template <typename F> float timeit(F f, int n = 200) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float tot = 0;
    for (int i = 0; i < n; ++i) {
        k_small<<<148, 128>>>(nullptr);        // something with a different configuration in between
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (i >= 10) tot += ms;
    }
    return 1e3f * tot / (n - 10);
}
int main() {
    cudaFuncSetAttribute(k_empty, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(k_regs, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    printf("small 148x128 0 B      : %.2f us\n", timeit([] { k_small<<<148, 128>>>(nullptr); }));
    for (int smem : {0, 48 * 1024, 100 * 1024, 160 * 1024, 227 * 1024}) {
        printf("empty 148x512 %6d B : %.2f us\n", smem, timeit([=] { k_empty<<<148, 512, smem>>>(nullptr); }));
        printf("empty 148x288 %6d B : %.2f us\n", smem, timeit([=] { k_empty<<<148, 288, smem>>>(nullptr); }));
    }
    printf("setmaxnreg 148x512 227K: %.2f us\n", timeit([] { k_regs<<<148, 512, 227 * 1024>>>(nullptr); }));
    return 0;
}
