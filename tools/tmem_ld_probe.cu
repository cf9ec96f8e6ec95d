// tmem_ld_probe.cu -- measures TMEM -> register bandwidth of tcgen05.ld on B200 (bytes / clock / SM).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_ld_probe.bin tools/tmem_ld_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD32(taddr, r)                                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                          \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                          \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"          \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),  \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),        \
                   "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),      \
                   "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),      \
                   "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                            \
                 : "r"(taddr)                                                                                       \
                 : "memory")

template <int DEPTH>
__global__ void probe(long long *cycles, uint32_t *sink, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t r[DEPTH][32];
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) LD32(base + (uint32_t)(((i * DEPTH + d) * 32) & 511 & ~31), r[d]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int d = 0; d < DEPTH; ++d)
#pragma unroll
            for (int k = 0; k < 32; k += 8) acc ^= r[d][k];
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
    }
}

template <int DEPTH>
void run(int warps, int sms) {
    long long *cyc; uint32_t *sink;
    cudaMalloc(&cyc, sizeof(long long) * sms); cudaMalloc(&sink, 4096);
    const int iters = 4000;
    probe<DEPTH><<<sms, warps * 32>>>(cyc, sink, iters);
    probe<DEPTH><<<sms, warps * 32>>>(cyc, sink, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < sms; ++i) c += h[i]; c /= sms;
    const double bytes = double(warps) * 32 * 32 * 4 * DEPTH * iters;
    printf("warps=%2d loads-in-flight=%d: %8.1f cycles/iter/warp-set, %7.1f B/clk/SM  (%s)\n", warps, DEPTH, c / iters, bytes / c,
           cudaGetErrorString(e));
    cudaFree(cyc); cudaFree(sink);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    for (int w : {4, 8, 16}) { run<1>(w, sms); run<2>(w, sms); run<4>(w, sms); }
    return 0;
}
