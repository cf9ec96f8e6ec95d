// Throughput of the legacy warp-level MMA path on sm_100a (mma.sync.m16n8k8 tf32), to size a tensor-core filter
// inside a SIMT search kernel.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/mma_probe.cu -o tools/mma_probe.bin
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int CHAINS>
__global__ void probe(float *out, int iters) {
    float d[CHAINS][4];
    unsigned a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i);
    for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.5f + threadIdx.x * 1e-3f + i);
    for (int c = 0; c < CHAINS; ++c)
        for (int i = 0; i < 4; ++i) d[c][i] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) mma_tf32(d[c], a, b);
    }
    float s = 0.f;
    for (int c = 0; c < CHAINS; ++c)
        for (int i = 0; i < 4; ++i) s += d[c][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
void run(int warps_per_sm, int sms, float clock_ghz) {
    const int iters = 20000;
    float *out;
    cudaMalloc(&out, size_t(sms) * warps_per_sm * 32 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<CHAINS><<<sms, warps_per_sm * 32>>>(out, 100);
    cudaEventRecord(e0);
    probe<CHAINS><<<sms, warps_per_sm * 32>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double mmas_per_sm = double(iters) * CHAINS * warps_per_sm;
    const double clk = ms * 1e-3 * clock_ghz * 1e9;
    printf("chains=%d warps/SM=%2d: %.3f ms  %.3f mma.m16n8k8.tf32 / clk / SM  (%.0f MAC/clk/SM)\n", CHAINS, warps_per_sm, ms,
           mmas_per_sm / clk, mmas_per_sm / clk * 1024);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const float ghz = khz * 1e-6f;
    printf("%s, %d SMs, %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
    for (int w : {4, 8, 16, 32}) run<1>(w, p.multiProcessorCount, ghz);
    for (int w : {4, 8, 16}) run<4>(w, p.multiProcessorCount, ghz);
    for (int w : {4, 8}) run<8>(w, p.multiProcessorCount, ghz);
    return 0;
}
