// alu_min_probe.cu -- throughput of FMNMX / FMNMX3 / FADD.SAT on B200 (lane-ops per SM per clock).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) probe(float *out, int iters, float seed) {
    float a[8], b[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] = seed + k + threadIdx.x; b[k] = seed * 0.5f + k; }
    const float c = seed * 3.0f, d = seed * 0.25f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0) a[k] = fminf(a[k], b[k] + 0.f), b[k] = fminf(b[k], c);                 // 2 x FMNMX
            else if (MODE == 1) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a[k]), "f"(b[k]), "f"(c)); a[k] = r;
                                  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(b[k]), "f"(a[k]), "f"(d)); b[k] = r; }  // 2 x FMNMX3
            else if (MODE == 2) { float r; asm("add.sat.f32 %0, %1, %2;" : "=f"(r) : "f"(c), "f"(a[k])); a[k] = r + b[k];
                                  asm("add.sat.f32 %0, %1, %2;" : "=f"(r) : "f"(d), "f"(b[k])); b[k] = r + a[k]; }      // 2 x (FADD.SAT + FADD)
            else if (MODE == 3) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a[k]), "f"(b[k]), "f"(c)); a[k] = r;   // FMNMX3 + FADD.SAT + FADD
                                  asm("add.sat.f32 %0, %1, %2;" : "=f"(r) : "f"(d), "f"(b[k])); b[k] = r + a[k]; }
        }
    }
    float s = 0; for (int k = 0; k < 8; ++k) s += a[k] + b[k];
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, double ops, int sms, double hz, float *out) {
    const int blocks = sms * 8, iters = 8000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); probe<MODE><<<blocks, 256>>>(out, iters, 1.0f + rep); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double rate = double(blocks) * 256.0 * iters * 8 * ops / (ms * 1e-3);
        if (rep && rate > best) best = rate;
    }
    printf("%-40s %7.1f lane-ops/SM/clk\n", name, best / sms / hz);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *out; cudaMalloc(&out, 4 * 256 * p.multiProcessorCount * 8);
    run<0>("FMNMX (2-input min)", 2, p.multiProcessorCount, khz * 1e3, out);
    run<1>("FMNMX3 (3-input min)", 2, p.multiProcessorCount, khz * 1e3, out);
    run<2>("FADD.SAT + FADD", 4, p.multiProcessorCount, khz * 1e3, out);
    run<3>("FMNMX3 + FADD.SAT + FADD mixed", 3, p.multiProcessorCount, khz * 1e3, out);
    return 0;
}
