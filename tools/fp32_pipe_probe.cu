// fp32_pipe_probe.cu -- measures the B200 FP32 pipe under the instruction mixes the NN kernel can use.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp32_pipe_probe tools/fp32_pipe_probe.cu
// Prints lane-operations per SM per clock (128 = nominal peak) for each mix.
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
template <int MODE>
__global__ void __launch_bounds__(256) probe(float2 *out, int iters, float seed) {
    float2 a[CHAINS], b[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) {
        a[k] = make_float2(seed + k + threadIdx.x, seed - k);
        b[k] = make_float2(seed * 0.5f + k, seed + 2 * k);
    }
    const float2 m = make_float2(0.999f, 1.001f), c = make_float2(1e-3f, -1e-3f);
    const float qs = seed * 0.25f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) {
            if (MODE == 0) {  // scalar FFMA x2 (3 distinct regs)
                a[k].x = fmaf(a[k].x, m.x, c.x);
                a[k].y = fmaf(a[k].y, m.y, c.y);
            } else if (MODE == 1) {  // FFMA2 a*m+c (3 distinct pairs)
                a[k] = __ffma2_rn(a[k], m, c);
            } else if (MODE == 2) {  // FFMA2 d*d+s
                a[k] = __ffma2_rn(b[k], b[k], a[k]);
            } else if (MODE == 3) {  // FMUL2 d*d
                a[k] = __fmul2_rn(a[k], a[k]);
            } else if (MODE == 4) {  // FADD2 scalar-broadcast + pair
                a[k] = __fadd2_rn(make_float2(qs, qs), make_float2(-a[k].x, -a[k].y));
            } else if (MODE == 5) {  // the NN mix per pair-of-candidates: 3 FADD2, 1 FMUL2, 2 FFMA2
                float2 dx = __fadd2_rn(make_float2(qs, qs), make_float2(-a[k].x, -a[k].y));
                float2 dy = __fadd2_rn(make_float2(seed, seed), make_float2(-b[k].x, -b[k].y));
                float2 dz = __fadd2_rn(make_float2(m.x, m.x), make_float2(-a[k].y, -b[k].x));
                float2 s = __fmul2_rn(dx, dx);
                s = __ffma2_rn(dy, dy, s);
                s = __ffma2_rn(dz, dz, s);
                a[k] = s;
                b[k].x += 1.0f;  // keep b changing (1 extra scalar op, counted below)
            } else if (MODE == 6) {  // same mix, scalar instructions
                float dx0 = qs - a[k].x, dx1 = qs - a[k].y, dy0 = seed - b[k].x, dy1 = seed - b[k].y;
                float dz0 = m.x - a[k].y, dz1 = m.x - b[k].x;
                float s0 = dx0 * dx0, s1 = dx1 * dx1;
                s0 = fmaf(dy0, dy0, s0); s1 = fmaf(dy1, dy1, s1);
                s0 = fmaf(dz0, dz0, s0); s1 = fmaf(dz1, dz1, s1);
                a[k] = make_float2(s0, s1);
                b[k].x += 1.0f;
            } else if (MODE == 7) {  // half packed / half scalar NN mix
                if (k & 1) {
                    float2 dx = __fadd2_rn(make_float2(qs, qs), make_float2(-a[k].x, -a[k].y));
                    float2 dy = __fadd2_rn(make_float2(seed, seed), make_float2(-b[k].x, -b[k].y));
                    float2 dz = __fadd2_rn(make_float2(m.x, m.x), make_float2(-a[k].y, -b[k].x));
                    float2 s = __fmul2_rn(dx, dx);
                    s = __ffma2_rn(dy, dy, s);
                    s = __ffma2_rn(dz, dz, s);
                    a[k] = s;
                } else {
                    float dx0 = qs - a[k].x, dx1 = qs - a[k].y, dy0 = seed - b[k].x, dy1 = seed - b[k].y;
                    float dz0 = m.x - a[k].y, dz1 = m.x - b[k].x;
                    float s0 = dx0 * dx0, s1 = dx1 * dx1;
                    s0 = fmaf(dy0, dy0, s0); s1 = fmaf(dy1, dy1, s1);
                    s0 = fmaf(dz0, dz0, s0); s1 = fmaf(dz1, dz1, s1);
                    a[k] = make_float2(s0, s1);
                }
                b[k].x += 1.0f;
            }
        }
    }
    float2 s = a[0];
#pragma unroll
    for (int k = 1; k < CHAINS; ++k) { s.x += a[k].x + b[k].x; s.y += a[k].y + b[k].y; }
    if (s.x == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, double lane_ops_per_chain_iter, int sms, double clock_hz, float2 *out, int bps) {
    const int blocks = sms * bps, iters = 8000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        probe<MODE><<<blocks, 256>>>(out, iters, 1.0f + rep);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double rate = double(blocks) * 256.0 * iters * CHAINS * lane_ops_per_chain_iter / (ms * 1e-3);
        if (rep && rate > best) best = rate;
    }
    printf("%-44s bps=%d  %8.2f T lane-op/s   %6.1f lane-ops/SM/clk @%.0f MHz\n", name, bps, best / 1e12,
           best / sms / clock_hz, clock_hz / 1e6);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double hz = khz * 1e3;
    float2 *out; cudaMalloc(&out, sizeof(float2) * sms * 8 * 256);
    for (int bps : {2, 4, 8}) {
        run<0>("scalar FFMA a*m+c", 2, sms, hz, out, bps);
        run<1>("FFMA2 a*m+c (3 pairs)", 2, sms, hz, out, bps);
        run<2>("FFMA2 d*d+s (2 pairs)", 2, sms, hz, out, bps);
        run<3>("FMUL2 d*d", 2, sms, hz, out, bps);
        run<4>("FADD2 scalar-bcast - pair", 2, sms, hz, out, bps);
        run<5>("NN mix packed (3 FADD2,1 FMUL2,2 FFMA2)+1", 13, sms, hz, out, bps);
        run<6>("NN mix scalar (12 ops)+1", 13, sms, hz, out, bps);
        run<7>("NN mix half packed / half scalar +1", 13, sms, hz, out, bps);
    }
    return 0;
}
