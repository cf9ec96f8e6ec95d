// common.cu -- error plumbing and device queries behind the C ABI.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace fpv {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }

struct ProfRec {
    char name[48];
    cudaEvent_t e0, e1;
    double bytes, items;
};
static bool g_prof = false;
static ProfRec g_recs[4096];
static int g_nrec = 0;
static bool g_open = false;

bool profile_on() { return g_prof; }
void profile_begin(const char *name, cudaStream_t st, double algo_bytes, double work_items) {
    if (!g_prof || g_nrec >= 4096) return;
    ProfRec &r = g_recs[g_nrec];
    snprintf(r.name, sizeof(r.name), "%s", name);
    r.bytes = algo_bytes;
    r.items = work_items;
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, st);
    g_open = true;
}
void profile_end(cudaStream_t st) {
    if (!g_open) return;
    cudaEventRecord(g_recs[g_nrec].e1, st);
    ++g_nrec;
    g_open = false;
}

// FP32 issue-rate probe: independent FFMA2 chains, no memory traffic.
__global__ void __launch_bounds__(256) ffma_probe_kernel(float2 *out, int iters, float seed) {
    float2 a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = make_float2(seed + k + threadIdx.x, seed - k);
    const float2 m = make_float2(0.999f, 1.001f), c = make_float2(1e-3f, -1e-3f);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = __ffma2_rn(a[k], m, c);
    }
    float2 s = a[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        s.x += a[k].x;
        s.y += a[k].y;
    }
    if (s.x == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace fpv

extern "C" {

unsigned long long fpv_launch_count(void) { return fpv::g_launches; }

int fpv_profile_enable(int on) {
    for (int i = 0; i < fpv::g_nrec; ++i) {
        cudaEventDestroy(fpv::g_recs[i].e0);
        cudaEventDestroy(fpv::g_recs[i].e1);
    }
    fpv::g_nrec = 0;
    fpv::g_open = false;
    fpv::g_prof = on != 0;
    return FPV_OK;
}

int fpv_profile_count(void) { return fpv::g_nrec; }

int fpv_profile_get(int i, char *name48, float *ms, double *algo_bytes, double *work_items) {
    FPV_CHECK_ARG(i >= 0 && i < fpv::g_nrec, "fpv_profile_get: record %d out of range", i);
    fpv::ProfRec &r = fpv::g_recs[i];
    FPV_CUDA(cudaEventSynchronize(r.e1));
    float t = 0.f;
    FPV_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
    if (name48) snprintf(name48, 48, "%s", r.name);
    if (ms) *ms = t;
    if (algo_bytes) *algo_bytes = r.bytes;
    if (work_items) *work_items = r.items;
    return FPV_OK;
}

/* Measures the FP32 pipe: returns fused multiply-add LANE operations per second (x2 = FLOP/s). */
int fpv_fp32_probe(double *lane_fma_per_s, fpv_stream_t stream) {
    FPV_CHECK_ARG(lane_fma_per_s, "fpv_fp32_probe: null output");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int blocks = fpv::sm_count() * 8, iters = 20000;
    float2 *out = nullptr;
    FPV_CUDA(cudaMalloc(&out, sizeof(float2) * size_t(blocks) * 256));
    cudaEvent_t e0, e1;
    FPV_CUDA(cudaEventCreate(&e0));
    FPV_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        FPV_CUDA(cudaEventRecord(e0, st));
        fpv::ffma_probe_kernel<<<blocks, 256, 0, st>>>(out, iters, 1.0f + rep);
        FPV_CUDA(cudaEventRecord(e1, st));
        FPV_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        FPV_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double rate = double(blocks) * 256.0 * iters * 8.0 * 2.0 / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *lane_fma_per_s = best;
    return FPV_OK;
}

const char *fpv_last_error(void) { return fpv::g_err; }

int fpv_abi_version(void) { return FPV_ABI_VERSION; }

int fpv_device_query(int device, int *sm_count, int *cc_major, int *cc_minor) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        fpv::set_error("no CUDA device visible (%s); this library has no CPU fallback",
                       e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return FPV_ERR_CUDA;
    }
    FPV_CHECK_ARG(device >= 0 && device < n, "device %d out of range [0,%d)", device, n);
    cudaDeviceProp p;
    FPV_CUDA(cudaGetDeviceProperties(&p, device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (p.major != 10) {
        fpv::set_error("device %d is sm_%d%d; kernels are built for sm_100a only", device, p.major, p.minor);
        return FPV_ERR_CUDA;
    }
    return FPV_OK;
}

}  // extern "C"
