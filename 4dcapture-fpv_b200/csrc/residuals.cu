// residuals.cu -- the memory-bound loss algebra around the chamfer term: robust contact mean,
// temporal finite-difference L1 residuals, and the rigid world transform.  sm_100a.
//
// Replaces (global_optimization.py): :295 robust contact loss; :266-267 / :381-382 parameter
// second difference; :304 world-joint first difference; :404-405 vertex second difference;
// :415-429 weighted leg velocity; :119-127 verts_transform.
// All scalar reductions are two-stage with a fixed block count and fixed summation order
// (double accumulators), so repeated runs are bitwise identical.
#include "common.cuh"

namespace fpv {

constexpr int RED_THREADS = 256;
constexpr int RED_MAX_BLOCKS = 1184;  // 8 x 148

static int red_blocks(int64_t n) {
    int64_t b = ceil_div(n, int64_t(RED_THREADS) * 8);
    return int(b < 1 ? 1 : (b > RED_MAX_BLOCKS ? RED_MAX_BLOCKS : b));
}

__global__ void finalize_mean_kernel(const double *__restrict__ partial, int nblk, double inv_count, float *out) {
    __shared__ double scratch[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < nblk; i += blockDim.x) s += partial[i];
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) out[0] = float(s * inv_count);
}

// ---- robust contact mean -------------------------------------------------------------------
__global__ void robust_partial_kernel(const float *__restrict__ d, int64_t n, float eps, double *partial) {
    __shared__ double scratch[32];
    double s = 0.0;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const float r = sqrtf(__fadd_rn(d[i], eps));
        s += double(__fdiv_rn(r, __fadd_rn(r, 1.0f)));
    }
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void robust_bwd_kernel(const float *__restrict__ d, int64_t n, float eps, const float *__restrict__ g_out,
                                  float *__restrict__ grad_d) {
    const float gs = g_out[0] / float(n);
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const float r = sqrtf(__fadd_rn(d[i], eps));
        const float r1 = r + 1.0f;
        // d/dd [ r/(r+1) ] = 1/(r+1)^2 * 1/(2r)
        grad_d[i] = gs / (2.0f * r * r1 * r1);
    }
}

// ---- temporal differences ------------------------------------------------------------------
__device__ __forceinline__ float tdiff_res(const float *__restrict__ x, int64_t F, int64_t t, int64_t f, int order,
                                           const float *__restrict__ w) {
    const float x0 = x[t * F + f], x1 = x[(t + 1) * F + f];
    if (order == 2) {
        const float x2 = x[(t + 2) * F + f];
        return __fsub_rn(__fsub_rn(x0, x1), __fsub_rn(x1, x2));
    }
    const float r = __fsub_rn(x0, x1);
    return w ? __fmul_rn(r, w[t + 1]) : r;
}

__global__ void tdiff_partial_kernel(const float *__restrict__ x, int64_t T, int64_t F, int order,
                                     const float *__restrict__ w, double *partial) {
    __shared__ double scratch[32];
    const int64_t n = (T - order) * F;
    double s = 0.0;
    for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += int64_t(gridDim.x) * blockDim.x) {
        const int64_t t = e / F, f = e - t * F;
        s += double(fabsf(tdiff_res(x, F, t, f, order, w)));
    }
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__device__ __forceinline__ float sgn(float r) { return float(r > 0.f) - float(r < 0.f); }

__global__ void tdiff_bwd_kernel(const float *__restrict__ x, int64_t T, int64_t F, int order,
                                 const float *__restrict__ w, const float *__restrict__ g_out,
                                 float *__restrict__ grad_x) {
    const int64_t n = T * F;
    const float gs = g_out[0] / float((T - order) * F);
    const int64_t R = T - order;  // residual rows
    for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += int64_t(gridDim.x) * blockDim.x) {
        const int64_t t = e / F, f = e - t * F;
        float acc = 0.f;
        if (order == 2) {
            // r_t = x_t - 2 x_{t+1} + x_{t+2}:  dL/dx_t = s_t - 2 s_{t-1} + s_{t-2}
            if (t < R) acc += sgn(tdiff_res(x, F, t, f, 2, nullptr));
            if (t >= 1 && t - 1 < R) acc -= 2.f * sgn(tdiff_res(x, F, t - 1, f, 2, nullptr));
            if (t >= 2 && t - 2 < R) acc += sgn(tdiff_res(x, F, t - 2, f, 2, nullptr));
        } else {
            // r_t = (x_t - x_{t+1}) w_{t+1}:  dL/dx_t = s_t w_{t+1} - s_{t-1} w_t
            if (t < R) acc += sgn(tdiff_res(x, F, t, f, 1, w)) * (w ? w[t + 1] : 1.f);
            if (t >= 1) acc -= sgn(tdiff_res(x, F, t - 1, f, 1, w)) * (w ? w[t] : 1.f);
        }
        grad_x[e] = gs * acc;
    }
}

// ---- rigid transform -----------------------------------------------------------------------
__global__ void transform_fwd_kernel(const float *__restrict__ v, const float *__restrict__ mats, int64_t P,
                                     float *__restrict__ out) {
    __shared__ float m[12];
    const int64_t t = blockIdx.y;
    if (threadIdx.x < 12) m[threadIdx.x] = mats[t * 16 + threadIdx.x];
    __syncthreads();
    const int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const float *s = v + (t * P + p) * 3;
    const float x = s[0], y = s[1], z = s[2];
    float *o = out + (t * P + p) * 3;
#pragma unroll
    for (int r = 0; r < 3; ++r)
        o[r] = __fmaf_rn(m[4 * r + 2], z, __fmaf_rn(m[4 * r + 1], y, __fmaf_rn(m[4 * r], x, m[4 * r + 3])));
}

constexpr int TB_THREADS = 256;
constexpr int TB_PER_THREAD = 4;  // points per thread in the backward reduction

// g_v = R^T g ; per-block partial sums of g (x) [v,1] -> partial[t][blk][12]
__global__ void transform_bwd_kernel(const float *__restrict__ v, const float *__restrict__ mats,
                                     const float *__restrict__ g, int64_t P, float *__restrict__ g_v,
                                     float *__restrict__ partial) {
    __shared__ float m[12];
    __shared__ float scratch[32];
    const int64_t t = blockIdx.y;
    if (threadIdx.x < 12) m[threadIdx.x] = mats[t * 16 + threadIdx.x];
    __syncthreads();
    float acc[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[k] = 0.f;
    const int64_t p0 = int64_t(blockIdx.x) * (TB_THREADS * TB_PER_THREAD);
#pragma unroll
    for (int u = 0; u < TB_PER_THREAD; ++u) {
        const int64_t p = p0 + int64_t(u) * TB_THREADS + threadIdx.x;
        if (p < P) {
            const float *gs = g + (t * P + p) * 3, *vs = v + (t * P + p) * 3;
            const float g0 = gs[0], g1 = gs[1], g2 = gs[2];
            const float x = vs[0], y = vs[1], z = vs[2];
            if (g_v) {
                float *o = g_v + (t * P + p) * 3;
                o[0] = m[0] * g0 + m[4] * g1 + m[8] * g2;
                o[1] = m[1] * g0 + m[5] * g1 + m[9] * g2;
                o[2] = m[2] * g0 + m[6] * g1 + m[10] * g2;
            }
            acc[0] += g0 * x; acc[1] += g0 * y; acc[2] += g0 * z; acc[3] += g0;
            acc[4] += g1 * x; acc[5] += g1 * y; acc[6] += g1 * z; acc[7] += g1;
            acc[8] += g2 * x; acc[9] += g2 * y; acc[10] += g2 * z; acc[11] += g2;
        }
    }
    if (partial) {
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const float s = block_sum(acc[k], scratch);
            if (threadIdx.x == 0) partial[(t * gridDim.x + blockIdx.x) * 12 + k] = s;
        }
    }
}

__global__ void transform_bwd_finalize_kernel(const float *__restrict__ partial, int nblk, float *__restrict__ g_mats) {
    const int64_t t = blockIdx.x;
    const int k = threadIdx.x;
    if (k < 12) {
        float s = 0.f;
        for (int b = 0; b < nblk; ++b) s += partial[(t * nblk + b) * 12 + k];
        g_mats[t * 16 + k] = s;
    } else if (k < 16) {
        g_mats[t * 16 + k] = 0.f;
    }
}

}  // namespace fpv

using namespace fpv;

extern "C" {

size_t fpv_reduce_workspace_bytes(int64_t n) {
    (void)n;
    return align_up(size_t(RED_MAX_BLOCKS) * sizeof(double), 256);
}

int fpv_robust_mean_fwd(const float *d, int64_t n, float eps, float *out, void *workspace, size_t workspace_bytes,
                        fpv_stream_t stream) {
    FPV_CHECK_ARG(d && out && n > 0, "fpv_robust_mean_fwd: empty input");
    FPV_CHECK_ARG(workspace && workspace_bytes >= fpv_reduce_workspace_bytes(n), "fpv_robust_mean_fwd: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double *partial = static_cast<double *>(workspace);
    const int nb = red_blocks(n);
    robust_partial_kernel<<<nb, RED_THREADS, 0, st>>>(d, n, eps, partial);
    FPV_LAUNCH_CHECK("robust_partial_kernel");
    finalize_mean_kernel<<<1, 256, 0, st>>>(partial, nb, 1.0 / double(n), out);
    FPV_LAUNCH_CHECK("finalize_mean_kernel");
    return FPV_OK;
}

int fpv_robust_mean_bwd(const float *d, int64_t n, float eps, const float *g_out, float *grad_d, fpv_stream_t stream) {
    FPV_CHECK_ARG(d && g_out && grad_d && n > 0, "fpv_robust_mean_bwd: empty input");
    robust_bwd_kernel<<<red_blocks(n), RED_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(d, n, eps, g_out, grad_d);
    FPV_LAUNCH_CHECK("robust_bwd_kernel");
    return FPV_OK;
}

int fpv_tdiff_l1_fwd(const float *x, int64_t T, int64_t F, int order, const float *frame_w, float *out,
                     void *workspace, size_t workspace_bytes, fpv_stream_t stream) {
    FPV_CHECK_ARG(x && out, "fpv_tdiff_l1_fwd: null pointer");
    FPV_CHECK_ARG(order == 1 || order == 2, "fpv_tdiff_l1_fwd: order must be 1 or 2");
    FPV_CHECK_ARG(T > order && F > 0, "fpv_tdiff_l1_fwd: need T > order and F > 0 (T=%lld F=%lld)", (long long)T, (long long)F);
    FPV_CHECK_ARG(!(frame_w && order != 1), "fpv_tdiff_l1_fwd: frame weights only apply to order 1");
    FPV_CHECK_ARG(workspace && workspace_bytes >= fpv_reduce_workspace_bytes(T * F), "fpv_tdiff_l1_fwd: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double *partial = static_cast<double *>(workspace);
    const int64_t n = (T - order) * F;
    const int nb = red_blocks(n);
    tdiff_partial_kernel<<<nb, RED_THREADS, 0, st>>>(x, T, F, order, frame_w, partial);
    FPV_LAUNCH_CHECK("tdiff_partial_kernel");
    finalize_mean_kernel<<<1, 256, 0, st>>>(partial, nb, 1.0 / double(n), out);
    FPV_LAUNCH_CHECK("finalize_mean_kernel");
    return FPV_OK;
}

int fpv_tdiff_l1_bwd(const float *x, int64_t T, int64_t F, int order, const float *frame_w, const float *g_out,
                     float *grad_x, fpv_stream_t stream) {
    FPV_CHECK_ARG(x && g_out && grad_x, "fpv_tdiff_l1_bwd: null pointer");
    FPV_CHECK_ARG(order == 1 || order == 2, "fpv_tdiff_l1_bwd: order must be 1 or 2");
    FPV_CHECK_ARG(T > order && F > 0, "fpv_tdiff_l1_bwd: need T > order and F > 0");
    FPV_CHECK_ARG(!(frame_w && order != 1), "fpv_tdiff_l1_bwd: frame weights only apply to order 1");
    const int64_t n = T * F;
    const int nb = int(ceil_div(n, RED_THREADS) < 148 * 16 ? ceil_div(n, RED_THREADS) : 148 * 16);
    tdiff_bwd_kernel<<<nb, RED_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x, T, F, order, frame_w, g_out, grad_x);
    FPV_LAUNCH_CHECK("tdiff_bwd_kernel");
    return FPV_OK;
}

int fpv_transform_fwd(const float *verts, const float *mats, int64_t T, int64_t P, float *out, fpv_stream_t stream) {
    FPV_CHECK_ARG(verts && mats && out, "fpv_transform_fwd: null pointer");
    FPV_CHECK_ARG(T > 0 && P > 0 && T <= 65535, "fpv_transform_fwd: bad shape T=%lld P=%lld", (long long)T, (long long)P);
    dim3 grid((unsigned)ceil_div(P, 256), (unsigned)T);
    transform_fwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(verts, mats, P, out);
    FPV_LAUNCH_CHECK("transform_fwd_kernel");
    return FPV_OK;
}

size_t fpv_transform_bwd_workspace_bytes(int64_t T, int64_t P) {
    if (T <= 0 || P <= 0) return 0;
    return align_up(size_t(T) * size_t(ceil_div(P, TB_THREADS * TB_PER_THREAD)) * 12 * sizeof(float), 256);
}

int fpv_transform_bwd(const float *verts, const float *mats, const float *g_out, int64_t T, int64_t P, float *g_verts,
                      float *g_mats, void *workspace, size_t workspace_bytes, fpv_stream_t stream) {
    FPV_CHECK_ARG(verts && mats && g_out, "fpv_transform_bwd: null pointer");
    FPV_CHECK_ARG(g_verts || g_mats, "fpv_transform_bwd: no gradient requested");
    FPV_CHECK_ARG(T > 0 && P > 0 && T <= 65535, "fpv_transform_bwd: bad shape");
    FPV_CHECK_ARG(!g_mats || (workspace && workspace_bytes >= fpv_transform_bwd_workspace_bytes(T, P)),
                  "fpv_transform_bwd: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nblk = int(ceil_div(P, TB_THREADS * TB_PER_THREAD));
    dim3 grid((unsigned)nblk, (unsigned)T);
    float *partial = g_mats ? static_cast<float *>(workspace) : nullptr;
    transform_bwd_kernel<<<grid, TB_THREADS, 0, st>>>(verts, mats, g_out, P, g_verts, partial);
    FPV_LAUNCH_CHECK("transform_bwd_kernel");
    if (g_mats) {
        transform_bwd_finalize_kernel<<<(unsigned)T, 32, 0, st>>>(partial, nblk, g_mats);
        FPV_LAUNCH_CHECK("transform_bwd_finalize_kernel");
    }
    return FPV_OK;
}

}  // extern "C"
