// adam.cu -- the optimiser update of the reference loop as capturable kernels (global_optimization.py:188, :592:
// torch.optim.Adam over [body_rotation_rec, scale, camera_ext, c_dct], lr 0.005, default betas / eps, no weight decay).
// The step counter lives on the device, so a captured fit step that ends with the update replays correctly: the body
// really moves from one replay to the next, as it does under the reference's optimiser.
#include "common.cuh"

namespace fpv {

__global__ void adam_tick_kernel(float *step) { *step += 1.0f; }

// torch.optim.Adam (amsgrad=False, maximize=False, weight_decay=0), single-tensor formulation:
//   m = b1 m + (1-b1) g ;  v = b2 v + (1-b2) g^2 ;  p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
__global__ void __launch_bounds__(256) adam_update_kernel(float *__restrict__ p, const float *__restrict__ g,
                                                          float *__restrict__ m, float *__restrict__ v, int64_t n,
                                                          float lr, float b1, float b2, float eps,
                                                          const float *__restrict__ step) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float t = *step;
    const float bc1 = 1.0f - powf(b1, t), bc2 = 1.0f - powf(b2, t);
    const float gi = g[i];
    const float mi = m[i] + (gi - m[i]) * (1.0f - b1);          // torch: exp_avg.lerp_(grad, 1 - beta1)
    const float vi = v[i] * b2 + gi * gi * (1.0f - b2);         // torch: exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    p[i] = p[i] - (lr / bc1) * (mi / denom);
}

}  // namespace fpv

using namespace fpv;

extern "C" {

/* step (device float, starts at 0) += 1: call once per optimiser step, before the fpv_adam_update calls of that step. */
int fpv_adam_tick(float *step, fpv_stream_t stream) {
    FPV_CHECK_ARG(step, "fpv_adam_tick: null pointer");
    adam_tick_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(step);
    FPV_LAUNCH_CHECK("adam_tick_kernel");
    return FPV_OK;
}

/* One Adam update of n parameters in place (torch.optim.Adam semantics, global_optimization.py:188, :592). */
int fpv_adam_update(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr,
                    float beta1, float beta2, float eps, const float *step, fpv_stream_t stream) {
    FPV_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && step && n > 0, "fpv_adam_update: bad arguments");
    adam_update_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step);
    FPV_LAUNCH_CHECK("adam_update_kernel");
    return FPV_OK;
}

}  // extern "C"
