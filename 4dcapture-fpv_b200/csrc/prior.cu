// prior.cu -- the parameter front-end and the trajectory prior either side of the SMPL-X forward (SURVEY.md
// section 8f rows f2, f3).  sm_100a.
//
//   rot6d -> axis-angle   convert_to_3D_rot (global_optimization.py:107-115): ContinousRotReprDecoder.decode
//                         (cvae.py:71-81, Gram-Schmidt on the two columns of view(-1,3,2)) followed by
//                         matrot2aa (cvae.py:84-93) = [3P] torchgeometry rotation_matrix_to_angle_axis
//                         (rotation matrix -> quaternion, four-branch; quaternion -> angle axis).
//   axis-angle -> rot6d   convert_to_6D_rot (:96-104): [3P] torchgeometry angle_axis_to_rotation_matrix, first
//                         two columns.
//   VPoser decode         [3P] human_body_prior v1 VPoser.decode(z, output_type='aa') (call site :270-271):
//                         Linear(32,512) lrelu(0.2) Linear(512,512) lrelu(0.2) Linear(512,21*6) -> the codec above.
//   DCT prior             FittingOP.cal_dctloss (:232-246): Geman-McClure residual of every joint trajectory
//                         against its low-order DCT reconstruction.
//
// The backward of the codec is forward-mode: the scalar routine is a template, instantiated once on float and
// once on a dual number carrying the six input partials, so the Jacobian follows exactly the branch the forward
// took (what autograd does through the reference's mask products).
#include <math_constants.h>

#include "common.cuh"

namespace fpv {

// ---------------------------------------------------------------------------------------------
// dual numbers (6 partials)
// ---------------------------------------------------------------------------------------------
struct Dual6 {
    float v;
    float d[6];
};
__device__ __forceinline__ Dual6 mk(float v) {
    Dual6 r;
    r.v = v;
#pragma unroll
    for (int i = 0; i < 6; ++i) r.d[i] = 0.f;
    return r;
}
__device__ __forceinline__ float val(float a) { return a; }
__device__ __forceinline__ float val(const Dual6 &a) { return a.v; }
__device__ __forceinline__ Dual6 operator+(const Dual6 &a, const Dual6 &b) {
    Dual6 r;
    r.v = a.v + b.v;
#pragma unroll
    for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
__device__ __forceinline__ Dual6 operator-(const Dual6 &a, const Dual6 &b) {
    Dual6 r;
    r.v = a.v - b.v;
#pragma unroll
    for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
__device__ __forceinline__ Dual6 operator-(const Dual6 &a) {
    Dual6 r;
    r.v = -a.v;
#pragma unroll
    for (int i = 0; i < 6; ++i) r.d[i] = -a.d[i];
    return r;
}
__device__ __forceinline__ Dual6 operator*(const Dual6 &a, const Dual6 &b) {
    Dual6 r;
    r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
}
__device__ __forceinline__ Dual6 operator/(const Dual6 &a, const Dual6 &b) {
    Dual6 r;
    const float inv = 1.0f / b.v;
    r.v = a.v * inv;
#pragma unroll
    for (int i = 0; i < 6; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
    return r;
}
__device__ __forceinline__ Dual6 operator+(float a, const Dual6 &b) {
    Dual6 r = b;
    r.v += a;
    return r;
}
__device__ __forceinline__ Dual6 operator-(float a, const Dual6 &b) {
    Dual6 r = -b;
    r.v += a;
    return r;
}
__device__ __forceinline__ Dual6 operator*(float a, const Dual6 &b) {
    Dual6 r;
    r.v = a * b.v;
#pragma unroll
    for (int i = 0; i < 6; ++i) r.d[i] = a * b.d[i];
    return r;
}
__device__ __forceinline__ Dual6 dsqrt(const Dual6 &a) {
    Dual6 r;
    r.v = sqrtf(a.v);
    const float s = 0.5f / r.v;
#pragma unroll
    for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * s;
    return r;
}
__device__ __forceinline__ float dsqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ Dual6 datan2(const Dual6 &y, const Dual6 &x) {
    Dual6 r;
    r.v = atan2f(y.v, x.v);
    const float inv = 1.0f / (x.v * x.v + y.v * y.v);
#pragma unroll
    for (int i = 0; i < 6; ++i) r.d[i] = (x.v * y.d[i] - y.v * x.d[i]) * inv;
    return r;
}
__device__ __forceinline__ float datan2(float y, float x) { return atan2f(y, x); }
// F.normalize denominator: max(||v||, 1e-12); the clamp branch has zero derivative
__device__ __forceinline__ Dual6 clamp_min(const Dual6 &a, float lo) { return a.v < lo ? mk(lo) : a; }
__device__ __forceinline__ float clamp_min(float a, float lo) { return a < lo ? lo : a; }

template <typename S>
__device__ __forceinline__ S mk1();
template <>
__device__ __forceinline__ float mk1<float>() { return 1.0f; }
template <>
__device__ __forceinline__ Dual6 mk1<Dual6>() { return mk(1.0f); }

// ---------------------------------------------------------------------------------------------
// rot6d (view(-1,3,2): x[2r + c] = element (row r, column c)) -> axis-angle
// ---------------------------------------------------------------------------------------------
template <typename S>
__device__ __forceinline__ void rot6d_to_aa(const S (&x)[6], S (&aa)[3]) {
    // ContinousRotReprDecoder.decode (cvae.py:71-81)
    const S a0 = x[0], a1 = x[2], a2 = x[4], c0 = x[1], c1 = x[3], c2 = x[5];
    const S na = clamp_min(dsqrt(a0 * a0 + a1 * a1 + a2 * a2), 1e-12f);
    const S b10 = a0 / na, b11 = a1 / na, b12 = a2 / na;
    const S dot = b10 * c0 + b11 * c1 + b12 * c2;
    const S u0 = c0 - dot * b10, u1 = c1 - dot * b11, u2 = c2 - dot * b12;
    const S nu = clamp_min(dsqrt(u0 * u0 + u1 * u1 + u2 * u2), 1e-12f);
    const S b20 = u0 / nu, b21 = u1 / nu, b22 = u2 / nu;
    const S b30 = b11 * b22 - b12 * b21, b31 = b12 * b20 - b10 * b22, b32 = b10 * b21 - b11 * b20;
    // R[r][c] = b_c[r]; torchgeometry works on m = R^T: m[i][j] = R[j][i] = b_i[j]
    const S m00 = b10, m01 = b11, m02 = b12, m10 = b20, m11 = b21, m12 = b22, m20 = b30, m21 = b31, m22 = b32;
    // [3P] torchgeometry.rotation_matrix_to_quaternion (eps = 1e-6)
    const bool d2 = val(m22) < 1e-6f, d0_d1 = val(m00) > val(m11), d0_nd1 = val(m00) < -val(m11);
    S qw, qx, qy, qz, t;
    if (d2 && d0_d1) {
        t = (1.0f + m00) - m11 - m22;
        qw = m12 - m21; qx = t; qy = m01 + m10; qz = m20 + m02;
    } else if (d2) {
        t = (1.0f - m00) + m11 - m22;
        qw = m20 - m02; qx = m01 + m10; qy = t; qz = m12 + m21;
    } else if (d0_nd1) {
        t = (1.0f - m00) - m11 + m22;
        qw = m01 - m10; qx = m20 + m02; qy = m12 + m21; qz = t;
    } else {
        t = (1.0f + m00) + m11 + m22;
        qw = t; qx = m12 - m21; qy = m20 - m02; qz = m01 - m10;
    }
    const S h = 0.5f * (mk1<S>() / dsqrt(t));
    qw = qw * h; qx = qx * h; qy = qy * h; qz = qz * h;
    // [3P] torchgeometry.quaternion_to_angle_axis
    const S s2 = qx * qx + qy * qy + qz * qz;
    if (val(s2) > 0.0f) {
        const S s = dsqrt(s2);
        const S two_theta = 2.0f * (val(qw) < 0.0f ? datan2(-s, -qw) : datan2(s, qw));
        const S k = two_theta / s;
        aa[0] = qx * k; aa[1] = qy * k; aa[2] = qz * k;
    } else {
        aa[0] = 2.0f * qx; aa[1] = 2.0f * qy; aa[2] = 2.0f * qz;
    }
}

__device__ __forceinline__ void rot6d_to_aa_f(const float *in6, float *aa) {
    float x[6], o[3];
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = in6[i];
    rot6d_to_aa<float>(x, o);
    aa[0] = o[0]; aa[1] = o[1]; aa[2] = o[2];
}

// g_in[j] = sum_i g_aa[i] * d aa_i / d x_j
__device__ __forceinline__ void rot6d_to_aa_vjp(const float *in6, const float *g_aa, float *g_in) {
    Dual6 x[6], o[3];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        x[i] = mk(in6[i]);
        x[i].d[i] = 1.0f;
    }
    rot6d_to_aa<Dual6>(x, o);
#pragma unroll
    for (int j = 0; j < 6; ++j) g_in[j] = g_aa[0] * o[0].d[j] + g_aa[1] * o[1].d[j] + g_aa[2] * o[2].d[j];
}

__global__ void rot6d_to_aa_kernel(const float *__restrict__ in6, int64_t n, float *__restrict__ aa) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) rot6d_to_aa_f(in6 + 6 * i, aa + 3 * i);
}

__global__ void rot6d_to_aa_bwd_kernel(const float *__restrict__ in6, int64_t n, const float *__restrict__ g_aa,
                                       float *__restrict__ g_in) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) rot6d_to_aa_vjp(in6 + 6 * i, g_aa + 3 * i, g_in + 6 * i);
}

// [3P] torchgeometry.angle_axis_to_rotation_matrix, first two columns, row-major (3,2)
__global__ void aa_to_rot6d_kernel(const float *__restrict__ aa, int64_t n, float *__restrict__ out6) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float rx = aa[3 * i], ry = aa[3 * i + 1], rz = aa[3 * i + 2];
    const float theta2 = rx * rx + ry * ry + rz * rz;
    float r00, r01, r10, r11, r20, r21;
    if (theta2 > 1e-6f) {
        const float theta = sqrtf(theta2);
        const float inv = 1.0f / (theta + 1e-6f);
        const float wx = rx * inv, wy = ry * inv, wz = rz * inv;
        const float c = cosf(theta), s = sinf(theta), k = 1.0f - c;
        r00 = c + wx * wx * k;
        r10 = wz * s + wx * wy * k;
        r20 = -wy * s + wx * wz * k;
        r01 = wx * wy * k - wz * s;
        r11 = c + wy * wy * k;
        r21 = wx * s + wy * wz * k;
    } else {  // first-order Taylor branch
        r00 = 1.0f; r01 = -rz;
        r10 = rz;   r11 = 1.0f;
        r20 = -ry;  r21 = rx;
    }
    float *o = out6 + 6 * i;
    o[0] = r00; o[1] = r01; o[2] = r10; o[3] = r11; o[4] = r20; o[5] = r21;
}

// ---------------------------------------------------------------------------------------------
// VPoser decoder: one CTA per frame, one thread per hidden unit; weights are read in the layout that makes
// the accesses of a warp contiguous (transposed copies forward, nn.Linear's own [out][in] backward).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float lrelu(float x) { return x > 0.f ? x : 0.2f * x; }

__global__ void vposer_fwd_kernel(const fpv_vposer_model m, const float *__restrict__ z, float *__restrict__ aa,
                                  float *__restrict__ saved) {
    extern __shared__ float sm[];
    const int H = m.hidden, Z = m.latent, O = 6 * m.joints;
    float *sz = sm, *h1 = sz + Z, *h2 = h1 + H, *y = h2 + H;
    const int64_t t = blockIdx.x;
    const int o = threadIdx.x;
    if (o < Z) sz[o] = z[t * Z + o];
    __syncthreads();
    if (o < H) {
        float acc = m.b1[o];
        for (int k = 0; k < Z; ++k) acc = fmaf(m.w1t[int64_t(k) * H + o], sz[k], acc);
        h1[o] = lrelu(acc);
    }
    __syncthreads();
    if (o < H) {
        float acc = m.b2[o];
        for (int k = 0; k < H; ++k) acc = fmaf(m.w2t[int64_t(k) * H + o], h1[k], acc);
        h2[o] = lrelu(acc);
    }
    __syncthreads();
    if (o < O) {
        float acc = m.b3[o];
        for (int k = 0; k < H; ++k) acc = fmaf(m.w3t[int64_t(k) * O + o], h2[k], acc);
        y[o] = acc;
    }
    __syncthreads();
    if (o < m.joints) rot6d_to_aa_f(y + 6 * o, aa + (t * m.joints + o) * 3);
    // saved for the backward: post-activations (leaky ReLU keeps the sign) and the 6D output
    float *sv = saved + t * int64_t(2 * H + O);
    if (o < H) {
        sv[o] = h1[o];
        sv[H + o] = h2[o];
    }
    if (o < O) sv[2 * H + o] = y[o];
}

__global__ void vposer_bwd_kernel(const fpv_vposer_model m, const float *__restrict__ saved,
                                  const float *__restrict__ g_aa, float *__restrict__ g_z) {
    extern __shared__ float sm[];
    const int H = m.hidden, Z = m.latent, O = 6 * m.joints;
    float *gy = sm, *g2 = gy + O, *g1 = g2 + H;
    const int64_t t = blockIdx.x;
    const int o = threadIdx.x;
    const float *sv = saved + t * int64_t(2 * H + O);
    if (o < m.joints) rot6d_to_aa_vjp(sv + 2 * H + 6 * o, g_aa + (t * m.joints + o) * 3, gy + 6 * o);
    __syncthreads();
    if (o < H) {
        float acc = 0.f;
        for (int k = 0; k < O; ++k) acc = fmaf(m.w3[int64_t(k) * H + o], gy[k], acc);
        g2[o] = sv[H + o] > 0.f ? acc : 0.2f * acc;
    }
    __syncthreads();
    if (o < H) {
        float acc = 0.f;
        for (int k = 0; k < H; ++k) acc = fmaf(m.w2[int64_t(k) * H + o], g2[k], acc);
        g1[o] = sv[o] > 0.f ? acc : 0.2f * acc;
    }
    __syncthreads();
    if (o < Z) {
        float acc = 0.f;
        for (int k = 0; k < H; ++k) acc = fmaf(m.w1[int64_t(k) * Z + o], g1[k], acc);
        g_z[t * Z + o] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// DCT prior: x [NB*F][C], basis [F][K], coef [NB][C][K];  loss = mean_{NB*C} sum_f e/(e+1), e = (x - basis.coef)^2
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float dct_residual(const float *__restrict__ x, const float *__restrict__ basis,
                                              const float *__restrict__ coef, int64_t F, int64_t C, int64_t K, int64_t e) {
    const int64_t row = e / C, c = e - row * C;
    const int64_t nb = row / F, f = row - nb * F;
    const float *cf = coef + (nb * C + c) * K, *bs = basis + f * K;
    float hat = 0.f;
    for (int64_t k = 0; k < K; ++k) hat = fmaf(bs[k], cf[k], hat);
    return __fsub_rn(x[e], hat);
}

__global__ void dct_partial_kernel(const float *__restrict__ x, const float *__restrict__ basis,
                                   const float *__restrict__ coef, int64_t NB, int64_t F, int64_t C, int64_t K,
                                   double *partial) {
    __shared__ double scratch[32];
    const int64_t n = NB * F * C;
    double s = 0.0;
    for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += int64_t(gridDim.x) * blockDim.x) {
        const float r = dct_residual(x, basis, coef, F, C, K, e);
        const float err = __fmul_rn(r, r);
        s += double(__fdiv_rn(err, __fadd_rn(err, 1.0f)));
    }
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void dct_final_kernel(const double *__restrict__ partial, int nblk, double inv_count, float *out) {
    __shared__ double scratch[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < nblk; i += blockDim.x) s += partial[i];
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) out[0] = float(s * inv_count);
}

// d/dr [ r^2 / (r^2 + 1) ] = 2 r / (r^2 + 1)^2
__global__ void dct_bwd_x_kernel(const float *__restrict__ x, const float *__restrict__ basis,
                                 const float *__restrict__ coef, int64_t NB, int64_t F, int64_t C, int64_t K,
                                 const float *__restrict__ g_out, float *__restrict__ grad_x) {
    const int64_t n = NB * F * C;
    const float gs = g_out[0] / float(NB * C);
    for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += int64_t(gridDim.x) * blockDim.x) {
        const float r = dct_residual(x, basis, coef, F, C, K, e);
        const float q = fmaf(r, r, 1.0f);
        grad_x[e] = gs * 2.0f * r / (q * q);
    }
}

// one thread per coefficient: fixed-order sum over the F frames of its window
__global__ void dct_bwd_coef_kernel(const float *__restrict__ x, const float *__restrict__ basis,
                                    const float *__restrict__ coef, int64_t NB, int64_t F, int64_t C, int64_t K,
                                    const float *__restrict__ g_out, float *__restrict__ grad_coef) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= NB * C * K) return;
    const int64_t k = i % K, c = (i / K) % C, nb = i / (K * C);
    const float gs = g_out[0] / float(NB * C);
    float acc = 0.f;
    for (int64_t f = 0; f < F; ++f) {
        const int64_t e = (nb * F + f) * C + c;
        const float r = dct_residual(x, basis, coef, F, C, K, e);
        const float q = fmaf(r, r, 1.0f);
        acc = fmaf(-2.0f * r / (q * q), basis[f * K + k], acc);
    }
    grad_coef[i] = gs * acc;
}

}  // namespace fpv

using namespace fpv;

extern "C" {

int fpv_rot6d_to_aa_fwd(const float *in6, int64_t n, float *aa, fpv_stream_t stream) {
    FPV_CHECK_ARG(in6 && aa && n > 0, "fpv_rot6d_to_aa_fwd: empty input");
    rot6d_to_aa_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(in6, n, aa);
    FPV_LAUNCH_CHECK("rot6d_to_aa_kernel");
    return FPV_OK;
}

int fpv_rot6d_to_aa_bwd(const float *in6, int64_t n, const float *g_aa, float *g_in6, fpv_stream_t stream) {
    FPV_CHECK_ARG(in6 && g_aa && g_in6 && n > 0, "fpv_rot6d_to_aa_bwd: empty input");
    rot6d_to_aa_bwd_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(in6, n, g_aa, g_in6);
    FPV_LAUNCH_CHECK("rot6d_to_aa_bwd_kernel");
    return FPV_OK;
}

int fpv_aa_to_rot6d(const float *aa, int64_t n, float *out6, fpv_stream_t stream) {
    FPV_CHECK_ARG(aa && out6 && n > 0, "fpv_aa_to_rot6d: empty input");
    aa_to_rot6d_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(aa, n, out6);
    FPV_LAUNCH_CHECK("aa_to_rot6d_kernel");
    return FPV_OK;
}

static int vposer_check(const fpv_vposer_model *m, const char *who) {
    FPV_CHECK_ARG(m && m->w1 && m->w2 && m->w3 && m->w1t && m->w2t && m->w3t && m->b1 && m->b2 && m->b3,
                  "%s: null weight pointer", who);
    FPV_CHECK_ARG(m->latent > 0 && m->latent <= 1024 && m->hidden > 0 && m->hidden <= 1024 && m->joints > 0 &&
                      6 * m->joints <= m->hidden && m->latent <= m->hidden,
                  "%s: unsupported decoder shape (latent %d, hidden %d, joints %d)", who, m->latent, m->hidden, m->joints);
    return FPV_OK;
}

size_t fpv_vposer_saved_floats(const fpv_vposer_model *m, int64_t T) {
    if (!m || T <= 0) return 0;
    return size_t(T) * size_t(2 * m->hidden + 6 * m->joints);
}

int fpv_vposer_decode_fwd(const fpv_vposer_model *m, const float *z, int64_t T, float *aa, float *saved,
                          fpv_stream_t stream) {
    if (int rc = vposer_check(m, "fpv_vposer_decode_fwd")) return rc;
    FPV_CHECK_ARG(z && aa && saved && T > 0, "fpv_vposer_decode_fwd: empty input");
    const int threads = int(align_up(size_t(m->hidden), 32));
    const size_t smem = size_t(m->latent + 2 * m->hidden + 6 * m->joints) * sizeof(float);
    vposer_fwd_kernel<<<(unsigned)T, threads, smem, static_cast<cudaStream_t>(stream)>>>(*m, z, aa, saved);
    FPV_LAUNCH_CHECK("vposer_fwd_kernel");
    return FPV_OK;
}

int fpv_vposer_decode_bwd(const fpv_vposer_model *m, const float *saved, const float *g_aa, int64_t T, float *g_z,
                          fpv_stream_t stream) {
    if (int rc = vposer_check(m, "fpv_vposer_decode_bwd")) return rc;
    FPV_CHECK_ARG(saved && g_aa && g_z && T > 0, "fpv_vposer_decode_bwd: empty input");
    const int threads = int(align_up(size_t(m->hidden), 32));
    const size_t smem = size_t(2 * m->hidden + 6 * m->joints) * sizeof(float);
    vposer_bwd_kernel<<<(unsigned)T, threads, smem, static_cast<cudaStream_t>(stream)>>>(*m, saved, g_aa, g_z);
    FPV_LAUNCH_CHECK("vposer_bwd_kernel");
    return FPV_OK;
}

static int dct_blocks(int64_t n) {
    const int64_t b = ceil_div(n, 256 * 4);
    return int(b < 1 ? 1 : (b > 1184 ? 1184 : b));
}

size_t fpv_dct_prior_workspace_bytes(int64_t NB, int64_t F, int64_t C) {
    if (NB <= 0 || F <= 0 || C <= 0) return 0;
    return align_up(size_t(dct_blocks(NB * F * C)) * sizeof(double), 256);
}

int fpv_dct_prior_fwd(const float *x, const float *basis, const float *coef, int64_t NB, int64_t F, int64_t C, int64_t K,
                      float *out, void *workspace, size_t workspace_bytes, fpv_stream_t stream) {
    FPV_CHECK_ARG(x && basis && coef && out, "fpv_dct_prior_fwd: null pointer");
    FPV_CHECK_ARG(NB > 0 && F > 0 && C > 0 && K > 0, "fpv_dct_prior_fwd: empty input");
    FPV_CHECK_ARG(workspace && workspace_bytes >= fpv_dct_prior_workspace_bytes(NB, F, C), "fpv_dct_prior_fwd: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nb = dct_blocks(NB * F * C);
    double *partial = static_cast<double *>(workspace);
    dct_partial_kernel<<<nb, 256, 0, st>>>(x, basis, coef, NB, F, C, K, partial);
    FPV_LAUNCH_CHECK("dct_partial_kernel");
    dct_final_kernel<<<1, 256, 0, st>>>(partial, nb, 1.0 / double(NB * C), out);
    FPV_LAUNCH_CHECK("dct_final_kernel");
    return FPV_OK;
}

int fpv_dct_prior_bwd(const float *x, const float *basis, const float *coef, int64_t NB, int64_t F, int64_t C, int64_t K,
                      const float *g_out, float *grad_x, float *grad_coef, fpv_stream_t stream) {
    FPV_CHECK_ARG(x && basis && coef && g_out, "fpv_dct_prior_bwd: null pointer");
    FPV_CHECK_ARG(NB > 0 && F > 0 && C > 0 && K > 0, "fpv_dct_prior_bwd: empty input");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (grad_x) {
        dct_bwd_x_kernel<<<dct_blocks(NB * F * C), 256, 0, st>>>(x, basis, coef, NB, F, C, K, g_out, grad_x);
        FPV_LAUNCH_CHECK("dct_bwd_x_kernel");
    }
    if (grad_coef) {
        dct_bwd_coef_kernel<<<(unsigned)ceil_div(NB * C * K, 128), 128, 0, st>>>(x, basis, coef, NB, F, C, K, g_out, grad_coef);
        FPV_LAUNCH_CHECK("dct_bwd_coef_kernel");
    }
    return FPV_OK;
}

}  // extern "C"
