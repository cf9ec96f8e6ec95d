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
// VPoser decoder: one CTA per FR consecutive frames, one thread per hidden unit; weights are read in the layout that
// makes the accesses of a warp contiguous (transposed copies forward, nn.Linear's own [out][in] backward).  The kernel is
// bound by the L2 reads of the 1.3 MB of weights per CTA: a weight fetched once serves FR frames (FR accumulators per
// thread, each frame's own k-ascending FMA chain: the result of a frame does not depend on FR).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float lrelu(float x) { return x > 0.f ? x : 0.2f * x; }

// acc[f] += sum_{k0 <= k < k1} w[k * ld + o] * s[f * S + k]  (k ascending; 16 independent weight loads in flight per
// thread: the loops are bound by the L2 latency of the weight fetches, not by bandwidth or arithmetic)
template <int FR>
__device__ __forceinline__ void vp_dot(const float *__restrict__ w, int ld, int o, int k0, int k1, const float *s, int S,
                                       float (&acc)[FR]) {
#pragma unroll 16
    for (int k = k0; k < k1; ++k) {
        const float wv = __ldg(w + int64_t(k) * ld + o);
#pragma unroll
        for (int f = 0; f < FR; ++f) acc[f] = fmaf(wv, s[f * S + k], acc[f]);
    }
}

// A layer with few outputs (NO << blockDim) would leave most of the CTA idle behind a K-long serial loop: the thread
// groups split K instead, partial sums meet in shared memory and are added in group order (fixed: deterministic).
// Returns through out[f * S_out + o] for o < NO.  `part` holds blockDim * FR floats.  Ends with a barrier.
template <int FR>
__device__ __forceinline__ void vp_layer_split(const float *__restrict__ w, const float *__restrict__ bias, int NO, int K,
                                               const float *s, int S, float *part, float *out, int S_out) {
    const int o = threadIdx.x, nthr = blockDim.x;
    const int NOP = (NO + 31) & ~31;
    const int G = nthr / NOP > 0 ? nthr / NOP : 1;
    const int g = o / NOP, oo = o - g * NOP;
    if (g < G && oo < NO) {
        const int chunk = (K + G - 1) / G;
        const int k0 = g * chunk, k1 = k0 + chunk < K ? k0 + chunk : K;
        float acc[FR];
#pragma unroll
        for (int f = 0; f < FR; ++f) acc[f] = (g == 0 && bias) ? bias[oo] : 0.f;
        vp_dot<FR>(w, NO, oo, k0, k1, s, S, acc);
#pragma unroll
        for (int f = 0; f < FR; ++f) part[f * nthr + o] = acc[f];
    }
    __syncthreads();
    if (o < NO) {
#pragma unroll
        for (int f = 0; f < FR; ++f) {
            float v = part[f * nthr + o];
            for (int gg = 1; gg < G; ++gg) v += part[f * nthr + gg * NOP + o];
            out[f * S_out + o] = v;
        }
    }
    __syncthreads();
}

template <int FR>
__global__ void vposer_fwd_kernel(const fpv_vposer_model m, const float *__restrict__ z, float *__restrict__ aa,
                                  float *__restrict__ saved, int64_t T) {
    extern __shared__ float sm[];
    const int H = m.hidden, Z = m.latent, O = 6 * m.joints;
    const int S = Z + 2 * H + O;  // per frame: sz[Z], h1[H], h2[H], y[O]
    float *part = sm + FR * S;    // blockDim * FR partial sums of the split output layer
    const int64_t t0 = int64_t(blockIdx.x) * FR;
    const int nf = int(T - t0 < FR ? T - t0 : FR);
    const int o = threadIdx.x;
    for (int i = o; i < FR * Z; i += blockDim.x) {
        const int f = i / Z, k = i - f * Z;
        sm[f * S + k] = f < nf ? z[(t0 + f) * Z + k] : 0.f;
    }
    __syncthreads();
    float acc[FR];
    if (o < H) {
#pragma unroll
        for (int f = 0; f < FR; ++f) acc[f] = m.b1[o];
        vp_dot<FR>(m.w1t, H, o, 0, Z, sm, S, acc);
#pragma unroll
        for (int f = 0; f < FR; ++f) sm[f * S + Z + o] = lrelu(acc[f]);
    }
    __syncthreads();
    if (o < H) {
#pragma unroll
        for (int f = 0; f < FR; ++f) acc[f] = m.b2[o];
        vp_dot<FR>(m.w2t, H, o, 0, H, sm + Z, S, acc);
#pragma unroll
        for (int f = 0; f < FR; ++f) sm[f * S + Z + H + o] = lrelu(acc[f]);
    }
    __syncthreads();
    vp_layer_split<FR>(m.w3t, m.b3, O, H, sm + Z + H, S, part, sm + Z + 2 * H, S);
    for (int i = o; i < nf * m.joints; i += blockDim.x) {
        const int f = i / m.joints, j = i - f * m.joints;
        rot6d_to_aa_f(sm + f * S + Z + 2 * H + 6 * j, aa + ((t0 + f) * m.joints + j) * 3);
    }
    // saved for the backward: post-activations (leaky ReLU keeps the sign) and the 6D output
    for (int f = 0; f < nf; ++f) {
        float *sv = saved + (t0 + f) * int64_t(2 * H + O);
        const float *fr = sm + f * S + Z;
        if (o < H) {
            sv[o] = fr[o];
            sv[H + o] = fr[H + o];
        }
        if (o < O) sv[2 * H + o] = fr[2 * H + o];
    }
}

template <int FR>
__global__ void vposer_bwd_kernel(const fpv_vposer_model m, const float *__restrict__ saved,
                                  const float *__restrict__ g_aa, float *__restrict__ g_z, int64_t T) {
    extern __shared__ float sm[];
    const int H = m.hidden, Z = m.latent, O = 6 * m.joints;
    const int S = O + 2 * H + Z;  // per frame: gy[O], g2[H], g1[H], gz[Z]
    float *part = sm + FR * S;
    const int64_t t0 = int64_t(blockIdx.x) * FR;
    const int nf = int(T - t0 < FR ? T - t0 : FR);
    const int o = threadIdx.x;
    const int64_t svs = int64_t(2 * H + O);
    for (int i = o; i < FR * m.joints; i += blockDim.x) {
        const int f = i / m.joints, j = i - f * m.joints;
        if (f < nf) {
            rot6d_to_aa_vjp(saved + (t0 + f) * svs + 2 * H + 6 * j, g_aa + ((t0 + f) * m.joints + j) * 3, sm + f * S + 6 * j);
        } else {
            for (int c = 0; c < 6; ++c) sm[f * S + 6 * j + c] = 0.f;
        }
    }
    __syncthreads();
    float acc[FR];
    if (o < H) {
#pragma unroll
        for (int f = 0; f < FR; ++f) acc[f] = 0.f;
        vp_dot<FR>(m.w3, H, o, 0, O, sm, S, acc);
#pragma unroll
        for (int f = 0; f < FR; ++f)
            sm[f * S + O + o] = (f < nf && saved[(t0 + f) * svs + H + o] > 0.f) ? acc[f] : 0.2f * acc[f];
    }
    __syncthreads();
    if (o < H) {
#pragma unroll
        for (int f = 0; f < FR; ++f) acc[f] = 0.f;
        vp_dot<FR>(m.w2, H, o, 0, H, sm + O, S, acc);
#pragma unroll
        for (int f = 0; f < FR; ++f)
            sm[f * S + O + H + o] = (f < nf && saved[(t0 + f) * svs + o] > 0.f) ? acc[f] : 0.2f * acc[f];
    }
    __syncthreads();
    vp_layer_split<FR>(m.w1, nullptr, Z, H, sm + O + H, S, part, sm + O + 2 * H, S);
    for (int i = o; i < nf * Z; i += blockDim.x) {
        const int f = i / Z, k = i - f * Z;
        g_z[(t0 + f) * Z + k] = sm[f * S + O + 2 * H + k];
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

// one warp per (window, channel): the lanes take the F frames of the window (residual and its weight once per frame, not
// once per coefficient), then every coefficient is one fixed-tree warp sum over the frames (deterministic)
__global__ void __launch_bounds__(128) dct_bwd_coef_kernel(const float *__restrict__ x, const float *__restrict__ basis,
                                                           const float *__restrict__ coef, int64_t NB, int64_t F, int64_t C,
                                                           int64_t K, const float *__restrict__ g_out,
                                                           float *__restrict__ grad_coef) {
    const int lane = threadIdx.x & 31;
    const int64_t i = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);  // (nb, c)
    if (i >= NB * C) return;
    const int64_t c = i % C, nb = i / C;
    const float gs = g_out[0] / float(NB * C);
    for (int64_t k0 = 0; k0 < K; k0 += 32) {  // coefficients in chunks of 32: lane k keeps chunk coefficient k
        float mine = 0.f;
        for (int64_t f0 = 0; f0 < F; f0 += 32) {
            const int64_t f = f0 + lane;
            float w = 0.f;
            if (f < F) {
                const float r = dct_residual(x, basis, coef, F, C, K, (nb * F + f) * C + c);
                const float q = fmaf(r, r, 1.0f);
                w = -2.0f * r / (q * q);
            }
            const int64_t kn = (K - k0 < 32) ? K - k0 : 32;
            for (int64_t k = 0; k < kn; ++k) {
                const float s = __shfl_sync(0xffffffffu, warp_sum(f < F ? w * basis[f * K + k0 + k] : 0.f), 0);  // total in lane 0
                if (lane == k) mine += s;
            }
        }
        if (k0 + lane < K) grad_coef[i * K + k0 + lane] = gs * mine;
    }
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

// frames per CTA: as many as keep about one CTA per SM (the weights are re-read from L2 once per CTA)
static int vposer_frames_per_cta(int64_t T, size_t smem_per_frame) {
    const int64_t sms = sm_count();
    int fr = T >= 4 * sms ? 4 : (T >= 2 * sms - 8 ? 2 : 1);
    while (fr > 1 && fr * smem_per_frame > 48 * 1024) fr >>= 1;  // stay inside the default dynamic shared-memory limit
    return fr;
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
    // per frame: activations + one partial sum per thread for the split output layer
    const size_t smem = size_t(m->latent + 2 * m->hidden + 6 * m->joints + threads) * sizeof(float);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int fr = vposer_frames_per_cta(T, smem);
    if (fr == 4)
        vposer_fwd_kernel<4><<<(unsigned)ceil_div(T, 4), threads, 4 * smem, st>>>(*m, z, aa, saved, T);
    else if (fr == 2)
        vposer_fwd_kernel<2><<<(unsigned)ceil_div(T, 2), threads, 2 * smem, st>>>(*m, z, aa, saved, T);
    else
        vposer_fwd_kernel<1><<<(unsigned)T, threads, smem, st>>>(*m, z, aa, saved, T);
    FPV_LAUNCH_CHECK("vposer_fwd_kernel");
    return FPV_OK;
}

int fpv_vposer_decode_bwd(const fpv_vposer_model *m, const float *saved, const float *g_aa, int64_t T, float *g_z,
                          fpv_stream_t stream) {
    if (int rc = vposer_check(m, "fpv_vposer_decode_bwd")) return rc;
    FPV_CHECK_ARG(saved && g_aa && g_z && T > 0, "fpv_vposer_decode_bwd: empty input");
    const int threads = int(align_up(size_t(m->hidden), 32));
    const size_t smem = size_t(m->latent + 2 * m->hidden + 6 * m->joints + threads) * sizeof(float);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int fr = vposer_frames_per_cta(T, smem);
    if (fr == 4)
        vposer_bwd_kernel<4><<<(unsigned)ceil_div(T, 4), threads, 4 * smem, st>>>(*m, saved, g_aa, g_z, T);
    else if (fr == 2)
        vposer_bwd_kernel<2><<<(unsigned)ceil_div(T, 2), threads, 2 * smem, st>>>(*m, saved, g_aa, g_z, T);
    else
        vposer_bwd_kernel<1><<<(unsigned)T, threads, smem, st>>>(*m, saved, g_aa, g_z, T);
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
        dct_bwd_coef_kernel<<<(unsigned)ceil_div(NB * C, 4), 128, 0, st>>>(x, basis, coef, NB, F, C, K, g_out, grad_coef);
        FPV_LAUNCH_CHECK("dct_bwd_coef_kernel");
    }
    return FPV_OK;
}

}  // extern "C"
