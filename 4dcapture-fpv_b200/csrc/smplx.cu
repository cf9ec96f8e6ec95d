// smplx.cu -- SMPL-X body-model forward and backward (batched over T frames).  sm_100a.
//
// Replaces the [3P] smplx forward the reference calls every optimiser step
// (global_optimization.py:280-283, :333-335, :396-398: ~200 torch-eager launches, a 54-step Python
// loop for the kinematic chain, a materialised [T,V,4,4] skinning tensor).  Here it is four launches:
//   1. pose_fwd      one CTA per frame: hand PCA, Rodrigues, rest joints from betas, the kinematic
//                    chain (level-synchronous in shared memory), skinning transforms A, and the
//                    512-wide coefficient row  [R_1..R_54 - I | betas,expr | 1 | 0].
//   2. blend GEMM    v_posed[T,3V] = coef[T,512] x basis[512,3V].  One contraction covers pose
//                    blend shapes (486 rows), shape blend shapes (20 rows) AND the template (the
//                    row multiplied by the constant 1) -- smplx.lbs steps 2 and 5 in one pass.
//   3. skin_fwd      per (frame, vertex): blend <= W joint transforms (ELL), apply, add transl.
//   4. extras        vertex-picked extra joints (VertexJointSelector).
// Backward mirrors it: skin_bwd (per vertex, R^T g) -> jointgrad (per (joint, frame) list reduction,
// fixed order) -> blend GEMM^T (split-K, fixed-order reduce) -> pose_bwd (chain reverse, Rodrigues
// backward, PCA backward).  Every reduction has a fixed order: gradients are run-to-run bitwise stable.
#include "common.cuh"

namespace fpv {

constexpr int NJ = FPV_SMPLX_JOINTS;
constexpr int KP = FPV_SMPLX_KPAD;
constexpr int NPF = FPV_SMPLX_POSE_FEAT;
constexpr int NSH = FPV_SMPLX_SHAPE;
constexpr int NTH = FPV_SMPLX_THETA;
constexpr int TH_LH = 75, TH_RH = 87, TH_BETA = 99, TH_TRANSL = 119;

int tc_gemm_3xtf32(const float *a_hi, const float *a_lo, int64_t lda, const float *b_hi, const float *b_lo,
                   int64_t ldb, int M, int N, int K, float *C, int64_t ldc, int ksplit, float *partial,
                   cudaStream_t st);  // tc_gemm.cu

// row pitch (floats) of every [*, 3V] operand of the tensor-core GEMMs: TMA needs 16-byte multiples
static inline int64_t pitch3v(int64_t V) { return (3 * V + 3) / 4 * 4; }

struct SavedLayout {
    size_t R, G, J, A, coef_hi, coef_lo, vposed, total;  // float offsets
};
static SavedLayout saved_layout(int64_t T, int64_t V) {
    SavedLayout L;
    size_t o = 0;
    auto take = [&](size_t n) {
        size_t r = o;
        o += align_up(n, 64);
        return r;
    };
    L.R = take(size_t(T) * NJ * 9);
    L.G = take(size_t(T) * NJ * 12);
    L.J = take(size_t(T) * NJ * 3);
    L.A = take(size_t(T) * NJ * 12);
    L.coef_hi = take(size_t(T) * KP);
    L.coef_lo = take(size_t(T) * KP);
    L.vposed = take(size_t(T) * pitch3v(V));
    L.total = o;
    return L;
}

__device__ __forceinline__ float tf32_rn(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// ---------------------------------------------------------------------------------------------
// 1. per-frame pose kernel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void rodrigues(const float *r, float *R) {
    const float eps = 1e-8f;
    const float ax = r[0] + eps, ay = r[1] + eps, az = r[2] + eps;
    const float angle = sqrtf(ax * ax + ay * ay + az * az);
    const float x = r[0] / angle, y = r[1] / angle, z = r[2] / angle;
    const float s = sinf(angle), omc = 1.0f - cosf(angle);
    // K = [[0,-z,y],[z,0,-x],[-y,x,0]] ; R = I + s K + (1-c) K K
    R[0] = 1.f + omc * (-(y * y + z * z));
    R[1] = s * (-z) + omc * (x * y);
    R[2] = s * (y) + omc * (x * z);
    R[3] = s * (z) + omc * (x * y);
    R[4] = 1.f + omc * (-(x * x + z * z));
    R[5] = s * (-x) + omc * (y * z);
    R[6] = s * (-y) + omc * (x * z);
    R[7] = s * (x) + omc * (y * z);
    R[8] = 1.f + omc * (-(x * x + y * y));
}

__global__ void __launch_bounds__(64) pose_fwd_kernel(const fpv_smplx_model_t m, const float *__restrict__ theta,
                                                      float *__restrict__ sR, float *__restrict__ sG,
                                                      float *__restrict__ sJ, float *__restrict__ sA,
                                                      float *__restrict__ coef_hi, float *__restrict__ coef_lo,
                                                      float *__restrict__ joints, int joints_stride) {
    __shared__ float th[NTH];
    __shared__ float fp[165];
    __shared__ float R[NJ][9];
    __shared__ float J[NJ][3];
    __shared__ float G[NJ][12];
    __shared__ int depth[NJ];
    __shared__ int par[NJ];
    __shared__ int maxdepth;
    const int64_t t = blockIdx.x;
    const int tid = threadIdx.x;
    for (int k = tid; k < NTH; k += 64) th[k] = theta[t * NTH + k];
    if (tid < NJ) par[tid] = m.parents[tid];
    __syncthreads();
    if (tid == 0) {
        int md = 0;
        depth[0] = 0;
        for (int j = 1; j < NJ; ++j) {
            depth[j] = depth[par[j]] + 1;
            md = depth[j] > md ? depth[j] : md;
        }
        maxdepth = md;
    }
    for (int k = tid; k < 165; k += 64) {
        float v;
        if (k < 75) {
            v = th[k];
        } else {
            const int hand = (k >= 120);
            const int c = k - (hand ? 120 : 75);
            const float *comp = m.hand_comps + hand * 12 * 45;
            const float *pc = th + (hand ? TH_RH : TH_LH);
            v = 0.f;
            for (int i = 0; i < 12; ++i) v = fmaf(pc[i], comp[i * 45 + c], v);
        }
        fp[k] = v + m.pose_mean[k];
    }
    __syncthreads();
    if (tid < NJ) {
        rodrigues(&fp[3 * tid], R[tid]);
        for (int c = 0; c < 3; ++c) {
            float v = m.j_template[tid * 3 + c];
            const float *sd = m.j_shapedirs + (tid * 3 + c) * NSH;
            for (int l = 0; l < NSH; ++l) v = fmaf(sd[l], th[TH_BETA + l], v);
            J[tid][c] = v;
        }
    }
    __syncthreads();
    for (int lvl = 0; lvl <= maxdepth; ++lvl) {
        if (tid < NJ && depth[tid] == lvl) {
            const int j = tid;
            if (j == 0) {
                for (int r = 0; r < 3; ++r) {
                    for (int c = 0; c < 3; ++c) G[0][4 * r + c] = R[0][3 * r + c];
                    G[0][4 * r + 3] = J[0][r];
                }
            } else {
                const int p = par[j];
                const float rel[3] = {J[j][0] - J[p][0], J[j][1] - J[p][1], J[j][2] - J[p][2]};
                for (int r = 0; r < 3; ++r) {
                    const float g0 = G[p][4 * r], g1 = G[p][4 * r + 1], g2 = G[p][4 * r + 2];
                    for (int c = 0; c < 3; ++c)
                        G[j][4 * r + c] = g0 * R[j][c] + g1 * R[j][3 + c] + g2 * R[j][6 + c];
                    G[j][4 * r + 3] = g0 * rel[0] + g1 * rel[1] + g2 * rel[2] + G[p][4 * r + 3];
                }
            }
        }
        __syncthreads();
    }
    if (tid < NJ) {
        const int j = tid;
        float *a = sA + (t * NJ + j) * 12;
        float *g = sG + (t * NJ + j) * 12;
        for (int r = 0; r < 3; ++r) {
            const float g0 = G[j][4 * r], g1 = G[j][4 * r + 1], g2 = G[j][4 * r + 2], gt = G[j][4 * r + 3];
            a[4 * r] = g0;
            a[4 * r + 1] = g1;
            a[4 * r + 2] = g2;
            a[4 * r + 3] = gt - (g0 * J[j][0] + g1 * J[j][1] + g2 * J[j][2]);
            g[4 * r] = g0;
            g[4 * r + 1] = g1;
            g[4 * r + 2] = g2;
            g[4 * r + 3] = gt;
            joints[(t * joints_stride + j) * 3 + r] = gt + th[TH_TRANSL + r];
            sJ[(t * NJ + j) * 3 + r] = J[j][r];
        }
        for (int k = 0; k < 9; ++k) sR[(t * NJ + j) * 9 + k] = R[j][k];
    }
    for (int k = tid; k < KP; k += 64) {
        float v = 0.f;
        if (k < NPF) {
            const int j = 1 + k / 9, e = k % 9;
            v = R[j][e] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f);
        } else if (k < NPF + NSH) {
            v = th[TH_BETA + (k - NPF)];
        } else if (k == NPF + NSH) {
            v = 1.f;
        }
        const float h = tf32_rn(v);  // TF32 hi/lo split feeding the 3xTF32 tensor-core contraction
        coef_hi[t * KP + k] = h;
        coef_lo[t * KP + k] = tf32_rn(v - h);
    }
}

// ---------------------------------------------------------------------------------------------
// 3. skinning
// ---------------------------------------------------------------------------------------------
constexpr int SKIN_THREADS = 128;

__device__ __forceinline__ void blend_transform(const float (*A)[12], const int32_t *__restrict__ ell_joint,
                                                const float *__restrict__ ell_weight, int W, int V, int v,
                                                float *Tm) {
#pragma unroll
    for (int k = 0; k < 12; ++k) Tm[k] = 0.f;
    for (int w = 0; w < W; ++w) {
        const int j = ell_joint[size_t(w) * V + v];
        if (j < 0) continue;
        const float wt = ell_weight[size_t(w) * V + v];
#pragma unroll
        for (int k = 0; k < 12; ++k) Tm[k] = fmaf(wt, A[j][k], Tm[k]);
    }
}

__global__ void __launch_bounds__(SKIN_THREADS) skin_fwd_kernel(const fpv_smplx_model_t m,
                                                                const float *__restrict__ theta,
                                                                const float *__restrict__ sA,
                                                                const float *__restrict__ vposed, int64_t ldp,
                                                                float *__restrict__ verts) {
    __shared__ float A[NJ][12];
    __shared__ float tr[3];
    const int64_t t = blockIdx.y;
    const int V = m.num_verts;
    for (int k = threadIdx.x; k < NJ * 12; k += SKIN_THREADS) (&A[0][0])[k] = sA[t * NJ * 12 + k];
    if (threadIdx.x < 3) tr[threadIdx.x] = theta[t * NTH + TH_TRANSL + threadIdx.x];
    __syncthreads();
    const int v = blockIdx.x * SKIN_THREADS + threadIdx.x;
    if (v >= V) return;
    float Tm[12];
    blend_transform(A, m.ell_joint, m.ell_weight, m.ell_width, V, v, Tm);
    const float *vp = vposed + t * ldp + 3 * v;
    const float x = vp[0], y = vp[1], z = vp[2];
    float *o = verts + (t * V + v) * 3;
#pragma unroll
    for (int r = 0; r < 3; ++r) o[r] = Tm[4 * r] * x + Tm[4 * r + 1] * y + Tm[4 * r + 2] * z + Tm[4 * r + 3] + tr[r];
}

__global__ void extras_fwd_kernel(const fpv_smplx_model_t m, const float *__restrict__ verts,
                                  float *__restrict__ joints, int joints_stride) {
    const int64_t t = blockIdx.x;
    const int e = threadIdx.x;
    if (e >= m.num_extra) return;
    const int v = m.extra_vertex_ids[e];
    for (int c = 0; c < 3; ++c)
        joints[(t * joints_stride + NJ + e) * 3 + c] = verts[(t * m.num_verts + v) * 3 + c];
}

// effective vertex gradient: g_vertices + the gradient of the extra joints picked from this vertex.  `mask` is a CTA-shared
// bitmap of the vertices that carry an extra joint (ex_mask_build): only those walk the list of extras -- one bit test for
// the other 10,4xx vertices instead of E comparisons each (the list walk was 2/3 of jointgrad_kernel's instructions).
constexpr int EX_MASK_WORDS = 512;  // covers 16,384 vertices; vertices beyond it always walk the list
__device__ __forceinline__ void ex_mask_build(unsigned *mask, const int *ex_id, int E) {
    for (int k = threadIdx.x; k < EX_MASK_WORDS; k += blockDim.x) mask[k] = 0u;
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        const int v = ex_id[e];
        if (v >= 0 && (v >> 5) < EX_MASK_WORDS) atomicOr(&mask[v >> 5], 1u << (v & 31));
    }
    __syncthreads();
}
__device__ __forceinline__ void eff_grad(const float *__restrict__ g_verts, int64_t t, int V, int v, int E,
                                         const int *ex_id, const float (*ex_g)[3], const unsigned *mask, float *g) {
    if (g_verts) {
        const float *s = g_verts + (t * V + v) * 3;
        g[0] = s[0];
        g[1] = s[1];
        g[2] = s[2];
    } else {
        g[0] = g[1] = g[2] = 0.f;
    }
    if (E > 0 && ((v >> 5) >= EX_MASK_WORDS || ((mask[v >> 5] >> (v & 31)) & 1u))) {
        for (int e = 0; e < E; ++e) {
            if (ex_id[e] == v) {
                g[0] += ex_g[e][0];
                g[1] += ex_g[e][1];
                g[2] += ex_g[e][2];
            }
        }
    }
}

constexpr int MAX_EXTRA = 64;

__global__ void __launch_bounds__(SKIN_THREADS) skin_bwd_kernel(const fpv_smplx_model_t m,
                                                                const float *__restrict__ sA,
                                                                const float *__restrict__ g_verts,
                                                                const float *__restrict__ g_joints, int joints_stride,
                                                                int64_t ldp, float *__restrict__ g_vposed_hi,
                                                                float *__restrict__ g_vposed_lo) {
    __shared__ float A[NJ][12];
    __shared__ int ex_id[MAX_EXTRA];
    __shared__ float ex_g[MAX_EXTRA][3];
    __shared__ unsigned ex_mask[EX_MASK_WORDS];
    const int64_t t = blockIdx.y;
    const int V = m.num_verts;
    const int E = g_joints ? m.num_extra : 0;
    for (int k = threadIdx.x; k < NJ * 12; k += SKIN_THREADS) (&A[0][0])[k] = sA[t * NJ * 12 + k];
    for (int e = threadIdx.x; e < E; e += SKIN_THREADS) {
        ex_id[e] = m.extra_vertex_ids[e];
        for (int c = 0; c < 3; ++c) ex_g[e][c] = g_joints[(t * joints_stride + NJ + e) * 3 + c];
    }
    __syncthreads();
    ex_mask_build(ex_mask, ex_id, E);
    const int v = blockIdx.x * SKIN_THREADS + threadIdx.x;
    if (v >= V) return;
    float g[3];
    eff_grad(g_verts, t, V, v, E, ex_id, ex_g, ex_mask, g);
    float Tm[12];
    blend_transform(A, m.ell_joint, m.ell_weight, m.ell_width, V, v, Tm);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float gv = Tm[c] * g[0] + Tm[4 + c] * g[1] + Tm[8 + c] * g[2];
        const float h = tf32_rn(gv);
        g_vposed_hi[t * ldp + 3 * v + c] = h;
        g_vposed_lo[t * ldp + 3 * v + c] = tf32_rn(gv - h);
    }
}

// grid (56, T): blocks 0..54 reduce the influence list of one joint to gA[t][j][12] = sum w * g (x) [vp,1];
// block 55 reduces sum_v g -> gT[t][3].  Strided accumulation + fixed tree: deterministic.
__global__ void __launch_bounds__(128) jointgrad_kernel(const fpv_smplx_model_t m, const float *__restrict__ vposed,
                                                        int64_t ldp, const float *__restrict__ g_verts,
                                                        const float *__restrict__ g_joints, int joints_stride,
                                                        float *__restrict__ gA, float *__restrict__ gT) {
    __shared__ int ex_id[MAX_EXTRA];
    __shared__ float ex_g[MAX_EXTRA][3];
    __shared__ float scratch[32];
    __shared__ float red[8][12];
    __shared__ unsigned ex_mask[EX_MASK_WORDS];
    const int64_t t = blockIdx.y;
    const int j = blockIdx.x;
    const int V = m.num_verts;
    const int E = g_joints ? m.num_extra : 0;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        ex_id[e] = m.extra_vertex_ids[e];
        for (int c = 0; c < 3; ++c) ex_g[e][c] = g_joints[(t * joints_stride + NJ + e) * 3 + c];
    }
    __syncthreads();
    ex_mask_build(ex_mask, ex_id, E);
    float acc[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[k] = 0.f;
    if (j < NJ) {
        const int beg = m.csr_ptr[j], end = m.csr_ptr[j + 1];
        for (int e = beg + threadIdx.x; e < end; e += blockDim.x) {
            const int v = m.csr_vert[e];
            const float w = m.csr_weight[e];
            float g[3];
            eff_grad(g_verts, t, V, v, E, ex_id, ex_g, ex_mask, g);
            const float *vp = vposed + t * ldp + 3 * v;
            const float x = vp[0], y = vp[1], z = vp[2];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float wg = w * g[r];
                acc[4 * r] = fmaf(wg, x, acc[4 * r]);
                acc[4 * r + 1] = fmaf(wg, y, acc[4 * r + 1]);
                acc[4 * r + 2] = fmaf(wg, z, acc[4 * r + 2]);
                acc[4 * r + 3] += wg;
            }
        }
        // all twelve sums in ONE block reduction (fixed shuffle tree per warp, warps added in order: deterministic)
        // instead of twelve with two barriers each
#pragma unroll
        for (int k = 0; k < 12; ++k) acc[k] = warp_sum(acc[k]);
        if ((threadIdx.x & 31) == 0) {
#pragma unroll
            for (int k = 0; k < 12; ++k) red[threadIdx.x >> 5][k] = acc[k];
        }
        __syncthreads();
        if (threadIdx.x < 12) {
            float s = 0.f;
            for (int w = 0; w < int(blockDim.x >> 5); ++w) s += red[w][threadIdx.x];
            gA[(t * NJ + j) * 12 + threadIdx.x] = s;
        }
    } else {
        for (int v = threadIdx.x; v < V; v += blockDim.x) {
            float g[3];
            eff_grad(g_verts, t, V, v, E, ex_id, ex_g, ex_mask, g);
            acc[0] += g[0];
            acc[1] += g[1];
            acc[2] += g[2];
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float s = block_sum(acc[k], scratch);
            if (threadIdx.x == 0) gT[t * 3 + k] = s;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// 4. per-frame pose backward
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void rodrigues_bwd(const float *r, const float *gR, float *gr) {
    const float eps = 1e-8f;
    const float ax = r[0] + eps, ay = r[1] + eps, az = r[2] + eps;
    const float a = sqrtf(ax * ax + ay * ay + az * az);
    const float x = r[0] / a, y = r[1] / a, z = r[2] / a;
    const float s = sinf(a), c = cosf(a), omc = 1.0f - c;
    const float K[9] = {0.f, -z, y, z, 0.f, -x, -y, x, 0.f};
    const float K2[9] = {-(y * y + z * z), x * y, x * z, x * y, -(x * x + z * z), y * z, x * z, y * z, -(x * x + y * y)};
    float g_s = 0.f, g_omc = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        g_s += gR[k] * K[k];
        g_omc += gR[k] * K2[k];
    }
    // gK = s gR + omc (gR K^T + K^T gR)
    float gK[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float u = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) u += gR[3 * i + k] * K[3 * j + k] + K[3 * k + i] * gR[3 * k + j];
            gK[3 * i + j] = s * gR[3 * i + j] + omc * u;
        }
    const float gdx = gK[7] - gK[5], gdy = gK[2] - gK[6], gdz = gK[3] - gK[1];
    float g_a = g_s * c + g_omc * s;
    g_a -= (gdx * r[0] + gdy * r[1] + gdz * r[2]) / (a * a);
    gr[0] = gdx / a + g_a * ax / a;
    gr[1] = gdy / a + g_a * ay / a;
    gr[2] = gdz / a + g_a * az / a;
}

__global__ void __launch_bounds__(64) pose_bwd_kernel(const fpv_smplx_model_t m, const float *__restrict__ theta,
                                                      const float *__restrict__ sR, const float *__restrict__ sG,
                                                      const float *__restrict__ sJ, const float *__restrict__ gA,
                                                      const float *__restrict__ gT, const float *__restrict__ gC,
                                                      const float *__restrict__ g_joints, int joints_stride,
                                                      float *__restrict__ g_theta) {
    __shared__ float th[NTH];
    __shared__ float fp[165];
    __shared__ float gfp[165];
    __shared__ float R[NJ][9];
    __shared__ float GR[NJ][9];
    __shared__ float J[NJ][3];
    __shared__ float gGR[NJ][9];
    __shared__ float gGt[NJ][3];
    __shared__ float gJ[NJ][3];
    __shared__ float gR[NJ][9];
    __shared__ int par[NJ];
    const int64_t t = blockIdx.x;
    const int tid = threadIdx.x;
    for (int k = tid; k < NTH; k += 64) th[k] = theta[t * NTH + k];
    if (tid < NJ) par[tid] = m.parents[tid];
    __syncthreads();
    for (int k = tid; k < 165; k += 64) {  // recompute full_pose (needed by the Rodrigues backward)
        float v;
        if (k < 75) {
            v = th[k];
        } else {
            const int hand = (k >= 120);
            const int c = k - (hand ? 120 : 75);
            const float *comp = m.hand_comps + hand * 12 * 45;
            const float *pc = th + (hand ? TH_RH : TH_LH);
            v = 0.f;
            for (int i = 0; i < 12; ++i) v = fmaf(pc[i], comp[i * 45 + c], v);
        }
        fp[k] = v + m.pose_mean[k];
    }
    if (tid < NJ) {
        const int j = tid;
        for (int k = 0; k < 9; ++k) R[j][k] = sR[(t * NJ + j) * 9 + k];
        for (int c = 0; c < 3; ++c) J[j][c] = sJ[(t * NJ + j) * 3 + c];
        const float *g = sG + (t * NJ + j) * 12;
        const float *ga = gA + (t * NJ + j) * 12;
        float gat[3];
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) GR[j][3 * r + c] = g[4 * r + c];
            gat[r] = ga[4 * r + 3];
        }
        // A = [G_R | G_t - G_R J]
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) gGR[j][3 * r + c] = ga[4 * r + c] - gat[r] * J[j][c];
            gGt[j][r] = gat[r] + (g_joints ? g_joints[(t * joints_stride + j) * 3 + r] : 0.f);
        }
        for (int c = 0; c < 3; ++c)
            gJ[j][c] = -(GR[j][c] * gat[0] + GR[j][3 + c] * gat[1] + GR[j][6 + c] * gat[2]);
    }
    __syncthreads();
    if (tid == 0) {  // reverse kinematic chain (children have larger indices than parents)
        for (int j = NJ - 1; j >= 1; --j) {
            const int p = par[j];
            const float rel[3] = {J[j][0] - J[p][0], J[j][1] - J[p][1], J[j][2] - J[p][2]};
            float grel[3];
            for (int r = 0; r < 3; ++r) {
                for (int c = 0; c < 3; ++c) {
                    // gR_j = G_R^p^T gGR_j
                    gR[j][3 * r + c] = GR[p][r] * gGR[j][c] + GR[p][3 + r] * gGR[j][3 + c] + GR[p][6 + r] * gGR[j][6 + c];
                }
                grel[r] = GR[p][r] * gGt[j][0] + GR[p][3 + r] * gGt[j][1] + GR[p][6 + r] * gGt[j][2];
            }
            for (int r = 0; r < 3; ++r) {
                for (int c = 0; c < 3; ++c) {
                    // gGR_p += gGR_j R_j^T + gGt_j (x) rel
                    gGR[p][3 * r + c] += gGR[j][3 * r] * R[j][3 * c] + gGR[j][3 * r + 1] * R[j][3 * c + 1] +
                                         gGR[j][3 * r + 2] * R[j][3 * c + 2] + gGt[j][r] * rel[c];
                }
                gGt[p][r] += gGt[j][r];
                gJ[j][r] += grel[r];
                gJ[p][r] -= grel[r];
            }
        }
        for (int k = 0; k < 9; ++k) gR[0][k] = gGR[0][k];
        for (int r = 0; r < 3; ++r) gJ[0][r] += gGt[0][r];
    }
    __syncthreads();
    if (tid < NJ) {
        const int j = tid;
        float g[9];
        for (int k = 0; k < 9; ++k) g[k] = gR[j][k] + (j >= 1 ? gC[t * KP + 9 * (j - 1) + k] : 0.f);
        rodrigues_bwd(&fp[3 * j], g, &gfp[3 * j]);
    }
    __syncthreads();
    float *out = g_theta + t * NTH;
    for (int k = tid; k < NTH; k += 64) {
        float v;
        if (k < 75) {
            v = gfp[k];
        } else if (k < TH_BETA) {
            const int hand = (k >= TH_RH);
            const int i = k - (hand ? TH_RH : TH_LH);
            const float *comp = m.hand_comps + hand * 12 * 45 + i * 45;
            const float *gs = gfp + (hand ? 120 : 75);
            v = 0.f;
            for (int c = 0; c < 45; ++c) v = fmaf(comp[c], gs[c], v);
        } else if (k < TH_TRANSL) {
            const int l = k - TH_BETA;
            v = gC[t * KP + NPF + l];
            for (int j = 0; j < NJ; ++j)
                for (int c = 0; c < 3; ++c) v = fmaf(m.j_shapedirs[(j * 3 + c) * NSH + l], gJ[j][c], v);
        } else {
            const int c = k - TH_TRANSL;
            v = gT[t * 3 + c];
            if (g_joints) {  // the 55 chain joints; the extras' share already arrived through gT
                for (int j = 0; j < NJ; ++j) v += g_joints[(t * joints_stride + j) * 3 + c];
            }
        }
        out[k] = v;
    }
}

constexpr int BWD_KSPLIT = 12;  // 3 x 4 output tiles x 12 k-slices = 144 CTAs ~ one wave of 148 SMs

struct BwdLayout {
    size_t gvp_hi, gvp_lo, gA, gT, gC, part, total;  // float offsets
    int nz;
};
static BwdLayout bwd_layout(int64_t T, int64_t V) {
    BwdLayout L;
    size_t o = 0;
    auto take = [&](size_t n) {
        size_t r = o;
        o += align_up(n, 64);
        return r;
    };
    L.nz = BWD_KSPLIT;
    L.gvp_hi = take(size_t(T) * pitch3v(V));
    L.gvp_lo = take(size_t(T) * pitch3v(V));
    L.gA = take(size_t(T) * NJ * 12);
    L.gT = take(size_t(T) * 3);
    L.gC = take(size_t(T) * KP);
    L.part = take(size_t(L.nz) * T * KP);
    L.total = o;
    return L;
}

}  // namespace fpv

using namespace fpv;

extern "C" {

size_t fpv_smplx_saved_bytes(const fpv_smplx_model_t *model, int64_t T) {
    if (!model || T <= 0) return 0;
    return saved_layout(T, model->num_verts).total * sizeof(float) + 256;
}

size_t fpv_smplx_workspace_bytes(const fpv_smplx_model_t *model, int64_t T) {
    if (!model || T <= 0) return 0;
    return bwd_layout(T, model->num_verts).total * sizeof(float) + 256;
}

static int check_model(const fpv_smplx_model_t *m) {
    FPV_CHECK_ARG(m, "smplx: null model");
    FPV_CHECK_ARG(m->num_verts > 0 && m->ell_width > 0 && m->num_extra >= 0 && m->num_extra <= MAX_EXTRA,
                  "smplx: bad model sizes (V=%d W=%d E=%d)", m->num_verts, m->ell_width, m->num_extra);
    FPV_CHECK_ARG(m->basis_nk_hi && m->basis_nk_lo && m->basis_kn_hi && m->basis_kn_lo,
                  "smplx: model lacks the TF32-split basis (tensor-core operands)");
    FPV_CHECK_ARG(m->j_template && m->j_shapedirs && m->parents && m->hand_comps && m->pose_mean &&
                      m->ell_joint && m->ell_weight && m->csr_ptr && m->csr_vert && m->csr_weight,
                  "smplx: model has null constant pointers");
    FPV_CHECK_ARG(m->num_extra == 0 || m->extra_vertex_ids, "smplx: extra_vertex_ids missing");
    return FPV_OK;
}

int fpv_smplx_fwd(const fpv_smplx_model_t *model, int64_t T, const float *theta, float *vertices, float *joints,
                  void *saved, void *workspace, size_t workspace_bytes, fpv_stream_t stream) {
    (void)workspace;
    (void)workspace_bytes;
    int rc = check_model(model);
    if (rc) return rc;
    FPV_CHECK_ARG(theta && vertices && joints && saved, "fpv_smplx_fwd: null pointer");
    FPV_CHECK_ARG(T > 0 && T <= 65535, "fpv_smplx_fwd: T=%lld out of range", (long long)T);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const fpv_smplx_model_t m = *model;
    const int V = m.num_verts;
    const SavedLayout L = saved_layout(T, V);
    float *S = static_cast<float *>(saved);
    const int jstride = NJ + m.num_extra;
    const int64_t ldp = pitch3v(V);
    pose_fwd_kernel<<<(unsigned)T, 64, 0, st>>>(m, theta, S + L.R, S + L.G, S + L.J, S + L.A, S + L.coef_hi, S + L.coef_lo,
                                                joints, jstride);
    FPV_LAUNCH_CHECK("pose_fwd_kernel");
    // v_posed[T,3V] = coef[T,512] x basis^T  (template + shape + pose blend shapes, one tcgen05 contraction)
    rc = tc_gemm_3xtf32(S + L.coef_hi, S + L.coef_lo, KP, m.basis_nk_hi, m.basis_nk_lo, KP, int(T), 3 * V, KP,
                        S + L.vposed, ldp, 1, nullptr, st);
    if (rc) return rc;
    {
        dim3 grid((unsigned)ceil_div(V, SKIN_THREADS), (unsigned)T);
        skin_fwd_kernel<<<grid, SKIN_THREADS, 0, st>>>(m, theta, S + L.A, S + L.vposed, ldp, vertices);
        FPV_LAUNCH_CHECK("skin_fwd_kernel");
    }
    if (m.num_extra > 0) {
        extras_fwd_kernel<<<(unsigned)T, MAX_EXTRA, 0, st>>>(m, vertices, joints, jstride);
        FPV_LAUNCH_CHECK("extras_fwd_kernel");
    }
    return FPV_OK;
}

int fpv_smplx_bwd(const fpv_smplx_model_t *model, int64_t T, const float *theta, const void *saved,
                  const float *g_vertices, const float *g_joints, float *g_theta, void *workspace,
                  size_t workspace_bytes, fpv_stream_t stream) {
    int rc = check_model(model);
    if (rc) return rc;
    FPV_CHECK_ARG(theta && saved && g_theta, "fpv_smplx_bwd: null pointer");
    FPV_CHECK_ARG(T > 0 && T <= 65535, "fpv_smplx_bwd: T=%lld out of range", (long long)T);
    FPV_CHECK_ARG(workspace && workspace_bytes >= fpv_smplx_workspace_bytes(model, T),
                  "fpv_smplx_bwd: workspace too small (%zu < %zu)", workspace_bytes,
                  fpv_smplx_workspace_bytes(model, T));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const fpv_smplx_model_t m = *model;
    const int V = m.num_verts;
    const SavedLayout L = saved_layout(T, V);
    const BwdLayout B = bwd_layout(T, V);
    const float *S = static_cast<const float *>(saved);
    float *W = static_cast<float *>(workspace);
    const int jstride = NJ + m.num_extra;
    const int64_t ldp = pitch3v(V);
    {
        dim3 grid((unsigned)ceil_div(V, SKIN_THREADS), (unsigned)T);
        skin_bwd_kernel<<<grid, SKIN_THREADS, 0, st>>>(m, S + L.A, g_vertices, g_joints, jstride, ldp, W + B.gvp_hi,
                                                       W + B.gvp_lo);
        FPV_LAUNCH_CHECK("skin_bwd_kernel");
    }
    {
        dim3 grid(NJ + 1, (unsigned)T);
        jointgrad_kernel<<<grid, 128, 0, st>>>(m, S + L.vposed, ldp, g_vertices, g_joints, jstride, W + B.gA, W + B.gT);
        FPV_LAUNCH_CHECK("jointgrad_kernel");
    }
    // gC[T,512] = g_vposed[T,3V] x basis  (split-K over 3V on the tensor cores, fixed-order slice reduction)
    rc = tc_gemm_3xtf32(W + B.gvp_hi, W + B.gvp_lo, ldp, m.basis_kn_hi, m.basis_kn_lo, ldp, int(T), KP, 3 * V, W + B.gC,
                        KP, B.nz, W + B.part, st);
    if (rc) return rc;
    pose_bwd_kernel<<<(unsigned)T, 64, 0, st>>>(m, theta, S + L.R, S + L.G, S + L.J, W + B.gA, W + B.gT, W + B.gC,
                                                g_joints, jstride, g_theta);
    FPV_LAUNCH_CHECK("pose_bwd_kernel");
    return FPV_OK;
}

}  // extern "C"
