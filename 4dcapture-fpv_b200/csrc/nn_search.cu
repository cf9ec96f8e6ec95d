// nn_search.cu -- exact brute-force nearest neighbour (fused min + argmin) and the chamfer
// forward/backward built on it.  sm_100a only.
//
// Replaces: chamfer_python.py:18-28 (distChamfer: three bmm + four reductions over a dense
// [bs,N,M] matrix) and the [3P] NmDistanceKernel behind ext.chamferDist()
// (global_optimization.py:292-294).  Design:
//   * candidates live in padded SoA planes; a producer warp streams 1024-point tiles of the three
//     planes into a 4-stage shared-memory ring with cp.async.bulk (TMA engine) + mbarriers;
//   * 8 consumer warps hold QPT queries per thread in registers and evaluate four candidates per
//     step from three broadcast LDS.128, with packed FADD2/FMUL2/FFMA2 (two candidates per
//     instruction) -- the canonical d = fma(dz,dz,fma(dy,dy,dx*dx)) bit for bit;
//   * the running minimum is a FMNMX3 chain; the argmin is only resolved on the (rare) steps in
//     which the minimum strictly improved, so ties keep the lowest index;
//   * candidate ranges can be split across CTAs; partial winners merge through a 64-bit
//     (distance bits << 32 | index) atomicMin -- the same key the multi-GPU combine uses.
#include <math_constants.h>

#include "common.cuh"

namespace fpv {

constexpr int NN_THREADS = 256;  // consumer threads (8 warps) + 1 producer warp
constexpr int NN_TILE = 1024;    // candidates per pipeline stage
constexpr int NN_STAGES = 4;
constexpr int NN_PAD = 64;       // plane length granularity (points; one culling tile of nn_culled.cu)
constexpr size_t NN_SMEM = size_t(NN_STAGES) * 3 * NN_TILE * sizeof(float) + 2 * NN_STAGES * sizeof(uint64_t);

static int g_tune_qpt = 0;      // 0 = heuristic
static int g_tune_nsplit = 0;   // 0 = heuristic
static int g_tune_packed = -1;  // -1 = default; else number of packed queries per thread
static int g_engine = 0;        // 0 = auto, 1 = FP32 SIMT brute force, 2 = tensor-core filter + exact re-check

// nn_tc.cu
size_t nn_tc_workspace_bytes(int64_t cand_batches);
void nn_tc_set_eshift(int e);
void nn_tc_set_subtile(int ns);
void nn_tc_set_debug(long long *dbg);
void nn_tc_plan(int64_t eb, int64_t eN, int64_t M, int nsplit_hint, int *nsplit, int64_t *chunk);
int nn_tc_launch(const float *queries, int64_t q_bstride, int64_t eb, int64_t eN, const float *planes,
                 int64_t plane_bstride, int64_t cand_batches, int64_t Mp, int64_t M, int64_t idx_base, float *dist,
                 void *idx, int idx_bytes, unsigned long long *keys, int keys_atomic, int nsplit, int64_t chunk,
                 float *ymax_ws, cudaStream_t st);

static bool use_tc(int64_t eb, int64_t eN, int64_t M) {
    if (g_engine == 1) return false;
    if (g_engine == 2) return true;
    // the filter pays once there are enough rows to fill the MMA tiles and enough candidates to amortise setup
    return eN >= 2048 && M >= 1024 && double(eb) * double(eN) * double(M) >= 2.0e9;
}

struct NNParams {
    const float *q;
    int64_t q_bstride;  // floats between query batches (0 = shared)
    int64_t N;          // queries per batch
    const float *planes;
    int64_t plane_bstride;  // floats between candidate batches (0 = shared)
    int64_t Mp;             // plane length
    int64_t M8;             // candidates scanned (M rounded up to 8; the pad is +inf)
    int64_t chunk;          // candidates per blockIdx.y slice (multiple of NN_TILE)
    int64_t idx_base;
    float *dist;
    void *idx;
    int idx_bytes;
    unsigned long long *keys;
    int keys_atomic;
};

__device__ __forceinline__ float d2_scalar(float x, float y, float z, float rx, float ry, float rz) {
    const float dx = __fsub_rn(x, rx), dy = __fsub_rn(y, ry), dz = __fsub_rn(z, rz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

// Four candidates (one float4 per plane) against QPT register-resident queries.
// The first NP queries use packed FADD2/FMUL2/FFMA2 (3 issue slots per pair, but the packed forms only
// sustain ~103 of the 128 lane-ops/SM/clk, profiles/r01_fp32_pipe_probe.txt); the rest use scalar
// FADD/FMUL/FFMA (full pipe rate, 6 issue slots per pair).  Mixing the two balances the issue port
// against the FP32 pipe.  Both evaluate the canonical expression, bit for bit.
template <int QPT, int NP>
__device__ __forceinline__ void nn_step(const float4 rx, const float4 ry, const float4 rz, const int j,
                                        const float (&qx)[QPT], const float (&qy)[QPT],
                                        const float (&qz)[QPT], float (&best)[QPT], int (&bidx)[QPT]) {
    float mnew[QPT];
    bool any = false;
    const float2 nx0 = make_float2(-rx.x, -rx.y), nx1 = make_float2(-rx.z, -rx.w);
    const float2 ny0 = make_float2(-ry.x, -ry.y), ny1 = make_float2(-ry.z, -ry.w);
    const float2 nz0 = make_float2(-rz.x, -rz.y), nz1 = make_float2(-rz.z, -rz.w);
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
        float m;
        if (q < NP) {
            const float2 bx = make_float2(qx[q], qx[q]), by = make_float2(qy[q], qy[q]), bz = make_float2(qz[q], qz[q]);
            const float2 dx0 = __fadd2_rn(bx, nx0), dx1 = __fadd2_rn(bx, nx1);
            const float2 dy0 = __fadd2_rn(by, ny0), dy1 = __fadd2_rn(by, ny1);
            const float2 dz0 = __fadd2_rn(bz, nz0), dz1 = __fadd2_rn(bz, nz1);
            float2 s0 = __fmul2_rn(dx0, dx0), s1 = __fmul2_rn(dx1, dx1);
            s0 = __ffma2_rn(dy0, dy0, s0);
            s1 = __ffma2_rn(dy1, dy1, s1);
            s0 = __ffma2_rn(dz0, dz0, s0);
            s1 = __ffma2_rn(dz1, dz1, s1);
            m = fmin3(best[q], s0.x, s0.y);
            m = fmin3(m, s1.x, s1.y);
        } else {
            const float e0 = d2_scalar(qx[q], qy[q], qz[q], rx.x, ry.x, rz.x);
            const float e1 = d2_scalar(qx[q], qy[q], qz[q], rx.y, ry.y, rz.y);
            const float e2 = d2_scalar(qx[q], qy[q], qz[q], rx.z, ry.z, rz.z);
            const float e3 = d2_scalar(qx[q], qy[q], qz[q], rx.w, ry.w, rz.w);
            m = fmin3(best[q], e0, e1);
            m = fmin3(m, e2, e3);
        }
        mnew[q] = m;
        any |= (m < best[q]);
    }
    if (any) {  // rare after the first few tiles: resolve WHICH candidate, lowest index first
#pragma unroll
        for (int q = 0; q < QPT; ++q) {
            if (mnew[q] < best[q]) {
                const float x = qx[q], y = qy[q], z = qz[q], m = mnew[q];
                const float e0 = d2_scalar(x, y, z, rx.x, ry.x, rz.x);
                const float e1 = d2_scalar(x, y, z, rx.y, ry.y, rz.y);
                const float e2 = d2_scalar(x, y, z, rx.z, ry.z, rz.z);
                bidx[q] = j + ((e0 == m) ? 0 : (e1 == m) ? 1 : (e2 == m) ? 2 : 3);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < QPT; ++q) best[q] = mnew[q];
}

template <int QPT, int NP>
__global__ void __launch_bounds__(NN_THREADS + 32, (QPT <= 4) ? 2 : 1) nn_search_kernel(const NNParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *tiles = reinterpret_cast<float *>(smem_raw);  // [STAGES][3][TILE]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + size_t(NN_STAGES) * 3 * NN_TILE * sizeof(float));
    uint64_t *empty = full + NN_STAGES;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t b = blockIdx.z;
    const int64_t r0 = int64_t(blockIdx.y) * p.chunk;
    const int64_t r1 = (r0 + p.chunk < p.M8) ? (r0 + p.chunk) : p.M8;
    const int ntiles = int((r1 - r0 + NN_TILE - 1) / NN_TILE);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NN_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NN_THREADS / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == NN_THREADS / 32) {  // ---- producer warp: one lane drives the TMA engine ----
        if (lane == 0) {
            const float *src = p.planes + b * p.plane_bstride + r0;
            for (int k = 0; k < ntiles; ++k) {
                const int s = k % NN_STAGES;
                const uint32_t ph = (k / NN_STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                const int64_t rem = r1 - r0 - int64_t(k) * NN_TILE;
                const uint32_t bytes = uint32_t(rem < NN_TILE ? rem : NN_TILE) * 4u;
                float *dst = tiles + size_t(s) * 3 * NN_TILE;
                const float *g = src + int64_t(k) * NN_TILE;
                mbar_arrive_expect_tx(&full[s], 3 * bytes);
                bulk_g2s(dst, g, bytes, &full[s]);
                bulk_g2s(dst + NN_TILE, g + p.Mp, bytes, &full[s]);
                bulk_g2s(dst + 2 * NN_TILE, g + 2 * p.Mp, bytes, &full[s]);
            }
        }
        return;
    }

    // ---- consumer warps ----
    float qx[QPT], qy[QPT], qz[QPT];
    float best[QPT];
    int bidx[QPT];
    const int64_t qbase = int64_t(blockIdx.x) * (NN_THREADS * QPT);
    const float *qsrc = p.q + b * p.q_bstride;
#pragma unroll
    for (int k = 0; k < QPT; ++k) {
        int64_t qi = qbase + int64_t(k) * NN_THREADS + tid;
        if (qi > p.N - 1) qi = p.N - 1;
        const float x = __ldg(qsrc + 3 * qi), y = __ldg(qsrc + 3 * qi + 1), z = __ldg(qsrc + 3 * qi + 2);
        qx[k] = x;
        qy[k] = y;
        qz[k] = z;
        best[k] = CUDART_INF_F;
        bidx[k] = 0;
    }

    for (int k = 0; k < ntiles; ++k) {
        const int s = k % NN_STAGES;
        const uint32_t ph = (k / NN_STAGES) & 1;
        mbar_wait(&full[s], ph);
        const float4 *X = reinterpret_cast<const float4 *>(tiles + size_t(s) * 3 * NN_TILE);
        const float4 *Y = X + NN_TILE / 4;
        const float4 *Z = Y + NN_TILE / 4;
        const int64_t rem = r1 - r0 - int64_t(k) * NN_TILE;
        const int cnt4 = int(rem < NN_TILE ? rem : NN_TILE) >> 2;  // even: M8 is a multiple of 8
        const int jbase = int(r0) + k * NN_TILE;
#pragma unroll 2
        for (int j4 = 0; j4 < cnt4; ++j4) {
            nn_step<QPT, NP>(X[j4], Y[j4], Z[j4], jbase + 4 * j4, qx, qy, qz, best, bidx);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }

#pragma unroll
    for (int k = 0; k < QPT; ++k) {
        const int64_t qi = qbase + int64_t(k) * NN_THREADS + tid;
        if (qi < p.N) {
            const int64_t o = b * p.N + qi;
            const int64_t gi = p.idx_base + bidx[k];
            if (p.keys) {
                const unsigned long long key =
                    (static_cast<unsigned long long>(__float_as_uint(best[k])) << 32) |
                    static_cast<unsigned long long>(static_cast<uint32_t>(gi));
                if (p.keys_atomic)
                    atomicMin(p.keys + o, key);
                else
                    p.keys[o] = key;
            } else {
                p.dist[o] = best[k];
                if (p.idx_bytes == 8)
                    static_cast<long long *>(p.idx)[o] = gi;
                else if (p.idx_bytes == 4)
                    static_cast<int *>(p.idx)[o] = int(gi);
            }
        }
    }
}

__global__ void pack_planes_kernel(const float *__restrict__ pts, int64_t M, int64_t Mp, float *__restrict__ planes) {
    const int64_t b = blockIdx.y;
    const int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= Mp) return;
    float x = CUDART_INF_F, y = CUDART_INF_F, z = CUDART_INF_F;
    if (j < M) {
        const float *s = pts + (b * M + j) * 3;
        x = s[0];
        y = s[1];
        z = s[2];
    }
    float *d = planes + b * 3 * Mp;
    d[j] = x;
    d[Mp + j] = y;
    d[2 * Mp + j] = z;
}

__global__ void unpack_keys_kernel(const unsigned long long *__restrict__ keys, int64_t n, float *dist, void *idx,
                                   int idx_bytes) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = keys[i];
    if (dist) dist[i] = __uint_as_float(static_cast<uint32_t>(k >> 32));
    const uint32_t lo = static_cast<uint32_t>(k);
    if (idx_bytes == 8)
        static_cast<long long *>(idx)[i] = static_cast<long long>(lo);
    else if (idx_bytes == 4)
        static_cast<int *>(idx)[i] = static_cast<int>(lo);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int64_t planes_len(int64_t M) { return ceil_div(M, NN_PAD) * NN_PAD; }

struct NNPlan {
    int qpt;
    int64_t qblocks;
    int nsplit;
    int64_t chunk, M8;
};

static NNPlan nn_plan(int64_t batches, int64_t N, int64_t M) {
    NNPlan pl;
    const int sms = sm_count();
    const int64_t total_q = batches * N;
    (void)total_q;
    pl.qpt = 4;  // measured: 4 queries/thread, 2 CTAs/SM beats 8 queries/thread, 1 CTA/SM on every shape (profiles/r01_nn_tune2.txt)
    if (g_tune_qpt == 4 || g_tune_qpt == 8) pl.qpt = g_tune_qpt;
    pl.qblocks = ceil_div(N, int64_t(NN_THREADS) * pl.qpt);
    pl.M8 = ceil_div(M, 8) * 8;
    const int64_t base = pl.qblocks * batches;
    const int64_t target = int64_t(sms) * 16;  // >= 8 waves at 2 CTAs/SM: tail below ~6%
    int64_t ns = ceil_div(target, base);
    const int64_t max_ns = pl.M8 / (4 * NN_TILE) > 1 ? pl.M8 / (4 * NN_TILE) : 1;
    if (ns > max_ns) ns = max_ns;
    if (ns > 65535) ns = 65535;
    if (g_tune_nsplit > 0) ns = g_tune_nsplit;
    pl.chunk = ceil_div(ceil_div(pl.M8, ns), NN_TILE) * NN_TILE;
    pl.nsplit = int(ceil_div(pl.M8, pl.chunk));
    return pl;
}

template <int QPT, int NP>
static cudaError_t nn_launch(const NNParams &p, dim3 grid, cudaStream_t st) {
    static PerDeviceOnce once;
    bool *configured = once.slot();
    if (!*configured) {
        cudaError_t e = cudaFuncSetAttribute(nn_search_kernel<QPT, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             int(NN_SMEM));
        if (e != cudaSuccess) return e;
        *configured = true;
    }
    nn_search_kernel<QPT, NP><<<grid, NN_THREADS + 32, NN_SMEM, st>>>(p);
    return cudaGetLastError();
}

constexpr int NN_DEFAULT_NP4 = 4, NN_DEFAULT_NP8 = 8;

static cudaError_t nn_dispatch(int qpt, const NNParams &p, dim3 grid, cudaStream_t st) {
    const int np = g_tune_packed >= 0 ? g_tune_packed : (qpt == 8 ? NN_DEFAULT_NP8 : NN_DEFAULT_NP4);
    if (qpt == 8) {
        switch (np) {
            case 0: return nn_launch<8, 0>(p, grid, st);
            case 2: return nn_launch<8, 2>(p, grid, st);
            case 4: return nn_launch<8, 4>(p, grid, st);
            case 6: return nn_launch<8, 6>(p, grid, st);
            default: return nn_launch<8, 8>(p, grid, st);
        }
    }
    switch (np) {
        case 0: return nn_launch<4, 0>(p, grid, st);
        case 1: return nn_launch<4, 1>(p, grid, st);
        case 2: return nn_launch<4, 2>(p, grid, st);
        case 3: return nn_launch<4, 3>(p, grid, st);
        default: return nn_launch<4, 4>(p, grid, st);
    }
}

// One direction.  Flattens the batch axis when one candidate set serves every batch.
static int nn_search_impl(const float *queries, int q_shared, int64_t batches, int64_t N, const float *ref_planes,
                          int64_t ref_batches, int64_t M, int64_t idx_base, float *dist, void *idx, int idx_bytes,
                          uint64_t *keys, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    FPV_CHECK_ARG(batches > 0 && N > 0 && M > 0, "nn_search: empty input (batches=%lld N=%lld M=%lld)",
                  (long long)batches, (long long)N, (long long)M);
    FPV_CHECK_ARG(ref_batches == 1 || ref_batches == batches, "nn_search: ref_batches must be 1 or batches");
    FPV_CHECK_ARG(idx_bytes == 0 || idx_bytes == 4 || idx_bytes == 8, "nn_search: idx_bytes must be 0, 4 or 8");
    FPV_CHECK_ARG(M + idx_base <= 0xffffffffll && idx_base >= 0, "nn_search: index range exceeds 32 bits");
    FPV_CHECK_ARG(keys || (dist && (idx || idx_bytes == 0)), "nn_search: no output requested");
    const int64_t Mp = planes_len(M);
    int64_t eb = batches, eN = N;
    int64_t q_bstride = q_shared ? 0 : N * 3;
    int64_t plane_bstride = (ref_batches == 1) ? 0 : 3 * Mp;
    if (ref_batches == 1 && !q_shared) {  // every query meets the same candidates: one flat batch
        eN = batches * N;
        eb = 1;
        q_bstride = 0;
    }
    FPV_CHECK_ARG(eb <= 65535, "nn_search: more than 65535 candidate batches");
    const int64_t total = eb * eN;
    const bool tc = use_tc(eb, eN, M);
    NNPlan pl = nn_plan(eb, eN, M);
    int tc_ns = 1;
    int64_t tc_chunk = 0;
    if (tc) nn_tc_plan(eb, eN, M, g_tune_nsplit, &tc_ns, &tc_chunk);
    const int nsplit = tc ? tc_ns : pl.nsplit;
    Arena ar(workspace, workspace_bytes);
    float *ymax_ws = nullptr;
    if (tc) {
        ymax_ws = ar.take<float>(size_t(ref_batches));
        if (!ymax_ws) {
            set_error("nn_search: workspace too small (%zu bytes)", workspace_bytes);
            return FPV_ERR_WORKSPACE;
        }
    }
    unsigned long long *kout = reinterpret_cast<unsigned long long *>(keys);
    unsigned long long *scratch_keys = nullptr;
    int keys_atomic = 0;
    if (nsplit > 1) {
        if (!kout) {
            scratch_keys = ar.take<unsigned long long>(size_t(total));
            if (!scratch_keys) {
                set_error("nn_search: workspace too small (%zu bytes, need %zu)", workspace_bytes,
                          align_up(size_t(total) * 8, 256) + 512);
                return FPV_ERR_WORKSPACE;
            }
            kout = scratch_keys;
        }
        keys_atomic = 1;
        FPV_CUDA(cudaMemsetAsync(kout, 0xff, size_t(total) * 8, st));
    }
    if (profile_on()) {
        // algorithmic bytes: every distinct query point and candidate once, outputs once
        const double qpts = double(q_shared ? N : batches * N), out_b = keys && !dist ? 8.0 : 4.0 + idx_bytes;
        char nm[48];
        snprintf(nm, sizeof(nm), "%s Q=%lld M=%lld", tc ? "nn_tc" : "nn_search<4>", (long long)total, (long long)M);
        profile_begin(nm, st, 12.0 * qpts + 12.0 * double(M) * double(ref_batches) + out_b * double(total),
                      double(total) * double(M));
    }
    if (tc) {
        int rc = nn_tc_launch(queries, q_bstride, eb, eN, ref_planes, plane_bstride, ref_batches, Mp, M, idx_base, dist,
                              idx, idx_bytes, kout, keys_atomic, tc_ns, tc_chunk, ymax_ws, st);
        profile_end(st);
        if (rc) return rc;
    } else {
        NNParams p;
        p.q = queries;
        p.q_bstride = q_bstride;
        p.N = eN;
        p.planes = ref_planes;
        p.plane_bstride = plane_bstride;
        p.Mp = Mp;
        p.M8 = pl.M8;
        p.chunk = pl.chunk;
        p.idx_base = idx_base;
        p.dist = dist;
        p.idx = idx;
        p.idx_bytes = idx_bytes;
        p.keys = kout;
        p.keys_atomic = keys_atomic;
        dim3 grid((unsigned)pl.qblocks, (unsigned)pl.nsplit, (unsigned)eb);
        cudaError_t e = nn_dispatch(pl.qpt, p, grid, st);
        profile_end(st);
        count_launch();
        if (e != cudaSuccess) {
            set_error("nn_search_kernel launch failed: %s", cudaGetErrorString(e));
            return FPV_ERR_CUDA;
        }
    }
    if (scratch_keys || (keys && dist)) {  // split merge, or caller wants keys AND plain outputs
        unpack_keys_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(kout, total, dist, idx, idx_bytes);
        FPV_LAUNCH_CHECK("unpack_keys_kernel");
    }
    return FPV_OK;
}

static size_t nn_search_ws(int64_t batches, int64_t N, int64_t M, int ref_shared) {
    int64_t eb = batches, eN = N;
    if (ref_shared) {
        eN = batches * N;
        eb = 1;
    }
    // worst case over both engines (the engine can be switched by the tuning hook after sizing)
    const NNPlan pl = nn_plan(eb, eN, M);
    int tc_ns = 1;
    int64_t tc_chunk = 0;
    nn_tc_plan(eb, eN, M, g_tune_nsplit, &tc_ns, &tc_chunk);
    const size_t keys = (pl.nsplit > 1 || tc_ns > 1) ? align_up(size_t(eb * eN) * 8, 256) : 0;
    return keys + nn_tc_workspace_bytes(batches) + 256;
}

static int pack_planes_impl(const float *pts, int64_t batches, int64_t M, float *planes, cudaStream_t st) {
    FPV_CHECK_ARG(batches > 0 && M > 0, "pack_planes: empty input");
    FPV_CHECK_ARG(batches <= 65535, "pack_planes: more than 65535 batches");
    const int64_t Mp = planes_len(M);
    dim3 grid((unsigned)ceil_div(Mp, 256), (unsigned)batches);
    pack_planes_kernel<<<grid, 256, 0, st>>>(pts, M, Mp, planes);
    FPV_LAUNCH_CHECK("pack_planes_kernel");
    return FPV_OK;
}

// ---------------------------------------------------------------------------------------------
// chamfer backward: gather for the direct term, fixed-point integer scatter for the argmin term
// ---------------------------------------------------------------------------------------------
// max |p[i]| over a flat array (NaNs ignored), as float bits through atomicMax: order-independent, deterministic
__global__ void absmax_kernel(const float *__restrict__ p, int64_t n, unsigned *out) {
    float m = 0.f;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
        m = fmaxf(m, fabsf(p[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

// power-of-two fixed-point exponent: |c| <= cmax < 2^e, at most 2^nb addends => sum < 2^(e+nb+k) <= 2^62
// bound[0..2] = max|g|, max|x coordinate|, max|y coordinate|  =>  |2 g (x - y)| <= 2 gmax (xmax + ymax)
__device__ __forceinline__ int fix_exponent(const unsigned *bound, int nb_bits) {
    const float cm = 2.000001f * __uint_as_float(bound[0]) * (__uint_as_float(bound[1]) + __uint_as_float(bound[2]));
    if (!(cm > 0.f)) return 0;
    int e;
    frexpf(cm, &e);
    int k = 62 - nb_bits - e;
    return k < -120 ? -120 : (k > 120 ? 120 : k);
}

// For every (b,i): j = idx[b,i]; c = 2 g (x_i - y_j).  to_y: acc[b*acc_bstride + 3j..] -= c ; else
// acc[b*acc_bstride + 3i..] += c (used when the target cloud is shared across batches).
// A warp walks rows of 32 consecutive queries (coalesced loads) and merges the lanes that hit the same target before
// touching memory.  Integer adds are associative, so the merge order cannot change the result.
constexpr int BWD_ROWS = 8;  // rows of 32 consecutive elements per warp

template <typename IdxT>
__global__ void __launch_bounds__(256) bwd_accum_kernel(const float *__restrict__ x, int64_t x_bstride, int64_t N,
                                                        const float *__restrict__ y, int64_t y_bstride,
                                                        const float *__restrict__ g, const IdxT *__restrict__ idx,
                                                        const unsigned *__restrict__ cmax_bits, int nb_bits, int to_y,
                                                        long long *acc, int64_t acc_bstride, int64_t gmul) {
    const int64_t b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int64_t row0 = (int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (32 * BWD_ROWS);
    if (row0 >= N) return;
    const float scale = ldexpf(1.f, fix_exponent(cmax_bits, nb_bits));
    unsigned long long *base = reinterpret_cast<unsigned long long *>(acc + b * acc_bstride);
    const float *xb = x + b * x_bstride, *yb = y + b * y_bstride;
#pragma unroll 2
    for (int r = 0; r < BWD_ROWS; ++r) {
        const int64_t i = row0 + r * 32 + lane;  // lane-consecutive: index, weight and point loads are coalesced
        if (row0 + r * 32 >= N) break;
        const bool valid = i < N;
        long long v[3] = {0, 0, 0};
        long long t = -1 - lane;  // a lane past the end is its own empty run
        if (valid) {
            const int64_t j = static_cast<int64_t>(idx[b * N + i]);
            t = to_y ? j : i;
            const float *xi = xb + 3 * i, *yj = yb + 3 * j;
            const float g2 = __fmul_rn(2.f, g[(b * N + i) * gmul]);  // gmul = 0: one broadcast weight
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float c = __fmul_rn(g2, __fsub_rn(xi[k], yj[k]));
                if (to_y) c = -c;
                v[k] = __float2ll_rn(__fmul_rn(c, scale));
            }
        }
        // lanes that hit the same target, consecutive or not, form one group (match.any); every group is summed through
        // the integer warp-reduce unit -- three 21-bit limbs per component, v = lo + mid 2^21 + hi 2^42 with hi signed,
        // exact for <= 32 addends -- and one lane per group touches memory.  When the queries arrive in spatial order
        // (the sorted scene of the spatial path) neighbours share their nearest body vertex and most atomics disappear.
        const unsigned grp = __match_any_sync(0xffffffffu, t);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int lo = int(v[k] & 0x1FFFFF), mid = int((v[k] >> 21) & 0x1FFFFF), hi = int(v[k] >> 42);
            const long long slo = __reduce_add_sync(grp, lo), smid = __reduce_add_sync(grp, mid);
            const long long shi = __reduce_add_sync(grp, hi);
            v[k] = slo + (smid << 21) + (shi << 42);
        }
        if (t >= 0 && lane == __ffs(grp) - 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (v[k] != 0) atomicAdd(base + 3 * t + k, static_cast<unsigned long long>(v[k]));
        }
    }
}

// grad[b,i,:] = (direct ? 2 g (x_i - y_idx) : 0) + acc * 2^-k
template <typename IdxT>
__global__ void bwd_finish_kernel(const float *__restrict__ x, int64_t N, const float *__restrict__ y,
                                  int64_t y_bstride, const float *__restrict__ g, const IdxT *__restrict__ idx,
                                  const unsigned *__restrict__ cmax_bits, int nb_bits, const long long *__restrict__ acc,
                                  float *__restrict__ grad, int64_t gmul) {
    const int64_t b = blockIdx.y;
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float d[3] = {0.f, 0.f, 0.f};
    if (g) {
        const int64_t j = static_cast<int64_t>(idx[b * N + i]);
        const float *xi = x + (b * N + i) * 3, *yj = y + b * y_bstride + 3 * j;
        const float g2 = __fmul_rn(2.f, g[(b * N + i) * gmul]);
#pragma unroll
        for (int k = 0; k < 3; ++k) d[k] = __fmul_rn(g2, __fsub_rn(xi[k], yj[k]));
    }
    if (acc) {
        const double inv = ldexp(1.0, -fix_exponent(cmax_bits, nb_bits));
#pragma unroll
        for (int k = 0; k < 3; ++k) d[k] = __fadd_rn(d[k], float(double(acc[(b * N + i) * 3 + k]) * inv));
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) grad[(b * N + i) * 3 + k] = d[k];
}

static int ilog2_ceil(int64_t n) {
    int b = 0;
    while ((int64_t(1) << b) < n) ++b;
    return b;
}

template <typename IdxT>
static int chamfer_bwd_impl(const float *a, const float *b, int64_t bs, int64_t N, int64_t M, int b_shared,
                            const float *g_b2a, const float *g_a2b, const IdxT *i_b2a, const IdxT *i_a2b,
                            float *grad_a, float *grad_b, void *workspace, size_t workspace_bytes, cudaStream_t st,
                            int g_broadcast) {
    // g_broadcast bit 0 / 1: g_b2a / g_a2b is ONE weight shared by every element (the gradient of a plain sum or
    // mean arrives as a stride-0 expanded scalar): nothing of size [bs,M] is materialised or re-read
    const int64_t gm1 = (g_broadcast & 1) ? 0 : 1, gm2 = (g_broadcast & 2) ? 0 : 1;
    Arena ar(workspace, workspace_bytes);
    unsigned *cmax = ar.take<unsigned>(8);  // [0..2] bound triple for grad_a, [4..6] for grad_b
    long long *acc_a = g_b2a ? ar.take<long long>(size_t(bs * N * 3)) : nullptr;
    const int64_t gb_batches = b_shared ? 1 : bs;
    const bool need_acc_b = grad_b && (g_a2b || (b_shared && g_b2a));
    long long *acc_b = need_acc_b ? ar.take<long long>(size_t(gb_batches * M * 3)) : nullptr;
    if (!cmax || (g_b2a && !acc_a) || (need_acc_b && !acc_b)) {
        set_error("chamfer_bwd: workspace too small (%zu bytes)", workspace_bytes);
        return FPV_ERR_WORKSPACE;
    }
    const int64_t b_bstride = b_shared ? 0 : M * 3;
    FPV_CUDA(cudaMemsetAsync(cmax, 0, 8 * sizeof(unsigned), st));
    auto absmax = [&](const float *ptr, int64_t n, unsigned *out) {
        int nb = int(ceil_div(n, 256 * 16));
        nb = nb < 1 ? 1 : (nb > 148 * 16 ? 148 * 16 : nb);
        absmax_kernel<<<nb, 256, 0, st>>>(ptr, n, out);
    };
    // the contribution bound only needs max|g| and the coordinate ranges: three cheap streaming reductions instead of a
    // pass that re-gathers every (query, winner) pair.  Nothing to bound when no scatter runs (pure gather).
    if (g_b2a || need_acc_b) {
        absmax(a, bs * N * 3, cmax + 1);
        absmax(b, gb_batches * M * 3, cmax + 2);
        FPV_CUDA(cudaMemcpyAsync(cmax + 5, cmax + 2, sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
        FPV_CUDA(cudaMemcpyAsync(cmax + 6, cmax + 1, sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
        count_launch();
        count_launch();
    }
    dim3 gridN((unsigned)ceil_div(N, 256), (unsigned)bs), gridM((unsigned)ceil_div(M, 256), (unsigned)bs);
    dim3 gridNr((unsigned)ceil_div(N, 256 * BWD_ROWS), (unsigned)bs);
    dim3 gridMr((unsigned)ceil_div(M, 256 * BWD_ROWS), (unsigned)bs);

    // ---- grad_a = 2 g_a2b (a_i - b_idx)  +  sum_{j: i_b2a[j]==i} 2 g_b2a[j] (a_i - b_j)
    const int nb_a = ilog2_ceil(M) + 1;
    if (g_b2a) {
        FPV_CUDA(cudaMemsetAsync(acc_a, 0, size_t(bs * N * 3) * sizeof(long long), st));
        absmax(g_b2a, gm1 ? bs * M : 1, cmax);
        FPV_LAUNCH_CHECK("absmax_kernel");
        bwd_accum_kernel<IdxT><<<gridMr, 256, 0, st>>>(b, b_bstride, M, a, N * 3, g_b2a, i_b2a, cmax, nb_a, 1, acc_a,
                                                      N * 3, gm1);
        FPV_LAUNCH_CHECK("bwd_accum_kernel");
    }
    bwd_finish_kernel<IdxT><<<gridN, 256, 0, st>>>(a, N, b, b_bstride, g_a2b, i_a2b, cmax, nb_a, acc_a, grad_a, gm2);
    FPV_LAUNCH_CHECK("bwd_finish_kernel");

    // ---- grad_b = 2 g_b2a (b_j - a_idx)  +  sum_{i: i_a2b[i]==j} 2 g_a2b[i] (b_j - a_i)
    if (grad_b) {
        const int nb_b = ilog2_ceil(b_shared ? bs * (N + 1) : N) + 1;
        if (acc_b) FPV_CUDA(cudaMemsetAsync(acc_b, 0, size_t(gb_batches * M * 3) * sizeof(long long), st));
        if (g_a2b) {
            absmax(g_a2b, gm2 ? bs * N : 1, cmax + 4);
            FPV_LAUNCH_CHECK("absmax_kernel");
        }
        if (b_shared && g_b2a) {
            absmax(g_b2a, gm1 ? bs * M : 1, cmax + 4);
            FPV_LAUNCH_CHECK("absmax_kernel");
        }
        if (g_a2b) {
            bwd_accum_kernel<IdxT><<<gridNr, 256, 0, st>>>(a, N * 3, N, b, b_bstride, g_a2b, i_a2b, cmax + 4, nb_b, 1,
                                                          acc_b, b_shared ? 0 : M * 3, gm2);
            FPV_LAUNCH_CHECK("bwd_accum_kernel");
        }
        if (b_shared) {
            if (g_b2a) {
                bwd_accum_kernel<IdxT><<<gridMr, 256, 0, st>>>(b, 0, M, a, N * 3, g_b2a, i_b2a, cmax + 4, nb_b, 0,
                                                              acc_b, 0, gm1);
                FPV_LAUNCH_CHECK("bwd_accum_kernel");
            }
            dim3 grid1((unsigned)ceil_div(M, 256), 1);
            bwd_finish_kernel<IdxT><<<grid1, 256, 0, st>>>(b, M, a, 0, nullptr, i_b2a, cmax + 4, nb_b, acc_b, grad_b, 1);
            FPV_LAUNCH_CHECK("bwd_finish_kernel");
        } else {
            bwd_finish_kernel<IdxT><<<gridM, 256, 0, st>>>(b, M, a, N * 3, g_b2a, i_b2a, cmax + 4, nb_b, acc_b, grad_b, gm1);
            FPV_LAUNCH_CHECK("bwd_finish_kernel");
        }
    }
    return FPV_OK;
}

}  // namespace fpv

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
using namespace fpv;

extern "C" {

size_t fpv_nn_planes_bytes(int64_t batches, int64_t M) {
    if (batches <= 0 || M <= 0) return 0;
    return align_up(size_t(batches) * 3 * size_t(planes_len(M)) * sizeof(float), 256);
}

int fpv_nn_pack_planes(const float *pts, int64_t batches, int64_t M, float *planes, fpv_stream_t stream) {
    FPV_CHECK_ARG(pts && planes, "fpv_nn_pack_planes: null pointer");
    return pack_planes_impl(pts, batches, M, planes, static_cast<cudaStream_t>(stream));
}

size_t fpv_nn_search_workspace_bytes(int64_t batches, int64_t N, int64_t M) {
    if (batches <= 0 || N <= 0 || M <= 0) return 0;
    const size_t a = nn_search_ws(batches, N, M, 0), b = nn_search_ws(batches, N, M, 1);
    return (a > b ? a : b) + 256;
}

int fpv_nn_search(const float *queries, int q_shared, int64_t batches, int64_t N, const float *ref_planes,
                  int64_t ref_batches, int64_t M, int64_t idx_base, float *dist, void *idx, int idx_bytes,
                  uint64_t *keys, void *workspace, size_t workspace_bytes, fpv_stream_t stream) {
    FPV_CHECK_ARG(queries && ref_planes, "fpv_nn_search: null input pointer");
    return nn_search_impl(queries, q_shared, batches, N, ref_planes, ref_batches, M, idx_base, dist, idx, idx_bytes,
                          keys, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int fpv_nn_unpack_keys(const uint64_t *keys, int64_t n, float *dist, void *idx, int idx_bytes, fpv_stream_t stream) {
    FPV_CHECK_ARG(keys && n > 0, "fpv_nn_unpack_keys: empty input");
    FPV_CHECK_ARG((idx_bytes == 4 || idx_bytes == 8) && idx, "fpv_nn_unpack_keys: idx_bytes must be 4 or 8");
    unpack_keys_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const unsigned long long *>(keys), n, dist, idx, idx_bytes);
    FPV_LAUNCH_CHECK("unpack_keys_kernel");
    return FPV_OK;
}

// Debug/tuning hook (bench sweeps): force queries-per-thread (4|8), the candidate split (0 = heuristic)
// and how many of a thread's queries use packed FP32x2 math (-1 = default).
int fpv_nn_set_tuning(int qpt, int nsplit, int packed) {
    g_tune_qpt = qpt;
    g_tune_nsplit = nsplit;
    g_tune_packed = packed;
    return FPV_OK;
}

// Debug: device buffer of 1024 int64 that receives a clock64 timeline of CTA 0 of nn_tc_kernel (NULL = off).
int fpv_nn_tc_debug(long long *dbg) {
    nn_tc_set_debug(dbg);
    return FPV_OK;
}

// engine: 0 = auto, 1 = FP32 SIMT brute force, 2 = tensor-core filter + exact re-check.
// tc_eshift: the filter's error bound is 2^-tc_eshift * max(|x|^2, max|y|^2) (0 = default 15).
int fpv_nn_set_engine(int engine, int tc_eshift) {
    g_engine = (engine >= 0 && engine <= 2) ? engine : 0;
    nn_tc_set_eshift((tc_eshift & 0xff) > 0 ? (tc_eshift & 0xff) : 15);
    nn_tc_set_subtile(tc_eshift >> 8);  // bits 8.. : accumulator sub-tile width (0 = default)
    return FPV_OK;
}

size_t fpv_chamfer_fwd_workspace_bytes(int64_t bs, int64_t N, int64_t M, int b_shared) {
    if (bs <= 0 || N <= 0 || M <= 0) return 0;
    size_t s = fpv_nn_planes_bytes(bs, N) + fpv_nn_planes_bytes(b_shared ? 1 : bs, M);
    const size_t w1 = nn_search_ws(bs, N, M, b_shared ? 1 : 0), w2 = nn_search_ws(bs, M, N, 0);
    return s + (w1 > w2 ? w1 : w2) + 512;
}

int fpv_chamfer_fwd(const float *a, const float *b, int64_t bs, int64_t N, int64_t M, int b_shared, float *d_b2a,
                    float *d_a2b, void *i_b2a, void *i_a2b, int idx_bytes, void *workspace, size_t workspace_bytes,
                    fpv_stream_t stream) {
    FPV_CHECK_ARG(a && b && d_b2a && d_a2b && i_b2a && i_a2b, "fpv_chamfer_fwd: null pointer");
    FPV_CHECK_ARG(bs > 0 && N > 0 && M > 0, "fpv_chamfer_fwd: empty cloud (bs=%lld N=%lld M=%lld)", (long long)bs,
                  (long long)N, (long long)M);
    FPV_CHECK_ARG(idx_bytes == 4 || idx_bytes == 8, "fpv_chamfer_fwd: idx_bytes must be 4 or 8");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Arena ar(workspace, workspace_bytes);
    const int64_t nb = b_shared ? 1 : bs;
    float *pa = ar.take<float>(size_t(bs) * 3 * planes_len(N));
    float *pb = ar.take<float>(size_t(nb) * 3 * planes_len(M));
    if (!pa || !pb) {
        set_error("fpv_chamfer_fwd: workspace too small (%zu bytes, need %zu)", workspace_bytes,
                  fpv_chamfer_fwd_workspace_bytes(bs, N, M, b_shared));
        return FPV_ERR_WORKSPACE;
    }
    int rc;
    if ((rc = pack_planes_impl(a, bs, N, pa, st))) return rc;
    if ((rc = pack_planes_impl(b, nb, M, pb, st))) return rc;
    void *rest = ar.base + ar.off;
    const size_t rest_bytes = ar.cap - ar.off;
    // a -> b : every a_i finds its nearest b_j
    if ((rc = nn_search_impl(a, 0, bs, N, pb, nb, M, 0, d_a2b, i_a2b, idx_bytes, nullptr, rest, rest_bytes, st)))
        return rc;
    // b -> a : every b_j finds its nearest a_i (per frame)
    if ((rc = nn_search_impl(b, b_shared ? 1 : 0, bs, M, pa, bs, N, 0, d_b2a, i_b2a, idx_bytes, nullptr, rest,
                             rest_bytes, st)))
        return rc;
    return FPV_OK;
}

size_t fpv_chamfer_bwd_workspace_bytes(int64_t bs, int64_t N, int64_t M, int b_shared, int want_grad_b) {
    if (bs <= 0 || N <= 0 || M <= 0) return 0;
    size_t s = 256 + align_up(size_t(bs * N * 3) * 8, 256);
    if (want_grad_b) s += align_up(size_t((b_shared ? 1 : bs) * M * 3) * 8, 256);
    return s + 256;
}

int fpv_chamfer_bwd_bcast(const float *a, const float *b, int64_t bs, int64_t N, int64_t M, int b_shared,
                          const float *g_b2a, const float *g_a2b, int g_broadcast, const void *i_b2a, const void *i_a2b,
                          int idx_bytes, float *grad_a, float *grad_b, void *workspace, size_t workspace_bytes,
                          fpv_stream_t stream) {
    FPV_CHECK_ARG(a && b && i_b2a && i_a2b && grad_a, "fpv_chamfer_bwd: null pointer");
    FPV_CHECK_ARG(bs > 0 && N > 0 && M > 0, "fpv_chamfer_bwd: empty cloud");
    FPV_CHECK_ARG(bs <= 65535, "fpv_chamfer_bwd: more than 65535 batches");
    FPV_CHECK_ARG(idx_bytes == 4 || idx_bytes == 8, "fpv_chamfer_bwd: idx_bytes must be 4 or 8");
    FPV_CHECK_ARG((g_broadcast & ~3) == 0, "fpv_chamfer_bwd: g_broadcast must be a combination of bits 0 and 1");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (idx_bytes == 8)
        return chamfer_bwd_impl<long long>(a, b, bs, N, M, b_shared, g_b2a, g_a2b,
                                           static_cast<const long long *>(i_b2a),
                                           static_cast<const long long *>(i_a2b), grad_a, grad_b, workspace,
                                           workspace_bytes, st, g_broadcast);
    return chamfer_bwd_impl<int>(a, b, bs, N, M, b_shared, g_b2a, g_a2b, static_cast<const int *>(i_b2a),
                                 static_cast<const int *>(i_a2b), grad_a, grad_b, workspace, workspace_bytes, st,
                                 g_broadcast);
}

int fpv_chamfer_bwd(const float *a, const float *b, int64_t bs, int64_t N, int64_t M, int b_shared, const float *g_b2a,
                    const float *g_a2b, const void *i_b2a, const void *i_a2b, int idx_bytes, float *grad_a,
                    float *grad_b, void *workspace, size_t workspace_bytes, fpv_stream_t stream) {
    return fpv_chamfer_bwd_bcast(a, b, bs, N, M, b_shared, g_b2a, g_a2b, 0, i_b2a, i_a2b, idx_bytes, grad_a, grad_b,
                                 workspace, workspace_bytes, stream);
}

}  // extern "C"
