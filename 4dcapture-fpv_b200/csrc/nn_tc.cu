// nn_tc.cu -- exact nearest neighbour with a tensor-core FILTER (tcgen05 + TMEM) and an exact FP32 re-check.
//
// north_star item 1: "uses the tensor-core GEMM form (|x|^2+|y|^2-2x.y) only if ncu shows it beats the FP32
// SIMT path, and then only with an exact FP32 re-check of the winning candidates".  This is that path.
// Results are bit-identical to nn_search_kernel (nn_search.cu): the tensor cores only decide which
// candidates are worth an exact evaluation.
//
//   approx  D~[i][j] = |x_i|^2 + |y_j|^2 - 2 x_i.y_j   as ONE K=16 TF32 contraction per pair:
//           every coordinate is split x = xh + xl (TF32 pieces) and the three products xh*yh, xh*yl, xl*yh
//           occupy three K slots; |x|^2 and |y|^2 are split in three TF32 pieces against constant 1s.
//           |D~ - d_canonical| <= E_i := 2^-e * max(|x_i|^2, max_j |y_j|^2)   (e = 15 by default; the
//           measured error is ~2^-19 of that scale, tests/test_nn_tc_gpu.py sweeps e to show the margin)
//   exact   per query row, per group of 32 accumulator columns: a FMNMX3 chain gives the group's minimum;
//           only when it is <= (best exact distance so far + E_i) are the columns re-read and the
//           candidates below the threshold re-evaluated with the canonical fp32 expression.  Candidates
//           are visited in increasing index order and updates use strict '<', so the lowest index wins.
//           Proof sketch: the true winner j* has D~[j*] <= d(j*) + E <= best + E whenever it is met.
//
// CTA = 12 warps: 0 bulk-copy producer (candidate planes), 1-2 operand converters (fp32 planes -> TF32
// split K=16 operand in the UMMA no-swizzle K-major layout), 3 TMEM allocator + MMA issuer
// (tcgen05.mma.kind::tf32 128x256x8, two per tile), 4-11 epilogue (tcgen05.ld 32x32b.x32).
// Each CTA keeps QT query tiles (128 rows each) resident as A operands and streams candidate tiles of 256.
#include <math_constants.h>

#include "common.cuh"

namespace fpv {

constexpr int TC_SETS = 4;         // epilogue warp sets (4 warps each, one per TMEM lane quarter)
constexpr int TC_TPS = 1;          // query tiles per epilogue set
constexpr int TC_QT = TC_TPS * TC_SETS;  // query tiles per CTA
constexpr int TC_M = 128;          // rows per query tile (MMA M)
constexpr int TC_N = 256;          // candidates per tile (MMA N)
constexpr int TC_K = 16;           // TF32 slots per pair
constexpr int TC_STAGES = 3;
constexpr int TC_THREADS = 128 + TC_SETS * 128;
constexpr int TC_PLANE_BYTES = 3 * TC_N * 4;             // 3 KB   fp32 xyz planes of one candidate tile
constexpr int TC_BOP_BYTES = TC_N * TC_K * 4;            // 16 KB  TF32 operand of one candidate tile
constexpr int TC_AOP_BYTES = TC_M * TC_K * 4;            // 8 KB   TF32 operand of one query tile
constexpr int TC_STAGE_BYTES = TC_PLANE_BYTES + TC_BOP_BYTES;
constexpr int TC_STATE_WORDS = 2;  // per (tile, row): best d, best j (final hand-off to the writer threads)
constexpr size_t TC_SMEM = size_t(TC_QT) * TC_AOP_BYTES + size_t(TC_STAGES) * TC_STAGE_BYTES +
                           size_t(TC_QT) * TC_M * TC_STATE_WORDS * 4 + 256 /*barriers*/ + 128 /*align*/;

static int g_tc_eshift = 15;
static long long *g_tc_dbg = nullptr;
static int g_tc_ns = 128;  // accumulator sub-tile width (64 | 128)

struct NNTCParams {
    const float *q;
    int64_t q_bstride;
    int64_t N;
    const float *planes;
    int64_t plane_bstride;
    int64_t Mp, M;
    int64_t chunk;  // candidates per blockIdx.y slice (multiple of TC_N)
    int64_t idx_base;
    const float *ymax;  // [candidate batches] max_j |y_j|^2 (finite part)
    int64_t ymax_bstride;
    float escale;  // 2^-e
    float *dist;
    void *idx;
    int idx_bytes;
    unsigned long long *keys;
    int keys_atomic;
    long long *dbg;  // optional timeline of CTA 0 (tools/nn_tc_timeline.py): [4][256] clock64 stamps
};

// ---- PTX wrappers (tcgen05) -------------------------------------------------------------------
__device__ __forceinline__ void tcx_alloc(uint32_t *slot_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tcx_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tcx_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcx_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcx_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tcx_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// TMEM -> registers, 32 lanes x 32 consecutive columns (one accumulator row segment per thread).  The load is
// asynchronous: tcx_ld32_issue starts it, tcx_ld32_wait makes the registers valid.  The wait takes the
// registers as read-write operands so the compiler cannot hoist a use above it.
__device__ __forceinline__ void tcx_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tcx_ld32_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                   "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                   "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}
// one lane of a converged warp (CUTLASS elect_one_sync idiom)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major operand without swizzle: 8-row x 16-byte core matrices; address(row r, 16B chunk c) =
// (r % 8) * 16 + (r / 8) * SBO + c * LBO.  One MMA (K = 8 tf32 = 32 B) reads chunks c and c+1.
__device__ __forceinline__ uint64_t umma_desc_plain(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return uint64_t((smem_addr & 0x3FFFF) >> 4) | (uint64_t(lbo_bytes >> 4) << 16) | (uint64_t(sbo_bytes >> 4) << 32) |
           (uint64_t(1) << 46);
}
__host__ __device__ constexpr uint32_t umma_idesc_tf32_mn(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

__device__ __forceinline__ float tf32r(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

__device__ __forceinline__ float d2_canon(float x, float y, float z, float rx, float ry, float rz) {
    const float dx = __fsub_rn(x, rx), dy = __fsub_rn(y, ry), dz = __fsub_rn(z, rz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

// write one operand row (16 tf32 values) into the chunk-major no-swizzle layout: 4 chunks of 16 bytes,
// chunk c of row r at  base + c * (rows*16) + (r/8)*128 + (r%8)*16
__device__ __forceinline__ void store_row16(unsigned char *base, int rows, int r, const float (&k)[16]) {
    unsigned char *p = base + (r >> 3) * 128 + (r & 7) * 16;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        *reinterpret_cast<float4 *>(p + size_t(c) * rows * 16) = make_float4(k[4 * c], k[4 * c + 1], k[4 * c + 2], k[4 * c + 3]);
}

__device__ __forceinline__ void split3(float v, float &h, float &m, float &l) {
    h = tf32r(v);
    const float r = v - h;
    m = tf32r(r);
    l = tf32r(r - m);
}

// Exact re-evaluation of one candidate column (canonical fp32 expression, strict '<': lowest index wins).
__device__ __forceinline__ void tc_eval(const float *px, const float *py, const float *pz, int col, int jglob, int jmax,
                                        float qx, float qy, float qz, float &bd, int &bj) {
    if (jglob < jmax) {
        const float d = d2_canon(qx, qy, qz, px[col], py[col], pz[col]);
        if (d < bd) {
            bd = d;
            bj = jglob;
        }
    }
}

// Exact re-check of the columns of one 32-column group that fall under the threshold (slow path).
__device__ __forceinline__ void tc_recheck(const uint32_t (&r)[32], const float (&pm)[8], const float *px,
                                           const float *py, const float *pz, int col0, int jglob0, int jmax, float qx,
                                           float qy, float qz, float qe, float &bd, int &bj) {
    float thresh = bd + qe;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        if (!(pm[c] > thresh)) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (!(__uint_as_float(r[4 * c + i]) > thresh)) {
                    tc_eval(px, py, pz, col0 + 4 * c + i, jglob0 + 4 * c + i, jmax, qx, qy, qz, bd, bj);
                    thresh = bd + qe;
                }
            }
        }
    }
}

__device__ __forceinline__ float tc_tree(const uint32_t (&r)[32], float (&pm)[8]) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
        pm[c] = fminf(fmin3(__uint_as_float(r[4 * c]), __uint_as_float(r[4 * c + 1]), __uint_as_float(r[4 * c + 2])),
                      __uint_as_float(r[4 * c + 3]));
    float m = fmin3(pm[0], pm[1], pm[2]);
    m = fmin3(m, pm[3], pm[4]);
    m = fmin3(m, pm[5], pm[6]);
    return fminf(m, pm[7]);
}

// One group of 32 accumulator columns of one query row: FMNMX3 tree, and -- only if the group's minimum
// can beat the best exact distance -- the exact re-check of the columns under the threshold.
__device__ __forceinline__ void tc_group(const uint32_t (&r)[32], const float *px, const float *py, const float *pz,
                                         int col0, int jglob0, int jmax, float qx, float qy, float qz, float qe,
                                         float &bd, int &bj) {
    float pm[8];
    const float m = tc_tree(r, pm);
    if (!(m > bd + qe))  // rare on long scans; no warp-aligned instruction inside, so divergence is legal
        tc_recheck(r, pm, px, py, pz, col0, jglob0, jmax, qx, qy, qz, qe, bd, bj);
}

// NS = accumulator sub-tile width (columns per MMA): 512/NS TMEM buffers are in flight, which is what hides the
// MMA -> commit -> epilogue -> release round trip (each sub-tile is only a K=16 contraction, ~NS/2 tensor cycles).
template <int NS>
__global__ void __launch_bounds__(TC_THREADS, 1) nn_tc_kernel(const NNTCParams p) {
    constexpr int NBUF = 512 / NS;        // TMEM accumulator buffers
    constexpr int NSUB = TC_N / NS;       // sub-tiles per 256-candidate stage
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base_u32 = smem_u32(smem_raw);
    unsigned char *sm = smem_raw + (((base_u32 + 127u) & ~127u) - base_u32);
    unsigned char *aop = sm;                                            // [QT][8 KB]
    unsigned char *stages = aop + size_t(TC_QT) * TC_AOP_BYTES;         // [STAGES][planes 3 KB | bop 16 KB]
    float *state = reinterpret_cast<float *>(stages + size_t(TC_STAGES) * TC_STAGE_BYTES);  // [2][QT][128]
    uint64_t *bars = reinterpret_cast<uint64_t *>(state + size_t(TC_QT) * TC_M * TC_STATE_WORDS);
    uint64_t *pl_full = bars;                      // [STAGES] producer -> converters, epilogue
    uint64_t *bop_full = pl_full + TC_STAGES;      // [STAGES] converters -> MMA
    uint64_t *stage_empty = bop_full + TC_STAGES;  // [STAGES] epilogue -> producer
    uint64_t *acc_full = stage_empty + TC_STAGES;  // [NBUF] MMA -> epilogue
    uint64_t *acc_empty = acc_full + 8;            // [NBUF] epilogue -> MMA
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t b = blockIdx.z;
    const int64_t r0 = int64_t(blockIdx.y) * p.chunk;
    const int64_t r1 = (r0 + p.chunk < p.M) ? (r0 + p.chunk) : p.M;
    const int ntiles = int((r1 - r0 + TC_N - 1) / TC_N);
    const int64_t qbase = int64_t(blockIdx.x) * (TC_QT * TC_M);
    const float *qsrc = p.q + b * p.q_bstride;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&pl_full[s], 1);
            mbar_init(&bop_full[s], 2);
            mbar_init(&stage_empty[s], 4 * TC_SETS);
        }
        for (int i = 0; i < NBUF; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4);
        }
        mbar_fence_init();
    }
    if (warp == 3) tcx_alloc(tmem_slot, 512);

    // ---- epilogue threads: set e (warps 4+4e .. 7+4e) owns query tiles e, e+TC_SETS, ..
    //      Each thread keeps the search state of its row in registers. ----
    constexpr int TPS = TC_TPS;
    // Every TMEM buffer must always be consumed by the same epilogue set: a waiter may only ever be one mbarrier
    // phase behind (parity waits alias modulo 2), which static ownership guarantees.
    static_assert(NBUF % TC_SETS == 0 && TC_QT % TC_SETS == 0, "accumulator buffers are statically owned by epilogue sets");
    const bool is_epi = warp >= 4;
    const int eset = (warp - 4) >> 2;  // which epilogue set
    const int quarter = warp & 3;      // TMEM lane quarter == warp index % 4
    const int row = quarter * 32 + lane;
    float qx[TPS], qy[TPS], qz[TPS], qe[TPS], bd[TPS];
    int bj[TPS];
    if (is_epi) {
        const float ymax = p.ymax[b * p.ymax_bstride];
#pragma unroll
        for (int u = 0; u < TPS; ++u) {
            const int t = TC_SETS * u + eset;
            int64_t qi = qbase + int64_t(t) * TC_M + row;
            if (qi > p.N - 1) qi = p.N - 1;
            const float x = __ldg(qsrc + 3 * qi), y = __ldg(qsrc + 3 * qi + 1), z = __ldg(qsrc + 3 * qi + 2);
            const float n = __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
            const float sc = fmaxf(n, ymax);
            qx[u] = x;
            qy[u] = y;
            qz[u] = z;
            // non-finite scale: the filter cannot be trusted for this row -> always re-check (bound = +inf)
            qe[u] = (tf32r(sc) < CUDART_INF_F) ? __fmul_rn(sc, p.escale) : CUDART_INF_F;
            bd[u] = CUDART_INF_F;
            bj[u] = 0;
            float k[16];
            float h, m, l;
            k[0] = k[1] = tf32r(x);
            k[2] = tf32r(x - k[0]);
            k[3] = k[4] = tf32r(y);
            k[5] = tf32r(y - k[3]);
            k[6] = k[7] = tf32r(z);
            k[8] = tf32r(z - k[6]);
            split3(n, h, m, l);
            k[9] = h;
            k[10] = m;
            k[11] = 1.f;
            k[12] = 1.f;
            k[13] = l;
            k[14] = 1.f;
            k[15] = 0.f;
            if (!(h < CUDART_INF_F)) {  // keep inf/NaN out of the tensor cores; this row re-checks everything
#pragma unroll
                for (int i = 0; i < 16; ++i) k[i] = 0.f;
            }
            store_row16(aop + size_t(t) * TC_AOP_BYTES, TC_M, row, k);
        }
        fence_proxy_async();
    }
    tcx_fence_before();
    __syncthreads();
    tcx_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------- producer: candidate planes -> smem ----------------
        if (lane == 0) {
            const float *src = p.planes + b * p.plane_bstride + r0;
            for (int k = 0; k < ntiles; ++k) {
                const int s = k % TC_STAGES;
                const uint32_t ph = (k / TC_STAGES) & 1;
                mbar_wait(&stage_empty[s], ph ^ 1);
                int64_t cnt = p.Mp - (r0 + int64_t(k) * TC_N);  // planes are padded to Mp (multiple of 32) with +inf
                if (cnt > TC_N) cnt = TC_N;
                const uint32_t bytes = uint32_t(cnt) * 4u;
                float *dst = reinterpret_cast<float *>(stages + size_t(s) * TC_STAGE_BYTES);
                const float *g = src + int64_t(k) * TC_N;
                mbar_arrive_expect_tx(&pl_full[s], 3 * bytes);
                bulk_g2s(dst, g, bytes, &pl_full[s]);
                bulk_g2s(dst + TC_N, g + p.Mp, bytes, &pl_full[s]);
                bulk_g2s(dst + 2 * TC_N, g + 2 * p.Mp, bytes, &pl_full[s]);
            }
        }
    } else if (warp == 1 || warp == 2) {
        // ---------------- converters: fp32 planes -> TF32-split B operand ----------------
        const int ct = (warp - 1) * 32 + lane;  // 0..63, four candidates each
        for (int k = 0; k < ntiles; ++k) {
            const int s = k % TC_STAGES;
            const uint32_t ph = (k / TC_STAGES) & 1;
            mbar_wait(&pl_full[s], ph);
            unsigned char *st = stages + size_t(s) * TC_STAGE_BYTES;
            const float *px = reinterpret_cast<const float *>(st), *py = px + TC_N, *pz = py + TC_N;
            unsigned char *bop = st + TC_PLANE_BYTES;
            const int64_t jt0 = r0 + int64_t(k) * TC_N;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = ct + u * 64;
                float x = 0.f, y = 0.f, z = 0.f, m = 1e30f;
                if (jt0 + c < r1) {
                    x = px[c];
                    y = py[c];
                    z = pz[c];
                    m = __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
                    if (!(tf32r(m) < CUDART_INF_F)) {  // non-finite candidate: park it far away; the exact re-check decides
                        x = y = z = 0.f;
                        m = 1e30f;
                    }
                }
                float k16[16];
                float h, mm, l;
                const float xh = tf32r(x), yh = tf32r(y), zh = tf32r(z);
                k16[0] = -2.f * xh;
                k16[1] = -2.f * tf32r(x - xh);
                k16[2] = k16[0];
                k16[3] = -2.f * yh;
                k16[4] = -2.f * tf32r(y - yh);
                k16[5] = k16[3];
                k16[6] = -2.f * zh;
                k16[7] = -2.f * tf32r(z - zh);
                k16[8] = k16[6];
                split3(m, h, mm, l);
                k16[9] = 1.f;
                k16[10] = 1.f;
                k16[11] = h;
                k16[12] = mm;
                k16[13] = 1.f;
                k16[14] = l;
                k16[15] = 0.f;
                store_row16(bop, TC_N, c, k16);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bop_full[s]);
        }
    } else if (warp == 3) {
        // ---------------- MMA issuer (whole warp converged; one elected lane issues) ----------------
        constexpr uint32_t idesc = umma_idesc_tf32_mn(TC_M, NS);
        const uint32_t aop_u32 = smem_u32(aop);
        uint64_t da0[TC_QT], da1[TC_QT];
#pragma unroll
        for (int t = 0; t < TC_QT; ++t) {
            da0[t] = umma_desc_plain(aop_u32 + uint32_t(t) * TC_AOP_BYTES, TC_M * 16, 128);
            da1[t] = umma_desc_plain(aop_u32 + uint32_t(t) * TC_AOP_BYTES + 2 * TC_M * 16, TC_M * 16, 128);
        }
        uint32_t seq = 0;
        for (int k = 0; k < ntiles; ++k) {
            const int s = k % TC_STAGES;
            const uint32_t ph = (k / TC_STAGES) & 1;
            mbar_wait(&bop_full[s], ph);
            tcx_fence_after();
            const uint32_t bop_u32 = smem_u32(stages + size_t(s) * TC_STAGE_BYTES + TC_PLANE_BYTES);
#pragma unroll 1
            for (int sub = 0; sub < NSUB; ++sub) {
                // rows [sub*NS, +NS) of the stage operand: NS/8 core-matrix groups of 128 bytes further on
                const uint32_t b_u32 = bop_u32 + uint32_t(sub) * (NS / 8) * 128;
                const uint64_t db0 = umma_desc_plain(b_u32, TC_N * 16, 128);
                const uint64_t db1 = umma_desc_plain(b_u32 + 2 * TC_N * 16, TC_N * 16, 128);
#pragma unroll
                for (int t = 0; t < TC_QT; ++t, ++seq) {
                    const uint32_t buf = seq % NBUF;
                    mbar_wait(&acc_empty[buf], ((seq / NBUF) & 1) ^ 1);
                    tcx_fence_after();
                    if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && seq < 256 && lane == 0)
                        p.dbg[seq] = clock64();
                    if (elect_one()) {
                        const uint32_t d = tmem_base + buf * NS;
                        tcx_mma_tf32(d, da0[t], db0, idesc, 0);
                        tcx_mma_tf32(d, da1[t], db1, idesc, 1);
                        tcx_commit(&acc_full[buf]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ---------------- epilogue: filter + exact re-check ----------------
        const int jmax = int(r1);
        const uint32_t tlane = tmem_base + (uint32_t(quarter * 32) << 16);
        for (int k = 0; k < ntiles; ++k) {
            const int s = k % TC_STAGES;
            const uint32_t ph = (k / TC_STAGES) & 1;
            mbar_wait(&pl_full[s], ph);  // acquire the TMA-written planes for the exact re-check
            const float *px = reinterpret_cast<const float *>(stages + size_t(s) * TC_STAGE_BYTES);
            const float *py = px + TC_N, *pz = py + TC_N;
            const int jt0 = int(r0) + k * TC_N;
#pragma unroll 1
            for (int sub = 0; sub < NSUB; ++sub) {
#pragma unroll
                for (int u = 0; u < TPS; ++u) {
                    const uint32_t seq = uint32_t((k * NSUB + sub) * TC_QT + TC_SETS * u + eset);
                    const uint32_t buf = seq % NBUF;
                    const bool stamp = p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && seq < 256 &&
                                       quarter == 0 && lane == 0;
                    if (stamp) p.dbg[256 + seq] = clock64();
                    mbar_wait(&acc_full[buf], (seq / NBUF) & 1);
                    if (stamp) p.dbg[512 + seq] = clock64();
                    tcx_fence_after();
                    const uint32_t taddr0 = tlane + buf * NS;
                    const int c0 = sub * NS;
                    uint32_t va[32];
#pragma unroll 1
                    for (int g = 0; g < NS / 32; ++g) {  // 4 warps per scheduler hide the TMEM load latency
                        tcx_ld32_issue(taddr0 + uint32_t(g * 32), va);
                        tcx_ld32_wait(va);
                        tc_group(va, px, py, pz, c0 + g * 32, jt0 + c0 + g * 32, jmax, qx[u], qy[u], qz[u], qe[u], bd[u], bj[u]);
                    }
                    tcx_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[buf]);
                    if (stamp) p.dbg[768 + seq] = clock64();
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&stage_empty[s]);
        }
#pragma unroll
        for (int u = 0; u < TPS; ++u) {
            const int t = TC_SETS * u + eset;
            state[t * TC_M + row] = bd[u];
            state[(TC_QT + t) * TC_M + row] = __int_as_float(bj[u]);
        }
    }
    tcx_fence_before();
    __syncthreads();
    if (warp == 3) {
        tcx_fence_after();
        tcx_dealloc(tmem_base, 512);
    }
    if (is_epi && eset == 0) {
        for (int t = 0; t < TC_QT; ++t) {
            const int64_t qi = qbase + int64_t(t) * TC_M + row;
            if (qi >= p.N) continue;
            const float d = state[t * TC_M + row];
            const int j = __float_as_int(state[(TC_QT + t) * TC_M + row]);
            const int64_t gi = p.idx_base + j;
            const int64_t o = b * p.N + qi;
            if (p.keys) {
                const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(d)) << 32) |
                                               static_cast<unsigned long long>(static_cast<uint32_t>(gi));
                if (p.keys_atomic)
                    atomicMin(p.keys + o, key);
                else
                    p.keys[o] = key;
            } else {
                p.dist[o] = d;
                if (p.idx_bytes == 8)
                    static_cast<long long *>(p.idx)[o] = gi;
                else if (p.idx_bytes == 4)
                    static_cast<int *>(p.idx)[o] = int(gi);
            }
        }
    }
}

// max_j |y_j|^2 over the finite candidates of each batch (scale of the filter's error bound)
__global__ void nn_tc_ymax_kernel(const float *__restrict__ planes, int64_t M, int64_t Mp, float *__restrict__ ymax) {
    const int64_t b = blockIdx.y;
    const float *X = planes + b * 3 * Mp, *Y = X + Mp, *Z = Y + Mp;
    float m = 0.f;
    for (int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; j < M; j += int64_t(gridDim.x) * blockDim.x) {
        const float n = __fmaf_rn(Z[j], Z[j], __fmaf_rn(Y[j], Y[j], __fmul_rn(X[j], X[j])));
        if (n < CUDART_INF_F) m = fmaxf(m, n);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<unsigned *>(ymax + b), __float_as_uint(m));
}

// ---------------------------------------------------------------------------------------------
// host side (called from nn_search.cu)
// ---------------------------------------------------------------------------------------------
size_t nn_tc_workspace_bytes(int64_t cand_batches) { return align_up(size_t(cand_batches) * sizeof(float), 256); }

void nn_tc_set_debug(long long *dbg) { g_tc_dbg = dbg; }
void nn_tc_set_eshift(int e) { g_tc_eshift = (e >= 8 && e <= 30) ? e : 15; }
void nn_tc_set_subtile(int ns) { g_tc_ns = (ns == 64 || ns == 128) ? ns : 128; }

template <int NS>
static cudaError_t nn_tc_dispatch(const NNTCParams &p, dim3 grid, cudaStream_t st) {
    static PerDeviceOnce once;
    bool *configured = once.slot();
    if (!*configured) {
        cudaError_t e = cudaFuncSetAttribute(nn_tc_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TC_SMEM));
        if (e != cudaSuccess) return e;
        *configured = true;
    }
    nn_tc_kernel<NS><<<grid, TC_THREADS, TC_SMEM, st>>>(p);
    return cudaGetLastError();
}

void nn_tc_plan(int64_t eb, int64_t eN, int64_t M, int nsplit_hint, int *nsplit, int64_t *chunk) {
    const int64_t qblocks = ceil_div(eN, int64_t(TC_QT) * TC_M);
    const int64_t ntiles = ceil_div(M, TC_N);
    int64_t ns = nsplit_hint > 0 ? nsplit_hint : 1;
    if (nsplit_hint <= 0) {
        const int64_t target = int64_t(sm_count()) * 8, base = qblocks * eb;
        ns = ceil_div(target, base);
        const int64_t max_ns = ntiles / 16 > 1 ? ntiles / 16 : 1;
        if (ns > max_ns) ns = max_ns;
    }
    if (ns > 65535) ns = 65535;
    if (ns > ntiles) ns = ntiles;
    *chunk = ceil_div(ntiles, ns) * TC_N;
    *nsplit = int(ceil_div(M, *chunk));
}

int nn_tc_launch(const float *queries, int64_t q_bstride, int64_t eb, int64_t eN, const float *planes,
                 int64_t plane_bstride, int64_t cand_batches, int64_t Mp, int64_t M, int64_t idx_base, float *dist,
                 void *idx, int idx_bytes, unsigned long long *keys, int keys_atomic, int nsplit, int64_t chunk,
                 float *ymax_ws, cudaStream_t st) {
    FPV_CUDA(cudaMemsetAsync(ymax_ws, 0, size_t(cand_batches) * sizeof(float), st));
    {
        int nb = int(ceil_div(M, 256 * 8));
        if (nb > 592) nb = 592;
        dim3 grid((unsigned)nb, (unsigned)cand_batches);
        nn_tc_ymax_kernel<<<grid, 256, 0, st>>>(planes, M, Mp, ymax_ws);
        FPV_LAUNCH_CHECK("nn_tc_ymax_kernel");
    }
    NNTCParams p;
    p.q = queries;
    p.q_bstride = q_bstride;
    p.N = eN;
    p.planes = planes;
    p.plane_bstride = plane_bstride;
    p.Mp = Mp;
    p.M = M;
    p.chunk = chunk;
    p.idx_base = idx_base;
    p.ymax = ymax_ws;
    p.ymax_bstride = (cand_batches > 1) ? 1 : 0;
    p.escale = ldexpf(1.f, -g_tc_eshift);
    p.dist = dist;
    p.idx = idx;
    p.idx_bytes = idx_bytes;
    p.keys = keys;
    p.keys_atomic = keys_atomic;
    p.dbg = g_tc_dbg;
    dim3 grid((unsigned)ceil_div(eN, int64_t(TC_QT) * TC_M), (unsigned)nsplit, (unsigned)eb);
    cudaError_t e = g_tc_ns == 64 ? nn_tc_dispatch<64>(p, grid, st) : nn_tc_dispatch<128>(p, grid, st);
    count_launch();
    if (e != cudaSuccess) {
        set_error("nn_tc_kernel launch failed: %s", cudaGetErrorString(e));
        return FPV_ERR_CUDA;
    }
    return FPV_OK;
}

}  // namespace fpv
