// tc_gemm.cu -- fp32-accurate GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   C[M,N] = sum_k A[m,k] * B[n,k]      (A and B both K-major, i.e. row-major [rows, K])
//
// with the 3xTF32 split  A = Ahi + Alo, B = Bhi + Blo (each part exactly representable in TF32):
//   C = Ahi.Bhi + Alo.Bhi + Ahi.Blo     (the dropped Alo.Blo term is ~2^-22 relative)
// so the result is within a few fp32 ulps of an fp32 GEMM -- the 1e-5 parity budget of the SMPL-X path
// (north_star item 2: "the pose-blendshape contraction ([T,486]x[486,3V]) runs on tensor cores").
//
// Used by smplx.cu for   v_posed[T,3V] = coef[T,512] x basis^T      (forward, 1 k-slice)
//                 and    gC[T,512]     = g_vposed[T,3V] x basis     (backward, split-K, fixed-order reduce)
//
// Kernel anatomy (one 128x128 output tile per CTA, 6 warps):
//   warp 0   TMA producer: per 32-float k-block four cp.async.bulk.tensor.2d (Ahi, Alo, Bhi, Blo tiles,
//            128 rows x 128 B, SWIZZLE_128B) into a 3-stage ring, completion on mbarriers
//   warp 1   allocates 128 TMEM columns; one lane issues 12 tcgen05.mma.kind::tf32 (128x128x8) per
//            k-block (3 products x 4 k-steps), tcgen05.commit releases the stage / signals the epilogue
//   warps 2-5  epilogue: tcgen05.ld 32x32b.x32 (one accumulator row per thread), stores to global
#include <cuda.h>

#include "common.cuh"

namespace fpv {

constexpr int TG_BM = 128, TG_BN = 128, TG_BK = 32;
constexpr int TG_STAGES = 3;
constexpr int TG_TILE_BYTES = TG_BM * TG_BK * 4;      // 16 KB
constexpr int TG_STAGE_BYTES = 4 * TG_TILE_BYTES;     // Ahi | Alo | Bhi | Blo
constexpr int TG_THREADS = 192;
constexpr size_t TG_SMEM = size_t(TG_STAGES) * TG_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

struct TGParams {
    float *C;              // output (or split-K partials)
    int64_t ldc;           // floats between rows of C
    int64_t slice_stride;  // floats between k-slices of C (0 when gridDim.z == 1)
    int M, N;
    int kblocks_total, kblocks_per_slice;
};

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *slot_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, TF32 inputs, fp32 accumulate; single-thread issue
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
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
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major operand tile, 128-byte rows, SWIZZLE_128B (what the TMA box above produces):
// start address >> 4 | LBO (unused for swizzled K-major, 1) | SBO = 1024 B between 8-row groups |
// version 1 (Blackwell) | layout type 2 (SWIZZLE_128B)         [cute::UMMA::SmemDescriptor]
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return uint64_t((smem_addr & 0x3FFFF) >> 4) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) |
           (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
// Instruction descriptor, kind::tf32: D=f32 (bits 4-5 = 1), A=B=TF32 (bits 7-9, 10-12 = 2), both K-major,
// N>>3 at bit 17, M>>4 at bit 24                                 [cute::UMMA::InstrDescriptor]
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

__global__ void __launch_bounds__(TG_THREADS, 1)
    tc_gemm_kernel(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
                   const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
                   const TGParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base_u32 = smem_u32(smem_raw);
    unsigned char *tiles = smem_raw + (((base_u32 + 1023u) & ~1023u) - base_u32);
    uint64_t *full = reinterpret_cast<uint64_t *>(tiles + size_t(TG_STAGES) * TG_STAGE_BYTES);
    uint64_t *empty = full + TG_STAGES;
    uint64_t *tmem_full = empty + TG_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // CTAs are numbered with the ROW tile fastest: the few row tiles (3 at T = 300) that share one column tile of the big
    // operand B are launched side by side, so its hi/lo tiles are fetched from HBM once and served to the others by L2
    // (with the column tile fastest every row tile streamed all of B again: 365 MB instead of ~130 MB, round-1 ncu)
    const int lin = blockIdx.x + gridDim.x * blockIdx.y;
    const int n0 = (lin / int(gridDim.y)) * TG_BN, m0 = (lin % int(gridDim.y)) * TG_BM;
    const int kb0 = blockIdx.z * p.kblocks_per_slice;
    const int kb1 = min(p.kblocks_total, kb0 + p.kblocks_per_slice);
    const int nkb = kb1 - kb0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TG_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(tmem_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TG_BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % TG_STAGES;
                const uint32_t ph = (i / TG_STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                unsigned char *st = tiles + size_t(s) * TG_STAGE_BYTES;
                const int kc = (kb0 + i) * TG_BK;
                mbar_arrive_expect_tx(&full[s], TG_STAGE_BYTES);
                tma_load_2d(st, &tmAhi, kc, m0, &full[s]);
                tma_load_2d(st + TG_TILE_BYTES, &tmAlo, kc, m0, &full[s]);
                tma_load_2d(st + 2 * TG_TILE_BYTES, &tmBhi, kc, n0, &full[s]);
                tma_load_2d(st + 3 * TG_TILE_BYTES, &tmBlo, kc, n0, &full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(TG_BM, TG_BN);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % TG_STAGES;
                const uint32_t ph = (i / TG_STAGES) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t st = smem_u32(tiles + size_t(s) * TG_STAGE_BYTES);
                const uint64_t dAhi = umma_desc_sw128(st), dAlo = umma_desc_sw128(st + TG_TILE_BYTES);
                const uint64_t dBhi = umma_desc_sw128(st + 2 * TG_TILE_BYTES), dBlo = umma_desc_sw128(st + 3 * TG_TILE_BYTES);
#pragma unroll
                for (int kk = 0; kk < TG_BK / 8; ++kk) {  // 8 tf32 = 32 bytes per MMA: advance the start address
                    const uint64_t adv = uint64_t((kk * 32) >> 4);
                    tc_mma_tf32(tmem_base, dAhi + adv, dBhi + adv, idesc, (i | kk) != 0);
                    tc_mma_tf32(tmem_base, dAlo + adv, dBhi + adv, idesc, 1);
                    tc_mma_tf32(tmem_base, dAhi + adv, dBlo + adv, idesc, 1);
                }
                tc_commit(&empty[s]);  // the stage is free once these MMAs have read it
            }
            tc_commit(tmem_full);
        }
    } else {
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int quarter = warp & 3;  // a warp may only touch TMEM lanes 32*(warp%4) .. +31
        const int row = m0 + quarter * 32 + lane;
        float *crow = p.C + int64_t(blockIdx.z) * p.slice_stride + int64_t(row) * p.ldc;
#pragma unroll 1
        for (int c = 0; c < TG_BN / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(c * 32), v);
            if (row < p.M && nkb > 0) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int col = n0 + c * 32 + i;
                    if (col < p.N) crow[col] = __uint_as_float(v[i]);
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TG_BN);
    }
}

// x -> (hi, lo), both exactly representable in TF32 (10 explicit mantissa bits), hi + lo ~= x to 2^-22
__device__ __forceinline__ float tf32_round(float x) {
    uint32_t u = __float_as_uint(x);
    u = (u + 0x1000u) & 0xFFFFE000u;
    return __uint_as_float(u);
}
__global__ void split_tf32_kernel(const float *__restrict__ src, int64_t n, float *__restrict__ hi,
                                  float *__restrict__ lo) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const float x = src[i];
        const float h = tf32_round(x);
        hi[i] = h;
        lo[i] = tf32_round(x - h);
    }
}

__global__ void splitk_reduce_ld_kernel(const float *__restrict__ part, int M, int N, int nz, float *__restrict__ out,
                                        int64_t ldc) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= int64_t(M) * N) return;
    float s = 0.f;
    for (int z = 0; z < nz; ++z) s += part[size_t(z) * M * N + i];
    out[(i / N) * ldc + (i % N)] = s;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [rows, K] fp32, row pitch ld floats -> 2-D map with a (32 floats x 128 rows) SWIZZLE_128B box
static int make_kmajor_map(CUtensorMap *tm, const float *ptr, int64_t rows, int64_t K, int64_t ld) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return FPV_ERR_CUDA;
    }
    FPV_CHECK_ARG((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * 4) % 16 == 0,
                  "tc_gemm: operand base / row pitch must be 16-byte aligned (ld=%lld)", (long long)ld);
    cuuint64_t gdim[2] = {cuuint64_t(K), cuuint64_t(rows)};
    cuuint64_t gstr[1] = {cuuint64_t(ld) * 4};
    cuuint32_t box[2] = {TG_BK, TG_BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", int(r));
        return FPV_ERR_CUDA;
    }
    return FPV_OK;
}

int tc_gemm_3xtf32(const float *a_hi, const float *a_lo, int64_t lda, const float *b_hi, const float *b_lo,
                   int64_t ldb, int M, int N, int K, float *C, int64_t ldc, int ksplit, float *partial,
                   cudaStream_t st) {
    FPV_CHECK_ARG(a_hi && a_lo && b_hi && b_lo && C, "tc_gemm: null pointer");
    FPV_CHECK_ARG(M > 0 && N > 0 && K > 0, "tc_gemm: empty problem");
    FPV_CHECK_ARG(ksplit >= 1 && (ksplit == 1 || partial), "tc_gemm: split-K needs a partial buffer");
    CUtensorMap tAh, tAl, tBh, tBl;
    int rc;
    if ((rc = make_kmajor_map(&tAh, a_hi, M, K, lda))) return rc;
    if ((rc = make_kmajor_map(&tAl, a_lo, M, K, lda))) return rc;
    if ((rc = make_kmajor_map(&tBh, b_hi, N, K, ldb))) return rc;
    if ((rc = make_kmajor_map(&tBl, b_lo, N, K, ldb))) return rc;
    static PerDeviceOnce once;
    bool *configured = once.slot();
    if (!*configured) {
        FPV_CUDA(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TG_SMEM)));
        *configured = true;
    }
    TGParams p;
    p.M = M;
    p.N = N;
    p.kblocks_total = int(ceil_div(K, TG_BK));
    p.kblocks_per_slice = int(ceil_div(p.kblocks_total, ksplit));
    const int nz = int(ceil_div(p.kblocks_total, p.kblocks_per_slice));
    if (nz == 1) {
        p.C = C;
        p.ldc = ldc;
        p.slice_stride = 0;
    } else {
        p.C = partial;
        p.ldc = N;
        p.slice_stride = int64_t(M) * N;
    }
    dim3 grid((unsigned)ceil_div(N, TG_BN), (unsigned)ceil_div(M, TG_BM), (unsigned)nz);
    tc_gemm_kernel<<<grid, TG_THREADS, TG_SMEM, st>>>(tAh, tAl, tBh, tBl, p);
    FPV_LAUNCH_CHECK("tc_gemm_kernel");
    if (nz > 1) {
        splitk_reduce_ld_kernel<<<(unsigned)ceil_div(int64_t(M) * N, 256), 256, 0, st>>>(partial, M, N, nz, C, ldc);
        FPV_LAUNCH_CHECK("splitk_reduce_ld_kernel");
    }
    return FPV_OK;
}

int split_tf32(const float *src, int64_t n, float *hi, float *lo, cudaStream_t st) {
    const int nb = int(ceil_div(n, 256) < 148 * 16 ? ceil_div(n, 256) : 148 * 16);
    split_tf32_kernel<<<nb, 256, 0, st>>>(src, n, hi, lo);
    FPV_LAUNCH_CHECK("split_tf32_kernel");
    return FPV_OK;
}

}  // namespace fpv

extern "C" {

int fpv_split_tf32(const float *src, int64_t n, float *hi, float *lo, fpv_stream_t stream) {
    FPV_CHECK_ARG(src && hi && lo && n > 0, "fpv_split_tf32: bad arguments");
    return fpv::split_tf32(src, n, hi, lo, static_cast<cudaStream_t>(stream));
}

size_t fpv_tc_gemm_workspace_bytes(int M, int N, int ksplit) {
    return ksplit > 1 ? fpv::align_up(size_t(ksplit) * M * N * sizeof(float), 256) : 0;
}

int fpv_tc_gemm_3xtf32(const float *a_hi, const float *a_lo, int64_t lda, const float *b_hi, const float *b_lo,
                       int64_t ldb, int M, int N, int K, float *C, int64_t ldc, int ksplit, void *workspace,
                       size_t workspace_bytes, fpv_stream_t stream) {
    FPV_CHECK_ARG(ksplit <= 1 || workspace_bytes >= fpv_tc_gemm_workspace_bytes(M, N, ksplit),
                  "fpv_tc_gemm_3xtf32: workspace too small");
    return fpv::tc_gemm_3xtf32(a_hi, a_lo, lda, b_hi, b_lo, ldb, M, N, K, C, ldc, ksplit,
                               static_cast<float *>(workspace), static_cast<cudaStream_t>(stream));
}

}  // extern "C"
