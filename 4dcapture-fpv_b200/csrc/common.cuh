// common.cuh -- shared helpers of the sm_100a kernels (error plumbing, PTX wrappers, reductions).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fpv_b200.h"

namespace fpv {

void set_error(const char *fmt, ...);

#define FPV_CHECK_ARG(cond, ...)                 \
    do {                                         \
        if (!(cond)) {                           \
            ::fpv::set_error(__VA_ARGS__);       \
            return FPV_ERR_INVALID;              \
        }                                        \
    } while (0)

#define FPV_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            ::fpv::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                             __LINE__);                                                        \
            return FPV_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define FPV_LAUNCH_CHECK(name)                                                           \
    do {                                                                                 \
        ::fpv::count_launch();                                                           \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess) {                                                        \
            ::fpv::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__)); \
            return FPV_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Bump allocator over a caller-provided workspace (256-byte aligned slices).
struct Arena {
    char *base;
    size_t cap, off;
    Arena(void *p, size_t n) : base(static_cast<char *>(p)), cap(n), off(0) {}
    template <typename T>
    T *take(size_t count) {
        size_t bytes = align_up(count * sizeof(T), 256);
        if (off + bytes > cap) return nullptr;
        T *r = reinterpret_cast<T *>(base + off);
        off += bytes;
        return r;
    }
};

int sm_count();
void count_launch();

// One-time per-DEVICE configuration flag: function attributes such as MaxDynamicSharedMemorySize are per device, so a
// process that uses cuda:1 after cuda:0 must set them again there.  slot() returns the flag of the current device.
struct PerDeviceOnce {
    bool done[64] = {};
    bool *slot() {
        static bool overflow;
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) {
            overflow = false;  // unknown device: configure every time (cheap)
            return &overflow;
        }
        return &done[d];
    }
};

// Optional per-kernel timing (bench.py's roofline leg): when enabled, instrumented launch sites bracket
// the kernel with CUDA events on the launching stream and record its algorithmic work.
bool profile_on();
void profile_begin(const char *name, cudaStream_t st, double algo_bytes, double work_items);
void profile_end(cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// mbarrier + bulk async copy (TMA engine, 1-D form)
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// Blocks until the phase with the given parity has completed.  The suspend-time hint lets the hardware park the
// thread until the barrier flips instead of re-issuing the probe every few dozen cycles: spinning waiters were
// measured to take a third of the issue slots of nn_tc_kernel (profiles/r01_nn_tc_ncu.md).
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"(0x989680u)
            : "memory");
    } while (!done);
}
// global -> shared bulk copy completing on an mbarrier; bytes % 16 == 0, both addresses 16B aligned
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Deterministic block-wide sum: fixed shuffle tree inside a warp, fixed order across warps.
// Every thread must call it; the result is valid in thread 0.  `scratch` holds >= 32 floats.
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
template <typename T>
__device__ __forceinline__ T block_sum(T v, T *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    T r = T(0);
    if (threadIdx.x == 0) {
        for (int w = 0; w < nwarps; ++w) r += scratch[w];
    }
    return r;
}

}  // namespace fpv
