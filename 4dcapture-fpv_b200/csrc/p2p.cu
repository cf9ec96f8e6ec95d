// p2p.cu -- peer-memory mailboxes for the scene-sharded step.  sm_100a, NVLink 5 / NVSwitch.
//
// One process per GPU.  Every rank owns one device allocation (its MAILBOX) that all other ranks map through CUDA IPC,
// so a kernel on rank r can store straight into rank q's memory over NVLink.  The sharded step uses it three ways:
//   keys  : the body -> scene search (nn_culled.cu) stores each query's packed (distance, global index) key into slot r
//           of EVERY rank's mailbox from its own epilogue -- the transfer is fused into the search, warp by warp;
//           after a flag barrier every rank takes the element-wise minimum over the slots of its own mailbox
//           (p2p_min_unpack_kernel), which IS the lexicographic (d, index) minimum over all scene shards;
//   grads : the small parameter-gradient vector is pushed the same way and summed in rank order (deterministic and
//           bit-identical on every rank);
//   flags : a system-scope release/acquire flag exchange is the barrier between "pushed" and "read".
// Nothing here calls NCCL, allocates at run time or synchronises with the host, so the whole sharded step can be captured
// in a CUDA graph (round 1's NCCL combine could not).  A rank that never arrives makes the barrier time out (error flag,
// checked by the host) instead of hanging the GPU.
#include <string.h>

#include "common.cuh"

namespace fpv {

struct BarrierParams {
    unsigned *flags[FPV_MAX_PEERS + 1];  // flags[q] = base of rank q's flag array ([world] words), own included
    unsigned *epoch;                     // local: number of barriers completed
    unsigned *error;                     // local: set to 1 on time-out
    int rank, world;
    unsigned long long timeout_ns;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// One CTA, one thread per rank: signal every peer, then wait for every peer's signal.
__global__ void p2p_barrier_kernel(const BarrierParams p) {
    __shared__ unsigned next;
    if (threadIdx.x == 0) next = *p.epoch + 1u;
    __syncthreads();
    const unsigned e = next;
    const int q = threadIdx.x;
    if (q < p.world) {
        __threadfence_system();  // everything this rank wrote before (earlier kernels, peer stores) is visible first
        unsigned *theirs = p.flags[q] + p.rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(e) : "memory");
        const unsigned *mine = p.flags[p.rank] + q;
        const unsigned long long t0 = globaltimer_ns();
        unsigned v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if (int(v - e) >= 0) break;
            if (globaltimer_ns() - t0 > p.timeout_ns) {
                *p.error = 1u;
                break;
            }
            __nanosleep(200);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *p.epoch = e;
}

// out = element-wise min over the `world` slots of the local mailbox half selected by *parity, unpacked to
// (dist, idx) and written through an optional permutation of the positions inside a row (sorted query position ->
// original position).  Thread 0 of the grid flips the parity word afterwards (the half is then free for the step after next).
__global__ void __launch_bounds__(256) p2p_min_unpack_kernel(const unsigned long long *__restrict__ slots, int world,
                                                             int64_t slot_stride, int64_t half_stride,
                                                             unsigned *parity, int64_t n, int64_t row,
                                                             const long long *__restrict__ perm, int64_t perm_bstride,
                                                             float *__restrict__ dist,
                                                             void *__restrict__ idx, int idx_bytes,
                                                             unsigned long long *__restrict__ keys_out) {
    const int64_t q = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const unsigned par = parity ? (*parity & 1u) : 0u;
    if (q < n) {
        const unsigned long long *s = slots + (par ? half_stride : 0) + q;
        unsigned long long k = s[0];
        for (int r = 1; r < world; ++r) {
            const unsigned long long v = s[int64_t(r) * slot_stride];
            k = v < k ? v : k;
        }
        int64_t o = q;
        if (perm) {
            const int64_t b = q / row;
            o = b * row + perm[b * perm_bstride + (q - b * row)];
        }
        if (keys_out) keys_out[o] = k;
        if (dist) dist[o] = __uint_as_float(unsigned(k >> 32));
        if (idx) {
            if (idx_bytes == 8)
                static_cast<long long *>(idx)[o] = (long long)(k & 0xFFFFFFFFull);
            else
                static_cast<int *>(idx)[o] = int(unsigned(k & 0xFFFFFFFFull));
        }
    }
}

__global__ void p2p_flip_kernel(unsigned *parity) { *parity ^= 1u; }

struct PushParams {
    float *dst[FPV_MAX_PEERS + 1];  // slot `rank` of every rank's mailbox (own included)
    int n_dst;
    const unsigned *parity;
    int64_t half_stride;
};

// src[n] -> the sender's slot in every mailbox
__global__ void __launch_bounds__(256) p2p_push_kernel(const float *__restrict__ src, int64_t n, const PushParams p) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t off = (p.parity && (*p.parity & 1u)) ? p.half_stride : 0;
    const float v = src[i];
    for (int r = 0; r < p.n_dst; ++r) p.dst[r][off + i] = v;
}

// out[i] = sum over ranks, in rank order: deterministic and identical on every rank
__global__ void __launch_bounds__(256) p2p_sum_kernel(const float *__restrict__ slots, int world, int64_t slot_stride,
                                                      int64_t half_stride, const unsigned *parity, int64_t n,
                                                      float *__restrict__ out) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *s = slots + ((parity && (*parity & 1u)) ? half_stride : 0) + i;
    float acc = s[0];
    for (int r = 1; r < world; ++r) acc = __fadd_rn(acc, s[int64_t(r) * slot_stride]);
    out[i] = acc;
}

struct GatherParams {
    int64_t offset[FPV_MAX_PEERS + 1];  // where rank r's piece goes in `out` (elements)
    int64_t count[FPV_MAX_PEERS + 1];   // its length
    int world;
};

// out[offset[r] + i] = slot r [i]: the pieces the ranks pushed (an all-gather), read from the local mailbox half
__global__ void __launch_bounds__(256) p2p_gather_kernel(const float *__restrict__ slots, int64_t slot_stride,
                                                         int64_t half_stride, const unsigned *parity, const GatherParams g,
                                                         float *__restrict__ out) {
    const int r = blockIdx.y;
    const float *s = slots + ((parity && (*parity & 1u)) ? half_stride : 0) + int64_t(r) * slot_stride;
    float *o = out + g.offset[r];
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < g.count[r]; i += int64_t(gridDim.x) * blockDim.x)
        o[i] = s[i];
}

}  // namespace fpv

using namespace fpv;

extern "C" {

/* Mailbox memory: a plain cudaMalloc allocation (exportable through CUDA IPC), zero-filled. */
int fpv_p2p_alloc(size_t bytes, void **ptr_host) {
    FPV_CHECK_ARG(ptr_host && bytes > 0, "fpv_p2p_alloc: bad arguments");
    void *p = nullptr;
    FPV_CUDA(cudaMalloc(&p, bytes));
    FPV_CUDA(cudaMemset(p, 0, bytes));
    FPV_CUDA(cudaDeviceSynchronize());
    *ptr_host = p;
    return FPV_OK;
}

int fpv_p2p_free(void *ptr) {
    if (ptr) FPV_CUDA(cudaFree(ptr));
    return FPV_OK;
}

/* 64-byte CUDA IPC handle of a mailbox; ship it to the other ranks by any host channel. */
int fpv_p2p_export(void *ptr, unsigned char *handle64_host) {
    FPV_CHECK_ARG(ptr && handle64_host, "fpv_p2p_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    FPV_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle64_host, &h, 64);
    return FPV_OK;
}

/* Map another rank's mailbox into this process (enables peer access on first use). */
int fpv_p2p_open(const unsigned char *handle64_host, void **peer_ptr_host) {
    FPV_CHECK_ARG(handle64_host && peer_ptr_host, "fpv_p2p_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64_host, 64);
    void *p = nullptr;
    FPV_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *peer_ptr_host = p;
    return FPV_OK;
}

int fpv_p2p_close(void *peer_ptr) {
    if (peer_ptr) FPV_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return FPV_OK;
}

/* Flag barrier across the ranks: flags_host[q] = device pointer to rank q's flag array ([world] uint32, own included,
 * all zero at start); epoch / error are local device words.  Enqueued on `stream`; no host synchronisation. */
int fpv_p2p_barrier(uint32_t *const *flags_host, int rank, int world, uint32_t *epoch, uint32_t *error,
                    double timeout_s, fpv_stream_t stream) {
    FPV_CHECK_ARG(flags_host && epoch && error, "fpv_p2p_barrier: null pointer");
    FPV_CHECK_ARG(world >= 1 && world <= FPV_MAX_PEERS + 1 && rank >= 0 && rank < world, "fpv_p2p_barrier: bad rank / world");
    BarrierParams p;
    for (int q = 0; q < FPV_MAX_PEERS + 1; ++q) p.flags[q] = q < world ? flags_host[q] : nullptr;
    p.epoch = epoch;
    p.error = error;
    p.rank = rank;
    p.world = world;
    p.timeout_ns = (unsigned long long)((timeout_s > 0 ? timeout_s : 20.0) * 1e9);
    p2p_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(p);
    FPV_LAUNCH_CHECK("p2p_barrier_kernel");
    return FPV_OK;
}

/* Combine: element-wise minimum over the `world` key slots ([world][slot_stride] uint64 per half) of the local mailbox,
 * unpacked into dist / idx (either may be NULL) and / or keys_out, written at perm[position in row] when perm is given
 * (row = queries per batch; perm is [row], or [n / row][row] when perm_batched).  parity (optional device word): bit 0 selects the half; flipped afterwards when flip != 0. */
int fpv_p2p_min_unpack(const uint64_t *slots, int world, int64_t slot_stride, int64_t half_stride, uint32_t *parity,
                       int flip, int64_t n, int64_t row, const long long *perm, int perm_batched, float *dist, void *idx,
                       int idx_bytes, uint64_t *keys_out, fpv_stream_t stream) {
    FPV_CHECK_ARG(slots && n > 0 && world >= 1 && row > 0, "fpv_p2p_min_unpack: bad arguments");
    FPV_CHECK_ARG(!idx || idx_bytes == 4 || idx_bytes == 8, "fpv_p2p_min_unpack: idx_bytes must be 4 or 8");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    p2p_min_unpack_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(
        reinterpret_cast<const unsigned long long *>(slots), world, slot_stride, half_stride, parity, n, row, perm,
        perm_batched ? row : 0, dist, idx, idx_bytes, reinterpret_cast<unsigned long long *>(keys_out));
    FPV_LAUNCH_CHECK("p2p_min_unpack_kernel");
    if (parity && flip) {
        p2p_flip_kernel<<<1, 1, 0, st>>>(parity);
        FPV_LAUNCH_CHECK("p2p_flip_kernel");
    }
    return FPV_OK;
}

/* src[n] floats -> the sender's slot in n_dst mailboxes (dst_host[r] = device pointer to that slot's first half). */
int fpv_p2p_push(const float *src, int64_t n, float *const *dst_host, int n_dst, const uint32_t *parity,
                 int64_t half_stride, fpv_stream_t stream) {
    FPV_CHECK_ARG(src && dst_host && n > 0 && n_dst >= 1 && n_dst <= FPV_MAX_PEERS + 1, "fpv_p2p_push: bad arguments");
    PushParams p;
    for (int r = 0; r < FPV_MAX_PEERS + 1; ++r) p.dst[r] = r < n_dst ? dst_host[r] : nullptr;
    p.n_dst = n_dst;
    p.parity = parity;
    p.half_stride = half_stride;
    p2p_push_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, n, p);
    FPV_LAUNCH_CHECK("p2p_push_kernel");
    return FPV_OK;
}

/* All-gather read-out: out[offset_host[r] .. + count_host[r]) = the piece rank r pushed into slot r of the local mailbox. */
int fpv_p2p_gather(const float *slots, int world, int64_t slot_stride, int64_t half_stride, uint32_t *parity, int flip,
                   const int64_t *offset_host, const int64_t *count_host, float *out, fpv_stream_t stream) {
    FPV_CHECK_ARG(slots && out && offset_host && count_host && world >= 1 && world <= FPV_MAX_PEERS + 1,
                  "fpv_p2p_gather: bad arguments");
    GatherParams g;
    int64_t mx = 1;
    for (int r = 0; r < FPV_MAX_PEERS + 1; ++r) {
        g.offset[r] = r < world ? offset_host[r] : 0;
        g.count[r] = r < world ? count_host[r] : 0;
        if (g.count[r] > mx) mx = g.count[r];
    }
    g.world = world;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int64_t bx = ceil_div(mx, 256 * 4);
    if (bx > 1024) bx = 1024;
    dim3 grid((unsigned)bx, (unsigned)world);
    p2p_gather_kernel<<<grid, 256, 0, st>>>(slots, slot_stride, half_stride, parity, g, out);
    FPV_LAUNCH_CHECK("p2p_gather_kernel");
    if (parity && flip) {
        p2p_flip_kernel<<<1, 1, 0, st>>>(parity);
        FPV_LAUNCH_CHECK("p2p_flip_kernel");
    }
    return FPV_OK;
}

/* out[n] = sum over the `world` float slots of the local mailbox, in rank order; flips parity afterwards when flip != 0. */
int fpv_p2p_sum(const float *slots, int world, int64_t slot_stride, int64_t half_stride, uint32_t *parity, int flip,
                int64_t n, float *out, fpv_stream_t stream) {
    FPV_CHECK_ARG(slots && out && n > 0 && world >= 1, "fpv_p2p_sum: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    p2p_sum_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(slots, world, slot_stride, half_stride, parity, n, out);
    FPV_LAUNCH_CHECK("p2p_sum_kernel");
    if (parity && flip) {
        p2p_flip_kernel<<<1, 1, 0, st>>>(parity);
        FPV_LAUNCH_CHECK("p2p_flip_kernel");
    }
    return FPV_OK;
}

}  // extern "C"
