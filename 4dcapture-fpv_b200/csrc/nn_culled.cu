// nn_culled.cu -- exact nearest neighbour over a spatially tiled candidate cloud with box culling.  sm_100a.
//
// SURVEY.md section 7 step 5 ("the exact uniform-grid/cell-sorted variant ... the only way to approach an
// HBM-type roofline"): brute force evaluates Q*M pairs; here the candidate cloud is pre-sorted along a
// Morton curve (the scene is static across all optimiser steps, global_optimization.py:175) and cut into
// tiles of 64 consecutive points with their axis-aligned bounding boxes.  Queries arrive in compact groups
// of 128 (also Morton-sorted by the caller).  One warp owns one group:
//   1. it computes the group's bounding box and, 32 tiles at a time, the box-to-box lower bound of the
//      squared distance to every tile; the tile with the smallest bound is searched first (seed);
//   2. it sweeps all tiles again and searches only those whose lower bound does not exceed the group's
//      current worst best-distance (re-evaluated after every searched tile).
// Inside a tile the arithmetic is the canonical fp32 expression of nn_search.cu, bit for bit.  Because
// tiles are visited out of index order, the winner is chosen by the full lexicographic rule
// (distance, ORIGINAL index): ties still resolve to the lowest original index, and a culled tile can
// never hold a winner because its lower bound (deflated by 2^-18 against fp32 rounding) exceeds every
// best distance of the group -- the result is identical to brute force.
//
// Two culling rules (template REP):
//   box mode (tiles of 64)  a tile is skipped when the box-to-box gap exceeds the GROUP's worst best-distance.
//                           Right for dense, static candidates queried from nearby (body -> scene: 1.4 % searched).
//   rep mode (tiles of 32)  every tile carries its first point ("representative") and the radius that covers the
//                           tile around it.  A query needs the tile only if |x - rep| <= sqrt(best_x) + radius --
//                           a PER-QUERY triangle-inequality test, far tighter than box gaps for far-field queries
//                           (scene -> body: ~13 % of the 32-vertex clusters survive, measured on the synthetic clip,
//                           where box culling keeps 77 %).  The representatives are real candidates, so a first pass
//                           over them alone seeds every query with a near-optimal best distance.
#include <math_constants.h>

#include <type_traits>

#include "common.cuh"

namespace fpv {

constexpr int CU_TILE = 64;     // candidates per tile, box mode
constexpr int CU_TILE_REP = 32; // candidates per tile, representative mode
constexpr int CU_SLOT = 6;      // floats per table entry
constexpr int CU_QPT = 4;       // queries per lane -> 128 queries per warp (group)
constexpr int CU_GROUP = 32 * CU_QPT;
constexpr int CU_WARPS = 4;     // groups per CTA

struct CulledParams {
    const float *q;       // [batches or 1][N][3] (grouped: 128 consecutive queries are spatially compact)
    int64_t q_bstride;    // floats between batches (0 = shared)
    int64_t N;
    const float *planes;  // [cand batches][3][Mp] Morton-sorted candidates (pad = +inf)
    int64_t plane_bstride, Mp, M;
    const float *boxes;   // [cand batches][ntile + nsuper][6]; tile entry = box (xmin ymin zmin xmax ymax zmax) or, in
    int64_t box_bstride;  //   rep mode, (x y z radius orig_idx -); a super box covers 32 tiles and follows the tile entries
    const int *oidx;      // [cand batches][M] original index of every sorted candidate
    int64_t oidx_bstride;
    int ntile, nsuper;
    int64_t idx_base;
    float *dist;
    void *idx;
    int idx_bytes;
    unsigned long long *tiles_searched;  // optional statistics counter
    const float *cand_orig;  // [cand batches][M][3] candidates in ORIGINAL order (needed to re-evaluate seeds), may be null
    int64_t cand_orig_bstride;
    int *seed;               // [batches][N] in/out winners of the previous call (original indices), may be null
    int seed_read;
    // packed-key epilogue (multi-GPU combine): when `keys` is set the kernel writes
    // (float_bits(d) << 32 | idx_base + idx) per query instead of dist / idx, and stores the same key into every
    // peer's mailbox slot (push_n remote buffers of the same layout, reached over NVLink) -- the transfer overlaps
    // the search warp by warp.  push_parity (device, may be null) selects the half of a double-buffered mailbox.
    unsigned long long *keys;
    unsigned long long *push[FPV_MAX_PEERS];
    int push_n;
    const unsigned *push_parity;
    int64_t push_half;       // elements between the two halves of a mailbox slot
};

__device__ __forceinline__ float cu_d2(float x, float y, float z, float rx, float ry, float rz) {
    const float dx = __fsub_rn(x, rx), dy = __fsub_rn(y, ry), dz = __fsub_rn(z, rz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// squared box-to-box gap, deflated so that fp32 rounding can never make it exceed a true pair distance
__device__ __forceinline__ float box_lb(const float *__restrict__ b, const float (&gmin)[3], const float (&gmax)[3]) {
    float s = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float g = fmaxf(0.f, fmaxf(b[a] - gmax[a], gmin[a] - b[3 + a]));
        s = fmaf(g, g, s);
    }
    return s * (1.0f - 1.0f / 262144.0f);
}

// Search one tile (TILE candidates staged in this warp's shared buffer) for the warp's 128 queries.
template <int TILE>
__device__ __forceinline__ void cu_search_tile(const float *sx, const float *sy, const float *sz,
                                               const int *__restrict__ oidx_tile, const float (&qx)[CU_QPT],
                                               const float (&qy)[CU_QPT], const float (&qz)[CU_QPT],
                                               float (&best)[CU_QPT], int (&bidx)[CU_QPT]) {
    const float4 *X = reinterpret_cast<const float4 *>(sx);
    const float4 *Y = reinterpret_cast<const float4 *>(sy);
    const float4 *Z = reinterpret_cast<const float4 *>(sz);
#pragma unroll 2
    for (int j4 = 0; j4 < TILE / 4; ++j4) {
        const float4 rx = X[j4], ry = Y[j4], rz = Z[j4];
        const float2 nx0 = make_float2(-rx.x, -rx.y), nx1 = make_float2(-rx.z, -rx.w);
        const float2 ny0 = make_float2(-ry.x, -ry.y), ny1 = make_float2(-ry.z, -ry.w);
        const float2 nz0 = make_float2(-rz.x, -rz.y), nz1 = make_float2(-rz.z, -rz.w);
        float m4[CU_QPT];
        bool any = false;
#pragma unroll
        for (int q = 0; q < CU_QPT; ++q) {
            const float2 bx = make_float2(qx[q], qx[q]), by = make_float2(qy[q], qy[q]), bz = make_float2(qz[q], qz[q]);
            const float2 dx0 = __fadd2_rn(bx, nx0), dx1 = __fadd2_rn(bx, nx1);
            const float2 dy0 = __fadd2_rn(by, ny0), dy1 = __fadd2_rn(by, ny1);
            const float2 dz0 = __fadd2_rn(bz, nz0), dz1 = __fadd2_rn(bz, nz1);
            float2 s0 = __fmul2_rn(dx0, dx0), s1 = __fmul2_rn(dx1, dx1);
            s0 = __ffma2_rn(dy0, dy0, s0);
            s1 = __ffma2_rn(dy1, dy1, s1);
            s0 = __ffma2_rn(dz0, dz0, s0);
            s1 = __ffma2_rn(dz1, dz1, s1);
            m4[q] = fminf(fmin3(s0.x, s0.y, s1.x), s1.y);
            any |= (m4[q] <= best[q]);  // '<=': an equal distance with a lower original index must win
        }
        if (any) {
#pragma unroll
            for (int q = 0; q < CU_QPT; ++q) {
                if (m4[q] <= best[q]) {
                    const float e[4] = {cu_d2(qx[q], qy[q], qz[q], rx.x, ry.x, rz.x), cu_d2(qx[q], qy[q], qz[q], rx.y, ry.y, rz.y),
                                        cu_d2(qx[q], qy[q], qz[q], rx.z, ry.z, rz.z), cu_d2(qx[q], qy[q], qz[q], rx.w, ry.w, rz.w)};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (e[c] <= best[q]) {
                            const int o = __ldg(oidx_tile + 4 * j4 + c);
                            if (e[c] < best[q] || o < bidx[q]) {
                                best[q] = e[c];
                                bidx[q] = o;
                            }
                        }
                    }
                }
            }
        }
    }
}

template <int TILE, bool REP>
__global__ void __launch_bounds__(CU_WARPS * 32) nn_culled_kernel(const CulledParams p) {
    __shared__ __align__(16) float stile[CU_WARPS][3][TILE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t b = blockIdx.y;
    const int64_t group = int64_t(blockIdx.x) * CU_WARPS + warp;
    const int64_t q0 = group * CU_GROUP;
    if (q0 >= p.N) return;  // whole warp leaves together
    const float *qsrc = p.q + b * p.q_bstride;
    const float *planes = p.planes + b * p.plane_bstride;
    const float *boxes = p.boxes + b * p.box_bstride;
    const int *oidx = p.oidx + b * p.oidx_bstride;
    float *sx = stile[warp][0], *sy = stile[warp][1], *sz = stile[warp][2];

    float qx[CU_QPT], qy[CU_QPT], qz[CU_QPT], best[CU_QPT];
    int bidx[CU_QPT];
    float gmin[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, gmax[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
    for (int k = 0; k < CU_QPT; ++k) {
        int64_t qi = q0 + k * 32 + lane;
        if (qi > p.N - 1) qi = p.N - 1;
        qx[k] = __ldg(qsrc + 3 * qi);
        qy[k] = __ldg(qsrc + 3 * qi + 1);
        qz[k] = __ldg(qsrc + 3 * qi + 2);
        best[k] = CUDART_INF_F;
        bidx[k] = 0;
        gmin[0] = fminf(gmin[0], qx[k]); gmax[0] = fmaxf(gmax[0], qx[k]);
        gmin[1] = fminf(gmin[1], qy[k]); gmax[1] = fmaxf(gmax[1], qy[k]);
        gmin[2] = fminf(gmin[2], qz[k]); gmax[2] = fmaxf(gmax[2], qz[k]);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        gmin[a] = warp_min(gmin[a]);
        gmax[a] = warp_max(gmax[a]);
    }
    // a NaN query coordinate poisons the box: fminf/fmaxf drop NaNs, so the box only covers the finite
    // queries; a NaN query never wins anything anyway (all its distances are NaN) -> result (+inf, 0).

    auto stage_and_search = [&](int t) {
        const int64_t j0 = int64_t(t) * TILE;
        __syncwarp();
#pragma unroll
        for (int c = 0; c < TILE; c += 32) {
            sx[c + lane] = planes[j0 + c + lane];
            sy[c + lane] = planes[p.Mp + j0 + c + lane];
            sz[c + lane] = planes[2 * p.Mp + j0 + c + lane];
        }
        __syncwarp();
        cu_search_tile<TILE>(sx, sy, sz, oidx + j0, qx, qy, qz, best, bidx);
    };
    auto group_worst = [&]() {
        float w = fmaxf(fmaxf(best[0], best[1]), fmaxf(best[2], best[3]));
        return warp_max(w);
    };

    const float *sboxes = boxes + int64_t(p.ntile) * CU_SLOT;
    unsigned long long searched = 0;
    if (REP) {
        // ---- phase A: the representatives alone (real candidates) seed every query's best distance ----
        for (int t = 0; t < p.ntile; ++t) {
            const float *e = boxes + int64_t(t) * CU_SLOT;
            const float rx = __ldg(e), ry = __ldg(e + 1), rz = __ldg(e + 2);
            const int o = __float_as_int(__ldg(e + 4));
#pragma unroll
            for (int k = 0; k < CU_QPT; ++k) {
                const float d = cu_d2(qx[k], qy[k], qz[k], rx, ry, rz);
                if (d <= best[k] && (d < best[k] || o < bidx[k])) {
                    best[k] = d;
                    bidx[k] = o;
                }
            }
        }
        // ---- phase B: a query needs a tile only if |x - rep| <= sqrt(best_x) + radius ----
        float sq[CU_QPT];
#pragma unroll
        for (int k = 0; k < CU_QPT; ++k) sq[k] = sqrtf(best[k]) * 1.000001f;
        float worst = group_worst();
        for (int s0 = 0; s0 < p.nsuper; s0 += 32) {
            const int sidx = s0 + lane;
            float slb = CUDART_INF_F;
            if (sidx < p.nsuper) slb = box_lb(sboxes + int64_t(sidx) * CU_SLOT, gmin, gmax);
            // (a lane past the table must never vote: with a query that has no finite best -- NaN / infinite / huge
            // coordinates -- `worst` is +inf and INF <= INF would send the warp into tiles that do not exist)
            unsigned smask = __ballot_sync(0xffffffffu, sidx < p.nsuper && slb <= worst);
            while (smask) {
                const int sl = __ffs(smask) - 1;
                smask &= smask - 1;
                if (!(__shfl_sync(0xffffffffu, slb, sl) <= worst)) continue;
                const int t0 = (s0 + sl) * 32;
                const int t1 = (t0 + 32 < p.ntile) ? t0 + 32 : p.ntile;
                for (int t = t0; t < t1; ++t) {
                    const float *e = boxes + int64_t(t) * CU_SLOT;
                    const float rx = __ldg(e), ry = __ldg(e + 1), rz = __ldg(e + 2), rad = __ldg(e + 3);
                    bool need = false;
#pragma unroll
                    for (int k = 0; k < CU_QPT; ++k) {
                        const float d = cu_d2(qx[k], qy[k], qz[k], rx, ry, rz);
                        const float lim = sq[k] + rad;
                        need |= (d <= lim * lim * 1.000001f);  // NaN queries never ask for a tile
                    }
                    if (__ballot_sync(0xffffffffu, need)) {
                        stage_and_search(t);
                        ++searched;
#pragma unroll
                        for (int k = 0; k < CU_QPT; ++k) sq[k] = sqrtf(best[k]) * 1.000001f;
                        worst = group_worst();
                    }
                }
            }
        }
    } else {
    // ---- phase A: seed = the tile with the smallest lower bound inside the super-tile with the smallest ----
    int seed;
    {
        float lmin = CUDART_INF_F;
        int smin = 0;
        for (int s0 = 0; s0 < p.nsuper; s0 += 32) {
            const int sidx = s0 + lane;
            if (sidx < p.nsuper) {
                const float lb = box_lb(sboxes + int64_t(sidx) * CU_SLOT, gmin, gmax);
                if (lb < lmin) {
                    lmin = lb;
                    smin = sidx;
                }
            }
        }
        float wmin = warp_min(lmin);
        unsigned m = __ballot_sync(0xffffffffu, lmin == wmin);
        smin = __shfl_sync(0xffffffffu, smin, m ? __ffs(m) - 1 : 0);
        const int t = smin * 32 + lane;
        const float lb = (t < p.ntile) ? box_lb(boxes + int64_t(t) * CU_SLOT, gmin, gmax) : CUDART_INF_F;
        wmin = warp_min(lb);
        m = __ballot_sync(0xffffffffu, lb == wmin);
        seed = smin * 32 + (m ? __ffs(m) - 1 : 0);
        if (seed >= p.ntile) seed = 0;
    }
    // seeds carried from the previous call on the same problem: every query starts from the exact distance to its old
    // winner, the group bound is tight before the first tile and the nearest-tile probe is not needed
    bool seeded = false;
    if (p.seed_read) {
        int sd[CU_QPT];
        bool ok = true;
#pragma unroll
        for (int k = 0; k < CU_QPT; ++k) {
            int64_t qi = q0 + k * 32 + lane;
            if (qi > p.N - 1) qi = p.N - 1;
            sd[k] = p.seed[b * p.N + qi];
            ok &= unsigned(sd[k]) < unsigned(p.M);
        }
        seeded = __all_sync(0xffffffffu, ok);
        if (seeded) {
            const float *co = p.cand_orig + b * p.cand_orig_bstride;
#pragma unroll
            for (int k = 0; k < CU_QPT; ++k) {
                const float *y = co + 3 * int64_t(sd[k]);
                best[k] = cu_d2(qx[k], qy[k], qz[k], __ldg(y), __ldg(y + 1), __ldg(y + 2));
                bidx[k] = sd[k];
                if (!(best[k] < CUDART_INF_F)) {
                    best[k] = CUDART_INF_F;
                    bidx[k] = 0;
                }
            }
            seed = -1;
        }
    }
    if (!seeded) {
        stage_and_search(seed);
        searched = 1;
    }
    float worst = group_worst();

    // ---- phase B: two-level sweep, searching only tiles that can still hold a winner ----
    for (int s0 = 0; s0 < p.nsuper; s0 += 32) {
        const int sidx = s0 + lane;
        float slb = CUDART_INF_F;
        if (sidx < p.nsuper) slb = box_lb(sboxes + int64_t(sidx) * CU_SLOT, gmin, gmax);
        // a lane past the table must never vote (worst can be +inf: INF <= INF), see the rep-mode sweep above
        unsigned smask = __ballot_sync(0xffffffffu, sidx < p.nsuper && slb <= worst);
        while (smask) {
            const int sl = __ffs(smask) - 1;
            smask &= smask - 1;
            if (!(__shfl_sync(0xffffffffu, slb, sl) <= worst)) continue;
            const int t0 = (s0 + sl) * 32;
            const int t = t0 + lane;
            const bool tv = t < p.ntile && t != seed;
            float lb = CUDART_INF_F;
            if (tv) lb = box_lb(boxes + int64_t(t) * CU_SLOT, gmin, gmax);
            unsigned mask = __ballot_sync(0xffffffffu, tv && lb <= worst);
            while (mask) {
                const int l = __ffs(mask) - 1;
                mask &= mask - 1;
                if (__shfl_sync(0xffffffffu, lb, l) <= worst) {  // the group's worst distance may have shrunk
                    // per-query refinement: point-to-box gap (deflated like box_lb) against the query's OWN best
                    const float *e = boxes + int64_t(t0 + l) * CU_SLOT;
                    const float lx = __ldg(e), ly = __ldg(e + 1), lz = __ldg(e + 2);
                    const float hx = __ldg(e + 3), hy = __ldg(e + 4), hz = __ldg(e + 5);
                    bool need = false;
#pragma unroll
                    for (int k = 0; k < CU_QPT; ++k) {
                        const float gx = fmaxf(fmaxf(lx - qx[k], qx[k] - hx), 0.f);
                        const float gy = fmaxf(fmaxf(ly - qy[k], qy[k] - hy), 0.f);
                        const float gz = fmaxf(fmaxf(lz - qz[k], qz[k] - hz), 0.f);
                        const float g2 = fmaf(gz, gz, fmaf(gy, gy, gx * gx)) * (1.0f - 1.0f / 262144.0f);
                        need |= g2 <= best[k];  // a NaN query never asks
                    }
                    if (!__ballot_sync(0xffffffffu, need)) continue;
                    stage_and_search(t0 + l);
                    worst = group_worst();
                    ++searched;
                }
            }
        }
    }
    }
    if (p.tiles_searched && lane == 0) atomicAdd(p.tiles_searched, searched);

    if (p.keys != nullptr) {
        const int64_t half = (p.push_parity != nullptr && (__ldg(p.push_parity) & 1u)) ? p.push_half : 0;
#pragma unroll
        for (int k = 0; k < CU_QPT; ++k) {
            const int64_t qi = q0 + k * 32 + lane;
            if (qi < p.N) {
                const int64_t o = b * p.N + qi;
                const unsigned long long key =
                    (static_cast<unsigned long long>(__float_as_uint(best[k])) << 32) | unsigned(p.idx_base + bidx[k]);
                p.keys[o + half] = key;
                for (int r = 0; r < p.push_n; ++r) p.push[r][o + half] = key;  // coalesced 8-byte stores into peer memory
                if (p.seed != nullptr) p.seed[o] = bidx[k];
            }
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < CU_QPT; ++k) {
        const int64_t qi = q0 + k * 32 + lane;
        if (qi < p.N) {
            const int64_t o = b * p.N + qi;
            const int64_t gi = p.idx_base + bidx[k];
            p.dist[o] = best[k];
            if (p.idx_bytes == 8)
                static_cast<long long *>(p.idx)[o] = gi;
            else
                static_cast<int *>(p.idx)[o] = int(gi);
            if (p.seed != nullptr) p.seed[o] = bidx[k];
        }
    }
}

// Per-tile table of Morton-sorted candidate planes (pad entries are +inf and are skipped).
//   box mode: bounding box.   rep mode: first point, covering radius around it (inflated), its original index.
template <int TILE, bool REP>
__global__ void tile_table_kernel(const float *__restrict__ planes, const int *__restrict__ oidx, int64_t M, int64_t Mp,
                                  int ntile, int nsuper, float *__restrict__ boxes) {
    const int64_t b = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntile) return;
    const float *P = planes + b * 3 * Mp;
    float *o = boxes + (b * int64_t(ntile + nsuper) + t) * CU_SLOT;
    const int64_t j0 = int64_t(t) * TILE;
    if (REP) {
        const float rx = P[j0], ry = P[Mp + j0], rz = P[2 * Mp + j0];
        float r2 = 0.f;
        for (int c = 1; c < TILE && j0 + c < M; ++c) {
            const float dx = P[j0 + c] - rx, dy = P[Mp + j0 + c] - ry, dz = P[2 * Mp + j0 + c] - rz;
            const float d = dx * dx + dy * dy + dz * dz;
            if (d == d) r2 = fmaxf(r2, d);  // a candidate with a NaN coordinate can never win: ignore it
        }
        o[0] = rx;
        o[1] = ry;
        o[2] = rz;
        o[3] = sqrtf(r2) * 1.00001f + 1e-30f;  // covering radius (an infinite member makes it +inf: always searched)
        o[4] = __int_as_float(oidx[b * Mp + j0]);
        o[5] = 0.f;
    } else {
        float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
        for (int c = 0; c < TILE && j0 + c < M; ++c) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float v = P[a * Mp + j0 + c];
                lo[a] = fminf(lo[a], v);  // NaNs are dropped; a candidate with a NaN coordinate can never win
                hi[a] = fmaxf(hi[a], v);
            }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            o[a] = lo[a];
            o[3 + a] = hi[a];
        }
    }
}

// Super boxes: bounding box of the points of 32 consecutive tiles.
template <int TILE>
__global__ void super_boxes_kernel(const float *__restrict__ planes, int64_t M, int64_t Mp, int ntile, int nsuper,
                                   float *__restrict__ boxes) {
    const int64_t b = blockIdx.y;
    const int sidx = blockIdx.x;
    if (sidx >= nsuper) return;
    const float *P = planes + b * 3 * Mp;
    float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    const int64_t j0 = int64_t(sidx) * 32 * TILE;
    for (int64_t j = j0 + threadIdx.x; j < j0 + 32 * TILE && j < M; j += 32) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = P[a * Mp + j];
            lo[a] = fminf(lo[a], v);
            hi[a] = fmaxf(hi[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = warp_min(lo[a]);
        hi[a] = warp_max(hi[a]);
    }
    if (threadIdx.x == 0) {
        float *o = boxes + (b * int64_t(ntile + nsuper) + ntile + sidx) * CU_SLOT;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            o[a] = lo[a];
            o[3 + a] = hi[a];
        }
    }
}


// =====================================================================================================
// Sphere-hierarchy mode (scene -> body): candidates are per-frame Morton-sorted body vertices cut into
// clusters of TILE points with a 4-level hierarchy of bounding spheres (TILE, 4*TILE, 16*TILE, 64*TILE points).
// A query needs a cluster only if |x - c| <= sqrt(best_x) + r  (triangle inequality, PER QUERY; the warp
// searches a cluster when any of its 128 spatially adjacent queries needs it).  When the query set is shared
// by consecutive frames of a clip, a warp walks a chunk of frames and seeds every query with the exact distance
// to its previous frame's winner (the body moves centimetres per frame), so only the first frame of a chunk pays
// for a seeding pass.  Results are identical to brute force (lexicographic rule on ORIGINAL indices).
//
// Both inner loops run on EXPANDED forms with a conservative slack, and fall back to the canonical
// (x-y)^2 evaluation only where the expanded form cannot rule a candidate out:
//   sphere test : |x|^2 - sq^2 - 2 x.c - 2 sq r  <=  r^2 - |c|^2           (4 packed FMAs per two queries)
//   tile search : |y|^2 - 2 x.y                  <=  best - |x|^2          (3 packed FMAs per two candidates)
// with sq = sqrt(best) inflated.  Every fp32 evaluation error of these forms is bounded by 2^-19.7 (|x|^2 + |c|^2
// + ...); the thresholds carry 2^-18 of the same magnitudes, so a cluster / candidate that the canonical
// arithmetic would accept is never dropped.  Candidates that pass the filter are re-evaluated canonically and
// compared lexicographically, so distances and indices stay bit-identical to brute force.  A warp holding a
// query with |x|^2 > 1e30 (where the expanded forms overflow) searches every cluster canonically instead.
// =====================================================================================================
constexpr float SPH_SLACK = 1.0f / 262144.0f;  // 2^-18

struct SphereParams {
    const float *q;
    int64_t q_bstride, N;
    const float *planes;
    int64_t plane_bstride, Mp, M;
    const float4 *table;    // [cand batches] { [n0p + n1p + n2p] sphere entries of two float4:
    int64_t table_bstride;  //   (-2cx, -2cy, -2cz, -2r) and (r^2 - |c|^2 + slack, (|c| + r)^2, -, -), every level padded to
                            //   whole sibling groups of 4 with never-asked entries (lim = NaN);
                            //   [n0p * TILE / 4] expanded candidate groups of four float4: -2x, -2y, -2z, |y|^2 of 4 points }
                            // stride in float4
    const int *oidx;
    int64_t oidx_bstride;
    const int *pos_of;       // [cand batches or 1][M] sorted position of every ORIGINAL candidate index (seeding), may be null
    int64_t pos_bstride;     //   0 when one ordering serves every batch
    int *seed;               // [batches][N] in/out: last call's winners (original indices; < 0 = none), may be null
    int seed_read;           // 0: the buffer holds nothing yet, only write it
    int n0, n1, n2, n3;
    int batches, frames_per_cta;
    int64_t idx_base;
    float *dist;
    void *idx;
    int idx_bytes;
    unsigned long long *tiles_searched;
    // fused-loss mode (MODE 2): nothing of size [batches][N] is written except the winners (into the seed buffer).  Per
    // batch the kernel accumulates the winners' distances (per-warp partial sums in double, reduced in fixed order
    // afterwards); s2b_accum_kernel then builds, per winning candidate, the number of queries it won and the fixed-point
    // sum of their coordinates -- everything the backward of  sum_j d(x_j, y_win(j))  needs:  d/dy_i = 2 (n_i y_i - S_i).
    double *sum_partial;          // [batches][groups]
    int64_t groups;               // query groups (warps) per batch
    unsigned long long *acc;      // [batches][M][4]: S_x, S_y, S_z (x * 2^fix_shift, two's complement), count
    float fix_scale;              // 2^fix_shift
};

// does any of this lane's 4 queries (two packed pairs) need the sphere?  a2 = |x|^2(1-s) - sq^2(1+s), sq2 = sq
__device__ __forceinline__ bool sphere_needed(const float4 e, const float lim, const float2 (&qx2)[2],
                                              const float2 (&qy2)[2], const float2 (&qz2)[2], const float2 (&sq2)[2],
                                              const float2 (&a2)[2]) {
    const float2 mx = make_float2(e.x, e.x), my = make_float2(e.y, e.y), mz = make_float2(e.z, e.z);
    const float2 mr = make_float2(e.w, e.w);
    float2 t[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        t[h] = __ffma2_rn(qx2[h], mx, a2[h]);
        t[h] = __ffma2_rn(qy2[h], my, t[h]);
        t[h] = __ffma2_rn(qz2[h], mz, t[h]);
        t[h] = __ffma2_rn(sq2[h], mr, t[h]);
    }
    // one comparison for the four queries: min drops NaNs (a NaN / infinite query never asks; a query next to it still does)
    return fminf(fmin3(t[0].x, t[0].y, t[1].x), t[1].y) <= lim;
}

// Search one cluster, read as expanded groups (-2y, |y|^2) straight from L1, for the warp's 128 queries; returns true when a query of this lane
// improved.  thr[q] = best[q] - |x_q|^2 (1 - s) + s (|c| + r)^2.
template <int TILE>
__device__ __forceinline__ bool sphere_search_tile(const float4 *__restrict__ xp_tile, const SphereParams &p, int b, int j0,
                                                   const float (&qx)[CU_QPT],
                                                   const float (&qy)[CU_QPT], const float (&qz)[CU_QPT],
                                                   const float (&thr)[CU_QPT], const int (&sgrp)[CU_QPT],
                                                   float (&best)[CU_QPT], int (&bidx)[CU_QPT]) {
    bool improved = false;
    // a query's SEED group (the four sorted candidates around its seed winner) was evaluated canonically when the seed
    // was taken, so its filter hits there -- the seed winner always passes its own threshold -- carry no news: masked
    int rel[CU_QPT];
#pragma unroll
    for (int q = 0; q < CU_QPT; ++q) rel[q] = sgrp[q] - (j0 >> 2);
#pragma unroll
    for (int j4 = 0; j4 < TILE / 4; ++j4) {
        // warp-uniform addresses: every load is one broadcast request served by L1
        const float4 mx = __ldg(xp_tile + 4 * j4), my = __ldg(xp_tile + 4 * j4 + 1), mz = __ldg(xp_tile + 4 * j4 + 2),
                     y2 = __ldg(xp_tile + 4 * j4 + 3);
        const float2 mx0 = make_float2(mx.x, mx.y), mx1 = make_float2(mx.z, mx.w);
        const float2 my0 = make_float2(my.x, my.y), my1 = make_float2(my.z, my.w);
        const float2 mz0 = make_float2(mz.x, mz.y), mz1 = make_float2(mz.z, mz.w);
        const float2 w0 = make_float2(y2.x, y2.y), w1 = make_float2(y2.z, y2.w);
        bool hit[CU_QPT];
        bool any = false;
#pragma unroll
        for (int q = 0; q < CU_QPT; ++q) {
            const float2 bx = make_float2(qx[q], qx[q]), by = make_float2(qy[q], qy[q]), bz = make_float2(qz[q], qz[q]);
            float2 s0 = __ffma2_rn(bx, mx0, w0), s1 = __ffma2_rn(bx, mx1, w1);
            s0 = __ffma2_rn(by, my0, s0);
            s1 = __ffma2_rn(by, my1, s1);
            s0 = __ffma2_rn(bz, mz0, s0);
            s1 = __ffma2_rn(bz, mz1, s1);
            hit[q] = (fminf(fmin3(s0.x, s0.y, s1.x), s1.y) <= thr[q]) && rel[q] != j4;
            any |= hit[q];
        }
        if (any) {  // rare: canonical re-evaluation of the four candidates, lexicographic update
            // (the plane / index addresses are formed here, not on the hot path)
            const float *planes_tile = p.planes + int64_t(b) * p.plane_bstride + j0;
            const int *oidx_tile = p.oidx + int64_t(b) * p.oidx_bstride + j0;
            const int64_t Mp = p.Mp;
            float rx[4], ry[4], rz[4];
            int o[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                rx[c] = __ldg(planes_tile + 4 * j4 + c);
                ry[c] = __ldg(planes_tile + Mp + 4 * j4 + c);
                rz[c] = __ldg(planes_tile + 2 * Mp + 4 * j4 + c);
                o[c] = __ldg(oidx_tile + 4 * j4 + c);
            }
#pragma unroll
            for (int q = 0; q < CU_QPT; ++q) {
                if (hit[q]) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float e = cu_d2(qx[q], qy[q], qz[q], rx[c], ry[c], rz[c]);
                        if (e <= best[q] && (e < best[q] || o[c] < bidx[q])) {
                            best[q] = e;
                            bidx[q] = o[c];
                            improved = true;
                        }
                    }
                }
            }
        }
    }
    return improved;
}

// MODE 0: distances + indices out.  MODE 2: only the winners (into the in/out seed buffer, -1 for a query without a
// finite winner) and the per-warp distance sums; a separate streaming kernel (s2b_accum_kernel) then builds the
// accumulators of the fused loss from the winners (accumulating inside this kernel was measured: +5.3 ms).
template <int TILE, int MINB, int MODE>
__global__ void __launch_bounds__(CU_WARPS * 32, MINB) nn_sphere_kernel(const SphereParams p) {
    constexpr int ST = TILE < 32 ? 32 : TILE;
    __shared__ __align__(16) float stile[CU_WARPS][3][ST];  // canonical fallback path only
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t group = int64_t(blockIdx.x) * CU_WARPS + warp;
    const int64_t q0 = group * CU_GROUP;
    if (q0 >= p.N) return;
    const int b0 = blockIdx.y * p.frames_per_cta;
    const int b1 = (b0 + p.frames_per_cta < p.batches) ? b0 + p.frames_per_cta : p.batches;
    float *smx = stile[warp][0], *smy = stile[warp][1], *smz = stile[warp][2];
    float qx[CU_QPT], qy[CU_QPT], qz[CU_QPT], best[CU_QPT];
    int bidx[CU_QPT];
    bool canonical = false;  // warp-uniform: a query too large for the expanded forms
    unsigned long long searched = 0;

    for (int b = b0; b < b1; ++b) {
        if (b == b0 || p.q_bstride != 0) {
            const float *qsrc = p.q + int64_t(b) * p.q_bstride;
            bool big = false;
#pragma unroll
            for (int k = 0; k < CU_QPT; ++k) {
                int64_t qi = q0 + k * 32 + lane;
                if (qi > p.N - 1) qi = p.N - 1;
                qx[k] = __ldg(qsrc + 3 * qi);
                qy[k] = __ldg(qsrc + 3 * qi + 1);
                qz[k] = __ldg(qsrc + 3 * qi + 2);
                const float x2 = fmaf(qz[k], qz[k], fmaf(qy[k], qy[k], qx[k] * qx[k]));
                big |= x2 > 1e30f;
            }
            canonical = __ballot_sync(0xffffffffu, big) != 0;
        }
        const float *planes = p.planes + int64_t(b) * p.plane_bstride;
        const float4 *tab = p.table + int64_t(b) * p.table_bstride;
        const int *oidx = p.oidx + int64_t(b) * p.oidx_bstride;
        // seed, in order of preference: the winner of the previous CALL for this (frame, query) (an optimiser moves the
        // body by millimetres per step), the winner of the previous FRAME of this call, a coarse pass over the frame.
        // A seed is taken together with the three sorted candidates that share its group of four: all four are evaluated
        // canonically here, and the traversal then ignores the query's filter hits in that group (sgrp).
        int sgrp[CU_QPT] = {-1, -1, -1, -1};
        auto seed_from = [&](const int (&sd)[CU_QPT]) {
            const int *pos = p.pos_of + int64_t(b) * p.pos_bstride;
            const float4 *px4 = reinterpret_cast<const float4 *>(planes);
            const float4 *py4 = reinterpret_cast<const float4 *>(planes + p.Mp);
            const float4 *pz4 = reinterpret_cast<const float4 *>(planes + 2 * p.Mp);
            const int4 *o4 = reinterpret_cast<const int4 *>(oidx);
#pragma unroll
            for (int k = 0; k < CU_QPT; ++k) {
                const int g = __ldg(pos + sd[k]) >> 2;
                const float4 cx = __ldg(px4 + g), cy = __ldg(py4 + g), cz = __ldg(pz4 + g);
                const int4 co = __ldg(o4 + g);
                const float ex[4] = {cx.x, cx.y, cx.z, cx.w}, ey[4] = {cy.x, cy.y, cy.z, cy.w}, ez[4] = {cz.x, cz.y, cz.z, cz.w};
                const int eo[4] = {co.x, co.y, co.z, co.w};
                best[k] = CUDART_INF_F;
                bidx[k] = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float e = cu_d2(qx[k], qy[k], qz[k], ex[c], ey[c], ez[c]);
                    if (e <= best[k] && (e < best[k] || eo[c] < bidx[k])) {  // (+inf padding / NaN never win)
                        best[k] = e;
                        bidx[k] = eo[c];
                    }
                }
                sgrp[k] = g;
            }
        };
        bool seeded = false;
        if (p.seed_read) {
            int sd[CU_QPT];
            bool ok = true;
#pragma unroll
            for (int k = 0; k < CU_QPT; ++k) {
                int64_t qi = q0 + k * 32 + lane;
                if (qi > p.N - 1) qi = p.N - 1;
                sd[k] = p.seed[int64_t(b) * p.N + qi];
                ok &= unsigned(sd[k]) < unsigned(p.M);
            }
            seeded = __all_sync(0xffffffffu, ok);
            if (seeded) seed_from(sd);
        }
        const bool temporal = (b > b0) && p.q_bstride == 0 && p.pos_of != nullptr;
        if (seeded) {
        } else if (temporal) {
            // seed: the previous frame's winner (and its group), re-evaluated on this frame's vertices
            int sd[CU_QPT];
#pragma unroll
            for (int k = 0; k < CU_QPT; ++k) sd[k] = (best[k] < CUDART_INF_F) ? bidx[k] : 0;
            seed_from(sd);
        } else {
            // seed: the first point of every level-2 group (real candidates)
#pragma unroll
            for (int k = 0; k < CU_QPT; ++k) {
                best[k] = CUDART_INF_F;
                bidx[k] = 0;
            }
            for (int m = 0; m < p.n2; ++m) {
                const int64_t j = int64_t(m) * 16 * TILE;
                const float rx = __ldg(planes + j), ry = __ldg(planes + p.Mp + j), rz = __ldg(planes + 2 * p.Mp + j);
                const int o = __ldg(oidx + j);
#pragma unroll
                for (int k = 0; k < CU_QPT; ++k) {
                    const float d = cu_d2(qx[k], qy[k], qz[k], rx, ry, rz);
                    if (d <= best[k] && (d < best[k] || o < bidx[k])) {
                        best[k] = d;
                        bidx[k] = o;
                    }
                }
            }
        }
        const float2 qx2[2] = {make_float2(qx[0], qx[1]), make_float2(qx[2], qx[3])};
        const float2 qy2[2] = {make_float2(qy[0], qy[1]), make_float2(qy[2], qy[3])};
        const float2 qz2[2] = {make_float2(qz[0], qz[1]), make_float2(qz[2], qz[3])};
        float2 sq2[2], a2[2];
        float thrb[CU_QPT];
        auto refresh = [&]() {
            float sq[CU_QPT], a[CU_QPT];
#pragma unroll
            for (int k = 0; k < CU_QPT; ++k) {
                const float xs = fmaf(qz[k], qz[k], fmaf(qy[k], qy[k], qx[k] * qx[k])) * (1.0f - SPH_SLACK);
                sq[k] = sqrtf(best[k]) * 1.000001f;
                a[k] = fmaf(-sq[k] * sq[k], 1.0f + SPH_SLACK, xs);
                thrb[k] = best[k] - xs;
            }
            sq2[0] = make_float2(sq[0], sq[1]);
            sq2[1] = make_float2(sq[2], sq[3]);
            a2[0] = make_float2(a[0], a[1]);
            a2[1] = make_float2(a[2], a[3]);
        };
        refresh();
        // entry index space: level 0 at [0, n0p), level 1 behind it, then level 2, then level 3 (TILE*64 points per
        // sphere); n0p = 4 n1, n1p = 4 n2, n2p = 4 n3: every sphere's four children are one aligned sibling group
        const int n0p = 4 * p.n1, n1p = 4 * p.n2, n2p = 4 * p.n3, n3p = (p.n3 + 3) & ~3;
        const float4 *xp = tab + 2 * (n0p + n1p + n2p + n3p);
        // The traversal is instantiated twice: the normal path carries no trace of the canonical fallback (no flag to
        // keep in a register or reload on the critical path of every test).
        auto traverse = [&](auto canon_tag) {
            constexpr bool CANON = decltype(canon_tag)::value;
            // four sibling spheres at a time (one 128-byte line, independent FMA chains); bit i is warp-uniform
            auto test4 = [&](int first) -> unsigned {
                const float4 *g = tab + 2 * first;
                float4 e[4];
                float lim[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    e[i] = __ldg(g + 2 * i);
                    lim[i] = __ldg(reinterpret_cast<const float *>(g + 2 * i + 1));
                }
                unsigned mine = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    // canonical mode asks for everything that is not padding (lim != NaN)
                    const bool need = CANON ? (lim[i] == lim[i]) : sphere_needed(e[i], lim[i], qx2, qy2, qz2, sq2, a2);
                    if (need) mine |= 1u << i;
                }
                return __reduce_or_sync(0xffffffffu, mine);   // one warp-wide OR instead of four votes
            };
            for (int w0 = 0; w0 < n3p; w0 += 4) {
              unsigned m3 = test4(n0p + n1p + n2p + w0);
              while (m3) {
                const int w = w0 + __ffs(m3) - 1;
                m3 &= m3 - 1;
                unsigned m2 = test4(n0p + n1p + 4 * w);
                while (m2) {
                    const int u = 4 * w + __ffs(m2) - 1;
                    m2 &= m2 - 1;
                    unsigned m1 = test4(n0p + 4 * u);
                    while (m1) {
                        const int m = 4 * u + __ffs(m1) - 1;
                        m1 &= m1 - 1;
                        unsigned m0 = test4(4 * m);
                        while (m0) {
                            const int c = 4 * m + __ffs(m0) - 1;
                            m0 &= m0 - 1;
                            const int j0 = c * TILE;
                            ++searched;
                            if (!CANON) {
                                const float cr2s = SPH_SLACK * __ldg(reinterpret_cast<const float *>(tab + 2 * c + 1) + 1);
                                float thr[CU_QPT];
#pragma unroll
                                for (int k = 0; k < CU_QPT; ++k) thr[k] = thrb[k] + cr2s;
                                if (sphere_search_tile<TILE>(xp + j0, p, b, j0, qx, qy, qz, thr, sgrp, best, bidx))
                                    refresh();
                            } else {
                                __syncwarp();
                                if (TILE >= 32 || lane < TILE) {
#pragma unroll
                                    for (int cc = 0; cc < TILE; cc += 32) {
                                        smx[cc + lane] = planes[j0 + cc + lane];
                                        smy[cc + lane] = planes[p.Mp + j0 + cc + lane];
                                        smz[cc + lane] = planes[2 * p.Mp + j0 + cc + lane];
                                    }
                                }
                                __syncwarp();
                                cu_search_tile<TILE>(smx, smy, smz, oidx + j0, qx, qy, qz, best, bidx);
                            }
                        }
                    }
                }
              }
            }
        };
        if (canonical)
            traverse(std::true_type{});
        else
            traverse(std::false_type{});
        if (MODE == 2) {
            double dsum = 0.0;
#pragma unroll
            for (int k = 0; k < CU_QPT; ++k) {
                const int64_t qi = q0 + k * 32 + lane;
                if (qi < p.N) {
                    dsum += double(best[k]);
                    p.seed[int64_t(b) * p.N + qi] = best[k] < CUDART_INF_F ? bidx[k] : -1;
                }
            }
            dsum = warp_sum(dsum);  // fixed shuffle tree: deterministic
            if (lane == 0) p.sum_partial[int64_t(b) * p.groups + group] = dsum;
        } else {
#pragma unroll
            for (int k = 0; k < CU_QPT; ++k) {
                const int64_t qi = q0 + k * 32 + lane;
                if (qi < p.N) {
                    const int64_t o = int64_t(b) * p.N + qi;
                    const int64_t gi = p.idx_base + bidx[k];
                    p.dist[o] = best[k];
                    if (p.idx_bytes == 8)
                        static_cast<long long *>(p.idx)[o] = gi;
                    else
                        static_cast<int *>(p.idx)[o] = int(gi);
                    if (p.seed != nullptr) p.seed[o] = bidx[k];
                }
            }
        }
    }
    if (p.tiles_searched && lane == 0) atomicAdd(p.tiles_searched, searched);
}

__host__ __device__ inline int64_t ceil_div_dev(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Padded level sizes of the sphere table (whole sibling groups of 4 on every level).
struct SphereLayout {
    int n0, n1, n2, n3, n0p, n1p, n2p, n3p;
    __host__ __device__ SphereLayout(int64_t M, int tile) {
        n0 = int((M + tile - 1) / tile);
        n1 = (n0 + 3) / 4;
        n2 = (n1 + 3) / 4;
        n3 = (n2 + 3) / 4;
        n0p = 4 * n1;
        n1p = 4 * n2;
        n2p = 4 * n3;
        n3p = (n3 + 3) & ~3;
    }
    __host__ __device__ int entries() const { return n0p + n1p + n2p + n3p; }
    __host__ __device__ int64_t float4s(int tile) const { return 2 * int64_t(entries()) + int64_t(n0p) * tile; }
};

// Sphere entries of all four levels: box centre of the finite member points, covering radius (inflated), stored in
// the expanded-test form (see above); padding entries are never asked for.  A level-0 entry (TILE points) is built by a
// group of 4 lanes, an entry of level l by 4 * 4^l lanes (up to a whole warp, which then strides over its 1024-point
// span): every entry costs a few dozen instructions on its critical path instead of a serial walk over its members.
__global__ void __launch_bounds__(128) sphere_table_kernel(const float *__restrict__ planes, int64_t M, int64_t Mp, int tile,
                                                           float4 *__restrict__ table) {
    const SphereLayout L(M, tile);
    const int64_t b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    // work item = (entry, lanes per entry); items are laid out level by level so that a warp never mixes widths
    const int64_t warp = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t w0 = ceil_div_dev(L.n0p, 8), w1 = ceil_div_dev(L.n1p, 2);   // 4 lanes / 16 lanes per entry
    int e, width, sub;
    int64_t span;
    int real;
    if (warp < w0) {
        width = 4; e = int(warp * 8 + (lane >> 2)); sub = lane & 3; span = tile; real = L.n0;
        if (e >= L.n0p) return;
    } else if (warp < w0 + w1) {
        width = 16; const int k = int((warp - w0) * 2 + (lane >> 4)); sub = lane & 15; span = int64_t(tile) * 4; real = L.n1;
        if (k >= L.n1p) return;
        e = L.n0p + k;
    } else if (warp < w0 + w1 + L.n2p) {
        width = 32; const int k = int(warp - w0 - w1); sub = lane; span = int64_t(tile) * 16; real = L.n2;
        e = L.n0p + L.n1p + k;
    } else if (warp < w0 + w1 + L.n2p + L.n3p) {
        width = 32; const int k = int(warp - w0 - w1 - L.n2p); sub = lane; span = int64_t(tile) * 64; real = L.n3;
        e = L.n0p + L.n1p + L.n2p + k;
    } else {
        return;
    }
    const int first = e - (e >= L.n0p + L.n1p + L.n2p ? L.n0p + L.n1p + L.n2p : (e >= L.n0p + L.n1p ? L.n0p + L.n1p : (e >= L.n0p ? L.n0p : 0)));
    const unsigned gmask = width == 32 ? 0xffffffffu : (((1u << width) - 1u) << (lane & ~(width - 1)));
    float4 *out = table + b * L.float4s(tile) + 2 * int64_t(e);
    if (first >= real) {  // padding: NaN limit, every comparison is false
        if (sub == 0) {
            out[0] = make_float4(0.f, 0.f, 0.f, 0.f);
            out[1] = make_float4(__int_as_float(0x7fc00000), 0.f, 0.f, 0.f);
        }
        return;
    }
    const int64_t j0 = int64_t(first) * span;
    const int64_t j1 = (j0 + span < M) ? j0 + span : M;
    const float *P = planes + b * 3 * Mp;
    float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    for (int64_t j = j0 + sub; j < j1; j += width) {
        const float x = P[j], y = P[Mp + j], z = P[2 * Mp + j];
        if (fabsf(x) < CUDART_INF_F && fabsf(y) < CUDART_INF_F && fabsf(z) < CUDART_INF_F) {  // finite points only
            lo[0] = fminf(lo[0], x); hi[0] = fmaxf(hi[0], x);
            lo[1] = fminf(lo[1], y); hi[1] = fmaxf(hi[1], y);
            lo[2] = fminf(lo[2], z); hi[2] = fmaxf(hi[2], z);
        }
    }
    for (int o = width >> 1; o > 0; o >>= 1) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(gmask, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(gmask, hi[a], o));
        }
    }
    if (lo[0] > hi[0]) {  // no finite member: nothing in here can win; asked for only by queries without any bound
        if (sub == 0) {
            out[0] = make_float4(0.f, 0.f, 0.f, -2e-30f);
            out[1] = make_float4(-CUDART_INF_F, 0.f, 0.f, 0.f);
        }
        return;
    }
    const float cx = 0.5f * lo[0] + 0.5f * hi[0], cy = 0.5f * lo[1] + 0.5f * hi[1], cz = 0.5f * lo[2] + 0.5f * hi[2];
    float r2 = 0.f;
    for (int64_t j = j0 + sub; j < j1; j += width) {
        const float x = P[j], y = P[Mp + j], z = P[2 * Mp + j];
        if (fabsf(x) < CUDART_INF_F && fabsf(y) < CUDART_INF_F && fabsf(z) < CUDART_INF_F) {
            const float dx = x - cx, dy = y - cy, dz = z - cz;
            r2 = fmaxf(r2, dx * dx + dy * dy + dz * dz);
        }
    }
    for (int o = width >> 1; o > 0; o >>= 1) r2 = fmaxf(r2, __shfl_xor_sync(gmask, r2, o));
    if (sub != 0) return;
    const float r = sqrtf(r2) * 1.00001f + 1e-30f;
    const float c2 = cx * cx + cy * cy + cz * cz, rr = r * r;
    if (!(c2 + rr < 1e30f)) {  // the expanded test would overflow: always asked for (lim = +inf), full slack
        out[0] = make_float4(0.f, 0.f, 0.f, -2e-30f);
        out[1] = make_float4(CUDART_INF_F, CUDART_INF_F, 0.f, 0.f);
        return;
    }
    const float cr = sqrtf(c2) + r;
    out[0] = make_float4(-2.f * cx, -2.f * cy, -2.f * cz, -2.f * r);
    out[1] = make_float4((rr - c2) + SPH_SLACK * (c2 + rr), cr * cr * 1.00001f, 0.f, 0.f);
}

// one thread per group of four sorted candidates: (-2x, -2y, -2z, |y|^2) x 4 behind the sphere entries
__global__ void sphere_expand_kernel(const float *__restrict__ planes, int64_t M, int64_t Mp, int tile,
                                     float4 *__restrict__ table) {
    const SphereLayout L(M, tile);
    const int64_t b = blockIdx.y;
    const int64_t g = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= int64_t(L.n0p) * tile / 4) return;
    const float *P = planes + b * 3 * Mp;
    float x[4], y[4], z[4], w[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int64_t j = 4 * g + c;
        const bool in = j < M;  // beyond M: +inf like the plane padding (never a candidate)
        const float px = in ? P[j] : CUDART_INF_F, py = in ? P[Mp + j] : CUDART_INF_F, pz = in ? P[2 * Mp + j] : CUDART_INF_F;
        x[c] = -2.0f * px;
        y[c] = -2.0f * py;
        z[c] = -2.0f * pz;
        w[c] = fmaf(pz, pz, fmaf(py, py, px * px));
    }
    float4 *out = table + b * L.float4s(tile) + 2 * int64_t(L.entries()) + 4 * g;
    out[0] = make_float4(x[0], x[1], x[2], x[3]);
    out[1] = make_float4(y[0], y[1], y[2], y[3]);
    out[2] = make_float4(z[0], z[1], z[2], z[3]);
    out[3] = make_float4(w[0], w[1], w[2], w[3]);
}

// ---- fused-loss helpers -----------------------------------------------------------------------------------------
// Accumulators of the fused scene -> body term from the winners (MODE 2): for every batch b and query j with winner
// w = win[b][j] >= 0:  acc[b][w] += (x_j * 2^k, 1).  A warp walks rows of 32 consecutive queries (coalesced index and
// point loads; the static query cloud is served by L2 after the first batch), groups the lanes that picked the same
// winner (match.any), sums every group through the integer warp-reduce unit, and one lane per group touches memory; a
// row with ONE winner (the far field) costs four atomics and no point loads.  Integer adds: the result does not depend
// on the order of arrival.
constexpr int S2B_ROWS = 8;

// Prepare pass (once per call; the query cloud is shared by every batch): xfix[j][0..2] = round(x_j * 2^k) as int64 --
// the accumulate pass then splits limbs with two integer instructions per coordinate instead of converting per batch --
// rowsum[r][0..2] = the sum over the 32 queries of row r (the static part of the one-winner fast path), and *wide != 0
// when some |xfix| >= 2^41 (then the 32-lane sums need three 21-bit limbs, otherwise two).
__global__ void __launch_bounds__(256) s2b_prepare_kernel(const float *__restrict__ x, int64_t N, float fix_scale,
                                                          long long *__restrict__ xfix, long long *__restrict__ rowsum,
                                                          long long *__restrict__ blocksum, unsigned *__restrict__ wide) {
    __shared__ long long srow[8][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = int64_t(blockIdx.x) * 8 + warp;   // 8 rows per CTA = the 256 queries one accumulate warp walks
    const int64_t i = row * 32 + lane;
    const bool live = row * 32 < N;
    bool big = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const long long v = i < N ? __float2ll_rn(__fmul_rn(__ldg(x + 3 * i + a), fix_scale)) : 0;
        if (i < N) xfix[3 * i + a] = v;
        big |= (v >= (1ll << 41)) || (v <= -(1ll << 41));
        const int lo = int(v & 0x1FFFFF), mid = int((v >> 21) & 0x1FFFFF), hi = int(v >> 42);
        const long long slo = __reduce_add_sync(0xffffffffu, lo), smid = __reduce_add_sync(0xffffffffu, mid);
        const long long shi = __reduce_add_sync(0xffffffffu, hi);
        const long long sum = slo + (smid << 21) + (shi << 42);
        if (lane == 0) {
            if (live) rowsum[row * 4 + a] = sum;
            srow[warp][a] = sum;
        }
    }
    if (lane == 0 && live) rowsum[row * 4 + 3] = 0;
    if (__any_sync(0xffffffffu, big) && lane == 0) atomicOr(wide, 1u);
    __syncthreads();
    if (threadIdx.x < 3) {   // blocksum[c][0..2]: the sum over the CTA's 256 queries (the all-rows-one-winner fast path)
        long long sum = 0;
        for (int w = 0; w < 8; ++w) sum += srow[w][threadIdx.x];
        blocksum[int64_t(blockIdx.x) * 4 + threadIdx.x] = sum;
    }
}

__global__ void __launch_bounds__(256) s2b_accum_kernel(const long long *__restrict__ xfix, int64_t N,
                                                        const int *__restrict__ win, int64_t M,
                                                        const long long *__restrict__ rowsum,
                                                        const long long *__restrict__ blocksum,
                                                        const unsigned *__restrict__ wide_flag,
                                                        unsigned long long *__restrict__ acc) {
    const int64_t b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int64_t row0 = (int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (32 * S2B_ROWS);
    if (row0 >= N) return;
    const bool wide = __ldg(wide_flag) != 0u;  // uniform
    unsigned long long *accb = acc + b * M * 4;
    const int *wb = win + b * N;
    int w[S2B_ROWS];
#pragma unroll
    for (int r = 0; r < S2B_ROWS; ++r) {   // all index loads of the warp's rows in flight before the first is used
        const int64_t i = row0 + r * 32 + lane;
        w[r] = i < N ? __ldg(wb + i) : -1;
    }
    {   // ONE winner for all 256 queries of the warp (deep far field): four atomics with the static block sums
        const int t0 = __shfl_sync(0xffffffffu, w[0], 0);
        bool same = true;
#pragma unroll
        for (int r = 0; r < S2B_ROWS; ++r) same &= w[r] == t0;
        if (__all_sync(0xffffffffu, same) && t0 >= 0 && row0 + 32 * S2B_ROWS <= N) {
            if (lane < 4) {
                const unsigned long long add =
                    lane < 3 ? static_cast<unsigned long long>(__ldg(blocksum + (row0 / (32 * S2B_ROWS)) * 4 + lane))
                             : static_cast<unsigned long long>(32 * S2B_ROWS);
                atomicAdd(accb + 4 * int64_t(t0) + lane, add);
            }
            return;
        }
    }
#pragma unroll
    for (int r = 0; r < S2B_ROWS; ++r) {
        if (row0 + r * 32 >= N) break;
        const int64_t i = row0 + r * 32 + lane;
        const bool valid = w[r] >= 0;
        const int t = valid ? w[r] : (-1 - lane);  // a lane past the end / without a winner is a group of its own
        // lanes that picked the same winner, consecutive or not, form one group: one set of atomics per group
        const unsigned grp = __match_any_sync(0xffffffffu, t);
        if (grp == 0xffffffffu && valid) {
            // ONE winner for the whole row (the far field): the row's coordinate sum is static -- four atomics, no point loads
            if (lane < 4) {
                const unsigned long long add =
                    lane < 3 ? static_cast<unsigned long long>(__ldg(rowsum + ((row0 >> 5) + r) * 4 + lane)) : 32ull;
                atomicAdd(accb + 4 * int64_t(t) + lane, add);
            }
            continue;
        }
        long long v[3] = {0, 0, 0};
        if (valid) {
            v[0] = __ldg(xfix + 3 * i);
            v[1] = __ldg(xfix + 3 * i + 1);
            v[2] = __ldg(xfix + 3 * i + 2);
        }
        // group sums through the integer warp-reduce unit (exact: <= 32 addends per limb): a low limb of 21 bits and
        // the signed rest when every |value| < 2^41, three limbs otherwise
        if (!wide) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int lo = int(v[k] & 0x1FFFFF), rest = int(v[k] >> 21);
                const long long slo = __reduce_add_sync(grp, lo), srest = __reduce_add_sync(grp, rest);
                v[k] = slo + (srest << 21);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int lo = int(v[k] & 0x1FFFFF), mid = int((v[k] >> 21) & 0x1FFFFF), hi = int(v[k] >> 42);
                const long long slo = __reduce_add_sync(grp, lo), smid = __reduce_add_sync(grp, mid);
                const long long shi = __reduce_add_sync(grp, hi);
                v[k] = slo + (smid << 21) + (shi << 42);
            }
        }
        if (valid && lane == __ffs(grp) - 1) {
            atomicAdd(accb + 4 * int64_t(t) + 0, static_cast<unsigned long long>(v[0]));
            atomicAdd(accb + 4 * int64_t(t) + 1, static_cast<unsigned long long>(v[1]));
            atomicAdd(accb + 4 * int64_t(t) + 2, static_cast<unsigned long long>(v[2]));
            atomicAdd(accb + 4 * int64_t(t) + 3, static_cast<unsigned long long>(__popc(grp)));
        }
    }
}

// sum_d[b] = sum over the query groups of the per-warp partial sums, in a fixed order (deterministic)
__global__ void __launch_bounds__(256) sphere_sum_kernel(const double *__restrict__ partial, int64_t groups,
                                                         float *__restrict__ sum_d) {
    __shared__ double scratch[32];
    const int64_t b = blockIdx.x;
    double s = 0.0;
    for (int64_t g = threadIdx.x; g < groups; g += blockDim.x) s += partial[b * groups + g];
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) sum_d[b] = float(s);
}

// d/dy of  sum_b g_b sum_j d(x_j, y_b,win(j)):  grad[b][i] = 2 g_b (n_i y_i - S_i), from the integer accumulators
__global__ void __launch_bounds__(256) scene2body_grad_kernel(const float *__restrict__ y,
                                                              const unsigned long long *__restrict__ acc,
                                                              double inv_scale, const float *__restrict__ g, int64_t M,
                                                              float *__restrict__ grad, int accumulate) {
    const int64_t b = blockIdx.y;
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const unsigned long long *a = acc + (b * M + i) * 4;
    const double n = double(a[3]);
    const double g2 = 2.0 * double(g[b]);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double S = double(static_cast<long long>(a[c])) * inv_scale;
        const float v = float(g2 * (n * double(y[(b * M + i) * 3 + c]) - S));
        float *o = grad + (b * M + i) * 3 + c;
        *o = accumulate ? __fadd_rn(*o, v) : v;
    }
}

// ---- ordering helpers: one launch each instead of dozens of elementwise torch ops per step -------------------
__device__ __forceinline__ unsigned spread10(unsigned v) {
    v &= 0x3FFu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// 30-bit Morton code on the grid (lo, inv_cell); non-finite coordinates go to the ends of the axis (they never win a
// search, their position in the order is irrelevant).  Same arithmetic as spatial.morton_keys.
__global__ void morton_keys_kernel(const float *__restrict__ pts, int64_t n, const float *__restrict__ lo,
                                   const float *__restrict__ inv_cell, long long *__restrict__ keys) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned q[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float v = (pts[3 * i + a] - lo[a]) * inv_cell[a];
        if (v != v || v == CUDART_INF_F) v = 1023.0f;   // nan_to_num(nan=1023, posinf=1023, neginf=0)
        if (v == -CUDART_INF_F) v = 0.0f;
        v = fminf(fmaxf(v, 0.0f), 1023.0f);
        q[a] = unsigned(v);                             // truncation, like .to(int64)
    }
    keys[i] = (long long)(spread10(q[0]) | (spread10(q[1]) << 1) | (spread10(q[2]) << 2));
}

// sorted[b][j] = pts[b][perm[j]], SoA planes (+inf padded) and the original-index table (INT32_MAX padded) in one pass
__global__ void gather_pack_kernel(const float *__restrict__ pts, const long long *__restrict__ perm, int64_t perm_bstride,
                                   int64_t M, int64_t Mp, float *__restrict__ sorted, float *__restrict__ planes,
                                   int *__restrict__ oidx) {
    const int64_t b = blockIdx.y;
    const int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= Mp) return;
    float x = CUDART_INF_F, y = CUDART_INF_F, z = CUDART_INF_F;
    int o = 0x7fffffff;
    if (j < M) {
        const long long src = perm[b * perm_bstride + j];
        const float *s = pts + (b * M + src) * 3;
        x = s[0];
        y = s[1];
        z = s[2];
        o = int(src);
        float *d = sorted + (b * M + j) * 3;
        d[0] = x;
        d[1] = y;
        d[2] = z;
    }
    float *pl = planes + b * 3 * Mp;
    pl[j] = x;
    pl[Mp + j] = y;
    pl[2 * Mp + j] = z;
    oidx[b * Mp + j] = o;
}

}  // namespace fpv

using namespace fpv;

extern "C" {

/* mode: 0 = box culling (tiles of 64), 1 = representative + radius culling (tiles of 32). */
int fpv_nn_culled_tile(int mode) { return mode ? CU_TILE_REP : CU_TILE; }

/* Number of floats of the table per candidate batch: 6 * (tiles + super-tiles of 32 tiles). */
size_t fpv_nn_tile_boxes_floats(int64_t M, int mode) {
    const int64_t ntile = ceil_div(M, mode ? CU_TILE_REP : CU_TILE);
    return size_t(ntile + ceil_div(ntile, 32)) * CU_SLOT;
}

/* planes: fpv_nn_pack_planes of the Morton-sorted candidates; orig_idx [batches][Mp]; table out. */
int fpv_nn_tile_boxes(const float *planes, const int32_t *orig_idx, int64_t batches, int64_t M, int mode, float *boxes,
                      fpv_stream_t stream) {
    FPV_CHECK_ARG(planes && boxes && orig_idx && batches > 0 && M > 0 && batches <= 65535, "fpv_nn_tile_boxes: bad arguments");
    const int64_t Mp = ceil_div(M, 64) * 64;
    const int tile = mode ? CU_TILE_REP : CU_TILE;
    const int ntile = int(ceil_div(M, tile));
    const int nsuper = (ntile + 31) / 32;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid((unsigned)ceil_div(ntile, 128), (unsigned)batches), grid2((unsigned)nsuper, (unsigned)batches);
    if (mode) {
        tile_table_kernel<CU_TILE_REP, true><<<grid, 128, 0, st>>>(planes, orig_idx, M, Mp, ntile, nsuper, boxes);
        FPV_LAUNCH_CHECK("tile_table_kernel");
        super_boxes_kernel<CU_TILE_REP><<<grid2, 32, 0, st>>>(planes, M, Mp, ntile, nsuper, boxes);
    } else {
        tile_table_kernel<CU_TILE, false><<<grid, 128, 0, st>>>(planes, orig_idx, M, Mp, ntile, nsuper, boxes);
        FPV_LAUNCH_CHECK("tile_table_kernel");
        super_boxes_kernel<CU_TILE><<<grid2, 32, 0, st>>>(planes, M, Mp, ntile, nsuper, boxes);
    }
    FPV_LAUNCH_CHECK("super_boxes_kernel");
    return FPV_OK;
}

/* Exact NN with culling.  queries: groups of 128 consecutive, spatially compact points
 * ([batches][N][3], or one shared [N][3] set when q_shared); candidates: Morton-sorted planes, their table
 * (fpv_nn_tile_boxes, same mode) and original indices [cand_batches][Mp] (Mp = M rounded up to 64, pad = INT32_MAX)
 * with cand_batches == batches or 1.  Returns, per query, the canonical distance and the
 * ORIGINAL index (+ idx_base) of the lexicographic (distance, original index) minimum -- identical to
 * fpv_nn_search on the unsorted cloud.  tiles_searched (optional, device) accumulates statistics. */
static int culled_search_impl(const float *queries, int q_shared, int64_t batches, int64_t N, const float *planes,
                              const float *boxes, const int32_t *orig_idx, int64_t cand_batches, int64_t M, int mode,
                              int64_t idx_base, float *dist, void *idx, int idx_bytes,
                              unsigned long long *tiles_searched, const float *cand_orig, int32_t *seed_inout,
                              int seed_valid, unsigned long long *keys, unsigned long long *const *push, int push_n,
                              const unsigned *push_parity, int64_t push_half, cudaStream_t st) {
    CulledParams p;
    int64_t eb = batches, eN = N;
    p.q_bstride = q_shared ? 0 : N * 3;
    if (cand_batches == 1 && !q_shared) {  // one candidate set for every batch: one flat batch of queries
        eN = batches * N;
        eb = 1;
        p.q_bstride = 0;
    }
    p.q = queries;
    p.N = eN;
    p.planes = planes;
    p.Mp = ceil_div(M, 64) * 64;
    p.M = M;
    p.ntile = int(ceil_div(M, mode ? CU_TILE_REP : CU_TILE));
    p.nsuper = (p.ntile + 31) / 32;
    p.plane_bstride = cand_batches == 1 ? 0 : 3 * p.Mp;
    p.boxes = boxes;
    p.box_bstride = cand_batches == 1 ? 0 : int64_t(p.ntile + p.nsuper) * CU_SLOT;
    p.oidx = orig_idx;
    p.oidx_bstride = cand_batches == 1 ? 0 : p.Mp;  // original indices are padded like the planes
    p.idx_base = idx_base;
    p.dist = dist;
    p.idx = idx;
    p.idx_bytes = idx_bytes;
    p.tiles_searched = tiles_searched;
    p.cand_orig = cand_orig;
    p.cand_orig_bstride = cand_batches == 1 ? 0 : M * 3;
    p.seed = seed_inout;
    p.seed_read = (seed_inout && seed_valid) ? 1 : 0;
    p.keys = keys;
    p.push_n = push_n;
    for (int r = 0; r < FPV_MAX_PEERS; ++r) p.push[r] = (r < push_n) ? push[r] : nullptr;
    p.push_parity = push_parity;
    p.push_half = push_half;
    dim3 grid((unsigned)ceil_div(ceil_div(eN, CU_GROUP), CU_WARPS), (unsigned)eb);
    if (profile_on()) {
        // algorithmic bytes per SURVEY 8(d): queries + candidates once + (d, i) out; seeds are not counted
        char nm[48];
        snprintf(nm, sizeof(nm), "nn_culled<%s> Q=%lld M=%lld", mode ? "rep" : "box", (long long)(eb * eN), (long long)M);
        profile_begin(nm, st,
                      12.0 * double(q_shared ? N : batches * N) + 12.0 * double(M) * double(cand_batches) +
                          (keys ? 8.0 * (1 + push_n) : 4.0 + idx_bytes) * double(eb * eN),
                      double(eb * eN) * double(M));
    }
    if (mode)
        nn_culled_kernel<CU_TILE_REP, true><<<grid, CU_WARPS * 32, 0, st>>>(p);
    else
        nn_culled_kernel<CU_TILE, false><<<grid, CU_WARPS * 32, 0, st>>>(p);
    profile_end(st);
    FPV_LAUNCH_CHECK("nn_culled_kernel");
    return FPV_OK;
}

int fpv_nn_culled_search(const float *queries, int q_shared, int64_t batches, int64_t N, const float *planes,
                         const float *boxes, const int32_t *orig_idx, int64_t cand_batches, int64_t M, int mode,
                         int64_t idx_base, float *dist, void *idx, int idx_bytes,
                         unsigned long long *tiles_searched, const float *cand_orig, int32_t *seed_inout,
                         int seed_valid, fpv_stream_t stream) {
    FPV_CHECK_ARG(queries && planes && boxes && orig_idx && dist && idx, "fpv_nn_culled_search: null pointer");
    FPV_CHECK_ARG(!seed_inout || (cand_orig && mode == 0), "fpv_nn_culled_search: seeds need cand_orig and box mode");
    FPV_CHECK_ARG(batches > 0 && N > 0 && M > 0 && batches <= 65535, "fpv_nn_culled_search: empty input");
    FPV_CHECK_ARG(cand_batches == 1 || cand_batches == batches, "fpv_nn_culled_search: cand_batches must be 1 or batches");
    FPV_CHECK_ARG(idx_bytes == 4 || idx_bytes == 8, "fpv_nn_culled_search: idx_bytes must be 4 or 8");
    return culled_search_impl(queries, q_shared, batches, N, planes, boxes, orig_idx, cand_batches, M, mode, idx_base, dist,
                              idx, idx_bytes, tiles_searched, cand_orig, seed_inout, seed_valid, nullptr, nullptr, 0,
                              nullptr, 0, static_cast<cudaStream_t>(stream));
}

/* The same search with the multi-GPU epilogue: every query's result is written as ONE packed key
 * (float_bits(d) << 32 | idx_base + index; canonical d >= 0, so unsigned integer order == lexicographic (d, index) order)
 * into keys[], and the same key is stored into push_n peer buffers of the same layout (device pointers into the other
 * ranks' mailboxes, mapped with fpv_p2p_open) straight from the search epilogue.  When push_parity is given, bit 0 of the
 * device word selects the half (offset push_half elements) of keys / push buffers (double-buffered mailbox). */
int fpv_nn_culled_search_keys(const float *queries, int q_shared, int64_t batches, int64_t N, const float *planes,
                              const float *boxes, const int32_t *orig_idx, int64_t cand_batches, int64_t M,
                              int64_t idx_base, uint64_t *keys, uint64_t *const *push_host, int push_n,
                              const uint32_t *push_parity, int64_t push_half, unsigned long long *tiles_searched,
                              const float *cand_orig, int32_t *seed_inout, int seed_valid, fpv_stream_t stream) {
    FPV_CHECK_ARG(queries && planes && boxes && orig_idx && keys, "fpv_nn_culled_search_keys: null pointer");
    FPV_CHECK_ARG(!seed_inout || cand_orig, "fpv_nn_culled_search_keys: seeds need cand_orig");
    FPV_CHECK_ARG(batches > 0 && N > 0 && M > 0 && batches <= 65535, "fpv_nn_culled_search_keys: empty input");
    FPV_CHECK_ARG(cand_batches == 1 || cand_batches == batches, "fpv_nn_culled_search_keys: cand_batches must be 1 or batches");
    FPV_CHECK_ARG(push_n >= 0 && push_n <= FPV_MAX_PEERS && (push_n == 0 || push_host), "fpv_nn_culled_search_keys: bad push list");
    FPV_CHECK_ARG(idx_base >= 0 && idx_base + M <= 0xFFFFFFFFll, "fpv_nn_culled_search_keys: global index does not fit 32 bits");
    return culled_search_impl(queries, q_shared, batches, N, planes, boxes, orig_idx, cand_batches, M, 0, idx_base, nullptr,
                              nullptr, 4, tiles_searched, cand_orig, seed_inout, seed_valid,
                              reinterpret_cast<unsigned long long *>(keys),
                              reinterpret_cast<unsigned long long *const *>(push_host), push_n, push_parity, push_half,
                              static_cast<cudaStream_t>(stream));
}


/* ---- sphere-hierarchy mode (tile = 16 or 32 points) ---- */
size_t fpv_nn_sphere_table_floats(int64_t M, int tile) {
    if (M <= 0 || (tile != 16 && tile != 32)) return 0;
    return size_t(SphereLayout(M, tile).float4s(tile)) * 4;
}

int fpv_nn_sphere_table(const float *planes, int64_t batches, int64_t M, int tile, float *table, fpv_stream_t stream) {
    FPV_CHECK_ARG(planes && table && batches > 0 && M > 0 && batches <= 65535, "fpv_nn_sphere_table: bad arguments");
    FPV_CHECK_ARG(tile == 16 || tile == 32, "fpv_nn_sphere_table: tile must be 16 or 32");
    FPV_CHECK_ARG((reinterpret_cast<uintptr_t>(table) & 15) == 0, "fpv_nn_sphere_table: table must be 16-byte aligned");
    const int64_t Mp = ceil_div(M, 64) * 64;
    const SphereLayout L(M, tile);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t warps = ceil_div(L.n0p, 8) + ceil_div(L.n1p, 2) + L.n2p + L.n3p;
    dim3 grid((unsigned)ceil_div(warps, 4), (unsigned)batches);
    sphere_table_kernel<<<grid, 128, 0, st>>>(planes, M, Mp, tile, reinterpret_cast<float4 *>(table));
    FPV_LAUNCH_CHECK("sphere_table_kernel");
    dim3 grid2((unsigned)ceil_div(int64_t(L.n0p) * tile / 4, 128), (unsigned)batches);
    sphere_expand_kernel<<<grid2, 128, 0, st>>>(planes, M, Mp, tile, reinterpret_cast<float4 *>(table));
    FPV_LAUNCH_CHECK("sphere_expand_kernel");
    return FPV_OK;
}

static int g_sphere_ctas_per_sm = 512;
static int g_sphere_minb = 7;   // resident CTAs per SM the fused tile-16 kernel is compiled for (8: 64 regs, 7: 72, 6: 80);
                                // measured at cfg 2: 11.19 / 10.96 / 11.28 ms incl. the accumulate pass (profiles/r02_sphere_order_tune.txt)

/* Frame chunking of the temporally seeded sphere search: the grid is sized to about ctas_per_sm CTAs per SM
 * (more CTAs = better load balance over the heavy-tailed per-group cost, but every chunk pays one unseeded frame). */
int fpv_nn_sphere_set_chunking(int ctas_per_sm) {
    if (ctas_per_sm >= -8 && ctas_per_sm <= -6) {   // tuning: negative values select the register budget instead
        g_sphere_minb = -ctas_per_sm;
        return FPV_OK;
    }
    FPV_CHECK_ARG(ctas_per_sm >= 1 && ctas_per_sm <= 4096, "fpv_nn_sphere_set_chunking: ctas_per_sm out of range");
    g_sphere_ctas_per_sm = ctas_per_sm;
    return FPV_OK;
}

static int sphere_search_impl(const float *queries, int q_shared, int64_t batches, int64_t N, const float *planes,
                              const float *table, const int32_t *orig_idx, const int32_t *pos_of, int pos_shared,
                              int32_t *seed_inout,
                              int seed_valid, int64_t M, int tile, int64_t idx_base, float *dist, void *idx, int idx_bytes,
                              unsigned long long *tiles_searched, double *sum_partial, unsigned long long *acc,
                              int fix_shift, cudaStream_t st) {
    const bool fused = sum_partial != nullptr;
    SphereParams p;
    p.q = queries;
    p.q_bstride = q_shared ? 0 : N * 3;
    p.N = N;
    p.planes = planes;
    p.Mp = ceil_div(M, 64) * 64;
    p.M = M;
    p.plane_bstride = 3 * p.Mp;
    const SphereLayout L(M, tile);
    p.n0 = L.n0;
    p.n1 = L.n1;
    p.n2 = L.n2;
    p.n3 = L.n3;
    p.table = reinterpret_cast<const float4 *>(table);
    p.table_bstride = L.float4s(tile);
    p.oidx = orig_idx;
    p.oidx_bstride = p.Mp;
    p.pos_of = pos_of;
    p.pos_bstride = pos_shared ? 0 : M;
    p.seed = seed_inout;
    p.seed_read = (seed_inout && seed_valid) ? 1 : 0;
    p.batches = int(batches);
    p.idx_base = idx_base;
    p.dist = dist;
    p.idx = idx;
    p.idx_bytes = idx_bytes;
    p.tiles_searched = tiles_searched;
    p.sum_partial = sum_partial;
    p.groups = ceil_div(N, CU_GROUP);
    p.acc = acc;
    p.fix_scale = ldexpf(1.0f, fix_shift);
    const int64_t ctas_x = ceil_div(ceil_div(N, CU_GROUP), CU_WARPS);
    int64_t nchunks = batches;
    if (q_shared && pos_of) {  // walk frames inside the warp, but keep >= ~2 resident waves of CTAs
        // with per-call seeds every frame starts seeded: frames need not share a CTA, so the grid can be much finer
        nchunks = ceil_div(int64_t(sm_count()) * g_sphere_ctas_per_sm * (p.seed_read ? 4 : 1), ctas_x);
        if (nchunks < 1) nchunks = 1;
        if (nchunks > batches) nchunks = batches;
    }
    p.frames_per_cta = int(ceil_div(batches, nchunks));
    nchunks = ceil_div(batches, p.frames_per_cta);
    dim3 grid((unsigned)ctas_x, (unsigned)nchunks);
    if (profile_on()) {
        // algorithmic bytes per SURVEY 8(d): queries once, candidates once per batch, outputs; the carried seeds are an
        // implementation device, not part of the algorithmic traffic, and are NOT counted
        char nm[48];
        snprintf(nm, sizeof(nm), "nn_sphere%s<%d> Q=%lld M=%lld", fused ? "_fused" : "", tile, (long long)(batches * N),
                 (long long)M);
        const double out_bytes = fused ? 8.0 * double(batches) + 12.0 * double(M) * double(batches)
                                       : (4.0 + idx_bytes) * double(batches * N);
        profile_begin(nm, st, 12.0 * double(q_shared ? N : batches * N) + 12.0 * double(M) * double(batches) + out_bytes,
                      double(batches * N) * double(M));
    }
    if (fused) {
        if (tile == 16 && g_sphere_minb == 7)
            nn_sphere_kernel<16, 7, 2><<<grid, CU_WARPS * 32, 0, st>>>(p);
        else if (tile == 16 && g_sphere_minb == 6)
            nn_sphere_kernel<16, 6, 2><<<grid, CU_WARPS * 32, 0, st>>>(p);
        else if (tile == 16)
            nn_sphere_kernel<16, 8, 2><<<grid, CU_WARPS * 32, 0, st>>>(p);
        else
            nn_sphere_kernel<32, 8, 2><<<grid, CU_WARPS * 32, 0, st>>>(p);
    } else if (tile == 16) {
        nn_sphere_kernel<16, 8, 0><<<grid, CU_WARPS * 32, 0, st>>>(p);  // 64 registers, 32 resident warps: latency-bound
    } else {
        nn_sphere_kernel<32, 8, 0><<<grid, CU_WARPS * 32, 0, st>>>(p);
    }
    profile_end(st);
    FPV_LAUNCH_CHECK("nn_sphere_kernel");
    return FPV_OK;
}

/* Exact NN through the sphere hierarchy.  pos_of (optional): the sorted position of every ORIGINAL candidate index
 * ([M] when pos_shared, else [batches][M]); it turns winners (original indices) back into table positions, which
 * enables seeding -- across consecutive batches (frames) of a shared query set, and from call to call (seed_inout). */
int fpv_nn_sphere_search(const float *queries, int q_shared, int64_t batches, int64_t N, const float *planes,
                         const float *table, const int32_t *orig_idx, const int32_t *pos_of, int pos_shared,
                         int32_t *seed_inout, int seed_valid, int64_t M, int tile, int64_t idx_base, float *dist,
                         void *idx, int idx_bytes, unsigned long long *tiles_searched, fpv_stream_t stream) {
    FPV_CHECK_ARG(queries && planes && table && orig_idx && dist && idx, "fpv_nn_sphere_search: null pointer");
    FPV_CHECK_ARG(!seed_inout || pos_of, "fpv_nn_sphere_search: seed_inout needs pos_of");
    FPV_CHECK_ARG(batches > 0 && N > 0 && M > 0 && batches <= 65535, "fpv_nn_sphere_search: empty input");
    FPV_CHECK_ARG(idx_bytes == 4 || idx_bytes == 8, "fpv_nn_sphere_search: idx_bytes must be 4 or 8");
    FPV_CHECK_ARG(tile == 16 || tile == 32, "fpv_nn_sphere_search: tile must be 16 or 32");
    FPV_CHECK_ARG((reinterpret_cast<uintptr_t>(table) & 15) == 0, "fpv_nn_sphere_search: table must be 16-byte aligned");
    return sphere_search_impl(queries, q_shared, batches, N, planes, table, orig_idx, pos_of, pos_shared, seed_inout, seed_valid, M,
                              tile, idx_base, dist, idx, idx_bytes, tiles_searched, nullptr, nullptr, 0,
                              static_cast<cudaStream_t>(stream));
}

/* Fused scene -> body term (SURVEY.md section 7 "hard parts", section 8d "fused-loss variant"): the same exact search
 * for ONE query set shared by every batch (the static scene, N points) against per-batch candidates (the body, M
 * vertices), but nothing of size [batches][N] is written except the in/out seeds:
 *   sum_d[b]      = sum_j min_i d(x_j, y_b,i)                                   (double accumulation, fixed order)
 *   acc[b][i][0..2] = sum over the queries that candidate i won of x_j * 2^fix_shift (two's complement), acc[b][i][3] = count
 * acc must be zero on entry ([batches][M][4] uint64).  fpv_scene2body_grad turns acc into d sum / d y.
 * fix_shift: choose 2^fix_shift * max|x| * (N + 1) < 2^62 (fpv_fix_shift_for).  Queries whose winner distance is not
 * finite add +inf / NaN to sum_d and nothing to acc. */
size_t fpv_nn_sphere_fused_workspace_bytes(int64_t batches, int64_t N) {
    if (batches <= 0 || N <= 0) return 0;
    return align_up(size_t(batches) * size_t(ceil_div(N, CU_GROUP)) * sizeof(double), 256) +
           align_up(size_t(ceil_div(N, 32)) * 4 * sizeof(long long), 256) + align_up(size_t(N) * 3 * sizeof(long long), 256) +
           align_up(size_t(ceil_div(N, 32 * S2B_ROWS)) * 4 * sizeof(long long), 256) + 256 + 256;
}

int fpv_fix_shift_for(float max_abs_coordinate, int64_t count) {
    int e = 0;
    if (max_abs_coordinate > 0.f && max_abs_coordinate <= 3.4e38f) frexpf(max_abs_coordinate, &e);
    int bits = 0;
    while ((int64_t(1) << bits) < count + 1) ++bits;
    int k = 61 - bits - e;
    return k < -100 ? -100 : (k > 100 ? 100 : k);
}

int fpv_nn_sphere_fused(const float *queries, int64_t batches, int64_t N, const float *planes, const float *table,
                        const int32_t *orig_idx, const int32_t *pos_of, int pos_shared, int32_t *seed_inout, int seed_valid, int64_t M,
                        int tile, int fix_shift, float *sum_d, unsigned long long *acc,
                        unsigned long long *tiles_searched, void *workspace, size_t workspace_bytes,
                        fpv_stream_t stream) {
    FPV_CHECK_ARG(queries && planes && table && orig_idx && pos_of && sum_d && acc && seed_inout,
                  "fpv_nn_sphere_fused: null pointer (the in/out seed buffer is required: it carries the winners)");
    FPV_CHECK_ARG(batches > 0 && N > 0 && M > 0 && batches <= 65535, "fpv_nn_sphere_fused: empty input");
    FPV_CHECK_ARG(tile == 16 || tile == 32, "fpv_nn_sphere_fused: tile must be 16 or 32");
    FPV_CHECK_ARG((reinterpret_cast<uintptr_t>(table) & 15) == 0, "fpv_nn_sphere_fused: table must be 16-byte aligned");
    FPV_CHECK_ARG(fix_shift >= -100 && fix_shift <= 100, "fpv_nn_sphere_fused: fix_shift out of range");
    Arena ar(workspace, workspace_bytes);
    const int64_t groups = ceil_div(N, CU_GROUP);
    double *partial = ar.take<double>(size_t(batches) * size_t(groups));
    long long *rowsum = ar.take<long long>(size_t(ceil_div(N, 32)) * 4);
    long long *xfix = ar.take<long long>(size_t(N) * 3);
    long long *blocksum = ar.take<long long>(size_t(ceil_div(N, 32 * S2B_ROWS)) * 4);
    unsigned *wide = ar.take<unsigned>(1);
    if (!partial || !rowsum || !xfix || !blocksum || !wide) {
        set_error("fpv_nn_sphere_fused: workspace too small (%zu bytes, need %zu)", workspace_bytes,
                  fpv_nn_sphere_fused_workspace_bytes(batches, N));
        return FPV_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = sphere_search_impl(queries, 1, batches, N, planes, table, orig_idx, pos_of, pos_shared, seed_inout, seed_valid, M, tile,
                                0, nullptr, nullptr, 4, tiles_searched, partial, acc, fix_shift, st);
    if (rc) return rc;
    {
        dim3 grid((unsigned)ceil_div(N, 256 * S2B_ROWS), (unsigned)batches);
        if (profile_on()) {
            char nm[48];
            snprintf(nm, sizeof(nm), "s2b_accum Q=%lld M=%lld", (long long)(batches * N), (long long)M);
            profile_begin(nm, st, 12.0 * double(N) + 4.0 * double(batches * N) + 32.0 * double(batches * M), double(batches * N));
        }
        // (the fixed-point copy and the row sums are static per query cloud; rebuilding them costs one pass of 36 N bytes,
        // ~10 us per million points)
        FPV_CUDA(cudaMemsetAsync(wide, 0, sizeof(unsigned), st));
        static_assert(S2B_ROWS == 8, "s2b_prepare_kernel sums one accumulate warp's queries per CTA of 8 warps");
        s2b_prepare_kernel<<<(unsigned)ceil_div(N, 32 * S2B_ROWS), 256, 0, st>>>(queries, N, ldexpf(1.0f, fix_shift), xfix, rowsum,
                                                                               blocksum, wide);
        count_launch();
        s2b_accum_kernel<<<grid, 256, 0, st>>>(xfix, N, seed_inout, M, rowsum, blocksum, wide, acc);
        profile_end(st);
        FPV_LAUNCH_CHECK("s2b_accum_kernel");
    }
    sphere_sum_kernel<<<(unsigned)batches, 256, 0, st>>>(partial, groups, sum_d);
    FPV_LAUNCH_CHECK("sphere_sum_kernel");
    return FPV_OK;
}

/* grad_y[b][i][:] (+)= 2 g[b] (count y - S 2^-fix_shift): the gradient of sum_b g[b] sum_d[b] w.r.t. the candidates
 * (accumulate != 0 adds to what grad already holds). */
int fpv_scene2body_grad(const float *cand /*[batches][M][3]*/, const unsigned long long *acc, int fix_shift,
                        const float *g /*[batches]*/, int64_t batches, int64_t M, float *grad /*[batches][M][3]*/,
                        int accumulate, fpv_stream_t stream) {
    FPV_CHECK_ARG(cand && acc && g && grad, "fpv_scene2body_grad: null pointer");
    FPV_CHECK_ARG(batches > 0 && M > 0 && batches <= 65535, "fpv_scene2body_grad: empty input");
    dim3 grid((unsigned)ceil_div(M, 256), (unsigned)batches);
    scene2body_grad_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(cand, acc, ldexp(1.0, -fix_shift), g, M, grad,
                                                                               accumulate);
    FPV_LAUNCH_CHECK("scene2body_grad_kernel");
    return FPV_OK;
}

/* 30-bit Morton keys of n points on the grid (lo[3], inv_cell[3] device pointers). */
int fpv_morton_keys(const float *pts, int64_t n, const float *lo, const float *inv_cell, long long *keys, fpv_stream_t stream) {
    FPV_CHECK_ARG(pts && lo && inv_cell && keys && n > 0, "fpv_morton_keys: bad arguments");
    morton_keys_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(pts, n, lo, inv_cell, keys);
    FPV_LAUNCH_CHECK("morton_keys_kernel");
    return FPV_OK;
}

/* Apply an ordering: sorted [B][M][3], planes (fpv_nn_planes_bytes) and orig_idx [B][Mp] from pts [B][M][3] and
 * perm ([M] shared by every batch entry when perm_shared, else [B][M]; int64 sorted position -> original index). */
int fpv_nn_gather_pack(const float *pts, const long long *perm, int perm_shared, int64_t batches, int64_t M, float *sorted,
                       float *planes, int32_t *orig_idx, fpv_stream_t stream) {
    FPV_CHECK_ARG(pts && perm && sorted && planes && orig_idx, "fpv_nn_gather_pack: null pointer");
    FPV_CHECK_ARG(batches > 0 && batches <= 65535 && M > 0, "fpv_nn_gather_pack: bad sizes");
    const int64_t Mp = ceil_div(M, 64) * 64;
    dim3 grid((unsigned)ceil_div(Mp, 256), (unsigned)batches);
    gather_pack_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(pts, perm, perm_shared ? 0 : M, M, Mp, sorted,
                                                                            planes, orig_idx);
    FPV_LAUNCH_CHECK("gather_pack_kernel");
    return FPV_OK;
}

}  // extern "C"
