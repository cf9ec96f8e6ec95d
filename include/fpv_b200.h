/*
 * fpv_b200.h -- C ABI of the B200-native hot path of 4DCapture-FPV's global_optimization stage.
 *
 * The reference has no FFI layer: its boundary is Python call signatures on torch tensors
 * (SURVEY.md section 8b).  This header is what a host in any language binds instead; the Python
 * mirror of the reference signatures (4dcapture-fpv_b200/*.py) is a ctypes client of exactly
 * these symbols.  Conventions:
 *   - every pointer is a DEVICE pointer unless the parameter name ends in _host;
 *   - nothing allocates: callers size scratch with the *_bytes() queries and pass it in;
 *   - `stream` is a cudaStream_t; all work is enqueued on it, no host synchronisation inside;
 *   - return value 0 = ok, non-zero = error, text via fpv_last_error() (thread-local);
 *   - fp32 data, int32 or int64 indices (idx_bytes = 4 | 8; the reference returns int64).
 * Reference citations are relative to /root/reference/.
 */
#ifndef FPV_B200_H
#define FPV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FPV_ABI_VERSION 1
#define FPV_OK 0
#define FPV_ERR_INVALID 1
#define FPV_ERR_CUDA 2
#define FPV_ERR_WORKSPACE 3
#define FPV_MAX_PEERS 7 /* other ranks of one 8-GPU box */

typedef void *fpv_stream_t; /* cudaStream_t */

const char *fpv_last_error(void);
int fpv_abi_version(void);
/* sm_count / compute capability of `device`; fails loudly when no sm_100 GPU is present. */
int fpv_device_query(int device, int *sm_count, int *cc_major, int *cc_minor);

/* Measurement hooks (bench.py): kernel launches issued by this library since load; per-kernel CUDA-event
 * timing of the instrumented launches with their algorithmic bytes / work items; an FP32 issue-rate probe
 * (fused multiply-add lane operations per second) that gives the SIMT roofline denominator. */
unsigned long long fpv_launch_count(void);
int fpv_profile_enable(int on);
int fpv_profile_count(void);
int fpv_profile_get(int i, char *name48, float *ms, double *algo_bytes, double *work_items);
int fpv_fp32_probe(double *lane_fma_per_s, fpv_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Chamfer nearest neighbour  (replaces chamfer_python.py:18-28 distChamfer and the [3P] CUDA op
 * ext.chamferDist() called at global_optimization.py:292-294, :349-353).
 * Canonical arithmetic: d = fma(dz,dz, fma(dy,dy, dx*dx)), dx = q - r in fp32; winner = the
 * lexicographic minimum of (d, index) -- lowest index on ties, like torch.min / strict '<'.
 * ------------------------------------------------------------------------------------------ */

/* Candidate clouds are searched from padded structure-of-arrays planes [batches][3][Mp]. */
size_t fpv_nn_planes_bytes(int64_t batches, int64_t M);
int fpv_nn_pack_planes(const float *pts /*[batches,M,3]*/, int64_t batches, int64_t M,
                       float *planes, fpv_stream_t stream);

/* One direction: for every query the nearest candidate.
 *   queries [q_batches,N,3] (q_shared != 0: one [N,3] set reused by every batch);
 *   ref_planes from fpv_nn_pack_planes with ref_batches == batches, or 1 (shared by all batches);
 *   idx_base is added to every index (scene shards keep global indices);
 *   dist [batches,N] / idx [batches,N] may be NULL when only keys are wanted;
 *   keys [batches,N] (optional) receives (float_bits(d) << 32) | idx, the multi-GPU combine key. */
size_t fpv_nn_search_workspace_bytes(int64_t batches, int64_t N, int64_t M);
int fpv_nn_search(const float *queries, int q_shared, int64_t batches, int64_t N,
                  const float *ref_planes, int64_t ref_batches, int64_t M, int64_t idx_base,
                  float *dist, void *idx, int idx_bytes, uint64_t *keys, void *workspace,
                  size_t workspace_bytes, fpv_stream_t stream);
int fpv_nn_unpack_keys(const uint64_t *keys, int64_t n, float *dist, void *idx, int idx_bytes,
                       fpv_stream_t stream);
/* Tuning hook for bench sweeps: force queries-per-thread (4 | 8; 0 = heuristic), the candidate split
 * (0 = heuristic) and how many queries per thread use packed FP32x2 math (-1 = default).
 * Results never depend on it. */
int fpv_nn_set_tuning(int qpt, int nsplit, int packed);
/* Search engine: 0 = auto, 1 = FP32 SIMT brute force (nn_search_kernel), 2 = tensor-core filter with exact
 * FP32 re-check (nn_tc_kernel).  Both return bit-identical results.  tc_eshift > 0 overrides the filter's
 * error-bound exponent (default 15); used by the tests to demonstrate the safety margin. */
int fpv_nn_set_engine(int engine, int tc_eshift);
/* Debug: device buffer of 1024 int64 receiving a clock64 pipeline timeline of CTA 0 of nn_tc_kernel (NULL = off). */
int fpv_nn_tc_debug(long long *dbg);

/* Spatially indexed exact search (nn_culled.cu): the candidate cloud is Morton-sorted by the caller and cut
 * into tiles of fpv_nn_culled_tile() points with bounding boxes; queries come in spatially compact groups of
 * 128.  Tiles that provably cannot hold a winner are skipped -- mode 0: box-to-box gap against the group's worst
 * best-distance; mode 1: per-query triangle inequality through each tile's representative point and covering
 * radius -- and the result (canonical distance, ORIGINAL candidate index, lowest original index on ties) equals
 * fpv_nn_search's.
 *   planes   : fpv_nn_pack_planes of the sorted candidates        boxes : fpv_nn_tile_boxes
 *   orig_idx : [cand_batches][Mp] original index of every sorted candidate, Mp = M rounded up to 64,
 *              padding = INT32_MAX */
int fpv_nn_culled_tile(int mode); /* mode 0: box culling, tiles of 64; mode 1: representative + radius, tiles of 32 */
size_t fpv_nn_tile_boxes_floats(int64_t M, int mode);
int fpv_nn_tile_boxes(const float *planes, const int32_t *orig_idx, int64_t batches, int64_t M, int mode,
                      float *boxes, fpv_stream_t stream);
int fpv_nn_culled_search(const float *queries, int q_shared, int64_t batches, int64_t N, const float *planes,
                         const float *boxes, const int32_t *orig_idx, int64_t cand_batches, int64_t M, int mode,
                         int64_t idx_base, float *dist, void *idx, int idx_bytes,
                         unsigned long long *tiles_searched, const float *cand_orig, int32_t *seed_inout,
                         int seed_valid, fpv_stream_t stream);
/*   cand_orig / seed_inout / seed_valid (mode 0, optional): the candidates in ORIGINAL order [cand_batches][M][3] and a
 *   [batches][N] int32 buffer of the previous call's winners (original indices, without idx_base), read as starting
 *   points when seed_valid and always overwritten -- hints only, see fpv_nn_sphere_search. */

/* The same search (box mode) with the multi-GPU epilogue (replaces the separate pack pass + NCCL
 * send buffer of the survey's design): every query's result is written as ONE packed key
 * (float_bits(d) << 32 | idx_base + index; canonical d >= 0, so unsigned integer order == lexicographic (d, index)
 * order) into keys[], and the same key is stored into push_n peer buffers of the same layout (push_host: HOST array of
 * device pointers into the other ranks' mailboxes, mapped with fpv_p2p_open) straight from the search epilogue, so the
 * NVLink transfer overlaps the search.  push_parity (optional device word): bit 0 selects the half (offset push_half
 * elements) of keys / push buffers of a double-buffered mailbox. */
int fpv_nn_culled_search_keys(const float *queries, int q_shared, int64_t batches, int64_t N, const float *planes,
                              const float *boxes, const int32_t *orig_idx, int64_t cand_batches, int64_t M,
                              int64_t idx_base, uint64_t *keys, uint64_t *const *push_host, int push_n,
                              const uint32_t *push_parity, int64_t push_half, unsigned long long *tiles_searched,
                              const float *cand_orig, int32_t *seed_inout, int seed_valid, fpv_stream_t stream);

/* Ordering helpers of the spatial engines (one launch each): 30-bit Morton keys of n points on the grid
 * (lo[3], inv_cell[3] are DEVICE pointers), and the application of an ordering -- sorted points, padded SoA planes
 * and the original-index table in one pass (perm: int64, [M] when perm_shared, else [batches][M]). */
int fpv_morton_keys(const float *pts, int64_t n, const float *lo, const float *inv_cell, long long *keys,
                    fpv_stream_t stream);
int fpv_nn_gather_pack(const float *pts, const long long *perm, int perm_shared, int64_t batches, int64_t M,
                       float *sorted, float *planes, int32_t *orig_idx, fpv_stream_t stream);

/* Frame chunking of fpv_nn_sphere_search with temporal seeding: size the grid to about ctas_per_sm CTAs per SM
 * (default 512).  More chunks balance the heavy-tailed per-group cost; every chunk pays one unseeded frame. */
int fpv_nn_sphere_set_chunking(int ctas_per_sm);

/* Sphere-hierarchy variant for moving candidate sets (scene -> body): clusters of `tile` (16 | 32) sorted points
 * with bounding spheres on four levels (tile, x4, x16, x64 points); a query needs a cluster only if |x - c| <= sqrt(best_x) + r.  With a
 * shared query set and pos_of (the sorted position of every ORIGINAL candidate index: [M] when pos_shared, else
 * [batches][M]) consecutive batches (frames) seed each other: every query starts from the exact distance to its
 * previous frame's winner and the three sorted candidates next to it.
 * seed_inout (optional, device, [batches][N] int32, needs pos_of): the winners of the previous CALL on the same
 * problem (original candidate indices; any value outside [0,M) = no seed, e.g. -1 on the first call); read as the
 * starting point of every (batch, query) when seed_valid != 0, and always overwritten with this call's winners.
 * Seeds are hints: any content yields the same exact result.
 * Table: 8 floats per sphere (expanded-test form, see nn_culled.cu).  tiles_searched (optional, device):
 * += clusters searched by a warp. */
size_t fpv_nn_sphere_table_floats(int64_t M, int tile);
int fpv_nn_sphere_table(const float *planes, int64_t batches, int64_t M, int tile, float *table,
                        fpv_stream_t stream);
int fpv_nn_sphere_search(const float *queries, int q_shared, int64_t batches, int64_t N, const float *planes,
                         const float *table, const int32_t *orig_idx, const int32_t *pos_of, int pos_shared,
                         int32_t *seed_inout, int seed_valid, int64_t M, int tile, int64_t idx_base, float *dist,
                         void *idx, int idx_bytes, unsigned long long *tiles_searched, fpv_stream_t stream);

/* Fused scene -> body term of the chamfer loss (SURVEY.md section 7 "hard parts", section 8d "fused-loss variant";
 * the direction of chamfer_python.py:28 that min-reduces over the BODY for every scene point): the same exact sphere
 * search for ONE query set shared by every batch (the static scene, N points) against per-batch candidates (the body,
 * M vertices, sphere table as for fpv_nn_sphere_search), but nothing of size [batches][N] is written except the in/out
 * seeds:
 *   sum_d[b]         = sum_j min_i d(x_j, y_b,i)                    (double accumulation in a fixed order -> float)
 *   acc[b][i][0..2] += sum over the queries candidate i won of x_j * 2^fix_shift (64-bit two's complement)
 *   acc[b][i][3]    += how many queries candidate i won             (integer atomics: order-independent, bitwise
 *                                                                    reproducible; replaces the [T,M] index pass of
 *                                                                    fpv_chamfer_bwd for this loss)
 * acc ([batches][M][4] uint64) must be zero on entry.  fpv_scene2body_grad turns it into d(sum_b g_b sum_d[b]) / d y.
 * fix_shift = fpv_fix_shift_for(max |query coordinate|, N) keeps every sum below 2^62.  A query whose winner distance
 * is not finite adds +inf / NaN to sum_d and nothing to acc. */
size_t fpv_nn_sphere_fused_workspace_bytes(int64_t batches, int64_t N);
int fpv_fix_shift_for(float max_abs_coordinate, int64_t count);
int fpv_nn_sphere_fused(const float *queries, int64_t batches, int64_t N, const float *planes, const float *table,
                        const int32_t *orig_idx, const int32_t *pos_of, int pos_shared, int32_t *seed_inout,
                        int seed_valid, int64_t M, int tile, int fix_shift, float *sum_d, unsigned long long *acc,
                        unsigned long long *tiles_searched, void *workspace, size_t workspace_bytes,
                        fpv_stream_t stream);
int fpv_scene2body_grad(const float *cand, const unsigned long long *acc, int fix_shift, const float *g,
                        int64_t batches, int64_t M, float *grad, int accumulate, fpv_stream_t stream);

/* distChamfer(a, b) forward, reference output order (chamfer_python.py:28):
 *   d_b2a [bs,M], d_a2b [bs,N], i_b2a [bs,M] (index into a), i_a2b [bs,N] (index into b).
 * b_shared != 0: b is ONE [M,3] cloud for all bs frames (the reference materialises T copies at
 * global_optimization.py:176; pass the copy-free form here). */
size_t fpv_chamfer_fwd_workspace_bytes(int64_t bs, int64_t N, int64_t M, int b_shared);
int fpv_chamfer_fwd(const float *a, const float *b, int64_t bs, int64_t N, int64_t M, int b_shared,
                    float *d_b2a, float *d_a2b, void *i_b2a, void *i_a2b, int idx_bytes,
                    void *workspace, size_t workspace_bytes, fpv_stream_t stream);

/* distChamfer backward: gradient of sum(g_b2a*d_b2a) + sum(g_a2b*d_a2b) (either g may be NULL).
 *   grad_a [bs,N,3] is overwritten; grad_b ([bs,M,3], or [M,3] when b_shared) may be NULL.
 * The scatter through the argmin indices is a deterministic segmented reduction: contributions are
 * converted to 64-bit fixed point against a device-computed bound and summed with integer adds,
 * so the result does not depend on the order in which threads arrive. */
size_t fpv_chamfer_bwd_workspace_bytes(int64_t bs, int64_t N, int64_t M, int b_shared, int want_grad_b);
int fpv_chamfer_bwd(const float *a, const float *b, int64_t bs, int64_t N, int64_t M, int b_shared,
                    const float *g_b2a, const float *g_a2b, const void *i_b2a, const void *i_a2b,
                    int idx_bytes, float *grad_a, float *grad_b, void *workspace,
                    size_t workspace_bytes, fpv_stream_t stream);
/* Same, with g_broadcast bit 0 (g_b2a) / bit 1 (g_a2b) set when that weight tensor is ONE value shared by every
 * element -- what autograd hands back for a plain sum or mean (a stride-0 expanded scalar): the pointer then
 * addresses a single float and no [bs,M] weight array is materialised or read. */
int fpv_chamfer_bwd_bcast(const float *a, const float *b, int64_t bs, int64_t N, int64_t M, int b_shared,
                          const float *g_b2a, const float *g_a2b, int g_broadcast, const void *i_b2a,
                          const void *i_a2b, int idx_bytes, float *grad_a, float *grad_b, void *workspace,
                          size_t workspace_bytes, fpv_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Loss algebra around the chamfer term (global_optimization.py).
 * ------------------------------------------------------------------------------------------ */

/* :295  out[0] = mean( s/(s+1) ), s = sqrt(d+eps)  (the caller applies weight_contact). */
size_t fpv_reduce_workspace_bytes(int64_t n);
int fpv_robust_mean_fwd(const float *d, int64_t n, float eps, float *out, void *workspace,
                        size_t workspace_bytes, fpv_stream_t stream);
int fpv_robust_mean_bwd(const float *d, int64_t n, float eps, const float *g_out /*[1]*/,
                        float *grad_d, fpv_stream_t stream);

/* Temporal finite-difference L1 means over x [T,F] (F = features per frame):
 *   order 2: mean |(x_t - x_{t+1}) - (x_{t+1} - x_{t+2})|   :266-267, :381-382, :404-405
 *   order 1: mean |x_t - x_{t+1}|                             :304
 *   order 1 with frame_w [T]: mean |(x_t - x_{t+1}) * w_{t+1}|  :415-429
 * out[0] = the mean; *_bwd writes grad_x = g_out[0] * d(mean)/dx. */
int fpv_tdiff_l1_fwd(const float *x, int64_t T, int64_t F, int order, const float *frame_w,
                     float *out, void *workspace, size_t workspace_bytes, fpv_stream_t stream);
int fpv_tdiff_l1_bwd(const float *x, int64_t T, int64_t F, int order, const float *frame_w,
                     const float *g_out, float *grad_x, fpv_stream_t stream);

/* :119-127 verts_transform: out[t,p] = M_t[:3,:3] * v[t,p] + M_t[:3,3]   (M [T,4,4] row-major). */
int fpv_transform_fwd(const float *verts, const float *mats, int64_t T, int64_t P, float *out,
                      fpv_stream_t stream);
size_t fpv_transform_bwd_workspace_bytes(int64_t T, int64_t P);
int fpv_transform_bwd(const float *verts, const float *mats, const float *g_out, int64_t T,
                      int64_t P, float *g_verts, float *g_mats /*[T,4,4], may be NULL*/,
                      void *workspace, size_t workspace_bytes, fpv_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * SMPL-X body model forward / backward  (replaces the [3P] smplx forward called at
 * global_optimization.py:280-283, :333-335, :396-398; construction :154-168).
 * ------------------------------------------------------------------------------------------ */

#define FPV_SMPLX_JOINTS 55
#define FPV_SMPLX_POSE_FEAT 486
#define FPV_SMPLX_SHAPE 20
#define FPV_SMPLX_KPAD 512  /* 486 pose features + 20 shape/expression + 1 (template) + 5 zero */
#define FPV_SMPLX_THETA 122 /* per-frame parameter row, see below */

/* Per-frame parameter row theta[t] (122 floats), the arguments of SMPLX.forward concatenated:
 *   [0:3) global_orient  [3:66) body_pose  [66:69) jaw  [69:72) leye  [72:75) reye
 *   [75:87) left_hand PCA  [87:99) right_hand PCA  [99:109) betas  [109:119) expression
 *   [119:122) transl */

/* Device-resident constants, prepared once by the host mirror (body_model.py). */
typedef struct fpv_smplx_model {
    int32_t num_verts;      /* V */
    int32_t num_extra;      /* E: vertex-picked extra joints appended after the 55 */
    int32_t ell_width;      /* W: max skinning influences per vertex */
    int32_t reserved;
    /* blend basis [512][3V]: rows posedirs(486) | shapedirs^T(20) | v_template | 0, stored as TF32 hi/lo
     * splits in both operand orientations of the tcgen05 GEMMs.  P = 3V rounded up to a multiple of 4. */
    const float *basis_kn;                  /* unused (reserved) */
    const float *basis_nk_hi, *basis_nk_lo; /* [3V][512]  forward  B operand (K = 512 contiguous) */
    const float *basis_kn_hi, *basis_kn_lo; /* [512][P]   backward B operand (K = 3V contiguous, pitch P) */
    const float *j_template;   /* [55,3]    J_regressor @ v_template */
    const float *j_shapedirs;  /* [55,3,20] J_regressor @ shapedirs */
    const int32_t *parents;    /* [55] */
    const float *hand_comps;   /* [2,12,45] left | right PCA components */
    const float *pose_mean;    /* [165] */
    const int32_t *ell_joint;  /* [W][V] joint index per influence (-1 = none) */
    const float *ell_weight;   /* [W][V] */
    const int32_t *csr_ptr;    /* [56] per-joint influence lists (the transpose of ell) */
    const int32_t *csr_vert;   /* [nnz] */
    const float *csr_weight;   /* [nnz] */
    const int32_t *extra_vertex_ids; /* [E] */
} fpv_smplx_model_t;

/* Bytes of the per-call state kept between forward and backward, and of scratch. */
size_t fpv_smplx_saved_bytes(const fpv_smplx_model_t *model_host, int64_t T);
size_t fpv_smplx_workspace_bytes(const fpv_smplx_model_t *model_host, int64_t T);
/* vertices [T,V,3], joints [T,55+E,3] (both include transl, as SMPLX.forward returns them). */
int fpv_smplx_fwd(const fpv_smplx_model_t *model_host, int64_t T, const float *theta,
                  float *vertices, float *joints, void *saved, void *workspace,
                  size_t workspace_bytes, fpv_stream_t stream);
/* g_vertices / g_joints may be NULL (= zeros); g_theta [T,122] is overwritten. */
int fpv_smplx_bwd(const fpv_smplx_model_t *model_host, int64_t T, const float *theta,
                  const void *saved, const float *g_vertices, const float *g_joints,
                  float *g_theta, void *workspace, size_t workspace_bytes, fpv_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Tensor-core GEMM with fp32-level accuracy (3xTF32 on tcgen05), the contraction engine of the SMPL-X
 * blend shapes:  C[M,N] = sum_k A[m,k] * B[n,k], A and B row-major with K contiguous, given as TF32
 * hi/lo splits (fpv_split_tf32).  Row pitches lda/ldb (floats) must be multiples of 4 and the bases 16-byte
 * aligned (TMA).  ksplit > 1 slices K across CTAs; partials are reduced in fixed order (deterministic).
 * ------------------------------------------------------------------------------------------ */
int fpv_split_tf32(const float *src, int64_t n, float *hi, float *lo, fpv_stream_t stream);
size_t fpv_tc_gemm_workspace_bytes(int M, int N, int ksplit);
int fpv_tc_gemm_3xtf32(const float *a_hi, const float *a_lo, int64_t lda, const float *b_hi,
                       const float *b_lo, int64_t ldb, int M, int N, int K, float *C, int64_t ldc,
                       int ksplit, void *workspace, size_t workspace_bytes, fpv_stream_t stream);

/* ---- parameter front-end and trajectory prior -------------------------
 * rot6d <-> axis-angle: convert_to_3D_rot / convert_to_6D_rot (global_optimization.py:96-115) on n rotations.
 * in6 / out6 [n][6] = the (3,2) matrix of the first two rotation columns, row-major, as cvae.py:72 views it. */
int fpv_rot6d_to_aa_fwd(const float *in6, int64_t n, float *aa, fpv_stream_t stream);
int fpv_rot6d_to_aa_bwd(const float *in6, int64_t n, const float *g_aa, float *g_in6, fpv_stream_t stream);
int fpv_aa_to_rot6d(const float *aa, int64_t n, float *out6, fpv_stream_t stream);

/* VPoser v1 decoder, output_type='aa' (call site global_optimization.py:270-271).  Device pointers to the three
 * nn.Linear layers in their own [out][in] layout (w*) and transposed [in][out] copies (w*t). */
typedef struct fpv_vposer_model {
    const float *w1, *b1, *w2, *b2, *w3, *b3; /* [hidden][latent], [hidden][hidden], [6*joints][hidden] */
    const float *w1t, *w2t, *w3t;             /* [latent][hidden], [hidden][hidden], [hidden][6*joints] */
    int latent, hidden, joints;               /* 32, 512, 21 */
} fpv_vposer_model;
size_t fpv_vposer_saved_floats(const fpv_vposer_model *m, int64_t T);
/* z [T][latent] -> aa [T][joints][3]; saved: fpv_vposer_saved_floats(m, T) floats kept for the backward */
int fpv_vposer_decode_fwd(const fpv_vposer_model *m, const float *z, int64_t T, float *aa, float *saved,
                          fpv_stream_t stream);
int fpv_vposer_decode_bwd(const fpv_vposer_model *m, const float *saved, const float *g_aa, int64_t T, float *g_z,
                          fpv_stream_t stream);

/* DCT trajectory prior, FittingOP.cal_dctloss (global_optimization.py:232-246): x [NB*F][C] trajectories (C = 23
 * joints x 3 axes), basis [F][K], coef [NB][C][K]; out = mean over NB*C of sum_f e/(e+1), e = (x - basis.coef)^2.
 * grad_x / grad_coef may be NULL. */
size_t fpv_dct_prior_workspace_bytes(int64_t NB, int64_t F, int64_t C);
int fpv_dct_prior_fwd(const float *x, const float *basis, const float *coef, int64_t NB, int64_t F, int64_t C,
                      int64_t K, float *out, void *workspace, size_t workspace_bytes, fpv_stream_t stream);
int fpv_dct_prior_bwd(const float *x, const float *basis, const float *coef, int64_t NB, int64_t F, int64_t C,
                      int64_t K, const float *g_out, float *grad_x, float *grad_coef, fpv_stream_t stream);

/* The optimiser update of the reference loop (torch.optim.Adam, global_optimization.py:188, :592) as capturable
 * kernels: the step counter is a device float, so a CUDA graph that ends with the update replays correctly. */
int fpv_adam_tick(float *step, fpv_stream_t stream);
int fpv_adam_update(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr,
                    float beta1, float beta2, float eps, const float *step, fpv_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Peer-memory mailboxes of the scene-sharded step (p2p.cu).  One process per GPU; every rank
 * allocates one mailbox, exports its CUDA IPC handle, and maps the other ranks' mailboxes.  Keys are pushed by the
 * search epilogue (fpv_nn_culled_search_keys), combined by fpv_p2p_min_unpack after fpv_p2p_barrier; the small
 * parameter-gradient vector goes through fpv_p2p_push / fpv_p2p_sum.  No NCCL, no host synchronisation: capturable.
 * Pointers named *_host are HOST arrays of device pointers.
 * ------------------------------------------------------------------------------------------ */
int fpv_p2p_alloc(size_t bytes, void **ptr_host);
int fpv_p2p_free(void *ptr);
int fpv_p2p_export(void *ptr, unsigned char *handle64_host);
int fpv_p2p_open(const unsigned char *handle64_host, void **peer_ptr_host);
int fpv_p2p_close(void *peer_ptr);
int fpv_p2p_barrier(uint32_t *const *flags_host, int rank, int world, uint32_t *epoch, uint32_t *error,
                    double timeout_s, fpv_stream_t stream);
int fpv_p2p_min_unpack(const uint64_t *slots, int world, int64_t slot_stride, int64_t half_stride, uint32_t *parity,
                       int flip, int64_t n, int64_t row, const long long *perm, int perm_batched, float *dist, void *idx,
                       int idx_bytes, uint64_t *keys_out, fpv_stream_t stream);
int fpv_p2p_push(const float *src, int64_t n, float *const *dst_host, int n_dst, const uint32_t *parity,
                 int64_t half_stride, fpv_stream_t stream);
int fpv_p2p_gather(const float *slots, int world, int64_t slot_stride, int64_t half_stride, uint32_t *parity, int flip,
                   const int64_t *offset_host, const int64_t *count_host, float *out, fpv_stream_t stream);
int fpv_p2p_sum(const float *slots, int world, int64_t slot_stride, int64_t half_stride, uint32_t *parity, int flip,
                int64_t n, float *out, fpv_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FPV_B200_H */
