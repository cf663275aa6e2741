/*
 * list_b200.h -- C ABI of the B200-native LIST per-query SDF hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI of its
 * own -- its hot path is a sequence of ATen calls inside two nn.Modules -- so every
 * entry point below cites the reference lines (relative to the reference repo root)
 * whose work it replaces.  The Python mirror of the reference's module API
 * (list_b200/network/{modules,models,executors}.py) binds these with ctypes;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are DEVICE pointers unless the
 *    name ends in _host; `stream` is a cudaStream_t passed as void*.
 *  - the caller owns every buffer (including workspaces); the library allocates
 *    nothing, keeps no pointer after return, and never synchronises: all work is
 *    enqueued on `stream`.
 *  - return 0 on success, a negative errno-style code otherwise; the message is
 *    available through list_b200_last_error() (thread-local).  Nothing throws.
 *  - re-entrant: no global mutable state besides the thread-local error string
 *    (nn.DataParallel replicas call from parallel host threads, reference
 *    train.py:126).
 *  - there is no CPU fallback and no other backend: sm_100a only.
 */
#ifndef LIST_B200_H
#define LIST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define LIST_API __attribute__((visibility("default")))
#else
#define LIST_API
#endif

#define LIST_B200_ABI_VERSION 3
#define LIST_MAX_LEVELS 8
#define LIST_MAX_MAPS 8
#define LIST_NUM_DISP 7          /* reference network/modules.py:205-212 */

enum { LIST_F32 = 0, LIST_BF16 = 1 };

enum {
  LIST_OK = 0,
  LIST_EINVAL = -22,             /* bad argument (shape, alignment, dtype) */
  LIST_ENOMEM = -12,             /* caller-provided workspace too small    */
  LIST_ECUDA = -5,               /* a CUDA call / launch failed            */
  LIST_ENOSYS = -38              /* not supported on this device           */
};

/* Per-image tensors the hot path reads, in the kernel layouts (channels-last).
 * Produced by list_prep_maps / list_prep_volume from the reference-layout tensors. */
typedef struct ListCtx {
  int32_t B;                     /* images                                                  */
  int32_t dtype;                 /* LIST_F32 | LIST_BF16: storage type of maps and vols      */
  int32_t map_size;              /* 137 (reference modules.py:16)                            */
  int32_t map_channels;          /* 1024 = 64+64+128+256+512, order f1..f5 (modules.py:53)   */
  const void* maps;              /* [B][map_size][map_size][map_channels]                    */
  int32_t n_levels;              /* 6                                                        */
  int32_t reserved0;
  int32_t vol_res[LIST_MAX_LEVELS];   /* 128,128,64,32,16,8                                  */
  int32_t vol_ch[LIST_MAX_LEVELS];    /* 1,16,32,64,128,128                                  */
  const void* vols[LIST_MAX_LEVELS];  /* [B][R][R][R][C] (D,H,W,C)                           */
  const float* trans_mat;        /* [B][4][3] fp32 (reference models.py:86)                  */
} ListCtx;

/* Column layout of one feature row X[n][0..k_pad) written by the gather kernels.
 * The reference order is [vox(c*7+d) | percep | q] (modules.py:272-275); the kernels use
 * [percep | level l, displacement d, channel c ... | scalar levels | q | zero pad] so that
 * every segment is 16-byte aligned.  perm[new] = reference column. */
typedef struct ListLayout {
  int32_t k_out;                 /* 3610                                                     */
  int32_t k_pad;                 /* k_out rounded up to 64 (3648)                            */
  int32_t map_off;               /* first column of the 2-D features                         */
  int32_t xyz_off;               /* first column of q                                        */
  int32_t vol_off[LIST_MAX_LEVELS];   /* first column of level l; +d*C+c inside              */
} ListLayout;

/* Kernel-format copies of the implicit MLP (reference modules.py:196-200); masters stay
 * nn.Parameters under the reference's state_dict keys in the Python wrapper. */
typedef struct ListWeights {
  int32_t dtype;                 /* LIST_F32 | LIST_BF16: type of w0,w1,w2                   */
  int32_t k_pad;                 /* columns of w0 (ListLayout.k_pad), zero padded            */
  int32_t n0, n1, n2;            /* 512, 256, 256                                            */
  int32_t reserved0;
  const void* w0;                /* [n0][k_pad]  fc_0, columns permuted by ListLayout        */
  const void* w1;                /* [n1][n0]     fc_1                                        */
  const void* w2;                /* [n2][n1]     fc_2                                        */
  const float* w3;               /* [n2]         fc_out (always fp32)                        */
  const float* b0;               /* [n0] */
  const float* b1;               /* [n1] */
  const float* b2;               /* [n2] */
  const float* b3;               /* [1]  */
} ListWeights;

/* Gradient buffers for list_sdf_bwd; all fp32, all ACCUMULATED INTO (caller zeroes). */
typedef struct ListGrads {
  float* d_maps;                 /* [B][S][S][Cm]        or NULL */
  float* d_vols[LIST_MAX_LEVELS];/* [B][R][R][R][C]      or NULL */
  float* d_trans_mat;            /* [B][4][3]            or NULL */
  float* d_w0;                   /* [n0][k_pad] (permuted columns) */
  float* d_w1;                   /* [n1][n0] */
  float* d_w2;                   /* [n2][n1] */
  float* d_w3;                   /* [n2] */
  float* d_b0; float* d_b1; float* d_b2; float* d_b3;
} ListGrads;

LIST_API int list_b200_abi_version(void);
LIST_API const char* list_b200_last_error(void);

/* Device check: 0 if the current device is sm_100 (B200), LIST_ENOSYS otherwise. */
LIST_API int list_b200_device_ok(void);

/* Column layout + permutation for a given channel configuration (host only, no CUDA).
 * perm may be NULL; otherwise it receives k_out entries, perm[new_col] = reference col. */
LIST_API int list_feature_layout(int32_t map_channels, int32_t n_levels, const int32_t* vol_ch,
                        ListLayout* layout, int32_t* perm);

/* a-1 hoisted (reference modules.py:25-35 + the per-chunk recomputation it implies):
 * n_maps NCHW fp32 maps -> one channels-last [B][S][S][sum C] tensor of `dtype`,
 * bilinear, align_corners=True.  Run once per image instead of once per chunk. */
LIST_API int list_prep_maps(const float* const* maps_nchw, const int32_t* ch, const int32_t* size,
                   int32_t n_maps, int32_t B, int32_t map_size, void* out, int32_t dtype,
                   void* stream);

/* NCDHW fp32 volume (reference modules.py:425-442 outputs) -> [B][R][R][R][C] of `dtype`. */
LIST_API int list_prep_volume(const float* vol_ncdhw, int32_t B, int32_t C, int32_t R, void* out,
                     int32_t dtype, void* stream);

/* a-9, adjoints of the two layout steps (autograd of reference modules.py:25-35 and of the voxel-encoder outputs feeding
 * modules.py:264-265), fp32: the channels-last gradients list_sdf_bwd produces -> gradients in the reference layout.
 *   list_prep_maps_bwd    g [B][S][S][sum ch] -> grads_nchw[i] [B][ch[i]][size[i]][size[i]] (bilinear align_corners adjoint,
 *                         gather form: deterministic, no atomics)
 *   list_prep_volume_bwd  g [B][R][R][R][C] -> [B][C][R][R][R] */
LIST_API int list_prep_maps_bwd(const float* g, const int32_t* ch, const int32_t* size, int32_t n_maps, int32_t B,
                       int32_t map_size, float* const* grads_nchw, void* stream);
LIST_API int list_prep_volume_bwd(const float* g, int32_t B, int32_t C, int32_t R, float* grad_ncdhw, void* stream);

/* a-8 grid (reference utils.py:84-95): points [begin, begin+count) of the res^3 grid,
 * x slowest / z fastest, float64 linspace rounded to fp32, written as (x,y,z) rows. */
LIST_API int list_grid_points(float* q, int32_t res, double bb_min, double bb_max, int64_t begin,
                     int64_t count, void* stream);

/* a-2..a-5 (reference modules.py:37-53, 256-275; models.py:91-92): one fused
 * transform-and-gather pass writing feature rows X[B*N][ldx] of ctx->dtype.
 * q: [B][N][3] fp32.  q_is_raw != 0: q holds raw (x,y,z) in [-0.5,0.5] and the kernel
 * applies the reference's [2,1,0] swap and *2; otherwise q is already swapped+scaled. */
LIST_API int list_gather_fwd(const ListCtx* ctx, const float* q, int32_t q_is_raw, void* X, int64_t ldx,
                    int32_t B, int64_t N, void* stream);

/* Same rows for grid points [begin, begin+count) of image `image` without reading q:
 * dense-grid specialisation of a-8 + a-2..a-5 (reference executors.py:215-220). */
LIST_API int list_gather_grid_fwd(const ListCtx* ctx, int32_t image, int32_t res, double bb_min,
                         double bb_max, int64_t begin, int64_t count, void* X, int64_t ldx,
                         void* stream);

/* a-6 (reference modules.py:276-282): sdf[r] = fc_out(relu(fc_2(relu(fc_1(relu(fc_0 X[r]))))))
 * for `rows` feature rows; result divided by out_div (1 = the reference's scaled SDF,
 * sdf_scale = executors.py:231).  LIST_BF16: tcgen05/TMEM kernel, needs no workspace.
 * LIST_F32: 3xTF32 tcgen05 GEMMs (fp32-level accuracy), workspace >= list_mlp_workspace_bytes(). */
LIST_API size_t list_mlp_workspace_bytes(const ListWeights* w, int64_t rows);
LIST_API int list_mlp_fwd(const ListWeights* w, const void* X, int64_t ldx, int64_t rows, float* sdf,
                 float out_div, void* workspace, size_t workspace_bytes, void* stream);
/* LIST_F32 forward of a TRAINING step (the activations it leaves in `workspace` feed list_sdf_bwd).  list_mlp_fwd computes
 * the fp32 mode on the tensor cores (three TF32 products per fp32 product, csrc/tgemm.cu), whose truncating accumulation
 * is inside the 1e-4 SDF bound but can flip the ReLU mask of units within 1e-5 of zero; this twin accumulates on the
 * FFMA pipe (round to nearest) so that the saved masks -- and with them the gradients -- match the fp32 reference. */
LIST_API int list_mlp_fwd_train(const ListWeights* w, const void* X, int64_t ldx, int64_t rows, float* sdf,
                       float out_div, void* workspace, size_t workspace_bytes, void* stream);

/* Diagnostic twin of list_mlp_fwd for the bf16 tensor-core kernel: additionally copies the hidden
 * activations relu(fc_0) [rows][n0], relu(fc_1) [rows][n1], relu(fc_2) [rows][n2] (fp32, before the
 * bf16 rounding the next layer sees) into h1/h2/h3 (each may be NULL) for layer-wise parity tests. */
LIST_API int list_mlp_fwd_debug(const ListWeights* w, const void* X, int64_t ldx, int64_t rows, float* sdf,
                       float out_div, float* h1, float* h2, float* h3, void* stream);

/* a-7: gather + MLP for explicit query points, chunked through `workspace`. */
LIST_API size_t list_sdf_workspace_bytes(const ListCtx* ctx, const ListWeights* w, int64_t chunk_rows);
LIST_API int list_sdf_fwd(const ListCtx* ctx, const ListWeights* w, const float* q, int32_t q_is_raw,
                 int32_t B, int64_t N, float* sdf, float out_div, int64_t chunk_rows,
                 void* workspace, size_t workspace_bytes, void* stream);

/* Hoisted fc_0 for dense grids in bf16 mode (csrc/hoist.cu).  fc_0 (reference modules.py:197,276) and the
 * samplers in front of it (modules.py:48-52, 264-265) are linear, so the maps and the coarse voxel levels
 * (R <= 16) are projected through their column blocks of W0 once per image and sampled per query as ONE
 * 512-wide addend block (which also carries the bias b0); the feature row shrinks from 3648 to
 * [addend 512 | remaining k_pad - hoist_cols columns] and fc_0 runs on the remaining columns only, the addend
 * being added to its accumulator in the epilogue.  list_sdf_grid does all of this internally; the calls below
 * expose the stages for tests and per-kernel timing.
 *   list_hoist_layout           hoist_cols (leading columns of the full row that are hoisted) and k_h (row width)
 *   list_hoist_bytes            size of the caller-owned buffer holding the projected tensors (0: not hoistable)
 *   list_hoist_prepare          projects every image of ctx
 *   list_hoist_gather_grid_fwd  hoisted rows X[count][ldx >= k_h] for grid points [begin, begin+count) of `image`;
 *                               parts: 1 = addend columns only, 2 = the remaining columns only, 3 = the whole row
 *   list_mlp_hoisted_fwd        a-6 on hoisted rows (w = the ORIGINAL weights) */
LIST_API int list_hoist_layout(const ListCtx* ctx, const ListWeights* w, int32_t* hoist_cols, int32_t* k_h);
LIST_API size_t list_hoist_bytes(const ListCtx* ctx, const ListWeights* w);
LIST_API int list_hoist_prepare(const ListCtx* ctx, const ListWeights* w, void* hoist_buf, size_t hoist_bytes, void* stream);
LIST_API int list_hoist_gather_grid_fwd(const ListCtx* ctx, const ListWeights* w, const void* hoist_buf, int32_t image,
                               int32_t res, double bb_min, double bb_max, int64_t begin, int64_t count,
                               void* X, int64_t ldx, int32_t parts, void* stream);
LIST_API int list_mlp_hoisted_fwd(const ListWeights* w, int32_t hoist_cols, const void* Xh, int64_t ldx, int64_t rows,
                         float* sdf, float out_div, void* stream);
/* Diagnostic twin: additionally records the phase timeline of CTA 0 -- trace[16 tiles][12] device int64 clock64() stamps:
 * 0 tile start, 1 fc_0 issued, 2 fc_1 start, 3 fc_1 issued, 4 fc_2 start, 5 fc_2 issued (MMA thread);
 * 6 fc_0 done seen, 7 ep0 done, 8 fc_1 done seen, 9 ep1 done, 10 fc_2 done seen, 11 ep2 done (epilogue warp). */
LIST_API int list_mlp_hoisted_trace(const ListWeights* w, int32_t hoist_cols, const void* Xh, int64_t ldx, int64_t rows,
                           float* sdf, float out_div, int64_t* trace, void* stream);

/* Line-table path for dense grids in bf16 mode (csrc/lines.cu + csrc/grid_tc.cu), the default of list_sdf_grid.
 * As above the maps and the coarse voxel levels (here R <= 32) are projected through their column blocks of W0 once per
 * image, but the 512-wide addend block is never materialised: along a z-line of the grid (reference utils.py:84-95) the
 * trilinear sample of a projected level is a two-tap interpolation in ONE column table per (level, W-shift class)
 * (list_lines_table), and the MLP kernel evaluates those interpolations and the bilinear taps of the projected map as
 * a small extra GEMM on the tensor cores, accumulating into fc_0's TMEM tile (list_grid_tc_fwd; reference
 * modules.py:48-53, 262-282).  list_sdf_grid runs the stages internally; they are exposed for tests and timing.
 *   list_lines_layout      hoist_cols, k_f = k_pad - hoist_cols (columns of the dense part Xr), rows per line table
 *   list_lines_hoist_bytes size of the caller-owned buffer with the projected tensors (0: configuration not covered)
 *   list_lines_prepare     projects every image of ctx into hoist_buf
 *   list_lines_table       G[lines touched by [begin, begin+count)][rows_per_line][512] bf16 of `image`
 *   list_lines_rest        Xr[count][ldx >= k_f]: the non-hoisted feature columns (fine levels, q, zero pad); the first three
 *                          pad columns hold 1.0: fc_0's bias rides in the MMA (list_lines_prepare writes a copy of the
 *                          non-hoisted weight columns with the bf16 hi / mid / lo parts of b0 in those columns)
 *   list_grid_plan         per tile of 128 consecutive steps of a z-line: the source rows its interpolation reads (rows of the
 *                          projected map and of G, as absolute addresses -- G is only used as an address here) and the <= 22
 *                          interpolation weights of every step; plan buffer >= list_grid_plan_bytes(), 256B aligned
 *   list_grid_tc_fwd       sdf[count] = MLP(Xr, interpolated hoisted terms) / out_div, from Xr, the plan (whose rows point
 *                          into hoist_buf and G: both must still be valid) and hoist_buf (weight copy).  dbg_h1 (or NULL): relu(fc_0) as fp32
 *                          [count][512]; trace (or NULL): device int64[16 tiles][24] clock64() timeline of CTA 0, slots
 *                          0-11 as in list_mlp_hoisted_trace, 13 interpolation chunks issued (MMA thread), 14 plan
 *                          loaded / 15 chunks filled / 16-23 phases of its first chunk (interp warp 0); stats (or NULL):
 *                          device uint64[2], += tile pairs processed, += 16-row k-steps of interpolation chunks issued (the padding tail
 *                          of a tile's last 64-row chunk is skipped) */
LIST_API int list_lines_layout(const ListCtx* ctx, const ListWeights* w, int32_t* hoist_cols, int32_t* k_f, int32_t* rows_per_line);
LIST_API size_t list_lines_hoist_bytes(const ListCtx* ctx, const ListWeights* w);
LIST_API int list_lines_prepare(const ListCtx* ctx, const ListWeights* w, void* hoist_buf, size_t hoist_bytes, void* stream);
LIST_API size_t list_lines_table_bytes(const ListCtx* ctx, const ListWeights* w, int32_t res, int64_t begin, int64_t count);
LIST_API int list_lines_table(const ListCtx* ctx, const ListWeights* w, const void* hoist_buf, int32_t image, int32_t res,
                     double bb_min, double bb_max, int64_t begin, int64_t count, void* G, size_t G_bytes, void* stream);
LIST_API int list_lines_rest(const ListCtx* ctx, const ListWeights* w, int32_t image, int32_t res, double bb_min, double bb_max,
                    int64_t begin, int64_t count, void* Xr, int64_t ldx, void* stream);
LIST_API size_t list_grid_plan_bytes(int32_t res, int64_t begin, int64_t count);
LIST_API int list_grid_plan(const ListCtx* ctx, const ListWeights* w, const void* hoist_buf, int32_t image, int32_t res,
                   double bb_min, double bb_max, int64_t begin, int64_t count, const void* G, void* plan, size_t plan_bytes,
                   void* stream);
LIST_API int list_grid_tc_fwd(const ListCtx* ctx, const ListWeights* w, const void* hoist_buf, int32_t res, double bb_min, double bb_max,
                     int64_t begin, int64_t count, const void* Xr, int64_t ldx, const void* plan, float* sdf, float out_div,
                     float* dbg_h1, int64_t* trace, uint64_t* stats, void* stream);

/* a-8 (reference executors.py:191-231): SDF of grid points [begin, begin+count) of every
 * image, sdf[B][count], divided by sdf_scale.  This is the per-rank shard of §8e.
 * Workspace: list_sdf_grid_workspace_bytes(ctx, w, res, chunk_rows) bytes -- what the path that runs for this
 * configuration and resolution needs (bf16 line-table path: two chunk buffers of ~1.6 KB per row + 0.3 GB of projected
 * tensors per image, e.g. 13 GB for 4 M-row chunks of a 256^3 grid); list_sdf_workspace_bytes(ctx, w, chunk_rows) is a
 * resolution-independent upper bound (full feature rows) and is accepted as well. */
LIST_API size_t list_sdf_grid_workspace_bytes(const ListCtx* ctx, const ListWeights* w, int32_t res, int64_t chunk_rows);
LIST_API int list_sdf_grid(const ListCtx* ctx, const ListWeights* w, int32_t res, double bb_min,
                  double bb_max, int64_t begin, int64_t count, float* sdf, float sdf_scale,
                  int64_t chunk_rows, void* workspace, size_t workspace_bytes, void* stream);

/* Staged variant for callers that are still producing the fine volumes (upload, all-gather ...) on another
 * stream: every level l with late_vols_ncdhw[l] != NULL is NOT yet valid in ctx->vols[l]; the call waits for
 * late_event (a cudaEvent_t recorded after the producer's last write, or NULL) right before its first kernel
 * that reads such a level -- the projection and the first chunk's line tables run before that -- and prepares
 * the level itself (list_prep_volume from the reference-layout fp32 DEVICE tensor late_vols_ncdhw[l] into
 * ctx->vols[l]).  sdf_host (pinned host memory, [B][count], or NULL) additionally receives every chunk's values
 * behind the next chunk's kernels.  Levels with R <= 32 and C % 64 == 0 may be read by the projection and cannot
 * be late (LIST_EINVAL).  late_vols_ncdhw == NULL and sdf_host == NULL: same as list_sdf_grid. */
LIST_API int list_sdf_grid_late(const ListCtx* ctx, const ListWeights* w, int32_t res, double bb_min,
                       double bb_max, int64_t begin, int64_t count, float* sdf, float sdf_scale,
                       int64_t chunk_rows, void* workspace, size_t workspace_bytes, void* stream,
                       void* late_event, const float* const* late_vols_ncdhw, float* sdf_host);

/* Same call with HOST buffers end to end (the e2e path of bench.py): reference-layout fp32
 * per-image tensors in (pinned) host memory -> H2D -> prep -> grid evaluation -> D2H of the
 * SDF grid.  dev_scratch is a caller-owned DEVICE arena of >= list_sdf_grid_host_bytes().
 * Transfers run on an internal copy stream, forked from and joined to `stream`: the fine
 * volumes upload while the projection and the first chunk's addend gather run, every chunk's
 * SDF values download behind the next chunk's kernels.  The host buffers must
 * stay untouched until the work enqueued on `stream` by this call is done; pinned memory is
 * needed for the overlap, not for correctness. */
LIST_API size_t list_sdf_grid_host_bytes(const int32_t* map_ch, const int32_t* map_size_in, int32_t n_maps,
                                int32_t map_size, int32_t n_levels, const int32_t* vol_ch,
                                const int32_t* vol_res, int32_t B, int32_t dtype, int64_t count,
                                int64_t chunk_rows);
LIST_API int list_sdf_grid_host(const float* const* maps_nchw_host, const int32_t* map_ch,
                       const int32_t* map_size_in, int32_t n_maps, int32_t map_size,
                       const float* const* vols_ncdhw_host, int32_t n_levels, const int32_t* vol_ch,
                       const int32_t* vol_res, const float* trans_mat_host, int32_t B, int32_t dtype,
                       const ListWeights* w_dev, int32_t res, double bb_min, double bb_max,
                       int64_t begin, int64_t count, float sdf_scale, int64_t chunk_rows,
                       float* sdf_host, void* dev_scratch, size_t dev_scratch_bytes, void* stream);

/* GPU marching cubes of a dense res^3 SDF grid (reference utils.py:172-182 hands the grid to
 * PyMCubes on the host: `mcubes.marching_cubes(-grid, 0)`).  Two calls because the output size is data dependent:
 *   list_mc_count     classifies cubes / grid edges, scans, and writes counts[0] = vertices, counts[1] = triangles
 *                     (DEVICE int64[2]; the caller reads them back and allocates the outputs)
 *   list_mc_generate  vertices[n][3] fp32 in index coordinates (i, j, k of sdf[i][j][k], like PyMCubes) -- one shared
 *                     vertex per crossed grid edge, ascending edge order -- and triangles[n][3] int32, ascending cube order
 * negate != 0 contours -sdf (what the reference passes).  workspace: >= list_mc_workspace_bytes(res), 256B aligned,
 * untouched between the two calls.  Case tables: csrc/mc_tables.h (scripts/gen_mc_tables.py). */
LIST_API size_t list_mc_workspace_bytes(int32_t res);
LIST_API int list_mc_count(const float* sdf, int32_t res, float iso, int32_t negate, void* workspace, size_t workspace_bytes,
                  int64_t* counts, void* stream);
LIST_API int list_mc_generate(const float* sdf, int32_t res, float iso, int32_t negate, const void* workspace,
                     size_t workspace_bytes, float* vertices, int64_t n_vertices, int32_t* triangles,
                     int64_t n_triangles, void* stream);

/* The GEMM of the fp32 path (a-6 in LIST_F32 mode and its backward; reference modules.py:276-280 and autograd): fp32-level
 * accuracy on the tensor cores through three TF32 products per fp32 product (csrc/tgemm.cu).  Exposed for tests:
 *   C[m][n] (+)= sum_k A[m*lda + k] B[n*ldb + k]   (both operands K-major; the backward transposes what comes otherwise)
 * lo_workspace: >= 4 * (M*lda + N*ldb) bytes (the low-order parts of both operands). */
LIST_API int list_gemm_f32_tc(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int32_t M, int32_t N,
                     int32_t K, int32_t accumulate, float* lo_workspace, size_t lo_bytes, void* stream);

/* a-9 backward of a-2..a-6 for training (reference train.py:72-85 via autograd):
 * given d_sdf[B*N] computes all gradients in ListGrads (fp32 path only).
 * Needs the fp32 feature rows X and the saved activations in `workspace` from
 * list_mlp_fwd on the same rows (same workspace pointer, untouched in between). */
LIST_API size_t list_bwd_workspace_bytes(const ListWeights* w, int64_t rows);
LIST_API int list_sdf_bwd(const ListCtx* ctx, const ListWeights* w, const float* q, int32_t q_is_raw,
                 int32_t B, int64_t N, const void* X, int64_t ldx, const void* fwd_workspace,
                 const float* d_sdf, const ListGrads* grads, void* workspace,
                 size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LIST_B200_H */
