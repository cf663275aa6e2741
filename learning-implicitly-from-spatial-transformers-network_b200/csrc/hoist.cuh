// Host-side interface of hoist.cu / lines.cu (hoisted fc_0 for dense grids, bf16 mode), shared with api.cu and grid_tc.cu.
#pragma once
#include "common.cuh"

namespace list {
namespace hoist {

constexpr int kMaxH = 2;            // hoisted voxel levels of the addend-kernel path (hoist_addend_kernel keeps them in registers)
constexpr int kMaxLev = 3;          // hoisted voxel levels of the line-table path (lines.cu + grid_tc.cu)

struct Plan {                       // what is hoisted and where it lives in the caller's buffer
  int nh;
  int lev[kMaxLev];
  int hoist_cols;                   // leading columns of the full row replaced by the addend
  int k_h;                          // hoisted row width (multiple of 64): 512 addend columns + the remaining ones
  size_t off_pmap, off_pvol[kMaxLev], off_zero, off_w0r, total;
  // fc_0's bias rides in the MMA of the line-table path: the non-hoisted row Xr carries 1.0 in three of its pad columns
  // [bias_col, +3) (relative to Xr) and the per-call weight copy W0r [512][k_h - 512] at off_w0r carries the bf16
  // hi / mid / lo parts of b0 there (their sum is b0 to fp32 precision), so the epilogue is a bare ReLU + convert
  int bias_col;
  // line tables (lines.cu): per z-line `rpl` rows of 512 bf16; level h starts at row rowbase[h], class c at + c * R
  int rpl, rowbase[kMaxLev];
};

// max_levels / max_res: how many coarse levels may be hoisted and up to which resolution (2 / 16 for the addend-kernel
// path, 3 / 32 for the line-table path)
int make_plan(const ListCtx* ctx, const ListWeights* w, Plan* pl, int max_levels = kMaxH, int max_res = 16);
int check_gather(const ListCtx* ctx, const Plan& pl, int res);
int prepare(const ListCtx* ctx, const ListWeights* w, const Plan& pl, void* buf, cudaStream_t st);
int gather(const ListCtx* ctx, const ListWeights* w, const Plan& pl, const void* buf, int image, int res, double bb_min, double bb_max,
           int64_t begin, int64_t count, void* X, int64_t ldx, int parts, cudaStream_t st);
constexpr int kPartAddend = 1, kPartRest = 2;      // `parts` bit mask: which of the two gather kernels to launch
constexpr int kPartOnes = 4;                       // with kPartRest: 1.0 in the bias columns [bias_col, +3) of the row

// lines.cu: per-line column tables G[line - line0][rpl][512] of the hoisted levels for the z-lines touched by
// grid points [begin, begin + count) of image `image`
size_t lines_bytes(const Plan& pl, int res, int64_t begin, int64_t count);
int lines(const ListCtx* ctx, const Plan& pl, const void* buf, int image, int res, double bb_min, double bb_max, int64_t begin,
          int64_t count, void* G, cudaStream_t st);

}  // namespace hoist

// grid_tc.cu.  grid_plan: per tile of 128 steps the list of source rows (rows of the projected map and of the line tables G)
// and the interpolation weights of every step; grid_tc_fwd: fused interpolation + MLP over the rows of grid points
// [begin, begin + count): Xr = the non-hoisted columns [count][ldx] (hoist::gather with kPartRest on a pointer shifted by
// -512 columns), plan_buf = grid_plan of the same range (its rows point into G = hoist::lines of the same range).
size_t grid_plan_bytes(int res, int64_t begin, int64_t count);
int grid_plan(const ListCtx* ctx, const hoist::Plan& pl, const void* hoist_buf, int image, int res, double bb_min, double bb_max,
              int64_t begin, int64_t count, const void* G, void* plan_buf, cudaStream_t st);
int grid_tc_fwd(const ListCtx* ctx, const ListWeights* w, const hoist::Plan& pl, const void* hoist_buf, int res, double bb_min, double bb_max, int64_t begin,
                int64_t count, const void* Xr, int64_t ldx, const void* plan_buf, float* sdf, float out_div, float* dbg1,
                long long* trace, unsigned long long* stats, cudaStream_t st);

}  // namespace list
