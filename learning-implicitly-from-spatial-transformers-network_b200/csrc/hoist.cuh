// Host-side interface of hoist.cu (hoisted fc_0 for dense grids, bf16 mode), shared with api.cu.
#pragma once
#include "common.cuh"

namespace list {
namespace hoist {

constexpr int kMaxH = 2;            // hoisted voxel levels

struct Plan {                       // what is hoisted and where it lives in the caller's buffer
  int nh;
  int lev[kMaxH];
  int hoist_cols;                   // leading columns of the full row replaced by the addend
  int k_h;                          // hoisted row width (multiple of 64)
  size_t off_pmap, off_pvol[kMaxH], total;
};

int make_plan(const ListCtx* ctx, const ListWeights* w, Plan* pl);
int check_gather(const ListCtx* ctx, const Plan& pl, int res);
int prepare(const ListCtx* ctx, const ListWeights* w, const Plan& pl, void* buf, cudaStream_t st);
int gather(const ListCtx* ctx, const ListWeights* w, const Plan& pl, const void* buf, int image, int res, double bb_min, double bb_max,
           int64_t begin, int64_t count, void* X, int64_t ldx, int parts, cudaStream_t st);
constexpr int kPartAddend = 1, kPartRest = 2;      // `parts` bit mask: which of the two gather kernels to launch

}  // namespace hoist
}  // namespace list
