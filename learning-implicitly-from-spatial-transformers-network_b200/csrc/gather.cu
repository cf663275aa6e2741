// Fused transform-and-gather (rows a-2..a-5 of SURVEY.md §8): for every query point
//   * localise it with the image's 4x3 spatial-transformer matrix, perspective divide, clamp
//     (reference network/modules.py:37-47),
//   * bilinearly sample the channels-last upsampled maps (modules.py:48-53),
//   * add the 7 displacements and trilinearly sample the 6 channels-last voxel volumes
//     (modules.py:256-265),
//   * write ONE row of the fc_0 input matrix X[n][k_pad] directly in the kernel's column
//     layout (ListLayout) -- no cat / reshape / cat materialisations (modules.py:270-275).
//
// Mapping: a CTA owns P consecutive points (z-adjacent on dense grids, so their taps share
// cache lines).  Work items are (point, displacement, 8-channel vector); consecutive lanes take
// consecutive channel vectors of the same tap, so every tap read is a contiguous 16 B-per-lane
// segment of the channels-last tensor and every store is a contiguous piece of the point's row.
#include "common.cuh"

namespace list {

struct GatherParams {
  const void* maps;
  const void* vols[LIST_MAX_LEVELS];
  const float* T;        // [B][12]
  const float* q;        // [B][N][3] or nullptr in grid mode
  void* X;
  int64_t ldx;           // row stride in elements
  int64_t N;             // points per image
  int S, Cm;
  int R[LIST_MAX_LEVELS], C[LIST_MAX_LEVELS], voff[LIST_MAX_LEVELS];
  int nlev;
  int map_off, xyz_off, k_out, k_pad;
  int q_raw;             // 1: q is raw (x,y,z) -> apply [2,1,0]*2 (reference models.py:91-92)
  // grid mode (q == nullptr): points are res^3 grid indices begin.. of image `image`
  int grid_res;
  int image;
  int64_t grid_begin;
  double bb_min, bb_max;
};

constexpr int kP = 32;          // points per CTA
constexpr int kThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kThreads) gather_fwd_kernel(const GatherParams p) {
  __shared__ float s_q[kP][3];
  __shared__ float s_uv[kP][2];
  const int tid = threadIdx.x;
  const int b = p.q ? blockIdx.y : p.image;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * kP;
  const int npts = static_cast<int>(min64(kP, p.N - n0));
  T* __restrict__ Xb = static_cast<T*>(p.X) + (static_cast<int64_t>(p.q ? blockIdx.y : 0) * p.N + n0) * p.ldx;

  // ---- phase 0: query point -> swapped/scaled frame, 2-D sample position ----
  if (tid < kP) {
    float q[3] = {0.f, 0.f, 0.f};
    if (tid < npts) {
      float raw[3];
      int is_raw = p.q_raw;
      if (p.q) {
        const float* src = p.q + (static_cast<int64_t>(b) * p.N + n0 + tid) * 3;
        raw[0] = __ldg(src); raw[1] = __ldg(src + 1); raw[2] = __ldg(src + 2);
      } else {
        const int64_t g = p.grid_begin + n0 + tid;     // x slowest, z fastest (utils.py:88-93)
        const int res = p.grid_res;
        raw[0] = linspace_f32(static_cast<int>(g / (static_cast<int64_t>(res) * res)), res, p.bb_min, p.bb_max);
        raw[1] = linspace_f32(static_cast<int>((g / res) % res), res, p.bb_min, p.bb_max);
        raw[2] = linspace_f32(static_cast<int>(g % res), res, p.bb_min, p.bb_max);
        is_raw = 1;
      }
      if (is_raw) {
        q[0] = raw[2] * 2.0f; q[1] = raw[1] * 2.0f; q[2] = raw[0] * 2.0f;
      } else {
        q[0] = raw[0]; q[1] = raw[1]; q[2] = raw[2];
      }
    }
    float ix, iy, h[3];
    localise(q, p.T + b * 12, p.S, ix, iy, h);
    s_q[tid][0] = q[0]; s_q[tid][1] = q[1]; s_q[tid][2] = q[2];
    s_uv[tid][0] = ix; s_uv[tid][1] = iy;
  }
  __syncthreads();

  // ---- phase 1: 2-D bilinear taps, zeros padding (a-3) ----
  {
    const int ncv = p.Cm >> 3;
    const T* __restrict__ maps = static_cast<const T*>(p.maps) + static_cast<size_t>(b) * p.S * p.S * p.Cm;
    for (int item = tid; item < npts * ncv; item += kThreads) {
      const int pt = item / ncv;
      const int cv = item - pt * ncv;
      const float ix = s_uv[pt][0], iy = s_uv[pt][1];
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (ix == ix && iy == iy) {   // NaN grid -> every tap out of bounds (ATen CUDA sampler)
        const float fx = floorf(ix), fy = floorf(iy);
        const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
        const float wx1 = ix - fx, wx0 = (fx + 1.0f) - ix;
        const float wy1 = iy - fy, wy0 = (fy + 1.0f) - iy;
        const int x1 = x0 + 1, y1 = y0 + 1;
        const int lim = p.S - 1;
        const bool okx1 = x1 <= lim, oky1 = y1 <= lim;       // x0,y0 are in [0,lim] after the clamp
        const float w[4] = {wx0 * wy0, okx1 ? wx1 * wy0 : 0.f, oky1 ? wx0 * wy1 : 0.f,
                            (okx1 && oky1) ? wx1 * wy1 : 0.f};
        const int xs[2] = {x0, min(x1, lim)}, ys[2] = {y0, min(y1, lim)};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float v[8];
          load8(maps + (static_cast<size_t>(ys[t >> 1]) * p.S + xs[t & 1]) * p.Cm + cv * 8, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(v[j], w[t], acc[j]);
        }
      }
      store8(Xb + static_cast<int64_t>(pt) * p.ldx + p.map_off + cv * 8, acc);
    }
  }

  // ---- phase 2: 3-D trilinear taps, border padding, 7 displacements (a-4, a-5) ----
  for (int l = 0; l < p.nlev; ++l) {
    const int C = p.C[l];
    if (C & 7) continue;   // scalar levels handled in phase 3
    const int R = p.R[l];
    const int ncv = C >> 3;
    const int per_pt = LIST_NUM_DISP * ncv;
    const T* __restrict__ vol = static_cast<const T*>(p.vols[l]) + static_cast<size_t>(b) * R * R * R * C;
    const int voff = p.voff[l];
    for (int item = tid; item < npts * per_pt; item += kThreads) {
      const int pt = item / per_pt;
      const int r = item - pt * per_pt;
      const int d = r / ncv;
      const int cv = r - d * ncv;
      const float q[3] = {s_q[pt][0], s_q[pt][1], s_q[pt][2]};
      float pd[3];
      displaced(q, d, pd);
      const Axis3 ax = axis_border(pd[0], R);   // -> W
      const Axis3 ay = axis_border(pd[1], R);   // -> H
      const Axis3 az = axis_border(pd[2], R);   // -> D
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const int zi[2] = {az.i0, az.i1}, yi[2] = {ay.i0, ay.i1}, xi[2] = {ax.i0, ax.i1};
      const float wz[2] = {az.w0, az.w1}, wy[2] = {ay.w0, ay.w1}, wx[2] = {ax.w0, ax.w1};
#pragma unroll
      for (int t = 0; t < 8; ++t) {   // tnw,tne,tsw,tse,bnw,bne,bsw,bse (ATen order)
        const int tz = t >> 2, ty = (t >> 1) & 1, tx = t & 1;
        const float w = (wx[tx] * wy[ty]) * wz[tz];
        float v[8];
        load8(vol + ((static_cast<size_t>(zi[tz]) * R + yi[ty]) * R + xi[tx]) * C + cv * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(v[j], w, acc[j]);
      }
      store8(Xb + static_cast<int64_t>(pt) * p.ldx + voff + d * C + cv * 8, acc);
    }
  }

  // ---- phase 3: scalar levels (C not a multiple of 8, i.e. the 1-channel occupancy volume),
  //      the query coordinates themselves (modules.py:257,275) and the zero padding ----
  {
    int tail0 = p.xyz_off;
    for (int l = 0; l < p.nlev; ++l)
      if (p.C[l] & 7) tail0 = min(tail0, p.voff[l]);
    const int ntail = p.k_pad - tail0;
    for (int item = tid; item < npts * ntail; item += kThreads) {
      const int pt = item / ntail;
      const int col = tail0 + (item - pt * ntail);
      float val = 0.f;
      if (col >= p.xyz_off) {
        if (col < p.xyz_off + 3) val = s_q[pt][col - p.xyz_off];
      } else {
        for (int l = 0; l < p.nlev; ++l) {
          const int C = p.C[l];
          if (!(C & 7)) continue;
          const int rel = col - p.voff[l];
          if (rel < 0 || rel >= LIST_NUM_DISP * C) continue;
          const int d = rel / C, c = rel - d * C;
          const int R = p.R[l];
          const T* __restrict__ vol = static_cast<const T*>(p.vols[l]) + static_cast<size_t>(b) * R * R * R * C;
          const float q[3] = {s_q[pt][0], s_q[pt][1], s_q[pt][2]};
          float pd[3];
          displaced(q, d, pd);
          const Axis3 ax = axis_border(pd[0], R), ay = axis_border(pd[1], R), az = axis_border(pd[2], R);
          const int zi[2] = {az.i0, az.i1}, yi[2] = {ay.i0, ay.i1}, xi[2] = {ax.i0, ax.i1};
          const float wz[2] = {az.w0, az.w1}, wy[2] = {ay.w0, ay.w1}, wx[2] = {ax.w0, ax.w1};
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int tz = t >> 2, ty = (t >> 1) & 1, tx = t & 1;
            const float w = (wx[tx] * wy[ty]) * wz[tz];
            val = fmaf(to_f32(vol[((static_cast<size_t>(zi[tz]) * R + yi[ty]) * R + xi[tx]) * C + c]), w, val);
          }
        }
      }
      T o;
      from_f32(o, val);
      Xb[static_cast<int64_t>(pt) * p.ldx + col] = o;
    }
  }
}

// a-8: utils.create_grid_points_from_bounds rows (x,y,z), fp32.
__global__ void grid_points_kernel(float* __restrict__ q, int res, double lo, double hi, int64_t begin,
                                   int64_t count) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int64_t g = begin + i;
  q[i * 3 + 0] = linspace_f32(static_cast<int>(g / (static_cast<int64_t>(res) * res)), res, lo, hi);
  q[i * 3 + 1] = linspace_f32(static_cast<int>((g / res) % res), res, lo, hi);
  q[i * 3 + 2] = linspace_f32(static_cast<int>(g % res), res, lo, hi);
}

int fill_gather_params(const ListCtx* ctx, GatherParams& p) {
  ListLayout lay;
  const int rc = list_feature_layout(ctx->map_channels, ctx->n_levels, ctx->vol_ch, &lay, nullptr);
  if (rc) return rc;
  p.maps = ctx->maps;
  p.T = ctx->trans_mat;
  p.S = ctx->map_size;
  p.Cm = ctx->map_channels;
  p.nlev = ctx->n_levels;
  for (int l = 0; l < ctx->n_levels; ++l) {
    p.vols[l] = ctx->vols[l];
    p.R[l] = ctx->vol_res[l];
    p.C[l] = ctx->vol_ch[l];
    p.voff[l] = lay.vol_off[l];
  }
  p.map_off = lay.map_off;
  p.xyz_off = lay.xyz_off;
  p.k_out = lay.k_out;
  p.k_pad = lay.k_pad;
  return LIST_OK;
}

int gather_fwd(const ListCtx* ctx, const float* q, int q_is_raw, void* X, int64_t ldx, int B, int64_t N,
               cudaStream_t st) {
  GatherParams p{};
  const int rc = fill_gather_params(ctx, p);
  if (rc) return rc;
  p.q = q;
  p.q_raw = q_is_raw;
  p.X = X;
  p.ldx = ldx;
  p.N = N;
  if (N == 0 || B == 0) return LIST_OK;
  dim3 grid(static_cast<unsigned>((N + kP - 1) / kP), B);
  if (ctx->dtype == LIST_F32) gather_fwd_kernel<float><<<grid, kThreads, 0, st>>>(p);
  else gather_fwd_kernel<__nv_bfloat16><<<grid, kThreads, 0, st>>>(p);
  LIST_LAUNCH_CHECK("gather_fwd_kernel");
  return LIST_OK;
}

int gather_grid_fwd(const ListCtx* ctx, int image, int res, double bb_min, double bb_max, int64_t begin,
                    int64_t count, void* X, int64_t ldx, cudaStream_t st) {
  GatherParams p{};
  const int rc = fill_gather_params(ctx, p);
  if (rc) return rc;
  p.q = nullptr;
  p.q_raw = 1;
  p.X = X;
  p.ldx = ldx;
  p.N = count;
  p.grid_res = res;
  p.image = image;
  p.grid_begin = begin;
  p.bb_min = bb_min;
  p.bb_max = bb_max;
  if (count == 0) return LIST_OK;
  dim3 grid(static_cast<unsigned>((count + kP - 1) / kP), 1);
  if (ctx->dtype == LIST_F32) gather_fwd_kernel<float><<<grid, kThreads, 0, st>>>(p);
  else gather_fwd_kernel<__nv_bfloat16><<<grid, kThreads, 0, st>>>(p);
  LIST_LAUNCH_CHECK("gather_fwd_kernel(grid)");
  return LIST_OK;
}

int grid_points(float* q, int res, double lo, double hi, int64_t begin, int64_t count, cudaStream_t st) {
  if (count == 0) return LIST_OK;
  grid_points_kernel<<<static_cast<unsigned>((count + 255) / 256), 256, 0, st>>>(q, res, lo, hi, begin, count);
  LIST_LAUNCH_CHECK("grid_points_kernel");
  return LIST_OK;
}

}  // namespace list
