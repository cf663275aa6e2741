// Backward of the fused transform-and-gather (row a-9 of SURVEY.md §8; the reference gets it
// from autograd over modules.py:37-53 and :256-275): given dX[rows][k_pad] (fp32, from the MLP
// backward) scatter-add into
//   d_vols  : trilinear taps, grid_sampler_3d_backward (border padding: clamped corners receive
//             the weight of the corner they replaced, i.e. exactly the forward's taps),
//   d_maps  : bilinear taps into the channels-last upsampled maps (zeros padding),
//   d_trans : through the 2-D sample coordinates, the clamp (no gradient where clamped), the
//             perspective divide and the 4x3 product.
// No gradient flows to the query points (they are data) -- same as the reference.
// fp32 only; all accumulation with vector red.global.add.f32.
#include "common.cuh"

namespace list {

struct GatherBwdParams {
  const void* maps;
  const float* T;
  const float* q;
  const float* dX;
  int64_t ldd;
  int64_t N;
  float* d_maps;
  float* d_vols[LIST_MAX_LEVELS];
  float* d_T;
  int S, Cm;
  int R[LIST_MAX_LEVELS], C[LIST_MAX_LEVELS], voff[LIST_MAX_LEVELS];
  int nlev;
  int map_off, xyz_off;
  int q_raw;
};

constexpr int kPB = 32;

__device__ __forceinline__ void red_add8(float* dst, const float g[8], float w) {
  atomicAdd(reinterpret_cast<float4*>(dst), make_float4(g[0] * w, g[1] * w, g[2] * w, g[3] * w));
  atomicAdd(reinterpret_cast<float4*>(dst) + 1, make_float4(g[4] * w, g[5] * w, g[6] * w, g[7] * w));
}

__global__ void __launch_bounds__(256) gather_bwd_kernel(const GatherBwdParams p) {
  __shared__ float s_q[kPB][3];
  __shared__ float s_uv[kPB][2];
  __shared__ float s_h[kPB][3];
  __shared__ float s_g[kPB][2];
  __shared__ float s_dT[12];
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * kPB;
  const int npts = static_cast<int>(min64(kPB, p.N - n0));
  const float* __restrict__ dXb = p.dX + (static_cast<int64_t>(b) * p.N + n0) * p.ldd;

  if (tid < 12) s_dT[tid] = 0.f;
  if (tid < kPB) {
    float q[3] = {0.f, 0.f, 0.f};
    if (tid < npts) {
      const float* src = p.q + (static_cast<int64_t>(b) * p.N + n0 + tid) * 3;
      const float r0 = __ldg(src), r1 = __ldg(src + 1), r2 = __ldg(src + 2);
      if (p.q_raw) { q[0] = r2 * 2.0f; q[1] = r1 * 2.0f; q[2] = r0 * 2.0f; }
      else { q[0] = r0; q[1] = r1; q[2] = r2; }
    }
    float ix, iy, h[3];
    localise(q, p.T + b * 12, p.S, ix, iy, h);
    for (int j = 0; j < 3; ++j) { s_q[tid][j] = q[j]; s_h[tid][j] = h[j]; }
    s_uv[tid][0] = ix; s_uv[tid][1] = iy;
    s_g[tid][0] = 0.f; s_g[tid][1] = 0.f;
  }
  __syncthreads();

  // ---- 2-D taps: d_maps and the coordinate gradient ----
  if (p.d_maps || p.d_T) {
    const int ncv = p.Cm >> 3;
    const size_t img = static_cast<size_t>(b) * p.S * p.S * p.Cm;
    const float* __restrict__ maps = static_cast<const float*>(p.maps) + img;
    float* dmaps = p.d_maps ? p.d_maps + img : nullptr;
    for (int item = tid; item < npts * ncv; item += 256) {
      const int pt = item / ncv;
      const int cv = item - pt * ncv;
      const float ix = s_uv[pt][0], iy = s_uv[pt][1];
      if (!(ix == ix && iy == iy)) continue;
      float g[8];
      load8(dXb + static_cast<int64_t>(pt) * p.ldd + p.map_off + cv * 8, g);
      const float fx = floorf(ix), fy = floorf(iy);
      const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
      const float wx1 = ix - fx, wx0 = (fx + 1.0f) - ix;
      const float wy1 = iy - fy, wy0 = (fy + 1.0f) - iy;
      const int lim = p.S - 1;
      const bool okx1 = (x0 + 1) <= lim, oky1 = (y0 + 1) <= lim;
      const bool ok[4] = {true, okx1, oky1, okx1 && oky1};
      const float w[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};
      const int xs[2] = {x0, min(x0 + 1, lim)}, ys[2] = {y0, min(y0 + 1, lim)};
      float gix = 0.f, giy = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (!ok[t]) continue;
        const size_t off = (static_cast<size_t>(ys[t >> 1]) * p.S + xs[t & 1]) * p.Cm + cv * 8;
        if (dmaps) red_add8(dmaps + off, g, w[t]);
        if (p.d_T) {
          float v[8];
          load8(maps + off, v);
          float dot = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) dot = fmaf(v[j], g[j], dot);
          // d out / d ix = (ne-nw)*wy0 + (se-sw)*wy1 ; d out / d iy = (sw-nw)*wx0 + (se-ne)*wx1
          const float sx = (t & 1) ? 1.f : -1.f, sy = (t >> 1) ? 1.f : -1.f;
          gix += sx * dot * ((t >> 1) ? wy1 : wy0);
          giy += sy * dot * ((t & 1) ? wx1 : wx0);
        }
      }
      if (p.d_T) {
        atomicAdd(&s_g[pt][0], gix);
        atomicAdd(&s_g[pt][1], giy);
      }
    }
  }

  // ---- 3-D taps ----
  for (int l = 0; l < p.nlev; ++l) {
    float* dvol_base = p.d_vols[l];
    if (!dvol_base) continue;
    const int C = p.C[l], R = p.R[l];
    float* dvol = dvol_base + static_cast<size_t>(b) * R * R * R * C;
    const int voff = p.voff[l];
    if (!(C & 7)) {
      const int ncv = C >> 3;
      const int per_pt = LIST_NUM_DISP * ncv;
      for (int item = tid; item < npts * per_pt; item += 256) {
        const int pt = item / per_pt;
        const int r = item - pt * per_pt;
        const int d = r / ncv;
        const int cv = r - d * ncv;
        const float q[3] = {s_q[pt][0], s_q[pt][1], s_q[pt][2]};
        float pd[3];
        displaced(q, d, pd);
        const Axis3 ax = axis_border(pd[0], R), ay = axis_border(pd[1], R), az = axis_border(pd[2], R);
        const int zi[2] = {az.i0, az.i1}, yi[2] = {ay.i0, ay.i1}, xi[2] = {ax.i0, ax.i1};
        const float wz[2] = {az.w0, az.w1}, wy[2] = {ay.w0, ay.w1}, wx[2] = {ax.w0, ax.w1};
        float g[8];
        load8(dXb + static_cast<int64_t>(pt) * p.ldd + voff + d * C + cv * 8, g);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int tz = t >> 2, ty = (t >> 1) & 1, tx = t & 1;
          const float w = (wx[tx] * wy[ty]) * wz[tz];
          if (w != 0.f)
            red_add8(dvol + ((static_cast<size_t>(zi[tz]) * R + yi[ty]) * R + xi[tx]) * C + cv * 8, g, w);
        }
      }
    } else {
      const int per_pt = LIST_NUM_DISP * C;
      for (int item = tid; item < npts * per_pt; item += 256) {
        const int pt = item / per_pt;
        const int r = item - pt * per_pt;
        const int d = r / C, c = r - d * C;
        const float q[3] = {s_q[pt][0], s_q[pt][1], s_q[pt][2]};
        float pd[3];
        displaced(q, d, pd);
        const Axis3 ax = axis_border(pd[0], R), ay = axis_border(pd[1], R), az = axis_border(pd[2], R);
        const int zi[2] = {az.i0, az.i1}, yi[2] = {ay.i0, ay.i1}, xi[2] = {ax.i0, ax.i1};
        const float wz[2] = {az.w0, az.w1}, wy[2] = {ay.w0, ay.w1}, wx[2] = {ax.w0, ax.w1};
        const float g = __ldg(dXb + static_cast<int64_t>(pt) * p.ldd + voff + r);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int tz = t >> 2, ty = (t >> 1) & 1, tx = t & 1;
          const float w = (wx[tx] * wy[ty]) * wz[tz];
          if (w != 0.f)
            atomicAdd(dvol + ((static_cast<size_t>(zi[tz]) * R + yi[ty]) * R + xi[tx]) * C + c, g * w);
        }
      }
    }
  }

  // ---- d_trans_mat ----
  if (p.d_T) {
    __syncthreads();
    if (tid < npts) {
      const float lim = static_cast<float>(p.S - 1);
      const float den = s_h[tid][2] + 1e-8f;
      const float xu = s_h[tid][0] / den, yu = s_h[tid][1] / den;
      // d ix / d x = ((S-1)/2) / half = 1; clamp passes gradient only inside [0, S-1]
      const float dx = (xu >= 0.f && xu <= lim) ? s_g[tid][0] : 0.f;
      const float dy = (yu >= 0.f && yu <= lim) ? s_g[tid][1] : 0.f;
      const float dh[3] = {dx / den, dy / den, -(dx * xu + dy * yu) / den};
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (dh[j] != 0.f) {
          atomicAdd(&s_dT[0 * 3 + j], s_q[tid][0] * dh[j]);
          atomicAdd(&s_dT[1 * 3 + j], s_q[tid][1] * dh[j]);
          atomicAdd(&s_dT[2 * 3 + j], s_q[tid][2] * dh[j]);
          atomicAdd(&s_dT[3 * 3 + j], dh[j]);
        }
      }
    }
    __syncthreads();
    if (tid < 12 && s_dT[tid] != 0.f) atomicAdd(p.d_T + b * 12 + tid, s_dT[tid]);
  }
}

int gather_bwd(const ListCtx* ctx, const float* q, int q_is_raw, int B, int64_t N, const float* dX,
               int64_t ldd, const ListGrads* g, cudaStream_t st) {
  ListLayout lay;
  const int rc = list_feature_layout(ctx->map_channels, ctx->n_levels, ctx->vol_ch, &lay, nullptr);
  if (rc) return rc;
  GatherBwdParams p{};
  p.maps = ctx->maps;
  p.T = ctx->trans_mat;
  p.q = q;
  p.dX = dX;
  p.ldd = ldd;
  p.N = N;
  p.d_maps = g->d_maps;
  p.d_T = g->d_trans_mat;
  p.S = ctx->map_size;
  p.Cm = ctx->map_channels;
  p.nlev = ctx->n_levels;
  for (int l = 0; l < ctx->n_levels; ++l) {
    p.d_vols[l] = g->d_vols[l];
    p.R[l] = ctx->vol_res[l];
    p.C[l] = ctx->vol_ch[l];
    p.voff[l] = lay.vol_off[l];
  }
  p.map_off = lay.map_off;
  p.xyz_off = lay.xyz_off;
  p.q_raw = q_is_raw;
  if (N == 0 || B == 0) return LIST_OK;
  dim3 grid(static_cast<unsigned>((N + kPB - 1) / kPB), B);
  gather_bwd_kernel<<<grid, 256, 0, st>>>(p);
  LIST_LAUNCH_CHECK("gather_bwd_kernel");
  return LIST_OK;
}

}  // namespace list
