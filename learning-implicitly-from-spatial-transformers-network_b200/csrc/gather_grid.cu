// Dense-grid specialisation of the fused transform-and-gather (rows a-8 + a-2..a-5 of SURVEY.md
// §8; reference network/executors.py:215-220 driving modules.py:37-53 and :256-275).
//
// The reference's grid (utils.py:84-95) is x-slowest / z-fastest, so consecutive points form
// z-runs with fixed (x, y).  In the swapped/scaled frame only the W coordinate of the voxel taps
// changes along a run, the (H, D) corners and their weights are constant.  A thread therefore
// owns one 8-channel vector of one (level, displacement) pair and WALKS the run:
//     G[xv] = sum over the 4 (H,D) corners of (wy*wz) * V[z_k][y_k][xv][c..c+8)   (in registers for
//             xv = x0 and x1; when x0 advances, G0 <- G1 and only ONE new row is fetched)
//     out   = G0 + wx1 * (G1 - G0)
// i.e. the trilinear sum evaluated separably.  That cuts the 56 tap reads per (point, level,
// channel vector) of the generic kernel to ~4*R/res per step and the FMAs from 56 to ~1 + 4*R/res.
// The 2-D taps are cached the same way (the pixel cell changes every ~2-3 steps at 256^3).
// Results differ from the generic kernel only by fp32 re-association (~1e-7 relative); a point's
// value never depends on how the grid is chunked or sharded (bit-exact composition).
//
// Mapping: grid = (ceil(N/64), roles), 128 threads.  blockIdx.y selects a ROLE -- a group of walkers
// with similar cost (the 2-D vectors; one or more voxel levels; the scalar tail) -- so every CTA is
// homogeneous, finishes on its own, and the hardware scheduler balances the roles.  Per-step
// coordinates (voxel index + weight per level and W-shift class; pixel cell + 4 weights) are
// computed once per point in phase 0 and shared through shared memory.
#include "common.cuh"

namespace list {

constexpr int kPz = 64;            // points per CTA
constexpr int kRoleThreads = 128;
constexpr int kMaxRoles = 12;
constexpr int kRoleLevels = 4;     // voxel levels one role may span

struct GridRole {
  int kind;                        // 0 = 2-D vectors, 1 = 3-D vectors, 2 = scalar tail
  int first, count;                // 2-D: channel-vector range; 3-D: item range in layout order
  int nlev;                        // 3-D: levels touched (for the phase-0 tables)
  int lev[kRoleLevels];
};

struct GridGatherParams {
  const void* maps;
  const void* vols[LIST_MAX_LEVELS];
  const float* T;        // [12] of the image
  void* X;
  int64_t ldx;
  int64_t N;             // points in this launch
  int64_t grid_begin;
  int S, Cm;
  int R[LIST_MAX_LEVELS], C[LIST_MAX_LEVELS], voff[LIST_MAX_LEVELS];
  int nlev;
  int map_off, xyz_off, k_pad, tail0;
  int res;
  double bb_min, bb_max;
  int nroles;
  GridRole roles[kMaxRoles];
};

__device__ __forceinline__ int disp_order(int i) {   // W-shifted displacements (1,2) last
  return i == 0 ? 0 : (i <= 4 ? i + 2 : i - 4);
}
__device__ __forceinline__ int shift_class(int d) { return d == 1 ? 1 : (d == 2 ? 2 : 0); }

template <typename T>
__device__ __forceinline__ void load_row8(const T* __restrict__ vol, const uint32_t base[4], const float wyz[4],
                                          int xv, int C, float out[8]) {
  float v[8];
  load8(vol + base[0] + static_cast<uint32_t>(xv) * C, v);
#pragma unroll
  for (int j = 0; j < 8; ++j) out[j] = v[j] * wyz[0];
#pragma unroll
  for (int k = 1; k < 4; ++k) {
    load8(vol + base[k] + static_cast<uint32_t>(xv) * C, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) out[j] = fmaf(v[j], wyz[k], out[j]);
  }
}

struct AxEntry { int i0; float w1; };          // voxel index along W and the weight of i0+1
struct UvEntry { int x0, y0; float w00, w01, w10, w11; };

template <typename T>
__global__ void __launch_bounds__(kRoleThreads) gather_grid_kernel(const GridGatherParams p) {
  __shared__ float s_q[kPz][3];
  __shared__ int s_new[kPz];
  __shared__ AxEntry s_ax[kRoleLevels][3][kPz];
  __shared__ UvEntry s_uv[kPz];
  const int tid = threadIdx.x;
  const GridRole& role = p.roles[blockIdx.y];
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * kPz;
  const int npts = static_cast<int>(min64(kPz, p.N - n0));
  T* __restrict__ Xb = static_cast<T*>(p.X) + n0 * p.ldx;

  // ---- phase 0a: grid index -> point (numpy linspace semantics), swap/scale ----
  if (tid < kPz) {
    float q[3] = {0.f, 0.f, 0.f};
    int fresh = 1;
    if (tid < npts) {
      const int64_t g = p.grid_begin + n0 + tid;
      const int res = p.res;
      const int gz = static_cast<int>(g % res);
      const float rx = linspace_f32(static_cast<int>(g / (static_cast<int64_t>(res) * res)), res, p.bb_min, p.bb_max);
      const float ry = linspace_f32(static_cast<int>((g / res) % res), res, p.bb_min, p.bb_max);
      const float rz = linspace_f32(gz, res, p.bb_min, p.bb_max);
      q[0] = rz * 2.0f; q[1] = ry * 2.0f; q[2] = rx * 2.0f;   // reference models.py:91-92
      fresh = (tid == 0 || gz == 0) ? 1 : 0;
    }
    s_q[tid][0] = q[0]; s_q[tid][1] = q[1]; s_q[tid][2] = q[2];
    s_new[tid] = fresh;
    if (role.kind == 0) {                                       // 2-D sample position and tap weights
      float ix, iy, h[3];
      localise(q, p.T, p.S, ix, iy, h);
      UvEntry e{0, 0, 0.f, 0.f, 0.f, 0.f};                      // NaN grid -> all taps out of bounds
      if (ix == ix && iy == iy) {
        const float fx = floorf(ix), fy = floorf(iy);
        e.x0 = static_cast<int>(fx); e.y0 = static_cast<int>(fy);
        const float wx1 = ix - fx, wx0 = (fx + 1.0f) - ix;
        const float wy1 = iy - fy, wy0 = (fy + 1.0f) - iy;
        const bool okx1 = (e.x0 + 1) <= p.S - 1, oky1 = (e.y0 + 1) <= p.S - 1;
        e.w00 = wx0 * wy0;
        e.w01 = okx1 ? wx1 * wy0 : 0.f;
        e.w10 = oky1 ? wx0 * wy1 : 0.f;
        e.w11 = (okx1 && oky1) ? wx1 * wy1 : 0.f;
      }
      s_uv[tid] = e;
    }
  }
  __syncthreads();
  // ---- phase 0b: per (level, W-shift class, point) voxel index and weight ----
  if (role.kind != 0) {
    for (int i = tid; i < role.nlev * 3 * kPz; i += kRoleThreads) {
      const int s = i % kPz, cls = (i / kPz) % 3, li = i / (3 * kPz);
      const float shift = cls == 0 ? 0.f : (cls == 1 ? -kDisplacement : kDisplacement);
      const float c = cls == 0 ? s_q[s][0] : s_q[s][0] + shift;
      const Axis3 ax = axis_border(c, p.R[role.lev[li]]);
      s_ax[li][cls][s] = AxEntry{ax.i0, ax.w1};
    }
    __syncthreads();
  }

  if (role.kind == 0) {
    // ================= 2-D walker: bilinear, zeros padding, cell cache =================
    if (tid >= role.count) return;
    const int cv = role.first + tid;
    const T* __restrict__ maps = static_cast<const T*>(p.maps) + cv * 8;
    const int lim = p.S - 1;
    int cx = -1, cy = -1;
    float v00[8], v01[8], v10[8], v11[8];
    T* __restrict__ dst = Xb + p.map_off + cv * 8;
    for (int s = 0; s < npts; ++s, dst += p.ldx) {
      const UvEntry e = s_uv[s];
      if (e.x0 != cx || e.y0 != cy) {
        cx = e.x0; cy = e.y0;
        const int x1 = min(cx + 1, lim), y1 = min(cy + 1, lim);
        load8(maps + (static_cast<size_t>(cy) * p.S + cx) * p.Cm, v00);
        load8(maps + (static_cast<size_t>(cy) * p.S + x1) * p.Cm, v01);
        load8(maps + (static_cast<size_t>(y1) * p.S + cx) * p.Cm, v10);
        load8(maps + (static_cast<size_t>(y1) * p.S + x1) * p.Cm, v11);
      }
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        acc[j] = fmaf(v11[j], e.w11, fmaf(v10[j], e.w10, fmaf(v01[j], e.w01, v00[j] * e.w00)));
      store8(dst, acc);
    }
  } else if (role.kind == 1) {
    // ================= 3-D vector walker =================
    if (tid >= role.count) return;
    int item = role.first + tid;
    int l = -1, li = -1, d = 0, cv = 0;
    for (int ll = p.nlev - 1; ll >= 0; --ll) {        // same level order as the row layout
      if (p.C[ll] & 7) continue;
      const int cnt = LIST_NUM_DISP * (p.C[ll] >> 3);
      if (item < cnt) {
        l = ll;
        const int ncv = p.C[ll] >> 3;
        d = disp_order(item / ncv);
        cv = item % ncv;
        break;
      }
      item -= cnt;
    }
    if (l < 0) return;
    for (int k = 0; k < role.nlev; ++k)
      if (role.lev[k] == l) li = k;
    const int R = p.R[l], C = p.C[l];
    const T* __restrict__ vol = static_cast<const T*>(p.vols[l]);
    const AxEntry* __restrict__ axs = s_ax[li][shift_class(d)];
    uint32_t base[4] = {0, 0, 0, 0};
    float wyz[4] = {0.f, 0.f, 0.f, 0.f};
    float G0[8], G1[8], D[8];
    int cx0 = -1;
    T* __restrict__ dst = Xb + p.voff[l] + d * C + cv * 8;
    for (int s = 0; s < npts; ++s, dst += p.ldx) {
      const AxEntry e = axs[s];
      bool reload = false;
      if (s_new[s]) {                                  // new (x, y) run: (H, D) corners and weights
        const float q[3] = {s_q[s][0], s_q[s][1], s_q[s][2]};
        float pd[3];
        displaced(q, d, pd);
        const Axis3 ay = axis_border(pd[1], R), az = axis_border(pd[2], R);
        const int zi[2] = {az.i0, az.i1}, yi[2] = {ay.i0, ay.i1};
        const float wz[2] = {az.w0, az.w1}, wy[2] = {ay.w0, ay.w1};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int tz = k >> 1, ty = k & 1;
          base[k] = (static_cast<uint32_t>(zi[tz]) * R + yi[ty]) * R * C + cv * 8;
          wyz[k] = wy[ty] * wz[tz];
        }
        reload = true;
      }
      if (reload || e.i0 != cx0) {
        const int i1 = min(e.i0 + 1, R - 1);
        if (!reload && e.i0 == cx0 + 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) G0[j] = G1[j];
        } else {
          load_row8(vol, base, wyz, e.i0, C, G0);
        }
        if (i1 != e.i0) load_row8(vol, base, wyz, i1, C, G1);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) G1[j] = G0[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) D[j] = G1[j] - G0[j];
        cx0 = e.i0;
      }
      float out[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) out[j] = fmaf(D[j], e.w1, G0[j]);
      store8(dst, out);
    }
  } else {
    // ================= tail: scalar levels, q, zero pad (one warp) =================
    if (tid >= 32) return;
    const int lane = tid;
    const int ntail = p.k_pad - p.tail0;               // <= 64 (checked by the launcher)
    const int nscal = p.xyz_off - p.tail0;             // scalar-level columns, <= 29
    int l = -1, li = 0, d = 0, c = 0;
    if (lane < nscal) {
      const int colabs = p.tail0 + lane;
      for (int ll = 0; ll < p.nlev; ++ll) {
        if (!(p.C[ll] & 7)) continue;
        const int rel = colabs - p.voff[ll];
        if (rel >= 0 && rel < LIST_NUM_DISP * p.C[ll]) { l = ll; d = rel / p.C[ll]; c = rel % p.C[ll]; }
      }
      for (int k = 0; k < role.nlev; ++k)
        if (role.lev[k] == l) li = k;
    }
    const int R = l >= 0 ? p.R[l] : 1, C = l >= 0 ? p.C[l] : 1;
    const T* __restrict__ vol = l >= 0 ? static_cast<const T*>(p.vols[l]) : nullptr;
    const AxEntry* __restrict__ axs = s_ax[li][shift_class(d)];
    uint32_t base[4] = {0, 0, 0, 0};
    float wyz[4] = {0.f, 0.f, 0.f, 0.f};
    float g0 = 0.f, g1 = 0.f;
    int cx0 = -1;
    T zero;
    from_f32(zero, 0.f);
    for (int s = 0; s < npts; ++s) {
      float val = 0.f;
      if (l >= 0) {
        const AxEntry e = axs[s];
        bool reload = false;
        if (s_new[s]) {
          const float q[3] = {s_q[s][0], s_q[s][1], s_q[s][2]};
          float pd[3];
          displaced(q, d, pd);
          const Axis3 ay = axis_border(pd[1], R), az = axis_border(pd[2], R);
          const int zi[2] = {az.i0, az.i1}, yi[2] = {ay.i0, ay.i1};
          const float wz[2] = {az.w0, az.w1}, wy[2] = {ay.w0, ay.w1};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int tz = k >> 1, ty = k & 1;
            base[k] = (static_cast<uint32_t>(zi[tz]) * R + yi[ty]) * R * C + c;
            wyz[k] = wy[ty] * wz[tz];
          }
          reload = true;
        }
        if (reload || e.i0 != cx0) {
          auto row = [&](int xv) {
            float r = to_f32(vol[base[0] + static_cast<uint32_t>(xv) * C]) * wyz[0];
#pragma unroll
            for (int k = 1; k < 4; ++k) r = fmaf(to_f32(vol[base[k] + static_cast<uint32_t>(xv) * C]), wyz[k], r);
            return r;
          };
          const int i1 = min(e.i0 + 1, R - 1);
          g0 = (!reload && e.i0 == cx0 + 1) ? g1 : row(e.i0);
          g1 = (i1 != e.i0) ? row(i1) : g0;
          cx0 = e.i0;
        }
        val = fmaf(g1 - g0, e.w1, g0);
      } else if (lane >= nscal && lane < nscal + 3) {
        val = s_q[s][lane - nscal];
      }
      T* row_out = Xb + static_cast<int64_t>(s) * p.ldx + p.tail0;
      T o;
      from_f32(o, val);
      if (lane < ntail) row_out[lane] = o;
      if (lane + 32 < ntail) row_out[lane + 32] = zero;   // columns beyond nscal+3 are padding (nscal+3 <= 32)
    }
  }
}

// Returns LIST_ENOSYS when the configuration does not fit the walker mapping (the caller then
// uses the generic kernel, which is correct for every configuration).
int gather_grid_walk(const ListCtx* ctx, int image, int res, double bb_min, double bb_max, int64_t begin, int64_t count,
                     void* X, int64_t ldx, cudaStream_t st) {
  ListLayout lay;
  const int rc = list_feature_layout(ctx->map_channels, ctx->n_levels, ctx->vol_ch, &lay, nullptr);
  if (rc) return rc;
  GridGatherParams p{};
  const size_t es = ctx->dtype == LIST_BF16 ? 2 : 4;
  p.maps = static_cast<const char*>(ctx->maps) + static_cast<size_t>(image) * ctx->map_size * ctx->map_size * ctx->map_channels * es;
  p.T = ctx->trans_mat + image * 12;
  p.X = X;
  p.ldx = ldx;
  p.N = count;
  p.grid_begin = begin;
  p.S = ctx->map_size;
  p.Cm = ctx->map_channels;
  p.nlev = ctx->n_levels;
  int tail0 = lay.xyz_off;
  for (int l = 0; l < ctx->n_levels; ++l) {
    const size_t vox = static_cast<size_t>(ctx->vol_res[l]) * ctx->vol_res[l] * ctx->vol_res[l] * ctx->vol_ch[l];
    if (vox >= (1ull << 32)) return LIST_ENOSYS;      // 32-bit element offsets inside a volume
    p.vols[l] = static_cast<const char*>(ctx->vols[l]) + static_cast<size_t>(image) * vox * es;
    p.R[l] = ctx->vol_res[l];
    p.C[l] = ctx->vol_ch[l];
    p.voff[l] = lay.vol_off[l];
    if (ctx->vol_ch[l] % 8 != 0) tail0 = lay.vol_off[l] < tail0 ? lay.vol_off[l] : tail0;
  }
  p.map_off = lay.map_off;
  p.xyz_off = lay.xyz_off;
  p.k_pad = lay.k_pad;
  p.tail0 = tail0;
  p.res = res;
  p.bb_min = bb_min;
  p.bb_max = bb_max;
  if (lay.xyz_off + 3 - tail0 > 32 || lay.k_pad - tail0 > 64) return LIST_ENOSYS;

  // ---- roles ----
  int nr = 0;
  for (int first = 0; first < ctx->map_channels / 8; first += kRoleThreads) {     // 2-D vectors
    if (nr >= kMaxRoles) return LIST_ENOSYS;
    GridRole& r = p.roles[nr++];
    r.kind = 0;
    r.first = first;
    r.count = (ctx->map_channels / 8 - first) < kRoleThreads ? (ctx->map_channels / 8 - first) : kRoleThreads;
  }
  {                                                   // 3-D vector levels, layout order, packed greedily
    int item0 = 0;
    GridRole cur{};
    cur.kind = 1;
    cur.first = 0;
    for (int l = ctx->n_levels - 1; l >= 0; --l) {
      if (ctx->vol_ch[l] % 8) continue;
      int cnt = LIST_NUM_DISP * (ctx->vol_ch[l] / 8);
      int done = 0;
      while (done < cnt) {
        const int room = kRoleThreads - cur.count;
        const bool level_fits = (cnt - done) <= room && cur.nlev < kRoleLevels;
        if (room == 0 || (!level_fits && cur.count > 0)) {            // close the current role
          if (nr >= kMaxRoles) return LIST_ENOSYS;
          p.roles[nr++] = cur;
          cur = GridRole{};
          cur.kind = 1;
          cur.first = item0;
          continue;
        }
        const int take = (cnt - done) < room ? (cnt - done) : room;
        if (cur.nlev == 0 || cur.lev[cur.nlev - 1] != l) cur.lev[cur.nlev++] = l;
        cur.count += take;
        done += take;
        item0 += take;
      }
    }
    if (cur.count > 0) {
      if (nr >= kMaxRoles) return LIST_ENOSYS;
      p.roles[nr++] = cur;
    }
  }
  {                                                   // scalar tail
    if (nr >= kMaxRoles) return LIST_ENOSYS;
    GridRole& r = p.roles[nr++];
    r.kind = 2;
    r.count = 32;
    for (int l = 0; l < ctx->n_levels; ++l)
      if (ctx->vol_ch[l] % 8) {
        if (r.nlev >= kRoleLevels) return LIST_ENOSYS;
        r.lev[r.nlev++] = l;
      }
  }
  p.nroles = nr;
  if (count == 0) return LIST_OK;
  dim3 grid(static_cast<unsigned>((count + kPz - 1) / kPz), nr);
  if (ctx->dtype == LIST_F32) gather_grid_kernel<float><<<grid, kRoleThreads, 0, st>>>(p);
  else gather_grid_kernel<__nv_bfloat16><<<grid, kRoleThreads, 0, st>>>(p);
  LIST_LAUNCH_CHECK("gather_grid_kernel");
  return LIST_OK;
}

}  // namespace list
