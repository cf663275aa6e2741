// Hoisted fc_0 for dense grids in the bf16 tensor-core mode.
//
// fc_0 is linear and so are the bilinear / trilinear samplers in front of it (reference
// network/modules.py:48-52, 264-265 feeding :276), hence for the perceptual maps and the coarse voxel
// levels
//        W0[:, cols(f)] · sample(f, p)  ==  sample(W0[:, cols(f)] · f, p)
// i.e. the 512-wide fc_0 pre-activation contribution of a feature tensor f can be computed by
// projecting f through its block of W0 ONCE per image (a small tensor-core GEMM over pixels / voxels,
// mlp_tc_project) and sampling the projected tensor per query.  On a dense grid this removes
//        1024 (maps) + 7*128 (level 16^3) + 7*128 (level 8^3) = 2816 of the 3610 K columns
// from the per-query GEMM and replaces them by ONE 512-wide "addend" block whose weight block is the
// identity: the feature row shrinks from 3648 to 512 + 832 = 1344 columns,
//        X_h = [ addend(512) | level 32^3 (448) | level 64^3 (224) | level 128^3 (112) | occupancy (7) | q (3) | 0-pad ]
//        W0h = [ I_512       | the same columns of W0 ............................................................ ]
// so that  W0h · X_h == W0 · X  exactly in real arithmetic.  The fp32 parity path is untouched.
//
// Kernels here:
//   build_w0h_kernel     W0h from W0 (identity block + verbatim tail columns).
//   hoist_addend_kernel  per 64-point tile of the grid, 64 threads, thread = 8 of the 512 addend channels:
//                        walks the z-run with the separable scheme of gather_grid.cu -- bilinear taps of the
//                        projected map (cell cache) + for each hoisted level the three W-shift classes
//                        {d=0,3,4,5,6}, {d=1}, {d=2} of the displacement table (modules.py:205-212): the five
//                        class-0 displacements share the voxel index and weight along the walk, so their
//                        (H,D)-interpolated projected columns are summed ONCE per voxel cell and a step costs one
//                        FMA per channel and class.  State per class: G1 (column at i0+1) and D = G1 - G0, a
//                        function of the cell only (bit-exact under any chunking / sharding of the grid).
//   hoist_rest_kernel    the remaining feature columns (levels 32^3, 64^3, 128^3, occupancy, q, pad) exactly as
//                        gather_grid.cu computes them, with break masks so the per-step loop is branch-free.
#include <cstdlib>

#include "hoist.cuh"

namespace list {

int mlp_tc_project(const ListWeights* w, int col0, int col_stride, int groups, int k, const void* X, int64_t ldx,
                   int64_t rows, void* out, cudaStream_t st);

namespace hoist {

constexpr int kPz = 64;             // points per CTA (one 64-bit break mask per walker class)
constexpr int kMaxRuns = 4;         // z-runs one tile may touch (launcher: res >= 32)
constexpr int kN0 = 512;            // fc_0 width == addend channels
constexpr int kAddThreads = kN0 / 8;
constexpr int kRestVecThreads = 128;
constexpr int kRestThreads = kRestVecThreads + 32;
constexpr int kRestLevels = 4;

struct Corner { uint32_t base; float w; };

struct AddParams {
  const __nv_bfloat16* pmap;        // image's [S][S][512]
  const __nv_bfloat16* pvol[kMaxH]; // image's slab of displacement 0: [R][R][R][512]
  uint32_t dstride[kMaxH];          // elements between displacement slabs
  const float* T;
  __nv_bfloat16* X;
  int64_t ldx, N, grid_begin;
  int S, res, nh;
  int R[kMaxH];
  double bb_min, bb_max;
};

struct RestParams {
  const __nv_bfloat16* vols[kRestLevels];
  int R[kRestLevels], C[kRestLevels], xoff[kRestLevels];   // xoff: first column of the level in the hoisted row
  int nlev;                         // levels handled here (vector levels first, then scalar ones)
  int nvec_items;                   // 16-byte items of the vector levels (<= 128)
  int tail0, xyz_off, k_h;          // tail region [tail0, k_h): scalar levels, q, zero pad
  __nv_bfloat16* X;
  int64_t ldx, N, grid_begin;
  int res;
  double bb_min, bb_max;
};

__device__ __forceinline__ int shift_class(int d) { return d == 1 ? 1 : (d == 2 ? 2 : 0); }

// grid index -> swapped/scaled query (reference utils.py:84-95, models.py:91-92)
__device__ __forceinline__ void grid_query(int64_t g, int res, double lo, double hi, float q[3], int& gz) {
  gz = static_cast<int>(g % res);
  const float rx = linspace_f32(static_cast<int>(g / (static_cast<int64_t>(res) * res)), res, lo, hi);
  const float ry = linspace_f32(static_cast<int>((g / res) % res), res, lo, hi);
  const float rz = linspace_f32(gz, res, lo, hi);
  q[0] = rz * 2.0f; q[1] = ry * 2.0f; q[2] = rx * 2.0f;
}

__device__ __forceinline__ int run_len(uint64_t mask, int s, int npts) {
  const uint64_t rest = (s + 1 < 64) ? (mask >> (s + 1)) : 0ull;
  const int n = rest ? __ffsll(static_cast<long long>(rest)) : 64;
  return min(n, npts - s);
}

// ------------------------------------------------------------------ W0h
__global__ void build_w0h_kernel(const __nv_bfloat16* __restrict__ w0, int k_pad, int hoist_cols, int k_h,
                                 __nv_bfloat16* __restrict__ w0h) {
  const int n = blockIdx.x;
  for (int j = threadIdx.x; j < k_h; j += blockDim.x) {
    __nv_bfloat16 v;
    if (j < kN0) v = __float2bfloat16_rn(j == n ? 1.0f : 0.0f);
    else v = w0[static_cast<size_t>(n) * k_pad + hoist_cols + (j - kN0)];
    w0h[static_cast<size_t>(n) * k_h + j] = v;
  }
}

// ------------------------------------------------------------------ addend
// V consecutive bf16 channels (V = 8: 16-byte, V = 4: 8-byte accesses)
template <int V>
__device__ __forceinline__ void loadv(const __nv_bfloat16* __restrict__ p, float v[V]) {
  if constexpr (V == 8) {
    load8(p, v);
  } else {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
    v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
  }
}
template <int V>
__device__ __forceinline__ void storev(__nv_bfloat16* __restrict__ p, const float v[V]) {
  if constexpr (V == 8) {
    store8(p, v);
  } else {
    uint2 u;
    u.x = pack_bf16x2(v[0], v[1]);
    u.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = u;
  }
}

template <int CLS, int V>
__device__ __forceinline__ void load_column(const __nv_bfloat16* __restrict__ pv, uint32_t dstride, const Corner* __restrict__ corners,
                                            int xv, int cv, float out[V]) {
  // corners: [7][4] of the current run and level; sums the class's displacements and 4 (H,D) corners
  constexpr int nd = CLS == 0 ? 5 : 1;
  const int dlist[5] = {CLS == 0 ? 0 : CLS, 3, 4, 5, 6};
#pragma unroll
  for (int j = 0; j < V; ++j) out[j] = 0.f;
#pragma unroll
  for (int di = 0; di < nd; ++di) {
    const int d = dlist[di];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const Corner c = corners[d * 4 + k];
      float v[V];
      loadv<V>(pv + static_cast<size_t>(d) * dstride + c.base + static_cast<uint32_t>(xv) * kN0 + cv * V, v);
#pragma unroll
      for (int j = 0; j < V; ++j) out[j] = fmaf(v[j], c.w, out[j]);
    }
  }
}

template <int V>
__global__ void __launch_bounds__(kN0 / V) hoist_addend_kernel(const AddParams p) {
  __shared__ float s_q[kPz][3];
  __shared__ __align__(16) float s_w[kPz][12];          // per step: w0 of the 6 classes, w00 w01 w10 w11, 2 pad
  __shared__ int s_i0[kMaxH * 3][kPz];
  __shared__ int s_xy[kPz];                             // y0 << 16 | x0
  __shared__ uint64_t s_mask[kMaxH * 3 + 1];
  __shared__ Corner s_corner[kMaxRuns][kMaxH][LIST_NUM_DISP * 4];
  const int tid = threadIdx.x;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * kPz;
  const int npts = static_cast<int>(min64(kPz, p.N - n0));
  const int64_t g0 = p.grid_begin + n0;
  const int gz0 = static_cast<int>(g0 % p.res);
  const int nruns = (gz0 + npts - 1) / p.res + 1;

  // ---- phase 0a: query, 2-D cell and weights ----
  if (tid < kPz) {
    float q[3] = {0.f, 0.f, 0.f};
    int gz = 0;
    if (tid < npts) grid_query(g0 + tid, p.res, p.bb_min, p.bb_max, q, gz);
    s_q[tid][0] = q[0]; s_q[tid][1] = q[1]; s_q[tid][2] = q[2];
    float ix, iy, h[3];
    localise(q, p.T, p.S, ix, iy, h);
    int x0 = 0, y0 = 0;
    float w00 = 0.f, w01 = 0.f, w10 = 0.f, w11 = 0.f;                    // NaN grid -> all taps out of bounds
    if (ix == ix && iy == iy) {
      const float fx = floorf(ix), fy = floorf(iy);
      x0 = static_cast<int>(fx); y0 = static_cast<int>(fy);
      const float wx1 = ix - fx, wx0 = (fx + 1.0f) - ix;
      const float wy1 = iy - fy, wy0 = (fy + 1.0f) - iy;
      const bool okx1 = (x0 + 1) <= p.S - 1, oky1 = (y0 + 1) <= p.S - 1;
      w00 = wx0 * wy0;
      w01 = okx1 ? wx1 * wy0 : 0.f;
      w10 = oky1 ? wx0 * wy1 : 0.f;
      w11 = (okx1 && oky1) ? wx1 * wy1 : 0.f;
    }
    s_xy[tid] = (y0 << 16) | x0;
    s_w[tid][6] = w00; s_w[tid][7] = w01; s_w[tid][8] = w10; s_w[tid][9] = w11;
    s_w[tid][10] = 0.f; s_w[tid][11] = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxH * 3; ++c) {                               // voxel index / weight along the walk
      const int hh = c / 3, cls = c % 3;
      int i0 = 0;
      float w0 = 0.f;
      if (hh < p.nh) {
        const float shift = cls == 0 ? 0.f : (cls == 1 ? -kDisplacement : kDisplacement);
        const Axis3 ax = axis_border(cls == 0 ? q[0] : q[0] + shift, p.R[hh]);
        i0 = ax.i0; w0 = ax.w0;
      }
      s_i0[c][tid] = i0;
      s_w[tid][c] = w0;
    }
  }
  __syncthreads();
  // ---- phase 0b: break masks (a step whose cell differs from the previous step's, or starts a z-run) ----
  if (tid < kPz) {
    const int s = tid;
    const bool fresh = s == 0 || ((gz0 + s) % p.res) == 0;
#pragma unroll
    for (int c = 0; c < kMaxH * 3; ++c) {
      const bool brk = s < npts && (fresh || s_i0[c][s] != s_i0[c][s - 1]);
      const uint32_t b = __ballot_sync(0xffffffffu, brk);
      if ((tid & 31) == 0) reinterpret_cast<uint32_t*>(&s_mask[c])[tid >> 5] = b;
    }
    const bool brk2 = s < npts && (s == 0 || s_xy[s] != s_xy[s - 1]);
    const uint32_t b2 = __ballot_sync(0xffffffffu, brk2);
    if ((tid & 31) == 0) reinterpret_cast<uint32_t*>(&s_mask[kMaxH * 3])[tid >> 5] = b2;
    // (H,D) corners and weights per (run, level, displacement)
    if (tid < nruns * p.nh * LIST_NUM_DISP) {
      const int d = tid % LIST_NUM_DISP, hh = (tid / LIST_NUM_DISP) % p.nh, r = tid / (LIST_NUM_DISP * p.nh);
      const int s0 = max(0, r * p.res - gz0);                            // first step of run r inside the tile
      const float q[3] = {s_q[s0][0], s_q[s0][1], s_q[s0][2]};
      float pd[3];
      displaced(q, d, pd);
      const int R = hh == 0 ? p.R[0] : p.R[1];
      const Axis3 ay = axis_border(pd[1], R), az = axis_border(pd[2], R);
      const int zi[2] = {az.i0, az.i1}, yi[2] = {ay.i0, ay.i1};
      const float wz[2] = {az.w0, az.w1}, wy[2] = {ay.w0, ay.w1};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int tz = k >> 1, ty = k & 1;
        Corner c;
        c.base = (static_cast<uint32_t>(zi[tz]) * R + yi[ty]) * R * kN0;
        c.w = wy[ty] * wz[tz];
        s_corner[r][hh][d * 4 + k] = c;
      }
    }
  }
  __syncthreads();

  const int cv = tid;
  uint64_t m[kMaxH * 3 + 1], many = 0;
#pragma unroll
  for (int c = 0; c <= kMaxH * 3; ++c) {
    m[c] = (c < kMaxH * 3 && c / 3 >= p.nh) ? 0ull : s_mask[c];
    many |= m[c];
  }
  float G1[kMaxH * 3][V], D[kMaxH * 3][V], gsum[V];
#pragma unroll
  for (int j = 0; j < V; ++j) gsum[j] = 0.f;
  float v00[V], v01[V], v10[V], v11[V];
  int cur[kMaxH * 3];
#pragma unroll
  for (int c = 0; c < kMaxH * 3; ++c) {
    cur[c] = -2;
#pragma unroll
    for (int j = 0; j < V; ++j) { G1[c][j] = 0.f; D[c][j] = 0.f; }
  }
#pragma unroll
  for (int j = 0; j < V; ++j) { v00[j] = v01[j] = v10[j] = v11[j] = 0.f; }

  __nv_bfloat16* __restrict__ dst = p.X + n0 * p.ldx + cv * V;
  const int lim = p.S - 1;
  int s = 0;
  while (s < npts) {
    const bool fresh = s == 0 || ((gz0 + s) % p.res) == 0;
    const int run = (gz0 + s) / p.res;
    bool any3 = false;
#pragma unroll
    for (int c = 0; c < kMaxH * 3; ++c) {
      if ((m[c] >> s) & 1ull) {
        any3 = true;
        const int hh = c / 3;
        const int R = p.R[hh];
        const int i0 = s_i0[c][s];
        const int i1 = min(i0 + 1, R - 1);
        const Corner* corners = s_corner[run][hh];
        float g0v[V], g1v[V];
        if (!fresh && i0 == cur[c] + 1) {
#pragma unroll
          for (int j = 0; j < V; ++j) g0v[j] = G1[c][j];
        } else {
          if (c % 3 == 0) load_column<0, V>(p.pvol[hh], p.dstride[hh], corners, i0, cv, g0v);
          else if (c % 3 == 1) load_column<1, V>(p.pvol[hh], p.dstride[hh], corners, i0, cv, g0v);
          else load_column<2, V>(p.pvol[hh], p.dstride[hh], corners, i0, cv, g0v);
        }
        if (i1 != i0) {
          if (c % 3 == 0) load_column<0, V>(p.pvol[hh], p.dstride[hh], corners, i1, cv, g1v);
          else if (c % 3 == 1) load_column<1, V>(p.pvol[hh], p.dstride[hh], corners, i1, cv, g1v);
          else load_column<2, V>(p.pvol[hh], p.dstride[hh], corners, i1, cv, g1v);
        } else {
#pragma unroll
          for (int j = 0; j < V; ++j) g1v[j] = g0v[j];
        }
#pragma unroll
        for (int j = 0; j < V; ++j) { D[c][j] = g1v[j] - g0v[j]; G1[c][j] = g1v[j]; }
        cur[c] = i0;
      }
    }
    if (any3) {
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float a = G1[0][j];
#pragma unroll
        for (int c = 1; c < kMaxH * 3; ++c) a += G1[c][j];
        gsum[j] = a;
      }
    }
    if ((m[kMaxH * 3] >> s) & 1ull) {
      const int xy = s_xy[s];
      const int cx = xy & 0xffff, cy = xy >> 16;
      const int x1 = min(cx + 1, lim), y1 = min(cy + 1, lim);
      const __nv_bfloat16* __restrict__ pm = p.pmap + cv * V;
      loadv<V>(pm + (static_cast<size_t>(cy) * p.S + cx) * kN0, v00);
      loadv<V>(pm + (static_cast<size_t>(cy) * p.S + x1) * kN0, v01);
      loadv<V>(pm + (static_cast<size_t>(y1) * p.S + cx) * kN0, v10);
      loadv<V>(pm + (static_cast<size_t>(y1) * p.S + x1) * kN0, v11);
    }
    const int n = run_len(many, s, npts);
#pragma unroll 2
    for (int k = 0; k < n; ++k, ++s, dst += p.ldx) {
      const float4 wa = *reinterpret_cast<const float4*>(&s_w[s][0]);
      const float4 wb = *reinterpret_cast<const float4*>(&s_w[s][4]);
      const float2 wc = *reinterpret_cast<const float2*>(&s_w[s][8]);
      const float w3[kMaxH * 3] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y};
      float acc[V];
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float a = fmaf(v00[j], wb.z, gsum[j]);
        a = fmaf(v01[j], wb.w, a);
        a = fmaf(v10[j], wc.x, a);
        a = fmaf(v11[j], wc.y, a);
#pragma unroll
        for (int c = 0; c < kMaxH * 3; ++c) a = fmaf(-w3[c], D[c][j], a);
        acc[j] = a;
      }
      storev<V>(dst, acc);
    }
  }
}

// ------------------------------------------------------------------ the remaining columns
__device__ __forceinline__ void load_row8(const __nv_bfloat16* __restrict__ vol, const uint32_t base[4], const float wyz[4],
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

__global__ void __launch_bounds__(kRestThreads) hoist_rest_kernel(const RestParams p) {
  __shared__ float s_q[kPz][3];
  __shared__ int s_i0[kRestLevels][3][kPz];
  __shared__ float s_w1[kRestLevels][3][kPz];
  __shared__ uint64_t s_mask[kRestLevels][3];
  const int tid = threadIdx.x;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * kPz;
  const int npts = static_cast<int>(min64(kPz, p.N - n0));
  const int64_t g0 = p.grid_begin + n0;
  const int gz0 = static_cast<int>(g0 % p.res);
  __nv_bfloat16* __restrict__ Xb = p.X + n0 * p.ldx;

  if (tid < kPz) {
    float q[3] = {0.f, 0.f, 0.f};
    int gz = 0;
    if (tid < npts) grid_query(g0 + tid, p.res, p.bb_min, p.bb_max, q, gz);
    s_q[tid][0] = q[0]; s_q[tid][1] = q[1]; s_q[tid][2] = q[2];
  }
  __syncthreads();
  for (int i = tid; i < p.nlev * 3 * kPz; i += kRestThreads) {
    const int s = i % kPz, cls = (i / kPz) % 3, li = i / (3 * kPz);
    const float shift = cls == 0 ? 0.f : (cls == 1 ? -kDisplacement : kDisplacement);
    const Axis3 ax = axis_border(cls == 0 ? s_q[s][0] : s_q[s][0] + shift, p.R[li]);
    s_i0[li][cls][s] = ax.i0;
    s_w1[li][cls][s] = ax.w1;
  }
  __syncthreads();
  for (int i = tid; i < p.nlev * 3 * kPz; i += kRestThreads) {         // 160 = 5 warps: each warp covers 32 steps of one class
    const int s = i % kPz, cls = (i / kPz) % 3, li = i / (3 * kPz);
    const bool fresh = s == 0 || ((gz0 + s) % p.res) == 0;
    const bool brk = s < npts && (fresh || s_i0[li][cls][s] != s_i0[li][cls][s - 1]);
    const uint32_t b = __ballot_sync(0xffffffffu, brk);
    if ((tid & 31) == 0) reinterpret_cast<uint32_t*>(&s_mask[li][cls])[s >> 5] = b;
  }
  __syncthreads();

  if (tid < kRestVecThreads) {
    // ================= 3-D vector walker =================
    if (tid >= p.nvec_items) return;
    int item = tid, li = -1, d = 0, cv = 0;
    for (int ll = 0; ll < p.nlev; ++ll) {
      if (p.C[ll] & 7) continue;
      const int ncv = p.C[ll] >> 3;
      const int cnt = LIST_NUM_DISP * ncv;
      if (item < cnt) {
        li = ll;
        const int di = item / ncv;
        d = di == 0 ? 0 : (di <= 4 ? di + 2 : di - 4);                  // W-shifted displacements (1,2) last
        cv = item % ncv;
        break;
      }
      item -= cnt;
    }
    if (li < 0) return;
    const int R = p.R[li], C = p.C[li];
    const int cls = shift_class(d);
    const __nv_bfloat16* __restrict__ vol = p.vols[li];
    const int* __restrict__ i0s = s_i0[li][cls];
    const float* __restrict__ w1s = s_w1[li][cls];
    const uint64_t m = s_mask[li][cls];
    uint32_t base[4] = {0, 0, 0, 0};
    float wyz[4] = {0.f, 0.f, 0.f, 0.f};
    float G0[8], G1[8], Dv[8];
    int cx0 = -2;
    __nv_bfloat16* __restrict__ dst = Xb + p.xoff[li] + d * C + cv * 8;
    int s = 0;
    while (s < npts) {
      const bool fresh = s == 0 || ((gz0 + s) % p.res) == 0;
      if (fresh) {                                       // new (x, y) run: (H, D) corners and weights
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
      }
      const int i0 = i0s[s];
      const int i1 = min(i0 + 1, R - 1);
      if (!fresh && i0 == cx0 + 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) G0[j] = G1[j];
      } else {
        load_row8(vol, base, wyz, i0, C, G0);
      }
      if (i1 != i0) load_row8(vol, base, wyz, i1, C, G1);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) G1[j] = G0[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) Dv[j] = G1[j] - G0[j];
      cx0 = i0;
      const int n = run_len(m, s, npts);
      for (int k = 0; k < n; ++k, ++s, dst += p.ldx) {
        const float w1 = w1s[s];
        float out[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) out[j] = fmaf(Dv[j], w1, G0[j]);
        store8(dst, out);
      }
    }
  } else {
    // ================= tail: scalar levels, q, zero pad (one warp) =================
    const int lane = tid - kRestVecThreads;
    const int ntail = p.k_h - p.tail0;                 // <= 64 (checked by the launcher)
    const int nscal = p.xyz_off - p.tail0;             // scalar-level columns, <= 29
    int li = -1, d = 0, c = 0;
    if (lane < nscal) {
      const int colabs = p.tail0 + lane;
      for (int ll = 0; ll < p.nlev; ++ll) {
        if (!(p.C[ll] & 7)) continue;
        const int rel = colabs - p.xoff[ll];
        if (rel >= 0 && rel < LIST_NUM_DISP * p.C[ll]) { li = ll; d = rel / p.C[ll]; c = rel % p.C[ll]; }
      }
    }
    const int R = li >= 0 ? p.R[li] : 1, C = li >= 0 ? p.C[li] : 1;
    const __nv_bfloat16* __restrict__ vol = li >= 0 ? p.vols[li] : nullptr;
    const int cls = shift_class(d);
    const int lsel = li >= 0 ? li : 0;
    uint32_t base[4] = {0, 0, 0, 0};
    float wyz[4] = {0.f, 0.f, 0.f, 0.f};
    float g0 = 0.f, g1 = 0.f;
    int cx0 = -2;
    const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
    for (int s = 0; s < npts; ++s) {
      float val = 0.f;
      if (li >= 0) {
        const int i0 = s_i0[lsel][cls][s];
        const bool fresh = s == 0 || ((gz0 + s) % p.res) == 0;
        if (fresh) {
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
        }
        if (fresh || i0 != cx0) {
          auto row = [&](int xv) {
            float r = __bfloat162float(vol[base[0] + static_cast<uint32_t>(xv) * C]) * wyz[0];
#pragma unroll
            for (int k = 1; k < 4; ++k) r = fmaf(__bfloat162float(vol[base[k] + static_cast<uint32_t>(xv) * C]), wyz[k], r);
            return r;
          };
          const int i1 = min(i0 + 1, R - 1);
          g0 = (!fresh && i0 == cx0 + 1) ? g1 : row(i0);
          g1 = (i1 != i0) ? row(i1) : g0;
          cx0 = i0;
        }
        val = fmaf(g1 - g0, s_w1[lsel][cls][s], g0);
      } else if (lane >= nscal && lane < nscal + 3) {
        val = s_q[s][lane - nscal];
      }
      __nv_bfloat16* row_out = Xb + static_cast<int64_t>(s) * p.ldx + p.tail0;
      if (lane < ntail) row_out[lane] = __float2bfloat16_rn(val);
      if (lane + 32 < ntail) row_out[lane + 32] = zero;   // columns beyond nscal+3 are padding (nscal+3 <= 32)
    }
  }
}

// ------------------------------------------------------------------ host side
// Which part of the row is hoisted: the maps and the leading vector levels of the layout (coarsest first)
// whose projected volumes stay small (R <= 16) and whose channel count feeds the tensor-core projection (C % 64).
int make_plan(const ListCtx* ctx, const ListWeights* w, Plan* pl) {
  if (ctx->dtype != LIST_BF16 || w->dtype != LIST_BF16 || w->n0 != kN0) return LIST_ENOSYS;
  if (ctx->map_channels % 64 != 0) return LIST_ENOSYS;
  ListLayout lay;
  const int rc = list_feature_layout(ctx->map_channels, ctx->n_levels, ctx->vol_ch, &lay, nullptr);
  if (rc) return rc;
  if (lay.map_off != 0 || w->k_pad != lay.k_pad) return LIST_ENOSYS;
  pl->nh = 0;
  int cols = ctx->map_channels;
  for (int l = ctx->n_levels - 1; l >= 0 && pl->nh < kMaxH; --l) {       // layout order of the vector levels
    if (ctx->vol_ch[l] % 8) continue;
    if (ctx->vol_res[l] > 16 || ctx->vol_ch[l] % 64 != 0 || lay.vol_off[l] != cols) break;
    pl->lev[pl->nh++] = l;
    cols += LIST_NUM_DISP * ctx->vol_ch[l];
  }
  if (cols % 64 != 0) return LIST_ENOSYS;
  pl->hoist_cols = cols;
  pl->k_h = kN0 + (lay.k_pad - cols);
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  size_t off = 0;
  pl->off_w0h = off; off += up(static_cast<size_t>(kN0) * pl->k_h * 2);
  pl->off_pmap = off; off += up(static_cast<size_t>(ctx->B) * ctx->map_size * ctx->map_size * kN0 * 2);
  for (int h = 0; h < pl->nh; ++h) {
    const size_t R = ctx->vol_res[pl->lev[h]];
    if (static_cast<size_t>(LIST_NUM_DISP) * ctx->B * R * R * R * kN0 >= (1ull << 32)) return LIST_ENOSYS;
    pl->off_pvol[h] = off;
    off += up(static_cast<size_t>(LIST_NUM_DISP) * ctx->B * R * R * R * kN0 * 2);
  }
  pl->total = off;
  return LIST_OK;
}

// Projects the maps and the hoisted levels of every image through their W0 blocks and builds W0h.
int prepare(const ListCtx* ctx, const ListWeights* w, const Plan& pl, void* buf, cudaStream_t st) {
  char* base = static_cast<char*>(buf);
  ListLayout lay;
  int rc = list_feature_layout(ctx->map_channels, ctx->n_levels, ctx->vol_ch, &lay, nullptr);
  if (rc) return rc;
  build_w0h_kernel<<<kN0, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(w->w0), w->k_pad, pl.hoist_cols, pl.k_h,
                                        reinterpret_cast<__nv_bfloat16*>(base + pl.off_w0h));
  LIST_LAUNCH_CHECK("build_w0h_kernel");
  const int64_t px = static_cast<int64_t>(ctx->B) * ctx->map_size * ctx->map_size;
  if ((rc = mlp_tc_project(w, lay.map_off, 0, 1, ctx->map_channels, ctx->maps, ctx->map_channels, px, base + pl.off_pmap, st)))
    return rc;
  for (int h = 0; h < pl.nh; ++h) {
    const int l = pl.lev[h];
    const int64_t R = ctx->vol_res[l];
    const int C = ctx->vol_ch[l];
    if ((rc = mlp_tc_project(w, lay.vol_off[l], C, LIST_NUM_DISP, C, ctx->vols[l], C, ctx->B * R * R * R,
                             base + pl.off_pvol[h], st)))
      return rc;
  }
  return LIST_OK;
}

// The non-hoisted levels / tail of the row for hoist_rest_kernel; LIST_ENOSYS if the mapping does not fit.
static int build_rest(const ListCtx* ctx, const Plan& pl, const ListLayout& lay, int image, RestParams* out) {
  RestParams& r = *out;
  int nl = 0, items = 0;
  const int shift = pl.hoist_cols - kN0;               // column in the full layout -> column in the hoisted row
  bool hoisted[LIST_MAX_LEVELS] = {};
  for (int h = 0; h < pl.nh; ++h) hoisted[pl.lev[h]] = true;
  int tail0 = lay.xyz_off;
  for (int pass = 0; pass < 2; ++pass) {               // vector levels (layout order) first, then scalar levels
    for (int l = ctx->n_levels - 1; l >= 0; --l) {
      if (hoisted[l]) continue;
      const bool vec = ctx->vol_ch[l] % 8 == 0;
      if (vec != (pass == 0)) continue;
      if (nl >= kRestLevels) return LIST_ENOSYS;
      const size_t vox = static_cast<size_t>(ctx->vol_res[l]) * ctx->vol_res[l] * ctx->vol_res[l] * ctx->vol_ch[l];
      if (vox >= (1ull << 32)) return LIST_ENOSYS;
      r.vols[nl] = static_cast<const __nv_bfloat16*>(ctx->vols[l]) + static_cast<size_t>(image) * vox;
      r.R[nl] = ctx->vol_res[l];
      r.C[nl] = ctx->vol_ch[l];
      r.xoff[nl] = lay.vol_off[l] - shift;
      if (vec) items += LIST_NUM_DISP * (ctx->vol_ch[l] / 8);
      else if (lay.vol_off[l] < tail0) tail0 = lay.vol_off[l];
      ++nl;
    }
  }
  if (items > kRestVecThreads) return LIST_ENOSYS;
  r.nlev = nl;
  r.nvec_items = items;
  r.tail0 = tail0 - shift;
  r.xyz_off = lay.xyz_off - shift;
  r.k_h = pl.k_h;
  if (r.xyz_off + 3 - r.tail0 > 32 || r.k_h - r.tail0 > 64) return LIST_ENOSYS;
  return LIST_OK;
}

// LIST_OK if gather() covers a res^3 grid of this configuration.
int check_gather(const ListCtx* ctx, const Plan& pl, int res) {
  if ((res - 1 + kPz - 1) / res + 1 > kMaxRuns) return LIST_ENOSYS;
  ListLayout lay;
  const int rc = list_feature_layout(ctx->map_channels, ctx->n_levels, ctx->vol_ch, &lay, nullptr);
  if (rc) return rc;
  RestParams r{};
  return build_rest(ctx, pl, lay, 0, &r);
}

// Feature rows X_h[count][ldx] of grid points [begin, begin+count) of image `image`.
int gather(const ListCtx* ctx, const Plan& pl, const void* buf, int image, int res, double bb_min, double bb_max,
           int64_t begin, int64_t count, void* X, int64_t ldx, cudaStream_t st) {
  if (count == 0) return LIST_OK;
  if ((res - 1 + kPz - 1) / res + 1 > kMaxRuns) return LIST_ENOSYS;
  ListLayout lay;
  const int rc = list_feature_layout(ctx->map_channels, ctx->n_levels, ctx->vol_ch, &lay, nullptr);
  if (rc) return rc;
  const char* base = static_cast<const char*>(buf);
  AddParams a{};
  a.pmap = reinterpret_cast<const __nv_bfloat16*>(base + pl.off_pmap) + static_cast<size_t>(image) * ctx->map_size * ctx->map_size * kN0;
  a.nh = pl.nh;
  for (int h = 0; h < pl.nh; ++h) {
    const size_t R = ctx->vol_res[pl.lev[h]];
    a.pvol[h] = reinterpret_cast<const __nv_bfloat16*>(base + pl.off_pvol[h]) + static_cast<size_t>(image) * R * R * R * kN0;
    a.dstride[h] = static_cast<uint32_t>(static_cast<size_t>(ctx->B) * R * R * R * kN0);
    a.R[h] = static_cast<int>(R);
  }
  for (int h = pl.nh; h < kMaxH; ++h) { a.pvol[h] = a.pvol[0]; a.dstride[h] = 0; a.R[h] = 1; }
  a.T = ctx->trans_mat + image * 12;
  a.X = static_cast<__nv_bfloat16*>(X);
  a.ldx = ldx;
  a.N = count;
  a.grid_begin = begin;
  a.S = ctx->map_size;
  a.res = res;
  a.bb_min = bb_min;
  a.bb_max = bb_max;

  RestParams r{};
  const int rc2 = build_rest(ctx, pl, lay, image, &r);
  if (rc2) return rc2;
  r.X = a.X;
  r.ldx = ldx;
  r.N = count;
  r.grid_begin = begin;
  r.res = res;
  r.bb_min = bb_min;
  r.bb_max = bb_max;

  const unsigned tiles = static_cast<unsigned>((count + kPz - 1) / kPz);
  static const int vec = []() { const char* e = getenv("LIST_B200_HOIST_VEC"); return (e && e[0] == '8') ? 8 : 4; }();
  if (vec == 8) hoist_addend_kernel<8><<<tiles, kN0 / 8, 0, st>>>(a);
  else hoist_addend_kernel<4><<<tiles, kN0 / 4, 0, st>>>(a);
  LIST_LAUNCH_CHECK("hoist_addend_kernel");
  hoist_rest_kernel<<<tiles, kRestThreads, 0, st>>>(r);
  LIST_LAUNCH_CHECK("hoist_rest_kernel");
  return LIST_OK;
}

}  // namespace hoist
}  // namespace list
