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
// from the per-query GEMM and replaces them by ONE 512-wide "addend" block (which also carries the bias b0) that
// the MLP kernel adds to fc_0's accumulator in its epilogue: the feature row shrinks from 3648 to 512 + 832 columns,
//        X_h = [ addend(512) | level 32^3 (448) | level 64^3 (224) | level 128^3 (112) | occupancy (7) | q (3) | 0-pad ]
// and fc_0 runs on K = 832 with the matching columns of W0:  relu(W0[:, 2816:] · X_h[512:] + addend) == relu(W0 · X + b0)
// in real arithmetic.  The fp32 parity path is untouched.
//
// Kernels here:
//   hoist_addend_kernel  per tile (z-line, segment of 128 steps), 128 threads, thread = 4 of the 512 addend channels:
//                        walks the z-run with the separable scheme of gather_grid.cu -- bilinear taps of the
//                        projected map (cell cache) + for each hoisted level the three W-shift classes
//                        {d=0,3,4,5,6}, {d=1}, {d=2} of the displacement table (modules.py:205-212): the five
//                        class-0 displacements share the voxel index and weight along the walk, so their
//                        (H,D)-interpolated projected columns are summed ONCE per voxel cell and a step costs one
//                        FMA per channel and class.  State per class: G1 (column at i0+1) and D = G1 - G0, a
//                        function of the cell only (bit-exact under any chunking / sharding of the grid).
//                        Steps at which a class enters a new cell are flagged in a per-step control word built
//                        once per tile, so the walk between two such steps is a branch-free FMA loop.
//   hoist_rest_kernel    the remaining feature columns (levels 32^3, 64^3, 128^3, occupancy, q, pad), bit-identical
//                        to gather_grid.cu's, in two divergence-free phases through a shared-memory column table.
#include <cstdlib>

#include "grid_common.cuh"
#include "hoist.cuh"

namespace list {

int mlp_tc_project(const ListWeights* w, int col0, int col_stride, int groups, int k, const void* X, int64_t ldx,
                   int64_t rows, void* out, cudaStream_t st);

namespace hoist {

constexpr int kTile = 128;          // max steps per tile (rest kernel)
#ifndef LIST_ADD_TILE
#define LIST_ADD_TILE 128
#endif
constexpr int kAddTile = LIST_ADD_TILE;   // steps per tile of the addend kernel (a power of two >= 128)
static_assert(kAddTile >= 128 && (kAddTile & (kAddTile - 1)) == 0, "tile decoding uses shifts");
constexpr int kN0 = 512;            // fc_0 width == addend channels
constexpr int kRestThreads = 256;
constexpr int kRestLevels = 4;      // non-hoisted levels (vector ones first)
constexpr int kMaxVec = 128;        // 16-byte items of the non-hoisted vector levels
constexpr int kMaxTail = 64;        // tail columns: scalar levels, q, zero pad

struct AddParams {
  const __nv_bfloat16* pmap;        // image's [S][S][512]
  const __nv_bfloat16* pvol[kMaxH]; // image's slab of displacement 0: [R][R][R][512]
  uint32_t dstride[kMaxH];          // elements between displacement slabs
  const float* T;
  const float* b0;                  // fc_0 bias, folded into the addend
  __nv_bfloat16* X;
  int64_t ldx;
  int S, nh;
  int R[kMaxH];
  TileMap tm;
};

struct RestParams {
  const __nv_bfloat16* vols[kRestLevels];
  int R[kRestLevels], C[kRestLevels], xoff[kRestLevels];   // xoff: first column of the level in the hoisted row
  int cells_max[kRestLevels];       // table rows per (level, displacement)
  int toff[kRestLevels];            // first float of the level's tables in shared memory
  // Tables of a level in shared memory: displacements 0, 1, 2 do not move (H, D), so they share ONE table (`cells_u` cells:
  // the union of the three W-shift classes' ranges); 3..6 have one each (cells_max cells).  dbase[d]: first float of
  // displacement d's table relative to toff.  Each table is followed by 4 floats of padding for the vector levels, so that
  // the 16-byte accesses of the lanes of a warp (one per (displacement, channel vector)) spread over the banks.
  int dbase[kRestLevels][LIST_NUM_DISP];
  int cells_u[kRestLevels];
  int g_lanes5[kRestLevels];        // phase G: kRestThreads / (5 tables * ncv)
  float inv_combos5[kRestLevels];
  int tab_floats;                   // floats of all column tables (the per-step tables follow)
  int nvl, nsl;                     // vector levels [0, nvl), scalar levels [nvl, nvl + nsl)
  int lwarp0[kRestLevels + 1];      // phase L: warps [lwarp0[i], lwarp0[i+1]) work on vector level i
  // division-free index decoding of the vector levels (ncv = C / 8 is a power of two; fast_div reciprocals)
  int lg_ncv[kRestLevels], g_lanes[kRestLevels], l_nblk[kRestLevels];
  float inv_combos[kRestLevels], inv_lnblk[kRestLevels];
  int nvec;                         // 16-byte items per row of the vector levels
  int g_items;                      // phase-G items of the vector levels
  int tail0, xyz_off, k_h;          // tail region [tail0, k_h): scalar levels, q, zero pad
  int ones;                         // 1: the three columns behind q hold 1.0 (Plan::bias_col)
  __nv_bfloat16* X;
  int64_t ldx;
  TileMap tm;
};

__device__ __forceinline__ int warp_of(int tid) { return tid >> 5; }
// floor(a / b) for 0 <= a < 2^20, 1 <= b < 2^10 through the float reciprocal (exact in that range)
__device__ __forceinline__ int fast_div(int a, float inv_b) { return static_cast<int>((static_cast<float>(a) + 0.5f) * inv_b); }

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

// Projected column of one W-shift class at voxel index xv: sum over the class's displacements and their 4 (H,D)
// corners.  corners: [7][4] of the tile and level.
template <int CLS, int V>
__device__ __forceinline__ void load_column(const __nv_bfloat16* __restrict__ pv, uint32_t dstride, const Corner* __restrict__ corners,
                                            int xv, int cv, float out[V]) {
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

// Control word of a step: bits 0-5 = class c (level c/3, W-shift class c%3) enters a new voxel cell,
// bit 6 = the 2-D sample enters a new pixel cell, bits 8-15 = steps until the next step with any of those bits.
#ifndef LIST_ADDEND_MINBLOCKS
#define LIST_ADDEND_MINBLOCKS 4        // <= 128 registers.  Measured at 256^3 (same box): 172 regs (uncapped) 21.7 ms, 128 regs 17.1 ms,
                                       // 96 regs 18.4 ms alone; pipelined with the MLP 48.5 / 42.7 / 46.8 ms per grid
#endif
template <int V>
__global__ void __launch_bounds__(kN0 / V, LIST_ADDEND_MINBLOCKS) hoist_addend_kernel(const AddParams p) {
  constexpr int NT = kN0 / V;
  constexpr int NC = kMaxH * 3;
  __shared__ __align__(16) float s_w[kAddTile][12];        // per step: w0 of the 6 classes, w00 w01 w10 w11, 2 pad
  __shared__ int s_i0[NC][kAddTile];
  __shared__ int s_xy[kAddTile];                           // y0 << 16 | x0
  __shared__ uint32_t s_ctl[kAddTile];
  __shared__ uint32_t s_any[kAddTile / 32];
  __shared__ Corner s_corner[kMaxH][LIST_NUM_DISP * 4];
  const int tid = threadIdx.x;
  TileSpan t;
  if (!tile_span(p.tm, blockIdx.x, t)) return;

  // ---- phase 0a: per step 2-D cell / weights and voxel index / weight of every class ----
  for (int s = tid; s < kAddTile; s += NT) {
    int x0 = 0, y0 = 0;
    float w00 = 0.f, w01 = 0.f, w10 = 0.f, w11 = 0.f;                    // NaN grid -> all taps out of bounds
    float q[3] = {0.f, t.qy, t.qz};
    if (s >= t.s_lo && s < t.s_hi) {
      q[0] = step_q0(p.tm, t, s);
      float ix, iy, h[3];
      localise(q, p.T, p.S, ix, iy, h);
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
    }
    s_xy[s] = (y0 << 16) | x0;
    s_w[s][6] = w00; s_w[s][7] = w01; s_w[s][8] = w10; s_w[s][9] = w11;
    s_w[s][10] = 0.f; s_w[s][11] = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int hh = c / 3, cls = c % 3;
      int i0 = 0;
      float w0 = 0.f;
      if (hh < p.nh) {
        const Axis3 ax = axis_border(cls == 0 ? q[0] : q[0] + class_shift(cls), hh == 0 ? p.R[0] : p.R[1]);
        i0 = ax.i0; w0 = ax.w0;
      }
      s_i0[c][s] = i0;
      s_w[s][c] = w0;
    }
  }
  if (tid < p.nh * LIST_NUM_DISP) {                                      // (H,D) corners per (level, displacement)
    const int d = tid % LIST_NUM_DISP, hh = tid / LIST_NUM_DISP;
    uint32_t base[4];
    float wyz[4];
    tile_corners(t.qy, t.qz, d, hh == 0 ? p.R[0] : p.R[1], kN0, base, wyz);
#pragma unroll
    for (int k = 0; k < 4; ++k) s_corner[hh][d * 4 + k] = Corner{base[k], wyz[k]};
  }
  __syncthreads();
  // ---- phase 0b: control words ----
  uint32_t myctl[kAddTile / NT];
#pragma unroll
  for (int it = 0; it < kAddTile / NT; ++it) {
    const int s = tid + it * NT;
    uint32_t ctl = 0;
    if (s >= t.s_lo && s < t.s_hi) {
      const bool first = s == t.s_lo;
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (c / 3 < p.nh && (first || s_i0[c][s] != s_i0[c][s - 1])) ctl |= 1u << c;
      if (first || s_xy[s] != s_xy[s - 1]) ctl |= 1u << 6;
    }
    myctl[it] = ctl;
    const uint32_t b = __ballot_sync(0xffffffffu, ctl != 0);
    if ((tid & 31) == 0) s_any[s >> 5] = b;
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < kAddTile / NT; ++it) {
    const int s = tid + it * NT;
    int nxt = kAddTile - s;                                                 // steps to the next break (clipped to s_hi later)
    const int w = s >> 5, bit = s & 31;
    const uint32_t rest = bit == 31 ? 0u : (s_any[w] >> (bit + 1));
    if (rest) nxt = __ffs(rest);
    else {
      for (int w2 = w + 1; w2 < kAddTile / 32; ++w2)
        if (s_any[w2]) { nxt = w2 * 32 + __ffs(s_any[w2]) - 1 - s; break; }
    }
    s_ctl[s] = myctl[it] | (static_cast<uint32_t>(nxt) << 8);       // distance < 2^16
  }
  __syncthreads();

  const int cv = tid;
  float G1[NC][V], D[NC][V], gsum[V];
  float v00[V], v01[V], v10[V], v11[V], bias[V];
  int cur[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    cur[c] = -2;
#pragma unroll
    for (int j = 0; j < V; ++j) { G1[c][j] = 0.f; D[c][j] = 0.f; }
  }
#pragma unroll
  for (int j = 0; j < V; ++j) { v00[j] = v01[j] = v10[j] = v11[j] = 0.f; bias[j] = __ldg(p.b0 + cv * V + j); gsum[j] = bias[j]; }

  __nv_bfloat16* __restrict__ dst = p.X + (t.g_tile0 + t.s_lo - p.tm.begin) * p.ldx + cv * V;
  const int lim = p.S - 1;
  int s = t.s_lo;
  while (s < t.s_hi) {
    const uint32_t ctl = s_ctl[s];
    if (ctl & 0x3fu) {
      const bool first = s == t.s_lo;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        if (ctl & (1u << c)) {
          const int hh = c / 3;
          const int R = hh == 0 ? p.R[0] : p.R[1];
          const int i0 = s_i0[c][s];
          const int i1 = min(i0 + 1, R - 1);
          const Corner* corners = s_corner[hh];
          const __nv_bfloat16* pv = hh == 0 ? p.pvol[0] : p.pvol[1];
          const uint32_t ds = hh == 0 ? p.dstride[0] : p.dstride[1];
          float g0v[V], g1v[V];
          if (!first && i0 == cur[c] + 1) {
#pragma unroll
            for (int j = 0; j < V; ++j) g0v[j] = G1[c][j];
          } else {
            if (c % 3 == 0) load_column<0, V>(pv, ds, corners, i0, cv, g0v);
            else if (c % 3 == 1) load_column<1, V>(pv, ds, corners, i0, cv, g0v);
            else load_column<2, V>(pv, ds, corners, i0, cv, g0v);
          }
          if (i1 != i0) {
            if (c % 3 == 0) load_column<0, V>(pv, ds, corners, i1, cv, g1v);
            else if (c % 3 == 1) load_column<1, V>(pv, ds, corners, i1, cv, g1v);
            else load_column<2, V>(pv, ds, corners, i1, cv, g1v);
          } else {
#pragma unroll
            for (int j = 0; j < V; ++j) g1v[j] = g0v[j];
          }
#pragma unroll
          for (int j = 0; j < V; ++j) { D[c][j] = g1v[j] - g0v[j]; G1[c][j] = g1v[j]; }
          cur[c] = i0;
        }
      }
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float a = bias[j];
#pragma unroll
        for (int c = 0; c < NC; ++c) a += G1[c][j];
        gsum[j] = a;
      }
    }
    if (ctl & 0x40u) {
      const int xy = s_xy[s];
      const int cx = xy & 0xffff, cy = xy >> 16;
      const int x1 = min(cx + 1, lim), y1 = min(cy + 1, lim);
      const __nv_bfloat16* __restrict__ pm = p.pmap + cv * V;
      loadv<V>(pm + (static_cast<size_t>(cy) * p.S + cx) * kN0, v00);
      loadv<V>(pm + (static_cast<size_t>(cy) * p.S + x1) * kN0, v01);
      loadv<V>(pm + (static_cast<size_t>(y1) * p.S + cx) * kN0, v10);
      loadv<V>(pm + (static_cast<size_t>(y1) * p.S + x1) * kN0, v11);
    }
    const int n = min(static_cast<int>((ctl >> 8) & 0xffffu), t.s_hi - s);
#pragma unroll 2
    for (int k = 0; k < n; ++k, ++s, dst += p.ldx) {
      const float4 wa = *reinterpret_cast<const float4*>(&s_w[s][0]);
      const float4 wb = *reinterpret_cast<const float4*>(&s_w[s][4]);
      const float2 wc = *reinterpret_cast<const float2*>(&s_w[s][8]);
      const float w3[NC] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y};
      float acc[V];
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float a = fmaf(v00[j], wb.z, gsum[j]);
        a = fmaf(v01[j], wb.w, a);
        a = fmaf(v10[j], wc.x, a);
        a = fmaf(v11[j], wc.y, a);
#pragma unroll
        for (int c = 0; c < NC; ++c) a = fmaf(-w3[c], D[c][j], a);
        acc[j] = a;
      }
      storev<V>(dst, acc);
    }
  }
}

// ------------------------------------------------------------------ the remaining columns
// Two phases per tile:
//   G: for every (level, displacement) the (H,D)-interpolated column  G[xv][c] = sum_4 (wy*wz) V[z_k][y_k][xv][c]
//      of every voxel index xv the tile's steps touch (+ its right neighbour), fp32, into shared memory;
//   L: out[s][col] = G[x0] + w1 * (G[x0+1] - G[x0]); a thread owns one 16-byte vector of the row and a contiguous
//      block of steps, keeps (G[x0], G[x0+1]-G[x0]) in registers and refreshes them from the table when x0 moves;
//      consecutive lanes write consecutive vectors of one row.
// Same arithmetic, in the same order, as the z-run walker of gather_grid.cu, so the columns are bit-identical to it.

__global__ void __launch_bounds__(kRestThreads) hoist_rest_kernel(const RestParams p) {
  extern __shared__ __align__(16) float s_tab[];
  __shared__ float s_q0[kTile];
  // per (level, class, step), sized by the tile length and placed behind the column tables in dynamic shared memory:
  // voxel index relative to the class's first one (fits a byte: cells_max <= 255) and the weight of its right neighbour
  uint2* const s_sw = reinterpret_cast<uint2*>(s_tab + p.tab_floats);               // [nlev*3][kPz]: {bits of w1, rel}
  __shared__ int s_first[kRestLevels][3], s_ncell[kRestLevels][3];
  __shared__ int s_coff[kRestLevels][3];                // first cell of the class relative to the shared table of d = 0, 1, 2
  __shared__ int s_firstu[kRestLevels], s_ncellu[kRestLevels];
  __shared__ Corner s_cor[kRestLevels * LIST_NUM_DISP][4];
  __shared__ int s_tailtab[kMaxTail];                   // scalar columns: table base | (level*3+class) << 24
  const int tid = threadIdx.x;
  TileSpan t;
  if (!tile_span(p.tm, blockIdx.x, t)) return;
  const int nlev = p.nvl + p.nsl;
  const int kPz = p.tm.kPz;

  // ---- phase 0: per-step voxel index / weight per (level, class); corners; descriptors ----
  if (tid < nlev * 3) {
    const int li = tid / 3, cls = tid % 3;
    int a = 0, n = 0, lo = 1 << 30, hi = 0;                           // this class's range and the union of the three
    for (int k = 0; k < 3; ++k) {
      const float sh = class_shift(k);
      const int ak = axis_border(step_q0(p.tm, t, t.s_lo) + sh, p.R[li]).i0;
      const int bk = axis_border(step_q0(p.tm, t, t.s_hi - 1) + sh, p.R[li]).i0;
      const int nk = min(min(bk + 1, p.R[li] - 1) - ak + 1, p.cells_max[li]);
      if (k == cls) { a = ak; n = nk; }
      lo = min(lo, ak);
      hi = max(hi, ak + nk);
    }
    s_first[li][cls] = a;
    s_ncell[li][cls] = n;
    s_coff[li][cls] = a - lo;
    if (cls == 0) { s_firstu[li] = lo; s_ncellu[li] = min(hi - lo, p.cells_u[li]); }
  } else if (tid >= 32 && tid < 32 + nlev * LIST_NUM_DISP) {
    const int pr = tid - 32, li = pr / LIST_NUM_DISP, d = pr % LIST_NUM_DISP;
    uint32_t base[4];
    float wyz[4];
    tile_corners(t.qy, t.qz, d, p.R[li], static_cast<uint32_t>(p.C[li]), base, wyz);
#pragma unroll
    for (int k = 0; k < 4; ++k) s_cor[pr][k] = Corner{base[k], wyz[k]};
  }
  for (int s = tid; s < kPz; s += kRestThreads) s_q0[s] = (s >= t.s_lo && s < t.s_hi) ? step_q0(p.tm, t, s) : 0.f;
  const int nscal = p.xyz_off - p.tail0;
  for (int j = tid; j < nscal; j += kRestThreads) {                     // scalar column -> table base, class
    const int col = p.tail0 + j;
    int val = 0;
    for (int li = p.nvl; li < nlev; ++li) {
      const int rel = col - p.xoff[li];
      if (rel >= 0 && rel < LIST_NUM_DISP * p.C[li]) {
        const int d = rel / p.C[li], c = rel % p.C[li];
        val = (p.toff[li] + p.dbase[li][d] + c) | (d < 3 ? 1 << 23 : 0) | ((li * 3 + shift_class(d)) << 24);   // bit 23: shared table
      }
    }
    s_tailtab[j] = val;
  }
  __syncthreads();
  for (int i = tid; i < nlev * 3 * kPz; i += kRestThreads) {
    const int s = i & (kPz - 1), lc = i >> p.tm.lg_kpz;
    const int li = lc / 3, cls = lc % 3;
    int rel = 0;
    float w1 = 0.f;
    if (s >= t.s_lo && s < t.s_hi) {
      const Axis3 ax = axis_border(cls == 0 ? s_q0[s] : s_q0[s] + class_shift(cls), p.R[li]);
      rel = min(ax.i0 - s_first[li][cls], p.cells_max[li] - 1);
      w1 = ax.w1;
    }
    s_sw[lc * kPz + s] = make_uint2(__float_as_uint(w1), static_cast<uint32_t>(rel));
  }

  // ---- phase G: vector levels.  A thread keeps one (displacement, channel vector) -- corners, weights and base pointers
  //      stay in registers -- and strides over the voxel cells of its class ----
  {
    for (int li = 0; li < p.nvl; ++li) {
      const int C = p.C[li], ncv = C >> 3;
      const int combos = 5 * ncv;                                       // tables 0 (shared by d = 0, 1, 2), 3, 4, 5, 6
      const int lanes = p.g_lanes5[li];                                 // kRestThreads / combos
      const int lane0 = fast_div(tid, p.inv_combos5[li]), combo = tid - lane0 * combos;
      if (lane0 >= lanes) continue;
      const int tt = combo >> p.lg_ncv[li], cvv = combo - (tt << p.lg_ncv[li]);
      const int d = tt == 0 ? 0 : tt + 2;
      const Corner* cor = s_cor[li * LIST_NUM_DISP + d];
      const Corner c0 = cor[0], c1 = cor[1], c2 = cor[2], c3 = cor[3];
      const int ncell = tt == 0 ? s_ncellu[li] : s_ncell[li][0];
      const int first = tt == 0 ? s_firstu[li] : s_first[li][0];
      const __nv_bfloat16* __restrict__ src = p.vols[li] + static_cast<uint32_t>(first + lane0) * C + cvv * 8;
      float* __restrict__ dstt = s_tab + p.toff[li] + p.dbase[li][d] + lane0 * C + cvv * 8;
      const int sstep = lanes * C;
      // the four corner loads of the next cell are issued before this cell's arithmetic (two register sets)
      auto issue = [&](const __nv_bfloat16* sp, uint4 (&r)[4]) {
        r[0] = __ldg(reinterpret_cast<const uint4*>(sp + c0.base));
        r[1] = __ldg(reinterpret_cast<const uint4*>(sp + c1.base));
        r[2] = __ldg(reinterpret_cast<const uint4*>(sp + c2.base));
        r[3] = __ldg(reinterpret_cast<const uint4*>(sp + c3.base));
      };
      auto work = [&](const uint4 (&r)[4], float* dp) {
        const uint32_t w0[4] = {r[0].x, r[0].y, r[0].z, r[0].w}, w1[4] = {r[1].x, r[1].y, r[1].z, r[1].w};
        const uint32_t w2[4] = {r[2].x, r[2].y, r[2].z, r[2].w}, w3[4] = {r[3].x, r[3].y, r[3].z, r[3].w};
        float g[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 e0 = bf16x2_to_f2(w0[j]), e1 = bf16x2_to_f2(w1[j]), e2 = bf16x2_to_f2(w2[j]), e3 = bf16x2_to_f2(w3[j]);
          g[2 * j] = fmaf(e3.x, c3.w, fmaf(e2.x, c2.w, fmaf(e1.x, c1.w, e0.x * c0.w)));
          g[2 * j + 1] = fmaf(e3.y, c3.w, fmaf(e2.y, c2.w, fmaf(e1.y, c1.w, e0.y * c0.w)));
        }
        *reinterpret_cast<float4*>(dp) = make_float4(g[0], g[1], g[2], g[3]);
        *reinterpret_cast<float4*>(dp + 4) = make_float4(g[4], g[5], g[6], g[7]);
      };
      uint4 ra[4], rb[4];
      int c = lane0;
      if (c < ncell) issue(src, ra);
#pragma unroll 1
      while (c < ncell) {
        if (c + lanes < ncell) issue(src + sstep, rb);
        work(ra, dstt);
        c += lanes; src += sstep; dstt += sstep;
        if (c >= ncell) break;
        if (c + lanes < ncell) issue(src + sstep, ra);
        work(rb, dstt);
        c += lanes; src += sstep; dstt += sstep;
      }
    }
    // scalar levels (C % 8 != 0): one (displacement, cell, channel) value per item
    for (int li = p.nvl; li < nlev; ++li) {
      const int C = p.C[li];
      const int per_t = p.cells_u[li] * C, cnt = 5 * per_t;               // five tables, enumerated with the shared one's size
      const float inv_per_t = 1.0f / static_cast<float>(per_t), inv_c = 1.0f / static_cast<float>(C);
      const __nv_bfloat16* __restrict__ vol = p.vols[li];
      for (int it = tid; it < cnt; it += kRestThreads) {
        const int tt = fast_div(it, inv_per_t);
        const int rem = it - tt * per_t;
        const int c = fast_div(rem, inv_c);
        const int ch = rem - c * C;
        const int d = tt == 0 ? 0 : tt + 2;
        if (c >= (tt == 0 ? s_ncellu[li] : s_ncell[li][0])) continue;
        const Corner* cor = s_cor[li * LIST_NUM_DISP + d];
        const __nv_bfloat16* src = vol + static_cast<uint32_t>((tt == 0 ? s_firstu[li] : s_first[li][0]) + c) * C + ch;
        float r = __bfloat162float(src[cor[0].base]) * cor[0].w;
#pragma unroll
        for (int k = 1; k < 4; ++k) r = fmaf(__bfloat162float(src[cor[k].base]), cor[k].w, r);
        s_tab[p.toff[li] + p.dbase[li][d] + c * C + ch] = r;
      }
    }
  }
  __syncthreads();

  // ---- phase L: vector columns.  Every warp works on ONE level (host plan: lwarp0 / lwarps), so all its lanes refresh
  //      their registers at that level's cadence; thread = (16-byte vector v of the level, block of steps) ----
  const int nsteps = t.s_hi - t.s_lo;
  __nv_bfloat16* __restrict__ Xb = p.X + (t.g_tile0 + t.s_lo - p.tm.begin) * p.ldx;
  {
    int li = 0;
    while (li + 1 < p.nvl && warp_of(tid) >= p.lwarp0[li + 1]) ++li;
    const int C = p.C[li], ncv = C >> 3, cm = p.cells_max[li];
    const int nvec = LIST_NUM_DISP * ncv;
    const int T = (p.lwarp0[li + 1] - p.lwarp0[li]) * 32;               // threads of this level
    const int lt = tid - p.lwarp0[li] * 32;
    const int nblk = p.l_nblk[li];                                      // max(1, T / nvec)
    const int per = fast_div(nsteps + nblk - 1, p.inv_lnblk[li]);
    const float inv_nvec = p.inv_combos[li];                            // nvec == combos of the level
    for (int it = lt; it < nvec * nblk; it += T) {
      const int blk = fast_div(it, inv_nvec), v = it - blk * nvec;
      const int d = v >> p.lg_ncv[li], cvv = v - (d << p.lg_ncv[li]);
      const int cls = shift_class(d);
      const int sr0 = blk * per, sr1 = min(nsteps, sr0 + per);
      const uint2* __restrict__ sw = s_sw + (li * 3 + cls) * kPz + t.s_lo;
      const int last = s_ncell[li][cls] - 1;
      const float* __restrict__ tab = s_tab + p.toff[li] + p.dbase[li][d] + (d < 3 ? s_coff[li][cls] * C : 0) + cvv * 8;
      // G0 / G1: the columns of the current voxel cell and of its right neighbour, Dv = G1 - G0 (packed pairs).  When the
      // walk moves on by one cell the old right neighbour becomes the new left one: only one column is read again.
      float2 G0[4], G1[4], Dv[4];
      int cc = -2;
      __nv_bfloat16* __restrict__ dst = Xb + static_cast<int64_t>(sr0) * p.ldx + p.xoff[li] + d * C + cvv * 8;
      const float2 minus1 = make_float2(-1.f, -1.f);
      for (int sr = sr0; sr < sr1; ++sr, dst += p.ldx) {
        const uint2 e = sw[sr];
        const int c = static_cast<int>(e.y);
        if (c != cc) {
          if (c == cc + 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) G0[j] = G1[j];
          } else {
            const float4 a0 = *reinterpret_cast<const float4*>(tab + c * C), a1 = *reinterpret_cast<const float4*>(tab + c * C + 4);
            G0[0] = make_float2(a0.x, a0.y); G0[1] = make_float2(a0.z, a0.w); G0[2] = make_float2(a1.x, a1.y); G0[3] = make_float2(a1.z, a1.w);
          }
          const float* g1p = tab + min(c + 1, last) * C;
          const float4 b0 = *reinterpret_cast<const float4*>(g1p), b1 = *reinterpret_cast<const float4*>(g1p + 4);
          G1[0] = make_float2(b0.x, b0.y); G1[1] = make_float2(b0.z, b0.w); G1[2] = make_float2(b1.x, b1.y); G1[3] = make_float2(b1.z, b1.w);
#pragma unroll
          for (int j = 0; j < 4; ++j) Dv[j] = ffma2(G0[j], minus1, G1[j]);          // G1 - G0, one rounding
          cc = c;
        }
        const float w1 = __uint_as_float(e.x);
        const float2 ww = make_float2(w1, w1);
        uint4 o;
        { const float2 r = ffma2(Dv[0], ww, G0[0]); o.x = pack_bf16x2(r.x, r.y); }
        { const float2 r = ffma2(Dv[1], ww, G0[1]); o.y = pack_bf16x2(r.x, r.y); }
        { const float2 r = ffma2(Dv[2], ww, G0[2]); o.z = pack_bf16x2(r.x, r.y); }
        { const float2 r = ffma2(Dv[3], ww, G0[3]); o.w = pack_bf16x2(r.x, r.y); }
        *reinterpret_cast<uint4*>(dst) = o;
      }
    }
  }
  // ---- tail: scalar levels, q, zero pad.  item = (step, 8-column group); the groups beyond the q columns are zeros
  //      and are enumerated separately so that a warp never mixes the two kinds ----
  {
    const int ngrp = (p.k_h - p.tail0) >> 3;             // tail0 and k_h are multiples of 8
    const int nlive = min(ngrp, (nscal + 3 + 3 * p.ones + 7) >> 3);   // groups holding scalar-level, q or bias columns
    const int ndead = ngrp - nlive;
    __nv_bfloat16* __restrict__ Xt = Xb + p.tail0;
    if (ndead > 0) {
      const float inv = 1.0f / static_cast<float>(ndead);
      for (int it = tid; it < nsteps * ndead; it += kRestThreads) {
        const int sr = fast_div(it, inv);
        const int gq = nlive + it - sr * ndead;
        *reinterpret_cast<uint4*>(Xt + static_cast<int64_t>(sr) * p.ldx + gq * 8) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    const float inv_live = 1.0f / static_cast<float>(nlive);
    for (int it = tid; it < nsteps * nlive; it += kRestThreads) {
      const int sr = fast_div(it, inv_live);
      const int gq = it - sr * nlive;
      const int s = t.s_lo + sr;
      float out[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = gq * 8 + j;                      // relative to tail0
        float val = 0.f;
        if (col < nscal) {
          const int tt = s_tailtab[col];
          const int lc = tt >> 24;
          const int li = lc / 3, cls = lc - li * 3;
          const int tab = (tt & 0x7fffff) + ((tt >> 23) & 1) * s_coff[li][cls] * p.C[li];
          const uint2 e = s_sw[lc * kPz + s];
          const int c = static_cast<int>(e.y);
          const int c1 = min(c + 1, s_ncell[li][cls] - 1);
          const float g0 = s_tab[tab + c * p.C[li]], g1 = s_tab[tab + c1 * p.C[li]];
          val = fmaf(g1 - g0, __uint_as_float(e.x), g0);
        } else if (col < nscal + 3) {
          const int a = col - nscal;
          val = a == 0 ? s_q0[s] : (a == 1 ? t.qy : t.qz);
        } else if (p.ones && col < nscal + 6) {
          val = 1.0f;
        }
        out[j] = val;
      }
      store8(Xt + static_cast<int64_t>(sr) * p.ldx + gq * 8, out);
    }
  }
}

// ------------------------------------------------------------------ host side
// Which part of the row is hoisted: the maps and the leading vector levels of the layout (coarsest first)
// whose projected volumes stay small (R <= 16) and whose channel count feeds the tensor-core projection (C % 64).
int make_plan(const ListCtx* ctx, const ListWeights* w, Plan* pl, int max_levels, int max_res) {
  if (ctx->dtype != LIST_BF16 || w->dtype != LIST_BF16 || w->n0 != kN0) return LIST_ENOSYS;
  if (ctx->map_channels % 64 != 0) return LIST_ENOSYS;
  if (max_levels > kMaxLev) max_levels = kMaxLev;
  ListLayout lay;
  const int rc = list_feature_layout(ctx->map_channels, ctx->n_levels, ctx->vol_ch, &lay, nullptr);
  if (rc) return rc;
  if (lay.map_off != 0 || w->k_pad != lay.k_pad) return LIST_ENOSYS;
  pl->nh = 0;
  pl->rpl = 0;
  int cols = ctx->map_channels;
  for (int l = ctx->n_levels - 1; l >= 0 && pl->nh < max_levels; --l) {       // layout order of the vector levels
    if (ctx->vol_ch[l] % 8) continue;
    if (ctx->vol_res[l] > max_res || ctx->vol_ch[l] % 64 != 0 || lay.vol_off[l] != cols) break;
    pl->rowbase[pl->nh] = pl->rpl;
    pl->rpl += 3 * ctx->vol_res[l];
    pl->lev[pl->nh++] = l;
    cols += LIST_NUM_DISP * ctx->vol_ch[l];
  }
  if (cols % 64 != 0) return LIST_ENOSYS;
  pl->hoist_cols = cols;
  pl->k_h = kN0 + (lay.k_pad - cols);
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  size_t off = 0;
  pl->off_pmap = off; off += up(static_cast<size_t>(ctx->B) * ctx->map_size * ctx->map_size * kN0 * 2);
  for (int h = 0; h < pl->nh; ++h) {
    const size_t R = ctx->vol_res[pl->lev[h]];
    if (static_cast<size_t>(LIST_NUM_DISP) * ctx->B * R * R * R * kN0 >= (1ull << 32)) return LIST_ENOSYS;
    pl->off_pvol[h] = off;
    off += up(static_cast<size_t>(LIST_NUM_DISP) * ctx->B * R * R * R * kN0 * 2);
  }
  pl->off_zero = off; off += up(kN0 * 2);                                  // one all-zero row (padding rows of grid_tc.cu)
  pl->off_w0r = off; off += up(static_cast<size_t>(kN0) * (lay.k_pad - cols) * 2);
  pl->bias_col = lay.xyz_off + 3 - cols;
  if (pl->bias_col + 3 > lay.k_pad - cols) return LIST_ENOSYS;            // no pad columns left for the bias
  pl->total = off;
  return LIST_OK;
}

// W0r[n][c] = W0[n][hoist_cols + c], with the bf16 hi / mid / lo parts of b0[n] in columns [bias_col, +3).
__global__ void w0r_kernel(const __nv_bfloat16* __restrict__ w0, int k_pad, int hoist_cols, int k_f, int bias_col,
                           const float* __restrict__ b0, __nv_bfloat16* __restrict__ w0r) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < k_f; c += blockDim.x) {
    __nv_bfloat16 v = w0[static_cast<size_t>(n) * k_pad + hoist_cols + c];
    if (c >= bias_col && c < bias_col + 3) {
      float r = b0[n];
      __nv_bfloat16 part = __float2bfloat16_rn(r);
      for (int i = 0; i < c - bias_col; ++i) {
        r -= __bfloat162float(part);
        part = __float2bfloat16_rn(r);
      }
      v = part;
    }
    w0r[static_cast<size_t>(n) * k_f + c] = v;
  }
}

// Projects the maps and the hoisted levels of every image through their W0 blocks.
int prepare(const ListCtx* ctx, const ListWeights* w, const Plan& pl, void* buf, cudaStream_t st) {
  char* base = static_cast<char*>(buf);
  ListLayout lay;
  int rc = list_feature_layout(ctx->map_channels, ctx->n_levels, ctx->vol_ch, &lay, nullptr);
  if (rc) return rc;
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
  LIST_CUDA(cudaMemsetAsync(base + pl.off_zero, 0, kN0 * 2, st));
  w0r_kernel<<<kN0, 128, 0, st>>>(static_cast<const __nv_bfloat16*>(w->w0), w->k_pad, pl.hoist_cols, pl.k_h - kN0, pl.bias_col, w->b0,
                                  reinterpret_cast<__nv_bfloat16*>(base + pl.off_w0r));
  LIST_LAUNCH_CHECK("w0r_kernel");
  return LIST_OK;
}

// Tile size and shared-memory plan of hoist_rest_kernel for a res^3 grid; LIST_ENOSYS if the mapping does not fit.
static int build_rest(const ListCtx* ctx, const Plan& pl, const ListLayout& lay, int image, int res, RestParams* out, size_t* smem) {
  RestParams& r = *out;
  const int shift = pl.hoist_cols - kN0;               // column in the full layout -> column in the hoisted row
  bool hoisted[LIST_MAX_LEVELS] = {};
  for (int h = 0; h < pl.nh; ++h) hoisted[pl.lev[h]] = true;
  int nl = 0, tail0 = lay.xyz_off;
  r.nvl = r.nsl = r.nvec = 0;
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
      if (vec) {
        const int ncv = ctx->vol_ch[l] / 8;
        if (ncv & (ncv - 1)) return LIST_ENOSYS;       // hoist_rest_kernel decodes items with shifts
        r.nvec += LIST_NUM_DISP * ncv;
        ++r.nvl;
      }
      else { if (lay.vol_off[l] < tail0) tail0 = lay.vol_off[l]; ++r.nsl; }
      ++nl;
    }
  }
  if (r.nvec > kMaxVec || r.nvec == 0) return LIST_ENOSYS;
  {
    // phase-L warp plan: warps per vector level in proportion to (vectors) x (cost of a step, which grows with the
    // number of voxel cells a step crosses), at least one each
    const int nw = kRestThreads / 32;
    if (r.nvl > nw) return LIST_ENOSYS;
    double cost[kRestLevels] = {}, total = 0;
    for (int i = 0; i < r.nvl; ++i) {
      const double cells_per_step = res > 1 ? static_cast<double>(r.R[i] - 1) / (res - 1) : 1.0;
      cost[i] = LIST_NUM_DISP * (r.C[i] / 8) * (20.0 + 22.0 * (cells_per_step < 1.0 ? cells_per_step : 1.0));
      total += cost[i];
    }
    int w[kRestLevels] = {}, used = 0;
    for (int i = 0; i < r.nvl; ++i) { w[i] = static_cast<int>(cost[i] / total * nw + 0.5); if (w[i] < 1) w[i] = 1; used += w[i]; }
    while (used > nw) { int k = 0; for (int i = 1; i < r.nvl; ++i) if (w[i] > w[k]) k = i; --w[k]; --used; }
    while (used < nw) { int k = 0; for (int i = 1; i < r.nvl; ++i) if (cost[i] / w[i] > cost[k] / w[k]) k = i; ++w[k]; ++used; }
    r.lwarp0[0] = 0;
    for (int i = 0; i < r.nvl; ++i) r.lwarp0[i + 1] = r.lwarp0[i] + w[i];
    for (int i = 0; i < r.nvl; ++i) {
      const int ncv = r.C[i] / 8, combos = LIST_NUM_DISP * ncv;
      int lg = 0;
      while ((1 << lg) < ncv) ++lg;
      r.lg_ncv[i] = lg;
      r.g_lanes[i] = kRestThreads / combos;
      r.inv_combos[i] = 1.0f / static_cast<float>(combos);
      r.g_lanes5[i] = kRestThreads / (5 * ncv);
      r.inv_combos5[i] = 1.0f / static_cast<float>(5 * ncv);
      const int nblk = (w[i] * 32) / combos > 1 ? (w[i] * 32) / combos : 1;
      r.l_nblk[i] = nblk;
      r.inv_lnblk[i] = 1.0f / static_cast<float>(nblk);
    }
  }
  r.tail0 = tail0 - shift;
  r.xyz_off = lay.xyz_off - shift;
  r.k_h = pl.k_h;
  if (r.tail0 % 8 != 0 || r.k_h - r.tail0 > kMaxTail || r.xyz_off + 3 > r.k_h) return LIST_ENOSYS;
  // the vector columns must be exactly [kN0, tail0): the kernel writes nothing else there
  if (kN0 + r.nvec * 8 != r.tail0) return LIST_ENOSYS;
  // largest tile whose column tables leave room for three CTAs per SM (the kernel is latency bound: warps matter more
  // than the per-tile overhead of shorter tiles)
  for (int kpz = kTile; kpz >= 16; kpz >>= 1) {
    size_t floats = 0;
    int items = 0;
    for (int i = 0; i < nl; ++i) {
      const int span = res > 1 ? static_cast<int>((static_cast<int64_t>(kpz - 1) * (r.R[i] - 1)) / (res - 1)) : 0;
      int cm = span + 3;
      if (cm > r.R[i]) cm = r.R[i];
      r.cells_max[i] = cm;
      // shared table of d = 0, 1, 2: the classes' first cells differ by at most the W shift in cells (+1 for rounding)
      int cu = cm + 2 * (static_cast<int>(kDisplacement * 0.5f * static_cast<float>(r.R[i] - 1)) + 2);
      if (cu > r.R[i]) cu = r.R[i];
      r.cells_u[i] = cu;
      r.toff[i] = static_cast<int>(floats);
      const int pad = i < r.nvl ? 4 : 0;
      int off = 0;
      for (int d = 0; d < LIST_NUM_DISP; ++d) {
        r.dbase[i][d] = d < 3 ? 0 : off;
        if (d == 2) off = cu * r.C[i] + pad;
        else if (d > 2) off += cm * r.C[i] + pad;
      }
      floats += static_cast<size_t>(off);
      if (i < r.nvl) items += 5 * cu * (r.C[i] / 8);
    }
    static const size_t budget = []() {                // LIST_B200_REST_SMEM_KB: table budget per CTA (tuning aid)
      const char* e = getenv("LIST_B200_REST_SMEM_KB");
      const long v = e ? atol(e) : 0;
      return static_cast<size_t>(v >= 8 && v <= 200 ? v : 64) * 1024;   // measured at 256^3: 96 KB 11.4 ms, 64 KB 10.9 ms, 48 KB 12.1 ms
    }();
    bool fits_byte = true;
    for (int i = 0; i < nl; ++i) fits_byte = fits_byte && r.cells_max[i] <= 255;
    if (floats * 4 <= budget && floats < (1u << 24) && items < (1 << 20) && fits_byte) {
      r.g_items = items;
      r.tm.kPz = kpz;
      r.tab_floats = static_cast<int>((floats + 3) / 4 * 4);
      *smem = static_cast<size_t>(r.tab_floats) * 4 + static_cast<size_t>(kRestLevels) * 3 * kpz * 8;   // + the per-step {w1, rel} tables
      return LIST_OK;
    }
  }
  return LIST_ENOSYS;
}

// LIST_OK if gather() covers a res^3 grid of this configuration.
int check_gather(const ListCtx* ctx, const Plan& pl, int res) {
  ListLayout lay;
  const int rc = list_feature_layout(ctx->map_channels, ctx->n_levels, ctx->vol_ch, &lay, nullptr);
  if (rc) return rc;
  RestParams r{};
  size_t smem = 0;
  return build_rest(ctx, pl, lay, 0, res, &r, &smem);
}

// Feature rows X_h[count][ldx] of grid points [begin, begin+count) of image `image`.
int gather(const ListCtx* ctx, const ListWeights* w, const Plan& pl, const void* buf, int image, int res, double bb_min, double bb_max,
           int64_t begin, int64_t count, void* X, int64_t ldx, int parts, cudaStream_t st) {
  if (count == 0) return LIST_OK;
  ListLayout lay;
  const int rc = list_feature_layout(ctx->map_channels, ctx->n_levels, ctx->vol_ch, &lay, nullptr);
  if (rc) return rc;
  const char* base = static_cast<const char*>(buf);
  AddParams a{};
  a.pmap = reinterpret_cast<const __nv_bfloat16*>(base + pl.off_pmap) + static_cast<size_t>(image) * ctx->map_size * ctx->map_size * kN0;
  if ((parts & kPartAddend) && pl.nh > kMaxH) {
    set_error("hoist::gather: the addend kernel covers at most %d hoisted levels (plan has %d)", kMaxH, pl.nh);
    return LIST_ENOSYS;
  }
  a.nh = pl.nh < kMaxH ? pl.nh : kMaxH;
  for (int h = 0; h < a.nh; ++h) {
    const size_t R = ctx->vol_res[pl.lev[h]];
    a.pvol[h] = reinterpret_cast<const __nv_bfloat16*>(base + pl.off_pvol[h]) + static_cast<size_t>(image) * R * R * R * kN0;
    a.dstride[h] = static_cast<uint32_t>(static_cast<size_t>(ctx->B) * R * R * R * kN0);
    a.R[h] = static_cast<int>(R);
  }
  for (int h = a.nh; h < kMaxH; ++h) { a.pvol[h] = a.pvol[0]; a.dstride[h] = 0; a.R[h] = 1; }
  a.T = ctx->trans_mat + image * 12;
  a.b0 = w->b0;
  a.X = static_cast<__nv_bfloat16*>(X);
  a.ldx = ldx;
  a.S = ctx->map_size;
  fill_tilemap(&a.tm, res, bb_min, bb_max, begin, count, kAddTile);

  RestParams r{};
  size_t smem = 0;
  const int rc2 = build_rest(ctx, pl, lay, image, res, &r, &smem);
  if (rc2) return rc2;
  r.X = a.X;
  r.ldx = ldx;
  r.ones = (parts & kPartOnes) ? 1 : 0;
  fill_tilemap(&r.tm, res, bb_min, bb_max, begin, count, r.tm.kPz);

  static const int vec = []() { const char* e = getenv("LIST_B200_HOIST_VEC"); return (e && e[0] == '8') ? 8 : 4; }();
  if (parts & kPartAddend) {
    LIST_CUDA(cudaFuncSetAttribute(hoist_addend_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    LIST_CUDA(cudaFuncSetAttribute(hoist_addend_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    if (vec == 8) hoist_addend_kernel<8><<<tile_count(a.tm), kN0 / 8, 0, st>>>(a);
    else hoist_addend_kernel<4><<<tile_count(a.tm), kN0 / 4, 0, st>>>(a);
    LIST_LAUNCH_CHECK("hoist_addend_kernel");
  }
  if (parts & kPartRest) {
    LIST_CUDA(cudaFuncSetAttribute(hoist_rest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    LIST_CUDA(cudaFuncSetAttribute(hoist_rest_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    hoist_rest_kernel<<<tile_count(r.tm), kRestThreads, smem, st>>>(r);
    LIST_LAUNCH_CHECK("hoist_rest_kernel");
  }
  return LIST_OK;
}

}  // namespace hoist
}  // namespace list
