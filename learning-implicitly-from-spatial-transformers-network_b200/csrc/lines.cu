// Per-line column tables of the hoisted voxel levels (bf16 dense-grid path, feeds grid_tc.cu).
//
// Along a z-line of the dense grid (reference utils.py:84-95: z fastest) the (H, D) position of every displaced copy of
// the query is constant, so the trilinear sample of a projected level (hoist.cu: W0[:, cols(level, d)] . volume, computed
// once per image) factors into
//        sample(p + disp_d) = w0(z) * G_d[i0(z)] + w1(z) * G_d[i0(z) + 1],     G_d[i] = sum_4 (wy*wz) PV_d[zc][yc][i][:]
// (reference modules.py:262-265: grid_sample, trilinear, border, align_corners).  The five displacements that do not move
// the W coordinate (d = 0, 3, 4, 5, 6; modules.py:205-212) share i0 / w0 / w1 along the line, so their G_d are summed:
// per line, level and W-shift class {0, -0.0722, +0.0722} there is ONE column of R rows x 512 channels.  This kernel
// writes those columns (fp32 accumulation, rounded once to bf16) as
//        G[line - line0][rowbase[h] + cls * R + i][512];
// grid_tc.cu then evaluates the z-interpolation of all of them -- and the bilinear sample of the projected feature map --
// as one small GEMM per tile on the tensor cores.
//
// A CTA produces the tables of LY consecutive lines of one x-plane (same D coordinate): their H coordinates lie within one
// voxel cell of the finest hoisted level, so per displacement the lines read the same 3 (H) x 2 (D) projected rows; those
// are loaded and reduced along D once, and every line takes its H-interpolation from the 3 reduced rows.  That cuts the
// L2 -> SM traffic of the naive per-line form (28 row reads per table row and line) by ~5x and leaves an FMA-bound kernel.
#include <cstdlib>

#include "grid_common.cuh"
#include "hoist.cuh"

namespace list {
namespace hoist {

constexpr int kLinesThreads = 256;
constexpr int kN0L = 512;
constexpr int kNY = 3;               // H nodes a CTA's lines can touch per displacement

struct LinesParams {
  const __nv_bfloat16* pvol[kMaxLev];   // image's slab of displacement 0: [R][R][R][512]
  uint32_t dstride[kMaxLev];            // elements between displacement slabs
  int R[kMaxLev], rowbase[kMaxLev];
  int nh, rpl, ly;                      // ly: lines per CTA (1, 2, 4 or 8)
  int groups;                           // CTAs per x-plane = ceil(res / ly)
  int64_t line_first, line_last;        // lines of the launch (inclusive)
  __nv_bfloat16* G;
  TileMap tm;
};

struct __align__(16) DispGeo {          // per (level, displacement) of a CTA
  uint32_t off0[kNY];                   // element offsets of the kNY rows (node 0, channel 0) in the first D plane, displacement slab included
  float wz0;
  uint32_t off1[kNY];                   // ... in the second D plane
  float wz1;
};

// 4 consecutive bf16 channels (8-byte access)
__device__ __forceinline__ void load4(const __nv_bfloat16* __restrict__ p, float v[4]) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
  v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
}
__device__ __forceinline__ void store4(__nv_bfloat16* __restrict__ p, const float v[4]) {
  uint2 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = u;
}

// Thread = (table row node i, 4-channel vector): 128 threads cover the 512 channels of a row, a CTA of 256 threads works
// on two nodes at a time.  Per displacement the thread reduces the 2 (D) x kNY (H) projected rows to kNY rows once and
// then takes every line's H-interpolation from them; when all lines of the CTA lie in one H cell (`two`: the third
// node has zero weight everywhere) the third row is neither loaded nor multiplied.
template <int LY>
__global__ void __launch_bounds__(kLinesThreads, 2) hoist_lines_kernel(const LinesParams p) {
  __shared__ DispGeo s_geo[kMaxLev][LIST_NUM_DISP];
  __shared__ __align__(16) float s_wy[kMaxLev][LIST_NUM_DISP][LY][4];        // H weights of every line on the kNY nodes; [3]: 1 if node 2 is unused
  const int tid = threadIdx.x;
  const int res = p.tm.res;
  const unsigned plane = blockIdx.x / static_cast<unsigned>(p.groups);
  const unsigned grp = blockIdx.x - plane * static_cast<unsigned>(p.groups);
  const int64_t lz = p.line_first / res + plane;
  const int ly0 = static_cast<int>(grp) * p.ly;
  // lines of this CTA: (lz, ly0 + j), j < ly, inside the launch's range
  if (tid < p.nh * LIST_NUM_DISP) {
    const int h = tid / LIST_NUM_DISP, d = tid - h * LIST_NUM_DISP;
    const int R = p.R[h];
    const float qz = linspace_f32_step(static_cast<int>(lz), res, p.tm.bb_min, p.tm.bb_max, p.tm.step) * 2.0f;
    float q[3] = {0.f, 0.f, qz}, pd[3];
    displaced(q, d, pd);
    const Axis3 az = axis_border(pd[2], R);
    DispGeo g;
    const uint32_t z0 = static_cast<uint32_t>(az.i0) * R * R * kN0L + static_cast<uint32_t>(d) * p.dstride[h];
    const uint32_t z1 = static_cast<uint32_t>(az.i1) * R * R * kN0L + static_cast<uint32_t>(d) * p.dstride[h];
    g.wz0 = az.w0; g.wz1 = az.w1;
    int ybase = 0;
    bool two = true;
    for (int j = 0; j < LY; ++j) {
      const int ly = min(ly0 + j, res - 1);
      q[1] = linspace_f32_step(ly, res, p.tm.bb_min, p.tm.bb_max, p.tm.step) * 2.0f;
      displaced(q, d, pd);
      const Axis3 ay = axis_border(pd[1], R);
      if (j == 0) ybase = ay.i0;
      // i0 - ybase is 0 or 1 (the lines of a CTA span less than one cell); i1 == i0 only at the border, where w1 == 0
      const int r0 = ay.i0 - ybase, r1 = r0 + (ay.i1 - ay.i0);
      float w[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) w[k] = (k == r0 ? ay.w0 : 0.f) + (k == r1 ? ay.w1 : 0.f);
      two = two && w[2] == 0.f;
      s_wy[h][d][j][0] = w[0]; s_wy[h][d][j][1] = w[1]; s_wy[h][d][j][2] = w[2];
    }
    for (int j = 0; j < LY; ++j) s_wy[h][d][j][3] = two ? 1.f : 0.f;
#pragma unroll
    for (int k = 0; k < kNY; ++k) {
      const uint32_t yo = static_cast<uint32_t>(min(ybase + k, R - 1)) * R * kN0L;
      g.off0[k] = z0 + yo;
      g.off1[k] = z1 + yo;
    }
    s_geo[h][d] = g;
  }
  __syncthreads();
  // valid lines of the CTA
  int jlo = 0, jhi = LY;
  {
    const int64_t l0 = lz * res + ly0;
    if (l0 < p.line_first) jlo = static_cast<int>(p.line_first - l0);
    const int64_t lend = min(static_cast<int64_t>(res) - ly0, p.line_last + 1 - l0);
    if (lend < jhi) jhi = static_cast<int>(lend);
    if (p.ly < jhi) jhi = p.ly;
  }
  if (jlo >= jhi) return;
  const int64_t out_line0 = lz * res + ly0 - p.line_first;       // may be negative for the lines below jlo
  const int line_stride = p.rpl * kN0L;                           // elements between the tables of consecutive lines
  const int v = tid & 127, ig = tid >> 7;                        // 4-channel vector, node group
  __nv_bfloat16* const gline = p.G + out_line0 * line_stride + v * 4;   // (before the buffer for the lines below jlo: never stored)
  // Steps of a thread: (node i, displacement in class order 0 3 4 5 6 | 1 | 2).  The six row loads of step n + 1 are issued
  // before step n's arithmetic (two register sets, the loop is unrolled by two), so the loads' latency -- the kernel's
  // dominant stall at 16 warps per SM -- hides behind ~100 FMAs; a class's rows are stored when its last step is done.
  for (int h = 0; h < p.nh; ++h) {
    const int R = p.R[h];
    const __nv_bfloat16* __restrict__ pv = p.pvol[h] + v * 4;
    if (ig >= R) continue;
    float2 acc[LY][2];
    auto issue = [&](int i, int sq, uint2 (&r0)[kNY], uint2 (&r1)[kNY]) {
      const int d = sq == 0 ? 0 : (sq < 5 ? sq + 2 : sq - 4);
      const uint4 o0 = *reinterpret_cast<const uint4*>(s_geo[h][d].off0), o1 = *reinterpret_cast<const uint4*>(s_geo[h][d].off1);
      const __nv_bfloat16* __restrict__ pd = pv + static_cast<uint32_t>(i) * kN0L;
      const bool two = s_wy[h][d][0][3] != 0.f;                   // uniform over the CTA
      r0[0] = __ldg(reinterpret_cast<const uint2*>(pd + o0.x));
      r1[0] = __ldg(reinterpret_cast<const uint2*>(pd + o1.x));
      r0[1] = __ldg(reinterpret_cast<const uint2*>(pd + o0.y));
      r1[1] = __ldg(reinterpret_cast<const uint2*>(pd + o1.y));
      if (!two) {
        r0[2] = __ldg(reinterpret_cast<const uint2*>(pd + o0.z));
        r1[2] = __ldg(reinterpret_cast<const uint2*>(pd + o1.z));
      }
    };
    auto work = [&](int i, int sq, const uint2 (&r0)[kNY], const uint2 (&r1)[kNY]) {
      const int d = sq == 0 ? 0 : (sq < 5 ? sq + 2 : sq - 4);
      const DispGeo& g = s_geo[h][d];
      const bool two = s_wy[h][d][0][3] != 0.f;
      if (sq == 0 || sq >= 5) {
#pragma unroll
        for (int j = 0; j < LY; ++j) { acc[j][0] = make_float2(0.f, 0.f); acc[j][1] = make_float2(0.f, 0.f); }
      }
      float2 u[kNY][2];
      const float2 wz0 = make_float2(g.wz0, g.wz0), wz1 = make_float2(g.wz1, g.wz1);
#pragma unroll
      for (int k = 0; k < kNY; ++k) {
        if (k < 2 || !two) {
          const float2 a0 = bf16x2_to_f2(r0[k].x), a1 = bf16x2_to_f2(r0[k].y), b0 = bf16x2_to_f2(r1[k].x), b1 = bf16x2_to_f2(r1[k].y);
          u[k][0] = ffma2(b0, wz1, make_float2(a0.x * g.wz0, a0.y * g.wz0));
          u[k][1] = ffma2(b1, wz1, make_float2(a1.x * g.wz0, a1.y * g.wz0));
        } else {
          u[k][0] = make_float2(0.f, 0.f); u[k][1] = make_float2(0.f, 0.f);
        }
      }
      if (two) {
#pragma unroll
        for (int j = 0; j < LY; ++j) {
          const float2 w = *reinterpret_cast<const float2*>(s_wy[h][d][j]);
          const float2 wx = make_float2(w.x, w.x), wy = make_float2(w.y, w.y);
#pragma unroll
          for (int c = 0; c < 2; ++c) acc[j][c] = ffma2(u[1][c], wy, ffma2(u[0][c], wx, acc[j][c]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < LY; ++j) {
          const float4 w = *reinterpret_cast<const float4*>(s_wy[h][d][j]);
          const float2 wx = make_float2(w.x, w.x), wy = make_float2(w.y, w.y), wz = make_float2(w.z, w.z);
#pragma unroll
          for (int c = 0; c < 2; ++c) acc[j][c] = ffma2(u[2][c], wz, ffma2(u[1][c], wy, ffma2(u[0][c], wx, acc[j][c])));
        }
      }
      if (sq >= 4) {                                                // last displacement of its class
        const int cls = sq - 4;
        __nv_bfloat16* const grow = gline + (p.rowbase[h] + cls * R + i) * kN0L;   // this table row in line j = 0
#pragma unroll
        for (int j = 0; j < LY; ++j)
          if (j >= jlo && j < jhi) {
            uint2 o;
            o.x = pack_bf16x2(acc[j][0].x, acc[j][0].y);
            o.y = pack_bf16x2(acc[j][1].x, acc[j][1].y);
            *reinterpret_cast<uint2*>(grow + j * line_stride) = o;
          }
      }
    };
    uint2 a0[kNY], a1[kNY], b0[kNY], b1[kNY];
    int i = ig, sq = 0;
    issue(i, sq, a0, a1);
#pragma unroll 1
    while (i < R) {
      int ni = i, ns = sq + 1;
      if (ns == LIST_NUM_DISP) { ns = 0; ni = i + kLinesThreads / 128; }
      if (ni < R) issue(ni, ns, b0, b1);
      work(i, sq, a0, a1);
      i = ni; sq = ns;
      if (i >= R) break;
      ni = i; ns = sq + 1;
      if (ns == LIST_NUM_DISP) { ns = 0; ni = i + kLinesThreads / 128; }
      if (ni < R) issue(ni, ns, a0, a1);
      work(i, sq, b0, b1);
      i = ni; sq = ns;
    }
  }
}

size_t lines_bytes(const Plan& pl, int res, int64_t begin, int64_t count) {
  if (count <= 0 || pl.rpl == 0) return 0;
  const int64_t nlines = (begin + count - 1) / res - begin / res + 1;
  return static_cast<size_t>(nlines) * pl.rpl * kN0L * 2;
}

int lines_tc(const ListCtx* ctx, const Plan& pl, const void* buf, int image, int res, double bb_min, double bb_max, int64_t begin,
             int64_t count, void* G, cudaStream_t st);   // lines_tc.cu

int lines(const ListCtx* ctx, const Plan& pl, const void* buf, int image, int res, double bb_min, double bb_max, int64_t begin,
          int64_t count, void* G, cudaStream_t st) {
  if (count == 0 || pl.nh == 0) return LIST_OK;
  // LIST_B200_LINES_TC=1: the tensor-core formulation (lines_tc.cu).  Correct, but measured slower than this file's SIMT
  // kernel on B200 (1.38 vs 1.35 ms per 4 M rows at 256^3, 1.5 vs 0.86 ms at 128^3: its pipeline skeleton alone takes
  // 0.7 ms), so it is opt-in; DESIGN.md section 4.4.
  const char* e = getenv("LIST_B200_LINES_TC");
  if (e && e[0] == '1') {
    const int rc = lines_tc(ctx, pl, buf, image, res, bb_min, bb_max, begin, count, G, st);
    if (rc != LIST_ENOSYS) return rc;
  }
  const char* base = static_cast<const char*>(buf);
  LinesParams p{};
  p.nh = pl.nh;
  p.rpl = pl.rpl;
  int rmax = 2;
  for (int h = 0; h < kMaxLev; ++h) {
    const int hh = h < pl.nh ? h : 0;
    const size_t R = ctx->vol_res[pl.lev[hh]];
    p.pvol[h] = reinterpret_cast<const __nv_bfloat16*>(base + pl.off_pvol[hh]) + static_cast<size_t>(image) * R * R * R * kN0L;
    p.dstride[h] = static_cast<uint32_t>(static_cast<size_t>(ctx->B) * R * R * R * kN0L);
    p.R[h] = static_cast<int>(R);
    p.rowbase[h] = pl.rowbase[hh];
    if (h < pl.nh && static_cast<int>(R) > rmax) rmax = static_cast<int>(R);
  }
  p.G = static_cast<__nv_bfloat16*>(G);
  fill_tilemap(&p.tm, res, bb_min, bb_max, begin, count, 128);
  p.line_first = begin / res;
  p.line_last = (begin + count - 1) / res;
  // lines per CTA: as many as fit in one H cell of the finest hoisted level (so that they touch at most kNY nodes)
  int ly = res > 1 ? (res - 1) / (rmax - 1) : 1;
  ly = ly >= 8 ? 8 : (ly >= 4 ? 4 : (ly >= 2 ? 2 : 1));
  p.ly = ly;
  p.groups = (res + ly - 1) / ly;
  const int64_t planes = p.line_last / res - p.line_first / res + 1;
  const int64_t ctas = planes * p.groups;
  LIST_CHECK_ARG(ctas < (1LL << 31), "hoist::lines: too many lines for one launch");
  const unsigned grid = static_cast<unsigned>(ctas);
  if (ly == 8) hoist_lines_kernel<8><<<grid, kLinesThreads, 0, st>>>(p);
  else if (ly == 4) hoist_lines_kernel<4><<<grid, kLinesThreads, 0, st>>>(p);
  else if (ly == 2) hoist_lines_kernel<2><<<grid, kLinesThreads, 0, st>>>(p);
  else hoist_lines_kernel<1><<<grid, kLinesThreads, 0, st>>>(p);
  LIST_LAUNCH_CHECK("hoist_lines_kernel");
  return LIST_OK;
}

}  // namespace hoist
}  // namespace list
