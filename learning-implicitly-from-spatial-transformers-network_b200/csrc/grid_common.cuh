// Geometry shared by the dense-grid kernels (hoist.cu, lines.cu, grid_tc.cu): the tile <-> grid-point mapping of a
// launch and the (H, D) corners of a z-line.  Reference: utils.py:84-95 (grid order: x slowest, z fastest),
// models.py:91-92 ([2,1,0] swap, *2), modules.py:205-212 (displacement table), :264-265 (trilinear, border).
#pragma once
#include "common.cuh"

namespace list {
namespace hoist {

// Tiles never straddle z-runs: tile t = (z-line, segment of kPz steps of that line), clipped to the launch's
// point range [begin, end).  Everything a tile computes is therefore a function of absolute grid positions
// only, which is what makes any chunking / sharding of the grid bit-identical.
struct TileMap {
  int64_t line0;                    // first z-line (flat index / res) touched by [begin, end); < res^2 <= 2^22
  int64_t begin, end;               // flat grid range of this launch
  int segs;                         // tiles per z-line
  int kPz, lg_kpz;                  // steps per tile (a power of two)
  int res;
  double bb_min, bb_max, step;      // step = (bb_max - bb_min) / (res - 1), the linspace increment
};

struct TileSpan {
  int64_t g_tile0;                  // flat grid index of step 0
  int gz0;                          // its position on the z-line
  int s_lo, s_hi;                   // steps of the tile inside [begin, end)
  float qy, qz;                     // swapped/scaled query components 1 (-> H) and 2 (-> D), constant over the tile
  unsigned line_rel;                // z-line of the tile relative to TileMap::line0
};

// 32-bit index arithmetic only (64-bit divisions cost ~100 instructions each and every thread of a tile runs this)
__device__ __forceinline__ bool tile_span(const TileMap& m, unsigned tile, TileSpan& t) {
  const unsigned lrel = tile / static_cast<unsigned>(m.segs);
  const unsigned seg = tile - lrel * static_cast<unsigned>(m.segs);
  const unsigned line = static_cast<unsigned>(m.line0) + lrel;
  const unsigned lz = line / static_cast<unsigned>(m.res), ly = line - lz * static_cast<unsigned>(m.res);
  t.line_rel = lrel;
  t.gz0 = static_cast<int>(seg) << m.lg_kpz;
  t.g_tile0 = static_cast<int64_t>(line) * m.res + t.gz0;
  const int full = min(m.kPz, m.res - t.gz0);
  t.s_lo = static_cast<int>(max(static_cast<int64_t>(0), m.begin - t.g_tile0));
  t.s_hi = static_cast<int>(min(static_cast<int64_t>(full), m.end - t.g_tile0));
  // reference utils.py:84-95 (x slowest, z fastest) and models.py:91-92 ([2,1,0] swap, *2)
  t.qy = linspace_f32_step(static_cast<int>(ly), m.res, m.bb_min, m.bb_max, m.step) * 2.0f;
  t.qz = linspace_f32_step(static_cast<int>(lz), m.res, m.bb_min, m.bb_max, m.step) * 2.0f;
  return t.s_lo < t.s_hi;
}
__device__ __forceinline__ float step_q0(const TileMap& m, const TileSpan& t, int s) {
  return linspace_f32_step(t.gz0 + s, m.res, m.bb_min, m.bb_max, m.step) * 2.0f;
}

inline void fill_tilemap(TileMap* tm, int res, double bb_min, double bb_max, int64_t begin, int64_t count, int kpz) {
  tm->line0 = begin / res;
  tm->begin = begin;
  tm->end = begin + count;
  tm->kPz = kpz;
  tm->lg_kpz = 0;
  while ((1 << tm->lg_kpz) < kpz) ++tm->lg_kpz;
  tm->segs = (res + kpz - 1) / kpz;
  tm->res = res;
  tm->bb_min = bb_min;
  tm->bb_max = bb_max;
  tm->step = res > 1 ? (bb_max - bb_min) / static_cast<double>(res - 1) : 0.0;
}
inline int64_t line_count(const TileMap& tm) { return (tm.end - 1) / tm.res - tm.line0 + 1; }
inline unsigned tile_count(const TileMap& tm) { return static_cast<unsigned>(line_count(tm) * tm.segs); }

struct Corner { uint32_t base; float w; };

// W-shift class of a displacement: the five displacements {0,3,4,5,6} share voxel index and weight along the walk
__device__ __forceinline__ int shift_class(int d) { return d == 1 ? 1 : (d == 2 ? 2 : 0); }
__device__ __forceinline__ float class_shift(int cls) { return cls == 0 ? 0.f : (cls == 1 ? -kDisplacement : kDisplacement); }

// (H, D) corners and weights of displacement d for a z-line (same arithmetic as gather_grid.cu)
__device__ __forceinline__ void tile_corners(float qy, float qz, int d, int R, uint32_t row_elems, uint32_t base[4], float wyz[4]) {
  const float q[3] = {0.f, qy, qz};
  float pd[3];
  displaced(q, d, pd);
  const Axis3 ay = axis_border(pd[1], R), az = axis_border(pd[2], R);
  const int zi[2] = {az.i0, az.i1}, yi[2] = {ay.i0, ay.i1};
  const float wz[2] = {az.w0, az.w1}, wy[2] = {ay.w0, ay.w1};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int tz = k >> 1, ty = k & 1;
    base[k] = (static_cast<uint32_t>(zi[tz]) * R + yi[ty]) * R * row_elems;
    wyz[k] = wy[ty] * wz[tz];
  }
}

}  // namespace hoist
}  // namespace list
