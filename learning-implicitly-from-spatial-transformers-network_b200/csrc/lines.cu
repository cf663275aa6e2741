// Per-line column tables of the hoisted voxel levels (bf16 dense-grid path, feeds grid_tc.cu).
//
// Along a z-line of the dense grid (reference utils.py:84-95: z fastest) the (H, D) position of every displaced copy of
// the query is constant, so the trilinear sample of a projected level (hoist.cu: W0[:, cols(level, d)] . volume, computed
// once per image) factors into
//        sample(p + disp_d) = w0(z) * G_d[i0(z)] + w1(z) * G_d[i0(z) + 1],     G_d[i] = sum_4 (wy*wz) PV_d[zc][yc][i][:]
// (reference modules.py:262-265: grid_sample, trilinear, border, align_corners).  The five displacements that do not move
// the W coordinate (d = 0, 3, 4, 5, 6; modules.py:205-212) share i0 / w0 / w1 along the line, so their G_d are summed:
// per line, level and W-shift class {0, -0.0722, +0.0722} there is ONE column of R rows x 512 channels.  This kernel
// writes those columns (fp32 accumulation in the tap order of gather_grid.cu, rounded once to bf16) as
//        G[line - line0][rowbase[h] + cls * R + i][512];
// grid_tc.cu then evaluates the z-interpolation of all of them -- and the bilinear sample of the projected feature map --
// as one small GEMM per tile on the tensor cores.
#include "grid_common.cuh"
#include "hoist.cuh"

namespace list {
namespace hoist {

constexpr int kLinesThreads = 256;
constexpr int kN0L = 512;

struct LinesParams {
  const __nv_bfloat16* pvol[kMaxLev];   // image's slab of displacement 0: [R][R][R][512]
  uint32_t dstride[kMaxLev];            // elements between displacement slabs
  int R[kMaxLev], rowbase[kMaxLev];
  int nh, rpl;
  __nv_bfloat16* G;
  TileMap tm;
};

template <int ND>
__device__ __forceinline__ void line_column(const __nv_bfloat16* __restrict__ pv, uint32_t dstride, const Corner* __restrict__ cor,
                                            const int (&dl)[ND], uint32_t off, float acc[8]) {
  float v[ND * 4][8];
#pragma unroll
  for (int di = 0; di < ND; ++di)
#pragma unroll
    for (int k = 0; k < 4; ++k) load8(pv + static_cast<size_t>(dl[di]) * dstride + cor[dl[di] * 4 + k].base + off, v[di * 4 + k]);
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
  for (int di = 0; di < ND; ++di)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float w = cor[dl[di] * 4 + k].w;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(v[di * 4 + k][j], w, acc[j]);
    }
}

__global__ void __launch_bounds__(kLinesThreads) hoist_lines_kernel(const LinesParams p) {
  __shared__ Corner s_cor[kMaxLev][LIST_NUM_DISP * 4];
  const int tid = threadIdx.x;
  const unsigned line = static_cast<unsigned>(p.tm.line0) + blockIdx.x;
  const unsigned lz = line / static_cast<unsigned>(p.tm.res), ly = line - lz * static_cast<unsigned>(p.tm.res);
  const float qy = linspace_f32_step(static_cast<int>(ly), p.tm.res, p.tm.bb_min, p.tm.bb_max, p.tm.step) * 2.0f;
  const float qz = linspace_f32_step(static_cast<int>(lz), p.tm.res, p.tm.bb_min, p.tm.bb_max, p.tm.step) * 2.0f;
  if (tid < p.nh * LIST_NUM_DISP) {
    const int h = tid / LIST_NUM_DISP, d = tid - h * LIST_NUM_DISP;
    uint32_t base[4];
    float wyz[4];
    tile_corners(qy, qz, d, p.R[h], kN0L, base, wyz);
#pragma unroll
    for (int k = 0; k < 4; ++k) s_cor[h][d * 4 + k] = Corner{base[k], wyz[k]};
  }
  __syncthreads();
  const int v = tid & 63, rg = tid >> 6;                       // 8-channel vector, row group
  __nv_bfloat16* __restrict__ out = p.G + static_cast<size_t>(blockIdx.x) * p.rpl * kN0L + v * 8;
  for (int row = rg; row < p.rpl; row += kLinesThreads / 64) {
    int h = 0;
#pragma unroll
    for (int i = 1; i < kMaxLev; ++i)
      if (i < p.nh && row >= p.rowbase[i]) h = i;
    const int rel = row - p.rowbase[h];
    const int R = p.R[h];
    const int cls = rel / R, node = rel - cls * R;
    const uint32_t off = static_cast<uint32_t>(node) * kN0L + v * 8;
    float acc[8];
    if (cls == 0) {
      const int dl[5] = {0, 3, 4, 5, 6};
      line_column<5>(p.pvol[h], p.dstride[h], s_cor[h], dl, off, acc);
    } else {
      const int dl[1] = {cls};
      line_column<1>(p.pvol[h], p.dstride[h], s_cor[h], dl, off, acc);
    }
    store8(out + static_cast<size_t>(row) * kN0L, acc);
  }
}

size_t lines_bytes(const Plan& pl, int res, int64_t begin, int64_t count) {
  if (count <= 0 || pl.rpl == 0) return 0;
  const int64_t nlines = (begin + count - 1) / res - begin / res + 1;
  return static_cast<size_t>(nlines) * pl.rpl * kN0L * 2;
}

int lines(const ListCtx* ctx, const Plan& pl, const void* buf, int image, int res, double bb_min, double bb_max, int64_t begin,
          int64_t count, void* G, cudaStream_t st) {
  if (count == 0 || pl.nh == 0) return LIST_OK;
  const char* base = static_cast<const char*>(buf);
  LinesParams p{};
  p.nh = pl.nh;
  p.rpl = pl.rpl;
  for (int h = 0; h < kMaxLev; ++h) {
    const int hh = h < pl.nh ? h : 0;
    const size_t R = ctx->vol_res[pl.lev[hh]];
    p.pvol[h] = reinterpret_cast<const __nv_bfloat16*>(base + pl.off_pvol[hh]) + static_cast<size_t>(image) * R * R * R * kN0L;
    p.dstride[h] = static_cast<uint32_t>(static_cast<size_t>(ctx->B) * R * R * R * kN0L);
    p.R[h] = static_cast<int>(R);
    p.rowbase[h] = pl.rowbase[hh];
  }
  p.G = static_cast<__nv_bfloat16*>(G);
  fill_tilemap(&p.tm, res, bb_min, bb_max, begin, count, 128);
  const int64_t nlines = line_count(p.tm);
  hoist_lines_kernel<<<static_cast<unsigned>(nlines), kLinesThreads, 0, st>>>(p);
  LIST_LAUNCH_CHECK("hoist_lines_kernel");
  return LIST_OK;
}

}  // namespace hoist
}  // namespace list
