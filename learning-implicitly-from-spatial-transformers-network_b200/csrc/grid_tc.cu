// Dense-grid SDF evaluation, bf16 tensor-core mode: interpolation of the hoisted fc_0 terms AND the implicit MLP in ONE
// persistent warp-specialised sm_100a kernel (rows a-3, a-5, a-6, a-8 of SURVEY.md §8; reference network/modules.py:48-53,
// 262-282 inside the chunk loop of network/executors.py:215-224).
//
// fc_0 is linear and so are the samplers in front of it, so the contribution of the perceptual maps and of the coarse
// voxel levels to fc_0's pre-activation is  sample(W0[:, cols] . tensor, p)  (hoist.cu projects the tensors once per image,
// lines.cu reduces the projected levels to one column table per z-line).  What is left per query is a SPARSE linear
// combination of 512-wide bf16 rows: 4 bilinear taps of the projected map P and, per hoisted level and W-shift class, the
// two neighbours of the line's column table G.  A tile of 128 consecutive steps of a z-line touches only ~120 distinct
// rows, so the combination is evaluated on the tensor cores, accumulating into the same TMEM tile as fc_0's dense part:
//
//     acc[128 x 512]  =  Xr[128 x K_F] . W0[:, hoisted..]^T            (F chunks: TMA operands, as in mlp_tc.cu)
//                      +  Aw[128 x rows] . Brows[rows x 512]           (I chunks: Aw = interpolation weights written by the
//                                                                       interp warps, Brows = rows of P / G copied with
//                                                                       cp.async; MN-major B operand)
// (fc_0's bias included: Xr carries 1.0 in three pad columns, the weight copy of hoist::prepare the bf16 hi / mid / lo
// parts of b0) followed by ReLU, fc_1, fc_2, fc_out as in mlp_tc.cu (activations stay in TMEM).  The 512-wide "addend"
// block of round 1 (34 GB through HBM per 256^3 grid) no longer exists.
//
//   Warp roles : warp 0 = TMA producer and ring allocator, warp 1 = TMEM allocator + MMA issuer (one thread),
//                warps 2..9 = epilogue (two per TMEM lane quarter, half of the columns each), warps 10..13 = interp warps
//                (Aw, Brows).
//   CG = 2     : CTA pair, tcgen05.mma.cta_group::2, M = 256.  CTA r evaluates tile 2p + r; the K space of the pair's
//                I chunks is the concatenation of both tiles' row lists (the other CTA's rows get zero weights), and each
//                CTA copies ITS half of the 512 channels of every listed row (B operand split along N).
//   Ring       : 12 units of 16 KB, allocated first-in first-out by the TMA thread in consumption order
//                  I chunk [Aw | Brows j=0 | Brows j=1] -> F chunk [X box | W0 box | W0 box] -> 8 + 4 weight boxes of fc_1 / fc_2.
//                I chunks are filled by the interp warps once the allocator has granted their units (grant barriers).  The
//                first two of each tile of the pair come first in the sequence, so that they are granted and filled during
//                the previous pair's fc_1 / fc_2 (the ring holds four such chunks); the others follow the F chunks and are
//                filled under them.  A tile's own chunks thus always sit at the same places relative to its F chunks,
//                whichever CTA of a pair evaluates it and whatever its partner needs (accumulation order is fixed).
//   Tile plans : grid_plan_kernel (one CTA per tile, launched before this kernel) writes per tile the list of source rows
//                (absolute addresses; voxel rows = contiguous node ranges of the line table per (level, class); pixel rows
//                = the nodes of the pixel cells the tile's path crosses, de-duplicated between consecutive cells) and per
//                step its <= 22 (chunk, position, bf16 weight) entries.  Planning is serial, latency-bound index work; as
//                its own kernel it runs at full occupancy instead of on four warps next to the MMA pipeline.
#include <cstdlib>

#include "grid_common.cuh"
#include "hoist.cuh"
#include "tc_common.cuh"

namespace list {
namespace gtc {

using namespace tc;
using hoist::TileMap;
using hoist::TileSpan;

constexpr int CG = 2;
constexpr int BM = 128, BK = 64;
constexpr int N0 = 512, N1 = 256, N2 = 256;
constexpr int UNIT_BYTES = 128 * BK * 2;           // 16 KB
constexpr int NU = 12, NB = 12;                    // ring units / chunk barriers
constexpr int U_FI = 3;                            // units of an F or I chunk
constexpr int N_W = (N0 + N1) / BK;                // fc_1 + fc_2 weight chunks per tile, one unit each
constexpr int kPre = 2;                            // I chunks whose first column half is issued under the previous tile's last epilogue
constexpr int kIFirst = 2;                         // I chunks of EACH tile of a pair that precede the F chunks in the ring order (the rest follow them)
constexpr int kEpiWarp0 = 2, kEpiWarps = 8, kIntWarp0 = kEpiWarp0 + kEpiWarps, kIntWarps = 4;
constexpr int kThreads = (kIntWarp0 + kIntWarps) * 32;   // 448
constexpr int kMaxLev = hoist::kMaxLev;
constexpr int NC = kMaxLev * 3;                    // (level, W-shift class) combinations
constexpr int kEnt = 2 * NC + 4;                   // weight entries of a step: two per combination + 4 pixel taps
constexpr int kEntPad = 24;                        // row pitch of the entry table (six 16-byte vectors)
static_assert(kEnt <= kEntPad, "entry table row");
constexpr int kMaxRows = 832;                      // rows of one tile's list: <= 3 * sum R voxel + 4 * 128 pixel, 64-aligned
constexpr uint32_t kEmpty = 0xff800000u;           // weight entry that matches no chunk
// weight entry: chunk << 23 | byte offset inside the row's 128 B of the (128B-swizzled, K-major) Aw unit << 16 | bf16 weight

// ---- tile plans in global memory (grid_plan_kernel -> grid_tc_kernel) ----
constexpr int kPlanRowBytes = kMaxRows * 8;        // uint64 source address per listed row
constexpr int kPlanEntBytes = BM * kEntPad * 4;    // uint32 entries [128 steps][kEntPad]
struct PlanBuf {
  int* hdr;                         // [n_tiles] 64-row chunks of the tile's list | 16-row k-steps used of the last chunk << 8
                                    // (0: tile outside the launch's range)
  unsigned long long* rows;         // [n_tiles][kMaxRows]
  uint32_t* ent;                    // [n_tiles][128][kEntPad]
};
inline size_t plan_hdr_bytes(unsigned n_tiles) { return (static_cast<size_t>(n_tiles) * 4 + 255) / 256 * 256; }
inline size_t plan_bytes(unsigned n_tiles) {
  return plan_hdr_bytes(n_tiles) + static_cast<size_t>(n_tiles) * (kPlanRowBytes + kPlanEntBytes);
}
inline PlanBuf plan_carve(void* buf, unsigned n_tiles) {
  char* b = static_cast<char*>(buf);
  PlanBuf pb;
  pb.hdr = reinterpret_cast<int*>(b);
  pb.rows = reinterpret_cast<unsigned long long*>(b + plan_hdr_bytes(n_tiles));
  pb.ent = reinterpret_cast<uint32_t*>(b + plan_hdr_bytes(n_tiles) + static_cast<size_t>(n_tiles) * kPlanRowBytes);
  return pb;
}

// ---- shared-memory carve-up of grid_tc_kernel (offsets from the 1024-byte aligned base) ----
constexpr int OFF_PAR = NU * UNIT_BYTES;                           // b1 b2 w3 (fp32; b0 rides in the MMA)
constexpr int PARAM_FLOATS = N1 + N2 + N2;
constexpr int OFF_ENT = OFF_PAR + PARAM_FLOATS * 4;                // this CTA's tile: uint32 [128][kEntPad]
constexpr int OFF_ROWS = OFF_ENT + kPlanEntBytes;                  // uint64 [2][kMaxRows]: both tiles' row lists
constexpr int OFF_PART = OFF_ROWS + 2 * kPlanRowBytes;               // float [128]: fc_out partial sums of the upper column halves
constexpr int OFF_BAR = OFF_PART + BM * 4;
constexpr int NUM_BARS = 4 * NB + 2;                               // full empty grant ifull | dfull hready
constexpr int SMEM_BYTES = OFF_BAR + NUM_BARS * 8 + 16 + 1024 /*align slack*/;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

// geometry of the interpolated terms, shared by the plan kernel and the fused kernel
struct Geo {
  TileMap tm;
  unsigned n_tiles;
  const __nv_bfloat16* pmap;        // projected map of the image [S*S][512]
  const __nv_bfloat16* G;           // line tables [lines][rpl][512]
  const __nv_bfloat16* zero;        // 512 zeros
  const float* T;
  int S, nh, rpl;
  int R[kMaxLev], rowbase[kMaxLev];
};

struct Params {
  const float *b1, *b2, *w3, *b3;
  float* sdf;                       // [count] of this image
  float out_div;
  int nkF;                          // F chunks: (k_pad - hoist_cols) / 64
  TileMap tm;
  unsigned n_tiles;
  PlanBuf plan;
  float* dbg1;                      // optional [count][512]: relu(fc_0) in fp32 before rounding
  long long* trace;
  int trace_stride;                 // every trace_stride-th tile pair of CTA 0 is recorded
  unsigned long long* stats;        // optional [2]: += pair tiles, += I chunks (executed-FLOP accounting of bench.py)
  uint32_t b_lbo, b_sbo, b_kadv;    // MN-major descriptor of the Brows units: bytes between 64-channel groups / 8-row groups / 16 k rows
};

__device__ __forceinline__ void tile_rows(const TileMap& m, unsigned tile, unsigned n_tiles, int64_t& g0, int& s_lo, int& s_hi) {
  if (tile >= n_tiles) { g0 = m.end; s_lo = 0; s_hi = 0; return; }
  const unsigned lrel = tile / static_cast<unsigned>(m.segs);
  const unsigned seg = tile - lrel * static_cast<unsigned>(m.segs);
  const int gz0 = static_cast<int>(seg) << m.lg_kpz;
  g0 = (m.line0 + lrel) * m.res + gz0;
  const int full = min(m.kPz, m.res - gz0);
  s_lo = static_cast<int>(max(static_cast<int64_t>(0), m.begin - g0));
  s_hi = static_cast<int>(min(static_cast<int64_t>(full), m.end - g0));
  if (s_hi < s_lo) s_hi = s_lo;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_prior1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void st_shared_zero16(uint32_t addr) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
}
__device__ __forceinline__ void st_shared_u16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(static_cast<unsigned short>(v)) : "memory");
}
// Remote arrive in the form CUTLASS uses for its cross-CTA pipeline barriers (default semantics).  The data hand-offs it
// signals are ordered by the proxy / tcgen05 fences executed before it; the explicit .release.cluster form compiles to
// MEMBAR.ALL.GPU, ~1k cycles on every hand-off.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ uint32_t bf16_bits(float x) { return static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16_rn(x))); }

// ------------------------------------------------------------------ tile plans
// One CTA of 128 threads per tile, thread = step.  Writes hdr[tile] (64-row chunks), rows[tile][..] (source address of
// every listed row, padded with the zero row to a whole chunk) and ent[tile][step][..] (weight entries).
__global__ void __launch_bounds__(BM, 12) grid_plan_kernel(const Geo p, const PlanBuf out) {
  __shared__ int s_key[BM], s_ckey[BM], s_cbase[BM], s_cmask[BM];
  __shared__ short s_cslot[BM * 4];
  __shared__ int s_first[NC], s_last[NC], s_scan[8];
  const unsigned tile = blockIdx.x;
  const int it = threadIdx.x;
  uint32_t* const ent = out.ent + (static_cast<size_t>(tile) * BM + it) * kEntPad;
  unsigned long long* const rows = out.rows + static_cast<size_t>(tile) * kMaxRows;
  TileSpan t;
  if (!tile_span(p.tm, tile, t)) {                                // uniform: tile outside [begin, end)
    if (it == 0) out.hdr[tile] = 0;
    return;
  }
  // The plan covers ALL steps of the tile that lie on its z-line, also those outside [begin, end): row list, slot order and
  // with them the order in which the tensor cores accumulate are then a function of the absolute tile only, which keeps any
  // chunking / sharding of the grid bit-identical (the steps outside the range are computed and not stored).
  t.s_lo = 0;
  t.s_hi = min(p.tm.kPz, p.tm.res - t.gz0);
  const int lane = it & 31, w = it >> 5;
  const int s = it;
  const bool valid = s < t.s_hi;
  const int S = p.S;
  const float q0 = valid ? step_q0(p.tm, t, s) : 0.f;
  // ---- pixel cell and bilinear weights (reference modules.py:37-52) ----
  int key = -1, x0 = 0, y0 = 0;
  float pw[4] = {0.f, 0.f, 0.f, 0.f};
  if (valid) {
    const float q[3] = {q0, t.qy, t.qz};
    float ix, iy, h[3];
    localise(q, p.T, S, ix, iy, h);
    if (ix == ix && iy == iy) {                                   // NaN grid -> all taps out of bounds
      const float fx = floorf(ix), fy = floorf(iy);
      x0 = static_cast<int>(fx); y0 = static_cast<int>(fy);
      const float wx1 = ix - fx, wx0 = (fx + 1.0f) - ix;
      const float wy1 = iy - fy, wy0 = (fy + 1.0f) - iy;
      pw[0] = wx0 * wy0; pw[1] = wx1 * wy0; pw[2] = wx0 * wy1; pw[3] = wx1 * wy1;
      key = y0 * S + x0;
    }
  }
  s_key[s] = key;
  // ---- voxel index / weights per (level, class) (reference modules.py:262-265) ----
  int vi0[NC];
  uint32_t vw[NC];                                                // bf16 w0 | bf16 w1 << 16, i1 == i0 flagged by w1 == 0xffff
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    vi0[c] = 0; vw[c] = 0;
    if (c / 3 < p.nh) {
      const int R = p.R[c / 3];
      const Axis3 ax = axis_border(q0 + hoist::class_shift(c % 3), R);
      vi0[c] = ax.i0;
      // the larger weight is rounded to bf16, the smaller one is its exact complement (representable): the pair sums to 1
      const bool big0 = ax.w0 >= ax.w1;
      const float a = __bfloat162float(__float2bfloat16_rn(big0 ? ax.w0 : ax.w1));
      const float b = 1.0f - a;
      const uint32_t w0b = bf16_bits(big0 ? a : b), w1b = bf16_bits(big0 ? b : a);
      vw[c] = w0b | ((ax.i1 != ax.i0 ? w1b : 0xffffu) << 16);
      if (s == t.s_lo) s_first[c] = ax.i0;
      if (s == t.s_hi - 1) s_last[c] = ax.i1;
    }
  }
  __syncthreads();
  int vbase[NC + 1];
  vbase[0] = 0;
#pragma unroll
  for (int c = 0; c < NC; ++c) vbase[c + 1] = vbase[c] + (c / 3 < p.nh ? s_last[c] - s_first[c] + 1 : 0);
  const int nvox = vbase[NC];
  // ---- pixel cells: a step opens a new cell when its key differs from the previous step's ----
  const bool isnew = key >= 0 && (s == 0 || s_key[s - 1] != key);
  const uint32_t bal = __ballot_sync(0xffffffffu, isnew);
  if (lane == 0) s_scan[w] = __popc(bal);
  __syncthreads();
  int woff = 0;
#pragma unroll
  for (int i = 0; i < BM / 32; ++i) woff += i < w ? s_scan[i] : 0;
  const int mycell = woff + __popc(bal & ((1u << lane) - 1u)) + (isnew ? 1 : 0) - 1;   // cell of this step (-1: none yet)
  if (isnew) s_ckey[mycell] = key;
  __syncthreads();
  // ---- which nodes of a new cell are new (not shared with the previous cell) ----
  int newmask = 0, validmask = 0;
  if (isnew) {
    const int pk = mycell > 0 ? s_ckey[mycell - 1] : -1;
    const int py0 = pk >= 0 ? pk / S : 0, px0 = pk >= 0 ? pk - py0 * S : 0;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int nx = x0 + (n & 1), ny = y0 + (n >> 1);
      const bool v = nx <= S - 1 && ny <= S - 1;
      const bool sh = pk >= 0 && static_cast<unsigned>(nx - px0) <= 1u && static_cast<unsigned>(ny - py0) <= 1u;
      if (v) validmask |= 1 << n;
      if (v && !sh) newmask |= 1 << n;
    }
  }
  const int cnt = __popc(newmask);
  int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) s_scan[4 + w] = inc;
  __syncthreads();
  int wbase = 0, npix = 0;
#pragma unroll
  for (int i = 0; i < BM / 32; ++i) { wbase += i < w ? s_scan[4 + i] : 0; npix += s_scan[4 + i]; }
  const int base = wbase + inc - cnt;
  if (isnew) { s_cbase[mycell] = base; s_cmask[mycell] = newmask | (validmask << 4); }
  __syncthreads();
  if (isnew) {
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      int slot = -1;
      if ((validmask >> n) & 1) {
        const int nx = x0 + (n & 1), ny = y0 + (n >> 1);
        if ((newmask >> n) & 1) {
          slot = base + __popc(newmask & ((1 << n) - 1));
          rows[nvox + slot] = reinterpret_cast<unsigned long long>(p.pmap + static_cast<size_t>(ny * S + nx) * N0);
        } else {
          for (int cc = mycell - 1; cc >= 0; --cc) {              // the node keeps the slot of the cell that introduced it
            const int ck = s_ckey[cc];
            const int cy0 = ck / S, cx0 = ck - cy0 * S;
            const int which = (nx - cx0) + 2 * (ny - cy0);
            const int m = s_cmask[cc];
            if ((m >> which) & 1) { slot = s_cbase[cc] + __popc(m & ((1 << which) - 1)); break; }
          }
        }
      }
      s_cslot[mycell * 4 + n] = static_cast<short>(slot);
    }
  }
  // ---- voxel rows: contiguous node ranges of the line table ----
  const size_t line_row0 = static_cast<size_t>(t.line_rel) * p.rpl;
  for (int e = it; e < nvox; e += BM) {
    int c = 0, vb = 0;
#pragma unroll
    for (int i = 1; i < NC; ++i)
      if (e >= vbase[i]) { c = i; vb = vbase[i]; }
    const int h = c / 3;
    rows[e] = reinterpret_cast<unsigned long long>(
        p.G + (line_row0 + static_cast<size_t>(p.rowbase[h] + (c - 3 * h) * p.R[h] + s_first[c] + (e - vb))) * N0);
  }
  const int n_rows = nvox + npix;
  const int n_chunks = (n_rows + BK - 1) / BK;
  for (int e = n_rows + it; e < n_chunks * BK; e += BM) rows[e] = reinterpret_cast<unsigned long long>(p.zero);
  if (it == 0) out.hdr[tile] = n_chunks | (((n_rows - (n_chunks - 1) * BK + 15) >> 4) << 8);
  __syncthreads();
  // ---- weight entries of this step: chunk << 23 | byte offset in the row's 128 B (16-byte pieces XOR-swizzled with the
  //      row index, as the K-major SWIZZLE_128B operand layout wants) << 16 | bf16 weight ----
  auto entry = [&](int slot, uint32_t wbits) -> uint32_t {
    const uint32_t col = static_cast<uint32_t>(slot) & 63u;
    const uint32_t off = (((col >> 3) ^ static_cast<uint32_t>(s & 7)) << 4) | ((col & 7u) << 1);
    return (static_cast<uint32_t>(slot >> 6) << 23) | (off << 16) | (wbits & 0xffffu);
  };
  uint32_t ev[kEntPad];
#pragma unroll
  for (int e = 0; e < kEntPad; ++e) ev[e] = kEmpty;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (c / 3 < p.nh && valid) {
      const int slot0 = vbase[c] + vi0[c] - s_first[c];
      ev[2 * c] = entry(slot0, vw[c] & 0xffffu);
      if ((vw[c] >> 16) != 0xffffu) ev[2 * c + 1] = entry(slot0 + 1, vw[c] >> 16);
    }
  }
  if (key >= 0) {
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int cs = s_cslot[mycell * 4 + n];
      if (cs >= 0) ev[2 * NC + n] = entry(nvox + cs, bf16_bits(pw[n]));
    }
  }
#pragma unroll
  for (int i = 0; i < kEntPad / 4; ++i)
    reinterpret_cast<uint4*>(ent)[i] = make_uint4(ev[4 * i], ev[4 * i + 1], ev[4 * i + 2], ev[4 * i + 3]);
}

// I chunk number ci of a tile pair in ring order -> (row list L, chunk c of that list).  Ring order: list 0 chunks [0, a0),
// list 1 chunks [0, a1) | the F chunks | list 0 chunks [a0, nI0), list 1 chunks [a1, nI1), with a = min(n, kIFirst).
__device__ __forceinline__ void chunk_of_seq(int ci, int nI0, int nI1, int& L, int& c) {
  const int a0 = min(nI0, kIFirst), a1 = min(nI1, kIFirst), nIa = a0 + a1;
  if (ci < a0) { L = 0; c = ci; }
  else if (ci < nIa) { L = 1; c = ci - a0; }
  else if (ci < nIa + (nI0 - a0)) { L = 0; c = ci - nIa + a0; }
  else { L = 1; c = ci - nIa - (nI0 - a0) + a1; }
}

// ------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(kThreads, 1)
grid_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW0,
               const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ volatile int s_flag;                    // interp warps: "the next chunk's units are already granted"
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* const gbase = smem_raw + (base - raw);
  float* const s_par = reinterpret_cast<float*>(gbase + OFF_PAR);
  float* const s_b1 = s_par;
  float* const s_b2 = s_b1 + N1;
  float* const s_w3 = s_b2 + N2;
  const uint32_t bar0 = base + OFF_BAR;
  auto full_bar = [&](uint32_t b) { return bar0 + 8u * b; };
  auto empty_bar = [&](uint32_t b) { return bar0 + 8u * (NB + b); };
  auto grant_bar = [&](uint32_t b) { return bar0 + 8u * (2 * NB + b); };
  auto ifull_bar = [&](uint32_t b) { return bar0 + 8u * (3 * NB + b); };
  const uint32_t dfull_bar = bar0 + 8u * (4 * NB);
  const uint32_t hready_bar = dfull_bar + 8u;
  const uint32_t tmem_slot = bar0 + 8u * NUM_BARS;
  volatile uint32_t* const tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + OFF_BAR + NUM_BARS * 8);
  auto unit_addr = [&](uint32_t u) { return base + (u % NU) * UNIT_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x / CG;
  const int num_clusters = gridDim.x / CG;
  const int n_pairs = static_cast<int>((p.n_tiles + 1) / 2);
  const int nkF = p.nkF;
  // 64-row interpolation chunks of a tile's row list (grid_plan_kernel); 0 for the tile past the end of an odd launch
  auto hdr_of = [&](unsigned tile) -> int { return tile < p.n_tiles ? __ldg(p.plan.hdr + tile) : 0; };
  auto chunks_of = [&](unsigned tile) -> int { return hdr_of(tile) & 0xff; };
  // I chunks of a pair as (before the F chunks) | (after them) << 16
  auto counts_of = [&](int pair) -> int {
    if (pair >= n_pairs) return 0;
    const int n0 = chunks_of(2u * pair), n1 = chunks_of(2u * pair + 1);
    const int a = min(n0, kIFirst) + min(n1, kIFirst);
    return a | ((n0 + n1 - a) << 16);
  };

  constexpr int kTraceTiles = 16, kTraceSlots = 24;
  const bool tracing = p.trace != nullptr && blockIdx.x == 0;
  auto stamp = [&](int tile_no, int slot) {
    if (tracing && tile_no % p.trace_stride == 0 && tile_no / p.trace_stride < kTraceTiles) {
      p.trace[(tile_no / p.trace_stride) * kTraceSlots + slot] = clock64();
      if (slot == 0) {                                            // slot 17: wall clock (ns) of the same instant -> SM clock under load
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        p.trace[(tile_no / p.trace_stride) * kTraceSlots + 17] = static_cast<long long>(ns);
      }
    }
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW0);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    for (int b = 0; b < NB; ++b) {
      mbar_init(full_bar(b), 1);                      // TMA chunks: the leader's expect_tx covers both CTAs' bytes
      mbar_init(empty_bar(b), 1);
      mbar_init(grant_bar(b), 1);
      mbar_init(ifull_bar(b), CG);                    // I chunks: one arrival per CTA (the interp warp that filled it)
    }
    mbar_init(dfull_bar, 1);
    mbar_init(hready_bar, kEpiWarps * CG);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<CG>(tmem_slot);
  if (warp >= kEpiWarp0 && warp < kIntWarp0) {
    for (int i = threadIdx.x - kEpiWarp0 * 32; i < PARAM_FLOATS; i += kEpiWarps * 32) {
      float v;
      if (i < N1) v = __ldg(p.b1 + i);
      else if (i < N1 + N2) v = __ldg(p.b2 + i - N1);
      else v = __ldg(p.w3 + i - N1 - N2);
      s_par[i] = v;
    }
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // =========================== TMA producer / ring allocator ===========================
    if (lane == 0) {
      uint32_t q = 0, head = 0, tail_q = 0;
      int free_units = NU;
      unsigned long long fifo = 0;                    // unit counts of the chunks in flight, oldest in bit 0 (1 = 3 units)
      int fifo_n = 0;
      auto make_room = [&](int n) {
        while (free_units < n) {
          mbar_wait(empty_bar(tail_q % NB), (tail_q / NB) & 1);
          free_units += (fifo & 1ull) ? U_FI : 1;
          fifo >>= 1;
          --fifo_n;
          ++tail_q;
        }
        free_units -= n;
        fifo |= static_cast<unsigned long long>(n == U_FI ? 1 : 0) << fifo_n;
        ++fifo_n;
      };
      // the leader announces the bytes of BOTH CTAs on its full barrier (its only arrival)
      auto announce = [&](uint32_t b, uint32_t bytes_per_cta) {
        if (rank == 0) mbar_expect_tx(full_bar(b), CG * bytes_per_cta);
      };
      int cnt = counts_of(cluster_id);
      for (int pair = cluster_id; pair < n_pairs; pair += num_clusters) {
        int64_t g0;
        int s_lo, s_hi;
        tile_rows(p.tm, 2u * pair + rank, p.n_tiles, g0, s_lo, s_hi);
        const int row0 = static_cast<int>(g0 - p.tm.begin);   // may be negative / past the end: TMA zero-fills those rows
        const int nIa = cnt & 0xffff, nIb = cnt >> 16;
        for (int ci = 0; ci < nIa; ++ci, ++q, head += U_FI) {     // units for the interp warps
          make_room(U_FI);
          mbar_arrive_local(grant_bar(q % NB));
        }
        for (int kc = 0; kc < nkF; ++kc, ++q, head += U_FI) {
          make_room(U_FI);
          const uint32_t b = q % NB;
          const uint32_t fb = mapa(full_bar(b), 0);
          announce(b, U_FI * UNIT_BYTES);
          tma_load_2d<CG>(&tmX, fb, unit_addr(head), kc * BK, row0);
#pragma unroll
          for (int j = 0; j < 2; ++j)
            tma_load_2d<CG>(&tmW0, fb, unit_addr(head + 1 + j), kc * BK, j * 256 + static_cast<int>(rank) * 128);
        }
        for (int ci = 0; ci < nIb; ++ci, ++q, head += U_FI) {
          make_room(U_FI);
          mbar_arrive_local(grant_bar(q % NB));
        }
        cnt = counts_of(pair + num_clusters);                     // next tile pair's counts, fetched under this tile's loads
#pragma unroll 1
        for (int layer = 1; layer <= 2; ++layer) {
          const CUtensorMap* tm = (layer == 1) ? &tmW1 : &tmW2;
          const int nk = (layer == 1 ? N0 : N1) / BK;
          for (int kc = 0; kc < nk; ++kc, ++q, ++head) {
            make_room(1);
            const uint32_t b = q % NB;
            announce(b, UNIT_BYTES);
            tma_load_2d<CG>(tm, mapa(full_bar(b), 0), unit_addr(head), kc * BK, static_cast<int>(rank) * 128);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA, one thread) ===========================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = umma_idesc(128 * CG, 128 * CG);
      constexpr uint32_t idesc_mn = idesc | (1u << 16);           // B operand MN-major (rows of P / G are N-contiguous)
      constexpr uint32_t kCols = 128 * CG;
      uint32_t q = 0, head = 0, hphase = 0, fphase = 0, iphase = 0;
      auto wait_full = [&]() {                                        // TMA chunk
        const uint32_t b = q % NB;
        mbar_wait(full_bar(b), (fphase >> b) & 1u);
        fphase ^= 1u << b;
        tc_fence_after();
      };
      auto wait_ifull = [&]() {                                       // I chunk: filled by the interp warps of both CTAs
        const uint32_t b = q % NB;
        mbar_wait(ifull_bar(b), (iphase >> b) & 1u);
        iphase ^= 1u << b;
        tc_fence_after();
      };
      int itn = 0;
      unsigned long long n_ksteps_i = 0;
      int h0 = hdr_of(2u * cluster_id), h1 = hdr_of(2u * cluster_id + 1);
      for (int pair = cluster_id; pair < n_pairs; pair += num_clusters, ++itn) {
        // ---- fc_0: D[0,512) = Aw . Brows (interpolated part, I chunks) + Xr . W0[:, hoisted..]^T (dense part, F chunks) ----
        const int nI0 = h0 & 0xff, nI1 = h1 & 0xff, kl0 = h0 >> 8, kl1 = h1 >> 8;
        const int nIa = min(nI0, kIFirst) + min(nI1, kIFirst), nI = nI0 + nI1;
        // 16-row k-steps of an I chunk: the tail of a tile's last chunk holds only padding rows and is skipped
        auto ksteps_of = [&](int ci) -> int {
          int L, c;
          chunk_of_seq(ci, nI0, nI1, L, c);
          return L == 0 ? (c == nI0 - 1 ? kl0 : BK / 16) : (c == nI1 - 1 ? kl1 : BK / 16);
        };
        // one 256-column half j of an I chunk
        auto issue_i_half = [&](uint32_t hd, int j, int ks, uint32_t acc) {
          const uint64_t ad = umma_desc_sw128(unit_addr(hd));
          const uint64_t bd = umma_desc_mn_sw128(unit_addr(hd + 1 + j), p.b_lbo, p.b_sbo);
          for (int k = 0; k < ks; ++k)                                  // 16 k rows = 2048 B further along K
            umma_ss<CG>(tmem_base + j * kCols, ad + 2 * k, bd + static_cast<uint64_t>((p.b_kadv >> 4) * k), idesc_mn,
                        acc | static_cast<uint32_t>(k != 0));
        };
        uint32_t acc = 0;                                             // the tile's first MMA overwrites the accumulator
        // Column half j = 0 of the first kPre chunks is issued BEFORE the previous tile's last epilogue has drained its
        // accumulator: that one lives in columns [256,512), and the tensor pipe executes MMAs in issue order, so the
        // previous fc_2 has read H2 (columns [0,128)) by the time these write columns [0,256).  (Only two I chunks fit
        // into the ring next to fc_2's weight boxes, so only two can be ready that early.)
        const int nPre = min(nIa, kPre);
        for (int ci = 0; ci < nPre; ++ci) {
          const uint32_t b = (q + ci) % NB;
          mbar_wait(ifull_bar(b), (iphase >> b) & 1u);
          iphase ^= 1u << b;
          tc_fence_after();
          issue_i_half(head + U_FI * ci, 0, ksteps_of(ci), ci != 0 ? 1u : 0u);
        }
        if (itn > 0) { mbar_wait(hready_bar, hphase); hphase ^= 1; }   // previous tile's accumulators drained
        tc_fence_after();
        stamp(itn, 0);
        for (int ci = 0; ci < nPre; ++ci) {
          const int ks = ksteps_of(ci);
          issue_i_half(head + U_FI * ci, 1, ks, ci != 0 ? 1u : 0u);
          umma_commit<CG>(empty_bar((q + ci) % NB));
          n_ksteps_i += static_cast<unsigned long long>(ks);
        }
        q += nPre;
        head += U_FI * nPre;
        if (nPre > 0) acc = 1;
        auto issue_i = [&](int c0, int c1) {
          for (int ci = c0; ci < c1; ++ci, head += U_FI) {
            wait_ifull();
            const int ks = ksteps_of(ci);
            const uint64_t ad = umma_desc_sw128(unit_addr(head));
            for (int k = 0; k < ks; ++k) {
#pragma unroll
              for (int j = 0; j < 2; ++j)
                umma_ss<CG>(tmem_base + j * kCols, ad + 2 * k,
                            umma_desc_mn_sw128(unit_addr(head + 1 + j), p.b_lbo, p.b_sbo) + static_cast<uint64_t>((p.b_kadv >> 4) * k),
                            idesc_mn, acc | static_cast<uint32_t>(k != 0));
            }
            acc = 1;
            umma_commit<CG>(empty_bar(q % NB));
            n_ksteps_i += static_cast<unsigned long long>(ks);
            ++q;
          }
        };
        issue_i(nPre, nIa);
        stamp(itn, 13);
        for (int kc = 0; kc < nkF; ++kc, head += U_FI) {
          wait_full();
          const uint64_t ad = umma_desc_sw128(unit_addr(head));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
              umma_ss<CG>(tmem_base + j * kCols, ad + 2 * k, umma_desc_sw128(unit_addr(head + 1 + j)) + 2 * k, idesc,
                          acc | static_cast<uint32_t>(k != 0));
          }
          acc = 1;
          umma_commit<CG>(empty_bar(q % NB));
          ++q;
        }
        stamp(itn, 12);
        issue_i(nIa, nI);
        h0 = hdr_of(2u * (pair + num_clusters));
        h1 = hdr_of(2u * (pair + num_clusters) + 1);
        umma_commit<CG>(dfull_bar);
        stamp(itn, 1);
        // ---- fc_1: D[256,512) = H1(TMEM [0,256)) . W1^T ;  fc_2: D[256,512) = H2(TMEM [0,128)) . W2^T ----
#pragma unroll 1
        for (int layer = 1; layer <= 2; ++layer) {
          const int nk = (layer == 1 ? N0 : N1) / BK;
          mbar_wait(hready_bar, hphase); hphase ^= 1;
          tc_fence_after();
          stamp(itn, 2 * layer);
          for (int kc = 0; kc < nk; ++kc, ++head) {
            wait_full();
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_ts<CG>(tmem_base + 256, tmem_base + kc * (BK / 2) + k * 8, umma_desc_sw128(unit_addr(head)) + 2 * k, idesc,
                          (kc | k) != 0 ? 1u : 0u);
            umma_commit<CG>(empty_bar(q % NB));
            ++q;
          }
          umma_commit<CG>(dfull_bar);
          stamp(itn, 2 * layer + 1);
        }
      }
      if (p.stats != nullptr) {
        atomicAdd(p.stats, static_cast<unsigned long long>(itn));
        atomicAdd(p.stats + 1, n_ksteps_i);
      }
    }
  } else if (warp < kIntWarp0) {
    // =========================== epilogue warps ===========================
    // Two warps per TMEM lane quarter (a warp may only touch lanes 32 * (warp % 4) .. + 31): `half` selects the columns.
    const int quarter = warp & 3, half = (warp - kEpiWarp0) >> 2;
    const uint32_t tq = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t hready_remote = mapa(hready_bar, 0);
    const float bias3 = __ldg(p.b3);
    const int r_in_tile = quarter * 32 + lane;
    float* const s_part = reinterpret_cast<float*>(gbase + OFF_PART);
    auto pair_sync = [&]() { named_bar_sync(3 + quarter, 64); };     // the two warps of this lane quarter
    auto arrive_hready = [&]() {
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(hready_remote);
    };
    uint32_t dphase = 0;
    int itn = 0;
    const bool estamp = warp == kEpiWarp0 && lane == 0;
    for (int pair = cluster_id; pair < n_pairs; pair += num_clusters, ++itn) {
      int64_t g0;
      int s_lo, s_hi;
      tile_rows(p.tm, 2u * pair + rank, p.n_tiles, g0, s_lo, s_hi);
      const bool live = r_in_tile >= s_lo && r_in_tile < s_hi;
      const int64_t orow = g0 + r_in_tile - p.tm.begin;
      // ---- after fc_0: H1 = relu(acc) -> bf16 -> TMEM [0,256), compacted in place (acc includes b0).  Step m: the pair reads accumulator
      //      columns [64 m, +64) and writes H1 columns [32 m, +32); the pair barrier between the loads and the stores keeps a
      //      warp from overwriting columns its partner has not read yet (step m's stores never reach a later step's loads). ----
      mbar_wait_warp(dfull_bar, dphase); dphase ^= 1;
      tc_fence_after();
      if (estamp) stamp(itn, 6);
      // The loads run one step ahead of the arithmetic (two register buffers): step m + 1's columns lie beyond everything
      // step m writes, so only the stores wait for the partner's loads.
      {
        uint32_t va[32], vb[32];
        auto finish = [&](int m, uint32_t (&v)[32]) {
          const int j = 2 * m + half;
          uint32_t u[16];                                           // the bias came in through the MMA (Plan::bias_col)
#pragma unroll
          for (int i = 0; i < 16; ++i) u[i] = pack_bf16x2_relu(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
          tmem_st16(tq + j * 16, u);
          if (p.dbg1 != nullptr && live) {
#pragma unroll
            for (int i = 0; i < 32; ++i) p.dbg1[orow * N0 + j * 32 + i] = fmaxf(__uint_as_float(v[i]), 0.f);
          }
        };
        tmem_ld32_nowait(tq + half * 32, va);
#pragma unroll 1
        for (int m = 0; m < N0 / 64; m += 2) {
          tmem_ld_wait32(va);
          pair_sync();
          tmem_ld32_nowait(tq + (2 * (m + 1) + half) * 32, vb);
          finish(m, va);
          tmem_ld_wait32(vb);
          pair_sync();
          if (m + 2 < N0 / 64) tmem_ld32_nowait(tq + (2 * (m + 2) + half) * 32, va);
          finish(m + 1, vb);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      if (estamp) stamp(itn, 7);
      arrive_hready();
      // ---- after fc_1: H2 = relu(acc + b1) -> bf16 -> TMEM [0,128) (reads [256,512): no overlap) ----
      mbar_wait_warp(dfull_bar, dphase); dphase ^= 1;
      tc_fence_after();
      if (estamp) stamp(itn, 8);
      {
        uint32_t va[32], vb[32];
        auto finish = [&](int m, uint32_t (&v)[32]) {
          const int j = half * (N1 / 64) + m;
          uint32_t u[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 bb = *reinterpret_cast<const float2*>(s_b1 + j * 32 + 2 * i);
            const float2 sum = fadd2(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), bb);
            u[i] = pack_bf16x2_relu(sum.x, sum.y);
          }
          tmem_st16(tq + j * 16, u);
        };
        const uint32_t src = tq + 256 + half * (N1 / 64) * 32;
        tmem_ld32_nowait(src, va);
#pragma unroll 1
        for (int m = 0; m < N1 / 64; m += 2) {
          tmem_ld_wait32(va);
          tmem_ld32_nowait(src + (m + 1) * 32, vb);
          finish(m, va);
          tmem_ld_wait32(vb);
          if (m + 2 < N1 / 64) tmem_ld32_nowait(src + (m + 2) * 32, va);
          finish(m + 1, vb);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      if (estamp) stamp(itn, 9);
      arrive_hready();
      // ---- after fc_2: sdf = (relu(acc + b2) . w3 + b3) / out_div; the upper column half hands its partial sum over ----
      mbar_wait_warp(dfull_bar, dphase); dphase ^= 1;
      tc_fence_after();
      if (estamp) stamp(itn, 10);
      float2 acc2 = make_float2(0.f, 0.f);
      {
        uint32_t va[32], vb[32];
        auto finish = [&](int m, uint32_t (&v)[32]) {
          const int j = half * (N2 / 64) + m;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 bb = *reinterpret_cast<const float2*>(s_b2 + j * 32 + 2 * i);
            const float2 ww = *reinterpret_cast<const float2*>(s_w3 + j * 32 + 2 * i);
            const float2 sum = fadd2(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), bb);
            acc2 = ffma2(make_float2(fmaxf(sum.x, 0.f), fmaxf(sum.y, 0.f)), ww, acc2);
          }
        };
        const uint32_t src = tq + 256 + half * (N2 / 64) * 32;
        tmem_ld32_nowait(src, va);
#pragma unroll 1
        for (int m = 0; m < N2 / 64; m += 2) {
          tmem_ld_wait32(va);
          tmem_ld32_nowait(src + (m + 1) * 32, vb);
          finish(m, va);
          tmem_ld_wait32(vb);
          if (m + 2 < N2 / 64) tmem_ld32_nowait(src + (m + 2) * 32, va);
          finish(m + 1, vb);
        }
      }
      tc_fence_before();
      if (estamp) stamp(itn, 11);
      arrive_hready();
      if (half == 1) s_part[r_in_tile] = acc2.x + acc2.y;
      pair_sync();
      if (half == 0 && live) p.sdf[orow] = __fdiv_rn(((acc2.x + acc2.y) + s_part[r_in_tile]) + bias3, p.out_div);
      pair_sync();                                                  // s_part is free again
    }
  } else {
    // =========================== interp warps ===========================
    const int it = threadIdx.x - kIntWarp0 * 32;                   // 0..127
    const int wq = it >> 5;
    const uint32_t ent_addr = base + OFF_ENT, rows_addr = base + OFF_ROWS;
    uint32_t* const s_ent = reinterpret_cast<uint32_t*>(gbase + OFF_ENT);
    const unsigned long long* const s_rows = reinterpret_cast<const unsigned long long*>(gbase + OFF_ROWS);
    // this lane's part of a listed row: j = which 256-channel block of N0, g = which 64-channel group of the CTA's 128, c16 = 16-byte piece
    const int bj = lane >> 4, bg = (lane >> 3) & 1, bc = lane & 7;
    const int col_off = bj * 256 + static_cast<int>(rank) * 128 + bg * 64 + bc * 8;
    auto ibar = [&]() { named_bar_sync(2, kIntWarps * 32); };
    uint32_t q = 0, head = 0, gphase = 0;
    int itn = 0;
    for (int pair = cluster_id; pair < n_pairs; pair += num_clusters, ++itn) {
      if (it == 0) stamp(itn, 16);
      ibar();                                                      // every warp is done with the previous tile's lists
      // ---- this tile pair's plan: both row lists, and the entries of this CTA's own tile ----
      const int nI0 = chunks_of(2u * pair), nI1 = chunks_of(2u * pair + 1);
      const unsigned own = 2u * pair + rank;
      if (own < p.n_tiles) {
        const char* src = reinterpret_cast<const char*>(p.plan.ent + static_cast<size_t>(own) * BM * kEntPad);
        for (int i = it; i < kPlanEntBytes / 16; i += kIntWarps * 32) cp_async16(ent_addr + 16u * i, src + 16 * i);
      }
#pragma unroll
      for (int L = 0; L < 2; ++L) {
        const int n16 = (L ? nI1 : nI0) * BK * 8 / 16;
        const char* src = reinterpret_cast<const char*>(p.plan.rows + static_cast<size_t>(2u * pair + L) * kMaxRows);
        for (int i = it; i < n16; i += kIntWarps * 32) cp_async16(rows_addr + static_cast<uint32_t>(L * kPlanRowBytes) + 16u * i, src + 16 * i);
      }
      cp_async_wait_all();
      ibar();
      if (it == 0) stamp(itn, 14);
      const int nI = nI0 + nI1;
      const int nIa = min(nI0, kIFirst) + min(nI1, kIFirst);
      // Chunks are filled one step ahead of their hand-over: chunk ci's row copies are in flight while chunk ci + 1 is
      // granted, its copies issued and its weights written; only then are ci's copies awaited and ci handed to the MMA
      // thread (the copies' L2 latency, ~1k cycles under load, is off the per-chunk critical path).  If the units of
      // chunk ci + 1 are not free yet, chunk ci is handed over first (the ring may be waiting for exactly that chunk).
      int pend_b = -1;
      auto hand_over = [&](int b) {                                // this thread's copies of the chunk have landed
        fence_proxy_async_smem();
        ibar();                                                    // all four quarters are in place
        if (it == 0) {
          if (rank == 0) mbar_arrive_local(ifull_bar(b));
          else mbar_arrive_remote(mapa(ifull_bar(b), 0));
        }
      };
      for (int ci = 0; ci < nI; ++ci, ++q, head += U_FI) {
        if (ci == nIa) { q += nkF; head += U_FI * nkF; }           // the pair's F chunks sit between its two groups of I chunks
        const uint32_t b = q % NB;
        const uint32_t par = (gphase >> b) & 1u;
        gphase ^= 1u << b;
        int L, c;
        chunk_of_seq(ci, nI0, nI1, L, c);
        if (pend_b >= 0) {
          if (it == 0) s_flag = mbar_test(grant_bar(b), par) ? 1 : 0;
          ibar();
          if (s_flag == 0) {
            cp_async_wait_all();
            hand_over(pend_b);
            pend_b = -1;
          }
        }
        mbar_wait_warp(grant_bar(b), par);
        if (it == 0 && ci == 0) stamp(itn, 19);
        // Every chunk is filled by all four warps, a quarter of its rows each: what matters is the latency from the grant
        // to the chunk being usable, not the throughput.
        // ---- Brows k in [16 wq, +16): this CTA's 256 channels of the listed rows, MN-major 128B-swizzled (row k of a
        //      64-channel group at (k / 8) * 1024 + (k % 8) * 128, 16-byte piece i at (i ^ (k % 8)) * 16; groups 8 KB apart) ----
        const uint32_t ub = unit_addr(head + 1 + bj) + static_cast<uint32_t>(bg) * 8192u + static_cast<uint32_t>(wq) * 2048u;
        const ulonglong2* const lrows = reinterpret_cast<const ulonglong2*>(s_rows + L * kMaxRows + c * BK + wq * 16);
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {
          const ulonglong2 rr2 = lrows[k2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int k = 2 * k2 + u;                              // row 16 wq + k of the chunk: (k >> 3) is the 8-row group inside this quarter
            cp_async16(ub + static_cast<uint32_t>(k >> 3) * 1024u + static_cast<uint32_t>(k & 7) * 128u +
                           (static_cast<uint32_t>(bc ^ (k & 7)) << 4),
                       reinterpret_cast<const __nv_bfloat16*>(u ? rr2.y : rr2.x) + col_off);
          }
        }
        cp_async_commit();
        if (it == 0 && ci == 0) stamp(itn, 20);
        // ---- Aw rows [32 wq, +32): zeros, then this CTA's weights if the list is its own tile's ----
        const uint32_t ua = unit_addr(head) + static_cast<uint32_t>(wq) * (32u * 128u);
#pragma unroll
        for (int i = 0; i < 32 * 128 / 16 / 32; ++i) st_shared_zero16(ua + static_cast<uint32_t>(i * 32 + lane) * 16u);
        __syncwarp();
        if (L == static_cast<int>(rank)) {
          const int row = wq * 32 + lane;
          const uint32_t ra = ua + static_cast<uint32_t>(lane) * 128u;
          const uint4* const ev = reinterpret_cast<const uint4*>(s_ent + row * kEntPad);
          uint4 e4[kEntPad / 4];
#pragma unroll
          for (int i = 0; i < kEntPad / 4; ++i) e4[i] = ev[i];
#pragma unroll
          for (int i = 0; i < kEntPad / 4; ++i) {
            const uint32_t vv[4] = {e4[i].x, e4[i].y, e4[i].z, e4[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t v = vv[j];
              if (static_cast<int>(v >> 23) == c) st_shared_u16(ra + ((v >> 16) & 0x7fu), v & 0xffffu);
            }
          }
        }
        if (it == 0 && ci == 0) stamp(itn, 21);
        if (pend_b >= 0) {
          cp_async_wait_prior1();                                  // the previous chunk's copies (this chunk's stay in flight)
          hand_over(pend_b);
        }
        pend_b = static_cast<int>(b);
        if (it == 0 && ci == 0) stamp(itn, 22);
      }
      if (pend_b >= 0) {
        cp_async_wait_all();
        hand_over(pend_b);
      }
      if (nI <= nIa) { q += nkF; head += U_FI * nkF; }
      if (it == 0) stamp(itn, 15);
      q += N_W;
      head = (head + N_W) % NU;
    }
  }

  // =========================== teardown ===========================
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base);
  }
}

}  // namespace gtc

static int fill_geo(const ListCtx* ctx, const hoist::Plan& pl, const void* hoist_buf, int image, int res, double bb_min, double bb_max,
                    int64_t begin, int64_t count, const void* G, gtc::Geo* g) {
  using namespace gtc;
  LIST_CHECK_ARG(pl.nh <= kMaxLev, "grid_tc: plan with %d hoisted levels", pl.nh);
  int vox_rows = 0;
  for (int h = 0; h < pl.nh; ++h) {
    LIST_CHECK_ARG(ctx->vol_res[pl.lev[h]] <= 32, "grid_tc: hoisted level with R = %d > 32", ctx->vol_res[pl.lev[h]]);
    vox_rows += 3 * ctx->vol_res[pl.lev[h]];
  }
  LIST_CHECK_ARG(vox_rows + 4 * BM <= kMaxRows, "grid_tc: %d voxel rows per tile exceed the row list", vox_rows);
  hoist::fill_tilemap(&g->tm, res, bb_min, bb_max, begin, count, BM);
  g->n_tiles = hoist::tile_count(g->tm);
  const char* hb = static_cast<const char*>(hoist_buf);
  g->pmap = reinterpret_cast<const __nv_bfloat16*>(hb + pl.off_pmap) + static_cast<size_t>(image) * ctx->map_size * ctx->map_size * N0;
  g->G = static_cast<const __nv_bfloat16*>(G);
  g->zero = reinterpret_cast<const __nv_bfloat16*>(hb + pl.off_zero);
  g->T = ctx->trans_mat + image * 12;
  g->S = ctx->map_size;
  g->nh = pl.nh;
  g->rpl = pl.rpl;
  for (int h = 0; h < kMaxLev; ++h) {
    g->R[h] = h < pl.nh ? ctx->vol_res[pl.lev[h]] : 1;
    g->rowbase[h] = h < pl.nh ? pl.rowbase[h] : 0;
  }
  return LIST_OK;
}

// Bytes of the tile plans of grid points [begin, begin + count) of a res^3 grid.
size_t grid_plan_bytes(int res, int64_t begin, int64_t count) {
  if (count <= 0) return 0;
  hoist::TileMap tm;
  hoist::fill_tilemap(&tm, res, 0.0, 1.0, begin, count, gtc::BM);
  return gtc::plan_bytes(hoist::tile_count(tm));
}

// Tile plans (row lists + weight entries) of grid points [begin, begin + count) of image `image`; G is only used as an
// address (the plans point into it), so this may run before or after hoist::lines.
int grid_plan(const ListCtx* ctx, const hoist::Plan& pl, const void* hoist_buf, int image, int res, double bb_min, double bb_max,
              int64_t begin, int64_t count, const void* G, void* plan_buf, cudaStream_t st) {
  using namespace gtc;
  if (count == 0) return LIST_OK;
  Geo g{};
  const int rc = fill_geo(ctx, pl, hoist_buf, image, res, bb_min, bb_max, begin, count, G, &g);
  if (rc) return rc;
  LIST_CHECK_ARG((reinterpret_cast<uintptr_t>(plan_buf) & 255) == 0, "grid_plan: plan buffer must be 256-byte aligned");
  grid_plan_kernel<<<g.n_tiles, BM, 0, st>>>(g, plan_carve(plan_buf, g.n_tiles));
  LIST_LAUNCH_CHECK("grid_plan_kernel");
  return LIST_OK;
}

// Fused interpolation + MLP over grid points [begin, begin + count) of image `image` (see the header comment).
int grid_tc_fwd(const ListCtx* ctx, const ListWeights* w, const hoist::Plan& pl, const void* hoist_buf, int res, double bb_min, double bb_max,
                int64_t begin, int64_t count, const void* Xr, int64_t ldx, const void* plan_buf, float* sdf, float out_div, float* dbg1,
                long long* trace, unsigned long long* stats, cudaStream_t st) {
  using namespace gtc;
  if (count == 0) return LIST_OK;
  LIST_CHECK_ARG(w->n0 == N0 && w->n1 == N1 && w->n2 == N2, "grid_tc: layer widths must be 512/256/256 (got %d/%d/%d)", w->n0, w->n1, w->n2);
  LIST_CHECK_ARG(count < (1LL << 31), "grid_tc: %lld rows are too many for one launch", (long long)count);
  const int k_f = pl.k_h - N0;
  LIST_CHECK_ARG(k_f >= BK && k_f % BK == 0 && ldx >= k_f && ldx % 8 == 0, "grid_tc: %d dense columns / ldx %lld invalid", k_f, (long long)ldx);
  LIST_CHECK_ARG((reinterpret_cast<uintptr_t>(Xr) & 15) == 0 && (reinterpret_cast<uintptr_t>(plan_buf) & 255) == 0,
                 "grid_tc: Xr must be 16-byte and the plan buffer 256-byte aligned");
  Params p{};
  p.b1 = w->b1; p.b2 = w->b2; p.w3 = w->w3; p.b3 = w->b3;       // b0: in the weight copy of hoist::prepare (Plan::bias_col)
  p.sdf = sdf;
  p.out_div = out_div;
  p.nkF = k_f / BK;
  hoist::fill_tilemap(&p.tm, res, bb_min, bb_max, begin, count, BM);
  p.n_tiles = hoist::tile_count(p.tm);
  p.plan = plan_carve(const_cast<void*>(plan_buf), p.n_tiles);
  p.dbg1 = dbg1;
  p.trace = trace;
  {
    const char* e = getenv("LIST_B200_TRACE_STRIDE");          // tuning aid of scripts/grid_tc_trace.py
    const int v = e ? atoi(e) : 1;
    p.trace_stride = v >= 1 ? v : 1;
  }
  p.stats = stats;
  p.b_lbo = 8192; p.b_sbo = 1024; p.b_kadv = 2048;
  CUtensorMap tmX, tmW0, tmW1, tmW2;
  int rc;
  if ((rc = make_map_bf16(&tmX, Xr, static_cast<uint64_t>(k_f), static_cast<uint64_t>(count), static_cast<uint64_t>(ldx)))) return rc;
  // fc_0's non-hoisted weight columns with b0 in the bias columns (hoist::prepare); Xr carries 1.0 there
  if ((rc = make_map_bf16(&tmW0, static_cast<const char*>(hoist_buf) + pl.off_w0r, static_cast<uint64_t>(k_f), N0, static_cast<uint64_t>(k_f)))) return rc;
  if ((rc = make_map_bf16(&tmW1, w->w1, N0, N1, N0))) return rc;
  if ((rc = make_map_bf16(&tmW2, w->w2, N1, N2, N1))) return rc;
  int dev = 0, sms = 0;
  LIST_CUDA(cudaGetDevice(&dev));
  LIST_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static thread_local int attr_dev = -1;
  if (attr_dev != dev) {
    LIST_CUDA(cudaFuncSetAttribute(grid_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_dev = dev;
  }
  const int n_pairs = static_cast<int>((p.n_tiles + 1) / 2);
  const int clusters = n_pairs < sms / CG ? n_pairs : sms / CG;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * CG);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LIST_CUDA(cudaLaunchKernelEx(&cfg, grid_tc_kernel, tmX, tmW0, tmW1, tmW2, p));
  return LIST_OK;
}

}  // namespace list
