// Per-line column tables of the hoisted voxel levels on the tensor cores (bf16 dense-grid path, feeds grid_tc.cu; same
// output as lines.cu's SIMT kernel, see its header for the algebra and the reference lines).
//
// For the LY lines (lz, ly0 .. ly0 + LY - 1) of one x-plane the table rows are
//      G[(cls, line)][node i][:] = sum_{d in cls} sum_{y} wy_{d,line}(y) * Rd_{d,y}[i][:],
//      Rd_{d,y}[i][:] = wz0_d * PV_d[z0_d][y][i][:] + wz1_d * PV_d[z1_d][y][i][:]          (the D reduction, shared by the lines)
// i.e. per node i a GEMM  [3 LY x K] . [K x 512]  with K = the (displacement, H node) pairs the lines touch (<= 64) and
// sparse H weights.  The SIMT kernel spends ~14 FMA per table element on it; here
//   producer warps  : Rd rows of one (node, 256-channel half) -> bf16 -> shared memory, MN-major 128B-swizzled B operand
//                     (2 global loads + 1 shared store per 8 channels and K row),
//   MMA thread      : tcgen05.mma M = 128 (rows (cls, line)), N = 256, K = 16 per step, A = the H weights (written once per
//                     work item), accumulators double buffered in TMEM,
//   epilogue warps  : tcgen05.ld -> bf16 -> shared-memory staging ([lines][64 channels], 128B swizzle) -> TMA tensor store
//                     into G viewed as [plane][line][row][512] (work items that lie inside the launch's range; the items
//                     at its ragged ends store straight from registers, a lane per (cls, line) row).
// What is left is the write of G: the kernel is HBM-bound instead of FMA-bound.
//
// Work item = (x-plane, group of LY lines, part of the node range); persistent CTAs stride over the items.
// Numerics: Rd and the H weights are rounded to bf16 before the MMA (fp32 accumulation); the two H weights of a
// (displacement, line) are quantised so that they sum to 1 exactly, as grid_plan_kernel does for the z taps.
#include "grid_common.cuh"
#include "hoist.cuh"
#include "tc_common.cuh"

namespace list {
namespace hoist {

using namespace tc;

namespace ltc {
constexpr int BM = 128, BN = 256, BK = 64;
constexpr int kStages = 3;                          // B ring
constexpr int B_BYTES = BK * BN * 2;                // 32 KB: [4 channel groups of 64][64 k rows][128 B]
constexpr int A_BYTES = BM * BK * 2;                // 16 KB per level: K-major, 128B swizzle
constexpr int kMaxLY = 32;
constexpr int kEpiWarp0 = 1, kEpiWarps = 4, kProdWarp0 = kEpiWarp0 + kEpiWarps, kProdWarps = 8;
constexpr int kThreads = (kProdWarp0 + kProdWarps) * 32;   // 416
constexpr int kN0 = 512;
constexpr int kGroupsPerWarp = (BN / 64) / (kEpiWarps / 4);   // 64-channel groups of a tile per epilogue warp
static_assert(kEpiWarps == 4 || kEpiWarps == 8, "epilogue warps");

struct KRow {                                       // one K row of a level: which projected rows it reduces
  uint32_t off0, off1;                              // element offsets (node 0, channel 0) of the two D planes' rows
  float w0, w1;                                     // D weights (0 for padding rows)
};
struct LineTap {                                    // H interpolation of one (level, displacement, line)
  int i0, i1;
  float w0, w1;
};

constexpr int kStageBuf = 4096;                     // epilogue staging: [32 rows][64 channels] bf16, 128B swizzle, two per warp
constexpr int OFF_STG = kStages * B_BYTES;
constexpr int OFF_A = OFF_STG + kEpiWarps * 2 * kStageBuf;
constexpr int OFF_KROW = OFF_A + kMaxLev * A_BYTES;
constexpr int OFF_TAP = OFF_KROW + kMaxLev * BK * static_cast<int>(sizeof(KRow));
constexpr int OFF_MISC = OFF_TAP + kMaxLev * LIST_NUM_DISP * kMaxLY * static_cast<int>(sizeof(LineTap));
constexpr int MISC_INTS = 64;
constexpr int OFF_BAR = OFF_MISC + MISC_INTS * 4;
constexpr int NUM_BARS = 2 * kStages + 4;           // b_full, b_empty | acc_full[2], acc_empty[2]
constexpr int SMEM_BYTES = OFF_BAR + NUM_BARS * 8 + 16 + 1024;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
}  // namespace ltc

struct LinesTcParams {
  const __nv_bfloat16* pvol[kMaxLev];   // image's slab of displacement 0: [R][R][R][512]
  uint32_t dstride[kMaxLev];            // elements between displacement slabs
  int R[kMaxLev], rowbase[kMaxLev];
  int nh, rpl;
  int ly;                               // lines per work item (<= 32)
  int groups;                           // line groups per x-plane = ceil(res / ly)
  int parts, nt_total, nt_part;         // the 2 * sum R (node, half) tiles of an item are split into `parts` ranges
  unsigned n_items;
  int use_tma;                          // tensor stores for the items inside the range (ly in {8, 16, 32})
  int64_t line_first, line_last;        // lines of the launch (inclusive)
  __nv_bfloat16* G;
  TileMap tm;
};

__device__ __forceinline__ uint32_t bf16_bits_rn(float x) { return static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16_rn(x))); }

// Tensor store issued by the lanes whose `on` is non-zero (predicated, not branched: a divergent `if (lane == 0)` around a
// blocking instruction reconverges through a slow path on sm_100, see profiles/r02_grid_tc_notes.md).
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3, uint32_t on) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %6, 0;\n\t"
               "@p cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n\t}"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(on) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(ltc::kThreads, 1) hoist_lines_tc_kernel(const __grid_constant__ CUtensorMap tmG, const LinesTcParams p) {
  using namespace ltc;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* const gbase = smem_raw + (base - raw);
  KRow* const s_krow = reinterpret_cast<KRow*>(gbase + OFF_KROW);               // [kMaxLev][BK]
  LineTap* const s_tap = reinterpret_cast<LineTap*>(gbase + OFF_TAP);           // [kMaxLev][7][kMaxLY]
  int* const s_misc = reinterpret_cast<int*>(gbase + OFF_MISC);                 // [h*8 + d]: kbase | ymin << 8 ; [32 + h]: K rows of the level
  const uint32_t bar0 = base + OFF_BAR;
  auto b_full = [&](int s) { return bar0 + 8u * s; };
  auto b_empty = [&](int s) { return bar0 + 8u * (kStages + s); };
  auto acc_full = [&](int a) { return bar0 + 8u * (2 * kStages + a); };
  auto acc_empty = [&](int a) { return bar0 + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bar0 + 8u * NUM_BARS;
  volatile uint32_t* const tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + OFF_BAR + NUM_BARS * 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int res = p.tm.res;
  if (tid == 0) {
    tma_prefetch_desc(&tmG);
    for (int s = 0; s < kStages; ++s) { mbar_init(b_full(s), kProdWarps); mbar_init(b_empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(acc_full(a), 1); mbar_init(acc_empty(a), kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc<1>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // pipeline counters (persist across work items): B stages produced / consumed, accumulators used
  uint32_t n_b = 0, n_acc = 0;

  for (unsigned item = blockIdx.x; item < p.n_items; item += gridDim.x) {
    const int part = static_cast<int>(item % static_cast<unsigned>(p.parts));
    const unsigned pg = item / static_cast<unsigned>(p.parts);
    const unsigned grp = pg % static_cast<unsigned>(p.groups);
    const unsigned plane = pg / static_cast<unsigned>(p.groups);
    const int64_t lz = p.line_first / res + plane;
    const int ly0 = static_cast<int>(grp) * p.ly;
    const int t_lo = part * p.nt_part, t_hi = min(p.nt_total, t_lo + p.nt_part);
    // lines of this item inside the launch's range: j in [jlo, jhi)
    int jlo = 0, jhi = p.ly;
    {
      const int64_t l0 = lz * res + ly0;
      if (l0 < p.line_first) jlo = static_cast<int>(min(static_cast<int64_t>(p.ly), p.line_first - l0));
      const int64_t lend = min(static_cast<int64_t>(res) - ly0, p.line_last + 1 - l0);
      if (lend < jhi) jhi = static_cast<int>(max(static_cast<int64_t>(0), lend));
    }
    if (jlo >= jhi || t_lo >= t_hi) continue;                      // uniform over the CTA
    const int64_t out_line0 = lz * res + ly0 - p.line_first;       // may be negative for the lines below jlo

    __syncthreads();                                               // the previous item's tables and A operands are free
    // ---- geometry 1: H taps of every (level, displacement, line) ----
    for (int e = tid; e < p.nh * LIST_NUM_DISP * p.ly; e += kThreads) {
      const int j = e % p.ly, hd = e / p.ly, d = hd % LIST_NUM_DISP, h = hd / LIST_NUM_DISP;
      const int ly = min(ly0 + j, res - 1);
      float q[3] = {0.f, linspace_f32_step(ly, res, p.tm.bb_min, p.tm.bb_max, p.tm.step) * 2.0f, 0.f}, pd[3];
      displaced(q, d, pd);
      const Axis3 ay = axis_border(pd[1], p.R[h]);
      // the larger weight is rounded to bf16, the smaller one is its exact complement: the pair sums to 1
      const bool big0 = ay.w0 >= ay.w1;
      const float a = __bfloat162float(__float2bfloat16_rn(big0 ? ay.w0 : ay.w1));
      const float b = 1.0f - a;
      LineTap t;
      t.i0 = ay.i0; t.i1 = ay.i1;
      t.w0 = big0 ? a : b; t.w1 = big0 ? b : a;
      s_tap[(h * LIST_NUM_DISP + d) * kMaxLY + j] = t;
    }
    __syncthreads();
    // ---- geometry 2: K rows of every level: (displacement, H node) pairs in displacement order ----
    if (tid < p.nh * LIST_NUM_DISP) {
      const int h = tid / LIST_NUM_DISP, d = tid - h * LIST_NUM_DISP;
      const LineTap* tp = s_tap + (h * LIST_NUM_DISP + d) * kMaxLY;
      const int ymin = min(tp[0].i0, tp[p.ly - 1].i0), ymax = max(tp[0].i1, tp[p.ly - 1].i1);
      s_misc[40 + h * 8 + d] = ymax - ymin + 1;
      s_misc[h * 8 + d] = ymin << 8;
    }
    // A operands: zero (the H weights follow below)
    for (int i = tid; i < p.nh * A_BYTES / 16; i += kThreads)
      *reinterpret_cast<uint4*>(gbase + OFF_A + 16 * i) = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    if (tid < p.nh * LIST_NUM_DISP) {
      const int h = tid / LIST_NUM_DISP, d = tid - h * LIST_NUM_DISP, R = p.R[h];
      int k = 0;
      for (int dd = 0; dd < d; ++dd) k += s_misc[40 + h * 8 + dd];
      const int ny = s_misc[40 + h * 8 + d], ymin = s_misc[h * 8 + d] >> 8;
      if (k + ny > BK) __trap();                                     // host-side bound on ly violated
      const float qz = linspace_f32_step(static_cast<int>(lz), res, p.tm.bb_min, p.tm.bb_max, p.tm.step) * 2.0f;
      float q[3] = {0.f, 0.f, qz}, pd[3];
      displaced(q, d, pd);
      const Axis3 az = axis_border(pd[2], R);
      const uint32_t slab = static_cast<uint32_t>(d) * p.dstride[h];
      for (int y = 0; y < ny; ++y) {
        KRow r;
        r.off0 = slab + (static_cast<uint32_t>(az.i0) * R + (ymin + y)) * R * kN0;
        r.off1 = slab + (static_cast<uint32_t>(az.i1) * R + (ymin + y)) * R * kN0;
        r.w0 = az.w0; r.w1 = az.w1;
        s_krow[h * BK + k + y] = r;
      }
      if (d == LIST_NUM_DISP - 1) {
        const int kend = k + ny, kpad = (kend + 15) & ~15;
        s_misc[32 + h] = kpad;
        for (int kk = kend; kk < kpad; ++kk) s_krow[h * BK + kk] = KRow{0u, 0u, 0.f, 0.f};
      }
    }
    __syncthreads();
    for (int e = tid; e < p.nh * LIST_NUM_DISP * p.ly; e += kThreads) {
      const int j = e % p.ly, hd = e / p.ly, d = hd % LIST_NUM_DISP, h = hd / LIST_NUM_DISP;
      int kbase = 0;
      for (int dd = 0; dd < d; ++dd) kbase += s_misc[40 + h * 8 + dd];
      const int ymin = s_misc[h * 8 + d] >> 8;
      const LineTap t = s_tap[(h * LIST_NUM_DISP + d) * kMaxLY + j];
      const int m = shift_class(d) * p.ly + j;
      uint8_t* const arow = gbase + OFF_A + h * A_BYTES + m * 128;
      auto put = [&](int k, float w) {
        *reinterpret_cast<unsigned short*>(arow + ((((k >> 3) ^ (m & 7)) << 4) | ((k & 7) << 1))) = static_cast<unsigned short>(bf16_bits_rn(w));
      };
      put(kbase + t.i0 - ymin, t.w0);
      if (t.i1 != t.i0) put(kbase + t.i1 - ymin, t.w1);
    }
    fence_proxy_async_smem();
    __syncthreads();

    // level / node / half of tile t
    auto tile_of = [&](int t, int& h, int& node, int& half) {
      h = 0;
      int t0 = 0;
      while (h + 1 < p.nh && t >= t0 + 2 * p.R[h]) { t0 += 2 * p.R[h]; ++h; }
      node = (t - t0) >> 1;
      half = (t - t0) & 1;
    };

    if (warp == 0) {
      // =========================== MMA issuer ===========================
      if (lane == 0) {
        constexpr uint32_t idesc = umma_idesc(BM, BN) | (1u << 16);   // B operand MN-major
        for (int t = t_lo; t < t_hi; ++t, ++n_b, ++n_acc) {
          int h, node, half;
          tile_of(t, h, node, half);
          const int s = n_b % kStages, a = n_acc & 1;
          mbar_wait(acc_empty(a), ((n_acc >> 1) & 1) ^ 1);            // first use of each accumulator passes
          mbar_wait(b_full(s), (n_b / kStages) & 1);
          tc_fence_after();
          const int ks = s_misc[32 + h] >> 4;
          const uint64_t ad = umma_desc_sw128(base + OFF_A + h * A_BYTES);
          const uint64_t bd = umma_desc_mn_sw128(base + s * B_BYTES, 8192, 1024);
          for (int k = 0; k < ks; ++k)
            umma_ss<1>(tmem_base + a * BN, ad + 2 * k, bd + static_cast<uint64_t>((2048 >> 4) * k), idesc, k != 0 ? 1u : 0u);
          umma_commit<1>(b_empty(s));
          umma_commit<1>(acc_full(a));
        }
      } else {
        n_b += t_hi - t_lo; n_acc += t_hi - t_lo;
      }
    } else if (warp < kProdWarp0) {
      // =========================== epilogue warps ===========================
      const int quarter = warp & 3, colhalf = (warp - kEpiWarp0) >> 2;   // kEpiWarps / 4 warps per TMEM lane quarter
      const uint32_t tq = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
      const int m = quarter * 32 + lane;
      const int cls = m / p.ly, j = m - cls * p.ly;
      const bool live = cls < 3 && j >= jlo && j < jhi;
      const bool any_row = (quarter * 32) / p.ly < 3;                // some lane of this warp holds a table row
      // tensor stores: the item lies inside the launch's range and a warp's 32 rows are whole groups of ly lines
      const bool boxes = p.use_tma && jlo == 0 && jhi == p.ly && ly0 + p.ly <= res;
      const uint32_t stg = base + OFF_STG + static_cast<uint32_t>(warp - kEpiWarp0) * (2 * kStageBuf);
      for (int t = t_lo; t < t_hi; ++t, ++n_acc) {
        int h, node, half;
        tile_of(t, h, node, half);
        const int a = n_acc & 1;
        mbar_wait_warp(acc_full(a), (n_acc >> 1) & 1);
        tc_fence_after();
        if (any_row) {
          __nv_bfloat16* const dst = p.G + (static_cast<size_t>(out_line0 + j) * p.rpl + (p.rowbase[h] + cls * p.R[h] + node)) * kN0 + half * BN;
          uint32_t va[32], vb[32];
          // c: 32-column chunk of the tile.  Two chunks fill one staging buffer (64 channels).
          auto flush = [&](int c, const uint32_t (&v)[32]) {
            uint4 u[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              u[i].x = pack_bf16x2(__uint_as_float(v[8 * i + 0]), __uint_as_float(v[8 * i + 1]));
              u[i].y = pack_bf16x2(__uint_as_float(v[8 * i + 2]), __uint_as_float(v[8 * i + 3]));
              u[i].z = pack_bf16x2(__uint_as_float(v[8 * i + 4]), __uint_as_float(v[8 * i + 5]));
              u[i].w = pack_bf16x2(__uint_as_float(v[8 * i + 6]), __uint_as_float(v[8 * i + 7]));
            }
            if (boxes) {
              const uint32_t buf = stg + static_cast<uint32_t>((c >> 1) & 1) * kStageBuf;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint32_t piece = static_cast<uint32_t>((c & 1) * 4 + i);
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(buf + lane * 128u + ((piece ^ (lane & 7u)) << 4)),
                             "r"(u[i].x), "r"(u[i].y), "r"(u[i].z), "r"(u[i].w) : "memory");
              }
            } else {
              if (live) {
#pragma unroll
                for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(dst + c * 32 + 8 * i) = u[i];
              }
            }
          };
          // this warp's 64-channel groups of the tile (kEpiWarps / 4 warps share a TMEM lane quarter)
#pragma unroll 1
          for (int gi = 0; gi < kGroupsPerWarp; ++gi) {
            const int c = (colhalf * kGroupsPerWarp + gi) * 2;
            tmem_ld32_nowait(tq + a * BN + c * 32, va);
            tmem_ld32_nowait(tq + a * BN + (c + 1) * 32, vb);
            if (boxes) {                                               // the store that last read this staging buffer is two groups back
              bulk_wait_read<1>();                                     // every lane (bulk groups are per thread; only lane 0's are non-empty)
              __syncwarp();
            }
            tmem_ld_wait32(va);
            tmem_ld_wait32(vb);
            flush(c, va);
            flush(c + 1, vb);
            if (boxes) {
              const uint32_t buf = stg + static_cast<uint32_t>((c >> 1) & 1) * kStageBuf;
              fence_proxy_async_smem();
              __syncwarp();
              for (int sub = 0; sub * p.ly < 32; ++sub) {
                const int bc = (quarter * 32 + sub * p.ly) / p.ly;
                const uint32_t on = (lane == 0 && bc < 3) ? 1u : 0u;
                tma_store_4d(&tmG, buf + static_cast<uint32_t>(sub * p.ly) * 128u, half * BN + (c >> 1) * 64,
                             p.rowbase[h] + bc * p.R[h] + node, ly0, static_cast<int>(plane), on);
              }
              bulk_commit();
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_local(acc_empty(a));
      }
      bulk_wait_read<0>();                                             // the staging buffers are free for the next item
      __syncwarp();
      n_b += t_hi - t_lo;
    } else {
      // =========================== producer warps ===========================
      const int pw = warp - kProdWarp0;
      const int cg = lane >> 3, c16 = lane & 7;                       // 64-channel group, 16-byte piece inside its 128 B row
      for (int t = t_lo; t < t_hi; ++t, ++n_b) {
        int h, node, half;
        tile_of(t, h, node, half);
        const int s = n_b % kStages;
        mbar_wait_warp(b_empty(s), ((n_b / kStages) & 1) ^ 1);
        const int kpad = s_misc[32 + h];
        const __nv_bfloat16* __restrict__ src = p.pvol[h] + static_cast<uint32_t>(node) * kN0 + half * BN + lane * 8;
        uint8_t* const bst = gbase + s * B_BYTES + cg * 8192;
        // kRpp K rows per pass: their 2 * kRpp 16-byte loads are in flight together
        constexpr int kRpp = 6;
#pragma unroll 1
        for (int k0 = pw; k0 < kpad; k0 += kRpp * kProdWarps) {
          KRow r[kRpp];
          uint4 a0[kRpp], a1[kRpp];
#pragma unroll
          for (int u = 0; u < kRpp; ++u) {
            const int k = k0 + u * kProdWarps;
            if (k < kpad) {
              r[u] = s_krow[h * BK + k];
              a0[u] = __ldg(reinterpret_cast<const uint4*>(src + r[u].off0));
              a1[u] = __ldg(reinterpret_cast<const uint4*>(src + r[u].off1));
            }
          }
#pragma unroll
          for (int u = 0; u < kRpp; ++u) {
            const int k = k0 + u * kProdWarps;
            if (k < kpad) {
              const uint32_t w0[4] = {a0[u].x, a0[u].y, a0[u].z, a0[u].w}, w1[4] = {a1[u].x, a1[u].y, a1[u].z, a1[u].w};
              uint32_t o[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float lo = fmaf(__uint_as_float(w1[i] << 16), r[u].w1, __uint_as_float(w0[i] << 16) * r[u].w0);
                const float hi = fmaf(__uint_as_float(w1[i] & 0xffff0000u), r[u].w1, __uint_as_float(w0[i] & 0xffff0000u) * r[u].w0);
                o[i] = pack_bf16x2(lo, hi);
              }
              *reinterpret_cast<uint4*>(bst + (k >> 3) * 1024 + (k & 7) * 128 + ((c16 ^ (k & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_local(b_full(s));
      }
      n_acc += t_hi - t_lo;
    }
  }

  if (warp >= kEpiWarp0 && warp < kProdWarp0) bulk_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base);
  }
}

// Host side: see hoist.cuh.  Returns LIST_ENOSYS if the geometry does not fit the kernel (lines() then uses the SIMT kernel).
int lines_tc(const ListCtx* ctx, const Plan& pl, const void* buf, int image, int res, double bb_min, double bb_max, int64_t begin,
             int64_t count, void* G, cudaStream_t st) {
  using namespace ltc;
  if (count == 0 || pl.nh == 0) return LIST_OK;
  const char* basep = static_cast<const char*>(buf);
  LinesTcParams p{};
  p.nh = pl.nh;
  p.rpl = pl.rpl;
  int rmax = 2;
  p.nt_total = 0;
  for (int h = 0; h < kMaxLev; ++h) {
    const int hh = h < pl.nh ? h : 0;
    const size_t R = ctx->vol_res[pl.lev[hh]];
    if (static_cast<size_t>(LIST_NUM_DISP) * ctx->B * R * R * R * kN0 >= (1ull << 32)) return LIST_ENOSYS;   // 32-bit element offsets
    p.pvol[h] = reinterpret_cast<const __nv_bfloat16*>(basep + pl.off_pvol[hh]) + static_cast<size_t>(image) * R * R * R * kN0;
    p.dstride[h] = static_cast<uint32_t>(static_cast<size_t>(ctx->B) * R * R * R * kN0);
    p.R[h] = static_cast<int>(R);
    p.rowbase[h] = pl.rowbase[hh];
    if (h < pl.nh) {
      if (static_cast<int>(R) > rmax) rmax = static_cast<int>(R);
      p.nt_total += 2 * static_cast<int>(R);
    }
  }
  p.G = static_cast<__nv_bfloat16*>(G);
  fill_tilemap(&p.tm, res, bb_min, bb_max, begin, count, 128);
  p.line_first = begin / res;
  p.line_last = (begin + count - 1) / res;
  // lines per item: the largest power of two whose lines touch at most 9 H nodes per displacement (7 * 9 <= BK K rows)
  int ly = kMaxLY;
  while (ly > 1 && (res > 1 ? (static_cast<int64_t>(ly - 1) * (rmax - 1)) / (res - 1) : 0) + 3 > 9) ly >>= 1;
  if ((res > 1 ? (static_cast<int64_t>(ly - 1) * (rmax - 1)) / (res - 1) : 0) + 3 > 9) return LIST_ENOSYS;
  p.ly = ly;
  p.groups = (res + ly - 1) / ly;
  p.parts = 4;
  p.nt_part = (p.nt_total + p.parts - 1) / p.parts;
  const int64_t planes = p.line_last / res - p.line_first / res + 1;
  const int64_t items = planes * p.groups * p.parts;
  LIST_CHECK_ARG(items < (1LL << 31), "hoist::lines: too many lines for one launch");
  p.n_items = static_cast<unsigned>(items);
  int dev = 0, sms = 0;
  LIST_CUDA(cudaGetDevice(&dev));
  LIST_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static thread_local int attr_dev = -1;
  if (attr_dev != dev) {
    LIST_CUDA(cudaFuncSetAttribute(hoist_lines_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_dev = dev;
  }
  // G as [plane][line of the plane][row][512]; the origin is the first line of the launch's first plane, which may lie
  // before the buffer: only items inside [line_first, line_last] are stored through the map
  CUtensorMap tmG;
  p.use_tma = (ly >= 8) ? 1 : 0;
  {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return LIST_ENOSYS; }
    const size_t row_bytes = static_cast<size_t>(kN0) * 2, line_bytes = row_bytes * pl.rpl;
    char* origin = static_cast<char*>(G) - static_cast<size_t>(p.line_first % res) * line_bytes;
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(kN0), static_cast<cuuint64_t>(pl.rpl), static_cast<cuuint64_t>(res),
                                static_cast<cuuint64_t>(planes)};
    const cuuint64_t strides[3] = {row_bytes, line_bytes, line_bytes * res};
    const cuuint32_t box[4] = {64, 1, static_cast<cuuint32_t>(p.use_tma ? ly : 8), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(&tmG, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, origin, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (line tables) failed (CUresult %d)", static_cast<int>(r)); return LIST_ECUDA; }
  }
  const unsigned grid = static_cast<unsigned>(items < sms ? items : sms);
  hoist_lines_tc_kernel<<<grid, kThreads, SMEM_BYTES, st>>>(tmG, p);
  LIST_LAUNCH_CHECK("hoist_lines_tc_kernel");
  return LIST_OK;
}

}  // namespace hoist
}  // namespace list
