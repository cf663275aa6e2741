// Fused dense-grid SDF kernel (bf16 tensor-core mode): rows a-8 + a-2..a-6 of SURVEY.md §8 in ONE
// persistent kernel -- the feature rows never reach HBM.
//
//   reference: network/executors.py:215-224 (chunk loop) calling modules.py:24-54 (perceptual
//   pooling) and modules.py:255-282 (7-tap voxel sampling + implicit MLP).
//
// Per CTA (one of a cta_group::2 pair) a tile of 128 consecutive grid points:
//   * the TMA warp streams only the WEIGHT tiles (W0/W1/W2 boxes, 128B swizzle) into the B ring;
//   * 8 GATHER warps produce the A operand of fc_0 -- the 128 x 3648 feature tile -- 64 columns (one
//     K chunk) at a time, directly in the 128B-swizzled shared-memory layout the MMA reads.
//     Two groups of 128 threads alternate chunks (group g: chunks g, g+2, ...; A stage = chunk % 4).
//     A voxel chunk is produced in two phases, which is the separable trilinear evaluation of
//     gather_grid.cu turned into batched, independent loads (the dependent "walk" of that kernel is
//     latency-bound with only 8 warps per SM):
//        G phase : G[xv][slot] = sum_{4 (H,D) corners} (wy*wz) * V[z_k][y_k][xv][8 channels]   for every
//                  voxel column xv the tile's z-run touches (<= 68), one (xv, slot) item per thread,
//                  4 independent 16-byte loads each, written to an fp32 scratch table in smem;
//        L phase : out[row][slot] = G[x0] + w1 * (G[x0+1] - G[x0]),  one (row, slot) item per thread,
//                  packed to bf16 and stored into the A stage.
//     The loads of chunk n+1 are issued before the L phase of chunk n (register prefetch).
//     2-D chunks (the first 16) read 4 taps per (row, slot) in two batches of 4 rows.
//   * the MMA thread issues tcgen05.mma (cta_group::2, M=256, N=256, K=16) once the A stage
//     (mbarrier arrivals of the gather threads of BOTH CTAs after fence.proxy.async) and the B stage
//     (TMA transaction bytes) are full; tcgen05.commit frees both;
//   * fc_1 / fc_2 / fc_out and the epilogue warps are those of mlp_tc.cu (activations stay in TMEM).
// The arithmetic per feature is exactly that of gather_grid.cu, so fused == chunked bit for bit.
#include <cstdlib>

#include "tc_common.cuh"

namespace list {
namespace fused {

using namespace tc;

constexpr int CG = 2;
constexpr int BM = 128;
constexpr int BK = 64;
constexpr int N0 = 512, N1 = 256, N2 = 256;
constexpr int A_BYTES = BM * BK * 2;            // 16 KB
constexpr int SUB_BYTES = 128 * BK * 2;         // 16 KB
constexpr int SUBS_L0 = (N0 / 128) / CG;        // 2
constexpr int SUBS_L12 = (N1 / 128) / CG;       // 1
constexpr int B_STAGE_BYTES = SUBS_L0 * SUB_BYTES;
constexpr int NA = 4;                           // A ring stages (stage = chunk % 4)
#ifndef LIST_FUSED_NB
#define LIST_FUSED_NB 3
#endif
#ifndef LIST_FUSED_SCR
#define LIST_FUSED_SCR 68
#endif
constexpr int NB = LIST_FUSED_NB;               // B ring stages
constexpr int kGatherWarp0 = 6;
constexpr int kGroups = 2;
constexpr int kGroupThreads = 160;
constexpr int kGatherThreads = kGroups * kGroupThreads;            // 320
constexpr int kThreads = kGatherWarp0 * 32 + kGatherThreads;       // 512 = 16 warps, 4 per SM sub-partition
constexpr int kXr = kGroupThreads / 8;                             // 20 voxel-column residues / row groups
constexpr int kLRows = (BM + kXr - 1) / kXr;                       // 7 rows per thread in the L phase
constexpr int kScrRows = LIST_FUSED_SCR;        // G rows (voxel columns) one segment may touch
constexpr int SCR_BYTES = kScrRows * 8 * 32;    // [xv_rel][slot][8 floats] = 17408
constexpr int kMaxTabLevels = 6;
constexpr int kItems = 4;                       // G items per thread: xv_rel = xr + 20 i  (4 * 20 = 80 >= kScrRows)
constexpr int kHalf = (kLRows + 1) / 2;         // 2-D rows per batch (4, then 3)
static_assert(kItems * kXr >= kScrRows && kHalf <= kItems, "gather geometry");

struct AxEntry { int i0; float w1; };
struct UvEntry { int x0, y0; float w00, w01, w10, w11; };

constexpr int PARAM_FLOATS = N0 + N1 + N2 + N2;
constexpr int OFF_A = 0;
constexpr int OFF_B = OFF_A + NA * A_BYTES;
constexpr int OFF_PAR = OFF_B + NB * B_STAGE_BYTES;
constexpr int OFF_Q = OFF_PAR + PARAM_FLOATS * 4;                 // float [128][3]
constexpr int OFF_NEW = OFF_Q + BM * 3 * 4;                        // int   [128]
constexpr int OFF_UV = OFF_NEW + BM * 4;                           // UvEntry [128]
constexpr int OFF_AX = OFF_UV + BM * 24;                           // AxEntry [level][3][128]
constexpr int OFF_SEG = OFF_AX + kMaxTabLevels * 3 * BM * 8;       // int [132]: nseg, seg starts..., 128
constexpr int OFF_SCR = OFF_SEG + 544;
constexpr int OFF_BAR = OFF_SCR + kGroups * SCR_BYTES;
constexpr int NUM_BARS = 2 * NA + 2 * NB + 2;
constexpr int SMEM_BYTES = OFF_BAR + NUM_BARS * 8 + 16 + 1024;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
static_assert(OFF_SCR % 16 == 0 && OFF_BAR % 8 == 0, "alignment");

struct FusedParams {
  const __nv_bfloat16* maps;                 // image's [S][S][Cm]
  const __nv_bfloat16* vols[LIST_MAX_LEVELS];
  const float* T;                            // [12]
  int S, Cm;
  int R[LIST_MAX_LEVELS], C[LIST_MAX_LEVELS], voff[LIST_MAX_LEVELS];
  int nlev;
  int map_off, xyz_off, k_pad, tail0;
  int res;
  double bb_min, bb_max;
  long long grid_begin, count;
  const float *b0, *b1, *b2, *w3, *b3;
  float* sdf;
  float out_div;
  int nk0;
  int scalar_level;      // the one level with C == 1 (its 7 columns start the tail region)
  int seg_rows;          // max rows per segment so that a segment touches <= kScrRows voxel columns
  int debug_skip;        // bit0: skip 2-D gather, bit1: skip voxel gather, bit2: skip tail (timing experiments)
};

__device__ __forceinline__ int shift_class(int d) { return d == 1 ? 1 : (d == 2 ? 2 : 0); }

// 16 bytes (8 bf16) of row `row`, 16B-slot `j` of a [128][64] bf16 tile in the 128B-swizzle layout
// that TMA would have produced (slot index XOR (row mod 8)).
__device__ __forceinline__ void sts_swizzled(uint8_t* stage, int row, int j, const float v[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(stage + row * 128 + ((j ^ (row & 7)) << 4)) = u;
}
__device__ __forceinline__ void unpack8(const uint4& u, float v[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// What one 16-byte slot of a chunk holds, for one segment of rows [sa, sb).
enum SlotType { T_2D = 0, T_VEC = 1, T_TAIL = 2 };
struct Plan {
  int type;
  int kc, sa, sb;
  int l, d, R, C, ch;        // T_VEC
  int xlo, xhi;              // voxel columns the segment touches (T_VEC)
  float wyz[4];
  uint32_t base[4];
};

__global__ void __launch_bounds__(kThreads, 1)
sdf_fused_kernel(const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* const gbase = smem_raw + (base - raw_addr);
  float* const s_par = reinterpret_cast<float*>(gbase + OFF_PAR);
  float* const s_b0 = s_par;
  float* const s_b1 = s_b0 + N0;
  float* const s_b2 = s_b1 + N1;
  float* const s_w3 = s_b2 + N2;
  float (*s_q)[3] = reinterpret_cast<float (*)[3]>(gbase + OFF_Q);
  int* const s_new = reinterpret_cast<int*>(gbase + OFF_NEW);
  UvEntry* const s_uv = reinterpret_cast<UvEntry*>(gbase + OFF_UV);
  AxEntry* const s_ax = reinterpret_cast<AxEntry*>(gbase + OFF_AX);     // [(l*3 + cls)*128 + row]
  int* const s_seg = reinterpret_cast<int*>(gbase + OFF_SEG);           // [0] = nseg, [1..nseg] starts, [nseg+1] = 128
  const uint32_t bar0 = base + OFF_BAR;
  auto afull_bar = [&](int s) { return bar0 + 8u * s; };
  auto aempty_bar = [&](int s) { return bar0 + 8u * (NA + s); };
  auto bfull_bar = [&](int s) { return bar0 + 8u * (2 * NA + s); };
  auto bempty_bar = [&](int s) { return bar0 + 8u * (2 * NA + NB + s); };
  const uint32_t dfull_bar = bar0 + 8u * (2 * NA + 2 * NB);
  const uint32_t hready_bar = dfull_bar + 8u;
  const uint32_t tmem_slot = hready_bar + 8u;
  volatile uint32_t* const tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + OFF_BAR + NUM_BARS * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x / CG;
  const int num_clusters = gridDim.x / CG;
  const long long rows_per_tile = static_cast<long long>(BM) * CG;
  const int num_tiles = static_cast<int>((p.count + rows_per_tile - 1) / rows_per_tile);
  const int nk0 = p.nk0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW0);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    for (int s = 0; s < NA; ++s) {
      mbar_init(afull_bar(s), (kGroupThreads / 32) * CG);   // one arrival per gather warp of the producing group, both CTAs
      mbar_init(aempty_bar(s), 1);
    }
    for (int s = 0; s < NB; ++s) {
      mbar_init(bfull_bar(s), 1);
      mbar_init(bempty_bar(s), 1);
    }
    mbar_init(dfull_bar, 1);
    mbar_init(hready_bar, 4 * CG);                  // one arrival per epilogue warp, both CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<CG>(tmem_slot);
  if (warp >= 2 && warp < kGatherWarp0) {
    for (int i = threadIdx.x - 64; i < PARAM_FLOATS; i += 128) {
      float v;
      if (i < N0) v = __ldg(p.b0 + i);
      else if (i < N0 + N1) v = __ldg(p.b1 + i - N0);
      else if (i < N0 + N1 + N2) v = __ldg(p.b2 + i - N0 - N1);
      else v = __ldg(p.w3 + i - N0 - N1 - N2);
      s_par[i] = v;
    }
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  auto stage_a = [&](int s) { return base + OFF_A + static_cast<uint32_t>(s) * A_BYTES; };
  auto stage_b = [&](int s) { return base + OFF_B + static_cast<uint32_t>(s) * B_STAGE_BYTES; };

  if (warp == 0) {
    // =========================== TMA producer: weights only ===========================
    if (lane == 0) {
      uint32_t slot = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
#pragma unroll 1
        for (int layer = 0; layer <= 2; ++layer) {
          const CUtensorMap* tm = layer == 0 ? &tmW0 : (layer == 1 ? &tmW1 : &tmW2);
          const int nk = layer == 0 ? nk0 : (layer == 1 ? N0 / BK : N1 / BK);
          const int subs = layer == 0 ? SUBS_L0 : SUBS_L12;
          for (int kc = 0; kc < nk; ++kc, ++slot) {
            const int s = slot % NB;
            mbar_wait(bempty_bar(s), ((slot / NB) & 1) ^ 1);
            const uint32_t fb = mapa(bfull_bar(s), 0);
            if (rank == 0) mbar_expect_tx(bfull_bar(s), CG * subs * SUB_BYTES);
            for (int j = 0; j < subs; ++j) {
              const int wrow = (layer == 0 ? j * 256 : 0) + static_cast<int>(rank) * 128;
              tma_load_2d<CG>(tm, fb, stage_b(s) + j * SUB_BYTES, kc * BK, wrow);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA, one thread) ===========================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = umma_idesc(128 * CG, 256);
      constexpr uint32_t kInstrB = (256 / CG) * BK * 2;
      uint32_t bslot = 0, hphase = 0, it = 0;
      bool first = true;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
        if (!first) { mbar_wait(hready_bar, hphase); hphase ^= 1; }
        first = false;
        tc_fence_after();
        for (int kc = 0; kc < nk0; ++kc, ++bslot) {
          // A stage sa = kc % NA; its n-th use (counted across tiles) is chunk kc of this tile:
          // n = it * ceil((nk0 - sa) / NA) + kc / NA   (the gather side counts the same way)
          const int sa = kc % NA, sb = bslot % NB;
          const uint32_t ause = it * static_cast<uint32_t>((nk0 - sa + NA - 1) / NA) + static_cast<uint32_t>(kc / NA);
          mbar_wait(afull_bar(sa), ause & 1);
          mbar_wait(bfull_bar(sb), (bslot / NB) & 1);
          tc_fence_after();
          const uint64_t ad = umma_desc_sw128(stage_a(sa));
          const uint64_t bd = umma_desc_sw128(stage_b(sb));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
#pragma unroll
            for (int i = 0; i < 2; ++i)
              umma_ss<CG>(tmem_base + i * 256, ad + 2 * k, bd + ((i * kInstrB) >> 4) + 2 * k, idesc, (kc | k) != 0 ? 1u : 0u);
          }
          umma_commit<CG>(aempty_bar(sa));
          umma_commit<CG>(bempty_bar(sb));
        }
        umma_commit<CG>(dfull_bar);
#pragma unroll 1
        for (int layer = 1; layer <= 2; ++layer) {
          const int nk = (layer == 1 ? N0 : N1) / BK;
          mbar_wait(hready_bar, hphase); hphase ^= 1;
          tc_fence_after();
          for (int kc = 0; kc < nk; ++kc, ++bslot) {
            const int sb = bslot % NB;
            mbar_wait(bfull_bar(sb), (bslot / NB) & 1);
            tc_fence_after();
            const uint64_t bd = umma_desc_sw128(stage_b(sb));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_ts<CG>(tmem_base + 256, tmem_base + kc * (BK / 2) + k * 8, bd + 2 * k, idesc, (kc | k) != 0 ? 1u : 0u);
            umma_commit<CG>(bempty_bar(sb));
          }
          umma_commit<CG>(dfull_bar);
        }
      }
    }
  } else if (warp < kGatherWarp0) {
    // =========================== epilogue warps (as in mlp_tc.cu) ===========================
    const int quarter = warp & 3;
    const uint32_t tq = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t hready_remote = mapa(hready_bar, 0);
    const float bias3 = __ldg(p.b3);
    uint32_t dphase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const long long row = tile * rows_per_tile + rank * BM + quarter * 32 + lane;
      mbar_wait_warp(dfull_bar, dphase); dphase ^= 1;
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < N0 / 32; ++j) {
        uint32_t v[32], u[16];
        tmem_ld32(tq + j * 32, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 bb = *reinterpret_cast<const float2*>(s_b0 + j * 32 + 2 * i);
          u[i] = pack_bf16x2(fmaxf(__uint_as_float(v[2 * i]) + bb.x, 0.f), fmaxf(__uint_as_float(v[2 * i + 1]) + bb.y, 0.f));
        }
        tmem_st16(tq + j * 16, u);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(hready_remote);
      mbar_wait_warp(dfull_bar, dphase); dphase ^= 1;
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < N1 / 32; ++j) {
        uint32_t v[32], u[16];
        tmem_ld32(tq + 256 + j * 32, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 bb = *reinterpret_cast<const float2*>(s_b1 + j * 32 + 2 * i);
          u[i] = pack_bf16x2(fmaxf(__uint_as_float(v[2 * i]) + bb.x, 0.f), fmaxf(__uint_as_float(v[2 * i + 1]) + bb.y, 0.f));
        }
        tmem_st16(tq + j * 16, u);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(hready_remote);
      mbar_wait_warp(dfull_bar, dphase); dphase ^= 1;
      tc_fence_after();
      float acc = 0.f;
#pragma unroll 1
      for (int j = 0; j < N2 / 32; ++j) {
        uint32_t v[32];
        tmem_ld32(tq + 256 + j * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i)
          acc = fmaf(fmaxf(__uint_as_float(v[i]) + s_b2[j * 32 + i], 0.f), s_w3[j * 32 + i], acc);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(hready_remote);
      if (row < p.count) p.sdf[row] = __fdiv_rn(acc + bias3, p.out_div);
    }
  } else {
    // =========================== gather warps ===========================
    const int gt = threadIdx.x - kGatherWarp0 * 32;          // 0..255
    const int grp = gt / kGroupThreads;                       // chunk parity this group produces
    const int tg = gt % kGroupThreads;
    const int j = tg & 7;                                     // 16-byte slot (8 columns) of the chunk: G role and L role
    const int xr = tg >> 3;                                   // G role: voxel-column residue (0..19); L role: row group
    const int r0 = xr * kLRows;                               // L role: rows r0 .. r0+6 (clipped to the tile)
    float* const scr = reinterpret_cast<float*>(gbase + OFF_SCR + grp * SCR_BYTES);
    const int lim = p.S - 1;
    const int bar_id = 2 + grp;
    uint32_t it = 0;
    uint4 raw[kItems * 4];                                    // taps in flight (G phase / 2-D batch)

    // ---- slot plan for (chunk kc, rows [sa, sb)) ----
    auto make_plan = [&](Plan& P, int kc, int sa, int sb) {
      P.kc = kc; P.sa = sa; P.sb = sb;
      const int col0 = kc * BK + j * 8;
      P.l = 0; P.d = 0; P.R = 1; P.C = 8; P.ch = 0; P.xlo = 0; P.xhi = -1;
      if (col0 < p.map_off + p.Cm) { P.type = T_2D; return; }
      if (col0 >= p.tail0) { P.type = T_TAIL; return; }
      P.type = T_VEC;
      for (int ll = 0; ll < p.nlev; ++ll)
        if (!(p.C[ll] & 7) && col0 >= p.voff[ll] && col0 < p.voff[ll] + LIST_NUM_DISP * p.C[ll]) P.l = ll;
      P.R = p.R[P.l]; P.C = p.C[P.l];
      const int rel = col0 - p.voff[P.l];
      P.d = rel / P.C; P.ch = rel - P.d * P.C;
      const AxEntry* axs = s_ax + (P.l * 3 + shift_class(P.d)) * BM;
      P.xlo = axs[sa].i0;
      P.xhi = min(min(axs[sb - 1].i0 + 1, P.R - 1), P.xlo + kScrRows - 1);
      const float pd1 = s_q[sa][1] + (P.d == 3 ? -kDisplacement : (P.d == 4 ? kDisplacement : 0.f));
      const float pd2 = s_q[sa][2] + (P.d == 5 ? -kDisplacement : (P.d == 6 ? kDisplacement : 0.f));
      const Axis3 ay = axis_border(P.d == 3 || P.d == 4 ? pd1 : s_q[sa][1], P.R);
      const Axis3 az = axis_border(P.d == 5 || P.d == 6 ? pd2 : s_q[sa][2], P.R);
      const int zi[2] = {az.i0, az.i1}, yi[2] = {ay.i0, ay.i1};
      const float wz[2] = {az.w0, az.w1}, wy[2] = {ay.w0, ay.w1};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        P.base[k] = (static_cast<uint32_t>(zi[k >> 1]) * P.R + yi[k & 1]) * P.R * P.C + P.ch;
        P.wyz[k] = wy[k & 1] * wz[k >> 1];
      }
    };
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      const long long row_base = tile * rows_per_tile + rank * BM;          // first point of my CTA's tile
      named_bar_sync(1, kGatherThreads);                      // previous tile's tables no longer in use
      // ---- per-row tables (phase 0 of gather_grid.cu) ----
      if (gt < BM) {
        const long long n = row_base + gt;
        float q[3] = {0.f, 0.f, 0.f};
        int fresh = (n == p.count) ? 1 : 0;
        if (n < p.count) {
          const long long g = p.grid_begin + n;
          const int res = p.res;
          const int gz = static_cast<int>(g % res);
          const float rx = linspace_f32(static_cast<int>(g / (static_cast<long long>(res) * res)), res, p.bb_min, p.bb_max);
          const float ry = linspace_f32(static_cast<int>((g / res) % res), res, p.bb_min, p.bb_max);
          const float rz = linspace_f32(gz, res, p.bb_min, p.bb_max);
          q[0] = rz * 2.0f; q[1] = ry * 2.0f; q[2] = rx * 2.0f;
          fresh = (gz == 0) ? 1 : 0;
        }
        s_q[gt][0] = q[0]; s_q[gt][1] = q[1]; s_q[gt][2] = q[2];
        s_new[gt] = fresh;
        float ix, iy, h[3];
        localise(q, p.T, p.S, ix, iy, h);
        UvEntry e{0, 0, 0.f, 0.f, 0.f, 0.f};
        if (ix == ix && iy == iy) {
          const float fx = floorf(ix), fy = floorf(iy);
          e.x0 = static_cast<int>(fx); e.y0 = static_cast<int>(fy);
          const float wx1 = ix - fx, wx0 = (fx + 1.0f) - ix;
          const float wy1 = iy - fy, wy0 = (fy + 1.0f) - iy;
          const bool okx1 = (e.x0 + 1) <= lim, oky1 = (e.y0 + 1) <= lim;
          e.w00 = wx0 * wy0;
          e.w01 = okx1 ? wx1 * wy0 : 0.f;
          e.w10 = oky1 ? wx0 * wy1 : 0.f;
          e.w11 = (okx1 && oky1) ? wx1 * wy1 : 0.f;
        }
        s_uv[gt] = e;
      }
      for (int i = gt; i < p.nlev * 3 * BM; i += kGatherThreads) {
        const int s = i % BM, cls = (i / BM) % 3, l = i / (3 * BM);
        const long long n = row_base + s;
        float q0 = 0.f;
        if (n < p.count) q0 = linspace_f32(static_cast<int>((p.grid_begin + n) % p.res), p.res, p.bb_min, p.bb_max) * 2.0f;
        const float c = cls == 0 ? q0 : q0 + (cls == 1 ? -kDisplacement : kDisplacement);
        const Axis3 ax = axis_border(c, p.R[l]);
        s_ax[(l * 3 + cls) * BM + s] = AxEntry{ax.i0, ax.w1};
      }
      named_bar_sync(1, kGatherThreads);
      if (gt == 0) {                                          // segments: constant (H,D) corners, bounded x span
        int nseg = 0, start = 0;
        for (int s = 1; s < BM; ++s)
          if (s_new[s] || s - start >= p.seg_rows) { s_seg[1 + nseg++] = start; start = s; }
        s_seg[1 + nseg++] = start;
        s_seg[1 + nseg] = BM;
        s_seg[0] = nseg;
      }
      named_bar_sync(1, kGatherThreads);
      const int nseg = s_seg[0];

      // ---- units = (chunk of this group) x (segment of rows) ----
      Plan cur;
      int kc = grp, sg = 0;
      while (kc < nk0) {
        make_plan(cur, kc, s_seg[1 + sg], s_seg[2 + sg]);
        const int sa_stage = kc % NA;
        uint8_t* const stage = gbase + OFF_A + sa_stage * A_BYTES;
        const uint32_t ause = it * static_cast<uint32_t>((nk0 - sa_stage + NA - 1) / NA) + static_cast<uint32_t>(kc / NA);
        const bool first_seg = (sg == 0), last_seg = (cur.type == T_2D) || (sg == nseg - 1);
        const bool skip = (p.debug_skip >> cur.type) & 1;

        if (cur.type == T_2D) {
          mbar_wait_warp(aempty_bar(sa_stage), (ause & 1) ^ 1);
          if (skip) {
            const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int r = r0; r < min(r0 + kLRows, BM); ++r) sts_swizzled(stage, r, j, z);
          } else {
            const __nv_bfloat16* __restrict__ maps = p.maps + (kc * BK + j * 8 - p.map_off);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int rbase = r0 + half * kHalf;
              const int nrow = half == 0 ? kHalf : kLRows - kHalf;
#pragma unroll
              for (int ri = 0; ri < kHalf; ++ri) {
                if (ri < nrow) {
                  const UvEntry e = s_uv[min(rbase + ri, BM - 1)];
                  const int x1 = min(e.x0 + 1, lim), y1 = min(e.y0 + 1, lim);
                  raw[ri * 4 + 0] = ldg16(maps + (static_cast<size_t>(e.y0) * p.S + e.x0) * p.Cm);
                  raw[ri * 4 + 1] = ldg16(maps + (static_cast<size_t>(e.y0) * p.S + x1) * p.Cm);
                  raw[ri * 4 + 2] = ldg16(maps + (static_cast<size_t>(y1) * p.S + e.x0) * p.Cm);
                  raw[ri * 4 + 3] = ldg16(maps + (static_cast<size_t>(y1) * p.S + x1) * p.Cm);
                }
              }
#pragma unroll
              for (int ri = 0; ri < kHalf; ++ri) {
                if (ri < nrow && rbase + ri < BM) {
                  const UvEntry e = s_uv[rbase + ri];
                  float v00[8], v01[8], v10[8], v11[8], acc[8];
                  unpack8(raw[ri * 4 + 0], v00); unpack8(raw[ri * 4 + 1], v01);
                  unpack8(raw[ri * 4 + 2], v10); unpack8(raw[ri * 4 + 3], v11);
#pragma unroll
                  for (int c = 0; c < 8; ++c)
                    acc[c] = fmaf(v11[c], e.w11, fmaf(v10[c], e.w10, fmaf(v01[c], e.w01, v00[c] * e.w00)));
                  sts_swizzled(stage, rbase + ri, j, acc);
                }
              }
            }
          }
          fence_proxy_async_smem();                             // every lane: its generic-proxy writes -> async proxy
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa(afull_bar(sa_stage), 0));
        } else {
          // ---------------- G phase: 4 independent taps per (voxel column, slot) item -> scratch table ----------------
          if (cur.type == T_VEC && !skip) {
            const __nv_bfloat16* __restrict__ vol = p.vols[cur.l];
#pragma unroll
            for (int i = 0; i < kItems; ++i) {
              const int xv = cur.xlo + xr + kXr * i;
              if (xv <= cur.xhi) {
#pragma unroll
                for (int k = 0; k < 4; ++k) raw[i * 4 + k] = ldg16(vol + cur.base[k] + static_cast<uint32_t>(xv) * cur.C);
              }
            }
#pragma unroll
            for (int i = 0; i < kItems; ++i) {
              const int xrel = xr + kXr * i;
              if (cur.xlo + xrel <= cur.xhi) {
                float v[8], g[8];
                unpack8(raw[i * 4 + 0], v);
#pragma unroll
                for (int c = 0; c < 8; ++c) g[c] = v[c] * cur.wyz[0];
#pragma unroll
                for (int k = 1; k < 4; ++k) {
                  unpack8(raw[i * 4 + k], v);
#pragma unroll
                  for (int c = 0; c < 8; ++c) g[c] = fmaf(v[c], cur.wyz[k], g[c]);
                }
                float4* dst = reinterpret_cast<float4*>(scr + (xrel * 8 + j) * 8);
                dst[0] = make_float4(g[0], g[1], g[2], g[3]);
                dst[1] = make_float4(g[4], g[5], g[6], g[7]);
              }
            }
          }
          const int ls = p.scalar_level;
          const int j0 = (p.tail0 - cur.kc * BK) >> 3;          // first tail slot of the last chunk
          if (cur.type == T_TAIL && !skip && j >= j0 && j <= j0 + 2) {
            // scalar G slots: j0 -> d in {0,3,4,5,6} (no W shift), j0+1 -> d=1, j0+2 -> d=2
            const int which = j - j0;
            const int Rs = p.R[ls];
            const __nv_bfloat16* __restrict__ vol = p.vols[ls];
            const AxEntry* axs = s_ax + (ls * 3 + which) * BM;
            const int xlo = axs[cur.sa].i0;
            const int xhi = min(min(axs[cur.sb - 1].i0 + 1, Rs - 1), xlo + kScrRows - 1);
            const int nd = which == 0 ? 5 : 1;
            for (int xrel = xr; xlo + xrel <= xhi; xrel += kXr) {
              float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int t = 0; t < 5; ++t) {
                if (t < nd) {
                  const int d = which == 0 ? (t == 0 ? 0 : t + 2) : which;
                  const float q1 = s_q[cur.sa][1], q2 = s_q[cur.sa][2];
                  const Axis3 ay = axis_border(d == 3 ? q1 - kDisplacement : (d == 4 ? q1 + kDisplacement : q1), Rs);
                  const Axis3 az = axis_border(d == 5 ? q2 - kDisplacement : (d == 6 ? q2 + kDisplacement : q2), Rs);
                  const int zi[2] = {az.i0, az.i1}, yi[2] = {ay.i0, ay.i1};
                  const float wz[2] = {az.w0, az.w1}, wy[2] = {ay.w0, ay.w1};
                  float a = 0.f;
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const uint32_t b = (static_cast<uint32_t>(zi[k >> 1]) * Rs + yi[k & 1]) * Rs;
                    a = fmaf(to_f32(vol[b + xlo + xrel]), wy[k & 1] * wz[k >> 1], a);
                  }
                  g[t] = a;
                }
              }
              float4* dst = reinterpret_cast<float4*>(scr + (xrel * 8 + j) * 8);
              dst[0] = make_float4(g[0], g[1], g[2], g[3]);
              dst[1] = make_float4(g[4], g[5], g[6], g[7]);
            }
          }
          named_bar_sync(bar_id, kGroupThreads);                // scratch table complete
          if (first_seg) mbar_wait_warp(aempty_bar(sa_stage), (ause & 1) ^ 1);
          // ---------------- L phase: rows of this segment ----------------
          const int ra = max(r0, cur.sa), rb = min(r0 + kLRows, cur.sb);
          if (skip) {
            const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int r = ra; r < rb; ++r) sts_swizzled(stage, r, j, z);
          } else if (cur.type == T_VEC) {
            const AxEntry* axs = s_ax + (cur.l * 3 + shift_class(cur.d)) * BM;
            for (int r = ra; r < rb; ++r) {
              const AxEntry e = axs[r];
              const int x0r = min(max(e.i0 - cur.xlo, 0), kScrRows - 1);
              const int x1r = min(max(min(e.i0 + 1, cur.R - 1) - cur.xlo, 0), kScrRows - 1);
              const float4* g0p = reinterpret_cast<const float4*>(scr + (x0r * 8 + j) * 8);
              const float4* g1p = reinterpret_cast<const float4*>(scr + (x1r * 8 + j) * 8);
              const float4 a0 = g0p[0], a1 = g0p[1], b0 = g1p[0], b1 = g1p[1];
              const float G0[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
              const float G1[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              float out[8];
#pragma unroll
              for (int c = 0; c < 8; ++c) out[c] = fmaf(G1[c] - G0[c], e.w1, G0[c]);
              sts_swizzled(stage, r, j, out);
            }
          } else {                                              // T_TAIL slots: [7 scalars | q | zeros]
            const int Rs = p.R[ls];
            int xlo3[3];
#pragma unroll
            for (int c3 = 0; c3 < 3; ++c3) xlo3[c3] = s_ax[(ls * 3 + c3) * BM + cur.sa].i0;
            for (int r = ra; r < rb; ++r) {
              float out[8];
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const int i = cur.kc * BK + j * 8 + c - p.tail0;  // column inside the tail region
                float v = 0.f;
                if (i < LIST_NUM_DISP) {
                  const int d = i;
                  const int cls = shift_class(d);
                  const int gslot = j0 + cls, pos = cls == 0 ? (d == 0 ? 0 : d - 2) : 0;
                  const AxEntry e = s_ax[(ls * 3 + cls) * BM + r];
                  const int x0r = min(max(e.i0 - xlo3[cls], 0), kScrRows - 1);
                  const int x1r = min(max(min(e.i0 + 1, Rs - 1) - xlo3[cls], 0), kScrRows - 1);
                  const float g0 = scr[(x0r * 8 + gslot) * 8 + pos], g1 = scr[(x1r * 8 + gslot) * 8 + pos];
                  v = fmaf(g1 - g0, e.w1, g0);
                } else if (i < LIST_NUM_DISP + 3) {
                  v = s_q[r][i - LIST_NUM_DISP];
                }
                out[c] = v;
              }
              sts_swizzled(stage, r, j, out);
            }
          }
          if (last_seg) {
            fence_proxy_async_smem();                           // generic-proxy writes -> visible to the async (tensor) proxy
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa(afull_bar(sa_stage), 0));
          }
          named_bar_sync(bar_id, kGroupThreads);                // scratch table free again
        }
        if (cur.type == T_2D || sg + 1 >= nseg) { kc += kGroups; sg = 0; } else { ++sg; }
      }
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

}  // namespace fused

// One launch evaluates grid points [begin, begin+count) of image `image` end to end.
// Returns LIST_ENOSYS for configurations the fused mapping does not cover (caller falls back to
// gather + MLP kernels).
int sdf_grid_fused(const ListCtx* ctx, const ListWeights* w, int image, int res, double bb_min, double bb_max,
                   int64_t begin, int64_t count, float* sdf, float out_div, cudaStream_t st) {
  using namespace fused;
  if (ctx->dtype != LIST_BF16 || w->dtype != LIST_BF16) return LIST_ENOSYS;
  if (w->n0 != N0 || w->n1 != N1 || w->n2 != N2) return LIST_ENOSYS;
  ListLayout lay;
  const int rc0 = list_feature_layout(ctx->map_channels, ctx->n_levels, ctx->vol_ch, &lay, nullptr);
  if (rc0) return rc0;
  if (w->k_pad != lay.k_pad || ctx->n_levels > kMaxTabLevels) return LIST_ENOSYS;
  if (ctx->map_channels % 64 != 0 || lay.map_off != 0) return LIST_ENOSYS;   // 2-D block = whole chunks
  FusedParams p{};
  p.maps = static_cast<const __nv_bfloat16*>(ctx->maps) + static_cast<size_t>(image) * ctx->map_size * ctx->map_size * ctx->map_channels;
  p.T = ctx->trans_mat + image * 12;
  p.S = ctx->map_size;
  p.Cm = ctx->map_channels;
  p.nlev = ctx->n_levels;
  int tail0 = lay.xyz_off;
  double rho = 0.0;
  for (int l = 0; l < ctx->n_levels; ++l) {
    const size_t vox = static_cast<size_t>(ctx->vol_res[l]) * ctx->vol_res[l] * ctx->vol_res[l] * ctx->vol_ch[l];
    if (vox >= (1ull << 32)) return LIST_ENOSYS;
    p.vols[l] = static_cast<const __nv_bfloat16*>(ctx->vols[l]) + static_cast<size_t>(image) * vox;
    p.R[l] = ctx->vol_res[l];
    p.C[l] = ctx->vol_ch[l];
    p.voff[l] = lay.vol_off[l];
    if (ctx->vol_ch[l] % 8 != 0) tail0 = lay.vol_off[l] < tail0 ? lay.vol_off[l] : tail0;
    const double r = res > 1 ? static_cast<double>(ctx->vol_res[l] - 1) / (res - 1) : 0.0;
    rho = r > rho ? r : rho;
  }
  // tail region = [one 1-channel level (7 columns) | q (3) | zero pad], starting on a 16-byte slot and
  // lying inside the last K chunk with >= 3 slots left for the scalar G tables
  int n_scalar = 0;
  for (int l = 0; l < ctx->n_levels; ++l)
    if (ctx->vol_ch[l] % 8 != 0) { n_scalar += 1; p.scalar_level = l; if (ctx->vol_ch[l] != 1) return LIST_ENOSYS; }
  if (n_scalar != 1 || tail0 % 8 != 0 || lay.xyz_off != tail0 + LIST_NUM_DISP) return LIST_ENOSYS;
  if (tail0 / BK != (lay.k_pad - 1) / BK || (lay.k_pad - tail0) / 8 < 3) return LIST_ENOSYS;
  p.map_off = lay.map_off;
  p.xyz_off = lay.xyz_off;
  p.k_pad = lay.k_pad;
  p.tail0 = tail0;
  p.res = res;
  p.bb_min = bb_min;
  p.bb_max = bb_max;
  p.grid_begin = begin;
  p.count = count;
  p.b0 = w->b0; p.b1 = w->b1; p.b2 = w->b2; p.w3 = w->w3; p.b3 = w->b3;
  p.sdf = sdf;
  p.out_div = out_div;
  p.nk0 = lay.k_pad / BK;
  {
    // rows per segment so that floor((L-1)*rho) + 3 <= kScrRows - 2 voxel columns are touched
    int L = BM;
    if (rho > 0.0) {
      const double lmax = (kScrRows - 5) / rho + 1.0;
      L = lmax < BM ? static_cast<int>(lmax) : BM;
    }
    p.seg_rows = L < 1 ? 1 : L;
    const char* e = getenv("LIST_B200_FUSED_SKIP");
    p.debug_skip = e ? atoi(e) : 0;
  }
  if (count == 0) return LIST_OK;
  CUtensorMap tmW0, tmW1, tmW2;
  int rc;
  if ((rc = tc::make_map_bf16(&tmW0, w->w0, w->k_pad, N0, w->k_pad))) return rc;
  if ((rc = tc::make_map_bf16(&tmW1, w->w1, N0, N1, N0))) return rc;
  if ((rc = tc::make_map_bf16(&tmW2, w->w2, N1, N2, N1))) return rc;
  int dev = 0, sms = 0;
  LIST_CUDA(cudaGetDevice(&dev));
  LIST_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  LIST_CUDA(cudaFuncSetAttribute(sdf_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  const int64_t tiles = (count + BM * CG - 1) / (BM * CG);
  const int clusters = static_cast<int>(tiles < (sms / CG) ? tiles : (sms / CG));
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
  LIST_CUDA(cudaLaunchKernelEx(&cfg, sdf_fused_kernel, tmW0, tmW1, tmW2, p));
  return LIST_OK;
}

}  // namespace list
