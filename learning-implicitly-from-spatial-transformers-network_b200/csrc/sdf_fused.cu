// Fused dense-grid SDF kernel (bf16 tensor-core mode): rows a-8 + a-2..a-6 of SURVEY.md §8 in ONE
// persistent kernel -- the feature rows never reach HBM.
//
//   reference: network/executors.py:215-224 (chunk loop) calling modules.py:24-54 (perceptual
//   pooling) and modules.py:255-282 (7-tap voxel sampling + implicit MLP).
//
// Per CTA (one of a cta_group::2 pair) a tile of 128 consecutive grid points:
//   * 8 GATHER warps walk the z-run exactly like gather_grid.cu (separable trilinear with register
//     cell caches, tables of per-row coordinates in shared memory) but write their 16-byte pieces
//     straight into the 128B-swizzled A-operand ring of fc_0 in shared memory, 64 feature columns
//     (one K chunk) at a time: group g of 64 threads owns ring stage g and produces chunks
//     g, g+4, g+8, ...; inside a group a thread owns one 8-column vector and a 16-row sub-run;
//   * the TMA warp streams only the WEIGHT tiles (W0/W1/W2 boxes, 128B swizzle) into the B ring;
//   * the MMA thread issues tcgen05.mma (cta_group::2, M=256, N=256, K=16) once the A stage
//     (mbarrier arrivals of the gather threads of BOTH CTAs after fence.proxy.async) and the B stage
//     (TMA transaction bytes) are full; tcgen05.commit frees both;
//   * fc_1 / fc_2 / fc_out and the epilogue warps are those of mlp_tc.cu (activations stay in TMEM).
// HBM traffic per query drops from 2 x 7.3 KB of feature rows to 4 B of SDF; the kernel is bound by
// the tensor pipe as long as the gather's instruction stream fits in the issue slots the MMA leaves
// idle (it needs ~1.4 k of the ~2.8 k warp-instruction slots per query).
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
constexpr int NA = 4;                           // A ring stages == gather groups
constexpr int NB = 4;                           // B ring stages
constexpr int kGatherWarp0 = 6;
#ifndef LIST_FUSED_GATHER_THREADS
#define LIST_FUSED_GATHER_THREADS 512
#endif
constexpr int kGatherThreads = LIST_FUSED_GATHER_THREADS;      // 256 or 512
constexpr int kGroupThreads = kGatherThreads / NA;             // threads filling one A stage
constexpr int kThreads = kGatherWarp0 * 32 + kGatherThreads;
constexpr int kSubRows = BM * 8 / kGroupThreads;               // rows one gather thread walks per chunk (16 or 8)
constexpr int kTabLevels = LIST_MAX_LEVELS;

struct AxEntry { int i0; float w1; };
struct UvEntry { int x0, y0; float w00, w01, w10, w11; };

constexpr int PARAM_FLOATS = N0 + N1 + N2 + N2;
constexpr int OFF_A = 0;
constexpr int OFF_B = OFF_A + NA * A_BYTES;
constexpr int OFF_PAR = OFF_B + NB * B_STAGE_BYTES;
constexpr int OFF_Q = OFF_PAR + PARAM_FLOATS * 4;                 // float [128][3]
constexpr int OFF_NEW = OFF_Q + BM * 3 * 4;                        // int   [128]
constexpr int OFF_UV = OFF_NEW + BM * 4;                           // UvEntry [128]
constexpr int OFF_AX = OFF_UV + BM * 24;                           // AxEntry [levels][3][128]
constexpr int OFF_BAR = OFF_AX + 6 * 3 * BM * 8;                   // 6 levels max in the tables
constexpr int NUM_BARS = 2 * NA + 2 * NB + 2;
constexpr int SMEM_BYTES = OFF_BAR + NUM_BARS * 8 + 16 + 1024;
constexpr int kMaxTabLevels = 6;

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
  int debug_skip;        // bit0: skip 2-D gather, bit1: skip 3-D vector gather, bit2: skip tail (timing experiments)
};

__device__ __forceinline__ int shift_class(int d) { return d == 1 ? 1 : (d == 2 ? 2 : 0); }

__device__ __forceinline__ void load_row8(const __nv_bfloat16* __restrict__ vol, const uint32_t base[4],
                                          const float wyz[4], int xv, int C, float out[8]) {
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

__global__ void __launch_bounds__(kThreads, 1)
sdf_fused_kernel(const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* const gbase = smem_raw + (base - raw);
  float* const s_par = reinterpret_cast<float*>(gbase + OFF_PAR);
  float* const s_b0 = s_par;
  float* const s_b1 = s_b0 + N0;
  float* const s_b2 = s_b1 + N1;
  float* const s_w3 = s_b2 + N2;
  float (*s_q)[3] = reinterpret_cast<float (*)[3]>(gbase + OFF_Q);
  int* const s_new = reinterpret_cast<int*>(gbase + OFF_NEW);
  UvEntry* const s_uv = reinterpret_cast<UvEntry*>(gbase + OFF_UV);
  AxEntry* const s_ax = reinterpret_cast<AxEntry*>(gbase + OFF_AX);     // [(l*3 + cls)*128 + row]
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
      mbar_init(afull_bar(s), kGroupThreads * CG);          // every gather thread of the stage's group, both CTAs
      mbar_init(aempty_bar(s), 1);
    }
    for (int s = 0; s < NB; ++s) {
      mbar_init(bfull_bar(s), 1);
      mbar_init(bempty_bar(s), 1);
    }
    mbar_init(dfull_bar, 1);
    mbar_init(hready_bar, 128 * CG);
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
          // A stage sa = kc % NA is produced by gather group sa; its n-th use (n counted across tiles)
          // is chunk kc of this tile: n = it * ceil((nk0 - sa) / NA) + kc / NA  (same count as `use` there)
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
      mbar_wait(dfull_bar, dphase); dphase ^= 1;
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
      mbar_arrive_cluster(hready_remote);
      mbar_wait(dfull_bar, dphase); dphase ^= 1;
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
      mbar_arrive_cluster(hready_remote);
      mbar_wait(dfull_bar, dphase); dphase ^= 1;
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
      mbar_arrive_cluster(hready_remote);
      if (row < p.count) p.sdf[row] = __fdiv_rn(acc + bias3, p.out_div);
    }
  } else {
    // =========================== gather warps ===========================
    const int gt = threadIdx.x - kGatherWarp0 * 32;          // 0..kGatherThreads-1
    const int grp = gt / kGroupThreads;                       // owns A stage `grp`
    const int tg = gt % kGroupThreads;
    const int j = tg & 7;                                     // 16-byte slot (8 columns) inside the chunk
    const int r0 = (tg >> 3) * kSubRows;                      // my sub-run of rows
    uint8_t* const my_stage = gbase + OFF_A + grp * A_BYTES;
    const uint32_t afull_remote = mapa(afull_bar(grp), 0);
    uint32_t use = 0;
    const int lim = p.S - 1;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const long long row_base = tile * rows_per_tile + rank * BM;          // first point of my CTA's tile
      named_bar_sync(1, kGatherThreads);                      // previous tile's tables no longer in use
      // ---- per-row tables (phase 0 of gather_grid.cu) ----
      if (gt < BM) {
        const long long n = row_base + gt;
        float q[3] = {0.f, 0.f, 0.f};
        int fresh = 1;
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

      for (int kc = grp; kc < nk0; kc += NA) {
        mbar_wait(aempty_bar(grp), (use & 1) ^ 1);
        ++use;
        const int col0 = kc * BK + j * 8;
        const int kind = col0 < p.map_off + p.Cm ? 0 : (col0 < p.tail0 ? 1 : 2);
        if ((p.debug_skip >> kind) & 1) {
          const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          for (int s = r0; s < r0 + kSubRows; ++s) sts_swizzled(my_stage, s, j, z);
        } else if (col0 < p.map_off + p.Cm) {
          // ---------------- 2-D vectors ----------------
          const __nv_bfloat16* __restrict__ maps = p.maps + (col0 - p.map_off);
          int cx = -1, cy = -1;
          float v00[8], v01[8], v10[8], v11[8];
#pragma unroll 1
          for (int s = r0; s < r0 + kSubRows; ++s) {
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
            for (int c = 0; c < 8; ++c)
              acc[c] = fmaf(v11[c], e.w11, fmaf(v10[c], e.w10, fmaf(v01[c], e.w01, v00[c] * e.w00)));
            sts_swizzled(my_stage, s, j, acc);
          }
        } else if (col0 < p.tail0) {
          // ---------------- 3-D vector level ----------------
          int l = 0;
          for (int ll = 0; ll < p.nlev; ++ll)
            if (!(p.C[ll] & 7) && col0 >= p.voff[ll] && col0 < p.voff[ll] + LIST_NUM_DISP * p.C[ll]) l = ll;
          const int R = p.R[l], C = p.C[l];
          const int rel = col0 - p.voff[l];
          const int d = rel / C, ch = rel - d * C;
          const __nv_bfloat16* __restrict__ vol = p.vols[l];
          const AxEntry* __restrict__ axs = s_ax + (l * 3 + shift_class(d)) * BM;
          uint32_t vb[4] = {0, 0, 0, 0};
          float wyz[4] = {0.f, 0.f, 0.f, 0.f};
          float G0[8], G1[8], D[8];
          int cx0 = -1;
#pragma unroll 1
          for (int s = r0; s < r0 + kSubRows; ++s) {
            const AxEntry e = axs[s];
            bool reload = false;
            if (s == r0 || s_new[s]) {
              const float q[3] = {s_q[s][0], s_q[s][1], s_q[s][2]};
              float pd[3];
              displaced(q, d, pd);
              const Axis3 ay = axis_border(pd[1], R), az = axis_border(pd[2], R);
              const int zi[2] = {az.i0, az.i1}, yi[2] = {ay.i0, ay.i1};
              const float wz[2] = {az.w0, az.w1}, wy[2] = {ay.w0, ay.w1};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int tz = k >> 1, ty = k & 1;
                vb[k] = (static_cast<uint32_t>(zi[tz]) * R + yi[ty]) * R * C + ch;
                wyz[k] = wy[ty] * wz[tz];
              }
              reload = true;
            }
            if (reload || e.i0 != cx0) {
              const int i1 = min(e.i0 + 1, R - 1);
              if (!reload && e.i0 == cx0 + 1) {
#pragma unroll
                for (int c = 0; c < 8; ++c) G0[c] = G1[c];
              } else {
                load_row8(vol, vb, wyz, e.i0, C, G0);
              }
              if (i1 != e.i0) load_row8(vol, vb, wyz, i1, C, G1);
              else {
#pragma unroll
                for (int c = 0; c < 8; ++c) G1[c] = G0[c];
              }
#pragma unroll
              for (int c = 0; c < 8; ++c) D[c] = G1[c] - G0[c];
              cx0 = e.i0;
            }
            float out[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) out[c] = fmaf(D[c], e.w1, G0[c]);
            sts_swizzled(my_stage, s, j, out);
          }
        } else {
          // ---------------- tail: the scalar (1-channel) level, q, zero padding ----------------
          // The 7 scalar columns are spread over the tail lanes of this octet (lanes of one sub-run):
          // tail lane number ti runs scalar walkers for displacements ti and ti + nt, the values are
          // exchanged with shuffles, and every tail lane assembles its own 8 columns.
          const int ls = p.scalar_level;
          const int R = p.R[ls];
          const __nv_bfloat16* __restrict__ vol = p.vols[ls];
          const int j0 = (p.tail0 - kc * BK) >> 3;               // first tail slot of this chunk (tail0 is inside it)
          const int nt = 8 - (j0 > 0 ? j0 : 0);                  // tail lanes per octet in this chunk
          const int ti = j - (j0 > 0 ? j0 : 0);
          const int sbase = (col0 - p.tail0) - ti * 8;           // scalar index of slot j0's first column (0 for the 1st tail chunk)
          const unsigned tail_mask = __ballot_sync(__activemask(), true);
          const int octet_lane0 = (lane & ~7) + (j0 > 0 ? j0 : 0);
          int dd[2] = {sbase + ti, sbase + ti + nt};             // my scalar columns == displacement ids
          uint32_t vb[2][4];
          float wyz[2][4], g0[2] = {0.f, 0.f}, g1[2] = {0.f, 0.f};
          int cx0[2] = {-1, -1};
#pragma unroll 1
          for (int s = r0; s < r0 + kSubRows; ++s) {
            float val[2] = {0.f, 0.f};
#pragma unroll
            for (int w = 0; w < 2; ++w) {
              const int d = dd[w];
              if (d < 0 || d >= LIST_NUM_DISP) continue;
              const AxEntry e = s_ax[(ls * 3 + shift_class(d)) * BM + s];
              bool reload = false;
              if (s == r0 || s_new[s]) {
                const float q[3] = {s_q[s][0], s_q[s][1], s_q[s][2]};
                float pd[3];
                displaced(q, d, pd);
                const Axis3 ay = axis_border(pd[1], R), az = axis_border(pd[2], R);
                const int zi[2] = {az.i0, az.i1}, yi[2] = {ay.i0, ay.i1};
                const float wz[2] = {az.w0, az.w1}, wy[2] = {ay.w0, ay.w1};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  vb[w][k] = (static_cast<uint32_t>(zi[k >> 1]) * R + yi[k & 1]) * R;
                  wyz[w][k] = wy[k & 1] * wz[k >> 1];
                }
                reload = true;
              }
              if (reload || e.i0 != cx0[w]) {
                const int i1 = min(e.i0 + 1, R - 1);
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  a0 = fmaf(to_f32(vol[vb[w][k] + e.i0]), wyz[w][k], a0);
                  a1 = fmaf(to_f32(vol[vb[w][k] + i1]), wyz[w][k], a1);
                }
                g0[w] = a0; g1[w] = a1;
                cx0[w] = e.i0;
              }
              val[w] = fmaf(g1[w] - g0[w], e.w1, g0[w]);
            }
            float out[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const int i = col0 - p.tail0 + c;                  // column index inside the tail region
              const int is = i < 0 ? 0 : i;
              const int src = octet_lane0 + (is - sbase >= 0 ? (is - sbase) % nt : 0);
              const float v0 = __shfl_sync(tail_mask, val[0], src);
              const float v1 = __shfl_sync(tail_mask, val[1], src);
              float v = 0.f;
              if (i >= 0 && i < LIST_NUM_DISP) v = ((i - sbase) / nt) ? v1 : v0;
              else if (i >= LIST_NUM_DISP && i < LIST_NUM_DISP + 3) v = s_q[s][i - LIST_NUM_DISP];
              out[c] = v;
            }
            sts_swizzled(my_stage, s, j, out);
          }
        }
        fence_proxy_async_smem();                 // generic-proxy writes -> visible to the tensor-core (async) proxy
        mbar_arrive_cluster(afull_remote);
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
  if (ctx->map_channels % 64 != 0) return LIST_ENOSYS;            // 2-D block must end on a chunk boundary
  FusedParams p{};
  p.maps = static_cast<const __nv_bfloat16*>(ctx->maps) + static_cast<size_t>(image) * ctx->map_size * ctx->map_size * ctx->map_channels;
  p.T = ctx->trans_mat + image * 12;
  p.S = ctx->map_size;
  p.Cm = ctx->map_channels;
  p.nlev = ctx->n_levels;
  int tail0 = lay.xyz_off;
  for (int l = 0; l < ctx->n_levels; ++l) {
    const size_t vox = static_cast<size_t>(ctx->vol_res[l]) * ctx->vol_res[l] * ctx->vol_res[l] * ctx->vol_ch[l];
    if (vox >= (1ull << 32)) return LIST_ENOSYS;
    p.vols[l] = static_cast<const __nv_bfloat16*>(ctx->vols[l]) + static_cast<size_t>(image) * vox;
    p.R[l] = ctx->vol_res[l];
    p.C[l] = ctx->vol_ch[l];
    p.voff[l] = lay.vol_off[l];
    if (ctx->vol_ch[l] % 8 != 0) tail0 = lay.vol_off[l] < tail0 ? lay.vol_off[l] : tail0;
  }
  // tail region = [one 1-channel level (7 columns) | q (3) | zero pad], starting on a 16-byte slot and
  // lying inside the last K chunk with >= 4 slots (so that each tail lane runs at most 2 scalar walkers)
  int n_scalar = 0;
  for (int l = 0; l < ctx->n_levels; ++l)
    if (ctx->vol_ch[l] % 8 != 0) { n_scalar += 1; p.scalar_level = l; if (ctx->vol_ch[l] != 1) return LIST_ENOSYS; }
  if (n_scalar != 1 || tail0 % 8 != 0 || lay.xyz_off != tail0 + LIST_NUM_DISP) return LIST_ENOSYS;
  if (tail0 / BK != (lay.k_pad - 1) / BK || (lay.k_pad - tail0) / 8 < 4) return LIST_ENOSYS;
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
    const char* e = getenv("LIST_B200_FUSED_SKIP");
    p.debug_skip = e ? atoi(e) : 0;
  }
  if (count == 0) return LIST_OK;
  if (count >= (1LL << 31) * 128) return LIST_ENOSYS;
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
