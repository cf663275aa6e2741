// bf16 tensor-core implicit MLP (row a-6, reference network/modules.py:196-201, 276-282):
//   sdf = fc_out(relu(fc_2(relu(fc_1(relu(fc_0 x))))))      3610(->3648) -> 512 -> 256 -> 256 -> 1
// as ONE persistent warp-specialised sm_100a kernel.  Per CTA a 128-row tile of X runs through
// all four layers without leaving the SM:
//
//   TMA (cp.async.bulk.tensor, 128B swizzle) streams 16 KB boxes (128 rows x 64 columns of X, of a weight matrix, or
//   of the hoisted rows' addend block) into a shared-memory ring of 12 units managed as a chunk FIFO (see the kernel);
//   tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) accumulates in TMEM;
//   fc_0's 128x512 fp32 accumulator fills all 512 TMEM columns;  the epilogue warps read it with
//   tcgen05.ld, add bias (or, on hoisted rows, the addend block read from the ring), ReLU, round to bf16 and write it
//   BACK to TMEM (tcgen05.st) as the A operand of fc_1 (A-from-TMEM MMA), likewise for fc_2;  fc_out (256 -> 1) is a
//   register dot product in the last epilogue.  The feature concat was already done by the gather kernel's
//   row layout, biases / activations / the final /sdf_scale are fused here.
//
//   TMEM columns : fc_0 acc [0,512) -> H1 bf16 [0,256) -> fc_1 acc [256,512) -> H2 bf16 [0,128)
//                  -> fc_2 acc [256,512).
//   Warp roles   : warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one thread),
//                  warps 2..5 = epilogue (one TMEM lane quarter each).
//   CG = 2       : a CTA pair (cluster 2x1x1) runs tcgen05.mma.cta_group::2 (M = 256): each CTA
//                  loads its own 128 rows of X and HALF of every weight tile, which halves the
//                  L2 -> SM weight traffic per row, since W0 is re-streamed for every row tile.
//   Measured per 128-row tile on hoisted rows (K = 832; list_mlp_hoisted_trace, profiles/r01_mlp_phase_trace.txt):
//   fc_0 7.5 us | ep0 3.9 | fc_1 2.6 | ep1 1.3 | fc_2 1.5 | ep2 1.2 | three hand-offs 2.1  = 20.0 us, 58 % of it MMA.
#include "tc_common.cuh"

namespace list {
namespace tc {

constexpr int BM = 128;                 // rows per CTA
constexpr int BK = 64;                  // bf16 elements per K chunk = 128 B = one swizzle atom row
constexpr int N0 = 512, N1 = 256, N2 = 256;
constexpr int UNIT_BYTES = 128 * BK * 2; // one 128-row x 64-column box (X, weight or addend), 16 KB
constexpr int kThreads = 192;
constexpr int kEpiWarp0 = 2;
constexpr int NA = N0 / BK;             // addend boxes per tile (hoisted rows)

template <int CG, int ST>
struct Cfg {
  static constexpr int SUBS_L0 = (N0 / 128) / CG;     // weight boxes per CTA per fc_0 chunk
  static constexpr int SUBS_L12 = (N1 / 128) / CG;    // per fc_1 / fc_2 chunk
  static constexpr int U0 = 1 + SUBS_L0;              // units of one fc_0 chunk (X box + weight boxes)
  static constexpr int NU = ST * U0;                  // ring units (ST fc_0 chunks deep)
  static constexpr int NB = NU;                       // chunk barriers (a chunk has >= 1 unit)
  static constexpr int RING_BYTES = NU * UNIT_BYTES;
  static constexpr int PARAM_FLOATS = N0 + N1 + N2 + N2;   // b0 b1 b2 w3
  static constexpr int BAR_OFF = RING_BYTES + PARAM_FLOATS * 4;
  static constexpr int NUM_BARS = 2 * NB + NA + 2;    // full[NB] empty[NB] afull[NA] dfull hready
  static constexpr int SMEM_BYTES = BAR_OFF + NUM_BARS * 8 + 16 + 1024 /*align slack*/;
};

// ------------------------------------------------------------------ kernel
// Operand ring.  Shared memory holds NU units of 16 KB; operands travel as CHUNKS of 1..U0 consecutive units
// (wrapping), allocated first-in first-out by the TMA producer:
//     fc_0 chunk kc : [X box | weight boxes]                      -> consumed by the MMA thread
//     addend box a  : 64 columns of the hoisted rows' addend block -> consumed by the epilogue warps (hoisted rows only)
//     fc_1 / fc_2   : weight boxes                                 -> consumed by the MMA thread
// in exactly that order per tile.  Chunk number q (counted over the whole kernel) owns barriers full[q % NB] (leader
// CTA; MMA-consumed chunks only), afull[a] (local; addend boxes) and empty[q % NB] (local in every CTA: completed by the
// multicast tcgen05.commit for MMA-consumed chunks, by an elected epilogue thread for addend boxes), so the empty phase
// of chunk q is q / NB everywhere.  Because the addend boxes and ALL of W1 / W2 fit in the ring at once (8 + 8 + 4 units
// of 12), the producer fetches them while fc_0 and the first epilogue run: the epilogue reads the addend from shared
// memory and fc_1 / fc_2 never wait for a weight tile.
template <int CG, int ST>
__global__ void __launch_bounds__(kThreads, 1)
mlp_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW0,
              const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
              const float* __restrict__ b0, const float* __restrict__ b1, const float* __restrict__ b2,
              const float* __restrict__ w3, const float* __restrict__ b3, float* __restrict__ sdf,
              long long rows, int nk0, float out_div, float* __restrict__ dbg1, float* __restrict__ dbg2,
              float* __restrict__ dbg3, __nv_bfloat16* __restrict__ proj_out, int proj_groups, int proj_w_col_stride,
              long long proj_out_group_stride, int xcol0, int has_add, long long* __restrict__ trace) {
  using C = Cfg<CG, ST>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;               // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* const gbase = smem_raw + (base - raw);
  float* const s_par = reinterpret_cast<float*>(gbase + C::RING_BYTES);
  float* const s_b0 = s_par;
  float* const s_b1 = s_b0 + N0;
  float* const s_b2 = s_b1 + N1;
  float* const s_w3 = s_b2 + N2;
  const uint32_t bar0 = base + C::BAR_OFF;
  auto full_bar = [&](uint32_t b) { return bar0 + 8u * b; };
  auto empty_bar = [&](uint32_t b) { return bar0 + 8u * (C::NB + b); };
  auto afull_bar = [&](uint32_t a) { return bar0 + 8u * (2 * C::NB + a); };
  const uint32_t dfull_bar = bar0 + 8u * (2 * C::NB + NA);
  const uint32_t hready_bar = dfull_bar + 8u;
  const uint32_t tmem_slot = hready_bar + 8u;
  volatile uint32_t* const tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + C::BAR_OFF + C::NUM_BARS * 8);
  auto unit_addr = [&](uint32_t u) { return base + (u % C::NU) * UNIT_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int cluster_id = blockIdx.x / CG;
  const int num_clusters = gridDim.x / CG;
  const long long rows_per_tile = static_cast<long long>(BM) * CG;
  // Projection mode (proj_out != nullptr, hoist.cu): only fc_0 runs, without bias/ReLU, and the raw
  // 128 x 512 accumulator is written as bf16 rows.  The tile index then also enumerates `proj_groups`
  // column blocks of W0 (block g starts at column g * proj_w_col_stride): every group multiplies the same
  // X rows and writes its own [rows][512] output slab.
  const bool proj = proj_out != nullptr;
  const bool add = has_add != 0;          // hoisted rows: columns [0, N0) of the X map are the addend block, fc_0's K starts at xcol0
  const int tiles_per_group = static_cast<int>((rows + rows_per_tile - 1) / rows_per_tile);
  const int num_tiles = proj ? tiles_per_group * proj_groups : tiles_per_group;
  // chunk sequence of one tile
  const int n_add = add ? NA : 0;
  const int n_l12 = proj ? 0 : (N0 + N1) / BK;                       // 8 fc_1 + 4 fc_2 weight chunks
  const int chunks_per_tile = nk0 + n_add + n_l12;
  const int units_per_tile = nk0 * C::U0 + n_add + n_l12 * C::SUBS_L12;
  auto units_of = [&](int i) { return i < nk0 ? C::U0 : (i < nk0 + n_add ? 1 : C::SUBS_L12); };

  // Diagnostic phase timeline (list_mlp_hoisted_trace): CTA 0 stamps clock64() at the phase boundaries of its first
  // kTraceTiles tiles: slots 0..5 by the MMA thread (tile start, fc_0 issued, fc_1 start, fc_1 issued, fc_2 start,
  // fc_2 issued), slots 6..11 by epilogue warp 2 (fc_0 done seen, ep0 done, fc_1 done seen, ep1 done, fc_2 done seen, ep2 done).
  constexpr int kTraceTiles = 16, kTraceSlots = 12;
  const bool tracing = trace != nullptr && blockIdx.x == 0;
  auto stamp = [&](int tile_no, int slot) {
    if (tracing && tile_no < kTraceTiles) trace[tile_no * kTraceSlots + slot] = clock64();
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW0);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    for (int b = 0; b < C::NB; ++b) {
      mbar_init(full_bar(b), 1);
      mbar_init(empty_bar(b), 1);
    }
    for (int a = 0; a < NA; ++a) mbar_init(afull_bar(a), 1);
    mbar_init(dfull_bar, 1);
    mbar_init(hready_bar, 4 * CG);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<CG>(tmem_slot);
  if (warp >= kEpiWarp0) {
    for (int i = threadIdx.x - kEpiWarp0 * 32; i < C::PARAM_FLOATS; i += 128) {
      float v;
      if (i < N0) v = __ldg(b0 + i);
      else if (i < N0 + N1) v = __ldg(b1 + i - N0);
      else if (i < N0 + N1 + N2) v = __ldg(b2 + i - N0 - N1);
      else v = __ldg(w3 + i - N0 - N1 - N2);
      s_par[i] = v;
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      uint32_t q = 0, head = 0;                 // next chunk number / first free unit
      uint32_t tail_q = 0;                      // oldest chunk whose units are not known to be free yet
      int tail_i = 0, free_units = C::NU;
      auto make_room = [&](int n) {
        while (free_units < n) {
          mbar_wait(empty_bar(tail_q % C::NB), (tail_q / C::NB) & 1);
          free_units += units_of(tail_i);
          if (++tail_i == chunks_per_tile) tail_i = 0;
          ++tail_q;
        }
        free_units -= n;
      };
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int group = tile / tiles_per_group;
        const int row0 = static_cast<int>((tile - group * tiles_per_group) * rows_per_tile + rank * BM);
        const int wcol0 = group * proj_w_col_stride;
        for (int kc = 0; kc < nk0; ++kc, ++q, head += C::U0) {            // fc_0: X box + W0 boxes
          make_room(C::U0);
          const uint32_t fb = (CG == 2) ? mapa(full_bar(q % C::NB), 0) : full_bar(q % C::NB);
          if (rank == 0) mbar_expect_tx(full_bar(q % C::NB), CG * C::U0 * UNIT_BYTES);
          tma_load_2d<CG>(&tmX, fb, unit_addr(head), xcol0 + kc * BK, row0);
#pragma unroll
          for (int j = 0; j < C::SUBS_L0; ++j) {
            const int wrow = (CG == 1) ? j * 128 : j * 256 + static_cast<int>(rank) * 128;
            tma_load_2d<CG>(&tmW0, fb, unit_addr(head + 1 + j), wcol0 + kc * BK, wrow);
          }
        }
        for (int a = 0; a < n_add; ++a, ++q, ++head) {                      // addend boxes: this CTA's rows, local barrier
          make_room(1);
          mbar_expect_tx(afull_bar(a), UNIT_BYTES);
          tma_load_2d<1>(&tmX, afull_bar(a), unit_addr(head), a * BK, row0);
        }
        if (proj) continue;
#pragma unroll 1
        for (int layer = 1; layer <= 2; ++layer) {                          // fc_1 / fc_2: weights only
          const CUtensorMap* tm = (layer == 1) ? &tmW1 : &tmW2;
          const int nk = (layer == 1 ? N0 : N1) / BK;
          for (int kc = 0; kc < nk; ++kc, ++q, head += C::SUBS_L12) {
            make_room(C::SUBS_L12);
            const uint32_t fb = (CG == 2) ? mapa(full_bar(q % C::NB), 0) : full_bar(q % C::NB);
            if (rank == 0) mbar_expect_tx(full_bar(q % C::NB), CG * C::SUBS_L12 * UNIT_BYTES);
#pragma unroll
            for (int j = 0; j < C::SUBS_L12; ++j) {
              const int wrow = (CG == 1) ? j * 128 : static_cast<int>(rank) * 128;
              tma_load_2d<CG>(tm, fb, unit_addr(head + j), kc * BK, wrow);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA, one thread) ===========================
    if (rank == 0 && lane == 0) {
      // one instruction covers 128*CG accumulator columns: weight box j of every CTA of the group
      constexpr uint32_t idesc = umma_idesc(128 * CG, 128 * CG);
      constexpr uint32_t kCols = 128 * CG;
      uint32_t q = 0, head = 0, hphase = 0;
      uint32_t fphase = 0;                                            // bit b: parity the next wait on full[b] expects
      auto wait_full = [&]() {
        const uint32_t b = q % C::NB;
        mbar_wait(full_bar(b), (fphase >> b) & 1u);
        fphase ^= 1u << b;
        tc_fence_after();
      };
      bool first = true;
      int tno = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++tno) {
        if (!first) { mbar_wait(hready_bar, hphase); hphase ^= 1; }   // previous tile's accumulators drained
        first = false;
        tc_fence_after();
        stamp(tno, 0);
        // ---- fc_0: D[0,512) = X · W0^T ----
        for (int kc = 0; kc < nk0; ++kc, head += C::U0) {
          wait_full();
          const uint64_t ad = umma_desc_sw128(unit_addr(head));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
#pragma unroll
            for (int j = 0; j < C::SUBS_L0; ++j) {
              umma_ss<CG>(tmem_base + j * kCols, ad + 2 * k, umma_desc_sw128(unit_addr(head + 1 + j)) + 2 * k, idesc,
                          (kc | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit<CG>(empty_bar(q % C::NB));
          ++q;
        }
        umma_commit<CG>(dfull_bar);
        stamp(tno, 1);
        q += n_add;                                                     // addend boxes belong to the epilogue
        head += n_add;
        if (proj) continue;
        // ---- fc_1: D[256,512) = H1(TMEM [0,256)) · W1^T ;  fc_2: D[256,512) = H2(TMEM [0,128)) · W2^T ----
#pragma unroll 1
        for (int layer = 1; layer <= 2; ++layer) {
          const int nk = (layer == 1 ? N0 : N1) / BK;
          mbar_wait(hready_bar, hphase); hphase ^= 1;
          tc_fence_after();
          stamp(tno, 2 * layer);
          for (int kc = 0; kc < nk; ++kc, head += C::SUBS_L12) {
            wait_full();
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
#pragma unroll
              for (int j = 0; j < C::SUBS_L12; ++j) {
                // A operand: 16 bf16 of K = 8 TMEM columns
                umma_ts<CG>(tmem_base + 256 + j * kCols, tmem_base + kc * (BK / 2) + k * 8,
                            umma_desc_sw128(unit_addr(head + j)) + 2 * k, idesc, (kc | k) != 0 ? 1u : 0u);
              }
            }
            umma_commit<CG>(empty_bar(q % C::NB));
            ++q;
          }
          umma_commit<CG>(dfull_bar);
          stamp(tno, 2 * layer + 1);
        }
      }
    }
  } else {
    // =========================== epilogue warps ===========================
    const int quarter = warp & 3;                                      // TMEM lane quarter this warp may touch
    const uint32_t tq = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t hready_remote = (CG == 2) ? mapa(hready_bar, 0) : hready_bar;
    const float bias3 = __ldg(b3);
    const int r_in_tile = quarter * 32 + lane;                         // row of the CTA's 128-row tile
    // one arrival per WARP (every lane has executed tcgen05.fence::before_thread_sync; __syncwarp orders them before
    // lane 0's release-arrive): 4*CG arrivals complete a phase instead of 128*CG, half of them remote
    auto arrive_hready = [&]() {
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(hready_remote);
        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(hready_bar) : "memory");
      }
    };
    uint32_t dphase = 0;
    int etno = 0;
    const bool estamp = warp == kEpiWarp0 && lane == 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++etno) {
      const int group = tile / tiles_per_group;
      const long long row = (tile - group * tiles_per_group) * rows_per_tile + rank * BM + r_in_tile;
      // ---- after fc_0: H1 = relu(acc + b0) -> bf16 -> TMEM [0,256) ----
      mbar_wait_warp(dfull_bar, dphase); dphase ^= 1;
      tc_fence_after();
      if (estamp) stamp(etno, 6);
      if (proj) {                                           // raw accumulator -> bf16 row of the group's slab
        __nv_bfloat16* const orow = proj_out + group * proj_out_group_stride + row * N0;
#pragma unroll 1
        for (int j = 0; j < N0 / 32; ++j) {
          uint32_t v[32];
          tmem_ld32(tq + j * 32, v);
          if (row < rows) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(v[8 * i + 0]), __uint_as_float(v[8 * i + 1]));
              u.y = pack_bf16x2(__uint_as_float(v[8 * i + 2]), __uint_as_float(v[8 * i + 3]));
              u.z = pack_bf16x2(__uint_as_float(v[8 * i + 4]), __uint_as_float(v[8 * i + 5]));
              u.w = pack_bf16x2(__uint_as_float(v[8 * i + 6]), __uint_as_float(v[8 * i + 7]));
              *reinterpret_cast<uint4*>(orow + j * 32 + 8 * i) = u;
            }
          }
        }
        tc_fence_before();
        arrive_hready();
        continue;
      }
      if (add) {
        // Hoisted rows (hoist.cu): fc_0's accumulator covers only the non-hoisted K columns; the 512-wide addend block of
        // the rows (projected maps / coarse levels + bias b0, bf16) arrived by TMA in NA boxes of 64 columns (128-byte
        // swizzle: row r at r*128, 16-byte slot s at (s ^ (r & 7)) * 16 -- conflict free for consecutive rows).
        const uint32_t q_add0 = static_cast<uint32_t>(etno) * chunks_per_tile + nk0;
        const uint32_t head_add0 = static_cast<uint32_t>((static_cast<long long>(etno) * units_per_tile + nk0 * C::U0) % C::NU);
#pragma unroll 1
        for (int a = 0; a < NA; ++a) {
          mbar_wait_warp(afull_bar(a), etno & 1);
          const uint32_t urow = unit_addr(head_add0 + a) + r_in_tile * 128;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int j = 2 * a + h;
            uint32_t v[32], u[16];
            uint4 pq[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(pq[i].x), "=r"(pq[i].y), "=r"(pq[i].z), "=r"(pq[i].w)
                           : "r"(urow + (((4 * h + i) ^ (r_in_tile & 7)) << 4)) : "memory");
            tmem_ld32(tq + j * 32, v);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t w4[4] = {pq[i].x, pq[i].y, pq[i].z, pq[i].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {                  // packed add, relu folded into the bf16 conversion
                const float2 sum = fadd2(make_float2(__uint_as_float(v[8 * i + 2 * e]), __uint_as_float(v[8 * i + 2 * e + 1])),
                                         bf16x2_to_f2(w4[e]));
                v[8 * i + 2 * e] = __float_as_uint(sum.x);
                v[8 * i + 2 * e + 1] = __float_as_uint(sum.y);
                u[4 * i + e] = pack_bf16x2_relu(sum.x, sum.y);
              }
            }
            tmem_st16(tq + j * 16, u);
            if (dbg1 != nullptr && row < rows) {
              float* const drow = dbg1 + row * N0 + j * 32;
#pragma unroll
              for (int i = 0; i < 32; ++i) drow[i] = fmaxf(__uint_as_float(v[i]), 0.f);
            }
          }
          named_bar_sync(1, 128);                            // all four epilogue warps are done with the box ...
          if (threadIdx.x == kEpiWarp0 * 32) mbar_arrive_local(empty_bar((q_add0 + a) % C::NB));   // ... hand its unit back
        }
      } else {
#pragma unroll 1
      for (int j = 0; j < N0 / 32; ++j) {
        uint32_t v[32], u[16];
        tmem_ld32(tq + j * 32, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 bb = *reinterpret_cast<const float2*>(s_b0 + j * 32 + 2 * i);
          const float2 sum = fadd2(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), bb);
          u[i] = pack_bf16x2_relu(sum.x, sum.y);
        }
        tmem_st16(tq + j * 16, u);
        if (dbg1 != nullptr && row < rows) {               // diagnostic copy of relu(fc_0) (fp32, pre-rounding)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            dbg1[row * N0 + j * 32 + i] = fmaxf(__uint_as_float(v[i]) + s_b0[j * 32 + i], 0.f);
        }
      }
      }
      tmem_wait_st();
      tc_fence_before();
      if (estamp) stamp(etno, 7);
      arrive_hready();
      // ---- after fc_1: H2 = relu(acc + b1) -> bf16 -> TMEM [0,128) ----
      mbar_wait_warp(dfull_bar, dphase); dphase ^= 1;
      tc_fence_after();
      if (estamp) stamp(etno, 8);
#pragma unroll 1
      for (int j = 0; j < N1 / 32; ++j) {
        uint32_t v[32], u[16];
        tmem_ld32(tq + 256 + j * 32, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 bb = *reinterpret_cast<const float2*>(s_b1 + j * 32 + 2 * i);
          const float2 sum = fadd2(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), bb);
          u[i] = pack_bf16x2_relu(sum.x, sum.y);
        }
        tmem_st16(tq + j * 16, u);
        if (dbg2 != nullptr && row < rows) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            dbg2[row * N1 + j * 32 + i] = fmaxf(__uint_as_float(v[i]) + s_b1[j * 32 + i], 0.f);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      if (estamp) stamp(etno, 9);
      arrive_hready();
      // ---- after fc_2: sdf = (relu(acc + b2) · w3 + b3) / out_div ----
      mbar_wait_warp(dfull_bar, dphase); dphase ^= 1;
      tc_fence_after();
      if (estamp) stamp(etno, 10);
      float2 acc2 = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int j = 0; j < N2 / 32; ++j) {
        uint32_t v[32];
        tmem_ld32(tq + 256 + j * 32, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {                       // even / odd columns accumulate in the two packed lanes
          const float2 bb = *reinterpret_cast<const float2*>(s_b2 + j * 32 + 2 * i);
          const float2 ww = *reinterpret_cast<const float2*>(s_w3 + j * 32 + 2 * i);
          const float2 sum = fadd2(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), bb);
          acc2 = ffma2(make_float2(fmaxf(sum.x, 0.f), fmaxf(sum.y, 0.f)), ww, acc2);
        }
        if (dbg3 != nullptr && row < rows) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            dbg3[row * N2 + j * 32 + i] = fmaxf(__uint_as_float(v[i]) + s_b2[j * 32 + i], 0.f);
        }
      }
      tc_fence_before();
      if (estamp) stamp(etno, 11);
      arrive_hready();
      if (row < rows) sdf[row] = __fdiv_rn((acc2.x + acc2.y) + bias3, out_div);
    }
  }

  // =========================== teardown ===========================
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base);
  }
}

// ------------------------------------------------------------------ host side
struct AddArgs {                        // hoisted rows: fc_0 over k columns of X only, addend block added in the epilogue
  const __nv_bfloat16* addend = nullptr;
  int64_t ld = 0;
  int k = 0;                            // K of fc_0 (columns of X / of the W0 view); 0 = w->k_pad
  long long* trace = nullptr;           // diagnostic phase timeline of CTA 0 (see the kernel)
};
struct ProjArgs {                       // projection mode (see the kernel); all zero = the full MLP
  __nv_bfloat16* out = nullptr;         // [groups][rows][512]
  int groups = 1;
  int w_col_stride = 0;                 // columns of W0 between consecutive groups
  int k = 0;                            // K of the projection (columns of X and of each W0 block)
};

template <int CG, int ST>
static int launch(const ListWeights* w, const void* X, int64_t ldx, int64_t rows, float* sdf, float out_div,
                  float* dbg1, float* dbg2, float* dbg3, const ProjArgs& pa, const AddArgs& aa, cudaStream_t st) {
  using C = Cfg<CG, ST>;
  CUtensorMap tmX, tmW0, tmW1, tmW2;
  int rc;
  const bool proj = pa.out != nullptr;
  const int k0 = proj ? pa.k : (aa.k > 0 ? aa.k : w->k_pad);
  const uint64_t w0_cols = proj ? static_cast<uint64_t>(pa.w_col_stride) * (pa.groups - 1) + pa.k : k0;
  // hoisted rows: ONE map over [addend N0 | k0 columns]; fc_0's chunks start at column N0, the addend boxes at 0
  const bool hoisted = aa.addend != nullptr;
  const void* xbase = hoisted ? static_cast<const void*>(aa.addend) : X;
  const uint64_t xcols = hoisted ? static_cast<uint64_t>(N0) + k0 : static_cast<uint64_t>(k0);
  if ((rc = make_map_bf16(&tmX, xbase, xcols, static_cast<uint64_t>(rows), static_cast<uint64_t>(ldx)))) return rc;
  if ((rc = make_map_bf16(&tmW0, w->w0, w0_cols, N0, w->k_pad))) return rc;
  if ((rc = make_map_bf16(&tmW1, w->w1, N0, N1, N0))) return rc;
  if ((rc = make_map_bf16(&tmW2, w->w2, N1, N2, N1))) return rc;
  int dev = 0, sms = 0;
  LIST_CUDA(cudaGetDevice(&dev));
  LIST_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  LIST_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<CG, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  // keep the SM's shared-memory carve-out at its maximum so that gather CTAs of the next chunk can co-reside (api.cu
  // run_chunks); otherwise the carve-out is sized for this kernel alone and nothing else fits until the SM drains
  LIST_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<CG, ST>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  const int64_t rows_per_tile = static_cast<int64_t>(BM) * CG;
  const int64_t tiles = (rows + rows_per_tile - 1) / rows_per_tile * (proj ? pa.groups : 1);
  const int clusters = static_cast<int>(tiles < (sms / CG) ? tiles : (sms / CG));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * CG);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const int nk0 = k0 / BK;
  LIST_CUDA(cudaLaunchKernelEx(&cfg, mlp_tc_kernel<CG, ST>, tmX, tmW0, tmW1, tmW2, w->b0, w->b1, w->b2, w->w3, w->b3,
                               sdf, static_cast<long long>(rows), nk0, out_div, dbg1, dbg2, dbg3, pa.out, pa.groups,
                               pa.w_col_stride, static_cast<long long>(rows) * N0, hoisted ? N0 : 0, hoisted ? 1 : 0, aa.trace));
  return LIST_OK;
}

}  // namespace tc

// variant: 1 = single-CTA tcgen05 (cta_group::1), 2 = CTA pair (cta_group::2)
int mlp_tc_fwd(const ListWeights* w, const void* X, int64_t ldx, int64_t rows, float* sdf, float out_div, int variant,
               float* dbg1, float* dbg2, float* dbg3, cudaStream_t st) {
  if (rows == 0) return LIST_OK;
  LIST_CHECK_ARG(w->n0 == tc::N0 && w->n1 == tc::N1 && w->n2 == tc::N2,
                 "mlp_tc: layer widths must be 512/256/256 (got %d/%d/%d)", w->n0, w->n1, w->n2);
  LIST_CHECK_ARG(w->k_pad % tc::BK == 0 && w->k_pad > 0, "mlp_tc: k_pad %d must be a positive multiple of 64", w->k_pad);
  LIST_CHECK_ARG(ldx % 8 == 0 && ldx >= w->k_pad, "mlp_tc: ldx %lld must be >= k_pad and a multiple of 8", (long long)ldx);
  LIST_CHECK_ARG((reinterpret_cast<uintptr_t>(X) & 15) == 0, "mlp_tc: X must be 16-byte aligned");
  LIST_CHECK_ARG(rows < (1LL << 31), "mlp_tc: rows %lld too large for one call", (long long)rows);
  const tc::ProjArgs none;
  const tc::AddArgs noadd;
  if (variant == 1) return tc::launch<1, 2>(w, X, ldx, rows, sdf, out_div, dbg1, dbg2, dbg3, none, noadd, st);
  // variant 2 = CTA pair with a 4-stage operand ring (192 KB); variant 3 = the same with 3 stages (144 KB),
  // which leaves shared memory for gather CTAs of the next chunk to co-reside on the SM (api.cu pipeline).
  if (variant == 3) return tc::launch<2, 3>(w, X, ldx, rows, sdf, out_div, dbg1, dbg2, dbg3, none, noadd, st);
  return tc::launch<2, 4>(w, X, ldx, rows, sdf, out_div, dbg1, dbg2, dbg3, none, noadd, st);
}

// Hoisted rows (hoist.cu): Xh[rows][ldx] = [addend n0 (incl. bias b0) | k columns]; fc_0 runs on W0[:, col0 : col0 + k]
// only and the addend block is added to its accumulator in the epilogue; fc_1, fc_2, fc_out as in mlp_tc_fwd.
int mlp_tc_fwd_hoisted(const ListWeights* w, int col0, int k, const void* Xh, int64_t ldx, int64_t rows, float* sdf,
                       float out_div, int variant, float* dbg1, float* dbg2, float* dbg3, long long* trace, cudaStream_t st) {
  if (rows == 0) return LIST_OK;
  LIST_CHECK_ARG(w->n0 == tc::N0 && w->n1 == tc::N1 && w->n2 == tc::N2,
                 "mlp_tc: layer widths must be 512/256/256 (got %d/%d/%d)", w->n0, w->n1, w->n2);
  LIST_CHECK_ARG(k > 0 && k % tc::BK == 0 && col0 % 8 == 0 && col0 + k <= w->k_pad, "mlp_tc hoisted: W0 columns [%d, +%d) invalid for k_pad %d",
                 col0, k, w->k_pad);
  LIST_CHECK_ARG(ldx % 8 == 0 && ldx >= tc::N0 + k, "mlp_tc hoisted: ldx %lld must be >= %d and a multiple of 8", (long long)ldx, tc::N0 + k);
  LIST_CHECK_ARG((reinterpret_cast<uintptr_t>(Xh) & 15) == 0, "mlp_tc hoisted: X must be 16-byte aligned");
  LIST_CHECK_ARG(rows < (1LL << 31), "mlp_tc hoisted: rows %lld too large for one call", (long long)rows);
  ListWeights wv = *w;
  wv.w0 = static_cast<const __nv_bfloat16*>(w->w0) + col0;
  const __nv_bfloat16* xh = static_cast<const __nv_bfloat16*>(Xh);
  tc::AddArgs aa;
  aa.addend = xh;
  aa.ld = ldx;
  aa.k = k;
  aa.trace = trace;
  const tc::ProjArgs none;
  if (variant == 3) return tc::launch<2, 3>(&wv, xh + tc::N0, ldx, rows, sdf, out_div, dbg1, dbg2, dbg3, none, aa, st);
  return tc::launch<2, 4>(&wv, xh + tc::N0, ldx, rows, sdf, out_div, dbg1, dbg2, dbg3, none, aa, st);
}

// Projection through fc_0 only (hoist.cu): out[g][r][0..512) = sum_k X[r][k] * W0[n][col0 + g*col_stride + k],
// g < groups, bf16 in, fp32 accumulate on tcgen05, bf16 out, no bias / activation.  `w` supplies W0
// (row pitch w->k_pad); X is [rows][ldx] bf16 with k <= ldx columns used.
int mlp_tc_project(const ListWeights* w, int col0, int col_stride, int groups, int k, const void* X, int64_t ldx,
                   int64_t rows, void* out, cudaStream_t st) {
  if (rows == 0 || groups == 0) return LIST_OK;
  LIST_CHECK_ARG(w->n0 == tc::N0, "mlp_tc_project: fc_0 width must be 512 (got %d)", w->n0);
  LIST_CHECK_ARG(k > 0 && k % tc::BK == 0 && ldx >= k && ldx % 8 == 0, "mlp_tc_project: k %d must be a positive multiple of 64 and <= ldx %lld",
                 k, (long long)ldx);
  LIST_CHECK_ARG(col0 % 8 == 0 && col_stride % 8 == 0 && col0 + static_cast<int64_t>(col_stride) * (groups - 1) + k <= w->k_pad,
                 "mlp_tc_project: W0 column block [%d + g*%d, +%d) outside k_pad %d or unaligned", col0, col_stride, k, w->k_pad);
  LIST_CHECK_ARG((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                 "mlp_tc_project: X/out must be 16-byte aligned");
  LIST_CHECK_ARG(rows * groups < (1LL << 31), "mlp_tc_project: too many rows");
  ListWeights wv = *w;                                    // view of W0 starting at column col0
  wv.w0 = static_cast<const __nv_bfloat16*>(w->w0) + col0;
  tc::ProjArgs pa;
  pa.out = static_cast<__nv_bfloat16*>(out);
  pa.groups = groups;
  pa.w_col_stride = col_stride;
  pa.k = k;
  return tc::launch<2, 4>(&wv, X, ldx, rows, nullptr, 1.0f, nullptr, nullptr, nullptr, pa, tc::AddArgs{}, st);
}

}  // namespace list
