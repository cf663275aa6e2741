// fp32-accurate GEMM family on the tensor cores ("3xTF32"), used by the fp32 parity path of the implicit MLP (a-6,
// reference network/modules.py:276-282) and by its backward (a-9, reference train.py:72-85 through autograd).
//
// tcgen05 has no fp32 MMA: kind::tf32 reads the upper 19 bits of a 32-bit operand (1 + 8 + 10).  Writing every operand as
// x = hi + lo with hi = those 19 bits (what the tensor core sees when it is handed x itself) and lo = x - hi (exact in
// fp32; at most 13 significant bits, rounded to the 11 the tensor core keeps),
//        a . b  ~=  hi(a) . hi(b) + lo(a) . hi(b) + hi(a) . lo(b)          (fp32 accumulation in TMEM)
// drops only lo . lo and the last two bits of lo: a relative error of ~2^-21 per product instead of TF32's 2^-11.  That
// keeps the 1e-4 parity bound over K = 3610 with three MMAs per product at the TF32 rate (1.1 PFLOP/s nominal), i.e.
// an order of magnitude above the FFMA pipe (sgemm.cuh, kept as the LIST_B200_F32_TC=0 reference).
//
//   C[m][n] (+)= epi( sum_k A[m*lda + k] * B[n*ldb + k] )      both operands K-major, epilogue as sgemm.cuh
// Both operands are K-major (MN-major TF32 operands returned zeros in bring-up; the backward transposes its operands
// instead: transpose_split_kernel).  The lo parts of the weights and of the transposed operands are separate tensors
// written by split_lo_kernel / transpose_split_kernel; lo(A) of an activation operand (Alo == nullptr) is computed in the
// kernel from the A box in shared memory by the epilogue warps.  TMA streams 128-byte
// swizzled fp32 boxes of x and lo(x) for both operands into a 2-stage ring (96 KB per stage: 128 x 32 of A, 256 x 32 of
// B, each twice), one thread issues three tcgen05.mma.kind::tf32 (M = 128, N = 256, K = 8) per k-step, four epilogue
// warps drain the 128 x 256 fp32 accumulator from TMEM with bias / ReLU / ReLU-mask / accumulate (atomics under split-K).
#include <cstdlib>

#include "tc_common.cuh"
#include "tgemm.cuh"

namespace list {
namespace tg {

using namespace tc;

constexpr int BM = 128, BN = 256, BKF = 32;        // tile; BKF fp32 = 128 B = one swizzle row
constexpr int ST = 2;
constexpr int A_BYTES = BM * BKF * 4, B_BYTES = BN * BKF * 4;
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;          // A, lo(A), B, lo(B)
constexpr int BAR_OFF = ST * STAGE_BYTES;
constexpr int SMEM_BYTES = BAR_OFF + (3 * ST + 1) * 8 + 16 + 1024;       // full, empty, ready (lo(A) in place) | dfull
constexpr int kThreads = 192;

// Instruction descriptor: D = f32, A = B = tf32, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t idesc_tf32() {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) | (static_cast<uint32_t>(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
// hi(x) = the upper 19 bits of x -- exactly what the tensor core makes of the 32-bit operand it is handed (measured on
// B200: it truncates; with lo = x - round_to_tf32(x) the corrected product is no better than plain TF32).  lo(x) = x - hi(x)
// is exact in fp32 with up to 13 significant bits, of which the tensor core again keeps the upper 11: lo is therefore
// rounded to nearest TF32 HERE, so that what is lost (<= 2^-23 |x|) is unbiased instead of a truncation that always
// points towards zero and adds up over K.
__device__ __forceinline__ float tf32_lo(float x) {
  const float lo = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(lo));
  return __uint_as_float(r);
}
// inline_lo_a: lo(A) is not streamed from memory but computed from the A box in shared memory by the four epilogue warps
// (idle during the main loop otherwise), so the activations need neither a lo tensor nor the pass that writes it.
__global__ void __launch_bounds__(kThreads, 1)
tgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAl,
             const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBl,
             float* __restrict__ C, int64_t ldc, int M, int N, int K, GemmEpilogue ep, int inline_lo_a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* const gbase = smem_raw + (base - raw);
  const uint32_t bar0 = base + BAR_OFF;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (ST + s); };
  auto ready_bar = [&](int s) { return bar0 + 8u * (2 * ST + s); };
  const uint32_t dfull_bar = bar0 + 8u * (3 * ST);
  const uint32_t tmem_slot = dfull_bar + 8u;
  volatile uint32_t* const tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + BAR_OFF + (3 * ST + 1) * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  // K range of this z-slice (whole chunks of BKF)
  const int nk_all = (K + BKF - 1) / BKF;
  const int nk_per = (nk_all + gridDim.z - 1) / gridDim.z;
  const int kc0 = blockIdx.z * nk_per;
  const int nk = min(nk_all - kc0, nk_per);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmAl); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmBl);
    for (int s = 0; s < ST; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); mbar_init(ready_bar(s), 4); }
    mbar_init(dfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (nk > 0) {
    if (warp == 0) {
      if (lane == 0) {
        for (int i = 0; i < nk; ++i) {
          const int s = i % ST;
          if (i >= ST) mbar_wait(empty_bar(s), ((i / ST) - 1) & 1);
          const uint32_t st = base + s * STAGE_BYTES;
          const int k0 = (kc0 + i) * BKF;
          mbar_expect_tx(full_bar(s), inline_lo_a ? STAGE_BYTES - A_BYTES : STAGE_BYTES);
          const CUtensorMap* am[2] = {&tmA, &tmAl};
          const CUtensorMap* bm[2] = {&tmB, &tmBl};
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t a = st + h * A_BYTES, b = st + 2 * A_BYTES + h * B_BYTES;
            if (h == 0 || !inline_lo_a) tma_load_2d<1>(am[h], full_bar(s), a, k0, m0);   // box {32 k, 128 m}
            tma_load_2d<1>(bm[h], full_bar(s), b, k0, n0);                       // box {32 k, 256 n}
          }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = idesc_tf32();
        for (int i = 0; i < nk; ++i) {
          const int s = i % ST;
          if (inline_lo_a) mbar_wait(ready_bar(s), (i / ST) & 1);     // the epilogue warps wrote lo(A) (they waited for `full`)
          else mbar_wait(full_bar(s), (i / ST) & 1);
          tc_fence_after();
          const uint32_t st = base + s * STAGE_BYTES;
          const uint32_t a[2] = {st, st + A_BYTES}, b[2] = {st + 2 * A_BYTES, st + 2 * A_BYTES + B_BYTES};
#pragma unroll
          for (int k = 0; k < BKF / 8; ++k) {
            uint64_t ad[2], bd[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              ad[h] = umma_desc_sw128(a[h]) + 2 * k;                                 // 8 k = 32 B along a row
              bd[h] = umma_desc_sw128(b[h]) + 2 * k;
            }
            umma_tf32(tmem_base, ad[0], bd[0], idesc, (i | k) != 0 ? 1u : 0u);   // hi . hi
            umma_tf32(tmem_base, ad[1], bd[0], idesc, 1u);                       // lo . hi
            umma_tf32(tmem_base, ad[0], bd[1], idesc, 1u);                       // hi . lo
          }
          umma_commit<1>(empty_bar(s));
        }
        umma_commit<1>(dfull_bar);
      }
    } else {
      // =========================== epilogue: one TMEM lane quarter per warp, lane = row ===========================
      const int quarter = warp & 3;
      const uint32_t tq = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
      const int m = m0 + quarter * 32 + lane;
      if (inline_lo_a) {
        // lo(A) box = elementwise transform of the A box (same 128B-swizzled layout, so the same offsets): 1024 16-byte
        // pieces per stage, 8 per thread
        const int t = (warp - 2) * 32 + lane;
        for (int i = 0; i < nk; ++i) {
          const int s = i % ST;
          mbar_wait_warp(full_bar(s), (i / ST) & 1);
          const float4* src = reinterpret_cast<const float4*>(gbase + s * STAGE_BYTES);
          float4* dst = reinterpret_cast<float4*>(gbase + s * STAGE_BYTES + A_BYTES);
#pragma unroll
          for (int j = 0; j < A_BYTES / 16 / 128; ++j) {
            const float4 v = src[t + j * 128];
            dst[t + j * 128] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive_local(ready_bar(s));
        }
      }
      mbar_wait_warp(dfull_bar, 0);
      tc_fence_after();
      const bool atomic = gridDim.z > 1;
#pragma unroll 1
      for (int j = 0; j < BN / 32; ++j) {
        uint32_t v[32];
        tmem_ld32(tq + j * 32, v);
        const int n = n0 + j * 32;
        if (m < M && n < N) {
          float* dst = C + static_cast<int64_t>(m) * ldc + n;
          const float* mk = ep.mask ? ep.mask + static_cast<int64_t>(m) * ep.ldmask + n : nullptr;
          if (n + 32 <= N && !atomic) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float x[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float t = __uint_as_float(v[4 * i + e]);
                if (ep.bias) t += __ldg(ep.bias + n + 4 * i + e);
                if (ep.relu) t = fmaxf(t, 0.f);
                if (mk) t = __ldg(mk + 4 * i + e) > 0.f ? t : 0.f;
                x[e] = t;
              }
              float4 o = make_float4(x[0], x[1], x[2], x[3]);
              if (ep.accumulate) {
                const float4 c = *reinterpret_cast<const float4*>(dst + 4 * i);
                o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
              }
              *reinterpret_cast<float4*>(dst + 4 * i) = o;
            }
          } else {
            for (int i = 0; i < 32 && n + i < N; ++i) {
              float t = __uint_as_float(v[i]);
              if (ep.bias) t += __ldg(ep.bias + n + i);
              if (ep.relu) t = fmaxf(t, 0.f);
              if (mk) t = __ldg(mk + i) > 0.f ? t : 0.f;
              if (atomic) atomicAdd(dst + i, t);
              else dst[i] = ep.accumulate ? dst[i] + t : t;
            }
          }
        }
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base);
  }
}

// ---------------------------------------------------------------- CTA-pair variant (cta_group::2, M = 256)
// Two CTAs of a cluster work on 256 x 256 of C: each loads its own 128 rows of A (and lo(A)) but only HALF of the B / lo(B)
// tile (the pair's tensor cores read both halves), which cuts the L2 -> SM traffic per CTA and k-chunk from 80 to 48 KB and
// leaves room for three stages.  B parts signal the leader's `full` barrier; A parts a barrier local to their CTA, where
// the four epilogue warps wait (and write lo(A) when it is computed in the kernel) before they report to the leader's
// `ready` barrier; one thread of the leader issues the MMAs for both.
namespace pair {
constexpr int PST = 3;
constexpr int BH_BYTES = (BN / 2) * BKF * 4;                      // this CTA's half of the B tile
constexpr int PSTAGE_BYTES = 2 * A_BYTES + 2 * BH_BYTES;          // A, lo(A), B half, lo(B) half
constexpr int PBAR_OFF = PST * PSTAGE_BYTES;
constexpr int PNUM_BARS = 4 * PST + 1;                              // full, afull, empty, ready | dfull
constexpr int PSMEM_BYTES = PBAR_OFF + PNUM_BARS * 8 + 16 + 1024;
}  // namespace pair

__device__ __forceinline__ void umma_tf32_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
tgemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAl,
                  const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBl,
                  float* __restrict__ C, int64_t ldc, int M, int N, int K, GemmEpilogue ep, int inline_lo_a) {
  using namespace pair;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* const gbase = smem_raw + (base - raw);
  const uint32_t bar0 = base + PBAR_OFF;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto afull_bar = [&](int s) { return bar0 + 8u * (PST + s); };
  auto empty_bar = [&](int s) { return bar0 + 8u * (2 * PST + s); };
  auto ready_bar = [&](int s) { return bar0 + 8u * (3 * PST + s); };
  const uint32_t dfull_bar = bar0 + 8u * (4 * PST);
  const uint32_t tmem_slot = dfull_bar + 8u;
  volatile uint32_t* const tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + PBAR_OFF + PNUM_BARS * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;           // blockIdx.x = 2 * pair + rank: this CTA's 128 rows
  const int nk_all = (K + BKF - 1) / BKF;
  const int nk_per = (nk_all + gridDim.z - 1) / gridDim.z;
  const int kc0 = blockIdx.z * nk_per;
  const int nk = min(nk_all - kc0, nk_per);                       // the same in both CTAs of a pair

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmAl); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmBl);
    for (int s = 0; s < PST; ++s) {
      mbar_init(full_bar(s), 1);                                  // leader: its own expect_tx for the B parts of both CTAs
      mbar_init(afull_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(ready_bar(s), 8);                                 // leader: four epilogue warps of each CTA
    }
    mbar_init(dfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<2>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (nk > 0) {
    if (warp == 0) {
      if (lane == 0) {
        const uint32_t full_leader = mapa(full_bar(0), 0);        // + 8 * s
        for (int i = 0; i < nk; ++i) {
          const int s = i % PST;
          if (i >= PST) mbar_wait(empty_bar(s), ((i / PST) - 1) & 1);
          const uint32_t st = base + s * PSTAGE_BYTES;
          const int k0 = (kc0 + i) * BKF;
          if (rank == 0) mbar_expect_tx(full_bar(s), 4 * BH_BYTES);
          mbar_expect_tx(afull_bar(s), inline_lo_a ? A_BYTES : 2 * A_BYTES);
          tma_load_2d<1>(&tmA, afull_bar(s), st, k0, m0);                                  // box {32 k, 128 m}
          if (!inline_lo_a) tma_load_2d<1>(&tmAl, afull_bar(s), st + A_BYTES, k0, m0);
          // this CTA's half of the B tile: rows n0 + 128 rank ..; the maps' box is {32 k, 128 n}
          tma_load_2d<2>(&tmB, full_leader + 8u * s, st + 2 * A_BYTES, k0, n0 + static_cast<int>(rank) * (BN / 2));
          tma_load_2d<2>(&tmBl, full_leader + 8u * s, st + 2 * A_BYTES + BH_BYTES, k0, n0 + static_cast<int>(rank) * (BN / 2));
        }
      }
    } else if (warp == 1) {
      if (rank == 0 && lane == 0) {
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) |
                                   (static_cast<uint32_t>((2 * BM) >> 4) << 24);
        for (int i = 0; i < nk; ++i) {
          const int s = i % PST;
          mbar_wait(full_bar(s), (i / PST) & 1);
          mbar_wait(ready_bar(s), (i / PST) & 1);
          tc_fence_after();
          const uint32_t st = base + s * PSTAGE_BYTES;
          const uint32_t a[2] = {st, st + A_BYTES}, b[2] = {st + 2 * A_BYTES, st + 2 * A_BYTES + BH_BYTES};
#pragma unroll
          for (int k = 0; k < BKF / 8; ++k) {
            uint64_t ad[2], bd[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              ad[h] = umma_desc_sw128(a[h]) + 2 * k;
              bd[h] = umma_desc_sw128(b[h]) + 2 * k;
            }
            umma_tf32_pair(tmem_base, ad[0], bd[0], idesc, (i | k) != 0 ? 1u : 0u);   // hi . hi
            umma_tf32_pair(tmem_base, ad[1], bd[0], idesc, 1u);                       // lo . hi
            umma_tf32_pair(tmem_base, ad[0], bd[1], idesc, 1u);                       // hi . lo
          }
          umma_commit<2>(empty_bar(s));
        }
        umma_commit<2>(dfull_bar);
      }
    } else {
      // =========================== epilogue warps: A relay / lo(A) during the main loop, then the drain ===========================
      const int quarter = warp & 3;
      const uint32_t tq = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
      const int m = m0 + quarter * 32 + lane;
      {
        const int t = (warp - 2) * 32 + lane;
        const uint32_t ready_leader = mapa(ready_bar(0), 0);
        for (int i = 0; i < nk; ++i) {
          const int s = i % PST;
          mbar_wait_warp(afull_bar(s), (i / PST) & 1);
          if (inline_lo_a) {
            const float4* src = reinterpret_cast<const float4*>(gbase + s * PSTAGE_BYTES);
            float4* dst = reinterpret_cast<float4*>(gbase + s * PSTAGE_BYTES + A_BYTES);
#pragma unroll
            for (int j = 0; j < A_BYTES / 16 / 128; ++j) {
              const float4 v = src[t + j * 128];
              dst[t + j * 128] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
            }
            fence_proxy_async_smem();
          }
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(ready_leader + 8u * s);
        }
      }
      mbar_wait_warp(dfull_bar, 0);
      tc_fence_after();
      const bool atomic = gridDim.z > 1;
#pragma unroll 1
      for (int j = 0; j < BN / 32; ++j) {
        uint32_t v[32];
        tmem_ld32(tq + j * 32, v);
        const int n = n0 + j * 32;
        if (m < M && n < N) {
          float* dst = C + static_cast<int64_t>(m) * ldc + n;
          const float* mk = ep.mask ? ep.mask + static_cast<int64_t>(m) * ep.ldmask + n : nullptr;
          if (n + 32 <= N && !atomic) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float x[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float tt = __uint_as_float(v[4 * i + e]);
                if (ep.bias) tt += __ldg(ep.bias + n + 4 * i + e);
                if (ep.relu) tt = fmaxf(tt, 0.f);
                if (mk) tt = __ldg(mk + 4 * i + e) > 0.f ? tt : 0.f;
                x[e] = tt;
              }
              float4 o = make_float4(x[0], x[1], x[2], x[3]);
              if (ep.accumulate) {
                const float4 c = *reinterpret_cast<const float4*>(dst + 4 * i);
                o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
              }
              *reinterpret_cast<float4*>(dst + 4 * i) = o;
            }
          } else {
            for (int i = 0; i < 32 && n + i < N; ++i) {
              float tt = __uint_as_float(v[i]);
              if (ep.bias) tt += __ldg(ep.bias + n + i);
              if (ep.relu) tt = fmaxf(tt, 0.f);
              if (mk) tt = __ldg(mk + i) > 0.f ? tt : 0.f;
              if (atomic) atomicAdd(dst + i, tt);
              else dst[i] = ep.accumulate ? dst[i] + tt : tt;
            }
          }
        }
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<2>(tmem_base);
  }
}

__global__ void __launch_bounds__(256) split_lo_kernel(const float* __restrict__ x, float* __restrict__ lo, int64_t n4) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    reinterpret_cast<float4*>(lo)[i] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
  }
}

// xT[c][r] = x[r][c], loT[c][r] = lo(x[r][c]) for an [rows][cols] matrix with row pitch ld; xT / loT have row pitch ldt.
// 32 x 32 tiles through shared memory (padded against bank conflicts).
__global__ void __launch_bounds__(256) transpose_split_kernel(const float* __restrict__ x, int64_t ld, int rows, int cols,
                                                              float* __restrict__ xT, float* __restrict__ loT, int64_t ldt) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;              // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    tile[ty + 8 * i][tx] = (r < rows && c < cols) ? __ldg(x + static_cast<int64_t>(r) * ld + c) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + tx;
    if (c < cols && r < ldt) {                                          // the pad columns [rows, ldt) are written as zeros
      const float v = tile[tx][ty + 8 * i];
      xT[static_cast<int64_t>(c) * ldt + r] = v;
      loT[static_cast<int64_t>(c) * ldt + r] = tf32_lo(v);
    }
  }
}

// fp32 row-major [outer][inner] with row pitch `pitch` elements; box = 32 x box_outer, 128B swizzle
static int make_map_f32(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch, uint32_t box_outer) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return LIST_ENOSYS; }
  const cuuint64_t dims[2] = {inner, outer};
  const cuuint64_t strides[1] = {pitch * 4};
  const cuuint32_t box[2] = {32, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (fp32) failed (CUresult %d)", static_cast<int>(r)); return LIST_ECUDA; }
  return LIST_OK;
}

static int launch(const float* A, const float* Alo, int64_t lda, const float* B, const float* Blo, int64_t ldb, float* C, int64_t ldc,
                  int M, int N, int K, const GemmEpilogue& ep, cudaStream_t st) {
  CUtensorMap tmA, tmAl, tmB, tmBl;
  int rc;
  if ((rc = make_map_f32(&tmA, A, K, M, lda, BM))) return rc;
  if ((rc = make_map_f32(&tmAl, Alo ? Alo : A, K, M, lda, BM))) return rc;   // Alo == nullptr: lo(A) is computed in the kernel
  if ((rc = make_map_f32(&tmB, B, K, N, ldb, BN))) return rc;
  if ((rc = make_map_f32(&tmBl, Blo, K, N, ldb, BN))) return rc;
  dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN);
  // split K when the output alone cannot fill the GPU and the epilogue is a plain accumulation (weight gradients)
  if (ep.accumulate && !ep.bias && !ep.relu && !ep.mask && K >= 2048) {
    const int tiles = grid.x * grid.y;
    int split = (148 + tiles - 1) / tiles;
    const int max_split = K / 512;
    if (split > max_split) split = max_split;
    if (split > 1) grid.z = split;
  }
  static thread_local int attr_dev = -1;
  int dev = 0;
  LIST_CUDA(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    LIST_CUDA(cudaFuncSetAttribute(tgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    LIST_CUDA(cudaFuncSetAttribute(tgemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pair::PSMEM_BYTES));
    attr_dev = dev;
  }
  // CTA pairs (M = 256 per MMA, half of the B tile per CTA) for everything with at least two row tiles; LIST_B200_TGEMM_PAIR=0
  // keeps the single-CTA kernel (A/B aid)
  const char* e = getenv("LIST_B200_TGEMM_PAIR");
  if (grid.x >= 2 && !(e && e[0] == '0')) {
    CUtensorMap tmBh, tmBlh;
    if ((rc = make_map_f32(&tmBh, B, K, N, ldb, BN / 2))) return rc;
    if ((rc = make_map_f32(&tmBlh, Blo, K, N, ldb, BN / 2))) return rc;
    grid.x = (grid.x + 1) / 2 * 2;
    tgemm_pair_kernel<<<grid, kThreads, pair::PSMEM_BYTES, st>>>(tmA, tmAl, tmBh, tmBlh, C, ldc, M, N, K, ep, Alo == nullptr ? 1 : 0);
    LIST_LAUNCH_CHECK("tgemm_pair_kernel");
    return LIST_OK;
  }
  tgemm_kernel<<<grid, kThreads, SMEM_BYTES, st>>>(tmA, tmAl, tmB, tmBl, C, ldc, M, N, K, ep, Alo == nullptr ? 1 : 0);
  LIST_LAUNCH_CHECK("tgemm_kernel");
  return LIST_OK;
}

}  // namespace tg

int split_lo(const float* x, float* lo, int64_t n, cudaStream_t st) {
  if (n <= 0) return LIST_OK;
  LIST_CHECK_ARG(n % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(lo)) & 15) == 0,
                 "split_lo: %lld elements / pointers must be multiples of 4 floats", (long long)n);
  const int64_t n4 = n / 4;
  const unsigned blocks = static_cast<unsigned>(n4 / 256 + 1 < 148 * 16 ? n4 / 256 + 1 : 148 * 16);
  tg::split_lo_kernel<<<blocks, 256, 0, st>>>(x, lo, n4);
  LIST_LAUNCH_CHECK("split_lo_kernel");
  return LIST_OK;
}

int transpose_split(const float* x, int64_t ld, int rows, int cols, float* xT, float* loT, int64_t ldt, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return LIST_OK;
  LIST_CHECK_ARG(ldt >= rows && ldt % 4 == 0, "transpose_split: ldt %lld must be >= rows and a multiple of 4", (long long)ldt);
  dim3 grid((cols + 31) / 32, static_cast<unsigned>((ldt + 31) / 32));
  tg::transpose_split_kernel<<<grid, 256, 0, st>>>(x, ld, rows, cols, xT, loT, ldt);
  LIST_LAUNCH_CHECK("transpose_split_kernel");
  return LIST_OK;
}

int tgemm(const float* A, const float* Alo, int64_t lda, const float* B, const float* Blo, int64_t ldb, float* C, int64_t ldc, int M,
          int N, int K, const GemmEpilogue& ep, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return LIST_OK;
  LIST_CHECK_ARG(lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0, "tgemm: leading dimensions must be multiples of 4 floats (TMA pitch)");
  LIST_CHECK_ARG(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Alo) | reinterpret_cast<uintptr_t>(B) |
                   reinterpret_cast<uintptr_t>(Blo) | reinterpret_cast<uintptr_t>(C)) & 15) == 0, "tgemm: operands must be 16-byte aligned");
  return tg::launch(A, Alo, lda, B, Blo, ldb, C, ldc, M, N, K, ep, st);
}

}  // namespace list
