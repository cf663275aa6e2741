// fp32 FFMA GEMM family: the reference implementation (LIST_B200_F32_TC=0) of the fp32 parity path of the implicit MLP
// (a-6, reference network/modules.py:276-282) and of its backward (a-9).  The default is tgemm.cu, which reaches the same
// accuracy on the tensor cores with three TF32 products per fp32 product.
//
//   C[m][n] (+)= epi( sum_k A(m,k) * B(k,n) )
//   A_KMAJOR : A(m,k) = A[m*lda + k]   else A(m,k) = A[k*lda + m]
//   B_KMAJOR : B(k,n) = B[n*ldb + k]   else B(k,n) = B[k*ldb + n]
// 128x128x16 tiles, 256 threads, 8x8 outputs per thread, register-prefetched double buffering.
// Split-K (gridDim.z > 1): every z-slice reduces its own K range and adds it to C with fp32 atomics -- used for the
// weight gradients (M x N = 512 x 3648 outputs over K = 16 384 rows would otherwise run on 116 CTAs); only valid for
// plain accumulation (accumulate = 1, no bias / ReLU / mask), which is what the launcher enforces.
#pragma once
#include "common.cuh"
#include "tgemm.cuh"

namespace list {

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, int64_t lda,
                                                    const float* __restrict__ B, int64_t ldb,
                                                    float* __restrict__ C, int64_t ldc, int M, int N,
                                                    int K, GemmEpilogue ep) {
  __shared__ __align__(16) float As[2][16][128];
  __shared__ __align__(16) float Bs[2][16][128];
  const int tid = threadIdx.x;
  const int bm = blockIdx.x * 128;
  const int bn = blockIdx.y * 128;
  const int tx = tid & 15, ty = tid >> 4;

  float4 ra[2], rb[2];
  auto load_tile = [&](int k0) {
    if (A_KMAJOR) {
      const int row = tid & 127, kh = (tid >> 7) * 8;
      const bool ok = (bm + row) < M;
      const float* src = A + static_cast<int64_t>(bm + row) * lda + k0 + kh;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int k = k0 + kh + i * 4;
        if (ok && k + 3 < K) ra[i] = __ldg(reinterpret_cast<const float4*>(src + i * 4));
        else {
          float t[4];
          for (int j = 0; j < 4; ++j) t[j] = (ok && k + j < K) ? __ldg(src + i * 4 + j) : 0.f;
          ra[i] = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int k = k0 + (tid >> 5) + i * 8, m = bm + (tid & 31) * 4;
        if (k < K && m + 3 < M) ra[i] = __ldg(reinterpret_cast<const float4*>(A + static_cast<int64_t>(k) * lda + m));
        else {
          float t[4];
          for (int j = 0; j < 4; ++j) t[j] = (k < K && m + j < M) ? __ldg(A + static_cast<int64_t>(k) * lda + m + j) : 0.f;
          ra[i] = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
    }
    if (B_KMAJOR) {
      const int row = tid & 127, kh = (tid >> 7) * 8;
      const bool ok = (bn + row) < N;
      const float* src = B + static_cast<int64_t>(bn + row) * ldb + k0 + kh;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int k = k0 + kh + i * 4;
        if (ok && k + 3 < K) rb[i] = __ldg(reinterpret_cast<const float4*>(src + i * 4));
        else {
          float t[4];
          for (int j = 0; j < 4; ++j) t[j] = (ok && k + j < K) ? __ldg(src + i * 4 + j) : 0.f;
          rb[i] = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int k = k0 + (tid >> 5) + i * 8, n = bn + (tid & 31) * 4;
        if (k < K && n + 3 < N) rb[i] = __ldg(reinterpret_cast<const float4*>(B + static_cast<int64_t>(k) * ldb + n));
        else {
          float t[4];
          for (int j = 0; j < 4; ++j) t[j] = (k < K && n + j < N) ? __ldg(B + static_cast<int64_t>(k) * ldb + n + j) : 0.f;
          rb[i] = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
    }
  };
  auto store_tile = [&](int buf) {
    if (A_KMAJOR) {
      const int row = tid & 127, kh = (tid >> 7) * 8;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        As[buf][kh + i * 4 + 0][row] = ra[i].x;
        As[buf][kh + i * 4 + 1][row] = ra[i].y;
        As[buf][kh + i * 4 + 2][row] = ra[i].z;
        As[buf][kh + i * 4 + 3][row] = ra[i].w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 2; ++i)
        *reinterpret_cast<float4*>(&As[buf][(tid >> 5) + i * 8][(tid & 31) * 4]) = ra[i];
    }
    if (B_KMAJOR) {
      const int row = tid & 127, kh = (tid >> 7) * 8;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        Bs[buf][kh + i * 4 + 0][row] = rb[i].x;
        Bs[buf][kh + i * 4 + 1][row] = rb[i].y;
        Bs[buf][kh + i * 4 + 2][row] = rb[i].z;
        Bs[buf][kh + i * 4 + 3][row] = rb[i].w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 2; ++i)
        *reinterpret_cast<float4*>(&Bs[buf][(tid >> 5) + i * 8][(tid & 31) * 4]) = rb[i];
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // K range of this z-slice (whole tiles of 16)
  const int nk_all = (K + 15) / 16;
  const int nk_per = (nk_all + gridDim.z - 1) / gridDim.z;
  const int kt0 = blockIdx.z * nk_per;
  const int nk = min(nk_all - kt0, nk_per);
  if (nk <= 0) return;
  load_tile(kt0 * 16);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile((kt0 + kt + 1) * 16);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = bm + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = bn + jh * 64 + tx * 4;
      if (n >= N) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x = acc[i][jh * 4 + j];
        if (n + j < N) {
          if (ep.bias) x += __ldg(ep.bias + n + j);
          if (ep.relu) x = fmaxf(x, 0.f);
          if (ep.mask) x = (__ldg(ep.mask + static_cast<int64_t>(m) * ep.ldmask + n + j) > 0.f) ? x : 0.f;
        }
        v[j] = x;
      }
      float* dst = C + static_cast<int64_t>(m) * ldc + n;
      if (gridDim.z > 1) {
        for (int j = 0; j < 4 && n + j < N; ++j) atomicAdd(dst + j, v[j]);
      } else if (n + 3 < N) {
        float4 o = make_float4(v[0], v[1], v[2], v[3]);
        if (ep.accumulate) {
          const float4 c = *reinterpret_cast<const float4*>(dst);
          o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
        }
        *reinterpret_cast<float4*>(dst) = o;
      } else {
        for (int j = 0; j < 4 && n + j < N; ++j) dst[j] = ep.accumulate ? dst[j] + v[j] : v[j];
      }
    }
  }
}

template <bool A_KMAJOR, bool B_KMAJOR>
inline int sgemm(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int M,
                 int N, int K, const GemmEpilogue& ep, cudaStream_t st) {
  if (M <= 0 || N <= 0) return LIST_OK;
  dim3 grid((M + 127) / 128, (N + 127) / 128);
  // split K when the output alone cannot fill the GPU and the epilogue is a plain accumulation
  if (ep.accumulate && !ep.bias && !ep.relu && !ep.mask && K >= 2048) {
    const int tiles = grid.x * grid.y;
    int split = (2 * 148 + tiles - 1) / tiles;                    // aim at ~2 CTAs per SM
    const int max_split = K / 1024;
    if (split > max_split) split = max_split;
    if (split > 1) grid.z = split;
  }
  sgemm_kernel<A_KMAJOR, B_KMAJOR><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, ep);
  LIST_LAUNCH_CHECK("sgemm_kernel");
  return LIST_OK;
}

}  // namespace list
