// Interface of the fp32 GEMM families (sgemm.cuh: FFMA pipe; tgemm.cu: 3xTF32 on the tensor cores).
#pragma once
#include "common.cuh"

namespace list {

struct GemmEpilogue {
  const float* bias;      // [N] added before the activation, or nullptr
  int relu;               // max(x,0)
  const float* mask;      // [M][ldmask]: result *= (mask > 0), or nullptr (ReLU backward)
  int64_t ldmask;
  int accumulate;         // C += result
};

// lo[i] = tf32_round(x[i] - upper_19_bits(x[i])); n a multiple of 4, 16-byte aligned pointers
int split_lo(const float* x, float* lo, int64_t n, cudaStream_t st);
// xT[c][r] = x[r][c] and loT = lo(xT) for x [rows][cols] with row pitch ld; row pitch ldt (>= rows, multiple of 4; the pad
// columns are written as zeros) of both outputs
int transpose_split(const float* x, int64_t ld, int rows, int cols, float* xT, float* loT, int64_t ldt, cudaStream_t st);

// C[m][n] (+)= epi(sum_k A[m*lda + k] B[n*ldb + k]) with fp32-level accuracy on the tensor cores; Alo / Blo = split_lo of
// A / B (same layout and leading dimension); Alo == nullptr: lo(A) is computed inside the kernel (activations).  Operands in
// another orientation go through transpose_split first.
int tgemm(const float* A, const float* Alo, int64_t lda, const float* B, const float* Blo, int64_t ldb, float* C, int64_t ldc, int M,
          int N, int K, const GemmEpilogue& ep, cudaStream_t st);

}  // namespace list
