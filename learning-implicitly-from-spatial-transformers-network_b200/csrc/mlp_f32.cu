// fp32 implicit MLP (a-6, reference network/modules.py:196-201, 276-282) and its backward
// (a-9): the parity path.  Forward keeps H1/H2/H3 in the caller's workspace, which is exactly
// what the backward needs.
#include "sgemm.cuh"

namespace list {

// sdf[r] = (sum_k H3[r][k]*w3[k] + b3) / out_div   -- one warp per row.
__global__ void __launch_bounds__(256) fc_out_kernel(const float* __restrict__ H3, int n2,
                                                     const float* __restrict__ w3,
                                                     const float* __restrict__ b3, float out_div,
                                                     float* __restrict__ sdf, int64_t rows) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float acc = 0.f;
  for (int k = lane; k < n2; k += 32) acc = fmaf(H3[r * n2 + k], __ldg(w3 + k), acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) sdf[r] = __fdiv_rn(acc + __ldg(b3), out_div);
}

size_t mlp_f32_workspace_bytes(const ListWeights* w, int64_t rows) {
  return static_cast<size_t>(rows) * (w->n0 + w->n1 + w->n2) * sizeof(float);
}

int mlp_f32_fwd(const ListWeights* w, const float* X, int64_t ldx, int64_t rows, float* sdf, float out_div,
                float* ws, cudaStream_t st) {
  float* H1 = ws;
  float* H2 = H1 + rows * w->n0;
  float* H3 = H2 + rows * w->n1;
  const int M = static_cast<int>(rows);
  GemmEpilogue ep{};
  ep.relu = 1;
  ep.bias = w->b0;
  int rc = sgemm<true, true>(X, ldx, static_cast<const float*>(w->w0), w->k_pad, H1, w->n0, M, w->n0, w->k_pad, ep, st);
  if (rc) return rc;
  ep.bias = w->b1;
  rc = sgemm<true, true>(H1, w->n0, static_cast<const float*>(w->w1), w->n0, H2, w->n1, M, w->n1, w->n0, ep, st);
  if (rc) return rc;
  ep.bias = w->b2;
  rc = sgemm<true, true>(H2, w->n1, static_cast<const float*>(w->w2), w->n1, H3, w->n2, M, w->n2, w->n1, ep, st);
  if (rc) return rc;
  fc_out_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, st>>>(H3, w->n2, w->w3, w->b3, out_div, sdf, rows);
  LIST_LAUNCH_CHECK("fc_out_kernel");
  return LIST_OK;
}

// ---------------------------------------------------------------- backward
// dZ3[r][k] = d_sdf[r] * w3[k] * (H3[r][k] > 0);  d_w3[k] += sum_r d_sdf[r]*H3[r][k];  d_b3 += sum_r d_sdf[r]
__global__ void __launch_bounds__(256) fc_out_bwd_kernel(const float* __restrict__ H3, int n2,
                                                         const float* __restrict__ w3,
                                                         const float* __restrict__ d_sdf, int64_t rows,
                                                         float* __restrict__ dZ3, float* __restrict__ d_w3,
                                                         float* __restrict__ d_b3) {
  // block = 256 columns (n2 <= 256 per y-block) x a slab of rows
  const int k = blockIdx.y * 256 + threadIdx.x;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 64;
  const int64_t r1 = min64(rows, r0 + 64);
  if (k >= n2) return;
  const float wk = __ldg(w3 + k);
  float gw = 0.f, gb = 0.f;
  for (int64_t r = r0; r < r1; ++r) {
    const float g = __ldg(d_sdf + r);
    const float h = H3[r * n2 + k];
    dZ3[r * n2 + k] = h > 0.f ? g * wk : 0.f;
    gw = fmaf(g, h, gw);
    gb += g;
  }
  if (d_w3) atomicAdd(d_w3 + k, gw);
  if (d_b3 && k == 0) atomicAdd(d_b3, gb);
}

// d_b[n] += sum_r dZ[r][n]
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dZ, int n, int64_t rows,
                                                     float* __restrict__ d_b) {
  const int k = blockIdx.y * 256 + threadIdx.x;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 128;
  const int64_t r1 = min64(rows, r0 + 128);
  if (k >= n) return;
  float s = 0.f;
  for (int64_t r = r0; r < r1; ++r) s += dZ[r * n + k];
  atomicAdd(d_b + k, s);
}

size_t mlp_f32_bwd_workspace_bytes(const ListWeights* w, int64_t rows) {
  // dZ0 [rows][n0], dZ1 [rows][n1], dZ2 [rows][n2], dX [rows][k_pad]
  return static_cast<size_t>(rows) * (w->n0 + w->n1 + w->n2 + w->k_pad) * sizeof(float);
}

// Returns dX (fp32 [rows][k_pad]) at the start of `ws`.
int mlp_f32_bwd(const ListWeights* w, const float* X, int64_t ldx, int64_t rows, const float* fwd_ws,
                const float* d_sdf, const ListGrads* g, float* ws, cudaStream_t st) {
  const float* H1 = fwd_ws;
  const float* H2 = H1 + rows * w->n0;
  const float* H3 = H2 + rows * w->n1;
  float* dX = ws;
  float* dZ0 = dX + rows * w->k_pad;
  float* dZ1 = dZ0 + rows * w->n0;
  float* dZ2 = dZ1 + rows * w->n1;
  const int M = static_cast<int>(rows);
  int rc;
  {
    dim3 grid(static_cast<unsigned>((rows + 63) / 64), (w->n2 + 255) / 256);
    fc_out_bwd_kernel<<<grid, 256, 0, st>>>(H3, w->n2, w->w3, d_sdf, rows, dZ2, g->d_w3, g->d_b3);
    LIST_LAUNCH_CHECK("fc_out_bwd_kernel");
  }
  auto colsum = [&](const float* dZ, int n, float* db) -> int {
    if (!db) return LIST_OK;
    dim3 grid(static_cast<unsigned>((rows + 127) / 128), (n + 255) / 256);
    colsum_kernel<<<grid, 256, 0, st>>>(dZ, n, rows, db);
    LIST_LAUNCH_CHECK("colsum_kernel");
    return LIST_OK;
  };
  GemmEpilogue acc{};
  acc.accumulate = 1;
  GemmEpilogue mask{};
  // layer 2: d_w2[n][k] += sum_r dZ2[r][n] H2[r][k];  dZ1 = (dZ2 · W2) * (H2 > 0)
  if ((rc = colsum(dZ2, w->n2, g->d_b2))) return rc;
  if (g->d_w2 && (rc = sgemm<false, false>(dZ2, w->n2, H2, w->n1, g->d_w2, w->n1, w->n2, w->n1, M, acc, st))) return rc;
  mask.mask = H2; mask.ldmask = w->n1;
  if ((rc = sgemm<true, false>(dZ2, w->n2, static_cast<const float*>(w->w2), w->n1, dZ1, w->n1, M, w->n1, w->n2, mask, st))) return rc;
  // layer 1
  if ((rc = colsum(dZ1, w->n1, g->d_b1))) return rc;
  if (g->d_w1 && (rc = sgemm<false, false>(dZ1, w->n1, H1, w->n0, g->d_w1, w->n0, w->n1, w->n0, M, acc, st))) return rc;
  mask.mask = H1; mask.ldmask = w->n0;
  if ((rc = sgemm<true, false>(dZ1, w->n1, static_cast<const float*>(w->w1), w->n0, dZ0, w->n0, M, w->n0, w->n1, mask, st))) return rc;
  // layer 0
  if ((rc = colsum(dZ0, w->n0, g->d_b0))) return rc;
  if (g->d_w0 && (rc = sgemm<false, false>(dZ0, w->n0, X, ldx, g->d_w0, w->k_pad, w->n0, w->k_pad, M, acc, st))) return rc;
  GemmEpilogue none{};
  if ((rc = sgemm<true, false>(dZ0, w->n0, static_cast<const float*>(w->w0), w->k_pad, dX, w->k_pad, M, w->k_pad, w->n0, none, st))) return rc;
  return LIST_OK;
}

}  // namespace list
