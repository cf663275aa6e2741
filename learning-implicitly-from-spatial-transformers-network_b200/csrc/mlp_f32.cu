// fp32 implicit MLP (a-6, reference network/modules.py:196-201, 276-282) and its backward
// (a-9): the parity path.  Forward keeps H1/H2/H3 in the caller's workspace, which is exactly
// what the backward needs.
#include <cstdlib>

#include "sgemm.cuh"

namespace list {

// LIST_B200_F32_TC=0 selects the FFMA GEMMs (sgemm.cuh) instead of the 3xTF32 tensor-core ones (tgemm.cu): A/B aid.
static bool f32_tc() {
  const char* e = getenv("LIST_B200_F32_TC");
  return !(e && e[0] == '0');
}
// floats of the lo(w) copies of the three weight matrices the tensor-core GEMMs read next to the forward's activations
// (lo of the activations themselves is computed inside tgemm_kernel)
static size_t fwd_lo_floats(const ListWeights* w, int64_t) {
  return static_cast<size_t>(w->n0) * w->k_pad + static_cast<size_t>(w->n1) * w->n0 + static_cast<size_t>(w->n2) * w->n1;
}

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
  // H1 H2 H3 | lo(W0) lo(W1) lo(W2)
  return (static_cast<size_t>(rows) * (w->n0 + w->n1 + w->n2) + fwd_lo_floats(w, rows)) * sizeof(float);
}

// exact != 0: FFMA GEMMs (round-to-nearest fp32 accumulation).  The forward of a training step uses it: the tensor cores
// accumulate with truncation, which over K = 3648 leaves a relative error of ~3e-5 in the pre-activations -- inside the
// 1e-4 SDF bound, but enough to flip the ReLU mask of units that sit within 1e-5 of zero, and a flipped mask changes a
// gradient entry by far more than the 1e-3 bound of the backward parity gate.  The backward GEMMs have no such decisions
// and run on the tensor cores either way.
int mlp_f32_fwd(const ListWeights* w, const float* X, int64_t ldx, int64_t rows, float* sdf, float out_div,
                float* ws, int exact, cudaStream_t st) {
  float* H1 = ws;
  float* H2 = H1 + rows * w->n0;
  float* H3 = H2 + rows * w->n1;
  const int M = static_cast<int>(rows);
  GemmEpilogue ep{};
  ep.relu = 1;
  ep.bias = w->b0;
  int rc;
  const float* W0 = static_cast<const float*>(w->w0);
  const float* W1 = static_cast<const float*>(w->w1);
  const float* W2 = static_cast<const float*>(w->w2);
  if (!exact && f32_tc() && ldx == w->k_pad) {
    float* W0l = H3 + rows * w->n2;
    float* W1l = W0l + static_cast<size_t>(w->n0) * w->k_pad;
    float* W2l = W1l + static_cast<size_t>(w->n1) * w->n0;
    // lo of the activations (X, H1, H2) is computed inside tgemm_kernel from the operand box in shared memory
    if ((rc = split_lo(W0, W0l, static_cast<int64_t>(w->n0) * w->k_pad, st))) return rc;
    if ((rc = split_lo(W1, W1l, static_cast<int64_t>(w->n1) * w->n0, st))) return rc;
    if ((rc = split_lo(W2, W2l, static_cast<int64_t>(w->n2) * w->n1, st))) return rc;
    if ((rc = tgemm(X, nullptr, ldx, W0, W0l, w->k_pad, H1, w->n0, M, w->n0, w->k_pad, ep, st))) return rc;
    ep.bias = w->b1;
    if ((rc = tgemm(H1, nullptr, w->n0, W1, W1l, w->n0, H2, w->n1, M, w->n1, w->n0, ep, st))) return rc;
    ep.bias = w->b2;
    if ((rc = tgemm(H2, nullptr, w->n1, W2, W2l, w->n1, H3, w->n2, M, w->n2, w->n1, ep, st))) return rc;
  } else {
    rc = sgemm<true, true>(X, ldx, W0, w->k_pad, H1, w->n0, M, w->n0, w->k_pad, ep, st);
    if (rc) return rc;
    ep.bias = w->b1;
    rc = sgemm<true, true>(H1, w->n0, W1, w->n0, H2, w->n1, M, w->n1, w->n0, ep, st);
    if (rc) return rc;
    ep.bias = w->b2;
    rc = sgemm<true, true>(H2, w->n1, W2, w->n1, H3, w->n2, M, w->n2, w->n1, ep, st);
    if (rc) return rc;
  }
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

static int64_t pad4(int64_t x) { return (x + 3) / 4 * 4; }

size_t mlp_f32_bwd_workspace_bytes(const ListWeights* w, int64_t rows) {
  // dX [rows][k_pad], dZ0 [rows][n0], dZ1 [rows][n1], dZ2 [rows][n2] | tensor-core extras, reused layer by layer:
  // lo(dZ) [rows][n0], dZ^T and lo [n0][ldr] each, X^T and lo [k_pad][ldr] each, W^T and lo [k_pad][n0] each
  const size_t ldr = static_cast<size_t>(pad4(rows));
  const size_t extra = static_cast<size_t>(rows) * w->n0 + 2 * static_cast<size_t>(w->n0) * ldr + 2 * static_cast<size_t>(w->k_pad) * ldr +
                       2 * static_cast<size_t>(w->k_pad) * w->n0;
  return (static_cast<size_t>(rows) * (w->k_pad + w->n0 + w->n1 + w->n2) + extra) * sizeof(float);
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
  GemmEpilogue none{};
  const float* W0 = static_cast<const float*>(w->w0);
  const float* W1 = static_cast<const float*>(w->w1);
  const float* W2 = static_cast<const float*>(w->w2);
  if (f32_tc() && w->n0 >= w->n1 && w->n0 >= w->n2 && w->k_pad >= w->n0) {
    // All GEMMs are K-major x K-major on the tensor cores; operands that come the other way round are transposed
    // (transpose_split writes x^T and lo(x^T) in one pass).  Per layer (n outputs, k inputs, Xin = the layer's input):
    //   d_w[n][k] += sum_r dZ[r][n] Xin[r][k]   =  dZ^T [n][rows] . (Xin^T [k][rows])^T
    //   dPrev[r][k] = sum_n dZ[r][n] W[n][k]     =  dZ [rows][n]  . (W^T [k][n])^T         (then the ReLU mask of Xin)
    const int64_t ldr = pad4(rows);
    float* dZl = dZ2 + rows * w->n2;
    float* dZT = dZl + rows * w->n0;
    float* dZTl = dZT + static_cast<size_t>(w->n0) * ldr;
    float* XT = dZTl + static_cast<size_t>(w->n0) * ldr;
    float* XTl = XT + static_cast<size_t>(w->k_pad) * ldr;
    float* WT = XTl + static_cast<size_t>(w->k_pad) * ldr;
    float* WTl = WT + static_cast<size_t>(w->k_pad) * w->n0;
    auto layer = [&](const float* dZ, int n, const float* Xin, int64_t ldin, int k, const float* W, float* dW, float* db,
                     float* dPrev, const GemmEpilogue& epPrev) -> int {
      int r;
      if ((r = colsum(dZ, n, db))) return r;
      if (dW) {
        if ((r = transpose_split(dZ, n, M, n, dZT, dZTl, ldr, st))) return r;
        if ((r = transpose_split(Xin, ldin, M, k, XT, XTl, ldr, st))) return r;
        if ((r = tgemm(dZT, dZTl, ldr, XT, XTl, ldr, dW, k, n, k, M, acc, st))) return r;
      }
      if ((r = transpose_split(W, k, n, k, WT, WTl, n, st))) return r;
      return tgemm(dZ, nullptr, n, WT, WTl, n, dPrev, k, M, k, n, epPrev, st);   // lo(dZ) in the kernel
    };
    mask.mask = H2; mask.ldmask = w->n1;
    if ((rc = layer(dZ2, w->n2, H2, w->n1, w->n1, W2, g->d_w2, g->d_b2, dZ1, mask))) return rc;
    mask.mask = H1; mask.ldmask = w->n0;
    if ((rc = layer(dZ1, w->n1, H1, w->n0, w->n0, W1, g->d_w1, g->d_b1, dZ0, mask))) return rc;
    return layer(dZ0, w->n0, X, ldx, w->k_pad, W0, g->d_w0, g->d_b0, dX, none);
  }
  // layer 2: d_w2[n][k] += sum_r dZ2[r][n] H2[r][k];  dZ1 = (dZ2 · W2) * (H2 > 0)
  if ((rc = colsum(dZ2, w->n2, g->d_b2))) return rc;
  if (g->d_w2 && (rc = sgemm<false, false>(dZ2, w->n2, H2, w->n1, g->d_w2, w->n1, w->n2, w->n1, M, acc, st))) return rc;
  mask.mask = H2; mask.ldmask = w->n1;
  if ((rc = sgemm<true, false>(dZ2, w->n2, W2, w->n1, dZ1, w->n1, M, w->n1, w->n2, mask, st))) return rc;
  // layer 1
  if ((rc = colsum(dZ1, w->n1, g->d_b1))) return rc;
  if (g->d_w1 && (rc = sgemm<false, false>(dZ1, w->n1, H1, w->n0, g->d_w1, w->n0, w->n1, w->n0, M, acc, st))) return rc;
  mask.mask = H1; mask.ldmask = w->n0;
  if ((rc = sgemm<true, false>(dZ1, w->n1, W1, w->n0, dZ0, w->n0, M, w->n0, w->n1, mask, st))) return rc;
  // layer 0
  if ((rc = colsum(dZ0, w->n0, g->d_b0))) return rc;
  if (g->d_w0 && (rc = sgemm<false, false>(dZ0, w->n0, X, ldx, g->d_w0, w->k_pad, w->n0, w->k_pad, M, acc, st))) return rc;
  if ((rc = sgemm<true, false>(dZ0, w->n0, W0, w->k_pad, dX, w->k_pad, M, w->k_pad, w->n0, none, st))) return rc;
  return LIST_OK;
}

}  // namespace list
