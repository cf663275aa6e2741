// Per-image layout kernels (run once per image, hoisted out of the chunk loop):
//   prep_maps   : a-1, reference network/modules.py:25-35 -- bilinear align_corners=True
//                 upsample of the 5 ResNet maps to map_size^2, fused with the NCHW ->
//                 channels-last transpose, the channel concat of modules.py:53 and the
//                 optional bf16 conversion.
//   prep_volume : NCDHW -> NDHWC (+ bf16) for the voxel-encoder pyramid
//                 (reference modules.py:425-442 outputs, consumed at modules.py:264-265).
// Both are HBM-bound transposes staged through shared memory so that reads (along W / along
// voxels) and writes (along C) are both coalesced.
#include "common.cuh"

namespace list {

// ATen UpSample.h area_pixel_compute_scale / guard_index_and_lambda, align_corners=True.
__device__ __forceinline__ void upsample_axis(int dst, int in_size, float scale, int& i0, int& i1,
                                              float& lam) {
  const float src = __fmul_rn(scale, static_cast<float>(dst));
  i0 = min(static_cast<int>(src), in_size - 1);
  lam = fminf(fmaxf(__fsub_rn(src, static_cast<float>(i0)), 0.0f), 1.0f);
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
}

// One CTA = one output row y of one image for a 32-channel slab of one source map.
// Stage the two source rows (y0, y1) of those 32 channels in shared memory with coalesced reads
// along W, then every thread produces (x, c) outputs with c fastest so the channels-last store
// is coalesced.
template <typename T>
__global__ void __launch_bounds__(256) prep_map_kernel(const float* __restrict__ in, int C, int H,
                                                       int W, int S, T* __restrict__ out, int Ctot,
                                                       int c_off) {
  extern __shared__ float smem[];   // [2][32][W+1]
  const int y = blockIdx.x;
  const int c0 = blockIdx.y * 32;
  const int b = blockIdx.z;
  const int Wp = W + 1;
  const float scale_h = (S > 1) ? static_cast<float>(H - 1) / static_cast<float>(S - 1) : 0.0f;
  const float scale_w = (S > 1) ? static_cast<float>(W - 1) / static_cast<float>(S - 1) : 0.0f;
  int y0, y1;
  float ly;
  upsample_axis(y, H, scale_h, y0, y1, ly);
  const int nch = min(32, C - c0);
  for (int i = threadIdx.x; i < 2 * nch * W; i += blockDim.x) {
    const int x = i % W;
    const int c = (i / W) % nch;
    const int r = i / (W * nch);
    const int ys = r ? y1 : y0;
    smem[(r * 32 + c) * Wp + x] = __ldg(in + ((static_cast<size_t>(b) * C + c0 + c) * H + ys) * W + x);
  }
  __syncthreads();
  const float hy0 = 1.0f - ly;
  for (int i = threadIdx.x; i < S * 32; i += blockDim.x) {
    const int c = i & 31;
    const int x = i >> 5;
    if (c >= nch) continue;
    int x0, x1;
    float lx;
    upsample_axis(x, W, scale_w, x0, x1, lx);
    const float hx0 = 1.0f - lx;
    const float* r0 = smem + (0 * 32 + c) * Wp;
    const float* r1 = smem + (1 * 32 + c) * Wp;
    // ATen upsample_bilinear2d (CUDA): h0*(w0*v00 + w1*v01) + h1*(w0*v10 + w1*v11)
    const float v = hy0 * (hx0 * r0[x0] + lx * r0[x1]) + ly * (hx0 * r1[x0] + lx * r1[x1]);
    T o;
    from_f32(o, v);
    out[((static_cast<size_t>(b) * S + y) * S + x) * Ctot + c_off + c0 + c] = o;
  }
}

// [B][C][V] -> [B][V][C]: 64 voxels x C channels per CTA through shared memory.
template <typename T>
__global__ void __launch_bounds__(256) prep_volume_kernel(const float* __restrict__ in, int C,
                                                          int64_t V, T* __restrict__ out) {
  extern __shared__ float smem[];   // [C][65]
  const int64_t v0 = static_cast<int64_t>(blockIdx.x) * 64;
  const int b = blockIdx.y;
  const int nv = static_cast<int>(min64(64, V - v0));
  const float* src = in + static_cast<size_t>(b) * C * V;
  for (int i = threadIdx.x; i < C * 64; i += blockDim.x) {
    const int v = i & 63;
    const int c = i >> 6;
    if (v < nv) smem[c * 65 + v] = __ldg(src + static_cast<size_t>(c) * V + v0 + v);
  }
  __syncthreads();
  T* dst = out + (static_cast<size_t>(b) * V + v0) * C;
  for (int i = threadIdx.x; i < C * 64; i += blockDim.x) {
    const int c = i % C;
    const int v = i / C;
    if (v < nv) {
      T o;
      from_f32(o, smem[c * 65 + v]);
      dst[static_cast<size_t>(v) * C + c] = o;
    }
  }
}

template <typename T>
static int launch_prep_maps(const float* const* maps, const int32_t* ch, const int32_t* size,
                            int n_maps, int B, int S, T* out, cudaStream_t st) {
  int ctot = 0;
  for (int i = 0; i < n_maps; ++i) ctot += ch[i];
  int off = 0;
  for (int i = 0; i < n_maps; ++i) {
    const size_t smem = static_cast<size_t>(2) * 32 * (size[i] + 1) * sizeof(float);
    if (smem > 48 * 1024) {
      LIST_CUDA(cudaFuncSetAttribute(prep_map_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
    }
    dim3 grid(S, (ch[i] + 31) / 32, B);
    prep_map_kernel<T><<<grid, 256, smem, st>>>(maps[i], ch[i], size[i], size[i], S, out, ctot, off);
    LIST_LAUNCH_CHECK("prep_map_kernel");
    off += ch[i];
  }
  return LIST_OK;
}

int prep_maps(const float* const* maps, const int32_t* ch, const int32_t* size, int n_maps, int B,
              int S, void* out, int dtype, cudaStream_t st) {
  if (dtype == LIST_F32) return launch_prep_maps<float>(maps, ch, size, n_maps, B, S, static_cast<float*>(out), st);
  return launch_prep_maps<__nv_bfloat16>(maps, ch, size, n_maps, B, S, static_cast<__nv_bfloat16*>(out), st);
}

int prep_volume(const float* in, int B, int C, int R, void* out, int dtype, cudaStream_t st) {
  const int64_t V = static_cast<int64_t>(R) * R * R;
  const size_t smem = static_cast<size_t>(C) * 65 * sizeof(float);
  dim3 grid(static_cast<unsigned>((V + 63) / 64), B);
  if (dtype == LIST_F32) {
    prep_volume_kernel<float><<<grid, 256, smem, st>>>(in, C, V, static_cast<float*>(out));
  } else {
    prep_volume_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(in, C, V, static_cast<__nv_bfloat16*>(out));
  }
  LIST_LAUNCH_CHECK("prep_volume_kernel");
  return LIST_OK;
}

}  // namespace list
