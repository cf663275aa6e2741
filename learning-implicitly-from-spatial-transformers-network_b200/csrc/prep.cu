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

// four consecutive channels as T (16 / 8 bytes)
__device__ __forceinline__ void store4_as(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4_as(__nv_bfloat16* p, float4 v) {
  uint2 u;
  u.x = pack_bf16x2(v.x, v.y);
  u.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}

// [B][C][V] -> [B][V][C] (FWD) or [B][V][C] -> [B][C][V] (the adjoint of the layout change, fp32 -> fp32): a tile of TV
// voxels x C channels per CTA through shared memory [C][TV + 4]; 16-byte accesses along V on the channel-major side and
// along C (when C % 4 == 0) on the channels-last side.
template <typename T, bool FWD>
__global__ void __launch_bounds__(256) prep_volume_kernel(const float* __restrict__ in, int C, int64_t V, int TV, T* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];   // [C][TV + 4]
  const int pitch = TV + 4;
  const int64_t v0 = static_cast<int64_t>(blockIdx.x) * TV;
  const int b = blockIdx.y;
  const int nv = static_cast<int>(min64(TV, V - v0));
  const bool v4 = (V & 3) == 0;                    // tile starts are multiples of 4 voxels then (TV % 4 == 0)
  const bool c4 = (C & 3) == 0;
  const size_t img = static_cast<size_t>(b) * C * V;
  if (FWD) {
    const float* src = in + img;
    if (v4) {
      const int q = TV >> 2;
      for (int i = threadIdx.x; i < C * q; i += blockDim.x) {
        const int c = i / q, v = (i - c * q) * 4;
        if (v < nv) *reinterpret_cast<float4*>(smem + c * pitch + v) = __ldg(reinterpret_cast<const float4*>(src + static_cast<size_t>(c) * V + v0 + v));
      }
    } else {
      for (int i = threadIdx.x; i < C * TV; i += blockDim.x) {
        const int c = i / TV, v = i - c * TV;
        if (v < nv) smem[c * pitch + v] = __ldg(src + static_cast<size_t>(c) * V + v0 + v);
      }
    }
    __syncthreads();
    T* dst = out + (static_cast<size_t>(b) * V + v0) * C;
    if (c4) {
      const int q = C >> 2;
      for (int i = threadIdx.x; i < nv * q; i += blockDim.x) {
        const int v = i / q, c = (i - v * q) * 4;
        const float4 val = make_float4(smem[c * pitch + v], smem[(c + 1) * pitch + v], smem[(c + 2) * pitch + v], smem[(c + 3) * pitch + v]);
        store4_as(dst + static_cast<size_t>(v) * C + c, val);
      }
    } else {
      for (int i = threadIdx.x; i < nv * C; i += blockDim.x) {
        const int v = i / C, c = i - v * C;
        T o;
        from_f32(o, smem[c * pitch + v]);
        dst[static_cast<size_t>(v) * C + c] = o;
      }
    }
  } else {
    const float* src = in + (static_cast<size_t>(b) * V + v0) * C;
    if (c4) {
      const int q = C >> 2;
      for (int i = threadIdx.x; i < nv * q; i += blockDim.x) {
        const int v = i / q, c = (i - v * q) * 4;
        const float4 val = __ldg(reinterpret_cast<const float4*>(src + static_cast<size_t>(v) * C + c));
        smem[c * pitch + v] = val.x; smem[(c + 1) * pitch + v] = val.y; smem[(c + 2) * pitch + v] = val.z; smem[(c + 3) * pitch + v] = val.w;
      }
    } else {
      for (int i = threadIdx.x; i < nv * C; i += blockDim.x) {
        const int v = i / C, c = i - v * C;
        smem[c * pitch + v] = __ldg(src + static_cast<size_t>(v) * C + c);
      }
    }
    __syncthreads();
    float* dst = reinterpret_cast<float*>(out) + img;
    if (v4) {
      const int q = TV >> 2;
      for (int i = threadIdx.x; i < C * q; i += blockDim.x) {
        const int c = i / q, v = (i - c * q) * 4;
        if (v < nv) *reinterpret_cast<float4*>(dst + static_cast<size_t>(c) * V + v0 + v) = *reinterpret_cast<const float4*>(smem + c * pitch + v);
      }
    } else {
      for (int i = threadIdx.x; i < C * TV; i += blockDim.x) {
        const int c = i / TV, v = i - c * TV;
        if (v < nv) dst[static_cast<size_t>(c) * V + v0 + v] = smem[c * pitch + v];
      }
    }
  }
}

// Adjoint of prep_map_kernel for one source map: grad_in[b][c][yi][xi] = sum over the output pixels (y, x) whose bilinear
// taps include (yi, xi) of weight * g[b][y][x][c_off + c] (reference modules.py:25-35 under autograd).  Gather form -- no
// atomics, deterministic: a warp owns one input pixel and 32 channels (lane = channel, so the channels-last gradient is
// read coalesced) and walks the window of output pixels that reference it (about (2 S / H)^2 of them).
constexpr int kMaxWindow = 256;       // output pixels along one axis that can tap one input pixel (<= map_size)
__global__ void __launch_bounds__(256) prep_map_bwd_kernel(const float* __restrict__ g, int Ctot, int c_off, int C, int H, int W, int S,
                                                           float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cblocks = (C + 31) / 32;
  const int b = blockIdx.z / cblocks, cb = blockIdx.z - b * cblocks;
  const int yi = blockIdx.y, xi = blockIdx.x * 8 + warp, c = cb * 32 + lane;
  if (xi >= W) return;                                              // uniform over the warp
  const float scale_h = (S > 1) ? static_cast<float>(H - 1) / static_cast<float>(S - 1) : 0.0f;
  const float scale_w = (S > 1) ? static_cast<float>(W - 1) / static_cast<float>(S - 1) : 0.0f;
  auto window = [&](int i, float scale, int& lo, int& hi) {          // superset of the output indices that tap input index i
    if (scale <= 0.f) { lo = 0; hi = S - 1; return; }
    lo = max(0, static_cast<int>(floorf(static_cast<float>(i - 1) / scale)) - 1);
    hi = min(S - 1, static_cast<int>(ceilf(static_cast<float>(i + 1) / scale)) + 1);
  };
  int ylo, yhi, xlo, xhi;
  window(yi, scale_h, ylo, yhi);
  window(xi, scale_w, xlo, xhi);
  // the x weights of the window are the same for every y: computed once per warp (lanes stride over the window)
  __shared__ float s_wx[8][kMaxWindow];
  if (xhi - xlo + 1 > kMaxWindow) xhi = xlo + kMaxWindow - 1;       // cannot happen for S <= kMaxWindow (checked on the host)
  for (int x = xlo + lane; x <= xhi; x += 32) {
    int x0, x1;
    float lx;
    upsample_axis(x, W, scale_w, x0, x1, lx);
    s_wx[warp][x - xlo] = (x0 == xi ? 1.0f - lx : 0.f) + (x1 == xi ? lx : 0.f);
  }
  __syncwarp();
  int xa = xlo, xb = xhi;                                            // trim the zero-weight ends (uniform over the warp)
  while (xa <= xb && s_wx[warp][xa - xlo] == 0.f) ++xa;
  while (xb >= xa && s_wx[warp][xb - xlo] == 0.f) --xb;
  const float* __restrict__ gb = g + static_cast<size_t>(b) * S * S * Ctot + c_off + c;
  float acc = 0.f;
  if (c < C) {
    for (int y = ylo; y <= yhi; ++y) {
      int y0, y1;
      float ly;
      upsample_axis(y, H, scale_h, y0, y1, ly);
      const float wy = (y0 == yi ? 1.0f - ly : 0.f) + (y1 == yi ? ly : 0.f);
      if (wy == 0.f) continue;
      const float* __restrict__ gr = gb + static_cast<size_t>(y) * S * Ctot;
      float row = 0.f;
#pragma unroll 4
      for (int x = xa; x <= xb; ++x) row = fmaf(s_wx[warp][x - xlo], __ldg(gr + static_cast<size_t>(x) * Ctot), row);
      acc = fmaf(wy, row, acc);
    }
  }
  if (c >= C) return;
  out[((static_cast<size_t>(b) * C + c) * H + yi) * W + xi] = acc;
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

// tile length along V: ~8192 elements per CTA, at most 48 KB of shared memory
static int volume_tile(int C) {
  int tv = 8192 / C;
  if (tv > 4096) tv = 4096;
  if (tv < 64) tv = 64;
  return tv & ~3;
}

int prep_volume(const float* in, int B, int C, int R, void* out, int dtype, cudaStream_t st) {
  const int64_t V = static_cast<int64_t>(R) * R * R;
  const int TV = volume_tile(C);
  const size_t smem = static_cast<size_t>(C) * (TV + 4) * sizeof(float);
  dim3 grid(static_cast<unsigned>((V + TV - 1) / TV), B);
  if (dtype == LIST_F32) {
    prep_volume_kernel<float, true><<<grid, 256, smem, st>>>(in, C, V, TV, static_cast<float*>(out));
  } else {
    prep_volume_kernel<__nv_bfloat16, true><<<grid, 256, smem, st>>>(in, C, V, TV, static_cast<__nv_bfloat16*>(out));
  }
  LIST_LAUNCH_CHECK("prep_volume_kernel");
  return LIST_OK;
}

// Adjoint of prep_volume (fp32): channels-last gradient [B][R^3][C] -> NCDHW [B][C][R^3].
int prep_volume_bwd(const float* g, int B, int C, int R, float* out, cudaStream_t st) {
  const int64_t V = static_cast<int64_t>(R) * R * R;
  const int TV = volume_tile(C);
  const size_t smem = static_cast<size_t>(C) * (TV + 4) * sizeof(float);
  dim3 grid(static_cast<unsigned>((V + TV - 1) / TV), B);
  prep_volume_kernel<float, false><<<grid, 256, smem, st>>>(g, C, V, TV, out);
  LIST_LAUNCH_CHECK("prep_volume_kernel (adjoint)");
  return LIST_OK;
}

// Adjoint of prep_maps (fp32): g [B][S][S][sum ch] -> one NCHW gradient per source map.
int prep_maps_bwd(const float* g, const int32_t* ch, const int32_t* size, int n_maps, int B, int S, float* const* outs, cudaStream_t st) {
  int ctot = 0;
  for (int i = 0; i < n_maps; ++i) ctot += ch[i];
  int off = 0;
  for (int i = 0; i < n_maps; ++i) {
    dim3 grid((size[i] + 7) / 8, size[i], B * ((ch[i] + 31) / 32));
    prep_map_bwd_kernel<<<grid, 256, 0, st>>>(g, ctot, off, ch[i], size[i], size[i], S, outs[i]);
    LIST_LAUNCH_CHECK("prep_map_bwd_kernel");
    off += ch[i];
  }
  return LIST_OK;
}

}  // namespace list
