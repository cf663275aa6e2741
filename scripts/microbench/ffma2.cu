// Microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) issue rate and latency on one SM partition.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2 ffma2.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float fma1(float a, float b, float c) {
  float d;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

template <int CH, bool PACKED>
__global__ void bench(float* out, long long* cycles, int iters, const float* wsrc) {
  float s[CH], xs[CH];
  unsigned long long p[CH], xp[CH];
  const float w = wsrc[threadIdx.x];                   // per-thread register operand (3-register form), like a weight from LDS
  const unsigned long long wp = (static_cast<unsigned long long>(__float_as_uint(w)) << 32) | __float_as_uint(w);
#pragma unroll
  for (int i = 0; i < CH; ++i) { s[i] = threadIdx.x + i; p[i] = (static_cast<unsigned long long>(__float_as_uint(s[i])) << 32) | __float_as_uint(s[i] + 1.f);
    xs[i] = wsrc[threadIdx.x + 32 * i]; xp[i] = (static_cast<unsigned long long>(__float_as_uint(xs[i])) << 32) | __float_as_uint(xs[i] * 0.5f); }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        if (PACKED) p[i] = fma2(xp[i], wp, p[i]);
        else s[i] = fma1(xs[i], w, s[i]);
      }
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) acc += PACKED ? __uint_as_float(static_cast<unsigned>(p[i])) + __uint_as_float(static_cast<unsigned>(p[i] >> 32)) : s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int CH, bool PACKED>
void run(const char* name, int threads, float* out, long long* cyc) {
  const int iters = 4096;
  bench<CH, PACKED><<<148, threads>>>(out, cyc, iters, out + 148 * 1024);
  bench<CH, PACKED><<<148, threads>>>(out, cyc, iters, out + 148 * 1024);
  cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  const double instr_per_warp = static_cast<double>(iters) * 8 * CH;
  const double warps_per_smsp = threads / 32 / 4.0;
  const double fma_lanes = instr_per_warp * (threads / 32) * 32 * (PACKED ? 2 : 1) / c;
  printf("%-28s threads %4d chains %d: %.0f cycles, %.2f cyc/instr/warp, %.2f cyc/instr/SMSP, %.1f fp32 FMA/clk/SM\n", name, threads, CH,
         static_cast<double>(c), c / instr_per_warp, c / (instr_per_warp * (warps_per_smsp < 1 ? 1 : warps_per_smsp)), fma_lanes);
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, (148 * 1024 + 2048) * sizeof(float));
  cudaMemset(out, 0, (148 * 1024 + 2048) * sizeof(float));
  cudaMalloc(&cyc, sizeof(long long));
  run<1, false>("FFMA  latency (1 chain)", 32, out, cyc);
  run<1, true>("FFMA2 latency (1 chain)", 32, out, cyc);
  run<8, false>("FFMA  8 chains", 128, out, cyc);
  run<8, true>("FFMA2 8 chains", 128, out, cyc);
  run<8, false>("FFMA  8 chains", 512, out, cyc);
  run<8, true>("FFMA2 8 chains", 512, out, cyc);
  run<8, false>("FFMA  8 chains", 1024, out, cyc);
  run<8, true>("FFMA2 8 chains", 1024, out, cyc);
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
