// Shared device/host helpers for the list_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/list_b200.h"

namespace list {

// ---- error reporting (thread-local string, SURVEY.md §8b "Error convention") ----
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define LIST_CHECK_ARG(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      ::list::set_error(__VA_ARGS__);             \
      return LIST_EINVAL;                         \
    }                                             \
  } while (0)

#define LIST_CUDA(call)                                         \
  do {                                                          \
    cudaError_t e__ = (call);                                   \
    if (e__ != cudaSuccess) return ::list::cuda_fail(e__, #call); \
  } while (0)

#define LIST_LAUNCH_CHECK(name)                                         \
  do {                                                                  \
    cudaError_t e__ = cudaGetLastError();                               \
    if (e__ != cudaSuccess) return ::list::cuda_fail(e__, "launch " name); \
  } while (0)

__host__ __device__ __forceinline__ int64_t min64(int64_t a, int64_t b) { return a < b ? a : b; }

constexpr float kDisplacement = 0.0722f;   // reference network/modules.py:205

// ---- packed fp32 pairs (sm_100 FFMA2 / FADD2) ----
// Two IEEE round-to-nearest operations per instruction, each lane bit-identical to the scalar fmaf / +.  Measured on
// B200 (scripts/microbench/ffma2.cu): FFMA issues one warp instruction per cycle and scheduler (127 FMA/clk/SM), FFMA2
// one per 2.17 cycles (118 FMA/clk/SM) -- the packed form saves issue slots, not pipe time.  It pays in the MLP
// epilogues (one warp per scheduler, issue bound); in the gather kernels it was neutral to slower
// (profiles/r01_exp_packed_fp32_and_addend_variants.txt).  ptxas folds a {s, s} operand into the scalar-broadcast form.
__device__ __forceinline__ unsigned long long f2_pack(float2 a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ float2 f2_unpack(unsigned long long r) {
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(r));
  return d;
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {        // a * b + c
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)), "l"(f2_pack(c)));
  return f2_unpack(d);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(d);
}
// two bf16 (one 32-bit word, low half first) -> fp32 pair
__device__ __forceinline__ float2 bf16x2_to_f2(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}

// ---- 8-wide channel vectors (16 B of bf16 / 32 B of fp32) ----
__device__ __forceinline__ void load8(const float* __restrict__ p, float v[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* __restrict__ p, float v[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void store8(float* __restrict__ p, const float v[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);   // .x = lo (low 16 bits)
  return *reinterpret_cast<const uint32_t*>(&h);
}
// max(x, 0) folded into the conversion (F2FP.RELU): negative -> +0, NaN stays NaN like torch's relu
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ void store8(__nv_bfloat16* __restrict__ p, const float v[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ float to_f32(float x) { return x; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 x) { return __bfloat162float(x); }
__device__ __forceinline__ void from_f32(float& d, float x) { d = x; }
__device__ __forceinline__ void from_f32(__nv_bfloat16& d, float x) { d = __float2bfloat16_rn(x); }

// Displacement d of the reference's table (modules.py:205-212): 0 = none, then
// (-x,+x,-y,+y,-z,+z) in the swapped/scaled frame, component 0 -> W.
__device__ __forceinline__ void displaced(const float q[3], int d, float out[3]) {
  out[0] = q[0]; out[1] = q[1]; out[2] = q[2];
  if (d > 0) {
    const int axis = (d - 1) >> 1;
    const float s = ((d - 1) & 1) ? kDisplacement : -kDisplacement;
    out[axis] = q[axis] + s;
  }
}

// ATen grid_sampler_compute_source_index, align_corners=True, padding 'border'
// (GridSampler.h): ((c+1)/2)*(R-1) clipped to [0,R-1]; corner i0=floor, i1=min(i0+1,R-1);
// weights w0=(i0+1)-i, w1=i-i0 (a corner index == R only occurs with weight 0).
struct Axis3 {
  int i0, i1;
  float w0, w1;
};
__device__ __forceinline__ Axis3 axis_border(float c, int R) {
  float i = ((c + 1.0f) * 0.5f) * static_cast<float>(R - 1);
  i = fminf(fmaxf(i, 0.0f), static_cast<float>(R - 1));
  const float f = floorf(i);
  Axis3 a;
  a.i0 = static_cast<int>(f);
  a.i1 = min(a.i0 + 1, R - 1);
  a.w1 = i - f;
  a.w0 = (f + 1.0f) - i;
  return a;
}

// reference modules.py:37-47 for one point: h=[q,1]·T (k-sequential FMA), perspective divide,
// clamp to [0,S-1] keeping NaN, normalise and un-normalise exactly as the reference +
// grid_sample do.  Returns pixel coordinates (ix -> W, iy -> H); NaN means "no tap".
__device__ __forceinline__ void localise(const float q[3], const float* __restrict__ T, int S,
                                         float& ix, float& iy, float h_out[3]) {
  float h[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float acc = __fmul_rn(q[0], T[0 * 3 + j]);
    acc = __fmaf_rn(q[1], T[1 * 3 + j], acc);
    acc = __fmaf_rn(q[2], T[2 * 3 + j], acc);
    h[j] = __fadd_rn(acc, T[3 * 3 + j]);
    h_out[j] = h[j];
  }
  const float den = __fadd_rn(h[2], 1e-8f);
  float x = __fdiv_rn(h[0], den);
  float y = __fdiv_rn(h[1], den);
  const float lim = static_cast<float>(S - 1);
  x = (x != x) ? x : fminf(fmaxf(x, 0.0f), lim);
  y = (y != y) ? y : fminf(fmaxf(y, 0.0f), lim);
  const float half = lim * 0.5f;
  const float gx = __fdiv_rn(__fsub_rn(x, half), half);
  const float gy = __fdiv_rn(__fsub_rn(y, half), half);
  ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.0f), 0.5f), lim);
  iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.0f), 0.5f), lim);
}

// a-8 axis value: numpy.linspace(lo, hi, res)[i] in float64 then rounded to fp32
// (reference utils.py:87, executors.py:194): i*step + lo, last point exactly hi.
__device__ __forceinline__ float linspace_f32(int i, int res, double lo, double hi) {
  if (res == 1) return static_cast<float>(lo);
  const double step = (hi - lo) / static_cast<double>(res - 1);
  const double v = (i == res - 1) ? hi : __dadd_rn(__dmul_rn(static_cast<double>(i), step), lo);
  return static_cast<float>(v);
}
// same value with the (double) step precomputed on the host: (hi - lo) / (res - 1) is an IEEE division on both sides
__device__ __forceinline__ float linspace_f32_step(int i, int res, double lo, double hi, double step) {
  if (res == 1) return static_cast<float>(lo);
  const double v = (i == res - 1) ? hi : __dadd_rn(__dmul_rn(static_cast<double>(i), step), lo);
  return static_cast<float>(v);
}

}  // namespace list
