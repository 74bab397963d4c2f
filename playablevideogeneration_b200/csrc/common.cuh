// Shared helpers for the pvg_b200 kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/pvg_b200.h"

namespace pvg {

void set_error(const std::string& msg);

#define PVG_CHECK_ARG(cond, msg)                                                     \
  do {                                                                               \
    if (!(cond)) { pvg::set_error(std::string(__func__) + ": " + (msg)); return -1; } \
  } while (0)

#define PVG_CUDA_OK(expr)                                                                            \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      pvg::set_error(std::string(__func__) + ": " #expr " -> " + cudaGetErrorString(_e));            \
      return -2;                                                                                     \
    }                                                                                                \
  } while (0)

#define PVG_LAUNCH_OK() PVG_CUDA_OK(cudaPeekAtLastError())

constexpr int kSMs = 148;   // B200: 2 dies x 74 SMs; grids are sized in multiples of this

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// grid for a grid-stride elementwise kernel: enough CTAs for full occupancy, never more than the work
inline int ew_grid(int64_t work_items, int threads) {
  int64_t need = ceil_div64(work_items, threads);
  int64_t cap = (int64_t)kSMs * 8;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// tf32 "hi" part of an fp32 value: round-to-nearest (ties away) to 10 explicit mantissa bits.  hi is exactly
// representable in tf32, so the tensor core's own conversion of hi is exact whatever its rounding mode, and
// lo = x - hi is exact in fp32 (it has <= 14 significant bits).
__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// fp16 pair / scalar from fp32, round to nearest, saturating at +-65504 in ONE instruction (F2FP.SATFINITE...PACK_AB) - the
// conversion every producer of fp16 operand planes uses (finite inputs: identical to clamp + cvt.rn)
__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float f16_round_sat(float v) {         // the value f16(v) carries, as fp32
  const uint32_t r = pack_f16x2_sat(v, 0.f);
  return __half2float(__ushort_as_half((unsigned short)(r & 0xffffu)));
}

__device__ __forceinline__ float act_fwd(float v, int act, float slope) {
  switch (act) {
    case PVG_ACT_LRELU:   return v > 0.f ? v : v * slope;
    case PVG_ACT_RELU:    return v > 0.f ? v : 0.f;
    case PVG_ACT_TANH:    return tanhf(v);
    case PVG_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default:              return v;
  }
}
// derivative expressed through the activation OUTPUT y (all activations on the path are invertible in sign)
__device__ __forceinline__ float act_bwd_from_out(float y, int act, float slope) {
  switch (act) {
    case PVG_ACT_LRELU:   return y > 0.f ? 1.f : slope;
    case PVG_ACT_RELU:    return y > 0.f ? 1.f : 0.f;
    case PVG_ACT_TANH:    return 1.f - y * y;
    case PVG_ACT_SIGMOID: return y * (1.f - y);
    default:              return 1.f;
  }
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void stg4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

}  // namespace pvg
