// Direct fp32 convolutions for the image-facing layers, where one side has 3 (or 12) channels and an implicit GEMM wastes
// its tile: these layers are HBM-bound (a 3->64 stem writes 256 B per pixel for 1.7 kFMA), so the kernels are organised
// around coalesced NHWC traffic - one thread per output pixel, weights broadcast from shared memory.
//   conv_small_cin  : Cin in {3, 12}, any Cout % 4 == 0  (E stem 3S->16, VGG conv1_1 3->64, data gradients of the 3-channel
//                     tanh heads: dY(3) -> 32/64/128 channels, 3x3 and 7x7)
//   conv_small_cout : Cout <= 4, Cin % 4 == 0             (tanh heads 128/64/32 -> 3 incl. the 7x7, data gradients of the
//                     stems: 16 -> 3S, 64 -> 3)
// Weights arrive in the common pack [Cout][R][S][Cin] (pvg_pack_conv_weight, unrounded fp32).
#include <stdlib.h>

#include "common.cuh"

namespace pvg {

constexpr int kDThreads = 128;

template <int CIN, int KS>
__global__ void __launch_bounds__(kDThreads) conv_small_cin_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                   const float* __restrict__ bias, float* __restrict__ y,
                                                                   int N, int H, int W, int Cout, int act, float slope) {
  constexpr int K = KS * KS * CIN;
  constexpr int PAD = KS / 2;
  extern __shared__ float ws[];                    // [K][Cout]: for a fixed k the output channels are contiguous
  for (int i = threadIdx.x; i < K * Cout; i += kDThreads) {
    int co = i / K, k = i - co * K;
    ws[k * Cout + co] = __ldg(w + i);
  }
  __syncthreads();
  const int64_t M = (int64_t)N * H * W;
  const int64_t p = (int64_t)blockIdx.x * kDThreads + threadIdx.x;
  if (p >= M) return;
  const int pw = (int)(p % W);
  const int64_t t = p / W;
  const int ph = (int)(t % H);
  const int64_t pn = t / H;
  float in[K];
#pragma unroll
  for (int r = 0; r < KS; ++r) {
    const int ih = ph + r - PAD;
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      const int iw = pw + s - PAD;
      const bool ok = ih >= 0 && ih < H && iw >= 0 && iw < W;
      const float* px = x + ((pn * H + ih) * W + iw) * CIN;
#pragma unroll
      for (int c = 0; c < CIN; ++c) in[(r * KS + s) * CIN + c] = ok ? __ldg(px + c) : 0.f;
    }
  }
  float* yp = y + p * Cout;
  for (int co0 = 0; co0 < Cout; co0 += 16) {
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = (bias != nullptr && co0 + j < Cout) ? __ldg(bias + co0 + j) : 0.f;
    if (co0 + 16 <= Cout) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float v = in[k];
        const float4* wr = reinterpret_cast<const float4*>(ws + k * Cout + co0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 wv = wr[j];
          acc[4 * j + 0] = fmaf(v, wv.x, acc[4 * j + 0]);
          acc[4 * j + 1] = fmaf(v, wv.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(v, wv.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(v, wv.w, acc[4 * j + 3]);
        }
      }
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        stg4(yp + co0 + j, make_float4(act_fwd(acc[j], act, slope), act_fwd(acc[j + 1], act, slope),
                                       act_fwd(acc[j + 2], act, slope), act_fwd(acc[j + 3], act, slope)));
    } else {                                   // tail of a Cout that is a multiple of 4 but not of 16
      const int rem = Cout - co0;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float v = in[k];
        for (int j = 0; j < rem; ++j) acc[j] = fmaf(v, ws[k * Cout + co0 + j], acc[j]);
      }
      for (int j = 0; j < rem; ++j) yp[co0 + j] = act_fwd(acc[j], act, slope);
    }
  }
}

// Cout <= 4: ws = [K][4] (zero-padded), K = KS*KS*Cin with Cin % 4 == 0; float4 loads along the input channels
template <int KS>
__global__ void __launch_bounds__(kDThreads) conv_small_cout_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                    const float* __restrict__ bias, float* __restrict__ y,
                                                                    int N, int H, int W, int Cin, int Cout, int act,
                                                                    float slope) {
  constexpr int PAD = KS / 2;
  const int K = KS * KS * Cin;
  extern __shared__ float ws[];                    // [K][4]
  for (int i = threadIdx.x; i < K * 4; i += kDThreads) {
    int k = i >> 2, co = i & 3;
    ws[i] = co < Cout ? __ldg(w + (int64_t)co * K + k) : 0.f;
  }
  __syncthreads();
  const int64_t M = (int64_t)N * H * W;
  const int64_t p = (int64_t)blockIdx.x * kDThreads + threadIdx.x;
  if (p >= M) return;
  const int pw = (int)(p % W);
  const int64_t t = p / W;
  const int ph = (int)(t % H);
  const int64_t pn = t / H;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int r = 0; r < KS; ++r) {
    const int ih = ph + r - PAD;
    if (ih < 0 || ih >= H) continue;
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      const int iw = pw + s - PAD;
      if (iw < 0 || iw >= W) continue;
      const float* px = x + ((pn * H + ih) * W + iw) * Cin;
      const float4* wk = reinterpret_cast<const float4*>(ws) + (r * KS + s) * Cin;
      for (int c = 0; c < Cin; c += 4) {
        const float4 v = ldg4(px + c);
        const float4 w0 = wk[c], w1 = wk[c + 1], w2 = wk[c + 2], w3 = wk[c + 3];
        a0 = fmaf(v.x, w0.x, a0); a1 = fmaf(v.x, w0.y, a1); a2 = fmaf(v.x, w0.z, a2); a3 = fmaf(v.x, w0.w, a3);
        a0 = fmaf(v.y, w1.x, a0); a1 = fmaf(v.y, w1.y, a1); a2 = fmaf(v.y, w1.z, a2); a3 = fmaf(v.y, w1.w, a3);
        a0 = fmaf(v.z, w2.x, a0); a1 = fmaf(v.z, w2.y, a1); a2 = fmaf(v.z, w2.z, a2); a3 = fmaf(v.z, w2.w, a3);
        a0 = fmaf(v.w, w3.x, a0); a1 = fmaf(v.w, w3.y, a1); a2 = fmaf(v.w, w3.z, a2); a3 = fmaf(v.w, w3.w, a3);
      }
    }
  }
  const float acc[4] = {a0, a1, a2, a3};
  float* yp = y + p * Cout;
  for (int j = 0; j < Cout; ++j) yp[j] = act_fwd(acc[j] + (bias ? __ldg(bias + j) : 0.f), act, slope);
}

// returns 1 when a direct kernel took the problem, 0 when the caller should fall back, < 0 on error
int conv2d_fwd_direct(const pvg_conv_desc* d, const float* x, const float* w, const float* bias, float* y, cudaStream_t st) {
  const int64_t M = (int64_t)d->N * d->H * d->W;
  const unsigned grid = (unsigned)ceil_div64(M, kDThreads);
  const int K = d->R * d->S * d->Cin;
  if (d->R != d->S) return 0;
  static int disabled = -1;
  if (disabled < 0) { const char* e = getenv("PVG_NO_DIRECT"); disabled = (e && atoi(e) == 1) ? 1 : 0; }
  if (disabled) return 0;
  if ((d->Cin == 3 || d->Cin == 12) && d->Cout % 4 == 0 && (size_t)K * d->Cout * 4 <= 96 * 1024) {
    const size_t smem = (size_t)K * d->Cout * 4;
#define PVG_LAUNCH_CIN(CIN, KS)                                                                                          \
  do {                                                                                                                   \
    static bool set = false;                                                                                             \
    if (!set) {                                                                                                          \
      PVG_CUDA_OK(cudaFuncSetAttribute(conv_small_cin_kernel<CIN, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); \
      set = true;                                                                                                        \
    }                                                                                                                    \
    conv_small_cin_kernel<CIN, KS><<<grid, kDThreads, smem, st>>>(x, w, bias, y, d->N, d->H, d->W, d->Cout, d->act, d->slope); \
    PVG_LAUNCH_OK();                                                                                                     \
    return 1;                                                                                                            \
  } while (0)
    // measured on B200 (tools/direct_bench.py): 3x3 stems 1.2-3.2x faster than the implicit GEMM; the 7x7 variants
    // (147 register-resident taps) were slower than the generic kernels and are not used
    if (d->Cin == 3 && d->R == 3) PVG_LAUNCH_CIN(3, 3);
    if (d->Cin == 12 && d->R == 3) PVG_LAUNCH_CIN(12, 3);
#undef PVG_LAUNCH_CIN
  }
  if (d->Cout <= 4 && d->Cin % 4 == 0 && (size_t)K * 16 <= 96 * 1024 && (((uintptr_t)x) & 15) == 0) {
    const size_t smem = (size_t)K * 16;
#define PVG_LAUNCH_COUT(KS)                                                                                              \
  do {                                                                                                                   \
    static bool set = false;                                                                                             \
    if (!set) {                                                                                                          \
      PVG_CUDA_OK(cudaFuncSetAttribute(conv_small_cout_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); \
      set = true;                                                                                                        \
    }                                                                                                                    \
    conv_small_cout_kernel<KS><<<grid, kDThreads, smem, st>>>(x, w, bias, y, d->N, d->H, d->W, d->Cin, d->Cout, d->act, d->slope); \
    PVG_LAUNCH_OK();                                                                                                     \
    return 1;                                                                                                            \
  } while (0)
    if (d->R == 1) PVG_LAUNCH_COUT(1);
    if (d->R == 3) PVG_LAUNCH_COUT(3);
#undef PVG_LAUNCH_COUT
  }
  return 0;
}

}  // namespace pvg
