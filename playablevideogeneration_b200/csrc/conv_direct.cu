// Direct fp32 convolutions for the image-facing layers, where one side has 3 (or 12) channels and an implicit GEMM wastes
// its tile: these layers are HBM-bound (a 3->64 stem writes 256 B per pixel for 1.7 kFMA), so the kernels are organised
// around coalesced NHWC traffic - one thread per output pixel, weights broadcast from shared memory.
//   conv_small_cin  : Cin in {3, 12}, any Cout % 4 == 0  (E stem 3S->16, VGG conv1_1 3->64, data gradients of the 3-channel
//                     tanh heads: dY(3) -> 32/64/128 channels, 3x3 and 7x7)
//   conv_small_cout : Cout <= 4, Cin % 4 == 0             (tanh heads 128/64/32 -> 3 incl. the 7x7, data gradients of the
//                     stems: 16 -> 3S, 64 -> 3)
// Weights arrive in the common pack [Cout][R][S][Cin] (pvg_pack_conv_weight, unrounded fp32).
#include <cuda_fp16.h>
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


// ---------------------------------------------------------------------------------------------------------------
// Shared-memory tiled variants for the layers where the one-pixel-per-thread kernels above are latency/L1 bound:
// the tanh heads (Cin -> 3, 3x3 and the 7x7 of the 256^2 head), the data gradient of the 7x7 head (3 -> Cin) and the VGG
// conv1_1 data gradient (64 -> 3 over 120 frames).  The input patch (with its halo) is staged once per CTA in a
// channel-planar layout so that a warp's reads are contiguous along x; every thread keeps a strip of outputs in registers.
// ---------------------------------------------------------------------------------------------------------------
template <int KS>
struct Cout4Cfg {
  static constexpr int TH = 16, TW = 64, CC = 8;               // output tile, input channels per shared-memory chunk
  static constexpr int PH = TH + KS - 1, PW = TW + KS - 1;      // staged patch
  static constexpr int NV = (4 + KS - 1 + 3) / 4;               // float4 loads per 4-pixel strip (4 + KS - 1 values)
  static constexpr int PWS = 60 + 4 * NV;                       // row stride: last strip starts at 60 and reads 4*NV values
  static constexpr int kXFloats = CC * PH * PWS;
  static constexpr int kSmemBytes = kXFloats * 4 + CC * KS * KS * 16;
  static_assert(PWS >= PW && PWS % 4 == 0, "row stride");
};

// Cout <= 3, Cin % 8 == 0.  128 threads = 16 rows x 8 strips; a thread owns pixels [4*tx, 4*tx+4) and [32+4*tx, 32+4*tx+4)
// of its row (two strips 32 apart keep the float4 shared-memory reads of a quarter-warp contiguous: no bank conflicts).
template <int KS>
__global__ void __launch_bounds__(kDThreads) conv_cout4_tiled_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                     const float* __restrict__ bias, float* __restrict__ y,
                                                                     int N, int H, int W, int Cin, int Cout, int act,
                                                                     float slope) {
  using C = Cout4Cfg<KS>;
  constexpr int PAD = KS / 2;
  extern __shared__ float4 smem4[];
  float* xs = reinterpret_cast<float*>(smem4);                       // [CC][PH][PWS]
  float4* ws = smem4 + C::kXFloats / 4;                              // [CC][KS*KS] -> (co0, co1, co2, co3)
  const int tiles_w = ceil_div(W, C::TW), tiles_h = ceil_div(H, C::TH);
  int t = blockIdx.x;
  const int tile_w = t % tiles_w; t /= tiles_w;
  const int tile_h = t % tiles_h;
  const int n = t / tiles_h;
  const int ow0 = tile_w * C::TW, oh0 = tile_h * C::TH;
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  const int K = KS * KS * Cin;
  float acc[2][4][3];
#pragma unroll
  for (int g = 0; g < 2; ++g)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[g][j][0] = acc[g][j][1] = acc[g][j][2] = 0.f;

  for (int c0 = 0; c0 < Cin; c0 += C::CC) {
    __syncthreads();                                                 // the previous chunk has been consumed
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      for (int i = threadIdx.x; i < C::PH * C::PW; i += kDThreads) {
        const int py = i / C::PW, px = i - py * C::PW;
        const int ih = oh0 + py - PAD, iw = ow0 + px - PAD;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = ldg4(x + (((int64_t)n * H + ih) * W + iw) * Cin + c0 + half * 4);
        float* d = xs + ((half * 4) * C::PH + py) * C::PWS + px;
        d[0] = v.x; d[C::PH * C::PWS] = v.y; d[2 * C::PH * C::PWS] = v.z; d[3 * C::PH * C::PWS] = v.w;
      }
    }
    for (int i = threadIdx.x; i < C::CC * KS * KS; i += kDThreads) {
      const int c = i / (KS * KS), tap = i - c * (KS * KS);
      const float* wp = w + tap * Cin + c0 + c;
      ws[i] = make_float4(__ldg(wp), Cout > 1 ? __ldg(wp + K) : 0.f, Cout > 2 ? __ldg(wp + 2 * K) : 0.f, 0.f);
    }
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < C::CC; ++c) {
#pragma unroll
      for (int r = 0; r < KS; ++r) {
        const float* row = xs + (c * C::PH + ty + r) * C::PWS + tx * 4;
        float xv[2][4 * C::NV];
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
          for (int i = 0; i < C::NV; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(row + g * 32 + 4 * i);
            xv[g][4 * i] = v.x; xv[g][4 * i + 1] = v.y; xv[g][4 * i + 2] = v.z; xv[g][4 * i + 3] = v.w;
          }
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          const float4 wv = ws[c * (KS * KS) + r * KS + s];
#pragma unroll
          for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[g][j][0] = fmaf(xv[g][s + j], wv.x, acc[g][j][0]);
              acc[g][j][1] = fmaf(xv[g][s + j], wv.y, acc[g][j][1]);
              acc[g][j][2] = fmaf(xv[g][s + j], wv.z, acc[g][j][2]);
            }
        }
      }
    }
  }
  const int oy = oh0 + ty;
  if (oy >= H) return;
  float b[3];
#pragma unroll
  for (int co = 0; co < 3; ++co) b[co] = (bias != nullptr && co < Cout) ? __ldg(bias + co) : 0.f;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int ox = ow0 + g * 32 + tx * 4;
    if (ox >= W) continue;
    float* yp = y + (((int64_t)n * H + oy) * W + ox) * Cout;
    float v[12];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int co = 0; co < 3; ++co) v[j * 3 + co] = act_fwd(acc[g][j][co] + b[co], act, slope);
    if (Cout == 3 && ox + 4 <= W && (((uintptr_t)yp) & 15) == 0) {
      stg4(yp, make_float4(v[0], v[1], v[2], v[3]));
      stg4(yp + 4, make_float4(v[4], v[5], v[6], v[7]));
      stg4(yp + 8, make_float4(v[8], v[9], v[10], v[11]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (ox + j < W)
#pragma unroll
          for (int co = 0; co < 3; ++co)
            if (co < Cout) yp[j * Cout + co] = v[j * 3 + co];
    }
  }
}

// Cin == 3 (data gradient of a tanh head), KS x KS, CO output channels per CTA (blockIdx.y walks the chunks of Cout).
// 128 threads = 4 rows x 32 lanes; a thread owns pixels (row, lane) and (row, lane + 32) and all CO channels of both.
template <int KS, int CO>
__global__ void __launch_bounds__(kDThreads) conv_cin3_tiled_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                    const float* __restrict__ bias, float* __restrict__ y,
                                                                    int N, int H, int W, int Cout, int act, float slope) {
  constexpr int PAD = KS / 2, TH = 4, TW = 64, PH = TH + KS - 1, PW = TW + KS - 1, PWS = PW + 2, K = KS * KS * 3;
  extern __shared__ float4 smem4[];
  float4* ws4 = smem4;                                               // [K][CO]
  float* ws = reinterpret_cast<float*>(smem4);
  float* xs = ws + K * CO;                                           // [3][PH][PWS]
  const int tiles_w = ceil_div(W, TW), tiles_h = ceil_div(H, TH);
  int t = blockIdx.x;
  const int tile_w = t % tiles_w; t /= tiles_w;
  const int tile_h = t % tiles_h;
  const int n = t / tiles_h;
  const int ow0 = tile_w * TW, oh0 = tile_h * TH;
  const int co0 = blockIdx.y * CO;
  const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < PH * PW; i += kDThreads) {
    const int py = i / PW, px = i - py * PW;
    const int ih = oh0 + py - PAD, iw = ow0 + px - PAD;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
      const float* px_ = x + (((int64_t)n * H + ih) * W + iw) * 3;
      v0 = __ldg(px_); v1 = __ldg(px_ + 1); v2 = __ldg(px_ + 2);
    }
    float* d = xs + py * PWS + px;
    d[0] = v0; d[PH * PWS] = v1; d[2 * PH * PWS] = v2;
  }
  for (int i = threadIdx.x; i < K * CO; i += kDThreads) {            // lanes walk co: conflict-free stores, L1-served loads
    const int k = i / CO, co = i - k * CO;
    ws[i] = (co0 + co < Cout) ? __ldg(w + (int64_t)(co0 + co) * K + k) : 0.f;
  }
  __syncthreads();
  float acc[2][CO];
#pragma unroll
  for (int j = 0; j < CO; ++j) acc[0][j] = acc[1][j] = 0.f;
#pragma unroll 1
  for (int r = 0; r < KS; ++r) {
#pragma unroll
    for (int s = 0; s < KS; ++s) {
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float* xp = xs + (ci * PH + row + r) * PWS + lane + s;
        const float x0 = xp[0], x1 = xp[32];
        const float4* wk = ws4 + ((r * KS + s) * 3 + ci) * (CO / 4);
#pragma unroll
        for (int q = 0; q < CO / 4; ++q) {
          const float4 wv = wk[q];
          acc[0][4 * q] = fmaf(x0, wv.x, acc[0][4 * q]); acc[0][4 * q + 1] = fmaf(x0, wv.y, acc[0][4 * q + 1]);
          acc[0][4 * q + 2] = fmaf(x0, wv.z, acc[0][4 * q + 2]); acc[0][4 * q + 3] = fmaf(x0, wv.w, acc[0][4 * q + 3]);
          acc[1][4 * q] = fmaf(x1, wv.x, acc[1][4 * q]); acc[1][4 * q + 1] = fmaf(x1, wv.y, acc[1][4 * q + 1]);
          acc[1][4 * q + 2] = fmaf(x1, wv.z, acc[1][4 * q + 2]); acc[1][4 * q + 3] = fmaf(x1, wv.w, acc[1][4 * q + 3]);
        }
      }
    }
  }
  const int oy = oh0 + row;
  if (oy >= H) return;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int ox = ow0 + lane + 32 * g;
    if (ox >= W) continue;
    float* yp = y + (((int64_t)n * H + oy) * W + ox) * Cout + co0;
#pragma unroll
    for (int q = 0; q < CO / 4; ++q) {
      if (co0 + 4 * q >= Cout) break;                               // Cout % 4 == 0
      float4 b = bias != nullptr ? ldg4(bias + co0 + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
      stg4(yp + 4 * q, make_float4(act_fwd(acc[g][4 * q] + b.x, act, slope), act_fwd(acc[g][4 * q + 1] + b.y, act, slope),
                                   act_fwd(acc[g][4 * q + 2] + b.z, act, slope), act_fwd(acc[g][4 * q + 3] + b.w, act, slope)));
    }
  }
}

// Weight gradient of a KS x KS head with Cout <= 3 and <= 32 input channels:
//   dW[co][ci][r][s] += sum_p dY[p][co] * X[p + (r,s) - pad][ci].
// Persistent CTAs of KS warps: warp = tap row r, lane = ci, 3*KS accumulators per thread; per 4 x 32-pixel tile the X patch
// ([y][x][ci]: a warp reads 32 consecutive ci) and the dY tile are staged in shared memory, a row of X is held in
// registers while the 32 pixels of the output row slide over it.  One atomic flush per CTA.
template <int KS>
__global__ void __launch_bounds__(KS * 32) wgrad_cout4_tiled_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                    float* __restrict__ dw, int N, int H, int W, int CinP,
                                                                    int Cin, int Cout) {
  constexpr int PAD = KS / 2, TH = 4, TW = 32, PH = TH + KS - 1, PW = TW + KS - 1, THREADS = KS * 32;
  extern __shared__ float4 smem4[];
  float4* xs4 = smem4;                                               // [PH][PW][32 ci]
  const float* xs = reinterpret_cast<const float*>(smem4);
  float4* gs = smem4 + PH * PW * 8;                                  // [TH][TW] -> (dy0, dy1, dy2, dy3)
  const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_w = ceil_div(W, TW), tiles_h = ceil_div(H, TH);
  const int tiles_total = tiles_w * tiles_h * N;
  float acc[KS][3];
#pragma unroll
  for (int s = 0; s < KS; ++s) acc[s][0] = acc[s][1] = acc[s][2] = 0.f;
  for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
    int t = tile;
    const int tile_w = t % tiles_w; t /= tiles_w;
    const int tile_h = t % tiles_h;
    const int n = t / tiles_h;
    const int ow0 = tile_w * TW, oh0 = tile_h * TH;
    __syncthreads();
    for (int i = threadIdx.x; i < PH * PW * 8; i += THREADS) {
      const int pix = i >> 3, q = i & 7;
      const int py = pix / PW, px = pix - py * PW;
      const int ih = oh0 + py - PAD, iw = ow0 + px - PAD;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q * 4 < CinP && ih >= 0 && ih < H && iw >= 0 && iw < W) v = ldg4(x + (((int64_t)n * H + ih) * W + iw) * CinP + q * 4);
      xs4[i] = v;
    }
    for (int i = threadIdx.x; i < TH * TW; i += THREADS) {
      const int py = i / TW, px = i - py * TW;
      const int oh = oh0 + py, ow = ow0 + px;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (oh < H && ow < W) {
        const float* g = dy + (((int64_t)n * H + oh) * W + ow) * Cout;
        v.x = __ldg(g);
        if (Cout > 1) v.y = __ldg(g + 1);
        if (Cout > 2) v.z = __ldg(g + 2);
      }
      gs[i] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int oy = 0; oy < TH; ++oy) {
      const float* xr = xs + (oy + r) * PW * 32 + lane;
      float xrow[PW];
#pragma unroll
      for (int k = 0; k < PW; ++k) xrow[k] = xr[k * 32];
#pragma unroll
      for (int ox = 0; ox < TW; ++ox) {
        const float4 g = gs[oy * TW + ox];
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          acc[s][0] = fmaf(xrow[ox + s], g.x, acc[s][0]);
          acc[s][1] = fmaf(xrow[ox + s], g.y, acc[s][1]);
          acc[s][2] = fmaf(xrow[ox + s], g.z, acc[s][2]);
        }
      }
    }
  }
  if (lane < Cin) {
#pragma unroll
    for (int s = 0; s < KS; ++s) {
#pragma unroll
      for (int co = 0; co < 3; ++co)
        if (co < Cout) atomicAdd(dw + (((int64_t)co * Cin + lane) * KS + r) * KS + s, acc[s][co]);
    }
  }
}

// Weight gradient of a 3x3 image stem (Cin = 3 or 12 input channels, e.g. the encoder's conv1 3S -> 16): the layer is
// HBM-bound (x and dY are each read once: 76 B per pixel for 3 -> 16), so persistent CTAs stream 8 x 32-pixel tiles: the x
// patch (with halo) goes through shared memory, thread = (output channel, 1 of 16 pixel lanes) keeps all 9*CIN taps of its
// channel in registers, dY is read straight from global memory (16 consecutive channels x 2 pixels per warp).  One
// shuffle + shared-memory reduction and one global atomic per output and CTA at the end.
template <int CIN>
__global__ void __launch_bounds__(256) wgrad_small_cin_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                              float* __restrict__ dw, int N, int H, int W, int Cout) {
  constexpr int TH = 8, TW = 32, PH = TH + 2, PW = TW + 2, KT = 9 * CIN;
  __shared__ float xs[PH * PW * CIN];
  __shared__ float red[16 * KT];
  const int co = threadIdx.x & 15, pl = threadIdx.x >> 4;
  const int co_g = blockIdx.y * 16 + co;
  const bool co_ok = co_g < Cout;
  const int tiles_w = ceil_div(W, TW), tiles_h = ceil_div(H, TH);
  const int tiles_total = tiles_w * tiles_h * N;
  float acc[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) acc[k] = 0.f;
  for (int i = threadIdx.x; i < 16 * KT; i += 256) red[i] = 0.f;
  for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
    int t = tile;
    const int tile_w = t % tiles_w; t /= tiles_w;
    const int tile_h = t % tiles_h;
    const int n = t / tiles_h;
    const int ow0 = tile_w * TW, oh0 = tile_h * TH;
    __syncthreads();
    for (int i = threadIdx.x; i < PH * PW * CIN; i += 256) {
      const int c = i % CIN, pix = i / CIN;
      const int py = pix / PW, px = pix - py * PW;
      const int ih = oh0 + py - 1, iw = ow0 + px - 1;
      xs[i] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? __ldg(x + (((int64_t)n * H + ih) * W + iw) * CIN + c) : 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int i = 0; i < (TH * TW) / 16; ++i) {
      const int pix = pl + 16 * i;
      const int py = pix / TW, px = pix - py * TW;
      const int oh = oh0 + py, ow = ow0 + px;
      const float g = (co_ok && oh < H && ow < W) ? __ldg(dy + (((int64_t)n * H + oh) * W + ow) * Cout + co_g) : 0.f;
      const float* xp = xs + (py * PW + px) * CIN;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < 3 * CIN; ++k) acc[r * 3 * CIN + k] = fmaf(g, xp[r * PW * CIN + k], acc[r * 3 * CIN + k]);
    }
  }
  // lanes l and l + 16 of a warp hold the same output channel
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    float v = acc[k] + __shfl_xor_sync(0xffffffffu, acc[k], 16);
    if ((threadIdx.x & 16) == 0) atomicAdd(&red[co * KT + k], v);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 16 * KT; i += 256) {
    const int c = i / KT, k = i - c * KT;          // k = (r*3 + s) * CIN + ci
    const int cg = blockIdx.y * 16 + c;
    if (cg < Cout) {
      const int tap = k / CIN, ci = k - tap * CIN;
      atomicAdd(dw + ((int64_t)cg * CIN + ci) * 9 + tap, red[i]);
    }
  }
}

// returns 1 when the tiled weight-gradient kernel took the problem, 0 otherwise, < 0 on error
int conv2d_wgrad_direct(const pvg_conv_desc* d, int Cin_logical, const float* x, const float* dy, float* dw, cudaStream_t st) {
  if (d->R == 3 && d->S == 3 && (d->Cin == 3 || d->Cin == 12) && Cin_logical == d->Cin) {      // image stems
    const int64_t tiles = (int64_t)ceil_div(d->W, 32) * ceil_div(d->H, 8) * d->N;
    const int64_t cap = (int64_t)kSMs * 4;
    dim3 grid((unsigned)(tiles < cap ? tiles : cap), ceil_div(d->Cout, 16));
    if (d->Cin == 3) wgrad_small_cin_kernel<3><<<grid, 256, 0, st>>>(x, dy, dw, d->N, d->H, d->W, d->Cout);
    else wgrad_small_cin_kernel<12><<<grid, 256, 0, st>>>(x, dy, dw, d->N, d->H, d->W, d->Cout);
    PVG_LAUNCH_OK();
    return 1;
  }
  if (!(d->R == 7 && d->S == 7 && d->Cout <= 3 && d->Cin <= 32 && d->Cin % 4 == 0 && (((uintptr_t)x) & 15) == 0)) return 0;
  constexpr int KS = 7, PH = 4 + KS - 1, PW = 32 + KS - 1;
  const int smem = PH * PW * 32 * 4 + 4 * 32 * 16;
  static bool set = false;
  if (!set) {
    PVG_CUDA_OK(cudaFuncSetAttribute(wgrad_cout4_tiled_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    set = true;
  }
  const int64_t tiles = (int64_t)ceil_div(d->W, 32) * ceil_div(d->H, 4) * d->N;
  const int64_t cap = (int64_t)kSMs * 4;
  const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
  wgrad_cout4_tiled_kernel<KS><<<grid, KS * 32, smem, st>>>(x, dy, dw, d->N, d->H, d->W, d->Cin, Cin_logical, d->Cout);
  PVG_LAUNCH_OK();
  return 1;
}

// returns 1 when a direct kernel took the problem, 0 when the caller should fall back, < 0 on error
// ---------------------------------------------------------------------------------------------------------------
// 3 -> 64 channels, 3x3 (VGG19 conv1_1, vgg.py:48-52 on the 256^2 frames: 240 frames per step per resolution).  The layer is
// bound by its OUTPUT: 64 fp32 channels per pixel plus the fp16 plane pair conv1_2 consumes = 512 B per pixel against 1 728 FMAs.
// Persistent blocks of 256 threads = 16 pixel groups x 16 channel quads.  A thread keeps the 27 x 4 weights of its channel quad
// in registers for the whole launch and computes 4 horizontally adjacent pixels x 4 channels at a time from five aligned
// float4 shared-memory reads per input row (r01/r02 captures: the per-pixel variant spent its issue slots on scalar LDS).
// A warp's store instruction covers two whole pixels (2 x 256 B of y, 2 x 128 B of each plane).  The haloed input tile
// (10 x 34 pixels x 3 channels) is double-buffered: the next tile is fetched into registers while the current one is computed.
// ---------------------------------------------------------------------------------------------------------------
struct Stem3Cfg {
  static constexpr int TH = 8, TW = 32, COUT = 64, THREADS = 256;
  static constexpr int PH = TH + 2, PWF = (TW + 2) * 3, PWS = 104;     // staged rows, floats per staged row, row stride
  static constexpr int STAGE = (PH * PWF + THREADS - 1) / THREADS;     // staged floats per thread
  static_assert(PWS % 4 == 0 && PWS >= (TW - 4) * 3 + 20, "aligned float4 reads stay inside the row");
};

__global__ void __launch_bounds__(Stem3Cfg::THREADS, 1) conv_stem3_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                          const float* __restrict__ bias, float* __restrict__ y,
                                                                          uint16_t* __restrict__ planes, int64_t y_numel, int N,
                                                                          int H, int W, int act, float slope) {
  using C = Stem3Cfg;
  __shared__ __align__(16) float xs[2][C::PH][C::PWS];
  const int tiles_w = ceil_div(W, C::TW), tiles_h = ceil_div(H, C::TH);
  const int tiles = tiles_w * tiles_h * N;
  const int cq = threadIdx.x & 15, pg = threadIdx.x >> 4;
  float wr[27][4];
#pragma unroll
  for (int k = 0; k < 27; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) wr[k][j] = __ldg(w + (cq * 4 + j) * 27 + k);
  float b4[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) b4[j] = bias != nullptr ? __ldg(bias + cq * 4 + j) : 0.f;

  float stage[C::STAGE];
  auto fetch = [&](int tile) {
    int t = tile;
    const int tile_w = t % tiles_w; t /= tiles_w;
    const int tile_h = t % tiles_h;
    const int n = t / tiles_h;
#pragma unroll
    for (int q = 0; q < C::STAGE; ++q) {
      const int i = threadIdx.x + q * C::THREADS;
      const int py = i / C::PWF, pf = i - py * C::PWF;
      const int ih = tile_h * C::TH + py - 1, f = (tile_w * C::TW - 1) * 3 + pf;          // f: float index inside the image row
      stage[q] = (i < C::PH * C::PWF && ih >= 0 && ih < H && f >= 0 && f < W * 3) ? __ldg(x + ((int64_t)n * H + ih) * W * 3 + f) : 0.f;
    }
  };
  auto commit = [&](int buf) {
#pragma unroll
    for (int q = 0; q < C::STAGE; ++q) {
      const int i = threadIdx.x + q * C::THREADS;
      const int py = i / C::PWF, pf = i - py * C::PWF;
      if (i < C::PH * C::PWF) xs[buf][py][pf] = stage[q];
    }
  };
  int tile = blockIdx.x;
  if (tile < tiles) { fetch(tile); commit(0); }
  __syncthreads();
  for (int buf = 0; tile < tiles; tile += gridDim.x, buf ^= 1) {
    const int next = tile + gridDim.x;
    if (next < tiles) fetch(next);
    int t = tile;
    const int tile_w = t % tiles_w; t /= tiles_w;
    const int tile_h = t % tiles_h;
    const int n = t / tiles_h;
    const int ow0 = tile_w * C::TW, oh0 = tile_h * C::TH;
#pragma unroll 1
    for (int it = 0; it < C::TH * C::TW / 64; ++it) {
      const int quad = it * 16 + pg;                       // 4 adjacent pixels of one tile row
      const int ph = quad / (C::TW / 4), pw = (quad % (C::TW / 4)) * 4;
      float acc[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[u][j] = b4[j];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        float in[20];
        const float4* row = reinterpret_cast<const float4*>(&xs[buf][ph + r][pw * 3]);
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const float4 v = row[q];
          in[4 * q] = v.x; in[4 * q + 1] = v.y; in[4 * q + 2] = v.z; in[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int q = 0; q < 9; ++q)
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[u][j] = fmaf(in[u * 3 + q], wr[r * 9 + q][j], acc[u][j]);
      }
      const int oh = oh0 + ph;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int ow = ow0 + pw + u;
        if (oh >= H || ow >= W) continue;
        const float4 v = make_float4(act_fwd(acc[u][0], act, slope), act_fwd(acc[u][1], act, slope),
                                     act_fwd(acc[u][2], act, slope), act_fwd(acc[u][3], act, slope));
        const int64_t e = (((int64_t)n * H + oh) * W + ow) * C::COUT + cq * 4;
        stg4(y + e, v);
        if (planes != nullptr) {              // PVG_CORR_FP16_ALL pair: { f16((v - f16(v)) * 2^12), f16(v) }, see pointwise.cu
          const uint32_t h01 = pack_f16x2_sat(v.x, v.y), h23 = pack_f16x2_sat(v.z, v.w);
          const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&h01)), f23 = __half22float2(*reinterpret_cast<const __half2*>(&h23));
          const uint32_t g01 = pack_f16x2_sat((v.x - f01.x) * 4096.f, (v.y - f01.y) * 4096.f);
          const uint32_t g23 = pack_f16x2_sat((v.z - f23.x) * 4096.f, (v.w - f23.y) * 4096.f);
          *reinterpret_cast<uint2*>(planes + e) = make_uint2(g01, g23);
          *reinterpret_cast<uint2*>(planes + y_numel + e) = make_uint2(h01, h23);
        }
      }
    }
    if (next < tiles) commit(buf ^ 1);
    __syncthreads();
  }
}

int conv2d_stem3_planes(const pvg_conv_desc* d, const float* x, const float* w, const float* bias, float* y, void* y_planes,
                        cudaStream_t st) {
  using C = Stem3Cfg;
  const int tiles = ceil_div(d->W, C::TW) * ceil_div(d->H, C::TH) * d->N;
  static int sms = 0;
  if (sms == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  const int grid = tiles < sms ? tiles : sms;        // one persistent block per SM (160 registers x 256 threads)
  conv_stem3_kernel<<<grid, C::THREADS, 0, st>>>(x, w, bias, y, (uint16_t*)y_planes, (int64_t)d->N * d->H * d->W * C::COUT, d->N,
                                                 d->H, d->W, d->act, d->slope);
  PVG_LAUNCH_OK();
  return 0;
}

int conv2d_fwd_direct(const pvg_conv_desc* d, const float* x, const float* w, const float* bias, float* y, cudaStream_t st) {
  const int64_t M = (int64_t)d->N * d->H * d->W;
  const unsigned grid = (unsigned)ceil_div64(M, kDThreads);
  const int K = d->R * d->S * d->Cin;
  if (d->R != d->S) return 0;
  static int disabled = -1;
  if (disabled < 0) { const char* e = getenv("PVG_NO_DIRECT"); disabled = (e && atoi(e) == 1) ? 1 : 0; }
  if (disabled) return 0;
  if (d->Cin == 3 && d->R == 7 && d->Cout % 16 == 0 && (((uintptr_t)y) & 15) == 0 &&
      (bias == nullptr || (((uintptr_t)bias) & 15) == 0)) {
    // data gradient of the 7x7 tanh head
    constexpr int TH = 4, TW = 64, PH = TH + 6, PWS = TW + 6 + 2, K7 = 49 * 3;
    const unsigned tiles = (unsigned)(ceil_div(d->W, TW) * ceil_div(d->H, TH) * d->N);
#define PVG_LAUNCH_CIN3(CO)                                                                                              \
  do {                                                                                                                   \
    const int smem = (K7 * CO + 3 * PH * PWS) * 4;                                                                       \
    static bool set = false;                                                                                             \
    if (!set) {                                                                                                          \
      PVG_CUDA_OK(cudaFuncSetAttribute(conv_cin3_tiled_kernel<7, CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
      set = true;                                                                                                        \
    }                                                                                                                    \
    conv_cin3_tiled_kernel<7, CO><<<dim3(tiles, ceil_div(d->Cout, CO)), kDThreads, smem, st>>>(                          \
        x, w, bias, y, d->N, d->H, d->W, d->Cout, d->act, d->slope);                                                     \
    PVG_LAUNCH_OK();                                                                                                     \
    return 1;                                                                                                            \
  } while (0)
    if (d->Cout % 32 == 0) PVG_LAUNCH_CIN3(32);
    PVG_LAUNCH_CIN3(16);
#undef PVG_LAUNCH_CIN3
  }
  if (d->Cout <= 3 && d->Cin % 8 == 0 && (d->R == 3 || d->R == 7) && (((uintptr_t)x) & 15) == 0) {
    const unsigned tiles = (unsigned)(ceil_div(d->W, 64) * ceil_div(d->H, 16) * d->N);
#define PVG_LAUNCH_COUT4(KS)                                                                                             \
  do {                                                                                                                   \
    static bool set = false;                                                                                             \
    if (!set) {                                                                                                          \
      PVG_CUDA_OK(cudaFuncSetAttribute(conv_cout4_tiled_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                       Cout4Cfg<KS>::kSmemBytes));                                                       \
      set = true;                                                                                                        \
    }                                                                                                                    \
    conv_cout4_tiled_kernel<KS><<<tiles, kDThreads, Cout4Cfg<KS>::kSmemBytes, st>>>(x, w, bias, y, d->N, d->H, d->W, d->Cin, \
                                                                                    d->Cout, d->act, d->slope);          \
    PVG_LAUNCH_OK();                                                                                                     \
    return 1;                                                                                                            \
  } while (0)
    if (d->R == 3) PVG_LAUNCH_COUT4(3);
    PVG_LAUNCH_COUT4(7);
#undef PVG_LAUNCH_COUT4
  }
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

namespace pvg {

// ---------------------------------------------------------------------------------------------------------------
// Weight gradient of 3x3 convolutions with 16 / 32 channels on both sides (the first encoder stage: 16 channels at 128 x 128
// over all B*T frames, residual_block.py:52-58).  The GEMM is 144..288 x 16..32 with K = every pixel of the batch: the
// tensor-core kernel spent 0.93 ms on 268 MB (r02 layer table: 10 TFLOP/s, 4 % of its floor).  Here a thread owns one
// (input channel, output channel) pair - or 2 / 4 of them - and keeps its 9 taps in registers while persistent blocks walk
// the image tiles; x slides through a 3 x 3 register window, so a pixel costs 3 + P shared-memory reads for 9 P FMAs.
// fp32 products and sums; one atomicAdd per accumulator and block into the packed scratch of the tensor-core path.
// ---------------------------------------------------------------------------------------------------------------
template <int CIN, int COUT>
struct WgradSmallCfg {
  static constexpr int TH = 4, TW = 32, THREADS = 256;
  static constexpr int COG = THREADS / CIN;               // output channels covered by one pass over the threads
  static constexpr int P = COUT / COG;                    // (ci, co) pairs per thread
  static constexpr int XROW = (TW + 2) * CIN;
  static_assert(COUT % COG == 0 && P >= 1, "thread mapping");
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(256) wgrad_small_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                          float* __restrict__ scratch, int N, int H, int W) {
  using C = WgradSmallCfg<CIN, COUT>;
  __shared__ __align__(16) float xs[C::TH + 2][C::XROW];
  __shared__ __align__(16) float gs[C::TH][C::TW * COUT];
  const int ci = threadIdx.x % CIN, cb = threadIdx.x / CIN;
  const int tiles_w = ceil_div(W, C::TW), tiles_h = ceil_div(H, C::TH);
  const int tiles = tiles_w * tiles_h * N;
  float acc[C::P][9];
#pragma unroll
  for (int k = 0; k < C::P; ++k)
#pragma unroll
    for (int q = 0; q < 9; ++q) acc[k][q] = 0.f;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    int t = tile;
    const int tile_w = t % tiles_w; t /= tiles_w;
    const int tile_h = t % tiles_h;
    const int n = t / tiles_h;
    const int ow0 = tile_w * C::TW, oh0 = tile_h * C::TH;
    __syncthreads();                                   // the previous tile has been consumed
    for (int i = threadIdx.x; i < (C::TH + 2) * (C::XROW / 4); i += C::THREADS) {
      const int py = i / (C::XROW / 4), pf = (i - py * (C::XROW / 4)) * 4;
      const int ih = oh0 + py - 1, iw = ow0 - 1 + pf / CIN;          // CIN % 4 == 0: a float4 stays inside one pixel
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = ldg4(x + (((int64_t)n * H + ih) * W + iw) * CIN + pf % CIN);
      *reinterpret_cast<float4*>(&xs[py][pf]) = v;
    }
    for (int i = threadIdx.x; i < C::TH * (C::TW * COUT / 4); i += C::THREADS) {
      const int py = i / (C::TW * COUT / 4), pf = (i - py * (C::TW * COUT / 4)) * 4;
      const int oh = oh0 + py, ow = ow0 + pf / COUT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (oh < H && ow < W) v = ldg4(g + (((int64_t)n * H + oh) * W + ow) * COUT + pf % COUT);
      *reinterpret_cast<float4*>(&gs[py][pf]) = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int ph = 0; ph < C::TH; ++ph) {
      float win[3][3];                                 // x[ph + r][pw + s][ci]
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        win[r][1] = xs[ph + r][ci];
        win[r][2] = xs[ph + r][CIN + ci];
      }
#pragma unroll 4
      for (int pw = 0; pw < C::TW; ++pw) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          win[r][0] = win[r][1]; win[r][1] = win[r][2];
          win[r][2] = xs[ph + r][(pw + 2) * CIN + ci];
        }
#pragma unroll
        for (int k = 0; k < C::P; ++k) {
          const float gv = gs[ph][pw * COUT + cb + k * C::COG];
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int q = 0; q < 3; ++q) acc[k][r * 3 + q] = fmaf(gv, win[r][q], acc[k][r * 3 + q]);
        }
      }
    }
  }
  constexpr int CinP = (CIN + 31) & ~31;
#pragma unroll
  for (int k = 0; k < C::P; ++k)
#pragma unroll
    for (int q = 0; q < 9; ++q) atomicAdd(scratch + ((size_t)(cb + k * C::COG) * 9 + q) * CinP + ci, acc[k][q]);
}

template <int CIN, int COUT>
static int launch_wgrad_small(const pvg_conv_desc* d, const float* x, const float* g, float* scratch, cudaStream_t st) {
  using C = WgradSmallCfg<CIN, COUT>;
  const int tiles = ceil_div(d->W, C::TW) * ceil_div(d->H, C::TH) * d->N;
  const int grid = tiles < 148 * 2 ? tiles : 148 * 2;      // persistent: two blocks per SM, one flush of the accumulators each
  wgrad_small_kernel<CIN, COUT><<<grid, C::THREADS, 0, st>>>(x, g, scratch, d->N, d->H, d->W);
  PVG_LAUNCH_OK();
  return 0;
}

}  // namespace pvg

using namespace pvg;

extern "C" int pvg_conv2d_wgrad_small(const pvg_conv_desc* d, int Cin_logical, const float* x, const float* g, float* scratch,
                                      float* dw_oihw, int accumulate, void* stream) {
  PVG_CHECK_ARG(d && x && g && scratch, "null argument");
  PVG_CHECK_ARG((d->Cin == 16 || d->Cin == 32) && (d->Cout == 16 || d->Cout == 32) && d->R == 3 && d->S == 3 && d->pad == 1,
                "16 / 32 channels on both sides, 3x3, 'same' padding only");
  PVG_CHECK_ARG((((uintptr_t)x | (uintptr_t)g) & 15) == 0, "x / g must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (d->Cin == 16) rc = d->Cout == 16 ? launch_wgrad_small<16, 16>(d, x, g, scratch, st) : launch_wgrad_small<16, 32>(d, x, g, scratch, st);
  else rc = d->Cout == 16 ? launch_wgrad_small<32, 16>(d, x, g, scratch, st) : launch_wgrad_small<32, 32>(d, x, g, scratch, st);
  if (rc || dw_oihw == nullptr) return rc;
  return pvg_unpack_dw(scratch, d->Cout, Cin_logical, 3, 3, d->Cin, dw_oihw, accumulate, stream);
}

extern "C" int pvg_conv2d_stem_planes(const pvg_conv_desc* d, const float* x, const float* w, const float* bias, float* y,
                                      void* y_planes, void* stream) {
  PVG_CHECK_ARG(d && x && w && y, "null argument");
  PVG_CHECK_ARG(d->Cin == 3 && d->Cout == 64 && d->R == 3 && d->S == 3 && d->pad == 1, "3 -> 64 channels, 3x3, 'same' padding only");
  PVG_CHECK_ARG((((uintptr_t)y | (uintptr_t)y_planes) & 15) == 0, "y / y_planes must be 16-byte aligned");
  return conv2d_stem3_planes(d, x, w, bias, y, y_planes, (cudaStream_t)stream);
}
