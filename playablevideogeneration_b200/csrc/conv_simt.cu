// fp32 CUDA-core implicit-GEMM convolution (forward / data-gradient via packed weights) and weight gradient.
//
// Role: (1) the kernel for shapes the tensor-core path does not take (Cin not a multiple of 32: the 3/12-channel image
// stems, the 16-channel encoder head - all HBM-bound layers); (2) the on-device cross-check of conv_umma.cu.
// GEMM view: M = N*H*W output pixels, N = Cout, K = R*S*Cin (ci fastest), A gathered on the fly from NHWC x with zero
// padding, B = packed weights [Cout][R][S][Cin].
#include "common.cuh"

namespace pvg {

constexpr int BM = 64, BK = 16, THREADS = 256;

template <int BN>
__global__ void __launch_bounds__(THREADS) conv_fwd_simt_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                const float* __restrict__ bias, float* __restrict__ y,
                                                                int N, int H, int W, int Cin, int Cout, int R, int S, int pad,
                                                                int act, float slope) {
  constexpr int TN = BN / 16;      // outputs per thread along Cout
  constexpr int TM = BM / 16;      // outputs per thread along pixels (4)
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int64_t M = (int64_t)N * H * W;
  const int K = R * S * Cin;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // A loader: thread -> (row = tid / 4, 4 consecutive k) ; B loader: (col = tid / 4 [+64...], 4 consecutive k)
  const int a_row = threadIdx.x / 4, a_k = (threadIdx.x % 4) * 4;
  const int64_t am = m0 + a_row;
  int an = 0, ah = 0, aw = 0;
  const bool a_valid = am < M;
  if (a_valid) { aw = (int)(am % W); int64_t t = am / W; ah = (int)(t % H); an = (int)(t / H); }

  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = k0 + a_k + j;
      float v = 0.f;
      if (a_valid && k < K) {
        int ci = k % Cin, tap = k / Cin;
        int s = tap % S, r = tap / S;
        int ih = ah + r - pad, iw = aw + s - pad;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = __ldg(x + (((int64_t)an * H + ih) * W + iw) * Cin + ci);
      }
      As[a_k + j][a_row] = v;
    }
    for (int col = threadIdx.x / 4; col < BN; col += THREADS / 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int k = k0 + a_k + j;
        int co = n0 + col;
        Bs[a_k + j][col] = (co < Cout && k < K) ? __ldg(w + (int64_t)co * K + k) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int64_t m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int co = n0 + tx * TN + j;
      if (co >= Cout) continue;
      float v = acc[i][j] + (bias ? __ldg(bias + co) : 0.f);
      y[m * Cout + co] = act_fwd(v, act, slope);
    }
  }
}

// Weight gradient: dW[co][ci][r][s] += sum_m dy[m][co] * x[m @ (r,s)][ci].
// GEMM view: rows = Cout (64 per CTA), cols = (tap, ci) (64 per CTA), reduction over pixels split across blockIdx.z.
__global__ void __launch_bounds__(THREADS) conv_wgrad_simt_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                  float* __restrict__ dw, int N, int H, int W, int CinP,
                                                                  int Cin, int Cout, int R, int S, int pad,
                                                                  int64_t pixels_per_split) {
  __shared__ float As[BK][64 + 4];     // dy[m][co]
  __shared__ float Bs[BK][64 + 4];     // x[m@tap][ci]
  const int64_t M = (int64_t)N * H * W;
  const int KC = R * S * Cin;          // number of (tap, ci) columns
  const int co0 = blockIdx.x * 64, col0 = blockIdx.y * 64;
  const int64_t p_begin = (int64_t)blockIdx.z * pixels_per_split;
  int64_t p_end = p_begin + pixels_per_split;
  if (p_end > M) p_end = M;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // loaders: 16 pixels x 64 columns per slab; thread -> (pixel = tid / 16, 4 consecutive columns = (tid % 16) * 4)
  const int lp = threadIdx.x / 16, lc = (threadIdx.x % 16) * 4;
  // decode this thread's 4 B-columns once
  int b_ci[4], b_r[4], b_s[4];
  bool b_ok[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int col = col0 + lc + j;
    b_ok[j] = col < KC;
    int tap = b_ok[j] ? col / Cin : 0;
    b_ci[j] = b_ok[j] ? col % Cin : 0;
    b_r[j] = tap / S;
    b_s[j] = tap % S;
  }
  for (int64_t p0 = p_begin; p0 < p_end; p0 += BK) {
    int64_t m = p0 + lp;
    bool mv = m < p_end;
    int pw = 0, ph = 0, pn = 0;
    if (mv) { pw = (int)(m % W); int64_t t = m / W; ph = (int)(t % H); pn = (int)(t / H); }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int co = co0 + lc + j;
      As[lp][lc + j] = (mv && co < Cout) ? __ldg(dy + m * Cout + co) : 0.f;
      float v = 0.f;
      if (mv && b_ok[j]) {
        int ih = ph + b_r[j] - pad, iw = pw + b_s[j] - pad;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = __ldg(x + (((int64_t)pn * H + ih) * W + iw) * CinP + b_ci[j]);
      }
      Bs[lp][lc + j] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int co = co0 + ty * 4 + i;
    if (co >= Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int col = col0 + tx * 4 + j;
      if (col >= KC) continue;
      int tap = col / Cin, ci = col % Cin;
      int r = tap / S, s = tap % S;
      atomicAdd(dw + (((int64_t)co * Cin + ci) * R + r) * S + s, acc[i][j]);
    }
  }
}

int conv2d_fwd_direct(const pvg_conv_desc* d, const float* x, const float* w, const float* bias, float* y, cudaStream_t st);
int conv2d_wgrad_direct(const pvg_conv_desc* d, int Cin_logical, const float* x, const float* dy, float* dw, cudaStream_t st);

int conv2d_fwd_simt(const pvg_conv_desc* d, const float* x, const float* w, const float* bias, float* y, cudaStream_t st) {
  // image-facing layers (3/12 channels on one side) have dedicated direct kernels (conv_direct.cu)
  const int taken = conv2d_fwd_direct(d, x, w, bias, y, st);
  if (taken != 0) return taken < 0 ? taken : 0;
  int64_t M = (int64_t)d->N * d->H * d->W;
  if (d->Cout <= 16) {
    dim3 grid((unsigned)ceil_div64(M, BM), ceil_div(d->Cout, 16));
    conv_fwd_simt_kernel<16><<<grid, THREADS, 0, st>>>(x, w, bias, y, d->N, d->H, d->W, d->Cin, d->Cout, d->R, d->S, d->pad,
                                                       d->act, d->slope);
  } else {
    dim3 grid((unsigned)ceil_div64(M, BM), ceil_div(d->Cout, 64));
    conv_fwd_simt_kernel<64><<<grid, THREADS, 0, st>>>(x, w, bias, y, d->N, d->H, d->W, d->Cin, d->Cout, d->R, d->S, d->pad,
                                                       d->act, d->slope);
  }
  PVG_LAUNCH_OK();
  return 0;
}

}  // namespace pvg

using namespace pvg;

extern "C" int pvg_conv2d_wgrad(const pvg_conv_desc* d, int Cin_logical, const float* x, const float* dy, float* dw_oihw,
                                void* stream) {
  PVG_CHECK_ARG(d && x && dy && dw_oihw, "null argument");
  PVG_CHECK_ARG(Cin_logical >= 1 && Cin_logical <= d->Cin, "Cin_logical out of range");
  {  // 7x7 tanh head (Cout = 3): tiled direct kernel (conv_direct.cu)
    const int taken = conv2d_wgrad_direct(d, Cin_logical, x, dy, dw_oihw, (cudaStream_t)stream);
    if (taken != 0) return taken < 0 ? taken : 0;
  }
  int64_t M = (int64_t)d->N * d->H * d->W;
  int gx = ceil_div(d->Cout, 64), gy = ceil_div(d->R * d->S * Cin_logical, 64);
  int64_t want_splits = ceil_div64((int64_t)kSMs * 4, (int64_t)gx * gy);
  int64_t max_splits = ceil_div64(M, 256);
  int64_t splits = want_splits < max_splits ? want_splits : max_splits;
  if (splits < 1) splits = 1;
  int64_t pps = ceil_div64(ceil_div64(M, splits), BK) * BK;
  splits = ceil_div64(M, pps);
  dim3 grid(gx, gy, (unsigned)splits);
  conv_wgrad_simt_kernel<<<grid, THREADS, 0, (cudaStream_t)stream>>>(x, dy, dw_oihw, d->N, d->H, d->W, d->Cin, Cin_logical,
                                                                    d->Cout, d->R, d->S, d->pad, pps);
  PVG_LAUNCH_OK();
  return 0;
}
