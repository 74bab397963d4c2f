// HBM-bound kernels of the CADDY hot path: BatchNorm (training/eval, forward/backward, optional fused avg-pool and
// residual/activation), bilinear resampling, max-pool, ConvLSTM point-wise cell, L1-type reductions, Adam, TF32 split,
// weight packing.  All tensors NHWC fp32.  Every kernel is a coalesced, float4-vectorised (when C % 4 == 0)
// grid-stride loop sized in multiples of the SM count; reductions accumulate in fp64 and finish with one atomic per
// CTA and channel.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace pvg {

thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

// ---------------------------------------------------------------------------------------------------------------
// small generic vector helpers: V = 4 (float4) or 1 (scalar) channel units
// ---------------------------------------------------------------------------------------------------------------
template <int V> struct Vec;
template <> struct Vec<4> {
  float v[4];
  __device__ static Vec load(const float* p) { float4 t = ldg4(p); return Vec{{t.x, t.y, t.z, t.w}}; }
  __device__ void store(float* p) const { stg4(p, make_float4(v[0], v[1], v[2], v[3])); }
};
template <> struct Vec<1> {
  float v[1];
  __device__ static Vec load(const float* p) { return Vec{{__ldg(p)}}; }
  __device__ void store(float* p) const { *p = v[0]; }
};

// hi != NULL: hi = rna_tf32(x), lo = x - hi (robust to any tensor-core input rounding).
// hi == NULL: lo = x - trunc_tf32(x); valid when the tensor core truncates raw fp32 operands, x itself is then "hi".
__device__ __forceinline__ float tf32_part(float v, bool rna) {
  return rna ? tf32_hi(v) : __uint_as_float(__float_as_uint(v) & 0xffffe000u);
}
// 16-bit correction operands of the nprod == 2 product: planes[0][i] = f16((x - trunc_tf32(x)) * 2^12), planes[1][i] = f16(x)
// with f16 = bf16 (round to nearest) or fp16 (round to nearest, saturating at +-65504).  The 2^12 keeps the residual in
// fp16's normal range; the tensor-core kernels fold it back (2^-12) when they add the correction accumulator.
// fp16 carries 11 significant bits: it holds the tf32 "hi" plane of a WEIGHT exactly, so the weight side adds no error
// that is coherent over the batch (with bf16 weights that error measured 30x the fp32 CPU oracle's on cancellation-heavy
// gradients); bf16 keeps fp32's exponent range and is what gradients (1e-6..1e-12) need.  kind::f16 cannot mix the two
// formats in one MMA (illegal instruction on sm_100a), so a conv uses one format for both operands.
// FMT: PVG_CORR_BF16 (0) | PVG_CORR_FP16 (1) | PVG_CORR_FP16_ALL (2, fp16 planes whose residual is taken w.r.t. f16(x), so
// that the plane pair alone carries x to 22 bits: ALL three products of the split then run as kind::f16 MMAs on the planes
// and the fp32 tensor is not read by the convolution at all)
template <int FMT>
__device__ __forceinline__ uint32_t f16x2_bits(float a, float b) {
  if (FMT != PVG_CORR_BF16) {
    return pack_f16x2_sat(a, b);
  }
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// the part of v that the "hi" operand of the main product carries: trunc_tf32(v) (the tensor core truncates the raw fp32
// operand) or, in the all-fp16 evaluation, f16(v)
template <int FMT>
__device__ __forceinline__ float hi_part(float v) {
  if (FMT == PVG_CORR_FP16_ALL) return f16_round_sat(v);
  return tf32_part(v, false);
}
template <int FMT>
__device__ __forceinline__ void store_16_planes(uint16_t* planes, int64_t n, int64_t i4, float4 v) {
  const float4 h = make_float4(hi_part<FMT>(v.x), hi_part<FMT>(v.y), hi_part<FMT>(v.z), hi_part<FMT>(v.w));
  constexpr float kS = 4096.f;
  uint2 lo = make_uint2(f16x2_bits<FMT>((v.x - h.x) * kS, (v.y - h.y) * kS), f16x2_bits<FMT>((v.z - h.z) * kS, (v.w - h.w) * kS));
  uint2 xb = make_uint2(f16x2_bits<FMT>(v.x, v.y), f16x2_bits<FMT>(v.z, v.w));
  *reinterpret_cast<uint2*>(planes + 4 * i4) = lo;
  *reinterpret_cast<uint2*>(planes + n + 4 * i4) = xb;
}

// Optional 16-bit plane outputs of a producer kernel: up to two plane pairs (e.g. fp16 "all" planes for the next forward conv
// and bf16 planes for its weight gradient) of the fp32 tensor it writes, so that no separate pvg_split_16 pass re-reads it.
struct PlaneOut {
  uint16_t* p[2];
  int fmt[2];
  int64_t n;            // elements of the tensor (offset of the second plane of a pair)
};
__device__ __forceinline__ void emit_planes4(const PlaneOut& po, int64_t elem, float4 v) {      // elem % 4 == 0
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (po.p[k] == nullptr) continue;
    if (po.fmt[k] == PVG_CORR_FP16_ALL) store_16_planes<2>(po.p[k], po.n, elem >> 2, v);
    else if (po.fmt[k] == PVG_CORR_FP16) store_16_planes<1>(po.p[k], po.n, elem >> 2, v);
    else store_16_planes<0>(po.p[k], po.n, elem >> 2, v);
  }
}
template <int V>
__device__ __forceinline__ void emit_planes(const PlaneOut& po, int64_t elem, const Vec<V>& o) {
  if (V == 4) emit_planes4(po, elem, make_float4(o.v[0], o.v[V > 1 ? 1 : 0], o.v[V > 2 ? 2 : 0], o.v[V > 3 ? 3 : 0]));
}
static PlaneOut no_planes() { PlaneOut po; po.p[0] = po.p[1] = nullptr; po.fmt[0] = po.fmt[1] = 0; po.n = 0; return po; }
static PlaneOut make_planes(void* a, int fa, void* b, int fb, int64_t n) {
  PlaneOut po; po.p[0] = (uint16_t*)a; po.p[1] = (uint16_t*)b; po.fmt[0] = fa; po.fmt[1] = fb; po.n = n; return po;
}

constexpr int kRedThreads = 256;

// Per-channel reduction of K quantities over the rows of one batch group.  `f(row, unit, acc)` adds the
// contributions of one V-wide channel unit of one row into acc[K][V] (doubles).
template <int V, int K, class F>
__device__ void channel_reduce(int64_t row_begin, int64_t row_end, int C, double* out /*[K][C]*/, F f) {
  const int U = C / V;
  const int lanes = kRedThreads / U;       // >= 1 (checked on the host)
  const int unit = threadIdx.x % U;
  const int lane = threadIdx.x / U;
  double acc[K][V];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int j = 0; j < V; ++j) acc[k][j] = 0.0;
  if (lane < lanes) {
    const int64_t st = (int64_t)gridDim.x * lanes;
    int64_t r = row_begin + (int64_t)blockIdx.x * lanes + lane;
    for (; r + 3 * st < row_end; r += 4 * st) {      // four rows per trip: four times the loads in flight per thread
      f(r, unit, acc);
      f(r + st, unit, acc);
      f(r + 2 * st, unit, acc);
      f(r + 3 * st, unit, acc);
    }
    for (; r < row_end; r += st) f(r, unit, acc);
  }
  // sum over lanes in shared memory, one (k, j) slice at a time to bound the footprint at 2 KB
  __shared__ double slice[kRedThreads];
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      slice[threadIdx.x] = (lane < lanes) ? acc[k][j] : 0.0;
      __syncthreads();
      if (threadIdx.x < U) {
        double s = 0.0;
        for (int l = 0; l < lanes; ++l) s += slice[l * U + threadIdx.x];
        atomicAdd(&out[(size_t)k * C + threadIdx.x * V + j], s);
      }
      __syncthreads();
    }
  }
}

// Row walk of the element-wise BatchNorm kernels: a thread keeps ONE V-wide channel unit (so its per-channel constants live in
// registers: `reload(g, c)` runs once per batch group the thread meets) and strides over the rows.  The first version decoded
// (row, channel, group) from a flat index with three 64-bit divisions and re-read up to 16 per-channel constants for every
// float4: the kernels ran at 1.5 - 2.2 TB/s (r02 launch list) while max-pool, with the same traffic pattern, reaches 6.
template <int N_>
struct RowCount { static constexpr int value = N_; };

// `body(r, step, c, RowCount<R>{})` handles the R rows r, r + step, ... of one batch group: it issues the loads of all of them
// before it computes and stores (R x the bytes in flight per thread; R = 1 for the tail and around group boundaries).
template <int V, int ROWS, class Reload, class Body>
__device__ __forceinline__ void walk_rows(int C, int64_t M, int64_t rows_per_group, Reload reload, Body body) {
  const int U = C / V;
  const int lanes = blockDim.x / U;
  const int unit = threadIdx.x % U, lane = threadIdx.x / U;
  if (lane >= lanes) return;
  const int c = unit * V;
  int64_t r = (int64_t)blockIdx.x * lanes + lane;
  const int64_t step = (int64_t)gridDim.x * lanes;
  if (r >= M) return;
  int g = (int)(r / rows_per_group);
  int64_t bound = (int64_t)(g + 1) * rows_per_group;
  reload(g, c);
  while (r < M) {
    if (r >= bound) {
      do { ++g; bound += rows_per_group; } while (r >= bound);
      reload(g, c);
    }
    const int64_t last = r + (ROWS - 1) * step;
    if (ROWS > 1 && last < M && last < bound) {
      body(r, step, c, RowCount<ROWS>{});
      r += ROWS * step;
    } else {
      body(r, step, c, RowCount<1>{});
      r += step;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// BatchNorm statistics
// ---------------------------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(kRedThreads) bn_stats_kernel(const float* __restrict__ x, int64_t rows_per_group,
                                                               int C, double* __restrict__ sums) {
  const int g = blockIdx.y;
  const int64_t r0 = (int64_t)g * rows_per_group;
  channel_reduce<V, 2>(r0, r0 + rows_per_group, C, sums + (size_t)g * 2 * C,
                       [&](int64_t r, int unit, double (*acc)[V]) {
                         Vec<V> t = Vec<V>::load(x + r * C + unit * V);
#pragma unroll
                         for (int j = 0; j < V; ++j) { acc[0][j] += t.v[j]; acc[1][j] += (double)t.v[j] * t.v[j]; }
                       });
}

template <int V>
__global__ void __launch_bounds__(kRedThreads) pool2_stats_kernel(const float* __restrict__ x, int H, int W, int C,
                                                                  int64_t rows_per_group, float* __restrict__ y,
                                                                  double* __restrict__ sums) {
  const int g = blockIdx.y;
  const int OH = H / 2, OW = W / 2;
  const int64_t r0 = (int64_t)g * rows_per_group;
  channel_reduce<V, 2>(r0, r0 + rows_per_group, C, sums + (size_t)g * 2 * C,
                       [&](int64_t r, int unit, double (*acc)[V]) {
                         int ow = (int)(r % OW);
                         int64_t t = r / OW;
                         int oh = (int)(t % OH);
                         int64_t n = t / OH;
                         const float* p = x + ((n * H + 2 * oh) * W + 2 * ow) * C + unit * V;
                         Vec<V> a = Vec<V>::load(p), b = Vec<V>::load(p + C);
                         Vec<V> c = Vec<V>::load(p + (int64_t)W * C), d = Vec<V>::load(p + (int64_t)W * C + C);
                         Vec<V> o;
#pragma unroll
                         for (int j = 0; j < V; ++j) {
                           o.v[j] = (a.v[j] + b.v[j] + c.v[j] + d.v[j]) * 0.25f;
                           acc[0][j] += o.v[j];
                           acc[1][j] += (double)o.v[j] * o.v[j];
                         }
                         o.store(y + r * C + unit * V);
                       });
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, int groups, int C, float eps,
                                   float momentum, float* running_mean, float* running_var, float* __restrict__ mean,
                                   float* __restrict__ invstd) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float rm = running_mean ? running_mean[c] : 0.f, rv = running_var ? running_var[c] : 0.f;
  for (int g = 0; g < groups; ++g) {
    double s = sums[((size_t)g * 2 + 0) * C + c], ss = sums[((size_t)g * 2 + 1) * C + c];
    double m = s / count;
    double var = ss / count - m * m;
    if (var < 0.0) var = 0.0;
    mean[(size_t)g * C + c] = (float)m;
    invstd[(size_t)g * C + c] = (float)(1.0 / sqrt(var + (double)eps));
    double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    rm = (1.f - momentum) * rm + momentum * (float)m;
    rv = (1.f - momentum) * rv + momentum * (float)unbiased;
  }
  if (running_mean) running_mean[c] = rm;
  if (running_var) running_var[c] = rv;
}

__global__ void bn_eval_prepare_kernel(const float* rm, const float* rv, int C, float eps, float* mean, float* invstd) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  mean[c] = rm[c];
  invstd[c] = 1.f / sqrtf(rv[c] + eps);
}

template <int V>
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, int64_t M, int64_t rows_per_group,
                                                       int C, const float* __restrict__ mean,
                                                       const float* __restrict__ invstd, const float* __restrict__ weight,
                                                       const float* __restrict__ bias, const float* __restrict__ residual,
                                                       int act, float slope, float* __restrict__ y, const PlaneOut po,
                                                       const int Cp) {
  // Cp <= C: channels that HAVE parameters (a 65-channel BatchNorm on a tensor physically padded to 72 channels so that its
  // convolutions run on the tensor cores); the padding channels are zero in, zero out
  float km[V], ki[V], kw[V], kb[V];
  walk_rows<V, 4>(
      C, M, rows_per_group,
      [&](int g, int c) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const bool real = c + j < Cp;
          kw[j] = real ? (weight ? __ldg(weight + c + j) : 1.f) : 0.f;
          kb[j] = (real && bias) ? __ldg(bias + c + j) : 0.f;
          km[j] = __ldg(mean + (size_t)g * C + c + j);
          ki[j] = __ldg(invstd + (size_t)g * C + c + j);
        }
      },
      [&](int64_t r, int64_t step, int c, auto rows) {
        constexpr int R = decltype(rows)::value;
        Vec<V> t[R], res[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
          t[i] = Vec<V>::load(x + (r + i * step) * C + c);
          if (residual) res[i] = Vec<V>::load(residual + (r + i * step) * C + c);
        }
#pragma unroll
        for (int i = 0; i < R; ++i) {
          Vec<V> o;
#pragma unroll
          for (int j = 0; j < V; ++j) {
            float v = (t[i].v[j] - km[j]) * ki[j] * kw[j] + kb[j];
            if (residual) v += res[i].v[j];
            o.v[j] = act_fwd(v, act, slope);
          }
          o.store(y + (r + i * step) * C + c);
          emit_planes<V>(po, (r + i * step) * C + c, o);
        }
      });
}

template <int V>
__global__ void __launch_bounds__(kRedThreads) bn_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                                    const float* __restrict__ x, int64_t rows_per_group,
                                                                    int C, const float* __restrict__ mean,
                                                                    const float* __restrict__ invstd, int act, float slope,
                                                                    double* __restrict__ sums2) {
  const int g = blockIdx.y;
  const int64_t r0 = (int64_t)g * rows_per_group;
  float km[V], ki[V];                     // this thread's channel unit is fixed (channel_reduce): its statistics live in registers
  {
    const int c = (int)(threadIdx.x % (C / V)) * V;
#pragma unroll
    for (int j = 0; j < V; ++j) { km[j] = __ldg(mean + (size_t)g * C + c + j); ki[j] = __ldg(invstd + (size_t)g * C + c + j); }
  }
  channel_reduce<V, 2>(r0, r0 + rows_per_group, C, sums2 + (size_t)g * 2 * C,
                       [&](int64_t r, int unit, double (*acc)[V]) {
                         int c = unit * V;
                         Vec<V> d = Vec<V>::load(dy + r * C + c), xv = Vec<V>::load(x + r * C + c), yv;
                         if (act != PVG_ACT_NONE) yv = Vec<V>::load(y + r * C + c);
#pragma unroll
                         for (int j = 0; j < V; ++j) {
                           float gg = d.v[j] * (act != PVG_ACT_NONE ? act_bwd_from_out(yv.v[j], act, slope) : 1.f);
                           float xhat = (xv.v[j] - km[j]) * ki[j];
                           acc[0][j] += gg;
                           acc[1][j] += (double)gg * xhat;
                         }
                       });
}

// bn_finalize + bn_apply in one launch (training mode): every CTA derives mean / invstd of all (group, channel) pairs from
// the fp64 sums into shared memory (groups * C divisions + rsqrt: negligible next to the apply pass), CTA 0 also publishes
// them for the backward pass and replays the running-statistics updates.  Saves one launch per BatchNorm call.
template <int V>
__global__ void __launch_bounds__(256) bn_finalize_apply_kernel(const float* __restrict__ x, int64_t M, int64_t rows_per_group,
                                                                int C, int groups, const double* __restrict__ sums, double count,
                                                                float eps, float momentum, float* running_mean,
                                                                float* running_var, float* __restrict__ mean,
                                                                float* __restrict__ invstd, const float* __restrict__ weight,
                                                                const float* __restrict__ bias, const float* __restrict__ residual,
                                                                int act, float slope, float* __restrict__ y, const PlaneOut po,
                                                                const int Cp) {
  extern __shared__ float sm_stats[];              // [groups][C] mean, then [groups][C] invstd
  float* s_mean = sm_stats;
  float* s_inv = sm_stats + (size_t)groups * C;
  for (int i = threadIdx.x; i < groups * C; i += blockDim.x) {
    const int g = i / C, c = i - g * C;
    const double s = sums[((size_t)g * 2 + 0) * C + c], ss = sums[((size_t)g * 2 + 1) * C + c];
    const double m = s / count;
    double var = ss / count - m * m;
    if (var < 0.0) var = 0.0;
    const float mf = (float)m, isf = (float)(1.0 / sqrt(var + (double)eps));
    s_mean[i] = mf; s_inv[i] = isf;
    if (blockIdx.x == 0) { mean[i] = mf; invstd[i] = isf; }
  }
  if (blockIdx.x == 0 && (running_mean || running_var)) {
    for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
      float rm = running_mean ? running_mean[c] : 0.f, rv = running_var ? running_var[c] : 0.f;
      for (int g = 0; g < groups; ++g) {
        const double s = sums[((size_t)g * 2 + 0) * C + c], ss = sums[((size_t)g * 2 + 1) * C + c];
        const double m = s / count;
        double var = ss / count - m * m;
        if (var < 0.0) var = 0.0;
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        rm = (1.f - momentum) * rm + momentum * (float)m;
        rv = (1.f - momentum) * rv + momentum * (float)unbiased;
      }
      if (running_mean) running_mean[c] = rm;
      if (running_var) running_var[c] = rv;
    }
  }
  __syncthreads();
  float km[V], ki[V], kw[V], kb[V];
  walk_rows<V, 4>(
      C, M, rows_per_group,
      [&](int g, int c) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const bool real = c + j < Cp;
          kw[j] = real ? (weight ? __ldg(weight + c + j) : 1.f) : 0.f;
          kb[j] = (real && bias) ? __ldg(bias + c + j) : 0.f;
          km[j] = s_mean[(size_t)g * C + c + j];
          ki[j] = s_inv[(size_t)g * C + c + j];
        }
      },
      [&](int64_t r, int64_t step, int c, auto rows) {
        constexpr int R = decltype(rows)::value;
        Vec<V> t[R], res[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
          t[i] = Vec<V>::load(x + (r + i * step) * C + c);
          if (residual) res[i] = Vec<V>::load(residual + (r + i * step) * C + c);
        }
#pragma unroll
        for (int i = 0; i < R; ++i) {
          Vec<V> o;
#pragma unroll
          for (int j = 0; j < V; ++j) {
            float v = (t[i].v[j] - km[j]) * ki[j] * kw[j] + kb[j];
            if (residual) v += res[i].v[j];
            o.v[j] = act_fwd(v, act, slope);
          }
          o.store(y + (r + i * step) * C + c);
          emit_planes<V>(po, (r + i * step) * C + c, o);
        }
      });
}

template <int V>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                           const float* __restrict__ x, int64_t M, int64_t rows_per_group,
                                                           int OH, int OW, int C, const float* __restrict__ mean,
                                                           const float* __restrict__ invstd, const float* __restrict__ weight,
                                                           int act, float slope, const double* __restrict__ sums2, int eval,
                                                           int unpool, float* __restrict__ dx, float* __restrict__ g_out,
                                                           int groups, float* __restrict__ dweight, float* __restrict__ dbias,
                                                           const int Cp, uint32_t* __restrict__ amax_out) {
  if (blockIdx.x == 0 && (dweight || dbias)) {       // bn_bwd_params folded in: one launch less per BatchNorm backward
    for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
      double sg = 0.0, sgx = 0.0;
      for (int g = 0; g < groups; ++g) { sg += sums2[((size_t)g * 2 + 0) * C + c]; sgx += sums2[((size_t)g * 2 + 1) * C + c]; }
      if (dweight) dweight[c] = (float)sgx;
      if (dbias) dbias[c] = (float)sg;
    }
  }
  const double inv_count = 1.0 / (double)rows_per_group;
  float amax = 0.f;              // largest |dx| this thread writes (the 0.25 of the un-pooling is applied at the end)
  float km[V], ki[V], kw[V], ksg[V], ksgx[V];
  const bool small = M < (int64_t)0x7fffffff;          // 32-bit divisions for the un-pooling coordinates
  walk_rows<V, 2>(
      C, M, rows_per_group,
      [&](int g, int c) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
          ki[j] = __ldg(invstd + (size_t)g * C + c + j);
          kw[j] = (c + j < Cp) ? (weight ? __ldg(weight + c + j) : 1.f) : 0.f;
          km[j] = __ldg(mean + (size_t)g * C + c + j);
          ksg[j] = eval ? 0.f : (float)(sums2[((size_t)g * 2 + 0) * C + c + j] * inv_count);
          ksgx[j] = eval ? 0.f : (float)(sums2[((size_t)g * 2 + 1) * C + c + j] * inv_count);
        }
      },
      [&](int64_t r0, int64_t step, int c, auto rows) {
        constexpr int R = decltype(rows)::value;
        Vec<V> d[R], xv[R], yv[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
          const int64_t r = r0 + i * step;
          d[i] = Vec<V>::load(dy + r * C + c);
          xv[i] = Vec<V>::load(x + r * C + c);
          if (act != PVG_ACT_NONE) yv[i] = Vec<V>::load(y + r * C + c);
        }
#pragma unroll
        for (int i = 0; i < R; ++i) {
          const int64_t r = r0 + i * step;
          Vec<V> o, go;
#pragma unroll
          for (int j = 0; j < V; ++j) {
            float gg = d[i].v[j] * (act != PVG_ACT_NONE ? act_bwd_from_out(yv[i].v[j], act, slope) : 1.f);
            go.v[j] = gg;
            float val;
            if (eval) {
              val = kw[j] * ki[j] * gg;
            } else {
              float xhat = (xv[i].v[j] - km[j]) * ki[j];
              val = kw[j] * ki[j] * (gg - ksg[j] - xhat * ksgx[j]);
            }
            o.v[j] = val;
            amax = fmaxf(amax, fabsf(val));
          }
          if (g_out) go.store(g_out + r * C + c);
          if (!unpool) {
            o.store(dx + r * C + c);
          } else {
#pragma unroll
            for (int j = 0; j < V; ++j) o.v[j] *= 0.25f;
            int ow, oh;
            int64_t n;
            if (small) {
              const uint32_t r32 = (uint32_t)r, t32 = r32 / (uint32_t)OW;
              ow = (int)(r32 - t32 * (uint32_t)OW);
              n = t32 / (uint32_t)OH;
              oh = (int)(t32 - (uint32_t)n * (uint32_t)OH);
            } else {
              ow = (int)(r % OW);
              const int64_t t = r / OW;
              oh = (int)(t % OH);
              n = t / OH;
            }
            const int W = OW * 2, H = OH * 2;
            float* p = dx + ((n * H + 2 * oh) * W + 2 * ow) * C + c;
            o.store(p); o.store(p + C); o.store(p + (int64_t)W * C); o.store(p + (int64_t)W * C + C);
          }
        }
      });
  if (amax_out != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0 && amax == amax) atomicMax(amax_out, __float_as_uint(unpool ? 0.25f * amax : amax));
  }
}

__global__ void bn_bwd_params_kernel(const double* sums2, int groups, int C, float* dweight, float* dbias) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double sg = 0.0, sgx = 0.0;
  for (int g = 0; g < groups; ++g) { sg += sums2[((size_t)g * 2 + 0) * C + c]; sgx += sums2[((size_t)g * 2 + 1) * C + c]; }
  if (dweight) dweight[c] = (float)sgx;
  if (dbias) dbias[c] = (float)sg;
}

// ---------------------------------------------------------------------------------------------------------------
// resampling
// ---------------------------------------------------------------------------------------------------------------
struct Lerp { int i0, i1; float l; };
// PyTorch area_pixel_compute_source_index(align_corners=False) + clamp to 0, as used by upsample_bilinear2d
__device__ __forceinline__ Lerp src_index(int dst, float scale, int in_size) {
  float s = scale * (dst + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  int i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  Lerp r;
  r.i0 = i0;
  r.i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  r.l = s - (float)i0;
  return r;
}

template <int V>
__global__ void __launch_bounds__(256) resize_bilinear_kernel(const float* __restrict__ x, int N, int H, int W, int C,
                                                              float* __restrict__ y, int OH, int OW, float sh, float sw,
                                                              const PlaneOut po) {
  const int U = C / V;
  const int64_t total = (int64_t)N * OH * OW * U;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % U) * V;
    int64_t r = i / U;
    int ox = (int)(r % OW); r /= OW;
    int oy = (int)(r % OH);
    int64_t n = r / OH;
    Lerp ly = src_index(oy, sh, H), lx = src_index(ox, sw, W);
    const float* b = x + n * H * W * C + c;
    Vec<V> v00 = Vec<V>::load(b + ((int64_t)ly.i0 * W + lx.i0) * C), v01 = Vec<V>::load(b + ((int64_t)ly.i0 * W + lx.i1) * C);
    Vec<V> v10 = Vec<V>::load(b + ((int64_t)ly.i1 * W + lx.i0) * C), v11 = Vec<V>::load(b + ((int64_t)ly.i1 * W + lx.i1) * C);
    Vec<V> o;
    float w0y = 1.f - ly.l, w1y = ly.l, w0x = 1.f - lx.l, w1x = lx.l;
#pragma unroll
    for (int j = 0; j < V; ++j)
      o.v[j] = w0y * (w0x * v00.v[j] + w1x * v01.v[j]) + w1y * (w0x * v10.v[j] + w1x * v11.v[j]);
    o.store(y + i * V);
    emit_planes<V>(po, i * V, o);
  }
}

// gradient of the x2 bilinear upsample: gather form (each input pixel visits the <= 4x4 outputs that read it)
template <int V>
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const float* __restrict__ dy, int N, int H, int W, int C,
                                                             float* __restrict__ dx) {
  const int U = C / V;
  const int OH = 2 * H, OW = 2 * W;
  const int64_t total = (int64_t)N * H * W * U;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % U) * V;
    int64_t r = i / U;
    int ix = (int)(r % W); r /= W;
    int iy = (int)(r % H);
    int64_t n = r / H;
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
    for (int oy = 2 * iy - 1; oy <= 2 * iy + 2; ++oy) {
      if (oy < 0 || oy >= OH) continue;
      Lerp ly = src_index(oy, 0.5f, H);
      float wy = (ly.i0 == iy ? 1.f - ly.l : 0.f) + (ly.i1 == iy ? ly.l : 0.f);
      if (wy == 0.f) continue;
      for (int ox = 2 * ix - 1; ox <= 2 * ix + 2; ++ox) {
        if (ox < 0 || ox >= OW) continue;
        Lerp lx = src_index(ox, 0.5f, W);
        float wx = (lx.i0 == ix ? 1.f - lx.l : 0.f) + (lx.i1 == ix ? lx.l : 0.f);
        if (wx == 0.f) continue;
        Vec<V> d = Vec<V>::load(dy + ((n * OH + oy) * OW + ox) * C + c);
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] += wy * wx * d.v[j];
      }
    }
    Vec<V> o;
#pragma unroll
    for (int j = 0; j < V; ++j) o.v[j] = acc[j];
    o.store(dx + i * V);
  }
}

template <int V>
__global__ void __launch_bounds__(256) maxpool2_fwd_kernel(const float* __restrict__ x, int N, int H, int W, int C,
                                                           float* __restrict__ y, const PlaneOut po) {
  const int U = C / V, OH = H / 2, OW = W / 2;
  const int64_t total = (int64_t)N * OH * OW * U;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % U) * V;
    int64_t r = i / U;
    int ox = (int)(r % OW); r /= OW;
    int oy = (int)(r % OH);
    int64_t n = r / OH;
    const float* p = x + ((n * H + 2 * oy) * W + 2 * ox) * C + c;
    Vec<V> a = Vec<V>::load(p), b = Vec<V>::load(p + C), cc = Vec<V>::load(p + (int64_t)W * C), d = Vec<V>::load(p + (int64_t)W * C + C), o;
#pragma unroll
    for (int j = 0; j < V; ++j) o.v[j] = fmaxf(fmaxf(a.v[j], b.v[j]), fmaxf(cc.v[j], d.v[j]));
    o.store(y + i * V);
    emit_planes<V>(po, i * V, o);
  }
}

template <int V>
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                           const float* __restrict__ y, int N, int H, int W, int C,
                                                           int relu_mask, float* __restrict__ dx) {
  const int U = C / V, OH = H / 2, OW = W / 2;
  const int64_t total = (int64_t)N * OH * OW * U;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % U) * V;
    int64_t r = i / U;
    int ox = (int)(r % OW); r /= OW;
    int oy = (int)(r % OH);
    int64_t n = r / OH;
    int64_t base = ((n * H + 2 * oy) * W + 2 * ox) * C + c;
    int64_t off[4] = {0, C, (int64_t)W * C, (int64_t)W * C + C};
    Vec<V> m = Vec<V>::load(y + i * V), d = Vec<V>::load(dy + i * V);
    Vec<V> in[4], out[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) in[k] = Vec<V>::load(x + base + off[k]);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      bool taken = false;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        bool hit = !taken && in[k].v[j] == m.v[j];
        taken = taken || hit;
        float gval = hit ? d.v[j] : 0.f;
        if (relu_mask && !(in[k].v[j] > 0.f)) gval = 0.f;
        out[k].v[j] = gval;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) out[k].store(dx + base + off[k]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// ConvLSTM cell point-wise part
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

template <int V>
__global__ void __launch_bounds__(256) lstm_fwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev,
                                                       int64_t M, int C, float* __restrict__ c_new, float* __restrict__ h_new) {
  const int U = C / V;
  const int64_t total = M * U;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / U;
    int c = (int)(i % U) * V;
    const float* gp = gates + r * 4 * C + c;
    Vec<V> gi = Vec<V>::load(gp), gf = Vec<V>::load(gp + C), go = Vec<V>::load(gp + 2 * C), gc = Vec<V>::load(gp + 3 * C);
    Vec<V> cp = Vec<V>::load(c_prev + r * C + c), cn, hn;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float ii = sigmoidf_(gi.v[j]), ff = sigmoidf_(gf.v[j]), oo = sigmoidf_(go.v[j]), cc = tanhf(gc.v[j]);
      cn.v[j] = ff * cp.v[j] + ii * cc;
      hn.v[j] = oo * tanhf(cn.v[j]);
    }
    cn.store(c_new + r * C + c);
    hn.store(h_new + r * C + c);
  }
}

template <int V>
__global__ void __launch_bounds__(256) lstm_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev,
                                                       const float* __restrict__ c_new, const float* __restrict__ dh,
                                                       const float* __restrict__ dc_new, int64_t M, int C,
                                                       float* __restrict__ dgates, float* __restrict__ dc_prev) {
  const int U = C / V;
  const int64_t total = M * U;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / U;
    int c = (int)(i % U) * V;
    const float* gp = gates + r * 4 * C + c;
    Vec<V> gi = Vec<V>::load(gp), gf = Vec<V>::load(gp + C), go = Vec<V>::load(gp + 2 * C), gc = Vec<V>::load(gp + 3 * C);
    Vec<V> cp = Vec<V>::load(c_prev + r * C + c), cn = Vec<V>::load(c_new + r * C + c);
    Vec<V> dhv, dcn, di, df, dog, dg, dcp;
    if (dh) dhv = Vec<V>::load(dh + r * C + c);
    if (dc_new) dcn = Vec<V>::load(dc_new + r * C + c);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float ii = sigmoidf_(gi.v[j]), ff = sigmoidf_(gf.v[j]), oo = sigmoidf_(go.v[j]), cc = tanhf(gc.v[j]);
      float tc = tanhf(cn.v[j]);
      float dhj = dh ? dhv.v[j] : 0.f;
      float dc = (dc_new ? dcn.v[j] : 0.f) + dhj * oo * (1.f - tc * tc);
      dog.v[j] = dhj * tc * oo * (1.f - oo);
      di.v[j] = dc * cc * ii * (1.f - ii);
      df.v[j] = dc * cp.v[j] * ff * (1.f - ff);
      dg.v[j] = dc * ii * (1.f - cc * cc);
      dcp.v[j] = dc * ff;
    }
    float* dp = dgates + r * 4 * C + c;
    di.store(dp); df.store(dp + C); dog.store(dp + 2 * C); dg.store(dp + 3 * C);
    dcp.store(dc_prev + r * C + c);
  }
}

// backward of the fused ConvLSTM epilogue (conv_h3.cu): activated gates in interleaved order [M][C][4] = (i, f, o, g)
__global__ void __launch_bounds__(256) lstm_bwd_act_kernel(const float* __restrict__ ga, const float* __restrict__ c_prev,
                                                           const float* __restrict__ c_new, const float* __restrict__ dh,
                                                           const float* __restrict__ dc_new, int64_t total,
                                                           float* __restrict__ dgates, float* __restrict__ dc_prev) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 g = ldg4(ga + 4 * i);
    const float ii = g.x, ff = g.y, oo = g.z, cc = g.w;
    const float tc = tanhf(__ldg(c_new + i));
    const float dhj = dh ? __ldg(dh + i) : 0.f;
    const float dc = (dc_new ? __ldg(dc_new + i) : 0.f) + dhj * oo * (1.f - tc * tc);
    stg4(dgates + 4 * i, make_float4(dc * cc * ii * (1.f - ii), dc * __ldg(c_prev + i) * ff * (1.f - ff), dhj * tc * oo * (1.f - oo),
                                     dc * ii * (1.f - cc * cc)));
    dc_prev[i] = dc * ff;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// L1-type reductions
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) absdiff_mean_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                               int64_t count, double* __restrict__ out) {
  const int n = blockIdx.y;
  const float* pa = a + (int64_t)n * count;
  const float* pb = b + (int64_t)n * count;
  double acc = 0.0;
  const bool vec = (count % 4 == 0) && ((((uintptr_t)pa | (uintptr_t)pb) & 15) == 0);
  if (vec) {
    int64_t q = count / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < q; i += (int64_t)gridDim.x * blockDim.x) {
      float4 x = ldg4(pa + 4 * i), y = ldg4(pb + 4 * i);
      float s = fabsf(x.x - y.x) + fabsf(x.y - y.y) + fabsf(x.z - y.z) + fabsf(x.w - y.w);
      acc += s;
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
      acc += fabsf(__ldg(pa + i) - __ldg(pb + i));
  }
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(out + n, red[0] / (double)count);
}

__global__ void __launch_bounds__(256) absdiff_mean_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                               const float* __restrict__ gout, int64_t count,
                                                               float* __restrict__ db) {
  const int n = blockIdx.y;
  const float g = __ldg(gout + n) / (float)count;
  const float* pa = a + (int64_t)n * count;
  const float* pb = b + (int64_t)n * count;
  float* pd = db + (int64_t)n * count;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    float d = __ldg(pa + i) - __ldg(pb + i);
    pd[i] = d > 0.f ? -g : (d < 0.f ? g : 0.f);      // d|a-b|/db = -sign(a-b)
  }
}

// ---------------------------------------------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi,
                                                         float* __restrict__ lo, int64_t n) {
  const bool rna = hi != nullptr;
  int64_t q = n / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < q; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = ldg4(x + 4 * i);
    float4 h = make_float4(tf32_part(v.x, rna), tf32_part(v.y, rna), tf32_part(v.z, rna), tf32_part(v.w, rna));
    if (hi) stg4(hi + 4 * i, h);
    stg4(lo + 4 * i, make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w));
  }
  for (int64_t i = q * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float h = tf32_part(x[i], rna);
    if (hi) hi[i] = h;
    lo[i] = x[i] - h;
  }
}

template <int FMT>
__global__ void __launch_bounds__(256) split_16_kernel(const float* __restrict__ x, uint16_t* __restrict__ planes, int64_t n) {
  const int64_t q = n / 4;                      // n % 8 == 0 (checked by the host wrapper)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < q; i += (int64_t)gridDim.x * blockDim.x)
    store_16_planes<FMT>(planes, n, i, ldg4(x + 4 * i));
}
// the same for a packed weight given its tf32 hi / residual lo planes (w = hi + lo exactly):
// planes = { f16(lo * 2^12), f16(hi) }, or { f16((w - f16(w)) * 2^12), f16(w) } in the all-fp16 evaluation
template <int FMT>
__global__ void __launch_bounds__(256) pack_16x2_kernel(const float* __restrict__ hi, const float* __restrict__ lo,
                                                        uint16_t* __restrict__ planes, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float h = hi[i], l = lo[i];
    if (FMT == PVG_CORR_FP16_ALL) {
      const float w = h + l;
      h = hi_part<FMT>(w);
      l = w - h;
    }
    const uint32_t v = f16x2_bits<FMT>(l * 4096.f, h);
    planes[i] = (uint16_t)(v & 0xffffu);
    planes[n + i] = (uint16_t)(v >> 16);
  }
}
template <int FMT>
__global__ void __launch_bounds__(256) act_bwd_split_16_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                               int act, float slope, float* __restrict__ g,
                                                               uint16_t* __restrict__ planes, int64_t n) {
  const int64_t q = n / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < q; i += (int64_t)gridDim.x * blockDim.x) {
    float4 d = ldg4(dy + 4 * i), o = ldg4(y + 4 * i);
    float4 v = make_float4(d.x * act_bwd_from_out(o.x, act, slope), d.y * act_bwd_from_out(o.y, act, slope),
                           d.z * act_bwd_from_out(o.z, act, slope), d.w * act_bwd_from_out(o.w, act, slope));
    stg4(g + 4 * i, v);
    store_16_planes<FMT>(planes, n, i, v);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Gradients as fp16 plane pairs: fp16 has no range for raw gradients (1e-6 .. 1e-12), so the tensor is scaled by a power of
// two S chosen from its largest magnitude, max|g| * S in [2^13, 2^14): the pair {f16((gS - f16(gS)) * 2^12), f16(gS)} then
// carries every element within 2^26 of the maximum to 22 bits (better than the 19 bits of a TF32 + bf16 pair) and smaller
// ones with an absolute error below 2^-50 of the maximum.  The consumer (data / weight gradient kernel) multiplies its result
// by 1 / S, read from device memory.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) amax_kernel(const float* __restrict__ x, int64_t n, uint32_t* __restrict__ amax_bits) {
  float m = 0.f;
  const int64_t q = n / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < q; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = ldg4(x + 4 * i);
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  for (int64_t i = q * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
    if (m == m) atomicMax(amax_bits, __float_as_uint(m));     // non-negative floats order like their bit patterns; NaN is dropped
  }
}
// S = 2^(14 - e) with amax = m * 2^e, m in [0.5, 1); 1 for an all-zero / non-finite tensor
__device__ __forceinline__ float grad_scale_from_amax(uint32_t amax_bits, float* inv) {
  const float a = __uint_as_float(amax_bits);
  if (!(a > 0.f) || !(a < 3.0e38f)) { *inv = 1.f; return 1.f; }
  int e;
  frexpf(a, &e);
  int sh = 14 - e;
  sh = sh > 120 ? 120 : (sh < -100 ? -100 : sh);
  *inv = ldexpf(1.f, -sh);
  return ldexpf(1.f, sh);
}
__global__ void __launch_bounds__(256) split_16_scaled_kernel(const float* __restrict__ x, uint16_t* __restrict__ planes, int64_t n,
                                                              const uint32_t* __restrict__ amax_bits, float* __restrict__ inv_scale) {
  float inv;
  const float S = grad_scale_from_amax(*amax_bits, &inv);
  if (blockIdx.x == 0 && threadIdx.x == 0) *inv_scale = inv;
  const int64_t q = n / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < q; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = ldg4(x + 4 * i);
    store_16_planes<2>(planes, n, i, make_float4(v.x * S, v.y * S, v.z * S, v.w * S));
  }
}
__global__ void __launch_bounds__(256) act_bwd_split_16_scaled_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                                      int act, float slope, float* __restrict__ g,
                                                                      uint16_t* __restrict__ planes, int64_t n,
                                                                      const uint32_t* __restrict__ amax_bits,
                                                                      float* __restrict__ inv_scale) {
  float inv;
  const float S = grad_scale_from_amax(*amax_bits, &inv);       // amax of dy: |dy * act'(y)| <= |dy| for every activation on the path
  if (blockIdx.x == 0 && threadIdx.x == 0) *inv_scale = inv;
  const int64_t q = n / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < q; i += (int64_t)gridDim.x * blockDim.x) {
    float4 d = ldg4(dy + 4 * i), o = ldg4(y + 4 * i);
    float4 v = make_float4(d.x * act_bwd_from_out(o.x, act, slope), d.y * act_bwd_from_out(o.y, act, slope),
                           d.z * act_bwd_from_out(o.z, act, slope), d.w * act_bwd_from_out(o.w, act, slope));
    if (g) stg4(g + 4 * i, v);
    store_16_planes<2>(planes, n, i, make_float4(v.x * S, v.y * S, v.z * S, v.w * S));
  }
}

__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, int act,
                                                      float slope, float* __restrict__ g, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    g[i] = __ldg(dy + i) * act_bwd_from_out(__ldg(y + i), act, slope);
}

// ---- a feature-matching L1 term folded into the activation backward of the conv that produced the feature ----------------
// The perceptual loss taps y = relu(conv(...)) (vgg.py:41-56) with mean|target - y| per sample (losses.py:465): y's gradient is
// dy (from the layers above) + tap_gout[n] / count * sign(y - target).  Adding the second term here replaces the
// absdiff backward kernel, autograd's accumulation of the two gradients and their three passes over the feature map each.
__device__ __forceinline__ float tap_grad(float target, float y, float gn) {
  const float d = target - y;                               // absdiff_mean_bwd_kernel: d|a-b|/db = -sign(a-b)
  return d > 0.f ? -gn : (d < 0.f ? gn : 0.f);
}
__device__ __forceinline__ float block_max_abs(const float* __restrict__ v, int n) {       // every thread returns the maximum
  __shared__ float part[32];
  float m = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(__ldg(v + i)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = m;
  __syncthreads();
  m = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, part[i]);
  return m;
}
// grid (x, N): sample n = blockIdx.y, count elements each (count % 4 == 0)
__global__ void __launch_bounds__(256) act_bwd_tap_kernel(const float* __restrict__ dy, const float* __restrict__ y, int act,
                                                          float slope, float* __restrict__ g, int64_t count,
                                                          const float* __restrict__ target, const float* __restrict__ tap_gout) {
  const int n = blockIdx.y;
  const float gn = __ldg(tap_gout + n) / (float)count;
  const int64_t base = (int64_t)n * count;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count / 4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 d = ldg4(dy + base + 4 * i), o = ldg4(y + base + 4 * i), t = ldg4(target + base + 4 * i);
    stg4(g + base + 4 * i, make_float4((d.x + tap_grad(t.x, o.x, gn)) * act_bwd_from_out(o.x, act, slope),
                                       (d.y + tap_grad(t.y, o.y, gn)) * act_bwd_from_out(o.y, act, slope),
                                       (d.z + tap_grad(t.z, o.z, gn)) * act_bwd_from_out(o.z, act, slope),
                                       (d.w + tap_grad(t.w, o.w, gn)) * act_bwd_from_out(o.w, act, slope)));
  }
}
__global__ void __launch_bounds__(256) act_bwd_tap_split_16_scaled_kernel(
    const float* __restrict__ dy, const float* __restrict__ y, int act, float slope, float* __restrict__ g,
    uint16_t* __restrict__ planes, int64_t count, int N, const uint32_t* __restrict__ amax_bits, float* __restrict__ inv_scale,
    const float* __restrict__ target, const float* __restrict__ tap_gout) {
  // scale from an upper bound of the sum: max|dy| + max_n |tap_gout[n]| / count (the same in every block)
  const float bound = __uint_as_float(*amax_bits) + block_max_abs(tap_gout, N) / (float)count;
  float inv;
  const float S = grad_scale_from_amax(__float_as_uint(bound), &inv);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *inv_scale = inv;
  const int n = blockIdx.y;
  const float gn = __ldg(tap_gout + n) / (float)count;
  const int64_t base = (int64_t)n * count, total = (int64_t)N * count;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count / 4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 d = ldg4(dy + base + 4 * i), o = ldg4(y + base + 4 * i), t = ldg4(target + base + 4 * i);
    const float4 v = make_float4((d.x + tap_grad(t.x, o.x, gn)) * act_bwd_from_out(o.x, act, slope),
                                 (d.y + tap_grad(t.y, o.y, gn)) * act_bwd_from_out(o.y, act, slope),
                                 (d.z + tap_grad(t.z, o.z, gn)) * act_bwd_from_out(o.z, act, slope),
                                 (d.w + tap_grad(t.w, o.w, gn)) * act_bwd_from_out(o.w, act, slope));
    if (g) stg4(g + base + 4 * i, v);
    store_16_planes<2>(planes, total, base / 4 + i, make_float4(v.x * S, v.y * S, v.z * S, v.w * S));
  }
}

// g = dy * act'(y) and, in the same pass, the 3xTF32 residual plane of g (lo = g - tf32(g); hi optional, see split)
__global__ void __launch_bounds__(256) act_bwd_split_kernel(const float* __restrict__ dy, const float* __restrict__ y, int act,
                                                            float slope, float* __restrict__ g, float* __restrict__ hi,
                                                            float* __restrict__ lo, int64_t n) {
  const bool rna = hi != nullptr;
  const int64_t q = n / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < q; i += (int64_t)gridDim.x * blockDim.x) {
    float4 d = ldg4(dy + 4 * i), o = ldg4(y + 4 * i);
    float4 v = make_float4(d.x * act_bwd_from_out(o.x, act, slope), d.y * act_bwd_from_out(o.y, act, slope),
                           d.z * act_bwd_from_out(o.z, act, slope), d.w * act_bwd_from_out(o.w, act, slope));
    float4 h = make_float4(tf32_part(v.x, rna), tf32_part(v.y, rna), tf32_part(v.z, rna), tf32_part(v.w, rna));
    stg4(g + 4 * i, v);
    if (hi) stg4(hi + 4 * i, h);
    stg4(lo + 4 * i, make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w));
  }
  for (int64_t i = q * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = dy[i] * act_bwd_from_out(y[i], act, slope);
    float h = tf32_part(v, rna);
    g[i] = v;
    if (hi) hi[i] = h;
    lo[i] = v - h;
  }
}

template <int V>
__global__ void __launch_bounds__(kRedThreads) channel_sum_kernel(const float* __restrict__ x, int64_t M, int C,
                                                                  double* __restrict__ out) {
  channel_reduce<V, 1>(0, M, C, out, [&](int64_t r, int unit, double (*acc)[V]) {
    Vec<V> t = Vec<V>::load(x + r * C + unit * V);
#pragma unroll
    for (int j = 0; j < V; ++j) acc[0][j] += t.v[j];
  });
}
__global__ void cast_d2f_kernel(const double* in, float* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}

// OIHW -> forward pack [Cout][R][S][CinK] (+lo) and data-gradient pack [CinRows][R][S][CoutK] with flipped taps (+lo);
// CinK >= Cin and CoutK >= Cout are the K-side channel counts of the consumer (padded with zeros), CinRows >= Cin the
// physical channel count of the activation the data gradient is taken with respect to
__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int R, int S,
                                                          int CinRows, int CinK, int CoutK, int round_hi, float* fwd_hi,
                                                          float* fwd_lo, float* bwd_hi, float* bwd_lo, int CoutRows) {
  const int64_t total_f = (int64_t)CoutRows * R * S * CinK;      // CoutRows >= Cout: zero rows for physically padded outputs
  const int64_t total_b = (int64_t)CinRows * R * S * CoutK;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_f + total_b; i += (int64_t)gridDim.x * blockDim.x) {
    const bool fwd = i < total_f;
    int64_t t = fwd ? i : i - total_f;
    int co, ci, r, s;
    if (fwd) {
      ci = (int)(t % CinK); t /= CinK;
      s = (int)(t % S); t /= S;
      r = (int)(t % R);
      co = (int)(t / R);
    } else {
      co = (int)(t % CoutK); t /= CoutK;
      s = S - 1 - (int)(t % S); t /= S;
      r = R - 1 - (int)(t % R);
      ci = (int)(t / R);
    }
    const float v = (ci < Cin && co < Cout) ? __ldg(w + (((int64_t)co * Cin + ci) * R + r) * S + s) : 0.f;
    const float hi = tf32_hi(v);
    const float h = round_hi ? hi : v;          // SIMT consumers want the full fp32 value, tensor-core consumers tf32(v)
    const int64_t o = fwd ? i : i - total_f;
    float* dh = fwd ? fwd_hi : bwd_hi;
    float* dl = fwd ? fwd_lo : bwd_lo;
    if (dh) dh[o] = h;
    if (dl) dl[o] = v - hi;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// channel concat of maps and broadcast vectors, zero-padded to a tensor-core-legal channel count, with optional 16-bit planes
// ---------------------------------------------------------------------------------------------------------------
struct ConcatParts {
  const float* src[PVG_CONCAT_MAX_PARTS];
  int64_t bstride[PVG_CONCAT_MAX_PARTS];      // elements between consecutive samples of the part
  int off[PVG_CONCAT_MAX_PARTS + 1];          // first output channel of each part; off[nparts] = total real channels
  int is_vec[PVG_CONCAT_MAX_PARTS];
  int nparts;
};
__global__ void __launch_bounds__(256) concat_pad_kernel(const ConcatParts cp, int64_t rows, int HW, int Cpad, float* __restrict__ y,
                                                         const PlaneOut po) {
  const int U = Cpad / 4;
  const int64_t total = rows * U;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / U;
    const int c0 = (int)(i % U) * 4;
    const int64_t n = r / HW, pix = r - n * HW;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + j;
      float val = 0.f;
      for (int k = 0; k < cp.nparts; ++k) {
        if (c >= cp.off[k] && c < cp.off[k + 1]) {
          const int ck = cp.off[k + 1] - cp.off[k];
          const float* sp = cp.src[k] + n * cp.bstride[k] + (cp.is_vec[k] ? 0 : pix * ck) + (c - cp.off[k]);
          val = __ldg(sp);
          break;
        }
      }
      v[j] = val;
    }
    const float4 o = make_float4(v[0], v[1], v[2], v[3]);
    stg4(y + r * Cpad + c0, o);
    emit_planes4(po, r * Cpad + c0, o);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// evaluator-side reductions and the frame conversion of the evaluation-dataset builder
// ---------------------------------------------------------------------------------------------------------------
// out[n] = mean over (c, p) of (a - b)^2 [* motion mask], n = b * T + t.  Element (c, p) of sample n lives at n * C * P + c * sc + p * sp
// (channels-last: sc = 1, sp = C; planar: sc = P, sp = 1).  With use_mask the weight of pixel p is sum_c |a_t - a_{t-1}| / C of the
// REFERENCE frames (0 for t = 0): evaluation/metrics/motion_mask.py:13-34, motion_masked_mse.py:23-26.
__global__ void __launch_bounds__(256) sqdiff_mean_kernel(const float* __restrict__ a, const float* __restrict__ b, int T, int C, int64_t P,
                                                          int64_t sc, int64_t sp, int use_mask, double* __restrict__ out) {
  const int n = blockIdx.y;
  const int t = n % T;
  const float* pa = a + (int64_t)n * C * P;
  const float* pb = b + (int64_t)n * C * P;
  const float* pprev = pa - (int64_t)C * P;          // previous frame of the same sequence (only read when t > 0)
  double acc = 0.0;
  if (!(use_mask && t == 0)) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
      float d2 = 0.f, m = 0.f;
      for (int c = 0; c < C; ++c) {
        const int64_t o = c * sc + p * sp;
        const float ra = __ldg(pa + o), d = ra - __ldg(pb + o);
        d2 += d * d;
        if (use_mask) m += fabsf(ra - __ldg(pprev + o));
      }
      acc += use_mask ? (double)d2 * (double)(m / (float)C) : (double)d2;
    }
  }
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(out + n, red[0] / ((double)C * (double)P));
}

// smallest element of x as ordered-int bits (atomicMin on the transformed pattern); *min_bits initialised to INT_MAX by the caller
__device__ __forceinline__ int float_to_ordered(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__global__ void __launch_bounds__(256) min_kernel(const float* __restrict__ x, int64_t n, int* __restrict__ min_bits) {
  float m = 3.0e38f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fminf(m, __ldg(x + i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMin(min_bits, float_to_ordered(m));
}
// uint8 frames of the evaluation-dataset builder: x in [-1, 1] (some value negative: (x + 1) / 2 first) or [0, 1] -> (uint8)(x * 255),
// truncating like numpy's astype (evaluation_dataset_builder.py:66-68,142-154); element order unchanged (channels-last in, HWC out)
__global__ void __launch_bounds__(256) frames_to_u8_kernel(const float* __restrict__ x, int64_t n, const int* __restrict__ min_bits,
                                                           uint8_t* __restrict__ out) {
  const bool shift = *min_bits < 0;                    // ordered-int pattern of a negative float is negative
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = __ldg(x + i);
    if (shift) v = (v + 1.f) / 2.f;
    v = v * 255.f;
    out[i] = (uint8_t)fminf(fmaxf(v, 0.f), 255.f);
  }
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                                   float wd, float bc1, float bc2_sqrt, float grad_scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float pi = p[i];
    float gi = g[i] * grad_scale + wd * pi;
    float mi = m[i] + (gi - m[i]) * (1.f - b1);          // exp_avg.lerp_(grad, 1 - beta1)
    float vi = v[i] * b2 + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

// hyper (device): [lr / bias_correction1, sqrt(bias_correction2), beta1, beta2, eps, weight_decay, grad_scale]
__global__ void __launch_bounds__(256) adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                       float* __restrict__ v, int64_t n, const float* __restrict__ hyper) {
  const float step_size = hyper[0], bc2_sqrt = hyper[1], b1 = hyper[2], b2 = hyper[3], eps = hyper[4], wd = hyper[5], gs = hyper[6];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float pi = p[i];
    float gi = g[i] * gs + wd * pi;
    float mi = m[i] + (gi - m[i]) * (1.f - b1);
    float vi = v[i] * b2 + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] = pi - step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  }
}

}  // namespace pvg

using namespace pvg;

#define DISPATCH_V(C, ...)                     \
  if ((C) % 4 == 0) { constexpr int V = 4; __VA_ARGS__; } else { constexpr int V = 1; __VA_ARGS__; }

static int red_grid_x(int64_t rows, int lanes, int groups) {
  int64_t want = ceil_div64(rows, (int64_t)lanes * 16);
  int64_t cap = (int64_t)kSMs * 4 / (groups > 0 ? groups : 1);
  if (cap < 1) cap = 1;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

// ---------------------------------------------------------------------------------------------------------------
// input pipeline: crop + ToTensor + Normalize of uint8 RGB frames, on the device
// ---------------------------------------------------------------------------------------------------------------
// src: [N][Hs][Ws][3] uint8 (what PIL / the PNG decoder produces), dst: [N][H][W][3] fp32 = ((u8 / 255) - mean) / std of the
// crop box whose top-left corner is (left, top).  Same operation order as torchvision's ToTensor + Normalize
// (dataset/transforms.py:90-108), IEEE divisions: bit-identical to the reference's CPU transform.  15 bytes per pixel.
__global__ void __launch_bounds__(256) frames_u8_kernel(const uint8_t* __restrict__ src, int64_t total, int Hs, int Ws, int left,
                                                        int top, int H, int W, float mean, float stdv, float* __restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int64_t t = i / W;
    const int y = (int)(t % H);
    const int64_t n = t / H;
    const uint8_t* p = src + ((n * Hs + (y + top)) * Ws + (x + left)) * 3;
    float* o = dst + i * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = __fdiv_rn(__fdiv_rn((float)p[c], 255.f) - mean, stdv);
  }
}

// One pass of PIL's two-pass ``Image.resize(size, BILINEAR)`` on 8-bit RGB (dataset/transforms.py:28; Pillow's
// ImagingResampleHorizontal_8bpc / Vertical_8bpc): out = clip8((2^21 + sum_k in[first + k] * coeff[k]) >> 22) with the 22-bit
// fixed-point coefficient table and the (first, count) bounds of every output position, both computed on the host exactly as
// Pillow's precompute_coeffs + normalize_coeffs_8bpc do (ops.pil_bilinear_coeffs).  `axis_stride` = elements between
// consecutive input positions along the resampled axis; the other axis and the channels are walked through (row, col) strides.
__global__ void __launch_bounds__(256) resample_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int64_t total,
                                                          int out_axis, int other, int64_t src_img, int64_t src_axis_stride,
                                                          int64_t src_other_stride, int64_t dst_img, int64_t dst_axis_stride,
                                                          int64_t dst_other_stride, const int* __restrict__ bounds,
                                                          const int* __restrict__ kk, int ksize) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % 3);
    int64_t t = i / 3;
    const int o = (int)(t % other); t /= other;
    const int a = (int)(t % out_axis);
    const int64_t n = t / out_axis;
    const int first = __ldg(bounds + 2 * a), count = __ldg(bounds + 2 * a + 1);
    const uint8_t* p = src + n * src_img + (int64_t)first * src_axis_stride + (int64_t)o * src_other_stride + c;
    const int* k = kk + (int64_t)a * ksize;
    int ss = 1 << 21;
    for (int x = 0; x < count; ++x) ss += (int)p[(int64_t)x * src_axis_stride] * __ldg(k + x);
    ss >>= 22;
    dst[n * dst_img + (int64_t)a * dst_axis_stride + (int64_t)o * dst_other_stride + c] = (uint8_t)(ss < 0 ? 0 : (ss > 255 ? 255 : ss));
  }
}

extern "C" {

const char* pvg_last_error(void) { return pvg::g_last_error.c_str(); }
int pvg_version(void) { return 100; }

int pvg_has_umma(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
  return p.major == 10 ? 1 : 0;
}

int pvg_bn_stats(const float* x, int N, int HW, int C, int groups, double* sums, void* stream) {
  PVG_CHECK_ARG(groups >= 1 && N % groups == 0, "N must be divisible by groups");
  int U = (C % 4 == 0) ? C / 4 : C;
  PVG_CHECK_ARG(U >= 1 && U <= kRedThreads, "unsupported channel count");
  int64_t rpg = (int64_t)(N / groups) * HW;
  dim3 grid(red_grid_x(rpg, kRedThreads / U, groups), groups);
  DISPATCH_V(C, (bn_stats_kernel<V><<<grid, kRedThreads, 0, (cudaStream_t)stream>>>(x, rpg, C, sums)));
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_pool2_stats(const float* x, int N, int H, int W, int C, float* y, int groups, double* sums, void* stream) {
  PVG_CHECK_ARG(groups >= 1 && N % groups == 0, "N must be divisible by groups");
  PVG_CHECK_ARG(H % 2 == 0 && W % 2 == 0, "avg_pool2d(2) needs even H, W");
  int U = (C % 4 == 0) ? C / 4 : C;
  PVG_CHECK_ARG(U >= 1 && U <= kRedThreads, "unsupported channel count");
  int64_t rpg = (int64_t)(N / groups) * (H / 2) * (W / 2);
  dim3 grid(red_grid_x(rpg, kRedThreads / U, groups), groups);
  DISPATCH_V(C, (pool2_stats_kernel<V><<<grid, kRedThreads, 0, (cudaStream_t)stream>>>(x, H, W, C, rpg, y, sums)));
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_bn_finalize(const double* sums, int64_t count, int groups, int C, float eps, float momentum, float* running_mean,
                    float* running_var, float* mean, float* invstd, void* stream) {
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(sums, (double)count, groups, C, eps, momentum,
                                                                         running_mean, running_var, mean, invstd);
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_bn_eval_prepare(const float* running_mean, const float* running_var, int C, float eps, float* mean, float* invstd,
                        void* stream) {
  bn_eval_prepare_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(running_mean, running_var, C, eps, mean, invstd);
  PVG_LAUNCH_OK();
  return 0;
}

static int planes_ok(const void* a, const void* b, int C) {
  if ((a || b) && C % 8 != 0) { pvg::set_error("16-bit plane outputs need C % 8 == 0"); return 0; }
  return 1;
}

int pvg_bn_apply_ex(const float* x, int N, int HW, int C, int groups, const float* mean, const float* invstd,
                    const float* weight, const float* bias, const float* residual, int act, float slope, float* y,
                    void* planes_a, int fmt_a, void* planes_b, int fmt_b, int Cparams, void* stream) {
  const int Cp = (Cparams > 0 && Cparams < C) ? Cparams : C;
  PVG_CHECK_ARG(groups >= 1 && N % groups == 0, "N must be divisible by groups");
  PVG_CHECK_ARG(((C % 4 == 0) ? C / 4 : C) <= 256, "unsupported channel count");
  if (!planes_ok(planes_a, planes_b, C)) return -1;
  int64_t M = (int64_t)N * HW, rpg = (int64_t)(N / groups) * HW;
  const PlaneOut po = make_planes(planes_a, fmt_a, planes_b, fmt_b, M * C);
  DISPATCH_V(C, (bn_apply_kernel<V><<<ew_grid((M * (C / V) + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(
                    x, M, rpg, C, mean, invstd, weight, bias, residual, act, slope, y, po, Cp)));
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_bn_apply(const float* x, int N, int HW, int C, int groups, const float* mean, const float* invstd,
                 const float* weight, const float* bias, const float* residual, int act, float slope, float* y,
                 void* stream) {
  return pvg_bn_apply_ex(x, N, HW, C, groups, mean, invstd, weight, bias, residual, act, slope, y, nullptr, 0, nullptr, 0, 0, stream);
}

int pvg_bn_finalize_apply_ex(const float* x, int N, int HW, int C, int groups, const double* sums, int64_t count, float eps,
                             float momentum, float* running_mean, float* running_var, float* mean, float* invstd,
                             const float* weight, const float* bias, const float* residual, int act, float slope, float* y,
                             void* planes_a, int fmt_a, void* planes_b, int fmt_b, int Cparams, void* stream) {
  PVG_CHECK_ARG(groups >= 1 && N % groups == 0, "N must be divisible by groups");
  PVG_CHECK_ARG(((C % 4 == 0) ? C / 4 : C) <= 256, "unsupported channel count");
  if (!planes_ok(planes_a, planes_b, C)) return -1;
  const int Cp = (Cparams > 0 && Cparams < C) ? Cparams : C;
  const size_t smem = (size_t)groups * C * 2 * sizeof(float);
  PVG_CHECK_ARG(smem <= 48 * 1024, "groups * C too large for the fused kernel: call pvg_bn_finalize + pvg_bn_apply");
  int64_t M = (int64_t)N * HW, rpg = (int64_t)(N / groups) * HW;
  const PlaneOut po = make_planes(planes_a, fmt_a, planes_b, fmt_b, M * C);
  DISPATCH_V(C, (bn_finalize_apply_kernel<V><<<ew_grid((M * (C / V) + 3) / 4, 256), 256, smem, (cudaStream_t)stream>>>(
                    x, M, rpg, C, groups, sums, (double)count, eps, momentum, running_mean, running_var, mean, invstd, weight,
                    bias, residual, act, slope, y, po, Cp)));
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_bn_finalize_apply(const float* x, int N, int HW, int C, int groups, const double* sums, int64_t count, float eps,
                          float momentum, float* running_mean, float* running_var, float* mean, float* invstd,
                          const float* weight, const float* bias, const float* residual, int act, float slope, float* y,
                          void* stream) {
  return pvg_bn_finalize_apply_ex(x, N, HW, C, groups, sums, count, eps, momentum, running_mean, running_var, mean, invstd, weight,
                                  bias, residual, act, slope, y, nullptr, 0, nullptr, 0, 0, stream);
}

int pvg_bn_bwd_reduce(const float* dy, const float* y, const float* x, int N, int HW, int C, int groups, const float* mean,
                      const float* invstd, int act, float slope, double* sums2, void* stream) {
  PVG_CHECK_ARG(groups >= 1 && N % groups == 0, "N must be divisible by groups");
  int U = (C % 4 == 0) ? C / 4 : C;
  PVG_CHECK_ARG(U >= 1 && U <= kRedThreads, "unsupported channel count");
  int64_t rpg = (int64_t)(N / groups) * HW;
  dim3 grid(red_grid_x(rpg, kRedThreads / U, groups), groups);
  DISPATCH_V(C, (bn_bwd_reduce_kernel<V><<<grid, kRedThreads, 0, (cudaStream_t)stream>>>(dy, y, x, rpg, C, mean, invstd, act,
                                                                                         slope, sums2)));
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_bn_bwd_apply(const float* dy, const float* y, const float* x, int N, int H, int W, int C, int groups,
                     const float* mean, const float* invstd, const float* weight, int act, float slope, const double* sums2,
                     int eval, int unpool, float* dx, float* g_out, float* dweight, float* dbias, void* stream) {
  return pvg_bn_bwd_apply_ex(dy, y, x, N, H, W, C, groups, mean, invstd, weight, act, slope, sums2, eval, unpool, dx, g_out, dweight,
                             dbias, 0, nullptr, stream);
}

int pvg_bn_bwd_apply_ex(const float* dy, const float* y, const float* x, int N, int H, int W, int C, int groups,
                        const float* mean, const float* invstd, const float* weight, int act, float slope, const double* sums2,
                        int eval, int unpool, float* dx, float* g_out, float* dweight, float* dbias, int Cparams, uint32_t* amax_out,
                        void* stream) {
  PVG_CHECK_ARG(groups >= 1 && N % groups == 0, "N must be divisible by groups");
  PVG_CHECK_ARG(((C % 4 == 0) ? C / 4 : C) <= 256, "unsupported channel count");
  const int Cp = (Cparams > 0 && Cparams < C) ? Cparams : C;
  int OH = unpool ? H / 2 : H, OW = unpool ? W / 2 : W;
  int64_t M = (int64_t)N * OH * OW, rpg = (int64_t)(N / groups) * OH * OW;
  DISPATCH_V(C, (bn_bwd_apply_kernel<V><<<ew_grid((M * (C / V) + 1) / 2, 256), 256, 0, (cudaStream_t)stream>>>(
                    dy, y, x, M, rpg, OH, OW, C, mean, invstd, weight, act, slope, sums2, eval, unpool, dx, g_out, groups, dweight,
                    dbias, Cp, amax_out)));
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_bn_bwd_params(const double* sums2, int groups, int C, float* dweight, float* dbias, void* stream) {
  bn_bwd_params_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(sums2, groups, C, dweight, dbias);
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_upsample2x_fwd(const float* x, int N, int H, int W, int C, float* y, void* stream) {
  return pvg_resize_bilinear(x, N, H, W, C, y, 2 * H, 2 * W, stream);
}

int pvg_resize_bilinear_ex(const float* x, int N, int H, int W, int C, float* y, int OH, int OW, void* planes_a, int fmt_a,
                           void* planes_b, int fmt_b, void* stream) {
  if (!planes_ok(planes_a, planes_b, C)) return -1;
  int64_t total = (int64_t)N * OH * OW * C;
  float sh = (float)H / (float)OH, sw = (float)W / (float)OW;
  const PlaneOut po = make_planes(planes_a, fmt_a, planes_b, fmt_b, total);
  DISPATCH_V(C, (resize_bilinear_kernel<V><<<ew_grid(total / V, 256), 256, 0, (cudaStream_t)stream>>>(x, N, H, W, C, y, OH,
                                                                                                    OW, sh, sw, po)));
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_upsample2x_bwd(const float* dy, int N, int H, int W, int C, float* dx, void* stream) {
  int64_t total = (int64_t)N * H * W * C;
  DISPATCH_V(C, (upsample2x_bwd_kernel<V><<<ew_grid(total / V, 256), 256, 0, (cudaStream_t)stream>>>(dy, N, H, W, C, dx)));
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_resize_bilinear(const float* x, int N, int H, int W, int C, float* y, int OH, int OW, void* stream) {
  int64_t total = (int64_t)N * OH * OW * C;
  float sh = (float)H / (float)OH, sw = (float)W / (float)OW;
  DISPATCH_V(C, (resize_bilinear_kernel<V><<<ew_grid(total / V, 256), 256, 0, (cudaStream_t)stream>>>(x, N, H, W, C, y, OH,
                                                                                                    OW, sh, sw, no_planes())));
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_maxpool2_fwd_ex(const float* x, int N, int H, int W, int C, float* y, void* planes_a, int fmt_a, void* stream) {
  PVG_CHECK_ARG(H >= 2 && W >= 2, "max_pool2d(2) needs H, W >= 2");     // odd sizes floor, like nn.MaxPool2d
  if (!planes_ok(planes_a, nullptr, C)) return -1;
  int64_t total = (int64_t)N * (H / 2) * (W / 2) * C;
  const PlaneOut po = make_planes(planes_a, fmt_a, nullptr, 0, total);
  DISPATCH_V(C, (maxpool2_fwd_kernel<V><<<ew_grid(total / V, 256), 256, 0, (cudaStream_t)stream>>>(x, N, H, W, C, y, po)));
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_maxpool2_fwd(const float* x, int N, int H, int W, int C, float* y, void* stream) {
  return pvg_maxpool2_fwd_ex(x, N, H, W, C, y, nullptr, 0, stream);
}

int pvg_maxpool2_bwd(const float* dy, const float* x, const float* y, int N, int H, int W, int C, int relu_mask, float* dx,
                     void* stream) {
  PVG_CHECK_ARG(H >= 2 && W >= 2, "max_pool2d(2) needs H, W >= 2");     // the caller zero-fills dx when H or W is odd
  int64_t total = (int64_t)N * (H / 2) * (W / 2) * C;
  DISPATCH_V(C, (maxpool2_bwd_kernel<V><<<ew_grid(total / V, 256), 256, 0, (cudaStream_t)stream>>>(dy, x, y, N, H, W, C,
                                                                                                 relu_mask, dx)));
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_concat_pad(const pvg_concat_desc* d, float* y, void* planes_a, int fmt_a, void* planes_b, int fmt_b, void* stream) {
  PVG_CHECK_ARG(d && y && d->nparts >= 1 && d->nparts <= PVG_CONCAT_MAX_PARTS, "bad part count");
  PVG_CHECK_ARG(d->Cpad % 4 == 0 && d->N > 0 && d->H > 0 && d->W > 0, "Cpad must be a multiple of 4");
  if (!planes_ok(planes_a, planes_b, d->Cpad)) return -1;
  ConcatParts cp;
  int off = 0;
  for (int k = 0; k < d->nparts; ++k) {
    PVG_CHECK_ARG(d->src[k] && d->c[k] > 0, "empty part");
    cp.src[k] = (const float*)d->src[k]; cp.bstride[k] = d->bstride[k]; cp.is_vec[k] = d->is_vec[k]; cp.off[k] = off;
    off += d->c[k];
  }
  cp.off[d->nparts] = off; cp.nparts = d->nparts;
  PVG_CHECK_ARG(off <= d->Cpad, "parts wider than Cpad");
  const int64_t rows = (int64_t)d->N * d->H * d->W;
  const PlaneOut po = make_planes(planes_a, fmt_a, planes_b, fmt_b, rows * d->Cpad);
  concat_pad_kernel<<<ew_grid(rows * (d->Cpad / 4), 256), 256, 0, (cudaStream_t)stream>>>(cp, rows, d->H * d->W, d->Cpad, y, po);
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_lstm_fwd(const float* gates, const float* c_prev, int64_t M, int C, float* c_new, float* h_new, void* stream) {
  DISPATCH_V(C, (lstm_fwd_kernel<V><<<ew_grid(M * (C / V), 256), 256, 0, (cudaStream_t)stream>>>(gates, c_prev, M, C, c_new,
                                                                                               h_new)));
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_lstm_bwd(const float* gates, const float* c_prev, const float* c_new, const float* dh, const float* dc_new, int64_t M,
                 int C, float* dgates, float* dc_prev, void* stream) {
  DISPATCH_V(C, (lstm_bwd_kernel<V><<<ew_grid(M * (C / V), 256), 256, 0, (cudaStream_t)stream>>>(gates, c_prev, c_new, dh,
                                                                                               dc_new, M, C, dgates, dc_prev)));
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_lstm_bwd_act(const float* gates_act, const float* c_prev, const float* c_new, const float* dh, const float* dc_new, int64_t M,
                     int C, float* dgates, float* dc_prev, void* stream) {
  PVG_CHECK_ARG(gates_act && c_prev && c_new && dgates && dc_prev, "null argument");
  PVG_CHECK_ARG((((uintptr_t)gates_act | (uintptr_t)dgates) & 15) == 0, "gate tensors must be 16-byte aligned");
  const int64_t total = M * C;
  lstm_bwd_act_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(gates_act, c_prev, c_new, dh, dc_new, total, dgates, dc_prev);
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_absdiff_mean_fwd(const float* a, const float* b, int N, int64_t count, double* out, void* stream) {
  int gx = (int)ceil_div64(count, 256 * 16);
  int cap = kSMs * 8 / (N > 0 ? N : 1);
  if (cap < 1) cap = 1;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  absdiff_mean_fwd_kernel<<<dim3(gx, N), 256, 0, (cudaStream_t)stream>>>(a, b, count, out);
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_absdiff_mean_bwd(const float* a, const float* b, const float* gout, int N, int64_t count, float* db, void* stream) {
  int gx = (int)ceil_div64(count, 256 * 8);
  int cap = kSMs * 8 / (N > 0 ? N : 1);
  if (cap < 1) cap = 1;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  absdiff_mean_bwd_kernel<<<dim3(gx, N), 256, 0, (cudaStream_t)stream>>>(a, b, gout, count, db);
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream) {
  PVG_CHECK_ARG((((uintptr_t)x | (uintptr_t)lo | (uintptr_t)hi) & 15) == 0, "pointers must be 16-byte aligned");
  split_tf32_kernel<<<ew_grid(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(x, hi, lo, n);
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_split_16(const float* x, void* planes, int64_t n, int fmt, void* stream) {
  PVG_CHECK_ARG(n % 8 == 0 && (((uintptr_t)x | (uintptr_t)planes) & 15) == 0, "n % 8 == 0 and 16-byte aligned pointers required");
  PVG_CHECK_ARG(fmt >= 0 && fmt <= 2, "unknown 16-bit plane format");
  if (fmt == PVG_CORR_FP16_ALL) split_16_kernel<2><<<ew_grid(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(x, (uint16_t*)planes, n);
  else if (fmt == PVG_CORR_FP16) split_16_kernel<1><<<ew_grid(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(x, (uint16_t*)planes, n);
  else split_16_kernel<0><<<ew_grid(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(x, (uint16_t*)planes, n);
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_pack_16x2(const float* hi, const float* lo, void* planes, int64_t n, int fmt, void* stream) {
  PVG_CHECK_ARG(hi && lo && planes, "null argument");
  PVG_CHECK_ARG(fmt >= 0 && fmt <= 2, "unknown 16-bit plane format");
  if (fmt == PVG_CORR_FP16_ALL) pack_16x2_kernel<2><<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(hi, lo, (uint16_t*)planes, n);
  else if (fmt == PVG_CORR_FP16) pack_16x2_kernel<1><<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(hi, lo, (uint16_t*)planes, n);
  else pack_16x2_kernel<0><<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(hi, lo, (uint16_t*)planes, n);
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_act_bwd_split_16(const float* dy, const float* y, int act, float slope, float* g, void* planes, int64_t n, int fmt,
                         void* stream) {
  PVG_CHECK_ARG(n % 8 == 0 && (((uintptr_t)dy | (uintptr_t)y | (uintptr_t)g | (uintptr_t)planes) & 15) == 0,
                "n % 8 == 0 and 16-byte aligned pointers required");
  PVG_CHECK_ARG(fmt >= 0 && fmt <= 2, "unknown 16-bit plane format");
  if (fmt == PVG_CORR_FP16_ALL)
    act_bwd_split_16_kernel<2><<<ew_grid(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(dy, y, act, slope, g, (uint16_t*)planes, n);
  else if (fmt == PVG_CORR_FP16)
    act_bwd_split_16_kernel<1><<<ew_grid(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(dy, y, act, slope, g, (uint16_t*)planes, n);
  else
    act_bwd_split_16_kernel<0><<<ew_grid(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(dy, y, act, slope, g, (uint16_t*)planes, n);
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_amax(const float* x, int64_t n, uint32_t* amax_bits, void* stream) {
  PVG_CHECK_ARG(x && amax_bits && n > 0 && (((uintptr_t)x) & 15) == 0, "null / misaligned argument");
  amax_kernel<<<ew_grid(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(x, n, amax_bits);
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_split_16_scaled(const float* x, void* planes, int64_t n, const uint32_t* amax_bits, float* inv_scale, void* stream) {
  PVG_CHECK_ARG(n % 8 == 0 && (((uintptr_t)x | (uintptr_t)planes) & 15) == 0 && amax_bits && inv_scale,
                "n % 8 == 0, 16-byte aligned pointers and the amax / scale scalars are required");
  split_16_scaled_kernel<<<ew_grid(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(x, (uint16_t*)planes, n, amax_bits, inv_scale);
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_act_bwd_split_16_scaled(const float* dy, const float* y, int act, float slope, float* g, void* planes, int64_t n,
                                const uint32_t* amax_bits, float* inv_scale, void* stream) {
  PVG_CHECK_ARG(n % 8 == 0 && (((uintptr_t)dy | (uintptr_t)y | (uintptr_t)g | (uintptr_t)planes) & 15) == 0 && amax_bits && inv_scale,
                "n % 8 == 0, 16-byte aligned pointers and the amax / scale scalars are required");
  act_bwd_split_16_scaled_kernel<<<ew_grid(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(dy, y, act, slope, g, (uint16_t*)planes, n,
                                                                                        amax_bits, inv_scale);
  PVG_LAUNCH_OK();
  return 0;
}

static dim3 tap_grid(int64_t count, int N) {
  int64_t gx = (count / 4 + 255) / 256;
  const int64_t cap = (148 * 16 + N - 1) / N;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  return dim3((unsigned)gx, (unsigned)N);
}

int pvg_act_bwd_tap(const float* dy, const float* y, int act, float slope, float* g, int N, int64_t count, const float* target,
                    const float* tap_gout, void* stream) {
  PVG_CHECK_ARG(dy && y && g && target && tap_gout && N > 0 && N <= 65535 && count > 0 && count % 4 == 0 &&
                    (((uintptr_t)dy | (uintptr_t)y | (uintptr_t)g | (uintptr_t)target) & 15) == 0,
                "count % 4 == 0, N <= 65535 and 16-byte aligned pointers required");
  act_bwd_tap_kernel<<<tap_grid(count, N), 256, 0, (cudaStream_t)stream>>>(dy, y, act, slope, g, count, target, tap_gout);
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_act_bwd_tap_split_16_scaled(const float* dy, const float* y, int act, float slope, float* g, void* planes, int N,
                                    int64_t count, const uint32_t* amax_bits, float* inv_scale, const float* target,
                                    const float* tap_gout, void* stream) {
  PVG_CHECK_ARG(dy && y && planes && target && tap_gout && amax_bits && inv_scale && N > 0 && N <= 65535 && count > 0 &&
                    count % 8 == 0 && (((uintptr_t)dy | (uintptr_t)y | (uintptr_t)g | (uintptr_t)planes | (uintptr_t)target) & 15) == 0,
                "count % 8 == 0, N <= 65535, 16-byte aligned pointers and the amax / scale scalars are required");
  act_bwd_tap_split_16_scaled_kernel<<<tap_grid(count, N), 256, 0, (cudaStream_t)stream>>>(
      dy, y, act, slope, g, (uint16_t*)planes, count, N, amax_bits, inv_scale, target, tap_gout);
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_act_bwd(const float* dy, const float* y, int act, float slope, float* g, int64_t n, void* stream) {
  act_bwd_kernel<<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(dy, y, act, slope, g, n);
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_act_bwd_split(const float* dy, const float* y, int act, float slope, float* g, float* hi, float* lo, int64_t n,
                      void* stream) {
  PVG_CHECK_ARG((((uintptr_t)dy | (uintptr_t)y | (uintptr_t)g | (uintptr_t)lo | (uintptr_t)hi) & 15) == 0, "pointers must be 16-byte aligned");
  act_bwd_split_kernel<<<ew_grid(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(dy, y, act, slope, g, hi, lo, n);
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_channel_sum(const float* x, int64_t M, int C, double* scratch, float* out, void* stream) {
  int U = (C % 4 == 0) ? C / 4 : C;
  PVG_CHECK_ARG(U >= 1 && U <= kRedThreads, "unsupported channel count");
  PVG_CUDA_OK(cudaMemsetAsync(scratch, 0, sizeof(double) * C, (cudaStream_t)stream));
  int gx = red_grid_x(M, kRedThreads / U, 1);
  DISPATCH_V(C, (channel_sum_kernel<V><<<gx, kRedThreads, 0, (cudaStream_t)stream>>>(x, M, C, scratch)));
  PVG_LAUNCH_OK();
  cast_d2f_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(scratch, out, C);
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_pack_conv_weight(const float* w_oihw, int Cout, int Cin, int R, int S, int CinRows, int CinK, int CoutK, int round_hi,
                         float* fwd_hi, float* fwd_lo, float* bwd_hi, float* bwd_lo, void* stream) {
  return pvg_pack_conv_weight_ex(w_oihw, Cout, Cin, R, S, CinRows, CinK, CoutK, Cout, round_hi, fwd_hi, fwd_lo, bwd_hi, bwd_lo, stream);
}

int pvg_pack_conv_weight_ex(const float* w_oihw, int Cout, int Cin, int R, int S, int CinRows, int CinK, int CoutK, int CoutRows,
                            int round_hi, float* fwd_hi, float* fwd_lo, float* bwd_hi, float* bwd_lo, void* stream) {
  PVG_CHECK_ARG(CinRows >= Cin && CinK >= CinRows && CoutK >= Cout && CoutRows >= Cout, "padded channel counts smaller than the real ones");
  int64_t total = (int64_t)CoutRows * R * S * CinK + (int64_t)CinRows * R * S * CoutK;
  pack_weight_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(w_oihw, Cout, Cin, R, S, CinRows, CinK, CoutK,
                                                                          round_hi, fwd_hi, fwd_lo, bwd_hi, bwd_lo, CoutRows);
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_frames_u8_to_nhwc(const uint8_t* src, int N, int Hs, int Ws, int left, int top, int H, int W, float mean, float stdv,
                          float* dst, void* stream) {
  PVG_CHECK_ARG(src && dst && N > 0 && H > 0 && W > 0, "empty problem");
  PVG_CHECK_ARG(left >= 0 && top >= 0 && left + W <= Ws && top + H <= Hs, "crop box outside the source frame");
  PVG_CHECK_ARG(stdv != 0.f, "std must not be zero");
  const int64_t total = (int64_t)N * H * W;
  frames_u8_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(src, total, Hs, Ws, left, top, H, W, mean, stdv, dst);
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_resample_u8(const uint8_t* src, int N, int Hs, int Ws, int left, int top, int Hin, int Win, int vertical, int out_size,
                    const int* bounds, const int* kk, int ksize, uint8_t* dst, void* stream) {
  PVG_CHECK_ARG(src && dst && bounds && kk && N > 0 && Hin > 0 && Win > 0 && out_size > 0 && ksize > 0, "bad argument");
  PVG_CHECK_ARG(left >= 0 && top >= 0 && left + Win <= Ws && top + Hin <= Hs, "input box outside the source frame");
  const uint8_t* s0 = src + ((int64_t)top * Ws + left) * 3;
  const int64_t src_img = (int64_t)Hs * Ws * 3;
  if (!vertical) {                 // [N][Hin][Win][3] -> [N][Hin][out_size][3]
    const int64_t total = (int64_t)N * Hin * out_size * 3;
    resample_u8_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(s0, dst, total, out_size, Hin, src_img, 3, (int64_t)Ws * 3,
                                                                               (int64_t)Hin * out_size * 3, 3, (int64_t)out_size * 3,
                                                                               bounds, kk, ksize);
  } else {                         // [N][Hin][Win][3] -> [N][out_size][Win][3]
    const int64_t total = (int64_t)N * out_size * Win * 3;
    resample_u8_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(s0, dst, total, out_size, Win, src_img, (int64_t)Ws * 3, 3,
                                                                               (int64_t)out_size * Win * 3, (int64_t)Win * 3, 3, bounds,
                                                                               kk, ksize);
  }
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_sqdiff_mean(const float* a, const float* b, int N, int T, int C, int64_t P, int channels_last, int use_mask, double* out,
                    void* stream) {
  PVG_CHECK_ARG(a && b && out && N > 0 && T > 0 && N % T == 0 && C > 0 && P > 0, "bad argument");
  int gx = (int)ceil_div64(P, 256 * 8);
  int cap = kSMs * 8 / N;
  if (cap < 1) cap = 1;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  sqdiff_mean_kernel<<<dim3(gx, N), 256, 0, (cudaStream_t)stream>>>(a, b, T, C, P, channels_last ? 1 : P, channels_last ? C : 1, use_mask, out);
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_frames_to_u8(const float* x, int64_t n, int* min_scratch, uint8_t* out, void* stream) {
  PVG_CHECK_ARG(x && out && min_scratch && n > 0, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int init = 0x7fffffff;
  PVG_CUDA_OK(cudaMemcpyAsync(min_scratch, &init, sizeof(int), cudaMemcpyHostToDevice, st));
  min_kernel<<<ew_grid(n, 256), 256, 0, st>>>(x, n, min_scratch);
  PVG_LAUNCH_OK();
  frames_to_u8_kernel<<<ew_grid(n, 256), 256, 0, st>>>(x, n, min_scratch, out);
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                  float weight_decay, int step, float grad_scale, void* stream) {
  float bc1 = 1.f - powf(beta1, (float)step);
  float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  adam_kernel<<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1,
                                                                bc2_sqrt, grad_scale);
  PVG_LAUNCH_OK();
  return 0;
}

int pvg_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, void* stream) {
  adam_dev_kernel<<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, hyper);
  PVG_LAUNCH_OK();
  return 0;
}

}  // extern "C"
