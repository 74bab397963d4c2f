// Weight gradient of the 'same' convolution on the tcgen05 tensor cores (sm_100a).
//
//   dW[co][r][s][ci] = sum over pixels p of dY[p][co] * X[p + (r,s) - pad][ci]
//
//   GEMM view   D[M = 128 rows of (tap, ci)][N = BN co] += A[M][K] * B[N][K] with K = pixels.  An M tile is FOUR groups of 32
//               input channels, each group with its own tap (r,s) - so M is always full whatever Cout is (the roles are
//               swapped w.r.t. the textbook dW = dY^T X: layers with 3..64 output channels would waste the 128-row MMA).
//               Both operands are "MN-major": in NHWC memory
//               the channel index is contiguous and the reduction index (pixel) strides by C.  A 4-D TMA box
//               {32 ch, tw, th, tn} (32 pixels) lands as 32 rows (pixels) x 128 B (32 channels) with the 32-byte-atom
//               128B swizzle (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), which is exactly the canonical MN-major
//               SWIZZLE_128B_BASE32B UMMA layout - the only one the tensor core takes for MN-major tf32 (4-pixel atoms of
//               512 B along K, SBO = 512; the next 32 channels are the next box, LBO = 4096 B).  The (r,s) shift and the zero padding of X are, as in the
//               forward kernel, TMA coordinates + out-of-bounds zero fill.
//   work split  CTA = (4 (tap, 32-ci) groups) x (BN-co tile) x (slice of the pixel patches); partial tiles are combined
//               with fp32 atomics into a packed [Cout][R*S][CinP] buffer (zeroed by the caller), unpacked to OIHW after.
//   accuracy    K (pixels) reaches 5e5: the accumulation chain in TMEM is cut every kDrain stages (32 MMAs) and drained
//               into fp32 registers exactly as in conv_umma.cu; NPROD=3 adds the A_lo*B_hi + A_hi*B_lo corrections,
//               NPROD=2 evaluates those two corrections as bf16 MMAs (kind::f16, K = 16) on the bf16 plane pairs
//               {f16(lo * 2^12), f16(x)} of both operands: MN-major 64-byte-swizzled {32 ch x 32 px} tiles, both planes of a
//               channel group in one 5-D TMA box.
#include "umma.cuh"

namespace pvg {

constexpr int kWThreads = 192;
constexpr int kPix = 32;                    // pixels (K) per pipeline stage
constexpr int kBoxBytes = kPix * 128;       // one {32 ch x 32 px} box
constexpr int kWSmemBudget = 200 * 1024;

template <int BN, int NPROD>
struct WCfg {
  // NPROD == 4: ALL three products as kind::f16 MMAs on fp16 plane pairs (x: PVG_CORR_FP16_ALL planes, the forward operand;
  // dY: the same format after a power-of-two scaling, pvg_split_16_scaled): no fp32 tiles - a {32 ch x 32 px} box of both
  // planes has the bytes of one fp32 box, so half of the shared-memory / L2 traffic of the TF32 + corrections evaluation
  static constexpr int kPlanes = (NPROD == 2 || NPROD == 3) ? 2 : 1;
  static constexpr int kABytes = 4 * kBoxBytes;                 // 4 groups of 32 input channels
  static constexpr int kBBytes = (BN / 32) * kBoxBytes;
  static constexpr int kStageBytes = kPlanes * (kABytes + kBBytes);
  static constexpr int kStagesRaw = kWSmemBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kDrain = 8;
  static constexpr int kAccCols = (NPROD >= 2 ? 3 : 2) * BN;    // [main0 | main1 | (correction)]
  static constexpr int kTmemCols = kAccCols <= 64 ? 64 : (kAccCols <= 128 ? 128 : (kAccCols <= 256 ? 256 : 512));
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 128, "BN must be 32..128 in steps of 32");
  static_assert(kStages >= 2, "need at least a double buffer");
};

struct WgradParams {
  int N, H, W, CinP, Cout, R, S, pad;
  int tw, th, tn;                 // 32-pixel patch
  int tiles_w, tiles_h, tiles_n;  // patches per dim
  int chunks, groups;             // 32-channel chunks of CinP; (tap, chunk) groups = R*S*chunks
  int patches_per_split;
  float* dwp;                     // [Cout][R*S][CinP]
  int corr_fp16;                  // NPROD == 2: 16-bit correction planes are fp16 (else bf16)
  const float* out_scale;         // NPROD == 4 (device, optional): 1 / S of the scaled dY planes
};

template <int BN, int NPROD>
__global__ void __launch_bounds__(kWThreads, 1)
conv_wgrad_umma_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmGlo,
                       const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmXlo,
                       const WgradParams p) {
  using C = WCfg<BN, NPROD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + C::kStages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tfull_bar = empty_bar + C::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int g0 = blockIdx.x * 4;            // first (tap, chunk) group of this M tile
  const int co0 = blockIdx.y * BN;
  const int total_patches = p.tiles_w * p.tiles_h * p.tiles_n;
  const int p_begin = blockIdx.z * p.patches_per_split;
  const int p_end = min(p_begin + p.patches_per_split, total_patches);
  const int iters = p_end - p_begin;          // >= 1 by construction

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmG); prefetch_tmap(&tmX);
    if (NPROD >= 2) { prefetch_tmap(&tmGlo); prefetch_tmap(&tmXlo); }
    for (int i = 0; i < C::kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&tfull_bar[0], 1); mbar_init(&tfull_bar[1], 1);
    mbar_init(&tempty_bar[0], 128); mbar_init(&tempty_bar[1], 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  const int periods = (iters + C::kDrain - 1) / C::kDrain;

  if (warp == 0) {
    const uint32_t leader = elect_one();      // see umma.cuh: single-lane issue without per-instruction election loops
    // Everything that does not change from stage to stage is decoded ONCE: the (tap, chunk) of this CTA's four row groups and
    // the patch coordinates, which then advance incrementally.  (ncu, round 2: with the integer divisions inside the loop -
    // ~11 per stage, each a MUFU.RCP sequence - the producer warp needed ~2000 clk per stage and the MMA warp spent 55 % of
    // its time waiting for data: tensor pipe 5 % active on the 64->32 @256^2 layer.)
    int gc0[4], gdw[4], gdh[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int gi = g0 + g;
      const bool ok = gi < p.groups;
      const int tap = ok ? gi / p.chunks : 0, cc = ok ? gi - tap * p.chunks : 0;
      const int r = tap / p.S, sx = tap - r * p.S;
      gc0[g] = ok ? cc * 32 : p.CinP;                  // a group past the end loads an all-out-of-bounds (zero) box
      gdw[g] = sx - p.pad; gdh[g] = r - p.pad;
    }
    int t0 = p_begin;
    int pw = t0 % p.tiles_w; t0 /= p.tiles_w;
    int ph = t0 % p.tiles_h;
    int pn = t0 / p.tiles_h;
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      const int w0 = pw * p.tw, h0 = ph * p.th, n0 = pn * p.tn;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (leader) {
        uint8_t* st = smem + stage * C::kStageBytes;
        mbar_expect_tx(&full_bar[stage], C::kStageBytes);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (NPROD == 4) {
            tma_load_5d(st + g * kBoxBytes, &tmXlo, &full_bar[stage], gc0[g], w0 + gdw[g], h0 + gdh[g], n0, 0);
            continue;
          }
          tma_load_4d(st + g * kBoxBytes, &tmX, &full_bar[stage], gc0[g], w0 + gdw[g], h0 + gdh[g], n0);
          if (NPROD == 3)
            tma_load_4d(st + C::kABytes + g * kBoxBytes, &tmXlo, &full_bar[stage], gc0[g], w0 + gdw[g], h0 + gdh[g], n0);
          if (NPROD == 2)     // [f16(lo * 2^12) 32 px x 64 B | f16(x) 32 px x 64 B] of this channel group
            tma_load_5d(st + C::kABytes + g * kBoxBytes, &tmXlo, &full_bar[stage], gc0[g], w0 + gdw[g], h0 + gdh[g], n0, 0);
        }
        uint8_t* sb = st + C::kPlanes * C::kABytes;
#pragma unroll
        for (int g = 0; g < BN / 32; ++g) {
          if (NPROD == 4) {
            tma_load_5d(sb + g * kBoxBytes, &tmGlo, &full_bar[stage], co0 + 32 * g, w0, h0, n0, 0);
            continue;
          }
          tma_load_4d(sb + g * kBoxBytes, &tmG, &full_bar[stage], co0 + 32 * g, w0, h0, n0);
          if (NPROD == 3) tma_load_4d(sb + C::kBBytes + g * kBoxBytes, &tmGlo, &full_bar[stage], co0 + 32 * g, w0, h0, n0);
          if (NPROD == 2) tma_load_5d(sb + C::kBBytes + g * kBoxBytes, &tmGlo, &full_bar[stage], co0 + 32 * g, w0, h0, n0, 0);
        }
      }
      __syncwarp();
      if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      if (++pw == p.tiles_w) { pw = 0; if (++ph == p.tiles_h) { ph = 0; ++pn; } }
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one();
    constexpr uint32_t idesc = make_idesc_tf32_ex<BN>(true, true);
    const uint32_t idesc16 = make_idesc_f16<BN>(p.corr_fp16 != 0, true, true);
    // descriptors = base + offset in 16-byte units (start-address field: low 14 bits, cannot carry out)
    const uint64_t d32 = make_mnmajor_sw128_desc(smem_u32(smem), kBoxBytes);
    const uint64_t d16 = make_mnmajor_sw64_desc(smem_u32(smem), kBoxBytes);
    constexpr uint32_t kStageU = C::kStageBytes >> 4, kAU = C::kABytes >> 4, kBU = C::kBBytes >> 4;
    constexpr uint32_t kHalfU = (kBoxBytes / 2) >> 4;          // the f16(x) tile follows the f16(lo) tile of each group
    int stage = 0; uint32_t phase = 0;
    const uint32_t corr = tmem_acc + 2 * BN;
    uint32_t corr_acc = 0;
    int it = 0;
    for (int per = 0; per < periods; ++per) {
      const int b = per & 1;
      mbar_wait(&tempty_bar[b], ((per >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t main_acc = tmem_acc + b * BN;
      const int it_end = min(it + C::kDrain, iters);
      uint32_t main_started = 0;
      for (; it < it_end; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (leader) {
          const uint32_t a_hi = stage * kStageU, a_lo = a_hi + kAU;
          const uint32_t b_hi = a_hi + C::kPlanes * kAU, b_lo = b_hi + kBU;
          if (NPROD == 4) {                            // [f16(lo * 2^12) | f16(v)] tiles per channel group, for both operands
            const uint32_t pa = a_hi, pb = a_hi + kAU;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              umma_bf16(corr, d16 + (pa + 64 * ks), d16 + (pb + kHalfU + 64 * ks), idesc16, corr_acc);
              corr_acc = 1;
            }
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) umma_bf16(corr, d16 + (pa + kHalfU + 64 * ks), d16 + (pb + 64 * ks), idesc16, 1);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              umma_bf16(main_acc, d16 + (pa + kHalfU + 64 * ks), d16 + (pb + kHalfU + 64 * ks), idesc16, main_started);
              main_started = 1;
            }
          }
          if (NPROD == 2) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {           // K = 16 pixels = two 8-row (512 B) atoms: 1024 B = 64 units per step
              umma_bf16(corr, d16 + (a_lo + 64 * ks), d16 + (b_lo + kHalfU + 64 * ks), idesc16, corr_acc);
              corr_acc = 1;
            }
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) umma_bf16(corr, d16 + (a_lo + kHalfU + 64 * ks), d16 + (b_lo + 64 * ks), idesc16, 1);
          }
          if (NPROD == 3) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              umma_tf32(corr, d32 + (a_lo + 64 * ks), d32 + (b_hi + 64 * ks), idesc, corr_acc);
              corr_acc = 1;
            }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_tf32(corr, d32 + (a_hi + 64 * ks), d32 + (b_lo + 64 * ks), idesc, 1);
          }
          if (NPROD != 4) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {             // K = 8 pixels = two 4-row (512 B) atoms: 1024 B per step
            umma_tf32(main_acc, d32 + (a_hi + 64 * ks), d32 + (b_hi + 64 * ks), idesc, main_started);
            main_started = 1;
          }
          }
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
      if (leader) umma_commit(&tfull_bar[b]);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;                 // TMEM lane quarter == group index inside the M tile
    const int gi = g0 + q;
    const bool ok = gi < p.groups;
    const int tap = ok ? gi / p.chunks : 0, cc = ok ? gi - tap * p.chunks : 0;
    const int ci = cc * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float acc[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) acc[j] = 0.f;
    for (int per = 0; per < periods; ++per) {
      const int b = per & 1;
      mbar_wait(&tfull_bar[b], (per >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < BN; c += 16) {
        float v[16];
        tmem_ld16(tmem_acc + lane_base + (uint32_t)(b * BN + c), v);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[c + j] += v[j];
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[b]);
    }
    // dwp[co][tap][ci]: for a fixed co the 32 lanes of a warp hit 32 consecutive ci -> one coalesced 128-byte reduction.
    // Bounds are tested once per 16 channels and the address advances by a constant stride (the per-element form cost ~8
    // instructions per value: with split-K CTAs of a few microseconds the epilogue was a third of the kernel).
    // NPROD == 2: the residual planes carry a 2^12 factor (pvg_split_16), so does the correction accumulator
    constexpr float kCorrScale = (NPROD == 2 || NPROD == 4) ? 0x1p-12f : 1.f;
    const float out_scale = (NPROD == 4 && p.out_scale != nullptr) ? __ldg(p.out_scale) : 1.f;
    const int64_t co_stride = (int64_t)p.R * p.S * p.CinP;
    float* out = p.dwp + (int64_t)tap * p.CinP + ci + (int64_t)co0 * co_stride;
#pragma unroll
    for (int c = 0; c < BN; c += 16) {
      float v[16];
      if (NPROD >= 2) {
        tmem_ld16(tmem_acc + lane_base + (uint32_t)(2 * BN + c), v);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaf(v[j], kCorrScale, acc[c + j]) * out_scale;
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = acc[c + j];
      }
      if (ok && co0 + c < p.Cout) {
        if (co0 + c + 16 <= p.Cout) {
#pragma unroll
          for (int j = 0; j < 16; ++j) atomicAdd(out + (c + j) * co_stride, v[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (co0 + c + j < p.Cout) atomicAdd(out + (c + j) * co_stride, v[j]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_acc, C::kTmemCols);
}

__global__ void __launch_bounds__(256) unpack_dw_kernel(const float* __restrict__ dwp, int Cout, int Cin, int R, int S, int CinP,
                                                        float* __restrict__ dw, int accumulate) {
  const int64_t total = (int64_t)Cout * Cin * R * S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int s = (int)(i % S);
    int64_t t = i / S;
    int r = (int)(t % R); t /= R;
    int ci = (int)(t % Cin);
    int co = (int)(t / Cin);
    const float v = dwp[((int64_t)co * R * S + r * S + s) * CinP + ci];
    dw[i] = accumulate ? dw[i] + v : v;
  }
}

static void choose_patch32(int N, int H, int W, int* tw, int* th, int* tn) {
  static const int cand[][3] = {{8, 4, 1}, {4, 8, 1}, {16, 2, 1}, {2, 16, 1}, {32, 1, 1}, {1, 32, 1}, {4, 4, 2}, {8, 2, 2},
                                {2, 8, 2}, {4, 2, 4}, {2, 4, 4}, {2, 2, 8}, {1, 1, 32}, {16, 1, 2}, {8, 1, 4}, {4, 1, 8}};
  double best = -1;
  for (auto& c : cand) {
    int64_t tiles = (int64_t)ceil_div(W, c[0]) * ceil_div(H, c[1]) * ceil_div(N, c[2]);
    double util = (double)N * H * W / ((double)tiles * 32.0);
    if (util > best + 1e-9) { best = util; *tw = c[0]; *th = c[1]; *tn = c[2]; }
  }
}

template <int BN, int NPROD>
static int launch_wgrad(const pvg_conv_desc* d, const float* x, const float* x_lo, const float* g, const float* g_lo,
                        float* dwp, cudaStream_t st, const float* out_scale = nullptr) {
  using C = WCfg<BN, NPROD>;
  WgradParams p;
  p.out_scale = out_scale;
  const int CinK = (d->Cin + 31) & ~31;     // K-side channel count padded to whole 32-channel groups (TMA zero-fills the rest)
  p.N = d->N; p.H = d->H; p.W = d->W; p.CinP = CinK; p.Cout = d->Cout; p.R = d->R; p.S = d->S; p.pad = d->pad; p.dwp = dwp; p.corr_fp16 = d->corr_fmt != PVG_CORR_BF16;
  choose_patch32(d->N, d->H, d->W, &p.tw, &p.th, &p.tn);
  p.tiles_w = ceil_div(d->W, p.tw); p.tiles_h = ceil_div(d->H, p.th); p.tiles_n = ceil_div(d->N, p.tn);
  p.chunks = CinK / 32;
  p.groups = d->R * d->S * p.chunks;
  const int total = p.tiles_w * p.tiles_h * p.tiles_n;
  const int gx = ceil_div(p.groups, 4), gy = ceil_div(d->Cout, BN);
  int want = ceil_div(kSMs * 2, gx * gy);
  int max_splits = ceil_div(total, 8);                 // at least 8 stages of work per CTA
  int splits = want < max_splits ? want : max_splits;
  if (splits < 1) splits = 1;
  p.patches_per_split = ceil_div(total, splits);
  splits = ceil_div(total, p.patches_per_split);
  CUtensorMap tmG, tmGlo, tmX, tmXlo;
  int rc;
  if (NPROD != 4) {
    if ((rc = encode_nhwc_map(&tmG, g, d->N, d->H, d->W, d->Cout, 32, p.tw, p.th, p.tn, true))) return rc;
    if ((rc = encode_nhwc_map(&tmX, x, d->N, d->H, d->W, d->Cin, 32, p.tw, p.th, p.tn, true))) return rc;
  }
  if (NPROD == 4) {
    if ((rc = encode_nhwc_16x2_map(&tmGlo, g_lo, d->N, d->H, d->W, d->Cout, p.tw, p.th, p.tn))) return rc;
    if ((rc = encode_nhwc_16x2_map(&tmXlo, x_lo, d->N, d->H, d->W, d->Cin, p.tw, p.th, p.tn))) return rc;
    tmG = tmGlo; tmX = tmXlo;
  } else if (NPROD == 3) {
    if ((rc = encode_nhwc_map(&tmGlo, g_lo, d->N, d->H, d->W, d->Cout, 32, p.tw, p.th, p.tn, true))) return rc;
    if ((rc = encode_nhwc_map(&tmXlo, x_lo, d->N, d->H, d->W, d->Cin, 32, p.tw, p.th, p.tn, true))) return rc;
  } else if (NPROD == 2) {
    if ((rc = encode_nhwc_16x2_map(&tmGlo, g_lo, d->N, d->H, d->W, d->Cout, p.tw, p.th, p.tn))) return rc;
    if ((rc = encode_nhwc_16x2_map(&tmXlo, x_lo, d->N, d->H, d->W, d->Cin, p.tw, p.th, p.tn))) return rc;
  } else {
    tmGlo = tmG; tmXlo = tmX;
  }
  static bool attr_set = false;
  if (!attr_set) {
    PVG_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_umma_kernel<BN, NPROD>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_set = true;
  }
  dim3 grid(gx, gy, splits);
  conv_wgrad_umma_kernel<BN, NPROD><<<grid, kWThreads, C::kSmemBytes, st>>>(tmG, tmGlo, tmX, tmXlo, p);
  PVG_LAUNCH_OK();
  return 0;
}

}  // namespace pvg

using namespace pvg;

// Tensor-core weight gradient.  x: [N,H,W,CinP] (CinP % 32 == 0), g = dY: [N,H,W,Cout] (Cout % 4 == 0), *_lo their
// 3xTF32 residual planes (nprod == 3).  scratch: float[Cout * R*S * roundup(CinP, 32)], zero-initialised by the caller.
// dw_oihw [Cout][Cin_logical][R][S] = (accumulate ? dw_oihw : 0) + unpack(scratch).
extern "C" int pvg_conv2d_wgrad_umma(const pvg_conv_desc* d, int Cin_logical, const float* x, const void* x_lo_, const float* g,
                                     const void* g_lo_, float* scratch, float* dw_oihw, int accumulate, void* stream) {
  const float* x_lo = (const float*)x_lo_;      // fp32 residual planes (nprod == 3) or bf16 plane pairs (nprod == 2)
  const float* g_lo = (const float*)g_lo_;
  PVG_CHECK_ARG(d && x && g && scratch && dw_oihw, "null argument");
  PVG_CHECK_ARG(d->Cin % (d->nprod == 2 ? 8 : 4) == 0 && d->Cout % 4 == 0, "tensor-core wgrad needs Cin % 4 == 0 (% 8 with 16-bit planes) and Cout % 4 == 0");
  PVG_CHECK_ARG((((uintptr_t)x | (uintptr_t)g) & 15) == 0, "operands must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  const int cin = d->Cout;      // BN tiles the OUTPUT channels (roles swapped, see the header comment)
  if (d->nprod == 3) {
    PVG_CHECK_ARG(x_lo && g_lo, "nprod == 3 needs x_lo and g_lo");
    if (cin <= 32) rc = launch_wgrad<32, 3>(d, x, x_lo, g, g_lo, scratch, st);
    else if (cin <= 64) rc = launch_wgrad<64, 3>(d, x, x_lo, g, g_lo, scratch, st);
    else if (cin <= 96) rc = launch_wgrad<96, 3>(d, x, x_lo, g, g_lo, scratch, st);
    else rc = launch_wgrad<128, 3>(d, x, x_lo, g, g_lo, scratch, st);
  } else if (d->nprod == 2) {
    PVG_CHECK_ARG(x_lo && g_lo, "nprod == 2 needs the bf16 plane pairs of x and g");
    PVG_CHECK_ARG(d->Cout % 8 == 0 && (((uintptr_t)x_lo | (uintptr_t)g_lo) & 15) == 0, "bf16 planes need Cout % 8 == 0 and 16-byte alignment");
    if (cin <= 32) rc = launch_wgrad<32, 2>(d, x, x_lo, g, g_lo, scratch, st);
    else if (cin <= 64) rc = launch_wgrad<64, 2>(d, x, x_lo, g, g_lo, scratch, st);
    else if (cin <= 96) rc = launch_wgrad<96, 2>(d, x, x_lo, g, g_lo, scratch, st);
    else rc = launch_wgrad<128, 2>(d, x, x_lo, g, g_lo, scratch, st);
  } else {
    if (cin <= 32) rc = launch_wgrad<32, 1>(d, x, nullptr, g, nullptr, scratch, st);
    else if (cin <= 64) rc = launch_wgrad<64, 1>(d, x, nullptr, g, nullptr, scratch, st);
    else if (cin <= 96) rc = launch_wgrad<96, 1>(d, x, nullptr, g, nullptr, scratch, st);
    else rc = launch_wgrad<128, 1>(d, x, nullptr, g, nullptr, scratch, st);
  }
  if (rc) return rc;
  int64_t total = (int64_t)d->Cout * Cin_logical * d->R * d->S;
  unpack_dw_kernel<<<ew_grid(total, 256), 256, 0, st>>>(scratch, d->Cout, Cin_logical, d->R, d->S, (d->Cin + 31) & ~31, dw_oihw, accumulate);
  PVG_LAUNCH_OK();
  return 0;
}

// Weight gradient from plane pairs only (see include/pvg_b200.h): x_planes / g_planes are PVG_CORR_FP16_ALL plane pairs, the
// latter of dY * S; the result is multiplied by *out_scale = 1 / S.
extern "C" int pvg_conv2d_wgrad_planes(const pvg_conv_desc* d, int Cin_logical, const void* x_planes, const void* g_planes,
                                       const float* out_scale, float* scratch, float* dw_oihw, int accumulate, void* stream) {
  PVG_CHECK_ARG(d && x_planes && g_planes && scratch, "null argument");
  PVG_CHECK_ARG(d->Cin % 8 == 0 && d->Cout % 8 == 0, "16-bit planes need Cin % 8 == 0 and Cout % 8 == 0");
  PVG_CHECK_ARG((((uintptr_t)x_planes | (uintptr_t)g_planes) & 15) == 0, "planes must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const float* xp = (const float*)x_planes;
  const float* gp = (const float*)g_planes;
  int rc;
  const int co = d->Cout;
  if (co <= 32) rc = launch_wgrad<32, 4>(d, nullptr, xp, nullptr, gp, scratch, st, out_scale);
  else if (co <= 64) rc = launch_wgrad<64, 4>(d, nullptr, xp, nullptr, gp, scratch, st, out_scale);
  else if (co <= 96) rc = launch_wgrad<96, 4>(d, nullptr, xp, nullptr, gp, scratch, st, out_scale);
  else rc = launch_wgrad<128, 4>(d, nullptr, xp, nullptr, gp, scratch, st, out_scale);
  if (rc) return rc;
  if (dw_oihw == nullptr) return 0;       // deferred: the split-K partials of further uses keep meeting in `scratch` (pvg_unpack_dw)
  int64_t total = (int64_t)d->Cout * Cin_logical * d->R * d->S;
  unpack_dw_kernel<<<ew_grid(total, 256), 256, 0, st>>>(scratch, d->Cout, Cin_logical, d->R, d->S, (d->Cin + 31) & ~31, dw_oihw, accumulate);
  PVG_LAUNCH_OK();
  return 0;
}

// packed [Cout][R*S][roundup(CinPhys, 32)] weight-gradient scratch -> OIHW (the second half of the weight-gradient entry points,
// for callers that let several uses of one weight accumulate in the same scratch and unpack once)
extern "C" int pvg_unpack_dw(const float* scratch, int Cout, int Cin_logical, int R, int S, int CinPhys, float* dw_oihw, int accumulate,
                             void* stream) {
  PVG_CHECK_ARG(scratch && dw_oihw && Cout > 0 && Cin_logical > 0 && CinPhys >= Cin_logical, "bad argument");
  int64_t total = (int64_t)Cout * Cin_logical * R * S;
  unpack_dw_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(scratch, Cout, Cin_logical, R, S, (CinPhys + 31) & ~31, dw_oihw, accumulate);
  PVG_LAUNCH_OK();
  return 0;
}
