// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05) for sm_100a.
//
//   GEMM view      M = N*H*W output pixels (128 per CTA: a tn x th x tw patch), N = Cout (BN per CTA),
//                  K = R*S*Cin walked tap by tap in 32-channel (128-byte) chunks.
//   A operand      NHWC activations.  One 4-D TMA box {32 ch, tw, th, tn} per (tap, chunk) lands the shifted patch in
//                  shared memory as 128 rows x 128 B with the 128-byte swizzle; the conv zero padding IS the TMA
//                  out-of-bounds fill (negative / past-the-end coordinates), so there is no im2col buffer and no
//                  boundary code in the main loop.
//   B operand      weights packed [Cout][R][S][Cin] (K-major), 2-D TMA box {32, BN}, same swizzle.
//   MMA            tcgen05.mma.cta_group::1.kind::tf32, M=128, N=BN, K=8, fp32 accumulators in TMEM (BN columns),
//                  issued by one thread.  NPROD=3 runs the error-compensated split
//                  A_lo*B_hi + A_hi*B_lo + A_hi*B_hi (fp32-equivalent accuracy at 3 MMAs per k-step).
//   pipeline       warp 0: TMA producer; warp 1: MMA issuer (+TMEM alloc); warps 2-5: epilogue.  smem ring of STAGES
//                  {A[,A_lo],B[,B_lo]} slots with full/empty mbarriers; tcgen05.commit releases slots and publishes the
//                  accumulator.
//   epilogue       tcgen05.ld 32 lanes x 16 columns per warp-instruction -> +bias -> activation -> NHWC global stores
//                  (each thread owns one output pixel and writes its channels contiguously).
//   NPROD=2        same split, but the two correction terms (2^-11 of the result, so 8 mantissa bits are plenty) run as
//                  16-bit MMAs: kind::f16, K = 16, on 16-bit copies {lo * 2^12, x} of both operands (pvg_split_16 /
//                  pvg_pack_16x2), bf16 or fp16 (pvg_conv_desc.corr_fmt; fp16 keeps the tf32 mantissa of a weight exactly, so
//                  the weight side adds no error that is coherent over the batch; bf16 has the range gradients need).  Per 32-channel chunk: 4 tf32 + 4 bf16 MMAs instead of 12 tf32, and
//                  the correction operands are half as wide in shared memory - the kernel's measured limiter.
//   split accum    (NPROD>=2) The tensor core adds into its fp32 accumulator with truncation, so an accumulation chain of L
//                  MMAs drifts by ~L * 2^-24 (measured on B200: 2.3e-5 at K = 4.7k).  To stay fp32-equivalent the main
//                  term A_hi*B_hi is accumulated in TMEM for only kDrain k-iterations (32 MMAs), then drained by the
//                  epilogue warps into fp32 REGISTER accumulators (round-to-nearest adds) while the MMA warp continues
//                  into a second TMEM buffer (ping-pong, tfull/tempty mbarriers).  The two correction terms, 2^-11
//                  smaller, accumulate in a third TMEM buffer for the whole K loop and are added once at the end.
#include <stdlib.h>

#include "umma.cuh"

namespace pvg {

int conv2d_fwd_simt(const pvg_conv_desc* d, const float* x, const float* w, const float* bias, float* y, cudaStream_t st);
int conv2d_fwd_h3(const pvg_conv_desc* d, const void* x_planes, const void* w_planes, const float* bias, float* y,
                  void* y_planes, const float* out_scale, cudaStream_t st);      // conv_h3.cu

constexpr int kThreads = 192;
constexpr int kTileM = 128;
constexpr int kSmemBudget = 200 * 1024;

// KC = channels per pipeline stage: 32 (128-byte rows, SWIZZLE_128B) or 16 (64-byte rows, SWIZZLE_64B: half the bytes per
// stage -> twice the stages in the same shared memory, i.e. a deeper TMA prefetch for the latency-bound 3xTF32 tiles)
template <int BN, int NPROD, int KC>
struct Cfg {
  // NPROD == 2: the second 'plane' holds the two 16-bit half-width tiles.  NPROD == 4 (all three products as kind::f16 MMAs on
  // the fp16 plane pair, PVG_CORR_FP16_ALL): the ONLY plane is that pair - 2 x 64-byte rows = the bytes of one fp32 row
  static constexpr int kPlanes = (NPROD == 2 || NPROD == 3) ? 2 : 1;
  static constexpr int kRowBytes = KC * 4;
  static constexpr int kABytes = kTileM * kRowBytes;
  static constexpr int kBBytes = BN * kRowBytes;
  static constexpr int kKSteps = KC / 8;                      // MMAs (K = 8) per operand pair and stage
  static constexpr int kStageBytes = kPlanes * (kABytes + kBBytes);
  static constexpr int kStagesRaw = kSmemBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr bool kSplitAcc = NPROD >= 2;               // see "split accum" in the header comment
  static constexpr int kDrain = 32 / kKSteps;                 // stages per TMEM accumulation chain (32 MMAs of the main term)
  static constexpr int kAccCols = kSplitAcc ? 3 * BN : BN;    // [main0 | main1 | correction] or [acc]
  static constexpr int kTmemCols = kAccCols <= 32 ? 32 : (kAccCols <= 64 ? 64 : (kAccCols <= 128 ? 128 : (kAccCols <= 256 ? 256 : 512)));
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(kStages >= 2, "need at least a double buffer");
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "invalid UMMA N");
  static_assert(kAccCols <= 512, "accumulators exceed TMEM");
  static_assert(KC == 32 || KC == 16, "KC must be 16 or 32");
  static_assert((NPROD != 2 && NPROD != 4) || KC == 32, "16-bit planes use 32-channel stages");
};

template <int BN, int NPROD, int KC>
__global__ void __launch_bounds__(kThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo, const ConvParams p) {
  using C = Cfg<BN, NPROD, KC>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  uint64_t* full_bar = (uint64_t*)(smem + C::kStages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tfull_bar = empty_bar + C::kStages;      // [2] accumulator buffer ready for the epilogue
  uint64_t* tempty_bar = tfull_bar + 2;              // [2] accumulator buffer drained, MMA may overwrite it
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile coordinates: the N tile index varies fastest, so the CTAs that re-read one activation patch run back to back
  // and hit L2 (with N-major order the second pass over a > 126 MB tensor came from DRAM again: 4x traffic on conv4)
  const int n_tiles = (p.Cout + BN - 1) / BN;
  const int n_tile = blockIdx.x % n_tiles;
  int t = blockIdx.x / n_tiles;
  const int tile_w = t % p.tiles_w; t /= p.tiles_w;
  const int tile_h = t % p.tiles_h; t /= p.tiles_h;
  const int tile_n = t;
  const int w0 = tile_w * p.tw, h0 = tile_h * p.th, n0 = tile_n * p.tn;
  const int co0 = n_tile * BN;
  const int chunks = p.Cin / KC;
  const int k_iters = p.R * p.S * chunks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB);
    if (NPROD >= 2) { prefetch_tmap(&tmAlo); prefetch_tmap(&tmBlo); }
    for (int s = 0; s < C::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tfull_bar[0], 1); mbar_init(&tfull_bar[1], 1);
    mbar_init(&tempty_bar[0], 128); mbar_init(&tempty_bar[1], 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    const uint32_t leader = elect_one();
    int stage = 0; uint32_t phase = 0;
    for (int k = 0; k < k_iters; ++k) {
      const int tap = k / chunks, cc = k - tap * chunks;
      const int r = tap / p.S, s = tap - r * p.S;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (leader) {
        uint8_t* st = stage_base + stage * C::kStageBytes;
        mbar_expect_tx(&full_bar[stage], C::kStageBytes);
        if (NPROD == 4) {             // fp16 plane pairs only: [f16(lo * 2^12) | f16(x)] tiles of A, then of B
          tma_load_5d(st, &tmAlo, &full_bar[stage], cc * KC, w0 + s - p.pad, h0 + r - p.pad, n0, 0);
          tma_load_3d(st + C::kABytes, &tmBlo, &full_bar[stage], k * KC, co0, 0);
        } else {
        tma_load_4d(st, &tmA, &full_bar[stage], cc * KC, w0 + s - p.pad, h0 + r - p.pad, n0);
        tma_load_2d(st + C::kPlanes * C::kABytes, &tmB, &full_bar[stage], k * KC, co0);
        }
        if (NPROD == 3) {
          tma_load_4d(st + C::kABytes, &tmAlo, &full_bar[stage], cc * KC, w0 + s - p.pad, h0 + r - p.pad, n0);
          tma_load_2d(st + 2 * C::kABytes + C::kBBytes, &tmBlo, &full_bar[stage], k * KC, co0);
        } else if (NPROD == 2) {      // both 16-bit planes of each operand in one box: [f16(lo * 2^12) tile | f16(x) tile]
          tma_load_5d(st + C::kABytes, &tmAlo, &full_bar[stage], cc * KC, w0 + s - p.pad, h0 + r - p.pad, n0, 0);
          tma_load_3d(st + 2 * C::kABytes + C::kBBytes, &tmBlo, &full_bar[stage], k * KC, co0, 0);
        }
      }
      __syncwarp();
      if (++stage == C::kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Descriptors are base + offset in 16-byte units (the start-address field is the low 14 bits; no carry can leave it).
    const uint32_t leader = elect_one();
    constexpr uint32_t idesc = make_idesc_tf32<BN>();
    const uint64_t d32 = make_kmajor_desc<KC>(smem_u32(stage_base));       // fp32 tiles (KC*4-byte rows)
    const uint64_t d16 = make_kmajor_desc<16>(smem_u32(stage_base));       // 16-bit tiles (64-byte rows, SWIZZLE_64B)
    constexpr uint32_t kStageU = C::kStageBytes >> 4, kAU = C::kABytes >> 4, kBU = C::kBBytes >> 4;
    int stage = 0; uint32_t phase = 0;
    if constexpr (C::kSplitAcc) {
      const uint32_t corr = tmem_acc + 2 * BN;
      const uint32_t idesc16 = make_idesc_f16<BN>(p.corr_fp16 != 0);
      uint32_t corr_acc = 0;
      const int periods = (k_iters + C::kDrain - 1) / C::kDrain;
      int k = 0;
      for (int per = 0; per < periods; ++per) {
        const int b = per & 1;
        mbar_wait(&tempty_bar[b], ((per >> 1) & 1) ^ 1);       // epilogue finished draining this buffer
        tc_fence_after();
        const uint32_t main_acc = tmem_acc + b * BN;
        const int k_end = min(k + C::kDrain, k_iters);
        uint32_t main_started = 0;
        for (; k < k_end; ++k) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (leader) {
            const uint32_t a_hi = stage * kStageU, a_lo = a_hi + kAU, b_hi = a_hi + 2 * kAU, b_lo = b_hi + kBU;
            if constexpr (NPROD == 4) {
              // stage = [f16(x_lo * 2^12) | f16(x)] tiles of A, then [f16(w_lo * 2^12) | f16(w)] tiles of B (64-byte rows):
              // corrections and main product are all kind::f16 MMAs with K = 16
              const uint32_t pa_lo = a_hi, pa_x = a_hi + kAU / 2, pb_lo = a_hi + kAU, pb_x = pb_lo + kBU / 2;
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                umma_bf16(corr, d16 + (pa_lo + 2 * ks), d16 + (pb_x + 2 * ks), idesc16, corr_acc);
                corr_acc = 1;
              }
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) umma_bf16(corr, d16 + (pa_x + 2 * ks), d16 + (pb_lo + 2 * ks), idesc16, 1);
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                umma_bf16(main_acc, d16 + (pa_x + 2 * ks), d16 + (pb_x + 2 * ks), idesc16, main_started);
                main_started = 1;
              }
            } else {
            if constexpr (NPROD == 2) {
              // a_lo region: [f16(x_lo * 2^12) | f16(x)] tiles, b_lo region: [f16(w_lo * 2^12) | f16(w_hi)] tiles, 64-byte rows,
              // K = 16 per MMA; the accumulator holds 2^12 x the correction
              const uint32_t a_xb = a_lo + kAU / 2, b_xb = b_lo + kBU / 2;
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                umma_bf16(corr, d16 + (a_lo + 2 * ks), d16 + (b_xb + 2 * ks), idesc16, corr_acc);
                corr_acc = 1;
              }
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) umma_bf16(corr, d16 + (a_xb + 2 * ks), d16 + (b_lo + 2 * ks), idesc16, 1);
            } else {
#pragma unroll
              for (int ks = 0; ks < C::kKSteps; ++ks) {
                umma_tf32(corr, d32 + (a_lo + 2 * ks), d32 + (b_hi + 2 * ks), idesc, corr_acc);
                corr_acc = 1;
              }
#pragma unroll
              for (int ks = 0; ks < C::kKSteps; ++ks) umma_tf32(corr, d32 + (a_hi + 2 * ks), d32 + (b_lo + 2 * ks), idesc, 1);
            }
#pragma unroll
            for (int ks = 0; ks < C::kKSteps; ++ks) {
              umma_tf32(main_acc, d32 + (a_hi + 2 * ks), d32 + (b_hi + 2 * ks), idesc, main_started);
              main_started = 1;
            }
            }
            umma_commit(&empty_bar[stage]);     // slot reusable once these MMAs have read it
          }
          __syncwarp();
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        if (leader) umma_commit(&tfull_bar[b]);   // this chain (and every earlier MMA, incl. corrections) is complete
        __syncwarp();
      }
    } else {
      uint32_t accumulate = 0;
      for (int k = 0; k < k_iters; ++k) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (leader) {
          const uint32_t a_hi = stage * kStageU, b_hi = a_hi + kAU;
#pragma unroll
          for (int ks = 0; ks < C::kKSteps; ++ks) {
            umma_tf32(tmem_acc, d32 + (a_hi + 2 * ks), d32 + (b_hi + 2 * ks), idesc, accumulate);
            accumulate = 1;
          }
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
      if (leader) umma_commit(&tfull_bar[0]);     // accumulator complete
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int wi = row % p.tw;
    const int hi = (row / p.tw) % p.th;
    const int ni = row / (p.tw * p.th);
    const int ow = w0 + wi, oh = h0 + hi, on = n0 + ni;
    const bool valid = ow < p.W && oh < p.H && on < p.N;
    float* yrow = p.y + (((int64_t)on * p.H + oh) * p.W + ow) * p.Cout;
    const bool vec_ok = (p.Cout % 4) == 0 && (((uintptr_t)p.y | (uintptr_t)p.bias) & 15) == 0;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    if constexpr (C::kSplitAcc) {
      float acc[BN];
#pragma unroll
      for (int j = 0; j < BN; ++j) acc[j] = 0.f;
      const int periods = (k_iters + C::kDrain - 1) / C::kDrain;
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
      // the last tfull commit also covers the correction MMAs (NPROD == 2: accumulated at 2^12 x their value)
      constexpr float kCorrScale = (NPROD == 2 || NPROD == 4) ? 0x1p-12f : 1.f;
#pragma unroll
      for (int c = 0; c < BN; c += 16) {
        float v[16];
        tmem_ld16(tmem_acc + lane_base + (uint32_t)(2 * BN + c), v);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaf(v[j], kCorrScale, acc[c + j]);
        finish16(v, p.bias, co0 + c, p.Cout, p.act, p.slope, yrow, valid, vec_ok);
      }
    } else {
      mbar_wait(&tfull_bar[0], 0);
      tc_fence_after();
#pragma unroll 2
      for (int c = 0; c < BN; c += 16) {
        float v[16];
        tmem_ld16(tmem_acc + lane_base + (uint32_t)c, v);
        finish16(v, p.bias, co0 + c, p.Cout, p.act, p.slope, yrow, valid, vec_ok);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_acc, C::kTmemCols);
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent variant of the 1-CTA split-product kernel (the default for 64- and 128-wide tiles; PVG_PERSISTENT=0 disables).  One CTA per SM walks
// tile = blockIdx.x, blockIdx.x + gridDim.x, ... so that (1) barrier init, TMEM allocation and tensor-map prefetch are
// paid once per SM instead of once per tile, (2) the TMA producer prefetches the next tile's first stages while the
// current tile drains, and (3) the bias / activation / store part of the epilogue overlaps the next tile's MMAs: the
// epilogue warps fold the correction accumulator into their registers first (TMEM reads only), release it through the
// `cfree` barrier, and only then do the arithmetic and the global stores.  profiles/r01_tile_model.md measures the
// non-overlapped fixed cost this removes at 3.7 us per tile (9-16 % of a 36-72 k-iteration tile).
// Barrier phases simply keep counting across tiles: `pg` = global period index (main-accumulator ping-pong), `it` = tiles
// done by this CTA (correction-accumulator hand-off).
// ---------------------------------------------------------------------------------------------------------------
template <int BN, int NPROD>
__global__ void __launch_bounds__(kThreads, 1)
conv_umma_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
                            const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo,
                            const ConvParams p, const int total_tiles) {
  static_assert(NPROD >= 2, "the persistent variant implements the split-product path only");
  constexpr int KC = 32;
  using C = Cfg<BN, NPROD, KC>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  uint64_t* full_bar = (uint64_t*)(smem + C::kStages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tfull_bar = empty_bar + C::kStages;      // [2] main accumulator buffer ready for the epilogue
  uint64_t* tempty_bar = tfull_bar + 2;              // [2] main accumulator buffer drained
  uint64_t* cfree_bar = tempty_bar + 2;              // [1] correction accumulator read by the epilogue: next tile may overwrite it
  uint32_t* tmem_slot = (uint32_t*)(cfree_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.Cout + BN - 1) / BN;
  const int chunks = p.Cin / KC;
  const int k_iters = p.R * p.S * chunks;
  const int periods = (k_iters + C::kDrain - 1) / C::kDrain;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB); prefetch_tmap(&tmAlo); prefetch_tmap(&tmBlo);
    for (int s = 0; s < C::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tfull_bar[0], 1); mbar_init(&tfull_bar[1], 1);
    mbar_init(&tempty_bar[0], 128); mbar_init(&tempty_bar[1], 128);
    mbar_init(cfree_bar, 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    const uint32_t leader = elect_one();
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_tile = tile % n_tiles;
      int t = tile / n_tiles;
      const int tile_w = t % p.tiles_w; t /= p.tiles_w;
      const int tile_h = t % p.tiles_h; t /= p.tiles_h;
      const int w0 = tile_w * p.tw, h0 = tile_h * p.th, n0 = t * p.tn;
      const int co0 = n_tile * BN;
      for (int k = 0; k < k_iters; ++k) {
        const int tap = k / chunks, cc = k - tap * chunks;
        const int r = tap / p.S, s = tap - r * p.S;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) {
          uint8_t* st = stage_base + stage * C::kStageBytes;
          mbar_expect_tx(&full_bar[stage], C::kStageBytes);
          tma_load_4d(st, &tmA, &full_bar[stage], cc * KC, w0 + s - p.pad, h0 + r - p.pad, n0);
          tma_load_2d(st + 2 * C::kABytes, &tmB, &full_bar[stage], k * KC, co0);
          if (NPROD == 3) {
            tma_load_4d(st + C::kABytes, &tmAlo, &full_bar[stage], cc * KC, w0 + s - p.pad, h0 + r - p.pad, n0);
            tma_load_2d(st + 2 * C::kABytes + C::kBBytes, &tmBlo, &full_bar[stage], k * KC, co0);
          } else {
            tma_load_5d(st + C::kABytes, &tmAlo, &full_bar[stage], cc * KC, w0 + s - p.pad, h0 + r - p.pad, n0, 0);
            tma_load_3d(st + 2 * C::kABytes + C::kBBytes, &tmBlo, &full_bar[stage], k * KC, co0, 0);
          }
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t leader = elect_one();
    constexpr uint32_t idesc = make_idesc_tf32<BN>();
    const uint32_t idesc16 = make_idesc_f16<BN>(p.corr_fp16 != 0);
    const uint64_t d32 = make_kmajor_desc<KC>(smem_u32(stage_base));
    const uint64_t d16 = make_kmajor_desc<16>(smem_u32(stage_base));
    constexpr uint32_t kStageU = C::kStageBytes >> 4, kAU = C::kABytes >> 4, kBU = C::kBBytes >> 4;
    const uint32_t corr = tmem_acc + 2 * BN;
    int stage = 0; uint32_t phase = 0;
    uint32_t pg = 0;                                   // global period index: main buffer = pg & 1
    int it = 0;                                        // tiles done by this CTA
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      if (it > 0) {                                    // the epilogue has folded the previous tile's corrections into registers
        mbar_wait(cfree_bar, (uint32_t)(it - 1) & 1);
        tc_fence_after();
      }
      uint32_t corr_acc = 0;
      int k = 0;
      for (int per = 0; per < periods; ++per, ++pg) {
        const uint32_t b = pg & 1;
        mbar_wait(&tempty_bar[b], ((pg >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t main_acc = tmem_acc + b * BN;
        const int k_end = min(k + C::kDrain, k_iters);
        uint32_t main_started = 0;
        for (; k < k_end; ++k) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (leader) {
            const uint32_t a_hi = stage * kStageU, a_lo = a_hi + kAU, b_hi = a_hi + 2 * kAU, b_lo = b_hi + kBU;
            if constexpr (NPROD == 2) {
              const uint32_t a_xb = a_lo + kAU / 2, b_xb = b_lo + kBU / 2;
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                umma_bf16(corr, d16 + (a_lo + 2 * ks), d16 + (b_xb + 2 * ks), idesc16, corr_acc);
                corr_acc = 1;
              }
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) umma_bf16(corr, d16 + (a_xb + 2 * ks), d16 + (b_lo + 2 * ks), idesc16, 1);
            } else {
#pragma unroll
              for (int ks = 0; ks < C::kKSteps; ++ks) {
                umma_tf32(corr, d32 + (a_lo + 2 * ks), d32 + (b_hi + 2 * ks), idesc, corr_acc);
                corr_acc = 1;
              }
#pragma unroll
              for (int ks = 0; ks < C::kKSteps; ++ks) umma_tf32(corr, d32 + (a_hi + 2 * ks), d32 + (b_lo + 2 * ks), idesc, 1);
            }
#pragma unroll
            for (int ks = 0; ks < C::kKSteps; ++ks) {
              umma_tf32(main_acc, d32 + (a_hi + 2 * ks), d32 + (b_hi + 2 * ks), idesc, main_started);
              main_started = 1;
            }
            umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        if (leader) umma_commit(&tfull_bar[b]);         // the tile's last commit also covers its correction MMAs
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int wi = row % p.tw;
    const int hi = (row / p.tw) % p.th;
    const int ni = row / (p.tw * p.th);
    const bool vec_ok = (p.Cout % 4) == 0 && (((uintptr_t)p.y | (uintptr_t)p.bias) & 15) == 0;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    constexpr float kCorrScale = NPROD == 2 ? 0x1p-12f : 1.f;
    uint32_t pg = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_tile = tile % n_tiles;
      int t = tile / n_tiles;
      const int tile_w = t % p.tiles_w; t /= p.tiles_w;
      const int tile_h = t % p.tiles_h; t /= p.tiles_h;
      const int ow = tile_w * p.tw + wi, oh = tile_h * p.th + hi, on = t * p.tn + ni;
      const int co0 = n_tile * BN;
      const bool valid = ow < p.W && oh < p.H && on < p.N;
      float* yrow = p.y + (((int64_t)on * p.H + oh) * p.W + ow) * p.Cout;
      float acc[BN];
#pragma unroll
      for (int j = 0; j < BN; ++j) acc[j] = 0.f;
      for (int per = 0; per < periods; ++per, ++pg) {
        const uint32_t b = pg & 1;
        mbar_wait(&tfull_bar[b], (pg >> 1) & 1);
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
      // TMEM reads only: fold the corrections into the registers, then hand the correction buffer back to the MMA warp
#pragma unroll
      for (int c = 0; c < BN; c += 16) {
        float v[16];
        tmem_ld16(tmem_acc + lane_base + (uint32_t)(2 * BN + c), v);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[c + j] = fmaf(v[j], kCorrScale, acc[c + j]);
      }
      tc_fence_before();
      mbar_arrive(cfree_bar);
      // bias / activation / stores overlap the next tile's main loop
#pragma unroll
      for (int c = 0; c < BN; c += 16) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = acc[c + j];
        finish16(v, p.bias, co0 + c, p.Cout, p.act, p.slope, yrow, valid, vec_ok);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_acc, C::kTmemCols);
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two CTAs of a cluster compute a 256-pixel x 128-channel tile with ONE MMA stream.
// Each CTA stages its own 128-pixel A patch and HALF of the weight tile (64 rows), so every MMA reads 4 KB (A) + 2 KB (B)
// of shared memory per SM instead of 4 + 4 KB and TMA fills 48 KB instead of 64 KB per k-iteration: the shared-memory
// bandwidth bound (profiles/r01_conv_umma_ncu_full.md) moves from 60 % to ~80 % tensor-pipe utilisation, and the smaller
// stage buys a 4th pipeline stage.  Protocol: both CTAs' TMA loads credit the LEADER's `full` barrier; the leader's MMA
// thread issues tcgen05.mma.cta_group::2 (M = 256: rows 0-127 land in the leader's TMEM, 128-255 in the peer's) and its
// commits are multicast to the `empty` / `tfull` barriers of BOTH CTAs; each CTA's epilogue drains its own TMEM and
// arrives on the leader's `tempty` barrier (256 arrivals).
// ---------------------------------------------------------------------------------------------------------------
template <int NPROD>
struct Cfg2 {
  static constexpr int BN = 128;                              // channels per pair tile
  static constexpr int kPlanes = NPROD >= 2 ? 2 : 1;
  static constexpr int kABytes = kTileM * 128;                // own 128-pixel patch, 32 channels
  static constexpr int kBBytes = (BN / 2) * 128;              // own half of the weight tile
  static constexpr int kStageBytes = kPlanes * (kABytes + kBBytes);
  static constexpr int kStagesRaw = kSmemBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr bool kSplitAcc = NPROD >= 2;
  static constexpr int kDrain = 8;
  static constexpr int kAccCols = kSplitAcc ? 3 * BN : BN;
  static constexpr int kTmemCols = kAccCols <= 128 ? 128 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

template <int NPROD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_umma2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
                  const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo, const ConvParams p) {
  using C = Cfg2<NPROD>;
  constexpr int BN = C::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  uint64_t* full_bar = (uint64_t*)(smem + C::kStages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tfull_bar = empty_bar + C::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  const int n_tiles = (p.Cout + BN - 1) / BN;
  const int cid = blockIdx.x >> 1;
  const int n_tile = cid % n_tiles;
  int t = (cid / n_tiles) * 2 + (int)rank;                 // this CTA's 128-pixel tile (may lie past the end: all OOB)
  const int tile_w = t % p.tiles_w; t /= p.tiles_w;
  const int tile_h = t % p.tiles_h; t /= p.tiles_h;
  const int tile_n = t;
  const int w0 = tile_w * p.tw, h0 = tile_h * p.th, n0 = tile_n * p.tn;
  const int co0 = n_tile * BN;
  const int chunks = p.Cin / 32;
  const int k_iters = p.R * p.S * chunks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB);
    if (NPROD >= 2) { prefetch_tmap(&tmAlo); prefetch_tmap(&tmBlo); }
    for (int s = 0; s < C::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tfull_bar[0], 1); mbar_init(&tfull_bar[1], 1);
    mbar_init(&tempty_bar[0], 256); mbar_init(&tempty_bar[1], 256);
    fence_barrier_init();
  }
  if (warp == 1) tmem2_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    const uint32_t elected = elect_one();
    int stage = 0; uint32_t phase = 0;
    for (int k = 0; k < k_iters; ++k) {
      const int tap = k / chunks, cc = k - tap * chunks;
      const int r = tap / p.S, s = tap - r * p.S;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elected) {
        uint8_t* st = stage_base + stage * C::kStageBytes;
        if (leader) mbar_expect_tx(&full_bar[stage], 2 * C::kStageBytes);      // bytes of BOTH CTAs land on this barrier
        tma2_load_4d(st, &tmA, &full_bar[stage], cc * 32, w0 + s - p.pad, h0 + r - p.pad, n0);
        tma2_load_2d(st + C::kPlanes * C::kABytes, &tmB, &full_bar[stage], k * 32, co0 + (int)rank * (BN / 2));
        if (NPROD == 3) {
          tma2_load_4d(st + C::kABytes, &tmAlo, &full_bar[stage], cc * 32, w0 + s - p.pad, h0 + r - p.pad, n0);
          tma2_load_2d(st + 2 * C::kABytes + C::kBBytes, &tmBlo, &full_bar[stage], k * 32, co0 + (int)rank * (BN / 2));
        } else if (NPROD == 2) {
          tma2_load_5d(st + C::kABytes, &tmAlo, &full_bar[stage], cc * 32, w0 + s - p.pad, h0 + r - p.pad, n0, 0);
          tma2_load_3d(st + 2 * C::kABytes + C::kBBytes, &tmBlo, &full_bar[stage], k * 32, co0 + (int)rank * (BN / 2), 0);
        }
      }
      __syncwarp();
      if (++stage == C::kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      const uint32_t elected = elect_one();
      constexpr uint32_t idesc = make_idesc_tf32_m256<BN>();
      const uint64_t d32 = make_kmajor_desc<32>(smem_u32(stage_base));
      const uint64_t d16 = make_kmajor_desc<16>(smem_u32(stage_base));
      constexpr uint32_t kStageU = C::kStageBytes >> 4, kAU = C::kABytes >> 4, kBU = C::kBBytes >> 4;
      int stage = 0; uint32_t phase = 0;
      if constexpr (C::kSplitAcc) {
        const uint32_t corr = tmem_acc + 2 * BN;
        const uint32_t idesc16 = make_idesc_f16<BN>(p.corr_fp16 != 0, false, false, 256);
        uint32_t corr_acc = 0;
        const int periods = (k_iters + C::kDrain - 1) / C::kDrain;
        int k = 0;
        for (int per = 0; per < periods; ++per) {
          const int b = per & 1;
          mbar_wait(&tempty_bar[b], ((per >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t main_acc = tmem_acc + b * BN;
          const int k_end = min(k + C::kDrain, k_iters);
          uint32_t main_started = 0;
          for (; k < k_end; ++k) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (elected) {
              const uint32_t a_hi = stage * kStageU, a_lo = a_hi + kAU, b_hi = a_hi + 2 * kAU, b_lo = b_hi + kBU;
              if constexpr (NPROD == 2) {
                const uint32_t a_xb = a_lo + kAU / 2, b_xb = b_lo + kBU / 2;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                  umma2_bf16(corr, d16 + (a_lo + 2 * ks), d16 + (b_xb + 2 * ks), idesc16, corr_acc);
                  corr_acc = 1;
                }
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) umma2_bf16(corr, d16 + (a_xb + 2 * ks), d16 + (b_lo + 2 * ks), idesc16, 1);
              } else {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  umma2_tf32(corr, d32 + (a_lo + 2 * ks), d32 + (b_hi + 2 * ks), idesc, corr_acc);
                  corr_acc = 1;
                }
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma2_tf32(corr, d32 + (a_hi + 2 * ks), d32 + (b_lo + 2 * ks), idesc, 1);
              }
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                umma2_tf32(main_acc, d32 + (a_hi + 2 * ks), d32 + (b_hi + 2 * ks), idesc, main_started);
                main_started = 1;
              }
              umma2_commit_both(&empty_bar[stage]);
            }
            __syncwarp();
            if (++stage == C::kStages) { stage = 0; phase ^= 1; }
          }
          if (elected) umma2_commit_both(&tfull_bar[b]);
          __syncwarp();
        }
      } else {
        uint32_t accumulate = 0;
        for (int k = 0; k < k_iters; ++k) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elected) {
            const uint32_t a_hi = stage * kStageU, b_hi = a_hi + kAU;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              umma2_tf32(tmem_acc, d32 + (a_hi + 2 * ks), d32 + (b_hi + 2 * ks), idesc, accumulate);
              accumulate = 1;
            }
            umma2_commit_both(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        if (elected) umma2_commit_both(&tfull_bar[0]);
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 of both CTAs; each CTA owns its 128 TMEM lanes) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int wi = row % p.tw;
    const int hi = (row / p.tw) % p.th;
    const int ni = row / (p.tw * p.th);
    const int ow = w0 + wi, oh = h0 + hi, on = n0 + ni;
    const bool valid = ow < p.W && oh < p.H && on < p.N;
    float* yrow = p.y + (((int64_t)on * p.H + oh) * p.W + ow) * p.Cout;
    const bool vec_ok = (p.Cout % 4) == 0 && (((uintptr_t)p.y | (uintptr_t)p.bias) & 15) == 0;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float acc[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) acc[j] = 0.f;
    if constexpr (C::kSplitAcc) {
      const int periods = (k_iters + C::kDrain - 1) / C::kDrain;
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
        mbar_arrive_leader(&tempty_bar[b]);
      }
    } else {
      mbar_wait(&tfull_bar[0], 0);
      tc_fence_after();
    }
    constexpr int corr_col = C::kSplitAcc ? 2 * BN : 0;
    constexpr float kCorrScale = NPROD == 2 ? 0x1p-12f : 1.f;
#pragma unroll
    for (int c = 0; c < BN; c += 16) {
      float v[16];
      tmem_ld16(tmem_acc + lane_base + (uint32_t)(corr_col + c), v);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaf(v[j], kCorrScale, acc[c + j]);
      finish16(v, p.bias, co0 + c, p.Cout, p.act, p.slope, yrow, valid, vec_ok);
    }
  }
  tc_fence_before();
  cluster_sync_all();          // the peer's shared memory / TMEM must stay alive until every MMA that reads it has retired
  if (warp == 1) tmem2_dealloc(tmem_acc, C::kTmemCols);
}

// Persistent variant of the CTA-pair kernel (default, like conv_umma_persistent_kernel).  One cluster per SM pair walks work items cid = cluster, cluster + #clusters, ...; the hand-off of the correction
// accumulators (one per CTA, both written by the leader's cta_group::2 MMAs) goes through the leader's `cfree` barrier with
// 256 arrivals, exactly like `tempty`.
template <int NPROD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_umma2_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
                             const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo,
                             const ConvParams p, const int total_items) {
  static_assert(NPROD >= 2, "the persistent variant implements the split-product path only");
  using C = Cfg2<NPROD>;
  constexpr int BN = C::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  uint64_t* full_bar = (uint64_t*)(smem + C::kStages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tfull_bar = empty_bar + C::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* cfree_bar = tempty_bar + 2;
  uint32_t* tmem_slot = (uint32_t*)(cfree_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int first_item = blockIdx.x >> 1, item_stride = gridDim.x >> 1;
  const int n_tiles = (p.Cout + BN - 1) / BN;
  const int chunks = p.Cin / 32;
  const int k_iters = p.R * p.S * chunks;
  const int periods = (k_iters + C::kDrain - 1) / C::kDrain;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB); prefetch_tmap(&tmAlo); prefetch_tmap(&tmBlo);
    for (int s = 0; s < C::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tfull_bar[0], 1); mbar_init(&tfull_bar[1], 1);
    mbar_init(&tempty_bar[0], 256); mbar_init(&tempty_bar[1], 256);
    mbar_init(cfree_bar, 256);
    fence_barrier_init();
  }
  if (warp == 1) tmem2_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    const uint32_t elected = elect_one();
    int stage = 0; uint32_t phase = 0;
    for (int cid = first_item; cid < total_items; cid += item_stride) {
      const int n_tile = cid % n_tiles;
      int t = (cid / n_tiles) * 2 + (int)rank;               // this CTA's 128-pixel tile (may lie past the end: all OOB)
      const int tile_w = t % p.tiles_w; t /= p.tiles_w;
      const int tile_h = t % p.tiles_h; t /= p.tiles_h;
      const int w0 = tile_w * p.tw, h0 = tile_h * p.th, n0 = t * p.tn;
      const int co0 = n_tile * BN;
      for (int k = 0; k < k_iters; ++k) {
        const int tap = k / chunks, cc = k - tap * chunks;
        const int r = tap / p.S, s = tap - r * p.S;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elected) {
          uint8_t* st = stage_base + stage * C::kStageBytes;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * C::kStageBytes);
          tma2_load_4d(st, &tmA, &full_bar[stage], cc * 32, w0 + s - p.pad, h0 + r - p.pad, n0);
          tma2_load_2d(st + 2 * C::kABytes, &tmB, &full_bar[stage], k * 32, co0 + (int)rank * (BN / 2));
          if (NPROD == 3) {
            tma2_load_4d(st + C::kABytes, &tmAlo, &full_bar[stage], cc * 32, w0 + s - p.pad, h0 + r - p.pad, n0);
            tma2_load_2d(st + 2 * C::kABytes + C::kBBytes, &tmBlo, &full_bar[stage], k * 32, co0 + (int)rank * (BN / 2));
          } else {
            tma2_load_5d(st + C::kABytes, &tmAlo, &full_bar[stage], cc * 32, w0 + s - p.pad, h0 + r - p.pad, n0, 0);
            tma2_load_3d(st + 2 * C::kABytes + C::kBBytes, &tmBlo, &full_bar[stage], k * 32, co0 + (int)rank * (BN / 2), 0);
          }
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      const uint32_t elected = elect_one();
      constexpr uint32_t idesc = make_idesc_tf32_m256<BN>();
      const uint32_t idesc16 = make_idesc_f16<BN>(p.corr_fp16 != 0, false, false, 256);
      const uint64_t d32 = make_kmajor_desc<32>(smem_u32(stage_base));
      const uint64_t d16 = make_kmajor_desc<16>(smem_u32(stage_base));
      constexpr uint32_t kStageU = C::kStageBytes >> 4, kAU = C::kABytes >> 4, kBU = C::kBBytes >> 4;
      const uint32_t corr = tmem_acc + 2 * BN;
      int stage = 0; uint32_t phase = 0;
      uint32_t pg = 0;
      int it = 0;
      for (int cid = first_item; cid < total_items; cid += item_stride, ++it) {
        if (it > 0) {                                   // both CTAs' epilogues have folded the previous corrections into registers
          mbar_wait(cfree_bar, (uint32_t)(it - 1) & 1);
          tc_fence_after();
        }
        uint32_t corr_acc = 0;
        int k = 0;
        for (int per = 0; per < periods; ++per, ++pg) {
          const uint32_t b = pg & 1;
          mbar_wait(&tempty_bar[b], ((pg >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t main_acc = tmem_acc + b * BN;
          const int k_end = min(k + C::kDrain, k_iters);
          uint32_t main_started = 0;
          for (; k < k_end; ++k) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (elected) {
              const uint32_t a_hi = stage * kStageU, a_lo = a_hi + kAU, b_hi = a_hi + 2 * kAU, b_lo = b_hi + kBU;
              if constexpr (NPROD == 2) {
                const uint32_t a_xb = a_lo + kAU / 2, b_xb = b_lo + kBU / 2;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                  umma2_bf16(corr, d16 + (a_lo + 2 * ks), d16 + (b_xb + 2 * ks), idesc16, corr_acc);
                  corr_acc = 1;
                }
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) umma2_bf16(corr, d16 + (a_xb + 2 * ks), d16 + (b_lo + 2 * ks), idesc16, 1);
              } else {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  umma2_tf32(corr, d32 + (a_lo + 2 * ks), d32 + (b_hi + 2 * ks), idesc, corr_acc);
                  corr_acc = 1;
                }
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma2_tf32(corr, d32 + (a_hi + 2 * ks), d32 + (b_lo + 2 * ks), idesc, 1);
              }
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                umma2_tf32(main_acc, d32 + (a_hi + 2 * ks), d32 + (b_hi + 2 * ks), idesc, main_started);
                main_started = 1;
              }
              umma2_commit_both(&empty_bar[stage]);
            }
            __syncwarp();
            if (++stage == C::kStages) { stage = 0; phase ^= 1; }
          }
          if (elected) umma2_commit_both(&tfull_bar[b]);
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 of both CTAs) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int wi = row % p.tw;
    const int hi = (row / p.tw) % p.th;
    const int ni = row / (p.tw * p.th);
    const bool vec_ok = (p.Cout % 4) == 0 && (((uintptr_t)p.y | (uintptr_t)p.bias) & 15) == 0;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    constexpr float kCorrScale = NPROD == 2 ? 0x1p-12f : 1.f;
    uint32_t pg = 0;
    for (int cid = first_item; cid < total_items; cid += item_stride) {
      const int n_tile = cid % n_tiles;
      int t = (cid / n_tiles) * 2 + (int)rank;
      const int tile_w = t % p.tiles_w; t /= p.tiles_w;
      const int tile_h = t % p.tiles_h; t /= p.tiles_h;
      const int ow = tile_w * p.tw + wi, oh = tile_h * p.th + hi, on = t * p.tn + ni;
      const int co0 = n_tile * BN;
      const bool valid = ow < p.W && oh < p.H && on < p.N;
      float* yrow = p.y + (((int64_t)on * p.H + oh) * p.W + ow) * p.Cout;
      float acc[BN];
#pragma unroll
      for (int j = 0; j < BN; ++j) acc[j] = 0.f;
      for (int per = 0; per < periods; ++per, ++pg) {
        const uint32_t b = pg & 1;
        mbar_wait(&tfull_bar[b], (pg >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < BN; c += 16) {
          float v[16];
          tmem_ld16(tmem_acc + lane_base + (uint32_t)(b * BN + c), v);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[c + j] += v[j];
        }
        tc_fence_before();
        mbar_arrive_leader(&tempty_bar[b]);
      }
#pragma unroll
      for (int c = 0; c < BN; c += 16) {
        float v[16];
        tmem_ld16(tmem_acc + lane_base + (uint32_t)(2 * BN + c), v);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[c + j] = fmaf(v[j], kCorrScale, acc[c + j]);
      }
      tc_fence_before();
      mbar_arrive_leader(cfree_bar);
#pragma unroll
      for (int c = 0; c < BN; c += 16) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = acc[c + j];
        finish16(v, p.bias, co0 + c, p.Cout, p.act, p.slope, yrow, valid, vec_ok);
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();          // the peer's shared memory / TMEM must stay alive until every MMA that reads it has retired
  if (warp == 1) tmem2_dealloc(tmem_acc, C::kTmemCols);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

int encode_nhwc_map(CUtensorMap* m, const float* x, int N, int H, int W, int C, int box_c, int tw, int th, int tn, bool atom32,
                    bool sw64) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled not available"); return -3; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  sw64 ? CU_TENSOR_MAP_SWIZZLE_64B : (atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(activations) failed: " + std::to_string((int)r)); return -3; }
  return 0;
}

int encode_nhwc_16x2_map(CUtensorMap* m, const void* planes, int N, int H, int W, int C, int tw, int th, int tn) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled not available"); return -3; }
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, 2};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)N * H * W * C * 2};
  cuuint32_t box[5] = {32, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tn, 2};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)planes, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(bf16 activation planes) failed: " + std::to_string((int)r)); return -3; }
  return 0;
}

int encode_w_16x2_map(CUtensorMap* m, const void* planes, int rows, int K, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled not available"); return -3; }
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)rows * K * 2};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)planes,      /* 16-bit payload: bf16 or fp16 */ dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(16-bit weight planes) failed: " + std::to_string((int)r)); return -3; }
  return 0;
}

static int encode_act_map(CUtensorMap* m, const float* x, int N, int H, int W, int C, int kc, int tw, int th, int tn) {
  return encode_nhwc_map(m, x, N, H, W, C, kc, tw, th, tn, false, kc == 16);
}

static int encode_w_map(CUtensorMap* m, const float* w, int Cout, int K, int bn, int kc) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled not available"); return -3; }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
  cuuint64_t strides[1] = {(cuuint64_t)K * 4};
  cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)bn};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  kc == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights) failed: " + std::to_string((int)r)); return -3; }
  return 0;
}

// choose the 128-pixel patch (tw, th, tn) that wastes the fewest MMA rows
void choose_patch(int N, int H, int W, int* tw, int* th, int* tn) {
  static const int cand[][3] = {{16, 8, 1}, {8, 16, 1}, {32, 4, 1}, {4, 32, 1}, {64, 2, 1}, {128, 1, 1}, {16, 4, 2},
                                {8, 8, 2},  {4, 16, 2}, {16, 2, 4}, {8, 4, 4},  {4, 8, 4},  {4, 4, 8},   {8, 2, 8},
                                {2, 8, 8},  {2, 2, 32}, {4, 2, 16}, {2, 4, 16}, {1, 1, 128}, {2, 1, 64}, {1, 2, 64}};
  double best = -1;
  for (auto& c : cand) {
    int64_t tiles = (int64_t)ceil_div(W, c[0]) * ceil_div(H, c[1]) * ceil_div(N, c[2]);
    double util = (double)N * H * W / ((double)tiles * 128.0);
    if (util > best + 1e-9) { best = util; *tw = c[0]; *th = c[1]; *tn = c[2]; }
  }
}

template <int BN, int NPROD, int KC>
static int launch_umma(const pvg_conv_desc* d, const float* x, const float* x_lo, const float* w, const float* w_lo,
                       const float* bias, float* y, cudaStream_t st) {
  using C = Cfg<BN, NPROD, KC>;
  ConvParams p;
  // Cin that is not a multiple of 32 (the 16-channel encoder block): the K loop runs over the padded count, the TMA box
  // {32 ch, ...} reads past the tensor's channel extent and gets zeros (out-of-bounds fill) - no padded copy in HBM
  const int CinK = (d->Cin + 31) & ~31;
  p.N = d->N; p.H = d->H; p.W = d->W; p.Cin = CinK; p.Cout = d->Cout; p.R = d->R; p.S = d->S; p.pad = d->pad;
  p.act = d->act; p.slope = d->slope; p.bias = bias; p.y = y; p.corr_fp16 = d->corr_fmt != PVG_CORR_BF16;
  choose_patch(d->N, d->H, d->W, &p.tw, &p.th, &p.tn);
  p.tiles_w = ceil_div(d->W, p.tw); p.tiles_h = ceil_div(d->H, p.th); p.tiles_n = ceil_div(d->N, p.tn);
  CUtensorMap tmA, tmAlo, tmB, tmBlo;
  const int K = d->R * d->S * CinK;
  int rc;
  if (NPROD != 4) {
    if ((rc = encode_act_map(&tmA, x, d->N, d->H, d->W, d->Cin, KC, p.tw, p.th, p.tn))) return rc;
    if ((rc = encode_w_map(&tmB, w, d->Cout, K, BN, KC))) return rc;
  }
  if (NPROD == 3) {
    if ((rc = encode_act_map(&tmAlo, x_lo, d->N, d->H, d->W, d->Cin, KC, p.tw, p.th, p.tn))) return rc;
    if ((rc = encode_w_map(&tmBlo, w_lo, d->Cout, K, BN, KC))) return rc;
  } else if (NPROD == 2 || NPROD == 4) {
    if ((rc = encode_nhwc_16x2_map(&tmAlo, x_lo, d->N, d->H, d->W, d->Cin, p.tw, p.th, p.tn))) return rc;
    if ((rc = encode_w_16x2_map(&tmBlo, w_lo, d->Cout, K, BN))) return rc;
    if (NPROD == 4) { tmA = tmAlo; tmB = tmBlo; }      // the fp32 tensors are not read
  } else {
    tmAlo = tmA; tmBlo = tmB;
  }
  static bool attr_set = false;
  if (!attr_set) {
    PVG_CUDA_OK(cudaFuncSetAttribute(conv_umma_kernel<BN, NPROD, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_set = true;
  }
  dim3 grid((unsigned)(p.tiles_w * p.tiles_h * p.tiles_n * ceil_div(d->Cout, BN)));
  conv_umma_kernel<BN, NPROD, KC><<<grid, kThreads, C::kSmemBytes, st>>>(tmA, tmAlo, tmB, tmBlo, p);
  PVG_LAUNCH_OK();
  return 0;
}

template <int BN, int NPROD>
static int launch_umma_persistent(const pvg_conv_desc* d, const float* x, const float* x_lo, const float* w, const float* w_lo,
                                  const float* bias, float* y, cudaStream_t st) {
  using C = Cfg<BN, NPROD, 32>;
  ConvParams p;
  const int CinK = (d->Cin + 31) & ~31;
  p.N = d->N; p.H = d->H; p.W = d->W; p.Cin = CinK; p.Cout = d->Cout; p.R = d->R; p.S = d->S; p.pad = d->pad;
  p.act = d->act; p.slope = d->slope; p.bias = bias; p.y = y; p.corr_fp16 = d->corr_fmt != PVG_CORR_BF16;
  choose_patch(d->N, d->H, d->W, &p.tw, &p.th, &p.tn);
  p.tiles_w = ceil_div(d->W, p.tw); p.tiles_h = ceil_div(d->H, p.th); p.tiles_n = ceil_div(d->N, p.tn);
  CUtensorMap tmA, tmAlo, tmB, tmBlo;
  const int K = d->R * d->S * CinK;
  int rc;
  if ((rc = encode_act_map(&tmA, x, d->N, d->H, d->W, d->Cin, 32, p.tw, p.th, p.tn))) return rc;
  if ((rc = encode_w_map(&tmB, w, d->Cout, K, BN, 32))) return rc;
  if (NPROD == 3) {
    if ((rc = encode_act_map(&tmAlo, x_lo, d->N, d->H, d->W, d->Cin, 32, p.tw, p.th, p.tn))) return rc;
    if ((rc = encode_w_map(&tmBlo, w_lo, d->Cout, K, BN, 32))) return rc;
  } else {
    if ((rc = encode_nhwc_16x2_map(&tmAlo, x_lo, d->N, d->H, d->W, d->Cin, p.tw, p.th, p.tn))) return rc;
    if ((rc = encode_w_16x2_map(&tmBlo, w_lo, d->Cout, K, BN))) return rc;
  }
  constexpr int kSmem = C::kSmemBytes + 64;            // + the cfree barrier
  static bool attr_set = false;
  if (!attr_set) {
    PVG_CUDA_OK(cudaFuncSetAttribute(conv_umma_persistent_kernel<BN, NPROD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_set = true;
  }
  const int total = p.tiles_w * p.tiles_h * p.tiles_n * ceil_div(d->Cout, BN);
  dim3 grid((unsigned)(total < kSMs ? total : kSMs));
  conv_umma_persistent_kernel<BN, NPROD><<<grid, kThreads, kSmem, st>>>(tmA, tmAlo, tmB, tmBlo, p, total);
  PVG_LAUNCH_OK();
  return 0;
}

template <int NPROD>
static int launch_umma2_persistent(const pvg_conv_desc* d, const float* x, const float* x_lo, const float* w, const float* w_lo,
                                   const float* bias, float* y, cudaStream_t st) {
  using C = Cfg2<NPROD>;
  ConvParams p;
  const int CinK = (d->Cin + 31) & ~31;
  p.N = d->N; p.H = d->H; p.W = d->W; p.Cin = CinK; p.Cout = d->Cout; p.R = d->R; p.S = d->S; p.pad = d->pad;
  p.act = d->act; p.slope = d->slope; p.bias = bias; p.y = y; p.corr_fp16 = d->corr_fmt != PVG_CORR_BF16;
  choose_patch(d->N, d->H, d->W, &p.tw, &p.th, &p.tn);
  p.tiles_w = ceil_div(d->W, p.tw); p.tiles_h = ceil_div(d->H, p.th); p.tiles_n = ceil_div(d->N, p.tn);
  CUtensorMap tmA, tmAlo, tmB, tmBlo;
  const int K = d->R * d->S * CinK;
  int rc;
  if ((rc = encode_act_map(&tmA, x, d->N, d->H, d->W, d->Cin, 32, p.tw, p.th, p.tn))) return rc;
  if ((rc = encode_w_map(&tmB, w, d->Cout, K, C::BN / 2, 32))) return rc;
  if (NPROD == 3) {
    if ((rc = encode_act_map(&tmAlo, x_lo, d->N, d->H, d->W, d->Cin, 32, p.tw, p.th, p.tn))) return rc;
    if ((rc = encode_w_map(&tmBlo, w_lo, d->Cout, K, C::BN / 2, 32))) return rc;
  } else {
    if ((rc = encode_nhwc_16x2_map(&tmAlo, x_lo, d->N, d->H, d->W, d->Cin, p.tw, p.th, p.tn))) return rc;
    if ((rc = encode_w_16x2_map(&tmBlo, w_lo, d->Cout, K, C::BN / 2))) return rc;
  }
  constexpr int kSmem = C::kSmemBytes + 64;
  static bool attr_set = false;
  if (!attr_set) {
    PVG_CUDA_OK(cudaFuncSetAttribute(conv_umma2_persistent_kernel<NPROD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_set = true;
  }
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int items = ceil_div(m_tiles, 2) * ceil_div(d->Cout, C::BN);        // (pair of M tiles) x N tile
  const int clusters = items < kSMs / 2 ? items : kSMs / 2;
  conv_umma2_persistent_kernel<NPROD><<<dim3((unsigned)(clusters * 2)), kThreads, kSmem, st>>>(tmA, tmAlo, tmB, tmBlo, p, items);
  PVG_LAUNCH_OK();
  return 0;
}

// Persistent tile loop (validated on B200, round 2: 120 conv parity cases; fixed cost per tile 3.7 -> 1.3 us for the 1-CTA
// kernel and 5.3 -> 1.3 us for the CTA pair, same time per k-iteration): the default for the split-product kernels.
// PVG_PERSISTENT=0 selects the one-tile-per-CTA kernels (A/B knob).
static bool want_persistent(const pvg_conv_desc* d) {
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("PVG_PERSISTENT");
    env = (e && atoi(e) == 0) ? 0 : 1;
  }
  return d->algo == PVG_ALGO_UMMA_PERSISTENT || env == 1;
}

template <int NPROD>
static int launch_umma2(const pvg_conv_desc* d, const float* x, const float* x_lo, const float* w, const float* w_lo,
                        const float* bias, float* y, cudaStream_t st) {
  using C = Cfg2<NPROD>;
  ConvParams p;
  // Cin that is not a multiple of 32 (the 16-channel encoder block): the K loop runs over the padded count, the TMA box
  // {32 ch, ...} reads past the tensor's channel extent and gets zeros (out-of-bounds fill) - no padded copy in HBM
  const int CinK = (d->Cin + 31) & ~31;
  p.N = d->N; p.H = d->H; p.W = d->W; p.Cin = CinK; p.Cout = d->Cout; p.R = d->R; p.S = d->S; p.pad = d->pad;
  p.act = d->act; p.slope = d->slope; p.bias = bias; p.y = y; p.corr_fp16 = d->corr_fmt != PVG_CORR_BF16;
  choose_patch(d->N, d->H, d->W, &p.tw, &p.th, &p.tn);
  p.tiles_w = ceil_div(d->W, p.tw); p.tiles_h = ceil_div(d->H, p.th); p.tiles_n = ceil_div(d->N, p.tn);
  CUtensorMap tmA, tmAlo, tmB, tmBlo;
  const int K = d->R * d->S * CinK;
  int rc;
  if ((rc = encode_act_map(&tmA, x, d->N, d->H, d->W, d->Cin, 32, p.tw, p.th, p.tn))) return rc;
  if ((rc = encode_w_map(&tmB, w, d->Cout, K, C::BN / 2, 32))) return rc;
  if (NPROD == 3) {
    if ((rc = encode_act_map(&tmAlo, x_lo, d->N, d->H, d->W, d->Cin, 32, p.tw, p.th, p.tn))) return rc;
    if ((rc = encode_w_map(&tmBlo, w_lo, d->Cout, K, C::BN / 2, 32))) return rc;
  } else if (NPROD == 2) {
    if ((rc = encode_nhwc_16x2_map(&tmAlo, x_lo, d->N, d->H, d->W, d->Cin, p.tw, p.th, p.tn))) return rc;
    if ((rc = encode_w_16x2_map(&tmBlo, w_lo, d->Cout, K, C::BN / 2))) return rc;
  } else {
    tmAlo = tmA; tmBlo = tmB;
  }
  static bool attr_set = false;
  if (!attr_set) {
    PVG_CUDA_OK(cudaFuncSetAttribute(conv_umma2_kernel<NPROD>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_set = true;
  }
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int pairs = ceil_div(m_tiles, 2);
  dim3 grid((unsigned)(pairs * ceil_div(d->Cout, C::BN) * 2));
  conv_umma2_kernel<NPROD><<<grid, kThreads, C::kSmemBytes, st>>>(tmA, tmAlo, tmB, tmBlo, p);
  PVG_LAUNCH_OK();
  return 0;
}

// The CTA-pair kernel pays two cluster barriers and a pair-wide TMEM allocation per tile (fixed cost 5.3 us per tile against
// 3.7 us) and runs a k-iteration in 952 clk against 1149 (tools/tile_model.py on B200, 16-bit corrections): it wins from
// 18 k-iterations on, given enough M tiles to fill the machine twice.  PVG_2CTA=1 / 0 forces it on / off.
static bool use_pairs(const pvg_conv_desc* d) {
  static int forced = -2;
  if (forced == -2) {
    const char* e = getenv("PVG_2CTA");
    forced = e ? atoi(e) : -1;
  }
  if (forced >= 0) return forced == 1;
  const int k_iters = d->R * d->S * ((d->Cin + 31) / 32);
  const int64_t tiles = (((int64_t)d->N * d->H * d->W + 127) / 128) * ((d->Cout + 127) / 128);     // CTAs of the 1-CTA kernel
  return k_iters >= 18 && tiles >= 2 * kSMs;
}

// PVG_KC=16 selects the 16-channel (SWIZZLE_64B) stage for the 3xTF32 128-wide tiles (A/B experiment knob)
static int stage_channels() {
  static int kc = 0;
  if (!kc) {
    const char* e = getenv("PVG_KC");
    kc = (e && atoi(e) == 16) ? 16 : 32;
  }
  return kc;
}

template <int NPROD>
static int dispatch_bn(const pvg_conv_desc* d, const float* x, const float* x_lo, const float* w, const float* w_lo,
                       const float* bias, float* y, cudaStream_t st) {
  const int co = d->Cout;
  if constexpr (NPROD == 4) {        // all-fp16 split product: 1-CTA kernel (experiment; superseded by conv_h3.cu)
    if (co <= 16) return launch_umma<16, 4, 32>(d, x, x_lo, w, w_lo, bias, y, st);
    if (co <= 32) return launch_umma<32, 4, 32>(d, x, x_lo, w, w_lo, bias, y, st);
    if (co <= 64) return launch_umma<64, 4, 32>(d, x, x_lo, w, w_lo, bias, y, st);
    if (co <= 80) return launch_umma<80, 4, 32>(d, x, x_lo, w, w_lo, bias, y, st);
    return launch_umma<128, 4, 32>(d, x, x_lo, w, w_lo, bias, y, st);
  } else {
  if constexpr (NPROD >= 2) {
    if (want_persistent(d) && co > 32 && co <= 64) return launch_umma_persistent<64, NPROD>(d, x, x_lo, w, w_lo, bias, y, st);
    if (want_persistent(d) && co > 80) {
      if (use_pairs(d)) return launch_umma2_persistent<NPROD>(d, x, x_lo, w, w_lo, bias, y, st);
      return launch_umma_persistent<128, NPROD>(d, x, x_lo, w, w_lo, bias, y, st);
    }
  }
  if (co <= 16) return launch_umma<16, NPROD, 32>(d, x, x_lo, w, w_lo, bias, y, st);
  if (co <= 32) return launch_umma<32, NPROD, 32>(d, x, x_lo, w, w_lo, bias, y, st);
  if (co <= 64) {
    if constexpr (NPROD == 3) {
      if (stage_channels() == 16) return launch_umma<64, NPROD, 16>(d, x, x_lo, w, w_lo, bias, y, st);
    }
    return launch_umma<64, NPROD, 32>(d, x, x_lo, w, w_lo, bias, y, st);
  }
  if (co <= 80) return launch_umma<80, NPROD, 32>(d, x, x_lo, w, w_lo, bias, y, st);
  if constexpr (NPROD == 1) {
    if (co % 256 == 0 && !use_pairs(d)) return launch_umma<256, NPROD, 32>(d, x, x_lo, w, w_lo, bias, y, st);
  }
  if (use_pairs(d)) return launch_umma2<NPROD>(d, x, x_lo, w, w_lo, bias, y, st);
  if constexpr (NPROD == 3) {
    if (stage_channels() == 16) return launch_umma<128, NPROD, 16>(d, x, x_lo, w, w_lo, bias, y, st);
  }
  return launch_umma<128, NPROD, 32>(d, x, x_lo, w, w_lo, bias, y, st);
  }
}

}  // namespace pvg

using namespace pvg;

extern "C" int pvg_conv2d_fwd(const pvg_conv_desc* d, const float* x, const void* x_lo_, const float* w, const void* w_lo_,
                              const float* bias, float* y, void* stream) {
  const float* x_lo = (const float*)x_lo_;      // fp32 residual plane (nprod == 3) or bf16 plane pair (nprod == 2)
  const float* w_lo = (const float*)w_lo_;
  const bool h3 = d && d->nprod == 2 && d->corr_fmt == PVG_CORR_FP16_ALL;      // reads the fp16 plane pairs only
  PVG_CHECK_ARG(d && y && (h3 || (x && w)), "null argument");
  PVG_CHECK_ARG(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "empty problem");
  PVG_CHECK_ARG(d->R == d->S && d->pad == (d->R - 1) / 2 && (d->R & 1), "only odd 'same' kernels are supported");
  cudaStream_t st = (cudaStream_t)stream;
  int algo = d->algo;
  // channel strides must be multiples of 16 bytes for TMA: Cin % 4 == 0 (fp32), % 8 == 0 with 16-bit planes
  const bool umma_ok = (d->Cin % (d->nprod == 2 ? 8 : 4)) == 0 && (((uintptr_t)x | (uintptr_t)w) & 15) == 0;
  if (algo == PVG_ALGO_AUTO) algo = (umma_ok && pvg_has_umma()) ? PVG_ALGO_UMMA : PVG_ALGO_SIMT;
  if (algo == PVG_ALGO_UMMA_PERSISTENT) PVG_CHECK_ARG(d->nprod >= 2, "the persistent kernel implements the split product only (nprod 2 or 3)");
  if (algo == PVG_ALGO_SIMT) return conv2d_fwd_simt(d, x, w, bias, y, st);
  PVG_CHECK_ARG(umma_ok, "tensor-core path needs Cin % 4 == 0 (% 8 with 16-bit correction planes) and 16-byte aligned operands");
  if (d->nprod == 3) {
    PVG_CHECK_ARG(x_lo && w_lo, "nprod == 3 needs x_lo and w_lo");
    return dispatch_bn<3>(d, x, x_lo, w, w_lo, bias, y, st);
  }
  if (d->nprod == 2) {      // x_lo / w_lo: bf16 plane pairs [2][numel] (pvg_split_16 / pvg_pack_16x2)
    PVG_CHECK_ARG(x_lo && w_lo, "nprod == 2 needs the bf16 plane pairs of x and w");
    PVG_CHECK_ARG((((uintptr_t)x_lo | (uintptr_t)w_lo) & 15) == 0, "bf16 planes must be 16-byte aligned");
    if (h3) {        // all-fp16 split product: conv_h3.cu (PVG_H3_LEGACY=1: the one-tile-per-CTA kernel of this file, A/B knob)
      static const bool legacy = getenv("PVG_H3_LEGACY") && atoi(getenv("PVG_H3_LEGACY")) == 1;
      if (!legacy) return conv2d_fwd_h3(d, x_lo, w_lo, bias, y, nullptr, nullptr, st);
      return dispatch_bn<4>(d, x, x_lo, w, w_lo, bias, y, st);
    }
    return dispatch_bn<2>(d, x, x_lo, w, w_lo, bias, y, st);
  }
  return dispatch_bn<1>(d, x, nullptr, w, nullptr, bias, y, st);
}
