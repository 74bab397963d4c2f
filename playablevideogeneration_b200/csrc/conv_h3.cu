// Forward implicit-GEMM convolution, all-fp16 split product ("h3"), persistent, with halo reuse of the activation tile and an
// optional CTA-pair (cta_group::2) form.  sm_100a: tcgen05 + TMEM + TMA.
//
//   arithmetic   x and w are consumed as fp16 plane pairs { f16((v - f16(v)) * 2^12), f16(v) } (PVG_CORR_FP16_ALL; 22 bits of
//                v).  y = x_hi*w_hi + (x_lo*w_hi + x_hi*w_lo): three kind::f16 MMAs (M = 128 | 256, N = BN, K = 16) with exact
//                products and fp32 accumulation in TMEM - the same arithmetic as the TF32-main / fp16-correction split of
//                conv_umma.cu at 3/4 of its tensor time and half of its shared-memory bytes (no fp32 tiles at all), which is what
//                bounded that kernel (1168 -> 681 clk per k-iteration before halo reuse, B200).  The main term is drained
//                from TMEM into fp32 registers every kDrain k-iterations (the tensor core accumulates with truncation, see
//                conv_umma.cu "split accum"); the correction term (2^12 x its value) accumulates in TMEM for the whole tile.
//   halo reuse   HALO (3x3 convs, tile = 8 wide x 16 high pixels of one image): ONE TMA box {32 ch, 10, 18} per 32-channel chunk
//                is shared by the 9 taps - tap (r, s) is the same shared-memory tile addressed from row r*10 + s with a stride
//                of 10 rows (640 B) between the 8-row groups of the K-major operand.  tcgen05 applies the 64B/128B swizzle to
//                the absolute shared-memory address (tools/probes/umma_desc_probe.cu, measured on B200: descriptor start
//                addresses offset by 1..3 rows and group strides of 10 / 18 rows read exactly the intended rows), which is what
//                makes the shifted views legal.  Activation fills drop from 16 KB to 2.6 KB per k-iteration.
//   rings        A ring (haloed tile per chunk: 3 slots x 24 KB; otherwise one 16 KB patch per k-iteration) and B ring
//                (weights of one (tap, chunk): [lo | hi] x BN rows x 64 B) with their own full/empty mbarriers; the producer
//                fetches the haloed tile of the NEXT chunk (or of the next tile's first chunk) before the 9 weight tiles of the
//                current one.
//   pair         PAIR: two CTAs of a cluster compute 256 pixels x BN channels with one MMA stream (cta_group::2): each CTA
//                stages its own pixel tile and half of the weight rows, so weight fills and weight operand reads halve again.
//   persistent   one CTA (pair) per SM walks the tiles; the epilogue of tile i (bias, activation, NHWC stores, optional fp16
//                planes of y for the next convolution) overlaps the main loop of tile i+1.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "umma.cuh"

namespace pvg {

constexpr int kH3Budget = 200 * 1024;       // bytes of shared memory for the two rings
constexpr int kHaloW = 10, kHaloH = 18;     // haloed 8 x 16 tile

template <int BN, bool HALO, bool PAIR>
struct H3Cfg {
  static constexpr int kBRows = PAIR ? BN / 2 : BN;                 // weight rows staged by this CTA
  static constexpr int kAPlane = HALO ? 12 * 1024 : 8 * 1024;       // one fp16 plane tile of A (halo: 180 rows x 64 B, padded to 1 KB)
  static constexpr int kABytes = 2 * kAPlane;
  static constexpr int kATx = HALO ? 2 * kHaloW * kHaloH * 64 : 2 * 8192;      // bytes TMA writes per A slot
  static constexpr int kBBytes = 2 * kBRows * 64;
  static constexpr int kASlots = HALO ? 3 : ((kH3Budget / (kABytes + kBBytes)) > 8 ? 8 : (kH3Budget / (kABytes + kBBytes)));
  static constexpr int kBSlotsRaw = HALO ? (kH3Budget - kASlots * kABytes) / kBBytes : kASlots;
  static constexpr int kBSlots = kBSlotsRaw > 12 ? 12 : kBSlotsRaw;
  static constexpr int kDrain = 8;                                  // k-iterations (16 main MMAs) per TMEM accumulation chain
  static constexpr int kAccCols = 3 * BN;                           // [main0 | main1 | correction]
  static constexpr int kTmemCols = kAccCols <= 32 ? 32 : (kAccCols <= 64 ? 64 : (kAccCols <= 128 ? 128 : (kAccCols <= 256 ? 256 : 512)));
  static constexpr int kBarBytes = (2 * kASlots + 2 * kBSlots + 5) * 8 + 16;
  static constexpr int kStatBytes = 2 * 4 * BN * 2 * 4;            // BatchNorm partial sums: [tile parity][4 lane quarters][BN][sum, sum^2]
  // Epilogue warps: four (one per TMEM lane quarter) or, from 64 columns on, eight - two per quarter, each taking about half of
  // the columns.  The epilogue drains the accumulator every period AND stores the finished tile: with four warps the store phase of
  // a 128-column tile with planes outlasted two periods of the next tile's main loop and stalled the MMA warp (r02 layer table:
  // planes cost 15 %, bias + ReLU 8 % on K = 2304 layers; layers with K <= 1152 ran at the epilogue's pace).
  static constexpr int kEpiWarps = BN >= 64 ? 8 : 4;
  static constexpr int kEpiThreads = 32 * kEpiWarps;
  static constexpr int kThreads = 64 + kEpiThreads;
  static constexpr int kSplitCol = kEpiWarps == 8 ? ((BN / 16 + 1) / 2) * 16 : BN;     // columns [0, kSplitCol) | [kSplitCol, BN)
  static constexpr int kHalfCols = kSplitCol;                       // the wider half: size of the register accumulator
  static constexpr int kSmemBytes = kASlots * kABytes + kBSlots * kBBytes + 1024 + kBarBytes + kStatBytes + 16;
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 128, "invalid N tile");
  static_assert(!PAIR || BN % 32 == 0, "a pair splits the weight rows in two MMA-legal halves");
  static_assert(kBSlots >= 3 && kASlots >= 2, "rings too shallow");
  static_assert((kBRows * 64) % 1024 == 0, "plane tiles of B must stay 1 KB aligned");
};

struct H3Tile { int w0, h0, n0, co0; };

template <bool PAIR>
__device__ __forceinline__ H3Tile h3_decode(int item, int rank, int n_tiles, int bn, const ConvParams& p) {
  H3Tile t;
  t.co0 = (item % n_tiles) * bn;
  int m = item / n_tiles;
  if (PAIR) m = m * 2 + rank;                 // this CTA's 128-pixel tile (may lie past the end: everything out of bounds)
  t.w0 = (m % p.tiles_w) * p.tw; m /= p.tiles_w;
  t.h0 = (m % p.tiles_h) * p.th; m /= p.tiles_h;
  t.n0 = m * p.tn;
  return t;
}

// K-major fp16 operand, 64-byte rows (32 channels), SWIZZLE_64B; `sbo` = bytes between consecutive 8-row groups
__device__ __forceinline__ uint64_t h3_desc(uint32_t smem_addr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;                      // SWIZZLE_64B
  return d;
}

__device__ __forceinline__ uint32_t h3_pack2(float a, float b) {
  return pack_f16x2_sat(a, b);
}

// fp16 plane pair of 16 finished output values (the PVG_CORR_FP16_ALL operand format of the next convolution)
__device__ __forceinline__ void store_planes16(const float (&v)[16], uint16_t* __restrict__ lo_row, uint16_t* __restrict__ hi_row,
                                               int co, int Cout) {
  if (co + 16 > Cout) {                        // ragged tail of the last N tile (e.g. 72 channels)
#pragma unroll
    for (int j = 0; j < 16; ++j) {             // static indices: a run-time trip count would push v[] into local memory
      if (j >= Cout - co) break;
      const uint32_t hb = pack_f16x2_sat(v[j], 0.f);
      const __half h = __ushort_as_half((unsigned short)(hb & 0xffffu));
      const uint32_t lb = pack_f16x2_sat((v[j] - __half2float(h)) * 4096.f, 0.f);
      const __half l = __ushort_as_half((unsigned short)(lb & 0xffffu));
      hi_row[co + j] = *reinterpret_cast<const uint16_t*>(&h);
      lo_row[co + j] = *reinterpret_cast<const uint16_t*>(&l);
    }
    return;
  }
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    hi[j] = h3_pack2(v[2 * j], v[2 * j + 1]);
    const __half2 h2 = *reinterpret_cast<const __half2*>(&hi[j]);
    lo[j] = h3_pack2((v[2 * j] - __low2float(h2)) * 4096.f, (v[2 * j + 1] - __high2float(h2)) * 4096.f);
  }
  uint4* ph = reinterpret_cast<uint4*>(hi_row + co);
  uint4* pl = reinterpret_cast<uint4*>(lo_row + co);
  ph[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]); ph[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
  pl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]); pl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
}

// ConvLSTM cell point-wise part fused into the gate convolution's epilogue (convolutional_lstm_cell.py:92-101).  The four gate
// convolutions are ONE GEMM whose output channels are interleaved: column 4 * c + {0, 1, 2, 3} = input / forget / output / cell
// gate of hidden channel c, so a thread that holds 16 consecutive accumulator columns of its pixel holds all four gates of 4
// hidden channels: bias, sigmoid / tanh, c' = f * c + i * g and h' = o * tanh(c') happen in registers.  The ACTIVATED gates are
// stored (training: the backward pass needs them; pass gates == NULL for inference).
__device__ __forceinline__ void lstm_finish16(const float (&v)[16], const float* __restrict__ bias, int co, int Cout, bool valid,
                                              const float* __restrict__ c_prev, float* __restrict__ c_new, float* __restrict__ h_new,
                                              float* __restrict__ gates, int64_t pix) {
  if (!valid || co >= Cout) return;
  const int C = Cout >> 2, ch0 = co >> 2;
  float4 cp = ldg4(c_prev + pix * C + ch0);
  const float cpv[4] = {cp.x, cp.y, cp.z, cp.w};
  float cn[4], hn[4], ga[16];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float4 b = bias != nullptr ? ldg4(bias + co + 4 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float gi = 1.f / (1.f + expf(-(v[4 * k] + b.x))), gf = 1.f / (1.f + expf(-(v[4 * k + 1] + b.y)));
    const float go = 1.f / (1.f + expf(-(v[4 * k + 2] + b.z))), gc = tanhf(v[4 * k + 3] + b.w);
    cn[k] = gf * cpv[k] + gi * gc;
    hn[k] = go * tanhf(cn[k]);
    ga[4 * k] = gi; ga[4 * k + 1] = gf; ga[4 * k + 2] = go; ga[4 * k + 3] = gc;
  }
  stg4(c_new + pix * C + ch0, make_float4(cn[0], cn[1], cn[2], cn[3]));
  stg4(h_new + pix * C + ch0, make_float4(hn[0], hn[1], hn[2], hn[3]));
  if (gates != nullptr) {
#pragma unroll
    for (int k = 0; k < 4; ++k) stg4(gates + pix * Cout + co + 4 * k, make_float4(ga[4 * k], ga[4 * k + 1], ga[4 * k + 2], ga[4 * k + 3]));
  }
}

// Sum of 16 per-lane values over the 32 lanes of a warp with 16 shuffles (instead of 80): every step exchanges half of the
// values a lane still holds with the lane `offset` away and keeps the other half, so after offsets 16, 8, 4, 2 each lane holds
// ONE channel's partial sum, completed by the offset-1 step.  Returns the full warp sum of channel
// ch = 8 * bit4 + 4 * bit3 + 2 * bit2 + bit1 of the lane index (lanes L and L ^ 1 return the same channel).
__device__ __forceinline__ float warp_sum16(const float (&v)[16], int lane) {
  float a[8], b[4], c[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float send = h16 ? v[j] : v[j + 8];
    const float keep = h16 ? v[j + 8] : v[j];
    a[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float send = h8 ? a[j] : a[j + 4];
    const float keep = h8 ? a[j + 4] : a[j];
    b[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float send = h4 ? b[j] : b[j + 2];
    const float keep = h4 ? b[j + 2] : b[j];
    c[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float send = h2 ? c[0] : c[1];
  const float keep = h2 ? c[1] : c[0];
  float d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  return d;
}

template <int BN, bool HALO, bool PAIR>
__global__ void __launch_bounds__(H3Cfg<BN, HALO, PAIR>::kThreads, 1)
conv_h3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvParams p,
               const int total_items) {
  using C = H3Cfg<BN, HALO, PAIR>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_base = smem;
  uint8_t* b_base = smem + C::kASlots * C::kABytes;
  uint64_t* a_full = (uint64_t*)(b_base + C::kBSlots * C::kBBytes);
  uint64_t* a_empty = a_full + C::kASlots;
  uint64_t* b_full = a_empty + C::kASlots;
  uint64_t* b_empty = b_full + C::kBSlots;
  uint64_t* tfull_bar = b_empty + C::kBSlots;       // [2] main accumulator buffer ready for the epilogue
  uint64_t* tempty_bar = tfull_bar + 2;             // [2] main accumulator buffer drained
  uint64_t* cfree_bar = tempty_bar + 2;             // [1] correction accumulator folded into registers: next tile may overwrite it
  uint32_t* tmem_slot = (uint32_t*)(cfree_bar + 1);
  float* stat_part = (float*)(((uintptr_t)(tmem_slot + 1) + 15) & ~(uintptr_t)15);     // [2][4][BN][2]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const bool leader = rank == 0;
  const int first_item = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int item_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int n_tiles = (p.Cout + BN - 1) / BN;
  const int chunks = p.Cin / 32;
  const int taps = p.R * p.S;
  const int k_iters = taps * chunks;
  constexpr uint32_t kEpiArrivals = (PAIR ? 2 : 1) * C::kEpiThreads;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB);
    for (int s = 0; s < C::kASlots; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < C::kBSlots; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(&tfull_bar[0], 1); mbar_init(&tfull_bar[1], 1);
    mbar_init(&tempty_bar[0], kEpiArrivals); mbar_init(&tempty_bar[1], kEpiArrivals);
    mbar_init(cfree_bar, kEpiArrivals);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) tmem2_alloc(tmem_slot, C::kTmemCols); else tmem_alloc(tmem_slot, C::kTmemCols);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs of a pair) =====================
    const uint32_t elected = elect_one();
    uint32_t a_slot = 0, a_phase = 0, b_slot = 0, b_phase = 0;
    auto load_a = [&](const H3Tile& t, int chunk, int r, int s) {
      // !HALO: the patch and the weights of a k-iteration travel in lockstep - they share the B ring's slot index and barriers
      const uint32_t slot = HALO ? a_slot : b_slot;
      uint64_t* full = HALO ? &a_full[slot] : &b_full[slot];
      mbar_wait(HALO ? &a_empty[slot] : &b_empty[slot], (HALO ? a_phase : b_phase) ^ 1);
      if (elected) {
        uint8_t* dst = a_base + slot * C::kABytes;
        if (leader) mbar_expect_tx(full, (PAIR ? 2 : 1) * (HALO ? C::kATx : C::kATx + C::kBBytes));      // a pair credits the leader's barrier
        if (HALO) {
          if (PAIR) {
            tma2_load_5d(dst, &tmA, full, chunk * 32, t.w0 - 1, t.h0 - 1, t.n0, 0);
            tma2_load_5d(dst + C::kAPlane, &tmA, full, chunk * 32, t.w0 - 1, t.h0 - 1, t.n0, 1);
          } else {
            tma_load_5d(dst, &tmA, full, chunk * 32, t.w0 - 1, t.h0 - 1, t.n0, 0);
            tma_load_5d(dst + C::kAPlane, &tmA, full, chunk * 32, t.w0 - 1, t.h0 - 1, t.n0, 1);
          }
        } else {                      // both planes of the shifted patch in one box: [f16(lo * 2^12) tile | f16(x) tile]
          if (PAIR) tma2_load_5d(dst, &tmA, full, chunk * 32, t.w0 + s - p.pad, t.h0 + r - p.pad, t.n0, 0);
          else tma_load_5d(dst, &tmA, full, chunk * 32, t.w0 + s - p.pad, t.h0 + r - p.pad, t.n0, 0);
        }
      }
      __syncwarp();
      if (HALO && ++a_slot == C::kASlots) { a_slot = 0; a_phase ^= 1; }
    };
    auto load_b = [&](int co0, int kcoord) {
      if (HALO) mbar_wait(&b_empty[b_slot], b_phase ^ 1);            // !HALO: load_a has waited for (and armed) this slot already
      if (elected) {
        uint8_t* dst = b_base + b_slot * C::kBBytes;
        if (HALO && leader) mbar_expect_tx(&b_full[b_slot], (PAIR ? 2 : 1) * C::kBBytes);
        if (PAIR) tma2_load_3d(dst, &tmB, &b_full[b_slot], kcoord, co0 + rank * C::kBRows, 0);
        else tma_load_3d(dst, &tmB, &b_full[b_slot], kcoord, co0, 0);
      }
      __syncwarp();
      if (++b_slot == C::kBSlots) { b_slot = 0; b_phase ^= 1; }
    };
    if (HALO) {
      if (first_item < total_items) load_a(h3_decode<PAIR>(first_item, rank, n_tiles, BN, p), 0, 0, 0);
      for (int item = first_item; item < total_items; item += item_stride) {
        const H3Tile t = h3_decode<PAIR>(item, rank, n_tiles, BN, p);
        for (int chunk = 0; chunk < chunks; ++chunk) {
          if (chunk + 1 < chunks) load_a(t, chunk + 1, 0, 0);                     // the NEXT haloed tile first
          else if (item + item_stride < total_items) load_a(h3_decode<PAIR>(item + item_stride, rank, n_tiles, BN, p), 0, 0, 0);
          for (int tap = 0; tap < taps; ++tap) load_b(t.co0, tap * p.Cin + chunk * 32);
        }
      }
    } else {
      for (int item = first_item; item < total_items; item += item_stride) {
        const H3Tile t = h3_decode<PAIR>(item, rank, n_tiles, BN, p);
        for (int chunk = 0; chunk < chunks; ++chunk)
          for (int tap = 0; tap < taps; ++tap) {
            const int r = tap / p.S, s = tap - r * p.S;
            load_a(t, chunk, r, s);
            load_b(t.co0, tap * p.Cin + chunk * 32);
          }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (the leader CTA of a pair) =====================
    // One warp-level loop iteration issues the MMAs of up to THREE k-iterations (for a 3x3 conv: one tap row): ncu showed the
    // first version bound by the latency of the issuing warp's own loop (barrier polls, warp reconvergence, uniform-register
    // set-up: ~500-700 clk per iteration against 384 clk of tensor work at N = 128 and 192 clk at N = 64), not by data or
    // by the tensor pipe.  A drain period is one chunk (9 k-iterations) for 3x3 convs, 8 k-iterations otherwise.
    if (leader) {
      const uint32_t elected = elect_one();
      const uint32_t idesc = make_idesc_f16<BN>(true, false, false, PAIR ? 256 : 128);
      const uint64_t dA = h3_desc(smem_u32(a_base), (HALO && !p.dbg) ? kHaloW * 64 : 512);
      const uint64_t dB = h3_desc(smem_u32(b_base), 512);
      constexpr uint32_t kAU = C::kABytes >> 4, kAPlaneU = C::kAPlane >> 4, kBU = C::kBBytes >> 4, kBPlaneU = (C::kBRows * 64) >> 4;
      const uint32_t corr = tmem_acc + 2 * BN;
      const int period = taps == 9 ? 9 : C::kDrain;
      uint32_t a_slot = 0, a_phase = 0, b_slot = 0, b_phase = 0, pg = 0;
      int it = 0;
      for (int item = first_item; item < total_items; item += item_stride, ++it) {
        if (it > 0) {                                   // the epilogue has folded the previous tile's corrections into registers
          mbar_wait(cfree_bar, (uint32_t)(it - 1) & 1);
          tc_fence_after();
        }
        uint32_t corr_acc = 0;
        for (int k = 0; k < k_iters; ++pg) {
          const uint32_t b = pg & 1;
          mbar_wait(&tempty_bar[b], ((pg >> 1) & 1) ^ 1);       // the epilogue finished draining this buffer
          if (HALO) mbar_wait(&a_full[a_slot], a_phase);        // HALO: a period is one chunk = one haloed tile
          const uint32_t main_acc = tmem_acc + b * BN;
          uint32_t main_started = 0;
          const int k_end = min(k + period, k_iters);
          uint32_t tap_off = 0;                         // descriptor offset of the tap row inside the haloed tile (16-byte units)
          while (k < k_end) {
            const int g = min(3, k_end - k);
            uint32_t sl[3], ph = b_phase, nxt = b_slot;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              sl[j] = nxt;
              if (j < g) {
                mbar_wait(&b_full[nxt], ph);
                if (++nxt == C::kBSlots) { nxt = 0; ph ^= 1; }
              }
            }
            tc_fence_after();
            const bool period_done = k + g == k_end;
            if (elected) {
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                if (j < g) {
                  const uint32_t a_lo = HALO ? a_slot * kAU + tap_off + 4 * j : sl[j] * kAU;      // 64-byte rows = 4 units
                  const uint32_t a_x = a_lo + kAPlaneU, b_lo = sl[j] * kBU, b_x = b_lo + kBPlaneU;
                  if (PAIR) {
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) { umma2_bf16(corr, dA + (a_lo + 2 * ks), dB + (b_x + 2 * ks), idesc, corr_acc); corr_acc = 1; }
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) umma2_bf16(corr, dA + (a_x + 2 * ks), dB + (b_lo + 2 * ks), idesc, 1);
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) { umma2_bf16(main_acc, dA + (a_x + 2 * ks), dB + (b_x + 2 * ks), idesc, main_started); main_started = 1; }
                    umma2_commit_both(&b_empty[sl[j]]);
                  } else {
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) { umma_bf16(corr, dA + (a_lo + 2 * ks), dB + (b_x + 2 * ks), idesc, corr_acc); corr_acc = 1; }
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) umma_bf16(corr, dA + (a_x + 2 * ks), dB + (b_lo + 2 * ks), idesc, 1);
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) { umma_bf16(main_acc, dA + (a_x + 2 * ks), dB + (b_x + 2 * ks), idesc, main_started); main_started = 1; }
                    umma_commit(&b_empty[sl[j]]);
                  }
                }
              }
              if (period_done) {                         // the chain (and, at the tile's end, the corrections) is complete
                if (PAIR) { if (HALO) umma2_commit_both(&a_empty[a_slot]); umma2_commit_both(&tfull_bar[b]); }
                else { if (HALO) umma_commit(&a_empty[a_slot]); umma_commit(&tfull_bar[b]); }
              }
            }
            __syncwarp();
            b_slot = nxt; b_phase = ph;
            k += g;
            if (HALO && !p.dbg) tap_off += (uint32_t)kHaloW * 4;     // next row of the 10-wide haloed tile
          }
          if (HALO && ++a_slot == C::kASlots) { a_slot = 0; a_phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..; in a pair each CTA drains its own 128 TMEM lanes) =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int half = (warp - 2) >> 2;       // which column half of the tile this warp owns (0 when there are four epilogue warps)
    const int c_begin = half ? C::kSplitCol : 0;
    const int c_width = half ? BN - C::kSplitCol : C::kSplitCol;
    const int e_idx = half * 128 + row;     // index among the epilogue threads
    const int wi = row % p.tw;
    const int hi = (row / p.tw) % p.th;
    const int ni = row / (p.tw * p.th);
    const bool vec_ok = (p.Cout % 4) == 0 && (((uintptr_t)p.y | (uintptr_t)p.bias) & 15) == 0;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int period = taps == 9 ? 9 : C::kDrain;
    const int periods = (k_iters + period - 1) / period;
    const float out_scale = p.out_scale != nullptr ? __ldg(p.out_scale) : 1.f;      // 1 / S of a scaled gradient operand
    uint32_t pg = 0;
    int tile_it = 0;
    for (int item = first_item; item < total_items; item += item_stride, ++tile_it) {
      const H3Tile t = h3_decode<PAIR>(item, rank, n_tiles, BN, p);
      const int ow = t.w0 + wi, oh = t.h0 + hi, on = t.n0 + ni;
      const bool valid = ow < p.W && oh < p.H && on < p.N;
      const int64_t pix = ((int64_t)on * p.H + oh) * p.W + ow;
      float* yrow = p.y + pix * p.Cout;
      float acc[C::kHalfCols];
      float tile_amax = 0.f;
#pragma unroll
      for (int j = 0; j < C::kHalfCols; ++j) acc[j] = 0.f;
      for (int per = 0; per < periods; ++per, ++pg) {
        const uint32_t b = pg & 1;
        mbar_wait(&tfull_bar[b], (pg >> 1) & 1);
        tc_fence_after();
        // TMEM -> registers 32 columns at a time: two loads in flight per wait (r02 capture of a K = 576 layer: the epilogue
        // warps sat on the long scoreboard - one tcgen05.ld round trip per 16 columns and one bias load per stored chunk)
#pragma unroll
        for (int cc = 0; cc < C::kHalfCols; cc += 32) {
          if (cc >= c_width) break;
          uint32_t r0[16], r1[16];
          const uint32_t ta = tmem_acc + lane_base + (uint32_t)(b * BN + c_begin + cc);
          const bool two = cc + 16 < C::kHalfCols && cc + 16 < c_width;
          tmem_ld16_nowait(ta, r0);
          if (two) tmem_ld16_nowait(ta + 16, r1);
          tmem_ld_wait();
          tmem_ld_use(r0);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[cc + j] += __uint_as_float(r0[j]);
          if (cc + 16 < C::kHalfCols) {
            if (two) {
              tmem_ld_use(r1);
#pragma unroll
              for (int j = 0; j < 16; ++j) acc[cc + 16 + j] += __uint_as_float(r1[j]);
            }
          }
        }
        tc_fence_before();
        if (PAIR) mbar_arrive_leader(&tempty_bar[b]); else mbar_arrive(&tempty_bar[b]);
      }
      // TMEM reads only: fold the corrections (accumulated at 2^12 x their value) into the registers, then hand the
      // correction buffer back to the MMA warp; activation / stores overlap the next tile's main loop.  The bias joins here:
      // its loads overlap the TMEM round trips instead of waiting behind the stores of the previous chunk.
      const bool bias_here = p.bias != nullptr && p.act != PVG_ACT_LSTM;
#pragma unroll
      for (int cc = 0; cc < C::kHalfCols; cc += 32) {
        if (cc >= c_width) break;
        uint32_t r0[16], r1[16];
        const uint32_t ta = tmem_acc + lane_base + (uint32_t)(2 * BN + c_begin + cc);
        const bool two = cc + 16 < C::kHalfCols && cc + 16 < c_width;
        tmem_ld16_nowait(ta, r0);
        if (two) tmem_ld16_nowait(ta + 16, r1);
        float bv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) bv[j] = 0.f;
        if (bias_here) {
          const int co = t.co0 + c_begin + cc;
          if (vec_ok && co + 16 <= p.Cout) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (j >= 16 && !(two && co + 32 <= p.Cout)) break;
              const float4 b4 = ldg4(p.bias + co + j);
              bv[j] = b4.x; bv[j + 1] = b4.y; bv[j + 2] = b4.z; bv[j + 3] = b4.w;
            }
            if (two && co + 32 > p.Cout) {
#pragma unroll
              for (int j = 16; j < 32; ++j)
                if (co + j < p.Cout) bv[j] = __ldg(p.bias + co + j);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (co + j < p.Cout) bv[j] = __ldg(p.bias + co + j);
          }
        }
        tmem_ld_wait();
        tmem_ld_use(r0);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[cc + j] = fmaf(__uint_as_float(r0[j]), 0x1p-12f, acc[cc + j]) * out_scale + bv[j];
        if (cc + 16 < C::kHalfCols) {
          if (two) {
            tmem_ld_use(r1);
#pragma unroll
            for (int j = 0; j < 16; ++j)
              acc[cc + 16 + j] = fmaf(__uint_as_float(r1[j]), 0x1p-12f, acc[cc + 16 + j]) * out_scale + bv[16 + j];
          }
        }
      }
      tc_fence_before();
      if (PAIR) mbar_arrive_leader(cfree_bar); else mbar_arrive(cfree_bar);
#pragma unroll
      for (int cc = 0; cc < C::kHalfCols; cc += 16) {
        if (cc >= c_width) break;
        const int c = c_begin + cc;
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = acc[cc + j];
        if (p.act == PVG_ACT_LSTM) {          // fused ConvLSTM cell: p.y = activated gates (optional), Cout = 4 * hidden channels
          lstm_finish16(v, p.bias, t.co0 + c, p.Cout, valid, p.lstm_c_prev, p.lstm_c_new, p.lstm_h_new, p.y, pix);
          continue;
        }
        finish16(v, nullptr, t.co0 + c, p.Cout, p.act, p.slope, yrow, valid, vec_ok);       // the bias is already in
        if (p.y_planes != nullptr && valid && t.co0 + c < p.Cout) {
          uint16_t* lo_row = p.y_planes + pix * p.Cout;          // Cout % 8 == 0 (checked by the host): 16-byte aligned rows
          store_planes16(v, lo_row, lo_row + p.y_numel, t.co0 + c, p.Cout);
        }
        if (p.amax_out != nullptr) {             // largest magnitude of the tile's outputs (scale of the next backward kernel)
#pragma unroll
          for (int j = 0; j < 16; ++j) tile_amax = fmaxf(tile_amax, valid && t.co0 + c + j < p.Cout ? fabsf(v[j]) : 0.f);
        }
        if (p.bn_sums != nullptr) {
          // BatchNorm statistics of the output in the epilogue (no separate pass over y): per-channel sum and sum of squares
          // of this warp's 32 pixels, 16 shuffles each
          float sq[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { v[j] = valid ? v[j] : 0.f; sq[j] = v[j] * v[j]; }
          const float s1 = warp_sum16(v, lane), s2 = warp_sum16(sq, lane);
          if ((lane & 1) == 0) {
            const int ch = c + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            float* dst = stat_part + (((tile_it & 1) * 4 + q) * BN + ch) * 2;
            dst[0] = s1; dst[1] = s2;
          }
        }
      }
      if (p.amax_out != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tile_amax = fmaxf(tile_amax, __shfl_xor_sync(0xffffffffu, tile_amax, o));
        if (lane == 0 && tile_amax == tile_amax) atomicMax(p.amax_out, __float_as_uint(tile_amax));
      }
      if (p.bn_sums != nullptr) {
        asm volatile("bar.sync 1, %0;" ::"n"(C::kEpiThreads) : "memory");            // the epilogue warps
        if (e_idx < BN && t.co0 + e_idx < p.Cout) {
          const float* src = stat_part + ((tile_it & 1) * 4 * BN + e_idx) * 2;
          const float s1 = src[0] + src[2 * BN] + src[4 * BN] + src[6 * BN];
          const float s2 = src[1] + src[2 * BN + 1] + src[4 * BN + 1] + src[6 * BN + 1];
          const int g = t.n0 / p.bn_samples_per_group;            // tiles never straddle groups (checked by the host)
          if (g < p.bn_groups) {
            double* o = (p.bn_replicas > 0 ? p.bn_ws + (size_t)(blockIdx.x % p.bn_replicas) * p.bn_groups * 2 * p.Cout : p.bn_sums) +
                        (size_t)g * 2 * p.Cout + t.co0 + e_idx;
            atomicAdd(o, (double)s1);
            atomicAdd(o + p.Cout, (double)s2);
          }
        }
        // the partials of tile i + 2 reuse this parity's slots: by then every warp has passed the barrier of tile i + 1
      }
    }
    if (p.bn_sums != nullptr && p.bn_replicas > 0) {
      // last CTA of the launch: replica copies -> bn_sums, and the workspace goes back to all-zero for the next launch
      __threadfence();
      asm volatile("bar.sync 1, %0;" ::"n"(C::kEpiThreads) : "memory");              // this CTA's atomics have all been issued and fenced
      uint32_t* flag = reinterpret_cast<uint32_t*>(stat_part);
      if (e_idx == 0) *flag = atomicAdd(p.bn_ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
      asm volatile("bar.sync 1, %0;" ::"n"(C::kEpiThreads) : "memory");
      if (*flag != 0u) {
        __threadfence();
        const int per = p.bn_groups * 2 * p.Cout;
        for (int i = e_idx; i < per; i += C::kEpiThreads) {
          double acc = 0.0;
          for (int r = 0; r < p.bn_replicas; ++r) {
            double* src = p.bn_ws + (size_t)r * per + i;
            acc += __ldcg(src);
            __stcg(src, 0.0);
          }
          p.bn_sums[i] += acc;
        }
        if (e_idx == 0) *p.bn_ticket = 0u;
      }
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();      // a peer's shared memory / TMEM must outlive every MMA that reads it
  if (warp == 1) {
    if (PAIR) tmem2_dealloc(tmem_acc, C::kTmemCols); else tmem_dealloc(tmem_acc, C::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
// 5-D map over the fp16 plane pair [2][N][H][W][C] with a haloed box {32 ch, 10, 18, 1 image, 1 plane}
static int encode_halo_map(CUtensorMap* m, const void* planes, int N, int H, int W, int C) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled not available"); return -3; }
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, 2};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)N * H * W * C * 2};
  cuuint32_t box[5] = {32, kHaloW, kHaloH, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 /* 16-bit payload */, 5, (void*)planes, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(haloed activation planes) failed: " + std::to_string((int)r)); return -3; }
  return 0;
}

// operands of the fused ConvLSTM epilogue (PVG_ACT_LSTM), set by pvg_convlstm_step around its call of conv2d_fwd_h3
struct LstmIO { const float* c_prev; float* c_new; float* h_new; };
static thread_local LstmIO g_lstm = {nullptr, nullptr, nullptr};

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// BatchNorm statistics request of the current pvg_conv2d_fwd_planes call (see its bn_sums argument)
struct BnStatsOut { double* sums; int groups; };
static thread_local BnStatsOut g_bn = {nullptr, 0};
static thread_local uint32_t* g_amax_out = nullptr;      // amax_out argument of the current pvg_conv2d_fwd_planes call

// Zero-kept workspace of the replicated BatchNorm sums (+ the ticket counter behind it), one per device, allocated at first use
// (an eager call: not inside a stream capture).  Launches that use it must be stream-ordered with respect to each other - every
// caller in this package runs its convolutions on one stream (or one captured graph).
constexpr int kBnReplicasMax = 32;
constexpr size_t kBnWsDoubles = 1u << 20;             // 8 MB: e.g. 32 copies x 16 groups x 2 x 1024 channels
static double* bn_workspace(cudaStream_t st) {
  static double* ws[64] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (ws[dev] == nullptr) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return nullptr; }
    double* pnew = nullptr;
    if (cudaMalloc(&pnew, (kBnWsDoubles + 2) * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cudaMemset(pnew, 0, (kBnWsDoubles + 2) * sizeof(double)) != cudaSuccess) { cudaGetLastError(); cudaFree(pnew); return nullptr; }
    ws[dev] = pnew;
  }
  return ws[dev];
}

template <int BN, bool HALO, bool PAIR>
static int launch_h3(const pvg_conv_desc* d, const void* x_planes, const void* w_planes, const float* bias, float* y,
                     void* y_planes, const float* out_scale, cudaStream_t st) {
  using C = H3Cfg<BN, HALO, PAIR>;
  ConvParams p;
  const int CinK = (d->Cin + 31) & ~31;       // the K loop runs over whole 32-channel chunks: TMA zero-fills past the tensor's extent
  p.N = d->N; p.H = d->H; p.W = d->W; p.Cin = CinK; p.Cout = d->Cout; p.R = d->R; p.S = d->S; p.pad = d->pad;
  p.act = d->act; p.slope = d->slope; p.bias = bias; p.y = y; p.corr_fp16 = 1;
  p.y_planes = (uint16_t*)y_planes; p.y_numel = (int64_t)d->N * d->H * d->W * d->Cout; p.out_scale = out_scale;
  p.lstm_c_prev = g_lstm.c_prev; p.lstm_c_new = g_lstm.c_new; p.lstm_h_new = g_lstm.h_new;
  p.bn_sums = nullptr; p.bn_groups = 0; p.bn_samples_per_group = 1;
  p.amax_out = g_amax_out;
  static const int dbg = env_int("PVG_H3_DBG", 0);      // timing experiment only (wrong results): haloed tile read without row offsets
  p.dbg = dbg;
  if (HALO) { p.tw = 8; p.th = 16; p.tn = 1; }
  else choose_patch(d->N, d->H, d->W, &p.tw, &p.th, &p.tn);
  p.tiles_w = ceil_div(d->W, p.tw); p.tiles_h = ceil_div(d->H, p.th); p.tiles_n = ceil_div(d->N, p.tn);
  bool stats_after = false;             // BatchNorm statistics requested but this tiling cannot produce them: separate pass
  p.bn_ws = nullptr; p.bn_ticket = nullptr; p.bn_replicas = 0;
  if (g_bn.sums != nullptr) {
    const int spg = d->N / g_bn.groups;
    if (spg % p.tn == 0) { p.bn_sums = g_bn.sums; p.bn_groups = g_bn.groups; p.bn_samples_per_group = spg; }   // tiles stay inside a group
    else stats_after = true;
  }
  CUtensorMap tmA, tmB;
  int rc;
  if (HALO) { if ((rc = encode_halo_map(&tmA, x_planes, d->N, d->H, d->W, d->Cin))) return rc; }
  else if ((rc = encode_nhwc_16x2_map(&tmA, x_planes, d->N, d->H, d->W, d->Cin, p.tw, p.th, p.tn))) return rc;
  if ((rc = encode_w_16x2_map(&tmB, w_planes, d->Cout, d->R * d->S * CinK, C::kBRows))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    PVG_CUDA_OK(cudaFuncSetAttribute(conv_h3_kernel<BN, HALO, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_set = true;
  }
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int n_tiles = ceil_div(d->Cout, BN);
  const int items = (PAIR ? ceil_div(m_tiles, 2) : m_tiles) * n_tiles;
  const int max_ctas = PAIR ? kSMs / 2 : kSMs;
  const int groups = items < max_ctas ? items : max_ctas;
  if (p.bn_sums != nullptr) {
    // replica count: enough copies that a copy sees few CTAs, bounded by the workspace; tiny launches keep the direct path
    const int ctas = PAIR ? groups * 2 : groups;
    const size_t per = (size_t)p.bn_groups * 2 * d->Cout;
    int rep = ctas >= 8 ? (ctas < kBnReplicasMax ? ctas : kBnReplicasMax) : 0;
    while (rep > 1 && per * rep > kBnWsDoubles) rep >>= 1;
    static const int no_rep = env_int("PVG_BN_NO_REPLICAS", 0);
    if (rep > 1 && !no_rep) {
      double* ws = bn_workspace(st);
      if (ws != nullptr) { p.bn_ws = ws; p.bn_ticket = reinterpret_cast<uint32_t*>(ws + kBnWsDoubles); p.bn_replicas = rep; }
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(PAIR ? groups * 2 : groups));
  cfg.blockDim = dim3(C::kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  PVG_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_h3_kernel<BN, HALO, PAIR>, tmA, tmB, p, items));
  PVG_LAUNCH_OK();
  if (stats_after) return pvg_bn_stats(y, d->N, d->H * d->W, d->Cout, g_bn.groups, g_bn.sums, (void*)st);
  return 0;
}

template <bool HALO, bool PAIR>
static int dispatch_h3_bn(const pvg_conv_desc* d, const void* xp, const void* wp, const float* bias, float* y, void* yp,
                          const float* os, cudaStream_t st) {
  const int co = d->Cout;
  if constexpr (PAIR) {
    if (co <= 32) return launch_h3<32, HALO, true>(d, xp, wp, bias, y, yp, os, st);
    if (co <= 64) return launch_h3<64, HALO, true>(d, xp, wp, bias, y, yp, os, st);
    return launch_h3<128, HALO, true>(d, xp, wp, bias, y, yp, os, st);
  } else {
    if (co <= 16) return launch_h3<16, HALO, false>(d, xp, wp, bias, y, yp, os, st);
    if (co <= 32) return launch_h3<32, HALO, false>(d, xp, wp, bias, y, yp, os, st);
    if (co <= 64) return launch_h3<64, HALO, false>(d, xp, wp, bias, y, yp, os, st);
    if (co <= 80) return launch_h3<80, HALO, false>(d, xp, wp, bias, y, yp, os, st);
    return launch_h3<128, HALO, false>(d, xp, wp, bias, y, yp, os, st);
  }
}

// all-fp16 forward convolution; called by pvg_conv2d_fwd (nprod == 2, corr_fmt == PVG_CORR_FP16_ALL) and pvg_conv2d_fwd_planes
int conv2d_fwd_h3(const pvg_conv_desc* d, const void* x_planes, const void* w_planes, const float* bias, float* y,
                  void* y_planes, const float* out_scale, cudaStream_t st) {
  static const int force_halo = env_int("PVG_H3_HALO", -1);       // A/B knobs: 0 / 1 force, -1 = heuristic
  static const int force_pair = env_int("PVG_H3_PAIR", -1);
  // halo reuse: 3x3 convs whose maps tile well into 8 x 16 single-image patches
  int tw, th, tn;
  choose_patch(d->N, d->H, d->W, &tw, &th, &tn);
  const double util_best = (double)d->N * d->H * d->W / ((double)ceil_div(d->W, tw) * ceil_div(d->H, th) * ceil_div(d->N, tn) * 128.0);
  const double util_halo = (double)d->H * d->W / ((double)ceil_div(d->W, 8) * ceil_div(d->H, 16) * 128.0);
  bool halo = d->R == 3 && d->S == 3 && d->pad == 1 && util_halo >= 0.8 * util_best;
  if (force_halo >= 0) halo = force_halo == 1 && d->R == 3 && d->S == 3 && d->pad == 1;
  const int64_t m_tiles = halo ? (int64_t)ceil_div(d->W, 8) * ceil_div(d->H, 16) * d->N
                               : (int64_t)ceil_div(d->W, tw) * ceil_div(d->H, th) * ceil_div(d->N, tn);
  // A pair (two CTAs, one MMA stream, half of the weight rows each) runs a k-iteration of its two tiles in ~0.8x the time one
  // CTA needs for one (tools/tile_model.py on B200: 550 vs 698 clk at N = 128, 362 vs 526 at N = 64): take it when its rounds
  // over the 74 clusters cost less than the 1-CTA rounds over the 148 SMs.
  const int64_t n_tiles = ceil_div(d->Cout, d->Cout <= 64 ? 64 : 128);
  const int64_t rounds1 = ceil_div64(m_tiles * n_tiles, kSMs), rounds2 = ceil_div64(ceil_div64(m_tiles, 2) * n_tiles, kSMs / 2);
  bool pair = d->Cout > 16 && rounds2 * 4 <= rounds1 * 5;          // N = 16 would leave 8-row weight tiles (below the 1 KB tile alignment)
  if (force_pair >= 0) pair = force_pair == 1 && d->Cout > 16;
  if (halo) return pair ? dispatch_h3_bn<true, true>(d, x_planes, w_planes, bias, y, y_planes, out_scale, st)
                        : dispatch_h3_bn<true, false>(d, x_planes, w_planes, bias, y, y_planes, out_scale, st);
  return pair ? dispatch_h3_bn<false, true>(d, x_planes, w_planes, bias, y, y_planes, out_scale, st)
              : dispatch_h3_bn<false, false>(d, x_planes, w_planes, bias, y, y_planes, out_scale, st);
}

}  // namespace pvg

using namespace pvg;

// y = act(bias + conv(x, w)) with x and w given ONLY as fp16 plane pairs (pvg_split_16 / pvg_pack_16x2 with PVG_CORR_FP16_ALL,
// or the y_planes of a previous call); y_planes (optional): the plane pair of y for the next convolution.
extern "C" int pvg_conv2d_fwd_planes(const pvg_conv_desc* d, const void* x_planes, const void* w_planes, const float* bias,
                                     float* y, void* y_planes, const float* out_scale, double* bn_sums, int bn_groups,
                                     uint32_t* amax_out, void* stream) {
  PVG_CHECK_ARG(!bn_sums || (bn_groups >= 1 && d && d->N % bn_groups == 0 && d->act == PVG_ACT_NONE),
                "BatchNorm statistics: N must be divisible by the group count and the epilogue must not apply an activation");
  PVG_CHECK_ARG(d && x_planes && w_planes && y, "null argument");
  PVG_CHECK_ARG(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "empty problem");
  PVG_CHECK_ARG(d->R == d->S && d->pad == (d->R - 1) / 2 && (d->R & 1), "only odd 'same' kernels are supported");
  PVG_CHECK_ARG(d->Cin % 8 == 0, "16-bit planes need Cin % 8 == 0 (16-byte TMA strides)");
  PVG_CHECK_ARG((((uintptr_t)x_planes | (uintptr_t)w_planes | (uintptr_t)y | (uintptr_t)y_planes) & 15) == 0, "operands must be 16-byte aligned");
  PVG_CHECK_ARG(!y_planes || d->Cout % 8 == 0, "y_planes needs Cout % 8 == 0");
  g_bn.sums = bn_sums; g_bn.groups = bn_groups; g_amax_out = amax_out;
  const int rc = conv2d_fwd_h3(d, x_planes, w_planes, bias, y, y_planes, out_scale, (cudaStream_t)stream);
  g_bn.sums = nullptr; g_bn.groups = 0; g_amax_out = nullptr;
  return rc;
}

// One ConvLSTM cell step: gates = conv3x3([inputs..., h]) as ONE implicit GEMM over interleaved gate columns, with the cell update
// fused into its epilogue (see lstm_finish16).  z_planes: PVG_CORR_FP16_ALL planes of the (zero-padded) channel concat
// [N,H,W,d->Cin]; w_planes: planes of the interleaved gate weight pack (d->Cout = 4 * hidden channels, a multiple of 16);
// bias: interleaved [4C]; c_prev / c_new / h_new: [N,H,W,C]; gates_act (NULL for inference): [N,H,W,4C] activated gates.
extern "C" int pvg_convlstm_step(const pvg_conv_desc* d, const void* z_planes, const void* w_planes, const float* bias,
                                 const float* c_prev, float* c_new, float* h_new, float* gates_act, void* stream) {
  PVG_CHECK_ARG(d && z_planes && w_planes && c_prev && c_new && h_new, "null argument");
  PVG_CHECK_ARG(d->Cout % 16 == 0 && d->Cin % 8 == 0 && d->R == d->S && d->pad == (d->R - 1) / 2 && (d->R & 1), "unsupported cell shape");
  PVG_CHECK_ARG((((uintptr_t)z_planes | (uintptr_t)w_planes | (uintptr_t)bias | (uintptr_t)c_prev | (uintptr_t)c_new | (uintptr_t)h_new |
                  (uintptr_t)gates_act) & 15) == 0, "operands must be 16-byte aligned");
  pvg_conv_desc dd = *d;
  dd.act = PVG_ACT_LSTM;
  g_lstm.c_prev = c_prev; g_lstm.c_new = c_new; g_lstm.h_new = h_new;
  const int rc = conv2d_fwd_h3(&dd, z_planes, w_planes, bias, gates_act, nullptr, nullptr, (cudaStream_t)stream);
  g_lstm.c_prev = nullptr; g_lstm.c_new = nullptr; g_lstm.h_new = nullptr;
  return rc;
}
