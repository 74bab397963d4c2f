// tcgen05 / TMEM / TMA / mbarrier PTX wrappers and descriptor builders shared by the tensor-core kernels (sm_100a).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace pvg {

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar), done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one lane of a CONVERGED warp (all 32 lanes must call it).  Single-thread work (TMA issue, tcgen05.mma issue) is written as
//   leader = elect_one();  loop { all lanes wait on the mbarrier; if (leader) { issue }; __syncwarp(); }
// so that ptxas knows exactly one lane is active and emits straight-line UTMALDG / UTCHMMA: behind an `if (lane == 0)` it
// wraps EVERY such instruction in an ELECT + BRA.U.ANY loop (~20 instructions per MMA), which made the MMA-issuing thread,
// not the tensor pipe, the limiter of the 128-wide tiles (ncu: 37 % pipe-active with the issuer never waiting for data).
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f16 with bf16 operands (K = 16 per instruction), fp32 accumulate: the correction terms of the split product
__device__ __forceinline__ void umma_bf16(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// The same load without the wait: issue several, then tmem_ld_wait() once, then tmem_ld_use() on every register block before
// its values are read (ties the uses to the wait: the compiler may otherwise schedule them ahead of it).
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_use(uint32_t* r) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                    "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// K-major swizzled operand tile of fp32/tf32: rows of KC*4 bytes (128 B -> SWIZZLE_128B, 64 B -> SWIZZLE_64B), 8-row
// swizzle atoms stacked along M/N.
template <int KC>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);   // start address, 16-byte units
  d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)((8 * KC * 4) >> 4) << 32;       // stride byte offset between 8-row groups (1024 B or 512 B)
  d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
  d |= (uint64_t)(KC == 32 ? 2 : 4) << 61;        // SWIZZLE_128B / SWIZZLE_64B
  return d;
}

template <int BN>
__host__ __device__ constexpr uint32_t make_idesc_tf32() {
  return (1u << 4)                 // D format: F32
         | (2u << 7)               // A format: TF32
         | (2u << 10)              // B format: TF32
         | ((uint32_t)(BN >> 3) << 17)   // N >> 3
         | ((uint32_t)(128 >> 4) << 24); // M >> 4 ; A and B K-major (bits 15, 16 = 0)
}


// kind::f16 instruction descriptor: (bf16 x bf16 | fp16 x fp16) -> fp32, M = 128; K-major operands unless a_mn / b_mn.
// The two operand formats must be equal (a bf16 x fp16 MMA traps as an illegal instruction on sm_100a).
template <int BN>
__host__ __device__ constexpr uint32_t make_idesc_f16(bool fp16, bool a_mn = false, bool b_mn = false, int m = 128) {
  return (1u << 4)                      // D format: F32
         | ((fp16 ? 0u : 1u) << 7)      // A format: F16 = 0, BF16 = 1
         | ((fp16 ? 0u : 1u) << 10)     // B format
         | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16)
         | ((uint32_t)(BN >> 3) << 17)
         | ((uint32_t)(m >> 4) << 24);
}

// ---- cta_group::2 (CTA pair) variants -------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the CTA-parity bit of a shared::cluster address -> the even (leader) CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA loads of a CTA pair: the transaction bytes are credited to the LEADER CTA's mbarrier (same offset, parity bit cleared)
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem2_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem2_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[256 x N] (rows 0-127 in the leader's TMEM, 128-255 in the peer's) += A (128 rows from each CTA) * B (N/2 rows from each)
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all prior MMAs of the pair -> arrive on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the LEADER CTA's barrier at this offset (from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}

// MN-major tf32 operand tile.  For 32-bit MN-major operands the only swizzle the tensor core accepts is
// SWIZZLE_128B_BASE32B (32-byte chunks XOR-ed with row % 4; CUTLASS: "for mn-major tf32 operands, SW128_32B is the only
// available smem layout"), which TMA produces with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  Element (mn, k) lives at row k
// (128 B = 32 consecutive mn values); 4-row (512 B) swizzle atoms are stacked along K, so one K=8 MMA spans two atoms
// (SBO = 512 B); the next 32 mn values start `mn_group_stride` bytes further (LBO).
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t mn_group_stride) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((mn_group_stride >> 4) & 0x3FFF) << 16;   // leading byte offset: stride between 32-element MN groups
  d |= (uint64_t)(512 >> 4) << 32;                           // stride byte offset: 512 B between 4-k swizzle atoms
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                                    // SWIZZLE_128B_BASE32B
  return d;
}

// MN-major bf16 operand tile with the 64-byte swizzle: element (mn, k) lives at row k (64 B = 32 consecutive mn values),
// 8-row (512 B) swizzle atoms stacked along K (SBO = 512 B, one K = 16 MMA spans two atoms); the next 32 mn values start
// `mn_group_stride` bytes further (LBO).  TMA produces it with CU_TENSOR_MAP_SWIZZLE_64B from a {32 ch, 32 px} bf16 box.
__device__ __forceinline__ uint64_t make_mnmajor_sw64_desc(uint32_t smem_addr, uint32_t mn_group_stride) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((mn_group_stride >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;                                    // SWIZZLE_64B
  return d;
}

// instruction descriptor of the pair MMA: M = 256 (128 rows per CTA)
template <int BN>
__host__ __device__ constexpr uint32_t make_idesc_tf32_m256() {
  return (make_idesc_tf32<BN>() & ~(0x1Fu << 24)) | ((uint32_t)(256 >> 4) << 24);
}

// instruction descriptor, kind::tf32, fp32 accumulate, M = 128; a_mn / b_mn select MN-major operands
template <int BN>
__host__ __device__ constexpr uint32_t make_idesc_tf32_ex(bool a_mn, bool b_mn) {
  return make_idesc_tf32<BN>() | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn();
// 4-D map over an NHWC fp32 tensor with box {box_c, tw, th, tn}, zero OOB fill; 128B swizzle with 16-byte atoms
// (K-major operands) or 32-byte atoms (atom32 = true, MN-major tf32 operands)
int encode_nhwc_map(CUtensorMap* m, const float* x, int N, int H, int W, int C, int box_c, int tw, int th, int tn,
                    bool atom32 = false, bool sw64 = false);
// 5-D map over the two 16-bit correction planes [2][N][H][W][C] of an activation tensor (pvg_split_16): box
// {32 ch, tw, th, tn, 2} lands as two consecutive 64-byte-row tiles (plane 0 = f16(lo * 2^12), plane 1 = f16(x)), SWIZZLE_64B
int encode_nhwc_16x2_map(CUtensorMap* m, const void* planes, int N, int H, int W, int C, int tw, int th, int tn);
// 3-D map over the two 16-bit planes [2][rows][K] of a packed weight: box {32, box_rows, 2}, SWIZZLE_64B
int encode_w_16x2_map(CUtensorMap* m, const void* planes, int rows, int K, int box_rows);
// the 128-pixel patch (tw, th, tn) that wastes the fewest MMA rows
void choose_patch(int N, int H, int W, int* tw, int* th, int* tn);

// ---------------------------------------------------------------------------------------------------------------
// shared by the implicit-GEMM forward kernels (conv_umma.cu, conv_h3.cu)
// ---------------------------------------------------------------------------------------------------------------
struct ConvParams {
  int N, H, W, Cin, Cout, R, S, pad, act;
  float slope;
  int tw, th, tn;               // M-tile patch (tw*th*tn == 128)
  int tiles_w, tiles_h, tiles_n;
  const float* bias;
  float* y;
  int corr_fp16;                // NPROD == 2: the 16-bit correction planes are fp16 (else bf16)
  uint16_t* y_planes;           // conv_h3.cu, optional: fp16 plane pair [2][N*H*W*Cout] of y (PVG_CORR_FP16_ALL), written by the
  int64_t y_numel;              //   epilogue so that the next convolution needs no separate split pass; y_numel = N*H*W*Cout
  const float* out_scale;       // conv_h3.cu, optional (device): accumulator scale (1 / S of a scaled gradient operand)
  const float* lstm_c_prev;     // conv_h3.cu, act == PVG_ACT_LSTM: fused ConvLSTM cell operands [N][H][W][Cout / 4]
  float* lstm_c_new;
  float* lstm_h_new;
  double* bn_sums;              // conv_h3.cu, optional: [groups][2][Cout] per-channel sum / sum of squares of y (BatchNorm statistics)
  int bn_groups, bn_samples_per_group;
  // Same-address fp64 atomics serialise in L2 (measured: ~2 ns per atomic with one [2][Cout] target, i.e. 0.9 ms for a 48-image
  // decoder launch): CTAs accumulate into one of bn_replicas copies [replica][groups][2][Cout] in a library-owned, zero-kept
  // workspace; the last CTA to finish (bn_ticket) adds the copies into bn_sums and zeroes them again.  bn_replicas == 0: direct.
  double* bn_ws;
  uint32_t* bn_ticket;
  int bn_replicas;
  uint32_t* amax_out;           // conv_h3.cu, optional: atomicMax of the bit patterns of |y| (zero-initialised by the caller)
  int dbg;                      // conv_h3.cu timing experiments
};

// bias + activation of 16 consecutive output channels and their NHWC store.  Everything that is uniform over the tile (bias
// pointer, activation kind, Cout bounds) is tested once per 16 channels, not per element: the straightforward per-element
// form compiled to ~66 instructions per output value and made the final epilogue 30 % of a 72-k-iteration tile (ncu).
__device__ __forceinline__ void finish16(float (&v)[16], const float* __restrict__ bias, int co, int Cout, int act, float slope,
                                         float* __restrict__ yrow, bool valid, bool vec_ok) {
  if (co >= Cout) return;
  const bool full = co + 16 <= Cout;
  if (bias != nullptr) {
    if (full && vec_ok) {                         // Cout % 4 == 0 and 16-byte aligned rows => bias + co is 16-byte aligned too
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 b = ldg4(bias + co + j);
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (co + j < Cout) v[j] += __ldg(bias + co + j);
    }
  }
  switch (act) {
    case PVG_ACT_LRELU:
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * slope;
      break;
    case PVG_ACT_RELU:
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
      break;
    case PVG_ACT_TANH:
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = tanhf(v[j]);
      break;
    case PVG_ACT_SIGMOID:
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 1.f / (1.f + expf(-v[j]));
      break;
    default: break;
  }
  if (!valid) return;
  if (full && vec_ok) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) stg4(yrow + co + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (co + j < Cout) yrow[co + j] = v[j];
  }
}


}  // namespace pvg
