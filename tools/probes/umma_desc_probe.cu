// Hardware probe for the halo-reuse design of the conv kernel: how does tcgen05.mma address a
// K-major SWIZZLE_128B operand whose descriptor START ADDRESS is offset by a number of 128-byte rows that is not a multiple
// of 8 (a tap shift inside one haloed tile), with a stride between 8-row groups (SBO) that is / is not a multiple of 1024
// bytes, and what does the descriptor's 3-bit "base offset" field (bits 49-51) do?
//
// Set-up: one CTA.  TMA loads X[rows = 256][32 fp32] (128-byte rows, SWIZZLE_128B) into a 1024-byte aligned tile, then ONE
// MMA (M = 128, N = 16, K = 8, kind::tf32) runs with
//     A = descriptor(start = tile + row_off * 128, SBO = sbo_rows * 128, base_offset = bo),
//     B = selector matrix: B[n][k] = (k == n) for n < 8  ->  D[m][n] = A[m][k = n]   (n < 8).
// X is filled once with its row index (X[r][c] = r) and once with its column index (X[r][c] = c): D then tells, for every
// MMA row m and every k, WHICH shared-memory row and column the tensor core actually read.  The host prints, per
// configuration, whether the mapping is the wanted "row = row_off + (m / 8) * sbo_rows + m % 8, column = k".
// A second table does the same for the MN-major tf32 operand of the weight-gradient kernel (pixel-row offsets = tap shifts
// along K inside a {32 ch x many px} tile).
//
// build + run (on a B200):  nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../playablevideogeneration_b200/csrc \
//                                -o /tmp/umma_desc_probe umma_desc_probe.cu -lcuda && /tmp/umma_desc_probe
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "umma.cuh"

namespace pvg {
void set_error(const std::string& m) { fprintf(stderr, "error: %s\n", m.c_str()); }
EncodeTiledFn get_encode_fn() {          // stand-alone copy of the lookup in conv_umma.cu
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return nullptr;
  return (EncodeTiledFn)ptr;
}
}  // namespace pvg
using namespace pvg;

constexpr int kRows = 256;                       // rows of X staged in shared memory (32 KB)
constexpr int kTileBytes = kRows * 128;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ bsel, int row_off, int sbo_rows, int base_off,
             int mn_major, float* __restrict__ out /* [128][8] */) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_tile = smem;                        // [256 rows][128 B], SWIZZLE_128B
  uint8_t* b_tile = smem + kTileBytes;           // [16 rows][128 B], SWIZZLE_128B (written by the threads below)
  uint64_t* bar = (uint64_t*)(b_tile + 16 * 128);
  uint64_t* done = bar + 1;
  uint32_t* tmem_slot = (uint32_t*)(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_barrier_init(); }
  // B tile: row n (128 B = 32 floats), 16-byte chunk j of row n lives at chunk (j ^ (n & 7)) (SWIZZLE_128B, rows n < 8 of one atom)
  for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) {
    const int n = i >> 5, k = i & 31;
    const int chunk = (k >> 2) ^ (n & 7);
    reinterpret_cast<float*>(b_tile + n * 128 + chunk * 16)[k & 3] = bsel[n * 32 + k];
  }
  fence_proxy_async();                           // generic-proxy writes of B visible to the tensor core (async proxy)
  if (warp == 0) tmem_alloc(tmem_slot, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  if (warp == 0) {
    const uint32_t leader = elect_one();
    if (leader) {
      mbar_expect_tx(bar, kTileBytes);
      tma_load_2d(a_tile, &tmX, bar, 0, 0);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    tc_fence_after();
    if (leader) {
      uint64_t ad = 0;
      const uint32_t a_addr = smem_u32(a_tile) + (uint32_t)row_off * 128u;
      ad |= (uint64_t)((a_addr & 0x3FFFF) >> 4);
      ad |= (uint64_t)1 << 16;
      ad |= (uint64_t)(((uint32_t)sbo_rows * 128u) >> 4) << 32;
      ad |= (uint64_t)1 << 46;
      ad |= (uint64_t)(base_off & 7) << 49;      // matrix base offset
      ad |= (uint64_t)2 << 61;                   // SWIZZLE_128B
      const uint64_t bd = make_kmajor_desc<32>(smem_u32(b_tile));
      if (mn_major) {
        // A = MN-major tf32 operand as in conv_wgrad_umma.cu: M = channels (32 per group, next group LBO = 4096 B further),
        // K = pixel rows, 4-row SWIZZLE_128B_BASE32B atoms (SBO = sbo_rows * 128, 512 B in the kernel); the tile was loaded
        // with the 32-byte-atom TMA swizzle.  D[m][n] = X[row_off + n][m % 32 (+ group * 32 rows further)].
        uint64_t md = 0;
        md |= (uint64_t)((a_addr & 0x3FFFF) >> 4);
        md |= (uint64_t)((4096 >> 4) & 0x3FFF) << 16;
        md |= (uint64_t)(((uint32_t)sbo_rows * 128u) >> 4) << 32;
        md |= (uint64_t)1 << 46;
        md |= (uint64_t)(base_off & 7) << 49;
        md |= (uint64_t)1 << 61;                 // SWIZZLE_128B_BASE32B
        umma_tf32(tmem_acc, md, bd, make_idesc_tf32_ex<16>(true, false), 0);
      } else {
        umma_tf32(tmem_acc, ad, bd, make_idesc_tf32<16>(), 0);
      }
      umma_commit(done);
    }
    __syncwarp();
  }
  mbar_wait(done, 0);
  tc_fence_after();
  float v[16];
  tmem_ld16(tmem_acc + ((uint32_t)(warp * 32) << 16), v);
  const int m = warp * 32 + lane;
  for (int n = 0; n < 8; ++n) out[m * 8 + n] = v[n];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_acc, 32);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s -> %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

int main() {
  std::vector<float> xr(kRows * 32), xc(kRows * 32), bsel(16 * 32, 0.f);
  for (int r = 0; r < kRows; ++r)
    for (int c = 0; c < 32; ++c) { xr[r * 32 + c] = (float)r; xc[r * 32 + c] = (float)c; }
  for (int n = 0; n < 8; ++n) bsel[n * 32 + n] = 1.f;
  float *d_xr, *d_xc, *d_b, *d_out;
  CK(cudaMalloc(&d_xr, xr.size() * 4)); CK(cudaMalloc(&d_xc, xc.size() * 4)); CK(cudaMalloc(&d_b, bsel.size() * 4));
  CK(cudaMalloc(&d_out, 128 * 8 * 4));
  CK(cudaMemcpy(d_xr, xr.data(), xr.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_xc, xc.data(), xc.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_b, bsel.data(), bsel.size() * 4, cudaMemcpyHostToDevice));
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { fprintf(stderr, "no cuTensorMapEncodeTiled\n"); return 1; }
  auto make_map = [&](float* p, CUtensorMap* m, bool atom32 = false) {
    cuuint64_t dims[2] = {32, (cuuint64_t)kRows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {32, (cuuint32_t)kRows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "encode failed %d\n", (int)r); exit(1); }
  };
  CUtensorMap mr, mc;
  make_map(d_xr, &mr); make_map(d_xc, &mc);
  const int smem = kTileBytes + 16 * 128 + 1024 + 64;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  std::vector<float> rows(128 * 8), cols(128 * 8);
  const int row_offs[] = {0, 1, 2, 3, 8, 9, 16, 17, 18};
  const int sbos[] = {8, 10, 16, 18};
  printf("row_off sbo_rows base_off | rows_ok cols_ok | first mismatching (m,k): got row/col, wanted row/col\n");
  for (int ro : row_offs)
    for (int sbo : sbos)
      for (int bo = 0; bo < 8; ++bo) {
        if (bo != 0 && bo != (ro & 7)) continue;               // the two candidates: no base offset / base offset = start phase
        CK(cudaMemset(d_out, 0, 128 * 8 * 4));
        probe_kernel<<<1, 128, smem>>>(mr, d_b, ro, sbo, bo, 0, d_out);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(rows.data(), d_out, rows.size() * 4, cudaMemcpyDeviceToHost));
        probe_kernel<<<1, 128, smem>>>(mc, d_b, ro, sbo, bo, 0, d_out);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(cols.data(), d_out, cols.size() * 4, cudaMemcpyDeviceToHost));
        int bad_r = 0, bad_c = 0, fm = -1, fk = -1;
        for (int m = 0; m < 128; ++m)
          for (int k = 0; k < 8; ++k) {
            const int want_r = ro + (m / 8) * sbo + (m % 8), want_c = k;
            if (want_r >= kRows) continue;
            const bool br = (int)rows[m * 8 + k] != want_r, bc = (int)cols[m * 8 + k] != want_c;
            bad_r += br; bad_c += bc;
            if ((br || bc) && fm < 0) { fm = m; fk = k; }
          }
        printf("%7d %8d %8d | %7s %7s |", ro, sbo, bo, bad_r ? "NO" : "yes", bad_c ? "NO" : "yes");
        if (fm >= 0)
          printf(" (%d,%d): got %d/%d, wanted %d/%d   [%d row, %d col mismatches]", fm, fk, (int)rows[fm * 8 + fk], (int)cols[fm * 8 + fk],
                 ro + (fm / 8) * sbo + fm % 8, fk, bad_r, bad_c);
        printf("\n");
      }
  // ---- MN-major operand (weight-gradient kernel): pixel-row offsets inside a 32-channel x many-pixel tile --------------
  CUtensorMap mr32, mc32;
  make_map(d_xr, &mr32, true); make_map(d_xc, &mc32, true);
  printf("\nMN-major tf32 (SWIZZLE_128B_BASE32B, TMA ATOM_32B): D[m][n] should be X[row_off + n][m] for m < 32\n");
  printf("row_off sbo_rows base_off | rows_ok cols_ok | first mismatching (m,n): got row/col, wanted row/col\n");
  const int mn_offs[] = {0, 1, 2, 3, 4, 5, 8, 34, 35};
  for (int ro : mn_offs)
    for (int bo = 0; bo < 8; ++bo) {
      if (bo != 0 && bo != (ro & 3) && bo != (ro & 7)) continue;
      CK(cudaMemset(d_out, 0, 128 * 8 * 4));
      probe_kernel<<<1, 128, smem>>>(mr32, d_b, ro, 4, bo, 1, d_out);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(rows.data(), d_out, rows.size() * 4, cudaMemcpyDeviceToHost));
      probe_kernel<<<1, 128, smem>>>(mc32, d_b, ro, 4, bo, 1, d_out);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(cols.data(), d_out, cols.size() * 4, cudaMemcpyDeviceToHost));
      int bad_r = 0, bad_c = 0, fm = -1, fk = -1;
      for (int m = 0; m < 32; ++m)                       // first channel group only (the others sit LBO = 32 rows further)
        for (int n = 0; n < 8; ++n) {
          const bool br = (int)rows[m * 8 + n] != ro + n, bc = (int)cols[m * 8 + n] != m;
          bad_r += br; bad_c += bc;
          if ((br || bc) && fm < 0) { fm = m; fk = n; }
        }
      printf("%7d %8d %8d | %7s %7s |", ro, 4, bo, bad_r ? "NO" : "yes", bad_c ? "NO" : "yes");
      if (fm >= 0)
        printf(" (%d,%d): got %d/%d, wanted %d/%d   [%d row, %d col mismatches]", fm, fk, (int)rows[fm * 8 + fk], (int)cols[fm * 8 + fk],
               ro + fk, fm, bad_r, bad_c);
      printf("\n");
    }
  return 0;
}
