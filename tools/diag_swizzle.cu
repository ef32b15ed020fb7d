// Hardware probe (not part of the product): does a tcgen05 shared-memory descriptor with the
// 128-byte swizzle accept a start address that is 128-byte but not 1024-byte aligned (i.e. a
// window that starts at an arbitrary ROW of a TMA-written SWIZZLE_128B tile), and which value of
// the descriptor's base-offset field makes it read the right bytes?  The halo-tile convolution
// (one TMA box per 64-channel chunk, nine tap-shifted operand views) depends on the answer.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I road_segmentation_unet_b200/csrc \
//        tools/diag_swizzle.cu -o tools/diag_swizzle && tools/diag_swizzle
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "ptx.cuh"

using namespace rsu;

constexpr int kRows = 192;  // rows of 128 B in the A tile (TMA box = 64 x 192)

struct Probe {
  int row_off;      // first smem row of the operand window
  int group_rows;   // rows between consecutive 8-row groups (SBO = group_rows * 128 B)
  int base_mode;    // 0: base_offset = 0, 1: base_offset = (addr >> 7) & 7
  int mn_major;     // 0: A is K-major (conv), 1: A and B are MN-major (wgrad)
  int lbo_rows;     // MN-major only: rows between the two 64-wide M atoms
};

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo,
                                               int base_mode) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  if (base_mode == 1) d |= static_cast<uint64_t>((saddr >> 7) & 0x7) << 49;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1)
    probe_kernel(const __grid_constant__ CUtensorMap a_map, const __grid_constant__ CUtensorMap b_map,
                 Probe pr, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_s = base;                    // 192 rows x 128 B = 24 KiB
  const uint32_t b_s = base + kRows * 128;      // 64 rows x 128 B  =  8 KiB (1024-aligned)
  const uint32_t bar = b_s + 64 * 128;
  const uint32_t mma_bar = bar + 8;
  const uint32_t slot = bar + 16;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (slot - base));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(mma_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, kRows * 128 + 64 * 128);
    tma_load_2d(a_s, &a_map, bar, 0, 0);
    tma_load_2d(b_s, &b_map, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t sbo = pr.group_rows * 128;
    if (!pr.mn_major) {
      // D[m, n] = sum_k A[row(m), k] * B[n, k]; K = 64 in four steps of 16 (32 B along the row)
      const uint32_t idesc = make_idesc_bf16(128, 64, false, false);
      for (int j = 0; j < 4; ++j) {
        const uint64_t ad = desc_sw128(a_s + pr.row_off * 128 + 32 * j, 16, sbo, pr.base_mode);
        const uint64_t bd = desc_sw128(b_s + 32 * j, 16, 1024, 0);
        umma_bf16(tmem, ad, bd, idesc, j ? 1u : 0u);
      }
    } else {
      // D[m, n] = sum_p A[row(p) + atom(m) * lbo_rows, m % 64] * B[p, n]; K = 64 pixels, 16 per step
      const uint32_t idesc = make_idesc_bf16(128, 64, true, true);
      for (int j = 0; j < 4; ++j) {
        const uint64_t ad = desc_sw128(a_s + (pr.row_off + 2 * j * pr.group_rows) * 128,
                                       pr.lbo_rows * 128, sbo, pr.base_mode);
        const uint64_t bd = desc_sw128(b_s + j * 2048, 8192, 1024, 0);
        umma_bf16(tmem, ad, bd, idesc, j ? 1u : 0u);
      }
    }
    umma_commit(mma_bar);
    mbar_wait(mma_bar, 0);
  }
  __syncthreads();
  tc_fence_after();
  for (int ch = 0; ch < 2; ++ch) {
    uint32_t r[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + ch * 32, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + ch * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float bf(float v) { return __bfloat162float(__float2bfloat16(v)); }

int main() {
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fnp);
  std::vector<float> A(kRows * 64), B(64 * 64);
  std::vector<__nv_bfloat16> Ah(kRows * 64), Bh(64 * 64);
  srand(7);
  for (int i = 0; i < kRows * 64; ++i) {
    A[i] = bf((rand() % 2001 - 1000) / 1000.f);
    Ah[i] = __float2bfloat16(A[i]);
  }
  for (int i = 0; i < 64 * 64; ++i) {
    B[i] = bf((rand() % 2001 - 1000) / 1000.f);
    Bh[i] = __float2bfloat16(B[i]);
  }
  __nv_bfloat16 *dA, *dB;
  float* dO;
  cudaMalloc(&dA, Ah.size() * 2);
  cudaMalloc(&dB, Bh.size() * 2);
  cudaMalloc(&dO, 128 * 64 * 4);
  cudaMemcpy(dA, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap am, bm;
  {
    cuuint64_t dims[2] = {64, kRows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, kRows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&am, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t dimsb[2] = {64, 64};
    cuuint32_t boxb[2] = {64, 64};
    CUresult r2 = enc(&bm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsb, strides, boxb, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
      printf("encode failed %d %d\n", (int)r, (int)r2);
      return 1;
    }
  }
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> O(128 * 64);
  int n_bad = 0;
  for (int mn = 0; mn < 2; ++mn)
    for (int group_rows : {8, 10})
      for (int row_off : {0, 1, 2, 3, 8, 10, 11, 21, 22})
        for (int base_mode = 0; base_mode < 2; ++base_mode) {
          Probe pr{row_off, group_rows, base_mode, mn, mn ? 1 : 0};
          if (mn && row_off + 8 * group_rows + 1 > kRows) continue;
          if (!mn && row_off + 16 * group_rows > kRows) continue;
          cudaMemset(dO, 0, 128 * 64 * 4);
          probe_kernel<<<1, 128, 48 * 1024>>>(am, bm, pr, dO);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("mn=%d group_rows=%d row_off=%d base_mode=%d : CUDA error %s\n", mn, group_rows,
                   row_off, base_mode, cudaGetErrorString(e));
            return 2;
          }
          cudaMemcpy(O.data(), dO, 128 * 64 * 4, cudaMemcpyDeviceToHost);
          double err = 0, ref2 = 0;
          for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 64; ++n) {
              double acc = 0;
              if (!mn) {
                const int row = row_off + (m / 8) * group_rows + (m % 8);
                for (int k = 0; k < 64; ++k) acc += (double)A[row * 64 + k] * B[n * 64 + k];
              } else {
                // K index p = 0..63 -> smem row row_off + (p / 8) * group_rows + p % 8 (+ atom shift)
                for (int p = 0; p < 64; ++p) {
                  const int row = row_off + (p / 8) * group_rows + (p % 8) + (m / 64) * pr.lbo_rows;
                  acc += (double)A[row * 64 + (m % 64)] * B[p * 64 + n];
                }
              }
              const double d = O[m * 64 + n] - acc;
              err += d * d;
              ref2 += acc * acc;
            }
          const double rel = sqrt(err / ref2);
          const bool ok = rel < 1e-3;
          if (!ok) ++n_bad;
          printf("%s group_rows=%2d row_off=%2d base_mode=%d : rel err %.3e %s\n",
                 mn ? "MN-major" : "K-major ", group_rows, row_off, base_mode, rel, ok ? "OK" : "WRONG");
        }
  printf("done, %d mismatching variants\n", n_bad);
  return 0;
}
