// Hardware probe: cycles per tcgen05.mma (kind::f16, bf16, M = 128, K = 16) as a function of N and
// of the operand major-ness, issued back to back by one thread from shared-memory operands (zeros).
// Answers: does an MN-major operand (the weight-gradient GEMMs, K = pixels) cost more tensor time
// than a K-major one?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I road_segmentation_unet_b200/csrc \
//        tools/mma_probe.cu -o tools/mma_probe && tools/mma_probe
#include <cstdio>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace rsu;

__global__ void __launch_bounds__(128, 1)
    probe_kernel(int N, int a_mn, int b_mn, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // A: 2 atoms x 16 KiB (MN-major needs 64-row atoms), B: 4 atoms x 16 KiB; all zeros
  const uint32_t a_base = smem_base, b_base = smem_base + 32768u, bar = smem_base + 98304u;
  const uint32_t slot = bar + 16u;
  for (uint32_t i = threadIdx.x; i < 98304u / 16u; i += blockDim.x)
    st_shared_v4(smem_base + i * 16u, make_uint4(0, 0, 0, 0));
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint8_t* gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (slot - smem_base));
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, a_mn != 0, b_mn != 0);
    const uint32_t hi = desc_hi_sw128(1024u);
    // K-major: LBO unused (16); MN-major: LBO = distance between 64-element M/N atoms
    const uint32_t a_lo = desc_lo_sw128(a_base, a_mn ? 16384u : 16u);
    const uint32_t b_lo = desc_lo_sw128(b_base, b_mn ? 16384u : 16u);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      // alternate between 4 K slices like a real K loop
      const uint32_t k = static_cast<uint32_t>(i & 3);
      const uint32_t ka = a_mn ? k * 128u : k * 2u, kb = b_mn ? k * 128u : k * 2u;
      umma_bf16_lohi(tmem + (i & 1) * 256, a_lo + ka, hi, b_lo + kb, hi, idesc, 1u);
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

// The issue pattern of wgrad_gemm_kernel without any data movement: per "stage" four MMAs
// (128 x N x 16, both operands MN-major, 64-pixel atoms 8 KiB apart) on one accumulator, then a
// tcgen05.commit onto that stage's mbarrier; before re-using a stage buffer the thread waits for
// the commit issued `depth` stages earlier (what the TMA producer's empty-barrier wait amounts to).
// mode 0: as the kernel (commit per stage), mode 1: commit every second stage.
__global__ void __launch_bounds__(128, 1)
    stage_probe_kernel(int N, int mn, int per_stage, int depth, int mode, int stages_total, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = 49152u;  // (2 + 4) x 8 KiB
  const uint32_t bar = smem_base + 4u * stage_bytes;
  const uint32_t slot = bar + 64u;
  for (uint32_t i = threadIdx.x; i < (4u * stage_bytes) / 16u; i += blockDim.x)
    st_shared_v4(smem_base + i * 16u, make_uint4(0, 0, 0, 0));
  fence_proxy_async();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(bar + 8u * i, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint8_t* gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (slot - smem_base));
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, mn != 0, mn != 0);
    const uint32_t hi = desc_hi_sw128(1024u);
    uint32_t phase[4] = {0, 0, 0, 0};
    long long t0 = clock64();
    for (int s = 0; s < stages_total; ++s) {
      const int st = s & 3;
      if (s >= depth && (mode == 0 || ((s - depth) & 1) == 1)) {
        const int w = (s - depth) & 3;
        mbar_wait(bar + 8u * w, phase[w]);
        phase[w] ^= 1u;
      }
      const uint32_t a_addr = smem_base + st * stage_bytes;
      const uint32_t a_lo = desc_lo_sw128(a_addr, mn ? 8192u : 16u);
      const uint32_t b_lo = desc_lo_sw128(a_addr + 16384u, mn ? 8192u : 16u);
      for (int j = 0; j < per_stage; ++j) {
        const uint32_t k = mn ? j * 128u : j * 2u;
        umma_bf16_lohi(tmem, a_lo + k, hi, b_lo + k, hi, idesc, 1u);
      }
      if (mode == 0 || (s & 1) == 1) umma_commit(bar + 8u * st);
    }
    // drain
    umma_commit(bar + 8u * 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
  const int iters = 4096;
  printf("%5s %5s %5s %12s %10s\n", "N", "A", "B", "cycles/MMA", "ideal N/2");
  for (int grid : {1, 148}) {
    printf("grid = %d CTA(s)\n", grid);
    for (int N : {64, 128, 256}) {
      for (int mode = 0; mode < 4; ++mode) {
        const int a_mn = mode & 1, b_mn = mode >> 1;
        probe_kernel<<<grid, 128, 100 * 1024>>>(N, a_mn, b_mn, iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("N=%d a_mn=%d b_mn=%d: %s\n", N, a_mn, b_mn, cudaGetErrorString(e));
          return 1;
        }
        long long h[148];
        cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%5d %5s %5s %12.1f %10d\n", N, a_mn ? "MN" : "K", b_mn ? "MN" : "K",
               static_cast<double>(mx) / iters, N / 2);
      }
    }
  }
  cudaFuncSetAttribute(stage_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("stage loop (issue pattern of wgrad_gemm, no data movement), grid 148\n");
  printf("%5s %4s %10s %6s %5s %14s %8s\n", "N", "mn", "mma/stage", "depth", "mode", "cycles/stage", "ideal");
  const int stages_total = 2000;
  for (int N : {128, 256}) {
    for (int mn = 0; mn < 2; ++mn) {
      for (int per_stage : {2, 4, 8}) {
        for (int depth : {1, 3}) {
          for (int mode = 0; mode < 2; ++mode) {
            stage_probe_kernel<<<148, 128, 198 * 1024>>>(N, mn, per_stage, depth, mode, stages_total, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
              printf("stage probe: %s\n", cudaGetErrorString(e));
              return 1;
            }
            long long h[148];
            cudaMemcpy(h, d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("%5d %4d %10d %6d %5d %14.1f %8d\n", N, mn, per_stage, depth, mode,
                   static_cast<double>(mx) / stages_total, per_stage * N / 2);
          }
        }
      }
    }
  }
  return 0;
}
