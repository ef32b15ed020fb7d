// Hardware probe for tcgen05.mma.ws (weight-stationary, UTCHMMA.WS) on sm_100a:
//   1. cycles per MMA (M = 128, K = 16, bf16) for N = 64 / 128 when consecutive instructions share
//      the B operand through collector buffer b0 (fill, lastuse / fill, use, use, lastuse) -- does
//      the reuse spare the shared-memory read of B (48 -> 40 cycles at N = 64)?
//   2. the accumulator layout: is row i of D still TMEM lane i, column j TMEM column j?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I road_segmentation_unet_b200/csrc \
//        tools/ws_probe.cu -o tools/ws_probe && tools/ws_probe
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace rsu;

// MODE 0: plain MMA; 1: .ws discard; 2: .ws pairs (fill, lastuse); 3: .ws quads (fill, use, use,
// lastuse); 4: plain MMA pairs that share A (collector::a::fill / lastuse) -- all fully unrolled so
// that the issuing thread is never the limit
template <int MODE>
__global__ void __launch_bounds__(128, 1) ws_time_kernel(int N, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base, b_base = smem_base + 65536u, bar = smem_base + 98304u;
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
    const uint32_t idesc = make_idesc_bf16(128, N, false, false);
    const uint32_t hi = desc_hi_sw128(1024u);
    const uint32_t a_lo = desc_lo_sw128(a_base, 16u), b_lo = desc_lo_sw128(b_base, 16u);
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        // K slice q / 2 (q / 4 for quads), A tile q % 2 (q % 4), accumulator = A tile
        const uint32_t k = MODE == 3 ? (q >> 2) : (q >> 1);
        const uint32_t t = MODE == 3 ? (q & 3) : (q & 1);
        const uint32_t a = a_lo + t * 1024u + k * 2u;  // A tiles 16 KiB apart
        const uint32_t b = b_lo + k * 2u;
        const uint32_t d = tmem + t * 128u;
        if (MODE == 0) umma_bf16_lohi(d, a, hi, b, hi, idesc, 1u);
        else if (MODE == 1) umma_ws_bf16_lohi<3>(d, a, hi, b, hi, idesc, 1u);
        else if (MODE == 2) {
          if (t == 0) umma_ws_bf16_lohi<0>(d, a, hi, b, hi, idesc, 1u);
          else umma_ws_bf16_lohi<2>(d, a, hi, b, hi, idesc, 1u);
        } else if (MODE == 3) {
          if (t == 0) umma_ws_bf16_lohi<0>(d, a, hi, b, hi, idesc, 1u);
          else if (t == 3) umma_ws_bf16_lohi<2>(d, a, hi, b, hi, idesc, 1u);
          else umma_ws_bf16_lohi<1>(d, a, hi, b, hi, idesc, 1u);
        } else {
          // same A (tile 0, slice k), two different B tiles 16 KiB apart, two accumulators
          const uint32_t a0 = a_lo + k * 2u, bb = b_lo + t * 1024u + k * 2u;
          if (t == 0)
            asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
                         "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], da, db, %5, p;\n\t}"
                         ::"r"(d), "r"(a0), "r"(hi), "r"(bb), "r"(hi), "r"(idesc), "r"(1u) : "memory");
          else
            asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
                         "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], da, db, %5, p;\n\t}"
                         ::"r"(d), "r"(a0), "r"(hi), "r"(bb), "r"(hi), "r"(idesc), "r"(1u) : "memory");
        }
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

// D = A B^T with A[i][0] = (i % 16) + 1, B[j][0] = (j % 8) + 1, everything else 0: D[i][j] is their
// product.  Two A tiles share B through the collector (fill, lastuse); out[t][i][j] as fp32.
__global__ void __launch_bounds__(128, 1) ws_layout_kernel(int N, int use_ws, float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base, b_base = smem_base + 65536u, bar = smem_base + 98304u;
  const uint32_t slot = bar + 16u;
  uint8_t* gen = smem_raw + (smem_base - smem_u32(smem_raw));
  for (uint32_t i = threadIdx.x; i < 98304u / 16u; i += blockDim.x)
    st_shared_v4(smem_base + i * 16u, make_uint4(0, 0, 0, 0));
  __syncthreads();
  auto put = [&](uint32_t tile_off, int row, float v) {  // element (row, k = 0) of a K-major SW128 tile
    const uint32_t off = tile_off + (row >> 3) * 1024u + (row & 7) * 128u + ((row & 7) << 4);
    *reinterpret_cast<__nv_bfloat16*>(gen + off) = __float2bfloat16(v);
  };
  {
    const int i = threadIdx.x;
    put(0u, i, static_cast<float>((i % 16) + 1));                 // A tile 0
    put(16384u, i, static_cast<float>(2 * ((i % 16) + 1)));       // A tile 1 = 2 x tile 0
    if (i < N) put(65536u, i, static_cast<float>((i % 8) + 1));   // B
  }
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (slot - smem_base));
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, false, false);
    const uint32_t hi = desc_hi_sw128(1024u);
    const uint32_t a_lo = desc_lo_sw128(a_base, 16u), b_lo = desc_lo_sw128(b_base, 16u);
    if (use_ws) {
      umma_ws_bf16_lohi<0>(tmem, a_lo, hi, b_lo, hi, idesc, 0u);
      umma_ws_bf16_lohi<2>(tmem + 128u, a_lo + 1024u, hi, b_lo, hi, idesc, 0u);
    } else {
      umma_bf16_lohi(tmem, a_lo, hi, b_lo, hi, idesc, 0u);
      umma_bf16_lohi(tmem + 128u, a_lo + 1024u, hi, b_lo, hi, idesc, 0u);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t = 0; t < 2; ++t) {
    for (int ch = 0; ch < N / 32; ++ch) {
      uint32_t r[32];
      tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + t * 128u + ch * 32, r);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j)
        out[(t * 128 + warp * 32 + lane) * N + ch * 32 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  float* o;
  cudaMalloc(&o, 2 * 128 * 256 * sizeof(float));
  cudaFuncSetAttribute(ws_time_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
  cudaFuncSetAttribute(ws_time_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
  cudaFuncSetAttribute(ws_time_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
  cudaFuncSetAttribute(ws_time_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
  cudaFuncSetAttribute(ws_time_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
  cudaFuncSetAttribute(ws_layout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
  static float h[2 * 128 * 256];
  for (int N : {64, 128}) {
    for (int ws = 0; ws < 2; ++ws) {
      cudaMemset(o, 0, sizeof(h));
      ws_layout_kernel<<<1, 128, 100 * 1024>>>(N, ws, o);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("layout N=%d ws=%d: %s\n", N, ws, cudaGetErrorString(e));
        return 1;
      }
      cudaMemcpy(h, o, 2 * 128 * N * sizeof(float), cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int t = 0; t < 2; ++t)
        for (int i = 0; i < 128; ++i)
          for (int j = 0; j < N; ++j) {
            const float want = (t + 1) * ((i % 16) + 1) * ((j % 8) + 1);
            if (h[(t * 128 + i) * N + j] != want) {
              if (bad < 6) printf("  N=%d ws=%d t=%d D[%d][%d] = %g, expected %g\n", N, ws, t, i, j,
                                  h[(t * 128 + i) * N + j], want);
              ++bad;
            }
          }
      printf("layout N=%d %s: %s (%d mismatches)\n", N, ws ? "tcgen05.mma.ws fill/lastuse" : "tcgen05.mma",
             bad ? "DIFFERENT" : "row i = lane i, column j = column j", bad);
    }
  }
  const int iters = 4096;
  const char* names[5] = {"tcgen05.mma", "ws discard", "ws fill,lastuse", "ws fill,use,use,lastuse",
                          "mma, A fill,lastuse"};
  for (int grid : {148}) {
    printf("grid = %d CTA(s): cycles per MMA (M = 128, K = 16)\n", grid);
    for (int N : {64, 128}) {
      for (int mode = 0; mode < 5; ++mode) {
        if (mode == 0) ws_time_kernel<0><<<grid, 128, 100 * 1024>>>(N, iters, d);
        if (mode == 1) ws_time_kernel<1><<<grid, 128, 100 * 1024>>>(N, iters, d);
        if (mode == 2) ws_time_kernel<2><<<grid, 128, 100 * 1024>>>(N, iters, d);
        if (mode == 3) ws_time_kernel<3><<<grid, 128, 100 * 1024>>>(N, iters, d);
        if (mode == 4) ws_time_kernel<4><<<grid, 128, 100 * 1024>>>(N, iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("N=%d mode=%d: %s\n", N, mode, cudaGetErrorString(e));
          return 1;
        }
        long long hh[148];
        cudaMemcpy(hh, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = hh[i] > mx ? hh[i] : mx;
        printf("  N=%3d %-26s %7.1f   (tensor-bound %d)\n", N, names[mode], static_cast<double>(mx) / iters, N / 2);
      }
    }
  }
  return 0;
}
