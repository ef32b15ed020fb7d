// First-layer (Cin = 3) 3x3 convolution and its weight gradient on tcgen05 tensor cores, with the
// im2col operand generated on the fly (sm_100a).
//
// The Cin = 3 layers (conv_0/conv1, conv_dilut_0/atrous_conv1: src/unet.py:22-23, 29-30, 34-35,
// 42-43) have K = 27: as GEMMs they are pure HBM traffic.  Through a materialised im2col tensor
// (64 bf16 per output pixel) the forward pass wrote 2.4 GB, read it back and wrote 2.4 GB of
// output; here four producer warps build each 128-pixel x 32-K operand tile directly in shared
// memory (colour adjust, dropout, bf16, 128-byte swizzle), so the only large HBM stream left is
// the output itself (forward) or dZ (weight gradient).  The image (fp32, 12 bytes per pixel, read
// 9 times) stays in L1 / L2.
//
//   forward : D[pixel, co]   = sum_k col[pixel, k] * W[co, k]      k = tap*3 + c  (k < 27)
//   wgrad   : G[k, co]      += sum_pixel col[pixel, k] * dZ[pixel, co]
// col[pixel, 27] = 1 (its row of G is BiasAddGrad; the packed forward weights are zero there).
//
// Shared-memory operand image: row m = pixel (m / 8, m % 8) of an 8 x 16 tile, 128 bytes per row,
// 16-byte chunk c stored at c ^ (m & 7) -- exactly what TMA writes for SWIZZLE_128B, so the same
// bytes serve as the K-major A operand of the forward GEMM and as the MN-major A operand (K =
// pixels) of the weight gradient.
//
// The image patch under a tile ((16 + 2d) rows x (8 + 2d) pixels x 3 floats) is fetched by TMA
// into a ring of small shared-memory buffers several tiles ahead, so the producers never wait
// for global memory (with direct loads one tile cost a full L2 round trip and the kernels ran at
// a third of the HBM rate).
//
// Warp roles (320 threads): warps 0-3 producers (one pixel row each), warps 4-7 epilogue,
// warp 8 barrier init / TMEM allocation / TMA of weights / MMA issue, warp 9 TMA of image
// patches (and of dZ in the weight gradient).
#include "gemm_params.h"
#include "host_common.h"
#include "ptx.cuh"

#include <string.h>

namespace rsu {

constexpr int kFirstThreads = 320;
constexpr int kFirstStages = 3;
constexpr int kFTW = 8, kFTH = 16;
constexpr uint32_t kFirstABytes = 16384;  // 128 rows x 128 B
constexpr int kImgStages = 6;             // image patches in flight
constexpr int kImgPitch = 40;             // floats per patch row: (8 + 2*2) pixels x 3, + up to 3 of
                                          // lead-in (TMA box starts must be 16-byte aligned)
constexpr uint32_t kImgStageBytes = 3200; // 40 * 20 * 4, 128-byte aligned

struct FirstParams {
  const float* img;  // [N, S, S, 3] fp32
  int N, S;
  const float* cw;  // device [3][3]: net0[m] = sum_c (x[c] - 0.5) * cw[c*3 + m] + cb[m]
  const float* cb;  // device [3]
  int d, oy, ox;  // dilation, window origin inside the image
  int Ho, Wo;     // output extent
  float keep;
  unsigned long long seed;
  int tiles_x, tiles_y;
  int cout;  // 64 or 128
  CUtensorMap img_map;  // 3-D (S*3, S, N) fp32, no swizzle, box {40, 16 + 2d, 1}
  int img_shift;        // (3 * ox) % 4: floats between the aligned box start and the first tap
  // forward
  CUtensorMap w_map;    // 2-D (64 k, cout) bf16, SWIZZLE_128B, box {64, cout}
  CUtensorMap out_map;  // 4-D (cout, Wo, Ho, N) bf16, box {64, 8, 16, 1}
  const float* bias;
  int relu;
  // weight gradient
  CUtensorMap g_map;  // 4-D dZ (cout, Wo, Ho, N) bf16, box {64, 8, 16, 1}
  float* dw;          // fp32 [>= 28 rows][ldo]
  int ldo;
};

// Counter-based uniform in [0,1) and the dropout scale: the same generator as elementwise.cu
// (rsu_dropout / rsu_dropout_mask), so a mask is a pure function of (seed, element index).
__device__ __forceinline__ float first_uniform01(unsigned long long seed, unsigned long long idx) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z = z ^ (z >> 31);
  return static_cast<float>(z >> 40) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ float first_keep_scale(unsigned long long seed, unsigned long long idx,
                                                  float keep) {
  return floorf(keep + first_uniform01(seed, idx)) / keep;
}

struct FirstColor {
  float w[9], b[3];
};
__device__ __forceinline__ FirstColor first_load_color(const FirstParams& p) {
  FirstColor c;
#pragma unroll
  for (int i = 0; i < 9; ++i) c.w[i] = p.cw != nullptr ? __ldg(p.cw + i) : (i % 4 == 0 ? 1.f : 0.f);
#pragma unroll
  for (int i = 0; i < 3; ++i) c.b[i] = p.cb != nullptr ? __ldg(p.cb + i) : 0.f;
  return c;
}

__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// The 27 raw image values under output pixel (ly, lx) of a tile, from its staged patch
// (row pitch kImgPitch floats; patch origin = the tile's first tap).
__device__ __forceinline__ void first_read_patch(uint32_t patch_addr, int ly, int lx, int d,
                                                 float (&raw)[27]) {  // patch_addr includes the lead-in
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const uint32_t a = patch_addr +
                       static_cast<uint32_t>(((ly + (t / 3) * d) * kImgPitch + (lx + (t % 3) * d) * 3) * 4);
    raw[t * 3 + 0] = ld_shared_f32(a);
    raw[t * 3 + 1] = ld_shared_f32(a + 4);
    raw[t * 3 + 2] = ld_shared_f32(a + 8);
  }
}

// One im2col row (pixel (y, x) of image n, raw image values in `raw`) -> chunks 0..3 of row m of
// the stage at `stage_addr`: colour adjust, dropout, bf16, 128-byte swizzle.
// IDENT: identity colour transform and no dropout (the folded first layer): col = raw - 0.5.
template <bool IDENT>
__device__ __forceinline__ void first_produce_row(const FirstParams& p, const FirstColor& cc,
                                                  const float (&raw)[27], uint32_t stage_addr, int m,
                                                  int n, int y, int x) {
  float col[32];
#pragma unroll
  for (int j = 28; j < 32; ++j) col[j] = 0.f;
  col[27] = 1.f;
  if (IDENT) {
#pragma unroll
    for (int k = 0; k < 27; ++k) col[k] = raw[k] - 0.5f;
  } else {
    const bool drop = p.keep < 1.0f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float a0 = raw[t * 3] - 0.5f, a1 = raw[t * 3 + 1] - 0.5f, a2 = raw[t * 3 + 2] - 0.5f;
#pragma unroll
      for (int mm = 0; mm < 3; ++mm) {
        float o = a0 * cc.w[0 * 3 + mm] + a1 * cc.w[1 * 3 + mm] + a2 * cc.w[2 * 3 + mm] + cc.b[mm];
        if (drop) {
          const int yy = y + p.oy + (t / 3) * p.d, xx = x + p.ox + (t % 3) * p.d;
          const long long pix = (1LL * n * p.S + yy) * p.S + xx;
          o *= first_keep_scale(p.seed, static_cast<unsigned long long>(pix * 3 + mm), p.keep);
        }
        col[t * 3 + mm] = o;
      }
    }
  }
  const uint32_t row = stage_addr + static_cast<uint32_t>(m) * 128u;
  const uint32_t swz = static_cast<uint32_t>(m & 7);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 v;
    v.x = pack_bf16x2(col[q * 8 + 0], col[q * 8 + 1]);
    v.y = pack_bf16x2(col[q * 8 + 2], col[q * 8 + 3]);
    v.z = pack_bf16x2(col[q * 8 + 4], col[q * 8 + 5]);
    v.w = pack_bf16x2(col[q * 8 + 6], col[q * 8 + 7]);
    st_shared_v4(row + ((static_cast<uint32_t>(q) ^ swz) << 4), v);
  }
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* map, uint32_t bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ------------------------------------------------------------------------------- forward
// smem: [A stages 3 x 16 KiB][weights cout x 128 B][out staging 2 x 16 KiB][image ring 6 x 3 KiB]
//       [barriers, bias]
template <bool IDENT>
__global__ void __launch_bounds__(kFirstThreads, 2)
    first_conv_kernel(const __grid_constant__ FirstParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t w_base = smem_base + kFirstStages * kFirstABytes;
  const uint32_t stg_base = w_base + static_cast<uint32_t>(p.cout) * 128u;
  const uint32_t img_base = stg_base + 2u * 16384u;
  const uint32_t bar_base = img_base + kImgStages * kImgStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kFirstStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kFirstStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kFirstStages + 2 + a); };
  const uint32_t wfull_bar = bar_base + 8u * (2 * kFirstStages + 4);
  auto ifull_bar = [&](int s) { return bar_base + 8u * (2 * kFirstStages + 5 + s); };
  auto iempty_bar = [&](int s) { return bar_base + 8u * (2 * kFirstStages + 5 + kImgStages + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kFirstStages + 5 + 2 * kImgStages);
  const uint32_t bias_base = tmem_slot + 16u;  // float [cout]
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  float* bias_s = reinterpret_cast<float*>(smem_gen + (bias_base - smem_base));

  const uint32_t tmem_cols = p.cout <= 64 ? 128u : 256u;  // two accumulator stages
  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&p.w_map);
      tma_prefetch_desc(&p.out_map);
      tma_prefetch_desc(&p.img_map);
      for (int s = 0; s < kFirstStages; ++s) {
        mbar_init(full_bar(s), 128);
        mbar_init(empty_bar(s), 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(tfull_bar(a), 1);
        mbar_init(tempty_bar(a), 4);
      }
      mbar_init(wfull_bar, 1);
      for (int s = 0; s < kImgStages; ++s) {
        mbar_init(ifull_bar(s), 1);
        mbar_init(iempty_bar(s), 128);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, tmem_cols);
  }
  for (int j = threadIdx.x; j < p.cout; j += kFirstThreads)
    bias_s[j] = p.bias != nullptr ? __ldg(p.bias + j) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int total_tiles = p.N * tiles_per_img;
  const uint32_t img_tx = static_cast<uint32_t>(kImgPitch * (kFTH + 2 * p.d) * 4);

  if (warp < 4) {
    // ------------------------------------------------------------ producers
    const int m = threadIdx.x;
    const int ly = m >> 3, lx = m & 7;
    const FirstColor cc = first_load_color(p);
    uint32_t stage = 0, phase = 0, is = 0, iphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int tx = tile % p.tiles_x;
      const int ty = (tile / p.tiles_x) % p.tiles_y;
      const int n = tile / tiles_per_img;
      float raw[27];
      mbar_wait(ifull_bar(is), iphase);
      first_read_patch(img_base + is * kImgStageBytes + p.img_shift * 4, ly, lx, p.d, raw);
      mbar_arrive(iempty_bar(is));
      if (++is == kImgStages) {
        is = 0;
        iphase ^= 1u;
      }
      mbar_wait(empty_bar(stage), phase ^ 1u);
      first_produce_row<IDENT>(p, cc, raw, smem_base + stage * kFirstABytes, m, n, ty * kFTH + ly,
                               tx * kFTW + lx);
      fence_proxy_async();
      mbar_arrive(full_bar(stage));
      if (++stage == kFirstStages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------ TMA of image patches
    // (out-of-range rows / columns of edge tiles arrive as zeros; their pixels are clipped by
    // the TMA store)
    uint32_t is = 0, iphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int tx = tile % p.tiles_x;
      const int ty = (tile / p.tiles_x) % p.tiles_y;
      const int n = tile / tiles_per_img;
      mbar_wait(iempty_bar(is), iphase ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(ifull_bar(is), img_tx);
        tma_load_3d(img_base + is * kImgStageBytes, &p.img_map, ifull_bar(is),
                    (tx * kFTW + p.ox) * 3 - p.img_shift, ty * kFTH + p.oy, n);
      }
      __syncwarp();
      if (++is == kImgStages) {
        is = 0;
        iphase ^= 1u;
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------ weights + MMA issue
    if (elect_one()) {
      mbar_expect_tx(wfull_bar, static_cast<uint32_t>(p.cout) * 128u);
      tma_load_2d(w_base, &p.w_map, wfull_bar, 0, 0);
    }
    __syncwarp();
    mbar_wait(wfull_bar, 0);
    const uint32_t idesc = make_idesc_bf16(kBlockM, p.cout, false, false);
    const uint32_t hi = desc_hi_sw128(1024u);
    const uint32_t b_lo = desc_lo_sw128(w_base, 16);
    uint32_t stage = 0, phase = 0, it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      mbar_wait(full_bar(stage), phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_lo = desc_lo_sw128(smem_base + stage * kFirstABytes, 16);
        const uint32_t d_tmem = tmem_base + acc * static_cast<uint32_t>(p.cout);
        // K = 32: two 16-element slices (chunks 0..3 of every row); chunks 4..7 are never read
        umma_bf16_lohi(d_tmem, a_lo, hi, b_lo, hi, idesc, 0u);
        umma_bf16_lohi(d_tmem, a_lo + 2u, hi, b_lo + 2u, hi, idesc, 1u);
        umma_commit(empty_bar(stage));
        umma_commit(tfull_bar(acc));
      }
      __syncwarp();
      if (++stage == kFirstStages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 4-7)
    // 64-channel x 128-pixel blocks staged in shared memory (128-byte swizzle), TMA stores
    const int wq = warp & 3;
    const int m = wq * 32 + lane;
    const bool leader = threadIdx.x == 128;
    const uint32_t row_off = static_cast<uint32_t>(m) * 128u;
    const uint32_t swz = static_cast<uint32_t>(m & 7);
    const int cbs = p.cout / 64;
    uint32_t it = 0;
    long long q = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
      const int tx = tile % p.tiles_x;
      const int ty = (tile / p.tiles_x) % p.tiles_y;
      const int n = tile / tiles_per_img;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) +
                             acc * static_cast<uint32_t>(p.cout);
      for (int cb = 0; cb < cbs; ++cb, ++q) {
        const uint32_t buf = static_cast<uint32_t>(q & 1);
        uint32_t r0[32], r1[32];
        tmem_ld32(t_row + cb * 64, r0);
        tmem_ld32(t_row + cb * 64 + 32, r1);
        tmem_ld_wait();
        // barrier A: the leader has seen the store that last used this staging buffer finish
        // reading it
        named_bar_sync(1, 128);
        uint4 packed[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j)
            v[j] = __uint_as_float(h == 0 ? r0[j] : r1[j]) + bias_s[cb * 64 + h * 32 + j];
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            packed[h * 4 + c].x = pack_bf16x2(v[c * 8 + 0], v[c * 8 + 1]);
            packed[h * 4 + c].y = pack_bf16x2(v[c * 8 + 2], v[c * 8 + 3]);
            packed[h * 4 + c].z = pack_bf16x2(v[c * 8 + 4], v[c * 8 + 5]);
            packed[h * 4 + c].w = pack_bf16x2(v[c * 8 + 6], v[c * 8 + 7]);
          }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
          st_shared_v4(stg_base + buf * 16384u + row_off + ((static_cast<uint32_t>(c) ^ swz) << 4),
                       packed[c]);
        fence_proxy_async();
        named_bar_sync(2, 128);  // barrier B: block complete in shared memory
        if (leader) {
          tma_store_4d(&p.out_map, stg_base + buf * 16384u, cb * 64, tx * kFTW, ty * kFTH, n);
          tma_store_commit();
          tma_store_wait_read<1>();  // the other staging buffer is free again
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
    if (leader) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------- weight gradient
// One accumulator [128 x cout] per CTA (rows 0..27 are used: M atom 0 = the im2col channels,
// atom 1 = a block of zeros), reduced into `dw` with fp32 atomics when the CTA has walked its
// pixel tiles.
// smem: [stages x (A 16 KiB + dZ atoms cout/64 x 16 KiB)][zeros 2 KiB][image ring][barriers]
template <bool IDENT>
__global__ void __launch_bounds__(kFirstThreads, 2)
    first_wgrad_kernel(const __grid_constant__ FirstParams p, int stages) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_b = p.cout / 64;
  const uint32_t stage_bytes = kFirstABytes * static_cast<uint32_t>(1 + n_b);
  const uint32_t zero_base = smem_base + stages * stage_bytes;
  const uint32_t img_base = zero_base + 2048u;
  const uint32_t bar_base = img_base + kImgStages * kImgStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * stages);
  auto ifull_bar = [&](int s) { return bar_base + 8u * (2 * stages + 1 + s); };
  auto iempty_bar = [&](int s) { return bar_base + 8u * (2 * stages + 1 + kImgStages + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * stages + 1 + 2 * kImgStages);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  // zero the A halves of every stage once (chunks 4..7 of each row stay zero for the whole
  // kernel: they are M rows 32..63 of the MN-major operand) and the zeros block
  {
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int s = 0; s < stages; ++s)
      for (int i = threadIdx.x; i < static_cast<int>(kFirstABytes / 16); i += kFirstThreads)
        st_shared_v4(smem_base + s * stage_bytes + i * 16, z);
    for (int i = threadIdx.x; i < 2048 / 16; i += kFirstThreads) st_shared_v4(zero_base + i * 16, z);
    fence_proxy_async();
  }
  const uint32_t tmem_cols = p.cout <= 64 ? 64u : 128u;
  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&p.g_map);
      tma_prefetch_desc(&p.img_map);
      for (int s = 0; s < stages; ++s) {
        mbar_init(full_bar(s), 129);  // 128 producer rows + the arrive.expect_tx of the dZ loads
        mbar_init(empty_bar(s), 1);
      }
      mbar_init(tfull_bar, 1);
      for (int s = 0; s < kImgStages; ++s) {
        mbar_init(ifull_bar(s), 1);
        mbar_init(iempty_bar(s), 128);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int total_tiles = p.N * tiles_per_img;
  const uint32_t img_tx = static_cast<uint32_t>(kImgPitch * (kFTH + 2 * p.d) * 4);

  if (warp < 4) {
    // ------------------------------------------------------------ producers
    // (rows past the ragged edge meet zero-filled dZ rows -- TMA out-of-bounds fill -- and are
    // themselves built from zero-filled image values: finite, so they contribute nothing)
    const int m = threadIdx.x;
    const int ly = m >> 3, lx = m & 7;
    const FirstColor cc = first_load_color(p);
    uint32_t stage = 0, phase = 0, is = 0, iphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int tx = tile % p.tiles_x;
      const int ty = (tile / p.tiles_x) % p.tiles_y;
      const int n = tile / tiles_per_img;
      float raw[27];
      mbar_wait(ifull_bar(is), iphase);
      first_read_patch(img_base + is * kImgStageBytes + p.img_shift * 4, ly, lx, p.d, raw);
      mbar_arrive(iempty_bar(is));
      if (++is == kImgStages) {
        is = 0;
        iphase ^= 1u;
      }
      mbar_wait(empty_bar(stage), phase ^ 1u);
      // clamp only what the dropout hash indexes (edge rows are multiplied by zero dZ anyway)
      first_produce_row<IDENT>(p, cc, raw, smem_base + stage * stage_bytes, m, n,
                               min(ty * kFTH + ly, p.Ho - 1), min(tx * kFTW + lx, p.Wo - 1));
      fence_proxy_async();
      mbar_arrive(full_bar(stage));
      if (++stage == static_cast<uint32_t>(stages)) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------ TMA: image patches and dZ
    uint32_t stage = 0, phase = 0, is = 0, iphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int tx = tile % p.tiles_x;
      const int ty = (tile / p.tiles_x) % p.tiles_y;
      const int n = tile / tiles_per_img;
      mbar_wait(iempty_bar(is), iphase ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(ifull_bar(is), img_tx);
        tma_load_3d(img_base + is * kImgStageBytes, &p.img_map, ifull_bar(is),
                    (tx * kFTW + p.ox) * 3 - p.img_shift, ty * kFTH + p.oy, n);
      }
      __syncwarp();
      if (++is == kImgStages) {
        is = 0;
        iphase ^= 1u;
      }
      mbar_wait(empty_bar(stage), phase ^ 1u);
      if (elect_one()) {
        const uint32_t b_addr = smem_base + stage * stage_bytes + kFirstABytes;
        mbar_expect_tx(full_bar(stage), kFirstABytes * static_cast<uint32_t>(n_b));
        for (int j = 0; j < n_b; ++j)
          tma_load_4d(b_addr + kFirstABytes * j, &p.g_map, full_bar(stage), j * 64, tx * kFTW,
                      ty * kFTH, n);
      }
      __syncwarp();
      if (++stage == static_cast<uint32_t>(stages)) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------ MMA issue
    const uint32_t idesc = make_idesc_bf16(kBlockM, p.cout, true, true);
    const uint32_t hi = desc_hi_sw128(1024u);
    const uint32_t zero16 = (zero_base >> 4) & 0x3FFFu;
    uint32_t stage = 0, phase = 0;
    bool first = true;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(full_bar(stage), phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a16 = ((smem_base + stage * stage_bytes) >> 4) & 0x3FFFu;
        const uint32_t b_lo = desc_lo_sw128(smem_base + stage * stage_bytes + kFirstABytes, kFirstABytes);
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // 16 pixels (K) per instruction
          const uint32_t s16 = a16 + j * 128u;
          // second M atom = the zeros block: leading byte offset = its distance from this slice
          const uint32_t a_lo = s16 | ((zero16 - s16) << 16);
          umma_bf16_lohi(tmem_base, a_lo, hi, b_lo + j * 128u, hi, idesc, (j != 0 || !first) ? 1u : 0u);
        }
        umma_commit(empty_bar(stage));
      }
      __syncwarp();
      first = false;
      if (++stage == static_cast<uint32_t>(stages)) {
        stage = 0;
        phase ^= 1u;
      }
    }
    if (elect_one()) umma_commit(tfull_bar);
    __syncwarp();
  } else if (warp == 4) {
    // ------------------------------------------------------------ epilogue: rows 0..27 -> dw
    if (blockIdx.x < total_tiles) {
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
      for (int ch = 0; ch < p.cout / 32; ++ch) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ch * 32, r);
        tmem_ld_wait();
        if (lane < 28) {
          float* orow = p.dw + static_cast<long long>(lane) * p.ldo + ch * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            red_add_v4(orow + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                       __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
      }
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, tmem_cols);
}

// 3-D fp32 map over the image rows: dims (S*3, S, N), box {box_w floats, box_h rows, 1}, no swizzle.
static int encode_image_map(CUtensorMap* map, const float* img, int N, int S, int box_w, int box_h) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return set_error(RSU_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  cuuint64_t dims[3] = {(cuuint64_t)S * 3, (cuuint64_t)S, (cuuint64_t)N};
  cuuint64_t strides[2] = {(cuuint64_t)S * 12, (cuuint64_t)S * S * 12};
  cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(img), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(RSU_ECUDA, "cuTensorMapEncodeTiled(image S=%d N=%d box %dx%d) -> %d", S, N, box_w,
                     box_h, (int)r);
  return RSU_OK;
}

static int first_fill_params(FirstParams* p, const float* img, int N, int S, const float* cw,
                             const float* cb, int dilation, int oy, int ox, int Ho, int Wo, int cout,
                             float keep, unsigned long long seed) {
  if (!img || N < 1 || S < 1) return set_error(RSU_EINVAL, "first_conv: empty image");
  if (dilation < 1 || oy < 0 || ox < 0 || Ho < 1 || Wo < 1 || oy + Ho + 2 * dilation > S ||
      ox + Wo + 2 * dilation > S)
    return set_error(RSU_EINVAL, "first_conv: window outside the %dx%d image", S, S);
  if (cout != 64 && cout != 128) return set_error(RSU_EINVAL, "first_conv: cout %d (64 or 128)", cout);
  if (!(keep > 0.f && keep <= 1.f)) return set_error(RSU_EINVAL, "first_conv: keep=%f", keep);
  memset(p, 0, sizeof(*p));
  p->img = img;
  p->N = N;
  p->S = S;
  if ((cw == nullptr) != (cb == nullptr))
    return set_error(RSU_EINVAL, "first_conv: cw and cb must both be given or both be null");
  if (!cw && keep < 1.f)
    return set_error(RSU_EINVAL, "first_conv: dropout needs the explicit colour transform");
  p->cw = cw;
  p->cb = cb;
  p->d = dilation;
  p->oy = oy;
  p->ox = ox;
  p->Ho = Ho;
  p->Wo = Wo;
  p->keep = keep;
  p->seed = seed;
  p->tiles_x = (Wo + kFTW - 1) / kFTW;
  p->tiles_y = (Ho + kFTH - 1) / kFTH;
  p->cout = cout;
  if (1LL * N * p->tiles_x * p->tiles_y > 0x7fffffffLL)
    return set_error(RSU_EINVAL, "first_conv: too many tiles");
  if (dilation > 2) return set_error(RSU_EINVAL, "first_conv: dilation %d > 2", dilation);
  // image as a 3-D fp32 tensor (S*3 floats, S rows, N images): TMA needs 16-byte strides
  if ((reinterpret_cast<uintptr_t>(img) & 15) || (S * 12) % 16 != 0)
    return set_error(RSU_EALIGN, "first_conv: image must be 16-byte aligned with S %% 4 == 0 (S = %d)", S);
  p->img_shift = (3 * ox) & 3;  // (tx * 8 + ox) * 3 floats = 96 tx bytes + 12 ox bytes
  return encode_image_map(&p->img_map, img, N, S, kImgPitch, kFTH + 2 * dilation);
}

}  // namespace rsu

using namespace rsu;

extern "C" {

int rsu_first_conv_fwd(const float* img, int N, int S, const float* cw, const float* cb,
                       int dilation, int oy, int ox, const void* w_packed, const float* bias,
                       int relu, const rsu_view* out, float keep, unsigned long long seed,
                       void* stream) {
  if (!out || !w_packed) return set_error(RSU_EINVAL, "first_conv_fwd: null argument");
  FirstParams p;
  int rc = first_fill_params(&p, img, N, S, cw, cb, dilation, oy, ox, out->H, out->W,
                             out->C, keep, seed);
  if (rc) return rc;
  if (out->N != N) return set_error(RSU_EINVAL, "first_conv_fwd: batch %d != %d", out->N, N);
  if (out->W < kFTW || out->H < kFTH) return set_error(RSU_EINVAL, "first_conv_fwd: output smaller than one tile");
  rc = encode_weight_map(&p.w_map, w_packed, 64, p.cout, p.cout);
  if (rc) return rc;
  rc = encode_act_map(&p.out_map, *out, kFTW, kFTH);
  if (rc) return rc;
  p.bias = bias;
  p.relu = relu;
  const int smem = 1024 + kFirstStages * kFirstABytes + p.cout * 128 + 2 * 16384 +
                   kImgStages * kImgStageBytes + 8 * (2 * kFirstStages + 5 + 2 * kImgStages) + 16 +
                   p.cout * 4;
  static bool attr_set = false;
  if (!attr_set) {
    RSU_CHECK_CUDA(cudaFuncSetAttribute(first_conv_kernel<true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    RSU_CHECK_CUDA(cudaFuncSetAttribute(first_conv_kernel<false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const long long total = 1LL * N * p.tiles_x * p.tiles_y;
  long long grid = 2LL * num_sms();  // two CTAs per SM fit (about 110 KiB of shared memory each)
  if (grid > total) grid = total;
  if (cw == nullptr)
    first_conv_kernel<true><<<static_cast<int>(grid), kFirstThreads, smem, (cudaStream_t)stream>>>(p);
  else
    first_conv_kernel<false><<<static_cast<int>(grid), kFirstThreads, smem, (cudaStream_t)stream>>>(p);
  return check_launch("first_conv_kernel");
}

int rsu_first_conv_wgrad(const float* img, int N, int S, const float* cw, const float* cb,
                         int dilation, int oy, int ox, const rsu_view* dz, float* dw, int ldo,
                         float keep, unsigned long long seed, void* stream) {
  if (!dz || !dw) return set_error(RSU_EINVAL, "first_conv_wgrad: null argument");
  if (ldo < dz->C || (ldo % 4) || (reinterpret_cast<uintptr_t>(dw) & 15))
    return set_error(RSU_EALIGN, "first_conv_wgrad: dw must be 16-byte aligned with ldo %% 4 == 0");
  FirstParams p;
  int rc = first_fill_params(&p, img, N, S, cw, cb, dilation, oy, ox, dz->H, dz->W, dz->C,
                             keep, seed);
  if (rc) return rc;
  if (dz->N != N) return set_error(RSU_EINVAL, "first_conv_wgrad: batch %d != %d", dz->N, N);
  if (dz->W < kFTW || dz->H < kFTH) return set_error(RSU_EINVAL, "first_conv_wgrad: dZ smaller than one tile");
  rc = encode_act_map(&p.g_map, *dz, kFTW, kFTH);
  if (rc) return rc;
  p.dw = dw;
  p.ldo = ldo;
  const int stage_bytes = kFirstABytes * (1 + p.cout / 64);
  // two CTAs per SM (the producers, not HBM, pace one CTA): 2 stages of 32 KiB at cout = 64
  int stages = (88 * 1024) / stage_bytes;
  if (stages < 2) stages = 2;
  if (stages > 4) stages = 4;
  const int smem = 1024 + stages * stage_bytes + 2048 + kImgStages * kImgStageBytes +
                   8 * (2 * stages + 1 + 2 * kImgStages) + 32;
  static bool attr_set = false;
  if (!attr_set) {
    RSU_CHECK_CUDA(cudaFuncSetAttribute(first_wgrad_kernel<true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    RSU_CHECK_CUDA(cudaFuncSetAttribute(first_wgrad_kernel<false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const long long total = 1LL * N * p.tiles_x * p.tiles_y;
  long long grid = (smem <= 113 * 1024 ? 2LL : 1LL) * num_sms();
  if (grid > total) grid = total;
  if (cw == nullptr)
    first_wgrad_kernel<true><<<static_cast<int>(grid), kFirstThreads, smem, (cudaStream_t)stream>>>(p, stages);
  else
    first_wgrad_kernel<false><<<static_cast<int>(grid), kFirstThreads, smem, (cudaStream_t)stream>>>(p, stages);
  return check_launch("first_wgrad_kernel");
}

}  // extern "C"
