// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   D[pixel, n] = sum over K = (tap, source, 64-channel chunk) of  A[pixel, k] * B[n, k]
//
// A rows are output pixels of one TW x TH spatial tile (<= 128 = TMEM lanes); for every K step
// one 4-D TMA box {64 ch, TW, TH, 1} is fetched at the tap-/crop-shifted coordinate, so im2col
// never exists in memory and out-of-range pixels arrive as zeros (that is the "full" padding
// of the data gradient).  B is the packed weight matrix [Ntot][Ktot] (K contiguous).  Both
// operands land in shared memory with the 128-byte swizzle and are consumed by
// tcgen05.mma.kind::f16 (bf16 x bf16 -> fp32) with the accumulator in tensor memory; two
// accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
//
// Warp roles (256 threads, 1 CTA / SM, persistent over tiles):
//   warp 0: TMA producer     warp 1: MMA issuer     warp 2: TMEM allocator
//   warps 4-7: epilogue (TMEM -> registers -> bias / ReLU / mask / accumulate -> global)
//
// Reference ops replaced: tf.layers.conv2d (+dilation_rate) fwd and Conv2DBackpropInput,
// tf.layers.conv2d_transpose fwd/bwd, crop + concat (src/unet.py:34-45, 67-91).
#include <stdlib.h>

#include "gemm_params.h"
#include "host_common.h"
#include "ptx.cuh"

namespace rsu {

constexpr int kConvThreads = 256;
constexpr int kAStageBytes = kBlockM * 128;  // 16 KiB
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;  // TMEM columns per accumulator stage

__global__ void __launch_bounds__(kConvThreads, 1)
    conv_gemm_kernel(const __grid_constant__ ConvGemmParams p, int stages) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t b_stage_bytes = static_cast<uint32_t>(p.BN) * 128u;
  const uint32_t stage_bytes = kAStageBytes + b_stage_bytes;
  const uint32_t bar_base = smem_base + stages * stage_bytes;
  // barrier layout: full[stages], empty[stages], tmem_full[2], tmem_empty[2]
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * stages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * stages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * stages + 4);
  const uint32_t bias_base = tmem_slot + 16u;  // float [2][256]
  const uint32_t ktab_base = bias_base + 2u * 256u * 4u;  // int4 [num_k]
  // generic pointers to the same carve-up
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  float* bias_s = reinterpret_cast<float*>(smem_gen + (bias_base - smem_base));
  int4* ktab = reinterpret_cast<int4*>(smem_gen + (ktab_base - smem_base));

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.n_src; ++s) tma_prefetch_desc(&p.a_map[s]);
    tma_prefetch_desc(&p.b_map);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int total_tiles = p.n_img * tiles_per_img * p.n_tiles_n;
  int chunks_total = 0;
  for (int s = 0; s < p.n_src; ++s) chunks_total += p.src_chunks[s];
  const int num_k = p.n_taps * chunks_total;
  const uint32_t a_bytes = static_cast<uint32_t>(p.TW * p.TH) * 128u;
  // K-step table (tile independent): {source index, first channel, x offset, y offset}.  The
  // single-warp producer loop is issue-latency bound, so everything that does not depend on the
  // tile is looked up with one 16-byte shared-memory load instead of being recomputed.
  for (int k = threadIdx.x; k < num_k; k += kConvThreads) {
    const int t = k / chunks_total;
    int cg = k % chunks_total, s = 0;
    while (s < p.n_src - 1 && cg >= p.src_chunks[s]) cg -= p.src_chunks[s++];
    ktab[k] = make_int4(s, cg * kBlockK, p.tap_dx[t] + p.src_off_x[s], p.tap_dy[t] + p.src_off_y[s]);
  }
  __syncthreads();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // (whole warp converged, one elected lane issues: keeps the loop on the uniform datapath)
    uint32_t stage = 0, phase = 0;
    const uint32_t tx_bytes = a_bytes + b_stage_bytes;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles_n;
      const int rest = tile / p.n_tiles_n;
      const int tx = rest % p.tiles_x;
      const int ty = (rest / p.tiles_x) % p.tiles_y;
      const int img = rest / tiles_per_img;
      const int x0 = tx * p.TW, y0 = ty * p.TH, n0 = n_tile * p.BN;
      for (int k = 0; k < num_k; ++k) {
        const int4 e = ktab[k];
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (elect_one()) {
          const uint32_t a_dst = smem_base + stage * stage_bytes;
          const uint32_t fb = full_bar(stage);
          mbar_expect_tx(fb, tx_bytes);
          tma_load_4d(a_dst, &p.a_map[e.x], fb, e.y, x0 + e.z, y0 + e.w, img);
          tma_load_2d(a_dst + kAStageBytes, &p.b_map, fb, k * kBlockK, n0);
        }
        __syncwarp();
        if (++stage == static_cast<uint32_t>(stages)) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc_bf16(kBlockM, p.BN, false, false);
    uint32_t stage = 0, phase = 0;
    uint32_t acc_it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++acc_it) {
      const uint32_t acc = acc_it & 1u;
      const uint32_t acc_phase = (acc_it >> 1) & 1u;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kAccStride;
      for (int k = 0; k < num_k; ++k) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_base + stage * stage_bytes;
          const uint64_t adesc = make_smem_desc_sw128(a_addr, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(a_addr + kAStageBytes, 16, 1024);
#pragma unroll
          for (int j = 0; j < kBlockK / 16; ++j) {
            // advancing 16 bf16 (32 B) along K inside the 128-byte swizzle row: +2 in >>4 units
            umma_bf16(d_tmem, adesc + 2u * j, bdesc + 2u * j, idesc, (k | j) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
        }
        __syncwarp();
        if (++stage == static_cast<uint32_t>(stages)) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (elect_one()) umma_commit(tfull_bar(acc));
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int wq = warp & 3;
    const int m = wq * 32 + lane;
    const int ly = m / p.TW, lx = m % p.TW;
    const bool in_tile = m < p.TW * p.TH;
    const int et = threadIdx.x - 128;
    uint32_t acc_it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++acc_it) {
      const uint32_t acc = acc_it & 1u;
      const uint32_t acc_phase = (acc_it >> 1) & 1u;
      const int n_tile = tile % p.n_tiles_n;
      const int rest = tile / p.n_tiles_n;
      const int tx = rest % p.tiles_x;
      const int ty = (rest / p.tiles_x) % p.tiles_y;
      const int img = rest / tiles_per_img;
      const int n0 = n_tile * p.BN;
      const int y = ty * p.TH + ly, x = tx * p.TW + lx;
      const bool valid = in_tile && y < p.H_out && x < p.W_out;

      float* bias_t = bias_s + acc * 256;
      if (p.bias != nullptr) {
        for (int j = et; j < p.BN; j += 128) {
          const int n = n0 + j;
          bias_t[j] = __ldg(p.bias + (p.shuffle_cout > 0 ? n % p.shuffle_cout : n));
        }
      }
      // ReLU-gradient mask bits of this tile, fetched while the MMAs are still running
      uint32_t mbits[2][4];
      // (a partial mask covers whole N tiles: mask_c0 / mask_nc are multiples of BN)
      const bool tile_masked =
          p.mask != nullptr && (p.mask_nc == 0 || (n0 >= p.mask_c0 && n0 < p.mask_c0 + p.mask_nc));
      if (tile_masked) {
        const __nv_bfloat16* mpx =
            p.mask + img * p.mask_sn + y * p.mask_sy + x * p.mask_sx + (n0 - p.mask_c0);
        load_mask_bits4(mpx, p.BN / 32, valid, mbits[0]);
        load_mask_bits4(mpx + 128, p.BN / 32 - 4, valid, mbits[1]);
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      named_bar_sync(1, 128);  // bias_t visible to the 4 epilogue warps

      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * kAccStride;
      for (int ch = 0; ch < p.BN / 32; ++ch) {
        uint32_t r[32];
        tmem_ld32(t_row + ch * 32, r);
        tmem_ld_wait();
        if (valid) {
          const int n = n0 + ch * 32;
          long long off;
          if (p.shuffle_cout > 0) {
            const int ab = n / p.shuffle_cout, co = n % p.shuffle_cout;
            off = img * p.out_sn + (2 * y + (ab >> 1)) * p.out_sy + (2 * x + (ab & 1)) * p.out_sx +
                  co;
          } else {
            off = img * p.out_sn + y * p.out_sy + x * p.out_sx + n;
          }
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (p.bias != nullptr) {
            add_bias32(v, bias_t + ch * 32);
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (tile_masked) {
            uint32_t mb = 0;
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (c == ch) mb = mbits[c >> 2][c & 3];
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (!((mb >> j) & 1u)) v[j] = 0.f;
          }
          uint4* op = reinterpret_cast<uint4*>(p.out + off);
          if (p.accumulate) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 ov = op[q];
              const uint32_t w[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                v[q * 8 + 2 * e] += bf16_lo(w[e]);
                v[q * 8 + 2 * e + 1] += bf16_hi(w[e]);
              }
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            o.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
            o.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
            o.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
            o.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
            op[q] = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------ host launcher
static int conv_smem_bytes(int BN, int num_k, int* stages_out) {
  const int stage_bytes = kAStageBytes + BN * 128;
  const int fixed = 1024 /*align slack*/ + 8 * (2 * 8 + 4) + 16 + 2 * 256 * 4 + 16 * num_k /*k-step table*/;
  int stages = (226 * 1024 - fixed) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  *stages_out = stages;
  return fixed + stages * stage_bytes;
}

int launch_conv_halo(const rsu_conv_gemm_desc* d, cudaStream_t stream, bool forced);  // conv_halo.cu
int launch_conv_halo2(const rsu_conv_gemm_desc* d, cudaStream_t stream);               // conv_halo2.cu
int launch_conv_gemm2(ConvGemmParams& p, const void* weights, int ktot, int ntot,
                      cudaStream_t stream);  // conv_gemm2.cu

}  // namespace rsu

using namespace rsu;

extern "C" int rsu_conv_gemm(const rsu_conv_gemm_desc* d, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d) return set_error(RSU_EINVAL, "null descriptor");
  if (d->n_src < 1 || d->n_src > kMaxSrc) return set_error(RSU_EINVAL, "n_src %d", d->n_src);
  if (d->n_taps < 1 || d->n_taps > kMaxTaps) return set_error(RSU_EINVAL, "n_taps %d", d->n_taps);
  if (d->H_out < 1 || d->W_out < 1 || d->N_img < 1)
    return set_error(RSU_EINVAL, "empty output %dx%dx%d", d->N_img, d->H_out, d->W_out);
  if (d->Ntot % 64 != 0) return set_error(RSU_EINVAL, "Ntot %d not a multiple of 64", d->Ntot);
  if ((reinterpret_cast<uintptr_t>(d->out) & 15) || (d->out_sx % 8) || (d->out_sy % 8) ||
      (d->out_sn % 8))
    return set_error(RSU_EALIGN, "output pointer/strides must be 16-byte aligned");
  if (d->mask && ((reinterpret_cast<uintptr_t>(d->mask) & 15) || (d->mask_sx % 8) ||
                  (d->mask_sy % 8) || (d->mask_sn % 8)))
    return set_error(RSU_EALIGN, "mask pointer/strides must be 16-byte aligned");
  if (d->shuffle_cout > 0 && (d->Ntot != 4 * d->shuffle_cout || d->shuffle_cout % 32 != 0))
    return set_error(RSU_EINVAL, "shuffle_cout %d inconsistent with Ntot %d", d->shuffle_cout,
                     d->Ntot);

  if (d->pool_done_host) *d->pool_done_host = 0;
  // Halo-tile kernel for the layers whose per-tap operand traffic (L2 -> SM) bounds them: few
  // output channels per tile and a large pixel grid.
  // (thresholds from the per-layer A/B table, profiles/r2_halo_experiments.txt)
  const bool want_halo =
      d->algo == 2 ||
      (d->algo == 0 && d->n_taps == 9 && d->H_out >= 64 && d->W_out >= 64 &&
       (d->Ntot <= 192 || (d->Ntot <= 384 && d->Ntot % 128 == 0 && d->H_out >= 190)));
  // CTA-pair halo kernel (conv_halo2.cu) wherever the halo kernel is wanted and the pair's halved
  // weight footprint fits in shared memory: x1.0 - 2.1 on the N = 64 / 128 layers of the flagship
  // network (tools/bench_halo_pair.py, profiles/r2_halo_pair_ab.txt); explicit with algo 4.
  // RSU_HALO_PAIR=0 disables.
  static const bool halo_pair_auto = [] {
    const char* e = getenv("RSU_HALO_PAIR");
    return !(e && e[0] == '0');
  }();
  if (d->algo == 4 || (want_halo && d->algo == 0 && halo_pair_auto)) {
    const int rc = launch_conv_halo2(d, stream);
    if (rc >= 0) return rc;
    if (d->algo == 4) return set_error(RSU_EINVAL, "shape not eligible for the CTA-pair halo kernel");
  }
  if (want_halo) {
    const int rc = launch_conv_halo(d, stream, d->algo == 2);
    if (rc >= 0) return rc;
    if (d->algo == 2) return set_error(RSU_EINVAL, "shape not eligible for the halo-tile kernel");
  }

  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  int max_tw = d->W_out, max_th = d->H_out;
  int ktot = 0;
  for (int s = 0; s < d->n_src; ++s) {
    const rsu_view& v = d->src[s];
    if (v.N != d->N_img) return set_error(RSU_EINVAL, "source %d batch %d != %d", s, v.N, d->N_img);
    if (v.W < max_tw) max_tw = v.W;
    if (v.H < max_th) max_th = v.H;
    ktot += v.C;
  }
  ktot *= d->n_taps;
  int TW, TH;
  pick_tile(d->W_out, d->H_out, max_tw, max_th, false, &TW, &TH);
  if (TW < 1 || TH < 1) return set_error(RSU_EINVAL, "no valid tile for %dx%d", d->W_out, d->H_out);
  p.TW = TW;
  p.TH = TH;
  p.tiles_x = (d->W_out + TW - 1) / TW;
  p.tiles_y = (d->H_out + TH - 1) / TH;
  p.n_img = d->N_img;
  p.BN = d->Ntot % 256 == 0 ? 256 : (d->Ntot % 128 == 0 ? 128 : 64);
  p.n_tiles_n = d->Ntot / p.BN;
  p.H_out = d->H_out;
  p.W_out = d->W_out;
  p.n_src = d->n_src;
  for (int s = 0; s < d->n_src; ++s) {
    int rc = encode_act_map(&p.a_map[s], d->src[s], TW, TH);
    if (rc) return rc;
    p.src_chunks[s] = d->src[s].C / 64;
    p.src_off_y[s] = d->src[s].off_y;
    p.src_off_x[s] = d->src[s].off_x;
  }
  p.n_taps = d->n_taps;
  for (int t = 0; t < d->n_taps; ++t) {
    p.tap_dy[t] = d->tap_dy[t];
    p.tap_dx[t] = d->tap_dx[t];
  }
  {
    int rc = encode_weight_map(&p.b_map, d->weights, ktot, d->Ntot, p.BN);
    if (rc) return rc;
  }
  p.out = static_cast<__nv_bfloat16*>(d->out);
  p.out_sn = d->out_sn;
  p.out_sy = d->out_sy;
  p.out_sx = d->out_sx;
  p.shuffle_cout = d->shuffle_cout;
  p.bias = d->bias;
  p.relu = d->relu;
  p.mask = static_cast<const __nv_bfloat16*>(d->mask);
  p.mask_sn = d->mask_sn;
  p.mask_sy = d->mask_sy;
  p.mask_sx = d->mask_sx;
  p.mask_c0 = d->mask_nc > 0 ? d->mask_c0 : 0;
  p.mask_nc = d->mask_nc;
  if (d->mask && d->mask_nc > 0 && (d->mask_c0 % p.BN || d->mask_nc % p.BN || d->mask_c0 < 0 ||
                                    d->mask_c0 + d->mask_nc > d->Ntot))
    return set_error(RSU_EINVAL, "mask channel range [%d, +%d) not a multiple of the N tile %d", d->mask_c0,
                     d->mask_nc, p.BN);
  p.accumulate = d->accumulate;
  // CTA pairs (conv_gemm2.cu) wherever this kernel would run with the 256-wide N tile: 3-13 %
  // faster on every such layer of the flagship network (tools/bench_pair.py,
  // profiles/r1_pair_ab.txt); at N = 128 the single-CTA kernels win.  RSU_CONV_PAIR=0 disables.
  static const bool pair_auto = [] {
    const char* e = getenv("RSU_CONV_PAIR");
    return !(e && e[0] == '0');
  }();
  if (d->algo == 3 || (d->algo == 0 && pair_auto && p.BN == 256))
    return launch_conv_gemm2(p, d->weights, ktot, d->Ntot, stream);

  int stages;
  const int smem = conv_smem_bytes(p.BN, ktot / kBlockK, &stages);
  if (smem > 227 * 1024) return set_error(RSU_EINVAL, "K = %d too deep for the k-step table", ktot);
  static bool attr_set = false;
  if (!attr_set) {
    RSU_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const long long total = 1LL * p.n_img * p.tiles_x * p.tiles_y * p.n_tiles_n;
  int grid = num_sms();
  if (total < grid) grid = static_cast<int>(total);
  conv_gemm_kernel<<<grid, kConvThreads, smem, stream>>>(p, stages);
  return check_launch("conv_gemm_kernel");
}
