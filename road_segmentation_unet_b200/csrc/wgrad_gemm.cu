// Weight-gradient GEMM on tcgen05 tensor cores (sm_100a).
//
//   D[(tap, src, c), co] += sum over pixels of  X_src[pixel + tap + off, c] * G[pixel + goff, co]
//
// The reduction (GEMM K) dimension is the pixel index, which is the slow axis of both NHWC
// operands, so both are consumed as MN-major: a TMA box {64 ch, TW, TH, 1} lands as TW*TH rows
// of 128 B and is exactly the canonical "MN-major, SWIZZLE_128B" atom stack (8-row K groups
// 1024 B apart).  One accumulator = 2 row atoms (128 weight rows: two (tap, 64-channel) blocks)
// x BN gradient channels; 16 pixels are retired per tcgen05.mma.  Pixel tiles are divided among
// `ksplit` CTAs whose partial sums are combined with fp32 atomics into the zero-initialised
// output, which has TensorFlow's HWIO kernel layout [(tap, cin), cout].
// Out-of-range pixels of either operand arrive as zeros from TMA and contribute nothing.
// BiasAddGrad (the column sums of G) rides along: one more 64-row atom whose operand is a constant
// block of ones in shared memory -- the free second atom of the last row tile when the atom count
// is odd, else one extra row tile -- so no separate pass over G is needed.
//
// Reference op replaced: Conv2DBackpropFilter of every tf.layers.conv2d / conv2d_transpose in
// src/unet.py:34-45, 67, 88-91.
#include "gemm_params.h"
#include "host_common.h"
#include "ptx.cuh"

#include <stdlib.h>

namespace rsu {

constexpr int kWgThreads = 256;
constexpr int kWgOnesBytes = 4096;  // two 16-row K slices of bf16 ones
constexpr int kWgTmemCols = 512;
constexpr int kWgAccStride = 256;

struct AtomCoord {
  int src, chunk, dy, dx;
};

__device__ __forceinline__ AtomCoord decode_atom(const WgradParams& p, int atom, int chunks_total) {
  AtomCoord a;
  const int tap = atom / chunks_total;
  int cg = atom % chunks_total;
  int s = 0;
  while (s < p.n_src - 1 && cg >= p.src_chunks[s]) {
    cg -= p.src_chunks[s];
    ++s;
  }
  a.src = s;
  a.chunk = cg;
  a.dy = p.tap_dy[tap] + p.src_off_y[s];
  a.dx = p.tap_dx[tap] + p.src_off_x[s];
  return a;
}

__global__ void __launch_bounds__(kWgThreads, 1)
    wgrad_gemm_kernel(const __grid_constant__ WgradParams p, int stages, uint32_t atom_bytes,
                      float* bias_grad) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int n_b_atoms = p.BN / 64;
  const uint32_t stage_bytes = static_cast<uint32_t>(2 + n_b_atoms) * atom_bytes;
  const uint32_t ones_base = smem_base + stages * stage_bytes;
  const uint32_t bar_base = ones_base + kWgOnesBytes;
  // (with a bias gradient the atom list is one longer: atom n_atoms = the ones block)
  const int n_atoms_all = p.n_atoms + (bias_grad != nullptr ? 1 : 0);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * stages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * stages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * stages + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.n_src; ++s) tma_prefetch_desc(&p.a_map[s]);
    tma_prefetch_desc(&p.b_map);
  }
  {
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem_raw + (ones_base - smem_u32(smem_raw)));
    for (int i = threadIdx.x; i < kWgOnesBytes / 4; i += kWgThreads) ones[i] = 0x3F803F80u;
    fence_proxy_async();
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 2);  // the A-operand and the B-operand producer warp
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kWgTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  int chunks_total = 0;
  for (int s = 0; s < p.n_src; ++s) chunks_total += p.src_chunks[s];
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int pix_tiles = p.n_img * tiles_per_img;
  const int total_units = p.n_tiles_m * p.n_tiles_n * p.ksplit;
  const int tile_pixels = p.TW * p.TH;
  const uint32_t box_bytes = static_cast<uint32_t>(tile_pixels) * 128u;
  const int mma_per_tile = tile_pixels / 16;

  // unit -> (ks, n_tile, m_tile); m fastest so that concurrently running CTAs share the G tile
  auto unit_range = [&](int unit, int* m_tile, int* n_tile, int* pt_begin, int* pt_end) {
    *m_tile = unit % p.n_tiles_m;
    const int rest = unit / p.n_tiles_m;
    *n_tile = rest % p.n_tiles_n;
    const int ks = rest / p.n_tiles_n;
    *pt_begin = static_cast<int>(1LL * pix_tiles * ks / p.ksplit);
    *pt_end = static_cast<int>(1LL * pix_tiles * (ks + 1) / p.ksplit);
  };

  if (warp == 0 || warp == 3) {
    // ------------------------------------------------------------ TMA producers
    // warp 0 loads the (up to) two A atoms of a stage, warp 3 its BN/64 B atoms: one warp issuing
    // all six boxes of a 64-pixel stage was busy three quarters of the time (ncu stall samples)
    // and could not run far enough ahead of the MMAs.  Each warp posts its own expect_tx.
    // (whole warp converged, one elected lane issues; see conv_gemm.cu)
    const bool is_a = warp == 0;
    uint32_t stage = 0, phase = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      int m_tile, n_tile, pt0, pt1;
      unit_range(unit, &m_tile, &n_tile, &pt0, &pt1);
      const int atom0 = m_tile * 2;
      // real (TMA-loaded) atoms of this row tile: 2, 1 (odd tail, or + the ones atom) or 0 (the
      // ones atom alone)
      const int n_a = atom0 + 1 < p.n_atoms ? 2 : (atom0 < p.n_atoms ? 1 : 0);
      const AtomCoord a0 = decode_atom(p, n_a >= 1 ? atom0 : 0, chunks_total);
      const AtomCoord a1 = decode_atom(p, n_a == 2 ? atom0 + 1 : 0, chunks_total);
      const int n0 = n_tile * p.BN;
      const uint32_t tx_bytes = static_cast<uint32_t>(is_a ? n_a : n_b_atoms) * box_bytes;
      const int bx = p.b_off_x, by = p.b_off_y;
      int img = pt0 / tiles_per_img;
      int r = pt0 % tiles_per_img;
      int ty = r / p.tiles_x, tx = r % p.tiles_x;
      for (int pt = pt0; pt < pt1; ++pt) {
        const int y0 = ty * p.TH, x0 = tx * p.TW;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (elect_one()) {
          const uint32_t dst = smem_base + stage * stage_bytes;
          const uint32_t fb = full_bar(stage);
          mbar_expect_tx(fb, tx_bytes);
          if (is_a) {
            if (n_a >= 1)
              tma_load_4d(dst, &p.a_map[a0.src], fb, a0.chunk * 64, x0 + a0.dx, y0 + a0.dy, img);
            if (n_a == 2)
              tma_load_4d(dst + atom_bytes, &p.a_map[a1.src], fb, a1.chunk * 64, x0 + a1.dx, y0 + a1.dy,
                          img);
          } else {
            for (int j = 0; j < n_b_atoms; ++j)
              tma_load_4d(dst + (2 + j) * atom_bytes, &p.b_map, fb, n0 + j * 64, x0 + bx, y0 + by, img);
          }
        }
        __syncwarp();
        if (++stage == static_cast<uint32_t>(stages)) {
          stage = 0;
          phase ^= 1u;
        }
        if (++tx == p.tiles_x) {
          tx = 0;
          if (++ty == p.tiles_y) {
            ty = 0;
            ++img;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc_bf16(kBlockM, p.BN, true, true);
    const uint32_t hi = desc_hi_sw128(1024u);
    const uint32_t ones16 = (ones_base >> 4) & 0x3FFFu;
    uint32_t stage = 0, phase = 0;
    uint32_t acc_it = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x, ++acc_it) {
      int m_tile, n_tile, pt0, pt1;
      unit_range(unit, &m_tile, &n_tile, &pt0, &pt1);
      const uint32_t acc = acc_it & 1u;
      const uint32_t acc_phase = (acc_it >> 1) & 1u;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kWgAccStride;
      // 0: no ones atom in this row tile, 1: second atom = ones, 2: first atom = ones
      const int ones_kind = bias_grad == nullptr ? 0
                            : (m_tile * 2 + 1 == p.n_atoms ? 1 : (m_tile * 2 == p.n_atoms ? 2 : 0));
      for (int pt = pt0; pt < pt1; ++pt) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_base + stage * stage_bytes;
          // A descriptor (lo word) of K slice 0 and its advance per 16-pixel slice:
          //   two real atoms / one real atom: start = stage, LBO = atom stride, + 2 KiB per slice
          //   real atom + ones atom: LBO = distance to the ones block (shrinks as the start advances)
          //   ones atom alone: start = ones block, never advances
          uint32_t a_lo, a_step;
          if (ones_kind == 0) {
            a_lo = desc_lo_sw128(a_addr, atom_bytes);
            a_step = 128u;
          } else if (ones_kind == 1) {
            const uint32_t s16 = (a_addr >> 4) & 0x3FFFu;
            a_lo = s16 | ((ones16 - s16) << 16);
            a_step = 128u - (128u << 16);
          } else {
            a_lo = ones16 | (128u << 16);
            a_step = 0u;
          }
          const uint32_t b_lo = desc_lo_sw128(a_addr + 2 * atom_bytes, atom_bytes);
          const uint32_t first = pt != pt0 ? 1u : 0u;
          // 16 pixels (K) per instruction = 16 rows of 128 B = 2 KiB further into every atom.
          // (separate loops per tile size: predicated-off tcgen05.mma still take a slot of the
          // instruction queue, and the 64-pixel stage is issue-paced)
          if (mma_per_tile == 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              umma_bf16_lohi(d_tmem, a_lo + j * a_step, hi, b_lo + j * 128u, hi, idesc, j != 0 ? 1u : first);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (j < mma_per_tile)
                umma_bf16_lohi(d_tmem, a_lo + j * a_step, hi, b_lo + j * 128u, hi, idesc, j != 0 ? 1u : first);
            }
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
    // ------------------------------------------------------------ epilogue: atomics to fp32
    const int wq = warp & 3;
    const int m = wq * 32 + lane;
    uint32_t acc_it = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x, ++acc_it) {
      int m_tile, n_tile, pt0, pt1;
      unit_range(unit, &m_tile, &n_tile, &pt0, &pt1);
      const uint32_t acc = acc_it & 1u;
      const uint32_t acc_phase = (acc_it >> 1) & 1u;
      const int atom = m_tile * 2 + (m >> 6);
      // the ones atom: its 64 rows all hold the column sums of G; row 0 goes to the bias gradient
      const bool is_bias = bias_grad != nullptr && atom == p.n_atoms && (m & 63) == 0;
      const bool valid = (atom < p.n_atoms || is_bias) && pt1 > pt0;
      float* orow = is_bias ? bias_grad + n_tile * p.BN
                            : p.out + (static_cast<long long>(atom) * 64 + (m & 63)) * p.ldo + n_tile * p.BN;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row =
          tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * kWgAccStride;
      for (int ch = 0; ch < p.BN / 32; ++ch) {
        uint32_t r[32];
        tmem_ld32(t_row + ch * 32, r);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            red_add_v4(orow + ch * 32 + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                       __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kWgTmemCols);
}

int launch_wgrad_halo(const rsu_wgrad_desc* d, cudaStream_t stream, int* bias_done);  // wgrad_halo.cu
int launch_wgrad_gemm2(WgradParams& p, float* bias_grad, cudaStream_t stream);         // wgrad_gemm2.cu

// pixels per K step of the per-tap kernel (RSU_WGRAD_TILE = 32 | 64 | 128 overrides, for A/B runs)
static int wgrad_tile_pixels() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("RSU_WGRAD_TILE");
    v = (e && atoi(e) == 128) ? 128 : ((e && atoi(e) == 32) ? 32 : 64);
  }
  return v;
}

}  // namespace rsu

using namespace rsu;

extern "C" int rsu_wgrad_gemm(const rsu_wgrad_desc* d, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d) return set_error(RSU_EINVAL, "null descriptor");
  if (d->n_src < 1 || d->n_src > kMaxSrc) return set_error(RSU_EINVAL, "n_src %d", d->n_src);
  if (d->n_taps < 1 || d->n_taps > kMaxTaps) return set_error(RSU_EINVAL, "n_taps %d", d->n_taps);
  if (d->grad.C % 64 != 0) return set_error(RSU_EINVAL, "grad channels %d % 64 != 0", d->grad.C);
  if (d->H < 1 || d->W < 1 || d->N_img < 1) return set_error(RSU_EINVAL, "empty pixel grid");
  if (reinterpret_cast<uintptr_t>(d->out) & 15)
    return set_error(RSU_EALIGN, "wgrad output must be 16-byte aligned");

  if (d->bias_done_host) *d->bias_done_host = 0;
  // (thresholds from the per-layer A/B tables, profiles/r2_wgrad_pair_ab.txt; the halo kernel also
  // produces the bias gradient, which is counted in its favour)
  int chunks_all = 0;
  for (int s = 0; s < d->n_src && s < kMaxSrc; ++s) chunks_all += d->src[s].C / 64;
  // (with 128-channel gradient tiles the halo kernel is not shared-memory bound any more, which
  // also pays for up to 6 input chunks: conv_9/conv1, 384 -> 128 channels)
  const int max_chunks = (d->grad.C % 128 == 0) ? 6 : 4;
  const bool want_halo =
      d->algo == 2 || (d->algo == 0 && d->n_taps == 9 && d->H >= 64 && d->W >= 64 && chunks_all <= max_chunks &&
                       d->grad.C <= 128);
  if (want_halo) {
    int bias_done = 0;
    const int rc = launch_wgrad_halo(d, stream, &bias_done);
    if (rc >= 0) {
      if (rc == RSU_OK && d->bias_done_host) *d->bias_done_host = bias_done;
      return rc;
    }
    if (d->algo == 2) return set_error(RSU_EINVAL, "shape not eligible for the halo-tile kernel");
  }

  WgradParams p;
  memset(&p, 0, sizeof(p));
  int max_tw = d->W, max_th = d->H;
  if (d->grad.W < max_tw) max_tw = d->grad.W;
  if (d->grad.H < max_th) max_th = d->grad.H;
  int chunks_total = 0;
  for (int s = 0; s < d->n_src; ++s) {
    if (d->src[s].W < max_tw) max_tw = d->src[s].W;
    if (d->src[s].H < max_th) max_th = d->src[s].H;
    chunks_total += d->src[s].C / 64;
  }
  // 64-pixel K steps: a stage is (2 + BN/64) x 8 KiB, so four stages fit even at BN = 256 (with
  // 128-pixel steps only two 96 KiB stages did, and the tensor pipe idled ~45 % of the time
  // waiting for loads -- profiles/r1_step_metrics.txt)
  // ... and at BN <= 128 a 64-pixel stage holds only 4 x 64 cycles of tensor work, less than the
  // issue loop's own cost per stage (tools/mma_probe.cu): those shapes take 128-pixel stages.
  const int cout_ = d->grad.C;
  const int bn_ = cout_ % 256 == 0 ? 256 : (cout_ % 128 == 0 ? 128 : 64);
  int TW, TH;
  pick_tile(d->W, d->H, max_tw, max_th, true, &TW, &TH, bn_ <= 128 ? 128 : wgrad_tile_pixels());
  if (TW < 1 || TH < 1 || (TW * TH) % 16 != 0)
    return set_error(RSU_EINVAL, "no valid pixel tile for %dx%d (need TW*TH %% 16 == 0)", d->W, d->H);
  p.TW = TW;
  p.TH = TH;
  p.tiles_x = (d->W + TW - 1) / TW;
  p.tiles_y = (d->H + TH - 1) / TH;
  p.n_img = d->N_img;
  p.n_src = d->n_src;
  for (int s = 0; s < d->n_src; ++s) {
    int rc = encode_act_map(&p.a_map[s], d->src[s], TW, TH);
    if (rc) return rc;
    p.src_chunks[s] = d->src[s].C / 64;
    p.src_off_y[s] = d->src[s].off_y;
    p.src_off_x[s] = d->src[s].off_x;
  }
  {
    int rc = encode_act_map(&p.b_map, d->grad, TW, TH);
    if (rc) return rc;
  }
  p.b_off_y = d->grad.off_y;
  p.b_off_x = d->grad.off_x;
  p.n_taps = d->n_taps;
  for (int t = 0; t < d->n_taps; ++t) {
    p.tap_dy[t] = d->tap_dy[t];
    p.tap_dx[t] = d->tap_dx[t];
  }
  p.n_atoms = d->n_taps * chunks_total;
  float* bias_grad = d->bias_grad;  // column sums of G through one more (ones) atom
  if (bias_grad && (reinterpret_cast<uintptr_t>(bias_grad) & 15)) bias_grad = nullptr;
  p.n_tiles_m = (p.n_atoms + (bias_grad ? 1 : 0) + 1) / 2;
  const int cout = d->grad.C;
  p.BN = cout % 256 == 0 ? 256 : (cout % 128 == 0 ? 128 : 64);
  p.n_tiles_n = cout / p.BN;
  p.out = d->out;
  p.ldo = d->ldo;

  // CTA-pair kernel (tcgen05.mma.cta_group::2, wgrad_gemm2.cu): each CTA of a pair loads half of
  // the gradient tile.  Faster than the single-CTA kernel on every 3x3 layer with a 256-wide
  // gradient tile and a small pixel grid (levels 3..8 of the flagship net: x1.00 - 1.11,
  // profiles/r2_wgrad_pair_ab.txt); the bias gradient rides along as one more (ones) unit.
  if (d->algo == 3 || (d->algo == 0 && p.BN == 256 && d->n_taps == 9 && d->H <= 104 && d->W <= 104)) {
    p.n_tiles_m = (p.n_atoms + 1) / 2;
    const int rc = launch_wgrad_gemm2(p, bias_grad, stream);
    if (rc == RSU_OK && bias_grad && d->bias_done_host) *d->bias_done_host = 1;
    return rc;
  }

  const int pix_tiles = p.n_img * p.tiles_x * p.tiles_y;
  const int mn_units = p.n_tiles_m * p.n_tiles_n;
  // cycles: one K step = (pixels / 16) MMAs of BN / 2 cycles; a unit ends with a BN-column
  // epilogue of fp32 reductions that overlaps the next unit only partly
  p.ksplit = choose_ksplit(mn_units, pix_tiles, num_sms(), (TW * TH / 16) * (p.BN / 2.0),
                           1500.0 + 12.0 * p.BN, 4);

  const int atom_bytes = ((TW * TH * 128) + 1023) & ~1023;
  const int stage_bytes = (2 + p.BN / 64) * atom_bytes;
  int stages = (220 * 1024 - kWgOnesBytes) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  const int smem = 1024 + stages * stage_bytes + kWgOnesBytes + 8 * (2 * stages + 4) + 16;
  static bool attr_set = false;
  if (!attr_set) {
    RSU_CHECK_CUDA(cudaFuncSetAttribute(wgrad_gemm_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const long long total = 1LL * mn_units * p.ksplit;
  int grid = num_sms();
  if (total < grid) grid = static_cast<int>(total);
  wgrad_gemm_kernel<<<grid, kWgThreads, smem, stream>>>(p, stages, static_cast<uint32_t>(atom_bytes),
                                                        bias_grad);
  if (bias_grad && d->bias_done_host) *d->bias_done_host = 1;
  return check_launch("wgrad_gemm_kernel");
}
