// Halo-tile weight-gradient GEMM on tcgen05 tensor cores (sm_100a).
//
//   D[(tap, src, c), co] += sum over pixels of  X_src[pixel + tap + off, c] * G[pixel + goff, co]
//
// As in wgrad_gemm.cu both operands are MN-major with the pixel index as GEMM K, but here one
// unit of work is (64-channel chunk of X, 64-channel tile of G, slice of the pixel tiles) and it
// produces ALL taps at once: per 8 x 16 pixel tile ONE TMA box {64 ch, 8 + span_x, 16 + span_y}
// of X (the halo tile) is loaded and the operand of tap (dy, dx) is the window of that tile that
// starts at smem row dy*Wh + dx (8-pixel-wide tiles make every 8-row K group one image row, so
// groups are uniformly Wh*128 bytes apart; the 128-byte swizzle is anchored at absolute
// addresses, see tools/diag_swizzle.cu).  Two taps share one 128-row accumulator (the second
// 64-row atom is "leading-byte-offset" away, which may be any tap-to-tap distance), so nine taps
// need five accumulators x 64 columns of tensor memory.  The spare atom of the fifth accumulator
// points at a constant block of ones: its rows are the column sums of G, i.e. BiasAddGrad for
// free.  L2 -> SM traffic per 128 pixels: 22.5 KiB + 16 KiB instead of 9/2 x (32 + 16) KiB.
//
//
// N = 64 instructions are shared-memory-operand bound (48 instead of 32 cycles, tools/mma_probe.cu).
// When Cout is a multiple of 128 the gradient tile is 128 channels wide instead; five accumulators
// x 128 columns no longer fit the 512 columns of tensor memory, so the taps are split into two
// groups -- accumulators {0,1} (taps 0..3) and {2,3,4} (taps 4..8 + ones) -- that are separate
// units of work (each re-reads the X halo tile and the G tile from L2, which is not the limit).
//
// Reference op replaced: Conv2DBackpropFilter + BiasAddGrad of the 3x3 convolutions in
// src/unet.py:34-45, 88-91.
#include "gemm_params.h"
#include "host_common.h"
#include "ptx.cuh"

namespace rsu {

constexpr int kWhThreads = 256;
constexpr int kWhTW = 8, kWhTH = 16;
constexpr int kWhBBytes = 16384;  // 128 pixels x 64 channels of G
constexpr int kOnesBytes = 4096;

struct WgradHaloParams {
  CUtensorMap a_map[kMaxSrc];  // 4-D (C, W, H, N) bf16, box {64, Wh, Hh, 1}
  CUtensorMap b_map;           // 4-D (C, W, H, N) bf16, box {64, 8, 16, 1}
  int n_src;
  int src_chunks[kMaxSrc];
  int src_off_y[kMaxSrc];  // crop offset + halo origin
  int src_off_x[kMaxSrc];
  int n_taps;
  int tap_row[kMaxTaps];  // ascending
  int b_off_y, b_off_x;
  int Wh, Hh;
  int tiles_x, tiles_y, n_img;
  int chunks_total, n_tiles_n, ksplit;
  int BN;      // gradient channels per unit: 64 or 128
  int groups;  // 1: all accumulators in one unit; 2: accumulators {0,1} and {2,3,4} (n_taps == 9)
  int stages;
  uint32_t a_stage_bytes;  // 1024-aligned
  float* out;
  int ldo;
  float* bias_grad;  // may be null
};

__global__ void __launch_bounds__(kWhThreads, 1)
    wgrad_halo_kernel(const __grid_constant__ WgradHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = p.stages;
  const int n_b_atoms = p.BN / 64;
  const uint32_t stage_bytes = p.a_stage_bytes + static_cast<uint32_t>(n_b_atoms) * kWhBBytes;
  const uint32_t ones_base = smem_base + stages * stage_bytes;
  const uint32_t bar_base = ones_base + kOnesBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * stages);
  const uint32_t tempty_bar = bar_base + 8u * (2 * stages + 1);
  const uint32_t tmem_slot = bar_base + 8u * (2 * stages + 2);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  // constant block of bf16 ones (the "second atom" of the last accumulator)
  {
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem_gen + (ones_base - smem_base));
    for (int i = threadIdx.x; i < kOnesBytes / 4; i += kWhThreads) ones[i] = 0x3F803F80u;
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.n_src; ++s) tma_prefetch_desc(&p.a_map[s]);
    tma_prefetch_desc(&p.b_map);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 4);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int pix_tiles = p.n_img * tiles_per_img;
  const int total_units = p.chunks_total * p.n_tiles_n * p.groups * p.ksplit;
  const int n_mtiles = (p.n_taps + 1) / 2;
  const uint32_t a_bytes = static_cast<uint32_t>(p.Wh * p.Hh) * 128u;

  // unit -> (ks, tap group, n_tile, chunk); chunk fastest so that concurrently running CTAs share
  // G and X.  i0 / i1: the unit's accumulator range.
  auto unit_range = [&](int unit, int* cg, int* n_tile, int* pt_begin, int* pt_end, int* i0,
                        int* i1) {
    *cg = unit % p.chunks_total;
    int rest = unit / p.chunks_total;
    *n_tile = rest % p.n_tiles_n;
    rest /= p.n_tiles_n;
    const int grp = rest % p.groups;
    const int ks = rest / p.groups;
    *i0 = (p.groups == 2 && grp == 1) ? 2 : 0;
    *i1 = (p.groups == 2 && grp == 0) ? 2 : n_mtiles;
    *pt_begin = static_cast<int>(1LL * pix_tiles * ks / p.ksplit);
    *pt_end = static_cast<int>(1LL * pix_tiles * (ks + 1) / p.ksplit);
  };
  auto chunk_src = [&](int cg, int* s, int* c) {
    int ss = 0;
    while (ss < p.n_src - 1 && cg >= p.src_chunks[ss]) {
      cg -= p.src_chunks[ss];
      ++ss;
    }
    *s = ss;
    *c = cg;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // (whole warp converged, one elected lane issues; see conv_gemm.cu)
    uint32_t stage = 0, phase = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      int cg, n_tile, pt0, pt1, s, c, i0, i1;
      unit_range(unit, &cg, &n_tile, &pt0, &pt1, &i0, &i1);
      chunk_src(cg, &s, &c);
      const int ox = p.src_off_x[s], oy = p.src_off_y[s];
      int img = pt0 / tiles_per_img;
      int r = pt0 % tiles_per_img;
      int ty = r / p.tiles_x, tx = r % p.tiles_x;
      for (int pt = pt0; pt < pt1; ++pt) {
        const int y0 = ty * kWhTH, x0 = tx * kWhTW;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (elect_one()) {
          const uint32_t dst = smem_base + stage * stage_bytes;
          const uint32_t fb = full_bar(stage);
          mbar_expect_tx(fb, a_bytes + static_cast<uint32_t>(n_b_atoms) * kWhBBytes);
          tma_load_4d(dst, &p.a_map[s], fb, c * 64, x0 + ox, y0 + oy, img);
          for (int a = 0; a < n_b_atoms; ++a)
            tma_load_4d(dst + p.a_stage_bytes + a * kWhBBytes, &p.b_map, fb, n_tile * p.BN + a * 64,
                        x0 + p.b_off_x, y0 + p.b_off_y, img);
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
    // Descriptors as (lo, hi) words.  lo = start address >> 4 | LBO >> 4 << 16: the LBO of
    // accumulator i is the distance between its two taps' windows (the ones block for the odd
    // tap), the start address advances by two image rows of the halo tile per 16-pixel K step.
    const uint32_t idesc = make_idesc_bf16(kBlockM, p.BN, true, true);
    const uint32_t bn = static_cast<uint32_t>(p.BN);
    const uint32_t a_hi = desc_hi_sw128(static_cast<uint32_t>(p.Wh) * 128u);
    const uint32_t b_hi = desc_hi_sw128(1024u);
    constexpr int kMaxM = (kMaxTaps + 1) / 2;
    uint32_t tile_lo[kMaxM];  // (tap window offset >> 4) | (LBO >> 4) << 16, without stage base
#pragma unroll
    for (int i = 0; i < kMaxM; ++i) {
      tile_lo[i] = 0;
      if (i < n_mtiles) {
        const uint32_t row0 = static_cast<uint32_t>(p.tap_row[2 * i]);
        const uint32_t lbo16 = (2 * i + 1 < p.n_taps)
                                   ? static_cast<uint32_t>(p.tap_row[2 * i + 1] - p.tap_row[2 * i]) * 8u
                                   : 0u;  // filled per K step below (ones block)
        tile_lo[i] = row0 * 8u + (lbo16 << 16);
      }
    }
    const bool odd = (p.n_taps & 1) != 0;
    const uint32_t kstep16 = static_cast<uint32_t>(2 * p.Wh) * 8u;
    uint32_t stage = 0, phase = 0;
    uint32_t unit_it = 0;
    const uint32_t stage16 = stage_bytes >> 4;
    const uint32_t a16_base = (smem_base >> 4) & 0x3FFFu;
    const uint32_t b_off16 = p.a_stage_bytes >> 4;
    const uint32_t ones16 = (ones_base >> 4) & 0x3FFFu;
    const uint32_t b_lbo = static_cast<uint32_t>(kWhBBytes >> 4) << 16;
    // per-K-step advance of every accumulator's A descriptor: + kstep16 in the start address; the
    // odd tile's second atom is the ones block, whose distance (LBO field) shrinks by as much
    const uint32_t d_pair = kstep16;
    const uint32_t d_last = odd ? kstep16 - (kstep16 << 16) : kstep16;
    uint32_t a16 = a16_base;  // start address >> 4 of the current stage
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x, ++unit_it) {
      int cg, n_tile, pt0, pt1, i0, i1;
      unit_range(unit, &cg, &n_tile, &pt0, &pt1, &i0, &i1);
      mbar_wait(tempty_bar, (unit_it & 1u) ^ 1u);
      tc_fence_after();
      for (int pt = pt0; pt < pt1; ++pt) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t acc = pt != pt0 ? 1u : 0u;
          const uint32_t b_lo0 = a16 + b_off16 + b_lbo;
          if (p.n_taps == 9) {
            // 5 accumulators: 4 tap pairs + (tap 8, ones); everything but a16 is loop invariant.
            // One fully unrolled instruction stream per accumulator range (no predicated-off
            // tcgen05.mma: they would still occupy the instruction queue).
            uint32_t a_lo[5];
#pragma unroll
            for (int i = 0; i < 4; ++i) a_lo[i] = a16 + tile_lo[i];
            {
              const uint32_t x = a16 + (tile_lo[4] & 0xFFFFu);
              a_lo[4] = x + ((ones16 - x) << 16);
            }
            if (i0 == 0 && i1 == 5) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {  // 16 pixels (two image rows of the tile) per instruction
#pragma unroll
                for (int i = 0; i < 5; ++i)
                  umma_bf16_lohi(tmem_base + i * bn, a_lo[i] + j * (i == 4 ? d_last : d_pair), a_hi,
                                 b_lo0 + j * 128u, b_hi, idesc, j != 0 ? 1u : acc);
              }
            } else if (i0 == 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
#pragma unroll
                for (int i = 0; i < 2; ++i)
                  umma_bf16_lohi(tmem_base + i * bn, a_lo[i] + j * d_pair, a_hi, b_lo0 + j * 128u, b_hi,
                                 idesc, j != 0 ? 1u : acc);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
#pragma unroll
                for (int i = 2; i < 5; ++i)
                  umma_bf16_lohi(tmem_base + (i - 2) * bn, a_lo[i] + j * (i == 4 ? d_last : d_pair), a_hi,
                                 b_lo0 + j * 128u, b_hi, idesc, j != 0 ? 1u : acc);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t b_lo = b_lo0 + j * 128u;
#pragma unroll
              for (int i = 0; i < kMaxM; ++i) {
                if (i < n_mtiles) {
                  uint32_t a_lo = a16 + tile_lo[i] + j * kstep16;
                  if (odd && i == n_mtiles - 1) a_lo += (ones16 - (a_lo & 0x3FFFu)) << 16;
                  umma_bf16_lohi(tmem_base + i * bn, a_lo, a_hi, b_lo, b_hi, idesc, j != 0 ? 1u : acc);
                }
              }
            }
          }
          umma_commit(empty_bar(stage));
        }
        __syncwarp();
        a16 += stage16;
        if (++stage == static_cast<uint32_t>(stages)) {
          stage = 0;
          phase ^= 1u;
          a16 = a16_base;
        }
      }
      if (elect_one()) umma_commit(tfull_bar);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue: atomics to fp32
    const int wq = warp & 3;
    const int m = wq * 32 + lane;
    uint32_t unit_it = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x, ++unit_it) {
      int cg, n_tile, pt0, pt1, i0, i1;
      unit_range(unit, &cg, &n_tile, &pt0, &pt1, &i0, &i1);
      mbar_wait(tfull_bar, unit_it & 1u);
      tc_fence_after();
      const bool nonempty = pt1 > pt0;
      for (int i = i0; i < i1; ++i) {
        const int tap = 2 * i + (m >> 6);
        const bool is_w = tap < p.n_taps && nonempty;
        const bool is_b = tap == p.n_taps && m == 64 && p.bias_grad != nullptr && cg == 0 && nonempty;
        float* orow = is_b ? p.bias_grad + n_tile * p.BN
                           : p.out + (static_cast<long long>(tap * p.chunks_total + cg) * 64 + (m & 63)) * p.ldo +
                                 n_tile * p.BN;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + (i - i0) * p.BN;
        for (int ch = 0; ch < p.BN / 32; ++ch) {
          uint32_t r[32];
          tmem_ld32(t_row + ch * 32, r);
          tmem_ld_wait();
          if (is_w || is_b) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              red_add_v4(orow + ch * 32 + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                         __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// Returns RSU_OK after launching, -1 when the shape is not eligible, or an RSU_E* code.
int launch_wgrad_halo(const rsu_wgrad_desc* d, cudaStream_t stream, int* bias_done) {
  *bias_done = 0;
  if (d->n_taps < 2) return -1;
  if (d->grad.C % 64 != 0) return -1;
  int min_dy = d->tap_dy[0], max_dy = d->tap_dy[0], min_dx = d->tap_dx[0], max_dx = d->tap_dx[0];
  for (int t = 1; t < d->n_taps; ++t) {
    min_dy = d->tap_dy[t] < min_dy ? d->tap_dy[t] : min_dy;
    max_dy = d->tap_dy[t] > max_dy ? d->tap_dy[t] : max_dy;
    min_dx = d->tap_dx[t] < min_dx ? d->tap_dx[t] : min_dx;
    max_dx = d->tap_dx[t] > max_dx ? d->tap_dx[t] : max_dx;
  }
  const int span_y = max_dy - min_dy, span_x = max_dx - min_dx;
  if (span_x > 8 || span_y > 8) return -1;

  WgradHaloParams p;
  memset(&p, 0, sizeof(p));
  p.Wh = kWhTW + span_x;
  p.Hh = kWhTH + span_y;
  p.a_stage_bytes = static_cast<uint32_t>((p.Wh * p.Hh * 128 + 1023) & ~1023);
  p.n_taps = d->n_taps;
  for (int t = 0; t < d->n_taps; ++t) {
    p.tap_row[t] = (d->tap_dy[t] - min_dy) * p.Wh + (d->tap_dx[t] - min_dx);
    if (t > 0 && p.tap_row[t] <= p.tap_row[t - 1]) return -1;  // LBO must be positive
  }
  if (d->grad.W < kWhTW || d->grad.H < kWhTH) return -1;
  p.n_src = d->n_src;
  for (int s = 0; s < d->n_src; ++s) {
    const rsu_view& v = d->src[s];
    if (v.W < p.Wh || v.H < p.Hh) return -1;
    int rc = encode_act_map(&p.a_map[s], v, p.Wh, p.Hh);
    if (rc) return rc;
    p.src_chunks[s] = v.C / 64;
    p.chunks_total += v.C / 64;
    p.src_off_y[s] = v.off_y + min_dy;
    p.src_off_x[s] = v.off_x + min_dx;
  }
  {
    int rc = encode_act_map(&p.b_map, d->grad, kWhTW, kWhTH);
    if (rc) return rc;
  }
  p.b_off_y = d->grad.off_y;
  p.b_off_x = d->grad.off_x;
  p.tiles_x = (d->W + kWhTW - 1) / kWhTW;
  p.tiles_y = (d->H + kWhTH - 1) / kWhTH;
  p.n_img = d->N_img;
  // 128-channel gradient tiles (two tap groups) when the channel count allows it
  p.BN = (d->grad.C % 128 == 0 && d->n_taps == 9) ? 128 : 64;
  p.groups = p.BN == 128 ? 2 : 1;
  p.n_tiles_n = d->grad.C / p.BN;
  p.out = d->out;
  p.ldo = d->ldo;
  p.bias_grad = (d->n_taps % 2 == 1) ? d->bias_grad : nullptr;

  const int pix_tiles = p.n_img * p.tiles_x * p.tiles_y;
  const int mn_units = p.chunks_total * p.n_tiles_n * p.groups;
  const int sms = num_sms();
  // cycles per 128-pixel tile: 8 K steps x accumulators MMAs of 128 x BN x 16 (48 cycles at BN = 64,
  // shared-memory operand bandwidth bound; 64 at BN = 128, 2.5 accumulators per unit on average);
  // the single accumulator set is drained by fp32 reductions before the next unit may start
  const double tile_cost = p.BN == 128 ? 8.0 * 2.5 * 64.0 : 8.0 * ((d->n_taps + 1) / 2) * 48.0;
  p.ksplit = choose_ksplit(mn_units, pix_tiles, sms, tile_cost, p.BN == 128 ? 9000.0 : 6000.0, 8);

  const int stage_bytes = static_cast<int>(p.a_stage_bytes) + (p.BN / 64) * kWhBBytes;
  int stages = (227 * 1024 - 1024 - kOnesBytes - 512) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return -1;
  p.stages = stages;
  const int smem = 1024 + stages * stage_bytes + kOnesBytes + 8 * (2 * stages + 2) + 16;
  static bool attr_set = false;
  if (!attr_set) {
    RSU_CHECK_CUDA(cudaFuncSetAttribute(wgrad_halo_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const long long total = 1LL * mn_units * p.ksplit;
  int grid = total < sms ? static_cast<int>(total) : sms;
  wgrad_halo_kernel<<<grid, kWhThreads, smem, stream>>>(p);
  *bias_done = p.bias_grad != nullptr ? 1 : 0;
  return check_launch("wgrad_halo_kernel");
}

}  // namespace rsu
