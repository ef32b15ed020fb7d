// Halo-tile implicit-GEMM 3x3 convolution on tcgen05 tensor cores (sm_100a).
//
// Same contraction as conv_gemm.cu,
//   D[pixel, n] = sum over K = (tap, source, 64-channel chunk) of  A[pixel, k] * B[n, k],
// but the activation operand of ALL taps comes from ONE TMA box per (source, chunk): the halo
// tile {64 ch, 8 + span_x, 16*MT + span_y} of an 8-wide x 16*MT-tall block of output pixels.
// A 128-row MMA operand for tap (dy, dx) is the window of that tile that starts at smem row
// dy*Wh + dx: with an 8-pixel-wide block every 8-row swizzle group is one image row, so the
// groups are uniformly Wh*128 bytes apart (the descriptor's stride-byte-offset) and the 128-byte
// swizzle is a function of the absolute shared-memory address, so a window may start at any row
// (tools/diag_swizzle.cu is the hardware probe for this).  Per 64-channel chunk the L2 -> SM
// traffic of A drops from 9 x 16 KiB to 22.5 KiB per 128 pixels, which is what bounds the
// Cout <= 128 layers of the network (profiles/r1_layers.json, per-tap vs halo columns of profiles/r2_halo_experiments.txt).
//
// MT = 2 gives two accumulators (256 pixels) per weight tile, halving the B traffic; when the
// whole weight matrix of a CTA's N tile fits in shared memory next to the A ring it is loaded
// once and stays resident (each CTA then keeps one N tile for its whole life).
//
// Warp roles as in conv_gemm.cu: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4-7 epilogue (TMEM -> registers -> bias / ReLU / mask / accumulate -> global).
//
// Reference ops replaced: tf.layers.conv2d (+dilation_rate) forward and Conv2DBackpropInput of the
// 3x3 convolutions in src/unet.py:34-45, 88-91, including crop + concat (:70-85).
#include "gemm_params.h"
#include "host_common.h"
#include "ptx.cuh"

namespace rsu {

constexpr int kHaloThreads = 256;
constexpr int kHaloTW = 8;    // block width in pixels = rows of one swizzle group
constexpr int kHaloTH = 16;   // image rows per 128-row accumulator
constexpr int kMaxBStages = 28;

struct ConvHaloParams {
  CUtensorMap a_map[kMaxSrc];  // 4-D (C, W, H, N) bf16, SWIZZLE_128B, box {64, Wh, Hh, 1}
  CUtensorMap b_map;           // 2-D (Ktot, Ntot) bf16, SWIZZLE_128B, box {64, BN}
  int n_src;
  int src_chunks[kMaxSrc];
  int src_off_y[kMaxSrc];  // crop offset + halo origin (min tap offset) per source
  int src_off_x[kMaxSrc];
  int n_taps;
  int tap_row[kMaxTaps];  // first smem row of the tap's window inside the halo tile
  int Wh, Hh;             // halo tile extent in pixels
  int MT;                 // accumulators (16-row blocks) per unit: 1 or 2
  int tiles_x, tiles_y, n_img;
  int n_tiles_n, BN;
  int resident;           // weights of one N tile loaded once per CTA (unit -> fixed N tile)
  int stages_a, stages_b;
  uint32_t a_stage_bytes;  // 1024-aligned
  int H_out, W_out;
  __nv_bfloat16* out;
  long long out_sn, out_sy, out_sx;
  const float* bias;
  int relu;
  const __nv_bfloat16* mask;
  long long mask_sn, mask_sy, mask_sx;
  int mask_c0, mask_nc;  // mask_nc > 0: the mask covers output channels [mask_c0, +mask_nc) only
  int accumulate;
  // Coalesced epilogue (used unless accumulate): 64-channel x 128-pixel blocks are staged in
  // shared memory (128-byte swizzle) and written with TMA stores; the ReLU-gradient mask blocks
  // arrive the same way (TMA loads, two blocks ahead).
  int tma_epilogue;
  CUtensorMap out_map;   // 4-D (C, W, H, N) bf16, box {64, 8, 16, 1}
  CUtensorMap mask_map;  // same geometry over the mask tensor
  // fused 2x2 max pool of the ReLU output (TMA epilogue only): 64-channel x 32-pixel blocks
  int pool;
  CUtensorMap pool_map;  // 4-D (C, W/2, H/2, N) bf16, box {64, 4, 8, 1}
};

struct HaloIssueCtx {
  uint32_t smem_base, b_base, bar_base, tmem_base;
  int u_begin, u_step, u_end, chunks_total;
};

// The MMA-issuing warp of conv_halo_kernel.  NT / MT / BN are compile-time for the shapes the
// network uses (0 = read them from the parameter block): with runtime tap counts the unrolled
// loop re-read its bounds from constant memory and branched before every group of four
// tcgen05.mma, and the issuing thread -- not shared memory or the tensor pipe -- set the pace
// (ncu source view: the warp never waited on a full barrier).  Descriptors are (lo, hi) words:
// hi (stride, version, swizzle) is constant, lo = start address >> 4 advances by the tap window
// offset / 16-row block / 32-byte K step.
template <int NT_, int MT_, int BN_>
__device__ __forceinline__ void halo_mma_issuer(const ConvHaloParams& p, const HaloIssueCtx& c) {
  const int NT = NT_ ? NT_ : p.n_taps;
  const int MT = MT_ ? MT_ : p.MT;
  const uint32_t bn = BN_ ? static_cast<uint32_t>(BN_) : static_cast<uint32_t>(p.BN);
  const int SA = p.stages_a, SB = p.stages_b;
  const uint32_t bar_base = c.bar_base;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (SA + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * SA + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * SA + SB + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * SA + 2 * SB + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * SA + 2 * SB + 2 + a); };

  const uint32_t idesc = make_idesc_bf16(kBlockM, static_cast<int>(bn), false, false);
  const uint32_t a_hi = desc_hi_sw128(static_cast<uint32_t>(p.Wh) * 128u);
  const uint32_t b_hi = desc_hi_sw128(1024u);
  uint32_t tap_off[kMaxTaps];
#pragma unroll
  for (int t = 0; t < kMaxTaps; ++t) tap_off[t] = t < NT ? static_cast<uint32_t>(p.tap_row[t]) * 8u : 0u;
  const uint32_t mt_off = static_cast<uint32_t>(kHaloTH * p.Wh) * 8u;
  const uint32_t b_stage_bytes = bn * 128u;
  const uint32_t b_step = b_stage_bytes >> 4;
  const uint32_t acc_set_cols = static_cast<uint32_t>(MT) * bn;
  const uint32_t a_stage16 = p.a_stage_bytes >> 4;
  const uint32_t a_lo_base = desc_lo_sw128(c.smem_base, 16);
  const uint32_t b_lo_base = desc_lo_sw128(c.b_base, 16);
  const bool resident = p.resident != 0;
  uint32_t sa = 0, sb = 0, pa = 0, pb = 0;
  uint32_t a_lo = a_lo_base, b_lo_s = b_lo_base;  // descriptors of stage sa / sb
  uint32_t acc_it = 0;
  bool first = true;
  for (int u = c.u_begin; u < c.u_end; u += c.u_step, ++acc_it) {
    const uint32_t acc = acc_it & 1u;
    const uint32_t acc_phase = (acc_it >> 1) & 1u;
    mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
    tc_fence_after();
    const uint32_t d_tmem = c.tmem_base + acc * acc_set_cols;
    uint32_t b_lo_c = b_lo_base;  // resident weights: descriptor of chunk cg, tap 0
    for (int cg = 0; cg < c.chunks_total; ++cg) {
      mbar_wait(a_full(sa), pa);
      if (resident) {
        // weights of this chunk live in stages [cg * NT, (cg + 1) * NT)
        if (first) {
          for (int t = 0; t < NT; ++t) mbar_wait(b_full(cg * NT + t), 0);
        }
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int t = 0; t < kMaxTaps; ++t) {
            if (t < NT) {
#pragma unroll
              for (int mt = 0; mt < 2; ++mt) {
                if (mt < MT) {
#pragma unroll
                  for (int j = 0; j < kBlockK / 16; ++j)
                    umma_bf16_lohi(d_tmem + mt * bn, a_lo + tap_off[t] + mt * mt_off + 2u * j, a_hi,
                                   b_lo_c + t * b_step + 2u * j, b_hi, idesc,
                                   (t | j) != 0 ? 1u : (cg != 0 ? 1u : 0u));
                }
              }
            }
          }
          umma_commit(a_empty(sa));
        }
        __syncwarp();
        b_lo_c += static_cast<uint32_t>(NT) * b_step;
      } else {
#pragma unroll
        for (int t = 0; t < kMaxTaps; ++t) {
          if (t < NT) {
            mbar_wait(b_full(sb), pb);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int mt = 0; mt < 2; ++mt) {
                if (mt < MT) {
#pragma unroll
                  for (int j = 0; j < kBlockK / 16; ++j)
                    umma_bf16_lohi(d_tmem + mt * bn, a_lo + tap_off[t] + mt * mt_off + 2u * j, a_hi,
                                   b_lo_s + 2u * j, b_hi, idesc, (t | j) != 0 ? 1u : (cg != 0 ? 1u : 0u));
                }
              }
              umma_commit(b_empty(sb));
              if (t == NT - 1) umma_commit(a_empty(sa));
            }
            __syncwarp();
            b_lo_s += b_step;
            if (++sb == static_cast<uint32_t>(SB)) {
              sb = 0;
              pb ^= 1u;
              b_lo_s = b_lo_base;
            }
          }
        }
      }
      a_lo += a_stage16;
      if (++sa == static_cast<uint32_t>(SA)) {
        sa = 0;
        pa ^= 1u;
        a_lo = a_lo_base;
      }
    }
    if (elect_one()) umma_commit(tfull_bar(acc));
    __syncwarp();
    first = false;
  }
}

__global__ void __launch_bounds__(kHaloThreads, 1)
    conv_halo_kernel(const __grid_constant__ ConvHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t b_stage_bytes = static_cast<uint32_t>(p.BN) * 128u;
  const uint32_t stg_bytes_k = (p.tma_epilogue ? (p.mask != nullptr ? 4u : 2u) * 16384u : 0u) +
                               (p.pool ? 2u * 4096u : 0u);
  const uint32_t b_base = smem_base + p.stages_a * p.a_stage_bytes + stg_bytes_k;
  const uint32_t bar_base = b_base + p.stages_b * b_stage_bytes;
  // barriers: a_full[SA] a_empty[SA] b_full[SB] b_empty[SB] tfull[2] tempty[2]
  const int SA = p.stages_a, SB = p.stages_b;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (SA + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * SA + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * SA + SB + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * SA + 2 * SB + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * SA + 2 * SB + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * SA + 2 * SB + 4);
  const uint32_t bias_base = tmem_slot + 16u;  // float [2][128]
  // epilogue staging: out[2] (16 KiB each) then mask[2]; placed before the barriers, 1024-aligned
  const uint32_t stg_base = b_base - stg_bytes_k;
  const uint32_t pool_stg = b_base - 2u * 4096u;  // (only meaningful when p.pool)
  const uint32_t mfull_base = bias_base + 2u * 128u * 4u;  // mask_full[2] mbarriers
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  float* bias_s = reinterpret_cast<float*>(smem_gen + (bias_base - smem_base));

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.n_src; ++s) tma_prefetch_desc(&p.a_map[s]);
    tma_prefetch_desc(&p.b_map);
    if (p.tma_epilogue) {
      tma_prefetch_desc(&p.out_map);
      if (p.mask != nullptr) tma_prefetch_desc(&p.mask_map);
      if (p.pool) tma_prefetch_desc(&p.pool_map);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < SA; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
      mbar_init(mfull_base + 8u * a, 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int spatial_units = p.n_img * tiles_per_img;
  int chunks_total = 0;
  for (int s = 0; s < p.n_src; ++s) chunks_total += p.src_chunks[s];
  const uint32_t a_bytes = static_cast<uint32_t>(p.Wh * p.Hh) * 128u;
  const uint32_t acc_set_cols = static_cast<uint32_t>(p.MT * p.BN);

  // Unit enumeration.  Streaming weights: unit = spatial * n_tiles_n + n_tile over the whole grid.
  // Resident weights: CTA b owns N tile (b % n_tiles_n) and walks the spatial blocks
  // (b / n_tiles_n) + i * (gridDim.x / n_tiles_n).
  int u_begin, u_step, u_end, fixed_n;
  if (p.resident) {
    fixed_n = blockIdx.x % p.n_tiles_n;
    u_begin = blockIdx.x / p.n_tiles_n;
    u_step = gridDim.x / p.n_tiles_n;
    u_end = spatial_units;
  } else {
    fixed_n = -1;
    u_begin = blockIdx.x;
    u_step = gridDim.x;
    u_end = spatial_units * p.n_tiles_n;
  }
  auto decode = [&](int u, int* n_tile, int* tx, int* ty, int* img) {
    int sp = u;
    if (fixed_n >= 0) {
      *n_tile = fixed_n;
    } else {
      *n_tile = u % p.n_tiles_n;
      sp = u / p.n_tiles_n;
    }
    *tx = sp % p.tiles_x;
    *ty = (sp / p.tiles_x) % p.tiles_y;
    *img = sp / tiles_per_img;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // (whole warp converged, one elected lane issues; see conv_gemm.cu)
    uint32_t sa = 0, sb = 0, pa = 0, pb = 0;
    bool first = true;
    for (int u = u_begin; u < u_end; u += u_step) {
      int n_tile, tx, ty, img;
      decode(u, &n_tile, &tx, &ty, &img);
      const int x0 = tx * kHaloTW, y0 = ty * kHaloTH * p.MT, n0 = n_tile * p.BN;
      int cg = 0;
      for (int s = 0; s < p.n_src; ++s) {
        const int cx = x0 + p.src_off_x[s], cy = y0 + p.src_off_y[s];
        for (int c = 0; c < p.src_chunks[s]; ++c, ++cg) {
          mbar_wait(a_empty(sa), pa ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(a_full(sa), a_bytes);
            tma_load_4d(smem_base + sa * p.a_stage_bytes, &p.a_map[s], a_full(sa), c * kBlockK, cx,
                        cy, img);
          }
          __syncwarp();
          if (++sa == static_cast<uint32_t>(SA)) {
            sa = 0;
            pa ^= 1u;
          }
          if (!p.resident || first) {
            for (int t = 0; t < p.n_taps; ++t) {
              mbar_wait(b_empty(sb), pb ^ 1u);
              if (elect_one()) {
                mbar_expect_tx(b_full(sb), b_stage_bytes);
                tma_load_2d(b_base + sb * b_stage_bytes, &p.b_map, b_full(sb),
                            (t * chunks_total + cg) * kBlockK, n0);
              }
              __syncwarp();
              if (++sb == static_cast<uint32_t>(SB)) {
                sb = 0;
                pb ^= 1u;
              }
            }
          }
        }
      }
      first = false;
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    // (specialised on taps / accumulators / N tile: the issue loop must cost well under one MMA
    // time -- 32..64 cycles -- per instruction or the tensor pipe starves; see halo_mma_issuer)
    const HaloIssueCtx c{smem_base, b_base, bar_base, tmem_base, u_begin, u_step, u_end, chunks_total};
    if (p.n_taps == 9 && p.MT == 2 && p.BN == 64) halo_mma_issuer<9, 2, 64>(p, c);
    else if (p.n_taps == 9 && p.MT == 2 && p.BN == 128) halo_mma_issuer<9, 2, 128>(p, c);
    else if (p.n_taps == 9 && p.MT == 1 && p.BN == 64) halo_mma_issuer<9, 1, 64>(p, c);
    else if (p.n_taps == 9 && p.MT == 1 && p.BN == 128) halo_mma_issuer<9, 1, 128>(p, c);
    else halo_mma_issuer<0, 0, 0>(p, c);
  } else if (warp >= 4 && p.tma_epilogue) {
    // ------------------------------------------------------------ epilogue (TMA stores)
    // A "block" is one 64-channel column block of one 16-row accumulator: 128 rows x 128 B in
    // shared memory, row m = pixel (m / 8, m % 8), 16-byte chunk c stored at c ^ (m & 7).
    const int wq = warp & 3;
    const int m = wq * 32 + lane;
    const int et = threadIdx.x - 128;
    const bool leader = et == 0;
    const bool has_mask = p.mask != nullptr;
    const int cbs = p.BN / 64;               // column blocks per accumulator
    const int bpu = p.MT * cbs;              // blocks per unit
    const uint32_t out_stg = stg_base;
    const uint32_t mask_stg = stg_base + 2u * 16384u;
    const uint32_t row_off = static_cast<uint32_t>(m) * 128u;
    const uint32_t swz = static_cast<uint32_t>(m & 7);
    // block q (counted over this CTA's whole life) -> tensor coordinates of its 8 x 16 pixel box
    auto block_coords = [&](long long q, int* c0, int* bx, int* by, int* bimg) -> bool {
      const long long ord = q / bpu;
      const int rem = static_cast<int>(q % bpu);
      const long long u = u_begin + ord * u_step;
      if (u >= u_end) return false;
      int n_tile, tx, ty, img;
      decode(static_cast<int>(u), &n_tile, &tx, &ty, &img);
      const int mt = rem / cbs, cb = rem % cbs;
      *c0 = n_tile * p.BN + cb * 64;
      *bx = tx * kHaloTW;
      *by = (ty * p.MT + mt) * kHaloTH;
      *bimg = img;
      return true;
    };
    // (with a partial mask every block still fetches a mask tile -- blocks outside the masked
    // channel range fetch channel 0 and ignore it -- so that the two-deep tile pipeline and its
    // barrier phases do not depend on the block)
    auto mask_chan = [&](int c0) -> int {
      if (p.mask_nc == 0) return c0;
      const int c = c0 - p.mask_c0;
      return (c >= 0 && c < p.mask_nc) ? c : 0;
    };
    auto issue_mask = [&](long long q) {
      int c0, bx, by, bimg;
      if (block_coords(q, &c0, &bx, &by, &bimg)) {
        const uint32_t bar = mfull_base + 8u * static_cast<uint32_t>(q & 1);
        mbar_expect_tx(bar, 16384u);
        tma_load_4d(mask_stg + static_cast<uint32_t>(q & 1) * 16384u, &p.mask_map, bar, mask_chan(c0), bx,
                    by, bimg);
      }
    };
    if (has_mask && leader) {
      issue_mask(0);
      issue_mask(1);
    }
    long long q = 0;
    uint32_t acc_it = 0;
    for (int u = u_begin; u < u_end; u += u_step, ++acc_it) {
      const uint32_t acc = acc_it & 1u;
      const uint32_t acc_phase = (acc_it >> 1) & 1u;
      int n_tile, tx, ty, img;
      decode(u, &n_tile, &tx, &ty, &img);
      const int n0 = n_tile * p.BN;
      float* bias_t = bias_s + acc * 128;
      if (p.bias != nullptr) {
        for (int j = et; j < p.BN; j += 128) bias_t[j] = __ldg(p.bias + n0 + j);
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      // (bias_t is published by the first named barrier of the first block below)
      for (int mt = 0; mt < p.MT; ++mt) {
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) +
                               acc * acc_set_cols + mt * p.BN;
        for (int cb = 0; cb < cbs; ++cb, ++q) {
          const uint32_t buf = static_cast<uint32_t>(q & 1);
          uint32_t r0[32], r1[32];
          tmem_ld32(t_row + cb * 64, r0);
          tmem_ld32(t_row + cb * 64 + 32, r1);
          if (has_mask) mbar_wait(mfull_base + 8u * buf, static_cast<uint32_t>((q >> 1) & 1));
          tmem_ld_wait();
          // barrier A: the leader has seen the store that last used out_stg[buf] finish reading,
          // and (first block of a unit) every warp's bias_t writes are visible
          named_bar_sync(1, 128);
          uint4 packed[8];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(h == 0 ? r0[j] : r1[j]);
            if (p.bias != nullptr) {
              add_bias32(v, bias_t + cb * 64 + h * 32);
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (has_mask && (p.mask_nc == 0 || (n0 + cb * 64 >= p.mask_c0 &&
                                                n0 + cb * 64 < p.mask_c0 + p.mask_nc))) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const uint4 mv = ld_shared_v4(mask_stg + buf * 16384u + row_off +
                                              ((static_cast<uint32_t>(h * 4 + c) ^ swz) << 4));
                const uint32_t w[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if (!(bf16_lo(w[e]) > 0.f)) v[c * 8 + 2 * e] = 0.f;
                  if (!(bf16_hi(w[e]) > 0.f)) v[c * 8 + 2 * e + 1] = 0.f;
                }
              }
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
            st_shared_v4(out_stg + buf * 16384u + row_off + ((static_cast<uint32_t>(c) ^ swz) << 4), packed[c]);
          if (p.pool) {
            // 2x2 max pool of the block: the window of pixel (ly, lx) = lanes {m, m^1, m^8} of this
            // warp (8-pixel-wide tile rows); ReLU outputs are >= 0, so the bf16 maximum is the
            // unsigned maximum of the 16-bit patterns.  Lanes with even (ly, lx) own a pooled pixel.
            uint4 pooled[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              uint32_t w[4] = {packed[c].x, packed[c].y, packed[c].z, packed[c].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                w[e] = __vmaxu2(w[e], __shfl_xor_sync(0xffffffffu, w[e], 1));
                w[e] = __vmaxu2(w[e], __shfl_xor_sync(0xffffffffu, w[e], 8));
              }
              pooled[c] = make_uint4(w[0], w[1], w[2], w[3]);
            }
            if ((m & 1) == 0 && (m & 8) == 0) {
              const uint32_t pr = static_cast<uint32_t>(((m >> 4) << 2) | ((m & 7) >> 1));  // (ly/2)*4 + lx/2
              const uint32_t prow = pool_stg + buf * 4096u + pr * 128u;
#pragma unroll
              for (int c = 0; c < 8; ++c)
                st_shared_v4(prow + ((static_cast<uint32_t>(c) ^ (pr & 7u)) << 4), pooled[c]);
            }
          }
          fence_proxy_async();
          named_bar_sync(2, 128);  // barrier B: block complete in smem, mask block fully consumed
          if (leader) {
            tma_store_4d(&p.out_map, out_stg + buf * 16384u, n0 + cb * 64, tx * kHaloTW,
                         (ty * p.MT + mt) * kHaloTH, img);
            if (p.pool)
              tma_store_4d(&p.pool_map, pool_stg + buf * 4096u, n0 + cb * 64, tx * (kHaloTW / 2),
                           (ty * p.MT + mt) * (kHaloTH / 2), img);
            tma_store_commit();
            if (has_mask) issue_mask(q + 2);
            tma_store_wait_read<1>();  // the other staging buffer is free again
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
    if (leader) tma_store_wait_all<0>();
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (direct stores)
    const int wq = warp & 3;
    const int m = wq * 32 + lane;
    const int ly = m >> 3, lx = m & 7;
    const int et = threadIdx.x - 128;
    uint32_t acc_it = 0;
    for (int u = u_begin; u < u_end; u += u_step, ++acc_it) {
      const uint32_t acc = acc_it & 1u;
      const uint32_t acc_phase = (acc_it >> 1) & 1u;
      int n_tile, tx, ty, img;
      decode(u, &n_tile, &tx, &ty, &img);
      const int n0 = n_tile * p.BN;
      const int x = tx * kHaloTW + lx;

      float* bias_t = bias_s + acc * 128;
      if (p.bias != nullptr) {
        for (int j = et; j < p.BN; j += 128) bias_t[j] = __ldg(p.bias + n0 + j);
      }
      // ReLU-gradient mask bits of this unit, fetched while the MMAs are still running
      uint32_t mbits[2][4];
      const bool tile_masked =
          p.mask != nullptr && (p.mask_nc == 0 || (n0 >= p.mask_c0 && n0 < p.mask_c0 + p.mask_nc));
      if (tile_masked) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int y = (ty * p.MT + mt) * kHaloTH + ly;
          const bool valid = mt < p.MT && y < p.H_out && x < p.W_out;
          load_mask_bits4(p.mask + img * p.mask_sn + y * p.mask_sy + x * p.mask_sx + (n0 - p.mask_c0),
                          p.BN / 32, valid, mbits[mt]);
        }
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      named_bar_sync(1, 128);  // bias_t visible to the 4 epilogue warps

      for (int mt = 0; mt < p.MT; ++mt) {
        const int y = (ty * p.MT + mt) * kHaloTH + ly;
        const bool valid = y < p.H_out && x < p.W_out;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) +
                               acc * acc_set_cols + mt * p.BN;
        for (int ch = 0; ch < p.BN / 32; ++ch) {
          uint32_t r[32];
          tmem_ld32(t_row + ch * 32, r);
          tmem_ld_wait();
          if (valid) {
            const int n = n0 + ch * 32;
            const long long off = img * p.out_sn + y * p.out_sy + x * p.out_sx + n;
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
              const uint32_t mb = mt == 0 ? (ch == 0 ? mbits[0][0] : ch == 1 ? mbits[0][1] : ch == 2 ? mbits[0][2] : mbits[0][3])
                                          : (ch == 0 ? mbits[1][0] : ch == 1 ? mbits[1][1] : ch == 2 ? mbits[1][2] : mbits[1][3]);
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
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------ host side
// Returns RSU_OK and launches, or a negative value (-1) when the shape is not eligible for the
// halo path (the caller then uses the per-tap kernel), or a positive RSU_E* code on error.
int launch_conv_halo(const rsu_conv_gemm_desc* d, cudaStream_t stream, bool forced) {
  if (d->n_taps < 2 || d->shuffle_cout > 0) return -1;
  if (d->Ntot % 64 != 0) return -1;
  int min_dy = d->tap_dy[0], max_dy = d->tap_dy[0], min_dx = d->tap_dx[0], max_dx = d->tap_dx[0];
  for (int t = 1; t < d->n_taps; ++t) {
    min_dy = d->tap_dy[t] < min_dy ? d->tap_dy[t] : min_dy;
    max_dy = d->tap_dy[t] > max_dy ? d->tap_dy[t] : max_dy;
    min_dx = d->tap_dx[t] < min_dx ? d->tap_dx[t] : min_dx;
    max_dx = d->tap_dx[t] > max_dx ? d->tap_dx[t] : max_dx;
  }
  const int span_y = max_dy - min_dy, span_x = max_dx - min_dx;
  if (span_x > 8 || span_y > 8) return -1;
  int chunks_total = 0;
  for (int s = 0; s < d->n_src; ++s) chunks_total += d->src[s].C / 64;
  const int num_k = d->n_taps * chunks_total;
  const int smem_max = 227 * 1024 - 1024 /*align*/ - 2048 /*barriers, bias*/;
  // fused max pool: ReLU epilogue without mask / accumulation, even output extents
  const bool want_pool = d->pool_out != nullptr && d->relu && !d->mask && !d->accumulate &&
                         d->H_out % 2 == 0 && d->W_out % 2 == 0;
  const int stg_full = (d->mask ? 4 : 2) * 16384 + (want_pool ? 2 * 4096 : 0);  // staging of the TMA epilogue

  ConvHaloParams p;
  memset(&p, 0, sizeof(p));
  p.Wh = kHaloTW + span_x;
  // N tiling: 128 when it divides, else 64 (two accumulator sets of MT*BN columns <= 512)
  p.BN = d->Ntot % 128 == 0 ? 128 : 64;
  p.n_tiles_n = d->Ntot / p.BN;
  auto a_stage = [&](int mt) {
    return static_cast<uint32_t>(((p.Wh * (kHaloTH * mt + span_y) * 128) + 1023) & ~1023);
  };
  // Candidate configurations in order of preference: resident weights (TMA epilogue, MT = 2 / 1;
  // then direct-store epilogue, which needs no staging buffers, MT = 2 / 1), else streaming
  // weights with MT = 2 and the TMA epilogue.
  const int b_stage = p.BN * 128;
  const int sms = num_sms();
  bool chosen = false;
  bool tma_epi = !d->accumulate;
  int stg_bytes = 0, smem_budget = smem_max;
  if (p.n_tiles_n <= sms && num_k <= kMaxBStages) {
    for (int epi = d->accumulate ? 0 : 1; epi >= 0 && !chosen; --epi) {
      const int budget = smem_max - (epi ? stg_full : 0);
      for (int mt = 2; mt >= 1 && !chosen; --mt) {
        if (mt == 2 && d->H_out <= kHaloTH) continue;  // a single block row suffices
        const int need = num_k * b_stage + 2 * static_cast<int>(a_stage(mt));
        if (need <= budget) {
          p.resident = 1;
          p.MT = mt;
          p.stages_b = num_k;
          p.stages_a = (budget - num_k * b_stage) / static_cast<int>(a_stage(mt));
          if (p.stages_a > 4) p.stages_a = 4;
          tma_epi = epi != 0;
          stg_bytes = epi ? stg_full : 0;
          smem_budget = budget;
          chosen = true;
        }
      }
    }
  }
  if (!chosen) {
    stg_bytes = tma_epi ? stg_full : 0;
    smem_budget = smem_max - stg_bytes;
    p.resident = 0;
    p.MT = (d->H_out > kHaloTH) ? 2 : 1;
    p.stages_a = 2;
    p.stages_b = (smem_budget - 2 * static_cast<int>(a_stage(p.MT))) / b_stage;
    if (p.stages_b > kMaxBStages) p.stages_b = kMaxBStages;
    if (p.stages_b < 3) return -1;
  }
  p.a_stage_bytes = a_stage(p.MT);
  p.Hh = kHaloTH * p.MT + span_y;
  if (p.Wh > 256 || p.Hh > 256) return -1;
  (void)forced;

  p.tiles_x = (d->W_out + kHaloTW - 1) / kHaloTW;
  p.tiles_y = (d->H_out + kHaloTH * p.MT - 1) / (kHaloTH * p.MT);
  p.n_img = d->N_img;
  p.H_out = d->H_out;
  p.W_out = d->W_out;
  p.n_src = d->n_src;
  for (int s = 0; s < d->n_src; ++s) {
    const rsu_view& v = d->src[s];
    if (v.N != d->N_img) return set_error(RSU_EINVAL, "source %d batch %d != %d", s, v.N, d->N_img);
    if (v.W < p.Wh || v.H < p.Hh) return -1;  // TMA box larger than the tensor
    int rc = encode_act_map(&p.a_map[s], v, p.Wh, p.Hh);
    if (rc) return rc;
    p.src_chunks[s] = v.C / 64;
    p.src_off_y[s] = v.off_y + min_dy;
    p.src_off_x[s] = v.off_x + min_dx;
  }
  p.n_taps = d->n_taps;
  for (int t = 0; t < d->n_taps; ++t)
    p.tap_row[t] = (d->tap_dy[t] - min_dy) * p.Wh + (d->tap_dx[t] - min_dx);
  {
    int rc = encode_weight_map(&p.b_map, d->weights, num_k * 64, d->Ntot, p.BN);
    if (rc) return rc;
  }
  p.out = static_cast<__nv_bfloat16*>(d->out);
  p.out_sn = d->out_sn;
  p.out_sy = d->out_sy;
  p.out_sx = d->out_sx;
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
  p.tma_epilogue = tma_epi ? 1 : 0;
  p.pool = (want_pool && tma_epi) ? 1 : 0;
  if (p.pool) {
    rsu_view pv;
    pv.ptr = d->pool_out;
    pv.C = d->Ntot;
    pv.H = d->H_out / 2;
    pv.W = d->W_out / 2;
    pv.N = d->N_img;
    pv.sn = d->pool_sn;
    pv.sy = d->pool_sy;
    pv.sx = d->pool_sx;
    pv.off_y = pv.off_x = 0;
    int rc = encode_act_map(&p.pool_map, pv, kHaloTW / 2, kHaloTH / 2);
    if (rc) return rc;
  }
  if (tma_epi) {
    rsu_view ov;
    ov.ptr = d->out;
    ov.C = d->Ntot;
    ov.H = d->H_out;
    ov.W = d->W_out;
    ov.N = d->N_img;
    ov.sn = d->out_sn;
    ov.sy = d->out_sy;
    ov.sx = d->out_sx;
    ov.off_y = ov.off_x = 0;
    int rc = encode_act_map(&p.out_map, ov, kHaloTW, kHaloTH);
    if (rc) return rc;
    if (d->mask) {
      ov.ptr = d->mask;
      if (d->mask_nc > 0) ov.C = d->mask_nc;
      ov.sn = d->mask_sn;
      ov.sy = d->mask_sy;
      ov.sx = d->mask_sx;
      rc = encode_act_map(&p.mask_map, ov, kHaloTW, kHaloTH);
      if (rc) return rc;
    }
  }

  const int smem = 1024 + p.stages_a * static_cast<int>(p.a_stage_bytes) + stg_bytes +
                   p.stages_b * b_stage + 8 * (2 * p.stages_a + 2 * p.stages_b + 4) + 16 +
                   2 * 128 * 4 + 16 /*mask_full*/;
  static bool attr_set = false;
  if (!attr_set) {
    RSU_CHECK_CUDA(cudaFuncSetAttribute(conv_halo_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const long long spatial = 1LL * p.n_img * p.tiles_x * p.tiles_y;
  int grid;
  if (p.resident) {
    long long per_n = sms / p.n_tiles_n;
    if (per_n > spatial) per_n = spatial;
    if (per_n < 1) per_n = 1;
    grid = static_cast<int>(per_n) * p.n_tiles_n;
  } else {
    const long long total = spatial * p.n_tiles_n;
    grid = total < sms ? static_cast<int>(total) : sms;
  }
  conv_halo_kernel<<<grid, kHaloThreads, smem, stream>>>(p);
  if (p.pool && d->pool_done_host) *d->pool_done_host = 1;
  return check_launch("conv_halo_kernel");
}

}  // namespace rsu
