// CTA-pair variant of the halo-tile convolution of conv_halo.cu (sm_100a).
//
// The halo kernels of the layers with few output channels (N = 64 / 128) are bound by the
// shared-memory pipe: an M = 128, N = 64 tcgen05.mma reads (128 + 64) x 32 B of operands in 32
// tensor cycles against 128 B/cycle, and the TMA writes and epilogue staging go through the same
// pipe.  Here a cluster of two CTAs runs ONE tcgen05.mma.cta_group::2
// instruction stream with M = 2 x 128:
//   * each CTA loads its own halo tile (its own 8 x 16*MT block of output pixels) and keeps HALF
//     of the resident weights (BN/2 rows per (chunk, tap)): (128 + 32) x 32 B per MMA and SM at
//     N = 64, and half of the weight footprint, which leaves room for a deeper halo-tile ring;
//   * all TMA loads of the pair complete on the LEADER's barriers; the leader's MMA thread issues
//     for both, its tcgen05.commit multicasts "stage free" / "accumulator full" to both CTAs;
//   * each CTA drains its own TMEM lanes with the TMA-store epilogue of conv_halo.cu (bias / ReLU /
//     ReLU-gradient mask / fused 2x2 max pool); the peer arrives remotely (relaxed) on the
//     leader's "accumulator empty" barrier.
// Restrictions (otherwise the caller falls back to conv_halo_kernel): resident weights, the
// TMA-store epilogue (no accumulation), 9 taps.
//
// Reference ops replaced: as conv_halo.cu (src/unet.py:34-45, 52, 88-91).
#include "gemm_params.h"
#include "host_common.h"
#include "ptx.cuh"

namespace rsu {

constexpr int kH2Threads = 256;
constexpr int kH2TW = 8;   // block width in pixels = rows of one swizzle group
constexpr int kH2TH = 16;  // image rows per 128-row accumulator
constexpr int kH2MaxB = 28;
constexpr uint32_t kH2PeerMask = 0xFEFFFFFFu;  // shared::cluster address of the even CTA of a pair

struct ConvHalo2Params {
  CUtensorMap a_map[kMaxSrc];  // 4-D (C, W, H, N) bf16, SWIZZLE_128B, box {64, Wh, Hh, 1}
  CUtensorMap b_map;           // 2-D (Ktot, Ntot) bf16, SWIZZLE_128B, box {64, BN / 2}
  CUtensorMap out_map;         // 4-D (C, W, H, N) bf16, box {64, 8, 16, 1}
  CUtensorMap mask_map;        // same geometry over the mask tensor
  CUtensorMap pool_map;        // 4-D (C, W/2, H/2, N) bf16, box {64, 4, 8, 1}
  int n_src;
  int src_chunks[kMaxSrc];
  int src_off_y[kMaxSrc];
  int src_off_x[kMaxSrc];
  int tap_row[kMaxTaps];  // first smem row of the tap's window inside the halo tile (9 taps)
  int Wh, Hh;
  int MT;  // accumulators (16-row blocks) per unit: 1 or 2
  int tiles_x, tiles_y, n_img;
  int n_tiles_n, BN;
  int stages_a, stages_b;  // stages_b = 9 * chunks (resident)
  uint32_t a_stage_bytes;
  const float* bias;
  int relu;
  int has_mask;
  int mask_c0, mask_nc;
  int pool;
};

namespace h2 {
__device__ __forceinline__ uint32_t ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\t"
               "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void umma2_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                           uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void commit_pair(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64"
      " [%0], %1;" ::"r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1,
                                             int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void arrive_leader_relaxed(uint32_t bar) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(bar));
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
}  // namespace h2

// The MMA-issuing thread of the pair (leader CTA), specialised like halo_mma_issuer of conv_halo.cu.
template <int MT_, int BN_>
__device__ __forceinline__ void halo2_mma_issuer(const ConvHalo2Params& p, uint32_t smem_base, uint32_t b_base,
                                                 uint32_t bar_base, uint32_t tmem_base, int n_items,
                                                 int chunks_total) {
  constexpr int NT = 9;
  const int MT = MT_ ? MT_ : p.MT;
  const uint32_t bn = BN_ ? static_cast<uint32_t>(BN_) : static_cast<uint32_t>(p.BN);
  const int SA = p.stages_a, SB = p.stages_b;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (SA + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * SA + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * SA + SB + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * SA + SB + 2 + a); };

  const uint32_t idesc = make_idesc_bf16(2 * kBlockM, static_cast<int>(bn), false, false);
  const uint32_t a_hi = desc_hi_sw128(static_cast<uint32_t>(p.Wh) * 128u);
  const uint32_t b_hi = desc_hi_sw128(1024u);
  uint32_t tap_off[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) tap_off[t] = static_cast<uint32_t>(p.tap_row[t]) * 8u;
  const uint32_t mt_off = static_cast<uint32_t>(kH2TH * p.Wh) * 8u;
  const uint32_t b_step = ((bn / 2u) * 128u) >> 4;  // half of the weight rows per CTA
  const uint32_t acc_set_cols = static_cast<uint32_t>(MT) * bn;
  const uint32_t a_stage16 = p.a_stage_bytes >> 4;
  const uint32_t a_lo_base = desc_lo_sw128(smem_base, 16);
  const uint32_t b_lo_base = desc_lo_sw128(b_base, 16);
  uint32_t sa = 0, pa = 0;
  uint32_t a_lo = a_lo_base;
  bool first = true;
  for (int it = 0; it < n_items; ++it) {
    const uint32_t acc = static_cast<uint32_t>(it) & 1u;
    const uint32_t acc_phase = (static_cast<uint32_t>(it) >> 1) & 1u;
    mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + acc * acc_set_cols;
    uint32_t b_lo_c = b_lo_base;
    for (int cg = 0; cg < chunks_total; ++cg) {
      mbar_wait(a_full(sa), pa);
      if (first) {
        for (int t = 0; t < NT; ++t) mbar_wait(b_full(cg * NT + t), 0);
      }
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            if (mt < MT) {
#pragma unroll
              for (int j = 0; j < kBlockK / 16; ++j)
                h2::umma2_lohi(d_tmem + mt * bn, a_lo + tap_off[t] + mt * mt_off + 2u * j, a_hi,
                               b_lo_c + t * b_step + 2u * j, b_hi, idesc,
                               (t | j) != 0 ? 1u : (cg != 0 ? 1u : 0u));
            }
          }
        }
        h2::commit_pair(a_empty(sa));
      }
      __syncwarp();
      b_lo_c += static_cast<uint32_t>(NT) * b_step;
      a_lo += a_stage16;
      if (++sa == static_cast<uint32_t>(SA)) {
        sa = 0;
        pa ^= 1u;
        a_lo = a_lo_base;
      }
    }
    if (elect_one()) h2::commit_pair(tfull_bar(acc));
    __syncwarp();
    first = false;
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kH2Threads, 1)
    conv_halo2_kernel(const __grid_constant__ ConvHalo2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = h2::ctarank();
  const bool leader = rank == 0;

  const uint32_t b_stage_bytes = static_cast<uint32_t>(p.BN / 2) * 128u;
  const uint32_t stg_bytes_k = (p.has_mask ? 4u : 2u) * 16384u + (p.pool ? 2u * 4096u : 0u);
  const uint32_t b_base = smem_base + p.stages_a * p.a_stage_bytes + stg_bytes_k;
  const uint32_t bar_base = b_base + p.stages_b * b_stage_bytes;
  // barriers: a_full[SA] a_empty[SA] b_full[SB] tfull[2] tempty[2]
  const int SA = p.stages_a, SB = p.stages_b;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (SA + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * SA + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * SA + SB + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * SA + SB + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * SA + SB + 4);
  const uint32_t bias_base = (tmem_slot + 16u + 15u) & ~15u;  // float [2][128], read as float4
  const uint32_t stg_base = b_base - stg_bytes_k;
  const uint32_t pool_stg = b_base - 2u * 4096u;
  const uint32_t mfull_base = bias_base + 2u * 128u * 4u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  float* bias_s = reinterpret_cast<float*>(smem_gen + (bias_base - smem_base));

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.n_src; ++s) tma_prefetch_desc(&p.a_map[s]);
    tma_prefetch_desc(&p.b_map);
    tma_prefetch_desc(&p.out_map);
    if (p.has_mask) tma_prefetch_desc(&p.mask_map);
    if (p.pool) tma_prefetch_desc(&p.pool_map);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < SA; ++s) {
      mbar_init(a_full(s), 1);   // leader: its producer's arrive + both CTAs' bytes
      mbar_init(a_empty(s), 1);  // one multicast commit per use
    }
    for (int s = 0; s < SB; ++s) mbar_init(b_full(s), 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8);  // leader: 4 epilogue warps of each CTA
      mbar_init(mfull_base + 8u * a, 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) h2::tmem_alloc2(tmem_slot, 512);
  tc_fence_before();
  h2::cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int spatial_units = p.n_img * tiles_per_img;
  int chunks_total = 0;
  for (int s = 0; s < p.n_src; ++s) chunks_total += p.src_chunks[s];
  const uint32_t a_bytes = static_cast<uint32_t>(p.Wh * p.Hh) * 128u;
  const uint32_t acc_set_cols = static_cast<uint32_t>(p.MT * p.BN);

  // The pair (blockIdx.x / 2) owns N tile pair_id % n_tiles_n and walks the spatial blocks two at
  // a time: item i -> blocks 2 k and 2 k + 1 with k = pair_id / n_tiles_n + i * pairs_per_n; this
  // CTA takes block 2 k + rank.  The odd CTA of a trailing half pair recomputes the last block
  // without storing it.
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int fixed_n = pair_id % p.n_tiles_n;
  const int k_begin = pair_id / p.n_tiles_n, k_step = n_pairs / p.n_tiles_n;
  const int k_end = (spatial_units + 1) >> 1;
  const int n_items = k_begin < k_end ? (k_end - 1 - k_begin) / k_step + 1 : 0;
  auto decode = [&](int item, int* tx, int* ty, int* img) -> bool {
    const int raw = 2 * (k_begin + item * k_step) + static_cast<int>(rank);
    const int sp = min(raw, spatial_units - 1);
    *tx = sp % p.tiles_x;
    *ty = (sp / p.tiles_x) % p.tiles_y;
    *img = sp / tiles_per_img;
    return raw < spatial_units;
  };
  const int n0 = fixed_n * p.BN;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    uint32_t sa = 0, pa = 0;
    for (int it = 0; it < n_items; ++it) {
      int tx, ty, img;
      decode(it, &tx, &ty, &img);
      const int x0 = tx * kH2TW, y0 = ty * kH2TH * p.MT;
      int cg = 0;
      for (int s = 0; s < p.n_src; ++s) {
        const int cx = x0 + p.src_off_x[s], cy = y0 + p.src_off_y[s];
        for (int c = 0; c < p.src_chunks[s]; ++c, ++cg) {
          mbar_wait(a_empty(sa), pa ^ 1u);
          if (elect_one()) {
            if (leader) mbar_expect_tx(a_full(sa), 2u * a_bytes);
            h2::tma2_load_4d(smem_base + sa * p.a_stage_bytes, &p.a_map[s], a_full(sa) & kH2PeerMask,
                             c * kBlockK, cx, cy, img);
          }
          __syncwarp();
          if (++sa == static_cast<uint32_t>(SA)) {
            sa = 0;
            pa ^= 1u;
          }
          if (it == 0) {
            // resident weights: this CTA's half of the rows of every (chunk, tap) slice, once
            if (elect_one()) {
              for (int t = 0; t < 9; ++t) {
                const int sb = cg * 9 + t;
                if (leader) mbar_expect_tx(b_full(sb), 2u * b_stage_bytes);
                h2::tma2_load_2d(b_base + sb * b_stage_bytes, &p.b_map, b_full(sb) & kH2PeerMask,
                                 (t * chunks_total + cg) * kBlockK,
                                 n0 + static_cast<int>(rank) * (p.BN / 2));
              }
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ------------------------------------------------------------ MMA issuer (leader CTA)
    if (p.MT == 2 && p.BN == 64) halo2_mma_issuer<2, 64>(p, smem_base, b_base, bar_base, tmem_base, n_items, chunks_total);
    else if (p.MT == 2 && p.BN == 128) halo2_mma_issuer<2, 128>(p, smem_base, b_base, bar_base, tmem_base, n_items, chunks_total);
    else halo2_mma_issuer<0, 0>(p, smem_base, b_base, bar_base, tmem_base, n_items, chunks_total);
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (TMA stores, both CTAs)
    // A "block" is one 64-channel column block of one 16-row accumulator: 128 rows x 128 B in
    // shared memory, row m = pixel (m / 8, m % 8), 16-byte chunk c stored at c ^ (m & 7).
    const int wq = warp & 3;
    const int m = wq * 32 + lane;
    const int et = threadIdx.x - 128;
    const bool lead_thread = et == 0;
    const bool has_mask = p.has_mask != 0;
    const int cbs = p.BN / 64;
    const int bpu = p.MT * cbs;
    const uint32_t out_stg = stg_base;
    const uint32_t mask_stg = stg_base + 2u * 16384u;
    const uint32_t row_off = static_cast<uint32_t>(m) * 128u;
    const uint32_t swz = static_cast<uint32_t>(m & 7);
    auto block_coords = [&](long long q, int* c0, int* bx, int* by, int* bimg) -> bool {
      const long long ord = q / bpu;
      const int rem = static_cast<int>(q % bpu);
      if (ord >= n_items) return false;
      int tx, ty, img;
      decode(static_cast<int>(ord), &tx, &ty, &img);
      const int mt = rem / cbs, cb = rem % cbs;
      *c0 = n0 + cb * 64;
      *bx = tx * kH2TW;
      *by = (ty * p.MT + mt) * kH2TH;
      *bimg = img;
      return true;
    };
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
        tma_load_4d(mask_stg + static_cast<uint32_t>(q & 1) * 16384u, &p.mask_map, bar, mask_chan(c0), bx, by,
                    bimg);
      }
    };
    if (has_mask && lead_thread) {
      issue_mask(0);
      issue_mask(1);
    }
    long long q = 0;
    for (int it = 0; it < n_items; ++it) {
      const uint32_t acc = static_cast<uint32_t>(it) & 1u;
      const uint32_t acc_phase = (static_cast<uint32_t>(it) >> 1) & 1u;
      int tx, ty, img;
      const bool live = decode(it, &tx, &ty, &img);
      float* bias_t = bias_s + acc * 128;
      if (p.bias != nullptr) {
        for (int j = et; j < p.BN; j += 128) bias_t[j] = __ldg(p.bias + n0 + j);
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      for (int mt = 0; mt < p.MT; ++mt) {
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * acc_set_cols + mt * p.BN;
        for (int cb = 0; cb < cbs; ++cb, ++q) {
          const uint32_t buf = static_cast<uint32_t>(q & 1);
          uint32_t r0[32], r1[32];
          tmem_ld32(t_row + cb * 64, r0);
          tmem_ld32(t_row + cb * 64 + 32, r1);
          if (has_mask) mbar_wait(mfull_base + 8u * buf, static_cast<uint32_t>((q >> 1) & 1));
          tmem_ld_wait();
          named_bar_sync(1, 128);
          uint4 packed[8];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(h == 0 ? r0[j] : r1[j]);
            if (p.bias != nullptr) add_bias32(v, bias_t + cb * 64 + h * 32);
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (has_mask && (p.mask_nc == 0 || (n0 + cb * 64 >= p.mask_c0 && n0 + cb * 64 < p.mask_c0 + p.mask_nc))) {
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
              const uint32_t pr = static_cast<uint32_t>(((m >> 4) << 2) | ((m & 7) >> 1));
              const uint32_t prow = pool_stg + buf * 4096u + pr * 128u;
#pragma unroll
              for (int c = 0; c < 8; ++c)
                st_shared_v4(prow + ((static_cast<uint32_t>(c) ^ (pr & 7u)) << 4), pooled[c]);
            }
          }
          fence_proxy_async();
          named_bar_sync(2, 128);
          if (lead_thread) {
            if (live) {
              tma_store_4d(&p.out_map, out_stg + buf * 16384u, n0 + cb * 64, tx * kH2TW,
                           (ty * p.MT + mt) * kH2TH, img);
              if (p.pool)
                tma_store_4d(&p.pool_map, pool_stg + buf * 4096u, n0 + cb * 64, tx * (kH2TW / 2),
                             (ty * p.MT + mt) * (kH2TH / 2), img);
            }
            tma_store_commit();
            if (has_mask) issue_mask(q + 2);
            tma_store_wait_read<1>();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader)
          mbar_arrive(tempty_bar(acc));
        else
          h2::arrive_leader_relaxed(tempty_bar(acc));
      }
    }
    if (lead_thread) tma_store_wait_all<0>();
  }

  // neither CTA may release its shared / tensor memory while the pair's MMAs can still read it
  tc_fence_before();
  h2::cluster_sync_all();
  if (warp == 2) h2::tmem_dealloc2(tmem_base, 512);
}

// Returns RSU_OK after launching, -1 when the shape is not eligible (the caller falls back to the
// single-CTA halo kernel), or an RSU_E* code.
int launch_conv_halo2(const rsu_conv_gemm_desc* d, cudaStream_t stream) {
  if (d->n_taps != 9 || d->shuffle_cout > 0 || d->accumulate) return -1;
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
  const int num_k = 9 * chunks_total;
  if (num_k > kH2MaxB) return -1;

  ConvHalo2Params p;
  memset(&p, 0, sizeof(p));
  p.Wh = kH2TW + span_x;
  p.BN = d->Ntot % 128 == 0 ? 128 : 64;
  p.n_tiles_n = d->Ntot / p.BN;
  const int sms = num_sms();
  const int n_pairs_total = sms / 2;
  if (p.n_tiles_n > n_pairs_total) return -1;
  const bool want_pool = d->pool_out != nullptr && d->relu && !d->mask && d->H_out % 2 == 0 && d->W_out % 2 == 0;
  const int stg_bytes = (d->mask ? 4 : 2) * 16384 + (want_pool ? 2 * 4096 : 0);
  const int budget = 227 * 1024 - 1024 - 2048 - stg_bytes;
  const int b_stage = (p.BN / 2) * 128;
  auto a_stage = [&](int mt) {
    return static_cast<int>(((p.Wh * (kH2TH * mt + span_y) * 128) + 1023) & ~1023);
  };
  p.MT = 0;
  for (int mt = 2; mt >= 1; --mt) {
    if (mt == 2 && d->H_out <= kH2TH) continue;
    if (num_k * b_stage + 2 * a_stage(mt) <= budget) {
      p.MT = mt;
      break;
    }
  }
  if (p.MT == 0) return -1;
  p.stages_b = num_k;
  p.stages_a = (budget - num_k * b_stage) / a_stage(p.MT);
  if (p.stages_a > 4) p.stages_a = 4;
  p.a_stage_bytes = static_cast<uint32_t>(a_stage(p.MT));
  p.Hh = kH2TH * p.MT + span_y;
  if (p.Wh > 256 || p.Hh > 256) return -1;
  p.tiles_x = (d->W_out + kH2TW - 1) / kH2TW;
  p.tiles_y = (d->H_out + kH2TH * p.MT - 1) / (kH2TH * p.MT);
  p.n_img = d->N_img;
  p.n_src = d->n_src;
  for (int s = 0; s < d->n_src; ++s) {
    const rsu_view& v = d->src[s];
    if (v.N != d->N_img) return set_error(RSU_EINVAL, "source %d batch %d != %d", s, v.N, d->N_img);
    if (v.W < p.Wh || v.H < p.Hh) return -1;
    int rc = encode_act_map(&p.a_map[s], v, p.Wh, p.Hh);
    if (rc) return rc;
    p.src_chunks[s] = v.C / 64;
    p.src_off_y[s] = v.off_y + min_dy;
    p.src_off_x[s] = v.off_x + min_dx;
  }
  for (int t = 0; t < 9; ++t) p.tap_row[t] = (d->tap_dy[t] - min_dy) * p.Wh + (d->tap_dx[t] - min_dx);
  {
    int rc = encode_weight_map(&p.b_map, d->weights, num_k * 64, d->Ntot, p.BN / 2);
    if (rc) return rc;
  }
  p.bias = d->bias;
  p.relu = d->relu;
  p.has_mask = d->mask != nullptr ? 1 : 0;
  p.mask_c0 = d->mask_nc > 0 ? d->mask_c0 : 0;
  p.mask_nc = d->mask_nc;
  if (d->mask && d->mask_nc > 0 && (d->mask_c0 % p.BN || d->mask_nc % p.BN || d->mask_c0 < 0 ||
                                    d->mask_c0 + d->mask_nc > d->Ntot))
    return set_error(RSU_EINVAL, "mask channel range [%d, +%d) not a multiple of the N tile %d", d->mask_c0,
                     d->mask_nc, p.BN);
  p.pool = want_pool ? 1 : 0;
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
  {
    int rc = encode_act_map(&p.out_map, ov, kH2TW, kH2TH);
    if (rc) return rc;
  }
  if (d->mask) {
    rsu_view mv = ov;
    mv.ptr = d->mask;
    if (d->mask_nc > 0) mv.C = d->mask_nc;
    mv.sn = d->mask_sn;
    mv.sy = d->mask_sy;
    mv.sx = d->mask_sx;
    int rc = encode_act_map(&p.mask_map, mv, kH2TW, kH2TH);
    if (rc) return rc;
  }
  if (p.pool) {
    rsu_view pv = ov;
    pv.ptr = d->pool_out;
    pv.H = d->H_out / 2;
    pv.W = d->W_out / 2;
    pv.sn = d->pool_sn;
    pv.sy = d->pool_sy;
    pv.sx = d->pool_sx;
    int rc = encode_act_map(&p.pool_map, pv, kH2TW / 2, kH2TH / 2);
    if (rc) return rc;
  }

  const int smem = 1024 + p.stages_a * static_cast<int>(p.a_stage_bytes) + stg_bytes + p.stages_b * b_stage +
                   8 * (2 * p.stages_a + p.stages_b + 4) + 32 + 2 * 128 * 4 + 16;
  static bool attr_set = false;
  if (!attr_set) {
    RSU_CHECK_CUDA(cudaFuncSetAttribute(conv_halo2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        227 * 1024));
    attr_set = true;
  }
  const long long spatial = 1LL * p.n_img * p.tiles_x * p.tiles_y;
  long long pairs_per_n = n_pairs_total / p.n_tiles_n;
  const long long items = (spatial + 1) / 2;
  if (pairs_per_n > items) pairs_per_n = items;
  if (pairs_per_n < 1) pairs_per_n = 1;
  const int grid = static_cast<int>(2 * pairs_per_n * p.n_tiles_n);
  conv_halo2_kernel<<<grid, kH2Threads, smem, stream>>>(p);
  if (p.pool && d->pool_done_host) *d->pool_done_host = 1;
  return check_launch("conv_halo2_kernel");
}

}  // namespace rsu
