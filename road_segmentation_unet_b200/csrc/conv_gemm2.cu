// Two-CTA variant of the implicit-GEMM convolution of conv_gemm.cu (sm_100a, opt-in: algo = 3).
//
// A cluster of two CTAs (one SM pair) computes two pixel tiles of the same N tile with ONE
// tcgen05.mma.cta_group::2 instruction stream, M = 2 x 128:
//   * each CTA loads its own A tile (128 pixels x 64 channels per K step) and HALF of the weight
//     tile (BN/2 rows); the tensor cores of the pair exchange the B halves, so the weight traffic
//     from L2 and the shared-memory operand reads per SM drop by a third (96 -> 64 B/cycle at
//     BN = 256);
//   * both CTAs' TMA loads complete on the LEADER's (cluster rank 0) full barrier; the leader's
//     MMA warp issues for the pair and its tcgen05.commit multicasts the "stage free" and
//     "accumulator full" arrivals to both CTAs;
//   * each CTA drains its own 128 TMEM lanes in its own epilogue warps; the peer's epilogue
//     warps arrive remotely on the leader's "accumulator empty" barrier.
// Everything else (K-step table, tile shapes, epilogue: bias / ReLU / ReLU-grad mask /
// accumulate / transpose-conv pixel shuffle) is the single-CTA kernel's.
//
// Reference ops replaced: as conv_gemm.cu (src/unet.py:34-45, 67-91).
#include "gemm_params.h"
#include "host_common.h"
#include "ptx.cuh"

namespace rsu {

constexpr int kConv2Threads = 256;
constexpr int kA2StageBytes = kBlockM * 128;  // 16 KiB
constexpr int kTmem2Cols = 512;
constexpr int kAcc2Stride = 256;  // TMEM columns per accumulator stage
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;  // shared::cluster address of the even CTA of a pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
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
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair once all MMAs
// issued so far have completed
__device__ __forceinline__ void umma2_commit_pair(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64"
      " [%0], %1;" ::"r"(bar), "h"(mask)
      : "memory");
}
// TMA loads whose completion is signalled on a barrier of the pair's leader CTA
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const void* map, uint32_t bar, int c0,
                                             int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const void* map, uint32_t bar, int c0,
                                             int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
// arrive on the barrier at this offset in the leader CTA (cluster rank 0) from its peer
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(bar));
  // relaxed: the barrier only tells the MMA thread that this CTA's tcgen05.ld of the accumulator
  // have completed (tcgen05.wait::ld + fence::before_thread_sync above).  A release at cluster
  // scope compiles to MEMBAR.ALL.GPU, i.e. the epilogue warp would wait once per tile until all
  // of its global stores / reductions have drained.
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote)
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConv2Threads, 1)
    conv_gemm2_kernel(const __grid_constant__ ConvGemmParams p, int stages) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  const uint32_t b_half_rows = static_cast<uint32_t>(p.BN) / 2u;
  const uint32_t b_stage_bytes = b_half_rows * 128u;
  const uint32_t stage_bytes = kA2StageBytes + b_stage_bytes;
  const uint32_t bar_base = smem_base + stages * stage_bytes;
  // barrier layout: full[stages], empty[stages], tmem_full[2], tmem_empty[2]
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * stages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * stages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * stages + 4);
  const uint32_t bias_base = tmem_slot + 16u;  // float [2][256]
  const uint32_t ktab_base = bias_base + 2u * 256u * 4u;  // int4 [num_k]
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
      mbar_init(full_bar(s), 1);   // used in the leader only: its producer's arrive + 2 CTAs' bytes
      mbar_init(empty_bar(s), 1);  // one multicast commit per use
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8);  // used in the leader only: 4 epilogue warps of each CTA
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc2(tmem_slot, kTmem2Cols);
  tc_fence_before();
  cluster_sync_all();  // barriers and TMEM of both CTAs exist before anything remote touches them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int m_tiles = p.n_img * tiles_per_img;
  const int m_pairs = (m_tiles + 1) >> 1;
  const int total_items = m_pairs * p.n_tiles_n;
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  int chunks_total = 0;
  for (int s = 0; s < p.n_src; ++s) chunks_total += p.src_chunks[s];
  const int num_k = p.n_taps * chunks_total;
  const uint32_t a_bytes = static_cast<uint32_t>(p.TW * p.TH) * 128u;
  for (int k = threadIdx.x; k < num_k; k += kConv2Threads) {
    const int t = k / chunks_total;
    int cg = k % chunks_total, s = 0;
    while (s < p.n_src - 1 && cg >= p.src_chunks[s]) cg -= p.src_chunks[s++];
    ktab[k] = make_int4(s, cg * kBlockK, p.tap_dx[t] + p.src_off_x[s], p.tap_dy[t] + p.src_off_y[s]);
  }
  __syncthreads();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    uint32_t stage = 0, phase = 0;
    const uint32_t pair_tx_bytes = 2u * (a_bytes + b_stage_bytes);
    for (int item = pair_id; item < total_items; item += n_pairs) {
      const int n_tile = item % p.n_tiles_n;
      // the odd CTA of the last pair of an odd tile count recomputes the last tile (not stored)
      const int m_tile = min(2 * (item / p.n_tiles_n) + static_cast<int>(rank), m_tiles - 1);
      const int tx = m_tile % p.tiles_x;
      const int ty = (m_tile / p.tiles_x) % p.tiles_y;
      const int img = m_tile / tiles_per_img;
      const int x0 = tx * p.TW, y0 = ty * p.TH;
      const int n0 = n_tile * p.BN + static_cast<int>(rank * b_half_rows);
      for (int k = 0; k < num_k; ++k) {
        const int4 e = ktab[k];
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (elect_one()) {
          const uint32_t a_dst = smem_base + stage * stage_bytes;
          const uint32_t fb = full_bar(stage) & kPeerMask;  // the leader's barrier
          if (leader) mbar_expect_tx(full_bar(stage), pair_tx_bytes);
          tma2_load_4d(a_dst, &p.a_map[e.x], fb, e.y, x0 + e.z, y0 + e.w, img);
          tma2_load_2d(a_dst + kA2StageBytes, &p.b_map, fb, k * kBlockK, n0);
        }
        __syncwarp();
        if (++stage == static_cast<uint32_t>(stages)) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ------------------------------------------------------------ MMA issuer (leader CTA)
    const uint32_t idesc = make_idesc_bf16(2 * kBlockM, p.BN, false, false);
    uint32_t stage = 0, phase = 0;
    uint32_t acc_it = 0;
    for (int item = pair_id; item < total_items; item += n_pairs, ++acc_it) {
      const uint32_t acc = acc_it & 1u;
      const uint32_t acc_phase = (acc_it >> 1) & 1u;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kAcc2Stride;
      for (int k = 0; k < num_k; ++k) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_base + stage * stage_bytes;
          const uint64_t adesc = make_smem_desc_sw128(a_addr, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(a_addr + kA2StageBytes, 16, 1024);
#pragma unroll
          for (int j = 0; j < kBlockK / 16; ++j)
            umma2_bf16(d_tmem, adesc + 2u * j, bdesc + 2u * j, idesc, (k | j) != 0 ? 1u : 0u);
          umma2_commit_pair(empty_bar(stage));
        }
        __syncwarp();
        if (++stage == static_cast<uint32_t>(stages)) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (elect_one()) umma2_commit_pair(tfull_bar(acc));
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (both CTAs)
    const int wq = warp & 3;
    const int m = wq * 32 + lane;
    const int ly = m / p.TW, lx = m % p.TW;
    const bool in_tile = m < p.TW * p.TH;
    const int et = threadIdx.x - 128;
    uint32_t acc_it = 0;
    for (int item = pair_id; item < total_items; item += n_pairs, ++acc_it) {
      const uint32_t acc = acc_it & 1u;
      const uint32_t acc_phase = (acc_it >> 1) & 1u;
      const int n_tile = item % p.n_tiles_n;
      const int m_raw = 2 * (item / p.n_tiles_n) + static_cast<int>(rank);
      const bool live = m_raw < m_tiles;
      const int m_tile = min(m_raw, m_tiles - 1);
      const int tx = m_tile % p.tiles_x;
      const int ty = (m_tile / p.tiles_x) % p.tiles_y;
      const int img = m_tile / tiles_per_img;
      const int n0 = n_tile * p.BN;
      const int y = ty * p.TH + ly, x = tx * p.TW + lx;
      const bool valid = live && in_tile && y < p.H_out && x < p.W_out;

      float* bias_t = bias_s + acc * 256;
      if (p.bias != nullptr) {
        for (int j = et; j < p.BN; j += 128) {
          const int n = n0 + j;
          bias_t[j] = __ldg(p.bias + (p.shuffle_cout > 0 ? n % p.shuffle_cout : n));
        }
      }
      uint32_t mbits[2][4];
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

      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * kAcc2Stride;
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
      if (lane == 0) {
        if (leader)
          mbar_arrive(tempty_bar(acc));
        else
          mbar_arrive_leader(tempty_bar(acc));
      }
    }
  }

  // Neither CTA may release its shared / tensor memory while the pair's MMAs can still read it --
  // and this final cluster barrier is also what keeps the PEER's shared memory alive for the
  // leader's late multicast arrivals on its empty barriers: the peer's producer has no tail that
  // waits for them, so nothing else orders the peer's exit after the leader's last commit.
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc2(tmem_base, kTmem2Cols);
}

// ------------------------------------------------------------------ host launcher
// Called from rsu_conv_gemm (conv_gemm.cu) with the parameter block already filled for the
// single-CTA kernel; re-encodes the weight map with a BN/2-row box and launches clusters of 2.
int launch_conv_gemm2(ConvGemmParams& p, const void* weights, int ktot, int ntot, cudaStream_t stream) {
  if (p.BN != 256 && p.BN != 128)
    return set_error(RSU_EINVAL, "two-CTA convolution needs an N tile of 128 or 256 (got %d)", p.BN);
  {
    int rc = encode_weight_map(&p.b_map, weights, ktot, ntot, p.BN / 2);
    if (rc) return rc;
  }
  const int num_k = ktot / kBlockK;
  const int stage_bytes = kA2StageBytes + (p.BN / 2) * 128;
  const int fixed = 1024 /*align slack*/ + 8 * (2 * 8 + 4) + 16 + 2 * 256 * 4 + 16 * num_k /*k-step table*/;
  int stages = (226 * 1024 - fixed) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return set_error(RSU_EINVAL, "K = %d too deep for the k-step table", ktot);
  const int smem = fixed + stages * stage_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    RSU_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm2_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const long long m_tiles = 1LL * p.n_img * p.tiles_x * p.tiles_y;
  const long long items = ((m_tiles + 1) / 2) * p.n_tiles_n;
  long long grid = (num_sms() / 2) * 2;
  if (2 * items < grid) grid = 2 * items;
  conv_gemm2_kernel<<<static_cast<unsigned>(grid), kConv2Threads, smem, stream>>>(p, stages);
  return check_launch("conv_gemm2_kernel");
}

}  // namespace rsu
