// Two-CTA variant of the weight-gradient GEMM of wgrad_gemm.cu (sm_100a).
//
// Validated on B200 (tests/test_kernels_gpu.py::test_wgrad_cta_pair: equal to the single-CTA kernel
// up to the split-K summation order) and chosen by rsu_wgrad_gemm for the deep 3x3 layers, where
// it is 1 - 11 % faster (profiles/r2_wgrad_pair_ab.txt).  The CTA-pair protocol is the one of
// conv_gemm2.cu.
//
//   D[(tap, src, c), co] += sum over pixels of  X_src[pixel + tap + off, c] * G[pixel + goff, co]
//
// A cluster of two CTAs shares one gradient (B) tile and one pixel range: each CTA owns its own
// row tile (two 64-row atoms of (tap, 64-channel) blocks: M = 2 x 128 for the pair) and loads
// HALF of the BN gradient channels; tcgen05.mma.cta_group::2 exchanges the halves.  Both
// operands are MN-major (K = pixels) exactly as in the single-CTA kernel.  BiasAddGrad (the column
// sums of G) rides along as one more unit per (gradient tile, pixel slice): a pair whose A operand
// is a constant block of ones in both CTAs (the A descriptor of a cta_group::2 MMA is shared, so
// the ones cannot take a free atom slot of one CTA as they do in wgrad_gemm.cu); row 0 of the
// leader's accumulator goes to the bias gradient.
//
// Reference op replaced: Conv2DBackpropFilter (src/unet.py:34-45, 67, 88-91).
#include "gemm_params.h"
#include "host_common.h"
#include "ptx.cuh"

namespace rsu {

constexpr int kWg2Threads = 256;
constexpr int kWg2TmemCols = 512;
constexpr int kWg2AccStride = 256;
constexpr int kWg2OnesBytes = 4096;  // two 16-row K slices of bf16 ones
constexpr uint32_t kWg2PeerMask = 0xFEFFFFFFu;  // shared::cluster address of the even CTA of a pair

namespace pair {
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
__device__ __forceinline__ void umma2_bf16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi,
                                                uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_pair(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64"
      " [%0], %1;" ::"r"(bar), "h"(mask)
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
}  // namespace pair

struct Atom2 {
  int src, chunk, dy, dx;
};
__device__ __forceinline__ Atom2 decode_atom2(const WgradParams& p, int atom, int chunks_total) {
  Atom2 a;
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

// p.n_tiles_m = row tiles of 2 atoms (no ones atom), p.ksplit computed for pairs by the launcher.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWg2Threads, 1)
    wgrad_gemm2_kernel(const __grid_constant__ WgradParams p, int stages, uint32_t atom_bytes,
                       float* bias_grad) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = pair::ctarank();
  const bool leader = rank == 0;

  const int n_b_half = p.BN / 128;  // 64-channel gradient atoms per CTA
  const uint32_t stage_bytes = static_cast<uint32_t>(2 + n_b_half) * atom_bytes;
  const uint32_t ones_base = smem_base + stages * stage_bytes;
  const uint32_t bar_base = ones_base + kWg2OnesBytes;
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
    // the second atom slot of a row tile with a single atom is read by the MMAs but never
    // loaded: keep it finite (its accumulator rows are not stored)
    uint32_t* z = reinterpret_cast<uint32_t*>(smem_gen);
    for (uint32_t i = threadIdx.x; i < stages * stage_bytes / 4u; i += kWg2Threads) z[i] = 0u;
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem_gen + (ones_base - smem_base));
    for (uint32_t i = threadIdx.x; i < kWg2OnesBytes / 4u; i += kWg2Threads) ones[i] = 0x3F803F80u;
    fence_proxy_async();
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 2);   // leader's A-operand and B-operand producer warps (+ all bytes)
      mbar_init(empty_bar(s), 1);  // one multicast commit per use
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8);  // 4 epilogue warps of each CTA (leader's barrier only)
    }
    fence_mbar_init();
  }
  if (warp == 2) pair::tmem_alloc2(tmem_slot, kWg2TmemCols);
  tc_fence_before();
  pair::cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  int chunks_total = 0;
  for (int s = 0; s < p.n_src; ++s) chunks_total += p.src_chunks[s];
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int pix_tiles = p.n_img * tiles_per_img;
  // (with a bias gradient the unit list is one pair longer: pair index m_real = the ones unit)
  const int m_real = (p.n_tiles_m + 1) >> 1;
  const int m_pairs = m_real + (bias_grad != nullptr ? 1 : 0);
  const int total_units = m_pairs * p.n_tiles_n * p.ksplit;
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int tile_pixels = p.TW * p.TH;
  const uint32_t box_bytes = static_cast<uint32_t>(tile_pixels) * 128u;
  const int mma_per_tile = tile_pixels / 16;

  // unit -> (ks, n_tile, m_pair); m fastest so that concurrently running pairs share the G tile
  auto unit_range = [&](int unit, int* m_pair, int* n_tile, int* pt_begin, int* pt_end) {
    *m_pair = unit % m_pairs;
    const int rest = unit / m_pairs;
    *n_tile = rest % p.n_tiles_n;
    const int ks = rest / p.n_tiles_n;
    *pt_begin = static_cast<int>(1LL * pix_tiles * ks / p.ksplit);
    *pt_end = static_cast<int>(1LL * pix_tiles * (ks + 1) / p.ksplit);
  };
  // real atoms of row tile m_tile: 2, 1 (odd tail) or 0 (the idle half of the last pair)
  auto atoms_of = [&](int m_tile) {
    const int atom0 = m_tile * 2;
    return atom0 + 1 < p.n_atoms ? 2 : (atom0 < p.n_atoms ? 1 : 0);
  };

  if (warp == 0 || warp == 3) {
    // ------------------------------------------------------------ TMA producers (both CTAs)
    // warp 0: this CTA's A atoms; warp 3: this CTA's half of the gradient atoms.  The leader's
    // warps post the expected bytes of BOTH CTAs on the leader's full barrier.
    const bool is_a = warp == 0;
    uint32_t stage = 0, phase = 0;
    for (int unit = pair_id; unit < total_units; unit += n_pairs) {
      int m_pair, n_tile, pt0, pt1;
      unit_range(unit, &m_pair, &n_tile, &pt0, &pt1);
      const int m_tile = 2 * m_pair + static_cast<int>(rank);
      const int atom0 = m_tile * 2;
      const int n_a = atoms_of(m_tile);
      const int n_a_pair = atoms_of(2 * m_pair) + atoms_of(2 * m_pair + 1);
      const Atom2 a0 = decode_atom2(p, n_a >= 1 ? atom0 : 0, chunks_total);
      const Atom2 a1 = decode_atom2(p, n_a == 2 ? atom0 + 1 : 0, chunks_total);
      const int n0 = n_tile * p.BN + static_cast<int>(rank) * (p.BN / 2);
      const uint32_t tx_bytes = static_cast<uint32_t>(is_a ? n_a_pair : 2 * n_b_half) * box_bytes;
      const int bx = p.b_off_x, by = p.b_off_y;
      int img = pt0 / tiles_per_img;
      int r = pt0 % tiles_per_img;
      int ty = r / p.tiles_x, tx = r % p.tiles_x;
      for (int pt = pt0; pt < pt1; ++pt) {
        const int y0 = ty * p.TH, x0 = tx * p.TW;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (elect_one()) {
          const uint32_t dst = smem_base + stage * stage_bytes;
          const uint32_t fb = full_bar(stage) & kWg2PeerMask;  // the leader's barrier
          if (leader) mbar_expect_tx(full_bar(stage), tx_bytes);
          if (is_a) {
            if (n_a >= 1)
              pair::tma2_load_4d(dst, &p.a_map[a0.src], fb, a0.chunk * 64, x0 + a0.dx, y0 + a0.dy, img);
            if (n_a == 2)
              pair::tma2_load_4d(dst + atom_bytes, &p.a_map[a1.src], fb, a1.chunk * 64, x0 + a1.dx,
                                 y0 + a1.dy, img);
          } else {
            for (int j = 0; j < n_b_half; ++j)
              pair::tma2_load_4d(dst + (2 + j) * atom_bytes, &p.b_map, fb, n0 + j * 64, x0 + bx, y0 + by,
                                 img);
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
  } else if (warp == 1 && leader) {
    // ------------------------------------------------------------ MMA issuer (leader CTA)
    const uint32_t idesc = make_idesc_bf16(2 * kBlockM, p.BN, true, true);
    const uint32_t hi = desc_hi_sw128(1024u);
    uint32_t stage = 0, phase = 0;
    uint32_t acc_it = 0;
    for (int unit = pair_id; unit < total_units; unit += n_pairs, ++acc_it) {
      int m_pair, n_tile, pt0, pt1;
      unit_range(unit, &m_pair, &n_tile, &pt0, &pt1);
      const uint32_t acc = acc_it & 1u;
      const uint32_t acc_phase = (acc_it >> 1) & 1u;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kWg2AccStride;
      const bool ones_unit = m_pair == m_real;
      for (int pt = pt0; pt < pt1; ++pt) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_base + stage * stage_bytes;
          // CTA-relative descriptors: the same offsets are valid in both CTAs of the pair.  The
          // ones unit reads the constant block (both 64-row atoms, 2 KiB apart) at every K slice.
          const uint32_t a_lo = ones_unit ? (((ones_base >> 4) & 0x3FFFu) | (128u << 16))
                                          : desc_lo_sw128(a_addr, atom_bytes);
          const uint32_t a_step = ones_unit ? 0u : 128u;
          const uint32_t b_lo = desc_lo_sw128(a_addr + 2 * atom_bytes, atom_bytes);
          const uint32_t first = pt != pt0 ? 1u : 0u;
          // 16 pixels (K) per instruction = 16 rows of 128 B = 2 KiB further into every atom
          if (mma_per_tile == 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              pair::umma2_bf16_lohi(d_tmem, a_lo + j * a_step, hi, b_lo + j * 128u, hi, idesc,
                                    j != 0 ? 1u : first);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (j < mma_per_tile)
                pair::umma2_bf16_lohi(d_tmem, a_lo + j * a_step, hi, b_lo + j * 128u, hi, idesc,
                                      j != 0 ? 1u : first);
            }
          }
          pair::umma2_commit_pair(empty_bar(stage));
        }
        __syncwarp();
        if (++stage == static_cast<uint32_t>(stages)) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (elect_one()) pair::umma2_commit_pair(tfull_bar(acc));
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (both CTAs): atomics
    const int wq = warp & 3;
    const int m = wq * 32 + lane;
    uint32_t acc_it = 0;
    for (int unit = pair_id; unit < total_units; unit += n_pairs, ++acc_it) {
      int m_pair, n_tile, pt0, pt1;
      unit_range(unit, &m_pair, &n_tile, &pt0, &pt1);
      const uint32_t acc = acc_it & 1u;
      const uint32_t acc_phase = (acc_it >> 1) & 1u;
      const int atom = (2 * m_pair + static_cast<int>(rank)) * 2 + (m >> 6);
      // the ones unit: every row of the leader's accumulator holds the column sums of G; row 0
      // goes to the bias gradient, the peer's half of the pair has nothing to store
      const bool is_bias = m_pair == m_real && leader && m == 0;
      const bool valid = ((m_pair < m_real && atom < p.n_atoms) || is_bias) && pt1 > pt0;
      float* orow = is_bias ? bias_grad + n_tile * p.BN
                            : p.out + (static_cast<long long>(atom) * 64 + (m & 63)) * p.ldo + n_tile * p.BN;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row =
          tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * kWg2AccStride;
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
      if (lane == 0) {
        if (leader)
          mbar_arrive(tempty_bar(acc));
        else
          pair::mbar_arrive_leader(tempty_bar(acc));
      }
    }
  }

  tc_fence_before();
  pair::cluster_sync_all();
  if (warp == 2) pair::tmem_dealloc2(tmem_base, kWg2TmemCols);
}

// Called from rsu_wgrad_gemm (wgrad_gemm.cu) with maps, taps, tile and output already filled in p
// (p.n_atoms set, BN chosen); fixes the row-tile count (no ones atom), the split-K factor for
// pairs and launches clusters of 2.
int launch_wgrad_gemm2(WgradParams& p, float* bias_grad, cudaStream_t stream) {
  if (p.BN != 256 && p.BN != 128)
    return set_error(RSU_EINVAL, "two-CTA weight gradient needs an N tile of 128 or 256 (got %d)", p.BN);
  p.n_tiles_m = (p.n_atoms + 1) / 2;
  const int m_pairs = (p.n_tiles_m + 1) / 2 + (bias_grad != nullptr ? 1 : 0);
  const int pix_tiles = p.n_img * p.tiles_x * p.tiles_y;
  const int mn_units = m_pairs * p.n_tiles_n;
  const int pairs = num_sms() / 2;
  p.ksplit = choose_ksplit(mn_units, pix_tiles, pairs, (p.TW * p.TH / 16) * (p.BN / 2.0),
                           1500.0 + 12.0 * p.BN, 4);
  const int atom_bytes = ((p.TW * p.TH * 128) + 1023) & ~1023;
  const int stage_bytes = (2 + p.BN / 128) * atom_bytes;
  int stages = (220 * 1024 - kWg2OnesBytes) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  const int smem = 1024 + stages * stage_bytes + kWg2OnesBytes + 8 * (2 * stages + 4) + 16;
  static bool attr_set = false;
  if (!attr_set) {
    RSU_CHECK_CUDA(cudaFuncSetAttribute(wgrad_gemm2_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const long long total = 1LL * mn_units * p.ksplit;
  long long grid = 2LL * pairs;
  if (2 * total < grid) grid = 2 * total;
  wgrad_gemm2_kernel<<<static_cast<unsigned>(grid), kWg2Threads, smem, stream>>>(
      p, stages, static_cast<uint32_t>(atom_bytes), bias_grad);
  return check_launch("wgrad_gemm2_kernel");
}

}  // namespace rsu
