// Kernel parameter blocks for the two tcgen05 kernels (passed as __grid_constant__).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace rsu {

constexpr int kMaxSrc = 4;    // sources walked by the K loop (fused crop + concat)
constexpr int kMaxTaps = 9;   // 3x3 filter taps (1 for 1x1 / transpose-conv GEMMs)
constexpr int kBlockM = 128;  // accumulator rows = TMEM lanes
constexpr int kBlockK = 64;   // bf16 elements per 128-byte swizzle row

// Implicit-GEMM convolution:  D[pixel, n] = sum_{tap, src, c} A_src[pixel + tap + off_src, c] * B[n, k]
// with k = (tap, src, c) flattened in exactly that loop order.
struct ConvGemmParams {
  CUtensorMap a_map[kMaxSrc];  // 4-D (C, W, H, N) bf16, SWIZZLE_128B, box {64, TW, TH, 1}
  CUtensorMap b_map;           // 2-D (Ktot, Ntot) bf16, SWIZZLE_128B, box {64, BN}
  int n_src;
  int src_chunks[kMaxSrc];  // channels / 64 per source
  int src_off_y[kMaxSrc];   // crop offset added to the pixel coordinate
  int src_off_x[kMaxSrc];
  int n_taps;
  int tap_dy[kMaxTaps];
  int tap_dx[kMaxTaps];
  int TW, TH;                   // spatial tile (TW*TH <= 128 pixels = accumulator rows)
  int tiles_x, tiles_y, n_img;  // tile grid over the output
  int n_tiles_n, BN;            // N tiling (BN divides Ntot, BN % 16 == 0, BN <= 256)
  int H_out, W_out;             // logical output extent (store mask)
  // output: element (img, y, x, n) -> out + img*out_sn + y*out_sy + x*out_sx + n   (bf16)
  __nv_bfloat16* out;
  long long out_sn, out_sy, out_sx;
  // transpose-conv epilogue: n = (a, b, co) with co < shuffle_cout; pixel (y,x) -> (2y+a, 2x+b)
  int shuffle_cout;
  const float* bias;  // [Ntot] (or [shuffle_cout] when shuffling), may be null
  int relu;
  // optional ReLU-gradient mask: keep value only where mask[...] > 0 (same indexing as out)
  const __nv_bfloat16* mask;
  long long mask_sn, mask_sy, mask_sx;
  int mask_c0, mask_nc;  // mask_nc > 0: mask covers output channels [mask_c0, mask_c0 + mask_nc) only
  int accumulate;  // out += result (read-modify-write)
};

// Weight-gradient GEMM (both operands MN-major, K = pixels):
//   D[(tap, src, c), n] += sum_{pixel} A_src[pixel + tap + off_src, c] * B[pixel + b_off, n]
struct WgradParams {
  CUtensorMap a_map[kMaxSrc];  // 4-D (C, W, H, N) bf16, box {64, TW, TH, 1}
  CUtensorMap b_map;           // 4-D (C, W, H, N) bf16, box {64, TW, TH, 1}
  int n_src;
  int src_chunks[kMaxSrc];
  int src_off_y[kMaxSrc];
  int src_off_x[kMaxSrc];
  int n_taps;
  int tap_dy[kMaxTaps];
  int tap_dx[kMaxTaps];
  int b_off_y, b_off_x;
  int TW, TH;  // TW*TH % 16 == 0, TW*TH <= 128
  int tiles_x, tiles_y, n_img;
  int n_atoms;    // n_taps * sum(src_chunks): 64-row blocks of the output
  int n_tiles_m;  // ceil(n_atoms / 2)
  int n_tiles_n, BN;
  int ksplit;     // pixel tiles are divided among ksplit CTAs; results combined with atomics
  float* out;     // [n_atoms*64, ldo] fp32, must be zero-initialised (or hold a running sum)
  int ldo;
};

}  // namespace rsu
