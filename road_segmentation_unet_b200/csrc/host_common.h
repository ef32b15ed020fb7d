// Host-side helpers shared by the C-ABI translation units: error reporting, launch counting,
// TMA tensor-map encoding through the driver entry point (no link-time libcuda dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/rsu_b200.h"

namespace rsu {

int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);  // cudaGetLastError -> RSU_ECUDA
int num_sms();

// 4-D bf16 activation map: dims (C, W, H, N), box {64, box_w, box_h, 1}, SWIZZLE_128B.
int encode_act_map(CUtensorMap* map, const rsu_view& v, int box_w, int box_h);
// 2-D bf16 weight map: dims (K, N), row stride K, box {64, box_n}, SWIZZLE_128B.
int encode_weight_map(CUtensorMap* map, const void* ptr, int K, int N, int box_n);

// Spatial tile (TW x TH <= max_pixels <= 128, TW <= max_tw, TH <= max_th) minimising the tile count over a
// W x H grid.  If mult16 is set TW*TH must be a multiple of 16 (pixels are the GEMM K dimension
// in the weight gradient).  Returns 0 x 0 when no admissible tile exists.
void pick_tile(int W, int H, int max_tw, int max_th, bool mult16, int* TW, int* TH,
               int max_pixels = 128);

// Split-K factor for `mn_units` independent output tiles, each reducing over `k_tiles` tiles, on
// `sms` persistent CTAs (unit u runs on CTA u % sms): minimises
//   waves x (tiles per unit x tile_cost + unit_cost),  waves = ceil(k * mn_units / sms),
// i.e. it trades the tail of a partially filled last wave against the per-unit epilogue.
int choose_ksplit(int mn_units, int k_tiles, int sms, double tile_cost, double unit_cost,
                  int min_tiles);

#define RSU_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return rsu::set_error(RSU_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));   \
  } while (0)

}  // namespace rsu
