#include "host_common.h"

#include <atomic>
#include <mutex>
#include <string.h>

namespace rsu {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
// launches per kernel name (the string literals handed to check_launch), for rsu_launch_histogram
constexpr int kMaxNames = 64;
static const char* g_names[kMaxNames];
static long long g_name_counts[kMaxNames];
static int g_n_names = 0;
static std::mutex g_name_mu;

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(RSU_ECUDA, "%s: %s", what, cudaGetErrorString(e));
  count_launch();
  {
    std::lock_guard<std::mutex> lock(g_name_mu);
    int i = 0;
    while (i < g_n_names && g_names[i] != what && strcmp(g_names[i], what) != 0) ++i;
    if (i == g_n_names && g_n_names < kMaxNames) {
      g_names[g_n_names] = what;
      g_name_counts[g_n_names++] = 0;
    }
    if (i < kMaxNames) ++g_name_counts[i];
  }
  return RSU_OK;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
  }
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int encode_act_map(CUtensorMap* map, const rsu_view& v, int box_w, int box_h) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(RSU_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if (v.C % 64 != 0) return set_error(RSU_EINVAL, "view channels %d not a multiple of 64", v.C);
  if ((reinterpret_cast<uintptr_t>(v.ptr) & 15) || (v.sx % 8) || (v.sy % 8) || (v.sn % 8))
    return set_error(RSU_EALIGN, "view pointer/strides must be 16-byte aligned");
  if (box_w < 1 || box_w > 256 || box_h < 1 || box_h > 256)
    return set_error(RSU_EINVAL, "bad TMA box %dx%d", box_w, box_h);
  cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.N};
  cuuint64_t strides[3] = {(cuuint64_t)v.sx * 2, (cuuint64_t)v.sy * 2, (cuuint64_t)v.sn * 2};
  // TMA wants non-degenerate strides even for extent-1 dims
  if (v.N == 1 && strides[2] == 0) strides[2] = strides[1] * v.H;
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(v.ptr), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(RSU_ECUDA,
                     "cuTensorMapEncodeTiled(act C=%d W=%d H=%d N=%d sx=%lld sy=%lld sn=%lld box "
                     "%dx%d) -> %d",
                     v.C, v.W, v.H, v.N, v.sx, v.sy, v.sn, box_w, box_h, (int)r);
  return RSU_OK;
}

int encode_weight_map(CUtensorMap* map, const void* ptr, int K, int N, int box_n) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(RSU_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if (K % 64 != 0) return set_error(RSU_EINVAL, "weight K %d not a multiple of 64", K);
  if (reinterpret_cast<uintptr_t>(ptr) & 15)
    return set_error(RSU_EALIGN, "weight pointer must be 16-byte aligned");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_n};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(RSU_ECUDA, "cuTensorMapEncodeTiled(weights K=%d N=%d box %d) -> %d", K, N,
                     box_n, (int)r);
  return RSU_OK;
}

void pick_tile(int W, int H, int max_tw, int max_th, bool mult16, int* TW, int* TH, int max_pixels) {
  long best_tiles = -1;
  int bw = 0, bh = 0;
  if (max_pixels > 128 || max_pixels < 1) max_pixels = 128;
  if (max_tw > max_pixels) max_tw = max_pixels;
  for (int tw = max_tw; tw >= 1; --tw) {
    if (tw < 8 && max_tw >= 8) break;  // keep contiguous runs >= 1 KiB when possible
    int th_max = max_pixels / tw;
    if (th_max > max_th) th_max = max_th;
    for (int th = th_max; th >= 1; --th) {
      if (mult16 && (tw * th) % 16 != 0) continue;
      const long tiles = (long)((W + tw - 1) / tw) * ((H + th - 1) / th);
      // fewest tiles wins (every tile costs a full 128-row MMA); ties keep the wider tile
      if (best_tiles < 0 || tiles < best_tiles) {
        best_tiles = tiles;
        bw = tw;
        bh = th;
      }
    }
  }
  *TW = bw;  // 0 x 0 when no admissible tile exists
  *TH = bh;
}

int choose_ksplit(int mn_units, int k_tiles, int sms, double tile_cost, double unit_cost,
                  int min_tiles) {
  if (min_tiles < 1) min_tiles = 1;
  int kmax = k_tiles / min_tiles;
  if (kmax < 1) kmax = 1;
  if (kmax > 8192) kmax = 8192;
  int best = 1;
  double best_cost = -1.0;
  for (int k = 1; k <= kmax; ++k) {
    const long long units = 1LL * k * mn_units;
    const long long waves = (units + sms - 1) / sms;
    const double per_unit = static_cast<double>((k_tiles + k - 1) / k) * tile_cost + unit_cost;
    const double cost = static_cast<double>(waves) * per_unit;
    if (best_cost < 0.0 || cost < best_cost * (1.0 - 1e-9)) {
      best_cost = cost;
      best = k;
    }
  }
  return best;
}

}  // namespace rsu

namespace rsu {
// CRC-32C (Castagnoli, reflected polynomial 0x82F63B78), slicing-by-8 on the host.
static uint32_t g_crc_tab[8][256];
static std::once_flag g_crc_once;
static void crc_init() {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c >> 1) ^ ((c & 1) ? 0x82F63B78u : 0u);
    g_crc_tab[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t)
      g_crc_tab[t][i] = (g_crc_tab[t - 1][i] >> 8) ^ g_crc_tab[0][g_crc_tab[t - 1][i] & 0xff];
}
}  // namespace rsu

extern "C" {
unsigned int rsu_crc32c_host(unsigned int crc, const void* data_host, unsigned long long n) {
  std::call_once(rsu::g_crc_once, rsu::crc_init);
  const unsigned char* p = static_cast<const unsigned char*>(data_host);
  uint32_t c = ~crc;
  while (n && (reinterpret_cast<uintptr_t>(p) & 7)) {
    c = (c >> 8) ^ rsu::g_crc_tab[0][(c ^ *p++) & 0xff];
    --n;
  }
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= c;
    c = rsu::g_crc_tab[7][w & 0xff] ^ rsu::g_crc_tab[6][(w >> 8) & 0xff] ^
        rsu::g_crc_tab[5][(w >> 16) & 0xff] ^ rsu::g_crc_tab[4][(w >> 24) & 0xff] ^
        rsu::g_crc_tab[3][(w >> 32) & 0xff] ^ rsu::g_crc_tab[2][(w >> 40) & 0xff] ^
        rsu::g_crc_tab[1][(w >> 48) & 0xff] ^ rsu::g_crc_tab[0][(w >> 56) & 0xff];
    p += 8;
    n -= 8;
  }
  while (n--) c = (c >> 8) ^ rsu::g_crc_tab[0][(c ^ *p++) & 0xff];
  return ~c;
}
const char* rsu_last_error(void) { return rsu::g_err; }
int rsu_version(void) { return 100; }
long long rsu_launch_count(void) { return rsu::g_launches.load(); }
void rsu_reset_launch_count(void) {
  rsu::g_launches.store(0);
  std::lock_guard<std::mutex> lock(rsu::g_name_mu);
  for (int i = 0; i < rsu::g_n_names; ++i) rsu::g_name_counts[i] = 0;
}
int rsu_launch_histogram(char* buf_host, int cap) {
  std::lock_guard<std::mutex> lock(rsu::g_name_mu);
  int need = 1;
  if (buf_host && cap > 0) buf_host[0] = 0;
  for (int i = 0; i < rsu::g_n_names; ++i) {
    if (rsu::g_name_counts[i] == 0) continue;
    char item[160];
    const int n = snprintf(item, sizeof(item), "%s=%lld;", rsu::g_names[i], rsu::g_name_counts[i]);
    if (buf_host && need + n <= cap) strcat(buf_host, item);
    need += n;
  }
  return need;
}
}
