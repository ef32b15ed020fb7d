// Geometry kernels replacing the NumPy / SciPy helpers of src/images.py (fp32 images, NHWC).
// All of them are pure data movement (bit-exact against the reference) except the two
// averaging kernels, which accumulate in fp64 in the reference's summation order.
//
// They are HBM-bound, so every kernel is built around bytes in flight: one block per image row
// (or 32 x 32 tile), 32-bit index arithmetic, 16-byte accesses where the row alignment allows it
// and several independent loads issued per thread before the first use.
#include "host_common.h"

namespace rsu {

// np.pad(..., "symmetric") index: reflect with the edge sample repeated, period 2*L.  The
// extension is symmetric about -1/2, so f(-1 - j) = f(j); the modulo only runs when the pad is
// wider than the image.
__device__ __forceinline__ int sym_index(int i, int L) {
  if (i < 0) i = -1 - i;
  if (i >= L) {
    const int m = i % (2 * L);
    i = m < L ? m : 2 * L - 1 - m;
  }
  return i;
}

__device__ __forceinline__ bool aligned16(const void* p) {
  return (reinterpret_cast<uintptr_t>(p) & 15) == 0;
}

// images.mirror_border (images.py:269-281).  One block per output row; a thread moves groups of
// four consecutive floats: interior groups are one 16-byte load (source contiguous) and border
// groups four reflected scalar loads; stores are 16 bytes.  VEC = 0: rows whose length or base
// is not 16-byte aligned use the same walk with scalar accesses.
template <int C_T, int VEC>
__global__ void __launch_bounds__(256)
    mirror_pad_kernel(const float* __restrict__ in, int H, int W, int C_rt, int pad,
                      float* __restrict__ out) {
  const int C = C_T ? C_T : C_rt;
  const int Ho = H + 2 * pad, Wo = W + 2 * pad;
  const int row = blockIdx.x;
  const int n = row / Ho, y = row - n * Ho;
  const int sy = sym_index(y - pad, H);
  const float* __restrict__ src = in + (1LL * n * H + sy) * W * C;
  float* __restrict__ dst = out + 1LL * row * Wo * C;
  const int row_elems = Wo * C;
  const int lo = pad * C, hi = (pad + W) * C;
  auto fetch = [&](int e) -> float {
    const int x = e / C, c = e - x * C;
    return __ldg(src + sym_index(x - pad, W) * C + c);
  };
  if (VEC) {
    const bool src_vec = aligned16(src);  // (lo is a multiple of 4 floats whenever VEC is chosen)
    constexpr int U = 2;
    const int n4 = row_elems >> 2;
    for (int g0 = threadIdx.x; g0 < n4; g0 += blockDim.x * U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int g = g0 + u * blockDim.x;
        const int e = g << 2;
        if (g < n4) {
          if (e >= lo && e + 3 < hi && src_vec && ((e - lo) & 3) == 0) {
            v[u] = __ldg(reinterpret_cast<const float4*>(src + (e - lo)));
          } else if (e >= lo && e + 3 < hi) {
            const float* s = src + (e - lo);
            v[u] = make_float4(__ldg(s), __ldg(s + 1), __ldg(s + 2), __ldg(s + 3));
          } else {
            v[u] = make_float4(fetch(e), fetch(e + 1), fetch(e + 2), fetch(e + 3));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int g = g0 + u * blockDim.x;
        if (g < n4) reinterpret_cast<float4*>(dst)[g] = v[u];
      }
    }
  } else {
    constexpr int U = 8;
    for (int e0 = threadIdx.x; e0 < row_elems; e0 += blockDim.x * U) {
      float v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = e0 + u * blockDim.x;
        if (e < row_elems) v[u] = (e >= lo && e < hi) ? __ldg(src + (e - lo)) : fetch(e);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = e0 + u * blockDim.x;
        if (e < row_elems) dst[e] = v[u];
      }
    }
  }
}

// Dihedral-group transform of square images, one op per image:
//   out = rot90(flipud(x) if (op & 4) else x, k = op & 3)     (counter-clockwise, like np.rot90)
// A T x T-pixel destination tile maps onto a T x T source tile; the source tile is read row-wise
// (coalesced) into shared memory and written out row-wise in destination order.
__device__ __forceinline__ void d4_src(int op, int S, int i, int j, int* si, int* sj) {
  int a, b;
  switch (op & 3) {
    case 0: a = i; b = j; break;
    case 1: a = j; b = S - 1 - i; break;
    case 2: a = S - 1 - i; b = S - 1 - j; break;
    default: a = S - 1 - j; b = i; break;
  }
  if (op & 4) a = S - 1 - a;
  *si = a;
  *sj = b;
}

// source tile of the destination tile [i0, i0+T) x [j0, j0+T): origin and extent
__device__ __forceinline__ void d4_src_tile(int op, int S, int i0, int j0, int i1, int j1, int* si0,
                                            int* sj0, int* sh, int* sw) {
  int ci[2], cj[2];
  d4_src(op, S, i0, j0, &ci[0], &cj[0]);
  d4_src(op, S, i1, j1, &ci[1], &cj[1]);
  *si0 = min(ci[0], ci[1]);
  *sj0 = min(cj[0], cj[1]);
  *sh = abs(ci[0] - ci[1]) + 1;
  *sw = abs(cj[0] - cj[1]) + 1;
}

// WORDS (elements of WordT per pixel) and T (tile side) are compile-time: every thread of the
// 32 x 8 block issues all of its T/8 * T*WORDS/32 loads before the first shared-memory store.
template <typename WordT, int WORDS, int T>
__global__ void __launch_bounds__(256)
    d4_transform_kernel(const WordT* __restrict__ in, WordT* __restrict__ out, int S,
                        const unsigned char* __restrict__ ops, int n_in) {
  constexpr int RW = T * WORDS;     // words per tile row
  constexpr int PITCH = RW + 1;
  constexpr int QW = (RW + 31) / 32;
  constexpr int QR = T / 8;
  __shared__ WordT tile[T * PITCH];
  const int n = blockIdx.z;
  const int op = ops[n];
  const int i0 = blockIdx.y * T, j0 = blockIdx.x * T;
  const int i1 = min(i0 + T - 1, S - 1), j1 = min(j0 + T - 1, S - 1);
  int si0, sj0, sh, sw;
  d4_src_tile(op, S, i0, j0, i1, j1, &si0, &sj0, &sh, &sw);
  // (output image n reads input image n % n_in: the 6-way ensemble transforms every input six times)
  const WordT* __restrict__ src = in + (1LL * (n % n_in) * S + si0) * S * WORDS + 1LL * sj0 * WORDS;
  WordT* __restrict__ dst = out + (1LL * n * S + i0) * S * WORDS + 1LL * j0 * WORDS;
  const int tx = threadIdx.x, ty = threadIdx.y;
  WordT v[QR][QW];
#pragma unroll
  for (int r = 0; r < QR; ++r)
#pragma unroll
    for (int q = 0; q < QW; ++q) {
      const int rr = ty + 8 * r, w = tx + 32 * q;
      if (rr < sh && w < sw * WORDS) v[r][q] = __ldg(src + 1LL * rr * S * WORDS + w);
    }
#pragma unroll
  for (int r = 0; r < QR; ++r)
#pragma unroll
    for (int q = 0; q < QW; ++q) {
      const int rr = ty + 8 * r, w = tx + 32 * q;
      if (rr < sh && w < sw * WORDS) tile[rr * PITCH + w] = v[r][q];
    }
  __syncthreads();
  const int th = i1 - i0 + 1, tw = j1 - j0 + 1;
#pragma unroll
  for (int r = 0; r < QR; ++r)
#pragma unroll
    for (int q = 0; q < QW; ++q) {
      const int rr = ty + 8 * r, w = tx + 32 * q;
      if (rr < th && w < tw * WORDS) {
        const int j = w / WORDS, e = w - j * WORDS;
        int si, sj;
        d4_src(op, S, i0 + rr, j0 + j, &si, &sj);
        dst[1LL * rr * S * WORDS + w] = tile[(si - si0) * PITCH + (sj - sj0) * WORDS + e];
      }
    }
}

// The same transform with VB-byte global accesses (16 for fp32 pixels, 4 for uint8 label masks):
// the element-wise kernel above issues ~25 instructions per 4 bytes moved and is bound by
// instruction issue at half of the HBM bandwidth.  Needs S * pixel bytes to be a multiple of VB
// (then every tile row of the source and of the destination starts VB-aligned) and VB-aligned
// base pointers.  The source tile is read row-wise into shared memory as vectors; every thread
// then gathers the words of one destination vector through the affine map
//   (source row, source column) - tile origin = (ai, aj) * r + (bi, bj) * j + (ci, cj)
// of the dihedral element, whose coefficients are computed once per block.
template <typename WordT, int WORDS, int T, typename VecT>
__global__ void __launch_bounds__(256)
    d4_transform_vec_kernel(const WordT* __restrict__ in, WordT* __restrict__ out, int S,
                            const unsigned char* __restrict__ ops, int n_in) {
  constexpr int VW = sizeof(VecT) / sizeof(WordT);  // words per vector
  constexpr int RW = T * WORDS;                     // words per full tile row
  constexpr int RV = RW / VW;                       // vectors per full tile row
  constexpr int PITCH = RW + (sizeof(WordT) == 4 ? 1 : 4);
  constexpr int Q = (T * RV + 255) / 256;
  static_assert(RW % VW == 0, "tile rows must be whole vectors");
  __shared__ WordT tile[T * PITCH];
  const int n = blockIdx.z;
  const int op = ops[n];
  const int i0 = blockIdx.y * T, j0 = blockIdx.x * T;
  const int i1 = min(i0 + T - 1, S - 1), j1 = min(j0 + T - 1, S - 1);
  int si0, sj0, sh, sw;
  d4_src_tile(op, S, i0, j0, i1, j1, &si0, &sj0, &sh, &sw);
  const WordT* __restrict__ src = in + (1LL * (n % n_in) * S + si0) * S * WORDS + 1LL * sj0 * WORDS;
  WordT* __restrict__ dst = out + (1LL * n * S + i0) * S * WORDS + 1LL * j0 * WORDS;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const long long row_words = 1LL * S * WORDS;
  // ---- source tile -> shared memory, row-wise, Q vectors per thread in flight
  const int src_rv = sw * WORDS / VW;  // vectors per source tile row (== RV for interior tiles)
  VecT v[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int idx = tid + q * 256;
    const int rr = src_rv == RV ? idx / RV : idx / src_rv;
    const int c = idx - rr * src_rv;
    if (rr < sh) v[q] = __ldg(reinterpret_cast<const VecT*>(src + rr * row_words) + c);
  }
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int idx = tid + q * 256;
    const int rr = src_rv == RV ? idx / RV : idx / src_rv;
    const int c = idx - rr * src_rv;
    if (rr < sh) {
      const WordT* w = reinterpret_cast<const WordT*>(&v[q]);
#pragma unroll
      for (int k = 0; k < VW; ++k) tile[rr * PITCH + c * VW + k] = w[k];
    }
  }
  __syncthreads();
  // ---- affine map of the block: shared-memory word index of destination (r, j, e)
  int o_i, o_j, r_i, r_j, c_i, c_j;
  d4_src(op, S, i0, j0, &o_i, &o_j);
  d4_src(op, S, i0 + 1, j0, &r_i, &r_j);
  d4_src(op, S, i0, j0 + 1, &c_i, &c_j);
  const int step_r = (r_i - o_i) * PITCH + (r_j - o_j) * WORDS;
  const int step_j = (c_i - o_i) * PITCH + (c_j - o_j) * WORDS;
  const int base = (o_i - si0) * PITCH + (o_j - sj0) * WORDS;
  const int th = i1 - i0 + 1, tw = j1 - j0 + 1;
  const int dst_rv = tw * WORDS / VW;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int idx = tid + q * 256;
    const int rr = dst_rv == RV ? idx / RV : idx / dst_rv;
    const int c = idx - rr * dst_rv;
    if (rr < th) {
      VecT o;
      WordT* w = reinterpret_cast<WordT*>(&o);
#pragma unroll
      for (int k = 0; k < VW; ++k) {
        const int word = c * VW + k;
        const int j = word / WORDS, e = word - j * WORDS;
        w[k] = tile[base + rr * step_r + j * step_j + e];
      }
      reinterpret_cast<VecT*>(dst + rr * row_words)[c] = o;
    }
  }
}

// any pixel size (fp64 images travel as 2 x 32-bit words per element): run-time word count
template <typename WordT>
__global__ void d4_transform_generic_kernel(const WordT* __restrict__ in, WordT* __restrict__ out,
                                            int S, int words, const unsigned char* __restrict__ ops,
                                            int n_in) {
  extern __shared__ uint8_t tile_raw[];  // [32][32*words + 1] words
  WordT* tile = reinterpret_cast<WordT*>(tile_raw);
  const int n = blockIdx.z;
  const int op = ops[n];
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int i1 = min(i0 + 31, S - 1), j1 = min(j0 + 31, S - 1);
  int si0, sj0, sh, sw;
  d4_src_tile(op, S, i0, j0, i1, j1, &si0, &sj0, &sh, &sw);
  const int pitch = 32 * words + 1;
  const WordT* src = in + 1LL * (n % n_in) * S * S * words;
  WordT* dst = out + 1LL * n * S * S * words;
  for (int r = threadIdx.y; r < sh; r += blockDim.y)
    for (int w = threadIdx.x; w < sw * words; w += blockDim.x)
      tile[r * pitch + w] = __ldg(src + (1LL * (si0 + r) * S + sj0) * words + w);
  __syncthreads();
  const int th = i1 - i0 + 1, tw = j1 - j0 + 1;
  for (int r = threadIdx.y; r < th; r += blockDim.y)
    for (int w = threadIdx.x; w < tw * words; w += blockDim.x) {
      const int j = w / words, e = w - j * words;
      int si, sj;
      d4_src(op, S, i0 + r, j0 + j, &si, &sj);
      dst[(1LL * (i0 + r) * S + j0) * words + w] = tile[(si - si0) * pitch + (sj - sj0) * words + e];
    }
}

// images.extract_patches (images.py:35-85): patch k of image n sits at column (k / side)*stride,
// row (k % side)*stride  (x is the OUTER loop in the reference).  One block per patch row: the
// source is a contiguous run of P*C floats, moved 16 bytes at a time when both ends allow it.
template <int VEC>
__global__ void __launch_bounds__(128)
    extract_patches_kernel(const float* __restrict__ in, int H, int W, int C, int P, int stride,
                           int side, long long k_begin, float* __restrict__ out) {
  const int row = blockIdx.x;  // (local patch, patch row)
  const int kl = row / P, py = row - kl * P;
  const long long k_all = k_begin + kl;
  const int per_img = side * side;
  const int n = static_cast<int>(k_all / per_img);
  const int k = static_cast<int>(k_all - 1LL * n * per_img);
  const int kx = k / side, ky = k - kx * side;
  const float* __restrict__ src = in + ((1LL * n * H + ky * stride + py) * W + kx * stride) * C;
  float* __restrict__ dst = out + 1LL * row * P * C;
  const int n_el = P * C;
  if (VEC && aligned16(src)) {  // (dst rows are 16-byte aligned whenever VEC is chosen)
    constexpr int U = 4;
    const int n4 = n_el >> 2;
    for (int g0 = threadIdx.x; g0 < n4; g0 += blockDim.x * U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int g = g0 + u * blockDim.x;
        if (g < n4) v[u] = __ldg(reinterpret_cast<const float4*>(src) + g);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int g = g0 + u * blockDim.x;
        if (g < n4) reinterpret_cast<float4*>(dst)[g] = v[u];
      }
    }
  } else {
    constexpr int U = 8;
    for (int e0 = threadIdx.x; e0 < n_el; e0 += blockDim.x * U) {
      float v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = e0 + u * blockDim.x;
        if (e < n_el) v[u] = __ldg(src + e);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = e0 + u * blockDim.x;
        if (e < n_el) dst[e] = v[u];
      }
    }
  }
}

// images.images_from_patches (images.py:131-164) in gather form: every output pixel sums the
// patches covering it in fp64 and divides by the hit count -- deterministic, no atomics.  A
// pixel is covered by up to (P/stride)^2 patches (1,089 at P = 388, stride 12) that lie
// ~600 KB apart, so the kernel is organised around which patch the whole block is reading:
//   * a warp owns 32 x VEC consecutive elements of one output row, the 4 rows of a block are
//     consecutive, and all of them walk the same warp-uniform (kx, ky) patch list with lanes
//     outside a patch predicated off -- every request is one contiguous 128 x VEC byte piece
//     of one patch row, and the block reads 4 adjacent rows of the same patch together;
//   * the patch list is flattened and taken eight patches at a time (eight independent
//     16-byte loads per lane in flight), the two warps that share a row take alternate groups
//     and their partial sums are added in a fixed order through shared memory.
// VEC = 4 needs P*C and stride*C to be multiples of 4 (every patch boundary then falls between
// lanes and all addresses are 16-byte aligned); VEC = 1 is the general form.
// k_lo/k_hi restrict the sum to patches whose global index (n*side*side + k) lies in
// [k_lo, k_hi) -- a rank of a sharded prediction holds only that slice (patches points at patch
// k_lo) and emits partial sums (normalize = 0) that are added and divided after the gather.
template <int VEC>
__global__ void __launch_bounds__(256)
    overlap_average_kernel(const float* __restrict__ patches, int side, int P, int C, int stride,
                           int S, long long total_rows, int n_chunks, long long k_lo,
                           long long k_hi, int normalize, float* __restrict__ out) {
  constexpr int U = 8, ROWS = 4;
  __shared__ double part[ROWS][VEC][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = warp & (ROWS - 1), half = warp >> 2;
  const int chunk = blockIdx.x % n_chunks;
  const long long row = 1LL * (blockIdx.x / n_chunks) * ROWS + r;  // n*S + y
  const int PC = P * C, sC = stride * C, row_elems = S * C;
  const long long patch_elems = 1LL * P * PC;
  const int xcw0 = chunk * 32 * VEC;                          // first element of the warp
  const int xcw1 = min(xcw0 + 32 * VEC, row_elems) - 1;       // last one
  const int xc = xcw0 + lane * VEC;
  const bool ok = row < total_rows && xc < row_elems;
  double acc[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) acc[e] = 0.0;
  int cnt = 1;
  if (row < total_rows) {
    const int n = static_cast<int>(row / S), y = static_cast<int>(row - 1LL * n * S);
    const long long img_k0 = 1LL * n * side * side;
    // patches covering this row / this lane / any lane of the warp (element units along x)
    const int ky_lo = (y - P + 1 <= 0) ? 0 : (y - P + stride) / stride;
    const int ky_hi = min(y / stride, side - 1);
    const int kx_lo = (xc - PC + 1 <= 0) ? 0 : (xc - PC + sC) / sC;
    const int kx_hi = min(xc / sC, side - 1);
    cnt = (kx_hi - kx_lo + 1) * (ky_hi - ky_lo + 1);
    int wkx_lo = (xcw0 - PC + 1 <= 0) ? 0 : (xcw0 - PC + sC) / sC;
    int wkx_hi = min(xcw1 / sC, side - 1);
    // this rank's slice in image-local patch indices k = kx*side + ky, and the patch columns
    // that can hold one of its patches
    const long long a = k_lo - img_k0, b = k_hi - img_k0, kk = 1LL * side * side;
    const int kl = static_cast<int>(min(max(a, 0LL), kk));
    const unsigned kspan = static_cast<unsigned>(static_cast<int>(min(max(b, 0LL), kk)) - kl);
    wkx_lo = max(wkx_lo, kl / side);
    wkx_hi = kspan ? min(wkx_hi, (kl + static_cast<int>(kspan) - 1) / side) : -1;
    // the two warps of a row split the patch columns; each walks its flattened (kx, ky) list
    const int nky = ky_hi - ky_lo + 1, nkx = wkx_hi - wkx_lo + 1;
    const int nkx0 = (nkx + 1) >> 1;
    const int h_lo = half ? wkx_lo + nkx0 : wkx_lo;
    const int h_hi = half ? wkx_hi : wkx_lo + nkx0 - 1;
    const int total = (nky > 0 && h_hi >= h_lo) ? (h_hi - h_lo + 1) * nky : 0;
    const int rounds = (total + U - 1) / U;
    const int lane_lo = max(kx_lo, h_lo), lane_hi = ok ? min(kx_hi, h_hi) : -1;
    const float* __restrict__ lane_base = patches + (1LL * y * PC + xc);
    // element offset of patch (kx, ky) relative to lane_base, stepped incrementally:
    // next ky = next patch, stride rows up; next kx = side patches on, stride*C elements left
    const long long ky_step = patch_elems - 1LL * stride * PC;
    const long long wrap_step = 1LL * side * patch_elems - sC - (nky - 1) * ky_step;
    const int wrap_k = side - (nky - 1);
    int kx = h_lo, ky = ky_lo, k = h_lo * side + ky_lo;
    long long off = (img_k0 + k - k_lo) * patch_elems - (1LL * ky * stride * PC + 1LL * kx * sC);
    // issue(): the loads of the next eight patches of the list, predicated per lane.  They are
    // software-pipelined one group ahead of the additions, so eight to sixteen 16-byte loads
    // per lane are always in flight (ptxas sinks a load next to its first use; here that use
    // is one loop iteration away).
    auto issue = [&](float (&v)[U][VEC]) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool in = kx >= lane_lo && kx <= lane_hi && static_cast<unsigned>(k - kl) < kspan;
        if constexpr (VEC == 4) {
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          if (in) t = __ldg(reinterpret_cast<const float4*>(lane_base + off));
          v[u][0] = t.x, v[u][1 % VEC] = t.y, v[u][2 % VEC] = t.z, v[u][3 % VEC] = t.w;
        } else {
          v[u][0] = in ? __ldg(lane_base + off) : 0.f;
        }
        const bool wrap = ky == ky_hi;
        ky = wrap ? ky_lo : ky + 1;
        kx += wrap ? 1 : 0;
        k += wrap ? wrap_k : 1;
        off += wrap ? wrap_step : ky_step;
      }
    };
    auto accumulate = [&](const float (&v)[U][VEC]) {
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] += static_cast<double>(v[u][e]);
    };
    if (rounds > 0) {
      float va[U][VEC], vb[U][VEC];
      issue(va);
      for (int rd = 0; rd < rounds; rd += 2) {
        if (rd + 1 < rounds) issue(vb);
        accumulate(va);
        if (rd + 1 >= rounds) break;
        if (rd + 2 < rounds) issue(va);
        accumulate(vb);
      }
    }
  }
  if (half == 1) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) part[r][e][lane] = acc[e];
  }
  __syncthreads();
  if (half == 0 && ok) {
    float res[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const double sum = acc[e] + part[r][e][lane];
      res[e] = static_cast<float>(normalize ? sum / static_cast<double>(cnt) : sum);
    }
    float* __restrict__ o = out + row * row_elems + xc;
    if constexpr (VEC == 4)
      *reinterpret_cast<float4*>(o) = make_float4(res[0], res[1 % VEC], res[2 % VEC], res[3 % VEC]);
    else
      o[0] = res[0];
  }
}

// images.quantize_mask (images.py:256-266) and the patch labels behind save_submission_csv
// (images.py:206-237: extract_patches(mask, p) -> labels_for_patches, images.py:88-99).
// One block per row of p x p cells of one mask (cells at the right / bottom edge are clipped
// like the reference's slices): a thread walks whole columns of the strip, so every warp reads
// and writes contiguous row segments; column sums meet in shared memory and one thread per cell
// adds its p columns in a fixed order.  rule 0: label = mean(v >= pixel_thr) > vote_thr
// (quantize_mask); rule 1: label = mean(v) > vote_thr (labels_for_patches).  Both means are
// taken in fp64, so the vote rule is exact (a count divided by the cell area).  The label is
// written to every pixel of the cell (quantized, optional) and / or to
// labels[n][x cell][y cell] -- x-outer, the order of extract_patches and of submission.csv.
template <typename T>
__global__ void __launch_bounds__(256)
    patch_vote_kernel(const T* __restrict__ masks, int S, int patch, int cells, int rule,
                      double pixel_thr, double vote_thr, T* __restrict__ quantized,
                      unsigned char* __restrict__ labels) {
  extern __shared__ double col_sum[];                                     // [S] then labels [cells]
  unsigned char* label_s = reinterpret_cast<unsigned char*>(col_sum + S);
  const int cy = blockIdx.y, n = blockIdx.x;
  const int y0 = cy * patch, h = min(patch, S - y0);
  const T* __restrict__ src = masks + (1LL * n * S + y0) * S;
  for (int x = threadIdx.x; x < S; x += blockDim.x) {
    double acc = 0.0;
    int yy = 0;
    for (; yy + 4 <= h; yy += 4) {  // four rows in flight per thread
      T v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = src[1LL * (yy + u) * S + x];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double d = static_cast<double>(v[u]);
        acc += rule == 0 ? (d >= pixel_thr ? 1.0 : 0.0) : d;
      }
    }
    for (; yy < h; ++yy) {
      const double d = static_cast<double>(src[1LL * yy * S + x]);
      acc += rule == 0 ? (d >= pixel_thr ? 1.0 : 0.0) : d;
    }
    col_sum[x] = acc;
  }
  __syncthreads();
  for (int cx = threadIdx.x; cx < cells; cx += blockDim.x) {
    const int x0 = cx * patch, w = min(patch, S - x0);
    double sum = 0.0;
    for (int xx = 0; xx < w; ++xx) sum += col_sum[x0 + xx];
    const unsigned char label = sum / static_cast<double>(w * h) > vote_thr ? 1 : 0;
    label_s[cx] = label;
    if (labels != nullptr) labels[(1LL * n * cells + cx) * cells + cy] = label;
  }
  if (quantized == nullptr) return;
  __syncthreads();
  T* __restrict__ dst = quantized + (1LL * n * S + y0) * S;
  for (int x = threadIdx.x; x < S; x += blockDim.x) {
    const T lv = static_cast<T>(label_s[x / patch]);
    for (int yy = 0; yy < h; ++yy) dst[1LL * yy * S + x] = lv;
  }
}

// crop_imgs(rotate_imgs(x, angle), crop) (images.py:313-373).  scipy.ndimage.rotate with
// order=0, reshape=True, mode='constant', cval=0: output pixel o maps to input coordinate
// R*o + offset (fp64); the sample is in[floor(y+0.5), floor(x+0.5)] when the unrounded
// coordinate lies in [0, H-1] on both axes, else 0.  Only the centre crop is ever materialised.
// One thread per output pixel (the fp64 coordinate is shared by the C channels); a block covers
// a 32 x 8 pixel patch so that the gathered source pixels share cache lines.
struct RotParams {
  double m00, m01, m10, m11, off0, off1;
  int out_side, crop0;
};
template <int C_T>
__global__ void __launch_bounds__(256)
    rotate_nn_crop_kernel(const float* __restrict__ in, int H, int C_rt, RotParams rp, int crop,
                          float* __restrict__ out) {
  const int C = C_T ? C_T : C_rt;
  const int n = blockIdx.z;
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if (x >= crop || y >= crop) return;
  const double oy = static_cast<double>(y + rp.crop0), ox = static_cast<double>(x + rp.crop0);
  // SciPy's NI_GeometricTransform order: cc = shift; cc += o[0]*m[i][0]; cc += o[1]*m[i][1]
  // (explicit _rn intrinsics: no FMA contraction, so ties resolve exactly as on the CPU)
  const double iy = __dadd_rn(__dadd_rn(rp.off0, __dmul_rn(oy, rp.m00)), __dmul_rn(ox, rp.m01));
  const double ix = __dadd_rn(__dadd_rn(rp.off1, __dmul_rn(oy, rp.m10)), __dmul_rn(ox, rp.m11));
  // mode='constant' is decided on the unrounded coordinate: outside [0, H-1] -> cval
  const double hi = static_cast<double>(H - 1);
  const bool inside = iy >= 0.0 && iy <= hi && ix >= 0.0 && ix <= hi;
  const long long ry = static_cast<long long>(floor(iy + 0.5));
  const long long rx = static_cast<long long>(floor(ix + 0.5));
  const float* __restrict__ s = in + ((1LL * n * H + ry) * H + rx) * C;
  float* __restrict__ d = out + ((1LL * n * crop + y) * crop + x) * C;
  if (C_T) {
    float v[C_T ? C_T : 1];
#pragma unroll
    for (int c = 0; c < C_T; ++c) v[c] = inside ? __ldg(s + c) : 0.f;
#pragma unroll
    for (int c = 0; c < C_T; ++c) d[c] = v[c];
  } else {
    for (int c = 0; c < C; ++c) d[c] = inside ? __ldg(s + c) : 0.f;
  }
}

// images.invert_image_augmentation_ensemble (images.py:399-417): undo the 6 variants and average
// (fp64 accumulation in the reference's order).  Per 32 x 32 output tile the six source tiles
// are read row-wise (coalesced, also for the transposing variants) into shared memory.
__global__ void __launch_bounds__(256)
    ensemble_invert_kernel(const float* __restrict__ masks, int N, int S, float* __restrict__ out) {
  __shared__ float t[6][32][33];
  const int n = blockIdx.z;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int i1 = min(i0 + 31, S - 1), j1 = min(j0 + 31, S - 1);
  const int tx = threadIdx.x, ty = threadIdx.y;
  // variant v -> inverse op: 0 id, 1 fliplr (=flipud+rot180), 2 flipud, 3 rot90^-1, 4 rot90^-2, 5 rot90^-3
  const int inv_ops[6] = {0, 4 | 2, 4, 3, 2, 1};
  const long long img = 1LL * S * S;
  float v[6][4];
  int si0[6], sj0[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    int sh, sw;
    d4_src_tile(inv_ops[k], S, i0, j0, i1, j1, &si0[k], &sj0[k], &sh, &sw);
    const float* __restrict__ src = masks + (1LL * k * N + n) * img + 1LL * si0[k] * S + sj0[k];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int rr = ty + 8 * r;
      v[k][r] = (rr < sh && tx < sw) ? __ldg(src + 1LL * rr * S + tx) : 0.f;
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k)
#pragma unroll
    for (int r = 0; r < 4; ++r) t[k][ty + 8 * r][tx] = v[k][r];
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty + 8 * r, j = j0 + tx;
    if (i <= i1 && j <= j1) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        int si, sj;
        d4_src(inv_ops[k], S, i, j, &si, &sj);
        acc += static_cast<double>(t[k][si - si0[k]][sj - sj0[k]]);
      }
      out[n * img + 1LL * i * S + j] = static_cast<float>(acc / 6.0);
    }
  }
}

// The same with 16-byte accesses (S a multiple of 4, 16-byte-aligned buffers): one float4 per thread
// and variant on the way in (six independent loads in flight), one float4 per thread on the way
// out; the six gathers go through per-variant affine maps computed once per block.
__global__ void __launch_bounds__(256)
    ensemble_invert_vec_kernel(const float* __restrict__ masks, int N, int S, float* __restrict__ out) {
  constexpr int T = 32, PITCH = 33;
  __shared__ float t[6][T * PITCH];
  const int n = blockIdx.z;
  const int i0 = blockIdx.y * T, j0 = blockIdx.x * T;
  const int i1 = min(i0 + T - 1, S - 1), j1 = min(j0 + T - 1, S - 1);
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int rr = tid >> 3, c4 = (tid & 7) << 2;  // row of the tile, first of four columns
  const int inv_ops[6] = {0, 4 | 2, 4, 3, 2, 1};
  const long long img = 1LL * S * S;
  float4 v[6];
  int base[6], step_r[6], step_j[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    int si0, sj0, sh, sw;
    d4_src_tile(inv_ops[k], S, i0, j0, i1, j1, &si0, &sj0, &sh, &sw);
    const float* __restrict__ src = masks + (1LL * k * N + n) * img + 1LL * si0 * S + sj0;
    v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rr < sh && c4 < sw) v[k] = __ldg(reinterpret_cast<const float4*>(src + 1LL * rr * S + c4));
    int o_i, o_j, r_i, r_j, c_i, c_j;
    d4_src(inv_ops[k], S, i0, j0, &o_i, &o_j);
    d4_src(inv_ops[k], S, i0 + 1, j0, &r_i, &r_j);
    d4_src(inv_ops[k], S, i0, j0 + 1, &c_i, &c_j);
    step_r[k] = (r_i - o_i) * PITCH + (r_j - o_j);
    step_j[k] = (c_i - o_i) * PITCH + (c_j - o_j);
    base[k] = (o_i - si0) * PITCH + (o_j - sj0);
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    float* row = &t[k][rr * PITCH + c4];
    row[0] = v[k].x, row[1] = v[k].y, row[2] = v[k].z, row[3] = v[k].w;
  }
  __syncthreads();
  const int i = i0 + rr, j = j0 + c4;
  if (i <= i1 && j <= j1) {  // (S % 4 == 0: a group of four columns is inside or outside as a whole)
    float res[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k)
        acc += static_cast<double>(t[k][base[k] + rr * step_r[k] + (c4 + e) * step_j[k]]);
      res[e] = static_cast<float>(acc / 6.0);
    }
    *reinterpret_cast<float4*>(out + n * img + 1LL * i * S + j) = make_float4(res[0], res[1], res[2], res[3]);
  }
}

// Windows of a batch of images addressed by a device job table: job j copies the win x win window
// of image jobs[j].x whose top-left pixel is (jobs[j].y, jobs[j].z) -- pixels outside the image
// read as zero -- into window jobs[j].w of `out`.  One block per (job, window row).  Serves both
// directions of the shared-window prediction (tf_aerial_images.shared_window_plan): cutting the
// enlarged input windows out of the mirror-padded images, and cutting the patch outputs out of the
// enlarged probability maps straight into their slots of the patch list.
template <int VEC>
__global__ void __launch_bounds__(128)
    copy_windows_kernel(const float* __restrict__ in, int H, int W, int C, int win,
                        const int4* __restrict__ jobs, float* __restrict__ out) {
  const int j = blockIdx.x / win, r = blockIdx.x - j * win;
  const int4 job = __ldg(jobs + j);
  const int y = job.y + r;
  const int n_el = win * C;
  float* __restrict__ dst = out + (1LL * job.w * win + r) * n_el;
  const bool row_in = y >= 0 && y < H;
  const int x_lo = max(0, -job.z), x_hi = min(win, W - job.z);  // window columns inside the image
  const int e_lo = row_in ? x_lo * C : n_el, e_hi = row_in ? max(x_hi * C, e_lo) : n_el;
  const float* __restrict__ src = in + ((1LL * job.x * H + (row_in ? y : 0)) * W + job.z) * C;
  if (VEC && e_lo == 0 && e_hi == n_el && aligned16(src)) {
    constexpr int U = 4;
    const int n4 = n_el >> 2;
    for (int g0 = threadIdx.x; g0 < n4; g0 += blockDim.x * U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int g = g0 + u * blockDim.x;
        if (g < n4) v[u] = __ldg(reinterpret_cast<const float4*>(src) + g);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int g = g0 + u * blockDim.x;
        if (g < n4) reinterpret_cast<float4*>(dst)[g] = v[u];
      }
    }
  } else {
    constexpr int U = 8;
    for (int e0 = threadIdx.x; e0 < n_el; e0 += blockDim.x * U) {
      float v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = e0 + u * blockDim.x;
        if (e < n_el) v[u] = (e >= e_lo && e < e_hi) ? __ldg(src + e) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = e0 + u * blockDim.x;
        if (e < n_el) dst[e] = v[u];
      }
    }
  }
}

// sums[n, y, x, c] /= number of sliding windows covering pixel (y, x): the analytic hit count of
// images_from_patches (images.py:154-162) for side x side patches of size P at `stride`.
__global__ void __launch_bounds__(256)
    divide_by_hits_kernel(float* __restrict__ sums, int S, int C, int side, int P, int stride,
                          long long total) {
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total;
       i += 1LL * gridDim.x * blockDim.x) {
    const long long px = i / C;
    const int x = static_cast<int>(px % S), y = static_cast<int>((px / S) % S);
    const int ky_lo = (y - P + 1 <= 0) ? 0 : (y - P + stride) / stride, ky_hi = min(y / stride, side - 1);
    const int kx_lo = (x - P + 1 <= 0) ? 0 : (x - P + stride) / stride, kx_hi = min(x / stride, side - 1);
    const int cnt = (ky_hi - ky_lo + 1) * (kx_hi - kx_lo + 1);
    sums[i] = static_cast<float>(static_cast<double>(sums[i]) / static_cast<double>(cnt));
  }
}

}  // namespace rsu

using namespace rsu;

extern "C" {

int rsu_copy_windows(const float* in, int N, int H, int W, int C, int win, long long n_jobs,
                     const int* jobs_dev, float* out, void* stream) {
  if (N < 1 || H < 1 || W < 1 || C < 1 || win < 1 || n_jobs < 0)
    return set_error(RSU_EINVAL, "copy_windows: shape");
  if (n_jobs == 0) return RSU_OK;
  if (reinterpret_cast<uintptr_t>(jobs_dev) & 15) return set_error(RSU_EALIGN, "copy_windows: job table");
  const long long rows = n_jobs * win;
  if (rows > 0x7fffffffLL || 1LL * win * C > 0x7fffffffLL)
    return set_error(RSU_EINVAL, "copy_windows: more than 2^31 window rows");
  const bool vec = (win * C) % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  const int4* jobs = reinterpret_cast<const int4*>(jobs_dev);
  if (vec)
    copy_windows_kernel<1><<<static_cast<unsigned>(rows), 128, 0, (cudaStream_t)stream>>>(in, H, W, C, win, jobs, out);
  else
    copy_windows_kernel<0><<<static_cast<unsigned>(rows), 128, 0, (cudaStream_t)stream>>>(in, H, W, C, win, jobs, out);
  return check_launch("copy_windows");
}

int rsu_divide_by_hits(float* sums, int N, int S, int C, int side, int P, int stride, void* stream) {
  if (N < 1 || S < 1 || C < 1 || side < 1 || P < 1 || stride < 1 || S != (side - 1) * stride + P)
    return set_error(RSU_EINVAL, "divide_by_hits: shape (S must be (side - 1) * stride + P)");
  const long long total = 1LL * N * S * S * C;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  divide_by_hits_kernel<<<static_cast<unsigned>(blocks), 256, 0, (cudaStream_t)stream>>>(sums, S, C, side, P,
                                                                                        stride, total);
  return check_launch("divide_by_hits");
}

int rsu_mirror_pad(const float* in, int N, int H, int W, int C, int pad, float* out, void* stream) {
  if (N < 1 || H < 1 || W < 1 || C < 1 || pad < 0) return set_error(RSU_EINVAL, "mirror_pad: shape");
  const long long rows = 1LL * N * (H + 2 * pad);
  if (rows > 0x7fffffffLL || 1LL * (W + 2 * pad) * C > 0x7fffffffLL)
    return set_error(RSU_EINVAL, "mirror_pad: more than 2^31 rows / row elements");
  const bool vec = ((W + 2 * pad) * C) % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  const dim3 grid(static_cast<unsigned>(rows));
  cudaStream_t st = (cudaStream_t)stream;
#define RSU_MP(CT, V) mirror_pad_kernel<CT, V><<<grid, 256, 0, st>>>(in, H, W, C, pad, out)
  if (vec) {
    if (C == 3) RSU_MP(3, 1);
    else if (C == 1) RSU_MP(1, 1);
    else RSU_MP(0, 1);
  } else {
    if (C == 3) RSU_MP(3, 0);
    else if (C == 1) RSU_MP(1, 0);
    else RSU_MP(0, 0);
  }
#undef RSU_MP
  return check_launch("mirror_pad");
}

int rsu_d4_transform(const void* in, void* out, int N, int S, int pixel_bytes,
                     const unsigned char* ops, int n_in, void* stream) {
  if (N < 1 || S < 1 || pixel_bytes < 1 || n_in < 0)
    return set_error(RSU_EINVAL, "d4_transform: bad shape N=%d S=%d pixel_bytes=%d", N, S, pixel_bytes);
  if (n_in == 0) n_in = N;
  if (in == out) return set_error(RSU_EINVAL, "d4_transform: in-place not supported");
  if (N > 65535) return set_error(RSU_EINVAL, "d4_transform: N > 65535");
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 block(32, 8);
  const dim3 g32((S + 31) / 32, (S + 31) / 32, N), g64((S + 63) / 64, (S + 63) / 64, N);
  const bool word4 = pixel_bytes % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 3) == 0 &&
                     (reinterpret_cast<uintptr_t>(out) & 3) == 0;
  const bool vec16 = word4 && (1LL * S * pixel_bytes) % 16 == 0 &&
                     ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (vec16 && pixel_bytes == 12) {  // RGB fp32, 16-byte accesses
    d4_transform_vec_kernel<uint32_t, 3, 32, uint4><<<g32, block, 0, st>>>(
        static_cast<const uint32_t*>(in), static_cast<uint32_t*>(out), S, ops, n_in);
  } else if (vec16 && pixel_bytes == 4) {  // fp32 masks, 16-byte accesses
    d4_transform_vec_kernel<uint32_t, 1, 64, uint4><<<g64, block, 0, st>>>(
        static_cast<const uint32_t*>(in), static_cast<uint32_t*>(out), S, ops, n_in);
  } else if (pixel_bytes == 1 && S % 4 == 0 &&
             ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 3) == 0) {
    // uint8 label masks, 4 pixels per access
    d4_transform_vec_kernel<uint8_t, 1, 64, uint32_t><<<g64, block, 0, st>>>(
        static_cast<const uint8_t*>(in), static_cast<uint8_t*>(out), S, ops, n_in);
  } else if (word4 && pixel_bytes == 12) {  // RGB fp32
    d4_transform_kernel<uint32_t, 3, 32><<<g32, block, 0, st>>>(
        static_cast<const uint32_t*>(in), static_cast<uint32_t*>(out), S, ops, n_in);
  } else if (word4 && pixel_bytes == 4) {  // fp32 masks
    d4_transform_kernel<uint32_t, 1, 64><<<g64, block, 0, st>>>(
        static_cast<const uint32_t*>(in), static_cast<uint32_t*>(out), S, ops, n_in);
  } else if (pixel_bytes == 1) {  // uint8 label masks
    d4_transform_kernel<uint8_t, 1, 64><<<g64, block, 0, st>>>(
        static_cast<const uint8_t*>(in), static_cast<uint8_t*>(out), S, ops, n_in);
  } else if (word4) {
    const int words = pixel_bytes / 4;
    const size_t smem = 32 * (32 * words + 1) * sizeof(uint32_t);
    if (smem > 48 * 1024) return set_error(RSU_EINVAL, "d4_transform: pixel too large");
    d4_transform_generic_kernel<uint32_t><<<g32, block, smem, st>>>(
        static_cast<const uint32_t*>(in), static_cast<uint32_t*>(out), S, words, ops, n_in);
  } else {  // byte-granular pixels
    const size_t smem = 32 * (32 * pixel_bytes + 1);
    if (smem > 48 * 1024) return set_error(RSU_EINVAL, "d4_transform: pixel too large");
    d4_transform_generic_kernel<uint8_t><<<g32, block, smem, st>>>(
        static_cast<const uint8_t*>(in), static_cast<uint8_t*>(out), S, pixel_bytes, ops, n_in);
  }
  return check_launch("d4_transform");
}

int rsu_extract_patches(const float* in, int N, int H, int W, int C, int patch, int stride,
                        long long k_begin, long long k_count, float* out, void* stream) {
  if (H != W) return set_error(RSU_EINVAL, "extract_patches: Assume square images");
  if (stride < 1 || patch > H || (H - patch) % stride != 0)
    return set_error(RSU_EINVAL, "extract_patches: Stride sliding should cover the whole image");
  const int side = (H - patch) / stride + 1;
  const long long all = 1LL * N * side * side;
  if (k_count < 0) k_count = all - k_begin;
  if (k_begin < 0 || k_begin + k_count > all || k_count < 1)
    return set_error(RSU_EINVAL, "extract_patches: patch range [%lld, +%lld) outside %lld", k_begin,
                     k_count, all);
  const long long rows = k_count * patch;
  if (rows > 0x7fffffffLL) return set_error(RSU_EINVAL, "extract_patches: more than 2^31 patch rows");
  const bool vec = (patch * C) % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  const dim3 grid(static_cast<unsigned>(rows));
  if (vec)
    extract_patches_kernel<1><<<grid, 128, 0, (cudaStream_t)stream>>>(in, H, W, C, patch, stride, side,
                                                                      k_begin, out);
  else
    extract_patches_kernel<0><<<grid, 128, 0, (cudaStream_t)stream>>>(in, H, W, C, patch, stride, side,
                                                                      k_begin, out);
  return check_launch("extract_patches");
}

int rsu_overlap_average(const float* patches, int N, int side, int P, int C, int stride,
                        long long k_begin, long long k_count, int normalize, float* out,
                        void* stream) {
  if (N < 1 || side < 1 || P < 1 || C < 1 || stride < 1)
    return set_error(RSU_EINVAL, "overlap_average: shape");
  const long long all = 1LL * N * side * side;
  if (k_count < 0) k_count = all - k_begin;
  if (k_begin < 0 || k_begin + k_count > all)
    return set_error(RSU_EINVAL, "overlap_average: patch range outside [0, %lld)", all);
  const int S = (side - 1) * stride + P;
  const long long rows = 1LL * N * S;
  if (1LL * P * P * C > 0x7fffffffLL || 1LL * S * C > 0x7fffffffLL || side > 46340)
    return set_error(RSU_EINVAL, "overlap_average: patch, row or patch grid too large");
  const bool vec = (P * C) % 4 == 0 && (stride * C) % 4 == 0 &&
                   ((reinterpret_cast<uintptr_t>(patches) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  const int lane_elems = vec ? 4 : 1;
  const int n_chunks = (S * C + 32 * lane_elems - 1) / (32 * lane_elems);
  const long long blocks = ((rows + 3) / 4) * n_chunks;
  if (blocks > 0x7fffffffLL) return set_error(RSU_EINVAL, "overlap_average: grid too large");
  const cudaStream_t st = (cudaStream_t)stream;
  if (vec)
    overlap_average_kernel<4><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
        patches, side, P, C, stride, S, rows, n_chunks, k_begin, k_begin + k_count, normalize, out);
  else
    overlap_average_kernel<1><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
        patches, side, P, C, stride, S, rows, n_chunks, k_begin, k_begin + k_count, normalize, out);
  return check_launch("overlap_average");
}

int rsu_rotate_nn_crop(const float* in, int N, int H, int C, const double* matrix_host,
                       const double* offset_host, int crop0, int crop, float* out, void* stream) {
  if (N < 1 || H < 1 || C < 1 || crop < 1 || crop0 < 0)
    return set_error(RSU_EINVAL, "rotate_nn_crop: shape");
  if (N > 65535) return set_error(RSU_EINVAL, "rotate_nn_crop: N > 65535");
  RotParams rp;
  rp.m00 = matrix_host[0];
  rp.m01 = matrix_host[1];
  rp.m10 = matrix_host[2];
  rp.m11 = matrix_host[3];
  rp.off0 = offset_host[0];
  rp.off1 = offset_host[1];
  rp.out_side = 0;
  rp.crop0 = crop0;
  const dim3 grid((crop + 31) / 32, (crop + 7) / 8, N), block(32, 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 3) rotate_nn_crop_kernel<3><<<grid, block, 0, st>>>(in, H, C, rp, crop, out);
  else if (C == 1) rotate_nn_crop_kernel<1><<<grid, block, 0, st>>>(in, H, C, rp, crop, out);
  else if (C == 6) rotate_nn_crop_kernel<6><<<grid, block, 0, st>>>(in, H, C, rp, crop, out);
  else if (C == 2) rotate_nn_crop_kernel<2><<<grid, block, 0, st>>>(in, H, C, rp, crop, out);
  else rotate_nn_crop_kernel<0><<<grid, block, 0, st>>>(in, H, C, rp, crop, out);
  return check_launch("rotate_nn_crop");
}

int rsu_ensemble_invert(const float* masks, int N, int S, float* out, void* stream) {
  if (N < 1 || S < 1) return set_error(RSU_EINVAL, "ensemble_invert: shape");
  if (N > 65535) return set_error(RSU_EINVAL, "ensemble_invert: N > 65535");
  const dim3 grid((S + 31) / 32, (S + 31) / 32, N), block(32, 8);
  if (S % 4 == 0 && ((reinterpret_cast<uintptr_t>(masks) | reinterpret_cast<uintptr_t>(out)) & 15) == 0)
    ensemble_invert_vec_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(masks, N, S, out);
  else
    ensemble_invert_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(masks, N, S, out);
  return check_launch("ensemble_invert");
}

int rsu_patch_vote(const void* masks, int elem_bytes, int N, int S, int patch, int rule,
                   double pixel_threshold, double vote_threshold, void* quantized,
                   unsigned char* labels, void* stream) {
  if (N < 1 || S < 1 || patch < 1 || (rule != 0 && rule != 1))
    return set_error(RSU_EINVAL, "patch_vote: shape / rule");
  if (elem_bytes != 4 && elem_bytes != 8) return set_error(RSU_EINVAL, "patch_vote: fp32 or fp64 masks");
  const int g = (S + patch - 1) / patch;
  if (g > 65535) return set_error(RSU_EINVAL, "patch_vote: more than 65535 cells per side");
  const size_t smem = sizeof(double) * S + ((g + 7) / 8) * 8;
  if (smem > 48 * 1024) return set_error(RSU_EINVAL, "patch_vote: mask side %d too large", S);
  const dim3 grid(N, g);  // (masks in grid.x: any number of them)
  const cudaStream_t st = (cudaStream_t)stream;
  if (elem_bytes == 4)
    patch_vote_kernel<float><<<grid, 256, smem, st>>>(static_cast<const float*>(masks), S, patch, g, rule,
                                                      pixel_threshold, vote_threshold,
                                                      static_cast<float*>(quantized), labels);
  else
    patch_vote_kernel<double><<<grid, 256, smem, st>>>(static_cast<const double*>(masks), S, patch, g, rule,
                                                       pixel_threshold, vote_threshold,
                                                       static_cast<double*>(quantized), labels);
  return check_launch("patch_vote");
}

}  // extern "C"
