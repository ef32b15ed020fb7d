// Geometry kernels replacing the NumPy / SciPy helpers of src/images.py (fp32 images, NHWC).
// All of them are pure data movement (bit-exact against the reference) except the two
// averaging kernels, which accumulate in fp64 in the reference's summation order.
#include "host_common.h"

namespace rsu {

static int geo_grid(long long items, int threads) {
  long long blocks = (items + threads - 1) / threads;
  const long long cap = 1LL * num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

// np.pad(..., "symmetric") index: reflect with the edge sample repeated, period 2*L.
__device__ __forceinline__ int sym_index(int i, int L) {
  int m = i % (2 * L);
  if (m < 0) m += 2 * L;
  return m < L ? m : 2 * L - 1 - m;
}

// images.mirror_border (images.py:269-281)
__global__ void mirror_pad_kernel(const float* __restrict__ in, int N, int H, int W, int C, int pad,
                                  float* __restrict__ out) {
  const int Ho = H + 2 * pad, Wo = W + 2 * pad;
  const long long row_elems = 1LL * Wo * C;
  const long long total = 1LL * N * Ho * row_elems;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total;
       i += 1LL * gridDim.x * blockDim.x) {
    const int xc = static_cast<int>(i % row_elems);
    const int x = xc / C, c = xc - x * C;
    const int y = static_cast<int>((i / row_elems) % Ho);
    const int n = static_cast<int>(i / (row_elems * Ho));
    const int sy = sym_index(y - pad, H), sx = sym_index(x - pad, W);
    out[i] = __ldg(in + ((1LL * n * H + sy) * W + sx) * C + c);
  }
}

// Dihedral-group transform of square images, one op per image:
//   out = rot90(flipud(x) if (op & 4) else x, k = op & 3)     (counter-clockwise, like np.rot90)
// A 32x32-pixel destination tile maps onto a 32x32 source tile; the source tile is read
// row-wise (coalesced) into shared memory and written out row-wise in destination order.
// words = 4-byte words per pixel.
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

template <typename WordT>
__global__ void d4_transform_kernel(const WordT* __restrict__ in, WordT* __restrict__ out, int S,
                                    int words, const unsigned char* __restrict__ ops) {
  extern __shared__ uint8_t tile_raw[];  // [32][32*words + 1] words
  WordT* tile = reinterpret_cast<WordT*>(tile_raw);
  const int n = blockIdx.z;
  const int op = ops[n];
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  // source tile origin = min over the tile corners of the mapped coordinates
  int ci[2], cj[2];
  d4_src(op, S, i0, j0, &ci[0], &cj[0]);
  const int i1 = min(i0 + 31, S - 1), j1 = min(j0 + 31, S - 1);
  d4_src(op, S, i1, j1, &ci[1], &cj[1]);
  const int si0 = min(ci[0], ci[1]), sj0 = min(cj[0], cj[1]);
  const int sh = abs(ci[0] - ci[1]) + 1, sw = abs(cj[0] - cj[1]) + 1;
  const int pitch = 32 * words + 1;
  const WordT* src = in + 1LL * n * S * S * words;
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
// row (k % side)*stride  (x is the OUTER loop in the reference).
__global__ void extract_patches_kernel(const float* __restrict__ in, int N, int H, int W, int C,
                                       int P, int stride, int side, long long k_begin,
                                       long long k_count, float* __restrict__ out) {
  const long long row_elems = 1LL * P * C;
  const long long total = k_count * P * row_elems;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total;
       i += 1LL * gridDim.x * blockDim.x) {
    const int xc = static_cast<int>(i % row_elems);
    const int py = static_cast<int>((i / row_elems) % P);
    const long long k_all = k_begin + i / (row_elems * P);
    const int k = static_cast<int>(k_all % (side * side));
    const int n = static_cast<int>(k_all / (side * side));
    const int x0 = (k / side) * stride, y0 = (k % side) * stride;
    out[i] = __ldg(in + ((1LL * n * H + y0 + py) * W + x0) * C + xc);
  }
}

// images.images_from_patches (images.py:131-164) in gather form: every output pixel sums the
// patches covering it in the reference's order (x outer, y inner) in fp64 and divides by the
// hit count -- deterministic, no atomics.
// k_lo/k_hi restrict the sum to patches whose global index (n*side*side + k) lies in
// [k_lo, k_hi) -- a rank of a sharded prediction holds only that slice (patches points at patch
// k_lo) and emits partial sums (normalize = 0) that are added and divided after the gather.
__global__ void overlap_average_kernel(const float* __restrict__ patches, int N, int side, int P,
                                       int C, int stride, int S, long long k_lo, long long k_hi,
                                       int normalize, float* __restrict__ out) {
  const long long row_elems = 1LL * S * C;
  const long long total = 1LL * N * S * row_elems;
  const long long patch_elems = 1LL * P * P * C;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total;
       i += 1LL * gridDim.x * blockDim.x) {
    const int xc = static_cast<int>(i % row_elems);
    const int x = xc / C, c = xc - x * C;
    const int y = static_cast<int>((i / row_elems) % S);
    const int n = static_cast<int>(i / (row_elems * S));
    int kx_lo = (x - P + stride) / stride;  // ceil((x-P+1)/stride)
    if (x - P + 1 <= 0) kx_lo = 0;
    int ky_lo = (y - P + stride) / stride;
    if (y - P + 1 <= 0) ky_lo = 0;
    const int kx_hi = min(x / stride, side - 1), ky_hi = min(y / stride, side - 1);
    double acc = 0.0;
    const long long img_k0 = 1LL * n * side * side;
    for (int kx = kx_lo; kx <= kx_hi; ++kx)
      for (int ky = ky_lo; ky <= ky_hi; ++ky) {
        const long long k = img_k0 + kx * side + ky;
        if (k < k_lo || k >= k_hi) continue;
        acc += static_cast<double>(__ldg(patches + (k - k_lo) * patch_elems +
                                         (1LL * (y - ky * stride) * P + (x - kx * stride)) * C + c));
      }
    const int cnt = (kx_hi - kx_lo + 1) * (ky_hi - ky_lo + 1);
    out[i] = static_cast<float>(normalize ? acc / static_cast<double>(cnt) : acc);
  }
}

// crop_imgs(rotate_imgs(x, angle), crop) (images.py:313-373).  scipy.ndimage.rotate with
// order=0, reshape=True, mode='constant', cval=0: output pixel o maps to input coordinate
// R*o + offset (fp64); the sample is in[floor(y+0.5), floor(x+0.5)] when the unrounded
// coordinate lies in [0, H-1] on both axes, else 0.  Only the centre crop is ever materialised.
struct RotParams {
  double m00, m01, m10, m11, off0, off1;
  int out_side, crop0;
};
__global__ void rotate_nn_crop_kernel(const float* __restrict__ in, int N, int H, int C, RotParams rp,
                                      int crop, float* __restrict__ out) {
  const long long row_elems = 1LL * crop * C;
  const long long total = 1LL * N * crop * row_elems;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total;
       i += 1LL * gridDim.x * blockDim.x) {
    const int xc = static_cast<int>(i % row_elems);
    const int x = xc / C, c = xc - x * C;
    const int y = static_cast<int>((i / row_elems) % crop);
    const int n = static_cast<int>(i / (row_elems * crop));
    const double oy = static_cast<double>(y + rp.crop0), ox = static_cast<double>(x + rp.crop0);
    // SciPy's NI_GeometricTransform order: cc = shift; cc += o[0]*m[i][0]; cc += o[1]*m[i][1]
    // (explicit _rn intrinsics: no FMA contraction, so ties resolve exactly as on the CPU)
    const double iy = __dadd_rn(__dadd_rn(rp.off0, __dmul_rn(oy, rp.m00)), __dmul_rn(ox, rp.m01));
    const double ix = __dadd_rn(__dadd_rn(rp.off1, __dmul_rn(oy, rp.m10)), __dmul_rn(ox, rp.m11));
    const long long ry = static_cast<long long>(floor(iy + 0.5));
    const long long rx = static_cast<long long>(floor(ix + 0.5));
    float v = 0.f;
    // mode='constant' is decided on the unrounded coordinate: outside [0, H-1] -> cval
    const double hi = static_cast<double>(H - 1);
    if (iy >= 0.0 && iy <= hi && ix >= 0.0 && ix <= hi)
      v = __ldg(in + ((1LL * n * H + ry) * H + rx) * C + c);
    out[i] = v;
  }
}

// images.invert_image_augmentation_ensemble (images.py:399-417): undo the 6 variants and average
// (fp64 accumulation in the reference's order).
__global__ void ensemble_invert_kernel(const float* __restrict__ masks, int N, int S,
                                       float* __restrict__ out) {
  const long long img = 1LL * S * S;
  const long long total = 1LL * N * img;
  for (long long idx = blockIdx.x * 1LL * blockDim.x + threadIdx.x; idx < total;
       idx += 1LL * gridDim.x * blockDim.x) {
    const int j = static_cast<int>(idx % S);
    const int i = static_cast<int>((idx / S) % S);
    const int n = static_cast<int>(idx / img);
    // variant v -> inverse op: 0 id, 1 fliplr (=flipud+rot180), 2 flipud, 3 rot90^-1, 4 rot90^-2, 5 rot90^-3
    const int inv_ops[6] = {0, 4 | 2, 4, 3, 2, 1};
    double acc = 0.0;
#pragma unroll
    for (int v = 0; v < 6; ++v) {
      int si, sj;
      d4_src(inv_ops[v], S, i, j, &si, &sj);
      acc += static_cast<double>(__ldg(masks + (1LL * v * N + n) * img + 1LL * si * S + sj));
    }
    out[idx] = static_cast<float>(acc / 6.0);
  }
}

}  // namespace rsu

using namespace rsu;

extern "C" {

int rsu_mirror_pad(const float* in, int N, int H, int W, int C, int pad, float* out, void* stream) {
  if (N < 1 || H < 1 || W < 1 || C < 1 || pad < 0) return set_error(RSU_EINVAL, "mirror_pad: shape");
  const long long total = 1LL * N * (H + 2 * pad) * (W + 2 * pad) * C;
  mirror_pad_kernel<<<geo_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(in, N, H, W, C, pad, out);
  return check_launch("mirror_pad");
}

int rsu_d4_transform(const void* in, void* out, int N, int S, int pixel_bytes,
                     const unsigned char* ops, void* stream) {
  if (N < 1 || S < 1 || pixel_bytes < 1)
    return set_error(RSU_EINVAL, "d4_transform: bad shape N=%d S=%d pixel_bytes=%d", N, S, pixel_bytes);
  if (in == out) return set_error(RSU_EINVAL, "d4_transform: in-place not supported");
  if (N > 65535) return set_error(RSU_EINVAL, "d4_transform: N > 65535");
  dim3 grid((S + 31) / 32, (S + 31) / 32, N), block(32, 8);
  const bool word4 = pixel_bytes % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 3) == 0 &&
                     (reinterpret_cast<uintptr_t>(out) & 3) == 0;
  if (word4) {
    const int words = pixel_bytes / 4;
    const size_t smem = 32 * (32 * words + 1) * sizeof(uint32_t);
    if (smem > 48 * 1024) return set_error(RSU_EINVAL, "d4_transform: pixel too large");
    d4_transform_kernel<uint32_t><<<grid, block, smem, (cudaStream_t)stream>>>(
        static_cast<const uint32_t*>(in), static_cast<uint32_t*>(out), S, words, ops);
  } else {  // byte-granular pixels (uint8 label masks)
    const size_t smem = 32 * (32 * pixel_bytes + 1);
    if (smem > 48 * 1024) return set_error(RSU_EINVAL, "d4_transform: pixel too large");
    d4_transform_kernel<uint8_t><<<grid, block, smem, (cudaStream_t)stream>>>(
        static_cast<const uint8_t*>(in), static_cast<uint8_t*>(out), S, pixel_bytes, ops);
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
  const long long total = k_count * patch * patch * C;
  extract_patches_kernel<<<geo_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(
      in, N, H, W, C, patch, stride, side, k_begin, k_count, out);
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
  const long long total = 1LL * N * S * S * C;
  overlap_average_kernel<<<geo_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(
      patches, N, side, P, C, stride, S, k_begin, k_begin + k_count, normalize, out);
  return check_launch("overlap_average");
}

int rsu_rotate_nn_crop(const float* in, int N, int H, int C, const double* matrix_host,
                       const double* offset_host, int crop0, int crop, float* out, void* stream) {
  if (N < 1 || H < 1 || C < 1 || crop < 1 || crop0 < 0)
    return set_error(RSU_EINVAL, "rotate_nn_crop: shape");
  RotParams rp;
  rp.m00 = matrix_host[0];
  rp.m01 = matrix_host[1];
  rp.m10 = matrix_host[2];
  rp.m11 = matrix_host[3];
  rp.off0 = offset_host[0];
  rp.off1 = offset_host[1];
  rp.out_side = 0;
  rp.crop0 = crop0;
  const long long total = 1LL * N * crop * crop * C;
  rotate_nn_crop_kernel<<<geo_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(in, N, H, C, rp, crop,
                                                                               out);
  return check_launch("rotate_nn_crop");
}

int rsu_ensemble_invert(const float* masks, int N, int S, float* out, void* stream) {
  if (N < 1 || S < 1) return set_error(RSU_EINVAL, "ensemble_invert: shape");
  ensemble_invert_kernel<<<geo_grid(1LL * N * S * S, 256), 256, 0, (cudaStream_t)stream>>>(masks, N, S,
                                                                                            out);
  return check_launch("ensemble_invert");
}

}  // extern "C"
