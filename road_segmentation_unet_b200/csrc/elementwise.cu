// HBM-bound fused elementwise / reduction kernels of the U-Net hot path (sm_100a).
// All activations are NHWC bf16; every thread moves 16-byte vectors (8 bf16) so that a warp
// covers 512 contiguous bytes.  Reference ops are cited per kernel.
#include "host_common.h"
#include "ptx.cuh"

namespace rsu {

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    f[2 * e] = bf16_lo(w[e]);
    f[2 * e + 1] = bf16_hi(w[e]);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]);
  o.w = pack_bf16x2(f[6], f[7]);
  return o;
}

// Counter-based uniform in [0,1): splitmix64 of (seed, index); 24 random bits.
__host__ __device__ __forceinline__ float uniform01(unsigned long long seed,
                                                    unsigned long long idx) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z = z ^ (z >> 31);
  return static_cast<float>(z >> 40) * (1.0f / 16777216.0f);
}
// tf.nn.dropout keeps an element iff floor(keep + U) == 1
__host__ __device__ __forceinline__ float keep_scale(unsigned long long seed,
                                                     unsigned long long idx, float keep) {
  return (keep + uniform01(seed, idx) >= 1.0f) ? 1.0f / keep : 0.0f;
}

// ------------------------------------------------------------------ weight packing
__global__ void pack_transpose_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                      int T, int R, int C, int ld) {
  __shared__ float tile[32][33];
  const int t = blockIdx.z;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < C) tile[i][threadIdx.x] = in[(static_cast<long long>(t) * R + r) * C + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C)
      out[static_cast<long long>(c) * ld + t * R + r] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}

struct Perm {
  int v[16];
};
__global__ void pack_permute_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                    int T, int R, int C, Perm perm) {
  const long long total = 1LL * T * R * C;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total;
       i += 1LL * gridDim.x * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long long tr = i / C;
    const int r = static_cast<int>(tr % R);
    const int t = static_cast<int>(tr / R);
    out[(static_cast<long long>(r) * T + perm.v[t]) * C + c] = __float2bfloat16(in[i]);
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                 long long n) {
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < n;
       i += 1LL * gridDim.x * blockDim.x)
    out[i] = __float2bfloat16(in[i]);
}

// All weight repacks of one optimizer step in ONE launch: every block looks its job up in a
// device table (block ranges are prefix sums), then runs the same tile code as the kernels above.
struct PackJobDev {
  const float* in;
  __nv_bfloat16* out;
  int kind;  // 0 = transpose [T][R][C] -> [C][ld], 1 = permute [T][R][C] -> [R][T][C], 2 = cast
  int T, R, C, ld;
  int block_begin, block_count;
};
__global__ void pack_batch_kernel(const PackJobDev* __restrict__ jobs, int n_jobs) {
  __shared__ float tile[64][33];
  __shared__ int s_job;
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    int lo = 0, hi = n_jobs - 1;
    const int b = blockIdx.x;
    while (lo < hi) {  // last job whose block_begin <= b
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].block_begin <= b) lo = mid;
      else hi = mid - 1;
    }
    s_job = lo;
  }
  __syncthreads();
  const PackJobDev j = jobs[s_job];
  const int lb = blockIdx.x - j.block_begin;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (j.kind == 0) {
    // [T][R][C] fp32 -> [C][ld] bf16 (row t*R + r): tiles of 64 r x 32 c; eight coalesced loads
    // per thread, then every warp writes 64 consecutive bf16 (128 bytes) of one output row
    const int tiles_c = (j.C + 31) / 32, tiles_r = (j.R + 63) / 64;
    const int t = lb / (tiles_c * tiles_r);
    const int rem = lb % (tiles_c * tiles_r);
    const int c0 = (rem % tiles_c) * 32, r0 = (rem / tiles_c) * 64;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = r0 + threadIdx.y + 8 * i, c = c0 + threadIdx.x;
      v[i] = (r < j.R && c < j.C) ? __ldg(j.in + (static_cast<long long>(t) * j.R + r) * j.C + c) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) tile[threadIdx.y + 8 * i][threadIdx.x] = v[i];
    __syncthreads();
    const bool pair_ok = ((j.ld | (t * j.R + r0)) & 1) == 0 && (reinterpret_cast<uintptr_t>(j.out) & 3) == 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + threadIdx.y + 8 * i, r = r0 + 2 * threadIdx.x;
      if (c < j.C) {
        __nv_bfloat16* o = j.out + static_cast<long long>(c) * j.ld + t * j.R + r;
        const float a0 = tile[2 * threadIdx.x][threadIdx.y + 8 * i];
        const float a1 = tile[2 * threadIdx.x + 1][threadIdx.y + 8 * i];
        if (pair_ok && r + 1 < j.R) {
          *reinterpret_cast<__nv_bfloat162*>(o) = __floats2bfloat162_rn(a0, a1);
        } else {
          if (r < j.R) o[0] = __float2bfloat16(a0);
          if (r + 1 < j.R) o[1] = __float2bfloat16(a1);
        }
      }
    }
  } else {
    const long long total = 1LL * j.T * j.R * j.C;
    const bool vec = (j.C % 8 == 0 || j.kind == 2) && total % 8 == 0 &&
                     (reinterpret_cast<uintptr_t>(j.in) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(j.out) & 15) == 0;
    if (vec) {
      // eight consecutive elements (same t, r; consecutive c) per thread: 2 x 16 B in, 16 B out
      const long long total8 = total >> 3;
      const int c8n = j.C >> 3;
      for (long long i8 = lb * 256LL + tid; i8 < total8; i8 += 256LL * j.block_count) {
        const float4 f0 = __ldg(reinterpret_cast<const float4*>(j.in) + 2 * i8);
        const float4 f1 = __ldg(reinterpret_cast<const float4*>(j.in) + 2 * i8 + 1);
        uint4 o;
        o.x = pack_bf16x2(f0.x, f0.y);
        o.y = pack_bf16x2(f0.z, f0.w);
        o.z = pack_bf16x2(f1.x, f1.y);
        o.w = pack_bf16x2(f1.z, f1.w);
        long long dst8 = i8;
        if (j.kind != 2) {
          const int c8 = static_cast<int>(i8 % c8n);
          const long long tr = i8 / c8n;
          const int r = static_cast<int>(tr % j.R);
          const int t = static_cast<int>(tr / j.R);
          dst8 = (static_cast<long long>(r) * j.T + t) * c8n + c8;
        }
        reinterpret_cast<uint4*>(j.out)[dst8] = o;
      }
    } else {
      for (long long i = lb * 256LL + tid; i < total; i += 256LL * j.block_count) {
        if (j.kind == 2) {
          j.out[i] = __float2bfloat16(j.in[i]);
        } else {
          const int c = static_cast<int>(i % j.C);
          const long long tr = i / j.C;
          const int r = static_cast<int>(tr % j.R);
          const int t = static_cast<int>(tr / j.R);
          j.out[(static_cast<long long>(r) * j.T + t) * j.C + c] = __float2bfloat16(j.in[i]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ first layer (Cin = 3)
// color_space_adjust (unet.py:22-23) + dropout (unet.py:29-30) + 3x3 im2col into 64 channels.
struct ColorW {
  float w[3][3];
  float b[3];
};
__device__ __forceinline__ void color_adjust(const float* __restrict__ px, const ColorW& cw,
                                             float (&o)[3]) {
  const float a0 = px[0] - 0.5f, a1 = px[1] - 0.5f, a2 = px[2] - 0.5f;
#pragma unroll
  for (int m = 0; m < 3; ++m) o[m] = a0 * cw.w[0][m] + a1 * cw.w[1][m] + a2 * cw.w[2][m] + cw.b[m];
}

__global__ void color_im2col_kernel(const float* __restrict__ img, int N, int S,
                                    const float* __restrict__ w1, const float* __restrict__ b1,
                                    int d, int oy, int ox, int Ho, int Wo,
                                    __nv_bfloat16* __restrict__ out, float keep,
                                    unsigned long long seed) {
  ColorW cw;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) cw.w[i][j] = __ldg(w1 + i * 3 + j);
    cw.b[i] = __ldg(b1 + i);
  }
  const long long total = 1LL * N * Ho * Wo;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total;
       i += 1LL * gridDim.x * blockDim.x) {
    const int x = static_cast<int>(i % Wo);
    const int y = static_cast<int>((i / Wo) % Ho);
    const int n = static_cast<int>(i / (1LL * Wo * Ho));
    float col[32];
#pragma unroll
    for (int j = 28; j < 32; ++j) col[j] = 0.f;
    // constant-one column: the weight-gradient GEMM im2col^T dZ then carries BiasAddGrad in row 27
    // (the packed forward weights are zero there, so the convolution itself is unaffected)
    col[27] = 1.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = y + oy + (t / 3) * d, xx = x + ox + (t % 3) * d;
      const long long pix = (1LL * n * S + yy) * S + xx;
      float o[3];
      color_adjust(img + pix * 3, cw, o);
      if (keep < 1.0f) {
#pragma unroll
        for (int m = 0; m < 3; ++m) o[m] *= keep_scale(seed, pix * 3 + m, keep);
      }
      col[t * 3 + 0] = o[0];
      col[t * 3 + 1] = o[1];
      col[t * 3 + 2] = o[2];
    }
    uint4* op = reinterpret_cast<uint4*>(out + i * 64);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 v;
      v.x = pack_bf16x2(col[q * 8 + 0], col[q * 8 + 1]);
      v.y = pack_bf16x2(col[q * 8 + 2], col[q * 8 + 3]);
      v.z = pack_bf16x2(col[q * 8 + 4], col[q * 8 + 5]);
      v.w = pack_bf16x2(col[q * 8 + 6], col[q * 8 + 7]);
      op[q] = v;
    }
    const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int q = 4; q < 8; ++q) op[q] = z;
  }
}

// d(color_space_adjust): gather d(net0) from d(im2col), then reduce dW1 = (img-0.5)^T d(net0).
__global__ void color_im2col_bwd_kernel(const float* __restrict__ img, int N, int S,
                                        const __nv_bfloat16* __restrict__ dcol, int d, int oy,
                                        int ox, int Ho, int Wo, float* __restrict__ dw1,
                                        float* __restrict__ db1, float keep,
                                        unsigned long long seed) {
  const int Hr = Ho + 2 * d, Wr = Wo + 2 * d;  // touched region of net0
  const long long total = 1LL * N * Hr * Wr;
  float acc[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) acc[j] = 0.f;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total;
       i += 1LL * gridDim.x * blockDim.x) {
    const int rx = static_cast<int>(i % Wr);
    const int ry = static_cast<int>((i / Wr) % Hr);
    const int n = static_cast<int>(i / (1LL * Wr * Hr));
    float g[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int y = ry - (t / 3) * d, x = rx - (t % 3) * d;
      if (y >= 0 && y < Ho && x >= 0 && x < Wo) {
        const __nv_bfloat16* p = dcol + ((1LL * n * Ho + y) * Wo + x) * 64 + t * 3;
        g[0] += __bfloat162float(p[0]);
        g[1] += __bfloat162float(p[1]);
        g[2] += __bfloat162float(p[2]);
      }
    }
    const long long pix = (1LL * n * S + ry + oy) * S + rx + ox;
    if (keep < 1.0f) {
#pragma unroll
      for (int m = 0; m < 3; ++m) g[m] *= keep_scale(seed, pix * 3 + m, keep);
    }
    const float* px = img + pix * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float a = px[c] - 0.5f;
#pragma unroll
      for (int m = 0; m < 3; ++m) acc[c * 3 + m] += a * g[m];
    }
#pragma unroll
    for (int m = 0; m < 3; ++m) acc[9 + m] += g[m];
  }
  __shared__ float red[12];
  if (threadIdx.x < 12) red[threadIdx.x] = 0.f;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    float v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&red[j], v);
  }
  __syncthreads();
  if (threadIdx.x < 9) atomicAdd(dw1 + threadIdx.x, red[threadIdx.x]);
  else if (threadIdx.x < 12) atomicAdd(db1 + threadIdx.x - 9, red[threadIdx.x]);
}

// Folded first layer (Cin = 3, no dropout): color_space_adjust (unet.py:22-23) is linear and the
// convolutions that follow are VALID, so
//   conv(W, (x - 0.5) W1 + b1) + b  ==  conv(W', x - 0.5) + b'
//   W'[tap, ci, co] = sum_c W1[ci, c] W[tap, c, co]      b'[co] = b[co] + sum_{tap, c} b1[c] W[tap, c, co]
// The im2col buffer then holds the raw centred image and the backward pass needs neither the data
// gradient of the convolution nor a pass over d(im2col): with Gx = im2col(x - 0.5)^T dZ (the
// weight-gradient GEMM that is computed anyway) and db = column sums of dZ,
//   dW[tap, c, co]  = sum_ci W1[ci, c] Gx[tap, ci, co] + b1[c] db[co]
//   dW1[ci, c]      = sum_{tap, co} W[tap, c, co] Gx[tap, ci, co]
//   db1[c]          = sum_{tap, co} W[tap, c, co] db[co]
// One thread per output channel.
__global__ void first_layer_fold_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                        const float* __restrict__ w1, const float* __restrict__ b1,
                                        int cout, __nv_bfloat16* __restrict__ wp,
                                        float* __restrict__ bias_eff) {
  const int co = blockIdx.x * blockDim.x + threadIdx.x;
  if (co >= cout) return;
  float m1[9], c1[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) m1[i] = __ldg(w1 + i);
#pragma unroll
  for (int i = 0; i < 3; ++i) c1[i] = __ldg(b1 + i);
  float be = __ldg(b + co);
  __nv_bfloat16* row = wp + static_cast<long long>(co) * 64;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    float wt[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) wt[c] = __ldg(w + (t * 3 + c) * cout + co);
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
      row[t * 3 + ci] =
          __float2bfloat16(m1[ci * 3 + 0] * wt[0] + m1[ci * 3 + 1] * wt[1] + m1[ci * 3 + 2] * wt[2]);
    be += c1[0] * wt[0] + c1[1] * wt[1] + c1[2] * wt[2];
  }
  for (int k = 27; k < 64; ++k) row[k] = __float2bfloat16(0.f);
  bias_eff[co] = be;
}

__global__ void first_layer_grads_kernel(const float* __restrict__ gx, int ldg,
                                         const float* __restrict__ w, const float* __restrict__ w1,
                                         const float* __restrict__ b1, int cout,
                                         float* __restrict__ dw, float* __restrict__ dbias,
                                         float* __restrict__ dw1, float* __restrict__ db1) {
  __shared__ float red[12];
  if (threadIdx.x < 12) red[threadIdx.x] = 0.f;
  __syncthreads();
  float acc[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) acc[j] = 0.f;
  float m1[9], c1[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) m1[i] = __ldg(w1 + i);
#pragma unroll
  for (int i = 0; i < 3; ++i) c1[i] = __ldg(b1 + i);
  for (int co = threadIdx.x; co < cout; co += blockDim.x) {
    const float dbv = __ldg(gx + 27LL * ldg + co);
    dbias[co] += dbv;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float g[3];
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) g[ci] = __ldg(gx + static_cast<long long>(t * 3 + ci) * ldg + co);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const long long wi = static_cast<long long>(t * 3 + c) * cout + co;
        const float wv = __ldg(w + wi);
        dw[wi] += m1[0 * 3 + c] * g[0] + m1[1 * 3 + c] * g[1] + m1[2 * 3 + c] * g[2] + c1[c] * dbv;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) acc[ci * 3 + c] += wv * g[ci];
        acc[9 + c] += wv * dbv;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    float v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&red[j], v);
  }
  __syncthreads();
  if (threadIdx.x < 9) atomicAdd(dw1 + threadIdx.x, red[threadIdx.x]);
  else if (threadIdx.x < 12) atomicAdd(db1 + threadIdx.x - 9, red[threadIdx.x]);
}

// ------------------------------------------------------------------ pooling
// tf.layers.max_pooling2d 2x2 / 2 (unet.py:52)
__global__ void maxpool2x2_kernel(const uint4* __restrict__ in, int N, int H, int W, int G,
                                  uint4* __restrict__ out) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = 1LL * N * Ho * Wo * G;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total;
       i += 1LL * gridDim.x * blockDim.x) {
    const int g = static_cast<int>(i % G);
    const long long p = i / G;
    const int x = static_cast<int>(p % Wo);
    const int y = static_cast<int>((p / Wo) % Ho);
    const int n = static_cast<int>(p / (1LL * Wo * Ho));
    const long long base = ((1LL * n * H + 2 * y) * W + 2 * x) * G + g;
    float a[8], b[8];
    unpack8(__ldg(in + base), a);
    unpack8(__ldg(in + base + G), b);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = fmaxf(a[e], b[e]);
    unpack8(__ldg(in + base + 1LL * W * G), b);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = fmaxf(a[e], b[e]);
    unpack8(__ldg(in + base + 1LL * W * G + G), b);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = fmaxf(a[e], b[e]);
    out[i] = pack8(a);
  }
}

// MaxPoolGrad + crop-pad of the concat gradient + ReluGrad in one pass over the skip tensor
// (unet.py:47-52, 70-85).  The pool gradient goes to the first maximum of each 2x2 window.
// One thread owns one 2x2 window x 8 channels: Y and dZ are touched exactly once, and the two
// horizontally adjacent pixels of a window are contiguous in NHWC, so a warp moves 1 KiB runs.
// WINDOWED = false (no pool below, odd extents allowed): one thread per pixel x 8 channels.
template <bool WINDOWED>
__global__ void __launch_bounds__(128, 8) skip_grad_kernel(const uint4* __restrict__ Y, int N, int H, int W, int G,
                                 const uint4* __restrict__ dP, const __nv_bfloat16* __restrict__ dC,
                                 long long c_sn, long long c_sy, long long c_sx, int Hc, int Wc,
                                 int crop_y, int crop_x, uint4* __restrict__ dZ) {
  constexpr int S = WINDOWED ? 2 : 1;
  const int Hw = H / S, Ww = W / S;
  // grid = (ceil(Ww*G / blockDim), Hw, N): no 64-bit divisions (this kernel was instruction-,
  // not HBM-bound with a flat grid-stride index)
  const int wy = blockIdx.y, n = blockIdx.z;
  const int xg = blockIdx.x * blockDim.x + threadIdx.x;
  if (xg < Ww * G) {
    const int wx = xg / G, g = xg - wx * G;
    const long long i = (1LL * n * Hw + wy) * (Ww * G) + xg;
    // all loads first (Y, pool gradient, cropped concat gradient: up to nine 16-byte loads in
    // flight per thread), then the arithmetic; out-of-crop pixels re-read their own Y vector (an
    // L1 hit) instead of branching around the load
    long long idx[S * S];
    uint4 yraw[S * S], craw[S * S];
    bool in_crop[S * S];
#pragma unroll
    for (int q = 0; q < S * S; ++q) {
      const int y = wy * S + q / S, x = wx * S + q % S;
      idx[q] = ((1LL * n * H + y) * W + x) * G + g;
      yraw[q] = __ldg(Y + idx[q]);
    }
    uint4 praw = make_uint4(0, 0, 0, 0);
    if (WINDOWED && dP != nullptr) praw = __ldg(dP + i);  // dP is [N, H/2, W/2, C]: same linear index as the window
#pragma unroll
    for (int q = 0; q < S * S; ++q) {
      const int cy = wy * S + q / S - crop_y, cx = wx * S + q % S - crop_x;
      in_crop[q] = dC != nullptr && cy >= 0 && cy < Hc && cx >= 0 && cx < Wc;
      const uint4* cp = in_crop[q] ? reinterpret_cast<const uint4*>(dC + n * c_sn + cy * c_sy + cx * c_sx) + g
                                   : Y + idx[q];
      craw[q] = __ldg(cp);
    }
    // arg-max position of every channel's 2x2 window (first maximum wins, like MaxPoolGrad),
    // two bits per channel; then one output vector at a time, so that only the raw 16-byte
    // vectors stay live (register pressure decides the occupancy of this HBM-bound kernel)
    uint32_t argbits = 0;
    if (WINDOWED && dP != nullptr) {
#pragma unroll
      for (int w = 0; w < 4; ++w) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          int arg = 0;
          float best = 0.f;
#pragma unroll
          for (int q = 0; q < S * S; ++q) {
            const uint32_t word = w == 0 ? yraw[q].x : w == 1 ? yraw[q].y : w == 2 ? yraw[q].z : yraw[q].w;
            const float v = h == 0 ? bf16_lo(word) : bf16_hi(word);
            if (q == 0 || v > best) {
              best = v;
              arg = q;
            }
          }
          argbits |= static_cast<uint32_t>(arg) << (2 * (2 * w + h));
        }
      }
    }
#pragma unroll
    for (int q = 0; q < S * S; ++q) {
      uint32_t ow[4];
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const uint32_t yw = w == 0 ? yraw[q].x : w == 1 ? yraw[q].y : w == 2 ? yraw[q].z : yraw[q].w;
        const uint32_t pw = w == 0 ? praw.x : w == 1 ? praw.y : w == 2 ? praw.z : praw.w;
        const uint32_t cw = w == 0 ? craw[q].x : w == 1 ? craw[q].y : w == 2 ? craw[q].z : craw[q].w;
        float g2[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float yv = h == 0 ? bf16_lo(yw) : bf16_hi(yw);
          float g = 0.f;
          if (WINDOWED && dP != nullptr) {
            const int arg = static_cast<int>((argbits >> (2 * (2 * w + h))) & 3u);
            if (arg == q) g = h == 0 ? bf16_lo(pw) : bf16_hi(pw);
          }
          if (in_crop[q]) g += h == 0 ? bf16_lo(cw) : bf16_hi(cw);
          g2[h] = yv > 0.f ? g : 0.f;
        }
        ow[w] = pack_bf16x2(g2[0], g2[1]);
      }
      dZ[idx[q]] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
  }
}

// ReluGrad on strided views: dZ (compact NHWC) = [Y > 0] * dY
__global__ void relu_mask_kernel(const __nv_bfloat16* __restrict__ Y, long long y_sn,
                                 long long y_sy, long long y_sx, const __nv_bfloat16* __restrict__ dY,
                                 long long d_sn, long long d_sy, long long d_sx, int N, int H, int W,
                                 int G, uint4* __restrict__ dZ) {
  const long long total = 1LL * N * H * W * G;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total;
       i += 1LL * gridDim.x * blockDim.x) {
    const int g = static_cast<int>(i % G);
    const long long p = i / G;
    const int x = static_cast<int>(p % W);
    const int y = static_cast<int>((p / W) % H);
    const int n = static_cast<int>(p / (1LL * W * H));
    float yv[8], dv[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(Y + n * y_sn + y * y_sy + x * y_sx) + g), yv);
    unpack8(__ldg(reinterpret_cast<const uint4*>(dY + n * d_sn + y * d_sy + x * d_sx) + g), dv);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (!(yv[e] > 0.f)) dv[e] = 0.f;
    dZ[i] = pack8(dv);
  }
}

// BiasAddGrad: out[c] += sum over pixels.  blockDim = G * k threads; thread (lane, g).  Blocks
// walk image rows (32-bit index math only), four independent 16-byte loads in flight per thread.
__global__ void bias_grad_kernel(const __nv_bfloat16* __restrict__ v, long long sn, long long sy,
                                 long long sx, int N, int H, int W, int G, int k,
                                 float* __restrict__ out) {
  extern __shared__ float sred[];  // [G*8]
  const int g = threadIdx.x % G, lane = threadIdx.x / G;
  for (int j = threadIdx.x; j < G * 8; j += blockDim.x) sred[j] = 0.f;
  __syncthreads();
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  const int rows = N * H;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int n = row / H, y = row - n * H;
    const __nv_bfloat16* __restrict__ base = v + n * sn + y * sy + g * 8;
    for (int x0 = lane; x0 < W; x0 += 4 * k) {
      uint4 raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int x = x0 + u * k;
        raw[u] = make_uint4(0, 0, 0, 0);
        if (x < W) raw[u] = __ldg(reinterpret_cast<const uint4*>(base + x * sx));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        unpack8(raw[u], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[e];
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) atomicAdd(&sred[g * 8 + e], acc[e]);
  __syncthreads();
  for (int j = threadIdx.x; j < G * 8; j += blockDim.x) atomicAdd(out + j, sred[j]);
}

// ------------------------------------------------------------------ head
// weight_output 1x1 conv (unet.py:95) + softmax P(road) (tf_aerial_images.py:147-148) + mean
// sparse softmax cross-entropy (:103-110) and all gradients that leave this layer.
// LP = C/8 lanes cooperate on one pixel (each owns 8 channels).
// Two classes: everything depends on the logit difference d = l1 - l0 only --
//   P(road) = sigmoid(d),  CE = log(1 + e^-|d|) + (label is the larger logit ? 0 : |d|),
//   dl1 = (P(road) - label) / count = -dl0,  dZ = dl1 * (w1 - w0),  dW[:,0] = -dW[:,1]
// -- which halves the arithmetic of this instruction-bound (not yet HBM-bound) kernel.  The
// separate logits l0, l1 are only evaluated when the caller asks for them (LOGITS).
template <int LP, bool LOGITS, int U_ = 4>
__global__ void __launch_bounds__(256, 3)
    head_kernel(const uint4* __restrict__ act, long long pixels, const float* __restrict__ w,
                const float* __restrict__ b, const unsigned char* __restrict__ labels,
                float* __restrict__ probs, float* __restrict__ logits, float* __restrict__ loss,
                uint4* __restrict__ dZ, float* __restrict__ dW, float* __restrict__ db,
                float inv_count) {
  constexpr int PPW = 32 / LP;  // pixels per warp
  const int lane = threadIdx.x & 31;
  const int sub = lane % LP;  // channel group
  const int pw = lane / LP;   // pixel slot in warp
  float w0[LOGITS ? 8 : 1], wd[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float a0 = __ldg(w + (sub * 8 + e) * 2 + 0), a1 = __ldg(w + (sub * 8 + e) * 2 + 1);
    if (LOGITS) w0[e] = a0;
    wd[e] = a1 - a0;
  }
  const float b0 = __ldg(b), bd = __ldg(b + 1) - b0;
  float aw1[8], ab1 = 0.f, aloss = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) aw1[e] = 0.f;

  const long long warp_global = (blockIdx.x * 1LL * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = (1LL * gridDim.x * blockDim.x) >> 5;
  // U pixel groups per iteration: all activation (and label) loads are issued before the first
  // dependent instruction
  // (U_ = 8 for prediction: no gradient state in registers, so twice the loads fit in flight)
  constexpr int U = U_;
  for (long long p0 = warp_global * (PPW * U); p0 < pixels; p0 += n_warps * (PPW * U)) {
    uint4 raw[U];
    int lab[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = p0 + u * PPW + pw;
      raw[u] = make_uint4(0, 0, 0, 0);
      lab[u] = 0;
      if (p < pixels) {
        raw[u] = __ldg(act + p * LP + sub);
        if (labels != nullptr) lab[u] = labels[p];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = p0 + u * PPW + pw;
      const bool ok = p < pixels;
      float a[8];
      unpack8(raw[u], a);
      float d = 0.f, l0 = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        d += a[e] * wd[e];
        if (LOGITS) l0 += a[e] * w0[e];
      }
#pragma unroll
      for (int o = LP / 2; o > 0; o >>= 1) {
        d += __shfl_xor_sync(0xffffffffu, d, o);
        if (LOGITS) l0 += __shfl_xor_sync(0xffffffffu, l0, o);
      }
      d += bd;
      const float ad = fabsf(d);
      const float en = __expf(-ad);         // e^-|d|  (softmax numerator of the smaller logit)
      const float inv = 1.f / (1.f + en);
      const float p1 = d >= 0.f ? inv : en * inv;
      if (ok && sub == 0) {
        if (probs) probs[p] = p1;
        if (LOGITS) {
          logits[p * 2] = l0 + b0;
          logits[p * 2 + 1] = l0 + b0 + d;
        }
      }
      if (labels != nullptr && ok) {
        const float dl1 = (p1 - (lab[u] ? 1.f : 0.f)) * inv_count;
        if (sub == 0) {
          aloss += __logf(1.f + en) + (((d >= 0.f) == (lab[u] != 0)) ? 0.f : ad);
          ab1 += dl1;
        }
        float gz[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          aw1[e] += a[e] * dl1;
          gz[e] = a[e] > 0.f ? dl1 * wd[e] : 0.f;
        }
        dZ[p * LP + sub] = pack8(gz);
      }
    }
  }
  if (labels == nullptr) return;
  // reduce across pixel slots of the warp, then across warps through shared memory
  __shared__ float s_w[LP * 8];
  __shared__ float s_s[2];
  for (int j = threadIdx.x; j < LP * 8; j += blockDim.x) s_w[j] = 0.f;
  if (threadIdx.x < 2) s_s[threadIdx.x] = 0.f;
  __syncthreads();
#pragma unroll
  for (int e = 0; e < 8; ++e) {
#pragma unroll
    for (int o = LP; o < 32; o <<= 1) aw1[e] += __shfl_xor_sync(0xffffffffu, aw1[e], o);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    aloss += __shfl_xor_sync(0xffffffffu, aloss, o);
    ab1 += __shfl_xor_sync(0xffffffffu, ab1, o);
  }
  if (pw == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(&s_w[sub * 8 + e], aw1[e]);
  }
  if (lane == 0) {
    atomicAdd(&s_s[0], aloss);
    atomicAdd(&s_s[1], ab1);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < LP * 8; j += blockDim.x) {
    atomicAdd(dW + 2 * j + 1, s_w[j]);
    atomicAdd(dW + 2 * j, -s_w[j]);
  }
  if (threadIdx.x == 0) {
    atomicAdd(loss, s_s[0] * inv_count);
    atomicAdd(db, -s_s[1]);
    atomicAdd(db + 1, s_s[1]);
  }
}

// ------------------------------------------------------------------ dropout / SGD
// tf.nn.dropout (unet.py:30, 65): y = x / keep * floor(keep + U)
__global__ void dropout_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long long n8,
                               float keep, unsigned long long seed) {
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < n8;
       i += 1LL * gridDim.x * blockDim.x) {
    float f[8];
    unpack8(__ldg(x + i), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] *= keep_scale(seed, i * 8 + e, keep);
    y[i] = pack8(f);
  }
}
__global__ void dropout_mask_kernel(float* __restrict__ m, long long n, float keep,
                                    unsigned long long seed) {
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < n;
       i += 1LL * gridDim.x * blockDim.x)
    m[i] = keep_scale(seed, i, keep);
}

// tf.train.MomentumOptimizer, non-Nesterov (tf_aerial_images.py:116-121)
__global__ void momentum_sgd_kernel(float4* __restrict__ w, float4* __restrict__ acc,
                                    const float4* __restrict__ g, long long n4, float lr,
                                    float momentum, float gscale) {
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < n4;
       i += 1LL * gridDim.x * blockDim.x) {
    float4 a = acc[i], ww = w[i];
    const float4 gg = __ldg(g + i);
    a.x = momentum * a.x + gg.x * gscale;
    a.y = momentum * a.y + gg.y * gscale;
    a.z = momentum * a.z + gg.z * gscale;
    a.w = momentum * a.w + gg.w * gscale;
    ww.x -= lr * a.x;
    ww.y -= lr * a.y;
    ww.z -= lr * a.z;
    ww.w -= lr * a.w;
    acc[i] = a;
    w[i] = ww;
  }
}
__global__ void momentum_sgd_tail_kernel(float* w, float* acc, const float* g, long long begin,
                                         long long n, float lr, float momentum, float gscale) {
  const long long i = begin + blockIdx.x * 1LL * blockDim.x + threadIdx.x;
  if (i < n) {
    const float a = momentum * acc[i] + g[i] * gscale;
    acc[i] = a;
    w[i] -= lr * a;
  }
}

static int grid_for(long long work_items, int threads) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = 1LL * num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace rsu

using namespace rsu;

extern "C" {

int rsu_pack_transpose(const float* in, void* out, int T, int R, int C, int ld, void* stream) {
  if (T < 1 || R < 1 || C < 1) return set_error(RSU_EINVAL, "pack_transpose: empty");
  if (ld == 0) ld = T * R;
  if (ld < T * R) return set_error(RSU_EINVAL, "pack_transpose: ld %d < %d", ld, T * R);
  dim3 grid((C + 31) / 32, (R + 31) / 32, T), block(32, 8);
  pack_transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(
      in, static_cast<__nv_bfloat16*>(out), T, R, C, ld);
  return check_launch("pack_transpose");
}

int rsu_pack_permute(const float* in, void* out, int T, int R, int C, const int* perm_host,
                     void* stream) {
  if (T < 1 || T > 16 || R < 1 || C < 1) return set_error(RSU_EINVAL, "pack_permute: bad T=%d", T);
  Perm perm;
  for (int t = 0; t < 16; ++t) perm.v[t] = perm_host && t < T ? perm_host[t] : t;
  const long long total = 1LL * T * R * C;
  pack_permute_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      in, static_cast<__nv_bfloat16*>(out), T, R, C, perm);
  return check_launch("pack_permute");
}

int rsu_cast_bf16(const float* in, void* out, long long n, void* stream) {
  if (n < 1) return set_error(RSU_EINVAL, "cast: empty");
  cast_bf16_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
      in, static_cast<__nv_bfloat16*>(out), n);
  return check_launch("cast_bf16");
}

int rsu_color_im2col(const float* img, int N, int S, const float* w1, const float* b1,
                     int dilation, int oy, int ox, int Ho, int Wo, void* out, float keep,
                     unsigned long long seed, void* stream) {
  if (oy < 0 || ox < 0 || oy + Ho + 2 * dilation > S || ox + Wo + 2 * dilation > S)
    return set_error(RSU_EINVAL, "color_im2col: window outside the %dx%d image", S, S);
  if (reinterpret_cast<uintptr_t>(out) & 15) return set_error(RSU_EALIGN, "color_im2col: out");
  const long long total = 1LL * N * Ho * Wo;
  color_im2col_kernel<<<grid_for(total, 128), 128, 0, (cudaStream_t)stream>>>(
      img, N, S, w1, b1, dilation, oy, ox, Ho, Wo, static_cast<__nv_bfloat16*>(out), keep, seed);
  return check_launch("color_im2col");
}

int rsu_color_im2col_bwd(const float* img, int N, int S, const void* dcol, int dilation, int oy,
                         int ox, int Ho, int Wo, float* dw1, float* db1, float keep,
                         unsigned long long seed, void* stream) {
  if (oy < 0 || ox < 0 || oy + Ho + 2 * dilation > S || ox + Wo + 2 * dilation > S)
    return set_error(RSU_EINVAL, "color_im2col_bwd: window outside the %dx%d image", S, S);
  const long long total = 1LL * N * (Ho + 2 * dilation) * (Wo + 2 * dilation);
  int grid = grid_for(total, 256);
  if (grid > num_sms() * 4) grid = num_sms() * 4;
  color_im2col_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      img, N, S, static_cast<const __nv_bfloat16*>(dcol), dilation, oy, ox, Ho, Wo, dw1, db1, keep,
      seed);
  return check_launch("color_im2col_bwd");
}

int rsu_pack_plan(const rsu_pack_job* jobs_host, int n_jobs, void* table_dev, int* total_blocks) {
  if (!jobs_host || n_jobs < 1 || n_jobs > 4096 || !table_dev || !total_blocks)
    return set_error(RSU_EINVAL, "pack_plan: bad arguments");
  PackJobDev* tab = new PackJobDev[n_jobs];
  int blocks = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const rsu_pack_job& h = jobs_host[i];
    if (h.kind < 0 || h.kind > 2 || h.T < 1 || h.R < 1 || h.C < 1) {
      delete[] tab;
      return set_error(RSU_EINVAL, "pack_plan: job %d invalid", i);
    }
    PackJobDev& d = tab[i];
    d.in = h.in;
    d.out = static_cast<__nv_bfloat16*>(h.out);
    d.kind = h.kind;
    d.T = h.T;
    d.R = h.R;
    d.C = h.C;
    d.ld = h.ld > 0 ? h.ld : h.T * h.R;
    d.block_begin = blocks;
    if (h.kind == 0) {
      d.block_count = h.T * ((h.R + 63) / 64) * ((h.C + 31) / 32);
    } else {
      const long long total = 1LL * h.T * h.R * h.C;
      long long b = (total + 256 * 32 - 1) / (256 * 32);  // 32 elements (4 vector groups) per thread
      d.block_count = static_cast<int>(b < 1 ? 1 : b);
    }
    blocks += d.block_count;
  }
  cudaError_t e = cudaMemcpy(table_dev, tab, sizeof(PackJobDev) * n_jobs, cudaMemcpyHostToDevice);
  delete[] tab;
  if (e != cudaSuccess) return set_error(RSU_ECUDA, "pack_plan: %s", cudaGetErrorString(e));
  *total_blocks = blocks;
  return RSU_OK;
}

int rsu_pack_plan_bytes(int n_jobs) { return static_cast<int>(sizeof(PackJobDev)) * n_jobs; }

int rsu_pack_run(const void* table_dev, int n_jobs, int total_blocks, void* stream) {
  if (!table_dev || n_jobs < 1 || total_blocks < 1) return set_error(RSU_EINVAL, "pack_run: bad arguments");
  pack_batch_kernel<<<total_blocks, dim3(32, 8), 0, (cudaStream_t)stream>>>(
      static_cast<const PackJobDev*>(table_dev), n_jobs);
  return check_launch("pack_batch");
}

int rsu_first_layer_fold(const float* w, const float* b, const float* w1, const float* b1, int cout,
                         void* w_packed, float* bias_eff, void* stream) {
  if (cout < 1) return set_error(RSU_EINVAL, "first_layer_fold: cout=%d", cout);
  first_layer_fold_kernel<<<(cout + 63) / 64, 64, 0, (cudaStream_t)stream>>>(
      w, b, w1, b1, cout, static_cast<__nv_bfloat16*>(w_packed), bias_eff);
  return check_launch("first_layer_fold");
}

int rsu_first_layer_grads(const float* gx, int ldg, const float* w, const float* w1, const float* b1,
                          int cout, float* dw, float* dbias, float* dw1, float* db1, void* stream) {
  if (cout < 1 || ldg < cout) return set_error(RSU_EINVAL, "first_layer_grads: cout=%d ldg=%d", cout, ldg);
  int threads = ((cout + 31) / 32) * 32;
  if (threads > 256) threads = 256;
  first_layer_grads_kernel<<<1, threads, 0, (cudaStream_t)stream>>>(gx, ldg, w, w1, b1, cout, dw,
                                                                    dbias, dw1, db1);
  return check_launch("first_layer_grads");
}

int rsu_maxpool2x2(const void* in, int N, int H, int W, int C, void* out, void* stream) {
  if (H % 2 || W % 2 || C % 8) return set_error(RSU_EINVAL, "maxpool: H=%d W=%d C=%d", H, W, C);
  const long long total = 1LL * N * (H / 2) * (W / 2) * (C / 8);
  maxpool2x2_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      static_cast<const uint4*>(in), N, H, W, C / 8, static_cast<uint4*>(out));
  return check_launch("maxpool2x2");
}

int rsu_skip_grad(const void* Y, int N, int H, int W, int C, const void* dP, const rsu_view* dCrop,
                  int crop_y, int crop_x, void* dZ, void* stream) {
  if (C % 8) return set_error(RSU_EINVAL, "skip_grad: C=%d", C);
  if (dP && (H % 2 || W % 2)) return set_error(RSU_EINVAL, "skip_grad: odd H/W with pool grad");
  if (dCrop && (dCrop->C != C || dCrop->N != N || (dCrop->sx % 8) || (dCrop->sy % 8) ||
                (dCrop->sn % 8) || (reinterpret_cast<uintptr_t>(dCrop->ptr) & 15)))
    return set_error(RSU_EINVAL, "skip_grad: bad crop view");
#define RSU_SKIP_ARGS                                                                          \
  static_cast<const uint4*>(Y), N, H, W, C / 8, static_cast<const uint4*>(dP),                   \
      dCrop ? static_cast<const __nv_bfloat16*>(dCrop->ptr) : nullptr, dCrop ? dCrop->sn : 0,    \
      dCrop ? dCrop->sy : 0, dCrop ? dCrop->sx : 0, dCrop ? dCrop->H : 0, dCrop ? dCrop->W : 0,  \
      crop_y, crop_x, static_cast<uint4*>(dZ)
  if (N > 65535 || H > 65535) return set_error(RSU_EINVAL, "skip_grad: N or H > 65535");
  if (dP) {
    const dim3 grid(((W / 2) * (C / 8) + 127) / 128, H / 2, N);
    skip_grad_kernel<true><<<grid, 128, 0, (cudaStream_t)stream>>>(RSU_SKIP_ARGS);
  } else {
    const dim3 grid((W * (C / 8) + 127) / 128, H, N);
    skip_grad_kernel<false><<<grid, 128, 0, (cudaStream_t)stream>>>(RSU_SKIP_ARGS);
  }
#undef RSU_SKIP_ARGS
  return check_launch("skip_grad");
}

int rsu_relu_mask(const rsu_view* Y, const rsu_view* dY, void* dZ, void* stream) {
  if (!Y || !dY || Y->C != dY->C || Y->H != dY->H || Y->W != dY->W || Y->N != dY->N || Y->C % 8)
    return set_error(RSU_EINVAL, "relu_mask: view mismatch");
  if ((Y->sx % 8) || (Y->sy % 8) || (Y->sn % 8) || (dY->sx % 8) || (dY->sy % 8) || (dY->sn % 8) ||
      (reinterpret_cast<uintptr_t>(Y->ptr) & 15) || (reinterpret_cast<uintptr_t>(dY->ptr) & 15))
    return set_error(RSU_EALIGN, "relu_mask: alignment");
  const long long total = 1LL * Y->N * Y->H * Y->W * (Y->C / 8);
  relu_mask_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      static_cast<const __nv_bfloat16*>(Y->ptr), Y->sn, Y->sy, Y->sx,
      static_cast<const __nv_bfloat16*>(dY->ptr), dY->sn, dY->sy, dY->sx, Y->N, Y->H, Y->W,
      Y->C / 8, static_cast<uint4*>(dZ));
  return check_launch("relu_mask");
}

int rsu_bias_grad(const rsu_view* v, float* out, void* stream) {
  if (!v || v->C % 8 || v->C / 8 > 1024) return set_error(RSU_EINVAL, "bias_grad: C");
  if ((v->sx % 8) || (v->sy % 8) || (v->sn % 8) || (reinterpret_cast<uintptr_t>(v->ptr) & 15))
    return set_error(RSU_EALIGN, "bias_grad: alignment");
  const int G = v->C / 8;
  int k = 256 / G;
  if (k < 1) k = 1;
  if (1LL * v->N * v->H > 0x7fffffffLL) return set_error(RSU_EINVAL, "bias_grad: too many rows");
  long long blocks = 1LL * v->N * v->H;
  const long long cap = 1LL * num_sms() * 8;
  if (blocks > cap) blocks = cap;
  bias_grad_kernel<<<static_cast<int>(blocks), G * k, G * 8 * sizeof(float), (cudaStream_t)stream>>>(
      static_cast<const __nv_bfloat16*>(v->ptr), v->sn, v->sy, v->sx, v->N, v->H, v->W, G, k, out);
  return check_launch("bias_grad");
}

int rsu_head(const void* act, int N, int H, int W, int C, const float* w, const float* b,
             const unsigned char* labels, float* probs, float* logits, float* loss, void* dZ,
             float* dW, float* db, void* stream) {
  const long long pixels = 1LL * N * H * W;
  if (labels && (!loss || !dZ || !dW || !db))
    return set_error(RSU_EINVAL, "head: training outputs missing");
  const float inv_count = 1.0f / static_cast<float>(pixels);
  const int threads = 256;
  int grid = grid_for(pixels * (C / 8), threads);
  if (grid > num_sms() * 6) grid = num_sms() * 6;  // 3 resident blocks per SM (80 registers): two full rounds
#define RSU_HEAD(LP)                                                                         \
  do {                                                                                       \
    if (logits)                                                                              \
      head_kernel<LP, true><<<grid, threads, 0, (cudaStream_t)stream>>>(                     \
          static_cast<const uint4*>(act), pixels, w, b, labels, probs, logits, loss,        \
          static_cast<uint4*>(dZ), dW, db, inv_count);                                      \
    else if (labels == nullptr)                                                              \
      head_kernel<LP, false, 8><<<grid, threads, 0, (cudaStream_t)stream>>>(                 \
          static_cast<const uint4*>(act), pixels, w, b, labels, probs, logits, loss,        \
          static_cast<uint4*>(dZ), dW, db, inv_count);                                      \
    else                                                                                     \
      head_kernel<LP, false><<<grid, threads, 0, (cudaStream_t)stream>>>(                    \
          static_cast<const uint4*>(act), pixels, w, b, labels, probs, logits, loss,        \
          static_cast<uint4*>(dZ), dW, db, inv_count);                                      \
  } while (0)
  if (C == 64) RSU_HEAD(8);
  else if (C == 128) RSU_HEAD(16);
  else if (C == 256) RSU_HEAD(32);
  else return set_error(RSU_EINVAL, "head: C=%d unsupported (64/128/256)", C);
#undef RSU_HEAD
  return check_launch("head");
}

int rsu_dropout(const void* x, void* y, long long n, float keep, unsigned long long seed,
                void* stream) {
  if (n % 8) return set_error(RSU_EINVAL, "dropout: n %% 8");
  if (!(keep > 0.f && keep <= 1.f)) return set_error(RSU_EINVAL, "dropout: keep=%f", keep);
  dropout_kernel<<<grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(y), n / 8, keep, seed);
  return check_launch("dropout");
}

int rsu_dropout_mask(float* m, long long n, float keep, unsigned long long seed, void* stream) {
  dropout_mask_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(m, n, keep, seed);
  return check_launch("dropout_mask");
}

int rsu_momentum_sgd(float* w, float* acc, const float* g, long long n, float lr, float momentum,
                     float gscale, void* stream) {
  if (n < 1) return set_error(RSU_EINVAL, "sgd: empty");
  if ((reinterpret_cast<uintptr_t>(w) & 15) || (reinterpret_cast<uintptr_t>(acc) & 15) ||
      (reinterpret_cast<uintptr_t>(g) & 15))
    return set_error(RSU_EALIGN, "sgd: pointers must be 16-byte aligned");
  const long long n4 = n / 4;
  if (n4 > 0) {
    momentum_sgd_kernel<<<grid_for(n4, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<float4*>(w), reinterpret_cast<float4*>(acc),
        reinterpret_cast<const float4*>(g), n4, lr, momentum, gscale);
    int rc = check_launch("momentum_sgd");
    if (rc) return rc;
  }
  if (n4 * 4 < n) {
    momentum_sgd_tail_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(w, acc, g, n4 * 4, n, lr, momentum,
                                                                 gscale);
    return check_launch("momentum_sgd_tail");
  }
  return RSU_OK;
}

}  // extern "C"
