/*
 * rsu_b200.h -- C ABI of the B200-native U-Net hot path (librsu_b200.so).
 *
 * The reference (aschneuw/road-segmentation-unet) has no FFI: its boundary is TensorFlow's
 * Session.run() into stock op kernels (src/tf_aerial_images.py:241-244, :312) plus NumPy/SciPy
 * helpers (src/images.py).  Every entry point below replaces one such op (or one NumPy helper)
 * and cites the reference call site it stands in for.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name says "host"; the caller owns every buffer;
 *   - activations are NHWC bf16, described by rsu_view (element strides, so channel slices and
 *     crops of a larger tensor are expressed without copies);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return value: 0 = ok, otherwise an RSU_E* code; rsu_last_error() gives the message.
 *   - there is no CPU fallback: unsupported shapes are errors.
 */
#ifndef RSU_B200_H
#define RSU_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSU_OK 0
#define RSU_EINVAL 1  /* bad shape / unsupported configuration            */
#define RSU_EALIGN 2  /* pointer or stride violates an alignment rule      */
#define RSU_ECUDA 3   /* CUDA runtime / driver error                       */
#define RSU_ENODEV 4  /* no sm_100 device                                  */

#define RSU_MAX_SRC 4
#define RSU_MAX_TAPS 9

const char* rsu_last_error(void);
int rsu_version(void);
/* Number of kernels launched by this library since load / since the last reset. */
long long rsu_launch_count(void);
void rsu_reset_launch_count(void);
/* The same count per kernel, as "name=count;name=count;..." written into buf_host (at most cap
 * bytes incl. the terminator); returns the size needed.  Lets the bench check that a committed ncu
 * capture describes the kernels that actually ran. */
int rsu_launch_histogram(char* buf_host, int cap);

/* CRC-32C (Castagnoli) of a HOST buffer, continuing from `crc` (0 to start): the checksum
 * TensorFlow's checkpoint bundle stores per tensor and per index block (tf.train.Saver,
 * tf_aerial_images.py:171, :343-379 -> tensorflow/core/util/tensor_bundle). */
unsigned int rsu_crc32c_host(unsigned int crc, const void* data_host, unsigned long long n);

/* NHWC bf16 activation view: element (n, y, x, c) lives at ptr + n*sn + y*sy + x*sx + c. */
typedef struct {
  const void* ptr;
  int C, H, W, N;       /* extents of this view (C % 64 == 0 for GEMM operands) */
  long long sn, sy, sx; /* strides in ELEMENTS (sx % 8 == 0)                    */
  int off_y, off_x;     /* offset added to the output pixel coordinate (crop)   */
} rsu_view;

/* ---------------------------------------------------------------- tcgen05 implicit GEMM ---- */

/* Generic implicit-GEMM convolution on the 5th-gen tensor cores:
 *   out[n, y, x, co] = epilogue( sum_{tap, src, c} src[n, y+dy[tap]+off_y, x+dx[tap]+off_x, c]
 *                                                  * weights[co][(tap, src, c)] )
 * Replaces tf.layers.conv2d / Conv2DBackpropInput (unet.py:34-45, 88-91) including the crop +
 * concat of unet.py:70-85 (several sources walked by the K loop) and, with shuffle_cout > 0,
 * tf.layers.conv2d_transpose k=2 s=2 (unet.py:67): GEMM column (a, b, co) is stored at output
 * pixel (2y+a, 2x+b).  Out-of-range source pixels read as zero (used by the data gradients). */
typedef struct {
  int n_src;
  rsu_view src[RSU_MAX_SRC];
  int n_taps;
  int tap_dy[RSU_MAX_TAPS], tap_dx[RSU_MAX_TAPS];
  const void* weights; /* bf16 [Ntot][Ktot], Ktot = n_taps * sum(src.C), K contiguous */
  int Ntot;
  int H_out, W_out, N_img; /* GEMM pixel grid */
  void* out;               /* bf16 */
  long long out_sn, out_sy, out_sx;
  int shuffle_cout;  /* 0, or Cout of a 2x2/s2 transpose conv (Ntot == 4*shuffle_cout) */
  const float* bias; /* fp32 [Ntot] ([shuffle_cout] when shuffling) or NULL */
  int relu;
  const void* mask; /* bf16, same indexing as out with its own strides: keep where mask > 0 */
  long long mask_sn, mask_sy, mask_sx;
  int accumulate; /* out += result */
  int algo;       /* 0 = choose, 1 = one TMA box per tap, 2 = halo tile shared by all taps,
                     3 = algorithm 1 on CTA pairs (tcgen05.mma.cta_group::2, N tile 128 / 256;
                     chosen by 0 whenever algorithm 1 would run with the 256-wide N tile),
                     4 = algorithm 2 on CTA pairs (each CTA keeps half of the resident weights) */
  /* mask_nc > 0: the mask tensor has mask_nc channels and covers output channels
   * [mask_c0, mask_c0 + mask_nc) only (the ReluGrad of one member of a concat gradient,
   * unet.py:70-85); both multiples of the N tile (64 / 128 / 256, the largest dividing Ntot). */
  int mask_c0, mask_nc;
  /* optional fused 2x2 / stride-2 max pool of the (ReLU) output (tf.layers.max_pooling2d,
   * unet.py:52): bf16 [N_img, H_out/2, W_out/2, Ntot] with element strides.  Only the halo-tile
   * kernel's coalesced epilogue implements it; *pool_done_host tells whether it was written
   * (0: the caller still has to run rsu_maxpool2x2). */
  void* pool_out;
  long long pool_sn, pool_sy, pool_sx;
  int* pool_done_host;
} rsu_conv_gemm_desc;
int rsu_conv_gemm(const rsu_conv_gemm_desc* d, void* stream);

/* Weight gradient on the tensor cores (Conv2DBackpropFilter, unet.py:34-45,67,88-91):
 *   out[(tap, src, c), co] += sum_{n,y,x} src[n, y+dy+off_y, x+dx+off_x, c] * grad[n, y+goff, x+goff, co]
 * out is fp32 [n_taps*sum(src.C), Cout] -- i.e. TensorFlow's HWIO kernel layout -- and must be
 * zero-initialised by the caller (partial sums over pixel splits are combined with atomics). */
typedef struct {
  int n_src;
  rsu_view src[RSU_MAX_SRC];
  int n_taps;
  int tap_dy[RSU_MAX_TAPS], tap_dx[RSU_MAX_TAPS];
  rsu_view grad;  /* C = Cout; off_y/off_x offset the grad pixel coordinate */
  int H, W, N_img; /* pixel grid that is summed over */
  float* out;
  int ldo; /* row stride of out (>= grad.C) */
  float* bias_grad; /* optional fp32 [Cout]: += column sums of grad over the pixel grid
                       (BiasAddGrad fused into the halo-tile kernel; NULL = not wanted).  On
                       return *bias_done tells whether the kernel produced it. */
  int* bias_done_host;
  int algo; /* 0 = choose, 1 = one TMA box per tap pair, 2 = halo tile shared by all taps,
               3 = algorithm 1 on CTA pairs (tcgen05.mma.cta_group::2; no bias gradient; chosen by 0
               for 3x3 layers with >= 256 gradient channels and pixel grids up to 104^2) */
} rsu_wgrad_desc;
int rsu_wgrad_gemm(const rsu_wgrad_desc* d, void* stream);

/* ---------------------------------------------------------------- weight packing ---------- */
/* fp32 in[T][R][C] -> bf16 out[C][ld] with out[c][t*R + r] (transpose of every [R][C] slab; conv
 * forward weights: HWIO [taps][Cin][Cout] -> [Cout][taps][Cin]).  ld = 0 means T*R; a larger ld
 * leaves the tail of every output row untouched (zero padding of the Cin = 3 layers). */
int rsu_pack_transpose(const float* in, void* out, int T, int R, int C, int ld, void* stream);
/* fp32 in[T][R][C] -> bf16 out[R][T'][C] with T' = perm[T] (conv data-gradient weights:
 * HWIO -> [Cin][tap'][Cout]; perm NULL = identity). */
int rsu_pack_permute(const float* in, void* out, int T, int R, int C, const int* perm_host,
                     void* stream);
/* fp32 -> bf16 cast (transpose-conv forward weights keep TensorFlow's [kh][kw][Cout][Cin]). */
int rsu_cast_bf16(const float* in, void* out, long long n, void* stream);

/* All repacks of one optimizer step in one launch.  rsu_pack_plan validates a host job list,
 * uploads the device table into caller-owned memory (rsu_pack_plan_bytes(n_jobs) bytes; done once,
 * synchronous) and returns the grid size; rsu_pack_run launches it.  kind 0 = rsu_pack_transpose,
 * 1 = rsu_pack_permute with the identity permutation, 2 = rsu_cast_bf16 (n = T*R*C). */
typedef struct {
  const float* in;
  void* out;
  int kind, T, R, C, ld;
} rsu_pack_job;
int rsu_pack_plan_bytes(int n_jobs);
int rsu_pack_plan(const rsu_pack_job* jobs_host, int n_jobs, void* table_dev, int* total_blocks);
int rsu_pack_run(const void* table_dev, int n_jobs, int total_blocks, void* stream);

/* ---------------------------------------------------------------- fused elementwise ------- */
/* color_space_adjust (unet.py:22-23) + optional dropout (unet.py:29-30) + im2col of the 3x3
 * (dilation d) neighbourhood into 64 bf16 channels (k = tap*3 + c for k < 27, channel 27 = 1.0
 * so that the weight-gradient GEMM also yields BiasAddGrad, channels 28..63 = 0):
 *   net0 = (img - 0.5) @ W1 + b1 ; out[n, y, x, tap*3+c] = net0[n, y+oy+dy*d, x+ox+dx*d, c] */
int rsu_color_im2col(const float* img, int N, int S, const float* w1 /* device [3][3] */,
                     const float* b1 /* device [3] */, int dilation, int oy, int ox, int Ho, int Wo,
                     void* out /* bf16 [N,Ho,Wo,64] */, float keep, unsigned long long seed,
                     void* stream);
/* Gradient of the block above w.r.t. W1/b1 given d(im2col) (bf16 [N,Ho,Wo,64]):
 * accumulates dW1 [3][3], db1 [3] (fp32, atomics). */
int rsu_color_im2col_bwd(const float* img, int N, int S, const void* dcol, int dilation, int oy,
                         int ox, int Ho, int Wo, float* dw1, float* db1, float keep,
                         unsigned long long seed, void* stream);

/* First layer without dropout: color_space_adjust (unet.py:22-23) folded into the Cin = 3
 * convolution that follows (unet.py:34, :42).  w: HWIO fp32 [3,3,3,cout], b [cout], w1 [3][3],
 * b1 [3] (device).  Writes the packed bf16 GEMM operand w_packed [cout][64] (k = tap*3 + ci,
 * zero padded) of the folded kernel W' = W1 . W and the folded bias b' = b + sum b1 . W. */
int rsu_first_layer_fold(const float* w, const float* b, const float* w1, const float* b1, int cout,
                         void* w_packed, float* bias_eff, void* stream);
/* Backward of the folded first layer from gx = im2col(x - 0.5)^T dZ (fp32 [>=28][ldg], row
 * tap*3 + ci; row 27 = column sums of dZ, produced by the constant-one column 27 that
 * rsu_color_im2col writes): accumulates dw [3,3,3,cout] (Conv2DBackpropFilter), dbias [cout]
 * (BiasAddGrad), dw1 [3][3] and db1 [3] (gradients of color_space_adjust). */
int rsu_first_layer_grads(const float* gx, int ldg, const float* w, const float* w1, const float* b1,
                          int cout, float* dw, float* dbias, float* dw1, float* db1, void* stream);

/* 2x2/2 max-pool, NHWC bf16 (tf.layers.max_pooling2d, unet.py:52). */
int rsu_maxpool2x2(const void* in, int N, int H, int W, int C, void* out, void* stream);

/* Gradient into a skip tensor Y [N,H,W,C] (ReLU output that feeds both the pool and the decoder
 * crop, unet.py:47-52,70-85):  dZ = [Y>0] * ( maxpool_bwd(dP)  +  pad(dCrop) ).
 * dP [N,H/2,W/2,C] may be NULL; dCrop is a view of the concat gradient (may be NULL) placed at
 * (crop_y, crop_x).  Gradient goes to the first maximum of each window (ties are zeros). */
int rsu_skip_grad(const void* Y, int N, int H, int W, int C, const void* dP, const rsu_view* dCrop,
                  int crop_y, int crop_x, void* dZ, void* stream);

/* dZ = [Y>0] * dY for strided views (ReluGrad). */
int rsu_relu_mask(const rsu_view* Y, const rsu_view* dY, void* dZ, void* stream);

/* Column sums of a bf16 view: out[c] += sum_{n,y,x} v[n,y,x,c]  (BiasAddGrad). */
int rsu_bias_grad(const rsu_view* v, float* out, void* stream);

/* Head (unet.py:95 + tf_aerial_images.py:103-110,147-148): 1x1 conv C->2, softmax, P(road),
 * mean cross-entropy and -- when labels != NULL -- the gradients: dZ (bf16, ReLU-masked grad of
 * the activation), dW [C][2], db [2], loss_sum (fp32 scalar, sum over pixels / (N*H*W)). */
int rsu_head(const void* act, int N, int H, int W, int C, const float* w /*[C][2]*/,
             const float* b /*[2]*/, const unsigned char* labels /* [N,H,W] or NULL */,
             float* probs /* [N,H,W] or NULL */, float* logits /* [N,H,W,2] or NULL */,
             float* loss /* scalar accumulate */, void* dZ, float* dW, float* db, void* stream);

/* Inverted dropout on a bf16 tensor with a counter-based RNG (tf.nn.dropout, unet.py:30,65):
 * y = x / keep * floor(keep + U);  the same (seed, site) regenerates the mask for backward. */
int rsu_dropout(const void* x, void* y, long long n, float keep, unsigned long long seed,
                void* stream);

/* The scale factors (0 or 1/keep) rsu_dropout applies, as fp32 -- lets a test feed the identical
 * mask to the CPU oracle (TensorFlow's own random stream is not reproducible). */
int rsu_dropout_mask(float* m, long long n, float keep, unsigned long long seed, void* stream);

/* Momentum SGD (tf.train.MomentumOptimizer, tf_aerial_images.py:116-121), fp32 master weights:
 *   acc = momentum*acc + g*gscale ; w -= lr*acc */
int rsu_momentum_sgd(float* w, float* acc, const float* g, long long n, float lr, float momentum,
                     float gscale, void* stream);

/* Data-parallel optimizer step over NVLink peer memory (one process per GPU; replaces
 * "all-reduce the flat gradient, then ApplyMomentum on every replica", tf_aerial_images.py:112-122):
 * for the elements [begin, end) of the flat parameter vector -- the calling rank's slice --
 *   g = sum_r grads[r][i] ; acc[i] = momentum*acc[i] + g*gscale ; w = params[rank][i] - lr*acc[i] ;
 *   params[r][i] = w for every rank r.
 * grads[r] / params[r]: device addresses of rank r's flat gradient / parameter buffers as mapped
 * into THIS process (symmetric memory; entry `rank` is the local buffer).  grads_mc / params_mc:
 * NVLS multicast addresses of the same buffers, or NULL (then plain peer loads / stores are used).
 * The caller orders the step across ranks: every rank's gradients are final before the call, and
 * no rank reads its parameters (or zeroes its gradients) before every rank's call has completed. */
#define RSU_MAX_PEERS 8
typedef struct {
  int world, rank;
  const float* grads[RSU_MAX_PEERS];
  float* params[RSU_MAX_PEERS];
  const float* grads_mc;
  float* params_mc;
} rsu_dp_peers;
int rsu_dp_momentum_sgd(const rsu_dp_peers* peers, float* acc, long long begin, long long end, float lr,
                        float momentum, float gscale, void* stream);

/* Zero-fill of a device range on `stream` (gradient and loss accumulators before a step). */
int rsu_fill_zero(void* ptr, long long bytes, void* stream);

/* First-layer (Cin = 3) 3x3 convolution with the im2col operand generated on the fly
 * (color_space_adjust + dropout + conv_0/conv1 or conv_dilut_0/atrous_conv1, src/unet.py:22-23,
 * 29-30, 34-35, 42-43).  img: fp32 [N,S,S,3]; cw / cb: DEVICE fp32 [3][3] / [3] colour transform
 * applied to (x - 0.5), or both NULL = identity (the transform folded into the kernel, keep = 1); the window of the
 * image starting at (oy, ox) is convolved with dilation `dilation`.  w_packed: bf16 [cout][64]
 * (k = tap*3 + c < 27, zero beyond); out: bf16 view [N,Ho,Wo,cout], cout = 64 or 128. */
int rsu_first_conv_fwd(const float* img, int N, int S, const float* cw, const float* cb,
                       int dilation, int oy, int ox, const void* w_packed, const float* bias,
                       int relu, const rsu_view* out, float keep, unsigned long long seed,
                       void* stream);
/* Its weight gradient: dw[k][co] += sum_pixels im2col[pixel][k] * dz[pixel][co] for k < 28 (row 27
 * is the constant-one column = BiasAddGrad); dw fp32 with row stride ldo, accumulated into. */
int rsu_first_conv_wgrad(const float* img, int N, int S, const float* cw, const float* cb,
                         int dilation, int oy, int ox, const rsu_view* dz, float* dw, int ldo,
                         float keep, unsigned long long seed, void* stream);

/* ---------------------------------------------------------------- geometry (images.py) ---- */
/* images.mirror_border (images.py:269-281): np.pad(..., "symmetric") on H and W; fp32. */
int rsu_mirror_pad(const float* in, int N, int H, int W, int C, int pad, float* out, void* stream);
/* One element of the dihedral group per image: op = flip_ud (bit 2) then rot90^k (bits 0-1),
 * counter-clockwise like np.rot90 / tf.image.rot90 (images.py:376-417,
 * tf_aerial_images.py:173-210).  pixel_bytes = bytes per pixel (C * sizeof element); any value. */
int rsu_d4_transform(const void* in, void* out, int N, int S, int pixel_bytes,
                     const unsigned char* ops /* device [N] */,
                     int n_in /* images in `in` (0 = N): output n transforms input n % n_in, so the
                                 6-way ensemble of images.py:376-396 is one launch over N = 6 n_in */,
                     void* stream);
/* images.extract_patches (images.py:35-85): x-outer / y-inner patch order; fp32.  Only patches
 * [k_begin, k_begin + k_count) of the N*side*side patch list are written (k_count < 0 = all that
 * follow k_begin), so a prediction batch never materialises the whole patch tensor. */
int rsu_extract_patches(const float* in, int N, int H, int W, int C, int patch, int stride,
                        long long k_begin, long long k_count, float* out, void* stream);
/* Windows addressed by a device job table (int32 x 4 per job: source image, top row, left column,
 * destination window index): out[dst] = in[img, y0 : y0 + win, x0 : x0 + win, :] with zeros outside
 * the image.  The data movement of the sliding-window loop of predict() (tf_aerial_images.py:296-315)
 * when aligned windows are evaluated together: enlarged input windows in, patch outputs out. */
int rsu_copy_windows(const float* in, int N, int H, int W, int C, int win, long long n_jobs,
                     const int* jobs_dev, float* out /* [*, win, win, C] */, void* stream);
/* In place: sums[n, y, x, c] /= hit count of pixel (y, x) under side x side windows of size P at
 * `stride` (the count_hits division of images.py:162, for partial sums reduced across ranks). */
int rsu_divide_by_hits(float* sums, int N, int S, int C, int side, int P, int stride, void* stream);
/* images.images_from_patches (images.py:131-164): overlap average in gather form (fp64 sums in
 * a fixed order, no atomics; 16-byte lanes when P*C and stride*C are multiples of 4 and both
 * pointers are 16-byte aligned).  `patches` points at patch k_begin of the list; with
 * normalize = 0 the un-divided partial sum of that slice is written (sharded prediction). */
int rsu_overlap_average(const float* patches, int N, int side, int P, int C, int stride,
                        long long k_begin, long long k_count, int normalize, float* out,
                        void* stream);
/* crop_imgs(rotate_imgs(x, angle), crop) (images.py:313-373): nearest-neighbour affine resampling
 * with scipy.ndimage.rotate(order=0, mode='constant', cval=0) semantics.  The caller passes
 * SciPy's geometry -- the 2x2 matrix [[c, s], [-s, c]] and the offset in_center - R*out_center
 * (host arrays) -- and the window [crop0, crop0+crop)^2 of the rotated image to produce. */
int rsu_rotate_nn_crop(const float* in, int N, int H, int C, const double* matrix_host,
                       const double* offset_host, int crop0, int crop, float* out, void* stream);
/* invert_image_augmentation_ensemble (images.py:399-417): average of the 6 un-transformed masks. */
int rsu_ensemble_invert(const float* masks /* [6N,S,S] */, int N, int S, float* out, void* stream);

/* images.quantize_mask (images.py:256-266) and the patch labels of save_submission_csv
 * (images.py:206-237 -> extract_patches + labels_for_patches, images.py:88-99) over the
 * patch x patch cells of N square single-channel masks (fp32 or fp64: elem_bytes 4 / 8).
 * rule 0: label = mean(v >= pixel_threshold) > vote_threshold; rule 1: label = mean(v) >
 * vote_threshold (fp64 means).  quantized (nullable, dtype of masks): the label written to every
 * pixel of its cell; labels (nullable): [N][cells along x][cells along y] -- x-outer, the row
 * order of submission.csv.  Edge cells are clipped when patch does not divide S. */
int rsu_patch_vote(const void* masks, int elem_bytes, int N, int S, int patch, int rule,
                   double pixel_threshold, double vote_threshold, void* quantized,
                   unsigned char* labels, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RSU_B200_H */
