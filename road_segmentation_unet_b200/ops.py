"""Operator-level host wrappers over the C ABI (one function per TensorFlow op of the reference).

Every function takes NHWC bf16 torch CUDA tensors purely as device buffers and launches
hand-written sm_100a kernels from librsu_b200.so on the current stream.  Weight layouts:

  HWIO fp32 master (TensorFlow's [kh, kw, Cin, Cout], what the optimizer updates)
    -> forward pack   [Cout][tap][Cin]   bf16   (pack_conv_fwd)
    -> dgrad pack     [Cin][tap][Cout]   bf16   (pack_conv_dgrad)
  transpose-conv master [2, 2, Cout, Cin] (TensorFlow's conv2d_transpose kernel)
    -> forward pack   [(a,b,co)][ci]     bf16   (cast only)
    -> dgrad pack     [ci][(a,b,co)]     bf16   (transpose)
"""
import ctypes as C

import torch

from . import _lib
from ._lib import ConvGemmDesc, View, WgradDesc, call, view


# ------------------------------------------------------------------ optional kernel timing
# bench.py brackets the two tcgen05 kernels with CUDA events on the launching stream to report
# the roofline of the dominant kernel; the engine tags every layer with its algorithmic FLOPs.
_PROFILE = None
_LAYER = (None, 0.0)


def set_layer(name, flops):
    global _LAYER
    _LAYER = (name, float(flops))


def profile_start():
    global _PROFILE
    _PROFILE = []


def profile_stop():
    """Returns [(kind, layer, alg_flops, milliseconds)] and disables profiling."""
    global _PROFILE
    rec, _PROFILE = _PROFILE or [], None
    torch.cuda.synchronize()
    return [(k, n, f, e0.elapsed_time(e1)) for (k, n, f, e0, e1) in rec]


def _timed(kind, fn, *args):
    if _PROFILE is None:
        return fn(*args)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    fn(*args)
    e1.record()
    _PROFILE.append((kind, _LAYER[0], _LAYER[1], e0, e1))


def _taps(desc, taps):
    desc.n_taps = len(taps)
    for i, (dy, dx) in enumerate(taps):
        desc.tap_dy[i] = dy
        desc.tap_dx[i] = dx


def conv_taps(dilation=1, sign=1):
    return [(sign * ky * dilation, sign * kx * dilation) for ky in range(3) for kx in range(3)]


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


ALGO_AUTO, ALGO_PER_TAP, ALGO_HALO, ALGO_PER_TAP_PAIR, ALGO_HALO_PAIR = 0, 1, 2, 3, 4


def conv_gemm(srcs, taps, weights, out, n_out, bias=None, relu=False, mask=None, accumulate=False,
              shuffle_cout=0, grid_hw=None, algo=ALGO_AUTO, mask_c0=0, pool_out=None):
    """Generic implicit GEMM (rsu_conv_gemm).  srcs: list of (tensor_or_View, off_y, off_x).
    algo: 0 = library's choice, 1 = one TMA box per tap, 2 = halo tile shared by all taps."""
    d = ConvGemmDesc()
    d.n_src = len(srcs)
    for i, (t, oy, ox) in enumerate(srcs):
        v = t if isinstance(t, View) else view(t)
        v.off_y, v.off_x = oy, ox
        d.src[i] = v
    _taps(d, taps)
    d.weights = _ptr(weights)
    d.Ntot = n_out
    n, h, w, _ = out.shape
    if grid_hw is None:
        grid_hw = (h // 2, w // 2) if shuffle_cout else (h, w)
    d.H_out, d.W_out, d.N_img = grid_hw[0], grid_hw[1], n
    d.out = _ptr(out)
    d.out_sn, d.out_sy, d.out_sx = out.stride()[:3]
    d.shuffle_cout = shuffle_cout
    d.bias = _ptr(bias)
    d.relu = int(relu)
    if mask is not None:
        # a mask with fewer channels than `out` covers output channels [mask_c0, mask_c0 + C_mask)
        assert mask.shape[:3] == out.shape[:3] and mask.shape[3] <= out.shape[3]
        d.mask = _ptr(mask)
        d.mask_sn, d.mask_sy, d.mask_sx = mask.stride()[:3]
        if mask.shape[3] != out.shape[3]:
            d.mask_c0, d.mask_nc = int(mask_c0), int(mask.shape[3])
    d.accumulate = int(accumulate)
    d.algo = int(algo)
    pooled = C.c_int(0)
    if pool_out is not None:
        assert pool_out.shape == (n, h // 2, w // 2, out.shape[3])
        d.pool_out = _ptr(pool_out)
        d.pool_sn, d.pool_sy, d.pool_sx = pool_out.stride()[:3]
        d.pool_done_host = C.pointer(pooled)
    _timed("conv_gemm", call, "rsu_conv_gemm", C.byref(d))
    return bool(pooled.value)


def wgrad_gemm(srcs, taps, grad, grad_off, out, grid_hw, bias_grad=None, algo=ALGO_AUTO):
    """Generic weight-gradient GEMM (rsu_wgrad_gemm); out: fp32 [rows, Cout], pre-zeroed.
    bias_grad (fp32 [Cout], accumulated into) is produced by the halo-tile kernel for free; the
    return value tells whether it was (False: the caller still has to run bias_grad())."""
    d = WgradDesc()
    d.n_src = len(srcs)
    for i, (t, oy, ox) in enumerate(srcs):
        v = t if isinstance(t, View) else view(t)
        v.off_y, v.off_x = oy, ox
        d.src[i] = v
    _taps(d, taps)
    g = grad if isinstance(grad, View) else view(grad)
    g.off_y, g.off_x = grad_off
    d.grad = g
    d.H, d.W = grid_hw
    d.N_img = g.N
    d.out = _ptr(out)
    d.ldo = out.stride(0)
    d.bias_grad = _ptr(bias_grad)
    done = C.c_int(0)
    d.bias_done_host = C.pointer(done)
    d.algo = int(algo)
    _timed("wgrad_gemm", call, "rsu_wgrad_gemm", C.byref(d))
    return bool(done.value)


# ------------------------------------------------------------------ weight packing
def pack_conv_fwd(w_hwio, out, taps, cin, cout, ld=0):
    """HWIO fp32 [taps, cin, cout] -> bf16 [cout][taps*cin] (row length ld)."""
    call("rsu_pack_transpose", _ptr(w_hwio), _ptr(out), taps, cin, cout, ld)


def pack_conv_dgrad(w_hwio, out, taps, cin, cout):
    """HWIO fp32 [taps, cin, cout] -> bf16 [cin][taps][cout]."""
    call("rsu_pack_permute", _ptr(w_hwio), _ptr(out), taps, cin, cout, None)


def cast_bf16(src, out):
    call("rsu_cast_bf16", _ptr(src), _ptr(out), src.numel())


PACK_TRANSPOSE, PACK_PERMUTE, PACK_CAST = 0, 1, 2


class PackPlan:
    """A fixed list of weight repacks (fp32 master -> bf16 operand layouts) run as one launch.
    jobs: [(kind, fp32 source tensor, bf16 destination tensor, T, R, C, ld)]."""

    def __init__(self, jobs, device):
        lib = _lib.load()
        n = len(jobs)
        arr = (_lib.PackJob * n)()
        for i, (kind, src, dst, T, R, Cc, ld) in enumerate(jobs):
            arr[i] = _lib.PackJob(src.data_ptr(), dst.data_ptr(), kind, T, R, Cc, ld)
        self._keep = [(src, dst) for (_, src, dst, _, _, _, _) in jobs]  # keep buffers alive
        self.table = torch.empty(lib.rsu_pack_plan_bytes(n), dtype=torch.uint8, device=device)
        blocks = C.c_int(0)
        _lib.check(lib.rsu_pack_plan(arr, n, C.c_void_p(self.table.data_ptr()), C.byref(blocks)))
        self.n, self.blocks = n, blocks.value

    def run(self):
        call("rsu_pack_run", _ptr(self.table), self.n, self.blocks)


# ------------------------------------------------------------------ conv 3x3 (unet.py:34-45,88-91)
def conv3x3_fwd(srcs, w_fwd, bias, out, dilation=1, relu=True, algo=ALGO_AUTO, pool_out=None):
    """srcs: [(tensor, off_y, off_x)] in concat order; out [N,Ho,Wo,Cout] bf16.  pool_out: the 2x2
    max pool of `out`, written by the same kernel when it can (returns True), else untouched."""
    return conv_gemm(srcs, conv_taps(dilation), w_fwd, out, out.shape[3], bias=bias, relu=relu, algo=algo,
                     pool_out=pool_out)


def conv3x3_dgrad(dz, w_dgrad, dx_window, dilation=1, mask=None, accumulate=False, algo=ALGO_AUTO,
                  mask_c0=0):
    """dx_window[u,v] = sum_taps dz[u - ky*d, v - kx*d] W[ky,kx]^T over the touched input window
    (extent = dz extent + 2*dilation); optional fused ReLU mask / accumulation."""
    conv_gemm([(dz, 0, 0)], conv_taps(dilation, -1), w_dgrad, dx_window, dx_window.shape[3],
              mask=mask, accumulate=accumulate, algo=algo, mask_c0=mask_c0)


def conv3x3_wgrad(srcs, dz, dw, dilation=1, bias_grad=None, algo=ALGO_AUTO):
    """dw fp32 [9*Cin_total, Cout] (HWIO), pre-zeroed; srcs as in conv3x3_fwd.  Returns True when
    bias_grad was filled by the same kernel."""
    return wgrad_gemm(srcs, conv_taps(dilation), dz, (0, 0), dw, (dz.shape[1], dz.shape[2]),
                      bias_grad=bias_grad, algo=algo)


# ------------------------------------------------------------------ conv2d_transpose (unet.py:67)
def _quad_views(t):
    """The four stride-2 phase views (a, b) of an NHWC tensor [N, 2H, 2W, C]."""
    n, h2, w2, c = t.shape
    sn, sy, sx, _ = t.stride()
    out = []
    for a in range(2):
        for b in range(2):
            out.append(View(C.c_void_p(t.data_ptr() + 2 * (a * sy + b * sx)), c, h2 // 2, w2 // 2,
                            n, sn, 2 * sy, 2 * sx, 0, 0))
    return out


def upconv2x2_fwd(x, w_fwd, bias, out):
    cout = out.shape[3]
    conv_gemm([(x, 0, 0)], [(0, 0)], w_fwd, out, 4 * cout, bias=bias, shuffle_cout=cout)


def upconv2x2_dgrad(dy, w_dgrad, dx, mask=None):
    conv_gemm([(v, 0, 0) for v in _quad_views(dy)], [(0, 0)], w_dgrad, dx, dx.shape[3], mask=mask)


def upconv2x2_wgrad(dy, x, dw):
    """dw fp32 [4*Cout, Cin] (TensorFlow's [2,2,Cout,Cin]), pre-zeroed."""
    wgrad_gemm([(v, 0, 0) for v in _quad_views(dy)], [(0, 0)], x, (0, 0), dw,
               (x.shape[1], x.shape[2]))


# ------------------------------------------------------------------ elementwise
def maxpool2x2(x, out):
    n, h, w, c = x.shape
    call("rsu_maxpool2x2", _ptr(x), n, h, w, c, _ptr(out))


def skip_grad(y, dpool, dcrop, crop_yx, dz):
    n, h, w, c = y.shape
    cv = None
    if dcrop is not None:
        cv = dcrop if isinstance(dcrop, View) else view(dcrop)
    call("rsu_skip_grad", _ptr(y), n, h, w, c, _ptr(dpool), C.byref(cv) if cv is not None else None,
         crop_yx[0], crop_yx[1], _ptr(dz))


def relu_mask(y, dy, dz):
    yv = y if isinstance(y, View) else view(y)
    dv = dy if isinstance(dy, View) else view(dy)
    call("rsu_relu_mask", C.byref(yv), C.byref(dv), _ptr(dz))


def bias_grad(v, out):
    vv = v if isinstance(v, View) else view(v)
    call("rsu_bias_grad", C.byref(vv), _ptr(out))


def head(act, w, b, labels=None, probs=None, logits=None, loss=None, dz=None, dw=None, db=None):
    n, h, wd, c = act.shape
    call("rsu_head", _ptr(act), n, h, wd, c, _ptr(w), _ptr(b), _ptr(labels), _ptr(probs),
         _ptr(logits), _ptr(loss), _ptr(dz), _ptr(dw), _ptr(db))


def dropout(x, y, keep, seed):
    call("rsu_dropout", _ptr(x), _ptr(y), x.numel(), float(keep), int(seed))


def dropout_mask(n, keep, seed, device="cuda"):
    m = torch.empty(n, dtype=torch.float32, device=device)
    call("rsu_dropout_mask", _ptr(m), n, float(keep), int(seed))
    return m


def fill_zero(t):
    """Zero a contiguous device tensor with a memset on the current stream (rsu_fill_zero)."""
    assert t.is_contiguous()
    call("rsu_fill_zero", _ptr(t), t.numel() * t.element_size())


def momentum_sgd(w, acc, g, lr, momentum, gscale=1.0):
    call("rsu_momentum_sgd", _ptr(w), _ptr(acc), _ptr(g), w.numel(), float(lr), float(momentum),
         float(gscale))


def color_im2col(img, w1, b1, dilation, oy, ox, out, keep=1.0, seed=0):
    n, s = img.shape[0], img.shape[1]
    call("rsu_color_im2col", _ptr(img), n, s, _ptr(w1), _ptr(b1), dilation, oy, ox, out.shape[1],
         out.shape[2], _ptr(out), float(keep), int(seed))


def color_im2col_bwd(img, dcol, dilation, oy, ox, dw1, db1, keep=1.0, seed=0):
    n, s = img.shape[0], img.shape[1]
    call("rsu_color_im2col_bwd", _ptr(img), n, s, _ptr(dcol), dilation, oy, ox, dcol.shape[1],
         dcol.shape[2], _ptr(dw1), _ptr(db1), float(keep), int(seed))


def first_conv_fwd(img, cw, cb, dilation, oy, ox, w_packed, bias, out, relu=True, keep=1.0, seed=0):
    """Cin = 3 convolution with the im2col operand built on the fly (rsu_first_conv_fwd): the
    window of img [N,S,S,3] fp32 at (oy, ox) -> out [N,Ho,Wo,cout] bf16, cout = 64 or 128.
    cw / cb: device fp32 colour transform applied to (x - 0.5)."""
    n, s = img.shape[0], img.shape[1]
    ov = view(out)

    def run(*a):
        call("rsu_first_conv_fwd", *a)
    _timed("conv_gemm", run, _ptr(img), n, s, _ptr(cw), _ptr(cb), int(dilation), int(oy), int(ox),
           _ptr(w_packed), _ptr(bias), int(relu), C.byref(ov), float(keep), int(seed))


def first_conv_wgrad(img, cw, cb, dilation, oy, ox, dz, dw, keep=1.0, seed=0):
    """dw[k, co] += im2col(img)^T dz for k < 28 (row 27 = BiasAddGrad); dw fp32 [>= 28, cout]."""
    n, s = img.shape[0], img.shape[1]
    gv = view(dz)

    def run(*a):
        call("rsu_first_conv_wgrad", *a)
    _timed("wgrad_gemm", run, _ptr(img), n, s, _ptr(cw), _ptr(cb), int(dilation), int(oy), int(ox),
           C.byref(gv), _ptr(dw), int(dw.stride(0)), float(keep), int(seed))


def first_layer_fold(w, b, w1, b1, w_packed, bias_eff):
    """Fold color_space_adjust into a Cin = 3 convolution (w: HWIO fp32 [3,3,3,cout])."""
    call("rsu_first_layer_fold", _ptr(w), _ptr(b), _ptr(w1), _ptr(b1), w.shape[3], _ptr(w_packed),
         _ptr(bias_eff))


def first_layer_grads(gx, w, w1, b1, dw, dbias, dw1, db1):
    """Gradients of the folded first layer from gx = im2col(x - 0.5)^T dZ (fp32 [rows, cout];
    row 27 = BiasAddGrad through the constant-one im2col column)."""
    call("rsu_first_layer_grads", _ptr(gx), gx.stride(0), _ptr(w), _ptr(w1), _ptr(b1),
         w.shape[3], _ptr(dw), _ptr(dbias), _ptr(dw1), _ptr(db1))


def launch_count():
    return _lib.launch_count()
