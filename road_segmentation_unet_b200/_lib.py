"""ctypes binding of librsu_b200.so (include/rsu_b200.h).

PyTorch tensors are used only as device buffers: every wrapper takes `tensor.data_ptr()` and the
current CUDA stream.  There is no CPU fallback -- if the library is missing or a call fails this
module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librsu_b200.so")

RSU_MAX_SRC = 4
RSU_MAX_TAPS = 9


class View(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("C", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("N", C.c_int), ("sn", C.c_longlong), ("sy", C.c_longlong), ("sx", C.c_longlong),
                ("off_y", C.c_int), ("off_x", C.c_int)]


class ConvGemmDesc(C.Structure):
    _fields_ = [("n_src", C.c_int), ("src", View * RSU_MAX_SRC), ("n_taps", C.c_int),
                ("tap_dy", C.c_int * RSU_MAX_TAPS), ("tap_dx", C.c_int * RSU_MAX_TAPS),
                ("weights", C.c_void_p), ("Ntot", C.c_int), ("H_out", C.c_int),
                ("W_out", C.c_int), ("N_img", C.c_int), ("out", C.c_void_p),
                ("out_sn", C.c_longlong), ("out_sy", C.c_longlong), ("out_sx", C.c_longlong),
                ("shuffle_cout", C.c_int), ("bias", C.c_void_p), ("relu", C.c_int),
                ("mask", C.c_void_p), ("mask_sn", C.c_longlong), ("mask_sy", C.c_longlong),
                ("mask_sx", C.c_longlong), ("accumulate", C.c_int), ("algo", C.c_int),
                ("mask_c0", C.c_int), ("mask_nc", C.c_int), ("pool_out", C.c_void_p),
                ("pool_sn", C.c_longlong), ("pool_sy", C.c_longlong), ("pool_sx", C.c_longlong),
                ("pool_done_host", C.POINTER(C.c_int))]


class WgradDesc(C.Structure):
    _fields_ = [("n_src", C.c_int), ("src", View * RSU_MAX_SRC), ("n_taps", C.c_int),
                ("tap_dy", C.c_int * RSU_MAX_TAPS), ("tap_dx", C.c_int * RSU_MAX_TAPS),
                ("grad", View), ("H", C.c_int), ("W", C.c_int), ("N_img", C.c_int),
                ("out", C.c_void_p), ("ldo", C.c_int), ("bias_grad", C.c_void_p),
                ("bias_done_host", C.POINTER(C.c_int)), ("algo", C.c_int)]


class PackJob(C.Structure):
    _fields_ = [("inp", C.c_void_p), ("out", C.c_void_p), ("kind", C.c_int), ("T", C.c_int),
                ("R", C.c_int), ("C", C.c_int), ("ld", C.c_int)]


RSU_MAX_PEERS = 8


class DpPeers(C.Structure):
    _fields_ = [("world", C.c_int), ("rank", C.c_int), ("grads", C.c_void_p * RSU_MAX_PEERS),
                ("params", C.c_void_p * RSU_MAX_PEERS), ("grads_mc", C.c_void_p), ("params_mc", C.c_void_p)]


class RsuError(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes)
_vp, _i, _ll, _f, _d, _ull = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double, C.c_ulonglong
_SIGNATURES = {
    "rsu_last_error": (C.c_char_p, []),
    "rsu_version": (_i, []),
    "rsu_launch_count": (_ll, []),
    "rsu_reset_launch_count": (None, []),
    "rsu_crc32c_host": (C.c_uint, [C.c_uint, _vp, _ull]),
    "rsu_launch_histogram": (_i, [C.c_char_p, _i]),
    "rsu_conv_gemm": (_i, [C.POINTER(ConvGemmDesc), _vp]),
    "rsu_wgrad_gemm": (_i, [C.POINTER(WgradDesc), _vp]),
    "rsu_pack_transpose": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "rsu_pack_permute": (_i, [_vp, _vp, _i, _i, _i, C.POINTER(C.c_int), _vp]),
    "rsu_cast_bf16": (_i, [_vp, _vp, _ll, _vp]),
    "rsu_color_im2col": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _f, _ull, _vp]),
    "rsu_color_im2col_bwd": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _f, _ull, _vp]),
    "rsu_first_conv_fwd": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _i, C.POINTER(View), _f, _ull, _vp]),
    "rsu_first_conv_wgrad": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _i, C.POINTER(View), _vp, _i, _f, _ull, _vp]),
    "rsu_pack_plan_bytes": (_i, [_i]),
    "rsu_pack_plan": (_i, [C.POINTER(PackJob), _i, _vp, C.POINTER(C.c_int)]),
    "rsu_pack_run": (_i, [_vp, _i, _i, _vp]),
    "rsu_first_layer_fold": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "rsu_first_layer_grads": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "rsu_maxpool2x2": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "rsu_skip_grad": (_i, [_vp, _i, _i, _i, _i, _vp, C.POINTER(View), _i, _i, _vp, _vp]),
    "rsu_relu_mask": (_i, [C.POINTER(View), C.POINTER(View), _vp, _vp]),
    "rsu_bias_grad": (_i, [C.POINTER(View), _vp, _vp]),
    "rsu_head": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rsu_dropout": (_i, [_vp, _vp, _ll, _f, _ull, _vp]),
    "rsu_dropout_mask": (_i, [_vp, _ll, _f, _ull, _vp]),
    "rsu_momentum_sgd": (_i, [_vp, _vp, _vp, _ll, _f, _f, _f, _vp]),
    "rsu_dp_momentum_sgd": (_i, [C.POINTER(DpPeers), _vp, _ll, _ll, _f, _f, _f, _vp]),
    "rsu_fill_zero": (_i, [_vp, _ll, _vp]),
    "rsu_mirror_pad": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "rsu_d4_transform": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp]),
    "rsu_copy_windows": (_i, [_vp, _i, _i, _i, _i, _i, _ll, _vp, _vp, _vp]),
    "rsu_divide_by_hits": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp]),
    "rsu_extract_patches": (_i, [_vp, _i, _i, _i, _i, _i, _i, _ll, _ll, _vp, _vp]),
    "rsu_overlap_average": (_i, [_vp, _i, _i, _i, _i, _i, _ll, _ll, _i, _vp, _vp]),
    "rsu_rotate_nn_crop": (_i, [_vp, _i, _i, _i, C.POINTER(C.c_double), C.POINTER(C.c_double), _i, _i, _vp, _vp]),
    "rsu_ensemble_invert": (_i, [_vp, _i, _i, _vp, _vp]),
    "rsu_patch_vote": (_i, [_vp, _i, _i, _i, _i, _i, C.c_double, C.c_double, _vp, _vp, _vp]),
}
EXPORTS = sorted(_SIGNATURES)


def load():
    """dlopen the library (building it first if the .so is absent and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    # (re)build when the sources changed since the .so was linked: a cheap digest check, so a stale
    # library is never used silently after an edit of csrc/ (needs nvcc only when it has to build)
    from . import build as _build
    try:
        _build.build()
    except Exception:
        if not os.path.exists(LIB_PATH):
            raise
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RsuError("rsu error %d: %s" % (rc, load().rsu_last_error().decode()))


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def call(name, *args):
    """Call an int-returning entry point on the current torch stream and raise on failure."""
    lib = load()
    check(getattr(lib, name)(*args, stream_ptr()))


def view(t, C_=None, c0=0, off_y=0, off_x=0):
    """rsu_view of an NHWC bf16 torch tensor (optionally a channel slice [c0, c0+C_))."""
    n, h, w, c = t.shape
    sn, sy, sx, sc = t.stride()
    assert sc == 1, "channels must be contiguous"
    cc = c if C_ is None else C_
    return View(C.c_void_p(t.data_ptr() + 2 * c0), cc, h, w, n, sn, sy, sx, off_y, off_x)


def launch_count():
    return load().rsu_launch_count()


def launch_histogram():
    """{kernel name: launches since the last reset}"""
    lib = load()
    buf = C.create_string_buffer(8192)
    lib.rsu_launch_histogram(buf, len(buf))
    out = {}
    for item in buf.value.decode().split(";"):
        if "=" in item:
            k, v = item.split("=")
            out[k] = int(v)
    return out
