"""CPU oracle for the TensorFlow half of the hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module; the product path (road_segmentation_unet_b200/) never does.

PARITY UNPINNED for this half: the arithmetic of the reference lives in tensorflow==1.4.0
(requirements.txt:18-19), which is not vendored under /root/reference, is not installable here
(no wheel, no network) and has no golden vectors or value-level tests in the reference
(src/test_images.py pins shapes of two NumPy helpers only).  This file therefore restates the
published semantics of the TF-1.4 ops at the reference's own call sites:

  unet.forward                       src/unet.py:12-97
  unet.input_size_needed             src/unet.py:100-115
  cross_entropy_loss / probabilities src/tf_aerial_images.py:103-110, 147-148
  optimize (momentum + lr decay)     src/tf_aerial_images.py:112-122
  stochastic_images_augmentation     src/tf_aerial_images.py:173-210

fp32 (or fp64) on CPU with torch ops as the arithmetic back end; `conv2d_valid_loops` is an
independent NumPy loop nest used by the tests to pin the torch-based functions on tiny shapes.
Tensors are NHWC, kernels HWIO ([kh, kw, Cin, Cout]) and transpose-conv kernels [kh, kw, Cout,
Cin] exactly as TensorFlow stores them, under TensorFlow's variable names.
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- sizes
def input_size_needed(output_size, num_layers):
    """src/unet.py:100-115 (same float arithmetic and the same assertion)."""
    for i in range(num_layers - 1):
        assert output_size % 2 == 0, 'expand layer {} has size {} not divisible by 2' \
            .format(num_layers - i, output_size)
        output_size = (output_size + 4) / 2
    for i in range(num_layers - 1):
        output_size = (output_size + 4) * 2
    return int(output_size + 4)


# ----------------------------------------------------------------------------- variables
def variable_shapes(num_layers, root_size, dilated_layers):
    """Variables in TensorFlow creation order (src/unet.py:23-95): name -> shape."""
    shapes = OrderedDict()
    shapes["color_space_adjust/kernel"] = (1, 1, 3, 3)
    shapes["color_space_adjust/bias"] = (3,)
    cin, f = 3, root_size
    for i in range(num_layers):
        if dilated_layers:
            shapes["conv_dilut_%d/atrous_conv1/kernel" % i] = (3, 3, cin, f)
            shapes["conv_dilut_%d/atrous_conv1/bias" % i] = (f,)
            shapes["conv_dilut_%d/atrous_conv2/kernel" % i] = (3, 3, f, f)
            shapes["conv_dilut_%d/atrous_conv2/bias" % i] = (f,)
        shapes["conv_%d/conv1/kernel" % i] = (3, 3, cin, f)
        shapes["conv_%d/conv1/bias" % i] = (f,)
        shapes["conv_%d/conv2/kernel" % i] = (3, 3, f, f)
        shapes["conv_%d/conv2/bias" % i] = (f,)
        cin, f = f, f * 2
    f = f // 2  # filters of the bottom block
    net_c = f
    for i in range(num_layers - 1):
        f = f // 2
        shapes["up_conv_%d/kernel" % i] = (2, 2, f, net_c)  # [kh, kw, Cout, Cin]
        shapes["up_conv_%d/bias" % i] = (f,)
        cat = f * (3 if dilated_layers else 2)
        j = num_layers + i
        shapes["conv_%d/conv1/kernel" % j] = (3, 3, cat, f)
        shapes["conv_%d/conv1/bias" % j] = (f,)
        shapes["conv_%d/conv2/kernel" % j] = (3, 3, f, f)
        shapes["conv_%d/conv2/bias" % j] = (f,)
        net_c = f
    shapes["weight_output/kernel"] = (1, 1, net_c, 2)
    shapes["weight_output/bias"] = (2,)
    return shapes


def dead_variables(num_layers, dilated_layers):
    """The deepest dilated pair is built but discarded (src/unet.py:56-59)."""
    if not dilated_layers:
        return []
    i = num_layers - 1
    return ["conv_dilut_%d/atrous_conv%d/%s" % (i, k, p) for k in (1, 2) for p in ("kernel", "bias")]


def init_params(num_layers, root_size, dilated_layers, seed=2017, dtype=np.float32):
    """Glorot-uniform kernels (tf.layers default), zero biases, drawn in creation order."""
    rs = np.random.RandomState(seed)
    params = OrderedDict()
    for name, shape in variable_shapes(num_layers, root_size, dilated_layers).items():
        if name.endswith("bias"):
            params[name] = np.zeros(shape, dtype=dtype)
        else:
            kh, kw, a, b = shape
            limit = np.sqrt(6.0 / (kh * kw * a + kh * kw * b))
            params[name] = rs.uniform(-limit, limit, size=shape).astype(dtype)
    return params


# ----------------------------------------------------------------------------- ops
def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1)


def conv2d_valid(x, w_hwio, b, dilation=1):
    """tf.layers.conv2d(padding='valid', dilation_rate=d): cross-correlation, NHWC in/out."""
    y = F.conv2d(_nchw(x), w_hwio.permute(3, 2, 0, 1), b, dilation=dilation)
    return _nhwc(y)


def conv2d_transpose_2x2(x, w, b):
    """tf.layers.conv2d_transpose(k=2, s=2, 'valid'); kernel [kh, kw, Cout, Cin]:
    y[2i+a, 2j+b, co] = sum_ci x[i, j, ci] * w[a, b, co, ci] + bias[co]."""
    y = F.conv_transpose2d(_nchw(x), w.permute(3, 2, 0, 1), b, stride=2)
    return _nhwc(y)


def max_pool_2x2(x):
    return _nhwc(F.max_pool2d(_nchw(x), 2, 2))


def center_crop(x, size):
    """tf.image.resize_image_with_crop_or_pad when cropping: offset floor((H - T) / 2)."""
    h, w = x.shape[1], x.shape[2]
    oy, ox = (h - size) // 2, (w - size) // 2
    return x[:, oy:oy + size, ox:ox + size, :]


def conv2d_valid_loops(x, w, b, dilation=1):
    """Independent NumPy loop nest for tiny shapes (pins conv2d_valid)."""
    n, h, wd, cin = x.shape
    kh, kw, _, cout = w.shape
    ho, wo = h - (kh - 1) * dilation, wd - (kw - 1) * dilation
    y = np.zeros((n, ho, wo, cout), dtype=np.float64)
    for ky in range(kh):
        for kx in range(kw):
            patch = x[:, ky * dilation:ky * dilation + ho, kx * dilation:kx * dilation + wo, :]
            y += np.einsum("nhwc,co->nhwo", patch.astype(np.float64), w[ky, kx].astype(np.float64))
    return y + b.astype(np.float64)


def conv2d_transpose_2x2_loops(x, w, b):
    n, h, wd, cin = x.shape
    cout = w.shape[2]
    y = np.zeros((n, 2 * h, 2 * wd, cout), dtype=np.float64)
    for a in range(2):
        for bb in range(2):
            y[:, a::2, bb::2, :] = np.einsum("nhwc,oc->nhwo", x.astype(np.float64),
                                             w[a, bb].astype(np.float64))
    return y + b.astype(np.float64)


class _RoundBF16(torch.autograd.Function):
    """Round-to-nearest-even to bfloat16 in the forward AND in the backward direction."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().to(g.dtype)


def _with_bf16_storage(conv_fn):
    """Storage-precision model of the device path: operands and results of every tensor-core
    convolution are bf16 in memory (fp32 accumulate), forward and backward.  The 1x1
    color_space_adjust and weight_output layers are evaluated in fp32 on the device too."""
    def wrapped(x, w, b, *a):
        if w.shape[0] == 1 and w.shape[1] == 1:
            return conv_fn(x, w, b, *a)
        r = _RoundBF16.apply
        return r(conv_fn(r(x), r(w), b, *a))
    return wrapped


# ----------------------------------------------------------------------------- forward
def forward(X, params, num_layers, root_size, dilated_layers, dropout_scales=None, keep=None,
            acts=None, storage=None):
    """src/unet.py:12-97.  X: torch [B,S,S,3]; params: name -> torch tensor.

    dropout_scales: optional list of per-site multiplicative masks (0 or 1/keep), in the order
    the reference calls tf.nn.dropout (one per encoder block, one per decoder block); TF's own
    random stream is not reproducible so masks are inputs.  acts (dict) collects activations.
    storage: None = the reference's fp32 everywhere; "bf16" = same algorithm with activations,
    weights and gradients of the 3x3 / transpose convolutions rounded to bf16 where the device
    path stores bf16 (used to separate kernel errors from the precision floor).
    """
    conv2d_valid = globals()["conv2d_valid"]
    conv2d_transpose_2x2 = globals()["conv2d_transpose_2x2"]
    if storage == "bf16":
        conv2d_valid = _with_bf16_storage(conv2d_valid)
        conv2d_transpose_2x2 = _with_bf16_storage(conv2d_transpose_2x2)
    else:
        assert storage is None

    def rec(name, t):
        if acts is not None:
            acts[name] = t
        return t

    site = [0]

    def drop(t):
        if dropout_scales is None:
            return t
        m = dropout_scales[site[0]]
        site[0] += 1
        return t * m

    net = X - 0.5                                                             # unet.py:22
    net = conv2d_valid(net, params["color_space_adjust/kernel"], params["color_space_adjust/bias"])
    rec("color_space_adjust", net)
    conv = []
    for i in range(num_layers):                                               # unet.py:28-54
        net = drop(net)
        dil = None
        if dilated_layers and i < num_layers - 1:
            # the pair at i = L-1 is dead code (unet.py:56-59) and is not evaluated
            p = "conv_dilut_%d/" % i
            dil = torch.relu(conv2d_valid(net, params[p + "atrous_conv1/kernel"],
                                          params[p + "atrous_conv1/bias"], 2))
            rec(p + "relu1", dil)
            dil = torch.relu(conv2d_valid(dil, params[p + "atrous_conv2/kernel"],
                                          params[p + "atrous_conv2/bias"], 2))
            rec(p + "relu2", dil)
        p = "conv_%d/" % i
        net = rec(p + "relu1", torch.relu(conv2d_valid(net, params[p + "conv1/kernel"],
                                                       params[p + "conv1/bias"])))
        net = rec(p + "relu2", torch.relu(conv2d_valid(net, params[p + "conv2/kernel"],
                                                       params[p + "conv2/bias"])))
        conv.append((net, dil))
        if i < num_layers - 1:
            net = rec("pool_%d" % i, max_pool_2x2(net))                       # last pool is dead
    net = conv.pop()[0]                                                       # unet.py:56-59
    for i in range(num_layers - 1):                                           # unet.py:61-91
        net = drop(net)
        net = rec("up_conv_%d" % i, conv2d_transpose_2x2(net, params["up_conv_%d/kernel" % i],
                                                         params["up_conv_%d/bias" % i]))
        trav, trav_dil = conv.pop()
        size = net.shape[1]
        parts = [center_crop(trav, size)]
        if dilated_layers:
            parts.append(center_crop(trav_dil, size))
        parts.append(net)
        net = torch.cat(parts, dim=3)                                         # [skip, dilated, up]
        p = "conv_%d/" % (num_layers + i)
        net = rec(p + "relu1", torch.relu(conv2d_valid(net, params[p + "conv1/kernel"],
                                                       params[p + "conv1/bias"])))
        net = rec(p + "relu2", torch.relu(conv2d_valid(net, params[p + "conv2/kernel"],
                                                       params[p + "conv2/bias"])))
    assert len(conv) == 0
    logits = conv2d_valid(net, params["weight_output/kernel"], params["weight_output/bias"])
    return rec("logits", logits)


def loss_and_probs(logits, labels):
    """tf_aerial_images.py:103-110 (mean sparse softmax CE) and :147-148 (P(road))."""
    b, h, w, _ = logits.shape
    loss = F.cross_entropy(logits.reshape(-1, 2), labels.reshape(-1).long(), reduction="mean")
    probs = torch.softmax(logits, dim=3)[..., 1]
    return loss, probs


def learning_rate(lr0, global_step):
    """tf.train.exponential_decay(lr, step, 1000, 0.95, staircase=True), tf_aerial_images.py:116."""
    return lr0 * 0.95 ** (global_step // 1000)


def momentum_step(params, grads, accs, lr, momentum):
    """tf.train.MomentumOptimizer (non-Nesterov): acc = m*acc + g ; w -= lr*acc (:120-121)."""
    for k in params:
        if k not in grads or grads[k] is None:
            continue
        accs[k] = momentum * accs[k] + grads[k]
        params[k] = params[k] - lr * accs[k]
    return params, accs


def d4_apply(x, flip, k):
    """Net effect of stochastic_images_augmentation (tf_aerial_images.py:173-210) on one sample:
    flip_up_down applied `flip` (XOR of the three coins, :188) times, then rot90 by k (ccw)."""
    x = np.asarray(x)
    if flip:
        x = x[::-1]
    return np.rot90(x, k=k, axes=(0, 1)).copy()


# ----------------------------------------------------------------------------- train step
def to_torch(params, dtype=torch.float32, requires_grad=False):
    out = OrderedDict()
    for k, v in params.items():
        t = torch.tensor(np.asarray(v), dtype=dtype)
        t.requires_grad_(requires_grad)
        out[k] = t
    return out


def train_step(X, labels, params, accs, num_layers, root_size, dilated_layers, lr, momentum,
               dropout_scales=None, dtype=torch.float32, want_acts=False, storage=None):
    """One fwd + bwd + momentum update.  Returns (loss, probs, grads, new_params, new_accs, acts)."""
    tp = to_torch(params, dtype, requires_grad=True)
    acts = OrderedDict() if want_acts else None
    Xt = torch.tensor(np.asarray(X), dtype=dtype)
    logits = forward(Xt, tp, num_layers, root_size, dilated_layers, dropout_scales, acts=acts,
                     storage=storage)
    loss, probs = loss_and_probs(logits, torch.tensor(np.asarray(labels)))
    if acts is not None:
        for t in acts.values():
            t.retain_grad()
    loss.backward()
    grads = OrderedDict((k, (v.grad.detach().numpy() if v.grad is not None else None))
                        for k, v in tp.items())
    new_params = OrderedDict((k, np.asarray(v)) for k, v in params.items())
    new_accs = OrderedDict((k, np.asarray(v)) for k, v in accs.items())
    new_params, new_accs = momentum_step(new_params, grads, new_accs, lr, momentum)
    return float(loss.detach()), probs.detach().numpy(), grads, new_params, new_accs, acts
