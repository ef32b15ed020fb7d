"""U-Net model builder -- host-side mirror of the reference's src/unet.py.

`forward(X, num_layers, root_size, dilated_layers, dropout_keep=None)` and
`input_size_needed(output_size, num_layers)` keep the reference signatures (src/unet.py:12, :100).
Instead of a TensorFlow graph, `forward` plans a `UNet` engine: a static schedule of hand-written
sm_100a kernels (librsu_b200.so) over preallocated NHWC bf16 buffers, with fp32 master weights
stored under TensorFlow's variable names and layouts.

Layer map (src/unet.py line -> kernel):
  :22-23  color_space_adjust   -> rsu_color_im2col (fused with the Cin = 3 im2col)
  :29-30  dropout              -> fused into rsu_color_im2col / rsu_dropout
  :34-39  dilated 3x3 pair     -> rsu_conv_gemm, tap stride 2, only the window the decoder crops
  :42-45  3x3 conv + ReLU      -> rsu_conv_gemm (tcgen05 implicit GEMM, bias + ReLU epilogue)
  :52     2x2 max pool         -> rsu_maxpool2x2
  :67-68  conv2d_transpose     -> rsu_conv_gemm with pixel-shuffle epilogue
  :70-85  crop + concat        -> never materialised: the consumer's K loop walks 2-3 sources
  :95     weight_output + loss -> rsu_head (1x1 conv + softmax + CE + gradients)
The deepest dilated pair and the last pool are dead code in the reference (:56-59): their weights
exist (checkpoint parity) but are never evaluated.
"""
from collections import OrderedDict

import numpy as np
import torch

from . import ops


def input_size_needed(output_size, num_layers):
    """Utility function to compute image size for a given U-Net output (src/unet.py:100-115)."""
    for i in range(num_layers - 1):
        assert output_size % 2 == 0, 'expand layer {} has size {} not divisible by 2' \
            .format(num_layers - i, output_size)
        output_size = (output_size + 4) / 2

    for i in range(num_layers - 1):
        output_size = (output_size + 4) * 2

    return int(output_size + 4)


def variable_shapes(num_layers, root_size, dilated_layers):
    """TensorFlow variables of src/unet.py:23-95 in creation order: name -> shape."""
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
    f //= 2
    net_c = f
    for i in range(num_layers - 1):
        f //= 2
        shapes["up_conv_%d/kernel" % i] = (2, 2, f, net_c)
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


def glorot_init(num_layers, root_size, dilated_layers, seed=2017):
    """tf.layers defaults: glorot-uniform kernels, zero biases, drawn in creation order."""
    rs = np.random.RandomState(seed)
    out = OrderedDict()
    for name, shape in variable_shapes(num_layers, root_size, dilated_layers).items():
        if name.endswith("bias"):
            out[name] = np.zeros(shape, dtype=np.float32)
        else:
            kh, kw, a, b = shape
            limit = np.sqrt(6.0 / (kh * kw * a + kh * kw * b))
            out[name] = rs.uniform(-limit, limit, size=shape).astype(np.float32)
    return out


def flat_layout(num_layers, root_size, dilated_layers):
    """Offsets of every variable inside the flat fp32 vectors (each 256-byte aligned) and their
    total length: (OrderedDict name -> offset, n_flat)."""
    offsets, off = OrderedDict(), 0
    for name, shape in variable_shapes(num_layers, root_size, dilated_layers).items():
        offsets[name] = off
        off += (int(np.prod(shape)) + 63) // 64 * 64
    return offsets, off


class _Conv:
    """One 3x3 convolution of the plan: where its operands live and how it is wired."""

    def __init__(self, name, cin, cout, dilation=1):
        self.name, self.cin, self.cout, self.dilation = name, cin, cout, dilation
        self.w_fwd = self.w_dgrad = None


class UNet:
    """Static execution plan + state of one U-Net instance on one GPU.

    State: `params` / `grads` / `momentum` are flat fp32 device vectors; `var(name)` returns the
    TensorFlow-layout view of one variable.  All activations live in preallocated buffers sized
    for (batch_size, input_size), so a step is a fixed sequence of kernel launches.
    """

    def __init__(self, num_layers, root_size, dilated_layers, batch_size, input_size,
                 device="cuda", seed=2017, training=True, params=None, weights_from=None,
                 flat_buffers=None):
        """params: dict name -> array (TensorFlow names / layouts); None = glorot init; {} = leave
        the weights zero (a checkpoint follows).  weights_from: a donor UNet of the same
        architecture whose master and packed weights this forward-only engine aliases (enlarged
        prediction windows: no second copy of the weights, no repack).  flat_buffers: {"params",
        "grads"} fp32 device vectors of flat_layout()'s length to use instead of fresh allocations
        (symmetric memory shared with the other data-parallel ranks, see dp.PeerOptimizer)."""
        if root_size not in (64, 128, 256):
            # the head kernel (rsu_head) is instantiated for 64 / 128 / 256 input channels and the
            # tcgen05 K chunk is 64 channels: anything else would only fail at the first forward
            raise ValueError("root_size must be 64, 128 or 256 on this engine; got %d" % root_size)
        if weights_from is not None:
            d = weights_from
            assert not training and (d.L, d.root, d.dilated) == (num_layers, root_size, bool(dilated_layers))
        self._donor = weights_from
        self._flat_buffers = flat_buffers
        if not torch.cuda.is_available():
            raise RuntimeError("the B200 U-Net engine needs a CUDA device; there is no CPU fallback")
        self.L, self.root, self.dilated = num_layers, root_size, bool(dilated_layers)
        self.B, self.S = batch_size, input_size
        self.device = torch.device(device)
        self.training = training
        self.seed = seed
        self.global_step = 0
        # data-parallel hook: called as on_bucket_ready(start, end) when the flat-gradient slice
        # [start, end) is final, so its all-reduce overlaps the rest of the backward pass
        self.on_bucket_ready = None
        self.on_backward_begin = None  # called at the start of backward() (arms the exchange)
        self._plan_geometry()
        self._flops = {k: v * batch_size for k, v in
                       plan_flops(num_layers, root_size, dilated_layers, self.P).items()}
        self._alloc_state(params)
        self._alloc_buffers()
        self.pack_weights()

    # ------------------------------------------------------------------ geometry
    def _plan_geometry(self):
        L, S = self.L, self.S
        self.f = [self.root * 2 ** i for i in range(L)]
        self.in_size = []
        s = S
        for i in range(L):
            self.in_size.append(s)
            skip = s - 4
            if skip <= 0:
                raise AssertionError("input size %d too small for %d layers" % (S, L))
            if i < L - 1:
                assert skip % 2 == 0, "level %d size %d not divisible by 2" % (i, skip)
                s = skip // 2
        self.skip_size = [s_ - 4 for s_ in self.in_size]
        net = self.skip_size[L - 1]
        self.up_size, self.dec_out = [], []
        for j in range(L - 1):
            t = 2 * net
            self.up_size.append(t)
            net = t - 4
            self.dec_out.append(net)
        self.P = net if L > 1 else self.skip_size[0]
        for j in range(L - 1):
            i = L - 2 - j
            assert self.skip_size[i] >= self.up_size[j]
            if self.dilated:
                assert self.in_size[i] - 8 >= self.up_size[j]

    def level_of_decoder(self, j):
        return self.L - 2 - j

    # ------------------------------------------------------------------ state
    def _alloc_state(self, params):
        shapes = variable_shapes(self.L, self.root, self.dilated)
        self.shapes = shapes
        self.offsets, off = flat_layout(self.L, self.root, self.dilated)
        self.n_flat = off
        dev = self.device
        if self._donor is not None:
            assert self._donor.n_flat == off
            self.params, self.grads, self.momentum = self._donor.params, None, None
            return
        if self._flat_buffers is not None:
            self.params, self.grads = self._flat_buffers["params"], self._flat_buffers["grads"]
            assert self.params.numel() == off == self.grads.numel() and self.params.dtype == torch.float32
        else:
            self.params = torch.zeros(off, dtype=torch.float32, device=dev)
            self.grads = torch.zeros(off, dtype=torch.float32, device=dev) if self.training else None
        self.momentum = torch.zeros(off, dtype=torch.float32, device=dev) if self.training else None
        init = params if params is not None else glorot_init(self.L, self.root, self.dilated, self.seed)
        self.load_state(init, strict=False)

    def _first_offset(self, prefix):
        for name, off in self.offsets.items():
            if name.startswith(prefix):
                return off
        raise KeyError(prefix)

    def _bucket_bounds(self):
        """Flat-gradient slices in the order the backward pass completes them: decoder blocks
        from the last (which also carries the head) to the first, then encoder levels L-1 .. 0
        (level 0 carries color_space_adjust)."""
        L = self.L
        starts = [0]
        for i in range(1, L):
            starts.append(self._first_offset("conv_dilut_%d/" % i if self.dilated else "conv_%d/" % i))
        for j in range(L - 1):
            starts.append(self._first_offset("up_conv_%d/" % j))
        ends = starts[1:] + [self.n_flat]
        enc = list(zip(starts[:L], ends[:L]))
        dec = list(zip(starts[L:], ends[L:]))
        return enc, dec

    def _ready(self, bounds):
        """A bucket of the flat gradient is final: hand its LIVE pieces to the data-parallel hook
        (the dead conv_dilut_{L-1} range -- 27 % of the flagship's parameters -- never receives a
        gradient and is not reduced)."""
        if self.on_bucket_ready is not None:
            for a, b in self.live_ranges():
                lo, hi = max(a, bounds[0]), min(b, bounds[1])
                if hi > lo:
                    self.on_bucket_ready(lo, hi)

    def var(self, name, which="params"):
        flat = getattr(self, which)
        shape = self.shapes[name]
        o = self.offsets[name]
        return flat[o:o + int(np.prod(shape))].view(shape)

    def load_state(self, params, momentum=None, strict=True):
        """Copy variables (and momentum slots) into the flat device vectors.  strict: every
        variable of the architecture must be present with its exact shape and nothing else may be
        (a partial or mismatched checkpoint must not restore silently with random-init layers)."""
        if strict:
            missing = [n for n in self.shapes if n not in params]
            unexpected = [n for n in params if n not in self.shapes]
            if missing or unexpected:
                raise ValueError("checkpoint does not match the model (num_layers=%d root_size=%d "
                                 "dilated_layers=%s): missing %s, unexpected %s"
                                 % (self.L, self.root, self.dilated, missing[:4], unexpected[:4]))
        for src, which in ((params, "params"), (momentum, "momentum")):
            if src is None or getattr(self, which) is None:
                continue
            for name in self.shapes:
                if name not in src:
                    continue
                a = np.asarray(src[name], dtype=np.float32)
                if tuple(a.shape) != tuple(self.shapes[name]):
                    raise ValueError("variable %s: checkpoint shape %s, model shape %s"
                                     % (name, tuple(a.shape), tuple(self.shapes[name])))
                self.var(name, which).copy_(torch.as_tensor(a))

    def state_dict(self, which="params"):
        return OrderedDict((n, self.var(n, which).detach().cpu().numpy().copy()) for n in self.shapes)

    def live_variables(self):
        dead = set()
        if self.dilated:
            i = self.L - 1
            dead = {"conv_dilut_%d/atrous_conv%d/%s" % (i, k, p) for k in (1, 2)
                    for p in ("kernel", "bias")}
        return [n for n in self.shapes if n not in dead]

    # ------------------------------------------------------------------ buffers
    def _bf(self, *shape):
        return torch.empty(shape, dtype=torch.bfloat16, device=self.device)

    def _alloc_buffers(self):
        L, B, f = self.L, self.B, self.f
        tr = self.training
        self.A1, self.A2, self.Pool = [], [], []
        self.D1, self.D2 = [], []
        self.dA1, self.dA2, self.dIn = [], [], []
        self.dD1, self.dD2 = [], []
        self.dil_off = []
        for i in range(L):
            s = self.in_size[i]
            self.A1.append(self._bf(B, s - 2, s - 2, f[i]))
            self.A2.append(self._bf(B, s - 4, s - 4, f[i]))
            self.Pool.append(self._bf(B, (s - 4) // 2, (s - 4) // 2, f[i]) if i < L - 1 else None)
            if tr:
                self.dA1.append(self._bf(B, s - 2, s - 2, f[i]))
                self.dA2.append(self._bf(B, s - 4, s - 4, f[i]))
                self.dIn.append(self._bf(B, s, s, f[i - 1]) if i > 0 else None)
            if self.dilated and i < L - 1:
                t = self.up_size[L - 2 - i]
                o2 = (s - 8 - t) // 2
                self.dil_off.append(o2)
                self.D1.append(self._bf(B, t + 4, t + 4, f[i]))
                self.D2.append(self._bf(B, t, t, f[i]))
                if tr:
                    self.dD1.append(self._bf(B, t + 4, t + 4, f[i]))
                    self.dD2.append(None)  # a channel slice of the concat gradient, set below
            else:
                self.dil_off.append(None)
                self.D1.append(None)
                self.D2.append(None)
                self.dD1.append(None)
                self.dD2.append(None)
        # first layer im2col buffers (Cin = 3 -> 64 padded channels)
        # (root 64 / 128: the Cin = 3 layers build their im2col operand inside the kernel,
        # rsu_first_conv_*, and no im2col tensor exists)
        s0 = self.in_size[0]
        # (TMA over the fp32 image rows needs 16-byte row strides: S % 4 == 0)
        self.fused_first = f[0] in (64, 128) and s0 % 4 == 0 and self.P >= 16
        self.col = None if self.fused_first else self._bf(B, s0 - 2, s0 - 2, 64)
        self.dcol = None  # gradient of the im2col matrix: only the dropout path needs it (lazy)
        if self.dilated and L > 1 and not self.fused_first:
            t0 = self.up_size[L - 2]
            self.colD = self._bf(B, t0 + 4, t0 + 4, 64)
        else:
            self.colD = None
        self.dcolD = None
        # identity colour transform: the folded first layer (keep == 1) feeds im2col(x - 0.5)
        self._eye3 = torch.eye(3, dtype=torch.float32, device=self.device)
        self._zero3 = torch.zeros(3, dtype=torch.float32, device=self.device)
        self.U, self.C1, self.C2 = [], [], []
        self.dCat, self.dC1, self.dC2 = [], [], []
        for j in range(L - 1):
            fo = f[L - 2 - j]
            t = self.up_size[j]
            self.U.append(self._bf(B, t, t, fo))
            self.C1.append(self._bf(B, t - 2, t - 2, fo))
            self.C2.append(self._bf(B, t - 4, t - 4, fo))
            if tr:
                self.dCat.append(self._bf(B, t, t, fo * (3 if self.dilated else 2)))
                self.dC1.append(self._bf(B, t - 2, t - 2, fo))
                self.dC2.append(self._bf(B, t - 4, t - 4, fo))
                if self.dilated:
                    # d(dilated skip): the data-gradient kernel of conv_{L+j}/conv1 applies the
                    # ReLU mask of D2 to this channel range while it writes the concat gradient
                    self.dD2[L - 2 - j] = self.dCat[j][..., fo:2 * fo]
        P = self.P
        dev = self.device
        self.probs = torch.empty(B, P, P, dtype=torch.float32, device=dev)
        self.logits = torch.empty(B, P, P, 2, dtype=torch.float32, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        # dropout scratch (dropped copies of pooled tensors / decoder inputs), allocated lazily
        self._drop_pool = [None] * L
        self._drop_net = [None] * max(L - 1, 0)
        # packed bf16 weights
        self.convs = OrderedDict()
        if self._donor is not None:
            self.convs = self._donor.convs
            return
        for name, shape in self.shapes.items():
            if not name.endswith("kernel") or name.startswith(("color_space", "weight_output")):
                continue
            key = name[:-len("/kernel")]
            if key.startswith("up_conv"):
                _, _, cout, cin = shape
                c = _Conv(key, cin, cout)
                c.w_fwd = self._bf(4 * cout, cin)
                c.w_dgrad = self._bf(cin, 4 * cout) if tr else None
            else:
                _, _, cin, cout = shape
                c = _Conv(key, cin, cout, 2 if "atrous" in key else 1)
                if cin == 3:
                    c.w_fwd = torch.zeros(cout, 64, dtype=torch.bfloat16, device=dev)
                    c.w_dgrad = torch.zeros(64, cout, dtype=torch.bfloat16, device=dev) if tr else None
                    c.dw_stage = torch.zeros(64, cout, dtype=torch.float32, device=dev) if tr else None
                    # color_space_adjust folded into the kernel / bias (used when keep == 1)
                    c.w_fold = torch.zeros(cout, 64, dtype=torch.bfloat16, device=dev)
                    c.bias_fold = torch.zeros(cout, dtype=torch.float32, device=dev)
                else:
                    c.w_fwd = self._bf(cout, 9 * cin)
                    c.w_dgrad = self._bf(cin, 9 * cout) if tr else None
            self.convs[key] = c
        if self.dilated:  # dead pair: weights exist, never packed or evaluated
            for k in (1, 2):
                self.convs.pop("conv_dilut_%d/atrous_conv%d" % (L - 1, k), None)

    # ------------------------------------------------------------------ weights
    def _pack_jobs(self):
        jobs = []
        for key, c in self.convs.items():
            w = self.var(key + "/kernel")
            if key.startswith("up_conv"):
                jobs.append((ops.PACK_CAST, w, c.w_fwd, 1, 1, w.numel(), 0))
                if c.w_dgrad is not None:
                    jobs.append((ops.PACK_TRANSPOSE, w, c.w_dgrad, 1, 4 * c.cout, c.cin, 0))
            elif c.cin == 3:
                jobs.append((ops.PACK_TRANSPOSE, w, c.w_fwd, 1, 27, c.cout, 64))
                if c.w_dgrad is not None:
                    jobs.append((ops.PACK_PERMUTE, w, c.w_dgrad, 1, 27, c.cout, 0))
            else:
                jobs.append((ops.PACK_TRANSPOSE, w, c.w_fwd, 9, c.cin, c.cout, 0))
                if c.w_dgrad is not None:
                    jobs.append((ops.PACK_PERMUTE, w, c.w_dgrad, 9, c.cin, c.cout, 0))
        return jobs

    def pack_weights(self):
        """fp32 master (TensorFlow layouts) -> bf16 GEMM operand layouts; run after every update:
        one table-driven launch for all repacks plus the two folded first-layer kernels."""
        if self._donor is not None:
            return  # the donor's packed operands are this engine's
        if getattr(self, "_pack_plan", None) is None:
            self._pack_plan = ops.PackPlan(self._pack_jobs(), self.device)
        self._pack_plan.run()
        for key, c in self.convs.items():
            if c.cin == 3 and not key.startswith("up_conv"):
                ops.first_layer_fold(self.var(key + "/kernel"), self.var(key + "/bias"),
                                     self.var("color_space_adjust/kernel"),
                                     self.var("color_space_adjust/bias"), c.w_fold, c.bias_fold)

    # ------------------------------------------------------------------ forward
    def _site_seed(self, site):
        return (self.seed * 1000003 + self.global_step * 131 + site) & 0xFFFFFFFFFFFF

    def forward(self, images, labels=None, keep=1.0, want_logits=False):
        """images: fp32 [B,S,S,3] device tensor in [0,1].  Fills self.probs (and self.loss and the
        head gradients when labels (uint8 [B,P,P]) are given).  Returns probs."""
        L, f = self.L, self.f
        assert images.shape == (self.B, self.S, self.S, 3) and images.dtype == torch.float32
        self._images = images
        self._keep = float(keep)
        w1 = self.var("color_space_adjust/kernel")
        b1 = self.var("color_space_adjust/bias")
        bias = lambda key: self.var(key + "/bias")
        drop = keep < 1.0
        site = 0
        net = None
        for i in range(L):
            reg1, reg2 = self.convs["conv_%d/conv1" % i], self.convs["conv_%d/conv2" % i]
            dil_live = self.dilated and i < L - 1
            o2 = self.dil_off[i]
            if i == 0:
                # without dropout color_space_adjust is folded into the convolution: the im2col
                # matrix holds x - 0.5 (identity transform) and the kernel / bias are W1.W, b + b1.W
                seed0 = self._site_seed(site)
                cw, cb = (w1, b1) if drop else (self._eye3, self._zero3)
                self._first_fwd(reg1, images, cw, cb, 1, 0, 0, self.col, self.A1[0], drop, keep, seed0)
                if dil_live:
                    d1 = self.convs["conv_dilut_0/atrous_conv1"]
                    self._first_fwd(d1, images, cw, cb, 2, o2, o2, self.colD, self.D1[0], drop, keep,
                                    seed0)
            else:
                src = self.Pool[i - 1]
                if drop:
                    if self._drop_pool[i] is None:
                        self._drop_pool[i] = torch.empty_like(src)
                    ops.dropout(src, self._drop_pool[i], keep, self._site_seed(site))
                    src = self._drop_pool[i]
                net = src
                self._tag(reg1.name)
                ops.conv3x3_fwd([(src, 0, 0)], reg1.w_fwd, bias(reg1.name), self.A1[i])
                if dil_live:
                    d1 = self.convs["conv_dilut_%d/atrous_conv1" % i]
                    self._tag(d1.name)
                    ops.conv3x3_fwd([(src, o2, o2)], d1.w_fwd, bias(d1.name), self.D1[i], dilation=2)
            site += 1
            if dil_live:
                d2 = self.convs["conv_dilut_%d/atrous_conv2" % i]
                self._tag(d2.name)
                ops.conv3x3_fwd([(self.D1[i], 0, 0)], d2.w_fwd, bias(d2.name), self.D2[i], dilation=2)
            self._tag(reg2.name)
            pooled = ops.conv3x3_fwd([(self.A1[i], 0, 0)], reg2.w_fwd, bias(reg2.name), self.A2[i],
                                     pool_out=self.Pool[i] if i < L - 1 else None)
            if i < L - 1 and not pooled:
                ops.maxpool2x2(self.A2[i], self.Pool[i])
        net = self.A2[L - 1]
        self._dec_in = []
        for j in range(L - 1):
            i = L - 2 - j
            up = self.convs["up_conv_%d" % j]
            c1, c2 = self.convs["conv_%d/conv1" % (L + j)], self.convs["conv_%d/conv2" % (L + j)]
            if drop:
                if self._drop_net[j] is None:
                    self._drop_net[j] = torch.empty_like(net)
                ops.dropout(net, self._drop_net[j], keep, self._site_seed(site))
                net = self._drop_net[j]
            site += 1
            self._dec_in.append(net)
            self._tag(up.name)
            ops.upconv2x2_fwd(net, up.w_fwd, bias(up.name), self.U[j])
            t = self.up_size[j]
            so = (self.skip_size[i] - t) // 2
            srcs = [(self.A2[i], so, so)]
            if self.dilated:
                srcs.append((self.D2[i], 0, 0))
            srcs.append((self.U[j], 0, 0))
            self._tag(c1.name)
            ops.conv3x3_fwd(srcs, c1.w_fwd, bias(c1.name), self.C1[j])
            self._tag(c2.name)
            ops.conv3x3_fwd([(self.C1[j], 0, 0)], c2.w_fwd, bias(c2.name), self.C2[j])
            net = self.C2[j]
        self._last = net
        wh = self.var("weight_output/kernel").view(-1, 2)
        bh = self.var("weight_output/bias")
        if labels is None:
            ops.head(net, wh, bh, probs=self.probs, logits=self.logits if want_logits else None)
        else:
            assert self.training
            ops.fill_zero(self.loss)
            dlast = self.dC2[L - 2] if L > 1 else self.dA2[0]
            ops.head(net, wh, bh, labels=labels, probs=self.probs,
                     logits=self.logits if want_logits else None, loss=self.loss, dz=dlast,
                     dw=self.var("weight_output/kernel", "grads").view(-1, 2),
                     db=self.var("weight_output/bias", "grads"))
        return self.probs

    def _first_fwd(self, conv, images, cw, cb, dilation, oy, ox, col, out, drop, keep, seed):
        """Cin = 3 convolution: im2col built inside the kernel (root 64 / 128), else through a
        materialised im2col tensor and the generic implicit GEMM."""
        w = conv.w_fwd if drop else conv.w_fold
        b = self.var(conv.name + "/bias") if drop else conv.bias_fold
        self._tag(conv.name)
        if self.fused_first:
            ops.first_conv_fwd(images, cw if drop else None, cb if drop else None, dilation, oy, ox, w, b,
                               out, relu=True, keep=keep, seed=seed)
        else:
            ops.color_im2col(images, cw, cb, dilation, oy, ox, col, keep, seed)
            ops.conv_gemm([(col, 0, 0)], [(0, 0)], w, out, out.shape[3], bias=b, relu=True)

    # ------------------------------------------------------------------ backward
    def _tag(self, name):
        ops.set_layer(name, self._flops.get(name, 0.0))

    def _conv_bwd(self, conv, srcs, dz, dx, mask=None, accumulate=False, need_dx=True, mask_c0=0):
        """wgrad + bias grad (+ dgrad) of one 3x3 convolution."""
        self._tag(conv.name)
        g = lambda n: self.var(conv.name + "/" + n, "grads")
        if not ops.conv3x3_wgrad(srcs, dz, g("kernel").view(-1, conv.cout), dilation=conv.dilation,
                                 bias_grad=g("bias")):
            ops.bias_grad(dz, g("bias"))
        if need_dx:
            ops.conv3x3_dgrad(dz, conv.w_dgrad, dx, dilation=conv.dilation, mask=mask,
                              accumulate=accumulate, mask_c0=mask_c0)

    def _first_bwd(self, conv, col, which_dcol, dz, dilation, oy, ox):
        """Cin = 3 convolution through its im2col matrix, plus d(color_space_adjust)."""
        self._tag(conv.name)
        g = lambda n: self.var(conv.name + "/" + n, "grads")
        ops.fill_zero(conv.dw_stage)
        if self.fused_first:
            drop = self._keep < 1.0
            cw = self.var("color_space_adjust/kernel") if drop else None
            cb = self.var("color_space_adjust/bias") if drop else None
            ops.first_conv_wgrad(self._images, cw, cb, dilation, oy, ox, dz, conv.dw_stage,
                                 keep=self._keep, seed=self._site_seed(0))
        else:
            ops.wgrad_gemm([(col, 0, 0)], [(0, 0)], dz, (0, 0), conv.dw_stage, (dz.shape[1], dz.shape[2]))
        if self._keep >= 1.0:
            # folded layer: everything follows from the 28 x Cout matrix im2col(x - 0.5)^T dZ
            # (row 27, the constant-one column, is the bias gradient)
            ops.first_layer_grads(conv.dw_stage, self.var(conv.name + "/kernel"),
                                  self.var("color_space_adjust/kernel"),
                                  self.var("color_space_adjust/bias"), g("kernel"), g("bias"),
                                  self.var("color_space_adjust/kernel", "grads"),
                                  self.var("color_space_adjust/bias", "grads"))
            return
        g("kernel").view(27, conv.cout).add_(conv.dw_stage[:27])
        g("bias").add_(conv.dw_stage[27])
        if getattr(self, which_dcol) is None:
            setattr(self, which_dcol, self._bf(dz.shape[0], dz.shape[1], dz.shape[2], 64))
        dcol = getattr(self, which_dcol)
        ops.conv_gemm([(dz, 0, 0)], [(0, 0)], conv.w_dgrad, dcol, 64)
        ops.color_im2col_bwd(self._images, dcol, dilation, oy, ox,
                             self.var("color_space_adjust/kernel", "grads"),
                             self.var("color_space_adjust/bias", "grads"), self._keep,
                             self._site_seed(0))

    def backward(self):
        """Gradients of the mean cross-entropy w.r.t. every live variable (into self.grads, which
        the caller zeroed before forward(labels=...))."""
        L, f = self.L, self.f
        keep = self._keep
        drop = keep < 1.0
        if self.on_backward_begin is not None and self.on_bucket_ready is not None:
            self.on_backward_begin(getattr(self, "_bwd_scale", 1.0))
        enc_buckets, dec_buckets = self._bucket_bounds()
        for j in range(L - 2, -1, -1):
            i = L - 2 - j
            fo = f[i]
            up = self.convs["up_conv_%d" % j]
            c1, c2 = self.convs["conv_%d/conv1" % (L + j)], self.convs["conv_%d/conv2" % (L + j)]
            t = self.up_size[j]
            so = (self.skip_size[i] - t) // 2
            self._conv_bwd(c2, [(self.C1[j], 0, 0)], self.dC2[j], self.dC1[j], mask=self.C1[j])
            srcs = [(self.A2[i], so, so)]
            if self.dilated:
                srcs.append((self.D2[i], 0, 0))
            srcs.append((self.U[j], 0, 0))
            # (dilated nets: ReluGrad of the dilated skip D2 is applied to channels [fo, 2 fo) of
            # the concat gradient by this kernel's epilogue; dD2 is that slice)
            self._conv_bwd(c1, srcs, self.dC1[j], self.dCat[j],
                           mask=self.D2[i] if self.dilated else None, mask_c0=fo)
            dcat = self.dCat[j]
            d_up = dcat[..., (2 if self.dilated else 1) * fo:]
            x_in = self._dec_in[j]
            g = lambda n: self.var(up.name + "/" + n, "grads")
            self._tag(up.name)
            ops.upconv2x2_wgrad(d_up, x_in, g("kernel").view(4 * up.cout, up.cin))
            ops.bias_grad(d_up, g("bias"))
            # gradient into the tensor that fed the transpose conv (previous decoder output, or
            # the bottom of the encoder), with its ReLU mask fused
            if j > 0:
                dst, act = self.dC2[j - 1], self.C2[j - 1]
            else:
                dst, act = self.dA2[L - 1], self.A2[L - 1]
            ops.upconv2x2_dgrad(d_up, up.w_dgrad, dst, mask=x_in if drop else act)
            if drop:
                ops.dropout(dst, dst, keep, self._site_seed(L + j))
            self._ready(dec_buckets[j])
        for i in range(L - 1, -1, -1):
            reg1, reg2 = self.convs["conv_%d/conv1" % i], self.convs["conv_%d/conv2" % i]
            dil_live = self.dilated and i < L - 1
            if i < L - 1:
                j = L - 2 - i
                t = self.up_size[j]
                so = (self.skip_size[i] - t) // 2
                d_in_next = self.dIn[i + 1]
                if drop:
                    ops.dropout(d_in_next, d_in_next, keep, self._site_seed(i + 1))
                ops.skip_grad(self.A2[i], d_in_next, self.dCat[j][..., :f[i]], (so, so), self.dA2[i])
            self._conv_bwd(reg2, [(self.A1[i], 0, 0)], self.dA2[i], self.dA1[i], mask=self.A1[i])
            if i > 0:
                src = self._drop_pool[i] if drop else self.Pool[i - 1]
                self._conv_bwd(reg1, [(src, 0, 0)], self.dA1[i], self.dIn[i])
            else:
                self._first_bwd(reg1, self.col, "dcol", self.dA1[0], 1, 0, 0)
            if dil_live:
                d1 = self.convs["conv_dilut_%d/atrous_conv1" % i]
                d2 = self.convs["conv_dilut_%d/atrous_conv2" % i]
                o2 = self.dil_off[i]
                self._conv_bwd(d2, [(self.D1[i], 0, 0)], self.dD2[i], self.dD1[i], mask=self.D1[i])
                if i > 0:
                    src = self._drop_pool[i] if drop else self.Pool[i - 1]
                    tt = self.D1[i].shape[1] + 4
                    win = self.dIn[i][:, o2:o2 + tt, o2:o2 + tt, :]
                    self._conv_bwd(d1, [(src, o2, o2)], self.dD1[i], win, accumulate=True)
                else:
                    self._first_bwd(d1, self.colD, "dcolD", self.dD1[0], 2, o2, o2)
            self._ready(enc_buckets[i])

    # ------------------------------------------------------------------ optimizer
    def learning_rate(self, lr0):
        """tf.train.exponential_decay(lr, global_step, 1000, 0.95, staircase=True)."""
        return lr0 * 0.95 ** (self.global_step // 1000)

    def live_ranges(self):
        """Slices of the flat parameter vector that ever receive a gradient: everything but the
        dead dilated pair of the deepest level (unet.py:57-59), whose gradient and momentum stay
        zero -- TensorFlow creates no ApplyMomentum op for it either."""
        if getattr(self, "_live_ranges", None) is None:
            if self.dilated:
                d0 = self.offsets["conv_dilut_%d/atrous_conv1/kernel" % (self.L - 1)]
                d1 = self.offsets["conv_%d/conv1/kernel" % (self.L - 1)]
                self._live_ranges = [(a, b) for a, b in ((0, d0), (d1, self.n_flat)) if b > a]
            else:
                self._live_ranges = [(0, self.n_flat)]
        return self._live_ranges

    def zero_grads(self):
        for a, b in self.live_ranges():
            ops.fill_zero(self.grads[a:b])

    def apply_gradients(self, lr0, momentum, grad_scale=1.0, peer=None):
        """tf.train.MomentumOptimizer step (+ global_step, + bf16 operand repack).  peer: a
        dp.PeerOptimizer -- the update then also IS the gradient exchange of the data-parallel
        ranks (each rank updates its slice from all ranks' gradients and writes everybody's
        weights)."""
        lr = self.learning_rate(lr0)
        if peer is not None:
            peer.step(self, lr, momentum, extra_scale=grad_scale)
        else:
            for a, b in self.live_ranges():
                ops.momentum_sgd(self.params[a:b], self.momentum[a:b], self.grads[a:b], lr, momentum,
                                 grad_scale)
        self.global_step += 1
        self.pack_weights()

    def accumulate_step(self, micro_batches, lr0=0.01, momentum=0.9, keep=1.0, peer=None, finish=None):
        """One optimizer step over several micro-batches [(images, labels), ...] of this engine's
        batch size: gradients accumulate in the flat fp32 vector (every weight-gradient kernel adds
        into it) and the update uses their mean -- a global batch that does not fit in HBM at once
        (2052^2 patches: ~20 GB of activations each).  peer / finish: the data-parallel exchange
        (dp.PeerOptimizer, or a callable that completes the bucketed all-reduce and returns its
        scale); the bucket hook only fires during the last micro-batch."""
        hook = self.on_bucket_ready
        self.zero_grads()
        self._bwd_scale = 1.0 / len(micro_batches)  # what an exchange armed during backward() applies
        for i, (images, labels) in enumerate(micro_batches):
            self.on_bucket_ready = hook if i == len(micro_batches) - 1 else None
            self.forward(images, labels, keep)
            self.backward()
        self.on_bucket_ready = hook
        self._bwd_scale = 1.0
        scale = (finish() if finish is not None else 1.0) / len(micro_batches)
        self.apply_gradients(lr0, momentum, scale, peer=peer)
        return self.loss

    def train_step(self, images, labels, lr0=0.01, momentum=0.9, keep=1.0, allreduce=None):
        """forward + backward + momentum update; returns the device scalar loss tensor."""
        self.zero_grads()
        self.forward(images, labels, keep)
        self.backward()
        scale = 1.0
        if allreduce is not None:
            scale = allreduce(self.grads)
        self.apply_gradients(lr0, momentum, scale)
        return self.loss


class Placeholder:
    """Stand-in for tf.placeholder: only carries the static shape [B, S, S, 3]."""

    def __init__(self, shape, name="patches"):
        self.shape, self.name = tuple(shape), name


def forward(X, num_layers, root_size, dilated_layers, dropout_keep=None, params=None, seed=2017):
    """Build the U-Net (src/unet.py:12-97).

    X is either a `Placeholder` (static shape, like the reference's graph mode) -- the planned
    `UNet` engine is returned -- or an array [B,S,S,3] in [0,1], in which case the network is run
    once (with `params`, a dict keyed by TensorFlow variable names, or a fresh glorot init) and the
    logits [B,P,P,2] are returned as a NumPy array.
    """
    if isinstance(X, Placeholder):
        b, s = X.shape[0], X.shape[1]
        return UNet(num_layers, root_size, dilated_layers, b, s, seed=seed, params=params)
    x = torch.as_tensor(np.asarray(X, dtype=np.float32)).cuda()
    net = UNet(num_layers, root_size, dilated_layers, x.shape[0], x.shape[1], seed=seed,
               training=False, params=params)
    keep = 1.0 if dropout_keep is None else float(dropout_keep)
    net.forward(x, keep=keep, want_logits=True)
    return net.logits.cpu().numpy()


def plan_flops(num_layers, root_size, dilated_layers, patch_size):
    """Algorithmic forward FLOPs per patch of every live layer (SURVEY.md 8(d) "F_min"):
    2 * k^2 * Cin * Cout * Hout^2 per convolution with Hout the extent actually consumed
    downstream (dilated branches only on the window the decoder crops), 2 * 4 * Cin * Cout * Hin^2
    per transpose convolution; dead layers count zero.  Returns an OrderedDict name -> FLOPs."""
    L, f0 = num_layers, root_size
    S = input_size_needed(patch_size, L)
    f = [f0 * 2 ** i for i in range(L)]
    in_size, s = [], S
    for i in range(L):
        in_size.append(s)
        s = (s - 4) // 2
    skip = [v - 4 for v in in_size]
    up, net = [], skip[L - 1]
    for j in range(L - 1):
        up.append(2 * net)
        net = 2 * net - 4
    out = OrderedDict()
    out["color_space_adjust"] = 2 * 3 * 3 * S * S
    for i in range(L):
        cin = 3 if i == 0 else f[i - 1]
        if dilated_layers and i < L - 1:
            t = up[L - 2 - i]
            out["conv_dilut_%d/atrous_conv1" % i] = 2 * 9 * cin * f[i] * (t + 4) ** 2
            out["conv_dilut_%d/atrous_conv2" % i] = 2 * 9 * f[i] * f[i] * t ** 2
        out["conv_%d/conv1" % i] = 2 * 9 * cin * f[i] * (in_size[i] - 2) ** 2
        out["conv_%d/conv2" % i] = 2 * 9 * f[i] * f[i] * (in_size[i] - 4) ** 2
    net_c, net = f[L - 1], skip[L - 1]
    for j in range(L - 1):
        fo = f[L - 2 - j]
        out["up_conv_%d" % j] = 2 * 4 * net_c * fo * net ** 2
        t = 2 * net
        cat = fo * (3 if dilated_layers else 2)
        out["conv_%d/conv1" % (L + j)] = 2 * 9 * cat * fo * (t - 2) ** 2
        out["conv_%d/conv2" % (L + j)] = 2 * 9 * fo * fo * (t - 4) ** 2
        net_c, net = fo, t - 4
    out["weight_output"] = 2 * net_c * 2 * net ** 2
    return out
