"""End-to-end parity of the B200 U-Net engine against the CPU oracle (oracle/unet_oracle.py):
same glorot weights (seed 2017), same synthetic inputs, one forward + backward + momentum step.

Gates (BASELINE.json north_star: 2e-2 relative, bf16 vs the reference's fp32):
  * every activation, the loss and P(road) against the fp32 oracle             < 2e-2
  * every weight / bias gradient and the momentum update against the oracle evaluated with the
    device's storage precision (oracle `storage="bf16"`: same algorithm, conv operands/results
    rounded to bf16, fp32 accumulate)                                            < 2e-2
  * gradients against the fp32 oracle: reported, and bounded by the precision floor -- the
    deviation that the bf16-storage ORACLE ITSELF shows against fp32 (max-pool arg-max and ReLU
    sign flips on near-ties make a randomly initialised network's gradient chaotic at the 3-10 %
    level for ANY bf16 implementation; see DESIGN.md "precision floor").
"""
import numpy as np
import pytest
import torch

from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu
TOL = 2e-2


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def synth(B, S, P, seed=2017):
    rs = np.random.RandomState(seed)
    X = rs.rand(B, S, S, 3).astype(np.float32)
    field = (rs.rand(B, P, P) < 0.25).astype(np.float32)
    k = np.ones((5, 5), dtype=np.float32) / 25.0
    sm = torch.nn.functional.conv2d(torch.tensor(field)[:, None], torch.tensor(k)[None, None], padding=2)
    labels = (sm[:, 0].numpy() >= 0.25).astype(np.uint8)
    return X, labels


def make_params(L, root, dil):
    params = O.init_params(L, root, dil, seed=2017)
    rs = np.random.RandomState(5)  # non-zero biases so the bias path is exercised
    for k in params:
        if k.endswith("bias"):
            params[k] = (0.05 * rs.randn(*params[k].shape)).astype(np.float32)
    return params


CASES = [
    # (L, root, dilated, P, B)
    (3, 64, False, 20, 2),
    (3, 64, True, 36, 2),
    (4, 64, True, 52, 1),
    (4, 64, False, 388, 1),  # BASELINE.json configs[0] at full size: 476^2 -> 388^2, batch 1
    (6, 64, True, 388, 1),   # the flagship architecture of configs[1..3] at full size, batch 1
]


@pytest.mark.parametrize("case", CASES)
def test_train_step_parity(case):
    from road_segmentation_unet_b200 import unet
    L, root, dil, P, B = case
    S = unet.input_size_needed(P, L)
    assert S == O.input_size_needed(P, L)
    params = make_params(L, root, dil)
    X, labels = synth(B, S, P)
    accs = {k: np.zeros_like(v) for k, v in params.items()}
    loss32, probs32, grads32, _, _, acts32 = O.train_step(
        X, labels, params, accs, L, root, dil, lr=0.01, momentum=0.9, want_acts=True)
    loss16, probs16, grads16, new_p16, _, acts16 = O.train_step(
        X, labels, params, accs, L, root, dil, lr=0.01, momentum=0.9, want_acts=True, storage="bf16")

    net = unet.UNet(L, root, dil, B, S, params=params)
    assert net.P == P
    own = unet.glorot_init(L, root, dil, 2017)  # the engine draws the same glorot weights
    ref_init = O.init_params(L, root, dil, 2017)
    for k in own:
        assert np.array_equal(own[k], ref_init[k]), k
    net.grads.zero_()
    net.forward(torch.tensor(X).cuda(), torch.tensor(labels).cuda(), keep=1.0, want_logits=True)
    net.backward()
    torch.cuda.synchronize()

    def act_pairs():
        yield "logits", net.logits, "logits", None
        for i in range(L):
            yield "conv_%d/relu1" % i, net.A1[i], "conv_%d/relu1" % i, None
            yield "conv_%d/relu2" % i, net.A2[i], "conv_%d/relu2" % i, None
            if dil and i < L - 1:
                yield "conv_dilut_%d/relu2" % i, net.D2[i], "conv_dilut_%d/relu2" % i, net.up_size[L - 2 - i]
        for j in range(L - 1):
            yield "up_conv_%d" % j, net.U[j], "up_conv_%d" % j, None
            yield "conv_%d/relu2" % (L + j), net.C2[j], "conv_%d/relu2" % (L + j), None

    # ---- activations vs the fp32 oracle (the reference's precision)
    errs = {}
    for name, got, key, crop in act_pairs():
        ref = acts32[key].detach()
        if crop is not None:
            ref = O.center_crop(ref, crop)
        errs[name] = rel(got.float().cpu().numpy(), ref.numpy())
    errs["probs"] = rel(net.probs.cpu().numpy(), probs32)
    print("activations vs fp32, worst:", sorted(errs.items(), key=lambda kv: -kv[1])[:4])
    bad = {k: v for k, v in errs.items() if not v < TOL}
    assert not bad, bad
    assert abs(net.loss.item() - loss32) < TOL * abs(loss32), (net.loss.item(), loss32)

    # ---- gradients vs the oracle at the device's storage precision
    live = net.live_variables()
    g_dev = {n: net.var(n, "grads").cpu().numpy() for n in live}
    e16 = {n: rel(g_dev[n], grads16[n]) for n in live}
    print("gradients vs bf16-storage oracle, worst:", sorted(e16.items(), key=lambda kv: -kv[1])[:4])
    # free-running chain: rare 1-ulp rounding differences flip a max-pool arg-max / ReLU sign and
    # the flip propagates; with the few pixels of these test shapes that is a few percent on the
    # deepest layers.  The strict 2e-2 per-layer gate is test_backward_per_layer_teacher_forced.
    FREE = 0.1
    bad = {k: v for k, v in e16.items() if not v < FREE}
    assert not bad, bad
    for name in O.dead_variables(L, dil):
        assert grads32[name] is None
        assert float(net.var(name, "grads").abs().max()) == 0.0

    # ---- gradients vs the fp32 oracle: bounded by the precision floor of bf16 storage
    e32 = {n: rel(g_dev[n], grads32[n]) for n in live}
    floor = {n: rel(grads16[n], grads32[n]) for n in live}
    print("gradients vs fp32 oracle, worst:", sorted(e32.items(), key=lambda kv: -kv[1])[:4])
    print("bf16-storage oracle vs fp32 oracle (floor), worst:", sorted(floor.items(), key=lambda kv: -kv[1])[:4])
    # Per layer: within the 2e-2 gate of the fp32 oracle, or -- where max-pool arg-max / ReLU
    # flips of near-ties dominate (few pixels: the deep levels at batch 1) -- no further from
    # fp32 than twice what the bf16-storage ORACLE itself is on that layer (both are
    # independent samples of the same flip noise).  test_flagship_free_running_batch4 holds the
    # production architecture to the plain 2e-2 gate with the noise averaged over 4 patches.
    bad = {n: (e32[n], floor[n]) for n in live if not (e32[n] < TOL or e32[n] < 2.0 * floor[n] + 5e-3)}
    assert not bad, bad
    if P >= 128 and L <= 4:
        # at the full size of BASELINE.json configs[0] the flip noise averages out: every weight
        # gradient of the free-running backward pass is within the 2e-2 gate of the fp32 oracle
        assert max(e32.values()) < TOL, sorted(e32.items(), key=lambda kv: -kv[1])[:4]
    assert np.mean(list(e32.values())) < 2.0 * np.mean(list(floor.values())) + 1e-2

    # ---- momentum update (tf.train.MomentumOptimizer): same gradients -> same delta
    net.apply_gradients(0.01, 0.9)
    torch.cuda.synchronize()
    for name in live:  # delta_w = -lr * (0.9 * 0 + g): exactly the device's own gradient
        # w' = w - lr * acc evaluated in fp32: identical up to the rounding of the subtraction
        want = params[name] - np.float32(0.01) * g_dev[name]
        assert np.allclose(net.var(name).cpu().numpy(), want, rtol=3e-7, atol=1e-9), name
        assert rel(net.var(name, "momentum").cpu().numpy(), g_dev[name]) < 1e-6, name
    assert net.global_step == 1


def test_dropout_parity():
    """tf.nn.dropout sites (unet.py:29-30, 64-65): the engine's masks are fed to the oracle."""
    from road_segmentation_unet_b200 import unet, ops
    L, root, dil, P, B, keep = 3, 64, True, 36, 2, 0.8
    S = unet.input_size_needed(P, L)
    params = make_params(L, root, dil)
    X, labels = synth(B, S, P, seed=3)
    net = unet.UNet(L, root, dil, B, S, params=params)
    net.grads.zero_()
    net.forward(torch.tensor(X).cuda(), torch.tensor(labels).cuda(), keep=keep)
    net.backward()
    torch.cuda.synchronize()
    scales = []
    shapes = [(B, S, S, 3)] + [tuple(net.Pool[i].shape) for i in range(L - 1)]
    shapes += [tuple(net.A2[L - 1].shape)] + [tuple(net.C2[j].shape) for j in range(L - 2)]
    for site, shp in enumerate(shapes):
        m = ops.dropout_mask(int(np.prod(shp)), keep, net._site_seed(site)).cpu().numpy()
        assert set(np.unique(m)).issubset({0.0, np.float32(1.0 / keep)})
        scales.append(torch.tensor(m.reshape(shp)))
    accs = {k: np.zeros_like(v) for k, v in params.items()}
    loss32, probs32, _, _, _, _ = O.train_step(
        X, labels, params, accs, L, root, dil, 0.01, 0.9, dropout_scales=scales)
    loss16, probs16, grads16, _, _, _ = O.train_step(
        X, labels, params, accs, L, root, dil, 0.01, 0.9, dropout_scales=scales, storage="bf16")
    assert abs(net.loss.item() - loss32) < TOL * abs(loss32)
    assert rel(net.probs.cpu().numpy(), probs32) < TOL
    e = {n: rel(net.var(n, "grads").cpu().numpy(), grads16[n]) for n in net.live_variables()}
    print("dropout: gradients vs bf16-storage oracle, worst:", sorted(e.items(), key=lambda kv: -kv[1])[:4])
    assert max(e.values()) < 8e-2 and float(np.median(list(e.values()))) < TOL, e


def test_forward_api_logits():
    """unet.forward keeps the reference signature and returns logits [B,P,P,2]."""
    from road_segmentation_unet_b200 import unet
    L, root, P = 3, 64, 20
    S = unet.input_size_needed(P, L)
    X = np.random.RandomState(0).rand(1, S, S, 3).astype(np.float32)
    logits = unet.forward(X, L, root, False)
    assert logits.shape == (1, P, P, 2)
    params = O.to_torch(O.init_params(L, root, False, 2017))
    ref = O.forward(torch.tensor(X), params, L, root, False).numpy()
    assert rel(logits, ref) < TOL


@pytest.mark.parametrize("case", [(3, 64, True, 36, 2), (4, 64, True, 52, 1), (3, 64, False, 20, 2),
                                  # the production layer shapes (BASELINE.json configs[1..3]): CTA-pair
                                  # kernel at Cout 1024 / 2048, resident-weight halo kernel with the fused
                                  # pool at 760^2, split-K weight gradients over 760^2 pixels
                                  (6, 64, True, 388, 1)])
def test_backward_per_layer_teacher_forced(case):
    """Per-layer gradient parity inside the real network (the 2e-2 gate of BASELINE.json): every
    layer's backward kernels are fed the ORACLE's incoming gradient (bf16-storage oracle, masked
    like the device stores it) and their outputs -- weight / bias gradients and the gradient handed
    to the previous layer -- are compared with the oracle's.  Forcing the incoming gradient removes
    the chaotic amplification of rare arg-max / ReLU flips along the chain, so the comparison is
    tight for every layer, including the deepest ones."""
    from road_segmentation_unet_b200 import ops, unet
    L, root, dil, P, B = case
    S = unet.input_size_needed(P, L)
    params = make_params(L, root, dil)
    X, labels = synth(B, S, P)
    accs = {k: np.zeros_like(v) for k, v in params.items()}
    _, _, grads, _, _, acts = O.train_step(X, labels, params, accs, L, root, dil, 0.01, 0.9,
                                           want_acts=True, storage="bf16")
    net = unet.UNet(L, root, dil, B, S, params=params)
    net.grads.zero_()
    lab_dev = torch.tensor(labels).cuda()
    net.forward(torch.tensor(X).cuda(), lab_dev, keep=1.0)
    torch.cuda.synchronize()
    f = net.f

    def a_val(name):
        return acts[name].detach().numpy()

    # Teacher forcing of the forward state too: every saved activation is overwritten with the
    # oracle's (bf16-storage, hence exactly representable) value after checking that the device's
    # own forward agrees with it, so ReLU masks and pooling arg-maxes are identical on both sides
    # and each backward kernel is compared on identical operands.
    def force(buf, name, crop=None):
        ref = a_val(name)
        if crop is not None:
            o, t = crop
            ref = ref[:, o:o + t, o:o + t, :]
        assert rel(buf.float().cpu().numpy(), ref) < TOL, name
        buf.copy_(torch.tensor(np.ascontiguousarray(ref)).cuda().to(torch.bfloat16))
        assert np.array_equal(buf.float().cpu().numpy(), ref), name  # bf16-exact

    for i in range(L):
        force(net.A1[i], "conv_%d/relu1" % i)
        force(net.A2[i], "conv_%d/relu2" % i)
        if i < L - 1:
            force(net.Pool[i], "pool_%d" % i)
            if dil:
                t_i, o_i = net.up_size[L - 2 - i], net.dil_off[i]
                force(net.D1[i], "conv_dilut_%d/relu1" % i, (o_i, t_i + 4))
                force(net.D2[i], "conv_dilut_%d/relu2" % i, (o_i, t_i))
    for j in range(L - 1):
        force(net.U[j], "up_conv_%d" % j)
        force(net.C1[j], "conv_%d/relu1" % (L + j))
        force(net.C2[j], "conv_%d/relu2" % (L + j))
    # head on the forced activation (loss + gradients that leave it)
    net.grads.zero_()
    net.loss.zero_()
    ops.head(net._last, net.var("weight_output/kernel").view(-1, 2), net.var("weight_output/bias"),
             labels=lab_dev, probs=net.probs, loss=net.loss, dz=net.dC2[L - 2],
             dw=net.var("weight_output/kernel", "grads").view(-1, 2),
             db=net.var("weight_output/bias", "grads"))

    def a_grad(name, masked=True, crop=None):
        g = acts[name].grad.numpy()
        if masked:
            g = g * (a_val(name) > 0)
        if crop is not None:
            o, t = crop
            g = g[:, o:o + t, o:o + t, :]
        return g

    def put(buf, arr):
        buf.copy_(torch.tensor(np.ascontiguousarray(arr), dtype=torch.float32).cuda().to(torch.bfloat16))

    def got(t):
        return t.float().cpu().numpy()

    report = {}
    failures = []

    def check(tag, a, b, tol=TOL, same=None):
        """same: boolean array marking elements whose ReLU mask agrees between device and oracle
        (a sign flip of a near-zero activation toggles the WHOLE gradient of that element, which is
        activation noise, not a backward-kernel error; the agreement itself is asserted >= 99 %)."""
        if same is not None:
            assert same.mean() > 0.99, (tag, same.mean())
            a, b = a[same], b[same]
        e = rel(a, b)
        report[tag] = e
        if not e < tol:
            failures.append((tag, e))

    def same_mask(dev_act, name, crop=None):
        ref = a_val(name)
        if crop is not None:
            o, t = crop
            ref = ref[:, o:o + t, o:o + t, :]
        return (got(dev_act) > 0) == (ref > 0)

    def check_vars(prefix):
        for suffix in ("kernel", "bias"):
            n = prefix + "/" + suffix
            check("grad " + n, got(net.var(n, "grads")), grads[n])

    # head (free running: it only depends on the device's own forward)
    check_vars("weight_output")
    last = "conv_%d/relu2" % (2 * L - 2)
    check("dZ " + last, got(net.dC2[L - 2]), a_grad(last), same=same_mask(net.C2[L - 2], last))

    for j in range(L - 2, -1, -1):
        i = L - 2 - j
        fo, t = f[i], net.up_size[j]
        so = (net.skip_size[i] - t) // 2
        c1, c2 = net.convs["conv_%d/conv1" % (L + j)], net.convs["conv_%d/conv2" % (L + j)]
        up = net.convs["up_conv_%d" % j]
        n1, n2 = "conv_%d/relu1" % (L + j), "conv_%d/relu2" % (L + j)
        # conv2
        net.grads.zero_()
        put(net.dC2[j], a_grad(n2))
        net._conv_bwd(c2, [(net.C1[j], 0, 0)], net.dC2[j], net.dC1[j], mask=net.C1[j])
        check_vars(c2.name)
        check("dZ " + n1, got(net.dC1[j]), a_grad(n1), same=same_mask(net.C1[j], n1))
        # conv1 over the (never materialised) concat
        net.grads.zero_()
        put(net.dC1[j], a_grad(n1))
        srcs = [(net.A2[i], so, so)] + ([(net.D2[i], 0, 0)] if dil else []) + [(net.U[j], 0, 0)]
        net._conv_bwd(c1, srcs, net.dC1[j], net.dCat[j], mask=net.D2[i] if dil else None, mask_c0=fo)
        check_vars(c1.name)
        dcat = net.dCat[j]
        nparts = 3 if dil else 2
        check("d up_conv_%d" % j, got(dcat[..., (nparts - 1) * fo:]), a_grad("up_conv_%d" % j, masked=False))
        if dil:
            # (the ReluGrad of the dilated skip is fused into the kernel that writes the concat
            # gradient: channels [fo, 2 fo) arrive masked, and dD2 is a view of that slice)
            od = (net.in_size[i] - 8 - t) // 2
            assert net.dD2[i].data_ptr() == dcat[..., fo:2 * fo].data_ptr()
            check("dZ conv_dilut_%d/relu2" % i, got(net.dD2[i]), a_grad("conv_dilut_%d/relu2" % i, crop=(od, t)),
                  same=same_mask(net.D2[i], "conv_dilut_%d/relu2" % i, (od, t)))
        # transpose conv
        net.grads.zero_()
        d_up = dcat[..., (nparts - 1) * fo:]
        d_up.copy_(torch.tensor(a_grad("up_conv_%d" % j, masked=False)).cuda().to(torch.bfloat16))
        x_in = net._dec_in[j]
        ops.upconv2x2_wgrad(d_up, x_in, net.var(up.name + "/kernel", "grads").view(4 * up.cout, up.cin))
        ops.bias_grad(d_up, net.var(up.name + "/bias", "grads"))
        check_vars(up.name)
        src_name = "conv_%d/relu2" % (L + j - 1) if j > 0 else "conv_%d/relu2" % (L - 1)
        dst = net.dC2[j - 1] if j > 0 else net.dA2[L - 1]
        ops.upconv2x2_dgrad(d_up, up.w_dgrad, dst, mask=x_in)
        check("dZ " + src_name, got(dst), a_grad(src_name),
              same=same_mask(net.C2[j - 1] if j > 0 else net.A2[L - 1], src_name))

    for i in range(L - 1, -1, -1):
        reg1, reg2 = net.convs["conv_%d/conv1" % i], net.convs["conv_%d/conv2" % i]
        n1, n2 = "conv_%d/relu1" % i, "conv_%d/relu2" % i
        if i < L - 1:
            j = L - 2 - i
            t = net.up_size[j]
            so = (net.skip_size[i] - t) // 2
            put(net.dIn[i + 1], a_grad("pool_%d" % i, masked=False))
            # dCat[j] still holds the device's own concat gradient (skip part) from above
            ops.skip_grad(net.A2[i], net.dIn[i + 1], net.dCat[j][..., :f[i]], (so, so), net.dA2[i])
            # ties inside a pooling window route the gradient differently: compare on the rest
            y = a_val(n2)
            win = y.reshape(B, y.shape[1] // 2, 2, y.shape[2] // 2, 2, y.shape[3])
            mx = win.max(axis=(2, 4), keepdims=True)
            # (an all-zero window is a 4-way tie too, but ReluGrad zeroes it on both sides)
            uniq = np.broadcast_to(((win == mx).sum(axis=(2, 4), keepdims=True) == 1) | (mx == 0),
                                   win.shape).reshape(y.shape)
            assert np.array_equal(got(net.A2[i]), y)
            # positive bf16 ties are rare at the shallow levels but common among the small
            # activations of the deep ones (12 % of the windows at level 4 of the flagship net)
            assert uniq.mean() > 0.8
            check("dZ(skip) " + n2, got(net.dA2[i])[uniq], a_grad(n2)[uniq])
        net.grads.zero_()
        put(net.dA2[i], a_grad(n2))
        net._conv_bwd(reg2, [(net.A1[i], 0, 0)], net.dA2[i], net.dA1[i], mask=net.A1[i])
        check_vars(reg2.name)
        check("dZ " + n1, got(net.dA1[i]), a_grad(n1), same=same_mask(net.A1[i], n1))
        net.grads.zero_()
        put(net.dA1[i], a_grad(n1))
        if dil and i < L - 1:
            t = net.up_size[L - 2 - i]
            o2 = net.dil_off[i]
            d1 = net.convs["conv_dilut_%d/atrous_conv1" % i]
            d2 = net.convs["conv_dilut_%d/atrous_conv2" % i]
            put(net.dD2[i], a_grad("conv_dilut_%d/relu2" % i, crop=(o2, t)))
            net._conv_bwd(d2, [(net.D1[i], 0, 0)], net.dD2[i], net.dD1[i], mask=net.D1[i])
            check_vars(d2.name)
            check("dZ conv_dilut_%d/relu1" % i, got(net.dD1[i]), a_grad("conv_dilut_%d/relu1" % i, crop=(o2, t + 4)),
                  same=same_mask(net.D1[i], "conv_dilut_%d/relu1" % i, (o2, t + 4)))
            put(net.dD1[i], a_grad("conv_dilut_%d/relu1" % i, crop=(o2, t + 4)))
        if i > 0:
            net._conv_bwd(reg1, [(net.Pool[i - 1], 0, 0)], net.dA1[i], net.dIn[i])
            check_vars(reg1.name)
            if dil and i < L - 1:
                tt = net.D1[i].shape[1] + 4
                net._conv_bwd(d1, [(net.Pool[i - 1], o2, o2)], net.dD1[i],
                              net.dIn[i][:, o2:o2 + tt, o2:o2 + tt, :], accumulate=True)
                check_vars(d1.name)
            check("d pool_%d" % (i - 1), got(net.dIn[i]), a_grad("pool_%d" % (i - 1), masked=False))
        else:
            net._first_bwd(reg1, net.col, net.dcol, net.dA1[0], 1, 0, 0)
            check_vars(reg1.name)
            if dil and L > 1:
                net._first_bwd(d1, net.colD, net.dcolD, net.dD1[0], 2, o2, o2)
                check_vars(d1.name)
            check_vars("color_space_adjust")
    torch.cuda.synchronize()
    print("teacher-forced per-layer errors, worst:", sorted(report.items(), key=lambda kv: -kv[1])[:8])
    assert not failures, failures
    assert len(report) >= 8 * L


def test_flagship_free_running_batch4():
    """The production architecture (L=6, root 64, dilated, 764^2 -> 388^2) at batch 4, free-running
    forward + backward against the reference's precision (fp32 oracle), 2e-2 relative per tensor:
    every activation, the loss, and the weight / bias gradient of every layer above the bottom
    block.  The bottom block (conv_5/*, 18^2 / 16^2 pixels per patch, and up_conv_0 that reads it)
    is reported and held to the floor the bf16-storage oracle itself shows against fp32 there:
    with ~1 k pixels per channel a single max-pool arg-max flip moves those gradients by percent."""
    from road_segmentation_unet_b200 import unet
    L, root, dil, P, B = 6, 64, True, 388, 4
    S = unet.input_size_needed(P, L)
    params = make_params(L, root, dil)
    X, labels = synth(B, S, P, seed=77)
    accs = {k: np.zeros_like(v) for k, v in params.items()}
    loss32, probs32, grads32, _, _, _ = O.train_step(X, labels, params, accs, L, root, dil, 0.01, 0.9)
    _, _, grads16, _, _, _ = O.train_step(X, labels, params, accs, L, root, dil, 0.01, 0.9, storage="bf16")
    net = unet.UNet(L, root, dil, B, S, params=params)
    net.grads.zero_()
    net.forward(torch.tensor(X).cuda(), torch.tensor(labels).cuda(), keep=1.0)
    net.backward()
    torch.cuda.synchronize()
    assert abs(net.loss.item() - loss32) < TOL * abs(loss32)
    assert rel(net.probs.cpu().numpy(), probs32) < TOL
    live = net.live_variables()
    e32 = {n: rel(net.var(n, "grads").cpu().numpy(), grads32[n]) for n in live}
    floor = {n: rel(grads16[n], grads32[n]) for n in live}
    table = sorted(((n, e32[n], floor[n]) for n in live), key=lambda r: -r[1])
    print("batch 4, gradients vs fp32 oracle (device, bf16-storage oracle):")
    for n, e, fl in table:
        print("   %-40s %.4f  %.4f" % (n, e, fl))
    within = [n for n in live if e32[n] < TOL]
    deep = [n for n in live if n not in within]
    print("%d of %d tensors within %.0e; beyond it: %s" % (len(within), len(live), TOL, sorted(set(
        n.rsplit("/", 1)[0] for n in deep))))
    # Where bf16 STORAGE itself moves the free-running gradient further than 2e-2 from fp32 (the
    # bf16-storage restatement of the reference shows the same deviation: rounding noise of ~40
    # chained bf16 tensors plus arg-max / ReLU flips, concentrated on the few-pixel deep levels),
    # the device must be no further from fp32 than that restatement is.  Each layer on its own --
    # same incoming gradient -- is held to 2e-2 at these shapes by
    # test_backward_per_layer_teacher_forced[(6, 64, True, 388, 1)] (observed 2e-3).
    bad = {n: (e32[n], floor[n]) for n in deep if not e32[n] < 1.25 * floor[n] + 2e-3}
    assert not bad, bad
    # the shallow, pixel-rich half of the network is inside the plain gate
    shallow = [n for n in live if n.startswith(("color_space_adjust", "conv_0/", "conv_1/", "conv_dilut_0/",
                                                "conv_dilut_1/", "conv_9/", "conv_10/", "up_conv_3", "up_conv_4",
                                                "weight_output"))]
    assert all(e32[n] < TOL for n in shallow), {n: e32[n] for n in shallow if not e32[n] < TOL}


def test_large_tile_forward_parity():
    """BASELINE.json configs[4], tiles as patches: the 6-layer dilated root-64 U-Net on a 1404^2
    input (P = 1028, the valid patch size next to 1024) -- every activation, the logits and
    P(road) of the device forward pass within 2e-2 of the fp32 oracle."""
    from road_segmentation_unet_b200 import unet
    L, root, dil, P, B = 6, 64, True, 1028, 1
    S = unet.input_size_needed(P, L)
    assert S == 1404
    params = make_params(L, root, dil)
    X, _ = synth(B, S, P, seed=5)
    acts = {}
    with torch.no_grad():
        logits = O.forward(torch.tensor(X), O.to_torch(params), L, root, dil, acts=acts)
        probs32 = torch.softmax(logits, dim=3)[..., 1].numpy()
    net = unet.UNet(L, root, dil, B, S, params=params, training=False)
    assert net.P == P
    net.forward(torch.tensor(X).cuda(), keep=1.0, want_logits=True)
    torch.cuda.synchronize()
    errs = {"logits": rel(net.logits.cpu().numpy(), logits.numpy()), "probs": rel(net.probs.cpu().numpy(), probs32)}
    for i in range(L):
        errs["conv_%d/relu2" % i] = rel(net.A2[i].float().cpu().numpy(), acts["conv_%d/relu2" % i].numpy())
        if i < L - 1:
            ref = O.center_crop(acts["conv_dilut_%d/relu2" % i], net.up_size[L - 2 - i])
            errs["conv_dilut_%d/relu2" % i] = rel(net.D2[i].float().cpu().numpy(), ref.numpy())
    for j in range(L - 1):
        errs["conv_%d/relu2" % (L + j)] = rel(net.C2[j].float().cpu().numpy(), acts["conv_%d/relu2" % (L + j)].numpy())
    print("P = 1028 activations vs fp32, worst:", sorted(errs.items(), key=lambda kv: -kv[1])[:4])
    assert max(errs.values()) < TOL, errs
