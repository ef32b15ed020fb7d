"""End-to-end parity of the B200 U-Net engine against the CPU oracle (oracle/unet_oracle.py):
same glorot weights (seed 2017), same synthetic inputs, one forward + backward + momentum step.
Gate (BASELINE.json north_star): per-layer activations and gradients within 2e-2 relative."""
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


CASES = [
    # (L, root, dilated, P, B)
    (3, 64, False, 20, 2),
    (3, 64, True, 36, 2),
    (4, 64, True, 52, 1),
]


@pytest.mark.parametrize("case", CASES)
def test_train_step_parity(case):
    from road_segmentation_unet_b200 import unet
    L, root, dil, P, B = case
    S = unet.input_size_needed(P, L)
    assert S == O.input_size_needed(P, L)
    params = O.init_params(L, root, dil, seed=2017)
    # non-zero biases so that bias gradients and the bias path are exercised
    rs = np.random.RandomState(5)
    for k in params:
        if k.endswith("bias"):
            params[k] = (0.05 * rs.randn(*params[k].shape)).astype(np.float32)
    X, labels = synth(B, S, P)
    accs = {k: np.zeros_like(v) for k, v in params.items()}
    loss_ref, probs_ref, grads_ref, new_p, new_a, acts = O.train_step(
        X, labels, params, accs, L, root, dil, lr=0.01, momentum=0.9, want_acts=True)

    net = unet.UNet(L, root, dil, B, S, params=params)
    assert net.P == P
    # the engine draws the same glorot weights by itself
    own = unet.glorot_init(L, root, dil, 2017)
    ref_init = O.init_params(L, root, dil, 2017)
    for k in own:
        assert np.array_equal(own[k], ref_init[k]), k
    xd = torch.tensor(X).cuda()
    ld = torch.tensor(labels).cuda()
    net.grads.zero_()
    net.forward(xd, ld, keep=1.0, want_logits=True)
    net.backward()
    torch.cuda.synchronize()

    errs = {}
    errs["logits"] = rel(net.logits.cpu().numpy(), acts["logits"].detach().numpy())
    errs["probs"] = rel(net.probs.cpu().numpy(), probs_ref)
    for i in range(L):
        errs["conv_%d/relu1" % i] = rel(net.A1[i].float().cpu().numpy(), acts["conv_%d/relu1" % i].detach().numpy())
        errs["conv_%d/relu2" % i] = rel(net.A2[i].float().cpu().numpy(), acts["conv_%d/relu2" % i].detach().numpy())
        if dil and i < L - 1:
            t = net.up_size[L - 2 - i]
            full = acts["conv_dilut_%d/relu2" % i].detach()
            errs["conv_dilut_%d/relu2" % i] = rel(net.D2[i].float().cpu().numpy(), O.center_crop(full, t).numpy())
    for j in range(L - 1):
        errs["up_conv_%d" % j] = rel(net.U[j].float().cpu().numpy(), acts["up_conv_%d" % j].detach().numpy())
        errs["conv_%d/relu2" % (L + j)] = rel(net.C2[j].float().cpu().numpy(),
                                             acts["conv_%d/relu2" % (L + j)].detach().numpy())
    assert abs(net.loss.item() - loss_ref) < TOL * abs(loss_ref), (net.loss.item(), loss_ref)
    # activation gradients (ReLU-masked dZ) of a few layers
    g_act = acts["conv_%d/relu2" % (L - 1)].grad.numpy() * (acts["conv_%d/relu2" % (L - 1)].detach().numpy() > 0)
    errs["d conv_%d/relu2" % (L - 1)] = rel(net.dA2[L - 1].float().cpu().numpy(), g_act)
    g_act = acts["conv_0/relu1"].grad.numpy() * (acts["conv_0/relu1"].detach().numpy() > 0)
    errs["d conv_0/relu1"] = rel(net.dA1[0].float().cpu().numpy(), g_act)
    # weight gradients of every live variable
    live = net.live_variables()
    for name in live:
        errs["grad " + name] = rel(net.var(name, "grads").cpu().numpy(), grads_ref[name])
    for name in O.dead_variables(L, dil):
        assert grads_ref[name] is None
        assert float(net.var(name, "grads").abs().max()) == 0.0
    bad = {k: v for k, v in errs.items() if not v < TOL}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    print("worst:", worst)
    assert not bad, bad
    # momentum update (tf.train.MomentumOptimizer): same gradients -> same delta
    net.apply_gradients(0.01, 0.9)
    torch.cuda.synchronize()
    for name in live:
        dw = net.var(name).cpu().numpy() - params[name]
        assert rel(dw, new_p[name] - params[name]) < TOL, name
    assert net.global_step == 1


def test_dropout_parity():
    """tf.nn.dropout sites (unet.py:29-30, 64-65): the engine's masks are fed to the oracle."""
    from road_segmentation_unet_b200 import unet, ops
    L, root, dil, P, B, keep = 3, 64, True, 36, 2, 0.8
    S = unet.input_size_needed(P, L)
    params = O.init_params(L, root, dil, seed=2017)
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
        scales.append(torch.tensor(m.reshape(shp)))
    accs = {k: np.zeros_like(v) for k, v in params.items()}
    loss_ref, probs_ref, grads_ref, _, _, _ = O.train_step(
        X, labels, params, accs, L, root, dil, 0.01, 0.9, dropout_scales=scales)
    assert abs(net.loss.item() - loss_ref) < TOL * abs(loss_ref)
    assert rel(net.probs.cpu().numpy(), probs_ref) < TOL
    for name in net.live_variables():
        assert rel(net.var(name, "grads").cpu().numpy(), grads_ref[name]) < TOL, name


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
