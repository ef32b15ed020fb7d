"""Pins the CPU oracle: images_oracle against golden vectors made by the reference's own
src/images.py (tests/golden/make_golden.py), unet_oracle against an independent NumPy loop nest
and against the properties the reference documents."""
import os

import numpy as np
import pytest
import torch

from oracle import images_oracle as IO
from oracle import unet_oracle as UO

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "images_golden.npz"))


def test_mirror_border_golden():
    assert np.array_equal(IO.mirror_border(G["mirror_in4"], 4), G["mirror_out4_n4"])
    assert np.array_equal(IO.mirror_border(G["mirror_in3"], 7), G["mirror_out3_n7"])
    # [1,0 | 0,1,2,3,4 | 4,3]
    row = np.arange(5, dtype=np.float32).reshape(1, 1, 5)
    assert IO.mirror_border(np.repeat(row, 5, axis=1), 2)[0, 2].tolist() == [1, 0, 0, 1, 2, 3, 4, 4, 3]


def test_extract_patches_golden():
    x = G["extract_in"]
    assert np.array_equal(IO.extract_patches(x, 8, stride=4), G["extract_p8_s4"])
    assert np.array_equal(IO.extract_patches(x, 10), G["extract_p10_nostride"])
    assert np.array_equal(IO.extract_patches(x[..., 0], 12, stride=8), G["extract3d_p12_s8"])
    assert IO.extract_patches(x, 8, stride=4).dtype == np.float64
    with pytest.raises(AssertionError):
        IO.extract_patches(x, 8, stride=5)


def test_reference_shape_tests():
    """The shape assertions of the reference's own src/test_images.py:11-121."""
    imgs = np.zeros((2, 608, 608, 3), dtype=np.float32)
    assert IO.extract_patches(imgs, 128, 16).shape == (2 * 31 * 31, 128, 128, 3)
    assert IO.extract_patches(imgs[:1], 32).shape == (361, 32, 32, 3)
    p = IO.extract_patches(np.zeros((1, 400, 400, 3), dtype=np.float32), 80)
    assert p.shape == (25, 80, 80, 3)
    assert IO.images_from_patches(p.reshape(1, 25, 80, 80, 3)).shape == (1, 400, 400, 3)


def test_images_from_patches_golden():
    assert np.array_equal(IO.images_from_patches(G["from_patches_in"], stride=4), G["from_patches_s4"])
    assert np.array_equal(IO.images_from_patches(G["from_patches_in"]), G["from_patches_nostride"])
    assert np.array_equal(IO.images_from_patches(G["from_patches_in64"], stride=3), G["from_patches64_s3"])
    x = np.random.RandomState(0).rand(2, 20, 20, 3)
    pt = IO.extract_patches(x, 8, stride=4).reshape(2, 16, 8, 8, 3)
    assert np.abs(IO.images_from_patches(pt, stride=4) - x).max() == 0.0  # exact round trip


def test_ensemble_golden():
    assert np.array_equal(IO.image_augmentation_ensemble(G["ens_in"]), G["ens_out"])
    assert np.array_equal(IO.invert_image_augmentation_ensemble(G["inv_in"]), G["inv_out"])
    m = np.random.RandomState(1).rand(3, 9, 9, 1)
    rt = IO.invert_image_augmentation_ensemble(IO.image_augmentation_ensemble(m))
    assert np.abs(rt - m).max() < 1e-15
    # flip_ud x rot90^k enumerates the 8 elements of D4, variants of the ensemble are 6 of them
    x = np.arange(16.0).reshape(4, 4)
    assert len({IO.d4(x, op).tobytes() for op in range(8)}) == 8
    ens_ops = [0, 4 | 2, 4, 1, 2, 3]
    e = IO.image_augmentation_ensemble(x.reshape(1, 4, 4, 1))
    for v, op in enumerate(ens_ops):
        assert np.array_equal(e[v, :, :, 0], IO.d4(x, op))


def test_crop_golden():
    assert np.array_equal(IO.crop_imgs(G["crop_in"], 8), G["crop_out8"])


@pytest.mark.parametrize("angle", [15, 30, 45, 60, 75, 90])
def test_rotate_golden(angle):
    got = IO.rotate_nn(G["rot_in"], angle)
    ref = G["rot_%d" % angle]
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)


def test_expand_and_rotate_golden():
    assert np.array_equal(IO.expand_and_rotate(G["expand_in"], [0, 15, 45, 75], 6), G["expand_off6"])
    assert np.array_equal(IO.expand_and_rotate(G["expand_in"][..., 0], [30, 60], 0), G["expand3d_off0"])


def test_quantize_golden():
    assert np.array_equal(IO.quantize_mask(G["quant_in"], 0.25, 16), G["quant_out"])
    assert IO.patch_f1(G["quant_in"], G["quant_in"]) == 1.0


# ------------------------------------------------------------------ U-Net oracle
def test_scoring_and_visual_golden():
    """labels, submission.csv text, overlays, confusion / error images."""
    assert np.array_equal(IO.labels_for_patches(G["labels_in"]), G["labels_out"])
    assert IO.submission_rows(G["csv_in"], 16) == bytes(G["csv_text"]).decode()
    for fade, key in ((0.95, "overlay_out_095"), (0.4, "overlay_out_040")):
        assert np.array_equal(IO.overlays(G["overlay_img"], G["overlay_mask"], fade), G[key])
    assert np.array_equal(IO.overlap_pred_true(G["confusion_pred"], G["confusion_true"]), G["confusion_out"])
    assert np.array_equal(IO.overlapp_error(G["confusion_pred"], G["confusion_true"]), G["error_out"])
    # the on-disk format of the reference's own submissions/*/submission.csv
    head = bytes(G["csv_reference_head"]).decode().splitlines()
    assert head[0] == "id,prediction"
    assert [r.split(",")[0] for r in head[1:]] == ["001_0_%d" % (16 * i) for i in range(38)] + ["001_16_0"]
    assert IO.submission_rows(np.zeros((1, 608, 608)), 16).splitlines()[:40] == \
        [head[0]] + [r.split(",")[0] + ",0" for r in head[1:]]


def submission_masks(labels):
    """[50, 38, 38] labels indexed [image, x cell, y cell] -> the quantised 608 x 608 masks."""
    return np.kron(labels.transpose(0, 2, 1), np.ones((16, 16), np.uint8)).astype(np.float64)


def test_reference_submission_known_answers():
    """The reference's OWN committed outputs (submissions/*/images_NNN.png + submission.csv, written
    by tf_aerial_images.py:446-458 from data/test/*.png; tests/golden/make_submission_golden.py):
    the overlay of the real test image with the quantised mask is reproduced byte for byte, and the
    masks the overlays locate give back every run's submission.csv byte for byte."""
    import hashlib
    S = np.load(os.path.join(os.path.dirname(__file__), "golden", "submission_golden.npz"))
    side = S["crop_rgb"].shape[1]
    for r in range(S["labels"].shape[0]):
        masks = submission_masks(S["labels"][r])
        text = IO.submission_rows(masks, 16)
        assert hashlib.sha256(text.encode()).hexdigest() == str(S["csv_sha256"][r]), str(S["runs"][r])
        if r == 0:
            assert text.startswith(bytes(S["csv_head"]).decode())
        assert np.array_equal(IO.quantize_mask(masks[:4, :, :, None], 0.25, 16)[..., 0], masks[:4])
        if r < S["crop_overlay"].shape[0]:
            for j, (k, top, left) in enumerate(S["crops"]):
                img = S["crop_rgb"][j:j + 1].astype(np.float32) / np.float32(255)  # = mpimg.imread
                m = masks[k:k + 1, top:top + side, left:left + side]
                assert 0.02 < m.mean() < 0.98  # the window shows road and background
                assert np.array_equal(IO.overlays(img, m, fade=0.4)[0], S["crop_overlay"][r, j])
                # the transposed mask is NOT what the reference drew: the test pins the cell order
                assert not np.array_equal(IO.overlays(img, m.transpose(0, 2, 1), fade=0.4)[0],
                                          S["crop_overlay"][r, j])


def test_input_size_needed():
    """unet.py:100-115; closed form S = P + 12 * 2^(L-1) - 8 (report/report.tex:50)."""
    assert UO.input_size_needed(388, 4) == 476
    assert UO.input_size_needed(388, 5) == 572
    assert UO.input_size_needed(388, 6) == 764
    for L in (2, 3, 4, 5, 6):
        assert UO.input_size_needed(2 ** L * 3 + 4, L) == 2 ** L * 3 + 4 + 12 * 2 ** (L - 1) - 8
    with pytest.raises(AssertionError):
        UO.input_size_needed(390, 4)


def test_variable_inventory():
    """Parameter counts of SURVEY.md appendix B / report.tex:50."""
    n = lambda L, r, d: sum(int(np.prod(s)) for s in UO.variable_shapes(L, r, d).values())
    assert n(6, 64, True) == 212403278
    assert n(4, 64, False) == 7697422
    assert n(5, 64, False) == 31031822
    dead = UO.dead_variables(6, True)
    shapes = UO.variable_shapes(6, 64, True)
    assert n(6, 64, True) - sum(int(np.prod(shapes[k])) for k in dead) == 155776078


def test_conv_against_loop_nest():
    rs = np.random.RandomState(3)
    x = rs.randn(2, 9, 9, 5).astype(np.float32)
    w = rs.randn(3, 3, 5, 4).astype(np.float32)
    b = rs.randn(4).astype(np.float32)
    for d in (1, 2):
        got = UO.conv2d_valid(torch.tensor(x), torch.tensor(w), torch.tensor(b), d).numpy()
        assert np.allclose(got, UO.conv2d_valid_loops(x, w, b, d), atol=1e-4)
    wt = rs.randn(2, 2, 6, 5).astype(np.float32)
    bt = rs.randn(6).astype(np.float32)
    got = UO.conv2d_transpose_2x2(torch.tensor(x), torch.tensor(wt), torch.tensor(bt)).numpy()
    assert np.allclose(got, UO.conv2d_transpose_2x2_loops(x, wt, bt), atol=1e-4)


@pytest.mark.parametrize("dilated", [False, True])
def test_forward_shapes_and_step(dilated):
    L, root, P = 3, 8, 20
    S = UO.input_size_needed(P, L)
    params = UO.init_params(L, root, dilated, seed=2017)
    rs = np.random.RandomState(0)
    X = rs.rand(2, S, S, 3).astype(np.float32)
    labels = (rs.rand(2, P, P) < 0.3).astype(np.int64)
    accs = {k: np.zeros_like(v) for k, v in params.items()}
    loss, probs, grads, new_p, new_a, acts = UO.train_step(
        X, labels, params, accs, L, root, dilated, lr=0.01, momentum=0.9, want_acts=True)
    assert acts["logits"].shape == (2, P, P, 2)
    assert probs.shape == (2, P, P) and 0 < loss < 5
    for k in UO.dead_variables(L, dilated):
        assert grads[k] is None and np.array_equal(new_p[k], params[k])
    k = "conv_0/conv1/kernel"
    assert np.allclose(new_a[k], grads[k]) and np.allclose(new_p[k], params[k] - 0.01 * grads[k])
    # fp64 run agrees with fp32 run (oracle self-consistency)
    loss64 = UO.train_step(X, labels, params, accs, L, root, dilated, 0.01, 0.9, dtype=torch.float64)[0]
    assert abs(loss64 - loss) < 1e-5


def test_learning_rate_staircase():
    assert UO.learning_rate(0.01, 999) == 0.01
    assert abs(UO.learning_rate(0.01, 1000) - 0.0095) < 1e-12
    assert abs(UO.learning_rate(0.01, 2500) - 0.01 * 0.95 ** 2) < 1e-12


@pytest.mark.parametrize("dilated", [False, True])
def test_shared_window_principle(dilated):
    """The property behind ConvolutionalModel._predict_shared, checked on the CPU oracle in fp64:
    sliding windows whose origins differ by a multiple of the pooling period 2^(L-1) are crops
    of ONE forward pass over the enlarged window (valid convolutions, aligned pooling, size-
    independent centre-crop offsets) -- and a shift that is not a multiple of the period is not."""
    L, root, P = 3, 8, 20
    S = UO.input_size_needed(P, L)            # 60
    period = 2 ** (L - 1)                     # 4
    q = 12                                    # lcm(stride 12, period 4): windows 12 pixels apart
    params = {k: torch.tensor(v, dtype=torch.float64) for k, v in UO.init_params(L, root, dilated, 2017).items()}
    rs = np.random.RandomState(1)
    img = torch.tensor(rs.rand(1, S + 2 * q, S + 2 * q, 3))
    with torch.no_grad():
        big = UO.forward(img, params, L, root, dilated)          # enlarged window: 3 x 3 positions
        assert big.shape[1] == P + 2 * q
        for jy in range(3):
            for jx in range(3):
                win = img[:, jy * q:jy * q + S, jx * q:jx * q + S]
                one = UO.forward(win, params, L, root, dilated)
                cut = big[:, jy * q:jy * q + P, jx * q:jx * q + P]
                assert float((one - cut).abs().max()) < 1e-12, (jy, jx)
        # a window shifted by half a period sees other pooling windows: not a crop of `big`
        off = period // 2
        one = UO.forward(img[:, off:off + S, :S], params, L, root, dilated)
        assert float((one - big[:, off:off + P, :P]).abs().max()) > 1e-6
