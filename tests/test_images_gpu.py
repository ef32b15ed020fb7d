"""Geometry kernels (csrc/geometry.cu through road_segmentation_unet_b200.images) against the
golden vectors produced by the reference's own src/images.py and against the NumPy oracle.
Bit-exact for padding, flips, 90-degree rotations, crops and patch indexing (BASELINE.json
north_star); <= 1e-6 for the two averaging helpers (fp64 accumulation on the device)."""
import os

import numpy as np
import pytest
import torch

from oracle import images_oracle as IO

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "images_golden.npz"))


@pytest.fixture(scope="module")
def images():
    from road_segmentation_unet_b200 import images as _images
    return _images


def test_mirror_border(images):
    out = images.mirror_border(G["mirror_in4"], 4)
    assert out.dtype == G["mirror_out4_n4"].dtype and np.array_equal(out, G["mirror_out4_n4"])
    assert np.array_equal(images.mirror_border(G["mirror_in3"], 7), G["mirror_out3_n7"])
    rs = np.random.RandomState(0)
    x64 = rs.rand(2, 33, 33, 3)  # float64 moves bit-exactly as word pairs
    assert np.array_equal(images.mirror_border(x64, 20), IO.mirror_border(x64, 20))
    # full-size: 6 x 604^2 x 3 -> 980^2 (config 3), pad larger than nothing special
    x = rs.rand(6, 604, 604, 3).astype(np.float32)
    assert np.array_equal(images.mirror_border(x, 188), np.pad(x, ((0, 0), (188, 188), (188, 188), (0, 0)), "symmetric"))
    # pad larger than the image (periodic reflection)
    small = rs.rand(1, 5, 5, 1).astype(np.float32)
    assert np.array_equal(images.mirror_border(small, 12), np.pad(small, ((0, 0), (12, 12), (12, 12), (0, 0)), "symmetric"))


def test_extract_patches(images):
    x = G["extract_in"]
    for got, ref in ((images.extract_patches(x, 8, stride=4), G["extract_p8_s4"]),
                     (images.extract_patches(x, 10), G["extract_p10_nostride"]),
                     (images.extract_patches(x[..., 0], 12, stride=8), G["extract3d_p12_s8"])):
        assert got.dtype == np.float64 and got.shape == ref.shape and np.array_equal(got, ref)
    with pytest.raises(AssertionError):
        images.extract_patches(x, 8, stride=5)
    with pytest.raises(AssertionError):
        images.extract_patches(np.zeros((1, 8, 9, 3), np.float32), 4)
    # the reference's own shape tests (src/test_images.py:11-45)
    imgs = np.random.RandomState(1).rand(2, 608, 608, 3).astype(np.float32)
    p = images.extract_patches(imgs, 128, 16)
    assert p.shape == (2 * 31 * 31, 128, 128, 3)
    # x-outer ordering: patch k of image n sits at (x = k // side * stride, y = k % side * stride)
    k = 31 * 5 + 7
    assert np.array_equal(p[k], imgs[0, 7 * 16:7 * 16 + 128, 5 * 16:5 * 16 + 128])
    assert np.array_equal(p[961 + k], imgs[1, 7 * 16:7 * 16 + 128, 5 * 16:5 * 16 + 128])
    # device-level range extraction = a slice of the full list
    xd = torch.tensor(imgs).cuda()
    part = images.extract_patches_dev(xd, 128, 16, k_begin=950, k_count=40).cpu().numpy()
    assert np.array_equal(part, p[950:990].astype(np.float32))


def test_images_from_patches(images):
    for key_in, stride, key_out in (("from_patches_in", 4, "from_patches_s4"),
                                    ("from_patches_in", None, "from_patches_nostride"),
                                    ("from_patches_in64", 3, "from_patches64_s3")):
        got = images.images_from_patches(G[key_in], stride=stride)
        assert got.shape == G[key_out].shape
        assert np.abs(got - G[key_out]).max() <= 1e-6
    # round trip with extract_patches is exact
    x = np.random.RandomState(2).rand(2, 44, 44, 1).astype(np.float32)
    pt = images.extract_patches(x, 20, stride=12).reshape(2, 9, 20, 20, 1)
    assert np.abs(images.images_from_patches(pt, stride=12) - x).max() == 0.0
    # config-3 geometry at reduced patch count: P=388, stride 12, side 5
    pr = np.random.RandomState(3).rand(1, 25, 388, 388, 1).astype(np.float32)
    got = images.images_from_patches(pr, stride=12)
    ref = IO.images_from_patches(pr.astype(np.float64), stride=12)
    assert got.shape == (1, 436, 436, 1) and np.abs(got - ref).max() <= 1e-6
    # sharded form: partial sums of two slices add up to the normalised result
    pd = torch.tensor(pr.reshape(25, 388, 388, 1)).cuda()
    a = images.images_from_patches_dev(pd[:11], 1, 5, 12, 0, 11, normalize=False)
    b = images.images_from_patches_dev(pd[11:], 1, 5, 12, 11, 14, normalize=False)
    hits = IO.images_from_patches(np.ones_like(pr, dtype=np.float64) * 0 + 1, stride=12)  # all ones
    cnt = np.zeros((436, 436))
    for kx in range(5):
        for ky in range(5):
            cnt[ky * 12:ky * 12 + 388, kx * 12:kx * 12 + 388] += 1
    assert np.abs((a + b).cpu().numpy()[0, :, :, 0] / cnt - ref[0, :, :, 0]).max() <= 1e-6
    assert hits.shape == ref.shape
    # 16-byte-lane path with channels (P*C, stride*C multiples of 4), several images, three
    # slices that cut through patch columns and images; and the same data through the scalar
    # path (patch pointer 4 bytes off a 16-byte boundary)
    pc = np.random.RandomState(6).rand(3, 16, 40, 40, 3).astype(np.float32)
    refc = IO.images_from_patches(pc.astype(np.float64), stride=8)
    assert np.abs(images.images_from_patches(pc, stride=8) - refc).max() <= 1e-6
    flat = torch.empty(pc.size + 1, dtype=torch.float32, device="cuda")
    for shift in (0, 1):
        pdc = flat[shift:shift + pc.size].view(48, 40, 40, 3)
        pdc.copy_(torch.tensor(pc.reshape(48, 40, 40, 3)))
        full = images.images_from_patches_dev(pdc, 3, 4, 8).cpu().numpy()
        assert np.abs(full - refc).max() <= 1e-6
        parts = [images.images_from_patches_dev(pdc[lo:hi], 3, 4, 8, lo, hi - lo, normalize=False)
                 for lo, hi in ((0, 7), (7, 30), (30, 48))]
        # the sharded pieces only line up with 16 bytes when the slice starts do; both are legal
        hits = images.images_from_patches_dev(torch.ones_like(pdc), 3, 4, 8, normalize=False)
        got = ((parts[0] + parts[1] + parts[2]) / hits).cpu().numpy()
        assert np.abs(got - refc).max() <= 1e-6
    # odd geometry (nothing divisible by 4): scalar lanes
    po = np.random.RandomState(7).rand(2, 9, 21, 21, 1).astype(np.float32)
    refo = IO.images_from_patches(po.astype(np.float64), stride=5)
    assert np.abs(images.images_from_patches(po, stride=5) - refo).max() <= 1e-6


def test_ensemble(images):
    got = images.image_augmentation_ensemble(G["ens_in"])
    assert got.dtype == np.float64 and np.array_equal(got, G["ens_out"])
    inv = images.invert_image_augmentation_ensemble(G["inv_in"])
    assert np.abs(inv - G["inv_out"]).max() <= 1e-6
    x = np.random.RandomState(4).rand(2, 604, 604, 3).astype(np.float32)
    assert np.array_equal(images.image_augmentation_ensemble(x), IO.image_augmentation_ensemble(x))
    m = np.random.RandomState(5).rand(2, 37, 37, 1).astype(np.float32)
    rt = images.invert_image_augmentation_ensemble(images.image_augmentation_ensemble(m))
    assert np.abs(rt - m).max() <= 1e-6
    for side in (604, 68):  # 16-byte path incl. partial edge tiles, random (not round-trip) masks
        big = np.random.RandomState(6).rand(12, side, side, 1).astype(np.float32)
        want = IO.invert_image_augmentation_ensemble(big.astype(np.float64).copy())
        assert np.abs(images.invert_image_augmentation_ensemble(big) - want).max() <= 1e-6


@pytest.mark.parametrize("dtype,shape", [(torch.float32, (9, 50, 50, 3)), (torch.uint8, (8, 37, 37)),
                                          (torch.float32, (8, 764, 764, 3)),
                                          # 16-byte / 4-byte vector paths incl. partial edge tiles
                                          (torch.float32, (8, 604, 604, 3)), (torch.float32, (8, 604, 604)),
                                          (torch.float32, (16, 68, 68, 1)), (torch.uint8, (16, 388, 388)),
                                          (torch.uint8, (8, 100, 100)), (torch.float32, (8, 36, 36, 3))])
def test_d4_transform(images, dtype, shape):
    """All 8 dihedral elements, bit-exact: out = rot90(flipud(x) if op & 4 else x, k = op & 3)
    (tf.image.flip_up_down / rot90 of tf_aerial_images.py:173-210, np.flip / np.rot90 of
    images.py:376-417)."""
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(shape, generator=g) * 255).to(dtype)
    ops_np = (np.arange(shape[0]) % 8).astype(np.uint8)
    out = images.d4_transform_dev(x.cuda(), torch.tensor(ops_np).cuda()).cpu().numpy()
    for i, op in enumerate(ops_np):
        assert np.array_equal(out[i], IO.d4(x[i].numpy(), int(op))), (i, op)


def test_crop(images):
    assert np.array_equal(images.crop_imgs(G["crop_in"], 8), G["crop_out8"])
    with pytest.raises(AssertionError):
        images.crop_imgs(G["crop_in"], 7)


@pytest.mark.parametrize("angle", [15, 30, 45, 60, 75, 90])
def test_rotate(images, angle):
    """Nearest-neighbour rotation against SciPy (golden) -- exact on the golden input; on a larger
    random image half-integer ties at 30/60 degrees may resolve differently."""
    got = images.rotate_imgs(G["rot_in"], angle)
    ref = G["rot_%d" % angle]
    assert got.shape == ref.shape and got.dtype == ref.dtype
    assert np.array_equal(got, ref)
    x = np.random.RandomState(6).rand(1, 200, 200, 3).astype(np.float32)
    big = images.rotate_imgs(x, angle)
    ref_big = IO.rotate_nn(x, angle)
    assert big.shape == ref_big.shape
    mismatch = float((big != ref_big).any(axis=-1).mean())
    assert mismatch <= (1e-3 if angle in (30, 60) else 0.0), mismatch


def test_expand_and_rotate(images):
    got = images.expand_and_rotate(G["expand_in"], [0, 15, 45, 75], 6)
    assert got.dtype == np.float64 and np.array_equal(got, G["expand_off6"])
    assert np.array_equal(images.expand_and_rotate(G["expand_in"][..., 0], [30, 60], 0), G["expand3d_off0"])
    # training-prep geometry of the README config on one synthetic 400^2 image (offset 188)
    x = np.random.RandomState(7).rand(1, 400, 400, 3).astype(np.float32)
    got = images.expand_and_rotate(x, [15, 45, 75], 188)
    ref = IO.expand_and_rotate(x, [15, 45, 75], 188)
    assert got.shape == (3, 776, 776, 3)
    assert float((got != ref).any(axis=-1).mean()) == 0.0


def test_scoring_rules(images, tmp_path):
    """rsu_patch_vote behind quantize_mask / labels_for_patches / save_submission_csv
    (images.py:88-99, 206-237, 256-266): bit-exact against the reference's own outputs."""
    for dt in (np.float64, np.float32):
        q = images.quantize_mask(G["quant_in"].astype(dt), 0.25, 16)
        assert q.dtype == dt and q.shape == G["quant_in"].shape
        assert np.array_equal(q, G["quant_out"].astype(dt) if dt == np.float32 else G["quant_out"])
    # a value just below 0.5 in fp64 must not be rounded up through fp32
    edge = np.full((1, 16, 16, 1), np.nextafter(0.5, 0.0))
    assert images.quantize_mask(edge, 0.25, 16).max() == 0.0
    assert images.quantize_mask(edge.astype(np.float32), 0.25, 16).min() == 1.0
    # cells clipped at the border when the patch size does not divide the image (reference slices)
    odd = np.random.RandomState(8).rand(2, 40, 40, 1)
    assert np.array_equal(images.quantize_mask(odd, 0.25, 16), IO.quantize_mask(odd, 0.25, 16))
    assert np.array_equal(images.labels_for_patches(G["labels_in"]), G["labels_out"])
    p = np.zeros((3, 4, 4))
    p[1] = 1.0
    p[2, :1] = 1.0  # mean 0.25 is NOT > 0.25
    assert images.labels_for_patches(p).tolist() == [0, 1, 0]
    # label grid [image, x cell, y cell] and the csv text
    lab = images.patch_labels(G["csv_in"], 16)
    want = np.array([[[G["csv_in"][n, y:y + 16, x:x + 16, 0].mean() > 0.25 for y in range(0, 48, 16)]
                      for x in range(0, 48, 16)] for n in range(2)])
    assert lab.dtype == np.int64 and np.array_equal(lab, want)
    images.save_submission_csv(G["csv_in"], str(tmp_path / "sub"), 16)
    with open(tmp_path / "sub" / "submission.csv", "rb") as f:
        assert f.read() == bytes(G["csv_text"])
    # 608^2 masks -> 38 x 38 labels per image, vote rule == patch F1 of the oracle
    rs = np.random.RandomState(9)
    pm, tm = rs.rand(3, 608, 608, 1).astype(np.float32), (rs.rand(3, 608, 608, 1) > 0.4) * 1.0
    lp = images.patch_labels(pm, 16, rule=images.RULE_VOTE)
    lt = images.patch_labels(tm, 16, rule=images.RULE_VOTE)
    assert lp.shape == (3, 38, 38)
    assert abs(images.patch_scores(lp, lt)[3] - IO.patch_f1(pm, tm)) < 1e-12


def test_copy_windows_and_divide_by_hits():
    """rsu_copy_windows (job-table window copies with zero fill outside the image: the data movement
    of the shared-window prediction) and rsu_divide_by_hits (analytic count_hits division of
    images.py:154-162) against NumPy, bit-exact / 1e-7."""
    from road_segmentation_unet_b200 import images
    rs = np.random.RandomState(9)
    for C_, win, H in ((3, 40, 52), (1, 17, 30), (3, 36, 33)):
        x = rs.rand(3, H, H, C_).astype(np.float32)
        jobs = [(0, 0, 0, 2), (1, 12, 8, 0), (2, H - win + 5, -3, 1), (1, -4, H - 10, 3)]
        out = torch.full((4, win, win, C_), -7.0, dtype=torch.float32, device="cuda")
        images.copy_windows_dev(torch.tensor(x).cuda(), win, torch.tensor(jobs, dtype=torch.int32).cuda(), out)
        got = out.cpu().numpy()
        for img, y0, x0, dst in jobs:
            want = np.zeros((win, win, C_), np.float32)
            ys, xs = np.arange(y0, y0 + win), np.arange(x0, x0 + win)
            vy, vx = (ys >= 0) & (ys < H), (xs >= 0) & (xs < H)
            want[np.ix_(vy, vx)] = x[img][np.ix_(ys[vy], xs[vx])]
            assert np.array_equal(got[dst], want), (C_, win, H, dst)
    # overlap-add partial sums / hit counts == the reference's average
    side, P, stride = 4, 20, 8
    patches = rs.rand(2, side * side, P, P, 1)
    ref = IO.images_from_patches(patches, stride=stride)
    p = torch.tensor(patches.reshape(-1, P, P, 1), dtype=torch.float32).cuda()
    sums = images.images_from_patches_dev(p, 2, side, stride, normalize=False)
    avg = images.divide_by_hits_dev(sums, side, P, stride).cpu().numpy()
    assert np.abs(avg - ref).max() < 1e-6
    # d4 with repeated inputs == the ensemble of the reference (variant-major), one launch
    x = rs.rand(2, 12, 12, 3).astype(np.float32)
    ens = images.image_augmentation_ensemble_dev(torch.tensor(x).cuda()).cpu().numpy()
    assert np.array_equal(ens, IO.image_augmentation_ensemble(x).astype(np.float32))
    # any number of patches through the vote kernel (masks ride in grid.x)
    many = (rs.rand(70000, 4, 4) > 0.6).astype(np.float32)
    assert np.array_equal(images.labels_for_patches(many), IO.labels_for_patches(many))
