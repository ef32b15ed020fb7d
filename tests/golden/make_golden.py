"""Generate tests/golden/images_golden.npz by running the reference's OWN src/images.py.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
The reference module is imported unmodified; matplotlib (absent here, used only by its load /
save helpers) is replaced by an empty stub module.  Inputs are seeded (seed 2017, the reference's
flag default) and small so the fixture stays a few hundred KB.
"""
import os
import sys
import types

import numpy as np

REF_SRC = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "images_golden.npz")


def import_reference_images():
    for name in ("matplotlib", "matplotlib.image"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].image = sys.modules["matplotlib.image"]
    sys.modules["matplotlib"].rcParams = {}
    sys.path.insert(0, REF_SRC)
    import images  # noqa: the reference's module
    return images


def main():
    import warnings
    warnings.simplefilter("ignore")
    ref = import_reference_images()
    rs = np.random.RandomState(2017)
    g = {}
    # mirror_border (images.py:269-281), 4-D and 3-D
    x4 = rs.rand(2, 10, 10, 3).astype(np.float32)
    x3 = rs.rand(2, 9, 9).astype(np.float32)
    g["mirror_in4"], g["mirror_out4_n4"] = x4, ref.mirror_border(x4, 4)
    g["mirror_in3"], g["mirror_out3_n7"] = x3, ref.mirror_border(x3, 7)
    # extract_patches (images.py:35-85)
    xe = rs.rand(2, 20, 20, 3).astype(np.float32)
    g["extract_in"] = xe
    g["extract_p8_s4"] = ref.extract_patches(xe, 8, stride=4)
    g["extract_p10_nostride"] = ref.extract_patches(xe, 10)
    g["extract3d_p12_s8"] = ref.extract_patches(xe[..., 0], 12, stride=8)
    # images_from_patches (images.py:131-164)
    pp = rs.rand(2, 16, 8, 8, 1).astype(np.float32)
    g["from_patches_in"] = pp
    g["from_patches_s4"] = ref.images_from_patches(pp, stride=4)
    g["from_patches_nostride"] = ref.images_from_patches(pp)
    pp64 = rs.rand(1, 9, 6, 6, 3)
    g["from_patches_in64"] = pp64
    g["from_patches64_s3"] = ref.images_from_patches(pp64, stride=3)
    # ensemble (images.py:376-417)
    xi = rs.rand(2, 7, 7, 3).astype(np.float32)
    g["ens_in"], g["ens_out"] = xi, ref.image_augmentation_ensemble(xi)
    mk = rs.rand(12, 7, 7, 1)
    g["inv_in"], g["inv_out"] = mk.copy(), ref.invert_image_augmentation_ensemble(mk.copy())
    # crop_imgs (images.py:354-373)
    xc = rs.rand(2, 13, 13, 2).astype(np.float32)
    g["crop_in"], g["crop_out8"] = xc, ref.crop_imgs(xc, 8)
    # rotate_imgs / expand_and_rotate (images.py:313-351)
    xr = rs.rand(2, 24, 24, 3).astype(np.float32)
    g["rot_in"] = xr
    for a in (15, 30, 45, 60, 75, 90):
        g["rot_%d" % a] = ref.rotate_imgs(xr, a)
    g["expand_in"] = xr
    g["expand_angles"] = np.array([0, 15, 45, 75])
    g["expand_off6"] = ref.expand_and_rotate(xr, [0, 15, 45, 75], 6)
    g["expand3d_off0"] = ref.expand_and_rotate(xr[..., 0], [30, 60], 0)
    # quantize_mask (images.py:256-266)
    qm = rs.rand(2, 32, 32, 1)
    g["quant_in"], g["quant_out"] = qm, ref.quantize_mask(qm, threshold=0.25, patch_size=16)
    # scoring rules and visual dumps (images.py:88-128, 167-180, 206-237, 284-310)
    lp = rs.rand(10, 16, 16) * 0.5
    g["labels_in"], g["labels_out"] = lp, ref.labels_for_patches(lp)
    g["pred_to_patches_in"] = np.arange(5)
    g["pred_to_patches_out"] = np.ascontiguousarray(ref.predictions_to_patches(np.arange(5), 4))
    oi = rs.rand(2, 12, 12, 3).astype(np.float32)
    om = rs.rand(2, 12, 12, 1)
    g["overlay_img"], g["overlay_mask"] = oi, om
    g["overlay_out_095"] = ref.overlays(oi, om)
    g["overlay_out_040"] = ref.overlays(oi, om, fade=0.4)
    pb, tb = (rs.rand(2, 12, 12) > 0.5) * 1, (rs.rand(2, 12, 12) > 0.5) * 1.0
    g["confusion_pred"], g["confusion_true"] = pb, tb
    g["confusion_out"] = ref.overlap_pred_true(pb, tb)
    g["error_out"] = ref.overlapp_error(pb, tb)
    import tempfile
    qd = ref.quantize_mask(rs.rand(2, 48, 48, 1), threshold=0.25, patch_size=16)
    with tempfile.TemporaryDirectory() as td:
        ref.save_submission_csv(qd, td, 16)
        with open(os.path.join(td, "submission.csv"), "rb") as f:
            g["csv_in"], g["csv_text"] = qd, np.frombuffer(f.read(), dtype=np.uint8)
    # the first rows of one of the reference's own submission files pin the on-disk format
    sub = sorted(os.listdir("/root/reference/submissions"))[0]
    with open(os.path.join("/root/reference/submissions", sub, "submission.csv"), "rb") as f:
        g["csv_reference_head"] = np.frombuffer(b"".join(f.readlines()[:40]), dtype=np.uint8)
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(g), "arrays")


if __name__ == "__main__":
    main()
