"""Generate tests/golden/submission_golden.npz from the reference's OWN committed outputs.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_submission_golden.py

/root/reference/submissions/<run>/ holds what the reference's `main` wrote for the 50 Kaggle test
images (tf_aerial_images.py:446-458): `images_NNN.png` = images.overlays(test image, quantised
mask, fade=0.4) saved by images.save_all, and `submission.csv` = images.save_submission_csv of the
same quantised masks.  Together with data/test/*.png (the inputs) they are known-answer vectors of
the reference for `load`, `img_float_to_uint8`, `overlays`, `save_all` and the row order / labels
of `save_submission_csv`: the overlay fixes where on the image every 16 x 16 cell of the mask lies,
the CSV must then list exactly those cells.

Stored (no reference source, data only):
  runs            names of the submission directories used (the five written by the final code;
                  the three older ones were blended with another fade)
  labels          [runs, 50, 38, 38] uint8: labels[r, k, x // 16, y // 16] of row "k+1_x_y"
  csv_sha256      sha-256 of every run's submission.csv
  csv_head        the first 80 lines of the first run's file
  crops           [n_crops, 3] (image index, top row, left column) of the stored 160 x 160 windows
  crop_rgb        [n_crops, 160, 160, 3] uint8 test-image pixels
  crop_overlay    [2, n_crops, 160, 160, 4] uint8 pixels of the overlay PNGs of the first two runs
"""
import csv
import glob
import hashlib
import os

import numpy as np
from PIL import Image

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "submission_golden.npz")
RUNS = [
    "2017-12-21T13h40m09s_multi_rotation_stochastic_aug_6_layers_epoch11_diluted_dropout1.0_93_994",
    "2017-12-17T10h55m14s_multi_rotation_stochastic_aug_6_layers_epoch_22_ensemble_dropout_1.0_94_124",
    "2017-12-17T01h53m02s_multi_rotation_stochastic_aug_6_layers_epoch_60_ensemble_93_642",
    "2017-12-16T11h01m42s_multi_rotation_stochastic_aug_6_layers_epoch_23_ensemble_93_623",
    "2017-12-15T21h16m47s_mutli_rotation_stochastic_aug_epoch_16_ensemble_93_504",
]
CROPS = [(0, 0, 448), (7, 208, 96), (49, 128, 256)]  # cell-aligned; the first touches the image edge
SIDE = 160
OVERLAY_RUNS = 2  # overlay pixels are kept for the first two runs, labels + CSV digests for all


def main():
    tests = sorted(glob.glob(os.path.join(REF, "data/test/*.png")))
    assert len(tests) == 50
    g = {"runs": np.array(RUNS), "crops": np.array(CROPS, np.int32)}
    g["crop_rgb"] = np.stack([np.array(Image.open(tests[k]).convert("RGB"))[r:r + SIDE, c:c + SIDE]
                              for k, r, c in CROPS])
    labels = np.zeros((len(RUNS), 50, 38, 38), np.uint8)
    shas, overlays = [], []
    for i, run in enumerate(RUNS):
        d = os.path.join(REF, "submissions", run)
        with open(os.path.join(d, "submission.csv"), "rb") as f:
            raw = f.read()
        shas.append(hashlib.sha256(raw).hexdigest())
        if i == 0:
            g["csv_head"] = np.frombuffer("".join(raw.decode().splitlines(True)[:80]).encode(), np.uint8)
        rows = list(csv.reader(raw.decode().splitlines()))
        assert rows[0] == ["id", "prediction"] and len(rows) == 1 + 50 * 38 * 38
        for rid, v in rows[1:]:
            k, x, y = rid.split("_")
            labels[i, int(k) - 1, int(x) // 16, int(y) // 16] = int(v)
        if i < OVERLAY_RUNS:
            overlays.append(np.stack([
                np.array(Image.open(os.path.join(d, "images_%03d.png" % (k + 1))))[r:r + SIDE, c:c + SIDE]
                for k, r, c in CROPS]))
    g["labels"] = labels
    g["csv_sha256"] = np.array(shas)
    g["crop_overlay"] = np.stack(overlays)
    assert g["crop_overlay"].shape == (OVERLAY_RUNS, len(CROPS), SIDE, SIDE, 4)
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
