"""Train the flagship model on the reference's own data set (data/training: 100 aerial images of
400^2 with road masks) through the reference-shaped API and score held-out images.

The PNG files are not part of this repository: copy /root/reference/data/training to
data_cache/training (git-ignored; it travels to the GPU box with the snapshot) and run

    python tools/train_real.py [--epochs 8] [--batch 4] [--angles 0,30,60] [--holdout 10]

Flow = tf_aerial_images.main (tf_aerial_images.py:403-430) with the README's flags (6 layers,
dilated, 388^2 patches, stride 12, image augmentation, dropout 1.0, ensemble prediction), the
patch list built angle by angle in float32 to bound host memory.  Prints the mean loss per epoch
and, for the held-out images, pixel accuracy and the patch-level scores of the submission rule
(16 x 16 cells, images.py:256-266, summary.py:141-147).
"""
import argparse
import contextlib
import io
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from road_segmentation_unet_b200 import images, unet, tf_aerial_images as tfa  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", default=os.path.join(ROOT, "data_cache", "training"))
    ap.add_argument("--epochs", type=int, default=8)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--angles", default="0,30,60")
    ap.add_argument("--holdout", type=int, default=10)
    ap.add_argument("--lr", type=float, default=0.01)
    args = ap.parse_args()

    flags = tfa.make_parser().parse_args([
        "--num_layers=6", "--dilated_layers", "--patch_size=388", "--stride=12", "--dropout=1.0",
        "--image_augmentation", "--ensemble_prediction", "--batch_size=%d" % args.batch, "--lr=%g" % args.lr,
        "--rotation_angles=" + args.angles, "--gpu=0", "--eval_every=1000000000",
        "--train_score_every=1000000000", "--save_path=/tmp/rsu_real_runs"])
    opts = tfa.Options(flags)
    imgs, gts = images.load_train_data(args.data)
    n_train = imgs.shape[0] - args.holdout
    tr_i, tr_g, ho_i, ho_g = imgs[:n_train], gts[:n_train], imgs[n_train:], gts[n_train:]
    print("train %d images, hold out %d; road fraction %.3f" % (n_train, args.holdout, float((gts >= 0.5).mean())))

    model = tfa.ConvolutionalModel(opts, None)
    S = model.input_size
    offset = (S - opts.patch_size) // 2
    t0 = time.time()
    patches, labels = [], []
    for angle in opts.rotation_angles:
        ext = images.expand_and_rotate(tr_i, [angle], offset)
        patches.append(images.extract_patches(ext, S, predict_patch_size=opts.patch_size,
                                              stride=opts.stride).astype(np.float32))
        lab = images.expand_and_rotate(tr_g, [angle], 0)
        labels.append(images.extract_patches(lab, opts.patch_size, stride=opts.stride).astype(np.float32))
    patches, labels = np.concatenate(patches), np.concatenate(labels)
    print("prepared %d patches of %d^2 (+ %d^2 masks) in %.1f s" % (patches.shape[0], S, opts.patch_size, time.time() - t0))

    for epoch in range(args.epochs):
        t0 = time.time()
        first = len(model.scalars)
        with contextlib.redirect_stdout(io.StringIO()):
            model.train(patches, labels, tr_i, tr_g)
        losses = [s[1] for s in model.scalars[first:]]
        dt = time.time() - t0
        print("epoch %d: %d steps, mean loss %.4f (last %.4f), %.1f s, %.0f patches/s incl. host batching"
              % (epoch, len(losses), float(np.mean(losses)), losses[-1], dt, len(losses) * args.batch / dt))

    t0 = time.time()
    with contextlib.redirect_stdout(io.StringIO()):
        masks = model.predict_batchwise(ho_i, 5)
    print("predicted %d held-out images (6-way ensemble) in %.2f s" % (ho_i.shape[0], time.time() - t0))
    truth = (ho_g >= 0.5).astype(np.float32)[..., None]
    pix_acc = float(np.mean((masks >= 0.5) == (truth >= 0.5)))
    lp = images.patch_labels(masks, 16, rule=images.RULE_VOTE)
    lt = images.patch_labels(truth, 16, rule=images.RULE_VOTE)
    acc, rec, prec, f1 = images.patch_scores(lp, lt)
    print("held-out: pixel accuracy %.4f; 16x16 patches: accuracy %.4f recall %.4f precision %.4f F1 %.4f"
          % (pix_acc, acc, rec, prec, f1))


if __name__ == "__main__":
    main()
