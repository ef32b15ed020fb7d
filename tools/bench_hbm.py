"""Achieved HBM bandwidth of the data-movement kernels (north_star item 3):
mirror pad, D4 transforms, patch extraction, overlap average, rotation + crop, ensemble inverse,
momentum SGD, max pool, head.  Algorithmic bytes = every input element read once + every output
element written once at the kernel's I/O dtype; the peak is MEASURED_PEAKS.json's copy bandwidth.
Working sets are larger than the 126 MB L2 (or the L2 is flushed between repetitions).

    python tools/bench_hbm.py [--json out.json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from road_segmentation_unet_b200 import images, ops  # noqa: E402


def peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


_FLUSH = None


def flush_l2():
    global _FLUSH
    if _FLUSH is None:
        _FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    _FLUSH.zero_()


def timed(fn, reps=5, flush=True):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        if flush:
            flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def run(n_img=48):  # BASELINE.json configs[2] with N = 8 images: 8 x 6 ensemble variants
    peak, src = peak_gbs()
    rows = []

    def add(name, nbytes, ms, note=""):
        gbs = nbytes / ms / 1e6
        rows.append({"kernel": name, "alg_bytes": int(nbytes), "ms": ms, "gbs": gbs, "frac": gbs / peak,
                     "note": note})

    g = torch.Generator(device="cuda").manual_seed(2017)
    # ---- prediction-side geometry at config-3 sizes: 604^2 RGB images, offset 188, stride 12
    x = torch.rand(n_img, 604, 604, 3, device="cuda", generator=g)
    ms = timed(lambda: images.mirror_border_dev(x, 188))
    add("mirror_pad", n_img * (604 ** 2 + 980 ** 2) * 3 * 4, ms, "%dx604^2x3 fp32, pad 188" % n_img)

    ops_t = torch.tensor([(0, 6, 4, 1, 2, 3)[i % 6] for i in range(n_img)], dtype=torch.uint8, device="cuda")
    ms = timed(lambda: images.d4_transform_dev(x, ops_t))
    add("d4_transform", 2 * x.numel() * 4, ms, "%dx604^2x3 fp32, ensemble ops" % n_img)

    lab = (torch.rand(64, 388, 388, device="cuda", generator=g) < 0.3).to(torch.uint8)
    ops_l = torch.tensor([i % 8 for i in range(64)], dtype=torch.uint8, device="cuda")
    big = torch.rand(32, 764, 764, 3, device="cuda", generator=g)
    ops_b = torch.tensor([i % 8 for i in range(32)], dtype=torch.uint8, device="cuda")
    ms = timed(lambda: images.d4_transform_dev(big, ops_b))
    add("d4_transform (train batch)", 2 * big.numel() * 4, ms, "32x764^2x3 fp32, all 8 ops")
    ms = timed(lambda: images.d4_transform_dev(lab, ops_l))
    add("d4_transform (labels u8)", 2 * lab.numel(), ms, "64x388^2 uint8")

    padded = images.mirror_border_dev(x[:1], 188)
    out = torch.empty(128, 764, 764, 3, dtype=torch.float32, device="cuda")
    ms = timed(lambda: images.extract_patches_dev(padded, 764, 12, 0, 128, out=out))
    # algorithmic bytes: the padded source image once (it stays in L2 across the overlapping
    # windows) + every patch element written once
    add("extract_patches", (padded.numel() + out.numel()) * 4, ms,
        "128 patches 764^2x3 fp32 out of one 980^2 image, stride 12: source read once + patches written once")

    # one predict() call of configs[2]: the 6 ensemble variants of a 604^2 image, 361 patches each
    preds = torch.rand(6 * 361, 388, 388, 1, device="cuda", generator=g)
    ms = timed(lambda: images.images_from_patches_dev(preds, 6, 19, 12))
    add("overlap_average", preds.numel() * 4 + 6 * 604 * 604 * 4, ms,
        "6 x 361 patches 388^2 -> 6 x 604^2, stride 12 (one predict call)")
    del preds

    # one angle of the README training-set preparation: 100 images 400^2, offset 188 (images.py:320-351)
    pad = images.mirror_border_dev(torch.rand(100, 400, 400, 3, device="cuda", generator=g), 216)
    ms = timed(lambda: images.rotate_crop_dev(pad, 30, 776))
    add("rotate_nn_crop", 100 * (832 ** 2 + 776 ** 2) * 3 * 4, ms, "100x832^2x3 -> 776^2, 30 deg")
    del pad

    probs = torch.rand(96, 608, 608, device="cuda", generator=g)
    ms = timed(lambda: images.patch_vote_dev(probs, 16, images.RULE_VOTE, 0.25, quantized=True, labels=True))
    add("patch_vote (quantize_mask)", 2 * probs.numel() * 4, ms, "96x608^2 fp32 -> quantised masks + 38x38 labels")

    masks = torch.rand(6 * 16, 604, 604, device="cuda", generator=g)  # two predict calls' worth
    ms = timed(lambda: images.invert_image_augmentation_ensemble_dev(masks))
    add("ensemble_invert", masks.numel() * 4 + 16 * 604 * 604 * 4, ms, "96x604^2 -> 16x604^2")

    # ---- training-side elementwise kernels at flagship sizes
    n = 155_776_078
    w = torch.rand(n, device="cuda", generator=g)
    acc = torch.zeros(n, device="cuda")
    gr = torch.rand(n, device="cuda", generator=g)
    ms = timed(lambda: ops.momentum_sgd(w, acc, gr, 0.01, 0.9))
    add("momentum_sgd", n * 20, ms, "155.8 M params: 12 B read + 8 B written each")

    a = torch.rand(16, 760, 760, 64, device="cuda", generator=g).to(torch.bfloat16)
    p = torch.empty(16, 380, 380, 64, dtype=torch.bfloat16, device="cuda")
    ms = timed(lambda: ops.maxpool2x2(a, p))
    add("maxpool2x2", (a.numel() + p.numel()) * 2, ms, "16x760^2x64 bf16")

    act = torch.rand(32, 388, 388, 64, device="cuda", generator=g).to(torch.bfloat16)
    wh = torch.rand(64, 2, device="cuda", generator=g)
    bh = torch.zeros(2, device="cuda")
    labels = (torch.rand(32, 388, 388, device="cuda", generator=g) < 0.3).to(torch.uint8)
    probs = torch.empty(32, 388, 388, device="cuda")
    loss = torch.zeros(1, device="cuda")
    dz = torch.empty_like(act)
    dw = torch.zeros(64, 2, device="cuda")
    db = torch.zeros(2, device="cuda")
    ms = timed(lambda: ops.head(act, wh, bh, labels=labels, probs=probs, loss=loss, dz=dz, dw=dw, db=db))
    add("head (1x1 conv + softmax CE + grads)", act.numel() * 4 + labels.numel() * 5, ms,
        "32x388^2x64 bf16 in, dZ out, labels, probs")
    ms = timed(lambda: ops.head(act, wh, bh, probs=probs))
    add("head (predict)", act.numel() * 2 + labels.numel() * 4, ms, "32x388^2x64 bf16 in, probs out")
    return {"peak_gbs": peak, "peak_source": src, "kernels": rows}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    res = run()
    print("%-40s %10s %9s %9s %6s  %s" % ("kernel", "MB", "ms", "GB/s", "frac", "workload"))
    for r in res["kernels"]:
        print("%-40s %10.1f %9.4f %9.1f %5.1f%%  %s" % (r["kernel"], r["alg_bytes"] / 1e6, r["ms"], r["gbs"],
                                                        100 * r["frac"], r["note"]))
    if a.json:
        with open(a.json, "w") as f:
            json.dump(res, f, indent=1)
