"""Per-layer A/B timing of the two tensor-core convolution algorithms (one TMA box per tap vs. the
halo tile) on the flagship network's layer shapes: forward, data gradient and weight gradient.

    python tools/bench_layers.py [--batch 32] [--out gpurun_out/layers_ab.json] [--only SUBSTR]

Inputs are random bf16 tensors of the real shapes (far larger than L2 for the big layers); every
variant is run 1 + `reps` times and the mean of the timed runs (CUDA events) is reported.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from road_segmentation_unet_b200 import ops, unet  # noqa: E402


def layer_list(L=6, root=64, P=388):
    """[(name, [(src_extent, channels, crop)], cout, dilation, out_extent)] of the dilated U-Net."""
    S = unet.input_size_needed(P, L)
    f = [root * 2 ** i for i in range(L)]
    in_size, s = [], S
    for i in range(L):
        in_size.append(s)
        s = (s - 4) // 2
    skip = [v - 4 for v in in_size]
    up, net = [], skip[L - 1]
    for j in range(L - 1):
        up.append(2 * net)
        net = 2 * net - 4
    out = []
    for i in range(L):
        cin = 64 if i == 0 else f[i - 1]  # level 0 conv1 runs on the 64-channel im2col (1 tap)
        if i > 0:
            out.append(("conv_%d/conv1" % i, [(in_size[i], cin, 0)], f[i], 1, in_size[i] - 2))
        out.append(("conv_%d/conv2" % i, [(in_size[i] - 2, f[i], 0)], f[i], 1, in_size[i] - 4))
        if i < L - 1:
            t = up[L - 2 - i]
            if i > 0:
                # reads the window [o2, o2 + t + 8) of the level input
                out.append(("conv_dilut_%d/atrous_conv1" % i, [(t + 8, cin, 0)], f[i], 2, t + 4))
            out.append(("conv_dilut_%d/atrous_conv2" % i, [(t + 4, f[i], 0)], f[i], 2, t))
    for j in range(L - 1):
        i = L - 2 - j
        fo, t = f[i], up[j]
        srcs = [(skip[i], fo, (skip[i] - t) // 2), (t, fo, 0), (t, fo, 0)]
        out.append(("conv_%d/conv1" % (L + j), srcs, fo, 1, t - 2))
        out.append(("conv_%d/conv2" % (L + j), [(t - 2, fo, 0)], fo, 1, t - 4))
    return out


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default="gpurun_out/layers_ab.json")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    B = args.batch
    res = {}
    for name, srcs, cout, d, ho in layer_list():
        if args.only and not any(o == name or (o.endswith("*") and name.startswith(o[:-1]))
                                 for o in args.only.split(",")):
            continue
        cin = sum(c for _, c, _ in srcs)
        xs = [torch.randn(B, e, e, c, device="cuda").to(torch.bfloat16) for e, c, _ in srcs]
        src_list = [(x, crop, crop) for x, (_, _, crop) in zip(xs, srcs)]
        w_fwd = (torch.randn(cout, 9 * cin, device="cuda") * 0.02).to(torch.bfloat16)
        bias = torch.zeros(cout, device="cuda")
        y = torch.empty(B, ho, ho, cout, device="cuda", dtype=torch.bfloat16)
        dz = torch.randn(B, ho, ho, cout, device="cuda").to(torch.bfloat16)
        flops = 2.0 * 9 * cin * cout * ho * ho * B
        row = {"gflop": flops / 1e9}
        for algo, tag in ((1, "tap"), (2, "halo")):
            try:
                ms = timed(lambda: ops.conv3x3_fwd(src_list, w_fwd, bias, y, dilation=d, algo=algo), args.reps)
                row["fwd_" + tag] = ms
            except Exception as e:  # not eligible
                row["fwd_" + tag] = None
                row["fwd_%s_err" % tag] = str(e)[:80]
        # data gradient w.r.t. a single-source input of cin channels (concat layers: all sources
        # at once, as the engine does, into one tensor of the up-sampled extent)
        e0 = srcs[-1][0]
        dx = torch.empty(B, e0, e0, cin, device="cuda", dtype=torch.bfloat16)
        w_dg = (torch.randn(cin, 9 * cout, device="cuda") * 0.02).to(torch.bfloat16)
        mask = xs[-1] if len(srcs) == 1 else None
        for algo, tag in ((1, "tap"), (2, "halo")):
            try:
                ms = timed(lambda: ops.conv3x3_dgrad(dz, w_dg, dx, dilation=d, mask=mask, algo=algo), args.reps)
                row["dgrad_" + tag] = ms
            except Exception as e:
                row["dgrad_" + tag] = None
                row["dgrad_%s_err" % tag] = str(e)[:80]
        dw = torch.zeros(9 * cin, cout, device="cuda")
        db = torch.zeros(cout, device="cuda")
        for algo, tag in ((1, "tap"), (2, "halo")):
            try:
                ms = timed(lambda: ops.conv3x3_wgrad(src_list, dz, dw, dilation=d, bias_grad=db, algo=algo),
                           args.reps)
                row["wgrad_" + tag] = ms
            except Exception as e:
                row["wgrad_" + tag] = None
                row["wgrad_%s_err" % tag] = str(e)[:80]
        row["bias_grad"] = timed(lambda: ops.bias_grad(dz, db), args.reps)
        res[name] = row
        fmt = lambda v: "   n/a " if v is None else "%7.3f" % v
        tf = lambda v: "  n/a" if v is None else "%5.0f" % (flops / v / 1e9)
        print("%-28s %8.1f GF | fwd %s %s ms (%s %s TF/s) | dgrad %s %s (%s %s) | wgrad %s %s (%s %s) | bias %.3f"
              % (name, flops / 1e9, fmt(row["fwd_tap"]), fmt(row["fwd_halo"]), tf(row["fwd_tap"]),
                 tf(row["fwd_halo"]), fmt(row["dgrad_tap"]), fmt(row["dgrad_halo"]), tf(row["dgrad_tap"]),
                 tf(row["dgrad_halo"]), fmt(row["wgrad_tap"]), fmt(row["wgrad_halo"]), tf(row["wgrad_tap"]),
                 tf(row["wgrad_halo"]), row["bias_grad"]), flush=True)
        del xs, src_list, y, dz, dx, dw
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
