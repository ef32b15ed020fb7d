"""A/B timing of the CTA-pair convolution kernel (algo 3, tcgen05.mma.cta_group::2) against the
library's current choice (algo 0: per-tap or halo kernel) on the flagship network's 3x3 layer
shapes, forward and data gradient.    python tools/bench_pair.py [--batch 32] [--out X.json]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_layers import layer_list, timed  # noqa: E402
from road_segmentation_unet_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default="gpurun_out/pair_ab.json")
    args = ap.parse_args()
    B, res = args.batch, {}
    for name, srcs, cout, d, ho in layer_list():
        cin = sum(c for _, c, _ in srcs)
        xs = [torch.randn(B, e, e, c, device="cuda").to(torch.bfloat16) for e, c, _ in srcs]
        src_list = [(x, crop, crop) for x, (_, _, crop) in zip(xs, srcs)]
        w_fwd = (torch.randn(cout, 9 * cin, device="cuda") * 0.02).to(torch.bfloat16)
        bias = torch.zeros(cout, device="cuda")
        y = torch.empty(B, ho, ho, cout, device="cuda", dtype=torch.bfloat16)
        dz = torch.randn(B, ho, ho, cout, device="cuda").to(torch.bfloat16)
        flops = 2.0 * 9 * cin * cout * ho * ho * B
        row = {"gflop": flops / 1e9}
        if cout % 128 == 0:
            for algo, tag in ((0, "auto"), (3, "pair")):
                row["fwd_" + tag] = timed(
                    lambda: ops.conv3x3_fwd(src_list, w_fwd, bias, y, dilation=d, algo=algo), args.reps)
        e0 = srcs[-1][0]
        if cin % 128 == 0:
            dx = torch.empty(B, e0, e0, cin, device="cuda", dtype=torch.bfloat16)
            w_dg = (torch.randn(cin, 9 * cout, device="cuda") * 0.02).to(torch.bfloat16)
            mask = xs[-1] if len(srcs) == 1 else None
            for algo, tag in ((0, "auto"), (3, "pair")):
                row["dgrad_" + tag] = timed(
                    lambda: ops.conv3x3_dgrad(dz, w_dg, dx, dilation=d, mask=mask, algo=algo), args.reps)
            del dx
        res[name] = row
        f = lambda k: "   -   " if k not in row else "%7.3f" % row[k]
        g = lambda a, b: "" if a not in row else " x%.3f" % (row[a] / row[b])
        print("%-28s %8.1f GF | fwd auto %s pair %s%s | dgrad auto %s pair %s%s"
              % (name, flops / 1e9, f("fwd_auto"), f("fwd_pair"), g("fwd_auto", "fwd_pair"),
                 f("dgrad_auto"), f("dgrad_pair"), g("dgrad_auto", "dgrad_pair")), flush=True)
        del xs, src_list, y, dz
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
