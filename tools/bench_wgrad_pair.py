"""A/B timing of the CTA-pair weight-gradient kernel (rsu_wgrad_desc.algo = 3) against the library's
choice (algo 0) on the flagship network's 3x3 layer shapes.   python tools/bench_wgrad_pair.py"""
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
    ap.add_argument("--out", default="gpurun_out/wgrad_pair_ab.json")
    args = ap.parse_args()
    B, res = args.batch, {}
    for name, srcs, cout, d, ho in layer_list():
        if cout % 128:
            continue
        cin = sum(c for _, c, _ in srcs)
        xs = [torch.randn(B, e, e, c, device="cuda").to(torch.bfloat16) for e, c, _ in srcs]
        src_list = [(x, crop, crop) for x, (_, _, crop) in zip(xs, srcs)]
        dz = torch.randn(B, ho, ho, cout, device="cuda").to(torch.bfloat16)
        dw = torch.zeros(9 * cin, cout, device="cuda")
        db = torch.zeros(cout, device="cuda")
        flops = 2.0 * 9 * cin * cout * ho * ho * B
        row = {"gflop": flops / 1e9}

        def auto():
            if not ops.conv3x3_wgrad(src_list, dz, dw, dilation=d, bias_grad=db, algo=0):
                ops.bias_grad(dz, db)

        def pair():
            if not ops.conv3x3_wgrad(src_list, dz, dw, dilation=d, bias_grad=db, algo=3):
                ops.bias_grad(dz, db)

        def tap():
            if not ops.conv3x3_wgrad(src_list, dz, dw, dilation=d, bias_grad=db, algo=1):
                ops.bias_grad(dz, db)

        row["auto"], row["tap"], row["pair"] = timed(auto, args.reps), timed(tap, args.reps), timed(pair, args.reps)
        res[name] = row
        print("%-28s %8.1f GF | auto %7.3f ms (%5.0f TF/s)  tap %7.3f (%5.0f)  pair+bias %7.3f (%5.0f)  x%.3f"
              % (name, flops / 1e9, row["auto"], flops / row["auto"] / 1e9, row["tap"], flops / row["tap"] / 1e9,
                 row["pair"], flops / row["pair"] / 1e9, row["auto"] / row["pair"]), flush=True)
        del xs, src_list, dz, dw
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
