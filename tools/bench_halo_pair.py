"""A/B timing of the CTA-pair halo kernel (algo 4, conv_halo2.cu) against the single-CTA halo kernel
(algo 2) on the flagship network's layers that run the halo kernel, forward (bias + ReLU, fused pool
where the network has one) and data gradient (ReLU-gradient mask).   python tools/bench_halo_pair.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_layers import layer_list, timed  # noqa: E402
from road_segmentation_unet_b200 import ops  # noqa: E402

B, res = 32, {}
names = ("conv_0/conv2", "conv_dilut_0/atrous_conv2", "conv_1/conv1", "conv_1/conv2", "conv_dilut_1/atrous_conv1",
         "conv_dilut_1/atrous_conv2", "conv_9/conv1", "conv_9/conv2", "conv_10/conv1", "conv_10/conv2")
for name, srcs, cout, d, ho in layer_list():
    if name not in names:
        continue
    cin = sum(c for _, c, _ in srcs)
    xs = [torch.randn(B, e, e, c, device="cuda").to(torch.bfloat16) for e, c, _ in srcs]
    src_list = [(x, crop, crop) for x, (_, _, crop) in zip(xs, srcs)]
    w_fwd = (torch.randn(cout, 9 * cin, device="cuda") * 0.02).to(torch.bfloat16)
    bias = torch.zeros(cout, device="cuda")
    y = torch.empty(B, ho, ho, cout, device="cuda", dtype=torch.bfloat16)
    pool = torch.empty(B, ho // 2, ho // 2, cout, device="cuda", dtype=torch.bfloat16) \
        if name in ("conv_0/conv2", "conv_1/conv2") else None
    dz = torch.randn(B, ho, ho, cout, device="cuda").to(torch.bfloat16)
    flops = 2.0 * 9 * cin * cout * ho * ho * B
    row = {"gflop": flops / 1e9}
    for algo, tag in ((2, "halo"), (4, "pair")):
        try:
            row["fwd_" + tag] = timed(lambda: ops.conv3x3_fwd(src_list, w_fwd, bias, y, dilation=d, algo=algo,
                                                              pool_out=pool), 5)
        except Exception as e:
            row["fwd_" + tag] = float("nan")
    e0 = srcs[-1][0]
    dx = torch.empty(B, e0, e0, cin, device="cuda", dtype=torch.bfloat16)
    w_dg = (torch.randn(cin, 9 * cout, device="cuda") * 0.02).to(torch.bfloat16)
    mask = xs[-1] if len(srcs) == 1 else None
    for algo, tag in ((2, "halo"), (4, "pair")):
        try:
            row["dgrad_" + tag] = timed(lambda: ops.conv3x3_dgrad(dz, w_dg, dx, dilation=d, mask=mask, algo=algo), 5)
        except Exception as e:
            row["dgrad_" + tag] = float("nan")
    res[name] = row
    print("%-28s %8.1f GF | fwd halo %6.3f pair %6.3f x%.3f (%4.0f -> %4.0f TF/s) | dgrad halo %6.3f pair %6.3f x%.3f"
          % (name, flops / 1e9, row["fwd_halo"], row["fwd_pair"], row["fwd_halo"] / row["fwd_pair"],
             flops / row["fwd_halo"] / 1e9, flops / row["fwd_pair"] / 1e9, row["dgrad_halo"], row["dgrad_pair"],
             row["dgrad_halo"] / row["dgrad_pair"]), flush=True)
    del xs, src_list, y, dz, dx
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/halo_pair_ab.json", "w"), indent=1)
