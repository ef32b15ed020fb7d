"""Timing of the transpose-convolution kernels (forward, data gradient, weight gradient) at the
flagship network's up_conv shapes, with the HBM floor of each.   python tools/bench_upconv.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_layers import timed  # noqa: E402
from road_segmentation_unet_b200 import ops  # noqa: E402

B = 32
for name, h, cin, cout in (("up_conv_4", 196, 128, 64), ("up_conv_3", 100, 256, 128), ("up_conv_2", 52, 512, 256),
                           ("up_conv_1", 28, 1024, 512), ("up_conv_0", 16, 2048, 1024)):
    x = torch.randn(B, h, h, cin, device="cuda").to(torch.bfloat16)
    w_fwd = (torch.randn(4 * cout, cin, device="cuda") * 0.05).to(torch.bfloat16)
    w_dg = (torch.randn(cin, 4 * cout, device="cuda") * 0.05).to(torch.bfloat16)
    bias = torch.zeros(cout, device="cuda")
    y = torch.empty(B, 2 * h, 2 * h, cout, device="cuda", dtype=torch.bfloat16)
    dy = torch.randn(B, 2 * h, 2 * h, cout, device="cuda").to(torch.bfloat16)
    dx = torch.empty_like(x)
    dw = torch.zeros(4 * cout, cin, device="cuda")
    t_f = timed(lambda: ops.upconv2x2_fwd(x, w_fwd, bias, y), 5)
    t_d = timed(lambda: ops.upconv2x2_dgrad(dy, w_dg, dx, mask=x), 5)
    t_w = timed(lambda: ops.upconv2x2_wgrad(dy, x, dw), 5)
    gb_f = (x.numel() + y.numel()) * 2 / 1e9
    gb_d = (dy.numel() + 2 * x.numel()) * 2 / 1e9
    gb_w = (dy.numel() + x.numel()) * 2 / 1e9
    fl = 2.0 * 4 * cin * cout * h * h * B
    print("%-10s fwd %.3f ms (%.2f GB -> floor %.3f, %4.0f TF/s) | dgrad %.3f ms (floor %.3f, %4.0f TF/s) | "
          "wgrad %.3f ms (floor %.3f, %4.0f TF/s)" % (name, t_f, gb_f, gb_f / 6.55, fl / t_f / 1e9, t_d, gb_d / 6.55,
                                                      fl / t_d / 1e9, t_w, gb_w / 6.55, fl / t_w / 1e9), flush=True)
