import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_layers import timed
from road_segmentation_unet_b200 import ops
for shp in ((32, 16, 16, 2048), (32, 18, 18, 2048), (32, 30, 30, 1024), (32, 28, 28, 1024), (32, 54, 54, 512), (32, 100, 100, 256), (32, 40, 40, 1024), (32, 88, 88, 512), (32, 392, 392, 64)):
    x = torch.randn(*shp, device="cuda").to(torch.bfloat16)
    out = torch.zeros(shp[3], device="cuda")
    ms = timed(lambda: ops.bias_grad(x, out), 20)
    ref = x.float().sum(dim=(0, 1, 2))
    out.zero_(); ops.bias_grad(x, out); torch.cuda.synchronize()
    err = float((out - ref).abs().max() / ref.abs().max())
    print("%-22s %7.1f MB  %6.1f us  (%.0f GB/s)  rel err %.1e" % (shp, x.numel() * 2 / 1e6, ms * 1e3, x.numel() * 2 / ms / 1e6, err))
