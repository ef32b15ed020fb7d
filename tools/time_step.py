"""Quick timing of the flagship train step (not the bench): per-phase CUDA-event times."""
import argparse, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from road_segmentation_unet_b200 import unet, ops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=6)
    ap.add_argument("--root", type=int, default=64)
    ap.add_argument("--dilated", type=int, default=1)
    ap.add_argument("--patch", type=int, default=388)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    S = unet.input_size_needed(a.patch, a.layers)
    t0 = time.time()
    net = unet.UNet(a.layers, a.root, bool(a.dilated), a.batch, S)
    torch.cuda.synchronize()
    print("built engine in %.1fs, mem %.1f GB" % (time.time() - t0, torch.cuda.memory_allocated() / 2**30))
    x = torch.rand(a.batch, S, S, 3, device="cuda")
    lab = (torch.rand(a.batch, a.patch, a.patch, device="cuda") < 0.3).to(torch.uint8)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    for it in range(a.steps):
        ops._lib.load().rsu_reset_launch_count()
        ev[0].record()
        net.zero_grads()
        net.forward(x, lab)
        ev[1].record()
        net.backward()
        ev[2].record()
        ops.momentum_sgd(net.params, net.momentum, net.grads, 0.01, 0.9)
        ev[3].record()
        net.pack_weights()
        ev[4].record()
        torch.cuda.synchronize()
        ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
        tot = sum(ts)
        fl = 3 * 680.82e9 * a.batch if (a.layers, a.root, a.dilated, a.patch) == (6, 64, 1, 388) else 0
        print("step %d: fwd %.1f ms  bwd %.1f ms  sgd %.2f ms  pack %.2f ms  total %.1f ms  -> %.1f patches/s  %.0f TFLOP/s  loss %.4f launches %d"
              % (it, ts[0], ts[1], ts[2], ts[3], tot, a.batch / tot * 1e3, fl / tot / 1e9, net.loss.item(),
                 ops.launch_count()))


if __name__ == "__main__":
    main()
