"""GPU diagnostic for the tcgen05 implicit-GEMM kernel: localises errors by row / column block."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from road_segmentation_unet_b200 import ops
from oracle import unet_oracle as O


def bf(x):
    return torch.tensor(x, dtype=torch.float32).bfloat16().float().numpy()


def dev(x, dt=torch.bfloat16):
    return torch.tensor(x, dtype=torch.float32).cuda().to(dt).contiguous()


def report(name, got, ref):
    err = np.abs(got - ref)
    rel = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)
    print("%-28s rel=%.3e max=%.3e  nan=%d" % (name, rel, err.max(), int(np.isnan(got).sum())))
    if rel > 1e-2:
        n, h, w, c = got.shape
        print("   err by y:", np.round(err.mean(axis=(0, 2, 3)), 3)[:24])
        print("   err by x:", np.round(err.mean(axis=(0, 1, 3)), 3)[:24])
        print("   err by c/8:", np.round(err.reshape(n, h, w, c // 8, 8).mean(axis=(0, 1, 2, 4)), 3)[:32])
        print("   got[0,0,0,:8]", got[0, 0, 0, :8], " ref", ref[0, 0, 0, :8])
    return rel


def main():
    torch.manual_seed(0)
    rs = np.random.RandomState(0)
    # 1x1 "conv": pure GEMM, one K step
    for (n, h, cin, cout) in [(1, 16, 64, 64), (1, 16, 128, 64), (1, 16, 64, 256), (2, 23, 192, 128)]:
        x = bf(rs.randn(n, h, h, cin).astype(np.float32))
        w = bf((rs.randn(1, 1, cin, cout) / np.sqrt(cin)).astype(np.float32))
        wp = torch.zeros(cout, cin, dtype=torch.bfloat16, device="cuda")
        ops.pack_conv_fwd(dev(w, torch.float32), wp, 1, cin, cout)
        assert np.array_equal(wp.float().cpu().numpy(), w.reshape(cin, cout).T), "pack_transpose"
        out = torch.zeros(n, h, h, cout, dtype=torch.bfloat16, device="cuda")
        ops.conv_gemm([(dev(x), 0, 0)], [(0, 0)], wp, out, cout)
        torch.cuda.synchronize()
        ref = O.conv2d_valid(torch.tensor(x), torch.tensor(w), None).numpy()
        report("1x1 n%d h%d %d->%d" % (n, h, cin, cout), out.float().cpu().numpy(), ref)
    # 3x3
    for (n, h, cin, cout, d) in [(1, 18, 64, 64, 1), (1, 21, 64, 64, 2), (2, 40, 128, 256, 1)]:
        x = bf(rs.randn(n, h, h, cin).astype(np.float32))
        w = bf((rs.randn(3, 3, cin, cout) / np.sqrt(9 * cin)).astype(np.float32))
        wp = torch.zeros(cout, 9 * cin, dtype=torch.bfloat16, device="cuda")
        ops.pack_conv_fwd(dev(w, torch.float32), wp, 9, cin, cout)
        ho = h - 2 * d
        out = torch.zeros(n, ho, ho, cout, dtype=torch.bfloat16, device="cuda")
        ops.conv3x3_fwd([(dev(x), 0, 0)], wp, None, out, dilation=d, relu=False)
        torch.cuda.synchronize()
        ref = O.conv2d_valid(torch.tensor(x), torch.tensor(w), None, d).numpy()
        report("3x3 n%d h%d %d->%d d%d" % (n, h, cin, cout, d), out.float().cpu().numpy(), ref)
    # wgrad 1 tap
    for (n, h, cin, cout) in [(1, 16, 64, 64), (1, 16, 128, 128), (2, 24, 64, 256)]:
        x = bf(rs.randn(n, h, h, cin).astype(np.float32))
        dz = bf(rs.randn(n, h, h, cout).astype(np.float32))
        out = torch.zeros(cin, cout, dtype=torch.float32, device="cuda")
        ops.wgrad_gemm([(dev(x), 0, 0)], [(0, 0)], dev(dz), (0, 0), out, (h, h))
        torch.cuda.synchronize()
        ref = np.einsum("nhwc,nhwo->co", x.astype(np.float64), dz.astype(np.float64))
        got = out.cpu().numpy()
        rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        print("wgrad 1tap n%d h%d %d->%d rel=%.3e" % (n, h, cin, cout, rel))
        if rel > 1e-2:
            err = np.abs(got - ref)
            print("   err by row/8:", np.round(err.reshape(cin // 8, 8, cout).mean(axis=(1, 2)), 3))
            print("   err by col/8:", np.round(err.reshape(cin, cout // 8, 8).mean(axis=(0, 2)), 3))
            print("   got[0,:8]", got[0, :8], "ref", ref[0, :8])


if __name__ == "__main__":
    main()
