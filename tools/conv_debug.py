"""Structured probes of the tcgen05 conv kernel (run on the GPU box when a conv parity test fails).

Identity weights on the centre tap turn the layer into a copy, so any pixel / channel permutation caused by
a wrong TMA box order, swizzle or UMMA descriptor shows up directly as a mapping."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from unet_dc_segmentation_b200 import layers  # noqa: E402
from unet_dc_segmentation_b200.model import pack_conv3x3  # noqa: E402


def probe(cin, cout, H, W, tap=(1, 1), d=1):
    w = torch.zeros(cout, cin, 3, 3)
    for c in range(min(cin, cout)):
        w[c, c, tap[0], tap[1]] = 1.0
    b = torch.zeros(cout)
    pix = torch.arange(H * W, dtype=torch.float32).reshape(1, H, W, 1) % 128
    x_p = pix.expand(1, H, W, cin).contiguous().bfloat16()
    x_c = (torch.arange(cin, dtype=torch.float32) % 64).reshape(1, 1, 1, cin).expand(1, H, W, cin).contiguous().bfloat16()
    for name, x in (("pixel-id", x_p), ("channel-id", x_c)):
        got = layers.conv3x3(x.cuda(), pack_conv3x3(w).cuda(), b.cuda(), dilation=d, relu=False)
        torch.cuda.synchronize()
        got = got.float().cpu()
        xs = torch.zeros_like(x.float())
        dy, dx = (tap[0] - 1) * d, (tap[1] - 1) * d
        ys0, ys1 = max(0, -dy), min(H, H - dy)
        xs0, xs1 = max(0, -dx), min(W, W - dx)
        xs[:, ys0:ys1, xs0:xs1] = x.float()[:, ys0 + dy:ys1 + dy, xs0 + dx:xs1 + dx]
        want = xs[..., :cout] if cout <= cin else torch.cat([xs, torch.zeros(1, H, W, cout - cin)], -1)
        bad = (got != want)
        print(f"[{name}] cin={cin} cout={cout} {H}x{W} tap={tap} d={d}: mismatches {int(bad.sum())}/{bad.numel()}")
        if bad.any():
            print("  got [h=0, w=0..15, c=0]:", got[0, 0, :16, 0].tolist())
            print("  got [h=0..7, w=0, c=0]:", got[0, :8, 0, 0].tolist())
            print("  got [h=0, w=0, c=0..31]:", got[0, 0, 0, :32].tolist())
            print("  want[h=0, w=0, c=0..31]:", want[0, 0, 0, :32].tolist())
            print("  got [h=1, w=3, c=0..15]:", got[0, 1, 3, :16].tolist())


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    probe(64, 64, 8, 16)
    probe(64, 64, 8, 16, tap=(0, 0))
    probe(64, 64, 8, 16, tap=(2, 1))
    probe(128, 128, 8, 16)
    probe(128, 256, 16, 32)
    probe(64, 64, 24, 40, tap=(1, 2), d=2)
    print("conv_debug done")
