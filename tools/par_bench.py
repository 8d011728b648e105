"""enc1.3 / dec1.3 at the benchmark shape (32 x 1024 x 1024 x 64): the 4-strip kernel (conv_halo2_kernel<64, 4>) against the
parity-class kernel with shared windows (conv_par2_kernel), CUDA events, median."""
import statistics
import sys

import torch

sys.path.insert(0, ".")
from unet_dc_segmentation_b200 import layers                                    # noqa: E402
from unet_dc_segmentation_b200.model import pack_conv3x3, pack_par3x3          # noqa: E402

B, H, W = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (32, 1024, 1024)
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
w = (torch.randn(64, 64, 3, 3, generator=g) / 24.0).bfloat16().float()
b = (torch.randn(64, generator=g) * 0.1).to(dev)
hw = (torch.randn(64, generator=g) * 0.2).to(dev)
pw, pp = pack_conv3x3(w).to(dev), pack_par3x3(w).to(dev)
x = torch.randn((B, H, W, 64), device=dev).bfloat16()
cat = torch.empty((B, H, W, 128), dtype=torch.bfloat16, device=dev)


def timed(fn, n=7):
    ts = []
    for _ in range(n):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); e.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    return statistics.median(ts[2:]), r


for name, fn in (("enc1.3 (store + pool)", lambda wp: layers.conv3x3(x, pw, b, out=cat, out_offset=64, pool=True, weight_par=wp)),
                 ("dec1.3 (head)", lambda wp: layers.conv3x3_head(x, pw, b, hw, 0.1, 0.3, weight_par=wp))):
    t0, r0 = timed(lambda: fn(None))
    t1, r1 = timed(lambda: fn(pp))
    d = max(float((a.float() - c.float()).abs().max()) for a, c in zip(r0, r1))
    print(f"{name}: 4-strip {t0:.3f} ms, parity classes {t1:.3f} ms, max |diff| {d:.4f}")
