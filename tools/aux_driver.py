"""Small driver for profiling the non-tensor stages alone (rolling ball, labelling + table, overlay stencil) at the
bench shape (32 x 1024^2) without the network:  python tools/aux_driver.py [--reps 3] [--size 1024] [--batch 32]"""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from unet_dc_segmentation_b200 import label_stats_device, overlay_stencil_device, rolling_ball_device   # noqa: E402
from unet_dc_segmentation_b200.synth import synthetic_image, synthetic_mask   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--overlay", action="store_true")
    a = ap.parse_args()
    S, B = a.size, a.batch
    dev = torch.device("cuda", 0)
    base_f = [synthetic_image(S, i) for i in range(4)]
    n_discs = int(3500 * (S / 1024.0) ** 2)
    base_m = [synthetic_mask(S, n_discs, seed=i, rmin=2, rmax=5) for i in range(4)]
    frames = torch.from_numpy(np.stack([np.roll(base_f[i % 4], 31 * (i // 4), axis=1) for i in range(B)])).to(dev)
    masks = torch.from_numpy(np.stack([np.roll(base_m[i % 4], 17 * (i // 4), axis=0) for i in range(B)])).to(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for rep in range(a.reps):
        ev[0].record()
        rolling_ball_device(frames, 50)
        ev[1].record()
        t = label_stats_device(masks, 1, 3.45)
        ev[2].record()
        if a.overlay:
            overlay_stencil_device(masks)
        ev[3].record()
        torch.cuda.synchronize()
        print(f"rep {rep}: rolling ball {ev[0].elapsed_time(ev[1]):.3f} ms, label+table {ev[1].elapsed_time(ev[2]):.3f} ms, "
              f"overlay {ev[2].elapsed_time(ev[3]):.3f} ms, droplets/frame {float(t.counts.float().mean()):.0f}")


if __name__ == "__main__":
    main()
