"""BASELINE config 5: UNetDC forward sweep over dilation sets and frame sizes 512-4096 px against the tensor roofline.

    python tools/sweep.py [--sizes 512,1024,2048,4096] [--out profiles/r01_sweep.md]

Forward only (dc_forward: stem + 17 conv3x3, the four transposed convs composed into dec{l}.0, threshold fused), u8 grayscale frames resident in HBM,
CUDA events, 3 warm-up + 5 timed calls per point.  Batch is chosen so the activation workspace stays under ~60 GB.
TFLOP/s uses the conservative in-bounds FLOP count; peak = MEASURED_PEAKS.json sustained bf16.
"""
import argparse
import json
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from unet_dc_segmentation_b200 import UNetDC, workload as wl   # noqa: E402
from unet_dc_segmentation_b200.synth import calibrated_state_dict   # noqa: E402

DILATION_SETS = {"reference (1,2,4,8,16)": (1, 2, 4, 8, 16), "plain UNet (1,1,1,1,1)": (1, 1, 1, 1, 1),
                 "(1,2,2,4,4)": (1, 2, 2, 4, 4), "(2,4,8,16,32)": (2, 4, 8, 16, 32)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="512,1024,2048,4096")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    peaks = json.loads((REPO / "MEASURED_PEAKS.json").read_text()) if (REPO / "MEASURED_PEAKS.json").exists() else {}
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    dev = torch.device("cuda", 0)
    sd = calibrated_state_dict(seed=0, calib_size=128, n_calib=1)
    rows = ["| dilations | frame | batch | ms / forward | TFLOP/s (in-bounds) | of sustained peak | frames/s |", "|---|---|---|---|---|---|---|"]
    for name, dil in DILATION_SETS.items():
        cls = type("UNetSweep", (UNetDC,), {"dilations": dil})
        m = cls(3, 1)
        m.load_state_dict(sd)
        m = m.to(dev).eval()
        for S in (int(v) for v in args.sizes.split(",")):
            B = max(1, min(32, int(60e9 // (0.93e9 * S * S / (1024 * 1024)))))
            x = torch.randint(0, 256, (B, S, S), dtype=torch.uint8, device=dev)
            mask = torch.empty((B, S, S), dtype=torch.uint8, device=dev)
            for _ in range(3):
                m.predict_u8(x, 0.3, mask_out=mask)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                m.predict_u8(x, 0.3, mask_out=mask)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            tf = B * wl.forward_flops(S, S, dil) / (ms * 1e9)
            rows.append(f"| {name} | {S}x{S} | {B} | {ms:.2f} | {tf:.0f} | {tf / peak:.2f} | {B / ms * 1e3:.1f} |")
            print(rows[-1], flush=True)
            del x, mask
        m._packed = None
        del m
        torch.cuda.empty_cache()
    if args.out:
        Path(args.out).write_text(__doc__.split("\n\n")[0] + "\n\n" + "\n".join(rows) + "\n")


if __name__ == "__main__":
    main()
