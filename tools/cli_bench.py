"""Files in -> files out through the drop-in CLI (cli.main), timed on the wall clock for both loops:
    python tools/cli_bench.py [--n 8] [--size 256] [--batch 8] [--img_size 512] [--reps 3]
BASELINE config 1 = 8 frames of 256x256, batch 8, --img_size 512 (the reference's IMG_SIZE)."""
import argparse
import json
import shutil
import sys
import tempfile
import time
from pathlib import Path

import torch
from PIL import Image

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from unet_dc_segmentation_b200 import cli   # noqa: E402
from unet_dc_segmentation_b200.synth import calibrated_state_dict, synthetic_image   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--img_size", type=int, default=512)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    d = Path(tempfile.mkdtemp(prefix="cli_bench_"))
    (d / "in").mkdir()
    for i in range(a.n):
        Image.fromarray(synthetic_image(a.size, i)).save(d / "in" / f"frame{i:04d}.png")
    torch.save(calibrated_state_dict(seed=0), d / "ckpt.pth")
    common = ["--img_dir", str(d / "in"), "--ckpt_path", str(d / "ckpt.pth"), "--batch", str(a.batch), "--img_size",
              str(a.img_size), "--skip_excel", "--skip_histogram", "--px_per_micron", "3.45"]
    out = {"frames": a.n, "size": a.size, "batch": a.batch, "img_size": a.img_size}
    for tag, extra in (("fast", []), ("reference_loop", ["--reference_loop"])):
        times = []
        for r in range(a.reps + 1):                                  # first repetition = warm-up (library load, cuda init)
            od = d / f"out_{tag}_{r}"
            t0 = time.perf_counter()
            cli.main(common + extra + ["--out_dir", str(od)])
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        out[tag] = {"seconds_first": times[0], "seconds_best": min(times[1:]), "images_per_s": a.n / min(times[1:])}
    print(json.dumps(out))
    shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
