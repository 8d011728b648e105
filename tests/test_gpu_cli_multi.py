"""GPU (needs >= 2 devices, skipped otherwise): the CLI under torchrun with two ranks -- frames sharded i -> rank
i mod 2, per-image files written by the owning rank, reports gathered on rank 0 -- gives the same reports as one rank."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent


def test_cli_two_ranks_equals_one_rank(tmp_path):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import pandas as pd
    from PIL import Image
    from unet_dc_segmentation_b200 import cli
    from unet_dc_segmentation_b200.synth import calibrated_state_dict, synthetic_image
    (tmp_path / "in").mkdir()
    for i in range(5):
        Image.fromarray(synthetic_image(96, 500 + i, n_droplets=10)).save(tmp_path / "in" / f"f{i}.png")
    torch.save(calibrated_state_dict(seed=0, calib_size=64, n_calib=2), tmp_path / "ckpt.pth")
    common = ["--img_dir", str(tmp_path / "in"), "--ckpt_path", str(tmp_path / "ckpt.pth"), "--batch", "2",
              "--px_per_micron", "3.45", "--skip_excel", "--skip_histogram", "--img_size", "96"]
    assert cli.main(common + ["--out_dir", str(tmp_path / "one")]) == 0
    env = dict(os.environ, PYTHONPATH=str(REPO), MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", "-m", "unet_dc_segmentation_b200.cli",
                        *common, "--out_dir", str(tmp_path / "two")], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for name in ("summary_per_image.csv", "all_droplets.csv", "droplet_size_stats.csv"):
        a = pd.read_csv(tmp_path / "one" / name, float_precision="round_trip")
        b = pd.read_csv(tmp_path / "two" / name, float_precision="round_trip")
        pd.testing.assert_frame_equal(a, b, check_exact=True)
    for i in range(5):
        assert (tmp_path / "two" / "predicted_masks" / f"f{i}_pred.png").read_bytes() == \
               (tmp_path / "one" / "predicted_masks" / f"f{i}_pred.png").read_bytes()
