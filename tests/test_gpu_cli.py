"""GPU: the drop-in CLI (unet_dc_segmentation_b200/cli.py <- reference quantify_droplets_batch.py:100-201):
same flags, same output files and columns; every per-image table equals the oracle's quantify() of the mask
PNG the CLI itself wrote (bit-exact through the CSV round trip)."""
import numpy as np
import pytest

import oracle
from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    import torch
    from PIL import Image
    from unet_dc_segmentation_b200.synth import calibrated_state_dict, synthetic_image
    d = tmp_path_factory.mktemp("cli")
    (d / "in").mkdir()
    for i in range(3):
        Image.fromarray(synthetic_image(96, 400 + i, n_droplets=10)).save(d / "in" / f"frame{i}.png")
    Image.fromarray(synthetic_image(128, 410, n_droplets=14)[:80, :112]).save(d / "in" / "wide.tif")
    (d / "in" / "notes.txt").write_text("not an image")
    torch.save(calibrated_state_dict(seed=0, calib_size=64, n_calib=2), d / "ckpt.pth")
    return d


def test_cli_native_size_outputs(cuda_device, workdir):
    import cv2
    import pandas as pd
    from unet_dc_segmentation_b200 import cli
    out = workdir / "out_native"
    rc = cli.main(["--img_dir", str(workdir / "in"), "--ckpt_path", str(workdir / "ckpt.pth"), "--out_dir", str(out),
                   "--batch", "2", "--px_per_micron", "3.45", "--save_overlays", "--skip_histogram", "--img_size", "96",
                   "--density_maps"])
    assert rc == 0
    names = ["frame0", "frame1", "frame2", "wide"]
    ref_cols = list(load_golden("reference_outputs.npz")["columns"])          # the reference's own all_droplets.csv header
    frames = []
    for n in names:
        mask = cv2.imread(str(out / "predicted_masks" / f"{n}_pred.png"), cv2.IMREAD_GRAYSCALE)
        assert mask is not None and set(np.unique(mask)) <= {0, 255}
        assert mask.shape == ((80, 112) if n == "wide" else (96, 96))
        # overlay = the reference's own three calls (qdb:75-77) on the frame and the mask that were written
        ov = cv2.imread(str(out / "overlays" / f"{n}_overlay.png"))
        src = [f for f in (workdir / "in").iterdir() if f.stem == n][0]
        want_ov = cv2.imread(str(src))
        cnts, _ = cv2.findContours(mask // 255, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        cv2.drawContours(want_ov, cnts, -1, (0, 255, 0), 2)
        np.testing.assert_array_equal(ov, want_ov, err_msg=f"{n} overlay")
        # density maps (quantify_pipline.py:131-141) = the restated reference functions on the same frame and mask
        from PIL import Image
        rgb = np.array(Image.open(src).convert("RGB"))
        roi, cy, cx = oracle.roi_mask(rgb)
        np.testing.assert_array_equal(np.load(out / f"{n}_radial_density.npy"), oracle.radial_density(mask // 255, roi, 10, cy, cx))
        np.testing.assert_array_equal(np.load(out / f"{n}_spatial_density.npy"), oracle.spatial_density(mask // 255, roi))
        df = pd.read_csv(out / f"{n}_droplets.csv", float_precision="round_trip")   # the default parser is 1 ulp sloppy
        want = oracle.quantify(mask // 255, 1, 3.45)
        if want.empty:
            assert len(df) == 0
            continue
        assert list(df.columns) == ref_cols
        for c in want.columns:
            np.testing.assert_array_equal(df[c].to_numpy(), want[c].to_numpy(), err_msg=f"{n}.{c}")
        frames.append(df)
    summary = pd.read_csv(out / "summary_per_image.csv")
    assert list(summary.columns) == ["filename", "droplet_count", "total_area_px"]
    assert list(summary["filename"]) == ["frame0.png", "frame1.png", "frame2.png", "wide.tif"]       # sorted, qdb:143
    combined = pd.read_csv(out / "all_droplets.csv", float_precision="round_trip")
    assert len(combined) == int(summary["droplet_count"].sum()) and combined["area"].sum() == summary["total_area_px"].sum()
    assert (out / "all_droplets.xlsx").exists() or (out / "all_droplets_noexcel.csv").exists()       # qdb:171-181
    stats = pd.read_csv(out / "droplet_size_stats.csv", index_col=0).iloc[:, 0]
    d = combined["eq_diam_micron"]
    np.testing.assert_allclose([stats["mean"], stats["median"], stats["std"]], [d.mean(), d.median(), d.std(ddof=1)])


def test_cli_default_img_size_resizes_like_the_reference(cuda_device, workdir):
    """IMG_SIZE = 512 (qdb:30): frames are up-sized for the network and the mask comes back at the original size."""
    import cv2
    from unet_dc_segmentation_b200 import cli
    out = workdir / "out_512"
    assert cli.main(["--img_dir", str(workdir / "in"), "--ckpt_path", str(workdir / "ckpt.pth"), "--out_dir", str(out),
                     "--batch", "8", "--skip_excel", "--skip_histogram"]) == 0
    mask = cv2.imread(str(out / "predicted_masks" / "wide_pred.png"), cv2.IMREAD_GRAYSCALE)
    assert mask.shape == (80, 112)
    assert not (out / "all_droplets.xlsx").exists() and not (out / "all_droplets_noexcel.csv").exists()
    assert not (out / "overlays").exists()


def test_cli_reference_loop_still_works(cuda_device, workdir):
    """--reference_loop keeps the reference-shaped per-image functions (preprocess / run_batch) on the CLI: same files,
    every table equal to the oracle's quantify() of the mask that loop wrote."""
    import cv2
    import pandas as pd
    from unet_dc_segmentation_b200 import cli
    out = workdir / "out_loop"
    assert cli.main(["--img_dir", str(workdir / "in"), "--ckpt_path", str(workdir / "ckpt.pth"), "--out_dir", str(out),
                     "--batch", "3", "--skip_excel", "--skip_histogram", "--img_size", "96", "--reference_loop",
                     "--save_overlays"]) == 0
    fast = workdir / "out_native"
    for n in ["frame0", "frame1", "frame2", "wide"]:
        mask = cv2.imread(str(out / "predicted_masks" / f"{n}_pred.png"), cv2.IMREAD_GRAYSCALE)
        df = pd.read_csv(out / f"{n}_droplets.csv", float_precision="round_trip")
        want = oracle.quantify(mask // 255, 1, None)
        assert len(df) == len(want)
        for c in want.columns:
            np.testing.assert_array_equal(df[c].to_numpy(), want[c].to_numpy(), err_msg=f"{n}.{c}")
        assert (out / "overlays" / f"{n}_overlay.png").exists()
        if (fast / "predicted_masks" / f"{n}_pred.png").exists():
            # the two loops feed the stem differently (f32 RGB vs folded u8 grey): a few threshold-edge pixels may differ
            other = cv2.imread(str(fast / "predicted_masks" / f"{n}_pred.png"), cv2.IMREAD_GRAYSCALE)
            assert (other != mask).mean() < 0.01


def test_cli_true_colour_frames(cuda_device, tmp_path):
    """A frame whose channels differ takes the [B,H,W,3] path (rolling ball per channel, qdb:41-43; u8 HWC stem), next
    to a grayscale frame of the same size: both end as masks + tables consistent with the oracle, and the colour
    frame's network input equals the reference's per-channel rolling ball (checked through the mask the f32 module
    path gives for the same preprocessed frame)."""
    import cv2
    import pandas as pd
    import torch
    from PIL import Image
    from unet_dc_segmentation_b200 import cli
    from unet_dc_segmentation_b200.synth import calibrated_state_dict, synthetic_image
    (tmp_path / "in").mkdir()
    g = synthetic_image(96, 420, n_droplets=10)
    rgb = np.stack([g, (g.astype(np.int32) * 3 // 4).astype(np.uint8), np.roll(g, 5, axis=1)], axis=-1)
    Image.fromarray(rgb).save(tmp_path / "in" / "colour.png")
    Image.fromarray(g).save(tmp_path / "in" / "grey.png")
    sd = calibrated_state_dict(seed=0, calib_size=64, n_calib=2)
    torch.save(sd, tmp_path / "ckpt.pth")
    out = tmp_path / "out"
    assert cli.main(["--img_dir", str(tmp_path / "in"), "--ckpt_path", str(tmp_path / "ckpt.pth"), "--out_dir", str(out),
                     "--batch", "4", "--skip_excel", "--skip_histogram", "--img_size", "96", "--save_overlays"]) == 0
    for n in ("colour", "grey"):
        mask = cv2.imread(str(out / "predicted_masks" / f"{n}_pred.png"), cv2.IMREAD_GRAYSCALE)
        assert mask.shape == (96, 96)
        df = pd.read_csv(out / f"{n}_droplets.csv", float_precision="round_trip")
        want = oracle.quantify(mask // 255, 1, None)
        assert len(df) == len(want)
        for c in want.columns:
            np.testing.assert_array_equal(df[c].to_numpy(), want[c].to_numpy(), err_msg=f"{n}.{c}")
        assert (out / "overlays" / f"{n}_overlay.png").exists()
    # the colour frame against the reference-shaped per-image functions (same kernels, f32 NCHW entry): masks agree up
    # to threshold-edge pixels
    m = cli.load_model(tmp_path / "ckpt.pth")
    t, _ = cli.preprocess(tmp_path / "in" / "colour.png", 50, 96)
    want_pre = oracle.rolling_ball_correction_rgb(rgb, 50)
    np.testing.assert_array_equal((t.permute(1, 2, 0) * 255.0).round().to(torch.uint8).cpu().numpy(), want_pre)
    probs = m(t[None])
    ref_mask = (probs[0, 0] > 0.3).to(torch.uint8).cpu().numpy()
    got = cv2.imread(str(out / "predicted_masks" / "colour_pred.png"), cv2.IMREAD_GRAYSCALE) // 255
    assert (got != ref_mask).mean() < 0.01
    summary = pd.read_csv(out / "summary_per_image.csv")
    assert list(summary["filename"]) == ["colour.png", "grey.png"]
