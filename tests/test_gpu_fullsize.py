"""GPU: BASELINE configs[1] at FULL size (32 frames of 1024 x 1024 per batch) through properties that need no
oracle run of that size: run-to-run determinism, batch independence (a frame's result does not depend on its batch
neighbours), mask == (prob > thresh), table checksums against the mask, raster ordering of the labels, and the
oracle's bit-exact table on a sample of the frames given the kernel's mask."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def run(cuda_device):
    import torch
    from unet_dc_segmentation_b200 import DropletPipeline, UNetDC
    from unet_dc_segmentation_b200.synth import calibrated_state_dict, synthetic_image
    sd = calibrated_state_dict(seed=0)
    m = UNetDC(3, 1)
    m.load_state_dict(sd)
    m = m.to(cuda_device).eval()
    base = np.stack([synthetic_image(1024, i) for i in range(4)])
    frames = np.concatenate([np.roll(base, 37 * r, axis=2) for r in range(8)])            # 32 distinct frames
    pipe = DropletPipeline(m, 50, 0.3, 1, 3.45)
    dev_frames = torch.from_numpy(frames).to(cuda_device)
    res = pipe.run_device(dev_frames, return_prob=True, want_labels=True)
    torch.cuda.synchronize()
    out = {"frames": frames, "dev_frames": dev_frames, "pipe": pipe, "model": m,
           "probs": res.probs.clone(), "masks": res.masks.clone(), "labels": res.tables.labels.clone(),
           "tables": res.tables.to_host()}
    return out


def test_shapes_and_threshold(run):
    import torch
    assert tuple(run["probs"].shape) == (32, 1, 1024, 1024) and tuple(run["masks"].shape) == (32, 1024, 1024)
    assert torch.equal(run["masks"], (run["probs"][:, 0] > 0.3).to(torch.uint8))
    frac = float(run["masks"].float().mean())
    assert 0.02 < frac < 0.4, f"degenerate mask (foreground {frac:.3f})"


def test_deterministic_and_batch_independent(run):
    import torch
    pipe = run["pipe"]
    again = pipe.run_device(run["dev_frames"], return_prob=True)
    assert torch.equal(again.probs, run["probs"]) and torch.equal(again.masks, run["masks"])
    for i in (0, 13, 31):                                   # the same frame alone, and in a different batch position
        solo = pipe.run_device(run["dev_frames"][i:i + 1], return_prob=True)
        assert torch.equal(solo.probs[0], run["probs"][i]), f"frame {i} depends on its batch"
        assert torch.equal(solo.masks[0], run["masks"][i])
    pair = pipe.run_device(run["dev_frames"][[31, 0]], return_prob=True)
    assert torch.equal(pair.probs[0], run["probs"][31]) and torch.equal(pair.probs[1], run["probs"][0])


def test_tables_consistent_with_masks(run):
    masks = run["masks"].cpu().numpy()
    labels = run["labels"].cpu().numpy()
    for b, t in enumerate(run["tables"]):
        n = len(t["label"])
        assert n > 100
        assert int(t["area"].sum()) == int(masks[b].sum())                               # checksum of checksums
        assert labels[b].max() == n and np.array_equal(labels[b] > 0, masks[b] > 0)
        cnt = np.bincount(labels[b].ravel(), minlength=n + 1)[1:]
        np.testing.assert_array_equal(cnt, t["area"])
        first = np.full(n + 1, labels[b].size, np.int64)
        flat = labels[b].ravel()
        idx = np.flatnonzero(flat)
        np.minimum.at(first, flat[idx], idx)
        assert np.all(np.diff(first[1:]) > 0)                                              # raster order of first pixels
        np.testing.assert_array_equal(t["area_sqmicron"], t["area"] / (3.45 ** 2))


def test_sampled_frames_match_oracle_tables(run):
    masks = run["masks"].cpu().numpy()
    labels = run["labels"].cpu().numpy()
    for b in (0, 17):
        want_l, want = oracle.quantify_arrays(masks[b], 1, 3.45)
        np.testing.assert_array_equal(labels[b], want_l)
        for c, v in want.items():
            np.testing.assert_array_equal(np.asarray(run["tables"][b][c]), v, err_msg=f"frame {b} column {c}")


def test_rolling_ball_stage_matches_cv2_at_full_size(run):
    """The first stage at full size against OpenCV itself (the library the reference calls), 3 sampled frames."""
    import cv2
    import torch
    from unet_dc_segmentation_b200 import rolling_ball_device
    got = rolling_ball_device(run["dev_frames"], 50).cpu().numpy()
    k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (50, 50))
    for b in (0, 9, 31):
        f = run["frames"][b]
        want = cv2.normalize(cv2.subtract(f, cv2.morphologyEx(f, cv2.MORPH_OPEN, k)), None, 0, 255, cv2.NORM_MINMAX)
        np.testing.assert_array_equal(got[b], want, err_msg=f"frame {b}")
