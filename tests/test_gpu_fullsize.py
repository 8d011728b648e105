"""GPU: BASELINE configs[1] at FULL size (32 frames of 1024 x 1024 per batch) through properties that need no
oracle run of that size: run-to-run determinism, batch independence (a frame's result does not depend on its batch
neighbours), mask == (prob > thresh), table checksums against the mask, raster ordering of the labels, and the
oracle's bit-exact table on a sample of the frames given the kernel's mask."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _fused(sd):
    """Composed weights of the decoder levels the CUDA forward runs as one launch each by default (UNetDC.fuse_level1,
    UNetDC.fuse_levels), for the emulation."""
    from unet_dc_segmentation_b200.model import fused_level_blobs
    cpu = {k: v.detach().float().cpu() for k, v in sd.items() if k.startswith(("dec", "upconv"))}
    return {lvl: fused_level_blobs(cpu, lvl) for lvl in (1, 2, 3, 4)}


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


def test_forward_parity_at_full_size(run):
    """The benchmarked shape itself (BASELINE configs[1]: 1024 x 1024 frames, batch 32): frames 0 and 17 of the batch
    against the fp32 reference network (oracle.unetdc_forward restates models/model_2.py:56-80) and against the bf16
    emulation.  At this size the bottleneck map is 64 x 64, so dilation 16 (model_2.py:16) and dilation 8
    (model_2.py:13) run with all nine taps in bounds, and every layer runs its full multi-tile schedule.
    Tolerances: tests/test_gpu_forward.py (PROB_TOL / EMU_TOL); masks may differ from the reference's only inside
    the PROB_TOL band around prob_thresh, and the droplet table is bit-exact given the kernel's mask."""
    import torch
    from test_gpu_forward import EMU_MEAN_TOL, EMU_TOL, MEAN_TOL, PROB_TOL
    from unet_dc_segmentation_b200 import rolling_ball_device
    from unet_dc_segmentation_b200.synth import calibrated_state_dict
    sd = calibrated_state_dict(seed=0)
    idx = [0, 17]
    pre = rolling_ball_device(run["dev_frames"][idx], 50).cpu().numpy()          # bit-exact vs cv2 (test above)
    x = torch.from_numpy(np.repeat(pre[:, None], 3, 1).astype(np.float32) / 255.0)
    ref = oracle.unetdc_forward(sd, x).numpy()[:, 0]
    emu = oracle.unetdc_forward(sd, x, emulate_bf16=True, fused_levels=_fused(sd), gray_input=True).numpy()[:, 0]
    got = run["probs"][idx, 0].cpu().numpy()
    masks = run["masks"][idx].cpu().numpy()
    e_ref, e_emu = np.abs(got - ref), np.abs(got - emu)
    ref_mask = (ref > 0.3).astype(np.uint8)
    diff = masks != ref_mask
    in_band = np.abs(ref - 0.3) <= PROB_TOL
    n_ref = [len(oracle.quantify_arrays(m, 1, None)[1]["label"]) for m in ref_mask]
    n_got = [len(run["tables"][i]["label"]) for i in idx]
    print(f"1024^2: vs fp32 max {e_ref.max():.4f} mean {e_ref.mean():.5f}; vs bf16 emulation max {e_emu.max():.4f} "
          f"mean {e_emu.mean():.5f}; mask pixels differing {int(diff.sum())} of {diff.size} "
          f"({int((diff & ~in_band).sum())} outside the band); droplets {n_got} vs reference {n_ref}")
    assert e_ref.max() <= PROB_TOL and e_ref.mean() <= MEAN_TOL
    assert e_emu.max() <= EMU_TOL and e_emu.mean() <= EMU_MEAN_TOL
    assert int((diff & ~in_band).sum()) == 0
    assert diff.mean() < 0.005
    for a, b in zip(n_got, n_ref):
        assert abs(a - b) <= 0.03 * b, f"droplet count {a} vs reference {b}"
