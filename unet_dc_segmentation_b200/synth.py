"""Synthetic inputs and a synthetic checkpoint for the droplet-quantification path.

The public UNetDC checkpoint is a Google-Drive link (reference README.md:54) and there is no
network, so measurement and parity use:

* grayscale microscopy-like images with planted Gaussian droplets (``synthetic_image``),
* binary disc masks for the labelling stage (``synthetic_mask``),
* a seeded default-init ``UNetDC`` whose BatchNorm running statistics are calibrated with one
  training-mode pass (``calibrated_state_dict``) -- a default-init network in eval mode outputs
  ~0.476 everywhere, which makes every mask all-ones and parity vacuous.

Nothing here produces predictions: ``calibrated_state_dict`` only manufactures a ``state_dict``
(the thing ``torch.load(ckpt)`` returns at reference quantify_droplets_batch.py:36), which is
then loaded through ``load_state_dict`` by whichever implementation is under test.
"""
from __future__ import annotations

import math

import numpy as np


def synthetic_image(size: int, index: int = 0, n_droplets: int | None = None, with_truth: bool = False):
    """u8 [size, size] grayscale frame: background 20 + N(0,3), planted Gaussian blobs.
    with_truth: also return the u8 {0,1} map of pixels where the planted blobs add >= 25 grey levels."""
    rs = np.random.RandomState(index)
    h = w = int(size)
    img = 20.0 + 3.0 * rs.standard_normal((h, w)).astype(np.float32)
    blobs = np.zeros((h, w), np.float32)
    if n_droplets is None:
        n_droplets = max(1, int(round(60 * (h * w) / (256.0 * 256.0))))
    amp = rs.uniform(60, 200, n_droplets)
    sig = rs.uniform(1.5, 5.0, n_droplets)
    cy = rs.uniform(0, h, n_droplets)
    cx = rs.uniform(0, w, n_droplets)
    for a, s, y, x in zip(amp, sig, cy, cx):
        r = int(math.ceil(4 * s))
        y0, y1 = max(0, int(y) - r), min(h, int(y) + r + 1)
        x0, x1 = max(0, int(x) - r), min(w, int(x) + r + 1)
        yy = np.arange(y0, y1, dtype=np.float32)[:, None] - np.float32(y)
        xx = np.arange(x0, x1, dtype=np.float32)[None, :] - np.float32(x)
        blobs[y0:y1, x0:x1] += np.float32(a) * np.exp(-(yy * yy + xx * xx) / np.float32(2 * s * s))
    out = np.clip(np.rint(img + blobs), 0, 255).astype(np.uint8)
    if with_truth:
        return out, (blobs >= 25.0).astype(np.uint8)
    return out


def synthetic_rgb(size: int, index: int = 0) -> np.ndarray:
    """u8 [size, size, 3]: what ``Image.open(p).convert("RGB")`` gives for a grayscale file (qdb:41)."""
    g = synthetic_image(size, index)
    return np.ascontiguousarray(np.repeat(g[:, :, None], 3, axis=2))


def synthetic_mask(size: int, n_discs: int, seed: int = 0, rmin: int = 2, rmax: int = 8) -> np.ndarray:
    """u8 {0,1} [size, size]: filled discs with uniform centres."""
    rs = np.random.RandomState(seed)
    m = np.zeros((size, size), np.uint8)
    rad = rs.randint(rmin, rmax + 1, n_discs)
    cy = rs.randint(0, size, n_discs)
    cx = rs.randint(0, size, n_discs)
    for r, y, x in zip(rad, cy, cx):
        y0, y1 = max(0, y - r), min(size, y + r + 1)
        x0, x1 = max(0, x - r), min(size, x + r + 1)
        yy = np.arange(y0, y1)[:, None] - y
        xx = np.arange(x0, x1)[None, :] - x
        m[y0:y1, x0:x1] |= (yy * yy + xx * xx <= r * r).astype(np.uint8)
    return m


# ------------------------------------------------------------------------------ checkpoint

_ENC = ["enc1", "enc2", "enc3", "enc4", "bottleneck"]


def default_init_state_dict(seed: int = 0):
    """state_dict of a default-initialised UNetDC(3, 1) (reference models/model_2.py:6-32)."""
    import torch
    from .model import UNetDC
    torch.manual_seed(seed)
    return {k: v.clone() for k, v in UNetDC(3, 1).state_dict().items()}


def calibrated_state_dict(seed: int = 0, calib_size: int = 512, n_calib: int = 1,
                          prob_thresh: float = 0.3, dilations=(1, 2, 4, 8, 16), fit_head: bool = True):
    """Default init + BatchNorm statistics calibration + a least-squares ``out_conv``.

    One training-mode pass (momentum 1.0) over ``n_calib`` synthetic frames sets every BN's
    running_mean / running_var to that batch's statistics (without it the net outputs ~0.476 everywhere).
    ``out_conv`` (the final 1x1 conv: 64 weights + 1 bias) is then fitted by ridge regression on the
    random 64-channel features so that the logit is about +4 on the planted droplets and -6 elsewhere:
    the 31 M convolution weights stay random-init, but the mask has droplet-like components
    (IoU ~0.74 with the planted droplets, ~3 k components per 1024^2 frame at prob_thresh 0.3) instead of
    ~100 k single-pixel speckles, which is what the labelling stage is sized for.  Calibrate at >= 512 px for
    frames of 1024 px and up (at small sizes the dilated layers see mostly padding and the statistics differ).
    """
    import torch
    import torch.nn.functional as F

    # The statistics and the least-squares fit are CPU reductions whose summation order follows the thread count:
    # pin it, so that every process (any rank count, any box) manufactures bit-identical weights.
    prev_threads = torch.get_num_threads()
    torch.set_num_threads(4)
    try:
        return _calibrated_state_dict(torch, F, seed, calib_size, n_calib, prob_thresh, dilations, fit_head)
    finally:
        torch.set_num_threads(prev_threads)


def _calibrated_state_dict(torch, F, seed, calib_size, n_calib, prob_thresh, dilations, fit_head):
    sd = default_init_state_dict(seed)
    frames, truth = [], []
    for i in range(n_calib):
        g, tr = synthetic_image(calib_size, 1000 + i, with_truth=True)
        g = g.astype(np.float32)
        g = np.clip(g - (np.median(g) - 9.0), 0.0, None)              # stand-in for the rolling-ball correction
        g = (g - g.min()) / max(1.0, float(g.max() - g.min()))
        frames.append(np.repeat(g[None], 3, axis=0))
        truth.append(tr)
    t = torch.from_numpy(np.stack(frames))

    def cbr(t, p, idx, d):
        t = F.conv2d(t, sd[f"{p}.{idx}.weight"], sd[f"{p}.{idx}.bias"], padding=d, dilation=d)
        mean = t.mean(dim=(0, 2, 3))
        var_b = t.var(dim=(0, 2, 3), unbiased=False)
        n = t.numel() // t.shape[1]
        sd[f"{p}.{idx + 1}.running_mean"] = mean.clone()
        sd[f"{p}.{idx + 1}.running_var"] = (var_b * n / max(1, n - 1)).clone()   # BN stores the unbiased one
        sd[f"{p}.{idx + 1}.num_batches_tracked"] = torch.tensor(1, dtype=torch.long)
        t = (t - mean[None, :, None, None]) / torch.sqrt(var_b[None, :, None, None] + 1e-5)
        t = t * sd[f"{p}.{idx + 1}.weight"][None, :, None, None] + sd[f"{p}.{idx + 1}.bias"][None, :, None, None]
        return F.relu(t)

    with torch.no_grad():
        skips = []
        for i, name in enumerate(_ENC[:4]):
            t = cbr(cbr(t, name, 0, dilations[i]), name, 3, dilations[i])
            skips.append(t)
            t = F.max_pool2d(t, 2)
        t = cbr(cbr(t, "bottleneck", 0, dilations[4]), "bottleneck", 3, dilations[4])
        for lvl in (4, 3, 2, 1):
            t = F.conv_transpose2d(t, sd[f"upconv{lvl}.weight"], sd[f"upconv{lvl}.bias"], stride=2)
            t = torch.cat([t, skips[lvl - 1]], dim=1)
            t = cbr(cbr(t, f"dec{lvl}", 0, 1), f"dec{lvl}", 3, 1)
        if fit_head:
            X = t.permute(0, 2, 3, 1).reshape(-1, 64).double()
            X = torch.cat([X, torch.ones(X.shape[0], 1, dtype=torch.float64)], 1)
            y = torch.from_numpy(np.stack(truth).reshape(-1).astype(np.float64)) * 10.0 - 6.0
            wgt = 1.0 + 2.0 * (y > 0).double()                          # droplets are ~10 % of the pixels
            A = (X * wgt[:, None]).T @ X + 1e-3 * X.shape[0] * torch.eye(65, dtype=torch.float64)
            sol = torch.linalg.solve(A, (X * wgt[:, None]).T @ y)
            sd["out_conv.weight"] = sol[:64].float().reshape(1, 64, 1, 1).clone()
            sd["out_conv.bias"] = sol[64:].float().clone()
        else:
            logits = F.conv2d(t, sd["out_conv.weight"], sd["out_conv.bias"])
            q = torch.quantile(logits.flatten(), 0.85)
            sd["out_conv.bias"] = sd["out_conv.bias"] + (math.log(prob_thresh / (1.0 - prob_thresh)) - q)
    return sd
