"""CPU oracle for the droplet-quantification hot path -- TEST INFRASTRUCTURE ONLY.

Restates, on the CPU, what the reference (malani86/unet-DC-segmentation) computes on the path
behind ``quantify_droplets_batch.py``.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this package; the
product package ``unet_dc_segmentation_b200`` never does (it fails loudly without its CUDA
library instead of falling back to anything here).

Parity pin: the reference ships no tests for this path, so the pin is
``tests/golden/*.npz`` -- outputs of the reference's OWN functions, generated in the build
container by ``tests/golden/make_golden.py`` (imports /root/reference unmodified under
sys.modules shims), plus the known-answer rows of the reference's ``outputs/all_droplets.csv``.

Integer/byte stages live in ``oracle.c`` (gcc); the network is a plain PyTorch fp32
restatement (``unetdc_forward``), the one floating-point kernel on the path.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    """Compile oracle.c -> liboracle.so with gcc (a few hundred ms)."""
    so = _HERE / "liboracle.so"
    src = _HERE / "oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B", "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(str(build()))
        i32p = ctypes.POINTER(ctypes.c_int32)
        i64p = ctypes.POINTER(ctypes.c_int64)
        f64p = ctypes.POINTER(ctypes.c_double)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        lib.orc_ellipse_rows.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        lib.orc_rolling_ball_u8.argtypes = [u8p, u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        lib.orc_label4.argtypes = [i32p, i32p, ctypes.c_int, ctypes.c_int]
        lib.orc_resize_linear_u8.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, u8p, ctypes.c_int, ctypes.c_int]
        lib.orc_overlay_stencil.argtypes = [u8p, ctypes.c_int, ctypes.c_int, u8p]
        lib.orc_quantify.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_double,
                                     i32p, ctypes.c_int, i64p, f64p, f64p, f64p, f64p, f64p]
        _LIB = lib
    return _LIB


def _ptr(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


# ----------------------------------------------------------------------------- rolling ball

def ellipse_rows(k: int):
    """[j1, j2) column span of each row of cv2.getStructuringElement(MORPH_ELLIPSE, (k, k))."""
    j1 = np.zeros(k, np.intc)
    j2 = np.zeros(k, np.intc)
    _lib().orc_ellipse_rows(k, _ptr(j1, ctypes.c_int), _ptr(j2, ctypes.c_int))
    return j1, j2


def rolling_ball_correction_rgb(image: np.ndarray, radius: int = 50) -> np.ndarray:
    """Follows reference utils/data_loader.py:11-24.  image: u8 [H, W, C]."""
    image = np.ascontiguousarray(image, dtype=np.uint8)
    if image.ndim == 2:
        image = image[:, :, None]
    H, W, C = image.shape
    out = np.empty_like(image)
    rc = _lib().orc_rolling_ball_u8(_ptr(image, ctypes.c_uint8), _ptr(out, ctypes.c_uint8), H, W, C, int(radius))
    if rc != 0:
        raise RuntimeError(f"orc_rolling_ball_u8 failed: {rc}")
    return out


def resize_linear_u8(src: np.ndarray, dsize) -> np.ndarray:
    """cv2.resize(src, dsize) with the default INTER_LINEAR on u8 -- what reference qdb:44 and qdb:57 effectively
    call (their interpolation flag lands in the `dst` slot).  dsize = (width, height) as in cv2."""
    src = np.ascontiguousarray(src, dtype=np.uint8)
    dw, dh = int(dsize[0]), int(dsize[1])
    sh, sw = src.shape[:2]
    cn = 1 if src.ndim == 2 else src.shape[2]
    out = np.empty((dh, dw) if src.ndim == 2 else (dh, dw, cn), np.uint8)
    rc = _lib().orc_resize_linear_u8(_ptr(src, ctypes.c_uint8), sh, sw, cn, _ptr(out, ctypes.c_uint8), dh, dw)
    if rc != 0:
        raise RuntimeError(f"orc_resize_linear_u8 failed: {rc}")
    return out


def overlay_stencil(mask: np.ndarray) -> np.ndarray:
    """Pixels painted by cv2.drawContours(img, cv2.findContours(mask, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)[0], -1,
    color, 2) -- reference qdb:76-77 -- as a u8 {0,1} map."""
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    H, W = mask.shape
    out = np.empty((H, W), np.uint8)
    rc = _lib().orc_overlay_stencil(_ptr(mask, ctypes.c_uint8), H, W, _ptr(out, ctypes.c_uint8))
    if rc != 0:
        raise RuntimeError(f"orc_overlay_stencil failed: {rc}")
    return out


# ----------------------------------------------------------------------------- labelling / table

def label4(img: np.ndarray) -> tuple[np.ndarray, int]:
    """skimage.measure.label(img, connectivity=1) semantics (reference qdb:82,86)."""
    img = np.ascontiguousarray(img, dtype=np.int32)
    H, W = img.shape
    out = np.empty((H, W), np.int32)
    n = _lib().orc_label4(_ptr(img, ctypes.c_int32), _ptr(out, ctypes.c_int32), H, W)
    if n < 0:
        raise RuntimeError(f"orc_label4 failed: {n}")
    return out, n


COLUMNS = ["label", "area", "equivalent_diameter", "centroid-0", "centroid-1"]
MICRON_COLUMNS = ["area_sqmicron", "eq_diam_micron"]


def quantify_arrays(bin_mask: np.ndarray, min_area: int = 1, px_per_um: float | None = None):
    """Reference quantify() (qdb:81-95) as plain arrays: (labels int32 [H,W], dict of columns)."""
    mask = np.ascontiguousarray(bin_mask, dtype=np.uint8)
    H, W = mask.shape
    cap = max(1, (H * W + 1) // 2)
    labels = np.empty((H, W), np.int32)
    area = np.empty(cap, np.int64)
    c0, c1, dia, aum, dum = (np.empty(cap, np.float64) for _ in range(5))
    n = _lib().orc_quantify(_ptr(mask, ctypes.c_uint8), H, W, int(min_area),
                            float(px_per_um) if px_per_um else 0.0,
                            _ptr(labels, ctypes.c_int32), cap, _ptr(area, ctypes.c_int64),
                            _ptr(c0, ctypes.c_double), _ptr(c1, ctypes.c_double), _ptr(dia, ctypes.c_double),
                            _ptr(aum, ctypes.c_double), _ptr(dum, ctypes.c_double))
    if n < 0:
        raise RuntimeError(f"orc_quantify failed: {n}")
    cols = {"label": np.arange(1, n + 1, dtype=np.int64), "area": area[:n].copy(),
            "equivalent_diameter": dia[:n].copy(), "centroid-0": c0[:n].copy(), "centroid-1": c1[:n].copy()}
    if px_per_um:
        cols["area_sqmicron"] = aum[:n].copy()
        cols["eq_diam_micron"] = dum[:n].copy()
    return labels, cols


def quantify(bin_mask: np.ndarray, min_area: int = 1, px_per_um: float | None = None):
    """Reference quantify() (qdb:81-95) -> pandas.DataFrame (empty, column-less, when no droplets)."""
    import pandas as pd
    _, cols = quantify_arrays(bin_mask, min_area, px_per_um)
    if len(cols["label"]) == 0:
        return pd.DataFrame()
    return pd.DataFrame(cols)


# ----------------------------------------------------------------------------- network (fp32)

_BLOCKS = [("enc1", 1), ("enc2", 2), ("enc3", 4), ("enc4", 8), ("bottleneck", 16)]


def composed_upconv_conv3x3(x, skip, comp, skipw, bias9, relu=True):
    """CPU evaluation of a fused upconv + conv layer (include/unetdc_b200.h dc_conv_upfused) from the composed weights:
    x f32 [B,2C,H,W], skip f32 [B,C,2H,2W], comp [4 classes][2][2][C][2C], skipw [3][3][C][C], bias9 [9,C]
    -> f32 [B,C,2H,2W].  Equal (up to rounding) to conv_transpose2d -> cat -> conv2d(padding=1) of
    models/model_2.py:76-77, restated per output parity class: class (py, px) is a 2x2 conv over x zero-padded by one
    pixel, read from (py, px), plus the 3x3 over skip, plus the bias of the pixel's border class."""
    import torch
    import torch.nn.functional as F

    B, _, H, W = x.shape
    C = skip.shape[1]
    xp = F.pad(x, (1, 1, 1, 1))
    out = torch.zeros((B, C, 2 * H, 2 * W), dtype=torch.float32)
    for py in range(2):
        for px in range(2):
            k = comp[py * 2 + px].to(torch.float32).permute(2, 3, 0, 1).contiguous()       # [co][cx][a][b]
            out[:, :, py::2, px::2] = F.conv2d(xp[:, :, py:py + H + 1, px:px + W + 1], k)
    out += F.conv2d(skip, skipw.to(torch.float32).permute(2, 3, 0, 1).contiguous(), padding=1)
    b9 = bias9.to(torch.float32).reshape(3, 3, C)
    rows = torch.ones(2 * H, dtype=torch.long); rows[0] = 0; rows[-1] = 2
    cols = torch.ones(2 * W, dtype=torch.long); cols[0] = 0; cols[-1] = 2
    out += b9[rows][:, cols].permute(2, 0, 1).unsqueeze(0)
    return F.relu(out) if relu else out


def unetdc_forward(state_dict, x, dilations=(1, 2, 4, 8, 16), emulate_bf16=False, gray_input=False, round_last=False,
                   fused_level1=None, fused_levels=None):
    """Plain PyTorch fp32 restatement of reference models/model_2.py:56-80 in eval mode.

    state_dict uses the reference's 136 keys; x: f32 [B,3,H,W]; returns f32 [B,1,H,W] probs.
    ``dilations`` = (1,1,1,1,1) gives reference models/model.py:35-50 (plain UNet).

    ``emulate_bf16``: the same network with values rounded to bf16 at exactly the points where the CUDA
    path stores bf16 (BatchNorm folded into the conv in fp32, folded weights -> bf16, every stored
    activation -> bf16, fp32 accumulation, the last feature map and the 1x1 head kept in fp32).  It separates
    "the kernels compute what they claim" (GPU vs this, tight) from "bf16 storage vs the fp32 reference"
    (this vs fp32, inherent to the precision choice).  ``gray_input``: x holds u8 grey levels / 255 replicated
    to 3 channels; the CUDA stem then feeds the exact integers and folds 1/255 and the 3 channels into the weights.
    ``round_last``: models with out_channels != 1 store the last feature map in bf16 before the 1x1 head kernel.
    ``fused_level1`` (emulation only): (comp, skipw, bias9) of the composed upconv1 + dec1.0 layer the CUDA path runs by
    default (model.compose_upconv; None = the two layers separately, `up` stored in bf16); ``fused_levels``: the same
    for any decoder level, {level: (comp, skipw, bias9)}."""
    import torch
    import torch.nn.functional as F

    sd = {k: v.detach().to(torch.float32).cpu() for k, v in state_dict.items()}
    r = (lambda t: t.to(torch.bfloat16).to(torch.float32)) if emulate_bf16 else (lambda t: t)

    def cbr(t, p, idx, d, first=False):                     # model_2.py:40-54 (one conv+BN+ReLU)
        w, b = sd[f"{p}.{idx}.weight"], sd[f"{p}.{idx}.bias"]
        g, beta = sd[f"{p}.{idx + 1}.weight"], sd[f"{p}.{idx + 1}.bias"]
        mu, var = sd[f"{p}.{idx + 1}.running_mean"], sd[f"{p}.{idx + 1}.running_var"]
        if not emulate_bf16:
            t = F.conv2d(t, w, b, padding=d, dilation=d)
            t = F.batch_norm(t, mu, var, g, beta, False, 0.0, 1e-5)
            return F.relu(t)
        scale = g / torch.sqrt(var + 1e-5)                  # fold in fp32, then round the weights
        wf = w * scale.view(-1, 1, 1, 1)
        bf = (b - mu) * scale + beta
        if first and gray_input:
            wq = r(wf.sum(1, keepdim=True) * (1.0 / 255.0))   # 3 identical channels -> one; 1/255 into the weights
            t = torch.round(t[:, :1] * 255.0)                 # the exact u8 grey levels
            return F.relu(F.conv2d(t, wq, bf, padding=d, dilation=d))
        return F.relu(F.conv2d(r(t), r(wf), bf, padding=d, dilation=d))

    def block(t, p, d, first=False, round_out=True):
        t = r(cbr(t, p, 0, d, first))
        t = cbr(t, p, 3, d)
        return r(t) if round_out else t

    fused_levels = dict(fused_levels or {})
    if fused_level1 is not None:
        fused_levels[1] = fused_level1
    with torch.no_grad():
        x = x.detach().to(torch.float32).cpu()
        skips = []
        t = x
        for i, (name, _) in enumerate(_BLOCKS[:4]):          # model_2.py:58-61
            t = block(t, name, dilations[i], first=(i == 0))
            skips.append(t)
            t = F.max_pool2d(t, 2)
        t = block(t, "bottleneck", dilations[4])            # model_2.py:64
        for lvl in (4, 3, 2, 1):                            # model_2.py:67-77
            fl = fused_levels.get(lvl) if emulate_bf16 else None
            if fl is not None:
                t = r(composed_upconv_conv3x3(t, skips[lvl - 1], *(b.cpu() for b in fl)))
                t = cbr(t, f"dec{lvl}", 3, 1)
                t = r(t) if (lvl > 1 or round_last) else t
                continue
            t = r(F.conv_transpose2d(t, r(sd[f"upconv{lvl}.weight"]), sd[f"upconv{lvl}.bias"], stride=2))
            t = torch.cat([t, skips[lvl - 1]], dim=1)
            t = block(t, f"dec{lvl}", 1, round_out=(lvl > 1 or round_last))
        t = F.conv2d(t, sd["out_conv.weight"], sd["out_conv.bias"])   # model_2.py:79
        return torch.sigmoid(t)                                        # model_2.py:80


def run_path(state_dict, images_u8, radius=50, prob_thresh=0.3, min_area=1, px_per_um=None,
             use_cv2=False):
    """Whole reference path at native size (IMG_SIZE == image size, identity resizes):
    preprocess (qdb:40-46) -> forward (qdb:52) -> threshold (qdb:56) -> quantify (qdb:61).

    images_u8: list of u8 [H,W,3].  Returns (probs f32 [B,H,W], masks u8 [B,H,W], tables)."""
    import torch
    if use_cv2:
        import cv2
        def rb(im, r):
            k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (r, r))
            chans = []
            for ch in cv2.split(im):
                bg = cv2.morphologyEx(ch, cv2.MORPH_OPEN, k)
                chans.append(cv2.normalize(cv2.subtract(ch, bg), None, 0, 255, cv2.NORM_MINMAX))
            return cv2.merge(chans)
    else:
        rb = rolling_ball_correction_rgb
    pre = [rb(im, radius).astype(np.float32) / 255.0 for im in images_u8]
    batch = torch.from_numpy(np.stack(pre)).permute(0, 3, 1, 2).contiguous()
    probs = unetdc_forward(state_dict, batch)[:, 0].numpy()
    masks = (probs > prob_thresh).astype(np.uint8)
    tables = [quantify(m, min_area, px_per_um) for m in masks]
    return probs, masks, tables


# ----------------------------------------------------------------------------- density maps (quantify_pipline.py)
# numpy restatement of the alternate front end's per-image maps.  Third-party arithmetic restated:
# OpenCV's u8 cvtColor / GaussianBlur / Otsu / rectangular morphology / moments and scipy.ndimage.gaussian_filter
# (neither vendored in /root/reference); pinned in tests/test_oracle_golden.py against cv2 / scipy themselves and
# against the reference's own functions (tests/golden/density.npz).

_GAUSS15_Q8 = np.array([1, 3, 6, 12, 20, 30, 36, 40, 36, 30, 20, 12, 6, 3, 1], np.int64)   # sums to 256


def _reflect101(i, n):
    """cv2 BORDER_REFLECT_101 (gfedcb|abcdefgh|gfedcba)."""
    if n == 1:
        return np.zeros_like(i)
    p = 2 * (n - 1)
    i = np.mod(i, p)
    return np.where(i >= n, p - i, i)


def _rect15(a, op, pad):
    """15x15 rectangular erode (np.minimum, pad 255) / dilate (np.maximum, pad 0), out-of-image taps ignored."""
    H, W = a.shape
    e = np.pad(a, ((0, 0), (7, 7)), constant_values=pad)
    h = e[:, 0:W].copy()
    for j in range(1, 15):
        h = op(h, e[:, j:j + W])
    e = np.pad(h, ((7, 7), (0, 0)), constant_values=pad)
    v = e[0:H].copy()
    for j in range(1, 15):
        v = op(v, e[j:j + H])
    return v


def otsu_threshold_u8(a: np.ndarray) -> int:
    """cv2.threshold(..., THRESH_OTSU)'s threshold for a u8 image (OpenCV getThreshVal_Otsu_8u, f64, same order)."""
    h = np.bincount(np.ascontiguousarray(a, dtype=np.uint8).ravel(), minlength=256)
    scale = 1.0 / a.size
    mu = 0.0
    for i in range(256):
        mu += i * float(h[i])
    mu *= scale
    mu1 = q1 = max_sigma = 0.0
    max_val = 0
    eps = float(np.finfo(np.float32).eps)
    for i in range(256):
        p_i = float(h[i]) * scale
        mu1 *= q1
        q1 += p_i
        q2 = 1.0 - q1
        if min(q1, q2) < eps or max(q1, q2) > 1.0 - eps:
            continue
        mu1 = (mu1 + i * p_i) / q1
        mu2 = (mu - q1 * mu1) / q2
        sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2)
        if sigma > max_sigma:
            max_sigma, max_val = sigma, i
    return max_val


def roi_blurred_gray(img_rgb: np.ndarray) -> np.ndarray:
    """cv2.GaussianBlur(cv2.cvtColor(img, COLOR_RGB2GRAY), (15, 15), 0)  (quantify_pipline.py:45-46): 15-bit gray
    coefficients, the 8-bit error-diffused kernel of sigma 2.6, rows then columns in integers, one rounding."""
    img = np.ascontiguousarray(img_rgb, dtype=np.uint8)
    r, g, b = (img[..., c].astype(np.int64) for c in range(3))
    gray = (r * 9798 + g * 19235 + b * 3735 + (1 << 14)) >> 15
    H, W = gray.shape
    e = gray[:, _reflect101(np.arange(-7, W + 7), W)]
    h = sum(e[:, j:j + W] * _GAUSS15_Q8[j] for j in range(15))
    e = h[_reflect101(np.arange(-7, H + 7), H), :]
    v = sum(e[j:j + H, :] * _GAUSS15_Q8[j] for j in range(15))
    return ((v + (1 << 15)) >> 16).astype(np.uint8)


def roi_mask(img_rgb: np.ndarray):
    """generate_roi_mask (quantify_pipline.py:44-51) + the centroid of quantify_pipline.py:134-137.
    Returns (u8 {0,1} [H,W], cy, cx)."""
    b = roi_blurred_gray(img_rgb)
    m = np.where(b > otsu_threshold_u8(b), 255, 0).astype(np.uint8)                 # :47
    m = _rect15(_rect15(m, np.maximum, 0), np.minimum, 255)                          # :49 MORPH_CLOSE
    m = _rect15(_rect15(m, np.minimum, 255), np.maximum, 0)                          # :50 MORPH_OPEN
    roi = (m > 0).astype(np.uint8)                                                   # :51
    H, W = roi.shape
    ys, xs = np.nonzero(roi)
    m00 = float(len(ys))
    cx = int(float(xs.sum()) / m00) if m00 else W // 2                               # :135
    cy = int(float(ys.sum()) / m00) if m00 else H // 2                               # :136
    return roi, cy, cx


def radial_density(mask: np.ndarray, roi: np.ndarray, nb_layers: int, cy: int, cx: int) -> np.ndarray:
    """get_targets (quantify_pipline.py:61-91): droplets per concentric ring painted onto the ring's ROI pixels."""
    _, cols = quantify_arrays(np.ascontiguousarray(mask, dtype=np.uint8), 1, None)   # :66-68 label + centroids
    c0, c1 = np.asarray(cols["centroid-0"], np.float64), np.asarray(cols["centroid-1"], np.float64)
    out = np.zeros(mask.shape, np.float32)
    ys, xs = np.nonzero(roi)
    if len(ys) == 0 or len(c0) == 0:                                                 # :71-72
        return out
    d = np.sqrt(((xs - cx) ** 2 + (ys - cy) ** 2).astype(np.float64))                # :75
    bounds = np.linspace(0, d.max(), nb_layers + 1)                                  # :76-77
    dc = np.sqrt((c1 - cx) ** 2 + (c0 - cy) ** 2)                                    # :80
    for i in range(nb_layers):                                                       # :83-90
        ring = (bounds[i] < d) & (d <= bounds[i + 1])
        if ring.any():
            out[ys[ring], xs[ring]] = np.sum((bounds[i] < dc) & (dc <= bounds[i + 1]))
    return out


def gaussian_weights(sigma: float, truncate: float = 4.0):
    """scipy.ndimage._filters._gaussian_kernel1d (order 0)."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum(), radius


def gaussian_filter_f32(a: np.ndarray, sigma: float) -> np.ndarray:
    """scipy.ndimage.gaussian_filter(float32 image, sigma) with the default mode='reflect': axis 0 then axis 1, each
    pass accumulated in f64 in NI_Correlate1D's symmetric order (centre, then tap pairs from the outside in) and
    rounded to f32."""
    w, r = gaussian_weights(sigma)
    out = np.ascontiguousarray(a, dtype=np.float32)
    for axis in (0, 1):
        x = np.moveaxis(out, axis, -1).astype(np.float64)
        n = x.shape[-1]
        idx = np.mod(np.arange(-r, n + r), 2 * n)
        idx = np.where(idx >= n, 2 * n - 1 - idx, idx)                               # (d c b a | a b c d | d c b a)
        e = x[..., idx]
        acc = e[..., r:r + n] * w[r]
        for jj in range(-r, 0):
            acc = acc + (e[..., r + jj:r + jj + n] + e[..., r - jj:r - jj + n]) * w[r + jj]
        out = np.moveaxis(acc.astype(np.float32), -1, axis)
    return np.ascontiguousarray(out)


def spatial_density(mask: np.ndarray, roi: np.ndarray, kernel_size: int = 21) -> np.ndarray:
    """density_maps (quantify_pipline.py:93-97)."""
    sigma = kernel_size / 6
    d = gaussian_filter_f32(mask.astype(np.float32), sigma)
    d = d / (gaussian_filter_f32(roi.astype(np.float32), sigma) + np.float32(1e-5))
    d *= np.float32(100)
    return d
