"""numpy model of csrc/morph.cu's morph_chord_kernel, driven by the ACTUAL plan the C++ host code builds
(dc_debug_rolling_ball_plan, host only).  Test infrastructure: it pins the plan builder (window fetch offsets and
shifts, chord nesting, row grouping, table levels) on the CPU, where the kernel itself cannot run."""
from __future__ import annotations

import ctypes as C

import numpy as np

TW = 64


def load_plan(radius: int, th: int) -> dict:
    from unet_dc_segmentation_b200 import _lib
    lib = _lib.load()
    buf = (C.c_int * 4096)()
    n = lib.dc_debug_rolling_ball_plan(radius, th, buf, 4096)
    if n < 0:
        raise RuntimeError(lib.dc_last_error().decode())
    v = list(buf[:n])
    keys = ["k", "an", "pad", "pitch", "RH", "level_words", "ntables", "nchords", "th", "HP"]
    p = dict(zip(keys, v[:10]))
    pos = 10
    p["chords"] = []
    for _ in range(p["nchords"]):
        nf = v[pos]
        f = [(v[pos + 1 + 2 * q], v[pos + 2 + 2 * q]) for q in range(4)]
        p["chords"].append({"nf": nf, "f": f, "rows": (v[pos + 9], v[pos + 10])})
        pos += 11
    p["rowoff"] = v[pos:]
    return p


def morph_pass(img: np.ndarray, plan: dict, is_max: bool) -> np.ndarray:
    """One erode (is_max False) / dilate pass over a u8 [H,W] plane, tile by tile, exactly as the kernel does it."""
    H, W = img.shape
    an, pad, pitch, RH, th, HP = (plan[k] for k in ("an", "pad", "pitch", "RH", "th", "HP"))
    ident = 0 if is_max else 255
    op = np.maximum if is_max else np.minimum
    out = np.empty_like(img)
    rowbytes = pitch * 4
    for ty in range(0, H, th):
        for tx in range(0, W, TW):
            x0, y0 = tx - an - pad, ty - an
            # range table 0: region bytes, identity outside the image
            t0 = np.full((RH, rowbytes), ident, np.uint8)
            ys = np.arange(RH) + y0
            xs = np.arange(rowbytes) + x0
            vy = (ys >= 0) & (ys < H)
            vx = (xs >= 0) & (xs < W)
            t0[np.ix_(vy, vx)] = img[np.ix_(ys[vy], xs[vx])]
            tables = [t0]
            for l in range(1, plan["ntables"]):
                prev = tables[-1]
                step = 1 << (l - 1)
                sh = np.full_like(prev, ident)
                sh[:, :rowbytes - step] = prev[:, step:]
                tables.append(op(prev, sh))
            flat = np.concatenate([t.reshape(-1) for t in tables])       # word-addressed like the kernel's smem
            hcur = np.full((RH, TW), ident, np.uint8)                     # running chord minimum per region row
            acc = np.full((th, TW), ident, np.uint8)
            rr = np.arange(RH)[:, None]
            xx = np.arange(TW)[None, :]
            for ch in plan["chords"]:
                slots = [0, 1] + ([2, 3] if ch["nf"] > 2 else [])
                for q in slots:
                    woff, shift = ch["f"][q]
                    assert shift in (0, 8, 16, 24)
                    # lane (row r, word xw): p = table0 + r * pitch + xw; window = bytes of words p[woff], p[woff + 1]
                    byte0 = (rr * pitch + woff) * 4 + shift // 8 + xx                 # xx = 4 * xw + pixel in word
                    # the two-word window must stay inside the row of the table it addresses
                    lvl = woff // plan["level_words"]
                    col = (woff % plan["level_words"])
                    assert col + (TW // 4 - 1) + 1 < pitch, "window fetch leaves the region row"
                    assert lvl < plan["ntables"]
                    hcur = op(hcur, flat[byte0])
                b, e = ch["rows"]
                for j in range(b, e):
                    o = plan["rowoff"][j]
                    assert o % HP == 0
                    r0 = o // HP                                          # dy + an
                    acc = op(acc, hcur[r0:r0 + th])
            hh, ww = min(th, H - ty), min(TW, W - tx)
            out[ty:ty + hh, tx:tx + ww] = acc[:hh, :ww]
    return out


def rolling_ball_plane(img: np.ndarray, radius: int, th: int = 128) -> np.ndarray:
    """erode -> dilate -> saturating subtract (the stretch is not modelled here)."""
    plan = load_plan(radius, th)
    er = morph_pass(img, plan, False)
    bg = morph_pass(er, plan, True)
    return np.where(img > bg, img - bg, 0).astype(np.uint8)
