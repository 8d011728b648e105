"""Golden vectors for the density maps of the reference's alternate front end, quantify_pipline.py,
made by its OWN functions imported unmodified from /root/reference (build container only):

    generate_roi_mask :44-51, the cv2.moments centroid :133-136, get_targets :61-91, density_maps :93-97

    python tests/golden/make_golden_density.py

Shims as in make_golden.py (matplotlib, albumentations, skimage.measure over scipy.ndimage).
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_golden import REF, install_shims   # noqa: E402


def cases():
    from scipy import ndimage as ndi
    rs = np.random.RandomState(5)
    out = []
    for name, (H, W), blob_sigma, n_drop in (("tissue_96x128", (96, 128), 14, 60), ("border_128x128", (128, 128), 25, 150),
                                              ("wide_64x200", (64, 200), 9, 40), ("small_33x29", (33, 29), 6, 8)):
        base = ndi.gaussian_filter(rs.rand(H, W), blob_sigma)
        base = (base - base.min()) / max(float(np.ptp(base)), 1e-9)
        img = np.clip((base > 0.5)[..., None] * rs.randint(90, 230, 3)[None, None] + 25 + rs.randn(H, W, 3) * 6, 0, 255).astype(np.uint8)
        mask = np.zeros((H, W), np.uint8)
        for _ in range(n_drop):
            cy, cx, r = rs.randint(0, H), rs.randint(0, W), rs.randint(1, 4)
            yy, xx = np.ogrid[:H, :W]
            mask[(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = 1
        out.append((name, img, mask))
    H, W = 40, 56
    out.append(("flat_image", np.full((H, W, 3), 77, np.uint8), (rs.rand(H, W) < 0.05).astype(np.uint8)))
    name, img, mask = out[0]
    out.append(("no_droplets", img, np.zeros_like(mask)))
    return out


def main():
    install_shims()
    for sub in ("", "models", "utils"):
        sys.path.insert(0, str(REF / sub))
    import cv2
    import quantify_pipline as qp

    g = {}
    for name, img, mask in cases():
        roi = qp.generate_roi_mask(img)                                          # :132
        M = cv2.moments(roi)                                                     # :133
        oh, ow = mask.shape
        cx = int(M["m10"] / M["m00"]) if M["m00"] else ow // 2                   # :134
        cy = int(M["m01"] / M["m00"]) if M["m00"] else oh // 2                   # :135
        g[f"{name}/img"] = img
        g[f"{name}/mask"] = mask
        g[f"{name}/roi"] = roi
        g[f"{name}/centroid"] = np.array([cy, cx], np.int64)
        g[f"{name}/radial"] = qp.get_targets(mask, roi, 10, cy, cx)             # :138
        g[f"{name}/spatial"] = qp.density_maps(mask, roi)                        # :139
        print(name, img.shape, "roi px", int(roi.sum()), "centroid", (cy, cx), "radial max", float(g[f"{name}/radial"].max()),
              "spatial max", float(g[f"{name}/spatial"].max()))
    np.savez_compressed(HERE / "density.npz", **g)


if __name__ == "__main__":
    main()
