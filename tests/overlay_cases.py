"""Shared by the CPU and GPU overlay tests: the reference's two OpenCV calls and the adversarial mask set."""
import numpy as np


def _cv2_overlay_stencil(mask):
    """The reference's own two calls (quantify_droplets_batch.py:76-77) on a black frame."""
    import cv2
    img = np.zeros(mask.shape + (3,), np.uint8)
    cnts, _ = cv2.findContours(mask, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    cv2.drawContours(img, cnts, -1, (0, 255, 0), 2)
    assert not img[..., 0].any() and not img[..., 2].any()
    return (img[..., 1] == 255).astype(np.uint8)


def _overlay_cases():
    from scipy import ndimage as ndi
    rs = np.random.RandomState(11)
    cases = [np.zeros((5, 7), np.uint8), np.ones((6, 4), np.uint8), np.ones((1, 1), np.uint8), np.zeros((1, 1), np.uint8),
             np.eye(9, dtype=np.uint8), np.fliplr(np.eye(9, dtype=np.uint8)).copy(),
             (np.indices((12, 13)).sum(0) % 2).astype(np.uint8)]
    ring = np.zeros((21, 21), np.uint8)                      # nested: ring, island in its hole, speck in the island's hole
    ring[2:19, 2:19] = 1; ring[5:16, 5:16] = 0; ring[8:13, 8:13] = 1; ring[10, 10] = 0
    cases.append(ring)
    cases.append(np.pad(ring, 3)[:, 2:])                     # same away from / touching the frame
    for it in range(120):
        H, W = rs.randint(1, 48, 2)
        if it % 2:
            m = (rs.rand(H, W) < rs.choice([0.05, 0.3, 0.5, 0.7, 0.95])).astype(np.uint8)
        else:
            m = (ndi.gaussian_filter(rs.rand(H, W), rs.choice([1.0, 1.5, 2.5])) > 0.5).astype(np.uint8)
        cases.append(m)
    return cases
