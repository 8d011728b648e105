"""dc_roi_mask / dc_radial_density / dc_spatial_density against the golden outputs of the reference's own
quantify_pipline.py functions, against the oracle restatement, and against OpenCV / scipy directly: bit-exact."""
import numpy as np
import pytest

import oracle
from conftest import load_golden

pytestmark = pytest.mark.gpu

CASES = ["tissue_96x128", "border_128x128", "wide_64x200", "small_33x29", "flat_image", "no_droplets"]


@pytest.mark.parametrize("name", CASES)
def test_density_functions_match_reference_golden(cuda_device, name):
    from unet_dc_segmentation_b200 import density
    g = load_golden("density.npz")
    img, mask = g[f"{name}/img"], g[f"{name}/mask"]
    roi = density.generate_roi_mask(img)
    np.testing.assert_array_equal(roi, g[f"{name}/roi"])
    cy, cx = (int(v) for v in g[f"{name}/centroid"])
    np.testing.assert_array_equal(density.get_targets(mask, roi, 10, cy, cx), g[f"{name}/radial"])
    np.testing.assert_array_equal(density.density_maps(mask, roi), g[f"{name}/spatial"])


def _tissue(rs, H, W):
    from scipy import ndimage as ndi
    base = ndi.gaussian_filter(rs.rand(H, W), rs.choice([3, 8, 15]))
    base = (base - base.min()) / max(float(np.ptp(base)), 1e-9)
    return np.clip(base[..., None] * rs.randint(100, 256, 3)[None, None] + rs.randn(H, W, 3) * 8, 0, 255).astype(np.uint8)


def test_roi_mask_batches_match_opencv(cuda_device):
    import cv2
    import torch
    from unet_dc_segmentation_b200 import density
    rs = np.random.RandomState(8)
    kernel = np.ones((15, 15), np.uint8)
    for (B, H, W) in [(3, 64, 80), (2, 257, 300), (4, 5, 9), (1, 1, 1), (2, 33, 513)]:
        imgs = np.stack([_tissue(rs, H, W) for _ in range(B)])
        roi, cen = density.roi_mask_device(torch.from_numpy(imgs).cuda())
        roi, cen = roi.cpu().numpy(), cen.cpu().numpy()
        for b in range(B):
            blurred = cv2.GaussianBlur(cv2.cvtColor(imgs[b], cv2.COLOR_RGB2GRAY), (15, 15), 0)
            _, m = cv2.threshold(blurred, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
            m = cv2.morphologyEx(cv2.morphologyEx(m, cv2.MORPH_CLOSE, kernel), cv2.MORPH_OPEN, kernel)
            want = (m > 0).astype(np.uint8)
            np.testing.assert_array_equal(roi[b], want, err_msg=f"{(B, H, W)} image {b}")
            M = cv2.moments(want)
            wc = (int(M["m01"] / M["m00"]), int(M["m10"] / M["m00"])) if M["m00"] else (H // 2, W // 2)
            assert tuple(cen[b]) == wc


def test_radial_and_spatial_batches_match_oracle_and_scipy(cuda_device):
    import torch
    from scipy.ndimage import gaussian_filter
    from unet_dc_segmentation_b200 import density, label_stats_device
    rs = np.random.RandomState(9)
    for (B, H, W) in [(3, 96, 128), (2, 40, 300), (1, 300, 41), (2, 16, 16)]:
        imgs = np.stack([_tissue(rs, H, W) for _ in range(B)])
        masks = (rs.rand(B, H, W) < 0.04).astype(np.uint8)
        masks[-1] = 0 if B > 1 else masks[-1]                       # a frame without droplets
        roi_d, cen_d = density.roi_mask_device(torch.from_numpy(imgs).cuda())
        m_d = torch.from_numpy(masks).cuda()
        t = label_stats_device(m_d, 1, None, 8192)
        radial = density.radial_density_device(roi_d, cen_d, t, 10).cpu().numpy()
        spatial = density.spatial_density_device(m_d, roi_d, 21).cpu().numpy()
        roi, cen = roi_d.cpu().numpy(), cen_d.cpu().numpy()
        for b in range(B):
            cy, cx = (int(v) for v in cen[b])
            np.testing.assert_array_equal(radial[b], oracle.radial_density(masks[b], roi[b], 10, cy, cx), err_msg=f"radial {b}")
            want = gaussian_filter(masks[b].astype(np.float32), sigma=21 / 6)
            want = want / (gaussian_filter(roi[b].astype(np.float32), sigma=21 / 6) + 1e-5)
            want *= 100
            np.testing.assert_array_equal(spatial[b], want, err_msg=f"spatial {b}")


def test_radial_density_other_ring_counts_and_off_centre(cuda_device):
    from unet_dc_segmentation_b200 import density
    rs = np.random.RandomState(10)
    H, W = 90, 70
    roi = np.zeros((H, W), np.uint8)
    roi[10:80, 5:60] = 1
    mask = (rs.rand(H, W) < 0.05).astype(np.uint8)
    for nb, (cy, cx) in [(1, (45, 30)), (3, (0, 0)), (10, (89, 69)), (64, (45, 30))]:
        np.testing.assert_array_equal(density.get_targets(mask, roi, nb, cy, cx), oracle.radial_density(mask, roi, nb, cy, cx))
    assert not density.get_targets(mask, np.zeros_like(roi), 10, 45, 30).any()        # empty ROI


def test_density_argument_checks(cuda_device):
    import torch
    from unet_dc_segmentation_b200 import density
    with pytest.raises(TypeError):
        density.roi_mask_device(torch.zeros((1, 8, 8), dtype=torch.uint8, device="cuda"))
    with pytest.raises(RuntimeError):
        density.roi_mask_device(torch.zeros((1, 8, 8, 3), dtype=torch.uint8))
    with pytest.raises(ValueError):
        density.generate_roi_mask(np.zeros((8, 8, 3), np.uint8), blur_kernel=5)
    with pytest.raises(RuntimeError, match="nb_layers"):
        density.get_targets(np.zeros((8, 8), np.uint8), np.ones((8, 8), np.uint8), 65, 4, 4)
