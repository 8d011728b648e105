"""GPU: randomized (hypothesis) parity of the integer stages against the oracle / OpenCV on small inputs --
ragged shapes, every radius, min_area filters, empty and full masks."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import oracle

pytestmark = pytest.mark.gpu
COMMON = dict(deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow], derandomize=True)


@settings(max_examples=60, **COMMON)
@given(h=st.integers(1, 70), w=st.integers(1, 70), p=st.floats(0.0, 1.0), min_area=st.integers(1, 6),
       seed=st.integers(0, 10 ** 6), px=st.sampled_from([None, 2.0, 3.45]))
def test_label_stats_random(cuda_device, h, w, p, min_area, seed, px):
    import torch
    from unet_dc_segmentation_b200 import quantify_arrays
    mask = (np.random.RandomState(seed).rand(h, w) < p).astype(np.uint8)
    tables, labels = quantify_arrays(torch.from_numpy(mask[None]).cuda(), min_area, px, want_labels=True)
    want_l, want = oracle.quantify_arrays(mask, min_area, px)
    np.testing.assert_array_equal(labels[0].cpu().numpy(), want_l)
    assert set(tables[0]) == set(want)
    for c, v in want.items():
        np.testing.assert_array_equal(np.asarray(tables[0][c]), v, err_msg=c)


@settings(max_examples=30, **COMMON)
@given(h=st.integers(1, 90), w=st.integers(1, 90), radius=st.integers(1, 70), seed=st.integers(0, 10 ** 6),
       smooth=st.booleans())
def test_rolling_ball_random(cuda_device, h, w, radius, seed, smooth):
    import torch
    from unet_dc_segmentation_b200 import rolling_ball_device
    rs = np.random.RandomState(seed)
    img = rs.randint(0, 256, (h, w)).astype(np.uint8)
    if smooth:
        img = (np.add.outer(np.arange(h), np.arange(w)) % 256).astype(np.uint8) // 2 + img // 8
    got = rolling_ball_device(torch.from_numpy(img[None]).cuda(), radius)[0].cpu().numpy()
    want = oracle.rolling_ball_correction_rgb(img[:, :, None], radius)[:, :, 0]
    np.testing.assert_array_equal(got, want)


@settings(max_examples=40, **COMMON)
@given(sh=st.integers(1, 80), sw=st.integers(1, 80), dh=st.integers(1, 120), dw=st.integers(1, 120),
       cn=st.sampled_from([1, 3]), seed=st.integers(0, 10 ** 6))
def test_resize_random_vs_cv2(cuda_device, sh, sw, dh, dw, cn, seed):
    import cv2
    import torch
    from unet_dc_segmentation_b200 import resize_linear_u8_device
    shape = (sh, sw, 3) if cn == 3 else (sh, sw)
    src = np.random.RandomState(seed).randint(0, 256, shape).astype(np.uint8)
    got = resize_linear_u8_device(torch.from_numpy(src[None]).cuda(), (dw, dh))[0].cpu().numpy()
    np.testing.assert_array_equal(got, cv2.resize(src, (dw, dh)).reshape(got.shape))
    np.testing.assert_array_equal(got, oracle.resize_linear_u8(src, (dw, dh)))


@settings(max_examples=30, **COMMON)
@given(B=st.integers(1, 3), H=st.integers(1, 40), W=st.integers(1, 40), cin=st.sampled_from([64, 128, 192, 256]),
       cout=st.sampled_from([64, 128, 256, 512]), d=st.integers(1, 6), relu=st.booleans(), seed=st.integers(0, 10 ** 6))
def test_conv3x3_random_shapes(cuda_device, B, H, W, cin, cout, d, relu, seed):
    """Ragged frame sizes and every (Cin, Cout, dilation) family: CTA-pair halo (resident / streamed weights, 1-4
    halves), and the generic per-tap kernel (dilation > 4), against fp32 conv2d on the same bf16 operands."""
    import torch
    import torch.nn.functional as F
    from unet_dc_segmentation_b200 import layers
    from unet_dc_segmentation_b200.model import pack_conv3x3
    g = torch.Generator().manual_seed(seed)
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (3.0 * cin ** 0.5)).bfloat16().float()
    b = torch.randn(cout, generator=g) * 0.1
    x = torch.randn(B, H, W, cin, generator=g).bfloat16()
    want = F.conv2d(x.float().permute(0, 3, 1, 2), w, b, padding=d, dilation=d)
    if relu:
        want = F.relu(want)
    want = want.permute(0, 2, 3, 1)
    got = layers.conv3x3(x.cuda(), pack_conv3x3(w).cuda(), b.cuda(), dilation=d, relu=relu).float().cpu()
    err = (got - want).abs()
    assert bool((err <= 1e-2 * want.abs().clamp(min=1.0)).all()), f"max err {float(err.max()):.4g}"


@settings(max_examples=40, **COMMON)
@given(B=st.integers(1, 3), H=st.integers(1, 70), W=st.integers(1, 70), p=st.sampled_from([0.1, 0.4, 0.6, 0.9]),
       blur=st.sampled_from([0.0, 1.0, 2.0]), seed=st.integers(0, 10 ** 6))
def test_overlay_stencil_random_vs_opencv(cuda_device, B, H, W, p, blur, seed):
    import cv2
    import torch
    from scipy import ndimage as ndi
    from unet_dc_segmentation_b200 import overlay_stencil_device
    rs = np.random.RandomState(seed)
    f = rs.rand(B, H, W)
    if blur:
        f = np.stack([ndi.gaussian_filter(x, blur) for x in f])
        f = (f - f.min()) / max(float(np.ptp(f)), 1e-9)
    masks = ((f > 0.5) if blur else (f < p)).astype(np.uint8)
    got = overlay_stencil_device(torch.from_numpy(masks).cuda()).cpu().numpy()
    for b in range(B):
        img = np.zeros((H, W, 3), np.uint8)
        cnts, _ = cv2.findContours(masks[b], cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        cv2.drawContours(img, cnts, -1, (0, 255, 0), 2)
        np.testing.assert_array_equal(got[b], (img[..., 1] > 0).astype(np.uint8))


@settings(max_examples=25, **COMMON)
@given(H=st.integers(1, 80), W=st.integers(1, 80), nb=st.integers(1, 20), ks=st.sampled_from([3, 9, 21, 45]),
       seed=st.integers(0, 10 ** 6))
def test_density_maps_random_vs_oracle(cuda_device, H, W, nb, ks, seed):
    from scipy import ndimage as ndi
    from unet_dc_segmentation_b200 import density
    rs = np.random.RandomState(seed)
    base = ndi.gaussian_filter(rs.rand(H, W), 4.0)
    base = (base - base.min()) / max(float(np.ptp(base)), 1e-9)
    img = np.clip(base[..., None] * rs.randint(60, 256, 3)[None, None] + rs.randn(H, W, 3) * 6, 0, 255).astype(np.uint8)
    mask = (rs.rand(H, W) < 0.06).astype(np.uint8)
    roi_want, cy, cx = oracle.roi_mask(img)
    roi = density.generate_roi_mask(img)
    np.testing.assert_array_equal(roi, roi_want)
    np.testing.assert_array_equal(density.get_targets(mask, roi, nb, cy, cx), oracle.radial_density(mask, roi, nb, cy, cx))
    np.testing.assert_array_equal(density.density_maps(mask, roi, ks), oracle.spatial_density(mask, roi, ks))
