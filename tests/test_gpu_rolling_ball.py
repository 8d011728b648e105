"""GPU: dc_rolling_ball (csrc/morph.cu) against the reference's cv2 output (golden) and the oracle.
u8 work: bit-exact."""
import numpy as np
import pytest

import oracle
from conftest import golden_cases, load_golden

pytestmark = pytest.mark.gpu
RB = load_golden("rolling_ball.npz")


@pytest.mark.parametrize("case", golden_cases(RB))
def test_golden(cuda_device, case):
    from unet_dc_segmentation_b200 import rolling_ball_correction_rgb
    img, radius, want = RB[f"{case}/in"], int(RB[f"{case}/radius"]), RB[f"{case}/out"]
    got = rolling_ball_correction_rgb(img, radius)
    assert got.shape == want.shape and got.dtype == np.uint8
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("shape,radius", [((3, 100, 130), 50), ((2, 256, 256), 50), ((1, 65, 33), 9),
                                           ((2, 40, 300), 30), ((1, 17, 19), 100), ((1, 300, 64), 64)])
def test_planar_batch_vs_oracle(cuda_device, shape, radius):
    import torch
    from unet_dc_segmentation_b200 import rolling_ball_device
    from unet_dc_segmentation_b200.synth import synthetic_image
    rs = np.random.RandomState(shape[1] * 7 + radius)
    B, H, W = shape
    imgs = np.stack([(synthetic_image(max(H, W), 20 + b)[:H, :W] if b % 2 == 0
                      else rs.randint(0, 256, (H, W)).astype(np.uint8)) for b in range(B)])
    got = rolling_ball_device(torch.from_numpy(imgs).cuda(), radius).cpu().numpy()
    for b in range(B):
        want = oracle.rolling_ball_correction_rgb(imgs[b][:, :, None], radius)[:, :, 0]
        np.testing.assert_array_equal(got[b], want, err_msg=f"image {b}")


def test_rgb_interleaved_channels_independent(cuda_device):
    import torch
    from unet_dc_segmentation_b200 import rolling_ball_device
    rs = np.random.RandomState(3)
    img = rs.randint(0, 256, (2, 90, 70, 3)).astype(np.uint8)
    got = rolling_ball_device(torch.from_numpy(img).cuda(), 21).cpu().numpy()
    for b in range(2):
        np.testing.assert_array_equal(got[b], oracle.rolling_ball_correction_rgb(img[b], 21))


def test_large_frame_radius50(cuda_device):
    """Radius 50 on a 1024^2 frame (BASELINE config 2/4 size) against the oracle on the full frame
    (the oracle's naive C loops take ~5 s here; 2048^2 is covered by the stretch/idempotence properties below)."""
    import torch
    from unet_dc_segmentation_b200 import rolling_ball_device
    from unet_dc_segmentation_b200.synth import synthetic_image
    img = synthetic_image(1024, 0)
    got = rolling_ball_device(torch.from_numpy(img[None]).cuda(), 50)[0].cpu().numpy()
    want = oracle.rolling_ball_correction_rgb(img[:, :, None], 50)[:, :, 0]
    np.testing.assert_array_equal(got, want)
    assert got.min() == 0 and got.max() == 255                               # min-max stretch property


def test_config3_size_2048_properties(cuda_device):
    """BASELINE config 3 size: properties that do not need the (slow) oracle at 2048^2 -- every 512^2
    interior crop far enough (>= 2*radius) from the crop border must match the oracle run on the crop
    before the stretch, which we check through the stretch-invariant ordering, and the output spans 0..255."""
    import torch
    from unet_dc_segmentation_b200 import rolling_ball_device
    from unet_dc_segmentation_b200.synth import synthetic_image
    img = synthetic_image(2048, 1)
    got = rolling_ball_device(torch.from_numpy(img[None]).cuda(), 50)[0].cpu().numpy()
    assert got.min() == 0 and got.max() == 255
    # the stretch is a monotone per-plane LUT: on an interior window the oracle's un-stretched
    # correction (computed on a crop with a 2*radius apron) must map to `got` through ONE monotone function
    y0, x0, n, ap = 700, 900, 256, 100
    crop = img[y0 - ap:y0 + n + ap, x0 - ap:x0 + n + ap]
    j1, j2 = oracle.ellipse_rows(50)
    import cv2
    k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (50, 50))
    assert np.array_equal(k.sum(1), j2 - j1)
    corr = cv2.subtract(crop, cv2.morphologyEx(crop, cv2.MORPH_OPEN, k))[ap:ap + n, ap:ap + n]
    win = got[y0:y0 + n, x0:x0 + n]
    lut = {}
    for c, g in zip(corr.ravel().tolist(), win.ravel().tolist()):
        assert lut.setdefault(c, g) == g
    keys = sorted(lut)
    assert all(lut[a] <= lut[b] for a, b in zip(keys, keys[1:]))


@pytest.mark.parametrize("shape,radius", [((1, 200, 260), 128), ((2, 150, 90), 150), ((1, 300, 300), None), ((1, 64, 64), 120)])
def test_large_radii_vs_cv2(cuda_device, shape, radius):
    """Radii beyond 100 (the reference CLI takes any --background_radius): shorter tiles, more chord levels;
    against OpenCV itself with the reference's four calls (utils/data_loader.py:17-21)."""
    import cv2
    import torch
    from unet_dc_segmentation_b200 import rolling_ball_device
    from unet_dc_segmentation_b200.morphology import max_radius
    radius = radius or max_radius()              # None: the largest element the kernel takes
    rs = np.random.RandomState(radius)
    B, H, W = shape
    imgs = rs.randint(0, 256, (B, H, W)).astype(np.uint8)
    imgs[:, H // 3:H // 2, W // 4:W // 2] //= 4
    got = rolling_ball_device(torch.from_numpy(imgs).cuda(), radius).cpu().numpy()
    k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (radius, radius))
    for b in range(B):
        want = cv2.normalize(cv2.subtract(imgs[b], cv2.morphologyEx(imgs[b], cv2.MORPH_OPEN, k)), None, 0, 255,
                             cv2.NORM_MINMAX)
        np.testing.assert_array_equal(got[b], want, err_msg=f"image {b}")


def test_radius_limit_is_reported(cuda_device):
    import torch
    from unet_dc_segmentation_b200 import rolling_ball_device
    from unet_dc_segmentation_b200.morphology import max_radius
    with pytest.raises(ValueError, match="radius must be in"):
        rolling_ball_device(torch.zeros((1, 32, 32), dtype=torch.uint8).cuda(), max_radius() + 1)
