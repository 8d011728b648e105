"""dc_overlay_stencil (GPU) against the oracle and against OpenCV's own findContours + drawContours
(reference quantify_droplets_batch.py:74-79): bit-exact, ragged sizes, batches, full-size frames."""
import numpy as np
import pytest

from overlay_cases import _cv2_overlay_stencil, _overlay_cases

pytestmark = pytest.mark.gpu


def _gpu(masks):
    import torch
    from unet_dc_segmentation_b200 import overlay_stencil_device
    m = torch.from_numpy(np.ascontiguousarray(masks)).cuda()
    return overlay_stencil_device(m).cpu().numpy()


def test_overlay_cases_match_oracle_and_opencv(cuda_device):
    import oracle
    for m in _overlay_cases():
        got = _gpu(m[None])[0]
        np.testing.assert_array_equal(got, oracle.overlay_stencil(m), err_msg=f"oracle, mask {m.shape}")
        np.testing.assert_array_equal(got, _cv2_overlay_stencil(m), err_msg=f"cv2, mask {m.shape}")


@pytest.mark.parametrize("shape", [(3, 31, 33), (2, 64, 64), (5, 97, 131), (1, 200, 300), (4, 32, 96)])
def test_overlay_batches_cross_tile_borders(cuda_device, shape):
    """Components, holes and diagonal steps straddling the 32-pixel tiles of both the labelling and the stencil kernel."""
    from scipy import ndimage as ndi
    rs = np.random.RandomState(shape[1] * 3 + shape[2])
    B, H, W = shape
    masks = np.stack([(ndi.gaussian_filter(rs.rand(H, W), 1.0 + b % 3) > 0.5).astype(np.uint8) * (255 if b % 2 else 1)
                      for b in range(B)])
    got = _gpu(masks)
    for b in range(B):
        np.testing.assert_array_equal(got[b], _cv2_overlay_stencil(masks[b]), err_msg=f"image {b}")


def test_overlay_long_serpentine_background(cuda_device):
    """Outer background reaching deep inside through a one-pixel corridor (long union-find chains), next to a sealed hole."""
    H = W = 129
    m = np.ones((H, W), np.uint8)
    for k, y in enumerate(range(1, H - 1, 2)):                # serpentine corridor open to the frame at the top-left
        m[y, 1:W - 1] = 0
        m[y + 1, (W - 2) if k % 2 == 0 else 1] = 0
    m[0, 1] = 0
    sealed = m.copy()
    sealed[0, 1] = 1                                          # same corridor, now a hole: nothing inside is outlined
    for case in (m, sealed):
        np.testing.assert_array_equal(_gpu(case[None])[0], _cv2_overlay_stencil(case))


def test_overlay_full_size_network_like_masks(cuda_device):
    from unet_dc_segmentation_b200.synth import synthetic_image
    masks = np.stack([(synthetic_image(1024, i) > 45).astype(np.uint8) for i in range(4)])
    got = _gpu(masks)
    assert 0.01 < got.mean() < 0.6
    for b in range(4):
        np.testing.assert_array_equal(got[b], _cv2_overlay_stencil(masks[b]), err_msg=f"image {b}")


def test_overlay_argument_checks(cuda_device):
    import torch
    from unet_dc_segmentation_b200 import overlay_stencil_device
    with pytest.raises(TypeError):
        overlay_stencil_device(torch.zeros((1, 8, 8), dtype=torch.float32, device="cuda"))
    with pytest.raises(RuntimeError):
        overlay_stencil_device(torch.zeros((1, 8, 8), dtype=torch.uint8))
    with pytest.raises(ValueError):
        overlay_stencil_device(torch.zeros((1, 8, 8), dtype=torch.uint8, device="cuda"),
                               out=torch.zeros((1, 8, 9), dtype=torch.uint8, device="cuda"))
