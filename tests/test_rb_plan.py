"""CPU: the chord plan csrc/morph.cu builds on the host (dumped through dc_debug_rolling_ball_plan) run through a
numpy model of the kernel (tests/rb_model.py) against OpenCV's own erode / dilate with the reference's structuring
element (utils/data_loader.py:17-19).  No GPU needed: this pins the plan builder; tests/test_gpu_rolling_ball.py pins
the kernel."""
import numpy as np
import pytest

from rb_model import load_plan, morph_pass


@pytest.mark.parametrize("radius,th", [(1, 128), (2, 128), (3, 64), (4, 128), (5, 32), (7, 128), (8, 128), (15, 128),
                                       (20, 64), (33, 128), (50, 128), (50, 64), (51, 128), (64, 128), (77, 128),
                                       (100, 128), (128, 128), (150, 64), (174, 32)])
def test_chord_plan_matches_cv2(radius, th):
    import cv2
    rs = np.random.RandomState(radius)
    img = rs.randint(0, 256, (150, 170)).astype(np.uint8)
    img[40:60, 50:90] = 255
    img[100:120, 10:30] = 0
    plan = load_plan(radius, th)
    se = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (radius, radius))
    np.testing.assert_array_equal(morph_pass(img, plan, False), cv2.erode(img, se))
    np.testing.assert_array_equal(morph_pass(img, plan, True), cv2.dilate(img, se))


def test_plan_shape_at_radius_50():
    p = load_plan(50, 128)
    assert p["nchords"] == 16 and p["ntables"] == 3 and p["pitch"] % 2 == 1
    # every element row is in exactly one chord; a chord's list is padded to an even length by repeating its last row
    assert sorted(set(p["rowoff"])) == [i * p["HP"] for i in range(50)]
    assert all((c["rows"][1] - c["rows"][0]) % 2 == 0 for c in p["chords"])


def test_max_radius_is_exposed():
    from unet_dc_segmentation_b200 import _lib
    from unet_dc_segmentation_b200.morphology import max_radius
    assert max_radius() == _lib.load().dc_rolling_ball_max_radius() >= 100
