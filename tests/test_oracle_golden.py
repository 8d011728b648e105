"""CPU: the oracle (oracle/oracle.c + oracle/__init__.py) against the golden vectors produced by the
reference's own functions (tests/golden/make_golden.py) -- this is what pins the oracle."""
import numpy as np
import pytest

import oracle
from conftest import assert_table_equal, golden_cases, golden_table, load_golden, unpack_mask

RB = load_golden("rolling_ball.npz")
QT = load_golden("quantify.npz")


@pytest.mark.parametrize("case", golden_cases(RB))
def test_rolling_ball_matches_reference(case):
    img, radius, want = RB[f"{case}/in"], int(RB[f"{case}/radius"]), RB[f"{case}/out"]
    got = oracle.rolling_ball_correction_rgb(img, radius)
    np.testing.assert_array_equal(got, want)


def test_ellipse_rows_match_survey_probe():
    # row widths of cv2.getStructuringElement(MORPH_ELLIPSE, (50, 50)), 1995 taps
    j1, j2 = oracle.ellipse_rows(50)
    widths = (j2 - j1).tolist()
    assert sum(widths) == 1995
    assert widths[:8] == [1, 15, 21, 25, 29, 31, 33, 35] and widths[-3:] == [25, 21, 15]
    assert widths[21:30] == [50] * 9


@pytest.mark.parametrize("case", golden_cases(QT))
def test_quantify_matches_reference(case):
    mask = unpack_mask(QT, case)
    px = float(QT[f"{case}/px"])
    labels, cols = oracle.quantify_arrays(mask, int(QT[f"{case}/min_area"]), px if px else None)
    assert_table_equal(cols, golden_table(QT, f"{case}/t/"), case)
    assert labels.max() == len(cols["label"])


RZ = load_golden("resize.npz")


@pytest.mark.parametrize("case", golden_cases(RZ))
def test_resize_matches_reference_call_form(case):
    """The reference's resize calls (flag in the dst slot => bilinear) against the fixed-point restatement."""
    src, dsize, want = RZ[f"{case}/in"], tuple(int(v) for v in RZ[f"{case}/dsize"]), RZ[f"{case}/out"]
    np.testing.assert_array_equal(oracle.resize_linear_u8(src, dsize), want)


def test_quantify_empty_frame_contract():
    df = oracle.quantify(np.zeros((8, 8), np.uint8), 1, 3.45)
    assert df.empty and len(df.columns) == 0            # qdb:87-88


def test_label4_is_raster_ordered_and_compacting():
    img = np.array([[0, 7, 0, 3],
                    [5, 7, 0, 3],
                    [5, 0, 9, 9]], np.int32)
    lab, n = oracle.label4(img)
    assert n == 4
    assert lab.tolist() == [[0, 1, 0, 2], [3, 1, 0, 2], [3, 0, 4, 4]]


@pytest.mark.parametrize("tag", ["unetdc", "unet"])
def test_forward_restatement_matches_reference_module(tag):
    import torch
    from unet_dc_segmentation_b200.synth import calibrated_state_dict
    g = load_golden("forward.npz")
    dil = tuple(int(v) for v in g[f"{tag}/dilations"])
    torch.manual_seed(0)
    sd = calibrated_state_dict(seed=0, calib_size=64, n_calib=2, dilations=dil)
    x = torch.from_numpy(np.repeat(g[f"{tag}/images"][:, None], 3, 1).astype(np.float32) / 255.0)
    y = oracle.unetdc_forward(sd, x, dil).numpy()
    # fp32 on both sides; only thread-count / instruction-set differences in the CPU conv kernels remain
    np.testing.assert_allclose(y, g[f"{tag}/probs"], atol=2e-4, rtol=0)


def test_whole_path_matches_reference():
    import torch
    from unet_dc_segmentation_b200.synth import calibrated_state_dict
    g = load_golden("end_to_end.npz")
    sd = calibrated_state_dict(seed=0, calib_size=64, n_calib=2)
    imgs = [np.repeat(im[:, :, None], 3, 2) for im in g["images"]]
    probs, masks, tables = oracle.run_path(sd, imgs, radius=50, prob_thresh=0.3, min_area=1, px_per_um=3.45)
    pre = np.stack([oracle.rolling_ball_correction_rgb(im, 50) for im in imgs]).transpose(0, 3, 1, 2)
    np.testing.assert_array_equal(pre, g["pre"])
    np.testing.assert_allclose(probs, g["probs"][:, 0], atol=2e-4, rtol=0)
    for i in range(2):
        want_mask = np.unpackbits(g[f"mask{i}"])[: 96 * 96].reshape(96, 96)
        near = np.abs(g["probs"][i, 0] - 0.3) < 1e-3
        assert np.all((masks[i] == want_mask) | near)
        # tables are bit-exact given the same mask: feed the oracle the reference's mask
        _, cols = oracle.quantify_arrays(want_mask, 1, 3.45)
        assert_table_equal(cols, golden_table(g, f"t{i}/"), f"image {i}")


def test_reference_sample_outputs_known_answers():
    """The reference's shipped outputs/all_droplets.csv pins the column contract and the formulas
    (SURVEY.md 4): eq.diam = sqrt(4*area/pi), area_sqmicron = area/3.45^2, eq_diam_micron = diam/3.45."""
    g = load_golden("reference_outputs.npz")
    assert list(g["columns"]) == ["filename"] + oracle.COLUMNS + oracle.MICRON_COLUMNS
    area = g["area"].astype(np.int64)
    px = 3.45
    # evaluate with the oracle's own arithmetic: a 1 x area strip has exactly that area
    for a in np.unique(area)[:40]:
        _, cols = oracle.quantify_arrays(np.ones((1, int(a)), np.uint8), 1, px)
        sel = area == a
        assert cols["area"][0] == a
        # the CSV text carries 16 significant digits, so a parsed value may be a few ulp off the f64 it printed
        for col in ("equivalent_diameter", "area_sqmicron", "eq_diam_micron"):
            np.testing.assert_allclose(g[col][sel], cols[col][0], rtol=2e-15, atol=0, err_msg=f"{col} area={a}")
    # labels are 1..n per image, and the summary is count / sum(area)
    for fn, cnt, tot in zip(g["summary_filename"], g["summary_count"], g["summary_area"]):
        sel = g["filename"] == fn
        assert sel.sum() == cnt and area[sel].sum() == tot
        assert g["label"][sel].tolist() == list(range(1, int(cnt) + 1))


@pytest.mark.parametrize("shape,p", [((64, 64), 0.5), ((97, 131), 0.6), ((200, 300), 0.4), ((33, 7), 0.7)])
def test_labelling_matches_opencv_connected_components(shape, p):
    """scikit-image is not in this image, so `label` / `regionprops_table` semantics are also pinned against a real
    third-party implementation: cv2.connectedComponentsWithStats(connectivity=4) numbers components in raster order of
    their first pixel exactly as skimage.measure.label(connectivity=1) does, and reports pixel-count
    areas and mean-coordinate centroids in f64."""
    import cv2
    rs = np.random.RandomState(shape[0] * 7 + shape[1])
    mask = (rs.rand(*shape) < p).astype(np.uint8)
    n_cv, lab_cv, stats, cent = cv2.connectedComponentsWithStats(mask, connectivity=4, ltype=cv2.CV_32S)
    lab, cols = oracle.quantify_arrays(mask, 1, None)
    assert n_cv - 1 == len(cols["label"])
    np.testing.assert_array_equal(lab, lab_cv)
    np.testing.assert_array_equal(cols["area"], stats[1:, cv2.CC_STAT_AREA])
    np.testing.assert_array_equal(cols["centroid-0"], cent[1:, 1])        # cv2 centroids are (x, y)
    np.testing.assert_array_equal(cols["centroid-1"], cent[1:, 0])


# ----------------------------------------------------------------------------- overlay stencil (qdb:74-79)
from overlay_cases import _cv2_overlay_stencil, _overlay_cases  # noqa: E402


def test_overlay_stencil_matches_opencv_contours():
    """oracle.overlay_stencil (no border following) paints exactly the pixels of cv2.findContours(RETR_EXTERNAL,
    CHAIN_APPROX_SIMPLE) + cv2.drawContours(thickness 2): nested components, frame-touching ones, thin diagonals,
    checkerboards, random and blob-like masks."""
    for m in _overlay_cases():
        np.testing.assert_array_equal(oracle.overlay_stencil(m), _cv2_overlay_stencil(m), err_msg=f"mask {m.shape}")


def test_overlay_stencil_on_droplet_like_mask():
    from unet_dc_segmentation_b200.synth import synthetic_image
    img = synthetic_image(256, 3)
    m = (img > 60).astype(np.uint8)
    assert 0 < m.mean() < 0.5
    np.testing.assert_array_equal(oracle.overlay_stencil(m * 255), _cv2_overlay_stencil(m))


# ----------------------------------------------------------------------------- density maps (quantify_pipline.py, N4)
DENSITY_CASES = ["tissue_96x128", "border_128x128", "wide_64x200", "small_33x29", "flat_image", "no_droplets"]


@pytest.mark.parametrize("name", DENSITY_CASES)
def test_density_maps_match_the_reference_functions(name):
    """oracle.roi_mask / radial_density / spatial_density against generate_roi_mask, the cv2.moments centroid,
    get_targets and density_maps of the reference's quantify_pipline.py (tests/golden/make_golden_density.py)."""
    g = load_golden("density.npz")
    img, mask = g[f"{name}/img"], g[f"{name}/mask"]
    roi, cy, cx = oracle.roi_mask(img)
    np.testing.assert_array_equal(roi, g[f"{name}/roi"])
    assert [cy, cx] == list(g[f"{name}/centroid"])
    np.testing.assert_array_equal(oracle.radial_density(mask, roi, 10, cy, cx), g[f"{name}/radial"])
    np.testing.assert_array_equal(oracle.spatial_density(mask, roi), g[f"{name}/spatial"])          # bit for bit


def test_roi_mask_steps_match_opencv():
    """Each OpenCV call of generate_roi_mask on its own: gray + 15x15 Gaussian (fixed point), the Otsu threshold, and
    the close / open pair, on random sizes down to 1 pixel."""
    import cv2
    from scipy import ndimage as ndi
    rs = np.random.RandomState(3)
    kernel = np.ones((15, 15), np.uint8)
    for it in range(40):
        H, W = rs.randint(1, 90, 2) if it % 4 else rs.randint(90, 260, 2)
        base = ndi.gaussian_filter(rs.rand(H, W), rs.choice([2, 5, 12]))
        base = (base - base.min()) / max(float(np.ptp(base)), 1e-9)
        img = np.clip(base[..., None] * rs.randint(100, 256, 3)[None, None] + rs.randn(H, W, 3) * 8, 0, 255).astype(np.uint8)
        blurred = cv2.GaussianBlur(cv2.cvtColor(img, cv2.COLOR_RGB2GRAY), (15, 15), 0)
        np.testing.assert_array_equal(oracle.roi_blurred_gray(img), blurred, err_msg=f"blur {H}x{W}")
        t, m = cv2.threshold(blurred, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
        assert oracle.otsu_threshold_u8(blurred) == int(t)
        m = cv2.morphologyEx(cv2.morphologyEx(m, cv2.MORPH_CLOSE, kernel), cv2.MORPH_OPEN, kernel)
        roi, cy, cx = oracle.roi_mask(img)
        np.testing.assert_array_equal(roi, (m > 0).astype(np.uint8), err_msg=f"roi {H}x{W}")
        M = cv2.moments(roi)
        assert (cy, cx) == ((int(M["m01"] / M["m00"]), int(M["m10"] / M["m00"])) if M["m00"] else (H // 2, W // 2))


def test_gaussian_filter_restatement_matches_scipy_bitwise():
    from scipy.ndimage import gaussian_filter
    rs = np.random.RandomState(4)
    for shape in [(64, 80), (7, 9), (200, 33), (1, 40), (29, 1)]:
        for sigma in (21 / 6, 1.0, 7.5):
            a = (rs.rand(*shape) < 0.3).astype(np.float32)
            np.testing.assert_array_equal(oracle.gaussian_filter_f32(a, sigma), gaussian_filter(a, sigma=sigma))
