"""GPU: dc_resize_linear_u8 (csrc/resize.cu) -- the two cv2.resize calls of reference qdb:44 and qdb:57 --
against the golden outputs of the reference's exact call forms, the oracle, and cv2 itself.  u8 work: bit-exact."""
import numpy as np
import pytest

import oracle
from conftest import golden_cases, load_golden

pytestmark = pytest.mark.gpu
RZ = load_golden("resize.npz")


@pytest.mark.parametrize("case", golden_cases(RZ))
def test_golden_reference_call_forms(cuda_device, case):
    import torch
    from unet_dc_segmentation_b200 import resize_linear_u8_device
    src, dsize, want = RZ[f"{case}/in"], tuple(int(v) for v in RZ[f"{case}/dsize"]), RZ[f"{case}/out"]
    got = resize_linear_u8_device(torch.from_numpy(src[None]).cuda(), dsize)[0].cpu().numpy()
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("sh,sw,dh,dw,cn", [(256, 256, 512, 512, 3), (512, 512, 256, 256, 1), (276, 408, 512, 512, 3),
                                              (512, 512, 276, 408, 1), (1024, 1024, 512, 512, 3), (512, 512, 2048, 2048, 1),
                                              (33, 7, 5, 100, 3), (1, 1, 16, 16, 1), (64, 64, 64, 64, 3)])
def test_batch_vs_oracle_and_cv2(cuda_device, sh, sw, dh, dw, cn):
    import cv2
    import torch
    from unet_dc_segmentation_b200 import resize_linear_u8_device
    rs = np.random.RandomState(sh * 31 + dw)
    shape = (3, sh, sw, 3) if cn == 3 else (3, sh, sw)
    src = rs.randint(0, 256, shape).astype(np.uint8) if cn == 3 else (rs.rand(*shape) < 0.4).astype(np.uint8)
    got = resize_linear_u8_device(torch.from_numpy(src).cuda(), (dw, dh)).cpu().numpy()
    for b in range(3):
        np.testing.assert_array_equal(got[b], oracle.resize_linear_u8(src[b], (dw, dh)), err_msg=f"oracle, image {b}")
        np.testing.assert_array_equal(got[b], cv2.resize(src[b], (dw, dh)).reshape(got[b].shape), err_msg=f"cv2, image {b}")


def test_as_shipped_flow_256_to_512_and_back(cuda_device):
    """BASELINE config 1 as shipped: 256^2 frames, IMG_SIZE = 512 (qdb:30,44,57).  The device flow with img_size=512
    must equal the oracle's: rolling ball -> bilinear up-size -> network -> threshold -> bilinear down-size of the
    0/1 mask -> quantify.  u8 stages bit-exact; the table bit-exact given the kernel's own 512^2 mask."""
    import torch
    from unet_dc_segmentation_b200 import DropletPipeline, UNetDC
    from unet_dc_segmentation_b200.synth import calibrated_state_dict, synthetic_image
    sd = calibrated_state_dict(seed=0, calib_size=64, n_calib=2)
    m = UNetDC(3, 1)
    m.load_state_dict(sd)
    m = m.to(cuda_device).eval()
    imgs = np.stack([synthetic_image(256, 70 + i) for i in range(2)])
    pipe = DropletPipeline(m, 50, 0.3, 1, 3.45, img_size=512)
    res = pipe.run_device(torch.from_numpy(imgs).to(cuda_device), want_labels=True)
    masks = res.masks.cpu().numpy()
    assert masks.shape == (2, 256, 256)
    tables = res.tables.to_host()
    # reproduce on the CPU with the kernel's own 512^2 mask (so bf16 noise in the network cannot enter)
    pipe512 = DropletPipeline(m, None, 0.3, 1, 3.45)
    for i in range(2):
        pre = oracle.rolling_ball_correction_rgb(imgs[i][:, :, None], 50)[:, :, 0]
        up = oracle.resize_linear_u8(pre, (512, 512))
        mask512 = pipe512.run_device(torch.from_numpy(up[None]).to(cuda_device)).masks[0].cpu().numpy()
        want_mask = oracle.resize_linear_u8(mask512, (256, 256))
        np.testing.assert_array_equal(masks[i], want_mask, err_msg=f"image {i}")
        labels, cols = oracle.quantify_arrays(want_mask, 1, 3.45)
        np.testing.assert_array_equal(res.tables.labels[i].cpu().numpy(), labels)
        for c, v in cols.items():
            np.testing.assert_array_equal(np.asarray(tables[i][c]), v, err_msg=f"image {i} column {c}")


def test_cli_preprocess_matches_oracle(cuda_device, tmp_path):
    from PIL import Image
    from unet_dc_segmentation_b200 import cli
    from unet_dc_segmentation_b200.synth import synthetic_image
    g = synthetic_image(128, 5)[:80, :112]
    Image.fromarray(g).save(tmp_path / "f.png")
    t, (oh, ow) = cli.preprocess(tmp_path / "f.png", 50, 96)
    assert (oh, ow) == (80, 112) and tuple(t.shape) == (3, 96, 96) and t.is_cuda
    rgb = np.repeat(g[:, :, None], 3, 2)
    want = oracle.resize_linear_u8(oracle.rolling_ball_correction_rgb(rgb, 50), (96, 96)).astype(np.float32) / 255.0
    np.testing.assert_array_equal(t.cpu().numpy(), want.transpose(2, 0, 1))
