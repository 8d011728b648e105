"""GPU (needs >= 2 devices, skipped otherwise): one process driving two GPUs -- per-device kernel attributes,
model handles and workspaces -- gives identical results on both."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_two_devices_in_one_process():
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from unet_dc_segmentation_b200 import DropletPipeline, UNetDC
    from unet_dc_segmentation_b200.synth import calibrated_state_dict, synthetic_image
    sd = calibrated_state_dict(seed=0, calib_size=64, n_calib=2)
    imgs = np.stack([synthetic_image(128, 40 + i) for i in range(3)])
    outs = []
    for d in (1, 0, 1):
        dev = torch.device("cuda", d)
        m = UNetDC(3, 1)
        m.load_state_dict(sd)
        m = m.to(dev).eval()
        res = DropletPipeline(m, 50, 0.3, 1, 3.45).run_device(torch.from_numpy(imgs).to(dev), return_prob=True)
        outs.append((res.probs.cpu(), res.masks.cpu(), res.tables.to_host()))
    for o in outs[1:]:
        assert torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1])
        for a, b in zip(o[2], outs[0][2]):
            for c in a:
                np.testing.assert_array_equal(a[c], b[c])
