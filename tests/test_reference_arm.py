"""CPU: oracle/_ref (the reference's own modules, byte-compiled where they lie) against the oracle port.  Whenever the
compiled reference is present -- in the build container always -- the port's whole path must reproduce it bit for
bit: probabilities, masks and every table column.  This is the strongest pin of the restatement: the unmodified
reference code (over real cv2 and real torch; skimage.measure through the scipy-backed shim of oracle/shims.py)."""
import numpy as np
import pytest

import oracle
from oracle import ref


@pytest.fixture(scope="module")
def reference():
    if not ref.available() and not ref.build_ref():
        pytest.skip("no /root/reference and no prebuilt oracle/_ref")
    return ref.load()


def test_reference_modules_are_the_reference(reference):
    assert reference.qdb.IMG_SIZE == 512 and reference.qdb.__name__ == "quantify_droplets_batch"
    assert reference.UNetDC.__module__ == "models.model_2"
    assert list(reference.UNetDC(3, 1).state_dict().keys())[:2] == ["enc1.0.weight", "enc1.0.bias"]


@pytest.mark.parametrize("min_area,px", [(1, 3.45), (4, None)])
def test_port_reproduces_the_reference_path(reference, min_area, px):
    from unet_dc_segmentation_b200.synth import calibrated_state_dict, synthetic_image
    sd = calibrated_state_dict(seed=0, calib_size=64, n_calib=2)
    imgs = [np.repeat(synthetic_image(96, 30 + i)[:, :, None], 3, 2) for i in range(2)]
    p_ref, m_ref, t_ref = ref.run_path(sd, imgs, 50, 0.3, min_area, px)
    p_port, m_port, t_port = oracle.run_path(sd, imgs, 50, 0.3, min_area, px, use_cv2=False)     # the C rolling ball
    np.testing.assert_array_equal(p_ref, p_port)
    np.testing.assert_array_equal(m_ref, m_port)
    for a, b in zip(t_ref, t_port):
        assert list(a.columns) == list(b.columns) and len(a) == len(b) > 0
        for c in a.columns:
            np.testing.assert_array_equal(a[c].to_numpy(), b[c].to_numpy(), err_msg=c)


def test_reference_quantify_empty_and_filtered(reference):
    import pandas as pd
    assert reference.quantify(np.zeros((16, 16), np.uint8), 1, None).equals(pd.DataFrame())
    cb = (np.add.outer(np.arange(16), np.arange(16)) % 2).astype(np.uint8)
    assert reference.quantify(cb, 2, 3.0).empty and oracle.quantify(cb, 2, 3.0).empty


@pytest.mark.parametrize("cin,cout", [(3, 1), (1, 1), (5, 2)])
def test_module_contract_matches_the_reference_for_any_channel_count(reference, cin, cout):
    """Same constructor, same state_dict keys / shapes / dtypes as the reference class it replaces (model_2.py:6-32)."""
    from unet_dc_segmentation_b200.model import UNetDC
    ours = UNetDC(cin, cout).state_dict()
    theirs = reference.UNetDC(cin, cout).state_dict()
    assert list(ours.keys()) == list(theirs.keys())
    for k in ours:
        assert ours[k].shape == theirs[k].shape and ours[k].dtype == theirs[k].dtype, k
