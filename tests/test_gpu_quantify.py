"""GPU: dc_label_stats (csrc/ccl.cu) against the oracle and the reference-generated golden tables.
Integer work: labels, counts, areas and the f64 columns must be bit-exact."""
import numpy as np
import pytest

import oracle
from conftest import assert_table_equal, golden_cases, golden_table, load_golden, unpack_mask

pytestmark = pytest.mark.gpu
QT = load_golden("quantify.npz")


def _gpu_quantify(masks, min_area, px, want_labels=True):
    import torch
    from unet_dc_segmentation_b200 import quantify_arrays
    t = torch.from_numpy(np.ascontiguousarray(masks)).cuda()
    tables, labels = quantify_arrays(t, min_area, px, want_labels=want_labels)
    return tables, (labels.cpu().numpy() if labels is not None else None)


@pytest.mark.parametrize("case", golden_cases(QT))
def test_golden_tables(cuda_device, case):
    mask = unpack_mask(QT, case)
    px = float(QT[f"{case}/px"]) or None
    min_area = int(QT[f"{case}/min_area"])
    tables, labels = _gpu_quantify(mask[None], min_area, px)
    assert_table_equal(tables[0], golden_table(QT, f"{case}/t/"), case)
    want_labels, _ = oracle.quantify_arrays(mask, min_area, px)
    np.testing.assert_array_equal(labels[0], want_labels)


@pytest.mark.parametrize("shape,p,min_area", [((33, 65), 0.5, 1), ((128, 128), 0.55, 1), ((100, 260), 0.6, 4),
                                               ((257, 31), 0.45, 2), ((1, 70), 0.5, 1), ((70, 1), 0.5, 1)])
def test_random_masks_batch(cuda_device, shape, p, min_area):
    rs = np.random.RandomState(hash(shape) & 0xffff)
    masks = (rs.rand(5, *shape) < p).astype(np.uint8)
    masks[3] = 0
    masks[4] = 1
    tables, labels = _gpu_quantify(masks, min_area, 3.45)
    for b in range(5):
        want_l, want = oracle.quantify_arrays(masks[b], min_area, 3.45)
        np.testing.assert_array_equal(labels[b], want_l, err_msg=f"image {b}")
        want["n"] = len(want["label"])
        assert_table_equal(tables[b], want, f"image {b}")


def test_quantify_dataframe_contract(cuda_device):
    from unet_dc_segmentation_b200 import quantify
    from unet_dc_segmentation_b200.synth import synthetic_mask
    m = synthetic_mask(200, 150, seed=5)
    df = quantify(m, 1, 3.45)
    want = oracle.quantify(m, 1, 3.45)
    assert list(df.columns) == list(want.columns)
    for c in df.columns:
        np.testing.assert_array_equal(df[c].to_numpy(), want[c].to_numpy())
    assert list(quantify(m, 1, None).columns) == oracle.COLUMNS
    empty = quantify(np.zeros((16, 16), np.uint8), 1, 3.45)
    assert empty.empty and len(empty.columns) == 0                         # qdb:87-88
    assert quantify(m, 10 ** 9, None).empty                                  # everything filtered


def test_config3_mask_2048(cuda_device):
    """BASELINE config 3 size: 2048^2 synthetic disc mask, ~10k droplets, vs the oracle (bit-exact),
    plus size-independent properties."""
    from unet_dc_segmentation_b200.synth import synthetic_mask
    m = synthetic_mask(2048, 14000, seed=0, rmin=2, rmax=5)              # 10,054 droplets
    tables, labels = _gpu_quantify(m[None], 1, 3.45)
    t = tables[0]
    want_l, want = oracle.quantify_arrays(m, 1, 3.45)
    assert len(want["label"]) == 10054
    np.testing.assert_array_equal(labels[0], want_l)
    want["n"] = len(want["label"])
    assert_table_equal(t, want, "2048^2")
    assert int(t["area"].sum()) == int(m.sum())                              # checksum of checksums
    first = np.full(len(t["label"]) + 1, m.size, np.int64)                   # raster order of first pixels
    flat = labels[0].ravel()
    idx = np.flatnonzero(flat)
    np.minimum.at(first, flat[idx], idx)
    assert np.all(np.diff(first[1:]) > 0)
    # idempotence: labelling the label image's support gives the same table
    tables2, _ = _gpu_quantify((labels[0] > 0).astype(np.uint8)[None], 1, 3.45, want_labels=False)
    for c in t:
        np.testing.assert_array_equal(tables2[0][c], t[c])


def test_capacity_overflow_reruns(cuda_device):
    import torch
    from unet_dc_segmentation_b200 import label_stats_device, quantify_arrays
    cb = (np.add.outer(np.arange(64), np.arange(64)) % 2).astype(np.uint8)   # 2048 single-pixel droplets
    t = label_stats_device(torch.from_numpy(cb[None]).cuda(), capacity=100)
    assert int(t.counts[0]) == 2048                                          # exact even though the table is full
    with pytest.raises(Exception):
        t.to_host()
    tables, _ = quantify_arrays(torch.from_numpy(cb[None]).cuda(), capacity=100)
    assert len(tables[0]["label"]) == 2048 and np.all(tables[0]["area"] == 1)
