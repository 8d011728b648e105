"""CPU: the run-based labelling scheme of csrc/ccl.cu (bit-packed 64-px words, 64 x 128 tiles, union-find over run
starts, packed per-tile moments, popcount-prefix numbering), modelled step by step in tests/ccl_model.py, against the
oracle.  This pins the DESIGN of the GPU algorithm where no GPU is available; tests/test_gpu_quantify.py pins the
kernels themselves."""
import numpy as np
import pytest

import oracle
from ccl_model import label_stats, nz4


def _cases():
    from unet_dc_segmentation_b200.synth import synthetic_mask
    rs = np.random.RandomState(0)
    spiral = np.zeros((200, 200), np.uint8)
    for k in range(0, 100, 4):
        spiral[k, k:200 - k] = 1
        spiral[k:200 - k, 199 - k] = 1
        spiral[199 - k, k:200 - k] = 1
        spiral[k + 4:200 - k, k] = 1
    return {
        "discs_300": synthetic_mask(300, 400, seed=1),
        "discs_300x200": synthetic_mask(300, 400, seed=2)[:, :200],
        "noise_260x130": (rs.rand(260, 130) < 0.5).astype(np.uint8),
        "dense_140x70": (rs.rand(140, 70) < 0.7).astype(np.uint8),
        "checkerboard_130x129": (np.add.outer(np.arange(130), np.arange(129)) % 2).astype(np.uint8),
        "ones_257x193": np.ones((257, 193), np.uint8),
        "zeros_10x10": np.zeros((10, 10), np.uint8),
        "spiral_200": spiral,
    }


def test_byte_to_bit_packing_trick():
    rs = np.random.RandomState(1)
    for _ in range(4000):
        v = int(rs.randint(0, 2 ** 32, dtype=np.uint64))
        for k in range(4):
            if rs.rand() < 0.5:
                v &= ~(0xFF << (8 * k))
        want = sum((((v >> (8 * k)) & 0xFF) != 0) << k for k in range(4))
        assert nz4(v) == want


@pytest.mark.parametrize("name", list(_cases()))
@pytest.mark.parametrize("min_area,word_bits", [(1, 32), (5, 32), (1, 64)])
def test_run_based_scheme_matches_oracle(name, min_area, word_bits):
    """word_bits 32: the droplet path (dc_label_stats); 64: the background labelling of dc_overlay_stencil."""
    m = _cases()[name]
    want_l, cols = oracle.quantify_arrays(m, min_area, None)
    labels, area, s0, s1 = label_stats(m, min_area, word_bits=word_bits)
    np.testing.assert_array_equal(labels, want_l)
    np.testing.assert_array_equal(area, cols["area"])
    if len(area):
        np.testing.assert_array_equal(s0 / area, cols["centroid-0"])
        np.testing.assert_array_equal(s1 / area, cols["centroid-1"])
