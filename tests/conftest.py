import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))
GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with `-m gpu` under gpurun")


def load_golden(name: str) -> dict:
    with np.load(GOLDEN / name, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def golden_cases(d: dict) -> list[str]:
    return sorted({k.split("/")[0] for k in d})


def unpack_mask(d: dict, case: str) -> np.ndarray:
    h, w = (int(v) for v in d[f"{case}/shape"])
    return np.unpackbits(d[f"{case}/mask"])[: h * w].reshape(h, w).astype(np.uint8)


def golden_table(d: dict, prefix: str) -> dict:
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


TABLE_FLOAT_COLS = ["equivalent_diameter", "centroid-0", "centroid-1", "area_sqmicron", "eq_diam_micron"]


def assert_table_equal(got: dict, want: dict, what: str = ""):
    """Bit-exact comparison of a droplet table (dict of numpy columns) with a golden one."""
    n = int(want["n"])
    assert len(got["label"]) == n, f"{what}: droplet count {len(got['label'])} != {n}"
    if n == 0:
        return
    np.testing.assert_array_equal(np.asarray(got["label"], np.int64), want["label"].astype(np.int64), err_msg=what)
    np.testing.assert_array_equal(np.asarray(got["area"], np.int64), want["area"].astype(np.int64), err_msg=what)
    for c in TABLE_FLOAT_COLS:
        if c in want:
            assert c in got, f"{what}: column {c} missing"
            a = np.asarray(got[c], np.float64)
            np.testing.assert_array_equal(a.view(np.int64), want[c].astype(np.float64).view(np.int64),
                                          err_msg=f"{what}: column {c} not bit-exact")
        else:
            assert c not in got or c in ("equivalent_diameter", "centroid-0", "centroid-1"), f"{what}: extra column {c}"


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)
