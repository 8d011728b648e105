"""sys.modules shims that let the UNMODIFIED reference modules import in this image -- TEST INFRASTRUCTURE ONLY.

Three packages the reference imports are absent here (no network to install them):
``matplotlib`` and ``albumentations`` (never called on the quantify_droplets_batch.py hot path) and
``skimage.measure``, whose ``label`` / ``regionprops_table`` are backed by scipy.ndimage: label-for-label identical to
scikit-image's 4-connectivity raster-order numbering; area = pixel count, centroid = mean coordinate,
equivalent_diameter = sqrt(4*area/pi) as published by scikit-image.  Used by tests/golden/make_golden.py (golden
vectors) and by oracle/ref.py (the reference arm of bench.py)."""
from __future__ import annotations

import sys
import types

import numpy as np


def _label(img, connectivity=1, **_):
    """skimage.measure.label(img, connectivity=1): equal-valued 4-neighbours connect, background 0, components
    numbered in raster order of their first pixel."""
    import scipy.ndimage as ndi
    assert connectivity == 1
    cross = ndi.generate_binary_structure(2, 1)
    img = np.asarray(img)
    comp, total = ndi.label(img != 0, structure=cross)
    comp = comp.astype(np.int64)
    if total:
        # a component of the non-zero pixels is a component of the image iff it holds a single value -- always the
        # case for a binary mask and for a label image with some labels zeroed (both calls of qdb:82,86)
        lo = ndi.minimum(img, comp, np.arange(1, total + 1))
        hi = ndi.maximum(img, comp, np.arange(1, total + 1))
        if not np.array_equal(lo, hi):                     # general multi-valued image: value by value
            comp = np.zeros(img.shape, np.int64)
            total = 0
            for v in np.unique(img):
                if v == 0:
                    continue
                lab, n = ndi.label(img == v, structure=cross)
                comp[lab > 0] = lab[lab > 0] + total
                total += n
    flat = comp.ravel()
    first = np.full(total + 1, flat.size, np.int64)
    idx = np.flatnonzero(flat)
    np.minimum.at(first, flat[idx], idx)
    order = np.argsort(first[1:], kind="stable")
    remap = np.zeros(total + 1, np.int64)
    remap[order + 1] = np.arange(1, total + 1)
    return remap[comp]


def _regionprops_table(lbl, properties=()):
    import scipy.ndimage as ndi
    lbl = np.asarray(lbl)
    ids = np.unique(lbl)
    ids = ids[ids != 0]
    res = {}
    area = ndi.sum_labels(np.ones(lbl.shape, np.float64), lbl, ids).astype(np.int64)
    com = np.array(ndi.center_of_mass(np.ones(lbl.shape, np.float64), lbl, ids), dtype=np.float64).reshape(-1, 2)
    for p in properties:
        if p == "label":
            res["label"] = ids.astype(np.int64)
        elif p == "area":
            res["area"] = area
        elif p == "equivalent_diameter":
            res["equivalent_diameter"] = np.sqrt(4.0 * area / np.pi)
        elif p == "centroid":
            res["centroid-0"] = com[:, 0]
            res["centroid-1"] = com[:, 1]
        else:
            raise KeyError(p)
    return res


def install_shims():
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)

    alb = types.ModuleType("albumentations")
    albt = types.ModuleType("albumentations.pytorch")
    albt.ToTensorV2 = object
    alb.pytorch = albt
    sys.modules.setdefault("albumentations", alb)
    sys.modules.setdefault("albumentations.pytorch", albt)

    sk = types.ModuleType("skimage")
    skm = types.ModuleType("skimage.measure")
    skm.label = _label
    skm.regionprops_table = _regionprops_table
    sk.measure = skm
    sys.modules.setdefault("skimage", sk)
    sys.modules.setdefault("skimage.measure", skm)
