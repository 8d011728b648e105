"""Generate the golden vectors under tests/golden/ by running the REFERENCE's own functions.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_golden.py

The reference (malani86/unet-DC-segmentation) ships no tests for the hot path, so the pin
for the oracle and for the CUDA kernels is the output of the reference functions themselves, imported
UNMODIFIED from /root/reference:

    utils/data_loader.py:11  rolling_ball_correction_rgb   (real OpenCV underneath)
    models/model_2.py:5      UNetDC                        (real torch fp32 underneath)
    models/model.py:7        UNet
    quantify_droplets_batch.py:40,81   preprocess, quantify

Three packages the reference imports are absent from this image and are shimmed in sys.modules:
``matplotlib`` and ``albumentations`` (never called on this path) and ``skimage.measure``, whose
``label`` / ``regionprops_table`` are backed by scipy.ndimage (label-for-label identical to skimage's
4-connectivity raster-order numbering; area = pixel count, centroid = mean coordinate,
equivalent_diameter = sqrt(4*area/pi) as published by scikit-image).
"""
from __future__ import annotations

import sys
import tempfile
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path("/root/reference")


# ----------------------------------------------------------------------------- shims
sys.path.insert(0, str(REPO))
from oracle.shims import install_shims   # noqa: E402  (matplotlib / albumentations / skimage.measure stand-ins)


def import_reference():
    install_shims()
    sys.path.insert(0, str(REF))
    import quantify_droplets_batch as qdb          # noqa: E402
    from models.model import UNet as RefUNet        # noqa: E402
    from models.model_2 import UNetDC as RefUNetDC  # noqa: E402
    from utils.data_loader import rolling_ball_correction_rgb as ref_rb  # noqa: E402
    return qdb, RefUNetDC, RefUNet, ref_rb


# ----------------------------------------------------------------------------- cases
def rolling_ball_cases():
    sys.path.insert(0, str(REPO))
    from unet_dc_segmentation_b200.synth import synthetic_image
    rs = np.random.RandomState(7)
    cases = []
    g = synthetic_image(96, 1)
    cases.append(("synthetic96_r50", np.repeat(g[:, :, None], 3, 2), 50))
    cases.append(("synthetic96_r15", np.repeat(g[:, :, None], 3, 2), 15))
    cases.append(("noise_80x128_r15", rs.randint(0, 256, (80, 128, 3)).astype(np.uint8), 15))
    cases.append(("noise_70x50_r7", rs.randint(0, 256, (70, 50, 3)).astype(np.uint8), 7))
    cases.append(("noise_40x40_r50", rs.randint(0, 256, (40, 40, 3)).astype(np.uint8), 50))
    ramp = (np.add.outer(np.arange(64), np.arange(96)) * 255 // (64 + 96 - 2)).astype(np.uint8)
    cases.append(("ramp_64x96_r20", np.repeat(ramp[:, :, None], 3, 2), 20))
    cases.append(("constant_48x48_r10", np.full((48, 48, 3), 77, np.uint8), 10))
    smooth = synthetic_image(160, 3, n_droplets=40)
    cases.append(("synthetic160_r50", np.repeat(smooth[:, :, None], 3, 2), 50))
    cases.append(("synthetic160_r2", np.repeat(smooth[:, :, None], 3, 2), 2))
    cases.append(("synthetic160_r1", np.repeat(smooth[:, :, None], 3, 2), 1))
    return cases


def mask_cases():
    sys.path.insert(0, str(REPO))
    from unet_dc_segmentation_b200.synth import synthetic_mask
    rs = np.random.RandomState(11)
    cases = []
    cases.append(("discs96", synthetic_mask(96, 40, seed=3), 1, None))
    cases.append(("discs96_min20_px", synthetic_mask(96, 40, seed=3), 20, 3.45))
    cases.append(("noise64_p50", (rs.rand(64, 64) < 0.5).astype(np.uint8), 1, None))
    cases.append(("noise64_p50_min3", (rs.rand(64, 64) < 0.5).astype(np.uint8), 3, 2.0))
    cases.append(("noise_37x53", (rs.rand(37, 53) < 0.6).astype(np.uint8), 2, 1.5))
    cb = (np.add.outer(np.arange(32), np.arange(48)) % 2).astype(np.uint8)
    cases.append(("checkerboard_32x48", cb, 1, 3.45))
    cases.append(("ones_40x40", np.ones((40, 40), np.uint8), 1, None))
    cases.append(("zeros_40x40", np.zeros((40, 40), np.uint8), 1, 3.45))
    cases.append(("all_filtered", cb, 2, None))
    sp = np.zeros((65, 65), np.uint8)      # serpentine: one long component crossing every 32-px tile border
    sp[::2, :] = 1
    for i in range(0, 64, 2):
        sp[i + 1, 64 if (i // 2) % 2 == 0 else 0] = 1
    cases.append(("serpentine_65", sp, 1, None))
    u = np.zeros((70, 70), np.uint8)       # nested U shapes: roots that only meet at the bottom
    for k in range(0, 30, 4):
        u[k:70 - k, k] = 1
        u[k:70 - k, 69 - k] = 1
        u[69 - k, k:70 - k] = 1
    cases.append(("nested_u_70", u, 1, 2.0))
    return cases


def table_to_arrays(df):
    if df.empty:
        return {"n": np.int64(0)}
    out = {"n": np.int64(len(df))}
    for c in df.columns:
        out[c] = df[c].to_numpy()
    return out


def main():
    import torch
    qdb, RefUNetDC, RefUNet, ref_rb = import_reference()
    sys.path.insert(0, str(REPO))
    from unet_dc_segmentation_b200.synth import calibrated_state_dict, synthetic_image

    # ---- state_dict contract of the reference module (keys, shapes, dtypes)
    import json
    ref_sd = RefUNetDC(3, 1).state_dict()
    assert list(ref_sd.keys()) == list(RefUNet(3, 1).state_dict().keys())
    (HERE / "state_dict_keys.json").write_text(json.dumps(
        [[k, list(v.shape), str(v.dtype)] for k, v in ref_sd.items()], indent=0))

    # ---- rolling ball (reference function, real cv2)
    rb = {}
    for name, img, radius in rolling_ball_cases():
        rb[f"{name}/in"] = img
        rb[f"{name}/radius"] = np.int64(radius)
        rb[f"{name}/out"] = ref_rb(img.copy(), radius)
    np.savez_compressed(HERE / "rolling_ball.npz", **rb)
    print("rolling_ball.npz", len(rb) // 3, "cases")

    # ---- quantify (reference function over the scipy-backed skimage shim)
    qd = {}
    for name, mask, min_area, px in mask_cases():
        df = qdb.quantify(mask.copy(), min_area, px)
        qd[f"{name}/mask"] = np.packbits(mask, axis=None)
        qd[f"{name}/shape"] = np.array(mask.shape, np.int64)
        qd[f"{name}/min_area"] = np.int64(min_area)
        qd[f"{name}/px"] = np.float64(px if px else 0.0)
        for k, v in table_to_arrays(df).items():
            qd[f"{name}/t/{k}"] = v
    np.savez_compressed(HERE / "quantify.npz", **qd)
    print("quantify.npz", len(mask_cases()), "cases")

    # ---- network: reference modules, fp32, calibrated synthetic checkpoint regenerated from its seed
    torch.manual_seed(0)
    torch.set_num_threads(1)
    nets = {}
    for tag, cls, dil in (("unetdc", RefUNetDC, (1, 2, 4, 8, 16)), ("unet", RefUNet, (1, 1, 1, 1, 1))):
        sd = calibrated_state_dict(seed=0, calib_size=64, n_calib=2, dilations=dil)
        ref = cls(3, 1) if tag == "unetdc" else cls(3, 1)
        ref.load_state_dict(sd)
        ref.eval()
        imgs = np.stack([synthetic_image(64, 50 + i, n_droplets=10) for i in range(2)])
        x = torch.from_numpy(np.repeat(imgs[:, None], 3, 1).astype(np.float32) / 255.0)
        with torch.no_grad():
            y = ref(x).numpy()
        nets[f"{tag}/images"] = imgs
        nets[f"{tag}/probs"] = y.astype(np.float32)
        nets[f"{tag}/dilations"] = np.array(dil, np.int64)
        # fingerprint of the regenerated checkpoint so a test can tell "weights differ" from "kernel differs"
        nets[f"{tag}/sd_checksum"] = np.float64(sum(float(v.double().abs().sum()) for v in sd.values()))
        print(tag, "probs range", float(y.min()), float(y.max()), "frac>0.3", float((y > 0.3).mean()))
    np.savez_compressed(HERE / "forward.npz", **nets)

    # ---- whole path at native size: preprocess (qdb:40-46) -> model -> threshold -> quantify
    sd = calibrated_state_dict(seed=0, calib_size=64, n_calib=2)
    ref = RefUNetDC(3, 1)
    ref.load_state_dict(sd)
    ref.eval()
    e2e = {}
    from PIL import Image
    with tempfile.TemporaryDirectory() as td:
        size = 96
        qdb.IMG_SIZE = size          # identity resizes (the kernel-parity variant of config 1)
        tensors, imgs = [], []
        for i in range(2):
            g = synthetic_image(size, 200 + i, n_droplets=12)
            p = Path(td) / f"img{i}.png"
            Image.fromarray(g).save(p)
            t, (oh, ow) = qdb.preprocess(p, 50)
            assert (oh, ow) == (size, size)
            tensors.append(t)
            imgs.append(g)
        with torch.no_grad():
            probs = ref(torch.stack(tensors)).numpy()
        e2e["images"] = np.stack(imgs)
        e2e["pre"] = (torch.stack(tensors).numpy() * 255.0).round().astype(np.uint8)   # rolling-ball output, CHW
        e2e["probs"] = probs.astype(np.float32)
        for i in range(2):
            mask = (probs[i, 0] > 0.3).astype(np.uint8)
            df = qdb.quantify(mask, 1, 3.45)
            e2e[f"mask{i}"] = np.packbits(mask, axis=None)
            for k, v in table_to_arrays(df).items():
                e2e[f"t{i}/{k}"] = v
            print("e2e image", i, "droplets", len(df), "fg", float(mask.mean()))
    np.savez_compressed(HERE / "end_to_end.npz", **e2e)

    # ---- the two resizes exactly as the reference calls them (interpolation flag in the `dst` slot, qdb:44,57)
    import cv2
    rs = np.random.RandomState(21)
    rz = {}
    for name, shape, dsize in (("rgb_69x102_to_128", (69, 102, 3), (128, 128)), ("rgb_64_to_128", (64, 64, 3), (128, 128)),
                               ("rgb_96_to_96", (96, 96, 3), (96, 96)), ("rgb_175x250_to_128", (175, 250, 3), (128, 128)),
                               ("rgb_17x23_to_64", (17, 23, 3), (64, 64))):
        im = rs.randint(0, 256, shape).astype(np.uint8)
        rz[f"{name}/in"] = im
        rz[f"{name}/dsize"] = np.array(dsize, np.int64)
        rz[f"{name}/out"] = cv2.resize(im, dsize, cv2.INTER_AREA)              # qdb:44, verbatim call form
    for name, shape, dsize in (("mask_128_to_102x69", (128, 128), (102, 69)), ("mask_128_to_64", (128, 128), (64, 64)),
                               ("mask_128_to_250x175", (128, 128), (250, 175)), ("mask_64_to_23x17", (64, 64), (23, 17))):
        m = (rs.rand(*shape) < 0.3).astype(np.uint8)
        rz[f"{name}/in"] = m
        rz[f"{name}/dsize"] = np.array(dsize, np.int64)
        rz[f"{name}/out"] = cv2.resize(m, dsize, cv2.INTER_NEAREST)            # qdb:57, verbatim call form
    np.savez_compressed(HERE / "resize.npz", **rz)
    print("resize.npz", len(rz) // 3, "cases")

    # ---- the reference's shipped sample outputs: formula / column known answers
    import pandas as pd
    df = pd.read_csv(REF / "outputs" / "all_droplets.csv")
    summ = pd.read_csv(REF / "outputs" / "summary_per_image.csv")
    np.savez_compressed(
        HERE / "reference_outputs.npz",
        columns=np.array(list(df.columns)),
        filename=df["filename"].to_numpy().astype(str), label=df["label"].to_numpy(),
        area=df["area"].to_numpy(), equivalent_diameter=df["equivalent_diameter"].to_numpy(),
        centroid0=df["centroid-0"].to_numpy(), centroid1=df["centroid-1"].to_numpy(),
        area_sqmicron=df["area_sqmicron"].to_numpy(), eq_diam_micron=df["eq_diam_micron"].to_numpy(),
        summary_columns=np.array(list(summ.columns)), summary_filename=summ["filename"].to_numpy().astype(str),
        summary_count=summ["droplet_count"].to_numpy(), summary_area=summ["total_area_px"].to_numpy())
    print("reference_outputs.npz", len(df), "rows")


if __name__ == "__main__":
    main()
