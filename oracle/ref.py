"""oracle/_ref: the reference's OWN modules, byte-compiled -- TEST / BASELINE INFRASTRUCTURE ONLY.

The reference (malani86/unet-DC-segmentation) is pure Python, so "compiling it where it lies" means
``py_compile``: ``build_ref()`` (run by ``__graft_entry__.build()`` in the build container, where /root/reference
exists) byte-compiles the five modules behind quantify_droplets_batch.py into ``oracle/_ref/*.refc`` (ordinary .pyc content; the extension differs because snapshot tools drop ``*.pyc``).  No reference
SOURCE is copied: the directory is git-ignored, holds compiled code objects only, and travels to the GPU box with the
snapshot like our own built ``.so`` files.  ``load()`` imports those code objects (under the sys.modules shims of
oracle/shims.py for the three third-party packages this image lacks) and returns the reference's functions, which
``bench.py --impl reference`` and the ``cpu_baseline`` leg time and whose outputs the bench's ``parity`` object is
measured against.  Nothing under ``unet_dc_segmentation_b200/`` imports this."""
from __future__ import annotations

import hashlib
import importlib.abc
import importlib.machinery
import importlib.util
import json
import py_compile
import sys
from pathlib import Path
from types import SimpleNamespace

_HERE = Path(__file__).resolve().parent
REF_SRC = Path("/root/reference")
REF_OUT = _HERE / "_ref"
# module name -> file relative to the reference root
EXT = ".refc"              # byte-compiled module (the .pyc format under a name snapshot tools do not filter out)
MODULES = {
    "quantify_droplets_batch": "quantify_droplets_batch.py",
    "models": "models/__init__.py",
    "models.model_2": "models/model_2.py",
    "models.model": "models/model.py",
    "utils": "utils/__init__.py",
    "utils.data_loader": "utils/data_loader.py",
}


def build_ref(force: bool = False) -> bool:
    """Byte-compile the reference modules into oracle/_ref/ (build container only).  Returns False when
    /root/reference is absent (the GPU box: it uses the prebuilt files)."""
    if not REF_SRC.exists():
        return False
    REF_OUT.mkdir(exist_ok=True)
    manifest = {"python": sys.version.split()[0], "magic": importlib.util.MAGIC_NUMBER.hex(), "modules": {}}
    for name, rel in MODULES.items():
        src = REF_SRC / rel
        dst = REF_OUT / (rel[:-3] + EXT)
        dst.parent.mkdir(parents=True, exist_ok=True)
        if force or not dst.exists() or dst.stat().st_mtime < src.stat().st_mtime:
            py_compile.compile(str(src), cfile=str(dst), dfile=f"<reference>/{rel}", doraise=True, optimize=0)
        manifest["modules"][name] = {"file": rel[:-3] + EXT, "source_sha256": hashlib.sha256(src.read_bytes()).hexdigest()}
    (REF_OUT / "MANIFEST.json").write_text(json.dumps(manifest, indent=1))
    return True


def available() -> bool:
    m = REF_OUT / "MANIFEST.json"
    if not m.exists():
        return False
    try:
        d = json.loads(m.read_text())
    except ValueError:
        return False
    return d.get("magic") == importlib.util.MAGIC_NUMBER.hex() and all(
        (REF_OUT / v["file"]).exists() for v in d["modules"].values())


class _RefFinder(importlib.abc.MetaPathFinder):
    """Resolves exactly the reference's module names to the compiled files under oracle/_ref."""

    def find_spec(self, fullname, path=None, target=None):
        rel = MODULES.get(fullname)
        if rel is None:
            return None
        pyc = REF_OUT / (rel[:-3] + EXT)
        if not pyc.exists():
            return None
        loader = importlib.machinery.SourcelessFileLoader(fullname, str(pyc))
        is_pkg = rel.endswith("__init__.py")
        return importlib.util.spec_from_file_location(fullname, str(pyc), loader=loader,
                                                      submodule_search_locations=[str(pyc.parent)] if is_pkg else None)


_REF = None


def load() -> SimpleNamespace:
    """The reference's functions: qdb (module), UNetDC, UNet, rolling_ball_correction_rgb, quantify, preprocess."""
    global _REF
    if _REF is not None:
        return _REF
    if not available():
        raise RuntimeError("oracle/_ref is missing or was compiled by another Python: run __graft_entry__.build() "
                           "in the build container (where /root/reference exists)")
    from .shims import install_shims
    install_shims()
    clash = [n for n in MODULES if n in sys.modules and not str(getattr(sys.modules[n], "__file__", "")).startswith(str(REF_OUT))]
    if clash:
        raise RuntimeError(f"modules {clash} are already imported from elsewhere; cannot load the reference beside them")
    sys.meta_path.insert(0, _RefFinder())
    import quantify_droplets_batch as qdb
    from models.model import UNet
    from models.model_2 import UNetDC
    from utils.data_loader import rolling_ball_correction_rgb
    _REF = SimpleNamespace(qdb=qdb, UNetDC=UNetDC, UNet=UNet, rolling_ball_correction_rgb=rolling_ball_correction_rgb,
                           quantify=qdb.quantify, preprocess=qdb.preprocess)
    return _REF


def run_path(state_dict, images_u8, radius=50, prob_thresh=0.3, min_area=1, px_per_um=None):
    """The reference's hot path on in-memory frames at native size, calling the reference's own functions in the
    order quantify_droplets_batch.py does (file decode / PNG + CSV writes left out, as in the GPU arm):
    rolling_ball_correction_rgb (qdb:43) -> /255, CHW (qdb:45-46) -> stack + UNetDC (qdb:51-52) -> `> thresh` (qdb:56)
    -> quantify (qdb:61).  The two cv2.resize calls (qdb:44,57) are the identity at native size and are skipped.
    images_u8: list of u8 [H,W,3].  Returns (probs f32 [B,H,W], masks u8 [B,H,W], list of DataFrames)."""
    import numpy as np
    import torch
    ref = load()
    model = ref.UNetDC(in_channels=3, out_channels=1)
    model.load_state_dict(state_dict)
    model = model.eval()
    tensors = []
    for im in images_u8:
        im = ref.rolling_ball_correction_rgb(im, radius)
        tensors.append(torch.from_numpy(im.astype(np.float32) / 255.0).permute(2, 0, 1))
    with torch.no_grad():
        logits = model(torch.stack(tensors))
    probs = logits[:, 0].cpu().numpy()
    masks = (probs > prob_thresh).astype(np.uint8)
    tables = [ref.quantify(m, min_area, px_per_um) for m in masks]
    return probs, masks, tables
