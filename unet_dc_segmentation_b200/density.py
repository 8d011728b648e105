"""Density maps of the reference's alternate front end, quantify_pipline.py, on the GPU.

    generate_roi_mask(img)                                   quantify_pipline.py:44-51
    get_targets(mask_thresh, mask_contour, nb_layers, centroid_y, centroid_x)   :61-91
    density_maps(mask_thresh, mask_contour, kernel_size=21)  :93-97
    normalize(img)                                           :53-57

The three functions above keep the reference's numpy-in / numpy-out signatures; the ``*_device`` variants are batched and
stay on the GPU.  All results are bit-exact against the OpenCV / numpy / scipy calls of the reference (csrc/density.cu).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .quantify import DEFAULT_CAPACITY, DropletTables, label_stats_device


def _ws(nbytes: int, dev) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=dev)


def roi_mask_device(rgb: torch.Tensor):
    """rgb: CUDA u8 [B,H,W,3] -> (roi u8 [B,H,W] {0,1}, centroid int32 [B,2] = (cy, cx)); qpl:44-51 and :133-136."""
    _lib.require_cuda(rgb, "rgb")
    if rgb.dtype != torch.uint8 or rgb.dim() != 4 or rgb.shape[-1] != 3:
        raise TypeError("rgb must be uint8 [B,H,W,3]")
    rgb = rgb.contiguous()
    B, H, W = rgb.shape[:3]
    dev = rgb.device
    lib = _lib.load()
    with torch.cuda.device(dev):
        need = C.c_size_t()
        _lib.check(lib.dc_roi_workspace_bytes(B, H, W, C.byref(need)))
        ws = _ws(need.value, dev)
        roi = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
        cen = torch.empty((B, 2), dtype=torch.int32, device=dev)
        args = _lib.RoiArgs(rgb.data_ptr(), B, H, W, roi.data_ptr(), cen.data_ptr(), ws.data_ptr(), ws.numel())
        _lib.check(lib.dc_roi_mask(C.byref(args), _lib.stream_ptr(dev)))
    return roi, cen


def radial_density_device(roi: torch.Tensor, centroid: torch.Tensor, tables: DropletTables, nb_layers: int = 10) -> torch.Tensor:
    """get_targets (qpl:61-91) for a batch: roi u8 [B,H,W], centroid int32 [B,2] (cy, cx), tables = label_stats_device
    of the droplet masks with min_area = 1 (every droplet must fit the tables' capacity).  -> f32 [B,H,W]."""
    _lib.require_cuda(roi, "roi")
    if roi.dtype != torch.uint8 or roi.dim() != 3:
        raise TypeError("roi must be uint8 [B,H,W]")
    roi = roi.contiguous()
    B, H, W = roi.shape
    dev = roi.device
    centroid = centroid.to(device=dev, dtype=torch.int32).contiguous()
    if centroid.shape != (B, 2) or tables.counts.shape[0] != B:
        raise ValueError("centroid / tables do not match the batch")
    lib = _lib.load()
    with torch.cuda.device(dev):
        need = C.c_size_t()
        _lib.check(lib.dc_radial_workspace_bytes(B, C.byref(need)))
        ws = _ws(need.value, dev)
        out = torch.empty((B, H, W), dtype=torch.float32, device=dev)
        args = _lib.RadialArgs(roi.data_ptr(), B, H, W, centroid.data_ptr(), tables.counts.data_ptr(),
                               tables.centroid0.data_ptr(), tables.centroid1.data_ptr(), int(tables.capacity), int(nb_layers),
                               out.data_ptr(), ws.data_ptr(), ws.numel())
        _lib.check(lib.dc_radial_density(C.byref(args), _lib.stream_ptr(dev)))
    return out


def gaussian_weights(sigma: float, truncate: float = 4.0):
    """The taps scipy.ndimage.gaussian_filter builds (scipy _gaussian_kernel1d, order 0): (f64 [2r+1], r)."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return np.ascontiguousarray(phi / phi.sum(), dtype=np.float64), radius


def spatial_density_device(mask: torch.Tensor, roi: torch.Tensor, kernel_size: int = 21) -> torch.Tensor:
    """density_maps (qpl:93-97) for a batch: mask, roi u8 [B,H,W] -> f32 [B,H,W]."""
    _lib.require_cuda(mask, "mask")
    _lib.require_cuda(roi, "roi")
    if mask.dtype != torch.uint8 or roi.dtype != torch.uint8 or mask.dim() != 3 or mask.shape != roi.shape:
        raise TypeError("mask and roi must be uint8 [B,H,W] of the same shape")
    mask, roi = mask.contiguous(), roi.contiguous()
    B, H, W = mask.shape
    dev = mask.device
    w, radius = gaussian_weights(kernel_size / 6)
    lib = _lib.load()
    with torch.cuda.device(dev):
        need = C.c_size_t()
        _lib.check(lib.dc_spatial_workspace_bytes(B, H, W, C.byref(need)))
        ws = _ws(need.value, dev)
        out = torch.empty((B, H, W), dtype=torch.float32, device=dev)
        args = _lib.SpatialArgs(mask.data_ptr(), roi.data_ptr(), B, H, W, radius, w.ctypes.data, out.data_ptr(),
                                ws.data_ptr(), ws.numel())
        _lib.check(lib.dc_spatial_density(C.byref(args), _lib.stream_ptr(dev)))
    return out


# ------------------------------------------------------------------ the reference's own signatures (numpy in / out)
def _dev(device):
    if not torch.cuda.is_available():
        raise RuntimeError("unet_dc_segmentation_b200 needs an sm_100 GPU; there is no CPU path")
    return torch.device(device)


def generate_roi_mask(img, blur_kernel: int = 15, device="cuda"):
    """quantify_pipline.py:44-51.  img: u8 [H,W,3] RGB -> u8 [H,W] {0,1}."""
    if blur_kernel != 15:
        raise ValueError("the kernels implement the reference's blur_kernel = 15")
    x = torch.from_numpy(np.ascontiguousarray(img, dtype=np.uint8)).to(_dev(device))[None]
    return roi_mask_device(x)[0][0].cpu().numpy()


def get_targets(mask_thresh, mask_contour, nb_layers, centroid_y, centroid_x, device="cuda"):
    """quantify_pipline.py:61-91.  -> f32 [H,W]."""
    dev = _dev(device)
    m = torch.from_numpy(np.ascontiguousarray(mask_thresh, dtype=np.uint8)).to(dev)[None]
    roi = torch.from_numpy(np.ascontiguousarray(mask_contour).astype(np.uint8)).to(dev)[None]
    t = label_stats_device(m, 1, None, DEFAULT_CAPACITY)
    n = int(t.counts.max().item())
    if n > t.capacity:
        t = label_stats_device(m, 1, None, n)
    cen = torch.tensor([[int(centroid_y), int(centroid_x)]], dtype=torch.int32, device=dev)
    return radial_density_device(roi, cen, t, nb_layers)[0].cpu().numpy()


def density_maps(mask_thresh, mask_contour, kernel_size: int = 21, device="cuda"):
    """quantify_pipline.py:93-97.  -> f32 [H,W]."""
    dev = _dev(device)
    m = torch.from_numpy(np.ascontiguousarray(mask_thresh, dtype=np.uint8)).to(dev)[None]
    roi = torch.from_numpy(np.ascontiguousarray(mask_contour).astype(np.uint8)).to(dev)[None]
    return spatial_density_device(m, roi, kernel_size)[0].cpu().numpy()


def normalize(img):
    """quantify_pipline.py:53-57 (host; feeds plt.imsave)."""
    img_min, img_max = np.min(img), np.max(img)
    if img_max > img_min:
        return (img - img_min) / (img_max - img_min)
    return img
