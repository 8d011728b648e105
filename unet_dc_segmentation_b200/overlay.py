"""Overlay contours of reference quantify_droplets_batch.py:74-79 on the GPU.

    cnts, _ = cv2.findContours(mask, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)      # qdb:76
    cv2.drawContours(img, cnts, -1, (0, 255, 0), 2)                                   # qdb:77

``overlay_stencil_device`` returns the set of pixels those two calls paint (bit-exact against OpenCV, see
csrc/ccl.cu and tests/test_gpu_overlay.py); ``draw_overlay`` applies it to a BGR frame.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

OVERLAY_BGR = (0, 255, 0)        # qdb:77
OVERLAY_THICKNESS = 2            # qdb:77 (the only thickness the kernels implement)


def overlay_workspace_bytes(B: int, H: int, W: int) -> int:
    need = C.c_size_t()
    _lib.check(_lib.load().dc_overlay_workspace_bytes(B, H, W, C.byref(need)))
    return int(need.value)


def overlay_stencil_device(masks: torch.Tensor, out: torch.Tensor | None = None,
                           workspace: torch.Tensor | None = None) -> torch.Tensor:
    """masks: CUDA u8 [B,H,W] (non-zero = foreground) -> u8 [B,H,W], 1 where the overlay is painted."""
    _lib.require_cuda(masks, "masks")
    if masks.dtype != torch.uint8 or masks.dim() != 3:
        raise TypeError("masks must be uint8 [B,H,W]")
    masks = masks.contiguous()
    B, H, W = masks.shape
    dev = masks.device
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
        elif out.shape != masks.shape or out.dtype != torch.uint8 or not out.is_contiguous() or out.device != dev:
            raise ValueError("out must be a contiguous uint8 [B,H,W] tensor on the masks' device")
        need = overlay_workspace_bytes(B, H, W)
        if workspace is None or workspace.numel() < need:
            workspace = torch.empty(need, dtype=torch.uint8, device=dev)
        args = _lib.OverlayArgs(masks.data_ptr(), B, H, W, out.data_ptr(), workspace.data_ptr(), workspace.numel())
        _lib.check(_lib.load().dc_overlay_stencil(C.byref(args), _lib.stream_ptr(dev)))
    return out


def draw_overlay(img_bgr: np.ndarray, stencil: np.ndarray) -> np.ndarray:
    """Paint the stencil into a BGR frame in place, as drawContours does (qdb:77), and return it."""
    img_bgr[stencil.astype(bool)] = OVERLAY_BGR
    return img_bgr
