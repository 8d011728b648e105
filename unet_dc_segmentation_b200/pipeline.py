"""Fused device pipeline: u8 images -> (u8 masks, droplet tables).

The device-side composition of reference quantify_droplets_batch.py: ``preprocess`` (:40-46, at native
size, where both cv2.resize calls are the identity) -> ``model(batch)`` (:52) -> ``> thresh`` (:56) ->
``quantify`` (:61, :81-95), for a whole batch at a time with nothing but the u8 images going up and
the u8 masks + per-droplet rows coming back.

Stages, all on the current CUDA stream (no host synchronisation until the tables are fetched):
  dc_rolling_ball   (grayscale input: one plane per image -- the three RGB channels the reference
                     builds with Image.convert("RGB") are identical, so they are corrected once)
  dc_forward        stem reads the corrected u8 plane and does the /255 (qdb:45); the last conv's
                     epilogue does out_conv + sigmoid + `> prob_thresh` and writes the u8 mask
  dc_label_stats    labelling, min_area filter, compaction, area / centroid / diameter (+ micron columns)
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .model import UNetDC
from .morphology import rolling_ball_device, rolling_ball_workspace_bytes
from .quantify import DEFAULT_CAPACITY, DropletTables, label_stats_device, label_workspace_bytes


@dataclass
class BatchResult:
    masks: torch.Tensor            # u8 [B,H,W] {0,1} (device)
    tables: DropletTables          # device-resident table
    probs: torch.Tensor | None     # f32 [B,1,H,W] when requested


class DropletPipeline:
    def __init__(self, model: UNetDC, background_radius: int | None = 50, prob_thresh: float = 0.3,
                 min_area: int = 1, px_per_micron: float | None = None, capacity: int = DEFAULT_CAPACITY):
        self.model = model
        self.background_radius = background_radius
        self.prob_thresh = float(prob_thresh)
        self.min_area = int(min_area)
        self.px_per_micron = px_per_micron
        self.capacity = int(capacity)
        self._rb_ws = None
        self._rb_out = None
        self._ccl_ws = None
        self._staging = None

    # ------------------------------------------------------------------ device-resident entry
    def run_device(self, images: torch.Tensor, return_prob: bool = False, want_labels: bool = False) -> BatchResult:
        """images: CUDA u8 [B,H,W] grayscale or [B,H,W,3]."""
        _lib.require_cuda(images, "images")
        x = images
        if self.background_radius:
            if self._rb_out is None or self._rb_out.shape != x.shape or self._rb_out.device != x.device:
                self._rb_out = torch.empty_like(x)
                self._rb_ws = None
            if self._rb_ws is None:
                need = rolling_ball_workspace_bytes(x.shape[0], x.shape[1], x.shape[2], 1 if x.dim() == 3 else x.shape[3])
                self._rb_ws = torch.empty(need, dtype=torch.uint8, device=x.device)
            x = rolling_ball_device(x, self.background_radius, out=self._rb_out, workspace=self._rb_ws)
        masks, probs = self.model.predict_u8(x, self.prob_thresh, return_prob=return_prob)
        need = label_workspace_bytes(*masks.shape)
        if self._ccl_ws is None or self._ccl_ws.device != masks.device or self._ccl_ws.numel() < need:
            self._ccl_ws = torch.empty(need, dtype=torch.uint8, device=masks.device)
        tables = label_stats_device(masks, self.min_area, self.px_per_micron, self.capacity,
                                    want_labels=want_labels, workspace=self._ccl_ws)
        return BatchResult(masks, tables, probs)

    # ------------------------------------------------------------------ host entry (what a user calls)
    def run_host(self, images: torch.Tensor | np.ndarray, device: torch.device | str = "cuda"):
        """images: host u8 [B,H,W] or [B,H,W,3] (pinned for async copies).  Returns
        (masks u8 numpy [B,H,W], list of per-image column dicts).  H2D of the images and D2H of the
        masks and table rows are part of this call."""
        if isinstance(images, np.ndarray):
            images = torch.from_numpy(np.ascontiguousarray(images))
        if images.dtype != torch.uint8:
            raise TypeError("images must be uint8")
        dev = torch.device(device)
        if self._staging is None or self._staging.shape != images.shape or self._staging.device != dev:
            self._staging = torch.empty(images.shape, dtype=torch.uint8, device=dev)
        self._staging.copy_(images, non_blocking=True)
        res = self.run_device(self._staging)
        masks = res.masks.to("cpu", non_blocking=False)
        tables = res.tables
        nmax = int(tables.counts.max().item())
        if nmax > tables.capacity:
            self.capacity = nmax
            tables = label_stats_device(res.masks, self.min_area, self.px_per_micron, nmax)
        return masks.numpy(), tables.to_host()
