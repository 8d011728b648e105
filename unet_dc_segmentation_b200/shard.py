"""Sharding of independent frames across the GPUs of one box (BASELINE config 4).

Frames are independent (nothing in reference quantify_droplets_batch.py:146-160 carries state from one image
to the next), so the path shards with NO data-path collective: rank r of G takes frames r, r+G, r+2G, ...
(image index i -> GPU i mod G), runs its own stream pipeline, and only the small per-droplet tables travel
back -- gathered on rank 0 and re-interleaved into frame order.  Labels are per image (1..n), so no
renumbering is needed.  torch.distributed is used for the gather only (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations


def shard_indices(n_items: int, rank: int, world: int) -> list[int]:
    """Frame indices owned by `rank`: i with i mod world == rank, ascending."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_items, world))


def batches(indices: list[int], batch: int) -> list[list[int]]:
    """Consecutive groups of at most `batch` of this rank's frames (the tail batch may be short, qdb:157-160)."""
    if batch < 1:
        raise ValueError("batch must be >= 1")
    return [indices[i:i + batch] for i in range(0, len(indices), batch)]


def merge_in_frame_order(per_rank: list[list[tuple[int, object]]], n_items: int) -> list[object]:
    """per_rank[r] = [(frame index, result), ...] -> results in frame order; checks every frame appears once."""
    out: list[object] = [None] * n_items
    seen = [False] * n_items
    for items in per_rank:
        for idx, res in items:
            if not 0 <= idx < n_items or seen[idx]:
                raise ValueError(f"frame {idx} missing from the shard map or delivered twice")
            seen[idx] = True
            out[idx] = res
    if not all(seen):
        raise ValueError(f"frames never delivered: {[i for i, s in enumerate(seen) if not s][:8]} ...")
    return out


def gather_results(local: list[tuple[int, object]], n_items: int, dst: int = 0):
    """Gather every rank's (frame index, result) pairs on `dst` and return them in frame order there
    (None on the other ranks).  With no process group initialised this is the single-GPU identity."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return merge_in_frame_order([local], n_items)
    world, rank = dist.get_world_size(), dist.get_rank()
    buf = [None] * world if rank == dst else None
    dist.gather_object(local, buf, dst=dst)
    if rank != dst:
        return None
    return merge_in_frame_order(buf, n_items)
