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


def gather_tables_in_frame_order(archive, n_frames_total: int, host_out=None, timing: dict | None = None):
    """Sharded runs (frame i -> rank i mod G): gather every rank's archived droplet tables on rank 0 and return them
    there in FRAME order as (rows, counts) -- rows f64 [sum(counts), C] (columns as DropletTables.compact_rows: area's
    int64 bit pattern, equivalent_diameter, centroid-0, centroid-1 [, area_sqmicron, eq_diam_micron]), counts int64
    [n_frames_total] -- both numpy on the host; (None, None) on the other ranks.

    `archive`: this rank's DropletTables holding its frames in order (image j of rank r is frame r + j*G), as filled
    by DropletPipeline.run_host_pipelined(tables_archive=...).  The tables never pass through the ranks' hosts: they are
    compacted on the device, gathered with ONE collective (NCCL on GPUs; gloo in the CPU tests), permuted into frame
    order on rank 0's device and read back once (into `host_out`, a pinned f64 [>= rows, C] tensor, when given).
    Labels are per image (row r of an image = label r + 1), so no renumbering is needed."""
    import time
    import torch
    import torch.distributed as dist
    t0 = time.perf_counter()
    dev = archive.counts.device

    def sync():
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)

    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    world = dist.get_world_size() if multi else 1
    rank = dist.get_rank() if multi else 0
    nf = (n_frames_total + world - 1) // world
    counts, rows = archive.compact_rows()                      # int64 [frames of this rank], f64 [R, C]
    ncol = rows.shape[1]
    sync()
    if timing is not None:
        timing["compact"] = time.perf_counter() - t0
    if multi:
        n_rows = torch.tensor([rows.shape[0]], dtype=torch.int64, device=dev)
        all_rows = [torch.zeros_like(n_rows) for _ in range(world)]
        dist.all_gather(all_rows, n_rows)
        maxr = int(max(int(v.item()) for v in all_rows))
        buf = torch.zeros((maxr, ncol), dtype=torch.float64, device=dev)
        buf[:rows.shape[0]] = rows
        cbuf = torch.zeros(nf, dtype=torch.int64, device=dev)
        cbuf[:counts.shape[0]] = counts
        gr = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
        gc = [torch.empty_like(cbuf) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, gr, dst=0)
        dist.gather(cbuf, gc, dst=0)
    else:
        gr, gc, maxr = [rows], [counts], rows.shape[0]
    sync()
    if timing is not None:
        timing["gather"] = time.perf_counter() - t0
    if rank != 0:
        return None, None
    # frame order on the device: frame i is image i // world of rank i % world
    cnt = torch.stack([torch.nn.functional.pad(c, (0, nf - c.shape[0])) for c in gc])          # [world, nf]
    off = torch.cumsum(cnt, 1) - cnt                                                           # row offset inside a rank
    fi = torch.arange(n_frames_total, device=dev)
    r_of, j_of = fi % world, fi // world
    f_cnt = cnt[r_of, j_of]
    f_start = r_of * maxr + off[r_of, j_of]                                                    # row in the stacked gather
    out_start = torch.cumsum(f_cnt, 0) - f_cnt
    total = int(f_cnt.sum().item())
    idx = torch.repeat_interleave(f_start - out_start, f_cnt, output_size=total) + torch.arange(total, device=dev)
    stacked = torch.cat([g[:maxr] for g in gr]) if multi else gr[0]
    merged_dev = stacked[idx]
    if host_out is not None and total <= host_out.shape[0]:
        host_out[:total].copy_(merged_dev, non_blocking=True)
        sync()
        merged = host_out[:total].numpy()
    else:
        merged = merged_dev.cpu().numpy()
    return merged, f_cnt.cpu().numpy()
