"""CPU, world_size 2 over gloo: the N > 1 host path -- frames sharded i -> rank i mod G, every rank works on
its own shard with no collective, the per-frame droplet tables are gathered on rank 0 in frame order."""
import os
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent

WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["REPO"])
import oracle                                    # the CPU stand-in for the per-rank device work in this test
from unet_dc_segmentation_b200 import shard
from unet_dc_segmentation_b200.synth import synthetic_mask

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
N = 7
mine = shard.shard_indices(N, rank, world)
local = []
for group in shard.batches(mine, 2):
    for i in group:
        _, cols = oracle.quantify_arrays(synthetic_mask(48, 12, seed=i), 1, 3.45)
        local.append((i, cols))
merged = shard.gather_results(local, N, dst=0)
if rank == 0:
    assert len(merged) == N
    for i, cols in enumerate(merged):
        _, want = oracle.quantify_arrays(synthetic_mask(48, 12, seed=i), 1, 3.45)
        for k in want:
            assert np.array_equal(cols[k], want[k]), (i, k)
    print("MERGED_OK", [len(c["label"]) for c in merged])
else:
    assert merged is None
dist.barrier()
dist.destroy_process_group()
"""


def test_two_ranks_shard_and_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, REPO=str(REPO), MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MERGED_OK" in r.stdout


TABLES_WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["REPO"])
from unet_dc_segmentation_b200 import shard
from unet_dc_segmentation_b200.quantify import DropletTables

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
N, CAP = 11, 9                                    # 11 frames over 2 ranks: rank 0 holds 6, rank 1 holds 5 (ragged)

def frame_table(i):                               # a function of the GLOBAL frame index only
    rs = np.random.RandomState(100 + i)
    n = int(rs.randint(0, CAP + 1))
    return n, rs.randint(1, 1000, (CAP, 6)).astype(np.float64) + i / 16.0

mine = shard.shard_indices(N, rank, world)
counts = torch.zeros(len(mine), dtype=torch.int32)
cols = [torch.zeros((len(mine), CAP), dtype=torch.float64) for _ in range(6)]
for j, i in enumerate(mine):
    n, t = frame_table(i)
    counts[j] = n
    for c in range(6):
        cols[c][j] = torch.from_numpy(t[:, c])
archive = DropletTables(counts, cols[0].view(torch.int64), cols[2], cols[3], cols[1], cols[4], cols[5], None, CAP)
rows, cnt = shard.gather_tables_in_frame_order(archive, N)
if rank == 0:
    want = [frame_table(i) for i in range(N)]
    assert cnt.tolist() == [n for n, _ in want]
    pos = 0
    for i, (n, t) in enumerate(want):
        assert np.array_equal(rows[pos:pos + n].view(np.int64), np.ascontiguousarray(t[:n]).view(np.int64)), i
        pos += n
    assert pos == rows.shape[0]
    print("TABLES_OK", cnt.tolist())
else:
    assert rows is None and cnt is None
dist.barrier()
dist.destroy_process_group()
"""


def test_two_ranks_gather_tables_in_frame_order(tmp_path):
    """shard.gather_tables_in_frame_order (what bench.py's config-4 job and sharded jobs use): compaction, one gather,
    frame-order permutation -- over gloo with CPU tensors, ragged shard sizes, frames without droplets."""
    script = tmp_path / "tables_worker.py"
    script.write_text(TABLES_WORKER)
    env = dict(os.environ, REPO=str(REPO), MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29519", str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "TABLES_OK" in r.stdout


def test_gather_tables_single_process():
    """No process group: the identity sharding (rank 0 of 1)."""
    import numpy as np
    import torch
    from unet_dc_segmentation_b200 import shard
    from unet_dc_segmentation_b200.quantify import DropletTables
    counts = torch.tensor([2, 0, 3], dtype=torch.int32)
    base = torch.arange(3 * 4, dtype=torch.float64).reshape(3, 4)
    t = DropletTables(counts, (base * 2).to(torch.int64), base + 0.25, base + 0.5, base + 0.125, None, None, None, 4)
    rows, cnt = shard.gather_tables_in_frame_order(t, 3)
    assert cnt.tolist() == [2, 0, 3] and rows.shape == (5, 4)
    np.testing.assert_array_equal(rows[:, 0].view(np.int64), [0, 2, 16, 18, 20])
    np.testing.assert_array_equal(rows[:, 2], [0.25, 1.25, 8.25, 9.25, 10.25])
