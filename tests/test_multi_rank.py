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
