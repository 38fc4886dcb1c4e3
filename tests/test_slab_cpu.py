"""Host side of the multi-GPU path on CPU: world_size-2 gloo processes agree on the slab partition and on the
rendezvous of the communicator id (the NCCL id itself needs a GPU box: tests/test_multigpu.py)."""
import os
import subprocess
import sys

import pytest

from spirit_b200 import slab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from spirit_b200 import slab
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
parts = slab.partition(257, world)
mine = parts[rank]
gathered = [None] * world
dist.all_gather_object(gathered, mine)
assert gathered == parts, (gathered, parts)
covered = sorted(c for b, n in gathered for c in range(b, b + n))
assert covered == list(range(257))
uid = slab.broadcast_unique_id(lambda: bytes(range(128)), dist, rank)
assert uid == bytes(range(128))
dist.barrier()
if rank == 0:
    print("SLAB_OK", parts)
dist.destroy_process_group()
'''


def test_partition_tiles_the_lattice():
    for nc in (1, 2, 7, 256, 257, 512):
        for world in (1, 2, 3, 4, 8):
            if world > nc:
                continue
            parts = slab.partition(nc, world)
            assert parts[0][0] == 0 and sum(n for _, n in parts) == nc
            for (b0, n0), (b1, _) in zip(parts, parts[1:]):
                assert b0 + n0 == b1
            assert max(n for _, n in parts) - min(n for _, n in parts) <= 1


def test_two_gloo_ranks_agree(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29713", str(script), ROOT],
        capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "SLAB_OK" in r.stdout
