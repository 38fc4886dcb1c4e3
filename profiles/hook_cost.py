#!/usr/bin/env python
"""Cost of the hook iteration (the last iteration of a block also produces energy, max torque and the projected effective field):
blocks of 1, 2, 5, 20 and 100 Depondt iterations on the 256^3 bench workload, CUDA-event time per block."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import fill_random, write_cfg  # noqa: E402
from spirit_b200 import capi, session as S  # noqa: E402

lib = capi.load_product()
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
lib.SpiritB200_Set_Device(local)
if world > 1:  # slabs of 256 planes per rank (weak scaling, as bench.py): python -m torch.distributed.run ... profiles/hook_cost.py
    import torch
    import torch.distributed as dist
    from spirit_b200 import slab
    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local))
    slab.init_comm(lib, dist, rank, world)
p = S.Session(lib, write_cfg(tempfile.mkdtemp(), (256, 256, 256), name="hook_%d.cfg" % rank))
if world > 1:
    assert lib.SpiritB200_Slab_Setup(p.state, rank * 256, world * 256, -1) == 0
fill_random(p, seed=20006 + rank)
p.upload()
p.iterate_device(S.SOLVER_DEPONDT, 20)
for n in (1, 2, 5, 20, 100):
    reps = max(3, 100 // n)
    ms = sum(p.iterate_device(S.SOLVER_DEPONDT, n) for _ in range(reps)) / reps
    if rank == 0:
        print("block of %3d iterations: %.4f ms per block, %.4f ms per iteration" % (n, ms, ms / n), flush=True)
p.close()
