#!/usr/bin/env python
"""configs[3] alone (bench.config_c4) on N ranks: python -m torch.distributed.run --nproc-per-node N profiles/c4_only.py"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from spirit_b200 import capi, slab  # noqa: E402

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
lib = capi.load_product()
lib.SpiritB200_Set_Device(local)
d = None
if world > 1:
    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local))
    slab.init_comm(lib, dist, rank, world)
    d = dist
out = bench.config_c4(lib, tempfile.mkdtemp(), 6553.0, d, rank, world)
if rank == 0:
    print(json.dumps({k: out[k] for k in ("n_gpus", "iterations_per_s", "barrier_meV", "max_torque") if k in out}), flush=True)
if world > 1:
    dist.destroy_process_group()
