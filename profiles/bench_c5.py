#!/usr/bin/env python
"""BASELINE configs[4]: simple cubic N^3 (default 512^3), exchange + DMI + DDI via the distributed FFT, LLG SIB, slab-decomposed
over the ranks of one box (launch with torch.distributed.run, one rank per GPU). Secondary measurement, not bench.py's line.
usage: python -m torch.distributed.run --nproc-per-node G profiles/bench_c5.py [--edge 512] [--steps 20]"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import fill_random  # noqa: E402
from spirit_b200 import capi, session as S, slab  # noqa: E402
from tests import cfgs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--edge", type=int, default=512)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--solver", default="SIB")
args = ap.parse_args()

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
lib = capi.load_product()
lib.SpiritB200_Set_Device(local)
if world > 1:
    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local))
    slab.init_comm(lib, dist, rank, world)
N = args.edge
c_begin, ncl = slab.partition(N, world)[rank]
tmp = tempfile.mkdtemp()
path = os.path.join(tmp, "c5_%d.cfg" % rank)
open(path, "w").write(cfgs.render("cubic256", n_basis_cells="%d %d %d" % (N, N, ncl), boundary_conditions="0 0 0", ddi_method="fft",
                                  ddi_n_periodic_images="0 0 0", anisotropy_magnitude=0, external_field_magnitude=25, llg_temperature=0))
p = S.Session(lib, path)
if world > 1:
    assert lib.SpiritB200_Slab_Setup(p.state, c_begin, N, -1) == 0
fill_random(p, seed=7 + rank)
t0 = time.time()
p.upload()
t_setup = time.time() - t0
solver = S.SOLVERS[args.solver]
p.iterate_device(solver, 3)
if world > 1:
    dist.barrier()
ms = p.iterate_device(solver, args.steps)
if world > 1:
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
if rank == 0:
    nos = N ** 3
    rate = nos * args.steps / ms * 1e3
    model = 120 + 2 * 1008  # stencil step + two pruned dipolar convolutions
    print(json.dumps({"config": "C5 %d^3 sc, exchange+DMI+DDI (distributed FFT), LLG %s, %d GPU(s), slabs of %d planes" % (N, args.solver, world, ncl),
                      "ms_per_iteration": ms / args.steps, "spin_steps_per_s": rate, "per_gpu_model_GBps": rate * model / 1e9 / world,
                      "ddi_plan_setup_s": t_setup}), flush=True)
p.close()
if world > 1:
    dist.destroy_process_group()
