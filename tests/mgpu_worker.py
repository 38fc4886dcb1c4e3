"""Worker of tests/test_multigpu.py (one process per GPU, launched by torch.distributed.run): an Nc-plane lattice cut
into world slabs must evolve bit-identically to the same lattice on one GPU."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from spirit_b200 import capi, session as S, slab  # noqa: E402
from tests import cfgs  # noqa: E402
from tests.test_parity_gpu import unit_random  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", rank))
    lib = capi.load_product()
    assert lib.SpiritB200_Device_Count() >= world
    lib.SpiritB200_Set_Device(local)
    slab.init_comm(lib, dist, rank, world)
    tmp = tempfile.mkdtemp()
    Na, Nb, Nc = 24, 10, 12
    failures = []
    for bc_c in (1, 0):
        for solver, temperature in (("Depondt", 0.0), ("Depondt", 5.0), ("Heun", 0.0), ("SIB", 0.0), ("RK4", 0.0), ("VP", 0.0)):
            over = dict(boundary_conditions="1 0 %d" % bc_c, llg_temperature=temperature, llg_n_iterations_amortize=4)
            s_global = unit_random(Na * Nb * Nc, 21)
            c_begin, nc_local = slab.partition(Nc, world)[rank]
            path = os.path.join(tmp, "slab_%d.cfg" % rank)
            open(path, "w").write(cfgs.render("cubic256", n_basis_cells="%d %d %d" % (Na, Nb, nc_local), **over))
            p = S.Session(lib, path)
            assert lib.SpiritB200_Slab_Setup(p.state, c_begin, Nc, -1) == 0
            plane = Na * Nb
            p.set_spins(s_global[c_begin * plane:(c_begin + nc_local) * plane])
            p.llg_start(S.SOLVERS[solver], n_iterations=8, n_iterations_log=8)
            mine = p.spins().copy()
            e_slab = p.energy()
            p.close()
            parts = [None] * world
            dist.all_gather_object(parts, mine)
            if rank == 0:
                gpath = os.path.join(tmp, "global.cfg")
                open(gpath, "w").write(cfgs.render("cubic256", n_basis_cells="%d %d %d" % (Na, Nb, Nc), **over))
                g = S.Session(lib, gpath)
                g.set_spins(s_global)
                g.llg_start(S.SOLVERS[solver], n_iterations=8, n_iterations_log=8)
                ref = g.spins().copy()
                dev = np.abs(np.concatenate(parts) - ref).max()
                moved = np.abs(ref - s_global).max()
                # VP couples all sites through two global sums whose summation order depends on the decomposition
                tol = 1e-12 if solver == "VP" else 0.0
                ok = dev <= tol and moved > 1e-4 and abs(e_slab - g.energy()) <= 1e-12 * abs(g.energy())
                print("bc_c=%d %-8s T=%g: max deviation %.3e, moved %.2e, E slab %.12e global %.12e %s" % (
                    bc_c, solver, temperature, dev, moved, e_slab, g.energy(), "OK" if ok else "FAIL"), flush=True)
                if not ok:
                    failures.append((bc_c, solver, temperature))
                g.close()
    dist.barrier()
    if rank == 0:
        print("MGPU_FAILURES %d" % len(failures), flush=True)
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
