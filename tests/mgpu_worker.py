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
    failures = []
    # the second lattice is cut into several c-segments per slab: the fused kernel launches the two end segments first and
    # exchanges two halo planes per side while the interior segments run
    cases = [(24, 10, 12, bc_c, solver, T) for bc_c in (1, 0)
             for solver, T in (("Depondt", 0.0), ("Depondt", 5.0), ("Heun", 0.0), ("SIB", 0.0), ("RK4", 0.0), ("VP", 0.0),
                               ("VP_OSO", 0.0), ("LBFGS_OSO", 0.0), ("LBFGS_Atlas", 0.0))]
    cases += [(40, 20, 48, bc_c, solver, T) for bc_c in (1, 0) for solver, T in (("Depondt", 0.0), ("Depondt", 5.0), ("SIB", 5.0))]
    for Na, Nb, Nc, bc_c, solver, temperature in cases:
        if True:
            over = dict(boundary_conditions="1 0 %d" % bc_c, llg_temperature=temperature, llg_n_iterations_amortize=4)
            s_global = unit_random(Na * Nb * Nc, 21)
            c_begin, nc_local = slab.partition(Nc, world)[rank]
            path = os.path.join(tmp, "slab_%d.cfg" % rank)
            open(path, "w").write(cfgs.render("cubic256", n_basis_cells="%d %d %d" % (Na, Nb, nc_local), **over))
            p = S.Session(lib, path)
            assert lib.SpiritB200_Slab_Setup(p.state, c_begin, Nc, -1) == 0
            plane = Na * Nb
            p.set_spins(s_global[c_begin * plane:(c_begin + nc_local) * plane])
            variant = p.step_variant(S.SOLVERS[solver])
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
                # (the minimisers too: every dot product of the L-BFGS recursion is a sum over all slabs)
                tol = 1e-12 if solver == "VP" else (1e-10 if "OSO" in solver or "Atlas" in solver else 0.0)
                ok = dev <= tol and moved > 1e-4 and abs(e_slab - g.energy()) <= 1e-12 * abs(g.energy())
                print("%dx%dx%d bc_c=%d %-8s T=%g (step variant %d): max deviation %.3e, moved %.2e, E slab %.12e global %.12e %s" % (
                    Na, Nb, Nc, bc_c, solver, temperature, variant, dev, moved, e_slab, g.energy(), "OK" if ok else "FAIL"), flush=True)
                if not ok:
                    failures.append((bc_c, solver, temperature))
                g.close()
    failures += gneb_sharded(lib, rank, world, tmp)
    failures += ddi_distributed(lib, rank, world, tmp)
    dist.barrier()
    if rank == 0:
        print("MGPU_FAILURES %d" % len(failures), flush=True)
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


def ddi_distributed(lib, rank, world, tmp):
    """dipole-dipole FFT convolution on a slab decomposition (all-to-all transposes) vs the same lattice on one GPU"""
    failures = []
    # the second lattice has padded lengths 128 x 64 x 64: the power-of-two pass kernels with the per-rank kb blocks
    for Na, Nb, Nc, bc, solver in ((16, 8, 8, "0 0 0", "Depondt"), (16, 8, 8, "1 1 0", "SIB"), (16, 8, 8, "0 0 0", "VP"),
                                   (64, 32, 32, "0 0 0", "Depondt"), (64, 32, 32, "1 1 0", "SIB"),
                                   # ka pencils (padded a of at least 128, fast kernels): open, c periodic (un-padded c), long a
                                   (64, 32, 32, "0 0 0", "SIB"), (64, 16, 64, "0 0 1", "Heun"), (256, 16, 16, "0 1 0", "Depondt"),
                                   (64, 32, 32, "0 0 0", "VP")):
        plane = Na * Nb
        over = dict(boundary_conditions=bc, ddi_method="fft", ddi_n_periodic_images="2 2 0", llg_n_iterations_amortize=3)
        s_global = unit_random(Na * Nb * Nc, 33)
        c_begin, nc_local = slab.partition(Nc, world)[rank]
        path = os.path.join(tmp, "ddi_%d.cfg" % rank)
        open(path, "w").write(cfgs.render("default", n_basis_cells="%d %d %d" % (Na, Nb, nc_local), **over))
        p = S.Session(lib, path)
        assert lib.SpiritB200_Slab_Setup(p.state, c_begin, Nc, -1) == 0
        p.set_spins(s_global[c_begin * plane:(c_begin + nc_local) * plane])
        p.llg_start(S.SOLVERS[solver], n_iterations=6, n_iterations_log=6)
        mine, e_slab = p.spins().copy(), p.energy()
        p.close()
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        if rank == 0:
            gpath = os.path.join(tmp, "ddi_global.cfg")
            open(gpath, "w").write(cfgs.render("default", n_basis_cells="%d %d %d" % (Na, Nb, Nc), **over))
            g = S.Session(lib, gpath)
            g.set_spins(s_global)
            g.llg_start(S.SOLVERS[solver], n_iterations=6, n_iterations_log=6)
            ref = g.spins().copy()
            dev = np.abs(np.concatenate(parts) - ref).max()
            moved = np.abs(ref - s_global).max()
            ok = dev <= 1e-13 and moved > 1e-4 and abs(e_slab - g.energy()) <= 1e-11 * abs(g.energy())
            print("DDI distributed %dx%dx%d bc=%s %-8s: max deviation %.3e, moved %.2e, E slab %.12e global %.12e %s" % (
                Na, Nb, Nc, bc, solver, dev, moved, e_slab, g.energy(), "OK" if ok else "FAIL"), flush=True)
            if not ok:
                failures.append(("ddi", bc, solver))
            g.close()
    return failures


def gneb_sharded(lib, rank, world, tmp):
    """whole GNEB images per GPU: 8 images over `world` ranks vs the same chain on one GPU"""
    from tests.test_gneb_gpu import make_chain
    failures = []
    noi = 8
    path = os.path.join(tmp, "gneb_%d.cfg" % rank)
    open(path, "w").write(cfgs.render("solvers", n_basis_cells="12 10 1", boundary_conditions="1 0 0"))
    # every rank builds the full initial chain on the host (cheap) and keeps its shard
    full = S.Session(lib, path)
    make_chain(full, noi=noi)
    images0 = [full.spins(i).copy() for i in range(noi)]
    full.close()
    i_begin, n_local = slab.partition(noi, world)[rank]
    for solver, n, types in (("VP", 30, {3: S.GNEB_CLIMBING}), ("Depondt", 8, {2: S.GNEB_FALLING, 5: S.GNEB_CLIMBING}), ("Heun", 8, {}),
                             ("VP_OSO", 20, {3: S.GNEB_CLIMBING}), ("LBFGS_OSO", 12, {4: S.GNEB_CLIMBING})):
        p = S.Session(lib, path)
        p.set_anisotropy(0.25, (0, 0, 1))
        p.chain_set_length(n_local)
        for i in range(n_local):
            p.set_spins(images0[i_begin + i], idx_image=i)
        for g, t in types.items():
            if i_begin <= g < i_begin + n_local:
                p.gneb_set_image_type(t, g - i_begin)
        assert lib.SpiritB200_Chain_Shard_Setup(p.state, i_begin, noi) == 0
        p.gneb_start(S.SOLVERS[solver], single_shot=True)
        p.n_shot(n)
        mine = np.stack([p.spins(i).copy() for i in range(n_local)])
        rx, e = p.chain_rx_e()
        tq = p.chain_max_torque()
        p.stop()
        p.close()
        parts = [None] * world
        dist.all_gather_object(parts, (mine, rx, e, tq))
        if rank == 0:
            g = S.Session(lib, path)
            make_chain(g, noi=noi)
            for gi, t in types.items():
                g.gneb_set_image_type(t, gi)
            g.gneb_start(S.SOLVERS[solver], single_shot=True)
            g.n_shot(n)
            ref = np.stack([g.spins(i).copy() for i in range(noi)])
            rx_ref, e_ref = g.chain_rx_e()
            tq_ref = g.chain_max_torque()
            g.stop()
            g.close()
            dev = np.abs(np.concatenate([x[0] for x in parts]) - ref).max()
            drx = np.abs(np.concatenate([x[1] for x in parts]) - rx_ref).max()
            de = np.abs(np.concatenate([x[2] for x in parts]) - e_ref).max()
            dtq = max(abs(x[3] - tq_ref) for x in parts)
            oso = "OSO" in solver  # dot products over all images: summation order depends on the sharding
            ok = dev <= (1e-10 if oso else 1e-13) and drx <= (1e-9 if oso else 1e-12) and de <= (1e-8 if oso else 1e-10) and dtq <= (1e-8 if oso else 1e-12) * tq_ref
            print("GNEB sharded %-8s: spins %.3e Rx %.3e E %.3e torque %.3e %s" % (solver, dev, drx, de, dtq, "OK" if ok else "FAIL"), flush=True)
            if not ok:
                failures.append(("gneb", solver))
    return failures


if __name__ == "__main__":
    main()
