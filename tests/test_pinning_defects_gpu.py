"""Pinned sites and defects on the device, against the reference built with -DSPIRIT_ENABLE_PINNING -DSPIRIT_ENABLE_DEFECTS
(oracle/_ref/libSpirit_ref_pd.so, `make -C oracle pd`): check_atom_type / idx_from_pair (Vectormath.hpp:406-528) in every
term of the Hamiltonian, the pinning mask on force and virtual force (Method_LLG.cpp:122-124, 222-224) and on the total
GNEB force (Method_GNEB.cpp:251-254, 386-389). Same tolerances as tests/test_parity_gpu.py."""
import numpy as np
import pytest

from spirit_b200 import session as S
from tests.test_parity_gpu import GRAD_RTOL, STEP_ATOL, unit_random

pytestmark = pytest.mark.gpu

PIN = ["pin_na_left 2", "pin_nb 1", "pinning_cell", "0 0.6 0.8", "n_pinned 2", "0 5 3 0  1 0 0", "0 6 4 1  0 0 -1"]
DEFECTS = ["n_defects 4", "0 4 4 0 -1", "0 7 2 1 -1", "0 3 5 0 2", "0 0 0 0 -1"]


def both(product, oracle_pd, path):
    return S.Session(product, path), S.Session(oracle_pd, path)


LATTICES = [
    ("default", {"n_basis_cells": "10 8 2"}),                                      # pair table, B = 25 T, open in c
    ("cubic256", {"n_basis_cells": "12 9 4", "external_field_magnitude": "7"}),     # neighbour shells + anisotropy, periodic
    ("cubic256", {"n_basis_cells": "11 8 3", "boundary_conditions": "1 0 1", "n_shells_exchange": "2", "jij": "10.0 -2.5",
                  "external_field_magnitude": "3"}),
]


@pytest.mark.parametrize("preset,overrides", LATTICES)
@pytest.mark.parametrize("block", [DEFECTS, PIN, PIN + DEFECTS], ids=["defects", "pinning", "both"])
def test_gradient_energy_and_terms(cfg, product, oracle_pd, preset, overrides, block):
    p, o = both(product, oracle_pd, cfg(preset, block=block, **overrides))
    assert p.stencil_variant() == 0  # the nearest-neighbour kernels step aside
    np.testing.assert_array_equal(p.atom_types(), o.atom_types())
    s = unit_random(p.nos, 4)
    gp, ep = p.gradient_and_energy(s)
    go, eo = o.gradient_and_energy(s)
    assert np.abs(gp - go).max() <= GRAD_RTOL * np.abs(go).max()
    if block is not PIN:
        assert np.all(gp[o.atom_types() < 0] == 0.0)
    cp, co = p.energy_contributions(s, per_spin=True), o.energy_contributions(s, per_spin=True)
    assert list(cp.keys()) == list(co.keys())
    abs_sum = sum(np.abs(v[1]).sum() for v in co.values())
    assert abs(ep - eo) <= 1e-12 * max(abs_sum, abs(eo))
    for name in co:
        assert np.abs(cp[name][1] - co[name][1]).max() <= 1e-12 * max(np.abs(co[name][1]).max(), 1e-300), name
    mp, mo = p.magnetization(), o.magnetization()
    assert np.abs(np.array(mp) - np.array(mo)).max() <= 1e-6
    for x in (p, o):
        x.close()


def test_defects_with_dipolar_interaction(cfg, product, oracle_pd):
    """defect sites carry no moment: they do not enter the dipolar convolution and feel no dipolar field"""
    block = ["n_defects 3", "0 2 2 1 -1", "1 1 3 2 -1", "0 3 0 4 1"]
    from tests import cfgs
    base = cfgs.PRESETS["ddi"]["_block"]
    p, o = both(product, oracle_pd, cfg("ddi", block=base + block, ddi_n_periodic_images="2 2 2"))
    s = unit_random(p.nos, 9)
    gp, ep = p.gradient_and_energy(s)
    go, eo = o.gradient_and_energy(s)
    assert np.abs(gp - go).max() <= 1e-11 * np.abs(go).max()
    assert abs(ep - eo) <= 1e-11 * abs(eo)
    cp, co = p.energy_contributions(s, per_spin=True), o.energy_contributions(s, per_spin=True)
    for name in co:
        assert np.abs(cp[name][1] - co[name][1]).max() <= 1e-11 * max(np.abs(co[name][1]).max(), 1e-300), name
    for x in (p, o):
        x.close()


@pytest.mark.parametrize("solver", ["Depondt", "Heun", "SIB", "RK4", "VP"])
@pytest.mark.parametrize("preset,overrides", LATTICES[:2])
def test_steps_with_pinned_sites(cfg, product, oracle_pd, solver, preset, overrides):
    """dynamics with pinned boundary layers and single sites: single shots and an amortised block"""
    p, o = both(product, oracle_pd, cfg(preset, block=PIN, llg_n_iterations_amortize=4, **overrides))
    s0 = unit_random(p.nos, 11)
    for x in (p, o):
        x.llg_set(temperature=0.0, damping=0.3, dt=1e-3)
        x.set_spins(s0)
        x.llg_start(S.SOLVERS[solver], single_shot=True)
        x.n_shot(20 if solver == "VP" else 5)
    sp, so = p.spins(), o.spins()
    moved = np.abs(so - s0).max(axis=1)
    assert moved.max() > 1e-4 and (moved == 0).sum() >= 2 * 8 * 2  # the pinned sites did not move
    assert np.abs(sp - so).max() < STEP_ATOL
    assert np.abs(sp[moved == 0] - s0[moved == 0]).max() < 4e-16  # (solvers that renormalise may touch the last bit)
    assert abs(p.energy() - o.energy()) <= 1e-11 * max(1.0, abs(o.energy()))
    assert abs(p.max_torque() - o.max_torque()) <= 1e-9 * max(1e-30, o.max_torque())
    fo = o.effective_field()
    assert np.abs(p.effective_field() - fo).max() <= 1e-9 * max(np.abs(fo).max(), 1e-300)
    for x in (p, o):
        x.stop()
    if solver != "VP":
        for x in (p, o):
            x.set_spins(s0)
            x.llg_start(S.SOLVERS[solver], n_iterations=12, n_iterations_log=12)
        assert np.abs(p.spins() - o.spins()).max() < STEP_ATOL
        assert abs(p.energy() - o.energy()) <= 1e-11 * max(1.0, abs(o.energy()))
    for x in (p, o):
        x.close()


@pytest.mark.parametrize("solver", ["Depondt", "SIB", "VP", "LBFGS_OSO", "VP_OSO"])
def test_minimisation_with_defects_and_pinning(cfg, product, oracle_pd, solver):
    """vacancies (and pinned sites) under the solvers that do not divide by mu_s: direct minimisation, VP and the OSO family"""
    preset, overrides = LATTICES[0]
    p, o = both(product, oracle_pd, cfg(preset, block=PIN + DEFECTS, **overrides))
    s0 = unit_random(p.nos, 13)
    for x in (p, o):
        x.llg_set(temperature=0.0, damping=0.3, dt=1e-3, direct_minimization=True)
        x.set_spins(s0)
        x.llg_start(S.SOLVERS[solver], single_shot=True)
        x.n_shot(8)
    so = o.spins()
    assert np.isfinite(so).all() and np.abs(so - s0).max() > 1e-4
    assert np.abs(p.spins() - so).max() < (1e-9 if "OSO" in solver else STEP_ATOL)
    assert abs(p.energy() - o.energy()) <= 1e-10 * max(1.0, abs(o.energy()))
    for x in (p, o):
        x.stop()
        x.close()


def test_pinning_set_at_run_time_and_thermal_noise(cfg, product, oracle_pd):
    """Configuration_Set_Pinned between two runs: the device tables follow; at T > 0 the pinned sites stay put"""
    p, o = both(product, oracle_pd, cfg("cubic256", n_basis_cells="10 10 4", llg_temperature=0))
    s0 = unit_random(p.nos, 17)
    for x in (p, o):
        x.llg_set(temperature=0.0, damping=0.3, dt=1e-3)
        x.set_spins(s0)
        x.llg_start(S.SOLVER_DEPONDT, n_iterations=3, n_iterations_log=3)
        x.set_pinned(True, pos=(1, 0, 0), cylindrical=3.2)
        x.set_atom_type(-1, pos=(-3, -3, 1), spherical=1.1)
        x.llg_set(direct_minimization=True)
        x.llg_start(S.SOLVER_DEPONDT, n_iterations=3, n_iterations_log=3)
    assert p.stencil_variant() == 0
    assert np.abs(p.spins() - o.spins()).max() < STEP_ATOL
    s1 = p.spins().copy()
    p.llg_set(temperature=30.0, direct_minimization=False)
    p.llg_start(S.SOLVER_DEPONDT, n_iterations=5, n_iterations_log=5)
    moved = np.abs(p.spins() - s1).max(axis=1)
    assert np.isfinite(p.spins()).all() and moved.max() > 1e-4
    frozen = moved == 0
    assert frozen.sum() >= 20 and frozen.sum() < p.nos // 2
    for x in (p, o):
        x.close()


def test_gneb_with_pinned_sites(cfg, product, oracle_pd):
    """the pinning mask on the total force of a chain (Method_GNEB.cpp:251-254): five images, VP"""
    p, o = both(product, oracle_pd, cfg("solvers", block=["pin_na 2", "pinning_cell", "0 0 1"]))
    for x in (p, o):
        x.plus_z()
        x.skyrmion(4.0, phase=-90.0)
        x.chain_set_length(5)
        x.jump_to_image(4)
        x.plus_z()
        x.transition_homogeneous(0, 4)
        x.gneb_start(S.SOLVER_VP, single_shot=True)
        x.n_shot(10)
    for i in range(5):
        assert np.abs(p.spins(i) - o.spins(i)).max() < 1e-9, i
    rx_p, e_p = p.chain_rx_e()
    rx_o, e_o = o.chain_rx_e()
    assert np.abs(np.array(e_p) - np.array(e_o)).max() <= 1e-9 * np.abs(e_o).max()
    for x in (p, o):
        x.stop_all()
        x.close()
