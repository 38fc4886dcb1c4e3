"""Pins the oracle (CPU, no GPU): the compiled reference (oracle/_ref) and the NumPy restatement (oracle/restatement.py)
against the committed golden vectors, against each other, and against the reference's own known-answer tests."""
import os

import numpy as np
import pytest

from oracle import restatement as R
from spirit_b200 import session as S

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold(name):
    return dict(np.load(os.path.join(GOLD, name)))


# dt and damping reached the reference through its float setters (Parameters_LLG_Set_Time_Step(float), SURVEY.md 8c
# hazard 5): the restatement must use the same narrowed values
F32 = dict(dt=float(np.float32(1e-3)), alpha=float(np.float32(0.3)))
MODELS = {
    "llg_solvers16.npz": R.Model((16, 16, 1), (1, 1, 0), **F32),
    "llg_default_12x10x3.npz": R.Model((12, 10, 3), (1, 1, 0), K=0.75, K4=0.5, **F32),
    "llg_cubic_8x6x5_periodic.npz": R.Model((8, 6, 5), (1, 1, 1), B=0.0, K=1.0, **F32),
}


@pytest.mark.parametrize("name", list(MODELS))
def test_restatement_matches_golden(name):
    g, m = gold(name), MODELS[name]
    s0 = g["spins0"]
    grad, E = m.gradient_and_energy(s0)
    assert np.abs(grad - g["gradient"]).max() <= 1e-12 * np.abs(g["gradient"]).max()
    assert abs(E - g["energy"]) <= 1e-12 * abs(g["energy"])
    for solver, f in (("Depondt", m.depondt), ("Heun", m.heun), ("SIB", m.sib), ("RK4", m.rk4)):
        s = s0.copy()
        for _ in range(5):
            s = f(s)
        assert np.abs(s - g["spins_" + solver]).max() < 1e-12, solver
    assert np.abs(m.vp_single_shots(s0.copy(), 20) - g["spins_VP"]).max() < 1e-12


def test_restatement_gneb_matches_golden():
    g = gold("gneb_7x10x10.npz")
    m = R.Model((10, 10, 1), (1, 1, 0), K=0.25)
    types = [R.NORMAL] * 7
    types[3] = R.CLIMBING
    imgs, E, Rx = R.gneb_vp_single_shots(m, list(g["images0"]), types, 1.0, 60)
    assert np.abs(np.stack(imgs) - g["images"]).max() < 1e-10


@pytest.mark.parametrize("name,preset,overrides,K", [
    ("llg_solvers16.npz", "solvers", {}, None),
    ("llg_default_12x10x3.npz", "default", {"n_basis_cells": "12 10 3"}, (0.75, 0.5)),
])
def test_compiled_reference_matches_golden(cfg, oracle, name, preset, overrides, K):
    """The golden files are reproducible from the reference build in this tree"""
    g = gold(name)
    o = S.Session(oracle, cfg(preset, **overrides))
    if K:
        o.set_anisotropy(K[0], (0, 0, 1))
        o.set_cubic_anisotropy(K[1])
    grad, E = o.gradient_and_energy(g["spins0"])
    assert np.abs(grad - g["gradient"]).max() <= 1e-14 * np.abs(g["gradient"]).max()
    assert abs(E - g["energy"]) <= 1e-12 * abs(g["energy"])  # summation order depends on the thread count
    o.llg_set(temperature=0.0, damping=0.3, dt=1e-3)
    o.set_spins(g["spins0"])
    o.llg_start(S.SOLVER_DEPONDT, single_shot=True)
    o.n_shot(5)
    assert np.abs(o.spins() - g["spins_Depondt"]).max() < 1e-14
    o.stop()
    o.close()


@pytest.mark.parametrize("solver", ["Heun", "Depondt", "SIB", "RK4"])
def test_reference_larmor_known_answer(cfg, oracle, solver):
    """core/test/test_physics.cpp:20-88 on the oracle: s_z = tanh(alpha dtg t B), s_x = cos(dtg t B) sqrt(1 - s_z^2)"""
    o = S.Session(oracle, cfg("larmor"))
    damping, dt, B = 0.3, 0.001, 1.0
    o.llg_set(damping=damping, dt=dt)
    o.domain((1.0, 0.0, 0.0))
    o.llg_start(S.SOLVERS[solver], single_shot=True)
    dtg = dt * oracle.Constants_gamma() / (1.0 + damping ** 2)
    for i in range(100):
        o.single_shot()
        s = o.spins()[0]
        sz = np.tanh(damping * dtg * (i + 1) * B)
        assert abs(s[0] - np.cos(dtg * (i + 1) * B) * np.sqrt(1 - sz ** 2)) < 1e-6
        assert abs(s[2] - sz) < 1e-6
    o.stop()
    o.close()


def test_reference_gradient_is_finite_difference_of_energy(cfg, oracle):
    """core/test/test_physics.cpp:90-142 (fd_pairs.cfg): analytic gradient == central finite difference of the energy"""
    o = S.Session(oracle, cfg("fd_pairs"))
    rng = np.random.default_rng(3)
    s = rng.normal(size=(o.nos, 3))
    s /= np.linalg.norm(s, axis=1, keepdims=True)
    g, _ = o.gradient_and_energy(s)
    delta = 1e-4
    for i in range(o.nos):
        for d in range(3):
            sp, sm = s.copy(), s.copy()
            sp[i, d] += delta
            sm[i, d] -= delta
            fd = (o.gradient_and_energy(sp)[1] - o.gradient_and_energy(sm)[1]) / (2 * delta)
            assert abs(fd - g[i, d]) < 1e-7 * max(1.0, abs(g[i, d]))
    o.close()


def test_reference_skyrmion_relaxation_golden_value(cfg, oracle):
    """core/test/test_solvers.cpp:44-45: E = -5849.69140625, M_z = 2 * 0.79977 (float precision golden values)"""
    o = S.Session(oracle, cfg("solvers"))
    o.plus_z()
    o.skyrmion(5.0, phase=-90.0)
    o.llg_set(direct_minimization=True)
    o.llg_start(S.SOLVER_VP)
    o.update_data()
    assert abs(o.energy() - (-5849.69140625)) < 1e-3
    assert abs(o.magnetization()[2] - 2 * 0.79977) < 1e-4
    o.close()


def _oracle_density(o):
    import ctypes
    n = o.lib.Quantity_Get_Topological_Charge_Density(o.state, None, None, -1, -1)
    q, tri = (ctypes.c_float * n)(), (ctypes.c_int * (3 * n))()
    o.lib.Quantity_Get_Topological_Charge_Density(o.state, q, tri, -1, -1)
    t = np.array(tri).reshape(-1, 3)
    return {tuple(sorted(map(int, row))): float(v) for row, v in zip(t, np.array(q))}


@pytest.mark.parametrize("lattice,bc", [("sc", "0 0 0"), ("sc", "1 1 0"), ("sc", "1 0 0"), ("hex2d", "0 0 0"), ("hex2d", "1 1 0")])
def test_restatement_topological_charge_matches_compiled_reference(cfg, oracle, lattice, bc):
    """pins oracle/restatement.py::topological_charge (cut of the cell, orientation, boundary rule) to the reference"""

    o = S.Session(oracle, cfg("cubic256", n_basis_cells="9 7 1", bravais_lattice=lattice, boundary_conditions=bc))
    periodic = [int(v) for v in bc.split()][:2]
    tb = (0.5, 0.5 * np.sqrt(3.0)) if lattice == "hex2d" else (0.0, 1.0)
    for make in (lambda x: (x.plus_z(), x.skyrmion(2.5, phase=-90.0)), lambda x: x.random()):
        make(o)
        total, per_triangle = R.topological_charge(o.spins(), (9, 7), periodic, tb=tb)
        ref = _oracle_density(o)
        assert set(ref) == set(per_triangle)
        assert max(abs(per_triangle[t] - ref[t]) for t in ref) < 1e-6  # the API returns floats
        assert abs(total - o.lib.Quantity_Get_Topological_Charge(o.state, -1, -1)) < 2e-6 * max(1.0, abs(total))
    o.close()


def test_restatement_direct_dipolar_sum_matches_compiled_reference(cfg, oracle):
    """pins oracle/restatement.py::ddi_gradient_direct to the reference's direct sum AND to its FFT convolution"""

    over = dict(n_basis_cells="5 4 3", boundary_conditions="0 0 0", jij="0", dij="0", n_shells_dmi="0", anisotropy_magnitude="0",
                external_field_magnitude="0", llg_temperature="0")
    for method in ({"ddi_method": "cutoff", "ddi_radius": "-1"}, {"ddi_method": "fft"}):
        o = S.Session(oracle, cfg("cubic256", **dict(over, **method)))
        s = o.spins().copy()
        g_ref, _ = o.gradient_and_energy(s)
        g = R.ddi_gradient_direct(s, (5, 4, 3), mu_s=2.0)
        assert np.abs(g - g_ref).max() <= 1e-12 * np.abs(g_ref).max()
        o.close()


@pytest.mark.parametrize("lattice,block,basis,bc", [
    ("hex2d", ["basis", "2", "0 0 0", "0.333 0.333 0.0"], [(0, 0), (0.333, 0.333)], "1 1 0"),
    ("hex2d", ["basis", "2", "0 0 0", "0.333 0.333 0.0"], [(0, 0), (0.333, 0.333)], "0 0 0"),
    ("sc", ["basis", "3", "0 0 0", "0.5 0.2 0.0", "0.2 0.6 0"], [(0, 0), (0.5, 0.2), (0.2, 0.6)], "1 0 0")])
def test_restatement_topological_charge_with_basis_matches_compiled_reference(tmp_path, oracle, lattice, block, basis, bc):
    from tests import cfgs
    path = tmp_path / "b.cfg"
    path.write_text(cfgs.render("cubic256", block=block, n_basis_cells="8 6 1", bravais_lattice=lattice, boundary_conditions=bc))
    o = S.Session(oracle, str(path))
    periodic = [int(v) for v in bc.split()][:2]
    tb = (0.5, 0.5 * np.sqrt(3.0)) if lattice == "hex2d" else (0.0, 1.0)
    for make in (lambda x: (x.plus_z(), x.skyrmion(2.0, phase=-90.0)), lambda x: x.random()):
        make(o)
        total, per_triangle = R.topological_charge_basis(o.spins(), (8, 6), periodic, basis, (1.0, 0.0), tb)
        ref = _oracle_density(o)
        assert set(ref) == set(per_triangle)
        assert max(abs(per_triangle[t] - ref[t]) for t in ref) < 1e-6
        assert abs(total - o.lib.Quantity_Get_Topological_Charge(o.state, -1, -1)) < 2e-6 * max(1.0, abs(total))
    o.close()
