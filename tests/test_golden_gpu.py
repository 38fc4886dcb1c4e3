"""The CUDA path against the committed golden vectors (tests/golden/*.npz, generated from the unmodified reference
by tests/golden/make_golden.py). These need no oracle library at run time."""
import os

import numpy as np
import pytest

from spirit_b200 import session as S

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")

CASES = {
    "llg_solvers16.npz": ("solvers", {}, None),
    "llg_default_12x10x3.npz": ("default", {"n_basis_cells": "12 10 3"}, (0.75, 0.5)),
    "llg_cubic_8x6x5_periodic.npz": ("cubic256", {"n_basis_cells": "8 6 5", "llg_temperature": "0"}, None),
}


@pytest.mark.parametrize("name", list(CASES))
def test_llg_golden(cfg, product, name):
    preset, overrides, K = CASES[name]
    g = dict(np.load(os.path.join(GOLD, name)))

    def fresh():
        p = S.Session(product, cfg(preset, **overrides))
        if K:
            p.set_anisotropy(K[0], (0, 0, 1))
            p.set_cubic_anisotropy(K[1])
        return p

    p = fresh()
    grad, E = p.gradient_and_energy(g["spins0"])
    assert np.abs(grad - g["gradient"]).max() <= 1e-12 * np.abs(g["gradient"]).max()
    assert abs(E - g["energy"]) <= 1e-12 * abs(g["energy"])
    for term, (tot, per) in p.energy_contributions(g["spins0"], per_spin=True).items():
        ref = g["E_" + term.replace(" ", "_")]
        assert np.abs(per - ref).max() <= 1e-12 * np.abs(ref).max(), term
    p.close()
    for solver, n in (("Depondt", 5), ("Heun", 5), ("SIB", 5), ("RK4", 5), ("VP", 20)):
        p = fresh()
        p.llg_set(temperature=0.0, damping=0.3, dt=1e-3)
        p.set_spins(g["spins0"])
        p.llg_start(S.SOLVERS[solver], single_shot=True)
        p.n_shot(n)
        assert np.abs(p.spins() - g["spins_" + solver]).max() < 1e-10, solver
        assert abs(p.energy() - g["energy_" + solver]) <= 1e-11 * abs(g["energy_" + solver]), solver
        assert abs(p.max_torque() - g["torque_" + solver]) <= 1e-9 * g["torque_" + solver], solver
        p.stop()
        p.close()
