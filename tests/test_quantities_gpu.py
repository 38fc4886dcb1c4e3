"""Topological charge (reference Vectormath.cpp:504-631, Quantities.cpp:62-133) on planar lattices with
one basis atom: total charge and the charge of every triangle against the reference. The reference returns floats."""
import ctypes

import numpy as np
import pytest

from spirit_b200 import session as S
from tests.test_parity_gpu import unit_random

pytestmark = pytest.mark.gpu

CASES = [
    ("solvers", {"n_basis_cells": "12 9 1", "boundary_conditions": "0 0 0"}),
    ("solvers", {"n_basis_cells": "12 9 1", "boundary_conditions": "1 1 0"}),
    ("solvers", {"n_basis_cells": "16 16 1", "boundary_conditions": "1 0 0"}),
    ("cubic256", {"n_basis_cells": "12 10 1", "bravais_lattice": "hex2d", "boundary_conditions": "0 0 0"}),
    ("cubic256", {"n_basis_cells": "12 10 1", "bravais_lattice": "hex2d", "boundary_conditions": "1 1 0"}),
]


def density(x):
    lib, st = x.lib, x.state
    n = lib.Quantity_Get_Topological_Charge_Density(st, None, None, -1, -1)
    q, tri = (ctypes.c_float * n)(), (ctypes.c_int * (3 * n))()
    assert lib.Quantity_Get_Topological_Charge_Density(st, q, tri, -1, -1) == n
    t = np.array(tri).reshape(-1, 3)
    return {tuple(sorted(map(int, row))): float(v) for row, v in zip(t, np.array(q))}, n


@pytest.mark.parametrize("preset,over", CASES)
def test_topological_charge_vs_reference(cfg, product, oracle, preset, over):
    path = cfg(preset, **over)
    p, o = S.Session(product, path), S.Session(oracle, path)
    for k, make in enumerate((lambda x: (x.plus_z(), x.skyrmion(3.0, phase=-90.0)), lambda x: x.set_spins(unit_random(x.nos, 21)),
                              lambda x: (x.minus_z(), x.skyrmion(2.5, order=2, phase=30.0, up_down=True)))):
        make(p), make(o)
        qp = p.lib.Quantity_Get_Topological_Charge(p.state, -1, -1)
        qo = o.lib.Quantity_Get_Topological_Charge(o.state, -1, -1)
        assert abs(qp - qo) <= 2e-6 * max(1.0, abs(qo)), (k, qp, qo)
        dp, n_p = density(p)
        do, n_o = density(o)
        assert n_p == n_o and set(dp) == set(do)  # the same triangles
        assert max(abs(dp[t] - do[t]) for t in do) <= 1e-6
    p.close(), o.close()


def test_topological_charge_of_a_skyrmion_on_a_periodic_film_is_an_integer(cfg, product):
    p = S.Session(product, cfg("solvers", n_basis_cells="64 64 1", boundary_conditions="1 1 0"))
    p.plus_z()
    p.skyrmion(8.0, phase=-90.0)
    assert abs(p.lib.Quantity_Get_Topological_Charge(p.state, -1, -1) + 1.0) < 1e-6
    p.plus_z()
    assert p.lib.Quantity_Get_Topological_Charge(p.state, -1, -1) == 0.0
    p.close()


BASIS_CASES = [
    ("hex2d", ["basis", "2", "0 0 0", "0.333 0.333 0.0"], "1 1 0"),
    ("hex2d", ["basis", "2", "0 0 0", "0.333 0.333 0.0"], "0 0 0"),
    ("sc", ["basis", "3", "0 0 0", "0.5 0.2 0.0", "0.2 0.6 0"], "1 0 0"),
]


@pytest.mark.parametrize("lattice,block,bc", BASIS_CASES)
def test_topological_charge_with_basis_vs_reference(tmp_path, product, oracle, lattice, block, bc):
    """several basis atoms: the triangle table of the cell (Delaunay, host) through k_topological_charge_table"""
    from tests import cfgs
    path = tmp_path / "b.cfg"
    path.write_text(cfgs.render("cubic256", block=block, n_basis_cells="10 8 1", bravais_lattice=lattice, boundary_conditions=bc))
    p, o = S.Session(product, str(path)), S.Session(oracle, str(path))
    for k, make in enumerate((lambda x: (x.plus_z(), x.skyrmion(2.5, phase=-90.0)), lambda x: x.set_spins(unit_random(x.nos, 5)))):
        make(p), make(o)
        qp = p.lib.Quantity_Get_Topological_Charge(p.state, -1, -1)
        qo = o.lib.Quantity_Get_Topological_Charge(o.state, -1, -1)
        assert abs(qp - qo) <= 2e-6 * max(1.0, abs(qo)), (k, qp, qo)
        dp, n_p = density(p)
        do, n_o = density(o)
        assert n_p == n_o and set(dp) == set(do)
        assert max(abs(dp[t] - do[t]) for t in do) <= 1e-6
    p.close(), o.close()


def test_the_reference_python_test_of_the_topological_charge(tmp_path, product):
    """core/python/test/quantities.py:35-41 on its own lattice (api.cfg: honeycomb, 50 x 50 cells, periodic): Q = -1"""
    from tests import cfgs
    path = tmp_path / "api.cfg"
    path.write_text(cfgs.render("cubic256", block=["basis", "2", "0 0 0", "0.333 0.333 0.0"], n_basis_cells="50 50 1",
                                bravais_lattice="hex2d", boundary_conditions="1 1 0"))
    p = S.Session(product, str(path))
    p.plus_z()
    p.skyrmion(5.0, pos=(1.5, 0, 0))
    q = p.lib.Quantity_Get_Topological_Charge(p.state, -1, -1)
    d, n = density(p)
    assert abs(q + 1.0) < 1e-6 and abs(q - sum(d.values())) < 1e-4
    p.close()
