"""Parity of the tile-staged marching kernels (spirit_b200/csrc/device/sc6_tile.cuh: planes staged in shared memory by
bulk copies) with the reference CPU build, and with the register march they replace.

The variant needs rows that start on AoSoA-32 block boundaries (Na % 32 == 0); the lattices here are chosen to hit its
cases: tile = whole row (wrap inside the tile) / tile narrower than the row (halo blocks, wrapped or not), partial
tiles along b, several c-segments, 2 / 3 stage buffers, 2-D systems, open and mixed boundaries, DMI not parallel to
the bonds, rare terms (cubic anisotropy, tilted anisotropy axis) that take the GENERAL variant.
Tolerances are BASELINE.json's: single-step spin deviation < 1e-10.
"""
import numpy as np
import pytest

from spirit_b200 import session as S
from tests.test_parity_gpu import make_case, unit_random

pytestmark = pytest.mark.gpu

STEP_ATOL = 1e-10

# preset, overrides, extra, environment (tile shape overrides, read when the image's device tables are built)
TILE_CASES = [
    ("cubic256", {"n_basis_cells": "32 7 5"}, None, {}),
    ("cubic256", {"n_basis_cells": "64 10 6", "boundary_conditions": "1 0 1"}, None, {"SPIRIT_B200_SC6T_BY": "4"}),
    ("cubic256", {"n_basis_cells": "96 5 4", "boundary_conditions": "0 0 0"}, "aniso", {"SPIRIT_B200_SC6T_BX": "32"}),
    ("cubic256", {"n_basis_cells": "96 9 7", "boundary_conditions": "1 1 1"}, None,
     {"SPIRIT_B200_SC6T_BX": "32", "SPIRIT_B200_SC6T_BY": "2", "SPIRIT_B200_SC6_LC": "3", "SPIRIT_B200_SC6T_NS": "2"}),
    ("cubic256", {"n_basis_cells": "32 3 9", "boundary_conditions": "1 1 0", "dm_chirality": "2"}, None,
     {"SPIRIT_B200_SC6_LC": "4"}),
    ("cubic256", {"n_basis_cells": "512 4 3", "boundary_conditions": "1 1 1"}, None, {}),
    ("cubic256", {"n_basis_cells": "512 3 2", "boundary_conditions": "0 1 0"}, None, {"SPIRIT_B200_SC6T_NS": "2"}),
    ("solvers", {"n_basis_cells": "64 64 1"}, None, {}),
    ("solvers", {"n_basis_cells": "64 33 1", "boundary_conditions": "0 0 0"}, None, {"SPIRIT_B200_SC6T_BY": "8"}),
    ("cubic256", {"n_basis_cells": "256 6 40"}, None, {"SPIRIT_B200_SC6_LC": "16"}),
]


def set_env(monkeypatch, env, tiled=True):
    for k in ("SPIRIT_B200_SC6T_BX", "SPIRIT_B200_SC6T_BY", "SPIRIT_B200_SC6T_NS", "SPIRIT_B200_SC6_LC"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    monkeypatch.setenv("SPIRIT_B200_SC6_TILED", "1" if tiled else "0")
    monkeypatch.setenv("SPIRIT_B200_GENERIC_STENCIL", "0")


@pytest.mark.parametrize("solver", ["Depondt", "Heun", "SIB", "RK4"])
@pytest.mark.parametrize("preset,overrides,extra,env", TILE_CASES)
def test_tile_blocks_match_reference(cfg, product, oracle, monkeypatch, solver, preset, overrides, extra, env):
    """Simulation_LLG_Start over amortised blocks (the iterations inside a block run the marching kernels): spins, energy
    and effective field against the reference after 9 iterations"""
    set_env(monkeypatch, env)
    kw = dict(overrides, llg_n_iterations_amortize=4, llg_temperature=0)
    p, o = make_case(cfg, product, oracle, preset, kw, extra)
    s0 = unit_random(p.nos, 5)
    assert p.stencil_variant() == 7  # marching kernels, tile-staged for one- and two-window stages
    for x in (p, o):
        x.llg_set(temperature=0.0, damping=0.3, dt=1e-3)
        x.set_spins(s0)
        x.llg_start(S.SOLVERS[solver], n_iterations=9, n_iterations_log=9)
    assert np.abs(o.spins() - s0).max() > 1e-4
    assert np.abs(p.spins() - o.spins()).max() < STEP_ATOL
    assert abs(p.energy() - o.energy()) <= 1e-11 * max(1.0, abs(o.energy()))
    fo = o.effective_field()
    assert np.abs(p.effective_field() - fo).max() <= 1e-9 * max(np.abs(fo).max(), 1e-300)
    p.close()
    o.close()


@pytest.mark.parametrize("solver", ["Depondt", "SIB"])
@pytest.mark.parametrize("preset,overrides,extra,env", [TILE_CASES[1], TILE_CASES[3], TILE_CASES[5], TILE_CASES[8]])
def test_tile_thermal_equals_register_march(cfg, product, monkeypatch, solver, preset, overrides, extra, env):
    """T > 0: the tile-staged and the register march evaluate the same arithmetic with the same counter-based noise
    (Philox keyed by site and plane), so 20 iterations from the same state agree to rounding of the fused operations
    -- and certainly far below the noise amplitude, which a wrong site index in the counter would not."""
    kw = dict(overrides, llg_n_iterations_amortize=5, llg_temperature=25)
    res = []
    for tiled in (True, False):
        set_env(monkeypatch, env, tiled)
        p = S.Session(product, cfg(preset, **kw))
        p.llg_set(temperature=25.0, damping=0.3, dt=1e-3)
        assert p.stencil_variant() == (7 if tiled else 1)
        p.set_spins(unit_random(p.nos, 3))
        p.llg_start(S.SOLVERS[solver], n_iterations=20, n_iterations_log=20)
        res.append(p.spins().copy())
        p.close()
    assert np.abs(res[0] - res[1]).max() < 1e-12
    monkeypatch.setenv("SPIRIT_B200_SC6_TILED", "1")


def test_tile_direct_minimisation_matches_reference(cfg, product, oracle, monkeypatch):
    """llg_direct_minimization (MODE = minimise) through the tile-staged kernels"""
    set_env(monkeypatch, {})
    p, o = make_case(cfg, product, oracle, "solvers", {"n_basis_cells": "32 16 1", "llg_n_iterations_amortize": 5}, None)
    for x in (p, o):
        x.plus_z()
        x.skyrmion(5.0, phase=-90.0)
        x.llg_set(direct_minimization=True)
        x.llg_start(S.SOLVER_DEPONDT, n_iterations=50, n_iterations_log=50)
    assert np.abs(p.spins() - o.spins()).max() < 1e-9
    p.close()
    o.close()
