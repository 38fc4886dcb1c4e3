"""Dipole-dipole interaction (zero-padded FFT convolution with hand-written FFT passes) against the reference's FFT path
and its O(N^2) direct sum (core/test/test_physics.cpp:144-181 checks exactly this pair on physics_ddi.cfg)."""
import numpy as np
import pytest

from spirit_b200 import session as S
from tests.test_parity_gpu import unit_random

pytestmark = pytest.mark.gpu

# (preset, overrides). Shapes keep Na >= Nb >= Nc: the reference's kissFFT path is wrong otherwise (SURVEY.md 8c hazard 1)
DDI_CASES = [
    ("ddi", {}),                                                                              # 5x5x5, 2-atom basis, skewed cell, BC 1 0 0, 4 images
    ("ddi", {"boundary_conditions": "0 0 0"}),
    ("ddi", {"boundary_conditions": "1 1 1", "ddi_n_periodic_images": "2 2 2", "n_basis_cells": "4 4 4"}),
    ("ddi", {"boundary_conditions": "1 1 0", "ddi_pb_zero_padding": "0", "ddi_n_periodic_images": "3 3 0", "n_basis_cells": "6 4 2"}),
    ("default", {"n_basis_cells": "8 6 4", "boundary_conditions": "0 0 0", "ddi_method": "fft"}),
    ("default", {"n_basis_cells": "16 16 1", "boundary_conditions": "1 1 0", "ddi_method": "fft", "ddi_n_periodic_images": "2 2 0"}),
    ("default", {"n_basis_cells": "12 10 1", "boundary_conditions": "0 0 0", "ddi_method": "fft"}),
    ("default", {"n_basis_cells": "10 1 1", "boundary_conditions": "0 0 0", "ddi_method": "fft"}),
    ("cubic256", {"n_basis_cells": "18 14 6", "boundary_conditions": "0 0 0", "ddi_method": "fft", "llg_temperature": "0"}),
    # padded lengths that are powers of two >= 64: the in-register radix-8 stages (64 = 8.8, 128 = 8.8.2, 256 = 8.8.4, 512, 1024)
    ("cubic256", {"n_basis_cells": "32 32 2", "boundary_conditions": "0 0 0", "ddi_method": "fft", "llg_temperature": "0"}),
    ("cubic256", {"n_basis_cells": "64 16 2", "boundary_conditions": "0 0 0", "ddi_method": "fft", "llg_temperature": "0"}),
    ("cubic256", {"n_basis_cells": "128 32 1", "boundary_conditions": "0 0 0", "ddi_method": "fft", "llg_temperature": "0"}),
    ("cubic256", {"n_basis_cells": "256 4 2", "boundary_conditions": "0 1 0", "ddi_method": "fft", "llg_temperature": "0",
                  "ddi_n_periodic_images": "0 2 0"}),
    ("cubic256", {"n_basis_cells": "512 2 1", "boundary_conditions": "0 0 0", "ddi_method": "fft", "llg_temperature": "0"}),
    ("cubic256", {"n_basis_cells": "64 64 32", "boundary_conditions": "1 1 0", "ddi_method": "fft", "llg_temperature": "0",
                  "ddi_n_periodic_images": "1 1 0", "ddi_pb_zero_padding": "0"}),
    # the fast pass kernels in mixed company: 2-atom basis (half-length a-pass over 6 components, shared-memory c-pass),
    # un-padded periodic a (no zero half), a length that is not a power of two next to two that are
    ("ddi", {"n_basis_cells": "64 32 2", "boundary_conditions": "0 0 0"}),
    ("ddi", {"n_basis_cells": "32 32 32", "boundary_conditions": "0 0 0"}),
    ("cubic256", {"n_basis_cells": "128 16 2", "boundary_conditions": "1 0 0", "ddi_method": "fft", "llg_temperature": "0",
                  "ddi_n_periodic_images": "2 0 0", "ddi_pb_zero_padding": "0"}),
    ("cubic256", {"n_basis_cells": "36 32 32", "boundary_conditions": "0 0 0", "ddi_method": "fft", "llg_temperature": "0"}),
]


@pytest.mark.parametrize("preset,overrides", DDI_CASES)
def test_ddi_gradient_energy_vs_reference_fft(cfg, product, oracle, preset, overrides):
    path = cfg(preset, **overrides)
    p, o = S.Session(product, path), S.Session(oracle, path)
    s = unit_random(p.nos, 3)
    gp, ep = p.gradient_and_energy(s)
    go, eo = o.gradient_and_energy(s)
    assert np.abs(gp - go).max() <= 1e-12 * np.abs(go).max()
    cp, co = p.energy_contributions(s, per_spin=True), o.energy_contributions(s, per_spin=True)
    assert list(cp) == list(co) and "DDI" in cp
    abs_sum = sum(np.abs(v[1]).sum() for v in co.values())
    assert abs(ep - eo) <= 1e-12 * max(abs_sum, abs(eo))
    assert np.abs(cp["DDI"][1] - co["DDI"][1]).max() <= 1e-11 * np.abs(co["DDI"][1]).max()
    p.close()
    o.close()


@pytest.mark.parametrize("preset,overrides", [DDI_CASES[1], DDI_CASES[4], DDI_CASES[6]])
def test_ddi_gradient_vs_reference_direct_sum(cfg, product, oracle, preset, overrides):
    """Open boundaries: the FFT convolution equals the plain O(N^2) sum (reference: ddi_method cutoff, radius < 0 ->
    Gradient_DDI_Direct, Hamiltonian_Heisenberg.cpp:870-877,1016-1071)"""
    p = S.Session(product, cfg(preset, **overrides))
    o = S.Session(oracle, cfg(preset, **dict(overrides, ddi_method="cutoff", ddi_radius="-1")))
    s = unit_random(p.nos, 4)
    gp, _ = p.gradient_and_energy(s)
    go, _ = o.gradient_and_energy(s)
    assert np.abs(gp - go).max() <= 1e-12 * np.abs(go).max()
    p.close()
    o.close()


@pytest.mark.parametrize("solver", ["Depondt", "Heun", "SIB", "RK4", "VP"])
@pytest.mark.parametrize("preset,overrides", [DDI_CASES[0], DDI_CASES[4], DDI_CASES[5]])
def test_ddi_steps(cfg, product, oracle, solver, preset, overrides):
    """single shots and an amortised block with the dipolar field in the loop"""
    path = cfg(preset, llg_n_iterations_amortize=3, **overrides)
    p, o = S.Session(product, path), S.Session(oracle, path)
    s0 = unit_random(p.nos, 9)
    for x in (p, o):
        x.llg_set(temperature=0.0, damping=0.3, dt=1e-3)
        x.set_spins(s0)
        x.llg_start(S.SOLVERS[solver], single_shot=True)
        x.n_shot(4)
        x.stop()
    assert np.abs(p.spins() - o.spins()).max() < 1e-10
    assert abs(p.energy() - o.energy()) <= 1e-11 * max(1.0, abs(o.energy()))
    if solver != "VP":
        for x in (p, o):
            x.llg_start(S.SOLVERS[solver], n_iterations=6, n_iterations_log=6)
        assert np.abs(p.spins() - o.spins()).max() < 1e-10
    p.close()
    o.close()


# ---------------------------------------------------------------------------------------------------------------------------
# Every per-length instantiation of the fast pass kernels (64 ... 4096) against the reference. The reference's FFT path
# needs Na >= Nb >= Nc, so long b- and c-axes are checked through a symmetry of the simple cubic lattice with isotropic
# exchange + dipolar coupling only (no DMI, no anisotropy, no field): exchanging two lattice axes AND the same two spin
# components maps a configuration and its gradient onto those of the lattice with the two axis lengths exchanged.
# ---------------------------------------------------------------------------------------------------------------------------
ISOTROPIC = {"boundary_conditions": "0 0 0", "ddi_method": "fft", "llg_temperature": "0", "dij": "0", "n_shells_dmi": "0",
             "anisotropy_magnitude": "0", "external_field_magnitude": "0"}
AXIS_SWAPS = {"a": (None, None), "b": ((0, 2, 1, 3), [1, 0, 2]), "c": ((2, 1, 0, 3), [2, 1, 0])}


def _swap(field, cells, axis):
    """field [nos][3] on the lattice `cells` = (Na, Nb, Nc) -> the same field on the lattice with `axis` and a exchanged"""
    perm, comps = AXIS_SWAPS[axis]
    if perm is None:
        return field
    f4 = field.reshape(cells[2], cells[1], cells[0], 3)
    return np.ascontiguousarray(f4.transpose(perm)[..., comps]).reshape(-1, 3)


@pytest.mark.parametrize("axis,n_long", [("a", 64), ("a", 256), ("a", 1024), ("a", 2048),
                                         ("b", 64), ("b", 128), ("b", 256), ("b", 512), ("b", 1024), ("b", 2048),
                                         ("c", 64), ("c", 128), ("c", 256), ("c", 512), ("c", 1024), ("c", 2048)])
def test_ddi_every_transform_length_vs_reference(cfg, product, oracle, axis, n_long):
    short = 8
    cells_o = (n_long, short, short)  # the reference always sees the long axis as a
    cells_p = {"a": (n_long, short, short), "b": (short, n_long, short), "c": (short, short, n_long)}[axis]
    p = S.Session(product, cfg("cubic256", n_basis_cells="%d %d %d" % cells_p, **ISOTROPIC))
    o = S.Session(oracle, cfg("cubic256", n_basis_cells="%d %d %d" % cells_o, **ISOTROPIC))
    s_p = unit_random(p.nos, 11)
    s_o = _swap(s_p, cells_p, axis)
    gp, ep = p.gradient_and_energy(s_p)
    go, eo = o.gradient_and_energy(s_o)
    assert np.abs(_swap(gp, cells_p, axis) - go).max() <= 1e-12 * np.abs(go).max()
    assert abs(ep - eo) <= 1e-12 * np.abs(go).sum()
    p.close()
    o.close()


def test_thin_film_reduced_size_vp_against_reference(cfg, product, oracle):
    """configs[2] at the survey's reduced size: 256 x 256 x 4 open film, exchange + DMI + field + dipolar FFT convolution
    (padded 512 x 512 x 8: the film kernels with the c-transforms in registers). Gradient and energy on a random state, then
    30 VP iterations and 5 Depondt iterations against the reference's FFT path."""
    path = cfg("cubic256", n_basis_cells="256 256 4", boundary_conditions="0 0 0", ddi_method="fft", ddi_n_periodic_images="0 0 0",
               external_field_magnitude=25, anisotropy_magnitude=0, llg_temperature=0, llg_n_iterations_amortize=10)
    p, o = S.Session(product, path), S.Session(oracle, path)
    s0 = unit_random(p.nos, 17)
    gp, ep = p.gradient_and_energy(s0)
    go, eo = o.gradient_and_energy(s0)
    assert np.abs(gp - go).max() <= 1e-12 * np.abs(go).max()
    assert abs(ep - eo) <= 1e-11 * abs(eo)
    for solver, n in (("VP", 30), ("Depondt", 5)):
        for x in (p, o):
            x.llg_set(temperature=0.0, damping=0.3, dt=1e-3)
            x.set_spins(s0)
            x.llg_start(S.SOLVERS[solver], n_iterations=n, n_iterations_log=n)
        assert np.abs(o.spins() - s0).max() > 1e-4
        assert np.abs(p.spins() - o.spins()).max() < 1e-10, solver
        assert abs(p.energy() - o.energy()) <= 1e-10 * abs(o.energy()), solver
    p.close()
    o.close()
