"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle cannot run 256^3 in seconds):

  * replication: a periodic lattice filled with an n-fold periodic repetition of a small configuration evolves as that
    repetition -- so the corner of the full-size run must equal the SMALL lattice evolved by the reference (T = 0; the
    counter-based noise is keyed by the site and is not periodic). This ties the full-size launch geometry (march
    segments, boundary / interior variants, tiles of the FFT passes, image batching) to the reference's result.
  * translation covariance, unit norm, energy dissipation for the LLG step at 256^3
  * linearity of the dipolar gradient at 2048 x 2048 x 4
Tolerances are BASELINE.json's (single-step spin deviation < 1e-10, gradient 1e-12 relative).
"""
import numpy as np
import pytest

from spirit_b200 import session as S
from tests.test_parity_gpu import unit_random

pytestmark = pytest.mark.gpu


def tiled(small, cells_small, reps):
    """small: [nos_small][3] in the reference's site order (a fastest) -> the reps-fold periodic repetition"""
    na, nb, nc = cells_small
    a = small.reshape(nc, nb, na, 3)
    return np.ascontiguousarray(np.tile(a, (reps[2], reps[1], reps[0], 1))).reshape(-1, 3)


def corner(big, cells_big, cells_small):
    na, nb, nc = cells_big
    a = big.reshape(nc, nb, na, 3)
    return a[:cells_small[2], :cells_small[1], :cells_small[0]].reshape(-1, 3)


@pytest.mark.parametrize("solver,n", [("Depondt", 10), ("SIB", 10), ("Heun", 10), ("RK4", 6), ("VP", 20)])
def test_256_cubed_replicates_the_reference_on_64_cubed(cfg, product, oracle, solver, n):
    """configs[1] at full size (256^3, J + DMI + K, periodic) at T = 0: a 4 x 4 x 4 repetition of a random 64^3 state"""
    small, big = (64, 64, 64), (256, 256, 256)
    s0 = unit_random(64 ** 3, 21)
    o = S.Session(oracle, cfg("cubic256", n_basis_cells="64 64 64", llg_temperature=0, llg_n_iterations_amortize=n))
    o.llg_set(temperature=0.0, damping=0.3, dt=1e-3)
    o.set_spins(s0)
    o.llg_start(S.SOLVERS[solver], n_iterations=n, n_iterations_log=n)
    ref = o.spins().copy()
    e_ref = o.energy()
    o.close()

    p = S.Session(product, cfg("cubic256", llg_temperature=0, llg_n_iterations_amortize=n))
    p.llg_set(temperature=0.0, damping=0.3, dt=1e-3)
    p.set_spins(tiled(s0, small, (4, 4, 4)))
    assert p.stencil_variant() == 1
    p.llg_start(S.SOLVERS[solver], n_iterations=n, n_iterations_log=n)
    out = p.spins()
    assert np.abs(ref - s0).max() > 1e-4
    assert np.abs(corner(out, big, small) - ref).max() < 1e-10
    # and the whole lattice is still the repetition of its corner, every site a unit vector
    assert np.abs(out - tiled(corner(out, big, small), small, (4, 4, 4))).max() < 1e-12
    assert np.abs(np.einsum("ij,ij->i", out, out) - 1.0).max() < 1e-12
    assert abs(p.energy() - 64 * e_ref) <= 1e-10 * abs(64 * e_ref)
    p.close()


def test_256_cubed_translation_covariance_and_dissipation(cfg, product):
    """Shifting the initial state by a lattice vector shifts the result (bit for bit: every site runs the same
    arithmetic wherever it sits in a CTA, a march segment or a boundary variant); damped dynamics lowers the energy"""
    p = S.Session(product, cfg("cubic256", llg_temperature=0, llg_n_iterations_amortize=8))
    p.llg_set(temperature=0.0, damping=0.3, dt=1e-3)
    s0 = unit_random(256 ** 3, 22)
    shift = (5, 37, 129)  # cells along a, b, c
    res = []
    for state in (s0, np.roll(s0.reshape(256, 256, 256, 3), (shift[2], shift[1], shift[0]), axis=(0, 1, 2)).reshape(-1, 3)):
        p.set_spins(np.ascontiguousarray(state))
        p.update_data()
        e0 = p.energy()
        p.llg_start(S.SOLVER_DEPONDT, n_iterations=8, n_iterations_log=8)
        assert p.energy() < e0
        res.append(p.spins().copy())
    back = np.roll(res[1].reshape(256, 256, 256, 3), (-shift[2], -shift[1], -shift[0]), axis=(0, 1, 2)).reshape(-1, 3)
    assert np.array_equal(back, res[0])
    p.close()


def test_thin_film_dipolar_gradient_is_linear_at_full_size(cfg, product):
    """configs[2]: 2048 x 2048 x 4 open film, dipole-dipole interaction only (FFT convolution): the gradient is linear in the
    spins, g(a s1 + b s2) = a g(s1) + b g(s2), and the energy is the quadratic form 1/2 s.g"""
    path = cfg("cubic256", n_basis_cells="2048 2048 4", boundary_conditions="0 0 0", ddi_method="fft",
               ddi_n_periodic_images="0 0 0", n_shells_exchange=0, jij="0.0", n_shells_dmi=0, dij="0.0", anisotropy_magnitude=0,
               external_field_magnitude=0, llg_temperature=0)
    p = S.Session(product, path)
    n = p.nos
    assert n == 2048 * 2048 * 4
    s1, s2 = unit_random(n, 31), unit_random(n, 32)
    g1, e1 = p.gradient_and_energy(s1)
    g2, _ = p.gradient_and_energy(s2)
    g12, _ = p.gradient_and_energy(0.75 * s1 - 1.5 * s2)
    scale = np.abs(g1).max()
    assert scale > 0
    assert np.abs(g12 - (0.75 * g1 - 1.5 * g2)).max() <= 1e-12 * scale * 3
    assert abs(e1 - 0.5 * np.einsum("ij,ij->", s1, g1)) <= 1e-11 * abs(e1)
    p.close()


def test_gneb_64_images_replicate_the_reference_chain(cfg, product, oracle):
    """configs[3] at full size: 64 images of 256 x 256 x 1 (climbing image set automatically). A 16 x 16 repetition of the
    reference's 16 x 16 skyrmion-collapse chain evolves as that repetition: tangents, spring and climbing forces and VP's
    projection are built from sums over the sites that scale consistently, so every image must equal the reference's small
    image, energies scale by 256 and the reaction coordinate by 16."""
    noi, reps = 64, 16

    def build(x):
        x.plus_z()
        x.skyrmion(5.0, phase=-90.0)
        x.chain_set_length(noi)
        x.jump_to_image(noi - 1)
        x.plus_z()
        x.jump_to_image(0)
        x.transition_homogeneous(0, noi - 1)

    o = S.Session(oracle, cfg("solvers", gneb_n_iterations_amortize=10))
    build(o)
    small0 = [o.spins(i).copy() for i in range(noi)]
    o.gneb_set_image_type(S.GNEB_CLIMBING, 20)
    o.gneb_start(S.SOLVER_VP, n_iterations=30, n_iterations_log=30)
    ref = [o.spins(i).copy() for i in range(noi)]
    rx_ref, e_ref = o.chain_rx_e()
    o.close()

    p = S.Session(product, cfg("solvers", n_basis_cells="256 256 1", gneb_n_iterations_amortize=10))
    p.plus_z()
    p.chain_set_length(noi)
    for i in range(noi):
        p.set_spins(tiled(small0[i], (16, 16, 1), (reps, reps, 1)), i)
    p.gneb_set_image_type(S.GNEB_CLIMBING, 20)
    p.gneb_start(S.SOLVER_VP, n_iterations=30, n_iterations_log=30)
    rx, e = p.chain_rx_e()
    assert np.abs(np.stack(ref) - np.stack(small0)).max() > 1e-4
    for i in range(noi):
        assert np.abs(corner(p.spins(i), (256, 256, 1), (16, 16, 1)) - ref[i]).max() < 1e-9, i
    assert np.abs(rx - reps * rx_ref).max() <= 1e-9 * reps * rx_ref.max()
    assert np.abs(e - reps * reps * e_ref).max() <= 1e-10 * np.abs(reps * reps * e_ref).max()
    p.close()


def test_dipolar_fast_passes_equal_generic_passes_at_256_cubed(cfg, product, monkeypatch):
    """BASELINE configs[4] per-GPU size: the per-length in-place pass kernels (half-length real a-pass, real tile-ordered
    tensor) against the generic two-buffer mixed-radix passes of the same library, which are pinned against the reference at
    the sizes it can run (tests/test_ddi_gpu.py). Two independent code paths, one answer."""
    import os
    over = dict(n_basis_cells="256 256 256", boundary_conditions="0 0 0", ddi_method="fft", llg_temperature="0")
    fast = S.Session(product, cfg("cubic256", **over))
    s = unit_random(fast.nos, 5)
    g_fast, e_fast = fast.gradient_and_energy(s)
    fast.close()
    monkeypatch.setenv("SPIRIT_B200_FFT_FAST", "0")
    generic = S.Session(product, cfg("cubic256", **over))
    g_gen, e_gen = generic.gradient_and_energy(s)
    generic.close()
    assert np.abs(g_fast).max() > 1.0
    assert np.abs(g_fast - g_gen).max() <= 1e-12 * np.abs(g_gen).max()
    assert abs(e_fast - e_gen) <= 1e-12 * np.abs(g_gen).sum()
