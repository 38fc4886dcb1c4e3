"""Parity of the CUDA path (through the C ABI) with the reference CPU build, on the same seeded inputs.

Tolerances are BASELINE.json's: gradient and energy 1e-12 relative (fp64), single-step spin deviation < 1e-10.
Both libraries are driven through identical C API calls so both see identical (float-narrowed) parameters
(SURVEY.md 8c hazard 5).
"""
import ctypes

import numpy as np
import pytest

from spirit_b200 import session as S

pytestmark = pytest.mark.gpu

GRAD_RTOL = 1e-12
ENERGY_RTOL = 1e-12
STEP_ATOL = 1e-10


def unit_random(n, seed):
    rng = np.random.default_rng(seed)
    z = rng.uniform(-1, 1, n)
    phi = rng.uniform(-np.pi, np.pi, n)
    r = np.sqrt(1 - z * z)
    return np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1)


@pytest.fixture(params=["fused", "two-pass", "generic"])
def stencil_path(request, monkeypatch):
    """Run the test through the three kernel families that can serve a nearest-neighbour Hamiltonian (all read when the
    device tables are built): the fused predictor + corrector kernel (sc6_fused.cuh; Depondt, Heun, SIB), the two-pass
    marching kernels (sc6.cuh; SPIRIT_B200_NO_FUSED=1) and the generic gather kernels (SPIRIT_B200_GENERIC_STENCIL=1)"""
    monkeypatch.setenv("SPIRIT_B200_GENERIC_STENCIL", "1" if request.param == "generic" else "0")
    monkeypatch.setenv("SPIRIT_B200_NO_FUSED", "1" if request.param == "two-pass" else "0")
    return request.param


def pair(product, oracle, path):
    return S.Session(product, path), S.Session(oracle, path)


CASES = [
    # preset, overrides, extra setup
    ("solvers", {}, None),
    ("default", {"n_basis_cells": "12 10 3"}, None),
    ("default", {"n_basis_cells": "33 7 5", "boundary_conditions": "1 1 1"}, None),
    ("default", {"n_basis_cells": "7 6 5", "boundary_conditions": "0 0 0"}, "aniso"),
    ("fd_pairs", {}, None),
    ("cubic256", {"n_basis_cells": "16 12 10"}, None),
    ("cubic256", {"n_basis_cells": "40 3 2", "boundary_conditions": "1 0 1", "n_shells_exchange": "3", "jij": "10.0 -2.5 1.25",
                  "n_shells_dmi": "2", "dij": "6.0 1.5", "dm_chirality": "2"}, "aniso"),
    ("cubic256", {"n_basis_cells": "1 1 1"}, None),
    ("cubic256", {"n_basis_cells": "2 1 1", "boundary_conditions": "1 1 1"}, None),
    ("ddi", {"ddi_method": "none"}, "pairs2"),
]


def setup_extra(s, what):
    if what == "aniso":
        s.set_anisotropy(0.75, (0.0, 0.6, 0.8))
        s.set_cubic_anisotropy(0.5)
    return s


def make_case(cfg, product, oracle, preset, overrides, extra):
    kw = dict(overrides)
    pairs = None
    if extra == "pairs2":
        # two basis atoms, general pair list between them
        pairs = ["i j   da db dc   Jij  Dij  Dijx Dijy Dijz",
                 "0 1   0  0  0    4.0  1.5  0.3  0.4  1.0",
                 "0 0   1  0  0    3.0  0.5  1.0  0.0  0.2",
                 "1 1   0  1  0   -2.0  0.7  0.0  1.0  0.0",
                 "1 0   1  0  1    1.0  0.2  0.5  0.5  0.5",
                 "0 1   -1 2  0    0.5  0.1  0.0  0.3  0.9"]
    path = cfg(preset, pairs=pairs, **kw)
    p, o = pair(product, oracle, path)
    setup_extra(p, extra)
    setup_extra(o, extra)
    return p, o


@pytest.mark.parametrize("preset,overrides,extra", CASES)
def test_gradient_and_energy(cfg, product, oracle, preset, overrides, extra):
    p, o = make_case(cfg, product, oracle, preset, overrides, extra)
    assert p.nos == o.nos
    for seed in (1, 2):
        s = unit_random(p.nos, seed)
        gp, ep = p.gradient_and_energy(s)
        go, eo = o.gradient_and_energy(s)
        scale = max(np.abs(go).max(), 1e-300)
        assert np.abs(gp - go).max() <= GRAD_RTOL * scale
        # the reference's own summation-order noise is 1e-13 relative (SURVEY.md 8c hazard 4): normalise by sum |e_i|
        contrib = o.energy_contributions(s, per_spin=True)
        abs_sum = sum(np.abs(v[1]).sum() for v in contrib.values())
        assert abs(ep - eo) <= ENERGY_RTOL * max(abs_sum, abs(eo))
        g2 = p.gradient(s)
        assert np.array_equal(g2, gp)
    p.close()
    o.close()


@pytest.mark.parametrize("table", [["n_anisotropy 2", "i K Kx Ky Kz K4", "0 1.5 0 0 1 0.0", "1 0.7 1 1 0 0.25"],
                                   ["n_anisotropy 2", "i Ka Kb Kc", "0 0.0 0.0 2.0", "1 0.5 0.5 0.0"]])
def test_gradient_with_anisotropy_table(tmp_path, product, oracle, table):
    """per-atom anisotropy tables (n_anisotropy): different axes and magnitudes on the two atoms of the basis, cubic term on one"""
    from tests import cfgs
    path = tmp_path / "a.cfg"
    path.write_text(cfgs.render("cubic256", block=["basis", "2", "0 0 0", "0.5 0.5 0.5"] + table, n_basis_cells="6 5 4"))
    p, o = S.Session(product, str(path)), S.Session(oracle, str(path))
    s = unit_random(p.nos, 3)
    gp, ep = p.gradient_and_energy(s)
    go, eo = o.gradient_and_energy(s)
    assert np.abs(gp - go).max() <= GRAD_RTOL * np.abs(go).max()
    assert abs(ep - eo) <= 1e-11 * abs(eo)
    cp, co = p.energy_contributions(s), o.energy_contributions(s)
    assert list(cp.keys()) == list(co.keys())
    for name in co:
        assert abs(cp[name][0] - co[name][0]) <= 1e-11 * max(abs(co[name][0]), 1.0), name
    p.close()
    o.close()


@pytest.mark.parametrize("preset,overrides,extra", [CASES[0], CASES[3], CASES[6], CASES[9]])
def test_energy_contributions(cfg, product, oracle, preset, overrides, extra):
    p, o = make_case(cfg, product, oracle, preset, overrides, extra)
    s = unit_random(p.nos, 7)
    cp = p.energy_contributions(s, per_spin=True)
    co = o.energy_contributions(s, per_spin=True)
    assert list(cp.keys()) == list(co.keys())
    for name in co:
        scale = max(np.abs(co[name][1]).max(), 1e-300)
        assert np.abs(cp[name][1] - co[name][1]).max() <= 1e-12 * scale, name
        assert abs(cp[name][0] - co[name][0]) <= 1e-12 * max(np.abs(co[name][1]).sum(), 1e-300), name
    p.close()
    o.close()


@pytest.mark.parametrize("solver", ["Depondt", "Heun", "SIB", "RK4", "VP"])
@pytest.mark.parametrize("preset,overrides,extra", [CASES[0], CASES[2], CASES[3], CASES[4], CASES[5], CASES[6], CASES[9]])
def test_single_steps(cfg, product, oracle, solver, preset, overrides, extra, stencil_path):
    """Simulation_SingleShot x 5 (VP: x 20) from the same random state: max spin-component deviation < 1e-10"""
    p, o = make_case(cfg, product, oracle, preset, overrides, extra)
    s0 = unit_random(p.nos, 11)
    n = 20 if solver == "VP" else 5
    for x in (p, o):
        x.llg_set(temperature=0.0, damping=0.3, dt=1e-3)
        x.set_spins(s0)
        x.llg_start(S.SOLVERS[solver], single_shot=True)
        for _ in range(n):
            x.single_shot()
    dev = np.abs(p.spins() - o.spins()).max()
    moved = np.abs(o.spins() - s0).max()
    assert moved > 1e-4  # the comparison is not vacuous
    assert dev < STEP_ATOL
    # energy after the last hook, torque and effective field
    assert abs(p.energy() - o.energy()) <= 1e-11 * max(1.0, abs(o.energy()))
    assert abs(p.max_torque() - o.max_torque()) <= 1e-9 * max(1e-30, o.max_torque())
    fo = o.effective_field()
    assert np.abs(p.effective_field() - fo).max() <= 1e-9 * max(np.abs(fo).max(), 1e-300)
    for x in (p, o):
        x.stop()
        x.close()


BLOCK_CASES = [CASES[0], CASES[1], CASES[2], CASES[3], CASES[4], CASES[5], CASES[6],
               ("cubic256", {"n_basis_cells": "130 5 9", "boundary_conditions": "1 0 1"}, None),
               ("cubic256", {"n_basis_cells": "20 20 40", "boundary_conditions": "0 1 0", "dm_chirality": "2"}, "aniso"),
               # whole tiles of the fused kernel for tile heights 13 and 16 (its interior variant), several c-segments
               ("cubic256", {"n_basis_cells": "64 208 6"}, None),
               ("cubic256", {"n_basis_cells": "96 52 11", "boundary_conditions": "0 0 0", "external_field_magnitude": "3"}, "aniso"),
               ("cubic256", {"n_basis_cells": "32 48 2", "boundary_conditions": "1 1 0"}, None),
               ("default", {"n_basis_cells": "100 100 1"}, None)]


@pytest.mark.parametrize("solver", ["Depondt", "Heun", "SIB", "RK4"])
@pytest.mark.parametrize("preset,overrides,extra", BLOCK_CASES)
def test_iterate_block(cfg, product, oracle, solver, preset, overrides, extra, stencil_path):
    """Simulation_LLG_Start over blocks of amortised iterations (only the last iteration of a block runs the hook
    variant of the kernels): spins, energy and torque against the reference after 12 iterations"""
    kw = dict(overrides, llg_n_iterations_amortize=4)
    p, o = make_case(cfg, product, oracle, preset, kw, extra)
    s0 = unit_random(p.nos, 5)
    for x in (p, o):
        x.llg_set(temperature=0.0, damping=0.3, dt=1e-3)
        x.set_spins(s0)
        x.llg_start(S.SOLVERS[solver], n_iterations=12, n_iterations_log=12)
    assert np.abs(o.spins() - s0).max() > 1e-4
    assert np.abs(p.spins() - o.spins()).max() < STEP_ATOL
    assert abs(p.energy() - o.energy()) <= 1e-11 * max(1.0, abs(o.energy()))
    fo = o.effective_field()
    assert np.abs(p.effective_field() - fo).max() <= 1e-9 * max(np.abs(fo).max(), 1e-300)
    p.close()
    o.close()


@pytest.mark.parametrize("solver", ["Depondt", "Heun", "SIB"])
@pytest.mark.parametrize("cells,bc,T", [("64 208 6", "1 1 1", 0.0), ("64 208 6", "1 1 1", 10.0), ("70 30 9", "1 0 0", 10.0),
                                        ("96 52 40", "0 1 1", 5.0), ("48 48 1", "1 1 0", 10.0)])
def test_fused_equals_two_pass(cfg, product, monkeypatch, solver, cells, bc, T):
    """The fused predictor + corrector kernel against the two-pass marching kernels from the same state, also at T > 0:
    the thermal field is a function of (site, plane, iteration) only, so even the rim sites that neighbouring CTAs
    recompute see the same noise. 9 iterations in blocks of 3 (hook variant on every third); spins equal to 1e-14,
    hook energy and torque to 1e-12."""
    out = []
    for no_fused in ("0", "1"):
        monkeypatch.setenv("SPIRIT_B200_GENERIC_STENCIL", "0")
        monkeypatch.setenv("SPIRIT_B200_NO_FUSED", no_fused)
        p = S.Session(product, cfg("cubic256", n_basis_cells=cells, boundary_conditions=bc, llg_n_iterations_amortize=3,
                                   llg_seed=4711))
        p.llg_set(temperature=T, damping=0.3, dt=1e-3)
        p.set_spins(unit_random(p.nos, 21))
        p.llg_start(S.SOLVERS[solver], n_iterations=9, n_iterations_log=9)
        out.append((p.spins().copy(), p.energy(), p.max_torque(), p.effective_field().copy()))
        p.close()
    (sf, ef, tf, ff), (st, et, tt, ft) = out
    assert np.abs(sf - st).max() < 1e-14
    assert abs(ef - et) <= 1e-12 * abs(et)
    assert abs(tf - tt) <= 1e-12 * tt
    assert np.abs(ff - ft).max() <= 1e-12 * np.abs(ft).max()


def test_vp_amortized_block(cfg, product, oracle):
    """VP trajectories depend on n_iterations_amortize through the in-place projected force (SURVEY.md 8c hazard 6)"""
    for amortize in (1, 10):
        path = cfg("solvers", llg_n_iterations_amortize=amortize)
        p, o = pair(product, oracle, path)
        for x in (p, o):
            x.plus_z()
            x.skyrmion(5.0, phase=-90.0)
            x.llg_start(S.SOLVER_VP, n_iterations=200, n_iterations_log=200)
        assert np.abs(p.spins() - o.spins()).max() < 1e-9, amortize
        p.close()
        o.close()


@pytest.mark.parametrize("solver", ["Heun", "Depondt", "SIB", "RK4"])
def test_larmor_known_answer(cfg, product, solver):
    """Closed form of a single damped precessing spin (core/test/test_physics.cpp:20-88), abs tol 1e-6"""
    from spirit_b200.capi import load_product
    p = S.Session(product, cfg("larmor"))
    damping, dt, B = 0.3, 0.001, 1.0
    p.llg_set(damping=damping, dt=dt)
    p.domain((1.0, 0.0, 0.0))
    p.llg_start(S.SOLVERS[solver], single_shot=True)
    gamma, mu_B = product.Constants_gamma(), product.Constants_mu_B()
    dtg = dt * gamma / (1.0 + damping ** 2)
    for i in range(100):
        p.single_shot()
        s = p.spins()[0]
        phi = dtg * (i + 1) * B
        sz = np.tanh(damping * dtg * (i + 1) * B)
        rxy = np.sqrt(1 - sz ** 2)
        assert abs(s[0] - np.cos(phi) * rxy) < 1e-6
        assert abs(s[2] - sz) < 1e-6
    p.stop()
    p.close()


def test_skyrmion_relaxation_golden(cfg, product):
    """core/test/test_solvers.cpp:44-45: all solvers relax the 16x16 skyrmion to E = -5849.69140625, Mz = 2*0.79977"""
    for solver in ("VP", "Depondt", "Heun", "SIB", "RK4"):
        p = S.Session(product, cfg("solvers"))
        p.plus_z()
        p.skyrmion(5.0, phase=-90.0)
        p.llg_set(direct_minimization=True)
        p.llg_start(S.SOLVERS[solver])
        p.update_data()
        assert abs(p.energy() - (-5849.69140625)) < 1e-5 * 5849.0, solver
        m = p.magnetization()
        assert abs(m[2] - 2 * 0.79977) < 1e-4, solver
        p.close()


def test_thermal_langevin_known_answer(cfg, product):
    """Non-interacting spins at T>0 obey <s_z> = coth x - 1/x, x = mu_s mu_B B/(k_B T): pins the
    amplitude of the Philox thermal field independently of the RNG stream. Oracle: 0.3987 +- 0.0030.
    16384 spins, 12 snapshots 2 ps apart (relaxation time 1 / (alpha gamma B) ~ 1.9 ps): Var(s_z) = 0.24 -> the standard
    error of the mean is 0.0011; the seed is fixed (without llg_seed the reference's default is libc random(), which
    depends on how many states the process has set up before)."""
    path = cfg("fd_pairs", pairs=["i j da db dc Jij"], n_basis_cells="128 128 1", external_field_magnitude="10",
               llg_temperature="10", llg_damping="0.3", llg_dt="1e-3", llg_n_iterations_amortize="100", llg_seed="20006")
    x = 2.0 * product.Constants_mu_B() * 10.0 / (product.Constants_k_B() * 10.0)
    expected = 1.0 / np.tanh(x) - 1.0 / x
    for solver in ("Depondt", "SIB", "Heun"):
        p = S.Session(product, path)
        p.plus_z()
        p.llg_start(S.SOLVERS[solver], n_iterations=8000, n_iterations_log=8000)
        means = []
        for _ in range(12):
            p.llg_start(S.SOLVERS[solver], n_iterations=2000, n_iterations_log=2000)
            means.append(p.spins()[:, 2].mean())
        means = np.array(means)
        err = means.std(ddof=1) / np.sqrt(len(means))
        assert abs(means.mean() - expected) < max(4 * err, 0.005), (solver, means.mean(), expected, err)
        p.close()


@pytest.mark.parametrize("solver", ["Depondt", "SIB", "Heun"])
def test_thermal_averages_match_reference_within_error_bars(cfg, product, oracle, solver):
    """BASELINE.json: at T > 0, time-averaged magnetisation and energy must agree within statistical error bars.
    Interacting system (J, D, K, B, T = 40 K on 16x16x8), independent noise streams (reference: serial mt19937; here:
    Philox), 3000 equilibration steps, then 30 block averages of 300 steps each. Error bars from the scatter of the
    blocks; the two means must agree within 4 combined standard errors."""
    path = cfg("cubic256", n_basis_cells="16 16 8", llg_temperature=40, llg_n_iterations_amortize=50, external_field_magnitude=5)
    stats = []
    for lib in (product, oracle):
        x = S.Session(lib, path)
        x.plus_z()
        x.llg_start(S.SOLVERS[solver], n_iterations=3000, n_iterations_log=3000)
        m, e = [], []
        for _ in range(30):
            x.llg_start(S.SOLVERS[solver], n_iterations=300, n_iterations_log=300)
            m.append(x.spins()[:, 2].mean())
            x.update_data()
            e.append(x.energy() / x.nos)
        stats.append((np.mean(m), np.std(m, ddof=1) / np.sqrt(len(m)), np.mean(e), np.std(e, ddof=1) / np.sqrt(len(e))))
        x.close()
    (mp, dmp, ep, dep), (mo, dmo, eo, deo) = stats
    assert 0.5 < mo < 0.99  # thermally disordered but not paramagnetic: the comparison is meaningful
    assert abs(mp - mo) < 4 * np.hypot(dmp, dmo), (solver, mp, dmp, mo, dmo)
    assert abs(ep - eo) < 4 * np.hypot(dep, deo), (solver, ep, dep, eo, deo)


def test_stencil_variant_probe(cfg, product, monkeypatch):
    """The nearest-neighbour Hamiltonians of the BASELINE configs are served by the marching kernels (and the tests
    parametrised over `stencil_path` do exercise both code paths); a general pair list by the gather kernels."""
    monkeypatch.setenv("SPIRIT_B200_GENERIC_STENCIL", "0")
    p = S.Session(product, cfg("cubic256", n_basis_cells="16 12 10"))
    assert p.stencil_variant() == 1
    p.close()
    monkeypatch.setenv("SPIRIT_B200_GENERIC_STENCIL", "1")
    p = S.Session(product, cfg("cubic256", n_basis_cells="16 12 10"))
    assert p.stencil_variant() == 0
    p.close()
    monkeypatch.setenv("SPIRIT_B200_GENERIC_STENCIL", "0")
    p = S.Session(product, cfg("cubic256", n_basis_cells="16 12 10", n_shells_exchange="2", jij="10.0 -2.5"))
    assert p.stencil_variant() == 0
    p.close()


OSO_CASES = [CASES[0], CASES[2], CASES[3], CASES[5], CASES[6], CASES[9]]


@pytest.fixture
def serial_oracle():
    """The reference's lbfgs_atlas_transform_direction updates rho[n] inside an OpenMP loop without a reduction
    (core/src/engine/Solver_Kernels.cpp:214-238): with several threads its result depends on the interleaving. Run the
    oracle on one thread where that kernel is compared (TEST INFRASTRUCTURE only)."""
    import ctypes
    gomp = ctypes.CDLL("libgomp.so.1")
    gomp.omp_get_max_threads.restype = ctypes.c_int
    before = gomp.omp_get_max_threads()
    gomp.omp_set_num_threads(1)
    yield
    gomp.omp_set_num_threads(before)


@pytest.mark.parametrize("solver", ["VP_OSO", "LBFGS_OSO", "LBFGS_Atlas"])
@pytest.mark.parametrize("preset,overrides,extra", OSO_CASES)
def test_oso_minimisers_match_reference(cfg, product, oracle, serial_oracle, solver, preset, overrides, extra):
    """Solver_VP_OSO.hpp:34-115 / Solver_LBFGS_OSO.hpp:39-77 + lbfgs_get_searchdir (Solver_Kernels.hpp:44-190):
    25 iterations in amortised blocks of 5 from the same random state (the L-BFGS memory wraps around several times and
    the step limiter is active at first). The scalars of the recursion are sums over all sites, folded in a different
    order than the reference's OpenMP reduction, so parity is to 1e-9 in the spins rather than bit-level."""
    kw = dict(overrides, llg_n_iterations_amortize=5)
    p, o = make_case(cfg, product, oracle, preset, kw, extra)
    s0 = unit_random(p.nos, 13)
    for x in (p, o):
        x.llg_set(temperature=0.0, damping=0.3, dt=1e-3)
        x.set_spins(s0)
        x.llg_start(S.SOLVERS[solver], n_iterations=25, n_iterations_log=25)
    assert np.abs(o.spins() - s0).max() > 1e-4
    assert np.abs(p.spins() - o.spins()).max() < 1e-9
    assert abs(p.energy() - o.energy()) <= 1e-10 * max(1.0, abs(o.energy()))
    assert abs(p.max_torque() - o.max_torque()) <= 1e-7 * max(1e-30, o.max_torque())
    p.close()
    o.close()


@pytest.mark.parametrize("solver", ["VP_OSO", "LBFGS_OSO", "LBFGS_Atlas"])
def test_oso_single_shots_match_reference(cfg, product, oracle, serial_oracle, solver):
    """Simulation_SingleShot x 12 (a hook after every iteration)"""
    p, o = make_case(cfg, product, oracle, "solvers", {}, None)
    for x in (p, o):
        x.plus_z()
        x.skyrmion(5.0, phase=-90.0)
        x.llg_start(S.SOLVERS[solver], single_shot=True)
        for _ in range(12):
            x.single_shot()
    assert np.abs(p.spins() - o.spins()).max() < 1e-9
    assert abs(p.energy() - o.energy()) <= 1e-10 * abs(o.energy())
    for x in (p, o):
        x.stop()
        x.close()


def test_oso_skyrmion_relaxation_golden(cfg, product):
    """core/test/test_solvers.cpp:39-72: LBFGS_Atlas, LBFGS_OSO and VP_OSO relax the 16x16 skyrmion to E = -5849.69140625,
    Mz = 2*0.79977"""
    for solver in ("LBFGS_Atlas", "LBFGS_OSO", "VP_OSO"):
        p = S.Session(product, cfg("solvers"))
        p.plus_z()
        p.skyrmion(5.0, phase=-90.0)
        p.llg_set(direct_minimization=True)
        p.llg_start(S.SOLVERS[solver])
        p.update_data()
        assert abs(p.energy() - (-5849.69140625)) < 1e-5 * 5849.0, solver
        m = p.magnetization()
        assert abs(m[2] - 2 * 0.79977) < 1e-4, solver
        p.close()


def test_temperature_gradient_langevin_profile(cfg, product):
    """Linear temperature gradient (Method_LLG.cpp:80-96, Vectormath::get_gradient_distribution): non-interacting spins in
    a field at a site temperature T(x) = T0 + g x obey <s_z>(x) = coth X - 1/X, X = mu_s mu_B B / (k_B T(x)). The
    profile along the gradient pins both the amplitude epsilon sqrt(T_i / mu_s) and its dependence on the position."""
    path = cfg("fd_pairs", pairs=["i j da db dc Jij"], n_basis_cells="64 256 1", external_field_magnitude="10",
               llg_temperature="5", llg_temperature_gradient_direction="1 0 0", llg_temperature_gradient_inclination="0.25",
               llg_damping="0.3", llg_dt="1e-3", llg_n_iterations_amortize="100", llg_seed="20006")
    # bands of 4096 spins, 12 snapshots 2 ps apart: standard error of a band mean ~ 0.002 (fixed seed, see the test above)
    T = 5.0 + 0.25 * np.arange(64)
    X = 2.0 * product.Constants_mu_B() * 10.0 / (product.Constants_k_B() * T)
    langevin = 1.0 / np.tanh(X) - 1.0 / X
    expected = langevin.reshape(4, 16).mean(axis=1)  # four bands of 16 columns
    assert expected[0] - expected[3] > 0.25  # the profile is far from flat
    p = S.Session(product, path)
    p.plus_z()
    p.llg_start(S.SOLVER_DEPONDT, n_iterations=8000, n_iterations_log=8000)
    bands = []
    for _ in range(12):
        p.llg_start(S.SOLVER_DEPONDT, n_iterations=2000, n_iterations_log=2000)
        sz = p.spins()[:, 2].reshape(256, 64)  # [b][a]
        bands.append(sz.reshape(256, 4, 16).mean(axis=(0, 2)))
    bands = np.array(bands)
    mean, err = bands.mean(axis=0), bands.std(axis=0, ddof=1) / np.sqrt(len(bands))
    for k in range(4):
        assert abs(mean[k] - expected[k]) < max(4 * err[k], 0.012), (k, mean[k], expected[k], err[k])
    p.close()


STT_CASES = [
    # lattice / boundary conditions of the finite differences: periodic, open (one-sided differences), single cells along an axis,
    # a two-atom basis on skewed Bravais vectors (inverse of the lattice matrix), several shells
    ("solvers", {"n_basis_cells": "12 9 1"}),
    ("cubic256", {"n_basis_cells": "9 7 5", "boundary_conditions": "0 1 0", "llg_temperature": "0"}),
    ("cubic256", {"n_basis_cells": "6 5 4", "boundary_conditions": "0 0 0", "llg_temperature": "0"}),
    ("ddi", {"ddi_method": "none", "n_basis_cells": "5 4 3"}),
]


@pytest.mark.parametrize("gradient", [0, 1])
@pytest.mark.parametrize("solver", ["Depondt", "Heun", "SIB", "RK4"])
@pytest.mark.parametrize("preset,overrides", STT_CASES)
def test_spin_transfer_torque(cfg, product, oracle, solver, preset, overrides, gradient):
    """Spin-transfer torque in Calculate_Force_Virtual: the monolayer approximation (Method_LLG.cpp:207-212) and the gradient
    approximation for in-plane currents (Method_LLG.cpp:184-205 with Vectormath::jacobian, Vectormath.cpp:816-903)"""
    extra = "pairs2" if preset == "ddi" else None
    kw = dict(overrides, llg_stt_use_gradient=gradient, llg_stt_magnitude="1.7", llg_stt_polarisation_normal="0.6 -0.3 0.74",
              llg_beta="0.12", llg_damping="0.25")
    p, o = make_case(cfg, product, oracle, preset, kw, extra)
    s0 = unit_random(p.nos, 5)
    for x in (p, o):
        x.set_spins(s0)
        x.llg_start(S.SOLVERS[solver], single_shot=True)
        x.n_shot(6)
    ref = o.spins().copy()
    dev = np.abs(p.spins() - ref).max()
    assert np.abs(ref - s0).max() > 1e-4
    assert dev < STEP_ATOL, dev
    for x in (p, o):
        x.stop()
    # the torque does something: the same steps without it end elsewhere
    o.set_spins(s0)
    o.lib.Parameters_LLG_Set_STT(o.state, bool(gradient), 0.0, (ctypes.c_float * 3)(0.6, -0.3, 0.74), -1, -1)
    o.llg_start(S.SOLVERS[solver], single_shot=True)
    o.n_shot(6)
    assert np.abs(o.spins() - ref).max() > 1e-6
    o.stop()
    p.close()
    o.close()
