"""GNEB image loop on the device against the reference: tangents, spring / climbing / falling forces, Rx, energies,
solver updates over all images. BASELINE tolerance: energy barrier within 1e-8 relative."""
import os

import numpy as np
import pytest

from spirit_b200 import session as S

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def make_chain(x, noi=7, radius=3.0, K=0.25):
    if K:
        x.set_anisotropy(K, (0, 0, 1))
    x.plus_z()
    x.skyrmion(radius, phase=-90.0)
    x.chain_set_length(noi)
    x.jump_to_image(noi - 1)
    x.plus_z()
    x.jump_to_image(0)
    x.transition_homogeneous(0, noi - 1)


def test_gneb_golden(cfg, product):
    """7 images of 10x10, image 3 climbing, 60 VP single shots: golden vectors from the reference"""
    g = dict(np.load(os.path.join(GOLD, "gneb_7x10x10.npz")))
    p = S.Session(product, cfg("solvers", n_basis_cells="10 10 1"))
    make_chain(p)
    assert np.abs(np.stack([p.spins(i) for i in range(7)]) - g["images0"]).max() < 1e-14
    p.gneb_set_image_type(S.GNEB_CLIMBING, 3)
    p.gneb_start(S.SOLVER_VP, single_shot=True)
    p.n_shot(60)
    imgs = np.stack([p.spins(i).copy() for i in range(7)])
    assert np.abs(imgs - g["images"]).max() < 1e-10
    rx, e = p.chain_rx_e()
    assert np.abs(rx - g["Rx"]).max() < 1e-10
    assert np.abs(e - g["E"]).max() <= 1e-11 * np.abs(g["E"]).max()
    assert abs(p.chain_max_torque() - g["max_torque"]) <= 1e-9 * g["max_torque"]
    p.stop()
    p.close()


TYPES = [{}, {3: S.GNEB_CLIMBING}, {2: S.GNEB_FALLING, 4: S.GNEB_CLIMBING, 5: S.GNEB_STATIONARY}]


@pytest.mark.parametrize("solver", ["Depondt", "Heun", "SIB"])
@pytest.mark.parametrize("types", TYPES)
def test_gneb_two_stage_solvers_match_restatement(cfg, product, solver, types):
    """Depondt / Heun / SIB over all images against the NumPy restatement. For SIB this is the only usable oracle: the
    compiled reference reads uninitialised end-image virtual forces (Solver_SIB.hpp:4-5), its results vary from run to
    run (oracle hazards)."""
    from oracle import restatement as R
    p = S.Session(product, cfg("solvers", n_basis_cells="12 10 1", boundary_conditions="1 0 0"))
    make_chain(p, noi=8)
    imgs0 = [p.spins(i).copy() for i in range(8)]
    t = [R.NORMAL] * 8
    for img, ty in types.items():
        p.gneb_set_image_type(ty, img)
        t[img] = ty
    p.gneb_start(S.SOLVERS[solver], single_shot=True)
    p.n_shot(10)
    # K reached the library through the float setter; dt comes from the cfg (not narrowed)
    m = R.Model((12, 10, 1), (1, 0, 0), K=0.25)
    imgs, E, Rx = R.gneb_two_stage_single_shots(m, imgs0, t, 1.0, 10, solver)
    assert np.abs(np.stack([p.spins(i) for i in range(8)]) - np.stack(imgs)).max() < 1e-10
    rx, e = p.chain_rx_e()
    assert np.abs(rx - np.array(Rx)).max() < 1e-10
    assert np.abs(e - np.array(E)).max() <= 1e-11 * np.abs(E).max()
    p.stop()
    p.close()


@pytest.mark.parametrize("solver,n", [("VP", 40), ("Depondt", 10), ("Heun", 10)])
@pytest.mark.parametrize("types", TYPES)
def test_gneb_single_shots_match_reference(cfg, product, oracle, solver, n, types):
    path = cfg("solvers", n_basis_cells="12 10 1", boundary_conditions="1 0 0")
    out = []
    for lib in (product, oracle):
        x = S.Session(lib, path)
        make_chain(x, noi=8)
        for img, t in types.items():
            x.gneb_set_image_type(t, img)
        x.gneb_start(S.SOLVERS[solver], single_shot=True)
        x.n_shot(n)
        rx, e = x.chain_rx_e()
        out.append((np.stack([x.spins(i).copy() for i in range(8)]), rx, e, x.chain_max_torque(),
                    np.stack([x.effective_field(i).copy() for i in range(8)])))
        x.stop()
        x.close()
    (sp, rxp, ep, tp, fp), (so, rxo, eo, to, fo) = out
    assert np.abs(sp - so).max() < 1e-10
    assert np.abs(rxp - rxo).max() < 1e-10
    assert np.abs(ep - eo).max() <= 1e-11 * np.abs(eo).max()
    assert abs(tp - to) <= 1e-9 * to
    assert np.abs(fp - fo).max() <= 1e-9 * np.abs(fo).max()


def test_gneb_block_iterations_and_interpolation(cfg, product, oracle):
    """Simulation_GNEB_Start over amortised blocks; Chain_Get_Rx / Energy(_Interpolated) as the API returns them"""
    import ctypes
    path = cfg("solvers", n_basis_cells="10 10 1", gneb_n_iterations_amortize=5)
    res = []
    for lib in (product, oracle):
        x = S.Session(lib, path)
        make_chain(x)
        x.gneb_start(S.SOLVER_VP, n_iterations=50, n_iterations_log=50)
        n_interp = 7 + 6 * lib.Parameters_GNEB_Get_N_Energy_Interpolations(x.state, -1)
        rx, e = (ctypes.c_float * n_interp)(), (ctypes.c_float * n_interp)()
        lib.Chain_Get_Rx_Interpolated(x.state, rx, -1)
        lib.Chain_Get_Energy_Interpolated(x.state, e, -1)
        res.append((np.stack([x.spins(i).copy() for i in range(7)]), np.array(rx[:]), np.array(e[:]), x.chain_rx_e()))
        x.close()
    assert np.abs(res[0][0] - res[1][0]).max() < 1e-9
    assert np.allclose(res[0][1], res[1][1], rtol=1e-6, atol=1e-6)
    assert np.allclose(res[0][2], res[1][2], rtol=1e-6, atol=1e-3)
    assert np.abs(res[0][3][1] - res[1][3][1]).max() <= 1e-10 * np.abs(res[1][3][1]).max()


def test_gneb_barrier_golden(cfg, product):
    """core/test/test_solvers.cpp:52-88: 9 images on the 16x16 skyrmion, 2e4 VP iterations, automatic climbing/falling
    images, converge: saddle point E = -5811.5244140625, M_z = 2 * 0.96657 => barrier ~ 38.167 meV"""
    p = S.Session(product, cfg("solvers"))
    p.plus_z()
    p.skyrmion(5.0, phase=-90.0)
    p.llg_set(direct_minimization=True)
    p.llg_start(S.SOLVER_VP)  # relax the initial skyrmion
    p.chain_set_length(9)
    p.jump_to_image(8)
    p.plus_z()
    p.jump_to_image(0)
    p.transition_homogeneous(0, 8)
    p.gneb_start(S.SOLVER_VP, n_iterations=20000, n_iterations_log=20000)
    p.gneb_set_image_type_automatically()
    p.gneb_start(S.SOLVER_VP)
    rx, e = p.chain_rx_e()
    i_max = int(np.argmax(e))
    assert abs(e[i_max] - (-5811.5244140625)) < 1e-3
    assert abs(e[0] - (-5849.69140625)) < 1e-3
    assert abs(p.magnetization(i_max)[2] - 2 * 0.96657) < 1e-4
    p.close()


def test_gneb_barrier_matches_reference(cfg, product, oracle):
    """Energy barrier max E - E[0] after the same GNEB run: 1e-8 relative (BASELINE.json)"""
    path = cfg("solvers", gneb_n_iterations_amortize=10)
    barrier = []
    for lib in (product, oracle):
        x = S.Session(lib, path)
        x.plus_z()
        x.skyrmion(5.0, phase=-90.0)
        x.llg_set(direct_minimization=True)
        x.llg_start(S.SOLVER_VP)
        x.chain_set_length(7)
        x.jump_to_image(6)
        x.plus_z()
        x.jump_to_image(0)
        x.transition_homogeneous(0, 6)
        x.gneb_start(S.SOLVER_VP, n_iterations=3000, n_iterations_log=3000)
        x.gneb_set_image_type_automatically()
        x.gneb_start(S.SOLVER_VP, n_iterations=3000, n_iterations_log=3000)
        rx, e = x.chain_rx_e()
        barrier.append(e.max() - e[0])
        x.close()
    assert barrier[1] > 30.0
    assert abs(barrier[0] - barrier[1]) <= 1e-8 * abs(barrier[1])


@pytest.fixture
def serial_oracle():
    """see tests/test_parity_gpu.py::serial_oracle (the reference's atlas chart change races under OpenMP)"""
    import ctypes
    gomp = ctypes.CDLL("libgomp.so.1")
    gomp.omp_get_max_threads.restype = ctypes.c_int
    before = gomp.omp_get_max_threads()
    gomp.omp_set_num_threads(1)
    yield
    gomp.omp_set_num_threads(before)


@pytest.mark.parametrize("solver", ["VP_OSO", "LBFGS_OSO", "LBFGS_Atlas"])
@pytest.mark.parametrize("types", TYPES)
def test_gneb_minimisers_match_reference(cfg, product, oracle, serial_oracle, solver, types):
    """VP_OSO / LBFGS_OSO / LBFGS_Atlas over all images (dot products and the step limit are taken over the whole chain,
    Solver_Kernels.hpp:44-190, Solver_LBFGS_OSO.hpp:59-69): 30 single shots against the reference"""
    path = cfg("solvers", n_basis_cells="12 10 1", boundary_conditions="1 0 0")
    out = []
    for lib in (product, oracle):
        x = S.Session(lib, path)
        make_chain(x, noi=8)
        for img, t in types.items():
            x.gneb_set_image_type(t, img)
        x.gneb_start(S.SOLVERS[solver], single_shot=True)
        x.n_shot(30)
        rx, e = x.chain_rx_e()
        out.append((np.stack([x.spins(i).copy() for i in range(8)]), rx, e, x.chain_max_torque()))
        x.stop()
        x.close()
    (sp, rxp, ep, tp), (so, rxo, eo, to) = out
    assert np.abs(sp - so).max() < 1e-8
    assert np.abs(rxp - rxo).max() < 1e-8
    assert np.abs(ep - eo).max() <= 1e-9 * np.abs(eo).max()
    assert abs(tp - to) <= 1e-6 * to


@pytest.mark.parametrize("solver", ["LBFGS_Atlas", "LBFGS_OSO", "VP_OSO"])
def test_gneb_barrier_golden_minimisers(cfg, product, solver):
    """core/test/test_solvers.cpp:74-115 for the three minimisers: saddle point E = -5811.5244140625, M_z = 2 * 0.96657"""
    p = S.Session(product, cfg("solvers"))
    p.plus_z()
    p.skyrmion(5.0, phase=-90.0)
    p.llg_set(direct_minimization=True)
    p.llg_start(S.SOLVERS[solver])  # relax the initial skyrmion
    p.chain_set_length(9)
    p.jump_to_image(8)
    p.plus_z()
    p.jump_to_image(0)
    p.transition_homogeneous(0, 8)
    p.gneb_start(S.SOLVERS[solver], n_iterations=20000, n_iterations_log=20000)
    p.gneb_set_image_type_automatically()
    p.gneb_start(S.SOLVERS[solver])
    rx, e = p.chain_rx_e()
    i_max = int(np.argmax(e))
    assert abs(e[i_max] - (-5811.5244140625)) < 1e-3
    assert abs(p.magnetization(i_max)[2] - 2 * 0.96657) < 1e-4
    p.close()


def _chain_run(lib, path, options, solver, n, types=None, noi=8):
    x = S.Session(lib, path)
    make_chain(x, noi=noi)
    for img, t in (types or {}).items():
        x.gneb_set_image_type(t, img)
    L, st = x.lib, x.state
    if "ratio" in options:
        L.Parameters_GNEB_Set_Spring_Force_Ratio(st, options["ratio"], -1)
    if "shortening" in options:
        L.Parameters_GNEB_Set_Path_Shortening_Constant(st, options["shortening"], -1)
    if options.get("moving"):
        L.Parameters_GNEB_Set_Moving_Endpoints(st, True, -1)
        L.Parameters_GNEB_Set_Equilibrium_Delta_Rx(st, options.get("dl", 1.0), options.get("dr", 1.0), -1)
    if options.get("translating"):
        L.Parameters_GNEB_Set_Translating_Endpoints(st, True, -1)
    x.gneb_start(S.SOLVERS[solver], single_shot=True)
    x.n_shot(n)
    rx, e = x.chain_rx_e()
    out = (np.stack([x.spins(i).copy() for i in range(noi)]), rx, e, x.chain_max_torque())
    x.stop()
    x.close()
    return out


FORCE_OPTIONS = [
    {"ratio": 0.3},                                            # energy-weighted springs (Method_GNEB.cpp:137-170)
    {"ratio": 1.0},
    {"shortening": 0.05},                                      # path shortening (:206-233); nos * constant above |F_go|
    {"shortening": 1e-6},                                      # ... below it
    {"moving": True, "dl": 0.4, "dr": 0.7},                    # moving endpoints (:261-355)
    {"moving": True, "translating": True, "dl": 0.5, "dr": 0.5},
    {"ratio": 0.5, "shortening": 1e-4, "moving": True, "translating": True},
]


@pytest.mark.parametrize("types", TYPES)
def test_gneb_rk4_matches_restatement(cfg, product, types):
    """RK4 over all images of a chain (Solver_RK4.hpp:41-147 with the GNEB force; Method_GNEB.cpp:749 instantiates the
    combination, but the reference's Simulation_GNEB_Start (Simulation.cpp:279-303) refuses solver RK4, so the compiled
    reference cannot run it: the NumPy restatement is the oracle here, as for SIB)"""
    from oracle import restatement as R
    p = S.Session(product, cfg("solvers", n_basis_cells="12 10 1", boundary_conditions="1 0 0"))
    make_chain(p, noi=8)
    imgs0 = [p.spins(i).copy() for i in range(8)]
    t = [R.NORMAL] * 8
    for img, ty in types.items():
        p.gneb_set_image_type(ty, img)
        t[img] = ty
    p.gneb_start(S.SOLVERS["RK4"], single_shot=True)
    p.n_shot(6)
    m = R.Model((12, 10, 1), (1, 0, 0), K=0.25)
    imgs, E, Rx = R.gneb_two_stage_single_shots(m, imgs0, t, 1.0, 6, "RK4")
    assert np.abs(np.stack([p.spins(i) for i in range(8)]) - np.stack(imgs)).max() < 1e-10
    rx, e = p.chain_rx_e()
    assert np.abs(rx - np.array(Rx)).max() < 1e-10
    assert np.abs(e - np.array(E)).max() <= 1e-11 * np.abs(E).max()
    p.stop()
    p.close()


@pytest.mark.parametrize("solver,n", [("VP", 30), ("Depondt", 8)])
@pytest.mark.parametrize("options", FORCE_OPTIONS)
def test_gneb_force_options_match_reference(cfg, product, oracle, solver, n, options):
    """The optional parts of the GNEB force -- energy-weighted springs, path shortening, moving and translating endpoints --
    against the compiled reference, single shots from the same chain (one climbing image)"""
    path = cfg("solvers", n_basis_cells="12 10 1", boundary_conditions="1 0 0")
    (sp, rxp, ep, tp), (so, rxo, eo, to) = [_chain_run(lib, path, options, solver, n, {3: S.GNEB_CLIMBING}) for lib in (product, oracle)]
    assert np.abs(sp - so).max() < 1e-9, np.abs(sp - so).max()
    assert np.abs(rxp - rxo).max() < 1e-9
    assert np.abs(ep - eo).max() <= 1e-10 * np.abs(eo).max()
    assert abs(tp - to) <= 1e-8 * to
    if options.get("moving"):
        # the end images did move
        x = S.Session(product, path)
        make_chain(x, noi=8)
        assert np.abs(sp[0] - x.spins(0)).max() > 1e-6
        x.close()


@pytest.mark.parametrize("solver,n", [("VP", 12), ("Depondt", 4)])
def test_gneb_with_dipolar_interaction_matches_reference(cfg, product, oracle, solver, n):
    """GNEB over images whose Hamiltonian has the dipolar FFT convolution (Method_GNEB.cpp:99-100 evaluates every image with
    its own Hamiltonian): one convolution per image and force evaluation"""
    path = cfg("solvers", n_basis_cells="12 10 1", boundary_conditions="0 0 0", ddi_method="fft", ddi_n_periodic_images="0 0 0")
    (sp, rxp, ep, tp), (so, rxo, eo, to) = [_chain_run(lib, path, {}, solver, n, {2: S.GNEB_CLIMBING}, noi=5) for lib in (product, oracle)]
    assert np.abs(sp - so).max() < 1e-9
    assert np.abs(rxp - rxo).max() < 1e-9
    assert np.abs(ep - eo).max() <= 1e-10 * np.abs(eo).max()
    assert abs(tp - to) <= 1e-8 * to
