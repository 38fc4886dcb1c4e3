"""The reference's own input files, VERBATIM, fed to both libraries: input/input.cfg (BASELINE configs[0]: 100x100x1, started like
ui-python/llg.py: +z with a skyrmion) and core/test/input/*.cfg. The files are copied next to the oracle by `make -C oracle
ref_tests` (oracle/_ref/ref_tests/run, git-ignored test artefacts that travel to the GPU box); nothing here reads /root/reference."""
import os
import shutil

import numpy as np
import pytest

from spirit_b200 import session as S
from tests.test_parity_gpu import unit_random

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUN = os.path.join(ROOT, "oracle", "_ref", "ref_tests", "run")


def sessions(tmp_path, monkeypatch, product, oracle, rel):
    """one working directory per library: the verbatim files ask for output and log files relative to the cwd"""
    src = os.path.join(RUN, rel)
    if not os.path.exists(src):
        pytest.fail("%s is missing: run `make -C oracle ref_tests` where /root/reference exists" % src)
    out = []
    for name, lib in (("product", product), ("oracle", oracle)):
        d = tmp_path / name
        (d / "output").mkdir(parents=True)
        shutil.copy(src, d / os.path.basename(rel))
        monkeypatch.chdir(d)
        out.append(S.Session(lib, str(d / os.path.basename(rel))))
    return out


def test_default_input_cfg_verbatim(tmp_path, monkeypatch, product, oracle):
    """BASELINE configs[0] exactly as shipped: gradient / energy on a random state, then the ui-python/llg.py scenario
    (PlusZ + skyrmion of radius 5) for 200 Depondt iterations: spins < 1e-10, energy 1e-11"""
    p, o = sessions(tmp_path, monkeypatch, product, oracle, "input/input.cfg")
    assert p.nos == o.nos == 100 * 100
    s = unit_random(p.nos, 2)
    gp, ep = p.gradient_and_energy(s)
    go, eo = o.gradient_and_energy(s)
    assert np.abs(gp - go).max() <= 1e-12 * np.abs(go).max()
    assert abs(ep - eo) <= 1e-11 * abs(eo)
    for x in (p, o):
        x.llg_no_output()
        x.plus_z()
        x.skyrmion(5.0, phase=-90.0)
        x.llg_start(S.SOLVER_DEPONDT, n_iterations=200, n_iterations_log=200)
    assert p.step_variant(S.SOLVER_DEPONDT) == 2  # the 2-D lattice runs the fused kernel
    assert np.abs(p.spins() - o.spins()).max() < 1e-10
    assert abs(p.energy() - o.energy()) <= 1e-11 * abs(o.energy())
    assert abs(p.max_torque() - o.max_torque()) <= 1e-9 * o.max_torque()
    p.close(), o.close()


@pytest.mark.parametrize("name", ["api.cfg", "fd_neighbours.cfg", "fd_pairs.cfg", "physics_ddi.cfg", "physics_larmor.cfg", "solvers.cfg"])
def test_reference_test_inputs_verbatim(tmp_path, monkeypatch, product, oracle, name):
    """core/test/input/*.cfg: gradient / energy on a random state and 5 single Depondt shots from it"""
    p, o = sessions(tmp_path, monkeypatch, product, oracle, os.path.join("core", "test", "input", name))
    assert p.nos == o.nos
    s = unit_random(p.nos, 4)
    gp, ep = p.gradient_and_energy(s)
    go, eo = o.gradient_and_energy(s)
    assert np.abs(gp - go).max() <= 1e-12 * max(np.abs(go).max(), 1e-300)
    assert abs(ep - eo) <= 1e-11 * max(abs(eo), 1e-300)
    if name == "api.cfg":
        # the API-test input defines no interaction at all (gradient and energy are zero); the reference terminates the process
        # when a simulation is started on it, so the comparison ends with the (zero) gradient
        assert np.abs(go).max() == 0.0 and eo == 0.0
        p.close(), o.close()
        return
    for x in (p, o):
        x.llg_no_output()
        x.set_spins(s)
        x.llg_start(S.SOLVER_DEPONDT, single_shot=True)
        for _ in range(5):
            x.single_shot()
    assert np.abs(p.spins() - o.spins()).max() < 1e-10
    for x in (p, o):
        x.stop()
        x.close()


def test_gaussian_input_is_refused(tmp_path, monkeypatch, product):
    """fd_gaussian.cfg asks for the Gaussian test Hamiltonian, which is outside the hot path: State_Setup fails instead of
    running other physics"""
    src = os.path.join(RUN, "core", "test", "input", "fd_gaussian.cfg")
    monkeypatch.chdir(tmp_path)
    assert not product.State_Setup(src.encode(), True)
