"""The reference's own Python package (core/python/spirit, ctypes) on top of this library, unchanged: every module must
import (each resolves all of its symbols at import time) and a host-side session must run through it. Runs only where the
reference tree is mounted (/root/reference: this container); the GPU box does not have it."""
import os
import shutil
import subprocess
import sys

import pytest

REF_PKG = "/root/reference/core/python/spirit"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir(REF_PKG), reason="reference tree not mounted")

DRIVER = r'''
import importlib, os, sys, numpy as np
mods = ["state", "system", "geometry", "configuration", "chain", "transition", "simulation", "hamiltonian", "quantities",
        "constants", "log", "io", "htst", "parameters.llg", "parameters.gneb", "parameters.mc", "parameters.mmf", "parameters.ema"]
for m in mods:
    importlib.import_module("spirit." + m)
from spirit import state, system, geometry, configuration, chain, io, log, constants, hamiltonian
from spirit.parameters import llg, gneb
cfg, out = sys.argv[1], sys.argv[2]
with state.State(cfg, quiet=True) as p:
    assert system.get_nos(p) == 35
    configuration.plus_z(p)
    configuration.skyrmion(p, 2.0, phase=-90.0)
    s = system.get_spin_directions(p)
    assert s.shape == (35, 3) and abs(np.linalg.norm(s, axis=1) - 1).max() < 1e-12 and s[:, 2].min() < 0
    assert list(geometry.get_n_cells(p)) == [7, 5, 1]
    llg.set_damping(p, 0.25)
    assert abs(llg.get_damping(p) - 0.25) < 1e-7
    hamiltonian.set_field(p, 3.0, [0, 0, 1])
    io.image_write(p, os.path.join(out, "a.ovf"), io.FILEFORMAT_OVF_TEXT, "from the python package")
    keep = s.copy()
    configuration.minus_z(p)
    io.image_read(p, os.path.join(out, "a.ovf"))
    assert abs(system.get_spin_directions(p) - keep).max() < 1e-11
    chain.image_to_clipboard(p)
    chain.set_length(p, 3)
    assert chain.get_noi(p) == 3
    io.chain_write(p, os.path.join(out, "c.ovf"))
    assert io.n_images_in_file(p, os.path.join(out, "c.ovf")) == 3
    log.send(p, log.LEVEL_INFO, log.SENDER_UI, "hello from the package")
    assert log.get_n_entries(p) > 0
    assert abs(constants.mu_B - 0.057883817555) < 1e-9
print("PYTHON_PACKAGE_OK")
'''


def test_reference_python_package_runs_on_this_library(tmp_path, cfg, product):
    pkg = tmp_path / "site" / "spirit"
    shutil.copytree(REF_PKG, pkg)
    shutil.copy(os.path.join(ROOT, "spirit_b200", "libSpirit.so"), pkg / "libSpirit.so")
    # the two files the reference's CMake writes next to the package (core/CMakeLists.txt:531-545)
    (pkg / "scalar.py").write_text("import ctypes\nscalar = ctypes.c_double\n")
    (pkg / "version.py").write_text('version_major = 2\nversion_minor = 2\nversion_patch = 0\nversion = "2.2.0"\nrevision = "spirit_b200"\n'
                                    'version_full = "2.2.0 (spirit_b200)"\ncompiler = "nvcc"\ncompiler_version = ""\ncompiler_full = "nvcc"\n'
                                    'scalartype = "double"\npinning = "OFF"\ndefects = "OFF"\ncuda = "ON"\nopenmp = "OFF"\nthreads = "OFF"\nfftw = "OFF"\n')
    if not (pkg / "__init__.py").exists():
        (pkg / "__init__.py").write_text("")
    driver = tmp_path / "driver.py"
    driver.write_text(DRIVER)
    out = tmp_path / "out"
    out.mkdir()
    env = dict(os.environ, PYTHONPATH=str(tmp_path / "site"))
    r = subprocess.run([sys.executable, str(driver), cfg("solvers", n_basis_cells="7 5 1"), str(out)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "PYTHON_PACKAGE_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])


REF_TESTS = "/root/reference/core/python/test"
# the reference's own unittest files that need no compute (the others -- system, quantities -- evaluate energies /
# magnetisation and need the device; simulation.py only checks that the calls return)
HOST_ONLY = ["state", "configuration", "constants", "geometry", "log", "parameters", "chain", "transition", "hamiltonian", "io_test", "simulation"]


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="reference tree not mounted")
@pytest.mark.parametrize("name", HOST_ONLY)
def test_reference_python_unittests_pass_on_this_library(tmp_path, name):
    """core/python/test/<name>.py, unchanged, with this library behind the unchanged package"""
    core = tmp_path / "core"
    (core / "python").mkdir(parents=True)
    shutil.copytree(REF_PKG, core / "python" / "spirit")
    shutil.copytree(REF_TESTS, core / "python" / "test")
    shutil.copytree("/root/reference/core/test/input", core / "test" / "input")
    pkg = core / "python" / "spirit"
    shutil.copy(os.path.join(ROOT, "spirit_b200", "libSpirit.so"), pkg / "libSpirit.so")
    (pkg / "scalar.py").write_text("import ctypes\nscalar = ctypes.c_double\n")
    (pkg / "version.py").write_text('version_major = 2\nversion_minor = 2\nversion_patch = 0\nversion = "2.2.0"\nrevision = "spirit_b200"\n'
                                    'version_full = "2.2.0 (spirit_b200)"\ncompiler = "nvcc"\ncompiler_version = ""\ncompiler_full = "nvcc"\n'
                                    'scalartype = "double"\npinning = "OFF"\ndefects = "OFF"\ncuda = "ON"\nopenmp = "OFF"\nthreads = "OFF"\nfftw = "OFF"\n')
    if not (pkg / "__init__.py").exists():
        (pkg / "__init__.py").write_text("")
    r = subprocess.run([sys.executable, str(core / "python" / "test" / (name + ".py"))], cwd=str(core), capture_output=True, text=True, timeout=300)
    tail = (r.stdout + r.stderr)[-1500:]
    assert r.returncode == 0 and "OK" in tail.splitlines()[-1], tail
