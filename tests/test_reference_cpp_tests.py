"""The reference's own Catch2 tests of the C API (core/test/test_io.cpp, test_api.cpp), compiled from where they lie against
the reference's headers and linked against THIS library: the C ABI is a drop-in at link level and the host-side behaviour
(OVF files in every format, chains, capitalisation, plain column files, segment counts, pair files; state / chain / configuration
API) passes the reference's assertions. Runs only where the reference tree is mounted; the cases that need the device
(topological charge in test_api.cpp, test_solvers.cpp, test_physics.cpp) are GPU work and are mirrored in tests/*_gpu.py."""
import os
import shutil
import subprocess

import pytest

REF = "/root/reference/core"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GEN = os.path.join(ROOT, "oracle", "_ref", "gen")

pytestmark = pytest.mark.skipif(not (os.path.isdir(REF) and os.path.isdir(GEN)), reason="reference tree / generated headers not present")

FLAGS = ["-std=c++14", "-O1", "-w", "-DFMT_HEADER_ONLY", "-DSPIRIT_USE_KISSFFT", "-Dkiss_fft_scalar=double",
         "-I" + REF + "/include", "-I" + GEN, "-I" + GEN + "/Spirit", "-I" + REF + "/test", "-I" + REF + "/thirdparty",
         "-I" + REF + "/thirdparty/kiss_fft", "-I" + REF + "/thirdparty/kiss_fft/tools", "-I" + REF + "/thirdparty/ovf/include",
         "-I" + REF + "/thirdparty/spectra/include", "-I" + REF + "/thirdparty/Eigen"]


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("refcpp")
    subprocess.run(["g++"] + FLAGS + ["-c", REF + "/test/main.cpp", "-o", str(d / "main.o")], check=True, timeout=600)
    run = d / "run" / "core" / "test"
    run.mkdir(parents=True)
    shutil.copytree(REF + "/test/input", run / "input")
    shutil.copytree(REF + "/test/io_test_files", run / "io_test_files")
    return d


def build_and_run(workdir, name, args=()):
    exe = workdir / name
    lib = os.path.join(ROOT, "spirit_b200")
    subprocess.run(["g++"] + FLAGS + [str(workdir / "main.o"), "%s/test/%s.cpp" % (REF, name), "-o", str(exe), "-L" + lib, "-lSpirit",
                                      "-Wl,-rpath," + lib], check=True, timeout=600)
    r = subprocess.run([str(exe)] + list(args), cwd=str(workdir / "run"), capture_output=True, text=True, timeout=300)
    return r.returncode, r.stdout[-1500:]  # Catch2 reports on stdout; the library logs to stderr


def test_reference_test_io_passes_on_this_library(workdir):
    rc, tail = build_and_run(workdir, "test_io")
    assert rc == 0 and "All tests passed" in tail, tail


def test_reference_test_api_host_cases_pass_on_this_library(workdir):
    rc, tail = build_and_run(workdir, "test_api", ["~Quantities"])  # Quantities: topological charge, needs the device
    assert rc == 0 and "All tests passed" in tail, tail
