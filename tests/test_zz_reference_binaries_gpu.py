"""The reference's own Catch2 tests of the C API (core/test/test_solvers.cpp, test_api.cpp, test_io.cpp), compiled by
`make -C oracle ref_tests` from the reference tree and linked against THIS library, run on the GPU: the reference's
assertions (all eight solvers relax the skyrmion to its golden energy, the GNEB saddle point, topological charge +-1, ...)
on the CUDA path. The binaries live in oracle/_ref/ref_tests (built where the reference is mounted, shipped with the
snapshot). They passed on the B200 at the end of round 1 (GPUTEST_r01: 3 xpassed) and are plain tests since."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RT = os.path.join(ROOT, "oracle", "_ref", "ref_tests")

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(os.path.join(RT, "test_solvers")), reason="oracle/_ref/ref_tests not built")]


@pytest.mark.parametrize("name", ["test_io", "test_api", "test_solvers"])
def test_reference_catch2_binary_on_the_gpu_library(name):
    r = subprocess.run([os.path.join(RT, name)], cwd=os.path.join(RT, "run"), capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0 and "All tests passed" in r.stdout[-2000:], r.stdout[-3000:]
