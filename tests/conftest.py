import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def product():
    """The CUDA library (spirit_b200/libSpirit.so). Missing library = hard failure, never a skip."""
    from spirit_b200 import capi
    return capi.load_product()


@pytest.fixture(scope="session")
def oracle():
    """TEST INFRASTRUCTURE: the unmodified reference CPU build (oracle/_ref/libSpirit_ref.so)."""
    from spirit_b200 import capi
    if not os.path.exists(capi.ORACLE_LIB):
        pytest.fail("oracle/_ref/libSpirit_ref.so is missing: run `make -C oracle` where /root/reference exists")
    return capi.load_oracle()


@pytest.fixture(scope="session")
def oracle_pd():
    """TEST INFRASTRUCTURE: the reference CPU build with -DSPIRIT_ENABLE_PINNING -DSPIRIT_ENABLE_DEFECTS (`make -C oracle pd`)."""
    from spirit_b200 import capi
    if not os.path.exists(capi.ORACLE_PD_LIB):
        pytest.fail("oracle/_ref/libSpirit_ref_pd.so is missing: run `make -C oracle pd` where /root/reference exists")
    return capi.load_oracle_pd()


@pytest.fixture
def cfg(tmp_path):
    """Factory writing an input.cfg into the test's tmp dir: cfg('solvers', n_basis_cells='8 8 1')"""
    from tests import cfgs

    def make(preset, **overrides):
        path = tmp_path / ("%s_%d.cfg" % (preset, len(list(tmp_path.iterdir()))))
        path.write_text(cfgs.render(preset, **overrides))
        return str(path)

    return make
