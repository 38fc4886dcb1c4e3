"""The normal variates of the thermal field as the stage kernels draw them (Philox4x32-10 + Box-Muller shaped with fp32 SFU
instructions), dumped through a probe: first four moments and the distribution function on 1.2e8 samples. The reference draws
std::normal_distribution<double> from a serial mt19937 (Method_LLG.cpp:98-108): the stream cannot be compared, the distribution
can. Also pinned: the tails up to 5 sigma, the bound of the radius (23-bit uniforms: 5.77 sigma) and the independence of the
three variates of a site."""
import ctypes
import math

import numpy as np
import pytest

from spirit_b200 import session as S

pytestmark = pytest.mark.gpu
N_SITES = 40_000_000  # x 3 variates = 1.2e8 samples


@pytest.fixture(scope="module")
def variates(product, tmp_path_factory):
    from tests import cfgs
    path = tmp_path_factory.mktemp("xi") / "xi.cfg"
    path.write_text(cfgs.render("cubic256", n_basis_cells="4 4 4", llg_seed=20006))
    p = S.Session(product, str(path))
    out = np.empty((N_SITES, 3), dtype=np.float32)
    assert product.SpiritB200_Thermal_Variates(p.state, 12345, N_SITES, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), -1) == 0
    p.close()
    return out


def test_moments(variates):
    x = variates.reshape(-1).astype(np.float64)
    n = x.size
    m1 = x.mean()
    c = x - m1
    m2, m3, m4 = (c ** 2).mean(), (c ** 3).mean(), (c ** 4).mean()
    # standard errors of the sample moments of a unit normal: 1/sqrt(n), sqrt(2/n), sqrt(6/n) (skew), sqrt(24/n) (excess kurtosis)
    assert abs(m1) < 5 / math.sqrt(n)
    assert abs(m2 - 1) < 5 * math.sqrt(2 / n)
    assert abs(m3 / m2 ** 1.5) < 5 * math.sqrt(6 / n)
    assert abs(m4 / m2 ** 2 - 3) < 5 * math.sqrt(24 / n)
    # sixth moment (15 for a normal): sensitive to the tails; its standard error is sqrt((10395 - 225) / n)
    assert abs((c ** 6).mean() - 15) < 5 * math.sqrt(10170 / n)


def test_distribution_function(variates):
    """Kolmogorov-Smirnov distance to the normal distribution function on 1.2e8 samples, through a fine histogram (4096 bins over
    +-8 sigma: the binning error of the empirical distribution function is zero at the bin edges, where it is compared)"""
    x = variates.reshape(-1)
    n = x.size
    edges = np.linspace(-8, 8, 4097)
    hist, _ = np.histogram(x, bins=edges)
    assert hist.sum() == n  # nothing beyond 8 sigma
    ecdf = np.cumsum(hist) / n
    cdf = np.array([0.5 * (1 + math.erf(e / math.sqrt(2))) for e in edges[1:]])
    d = np.abs(ecdf - cdf).max()
    # Kolmogorov: P(sqrt(n) D > 1.95) = 0.001
    assert math.sqrt(n) * d < 1.95, d


def test_tails_and_independence(variates):
    x = variates.astype(np.float64)
    n = x.size
    # two-sided tail beyond 4 and 5 sigma against the normal law (Poisson errors)
    for t in (4.0, 5.0):
        expected = n * math.erfc(t / math.sqrt(2))
        got = float((np.abs(x) > t).sum())
        assert abs(got - expected) < 5 * math.sqrt(expected) + 1, (t, got, expected)
    assert np.abs(x).max() < 5.8  # the radius is bounded by sqrt(-2 ln 2^-24) = 5.77 (llg.cuh, unit_open)
    # the three variates of a site (two share a Box-Muller radius, the third comes from the second pair) are uncorrelated, and so
    # are their squares (a shared radius would show up there if sin / cos came from the same angle uniform with an offset)
    for a, b in ((0, 1), (0, 2), (1, 2)):
        r = np.corrcoef(x[:, a], x[:, b])[0, 1]
        r2 = np.corrcoef(x[:, a] ** 2, x[:, b] ** 2)[0, 1]
        assert abs(r) < 5 / math.sqrt(len(x)), (a, b, r)
        assert abs(r2) < 5 / math.sqrt(len(x)), (a, b, r2)
    # neighbouring counters are independent
    r = np.corrcoef(x[:-1, 0], x[1:, 0])[0, 1]
    assert abs(r) < 5 / math.sqrt(len(x))
