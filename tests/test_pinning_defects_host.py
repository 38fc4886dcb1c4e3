"""Pinned sites and defects, host side (no GPU): what the configuration file and the API calls make of the lattice must equal
what the reference built with -DSPIRIT_ENABLE_PINNING -DSPIRIT_ENABLE_DEFECTS makes of it (oracle/_ref/libSpirit_ref_pd.so;
Configparser.cpp:389-441, 573-699, Geometry.cpp:50-83, 486-566, 807-832, Configurations.cpp:567-599)."""
import numpy as np

from spirit_b200 import session as S

PIN_BLOCK = ["pin_na_left 2", "pin_nb 1", "pinning_cell", "0 1 0",
             "n_pinned 2", "0 5 3 0  1 0 0", "0 6 4 1  0 0 -1"]
DEFECT_BLOCK = ["n_defects 3", "0 4 4 0 -1", "0 7 2 1 -1", "0 3 5 0 2"]


def both(product, oracle_pd, path):
    return S.Session(product, path), S.Session(oracle_pd, path)


def test_build_options_reported(product):
    assert product.Spirit_Pinning() == b"ON" and product.Spirit_Defects() == b"ON"


def test_pinning_from_config_and_configurations(product, oracle_pd, cfg):
    path = cfg("default", n_basis_cells="10 8 2", block=PIN_BLOCK)
    p, o = both(product, oracle_pd, path)
    for x in (p, o):
        x.plus_z()
    np.testing.assert_array_equal(p.spins(), o.spins())
    pinned = np.any(o.spins() != np.array([0.0, 0.0, 1.0]), axis=1)
    assert pinned.sum() == 2 * 8 * 2 + 2 * (10 - 2) * 2 + 2 - 0  # two a-layers, one b-layer on each side, two single sites
    for x in (p, o):
        x.random()
        x.skyrmion(3.0)
    # random() draws from each library's own generator: only the pinned sites must agree, and they must be untouched
    np.testing.assert_array_equal(p.spins()[pinned], o.spins()[pinned])
    for x in (p, o):
        x.close()


def test_defects_from_config(product, oracle_pd, cfg):
    path = cfg("default", n_basis_cells="10 8 2", block=DEFECT_BLOCK)
    p, o = both(product, oracle_pd, path)
    np.testing.assert_array_equal(p.atom_types(), o.atom_types())
    assert (o.atom_types() < 0).sum() == 2 and (o.atom_types() == 2).sum() == 1
    for x in (p, o):
        x.close()


def test_set_pinned_and_set_atom_type(product, oracle_pd, cfg):
    path = cfg("default", n_basis_cells="12 12 1")
    p, o = both(product, oracle_pd, path)
    for x in (p, o):
        x.plus_z()
        x.skyrmion(4.0)
        x.set_pinned(True, pos=(2, 1, 0), cylindrical=2.5)  # the sites inside the cylinder keep their present orientation
        x.set_atom_type(-1, pos=(-3, -2, 0), rect=(1.2, 0.7, -1))
        x.set_atom_type(3, pos=(4, -4, 0), spherical=1.1)
        x.minus_z()
    np.testing.assert_array_equal(p.atom_types(), o.atom_types())
    np.testing.assert_array_equal(p.spins(), o.spins())
    assert (o.atom_types() < 0).sum() > 0 and np.any(o.spins()[:, 2] > -1.0)
    for x in (p, o):
        x.set_pinned(False, pos=(2, 1, 0), cylindrical=2.5)
        x.plus_z()
    np.testing.assert_array_equal(p.spins(), o.spins())
    for x in (p, o):
        x.close()


def test_disordered_cell_is_refused(product, cfg):
    path = cfg("default", n_basis_cells="4 4 1", block=["atom_types 1", "0 1 2.0 0.5"])
    assert not product.State_Setup(path.encode(), True)
