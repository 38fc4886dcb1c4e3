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


def test_tables_from_separate_files(product, oracle_pd, cfg, tmp_path):
    """pinned_from_file / defects_from_file: tables without a count line, read from the top of their own files"""
    pinned = tmp_path / "pinned.txt"
    pinned.write_text("0 1 1 0   0 1 0\n0 2 3 1   -1 0 0\n0 9 7 1   0 0 -1\n")
    defects = tmp_path / "defects.txt"
    defects.write_text("0 0 0 0 -1\n0 5 5 1 -1\n0 6 1 0 4\n")
    path = cfg("default", n_basis_cells="10 8 2", block=["pinned_from_file %s" % pinned, "defects_from_file %s" % defects])
    p, o = both(product, oracle_pd, path)
    np.testing.assert_array_equal(p.atom_types(), o.atom_types())
    assert sorted(o.atom_types()[o.atom_types() != 0].tolist()) == [-1, -1, 4]
    for x in (p, o):
        x.plus_z()
    np.testing.assert_array_equal(p.spins(), o.spins())
    assert (np.any(o.spins() != np.array([0.0, 0.0, 1.0]), axis=1)).sum() == 3
    for x in (p, o):
        x.close()


def test_zero_vectors_in_spin_files_become_vacancies(product, oracle_pd, cfg, tmp_path):
    """IO_Image_Read of an OVF file and of a plain column file with (near-)zero vectors: +z spins, atom type -1 (IO.cpp:265-275,
    Dataparser.cpp:39-46 of a build with defects)"""
    from tests.test_io import HAND
    rng = np.random.default_rng(5)
    v = rng.standard_normal((8, 3))
    v[2] = 0.0
    v[6] = 1e-7
    rows = "\n".join("%.17g %.17g %.17g" % tuple(r) for r in v)
    ovf_file, col_file = tmp_path / "z.ovf", tmp_path / "z.txt"
    ovf_file.write_text(HAND % (rows, rows.replace(" ", ", ")))
    col_file.write_text(rows + "\n")
    for f in (ovf_file, col_file):
        p, o = both(product, oracle_pd, cfg("default", n_basis_cells="4 2 1"))
        for x in (p, o):
            x.image_read(f, 0)
        np.testing.assert_array_equal(p.spins(), o.spins())
        np.testing.assert_array_equal(p.atom_types(), o.atom_types())
        assert (o.atom_types() == -1).sum() == 2 and np.array_equal(o.spins()[2], [0, 0, 1])
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


def test_restatement_with_defects_and_pinning_matches_the_reference(oracle_pd, cfg):
    """pins the masks of oracle/restatement.py (vacancies in the pair terms and the on-site terms, no moment at defect sites,
    pinned sites under Depondt dynamics) to the compiled reference with both options"""
    from oracle import restatement as R
    from tests.test_parity_gpu import unit_random
    n = (9, 7, 3)
    o = S.Session(oracle_pd, cfg("cubic256", n_basis_cells="%d %d %d" % n, boundary_conditions="1 0 1", external_field_magnitude="7",
                                 block=["n_defects 3", "0 4 4 0 -1", "0 7 2 1 -1", "0 3 5 2 2"]))
    types = o.atom_types()
    defect = np.zeros(o.nos, bool)
    for a, b, c in ((4, 4, 0), (7, 2, 1), (3, 5, 2)):
        defect[a + n[0] * (b + n[1] * c)] = True
    m = R.Model(n, (1, 0, 1), J=10.0, D=6.0, B=7.0, mu_s=2.0, K=1.0, atom_types=types, defect_sites=defect)
    s = unit_random(o.nos, 21)
    go, eo = o.gradient_and_energy(s)
    gr, er = m.gradient_and_energy(s)
    assert np.abs(gr - go).max() <= 1e-12 * np.abs(go).max()
    assert abs(er - eo) <= 1e-12 * abs(eo)
    o.close()
    # pinned boundary layers and one pinned site, three Depondt steps
    o = S.Session(oracle_pd, cfg("cubic256", n_basis_cells="%d %d %d" % n, boundary_conditions="1 0 1", external_field_magnitude="7",
                                 llg_temperature=0, block=["pin_na_left 2", "pin_nc_right 1", "pinning_cell", "0 0 1", "n_pinned 1", "0 5 3 1  1 0 0"]))
    a, b, c = np.meshgrid(np.arange(n[0]), np.arange(n[1]), np.arange(n[2]), indexing="ij")
    pinned = ((a < 2) | (c >= n[2] - 1) | ((a == 5) & (b == 3) & (c == 1))).transpose(2, 1, 0).reshape(-1)
    f32 = lambda x: float(np.float32(x))  # the C API narrows dt and damping to float (SURVEY.md 8c hazard 5)
    m = R.Model(n, (1, 0, 1), J=10.0, D=6.0, B=7.0, mu_s=2.0, K=1.0, dt=f32(1e-3), alpha=f32(0.3), pinned=pinned)
    o.llg_set(temperature=0.0, damping=0.3, dt=1e-3)
    o.set_spins(s)
    o.llg_start(S.SOLVER_DEPONDT, single_shot=True)
    o.n_shot(3)
    sr = s
    for _ in range(3):
        sr = m.depondt(sr)
    assert np.abs(o.spins() - s).max() > 1e-4 and np.abs(sr - o.spins()).max() < 1e-12
    assert np.array_equal(o.spins()[pinned], s[pinned])
    o.stop()
    o.close()


def test_disordered_cell_is_refused(product, cfg):
    path = cfg("default", n_basis_cells="4 4 1", block=["atom_types 1", "0 1 2.0 0.5"])
    assert not product.State_Setup(path.encode(), True)
