"""Host-side logic of the library against the reference (CPU, no GPU): input.cfg parsing, pair lists, configurations,
chain operations. These produce the inputs of the hot path and must be identical bit for bit."""
import numpy as np
import pytest

from spirit_b200 import session as S

CASES = [
    ("solvers", {}),
    ("default", {"n_basis_cells": "12 10 3"}),
    ("fd_pairs", {}),
    ("cubic256", {"n_basis_cells": "6 5 4"}),
    ("cubic256", {"n_basis_cells": "6 5 4", "n_shells_exchange": "4", "jij": "10 5 2.5 1.25", "n_shells_dmi": "3", "dij": "6 3 1", "dm_chirality": "-2"}),
    ("cubic256", {"n_basis_cells": "5 5 5", "bravais_lattice": "fcc", "n_shells_exchange": "2", "jij": "10 5", "dm_chirality": "-1"}),
    ("cubic256", {"n_basis_cells": "5 5 2", "bravais_lattice": "bcc", "n_shells_exchange": "2", "jij": "10 5"}),
    ("cubic256", {"n_basis_cells": "7 7 1", "bravais_lattice": "hex2d", "boundary_conditions": "1 1 0", "n_shells_exchange": "3", "jij": "10 5 1", "n_shells_dmi": "2", "dij": "6 1", "dm_chirality": "2"}),
    ("ddi", {"ddi_method": "none"}),
]


@pytest.mark.parametrize("preset,overrides", CASES)
def test_pair_lists_match_reference(cfg, product, oracle, preset, overrides):
    """Hamiltonian_Heisenberg::Update_Interactions + Neighbours::Get_Neighbours_in_Shells + DMI_Normal_from_Pair"""
    path = cfg(preset, **overrides)
    p, o = S.Session(product, path), S.Session(oracle, path)
    assert p.nos == o.nos
    for kind in (0, 1):
        a, b = p.pairs(kind), o.pairs(kind)
        assert np.array_equal(a[0], b[0]), "pair indices/translations, kind %d" % kind
        assert np.array_equal(a[1], b[1]), "magnitudes"
        assert np.abs(a[2] - b[2]).max(initial=0) <= 1e-15, "DMI normals"
    p.close()
    o.close()


@pytest.mark.parametrize("preset,overrides", [CASES[0], CASES[1], CASES[5], CASES[8]])
def test_configurations_match_reference(cfg, product, oracle, preset, overrides):
    """Utility::Configurations: Random (same mt19937 stream), PlusZ, Domain, Skyrmion"""
    path = cfg(preset, **overrides)
    p, o = S.Session(product, path), S.Session(oracle, path)
    assert np.array_equal(p.spins(), o.spins())  # State_Setup ends with Configuration_Random
    for x in (p, o):
        x.random()
    assert np.array_equal(p.spins(), o.spins())
    for x in (p, o):
        x.domain((0.3, -0.2, 0.9))
    assert np.abs(p.spins() - o.spins()).max() < 1e-15
    for x in (p, o):
        x.plus_z()
        x.skyrmion(4.0, order=1, phase=-90.0, pos=(1.0, -1.0, 0.0))
    assert np.abs(p.spins() - o.spins()).max() < 1e-14
    for x in (p, o):
        x.minus_z()
        x.skyrmion(3.0, order=2, phase=30.0, up_down=True, achiral=True, rl=True)
    assert np.abs(p.spins() - o.spins()).max() < 1e-14
    p.close()
    o.close()


def test_chain_operations_match_reference(cfg, product, oracle):
    """Chain_Image_to_Clipboard / Chain_Set_Length / Chain_Jump_To_Image / Transition_Homogeneous"""
    path = cfg("solvers", n_basis_cells="10 10 1")
    out = []
    for lib in (product, oracle):
        x = S.Session(lib, path)
        x.plus_z()
        x.skyrmion(3.0, phase=-90.0)
        x.chain_set_length(7)
        assert x.noi == 7
        x.jump_to_image(6)
        x.plus_z()
        x.jump_to_image(0)
        x.transition_homogeneous(0, 6)
        out.append(np.stack([x.spins(i).copy() for i in range(7)]))
        x.close()
    assert np.abs(out[0] - out[1]).max() < 1e-14


def test_setters_and_getters_round_trip(cfg, product, oracle):
    """Hamiltonian_Set_* / Get_* and Parameters_LLG_Set_* / Get_*: float narrowing as in the reference"""
    import ctypes
    path = cfg("solvers")
    vals = []
    for lib in (product, oracle):
        x = S.Session(lib, path)
        x.set_field(12.5, (0.0, 1.0, 1.0))
        x.set_anisotropy(0.7, (1.0, 1.0, 0.0))
        x.set_exchange([9.0, 1.5])
        x.set_dmi([5.0], S.CHIRALITY_NEEL)
        x.llg_set(dt=2e-3, damping=0.25, temperature=3.0, convergence=1e-7)
        mag, nrm = ctypes.c_float(), (ctypes.c_float * 3)()
        lib.Hamiltonian_Get_Field(x.state, ctypes.byref(mag), nrm, -1, -1)
        k, kn = ctypes.c_float(), (ctypes.c_float * 3)()
        lib.Hamiltonian_Get_Anisotropy(x.state, ctypes.byref(k), kn, -1, -1)
        vals.append((mag.value, list(nrm), k.value, list(kn), lib.Hamiltonian_Get_Exchange_N_Pairs(x.state, -1, -1),
                     lib.Hamiltonian_Get_DMI_N_Pairs(x.state, -1, -1), lib.Parameters_LLG_Get_Time_Step(x.state, -1, -1),
                     lib.Parameters_LLG_Get_Damping(x.state, -1, -1), lib.Parameters_LLG_Get_Temperature(x.state, -1, -1),
                     lib.Parameters_LLG_Get_Convergence(x.state, -1, -1), x.pairs(0)[1].tolist(), x.pairs(1)[2].tolist()))
        x.close()
    assert vals[0] == vals[1]


TOPOLOGY_CASES = [
    ("solvers", None, {"n_basis_cells": "6 5 1", "boundary_conditions": "0 0 0"}),
    ("solvers", None, {"n_basis_cells": "6 5 1", "boundary_conditions": "1 1 0"}),
    ("cubic256", None, {"n_basis_cells": "6 5 1", "bravais_lattice": "hex2d", "boundary_conditions": "1 0 0"}),
    ("cubic256", ["basis", "2", "0 0 0", "0.333 0.333 0.0"], {"n_basis_cells": "5 4 1", "bravais_lattice": "hex2d", "boundary_conditions": "0 0 0"}),
    ("cubic256", ["basis", "2", "0 0 0", "0.333 0.333 0.0"], {"n_basis_cells": "5 4 1", "bravais_lattice": "hex2d", "boundary_conditions": "1 1 0"}),
    ("cubic256", ["basis", "2", "0 0 0", "0.5 0.5 0.0"], {"n_basis_cells": "4 3 1", "bravais_lattice": "sc", "boundary_conditions": "1 0 0"}),
    ("cubic256", ["basis", "3", "0 0 0", "0.5 0.2 0.0", "0.2 0.6 0"], {"n_basis_cells": "4 3 1", "bravais_lattice": "sc", "boundary_conditions": "0 0 0"}),
]


@pytest.mark.parametrize("preset,block,overrides", TOPOLOGY_CASES)
def test_topological_charge_triangles_match_reference(tmp_path, product, oracle, preset, block, overrides):
    """the triangles the topological charge is summed over (Delaunay triangulation of the basis cell + boundary rule,
    Vectormath.cpp:504-631) -- host side, no device: the same set of site triples as the reference uses"""
    import ctypes
    from tests import cfgs
    path = tmp_path / "t.cfg"
    path.write_text(cfgs.render(preset, block=block, **overrides))
    p, o = S.Session(product, str(path)), S.Session(oracle, str(path))
    n_o = o.lib.Quantity_Get_Topological_Charge_Density(o.state, None, None, -1, -1)
    q, tri_o = (ctypes.c_float * n_o)(), (ctypes.c_int * (3 * n_o))()
    o.lib.Quantity_Get_Topological_Charge_Density(o.state, q, tri_o, -1, -1)
    n_p = p.lib.SpiritB200_Topology_Triangles(p.state, None, -1)
    assert n_p == n_o and n_o > 0
    tri_p = (ctypes.c_int * (3 * n_p))()
    assert p.lib.SpiritB200_Topology_Triangles(p.state, tri_p, -1) == n_p
    ref = sorted(tuple(sorted(map(int, r))) for r in np.array(tri_o).reshape(-1, 3))
    got = sorted(tuple(sorted(map(int, r))) for r in np.array(tri_p).reshape(-1, 3))
    assert got == ref
    p.close(), o.close()


ANISOTROPY_TABLES = [
    ["n_anisotropy 2", "i K Kx Ky Kz K4", "0 1.5 0 0 1 0.0", "1 0.7 1 1 0 0.25"],
    ["n_anisotropy 2", "i Ka Kb Kc", "0 0.0 0.0 2.0", "1 0.5 0.5 0.0"],
    ["n_anisotropy 1", "K4 i Kz Ky Kx", "0.3 1 0.6 0.0 0.8"],
]


@pytest.mark.parametrize("table", ANISOTROPY_TABLES)
def test_anisotropy_table_is_parsed_like_the_reference(tmp_path, product, oracle, table):
    """n_anisotropy (per-atom anisotropy table, Configparser.cpp:1352-1385, Dataparser.cpp:97-260): what the getters report
    (the first entry) equals the reference; the gradient with the whole table is compared on the GPU (tests/test_parity_gpu.py)"""
    import ctypes
    from tests import cfgs
    path = tmp_path / "a.cfg"
    path.write_text(cfgs.render("cubic256", block=["basis", "2", "0 0 0", "0.5 0.5 0.5"] + table, n_basis_cells="4 3 2"))
    vals = []
    for lib in (product, oracle):
        x = S.Session(lib, str(path))
        mag, nrm, k4 = ctypes.c_float(0), (ctypes.c_float * 3)(), ctypes.c_float(0)
        lib.Hamiltonian_Get_Anisotropy(x.state, ctypes.byref(mag), nrm, -1, -1)
        lib.Hamiltonian_Get_Cubic_Anisotropy(x.state, ctypes.byref(k4), -1, -1)
        vals.append((mag.value, tuple(nrm), k4.value))
        x.close()
    assert vals[0] == vals[1]


def test_unsupported_hamiltonian_fails_state_setup(tmp_path, product):
    """Input that asks for physics outside the library must not produce a State with a silently different Hamiltonian
    (DESIGN.md 8): State_Setup returns NULL, as the reference does for a Hamiltonian it cannot build"""
    from tests import cfgs
    for extra in (["hamiltonian gaussian"], ["n_interaction_quadruplets 1", "i j da_j db_j dc_j k da_k db_k dc_k l da_l db_l dc_l Q",
                                             "0 0 1 0 0 0 0 1 0 0 1 1 0 0 3.0"]):
        path = tmp_path / "u.cfg"
        text = cfgs.render("cubic256", n_basis_cells="4 3 2")
        if extra[0].startswith("hamiltonian"):
            text = "\n".join(l for l in text.split("\n") if not l.startswith("hamiltonian ")) + "\n"
        path.write_text(text + "\n".join(extra) + "\n")
        assert not product.State_Setup(str(path).encode(), True)
