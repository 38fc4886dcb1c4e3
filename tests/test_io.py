"""OVF files (Spirit/IO.h, SURVEY.md 8f rank 2: the data format either side of the hot path) against the reference's reader
and writer (its bundled ovf library): what one library writes the other must read back to the same spins, in every format,
for single images, appended segments and chains. CPU only -- the configurations are the host copies of the API."""
import numpy as np
import pytest

from spirit_b200 import session as S

BIN, BIN4, BIN8, TEXT, CSV = 0, 1, 2, 3, 4
FORMATS = [(BIN, 3e-16), (BIN4, 1e-7), (BIN8, 3e-16), (TEXT, 1e-11), (CSV, 1e-11)]  # (the readers normalise what they read)
LATTICES = [("solvers", {"n_basis_cells": "7 5 1"}), ("cubic256", {"n_basis_cells": "6 5 4"}), ("ddi", {"ddi_method": "none", "n_basis_cells": "4 3 2"})]


def pair(cfg, product, oracle, preset="solvers", **over):
    path = cfg(preset, **over)
    return S.Session(product, path), S.Session(oracle, path)


@pytest.mark.parametrize("fmt,tol", FORMATS)
@pytest.mark.parametrize("preset,over", LATTICES)
def test_image_written_here_is_read_by_the_reference_and_back(cfg, product, oracle, tmp_path, preset, over, fmt, tol):
    p, o = pair(cfg, product, oracle, preset, **over)
    p.random()
    want = p.spins().copy()
    ours, theirs = tmp_path / "ours.ovf", tmp_path / "theirs.ovf"
    p.image_write(ours, fmt, "written by spirit_b200")
    assert p.n_images_in_file(ours) == 1 and o.n_images_in_file(ours) == 1
    o.plus_z()
    o.image_read(ours)
    assert np.abs(o.spins() - want).max() <= tol
    # the other direction: the reference writes, this library reads
    o.image_write(theirs, fmt, "written by the reference")
    written = o.spins().copy()
    p.plus_z()
    p.image_read(theirs)
    assert np.abs(p.spins() - written).max() <= tol
    # and both readers agree on both files to the last bit
    for path in (ours, theirs):
        p.plus_z(), o.plus_z()
        p.image_read(path), o.image_read(path)
        assert np.array_equal(p.spins(), o.spins())
    p.close(), o.close()


def test_append_counts_segments_in_files_of_either_library(cfg, product, oracle, tmp_path):
    p, o = pair(cfg, product, oracle)
    states = []
    mixed = tmp_path / "mixed.ovf"
    for k, (who, fmt) in enumerate([(p, BIN8), (o, TEXT), (p, CSV), (o, BIN4), (p, BIN4), (o, BIN8), (p, TEXT)]):
        who.random()
        states.append(who.spins().copy())
        who.image_append(mixed, fmt, "segment %d" % k)  # the first append creates the file
        assert p.n_images_in_file(mixed) == k + 1 and o.n_images_in_file(mixed) == k + 1
    for k, want in enumerate(states):
        for who in (p, o):
            who.plus_z()
            who.image_read(mixed, idx_image_infile=k)
            assert np.abs(who.spins() - want).max() <= 1e-7
        assert np.array_equal(p.spins(), o.spins())
    # an image write replaces the file
    p.image_write(mixed, BIN8)
    assert o.n_images_in_file(mixed) == 1
    p.close(), o.close()


@pytest.mark.parametrize("fmt", [TEXT, BIN8])
def test_chain_files(cfg, product, oracle, tmp_path, fmt):
    tol = 1e-11 if fmt == TEXT else 3e-16  # 12 decimals in text files; readers normalise what they read (last bit)
    p, o = pair(cfg, product, oracle)
    for x in (p, o):
        x.chain_set_length(4)
    images = []
    for i in range(4):
        p.jump_to_image(i)
        p.random()
        images.append(p.spins().copy())
    ours = tmp_path / "chain.ovf"
    p.chain_write(ours, fmt, "a chain")
    assert o.n_images_in_file(ours) == 4
    o.chain_read(ours)
    assert o.noi == 4
    for i in range(4):
        assert np.abs(o.spins(i) - images[i]).max() <= tol
    # the reference writes the chain, a one-image state here grows to hold it (IO.cpp:511-520)
    theirs = tmp_path / "chain_ref.ovf"
    o.chain_write(theirs, fmt, "reference chain")
    q = S.Session(product, cfg("solvers"))
    assert q.noi == 1
    q.chain_read(theirs)
    assert q.noi == 4
    for i in range(4):
        assert np.abs(q.spins(i) - o.spins(i)).max() <= tol
    # a sub-range of the file, and appending a chain to a file
    r, ro = pair(cfg, product, oracle)
    for x in (r, ro):
        x.chain_read(theirs, start=1, end=2)
        assert x.noi == 2
    for i in range(2):
        assert np.array_equal(r.spins(i), ro.spins(i)) and np.abs(r.spins(i) - o.spins(i + 1)).max() <= tol
    p.chain_append(ours, fmt, "again")
    assert p.n_images_in_file(ours) == 8 and o.n_images_in_file(ours) == 8
    for x in (p, o, q, r, ro):
        x.close()


def test_reading_a_file_of_another_size_or_shape(cfg, product, oracle, tmp_path):
    """more / fewer rows than spins: the common part is read (IO.cpp:233-251); the rest keeps its values, all are normalised"""
    small_p, small_o = pair(cfg, product, oracle, n_basis_cells="4 4 1")
    big_p, big_o = pair(cfg, product, oracle, n_basis_cells="6 5 1")
    small_p.random(), big_p.random()
    f_small, f_big = tmp_path / "small.ovf", tmp_path / "big.ovf"
    small_p.image_write(f_small, BIN8), big_p.image_write(f_big, TEXT)
    for a, b, f in ((big_p, big_o, f_small), (small_p, small_o, f_big)):
        a.plus_z(), b.plus_z()
        a.image_read(f), b.image_read(f)
        assert np.array_equal(a.spins(), b.spins())
        assert np.abs(np.linalg.norm(a.spins(), axis=1) - 1).max() < 1e-15
    # not an OVF file, a missing file, a segment that does not exist: nothing changes, nothing is thrown
    junk = tmp_path / "junk.ovf"
    junk.write_text("1 2 3\n4 5 6\n")
    before = small_p.spins().copy()
    assert small_p.n_images_in_file(junk) == -1 and small_o.n_images_in_file(junk) == -1
    small_p.image_read(tmp_path / "missing.ovf")
    small_p.image_read(f_small, idx_image_infile=5)
    assert np.array_equal(small_p.spins(), before)
    for x in (small_p, small_o, big_p, big_o):
        x.close()


def test_header_of_a_written_file(cfg, product, tmp_path):
    """the keywords a third-party OVF 2.0 reader needs, with the reference's conventions (basis atoms folded into xnodes, nm)"""
    p = S.Session(product, cfg("ddi", ddi_method="none", n_basis_cells="4 3 2"))
    f = tmp_path / "h.ovf"
    p.image_write(f, TEXT, "a remark")
    text = f.read_text()
    lines = [l.strip() for l in text.splitlines()]
    assert lines[0] == "# OOMMF OVF 2.0" and "# Segment count: 000001" in lines
    for want in ("# Begin: Segment", "# Begin: Header", "# Desc: a remark", "# valuedim: 3   ## field dimensionality",
                 "# valueunits: none none none", "# valuelabels: spin_x spin_y spin_z", "# meshunit: nm", "# meshtype: rectangular",
                 "# xnodes: 8", "# ynodes: 3", "# znodes: 2", "# End: Header", "# Begin: Data Text", "# End: Data Text", "# End: Segment"):
        assert want in lines, want
    data = [l for l in text.splitlines() if l and not l.startswith("#")]
    assert len(data) == p.nos and all(len(l) == 66 for l in data)  # three columns of width 22
    p.close()


@pytest.mark.gpu
@pytest.mark.parametrize("solver", [S.SOLVER_DEPONDT, S.SOLVER_VP])
def test_llg_output_files_match_the_reference(cfg, product, oracle, tmp_path, solver):
    """llg_output_*: same file names, spins within the single-step tolerance, same energy tables (Method_LLG.cpp:310-500)"""
    import os
    folders = {}
    for name, lib in (("product", product), ("oracle", oracle)):
        out = tmp_path / name
        out.mkdir()
        folders[name] = out
        path = cfg("solvers", llg_output_any=1, llg_output_initial=1, llg_output_final=1, llg_output_configuration_step=1,
                   llg_output_configuration_archive=1, llg_output_energy_step=1, llg_output_energy_archive=1,
                   llg_output_energy_divide_by_nspins=0, llg_output_configuration_filetype=3, llg_output_folder=str(out),
                   output_file_tag="run", llg_n_iterations=40, llg_n_iterations_log=10, llg_force_convergence="1e-14")
        x = S.Session(lib, path, quiet=False)  # (a quiet state writes no files, State.cpp:60-75)
        x.plus_z()
        x.skyrmion(5.0, phase=-90.0)
        x.llg_start(solver, n_iterations=40, n_iterations_log=10)
        x.close()
    names = {k: sorted(os.listdir(v)) for k, v in folders.items()}
    assert names["product"] == names["oracle"] and len(names["product"]) >= 12, names
    reader_p, reader_o = S.Session(product, cfg("solvers")), S.Session(oracle, cfg("solvers"))
    for f in names["product"]:
        fp, fo = folders["product"] / f, folders["oracle"] / f
        if f.endswith(".ovf"):
            n = reader_o.n_images_in_file(fo)
            assert reader_p.n_images_in_file(fp) == n and reader_o.n_images_in_file(fp) == n
            for k in range(n):
                reader_p.image_read(fp, k), reader_o.image_read(fo, k)
                assert np.abs(reader_p.spins() - reader_o.spins()).max() < 1e-9, (f, k)
        else:
            tp, to = fp.read_text().splitlines(), fo.read_text().splitlines()
            assert len(tp) == len(to) and len(tp) >= 2, f
            assert tp[0].split() == to[0].split(), f  # column titles
            for lp, lo in zip(tp[1:], to[1:]):
                vp, vo = [float(v) for v in lp.split()], [float(v) for v in lo.split()]
                assert vp[0] == vo[0] and abs(vp[1] - vo[1]) <= 1e-8 * max(1.0, abs(vo[1])), (f, lp, lo)  # iteration, E_tot
                assert len(vp) == len(vo)
    reader_p.close(), reader_o.close()


@pytest.mark.gpu
def test_gneb_output_files_match_the_reference(cfg, product, oracle, tmp_path):
    """gneb_output_*: chain files (one OVF segment per image) and chain energy tables (Method_GNEB.cpp:600-715)"""
    import os
    from tests.test_gneb_gpu import make_chain
    folders = {}
    for name, lib in (("product", product), ("oracle", oracle)):
        out = tmp_path / name
        out.mkdir()
        folders[name] = out
        path = cfg("solvers", n_basis_cells="10 10 1", gneb_output_any=1, gneb_output_initial=1, gneb_output_final=1,
                   gneb_output_chain_step=1, gneb_output_energies_step=1, gneb_output_energies_divide_by_nspins=0,
                   gneb_output_chain_filetype=3, gneb_output_folder=str(out), output_file_tag="path",
                   gneb_n_iterations=30, gneb_n_iterations_log=10, gneb_force_convergence="1e-14")
        x = S.Session(lib, path, quiet=False)
        make_chain(x, noi=5)
        x.gneb_start(S.SOLVER_VP, n_iterations=30, n_iterations_log=10)
        x.close()
    names = {k: sorted(os.listdir(v)) for k, v in folders.items()}
    assert names["product"] == names["oracle"] and len(names["product"]) >= 10, names
    reader_p, reader_o = S.Session(product, cfg("solvers", n_basis_cells="10 10 1")), S.Session(oracle, cfg("solvers", n_basis_cells="10 10 1"))
    for f in names["product"]:
        fp, fo = folders["product"] / f, folders["oracle"] / f
        if f.endswith(".ovf"):
            assert reader_p.n_images_in_file(fp) == 5 and reader_o.n_images_in_file(fo) == 5 and reader_o.n_images_in_file(fp) == 5
            for k in range(5):
                reader_p.image_read(fp, k), reader_o.image_read(fo, k)
                assert np.abs(reader_p.spins() - reader_o.spins()).max() < 1e-9, (f, k)
        else:
            tp, to = fp.read_text().splitlines(), fo.read_text().splitlines()
            assert len(tp) == len(to) == 3 + 5, f  # separator, titles, separator, one line per image
            assert tp[0] == to[0] and tp[1].split("|")[:4] == to[1].split("|")[:4], f
            for lp, lo in zip(tp[3:], to[3:]):
                vp, vo = [float(v) for v in lp.split()], [float(v) for v in lo.split()]
                assert vp[0] == vo[0] and abs(vp[1] - vo[1]) <= 1e-8 and abs(vp[2] - vo[2]) <= 1e-8 * max(1.0, abs(vo[2])), (f, lp, lo)
    reader_p.close(), reader_o.close()


@pytest.mark.gpu
def test_energy_files_match_the_reference(cfg, product, oracle, tmp_path):
    """IO_Image_Write_Energy, IO_Image_Write_Energy_per_Spin, IO_Chain_Write_Energies (IO.cpp:851-992)"""
    p, o = pair(cfg, product, oracle, "cubic256", n_basis_cells="6 5 4", llg_temperature=0)
    s = p.spins().copy()  # State_Setup ends with the same random configuration in both
    assert np.array_equal(s, o.spins())
    files = {}
    for name, x in (("p", p), ("o", o)):
        x.update_data()
        e, es, ce = tmp_path / (name + "_E.txt"), tmp_path / (name + "_Es.ovf"), tmp_path / (name + "_chain.txt")
        x.lib.IO_Image_Write_Energy(x.state, str(e).encode(), -1, -1)
        x.lib.IO_Image_Write_Energy_per_Spin(x.state, str(es).encode(), BIN8, -1, -1)
        x.chain_update_data()
        x.lib.IO_Chain_Write_Energies(x.state, str(ce).encode(), -1)
        files[name] = (e, es, ce)
    for k in (0, 2):  # the two tables: identical titles, numbers to 1e-9 relative
        tp, to = files["p"][k].read_text().splitlines(), files["o"][k].read_text().splitlines()
        assert len(tp) == len(to) and tp[:3] == to[:3]
        for lp, lo in zip(tp[3:], to[3:]):
            vp = np.array([float(v) for v in lp.replace("|", " ").split()])
            vo = np.array([float(v) for v in lo.replace("|", " ").split()])
            assert vp.shape == vo.shape and np.abs(vp - vo).max() <= 1e-9 * max(1.0, np.abs(vo).max())
    # per-spin energies: an OVF field with 1 + n_terms columns
    hp, ho = files["p"][1].read_bytes(), files["o"][1].read_bytes()
    def block(raw):
        head, _, rest = raw.partition(b"# Begin: Data Binary 8\n")
        labels = [l for l in head.decode().splitlines() if l.startswith("# valuelabels:")][0].split(":")[1].split()
        data = np.frombuffer(rest[8:8 + 8 * len(labels) * 120], dtype="<f8").reshape(120, len(labels))
        return labels, data
    lp, dp = block(hp)
    lo, do = block(ho)
    assert lp == lo and lp[0] == "Total" and len(lp) >= 4
    assert np.abs(dp - do).max() <= 1e-12 * np.abs(do).max()
    p.close(), o.close()


HAND = """# OOMMF OVF 2.0
#
# Segment count: 000002
#
# Begin: Segment
# Begin: Header
#
# Title: by hand
# Desc: first line of the description
# Desc: second line
## a remark line inside the header
# valuedim: 3   ## field dimensionality
# valueunits: none none none
# valuelabels: spin_x spin_y spin_z
# meshunit: nm
# xmin: 0
# ymin: 0
# zmin: 0
# xmax: 1
# ymax: 1
# zmax: 1
# meshtype: rectangular
# xbase: 0
# ybase: 0
# zbase: 0
# xstepsize: 1
# ystepsize: 1
# zstepsize: 1
# xnodes: 4
# ynodes: 2
# znodes: 1
# End: Header
# Begin: Data Text
%s
# End: Data Text
# End: Segment
#
# Begin: Segment
# Begin: Header
# Title: second segment, CSV, upper-case keywords are not required by the format
# valuedim: 3
# valueunits: none none none
# valuelabels: spin_x spin_y spin_z
# meshunit: nm
# xmin: 0
# ymin: 0
# zmin: 0
# xmax: 1
# ymax: 1
# zmax: 1
# meshtype: rectangular
# xbase: 0
# ybase: 0
# zbase: 0
# xstepsize: 1
# ystepsize: 1
# zstepsize: 1
# xnodes: 4
# ynodes: 2
# znodes: 1
# End: Header
# Begin: Data CSV
%s
# End: Data CSV
# End: Segment
"""


def test_hand_written_files_are_read_like_the_reference_reads_them(cfg, product, oracle, tmp_path):
    """text and CSV blocks with irregular spacing, scientific notation, unnormalised and zero vectors, multi-line
    descriptions, remark lines; CRLF line ends"""
    rng = np.random.default_rng(3)
    v = rng.standard_normal((8, 3)) * 3.0
    v[5] = 0.0  # a zero vector: read as +z (IO.cpp:267-270)
    v = np.array([[float("%.9e" % b) if i == 1 else b for i, b in enumerate(row)] for row in v])  # column 2 is written with 10 digits
    text = "\n".join("  %.17g   %.9e\t%r" % tuple(float(x) for x in row) for row in v)
    csv = "\n".join("%.17g, %.17g ,%.17g," % tuple(row[::-1]) for row in v)
    p, o = pair(cfg, product, oracle, n_basis_cells="4 2 1")
    for name, content in (("unix.ovf", HAND % (text, csv)), ("dos.ovf", (HAND % (text, csv)).replace("\n", "\r\n"))):
        f = tmp_path / name
        f.write_bytes(content.encode())
        assert p.n_images_in_file(f) == o.n_images_in_file(f) == 2, name
        for k in range(2):
            p.plus_z(), o.plus_z()
            p.image_read(f, k), o.image_read(f, k)
            got, ref = p.spins(), o.spins()
            assert np.array_equal(got, ref), (name, k)
            want = (v if k == 0 else v[:, ::-1]).copy()
            want[5] = (0, 0, 1)
            want /= np.linalg.norm(want, axis=1)[:, None]
            assert np.abs(got - want).max() < 1e-9
    p.close(), o.close()


@pytest.mark.parametrize("preset,over", [("solvers", {}), ("cubic256", {"n_basis_cells": "4 4 4", "n_shells_exchange": "3", "jij": "10 5 2.5",
                                                                      "n_shells_dmi": "2", "dij": "6 3"}), ("default", {"n_basis_cells": "6 5 2"}),
                                         ("ddi", {"ddi_method": "none"})])
def test_neighbour_files_are_identical_to_the_reference(cfg, product, oracle, tmp_path, preset, over):
    """IO_Image_Write_Neighbours_Exchange / _DMI: the pair lists the stencil kernels work on, byte for byte"""
    p, o = pair(cfg, product, oracle, preset, **over)
    for kind in ("Exchange", "DMI"):
        fp, fo = tmp_path / ("p_%s.txt" % kind), tmp_path / ("o_%s.txt" % kind)
        getattr(p.lib, "IO_Image_Write_Neighbours_" + kind)(p.state, str(fp).encode(), -1, -1)
        getattr(o.lib, "IO_Image_Write_Neighbours_" + kind)(o.state, str(fo).encode(), -1, -1)
        assert fp.read_text() == fo.read_text(), kind
        assert len(fp.read_text().splitlines()) >= 2
    p.close(), o.close()


def test_system_from_config_matches_the_reference(cfg, product, oracle):
    """IO_System_From_Config (IO.cpp:36-85): the image takes geometry, Hamiltonian and parameters of another input file"""
    first = cfg("solvers", n_basis_cells="6 6 1")
    second = cfg("cubic256", n_basis_cells="4 3 3", n_shells_exchange="2", jij="7 2", llg_temperature="0", llg_seed="77")
    wrong = cfg("solvers", n_basis_cells="5 5 1")
    p, o = S.Session(product, first), S.Session(oracle, first)
    for x in (p, o):
        assert x.lib.IO_System_From_Config(x.state, wrong.encode(), -1, -1) == 0  # another number of spins: refused
        assert x.lib.IO_System_From_Config(x.state, second.encode(), -1, -1) == 1
    assert p.nos == o.nos == 36
    for kind in (0, 1):
        a, b = p.pairs(kind), o.pairs(kind)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(p.spins(), o.spins())  # the random configuration of the new system's generator
    assert abs(np.linalg.norm(p.spins(), axis=1) - 1).max() < 1e-15
    p.close(), o.close()


def test_random_sequences_of_writes_and_appends_round_trip(cfg, product, oracle, tmp_path):
    """property test: any sequence of write / append calls in any format leaves a file whose segments both libraries read back
    as the configurations that were written (tolerance of the format)"""
    from hypothesis import given, settings, strategies as st
    p, o = pair(cfg, product, oracle, n_basis_cells="5 3 2")
    tol = {BIN: 3e-16, BIN4: 1e-7, BIN8: 3e-16, TEXT: 1e-11, CSV: 1e-11}
    counter = [0]

    @settings(max_examples=25, deadline=None)
    @given(st.lists(st.tuples(st.booleans(), st.sampled_from([BIN, BIN4, BIN8, TEXT, CSV]), st.integers(0, 2 ** 31 - 1)), min_size=1, max_size=6))
    def run(ops):
        counter[0] += 1
        f = tmp_path / ("seq_%d.ovf" % counter[0])
        expected = []
        for append, fmt, seed in ops:
            rng = np.random.default_rng(seed)
            s = rng.standard_normal((p.nos, 3))
            s /= np.linalg.norm(s, axis=1)[:, None]
            p.set_spins(s)
            if append:
                p.image_append(f, fmt, "seed %d" % seed)
                expected.append((p.spins().copy(), tol[fmt]))
            else:
                p.image_write(f, fmt, "seed %d" % seed)
                expected = [(p.spins().copy(), tol[fmt])]
        assert p.n_images_in_file(f) == len(expected) == o.n_images_in_file(f)
        for k, (want, t) in enumerate(expected):
            p.plus_z(), o.plus_z()
            p.image_read(f, k), o.image_read(f, k)
            assert np.abs(p.spins() - want).max() <= t and np.array_equal(p.spins(), o.spins())

    run()
    p.close(), o.close()
