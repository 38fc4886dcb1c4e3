import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, tempfile
from spirit_b200 import capi, session as S
from tests import cfgs
product = capi.load_product()
d = tempfile.mkdtemp(); path = os.path.join(d, "t.cfg")
open(path, "w").write(cfgs.render("fd_pairs", pairs=["i j da db dc Jij"], n_basis_cells="64 64 1", external_field_magnitude="10",
               llg_temperature="10", llg_damping="0.3", llg_dt="1e-3", llg_n_iterations_amortize="100"))
for solver in ("Depondt", "SIB", "Heun"):
    p = S.Session(product, path)
    p.plus_z()
    p.llg_start(S.SOLVERS[solver], n_iterations=6000, n_iterations_log=6000)
    means = []
    for _ in range(12):
        p.llg_start(S.SOLVERS[solver], n_iterations=500, n_iterations_log=500)
        means.append(p.spins()[:, 2].mean())
    print(solver, "%.6f" % np.mean(means), " ".join("%.4f" % m for m in means), flush=True)
    p.close()
