#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libSpirit_ref.so, built by oracle/Makefile
from /root/reference). Run in the build container: python tests/golden/make_golden.py
Each case: the input spins, the reference's gradient + energy, the per-term energies, and the spins / energy / max
torque after n Simulation_SingleShot calls for every in-scope solver. GNEB: a 7-image chain after 60 VP single shots."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from spirit_b200 import capi, session as S  # noqa: E402
from tests import cfgs  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

LLG_CASES = {
    # name: (preset, overrides, anisotropy K, cubic K4)   -- K, K4 float-representable on purpose (SURVEY.md 8c hazard 5)
    "solvers16": ("solvers", {}, None, None),
    "default_12x10x3": ("default", {"n_basis_cells": "12 10 3"}, 0.75, 0.5),
    "cubic_8x6x5_periodic": ("cubic256", {"n_basis_cells": "8 6 5", "llg_temperature": "0"}, None, None),
}


def unit_random(n, seed):
    rng = np.random.default_rng(seed)
    z = rng.uniform(-1, 1, n)
    phi = rng.uniform(-np.pi, np.pi, n)
    r = np.sqrt(1 - z * z)
    return np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1)


def main():
    oracle = capi.load_oracle()
    tmp = tempfile.mkdtemp()
    for name, (preset, overrides, K, K4) in LLG_CASES.items():
        path = os.path.join(tmp, name + ".cfg")
        open(path, "w").write(cfgs.render(preset, **overrides))
        out = {}

        def fresh():
            o = S.Session(oracle, path)
            if K is not None:
                o.set_anisotropy(K, (0, 0, 1))
                o.set_cubic_anisotropy(K4)
            return o

        o = fresh()
        s0 = unit_random(o.nos, 1)
        out["spins0"] = s0
        g, e = o.gradient_and_energy(s0)
        out["gradient"], out["energy"] = g, e
        for term, (tot, per) in o.energy_contributions(s0, per_spin=True).items():
            out["E_" + term.replace(" ", "_")] = per
        o.close()
        for solver, n in (("Depondt", 5), ("Heun", 5), ("SIB", 5), ("RK4", 5), ("VP", 20)):
            o = fresh()
            o.llg_set(temperature=0.0, damping=0.3, dt=1e-3)
            o.set_spins(s0)
            o.llg_start(S.SOLVERS[solver], single_shot=True)
            o.n_shot(n)
            out["spins_" + solver] = o.spins().copy()
            out["energy_" + solver] = o.energy()
            out["torque_" + solver] = o.max_torque()
            o.stop()
            o.close()
        np.savez_compressed(os.path.join(HERE, "llg_%s.npz" % name), **out)
        print("wrote", name, {k: np.shape(v) for k, v in out.items() if k.startswith("spins_")})

    # GNEB: 7 images of 10x10x1 (solvers.cfg physics + K = 0.25), skyrmion -> +z, image 3 climbing, 60 VP single shots
    path = os.path.join(tmp, "gneb.cfg")
    open(path, "w").write(cfgs.render("solvers", n_basis_cells="10 10 1"))
    o = S.Session(oracle, path)
    o.set_anisotropy(0.25, (0, 0, 1))
    o.plus_z()
    o.skyrmion(3.0, phase=-90.0)
    o.chain_set_length(7)
    o.jump_to_image(6)
    o.plus_z()
    o.jump_to_image(0)
    o.transition_homogeneous(0, 6)
    out = {"images0": np.stack([o.spins(i).copy() for i in range(7)])}
    o.gneb_set_image_type(S.GNEB_CLIMBING, 3)
    o.gneb_start(S.SOLVER_VP, single_shot=True)
    o.n_shot(60)
    out["images"] = np.stack([o.spins(i).copy() for i in range(7)])
    rx, e = o.chain_rx_e()
    out["Rx"], out["E"] = rx, e
    out["max_torque"] = o.chain_max_torque()
    o.stop()
    o.close()
    np.savez_compressed(os.path.join(HERE, "gneb_7x10x10.npz"), **out)
    print("wrote gneb", out["Rx"], out["E"])


if __name__ == "__main__":
    main()
