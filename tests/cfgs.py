"""input.cfg fixtures for the parity tests, written out on demand.

The parameter sets restate the reference's own test fixtures (same physics, none of their text):
  larmor      core/test/input/physics_larmor.cfg  (1 spin, B = 1 T, alpha = 0)
  fd_pairs    core/test/input/fd_pairs.cfg        (2x2x1 open, explicit pair table)
  solvers     core/test/input/solvers.cfg         (16x16x1, J=10, D=6 Bloch, B=25 T, periodic ab)
  ddi         core/test/input/physics_ddi.cfg     (5x5x5, 2-atom basis, skewed Bravais vectors, DDI FFT)
  default     input/input.cfg                     (100x100x1, BC 1 1 0, pair table, B=25 T)  -- BASELINE config 1
  cubic256    BASELINE config 2 (sc, neighbours J=10 D=6, K=1, mu_s=2, T>0), n_basis_cells overridable
`render(preset, key=value, ...)` overrides or adds single-line keys; `pairs=[...]` replaces the pair table.
Logging to file is always off so tests leave no files behind.
"""

COMMON = {
    "log_to_console": "0",
    "log_to_file": "0",
    "log_console_level": "1",
    "llg_output_any": "0",
    "llg_output_initial": "0",
    "llg_output_final": "0",
    "gneb_output_any": "0",
    "gneb_output_initial": "0",
    "gneb_output_final": "0",
    "llg_max_walltime": "0:0:0",
    "gneb_max_walltime": "0:0:0",
}

NEIGHBOURS = {
    "hamiltonian": "heisenberg_neighbours",
    "bravais_lattice": "sc",
    "mu_s": "2.0",
    "external_field_normal": "0.0 0.0 1.0",
    "anisotropy_magnitude": "0.0",
    "anisotropy_normal": "0.0 0.0 1.0",
    "n_shells_exchange": "1",
    "jij": "10.0",
    "dm_chirality": "1",
    "n_shells_dmi": "1",
    "dij": "6.0",
    "llg_seed": "20006",
    "llg_dt": "1e-3",
    "llg_temperature": "0",
    "llg_n_iterations": "2000000",
    "llg_n_iterations_log": "1000",
}

PRESETS = {
    "larmor": dict(NEIGHBOURS, n_basis_cells="1 1 1", boundary_conditions="1 1 0", external_field_magnitude="1",
                   llg_damping="0.0", llg_force_convergence="1e-8"),
    "solvers": dict(NEIGHBOURS, n_basis_cells="16 16 1", boundary_conditions="1 1 0", external_field_magnitude="25",
                    llg_damping="0.3", llg_force_convergence="1e-8", gneb_spring_constant="1.0",
                    gneb_force_convergence="1e-6", gneb_n_iterations="200000", gneb_n_iterations_log="1000"),
    "cubic256": dict(NEIGHBOURS, n_basis_cells="256 256 256", boundary_conditions="1 1 1", external_field_magnitude="0",
                     anisotropy_magnitude="1.0", llg_damping="0.3", llg_temperature="10", llg_force_convergence="1e-12",
                     llg_n_iterations_amortize="100"),
    "fd_pairs": {
        "hamiltonian": "heisenberg_pairs", "bravais_lattice": "sc", "n_basis_cells": "2 2 1",
        "boundary_conditions": "0 0 0", "external_field_magnitude": "25.0", "external_field_normal": "0.0 0.0 1.0",
        "mu_s": "2.0", "anisotropy_magnitude": "0.0", "anisotropy_normal": "0.0 0.0 1.0",
        "_pairs": ["i j   da db dc   Dijx Dijy Dijz   Jij",
                   "0 0   1  0  0    6.0  0.0  0.0    10.0",
                   "0 0   0  1  0    0.0  6.0  0.0    10.0",
                   "0 0   0  0  1    0.0  0.0  6.0    10.0"],
    },
    "default": {
        "hamiltonian": "heisenberg_pairs", "bravais_lattice": "sc", "n_basis_cells": "100 100 1",
        "boundary_conditions": "1 1 0", "external_field_magnitude": "25.0", "external_field_normal": "0.0 0.0 1.0",
        "mu_s": "2.0", "anisotropy_magnitude": "0.0", "anisotropy_normal": "0.0 0.0 1.0",
        "llg_damping": "0.3", "llg_dt": "1.0E-3", "llg_temperature": "0", "llg_seed": "20006",
        "llg_force_convergence": "10e-9", "llg_n_iterations": "2000000", "llg_n_iterations_log": "2000",
        "llg_stt_use_gradient": "0", "llg_stt_magnitude": "0.0",
        "_pairs": ["i j   da db dc    Jij   Dij  Dijx Dijy Dijz",
                   "0 0    1  0  0   10.0   6.0   1.0  0.0  0.0",
                   "0 0    0  1  0   10.0   6.0   0.0  1.0  0.0",
                   "0 0    0  0  1   10.0   6.0   0.0  0.0  1.0"],
    },
    "ddi": {
        "hamiltonian": "heisenberg_pairs", "boundary_conditions": "1 0 0", "external_field_magnitude": "0.0",
        "external_field_normal": "0.0 0.0 1.0", "mu_s": "2.17", "anisotropy_magnitude": "0.0",
        "anisotropy_normal": "0.0 0.0 1.0", "ddi_method": "fft", "ddi_n_periodic_images": "4 4 4",
        "ddi_radius": "10.0", "ddi_pb_zero_padding": "1", "lattice_constant": "2.77", "n_basis_cells": "5 5 5",
        "llg_damping": "0.3", "llg_dt": "1.0E-3", "llg_seed": "20006", "llg_force_convergence": "10e-9",
        "_pairs": [],
        "_block": ["bravais_vectors", " 3   0 -1", " 0   1  0", "0.3  -2  1", "basis", "2", "0 0 0", "0.5 0.2 0.7"],
    },
}


def render(preset, pairs=None, block=None, **overrides):
    p = dict(COMMON)
    p.update(PRESETS[preset])
    p.update({k: str(v) for k, v in overrides.items()})
    if pairs is not None:
        p["_pairs"] = pairs
    if block is not None:
        p["_block"] = block
    lines = []
    for k, v in p.items():
        if not k.startswith("_"):
            lines.append("%s %s" % (k, v))
    if "_pairs" in p:
        rows = p["_pairs"]
        lines.append("n_interaction_pairs %d" % max(0, len(rows) - 1))
        lines.extend(rows)
    lines.extend(p.get("_block", []))
    return "\n".join(lines) + "\n"
