"""The C-ABI boundary (CPU, no compute): the library loads, exports every symbol include/*.h declares, and fails
loudly -- never falls back to a CPU path -- when no CUDA device is present."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from spirit_b200 import capi, session as S


def test_every_declared_symbol_is_exported(product):
    api, ext = capi.declared_prototypes()
    assert len(api) > 150 and len(ext) >= 15
    assert product.missing == []
    out = subprocess.run(["nm", "-D", "--defined-only", capi.PRODUCT_LIB], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for name in list(api) + list(ext):
        assert name in exported, name
    # nothing but the C API is exported (C++ internals are hidden)
    assert not [s for s in exported if s.startswith("_Z")]


def test_north_star_entry_points_have_reference_signatures(product):
    api, _ = capi.declared_prototypes()
    c = ctypes
    P = c.POINTER
    assert api["State_Setup"] == (c.c_void_p, [c.c_char_p, c.c_bool])
    assert api["Simulation_LLG_Start"][1] == [c.c_void_p, c.c_int, c.c_int, c.c_int, c.c_bool, P(capi.Simulation_Run_Info), c.c_int, c.c_int]
    assert api["Simulation_GNEB_Start"][1] == [c.c_void_p, c.c_int, c.c_int, c.c_int, c.c_bool, P(capi.Simulation_Run_Info), c.c_int]
    assert api["Hamiltonian_Set_DMI"][1] == [c.c_void_p, c.c_int, P(c.c_float), c.c_int, c.c_int, c.c_int]
    assert api["Hamiltonian_Set_DDI"][1] == [c.c_void_p, c.c_int, P(c.c_int), c.c_float, c.c_bool, c.c_int, c.c_int]
    assert api["System_Get_Spin_Directions"] == (P(c.c_double), [c.c_void_p, c.c_int, c.c_int])
    assert api["System_Get_Energy"][0] is c.c_float  # the reference narrows energies to float at the ABI


def test_constants_match_reference_digits(product):
    """core/include/utility/Constants.hpp:18-46"""
    assert product.Constants_mu_B() == 0.057883817555
    assert product.Constants_k_B() == 0.08617330350
    assert product.Constants_gamma() == 0.1760859644
    assert product.Constants_mu_0() == 2.0133545e-28


def test_no_cpu_fallback(cfg, product):
    """Without a CUDA device every compute entry point reports failure; nothing is computed on the host"""
    if product.SpiritB200_Device_Count() > 0:
        pytest.skip("a CUDA device is present; the no-device behaviour is checked on the CPU box")
    p = S.Session(product, cfg("solvers"))
    n_err = product.Log_Get_N_Errors(p.state)
    with pytest.raises(RuntimeError):
        p.gradient_and_energy()
    with pytest.raises(RuntimeError):
        p.upload()
    before = p.spins().copy()
    p.llg_start(S.SOLVER_DEPONDT, n_iterations=3, n_iterations_log=3)
    assert np.array_equal(before, p.spins())  # the simulation did not run
    assert not product.Simulation_Running_On_Image(p.state, -1, -1)
    assert product.Log_Get_N_Errors(p.state) > n_err
    p.close()
