"""A small host-side driver over the Spirit C API (either library of capi.py).

It plays the role of the reference's ctypes package (core/python/spirit/{state,system,simulation,hamiltonian,
configuration,chain,transition,parameters}.py): same C entry points, same argument meaning, one object per `State*`.
Arrays returned by `spins` / `effective_field` are live numpy views into library storage, as in
core/python/spirit/system.py:49-63.
"""
import ctypes
import os

import numpy as np

from . import capi

# core/include/Spirit/Simulation.h:33-54
SOLVER_VP, SOLVER_SIB, SOLVER_DEPONDT, SOLVER_HEUN, SOLVER_RK4 = 0, 1, 2, 3, 4
SOLVER_LBFGS_OSO, SOLVER_LBFGS_ATLAS, SOLVER_VP_OSO = 5, 6, 7  # core/include/Spirit/Simulation.h:33-54
SOLVERS = {"VP": 0, "SIB": 1, "Depondt": 2, "Heun": 3, "RK4": 4, "LBFGS_OSO": 5, "LBFGS_Atlas": 6, "VP_OSO": 7}
# core/include/Spirit/Hamiltonian.h:31-57
CHIRALITY_BLOCH, CHIRALITY_NEEL = 1, 2
DDI_NONE, DDI_FFT, DDI_FMM, DDI_CUTOFF = 0, 1, 2, 3
# core/include/Spirit/Parameters_GNEB.h: image types
GNEB_NORMAL, GNEB_CLIMBING, GNEB_FALLING, GNEB_STATIONARY = 0, 1, 2, 3


def _f3(v):
    return (ctypes.c_float * 3)(*[float(x) for x in v])


class Session:
    def __init__(self, lib, cfg="", quiet=True):
        self.lib = lib
        self.state = lib.State_Setup(os.fsencode(cfg), bool(quiet))
        if not self.state:
            raise RuntimeError("State_Setup failed for %r" % cfg)

    def close(self):
        if self.state:
            self.lib.State_Delete(self.state)
            self.state = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- system -----------------------------------------------------------------------------------------------
    @property
    def nos(self):
        return self.lib.System_Get_NOS(self.state, -1, -1)

    def _view(self, ptr, n):
        return np.ctypeslib.as_array(ptr, shape=(n, 3))

    def spins(self, idx_image=-1):
        """live [nos][3] view (System_Get_Spin_Directions, System.h:31)"""
        return self._view(self.lib.System_Get_Spin_Directions(self.state, idx_image, -1), self.nos)

    def set_spins(self, array, idx_image=-1):
        self.spins(idx_image)[:] = np.asarray(array, dtype=np.float64).reshape(self.nos, 3)

    def effective_field(self, idx_image=-1):
        return self._view(self.lib.System_Get_Effective_Field(self.state, idx_image, -1), self.nos)

    def update_data(self, idx_image=-1):
        self.lib.System_Update_Data(self.state, idx_image, -1)

    def energy_float(self, idx_image=-1):
        return self.lib.System_Get_Energy(self.state, idx_image, -1)

    # ---- double-precision probes (include/spirit_b200.h <-> oracle/ref_shim.cpp) ---------------------------------
    def gradient_and_energy(self, spins=None, idx_image=-1):
        n = self.nos
        g = np.zeros((n, 3))
        e = ctypes.c_double(0)
        sp = None if spins is None else np.ascontiguousarray(spins, dtype=np.float64).ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        r = self.lib.SpiritB200_Gradient_and_Energy(self.state, sp, g.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.byref(e), idx_image)
        if r < 0:
            raise RuntimeError("Gradient_and_Energy failed")
        return g, e.value

    def gradient(self, spins=None, idx_image=-1):
        n = self.nos
        g = np.zeros((n, 3))
        sp = None if spins is None else np.ascontiguousarray(spins, dtype=np.float64).ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        r = self.lib.SpiritB200_Gradient(self.state, sp, g.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), idx_image)
        if r < 0:
            raise RuntimeError("Gradient failed")
        return g

    def energy_contributions(self, spins=None, per_spin=False, idx_image=-1):
        n = self.nos
        names = ctypes.create_string_buffer(32 * 8)
        totals = np.zeros(8)
        ps = np.zeros((8, n)) if per_spin else None
        sp = None if spins is None else np.ascontiguousarray(spins, dtype=np.float64).ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        nt = self.lib.SpiritB200_Energy_Contributions(
            self.state, sp, 8, names, totals.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
            ps.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if per_spin else None, idx_image)
        if nt < 0:
            raise RuntimeError("Energy_Contributions failed")
        out = {}
        for t in range(nt):
            name = names.raw[32 * t:32 * t + 32].split(b"\0")[0].decode()
            out[name] = (totals[t], ps[t].copy() if per_spin else None)
        return out

    def energy(self, idx_image=-1):
        return self.lib.SpiritB200_Get_Energy(self.state, idx_image)

    def max_torque(self, idx_image=-1):
        if idx_image == -1:
            idx_image = self.lib.System_Get_Index(self.state)
        return self.lib.SpiritB200_Get_MaxTorque(self.state, idx_image)

    def chain_max_torque(self):
        return self.lib.SpiritB200_Get_MaxTorque(self.state, -2)

    def chain_rx_e(self):
        noi = self.noi
        rx, e = np.zeros(noi), np.zeros(noi)
        P = ctypes.POINTER(ctypes.c_double)
        self.lib.SpiritB200_Chain_Get_Rx_E(self.state, rx.ctypes.data_as(P), e.ctypes.data_as(P))
        return rx, e

    def magnetization(self, idx_image=-1):
        m = np.zeros(3)
        self.lib.SpiritB200_Get_Magnetization(self.state, m.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), idx_image)
        return m

    def pairs(self, kind, idx_image=-1):
        """redundant pair lists: kind 0 exchange, 1 DMI -> (ijt [n][5], magnitudes [n], normals [n][3])"""
        cap = 4096
        ijt = np.zeros((cap, 5), dtype=np.int32)
        mag, nrm = np.zeros(cap), np.zeros((cap, 3))
        n = self.lib.SpiritB200_Get_Pairs(
            self.state, kind, cap, ijt.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
            mag.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), nrm.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), idx_image)
        return ijt[:n], mag[:n], nrm[:n]

    # ---- configurations (Configurations.h:42-150) -------------------------------------------------------------------
    _pos0, _rect0 = (0, 0, 0), (-1, -1, -1)

    def plus_z(self, idx_image=-1):
        self.lib.Configuration_PlusZ(self.state, _f3(self._pos0), _f3(self._rect0), -1, -1, False, idx_image, -1)

    def minus_z(self, idx_image=-1):
        self.lib.Configuration_MinusZ(self.state, _f3(self._pos0), _f3(self._rect0), -1, -1, False, idx_image, -1)

    def domain(self, direction, idx_image=-1):
        self.lib.Configuration_Domain(self.state, _f3(direction), _f3(self._pos0), _f3(self._rect0), -1, -1, False, idx_image, -1)

    def random(self, idx_image=-1):
        self.lib.Configuration_Random(self.state, _f3(self._pos0), _f3(self._rect0), -1, -1, False, False, idx_image, -1)

    def skyrmion(self, radius, order=1, phase=0, up_down=False, achiral=False, rl=False, pos=(0, 0, 0), idx_image=-1):
        self.lib.Configuration_Skyrmion(
            self.state, radius, order, phase, up_down, achiral, rl, _f3(pos), _f3(self._rect0), -1, -1, False, idx_image, -1)

    # ---- pinning and defects (Configurations.h:44-48, Geometry.h:42) ----------------------------------------------------
    def set_pinned(self, pinned, pos=(0, 0, 0), rect=(-1, -1, -1), cylindrical=-1, spherical=-1, inverted=False, idx_image=-1):
        self.lib.Configuration_Set_Pinned(self.state, bool(pinned), _f3(pos), _f3(rect), cylindrical, spherical, inverted, idx_image, -1)

    def set_atom_type(self, atom_type, pos=(0, 0, 0), rect=(-1, -1, -1), cylindrical=-1, spherical=-1, inverted=False, idx_image=-1):
        self.lib.Configuration_Set_Atom_Type(self.state, int(atom_type), _f3(pos), _f3(rect), cylindrical, spherical, inverted, idx_image, -1)

    def atom_types(self, idx_image=-1):
        ptr = self.lib.Geometry_Get_Atom_Types(self.state, idx_image, -1)
        return np.ctypeslib.as_array(ptr, shape=(self.nos,)).copy()

    # ---- hamiltonian (Hamiltonian.h:65-101) --------------------------------------------------------------------------
    def set_boundary_conditions(self, bc, idx_image=-1):
        self.lib.Hamiltonian_Set_Boundary_Conditions(self.state, (ctypes.c_bool * 3)(*[bool(b) for b in bc]), idx_image, -1)

    def set_field(self, magnitude, normal, idx_image=-1):
        self.lib.Hamiltonian_Set_Field(self.state, magnitude, _f3(normal), idx_image, -1)

    def set_anisotropy(self, magnitude, normal, idx_image=-1):
        self.lib.Hamiltonian_Set_Anisotropy(self.state, magnitude, _f3(normal), idx_image, -1)

    def set_cubic_anisotropy(self, magnitude, idx_image=-1):
        self.lib.Hamiltonian_Set_Cubic_Anisotropy(self.state, magnitude, idx_image, -1)

    def set_exchange(self, jij, idx_image=-1):
        self.lib.Hamiltonian_Set_Exchange(self.state, len(jij), (ctypes.c_float * max(1, len(jij)))(*jij), idx_image, -1)

    def set_dmi(self, dij, chirality=CHIRALITY_BLOCH, idx_image=-1):
        self.lib.Hamiltonian_Set_DMI(self.state, len(dij), (ctypes.c_float * max(1, len(dij)))(*dij), chirality, idx_image, -1)

    def set_ddi(self, method, n_periodic_images=(4, 4, 4), cutoff=0.0, zero_padding=True, idx_image=-1):
        self.lib.Hamiltonian_Set_DDI(self.state, method, (ctypes.c_int * 3)(*n_periodic_images), cutoff, zero_padding, idx_image, -1)

    # ---- LLG parameters (Parameters_LLG.h) ---------------------------------------------------------------------------
    def llg_set(self, dt=None, damping=None, temperature=None, convergence=None, direct_minimization=None,
                n_iterations=None, idx_image=-1):
        L, s = self.lib, self.state
        if dt is not None:
            L.Parameters_LLG_Set_Time_Step(s, dt, idx_image, -1)
        if damping is not None:
            L.Parameters_LLG_Set_Damping(s, damping, idx_image, -1)
        if temperature is not None:
            L.Parameters_LLG_Set_Temperature(s, temperature, idx_image, -1)
        if convergence is not None:
            L.Parameters_LLG_Set_Convergence(s, convergence, idx_image, -1)
        if direct_minimization is not None:
            L.Parameters_LLG_Set_Direct_Minimization(s, direct_minimization, idx_image, -1)
        if n_iterations is not None:
            L.Parameters_LLG_Set_N_Iterations(s, n_iterations[0], n_iterations[1], idx_image, -1)

    def llg_no_output(self, idx_image=-1):
        self.lib.Parameters_LLG_Set_Output_General(self.state, False, False, False, idx_image, -1)

    def gneb_no_output(self):
        self.lib.Parameters_GNEB_Set_Output_General(self.state, False, False, False, -1)

    # ---- simulation (Simulation.h:81-218) ----------------------------------------------------------------------------
    def llg_start(self, solver, n_iterations=-1, n_iterations_log=-1, single_shot=False, idx_image=-1):
        info = capi.Simulation_Run_Info()
        self.lib.Simulation_LLG_Start(self.state, solver, n_iterations, n_iterations_log, single_shot, ctypes.byref(info), idx_image, -1)
        return info

    def gneb_start(self, solver, n_iterations=-1, n_iterations_log=-1, single_shot=False):
        info = capi.Simulation_Run_Info()
        self.lib.Simulation_GNEB_Start(self.state, solver, n_iterations, n_iterations_log, single_shot, ctypes.byref(info), -1)
        return info

    def single_shot(self, idx_image=-1):
        self.lib.Simulation_SingleShot(self.state, idx_image, -1)

    def n_shot(self, n, idx_image=-1):
        self.lib.Simulation_N_Shot(self.state, n, idx_image, -1)

    def stop(self, idx_image=-1):
        self.lib.Simulation_Stop(self.state, idx_image, -1)

    def stop_all(self):
        self.lib.Simulation_Stop_All(self.state)

    # ---- chain (Chain.h, Transitions.h) ----------------------------------------------------------------------------
    @property
    def noi(self):
        return self.lib.Chain_Get_NOI(self.state, -1)

    def chain_set_length(self, n):
        """Chain_Image_to_Clipboard + Chain_Set_Length (Chain.h:60,49)"""
        self.lib.Chain_Image_to_Clipboard(self.state, -1, -1)
        self.lib.Chain_Set_Length(self.state, n, -1)

    def jump_to_image(self, idx):
        self.lib.Chain_Jump_To_Image(self.state, idx, -1)

    def transition_homogeneous(self, first, last):
        self.lib.Transition_Homogeneous(self.state, first, last, -1)

    def chain_update_data(self):
        self.lib.Chain_Update_Data(self.state, -1)

    def gneb_set_image_type(self, image_type, idx_image=-1):
        self.lib.Parameters_GNEB_Set_Climbing_Falling(self.state, image_type, idx_image, -1)

    def gneb_set_image_type_automatically(self):
        self.lib.Parameters_GNEB_Set_Image_Type_Automatically(self.state, -1)

    # ---- device (product only) ---------------------------------------------------------------------------------------
    # ---- OVF files (Spirit/IO.h) ----------------------------------------------------------------------------------------
    def n_images_in_file(self, path):
        return self.lib.IO_N_Images_In_File(self.state, str(path).encode(), -1, -1)

    def image_write(self, path, fmt=2, comment="-", idx_image=-1):
        self.lib.IO_Image_Write(self.state, str(path).encode(), fmt, comment.encode(), idx_image, -1)

    def image_append(self, path, fmt=2, comment="-", idx_image=-1):
        self.lib.IO_Image_Append(self.state, str(path).encode(), fmt, comment.encode(), idx_image, -1)

    def image_read(self, path, idx_image_infile=0, idx_image=-1):
        self.lib.IO_Image_Read(self.state, str(path).encode(), idx_image_infile, idx_image, -1)

    def chain_write(self, path, fmt=3, comment="-"):
        self.lib.IO_Chain_Write(self.state, str(path).encode(), fmt, comment.encode(), -1)

    def chain_append(self, path, fmt=3, comment="-"):
        self.lib.IO_Chain_Append(self.state, str(path).encode(), fmt, comment.encode(), -1)

    def chain_read(self, path, start=0, end=-1, insert_idx=0):
        self.lib.IO_Chain_Read(self.state, str(path).encode(), start, end, insert_idx, -1)

    def upload(self, idx_image=-1):
        if self.lib.SpiritB200_Upload(self.state, idx_image) < 0:
            raise RuntimeError("SpiritB200_Upload failed (no CUDA device?)")

    def download(self, idx_image=-1):
        if self.lib.SpiritB200_Download(self.state, idx_image) < 0:
            raise RuntimeError("SpiritB200_Download failed")

    def iterate_device(self, solver, n_iterations, idx_image=-1):
        """n iterations on HBM-resident spins; returns milliseconds from CUDA events on the image's stream"""
        ms = self.lib.SpiritB200_LLG_Iterate_Device(self.state, solver, n_iterations, idx_image)
        if ms < 0:
            raise RuntimeError("SpiritB200_LLG_Iterate_Device failed")
        return ms

    def stencil_variant(self, idx_image=-1):
        """1: nearest-neighbour marching kernels, 0: generic gather kernels (include/spirit_b200.h)"""
        return self.lib.SpiritB200_Stencil_Variant(self.state, idx_image)

    def step_variant(self, solver, idx_image=-1):
        """2: one fused predictor + corrector kernel per iteration, 1: one marching kernel per stage, 0: generic gather"""
        return self.lib.SpiritB200_Step_Variant(self.state, solver, idx_image)

    def kernel_launches(self, idx_image=-1):
        return self.lib.SpiritB200_Kernel_Launches(self.state, idx_image)
