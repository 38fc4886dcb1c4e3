"""oracle/restatement.py -- TEST INFRASTRUCTURE, not product code.

A NumPy restatement of the reference's hot-path arithmetic for the nearest-neighbour simple-cubic Heisenberg
Hamiltonian (exchange + bond-parallel DMI + uniaxial / cubic anisotropy + Zeeman), the LLG virtual force and the
Depondt / Heun / SIB / RK4 / VP solver updates, and the GNEB force. It is an independent third opinion beside the
compiled reference (oracle/_ref/libSpirit_ref.so) and documents the mathematics the CUDA kernels implement.

Pinned: tests/test_oracle.py checks every function here against the compiled reference and against the committed
golden vectors (tests/golden/*.npz, produced by tests/golden/make_golden.py from the reference itself).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

Reference locations (relative to /root/reference/core):
  gradient terms            src/engine/Hamiltonian_Heisenberg.cpp:768-864
  energy                    src/engine/Hamiltonian_Heisenberg.cpp:704-766 (1/2 g.s for bilinear terms)
  idx_from_pair / BC        include/engine/Vectormath.hpp:437-528
  virtual force             src/engine/Method_LLG.cpp:131-226
  Depondt / Heun / SIB      include/engine/Solver_Depondt.hpp:29-77, Solver_Heun.hpp:30-81, Solver_SIB.hpp:22-50,
                            src/engine/Solver_Kernels.cpp:15-48 (sib_transform)
  RK4 / VP                  include/engine/Solver_RK4.hpp:41-147, Solver_VP.hpp:29-114
  rotate                    src/engine/Vectormath.cpp:474-485
  GNEB force, tangents      src/engine/Method_GNEB.cpp:87-258, src/engine/Manifoldmath.cpp:115-213
  constants                 include/utility/Constants.hpp:18-46
  topological charge        src/engine/Vectormath.cpp:462-472,504-631
  dipolar direct sum        src/engine/Hamiltonian_Heisenberg.cpp:1016-1071
  pinning / defects         include/engine/Vectormath.hpp:406-528 (check_atom_type in idx_from_pair and in every on-site term),
                            src/data/Geometry.cpp:72-82 (mu_s = 0 at defect sites), src/engine/Method_LLG.cpp:122-124, 222-224
                            (mask_unpinned on gradient and virtual force) -- pinned against the reference built with
                            -DSPIRIT_ENABLE_PINNING -DSPIRIT_ENABLE_DEFECTS (oracle/_ref/libSpirit_ref_pd.so)
"""
import numpy as np

mu_B = 0.057883817555
gamma = 0.1760859644
k_B = 0.08617330350


class Model:
    """sc lattice n = (Na, Nb, Nc), periodic flags bc, first-shell J, Bloch DMI D (d parallel to the bond),
    uniaxial K along Kn, cubic K4, field B (Tesla) along Bn, mu_s (mu_B)."""

    def __init__(self, n, bc, J=10.0, D=6.0, B=25.0, Bn=(0, 0, 1), mu_s=2.0, K=0.0, Kn=(0, 0, 1), K4=0.0,
                 dt=1e-3, alpha=0.3, atom_types=None, defect_sites=None, pinned=None):
        self.n, self.bc = tuple(n), tuple(bc)
        self.J, self.D, self.B, self.Bn = J, D, B, np.asarray(Bn, float)
        self.mu_s, self.K, self.Kn, self.K4 = mu_s, K, np.asarray(Kn, float), K4
        self.dt, self.alpha = dt, alpha
        # per-site masks in the reference's site order (a fastest): present = atom type >= 0 (check_atom_type), moment =
        # mu_s != 0 (zero at EVERY defect site, whatever its type), free = mask_unpinned
        nos = int(np.prod(self.n))
        types = np.zeros(nos, int) if atom_types is None else np.asarray(atom_types, int)
        self.present = types >= 0
        self.moment = np.ones(nos, bool) if defect_sites is None else ~np.asarray(defect_sites, bool)
        self.free = np.ones(nos, bool) if pinned is None else ~np.asarray(pinned, bool)

    # ---- Hamiltonian ----------------------------------------------------------------------------------------------
    def pair_gradient(self, S):
        """exchange + DMI, S shaped (Nc, Nb, Na, 3): g_i -= J s_j + D s_j x d_ij over the (redundant) neighbours"""
        g = np.zeros_like(S)
        present = self.present.reshape(S.shape[:3])
        for axis, (N, per) in zip((2, 1, 0), zip(self.n, self.bc)):
            d = np.zeros(3)
            d[2 - axis] = 1.0
            for sign in (+1, -1):
                if N == 1 and not per:
                    continue
                Sj = np.roll(S, -sign, axis=axis)
                mask = present & np.roll(present, -sign, axis=axis)  # idx_from_pair: both atoms of the pair must be there
                if not per:
                    idx = [slice(None)] * 3
                    idx[axis] = (N - 1) if sign > 0 else 0
                    mask[tuple(idx)] = False
                g -= (self.J * Sj + self.D * np.cross(Sj, sign * d)) * mask[..., None]
        return g

    def gradient_and_energy(self, s):
        Na, Nb, Nc = self.n
        S = s.reshape(Nc, Nb, Na, 3)
        here = self.present.reshape(Nc, Nb, Na, 1)
        gp = self.pair_gradient(S)
        ga = -2 * self.K * (S @ self.Kn)[..., None] * self.Kn * here
        gc = -2 * self.K4 * S ** 3 * here
        gz = -self.mu_s * mu_B * self.B * self.Bn * np.ones_like(S) * (here & self.moment.reshape(Nc, Nb, Na, 1))
        E = 0.5 * np.sum((gp + ga) * S) - 0.5 * self.K4 * np.sum(S ** 4 * here) + np.sum(gz * S)
        return (gp + ga + gc + gz).reshape(-1, 3), E

    def gradient(self, s):
        return self.gradient_and_energy(s)[0]

    # ---- LLG ------------------------------------------------------------------------------------------------------------
    def virtual_force(self, s, xi=None):
        dtg = self.dt * gamma / mu_B / (1 + self.alpha ** 2)
        F = -self.gradient(s) * self.free[:, None]  # Method_LLG.cpp:122-124
        fv = (dtg * F + dtg * self.alpha * np.cross(s, F)) / self.mu_s
        if xi is not None:
            fv = fv + xi + self.alpha * np.cross(s, xi)
        return fv * self.free[:, None]  # Method_LLG.cpp:222-224

    def thermal_amplitude(self, T):
        """epsilon * sqrt(T / mu_s), Method_LLG.cpp:74-75,105"""
        return np.sqrt(2 * self.alpha * self.dt * gamma / mu_B * k_B) / (1 + self.alpha ** 2) * np.sqrt(T / self.mu_s)

    def depondt(self, s, xi=None):
        H1 = self.virtual_force(s, xi)
        sp = rotate(s, H1)
        return rotate(s, 0.5 * (H1 + self.virtual_force(sp, xi)))

    def heun(self, s):
        k1 = -np.cross(s, self.virtual_force(s))
        sp = normalize(s + k1)
        return normalize(s + 0.5 * k1 - 0.5 * np.cross(sp, self.virtual_force(sp)))

    def sib(self, s):
        sp = 0.5 * (s + sib_transform(s, self.virtual_force(s)))
        return sib_transform(s, self.virtual_force(sp))

    def rk4(self, s):
        fv = self.virtual_force
        k1 = -np.cross(s, fv(s))
        s1 = normalize(s + 0.5 * k1)
        k2 = -np.cross(s1, fv(s1))
        s2 = normalize(s + 0.5 * k2)
        k3 = -np.cross(s2, fv(s2))
        s3 = normalize(s + k3)
        k4 = -np.cross(s3, fv(s3))
        return normalize(s + k1 / 6 + k2 / 3 + k3 / 3 + k4 / 6)

    def vp_single_shots(self, s, n_steps):
        """VP with the post-iteration hook after every iteration (Simulation_SingleShot): the hook projects the force
        in place, and the projected force is next iteration's F_prev (SURVEY.md 8c hazard 6)"""
        F = project_tangential(-self.gradient(s), s)
        v = np.zeros_like(s)
        for _ in range(n_steps):
            Fp, F = F, -self.gradient(s)
            v = v + 0.5 * (Fp + F)
            p_, f2 = np.sum(v * F), np.sum(F * F)
            v = F * (p_ / f2) if p_ > 0 else np.zeros_like(s)
            s = normalize(s + self.dt * v + 0.5 * self.dt * F)
            F = project_tangential(F, s)
        return s


def normalize(x):
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def project_tangential(F, s):
    return F - np.sum(F * s, axis=1, keepdims=True) * s


def rotate(v, H):
    """Rodrigues rotation of v about H/|H| by |H| (zero force: identity)"""
    th = np.linalg.norm(H, axis=1, keepdims=True)
    k = np.divide(H, th, out=np.zeros_like(H), where=th > 0)
    return v * np.cos(th) + np.cross(k, v) * np.sin(th) + k * np.sum(k * v, axis=1, keepdims=True) * (1 - np.cos(th))


def sib_transform(s, H):
    A = 0.5 * H
    a = s - np.cross(s, A)
    Ax, Ay, Az = A.T
    x, y, z = a.T
    o = np.stack([x * (Ax * Ax + 1) + y * (Ax * Ay - Az) + z * (Ax * Az + Ay),
                  x * (Ay * Ax + Az) + y * (Ay * Ay + 1) + z * (Ay * Az - Ax),
                  x * (Az * Ax - Ay) + y * (Az * Ay + Ax) + z * (Az * Az + 1)], axis=1)
    return o / (1 + np.sum(A * A, axis=1))[:, None]


# ---- GNEB -------------------------------------------------------------------------------------------------------------------
NORMAL, CLIMBING, FALLING, STATIONARY = 0, 1, 2, 3


def geodesic_distance(a, b):
    return np.sqrt(np.sum(np.arccos(np.clip(np.sum(a * b, axis=1), -1, 1)) ** 2))


def tangents(imgs, E):
    """energy-weighted tangents of the interior images, projected and normalised in 3N space (Manifoldmath.cpp:143-211)"""
    T = [None] * len(imgs)
    for i in range(1, len(imgs) - 1):
        tp, tm = imgs[i + 1] - imgs[i], imgs[i] - imgs[i - 1]
        Em, Ep, Emi = E[i], E[i + 1], E[i - 1]
        if (Ep < Em and Em > Emi) or (Ep > Em and Em < Emi):
            Emax, Emin = max(abs(Ep - Em), abs(Emi - Em)), min(abs(Ep - Em), abs(Emi - Em))
            t = Emax * tp + Emin * tm if Ep > Emi else Emin * tp + Emax * tm
        elif Ep > Em > Emi:
            t = tp
        elif Ep < Em < Emi:
            t = tm
        else:
            t = tp + tm
        t = project_tangential(t, imgs[i])
        T[i] = t / np.sqrt(np.sum(t * t))
    return T


def gneb_force(model, imgs, types, k_spring):
    gE = [model.gradient_and_energy(s) for s in imgs]
    E = [e for _, e in gE]
    Rx = [0.0]
    for i in range(1, len(imgs)):
        Rx.append(Rx[-1] + geodesic_distance(imgs[i], imgs[i - 1]))
    T = tangents(imgs, E)
    F = [np.zeros_like(imgs[0]) for _ in imgs]
    for i in range(1, len(imgs) - 1):
        Fg = project_tangential(-gE[i][0], imgs[i])
        if types[i] == CLIMBING:
            F[i] = Fg - 2 * np.sum(Fg * T[i]) * T[i]
        elif types[i] == FALLING:
            F[i] = Fg
        elif types[i] == NORMAL:
            F[i] = Fg - np.sum(Fg * T[i]) * T[i] + k_spring * (Rx[i + 1] - 2 * Rx[i] + Rx[i - 1]) * T[i]
    return F, E, Rx


def gneb_vp_single_shots(model, imgs, types, k_spring, n_steps):
    """GNEB + VP coupled over all images, hook (in-place projection of the forces) after every iteration"""
    noi = len(imgs)
    imgs = [x.copy() for x in imgs]
    F = [np.zeros_like(imgs[0]) for _ in imgs]
    v = [np.zeros_like(imgs[0]) for _ in imgs]
    E = Rx = None
    for _ in range(n_steps):
        Fp = F
        F, E, Rx = gneb_force(model, imgs, types, k_spring)
        v = [v[i] + 0.5 * (Fp[i] + F[i]) for i in range(noi)]
        pf = sum(np.sum(v[i] * F[i]) for i in range(noi))
        f2 = sum(np.sum(F[i] * F[i]) for i in range(noi))
        for i in range(noi):
            v[i] = F[i] * (pf / f2) if pf > 0 else np.zeros_like(F[i])
            imgs[i] = normalize(imgs[i] + model.dt * v[i] + 0.5 * model.dt * F[i])
        F = [project_tangential(F[i], imgs[i]) for i in range(noi)]
    return imgs, E, Rx


def gneb_two_stage_single_shots(model, imgs, types, k_spring, n_steps, solver):
    """GNEB with Depondt / Heun / SIB over all images (Solver_*.hpp with noi images). Virtual force dt gamma/mu_B s x F,
    ZERO for the two end images (Method_GNEB.cpp:359-391 skips them).
    Note: the compiled reference leaves forces_virtual of the end images UNINITIALISED for SIB
    (Solver_SIB.hpp:4-5 allocates without a fill value), so its GNEB + SIB results are not reproducible; this function
    states the intended semantics and is the oracle for that combination."""
    imgs = [x.copy() for x in imgs]
    dtg = model.dt * gamma / mu_B

    def fv(conf):
        F, E, Rx = gneb_force(model, conf, types, k_spring)
        out = [dtg * np.cross(s, f) for s, f in zip(conf, F)]
        out[0] = np.zeros_like(out[0])
        out[-1] = np.zeros_like(out[-1])
        return out, E, Rx

    E = Rx = None
    for _ in range(n_steps):
        Fv, _, _ = fv(imgs)
        if solver == "SIB":
            pred = [0.5 * (s + sib_transform(s, f)) for s, f in zip(imgs, Fv)]
            Fvp, E, Rx = fv(pred)
            imgs = [sib_transform(s, f) for s, f in zip(imgs, Fvp)]
        elif solver == "Depondt":
            pred = [rotate(s, f) for s, f in zip(imgs, Fv)]
            Fvp, E, Rx = fv(pred)
            imgs = [rotate(s, 0.5 * (f + fp)) for s, f, fp in zip(imgs, Fv, Fvp)]
        elif solver == "Heun":
            k1 = [-np.cross(s, f) for s, f in zip(imgs, Fv)]
            pred = [normalize(s + k) for s, k in zip(imgs, k1)]
            Fvp, E, Rx = fv(pred)
            imgs = [normalize(s + 0.5 * k - 0.5 * np.cross(sp, fp)) for s, k, sp, fp in zip(imgs, k1, pred, Fvp)]
        elif solver == "RK4":
            # Solver_RK4.hpp:41-147 over noi images (Method_GNEB.cpp:749 instantiates it; Simulation_GNEB_Start does not
            # dispatch to it, so the compiled reference cannot be the oracle for this combination)
            k1 = [-np.cross(s, f) for s, f in zip(imgs, Fv)]
            pred = [normalize(s + 0.5 * k) for s, k in zip(imgs, k1)]
            Fvp, _, _ = fv(pred)
            k2 = [-np.cross(sp, fp) for sp, fp in zip(pred, Fvp)]
            pred = [normalize(s + 0.5 * k) for s, k in zip(imgs, k2)]
            Fvp, _, _ = fv(pred)
            k3 = [-np.cross(sp, fp) for sp, fp in zip(pred, Fvp)]
            pred = [normalize(s + k) for s, k in zip(imgs, k3)]
            Fvp, E, Rx = fv(pred)
            k4 = [-np.cross(sp, fp) for sp, fp in zip(pred, Fvp)]
            imgs = [normalize(s + a / 6 + b / 3 + c / 3 + d / 6) for s, a, b, c, d in zip(imgs, k1, k2, k3, k4)]
        else:
            raise ValueError(solver)
    return imgs, E, Rx


# ---------------------------------------------------------------------------------------------------------------------------
# Topological charge of a planar lattice with one basis atom (src/engine/Vectormath.cpp:462-472,504-631): the cell
# parallelogram (0, a, b, a+b) is cut along the a-b diagonal (the Delaunay choice for square and hexagonal cells, the
# reference's library picks it for the degenerate square cell too); a triangle counts when its translations are inside the
# lattice or allowed by the boundary conditions; charge = sign / (4 pi) * 2 atan2( s1.(s2 x s3), 1 + s1.s2 + s1.s3 + s2.s3 ).
# ---------------------------------------------------------------------------------------------------------------------------
def solid_angle(s1, s2, s3):
    x = np.einsum("...i,...i", s1, np.cross(s2, s3))
    y = 1 + np.einsum("...i,...i", s1, s2) + np.einsum("...i,...i", s1, s3) + np.einsum("...i,...i", s2, s3)
    return 2 * np.arctan2(x, y)


def topological_charge(spins, n_cells, bc, ta=(1.0, 0.0), tb=(0.0, 1.0)):
    """spins [Nb*Na][3] (a fastest), bc = (periodic a, periodic b); ta, tb: the in-plane bravais vectors.
    Returns (total charge, {sorted site triple: charge})."""
    Na, Nb = n_cells
    s = np.asarray(spins, dtype=float).reshape(Nb, Na, 3)
    ta, tb = np.asarray(ta, dtype=float), np.asarray(tb, dtype=float)

    def orientation(p0, p1, p2):  # z of (p0 - p1) x (p0 - p2)
        u, v = p0 - p1, p0 - p2
        return 1.0 if u[0] * v[1] - u[1] * v[0] > 0 else -1.0

    k = 0.1  # the reference stretches the corners of the cell away from its centre before triangulating
    P0, Pab, Pb, Pa = -k * (ta + tb), (1 + k) * (ta + tb), tb - k * (ta - tb), ta + k * (ta - tb)
    sign = (orientation(Pa, Pb, Pab), orientation(Pa, Pb, P0))
    total, per_triangle = 0.0, {}
    for b in range(Nb):
        for a in range(Na):
            if not ((a + 1 < Na or bc[0]) and (b + 1 < Nb or bc[1])):
                continue
            an, bn = (a + 1) % Na, (b + 1) % Nb
            i0, ia, ib, iab = a + Na * b, an + Na * b, a + Na * bn, an + Na * bn
            for sg, (i1, i2, i3), (v1, v2, v3) in ((sign[0], (ia, ib, iab), (s[b, an], s[bn, a], s[bn, an])),
                                                   (sign[1], (ia, ib, i0), (s[b, an], s[bn, a], s[b, a]))):
                q = sg / (4 * np.pi) * solid_angle(v1, v2, v3)
                per_triangle[tuple(sorted((i1, i2, i3)))] = q
                total += q
    return total, per_triangle


# ---------------------------------------------------------------------------------------------------------------------------
# Dipole-dipole gradient as the plain O(N^2) sum over an open simple-cubic lattice (Gradient_DDI_Direct,
# src/engine/Hamiltonian_Heisenberg.cpp:1016-1071):  g_i -= mu_i C sum_j mu_j (3 (s_j.r) r / r^5 - s_j / r^3),
# C = mu_0 mu_B^2 / (4 pi 1e-30), r in Angstrom. What the zero-padded FFT convolution of the product must reproduce.
# ---------------------------------------------------------------------------------------------------------------------------
def ddi_gradient_direct(spins, n_cells, mu_s=2.0, lattice_constant=1.0):
    Na, Nb, Nc = n_cells
    mu_0 = 2.0133545e-28  # T^2 m^3 / meV (include/utility/Constants.hpp)
    C = mu_0 * mu_B ** 2 / (4 * np.pi * 1e-30)
    idx = np.arange(Na * Nb * Nc)
    pos = np.stack([idx % Na, (idx // Na) % Nb, idx // (Na * Nb)], axis=1) * float(lattice_constant)
    s = np.asarray(spins, dtype=float)
    g = np.zeros_like(s)
    for i in range(len(s)):
        r = pos - pos[i]
        d = np.linalg.norm(r, axis=1)
        d[i] = np.inf
        sr = np.einsum("ij,ij->i", s, r)
        g[i] = -mu_s * mu_s * C * ((3 * sr / d ** 5)[:, None] * r - s / (d ** 3)[:, None]).sum(axis=0)
    return g


def topological_charge_basis(spins, n_cells, bc, basis, ta, tb):
    """The same for a cell with several basis atoms (lattice coordinates `basis`, atom 0 at the origin): Delaunay triangulation
    of the basis atoms and the corners a+b, b, a of the cell (corners stretched by 10 %, Vectormath.cpp:516-548), every
    triple whose circumcircle holds no other point; site order ib + NB (a + Na b). Returns (total, {sorted triple: charge})."""
    import itertools
    Na, Nb = n_cells
    NB = len(basis)
    ta, tb = np.asarray(ta, dtype=float), np.asarray(tb, dtype=float)
    s = np.asarray(spins, dtype=float).reshape(Nb, Na, NB, 3)
    pos = [b[0] * ta + b[1] * tb for b in basis]
    k = 0.1
    pts = [p.copy() for p in pos]
    pts[0] = pts[0] - k * (ta + tb)
    pts += [ta + tb + pos[0] + k * (ta + tb), tb + pos[0] - k * (ta - tb), ta + pos[0] + k * (ta - tb)]
    triangles = []
    for i, j, l in itertools.combinations(range(len(pts)), 3):
        A, B, C = pts[i], pts[j], pts[l]
        d = 2 * (A[0] * (B[1] - C[1]) + B[0] * (C[1] - A[1]) + C[0] * (A[1] - B[1]))
        if abs(d) < 1e-12:
            continue
        ux = ((A @ A) * (B[1] - C[1]) + (B @ B) * (C[1] - A[1]) + (C @ C) * (A[1] - B[1])) / d
        uy = ((A @ A) * (C[0] - B[0]) + (B @ B) * (A[0] - C[0]) + (C @ C) * (B[0] - A[0])) / d
        centre = np.array([ux, uy])
        r2 = (A - centre) @ (A - centre)
        if all((pts[m] - centre) @ (pts[m] - centre) > r2 * (1 + 1e-9) for m in range(len(pts)) if m not in (i, j, l)):
            nz = (A[0] - B[0]) * (A[1] - C[1]) - (A[1] - B[1]) * (A[0] - C[0])
            triangles.append(((i, j, l), 1.0 if nz > 0 else -1.0))
    total, per_triangle = 0.0, {}
    for (verts, sign) in triangles:
        for b in range(Nb):
            for a in range(Na):
                a_ok, b_ok = (a + 1 < Na or bc[0]), (b + 1 < Nb or bc[1])
                an, bn = (a + 1) % Na, (b + 1) % Nb
                sites, vecs = [], []
                for v in verts:
                    if v < NB:
                        sites.append(v + NB * (a + Na * b)), vecs.append(s[b, a, v])
                    elif v == NB + 2 and a_ok:
                        sites.append(NB * (an + Na * b)), vecs.append(s[b, an, 0])
                    elif v == NB + 1 and b_ok:
                        sites.append(NB * (a + Na * bn)), vecs.append(s[bn, a, 0])
                    elif v == NB and a_ok and b_ok:
                        sites.append(NB * (an + Na * bn)), vecs.append(s[bn, an, 0])
                    else:
                        sites = None
                        break
                if sites:
                    q = sign / (4 * np.pi) * solid_angle(*vecs)
                    per_triangle[tuple(sorted(sites))] = q
                    total += q
    return total, per_triangle
