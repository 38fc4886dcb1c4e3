// Plain-old-data parameter blocks shared between the host library and the CUDA kernels.
// They are passed BY VALUE as __grid_constant__ kernel parameters, so every thread reads them
// through the constant bank (uniform loads, no global traffic, no per-image __constant__ symbol
// to keep in sync when several images with different Hamiltonians live in one process).
#pragma once

#include <cstdint>

namespace sb
{
namespace dev
{

// Solver ids of the C API (core/include/Spirit/Simulation.h:33-54)
enum Solver
{
    Solver_VP      = 0,
    Solver_SIB     = 1,
    Solver_Depondt = 2,
    Solver_Heun    = 3,
    Solver_RK4     = 4,
    Solver_LBFGS_OSO   = 5,
    Solver_LBFGS_Atlas = 6,
    Solver_VP_OSO      = 7
};

constexpr int MAX_BASIS = 8;   // basis atoms per cell handled by the stencil kernels
constexpr int MAX_NEIGH = 160; // merged (exchange + DMI) neighbour entries over all basis atoms
constexpr int MAX_ANISO = 16;  // uniaxial anisotropy entries

// One gathered neighbour of basis atom `ib`: the spin (jb, a+da, b+db, c+dc) contributes
//     g_i -= J * s_j + s_j x D          (D = D_magnitude * normal)
// which merges the reference's Gradient_Exchange and Gradient_DMI loops
// (core/src/engine/Hamiltonian_Heisenberg.cpp:822-864) for pairs with identical (i, j, translations).
struct Neighbour
{
    int jb;
    int da, db, dc;
    double J;
    double Dx, Dy, Dz;
};

struct Anisotropy
{
    int ib;
    int flags; // bit k set: component k of the normal is non-zero (lets the kernels skip the zero terms)
    double K;
    double nx, ny, nz;
};

// Everything the gradient / energy stencil needs. Site order is the reference's
// (core/include/engine/Vectormath.hpp:61-74): x = ib + NB*a is the contiguous index of a row,
// rows are ordered b + Nb*c.
struct StencilParams
{
    int Na, Nb, Nc, NB;
    int bc[3]; // periodic (1) or open (0) along a, b, c -- idx_from_pair, Vectormath.hpp:437-528
    int n_neigh;
    int neigh_begin[MAX_BASIS + 1]; // entries of basis atom ib are neigh[neigh_begin[ib] .. neigh_begin[ib+1])
    int n_aniso;
    int has_cubic;
    int has_zeeman;
    int has_ddi; // a precomputed DDI gradient field is added (device/ddi_fft.cu)

    // Slab decomposition along c (multi-GPU): this device stores planes [c_begin - halo, c_begin + nc_local + halo).
    // Single device: c_begin = 0, nc_local = Nc, halo = 0 and periodic c wraps locally.
    int c_begin, nc_local, halo;
    int plane_stride; // storage sites per plane: Na*NB*Nb rounded up to a multiple of 32

    // Nearest-neighbour ("7-point") structure, detected on the host: one basis atom and the merged neighbour list is
    // a subset of {+-a, +-b, +-c} with J(+) == J(-) and D(+) == -D(-). Served by the marching kernels of sc6.cuh.
    int sc6;
    int sc6_axis[3];   // neighbours along a / b / c present
    int sc6_dflags[3]; // bit k set: component k of sc6_D[axis] is non-zero
    int sc6_extras;    // has_cubic || has_ddi: rare terms behind one flag
    int sc6_aniso_full; // sc6_A has off-diagonal elements
    double sc6_J[3];
    double sc6_nJ[3];   // -J
    double sc6_D[3][3]; // D_magnitude * normal of the +direction neighbour of each axis
    double sc6_A[6];    // on-site quadratic form -2 sum_k K_k n_k n_k^T of basis atom 0: xx, yy, zz, xy, xz, yz
    double sc6_g0[3];   // -mu_s B n: start value of the gradient accumulation

    // Pinned sites and defects (Geometry::site_flags; SITE_* bits below), one byte per STORAGE index; null on a lattice without
    // either. With flags the nearest-neighbour kernels step aside (sc6 = 0) and the generic kernels test them:
    // a vacancy takes part in no interaction (check_atom_type / idx_from_pair, Vectormath.hpp:406-528), a site without moment
    // has no Zeeman and no dipolar term, force and virtual force of a pinned site are zero (Method_LLG.cpp:122-124, 222-224).
    const unsigned char * site_flags;

    double K4[MAX_BASIS];        // cubic anisotropy per basis atom (Hamiltonian_Heisenberg.cpp:802-820)
    double zeeman[MAX_BASIS][3]; // mu_s[ib] * (B mu_B) * n_B   (Hamiltonian_Heisenberg.cpp:768-783)
    double mu_s[MAX_BASIS];
    Anisotropy aniso[MAX_ANISO];
    Neighbour neigh[MAX_NEIGH];
};

constexpr unsigned FLAG_VACANT = 1, FLAG_PINNED = 2, FLAG_NO_MU_S = 4; // = sb::SITE_* (core/geometry.hpp)

// LLG virtual-force parameters (core/src/engine/Method_LLG.cpp:131-226)
struct LLGParams
{
    double dtg;       // dt*gamma/mu_B/(1+alpha^2)   (dynamics)   or   dt*gamma/mu_B (direct minimisation)
    double damping;   // alpha
    double dt;        // raw llg_dt (VP uses it directly, Solver_VP.hpp:95-110)
    int direct_minimization; // Fv = dtg * s x F
    int has_stt;      // 1: monolayer spin-transfer torque (Method_LLG.cpp:207-212), 2: gradient approximation (:184-205)
    double stt_c1;    // -dtg*a_j*(alpha-beta)
    double stt_c2;    // -dtg*a_j*(1+beta*alpha)
    double stt_pol[3];
    // gradient approximation: s_c_grad = jacobian(s) je = sum_t stt_w[t] f_t (s(+t) - s(-t)), t = a, b, c lattice translations,
    // stt_w = (lattice_constant [ta tb tc])^-1 je, f_t = 1/2 (central) or 1 (one-sided at an open boundary) -- Vectormath.cpp:816-903;
    // Fv += stt_g1 s_c_grad + stt_g2 s_c_grad x s with stt_g1 = dtg a_j (alpha - beta), stt_g2 = dtg a_j (1 + beta alpha)
    double stt_w[3];
    double stt_g1, stt_g2;
    int has_thermal;  // Method_LLG.cpp:65-110,215-219
    int pad;
    double thermal_scale[MAX_BASIS]; // epsilon*sqrt(T/mu_s[ib])
    double inv_mu_s[MAX_BASIS];
    double c1[MAX_BASIS];    // dtg / mu_s[ib]
    double c2[MAX_BASIS];    // alpha * dtg / mu_s[ib]
    double nc1[MAX_BASIS];   // -c1
    double nc2[MAX_BASIS];   // -c2
    double half_nc1[MAX_BASIS]; // -c1/2, -c2/2, alpha/2, -dtg/2: Depondt's corrector averages two virtual forces
    double half_nc2[MAX_BASIS];
    double half_damping;
    double half_ndtg;
    float thermal_k[MAX_BASIS]; // -2 ln2 thermal_scale^2: Box-Muller radius sqrt(k lg2 u) already carries the amplitude
    // Linear temperature gradient (Method_LLG.cpp:80-96, Vectormath.cpp:633-652): T_i = clamp( tgrad_T0 + tgrad_cell .
    // (a, b, c_global) + tgrad_basis[ib], 0, 1e30 ); the radius constant becomes thermal_k_per_T[ib] * T_i
    int has_tgrad;
    int pad2;
    double tgrad_T0;
    double tgrad_cell[3];
    double tgrad_basis[MAX_BASIS];
    float thermal_k_per_T[MAX_BASIS]; // -2 ln2 epsilon^2 / mu_s[ib]
    unsigned philox_key[10][2]; // round keys of Philox4x32-10: (seed_lo + r W0, seed_hi + r W1)
    std::uint64_t seed;      // Philox key
    std::uint64_t iteration; // Philox counter high words: one xi per iteration, shared by all stages
};

} // namespace dev
} // namespace sb
