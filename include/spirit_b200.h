/* spirit_b200 extensions to the Spirit C API (plain C ABI, no torch types).
 *
 * (1) Double-precision probes. The reference API narrows energies, torques and magnetisation to float
 *     (core/include/Spirit/System.h:55, Quantities.h:21, Simulation.h:58-71, Chain.h:86-96), which cannot carry a
 *     1e-12 parity check. The reference's own tests reach below the API for that
 *     (core/test/test_anisotropy.cpp:137-149 calls hamiltonian->Gradient_and_Energy on doubles); these probes are
 *     the same thing for this library. Every probe has a twin `refshim_*` with the same signature in
 *     oracle/ref_shim.cpp that calls the unmodified reference engine.
 * (2) Device control and device-resident stepping for benchmarks: run iterations on spins that already live in
 *     HBM and time them with CUDA events on the image's own stream.
 *
 * All functions return a negative value (or 0.0) on error and never throw.
 */
#ifndef SPIRIT_B200_H
#define SPIRIT_B200_H
#include "Spirit/Export.h"
struct State;
typedef struct State State;

/* --- (1) probes --------------------------------------------------------------------------------------------- */
/* gradient[nos][3] and total energy of `spins` ([nos][3], NULL: the image's own spins).
 * replaces Engine::Hamiltonian_Heisenberg::Gradient_and_Energy, core/src/engine/Hamiltonian_Heisenberg.cpp:704-766 */
SPIRIT_API int SpiritB200_Gradient_and_Energy( State * state, const double * spins, double * gradient, double * energy, int idx_image ) SPIRIT_NOEXCEPT;
/* replaces Hamiltonian_Heisenberg::Gradient, Hamiltonian_Heisenberg.cpp:670-702 */
SPIRIT_API int SpiritB200_Gradient( State * state, const double * spins, double * gradient, int idx_image ) SPIRIT_NOEXCEPT;
/* per-term energies: names [max_terms][32], totals[max_terms], per_spin (nullable) [n_terms][nos]; returns n_terms.
 * replaces Energy_Contributions_per_Spin, Hamiltonian_Heisenberg.cpp:262-404 */
SPIRIT_API int SpiritB200_Energy_Contributions( State * state, const double * spins, int max_terms, char * names, double * totals, double * per_spin, int idx_image ) SPIRIT_NOEXCEPT;
/* image->E in double */
SPIRIT_API double SpiritB200_Get_Energy( State * state, int idx_image ) SPIRIT_NOEXCEPT;
/* (redundant) pair lists after Update_Interactions (Hamiltonian_Heisenberg.cpp:101-198). kind 0 exchange, 1 DMI.
 * ijt [max][5] = i j da db dc; magnitudes[max]; normals[max][3] (DMI). Returns the number of pairs. Host only. */
SPIRIT_API int SpiritB200_Get_Pairs( State * state, int kind, int max_pairs, int * ijt, double * magnitudes, double * normals, int idx_image ) SPIRIT_NOEXCEPT;
/* max torque of the running / last method on the image (idx_image == -2: the chain method), in double */
SPIRIT_API double SpiritB200_Get_MaxTorque( State * state, int idx_image ) SPIRIT_NOEXCEPT;
/* reaction coordinate and energies of all images in double; returns noi */
SPIRIT_API int SpiritB200_Chain_Get_Rx_E( State * state, double * Rx, double * E ) SPIRIT_NOEXCEPT;
/* mean of mu_s*s in double (Vectormath::Magnetization, core/src/engine/Vectormath.cpp:495-502) */
SPIRIT_API int SpiritB200_Get_Magnetization( State * state, double * m, int idx_image ) SPIRIT_NOEXCEPT;

/* --- (2) device ------------------------------------------------------------------------------------------------ */
SPIRIT_API int SpiritB200_Device_Count( void ) SPIRIT_NOEXCEPT;          /* 0 when there is no usable CUDA device */
SPIRIT_API int SpiritB200_Set_Device( int device ) SPIRIT_NOEXCEPT;      /* device for subsequently created images */
SPIRIT_API const char * SpiritB200_Device_Name( void ) SPIRIT_NOEXCEPT;
/* Number of kernels this library launched for the image since it was created */
SPIRIT_API unsigned long long SpiritB200_Kernel_Launches( State * state, int idx_image ) SPIRIT_NOEXCEPT;
/* Which stage kernels serve the image's Hamiltonian: 1 the nearest-neighbour marching kernels, 0 the generic gather
 * kernels; < 0 on error. No counterpart in the reference (its CUDA backend has one kernel per term). */
SPIRIT_API int SpiritB200_Stencil_Variant( State * state, int idx_image ) SPIRIT_NOEXCEPT;
/* Test probe: 3 * count normal variates of unit standard deviation exactly as the stage kernels draw them for the thermal
 * field (Philox4x32-10 keyed by the image's llg_seed, counters 0 .. count - 1 of iteration `iteration`, Box-Muller). The
 * reference draws std::normal_distribution<double> from one serial mt19937 (Method_LLG.cpp:98-108): only the distribution
 * can be compared. Returns 0, or < 0 on error. */
SPIRIT_API int SpiritB200_Thermal_Variates( State * state, unsigned long long iteration, unsigned long long count, float * variates, int idx_image ) SPIRIT_NOEXCEPT;
/* Which kernels run one iteration of `solver_type` on the image with its current Hamiltonian and LLG parameters:
 * 2 ONE fused predictor + corrector kernel per iteration (Depondt / Heun / SIB on the nearest-neighbour stencil: spins read
 * once and written once), 1 one marching kernel per solver stage, 0 the generic gather kernels; < 0 on error. */
SPIRIT_API int SpiritB200_Step_Variant( State * state, int solver_type, int idx_image ) SPIRIT_NOEXCEPT;
/* Upload the image's host spins to HBM (and build the device tables) / download spins + effective field */
SPIRIT_API int SpiritB200_Upload( State * state, int idx_image ) SPIRIT_NOEXCEPT;
SPIRIT_API int SpiritB200_Download( State * state, int idx_image ) SPIRIT_NOEXCEPT;
/* Run n_iterations of an LLG solver on the device-resident spins of the image (no host<->device copies, no hooks
 * except after the last iteration) and return the elapsed milliseconds measured with CUDA events on the image's
 * stream; < 0 on error. The image's LLG parameters (dt, damping, temperature, seed, ...) apply.
 * Same arithmetic as Simulation_LLG_Start: Method_Solver<solver>::Iteration, core/include/engine/Solver_*.hpp */
SPIRIT_API double SpiritB200_LLG_Iterate_Device( State * state, int solver_type, int n_iterations, int idx_image ) SPIRIT_NOEXCEPT;
/* --- (3) slab decomposition over several GPUs, one process per GPU ------------------------------------------------------
 * The pair stencils shard along c: rank r holds the planes [c_begin, c_begin + nc_local) of a lattice with Nc_global
 * planes; its State is set up with the LOCAL slab (n_basis_cells a b nc_local). After every kernel that writes a
 * configuration the first / last plane travels to the neighbouring ranks over NCCL (NVLink); energies, torques and the
 * VP projections are all-reduced. The thermal noise is keyed by (site in plane, GLOBAL plane), so results do not depend on the
 * decomposition. The reference has no multi-GPU path; these entry points have no counterpart there. */
/* 128-byte NCCL unique id, to be created on rank 0 and distributed by the launcher */
SPIRIT_API int SpiritB200_Comm_Unique_Id( char * id128 ) SPIRIT_NOEXCEPT;
SPIRIT_API int SpiritB200_Comm_Init( int rank, int world, const char * id128 ) SPIRIT_NOEXCEPT;
/* Declare the image to be the slab starting at plane c_begin of a lattice with Nc_global planes. Call before any compute. */
SPIRIT_API int SpiritB200_Slab_Setup( State * state, int c_begin, int Nc_global, int idx_image ) SPIRIT_NOEXCEPT;

/* GNEB: whole images per GPU. This process's chain holds the images [i_begin, i_begin + noi) of a chain of noi_global
 * images; per force evaluation the first / last local image travels to the neighbouring rank (tangents, geodesic
 * distances), per-image scalars (E, Rx, projections) are all-reduced, VP's projections are summed over all images. */
SPIRIT_API int SpiritB200_Chain_Shard_Setup( State * state, int i_begin, int noi_global ) SPIRIT_NOEXCEPT;

/* The same iterations with a CUDA event between the stage kernels: stage_ms[k] = mean milliseconds of stage k+1
 * (Depondt/Heun/SIB: 2 stages, RK4: 4). Returns the number of stages, < 0 on error. For per-kernel rooflines. */
SPIRIT_API int SpiritB200_LLG_Profile_Stages( State * state, int solver_type, int n_iterations, double * stage_ms, int max_stages, int idx_image ) SPIRIT_NOEXCEPT;
/* Host-side probe: the triangles (3 site indices each) the topological charge is summed over (Vectormath.cpp:504-631:
 * Delaunay triangulation of the basis cell, repeated over the lattice with the boundary rule). Returns their number;
 * triangle_indices may be NULL. No device needed. */
SPIRIT_API int SpiritB200_Topology_Triangles( State * state, int * triangle_indices, int idx_image ) SPIRIT_NOEXCEPT;
#endif
