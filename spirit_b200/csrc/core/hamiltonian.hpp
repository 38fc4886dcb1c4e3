// Host-side description of the Heisenberg Hamiltonian: parameters and pair lists.
// Semantics follow Engine::Hamiltonian_Heisenberg (core/include/engine/Hamiltonian_Heisenberg.hpp:31-229,
// core/src/engine/Hamiltonian_Heisenberg.cpp:26-260). The arithmetic lives on the device
// (device/hamiltonian_kernels.cu); this class only builds the tables the kernels consume.
//
// Out of scope: quadruplets, Hessians, DDI cutoff.
#pragma once

#include "geometry.hpp"

#include <memory>
#include <string>

namespace sb
{

enum class DDI_Method
{
    None   = 0, // SPIRIT_DDI_METHOD_NONE   (Hamiltonian.h:48-57)
    FFT    = 1,
    FMM    = 2,
    Cutoff = 3
};

struct Hamiltonian
{
    Hamiltonian( std::shared_ptr<Geometry> geometry );

    std::shared_ptr<Geometry> geometry;
    std::array<int, 3> boundary_conditions{ 0, 0, 0 };

    // Zeeman. As in the reference the stored magnitude is already multiplied by mu_B
    // (Hamiltonian_Heisenberg.cpp:35): units meV per mu_B.
    double external_field_magnitude = 0;
    Vec3 external_field_normal{ 0, 0, 1 };

    // Uniaxial / cubic anisotropy, indexed by basis atom
    intfield anisotropy_indices;
    scalarfield anisotropy_magnitudes;
    vectorfield anisotropy_normals;
    intfield cubic_anisotropy_indices;
    scalarfield cubic_anisotropy_magnitudes;

    // Exchange: either shells or an explicit (symmetry-reduced) pair list
    scalarfield exchange_shell_magnitudes;
    pairfield exchange_pairs_in;
    scalarfield exchange_magnitudes_in;
    pairfield exchange_pairs; // redundant (both directions), as in the reference's OpenMP / CUDA builds
    scalarfield exchange_magnitudes;

    // DMI
    scalarfield dmi_shell_magnitudes;
    int dmi_shell_chirality = 0;
    pairfield dmi_pairs_in;
    scalarfield dmi_magnitudes_in;
    vectorfield dmi_normals_in;
    pairfield dmi_pairs;
    scalarfield dmi_magnitudes;
    vectorfield dmi_normals;

    // Dipole-dipole
    DDI_Method ddi_method = DDI_Method::None;
    std::array<int, 3> ddi_n_periodic_images{ 4, 4, 4 };
    bool ddi_pb_zero_padding = true;
    double ddi_cutoff_radius = 0;

    // Rebuild the (redundant) pair lists from shells or input pairs (Hamiltonian_Heisenberg.cpp:101-198).
    // Increments `revision`, which the device side uses to notice that its tables are stale.
    void Update_Interactions();
    // Which terms contribute, in the reference's order (Hamiltonian_Heisenberg.cpp:200-260)
    void Update_Energy_Contributions();

    std::vector<std::string> contribution_names;
    int idx_zeeman = -1, idx_anisotropy = -1, idx_cubic_anisotropy = -1, idx_exchange = -1, idx_dmi = -1,
        idx_ddi = -1;

    std::uint64_t revision = 0;
};

} // namespace sb
