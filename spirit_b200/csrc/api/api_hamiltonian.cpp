// C API: Hamiltonian setters / getters. Reference behaviour: core/src/Spirit/Hamiltonian.cpp:30-700.
// Setters narrow through `float` exactly like the reference (SURVEY.md 8c hazard 5), lock the image and bump the
// Hamiltonian revision, so the device tables are rebuilt before the next kernel launch.
#include "api_common.hpp"

#include <Spirit/Hamiltonian.h>

using namespace sb;

void Hamiltonian_Set_Boundary_Conditions( State * state, const bool * periodical, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    for( int d = 0; d < 3; ++d )
        image->hamiltonian->boundary_conditions[d] = periodical[d] ? 1 : 0;
    // Hamiltonian.cpp:30-62 calls Update_Interactions (the DDI tensor depends on the boundary conditions)
    image->hamiltonian->Update_Interactions();
}
SB_API_CATCH_VOID

void Hamiltonian_Set_Field( State * state, float magnitude, const float * normal, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    Vec3 n{ normal[0], normal[1], normal[2] };
    n.normalize();
    image->hamiltonian->external_field_magnitude = double( magnitude ) * constants::mu_B;
    image->hamiltonian->external_field_normal    = n;
    image->hamiltonian->Update_Energy_Contributions();
}
SB_API_CATCH_VOID

void Hamiltonian_Set_Anisotropy( State * state, float magnitude, const float * normal, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    const int n_cell_atoms = image->geometry->n_cell_atoms;
    Vec3 n{ normal[0], normal[1], normal[2] };
    n.normalize();
    auto & ham = *image->hamiltonian;
    ham.anisotropy_indices.resize( n_cell_atoms );
    ham.anisotropy_magnitudes.assign( n_cell_atoms, double( magnitude ) );
    ham.anisotropy_normals.assign( n_cell_atoms, n );
    for( int i = 0; i < n_cell_atoms; ++i )
        ham.anisotropy_indices[i] = i;
    ham.Update_Energy_Contributions();
}
SB_API_CATCH_VOID

void Hamiltonian_Set_Cubic_Anisotropy( State * state, float magnitude, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    const int n_cell_atoms = image->geometry->n_cell_atoms;
    auto & ham             = *image->hamiltonian;
    ham.cubic_anisotropy_indices.resize( n_cell_atoms );
    ham.cubic_anisotropy_magnitudes.assign( n_cell_atoms, double( magnitude ) );
    for( int i = 0; i < n_cell_atoms; ++i )
        ham.cubic_anisotropy_indices[i] = i;
    ham.Update_Energy_Contributions();
}
SB_API_CATCH_VOID

void Hamiltonian_Set_Exchange( State * state, int n_shells, const float * jij, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    auto & ham = *image->hamiltonian;
    ham.exchange_shell_magnitudes.assign( jij, jij + n_shells );
    ham.exchange_pairs_in.clear();
    ham.exchange_magnitudes_in.clear();
    ham.Update_Interactions();
}
SB_API_CATCH_VOID

void Hamiltonian_Set_DMI( State * state, int n_shells, const float * dij, int chirality, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    if( chirality != SPIRIT_CHIRALITY_BLOCH && chirality != SPIRIT_CHIRALITY_NEEL && chirality != SPIRIT_CHIRALITY_BLOCH_INVERSE
        && chirality != SPIRIT_CHIRALITY_NEEL_INVERSE )
    {
        Log( Log_Level::Error, Log_Sender::API, "Hamiltonian_Set_DMI: Invalid DM chirality " + std::to_string( chirality ), idx_image, idx_chain );
        return;
    }
    ImageLock lock( *image );
    auto & ham = *image->hamiltonian;
    ham.dmi_shell_magnitudes.assign( dij, dij + n_shells );
    ham.dmi_shell_chirality = chirality;
    ham.dmi_pairs_in.clear();
    ham.dmi_magnitudes_in.clear();
    ham.dmi_normals_in.clear();
    ham.Update_Interactions();
}
SB_API_CATCH_VOID

void Hamiltonian_Set_DDI(
    State * state, int ddi_method, int n_periodic_images[3], float cutoff_radius, bool pb_zero_padding, int idx_image,
    int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    auto & ham     = *image->hamiltonian;
    ham.ddi_method = DDI_Method( ddi_method );
    for( int d = 0; d < 3; ++d )
        ham.ddi_n_periodic_images[d] = n_periodic_images[d];
    ham.ddi_cutoff_radius   = cutoff_radius;
    ham.ddi_pb_zero_padding = pb_zero_padding;
    ham.Update_Interactions();
}
SB_API_CATCH_VOID

const char * Hamiltonian_Get_Name( State * state, int idx_image, int idx_chain ) noexcept
try
{
    resolve( state, idx_image, idx_chain );
    return "Heisenberg";
}
SB_API_CATCH_RET( nullptr )

void Hamiltonian_Get_Boundary_Conditions( State * state, bool * periodical, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    for( int d = 0; d < 3; ++d )
        periodical[d] = image->hamiltonian->boundary_conditions[d] != 0;
}
SB_API_CATCH_VOID

// Hamiltonian.cpp:430-470: magnitude in T; a zero field reports normal (0,0,1)
void Hamiltonian_Get_Field( State * state, float * magnitude, float * normal, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    auto & ham = *image->hamiltonian;
    if( ham.external_field_magnitude > 0 )
    {
        *magnitude = float( ham.external_field_magnitude / constants::mu_B );
        for( int d = 0; d < 3; ++d )
            normal[d] = float( ham.external_field_normal[d] );
    }
    else
    {
        *magnitude = 0;
        normal[0]  = 0;
        normal[1]  = 0;
        normal[2]  = 1;
    }
}
SB_API_CATCH_VOID

void Hamiltonian_Get_Anisotropy( State * state, float * magnitude, float * normal, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    auto & ham = *image->hamiltonian;
    if( !ham.anisotropy_indices.empty() )
    {
        *magnitude = float( ham.anisotropy_magnitudes[0] );
        for( int d = 0; d < 3; ++d )
            normal[d] = float( ham.anisotropy_normals[0][d] );
    }
    else
    {
        *magnitude = 0;
        normal[0]  = 0;
        normal[1]  = 0;
        normal[2]  = 1;
    }
}
SB_API_CATCH_VOID

void Hamiltonian_Get_Cubic_Anisotropy( State * state, float * magnitude, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    auto & ham = *image->hamiltonian;
    *magnitude = ham.cubic_anisotropy_indices.empty() ? 0.0f : float( ham.cubic_anisotropy_magnitudes[0] );
}
SB_API_CATCH_VOID

void Hamiltonian_Get_Exchange_Shells( State * state, int * n_shells, float * jij, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    auto & ham = *image->hamiltonian;
    *n_shells  = int( ham.exchange_shell_magnitudes.size() );
    for( int i = 0; i < *n_shells; ++i )
        jij[i] = float( ham.exchange_shell_magnitudes[i] );
}
SB_API_CATCH_VOID

int Hamiltonian_Get_Exchange_N_Pairs( State * state, int idx_image, int idx_chain ) noexcept
try
{
    return int( resolve( state, idx_image, idx_chain ).image->hamiltonian->exchange_pairs.size() );
}
SB_API_CATCH_RET( 0 )

void Hamiltonian_Get_Exchange_Pairs(
    State * state, int idx[][2], int translations[][3], float * Jij, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    auto & ham = *image->hamiltonian;
    for( std::size_t p = 0; p < ham.exchange_pairs.size(); ++p )
    {
        idx[p][0] = ham.exchange_pairs[p].i;
        idx[p][1] = ham.exchange_pairs[p].j;
        for( int d = 0; d < 3; ++d )
            translations[p][d] = ham.exchange_pairs[p].translations[d];
        Jij[p] = float( ham.exchange_magnitudes[p] );
    }
}
SB_API_CATCH_VOID

void Hamiltonian_Get_DMI_Shells( State * state, int * n_shells, float * dij, int * chirality, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    auto & ham = *image->hamiltonian;
    *n_shells  = int( ham.dmi_shell_magnitudes.size() );
    *chirality = ham.dmi_shell_chirality;
    for( int i = 0; i < *n_shells; ++i )
        dij[i] = float( ham.dmi_shell_magnitudes[i] );
}
SB_API_CATCH_VOID

int Hamiltonian_Get_DMI_N_Pairs( State * state, int idx_image, int idx_chain ) noexcept
try
{
    return int( resolve( state, idx_image, idx_chain ).image->hamiltonian->dmi_pairs.size() );
}
SB_API_CATCH_RET( 0 )

void Hamiltonian_Get_DDI(
    State * state, int * ddi_method, int n_periodic_images[3], float * cutoff_radius, bool * pb_zero_padding, int idx_image,
    int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    auto & ham = *image->hamiltonian;
    *ddi_method = int( ham.ddi_method );
    for( int d = 0; d < 3; ++d )
        n_periodic_images[d] = ham.ddi_n_periodic_images[d];
    *cutoff_radius   = float( ham.ddi_cutoff_radius );
    *pb_zero_padding = ham.ddi_pb_zero_padding;
}
SB_API_CATCH_VOID
