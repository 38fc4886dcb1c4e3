// C API: Configuration_* initialisers. Reference behaviour: core/src/Spirit/Configurations.cpp:120-700.
// They fill the image's host spins (the live array System_Get_Spin_Directions exposes); the next
// Simulation_* / System_Update_Data call uploads them to the GPU.
#include "api_common.hpp"

#include <Spirit/Configurations.h>

using namespace sb;
using configurations::get_filter;

namespace
{
Vec3 v3( const float * p )
{
    return Vec3{ p[0], p[1], p[2] };
}
} // namespace

void Configuration_To_Clipboard( State * state, int idx_image, int idx_chain ) noexcept
try
{
    auto image             = resolve( state, idx_image, idx_chain ).image;
    state->clipboard_spins = std::make_shared<std::vector<Vec3>>( image->spins.data(), image->spins.data() + image->nos );
}
SB_API_CATCH_VOID

void Configuration_From_Clipboard(
    State * state, const float position[3], const float r_cut_rectangular[3], float r_cut_cylindrical, float r_cut_spherical,
    bool inverted, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    if( !state->clipboard_spins )
    {
        Log( Log_Level::Warning, Log_Sender::API, "Tried to insert configuration, but clipboard was empty.", idx_image, idx_chain );
        return;
    }
    const Vec3 vpos = image->geometry->center + v3( position );
    auto filter     = get_filter( vpos, r_cut_rectangular, r_cut_cylindrical, r_cut_spherical, inverted );
    ImageLock lock( *image );
    PinGuard pin( *image );
    configurations::Insert( *image, *state->clipboard_spins, 0, filter );
}
SB_API_CATCH_VOID

bool Configuration_From_Clipboard_Shift(
    State * state, const float shift[3], const float position[3], const float r_cut_rectangular[3], float r_cut_cylindrical,
    float r_cut_spherical, bool inverted, int idx_image, int idx_chain ) noexcept
try
{
    if( !state || !state->clipboard_spins )
    {
        Log( Log_Level::Warning, Log_Sender::API, "Tried to insert configuration, but clipboard was empty.", idx_image, idx_chain );
        return false;
    }
    auto image = resolve( state, idx_image, idx_chain ).image;
    auto & g   = *image->geometry;
    // decompose the shift into lattice translations: solve [a b c] x = shift (Cramer)
    const Vec3 a = g.bravais_vectors[0], b = g.bravais_vectors[1], c = g.bravais_vectors[2], s = v3( shift );
    const double det = a.dot( b.cross( c ) );
    const int da     = int( std::round( s.dot( b.cross( c ) ) / det ) );
    const int db     = int( std::round( a.dot( s.cross( c ) ) / det ) );
    const int dc     = int( std::round( a.dot( b.cross( s ) ) / det ) );
    if( da == 0 && db == 0 && dc == 0 )
        return false;
    const int delta = g.n_cell_atoms * da + g.n_cell_atoms * g.n_cells[0] * db + g.n_cell_atoms * g.n_cells[0] * g.n_cells[1] * dc;
    auto filter     = get_filter( v3( position ), r_cut_rectangular, r_cut_cylindrical, r_cut_spherical, inverted );
    ImageLock lock( *image );
    PinGuard pin( *image );
    configurations::Insert( *image, *state->clipboard_spins, delta, filter );
    return true;
}
SB_API_CATCH_RET( false )

void Configuration_Domain(
    State * state, const float direction[3], const float position[3], const float r_cut_rectangular[3], float r_cut_cylindrical,
    float r_cut_spherical, bool inverted, int idx_image, int idx_chain ) noexcept
try
{
    auto image      = resolve( state, idx_image, idx_chain ).image;
    const Vec3 vpos = image->geometry->center + v3( position );
    auto filter     = get_filter( vpos, r_cut_rectangular, r_cut_cylindrical, r_cut_spherical, inverted );
    ImageLock lock( *image );
    PinGuard pin( *image );
    configurations::Domain( *image, v3( direction ), filter );
}
SB_API_CATCH_VOID

void Configuration_PlusZ(
    State * state, const float position[3], const float r_cut_rectangular[3], float r_cut_cylindrical, float r_cut_spherical,
    bool inverted, int idx_image, int idx_chain ) noexcept
{
    const float dir[3] = { 0, 0, 1 };
    Configuration_Domain( state, dir, position, r_cut_rectangular, r_cut_cylindrical, r_cut_spherical, inverted, idx_image, idx_chain );
}

void Configuration_MinusZ(
    State * state, const float position[3], const float r_cut_rectangular[3], float r_cut_cylindrical, float r_cut_spherical,
    bool inverted, int idx_image, int idx_chain ) noexcept
{
    const float dir[3] = { 0, 0, -1 };
    Configuration_Domain( state, dir, position, r_cut_rectangular, r_cut_cylindrical, r_cut_spherical, inverted, idx_image, idx_chain );
}

void Configuration_Random(
    State * state, const float position[3], const float r_cut_rectangular[3], float r_cut_cylindrical, float r_cut_spherical,
    bool inverted, bool, int idx_image, int idx_chain ) noexcept
try
{
    auto image      = resolve( state, idx_image, idx_chain ).image;
    const Vec3 vpos = image->geometry->center + v3( position );
    auto filter     = get_filter( vpos, r_cut_rectangular, r_cut_cylindrical, r_cut_spherical, inverted );
    ImageLock lock( *image );
    PinGuard pin( *image );
    // `external` makes no difference in the reference either (Configurations.cpp:99-129 uses the LLG prng in both branches)
    configurations::Random( *image, filter );
}
SB_API_CATCH_VOID

void Configuration_SpinSpiral(
    State * state, const char * direction_type, float q[3], float axis[3], float theta, const float position[3],
    const float r_cut_rectangular[3], float r_cut_cylindrical, float r_cut_spherical, bool inverted, int idx_image,
    int idx_chain ) noexcept
try
{
    auto image      = resolve( state, idx_image, idx_chain ).image;
    const Vec3 vpos = image->geometry->center + v3( position );
    auto filter     = get_filter( vpos, r_cut_rectangular, r_cut_cylindrical, r_cut_spherical, inverted );
    ImageLock lock( *image );
    PinGuard pin( *image );
    configurations::SpinSpiral( *image, direction_type ? direction_type : "", v3( q ), v3( axis ), theta, filter );
}
SB_API_CATCH_VOID

void Configuration_SpinSpiral_2q(
    State *, const char *, float[3], float[3], float[3], float, const float[3], const float[3], float, float, bool, int idx_image,
    int idx_chain ) noexcept
{
    Log( Log_Level::Error, Log_Sender::API, "Configuration_SpinSpiral_2q is not implemented in spirit_b200", idx_image, idx_chain );
}

void Configuration_Add_Noise_Temperature(
    State * state, float temperature, const float position[3], const float r_cut_rectangular[3], float r_cut_cylindrical,
    float r_cut_spherical, bool inverted, int idx_image, int idx_chain ) noexcept
try
{
    auto image      = resolve( state, idx_image, idx_chain ).image;
    const Vec3 vpos = image->geometry->center + v3( position );
    auto filter     = get_filter( vpos, r_cut_rectangular, r_cut_cylindrical, r_cut_spherical, inverted );
    ImageLock lock( *image );
    PinGuard pin( *image );
    configurations::Add_Noise_Temperature( *image, temperature, 0, filter );
}
SB_API_CATCH_VOID

void Configuration_Displace_Eigenmode( State *, int, int idx_image, int idx_chain ) noexcept
{
    Log( Log_Level::Error, Log_Sender::API, "Configuration_Displace_Eigenmode: eigenmodes are outside the hot path of spirit_b200", idx_image, idx_chain );
}

void Configuration_Skyrmion(
    State * state, float r, float order, float phase, bool upDown, bool achiral, bool rl, const float position[3],
    const float r_cut_rectangular[3], float r_cut_cylindrical, float r_cut_spherical, bool inverted, int idx_image,
    int idx_chain ) noexcept
try
{
    auto image      = resolve( state, idx_image, idx_chain ).image;
    const Vec3 vpos = image->geometry->center + v3( position );
    if( r_cut_cylindrical < 0 )
        r_cut_cylindrical = r;
    auto filter = get_filter( vpos, r_cut_rectangular, r_cut_cylindrical, r_cut_spherical, inverted );
    ImageLock lock( *image );
    PinGuard pin( *image );
    configurations::Skyrmion( *image, vpos, r, order, phase, upDown, achiral, rl, filter );
}
SB_API_CATCH_VOID

void Configuration_DW_Skyrmion(
    State * state, float dw_radius, float dw_width, float order, float phase, bool upDown, bool achiral, bool rl,
    const float position[3], const float r_cut_rectangular[3], float r_cut_cylindrical, float r_cut_spherical, bool inverted,
    int idx_image, int idx_chain ) noexcept
try
{
    auto image      = resolve( state, idx_image, idx_chain ).image;
    const Vec3 vpos = image->geometry->center + v3( position );
    if( r_cut_cylindrical < 0 )
        r_cut_cylindrical = std::max( 3 * dw_radius, 3 * dw_width );
    auto filter = get_filter( vpos, r_cut_rectangular, r_cut_cylindrical, r_cut_spherical, inverted );
    ImageLock lock( *image );
    PinGuard pin( *image );
    configurations::DW_Skyrmion( *image, vpos, dw_radius, dw_width, order, phase, upDown, achiral, rl, filter );
}
SB_API_CATCH_VOID

void Configuration_Hopfion(
    State * state, float r, int order, const float position[3], const float r_cut_rectangular[3], float r_cut_cylindrical,
    float r_cut_spherical, bool inverted, const float normal[3], int idx_image, int idx_chain ) noexcept
try
{
    auto image      = resolve( state, idx_image, idx_chain ).image;
    const Vec3 vpos = image->geometry->center + v3( position );
    if( r_cut_spherical < 0 )
        r_cut_spherical = r * float( constants::Pi ); // Configurations.cpp: the hopfion fills a sphere of radius pi*r
    auto filter = get_filter( vpos, r_cut_rectangular, r_cut_cylindrical, r_cut_spherical, inverted );
    ImageLock lock( *image );
    PinGuard pin( *image );
    configurations::Hopfion( *image, vpos, r, order, v3( normal ), filter );
}
SB_API_CATCH_VOID

// Pinning (Configurations.cpp:697-726, Utility::Configurations::Set_Pinned): the filtered sites are (un)pinned at their current
// orientation
void Configuration_Set_Pinned(
    State * state, bool pinned, const float position[3], const float r_cut_rectangular[3], float r_cut_cylindrical, float r_cut_spherical,
    bool inverted, int idx_image, int idx_chain ) noexcept
try
{
    auto image      = resolve( state, idx_image, idx_chain ).image;
    const Vec3 vpos = image->geometry->center + v3( position );
    auto filter     = get_filter( vpos, r_cut_rectangular, r_cut_cylindrical, r_cut_spherical, inverted );
    ImageLock lock( *image );
    const auto & positions = image->geometry->positions();
    for( int i = 0; i < image->nos; ++i )
        if( filter( image->spins[i], positions[i] ) )
            image->geometry->set_pinned( i, pinned, image->spins[i] );
    Log( Log_Level::Info, Log_Sender::API, "Set pinned spins.", idx_image, idx_chain );
}
SB_API_CATCH_VOID

// Defects (Configurations.cpp:728-758, Utility::Configurations::Set_Atom_Types): type < 0 makes the filtered sites vacancies
void Configuration_Set_Atom_Type(
    State * state, int atom_type, const float position[3], const float r_cut_rectangular[3], float r_cut_cylindrical,
    float r_cut_spherical, bool inverted, int idx_image, int idx_chain ) noexcept
try
{
    auto image      = resolve( state, idx_image, idx_chain ).image;
    const Vec3 vpos = image->geometry->center + v3( position );
    auto filter     = get_filter( vpos, r_cut_rectangular, r_cut_cylindrical, r_cut_spherical, inverted );
    ImageLock lock( *image );
    const auto & positions = image->geometry->positions();
    for( int i = 0; i < image->nos; ++i )
        if( filter( image->spins[i], positions[i] ) )
            image->geometry->set_atom_type( i, atom_type );
    Log( Log_Level::Info, Log_Sender::API, "Set atom types to " + std::to_string( atom_type ) + ".", idx_image, idx_chain );
}
SB_API_CATCH_VOID
