// C API: System, Quantities, Geometry.
// Reference behaviour: core/src/Spirit/System.cpp, Quantities.cpp:17-60, Geometry.cpp:13-560.
#include "api_common.hpp"

#include <Spirit/Geometry.h>
#include <Spirit/Quantities.h>
#include <Spirit/Simulation.h>
#include <Spirit/System.h>
#include <spirit_b200.h>

#include <cstring>

using namespace sb;

int System_Get_Index( State * state ) noexcept
{
    return state ? state->idx_active_image : -1;
}

int System_Get_NOS( State * state, int idx_image, int idx_chain ) noexcept
try
{
    return resolve( state, idx_image, idx_chain ).image->nos;
}
SB_API_CATCH_RET( 0 )

scalar * System_Get_Spin_Directions( State * state, int idx_image, int idx_chain ) noexcept
try
{
    return resolve( state, idx_image, idx_chain ).image->spins.scalars();
}
SB_API_CATCH_RET( nullptr )

scalar * System_Get_Effective_Field( State * state, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    if( image->effective_field_stale )
    {
        ImageLock lock( *image );
        image->refresh_effective_field_mirror(); // the field of the last hook is still on the device
    }
    return image->effective_field.scalars();
}
SB_API_CATCH_RET( nullptr )

float System_Get_Rx( State * state, int idx_image, int idx_chain ) noexcept
try
{
    auto r = resolve( state, idx_image, idx_chain );
    return float( r.chain->Rx[idx_image] );
}
SB_API_CATCH_RET( 0 )

float System_Get_Energy( State * state, int idx_image, int idx_chain ) noexcept
try
{
    return float( resolve( state, idx_image, idx_chain ).image->E );
}
SB_API_CATCH_RET( 0 )

// System.cpp:137-177
int System_Get_Energy_Array_Names( State * state, char * names, int idx_image, int idx_chain ) noexcept
try
{
    auto image       = resolve( state, idx_image, idx_chain ).image;
    int n_char_array = -1;
    for( const auto & e : image->E_array )
        n_char_array += int( e.first.size() ) + 1;
    if( names == nullptr )
        return n_char_array;
    int idx = 0;
    for( std::size_t i = 0; i < image->E_array.size(); ++i )
    {
        for( char c : image->E_array[i].first )
            names[idx++] = c;
        if( i + 1 != image->E_array.size() )
            names[idx++] = '|';
    }
    return -1;
}
SB_API_CATCH_RET( -1 )

// System.cpp:179-211
int System_Get_Energy_Array( State * state, float * energies, bool divide_by_nspins, int idx_image, int idx_chain ) noexcept
try
{
    auto image     = resolve( state, idx_image, idx_chain ).image;
    double nd      = divide_by_nspins ? 1.0 / double( image->nos ) : 1.0;
    if( energies == nullptr )
        return int( image->E_array.size() );
    for( std::size_t i = 0; i < image->E_array.size(); ++i )
        energies[i] = float( nd * image->E_array[i].second );
    return -1;
}
SB_API_CATCH_RET( -1 )

void System_Print_Energy_Array( State * state, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    std::fprintf( stdout, "E_tot = %.10g meV/spin\n", image->E / image->nos );
    for( const auto & e : image->E_array )
        std::fprintf( stdout, "  %-18s %.10g meV/spin\n", e.first.c_str(), e.second / image->nos );
}
SB_API_CATCH_VOID

// System.cpp:255-278: per-term energies of the current spins (evaluated on the GPU)
void System_Update_Data( State * state, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    image->refresh_effective_field_mirror();
    image->UpdateEnergy();
}
SB_API_CATCH_VOID

// ---- Quantities ----
void Quantity_Get_Average_Spin( State * state, float s[3], int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    double m[3];
    image->sync_to_device();
    image->device().magnetization( m, false );
    for( int d = 0; d < 3; ++d )
        s[d] = float( m[d] );
}
SB_API_CATCH_VOID

void Quantity_Get_Magnetization( State * state, float m[3], int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    double mag[3];
    image->sync_to_device();
    image->device().magnetization( mag, true );
    image->M = Vec3{ mag[0], mag[1], mag[2] };
    for( int d = 0; d < 3; ++d )
        m[d] = float( mag[d] );
}
SB_API_CATCH_VOID

namespace
{
// The cut of the cell parallelogram and the orientation of its two triangles, as the reference's Delaunay triangulation of the
// (stretched) cell corners gives them (Vectormath.cpp:516-575). Observed on the reference for the degenerate square cell: a-b.
struct CellTriangulation
{
    int diag;       // 0: triangles (a, b, a+b), (a, b, 0); 1: (0, a, a+b), (0, a+b, b)
    double sign[2]; // z-orientation of the two triangles in that vertex order
};
CellTriangulation cell_triangulation( const Geometry & g )
{
    const Vec3 ta = g.bravais_vectors[0] * g.lattice_constant, tb = g.bravais_vectors[1] * g.lattice_constant;
    const Vec3 p0 = g.positions()[0];
    const double k = 0.1; // the reference stretches the corners away from the centre
    const Vec3 P0 = p0 - ( ta + tb ) * k, Pab = ta + tb + p0 + ( ta + tb ) * k, Pb = tb + p0 - ( ta - tb ) * k, Pa = ta + p0 + ( ta - tb ) * k;
    auto angle = []( const Vec3 & at, const Vec3 & u, const Vec3 & v )
    {
        const double ux = u.x - at.x, uy = u.y - at.y, vx = v.x - at.x, vy = v.y - at.y;
        return std::atan2( std::abs( ux * vy - uy * vx ), ux * vx + uy * vy );
    };
    // Delaunay: the diagonal a-b is kept when the angles opposite to it (at 0 and at a+b) sum to at most pi
    const double opposite = angle( P0, Pa, Pb ) + angle( Pab, Pa, Pb );
    CellTriangulation t;
    t.diag      = opposite <= constants::Pi + 1e-9 ? 0 : 1;
    auto orient = []( const Vec3 & q0, const Vec3 & q1, const Vec3 & q2 )
    {
        const double nz = ( q0.x - q1.x ) * ( q0.y - q2.y ) - ( q0.y - q1.y ) * ( q0.x - q2.x );
        return nz < 0 ? -1.0 : 1.0;
    };
    if( t.diag == 0 )
        t.sign[0] = orient( Pa, Pb, Pab ), t.sign[1] = orient( Pa, Pb, P0 );
    else
        t.sign[0] = orient( P0, Pa, Pab ), t.sign[1] = orient( P0, Pab, Pb );
    return t;
}
// Any number of basis atoms: Delaunay triangulation of the basis atoms of one cell and the corners a+b, b, a (ids NB,
// NB + 1, NB + 2), corners stretched by 10 % as in the reference (Vectormath.cpp:516-548). Few points: every triple whose
// circumcircle contains no other point is a triangle. One basis atom goes through cell_triangulation (the square cell is
// degenerate: four points on one circle) and is expressed in the same vertex ids.
struct CellTriangles
{
    int n = 0;
    int vertex[16][3];
    double sign[16];
};
CellTriangles cell_triangles( const Geometry & g )
{
    CellTriangles out;
    const int NB = g.n_cell_atoms;
    if( NB == 1 )
    {
        const CellTriangulation t = cell_triangulation( g );
        const int ids[2][2][3]    = { { { NB + 2, NB + 1, NB }, { NB + 2, NB + 1, 0 } }, { { 0, NB + 2, NB }, { 0, NB, NB + 1 } } };
        out.n                     = 2;
        for( int k = 0; k < 2; ++k )
        {
            for( int c = 0; c < 3; ++c )
                out.vertex[k][c] = ids[t.diag][k][c];
            out.sign[k] = t.sign[k];
        }
        return out;
    }
    const Vec3 ta = g.bravais_vectors[0] * g.lattice_constant, tb = g.bravais_vectors[1] * g.lattice_constant;
    const auto & positions = g.positions();
    const Vec3 p0          = positions[0];
    const double k         = 0.1;
    std::vector<std::array<double, 2>> pts;
    for( int i = 0; i < NB; ++i )
        pts.push_back( { positions[i].x, positions[i].y } );
    pts[0] = { p0.x - k * ( ta.x + tb.x ), p0.y - k * ( ta.y + tb.y ) };
    pts.push_back( { ta.x + tb.x + p0.x + k * ( ta.x + tb.x ), ta.y + tb.y + p0.y + k * ( ta.y + tb.y ) } ); // a + b
    pts.push_back( { tb.x + p0.x - k * ( ta.x - tb.x ), tb.y + p0.y - k * ( ta.y - tb.y ) } );               // b
    pts.push_back( { ta.x + p0.x + k * ( ta.x - tb.x ), ta.y + p0.y + k * ( ta.y - tb.y ) } );               // a
    const int n_pts = int( pts.size() );
    double area_triangles = 0;
    for( int i = 0; i < n_pts; ++i )
        for( int j = i + 1; j < n_pts; ++j )
            for( int l = j + 1; l < n_pts; ++l )
            {
                const auto &A = pts[i], &B = pts[j], &C = pts[l];
                const double d = 2 * ( A[0] * ( B[1] - C[1] ) + B[0] * ( C[1] - A[1] ) + C[0] * ( A[1] - B[1] ) );
                if( std::abs( d ) < 1e-12 )
                    continue; // collinear
                const double a2 = A[0] * A[0] + A[1] * A[1], b2 = B[0] * B[0] + B[1] * B[1], c2 = C[0] * C[0] + C[1] * C[1];
                const double ux = ( a2 * ( B[1] - C[1] ) + b2 * ( C[1] - A[1] ) + c2 * ( A[1] - B[1] ) ) / d;
                const double uy = ( a2 * ( C[0] - B[0] ) + b2 * ( A[0] - C[0] ) + c2 * ( B[0] - A[0] ) ) / d;
                const double r2 = ( A[0] - ux ) * ( A[0] - ux ) + ( A[1] - uy ) * ( A[1] - uy );
                bool empty      = true;
                for( int m = 0; m < n_pts && empty; ++m )
                    if( m != i && m != j && m != l )
                        empty = ( pts[m][0] - ux ) * ( pts[m][0] - ux ) + ( pts[m][1] - uy ) * ( pts[m][1] - uy ) > r2 * ( 1 + 1e-9 );
                if( !empty )
                    continue;
                if( out.n >= 16 )
                    throw std::runtime_error( "topological charge: too many triangles per cell" );
                out.vertex[out.n][0] = i, out.vertex[out.n][1] = j, out.vertex[out.n][2] = l;
                // orientation of the triangle in this vertex order: z of (p0 - p1) x (p0 - p2)
                const double nz = ( A[0] - B[0] ) * ( A[1] - C[1] ) - ( A[1] - B[1] ) * ( A[0] - C[0] );
                out.sign[out.n] = nz < 0 ? -1.0 : 1.0;
                area_triangles += 0.5 * std::abs( nz );
                ++out.n;
            }
    // a triangulation of the stretched parallelogram covers it exactly; anything else is a degenerate (co-circular) cell
    const auto &Q0 = pts[0], &Qab = pts[NB], &Qb = pts[NB + 1], &Qa = pts[NB + 2];
    const double area_cell = 0.5 * std::abs( ( Qa[0] - Q0[0] ) * ( Qab[1] - Q0[1] ) - ( Qa[1] - Q0[1] ) * ( Qab[0] - Q0[0] ) )
                             + 0.5 * std::abs( ( Qab[0] - Q0[0] ) * ( Qb[1] - Q0[1] ) - ( Qab[1] - Q0[1] ) * ( Qb[0] - Q0[0] ) );
    if( std::abs( area_triangles - area_cell ) > 1e-6 * area_cell )
        throw std::runtime_error( "topological charge: the basis cell has no unique Delaunay triangulation (or a basis atom lies outside the cell)" );
    return out;
}
// site index of vertex `id` of the cell (a, b), or -1 when its translation is not allowed (Vectormath.cpp:577-606)
int triangle_site( int id, int NB, int a, int b, int Na, int Nb, const std::array<int, 3> & bc )
{
    const bool a_ok = a + 1 < Na || bc[0], b_ok = b + 1 < Nb || bc[1];
    if( id < NB )
        return id + NB * ( a + Na * b );
    if( id == NB + 2 )
        return a_ok ? NB * ( ( a + 1 ) % Na + Na * b ) : -1;
    if( id == NB + 1 )
        return b_ok ? NB * ( a + Na * ( ( b + 1 ) % Nb ) ) : -1;
    return a_ok && b_ok ? NB * ( ( a + 1 ) % Na + Na * ( ( b + 1 ) % Nb ) ) : -1;
}
// all triangles that count, reference order (triangle of the cell outermost, then b, then a): 3 site indices each
std::vector<std::array<int, 4>> counted_triangles( const Geometry & g, const std::array<int, 3> & bc, const CellTriangles & t )
{
    std::vector<std::array<int, 4>> out; // site0, site1, site2, index into the density array [k][b][a]
    const int Na = g.n_cells[0], Nb = g.n_cells[1], NB = g.n_cell_atoms;
    for( int k = 0; k < t.n; ++k )
        for( int b = 0; b < Nb; ++b )
            for( int a = 0; a < Na; ++a )
            {
                const int s0 = triangle_site( t.vertex[k][0], NB, a, b, Na, Nb, bc ), s1 = triangle_site( t.vertex[k][1], NB, a, b, Na, Nb, bc ),
                          s2 = triangle_site( t.vertex[k][2], NB, a, b, Na, Nb, bc );
                if( s0 >= 0 && s1 >= 0 && s2 >= 0 )
                    out.push_back( { s0, s1, s2, ( k * Nb + b ) * Na + a } );
            }
    return out;
}
} // namespace

// Host-side probe (include/spirit_b200.h): the triangles the topological charge is summed over; no device needed
int SpiritB200_Topology_Triangles( State * state, int * triangle_indices, int idx_image ) noexcept
{
    int idx_chain = -1;
    try
    {
        auto image = resolve( state, idx_image, idx_chain ).image;
        if( image->geometry->dimensionality != 2 )
            return 0;
        const auto tris = counted_triangles( *image->geometry, image->hamiltonian->boundary_conditions, cell_triangles( *image->geometry ) );
        if( triangle_indices )
            for( std::size_t i = 0; i < tris.size(); ++i )
                for( int c = 0; c < 3; ++c )
                    triangle_indices[3 * i + c] = tris[i][c];
        return int( tris.size() );
    }
    SB_API_CATCH_RET( -1 )
}

// Quantities.cpp:62-87
float Quantity_Get_Topological_Charge( State * state, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    if( image->geometry->dimensionality != 2 )
        return 0;
    if( image->geometry->n_cell_atoms > 1 )
    {
        const CellTriangles t = cell_triangles( *image->geometry );
        image->sync_to_device();
        return float( image->device().topological_charge_table( t.n, t.vertex, t.sign, nullptr ) );
    }
    const CellTriangulation t = cell_triangulation( *image->geometry );
    image->sync_to_device();
    return float( image->device().topological_charge( t.diag, t.sign[0], t.sign[1], nullptr ) );
}
SB_API_CATCH_RET( 0 )

// Quantities.cpp:89-133: charge and site indices of every triangle that counts; returns their number. The triangles of a
// family are in the reference's order (b outer, a inner); which family comes first is the triangulation library's choice there.
int Quantity_Get_Topological_Charge_Density( State * state, float * charge_density, int * triangle_indices, int idx_image, int idx_chain ) noexcept
try
{
    auto image         = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    const Geometry & g = *image->geometry;
    if( g.dimensionality != 2 )
        return 0;
    if( g.n_cell_atoms > 1 )
    {
        const CellTriangles t = cell_triangles( g );
        const auto tris       = counted_triangles( g, image->hamiltonian->boundary_conditions, t );
        if( charge_density && triangle_indices )
        {
            std::vector<double> all( std::size_t( t.n ) * g.n_cells[0] * g.n_cells[1] );
            image->sync_to_device();
            image->device().topological_charge_table( t.n, t.vertex, t.sign, all.data() );
            for( std::size_t i = 0; i < tris.size(); ++i )
            {
                charge_density[i] = float( all[std::size_t( tris[i][3] )] );
                for( int c = 0; c < 3; ++c )
                    triangle_indices[3 * i + c] = tris[i][c];
            }
        }
        return int( tris.size() );
    }
    const CellTriangulation t = cell_triangulation( g );
    const int Na = g.n_cells[0], Nb = g.n_cells[1];
    const auto & bc = image->hamiltonian->boundary_conditions;
    std::vector<double> density;
    if( charge_density && triangle_indices )
    {
        density.resize( 2 * std::size_t( Na ) * Nb );
        image->sync_to_device();
        image->device().topological_charge( t.diag, t.sign[0], t.sign[1], density.data() );
    }
    int n = 0;
    for( int family = 0; family < 2; ++family )
        for( int b = 0; b < Nb; ++b )
            for( int a = 0; a < Na; ++a )
            {
                if( !( ( a + 1 < Na || bc[0] ) && ( b + 1 < Nb || bc[1] ) ) )
                    continue;
                if( charge_density && triangle_indices )
                {
                    const int i0 = a + Na * b, ia = ( a + 1 ) % Na + Na * b, ib = a + Na * ( ( b + 1 ) % Nb ), iab = ( a + 1 ) % Na + Na * ( ( b + 1 ) % Nb );
                    const int tri[2][2][3] = { { { ia, ib, iab }, { ia, ib, i0 } }, { { i0, ia, iab }, { i0, iab, ib } } };
                    charge_density[n] = float( density[std::size_t( family ) * Na * Nb + i0] );
                    for( int k = 0; k < 3; ++k )
                        triangle_indices[3 * n + k] = tri[t.diag][family][k];
                }
                ++n;
            }
    return n;
}
SB_API_CATCH_RET( 0 )

// ---- Geometry ----
namespace
{
// Vectormath::change_dimensions (Vectormath.hpp:571-612): keep the overlap of old and new lattice
void change_dimensions(
    HostField & field, int nb_old, const std::array<int, 3> & n_old, int nb_new, const std::array<int, 3> & n_new, Vec3 default_value )
{
    HostField out( std::size_t( nb_new ) * n_new[0] * n_new[1] * n_new[2] );
    for( std::size_t i = 0; i < out.size(); ++i )
        out[i] = default_value;
    for( int c = 0; c < n_new[2]; ++c )
        for( int b = 0; b < n_new[1]; ++b )
            for( int a = 0; a < n_new[0]; ++a )
                for( int ib = 0; ib < nb_new; ++ib )
                    if( ib < nb_old && a < n_old[0] && b < n_old[1] && c < n_old[2] )
                        out[ib + std::size_t( nb_new ) * ( a + std::size_t( n_new[0] ) * ( b + std::size_t( n_new[1] ) * c ) )]
                            = field[ib + std::size_t( nb_old ) * ( a + std::size_t( n_old[0] ) * ( b + std::size_t( n_old[1] ) * c ) )];
    field = out;
}

// Geometry.cpp:13-37
void system_set_geometry( Spin_System & system, const Geometry & new_geometry )
{
    const Geometry old = *system.geometry;
    system.nos         = new_geometry.nos;
    change_dimensions( system.spins, old.n_cell_atoms, old.n_cells, new_geometry.n_cell_atoms, new_geometry.n_cells, { 0, 0, 1 } );
    change_dimensions( system.effective_field, old.n_cell_atoms, old.n_cells, new_geometry.n_cell_atoms, new_geometry.n_cells, { 0, 0, 0 } );
    *system.geometry = new_geometry;
    system.drop_device();
    system.hamiltonian->Update_Interactions();
}

// Geometry.cpp:39-108
void state_set_geometry( State & state, const Geometry & new_geometry )
{
    Simulation_Stop_All( &state );
    state.chain->Lock();
    try
    {
        for( auto & system : state.chain->images )
            system_set_geometry( *system, new_geometry );
    }
    catch( ... )
    {
        handle_exception_api( "Geometry_Set" );
    }
    state.chain->Unlock();
    state.nos = state.active_image->nos;
    if( state.clipboard_image )
        system_set_geometry( *state.clipboard_image, new_geometry );
    if( state.clipboard_spins )
    {
        // the clipboard configuration lives in a plain vector: convert through a HostField
        state.clipboard_spins.reset();
    }
    state.method_image.assign( state.chain->noi, nullptr );
    state.method_chain.reset();
}

std::shared_ptr<Geometry> active_geometry( State * state )
{
    if( !state || !state->active_image )
        throw std::runtime_error( "The State pointer is invalid" );
    return state->active_image->geometry;
}
} // namespace

void Geometry_Set_Bravais_Lattice_Type( State * state, Bravais_Lattice_Type lattice_type ) noexcept
try
{
    auto old = active_geometry( state );
    std::vector<Vec3> bv;
    switch( lattice_type )
    {
        case Bravais_Lattice_SC: bv = Geometry::BravaisVectorsSC(); break;
        case Bravais_Lattice_Hex2D:
        case Bravais_Lattice_Hex2D_60: bv = Geometry::BravaisVectorsHex2D60(); break;
        case Bravais_Lattice_Hex2D_120: bv = Geometry::BravaisVectorsHex2D120(); break;
        case Bravais_Lattice_BCC: bv = Geometry::BravaisVectorsBCC(); break;
        case Bravais_Lattice_FCC: bv = Geometry::BravaisVectorsFCC(); break;
        default:
            Log( Log_Level::Warning, Log_Sender::API, "Geometry_Set_Bravais_Lattice_Type: cannot set this lattice type" );
            return;
    }
    Geometry g( bv, old->n_cells, old->cell_atoms, old->cell_mu_s, old->lattice_constant );
    g.cell_atom_types = old->cell_atom_types;
    state_set_geometry( *state, g );
}
catch( ... )
{
    handle_exception_api( __func__ );
}

void Geometry_Set_N_Cells( State * state, int n_cells[3] ) noexcept
try
{
    auto old = active_geometry( state );
    Geometry g( old->bravais_vectors, { n_cells[0], n_cells[1], n_cells[2] }, old->cell_atoms, old->cell_mu_s, old->lattice_constant );
    g.cell_atom_types = old->cell_atom_types;
    state_set_geometry( *state, g );
}
catch( ... )
{
    handle_exception_api( __func__ );
}

void Geometry_Set_Cell_Atoms( State * state, int n_atoms, float ** atoms ) noexcept
try
{
    auto old = active_geometry( state );
    if( n_atoms < 1 )
    {
        Log( Log_Level::Error, Log_Sender::API, "Cannot set number of atoms to less than one." );
        return;
    }
    std::vector<Vec3> cell_atoms( n_atoms );
    for( int i = 0; i < n_atoms; ++i )
        cell_atoms[i] = Vec3{ atoms[i][0], atoms[i][1], atoms[i][2] };
    // Geometry.cpp:246-262: keep the composition of the atoms that still exist, new ones copy atom 0
    std::vector<double> mu_s( n_atoms );
    for( int i = 0; i < n_atoms; ++i )
        mu_s[i] = i < old->n_cell_atoms ? old->cell_mu_s[i] : old->cell_mu_s[0];
    Geometry g( old->bravais_vectors, old->n_cells, cell_atoms, mu_s, old->lattice_constant );
    state_set_geometry( *state, g );
}
catch( ... )
{
    handle_exception_api( __func__ );
}

void Geometry_Set_mu_s( State * state, float mu_s, int idx_image, int idx_chain ) noexcept
try
{
    auto old = active_geometry( state );
    std::vector<double> new_mu_s( old->n_cell_atoms, double( mu_s ) );
    Geometry g( old->bravais_vectors, old->n_cells, old->cell_atoms, new_mu_s, old->lattice_constant );
    g.cell_atom_types = old->cell_atom_types;
    state_set_geometry( *state, g );
}
SB_API_CATCH_VOID

void Geometry_Set_Cell_Atom_Types( State * state, int n_atoms, int * atom_types ) noexcept
try
{
    auto old = active_geometry( state );
    Geometry g = *old;
    for( int i = 0; i < n_atoms && i < g.n_cell_atoms; ++i )
        g.cell_atom_types[i] = atom_types[i];
    state_set_geometry( *state, g );
}
catch( ... )
{
    handle_exception_api( __func__ );
}

void Geometry_Set_Bravais_Vectors( State * state, float ta[3], float tb[3], float tc[3] ) noexcept
try
{
    auto old = active_geometry( state );
    std::vector<Vec3> bv{ Vec3{ ta[0], ta[1], ta[2] }, Vec3{ tb[0], tb[1], tb[2] }, Vec3{ tc[0], tc[1], tc[2] } };
    Geometry g( bv, old->n_cells, old->cell_atoms, old->cell_mu_s, old->lattice_constant );
    g.cell_atom_types = old->cell_atom_types;
    state_set_geometry( *state, g );
}
catch( ... )
{
    handle_exception_api( __func__ );
}

void Geometry_Set_Lattice_Constant( State * state, float lattice_constant ) noexcept
try
{
    auto old = active_geometry( state );
    Geometry g( old->bravais_vectors, old->n_cells, old->cell_atoms, old->cell_mu_s, double( lattice_constant ) );
    g.cell_atom_types = old->cell_atom_types;
    state_set_geometry( *state, g );
}
catch( ... )
{
    handle_exception_api( __func__ );
}

int Geometry_Get_NOS( State * state ) noexcept
try
{
    return active_geometry( state )->nos;
}
catch( ... )
{
    handle_exception_api( __func__ );
    return 0;
}

scalar * Geometry_Get_Positions( State * state, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    return const_cast<scalar *>( reinterpret_cast<const scalar *>( image->geometry->positions().data() ) );
}
SB_API_CATCH_RET( nullptr )

int * Geometry_Get_Atom_Types( State * state, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    return const_cast<int *>( image->geometry->atom_types().data() );
}
SB_API_CATCH_RET( nullptr )

void Geometry_Get_Bounds( State * state, float min[3], float max[3], int idx_image, int idx_chain ) noexcept
try
{
    auto g = resolve( state, idx_image, idx_chain ).image->geometry;
    for( int d = 0; d < 3; ++d )
    {
        min[d] = float( g->bounds_min[d] );
        max[d] = float( g->bounds_max[d] );
    }
}
SB_API_CATCH_VOID

void Geometry_Get_Center( State * state, float center[3], int idx_image, int idx_chain ) noexcept
try
{
    auto g = resolve( state, idx_image, idx_chain ).image->geometry;
    for( int d = 0; d < 3; ++d )
        center[d] = float( g->center[d] );
}
SB_API_CATCH_VOID

void Geometry_Get_Cell_Bounds( State * state, float min[3], float max[3], int idx_image, int idx_chain ) noexcept
try
{
    auto g = resolve( state, idx_image, idx_chain ).image->geometry;
    for( int d = 0; d < 3; ++d )
    {
        min[d] = float( g->cell_bounds_min[d] );
        max[d] = float( g->cell_bounds_max[d] );
    }
}
SB_API_CATCH_VOID

Bravais_Lattice_Type Geometry_Get_Bravais_Lattice_Type( State * state, int idx_image, int idx_chain ) noexcept
try
{
    return Bravais_Lattice_Type( int( resolve( state, idx_image, idx_chain ).image->geometry->classifier ) );
}
SB_API_CATCH_RET( Bravais_Lattice_Irregular )

void Geometry_Get_Bravais_Vectors( State * state, float a[3], float b[3], float c[3], int idx_image, int idx_chain ) noexcept
try
{
    auto g = resolve( state, idx_image, idx_chain ).image->geometry;
    for( int d = 0; d < 3; ++d )
    {
        a[d] = float( g->bravais_vectors[0][d] );
        b[d] = float( g->bravais_vectors[1][d] );
        c[d] = float( g->bravais_vectors[2][d] );
    }
}
SB_API_CATCH_VOID

int Geometry_Get_Dimensionality( State * state, int idx_image, int idx_chain ) noexcept
try
{
    return resolve( state, idx_image, idx_chain ).image->geometry->dimensionality;
}
SB_API_CATCH_RET( 0 )

void Geometry_Get_mu_s( State * state, float * mu_s, int idx_image, int idx_chain ) noexcept
try
{
    auto g = resolve( state, idx_image, idx_chain ).image->geometry;
    for( int i = 0; i < g->n_cell_atoms; ++i )
        mu_s[i] = float( g->cell_mu_s[i] );
}
SB_API_CATCH_VOID

void Geometry_Get_N_Cells( State * state, int n_cells[3], int idx_image, int idx_chain ) noexcept
try
{
    auto g = resolve( state, idx_image, idx_chain ).image->geometry;
    for( int d = 0; d < 3; ++d )
        n_cells[d] = g->n_cells[d];
}
SB_API_CATCH_VOID

int Geometry_Get_N_Cell_Atoms( State * state, int idx_image, int idx_chain ) noexcept
try
{
    return resolve( state, idx_image, idx_chain ).image->geometry->n_cell_atoms;
}
SB_API_CATCH_RET( 0 )

int Geometry_Get_Cell_Atoms( State * state, scalar ** atoms, int idx_image, int idx_chain ) noexcept
try
{
    auto g = resolve( state, idx_image, idx_chain ).image->geometry;
    if( atoms != nullptr )
        *atoms = reinterpret_cast<scalar *>( g->cell_atoms.data() );
    return g->n_cell_atoms;
}
SB_API_CATCH_RET( 0 )
