#include "neighbours.hpp"

#include <algorithm>
#include <limits>

namespace sb
{
namespace neighbours
{

namespace
{
constexpr double min_shell_width = 1e-3;

struct SearchBox
{
    Vec3 ta, tb, tc;
    int i_max, j_max, k_max;
};

// Search range of translations: n_shells + 2, limited by the lattice size (Neighbours.cpp:25-38,95-113)
SearchBox search_box( const Geometry & g, std::size_t n_shells )
{
    SearchBox s;
    s.ta                   = g.lattice_constant * g.bravais_vectors[0];
    s.tb                   = g.lattice_constant * g.bravais_vectors[1];
    s.tc                   = g.lattice_constant * g.bravais_vectors[2];
    int max_n_translations = int( n_shells ) + 2;
    s.i_max                = std::min( max_n_translations, g.n_cells[0] - 1 );
    s.j_max                = std::min( max_n_translations, g.n_cells[1] - 1 );
    s.k_max                = std::min( max_n_translations, g.n_cells[2] - 1 );
    return s;
}
} // namespace

// Neighbours.cpp:15-81. Note that, like the reference, the basis positions enter in units of the
// Bravais vectors (geometry.cell_atoms), not as absolute positions.
std::vector<double> get_shell_radii( const Geometry & g, std::size_t n_shells )
{
    std::vector<double> shell_radii( n_shells );
    SearchBox s = search_box( g, n_shells );
    int i_max = s.i_max, j_max = s.j_max, k_max = s.k_max;
    if( s.ta.norm() == 0.0 )
        i_max = 0;
    if( s.tb.norm() == 0.0 )
        j_max = 0;
    if( s.tc.norm() == 0.0 )
        k_max = 0;

    double outermost_radius = 0, previous_radius = 0;
    for( auto & shell_radius : shell_radii )
    {
        previous_radius  = outermost_radius;
        outermost_radius = std::numeric_limits<double>::max();
        for( int atom_one = 0; atom_one < g.n_cell_atoms; ++atom_one )
        {
            Vec3 pos_one = g.cell_atoms[atom_one];
            // By symmetry only half the space needs to be searched
            for( int i = i_max; i >= 0; --i )
                for( int j = j_max; j >= -j_max; --j )
                    for( int k = k_max; k >= -k_max; --k )
                        for( int atom_two = 0; atom_two < g.n_cell_atoms; ++atom_two )
                        {
                            if( atom_one == atom_two && i == 0 && j == 0 && k == 0 )
                                continue;
                            Vec3 pos_two = g.cell_atoms[atom_two] + double( i ) * s.ta + double( j ) * s.tb
                                           + double( k ) * s.tc;
                            double pos_delta = ( pos_one - pos_two ).norm();
                            if( pos_delta - previous_radius > min_shell_width && pos_delta < outermost_radius )
                            {
                                outermost_radius = pos_delta;
                                shell_radius     = pos_delta;
                            }
                        }
        }
    }
    return shell_radii;
}

// Neighbours.cpp:83-156
void get_neighbours_in_shells(
    const Geometry & g, std::size_t n_shells, pairfield & neighbours, intfield & shells, bool use_redundant_neighbours )
{
    auto shell_radii = get_shell_radii( g, n_shells );
    SearchBox s      = search_box( g, n_shells );
    int i_max = s.i_max, j_max = s.j_max, k_max = s.k_max;
    // The lower bounds are taken before the zero-vector abort conditions, as in the reference
    int i_min = -i_max, j_min = -j_max, k_min = -k_max;
    if( s.ta.norm() == 0.0 )
        i_max = 0;
    if( s.tb.norm() == 0.0 )
        j_max = 0;
    if( s.tc.norm() == 0.0 )
        k_max = 0;

    int second_atom_min = 0;
    for( int atom_one = 0; atom_one < g.n_cell_atoms; ++atom_one )
    {
        if( !use_redundant_neighbours )
            second_atom_min = atom_one;
        Vec3 pos_one = g.cell_atoms[atom_one];
        for( std::size_t ishell = 0; ishell < n_shells; ++ishell )
        {
            double radius = shell_radii[ishell];
            for( int i = i_max; i >= i_min; --i )
                for( int j = j_max; j >= j_min; --j )
                    for( int k = k_max; k >= k_min; --k )
                        for( int atom_two = second_atom_min; atom_two < g.n_cell_atoms; ++atom_two )
                        {
                            if( ( atom_two > atom_one )
                                || ( i > 0 || ( i == 0 && j > 0 ) || ( i == 0 && j == 0 && k > 0 ) )
                                || use_redundant_neighbours )
                            {
                                Vec3 pos_two = g.cell_atoms[atom_two] + double( i ) * s.ta + double( j ) * s.tb
                                               + double( k ) * s.tc;
                                double pos_delta = ( pos_one - pos_two ).norm();
                                if( std::abs( pos_delta - radius ) < min_shell_width )
                                {
                                    neighbours.push_back( Pair{ atom_one, atom_two, { i, j, k } } );
                                    shells.push_back( int( ishell ) );
                                }
                            }
                        }
        }
    }
}

// Neighbours.cpp:249-286. Chirality: +-1 Bloch, +-2 Neel, else zero vector.
Vec3 dmi_normal_from_pair( const Geometry & g, const Pair & pair, int chirality )
{
    Vec3 ta = g.lattice_constant * g.bravais_vectors[0];
    Vec3 tb = g.lattice_constant * g.bravais_vectors[1];
    Vec3 tc = g.lattice_constant * g.bravais_vectors[2];

    Vec3 ipos = g.position_of( 0, 0, 0, pair.i );
    Vec3 jpos = g.position_of( 0, 0, 0, pair.j ) + double( pair.translations[0] ) * ta
                + double( pair.translations[1] ) * tb + double( pair.translations[2] ) * tc;

    if( chirality == 1 )
        return ( jpos - ipos ).normalized();
    else if( chirality == -1 )
        return ( ipos - jpos ).normalized();
    else if( chirality == 2 )
        return ( jpos - ipos ).normalized().cross( Vec3{ 0, 0, 1 } );
    else if( chirality == -2 )
        return Vec3{ 0, 0, 1 }.cross( ( jpos - ipos ).normalized() );
    return Vec3{ 0, 0, 0 };
}

} // namespace neighbours
} // namespace sb
