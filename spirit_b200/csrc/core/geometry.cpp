#include "geometry.hpp"

#include <algorithm>
#include <limits>
#include <stdexcept>

namespace sb
{

Geometry::Geometry(
    const std::vector<Vec3> & bravais_vectors_, const std::array<int, 3> & n_cells_,
    const std::vector<Vec3> & cell_atoms_, const std::vector<double> & cell_mu_s_, double lattice_constant_ )
        : bravais_vectors( bravais_vectors_ ),
          lattice_constant( lattice_constant_ ),
          n_cells( n_cells_ ),
          n_cell_atoms( int( cell_atoms_.size() ) ),
          cell_atoms( cell_atoms_ ),
          cell_mu_s( cell_mu_s_ ),
          cell_atom_types( cell_atoms_.size(), 0 )
{
    if( n_cell_atoms < 1 || n_cells[0] < 1 || n_cells[1] < 1 || n_cells[2] < 1 )
        throw std::runtime_error( "Geometry: need at least one basis atom and one cell per direction" );
    std::int64_t total = std::int64_t( n_cell_atoms ) * n_cells[0] * n_cells[1] * n_cells[2];
    if( total > std::int64_t( std::numeric_limits<int>::max() ) )
        throw std::runtime_error( "Geometry: number of spins exceeds the C API's int range" );
    cell_mu_s.resize( n_cell_atoms, cell_mu_s.empty() ? 1.0 : cell_mu_s[0] );

    nos           = int( total );
    nos_nonvacant = nos;
    n_cells_total = n_cells[0] * n_cells[1] * n_cells[2];

    // Two basis atoms must not coincide modulo lattice translations (Geometry.cpp:94-143)
    const double epsilon = 1e-6;
    int max_a = std::min( 10, n_cells[0] ), max_b = std::min( 10, n_cells[1] ), max_c = std::min( 10, n_cells[2] );
    for( int i = 0; i < n_cell_atoms; ++i )
        for( int j = 0; j < n_cell_atoms; ++j )
            for( int da = -max_a; da <= max_a; ++da )
                for( int db = -max_b; db <= max_b; ++db )
                    for( int dc = -max_c; dc <= max_c; ++dc )
                    {
                        Vec3 diff = cell_atoms[i] - ( cell_atoms[j] + Vec3{ double( da ), double( db ), double( dc ) } );
                        bool same = std::abs( diff.x ) < epsilon && std::abs( diff.y ) < epsilon
                                    && std::abs( diff.z ) < epsilon;
                        if( same && ( i != j || da != 0 || db != 0 || dc != 0 ) )
                            throw std::runtime_error(
                                "Geometry: two basis atoms occupy the same position modulo a lattice translation" );
                    }

    calculateBounds();
    calculateUnitCellBounds();
    calculateDimensionality();
    center = ( bounds_min + bounds_max ) * 0.5;
    calculateGeometryType();
}

Vec3 Geometry::position_of( std::int64_t a, std::int64_t b, std::int64_t c, int iatom ) const
{
    return lattice_constant
           * ( ( double( a ) + cell_atoms[iatom][0] ) * bravais_vectors[0]
               + ( double( b ) + cell_atoms[iatom][1] ) * bravais_vectors[1]
               + ( double( c ) + cell_atoms[iatom][2] ) * bravais_vectors[2] );
}

const vectorfield & Geometry::positions() const
{
    if( _positions.size() != std::size_t( nos ) )
    {
        _positions.resize( nos );
        std::int64_t Na = n_cells[0], Nb = n_cells[1], Nc = n_cells[2], N = n_cell_atoms;
#pragma omp parallel for collapse( 2 )
        for( std::int64_t c = 0; c < Nc; ++c )
            for( std::int64_t b = 0; b < Nb; ++b )
                for( std::int64_t a = 0; a < Na; ++a )
                    for( int iatom = 0; iatom < N; ++iatom )
                        _positions[iatom + N * ( a + Na * ( b + Nb * c ) )] = position_of( a, b, c, iatom );
    }
    return _positions;
}

const scalarfield & Geometry::mu_s() const
{
    if( _mu_s.size() != std::size_t( nos ) )
    {
        _mu_s.resize( nos );
        for( std::int64_t i = 0; i < nos; ++i )
            _mu_s[i] = ( !site_flags.empty() && ( site_flags[i] & SITE_NO_MU_S ) ) ? 0.0 : cell_mu_s[i % n_cell_atoms];
    }
    return _mu_s;
}

const intfield & Geometry::atom_types() const
{
    if( _atom_types.size() != std::size_t( nos ) )
    {
        _atom_types.resize( nos );
        for( std::int64_t i = 0; i < nos; ++i )
            _atom_types[i] = cell_atom_types[i % n_cell_atoms];
    }
    return _atom_types;
}

// ---- pinning and defects (Geometry.cpp:50-83, 486-566, 807-832) ----
int Geometry::site_index( const LatticeSite & site ) const
{
    return site.i + n_cell_atoms * ( site.translations[0] + n_cells[0] * ( site.translations[1] + n_cells[1] * site.translations[2] ) );
}

void Geometry::need_site_flags()
{
    if( site_flags.empty() )
    {
        site_flags.assign( nos, 0 );
        mask_pinned_cells.assign( nos, Vec3{ 0, 0, 0 } );
    }
}

void Geometry::set_pinning_and_defects( const Pinning & pinning_, const Defects & defects_ )
{
    pinning = pinning_;
    defects = defects_;
    pinning.pinned_cell.resize( n_cell_atoms, Vec3{ 0, 0, 1 } );
    site_flags.clear();
    mask_pinned_cells.clear();
    _atom_types.clear();
    _mu_s.clear();
    nos_nonvacant = nos;
    ++site_revision;
    const bool boundary = pinning.na_left > 0 || pinning.na_right > 0 || pinning.nb_left > 0 || pinning.nb_right > 0 || pinning.nc_left > 0
                          || pinning.nc_right > 0;
    if( !boundary && pinning.sites.empty() && defects.sites.empty() )
        return;
    need_site_flags();
    const std::int64_t Na = n_cells[0], Nb = n_cells[1], Nc = n_cells[2], N = n_cell_atoms;
    if( boundary )
        for( std::int64_t c = 0; c < Nc; ++c )
            for( std::int64_t b = 0; b < Nb; ++b )
                for( std::int64_t a = 0; a < Na; ++a )
                    if( a < pinning.na_left || a >= Na - pinning.na_right || b < pinning.nb_left || b >= Nb - pinning.nb_right
                        || c < pinning.nc_left || c >= Nc - pinning.nc_right )
                        for( int iatom = 0; iatom < N; ++iatom )
                        {
                            const std::int64_t ispin = iatom + N * ( a + Na * ( b + Nb * c ) );
                            site_flags[ispin] |= SITE_PINNED;
                            mask_pinned_cells[ispin] = pinning.pinned_cell[iatom];
                        }
    for( std::size_t k = 0; k < pinning.sites.size(); ++k )
    {
        const int ispin = site_index( pinning.sites[k] );
        if( ispin < 0 || ispin >= nos )
            throw std::runtime_error( "Geometry: pinned site outside of the lattice" );
        site_flags[ispin] |= SITE_PINNED;
        mask_pinned_cells[ispin] = pinning.spins[k];
    }
    atom_types(); // per-site types: the basis cell's, then the defects'
    for( std::size_t k = 0; k < defects.sites.size(); ++k )
    {
        const int ispin = site_index( defects.sites[k] );
        if( ispin < 0 || ispin >= nos )
            throw std::runtime_error( "Geometry: defect site outside of the lattice" );
        _atom_types[ispin] = defects.types[k];
        site_flags[ispin] |= SITE_NO_MU_S; // Geometry.cpp:80-81: mu_s = 0 at every defect site
        if( defects.types[k] < 0 )
            site_flags[ispin] |= SITE_VACANT;
        else
            site_flags[ispin] &= static_cast<unsigned char>( ~SITE_VACANT );
    }
}

void Geometry::set_pinned( int ispin, bool pinned, const Vec3 & orientation )
{
    need_site_flags();
    if( pinned )
        site_flags[ispin] |= SITE_PINNED;
    else
        site_flags[ispin] &= static_cast<unsigned char>( ~SITE_PINNED );
    mask_pinned_cells[ispin] = orientation;
    ++site_revision;
}

void Geometry::set_atom_type( int ispin, int type )
{
    need_site_flags();
    atom_types();
    _atom_types[ispin] = type;
    if( type < 0 )
    {
        if( !( site_flags[ispin] & SITE_VACANT ) )
            --nos_nonvacant;
        site_flags[ispin] |= SITE_VACANT | SITE_NO_MU_S; // Configurations.cpp:577-579: mu_s = 0 for vacancies
    }
    else
    {
        if( site_flags[ispin] & SITE_VACANT )
            ++nos_nonvacant;
        site_flags[ispin] &= static_cast<unsigned char>( ~SITE_VACANT ); // (mu_s stays as it is, as in the reference)
    }
    _mu_s.clear();
    ++site_revision;
}

void Geometry::set_vacancy_read_from_file( int ispin )
{
    need_site_flags();
    atom_types();
    if( _atom_types[ispin] == -1 && ( site_flags[ispin] & SITE_VACANT ) )
        return;
    _atom_types[ispin] = -1;
    site_flags[ispin] |= SITE_VACANT; // (mu_s and nos_nonvacant stay as they are: the reference sets the atom type only)
    ++site_revision;
}

void Geometry::apply_pinning( Vec3 * spins ) const
{
    if( site_flags.empty() )
        return;
    for( std::int64_t i = 0; i < nos; ++i )
        if( site_flags[i] & SITE_PINNED )
            spins[i] = mask_pinned_cells[i];
}

bool Geometry::mu_s_homogeneous() const
{
    for( double m : cell_mu_s )
        if( m != cell_mu_s[0] )
            return false;
    return true;
}

// Geometry.cpp:730-745 takes min/max over all positions starting from zero. The positions are affine in
// (a,b,c), so the extrema are attained at corner cells; evaluating only those gives identical values.
void Geometry::calculateBounds()
{
    bounds_min = { 0, 0, 0 };
    bounds_max = { 0, 0, 0 };
    for( int ca = 0; ca < 2; ++ca )
        for( int cb = 0; cb < 2; ++cb )
            for( int cc = 0; cc < 2; ++cc )
                for( int iatom = 0; iatom < n_cell_atoms; ++iatom )
                {
                    Vec3 p = position_of(
                        ca * ( n_cells[0] - 1 ), cb * ( n_cells[1] - 1 ), cc * ( n_cells[2] - 1 ), iatom );
                    for( int dim = 0; dim < 3; ++dim )
                    {
                        bounds_min[dim] = std::min( bounds_min[dim], p[dim] );
                        bounds_max[dim] = std::max( bounds_max[dim], p[dim] );
                    }
                }
}

// Geometry.cpp:747-773
void Geometry::calculateUnitCellBounds()
{
    cell_bounds_min = { 0, 0, 0 };
    cell_bounds_max = { 0, 0, 0 };
    for( const auto & bv : bravais_vectors )
        for( int iatom = 0; iatom < n_cell_atoms; ++iatom )
        {
            Vec3 p  = position_of( 0, 0, 0, iatom );
            Vec3 n1 = p + lattice_constant * bv;
            Vec3 n2 = p - lattice_constant * bv;
            for( int dim = 0; dim < 3; ++dim )
            {
                cell_bounds_min[dim] = std::min( { cell_bounds_min[dim], n1[dim], n2[dim] } );
                cell_bounds_max[dim] = std::max( { cell_bounds_max[dim], n1[dim], n2[dim] } );
            }
        }
    cell_bounds_min = cell_bounds_min * 0.5;
    cell_bounds_max = cell_bounds_max * 0.5;
}

// Geometry.cpp:566-728
void Geometry::calculateDimensionality()
{
    Vec3 test_vec_basis, test_vec_translations;
    const double epsilon = std::numeric_limits<double>::epsilon();

    if( n_cell_atoms == 1 )
        dimensionality_basis = 0;
    else if( n_cell_atoms == 2 )
    {
        dimensionality_basis = 1;
        test_vec_basis       = position_of( 0, 0, 0, 0 ) - position_of( 0, 0, 0, 1 );
    }
    else
    {
        Vec3 v0 = position_of( 0, 0, 0, 0 );
        std::vector<Vec3> b_vectors( n_cell_atoms - 1 );
        for( int i = 1; i < n_cell_atoms; ++i )
            b_vectors[i - 1] = ( position_of( 0, 0, 0, i ) - v0 ).normalized();
        test_vec_basis         = b_vectors[0];
        std::size_t n_parallel = 0;
        for( std::size_t i = 1; i < b_vectors.size(); ++i )
        {
            if( 1 - std::abs( b_vectors[i].dot( test_vec_basis ) ) < epsilon )
                ++n_parallel;
            else
                break;
        }
        if( n_parallel == b_vectors.size() - 1 )
            dimensionality_basis = 1;
        else
        {
            test_vec_basis         = b_vectors[0].cross( b_vectors[n_parallel + 1] );
            std::size_t n_in_plane = 0;
            for( std::size_t i = 2; i < b_vectors.size(); ++i )
                if( std::abs( b_vectors[i].dot( test_vec_basis ) ) < epsilon )
                    ++n_in_plane;
            if( std::int64_t( n_in_plane ) == std::int64_t( b_vectors.size() ) - 2 )
                dimensionality_basis = 2;
            else
            {
                dimensionality_basis = 3;
                dimensionality       = 3;
                return;
            }
        }
    }

    double t01 = std::abs( bravais_vectors[0].normalized().dot( bravais_vectors[1].normalized() ) ) - 1.0;
    double t02 = std::abs( bravais_vectors[0].normalized().dot( bravais_vectors[2].normalized() ) ) - 1.0;
    double t12 = std::abs( bravais_vectors[1].normalized().dot( bravais_vectors[2].normalized() ) ) - 1.0;

    int dims_translations   = 0;
    int n_independent_pairs = 0;
    if( ( t01 < epsilon ) && ( n_cells[0] > 1 ) && ( n_cells[1] > 1 ) )
        ++n_independent_pairs;
    if( ( t02 < epsilon ) && ( n_cells[0] > 1 ) && ( n_cells[2] > 1 ) )
        ++n_independent_pairs;
    if( ( t12 < epsilon ) && ( n_cells[1] > 1 ) && ( n_cells[2] > 1 ) )
        ++n_independent_pairs;

    if( ( n_cells[0] == 1 ) && ( n_cells[1] == 1 ) && ( n_cells[2] == 1 ) )
        dims_translations = 0;
    else if( n_independent_pairs == 0 )
    {
        dims_translations = 1;
        for( int i = 0; i < 3; ++i )
            if( n_cells[i] > 1 )
                test_vec_translations = bravais_vectors[i];
    }
    else if( n_independent_pairs < 3 )
    {
        dims_translations = 2;
        int n             = 0;
        std::vector<Vec3> plane( 2 );
        for( int i = 0; i < 3; ++i )
            if( n_cells[i] > 1 && n < 2 )
                plane[n++] = bravais_vectors[i];
        test_vec_translations = plane[0].cross( plane[1] );
    }
    else
    {
        dimensionality = 3;
        return;
    }

    test_vec_basis.normalize();
    test_vec_translations.normalize();
    if( dimensionality_basis == 0 )
        dimensionality = dims_translations;
    else if( dims_translations == 0 )
        dimensionality = dimensionality_basis;
    else if( dimensionality_basis == dims_translations )
    {
        if( std::abs( test_vec_basis.dot( test_vec_translations ) ) - 1 < epsilon )
            dimensionality = dimensionality_basis;
        else
            dimensionality = dimensionality_basis + 1;
    }
    else if(
        ( dimensionality_basis == 1 && dims_translations == 2 )
        || ( dimensionality_basis == 2 && dims_translations == 1 ) )
    {
        if( std::abs( test_vec_basis.dot( test_vec_translations ) ) < epsilon )
            dimensionality = 2;
        else
            dimensionality = 3;
    }
}

// Geometry.cpp:775-800
void Geometry::calculateGeometryType()
{
    const double epsilon = std::numeric_limits<double>::epsilon();
    classifier           = BravaisLatticeType::Irregular;
    if( cell_atoms.size() == 1 )
    {
        if( ( std::abs( bravais_vectors[0].normalized().dot( bravais_vectors[1].normalized() ) ) < epsilon )
            && ( std::abs( bravais_vectors[0].normalized().dot( bravais_vectors[2].normalized() ) ) < epsilon ) )
        {
            if( ( bravais_vectors[0].norm() == bravais_vectors[1].norm() )
                && ( bravais_vectors[1].norm() == bravais_vectors[2].norm() ) )
                classifier = BravaisLatticeType::SC;
            else
                classifier = BravaisLatticeType::Rectilinear;
        }
    }
}

std::vector<Vec3> Geometry::BravaisVectorsSC()
{
    return { { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } };
}
std::vector<Vec3> Geometry::BravaisVectorsFCC()
{
    return { { 0.5, 0.0, 0.5 }, { 0.5, 0.5, 0.0 }, { 0.0, 0.5, 0.5 } };
}
std::vector<Vec3> Geometry::BravaisVectorsBCC()
{
    return { { 0.5, 0.5, -0.5 }, { -0.5, 0.5, -0.5 }, { 0.5, -0.5, -0.5 } };
}
std::vector<Vec3> Geometry::BravaisVectorsHex2D60()
{
    return { { 0.5 * std::sqrt( 3.0 ), -0.5, 0 }, { 0.5 * std::sqrt( 3.0 ), 0.5, 0 }, { 0, 0, 1 } };
}
std::vector<Vec3> Geometry::BravaisVectorsHex2D120()
{
    return { { 0.5, -0.5 * std::sqrt( 3.0 ), 0 }, { 0.5, 0.5 * std::sqrt( 3.0 ), 0 }, { 0, 0, 1 } };
}

} // namespace sb
