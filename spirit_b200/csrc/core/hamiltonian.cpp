#include "hamiltonian.hpp"
#include "neighbours.hpp"

#include <cmath>

namespace sb
{

Hamiltonian::Hamiltonian( std::shared_ptr<Geometry> geometry_ ) : geometry( std::move( geometry_ ) ) {}

void Hamiltonian::Update_Interactions()
{
    // Every spin gathers from all of its neighbours (race-free), so every pair is listed in both
    // directions -- the reference does the same whenever it runs in parallel
    // (Hamiltonian_Heisenberg.cpp:103-109).
    const bool use_redundant_neighbours = true;

    // Exchange (Hamiltonian_Heisenberg.cpp:111-141)
    exchange_pairs.clear();
    exchange_magnitudes.clear();
    if( !exchange_shell_magnitudes.empty() )
    {
        intfield exchange_shells;
        neighbours::get_neighbours_in_shells(
            *geometry, exchange_shell_magnitudes.size(), exchange_pairs, exchange_shells, use_redundant_neighbours );
        for( std::size_t ipair = 0; ipair < exchange_pairs.size(); ++ipair )
            exchange_magnitudes.push_back( exchange_shell_magnitudes[exchange_shells[ipair]] );
    }
    else
    {
        exchange_pairs      = exchange_pairs_in;
        exchange_magnitudes = exchange_magnitudes_in;
        for( std::size_t i = 0; i < exchange_pairs_in.size(); ++i )
        {
            const auto & p = exchange_pairs_in[i];
            const auto & t = p.translations;
            exchange_pairs.push_back( Pair{ p.j, p.i, { -t[0], -t[1], -t[2] } } );
            exchange_magnitudes.push_back( exchange_magnitudes_in[i] );
        }
    }

    // DMI (Hamiltonian_Heisenberg.cpp:143-177)
    dmi_pairs.clear();
    dmi_magnitudes.clear();
    dmi_normals.clear();
    if( !dmi_shell_magnitudes.empty() )
    {
        intfield dmi_shells;
        neighbours::get_neighbours_in_shells(
            *geometry, dmi_shell_magnitudes.size(), dmi_pairs, dmi_shells, use_redundant_neighbours );
        for( std::size_t ineigh = 0; ineigh < dmi_pairs.size(); ++ineigh )
        {
            dmi_normals.push_back( neighbours::dmi_normal_from_pair( *geometry, dmi_pairs[ineigh], dmi_shell_chirality ) );
            dmi_magnitudes.push_back( dmi_shell_magnitudes[dmi_shells[ineigh]] );
        }
    }
    else
    {
        dmi_pairs      = dmi_pairs_in;
        dmi_magnitudes = dmi_magnitudes_in;
        dmi_normals    = dmi_normals_in;
        for( std::size_t i = 0; i < dmi_pairs_in.size(); ++i )
        {
            const auto & p = dmi_pairs_in[i];
            const auto & t = p.translations;
            dmi_pairs.push_back( Pair{ p.j, p.i, { -t[0], -t[1], -t[2] } } );
            dmi_magnitudes.push_back( dmi_magnitudes_in[i] );
            dmi_normals.push_back( -dmi_normals_in[i] );
        }
    }

    Update_Energy_Contributions();
    ++revision;
}

void Hamiltonian::Update_Energy_Contributions()
{
    contribution_names.clear();
    auto add = [this]( const char * name ) {
        contribution_names.emplace_back( name );
        return int( contribution_names.size() ) - 1;
    };
    idx_zeeman           = std::abs( external_field_magnitude ) > 1e-60 ? add( "Zeeman" ) : -1;
    idx_anisotropy       = !anisotropy_indices.empty() ? add( "Anisotropy" ) : -1;
    idx_cubic_anisotropy = !cubic_anisotropy_indices.empty() ? add( "Cubic anisotropy" ) : -1;
    idx_exchange         = !exchange_pairs.empty() ? add( "Exchange" ) : -1;
    idx_dmi              = !dmi_pairs.empty() ? add( "DMI" ) : -1;
    idx_ddi              = ddi_method != DDI_Method::None ? add( "DDI" ) : -1;
    ++revision;
}

} // namespace sb
