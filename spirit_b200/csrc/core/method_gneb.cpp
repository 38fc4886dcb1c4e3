#include "method_gneb.hpp"

#include "io.hpp"
#include "ovf.hpp"
#include "constants.hpp"
#include "logging.hpp"

#include <algorithm>

#include <cmath>
#include <stdexcept>

namespace sb
{

// core/src/utility/Cubic_Hermite_Spline.cpp:11-48
std::vector<std::vector<double>> cubic_hermite_interpolate(
    const std::vector<double> & x, const std::vector<double> & p, const std::vector<double> & m, int n_interpolations )
{
    const std::size_t n_points = p.size() + ( p.size() - 1 ) * n_interpolations;
    std::vector<std::vector<double>> result( 2, std::vector<double>( n_points ) );
    for( std::size_t i = 0; i + 1 < p.size(); ++i )
    {
        const double x0 = x[i], x1 = x[i + 1], p0 = p[i], p1 = p[i + 1], m0 = m[i], m1 = m[i + 1];
        for( int j = 0; j < n_interpolations + 1; ++j )
        {
            const double t   = j / double( n_interpolations + 1 );
            const double t2  = t * t, t3 = t2 * t;
            const double h00 = 2 * t3 - 3 * t2 + 1, h10 = -2 * t3 + 3 * t2, h01 = t3 - 2 * t2 + t, h11 = t3 - t2;
            const std::size_t idx = i * ( n_interpolations + 1 ) + j;
            result[0][idx]        = x0 + t * ( x1 - x0 );
            result[1][idx]        = h00 * p0 + h10 * p1 + h01 * m0 * ( x0 - x1 ) + h11 * m1 * ( x0 - x1 );
        }
    }
    result[0].back() = x.back();
    result[1].back() = p.back();
    return result;
}


// ---------------------------------------------------------------------------------------------
// Method_GNEB (core/src/engine/Method_GNEB.cpp:23-456)
// ---------------------------------------------------------------------------------------------
Method_GNEB::Method_GNEB( std::shared_ptr<Chain> chain_, int solver_, int idx_chain_ )
        : Method( chain_->gneb_parameters, -1, idx_chain_ ), chain( std::move( chain_ ) )
{
    solver          = solver_;
    const auto & P  = *chain->gneb_parameters;

    // We assume that the chain is not converged before the first iteration (Method_GNEB.cpp:53-55)
    max_torque     = P.force_convergence + 1.0;
    max_torque_all = std::vector<double>( chain->noi, 0.0 );

    // The device chain evaluates every image with ONE set of tables (images of a chain are copies of one another). The reference
    // calls each image's own Hamiltonian and llg dt (Method_GNEB.cpp:99-100, 359-391): a chain whose images differ in what the
    // tables are built from is refused instead of being computed with image 0's parameters.
    {
        const auto & h0 = *chain->images[0]->hamiltonian;
        for( int i = 1; i < chain->noi; ++i )
        {
            const auto & h = *chain->images[i]->hamiltonian;
            const bool same = h.boundary_conditions == h0.boundary_conditions && h.external_field_magnitude == h0.external_field_magnitude
                              && h.external_field_normal.x == h0.external_field_normal.x && h.external_field_normal.y == h0.external_field_normal.y
                              && h.external_field_normal.z == h0.external_field_normal.z && h.anisotropy_magnitudes == h0.anisotropy_magnitudes
                              && h.cubic_anisotropy_magnitudes == h0.cubic_anisotropy_magnitudes && h.exchange_magnitudes == h0.exchange_magnitudes
                              && h.dmi_magnitudes == h0.dmi_magnitudes && h.ddi_method == h0.ddi_method
                              && chain->images[i]->llg_parameters->dt == chain->images[0]->llg_parameters->dt;
            if( chain->images[i]->geometry->site_flags != chain->images[0]->geometry->site_flags )
                throw std::runtime_error(
                    "spirit_b200: GNEB over images with different pinned sites or defects (image " + std::to_string( i )
                    + " differs from image 0) is not implemented: the device chain uses one set of site flags" );
            if( !same )
                throw std::runtime_error(
                    "spirit_b200: GNEB over images with different Hamiltonian parameters or llg_dt (image " + std::to_string( i )
                    + " differs from image 0) is not implemented: the device chain uses one set of interaction tables" );
        }
    }
    device_ = std::make_unique<dev::DeviceChain>( *chain->images[0]->geometry, chain->noi, chain->shard_begin, chain->shard_noi_global );
    device_->set_hamiltonian( *chain->images[0]->hamiltonian );
    Sync_Device();
    device_->vp_reset();

    // Data of the border images, which are not updated (Method_GNEB.cpp:68-72)
    if( chain->shard_noi_global < 0 )
    {
        chain->images.front()->UpdateEffectiveField();
        chain->images.back()->UpdateEffectiveField();
    }
}

Method_GNEB::~Method_GNEB() = default;

static dev::GNEBParams gneb_params( const Chain & chain )
{
    dev::GNEBParams g;
    g.spring_constant = chain.gneb_parameters->spring_constant;
    g.dt              = chain.images[0]->llg_parameters->dt;
    g.dtg             = g.dt * constants::gamma / constants::mu_B;
    for( int i = 0; i < chain.noi; ++i )
        g.image_type.push_back( int( chain.image_type[i] ) );
    const auto & P               = *chain.gneb_parameters;
    g.spring_force_ratio         = P.spring_force_ratio;
    g.path_shortening_constant   = P.path_shortening_constant;
    g.moving_endpoints           = P.moving_endpoints;
    g.translating_endpoints      = P.translating_endpoints;
    g.escape_first               = P.escape_first;
    g.equilibrium_delta_Rx_left  = P.equilibrium_delta_Rx_left;
    g.equilibrium_delta_Rx_right = P.equilibrium_delta_Rx_right;
    return g;
}

void Method_GNEB::Iteration( bool hook_follows )
{
    device_->set_hamiltonian( *chain->images[0]->hamiltonian );
    device_->iterate( solver, gneb_params( *chain ), 1, hook_follows, hook_follows ? &pending_ : nullptr );
    hook_pending_ = hook_follows;
    evaluated_    = true;
}

// Method_GNEB.cpp:410-456
void Method_GNEB::Hook_Post_Iteration()
{
    if( !hook_pending_ )
        return;
    hook_pending_ = false;
    if( pending_.degenerate )
    {
        Log( Log_Level::Error, Log_Sender::GNEB, "The geodesic distance between two images is zero! Stopping...", -1, idx_chain );
        chain->iteration_allowed = false;
        return;
    }
    for( int img = 0; img < chain->noi; ++img )
        max_torque_all[img] = pending_.max_torque[img];
    max_torque = pending_.max_torque_chain; // over all images, also those held by other ranks
    auto interp = cubic_hermite_interpolate( pending_.Rx, pending_.energy, pending_.dE_dRx, chain->gneb_parameters->n_E_interpolations );
    chain->Rx   = pending_.Rx;
    for( int img = 0; img < chain->noi; ++img )
        chain->images[img]->E = pending_.energy[img];
    chain->Rx_interpolated = interp[0];
    chain->E_interpolated  = interp[1];
}

bool Method_GNEB::Converged()
{
    return max_torque < chain->gneb_parameters->force_convergence;
}

void Method_GNEB::Finalize()
{
    chain->iteration_allowed = false;
}

// Method_GNEB.cpp:600-715: histories, and with gneb_output_any the chain (one OVF segment per image) and its energy table at
// the start, at the end and per log step. Interpolated energy tables (gneb_output_energies_interpolated) are not written.
void Method_GNEB::Save_Current( bool initial, bool final )
{
    history_iteration.push_back( int( iteration ) );
    history_max_torque.push_back( max_torque );
    history_energy.push_back( chain->images[chain->idx_active_image]->E );

    const Parameters_GNEB & P = *chain->gneb_parameters;
    if( !P.output_any )
        return;
    char s_iter[32];
    std::snprintf( s_iter, sizeof( s_iter ), "%06ld", iteration );
    const std::string tag  = P.output_file_tag == "<time>" ? starttime + "_" : ( P.output_file_tag.empty() ? "" : P.output_file_tag + "_" );
    const std::string pre  = P.output_folder + "/" + tag + "Chain";
    const std::string base = "GNEB simulation (" + SolverFullName() + " solver)\n# Desc:      Iteration: " + std::to_string( iteration )
                             + "\n# Desc:      Maximum torque: " + io::shortest( max_torque );
    auto write_chain = [&]( const std::string & suffix )
    {
        try
        {
            ovf::File file( pre + suffix + ".ovf", ovf::File::ForWriting{} );
            for( int i = 0; i < chain->noi; ++i )
            {
                const ovf::Segment seg = io::spin_segment(
                    *chain->images[i], base + "\n# Desc: Image " + std::to_string( i ) + " of " + std::to_string( chain->noi ) );
                if( i == 0 )
                    file.write_segment( seg, chain->images[i]->spins.scalars(), P.output_vf_filetype );
                else
                    file.append_segment( seg, chain->images[i]->spins.scalars(), P.output_vf_filetype );
            }
        }
        catch( const std::exception & e )
        {
            Log( Log_Level::Error, Log_Sender::GNEB, std::string( "GNEB output failed: " ) + e.what(), -1, idx_chain );
        }
    };
    auto write_energies = [&]( const std::string & suffix )
    {
        io::write_chain_energies( *chain, pre + "_Energies" + suffix + ".txt", P.output_energies_divide_by_nspins, P.output_energies_add_readability_lines );
        if( P.output_energies_interpolated )
            Log( Log_Level::Warning, Log_Sender::GNEB, "gneb_output_energies_interpolated: interpolated energy tables are not written", -1, idx_chain );
    };
    if( initial && P.output_initial )
    {
        write_chain( "-initial" );
        write_energies( "-initial" );
    }
    else if( final && P.output_final )
    {
        write_chain( "-final" );
        write_energies( "-final" );
    }
    if( P.output_chain_step )
        write_chain( std::string( "_" ) + s_iter );
    if( P.output_energies_step )
        write_energies( std::string( "_" ) + s_iter );
}

void Method_GNEB::Sync_Host()
{
    for( int img = 0; img < chain->noi; ++img )
    {
        device_->download_image( img, chain->images[img]->spins.scalars() );
        if( evaluated_ )
            device_->download_effective_field( img, chain->images[img]->effective_field.scalars() );
    }
}

void Method_GNEB::Sync_Device()
{
    for( int img = 0; img < chain->noi; ++img )
        device_->upload_image( img, chain->images[img]->spins.scalars() );
}

} // namespace sb
