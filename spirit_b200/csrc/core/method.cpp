#include "method.hpp"

#include "io.hpp"
#include "ovf.hpp"
#include "constants.hpp"
#include "logging.hpp"

#include <algorithm>
#include <cmath>
#include <fstream>

namespace sb
{

Method::Method( std::shared_ptr<Parameters_Method> parameters_, int idx_image_, int idx_chain_ )
        : idx_image( idx_image_ ), idx_chain( idx_chain_ ), parameters( std::move( parameters_ ) )
{
    // Method.cpp:15-47
    n_iterations_amortize = parameters->n_iterations_amortize;
    n_iterations          = std::max( 1L, parameters->n_iterations );
    n_iterations_log      = std::min( parameters->n_iterations_log, n_iterations );
    if( n_iterations_log <= 0 )
        n_iterations_log = n_iterations;
    n_log = n_iterations / n_iterations_log;
    if( n_iterations_amortize > n_iterations_log )
        n_iterations_amortize = n_iterations_log;
    if( n_iterations_amortize < 1 )
        n_iterations_amortize = 1;
    for( int i = 0; i < 7; ++i )
        t_iterations.push_back( std::chrono::system_clock::now() );
    t_start = t_last = std::chrono::system_clock::now();
    starttime        = io::current_date_time();
}

void Method::Iterate()
{
    starttime = io::current_date_time();
    t_start = t_last = std::chrono::system_clock::now();
    auto t_current   = t_start;

    Save_Current( true, false );

    for( iteration = 0;
         ContinueIterating() && !Walltime_Expired( std::chrono::duration<double>( t_current - t_start ).count() );
         iteration += n_iterations_amortize )
    {
        t_current = std::chrono::system_clock::now();
        Lock();
        Hook_Pre_Iteration();
        for( long i = 0; i < n_iterations_amortize; ++i )
            Iteration( i == n_iterations_amortize - 1 );
        Hook_Post_Iteration();

        t_iterations.pop_front();
        t_iterations.push_back( std::chrono::system_clock::now() );

        if( n_iterations_log > 0 && iteration > 0 && 0 == std::fmod( double( iteration ), double( n_iterations_log ) ) )
        {
            ++step;
            Sync_Host();
            Save_Current( false, false );
            t_last = std::chrono::system_clock::now();
        }
        Unlock();
    }

    Sync_Host();
    Finalize();
    step = iteration / n_iterations_log;
    Save_Current( false, true );
}

void Method::Save_Current( bool, bool ) {}

std::string Method::SolverName()
{
    switch( solver )
    {
        case dev::Solver_VP: return "VP";
        case dev::Solver_SIB: return "SIB";
        case dev::Solver_Depondt: return "Depondt";
        case dev::Solver_Heun: return "Heun";
        case dev::Solver_RK4: return "RK4";
        case dev::Solver_LBFGS_OSO: return "LBFGS_OSO";
        case dev::Solver_VP_OSO: return "VP_OSO";
        case dev::Solver_LBFGS_Atlas: return "LBFGS_Atlas";
        default: return "--";
    }
}

std::string Method::SolverFullName()
{
    switch( solver )
    {
        case dev::Solver_VP: return "Velocity Projection";
        case dev::Solver_SIB: return "Semi-implicit B";
        case dev::Solver_Depondt: return "Depondt";
        case dev::Solver_Heun: return "Heun";
        case dev::Solver_RK4: return "Runge Kutta (4th order)";
        case dev::Solver_LBFGS_OSO: return "Limited memory Broyden-Fletcher-Goldfarb-Shanno using exponential transforms";
        case dev::Solver_VP_OSO: return "Velocity Projection using exponential transforms";
        case dev::Solver_LBFGS_Atlas: return "Limited memory Broyden-Fletcher-Goldfarb-Shanno using stereographic atlas";
        default: return "--";
    }
}

// Method.cpp:178-200
bool Method::ContinueIterating()
{
    if( !( iteration < n_iterations && Iterations_Allowed() ) )
        return false;
    std::ifstream f( "STOP" );
    if( f.good() )
        return false;
    return !Converged();
}

bool Method::Walltime_Expired( double seconds ) const
{
    if( parameters->max_walltime_sec <= 0 )
        return false;
    return seconds > double( parameters->max_walltime_sec );
}

double Method::getIterationsPerSecond()
{
    double l_ips = 0;
    for( std::size_t i = 0; i + 1 < t_iterations.size(); ++i )
        l_ips += std::chrono::duration<double>( t_iterations[i + 1] - t_iterations[i] ).count();
    return 1.0 / ( l_ips / double( t_iterations.size() - 1 ) ) * double( n_iterations_amortize );
}

std::int64_t Method::getWallTime() const
{
    auto dt = std::chrono::system_clock::now() - t_start;
    return std::chrono::duration_cast<std::chrono::milliseconds>( dt ).count();
}

// ---------------------------------------------------------------------------------------------
// Method_LLG
// ---------------------------------------------------------------------------------------------
Method_LLG::Method_LLG( std::shared_ptr<Spin_System> system_, int solver_, int idx_image_, int idx_chain_ )
        : Method( system_->llg_parameters, idx_image_, idx_chain_ ), system( std::move( system_ ) )
{
    solver = solver_;
    if( solver != dev::Solver_VP && solver != dev::Solver_SIB && solver != dev::Solver_Depondt
        && solver != dev::Solver_Heun && solver != dev::Solver_RK4 && solver != dev::Solver_LBFGS_OSO
        && solver != dev::Solver_LBFGS_Atlas && solver != dev::Solver_VP_OSO )
        throw std::runtime_error(
            "Solver " + std::to_string( solver )
            + " is not implemented in spirit_b200 (available: VP 0, SIB 1, Depondt 2, Heun 3, RK4 4, LBFGS_OSO 5, LBFGS_Atlas 6, VP_OSO 7)" );

    // We assume it is not converged before the first iteration (Method_LLG.cpp:44-46)
    max_torque = system->llg_parameters->force_convergence + 1.0;

    // Constructor-time force evaluation + hook (Method_LLG.cpp:57-62)
    system->device_is_newer = false; // a new method starts from the host configuration
    system->sync_to_device();
    system->device().oso_reset(); // Method_Solver<...OSO>::Initialize: zero velocity / empty L-BFGS memory
    llg_          = make_params( *system, solver );
    hook_pending_ = true;
    system->device().llg_initial_hook( solver, llg_, &pending_hook_ );
    ++system->llg_parameters->philox_counter;
    Hook_Post_Iteration();
    // the constructor-time hook does not count as simulated time (Method_LLG.cpp:24 starts at 0 and the hook adds dt;
    // the reference has the same off-by-one, which Simulation_Get_Time exposes -- keep it)
}

// Method_LLG.cpp:131-226 (prefactors) and :65-110 (thermal amplitude)
dev::LLGParams Method_LLG::make_params( const Spin_System & system, int solver )
{
    namespace C    = constants;
    const auto & P = *system.llg_parameters;
    const auto & g = *system.geometry;
    dev::LLGParams l{};
    const bool lbfgs      = solver == dev::Solver_LBFGS_OSO || solver == dev::Solver_LBFGS_Atlas;
    const bool minimise   = P.direct_minimization || solver == dev::Solver_VP || solver == dev::Solver_VP_OSO || lbfgs;
    l.damping             = P.damping;
    l.dt                  = P.dt;
    l.direct_minimization = minimise ? 1 : 0;
    if( lbfgs )
        l.dtg = 1.0; // Fv = s x F (Method_LLG.cpp:163-166)
    else if( minimise )
        l.dtg = P.dt * C::gamma / C::mu_B;
    else
        l.dtg = P.dt * C::gamma / C::mu_B / ( 1 + P.damping * P.damping );

    // STT, monolayer approximation only (the gradient approximation is SURVEY.md 8f rank 4)
    l.has_stt = 0;
    if( !minimise && P.stt_magnitude > 0 )
    {
        l.has_stt = 1;
        if( P.stt_use_gradient )
        {
            // gradient approximation for in-plane currents (Method_LLG.cpp:184-205): s_c_grad = jacobian(s) je, the jacobian from
            // finite differences along the lattice translations times the inverse of the matrix of lattice vectors
            if( const_cast<Spin_System &>( system ).device().is_slab() )
                throw std::runtime_error( "spirit_b200: llg_stt_use_gradient 1 is not implemented on a lattice cut into slabs over GPUs" );
            double B[3][3]; // columns: lattice_constant * bravais vectors
            for( int c = 0; c < 3; ++c )
                for( int r = 0; r < 3; ++r )
                    B[r][c] = g.lattice_constant * g.bravais_vectors[c][r];
            const double det = B[0][0] * ( B[1][1] * B[2][2] - B[1][2] * B[2][1] ) - B[0][1] * ( B[1][0] * B[2][2] - B[1][2] * B[2][0] )
                               + B[0][2] * ( B[1][0] * B[2][1] - B[1][1] * B[2][0] );
            double inv[3][3];
            inv[0][0] = ( B[1][1] * B[2][2] - B[1][2] * B[2][1] ) / det, inv[0][1] = ( B[0][2] * B[2][1] - B[0][1] * B[2][2] ) / det;
            inv[0][2] = ( B[0][1] * B[1][2] - B[0][2] * B[1][1] ) / det, inv[1][0] = ( B[1][2] * B[2][0] - B[1][0] * B[2][2] ) / det;
            inv[1][1] = ( B[0][0] * B[2][2] - B[0][2] * B[2][0] ) / det, inv[1][2] = ( B[0][2] * B[1][0] - B[0][0] * B[1][2] ) / det;
            inv[2][0] = ( B[1][0] * B[2][1] - B[1][1] * B[2][0] ) / det, inv[2][1] = ( B[0][1] * B[2][0] - B[0][0] * B[2][1] ) / det;
            inv[2][2] = ( B[0][0] * B[1][1] - B[0][1] * B[1][0] ) / det;
            for( int t = 0; t < 3; ++t )
                l.stt_w[t] = inv[t][0] * P.stt_polarisation_normal[0] + inv[t][1] * P.stt_polarisation_normal[1]
                             + inv[t][2] * P.stt_polarisation_normal[2];
            l.stt_g1  = l.dtg * P.stt_magnitude * ( P.damping - P.beta );
            l.stt_g2  = l.dtg * P.stt_magnitude * ( 1 + P.beta * P.damping );
            l.has_stt = 2;
        }
        l.stt_c1  = -l.dtg * P.stt_magnitude * ( P.damping - P.beta );
        l.stt_c2  = -l.dtg * P.stt_magnitude * ( 1 + P.beta * P.damping );
        for( int d = 0; d < 3; ++d )
            l.stt_pol[d] = P.stt_polarisation_normal[d];
    }

    // Method_LLG.cpp:72: a thermal field exists for T > 0 or a non-zero temperature gradient
    l.has_tgrad   = ( !minimise && P.temperature_gradient_inclination != 0 ) ? 1 : 0;
    l.has_thermal = ( !minimise && ( P.temperature > 0 || l.has_tgrad ) ) ? 1 : 0;
    const double epsilon = std::sqrt( 2 * P.damping * P.dt * C::gamma / C::mu_B * C::k_B ) / ( 1 + P.damping * P.damping );
    for( int ib = 0; ib < dev::MAX_BASIS; ++ib )
    {
        const double mu      = ib < g.n_cell_atoms ? g.cell_mu_s[ib] : 1.0;
        l.inv_mu_s[ib]       = 1.0 / mu;
        l.c1[ib]             = l.dtg / mu;
        l.c2[ib]             = l.damping * l.dtg / mu;
        l.nc1[ib]            = -l.c1[ib];
        l.nc2[ib]            = -l.c2[ib];
        l.thermal_scale[ib]  = l.has_thermal ? epsilon * std::sqrt( P.temperature / mu ) : 0.0;
        l.half_nc1[ib]       = 0.5 * l.nc1[ib];
        l.half_nc2[ib]       = 0.5 * l.nc2[ib];
        l.thermal_k[ib]      = float( -2.0 * 0.69314718055994531 * l.thermal_scale[ib] * l.thermal_scale[ib] );
    }
    if( l.has_tgrad )
    {
        // T_i = inclination * (d . r_i) + T - inclination * min(d . bounds_min, d . bounds_max)  (Vectormath.cpp:633-652),
        // and r_i is affine in the cell indices: r_i = lc * ( (a + u_a) t_a + (b + u_b) t_b + (c + u_c) t_c )
        Vec3 d = P.temperature_gradient_direction;
        d.normalize();
        const double incl = P.temperature_gradient_inclination;
        const double dmin = std::min( g.bounds_min.dot( d ), g.bounds_max.dot( d ) );
        l.tgrad_T0        = P.temperature - incl * dmin;
        for( int k = 0; k < 3; ++k )
            l.tgrad_cell[k] = incl * g.lattice_constant * g.bravais_vectors[k].dot( d );
        for( int ib = 0; ib < dev::MAX_BASIS; ++ib )
        {
            l.tgrad_basis[ib] = 0;
            if( ib < g.n_cell_atoms )
                for( int k = 0; k < 3; ++k )
                    l.tgrad_basis[ib] += incl * g.lattice_constant * g.cell_atoms[ib][k] * g.bravais_vectors[k].dot( d );
            const double mu        = ib < g.n_cell_atoms ? g.cell_mu_s[ib] : 1.0;
            l.thermal_k_per_T[ib] = float( -2.0 * 0.69314718055994531 * epsilon * epsilon / mu );
        }
    }
    l.half_damping = 0.5 * l.damping;
    l.half_ndtg    = -0.5 * l.dtg;
    l.seed      = std::uint64_t( std::uint32_t( P.rng_seed ) ) | ( std::uint64_t( 0x5b200 ) << 32 );
    l.iteration = P.philox_counter;
    for( unsigned r = 0; r < 10; ++r )
    {
        l.philox_key[r][0] = unsigned( l.seed ) + r * 0x9E3779B9u;
        l.philox_key[r][1] = unsigned( l.seed >> 32 ) + r * 0xBB67AE85u;
    }
    return l;
}

void Method_LLG::Iteration( bool hook_follows )
{
    llg_ = make_params( *system, solver );
    system->device().set_hamiltonian( *system->hamiltonian );
    if( solver == dev::Solver_LBFGS_OSO || solver == dev::Solver_VP_OSO || solver == dev::Solver_LBFGS_Atlas )
        system->device().oso_iterate( solver, llg_, 1, hook_follows, hook_follows ? &pending_hook_ : nullptr );
    else
        system->device().llg_iterate( solver, llg_, 1, hook_follows, hook_follows ? &pending_hook_ : nullptr );
    ++system->llg_parameters->philox_counter;
    hook_pending_           = hook_follows;
    system->device_is_newer = true;
}

double Method_LLG::Iterate_Device_Resident( const std::shared_ptr<Spin_System> & system, int solver, int n_iterations )
{
    if( solver < dev::Solver_VP || solver > dev::Solver_RK4 )
        throw std::runtime_error( "Solver " + std::to_string( solver ) + " is not implemented in spirit_b200" );
    auto & d = system->device();
    d.set_hamiltonian( *system->hamiltonian );
    dev::LLGParams l = make_params( *system, solver );
    dev::HookResult result;
    if( solver == dev::Solver_VP )
        d.llg_initial_hook( solver, l, &result ); // F_prev of the first VP iteration
    d.timer_start();
    d.llg_iterate( solver, l, n_iterations, true, &result );
    const double ms = d.timer_stop();
    system->llg_parameters->philox_counter = l.iteration;
    system->E                              = result.energy;
    system->device_is_newer                = true; // until SpiritB200_Download / the next upload
    return ms;
}

// Method_LLG.cpp:246-301
void Method_LLG::Hook_Post_Iteration()
{
    picoseconds_passed += system->llg_parameters->dt;
    if( !hook_pending_ )
        return;
    hook_pending_    = false;
    force_converged_ = false;
    double fmax      = pending_hook_.max_torque;
    max_torque       = fmax > 0 ? fmax : 0;
    if( fmax < system->llg_parameters->force_convergence )
        force_converged_ = true;
    system->E = pending_hook_.energy;
}

bool Method_LLG::Converged()
{
    return force_converged_;
}

void Method_LLG::Finalize()
{
    system->iteration_allowed = false;
}

// Method_LLG.cpp:310-500: the histories, and -- with llg_output_any -- the files of the reference: spins at the start, at the
// end, per log step and as an appended archive (OVF, format llg_output_vf_filetype), energy tables next to them. The spins
// are the host copies (Sync_Host precedes every call). Per-spin energy files (llg_output_energy_spin_resolved) are not written.
void Method_LLG::Save_Current( bool initial, bool final )
{
    history_iteration.push_back( int( iteration ) );
    history_max_torque.push_back( max_torque );
    history_energy.push_back( system->E );

    const Parameters_LLG & P = *system->llg_parameters;
    if( !P.output_any )
        return;
    char s_img[16];
    std::snprintf( s_img, sizeof( s_img ), "%02d", idx_image );
    const int width = P.n_iterations > 0 ? int( std::log10( double( P.n_iterations ) ) ) : 0;
    char s_iter[32];
    std::snprintf( s_iter, sizeof( s_iter ), "%0*ld", width, iteration );
    const std::string tag    = P.output_file_tag == "<time>" ? starttime + "_" : ( P.output_file_tag.empty() ? "" : P.output_file_tag + "_" );
    const std::string spins  = P.output_folder + "/" + tag + "Image-" + s_img + "_Spins";
    const std::string energy = P.output_folder + "/" + tag + "Image-" + s_img + "_Energy";

    auto write_configuration = [&]( const std::string & suffix, bool append )
    {
        try
        {
            ovf::Segment seg = io::spin_segment(
                *system, "LLG simulation (" + SolverFullName() + " solver)\n# Desc:      Iteration: " + std::to_string( iteration )
                             + "\n# Desc:      Maximum torque: " + io::shortest( max_torque ) );
            ovf::File file( spins + suffix + ".ovf", ovf::File::ForWriting{} );
            if( append )
                file.append_segment( seg, system->spins.scalars(), P.output_vf_filetype );
            else
                file.write_segment( seg, system->spins.scalars(), P.output_vf_filetype );
        }
        catch( const std::exception & e )
        {
            Log( Log_Level::Error, Log_Sender::LLG, std::string( "LLG output failed: " ) + e.what(), idx_image, idx_chain );
        }
    };
    auto write_energy = [&]( const std::string & suffix, bool append )
    {
        const std::string file = energy + suffix + ".txt";
        if( !append || !std::ifstream( file ).good() )
            io::write_energy_header( *system, file, { "iteration", "E_tot" }, P.output_energy_add_readability_lines );
        io::append_image_energy( *system, iteration, file, P.output_energy_divide_by_nspins, P.output_energy_add_readability_lines );
        if( !append && P.output_energy_spin_resolved )
            Log( Log_Level::Warning, Log_Sender::LLG, "llg_output_energy_spin_resolved: per-spin energy files are not written", idx_image, idx_chain );
    };
    if( initial && P.output_initial )
    {
        write_configuration( "-initial", false );
        write_energy( "-initial", false );
    }
    else if( final && P.output_final )
    {
        write_configuration( "-final", false );
        write_energy( "-final", false );
    }
    if( P.output_configuration_step )
        write_configuration( std::string( "_" ) + s_iter, false );
    if( P.output_energy_step )
        write_energy( std::string( "_" ) + s_iter, false );
    if( P.output_configuration_archive )
        write_configuration( "-archive", true );
    if( P.output_energy_archive )
        write_energy( "-archive", true );
}

void Method_LLG::Sync_Host()
{
    auto & d = system->device();
    d.download_spins( system->spins.scalars() );
    system->effective_field_stale = true; // mirrored on demand (Spin_System::refresh_effective_field_mirror)
    system->device_is_newer       = false;
}

void Method_LLG::Sync_Device()
{
    system->device_is_newer = false;
    system->sync_to_device();
}

} // namespace sb
