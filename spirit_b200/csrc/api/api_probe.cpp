// spirit_b200 extensions (include/spirit_b200.h): double-precision probes and device control.
#include "api_common.hpp"

#include "../core/method_gneb.hpp"

#include <spirit_b200.h>

#include <cstring>

using namespace sb;

namespace
{
thread_local std::string device_name_buffer;

// Upload either the caller's spins or the image's own to the device
dev::DeviceImage & device_with_spins( Spin_System & image, const double * spins )
{
    auto & d = image.device();
    d.set_hamiltonian( *image.hamiltonian );
    d.upload_spins( spins ? spins : image.spins.scalars() );
    return d;
}
} // namespace

int SpiritB200_Gradient_and_Energy( State * state, const double * spins, double * gradient, double * energy, int idx_image ) noexcept
try
{
    int idx_chain = -1;
    auto image    = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    device_with_spins( *image, spins ).gradient_and_energy( gradient, energy );
    return image->nos;
}
catch( ... )
{
    handle_exception_api( __func__, idx_image );
    return -1;
}

int SpiritB200_Gradient( State * state, const double * spins, double * gradient, int idx_image ) noexcept
try
{
    int idx_chain = -1;
    auto image    = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    device_with_spins( *image, spins ).gradient_and_energy( gradient, nullptr );
    return image->nos;
}
catch( ... )
{
    handle_exception_api( __func__, idx_image );
    return -1;
}

int SpiritB200_Energy_Contributions(
    State * state, const double * spins, int max_terms, char * names, double * totals, double * per_spin, int idx_image ) noexcept
try
{
    int idx_chain = -1;
    auto image    = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    auto & d          = device_with_spins( *image, spins );
    const auto & ham  = *image->hamiltonian;
    const int n_terms = int( ham.contribution_names.size() );
    if( n_terms > max_terms )
        throw std::runtime_error( "SpiritB200_Energy_Contributions: max_terms is smaller than the number of contributions" );
    std::vector<double> t( std::max( 1, n_terms ), 0.0 );
    d.energy_contributions( ham, t.data(), per_spin );
    for( int i = 0; i < n_terms; ++i )
    {
        std::strncpy( names + 32 * i, ham.contribution_names[i].c_str(), 31 );
        names[32 * i + 31] = 0;
        totals[i]          = t[i];
    }
    return n_terms;
}
catch( ... )
{
    handle_exception_api( __func__, idx_image );
    return -1;
}

double SpiritB200_Get_Energy( State * state, int idx_image ) noexcept
try
{
    int idx_chain = -1;
    return resolve( state, idx_image, idx_chain ).image->E;
}
catch( ... )
{
    handle_exception_api( __func__, idx_image );
    return 0;
}

int SpiritB200_Get_Pairs( State * state, int kind, int max_pairs, int * ijt, double * magnitudes, double * normals, int idx_image ) noexcept
try
{
    int idx_chain      = -1;
    auto image         = resolve( state, idx_image, idx_chain ).image;
    const auto & ham   = *image->hamiltonian;
    const auto & pairs = kind == 0 ? ham.exchange_pairs : ham.dmi_pairs;
    const auto & mags  = kind == 0 ? ham.exchange_magnitudes : ham.dmi_magnitudes;
    const int n        = int( pairs.size() );
    for( int p = 0; p < n && p < max_pairs; ++p )
    {
        ijt[5 * p + 0] = pairs[p].i;
        ijt[5 * p + 1] = pairs[p].j;
        for( int d = 0; d < 3; ++d )
            ijt[5 * p + 2 + d] = pairs[p].translations[d];
        magnitudes[p] = mags[p];
        if( kind == 1 && normals )
            for( int d = 0; d < 3; ++d )
                normals[3 * p + d] = ham.dmi_normals[p][d];
    }
    return n;
}
catch( ... )
{
    handle_exception_api( __func__, idx_image );
    return -1;
}

double SpiritB200_Get_MaxTorque( State * state, int idx_image ) noexcept
try
{
    if( idx_image == -2 )
        return state->method_chain ? state->method_chain->max_torque : 0.0;
    int idx_chain = -1;
    resolve( state, idx_image, idx_chain );
    auto & m = state->method_image[idx_image];
    return m ? m->max_torque : 0.0;
}
catch( ... )
{
    handle_exception_api( __func__, idx_image );
    return 0;
}

int SpiritB200_Chain_Get_Rx_E( State * state, double * Rx, double * E ) noexcept
try
{
    int idx_image = -1, idx_chain = -1;
    auto chain = resolve( state, idx_image, idx_chain ).chain;
    for( int i = 0; i < chain->noi; ++i )
    {
        Rx[i] = chain->Rx[i];
        E[i]  = chain->images[i]->E;
    }
    return chain->noi;
}
catch( ... )
{
    handle_exception_api( __func__ );
    return -1;
}

int SpiritB200_Get_Magnetization( State * state, double * m, int idx_image ) noexcept
try
{
    int idx_chain = -1;
    auto image    = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    device_with_spins( *image, nullptr ).magnetization( m, true );
    return 0;
}
catch( ... )
{
    handle_exception_api( __func__, idx_image );
    return -1;
}

// ---------------------------------------------------------------------------------------------
int SpiritB200_Device_Count() noexcept
{
    return dev::device_count();
}

int SpiritB200_Set_Device( int device ) noexcept
try
{
    dev::set_device( device );
    return 0;
}
catch( ... )
{
    handle_exception_api( __func__ );
    return -1;
}

const char * SpiritB200_Device_Name() noexcept
try
{
    device_name_buffer = dev::device_name();
    return device_name_buffer.c_str();
}
catch( ... )
{
    handle_exception_api( __func__ );
    return "";
}

unsigned long long SpiritB200_Kernel_Launches( State * state, int idx_image ) noexcept
try
{
    int idx_chain = -1;
    auto image    = resolve( state, idx_image, idx_chain ).image;
    return image->has_device() ? image->device().kernel_launches() : 0ull;
}
catch( ... )
{
    handle_exception_api( __func__, idx_image );
    return 0;
}

int SpiritB200_Thermal_Variates( State * state, unsigned long long iteration, unsigned long long count, float * variates, int idx_image ) noexcept
try
{
    int idx_chain = -1;
    auto image    = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    dev::LLGParams l = Method_LLG::make_params( *image, dev::Solver_Depondt );
    l.iteration      = iteration;
    image->device().dump_thermal_variates( l, std::size_t( count ), variates );
    return 0;
}
catch( ... )
{
    handle_exception_api( __func__, idx_image );
    return -1;
}

int SpiritB200_Step_Variant( State * state, int solver_type, int idx_image ) noexcept
try
{
    int idx_chain = -1;
    auto image    = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    auto & d = image->device();
    d.set_hamiltonian( *image->hamiltonian );
    if( solver_type >= dev::Solver_VP && solver_type <= dev::Solver_RK4 )
    {
        const dev::LLGParams l = Method_LLG::make_params( *image, solver_type );
        if( d.fused_usable( solver_type, l ) )
            return 2;
    }
    return d.stencil_variant();
}
catch( ... )
{
    handle_exception_api( __func__, idx_image );
    return -1;
}

int SpiritB200_Stencil_Variant( State * state, int idx_image ) noexcept
try
{
    int idx_chain = -1;
    auto image    = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    auto & d = image->device(); // creates the device image if there is none yet; throws without a CUDA device
    d.set_hamiltonian( *image->hamiltonian );
    return d.stencil_variant();
}
catch( ... )
{
    handle_exception_api( __func__, idx_image );
    return -1;
}

int SpiritB200_Upload( State * state, int idx_image ) noexcept
try
{
    int idx_chain = -1;
    auto image    = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    image->device_is_newer = false; // an explicit upload: the host copy wins
    image->sync_to_device();
    image->device().synchronize();
    return 0;
}
catch( ... )
{
    handle_exception_api( __func__, idx_image );
    return -1;
}

int SpiritB200_Download( State * state, int idx_image ) noexcept
try
{
    int idx_chain = -1;
    auto image    = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    image->device().download_spins( image->spins.scalars() );
    image->device_is_newer = false;
    return 0;
}
catch( ... )
{
    handle_exception_api( __func__, idx_image );
    return -1;
}

// Device-resident stepping: same kernels as Simulation_LLG_Start, timed with CUDA events on the image's stream
double SpiritB200_LLG_Iterate_Device( State * state, int solver_type, int n_iterations, int idx_image ) noexcept
try
{
    int idx_chain = -1;
    auto image    = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    return Method_LLG::Iterate_Device_Resident( image, solver_type, n_iterations );
}
catch( ... )
{
    handle_exception_api( __func__, idx_image );
    return -1;
}

int SpiritB200_LLG_Profile_Stages( State * state, int solver_type, int n_iterations, double * stage_ms, int max_stages, int idx_image ) noexcept
try
{
    int idx_chain = -1;
    auto image    = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    auto & d = image->device();
    d.set_hamiltonian( *image->hamiltonian );
    dev::LLGParams l = Method_LLG::make_params( *image, solver_type );
    const int n      = d.llg_profile_stages( solver_type, l, n_iterations, stage_ms, max_stages );
    image->llg_parameters->philox_counter = l.iteration;
    return n;
}
catch( ... )
{
    handle_exception_api( __func__, idx_image );
    return -1;
}

// ---------------------------------------------------------------------------------------------
int SpiritB200_Comm_Unique_Id( char * id128 ) noexcept
try
{
    dev::comm_unique_id( id128 );
    return 0;
}
catch( ... )
{
    handle_exception_api( __func__ );
    return -1;
}

int SpiritB200_Comm_Init( int rank, int world, const char * id128 ) noexcept
try
{
    dev::comm_init( rank, world, id128 );
    return 0;
}
catch( ... )
{
    handle_exception_api( __func__ );
    return -1;
}

int SpiritB200_Slab_Setup( State * state, int c_begin, int Nc_global, int idx_image ) noexcept
try
{
    int idx_chain = -1;
    auto image    = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    image->drop_device();
    image->device().set_slab( c_begin, Nc_global );
    return 0;
}
catch( ... )
{
    handle_exception_api( __func__, idx_image );
    return -1;
}

int SpiritB200_Chain_Shard_Setup( State * state, int i_begin, int noi_global ) noexcept
try
{
    int idx_image = -1, idx_chain = -1;
    auto chain = resolve( state, idx_image, idx_chain ).chain;
    if( i_begin < 0 || i_begin + chain->noi > noi_global )
        throw std::runtime_error( "SpiritB200_Chain_Shard_Setup: shard outside of the global chain" );
    chain->shard_begin      = i_begin;
    chain->shard_noi_global = noi_global;
    return 0;
}
catch( ... )
{
    handle_exception_api( __func__ );
    return -1;
}
