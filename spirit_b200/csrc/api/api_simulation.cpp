// C API: Simulation_*. Reference behaviour: core/src/Spirit/Simulation.cpp:16-800.
#include "api_common.hpp"

#include "../core/method_gneb.hpp"

#include <Spirit/Simulation.h>

#include <algorithm>
#include <cmath>

using namespace sb;

void free_run_info( Simulation_Run_Info info ) noexcept
{
    delete[] info.history_energy;
    delete[] info.history_iteration;
    delete[] info.history_max_torque;
}

namespace
{
// Simulation.cpp:24-75
void run_method( const std::shared_ptr<Method> & method, bool singleshot, Simulation_Run_Info * info )
{
    if( singleshot )
    {
        method->t_start = method->t_last = std::chrono::system_clock::now();
        method->iteration                = 0;
        method->Save_Current( true, false );
        return;
    }
    method->Iterate();
    if( info )
    {
        info->max_torque       = float( method->max_torque );
        info->total_iterations = int( method->iteration );
        info->total_walltime   = int( method->getWallTime() );
        info->total_ips        = float( info->total_iterations ) / float( info->total_walltime ) * 1000.0f;
        if( !method->history_iteration.empty() )
        {
            info->n_history_iteration = int( method->history_iteration.size() );
            info->history_iteration   = new int[method->history_iteration.size()];
            std::copy( method->history_iteration.begin(), method->history_iteration.end(), info->history_iteration );
        }
        if( !method->history_max_torque.empty() )
        {
            info->n_history_max_torque = int( method->history_max_torque.size() );
            info->history_max_torque   = new float[method->history_max_torque.size()];
            std::copy( method->history_max_torque.begin(), method->history_max_torque.end(), info->history_max_torque );
        }
        if( !method->history_energy.empty() )
        {
            info->n_history_energy = int( method->history_energy.size() );
            info->history_energy   = new float[method->history_energy.size()];
            std::copy( method->history_energy.begin(), method->history_energy.end(), info->history_energy );
        }
    }
}

void out_of_scope( const char * what, int idx_image, int idx_chain )
{
    Log( Log_Level::Error, Log_Sender::API,
         std::string( what ) + " is outside the accelerated hot path of spirit_b200 (LLG and GNEB with VP/SIB/Depondt/Heun/RK4) and was not started",
         idx_image, idx_chain );
}

// Which method is responsible for the image: its own, or the chain's (Simulation.cpp:600-800)
std::shared_ptr<Method> active_method( State * state, int & idx_image, int & idx_chain )
{
    auto r = resolve( state, idx_image, idx_chain );
    if( r.image->iteration_allowed && idx_image < int( state->method_image.size() ) )
        return state->method_image[idx_image];
    if( r.chain->iteration_allowed )
        return state->method_chain;
    return nullptr;
}
} // namespace

void Simulation_MC_Start( State *, int, int, bool, Simulation_Run_Info *, int idx_image, int idx_chain ) noexcept
{
    out_of_scope( "Monte Carlo", idx_image, idx_chain );
}
void Simulation_MMF_Start( State *, int, int, int, bool, Simulation_Run_Info *, int idx_image, int idx_chain ) noexcept
{
    out_of_scope( "Minimum mode following", idx_image, idx_chain );
}
void Simulation_EMA_Start( State *, int, int, bool, Simulation_Run_Info *, int idx_image, int idx_chain ) noexcept
{
    out_of_scope( "Eigenmode analysis", idx_image, idx_chain );
}

// Simulation.cpp:134-217
void Simulation_LLG_Start(
    State * state, int solver_type, int n_iterations, int n_iterations_log, bool singleshot, Simulation_Run_Info * info,
    int idx_image, int idx_chain ) noexcept
try
{
    auto r = resolve( state, idx_image, idx_chain );
    if( r.image->iteration_allowed )
    {
        Log( Log_Level::Warning, Log_Sender::API, "LLG simulation is already running on this image. No action taken.", idx_image, idx_chain );
        return;
    }
    if( r.chain->iteration_allowed )
    {
        Log( Log_Level::Warning, Log_Sender::API, "A simulation is already running on this chain. No action taken.", idx_image, idx_chain );
        return;
    }
    if( solver_type < 0 || solver_type > Solver_VP_OSO )
    {
        Log( Log_Level::Error, Log_Sender::API,
             "Solver " + std::to_string( solver_type )
                 + " is not available in spirit_b200 (solver ids 0 ... 7, core/include/Spirit/Simulation.h:33-54). No action taken.",
             idx_image, idx_chain );
        return;
    }
    std::shared_ptr<Method> method;
    {
        ImageLock lock( *r.image );
        r.image->iteration_allowed  = true;
        r.image->singleshot_allowed = singleshot;
        if( n_iterations > 0 )
            r.image->llg_parameters->n_iterations = n_iterations;
        if( n_iterations_log > 0 )
            r.image->llg_parameters->n_iterations_log = n_iterations_log;
        try
        {
            method = std::make_shared<Method_LLG>( r.image, solver_type, idx_image, idx_chain );
        }
        catch( ... )
        {
            r.image->iteration_allowed  = false;
            r.image->singleshot_allowed = false;
            throw;
        }
    }
    if( int( state->method_image.size() ) <= idx_image )
        state->method_image.resize( idx_image + 1 );
    state->method_image[idx_image] = method;
    run_method( method, singleshot, info );
}
SB_API_CATCH_VOID

// Simulation.cpp:219-315
void Simulation_GNEB_Start(
    State * state, int solver_type, int n_iterations, int n_iterations_log, bool singleshot, Simulation_Run_Info * info,
    int idx_chain ) noexcept
try
{
    int idx_image = -1;
    auto r        = resolve( state, idx_image, idx_chain );
    if( r.chain->iteration_allowed )
    {
        Log( Log_Level::Warning, Log_Sender::API, "GNEB simulation is already running on this chain. No action taken.", -1, idx_chain );
        return;
    }
    for( auto & image : r.chain->images )
        if( image->iteration_allowed )
        {
            Log( Log_Level::Warning, Log_Sender::API, "A simulation is running on an image of this chain. No action taken.", -1, idx_chain );
            return;
        }
    if( ( r.chain->shard_noi_global < 0 ? r.chain->noi : r.chain->shard_noi_global ) < 3 )
    {
        Log( Log_Level::Error, Log_Sender::API, "There are less than 3 images in the chain. GNEB cannot be started.", -1, idx_chain );
        return;
    }
    // The reference's API offers VP, SIB, Depondt, Heun and the OSO / LBFGS solvers for GNEB (Simulation.cpp:278-303); its
    // engine also instantiates Method_GNEB<RK4> (Method_GNEB.cpp:749) without an API branch for it. Accepted here.
    if( solver_type != Solver_VP && solver_type != Solver_SIB && solver_type != Solver_Depondt && solver_type != Solver_Heun
        && solver_type != sb::dev::Solver_RK4 && solver_type != Solver_LBFGS_OSO && solver_type != Solver_LBFGS_Atlas && solver_type != Solver_VP_OSO )
    {
        Log( Log_Level::Error, Log_Sender::API,
             "Solver " + std::to_string( solver_type ) + " is not available for GNEB in spirit_b200 (VP 0, SIB 1, Depondt 2, Heun 3, RK4 4, LBFGS_OSO 5, LBFGS_Atlas 6, VP_OSO 7). No action taken.",
             -1, idx_chain );
        return;
    }
    std::shared_ptr<Method> method;
    r.chain->Lock();
    try
    {
        r.chain->iteration_allowed  = true;
        r.chain->singleshot_allowed = singleshot;
        if( n_iterations > 0 )
            r.chain->gneb_parameters->n_iterations = n_iterations;
        if( n_iterations_log > 0 )
            r.chain->gneb_parameters->n_iterations_log = n_iterations_log;
        method = std::make_shared<Method_GNEB>( r.chain, solver_type, idx_chain );
    }
    catch( ... )
    {
        r.chain->iteration_allowed  = false;
        r.chain->singleshot_allowed = false;
        r.chain->Unlock();
        throw;
    }
    r.chain->Unlock();
    state->method_chain = method;
    run_method( method, singleshot, info );
}
catch( ... )
{
    handle_exception_api( __func__, -1, idx_chain );
}

void Simulation_SingleShot( State * state, int idx_image, int idx_chain ) noexcept
{
    Simulation_N_Shot( state, 1, idx_image, idx_chain );
}

// Simulation.cpp:454-540: hook after every iteration, no amortisation. The caller may have written to the live
// spin array since the last shot, so the spins are re-uploaded first and downloaded afterwards.
void Simulation_N_Shot( State * state, int N, int idx_image, int idx_chain ) noexcept
try
{
    auto r = resolve( state, idx_image, idx_chain );
    std::shared_ptr<Method> method;
    if( r.image->iteration_allowed && r.image->singleshot_allowed )
        method = state->method_image[idx_image];
    else if( r.chain->iteration_allowed && r.chain->singleshot_allowed )
        method = state->method_chain;
    else
    {
        Log( Log_Level::Warning, Log_Sender::API, "No simulation has been started in single-shot mode on this image or chain. No action taken.", idx_image, idx_chain );
        return;
    }

    auto now = std::chrono::system_clock::now();
    if( method->ContinueIterating() && !method->Walltime_Expired( std::chrono::duration<double>( now - method->t_start ).count() ) )
    {
        method->Lock();
        method->Sync_Device();
        for( int i = 0; i < N; ++i )
        {
            method->Hook_Pre_Iteration();
            method->Iteration( true );
            method->Hook_Post_Iteration();
            method->t_iterations.pop_front();
            method->t_iterations.push_back( std::chrono::system_clock::now() );
            if( method->n_iterations_log > 0 && method->iteration > 0
                && 0 == std::fmod( double( method->iteration ), double( method->n_iterations_log ) ) )
            {
                ++method->step;
                method->Sync_Host(); // the files of a log step hold the spins of that step, as in Method::Iterate
                method->Save_Current( false, false );
            }
            ++method->iteration;
        }
        method->Sync_Host();
        method->Unlock();
    }

    now = std::chrono::system_clock::now();
    if( !method->ContinueIterating() || method->Walltime_Expired( std::chrono::duration<double>( now - method->t_start ).count() ) )
    {
        method->step = method->iteration / method->n_iterations_log;
        method->Save_Current( false, true );
        method->Finalize();
        r.image->singleshot_allowed = false;
        r.chain->singleshot_allowed = false;
    }
}
SB_API_CATCH_VOID

// Simulation.cpp:542-598
void Simulation_Stop( State * state, int idx_image, int idx_chain ) noexcept
try
{
    auto r = resolve( state, idx_image, idx_chain );
    if( r.image->iteration_allowed )
    {
        // no lock before clearing the flag: a running Iterate() holds the image lock per outer iteration
        r.image->iteration_allowed = false;
        ImageLock lock( *r.image );
        if( r.image->singleshot_allowed )
        {
            r.image->singleshot_allowed = false;
            auto method                 = state->method_image[idx_image];
            if( method )
            {
                method->step = method->iteration / std::max( 1L, method->n_iterations_log );
                method->Save_Current( false, true );
                method->Finalize();
            }
        }
    }
    else if( r.chain->iteration_allowed )
    {
        r.chain->iteration_allowed = false;
        r.chain->Lock();
        if( r.chain->singleshot_allowed )
        {
            r.chain->singleshot_allowed = false;
            auto method                 = state->method_chain;
            if( method )
            {
                method->step = method->iteration / std::max( 1L, method->n_iterations_log );
                method->Save_Current( false, true );
                method->Finalize();
            }
        }
        r.chain->Unlock();
    }
}
SB_API_CATCH_VOID

void Simulation_Stop_All( State * state ) noexcept
try
{
    if( !state || !state->chain )
        return;
    state->chain->iteration_allowed = false;
    for( int i = 0; i < state->chain->noi; ++i )
        Simulation_Stop( state, i, -1 );
    Simulation_Stop( state, -1, -1 );
}
catch( ... )
{
    handle_exception_api( __func__ );
}

// ---- queries (Simulation.cpp:600-800) ----
float Simulation_Get_MaxTorqueComponent( State * state, int idx_image, int idx_chain ) noexcept
{
    return Simulation_Get_MaxTorqueNorm( state, idx_image, idx_chain );
}

void Simulation_Get_Chain_MaxTorqueComponents( State * state, float * torques, int idx_chain ) noexcept
{
    Simulation_Get_Chain_MaxTorqueNorms( state, torques, idx_chain );
}

float Simulation_Get_MaxTorqueNorm( State * state, int idx_image, int idx_chain ) noexcept
try
{
    auto method = active_method( state, idx_image, idx_chain );
    return method ? float( method->max_torque ) : 0.0f;
}
SB_API_CATCH_RET( 0 )

void Simulation_Get_Chain_MaxTorqueNorms( State * state, float * torques, int idx_chain ) noexcept
try
{
    int idx_image = -1;
    auto r        = resolve( state, idx_image, idx_chain );
    if( r.chain->iteration_allowed && state->method_chain )
    {
        auto all = state->method_chain->getTorqueMaxNorm_All();
        for( std::size_t i = 0; i < all.size(); ++i )
            torques[i] = float( all[i] );
    }
}
catch( ... )
{
    handle_exception_api( __func__, -1, idx_chain );
}

float Simulation_Get_IterationsPerSecond( State * state, int idx_image, int idx_chain ) noexcept
try
{
    auto method = active_method( state, idx_image, idx_chain );
    return method ? float( method->getIterationsPerSecond() ) : 0.0f;
}
SB_API_CATCH_RET( 0 )

int Simulation_Get_Iteration( State * state, int idx_image, int idx_chain ) noexcept
try
{
    auto method = active_method( state, idx_image, idx_chain );
    return method ? int( method->iteration ) : 0;
}
SB_API_CATCH_RET( 0 )

float Simulation_Get_Time( State * state, int idx_image, int idx_chain ) noexcept
try
{
    auto method = active_method( state, idx_image, idx_chain );
    return method ? float( method->get_simulated_time() ) : 0.0f;
}
SB_API_CATCH_RET( 0 )

int Simulation_Get_Wall_Time( State * state, int idx_image, int idx_chain ) noexcept
try
{
    auto method = active_method( state, idx_image, idx_chain );
    return method ? int( method->getWallTime() ) : 0;
}
SB_API_CATCH_RET( 0 )

namespace
{
thread_local std::string name_buffer;
}

const char * Simulation_Get_Solver_Name( State * state, int idx_image, int idx_chain ) noexcept
try
{
    auto method = active_method( state, idx_image, idx_chain );
    name_buffer = method ? method->SolverName() : "";
    return name_buffer.c_str();
}
SB_API_CATCH_RET( "" )

const char * Simulation_Get_Method_Name( State * state, int idx_image, int idx_chain ) noexcept
try
{
    auto method = active_method( state, idx_image, idx_chain );
    name_buffer = method ? method->Name() : "";
    return name_buffer.c_str();
}
SB_API_CATCH_RET( "" )

bool Simulation_Running_On_Image( State * state, int idx_image, int idx_chain ) noexcept
try
{
    return resolve( state, idx_image, idx_chain ).image->iteration_allowed;
}
SB_API_CATCH_RET( false )

bool Simulation_Running_On_Chain( State * state, int idx_chain ) noexcept
try
{
    int idx_image = -1;
    return resolve( state, idx_image, idx_chain ).chain->iteration_allowed;
}
catch( ... )
{
    handle_exception_api( __func__, -1, idx_chain );
    return false;
}

bool Simulation_Running_Anywhere_On_Chain( State * state, int idx_chain ) noexcept
try
{
    int idx_image = -1;
    auto chain    = resolve( state, idx_image, idx_chain ).chain;
    if( chain->iteration_allowed )
        return true;
    for( auto & image : chain->images )
        if( image->iteration_allowed )
            return true;
    return false;
}
catch( ... )
{
    handle_exception_api( __func__, -1, idx_chain );
    return false;
}
