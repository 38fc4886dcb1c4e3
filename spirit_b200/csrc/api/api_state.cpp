// C API: State, Version, Constants, Log.
// Reference behaviour: core/src/Spirit/State.cpp:18-253, Version.cpp, Constants.cpp, Log.cpp.
#include "api_common.hpp"

#include <Spirit/Constants.h>
#include <Spirit/Log.h>
#include <Spirit/Simulation.h>
#include <Spirit/State.h>
#include <Spirit/Version.h>

#include <ctime>
#include <fstream>

using namespace sb;

State * State_Setup( const char * config_file, bool quiet ) noexcept
try
{
    auto * state        = new State();
    state->config_file  = config_file ? config_file : "";
    state->quiet        = quiet;
    {
        std::time_t t = std::chrono::system_clock::to_time_t( state->datetime_creation );
        char buf[64];
        std::strftime( buf, sizeof( buf ), "%Y-%m-%d_%H-%M-%S", std::localtime( &t ) );
        state->datetime_creation_string = buf;
    }
    // A config file that cannot be opened means defaults (State.cpp:31-42)
    if( !state->config_file.empty() )
    {
        std::ifstream f( state->config_file );
        if( !f.is_open() )
        {
            Log( Log_Level::Error, Log_Sender::All, "Could not open config file \"" + state->config_file + "\". Using defaults." );
            state->config_file = "";
        }
    }
    config::Log_from_Config( state->config_file, quiet );
    Log( Log_Level::Info, Log_Sender::All, "spirit_b200 (B200-native Spirit hot path), scalar type double" );

    state->active_image = config::Spin_System_from_Config( state->config_file );
    // Random initial configuration from the LLG prng (State.cpp:121)
    configurations::Random( *state->active_image, []( const Vec3 &, const Vec3 & ) { return true; } );

    auto chain             = std::make_shared<Chain>();
    chain->gneb_parameters = config::Parameters_GNEB_from_Config( state->config_file );
    chain->images.push_back( state->active_image );
    chain->noi        = 1;
    chain->image_type = { GNEB_Image_Type::Normal };
    chain->Rx         = { 0.0 };
    chain->Setup_Interpolation();
    state->chain = chain;

    state->idx_active_image = 0;
    state->noi              = 1;
    state->nos              = state->active_image->nos;
    state->method_image.assign( 1, nullptr );
    state->method_chain.reset();
    if( quiet )
    {
        state->active_image->llg_parameters->output_any = false;
        chain->gneb_parameters->output_any              = false;
    }
    return state;
}
catch( ... )
{
    handle_exception_api( "State_Setup" );
    return nullptr;
}

void State_Delete( State * state ) noexcept
try
{
    if( !state )
        return;
    Simulation_Stop_All( state );
    delete state;
}
catch( ... )
{
    handle_exception_api( "State_Delete" );
}

// State.cpp:226-252
void State_Update( State * state ) noexcept
try
{
    if( !state || !state->chain )
        return;
    state->noi = state->chain->noi;
    if( state->idx_active_image >= state->noi )
        state->idx_active_image = state->noi - 1;
    state->chain->idx_active_image = state->idx_active_image;
    state->active_image            = state->chain->images[state->idx_active_image];
    state->nos                     = state->active_image->nos;
    state->method_image.resize( state->noi );
}
catch( ... )
{
    handle_exception_api( "State_Update" );
}

void State_To_Config( State *, const char * config_file, const char * ) noexcept
{
    Log( Log_Level::Warning, Log_Sender::API,
         std::string( "State_To_Config: writing config files is outside the hot path of spirit_b200; \"" )
             + ( config_file ? config_file : "" ) + "\" was not written" );
}

const char * State_DateTime( State * state ) noexcept
{
    return state ? state->datetime_creation_string.c_str() : "";
}

// ---- Version (core/src/Spirit/Version.cpp) ----
const int Spirit_Version_Major() noexcept { return 2; }
const int Spirit_Version_Minor() noexcept { return 2; }
const int Spirit_Version_Patch() noexcept { return 0; }
const char * Spirit_Version() noexcept { return "2.2.0"; }
const char * Spirit_Version_Revision() noexcept { return "spirit_b200"; }
const char * Spirit_Version_Full() noexcept { return "2.2.0 (spirit_b200)"; }
const char * Spirit_Compiler() noexcept { return "nvcc+g++"; }
const char * Spirit_Compiler_Version() noexcept { return __VERSION__; }
const char * Spirit_Compiler_Full() noexcept { return "nvcc 12.9 + g++ " __VERSION__; }
const char * Spirit_Scalar_Type() noexcept { return "double"; }
const char * Spirit_Defects() noexcept { return "ON"; } // (compile-time options of the reference; always built here)
const char * Spirit_Pinning() noexcept { return "ON"; }
const char * Spirit_Cuda() noexcept { return "ON"; }
const char * Spirit_OpenMP() noexcept { return "OFF"; }
int Spirit_OpenMP_Get_Num_Threads() noexcept { return 1; }
const char * Spirit_Threads() noexcept { return "OFF"; }
const char * Spirit_FFTW() noexcept { return "OFF"; }

// ---- Constants (core/src/Spirit/Constants.cpp) ----
scalar Constants_mu_B() noexcept { return constants::mu_B; }
scalar Constants_mu_0() noexcept { return constants::mu_0; }
scalar Constants_k_B() noexcept { return constants::k_B; }
scalar Constants_hbar() noexcept { return constants::hbar; }
scalar Constants_mRy() noexcept { return constants::mRy; }
scalar Constants_gamma() noexcept { return constants::gamma; }
scalar Constants_g_e() noexcept { return constants::g_e; }
scalar Constants_Pi() noexcept { return constants::Pi; }

// ---- Log ----
void Log_Send( State *, Spirit_Log_Level level, Spirit_Log_Sender sender, const char * message, int idx_image, int idx_chain ) noexcept
try
{
    Log( Log_Level( int( level ) ), Log_Sender( int( sender ) ), message ? message : "", idx_image, idx_chain );
}
catch( ... )
{
}
void Log_Set_Output_To_Console( State *, bool output, int level ) noexcept
{
    Log.messages_to_console = output;
    Log.level_console       = Log_Level( std::max( 0, std::min( 6, level ) ) );
}
// Log.h:70-122
void Log_Append( State * ) noexcept
{
    Log.Append_to_File();
}
void Log_Dump( State * ) noexcept
{
    Log.Dump_to_File();
}
void Log_Set_Output_File_Tag( State *, const char * tag ) noexcept
{
    Log.file_tag = tag ? tag : "";
}
void Log_Set_Output_Folder( State *, const char * folder ) noexcept
{
    Log.output_folder = folder ? folder : ".";
}
void Log_Set_Output_To_File( State *, bool output, int level ) noexcept
{
    Log.messages_to_file = output;
    Log.level_file       = Log_Level( std::max( 0, std::min( 6, level ) ) );
}
const char * Log_Get_Output_File_Tag( State * ) noexcept
{
    return Log.file_tag.c_str();
}
const char * Log_Get_Output_Folder( State * ) noexcept
{
    return Log.output_folder.c_str();
}
bool Log_Get_Output_To_Console( State * ) noexcept
{
    return Log.messages_to_console;
}
int Log_Get_Output_Console_Level( State * ) noexcept
{
    return int( Log.level_console );
}
bool Log_Get_Output_To_File( State * ) noexcept
{
    return Log.messages_to_file;
}
int Log_Get_Output_File_Level( State * ) noexcept
{
    return int( Log.level_file );
}
int Log_Get_N_Entries( State * ) noexcept { return Log.n_entries; }
int Log_Get_N_Errors( State * ) noexcept { return Log.n_errors; }
int Log_Get_N_Warnings( State * ) noexcept { return Log.n_warnings; }
