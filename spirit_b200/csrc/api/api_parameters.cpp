// C API: Parameters_LLG_* and Parameters_GNEB_*.
// Reference behaviour: core/src/Spirit/Parameters_LLG.cpp, core/src/Spirit/Parameters_GNEB.cpp.
// Setters take float like the reference; they lock the image / chain. Running methods pick the new values up
// at their next iteration (the kernel parameter block is rebuilt from the parameter struct every iteration).
#include "api_common.hpp"

#include <Spirit/Parameters_GNEB.h>
#include <Spirit/Parameters_LLG.h>

using namespace sb;

// ---------------------------------------------------------------------------------------------
// LLG
// ---------------------------------------------------------------------------------------------
#define LLG_SETTER( NAME, ARGS, BODY )                                                                                 \
    void NAME( State * state, ARGS, int idx_image, int idx_chain ) noexcept                                            \
    try                                                                                                                \
    {                                                                                                                  \
        auto image = resolve( state, idx_image, idx_chain ).image;                                                     \
        ImageLock lock( *image );                                                                                      \
        auto & p = *image->llg_parameters;                                                                             \
        BODY;                                                                                                          \
    }                                                                                                                  \
    SB_API_CATCH_VOID

#define COMMA ,

LLG_SETTER( Parameters_LLG_Set_Output_Tag, const char * tag, p.output_file_tag = tag )
LLG_SETTER( Parameters_LLG_Set_Output_Folder, const char * folder, p.output_folder = folder )
LLG_SETTER( Parameters_LLG_Set_Output_General, bool any COMMA bool initial COMMA bool final_,
            p.output_any = any; p.output_initial = initial; p.output_final = final_ )
LLG_SETTER( Parameters_LLG_Set_Output_Energy,
            bool energy_step COMMA bool energy_archive COMMA bool energy_spin_resolved COMMA bool energy_divide_by_nos COMMA bool energy_add_readability_lines,
            p.output_energy_step = energy_step; p.output_energy_archive = energy_archive;
            p.output_energy_spin_resolved = energy_spin_resolved; p.output_energy_divide_by_nspins = energy_divide_by_nos;
            p.output_energy_add_readability_lines = energy_add_readability_lines )
LLG_SETTER( Parameters_LLG_Set_Output_Configuration, bool configuration_step COMMA bool configuration_archive COMMA int configuration_filetype,
            p.output_configuration_step = configuration_step; p.output_configuration_archive = configuration_archive;
            p.output_vf_filetype = configuration_filetype )
LLG_SETTER( Parameters_LLG_Set_N_Iterations, int n_iterations COMMA int n_iterations_log,
            p.n_iterations = n_iterations; p.n_iterations_log = n_iterations_log )
LLG_SETTER( Parameters_LLG_Set_Direct_Minimization, bool direct, p.direct_minimization = direct )
LLG_SETTER( Parameters_LLG_Set_Convergence, float convergence, p.force_convergence = convergence )
LLG_SETTER( Parameters_LLG_Set_Time_Step, float dt, p.dt = dt )
LLG_SETTER( Parameters_LLG_Set_Damping, float damping, p.damping = damping )
LLG_SETTER( Parameters_LLG_Set_Non_Adiabatic_Damping, float beta, p.beta = beta )
LLG_SETTER( Parameters_LLG_Set_STT, bool use_gradient COMMA float magnitude COMMA const float * normal,
            p.stt_use_gradient = use_gradient; p.stt_magnitude = magnitude;
            p.stt_polarisation_normal = Vec3{ normal[0] COMMA normal[1] COMMA normal[2] };
            if( p.stt_polarisation_normal.norm() < 0.9 ) {
                p.stt_polarisation_normal = Vec3{ 0 COMMA 0 COMMA 1 };
                Log( Log_Level::Warning, Log_Sender::API, "s_c_vec = {0,0,0} replaced by {0,0,1}" );
            } else p.stt_polarisation_normal.normalize() )
LLG_SETTER( Parameters_LLG_Set_Temperature, float T, p.temperature = T )
LLG_SETTER( Parameters_LLG_Set_Temperature_Gradient, float inclination COMMA const float * direction,
            p.temperature_gradient_inclination = inclination;
            p.temperature_gradient_direction = Vec3{ direction[0] COMMA direction[1] COMMA direction[2] } )

#define LLG_GETTER( TYPE, NAME, EXPR, FAIL )                                                                           \
    TYPE NAME( State * state, int idx_image, int idx_chain ) noexcept                                                  \
    try                                                                                                                \
    {                                                                                                                  \
        auto image = resolve( state, idx_image, idx_chain ).image;                                                     \
        auto & p   = *image->llg_parameters;                                                                           \
        return EXPR;                                                                                                   \
    }                                                                                                                  \
    SB_API_CATCH_RET( FAIL )

LLG_GETTER( const char *, Parameters_LLG_Get_Output_Tag, p.output_file_tag.c_str(), nullptr )
LLG_GETTER( const char *, Parameters_LLG_Get_Output_Folder, p.output_folder.c_str(), nullptr )
LLG_GETTER( bool, Parameters_LLG_Get_Direct_Minimization, p.direct_minimization, false )
LLG_GETTER( float, Parameters_LLG_Get_Convergence, float( p.force_convergence ), 0 )
LLG_GETTER( float, Parameters_LLG_Get_Time_Step, float( p.dt ), 0 )
LLG_GETTER( float, Parameters_LLG_Get_Damping, float( p.damping ), 0 )
LLG_GETTER( float, Parameters_LLG_Get_Non_Adiabatic_Damping, float( p.beta ), 0 )
LLG_GETTER( float, Parameters_LLG_Get_Temperature, float( p.temperature ), 0 )

#define LLG_GETTER_VOID( NAME, ARGS, BODY )                                                                            \
    void NAME( State * state, ARGS, int idx_image, int idx_chain ) noexcept                                            \
    try                                                                                                                \
    {                                                                                                                  \
        auto image = resolve( state, idx_image, idx_chain ).image;                                                     \
        auto & p   = *image->llg_parameters;                                                                           \
        BODY;                                                                                                          \
    }                                                                                                                  \
    SB_API_CATCH_VOID

LLG_GETTER_VOID( Parameters_LLG_Get_Output_General, bool * any COMMA bool * initial COMMA bool * final_,
                 *any = p.output_any; *initial = p.output_initial; *final_ = p.output_final )
LLG_GETTER_VOID( Parameters_LLG_Get_Output_Energy,
                 bool * energy_step COMMA bool * energy_archive COMMA bool * energy_spin_resolved COMMA bool * energy_divide_by_nos COMMA bool * energy_add_readability_lines,
                 *energy_step = p.output_energy_step; *energy_archive = p.output_energy_archive;
                 *energy_spin_resolved = p.output_energy_spin_resolved; *energy_divide_by_nos = p.output_energy_divide_by_nspins;
                 *energy_add_readability_lines = p.output_energy_add_readability_lines )
LLG_GETTER_VOID( Parameters_LLG_Get_Output_Configuration, bool * configuration_step COMMA bool * configuration_archive COMMA int * configuration_filetype,
                 *configuration_step = p.output_configuration_step; *configuration_archive = p.output_configuration_archive;
                 *configuration_filetype = p.output_vf_filetype )
LLG_GETTER_VOID( Parameters_LLG_Get_N_Iterations, int * iterations COMMA int * iterations_log,
                 *iterations = int( p.n_iterations ); *iterations_log = int( p.n_iterations_log ) )
LLG_GETTER_VOID( Parameters_LLG_Get_Temperature_Gradient, float * inclination COMMA float * direction,
                 *inclination = float( p.temperature_gradient_inclination );
                 for( int d = 0; d < 3; ++d ) direction[d] = float( p.temperature_gradient_direction[d] ) )
LLG_GETTER_VOID( Parameters_LLG_Get_STT, bool * use_gradient COMMA float * magnitude COMMA float * normal,
                 *use_gradient = p.stt_use_gradient; *magnitude = float( p.stt_magnitude );
                 for( int d = 0; d < 3; ++d ) normal[d] = float( p.stt_polarisation_normal[d] ) )

// ---------------------------------------------------------------------------------------------
// GNEB (chain-wide parameters; the reference keeps one parameter struct per chain)
// ---------------------------------------------------------------------------------------------
namespace
{
struct ChainLock
{
    explicit ChainLock( Chain & c ) : c_( c )
    {
        c_.Lock();
    }
    ~ChainLock()
    {
        c_.Unlock();
    }
    Chain & c_;
};
} // namespace

#define GNEB_SETTER( NAME, ARGS, BODY )                                                                                \
    void NAME( State * state, ARGS, int idx_chain ) noexcept                                                           \
    try                                                                                                                \
    {                                                                                                                  \
        int idx_image = -1;                                                                                            \
        auto chain    = resolve( state, idx_image, idx_chain ).chain;                                                  \
        ChainLock lock( *chain );                                                                                      \
        auto & p = *chain->gneb_parameters;                                                                            \
        BODY;                                                                                                          \
    }                                                                                                                  \
    catch( ... )                                                                                                       \
    {                                                                                                                  \
        handle_exception_api( __func__, -1, idx_chain );                                                               \
    }

GNEB_SETTER( Parameters_GNEB_Set_Output_Tag, const char * tag, p.output_file_tag = tag )
GNEB_SETTER( Parameters_GNEB_Set_Output_Folder, const char * folder, p.output_folder = folder )
GNEB_SETTER( Parameters_GNEB_Set_Output_General, bool any COMMA bool initial COMMA bool final_,
             p.output_any = any; p.output_initial = initial; p.output_final = final_ )
GNEB_SETTER( Parameters_GNEB_Set_Output_Energies, bool step COMMA bool interpolated COMMA bool divide_by_nos COMMA bool add_readability_lines,
             p.output_energies_step = step; p.output_energies_interpolated = interpolated;
             p.output_energies_divide_by_nspins = divide_by_nos; p.output_energies_add_readability_lines = add_readability_lines )
GNEB_SETTER( Parameters_GNEB_Set_Output_Chain, bool step COMMA int filetype, p.output_chain_step = step; p.output_vf_filetype = filetype )
GNEB_SETTER( Parameters_GNEB_Set_N_Iterations, int n_iterations COMMA int n_iterations_log,
             p.n_iterations = n_iterations; p.n_iterations_log = n_iterations_log )
GNEB_SETTER( Parameters_GNEB_Set_Spring_Force_Ratio, float ratio, p.spring_force_ratio = ratio )
GNEB_SETTER( Parameters_GNEB_Set_Path_Shortening_Constant, float path_shortening_constant, p.path_shortening_constant = path_shortening_constant )
GNEB_SETTER( Parameters_GNEB_Set_Moving_Endpoints, bool moving_endpoints, p.moving_endpoints = moving_endpoints )
GNEB_SETTER( Parameters_GNEB_Set_Translating_Endpoints, bool translating_endpoints, p.translating_endpoints = translating_endpoints )
GNEB_SETTER( Parameters_GNEB_Set_Equilibrium_Delta_Rx, float delta_Rx_left COMMA float delta_Rx_right,
             p.equilibrium_delta_Rx_left = delta_Rx_left; p.equilibrium_delta_Rx_right = delta_Rx_right )

void Parameters_GNEB_Set_Convergence( State * state, float convergence, int idx_image, int idx_chain ) noexcept
try
{
    auto chain = resolve( state, idx_image, idx_chain ).chain;
    ChainLock lock( *chain );
    chain->gneb_parameters->force_convergence = convergence;
}
SB_API_CATCH_VOID

void Parameters_GNEB_Set_Spring_Constant( State * state, float spring_constant, int idx_image, int idx_chain ) noexcept
try
{
    auto chain = resolve( state, idx_image, idx_chain ).chain;
    ChainLock lock( *chain );
    chain->gneb_parameters->spring_constant = spring_constant;
}
SB_API_CATCH_VOID

void Parameters_GNEB_Set_Climbing_Falling( State * state, int image_type, int idx_image, int idx_chain ) noexcept
try
{
    auto chain = resolve( state, idx_image, idx_chain ).chain;
    ChainLock lock( *chain );
    chain->image_type[idx_image] = GNEB_Image_Type( image_type );
}
SB_API_CATCH_VOID

// Parameters_GNEB.cpp:338-367: maxima climb, minima fall, using the images' current energies
void Parameters_GNEB_Set_Image_Type_Automatically( State * state, int idx_chain ) noexcept
try
{
    int idx_image = -1;
    auto chain    = resolve( state, idx_image, idx_chain ).chain;
    for( int img = 1; img < chain->noi - 1; ++img )
    {
        const double E0 = chain->images[img - 1]->E, E1 = chain->images[img]->E, E2 = chain->images[img + 1]->E;
        if( E0 < E1 && E1 > E2 )
            chain->image_type[img] = GNEB_Image_Type::Climbing;
        else if( E0 > E1 && E1 < E2 )
            chain->image_type[img] = GNEB_Image_Type::Falling;
        else if( chain->image_type[img] != GNEB_Image_Type::Stationary )
            chain->image_type[img] = GNEB_Image_Type::Normal;
    }
}
catch( ... )
{
    handle_exception_api( __func__, -1, idx_chain );
}

// Parameters_GNEB.cpp:369-395: also resizes the interpolation arrays
void Parameters_GNEB_Set_N_Energy_Interpolations( State * state, int n, int idx_chain ) noexcept
try
{
    int idx_image = -1;
    auto chain    = resolve( state, idx_image, idx_chain ).chain;
    ChainLock lock( *chain );
    chain->gneb_parameters->n_E_interpolations = n;
    const int size_interpolated                = chain->noi + ( chain->noi - 1 ) * n;
    chain->Rx_interpolated.assign( size_interpolated, 0.0 );
    chain->E_interpolated.assign( size_interpolated, 0.0 );
    chain->E_array_interpolated.assign( 7, std::vector<double>( size_interpolated, 0.0 ) );
}
catch( ... )
{
    handle_exception_api( __func__, -1, idx_chain );
}

#define GNEB_GETTER( TYPE, NAME, EXPR, FAIL )                                                                          \
    TYPE NAME( State * state, int idx_chain ) noexcept                                                                 \
    try                                                                                                                \
    {                                                                                                                  \
        int idx_image = -1;                                                                                            \
        auto chain    = resolve( state, idx_image, idx_chain ).chain;                                                  \
        auto & p      = *chain->gneb_parameters;                                                                       \
        return EXPR;                                                                                                   \
    }                                                                                                                  \
    catch( ... )                                                                                                       \
    {                                                                                                                  \
        handle_exception_api( __func__, -1, idx_chain );                                                               \
        return FAIL;                                                                                                   \
    }

GNEB_GETTER( const char *, Parameters_GNEB_Get_Output_Tag, p.output_file_tag.c_str(), nullptr )
GNEB_GETTER( const char *, Parameters_GNEB_Get_Output_Folder, p.output_folder.c_str(), nullptr )
GNEB_GETTER( float, Parameters_GNEB_Get_Spring_Force_Ratio, float( p.spring_force_ratio ), 0 )
GNEB_GETTER( float, Parameters_GNEB_Get_Path_Shortening_Constant, float( p.path_shortening_constant ), 0 )
GNEB_GETTER( bool, Parameters_GNEB_Get_Moving_Endpoints, p.moving_endpoints, false )
GNEB_GETTER( bool, Parameters_GNEB_Get_Translating_Endpoints, p.translating_endpoints, false )
GNEB_GETTER( int, Parameters_GNEB_Get_N_Energy_Interpolations, p.n_E_interpolations, 0 )

#define GNEB_GETTER_VOID( NAME, ARGS, BODY )                                                                           \
    void NAME( State * state, ARGS, int idx_chain ) noexcept                                                           \
    try                                                                                                                \
    {                                                                                                                  \
        int idx_image = -1;                                                                                            \
        auto chain    = resolve( state, idx_image, idx_chain ).chain;                                                  \
        auto & p      = *chain->gneb_parameters;                                                                       \
        BODY;                                                                                                          \
    }                                                                                                                  \
    catch( ... )                                                                                                       \
    {                                                                                                                  \
        handle_exception_api( __func__, -1, idx_chain );                                                               \
    }

GNEB_GETTER_VOID( Parameters_GNEB_Get_Output_General, bool * any COMMA bool * initial COMMA bool * final_,
                  *any = p.output_any; *initial = p.output_initial; *final_ = p.output_final )
GNEB_GETTER_VOID( Parameters_GNEB_Get_Output_Energies, bool * step COMMA bool * interpolated COMMA bool * divide_by_nos COMMA bool * add_readability_lines,
                  *step = p.output_energies_step; *interpolated = p.output_energies_interpolated;
                  *divide_by_nos = p.output_energies_divide_by_nspins; *add_readability_lines = p.output_energies_add_readability_lines )
GNEB_GETTER_VOID( Parameters_GNEB_Get_Output_Chain, bool * step COMMA int * filetype, *step = p.output_chain_step; *filetype = p.output_vf_filetype )
GNEB_GETTER_VOID( Parameters_GNEB_Get_N_Iterations, int * iterations COMMA int * iterations_log,
                  *iterations = int( p.n_iterations ); *iterations_log = int( p.n_iterations_log ) )
GNEB_GETTER_VOID( Parameters_GNEB_Get_Equilibrium_Delta_Rx, float * delta_Rx_left COMMA float * delta_Rx_right,
                  *delta_Rx_left = float( p.equilibrium_delta_Rx_left ); *delta_Rx_right = float( p.equilibrium_delta_Rx_right ) )

float Parameters_GNEB_Get_Convergence( State * state, int idx_image, int idx_chain ) noexcept
try
{
    return float( resolve( state, idx_image, idx_chain ).chain->gneb_parameters->force_convergence );
}
SB_API_CATCH_RET( 0 )

float Parameters_GNEB_Get_Spring_Constant( State * state, int idx_image, int idx_chain ) noexcept
try
{
    return float( resolve( state, idx_image, idx_chain ).chain->gneb_parameters->spring_constant );
}
SB_API_CATCH_RET( 0 )

int Parameters_GNEB_Get_Climbing_Falling( State * state, int idx_image, int idx_chain ) noexcept
try
{
    auto chain = resolve( state, idx_image, idx_chain ).chain;
    return int( chain->image_type[idx_image] );
}
SB_API_CATCH_RET( 0 )
